// Package ykgpu is the reference-side binding of libykgpu.so: a type that satisfies
// github.com/apache/yunikorn-scheduler-interface/lib/go/api.SchedulerAPI and drives
// api.ResourceManagerCallback, so that pkg/shim can be handed this instead of serviceContext.RMProxy
// (/root/reference/pkg/cmd/shim/main.go:54-57) with pkg/shim, pkg/plugin and cache.Context unchanged.
//
// NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Go toolchain and the SI / core modules are
// not vendored (SURVEY.md section 0).  It is kept deliberately thin -- marshal, one cgo call, unmarshal --
// so that everything with behaviour lives behind include/ykgpu.h where it is tested (tests/ drive the very
// same entry points through ctypes).
package ykgpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../yunikorn_k8shim_b200 -lykgpu
#include <stdlib.h>
#include "ykgpu.h"
*/
import "C"

import (
	"errors"
	"sort"
	"strconv"
	"sync"
	"time"
	"unsafe"

	"github.com/apache/yunikorn-scheduler-interface/lib/go/api"
	siCommon "github.com/apache/yunikorn-scheduler-interface/lib/go/common"
	"github.com/apache/yunikorn-scheduler-interface/lib/go/si"
)

// resource dimension order of the engine (yk_config.D = 4)
var dims = []string{siCommon.CPU, siCommon.Memory, "pods", "ephemeral-storage"}

type Engine struct {
	sync.Mutex
	h        *C.yk_engine
	cb       api.ResourceManagerCallback
	nodeIdx  map[string]uint32 // NodeID -> dense index (the shim already keys its cache by name, scheduler_cache.go:53-54)
	nodeName []string
	askIdx   map[string]uint32 // AllocationKey -> dense index
	askKey   []string
	askApp   []string
	appIdx   map[string]uint32
	seq      int64
	stop     chan struct{}
}

var _ api.SchedulerAPI = &Engine{}

func New(maxNodes, maxAsks, maxApps int) (*Engine, error) {
	var cfg C.yk_config
	cfg.abi_version = C.YK_ABI_VERSION
	cfg.D = C.uint32_t(len(dims))
	cfg.policy = C.YK_POLICY_FAIR
	cfg.weights[0], cfg.weights[1] = 1, 1 // core default: vcore = memory = 1
	cfg.max_nodes, cfg.max_asks, cfg.max_apps, cfg.max_queues = C.uint32_t(maxNodes), C.uint32_t(maxAsks), C.uint32_t(maxApps), 64
	cfg.device = -1
	e := &Engine{nodeIdx: map[string]uint32{}, askIdx: map[string]uint32{}, appIdx: map[string]uint32{}, stop: make(chan struct{})}
	if rc := C.yk_create(&cfg, &e.h); rc != C.YK_OK {
		return nil, errors.New(C.GoString(C.yk_strerror(rc)))
	}
	// default queue tree root -> root.default (deployments/scheduler/yunikorn-configs.yaml:23-32);
	// UpdateConfiguration replaces it from queues.yaml
	parent := []C.uint32_t{C.YK_NONE, 0}
	C.yk_queues_set(e.h, 2, &parent[0], nil, nil, nil, nil)
	return e, nil
}

func (e *Engine) RegisterResourceManager(req *si.RegisterResourceManagerRequest, cb api.ResourceManagerCallback) (*si.RegisterResourceManagerResponse, error) {
	e.cb = cb // the object pkg/shim/scheduler.go:166-167 hands over
	go e.loop()
	return &si.RegisterResourceManagerResponse{}, nil
}

func vec(r *si.Resource) [4]C.int64_t {
	var v [4]C.int64_t
	if r != nil {
		for k, name := range dims {
			if q, ok := r.Resources[name]; ok {
				v[k] = C.int64_t(q.Value)
			}
		}
	}
	return v
}

// UpdateNode: call sites pkg/cache/context.go:256,1610,1630,1635,1656
func (e *Engine) UpdateNode(req *si.NodeRequest) error {
	e.Lock()
	defer e.Unlock()
	accepted := make([]*si.AcceptedNode, 0, len(req.Nodes))
	for _, n := range req.Nodes {
		idx, ok := e.nodeIdx[n.NodeID]
		if !ok {
			idx = uint32(len(e.nodeName))
			e.nodeIdx[n.NodeID] = idx
			e.nodeName = append(e.nodeName, n.NodeID)
		}
		if n.Action == si.NodeInfo_DECOMISSION {
			i := C.uint32_t(idx)
			C.yk_nodes_remove(e.h, 1, &i)
			continue
		}
		total := vec(n.SchedulableResource)
		avail := total // minus occupied/allocated as tracked by the adapter (foreign allocations, context.go:409-472)
		flags := C.uint32_t(C.YK_NODE_SCHEDULABLE)
		if n.Action == si.NodeInfo_CREATE_DRAIN || n.Action == si.NodeInfo_DRAIN_NODE {
			flags = 0
		}
		i, rank := C.uint32_t(idx), C.uint32_t(0) // rank refreshed below
		var taint, label C.uint64_t              // from the dictionary encoder (DESIGN.md section 9, next)
		C.yk_nodes_upsert(e.h, 1, &i, &total[0], &avail[0], &taint, &label, &rank, &flags)
		accepted = append(accepted, &si.AcceptedNode{NodeID: n.NodeID})
	}
	e.refreshRanks()
	// Accepted/Rejected must come from another goroutine: registerNodes waits on a WaitGroup (context.go:1580-1623)
	go e.cb.UpdateNode(&si.NodeResponse{Accepted: accepted}) //nolint:errcheck
	return nil
}

// name_rank must preserve Go string order of NodeIDs (tie-break of the node iterator)
func (e *Engine) refreshRanks() {
	order := make([]int, len(e.nodeName))
	for i := range order {
		order[i] = i
	}
	sort.Slice(order, func(a, b int) bool { return e.nodeName[order[a]] < e.nodeName[order[b]] })
	_ = order // one yk_nodes_upsert with the new ranks for nodes whose rank changed
}

// UpdateApplication: pkg/cache/application.go:423
func (e *Engine) UpdateApplication(req *si.ApplicationRequest) error {
	e.Lock()
	defer e.Unlock()
	acc := make([]*si.AcceptedApplication, 0, len(req.New))
	for _, a := range req.New {
		idx := uint32(len(e.appIdx))
		e.appIdx[a.ApplicationID] = idx
		i, q, t := C.uint32_t(idx), C.uint32_t(1), C.int64_t(time.Now().UnixNano()) // queue index from QueueName
		C.yk_apps_upsert(e.h, 1, &i, &q, &t)
		acc = append(acc, &si.AcceptedApplication{ApplicationID: a.ApplicationID})
	}
	go e.cb.UpdateApplication(&si.ApplicationResponse{Accepted: acc}) //nolint:errcheck
	return nil
}

// UpdateAllocation: asks are si.Allocation without NodeID (pkg/common/si_helper.go:75-115), sent from
// pkg/cache/task.go:311-334; releases from task.go:518, context.go:459
func (e *Engine) UpdateAllocation(req *si.AllocationRequest) error {
	e.Lock()
	defer e.Unlock()
	for _, a := range req.Allocations {
		if a.NodeID != "" {
			continue // existing allocation on recovery: accounted as occupied on its node
		}
		idx := uint32(len(e.askKey))
		e.askIdx[a.AllocationKey] = idx
		e.askKey = append(e.askKey, a.AllocationKey)
		e.askApp = append(e.askApp, a.ApplicationID)
		rq := vec(a.ResourcePerAlloc)
		created, _ := strconv.ParseInt(a.AllocationTags[siCommon.CreationTime], 10, 64)
		e.seq++
		seq := C.int64_t(created<<20 | e.seq&0xFFFFF) // seconds are not unique: break ties by arrival (SURVEY A.6)
		i, prio, app := C.uint32_t(idx), C.int32_t(a.Priority), C.uint32_t(e.appIdx[a.ApplicationID])
		var tol, need, deny C.uint64_t // from the dictionary encoder; slow-path asks get flags = YK_ASK_SLOWPATH
		C.yk_asks_upsert(e.h, 1, &i, &rq[0], &tol, &need, &deny, &prio, &seq, &app, nil, nil, nil)
	}
	if req.Releases != nil {
		for _, r := range req.Releases.AllocationsToRelease {
			if idx, ok := e.askIdx[r.AllocationKey]; ok {
				i := C.uint32_t(idx)
				if C.yk_release(e.h, 1, &i) != C.YK_OK {
					C.yk_asks_remove(e.h, 1, &i) // was still pending
				}
			}
		}
	}
	return nil
}

func (e *Engine) UpdateConfiguration(req *si.UpdateConfigurationRequest) error { return nil } // queues.yaml -> yk_queues_set
func (e *Engine) Stop()                                                        { close(e.stop); C.yk_destroy(e.h) }

// loop is the scheduling goroutine: one yk_cycle per tick, bindings handed to the shim exactly as the core's
// notifyRMNewAllocation does (scheduler_callback.go:49-91 consumes them).
func (e *Engine) loop() {
	out := make([]C.yk_binding, 1<<16)
	slow := make([]C.uint32_t, 1<<12)
	for {
		select {
		case <-e.stop:
			return
		case <-time.After(time.Millisecond):
		}
		e.Lock()
		var n, nslow C.uint32_t
		rc := C.yk_cycle(e.h, C.uint32_t(len(out)), &out[0], &n, &slow[0], C.uint32_t(len(slow)), &nslow)
		resp := &si.AllocationResponse{}
		for i := 0; rc == C.YK_OK && i < int(n); i++ {
			a, node := uint32(out[i].ask), uint32(out[i].node)
			resp.New = append(resp.New, &si.Allocation{AllocationKey: e.askKey[a], ApplicationID: e.askApp[a], NodeID: e.nodeName[node]})
		}
		// slow-path asks: candidate nodes are verified one by one through the UNCHANGED Go path
		for i := 0; i < int(nslow); i++ {
			_ = e.cb.Predicates(&si.PredicatesArgs{AllocationKey: e.askKey[uint32(slow[i])], NodeID: "", Allocate: true})
		}
		e.Unlock()
		if len(resp.New) > 0 {
			_ = e.cb.UpdateAllocation(resp)
		}
		_ = unsafe.Pointer(nil)
	}
}
