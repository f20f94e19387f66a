// Package ykgpu is the reference-side binding of libykgpu.so: a type that satisfies
// github.com/apache/yunikorn-scheduler-interface/lib/go/api.SchedulerAPI (method set:
// pkg/common/test/schedulerapi_mock.go:88-146) and drives api.ResourceManagerCallback
// (pkg/cache/scheduler_callback.go:42-43), so that pkg/shim is handed this instead of serviceContext.RMProxy
// (pkg/cmd/shim/main.go:56-57) with pkg/shim, pkg/plugin and cache.Context unchanged.
//
// NOT COMPILED IN THIS REPOSITORY: the build image has no Go toolchain and neither the SI nor the Kubernetes modules are
// vendored (SURVEY.md section 0).  The file is complete -- every SchedulerAPI call, the scheduling loop, the snapshot
// builder (strings -> bit sets through the dictionary encoder that ships in the same library), foreign-pod occupancy,
// the slow-path bridge to the unchanged Go PredicateManager and the scheduling-state notifications -- and is written to
// be reviewed next to include/ykgpu.h, include/ykgpu_dict.h and INTEGRATION.md.  Everything with scheduling behaviour
// lives behind the C ABI, where it is tested (tests/ drive the same entry points through ctypes).
//
// Concurrency: one mutex guards the index maps and every cgo call; callbacks into the shim are made WITHOUT it (the
// shim's handlers call back into SchedulerAPI, e.g. task.go:518 releases from inside UpdateAllocation handling).
package ykgpu

/*
#cgo CFLAGS: -I${SRCDIR}/../../include
#cgo LDFLAGS: -L${SRCDIR}/../../yunikorn_k8shim_b200 -lykgpu
#include <stdlib.h>
#include <string.h>
#include "ykgpu.h"
#include "ykgpu_dict.h"

// cgo cannot build arrays of C structs that contain pointers to Go memory: the spec is assembled in C memory
static yk_requirement* yk_go_reqs(uint32_t n) { return (yk_requirement*)calloc(n ? n : 1, sizeof(yk_requirement)); }
static yk_selector_term* yk_go_terms(uint32_t n) { return (yk_selector_term*)calloc(n ? n : 1, sizeof(yk_selector_term)); }
static yk_toleration* yk_go_tols(uint32_t n) { return (yk_toleration*)calloc(n ? n : 1, sizeof(yk_toleration)); }
static yk_taint* yk_go_taints(uint32_t n) { return (yk_taint*)calloc(n ? n : 1, sizeof(yk_taint)); }
static const char** yk_go_strs(uint32_t n) { return (const char**)calloc(n ? n : 1, sizeof(char*)); }
*/
import "C"

import (
	"errors"
	"fmt"
	"sort"
	"strconv"
	"strings"
	"sync"
	"time"
	"unsafe"

	"go.yaml.in/yaml/v3"
	v1 "k8s.io/api/core/v1"

	"github.com/apache/yunikorn-scheduler-interface/lib/go/api"
	siCommon "github.com/apache/yunikorn-scheduler-interface/lib/go/common"
	"github.com/apache/yunikorn-scheduler-interface/lib/go/si"
)

// resource dimension order of the engine (yk_config.D = 4): what pkg/common/resource.go puts into si.Resource
var dims = []string{siCommon.CPU, siCommon.Memory, "pods", "ephemeral-storage"}

const none = ^uint32(0)

// ObjectSource is what the adapter reads from the shim's cache to build the bit sets: the pod behind an AllocationKey
// (= pod UID, pkg/common/si_helper.go:75-115) and the node object behind a NodeID.  *external.SchedulerCache satisfies
// it through a two-line wrapper (INTEGRATION.md): GetPod(uid) and GetNode(name).Node().
type ObjectSource interface {
	GetPod(uid string) *v1.Pod
	GetNodeObject(name string) *v1.Node
}

type nodeRec struct {
	name          string
	total         [4]int64
	occupied      [4]int64 // foreign pods (context.go:409-472): part of total the core never sees as available
	schedulable   bool
	present       bool
	rank          uint32
	label, taint  uint64
}

type askRec struct {
	key, app  string
	req       [4]int64
	foreign   bool   // an occupancy record, not a schedulable ask
	node      uint32 // foreign / recovered: the node it sits on
	lastState uint8  // last state reported through UpdateContainerSchedulingState
	slow      bool
	present   bool
}

type Engine struct {
	mu    sync.Mutex
	h     *C.yk_engine
	dict  *C.yk_dict
	gen   C.uint64_t // dictionary generation the node / ask bit sets were written under
	cb    api.ResourceManagerCallback
	src   ObjectSource
	rmID  string

	nodeIdx  map[string]uint32
	nodes    []nodeRec
	askIdx   map[string]uint32
	asks     []askRec
	freeAsks []uint32
	appIdx   map[string]uint32
	appName  []string
	appQueue []uint32
	queueIdx map[string]uint32 // "root.a.b" -> index
	queueLeaf []bool
	seq      int64
	stop     chan struct{}
	kick     chan struct{}
}

var _ api.SchedulerAPI = &Engine{}

// New creates the engine.  There is no CPU fallback: without a CUDA device yk_create fails and so does New.
func New(maxNodes, maxAsks, maxApps, maxQueues int, src ObjectSource) (*Engine, error) {
	var cfg C.yk_config
	cfg.abi_version = C.YK_ABI_VERSION
	cfg.D = C.uint32_t(len(dims))
	cfg.policy = C.YK_POLICY_FAIR
	cfg.weights[0], cfg.weights[1] = 1, 1 // core default node-sort weights: vcore = memory = 1 (SURVEY A.3)
	cfg.max_nodes, cfg.max_asks, cfg.max_apps, cfg.max_queues = C.uint32_t(maxNodes), C.uint32_t(maxAsks), C.uint32_t(maxApps), C.uint32_t(maxQueues)
	cfg.device = -1
	e := &Engine{src: src, nodeIdx: map[string]uint32{}, askIdx: map[string]uint32{}, appIdx: map[string]uint32{},
		queueIdx: map[string]uint32{}, stop: make(chan struct{}), kick: make(chan struct{}, 1)}
	if rc := C.yk_create(&cfg, &e.h); rc != C.YK_OK {
		return nil, errors.New(C.GoString(C.yk_strerror(rc)))
	}
	e.dict = C.yk_dict_create()
	e.gen = C.yk_dict_generation(e.dict)
	// default tree root -> root.default (deployments/scheduler/yunikorn-configs.yaml:23-32) until the config arrives
	if err := e.setQueues(defaultQueues()); err != nil {
		return nil, err
	}
	return e, nil
}

// SetObjectSource hands over the shim's cache once the shim has built it (the cache does not exist yet when New runs:
// pkg/cmd/shim/main.go creates the SchedulerAPI first, then the shim around it).
func (e *Engine) SetObjectSource(src ObjectSource) {
	e.mu.Lock()
	defer e.mu.Unlock()
	e.src = src
}

func (e *Engine) ck(rc C.int, what string) error {
	if rc == C.YK_OK {
		return nil
	}
	return fmt.Errorf("ykgpu %s: %s (%s)", what, C.GoString(C.yk_strerror(rc)), C.GoString(C.yk_last_error(e.h)))
}

// ---------------------------------------------------------------------------------------------------------------
// queues.yaml (the core's configs.SchedulerConfig, schema evidence: test/e2e/framework/helpers/common/
// test_queues_configs.go:42-60) -> yk_queues_set / yk_queues_priority
// ---------------------------------------------------------------------------------------------------------------

type queueConf struct {
	Name       string            `yaml:"name"`
	Parent     bool              `yaml:"parent"`
	Resources  struct{ Guaranteed, Max map[string]string } `yaml:"resources"`
	Properties map[string]string `yaml:"properties"`
	Queues     []queueConf       `yaml:"queues"`
}
type schedulerConf struct {
	Partitions []struct {
		Name            string      `yaml:"name"`
		Queues          []queueConf `yaml:"queues"`
		NodeSortPolicy  struct{ Type string `yaml:"type"` } `yaml:"nodesortpolicy"`
	} `yaml:"partitions"`
}

func defaultQueues() []queueConf {
	return []queueConf{{Name: "root", Parent: true, Queues: []queueConf{{Name: "default"}}}}
}

// quantity: the core's resources.NewResourceFromConf -- plain integers, "vcore" in milli units when suffixed with m,
// memory with the binary / decimal suffixes of resource.Quantity.  The library parses Kubernetes quantities
// (include/ykgpu_pod.h yk_quantity_value); the configuration only uses the subset below.
func quantity(name, s string) int64 {
	s = strings.TrimSpace(s)
	mult := int64(1)
	for suf, m := range map[string]int64{"Ki": 1 << 10, "Mi": 1 << 20, "Gi": 1 << 30, "Ti": 1 << 40, "k": 1e3, "M": 1e6, "G": 1e9, "T": 1e12} {
		if strings.HasSuffix(s, suf) {
			s, mult = strings.TrimSuffix(s, suf), m
			break
		}
	}
	if name == siCommon.CPU {
		if strings.HasSuffix(s, "m") {
			v, _ := strconv.ParseInt(strings.TrimSuffix(s, "m"), 10, 64)
			return v
		}
		v, _ := strconv.ParseInt(s, 10, 64)
		return v * 1000
	}
	v, _ := strconv.ParseInt(s, 10, 64)
	return v * mult
}

func resVec(m map[string]string) [4]int64 {
	v := [4]int64{-1, -1, -1, -1} // -1 = not set
	for k, name := range dims {
		if s, ok := m[name]; ok {
			v[k] = quantity(name, s)
		}
	}
	return v
}

func (e *Engine) setQueues(root []queueConf) error {
	var parent []C.uint32_t
	var guar, max [][4]int64
	var sortPol []C.uint8_t
	var off []C.int32_t
	var fence []C.uint8_t
	idx := map[string]uint32{}
	var leaf []bool
	var walk func(q queueConf, p uint32, path string)
	walk = func(q queueConf, p uint32, path string) {
		me := uint32(len(parent))
		parent = append(parent, C.uint32_t(p))
		guar, max = append(guar, resVec(q.Resources.Guaranteed)), append(max, resVec(q.Resources.Max))
		s := C.uint8_t(C.YK_SORT_FIFO)
		if q.Properties["application.sort.policy"] == "fair" {
			s = C.YK_SORT_FAIR
		}
		sortPol = append(sortPol, s)
		o, _ := strconv.ParseInt(q.Properties["priority.offset"], 10, 32)
		off = append(off, C.int32_t(o))
		f := C.uint8_t(0)
		if q.Properties["priority.policy"] == "fence" {
			f = 1
		}
		fence = append(fence, f)
		idx[path] = me
		leaf = append(leaf, len(q.Queues) == 0 && !q.Parent)
		for _, c := range q.Queues {
			walk(c, me, path+"."+c.Name)
		}
	}
	for _, q := range root {
		walk(q, none, q.Name)
	}
	n := len(parent)
	g, m := make([]C.int64_t, 4*n), make([]C.int64_t, 4*n) // [D][q] column-major
	for i := 0; i < n; i++ {
		for k := 0; k < 4; k++ {
			g[k*n+i], m[k*n+i] = C.int64_t(guar[i][k]), C.int64_t(max[i][k])
		}
	}
	// allocated == NULL: the library sums what the present applications hold up the new tree (yk_queues_set)
	if err := e.ck(C.yk_queues_set(e.h, C.uint32_t(n), &parent[0], &g[0], &m[0], nil, &sortPol[0]), "yk_queues_set"); err != nil {
		return err
	}
	if err := e.ck(C.yk_queues_priority(e.h, C.uint32_t(n), &off[0], &fence[0]), "yk_queues_priority"); err != nil {
		return err
	}
	e.queueIdx, e.queueLeaf = idx, leaf
	return nil
}

func (e *Engine) applyConfig(cfg string) error {
	if strings.TrimSpace(cfg) == "" {
		return nil
	}
	var sc schedulerConf
	if err := yaml.Unmarshal([]byte(cfg), &sc); err != nil {
		return err
	}
	if len(sc.Partitions) == 0 {
		return nil
	}
	// one engine = one partition (a second partition is a second Engine on another GPU: DESIGN.md "multi-GPU")
	return e.setQueues(sc.Partitions[0].Queues)
}

// RegisterResourceManager: pkg/shim/scheduler.go:147-167 hands over the callback object and the queues.yaml text
func (e *Engine) RegisterResourceManager(req *si.RegisterResourceManagerRequest, cb api.ResourceManagerCallback) (*si.RegisterResourceManagerResponse, error) {
	e.mu.Lock()
	defer e.mu.Unlock()
	e.cb, e.rmID = cb, req.RmID
	if err := e.applyConfig(req.Config); err != nil {
		return nil, err
	}
	go e.loop()
	return &si.RegisterResourceManagerResponse{}, nil
}

// UpdateConfiguration: a changed queues.yaml.  yk_queues_set refuses a tree that drops a queue applications still sit
// in (YK_ERR_STATE): the error goes back to the shim, which keeps the old configuration (as the core does).
func (e *Engine) UpdateConfiguration(req *si.UpdateConfigurationRequest) error {
	e.mu.Lock()
	defer e.mu.Unlock()
	return e.applyConfig(req.Config)
}

// ---------------------------------------------------------------------------------------------------------------
// nodes
// ---------------------------------------------------------------------------------------------------------------

func vec(r *si.Resource) [4]int64 {
	var v [4]int64
	if r != nil {
		for k, name := range dims {
			if q, ok := r.Resources[name]; ok {
				v[k] = q.Value
			}
		}
	}
	return v
}

// nodeBits: labels / taints / spec.unschedulable of the node object -> the dictionary's bit sets
func (e *Engine) nodeBits(idx uint32, name string) (label, taint uint64) {
	var obj *v1.Node
	if e.src != nil {
		obj = e.src.GetNodeObject(name)
	}
	cName := C.CString(name)
	defer C.free(unsafe.Pointer(cName))
	var keys, vals **C.char
	var taints *C.yk_taint
	nl, nt := 0, 0
	unsched := C.int32_t(0)
	var held []unsafe.Pointer
	hold := func(s string) *C.char { p := C.CString(s); held = append(held, unsafe.Pointer(p)); return p }
	defer func() {
		for _, p := range held {
			C.free(p)
		}
	}()
	if obj != nil {
		nl = len(obj.Labels)
		keys, vals = C.yk_go_strs(C.uint32_t(nl)), C.yk_go_strs(C.uint32_t(nl))
		defer C.free(unsafe.Pointer(keys))
		defer C.free(unsafe.Pointer(vals))
		ks := unsafe.Slice(keys, nl+1)
		vs := unsafe.Slice(vals, nl+1)
		i := 0
		for k, v := range obj.Labels {
			ks[i], vs[i] = hold(k), hold(v)
			i++
		}
		nt = len(obj.Spec.Taints)
		taints = C.yk_go_taints(C.uint32_t(nt))
		defer C.free(unsafe.Pointer(taints))
		ts := unsafe.Slice(taints, nt+1)
		for i, t := range obj.Spec.Taints {
			ts[i].key, ts[i].value, ts[i].effect = hold(t.Key), hold(t.Value), effectOf(t.Effect)
		}
		if obj.Spec.Unschedulable {
			unsched = 1
		}
	}
	var lb, tb C.uint64_t
	C.yk_dict_node(e.dict, C.uint32_t(idx), cName, C.uint32_t(nl), keys, vals, C.uint32_t(nt), taints, unsched, &lb, &tb)
	return uint64(lb), uint64(tb)
}

func effectOf(ef v1.TaintEffect) C.uint32_t {
	switch ef {
	case v1.TaintEffectNoSchedule:
		return C.YK_EFFECT_NO_SCHEDULE
	case v1.TaintEffectPreferNoSchedule:
		return C.YK_EFFECT_PREFER_NO_SCHEDULE
	case v1.TaintEffectNoExecute:
		return C.YK_EFFECT_NO_EXECUTE
	}
	return C.YK_EFFECT_ALL
}

// pushNode writes one node to the engine.  avail = total - occupied - (what the engine itself has allocated there):
// the engine's current availability moves by the change of (total - occupied), it is never reset to total.
func (e *Engine) pushNode(idx uint32, n *nodeRec, availDelta [4]int64, isNew bool) error {
	i := C.uint32_t(idx)
	var total, avail [4]C.int64_t
	if isNew {
		for k := 0; k < 4; k++ {
			total[k], avail[k] = C.int64_t(n.total[k]), C.int64_t(n.total[k]-n.occupied[k])
		}
	} else {
		var cur [4]C.int64_t
		if err := e.ck(C.yk_nodes_available(e.h, 1, &i, &cur[0]), "yk_nodes_available"); err != nil {
			return err
		}
		for k := 0; k < 4; k++ {
			total[k], avail[k] = C.int64_t(n.total[k]), cur[k]+C.int64_t(availDelta[k])
		}
	}
	flags := C.uint32_t(0)
	if n.schedulable {
		flags = C.YK_NODE_SCHEDULABLE
	}
	taint, label, rank := C.uint64_t(n.taint), C.uint64_t(n.label), C.uint32_t(n.rank)
	return e.ck(C.yk_nodes_upsert(e.h, 1, &i, &total[0], &avail[0], &taint, &label, &rank, &flags), "yk_nodes_upsert")
}

// refreshRanks: name_rank must preserve the Go string order of the NodeIDs (tie-break of the node iterator, SURVEY
// A.3).  Ranks are re-dealt with gaps when a new name does not fit between its neighbours; only nodes whose rank
// changed are written again.
func (e *Engine) refreshRanks() error {
	order := make([]int, 0, len(e.nodes))
	for i := range e.nodes {
		if e.nodes[i].present {
			order = append(order, i)
		}
	}
	sort.Slice(order, func(a, b int) bool { return e.nodes[order[a]].name < e.nodes[order[b]].name })
	for pos, i := range order {
		want := uint32(pos+1) << 8 // room for 255 insertions between neighbours before the next full re-deal
		n := &e.nodes[i]
		lo, hi := uint32(0), ^uint32(0)
		if pos > 0 {
			lo = e.nodes[order[pos-1]].rank
		}
		if pos+1 < len(order) && e.nodes[order[pos+1]].rank != 0 {
			hi = e.nodes[order[pos+1]].rank
		}
		if n.rank > lo && n.rank < hi && n.rank != 0 {
			continue // still in order: keep
		}
		if lo+1 < hi && n.rank == 0 && hi != ^uint32(0) {
			want = lo + (hi-lo)/2
		}
		n.rank = want
		if err := e.pushNode(uint32(i), n, [4]int64{}, false); err != nil {
			return err
		}
	}
	return nil
}

// UpdateNode: call sites pkg/cache/context.go:256 (drain), :1610,:1630,:1635 (register / update), :1656 (remove)
func (e *Engine) UpdateNode(req *si.NodeRequest) error {
	e.mu.Lock()
	accepted := make([]*si.AcceptedNode, 0, len(req.Nodes))
	rejected := make([]*si.RejectedNode, 0)
	needRanks := false
	for _, n := range req.Nodes {
		idx, known := e.nodeIdx[n.NodeID]
		switch n.Action {
		case si.NodeInfo_DECOMISSION:
			if known {
				i := C.uint32_t(idx)
				C.yk_nodes_remove(e.h, 1, &i)
				C.yk_dict_node_remove(e.dict, i)
				e.nodes[idx].present = false
				delete(e.nodeIdx, n.NodeID)
			}
			continue
		case si.NodeInfo_DRAIN_NODE, si.NodeInfo_DRAIN_TO_SCHEDULABLE:
			// cordon / uncordon carry no resources (context.go:247-257): a flag-only update
			if !known {
				rejected = append(rejected, &si.RejectedNode{NodeID: n.NodeID, Reason: "unknown node"})
				continue
			}
			rec := &e.nodes[idx]
			rec.schedulable = n.Action == si.NodeInfo_DRAIN_TO_SCHEDULABLE
			rec.label, rec.taint = e.nodeBits(idx, rec.name) // spec.unschedulable is a taint bit as well
			if err := e.pushNode(idx, rec, [4]int64{}, false); err != nil {
				rejected = append(rejected, &si.RejectedNode{NodeID: n.NodeID, Reason: err.Error()})
			}
			continue
		}
		// CREATE, CREATE_DRAIN, UPDATE
		isNew := !known
		if isNew {
			idx = uint32(len(e.nodes))
			for i := range e.nodes { // reuse the slot of a removed node
				if !e.nodes[i].present {
					idx = uint32(i)
					break
				}
			}
			if int(idx) == len(e.nodes) {
				e.nodes = append(e.nodes, nodeRec{})
			}
			e.nodes[idx] = nodeRec{name: n.NodeID, present: true, schedulable: n.Action != si.NodeInfo_CREATE_DRAIN}
			e.nodeIdx[n.NodeID] = idx
			needRanks = true
		}
		rec := &e.nodes[idx]
		newTotal := vec(n.SchedulableResource)
		var delta [4]int64
		for k := 0; k < 4; k++ {
			delta[k] = newTotal[k] - rec.total[k] // capacity change moves availability by the same amount
		}
		rec.total = newTotal
		rec.label, rec.taint = e.nodeBits(idx, rec.name)
		if err := e.pushNode(idx, rec, delta, isNew); err != nil {
			rejected = append(rejected, &si.RejectedNode{NodeID: n.NodeID, Reason: err.Error()})
			continue
		}
		accepted = append(accepted, &si.AcceptedNode{NodeID: n.NodeID})
	}
	var err error
	if needRanks {
		err = e.refreshRanks()
	}
	cb := e.cb
	e.mu.Unlock()
	// Accepted / Rejected must come from another goroutine: registerNodes waits on a WaitGroup (context.go:1580-1623)
	if cb != nil && (len(accepted) > 0 || len(rejected) > 0) {
		go cb.UpdateNode(&si.NodeResponse{Accepted: accepted, Rejected: rejected}) //nolint:errcheck
	}
	e.wake()
	return err
}

// ---------------------------------------------------------------------------------------------------------------
// applications
// ---------------------------------------------------------------------------------------------------------------

// UpdateApplication: pkg/cache/application.go:423 (New, with the Ugi the shim resolved, :430) and :597 (Remove)
func (e *Engine) UpdateApplication(req *si.ApplicationRequest) error {
	e.mu.Lock()
	acc := make([]*si.AcceptedApplication, 0, len(req.New))
	rej := make([]*si.RejectedApplication, 0)
	for _, a := range req.New {
		q, ok := e.queueIdx[a.QueueName]
		if !ok || !e.queueLeaf[q] {
			// the core's placement rules would create / pick a queue; this adapter takes the queue the shim names
			rej = append(rej, &si.RejectedApplication{ApplicationID: a.ApplicationID, Reason: "queue " + a.QueueName + " is not a leaf of the configured tree"})
			continue
		}
		idx, known := e.appIdx[a.ApplicationID]
		if !known {
			idx = uint32(len(e.appName))
			e.appName = append(e.appName, a.ApplicationID)
			e.appQueue = append(e.appQueue, q)
			e.appIdx[a.ApplicationID] = idx
		}
		i, cq, t := C.uint32_t(idx), C.uint32_t(q), C.int64_t(time.Now().UnixNano()) // submission time: unique (SURVEY A.6)
		if err := e.ck(C.yk_apps_upsert(e.h, 1, &i, &cq, &t), "yk_apps_upsert"); err != nil {
			rej = append(rej, &si.RejectedApplication{ApplicationID: a.ApplicationID, Reason: err.Error()})
			continue
		}
		acc = append(acc, &si.AcceptedApplication{ApplicationID: a.ApplicationID})
	}
	for _, r := range req.Remove {
		if idx, ok := e.appIdx[r.ApplicationID]; ok {
			i := C.uint32_t(idx)
			C.yk_apps_remove(e.h, 1, &i)
			delete(e.appIdx, r.ApplicationID)
		}
	}
	cb := e.cb
	e.mu.Unlock()
	if cb != nil && (len(acc) > 0 || len(rej) > 0) {
		go cb.UpdateApplication(&si.ApplicationResponse{Accepted: acc, Rejected: rej}) //nolint:errcheck
	}
	return nil
}

// ---------------------------------------------------------------------------------------------------------------
// asks, allocations, releases
// ---------------------------------------------------------------------------------------------------------------

// needsGoPredicates: what does not reduce to the bit sets (SURVEY A.4) stays with the unchanged PredicateManager
func needsGoPredicates(p *v1.Pod) bool {
	if p == nil {
		return true // nothing known about the pod: let the Go path decide
	}
	for _, c := range p.Spec.Containers {
		for _, port := range c.Ports {
			if port.HostPort != 0 {
				return true
			}
		}
	}
	for _, v := range p.Spec.Volumes {
		if v.PersistentVolumeClaim != nil || v.Ephemeral != nil {
			return true
		}
	}
	if a := p.Spec.Affinity; a != nil && (a.PodAffinity != nil || a.PodAntiAffinity != nil) {
		return true
	}
	return len(p.Spec.TopologySpreadConstraints) > 0
}

// podMasks: nodeSelector + required node affinity + tolerations + spec.nodeName -> yk_pod_masks
func (e *Engine) podMasks(p *v1.Pod) C.yk_pod_masks {
	var out C.yk_pod_masks
	out.required_node = C.YK_NONE
	if p == nil {
		out.flags = C.YK_ASK_SLOWPATH
		return out
	}
	var held []unsafe.Pointer
	hold := func(s string) *C.char { c := C.CString(s); held = append(held, unsafe.Pointer(c)); return c }
	keep := func(p unsafe.Pointer) { held = append(held, p) }
	defer func() {
		for _, x := range held {
			C.free(x)
		}
	}()
	var spec C.yk_pod_spec
	ns := len(p.Spec.NodeSelector)
	sk, sv := C.yk_go_strs(C.uint32_t(ns)), C.yk_go_strs(C.uint32_t(ns))
	keep(unsafe.Pointer(sk))
	keep(unsafe.Pointer(sv))
	i := 0
	for k, v := range p.Spec.NodeSelector {
		unsafe.Slice(sk, ns+1)[i], unsafe.Slice(sv, ns+1)[i] = hold(k), hold(v)
		i++
	}
	spec.n_selector, spec.selector_keys, spec.selector_values = C.uint32_t(ns), sk, sv
	reqs := func(in []v1.NodeSelectorRequirement) (*C.yk_requirement, C.uint32_t) {
		r := C.yk_go_reqs(C.uint32_t(len(in)))
		keep(unsafe.Pointer(r))
		rs := unsafe.Slice(r, len(in)+1)
		for i, x := range in {
			vals := C.yk_go_strs(C.uint32_t(len(x.Values)))
			keep(unsafe.Pointer(vals))
			for j, v := range x.Values {
				unsafe.Slice(vals, len(x.Values)+1)[j] = hold(v)
			}
			rs[i].key, rs[i].n_values, rs[i].values = hold(x.Key), C.uint32_t(len(x.Values)), vals
			switch x.Operator {
			case v1.NodeSelectorOpIn:
				rs[i].op = C.YK_OP_IN
			case v1.NodeSelectorOpNotIn:
				rs[i].op = C.YK_OP_NOT_IN
			case v1.NodeSelectorOpExists:
				rs[i].op = C.YK_OP_EXISTS
			case v1.NodeSelectorOpDoesNotExist:
				rs[i].op = C.YK_OP_DOES_NOT_EXIST
			case v1.NodeSelectorOpGt:
				rs[i].op = C.YK_OP_GT
			case v1.NodeSelectorOpLt:
				rs[i].op = C.YK_OP_LT
			}
		}
		return r, C.uint32_t(len(in))
	}
	if a := p.Spec.Affinity; a != nil && a.NodeAffinity != nil && a.NodeAffinity.RequiredDuringSchedulingIgnoredDuringExecution != nil {
		spec.has_required_affinity = 1
		in := a.NodeAffinity.RequiredDuringSchedulingIgnoredDuringExecution.NodeSelectorTerms
		terms := C.yk_go_terms(C.uint32_t(len(in)))
		keep(unsafe.Pointer(terms))
		for i, t := range in {
			ts := unsafe.Slice(terms, len(in)+1)
			ts[i].expressions, ts[i].n_expressions = reqs(t.MatchExpressions)
			ts[i].fields, ts[i].n_fields = reqs(t.MatchFields)
		}
		spec.n_terms, spec.terms = C.uint32_t(len(in)), terms
	}
	tols := C.yk_go_tols(C.uint32_t(len(p.Spec.Tolerations)))
	keep(unsafe.Pointer(tols))
	for i, t := range p.Spec.Tolerations {
		ts := unsafe.Slice(tols, len(p.Spec.Tolerations)+1)
		ts[i].key, ts[i].value, ts[i].effect = hold(t.Key), hold(t.Value), effectOf(t.Effect)
		if t.Operator == v1.TolerationOpExists {
			ts[i].op = C.YK_TOL_EXISTS
		}
	}
	spec.n_tolerations, spec.tolerations = C.uint32_t(len(p.Spec.Tolerations)), tols
	if p.Spec.NodeName != "" {
		spec.node_name = hold(p.Spec.NodeName)
	}
	C.yk_dict_pod(e.dict, &spec, &out)
	if needsGoPredicates(p) {
		out.flags |= C.YK_ASK_SLOWPATH
	}
	return out
}

// dictionary handed out new bits while compiling a pod: node bit sets (and key-only tolerations) were extended
func (e *Engine) syncDictionary() error {
	g := C.yk_dict_generation(e.dict)
	if g == e.gen {
		return nil
	}
	e.gen = g
	for i := range e.nodes {
		n := &e.nodes[i]
		if !n.present {
			continue
		}
		var lb, tb C.uint64_t
		C.yk_dict_node_bits(e.dict, C.uint32_t(i), &lb, &tb)
		if uint64(lb) != n.label || uint64(tb) != n.taint {
			n.label, n.taint = uint64(lb), uint64(tb)
			if err := e.pushNode(uint32(i), n, [4]int64{}, false); err != nil {
				return err
			}
		}
	}
	// pending asks whose tolerations name a key that just got a bit are re-encoded (cheap: only pending, only on growth)
	for i := range e.asks {
		a := &e.asks[i]
		if a.present && !a.foreign && a.lastState != C.YK_ST_ALLOCATED {
			if err := e.pushAsk(uint32(i), a, nil); err != nil {
				return err
			}
		}
	}
	return nil
}

func (e *Engine) newAskSlot() uint32 {
	if n := len(e.freeAsks); n > 0 {
		idx := e.freeAsks[n-1]
		e.freeAsks = e.freeAsks[:n-1]
		return idx
	}
	e.asks = append(e.asks, askRec{})
	return uint32(len(e.asks) - 1)
}

// pushAsk (re)writes a pending ask; alloc != nil on the first write (priority, creation time)
func (e *Engine) pushAsk(idx uint32, a *askRec, alloc *si.Allocation) error {
	var pod *v1.Pod
	if e.src != nil {
		pod = e.src.GetPod(a.key)
	}
	m := e.podMasks(pod)
	a.slow = m.flags&C.YK_ASK_SLOWPATH != 0
	app, ok := e.appIdx[a.app]
	if !ok {
		return fmt.Errorf("ask %s: unknown application %s", a.key, a.app)
	}
	var rq [4]C.int64_t
	for k := 0; k < 4; k++ {
		rq[k] = C.int64_t(a.req[k])
	}
	prio, seq := C.int32_t(0), C.int64_t(0)
	if alloc != nil {
		created, _ := strconv.ParseInt(alloc.AllocationTags[siCommon.CreationTime], 10, 64)
		e.seq++
		prio, seq = C.int32_t(alloc.Priority), C.int64_t(created<<20|e.seq&0xFFFFF) // seconds tie: break by arrival (SURVEY A.6)
		a.lastState = C.YK_ST_PENDING
	}
	i, capp := C.uint32_t(idx), C.uint32_t(app)
	tol, need, deny, rn, fl := m.tolerated_bits, m.required_bits, m.forbidden_bits, m.required_node, m.flags
	return e.ck(C.yk_asks_upsert(e.h, 1, &i, &rq[0], &tol, &need, &deny, &prio, &seq, &capp, &rn, &fl, nil), "yk_asks_upsert")
}

// occupy: a pod that holds resources on a node without being scheduled by this engine -- a foreign pod
// (context.go:409-472, tagged siCommon.Foreign) or a YuniKorn pod found bound on recovery: the node's availability
// drops by its request; the release gives it back.
func (e *Engine) occupy(a *askRec, sign int64) error {
	n := &e.nodes[a.node]
	var delta [4]int64
	for k := 0; k < 4; k++ {
		n.occupied[k] += sign * a.req[k]
		delta[k] = -sign * a.req[k]
	}
	return e.pushNode(a.node, n, delta, false)
}

// UpdateAllocation: asks are si.Allocation without NodeID (si_helper.go:75-115, sent from task.go:311-334); with a
// NodeID they are existing / foreign allocations; releases come from task.go:518 and context.go:459
func (e *Engine) UpdateAllocation(req *si.AllocationRequest) error {
	e.mu.Lock()
	defer e.mu.Unlock()
	var firstErr error
	note := func(err error) {
		if err != nil && firstErr == nil {
			firstErr = err
		}
	}
	for _, al := range req.Allocations {
		if _, dup := e.askIdx[al.AllocationKey]; dup {
			continue
		}
		rec := askRec{key: al.AllocationKey, app: al.ApplicationID, req: vec(al.ResourcePerAlloc), present: true}
		if al.NodeID != "" {
			nidx, ok := e.nodeIdx[al.NodeID]
			if !ok {
				note(fmt.Errorf("allocation %s on unknown node %s", al.AllocationKey, al.NodeID))
				continue
			}
			rec.foreign, rec.node = true, nidx
			idx := e.newAskSlot()
			e.asks[idx] = rec
			e.askIdx[al.AllocationKey] = idx
			note(e.occupy(&e.asks[idx], +1))
			continue
		}
		idx := e.newAskSlot()
		e.asks[idx] = rec
		e.askIdx[al.AllocationKey] = idx
		note(e.pushAsk(idx, &e.asks[idx], al))
	}
	note(e.syncDictionary())
	if req.Releases != nil {
		for _, r := range req.Releases.AllocationsToRelease {
			idx, ok := e.askIdx[r.AllocationKey]
			if !ok {
				continue
			}
			a := &e.asks[idx]
			switch {
			case a.foreign:
				note(e.occupy(a, -1))
			case a.lastState == C.YK_ST_ALLOCATED:
				i := C.uint32_t(idx)
				note(e.ck(C.yk_release(e.h, 1, &i), "yk_release"))
			default:
				i := C.uint32_t(idx)
				note(e.ck(C.yk_asks_remove(e.h, 1, &i), "yk_asks_remove"))
			}
			a.present = false
			delete(e.askIdx, r.AllocationKey)
			e.freeAsks = append(e.freeAsks, idx)
		}
	}
	e.wake()
	return firstErr
}

// ---------------------------------------------------------------------------------------------------------------
// the scheduling goroutine
// ---------------------------------------------------------------------------------------------------------------

func (e *Engine) wake() {
	select {
	case e.kick <- struct{}{}:
	default:
	}
}

func (e *Engine) Stop() {
	close(e.stop)
	e.mu.Lock()
	defer e.mu.Unlock()
	C.yk_destroy(e.h)
	C.yk_dict_destroy(e.dict)
}

// loop: one yk_cycle per wake-up; what it returns is handed to the shim exactly as the core's
// notifyRMNewAllocation / UpdateContainerSchedulingState do (scheduler_callback.go:49-91, :218-222 consume them).
func (e *Engine) loop() {
	out := make([]C.yk_binding, 1<<16)
	slow := make([]C.uint32_t, 1<<12)
	for {
		select {
		case <-e.stop:
			return
		case <-e.kick:
		case <-time.After(100 * time.Millisecond): // asks that failed are tried again (node changes arrive as kicks)
		}
		e.mu.Lock()
		var n, nslow C.uint32_t
		rc := C.yk_cycle(e.h, C.uint32_t(len(out)), &out[0], &n, &slow[0], C.uint32_t(len(slow)), &nslow)
		resp := &si.AllocationResponse{}
		// a cycle that broke off still returns the bindings it made: they are real, hand them over
		for i := 0; i < int(n); i++ {
			a, node := uint32(out[i].ask), uint32(out[i].node)
			rec := &e.asks[a]
			rec.lastState = C.YK_ST_ALLOCATED
			resp.New = append(resp.New, &si.Allocation{AllocationKey: rec.key, ApplicationID: rec.app, NodeID: e.nodes[node].name,
				ResourcePerAlloc: nil /* the shim only reads key, application and node: scheduler_callback.go:53-76 */})
		}
		var states []*si.UpdateContainerSchedulingStateRequest
		if rc == C.YK_OK {
			states = e.collectStates()
		}
		slowKeys := make([]uint32, 0, int(nslow))
		for i := 0; i < int(nslow); i++ {
			slowKeys = append(slowKeys, uint32(slow[i]))
		}
		cb := e.cb
		e.mu.Unlock()
		if rc != C.YK_OK {
			// the error text is in yk_last_error; the loop goes on with the next kick
			_ = rc
		}
		if cb == nil {
			continue
		}
		if len(resp.New) > 0 {
			_ = cb.UpdateAllocation(resp) // -> AssumePod + bind, one pod at a time (scheduler_callback.go:53-91, task.go:348)
		}
		for _, s := range states {
			cb.UpdateContainerSchedulingState(s)
		}
		e.slowPath(cb, slowKeys)
	}
}

// collectStates: asks that ended the cycle NOFIT / SKIPPED and were not reported in that state yet ->
// UpdateContainerSchedulingState FAILED / SKIPPED (scheduler_callback.go:218-222 -> context.go:1232-1272: the FAILED
// state is what makes the shim mark the pod unschedulable and trigger the autoscaler)
func (e *Engine) collectStates() []*si.UpdateContainerSchedulingStateRequest {
	var idx []C.uint32_t
	for i := range e.asks {
		if a := &e.asks[i]; a.present && !a.foreign && a.lastState != C.YK_ST_ALLOCATED {
			idx = append(idx, C.uint32_t(i))
		}
	}
	if len(idx) == 0 {
		return nil
	}
	st := make([]C.uint8_t, len(idx))
	if C.yk_ask_states(e.h, C.uint32_t(len(idx)), &idx[0], &st[0]) != C.YK_OK {
		return nil
	}
	var out []*si.UpdateContainerSchedulingStateRequest
	for j, i := range idx {
		a := &e.asks[i]
		s := uint8(st[j])
		if s == a.lastState {
			continue
		}
		a.lastState = s
		switch s {
		case C.YK_ST_NOFIT:
			out = append(out, &si.UpdateContainerSchedulingStateRequest{ApplicationID: a.app, AllocationKey: a.key,
				State: si.UpdateContainerSchedulingStateRequest_FAILED, Reason: "no node passes the predicates with enough resources"})
		case C.YK_ST_SKIPPED:
			out = append(out, &si.UpdateContainerSchedulingStateRequest{ApplicationID: a.app, AllocationKey: a.key,
				State: si.UpdateContainerSchedulingStateRequest_SKIPPED, Reason: "request exceeds the queue headroom"})
		}
	}
	return out
}

// slowPath: pods whose predicates do not reduce to the bit sets.  Candidates are tried in the engine's node order
// (ascending float64 score, then NodeID) through the UNCHANGED Go path -- callback.Predicates ->
// Context.IsPodFitNode -> PredicateManager (scheduler_callback.go:196-198) -- and the first node that passes is pinned:
// the ask goes back to the engine naming that node (spec.nodeName semantics), everything else tolerated, so the next
// cycle re-checks the resources, commits and reports the binding like any other.
func (e *Engine) slowPath(cb api.ResourceManagerCallback, askIdxs []uint32) {
	if len(askIdxs) == 0 {
		return
	}
	e.mu.Lock()
	live := make([]uint32, 0, len(e.nodes))
	for i := range e.nodes {
		if e.nodes[i].present && e.nodes[i].schedulable {
			live = append(live, uint32(i))
		}
	}
	ci := make([]C.uint32_t, len(live))
	for i, x := range live {
		ci[i] = C.uint32_t(x)
	}
	scores := make([]C.double, len(live))
	if len(live) > 0 {
		C.yk_node_scores(e.h, C.uint32_t(len(live)), &ci[0], &scores[0])
	}
	sort.Slice(live, func(a, b int) bool {
		if scores[a] != scores[b] {
			return scores[a] < scores[b]
		}
		return e.nodes[live[a]].name < e.nodes[live[b]].name
	})
	names := make([]string, len(live))
	for i, x := range live {
		names[i] = e.nodes[x].name
	}
	type job struct {
		idx uint32
		key string
	}
	jobs := make([]job, 0, len(askIdxs))
	for _, a := range askIdxs {
		if int(a) < len(e.asks) && e.asks[a].present {
			jobs = append(jobs, job{a, e.asks[a].key})
		}
	}
	e.mu.Unlock()
	for _, j := range jobs {
		chosen := ""
		for _, node := range names {
			if cb.Predicates(&si.PredicatesArgs{AllocationKey: j.key, NodeID: node, Allocate: true}) == nil {
				chosen = node
				break
			}
		}
		e.mu.Lock()
		a := &e.asks[j.idx]
		if chosen == "" || !a.present || a.key != j.key {
			e.mu.Unlock()
			continue // no node passes today: the ask stays SLOWPATH and is tried again next cycle
		}
		i, rn, all, zero, fl := C.uint32_t(j.idx), C.uint32_t(e.nodeIdx[chosen]), ^C.uint64_t(0), C.uint64_t(0), C.uint32_t(0)
		var rq [4]C.int64_t
		for k := 0; k < 4; k++ {
			rq[k] = C.int64_t(a.req[k])
		}
		app := C.uint32_t(e.appIdx[a.app])
		prio, seq := C.int32_t(0), C.int64_t(0) // first in its application: it was reached in order already
		C.yk_asks_upsert(e.h, 1, &i, &rq[0], &all, &zero, &zero, &prio, &seq, &app, &rn, &fl, nil)
		e.mu.Unlock()
		e.wake()
	}
}
