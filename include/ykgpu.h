/*
 * ykgpu.h -- C ABI of the B200 scheduling-cycle engine (libykgpu.so).
 *
 * This is the drop-in boundary for YuniKorn's pod->node allocation hot path.  A thin Go type
 * implementing api.SchedulerAPI (method set: /root/reference/pkg/common/test/schedulerapi_mock.go:88-146)
 * binds these entry points through cgo and drives api.ResourceManagerCallback
 * (/root/reference/pkg/cache/scheduler_callback.go:42-43) with what yk_cycle returns; pkg/shim,
 * pkg/plugin and cache.Context stay unchanged.  INTEGRATION.md shows the cgo side.
 *
 * What each entry point replaces in the reference:
 *   yk_nodes_upsert / yk_nodes_remove   SchedulerAPI.UpdateNode(*si.NodeRequest) -- call sites
 *                                       pkg/cache/context.go:256,1610,1630,1635,1656; node vector
 *                                       = common.GetNodeResource, pkg/common/resource.go:188-195
 *   yk_queues_set                       the queues.yaml half of RegisterResourceManager /
 *                                       UpdateConfiguration (pkg/shim/scheduler.go:147-167)
 *   yk_apps_upsert / yk_apps_remove     SchedulerAPI.UpdateApplication (pkg/cache/application.go:423)
 *   yk_asks_upsert / yk_asks_remove     SchedulerAPI.UpdateAllocation(*si.AllocationRequest) with
 *                                       Allocations without NodeID = asks (pkg/cache/task.go:311-334,
 *                                       pkg/common/si_helper.go:75-115); ask vector = common.GetPodResource,
 *                                       pkg/common/resource.go:56-109
 *   yk_release                          AllocationRequest.Releases (pkg/cache/task.go:518, context.go:459)
 *   yk_cycle                            yunikorn-core partition.tryAllocate loop [EXT, SURVEY 3.3] INCLUDING
 *                                       the per-(ask,node) callback ResourceManagerCallback.Predicates
 *                                       (pkg/cache/scheduler_callback.go:196-198 -> pkg/cache/context.go:683-703
 *                                       -> pkg/plugin/predicates/predicate_manager.go:130-283) for the
 *                                       bitmaskable plugin set, and the node sorter/scorer.  Its output is the
 *                                       payload of ResourceManagerCallback.UpdateAllocation(New: ...)
 *                                       (pkg/cache/scheduler_callback.go:49-91).
 *   yk_ask_states                       UpdateContainerSchedulingState FAILED / SKIPPED
 *                                       (pkg/cache/scheduler_callback.go:218-222, context.go:1232-1272)
 *   yk_evaluate                         one ResourceManagerCallback.Predicates(ask,node) answer, for the
 *                                       slow-path bridge and for known-answer tests
 *
 * Conventions: every call returns YK_OK (0) or a negative yk_status; no exceptions or aborts cross the
 * boundary.  Inputs are caller-owned host memory, fully consumed before return (cgo: no Go pointer is
 * retained).  Outputs are caller-allocated.  All resource matrices are COLUMN-MAJOR: element (k, i) of a
 * [D][count] matrix is at m[k*count + i].  Strings never cross: nodes, asks, apps and queues are dense
 * uint32 indices below the capacities given at yk_create (the Go side already owns the name->index maps,
 * pkg/cache/external/scheduler_cache.go:53-54).  One writer at a time per engine (internal mutex).
 * A missing CUDA device or kernel image is an error (YK_ERR_CUDA), never a CPU fallback.
 */
#ifndef YKGPU_H
#define YKGPU_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define YK_ABI_VERSION 3
#define YK_MAX_D 8
#define YK_NONE 0xFFFFFFFFu          /* "no node" / "no gang" index */

typedef enum {
    YK_OK = 0,
    YK_ERR_ARG = -1,        /* null pointer, index out of capacity, bad enum */
    YK_ERR_CUDA = -2,       /* no device, no sm_100a image, launch or copy failed */
    YK_ERR_NOMEM = -3,
    YK_ERR_STATE = -4,      /* e.g. release of an ask that holds no allocation */
    YK_ERR_RANGE = -5,      /* NaN node score (zero total on a weighted resource with non-zero available) */
    YK_ERR_COMM = -6        /* multi-GPU exchange failed */
} yk_status;

/* node sort policy (core NodeSortingPolicy) */
#define YK_POLICY_FAIR 0u
#define YK_POLICY_BINPACKING 1u
/* leaf application sort policy */
#define YK_SORT_FIFO 0u
#define YK_SORT_FAIR 1u
/* node flags */
#define YK_NODE_SCHEDULABLE 1u   /* cleared by DRAIN_NODE / cordon (pkg/cache/context.go:247-257) */
#define YK_NODE_RESERVED 2u      /* hidden from the normal node iterator */
/* ask flags */
#define YK_ASK_SLOWPATH 1u       /* carries a predicate that does not reduce to the masks (ports, volumes,
                                    pod (anti-)affinity, topology spread): never bound by yk_cycle, returned
                                    in its slow-path list for the unchanged Go PredicateManager */
/* per-ask state reported by yk_ask_states */
#define YK_ST_PENDING 0u
#define YK_ST_ALLOCATED 1u
#define YK_ST_NOFIT 2u           /* -> UpdateContainerSchedulingState FAILED */
#define YK_ST_SKIPPED 3u         /* queue headroom; -> SchedulingState SKIPPED */
#define YK_ST_SLOWPATH 4u
#define YK_ST_INVALID 5u         /* request not strictly greater than zero */
#define YK_ST_ABSENT 255u

/* yk_evaluate answers: 0 = fits, else first failing step in the reference's order */
#define YK_FAIL_NODE_NOT_SCHEDULABLE 1
#define YK_FAIL_TOTAL 2
#define YK_FAIL_REQUEST_NOT_POSITIVE 3
#define YK_FAIL_AVAILABLE 4
#define YK_FAIL_NODENAME 5
#define YK_FAIL_TAINT 6
#define YK_FAIL_AFFINITY 7
#define YK_FAIL_RESOURCES 8
#define YK_FAIL_ABSENT 9         /* pod or node not in the cache (context.go:686-694) */

/* yk_config.flags */
#define YK_FLAG_NO_ROW_SHARING 1u   /* sweep one row per ask even when asks of a batch have identical predicate inputs
                                      (requests, tolerations, label masks, node name); default: one row per distinct set */
/* Where the ordered commit runs.  Default (neither flag): automatic -- a cycle whose order consists of long UNIFORM RUNS
   (>= 2048 consecutive asks with one request vector and one predicate signature: the replicas of a deployment, the executors of
   a job, the reference's own benchmark) and that is eligible (fair node sort, non-negative weights and weighted totals, unique
   NodeID ranks, every gang's members requesting one vector) is decided on the device by a grid-wide sort (csrc/yk_uniform.cuh);
   every other cycle takes the sweep + host commit path.  The bindings are the same either way. */
#define YK_FLAG_HOST_COMMIT 2u      /* never commit on the device */
#define YK_FLAG_DEVICE_COMMIT 4u    /* commit on the device whenever the cycle is eligible, also where the order is not made of
                                      uniform runs (yk_lattice_kernel: exact, no bitmap read-back, no host work per ask, but
                                      measured slower than the host commit there) */

typedef struct yk_engine yk_engine;

typedef struct {
    uint32_t abi_version;        /* YK_ABI_VERSION */
    uint32_t D;                  /* resource dimensions 1..YK_MAX_D; by convention 0 vcore(milli) 1 memory 2 pods 3 ephemeral-storage */
    uint32_t policy;             /* YK_POLICY_* */
    uint32_t batch;              /* asks per sweep batch; 0 = default */
    double weights[YK_MAX_D];    /* node-sort resource weights; core default vcore=1, memory=1 */
    uint32_t max_nodes, max_asks, max_apps, max_queues;
    int32_t device;              /* CUDA ordinal, -1 = current device */
    uint32_t flags;              /* YK_FLAG_* (0 = defaults) */
    /* multi-GPU (one process per GPU): asks of every batch are split in `world` contiguous shards, this
       engine sweeps shard `rank`; the caller wires the exchange with yk_set_exchange.  world<=1: single GPU */
    uint32_t rank, world;
} yk_config;

typedef struct { uint32_t ask; uint32_t node; } yk_binding;

typedef struct {
    uint64_t cycles, batches, allocations, nofit, skipped;
    uint64_t evaluations;        /* (ask,node) pairs swept by the kernel */
    uint64_t sweep_launches, other_launches;   /* kernels launched by this library */
    uint64_t h2d_bytes, d2h_bytes;
    double sweep_ms, sort_ms, commit_ms, total_ms;   /* CUDA-event / wall accumulations */
    double last_sweep_ms;
    uint64_t last_sweep_pairs;
    /* host wall-clock split of yk_cycle: 0 table upload + initial device sort, 1 orderer begin_cycle,
       2 orderer fill/rewind, 3 waiting for device results, 4 ordered commit, 5 order merge + state push, 6 launch,
       7 orderer + launch time spent on the helper thread (overlapped with the commit) */
    double host_ms[8];
    /* commit counters: 0 bitmap words scanned, 1 re-scored candidates examined, 2 asks won by a re-scored node,
       3 re-keys of an already re-scored node */
    uint64_t dbg[4];
    /* with YK_PROFILE_COMMIT set: TSC cycles in the commit loop -- 0 clean scan, 1 re-scored walk, 2 choose + undo log
       + erase, 3 subtract + float64 re-score, 4 re-insert + bookkeeping, 5 asks counted */
    uint64_t prof[6];
    /* row sharing: asks that went through a sweep batch, and the rows (distinct predicate signatures) actually swept */
    uint64_t asks_swept, rows_swept;
    /* device-resident ordered commit (yk_lattice_kernel): launches, sub-runs (scan + lattice + chain + patch rounds), asks it
       decided, lattice elements evaluated, sorts, full-order scans (an ask with no candidate near the front), asks decided by
       the capacity bound alone, cycles handed over to the host commit (a gang that has to be rolled back), CUDA-event time */
    uint64_t lattice_launches, lattice_subruns, lattice_asks, lattice_elements, lattice_sorts, lattice_fullscans,
             lattice_quick, lattice_handoffs;
    double lattice_ms;
    uint64_t lattice_cycles;     /* cycles that started on the device commit */
    /* uniform runs of the device commit (consecutive asks with one request vector and one predicate signature, decided by a
       grid-wide sort instead of the sequential chain): runs, asks they decided, elements generated, deeper retries */
    uint64_t uniform_runs, uniform_asks, uniform_elements, uniform_retries;
} yk_stats_t;

int yk_create(const yk_config* cfg, yk_engine** out);
void yk_destroy(yk_engine* e);

/* nodes: total/avail are [D][n]; name_rank = rank of the NodeID in Go string (bytewise) order among all
 * nodes -- it must be order-preserving, not dense.  avail = total - allocated - occupied as the core
 * tracks it. */
int yk_nodes_upsert(yk_engine* e, uint32_t n, const uint32_t* idx, const int64_t* total, const int64_t* avail,
                    const uint64_t* taint_bits, const uint64_t* label_bits, const uint32_t* name_rank,
                    const uint32_t* flags);
int yk_nodes_remove(yk_engine* e, uint32_t n, const uint32_t* idx);

/* queues: whole tree at once.  parent[0] = YK_NONE (root), parent[i] < i.  guaranteed/max/allocated are
 * [D][q] with -1 = not set (allocated: 0).  sort = leaf application sort policy. */
int yk_queues_set(yk_engine* e, uint32_t q, const uint32_t* parent, const int64_t* guaranteed,
                  const int64_t* max, const int64_t* allocated, const uint8_t* sort);

/* optional queue priority properties of the core's queue configuration [EXT yunikorn-core configs: properties
 * priority.offset, priority.policy]: a parent sorts its children by their current priority first (highest priority among the
 * asks pending below the child, plus the child's offset; a child with policy "fence" shows only its offset), then by share.
 * Both arrays are [q] in yk_queues_set order and may be NULL (= 0 / no fence); yk_queues_set resets them.  Behaviour
 * checked by the reference in test/e2e/priority_scheduling/priority_scheduling_test.go:70-251. */
int yk_queues_priority(yk_engine* e, uint32_t q, const int32_t* offset, const uint8_t* fence);

/* applications live in LEAF queues (as in the core: placement into a parent queue is rejected there; here an application
 * whose queue has children is accepted but never scheduled -- the adapter must not produce one) */
int yk_apps_upsert(yk_engine* e, uint32_t n, const uint32_t* idx, const uint32_t* queue,
                   const int64_t* submit_time);
int yk_apps_remove(yk_engine* e, uint32_t n, const uint32_t* idx);

/* User / group resource limits: the core's queue `limits:` entries [EXT yunikorn-core ugm]; the shim sends the user and its
 * groups with every application (si.AddApplicationRequest.Ugi, pkg/cache/application.go:430).  yk_apps_user names the user
 * of each application (dense index, YK_NONE = none).  yk_user_limits_set replaces the whole table: entry l says "user
 * user[l] may hold at most max[.][l] ([D][n], -1 = not limited in that dimension) below queue queue[l]", summed over the
 * user's applications in that queue's subtree; an ask that exceeds what is left is SKIPPED for the cycle exactly like one
 * that exceeds the queue headroom.  The adapter resolves which entry applies to a user (its own, its group's, the
 * wildcard).  held[.][l] = what the user holds there now ([D][n], NULL = 0); the engine keeps it current afterwards
 * (cycle, yk_release).  maxapplications is not modelled (it gates application acceptance, not this path). */
int yk_apps_user(yk_engine* e, uint32_t n, const uint32_t* idx, const uint32_t* user);
int yk_user_limits_set(yk_engine* e, uint32_t n, const uint32_t* queue, const uint32_t* user, const int64_t* max,
                       const int64_t* held);

/* asks: req is [D][a].  create_seq is the total-order key within an application (ties pre-broken by the
 * caller: the reference's CreationTime is second-granular, pkg/common/si_helper.go:109).
 * required_node = YK_NONE or the node index pod.Spec.NodeName names.  gang = YK_NONE or a gang id
 * (all-or-nothing group).  Any of tolerated/required/forbidden/priority/required_node/flags/gang may be
 * NULL = all zero / YK_NONE. */
int yk_asks_upsert(yk_engine* e, uint32_t a, const uint32_t* idx, const int64_t* req,
                   const uint64_t* tolerated_bits, const uint64_t* required_bits, const uint64_t* forbidden_bits,
                   const int32_t* priority, const int64_t* create_seq, const uint32_t* app,
                   const uint32_t* required_node, const uint32_t* flags, const uint32_t* gang);
int yk_asks_remove(yk_engine* e, uint32_t a, const uint32_t* idx);

/* give the resources of allocated asks back to their nodes and queues, and forget the asks */
int yk_release(yk_engine* e, uint32_t n, const uint32_t* ask_idx);

/* One scheduling cycle: runs schedule() passes until no pending ask can be placed or max_bindings is
 * reached.  out[0..*n_out) are the bindings in the reference's commit order.  slow_path_asks (may be NULL)
 * receives up to slow_cap asks flagged YK_ASK_SLOWPATH that were reached in order.
 * max_bindings ends the cycle at the first placement it cannot hold (an ask, or a gang that passed its queue-side checks
 * and has more members than bindings are left); the bindings are a prefix of the uncut cycle's.
 * The engine looks one batch ahead, so a few asks just behind the cut may already carry a cause (INVALID / SLOWPATH /
 * SKIPPED / NOFIT) that the uncut order would have given them later; every cycle starts all unallocated asks as pending
 * again, so this never changes a later binding. */
int yk_cycle(yk_engine* e, uint32_t max_bindings, yk_binding* out, uint32_t* n_out,
             uint32_t* slow_path_asks, uint32_t slow_cap, uint32_t* n_slow);

int yk_ask_states(yk_engine* e, uint32_t n, const uint32_t* idx, uint8_t* state_out);
/* current node available, [D][n] */
int yk_nodes_available(yk_engine* e, uint32_t n, const uint32_t* idx, int64_t* avail_out);
/* one (ask,node) answer on the current state, evaluated on the device: 0 fits / YK_FAIL_* / <0 error */
int yk_evaluate(yk_engine* e, uint32_t ask, uint32_t node);
/* the same answer for the reservation phase, Predicates(Allocate = false): the reference then runs the same plugins WITHOUT
 * NodeResourcesFit (pkg/plugin/predicates/predicate_manager.go:130-135, reservation filter set :353-368) and the core does
 * not ask for available resources: YK_FAIL_AVAILABLE / YK_FAIL_RESOURCES cannot come back.  (Reservations themselves --
 * an old unplaceable ask pinning a node -- are not made by yk_cycle: SURVEY A.1, DESIGN.md.) */
int yk_evaluate_reserve(yk_engine* e, uint32_t ask, uint32_t node);
/* node sort keys as the device computes them (float64 score bits), for known-answer tests */
int yk_node_scores(yk_engine* e, uint32_t n, const uint32_t* idx, double* score_out);

/* Preemption victim search, batched (ResourceManagerCallback.PreemptionPredicates,
 * pkg/cache/scheduler_callback.go:200-209 -> context.go:705-729 -> predicate_manager.go:137-175).  Query q: the
 * smallest index i >= start[q] such that ask[q] passes every predicate on node[q] once the victims
 * victim_off[q] .. victim_off[q]+i (in the given order) are removed, or -1.  victim_req is [D][victim_off[n]]: what
 * each victim gives back to the node.  Known answers: TestPreemptionPredicates(+Empty), predicate_manager_test.go:67-134. */
int yk_preemption_search(yk_engine* e, uint32_t n_queries, const uint32_t* ask, const uint32_t* node,
                         const uint32_t* victim_off, const int64_t* victim_req, const uint32_t* start, int32_t* index_out);

/* multi-GPU exchange hook: called once per batch on every rank, after the local shard's rows of `buf`
 * (device pointer, row-major, `row_bytes` per ask, rows [first_row, first_row+n_rows) are this rank's)
 * have been written on `stream`; must make all `total_rows` rows valid on every rank (an all-gather). */
typedef int (*yk_allgather_fn)(void* ctx, void* device_buf, uint64_t row_bytes, uint32_t first_row,
                               uint32_t n_rows, uint32_t total_rows, void* cuda_stream);
int yk_set_exchange(yk_engine* e, yk_allgather_fn fn, void* ctx);

/* Peer-to-peer exchange (preferred over yk_set_exchange when all ranks sit on one NVLink / NVSwitch node): every
 * rank exports its fit buffers and sync block as CUDA IPC handles, imports every other rank's, and from then on the
 * sweep kernel stores its rows straight into all ranks' buffers -- the all-gather is fused into the sweep, ordered by
 * sequence flags in peer memory, with no collective launch and no host involvement.  Call yk_peer_export on every
 * rank, exchange the blobs (any transport), call yk_peer_import once per other rank, then yk_peer_enable. */
typedef struct { unsigned char blob[3][64]; } yk_peer_handles;   /* cudaIpcMemHandle_t x {slot 0, slot 1, sync} */
int yk_peer_export(yk_engine* e, yk_peer_handles* out);
int yk_peer_import(yk_engine* e, uint32_t peer_rank, const yk_peer_handles* in);
int yk_peer_enable(yk_engine* e);

int yk_stats(yk_engine* e, yk_stats_t* out);
int yk_stats_reset(yk_engine* e);
const char* yk_strerror(int status);
const char* yk_last_error(yk_engine* e);
uint32_t yk_abi_version(void);

#ifdef __cplusplus
}
#endif
#endif
