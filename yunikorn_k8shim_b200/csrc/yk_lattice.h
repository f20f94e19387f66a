// yk_lattice.h -- the device-resident ordered commit ("lattice commit"), single source:
//   * compiled by nvcc it is the body of yk_lattice_kernel (one persistent CTA of 1024 threads, sm_100a);
//   * compiled by g++ (tests/host/lattice_shim.cpp) the SAME code runs as a sequential emulation, so the CPU test suite
//     fuzzes the algorithm against the oracle without a GPU (test infrastructure; the product only ever runs the kernel).
//
// What it replaces: the reference's one-allocation-per-pass loop [EXT yunikorn-core Application.tryNodes / tryNode, entered
// per (ask,node) pair through pkg/cache/scheduler_callback.go:196-198 -> pkg/cache/context.go:683-703 ->
// pkg/plugin/predicates/predicate_manager.go:130-283] under the `fair` node-sort policy, for a WINDOW of consecutive asks
// that request at most SMAX distinct resource vectors ("shapes": the replicas of a few deployments / jobs / task groups).
//
// Why a window is parallel.  Under `fair` a node's key (float64 usage score) never decreases when something is
// allocated on it.  Within a window a node is in a STATE c = (c_1..c_S): how many allocations of each shape it has taken.
// State c of node n has a key k_n(c) that is known in advance (available - sum c_s r_s, re-scored), and k_n is monotone in
// every c_s.  The sequential loop "take the minimum (key, NodeID) over the nodes that pass, allocate, re-key" therefore
// walks a LATTICE of elements (node, state): an ask of shape s takes the first element, in (key, NodeID) order, that is
// currently AVAILABLE (its node is in exactly that state), has room for one more s, and whose node passes the ask's own
// taint / affinity / node-name masks; the element becomes unavailable and its child (node, c + e_s) becomes available.
// All of that is bit arithmetic on a sorted element list:   first set bit of  AVAIL & ROOM_s & ACCEPT_sig   per ask --
// a few register operations for one warp -- while everything else (which states exist, their float64 keys, the room and
// mask tests, the child links, the new node order) is evaluated in parallel by the whole CTA.
//
// One sub-run =
//   1. stage   the next asks up to SMAX distinct shapes / SIGCAP distinct signatures; never cut inside a gang
//   2. scan    the first P positions of the node order (ascending (key, NodeID)), gather their records
//   3. window  choose a key threshold tau (largest of eight candidates whose state count fits) and per node / shape the
//              depth d_s = how many allocations of s alone stay below tau; the node's states are the box [0,d_1]x..x[0,d_S]
//   4. bound   the smallest key just outside any box ((d_s+1) e_s: boxes are cut along the axes and keys are monotone)
//              and the key of position P: every element below the bound is known, so the sorted list below it is exact
//   5. lattice evaluate the box states (reachable? key below the bound?), sort when there is more than the bases
//   6. links   child of (element, shape), ROOM_s bitmaps, one ACCEPT bitmap per distinct predicate signature
//   7. chain   warp 0, sequential over the asks; gangs snapshot / restore the AVAIL words
//   8. apply   available -= sum c_s r_s on the touched nodes, float64 re-score, bindings out
//   9. patch   the touched nodes leave their positions and are merged back by their new keys
// An ask that finds nothing below the bound ends the sub-run there (the next one starts from the new order); an ask that
// finds nothing in a fresh sub-run is decided by a full scan of the order (exact first fit, or certain NOFIT, which also
// refreshes the per-dimension capacity bound that rejects later hopeless asks without a scan).
#pragma once
#include <stdint.h>
#include "yk_score.h"

#if defined(__CUDACC__)
#define LT_DEV 1
#define LT_HD __host__ __device__ __forceinline__
#define LT_FN __device__ __forceinline__
#define LT_FOR(i, n) for (int i = (int)threadIdx.x; i < (int)(n); i += yklt::THREADS)
#define LT_SYNC() __syncthreads()
#define LT_ONE if (threadIdx.x == 0)
#else
#define LT_DEV 0
#define LT_HD inline
#define LT_FN inline
#define LT_FOR(i, n) for (int i = 0; i < (int)(n); ++i)
#define LT_SYNC() ((void)0)
#define LT_ONE if (true)
#endif

namespace yklt {

constexpr int THREADS = 1024;
constexpr int LCAP = 2048;              // lattice elements of one sub-run
constexpr int FW = LCAP / 32;           // words per bitmap over the elements
constexpr int KCAP = 512;               // asks per sub-run
constexpr int SIGCAP = 128;             // distinct (consecutive) predicate signatures per sub-run
constexpr int SMAX = 8;                 // distinct request vectors ("shapes") per sub-run
constexpr int DCAP = 63;                // deepest a box goes along one shape
constexpr int VCAP = 256;               // states per node
constexpr int SHAPE_IDS = 2048;         // shape ids a cycle may use (the host numbers the distinct request vectors)
constexpr int NCAND = 8;                // candidate key thresholds per sub-run
constexpr uint32_t NONE = 0xFFFFFFFFu;
constexpr uint16_t NO_CHILD = 0xFFFFu;
constexpr uint64_t KEY_INF = 0xFFFFFFFFFFFFFFFFull;

// meta[] word of a batch entry (written by the host, which knows requests, signatures and gangs):
//   bits 0-7 flags, bits 16-31 shape id (dense number of the request vector within the cycle, < SHAPE_IDS)
constexpr uint32_t M_RUN = 1;           // request vector differs from the previous entry's (or first entry)
constexpr uint32_t M_SIG = 2;           // predicate signature differs from the previous entry's (always set with M_RUN)
constexpr uint32_t M_GSTART = 4;        // first member of a gang
constexpr uint32_t M_GANG = 8;          // member of a gang

// hdr[] words
enum { H_CONSUMED = 0, H_STATUS = 1, H_SUBRUNS = 2, H_FULLSCANS = 3, H_SORTS = 4, H_ELEMS = 5, H_QUICK = 6, H_ESC = 7, H_WORDS = 8 };
enum { ST_DONE = 0, ST_STOPPED = 1, ST_HANDOFF = 2, ST_NAN = 3 };

struct Ent { uint64_t key; uint64_t rn; };   // one position of the node order: sort key, NodeID rank << 32 | node index

LT_HD bool ent_less(uint64_t k0, uint64_t r0, uint64_t k1, uint64_t r1) { return k0 < k1 || (k0 == k1 && r0 < r1); }

struct Args {
    uint32_t policy;
    double w[8];
    int64_t* rec; int RS;                  // node records: [0,D) available, [D,2D) total, [2D] taint, [2D+1] label, [2D+2] rank<<32|flags
    Ent* ord[2]; int* cur;                 // the node order, ping-pong; *cur = which buffer is current
    int nlive;
    const int64_t* a_req; size_t lda;      // ask table (column-major, as in yk_kernels.cuh)
    const uint64_t* a_tol; const uint64_t* a_need; const uint64_t* a_deny; const uint32_t* a_node;
    const uint32_t* asks; const uint32_t* meta; int B;
    uint32_t* res;                         // [B] node index or NONE
    int* hdr;                              // [H_WORDS]
    int64_t* ub;                           // [8] per-dimension upper bound of what any node can still hold (exact after a full scan)
    int insensitive;                       // 1: a failed ask does not end the batch (placement-insensitive order)
};

// can a node in this state (available, total) take `req` once more (same fold as yk_gather_kernel: unusable -> nothing
// fits; request <= min(max(0,total), max(0,available)) on every dimension)
template <int D>
LT_HD bool fits(bool usable, const int64_t* avail, const int64_t* total, const int64_t* req) {
    if (!usable) return false;
    for (int k = 0; k < D; ++k) {
        const int64_t r = req[k];
        const int64_t t = total[k] < 0 ? 0 : total[k];
        int64_t a = avail[k];
        if (a < 0) a = 0;
        if (r > t || r > a) return false;
    }
    return true;
}

LT_HD bool accepts(uint64_t taint, uint64_t label, uint32_t node, uint64_t tol, uint64_t need, uint64_t deny, uint32_t want) {
    if ((taint & ~tol) | (~label & need) | (label & deny)) return false;
    return want == NONE || want == node;
}

template <int D>
struct Shared {
    static constexpr int PMAX = D <= 4 ? 512 : 256;
    // ---- scalars (written inside LT_ONE or by shared atomics, read after LT_SYNC) ----
    int n, P, nrows, nshapes, cut, gstart_before_cut;
    int nbase, nvalid, nt, tdone, need_more, status, hit, esc_pos, cnt_elems, total_slots, pick;
    unsigned long long bound_key, bound_sec, pos_key;
    unsigned long long ubx[8];
    int sumv[NCAND];
    double tau[NCAND];
    int64_t sreq[SMAX * 8];       // request vector of each shape of the sub-run
    // ---- staged positions ----
    uint64_t okey[PMAX], orn[PMAX], taint[PMAX], label[PMAX];
    int64_t avail[PMAX * D], total[PMAX * D];
    uint8_t usable[PMAX];
    uint8_t depth[PMAX * SMAX];   // box of the node: states c with c_s <= depth[s]
    alignas(4) uint16_t cnt[PMAX * SMAX];   // allocations of each shape the chain gave to the position
    uint32_t off[PMAX + 1];       // first slot of the node's box
    uint32_t tb[PMAX + 1];        // touched positions before p
    // ---- lattice elements: key, sec = NodeID rank << 32 | allocations in the state << 22 | slot ----
    unsigned long long e_key[LCAP], e_sec[LCAP];
    uint16_t slot_p[LCAP];        // slot -> position
    uint16_t idx_of[LCAP];        // slot -> sorted element (NO_CHILD: not below the bound)
    uint16_t child[LCAP * SMAX];  // (element, shape) -> element of the node's next state
    uint32_t ROOM[SMAX * FW];     // per shape: the element's node can take one more of it
    uint32_t BASE[FW];            // elements that are a node's initial state
    alignas(16) uint32_t F[SIGCAP * FW];   // ACCEPT rows; dead after the chain: scratch for the apply / patch phases
    uint32_t scan[LCAP];
    uint64_t r_tol[SIGCAP], r_need[SIGCAP], r_deny[SIGCAP];
    uint32_t r_want[SIGCAP];
    uint8_t k_row[KCAP], k_meta[KCAP], k_ls[KCAP];
    int32_t sel[KCAP];            // element taken by the ask, -1 none
    int shape_first[SHAPE_IDS];
    uint8_t shape_loc[SHAPE_IDS];
    // touched nodes sorted by their new (key, rn): lives in F after the chain
    unsigned long long* t_key() { return reinterpret_cast<unsigned long long*>(F + LCAP); }
    unsigned long long* t_rn() { return reinterpret_cast<unsigned long long*>(F + LCAP) + PMAX; }
    const unsigned long long* t_key() const { return reinterpret_cast<const unsigned long long*>(F + LCAP); }
    const unsigned long long* t_rn() const { return reinterpret_cast<const unsigned long long*>(F + LCAP) + PMAX; }
    uint32_t* tmp() { return F; }   // LCAP words
};
static_assert(SIGCAP * FW >= LCAP + 4 * 512, "scratch after the chain must fit the ACCEPT rows");
static_assert(512 * SMAX <= 2 * LCAP, "corner keys use the element arrays as scratch");

// ---- block-wide helpers (each is a sequence of LT_FOR regions: runs unchanged as plain loops on the host) -------
// exclusive prefix sum of x[0..n) (n <= LCAP) in place, total in *tot (valid on thread 0 / the host); tmp has n words
LT_FN void block_scan(uint32_t* x, uint32_t* tmp, int n, int* tot) {
#if LT_DEV
    uint32_t* in = x; uint32_t* out = tmp;   // Hillis-Steele, ping-pong between x and tmp
    for (int off = 1; off < n; off <<= 1) {
        LT_FOR(i, n) out[i] = in[i] + (i >= off ? in[i - off] : 0u);
        LT_SYNC();
        uint32_t* sw = in; in = out; out = sw;
    }
    LT_FOR(i, n) out[i] = i ? in[i - 1] : 0u;   // inclusive -> exclusive, into the other buffer
    LT_ONE *tot = n ? (int)in[n - 1] : 0;
    LT_SYNC();
    if (out != x) { LT_FOR(i, n) x[i] = out[i]; LT_SYNC(); }
#else
    uint32_t run = 0;
    for (int i = 0; i < n; ++i) { const uint32_t v = x[i]; x[i] = run; run += v; }
    *tot = (int)run;
    (void)tmp;
#endif
}

// ascending bitonic sort of (key, sec) pairs, n a power of two
LT_FN void block_sort(unsigned long long* key, unsigned long long* sec, int n) {
    for (int k = 2; k <= n; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            LT_FOR(x, n >> 1) {
                const int i = ((x / j) * 2 * j) + (x % j);   // lower index of the pair
                const int l = i + j;
                const bool up = (i & k) == 0;
                const unsigned long long ki = key[i], kl = key[l], si = sec[i], sl = sec[l];
                const bool gt = ki > kl || (ki == kl && si > sl);
                if (gt == up) { key[i] = kl; key[l] = ki; sec[i] = sl; sec[l] = si; }
            }
            LT_SYNC();
        }
}

LT_FN void smin64(unsigned long long* p, unsigned long long v) {
#if LT_DEV
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
LT_FN void smax64(unsigned long long* p, unsigned long long v) {
#if LT_DEV
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
LT_FN void smin32(int* p, int v) {
#if LT_DEV
    atomicMin(p, v);
#else
    if (v < *p) *p = v;
#endif
}
LT_FN void smax32(int* p, int v) {
#if LT_DEV
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
LT_FN int sadd32(int* p, int v) {
#if LT_DEV
    return atomicAdd(p, v);
#else
    const int o = *p; *p += v; return o;
#endif
}
LT_FN void sor32(uint32_t* p, uint32_t v) {
#if LT_DEV
    atomicOr(p, v);
#else
    *p |= v;
#endif
}
LT_FN void sinc16(uint16_t* p) {   // 16-bit counter inside a 32-bit word of shared memory
#if LT_DEV
    const size_t a = (size_t)p;
    uint32_t* w = (uint32_t*)(a & ~(size_t)3);
    atomicAdd(w, (a & 2) ? 0x10000u : 1u);
#else
    ++*p;
#endif
}

template <typename T>
LT_FN T ldg(const T* p) {
#if LT_DEV
    return __ldcg(p);     // L2: the records and the order are rewritten by this kernel, L1 lines could be stale
#else
    return *p;
#endif
}

// ---- per node: the depth of its box along every shape for a key threshold tau (score units) ------------------
// d_s = how many allocations of shape s alone keep the (linearly estimated) score below tau, capped; then the box is
// shrunk (deepest side halved) until it has at most VCAP states.  A heuristic: exactness never depends on the depths,
// only on the bound computed from the true keys just outside the boxes.  Returns the number of states.
template <int D>
LT_HD int box_of(double sc, double tau, const double* dlt /*[S]*/, int S, uint8_t* d /*[SMAX]*/) {
    int v = 1;
    for (int s = 0; s < S; ++s) {
        int x = 0;
        if (sc < tau) {
            if (dlt[s] > 0.0) {
                const double q = (tau - sc) / dlt[s];
                x = q >= (double)DCAP ? DCAP : (int)q;
            } else x = DCAP;   // the shape does not move the key at all
        }
        d[s] = (uint8_t)x;
        v *= x + 1;
        if (v > (1 << 24)) v = 1 << 24;
    }
    while (v > VCAP) {
        int m = 0;
        for (int s = 1; s < S; ++s) if (d[s] > d[m]) m = s;
        d[m] = (uint8_t)(d[m] >> 1);
        v = 1;
        for (int s = 0; s < S; ++s) v *= d[s] + 1;
    }
    return v;
}

// estimated score increase of one allocation of `req` on the node (the weighted mean of req/total, as yk_node_score)
template <int D>
LT_HD double delta_of(const double* w, const int64_t* total, const int64_t* req) {
    double u = 0.0, tw = 0.0;
    for (int k = 0; k < D; ++k) {
        if (!(w[k] > 0.0) || total[k] <= 0) continue;
        u += w[k] * ((double)req[k] / (double)total[k]);
        tw += w[k];
    }
    return tw > 0.0 ? u / tw : 0.0;
}

// number of touched entries (sorted t_key/t_rn[0..nt)) that sort before (key, rn)
template <int D>
LT_FN int touched_before(const Shared<D>& s, int nt, uint64_t key, uint64_t rn) {
    const unsigned long long* tk = s.t_key();
    const unsigned long long* tr = s.t_rn();
    int lo = 0, hi = nt;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (ent_less(tk[mid], tr[mid], key, rn)) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// ---- patch: src order minus the touched positions, merged with the touched nodes under their new keys -> dst ------
// touched positions: sub-run mode (esc_pos < 0): positions p < P with tb[p+1] > tb[p];
//                    single mode (esc_pos >= 0): exactly position esc_pos
template <int D>
LT_FN void patch_order(const Args& a, Shared<D>& s, const Ent* src, Ent* dst) {
    const int nt = s.nt, P = s.P, esc = s.esc_pos, nlive = a.nlive;
    const unsigned long long* tk = s.t_key();
    const unsigned long long* tr = s.t_rn();
    LT_FOR(p, nlive) {
        Ent e;
        e.key = ldg(&src[p].key); e.rn = ldg(&src[p].rn);
        const bool touched = esc >= 0 ? (p == esc) : (p < P && s.tb[p + 1] != s.tb[p]);
        const int tbp = esc >= 0 ? (p > esc ? 1 : 0) : (p < P ? (int)s.tb[p] : nt);
        const int ib = touched_before(s, nt, e.key, e.rn);
        // the touched entries that fall between the old tuples of positions p-1 and p are written by position p
        int ibp = 0;
        if (p > 0) {
            const uint64_t pk = ldg(&src[p - 1].key), pr = ldg(&src[p - 1].rn);
            ibp = touched_before(s, nt, pk, pr);
        }
        const int before = p - tbp;   // untouched old entries before position p
        for (int i = ibp; i < ib; ++i) { Ent te; te.key = tk[i]; te.rn = tr[i]; dst[i + before] = te; }
        if (!touched) dst[before + ib] = e;
        if (p == nlive - 1) {   // touched entries behind every old tuple
            const int all = before + (touched ? 0 : 1);
            for (int i = ib; i < nt; ++i) { Ent te; te.key = tk[i]; te.rn = tr[i]; dst[i + all] = te; }
        }
    }
    LT_SYNC();
}

// ---- chain: sequential first fit over the sorted elements --------------------------------------------------------
// asks [0, n) of the sub-run.  Sets sel[], tdone (asks decided), need_more (an ask found nothing below the bound: the
// sub-run ends before it -- before its gang when it is a gang member).
template <int D>
LT_FN void chain(Shared<D>& s) {
    const int n = s.n, nw = (s.nvalid + 31) >> 5;
#if LT_DEV
    if (threadIdx.x < 32) {
        const int lane = threadIdx.x;
        uint32_t a0 = lane < nw ? s.BASE[lane] : 0u, a1 = lane + 32 < nw ? s.BASE[lane + 32] : 0u;   // AVAIL words lane, lane+32
        uint32_t g0 = 0, g1 = 0;   // snapshot at the gang start
        int gbeg = -1, done = n, more = 0;
        for (int i = 0; i < n; ++i) {
            const uint32_t m = s.k_meta[i];
            if (m & M_GSTART) { gbeg = i; g0 = a0; g1 = a1; }
            else if (!(m & M_GANG)) gbeg = -1;
            const int sh = s.k_ls[i];
            const uint32_t* row = s.F + (int)s.k_row[i] * FW;
            const uint32_t* room = s.ROOM + sh * FW;
            const uint32_t f0 = lane < nw ? (row[lane] & room[lane] & a0) : 0u;
            const uint32_t f1 = lane + 32 < nw ? (row[lane + 32] & room[lane + 32] & a1) : 0u;
            const uint32_t b0 = __ballot_sync(0xFFFFFFFFu, f0 != 0u);
            int e = -1, ch = -1;
            if (b0) {
                const int wl = __ffs((int)b0) - 1;
                if (lane == wl) { const int bit = __ffs((int)f0) - 1; a0 &= ~(1u << bit); e = wl * 32 + bit; ch = s.child[e * SMAX + sh]; }
                e = __shfl_sync(0xFFFFFFFFu, e, wl); ch = __shfl_sync(0xFFFFFFFFu, ch, wl);
            } else {
                const uint32_t b1 = __ballot_sync(0xFFFFFFFFu, f1 != 0u);
                if (b1) {
                    const int wl = __ffs((int)b1) - 1;
                    if (lane == wl) { const int bit = __ffs((int)f1) - 1; a1 &= ~(1u << bit); e = (wl + 32) * 32 + bit; ch = s.child[e * SMAX + sh]; }
                    e = __shfl_sync(0xFFFFFFFFu, e, wl); ch = __shfl_sync(0xFFFFFFFFu, ch, wl);
                }
            }
            if (e < 0) {
                more = 1;
                if (gbeg >= 0) { done = gbeg; a0 = g0; a1 = g1; }   // the gang is decided as a whole by a later sub-run
                else done = i;
                break;
            }
            if (ch != (int)NO_CHILD) {   // the node's next state becomes available
                const int wd = ch >> 5;
                if (wd == lane) a0 |= 1u << (ch & 31);
                else if (wd == lane + 32) a1 |= 1u << (ch & 31);
            }
            if (lane == 0) s.sel[i] = e;
        }
        if (lane == 0) { s.tdone = done; s.need_more = more; }
    }
    LT_SYNC();
#else
    uint32_t av[FW], snap[FW];
    for (int w = 0; w < FW; ++w) av[w] = w < nw ? s.BASE[w] : 0u;
    int gbeg = -1, done = n, more = 0;
    for (int i = 0; i < n; ++i) {
        const uint32_t m = s.k_meta[i];
        if (m & M_GSTART) { gbeg = i; for (int w = 0; w < FW; ++w) snap[w] = av[w]; }
        else if (!(m & M_GANG)) gbeg = -1;
        const int sh = s.k_ls[i];
        const uint32_t* row = s.F + (int)s.k_row[i] * FW;
        const uint32_t* room = s.ROOM + sh * FW;
        int e = -1;
        for (int w = 0; w < nw; ++w) {
            const uint32_t f = row[w] & room[w] & av[w];
            if (f) { const int bit = __builtin_ctz(f); av[w] &= ~(1u << bit); e = w * 32 + bit; break; }
        }
        if (e < 0) {
            more = 1;
            if (gbeg >= 0) { done = gbeg; for (int w = 0; w < FW; ++w) av[w] = snap[w]; }
            else done = i;
            break;
        }
        const int ch = s.child[e * SMAX + sh];
        if (ch != (int)NO_CHILD) av[ch >> 5] |= 1u << (ch & 31);
        s.sel[i] = e;
    }
    s.tdone = done; s.need_more = more;
#endif
}

// ---- the batch --------------------------------------------------------------------------------------------------
template <int D>
LT_FN void lattice_batch(const Args& a, Shared<D>& s) {
    constexpr int PMAX = Shared<D>::PMAX;
    const int RS = a.RS, nlive = a.nlive, B = a.B;
    int buf = ldg(a.cur);
    int t = 0;                 // next undecided batch entry
    int status = ST_DONE;
    int subruns = 0, fullscans = 0, sorts = 0, quick = 0, escs = 0;
    long long elems = 0;
    int fresh = 0;             // 1: the previous sub-run decided nothing for entry t: decide it by a full scan
    int dead_sig = 0;          // entry t-1 was a certain NOFIT and nothing was committed since: same signature -> NOFIT too
    LT_ONE s.status = ST_DONE;
    LT_FOR(i, SHAPE_IDS) s.shape_first[i] = 0x7FFFFFFF;
    LT_SYNC();

    while (t < B && status == ST_DONE) {
        // ================= stage the asks of the sub-run =================
        const int navail = B - t < KCAP ? B - t : KCAP;
        LT_ONE { s.cut = navail; s.gstart_before_cut = 0; s.nshapes = 0; }
        LT_FOR(i, navail) {
            const uint32_t m = ldg(&a.meta[t + i]);
            s.k_meta[i] = (uint8_t)m;
            s.sel[i] = (int32_t)(m >> 16);   // shape id, until the chain needs sel[]
        }
        LT_SYNC();
        LT_FOR(i, navail) { if (s.sel[i] >= SHAPE_IDS) smin32(&s.cut, i); else smin32(&s.shape_first[s.sel[i]], i); }
        LT_SYNC();
        // local shape numbers in order of first appearance; rows: a new one wherever the signature changes
        LT_FOR(i, navail) s.scan[i] = (s.sel[i] < SHAPE_IDS && s.shape_first[s.sel[i]] == i) ? 1u : 0u;
        LT_SYNC();
        { int tot; block_scan(s.scan, (uint32_t*)s.e_sec, navail, &tot); (void)tot; }
        LT_FOR(i, navail) if (s.sel[i] < SHAPE_IDS && s.shape_first[s.sel[i]] == i) {
            const int loc = (int)s.scan[i];
            s.shape_loc[s.sel[i]] = (uint8_t)(loc < 255 ? loc : 255);
            if (loc < SMAX) {
                const uint32_t ask = ldg(&a.asks[t + i]);
                for (int k = 0; k < D; ++k) s.sreq[loc * 8 + k] = ldg(&a.a_req[(size_t)k * a.lda + ask]);
            }
        }
        LT_SYNC();
        LT_FOR(i, navail) if (s.sel[i] < SHAPE_IDS) {
            const int loc = s.shape_loc[s.sel[i]];
            if (loc >= SMAX) smin32(&s.cut, i); else s.k_ls[i] = (uint8_t)loc;
        }
        LT_FOR(i, navail) s.scan[i] = (i == 0 || (s.k_meta[i] & M_SIG)) ? 1u : 0u;
        LT_SYNC();
        { int tot; block_scan(s.scan, (uint32_t*)s.e_sec, navail, &tot); (void)tot; }
        LT_FOR(i, navail) {
            const int row = (int)s.scan[i] + ((i == 0 || (s.k_meta[i] & M_SIG)) ? 1 : 0) - 1;
            if (row >= SIGCAP) smin32(&s.cut, i); else s.k_row[i] = (uint8_t)row;
            if (s.sel[i] < SHAPE_IDS) s.shape_first[s.sel[i]] = 0x7FFFFFFF;   // leave the table clean for the next sub-run
        }
        LT_SYNC();
        int n = s.cut;
        // never cut inside a gang
        const uint32_t mnext = n < navail ? (uint32_t)s.k_meta[n] : (t + n < B ? ldg(&a.meta[t + n]) : 0u);
        if (n > 0 && (mnext & M_GANG) && !(mnext & M_GSTART)) {
            LT_FOR(i, n) if (s.k_meta[i] & M_GSTART) smax32(&s.gstart_before_cut, i);
            LT_SYNC();
            n = s.gstart_before_cut;   // 0: the gang does not fit a sub-run
        }
        if (n == 0) { status = ST_HANDOFF; break; }
        LT_FOR(i, n) {
            smax32(&s.nshapes, (int)s.k_ls[i] + 1);
            s.sel[i] = -1;
            if (i == 0 || (s.k_meta[i] & M_SIG)) {
                const uint32_t ask = ldg(&a.asks[t + i]);
                const int row = s.k_row[i];
                s.r_tol[row] = ldg(&a.a_tol[ask]); s.r_need[row] = ldg(&a.a_need[ask]); s.r_deny[row] = ldg(&a.a_deny[ask]);
                s.r_want[row] = ldg(&a.a_node[ask]);
            }
        }
        LT_ONE { s.n = n; s.nrows = (int)s.k_row[n - 1] + 1; }   // rows are numbered in entry order
        LT_SYNC();
        const int nrows = s.nrows, S = s.nshapes;
        const bool gang0 = (s.k_meta[0] & M_GANG) != 0;
        const int64_t* req0 = s.sreq;   // entry 0 is the first appearance of its shape: local number 0

        // ================= hopeless request: some dimension exceeds what any node has left =================
        bool hopeless = false;
        for (int k = 0; k < D; ++k) if (req0[k] > ldg(&a.ub[k])) hopeless = true;
        if (dead_sig && !(s.k_meta[0] & M_SIG) && !gang0) hopeless = true;   // same signature as the NOFIT just decided, same state
        if (hopeless) {
            // the first entry (its whole gang when it is a gang member) is a certain NOFIT
            int g1 = 1;
            if (gang0) while (g1 < n && (s.k_meta[g1] & M_GANG) && !(s.k_meta[g1] & M_GSTART)) ++g1;
            LT_FOR(i, g1) a.res[t + i] = NONE;
            LT_SYNC();
            ++quick;
            t += g1;
            dead_sig = gang0 ? 0 : 1;
            if (!a.insensitive) status = ST_STOPPED;
            continue;
        }

        if (!fresh) {
            // ================= scan: stage the first P positions =================
            const int P = nlive < PMAX ? nlive : PMAX;
            const Ent* cur = a.ord[buf];
            LT_ONE {
                s.P = P; s.nt = 0; s.esc_pos = -1; s.bound_sec = KEY_INF;
                // every element at or after the first unscanned position is unknown
                s.pos_key = P < nlive ? ldg(&cur[P].key) : KEY_INF;
                s.bound_key = s.pos_key;
                for (int c = 0; c < NCAND; ++c) s.sumv[c] = 0;
            }
            LT_FOR(p, P) {
                const uint64_t key = ldg(&cur[p].key), rn = ldg(&cur[p].rn);
                const uint32_t node = (uint32_t)rn;
                const int64_t* r = a.rec + (size_t)node * RS;
                for (int k = 0; k < D; ++k) { s.avail[p * D + k] = ldg(&r[k]); s.total[p * D + k] = ldg(&r[D + k]); }
                s.okey[p] = key; s.orn[p] = rn;
                s.taint[p] = (uint64_t)ldg(&r[2 * D]); s.label[p] = (uint64_t)ldg(&r[2 * D + 1]);
                const uint32_t fl = (uint32_t)(uint64_t)ldg(&r[2 * D + 2]);
                s.usable[p] = ((fl & 1u) && !(fl & 2u)) ? 1 : 0;
                for (int x = 0; x < SMAX; ++x) s.cnt[p * SMAX + x] = 0;
            }
            LT_SYNC();
            // ================= window: key threshold and the boxes =================
            // candidates: the keys at positions P, P/2, P/4 ...; the largest whose boxes hold at most LCAP states wins
            LT_ONE {
                for (int c = 0; c < NCAND; ++c) {
                    const int m = P >> c;
                    s.tau[c] = m >= P ? (P < nlive ? yk_key_to_score(s.pos_key) : 1e300) : yk_key_to_score(s.okey[m > 0 ? m : 0]);
                }
            }
            LT_SYNC();
            LT_FOR(p, P) {
                double dlt[SMAX];
                uint8_t d[SMAX];
                for (int x = 0; x < S; ++x) dlt[x] = delta_of<D>(a.w, s.total + p * D, s.sreq + x * 8);
                const double sc = yk_key_to_score(s.okey[p]);
                for (int c = 0; c < NCAND; ++c) {
                    const int v = s.usable[p] ? box_of<D>(sc, s.tau[c], dlt, S, d) : 1;
                    sadd32(&s.sumv[c], v);
                }
            }
            LT_SYNC();
            LT_ONE {
                int pick = NCAND - 1;
                for (int c = NCAND - 1; c >= 0; --c) if (s.sumv[c] <= LCAP) pick = c;
                s.pick = pick;
            }
            LT_SYNC();
            const double tau = s.tau[s.pick];
            LT_FOR(p, P) {
                double dlt[SMAX];
                uint8_t d[SMAX];
                for (int x = 0; x < S; ++x) { dlt[x] = delta_of<D>(a.w, s.total + p * D, s.sreq + x * 8); d[x] = 0; }
                const int v = s.usable[p] ? box_of<D>(yk_key_to_score(s.okey[p]), tau, dlt, S, d) : 1;
                for (int x = 0; x < SMAX; ++x) s.depth[p * SMAX + x] = x < S && s.usable[p] ? d[x] : 0;
                s.scan[p] = (uint32_t)v;
            }
            LT_SYNC();
            { int tot; block_scan(s.scan, s.tmp(), P, &tot); LT_ONE s.total_slots = tot; LT_SYNC(); }
            int total_slots = s.total_slots;
            if (total_slots > LCAP) {   // cannot happen with the smallest candidate unless VCAP * P/2^(NCAND-1) + P > LCAP
                LT_FOR(p, P) { for (int x = 0; x < SMAX; ++x) s.depth[p * SMAX + x] = 0; s.scan[p] = (uint32_t)p; }
                LT_SYNC();
                total_slots = P;
            }
            LT_FOR(p, P) {
                s.off[p] = s.scan[p];
                int v = 1;
                for (int x = 0; x < S; ++x) v *= s.depth[p * SMAX + x] + 1;
                for (int x = 0; x < v; ++x) s.slot_p[s.scan[p] + x] = (uint16_t)p;
            }
            LT_ONE s.off[P] = (uint32_t)total_slots;
            LT_SYNC();
            // ================= bound: the smallest key just outside any box =================
            // element order = (key, NodeID rank, allocations in the state); sec packs rank << 32 | allocations << 22 | slot
            LT_FOR(x, P * S) {
                const int p = x / S, sh = x % S;
                const int j = s.depth[p * SMAX + sh] + 1;
                int64_t av[D];
                bool ok = s.usable[p] != 0;
                for (int k = 0; k < D && ok; ++k) {
                    const int64_t r = s.sreq[sh * 8 + k];
                    if (r > 0 && r > INT64_MAX / j) ok = false;
                    av[k] = s.avail[p * D + k] - (int64_t)(j - 1) * r;   // the state before the j-th allocation
                }
                uint64_t tk = KEY_INF;
                if (ok && fits<D>(true, av, s.total + p * D, s.sreq + sh * 8)) {
                    for (int k = 0; k < D; ++k) av[k] -= s.sreq[sh * 8 + k];
                    tk = yk_key_bits(yk_node_score(D, a.policy, a.w, s.total + p * D, av, 1));
                    smin64(&s.bound_key, tk);
                }
                if (x < LCAP) s.e_key[x] = tk; else s.e_sec[x - LCAP] = tk;   // scratch until the elements are written (P * S <= 2 LCAP)
            }
            LT_SYNC();
            {
                const uint64_t bk0 = s.bound_key;
                LT_FOR(x, P * S) if ((x < LCAP ? s.e_key[x] : s.e_sec[x - LCAP]) == bk0 && bk0 != KEY_INF) {
                    const int p = x / S, sh = x % S;
                    smin64(&s.bound_sec, ((s.orn[p] >> 32) << 32) | ((unsigned long long)(s.depth[p * SMAX + sh] + 1) << 22));
                }
                LT_ONE if (P < nlive && s.pos_key == bk0) smin64(&s.bound_sec, (ldg(&cur[P].rn) >> 32) << 32);
                LT_SYNC();
            }
            const uint64_t bk = s.bound_key, bs = s.bound_sec;
            // ================= lattice: the states of every box =================
            // bases (state 0) keep their position order; the other states are appended and sorted in afterwards
            LT_FOR(p, P) {
                const uint64_t sec = ((s.orn[p] >> 32) << 32) | (unsigned long long)s.off[p];
                s.scan[p] = (s.usable[p] && ent_less(s.okey[p], sec, bk, bs)) ? 1u : 0u;
            }
            LT_SYNC();
            { int nb; block_scan(s.scan, s.tmp(), P, &nb); LT_ONE { s.nbase = nb; s.cnt_elems = nb; } LT_SYNC(); }
            const int nbase = s.nbase;
            LT_FOR(p, P) {
                const uint64_t sec = ((s.orn[p] >> 32) << 32) | (unsigned long long)s.off[p];
                if (s.usable[p] && ent_less(s.okey[p], sec, bk, bs)) { const int x = (int)s.scan[p]; s.e_key[x] = s.okey[p]; s.e_sec[x] = sec; }
            }
            LT_SYNC();
            LT_FOR(x, total_slots) {
                const int p = s.slot_p[x];
                int loc = x - (int)s.off[p];
                if (loc == 0) continue;   // the base
                int64_t av[D];
                for (int k = 0; k < D; ++k) av[k] = s.avail[p * D + k];
                int csum = 0;
                bool ok = true;
                // the state is reachable iff its allocations fit one after the other (any order: the tests are monotone)
                for (int sh = 0; sh < S; ++sh) {
                    const int dim = s.depth[p * SMAX + sh] + 1;
                    const int c = loc % dim;
                    loc /= dim;
                    csum += c;
                    for (int k = 0; k < D && ok; ++k) {
                        const int64_t r = s.sreq[sh * 8 + k];
                        if (c > 0 && r > 0) {
                            if (r > INT64_MAX / c) { ok = false; break; }
                            const int64_t tt = s.total[p * D + k] < 0 ? 0 : s.total[p * D + k];
                            if (r > tt) { ok = false; break; }
                            av[k] -= (int64_t)c * r;
                            if (av[k] < 0) { ok = false; break; }
                        }
                    }
                }
                if (!ok) continue;
                const uint64_t key = yk_key_bits(yk_node_score(D, a.policy, a.w, s.total + p * D, av, 1));
                const uint64_t sec = ((s.orn[p] >> 32) << 32) | ((unsigned long long)csum << 22) | (unsigned long long)x;
                if (ent_less(key, sec, bk, bs)) {
                    const int y = sadd32(&s.cnt_elems, 1);
                    s.e_key[y] = key; s.e_sec[y] = sec;
                }
            }
            LT_SYNC();
            const int nvalid = s.cnt_elems;
            if (nvalid > nbase) {   // deeper states interleave with the bases: sort (bases alone are already in order)
                int pn = 1;
                while (pn < nvalid) pn <<= 1;
                LT_FOR(x, pn - nvalid) { s.e_key[nvalid + x] = KEY_INF; s.e_sec[nvalid + x] = KEY_INF; }
                LT_SYNC();
                block_sort(s.e_key, s.e_sec, pn);
                ++sorts;
            }
            LT_ONE s.nvalid = nvalid;
            elems += nvalid;
            // ================= links: slot -> element, child of (element, shape), ROOM and BASE bitmaps =================
            const int nw = (nvalid + 31) >> 5;
            LT_FOR(x, total_slots) s.idx_of[x] = NO_CHILD;
            LT_FOR(x, (SMAX + 1) * FW) { if (x < SMAX * FW) s.ROOM[x] = 0u; else s.BASE[x - SMAX * FW] = 0u; }
            LT_SYNC();
            LT_FOR(e, nvalid) s.idx_of[(int)(s.e_sec[e] & 0x3FFFFFu)] = (uint16_t)e;
            LT_SYNC();
            LT_FOR(x, nvalid * S) {
                const int e = x / S, sh = x % S;
                const int slot = (int)(s.e_sec[e] & 0x3FFFFFu);
                const int p = s.slot_p[slot];
                int loc = slot - (int)s.off[p];
                if (sh == 0 && loc == 0) sor32(&s.BASE[e >> 5], 1u << (e & 31));
                // decode the state, test one more allocation of shape sh
                int64_t av[D];
                for (int k = 0; k < D; ++k) av[k] = s.avail[p * D + k];
                int stride = 1, my_c = 0, my_stride = 0;
                for (int q = 0; q < S; ++q) {
                    const int dim = s.depth[p * SMAX + q] + 1;
                    const int c = loc % dim;
                    loc /= dim;
                    if (q == sh) { my_c = c; my_stride = stride; }
                    stride *= dim;
                    for (int k = 0; k < D; ++k) av[k] -= (int64_t)c * s.sreq[q * 8 + k];
                }
                const bool room = fits<D>(true, av, s.total + p * D, s.sreq + sh * 8);
                if (room) sor32(&s.ROOM[sh * FW + (e >> 5)], 1u << (e & 31));
                s.child[e * SMAX + sh] = (room && my_c < s.depth[p * SMAX + sh]) ? s.idx_of[slot + my_stride] : NO_CHILD;
            }
            // ================= ACCEPT rows =================
            LT_FOR(x, nrows * nw) {
                const int row = x / nw, wd = x % nw;
                const uint64_t tol = s.r_tol[row], need = s.r_need[row], deny = s.r_deny[row];
                const uint32_t want = s.r_want[row];
                uint32_t bits = 0;
                const int e1 = wd * 32 + 32 < nvalid ? wd * 32 + 32 : nvalid;
                for (int e = wd * 32; e < e1; ++e) {
                    const int p = s.slot_p[(int)(s.e_sec[e] & 0x3FFFFFu)];
                    if (accepts(s.taint[p], s.label[p], (uint32_t)s.orn[p], tol, need, deny, want)) bits |= 1u << (e & 31);
                }
                s.F[row * FW + wd] = bits;
            }
            LT_SYNC();
            // ================= chain =================
            chain<D>(s);
            ++subruns;
            const int tdone = s.tdone;
            if (tdone == 0) {   // nothing decided: the first entry needs the full scan
                fresh = 1;
                continue;
            }
            // ================= apply =================
            LT_FOR(i, tdone) {
                const int e = s.sel[i];
                const int p = s.slot_p[(int)(s.e_sec[e] & 0x3FFFFFu)];
                a.res[t + i] = (uint32_t)s.orn[p];
                sinc16(&s.cnt[p * SMAX + s.k_ls[i]]);
            }
            LT_SYNC();
            LT_FOR(p, P) {
                uint32_t any = 0;
                for (int x = 0; x < S; ++x) any |= s.cnt[p * SMAX + x];
                s.scan[p] = any ? 1u : 0u;
            }
            LT_SYNC();
            { int nt; block_scan(s.scan, s.tmp(), P, &nt); LT_ONE s.nt = nt; LT_SYNC(); }
            const int nt = s.nt;
            {
                unsigned long long* tk = s.t_key();
                unsigned long long* tr = s.t_rn();
                LT_FOR(p, P) {
                    s.tb[p] = s.scan[p];
                    uint32_t any = 0;
                    for (int x = 0; x < S; ++x) any |= s.cnt[p * SMAX + x];
                    if (any) {
                        int64_t na[D];
                        int64_t* r = a.rec + (size_t)(uint32_t)s.orn[p] * RS;
                        for (int k = 0; k < D; ++k) {
                            int64_t v = s.avail[p * D + k];
                            for (int x = 0; x < S; ++x) v -= (int64_t)s.cnt[p * SMAX + x] * s.sreq[x * 8 + k];
                            na[k] = v; r[k] = v;
                        }
                        const uint64_t nk = yk_key_bits(yk_node_score(D, a.policy, a.w, s.total + p * D, na, 1));
                        if (nk == YK_KEY_NAN) s.status = ST_NAN;
                        const int x = (int)s.scan[p];
                        tk[x] = nk; tr[x] = s.orn[p];
                    }
                }
                LT_ONE { s.tb[P] = (uint32_t)nt; }
                LT_SYNC();
                if (s.status == ST_NAN) { status = ST_NAN; break; }
                int pn = 1;   // touched nodes by their new keys
                while (pn < nt) pn <<= 1;
                LT_FOR(x, pn - nt) { tk[nt + x] = KEY_INF; tr[nt + x] = KEY_INF; }
                LT_SYNC();
                if (pn > 1) block_sort(tk, tr, pn);
            }
            patch_order<D>(a, s, a.ord[buf], a.ord[buf ^ 1]);
            buf ^= 1;
            t += tdone;
            dead_sig = 0;
            continue;
        }

        // ================= full scan: exact first fit of entry t over the whole order =================
        fresh = 0;
        ++fullscans;
        {
            const Ent* cur = a.ord[buf];
            const uint64_t tol = s.r_tol[0], need = s.r_need[0], deny = s.r_deny[0];
            const uint32_t want = s.r_want[0];
            LT_ONE { s.hit = nlive; for (int k = 0; k < D; ++k) s.ubx[k] = 0; }
            LT_SYNC();
            for (int base = 0; base < nlive && s.hit == nlive; base += THREADS * 4) {
                const int lim = nlive - base < THREADS * 4 ? nlive - base : THREADS * 4;
                LT_FOR(x, lim) {
                    const int p = base + x;
                    const uint64_t rn = ldg(&cur[p].rn);
                    const uint32_t node = (uint32_t)rn;
                    const int64_t* r = a.rec + (size_t)node * RS;
                    int64_t av[D], to[D];
                    for (int k = 0; k < D; ++k) { av[k] = ldg(&r[k]); to[k] = ldg(&r[D + k]); }
                    const uint32_t fl = (uint32_t)(uint64_t)ldg(&r[2 * D + 2]);
                    const bool usable = (fl & 1u) && !(fl & 2u);
                    if (usable)
                        for (int k = 0; k < D; ++k) {
                            const int64_t aa = av[k] < 0 ? 0 : av[k], tt = to[k] < 0 ? 0 : to[k];
                            smax64(&s.ubx[k], (unsigned long long)(aa < tt ? aa : tt));
                        }
                    if (fits<D>(usable, av, to, req0) &&
                        accepts((uint64_t)ldg(&r[2 * D]), (uint64_t)ldg(&r[2 * D + 1]), node, tol, need, deny, want))
                        smin32(&s.hit, p);
                }
                LT_SYNC();
            }
            const int hit = s.hit;
            if (hit == nlive) {   // certain NOFIT (of the whole gang when the entry leads one); every node was seen: the
                int g1 = 1;       // capacity bound is exact now
                if (gang0) while (g1 < n && (s.k_meta[g1] & M_GANG) && !(s.k_meta[g1] & M_GSTART)) ++g1;
                LT_FOR(i, g1) a.res[t + i] = NONE;
                LT_ONE { for (int k = 0; k < D; ++k) a.ub[k] = (int64_t)s.ubx[k]; }
                LT_SYNC();
                t += g1;
                dead_sig = gang0 ? 0 : 1;
                if (!a.insensitive) status = ST_STOPPED;
                continue;
            }
            // a gang whose first member fits somewhere but that could not be placed from the front of the order: the
            // host path decides it (it can roll a partly placed gang back)
            if (gang0) { status = ST_HANDOFF; break; }
            // commit the ask to the node at position `hit`
            ++escs;
            LT_ONE {
                const uint64_t rn = ldg(&cur[hit].rn);
                const uint32_t node = (uint32_t)rn;
                int64_t* r = a.rec + (size_t)node * RS;
                int64_t na[D], to[D];
                for (int k = 0; k < D; ++k) { na[k] = ldg(&r[k]) - req0[k]; to[k] = ldg(&r[D + k]); r[k] = na[k]; }
                const uint64_t nk = yk_key_bits(yk_node_score(D, a.policy, a.w, to, na, 1));
                if (nk == YK_KEY_NAN) s.status = ST_NAN;
                s.t_key()[0] = nk; s.t_rn()[0] = rn;
                s.nt = 1; s.esc_pos = hit; s.P = 0;
                a.res[t] = node;
            }
            LT_SYNC();
            if (s.status == ST_NAN) { status = ST_NAN; break; }
            patch_order<D>(a, s, a.ord[buf], a.ord[buf ^ 1]);
            buf ^= 1;
            t += 1;
            dead_sig = 0;
        }
    }
    LT_ONE {
        *a.cur = buf;
        a.hdr[H_CONSUMED] = t;
        a.hdr[H_STATUS] = status;
        a.hdr[H_SUBRUNS] += subruns; a.hdr[H_FULLSCANS] += fullscans; a.hdr[H_SORTS] += sorts;
        a.hdr[H_ELEMS] += (int)(elems > 0x7FFFFFFF ? 0x7FFFFFFF : elems); a.hdr[H_QUICK] += quick; a.hdr[H_ESC] += escs;
    }
    LT_SYNC();
}

}  // namespace yklt
