// yk_kernels.cuh -- sm_100a device code of the scheduling-cycle engine.
//
// Data layout in HBM (DESIGN.md "layout"): everything is structure-of-arrays, column-major.
//   node table (node-index space, resident):  total[D][ldn] i64, avail[D][ldn] i64, taint[ldn] u64,
//                                             label[ldn] u64, flags[ldn] u32
//   ask table  (ask-index space, resident):   req[D][lda] i64, tol/need/deny[lda] u64, node[lda] u32
//   sorted view (rebuilt per batch):          cap[D][Np] i64, taint/label[Np] u64, node[Np] u32, key[Np] u64
//                                             position p = p-th node in ascending (score, NodeID) order
//   outputs per batch:                        fit[B][W] u32 bitmap (bit p of row i = ask i fits sorted node p),
//                                             first[B] u32 = lowest set position = the frozen-snapshot argmin
//
// The sweep is integer compare + reduce: no tensor cores (not a contraction).  The node tile lives in
// registers and is reused across the whole ask chunk; the ask chunk is staged once in shared memory and
// read back as warp-wide broadcasts.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "yk_score.h"

#define YK_NONE_U32 0xFFFFFFFFu
#define YK_SWEEP_THREADS 256

struct YkWeights { double w[8]; };

// ---- per-node sort key --------------------------------------------------------------------------------
// key_in[i] / val_in[i] for the i-th live node in NodeID-rank order; a stable sort by key then yields
// ascending (score, NodeID) -- the order of the core's node btree.
__global__ void yk_key_kernel(int D, uint32_t policy, YkWeights w, const int64_t* __restrict__ total,
                              const int64_t* __restrict__ avail, size_t ldn, const uint32_t* __restrict__ by_rank,
                              int nlive, uint64_t* __restrict__ key_in, uint32_t* __restrict__ val_in,
                              int* __restrict__ nan_flag) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nlive) return;
    uint32_t n = by_rank[i];
    double s = yk_node_score(D, policy, w.w, total + n, avail + n, ldn);
    uint64_t k = yk_key_bits(s);
    if (k == YK_KEY_NAN) atomicExch(nan_flag, 1);
    key_in[i] = k;
    val_in[i] = n;
}

__global__ void yk_score_kernel(int D, uint32_t policy, YkWeights w, const int64_t* __restrict__ total,
                                const int64_t* __restrict__ avail, size_t ldn, const uint32_t* __restrict__ idx,
                                int n, double* __restrict__ out) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t node = idx[i];
    out[i] = yk_node_score(D, policy, w.w, total + node, avail + node, ldn);
}

// ---- sorted view ---------------------------------------------------------------------------------------
// cap_k = min(max(0,total_k), max(0,avail_k)) folds the core's FitInNode(total) and preAllocateCheck
// available.FitIn (FitIn clamps negatives to 0) and the shim's NodeResourcesFit into ONE
// compare per dimension; cap = -1 on every dimension for nodes that are unschedulable, reserved or padding,
// so no request (requests are >= 0 by the time they reach the sweep) can fit them.
__global__ void yk_gather_kernel(int D, const int64_t* __restrict__ total, const int64_t* __restrict__ avail,
                                 size_t ldn, const uint64_t* __restrict__ taint, const uint64_t* __restrict__ label,
                                 const uint32_t* __restrict__ flags, const uint32_t* __restrict__ sorted_node,
                                 int nlive, int Np, int64_t* __restrict__ s_cap, uint64_t* __restrict__ s_taint,
                                 uint64_t* __restrict__ s_label, uint32_t* __restrict__ s_node) {
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= Np) return;
    if (p < nlive) {
        uint32_t n = sorted_node[p];
        bool usable = (flags[n] & 1u) && !(flags[n] & 2u);
        for (int k = 0; k < D; ++k) {
            int64_t t = total[(size_t)k * ldn + n], a = avail[(size_t)k * ldn + n];
            t = t < 0 ? 0 : t;
            a = a < 0 ? 0 : a;
            s_cap[(size_t)k * Np + p] = usable ? (a < t ? a : t) : -1;
        }
        s_taint[p] = taint[n];
        s_label[p] = label[n];
        s_node[p] = n;
    } else {
        for (int k = 0; k < D; ++k) s_cap[(size_t)k * Np + p] = -1;
        s_taint[p] = ~0ull;
        s_label[p] = 0;
        s_node[p] = YK_NONE_U32;
    }
}

// ---- the fused sweep: resource fit + taint/affinity masks + node name + first-fit argmin ---------------
struct YkSweepArgs {
    // sorted node view
    const int64_t* s_cap;     // [D][Np]
    const uint64_t* s_taint;  // [Np]
    const uint64_t* s_label;  // [Np]
    const uint32_t* s_node;   // [Np]
    int Np;                   // multiple of YK_SWEEP_THREADS * NPT
    // ask table + batch order
    const int64_t* a_req;     // [D][lda]
    const uint64_t* a_tol;
    const uint64_t* a_need;
    const uint64_t* a_deny;
    const uint32_t* a_node;
    size_t lda;
    const uint32_t* batch;    // [B] ask indices in commit order
    int row0, rows;           // this launch handles batch rows [row0, row0+rows)
    // outputs
    uint32_t* fit;            // [B][W]
    uint32_t* first;          // [B], pre-set to YK_NONE_U32
    int W;                    // words per row = Np/32
};

template <int D, int NPT, int AC>
__global__ void __launch_bounds__(YK_SWEEP_THREADS) yk_sweep_kernel(const YkSweepArgs p) {
    __shared__ int64_t sh_req[D][AC];
    __shared__ uint64_t sh_tol[AC], sh_need[AC], sh_deny[AC];
    __shared__ uint32_t sh_node[AC];
    __shared__ uint32_t sh_first[AC];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int a0 = p.row0 + blockIdx.y * AC;
    const int na = min(AC, p.row0 + p.rows - a0);

    // stage the ask chunk (gather by batch order; rows past the end can never fit)
    for (int i = tid; i < AC; i += YK_SWEEP_THREADS) {
        if (i < na) {
            const uint32_t a = p.batch[a0 + i];
#pragma unroll
            for (int k = 0; k < D; ++k) sh_req[k][i] = p.a_req[(size_t)k * p.lda + a];
            sh_tol[i] = p.a_tol[a];
            sh_need[i] = p.a_need[a];
            sh_deny[i] = p.a_deny[a];
            sh_node[i] = p.a_node[a];
        } else {
#pragma unroll
            for (int k = 0; k < D; ++k) sh_req[k][i] = INT64_MAX;
            sh_tol[i] = 0; sh_need[i] = ~0ull; sh_deny[i] = ~0ull; sh_node[i] = YK_NONE_U32;
        }
        sh_first[i] = YK_NONE_U32;
    }

    // node tile -> registers (coalesced column reads: consecutive lanes = consecutive positions)
    int64_t cap[NPT][D];
    uint64_t ntaint[NPT], nlabel[NPT];
    uint32_t nidx[NPT];
    const int pos0 = blockIdx.x * (YK_SWEEP_THREADS * NPT) + tid;
#pragma unroll
    for (int j = 0; j < NPT; ++j) {
        const int pos = pos0 + j * YK_SWEEP_THREADS;
#pragma unroll
        for (int k = 0; k < D; ++k) cap[j][k] = __ldg(p.s_cap + (size_t)k * p.Np + pos);
        ntaint[j] = __ldg(p.s_taint + pos);
        nlabel[j] = __ldg(p.s_label + pos);
        nidx[j] = __ldg(p.s_node + pos);
    }
    __syncthreads();

    const int word0 = blockIdx.x * (YK_SWEEP_THREADS * NPT / 32) + warp;
    for (int ab = 0; ab < na; ab += 32) {
        uint32_t keep[NPT];
#pragma unroll
        for (int j = 0; j < NPT; ++j) keep[j] = 0;
#pragma unroll 8
        for (int l = 0; l < 32; ++l) {
            const int i = ab + l;
            int64_t rq[D];
#pragma unroll
            for (int k = 0; k < D; ++k) rq[k] = sh_req[k][i];
            const uint64_t tol = sh_tol[i], need = sh_need[i], deny = sh_deny[i];
            const uint32_t want = sh_node[i];
#pragma unroll
            for (int j = 0; j < NPT; ++j) {
                bool ok = true;
#pragma unroll
                for (int k = 0; k < D; ++k) ok = ok && (rq[k] <= cap[j][k]);
                const uint64_t bad = (ntaint[j] & ~tol) | (~nlabel[j] & need) | (nlabel[j] & deny);
                ok = ok && (bad == 0ull) && (want == YK_NONE_U32 || want == nidx[j]);
                const uint32_t w = __ballot_sync(0xFFFFFFFFu, ok);
                if (lane == l) keep[j] = w;
            }
        }
        // lane l now holds the NPT bitmap words of ask (ab + l) for this warp's positions
        const int i = ab + lane;
        if (i < na) {
            uint32_t* row = p.fit + (size_t)(a0 + i) * p.W;
            uint32_t best = YK_NONE_U32;
#pragma unroll
            for (int j = NPT - 1; j >= 0; --j) {
                const int widx = word0 + j * (YK_SWEEP_THREADS / 32);
                row[widx] = keep[j];
                if (keep[j]) best = (uint32_t)widx * 32u + (uint32_t)(__ffs((int)keep[j]) - 1);
            }
            if (best != YK_NONE_U32) atomicMin(&sh_first[i], best);
        }
    }
    __syncthreads();
    for (int i = tid; i < na; i += YK_SWEEP_THREADS)
        if (sh_first[i] != YK_NONE_U32) atomicMin(&p.first[a0 + i], sh_first[i]);
}

// ---- one (ask,node) answer in the reference's step order, on the device --------------------------------
// order: core tryNodes (schedulable, FitInNode total) -> tryNode preAllocateCheck (request > 0, available)
// -> shim Predicates: NodeName, TaintToleration/NodeUnschedulable, NodeAffinity, NodeResourcesFit
// (predicate_manager.go:339-351 filter order).
__global__ void yk_evaluate_kernel(int D, const int64_t* __restrict__ total, const int64_t* __restrict__ avail,
                                   size_t ldn, const uint64_t* __restrict__ taint, const uint64_t* __restrict__ label,
                                   const uint32_t* __restrict__ flags, const int64_t* __restrict__ a_req,
                                   const uint64_t* __restrict__ a_tol, const uint64_t* __restrict__ a_need,
                                   const uint64_t* __restrict__ a_deny, const uint32_t* __restrict__ a_node, size_t lda,
                                   uint32_t ask, uint32_t node, int* __restrict__ out) {
    if (threadIdx.x || blockIdx.x) return;
    int r = 0;
    bool pos = false, neg = false, fit_total = true, fit_avail = true;
    for (int k = 0; k < D; ++k) {
        int64_t q = a_req[(size_t)k * lda + ask];
        int64_t t = total[(size_t)k * ldn + node], a = avail[(size_t)k * ldn + node];
        if (q < 0) neg = true;
        if (q > 0) pos = true;
        if (q > (t < 0 ? 0 : t)) fit_total = false;
        if (q > (a < 0 ? 0 : a)) fit_avail = false;
    }
    if (!(flags[node] & 1u)) r = 1;
    else if (!fit_total) r = 2;
    else if (neg || !pos) r = 3;
    else if (!fit_avail) r = 4;
    else if (a_node[ask] != YK_NONE_U32 && a_node[ask] != node) r = 5;
    else if (taint[node] & ~a_tol[ask]) r = 6;
    else if ((label[node] & a_need[ask]) != a_need[ask] || (label[node] & a_deny[ask])) r = 7;
    *out = r;
}

// scatter new availability for a list of nodes (after the ordered commit)
__global__ void yk_apply_avail_kernel(int D, int64_t* __restrict__ avail, size_t ldn, const uint32_t* __restrict__ nodes,
                                      const int64_t* __restrict__ vals /*[D][n]*/, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    uint32_t node = nodes[i];
    for (int k = 0; k < D; ++k) avail[(size_t)k * ldn + node] = vals[(size_t)k * n + i];
}

__global__ void yk_fill_u32_kernel(uint32_t* __restrict__ p, uint32_t v, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}
