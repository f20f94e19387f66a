// Host emulation of the device-resident ordered commit for the CPU test suite: csrc/yk_lattice.h is single-source -- nvcc
// turns it into yk_lattice_kernel, g++ turns the SAME code into plain loops -- so the lattice algorithm (bound, element
// order, acceptance rows, chain, gang snapshots, capacity bound, order patch) is fuzzed against the oracle here without a
// GPU.  The driver around it is the real ordering engine (csrc/yk_orderer.hpp) and the real meta builder
// (csrc/yk_lattice_host.hpp).  Test infrastructure: the product only ever runs the kernel.
//
// Where the engine hands a batch over to its host commit (status HANDOFF: a gang that cannot be placed from the front of
// the order), this shim finishes the batch with a plain sequential first-fit over the current order instead.
#include "../../yunikorn_k8shim_b200/csrc/yk_lattice_host.hpp"
#include "../../yunikorn_k8shim_b200/csrc/yk_uniform.h"
#include "../../yunikorn_k8shim_b200/csrc/yk_orderer.hpp"

#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <memory>
#include <vector>

static const int32_t* g_q_prio_offset = nullptr;
static const uint8_t* g_q_prio_fence = nullptr;
extern "C" void host_set_queue_priority(const int32_t* offset, const uint8_t* fence) { g_q_prio_offset = offset; g_q_prio_fence = fence; }
// user / group limits for the next run (same convenience): application users, then the entries [D][n] column-major
static const uint32_t* g_p_user = nullptr; static uint32_t g_n_ul = 0;
static const uint32_t* g_ul_queue = nullptr; static const uint32_t* g_ul_user = nullptr;
static const int64_t* g_ul_max = nullptr; static int64_t* g_ul_alloc = nullptr;
extern "C" void host_set_user_limits(const uint32_t* p_user, uint32_t n, const uint32_t* queue, const uint32_t* user, const int64_t* max, int64_t* alloc) {
    g_p_user = p_user; g_n_ul = n; g_ul_queue = queue; g_ul_user = user; g_ul_max = max; g_ul_alloc = alloc;
}

// shortest uniform run that takes the uniform-run path (0: never), and how deep the first attempt generates (0: the engine's rule)
static int g_un_min = 0, g_un_depth = 0;
extern "C" void host_set_uniform(int min_run, int first_depth) { g_un_min = min_run; g_un_depth = first_depth; }

namespace {

// The uniform-run path as the engine drives it (csrc/yk_engine.cu lt_uniform), with the kernels of yk_uniform.cuh replaced by
// loops over the same per-item bodies and the two cub radix sorts by std::stable_sort.
template <int D>
int uniform_host(const yklt::Args& la, const uint32_t* asks, int off, int R, bool has_gang, uint32_t* res, size_t* consumed, uint64_t* stats) {
    const int nlive = la.nlive;
    *consumed = 0;
    if (nlive == 0 || R == 0) return ykun::U_FALLBACK;
    const uint32_t ask = asks[off];
    std::vector<uint32_t> byrank((size_t)nlive);
    {
        const yklt::Ent* e = la.ord[*la.cur & 1];
        std::vector<uint64_t> rn((size_t)nlive);
        for (int p = 0; p < nlive; ++p) rn[(size_t)p] = e[p].rn;
        std::sort(rn.begin(), rn.end());
        for (int p = 0; p < nlive; ++p) byrank[(size_t)p] = (uint32_t)rn[(size_t)p];
    }
    uint32_t max_node = 0;
    for (uint32_t n : byrank) max_node = std::max(max_node, n);
    std::vector<uint32_t> cnt((size_t)max_node + 1, 0);
    std::vector<unsigned long long> bk((size_t)nlive), rkey((size_t)nlive), rrn((size_t)nlive), okey((size_t)nlive), orn((size_t)nlive);
    ykun::Globals g;
    ykun::Args a{};
    a.policy = la.policy;
    for (int k = 0; k < 8; ++k) { a.w[k] = la.w[k]; a.req[k] = k < D ? la.a_req[(size_t)k * la.lda + ask] : 0; }
    a.rec = la.rec; a.RS = la.RS; a.ord[0] = la.ord[0]; a.ord[1] = la.ord[1]; a.cur = la.cur; a.nlive = nlive; a.byrank = byrank.data();
    a.tol = la.a_tol[ask]; a.need = la.a_need[ask]; a.deny = la.a_deny[ask]; a.want = la.a_node[ask];
    a.R = R; a.insensitive = la.insensitive; a.has_gang = has_gang ? 1 : 0;
    a.bk = bk.data(); a.cnt = cnt.data(); a.rkey = rkey.data(); a.rrn = rrn.data(); a.okey = okey.data(); a.orn = orn.data();
    a.res = res + off; a.g = &g;
    int L = g_un_depth > 0 ? g_un_depth : ykun::first_depth(R, nlive);
    for (;;) {
        if ((size_t)L * (size_t)nlive > ((size_t)4 << 20)) return ykun::U_FALLBACK;
        a.L = L;
        const size_t ne = (size_t)nlive * (size_t)L;
        std::vector<unsigned long long> ekey(ne), skey(ne);
        std::vector<uint32_t> enode(ne), snode(ne), idx(ne);
        a.ekey = ekey.data(); a.enode = enode.data(); a.skey = skey.data(); a.snode = snode.data();
        g.bkey = ykun::KEY_INF; g.brank = ykun::KEY_INF; g.last_key = 0; g.last_rank = 0; g.valid = 0; g.nan = 0; g.status = ykun::U_RETRY; g.consumed = 0;
        for (long long x = 0; x < (long long)ne; ++x) {
            const ykun::DepthOut o = ykun::element_item<D>(a, x);
            g.valid += o.valid;
            if (o.bk < g.bkey) g.bkey = o.bk;
            if (o.nan) g.nan = 1;
        }
        for (int i = 0; i < nlive; ++i) { const unsigned long long r = ykun::brank_item<D>(a, i, g.bkey); if (r < g.brank) g.brank = r; }
        for (size_t x = 0; x < ne; ++x) idx[x] = (uint32_t)x;
        std::stable_sort(idx.begin(), idx.end(), [&](uint32_t x, uint32_t y) { return ekey[x] < ekey[y]; });
        for (size_t x = 0; x < ne; ++x) { skey[x] = ekey[idx[x]]; snode[x] = enode[idx[x]]; }
        for (int q = 0; q < R; ++q) ykun::select_item<D>(a, q, g.valid);
        ykun::decide(a);
        int nan_after = 0;
        for (int i = 0; i < nlive; ++i) { ykun::apply_item<D>(a, i, g.status); nan_after |= ykun::rekey_item<D>(a, i); }
        {
            std::vector<uint32_t> ix((size_t)nlive);
            for (int i = 0; i < nlive; ++i) ix[(size_t)i] = (uint32_t)i;
            std::stable_sort(ix.begin(), ix.end(), [&](uint32_t x, uint32_t y) { return rkey[x] < rkey[y]; });
            yklt::Ent* o = la.ord[*la.cur & 1];
            for (int p = 0; p < nlive; ++p) { o[p].key = rkey[ix[(size_t)p]]; o[p].rn = rrn[ix[(size_t)p]]; }
        }
        for (uint32_t c : cnt) if (c) return -99;   // counters must be clean after every attempt
        stats[8] += 1;
        if (g.status == ykun::U_RETRY) { stats[9] += 1; L *= 4; continue; }
        if (g.status == ykun::U_NAN || g.status == ykun::U_FALLBACK) return ykun::U_FALLBACK;
        if (nan_after) return -5;
        *consumed = (size_t)g.consumed;
        stats[10] += (uint64_t)g.consumed;
        return g.status;
    }
}

template <int D>
int run_d(uint32_t policy, const double* weights,
          uint32_t nN, const int64_t* n_total, const int64_t* n_avail, const uint64_t* n_taint, const uint64_t* n_label,
          const uint32_t* n_flags, const uint32_t* n_rank,
          uint32_t nA, uint32_t nP, uint32_t nQ, const int64_t* a_req, const uint64_t* a_tol, const uint64_t* a_need,
          const uint64_t* a_deny, const uint32_t* a_node, const int32_t* a_prio, const int64_t* a_create, const uint32_t* a_app,
          const uint32_t* a_flags, const uint32_t* a_gang, const uint32_t* p_queue, const int64_t* p_submit,
          const uint32_t* q_parent, const int64_t* q_guar, const int64_t* q_max, int64_t* q_alloc, const uint8_t* q_sort,
          uint32_t batch, uint32_t max_bindings,
          uint32_t* out_ask, uint32_t* out_node, uint32_t* n_out, uint8_t* state_out, int64_t* avail_out, uint64_t* stats_out) {
    yk::Orderer o;
    std::vector<uint8_t> state(nA, yk::ST_PENDING), present(nP, 1), npresent(nN, 1);
    std::vector<int64_t> p_alloc((size_t)D * nP, 0);
    o.t.D = D; o.t.maxA = nA; o.t.maxP = nP; o.t.nq = nQ;
    o.t.a_req = a_req; o.t.a_prio = a_prio; o.t.a_create = a_create; o.t.a_app = a_app; o.t.a_flags = a_flags; o.t.a_gang = a_gang;
    // the per-ask cause table the engine keeps at upsert time (yk::Tables::a_cause): same rule, so the tests walk that branch
    std::vector<uint8_t> a_cause(nA, 0);
    for (uint32_t a = 0; a < nA; ++a) {
        bool pos = false, neg = false;
        for (int k = 0; k < D; ++k) { const int64_t v = a_req[(size_t)k * nA + a]; neg = neg || v < 0; pos = pos || v > 0; }
        a_cause[a] = (a_flags[a] & 1u) ? yk::ST_SLOWPATH : ((neg || !pos) ? yk::ST_INVALID : 0);
    }
    o.t.a_cause = a_cause.data();
    o.t.a_state = state.data(); o.t.p_queue = p_queue; o.t.p_submit = p_submit; o.t.p_present = present.data();
    o.t.q_parent = q_parent; o.t.q_guar = q_guar; o.t.q_max = q_max; o.t.q_alloc = q_alloc; o.t.p_alloc = p_alloc.data(); o.t.q_sort = q_sort;
    o.t.q_prio_offset = g_q_prio_offset; o.t.q_prio_fence = g_q_prio_fence;
    o.t.p_user = g_p_user; o.t.n_ul = g_n_ul; o.t.ul_queue = g_ul_queue; o.t.ul_user = g_ul_user; o.t.ul_max = g_ul_max; o.t.ul_alloc = g_ul_alloc;
    std::vector<uint32_t> pending(nA);
    for (uint32_t i = 0; i < nA; ++i) pending[i] = i;
    auto now_ns = [] { return (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    uint64_t t_begin = now_ns(), t_fill = 0, t_settle = 0, t_finish = 0;
    o.begin_cycle(pending);
    t_begin = now_ns() - t_begin;
    const bool ins = o.insensitive;

    yk::CommitTables ct;
    ct.D = D; ct.policy = policy; ct.w = weights; ct.lda = nA;
    ct.a_req = a_req; ct.a_tol = a_tol; ct.a_need = a_need; ct.a_deny = a_deny; ct.a_node = a_node; ct.a_gang = a_gang; ct.a_app = a_app;
    const yklt::Eligibility el = yklt::eligible(ct, nN, npresent.data(), n_total, nN, n_rank, pending);
    if (!el.ok) return 100;
    std::vector<uint64_t> a_sig(nA);
    std::vector<uint32_t> a_shape(nA, 0), a_sigid(nA, 0);
    for (uint32_t a = 0; a < nA; ++a) a_sig[a] = yk::ask_signature(ct, a);
    uint32_t n_shapes = 0;
    yklt::assign_shapes(ct, pending, a_shape, &n_shapes);
    yklt::assign_sigs(ct, a_sig.data(), pending, a_sigid);

    // ---- device state: node records, the order (what yk_key_kernel + the radix sort + yk_lt_init_kernel build) ----
    const int RS = (2 * D + 3 + 3) / 4 * 4;
    std::vector<int64_t> rec((size_t)std::max<uint32_t>(nN, 1) * RS, 0);
    std::vector<yklt::Ent> ord0(std::max<uint32_t>(nN, 1)), ord1(std::max<uint32_t>(nN, 1));
    for (uint32_t n = 0; n < nN; ++n) {
        int64_t* r = rec.data() + (size_t)n * RS;
        for (int k = 0; k < D; ++k) { r[k] = n_avail[(size_t)k * nN + n]; r[D + k] = n_total[(size_t)k * nN + n]; }
        r[2 * D] = (int64_t)n_taint[n]; r[2 * D + 1] = (int64_t)n_label[n];
        r[2 * D + 2] = (int64_t)(((uint64_t)n_rank[n] << 32) | n_flags[n]);
        const double sc = yk_node_score(D, policy, weights, n_total + n, n_avail + n, nN);
        ord0[n].key = yk_key_bits(sc); ord0[n].rn = ((uint64_t)n_rank[n] << 32) | n;
        if (ord0[n].key == YK_KEY_NAN) return -5;
    }
    auto ent_lt = [](const yklt::Ent& x, const yklt::Ent& y) { return x.key < y.key || (x.key == y.key && x.rn < y.rn); };
    std::sort(ord0.begin(), ord0.begin() + nN, ent_lt);
    int cur = 0;
    int hdr[yklt::H_WORDS] = {0, 0, 0, 0, 0, 0, 0, 0};
    int64_t ub[8];
    for (int k = 0; k < 8; ++k) ub[k] = INT64_MAX;
    std::unique_ptr<yklt::Shared<D>> sh(new yklt::Shared<D>());

    yklt::Args la{};
    la.policy = policy;
    for (int k = 0; k < 8; ++k) la.w[k] = k < D ? weights[k] : 0.0;
    la.rec = rec.data(); la.RS = RS; la.ord[0] = ord0.data(); la.ord[1] = ord1.data(); la.cur = &cur; la.nlive = (int)nN;
    la.a_req = a_req; la.lda = nA; la.a_tol = a_tol; la.a_need = a_need; la.a_deny = a_deny; la.a_node = a_node;
    la.hdr = hdr; la.ub = ub; la.insensitive = ins ? 1 : 0;

    // the order must stay exactly: every node once, ascending by (CURRENT key, rank)
    auto order_ok = [&]() -> bool {
        const yklt::Ent* e = la.ord[cur];
        std::vector<uint8_t> seen(nN, 0);
        for (uint32_t p = 0; p < nN; ++p) {
            const uint32_t n = (uint32_t)e[p].rn;
            if (n >= nN || seen[n]) return false;
            seen[n] = 1;
            const int64_t* r = rec.data() + (size_t)n * RS;
            if (e[p].key != yk_key_bits(yk_node_score(D, policy, weights, r + D, r, 1)) || (e[p].rn >> 32) != n_rank[n]) return false;
            if (p > 0 && !ent_lt(e[p - 1], e[p])) return false;
        }
        return true;
    };
    // plain sequential first fit on the current order (stands in for the engine's host commit after a hand-off)
    auto resort = [&]() {
        yklt::Ent* e = la.ord[cur];
        for (uint32_t p = 0; p < nN; ++p) {
            const uint32_t n = (uint32_t)e[p].rn;
            const int64_t* r = rec.data() + (size_t)n * RS;
            e[p].key = yk_key_bits(yk_node_score(D, policy, weights, r + D, r, 1));
        }
        std::sort(e, e + nN, ent_lt);
    };
    auto seq_place = [&](uint32_t ask) -> uint32_t {
        int64_t rq[D];
        for (int k = 0; k < D; ++k) rq[k] = a_req[(size_t)k * nA + ask];
        const yklt::Ent* e = la.ord[cur];
        for (uint32_t p = 0; p < nN; ++p) {
            const uint32_t n = (uint32_t)e[p].rn;
            int64_t* r = rec.data() + (size_t)n * RS;
            const uint32_t fl = (uint32_t)(uint64_t)r[2 * D + 2];
            if (!yklt::fits<D>((fl & 1u) && !(fl & 2u), r, r + D, rq)) continue;
            if (!yklt::accepts((uint64_t)r[2 * D], (uint64_t)r[2 * D + 1], n, a_tol[ask], a_need[ask], a_deny[ask], a_node[ask])) continue;
            for (int k = 0; k < D; ++k) r[k] -= rq[k];
            resort();
            return n;
        }
        return yk::CNONE;
    };

    std::vector<uint32_t> asks, meta, shp, sig, result;
    yk::Orderer::Snap snap;
    size_t bsz = batch;
    uint32_t n = 0;
    uint64_t handoffs = 0, batches = 0;
    uint64_t ustats[11] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};   // [8] uniform attempts, [9] deeper retries, [10] asks decided by uniform runs
    while (n < max_bindings) {
        { const uint64_t t0 = now_ns(); o.fill(bsz, (size_t)max_bindings - n, asks, snap); t_fill += now_ns() - t0; }
        if (o.oversize_gang) {
            if (bsz < batch) { bsz = batch; o.fill(bsz, (size_t)max_bindings - n, asks, snap); }
            if (o.oversize_gang) return -1;
        }
        if (asks.empty()) break;
        ++batches;
        const size_t B = asks.size();
        meta.resize(B); shp.resize(B); sig.resize(B);
        yklt::build_meta(ct, a_shape.data(), a_sigid.data(), asks, meta.data(), shp.data(), sig.data());
        result.assign(B, yk::CNONE);
        la.asks = asks.data(); la.meta = meta.data(); la.shp = shp.data(); la.sig = sig.data(); la.B = (int)B; la.res = result.data();
        size_t consumed = 0;
        if (nN == 0) {
            consumed = ins ? B : 1;
            if (!ins && a_gang[asks[0]] != yk::CNONE)
                while (consumed < B && a_gang[asks[consumed]] == a_gang[asks[0]] && a_app[asks[consumed]] == a_app[asks[0]]) ++consumed;
        } else {
            // the engine's lt_batch: uniform runs through the uniform-run path, the stretches between them through the kernel
            std::vector<ykun::Segment> segs;
            ykun::plan_segments(meta.data(), shp.data(), sig.data(), (int)B, g_un_min > 0 ? g_un_min : (int)B + 1, segs);
            int status = yklt::ST_DONE;
            for (const ykun::Segment& sg : segs) {
                size_t cons = 0;
                int st = yklt::ST_DONE;
                bool windowed = !sg.uniform;
                if (sg.uniform) {
                    bool has_gang = false;
                    for (int i = 0; i < sg.len; ++i) has_gang = has_gang || (meta[(size_t)sg.off + i] & yklt::M_GANG);
                    const int ust = uniform_host<D>(la, asks.data(), sg.off, sg.len, has_gang, result.data(), &cons, ustats);
                    if (ust < 0) return ust;
                    if (ust == ykun::U_FALLBACK) windowed = true;
                    else st = ust == ykun::U_STOPPED ? yklt::ST_STOPPED : yklt::ST_DONE;
                }
                if (windowed) {
                    la.asks = asks.data() + sg.off; la.meta = meta.data() + sg.off; la.shp = shp.data() + sg.off; la.sig = sig.data() + sg.off;
                    la.B = sg.len; la.res = result.data() + sg.off;
                    yklt::lattice_batch<D>(la, *sh);
                    st = hdr[yklt::H_STATUS];
                    cons = (size_t)hdr[yklt::H_CONSUMED];
                }
                if (st == yklt::ST_NAN) return -5;
                if (!order_ok()) return -9;
                consumed = (size_t)sg.off + cons;
                status = st;
                if (st != yklt::ST_DONE || cons < (size_t)sg.len) break;
            }
            hdr[yklt::H_STATUS] = status;
            hdr[yklt::H_CONSUMED] = (int)consumed;
            if (hdr[yklt::H_STATUS] == yklt::ST_HANDOFF) {
                ++handoffs;
                bool stop = false;
                size_t i = consumed;
                while (i < B && !stop) {
                    const uint32_t a = asks[i];
                    if (a_gang[a] == yk::CNONE) {
                        result[i] = seq_place(a);
                        if (result[i] == yk::CNONE && !ins) stop = true;
                        ++i;
                        continue;
                    }
                    size_t g1 = i;
                    while (g1 < B && a_gang[asks[g1]] == a_gang[a] && a_app[asks[g1]] == a_app[a]) ++g1;
                    std::vector<int64_t> keep(rec);
                    bool okg = true;
                    for (size_t x = i; x < g1 && okg; ++x) { result[x] = seq_place(asks[x]); okg = result[x] != yk::CNONE; }
                    if (!okg) {
                        rec = keep; la.rec = rec.data();
                        resort();
                        for (size_t x = i; x < g1; ++x) result[x] = yk::CNONE;
                        if (!ins) stop = true;
                    }
                    i = g1;
                }
                consumed = i;
                for (int k = 0; k < 8; ++k) ub[k] = INT64_MAX;   // a roll-back may have handed capacity back
            }
        }
        bool failed = false;
        if (!ins && consumed > 0 && result[consumed - 1] == yk::CNONE) {
            size_t j = consumed - 1;
            while (j > 0 && a_gang[asks[j]] != yk::CNONE && a_gang[asks[j - 1]] == a_gang[asks[j]] &&
                   a_app[asks[j - 1]] == a_app[asks[j]] && result[j - 1] == yk::CNONE) --j;
            o.rewind(snap, nullptr, asks, j);
            failed = true;
        } else if (!ins && consumed < B) {
            return -12;   // a placement-sensitive batch may only end early on a failure
        }
        const uint64_t t_s0 = now_ns();
        for (size_t i = 0; i < consumed; ++i) {
            const uint32_t a = asks[i];
            if (result[i] == yk::CNONE) { if (ins) o.fail_in_place(a); continue; }
            o.confirm(a);
            out_ask[n] = a; out_node[n] = result[i]; ++n;
        }
        t_settle += now_ns() - t_s0;
        if (ins && consumed < B) return -13;
        bsz = failed ? std::max<size_t>(std::min<size_t>(64, batch), bsz / 4) : std::min<size_t>(batch, bsz * 2);
    }
    t_finish = now_ns();
    o.finish();
    t_finish = now_ns() - t_finish;
    *n_out = n;
    for (uint32_t i = 0; i < nA; ++i) state_out[i] = state[i];
    for (uint32_t nn = 0; nn < nN; ++nn)
        for (int k = 0; k < D; ++k) avail_out[(size_t)k * nN + nn] = rec[(size_t)nn * RS + k];
    if (stats_out) {
        stats_out[0] = (uint64_t)hdr[yklt::H_SUBRUNS]; stats_out[1] = (uint64_t)hdr[yklt::H_FULLSCANS]; stats_out[2] = (uint64_t)hdr[yklt::H_SORTS];
        stats_out[3] = (uint64_t)hdr[yklt::H_ELEMS]; stats_out[4] = (uint64_t)hdr[yklt::H_QUICK]; stats_out[5] = (uint64_t)hdr[yklt::H_ESC];
        stats_out[6] = handoffs; stats_out[7] = batches;
        stats_out[8] = ustats[8]; stats_out[9] = ustats[9]; stats_out[10] = ustats[10];
        stats_out[11] = t_begin; stats_out[12] = t_fill; stats_out[13] = t_settle; stats_out[14] = t_finish;   // ns (development aid)
    }
    return 0;
}

}  // namespace

// same argument list as engine_host_run (tests/host/engine_shim.cpp) so that the Python harness is shared; epoch_limit,
// speculate and share_rows do not apply.  Returns 100 when the snapshot is not eligible for the lattice commit.
extern "C" int lattice_host_run(int D, uint32_t policy, const double* weights,
    uint32_t nN, const int64_t* n_total, const int64_t* n_avail, const uint64_t* n_taint, const uint64_t* n_label,
    const uint32_t* n_flags, const uint32_t* n_rank,
    uint32_t nA, uint32_t nP, uint32_t nQ, const int64_t* a_req, const uint64_t* a_tol, const uint64_t* a_need,
    const uint64_t* a_deny, const uint32_t* a_node, const int32_t* a_prio, const int64_t* a_create, const uint32_t* a_app,
    const uint32_t* a_flags, const uint32_t* a_gang, const uint32_t* p_queue, const int64_t* p_submit,
    const uint32_t* q_parent, const int64_t* q_guar, const int64_t* q_max, int64_t* q_alloc, const uint8_t* q_sort,
    uint32_t batch, uint32_t /*epoch_limit*/, int /*speculate*/, int /*share_rows*/, uint32_t max_bindings,
    uint32_t* out_ask, uint32_t* out_node, uint32_t* n_out, uint8_t* state_out, int64_t* avail_out, uint64_t* stats_out) {
#define RUN(DD) case DD: return run_d<DD>(policy, weights, nN, n_total, n_avail, n_taint, n_label, n_flags, n_rank, nA, nP, nQ, a_req, a_tol, a_need, a_deny, a_node, a_prio, a_create, a_app, a_flags, a_gang, p_queue, p_submit, q_parent, q_guar, q_max, q_alloc, q_sort, batch, max_bindings, out_ask, out_node, n_out, state_out, avail_out, stats_out);
    switch (D) { RUN(1) RUN(2) RUN(3) RUN(4) RUN(5) RUN(6) RUN(7) RUN(8) default: return -2; }
#undef RUN
}

// ykun::plan_segments / cap_of_d / first_depth for the unit tests
extern "C" int host_plan_segments(const uint32_t* meta, const uint32_t* shp, const uint32_t* sig, int B, int min_run, int* off, int* len, int* uniform, int cap) {
    std::vector<ykun::Segment> segs;
    ykun::plan_segments(meta, shp, sig, B, min_run, segs);
    int n = 0;
    for (const ykun::Segment& s : segs) { if (n < cap) { off[n] = s.off; len[n] = s.len; uniform[n] = s.uniform ? 1 : 0; } ++n; }
    return n;
}
extern "C" int64_t host_cap_of(int D, int usable, const int64_t* avail, const int64_t* total, const int64_t* req, int64_t limit) {
    return ykun::cap_of_d(D, usable != 0, avail, total, req, limit);
}
extern "C" int host_first_depth(int R, int nlive) { return ykun::first_depth(R, nlive); }
