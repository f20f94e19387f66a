// Host-only harness around yunikorn_k8shim_b200/csrc/yk_orderer.hpp so the ordering engine can be tested
// on a CPU box: it plays the role of the device and answers "placed" (or "no node" for asks listed in
// fail[]) for every ask the orderer proposes, with rewind on failure exactly as yk_cycle does.
#include "../../yunikorn_k8shim_b200/csrc/yk_orderer.hpp"
#include <cstdint>
#include <vector>


// queue priority properties for the next run (priority.offset / priority.policy = fence); NULL = defaults.  Test
// harness convenience: keeps the long argument lists of the run functions unchanged.
static const int32_t* g_q_prio_offset = nullptr;
static const uint8_t* g_q_prio_fence = nullptr;
extern "C" void host_set_queue_priority(const int32_t* offset, const uint8_t* fence) { g_q_prio_offset = offset; g_q_prio_fence = fence; }

extern "C" int orderer_run(int D, uint32_t nA, uint32_t nP, uint32_t nQ, const int64_t* a_req /*[D][nA]*/,
                           const int32_t* a_prio, const int64_t* a_create, const uint32_t* a_app, const uint32_t* a_flags, const uint32_t* a_gang,
                           const uint32_t* p_queue, const int64_t* p_submit, const uint32_t* q_parent,
                           const int64_t* q_guar, const int64_t* q_max, int64_t* q_alloc, const uint8_t* q_sort,
                           const uint8_t* fail /*[nA] 1 = device finds no node*/, uint32_t batch, int speculate,
                           uint32_t* out_order, uint32_t* n_out, uint8_t* state_out, int* insensitive_out) {
    yk::Orderer o;
    std::vector<uint8_t> state(nA, yk::ST_PENDING), present(nP, 1);
    o.t.D = D; o.t.maxA = nA; o.t.maxP = nP; o.t.nq = nQ;
    o.t.a_req = a_req; o.t.a_prio = a_prio; o.t.a_create = a_create; o.t.a_app = a_app; o.t.a_flags = a_flags; o.t.a_gang = a_gang;
    o.t.a_state = state.data(); o.t.p_queue = p_queue; o.t.p_submit = p_submit; o.t.p_present = present.data();
    o.t.q_parent = q_parent; o.t.q_guar = q_guar; o.t.q_max = q_max; o.t.q_alloc = q_alloc; o.t.q_sort = q_sort;
    o.t.q_prio_offset = g_q_prio_offset; o.t.q_prio_fence = g_q_prio_fence;
    std::vector<uint32_t> pending(nA);
    for (uint32_t i = 0; i < nA; ++i) pending[i] = i;
    o.begin_cycle(pending);
    *insensitive_out = o.insensitive ? 1 : 0;
    std::vector<uint32_t> b;
    uint32_t n = 0;
    auto same_gang = [&](uint32_t x, uint32_t y) { return a_gang[x] != yk::NONE && a_gang[x] == a_gang[y] && a_app[x] == a_app[y]; };
    // the "device + commit": an ask fails if fail[] says so; a gang fails whole if any member does.
    // Mirrors yk_cycle's control flow: two slots, the next batch is filled (speculatively) BEFORE the current one is
    // decided; a failure in a placement-sensitive order rewinds both and drops the speculated batch.
    struct Slot { std::vector<uint32_t> asks; yk::Orderer::Snap snap; };
    Slot slot[2];
    size_t bsz = batch;
    const bool ins = o.insensitive;
    auto next_batch = [&](Slot& sl) -> int {
        sl.asks.clear();
        o.fill(bsz, (size_t)-1, sl.asks, sl.snap);
        if (o.oversize_gang) {
            if (bsz < batch) { bsz = batch; o.fill(bsz, (size_t)-1, sl.asks, sl.snap); }
            if (o.oversize_gang) return -1;
        }
        return 0;
    };
    int cur = 0;
    if (next_batch(slot[0])) return -1;
    while (!slot[cur].asks.empty()) {
        Slot& A = slot[cur];
        Slot& Nx = slot[cur ^ 1];
        Nx.asks.clear();
        if (speculate && next_batch(Nx)) return -1;
        const bool forked = speculate != 0;
        const std::vector<uint32_t>& b = A.asks;
        std::vector<uint8_t> bad(b.size(), 0);
        for (size_t i = 0; i < b.size();) {
            size_t j = i + 1;
            while (j < b.size() && same_gang(b[i], b[j])) ++j;
            bool f = false;
            for (size_t x = i; x < j; ++x) f = f || fail[b[x]];
            for (size_t x = i; x < j; ++x) bad[x] = f;
            i = j;
        }
        size_t consumed = b.size(), first_bad = b.size();
        if (!ins)
            for (size_t i = 0; i < b.size(); ++i)
                if (bad[i]) { first_bad = i; consumed = i + 1; while (consumed < b.size() && same_gang(b[i], b[consumed])) ++consumed; break; }
        bool failed = false;
        if (first_bad < b.size()) {
            o.rewind(A.snap, forked ? &Nx.snap : nullptr, b, first_bad);
            Nx.asks.clear();
            failed = true;
        }
        for (size_t i = 0; i < consumed; ++i) {
            uint32_t a = b[i];
            if (bad[i]) { if (ins) o.fail_in_place(a); continue; }
            o.confirm(a);
            out_order[n++] = a;
        }
        bsz = failed ? std::max<size_t>(1, bsz / 4) : std::min<size_t>(batch, bsz * 2);
        if (Nx.asks.empty() && next_batch(Nx)) return -1;
        cur ^= 1;
    }
    if (o.oversize_gang) return -1;
    o.finish();
    *n_out = n;
    for (uint32_t i = 0; i < nA; ++i) state_out[i] = state[i];
    return 0;
}
