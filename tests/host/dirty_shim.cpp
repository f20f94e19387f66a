// Host-only model check of csrc/yk_dirty.hpp: random insert / erase / walk sequences against std::set.
#include "../../yunikorn_k8shim_b200/csrc/yk_dirty.hpp"
#include <cstdint>
#include <random>
#include <set>
extern "C" int dirty_model_check(uint64_t seed, int n_keys, int ops, int key_range) {
    std::mt19937_64 r(seed);
    std::vector<uint64_t> keys((size_t)n_keys);
    for (auto& k : keys) k = r() % (uint64_t)key_range;
    std::sort(keys.begin(), keys.end());
    yk::DirtyIndex d;
    d.clear();
    std::set<unsigned __int128> model;
    std::vector<yk::DirtyRef> live;
    for (int t = 0; t < ops; ++t) {
        const int op = (int)(r() % 10);
        if (op < 5 || live.empty()) {   // insert a new (key, rank): ranks unique by construction
            yk::DirtyRef x(r() % (uint64_t)(key_range + 3), (uint32_t)t, (uint32_t)t);
            d.insert(x); model.insert(x.w); live.push_back(x);
        } else if (op < 7) {            // general erase of a random live entry
            size_t i = r() % live.size();
            d.erase(live[i]); model.erase(live[i].w); live[i] = live.back(); live.pop_back();
        } else if (op == 7) {           // seek to a random point, walk a few entries, compare with lower_bound
            yk::DirtyRef x(r() % (uint64_t)(key_range + 3), (uint32_t)(r() % (uint64_t)(t + 1)), 0);
            yk::DirtyIndex::Cursor c;
            const yk::DirtyRef* e = d.seek(c, x);
            auto it = model.lower_bound(x.w);
            int k = (int)(r() % 6);
            for (int s = 0; s <= k; ++s) {
                if ((e == nullptr) != (it == model.end())) return -6;
                if (!e) break;
                if (*it != e->w) return -7;
                e = d.next(c); ++it;
            }
        } else {                        // walk k entries from the front, compare with the model, maybe erase where we stand
            yk::DirtyIndex::Cursor c;
            const yk::DirtyRef* e = d.first(c);
            auto it = model.begin();
            int k = (int)(r() % 12);
            for (int s = 0; s < k && e; ++s) { if (it == model.end() || *it != e->w) return -1; e = d.next(c); ++it; }
            if ((e == nullptr) != (it == model.end())) return -2;
            if (e && (r() & 1)) {
                if (*it != e->w) return -3;
                const unsigned __int128 w = e->w;
                d.erase_at(c); model.erase(w);
                for (size_t i = 0; i < live.size(); ++i) if (live[i].w == w) { live[i] = live.back(); live.pop_back(); break; }
            }
        }
        if (d.size() != model.size()) return -4;
    }
    auto it = model.begin();
    int bad = 0;
    d.for_each([&](const yk::DirtyRef& x) { if (it == model.end() || *it != x.w) bad = 1; else ++it; });
    if (bad || it != model.end()) return -5;
    return 0;
}

// The commit's own pattern, for long: take an entry near the front out (erase_at) and put it back further behind under a larger
// key.  Front ranges run empty and are dropped all the time; their slots are recycled by later splits.
extern "C" int dirty_model_drain(uint64_t seed, int n, int ops, int max_step) {
    std::mt19937_64 r(seed);
    yk::DirtyIndex d;
    d.clear();
    std::set<unsigned __int128> model;
    std::vector<uint64_t> key((size_t)n);
    for (int i = 0; i < n; ++i) { key[(size_t)i] = (uint64_t)i * 7; yk::DirtyRef x(key[(size_t)i], (uint32_t)i, (uint32_t)i); d.insert(x); model.insert(x.w); }
    for (int t = 0; t < ops; ++t) {
        yk::DirtyIndex::Cursor c;
        const yk::DirtyRef* e = d.first(c);
        auto it = model.begin();
        const int skip = (int)(r() % 4);
        for (int s = 0; s < skip && e; ++s) { if (*it != e->w) return -1; e = d.next(c); ++it; }
        if (!e || it == model.end() || *it != e->w) return -2;
        const uint32_t node = e->node();
        d.erase_at(c); model.erase(it);
        key[node] += 1 + r() % (uint64_t)max_step;
        yk::DirtyRef x(key[node], node, node);
        d.insert(x); model.insert(x.w);
        if (d.size() != model.size()) return -3;
        if ((t & 1023) == 0) {   // full order check now and then
            auto jt = model.begin();
            int bad = 0;
            d.for_each([&](const yk::DirtyRef& y) { if (jt == model.end() || *jt != y.w) bad = 1; else ++jt; });
            if (bad || jt != model.end()) return -4;
        }
    }
    auto jt = model.begin();
    int bad = 0;
    d.for_each([&](const yk::DirtyRef& y) { if (jt == model.end() || *jt != y.w) bad = 1; else ++jt; });
    return (bad || jt != model.end()) ? -5 : 0;
}
