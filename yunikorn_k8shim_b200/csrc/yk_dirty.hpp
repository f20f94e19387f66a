// yk_dirty.hpp -- the ordered set of nodes re-scored during the current epoch ("dirty" nodes), host side.
//
// What the ordered commit needs from it, per ask: walk the entries from the smallest (score key, NodeID rank)
// upwards until one fits or the bound is passed (usually one or two entries), take the chosen one out, and put a
// node back under its new key -- which, after a commit, lands at an essentially random place further back.  A
// sorted container pays a search + shift in cold memory for that insert; here the entries are kept in key RANGES
// (split when they pass CAP entries, like B-tree leaves) whose lower bounds sit in one contiguous array: an insert is
// "find the range (branchless search that stays in L1) + append", and a range is put in order only when the walk
// actually reaches it (adaptive insertion sort: it was in order before the last few appends).  The order seen by the
// walk and by for_each() is exactly ascending (key, rank).
#pragma once
#include <algorithm>
#include <cstdint>
#include <cstring>
#include <vector>

namespace yk {

struct DirtyRef {
    unsigned __int128 w;   // key << 64 | rank << 32 | node : ONE unsigned compare orders by (key, rank)
    DirtyRef() : w(0) {}
    DirtyRef(uint64_t key, uint32_t rank, uint32_t node) : w(((unsigned __int128)key << 64) | ((uint64_t)rank << 32) | node) {}
    uint64_t key() const { return (uint64_t)(w >> 64); }
    uint32_t rank() const { return (uint32_t)((uint64_t)w >> 32); }
    uint32_t node() const { return (uint32_t)(uint64_t)w; }
    bool operator<(const DirtyRef& o) const { return w < o.w; }
    bool operator==(const DirtyRef& o) const { return w == o.w; }
};

class DirtyIndex {
public:
    static constexpr uint32_t CAP = 64;   // a range splits when it is full

    void clear() { used_ = 0; lo_.clear(); id_.clear(); count_ = 0; }
    size_t size() const { return count_; }

    void insert(const DirtyRef& x) {
        if (lo_.empty()) { lo_.push_back(0); id_.push_back(fresh()); }
        uint32_t r = range_of(x.w);
        Range* g = &pool_[id_[r]];
        if (g->n == CAP) {   // split: put in order, move the upper half into a new range right after this one
            sort_range(*g);
            const uint32_t nid = fresh();
            g = &pool_[id_[r]];
            Range& u = pool_[nid];
            u.n = CAP - CAP / 2; u.sorted = 1;
            memcpy(u.v, g->v + CAP / 2, sizeof(DirtyRef) * u.n);
            g->n = CAP / 2;
            lo_.insert(lo_.begin() + r + 1, u.v[0].w);
            id_.insert(id_.begin() + r + 1, nid);
            if (!(x.w < u.v[0].w)) { ++r; g = &u; }
        }
        // append; the range is put back in order only if and when the walk reaches it
        if (g->sorted && g->n && x.w < g->v[g->n - 1].w) g->sorted = 0;
        g->v[g->n++] = x;
        ++count_;
    }
    // general erase (re-keying an entry the walk did not just stand on, gang roll-back)
    void erase(const DirtyRef& x) {
        const uint32_t r = range_of(x.w);
        Range& g = pool_[id_[r]];
        uint32_t i = 0;
        while (i < g.n && !(g.v[i] == x)) ++i;
        if (g.sorted) memmove(g.v + i, g.v + i + 1, sizeof(DirtyRef) * (g.n - i - 1));
        else g.v[i] = g.v[g.n - 1];
        --g.n;
        --count_;
        if (g.n == 0 && lo_.size() > 1) drop_range(r);
    }

    // ---- ordered walk from the smallest entry ----
    struct Cursor { uint32_t r = 0; uint32_t i = 0; };
    const DirtyRef* first(Cursor& c) { c.r = 0; c.i = 0; return settle(c); }
    const DirtyRef* next(Cursor& c) { ++c.i; return settle(c); }
    // take out the entry the cursor stands on
    void erase_at(const Cursor& c) {
        Range& g = pool_[id_[c.r]];
        memmove(g.v + c.i, g.v + c.i + 1, sizeof(DirtyRef) * (g.n - c.i - 1));
        --g.n;
        --count_;
        if (g.n == 0 && lo_.size() > 1) drop_range(c.r);
    }

    template <typename F>
    void for_each(F&& f) {   // all entries, ascending (puts every range in order: epoch end)
        for (uint32_t r = 0; r < lo_.size(); ++r) {
            Range& g = pool_[id_[r]];
            sort_range(g);
            for (uint32_t i = 0; i < g.n; ++i) f(g.v[i]);
        }
    }

private:
    struct Range { uint32_t n; uint32_t sorted; uint64_t pad; DirtyRef v[CAP]; };   // storage inline: no pointer chase
    uint32_t fresh() {
        if (used_ == pool_.size()) pool_.emplace_back();
        pool_[used_].n = 0; pool_[used_].sorted = 1;
        return used_++;
    }
    void drop_range(uint32_t r) {
        lo_.erase(lo_.begin() + r); id_.erase(id_.begin() + r);
        lo_[0] = 0;   // the first range always starts at -inf
    }
    static void sort_range(Range& g) {
        if (g.sorted) return;
        // the range was in order before a few appends: insertion sort is adaptive to that
        for (uint32_t a = 1; a < g.n; ++a) {
            const DirtyRef x = g.v[a];
            uint32_t b = a;
            for (; b > 0 && x.w < g.v[b - 1].w; --b) g.v[b] = g.v[b - 1];
            g.v[b] = x;
        }
        g.sorted = 1;
    }
    uint32_t range_of(unsigned __int128 w) const {   // last range whose lower bound is <= w (branchless)
        const unsigned __int128* base = lo_.data();
        uint32_t len = (uint32_t)lo_.size();
        while (len > 1) {
            const uint32_t half = len >> 1;
            base += (base[half] <= w) ? half : 0;
            len -= half;
        }
        return (uint32_t)(base - lo_.data());
    }
    const DirtyRef* settle(Cursor& c) {
        while (c.r < lo_.size()) {
            Range& g = pool_[id_[c.r]];
            if (c.i < g.n) { sort_range(g); return &g.v[c.i]; }
            ++c.r; c.i = 0;
        }
        return nullptr;
    }
    std::vector<unsigned __int128> lo_;   // lower bound of each range, ascending, contiguous (the search stays in L1)
    std::vector<uint32_t> id_;            // range -> slot in pool_
    std::vector<Range> pool_;             // range storage, reused across epochs
    uint32_t used_ = 0;
    size_t count_ = 0;
};

}  // namespace yk
