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
    #ifndef YK_DIRTY_CAP
#define YK_DIRTY_CAP 64
#endif
    static constexpr uint32_t CAP = YK_DIRTY_CAP;   // a range splits when it is full

    void clear() { used_ = 0; free_.clear(); lo_.clear(); id_.clear(); count_ = 0; }
    size_t size() const { return count_; }

    void insert(const DirtyRef& x) {
        if (lo_.empty()) { lo_.push_back(0); id_.push_back(fresh()); }
        uint32_t r = range_of(x.w);
        uint32_t s = id_[r];
        Meta* m = &meta_[s];
        if (m->n == CAP && m->b > 0) {   // the walk has eaten entries off the front: reclaim that room first
            DirtyRef* v = body(s);
            memmove(v, v + m->b, sizeof(DirtyRef) * (m->n - m->b));
            m->n -= m->b; m->b = 0;
        }
        if (m->n == CAP) {   // split: put in order, move the upper half into a new range right after this one
            sort_range(s);
            const uint32_t ns = fresh();   // may move meta_ / body_
            m = &meta_[s];
            Meta& u = meta_[ns];
            u.n = CAP - CAP / 2; u.b = 0; u.sorted = 1;
            memcpy(body(ns), body(s) + CAP / 2, sizeof(DirtyRef) * u.n);
            u.max = body(ns)[u.n - 1].w;
            m->n = CAP / 2;
            m->max = body(s)[m->n - 1].w;
            lo_.insert(lo_.begin() + r + 1, body(ns)[0].w);
            id_.insert(id_.begin() + r + 1, ns);
            if (!(x.w < body(ns)[0].w)) { ++r; s = ns; m = &u; }
        }
        // append; the range is put back in order only if and when the walk reaches it.  Everything decided here comes
        // from the small per-range records (L1); the range body itself only receives one store.
        if (m->n > m->b && x.w < m->max) m->sorted = 0; else m->max = x.w;   // max may be stale-high after erases: harmless
        body(s)[m->n++] = x;
        ++count_;
    }
    // general erase (re-keying an entry the walk did not just stand on, gang roll-back)
    void erase(const DirtyRef& x) {
        const uint32_t r = range_of(x.w);
        const uint32_t s = id_[r];
        Meta& m = meta_[s];
        DirtyRef* v = body(s);
        uint32_t i = m.b;
        while (i < m.n && !(v[i] == x)) ++i;
        if (i == m.b) ++m.b;
        else {
            if (m.sorted) memmove(v + i, v + i + 1, sizeof(DirtyRef) * (m.n - i - 1));
            else v[i] = v[m.n - 1];
            --m.n;
        }
        --count_;
        if (m.n == m.b) { m.n = m.b = 0; m.sorted = 1; if (lo_.size() > 1) drop_range(r); }
    }

    // ---- ordered walk from the smallest entry ----
    struct Cursor { uint32_t r = 0; uint32_t i = 0; bool enter = true; };
    const DirtyRef* first(Cursor& c) { c.r = 0; c.i = 0; c.enter = true; return settle(c); }
    const DirtyRef* next(Cursor& c) { ++c.i; return settle(c); }
    // start the walk at the first entry >= x instead (entries before it are known to be of no interest)
    const DirtyRef* seek(Cursor& c, const DirtyRef& x) {
        if (lo_.empty()) return nullptr;
        c.r = range_of(x.w);
        const uint32_t s = id_[c.r];
        sort_range(s);
        const Meta& m = meta_[s];
        const DirtyRef* v = body(s);
        uint32_t lo = m.b, hi = m.n;
        while (lo < hi) { const uint32_t mid = (lo + hi) >> 1; if (v[mid].w < x.w) lo = mid + 1; else hi = mid; }
        c.i = lo; c.enter = false;
        return settle(c);
    }
    // take out the entry the cursor stands on
    void erase_at(const Cursor& c) {
        const uint32_t s = id_[c.r];
        Meta& m = meta_[s];
        if (c.i == m.b) ++m.b;   // the usual case: the walk takes the smallest entry -- no shifting
        else {
            DirtyRef* v = body(s);
            memmove(v + c.i, v + c.i + 1, sizeof(DirtyRef) * (m.n - c.i - 1));
            --m.n;
        }
        --count_;
        if (m.n == m.b) { m.n = m.b = 0; m.sorted = 1; if (lo_.size() > 1) drop_range(c.r); }
    }

    template <typename F>
    void for_each(F&& f) {   // all entries, ascending (puts every range in order: epoch end)
        for (uint32_t r = 0; r < lo_.size(); ++r) {
            const uint32_t s = id_[r];
            sort_range(s);
            const Meta& m = meta_[s];
            const DirtyRef* v = body(s);
            for (uint32_t i = m.b; i < m.n; ++i) f(v[i]);
        }
    }

private:
    struct Meta { unsigned __int128 max; uint32_t n, b, sorted, pad; };   // live entries of the range: body[b, n)
    DirtyRef* body(uint32_t s) { return body_.data() + (size_t)s * CAP; }
    uint32_t fresh() {
        uint32_t s;
        if (!free_.empty()) { s = free_.back(); free_.pop_back(); }   // a slot an emptied range gave back: the working set stays the live ranges
        else {
            if (used_ == meta_.size()) {   // grow geometrically: a long epoch splits thousands of times
                const size_t want = std::max<size_t>(64, meta_.size() * 2);
                meta_.resize(want); body_.resize(want * CAP);
            }
            s = used_++;
        }
        Meta& m = meta_[s];
        m.n = 0; m.b = 0; m.sorted = 1; m.max = 0;
        return s;
    }
    void drop_range(uint32_t r) {
        free_.push_back(id_[r]);
        lo_.erase(lo_.begin() + r); id_.erase(id_.begin() + r);
        lo_[0] = 0;   // the first range always starts at -inf
    }
    void sort_range(uint32_t s) {
        Meta& m = meta_[s];
        if (m.sorted) return;
        DirtyRef* v = body(s);
        // the range was in order before a few appends: insertion sort is adaptive to that
        for (uint32_t a = m.b + 1; a < m.n; ++a) {
            const DirtyRef x = v[a];
            uint32_t b = a;
            for (; b > m.b && x.w < v[b - 1].w; --b) v[b] = v[b - 1];
            v[b] = x;
        }
        m.sorted = 1;
        m.max = v[m.n - 1].w;
    }
    // last range whose lower bound is <= w.  16-ary counting search: every level issues up to 15 independent
    // compares (the bounds are sorted, so "how many samples are <= w" IS the child to descend into) instead of the 8-11
    // dependent load-compare steps of a binary search; two levels cover 256 ranges, three cover 4096.
    uint32_t range_of(unsigned __int128 w) const {
        const unsigned __int128* lo = lo_.data();
        uint32_t base = 0, len = (uint32_t)lo_.size();   // invariant: lo[base] <= w, the answer is in [base, base+len)
        while (len > 16) {
            const uint32_t step = (len + 15) >> 4;
            uint32_t c = 0;
            for (uint32_t k = 1; k < 16; ++k) {
                const uint32_t idx = base + k * step;
                c += (k * step < len && lo[idx] <= w) ? 1u : 0u;
            }
            base += c * step;
            len = std::min(step, len - c * step);
        }
        uint32_t c = 0;
        for (uint32_t k = 1; k < len; ++k) c += (lo[base + k] <= w) ? 1u : 0u;
        return base + c;
    }
    const DirtyRef* settle(Cursor& c) {
        while (c.r < lo_.size()) {
            const uint32_t s = id_[c.r];
            const Meta& m = meta_[s];
            if (c.enter) { c.i = m.b; c.enter = false; }
            if (c.i < m.n) { sort_range(s); return body(s) + c.i; }
            ++c.r; c.enter = true;
        }
        return nullptr;
    }
    std::vector<unsigned __int128> lo_;   // lower bound of each range, ascending, contiguous (the search stays in L1)
    std::vector<uint32_t> id_;            // range -> slot
    std::vector<Meta> meta_;              // per slot: counts, order flag, largest entry (small: stays in L1)
    std::vector<DirtyRef> body_;          // per slot: CAP entries; reused across epochs
    std::vector<uint32_t> free_;          // slots of ranges that ran empty, reused before new ones are taken
    uint32_t used_ = 0;
    size_t count_ = 0;
};

}  // namespace yk
