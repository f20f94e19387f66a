// yk_lattice_host.hpp -- host-side companions of the lattice commit (yk_lattice.h): the per-entry meta words the kernel
// reads, and the test that decides whether a cycle may use the lattice commit at all.  No CUDA in here: csrc/yk_engine.cu
// and tests/host/lattice_shim.cpp both include it.
#pragma once
#include <cstdint>
#include <unordered_map>
#include <vector>

#include "yk_commit.hpp"
#include "yk_lattice.h"

namespace yklt {

inline bool same_request(const yk::CommitTables& t, uint32_t a, uint32_t b) {
    for (int k = 0; k < t.D; ++k)
        if (t.a_req[(size_t)k * t.lda + a] != t.a_req[(size_t)k * t.lda + b]) return false;
    return true;
}

// Dense numbers for the distinct request vectors ("shapes") of the cycle's pending asks: a_shape[ask].  Equal numbers <=>
// equal vectors (hash table, equality always confirmed on the full vector).
inline void assign_shapes(const yk::CommitTables& t, const std::vector<uint32_t>& pending, std::vector<uint32_t>& a_shape, uint32_t* n_shapes) {
    struct Slot { uint32_t ask; uint32_t id; };
    size_t cap = 64;
    while (cap < 4 * pending.size()) cap <<= 1;
    std::vector<Slot> shapes(cap, Slot{yk::CNONE, 0});
    uint32_t ns = 0;
    for (uint32_t a : pending) {
        uint64_t h = 0x9E3779B97F4A7C15ull;
        for (int k = 0; k < t.D; ++k) { h ^= (uint64_t)t.a_req[(size_t)k * t.lda + a] + 0x9E3779B97F4A7C15ull + (h << 6) + (h >> 2); h *= 0xD6E8FEB86659FD93ull; }
        size_t x = (size_t)(h ^ (h >> 29)) & (cap - 1);
        for (;;) {
            Slot& s = shapes[x];
            if (s.ask == yk::CNONE) { s.ask = a; s.id = ns++; a_shape[a] = s.id; break; }
            if (same_request(t, s.ask, a)) { a_shape[a] = s.id; break; }
            x = (x + 1) & (cap - 1);
        }
    }
    *n_shapes = ns;
}

// Dense numbers for the distinct predicate signatures of the given asks (the engine keeps these up to date as asks are
// upserted; the CPU test harness computes them here).  a_sig = the 64-bit signature hashes.
inline void assign_sigs(const yk::CommitTables& t, const uint64_t* a_sig, const std::vector<uint32_t>& pending, std::vector<uint32_t>& a_sigid) {
    struct Slot { uint32_t ask; uint32_t id; };
    size_t cap = 64;
    while (cap < 4 * pending.size()) cap <<= 1;
    std::vector<Slot> sigs(cap, Slot{yk::CNONE, 0});
    uint32_t ng = 0;
    for (uint32_t a : pending) {
        const uint64_t g = a_sig[a];
        size_t x = (size_t)(g ^ (g >> 29)) & (cap - 1);
        for (;;) {
            Slot& s = sigs[x];
            if (s.ask == yk::CNONE) { s.ask = a; s.id = ng++; a_sigid[a] = s.id; break; }
            if (a_sig[s.ask] == g && yk::same_signature(t, s.ask, a)) { a_sigid[a] = s.id; break; }
            x = (x + 1) & (cap - 1);
        }
    }
}

// the three words per batch entry that the kernel reads (yk_lattice.h): gang flags, shape number, signature number
inline void build_meta(const yk::CommitTables& t, const uint32_t* a_shape, const uint32_t* a_sigid, const std::vector<uint32_t>& batch,
                       uint32_t* meta, uint32_t* shp, uint32_t* sig) {
    const size_t B = batch.size();
    for (size_t i = 0; i < B; ++i) {
        const uint32_t a = batch[i];
        uint32_t m = 0;
        if (t.a_gang[a] != yk::CNONE) {
            m |= M_GANG;
            const uint32_t b = i ? batch[i - 1] : a;
            if (i == 0 || t.a_gang[b] != t.a_gang[a] || t.a_app[b] != t.a_app[a]) m |= M_GSTART;
        }
        meta[i] = m; shp[i] = a_shape[a]; sig[i] = a_sigid[a];
    }
}

// How many sub-run windows the kernel's staging rules (at most SMAX shapes, SIGCAP signatures, KCAP asks) cut the batch
// into at least: the engine only starts a cycle on the device commit when the windows are long enough to pay for a sub-run.
inline size_t estimate_windows(const uint32_t* shp, const uint32_t* sig, size_t B) {
    size_t windows = 0, i = 0;
    std::vector<uint32_t> seen(512, 0xFFFFFFFFu);
    while (i < B) {
        uint32_t shapes[SMAX];
        int ns = 0, rows = 0;
        std::fill(seen.begin(), seen.end(), 0xFFFFFFFFu);
        size_t j = i;
        for (; j < B && j - i < (size_t)KCAP; ++j) {
            int k = 0;
            while (k < ns && shapes[k] != shp[j]) ++k;
            if (k == ns) { if (ns == SMAX) break; shapes[ns++] = shp[j]; }
            size_t x = (sig[j] * 2654435761u) >> 23;   // 9 bits
            while (seen[x] != 0xFFFFFFFFu && seen[x] != sig[j]) x = (x + 1) & 511;
            if (seen[x] == 0xFFFFFFFFu) { if (rows == SIGCAP) break; seen[x] = sig[j]; ++rows; }
        }
        ++windows;
        i = j > i ? j : i + 1;
    }
    return windows;
}

// What the lattice commit relies on (DESIGN.md "lattice commit"):
//   fair node sort; weights >= 0 and weighted totals >= 0 (a node's key must not decrease when it is allocated to);
//   NodeID ranks unique (element order is (key, rank, j)); every gang's members request the same vector (a gang never
//   straddles two runs, so a failed gang is undone inside one sub-run).
struct Eligibility {
    bool ok = false;
    const char* why = "";
};

// ranks_unique: when the caller already knows whether the live nodes' ranks are distinct (the engine checks it where it sorts
// the nodes by rank), else nullptr; any_gang: when it already knows whether any pending ask is a gang member.
inline Eligibility eligible(const yk::CommitTables& t, uint32_t n_hi, const uint8_t* n_present, const int64_t* n_total,
                            size_t ldn, const uint32_t* n_rank, const std::vector<uint32_t>& pending, const bool* ranks_unique = nullptr,
                            const bool* any_gang = nullptr) {
    Eligibility e;
    if (t.policy != 0u) { e.why = "binpacking node sort"; return e; }
    for (int k = 0; k < t.D; ++k) if (t.w[k] < 0.0 || t.w[k] != t.w[k]) { e.why = "negative node-sort weight"; return e; }
    for (int k = 0; k < t.D; ++k) {
        if (t.w[k] == 0.0) continue;
        const int64_t* col = n_total + (size_t)k * ldn;
        for (uint32_t n = 0; n < n_hi; ++n)
            if (n_present[n] && col[n] < 0) { e.why = "negative total on a weighted resource"; return e; }
    }
    if (ranks_unique) {
        if (!*ranks_unique) { e.why = "duplicate NodeID ranks"; return e; }
    } else {
        std::vector<uint32_t> ranks;
        ranks.reserve(n_hi);
        for (uint32_t n = 0; n < n_hi; ++n) if (n_present[n]) ranks.push_back(n_rank[n]);
        std::sort(ranks.begin(), ranks.end());
        for (size_t i = 1; i < ranks.size(); ++i) if (ranks[i] == ranks[i - 1]) { e.why = "duplicate NodeID ranks"; return e; }
    }
    if (!any_gang || *any_gang) {
        std::unordered_map<uint64_t, uint32_t> first;   // (app, gang) -> first member seen
        for (uint32_t a : pending) {
            if (t.a_gang[a] == yk::CNONE) continue;
            const uint64_t key = ((uint64_t)t.a_app[a] << 32) | t.a_gang[a];
            auto it = first.find(key);
            if (it == first.end()) first.emplace(key, a);
            else if (!same_request(t, a, it->second)) { e.why = "a gang mixes request vectors"; return e; }
        }
    }
    e.ok = true;
    return e;
}

}  // namespace yklt
