// yk_podres.cpp -- request-vector builder behind include/ykgpu_pod.h (host code, part of libykgpu.so).
//
// Restates, with identical results, the integer arithmetic the shim applies to a pod before the core ever sees it:
//   /root/reference/pkg/common/resource.go:56-109   GetPodResource
//                                         :111-127  computeContainerResource
//                                         :145-162  updateMax
//                                         :164-182  checkInitContainerRequest
//                                         :188-195  GetNodeResource
//                                         :273-301  getResource / getPodLevelResource
//                                         :331-351  Add
// and the text form of k8s quantities [EXT k8s.io/apimachinery pkg/api/resource quantity.go, suffix.go].
// Nothing here is copied: the reference works on protobuf maps of *si.Quantity, this works on one sorted map of int64.
#include <cstdint>
#include <cstring>
#include <map>
#include <new>
#include <string>
#include <vector>

#include "../../include/ykgpu.h"
#include "../../include/ykgpu_pod.h"

namespace {

using u128 = unsigned __int128;
using ResMap = std::map<std::string, int64_t>;   // resource name (scheduler vocabulary) -> amount

constexpr int64_t I64_MAX = INT64_MAX;

bool mul_ok(u128& x, u128 m) {
    if (x != 0 && m > (~(u128)0) / x) return false;
    x *= m;
    return true;
}

// ceil(mant * 2^bexp * 10^e10) for mant >= 0; false when it does not fit int64
bool scaled_ceil(u128 mant, int bexp, int e10, int64_t* out) {
    if (mant == 0) { *out = 0; return true; }
    u128 num = mant, den = 1;
    // cancel what can be cancelled first so that intermediate products stay small
    while (e10 < 0 && num % 10 == 0) { num /= 10; ++e10; }
    int neg = e10 < 0 ? -e10 : 0, pos = e10 > 0 ? e10 : 0;
    int twos = bexp < neg ? bexp : neg;   // 10^neg = 2^neg * 5^neg
    bexp -= twos;
    for (int i = 0; i < neg; ++i) if (!mul_ok(den, 5)) return false;
    for (int i = 0; i < neg - twos; ++i) if (!mul_ok(den, 2)) return false;
    for (int i = 0; i < bexp; ++i) if (!mul_ok(num, 2)) return false;
    for (int i = 0; i < pos; ++i) if (!mul_ok(num, 10)) return false;
    u128 q = num / den;
    if (num % den != 0) ++q;   // away from zero
    if (q > (u128)I64_MAX) return false;
    *out = (int64_t)q;
    return true;
}

struct Parsed { bool neg = false; u128 mant = 0; int frac = 0; int bexp = 0; int e10 = 0; };

// <sign><digits>[.<digits>]<suffix>; at least one digit; suffix from the k8s tables or e/E<signed int>
int parse_quantity(const char* s, Parsed& q) {
    if (!s || !*s) return YK_ERR_ARG;
    const char* p = s;
    if (*p == '+' || *p == '-') { q.neg = *p == '-'; ++p; }
    int digits = 0;
    for (; *p >= '0' && *p <= '9'; ++p) {
        if (q.mant > ((~(u128)0) - 9) / 10) return YK_ERR_RANGE;
        q.mant = q.mant * 10 + (u128)(*p - '0');
        ++digits;
    }
    if (*p == '.') {
        ++p;
        for (; *p >= '0' && *p <= '9'; ++p) {
            if (q.mant > ((~(u128)0) - 9) / 10) return YK_ERR_RANGE;
            q.mant = q.mant * 10 + (u128)(*p - '0');
            ++q.frac;
            ++digits;
        }
    }
    if (digits == 0) return YK_ERR_ARG;
    const std::string suf(p);
    static const struct { const char* s; int bexp; int e10; } table[] = {
        {"", 0, 0},    {"Ki", 10, 0}, {"Mi", 20, 0}, {"Gi", 30, 0}, {"Ti", 40, 0}, {"Pi", 50, 0}, {"Ei", 60, 0},
        {"n", 0, -9},  {"u", 0, -6},  {"m", 0, -3},  {"k", 0, 3},   {"M", 0, 6},   {"G", 0, 9},   {"T", 0, 12},
        {"P", 0, 15},  {"E", 0, 18}};
    for (const auto& t : table)
        if (suf == t.s) { q.bexp = t.bexp; q.e10 = t.e10; return YK_OK; }
    if (suf.size() > 1 && (suf[0] == 'e' || suf[0] == 'E')) {
        const char* e = suf.c_str() + 1;
        bool eneg = false;
        if (*e == '+' || *e == '-') { eneg = *e == '-'; ++e; }
        if (!*e) return YK_ERR_ARG;
        long v = 0;
        for (; *e; ++e) {
            if (*e < '0' || *e > '9') return YK_ERR_ARG;
            v = v * 10 + (*e - '0');
            if (v > 100000) return YK_ERR_RANGE;
        }
        q.e10 = (int)(eneg ? -v : v);
        return YK_OK;
    }
    return YK_ERR_ARG;
}

int quantity_amounts(const char* text, int64_t* value, int64_t* milli) {
    Parsed q;
    const int rc = parse_quantity(text, q);
    if (rc) return rc;
    const int e = q.e10 - q.frac;
    int64_t v = 0, m = 0;
    // a scaled value that does not fit int64 saturates, as Quantity.ScaledValue does [EXT, unpinned by the reference's tests]
    if (e > 60 || !scaled_ceil(q.mant, q.bexp, e, &v)) v = q.mant == 0 ? 0 : I64_MAX;
    if (e > 57 || !scaled_ceil(q.mant, q.bexp, e + 3, &m)) m = q.mant == 0 ? 0 : I64_MAX;
    if (value) *value = q.neg ? -v : v;
    if (milli) *milli = q.neg ? -m : m;
    return YK_OK;
}

// getResource (resource.go:273-285): cpu -> vcore in milli units, everything else by name in units
const char* scheduler_name(const char* resource) { return strcmp(resource, "cpu") == 0 ? "vcore" : resource; }

void add_into(ResMap& left, const ResMap& right) {            // Add (resource.go:331-351): union, summed
    for (const auto& kv : right) {   // Go's int64 addition wraps; keep that defined here too
        int64_t& l = left[kv.first];
        l = (int64_t)((uint64_t)l + (uint64_t)kv.second);
    }
}
void max_into(ResMap& left, const ResMap& right) {            // updateMax (resource.go:145-162): union, larger value wins
    for (const auto& kv : right) {
        auto it = left.find(kv.first);
        if (it == left.end()) left.emplace(kv.first, kv.second);
        else if (kv.second > it->second) it->second = kv.second;
    }
}

struct Container {
    std::string name;
    uint32_t kind = YK_CONTAINER;
    ResMap requests, allocated, status_requests;
};
struct Status { bool has_resources = false; };

}  // namespace

struct yk_podspec {
    std::vector<Container> containers;
    std::map<std::string, Status> statuses;   // ContainerStatuses + InitContainerStatuses, by name (resource.go:66-72)
    ResMap pod_requests, overhead;
    bool has_pod_requests = false, has_overhead = false, infeasible = false;
    ResMap result;
    std::vector<const std::string*> order;     // result entries by name (std::map order), for yk_podspec_result
    bool computed = false;
};

namespace {

// computeContainerResource (resource.go:111-127)
ResMap container_resource(const yk_podspec& p, const Container& c) {
    ResMap combined = c.requests;
    auto st = p.statuses.find(c.name);
    if (st != p.statuses.end()) {
        if (p.infeasible && st->second.has_resources) return c.status_requests;   // resize denied: the status is the truth
        max_into(combined, c.allocated);
        if (st->second.has_resources) max_into(combined, c.status_requests);
    }
    return combined;
}

bool pod_level_supported(const std::string& scheduler_resource) {
    // helpers.IsSupportedPodLevelResource [EXT k8s.io/component-helpers/resource v0.34]: cpu, memory, hugepages-*
    return scheduler_resource == "vcore" || scheduler_resource == "memory" || scheduler_resource.rfind("hugepages-", 0) == 0;
}

}  // namespace

extern "C" {

int yk_quantity_parse(const char* text, int64_t* value, int64_t* milli) { return quantity_amounts(text, value, milli); }

yk_podspec* yk_podspec_new(void) { return new (std::nothrow) yk_podspec(); }
void yk_podspec_free(yk_podspec* p) { delete p; }
void yk_podspec_clear(yk_podspec* p) {
    if (p) *p = yk_podspec();
}

int yk_podspec_container(yk_podspec* p, const char* name, uint32_t kind) {
    if (!p || !name || kind > YK_INIT_SIDECAR) return YK_ERR_ARG;
    Container c;
    c.name = name;
    c.kind = kind;
    p->containers.push_back(std::move(c));
    p->computed = false;
    return (int)p->containers.size() - 1;
}

int yk_podspec_status(yk_podspec* p, const char* container_name, int32_t has_resources) {
    if (!p || !container_name) return YK_ERR_ARG;
    p->statuses[container_name].has_resources = has_resources != 0;
    p->computed = false;
    return YK_OK;
}

int yk_podspec_amount(yk_podspec* p, uint32_t list, int32_t container, const char* resource, int64_t value, int64_t milli) {
    if (!p || !resource || list > YK_LIST_OVERHEAD) return YK_ERR_ARG;
    const bool cpu = strcmp(resource, "cpu") == 0;
    const int64_t amount = cpu ? milli : value;
    const std::string name = scheduler_name(resource);
    ResMap* dst = nullptr;
    if (list == YK_LIST_POD_REQUESTS) { dst = &p->pod_requests; p->has_pod_requests = true; }
    else if (list == YK_LIST_OVERHEAD) { dst = &p->overhead; p->has_overhead = true; }
    else {
        if (container < 0 || (size_t)container >= p->containers.size()) return YK_ERR_ARG;
        Container& c = p->containers[(size_t)container];
        dst = list == YK_LIST_REQUESTS ? &c.requests : list == YK_LIST_ALLOCATED ? &c.allocated : &c.status_requests;
    }
    (*dst)[name] = amount;   // a resource list is a map: a repeated name replaces
    p->computed = false;
    return YK_OK;
}

int yk_podspec_quantity(yk_podspec* p, uint32_t list, int32_t container, const char* resource, const char* quantity) {
    int64_t v = 0, m = 0;
    const int rc = quantity_amounts(quantity, &v, &m);
    if (rc) return rc;
    return yk_podspec_amount(p, list, container, resource, v, m);
}

int yk_podspec_resize_infeasible(yk_podspec* p, int32_t yes) {
    if (!p) return YK_ERR_ARG;
    p->infeasible = yes != 0;
    p->computed = false;
    return YK_OK;
}

int yk_podspec_compute(yk_podspec* p) {
    if (!p) return YK_ERR_ARG;
    ResMap pod;
    pod["pods"] = 1;                                                       // resource.go:59-61
    for (const Container& c : p->containers)                               // :75-77
        if (c.kind == YK_CONTAINER) add_into(pod, container_resource(*p, c));
    bool any_init = false;
    for (const Container& c : p->containers) any_init = any_init || c.kind != YK_CONTAINER;
    if (any_init) {                                                        // checkInitContainerRequest :164-182
        ResMap init_max, sidecars;
        for (const Container& c : p->containers) {
            if (c.kind == YK_CONTAINER) continue;
            const ResMap own = container_resource(*p, c);
            ResMap current = own;
            add_into(current, sidecars);                                   // plus the sidecars already running
            if (c.kind == YK_INIT_SIDECAR) add_into(sidecars, own);        // it keeps running
            max_into(init_max, current);
        }
        add_into(pod, sidecars);
        max_into(pod, init_max);
    }
    if (p->has_pod_requests && !p->pod_requests.empty())                   // :87-93 pod-level requests override cpu / memory
        for (const auto& kv : p->pod_requests)
            if (pod_level_supported(kv.first)) pod[kv.first] = kv.second;
    if (p->has_overhead) add_into(pod, p->overhead);                       // :97-106
    p->result.swap(pod);
    p->order.clear();
    for (const auto& kv : p->result) p->order.push_back(&kv.first);
    p->computed = true;
    return (int)p->result.size();
}

int yk_podspec_result(const yk_podspec* p, uint32_t i, const char** name, int64_t* value) {
    if (!p || !p->computed || i >= p->order.size()) return YK_ERR_ARG;
    if (name) *name = p->order[i]->c_str();
    if (value) *value = p->result.at(*p->order[i]);
    return YK_OK;
}

int yk_podspec_vector(const yk_podspec* p, uint32_t n_dims, const char* const* dim_names, int64_t* out, uint32_t* n_unmapped) {
    if (!p || !p->computed || (n_dims && (!dim_names || !out))) return YK_ERR_ARG;
    uint32_t mapped = 0;
    for (uint32_t k = 0; k < n_dims; ++k) {
        if (!dim_names[k]) return YK_ERR_ARG;
        auto it = p->result.find(dim_names[k]);
        out[k] = it == p->result.end() ? 0 : it->second;
        if (it != p->result.end()) ++mapped;
    }
    if (n_unmapped) *n_unmapped = (uint32_t)p->result.size() - mapped;
    return YK_OK;
}

int yk_node_quantity(const char* resource, const char* quantity, const char** name_out, int64_t* amount) {
    if (!resource || !quantity) return YK_ERR_ARG;
    int64_t v = 0, m = 0;
    const int rc = quantity_amounts(quantity, &v, &m);
    if (rc) return rc;
    const bool cpu = strcmp(resource, "cpu") == 0;
    if (name_out) *name_out = cpu ? "vcore" : resource;
    if (amount) *amount = cpu ? m : v;
    return YK_OK;
}

}  // extern "C"
