"""ctypes stub of the request-vector builder (include/ykgpu_pod.h, csrc/yk_podres.cpp): pod containers -> int64 request
vector, node allocatable -> int64 total vector.  Host code inside libykgpu.so; needs no GPU.

A pod is described the way the reference's GetPodResource reads it (/root/reference/pkg/common/resource.go:56-109):

    {"containers":      [{"name": "c1", "requests": {"cpu": "500m", "memory": "1Gi"}}, ...],
     "initContainers":  [{"name": "i1", "requests": {...}, "restartPolicy": "Always"}, ...],
     "statuses":        [{"name": "c1", "allocated": {...} | None, "resources": {"cpu": ...} | None}, ...],
     "podRequests":     {...} | None,        # pod.Spec.Resources.Requests
     "overhead":        {...} | None,        # pod.Spec.Overhead
     "resizeInfeasible": False}
"""
from __future__ import annotations

import ctypes as C

from .engine import load_library, YkError

POD_EXPORTS = ["yk_quantity_parse", "yk_podspec_new", "yk_podspec_free", "yk_podspec_clear", "yk_podspec_container",
               "yk_podspec_status", "yk_podspec_quantity", "yk_podspec_amount", "yk_podspec_resize_infeasible",
               "yk_podspec_compute", "yk_podspec_result", "yk_podspec_vector", "yk_node_quantity"]

CONTAINER, INIT_CONTAINER, INIT_SIDECAR = 0, 1, 2
LIST_REQUESTS, LIST_ALLOCATED, LIST_STATUS_REQUESTS, LIST_POD_REQUESTS, LIST_OVERHEAD = 0, 1, 2, 3, 4


def _lib():
    lib = load_library()
    lib.yk_podspec_new.restype = C.c_void_p
    lib.yk_podspec_free.argtypes = [C.c_void_p]
    lib.yk_podspec_free.restype = None
    lib.yk_podspec_clear.argtypes = [C.c_void_p]
    lib.yk_podspec_clear.restype = None
    lib.yk_strerror.restype = C.c_char_p
    return lib


def _ck(lib, rc):
    if rc < 0:
        raise YkError(rc, lib.yk_strerror(rc).decode())
    return rc


def parse_quantity(text: str):
    """-> (Value(), MilliValue()) of a k8s resource.Quantity in text form"""
    lib = _lib()
    v, m = C.c_int64(0), C.c_int64(0)
    _ck(lib, lib.yk_quantity_parse(text.encode(), C.byref(v), C.byref(m)))
    return v.value, m.value


def pod_resource(pod: dict) -> dict:
    """GetPodResource: -> {"vcore": milli, "memory": bytes, "pods": 1, ...}"""
    lib = _lib()
    h = C.c_void_p(lib.yk_podspec_new())
    if not h:
        raise MemoryError("yk_podspec_new")
    try:
        index = {}

        def add_list(which, ci, lst):
            for res, q in (lst or {}).items():
                _ck(lib, lib.yk_podspec_quantity(h, C.c_uint32(which), C.c_int32(ci), res.encode(), str(q).encode()))

        for c in pod.get("containers", []):
            ci = _ck(lib, lib.yk_podspec_container(h, c["name"].encode(), C.c_uint32(CONTAINER)))
            index[c["name"]] = ci
            add_list(LIST_REQUESTS, ci, c.get("requests"))
        for c in pod.get("initContainers", []):
            kind = INIT_SIDECAR if c.get("restartPolicy") == "Always" else INIT_CONTAINER
            ci = _ck(lib, lib.yk_podspec_container(h, c["name"].encode(), C.c_uint32(kind)))
            index[c["name"]] = ci
            add_list(LIST_REQUESTS, ci, c.get("requests"))
        for st in pod.get("statuses") or []:
            _ck(lib, lib.yk_podspec_status(h, st["name"].encode(), C.c_int32(1 if st.get("resources") is not None else 0)))
            ci = index.get(st["name"])
            if ci is None:
                continue                      # a status for a container the spec does not have: never looked up
            add_list(LIST_ALLOCATED, ci, st.get("allocated"))
            add_list(LIST_STATUS_REQUESTS, ci, st.get("resources"))
        if pod.get("podRequests") is not None:
            add_list(LIST_POD_REQUESTS, -1, pod["podRequests"])
        if pod.get("overhead") is not None:
            add_list(LIST_OVERHEAD, -1, pod["overhead"])
        _ck(lib, lib.yk_podspec_resize_infeasible(h, C.c_int32(1 if pod.get("resizeInfeasible") else 0)))
        n = _ck(lib, lib.yk_podspec_compute(h))
        out = {}
        for i in range(n):
            name, val = C.c_char_p(), C.c_int64(0)
            _ck(lib, lib.yk_podspec_result(h, C.c_uint32(i), C.byref(name), C.byref(val)))
            out[name.value.decode()] = val.value
        return out
    finally:
        lib.yk_podspec_free(h)


def pod_vector(pod: dict, dims):
    """-> (request vector in the order of `dims`, number of requested resources no dimension names)"""
    res = pod_resource(pod)
    return [res.get(d, 0) for d in dims], sum(1 for k in res if k not in dims)


def node_resource(allocatable: dict) -> dict:
    """GetNodeResource(node.Status.Allocatable)"""
    lib = _lib()
    out = {}
    for res, q in allocatable.items():
        name, val = C.c_char_p(), C.c_int64(0)
        _ck(lib, lib.yk_node_quantity(res.encode(), str(q).encode(), C.byref(name), C.byref(val)))
        out[name.value.decode()] = val.value
    return out
