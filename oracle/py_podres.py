"""Pure-Python restatement of the shim's pod / node resource arithmetic -- TEST INFRASTRUCTURE ONLY.

Follows /root/reference/pkg/common/resource.go (GetPodResource :56-109, computeContainerResource :111-127,
isResizeInfeasible :132-142, updateMax :145-162, checkInitContainerRequest :164-182, GetNodeResource :188-195,
getResource :273-285, getPodLevelResource :287-301, Add :331-351) with plain dicts, and the k8s quantity text form
[EXT k8s.io/apimachinery pkg/api/resource] with exact rationals (fractions.Fraction), so that it shares no code and no
number representation with csrc/yk_podres.cpp.  Pinned by tests/golden/pod_resources.json (the reference's own tests).
Imported only by tests/.
"""
from __future__ import annotations

import math
import re
from fractions import Fraction

I64_MAX = (1 << 63) - 1
_BIN = {"Ki": 10, "Mi": 20, "Gi": 30, "Ti": 40, "Pi": 50, "Ei": 60}
_DEC = {"n": -9, "u": -6, "m": -3, "": 0, "k": 3, "M": 6, "G": 9, "T": 12, "P": 15, "E": 18}
_NUM = re.compile(r"^([+-]?)(\d*)(?:\.(\d*))?(.*)$", re.S)


def quantity(text: str) -> Fraction:
    m = _NUM.match(text)
    if not m:
        raise ValueError(text)
    sign, whole, frac, suffix = m.group(1), m.group(2), m.group(3) or "", m.group(4)
    if whole == "" and frac == "":
        raise ValueError(text)
    if "." in text and m.group(3) is None:
        raise ValueError(text)
    x = Fraction(int((whole or "0") + frac), 10 ** len(frac))
    if suffix in _BIN:
        x *= 2 ** _BIN[suffix]
    elif suffix in _DEC:
        x *= Fraction(10) ** _DEC[suffix]
    elif len(suffix) > 1 and suffix[0] in "eE" and re.fullmatch(r"[+-]?\d+", suffix[1:]):
        e = int(suffix[1:])
        if abs(e) > 100000:
            raise OverflowError(text)
        x *= Fraction(10) ** e if abs(e) < 400 else (Fraction(10) ** 400 if e > 0 else Fraction(0) if x == 0 else Fraction(1, 10 ** 400))
    else:
        raise ValueError(text)
    return -x if sign == "-" else x


def _away(x: Fraction) -> int:
    v = math.ceil(abs(x))
    v = min(v, I64_MAX)
    return -v if x < 0 else v


def value(text: str) -> int:          # Quantity.Value()
    return _away(quantity(text))


def milli_value(text: str) -> int:    # Quantity.MilliValue()
    return _away(quantity(text) * 1000)


def get_resource(lst):                # getResource: cpu -> vcore in milli, the rest by name in units
    out = {}
    for name, q in (lst or {}).items():
        if name == "cpu":
            out["vcore"] = milli_value(str(q))
        else:
            out[name] = value(str(q))
    return out


def add(left, right):
    out = dict(right or {})
    for k, v in (left or {}).items():
        out[k] = _wrap(out[k] + v) if k in out else v
    return out


def _wrap(x: int) -> int:             # Go's int64 addition wraps around
    x &= (1 << 64) - 1
    return x - (1 << 64) if x >= (1 << 63) else x


def update_max(left, right):
    for k, v in (right or {}).items():
        if k not in left or v > left[k]:
            left[k] = v


def _container(pod, c, statuses):
    combined = {}
    update_max(combined, get_resource(c.get("requests")))
    st = statuses.get(c["name"])
    if st is not None:
        if pod.get("resizeInfeasible") and st.get("resources") is not None:
            return get_resource(st["resources"])
        update_max(combined, get_resource(st.get("allocated")))
        if st.get("resources") is not None:
            update_max(combined, get_resource(st["resources"]))
    return combined


def pod_resource(pod: dict) -> dict:
    res = {"pods": 1}
    statuses = {s["name"]: s for s in (pod.get("statuses") or [])}
    for c in pod.get("containers", []):
        res = add(res, _container(pod, c, statuses))
    inits = pod.get("initContainers") or []
    if inits:
        init_max, sidecars = {}, None
        for c in inits:
            own = _container(pod, c, statuses)
            current = add(own, sidecars)
            if c.get("restartPolicy") == "Always":
                sidecars = add(sidecars, own)
            update_max(init_max, current)
        res = add(res, sidecars)
        update_max(res, init_max)
    pr = pod.get("podRequests")
    if pr:
        for name, v in get_resource(pr).items():
            if name in ("vcore", "memory") or name.startswith("hugepages-"):
                res[name] = v
    if pod.get("overhead") is not None:
        res = add(res, get_resource(pod["overhead"]))
    return res


def node_resource(allocatable: dict) -> dict:
    return get_resource(allocatable)
