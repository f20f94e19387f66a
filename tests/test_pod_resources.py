"""Request vectors in: the pod / node resource arithmetic of the shim
(/root/reference/pkg/common/resource.go) as csrc/yk_podres.cpp computes it behind include/ykgpu_pod.h, against
 (1) the reference's own known answers, transcribed into tests/golden/pod_resources.json (pinned parity), and
 (2) an independent exact-rational Python restatement (oracle/py_podres.py) on randomized pods and quantities.
Host code only: runs without a GPU."""
import json
import os
import random

import pytest

from oracle import py_podres
from yunikorn_k8shim_b200 import podres

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "pod_resources.json")))


def _check(case, got):
    for k, v in case["expect"].items():
        assert got.get(k) == v, (case["name"], case["source"], k, got)
    if "len" in case:
        assert len(got) == case["len"], (case["name"], got)
    for k in case.get("absent", []):
        assert k not in got, (case["name"], k)


@pytest.mark.parametrize("case", GOLD["pods"], ids=[c["name"] for c in GOLD["pods"]])
def test_oracle_matches_reference_known_answers(case):
    _check(case, py_podres.pod_resource(case["pod"]))


@pytest.mark.parametrize("case", GOLD["pods"], ids=[c["name"] for c in GOLD["pods"]])
def test_library_matches_reference_known_answers(case):
    got = podres.pod_resource(case["pod"])
    _check(case, got)
    assert got == py_podres.pod_resource(case["pod"])        # and agrees with the oracle on the keys the test leaves open


def test_node_resource_known_answer():
    for case in GOLD["nodes"]:
        assert py_podres.node_resource(case["allocatable"]) == case["expect"]
        assert podres.node_resource(case["allocatable"]) == case["expect"]


def test_quantities_used_by_the_reference_tests():
    for text, value, milli in GOLD["quantities"]["cases"]:
        v, m = podres.parse_quantity(text)
        if value is not None:
            assert v == value == py_podres.value(text), text
        if milli is not None:
            assert m == milli == py_podres.milli_value(text), text


SUFFIXES = ["", "", "", "m", "k", "M", "G", "T", "Ki", "Mi", "Gi", "Ti", "n", "u", "P", "E", "Pi", "Ei", "e3", "E2", "e-2", "e+1", "e0"]


def _rand_quantity(rng):
    whole = str(rng.randrange(0, 10 ** rng.randrange(1, 8))) if rng.random() < 0.9 else ""
    frac = ""
    if rng.random() < 0.5 or whole == "":
        frac = "." + "".join(rng.choice("0123456789") for _ in range(rng.randrange(0 if whole else 1, 7)))
    sign = rng.choice(["", "", "", "+"])
    return sign + whole + frac + rng.choice(SUFFIXES)


def test_quantity_parser_against_exact_rationals():
    rng = random.Random(11)
    n = 0
    for _ in range(20000):
        text = _rand_quantity(rng)
        try:
            want = (py_podres.value(text), py_podres.milli_value(text))
        except ValueError:
            with pytest.raises(Exception):
                podres.parse_quantity(text)
            continue
        assert podres.parse_quantity(text) == want, text
        n += 1
    assert n > 15000
    for bad in ["", ".", "abc", "1.2.3", "1e", "1Kii", "--1", "1 ", " 1", "1e1.5", "1ki", "1mi", "1K", "e3", "1ee3", "1E+"]:
        with pytest.raises(Exception):
            podres.parse_quantity(bad)
        with pytest.raises(Exception):
            py_podres.value(bad)
    # saturation and sign
    assert podres.parse_quantity("1E") == (10 ** 18, (1 << 63) - 1) == (py_podres.value("1E"), py_podres.milli_value("1E"))
    assert podres.parse_quantity("-1.5") == (-2, -1500) == (py_podres.value("-1.5"), py_podres.milli_value("-1.5"))
    assert podres.parse_quantity("1e30")[0] == (1 << 63) - 1


def _rand_list(rng, names):
    return {n: _rand_quantity(rng).lstrip("+") for n in rng.sample(names, rng.randrange(0, len(names) + 1))}


def test_random_pods_against_the_oracle():
    rng = random.Random(5)
    names = ["cpu", "memory", "nvidia.com/gpu", "ephemeral-storage", "hugepages-2Mi", "example.com/foo"]
    for i in range(3000):
        pod = {"containers": [{"name": f"c{k}", "requests": _rand_list(rng, names)} for k in range(rng.randrange(1, 4))]}
        if rng.random() < 0.6:
            pod["initContainers"] = []
            for k in range(rng.randrange(1, 5)):
                c = {"name": f"i{k}", "requests": _rand_list(rng, names)}
                if rng.random() < 0.5:
                    c["restartPolicy"] = "Always"
                pod["initContainers"].append(c)
        if rng.random() < 0.5:
            pod["statuses"] = []
            for c in pod["containers"] + pod.get("initContainers", []):
                if rng.random() < 0.7:
                    pod["statuses"].append({"name": c["name"],
                                            "allocated": _rand_list(rng, names) if rng.random() < 0.7 else None,
                                            "resources": _rand_list(rng, names) if rng.random() < 0.7 else None})
            if rng.random() < 0.2:
                pod["statuses"].append({"name": "ghost", "allocated": {"cpu": "64"}, "resources": {"cpu": "64"}})
        if rng.random() < 0.3:
            pod["podRequests"] = _rand_list(rng, names)
        if rng.random() < 0.3:
            pod["overhead"] = _rand_list(rng, names)
        pod["resizeInfeasible"] = rng.random() < 0.3
        assert podres.pod_resource(pod) == py_podres.pod_resource(pod), (i, pod)


def test_vector_for_the_engine():
    pod = GOLD["pods"][0]["pod"]
    vec, unmapped = podres.pod_vector(pod, ["vcore", "memory", "pods", "ephemeral-storage"])
    assert vec == [3000, 1524000000, 1, 0]
    assert unmapped == 1          # nvidia.com/gpu has no dimension: such a pod must take the slow path
    vec, unmapped = podres.pod_vector(pod, ["vcore", "memory", "pods", "nvidia.com/gpu"])
    assert vec == [3000, 1524000000, 1, 5] and unmapped == 0
