"""String-level restatement of the Kubernetes filter plugins the fast path covers -- TEST INFRASTRUCTURE ONLY.

What `predicateManagerImpl.Predicates` (/root/reference/pkg/plugin/predicates/predicate_manager.go:130-283, plugin list
:339-351) asks of NodeAffinity, TaintToleration, NodeUnschedulable and NodeName, evaluated directly on label / taint /
toleration strings [EXT k8s.io/kubernetes v1.34.6 plugins nodeaffinity, tainttoleration, nodeunschedulable, nodename;
k8s.io/component-helpers/scheduling/corev1/nodeaffinity; k8s.io/api core/v1 toleration.go] as summarised in SURVEY.md
Appendix A.4.  It knows nothing about bit sets: tests/test_dictionary_random.py uses it to check that the masks produced
by csrc/yk_dict.cpp give the same verdict for every (pod, node) pair.  Pinned by the same reference tables as the encoder
(tests/golden/pod_fits_selector.json).  Imported only by tests/.
"""
from __future__ import annotations

import re

UNSCHEDULABLE_KEY = "node.kubernetes.io/unschedulable"


def _int(s):
    try:
        if s.strip() != s or s == "" or s[0] == "+" and len(s) == 1:
            return None
        return int(s, 10)
    except ValueError:
        return None


_VALUE = re.compile(r"^([A-Za-z0-9]([-A-Za-z0-9_.]*[A-Za-z0-9])?)?$")


def _valid_value(v):                   # validation.IsValidLabelValue [EXT k8s.io/apimachinery]
    return len(v) <= 63 and _VALUE.match(v) is not None


def requirement_matches(req, labels):
    """one matchExpressions entry against the node labels; None = the requirement is malformed"""
    key, op, values = req["key"], req["op"], req.get("values") or []
    if any(not _valid_value(v) for v in values):
        return None
    if op == "In":
        return None if not values else (key in labels and labels[key] in values)
    if op == "NotIn":
        return None if not values else (key not in labels or labels[key] not in values)
    if op == "Exists":
        return None if values else key in labels
    if op == "DoesNotExist":
        return None if values else key not in labels
    if op in ("Gt", "Lt"):
        if len(values) != 1 or _int(values[0]) is None:
            return None
        if key not in labels or _int(labels[key]) is None:
            return False
        return _int(labels[key]) > _int(values[0]) if op == "Gt" else _int(labels[key]) < _int(values[0])
    return None


def field_matches(req, node_name):
    key, op, values = req["key"], req["op"], req.get("values") or []
    if key != "metadata.name" or len(values) != 1 or op not in ("In", "NotIn"):
        return None
    return (node_name == values[0]) if op == "In" else (node_name != values[0])


def term_matches(term, labels, node_name):
    ex, fl = term.get("expr") or [], term.get("fields") or []
    if not ex and not fl:
        return False                       # an empty term matches no objects
    results = [requirement_matches(r, labels) for r in ex] + [field_matches(r, node_name) for r in fl]
    if any(r is None for r in results):
        return False                       # a term that cannot be parsed matches nothing
    return all(results)


def node_affinity(pod, labels, node_name):
    for k, v in (pod.get("node_selector") or {}).items():
        if labels.get(k) != v:
            return False
    if pod.get("has_affinity"):
        terms = pod.get("affinity_terms") or []
        if not any(term_matches(t, labels, node_name) for t in terms):
            return False
    return True


def tolerates(tol, taint):
    if tol.get("effect", "") not in ("", taint.get("effect", "NoSchedule")):
        return False
    if tol.get("key", "") != "" and tol["key"] != taint["key"]:
        return False
    op = tol.get("op", "") or "Equal"
    if tol.get("key", "") == "" and op != "Exists":
        return False                       # an empty key is only legal (and only matches everything) with Exists
    if op == "Exists":
        return True
    return tol.get("value", "") == taint.get("value", "")


def taint_toleration(pod, taints, unschedulable):
    todo = [t for t in taints if t.get("effect", "NoSchedule") in ("NoSchedule", "NoExecute")]
    if unschedulable:
        todo.append({"key": UNSCHEDULABLE_KEY, "value": "", "effect": "NoSchedule"})
    tols = pod.get("tolerations") or []
    return all(any(tolerates(tl, t) for tl in tols) for t in todo)


def fits(pod, node):
    if pod.get("node_name") and pod["node_name"] != node["name"]:
        return False
    return node_affinity(pod, node.get("labels") or {}, node["name"]) and \
        taint_toleration(pod, node.get("taints") or [], node.get("unschedulable", False))
