"""The graph-dependency tool (tools/graph_deps.py) on the committed dump of the captured step: node / edge parsing and the
longest-path computation that found the false dependency of round 2."""
import gzip
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_graph_dump_parses_and_has_one_optimiser_tail(tmp_path):
    import graph_deps as G
    dot = tmp_path / "step.dot"
    dot.write_bytes(gzip.open(os.path.join(ROOT, "profiles", "step_graph_r02_c.phase2.dot.gz")).read())
    nodes, pred, succ, d_of = G.load(str(dot), os.path.join(ROOT, "profiles", "step_timeline_r02_c.csv"))
    kinds = {}
    for n in nodes.values():
        kinds[n["kind"]] = kinds.get(n["kind"], 0) + 1
    assert kinds["KERNEL"] > 300 and kinds.get("MEMCPY", 0) == 2          # the optimiser's two pointer-table copies
    names = [n["short"] for n in nodes.values()]
    assert sum("htem_fwd_kernel" in s for s in names) == 8 and sum("cap_route2_fwd_kernel" in s for s in names) == 4
    # the graph is a DAG with a single sink: the Adam kernel, which (transitively) depends on every node
    sinks = [i for i in nodes if not succ[i]]
    assert len(sinks) == 1 and "opt_adam_kernel" in nodes[sinks[0]]["short"]
    seen, stack = set(), [sinks[0]]
    while stack:
        i = stack.pop()
        if i in seen:
            continue
        seen.add(i)
        stack.extend(pred[i])
    assert len(seen) == len(nodes)
    # main-chain kernels carry the capture stream's high priority, the table kernels the default one
    prios = {}
    for n in nodes.values():
        if n["kind"] == "KERNEL":
            prios.setdefault(n["short"], set()).add(n["prio"])
    assert prios["htf::htem_fwd_kernel"] == {-1} and 0 in prios["sm::table_dpool_kernel"]
    assert d_of(nodes[sinks[0]]) > 1.0
