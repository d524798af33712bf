"""Parse a cudaGraphDebugDotPrint dump of the captured step (tools/graph_dot.py): nodes (capture order, kernel, priority) and
dependency edges; prints the longest dependency path with per-kernel durations taken from a step timeline.
usage: python tools/graph_deps.py step_graph.phase2.dot [timeline.csv]"""
import re, sys, collections, subprocess, csv

def parse(path):
    txt = open(path).read()
    nodes = {}
    for m in re.finditer(r'"graph_1_node_(\d+)"\[[^\]]*?label="\{(\w+)(.*?)\}"\];', txt, re.S):
        nid, kind, body = int(m.group(1)), m.group(2), m.group(3)
        name, grid, prio = kind, "", None
        if kind == "KERNEL":
            mm = re.search(r'\| (\S+?)\\<\\<\\<(.*?)\\>\\>\\>', body)
            if mm:
                name, grid = mm.group(1), mm.group(2).replace("\\", "")
            pm = re.search(r'\{priority \| (-?\d+)\}', body)
            prio = int(pm.group(1)) if pm else None
        nodes[nid] = dict(kind=kind, name=name, grid=grid, prio=prio)
    for m in re.finditer(r'"graph_1_node_(\d+)"\[[^\]]*?label="\{\s*(MEMCPY|MEMSET|EVENT\w*|EMPTY|HOST)', txt):
        nodes.setdefault(int(m.group(1)), dict(kind=m.group(2), name=m.group(2), grid="", prio=None))
    edges = [(int(a), int(b)) for a, b in re.findall(r'"graph_1_node_(\d+)" -> "graph_1_node_(\d+)"', txt)]
    return nodes, edges

def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.split("\n")
    return dict(zip(names, out))

def short(d):
    d = re.sub(r'\(.*', '', d)
    d = re.sub(r'<.*', '', d)
    return d.replace("void ", "").replace("gptst::", "")

def load(path, timeline=None):
    nodes, edges = parse(path)
    dm = demangle(sorted({n["name"] for n in nodes.values()}))
    for n in nodes.values():
        n["short"] = short(dm.get(n["name"], n["name"]))
    pred = collections.defaultdict(list); succ = collections.defaultdict(list)
    for a, b in edges:
        pred[b].append(a); succ[a].append(b)
    dur = {}
    if timeline:
        acc = collections.defaultdict(list)
        for r in csv.DictReader(open(timeline)):
            nm = r["name"].replace("void ", "").replace("gptst::", "")
            acc[nm].append(float(r["dur_us"]))
        for k, v in acc.items():       # 20th percentile: side-stream kernels are stretched when they share the GPU
            v.sort(); dur[k] = v[len(v) // 5]
    def d_of(n):
        s = n["short"]
        if s in dur: return dur[s]
        for k in dur:
            if k and (k.startswith(s) or s.startswith(k)): return dur[k]
        return 3.0 if n["kind"] == "KERNEL" else 2.0
    return nodes, pred, succ, d_of

if __name__ == "__main__":
    nodes, pred, succ, d_of = load(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
    print(len(nodes), "nodes; kinds", collections.Counter(n["kind"] for n in nodes.values()))
    indeg = {i: len(pred[i]) for i in nodes}
    q = [i for i in nodes if indeg[i] == 0]; topo = []
    while q:
        i = q.pop(); topo.append(i)
        for j in succ[i]:
            indeg[j] -= 1
            if indeg[j] == 0: q.append(j)
    fin = {}; best = {}
    for i in topo:
        st = max((fin[p] for p in pred[i]), default=0.0)
        best[i] = max(pred[i], key=lambda p: fin[p]) if pred[i] else None
        fin[i] = st + d_of(nodes[i])
    end = max(fin, key=fin.get)
    print(f"longest path (dependency-only, per-name durations): {fin[end]:.1f} us")
    chain = []
    i = end
    while i is not None:
        chain.append(i); i = best[i]
    for i in reversed(chain):
        n = nodes[i]
        print(f"  {fin[i]:8.1f}  +{d_of(n):6.1f}  node {i:4d} prio {n['prio']} {n['short'][:60]} <<<{n['grid']}>>> preds={pred[i]}")
