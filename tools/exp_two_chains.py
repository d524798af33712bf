"""Experiment: how much throughput is left in overlapping two independent half-batch chains on one GPU?
Two models / two PretrainStep graphs (batch B/2 each) replayed concurrently on two streams vs one graph at batch B."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gptst_b200.GPTST import GPTST_Model
from gptst_b200.train import PretrainStep

N, D, B = bench.WORKLOADS["pems08"]
def mk(b):
    m = GPTST_Model(bench.make_cfg(N, D, "cuda")).cuda()
    bench.run_init(m, 0)
    st = PretrainStep(m)
    x = torch.randn(b, 12, N, 3, device="cuda")
    for _ in range(6):
        st(x, 200)
    torch.cuda.synchronize()
    return st, x

def timeit(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

full, xf = mk(B)
t_full = timeit(lambda: full(xf, 200))
print(f"one graph, batch {B}: {t_full:.4f} ms/step -> {B / t_full * 1e3:.0f} samples/s")
for nb in (2, 4):
    hs = [mk(B // nb) for _ in range(nb)]
    t1 = timeit(lambda: hs[0][0](hs[0][1], 200))
    print(f"one graph, batch {B // nb}: {t1:.4f} ms/step -> {B // nb / t1 * 1e3:.0f} samples/s")
    streams = [torch.cuda.Stream(priority=-1) for _ in range(nb)]
    def both():
        cur = torch.cuda.current_stream()
        for s, (st, x) in zip(streams, hs):
            s.wait_stream(cur)
            with torch.cuda.stream(s):
                st(x, 200)
        for s in streams:
            cur.wait_stream(s)
    t2 = timeit(both)
    print(f"{nb} graphs x batch {B // nb} concurrently: {t2:.4f} ms per pair -> {B / t2 * 1e3:.0f} samples/s  ({t_full / t2:.3f}x of the single graph)")
    del hs
