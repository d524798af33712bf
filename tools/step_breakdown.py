"""Where does the graph-captured step go?  Times (CUDA graph replays) forward only / forward+backward / full step."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gptst_b200 import ops
from gptst_b200.GPTST import GPTST_Model
from gptst_b200.train import PretrainStep

epoch = int(sys.argv[1]) if len(sys.argv) > 1 else 200
N, D, B = bench.WORKLOADS["pems08"]
model = GPTST_Model(bench.make_cfg(N, D, "cuda")).cuda()
bench.run_init(model, 0)
x = torch.randn(B, 12, N, 3, device="cuda")

def timeit(fn, n=30):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n

def capture(body):
    s = torch.cuda.Stream(); s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3): body()
    torch.cuda.current_stream().wait_stream(s); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): body()
    return g

import random
n = B * 12 * N
plan = model.encoder.mask_plan(n, epoch).cuda()
model.encoder.plan_override = plan
params = [p for p in model.parameters()]
def fwd():
    with torch.no_grad():
        return model(x, x, None, epoch)
def fwd_loss():
    with torch.no_grad():
        outs = model(x, x, None, epoch); return ops.fused_probe_loss(outs, x, epoch > 10)
def fwd_bwd():
    for p in params: p.grad = None
    outs = model(x, x, None, epoch); l = ops.fused_probe_loss(outs, x, epoch > 10); l.backward(); return l
g1 = capture(fwd); t1 = timeit(g1.replay)
g2 = capture(fwd_bwd); t2 = timeit(g2.replay)
model.encoder.plan_override = None
step = PretrainStep(model)
for _ in range(6): step(x, epoch)
t3 = timeit(lambda: step(x, epoch))
os.environ["GPTST_B200_SIDE_STREAMS"] = "1"
print(f"epoch arg {epoch}: forward {t1:.3f} ms | forward+loss+backward {t2:.3f} ms | full step (+clip+Adam) {t3:.3f} ms")
