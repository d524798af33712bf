"""Per-kernel device time of ONE eager pre-training step (torch.profiler, warm caches) -- cheap alternative to an ncu
launch list for finding where the step time goes.  usage: python tools/prof_step.py [epoch=200] [B=64]"""
import os, sys, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from gptst_b200.GPTST import GPTST_Model
from gptst_b200.train import PretrainStep

epoch = int(sys.argv[1]) if len(sys.argv) > 1 else 200
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
N, D, _ = bench.WORKLOADS["pems08"]
model = GPTST_Model(bench.make_cfg(N, D, "cuda")).cuda()
bench.run_init(model, 0)
step = PretrainStep(model, use_graph=False)
x = torch.randn(B, 12, N, 3, device="cuda")
for _ in range(4):
    step(x, epoch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(x, epoch)
    torch.cuda.synchronize()
agg = collections.defaultdict(lambda: [0, 0.0])
for e in prof.events():
    if e.device_type.name == "CUDA":
        n = re.sub(r"<.*", "", e.name)
        n = re.sub(r"\(.*", "", n)
        agg[n][0] += 1
        agg[n][1] += e.device_time
tot = sum(v[1] for v in agg.values())
cnt = sum(v[0] for v in agg.values())
mine = sum(v[1] for k, v in agg.items() if "gptst::" in k)
print(f"one step: {cnt} kernels, {tot/1e3:.2f} ms device time; gptst kernels {mine/1e3:.2f} ms, torch/library {(tot-mine)/1e3:.2f} ms")
for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:45]:
    print(f"{t:9.1f} us {c:5d}  {n[:110]}")
print("---- by operator (self device time)")
ka = prof.key_averages()
rows = sorted(ka, key=lambda e: -e.self_device_time_total)
for e in rows[:40]:
    if e.self_device_time_total > 0:
        print(f"{e.self_device_time_total:9.1f} us {e.count:5d}  {e.key[:90]}")
