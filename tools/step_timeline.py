"""Timeline of ONE graph-replayed pre-training step (torch.profiler / CUPTI): start offset, duration and stream of
every kernel, written as CSV so the critical path can be read offline.
usage: python tools/step_timeline.py [out.csv] [epoch=200] [workload=pems08] [batch]"""
import os, sys, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
import bench
from gptst_b200.GPTST import GPTST_Model
from gptst_b200.train import PretrainStep

out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_timeline.csv"
epoch = int(sys.argv[2]) if len(sys.argv) > 2 else 200
N, D, B = bench.WORKLOADS[sys.argv[3] if len(sys.argv) > 3 else "pems08"]
if len(sys.argv) > 4:
    B = int(sys.argv[4])
model = GPTST_Model(bench.make_cfg(N, D, "cuda")).cuda()
bench.run_init(model, 0)
step = PretrainStep(model)
x = torch.randn(B, 12, N, 3, device="cuda")
for _ in range(8):
    step(x, epoch)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    step(x, epoch)
    torch.cuda.synchronize()
rows = []
try:   # kineto events carry the stream id
    for k in prof.profiler.kineto_results.events():
        if "CUDA" in str(k.device_type()) and k.duration_ns() > 0:
            rows.append((k.start_ns() / 1e3, k.duration_ns() / 1e3, k.device_resource_id(), k.name()))
except Exception as exc:  # older/newer torch: fall back to FunctionEvents (no stream id)
    print("kineto_results unavailable:", exc)
    rows = [(e.time_range.start, e.time_range.end - e.time_range.start, -1, e.name) for e in prof.events()
            if e.device_type.name == "CUDA"]
rows.sort()
t0 = rows[0][0]
os.makedirs(os.path.dirname(out) or ".", exist_ok=True)
with open(out, "w") as f:
    f.write("start_us,dur_us,stream,name\n")
    for st, du, sid, name in rows:
        n = re.sub(r"\(.*", "", re.sub(r"<.*", "", name))[:80].replace(",", ";")
        f.write(f"{st - t0:.1f},{du:.1f},{sid},{n}\n")
end = max(st + du for st, du, _, _ in rows) - t0
busy = sum(du for _, du, _, _ in rows)
print(f"{len(rows)} kernels, span {end/1e3:.3f} ms, summed kernel time {busy/1e3:.3f} ms -> {out}")
