"""A/B of the scheduling knobs of the captured pre-training step in ONE process (PEMS08 geometry, batch 64): per variant a fresh
model + PretrainStep, 3 rounds of 30 graph-replayed steps with resident inputs and 30 through host buffers + float(loss);
prints one JSON line per variant (min / median ms per step).  Not a bench line -- bench.py is."""
import json, os, random, statistics, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from gptst_b200.GPTST import GPTST_Model
from gptst_b200.train import PretrainStep

# (library-side knobs such as GPTST_B200_PDL are read once per process by libgptst_b200.so: A/B those with separate bench.py runs)
KNOBS = ("GPTST_B200_OPT_PREFETCH", "GPTST_B200_SCORER_PRIO", "GPTST_B200_HTEM", "GPTST_B200_CAP", "GPTST_B200_CAP_Z")
VARIANTS = [{}, {"GPTST_B200_OPT_PREFETCH": "1"}, {"GPTST_B200_SCORER_PRIO": "low"}, {"GPTST_B200_CAP_Z": "recompute"},
            {"GPTST_B200_HTEM": "split"}, {"GPTST_B200_CAP": "split"}]
if len(sys.argv) > 1:          # python tools/ab_knobs.py '[{}, {"GPTST_B200_HTEM": "split"}]'
    VARIANTS = json.loads(sys.argv[1])
N, D, B = bench.WORKLOADS["pems08"]
g = torch.Generator().manual_seed(100)
host = [torch.randn(B, 12, N, 3, generator=g).pin_memory() for _ in range(4)]
res = [h.cuda() for h in host]


def timed(step, n, e2e):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for i in range(n):
        last = float(step(host[i % 4], 200)) if e2e else step(res[i % 4], 200)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, float(last)


for v in VARIANTS:
    for k in KNOBS:
        os.environ.pop(k, None)
    os.environ.update(v)
    model = GPTST_Model(bench.make_cfg(N, D, "cuda")).cuda()
    bench.run_init(model, 0)
    step = PretrainStep(model, lr=3e-3, max_grad_norm=5.0, loss="probe")
    random.seed(1234)
    torch.manual_seed(1234)
    first = [float(step(res[i % 4], 200)) for i in range(6)]
    r = [timed(step, 30, False) for _ in range(3)]
    e = [timed(step, 30, True) for _ in range(2)]
    print(json.dumps({"env": v, "ms_min": min(x[0] for x in r), "ms_med": statistics.median(x[0] for x in r),
                      "e2e_ms_min": min(x[0] for x in e), "first_losses": first, "loss_after_96": r[-1][1], "loss_after_156": e[-1][1]}), flush=True)
    del step, model
    torch.cuda.empty_cache()
