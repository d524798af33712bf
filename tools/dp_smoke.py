"""Staged multi-GPU smoke test of the data-parallel graph step (prints progress; dumps stacks if a stage stalls).
   torchrun --nproc-per-node 2 tools/dp_smoke.py [graph|eager]"""
import faulthandler, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(50, exit=True)
import torch
import torch.distributed as dist
import bench
from gptst_b200 import dp
from gptst_b200.GPTST import GPTST_Model
from gptst_b200.train import PretrainStep

mode = sys.argv[1] if len(sys.argv) > 1 else "graph"
t0 = time.time()
def log(msg):
    print(f"[rank {os.environ.get('RANK')}] +{time.time()-t0:5.1f}s {msg}", flush=True)
rank, local, world = dp.init_from_env()
log(f"init done world={world}")
dev = torch.device("cuda", local)
model = GPTST_Model(bench.make_cfg(170, 64, "cuda")).to(dev)
bench.run_init(model, 0)
dp.broadcast_parameters(model)
torch.cuda.synchronize(); log("broadcast done")
red = dp.FlatGradAllReduce(model.parameters())
step = PretrainStep(model, use_graph=(mode == "graph"), reducer=red)
x = torch.randn(8, 12, 170, 3, device=dev)
for i in range(6):
    loss = step(x, 200)
    torch.cuda.synchronize()
    log(f"step {i} loss {float(loss):.4f} replays {step.replays}")
dist.barrier(); torch.cuda.synchronize(); log("barrier done")
# gradient all-reduce really averaged: parameters must stay identical across ranks
p = next(model.parameters()).detach().clone()
q = p.clone(); dist.broadcast(q, src=0)
log(f"param sync max diff {float((p-q).abs().max()):.3e}")
log("done")
sys.stdout.flush()
os._exit(0)
