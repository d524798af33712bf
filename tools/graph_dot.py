"""Dump the captured pre-training step as a DOT file (cudaGraphDebugDotPrint) for dependency analysis.
usage: python tools/graph_dot.py [out_prefix=gpurun_out/step_graph] [epoch=200]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
out = sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/step_graph"
os.environ["GPTST_B200_GRAPH_DOT"] = out
import torch
import bench
from gptst_b200.GPTST import GPTST_Model
from gptst_b200.train import PretrainStep
epoch = int(sys.argv[2]) if len(sys.argv) > 2 else 200
N, D, B = bench.WORKLOADS["pems08"]
model = GPTST_Model(bench.make_cfg(N, D, "cuda")).cuda()
bench.run_init(model, 0)
step = PretrainStep(model)
x = torch.randn(B, 12, N, 3, device="cuda")
for _ in range(6):
    step(x, epoch)
torch.cuda.synchronize()
print("dumped", [f for f in os.listdir(os.path.dirname(out) or ".") if f.endswith(".dot")])
