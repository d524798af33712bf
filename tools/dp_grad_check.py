"""1-vs-N-GPU gradient equality on the REAL model over NCCL (SURVEY.md section 4 "distributed").

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tools/dp_grad_check.py [epoch]

Every rank builds the same model, takes rows [rank::world] of one global batch, runs forward + probe loss + backward with the
data-parallel reducer (one flat NCCL all-reduce at the end of backward, gptst_b200.dp) and keeps the reduced gradients.  Rank 0
then recomputes, ALONE, what data parallelism means for this model -- each shard with its own exact-count mask built from the same
injected draws, per-shard loss, mean of the shard gradients -- and compares: the NCCL result must equal the single-GPU result
to fp32 summation order (sum of `world` terms: 2e-6 x abs-max), parameters without a gradient must stay None on every rank, and
a second pass through the CUDA-graph PretrainStep with the reducer inside the graph must leave all ranks with identical
parameters.  Prints one JSON line on rank 0; exit code 1 on mismatch."""
import json, os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
import bench
from gptst_b200 import dp, ops
from gptst_b200.GPTST import GPTST_Model
from gptst_b200.train import PretrainStep

epoch = int(sys.argv[1]) if len(sys.argv) > 1 else 200
rank, local, world = dp.init_from_env()
dev = torch.device("cuda", local)
N, D = 170, 64
Bg = 4 * world
cfg = bench.make_cfg(N, D, "cuda")
model = GPTST_Model(cfg).to(dev)
bench.run_init(model, 0)
dp.broadcast_parameters(model)
gen = torch.Generator().manual_seed(7)
xg = torch.randn(Bg, 12, N, 3, generator=gen)
n_shard = (Bg // world) * 12 * N


def shard_run(m, r, reducer=None):
    """forward + loss + backward of shard r with that shard's injected draws; returns the gradients (None kept)."""
    g = torch.Generator().manual_seed(1000 + r)
    m.encoder.draws_override = {"u1": torch.rand(n_shard, generator=g).to(dev), "u2": torch.rand(n_shard, generator=g).to(dev)}
    random.seed(50 + r)
    for p in m.parameters():
        p.grad = None
    x = xg[r::world].to(dev)
    outs = m(x, x, None, epoch)
    ops.fused_probe_loss(outs, x, epoch > 10).backward()
    if reducer is not None:
        reducer.reduce()
    torch.cuda.synchronize()
    m.encoder.draws_override = None
    return [None if p.grad is None else p.grad.detach().clone() for p in m.parameters()]


red = dp.FlatGradAllReduce(model.parameters())
got = shard_run(model, rank, red)
ok, worst, n_none = True, 0.0, sum(g is None for g in got)
none_counts = [torch.zeros(1, device=dev) for _ in range(world)]
dist.all_gather(none_counts, torch.tensor([float(n_none)], device=dev))
if rank == 0:
    acc = None
    for r in range(world):
        gs = shard_run(model, r)
        acc = gs if acc is None else [None if a is None else a + b for a, b in zip(acc, gs)]
    for (k, _), a, b in zip(model.named_parameters(), acc, got):
        if (a is None) != (b is None):
            ok = False
            print("None mismatch", k)
            continue
        if a is None:
            continue
        a = a / world
        err = (a - b).abs().max().item() / max(1e-30, a.abs().max().item())
        worst = max(worst, err)
        if err > 2e-6:
            ok = False
            print(f"gradient mismatch {k}: {err:.3e}")
    ok = ok and len({int(c.item()) for c in none_counts}) == 1

# the reducers inside the captured graph: ranks must stay in lock step, and the bucketed exchange (rank-summed gradients left in
# the flat buffers, 1/world folded into the clip coefficient, decoder bucket overlapped with the encoder backward) must land on
# the same parameters as the flat one (same draws: seeded torch / python RNG before each run)
import copy
xs = xg[rank::world].to(dev)


def graph_run(m, reducer, steps=7):
    torch.manual_seed(123 + rank)
    random.seed(77)
    st = PretrainStep(m, lr=3e-3, max_grad_norm=5.0, loss="probe", use_graph=True, reducer=reducer)
    for _ in range(steps):
        st(xs, epoch)
    torch.cuda.synchronize()
    return st, torch.cat([p.detach().reshape(-1) for p in m.parameters()])


model_b = copy.deepcopy(model)
stepper, flat = graph_run(model, red)
red_b = dp.BucketedGradAllReduce(dp.pretrain_buckets(model_b))
stepper_b, flat_b = graph_run(model_b, red_b)
ref = flat.clone()
dist.broadcast(ref, src=0)
sync_err = (flat - ref).abs().max()
ref_b = flat_b.clone()
dist.broadcast(ref_b, src=0)
sync_err = torch.maximum(sync_err, (flat_b - ref_b).abs().max())
dist.all_reduce(sync_err, op=dist.ReduceOp.MAX)
bucket_vs_flat = ((flat_b - flat).abs().max() / flat.abs().max()).item()
if rank == 0:
    ok = ok and sync_err.item() == 0.0 and bucket_vs_flat <= 1e-4
    print(json.dumps({"world": world, "epoch": epoch, "ok": bool(ok), "worst_rel_grad_err_vs_single_gpu": worst,
                      "params_without_grad": n_none, "all_reduce_numel": red.last_numel, "graph_replays": stepper.replays,
                      "param_max_diff_across_ranks_after_7_graph_steps": sync_err.item(),
                      "bucketed_vs_flat_param_rel_diff_after_7_graph_steps": bucket_vs_flat,
                      "bucketed_all_reduce_numel": red_b.last_numel}), flush=True)
stepper_b._graphs.clear()
stepper._graphs.clear()
import gc
gc.collect()
torch.cuda.synchronize()
dist.barrier()
torch.cuda.synchronize()
sys.stdout.flush()
import time
time.sleep(3.0)          # see bench.py: no rank tears its context down while a peer may still be inside the barrier
os._exit(0 if (rank != 0 or ok) else 1)
