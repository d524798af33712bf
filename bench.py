#!/usr/bin/env python
"""bench.py -- pre-training samples/s of the GPT-ST hot path on B200 (BASELINE.json metric) + roofline of the
fused hypergraph / adaptive-GCN block.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One "step" = one full pre-training step on one synthetic batch: zero_grad, GPTST_Model forward (mask scoring,
masking, encoder + decoder STHCN), loss, backward, (N>1: one flat-gradient NCCL all-reduce),
clip_grad_norm_(5), Adam(lr 3e-3) -- the step of the reference trainer (BasicTrainer.py:72-103) on the synthetic
inputs of the driver's GPU probe.  Workload = BASELINE.json configs[1]: PEMS08 geometry,
batch 64 per GPU, N=170, T=12, D=64, epoch argument 200 (adaptive-mask + KL phase, 290 of the 300 epochs).
Weak scaling: every rank keeps batch 64; `value` = world * 64 * K / max-over-ranks device time.

--impl reference times the reference's own CPU PyTorch implementation of the same step on the host cores
(rank 0 only).  See DESIGN.md section "Measurement" for every field of the JSON line.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time
import types

REPO = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, REPO)

import torch  # noqa: E402

T_STEPS = 12
WORKLOADS = {
    # name: (N, D, per-GPU batch at one GPU)
    "pems08": (170, 64, 64),          # BASELINE.json configs[1]; weak scaling (64 per GPU)
    "metr_la": (207, 64, 64),         # configs[2]; weak scaling
    "synthetic2048": (2048, 128, 128),  # configs[3]; STRONG scaling: global batch 128 split over the ranks
}
STRONG = {"synthetic2048"}


def per_gpu_batch(name, world, override=0):
    B = WORKLOADS[name][2]
    if override:
        return override
    if name in STRONG:
        if B % world:
            raise SystemExit(f"{name}: global batch {B} is not divisible by {world} ranks")
        return B // world
    return B


def make_cfg(N, D, device):
    return types.SimpleNamespace(num_nodes=N, input_base_dim=1, input_extra_dim=2, hidden_dim=D, output_dim=1, horizon=12,
                                 lag=12, embed_dim=16, embed_dim_spa=4, HS=10, HT=16, HT_Tem=8, num_route=2, mode="pretrain",
                                 model="TGCN", device=device, scaler_zeros=-1.5767, interval=5, week_day=7, mask_ratio=0.25,
                                 ada_mask_ratio=0.5, ada_type="all", change_epoch=10, epochs=300)


def run_init(model, seed=0):
    torch.manual_seed(seed)
    for p in model.parameters():  # reference Run.py:79-85
        if p.dim() > 1:
            torch.nn.init.xavier_uniform_(p)
        else:
            torch.nn.init.uniform_(p)


# ------------------------------------------------------------------------------------------------------
# clocks (B200_PROFILING.md recipe)
# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx, self.proc, self.path = gpu_index, None, None

    def start(self):
        try:
            fd, self.path = tempfile.mkstemp(suffix=".csv")
            os.close(fd)
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.idx), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        try:
            for line in open(self.path):
                f = [c.strip() for c in line.split(",")]
                if len(f) < 9:
                    continue
                try:
                    sm.append(float(f[1]))
                    mx.append(float(f[2]))
                except ValueError:
                    continue
                for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            os.unlink(self.path)
        except Exception:
            pass
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), samples=len(sm))
        out["reasons"] = sorted(reasons)
        return out


# ------------------------------------------------------------------------------------------------------
# reference arm / CPU baseline: the reference's own PyTorch code on the host cores
# ------------------------------------------------------------------------------------------------------
def cpu_reference_stepper(N, D, B, epoch):
    """Returns (step_fn, kind).  kind 'reference' = unmodified reference module from the driver-provided install
    (baseline/_ref) with the six 'cuda:0' literals rewritten to 'cpu'; 'port' = the oracle restatement."""
    from oracle import gptst_oracle as O
    from oracle.ref_import import load_reference, reference_root
    cfg = make_cfg(N, D, "cpu")
    x = torch.randn(B, T_STEPS, N, 3, generator=torch.Generator().manual_seed(0))
    kl = torch.nn.KLDivLoss(reduction="sum")
    root = reference_root()
    if root is not None:
        ref = load_reference("cpu", root)
        torch.manual_seed(0)
        m = ref.GPTST_Model(cfg)
        run_init(m, 0)
        opt = torch.optim.Adam(m.parameters(), lr=3e-3, eps=1e-8)

        def step():
            opt.zero_grad()
            o, _, mask, prob, hs = m(x, x, 1, epoch)
            loss = ((o - x[..., :1]) * mask).abs().mean()
            if epoch > 10:
                loss = loss + 0.1 * kl(prob.log(), hs)
            loss.backward()
            torch.nn.utils.clip_grad_norm_(m.parameters(), 5)
            opt.step()
            return float(loss.detach())
        return step, "reference"
    import random
    P = {k: v.requires_grad_() for k, v in O.init_params(cfg, 0).items()}
    opt = torch.optim.Adam(list(P.values()), lr=3e-3, eps=1e-8)
    gen, pyr = torch.Generator().manual_seed(0), random.Random(0)

    def step():
        opt.zero_grad()
        n = B * T_STEPS * N
        dr = O.Draws.sample(n, n, cfg.HS, epoch > cfg.change_epoch, gen, pyr)
        outs = O.model_forward(P, cfg, x, epoch, dr)
        loss = O.synthetic_loss(outs, x, epoch)
        loss.backward()
        torch.nn.utils.clip_grad_norm_([p for p in P.values() if p.grad is not None], 5)
        opt.step()
        return float(loss.detach())
    return step, "port"


def time_cpu(step, steps, warmup, budget_s=None):
    for _ in range(warmup):
        step()
    t0 = time.perf_counter()
    done = 0
    for _ in range(steps):
        step()
        done += 1
        if budget_s is not None and time.perf_counter() - t0 > budget_s:
            break
    return (time.perf_counter() - t0) / done, done


def run_reference_arm(args):
    """The reference's own CPU implementation of the step on the host cores, on OUR arm's config: same workload dict, same
    GLOBAL batch as our arm runs at this --gpus (weak scaling: 64 x N; strong: 128).  Rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    N, D, _ = WORKLOADS[args.workload]
    world = max(1, args.gpus)
    Bg = per_gpu_batch(args.workload, world, args.batch)
    B = Bg * world                                    # the CPU runs the whole global batch of one step
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    step, kind = cpu_reference_stepper(N, D, B, args.epoch)
    sec, done = time_cpu(step, args.steps, args.warmup, budget_s=150.0)
    value = B / sec
    line = {
        "impl": "reference", "metric": "pretrain samples/sec", "value": value, "unit": "samples/s", "n_gpus": args.gpus,
        "steps": done, "warmup": args.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True,
        "scaling": "strong" if args.workload in STRONG else "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args.workload, Bg, world, args.epoch),
        "cpu_baseline": {"value": value, "unit": "samples/s", "cores": cores, "kind": kind,
                         "sample": f"{done} full training steps at global batch {B} after {args.warmup} warm-up, all {cores} host threads"},
        "e2e": {"value": value, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)
    return 0


L2_NOTE = ("no flush between steps: one step touches ~2 GB of activations/gradients >> 126 MB L2; "
           "kernel roofline timed on rotating buffer sets > L2")


def workload_config(name, B, world, epoch):
    """The SAME dict on both arms (ours / --impl reference): the driver compares them key by key."""
    N, D, _ = WORKLOADS[name]
    return {"l2": L2_NOTE, "workload": f"{name}: GPT-ST pretrain step, N={N}, T=12, D={D}, batch {B}/GPU (BASELINE.json configs[1] geometry)"
            if name == "pems08" else f"{name}: GPT-ST pretrain step, N={N}, T=12, D={D}, batch {B}/GPU",
            "global_batch": B * world, "per_gpu_batch": B, "num_nodes": N, "hidden_dim": D, "epoch_arg": epoch,
            "mask_phase": "adaptive+KL" if epoch > 10 else "random", "optimizer": "Adam lr 3e-3, clip_grad_norm 5",
            "parallelism": f"dp{world}"}


# ------------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------------
def time_graph_replays(fns, iters=20):
    """Median CUDA-event time of replaying one captured graph per rotating input set (fns[i]() is captured once).  The product
    runs these kernels inside the captured step; timing eager Python calls instead leaves host gaps (tensor allocation, ctypes)
    between event record and launch that are not kernel time."""
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        for f in fns:
            f()                                  # warm-up / allocator
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graphs = []
    for f in fns:
        g_ = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_):
            f()
        graphs.append(g_)
    times = []
    for i in range(3 + iters):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        graphs[i % len(graphs)].replay()
        e1.record()
        e1.synchronize()
        if i >= 3:
            times.append(e0.elapsed_time(e1))
    return statistics.median(times)


def cap_forward_roofline(N, D, B, iters=20):
    """cap forward (GPTST.py:100-141) timed alone with CUDA events on the launching stream, on rotating buffer
    sets larger than L2 so every launch reads x from HBM.  Algorithmic bytes: 4*B*T*N*(2D+H)."""
    from gptst_b200 import ops
    H, HT, d, ds, T = 10, 16, 16, 4, T_STEPS
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(0)
    nset = max(3, int(400e6 // (2 * 4 * B * T * N * D)) + 1)
    xs = [torch.randn(B, T, N, D, device=dev, generator=g) for _ in range(nset)]
    Wp = torch.randn(D, D, device=dev, generator=g) * D ** -0.5
    bp = torch.rand(D, device=dev, generator=g)
    dadj = torch.randn(B, T, H, N, device=dev, generator=g)
    dyn = torch.randn(B, HT, T * H, device=dev, generator=g) * 0.3
    Wn = torch.randn(N, D, D, device=dev, generator=g) * D ** -0.5
    bn = torch.rand(N, D, device=dev, generator=g)
    prec = ops.default_precision()
    fused = ops.cap_fused_enabled(N, D, H, T, prec)
    # the fragment-ordered copy of W_n is parameter-side work (like the tables themselves): the model packs it on the block's
    # table stream, off the main chain, so it is prepared outside the timed chain here as well
    wnf = ops.cap_pack_wn(Wn) if fused else None
    for x in xs:
        x.requires_grad_(True)          # the TRAINING flavour: `recon` is written for the backward, as inside the timed step
    with torch.enable_grad():
        ms = time_graph_replays([(lambda x=x: ops.cap_core(x, Wp, bp, dadj, dyn, Wn, bn, 2, prec, wnf=wnf)) for x in xs], iters)
    algo = 4 * B * T * N * (2 * D + H)
    return algo, ms


def hypertem_forward_time(N, D, B, iters=20):
    from gptst_b200 import ops
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(1)
    nset = max(3, int(400e6 // (2 * 4 * B * T_STEPS * N * D)) + 1)
    xs = [torch.randn(B, T_STEPS, N, D, device=dev, generator=g) for _ in range(nset)]
    Mn = torch.randn(N, 12, 12, device=dev, generator=g) * 0.2
    W = torch.randn(B, 12, D, D, device=dev, generator=g) * D ** -0.5
    b = torch.rand(B, 12, D, device=dev, generator=g)
    prec = ops.default_precision()
    with torch.no_grad():
        ms = time_graph_replays([(lambda x=x: ops.hypertem_core(x, Mn, W, b, prec)) for x in xs], iters)
    return 8 * B * T_STEPS * N * D, ms


def kernel_rooflines(N, D, B, peak, iters=20):
    """Other kernels of the step timed alone (CUDA events, graph replays, rotating buffer sets > L2) against their algorithmic
    bytes: the fused hyperTem backward on the main chain (reads dOut, writes deb + dret: 3A; SURVEY.md 8d counts 12*B*T*N*D = 3A
    for the whole hyperTem backward), its two side-stream parameter-gradient kernels, and cap's node-grouped projection backward."""
    from gptst_b200 import _lib, ops
    dev = torch.device("cuda")
    g = torch.Generator(device="cuda").manual_seed(2)
    A = 4 * B * T_STEPS * N * D
    nset = max(3, int(400e6 // (5 * A)) + 1)
    sets = [[torch.randn(B, T_STEPS, N, D, device=dev, generator=g) for _ in range(4)] for _ in range(nset)]
    prec = ops.default_precision()
    out = {}

    def timeit(fn):
        return time_graph_replays([(lambda st=st: fn(st)) for st in sets], iters)

    def entry(ms, algo, launches):
        return {"ms": ms, "algorithmic_bytes": algo, "achieved": algo / (ms * 1e-3) / 1e9, "frac": algo / (ms * 1e-3) / 1e9 / peak,
                "unit": "GB/s", "launches_per_step": launches}

    with torch.no_grad():
        if ops.hypertem_fused_enabled(D, T_STEPS, prec):
            L = _lib.lib()
            W = torch.randn(B, 12, D, D, device=dev, generator=g) * D ** -0.5
            Mn = torch.randn(N, 12, 12, device=dev, generator=g) * 0.2
            nb = L.gptst_hypertem_wfrag_bytes(B * 12)
            wf, wb = (torch.empty(nb, dtype=torch.uint8, device=dev) for _ in range(2))
            st0 = torch.cuda.current_stream().cuda_stream
            _lib.check(L.gptst_hypertem_pack_w(W.data_ptr(), wf.data_ptr(), wb.data_ptr(), B * 12, st0), "pack")
            npad = (N + 15) // 16 * 16
            mask = torch.randint(-2 ** 31, 2 ** 31 - 1, (B * 12, npad, 2), device=dev, dtype=torch.int32)
            sp_w, sp_m = L.gptst_gproj_splits(B * 12, N, D), L.gptst_tmix_bwd_splits(B, N)
            dWp = torch.empty((sp_w, B * 12, D, D), device=dev)
            dbp = torch.empty((sp_w, B * 12, D), device=dev)
            dMp = torch.empty((sp_m, N, 12, 12), device=dev)
            cs = lambda: torch.cuda.current_stream().cuda_stream
            ms = timeit(lambda s: _lib.check(L.gptst_hypertem_bwd(s[0].data_ptr(), mask.data_ptr(), Mn.data_ptr(), wb.data_ptr(), s[1].data_ptr(),
                                                                  s[2].data_ptr(), B, 12, N, D, cs()), "bwd"))
            out["hypertem_bwd_fused"] = entry(ms, 3 * A, 8)
            ms = timeit(lambda s: _lib.check(L.gptst_hypertem_dw(s[0].data_ptr(), mask.data_ptr(), s[3].data_ptr(), dWp.data_ptr(), dbp.data_ptr(),
                                                                 B, 12, N, D, npad, sp_w, cs()), "dw"))
            out["hypertem_dw_side_stream"] = entry(ms, 2 * A + 4 * B * 12 * D * D, 8)
            ms = timeit(lambda s: _lib.check(L.gptst_tmix_dM2(s[2].data_ptr(), s[3].data_ptr(), dMp.data_ptr(), B, 12, N, D, sp_m, cs()), "dM"))
            out["hypertem_dM_side_stream"] = entry(ms, 2 * A, 8)
        Wn = torch.randn(N, D, D, device=dev, generator=g) * D ** -0.5
        ms = timeit(lambda s: ops.gproj_bwd(s[0], s[1], s[2], Wn, node_grouped=True, act=True, prec=prec, want_dres=True, sum_parts=False))
        out["gproj_bwd_node_grouped"] = entry(ms, 5 * A, 4)
    return out


def ncu_traffic(workload, B):
    """dram__bytes_read.sum + dram__bytes_write.sum of the kernels of ONE cap forward, per launch of the chain, from the
    committed `ncu --set full` summary of `python bench.py --cap-only` (tools/ncu_cap_traffic.sh writes
    profiles/ncu_cap_forward_traffic.json with the kernel list, the geometry and the commit it was taken at).  None when no
    capture exists for this geometry: bench.py cannot run ncu on itself, so this is a committed measurement, not a live one."""
    try:
        d = json.load(open(os.path.join(REPO, "profiles", "ncu_cap_forward_traffic.json")))
        if d.get("workload") == workload and int(d.get("batch", -1)) == B:
            return float(d["dram_bytes_per_chain"]), d.get("source")
    except Exception:
        pass
    return None, None


def measured_peak():
    p = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def run_ours(args):
    import torch.distributed as dist
    from gptst_b200 import dp, ops
    from gptst_b200.GPTST import GPTST_Model
    from gptst_b200.losses import probe_loss

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: the product path needs a CUDA device (no CPU fallback); use --impl reference for the CPU arm")
    rank, local, world = dp.init_from_env()
    if world != args.gpus and world > 1:
        args.gpus = world
    dev = torch.device("cuda", local)
    torch.cuda.set_device(dev)
    N, D, _ = WORKLOADS[args.workload]
    B = per_gpu_batch(args.workload, world, args.batch)
    epoch = args.epoch
    if args.cap_only:              # the roofline leg alone (what tools/ncu_cap_traffic.sh profiles); never a bench line
        algo, cap_ms = cap_forward_roofline(N, D, B, iters=6)
        print(json.dumps({"cap_only": True, "workload": args.workload, "batch": B, "algorithmic_bytes": algo, "ms": cap_ms}), flush=True)
        return 0
    cfg = make_cfg(N, D, "cuda")
    model = GPTST_Model(cfg).to(dev)
    run_init(model, 0)
    dp.broadcast_parameters(model)
    from gptst_b200.train import PretrainStep
    # two buckets (decoder | encoder + scorer): rank-summed gradients stay in the flat all-reduce buffers and the fused optimiser
    # reads them there with 1/world folded into its clip coefficient; the decoder bucket's collective overlaps the encoder backward
    if world > 1 and not args.eager:
        reducer = dp.BucketedGradAllReduce(dp.pretrain_buckets(model))
    else:
        reducer = dp.FlatGradAllReduce(model.parameters()) if world > 1 else None
    stepper = PretrainStep(model, lr=3e-3, max_grad_norm=5.0, loss="probe", use_graph=not args.eager, reducer=reducer)

    # synthetic inputs (standard normal); distinct per rank
    gcpu = torch.Generator().manual_seed(100 + rank)
    nbuf = 4
    host = [torch.randn(B, T_STEPS, N, 3, generator=gcpu).pin_memory() for _ in range(nbuf)]
    resident = [h.to(dev) for h in host]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(nsteps, e2e):
        barrier()
        l0, r0 = ops.launch_count(), stepper.replays
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if args.profiler_range and not e2e:
            torch.cuda.profiler.start()     # `ncu --profile-from-start off` then profiles exactly the timed graph replays
        e0.record()
        last = None
        if e2e and not args.eager:
            # public API with HOST buffers, software-pipelined the way a training loop logs: the H2D copy of batch i+1 is
            # announced with `prefetch=` and runs under step i, and the loss of step i is read back (D2H into pinned memory,
            # then float()) while step i+1 is already queued -- every step's input copy and loss read stay inside the timed region
            loss_host = [torch.zeros((), dtype=torch.float32).pin_memory() for _ in range(2)]
            evs = [torch.cuda.Event(), torch.cuda.Event()]
            for i in range(nsteps):
                dl = stepper(host[i % nbuf], epoch, prefetch=host[(i + 1) % nbuf])
                loss_host[i % 2].copy_(dl, non_blocking=True)
                evs[i % 2].record()
                if i > 0:
                    evs[(i - 1) % 2].synchronize()
                    last = float(loss_host[(i - 1) % 2])
            evs[(nsteps - 1) % 2].synchronize()
            last = float(loss_host[(nsteps - 1) % 2])
        else:
            for i in range(nsteps):
                if e2e:
                    # eager mode: H2D of this step's batch from pinned memory, D2H read of the loss
                    last = float(stepper(host[i % nbuf], epoch))
                else:
                    last = stepper(resident[i % nbuf], epoch)
        e1.record()
        torch.cuda.synchronize()
        if args.profiler_range and not e2e:
            torch.cuda.profiler.stop()
        ms = e0.elapsed_time(e1)
        barrier()
        if world > 1:
            t = torch.tensor([ms], device=dev)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        launches = (ops.launch_count() - l0) + (stepper.replays - r0) * stepper.launches_per_step
        return ms, launches, float(last)

    import random
    random.seed(1234 + rank)
    torch.manual_seed(1234 + rank)
    for i in range(max(4, args.warmup)):          # >= 3 eager warm-up steps + graph capture + first replay
        stepper(resident[i % nbuf], epoch)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ms, launches, last_loss = timed(args.steps, e2e=False)
    ms_e2e, _, _ = timed(args.steps, e2e=True)
    clocks = sampler.stop() if rank == 0 else {}
    value = world * B * args.steps / (ms * 1e-3)
    value_e2e = world * B * args.steps / (ms_e2e * 1e-3)
    h2d_bytes = resident[0].numel() * 4 * world

    line = None
    if rank == 0:
        if args.no_rooflines:      # A/B runs: step throughput only
            print(json.dumps({"value": value, "ms_per_step": ms / args.steps, "e2e": value_e2e, "e2e_ms_per_step": ms_e2e / args.steps,
                              "last_loss": last_loss, "gpu_launches": launches, "clocks": clocks,
                              "env": {k: v for k, v in os.environ.items() if k.startswith("GPTST_B200_")}}), flush=True)
    if rank == 0 and not args.no_rooflines:
        peak, peak_src = measured_peak()
        algo, cap_ms = cap_forward_roofline(N, D, B)
        ht_algo, ht_ms = hypertem_forward_time(N, D, B)
        achieved = algo / (cap_ms * 1e-3) / 1e9
        step_bytes = (60 * 4 * B * T_STEPS * N * D) + (8 * 4 * B * T_STEPS * N * 10)
        line = {
            "metric": "pretrain samples/sec", "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps,
            "warmup": max(4, args.warmup), "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if args.workload in STRONG else "weak",
            "vs_baseline": None, "dtype": "f32 (tensor-core contractions as three-term fp16 splits, fp32 accumulate; fp32 FMA elsewhere)"
            if ops.default_precision() == 3 else "f32 storage, single-term fp16/tf32 tensor-core operands, fp32 accumulate",
            "data": "synthetic", "config": workload_config(args.workload, B, world, epoch),
            "e2e": {"value": value_e2e, "unit": "samples/s", "h2d_bytes_per_step": h2d_bytes,
                    "d2h_bytes_per_step": 4 * world, "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": launches, "step_mode": "eager" if args.eager else "one CUDA graph per step (gptst_b200.train.PretrainStep)",
            "clocks": clocks,
            "roofline": {"bound": "hbm", "kernel": ("cap forward, training flavour (gptst_cap_route_fwd + gptst_cap_hop_ev + gptst_cap_recon_proj)"
                                                    if ops.cap_fused_enabled(N, D, 10, T_STEPS, ops.default_precision()) else
                                                    "cap forward (gptst_cap_route_fwd + gptst_cap_hop_e1 + gptst_cap_recon_hop + gptst_gproj_fwd)")
                         + ", the hypergraph + node-adaptive GCN block named by BASELINE.json", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(args.workload, B)[0], "traffic_source": ncu_traffic(args.workload, B)[1],
                         "algorithmic_bytes": algo, "ms": cap_ms, "peak_source": peak_src},
            "roofline_hypertem_fwd": {"achieved": ht_algo / (ht_ms * 1e-3) / 1e9, "unit": "GB/s", "ms": ht_ms,
                                      "algorithmic_bytes": ht_algo, "frac": ht_algo / (ht_ms * 1e-3) / 1e9 / peak},
            "roofline_kernels": kernel_rooflines(N, D, B, peak),
            "step_roofline": {"algorithmic_bytes": step_bytes, "frac": step_bytes / (ms / args.steps * 1e-3) / 1e9 / peak},
            "last_loss": last_loss,
        }
        if world == 1 and not args.no_cpu_baseline:
            cores = os.cpu_count() or 1
            torch.set_num_threads(cores)
            step, kind = cpu_reference_stepper(N, D, B, epoch)
            sec, done = time_cpu(step, 12, 1, budget_s=15.0)
            line["cpu_baseline"] = {"value": B / sec, "unit": "samples/s", "cores": cores, "kind": kind,
                                    "sample": f"{done} full training steps at batch {B} (after 1 warm-up) on {cores} host threads"}
        print(json.dumps(line), flush=True)
    if world > 1:
        # a CUDA graph that captured NCCL kernels must be released before the communicator goes away; destroying the
        # process group with the graph alive dead-locks (observed), so: drop graphs, sync, barrier, hard-exit.
        stepper._graphs.clear()
        import gc
        gc.collect()
        torch.cuda.synchronize()
        dist.barrier()
        torch.cuda.synchronize()
        sys.stdout.flush()
        # a rank whose barrier kernel has finished must not tear its context down while a peer's is still reading its
        # buffers over NVLink (observed: one rank exits, the other hangs in the barrier forever): give the peers a moment
        time.sleep(3.0)
        os._exit(0)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="pems08", choices=sorted(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="per-GPU batch override")
    ap.add_argument("--epoch", type=int, default=200, help="epoch argument passed to the model (mask phase)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--profiler-range", action="store_true",
                    help="bracket the timed (resident-input) steps with cudaProfilerStart/Stop: `ncu --profile-from-start off "
                         "--metrics gpu__time_duration.sum ... python bench.py --steps 2 --warmup 1 --profiler-range --no-rooflines "
                         "--no-cpu-baseline` lists exactly the kernels of the timed graph replays (never a bench value)")
    ap.add_argument("--cap-only", action="store_true", help="run only the cap-forward roofline leg (for ncu captures; not a bench line)")
    ap.add_argument("--no-rooflines", action="store_true", help="A/B runs: print only the step numbers (not a bench line)")
    ap.add_argument("--eager", action="store_true", help="time the eager step (what an unmodified Run.py loop launches) instead of the graph")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_ours(args)


if __name__ == "__main__":
    sys.exit(main())
