"""Per-kernel device time (torch.profiler / CUPTI, warm) of one forward+backward of the cap and hyperTem blocks.
usage: python tools/prof_kernels.py [B=64] [N=170]"""
import os, sys, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from gptst_b200 import ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 170
D, T, H, HT, prec = 64, 12, 10, 16, 3
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
R = lambda *s, sc=1.0: (torch.randn(*s, device=dev, generator=g) * sc).requires_grad_()
x, go = R(B, T, N, D), torch.randn(B, T, N, D, device=dev, generator=g)
Wp, bp, dadj, dyn = R(D, D, sc=D ** -0.5), R(D), R(B, T, H, N), R(B, HT, T * H, sc=0.3)
Wn, bn = R(N, D, D, sc=D ** -0.5), R(N, D)
Mn, W, b = R(N, 12, 12, sc=0.2), R(B, 12, D, D, sc=D ** -0.5), R(B, 12, D)
def cap():
    o, _ = ops.cap_core(x, Wp, bp, dadj, dyn, Wn, bn, 2, prec); o.backward(go)
def ht():
    ops.hypertem_core(x, Mn, W, b, prec).backward(go)
for name, fn in (("cap", cap), ("hyperTem", ht)):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        fn(); torch.cuda.synchronize()
    rows = [(k.start_ns(), k.duration_ns() / 1e3, k.name()) for k in prof.profiler.kineto_results.events()
            if "CUDA" in str(k.device_type()) and k.duration_ns() > 0]
    rows.sort()
    print(f"---- {name} fwd+bwd: {sum(r[1] for r in rows):.1f} us in {len(rows)} kernels")
    for _, du, nm in rows:
        print(f"  {du:7.1f} us  {re.sub(r'[(<].*', '', nm)[:70]}")
