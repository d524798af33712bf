"""Run the heavy blocks alone (PEMS08 geometry) so that ncu can capture single kernels cheaply.
    ncu --set full -k regex:cap_route_fwd -c 1 ... python tools/prof_blocks.py [cap|hypertem|all] [B] [prec]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gptst_b200 import ops

which = sys.argv[1] if len(sys.argv) > 1 else "all"
B = int(sys.argv[2]) if len(sys.argv) > 2 else 64
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 3
N, D, T, H, HT = int(os.environ.get("N", 170)), int(os.environ.get("D", 64)), 12, 10, 16
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
x = torch.randn(B, T, N, D, device=dev, generator=g, requires_grad=True)
go = torch.randn(B, T, N, D, device=dev, generator=g)
def timeit(fn, n=10):
    for _ in range(3): fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); e1.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
if which in ("cap", "all"):
    Wp = (torch.randn(D, D, device=dev, generator=g) * D ** -0.5).requires_grad_()
    bp = torch.rand(D, device=dev, generator=g).requires_grad_()
    dadj = torch.randn(B, T, H, N, device=dev, generator=g).requires_grad_()
    dyn = (torch.randn(B, HT, T * H, device=dev, generator=g) * 0.3).requires_grad_()
    Wn = (torch.randn(N, D, D, device=dev, generator=g) * D ** -0.5).requires_grad_()
    bn = torch.rand(N, D, device=dev, generator=g).requires_grad_()
    def f():
        with torch.no_grad(): ops.cap_core(x, Wp, bp, dadj, dyn, Wn, bn, 2, prec)
    def fb():
        o, _ = ops.cap_core(x, Wp, bp, dadj, dyn, Wn, bn, 2, prec); o.backward(go)
    print(f"cap fwd {timeit(f):.1f} us   fwd+bwd {timeit(fb):.1f} us  (B={B} N={N} D={D} prec={prec})")
if which in ("hypertem", "all"):
    Mn = (torch.randn(N, 12, 12, device=dev, generator=g) * 0.2).requires_grad_()
    W = (torch.randn(B, 12, D, D, device=dev, generator=g) * D ** -0.5).requires_grad_()
    b = torch.rand(B, 12, D, device=dev, generator=g).requires_grad_()
    def f():
        with torch.no_grad(): ops.hypertem_core(x, Mn, W, b, prec)
    def fb():
        ops.hypertem_core(x, Mn, W, b, prec).backward(go)
    print(f"hyperTem fwd {timeit(f):.1f} us   fwd+bwd {timeit(fb):.1f} us")
torch.cuda.synchronize()
