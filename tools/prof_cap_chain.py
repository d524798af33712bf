"""Per-launch CUDA-event timing of the cap forward chain (route_fwd -> hop_e1 -> recon_hop -> gproj_fwd) on rotating
buffer sets larger than L2.  usage: python tools/prof_cap_chain.py [B=64] [N=170] [prec=3]"""
import os, sys, statistics
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from gptst_b200 import _lib, ops

B = int(sys.argv[1]) if len(sys.argv) > 1 else 64
N = int(sys.argv[2]) if len(sys.argv) > 2 else 170
prec = int(sys.argv[3]) if len(sys.argv) > 3 else 3
D, T, H, HT, R = 64, 12, 10, 16, 2
dev = torch.device("cuda")
g = torch.Generator(device="cuda").manual_seed(0)
nset = max(3, int(400e6 // (2 * 4 * B * T * N * D)) + 1)
xs = [torch.randn(B, T, N, D, device=dev, generator=g) for _ in range(nset)]
recons = [torch.empty(B, T, N, D, device=dev) for _ in range(nset)]
Wp = torch.randn(D, D, device=dev, generator=g) * D ** -0.5
bp = torch.rand(D, device=dev, generator=g)
dadj = torch.randn(B, T, H, N, device=dev, generator=g)
dyn = torch.randn(B, HT, T * H, device=dev, generator=g) * 0.3
Wn = torch.randn(N, D, D, device=dev, generator=g) * D ** -0.5
bn = torch.rand(N, D, device=dev, generator=g)
c = torch.empty(B, T, H, N, device=dev)
s = torch.empty(B, T, H, D, device=dev)
v = torch.empty_like(s)
e1 = torch.empty(B, HT, D, device=dev)
L = _lib.lib()
p = lambda t: t.data_ptr()
st = torch.cuda.current_stream().cuda_stream
names = ["route_fwd", "hop_e1", "recon_hop", "gproj_fwd", "chain"]
times = {n: [] for n in names}
for i in range(25):
    x, recon = xs[i % nset], recons[i % nset]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    ev[0].record()
    _lib.check(L.gptst_cap_route_fwd(p(x), p(Wp), p(bp), p(dadj), p(c), p(s), B, T, N, D, H, R, prec, st), "route")
    ev[1].record()
    _lib.check(L.gptst_cap_hop_e1(p(s), p(dyn), p(e1), B, T, D, H, HT, st), "e1")
    ev[2].record()
    _lib.check(L.gptst_cap_recon_hop(p(c), p(s), p(dyn), p(e1), p(v), p(recon), B, T, N, D, H, HT, st), "recon_hop")
    ev[3].record()
    out = ops.gproj_fwd(recon, Wn, bn, x, node_grouped=True, act=True, prec=prec)
    ev[4].record()
    ev[4].synchronize()
    if i >= 5:
        for k in range(4):
            times[names[k]].append(ev[k].elapsed_time(ev[k + 1]) * 1e3)
        times["chain"].append(ev[0].elapsed_time(ev[4]) * 1e3)
algo = 4 * B * T * N * (2 * D + H)
print(f"cap forward chain B={B} N={N} D={D} prec={prec} (rotating over {nset} buffer sets; events between launches add ~2 us each)")
for n in names:
    print(f"  {n:10s} {statistics.median(times[n]):8.1f} us")
ch = statistics.median(times["chain"])
print(f"  algorithmic bytes {algo} -> {algo / ch / 1e3:.1f} GB/s")
