"""Per-parameter gradient error of the CUDA model against the fp64 oracle at the BASELINE geometry (batch 64, N=170, D=64):
max |err| / abs-max of the reference gradient, for each parameter tensor.  usage: python tools/grad_err_report.py [epoch]"""
import os, random, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from oracle import gptst_oracle as O
from util import make_cfg
from test_parity_r2_gpu import build, inject, clear, oracle_step

epoch = int(sys.argv[1]) if len(sys.argv) > 1 else 1
cfg = make_cfg(N=170, D=64)
B = 64
m, P = build(cfg, seed=3)
src = torch.randn(B, 12, 170, 3, generator=torch.Generator().manual_seed(2))
n = B * 12 * 170
draws = O.Draws.sample(n, n, cfg.HS, epoch > cfg.change_epoch, torch.Generator().manual_seed(40 + epoch), random.Random(40 + epoch))
label_c = None
if epoch > cfg.change_epoch:
    with torch.no_grad():
        prob = O.encoder(P, cfg, src, epoch, draws)[2]
    label_c = torch.sort(prob, dim=-1, descending=True)[1][..., 0]
    m.encoder.label_c_override = label_c.cuda()
inject(m, cfg, draws, epoch, n)
outs = m(src.cuda(), None, 1, epoch)
O.synthetic_loss(outs, src.cuda(), epoch).backward()
clear(m)
ref, ref_loss, gref = oracle_step(P, cfg, src, epoch, draws, label_c, torch.float64)
rows = []
for k, p in m.named_parameters():
    if p.grad is None:
        continue
    g, r = p.grad.double().cpu(), gref[k]
    rows.append(((g - r).abs().max().item() / max(1e-30, r.abs().max().item()), r.abs().max().item(), k))
rows.sort(reverse=True)
print(f"HTEM={os.environ.get('GPTST_B200_HTEM', 'fused')} epoch {epoch}: worst 12 of {len(rows)} gradients (max err / abs-max, abs-max, name)")
for e, a, k in rows[:12]:
    print(f"  {e:9.2e}  {a:9.2e}  {k}")
