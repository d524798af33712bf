O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_eval_path_gpu.py "tests/test_kernels_gpu.py::test_eval_glue_fusion_matches_reference_formula" -m gpu -q -p no:cacheprovider 2>&1 | grep -v "Warning\|warnings.warn\|run_backward\|^$" | tail -30 > $O/pytest_r02_f4.log
tail -25 $O/pytest_r02_f4.log
python - <<'PY'
import torch, time, sys
sys.path.insert(0,'.')
from torch.profiler import profile, ProfilerActivity
from gptst_b200.fusion import TemporalConvGLU, Fusion
for (B,ci,co) in [(64,64,32),(64,128,128)]:
    l=TemporalConvGLU(3,ci,co).cuda()
    x=torch.randn(B,ci,12,170,device='cuda',requires_grad=True)
    ref=torch.nn.Conv2d(ci,2*co,(3,1),1,padding=[1,0]).cuda()
    for name,f in (("ours",lambda: l(x)),("torch conv only",lambda: ref(x))):
        for _ in range(3): y=f(); y.sum().backward()
        torch.cuda.synchronize()
        e0,e1,e2=[torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e0.record(); y=f(); e1.record(); y.backward(torch.ones_like(y)); e2.record(); torch.cuda.synchronize()
        print(f"GLU tconv {ci}->{co} B={B} {name}: fwd {e0.elapsed_time(e1)*1e3:.1f} us  bwd {e1.elapsed_time(e2)*1e3:.1f} us")
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        y=l(x); y.backward(torch.ones_like(y)); torch.cuda.synchronize()
    print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=8, max_name_column_width=60))
PY
