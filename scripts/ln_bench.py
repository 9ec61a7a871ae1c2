import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientvlm_b200 import kernels as K
dev = torch.device("cuda", 0)
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for rows, H, dyb in ((25216, 768, True), (5120, 768, False), (15360, 768, False), (25216, 768, False)):
    NB = 3
    xs = [torch.randn(rows, H, device=dev) for _ in range(NB)]
    dys = [torch.randn(rows, H, device=dev).to(torch.bfloat16 if dyb else torch.float32) for _ in range(NB)]
    dres = [torch.randn(rows, H, device=dev) for _ in range(NB)]
    w, b = torch.randn(H, device=dev), torch.randn(H, device=dev)
    _, _, mean, rstd = K.layernorm_fwd(xs[0], w, b, 1e-5, want_f32=True)
    dg, dbt = torch.zeros(H, device=dev), torch.zeros(H, device=dev)
    i = [0]
    def f():
        j = i[0] = (i[0] + 1) % NB
        K.layernorm_bwd(dys[j], xs[j], w, mean, rstd, dres=dres[j], want_f32=True, want_bf16=True, dgamma=dg, dbeta=dbt)
    def g():
        j = i[0] = (i[0] + 1) % NB
        K.layernorm_fwd(xs[j], w, b, 1e-5, want_f32=True, want_bf16=True)
    tb, tf = timeit(f), timeit(g)
    mb_b = rows * H * ((2 if dyb else 4) + 4 + 4 + 4 + 2) / 1e6
    mb_f = rows * H * (4 + 4 + 2) / 1e6
    print("rows %6d H %d dy_bf16 %d: bwd %6.1f us (%.0f GB/s)  fwd %6.1f us (%.0f GB/s)" % (rows, H, dyb, tb, mb_b / tb * 1e3, tf, mb_f / tf * 1e3))
