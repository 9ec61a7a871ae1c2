# GPU debug: zero-skip FFN vs dense at the GEMM level
import sys, torch
sys.path.insert(0, ".")
from efficientvlm_b200 import kernels as K, ops
bf16, f32 = torch.bfloat16, torch.float32
g = torch.Generator().manual_seed(3)
T, H, I = 200, 768, 3072
def rel(a, b): return float((a.float() - b.float()).norm() / b.float().norm())
x16 = (torch.randn(T, H, generator=g) * 1.0).to(bf16).cuda()
W1 = (torch.randn(I, H, generator=g) * 0.03).to(bf16).cuda()
W2 = (torch.randn(H, I, generator=g) * 0.03).to(bf16).cuda()
b1 = (torch.randn(I, generator=g) * 0.1).cuda()
b2 = (torch.randn(H, generator=g) * 0.1).cuda()
res = torch.randn(T, H, generator=g).cuda()
z = torch.rand(I, generator=g)
z[torch.rand(I, generator=g) < 0.17] = 0
z = z.cuda()
for act, mode in ((1, 1), (2, 2)):
    gd = ops.alloc16(T, I, "cuda"); ud = ops.alloc16(T, I, "cuda")
    K.gemm(x16, W1, gd, T, I, H, bias=b1, act=act, gate=z, gate_mode=mode, aux_out=ud)
    hd = torch.empty(T, H, device="cuda")
    K.gemm(gd, W2, hd, T, H, I, bias=b2, residual=res)
    cp = ops._Compact(z, W1, b1, W2)
    cnt = int(cp.count)
    gs = ops.alloc16(T, I, "cuda"); us = ops.alloc16(T, I, "cuda")
    K.gemm(x16, cp.W1, gs, T, I, H, bias=cp.b1, act=act, gate=cp.z, gate_mode=mode, aux_out=us, n_limit=cp.count)
    hs = torch.empty(T, H, device="cuda")
    K.gemm(gs, cp.W2, hs, T, H, I, bias=b2, residual=res, k_limit=cp.count)
    kept = cp.idx[:cnt].long()
    print("act", act, "count", cnt, "fc1 kept cols rel", rel(gs[:, :cnt], gd[:, kept]), "u rel", rel(us[:, :cnt], ud[:, kept]),
          "tail finite", bool(torch.isfinite(gs[:, cnt:(cnt + 63) // 64 * 64].float()).all()), float(gs[:, cnt:(cnt + 63) // 64 * 64].float().abs().max()),
          "fc2 rel", rel(hs, hd), "fc2-res rel", rel(hs - res, hd - res))
    ref = (gd.float() @ W2.float().t() + b2 + res)
    print("   dense vs torch", rel(hd, ref), " skip vs torch", rel(hs, ref))
