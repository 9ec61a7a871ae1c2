"""Attention microbenchmark on the GD step's shapes (CUDA events, inputs rotated through > L2-size buffers)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientvlm_b200 import kernels as K
dev = torch.device("cuda", 0)
bf16 = torch.bfloat16
SHAPES = [("vit_self", 128, 12, 197, 197, True), ("text_self", 128, 12, 40, 40, True), ("cross", 128, 12, 40, 197, True),
          ("itm_self", 384, 12, 40, 40, True), ("itm_cross", 384, 12, 40, 197, True), ("vit_noprobs", 128, 12, 197, 197, False)]
only = sys.argv[1:] 
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
for name, B, H, Lq, Lk, probs in SHAPES:
    if only and name not in only: continue
    E = H * 64
    NB = 4   # rotate buffers so that inputs are not L2-resident
    qs = [torch.randn(B * Lq, E, device=dev).to(bf16) for _ in range(NB)]
    ks = [torch.randn(B * Lk, E, device=dev).to(bf16) for _ in range(NB)]
    vs = [torch.randn(B * Lk, E, device=dev).to(bf16) for _ in range(NB)]
    mask = torch.zeros(B, Lk, device=dev)
    i = [0]
    def fwd():
        j = i[0] = (i[0] + 1) % NB
        return K.attention_fwd(qs[j], ks[j], vs[j], B, H, Lq, Lk, 0.125, key_mask=mask, want_probs=probs, dropout_p=0.1 if Lq == 40 else 0.0, seed=5, stream_id=1)
    t_f = timeit(fwd)
    ctx, P, lse = fwd()
    dctx = torch.randn_like(ctx)
    dP = torch.randn_like(P) * 1e-3 if probs else None
    dq, dk, dv = torch.empty_like(qs[0]), torch.empty_like(ks[0]), torch.empty_like(vs[0])
    j = i[0]
    def bwd():
        K.attention_bwd(qs[j], ks[j], vs[j], ctx, lse, dctx, dq, dk, dv, B, H, Lq, Lk, 0.125, probs=P, dprobs=dP, key_mask=mask,
                        dropout_p=0.1 if Lq == 40 else 0.0, seed=5, stream_id=1)
    t_b = timeit(bwd)
    io_f = (B * Lq * E * 2 * 2 + 2 * B * Lk * E * 2 + (B * H * Lq * Lk * 4 if probs else 0)) / 1e6
    print("%-12s B=%d Lq=%d Lk=%d probs=%d  fwd %7.1f us (%.0f MB -> %.0f GB/s)   bwd %7.1f us" % (name, B, Lq, Lk, probs, t_f, io_f, io_f / t_f * 1e3, t_b))
