"""Packed vs unpacked attention on the fusion-pass shapes (4B = 512 rows of 40 tokens)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from efficientvlm_b200 import kernels as K
dev = torch.device("cuda", 0); bf16 = torch.bfloat16
def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
B, H, L, E = 512, 12, 40, 768
q = torch.randn(B * L, 3 * E, device=dev).to(bf16)
mask = torch.zeros(B, L, device=dev)
groups = (B + 2) // 3
idx = torch.arange(groups * 3, dtype=torch.int32).view(groups, 3)
pack = torch.where(idx < B, idx, torch.full_like(idx, -1)).to(dev).contiguous()
for name, kw in (("self unpacked", {}), ("self packed x3", dict(pack_items=pack, pack_own_kv=True))):
    f = lambda: K.attention_fwd(q[:, :E], q[:, E:2 * E], q[:, 2 * E:], B, H, L, L, 0.125, key_mask=mask, want_probs=True, dropout_p=0.1, seed=3, stream_id=0, **kw)
    tf = timeit(f)
    c, P, lse = f()
    dc = torch.randn_like(c); dP = torch.randn_like(P) * 1e-3; dqkv = torch.empty_like(q)
    g = lambda: K.attention_bwd(q[:, :E], q[:, E:2 * E], q[:, 2 * E:], c, lse, dc, dqkv[:, :E], dqkv[:, E:2 * E], dqkv[:, 2 * E:], B, H, L, L, 0.125,
                                probs=P, dprobs=dP, key_mask=mask, dropout_p=0.1, seed=3, stream_id=0, **kw)
    print("%-16s fwd %7.1f us   bwd(+delta) %7.1f us" % (name, tf, timeit(g)))
