import sys, torch
sys.path.insert(0, ".")
from efficientvlm_b200 import kernels as K, ops
from efficientvlm_b200.eff_bert import BertConfig, BertLayer
from efficientvlm_b200.eff_vit import CLIPEncoderLayer
from oracle.det_init import det_init_module_
def rel(a, b): return float((a.float() - b.float()).norm() / b.float().norm())
H, I, nh, B, N = 768, 3072, 12, 4, 50
g = torch.Generator().manual_seed(17)
vit = CLIPEncoderLayer(H, "quick_gelu", nh, 0.0, I).eval(); det_init_module_(vit); vit.cuda()
cfg = BertConfig(vocab_size=64, hidden_size=H, num_hidden_layers=1, num_attention_heads=nh, intermediate_size=I, max_position_embeddings=64)
cfg.fusion_layer, cfg.encoder_width = 1, H
bert = BertLayer(cfg, 0).eval(); det_init_module_(bert); bert.cuda()
x = torch.randn(B, N, H, generator=g).cuda()
z = torch.rand(2, I, generator=g); z[torch.rand(2, I, generator=g) < 0.17] = 0; z = z.cuda()
out = {}
for skip in (False, True, False):
    ops.ZERO_SKIP = skip
    with torch.no_grad():
        hv = vit(x, None, False, mlp_z=z[0].view(1, 1, I))[0]
        hb = bert(x, attention_mask=None, mlp_z=z[1].view(1, 1, I))[0]
        hb2 = bert(hv, attention_mask=None, mlp_z=z[1].view(1, 1, I))[0]
    out.setdefault(skip, []).append((hv, hb, hb2))
d0, d1 = out[False]; s = out[True][0]
print("dense vs dense (determinism):", [rel(a, b) for a, b in zip(d0, d1)])
print("skip vs dense: vit %.3e  bert(x) %.3e  bert(vit) %.3e" % tuple(rel(a, b) for a, b in zip(s, d0)))
print("norms", [float(t.norm()) for t in d0], "x", float(x.norm()), "absmax hv", float(d0[0].abs().max()))
