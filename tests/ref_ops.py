"""TEST-ONLY operator backend: plain-torch (oracle) implementations of the `efficientvlm_b200.ops` / `.kernels` entry points
the product modules call, so the HOST LOGIC (gate routing, mode slicing, hidden-state bookkeeping, loss assembly, state_dict
layout) can be checked on a CPU-only box against the reference-generated goldens.  The product never imports this file and
has no CPU path of its own; `install(monkeypatch)` swaps the functions in for the duration of one test.
"""
import math

import torch
import torch.nn.functional as F

from oracle import xvlm_oracle as O


def _sd(names, tensors, prefix="p"):
    return {prefix + "." + n: t for n, t in zip(names, tensors)}


VIT_NAMES = ["layer_norm1.weight", "layer_norm1.bias", "self_attn.q_proj.weight", "self_attn.q_proj.bias", "self_attn.k_proj.weight",
             "self_attn.k_proj.bias", "self_attn.v_proj.weight", "self_attn.v_proj.bias", "self_attn.out_proj.weight",
             "self_attn.out_proj.bias", "layer_norm2.weight", "layer_norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight",
             "mlp.fc2.bias"]
ATT_NAMES = ["self.query.weight", "self.query.bias", "self.key.weight", "self.key.bias", "self.value.weight", "self.value.bias",
             "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight", "output.LayerNorm.bias"]
FFN_NAMES = ["intermediate.dense.weight", "intermediate.dense.bias", "output.dense.weight", "output.dense.bias", "output.LayerNorm.weight",
             "output.LayerNorm.bias"]


def vit_layer(h, key_mask, head_z, head_layer_z, mlp_z, cfg, params):
    sd = _sd(VIT_NAMES, params)
    mask = None if key_mask is None else key_mask[:, None, None, :]
    out, probs = O.vit_layer(sd, "p", h, cfg.num_heads, mask, head_z, head_layer_z, mlp_z, eps=cfg.eps)
    return out, (probs if cfg.want_probs else None)


def vit_embed(x, patch_w, cls, pos, lnw, lnb, eps=1e-5):
    B = x.shape[0]
    pe = F.conv2d(x, patch_w, stride=patch_w.shape[-1]).flatten(2).transpose(1, 2)
    h = torch.cat([cls.expand(B, 1, -1), pe], 1) + pos[None]
    return F.layer_norm(h, (h.shape[-1],), lnw, lnb, eps)


def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def linear(x, w, b=None, act=0):
    y = F.linear(x, w, b)
    if act == 1:
        y = O.quick_gelu(y)
    elif act == 2:
        y = O.gelu(y)
    return y


def gelu(x):
    return O.gelu(x)


def bert_embed(ids, type_ids, pos_ids, word, type_emb, pos_emb, lnw, lnb, eps, p_drop, past_len=0, padding_idx=None):
    sd = {"e.word_embeddings.weight": word, "e.token_type_embeddings.weight": type_emb, "e.position_embeddings.weight": pos_emb,
          "e.LayerNorm.weight": lnw, "e.LayerNorm.bias": lnb}
    return O.bert_embeddings(sd, "e", ids, type_ids, pos_ids, past_len, eps, padding_idx)


def bert_layer(x, key_mask, enc, enc_mask, self_head_z, cross_head_z, mlp_z, past_kv, cfg, params, enc_index=None):
    import efficientvlm_b200.ops as _ops
    if isinstance(enc_index, _ops.UniformGroups):   # equal consecutive groups: a launch-geometry hint, same result as the explicit index
        enc_index = enc_index.index
    if isinstance(enc_index, tuple):   # (row -> image index, packed row groups): packing is a launch-geometry hint only
        enc_index = enc_index[0]
    if enc_index is not None:   # shared-K/V extension: same result as feeding the gathered encoder states
        enc = enc.index_select(0, enc_index.long())
    sd = _sd(ATT_NAMES, params[:10], "p.attention")
    if cfg.has_cross:
        sd.update(_sd(ATT_NAMES, params[10:20], "p.crossattention"))
    sd.update(_sd(FFN_NAMES, params[-6:], "p"))
    B, L, _ = x.shape
    Lk = L + (past_kv[0].shape[2] if past_kv is not None else 0)
    ext = torch.zeros(B, 1, L, Lk, device=x.device)
    if key_mask is not None:
        ext = ext + key_mask[:, None, None, :]
    if cfg.causal:
        i = torch.arange(L, device=x.device)[:, None]
        j = torch.arange(Lk, device=x.device)[None, :]
        ext = ext + (j > i + (Lk - L)).float()[None, None] * -10000.0
    em = None if enc_mask is None else enc_mask[:, None, None, :]
    hz = (self_head_z, cross_head_z) if (cfg.has_cross and self_head_z is not None) else self_head_z
    out, sp, cp, kv = O.bert_layer(sd, "p", x, cfg.num_heads, ext, cfg.has_cross, 0, 0, enc, em, hz, mlp_z, past_kv, cfg.eps,
                                   cross_heads=cfg.cross_heads if cfg.has_cross else None)
    return out, (sp if cfg.want_probs else None), (cp if cfg.want_probs else None), kv


def mse_pairs(students, teachers, scales):
    return torch.stack([F.mse_loss(s, t.detach()) * w for s, t, w in zip(students, teachers, scales)])


def xent_rows(logits, labels, ignore_index=-100, label_smoothing=0.0):
    if label_smoothing > 0:
        return O.label_smooth_ce(logits, labels, label_smoothing, "none", ignore_index)
    return F.cross_entropy(logits, labels, reduction="none", ignore_index=ignore_index)


def kl_rows(s, t, inv_temp=1.0):
    ls, lt = F.log_softmax(s * inv_temp, -1), F.log_softmax(t.detach() * inv_temp, -1)
    return (lt.exp() * (lt - ls)).sum(-1)


def soft_xent_rows(logits, labels):
    return -(F.log_softmax(logits, -1) * labels).sum(-1)


def sum_scaled(x, scale=1.0):
    return x.sum() * scale


def l2_normalize(x):
    return F.normalize(x, dim=-1)


def sim_over_temp(a, b, temp):
    return a @ b.t() / temp


def l0_sample(loga, u, temperature):
    return O.l0_sample_z(loga, u, temperature)


def l0_expected_size(logas, weights, temperature):
    return sum(torch.sum(1 - O.cdf_qz(0, la, temperature)) * w for la, w in zip(logas, weights))


def k_l0_deterministic(loga, temperature, magical_number):
    rows = [O.l0_deterministic_z(loga.shape[1], loga[l], temperature, magical_number) for l in range(loga.shape[0])]
    m = torch.stack(rows)
    return m, m.sum(1).to(torch.int32)


def k_itm_sample_neg(sim, idx, u):
    w = F.softmax(sim, 1) + 1e-5
    B = sim.shape[0]
    if idx is None:
        w = w.masked_fill(torch.eye(B, dtype=torch.bool, device=sim.device), 0)
    else:
        w = w.masked_fill(idx.view(-1, 1) == idx.view(1, -1), 0)
    c = torch.cumsum(w, 1)
    return (c > (u * c[:, -1]).unsqueeze(1)).float().argmax(1)


# ---- optimizer kernels (csrc/l0_optim.cu) — what transformers 4.12.5's AdamW + torch.nn.utils.clip_grad_norm_ do (reference optim.py:67,
# accelerators/apex_ddp_accelerator.py:96-101), on the flat arenas
def k_store_f32(dst, values):
    dst[:len(values)] = torch.tensor(values, dtype=torch.float32, device=dst.device)


def k_sumsq(x, out):
    out += x.double().pow(2).sum().float()


def k_clip_coef(sumsq_t, max_norm, coef):
    coef.copy_(torch.clamp(max_norm / (sumsq_t.sqrt() + 1e-6), max=1.0))


def k_adamw_step(groups, grad_scale=None, hyper_dev=None):
    for i, gr in enumerate(groups):
        g = gr["g"] * grad_scale if grad_scale is not None else gr["g"]
        b1, b2, t = gr["beta1"], gr["beta2"], gr["step"]
        if hyper_dev is not None:
            step_size, decay = hyper_dev[2 * i], hyper_dev[2 * i + 1]
        else:
            step_size, decay = gr["lr"] * (1.0 - b2 ** t) ** 0.5 / (1.0 - b1 ** t), gr["lr"] * gr["weight_decay"]
        gr["m"].mul_(b1).add_(g, alpha=1.0 - b1)
        gr["v"].mul_(b2).addcmul_(g, g, value=1.0 - b2)
        gr["p"].sub_(step_size * gr["m"] / (gr["v"].sqrt() + gr["eps"]))
        if gr["weight_decay"] != 0.0:
            gr["p"].sub_(decay * gr["p"])


def install_optimizer(monkeypatch):
    """Torch versions of the FlatAdamW kernels: lets CPU tests run whole training loops (tests/test_dropin.py, gd_train_loop)."""
    import efficientvlm_b200.kernels as K
    for name in ("store_f32", "sumsq", "clip_coef", "adamw_step"):
        monkeypatch.setattr(K, name, globals()["k_" + name])


def install(monkeypatch):
    import efficientvlm_b200.kernels as K
    import efficientvlm_b200.ops as ops
    for name in ("vit_layer", "vit_embed", "layer_norm", "linear", "gelu", "bert_embed", "bert_layer", "mse_pairs", "xent_rows", "kl_rows",
                 "soft_xent_rows", "sum_scaled", "l2_normalize", "sim_over_temp", "l0_sample", "l0_expected_size"):
        monkeypatch.setattr(ops, name, globals()[name])
    monkeypatch.setattr(K, "l0_deterministic", k_l0_deterministic)
    monkeypatch.setattr(K, "itm_sample_neg", k_itm_sample_neg)
