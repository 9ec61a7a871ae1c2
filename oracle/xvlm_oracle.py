"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product package).

CPU fp32 re-statement, in plain torch tensor ops, of the EfficientVLM hot path:
  CLIP-ViT encoder (eff_vit.py), BERT text/fusion encoder + MLM / LM heads (eff_bert.py),
  X-VLM feature / ITC / ITM / MLM glue (xvlm.py), hard-concrete L0 gates (xvlm_l0_module.py),
  KD losses (GeneralDistill.py).
All file:line citations are relative to /root/reference.  Every function is *functional*: parameters come
from a flat dict keyed by the reference's own state_dict names (`sd`) plus a `prefix`, so the same
weights can be pushed through the reference, the oracle and the CUDA product.

Parity status: PINNED.  tests/test_oracle_golden.py checks every function here against fixtures in
tests/golden/ that were produced by running the unmodified reference modules (oracle/make_golden.py,
through oracle/ref_shim.py) on seeded inputs.
"""
import math

import torch
import torch.nn.functional as F

LIMIT_A, LIMIT_B, EPSILON = -0.1, 1.1, 1e-6  # xvlm_l0_module.py:16


# ----------------------------------------------------------------------------------------------
# activations (transformers 4.12.5 ACT2FN; un-vendored third-party arithmetic, published formulas)
# ----------------------------------------------------------------------------------------------
def quick_gelu(x):  # ACT2FN['quick_gelu'], call site eff_vit.py:210,218
    return x * torch.sigmoid(1.702 * x)


def gelu(x):  # ACT2FN['gelu'] (erf form), call sites eff_bert.py:441,717
    return 0.5 * x * (1.0 + torch.erf(x / math.sqrt(2.0)))


def layer_norm(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def linear(x, sd, name):
    return F.linear(x, sd[name + ".weight"], sd.get(name + ".bias"))


# ----------------------------------------------------------------------------------------------
# CLIP ViT — eff_vit.py:82-474
# ----------------------------------------------------------------------------------------------
def vit_attention(sd, p, h, num_heads, attention_mask=None, head_z=None, head_layer_z=None):
    """eff_vit.py:123-204. Returns (out, probs[B,h,N,N])."""
    B, N, E = h.shape
    d = sd[p + ".q_proj.weight"].shape[0] // num_heads
    scale = d ** -0.5
    q = linear(h, sd, p + ".q_proj") * scale                                   # :137
    k = linear(h, sd, p + ".k_proj")
    v = linear(h, sd, p + ".v_proj")
    sh = lambda t: t.view(B, N, num_heads, d).transpose(1, 2)                  # :102-103
    q, k, v = sh(q), sh(k), sh(v)
    w = q @ k.transpose(-1, -2)                                                # :147
    if attention_mask is not None:                                             # :163-169
        w = w + attention_mask
    probs = torch.softmax(w, dim=-1)                                           # :171
    ctx = probs @ v                                                            # :185 (dropout p=0.0)
    if head_z is not None:
        ctx = ctx * head_z                                                     # :194-195
    ctx = ctx.transpose(1, 2).reshape(B, N, num_heads * d)
    out = linear(ctx, sd, p + ".out_proj")                                     # :199
    if head_layer_z is not None:
        out = out * head_layer_z                                               # :201-202
    return out, probs


def vit_layer(sd, p, h, num_heads, attention_mask=None, head_z=None, head_layer_z=None, mlp_z=None, eps=1e-5):
    """eff_vit.py:231-273 (pre-LN)."""
    a = layer_norm(h, sd[p + ".layer_norm1.weight"], sd[p + ".layer_norm1.bias"], eps)
    o, probs = vit_attention(sd, p + ".self_attn", a, num_heads, attention_mask, head_z, head_layer_z)
    h = h + o
    m = layer_norm(h, sd[p + ".layer_norm2.weight"], sd[p + ".layer_norm2.bias"], eps)
    u = linear(m, sd, p + ".mlp.fc1")                                          # :215
    if mlp_z is not None:
        u = u * mlp_z                                                          # :216-217 gate BEFORE activation
    u = quick_gelu(u)
    h = h + linear(u, sd, p + ".mlp.fc2")                                      # :219,266
    return h, probs


def vit_forward(sd, p, x, num_heads, num_layers, patch=16, head_z=None, head_layer_z=None, mlp_z=None,
                idx_to_group_img=None, image_atts=None, local_attn_depth=0):
    """CLIPVisionTransformer.forward eff_vit.py:432-474 + CLIPEncoder.forward :291-383.
    Returns (out, hidden_states tuple (nl+1), attentions tuple (nl)[, out_fullatts])."""
    pref = (p + ".") if p else ""
    B = x.shape[0]
    pe = F.conv2d(x, sd[pref + "patch_embed.weight"], stride=patch).flatten(2).transpose(1, 2)     # :445-446
    cls = sd[pref + "class_embedding"].expand(B, 1, -1)
    h = torch.cat([cls, pe], 1) + sd[pref + "pos_embed.weight"][None]                               # :448-450
    h = layer_norm(h, sd[pref + "pre_layrnorm.weight"], sd[pref + "pre_layrnorm.bias"], 1e-5)       # :452
    do_gather = idx_to_group_img is not None
    blk_mask = None
    if do_gather and image_atts is not None:                                                        # :334-341
        full = torch.ones(h.shape[:2], dtype=h.dtype, device=h.device)
        m = torch.cat([image_atts.to(h.dtype), full], 0)[:, None, None, :]
        blk_mask = ((1.0 - m) * -10000.0).expand(-1, -1, m.size(-1), -1)
    hidden, atts = [], []
    for l in range(num_layers):
        hidden.append(h)                                                                            # :351-352 BEFORE the layer
        hz = head_z[l] if head_z is not None else None
        hlz = head_layer_z[l] if head_layer_z is not None else None
        mz = mlp_z[l] if mlp_z is not None else None
        lp = pref + "encoder.layers.%d" % l
        if local_attn_depth > 0 and l >= num_layers - local_attn_depth:                             # :353-367
            if do_gather:
                do_gather = False
                hb = torch.gather(h, 0, idx_to_group_img.view(-1, 1, 1).expand(-1, h.shape[1], h.shape[2]))
                h = torch.cat([hb, h], 0)
            h, pr = vit_layer(sd, lp, h, num_heads, blk_mask, hz, hlz, mz)
        else:
            h, pr = vit_layer(sd, lp, h, num_heads, None, hz, hlz, mz)
        atts.append(pr)
    hidden.append(h)                                                                                # :377-378
    out = layer_norm(h, sd[pref + "post_layernorm.weight"], sd[pref + "post_layernorm.bias"], 1e-5)  # :467
    if idx_to_group_img is not None:                                                                # :469-472
        bs = len(idx_to_group_img)
        return out[:bs], tuple(hidden), tuple(atts), out[bs:]
    return out, tuple(hidden), tuple(atts)


# ----------------------------------------------------------------------------------------------
# BERT — eff_bert.py:188-694, 953-1162
# ----------------------------------------------------------------------------------------------
def bert_embeddings(sd, p, input_ids, token_type_ids=None, position_ids=None, past_len=0, eps=1e-12, padding_idx=0):
    """eff_bert.py:188-215 (dropout omitted = eval mode).  padding_idx = config.pad_token_id (eff_bert.py:173): that word row
    receives no gradient from look-ups (None: plain indexing)."""
    B, L = input_ids.shape
    if position_ids is None:
        position_ids = torch.arange(past_len, past_len + L, device=input_ids.device)[None]
    if token_type_ids is None:
        token_type_ids = torch.zeros_like(input_ids)
    e = F.embedding(input_ids, sd[p + ".word_embeddings.weight"], padding_idx=padding_idx) + sd[p + ".token_type_embeddings.weight"][token_type_ids]
    e = e + sd[p + ".position_embeddings.weight"][position_ids]
    return layer_norm(e, sd[p + ".LayerNorm.weight"], sd[p + ".LayerNorm.bias"], eps)


def extended_attention_mask(attention_mask, is_decoder=False):
    """get_extended_attention_mask eff_bert.py:953-1013 -> additive [B,1,1|L,Lk]."""
    if attention_mask.dim() == 3:
        ext = attention_mask[:, None, :, :]
    elif is_decoder:
        B, L = attention_mask.shape[0], attention_mask.shape[1]
        ids = torch.arange(L, device=attention_mask.device)
        causal = (ids[None, None, :].repeat(B, L, 1) <= ids[None, :, None]).to(attention_mask.dtype)
        ext = causal[:, None, :, :] * attention_mask[:, None, None, :]
    else:
        ext = attention_mask[:, None, None, :]
    return (1.0 - ext.float()) * -10000.0


def invert_attention_mask(mask, neg=-10000.0):
    """transformers ModuleUtilsMixin.invert_attention_mask (call site eff_bert.py:1106-1111). The additive
    constant is -1e4 (fp16) / -1e9 (fp32) in 4.12.5 and finfo.min in 5.x: indistinguishable after softmax
    unless an entire row is masked."""
    m = mask[:, None, None, :] if mask.dim() == 2 else mask[:, None, :, :]
    return (1.0 - m.float()) * neg


def bert_self_attention(sd, p, h, num_heads, ext_mask, enc=None, enc_mask=None, head_z=None, past_kv=None, fp16_prescale=False):
    """BertSelfAttention.forward eff_bert.py:266-364 (eval: no dropout). Returns (ctx, probs, (k,v))."""
    B, L, _ = h.shape
    d = sd[p + ".query.weight"].shape[0] // num_heads
    q = linear(h, sd, p + ".query")
    src = enc if enc is not None else h
    k = linear(src, sd, p + ".key")
    v = linear(src, sd, p + ".value")
    sh = lambda t: t.view(t.shape[0], t.shape[1], num_heads, d).permute(0, 2, 1, 3)
    q, k, v = sh(q), sh(k), sh(v)
    mask = enc_mask if enc is not None else ext_mask
    if enc is None and past_kv is not None:                                   # :288-292
        k = torch.cat([past_kv[0], k], 2)
        v = torch.cat([past_kv[1], v], 2)
    if fp16_prescale:                                                         # :297-302 (quirk Q9)
        q = q / math.sqrt(d)
    s = q @ k.transpose(-1, -2)
    if not fp16_prescale:
        s = s / math.sqrt(d)                                                  # :330-331
    if mask is not None:
        s = s + mask                                                          # :333-335
    probs = torch.softmax(s, -1)                                              # :338
    ctx = probs @ v                                                           # :352
    if head_z is not None:
        ctx = ctx * head_z                                                    # :354-355
    ctx = ctx.permute(0, 2, 1, 3).reshape(B, L, num_heads * d)
    return ctx, probs, (k, v)


def bert_attention(sd, p, h, num_heads, ext_mask, enc=None, enc_mask=None, head_z=None, head_layer_z=None, past_kv=None,
                   eps=1e-12, fp16_prescale=False):
    """BertAttention eff_bert.py:409-433 + BertSelfOutput :374-381."""
    ctx, probs, kv = bert_self_attention(sd, p + ".self", h, num_heads, ext_mask, enc, enc_mask, head_z, past_kv, fp16_prescale)
    o = linear(ctx, sd, p + ".output.dense")
    if head_layer_z is not None:
        o = o * head_layer_z
    out = layer_norm(o + h, sd[p + ".output.LayerNorm.weight"], sd[p + ".output.LayerNorm.bias"], eps)
    return out, probs, kv


def bert_layer(sd, p, h, num_heads, ext_mask, has_cross, layer_num=0, fusion_layer=0, enc=None, enc_mask=None, head_z=None,
               mlp_z=None, past_kv=None, eps=1e-12, fp16_prescale=False, cross_heads=None):
    """BertLayer.forward eff_bert.py:480-560. Returns (h, self_probs, cross_probs|None, present_kv).  cross_heads: head count of
    the cross-attention block when prune_heads left it different from the self-attention's (eff_bert.py:391-407)."""
    cross_z = None
    if has_cross and head_z is not None:
        head_z, cross_z = head_z                                              # :494-496
    a, sp, kv = bert_attention(sd, p + ".attention", h, num_heads, ext_mask, head_z=head_z,
                               past_kv=past_kv[:2] if past_kv is not None else None, eps=eps, fp16_prescale=fp16_prescale)
    cp = None
    if has_cross:
        if isinstance(enc, list):                                             # :518-529 (NLVR)
            j = (layer_num - fusion_layer) % len(enc)
            e, em = enc[j], enc_mask[j]
        else:
            e, em = enc, enc_mask
        a, cp, _ = bert_attention(sd, p + ".crossattention", a, cross_heads or num_heads, ext_mask, e, em, head_z=cross_z, eps=eps,
                                  fp16_prescale=fp16_prescale)
    u = gelu(linear(a, sd, p + ".intermediate.dense"))                       # :445-448
    if mlp_z is not None:
        u = u * mlp_z                                                         # :555-556 gate AFTER activation
    o = linear(u, sd, p + ".output.dense")
    out = layer_norm(o + a, sd[p + ".output.LayerNorm.weight"], sd[p + ".output.LayerNorm.bias"], eps)   # :458-462
    return out, sp, cp, kv


def bert_encoder(sd, p, h, num_heads, num_layers, fusion_layer, ext_mask, enc=None, enc_mask=None, mode="multi_modal",
                 head_z=None, mlp_z=None, past_key_values=None, eps=1e-12, fp16_prescale=False):
    """BertEncoder.forward eff_bert.py:570-694. Returns dict(last, hidden, attentions, cross_attentions, cache)."""
    if mode == "text":
        start, end = 0, fusion_layer
    elif mode == "fusion":
        start, end = fusion_layer, num_layers
    elif mode == "multi_modal":
        start, end = 0, num_layers
    else:
        raise ValueError("mode %s is not supported" % mode)
    hidden, atts, catts, cache = [], [], [], []
    for i in range(start, end):
        hidden.append(h)
        if i >= fusion_layer and head_z is not None:                          # :612-615 (quirk Q1)
            first = (i - fusion_layer) * 2
            hz, mz = (head_z[first], head_z[first + 1]), mlp_z[i - fusion_layer]
        elif head_z is not None:
            hz, mz = head_z[i], mlp_z[i]
        else:
            hz, mz = None, None
        pkv = past_key_values[i] if past_key_values is not None else None
        h, sp, cp, kv = bert_layer(sd, p + ".layer.%d" % i, h, num_heads, ext_mask, i >= fusion_layer, i, fusion_layer, enc,
                                   enc_mask, hz, mz, pkv, eps, fp16_prescale)
        atts.append(sp)
        cache.append(kv)
        if cp is not None:
            catts.append(cp)
    hidden.append(h)
    return dict(last=h, hidden=tuple(hidden), attentions=tuple(atts), cross_attentions=tuple(catts), cache=tuple(cache))


def bert_model(sd, p, num_heads, num_layers, fusion_layer, input_ids=None, attention_mask=None, encoder_embeds=None,
               encoder_hidden_states=None, encoder_attention_mask=None, is_decoder=False, mode="multi_modal", head_z=None,
               mlp_z=None, past_key_values=None, fp16_prescale=False):
    """BertModel.forward eff_bert.py:1015-1162."""
    past_len = past_key_values[0][0].shape[2] if past_key_values is not None else 0
    if encoder_embeds is None:
        h = bert_embeddings(sd, p + ".embeddings", input_ids, past_len=past_len)
        B, L = input_ids.shape
    else:
        h = encoder_embeds
        B, L = h.shape[:2]
    if attention_mask is None:
        attention_mask = torch.ones(B, L + past_len, device=h.device)
    if is_decoder and attention_mask.dim() == 2 and past_len > 0:
        # :976-996 causal mask over the new positions with an all-ones prefix for the cached ones
        ids = torch.arange(L, device=h.device)
        causal = (ids[None, None, :].repeat(B, L, 1) <= ids[None, :, None]).to(attention_mask.dtype)
        causal = torch.cat([torch.ones(B, L, past_len, dtype=causal.dtype, device=h.device), causal], -1)
        ext = (1.0 - (causal[:, None] * attention_mask[:, None, None, :]).float()) * -10000.0
    else:
        ext = extended_attention_mask(attention_mask, is_decoder)
    enc_mask = None
    if encoder_hidden_states is not None:
        if isinstance(encoder_attention_mask, list):
            enc_mask = [invert_attention_mask(m) for m in encoder_attention_mask]
        elif encoder_attention_mask is None:
            enc_mask = invert_attention_mask(torch.ones(encoder_hidden_states.shape[:2], device=encoder_hidden_states.device))
        else:
            enc_mask = invert_attention_mask(encoder_attention_mask)
    return bert_encoder(sd, p + ".encoder", h, num_heads, num_layers, fusion_layer, ext, encoder_hidden_states, enc_mask, mode,
                        head_z, mlp_z, past_key_values, fp16_prescale=fp16_prescale)


def mlm_head(sd, p, x, eps=1e-12):
    """BertOnlyMLMHead / BertLMPredictionHead eff_bert.py:712-746 (decoder weight tied to word embeddings by the caller)."""
    t = gelu(linear(x, sd, p + ".predictions.transform.dense"))
    t = layer_norm(t, sd[p + ".predictions.transform.LayerNorm.weight"], sd[p + ".predictions.transform.LayerNorm.bias"], eps)
    return F.linear(t, sd[p + ".predictions.decoder.weight"], sd[p + ".predictions.bias"])


def masked_lm_forward(sd, p, num_heads, num_layers, fusion_layer, input_ids, attention_mask, enc, enc_atts, masked_pos, labels,
                      **kw):
    """BertForMaskedLM.forward eff_bert.py:1634-1714. Returns (loss, logits, encoder dict)."""
    o = bert_model(sd, p + ".bert", num_heads, num_layers, fusion_layer, input_ids, attention_mask, encoder_hidden_states=enc,
                   encoder_attention_mask=enc_atts, **kw)
    seq = o["last"]
    if masked_pos is not None:
        seq = torch.gather(seq, 1, masked_pos.unsqueeze(2).expand(-1, -1, seq.size(-1)))       # :1631-1632
    logits = mlm_head(sd, p + ".cls", seq)
    loss = None
    if labels is not None:
        loss = F.cross_entropy(logits.view(-1, logits.shape[-1]), labels.view(-1))            # :1701-1702 (ignore -100)
    return loss, logits, o


def label_smooth_ce(logits, label, lb_smooth=0.1, reduction="mean", ignore_index=-100):
    """LabelSmoothSoftmaxCEV1 eff_bert.py:1263-1302."""
    logits = logits.float()
    num_classes = logits.size(1)
    label = label.clone()
    ignore = label.eq(ignore_index)
    n_valid = ignore.eq(0).sum()
    label[ignore] = 0
    lb_pos, lb_neg = 1.0 - lb_smooth, lb_smooth / num_classes
    one_hot = torch.empty_like(logits).fill_(lb_neg).scatter_(1, label.unsqueeze(1), lb_pos)
    loss = -torch.sum(torch.log_softmax(logits, 1) * one_hot, dim=1)
    loss[ignore] = 0
    if reduction == "mean":
        loss = loss.sum() / n_valid
    if reduction == "sum":
        loss = loss.sum()
    return loss


def lm_head_forward(sd, p, num_heads, num_layers, fusion_layer, input_ids, attention_mask, enc, enc_atts, labels=None,
                    label_smoothing=0.0, reduction="mean", **kw):
    """BertLMHeadModel.forward eff_bert.py:1332-1443 (is_decoder=True). Returns (loss, logits, encoder dict)."""
    o = bert_model(sd, p + ".bert", num_heads, num_layers, fusion_layer, input_ids, attention_mask, encoder_hidden_states=enc,
                   encoder_attention_mask=enc_atts, is_decoder=True, **kw)
    logits = mlm_head(sd, p + ".cls", o["last"])
    loss = None
    if labels is not None:
        sl = logits[:, :-1, :].contiguous()
        lb = labels[:, 1:].contiguous()
        V = sl.shape[-1]
        if label_smoothing > 0:
            loss = label_smooth_ce(sl.view(-1, V), lb.view(-1), label_smoothing, reduction)
        else:
            loss = F.cross_entropy(sl.view(-1, V), lb.view(-1), reduction=reduction)
        if reduction == "none":
            loss = loss.view(logits.size(0), -1).sum(1)                                       # :1429-1430
    return loss, logits, o


# ----------------------------------------------------------------------------------------------
# X-VLM glue — xvlm.py:54-83, 375-518
# ----------------------------------------------------------------------------------------------
def get_features(sd, image_embeds=None, text_embeds=None):
    """xvlm.py:375-382."""
    out = []
    if image_embeds is not None:
        out.append(F.normalize(linear(image_embeds[:, 0, :], sd, "vision_proj"), dim=-1))
    if text_embeds is not None:
        out.append(F.normalize(linear(text_embeds[:, 0, :], sd, "text_proj"), dim=-1))
    return out[0] if len(out) == 1 else tuple(out)


def contrastive_loss(image_feat_all, text_feat_all, temp, idx_all=None):
    """get_contrastive_loss xvlm.py:384-416 on already-gathered features."""
    logits = image_feat_all @ text_feat_all.t() / temp
    n = logits.shape[0]
    if idx_all is None:
        labels = torch.arange(n, device=logits.device)
        return (F.cross_entropy(logits, labels) + F.cross_entropy(logits.t(), labels)) / 2
    idx_all = idx_all.view(-1, 1)
    pos = torch.eq(idx_all, idx_all.t()).float()
    labels = pos / pos.sum(1, keepdim=True)
    l_i2t = -torch.sum(F.log_softmax(logits, 1) * labels, 1).mean()
    l_t2i = -torch.sum(F.log_softmax(logits.t(), 1) * labels, 1).mean()
    return (l_i2t + l_t2i) / 2


def itm_negative_weights(image_feat, text_feat, temp, idx=None):
    """xvlm.py:422-437: sampling weights (i2t, t2i) for the hard negatives."""
    sim_i2t = image_feat @ text_feat.t() / temp
    sim_t2i = text_feat @ image_feat.t() / temp
    w_i2t = F.softmax(sim_i2t, 1) + 1e-5
    w_t2i = F.softmax(sim_t2i, 1) + 1e-5
    if idx is None:
        w_i2t.fill_diagonal_(0)
        w_t2i.fill_diagonal_(0)
    else:
        idx = idx.view(-1, 1)
        m = torch.eq(idx, idx.t())
        w_i2t.masked_fill_(m, 0)
        w_t2i.masked_fill_(m, 0)
    return w_i2t, w_t2i


def build_mlp_forward(sd, p, x):
    """build_mlp xvlm.py:77-83: Linear -> LayerNorm(1e-5) -> GELU -> Linear."""
    h = linear(x, sd, p + ".0")
    h = layer_norm(h, sd[p + ".1.weight"], sd[p + ".1.bias"], 1e-5)
    return linear(F.gelu(h), sd, p + ".3")


def itm_batches(image_embeds, image_atts, text_embeds, text_atts, neg_img_idx, neg_txt_idx):
    """xvlm.py:439-463: assemble the (2B) negative batch from sampled indices."""
    ie_neg, ia_neg = image_embeds[neg_img_idx], image_atts[neg_img_idx]
    te_neg, ta_neg = text_embeds[neg_txt_idx], text_atts[neg_txt_idx]
    return (torch.cat([ie_neg, image_embeds], 0), torch.cat([ia_neg, image_atts], 0),
            torch.cat([text_embeds, te_neg], 0), torch.cat([text_atts, ta_neg], 0))


# ----------------------------------------------------------------------------------------------
# L0 — xvlm_l0_module.py:174-271,321-341
# ----------------------------------------------------------------------------------------------
def cdf_qz(x, loga, temperature=2.0 / 3.0):
    xn = (x - LIMIT_A) / (LIMIT_B - LIMIT_A)
    logits = math.log(xn) - math.log(1 - xn)
    return torch.sigmoid(logits * temperature - loga).clamp(min=EPSILON, max=1 - EPSILON)


def l0_sample_z(loga, eps, temperature=2.0 / 3.0):
    y = torch.sigmoid((torch.log(eps) - torch.log(1 - eps) + loga) / temperature)
    return F.hardtanh(y * (LIMIT_B - LIMIT_A) + LIMIT_A, min_val=0, max_val=1)


def l0_deterministic_z(size, loga, temperature=2.0 / 3.0, magical_number=0.8):
    """_deterministic_z :253-271 for ONE layer row (python round = banker's rounding; CPU topk tie order)."""
    expected_num_nonzeros = torch.sum(1 - cdf_qz(0, loga, temperature))
    num_zeros = round(size - expected_num_nonzeros.item())
    soft = torch.sigmoid(loga / temperature * magical_number)
    if num_zeros > 0:
        _, ind = torch.topk(soft, k=num_zeros, largest=False)
        soft = torch.ones_like(soft)
        soft[ind] = 0.0
        return soft
    return torch.ones_like(soft)


def l0_expected_size(logas, params_per_dim, temperature=2.0 / 3.0):
    """get_num_parameters_and_constraint :198-216. logas / params_per_dim: dicts keyed by type."""
    n = 0
    for k, la in logas.items():
        n = n + torch.sum(1 - cdf_qz(0, la, temperature)) * params_per_dim[k]
    return n


def l0_lagrangian(logas, params_per_dim, prunable_model_size, lambda_1, lambda_2, target_sparsity, pruned_steps,
                  lagrangian_warmup, start_sparsity=0.0):
    """lagrangian_regularization :225-237."""
    expected_sparsity = 1 - l0_expected_size(logas, params_per_dim) / prunable_model_size
    t = target_sparsity
    if lagrangian_warmup > 0:
        t = (target_sparsity - start_sparsity) * min(1, pruned_steps / lagrangian_warmup) + start_sparsity
    loss = lambda_1 * (expected_sparsity - t) + lambda_2 * (expected_sparsity - t) ** 2
    return loss, expected_sparsity, t


# ----------------------------------------------------------------------------------------------
# KD — GeneralDistill.py:60-104
# ----------------------------------------------------------------------------------------------
def get_cor_teacher(teacher_reps, student_reps, is_attn=False):
    t = [r.detach() for r in teacher_reps]
    tn, sn = len(t), len(student_reps)
    if is_attn:
        assert tn % sn == 0
        k = tn // sn
        return [t[i * k + k - 1] for i in range(sn)]
    assert (tn - 1) % (sn - 1) == 0
    k = (tn - 1) // (sn - 1)
    return [t[i * k] for i in range(sn)]


def get_kd_loss(student_reps, teacher_reps, is_attn=False, is_img=False):
    loss = 0
    if is_attn:
        for s, t in zip(student_reps, teacher_reps):
            s = torch.where(s <= -1e2, torch.zeros_like(s), s)
            t = torch.where(t <= -1e2, torch.zeros_like(t), t)
            loss = loss + F.mse_loss(s, t) * s.shape[-1]
    elif is_img:
        for layer, (s, t) in enumerate(zip(student_reps, teacher_reps)):
            if layer != 6:
                loss = loss + F.mse_loss(s, t)
    else:
        for s, t in zip(student_reps, teacher_reps):
            loss = loss + F.mse_loss(s, t)
    return loss


def soft_cross_entropy(predicts, targets):
    sl = F.log_softmax(predicts, dim=-1)
    tp = F.softmax(targets, dim=-1)
    return F.kl_div(sl.view(-1, predicts.shape[-1]), tp.view(-1, targets.shape[-1]), reduction="batchmean")
