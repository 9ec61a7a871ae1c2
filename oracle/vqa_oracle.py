"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU fp32 re-statement of the VQA task model and its pruning/KD step:
`efficient_models/model_generation.py:23-300` (EffXVLMForVQA: gated train forward, eval forward, rank_answer, tile),
`models/model_generation.py:228-443` (XVLMForVQA teacher = the same with no gates) and the loss assembly of
`Eff_VQA.py:105-181`, composed from oracle/xvlm_oracle.py.

Parity status: PINNED — tests/test_oracle_golden.py::test_vqa_oracle checks every function here against
tests/golden/vqa_tiny.pt, which oracle/make_golden_vqa.py produced by running the UNMODIFIED reference classes.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this file.
"""
import torch
import torch.nn.functional as F

from . import xvlm_oracle as O


def tile(x, dim, n_tile):
    """model_generation.py:15-21."""
    init_dim = x.size(dim)
    repeat_idx = [1] * x.dim()
    repeat_idx[dim] = n_tile
    x = x.repeat(*repeat_idx)
    order = torch.cat([init_dim * torch.arange(n_tile) + i for i in range(init_dim)])
    return torch.index_select(x, dim, order)


def gates_from_zs(zs):
    """model_generation.py:104-111,128-129,150-151: the eight L0 gate tensors routed to the four sub-networks."""
    if zs is None:
        return dict(vision_head=None, vision_mlp=None, enc_head=None, enc_mlp=None, dec_head=None, dec_mlp=None)
    return dict(vision_head=zs["vision_head_z"], vision_mlp=zs["vision_intermediate_z"],
                enc_head=torch.cat((zs["text_head_z"], zs["cross_head_z"]), 0),
                enc_mlp=torch.cat((zs["text_intermediate_z"], zs["cross_intermediate_z"]), 0),
                dec_head=zs["decoder_head_z"], dec_mlp=zs["decoder_intermediate_z"])


def encode(sd, cfg, image, q_ids, q_atts, z):
    """vision encoder + question encoder (multi_modal mode, concatenated gates -> quirk Q1); model_generation.py:112-129."""
    img, img_hidden, img_att = O.vit_forward(sd, "vision_encoder", image, cfg["vit_heads"], cfg["vit_layers"], head_z=z["vision_head"],
                                             mlp_z=z["vision_mlp"])
    nl = cfg["text_layers"]
    q = O.bert_model(sd, "text_encoder", cfg["text_heads"], nl, nl // 2, q_ids, q_atts, encoder_hidden_states=img,
                     encoder_attention_mask=torch.ones(img.shape[:2]), head_z=z["enc_head"], mlp_z=z["enc_mlp"])
    return img_hidden, img_att, q


def train_forward(sd, cfg, image, q_ids, q_atts, a_ids, a_atts, k, weights, zs=None, pad_token_id=0):
    """model_generation.py:98-180 with KD outputs.  cfg = dict(vit_layers, vit_heads, text_layers, text_heads, dec_layers)."""
    z = gates_from_zs(zs)
    img_hidden, img_att, q = encode(sd, cfg, image, q_ids, q_atts, z)
    targets = a_ids.masked_fill(a_ids == pad_token_id, -100)
    states, atts = [], []
    for b, n in enumerate(k):                                                      # :134-139
        states += [q["last"][b]] * n
        atts += [q_atts[b]] * n
    states, atts = torch.stack(states, 0), torch.stack(atts, 0)
    loss_rows, logits, dec = O.lm_head_forward(sd, "text_decoder", cfg["text_heads"], cfg["dec_layers"], 0, a_ids, a_atts, states, atts,
                                               labels=targets, reduction="none", head_z=z["dec_head"], mlp_z=z["dec_mlp"])
    loss = (weights * loss_rows).sum() / image.size(0)                             # :169-170
    return {"loss": loss,
            "hidden_dict": {"image_hidden_states": img_hidden, "text_hidden_states": q["hidden"], "decoder_hidden_states": dec["hidden"]},
            "attention_dict": {"image_attentions": img_att, "text_attentions": q["attentions"], "decoder_attentions": dec["attentions"]},
            "cross_attention_dict": {"cross_attentions": q["cross_attentions"], "decoder_cross_attentions": dec["cross_attentions"]},
            "logits_dict": {"logits": logits}}


def rank_answer(sd, cfg, question_states, question_atts, answer_ids, answer_atts, k, z, pad_token_id=0):
    """model_generation.py:233-300."""
    num_ques = question_states.size(0)
    start_ids = answer_ids[0, 0].repeat(num_ques, 1)
    _, logits, _ = O.lm_head_forward(sd, "text_decoder", cfg["text_heads"], cfg["dec_layers"], 0, start_ids, None, question_states,
                                     question_atts, head_z=z["dec_head"], mlp_z=z["dec_mlp"])
    logits = logits[:, 0, :]
    prob_first = F.softmax(logits, dim=1).index_select(1, answer_ids[:, 1])
    topk_probs, topk_ids = prob_first.topk(k, dim=1)
    input_ids = torch.cat([answer_ids.index_select(0, t) for t in topk_ids], 0)
    input_atts = torch.cat([answer_atts.index_select(0, t) for t in topk_ids], 0)
    targets = input_ids.masked_fill(input_ids == pad_token_id, -100)
    loss_rows, _, _ = O.lm_head_forward(sd, "text_decoder", cfg["text_heads"], cfg["dec_layers"], 0, input_ids, input_atts,
                                        tile(question_states, 0, k), tile(question_atts, 0, k), labels=targets, reduction="none",
                                        head_z=z["dec_head"], mlp_z=z["dec_mlp"])
    log_probs = torch.cat([topk_probs.view(-1, 1).log(), -loss_rows.view(input_ids.size(0), -1)], 1)
    probs = F.softmax(log_probs.sum(1).view(num_ques, k), dim=-1)
    probs, rerank = probs.topk(k, dim=1)
    return torch.gather(topk_ids, 1, rerank), probs


def eval_forward(sd, cfg, image, q_ids, q_atts, list_ids, list_atts, k, zs=None):
    """model_generation.py:188-212 (zs = deterministic masks of l0_module.forward(training=False), or None for the teacher)."""
    z = gates_from_zs(zs)
    _, _, q = encode(sd, cfg, image, q_ids, q_atts, z)
    return rank_answer(sd, cfg, q["last"], q_atts, list_ids, list_atts, k, z)


def kd_total_loss(so, to, temperature=1.0):
    """Eff_VQA.py:105-176.  Returns (loss without the Lagrangian term, dict of the eleven KD terms)."""
    sh, th, sa, ta = so["hidden_dict"], to["hidden_dict"], so["attention_dict"], to["attention_dict"]
    sc, tc = so["cross_attention_dict"], to["cross_attention_dict"]
    s_text_h = sh["text_hidden_states"]
    t_text_h = O.get_cor_teacher(th["text_hidden_states"], s_text_h)
    s_text_a = sa["text_attentions"]
    t_text_a = O.get_cor_teacher(ta["text_attentions"], s_text_a, is_attn=True)
    s_cross_a = sc["cross_attentions"]
    t_cross_a = O.get_cor_teacher(tc["cross_attentions"], s_cross_a, is_attn=True)
    p = {}
    p["text_hidden"] = O.get_kd_loss(s_text_h[:4], t_text_h[:4])
    p["text_attention"] = O.get_kd_loss(s_text_a[:3], t_text_a[:3], is_attn=True)
    p["cross_hidden"] = O.get_kd_loss(s_text_h[4:], t_text_h[4:])
    p["cross_self_attention"] = O.get_kd_loss(s_text_a[3:], t_text_a[3:], is_attn=True)
    p["cross_attention"] = O.get_kd_loss(s_cross_a, t_cross_a, is_attn=True)
    p["image_hidden"] = O.get_kd_loss(sh["image_hidden_states"], O.get_cor_teacher(th["image_hidden_states"], sh["image_hidden_states"]),
                                      is_img=True)
    p["image_attention"] = O.get_kd_loss(sa["image_attentions"], O.get_cor_teacher(ta["image_attentions"], sa["image_attentions"], True),
                                         is_attn=True)
    p["decoder_hidden"] = O.get_kd_loss(sh["decoder_hidden_states"],
                                        O.get_cor_teacher(th["decoder_hidden_states"], sh["decoder_hidden_states"]), is_img=True)
    p["decoder_attention"] = O.get_kd_loss(sa["decoder_attentions"],
                                           O.get_cor_teacher(ta["decoder_attentions"], sa["decoder_attentions"], True), is_attn=True)
    p["decoder_cross"] = O.get_kd_loss(sc["decoder_cross_attentions"],
                                       O.get_cor_teacher(tc["decoder_cross_attentions"], sc["decoder_cross_attentions"], True), is_attn=True)
    p["logits"] = O.soft_cross_entropy(so["logits_dict"]["logits"] / temperature, to["logits_dict"]["logits"] / temperature)
    loss_text_kd = p["text_attention"] + p["text_hidden"]
    loss_img_kd = p["image_attention"] + p["image_hidden"] * 0.2
    loss_cross_kd = (p["cross_hidden"] + p["cross_self_attention"] + p["cross_attention"]) * 0.5
    loss_decoder_kd = p["decoder_attention"] + p["decoder_hidden"] + p["decoder_cross"]
    loss_kd = p["logits"] + loss_text_kd + loss_img_kd + loss_cross_kd + loss_decoder_kd
    return loss_kd * 0.4 + so["loss"] * 0.6, p


def vqa_step(student_sd, teacher_sd, s_cfg, t_cfg, batch, zs, lagrangian=None, temperature=1.0):
    """One Eff_VQA.py training step's loss: batch = (image, q_ids, q_atts, a_ids, a_atts, k, weights); zs = the sampled gates;
    lagrangian = callable returning the Lagrangian term (or None).  Returns (loss, student outputs, teacher outputs)."""
    so = train_forward(student_sd, s_cfg, *batch, zs=zs)
    with torch.no_grad():
        to = train_forward(teacher_sd, t_cfg, *batch, zs=None)
    loss, _ = kd_total_loss(so, to, temperature)
    if lagrangian is not None:
        loss = loss + lagrangian()
    return loss, so, to


def l0_layout(hidden, intermediate, heads, vision_layers, text_layers):
    """generation_l0_module.py:41-62,117-168: gate types in registration order with (layers, size, params per dim, shape) and
    the prunable model size.  text_layers = text + cross (6 or 12); the decoder has as many layers as the cross part."""
    n_text = text_layers // 2
    n_cross = text_layers - n_text
    n_dec = n_cross
    per_head_layer = hidden * hidden * 4 + hidden * 4
    per_head = per_head_layer // heads
    per_mlp_layer = hidden * intermediate * 2 + hidden + hidden * 4
    per_int = per_mlp_layer // intermediate
    types, prunable = {}, 0
    for name, rows in (("vision_head", vision_layers), ("text_head", n_text), ("cross_head", n_cross * 2), ("decoder_head", n_dec * 2)):
        types[name] = dict(rows=rows, size=heads, per_dim=per_head, shape=[rows, 1, heads, 1, 1])
        prunable += per_head * rows * heads
    for name, rows in (("vision_intermediate", vision_layers), ("text_intermediate", n_text), ("cross_intermediate", n_cross),
                       ("decoder_intermediate", n_dec)):
        types[name] = dict(rows=rows, size=intermediate, per_dim=per_int, shape=[rows, 1, 1, intermediate])
        prunable += per_mlp_layer * rows
    return types, prunable


def sample_gates(layout, logas, eps):
    """l0_module.forward(training=True), generation_l0_module.py:347-354: one hard-concrete sample per type, reshaped."""
    return {t + "_z": O.l0_sample_z(logas[t], eps[t]).reshape(layout[t]["shape"]) for t in layout}


def deterministic_gates(layout, logas):
    """l0_module.forward(training=False), generation_l0_module.py:355-366: per-layer deterministic masks."""
    out = {}
    for t, spec in layout.items():
        rows = [O.l0_deterministic_z(spec["size"], logas[t][l].detach()) for l in range(spec["rows"])]
        out[t + "_z"] = torch.stack(rows).reshape(spec["shape"])
    return out
