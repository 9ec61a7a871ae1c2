"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU fp32 re-statement of one NLVR2 pruning step: `efficient_models/model_nlvr.py:190-245`
(L0-gated student, two images per text, fusion layers alternating between the images), `models/model_nlvr.py:187-228` (un-gated
teacher) and the loss assembly of `Eff_NLVR.py:100-157`, composed from oracle/xvlm_oracle.py.

Parity status: PINNED — tests/test_oracle_golden.py::test_nlvr_oracle checks it against tests/golden/nlvr_kd_tiny.pt, produced by
oracle/make_golden_nlvr.py from the reference classes (teacher constructed through the efficient_models text-encoder builder, see
quirk Q12 there).  Only tests/, smoke() and bench.py's CPU legs may import this.
"""
import torch
import torch.nn.functional as F

from . import xvlm_oracle as O


def nlvr_forward(sd, cfg, image, text_ids, text_atts, targets, zs=None):
    """cfg = dict(vit_layers, vit_heads, text_layers (nominal: 6 / 12), text_heads).  The encoder has text + 2 * cross layers; the
    state dict lists the tied cross-attention key / value tensors under both layer names (share_cross_attention)."""
    z = zs or {}
    n_text = cfg["text_layers"] // 2
    n_layers = n_text + 2 * (cfg["text_layers"] - n_text)
    img, img_hidden, img_att = O.vit_forward(sd, "vision_encoder", image, cfg["vit_heads"], cfg["vit_layers"], head_z=z.get("vision_head_z"),
                                             mlp_z=z.get("vision_intermediate_z"))
    B = targets.size(0)
    img0, img1 = torch.split(img, B)
    atts = [torch.ones(img0.shape[:2]), torch.ones(img1.shape[:2])]
    head_z = mlp_z = None
    if zs is not None:
        head_z = torch.cat((z["text_head_z"], z["cross_head_z"]), 0)
        mlp_z = torch.cat((z["text_intermediate_z"], z["cross_intermediate_z"]), 0)
    o = O.bert_model(sd, "text_encoder", cfg["text_heads"], n_layers, n_text, text_ids, text_atts, encoder_hidden_states=[img0, img1],
                     encoder_attention_mask=atts, head_z=head_z, mlp_z=mlp_z)
    pred = O.build_mlp_forward(sd, "cls_head", o["last"][:, 0, :])
    return {"loss": F.cross_entropy(pred, targets),
            "hidden_dict": {"image_hidden_states": img_hidden, "text_hidden_states": o["hidden"]},
            "attention_dict": {"image_attentions": img_att, "text_attentions": o["attentions"]},
            "cross_attention_dict": {"cross_attentions": o["cross_attentions"]}, "logits_dict": {"cls_head_logits": pred}}


def nlvr_total_loss(so, to, temperature=1.0):
    """Eff_NLVR.py:110-155 without the Lagrangian term.  Returns (loss, dict of the eight KD terms)."""
    sh, th, sa, ta = so["hidden_dict"], to["hidden_dict"], so["attention_dict"], to["attention_dict"]
    s_text_h = sh["text_hidden_states"]
    t_text_h = O.get_cor_teacher(th["text_hidden_states"], s_text_h)
    s_text_a = sa["text_attentions"]
    t_text_a = O.get_cor_teacher(ta["text_attentions"], s_text_a, is_attn=True)
    s_cross_a = so["cross_attention_dict"]["cross_attentions"]
    t_cross_a = O.get_cor_teacher(to["cross_attention_dict"]["cross_attentions"], s_cross_a, is_attn=True)
    p = dict(text_hidden=O.get_kd_loss(s_text_h[:4], t_text_h[:4]), text_attention=O.get_kd_loss(s_text_a[:3], t_text_a[:3], is_attn=True),
             cross_hidden=O.get_kd_loss(s_text_h[4:], t_text_h[4:]), cross_self_attention=O.get_kd_loss(s_text_a[3:], t_text_a[3:], is_attn=True),
             cross_attention=O.get_kd_loss(s_cross_a, t_cross_a, is_attn=True),
             image_hidden=O.get_kd_loss(sh["image_hidden_states"], O.get_cor_teacher(th["image_hidden_states"], sh["image_hidden_states"]), is_img=True),
             image_attention=O.get_kd_loss(sa["image_attentions"], O.get_cor_teacher(ta["image_attentions"], sa["image_attentions"], True), is_attn=True),
             logits=O.soft_cross_entropy(so["logits_dict"]["cls_head_logits"] / temperature, to["logits_dict"]["cls_head_logits"].detach() / temperature))
    loss_text_kd = p["text_attention"] + p["text_hidden"]
    loss_img_kd = p["image_attention"] + p["image_hidden"] * 0.1
    loss_cross_kd = (p["cross_hidden"] + p["cross_self_attention"] + p["cross_attention"]) * 0.5
    loss_kd = p["logits"] + loss_text_kd + (loss_img_kd + loss_cross_kd) * 0.33
    return 0.8 * so["loss"] + 0.2 * loss_kd, p
