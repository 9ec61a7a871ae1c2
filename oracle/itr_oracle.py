"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU fp32 re-statement of one ITR-COCO pruning step (BASELINE config 4):
`efficient_models/model_retrieval.py:24-92` (L0-gated student with KD outputs), `models/model_retrieval.py:20-63` (un-gated
teacher, KD outputs without the ITC loss) and the loss assembly of `Eff_Retrieval.py:96-178`, composed from oracle/xvlm_oracle.py.

Parity status: PINNED — tests/test_oracle_golden.py::test_itr_oracle checks it against tests/golden/itr_kd_tiny.pt, produced by
oracle/make_golden_itr.py from the UNMODIFIED reference classes.  Only tests/, smoke() and bench.py's CPU legs may import this.
"""
import torch
import torch.nn.functional as F

from . import xvlm_oracle as O


def retrieval_forward(sd, cfg, image, text_ids, text_atts, idx=None, zs=None, negatives=None, with_losses=True):
    """KD-output forward.  cfg = dict(vit_layers, vit_heads, text_layers, text_heads); zs = the six gate tensors (student) or None
    (teacher); negatives = (neg_img_idx, neg_txt_idx) or None -> argmax of the sampling weights (the fixtures' deterministic draw)."""
    z = zs or {}
    nl, nh = cfg["text_layers"], cfg["text_heads"]
    fl = nl // 2
    img, img_hidden, img_att = O.vit_forward(sd, "vision_encoder", image, cfg["vit_heads"], cfg["vit_layers"], head_z=z.get("vision_head_z"),
                                             mlp_z=z.get("vision_intermediate_z"))
    B = image.shape[0]
    image_atts = torch.ones(img.shape[:2])
    te = O.bert_model(sd, "text_encoder", nh, nl, fl, text_ids, text_atts, mode="text", head_z=z.get("text_head_z"),
                      mlp_z=z.get("text_intermediate_z"))
    text_embeds = te["last"]
    image_feat, text_feat = O.get_features(sd, img, text_embeds)
    if negatives is None:
        w_i2t, w_t2i = O.itm_negative_weights(image_feat.detach(), text_feat.detach(), sd["temp"].detach(), idx)
        negatives = (w_t2i.argmax(1), w_i2t.argmax(1))                               # xvlm.py:439-455 with a deterministic draw
    ie_all, ia_all, te_all, ta_all = O.itm_batches(img, image_atts, text_embeds, text_atts, *negatives)
    gates = dict(head_z=z.get("cross_head_z"), mlp_z=z.get("cross_intermediate_z"))
    pos = O.bert_model(sd, "text_encoder", nh, nl, fl, attention_mask=text_atts, encoder_embeds=text_embeds, encoder_hidden_states=img,
                       encoder_attention_mask=image_atts, mode="fusion", **gates)
    neg = O.bert_model(sd, "text_encoder", nh, nl, fl, attention_mask=ta_all, encoder_embeds=te_all, encoder_hidden_states=ie_all,
                       encoder_attention_mask=ia_all, mode="fusion", **gates)
    itm_logits = O.build_mlp_forward(sd, "itm_head", torch.cat([pos["last"][:, 0], neg["last"][:, 0]], 0))
    out = {"hidden_dict": {"image_hidden_states": img_hidden, "text_hidden_states": te["hidden"], "itm_pos_hidden_states": pos["hidden"],
                           "itm_neg_hidden_states": neg["hidden"]},
           "attention_dict": {"image_attentions": img_att, "text_attentions": te["attentions"], "itm_pos_attentions": pos["attentions"],
                              "itm_neg_attentions": neg["attentions"]},
           "cross_attention_dict": {"itm_pos_cross_attentions": pos["cross_attentions"], "itm_neg_cross_attentions": neg["cross_attentions"]},
           "logits_dict": {"itm_head_logits": itm_logits}}
    if with_losses:
        itm_labels = torch.cat([torch.ones(B, dtype=torch.long), torch.zeros(2 * B, dtype=torch.long)])
        out["loss"] = {"loss_itc": O.contrastive_loss(image_feat, text_feat, sd["temp"], idx), "loss_itm": F.cross_entropy(itm_logits, itm_labels)}
    return out


def itr_total_loss(so, to, temperature=1.0):
    """Eff_Retrieval.py:112-178 without the Lagrangian term.  Returns (loss, dict of the eleven KD terms)."""
    sh, th, sa, ta = so["hidden_dict"], to["hidden_dict"], so["attention_dict"], to["attention_dict"]
    sc, tc = so["cross_attention_dict"], to["cross_attention_dict"]

    def hid(name, is_img=False):
        return O.get_kd_loss(sh[name], O.get_cor_teacher(th[name], sh[name]), is_img=is_img)

    def att(s, t):
        return O.get_kd_loss(s, O.get_cor_teacher(t, s, is_attn=True), is_attn=True)
    p = dict(text_hidden=hid("text_hidden_states"), text_attention=att(sa["text_attentions"], ta["text_attentions"]),
             image_hidden=hid("image_hidden_states", True), image_attention=att(sa["image_attentions"], ta["image_attentions"]),
             itm_pos_hidden=hid("itm_pos_hidden_states"), itm_pos_attn=att(sa["itm_pos_attentions"], ta["itm_pos_attentions"]),
             itm_pos_cross=att(sc["itm_pos_cross_attentions"], tc["itm_pos_cross_attentions"]),
             itm_neg_hidden=hid("itm_neg_hidden_states"), itm_neg_attn=att(sa["itm_neg_attentions"], ta["itm_neg_attentions"]),
             itm_neg_cross=att(sc["itm_neg_cross_attentions"], tc["itm_neg_cross_attentions"]),
             itm_logits=O.soft_cross_entropy(so["logits_dict"]["itm_head_logits"] / temperature,
                                             to["logits_dict"]["itm_head_logits"].detach() / temperature))
    loss_text_kd = p["text_hidden"] + p["text_attention"]
    loss_img_kd = 0.2 * p["image_hidden"] + p["image_attention"]
    loss_cross_kd = (p["itm_neg_hidden"] + p["itm_pos_hidden"] + p["itm_pos_attn"] + p["itm_pos_cross"] + p["itm_neg_attn"] + p["itm_neg_cross"]) * 0.5
    loss_kd = p["itm_logits"] + (loss_text_kd + loss_img_kd + loss_cross_kd) * 0.33
    loss_small = so["loss"]["loss_itc"] + so["loss"]["loss_itm"]
    return (loss_kd + loss_small) * 0.5, p


def itr_step(student_sd, teacher_sd, s_cfg, t_cfg, batch, zs, lagrangian=None, temperature=1.0):
    """batch = (image, text_ids, text_atts, idx).  Returns (loss, student outputs, teacher outputs)."""
    so = retrieval_forward(student_sd, s_cfg, *batch, zs=zs)
    with torch.no_grad():
        to = retrieval_forward(teacher_sd, t_cfg, *batch, zs=None, with_losses=False)
    loss, _ = itr_total_loss(so, to, temperature)
    if lagrangian is not None:
        loss = loss + lagrangian()
    return loss, so, to
