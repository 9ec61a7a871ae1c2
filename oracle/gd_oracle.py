"""ORACLE — TEST INFRASTRUCTURE ONLY.  CPU fp32 re-statement of one general-distillation step
(models/model_pretrain.py:11-82 forward for teacher and student + GeneralDistill.py:300-376 loss mix), composed from
oracle/xvlm_oracle.py.  Parity status: PINNED — tests/test_oracle_golden.py::test_gd_oracle checks the whole step (every loss term,
total, 13 gradients) against tests/golden/gd_kd_tiny.pt, produced by oracle/make_golden_gd.py from the UNMODIFIED
`models/model_pretrain.py::XVLM` and the reference's own train-loop statements (`GeneralDistill.py:300-376`, lifted with `ast`).  Used (a) by tests to check the CUDA product's full step (loss and gradients) and (b) by bench.py as the
reported CPU baseline / `--impl reference` arm ("port" kind: the reference's own modules need /root/reference, which does not
exist on the GPU box).
"""
import torch
import torch.nn.functional as F

from . import xvlm_oracle as O


def bbox_losses(output_coord, target_bbox, is_image=None):
    """models/xvlm.py:587-612 (L1 + GIoU of row-aligned boxes; models/box_ops.py:9-57 restated for the diagonal only)."""
    loss_bbox = (output_coord - target_bbox).abs()

    def xyxy(b):                                                     # box_ops.py:9-13
        cx, cy, w, h = b.unbind(-1)
        return torch.stack([cx - 0.5 * w, cy - 0.5 * h, cx + 0.5 * w, cy + 0.5 * h], -1)

    b1, b2 = xyxy(output_coord), xyxy(target_bbox)
    if bool((b1[:, 2:] < b1[:, :2]).any()) or bool((b2[:, 2:] < b2[:, :2]).any()):          # xvlm.py:598-601
        loss_giou = torch.zeros(output_coord.size(0), device=output_coord.device)
    else:
        area1 = (b1[:, 2] - b1[:, 0]) * (b1[:, 3] - b1[:, 1])
        area2 = (b2[:, 2] - b2[:, 0]) * (b2[:, 3] - b2[:, 1])
        wh = (torch.min(b1[:, 2:], b2[:, 2:]) - torch.max(b1[:, :2], b2[:, :2])).clamp(min=0)   # box_ops.py:28-32
        inter = wh[:, 0] * wh[:, 1]
        union = area1 + area2 - inter
        hull = (torch.max(b1[:, 2:], b2[:, 2:]) - torch.min(b1[:, :2], b2[:, :2])).clamp(min=0)  # box_ops.py:51-55
        hull_area = hull[:, 0] * hull[:, 1]
        loss_giou = 1 - (inter / union - (hull_area - union) / hull_area)
    if is_image is None:
        num_boxes = target_bbox.size(0)
    else:                                                            # xvlm.py:607-610: whole-image rows carry no box
        num_boxes = torch.sum(1 - is_image)
        loss_bbox = loss_bbox * (1 - is_image.view(-1, 1))
        loss_giou = loss_giou * (1 - is_image)
    return loss_bbox.sum() / num_boxes, loss_giou.sum() / num_boxes


def pretrain_forward(sd, cfg, image, text_ids, text_atts, text_ids_masked, masked_pos, masked_ids, neg_img=None, neg_txt=None,
                     region=None):
    """models/model_pretrain.py:11-82 with KD outputs; `sd` is the model's state_dict (reference key names),
    cfg = dict(vit_layers, vit_heads, text_layers, text_heads[, local_attn_depth]); ITM negatives are injected (quirk Q4).
    `region` = dict(idx_to_group_img, image_atts, target_bbox, is_image) selects the `ret_bbox_loss=True` branch (:15-17, :62-74):
    `image` then holds the distinct images, every other input has one row per region."""
    nl, nh = cfg["text_layers"], cfg["text_heads"]
    fl = nl // 2
    if region is None:
        img, img_hidden, img_att = O.vit_forward(sd, "vision_encoder", image, cfg["vit_heads"], cfg["vit_layers"])
        image_atts = torch.ones(img.shape[:2], device=img.device)
    else:                                                            # xvlm.py:331-364 with image_atts and idx_to_group_img
        idx = region["idx_to_group_img"]
        image_atts = region["image_atts"]
        img, img_hidden, img_att, img_full = O.vit_forward(sd, "vision_encoder", image, cfg["vit_heads"], cfg["vit_layers"],
                                                           idx_to_group_img=idx, image_atts=image_atts,
                                                           local_attn_depth=cfg.get("local_attn_depth", 0))
        img_full = torch.gather(img_full, 0, idx.view(-1, 1, 1).expand(-1, img_full.shape[1], img_full.shape[2]))
    B = text_ids.shape[0]
    te = O.bert_model(sd, "text_encoder.bert", nh, nl, fl, text_ids, text_atts, mode="text")
    text_embeds = te["last"]
    temp = sd["temp"]
    image_feat, text_feat = O.get_features(sd, img, text_embeds)
    loss_itc = O.contrastive_loss(image_feat, text_feat, temp)
    if neg_img is None:      # xvlm.py:439-455 with a deterministic draw (the fixtures patch torch.multinomial to argmax)
        w_i2t, w_t2i = O.itm_negative_weights(image_feat.detach(), text_feat.detach(), temp.detach(), None)
        neg_img, neg_txt = w_t2i.argmax(1), w_i2t.argmax(1)
    ie_all, ia_all, te_all, ta_all = O.itm_batches(img, image_atts, text_embeds, text_atts, neg_img, neg_txt)
    pos = O.bert_model(sd, "text_encoder.bert", nh, nl, fl, attention_mask=text_atts, encoder_embeds=text_embeds,
                       encoder_hidden_states=img, encoder_attention_mask=image_atts, mode="fusion")
    neg = O.bert_model(sd, "text_encoder.bert", nh, nl, fl, attention_mask=ta_all, encoder_embeds=te_all, encoder_hidden_states=ie_all,
                       encoder_attention_mask=ia_all, mode="fusion")
    itm_logits = O.build_mlp_forward(sd, "itm_head", torch.cat([pos["last"][:, 0], neg["last"][:, 0]], 0))
    itm_labels = torch.cat([torch.ones(B, dtype=torch.long, device=img.device), torch.zeros(2 * B, dtype=torch.long, device=img.device)])
    loss_itm = F.cross_entropy(itm_logits, itm_labels)
    loss_mlm, mlm_logits, mlm = O.masked_lm_forward(sd, "text_encoder", nh, nl, fl, text_ids_masked, text_atts, img, image_atts,
                                                    masked_pos, masked_ids)
    out = {
        "loss": {"loss_itc": loss_itc, "loss_itm": loss_itm, "loss_mlm": loss_mlm},
        "hidden_dict": {"image_hidden_states": img_hidden, "text_hidden_states": te["hidden"], "itm_pos_hidden_states": pos["hidden"],
                        "itm_neg_hidden_states": neg["hidden"], "mlm_hidden_states": mlm["hidden"]},
        "attention_dict": {"image_attentions": img_att, "text_attentions": te["attentions"], "itm_pos_attentions": pos["attentions"],
                           "itm_neg_attentions": neg["attentions"], "mlm_attentions": mlm["attentions"]},
        "logits_dict": {"itm_head_logits": itm_logits, "mlm_logits": mlm_logits},
        "feats": (image_feat, text_feat),
    }
    if region is not None:                                           # model_pretrain.py:62-74, xvlm.py:566-584
        box = O.bert_model(sd, "text_encoder.bert", nh, nl, fl, attention_mask=text_atts, encoder_embeds=text_embeds,
                           encoder_hidden_states=img_full, encoder_attention_mask=torch.ones(img_full.shape[:2], device=img.device),
                           mode="fusion")
        output_coord = O.build_mlp_forward(sd, "bbox_head", box["last"][:, 0]).sigmoid()
        out["loss"]["loss_bbox"], out["loss"]["loss_giou"] = bbox_losses(output_coord, region["target_bbox"], region.get("is_image"))
        out["hidden_dict"]["bbox_hidden_states"] = box["hidden"]
        out["attention_dict"]["bbox_attentions"] = box["attentions"]
        out["output_coord"] = output_coord
    return out


def gd_total_loss(so, to, temperature=1.0):
    """GeneralDistill.py:300-376."""
    def kd(name_h, name_a, is_img=False):
        sh, th = so["hidden_dict"][name_h], to["hidden_dict"][name_h]
        sa, ta = so["attention_dict"][name_a], to["attention_dict"][name_a]
        h = O.get_kd_loss(sh, O.get_cor_teacher(th, sh), is_img=is_img)
        a = O.get_kd_loss(sa, O.get_cor_teacher(ta, sa, is_attn=True), is_attn=True)
        return h, a

    text_h, text_a = kd("text_hidden_states", "text_attentions")
    img_h, img_a = kd("image_hidden_states", "image_attentions", is_img=True)
    pos_h, pos_a = kd("itm_pos_hidden_states", "itm_pos_attentions")
    neg_h, neg_a = kd("itm_neg_hidden_states", "itm_neg_attentions")
    mlm_h, mlm_a = kd("mlm_hidden_states", "mlm_attentions")
    mlm_kl = O.soft_cross_entropy(so["logits_dict"]["mlm_logits"] / temperature, to["logits_dict"]["mlm_logits"].detach() / temperature)
    itm_kl = O.soft_cross_entropy(so["logits_dict"]["itm_head_logits"] / temperature,
                                  to["logits_dict"]["itm_head_logits"].detach() / temperature)
    loss = so["loss"]
    loss_small = loss["loss_itc"] + loss["loss_itm"] + loss["loss_mlm"]
    if "loss_bbox" in loss:                                          # the region branch's mix, GeneralDistill.py:257
        loss_small = loss_small + loss["loss_bbox"] + loss["loss_giou"]
    loss_kd = itm_kl + mlm_kl + (text_a + text_h) + (img_a + 0.1 * img_h) + (neg_a + neg_h + pos_a + pos_h + mlm_a + mlm_h)
    return loss_small * 0.6 + loss_kd * 0.4, dict(loss_small=loss_small, loss_kd=loss_kd)


def gd_step(student_sd, teacher_sd, s_cfg, t_cfg, batch, negs_s, negs_t, temperature=1.0, region=None):
    """One oracle GD step: returns (total loss, components, student outputs). Gradients flow into the tensors of student_sd
    that require grad.  With `region` (see pretrain_forward) it is the region-batch half of the iteration (GeneralDistill.py:158-260)."""
    negs_s, negs_t = negs_s or (None, None), negs_t or (None, None)     # None: argmax of each model's own sampling weights
    with torch.no_grad():
        to = pretrain_forward(teacher_sd, t_cfg, *batch, *negs_t, region=region)
    so = pretrain_forward(student_sd, s_cfg, *batch, *negs_s, region=region)
    total, parts = gd_total_loss(so, to, temperature)
    return total, parts, so
