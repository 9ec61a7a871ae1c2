"""Pins the oracle (oracle/xvlm_oracle.py) against fixtures produced by the UNMODIFIED reference
(oracle/make_golden.py).  CPU only; tolerance 1e-5 relative (fp32 vs fp32, different op order), bit-exact masks."""
import torch

from oracle import xvlm_oracle as O
from tests.helpers import assert_close, load_golden, sd_from_spec

TOL = 2e-5


def test_vit_forward_and_grads():
    g = load_golden("vit_tiny")
    sd = {k: v.requires_grad_() if torch.is_floating_point(v) else v for k, v in sd_from_spec(g["sd_spec"]).items()}
    hz, mz = g["head_z"].clone().requires_grad_(), g["mlp_z"].clone().requires_grad_()
    out, hid, att = O.vit_forward(sd, "", g["x"], 2, 2, head_z=hz, mlp_z=mz)
    assert_close(out, g["out"], TOL, "vit out")
    for i, (a, b) in enumerate(zip(hid, g["hidden"])):
        assert_close(a, b, TOL, "vit hidden %d" % i)
    for i, (a, b) in enumerate(zip(att, g["attn"])):
        assert_close(a, b, TOL, "vit attn %d" % i)
    loss = out.pow(2).mean() + sum(a.pow(2).mean() for a in att) * 3.0
    assert_close(loss, g["loss"], TOL, "vit loss")
    wrt = [sd[n] for n in g["grad_names"][:-2]] + [hz, mz]
    grads = torch.autograd.grad(loss, wrt)
    for n, a, b in zip(g["grad_names"], grads, g["grads"]):
        assert_close(a, b, 1e-4, "vit grad " + n)
    out2, hid2, att2 = O.vit_forward(sd, "", g["x"], 2, 2)
    assert_close(out2, g["out_nogate"], TOL, "vit out nogate")
    assert_close(att2[1], g["attn_nogate"][1], TOL, "vit attn nogate")


def test_vit_region_batch():
    g = load_golden("vit_tiny")
    sd = sd_from_spec(g["sd_spec"])
    out, hid, att, full = O.vit_forward(sd, "", g["x"], 2, 2, idx_to_group_img=g["idx_to_group"], image_atts=g["image_atts"],
                                        local_attn_depth=1)
    assert_close(out, g["region_out"], TOL, "region out")
    assert_close(full, g["region_full"], TOL, "region fullatts")
    assert_close(att[1], g["region_attn"][1], TOL, "region attn")
    assert hid[-1].shape == g["region_hidden"][-1].shape


def _bert_cfg(g):
    c = g["cfg"]
    return c["num_attention_heads"], c["num_hidden_layers"], c["fusion_layer"]


def test_bert_modes_gates_and_grads():
    g = load_golden("bert_tiny")
    nh, nl, fl = _bert_cfg(g)
    sd = {"bert." + k: (v.requires_grad_() if torch.is_floating_point(v) else v) for k, v in sd_from_spec(g["sd_spec"]).items()}
    thz, tmz = g["text_head_z"].clone().requires_grad_(), g["text_mlp_z"].clone().requires_grad_()
    chz, cmz = g["cross_head_z"].clone().requires_grad_(), g["cross_mlp_z"].clone().requires_grad_()
    ot = O.bert_model(sd, "bert", nh, nl, fl, g["ids"], g["atts"], mode="text", head_z=thz, mlp_z=tmz)
    assert_close(ot["last"], g["text_last"], TOL, "text last")
    assert len(ot["hidden"]) == len(g["text_hidden"]) and len(ot["attentions"]) == len(g["text_attn"])
    for a, b in zip(ot["attentions"], g["text_attn"]):
        assert_close(a, b, TOL, "text attn")
    of = O.bert_model(sd, "bert", nh, nl, fl, attention_mask=g["atts"], encoder_embeds=ot["last"], encoder_hidden_states=g["img"],
                      encoder_attention_mask=g["img_atts"], mode="fusion", head_z=chz, mlp_z=cmz)
    assert_close(of["last"], g["fus_last"], TOL, "fusion last")
    for a, b in zip(of["hidden"], g["fus_hidden"]):
        assert_close(a, b, TOL, "fusion hidden")
    for a, b in zip(of["cross_attentions"], g["fus_cross"]):
        assert_close(a, b, TOL, "fusion cross attn")
    loss = of["last"][:, 0].pow(2).mean() + sum(a.pow(2).mean() for a in of["cross_attentions"]) \
        + sum(a.pow(2).mean() for a in ot["attentions"])
    assert_close(loss, g["loss"], TOL, "bert loss")
    wrt = [sd["bert." + n] for n in g["grad_names"][:-4]] + [thz, tmz, chz, cmz]
    grads = torch.autograd.grad(loss, wrt)
    for n, a, b in zip(g["grad_names"], grads, g["grads"]):
        assert_close(a, b, 1e-4, "bert grad " + n)


def test_bert_multimodal_quirk_q1_and_list_inputs():
    g = load_golden("bert_tiny")
    nh, nl, fl = _bert_cfg(g)
    sd = {"bert." + k: v for k, v in sd_from_spec(g["sd_spec"]).items()}
    hz = torch.cat([g["text_head_z"], g["cross_head_z"]])
    mz = torch.cat([g["text_mlp_z"], g["cross_mlp_z"]])
    o = O.bert_model(sd, "bert", nh, nl, fl, g["ids"], g["atts"], encoder_hidden_states=g["img"],
                     encoder_attention_mask=g["img_atts"], mode="multi_modal", head_z=hz, mlp_z=mz)
    assert_close(o["last"], g["mm_last"], TOL, "mm last")
    for a, b in zip(o["cross_attentions"], g["mm_cross"]):
        assert_close(a, b, TOL, "mm cross")
    o2 = O.bert_model(sd, "bert", nh, nl, fl, g["ids"], g["atts"], encoder_hidden_states=g["img"],
                      encoder_attention_mask=g["img_atts"], mode="multi_modal")
    assert_close(o2["last"], g["mm_nogate_last"], TOL, "mm nogate")
    o3 = O.bert_model(sd, "bert", nh, nl, fl, g["ids"], g["atts"], encoder_hidden_states=[g["img"], g["img2"]],
                      encoder_attention_mask=[g["img_atts"], g["img_atts"]], mode="multi_modal")
    assert_close(o3["last"], g["list_last"], TOL, "nlvr list")


def test_mlm_and_lm_heads():
    g = load_golden("heads_tiny")
    b = load_golden("bert_tiny")
    nh, nl, fl = _bert_cfg(b)
    sd = {"te." + k: (v.requires_grad_() if torch.is_floating_point(v) else v) for k, v in sd_from_spec(g["mlm_sd_spec"]).items()}
    sd["te.cls.predictions.decoder.weight"] = sd["te.bert.embeddings.word_embeddings.weight"]
    loss, logits, enc = O.masked_lm_forward(sd, "te", nh, nl, fl, b["ids"], b["atts"], b["img"], b["img_atts"], g["masked_pos"],
                                            g["labels"])
    assert_close(loss, g["mlm_loss"], TOL, "mlm loss")
    assert_close(logits, g["mlm_logits"], TOL, "mlm logits")
    grads = torch.autograd.grad(loss, [sd["te.bert.embeddings.word_embeddings.weight"], sd["te.cls.predictions.bias"],
                                       sd["te.cls.predictions.transform.dense.weight"]])
    for a, r in zip(grads, g["mlm_grads"]):
        assert_close(a, r, 1e-4, "mlm grad")
    sdd = {"td." + k: v for k, v in sd_from_spec(g["dec_sd_spec"]).items()}
    sdd["td.cls.predictions.decoder.weight"] = sdd["td.bert.embeddings.word_embeddings.weight"]
    l1, lg1, _ = O.lm_head_forward(sdd, "td", nh, nl, fl, b["ids"], b["atts"], b["img"], b["img_atts"], g["dlabels"], 0.1, "none",
                                   head_z=g["dec_head_z"], mlp_z=g["dec_mlp_z"])
    assert_close(l1, g["dec_loss_none_ls"], TOL, "decoder loss (label smoothing, none)")
    assert_close(lg1, g["dec_logits"], TOL, "decoder logits")
    l2, lg2, _ = O.lm_head_forward(sdd, "td", nh, nl, fl, b["ids"], b["atts"], b["img"], b["img_atts"], g["dlabels"], 0.0, "mean")
    assert_close(l2, g["dec_loss_mean"], TOL, "decoder loss mean")
    # KV-cache step == full recompute
    B = b["ids"].shape[0]
    _, _, e1 = O.lm_head_forward(sdd, "td", nh, nl, fl, b["ids"][:, :4], torch.ones(B, 4), b["img"], b["img_atts"])
    _, lg_step, _ = O.lm_head_forward(sdd, "td", nh, nl, fl, b["ids"][:, 4:5], torch.ones(B, 5), b["img"], b["img_atts"],
                                      past_key_values=e1["cache"])
    assert_close(lg_step, g["step2_logits"], TOL, "cached decode step")
    assert_close(lg_step[:, -1], g["full5_logits"][:, -1], 1e-4, "cache == full")


def test_l0_module():
    g = load_golden("l0_tiny")
    logas = {k: v.clone().requires_grad_() for k, v in g["logas"].items()}
    zs = {}
    for t in g["types"]:
        zs[t + "_z"] = O.l0_sample_z(logas[t], g["eps"][t]).reshape(g["shapes"][t])
    assert list(zs.keys()) == g["zs_order"]
    for k in zs:
        assert_close(zs[k], g["zs_train"][k], 1e-6, "z train " + k)
    for t in g["types"]:
        rows = [O.l0_deterministic_z(g["sizes"][t], logas[t][l].detach()).reshape(g["shapes"][t][1:]) for l in range(logas[t].shape[0])]
        assert torch.equal(torch.stack(rows), g["zs_eval"][t + "_z"]), "deterministic mask must be bit-exact: " + t
    l1 = torch.tensor(g["lambda_1"], requires_grad=True)
    l2 = torch.tensor(g["lambda_2"], requires_grad=True)
    lag, es, ts = O.l0_lagrangian(logas, g["params_per_dim"], g["prunable_model_size"], l1, l2, g["target"], g["step"], g["warmup"])
    assert_close(lag, g["lagrangian"], 1e-5, "lagrangian")
    assert_close(es, g["expected_sparsity"], 1e-6, "expected sparsity")
    assert abs(ts - g["target_sparsity"]) < 1e-9
    ltot = lag + sum((z * torch.arange(z.numel()).view(z.shape) / z.numel()).sum() for z in zs.values())
    grads = torch.autograd.grad(ltot, [logas[t] for t in g["types"]] + [l1, l2])
    for a, r in zip(grads, g["grads"]):
        assert_close(a, r, 1e-4, "l0 grad")


def test_kd_losses():
    g = load_golden("retrieval_tiny")["kd"]
    th = O.get_cor_teacher(g["t_hidden"], g["s_hidden"])
    ta = O.get_cor_teacher(g["t_att"], g["s_att"], is_attn=True)
    assert_close(O.get_kd_loss(g["s_hidden"], th), g["hid"], 1e-6, "hidden kd")
    assert_close(O.get_kd_loss(g["s_hidden"], th, is_img=True), g["hid_img"], 1e-6, "image hidden kd")
    assert_close(O.get_kd_loss(g["s_att"], ta, is_attn=True), g["att"], 1e-6, "attn kd")
    assert_close(O.soft_cross_entropy(g["s_logits"] / 2.0, g["t_logits"] / 2.0), g["kl"], 1e-6, "kl")


def test_vqa_oracle():
    """oracle/vqa_oracle.py (train forward with KD outputs, Eff_VQA loss assembly, Lagrangian, eval + rank_answer) against the
    fixture the unmodified reference classes produced (oracle/make_golden_vqa.py)."""
    from oracle import vqa_oracle as V
    from tests.helpers import sd_from_spec
    g = load_golden("vqa_tiny")
    ssd, tsd = sd_from_spec(g["s_sd_spec"]), sd_from_spec(g["t_sd_spec"])
    for sd in (ssd, tsd):
        sd["text_decoder.cls.predictions.decoder.weight"] = sd["text_decoder.bert.embeddings.word_embeddings.weight"]
    for n in g["grad_names"]:
        if not n.startswith("l0_module."):
            ssd[n].requires_grad_()
    b, vis, tvis = g["bert"], g["vis"], g["tvis"]
    s_cfg = dict(vit_layers=vis["num_hidden_layers"], vit_heads=vis["num_attention_heads"], text_layers=6, text_heads=b["num_attention_heads"], dec_layers=3)
    t_cfg = dict(vit_layers=tvis["num_hidden_layers"], vit_heads=tvis["num_attention_heads"], text_layers=12, text_heads=b["num_attention_heads"], dec_layers=6)
    layout, prunable = V.l0_layout(b["hidden_size"], b["intermediate_size"], b["num_attention_heads"], vis["num_hidden_layers"], 6)
    logas = {k: v.clone().requires_grad_() for k, v in g["l0_logas"].items()}
    assert list(layout) == list(g["l0_logas"])
    zs = V.sample_gates(layout, logas, g["eps"])
    batch = (g["image"], g["q_ids"], g["q_atts"], g["a_ids"], g["a_atts"], g["k"], g["weights"])
    l1, l2 = torch.tensor(g["lambda_1"], requires_grad=True), torch.tensor(g["lambda_2"], requires_grad=True)
    order = [t for t in layout if t.endswith("_head")] + [t for t in layout if t.endswith("_intermediate")]

    def lagrangian():
        return O.l0_lagrangian({t: logas[t] for t in order}, {t: layout[t]["per_dim"] for t in order}, prunable, l1, l2, g["scfg"]["sparsity"],
                               g["step"], g["warmup"])[0]
    total, so, to = V.vqa_step(ssd, tsd, s_cfg, t_cfg, batch, zs, lagrangian)
    assert_close(so["loss"], g["s_loss"], TOL, "student loss")
    assert_close(to["loss"], g["t_loss"], TOL, "teacher loss")
    assert_close(so["logits_dict"]["logits"], g["s_logits"], TOL, "student logits")
    assert_close(to["cross_attention_dict"]["decoder_cross_attentions"][-1], g["t_decoder_cross_last"], TOL, "teacher decoder cross attention")
    _, parts = V.kd_total_loss(so, to)
    for name, v in parts.items():
        assert_close(v, g["parts"][name], 1e-5, name)
    assert_close(lagrangian(), g["parts"]["lagrangian"], 1e-5, "lagrangian")
    assert_close(total, g["total"], 1e-5, "total")
    wrt = [logas[n[len("l0_module."):].replace("_int_loga", "_intermediate").replace("_head_loga", "_head")] if "loga" in n
           else (l1 if n.endswith("lambda_1") else l2 if n.endswith("lambda_2") else ssd[n]) for n in g["grad_names"]]
    grads = torch.autograd.grad(total, wrt)
    for n, a, r in zip(g["grad_names"], grads, g["grads"]):
        assert_close(a, r, 2e-4, "grad " + n)
    ze = V.deterministic_gates(layout, logas)
    for k in ze:
        assert torch.equal(ze[k], g["zs_eval"][k]), k
    with torch.no_grad():
        ids, probs = V.eval_forward(ssd, s_cfg, g["image"], g["q_ids"], g["q_atts"], g["l_ids"], g["l_atts"], g["k_test"], ze)
        t_ids, t_probs = V.eval_forward(tsd, t_cfg, g["image"], g["q_ids"], g["q_atts"], g["l_ids"], g["l_atts"], g["k_test"], None)
    assert torch.equal(ids, g["topk_ids"]) and torch.equal(t_ids, g["t_topk_ids"])
    assert_close(probs, g["topk_probs"], 1e-4, "topk probs")
    assert_close(t_probs, g["t_topk_probs"], 1e-4, "teacher topk probs")


def test_itr_oracle():
    """oracle/itr_oracle.py (gated student / un-gated teacher KD forward, Eff_Retrieval loss assembly) against the fixture from
    the unmodified reference classes (oracle/make_golden_itr.py)."""
    from oracle import itr_oracle as R
    g = load_golden("itr_kd_tiny")
    ssd, tsd = sd_from_spec(g["s_sd_spec"]), sd_from_spec(g["t_sd_spec"])
    for n in g["grad_names"]:
        if not n.startswith("l0_module."):
            ssd[n].requires_grad_()
    b, vis, tvis = g["bert"], g["vis"], g["tvis"]
    s_cfg = dict(vit_layers=vis["num_hidden_layers"], vit_heads=vis["num_attention_heads"], text_layers=6, text_heads=b["num_attention_heads"])
    t_cfg = dict(vit_layers=tvis["num_hidden_layers"], vit_heads=tvis["num_attention_heads"], text_layers=12, text_heads=b["num_attention_heads"])
    logas = {k: v.clone().requires_grad_() for k, v in g["l0_logas"].items()}
    heads, inter, H = b["num_attention_heads"], b["intermediate_size"], b["hidden_size"]
    shapes = {k: ([v.shape[0], 1, heads, 1, 1] if k.endswith("_head") else [v.shape[0], 1, 1, inter]) for k, v in logas.items()}
    zs = {k + "_z": O.l0_sample_z(logas[k], g["eps"][k]).reshape(shapes[k]) for k in logas}
    per_head = (H * H * 4 + H * 4) // heads
    per_mlp_layer = H * inter * 2 + H + H * 4
    per_dim = {k: (per_head if k.endswith("_head") else per_mlp_layer // inter) for k in logas}
    prunable = sum(per_head * v.shape[0] * heads if k.endswith("_head") else per_mlp_layer * v.shape[0] for k, v in logas.items())
    l1, l2 = torch.tensor(g["lambda_1"], requires_grad=True), torch.tensor(g["lambda_2"], requires_grad=True)
    order = [k for k in logas if k.endswith("_head")] + [k for k in logas if k.endswith("_intermediate")]

    def lagrangian():
        return O.l0_lagrangian({k: logas[k] for k in order}, per_dim, prunable, l1, l2, g["scfg"]["sparsity"], g["step"], g["warmup"])[0]
    batch = (g["image"], g["text_ids"], g["text_atts"], g["idx"])
    total, so, to = R.itr_step(ssd, tsd, s_cfg, t_cfg, batch, zs, lagrangian)
    assert_close(to["logits_dict"]["itm_head_logits"], g["t_itm_logits"], 1e-4, "teacher itm logits")
    _, parts = R.itr_total_loss(so, to)
    for name, v in parts.items():
        assert_close(v, g["parts"][name], 1e-5, name)
    assert_close(so["loss"]["loss_itc"], g["parts"]["loss_itc"], TOL, "itc")
    assert_close(so["loss"]["loss_itm"], g["parts"]["loss_itm"], 1e-4, "itm")
    assert_close(lagrangian(), g["parts"]["lagrangian"], 1e-5, "lagrangian")
    assert_close(total, g["total"], 1e-5, "total")
    wrt = [logas[n[len("l0_module."):].replace("_int_loga", "_intermediate").replace("_head_loga", "_head")] if "loga" in n
           else (l1 if n.endswith("lambda_1") else l2 if n.endswith("lambda_2") else ssd[n]) for n in g["grad_names"]]
    grads = torch.autograd.grad(total, wrt)
    for n, a, r in zip(g["grad_names"], grads, g["grads"]):
        assert_close(a, r, 2e-4, "grad " + n)


def test_nlvr_oracle():
    """oracle/nlvr_oracle.py against the reference-generated NLVR2 fixture (oracle/make_golden_nlvr.py)."""
    from oracle import nlvr_oracle as N
    g = load_golden("nlvr_kd_tiny")
    ssd, tsd = sd_from_spec(g["s_sd_spec"]), sd_from_spec(g["t_sd_spec"])
    for sd, nominal in ((ssd, 6), (tsd, 12)):      # tied cross-attention K / V: the shared tensor carries the SECOND name's values
        n_text = nominal // 2
        for i in range(nominal - n_text):
            a, b2 = n_text + 2 * i, n_text + 2 * i + 1
            for kv in ("key", "value"):
                for wb in ("weight", "bias"):
                    sd["text_encoder.encoder.layer.%d.crossattention.self.%s.%s" % (a, kv, wb)] = \
                        sd["text_encoder.encoder.layer.%d.crossattention.self.%s.%s" % (b2, kv, wb)]
    b, vis, tvis = g["bert"], g["vis"], g["tvis"]
    s_cfg = dict(vit_layers=vis["num_hidden_layers"], vit_heads=vis["num_attention_heads"], text_layers=6, text_heads=b["num_attention_heads"])
    t_cfg = dict(vit_layers=tvis["num_hidden_layers"], vit_heads=tvis["num_attention_heads"], text_layers=12, text_heads=b["num_attention_heads"])
    heads, inter = b["num_attention_heads"], b["intermediate_size"]
    logas = g["l0_logas"]
    shapes = {k: ([v.shape[0], 1, heads, 1, 1] if k.endswith("_head") else [v.shape[0], 1, 1, inter]) for k, v in logas.items()}
    zs = {k + "_z": O.l0_sample_z(logas[k], g["eps"][k]).reshape(shapes[k]) for k in logas}
    batch = (g["image"], g["text_ids"], g["text_atts"], g["targets"])
    so = N.nlvr_forward(ssd, s_cfg, *batch, zs=zs)
    with torch.no_grad():
        to = N.nlvr_forward(tsd, t_cfg, *batch, zs=None)
    assert_close(so["logits_dict"]["cls_head_logits"], g["s_logits"], 1e-4, "student logits")
    assert_close(to["logits_dict"]["cls_head_logits"], g["t_logits"], 1e-4, "teacher logits")
    total, parts = N.nlvr_total_loss(so, to)
    for name, v in parts.items():
        assert_close(v, g["parts"][name], 1e-5, name)
    assert_close(total + g["parts"]["lagrangian"], g["total"], 1e-5, "total")


GD_PARTS = ("text_hidden_loss", "text_attention_loss", "image_hidden_loss", "image_attention_loss", "itm_pos_hidden_loss", "itm_pos_attn_loss",
            "itm_neg_hidden_loss", "itm_neg_attn_loss", "mlm_hidden_loss", "mlm_attn_loss", "mlm_logits_loss", "itm_logits_loss")


def test_gd_oracle():
    """oracle/gd_oracle.py — the HEADLINE workload's oracle (bench.py's cpu_baseline / --impl reference arm, the GPU GD-step test's
    checker) — against the fixture from the unmodified `models/model_pretrain.py::XVLM` student / teacher and the reference's own
    train-loop statements `GeneralDistill.py:300-376` (oracle/make_golden_gd.py): ITC / ITM / MLM, both logit KLs, every KD MSE term
    the loop computes (re-derived here from the oracle's outputs), the 0.6 / 0.4 mix and 13 gradients."""
    from oracle import gd_oracle as G
    g = load_golden("gd_kd_tiny")
    ssd, tsd = sd_from_spec(g["s_sd_spec"]), sd_from_spec(g["t_sd_spec"])
    for sd in (ssd, tsd):
        sd["text_encoder.cls.predictions.decoder.weight"] = sd["text_encoder.bert.embeddings.word_embeddings.weight"]
    for n in g["grad_names"]:
        ssd[n].requires_grad_()
    ssd["text_encoder.cls.predictions.decoder.weight"] = ssd["text_encoder.bert.embeddings.word_embeddings.weight"]   # tied (same tensor)
    b, vis, tvis = g["bert"], g["vis"], g["tvis"]
    s_cfg = dict(vit_layers=vis["num_hidden_layers"], vit_heads=vis["num_attention_heads"], text_layers=6, text_heads=b["num_attention_heads"])
    t_cfg = dict(vit_layers=tvis["num_hidden_layers"], vit_heads=tvis["num_attention_heads"], text_layers=12, text_heads=b["num_attention_heads"])
    bt = g["batch"]
    batch = [bt[k] for k in ("image", "text_ids", "text_atts", "text_ids_masked", "masked_pos", "masked_ids")]
    total, parts, so = G.gd_step(ssd, tsd, s_cfg, t_cfg, batch, None, None)
    with torch.no_grad():
        to = G.pretrain_forward(tsd, t_cfg, *batch)
    for k in ("loss_itc", "loss_itm", "loss_mlm"):
        assert_close(so["loss"][k], g["loss"][k], 1e-5, k)
    assert_close(so["logits_dict"]["itm_head_logits"], g["s_itm_logits"], 1e-4, "student itm logits")
    assert_close(so["logits_dict"]["mlm_logits"], g["s_mlm_logits"], 1e-4, "student mlm logits")
    assert_close(to["logits_dict"]["itm_head_logits"], g["t_itm_logits"], 1e-4, "teacher itm logits")
    assert_close(to["logits_dict"]["mlm_logits"], g["t_mlm_logits"], 1e-4, "teacher mlm logits")
    assert_close(so["hidden_dict"]["mlm_hidden_states"][-1], g["s_mlm_hidden_last"], 1e-4, "mlm hidden")
    assert_close(so["attention_dict"]["itm_neg_attentions"][-1], g["s_neg_attn_last"], 1e-4, "itm neg attention")
    for d in ("hidden_dict", "attention_dict"):
        for k, v in so[d].items():
            assert len(v) == g["counts"][k], k
            assert len(to[d][k]) == g["t_counts"][k], k

    def kd(name_h, name_a, is_img=False):
        sh, th, sa, ta = so["hidden_dict"][name_h], to["hidden_dict"][name_h], so["attention_dict"][name_a], to["attention_dict"][name_a]
        return (O.get_kd_loss(sh, O.get_cor_teacher(th, sh), is_img=is_img), O.get_kd_loss(sa, O.get_cor_teacher(ta, sa, is_attn=True), is_attn=True))
    mine = {}
    mine["text_hidden_loss"], mine["text_attention_loss"] = kd("text_hidden_states", "text_attentions")
    mine["image_hidden_loss"], mine["image_attention_loss"] = kd("image_hidden_states", "image_attentions", True)
    mine["itm_pos_hidden_loss"], mine["itm_pos_attn_loss"] = kd("itm_pos_hidden_states", "itm_pos_attentions")
    mine["itm_neg_hidden_loss"], mine["itm_neg_attn_loss"] = kd("itm_neg_hidden_states", "itm_neg_attentions")
    mine["mlm_hidden_loss"], mine["mlm_attn_loss"] = kd("mlm_hidden_states", "mlm_attentions")
    mine["mlm_logits_loss"] = O.soft_cross_entropy(so["logits_dict"]["mlm_logits"], to["logits_dict"]["mlm_logits"])
    mine["itm_logits_loss"] = O.soft_cross_entropy(so["logits_dict"]["itm_head_logits"], to["logits_dict"]["itm_head_logits"])
    for name in GD_PARTS:
        assert_close(mine[name], g["parts"][name], 1e-5, name)
    assert_close(parts["loss_small"], g["parts"]["loss_small"], 1e-5, "loss_small")
    assert_close(parts["loss_kd"], g["parts"]["loss_kd"], 1e-5, "loss_kd")
    assert_close(total, g["total"], 1e-5, "loss_in_total")
    grads = torch.autograd.grad(total, [ssd[n] for n in g["grad_names"]])
    for n, a, r in zip(g["grad_names"], grads, g["grads"]):
        assert_close(a, r, 2e-4, "grad " + n)


def test_gd_region_oracle():
    """The region-batch half of the GD iteration (`ret_bbox_loss=True`, GeneralDistill.py:158-260) in oracle/gd_oracle.py against
    tests/golden/gd_region_tiny.pt (unmodified reference model class + the reference's own loop statements): images replicated per
    region inside the local ViT layers under a patch-subset mask, bbox head, L1 + GIoU, the five-term `loss_small`, 16 gradients.
    This oracle is `bench.py --workload gd_region`'s CPU arm."""
    from oracle import gd_oracle as G
    g = load_golden("gd_region_tiny")
    ssd, tsd = sd_from_spec(g["s_sd_spec"]), sd_from_spec(g["t_sd_spec"])
    for n in g["grad_names"]:
        ssd[n].requires_grad_()
    for sd in (ssd, tsd):
        sd["text_encoder.cls.predictions.decoder.weight"] = sd["text_encoder.bert.embeddings.word_embeddings.weight"]
    b, vis, tvis = g["bert"], g["vis"], g["tvis"]
    s_cfg = dict(vit_layers=vis["num_hidden_layers"], vit_heads=vis["num_attention_heads"], text_layers=6, text_heads=b["num_attention_heads"],
                 local_attn_depth=vis["local_attn_depth"])
    t_cfg = dict(vit_layers=tvis["num_hidden_layers"], vit_heads=tvis["num_attention_heads"], text_layers=12, text_heads=b["num_attention_heads"],
                 local_attn_depth=tvis["local_attn_depth"])
    bt = g["batch"]
    batch = [bt[k] for k in ("image", "text_ids", "text_atts", "text_ids_masked", "masked_pos", "masked_ids")]
    region = {k: bt[k] for k in ("idx_to_group_img", "image_atts", "target_bbox", "is_image")}
    total, parts, so = G.gd_step(ssd, tsd, s_cfg, t_cfg, batch, None, None, region=region)
    with torch.no_grad():
        to = G.pretrain_forward(tsd, t_cfg, *batch, region=region)
    for k in ("loss_itc", "loss_itm", "loss_mlm", "loss_bbox", "loss_giou"):
        assert_close(so["loss"][k], g["loss"][k], 1e-5, k)
    assert_close(so["logits_dict"]["itm_head_logits"], g["s_itm_logits"], 1e-4, "student itm logits")
    assert_close(to["logits_dict"]["itm_head_logits"], g["t_itm_logits"], 1e-4, "teacher itm logits")
    assert [tuple(h.shape) for h in so["hidden_dict"]["image_hidden_states"]] == g["s_image_hidden_shapes"]
    assert [tuple(a.shape) for a in so["attention_dict"]["image_attentions"]] == g["s_image_attn_shapes"]
    assert_close(so["hidden_dict"]["bbox_hidden_states"][-1], g["s_bbox_hidden_last"], 1e-4, "bbox fusion hidden")
    for d in ("hidden_dict", "attention_dict"):
        for k, v in so[d].items():
            assert len(v) == g["counts"][k], k
    assert_close(parts["loss_small"], g["parts"]["loss_small"], 1e-5, "loss_small")
    assert_close(parts["loss_kd"], g["parts"]["loss_kd"], 1e-5, "loss_kd")
    assert_close(total, g["total"], 1e-5, "loss_in_total")
    grads = torch.autograd.grad(total, [ssd[n] for n in g["grad_names"]])
    for n, a, r in zip(g["grad_names"], grads, g["grads"]):
        assert_close(a, r, 2e-4, "grad " + n)


def test_gd_oracle_is_device_agnostic():
    """The GD oracle builds every helper tensor on its inputs' device (checked by a dry run on the `meta` device, where mixing in a CPU
    tensor raises): `bench.py --torch-gpu-baseline` runs the same port as eager PyTorch on the GPU (SURVEY 8d's same-box comparator)."""
    from oracle import gd_oracle as G
    g = load_golden("gd_kd_tiny")
    ssd = {k: v.to("meta") for k, v in sd_from_spec(g["s_sd_spec"]).items()}
    tsd = {k: v.to("meta") for k, v in sd_from_spec(g["t_sd_spec"]).items()}
    for sd in (ssd, tsd):
        sd["text_encoder.cls.predictions.decoder.weight"] = sd["text_encoder.bert.embeddings.word_embeddings.weight"]
    b, vis, tvis = g["bert"], g["vis"], g["tvis"]
    s_cfg = dict(vit_layers=vis["num_hidden_layers"], vit_heads=vis["num_attention_heads"], text_layers=6, text_heads=b["num_attention_heads"])
    t_cfg = dict(vit_layers=tvis["num_hidden_layers"], vit_heads=tvis["num_attention_heads"], text_layers=12, text_heads=b["num_attention_heads"])
    batch = [g["batch"][k].to("meta") for k in ("image", "text_ids", "text_atts", "text_ids_masked", "masked_pos", "masked_ids")]
    B = batch[0].shape[0]
    negs = (torch.roll(torch.arange(B), 1).to("meta"), torch.roll(torch.arange(B), -2).to("meta"))
    total, parts, so = G.gd_step(ssd, tsd, s_cfg, t_cfg, batch, negs, negs)
    assert total.device.type == "meta" and total.shape == ()
