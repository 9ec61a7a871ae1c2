"""CPU tests of the product's HOST side: reference-compatible state_dict layout, gate routing / mode slicing / KD output
structures (checked against the reference-generated goldens through a test-only torch op backend, tests/ref_ops.py), the
no-fallback guarantee, and the C-ABI export table.  The kernels themselves are tested on the GPU (tests/test_gpu_*.py)."""
import os
import re

import pytest
import torch

from tests import ref_ops
from tests.helpers import assert_close, load_golden, sd_from_spec

TOL = 2e-5
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _bert_config(c):
    from efficientvlm_b200.eff_bert import BertConfig
    cfg = BertConfig(**{k: v for k, v in c.items() if k not in ("fusion_layer", "encoder_width")})
    cfg.fusion_layer, cfg.encoder_width = c["fusion_layer"], c["encoder_width"]
    return cfg


def _vit(g):
    from efficientvlm_b200.eff_vit import CLIPVisionTransformer
    v = g["cfg"]
    m = CLIPVisionTransformer(v["image_res"], v["patch_size"], v["vision_width"], v["hidden_act"], v["num_attention_heads"],
                              v["attention_dropout"], v["intermediate_size"], v["num_hidden_layers"], local_attn_depth=v["local_attn_depth"])
    m.load_state_dict(sd_from_spec(g["sd_spec"]), strict=True)   # strict: same keys and shapes as the reference
    return m.eval()


def test_abi_exports_every_declared_symbol():
    from efficientvlm_b200 import _lib
    header = open(os.path.join(ROOT, "include", "evlm.h")).read()
    declared = set(re.findall(r"\b(evlm_[a-z0-9_]+)\s*\(", header))
    declared -= {"evlm_gemm_args", "evlm_attn_args", "evlm_mse_pair", "evlm_adamw_group"}
    lib = _lib.load()
    for name in sorted(declared):
        assert hasattr(lib, name), "libevlm_b200.so does not export %s" % name
        assert name in _lib.PROTOTYPES, "no ctypes prototype for %s" % name
    assert set(_lib.PROTOTYPES) <= declared
    header_version = int(re.search(r"#define\s+EVLM_ABI_VERSION\s+(\d+)", header).group(1))
    assert lib.evlm_abi_version() == header_version == _lib.ABI_VERSION


def test_graft_entry_build_runs():
    # the driver's "does it build" check: make is a no-op when the library is current, and build() must accept the current ABI
    import __graft_entry__ as entry
    entry.build()


def test_no_cpu_fallback():
    g = load_golden("vit_tiny")
    vit = _vit(g)
    with pytest.raises(RuntimeError, match="CUDA"):
        vit(g["x"])


def test_state_dict_layouts_match_reference():
    from efficientvlm_b200.eff_bert import BertForMaskedLM, BertLMHeadModel, BertModel
    g = load_golden("bert_tiny")
    cfg = _bert_config(g["cfg"])
    m = BertModel(cfg, add_pooling_layer=False)
    assert {k: tuple(v.shape) for k, v in m.state_dict().items()} == {k: s for k, (s, _) in g["sd_spec"].items()}
    h = load_golden("heads_tiny")
    for cls, key in ((BertForMaskedLM, "mlm_sd_spec"), (BertLMHeadModel, "dec_sd_spec")):
        mm = cls(_bert_config(g["cfg"]))
        assert {k: tuple(v.shape) for k, v in mm.state_dict().items()} == {k: s for k, (s, _) in h[key].items()}
        assert mm.cls.predictions.decoder.weight is mm.bert.embeddings.word_embeddings.weight      # tied


def test_vit_routing(monkeypatch):
    ref_ops.install(monkeypatch)
    g = load_golden("vit_tiny")
    vit = _vit(g)
    out, hid, att = vit(g["x"], output_attentions=True, output_hidden_states=True, head_z=g["head_z"], mlp_z=g["mlp_z"])
    assert_close(out, g["out"], TOL, "vit out")
    assert len(hid) == len(g["hidden"]) and len(att) == len(g["attn"])
    for a, b in zip(hid, g["hidden"]):
        assert_close(a, b, TOL, "hidden")
    for a, b in zip(att, g["attn"]):
        assert_close(a, b, TOL, "attn")
    out2, hid2, att2 = vit(g["x"])
    assert hid2 is None and att2 is None
    assert_close(out2, g["out_nogate"], TOL, "vit nogate")
    r = vit(g["x"], idx_to_group_img=g["idx_to_group"], image_atts=g["image_atts"], output_attentions=True, output_hidden_states=True)
    assert_close(r[0], g["region_out"], TOL, "region out")
    assert_close(r[3], g["region_full"], TOL, "region full")
    assert_close(r[2][1], g["region_attn"][1], TOL, "region attn")


def test_vit_prune_heads_matches_masked(monkeypatch):
    ref_ops.install(monkeypatch)
    g = load_golden("vit_tiny")
    vit = _vit(g)
    hz = torch.ones(2, 1, 2, 1, 1)
    hz[0, 0, 1] = 0
    masked = vit(g["x"], head_z=hz)[0]
    vit.prune_heads({0: [1]})
    assert vit.encoder.layers[0].self_attn.q_proj.weight.shape == (64, 128)
    pruned = vit(g["x"])[0]
    assert_close(pruned, masked, 1e-5, "materialised == masked")


def test_bert_routing(monkeypatch):
    ref_ops.install(monkeypatch)
    from efficientvlm_b200.eff_bert import BertModel
    g = load_golden("bert_tiny")
    bert = BertModel(_bert_config(g["cfg"]), add_pooling_layer=False).eval()
    bert.load_state_dict(sd_from_spec(g["sd_spec"]), strict=True)
    ot = bert(g["ids"], attention_mask=g["atts"], return_dict=True, mode="text", output_attentions=True, output_hidden_states=True,
              head_z=g["text_head_z"], mlp_z=g["text_mlp_z"])
    assert_close(ot.last_hidden_state, g["text_last"], TOL, "text")
    assert len(ot.hidden_states) == 4 and len(ot.attentions) == 3 and len(ot.cross_attentions) == 0
    of = bert(encoder_embeds=ot.last_hidden_state, attention_mask=g["atts"], encoder_hidden_states=g["img"],
              encoder_attention_mask=g["img_atts"], return_dict=True, mode="fusion", output_attentions=True, output_hidden_states=True,
              head_z=g["cross_head_z"], mlp_z=g["cross_mlp_z"])
    assert_close(of.last_hidden_state, g["fus_last"], TOL, "fusion")
    for a, b in zip(of.cross_attentions, g["fus_cross"]):
        assert_close(a, b, TOL, "cross attn")
    for a, b in zip(of.hidden_states, g["fus_hidden"]):
        assert_close(a, b, TOL, "fusion hidden")
    hz = torch.cat([g["text_head_z"], g["cross_head_z"]])
    mz = torch.cat([g["text_mlp_z"], g["cross_mlp_z"]])
    om = bert(g["ids"], attention_mask=g["atts"], encoder_hidden_states=g["img"], encoder_attention_mask=g["img_atts"], return_dict=True,
              mode="multi_modal", output_attentions=True, output_hidden_states=True, head_z=hz, mlp_z=mz)
    assert_close(om.last_hidden_state, g["mm_last"], TOL, "multi_modal (quirk Q1)")
    assert len(om.hidden_states) == len(g["mm_hidden"]) and len(om.cross_attentions) == len(g["mm_cross"])
    ol = bert(g["ids"], attention_mask=g["atts"], encoder_hidden_states=[g["img"], g["img2"]],
              encoder_attention_mask=[g["img_atts"], g["img_atts"]], return_dict=True, mode="multi_modal")
    assert_close(ol.last_hidden_state, g["list_last"], TOL, "nlvr list")
    tup = bert(g["ids"], attention_mask=g["atts"], return_dict=False, mode="text")
    assert_close(tup[0], bert(g["ids"], attention_mask=g["atts"], mode="text")[0], 1e-7, "tuple output")
    with pytest.raises(ValueError):
        bert(g["ids"], attention_mask=g["atts"], mode="bogus")


def _tied(sd):
    """The fixture models tie decoder.weight to the word embeddings AFTER initialisation (as transformers 4.12.5's
    init_weights did); a tied module loads both keys into one tensor, so make them agree first."""
    sd["cls.predictions.decoder.weight"] = sd["bert.embeddings.word_embeddings.weight"]
    return sd


def test_heads(monkeypatch):
    ref_ops.install(monkeypatch)
    from efficientvlm_b200.eff_bert import BertForMaskedLM, BertLMHeadModel
    g, b = load_golden("heads_tiny"), load_golden("bert_tiny")
    mlm = BertForMaskedLM(_bert_config(b["cfg"])).eval()
    mlm.load_state_dict(_tied(sd_from_spec(g["mlm_sd_spec"])), strict=True)
    mo = mlm(b["ids"], attention_mask=b["atts"], encoder_hidden_states=b["img"], encoder_attention_mask=b["img_atts"], return_dict=True,
             labels=g["labels"], masked_pos=g["masked_pos"], output_attentions=True, output_hidden_states=True)
    assert_close(mo.loss, g["mlm_loss"], TOL, "mlm loss")
    assert_close(mo.logits, g["mlm_logits"], TOL, "mlm logits")
    dec = BertLMHeadModel(_bert_config(b["cfg"]), label_smoothing=0.1).eval()
    dec.load_state_dict(_tied(sd_from_spec(g["dec_sd_spec"])), strict=True)
    do = dec(b["ids"], attention_mask=b["atts"], encoder_hidden_states=b["img"], encoder_attention_mask=b["img_atts"], labels=g["dlabels"],
             return_dict=True, reduction="none", head_z=g["dec_head_z"], mlp_z=g["dec_mlp_z"])
    assert_close(do.loss, g["dec_loss_none_ls"], TOL, "decoder loss")
    assert_close(do.logits, g["dec_logits"], TOL, "decoder logits")
    dec.label_smoothing = 0.0
    do0 = dec(b["ids"], attention_mask=b["atts"], encoder_hidden_states=b["img"], encoder_attention_mask=b["img_atts"], labels=g["dlabels"],
              return_dict=True, reduction="mean")
    assert_close(do0.loss, g["dec_loss_mean"], TOL, "decoder loss mean")
    B = b["ids"].shape[0]
    s1 = dec(b["ids"][:, :4], attention_mask=torch.ones(B, 4, dtype=torch.long), encoder_hidden_states=b["img"],
             encoder_attention_mask=b["img_atts"], return_dict=True, use_cache=True)
    assert s1.past_key_values[0][0].shape[2] == 4
    s2 = dec(b["ids"][:, 4:5], attention_mask=torch.ones(B, 5, dtype=torch.long), encoder_hidden_states=b["img"],
             encoder_attention_mask=b["img_atts"], return_dict=True, use_cache=True, past_key_values=s1.past_key_values)
    assert_close(s2.logits, g["step2_logits"], TOL, "cached decode step")


def _retrieval_model(g):
    from efficientvlm_b200.distill import EffXVLMforRetrieval
    cfg = dict(g["cfg"])
    cfg["vision_config"] = dict(g["vis"])
    cfg["text_encoder"] = None
    import efficientvlm_b200.eff_bert as eb
    orig = eb.BertConfig.__init__

    def patched(self, **kw):
        merged = dict(g["bert"])
        merged.update(kw)
        orig(self, **merged)
    eb.BertConfig.__init__ = patched          # the fixture's tiny BERT instead of bert-base-uncased defaults
    try:
        m = EffXVLMforRetrieval(cfg)
    finally:
        eb.BertConfig.__init__ = orig
    sd = sd_from_spec(g["sd_spec"])
    for k, v in g["l0_logas"].items():
        name = {"vision_head": "vision_head_loga", "text_head": "text_head_loga", "cross_head": "cross_head_loga",
                "vision_intermediate": "vision_int_loga", "text_intermediate": "text_int_loga", "cross_intermediate": "cross_int_loga"}[k]
        sd["l0_module." + name] = v
    m.load_state_dict(sd, strict=True)
    return m.eval()


def _argmax_negatives(model):
    from oracle import xvlm_oracle as O

    def sampler(image_feat, text_feat, idx=None):
        w_i2t, w_t2i = O.itm_negative_weights(image_feat.detach(), text_feat.detach(), model.temp.detach(), idx)
        return w_t2i.argmax(1), w_i2t.argmax(1)
    return sampler


def test_retrieval_model_losses_and_kd_structure(monkeypatch):
    ref_ops.install(monkeypatch)
    g = load_golden("retrieval_tiny")
    model = _retrieval_model(g)
    model.sample_itm_negatives = _argmax_negatives(model)
    loss_itc, loss_itm = model(g["image"], g["text_ids"], g["text_atts"], idx=g["idx"])
    assert_close(loss_itc, g["loss_itc"], TOL, "itc (idx, deterministic masks)")
    assert_close(loss_itm, g["loss_itm"], 1e-4, "itm (idx)")
    l2, m2 = model(g["image"], g["text_ids"], g["text_atts"], idx=None)
    assert_close(l2, g["loss_itc_noidx"], TOL, "itc")
    assert_close(m2, g["loss_itm_noidx"], 1e-4, "itm")
    it = iter([g["eps"][k] for k in model.l0_module.types])
    model.l0_module.get_eps = lambda size: next(it)
    res = model(g["image"], g["text_ids"], g["text_atts"], idx=g["idx"], output_attentions=True, output_hidden_states=True)
    assert_close(res["loss"]["loss_itc"], g["kd_loss_itc"], TOL, "kd itc")
    assert_close(res["loss"]["loss_itm"], g["kd_loss_itm"], 1e-4, "kd itm")
    assert_close(res["logits_dict"]["itm_head_logits"], g["kd_itm_logits"], 1e-4, "itm logits")
    assert len(res["hidden_dict"]["image_hidden_states"]) == len(g["kd_image_hidden"])
    for a, b in zip(res["attention_dict"]["text_attentions"], g["kd_text_attn"]):
        assert_close(a, b, TOL, "text attn")
    for a, b in zip(res["cross_attention_dict"]["itm_neg_cross_attentions"], g["kd_neg_cross"]):
        assert_close(a, b, 1e-4, "neg cross attn")
    tot = res["loss"]["loss_itc"] + res["loss"]["loss_itm"]
    params = dict(model.named_parameters())
    grads = torch.autograd.grad(tot, [params[n] for n in g["grad_names"]])
    for n, a, b in zip(g["grad_names"], grads, g["grads"]):
        assert_close(a, b, 2e-4, "grad " + n)


def test_l0_module(monkeypatch):
    ref_ops.install(monkeypatch)
    g = load_golden("retrieval_tiny")
    l0g = load_golden("l0_tiny")
    model = _retrieval_model(g)
    l0 = model.l0_module
    assert list(l0.types) == l0g["types"]
    assert l0.prunable_model_size == l0g["prunable_model_size"]
    assert {k: int(v) for k, v in l0.parameters_per_dim.items()} == {k: int(v) for k, v in l0g["params_per_dim"].items()}
    with torch.no_grad():
        for k in l0.types:
            l0.z_logas[k].copy_(l0g["logas"][k])
        l0.lambda_1.fill_(l0g["lambda_1"])
        l0.lambda_2.fill_(l0g["lambda_2"])
    it = iter([l0g["eps"][k] for k in l0.types])
    l0.get_eps = lambda size: next(it)
    zs = l0(training=True)
    assert list(zs.keys()) == l0g["zs_order"]
    for k in zs:
        assert zs[k].shape == l0g["zs_train"][k].shape
        assert_close(zs[k], l0g["zs_train"][k], 1e-6, k)
    ze = l0(training=False)
    for k in ze:
        assert torch.equal(ze[k], l0g["zs_eval"][k]), k
    l0.set_lagrangian_warmup_steps(l0g["warmup"])
    lag, es, ts = l0.lagrangian_regularization(l0g["step"])
    assert_close(lag, l0g["lagrangian"], 1e-5, "lagrangian")
    assert abs(ts - l0g["target_sparsity"]) < 1e-9
    assert l0.calculate_model_size(ze) == l0g["model_size"]
    names = [n for n, _ in l0.named_parameters()]
    assert names[-2:] == ["lambda_1", "lambda_2"] and all("lambda" not in n for n in names[:-2])


def test_kd_helpers(monkeypatch):
    ref_ops.install(monkeypatch)
    from efficientvlm_b200 import distill as D
    g = load_golden("retrieval_tiny")["kd"]
    th = D.get_cor_teacher(g["t_hidden"], g["s_hidden"])
    ta = D.get_cor_teacher(g["t_att"], g["s_att"], is_attn=True)
    assert_close(D.get_kd_loss(g["s_hidden"], th, False), g["hid"], 1e-6, "hidden")
    assert_close(D.get_kd_loss(g["s_hidden"], th, False, is_img=True), g["hid_img"], 1e-6, "image hidden (drops idx 6)")
    assert_close(D.get_kd_loss(g["s_att"], ta, True), g["att"], 1e-6, "attention (x key_len)")
    assert_close(D.soft_cross_entropy(g["s_logits"] / 2.0, g["t_logits"] / 2.0), g["kl"], 1e-6, "kl")


def test_outputs_container():
    from efficientvlm_b200.outputs import MaskedLMOutput
    o = MaskedLMOutput(loss=None, logits=torch.zeros(1), hidden_states=(1, 2))
    assert o[0] is o.logits and o[1] == (1, 2) and len(o) == 2 and o["logits"] is o.logits


def test_vqa_models_vs_reference_golden(monkeypatch):
    """EffXVLMForVQA / XVLMForVQA host logic (gate routing, k-answer replication through the cross-attention index, loss mix of
    Eff_VQA.py:105-181, rank_answer) against the reference-generated fixture; arithmetic = the test-only torch backend."""
    from tests.helpers import Tokens, arm_eps, vqa_models
    from efficientvlm_b200.vqa import tile, vqa_loss
    ref_ops.install(monkeypatch)
    g = load_golden("vqa_tiny")
    student, teacher = vqa_models(g)
    q, a = Tokens(g["q_ids"], g["q_atts"]), Tokens(g["a_ids"], g["a_atts"])
    arm_eps(student.l0_module, g["eps"])
    so = student(g["image"], q, a, train=True, k=g["k"], weights=g["weights"], output_attentions=True, output_hidden_states=True)
    with torch.no_grad():
        to = teacher(g["image"], q, a, train=True, k=g["k"], weights=g["weights"], output_attentions=True, output_hidden_states=True)
    assert_close(so["loss"], g["s_loss"], TOL, "student task loss")
    assert_close(to["loss"], g["t_loss"], TOL, "teacher task loss")
    assert_close(so["logits_dict"]["logits"], g["s_logits"], TOL, "student logits")
    assert_close(to["logits_dict"]["logits"], g["t_logits"], TOL, "teacher logits")
    c = g["counts"]
    assert [len(so["hidden_dict"][n]) for n in ("image_hidden_states", "text_hidden_states", "decoder_hidden_states")] == [c["s_img_h"], c["s_text_h"], c["s_dec_h"]]
    assert [len(to["hidden_dict"][n]) for n in ("image_hidden_states", "text_hidden_states", "decoder_hidden_states")] == [c["t_img_h"], c["t_text_h"], c["t_dec_h"]]
    assert [len(so["attention_dict"][n]) for n in ("image_attentions", "text_attentions", "decoder_attentions")] == [g["vis"]["num_hidden_layers"], c["s_text_a"], c["s_dec_a"]]
    assert [len(so["cross_attention_dict"][n]) for n in ("cross_attentions", "decoder_cross_attentions")] == [c["s_cross_a"], c["s_dec_c"]]
    assert_close(so["hidden_dict"]["image_hidden_states"][-1], g["s_image_hidden_last"], TOL, "image hidden")
    assert_close(so["hidden_dict"]["text_hidden_states"][-1], g["s_text_hidden_last"], TOL, "question hidden")
    for name, key in (("decoder_hidden_states", "s_decoder_hidden"),):
        for x, y in zip(so["hidden_dict"][name], g[key]):
            assert_close(x, y, TOL, name)
    for x, y in zip(so["attention_dict"]["decoder_attentions"], g["s_decoder_attn"]):
        assert_close(x, y, TOL, "decoder self attention")
    for x, y in zip(so["cross_attention_dict"]["decoder_cross_attentions"], g["s_decoder_cross"]):
        assert_close(x, y, TOL, "decoder cross attention")
    for x, y in zip(so["cross_attention_dict"]["cross_attentions"], g["s_cross_attn"]):
        assert_close(x, y, TOL, "question cross attention")
    total, parts = vqa_loss(so, to, student.l0_module, g["step"], 1.0)
    for name in ("text_hidden", "text_attention", "cross_hidden", "cross_self_attention", "cross_attention", "image_hidden", "image_attention",
                 "decoder_hidden", "decoder_attention", "decoder_cross", "logits"):
        assert_close(parts["kd_" + name], g["parts"][name], 1e-4, name)
    assert_close(parts["loss_lagrangian"], g["parts"]["lagrangian"], 1e-4, "lagrangian")
    assert abs(parts["target_sparsity"] - g["parts"]["target_sparsity"]) < 1e-9
    assert_close(total, g["total"], 1e-5, "total loss")
    sp = dict(student.named_parameters())
    grads = torch.autograd.grad(total, [sp[n] for n in g["grad_names"]])
    for n, x, y in zip(g["grad_names"], grads, g["grads"]):
        assert_close(x, y, 2e-4, "grad " + n)
    arm_eps(student.l0_module, g["eps"])
    assert_close(student(g["image"], q, a, train=True, k=g["k"], weights=g["weights"]), g["loss_plain"], TOL, "task loss without KD outputs")
    assert_close(student(g["image"], q, a, train=True, k=g["k"], weights=g["weights"], stop_prune=True), g["loss_stop"], TOL, "stop_prune")
    al = Tokens(g["l_ids"], g["l_atts"])
    ids, probs = student(g["image"], q, al, train=False, k=g["k_test"])
    assert torch.equal(ids, g["topk_ids"])
    assert_close(probs, g["topk_probs"], 1e-4, "re-ranked probabilities")
    t_ids, t_probs = teacher(g["image"], q, al, train=False, k=g["k_test"])
    assert torch.equal(t_ids, g["t_topk_ids"])
    assert_close(t_probs, g["t_topk_probs"], 1e-4, "teacher re-ranked probabilities")
    x = torch.arange(6).view(3, 2)
    assert torch.equal(tile(x, 0, 2), x[[0, 0, 1, 1, 2, 2]])


def _pruned_vqa_model(g):
    """6-vision-layer tiny student of tests/golden/vqa_pruned_tiny.pt, materialised with efficientvlm_b200.prune."""
    from tests.helpers import build_with_tiny_bert
    from efficientvlm_b200 import prune
    from efficientvlm_b200.vqa import EffXVLMForVQA
    v = load_golden("vqa_tiny")
    cfg = dict(v["scfg"], vision_config=dict(g["vis"]), text_encoder=None)
    m = build_with_tiny_bert(EffXVLMForVQA, cfg, v["bert"])
    sd = sd_from_spec(g["sd_spec"])
    sd["text_decoder.cls.predictions.decoder.weight"] = sd["text_decoder.bert.embeddings.word_embeddings.weight"]
    m.load_state_dict(sd, strict=True)
    m.eval()
    prune.update_params(m, g["zs"])
    prune.prune_model_with_z(g["zs"], m)
    return m, v


def test_materialised_vqa_model_matches_reference_utilities(monkeypatch):
    """prune.update_params + prune.prune_model_with_z reproduce utils/vqa_utils.py:37-313 (run by oracle/make_golden_vqa.py on the
    reference model): same pruned parameter shapes, same folded weights, same fake_forward ranking."""
    from tests.helpers import Tokens
    ref_ops.install(monkeypatch)
    g = load_golden("vqa_pruned_tiny")
    m, v = _pruned_vqa_model(g)
    shapes = {k: tuple(t.shape) for k, t in m.state_dict().items()}
    assert shapes == g["pruned_shapes"]
    for k, t in g["probe"].items():
        assert_close(m.state_dict()[k], t, 1e-6, "folded weight " + k)
    ids, probs, _ = m.fake_forward(v["image"], Tokens(v["q_ids"], v["q_atts"]), Tokens(v["l_ids"], v["l_atts"]), k=v["k_test"])
    assert torch.equal(ids, g["topk_ids"])
    assert_close(probs, g["topk_probs"], 1e-4, "pruned-model probabilities")
    # a fully pruned layer is refused loudly instead of producing a module the forward cannot run
    from efficientvlm_b200 import prune
    zs = {"vision_head_z": torch.zeros(6, 1, 2, 1, 1)}
    with pytest.raises(NotImplementedError):
        prune.prune_model_with_z(zs, m)


def test_itr_kd_step_vs_reference_golden(monkeypatch):
    """EffXVLMforRetrieval student + XVLMforRetrieval teacher + the loss mix of Eff_Retrieval.py:96-178 (host logic)."""
    from tests.helpers import run_itr_kd_step
    ref_ops.install(monkeypatch)
    run_itr_kd_step(load_golden("itr_kd_tiny"), "cpu", 1e-4, 1e-5, 2e-4)


def test_nlvr_kd_step_vs_reference_golden(monkeypatch):
    """EffXVLMForNLVR student + XVLMForNLVR teacher (two images per text, tied cross-attention K/V, list-of-images fusion
    layers) + the loss mix of Eff_NLVR.py:100-157 (host logic)."""
    from tests.helpers import run_nlvr_kd_step
    ref_ops.install(monkeypatch)
    run_nlvr_kd_step(load_golden("nlvr_kd_tiny"), "cpu", 1e-4, 1e-5, 2e-4)


def test_caption_kd_step_vs_reference_golden(monkeypatch):
    """EffXVLMForCaptioning student + XVLMForCaptioning teacher, the loss mix of Eff_Captioning.py:96-148 and the greedy decode
    loop (host logic; a whitespace tokenizer stands in for bert-base-uncased on both sides)."""
    from tests.helpers import run_caption_kd_step
    ref_ops.install(monkeypatch)
    run_caption_kd_step(load_golden("caption_kd_tiny"), "cpu", 1e-4, 1e-5, 2e-4, exact_decode=True)


def test_beam_search_vs_published_algorithm(monkeypatch):
    """`generate(num_beams > 1)` — what `Eff_Captioning.py:201-202` evaluates with — against oracle/beam_search_oracle.py (host logic;
    transformers 4.12.5's beam search restated, parity unpinned): identical sequences for four (beams, max_length, min_length,
    repetition_penalty) settings, including one that hits max_length on the first step."""
    from tests.helpers import run_beam_search_vs_oracle
    ref_ops.install(monkeypatch)
    run_beam_search_vs_oracle(load_golden("caption_kd_tiny"), "cpu", exact=True)


def test_itr_rerank_evaluation_vs_reference_golden(monkeypatch):
    """retrieval_eval.evaluation / rerank_scores / itm_eval against the reference driver's own `evaluation` and `itm_eval`
    (Eff_Retrieval.py:216-378, extracted and run by oracle/make_golden_itr_eval.py): similarity matrix, candidate sets, ITM scores,
    recall metrics, the two-rank row split and every batching mode (host logic)."""
    from tests.helpers import run_itr_eval
    ref_ops.install(monkeypatch)
    run_itr_eval(load_golden("itr_eval_tiny"), "cpu", 1e-5, 1e-4, exact_candidates=True)


def test_itm_eval_matches_reference_loop():
    """itm_eval (vectorised ranks) == the reference's per-row loop (Eff_Retrieval.py:335-378) on random score matrices with -100 fills."""
    import numpy as np
    from efficientvlm_b200.retrieval_eval import itm_eval
    rng = np.random.default_rng(3)
    n_img, per = 23, 5
    n_txt = n_img * per
    s_i2t = np.where(rng.random((n_img, n_txt)) < 0.3, rng.standard_normal((n_img, n_txt)), -100.0).astype(np.float32)
    s_t2i = np.where(rng.random((n_txt, n_img)) < 0.5, rng.standard_normal((n_txt, n_img)), -100.0).astype(np.float32)
    img2txt = {i: list(range(i * per, (i + 1) * per)) for i in range(n_img)}
    txt2img = {t: t // per for t in range(n_txt)}
    ranks = np.zeros(n_img)
    for index, score in enumerate(s_i2t):
        inds = np.argsort(score)[::-1]
        ranks[index] = min(np.where(inds == i)[0][0] for i in img2txt[index])
    tr = [100.0 * len(np.where(ranks < n)[0]) / len(ranks) for n in (1, 5, 10)]
    ranks = np.zeros(n_txt)
    for index, score in enumerate(s_t2i):
        ranks[index] = np.where(np.argsort(score)[::-1] == txt2img[index])[0][0]
    ir = [100.0 * len(np.where(ranks < n)[0]) / len(ranks) for n in (1, 5, 10)]
    out = itm_eval(s_i2t, s_t2i, txt2img, img2txt)
    assert [out["txt_r1"], out["txt_r5"], out["txt_r10"]] == tr
    assert [out["img_r1"], out["img_r5"], out["img_r10"]] == ir
    assert out["r_mean"] == (sum(tr) / 3 + sum(ir) / 3) / 2


@pytest.mark.parametrize("batch_passes", [True, False])
def test_gd_kd_step_vs_reference_golden(monkeypatch, batch_passes):
    """The HEADLINE workload's host logic — `distill.XVLM` student / teacher (models/model_pretrain.py), `gd_kd_losses`, `gd_loss` — against
    the fixture from the unmodified reference model class and the reference's own train-loop statements (GeneralDistill.py:300-376,
    oracle/make_golden_gd.py), in the batched-pass schedule the bench runs (text 2B, fusion 4B) and pass by pass."""
    from tests.helpers import run_gd_kd_step
    ref_ops.install(monkeypatch)
    run_gd_kd_step(load_golden("gd_kd_tiny"), "cpu", 1e-4, 1e-5, 2e-4, batch_passes=batch_passes)


def test_materialised_retrieval_model_reproduces_masked_evaluation(monkeypatch):
    """prune.materialize on the retrieval student (layer counts taken from the gate tensors, not the literals 6 / 3 / 3 of
    utils/xvlm_utils.py:37-245): the physically smaller, gate-free model reproduces the similarity matrix and the re-rank scores the
    REFERENCE's masked evaluation produced (tests/golden/itr_eval_tiny.pt) — text mode + fusion mode, where quirk Q1 does not apply."""
    from efficientvlm_b200 import prune
    from efficientvlm_b200 import retrieval_eval as RE
    from tests.helpers import itr_eval_setup
    ref_ops.install(monkeypatch)
    g = load_golden("itr_eval_tiny")
    model, _, tokenizer = itr_eval_setup(g, "cpu")
    with torch.no_grad():
        zs = model.l0_module.forward(training=False)
        before = sum(p.numel() for n, p in model.named_parameters() if not n.startswith("l0_module."))
        prune.materialize(model, zs)
        after = sum(p.numel() for n, p in model.named_parameters() if not n.startswith("l0_module."))
        kept = sum(int(zs[k].sum()) for k in zs if k.endswith("intermediate_z"))
        total = sum(zs[k].numel() for k in zs if k.endswith("intermediate_z"))
        H = g["bert"]["hidden_size"]
        assert before - after == (total - kept) * (2 * H + 1)        # fc1 row + bias + fc2 column per pruned FFN unit (no head is pruned here)
        enc = tokenizer(g["texts"], padding="max_length", truncation=True, max_length=g["config"]["max_tokens"], return_tensors="pt")
        text_feats = model.get_text_embeds(enc.input_ids, enc.attention_mask)
        image_feats, _ = model.get_vision_embeds(g["images"])
        sims = model.get_features(image_embeds=image_feats) @ model.get_features(text_embeds=text_feats).t()
        assert_close(sims, g["sims"], 1e-5, "similarity matrix of the materialised model")
        a, b = RE.rerank_scores(model, image_feats, text_feats, enc.attention_mask, g["sims"], g["config"]["k_test"], rank=0, world=1, group_rows=3)
    assert_close(a, g["score_i2t"], 1e-4, "image->text re-rank scores")
    assert_close(b, g["score_t2i"], 1e-4, "text->image re-rank scores")


def test_gd_region_step_vs_reference_golden(monkeypatch):
    """SURVEY 8(f) row 2 — the region / bbox half of a GD iteration (`ret_bbox_loss=True`): local-attention ViT layers on
    [regions + images] rows, `predict_bbox`, `get_bbox_loss` (L1 + GIoU, `is_image` rows excluded), the loss mix of
    GeneralDistill.py:257-259 and 16 gradients against the fixture from the unmodified reference (oracle/make_golden_gd.py)."""
    from tests.helpers import run_gd_region_step
    ref_ops.install(monkeypatch)
    run_gd_region_step(load_golden("gd_region_tiny"), "cpu", 1e-4, 1e-5, 2e-4)


def test_flat_adamw_state_dict_roundtrip_and_reference_layout():
    """the drivers checkpoint `optimizer.state_dict()` (GeneralDistill.py:422,430) and resume with
    `optimizer.load_state_dict()` (:517).  The layout is torch's ({"state": {i: step/exp_avg/exp_avg_sq}, "param_groups": [...]}),
    loading copies INTO the arenas, and a state written by a torch AdamW over the same groups loads."""
    import io
    from efficientvlm_b200.optim import FlatAdamW
    torch.manual_seed(0)
    net = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.LayerNorm(7), torch.nn.Linear(7, 3))
    groups = [{"params": [p for n, p in net.named_parameters() if n.endswith("weight")], "weight_decay": 0.01, "lr": 1e-3},
              {"params": [p for n, p in net.named_parameters() if n.endswith("bias")], "weight_decay": 0.0, "lr": 2e-3}]
    opt = FlatAdamW(groups)
    assert opt.state_dict()["state"] == {} and [g["params"] for g in opt.state_dict()["param_groups"]] == [[0, 1, 2], [3, 4, 5]]
    # pretend 4 steps happened (the update kernel itself is CUDA-only; its state is what we checkpoint)
    opt.state_step = 4
    for g in opt.param_groups:
        g["m"].normal_()
        g["v"].uniform_()
        g["lr"] *= 0.5
    sd = opt.state_dict()
    assert sd["state"][4]["step"] == 4 and sd["state"][1]["exp_avg"].shape == net[1].weight.shape
    assert sd["param_groups"][1]["lr"] == 1e-3 and sd["param_groups"][1]["initial_lr"] == 2e-3 and sd["param_groups"][0]["betas"] == (0.9, 0.98)
    buf = io.BytesIO()
    torch.save({"optimizer": sd}, buf)         # what utils/checkpointer.py does with save_obj
    buf.seek(0)
    loaded = torch.load(buf, weights_only=False)["optimizer"]
    net2 = torch.nn.Sequential(torch.nn.Linear(5, 7), torch.nn.LayerNorm(7), torch.nn.Linear(7, 3))
    opt2 = FlatAdamW([{"params": [p for n, p in net2.named_parameters() if n.endswith("weight")], "weight_decay": 0.01, "lr": 1e-3},
                      {"params": [p for n, p in net2.named_parameters() if n.endswith("bias")], "weight_decay": 0.0, "lr": 2e-3}])
    ptrs = [g["m"].data_ptr() for g in opt2.param_groups]
    opt2.load_state_dict(loaded)
    assert opt2.state_step == 4 and ptrs == [g["m"].data_ptr() for g in opt2.param_groups]
    for a, b in zip(opt.param_groups, opt2.param_groups):
        assert a["lr"] == b["lr"] and a["initial_lr"] == b["initial_lr"]
        for q, off in zip(a["params"], a["offsets"]):           # (the alignment padding between parameters is not part of the state)
            sl = slice(off, off + q.numel())
            assert torch.equal(a["m"][sl], b["m"][sl]) and torch.equal(a["v"][sl], b["v"][sl])
    # a torch.optim.AdamW checkpoint over the same groups
    ref = torch.optim.AdamW([{"params": g["params"], "lr": g["lr"], "weight_decay": g["weight_decay"]} for g in groups], betas=(0.9, 0.98))
    for p in net.parameters():
        p.grad = torch.randn_like(p)
    ref.step()
    opt2.load_state_dict(ref.state_dict())
    assert opt2.state_step == 1
    off = opt2.param_groups[0]["offsets"][2]
    want = ref.state_dict()["state"][2]["exp_avg"]
    assert torch.equal(opt2.param_groups[0]["m"][off:off + want.numel()].view_as(want), want)
    with pytest.raises(ValueError):
        opt2.load_state_dict({"state": {}, "param_groups": sd["param_groups"][:1]})


def test_later_flat_adamw_takes_ownership_of_shared_parameters():
    """the reference lists the l0_module gates in the main optimizer AND in the L0 optimizers (quirk Q11).  Here the
    optimizer built last owns the parameter; the earlier one must not re-bind p.grad / p._evlm_main_grad back into its own arena."""
    from efficientvlm_b200.optim import FlatAdamW
    w, gate = torch.nn.Parameter(torch.ones(4)), torch.nn.Parameter(torch.zeros(3))
    main = FlatAdamW([{"params": [w, gate]}])
    l0 = FlatAdamW([{"params": [gate]}])
    assert gate._evlm_owner is l0 and w._evlm_owner is main
    view = gate.grad.data_ptr()
    assert view == l0.param_groups[0]["g"].data_ptr()
    main.zero_grad()
    main._gather_stray_grads()
    assert gate.grad.data_ptr() == view and gate._evlm_main_grad.data_ptr() == view
    gate.grad.add_(1.0)
    assert float(l0.param_groups[0]["g"][:3].sum()) == 3.0 and float(main.param_groups[0]["g"].sum()) == 0.0


def test_scheduler_keeps_the_lambdalr_surface():
    """Captioning_pretrain.py:32-50 / NLVR_pretrain.py:175 read `scheduler.optimizer` and re-run
    `scheduler.__init__(optimizer, lr_lambda, last_epoch=-1)`; checkpoints hold LambdaLR's `last_epoch`."""
    import types
    from efficientvlm_b200.optim import LinearWarmupDecay
    opt = types.SimpleNamespace(param_groups=[{"lr": 0.5, "initial_lr": 0.5}, {"lr": 0.02, "initial_lr": 0.02}])
    s = LinearWarmupDecay(opt, 20, 4)
    assert s.optimizer is opt and s.last_epoch == 0 and s.get_last_lr() == [0.0, 0.0]
    seq = []
    for _ in range(7):
        s.step()
        seq.append(opt.param_groups[0]["lr"])

    def lr_lambda(step):
        return step / 4.0 if step < 4 else max(0.0, (20 - step) / 16.0)
    if s.optimizer == opt:
        s.__init__(opt, lr_lambda, last_epoch=-1)
    assert s.last_epoch == 0 and opt.param_groups[1]["lr"] == 0.0
    seq2 = []
    for _ in range(7):
        s.step()
        seq2.append(opt.param_groups[0]["lr"])
    assert seq == seq2
    # a torch LambdaLR state loads
    w = torch.nn.Parameter(torch.zeros(1))
    topt = torch.optim.SGD([{"params": [w], "lr": 0.5}, {"params": [torch.nn.Parameter(torch.zeros(1))], "lr": 0.02}], lr=0.5)
    ref = torch.optim.lr_scheduler.LambdaLR(topt, lr_lambda, last_epoch=-1)
    for _ in range(11):
        topt.step()
        ref.step()
    opt3 = types.SimpleNamespace(param_groups=[{"lr": 0.5, "initial_lr": 0.5}, {"lr": 0.02, "initial_lr": 0.02}])
    s3 = LinearWarmupDecay(opt3, 20, 4)
    s3.load_state_dict(ref.state_dict())
    assert s3.last_epoch == ref.last_epoch == 11 and [g["lr"] for g in opt3.param_groups] == ref.get_last_lr()
    s3.step(); topt.step(); ref.step()
    assert [g["lr"] for g in opt3.param_groups] == ref.get_last_lr()
    # our own state round-trips, and the round-1 key still loads
    s4 = LinearWarmupDecay(types.SimpleNamespace(param_groups=[{"lr": 0.5, "initial_lr": 0.5}, {"lr": 0.02, "initial_lr": 0.02}]), 20, 4)
    s4.load_state_dict(s3.state_dict())
    assert s4.get_last_lr() == s3.get_last_lr() and s4.last_epoch == s3.last_epoch
    s4.load_state_dict({"last_step": 3, "total": 20, "warm": 4})
    assert s4.last_epoch == 3 and s4.get_last_lr()[0] == 0.5 * 0.75


def test_device_decode_checker_runs_on_the_reference_backend(monkeypatch):
    """The `-m gpu` captioning test compares device decodes with an on-box fp32 reference pass (tests/helpers.py::
    _decode_on_device_vs_fp32_reference); here the same checker runs with the oracle operators standing in for the device too, so its
    own logic (recorded draws, teacher-forced re-score, first-divergence margin) is exercised on every CPU run."""
    from tests import helpers as H
    ref_ops.install(monkeypatch)
    g = load_golden("caption_kd_tiny")
    student, _ = H.caption_models(g)
    student.tokenizer(g["captions"])
    caps = student.generate(g["image"], greedy=True, max_length=10)
    assert caps == g["greedy_captions"]
    H._decode_on_device_vs_fp32_reference(g, student, g["image"], caps)
    # a flipped token with a wide reference margin is NOT accepted as a tie
    orig = student.generate

    def flipped(image, **kw):
        out = orig(image, **kw)
        if kw.get("greedy") and kw.get("return_ids"):
            ids = out[1].clone()
            ids[0, student.prompt_length] = (ids[0, student.prompt_length] + 1) % g["bert"]["vocab_size"]
            return out[0], ids
        return out
    student.generate = flipped
    with pytest.raises(AssertionError):
        H._decode_on_device_vs_fp32_reference(g, student, g["image"], caps)


def test_data_inplace_optimizers_invalidate_the_weight_shadows():
    """`p.data.add_()` (how transformers-4.12.5 AdamW updates, /root/reference/optim.py:1,67) leaves `_version` and
    `data_ptr()` unchanged, so the bf16 shadow cache cannot see it; the global optimizer post-step hook bumps the per-parameter epoch."""
    from efficientvlm_b200 import ops

    class DataAdamLike(torch.optim.Optimizer):          # updates the way HF AdamW does
        def __init__(self, params):
            super().__init__(params, dict(lr=0.1))

        def step(self, closure=None):
            for g in self.param_groups:
                for p in g["params"]:
                    if p.grad is not None:
                        p.data.add_(p.grad.data, alpha=-g["lr"])
    w, frozen = torch.nn.Parameter(torch.ones(3)), torch.nn.Parameter(torch.ones(3))
    w.grad = torch.ones(3)
    opt = DataAdamLike([w])
    ver, e0, f0 = w._version, ops._pepoch.get(id(w), 0), ops._pepoch.get(id(frozen), 0)
    opt.step()
    assert w._version == ver, "the in-place .data update is invisible to the version counter (the reason for the hook)"
    assert ops._pepoch.get(id(w), 0) == e0 + 1 and ops._pepoch.get(id(frozen), 0) == f0
    torch.optim.SGD([w], lr=0.1).step()
    assert ops._pepoch.get(id(w), 0) == e0 + 2


def test_second_backward_fails_with_a_clear_message(monkeypatch):
    """the custom Functions free their layer buffers after the first backward; a second backward through the same
    graph must say so (RuntimeError naming retain_graph) instead of dying on a None unpack."""
    from efficientvlm_b200 import ops

    class Ctx:
        saved = None
        students = None
    with pytest.raises(RuntimeError, match="second time"):
        ops._saved_or_raise(Ctx())
    with pytest.raises(RuntimeError, match="retain_graph"):
        ops._saved_or_raise(Ctx(), "students")
    ctx = Ctx()
    ctx.saved = (1, 2)
    assert ops._saved_or_raise(ctx) == (1, 2)


def test_overlap_stage_boundaries_follow_the_arena_layout(monkeypatch):
    """FlatAdamW.enable_overlap: in every arena the vision tower is a prefix, and inside it the upper half of the tower (layers depth/2 ..
    and the final LayerNorm) is the contiguous run that ends the prefix — the two ranges that leave under the backward — while what is
    exchanged last ([0, split2)) holds the embeddings and the lower layers only."""
    from efficientvlm_b200.optim import create_optimizer
    from tests.helpers import gd_models
    ref_ops.install(monkeypatch)
    student, _ = gd_models(load_golden("gd_kd_tiny"))
    opt = create_optimizer(dict(lr=1e-4, weight_decay=0.01, lr_mult=2), student)
    opt.enable_overlap(student, "vision_encoder.")
    depth = len(student.vision_encoder.encoder.layers)
    mid = depth // 2
    upper = tuple("vision_encoder.encoder.layers.%d." % i for i in range(mid, depth)) + ("vision_encoder.post_layernorm.",)
    seen_upper = 0
    for g in opt.param_groups:
        assert 0 <= g["split2"] <= g["split"] <= g["size"]
        for name, off in zip(g["names"], g["offsets"]):
            if off >= g["split"]:
                assert not name.startswith("vision_encoder."), name
            elif off >= g["split2"]:
                assert name.startswith(upper), name
                seen_upper += 1
            else:
                assert name.startswith("vision_encoder.") and not name.startswith(upper), name
    assert seen_upper == sum(1 for n, _ in student.named_parameters() if n.startswith(upper))
    assert student.vision_encoder.encoder._evlm_grad_mid[0] == mid and student.vision_encoder._evlm_grad_ready


def test_kv_cache_registry_recognises_only_live_views_of_its_buffers():
    """ops._kv_cache_of: the decode loop's `past` tensors are accepted as an in-place cache only when they are the K and V views
    ([B, heads, Lp, 64], strides of the [B, capacity, K | V] buffer) of a LIVE registered buffer; re-ordered copies (beam search),
    other tensors at a recycled address and wrong geometries fall back to the concatenating path (None)."""
    import weakref
    from efficientvlm_b200 import ops
    B, nh, cap, Lp = 3, 2, 10, 4
    E = nh * 64
    cache = torch.zeros(B, cap, 2 * E, dtype=torch.bfloat16)
    ops._kv_caches[cache.data_ptr()] = weakref.ref(cache)
    pk = cache[:, :Lp, :E].view(B, Lp, nh, 64).permute(0, 2, 1, 3)
    pv = cache[:, :Lp, E:].view(B, Lp, nh, 64).permute(0, 2, 1, 3)
    assert ops._kv_cache_of(pk, pv, B, nh, E) is cache
    assert ops._kv_cache_of(pk, pk, B, nh, E) is None                          # V view does not start at the V half
    idx = torch.tensor([2, 0, 1])
    assert ops._kv_cache_of(pk.index_select(0, idx), pv.index_select(0, idx), B, nh, E) is None    # re-ordered copies (beam search)
    assert ops._kv_cache_of(pk[:2], pv[:2], 2, nh, E) is None                  # another batch size
    assert ops._kv_cache_of(pk.float(), pv.float(), B, nh, E) is None          # another dtype (and address)
    ptr = cache.data_ptr()
    del cache, pk, pv
    other = torch.zeros(B, nh, Lp, 64, dtype=torch.bfloat16)
    ref = ops._kv_caches.get(ptr)
    assert ref is None or ref() is None                                        # the registry holds no strong reference
    assert ops._kv_cache_of(other, other, B, nh, E) is None


def test_ctypes_structs_match_the_c_header_layout(tmp_path):
    """The ctypes mirrors in efficientvlm_b200/_lib.py against include/evlm.h as gcc lays it out: size of every struct and the offset of
    every field (a field added to one side only — an ABI drift — would shift everything behind it silently)."""
    import ctypes
    import re
    import shutil
    import subprocess
    from efficientvlm_b200 import _lib
    if shutil.which("gcc") is None:
        pytest.skip("no C compiler")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "evlm.h")).read()
    pairs = {"evlm_gemm_args": _lib.GemmArgs, "evlm_attn_args": _lib.AttnArgs, "evlm_mse_pair": _lib.MsePair,
             "evlm_cast_entry": _lib.CastEntry, "evlm_adamw_group": _lib.AdamWGroup}
    lines = ['#include <stdio.h>', '#include <stddef.h>', '#include "evlm.h"', "int main(void) {"]
    c_fields = {}
    for cname in pairs:
        body = re.search(r"typedef struct %s \{(.*?)\} %s;" % (cname, cname), header, re.S).group(1)
        body = re.sub(r"/\*.*?\*/", "", body, flags=re.S)
        names = []
        for decl in body.split(";"):
            decl = decl.strip()
            if not decl:
                continue
            for part in decl.split(","):
                names.append(re.findall(r"[A-Za-z_][A-Za-z_0-9]*", part)[-1])
        c_fields[cname] = names
        lines.append('  printf("%s %%zu\\n", sizeof(%s));' % (cname, cname))
        for n in names:
            lines.append('  printf("%s.%s %%zu\\n", offsetof(%s, %s));' % (cname, n, cname, n))
    lines += ["  return 0;", "}"]
    src = tmp_path / "layout.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "layout"
    subprocess.run(["gcc", "-I", os.path.join(root, "include"), str(src), "-o", str(exe)], check=True)
    out = dict(line.split() for line in subprocess.run([str(exe)], capture_output=True, text=True, check=True).stdout.splitlines())
    for cname, cls in pairs.items():
        assert int(out[cname]) == ctypes.sizeof(cls), (cname, out[cname], ctypes.sizeof(cls))
        py_fields = [f[0] for f in cls._fields_]
        assert [f for f in py_fields if f != "pad"] == [f for f in c_fields[cname] if f != "pad"], (cname, py_fields, c_fields[cname])
        for f in py_fields:
            if f in c_fields[cname]:
                assert int(out["%s.%s" % (cname, f)]) == getattr(cls, f).offset, (cname, f)
