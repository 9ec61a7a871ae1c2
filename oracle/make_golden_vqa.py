"""TEST INFRASTRUCTURE ONLY — generates tests/golden/vqa_tiny.pt by running the UNMODIFIED reference VQA task models
(`efficient_models/model_generation.py::EffXVLMForVQA` student, `models/model_generation.py::XVLMForVQA` teacher) and the
reference's own KD helpers (`Eff_VQA.py:28-71`, extracted from the source with `ast`) through oracle/ref_shim.py on seeded
synthetic inputs with tiny random-init models.

    python oracle/make_golden_vqa.py        # rewrites tests/golden/vqa_tiny.pt only

Weights are rebuilt on both sides by oracle/det_init.py, so only inputs and outputs are stored.
"""
import ast
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle.det_init import det_init_module_  # noqa: E402
from oracle.make_golden import BERT, VIS, cpu, save, spec  # noqa: E402

TEACHER_VIS = dict(VIS, num_hidden_layers=4, local_attn_depth=0)


def main():
    ref_shim.install()
    # `efficient_models/model_generation.py:7` imports `dataset.build_tokenizer` (only the captioning model calls it); the real
    # package drags in skimage / pycocotools, so a stub stands in for it.
    ds = types.ModuleType("dataset")
    ds.build_tokenizer = lambda *a, **k: None
    sys.modules["dataset"] = ds
    g = torch.Generator().manual_seed(11)
    vj, td = ref_shim.make_config_dir(dict(VIS, local_attn_depth=0), BERT)
    tvj, _ = ref_shim.make_config_dir(TEACHER_VIS, BERT)
    scfg = dict(text_encoder=td, vision_config=vj, patch_size=16, image_res=32, use_clip_vit=True, use_swin=False,
                text_num_hidden_layers=6, num_dec_layers=3, pad_token_id=0, sparsity=0.35)
    tcfg = dict(scfg, vision_config=tvj, text_num_hidden_layers=12, num_dec_layers=6)
    cwd = os.getcwd()
    os.chdir(ref_shim.REF_ROOT)
    from efficient_models.model_generation import EffXVLMForVQA
    from models.model_generation import XVLMForVQA
    torch.manual_seed(5)
    student = EffXVLMForVQA(scfg).eval()
    teacher = XVLMForVQA(tcfg).eval()
    os.chdir(cwd)
    det_init_module_(student)
    det_init_module_(teacher)
    for m in (student, teacher):   # BertLMHeadModel ties decoder <-> word embeddings (eff_bert.py:1621)
        m.text_decoder.cls.predictions.decoder.weight = m.text_decoder.bert.embeddings.word_embeddings.weight
    with torch.no_grad():
        for k, la in student.l0_module.z_logas.items():
            la.copy_(torch.randn(la.shape, generator=g) * 1.5 + 0.5)
        student.l0_module.lambda_1.fill_(0.4)
        student.l0_module.lambda_2.fill_(-0.2)
    student.l0_module.set_lagrangian_warmup_steps(50)

    B, Lq, La = 3, 7, 4
    image = torch.randn(B, 3, 32, 32, generator=g)
    q_ids = torch.randint(1, BERT["vocab_size"], (B, Lq), generator=g)
    q_atts = torch.ones(B, Lq, dtype=torch.long)
    q_atts[1, 5:] = 0
    q_ids[1, 5:] = 0
    question = types.SimpleNamespace(input_ids=q_ids, attention_mask=q_atts)
    k = [2, 1, 3]
    n_ans = sum(k)
    a_ids = torch.randint(1, BERT["vocab_size"], (n_ans, La), generator=g)
    a_ids[:, 0] = 5
    a_atts = torch.ones(n_ans, La, dtype=torch.long)
    a_atts[2, 3:] = 0
    a_ids[2, 3:] = 0
    a_atts[4, 2:] = 0
    a_ids[4, 2:] = 0
    answer = types.SimpleNamespace(input_ids=a_ids, attention_mask=a_atts)
    weights = torch.tensor([0.5, 0.5, 1.0, 0.2, 0.3, 0.5])

    eps = {kk: torch.rand(la.shape, generator=g).clamp(1e-6, 1 - 1e-6) for kk, la in student.l0_module.z_logas.items()}

    def arm_eps():
        it = iter([eps[t] for t in student.l0_module.types])
        student.l0_module.get_eps = lambda size: next(it)

    # ---- train, KD outputs (Eff_VQA.py:102-103) ----
    arm_eps()
    so = student(image, question, answer, train=True, k=k, weights=weights, output_attentions=True, output_hidden_states=True)
    with torch.no_grad():
        to = teacher(image, question, answer, train=True, k=k, weights=weights, output_attentions=True, output_hidden_states=True)
    src = open(os.path.join(ref_shim.REF_ROOT, "Eff_VQA.py")).read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("get_kd_loss", "soft_cross_entropy", "get_cor_teacher")]
    ns = {"torch": torch, "KLDivLoss": torch.nn.KLDivLoss}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "Eff_VQA.py", "exec"), ns)   # the reference's own helpers, unmodified
    get_kd_loss, soft_cross_entropy, get_cor_teacher = ns["get_kd_loss"], ns["soft_cross_entropy"], ns["get_cor_teacher"]
    mse, dev, temperature = torch.nn.MSELoss(), "cpu", 1.0
    # ---- the loss assembly of Eff_VQA.py:105-176, statement by statement ----
    sh, th, sa, ta = so["hidden_dict"], to["hidden_dict"], so["attention_dict"], to["attention_dict"]
    sc, tc = so["cross_attention_dict"], to["cross_attention_dict"]
    s_text_h = sh["text_hidden_states"]
    t_text_h = get_cor_teacher(th["text_hidden_states"], s_text_h)
    s_cross_h, t_cross_h = s_text_h[4:], t_text_h[4:]
    s_text_a = sa["text_attentions"]
    t_text_a = get_cor_teacher(ta["text_attentions"], s_text_a, is_attn=True)
    s_cross_sa, t_cross_sa = s_text_a[3:], t_text_a[3:]
    s_cross_a = sc["cross_attentions"]
    t_cross_a = get_cor_teacher(tc["cross_attentions"], s_cross_a, is_attn=True)
    text_hidden_loss = get_kd_loss(s_text_h[:4], t_text_h[:4], False, mse, dev)
    text_attention_loss = get_kd_loss(s_text_a[:3], t_text_a[:3], True, mse, dev)
    cross_hidden_loss = get_kd_loss(s_cross_h, t_cross_h, False, mse, dev)
    cross_self_attention_loss = get_kd_loss(s_cross_sa, t_cross_sa, True, mse, dev)
    cross_attention_loss = get_kd_loss(s_cross_a, t_cross_a, True, mse, dev)
    s_img_h = sh["image_hidden_states"]
    t_img_h = get_cor_teacher(th["image_hidden_states"], s_img_h)
    s_img_a = sa["image_attentions"]
    t_img_a = get_cor_teacher(ta["image_attentions"], s_img_a, is_attn=True)
    image_hidden_loss = get_kd_loss(s_img_h, t_img_h, False, mse, dev, is_img=True)
    image_attention_loss = get_kd_loss(s_img_a, t_img_a, True, mse, dev)
    s_dec_h = sh["decoder_hidden_states"]
    t_dec_h = get_cor_teacher(th["decoder_hidden_states"], s_dec_h)
    s_dec_a = sa["decoder_attentions"]
    t_dec_a = get_cor_teacher(ta["decoder_attentions"], s_dec_a, is_attn=True)
    s_dec_c = sc["decoder_cross_attentions"]
    t_dec_c = get_cor_teacher(tc["decoder_cross_attentions"], s_dec_c, is_attn=True)
    decoder_hidden_loss = get_kd_loss(s_dec_h, t_dec_h, False, mse, dev, is_img=True)
    decoder_attention_loss = get_kd_loss(s_dec_a, t_dec_a, True, mse, dev)
    decoder_cross_loss = get_kd_loss(s_dec_c, t_dec_c, True, mse, dev)
    logits_loss = soft_cross_entropy(so["logits_dict"]["logits"] / temperature, to["logits_dict"]["logits"] / temperature)
    loss_small = so["loss"]
    loss_text_kd = text_attention_loss + text_hidden_loss
    loss_img_kd = image_attention_loss + image_hidden_loss * 0.2
    loss_cross_kd = (cross_hidden_loss + cross_self_attention_loss + cross_attention_loss) * 0.5
    loss_decoder_kd = decoder_attention_loss + decoder_hidden_loss + decoder_cross_loss
    loss_kd = logits_loss + loss_text_kd + loss_img_kd + loss_cross_kd + loss_decoder_kd
    loss = loss_kd * 0.4 + loss_small * 0.6
    lagrangian_loss, exp_sparsity, tgt_sparsity = student.l0_module.lagrangian_regularization(20)
    loss = loss + lagrangian_loss
    gn = ["vision_encoder.encoder.layers.1.mlp.fc1.weight", "text_encoder.encoder.layer.1.attention.self.query.weight",
          "text_encoder.encoder.layer.4.crossattention.self.key.weight", "text_decoder.bert.encoder.layer.2.crossattention.self.value.weight",
          "text_decoder.bert.encoder.layer.0.intermediate.dense.weight", "text_decoder.cls.predictions.transform.dense.weight",
          "text_decoder.bert.embeddings.word_embeddings.weight", "l0_module.vision_head_loga", "l0_module.cross_head_loga",
          "l0_module.decoder_head_loga", "l0_module.decoder_int_loga", "l0_module.lambda_1", "l0_module.lambda_2"]
    sp = dict(student.named_parameters())
    grads = torch.autograd.grad(loss, [sp[n] for n in gn])

    # ---- train, task loss only (Eff_VQA fine-tune without KD outputs) and the stop_prune branch ----
    arm_eps()
    loss_plain = student(image, question, answer, train=True, k=k, weights=weights)
    loss_stop = student(image, question, answer, train=True, k=k, weights=weights, stop_prune=True)

    # ---- eval: deterministic masks + rank_answer over a 9-entry answer list, k_test = 4 ----
    n_list, k_test = 9, 4
    l_ids = torch.randint(1, BERT["vocab_size"], (n_list, La), generator=g)
    l_ids[:, 0] = 5
    l_ids[:, 1] = torch.randperm(BERT["vocab_size"] - 1, generator=g)[:n_list] + 1   # distinct first tokens
    l_atts = torch.ones(n_list, La, dtype=torch.long)
    l_atts[3, 3:] = 0
    l_ids[3, 3:] = 0
    l_atts[7, 2:] = 0
    l_ids[7, 2:] = 0
    alist = types.SimpleNamespace(input_ids=l_ids, attention_mask=l_atts)
    with torch.no_grad():
        topk_ids, topk_probs = student(image, question, alist, train=False, k=k_test)
        t_topk_ids, t_topk_probs = teacher(image, question, alist, train=False, k=k_test)
        zs_eval = student.l0_module.forward(training=False)

    save("vqa_tiny", dict(
        scfg=dict(scfg, text_encoder=None, vision_config=None), tcfg=dict(tcfg, text_encoder=None, vision_config=None), vis=dict(VIS, local_attn_depth=0),
        tvis=TEACHER_VIS, bert=BERT, s_sd_spec=spec(student), t_sd_spec=spec(teacher),
        l0_logas={kk: cpu(v) for kk, v in student.l0_module.z_logas.items()}, lambda_1=0.4, lambda_2=-0.2, warmup=50, step=20, eps=eps,
        image=image, q_ids=q_ids, q_atts=q_atts, a_ids=a_ids, a_atts=a_atts, k=k, weights=weights,
        s_loss=cpu(so["loss"]), t_loss=cpu(to["loss"]), s_logits=cpu(so["logits_dict"]["logits"]), t_logits=cpu(to["logits_dict"]["logits"]),
        s_image_hidden_last=cpu(s_img_h[-1]), s_text_hidden_last=cpu(s_text_h[-1]), s_decoder_hidden=cpu(s_dec_h),
        s_decoder_attn=cpu(s_dec_a), s_decoder_cross=cpu(s_dec_c), s_cross_attn=cpu(s_cross_a), t_decoder_cross_last=cpu(to["cross_attention_dict"]["decoder_cross_attentions"][-1]),
        counts=dict(s_img_h=len(s_img_h), s_text_h=len(s_text_h), s_text_a=len(s_text_a), s_cross_a=len(s_cross_a), s_dec_h=len(s_dec_h),
                    s_dec_a=len(s_dec_a), s_dec_c=len(s_dec_c), t_img_h=len(th["image_hidden_states"]), t_text_h=len(th["text_hidden_states"]),
                    t_text_a=len(ta["text_attentions"]), t_cross_a=len(tc["cross_attentions"]), t_dec_h=len(th["decoder_hidden_states"]),
                    t_dec_a=len(ta["decoder_attentions"]), t_dec_c=len(tc["decoder_cross_attentions"])),
        parts=dict(text_hidden=cpu(text_hidden_loss), text_attention=cpu(text_attention_loss), cross_hidden=cpu(cross_hidden_loss),
                   cross_self_attention=cpu(cross_self_attention_loss), cross_attention=cpu(cross_attention_loss),
                   image_hidden=cpu(image_hidden_loss), image_attention=cpu(image_attention_loss), decoder_hidden=cpu(decoder_hidden_loss),
                   decoder_attention=cpu(decoder_attention_loss), decoder_cross=cpu(decoder_cross_loss), logits=cpu(logits_loss),
                   loss_kd=cpu(loss_kd), lagrangian=cpu(lagrangian_loss), expected_sparsity=cpu(exp_sparsity), target_sparsity=tgt_sparsity),
        total=cpu(loss), grad_names=gn, grads=cpu(grads), loss_plain=cpu(loss_plain), loss_stop=cpu(loss_stop),
        l_ids=l_ids, l_atts=l_atts, k_test=k_test, topk_ids=cpu(topk_ids), topk_probs=cpu(topk_probs), t_topk_ids=cpu(t_topk_ids),
        t_topk_probs=cpu(t_topk_probs), zs_eval={kk: cpu(v) for kk, v in zs_eval.items()}))
    make_pruned(g, scfg, question, alist, image, k_test)
    print("done")


def make_pruned(g, scfg, question, alist, image, k_test):
    """Second fixture: the reference's OWN materialisation utilities (utils/vqa_utils.py:37-313, literal layer counts 6/3/3/3, so
    the vision tower gets 6 layers here) applied to the reference student, then `fake_forward` (no gates) on the pruned model."""
    import efficient_models.model_generation as mg
    import utils.vqa_utils as vu
    vis6 = dict(VIS, num_hidden_layers=6, local_attn_depth=0)
    vj6, td = ref_shim.make_config_dir(vis6, BERT)
    cfg6 = dict(scfg, vision_config=vj6, text_encoder=td)
    cwd = os.getcwd()
    os.chdir(ref_shim.REF_ROOT)
    torch.manual_seed(6)
    m = mg.EffXVLMForVQA(cfg6).eval()
    os.chdir(cwd)
    det_init_module_(m)
    m.text_decoder.cls.predictions.decoder.weight = m.text_decoder.bert.embeddings.word_embeddings.weight
    sd_spec = spec(m)
    heads, inter = BERT["num_attention_heads"], BERT["intermediate_size"]

    def head_gate(layers, pattern):     # one head at most is pruned per layer (2-head fixture); kept gates are non-binary
        z = torch.rand(layers, 1, heads, 1, 1, generator=g) * 0.8 + 0.2
        for layer, hd in pattern.items():
            z[layer, 0, hd, 0, 0] = 0.0
        return z

    def int_gate(layers):
        z = torch.rand(layers, 1, 1, inter, generator=g)
        z[z < 0.3] = 0.0
        z[z > 0.8] = 1.0
        return z
    zs = {"vision_head_z": head_gate(6, {0: 0, 2: 1, 5: 0}), "text_head_z": head_gate(3, {1: 1}), "cross_head_z": head_gate(6, {0: 0, 3: 1, 4: 0}),
          "decoder_head_z": head_gate(6, {1: 1, 2: 0}), "vision_intermediate_z": int_gate(6), "text_intermediate_z": int_gate(3),
          "cross_intermediate_z": int_gate(3), "decoder_intermediate_z": int_gate(3)}
    vu.update_params(m, zs)
    vu.prune_model_with_z(zs, m)
    with torch.no_grad():
        ids, probs, _ = m.fake_forward(image, question, alist, k=k_test)
    save("vqa_pruned_tiny", dict(vis=vis6, sd_spec=sd_spec, zs=zs, pruned_shapes={k: tuple(v.shape) for k, v in m.state_dict().items()},
                                 topk_ids=cpu(ids), topk_probs=cpu(probs),
                                 probe={k: cpu(v) for k, v in m.state_dict().items() if k in (
                                     "vision_encoder.encoder.layers.0.self_attn.v_proj.weight", "vision_encoder.encoder.layers.2.mlp.fc2.weight",
                                     "text_encoder.encoder.layer.4.crossattention.self.value.bias", "text_decoder.bert.encoder.layer.1.output.dense.weight")}))


if __name__ == "__main__":
    main()
