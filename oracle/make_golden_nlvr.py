"""TEST INFRASTRUCTURE ONLY — generates tests/golden/nlvr_kd_tiny.pt: one NLVR2 pruning step's loss (Eff_NLVR.py:93-157) computed by
the UNMODIFIED reference classes — student `efficient_models/model_nlvr.py::EffXVLMForNLVR`, teacher
`models/model_nlvr.py::XVLMForNLVR` — and the reference's own KD helpers (extracted from Eff_NLVR.py with `ast`).

    python oracle/make_golden_nlvr.py
"""
import ast
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle.det_init import det_init_module_  # noqa: E402
from oracle.make_golden import BERT, VIS, cpu, save, spec  # noqa: E402

TEACHER_VIS = dict(VIS, num_hidden_layers=4, local_attn_depth=0)


def main():
    ref_shim.install()
    g = torch.Generator().manual_seed(31)
    vj, td = ref_shim.make_config_dir(dict(VIS, local_attn_depth=0), BERT)
    tvj, _ = ref_shim.make_config_dir(TEACHER_VIS, BERT)
    scfg = dict(text_encoder=td, vision_config=vj, patch_size=16, image_res=32, use_clip_vit=True, use_swin=False,
                text_num_hidden_layers=6, embed_dim=64, sparsity=0.25)
    tcfg = dict(scfg, vision_config=tvj, text_num_hidden_layers=12)
    cwd = os.getcwd()
    os.chdir(ref_shim.REF_ROOT)
    from efficient_models.model_nlvr import EffXVLMForNLVR
    from models.model_nlvr import XVLMForNLVR
    torch.manual_seed(9)
    student = EffXVLMForNLVR(scfg).eval()
    # Quirk Q12: as shipped, `models/xvlm.py::build_text_encoder` (:198-200) overwrites `config_text.num_hidden_layers` with
    # config['text_num_hidden_layers'] even when the caller passed a config_text, so `models/model_nlvr.py::XVLMForNLVR` builds a
    # 12-layer encoder and its own `share_cross_attention` raises IndexError (layer 12 of the 6 + 2*6 = 18 it expects).  The
    # efficient_models builder (`efficient_models/xvlm.py:142-155`) honours config_text; it is swapped in to make the reference
    # teacher constructible — its forward is the unmodified reference code.
    import efficient_models.xvlm as exvlm
    import models.xvlm as mxvlm
    mxvlm.build_text_encoder = exvlm.build_text_encoder
    teacher = XVLMForNLVR(tcfg).eval()
    os.chdir(cwd)
    det_init_module_(student)
    det_init_module_(teacher)
    with torch.no_grad():
        for k, la in student.l0_module.z_logas.items():
            la.copy_(torch.randn(la.shape, generator=g) * 1.5 + 1.0)
        student.l0_module.lambda_1.fill_(0.5)
        student.l0_module.lambda_2.fill_(0.25)
    student.l0_module.set_lagrangian_warmup_steps(30)
    B = 3
    image = torch.randn(2 * B, 3, 32, 32, generator=g)          # B first images, then B second images
    text_ids = torch.randint(1, BERT["vocab_size"], (B, 8), generator=g)
    text_atts = torch.ones(B, 8, dtype=torch.long)
    text_atts[1, 6:] = 0
    targets = torch.tensor([1, 0, 1])
    eps = {k: torch.rand(la.shape, generator=g).clamp(1e-6, 1 - 1e-6) for k, la in student.l0_module.z_logas.items()}
    it = iter([eps[k] for k in student.l0_module.types])
    student.l0_module.get_eps = lambda size: next(it)
    so = student(image, text_ids, text_atts, targets=targets, train=True, output_attentions=True, output_hidden_states=True)
    with torch.no_grad():
        to = teacher(image, text_ids, text_atts, targets=targets, train=True, output_attentions=True, output_hidden_states=True)
    src = open(os.path.join(ref_shim.REF_ROOT, "Eff_NLVR.py")).read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("get_kd_loss", "soft_cross_entropy", "get_cor_teacher")]
    ns = {"torch": torch, "KLDivLoss": torch.nn.KLDivLoss}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "Eff_NLVR.py", "exec"), ns)
    get_kd_loss, soft_cross_entropy, get_cor_teacher = ns["get_kd_loss"], ns["soft_cross_entropy"], ns["get_cor_teacher"]
    mse, dev, temperature = torch.nn.MSELoss(), "cpu", 1.0
    sh, th, sa, ta = so["hidden_dict"], to["hidden_dict"], so["attention_dict"], to["attention_dict"]
    sc, tc = so["cross_attention_dict"], to["cross_attention_dict"]
    # ---- Eff_NLVR.py:110-155, statement by statement ----
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
    logits_loss = soft_cross_entropy(so["logits_dict"]["cls_head_logits"] / temperature, to["logits_dict"]["cls_head_logits"] / temperature)
    loss_small = so["loss"]
    loss_text_kd = text_attention_loss + text_hidden_loss
    loss_img_kd = image_attention_loss + image_hidden_loss * 0.1
    loss_cross_kd = (cross_hidden_loss + cross_self_attention_loss + cross_attention_loss) * 0.5
    loss_kd = logits_loss + loss_text_kd + (loss_img_kd + loss_cross_kd) * 0.33
    loss = 0.8 * loss_small + 0.2 * loss_kd
    lagrangian_loss, _, _ = student.l0_module.lagrangian_regularization(12)
    loss = loss + lagrangian_loss
    gn = ["vision_encoder.encoder.layers.1.self_attn.out_proj.weight", "text_encoder.encoder.layer.0.intermediate.dense.weight",
          "text_encoder.encoder.layer.3.crossattention.self.key.weight", "text_encoder.encoder.layer.7.crossattention.self.query.weight",
          "text_encoder.encoder.layer.8.output.dense.weight", "cls_head.0.weight", "l0_module.cross_head_loga", "l0_module.vision_int_loga",
          "l0_module.lambda_2"]
    sp = dict(student.named_parameters())
    grads = torch.autograd.grad(loss, [sp[n] for n in gn])
    with torch.no_grad():
        pred_eval = student(image, text_ids, text_atts, targets=targets, train=False)
    save("nlvr_kd_tiny", dict(
        scfg=dict(scfg, text_encoder=None, vision_config=None), tcfg=dict(tcfg, text_encoder=None, vision_config=None), vis=dict(VIS, local_attn_depth=0),
        tvis=TEACHER_VIS, bert=BERT, s_sd_spec=spec(student), t_sd_spec=spec(teacher), s_param_names=[n for n, _ in student.named_parameters()],
        l0_logas={k: cpu(v) for k, v in student.l0_module.z_logas.items()}, lambda_1=0.5, lambda_2=0.25, warmup=30, step=12, eps=eps,
        image=image, text_ids=text_ids, text_atts=text_atts, targets=targets, total=cpu(loss),
        counts=dict(s_text_h=len(s_text_h), s_text_a=len(s_text_a), s_cross_a=len(s_cross_a), t_text_h=len(th["text_hidden_states"]),
                    t_text_a=len(ta["text_attentions"]), t_cross_a=len(tc["cross_attentions"])),
        parts=dict(text_hidden=cpu(text_hidden_loss), text_attention=cpu(text_attention_loss), cross_hidden=cpu(cross_hidden_loss),
                   cross_self_attention=cpu(cross_self_attention_loss), cross_attention=cpu(cross_attention_loss),
                   image_hidden=cpu(image_hidden_loss), image_attention=cpu(image_attention_loss), logits=cpu(logits_loss),
                   loss_small=cpu(loss_small), lagrangian=cpu(lagrangian_loss)),
        s_logits=cpu(so["logits_dict"]["cls_head_logits"]), t_logits=cpu(to["logits_dict"]["cls_head_logits"]),
        s_cross_last=cpu(s_cross_a[-1]), pred_eval=cpu(pred_eval), grad_names=gn, grads=cpu(grads)))
    print("done")


if __name__ == "__main__":
    main()
