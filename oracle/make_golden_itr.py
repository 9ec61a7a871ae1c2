"""TEST INFRASTRUCTURE ONLY — generates tests/golden/itr_kd_tiny.pt: one ITR pruning step's loss (Eff_Retrieval.py:96-178) computed
by the UNMODIFIED reference classes — student `efficient_models/model_retrieval.py::EffXVLMforRetrieval`, teacher
`models/model_retrieval.py::XVLM` — and the reference's own KD helpers (extracted from Eff_Retrieval.py with `ast`).

    python oracle/make_golden_itr.py
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
    g = torch.Generator().manual_seed(21)
    vj, td = ref_shim.make_config_dir(dict(VIS, local_attn_depth=0), BERT)
    tvj, _ = ref_shim.make_config_dir(TEACHER_VIS, BERT)
    scfg = dict(text_encoder=td, vision_config=vj, patch_size=16, image_res=32, use_clip_vit=True, use_swin=False,
                text_num_hidden_layers=6, embed_dim=64, temp=0.07, sparsity=0.25)
    tcfg = dict(scfg, vision_config=tvj, text_num_hidden_layers=12)
    cwd = os.getcwd()
    os.chdir(ref_shim.REF_ROOT)
    from efficient_models.model_retrieval import EffXVLMforRetrieval
    from models.model_retrieval import XVLM as TeacherXVLM
    torch.manual_seed(8)
    student = EffXVLMforRetrieval(scfg).eval()
    teacher = TeacherXVLM(tcfg).eval()
    os.chdir(cwd)
    det_init_module_(student)
    det_init_module_(teacher)
    with torch.no_grad():
        for k, la in student.l0_module.z_logas.items():
            la.copy_(torch.randn(la.shape, generator=g) * 1.5 + 1.0)
        student.l0_module.lambda_1.fill_(-0.3)
        student.l0_module.lambda_2.fill_(0.6)
    student.l0_module.set_lagrangian_warmup_steps(40)
    B = 4
    image = torch.randn(B, 3, 32, 32, generator=g)
    text_ids = torch.randint(1, BERT["vocab_size"], (B, 9), generator=g)
    text_atts = torch.ones(B, 9, dtype=torch.long)
    text_atts[2, 6:] = 0
    idx = torch.tensor([3, 4, 5, 4])
    eps = {k: torch.rand(la.shape, generator=g).clamp(1e-6, 1 - 1e-6) for k, la in student.l0_module.z_logas.items()}
    it = iter([eps[k] for k in student.l0_module.types])
    student.l0_module.get_eps = lambda size: next(it)
    orig_multinomial = torch.multinomial
    torch.multinomial = lambda w, n, *a, **k: torch.argmax(w, dim=-1, keepdim=True)   # deterministic hard negatives on both sides
    try:
        so = student(image, text_ids, text_atts, idx=idx, output_attentions=True, output_hidden_states=True)
        with torch.no_grad():
            to = teacher(image, text_ids, text_atts, idx=idx, output_attentions=True, output_hidden_states=True)
    finally:
        torch.multinomial = orig_multinomial
    src = open(os.path.join(ref_shim.REF_ROOT, "Eff_Retrieval.py")).read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("get_kd_loss", "soft_cross_entropy", "get_cor_teacher")]
    ns = {"torch": torch, "KLDivLoss": torch.nn.KLDivLoss}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "Eff_Retrieval.py", "exec"), ns)
    get_kd_loss, soft_cross_entropy, get_cor_teacher = ns["get_kd_loss"], ns["soft_cross_entropy"], ns["get_cor_teacher"]
    mse, dev, temperature = torch.nn.MSELoss(), "cpu", 1.0
    sh, th, sa, ta = so["hidden_dict"], to["hidden_dict"], so["attention_dict"], to["attention_dict"]
    sc, tc = so["cross_attention_dict"], to["cross_attention_dict"]

    def hid(name, is_img=False):
        return get_kd_loss(sh[name], get_cor_teacher(th[name], sh[name]), False, mse, dev, is_img=is_img)

    def att(s, t):
        return get_kd_loss(s, get_cor_teacher(t, s, is_attn=True), True, mse, dev)
    # ---- Eff_Retrieval.py:112-178, statement by statement ----
    text_hidden_loss, text_attention_loss = hid("text_hidden_states"), att(sa["text_attentions"], ta["text_attentions"])
    image_hidden_loss, image_attention_loss = hid("image_hidden_states", True), att(sa["image_attentions"], ta["image_attentions"])
    itm_pos_hidden_loss, itm_pos_attn_loss = hid("itm_pos_hidden_states"), att(sa["itm_pos_attentions"], ta["itm_pos_attentions"])
    itm_pos_cross_loss = att(sc["itm_pos_cross_attentions"], tc["itm_pos_cross_attentions"])
    itm_neg_hidden_loss, itm_neg_attn_loss = hid("itm_neg_hidden_states"), att(sa["itm_neg_attentions"], ta["itm_neg_attentions"])
    itm_neg_cross_loss = att(sc["itm_neg_cross_attentions"], tc["itm_neg_cross_attentions"])
    itm_logits_loss = soft_cross_entropy(so["logits_dict"]["itm_head_logits"] / temperature, to["logits_dict"]["itm_head_logits"] / temperature)
    loss_itc, loss_itm = so["loss"]["loss_itc"], so["loss"]["loss_itm"]
    loss_text_kd = text_hidden_loss + text_attention_loss
    loss_img_kd = 0.2 * image_hidden_loss + image_attention_loss
    loss_cross_kd = (itm_neg_hidden_loss + itm_pos_hidden_loss + itm_pos_attn_loss + itm_pos_cross_loss + itm_neg_attn_loss + itm_neg_cross_loss) * 0.5
    loss_kd = itm_logits_loss + (loss_text_kd + loss_img_kd + loss_cross_kd) * 0.33
    loss_small = loss_itc + loss_itm
    loss = (loss_kd + loss_small) * 0.5
    lagrangian_loss, _, _ = student.l0_module.lagrangian_regularization(17)
    loss = loss + lagrangian_loss
    gn = ["vision_encoder.encoder.layers.0.self_attn.k_proj.weight", "text_encoder.encoder.layer.2.output.dense.weight",
          "text_encoder.encoder.layer.5.crossattention.self.query.weight", "itm_head.3.weight", "text_proj.weight", "temp",
          "l0_module.text_head_loga", "l0_module.cross_int_loga", "l0_module.lambda_1"]
    sp = dict(student.named_parameters())
    grads = torch.autograd.grad(loss, [sp[n] for n in gn])
    save("itr_kd_tiny", dict(
        scfg=dict(scfg, text_encoder=None, vision_config=None), tcfg=dict(tcfg, text_encoder=None, vision_config=None), vis=dict(VIS, local_attn_depth=0),
        tvis=TEACHER_VIS, bert=BERT, s_sd_spec=spec(student), t_sd_spec=spec(teacher),
        l0_logas={k: cpu(v) for k, v in student.l0_module.z_logas.items()}, lambda_1=-0.3, lambda_2=0.6, warmup=40, step=17, eps=eps,
        image=image, text_ids=text_ids, text_atts=text_atts, idx=idx, total=cpu(loss),
        parts=dict(text_hidden=cpu(text_hidden_loss), text_attention=cpu(text_attention_loss), image_hidden=cpu(image_hidden_loss),
                   image_attention=cpu(image_attention_loss), itm_pos_hidden=cpu(itm_pos_hidden_loss), itm_pos_attn=cpu(itm_pos_attn_loss),
                   itm_pos_cross=cpu(itm_pos_cross_loss), itm_neg_hidden=cpu(itm_neg_hidden_loss), itm_neg_attn=cpu(itm_neg_attn_loss),
                   itm_neg_cross=cpu(itm_neg_cross_loss), itm_logits=cpu(itm_logits_loss), loss_kd=cpu(loss_kd), loss_itc=cpu(loss_itc),
                   loss_itm=cpu(loss_itm), lagrangian=cpu(lagrangian_loss)),
        t_itm_logits=cpu(to["logits_dict"]["itm_head_logits"]), t_neg_cross_last=cpu(tc["itm_neg_cross_attentions"][-1]),
        grad_names=gn, grads=cpu(grads)))
    print("done")


if __name__ == "__main__":
    main()
