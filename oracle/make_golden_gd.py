"""TEST INFRASTRUCTURE ONLY — generates tests/golden/gd_kd_tiny.pt: one general-distillation step (BASELINE config 2, the headline
workload) computed by the UNMODIFIED reference: student and teacher are `models/model_pretrain.py::XVLM` (ViT-6 + BERT-3/3 vs ViT-12 +
BERT-6/6 at width 128), and the loss is the reference's OWN train-loop code — the statements of `GeneralDistill.py:300-376` (from
`student_hidden = student_outputs['hidden_dict']` to `loss_in_total = ...`) are lifted out of `train()` with `ast` and executed as they
stand, together with the file's `get_kd_loss` / `get_cor_teacher` / `soft_cross_entropy`.

A second fixture, tests/golden/gd_region_tiny.pt, is the region-batch half of the iteration (`GeneralDistill.py:158-260`: `ret_bbox_loss=True`,
per-region image replication in the local ViT layers, bbox head, L1 + GIoU), lifted the same way.

    python oracle/make_golden_gd.py
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

# `models/model_pretrain.py::XVLM` always loads "pretrained" towers (load_vision_params / load_text_params = True) and only knows
# 6- and 12-layer towers (`models/xvlm.py:123-126,198-200`): student ViT-6 + BERT-3/3, teacher ViT-12 + BERT-6/6, width 128, and
# stand-in checkpoint files (a CLIP file holding only the position embedding, an empty bert-base-uncased file); every weight is then
# overwritten by the deterministic initialiser, exactly as in the other fixtures.
STUDENT_VIS = dict(VIS, num_hidden_layers=6, local_attn_depth=0)
TEACHER_VIS = dict(VIS, num_hidden_layers=12, local_attn_depth=0)


def config_dirs(vis):
    """(vision json with a stand-in 'ckpt', text dir whose path contains 'bert-base-uncased' with an empty pytorch_model.bin)."""
    vj, td = ref_shim.make_config_dir(vis, BERT)
    base = os.path.dirname(td)
    ckpt = os.path.join(base, "clip_stub.bin")
    n = (32 // 16) ** 2 + 1
    torch.save({"vision_model.embeddings.position_embedding.weight": torch.zeros(n, vis["vision_width"])}, ckpt)
    import json
    with open(vj, "w") as f:
        json.dump(dict(vis, ckpt=ckpt), f)
    bert_dir = os.path.join(base, "bert-base-uncased")
    os.symlink(td, bert_dir)
    torch.save({}, os.path.join(td, "pytorch_model.bin"))
    return vj, bert_dir


def reference_loss_code(region=False):
    """(helper namespace, code object of GeneralDistill.py's loss statements inside train()'s batch loop).  region=False: the main
    batch (lines 300-376); region=True: the region-batch branch `if random.random() < config['regions']['iter_perc']:` (lines 184-260)."""
    src = open(os.path.join(ref_shim.REF_ROOT, "GeneralDistill.py")).read()
    tree = ast.parse(src)
    helpers = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("get_kd_loss", "soft_cross_entropy", "get_cor_teacher")]
    ns = {"torch": torch, "KLDivLoss": torch.nn.KLDivLoss, "MSELoss": torch.nn.MSELoss}
    exec(compile(ast.Module(body=helpers, type_ignores=[]), "GeneralDistill.py", "exec"), ns)
    train = next(n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "train")

    def assigns(s, name):
        return isinstance(s, ast.Assign) and getattr(s.targets[0], "id", None) == name
    kind = ast.If if region else ast.For
    block = next(n for n in ast.walk(train) if isinstance(n, kind) and any(assigns(s, "loss_in_total") for s in n.body))
    first = next(i for i, s in enumerate(block.body) if assigns(s, "student_hidden"))
    last = next(i for i, s in enumerate(block.body) if assigns(s, "loss_in_total"))
    stmts = block.body[first:last + 1]
    print("lifted GeneralDistill.py lines %d-%d (%d statements)" % (stmts[0].lineno, stmts[-1].end_lineno, len(stmts)))
    return ns, compile(ast.Module(body=stmts, type_ignores=[]), "GeneralDistill.py", "exec")


GRAD_NAMES = ["vision_encoder.encoder.layers.0.self_attn.q_proj.weight", "vision_encoder.encoder.layers.1.mlp.fc2.weight",
              "vision_encoder.patch_embed.weight", "text_encoder.bert.embeddings.word_embeddings.weight",
              "text_encoder.bert.encoder.layer.1.attention.self.value.weight", "text_encoder.bert.encoder.layer.4.crossattention.self.key.weight",
              "text_encoder.bert.encoder.layer.5.output.dense.weight", "text_encoder.cls.predictions.transform.dense.weight",
              "itm_head.0.weight", "itm_head.3.bias", "vision_proj.weight", "text_proj.bias", "temp"]
PART_NAMES = ["text_hidden_loss", "text_attention_loss", "image_hidden_loss", "image_attention_loss", "itm_pos_hidden_loss",
              "itm_pos_attn_loss", "itm_neg_hidden_loss", "itm_neg_attn_loss", "mlm_hidden_loss", "mlm_attn_loss", "mlm_logits_loss",
              "itm_logits_loss", "loss_small", "loss_text_kd", "loss_img_kd", "loss_cross_kd", "loss_kd"]


def build_pair(student_vis, teacher_vis):
    vj, td = config_dirs(student_vis)
    tvj, _ = config_dirs(teacher_vis)
    scfg = dict(text_encoder=td, vision_config=vj, patch_size=16, image_res=32, use_clip_vit=True, use_swin=False,
                text_num_hidden_layers=6, embed_dim=64, temp=0.07, max_tokens=9)
    tcfg = dict(scfg, vision_config=tvj, text_num_hidden_layers=12)
    cwd = os.getcwd()
    os.chdir(ref_shim.REF_ROOT)
    from models.model_pretrain import XVLM
    torch.manual_seed(3)
    student = XVLM(scfg).eval()          # eval(): BERT's 0.1 dropouts are Philox-stream dependent, compared with p = 0
    teacher = XVLM(tcfg).eval()
    os.chdir(cwd)
    det_init_module_(student)
    det_init_module_(teacher)
    for m in (student, teacher):
        m.text_encoder.cls.predictions.decoder.weight = m.text_encoder.bert.embeddings.word_embeddings.weight
    return student, teacher, scfg, tcfg


def text_batch(B, L, g, short_rows):
    text_ids = torch.randint(1, BERT["vocab_size"], (B, L), generator=g)
    text_atts = torch.ones(B, L, dtype=torch.long)
    for row, n in short_rows:
        text_atts[row, n:] = 0
    masked_pos = torch.stack([torch.randperm(L - 1, generator=g)[:3].sort().values + 1 for _ in range(B)])
    masked_ids = torch.gather(text_ids, 1, masked_pos)
    masked_ids[2, 1] = -100                                   # padding of the masked-token list
    text_ids_masked = text_ids.clone().scatter_(1, masked_pos, 103)
    return dict(text_ids=text_ids, text_atts=text_atts, text_ids_masked=text_ids_masked, masked_pos=masked_pos, masked_ids=masked_ids)


def region_fixture():
    """tests/golden/gd_region_tiny.pt: the region-batch half of a GD iteration (GeneralDistill.py:158-260): `ret_bbox_loss=True`, images
    replicated per region inside the last `local_attn_depth` ViT layers with a patch-subset mask, bbox head, L1 + GIoU."""
    svis, tvis = dict(STUDENT_VIS, local_attn_depth=2), dict(TEACHER_VIS, local_attn_depth=4)      # the ratios of config_clipvit{_small,B}.json
    student, teacher, scfg, tcfg = build_pair(svis, tvis)
    g = torch.Generator().manual_seed(77)
    n_img, L = 3, 9
    idx_to_group_img = torch.tensor([0, 0, 1, 2, 2])
    R = idx_to_group_img.numel()
    image = torch.randn(n_img, 3, 32, 32, generator=g)
    image_atts = torch.ones(R, 5, dtype=torch.long)
    image_atts[0, 2:4] = 0
    image_atts[1, 1] = 0
    image_atts[3, 3:] = 0                                     # rows 2 and 4 are whole-image "regions"
    cxcy = torch.rand(R, 2, generator=g) * 0.4 + 0.3
    wh = torch.rand(R, 2, generator=g) * 0.3 + 0.1
    target_bbox = torch.cat([cxcy, wh], 1)
    is_image = torch.tensor([0., 0., 1., 0., 1.])
    batch = dict(image=image, idx_to_group_img=idx_to_group_img, image_atts=image_atts, target_bbox=target_bbox, is_image=is_image,
                 **text_batch(R, L, g, [(1, 5), (3, 7)]))
    kw = dict(text_ids_masked=batch["text_ids_masked"], masked_pos=batch["masked_pos"], masked_ids=batch["masked_ids"], image_atts=image_atts,
              idx_to_group_img=idx_to_group_img, target_bbox=target_bbox, is_image=is_image, ret_bbox_loss=True, output_attentions=True,
              output_hidden_states=True)
    orig_multinomial = torch.multinomial
    torch.multinomial = lambda w, n, *a, **k: torch.argmax(w, dim=-1, keepdim=True)
    try:
        student_outputs = student(image, batch["text_ids"], batch["text_atts"], **kw)
        with torch.no_grad():
            teacher_outputs = teacher(image, batch["text_ids"], batch["text_atts"], **kw)
    finally:
        torch.multinomial = orig_multinomial
    ns, code = reference_loss_code(region=True)
    ns.update(student_outputs=student_outputs, teacher_outputs=teacher_outputs, device="cpu", args=types.SimpleNamespace(temperature=1.0))
    exec(code, ns)
    total = ns["loss_in_total"]
    sp = dict(student.named_parameters())
    gn = GRAD_NAMES + ["bbox_head.0.weight", "bbox_head.3.bias", "vision_encoder.encoder.layers.5.self_attn.v_proj.weight"]
    grads = torch.autograd.grad(total, [sp[n] for n in gn])
    so, to = student_outputs, teacher_outputs
    save("gd_region_tiny", dict(
        scfg=dict(scfg, text_encoder=None, vision_config=None), tcfg=dict(tcfg, text_encoder=None, vision_config=None), vis=svis, tvis=tvis,
        bert=BERT, s_sd_spec=spec(student), t_sd_spec=spec(teacher), batch=batch, total=cpu(total),
        parts={k: cpu(ns[k]) for k in PART_NAMES}, loss={k: cpu(v) for k, v in so["loss"].items()},
        s_itm_logits=cpu(so["logits_dict"]["itm_head_logits"]), t_itm_logits=cpu(to["logits_dict"]["itm_head_logits"]),
        s_image_hidden_shapes=[tuple(h.shape) for h in so["hidden_dict"]["image_hidden_states"]],
        s_image_attn_shapes=[tuple(a.shape) for a in so["attention_dict"]["image_attentions"]],
        s_bbox_hidden_last=cpu(so["hidden_dict"]["bbox_hidden_states"][-1]),
        counts={k: len(v) for d in (so["hidden_dict"], so["attention_dict"], so["cross_attention_dict"]) for k, v in d.items()},
        grad_names=gn, grads=cpu(grads)))
    print("region total", float(total), {k: round(float(v), 5) for k, v in so["loss"].items()})


def main():
    ref_shim.install()
    student, teacher, scfg, tcfg = build_pair(STUDENT_VIS, TEACHER_VIS)
    g = torch.Generator().manual_seed(2024)
    B, L = 5, 9
    image = torch.randn(B, 3, 32, 32, generator=g)
    batch = dict(image=image, **text_batch(B, L, g, [(1, 6), (4, 4)]))
    kw = dict(text_ids_masked=batch["text_ids_masked"], masked_pos=batch["masked_pos"], masked_ids=batch["masked_ids"], output_attentions=True,
              output_hidden_states=True)
    orig_multinomial = torch.multinomial
    torch.multinomial = lambda w, n, *a, **k: torch.argmax(w, dim=-1, keepdim=True)   # deterministic hard negatives on both sides
    try:
        student_outputs = student(image, batch["text_ids"], batch["text_atts"], **kw)
        with torch.no_grad():
            teacher_outputs = teacher(image, batch["text_ids"], batch["text_atts"], **kw)
    finally:
        torch.multinomial = orig_multinomial
    ns, code = reference_loss_code()
    ns.update(student_outputs=student_outputs, teacher_outputs=teacher_outputs, device="cpu", args=types.SimpleNamespace(temperature=1.0))
    exec(code, ns)                                             # GeneralDistill.py:300-376, as written
    total = ns["loss_in_total"]
    sp = dict(student.named_parameters())
    grads = torch.autograd.grad(total, [sp[n] for n in GRAD_NAMES])
    so, to = student_outputs, teacher_outputs
    save("gd_kd_tiny", dict(
        scfg=dict(scfg, text_encoder=None, vision_config=None), tcfg=dict(tcfg, text_encoder=None, vision_config=None), vis=STUDENT_VIS,
        tvis=TEACHER_VIS, bert=BERT, s_sd_spec=spec(student), t_sd_spec=spec(teacher), batch=batch, total=cpu(total),
        parts={k: cpu(ns[k]) for k in PART_NAMES}, loss={k: cpu(v) for k, v in so["loss"].items()},
        s_itm_logits=cpu(so["logits_dict"]["itm_head_logits"]), s_mlm_logits=cpu(so["logits_dict"]["mlm_logits"]),
        t_itm_logits=cpu(to["logits_dict"]["itm_head_logits"]), t_mlm_logits=cpu(to["logits_dict"]["mlm_logits"]),
        s_mlm_hidden_last=cpu(so["hidden_dict"]["mlm_hidden_states"][-1]), s_neg_attn_last=cpu(so["attention_dict"]["itm_neg_attentions"][-1]),
        counts={k: len(v) for d in (so["hidden_dict"], so["attention_dict"], so["cross_attention_dict"]) for k, v in d.items()},
        t_counts={k: len(v) for d in (to["hidden_dict"], to["attention_dict"], to["cross_attention_dict"]) for k, v in d.items()},
        grad_names=GRAD_NAMES, grads=cpu(grads)))
    print("total", float(total.detach()), {k: round(float(ns[k].detach()), 5) for k in ("loss_small", "loss_kd")})
    region_fixture()


if __name__ == "__main__":
    main()
