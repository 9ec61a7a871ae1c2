"""TEST INFRASTRUCTURE ONLY — generates tests/golden/caption_kd_tiny.pt: one captioning pruning step's loss
(Eff_Captioning.py:91-148) and a greedy decode, computed by the UNMODIFIED reference classes — student
`efficient_models/model_generation.py::EffXVLMForCaptioning`, teacher `models/model_generation.py::XVLMForCaptioning` — with the
reference's own KD helpers (extracted from Eff_Captioning.py with `ast`) and oracle/fake_tokenizer.py standing in for the
bert-base-uncased tokenizer (not available offline).

    python oracle/make_golden_caption.py
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
from oracle.fake_tokenizer import FakeTokenizer  # noqa: E402
from oracle.make_golden import BERT, VIS, cpu, save, spec  # noqa: E402

TEACHER_VIS = dict(VIS, num_hidden_layers=4, local_attn_depth=0)
CAPTIONS = ["a picture of two dogs running on the beach", "a picture of a red bus", "a picture of an old man reading a very long newspaper outside"]


def main():
    ref_shim.install()
    ds = types.ModuleType("dataset")
    ds.build_tokenizer = lambda *a, **k: FakeTokenizer(BERT["vocab_size"])
    sys.modules["dataset"] = ds
    g = torch.Generator().manual_seed(41)
    vj, td = ref_shim.make_config_dir(dict(VIS, local_attn_depth=0), BERT)
    tvj, _ = ref_shim.make_config_dir(TEACHER_VIS, BERT)
    base = os.path.dirname(td)
    # the reference insists on config['text_encoder'] == 'data/bert-base-uncased' (a relative path): run from a directory that has it
    os.makedirs(os.path.join(base, "data"), exist_ok=True)
    os.symlink(td, os.path.join(base, "data", "bert-base-uncased"))
    os.symlink(os.path.join(ref_shim.REF_ROOT, "configs"), os.path.join(base, "configs"))
    scfg = dict(text_encoder="data/bert-base-uncased", vision_config=vj, patch_size=16, image_res=32, use_clip_vit=True, use_swin=False,
                text_num_hidden_layers=6, prompt="a picture of ", max_tokens=12, label_smoothing=0.1, sparsity=0.3)
    tcfg = dict(scfg, vision_config=tvj, text_num_hidden_layers=12)
    cwd = os.getcwd()
    os.chdir(base)
    from efficient_models.model_generation import EffXVLMForCaptioning
    from models.model_generation import XVLMForCaptioning
    # The greedy loop (eff_bert.py:1535-1537) calls transformers 4.12.5's GenerationMixin._update_model_kwargs_for_generation, which
    # 5.x PreTrainedModel no longer has.  Restated from the published 4.12.5 behaviour (un-vendored third-party code: parity
    # unpinned): carry the KV cache over as `past` and extend the attention mask by one column.
    import efficient_models.eff_bert as eb
    import models.xbert as xb

    def _update_model_kwargs_for_generation(self, outputs, model_kwargs, is_encoder_decoder=False):
        model_kwargs["past"] = outputs.past_key_values if getattr(outputs, "past_key_values", None) is not None else None
        if not is_encoder_decoder and model_kwargs.get("attention_mask", None) is not None:
            am = model_kwargs["attention_mask"]
            model_kwargs["attention_mask"] = torch.cat([am, am.new_ones((am.shape[0], 1))], dim=-1)
        return model_kwargs
    for mod in (eb, xb):
        mod.BertLMHeadModel._update_model_kwargs_for_generation = _update_model_kwargs_for_generation
    torch.manual_seed(10)
    student = EffXVLMForCaptioning(scfg).eval()
    teacher = XVLMForCaptioning(tcfg).eval()
    os.chdir(cwd)
    det_init_module_(student)
    det_init_module_(teacher)
    for m in (student, teacher):
        m.text_decoder.cls.predictions.decoder.weight = m.text_decoder.bert.embeddings.word_embeddings.weight
    with torch.no_grad():
        for k, la in student.l0_module.z_logas.items():
            la.copy_(torch.randn(la.shape, generator=g) * 1.5 + 1.0)
        student.l0_module.lambda_1.fill_(0.2)
        student.l0_module.lambda_2.fill_(0.9)
    student.l0_module.set_lagrangian_warmup_steps(25)
    image = torch.randn(len(CAPTIONS), 3, 32, 32, generator=g)
    eps = {k: torch.rand(la.shape, generator=g).clamp(1e-6, 1 - 1e-6) for k, la in student.l0_module.z_logas.items()}

    def arm():
        it = iter([eps[k] for k in student.l0_module.types])
        student.l0_module.get_eps = lambda size: next(it)
    arm()
    so = student(image, CAPTIONS, output_attentions=True, output_hidden_states=True)
    with torch.no_grad():
        to = teacher(image, CAPTIONS, output_attentions=True, output_hidden_states=True)
    src = open(os.path.join(ref_shim.REF_ROOT, "Eff_Captioning.py")).read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("get_kd_loss", "soft_cross_entropy", "get_cor_teacher")]
    ns = {"torch": torch, "KLDivLoss": torch.nn.KLDivLoss}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "Eff_Captioning.py", "exec"), ns)
    get_kd_loss, soft_cross_entropy, get_cor_teacher = ns["get_kd_loss"], ns["soft_cross_entropy"], ns["get_cor_teacher"]
    mse, dev, temperature = torch.nn.MSELoss(), "cpu", 1.0
    sh, th, sa, ta = so["hidden_dict"], to["hidden_dict"], so["attention_dict"], to["attention_dict"]
    sc, tc = so["cross_attention_dict"], to["cross_attention_dict"]
    # ---- Eff_Captioning.py:110-148, statement by statement ----
    image_hidden_loss = get_kd_loss(sh["image_hidden_states"], get_cor_teacher(th["image_hidden_states"], sh["image_hidden_states"]), False, mse, dev, is_img=True)
    image_attention_loss = get_kd_loss(sa["image_attentions"], get_cor_teacher(ta["image_attentions"], sa["image_attentions"], is_attn=True), True, mse, dev)
    decoder_hidden_loss = get_kd_loss(sh["decoder_hidden_states"], get_cor_teacher(th["decoder_hidden_states"], sh["decoder_hidden_states"]), False, mse, dev, is_img=True)
    decoder_attention_loss = get_kd_loss(sa["decoder_attentions"], get_cor_teacher(ta["decoder_attentions"], sa["decoder_attentions"], is_attn=True), True, mse, dev)
    decoder_cross_loss = get_kd_loss(sc["decoder_cross_attentions"], get_cor_teacher(tc["decoder_cross_attentions"], sc["decoder_cross_attentions"], is_attn=True), True, mse, dev)
    logits_loss = soft_cross_entropy(so["logits_dict"]["logits"] / temperature, to["logits_dict"]["logits"] / temperature)
    loss_small = so["loss"]
    loss_img_kd = image_attention_loss + image_hidden_loss * 0.1
    loss_decoder_kd = decoder_attention_loss + decoder_hidden_loss + decoder_cross_loss
    loss_kd = logits_loss + loss_img_kd + loss_decoder_kd
    loss = loss_kd * 0.3 + loss_small * 0.7
    lagrangian_loss, _, _ = student.l0_module.lagrangian_regularization(9)
    loss = loss + lagrangian_loss
    gn = ["vision_encoder.encoder.layers.0.mlp.fc2.weight", "text_decoder.bert.encoder.layer.1.attention.self.value.weight",
          "text_decoder.bert.encoder.layer.4.crossattention.self.query.weight", "text_decoder.cls.predictions.transform.dense.weight",
          "text_decoder.bert.embeddings.word_embeddings.weight", "l0_module.text_head_loga", "l0_module.cross_int_loga", "l0_module.lambda_1"]
    sp = dict(student.named_parameters())
    grads = torch.autograd.grad(loss, [sp[n] for n in gn])
    arm()
    loss_plain = student(image, CAPTIONS)
    with torch.no_grad():
        caps = student.generate(image, greedy=True, max_length=10)
        caps_rp = student.generate(image, greedy=True, max_length=12, repetition_penalty=1.3)
        # sampling branch (model_generation.py:455-469: do_sample=True, temperature 1, no top-k / top-p): one torch.multinomial draw
        # per step over the batch, so a seeded CPU generator gives the same tokens on both sides
        torch.manual_seed(1234)
        caps_sample, logprobs_sample = student.generate(image, sample=True, max_length=12, repetition_penalty=1.1)
        # (the teacher's generate() is broken as shipped: models/model_generation.py:160 keeps the vision tower's output TUPLE)
    tok = student.tokenizer(CAPTIONS, padding="longest", truncation=True, max_length=12, return_tensors="pt")
    save("caption_kd_tiny", dict(
        scfg=dict(scfg, text_encoder=None, vision_config=None), tcfg=dict(tcfg, text_encoder=None, vision_config=None), vis=dict(VIS, local_attn_depth=0),
        tvis=TEACHER_VIS, bert=BERT, s_sd_spec=spec(student), t_sd_spec=spec(teacher), captions=CAPTIONS, input_ids=tok.input_ids,
        prompt_length=student.prompt_length, l0_logas={k: cpu(v) for k, v in student.l0_module.z_logas.items()}, lambda_1=0.2, lambda_2=0.9,
        warmup=25, step=9, eps=eps, image=image, total=cpu(loss), loss_plain=cpu(loss_plain), s_logits=cpu(so["logits_dict"]["logits"]),
        t_logits=cpu(to["logits_dict"]["logits"]), s_dec_cross_last=cpu(sc["decoder_cross_attentions"][-1]),
        parts=dict(image_hidden=cpu(image_hidden_loss), image_attention=cpu(image_attention_loss), decoder_hidden=cpu(decoder_hidden_loss),
                   decoder_attention=cpu(decoder_attention_loss), decoder_cross=cpu(decoder_cross_loss), logits=cpu(logits_loss),
                   loss_small=cpu(loss_small), lagrangian=cpu(lagrangian_loss)),
        greedy_captions=caps, greedy_captions_rp13=caps_rp, sample_seed=1234, sample_captions=caps_sample, sample_logprobs=cpu(logprobs_sample),
        grad_names=gn, grads=cpu(grads)))
    print("student greedy:", caps, caps_rp)
    print("student sample:", caps_sample, logprobs_sample)
    print("done")


if __name__ == "__main__":
    main()
