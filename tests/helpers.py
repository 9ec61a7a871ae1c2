"""Shared test helpers: golden loading, deterministic weights, error metrics."""
import os

import torch

from oracle.det_init import det_state_dict

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    return torch.load(os.path.join(GOLDEN, name + ".pt"), map_location="cpu", weights_only=False)


def sd_from_spec(spec):
    """Rebuild the reference's weights from a fixture's {name: (shape, dtype)} spec."""
    raw = {k: torch.empty(shape, dtype=getattr(torch, dt.replace("torch.", ""))) for k, (shape, dt) in spec.items()}
    out = det_state_dict(raw)
    for k, v in out.items():
        # entries det_state_dict leaves alone keep the reference constructor's values in the fixtures
        if k.endswith("temp"):
            out[k] = torch.full_like(v, 0.07)
        elif "lambda" in k or "loga" in k:
            out[k] = torch.zeros_like(v)
        if k.endswith("position_ids"):
            out[k] = torch.arange(v.shape[-1]).expand(v.shape).clone()
    return out


def rel_err(a, b):
    a, b = a.detach().float().cpu(), b.detach().float().cpu()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


# ---- the measured bf16 noise of the reference's own PyTorch path -----------------------------------------------------------------
# north_star: "within 1e-2 relative (bf16)" of "the reference PyTorch path".  At bf16 that path is the reference's arithmetic as eager
# PyTorch under torch.autocast(bfloat16) on the GPU; its own distance from the fp32 fixture is the noise floor of ANY bf16
# implementation of the same graph (operand rounding through N layers at width 128: little averaging).  `with_reference_bf16_noise`
# runs a test body twice: first with the oracle's operators (tests/ref_ops.py -> oracle/xvlm_oracle.py) under autocast on the GPU,
# recording the relative error of every compared tensor against the fixture; then on the CUDA product, where each tensor must be within
# max(tol, NOISE_KAPPA x the reference's own bf16 error for THAT tensor).  tol is north_star's 1e-2 everywhere.
#
# Both errors are single draws of rounding noise, so their per-tensor ratio scatters around 1.  Measured on the B200 over all 80 compared
# tensors whose error exceeds 1e-2 (profiles/r02_parity_calibration.txt): median ours / reference = 0.88, largest 1.38 on tensors and
# 2.02 on a scalar (d loss / d temp, one number); per test the geometric mean of the ratio is 0.75 - 1.29 (0.75 - 1.11 for the tests with
# at least 8 such tensors).  The bars: 1.5 x per tensor, 2.5 x for fewer than 8 elements, AND over a whole test with >= 8 tensors above
# 1e-2 the geometric mean of ours / reference must stay below NOISE_GEOMEAN — the product may not be systematically noisier than the
# reference's own bf16 path.
NOISE_KAPPA = 1.5
NOISE_KAPPA_SCALAR = 2.5
NOISE_GEOMEAN = 1.25
_mode = {"kind": None, "noise": None, "seen": None, "ratios": None}


def recording():
    """True while the body runs on the eager-PyTorch bf16-autocast reference (exact-id checks of the PRODUCT are skipped there)."""
    return _mode["kind"] == "record"


def _key(what):
    n = _mode["seen"].get(what, 0)
    _mode["seen"][what] = n + 1
    return "%s#%d" % (what, n)


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    if _mode["kind"] == "record":
        _mode["noise"][_key(what)] = e
        return
    noise = None
    if _mode["kind"] == "bound":
        noise = _mode["noise"].get(_key(what))
    kappa = NOISE_KAPPA if (torch.is_tensor(a) and a.numel() >= 8) else NOISE_KAPPA_SCALAR
    bar = tol if noise is None else max(tol, kappa * noise)
    if noise is not None and e > tol and noise > 0:
        _mode["ratios"].append(e / noise)
    log = os.environ.get("EVLM_CALIBRATE_LOG")
    if log:  # calibration pass: record every measured error next to its bar (scripts/gpu_r2_first.sh); EVLM_CALIBRATE_NOFAIL=1 keeps going
        import json
        with open(log, "a") as f:
            f.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", ""), "what": what, "err": e, "tol": tol,
                                "ref_bf16_noise": noise}) + "\n")
        if os.environ.get("EVLM_CALIBRATE_NOFAIL"):
            return
    if noise is None:
        assert e <= tol, "%s: relative error %.3e > %.1e" % (what, e, tol)
    else:
        assert e <= bar, "%s: relative error %.3e > max(%.1e, %.2f x %.3e = the reference's own bf16-autocast error)" % (
            what, e, tol, kappa, noise)


def with_reference_bf16_noise(fn):
    """Decorator for a `-m gpu` test (see the block comment above)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*args, **kw):
        import efficientvlm_b200.kernels as K
        import efficientvlm_b200.ops as ops
        from tests import ref_ops
        saved = []

        class MP:
            def setattr(self, obj, name, val):
                if isinstance(obj, str):
                    raise NotImplementedError
                saved.append((obj, name, getattr(obj, name)))
                setattr(obj, name, val)
        noise = {}
        _mode.update(kind="record", noise=noise, seen={})
        incomplete = None
        try:
            ref_ops.install(MP())
            with torch.autocast("cuda", dtype=torch.bfloat16):
                fn(*args, **kw)
        except Exception as exc:   # an exact-id check of the product or a CUDA-only helper inside the body: what was recorded stands,
            incomplete = exc       # every tensor after it gets the plain 1e-2 bar (the stricter direction)
        finally:
            for obj, name, val in reversed(saved):
                setattr(obj, name, val)
            _mode.update(kind=None, noise=None, seen=None)
        if os.environ.get("EVLM_CALIBRATE_LOG") and incomplete is not None:
            import json
            with open(os.environ["EVLM_CALIBRATE_LOG"], "a") as f:
                f.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", ""), "record_pass_stopped": repr(incomplete)[:300],
                                    "recorded": len(noise)}) + "\n")
        ops.invalidate_weight_cache()
        ratios = []
        _mode.update(kind="bound", noise=noise, seen={}, ratios=ratios)
        try:
            out = fn(*args, **kw)
        finally:
            _mode.update(kind=None, noise=None, seen=None, ratios=None)
        if ratios:
            import math
            gm = math.exp(sum(math.log(r) for r in ratios) / len(ratios))
            if os.environ.get("EVLM_CALIBRATE_LOG"):
                import json
                with open(os.environ["EVLM_CALIBRATE_LOG"], "a") as f:
                    f.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", ""), "tensors_above_tol": len(ratios),
                                        "geomean_ours_over_reference_bf16": gm, "max_ratio": max(ratios)}) + "\n")
            assert len(ratios) < 8 or gm <= NOISE_GEOMEAN or os.environ.get("EVLM_CALIBRATE_NOFAIL"), \
                "over the %d tensors above tolerance the product is %.2f x noisier than the reference's own bf16 path" % (len(ratios), gm)
        return out
    return wrapper


class Tokens:
    """Stand-in for the tokenizer output the VQA models read (`.input_ids`, `.attention_mask`)."""

    def __init__(self, input_ids, attention_mask):
        self.input_ids, self.attention_mask = input_ids, attention_mask

    def to(self, device):
        return Tokens(self.input_ids.to(device), self.attention_mask.to(device))


L0_PARAM = {"vision_head": "vision_head_loga", "text_head": "text_head_loga", "cross_head": "cross_head_loga", "decoder_head": "decoder_head_loga",
            "vision_intermediate": "vision_int_loga", "text_intermediate": "text_int_loga", "cross_intermediate": "cross_int_loga",
            "decoder_intermediate": "decoder_int_loga"}


def build_with_tiny_bert(cls, cfg, bert_kwargs):
    """Construct a task model with the fixture's tiny BERT instead of the bert-base-uncased defaults."""
    import efficientvlm_b200.eff_bert as eb
    orig = eb.BertConfig.__init__

    def patched(self, **kw):
        merged = dict(bert_kwargs)
        merged.update(kw)
        orig(self, **merged)
    eb.BertConfig.__init__ = patched
    try:
        return cls(cfg)
    finally:
        eb.BertConfig.__init__ = orig


def vqa_models(g):
    """(student EffXVLMForVQA, teacher XVLMForVQA) of tests/golden/vqa_tiny.pt with the fixture's weights, strict key check."""
    from efficientvlm_b200.vqa import EffXVLMForVQA, XVLMForVQA
    out = []
    for cls, cfg, vis, spec_key in ((EffXVLMForVQA, g["scfg"], g["vis"], "s_sd_spec"), (XVLMForVQA, g["tcfg"], g["tvis"], "t_sd_spec")):
        cfg = dict(cfg, vision_config=dict(vis), text_encoder=None)
        m = build_with_tiny_bert(cls, cfg, g["bert"])
        sd = sd_from_spec(g[spec_key])
        sd["text_decoder.cls.predictions.decoder.weight"] = sd["text_decoder.bert.embeddings.word_embeddings.weight"]
        if cls is EffXVLMForVQA:
            for k, v in g["l0_logas"].items():
                sd["l0_module." + L0_PARAM[k]] = v
            sd["l0_module.lambda_1"] = torch.tensor(g["lambda_1"])
            sd["l0_module.lambda_2"] = torch.tensor(g["lambda_2"])
        m.load_state_dict(sd, strict=True)
        out.append(m.eval())
    out[0].l0_module.set_lagrangian_warmup_steps(g["warmup"])
    return out


def arm_eps(l0, eps):
    it = iter([eps[t] for t in l0.types])
    l0.get_eps = lambda size: next(it)


def itr_models(g):
    """(student EffXVLMforRetrieval, teacher XVLMforRetrieval) of tests/golden/itr_kd_tiny.pt, strict key check."""
    from efficientvlm_b200.distill import EffXVLMforRetrieval, XVLMforRetrieval
    out = []
    for cls, cfg, vis, spec_key in ((EffXVLMforRetrieval, g["scfg"], g["vis"], "s_sd_spec"), (XVLMforRetrieval, g["tcfg"], g["tvis"], "t_sd_spec")):
        cfg = dict(cfg, vision_config=dict(vis), text_encoder=None)
        m = build_with_tiny_bert(cls, cfg, g["bert"])
        sd = sd_from_spec(g[spec_key])
        if cls is EffXVLMforRetrieval:
            for k, v in g["l0_logas"].items():
                sd["l0_module." + L0_PARAM[k]] = v
            sd["l0_module.lambda_1"] = torch.tensor(g["lambda_1"])
            sd["l0_module.lambda_2"] = torch.tensor(g["lambda_2"])
        m.load_state_dict(sd, strict=True)
        out.append(m.eval())
    out[0].l0_module.set_lagrangian_warmup_steps(g["warmup"])
    return out


def argmax_negatives(model):
    """Deterministic ITM hard negatives (the fixtures patch torch.multinomial to argmax on the reference side)."""
    from oracle import xvlm_oracle as O

    def sampler(image_feat, text_feat, idx=None):
        w_i2t, w_t2i = O.itm_negative_weights(image_feat.detach(), text_feat.detach(), model.temp.detach(), idx)
        return w_t2i.argmax(1), w_i2t.argmax(1)
    return sampler


ITR_KD_TERMS = ("text_hidden", "text_attention", "image_hidden", "image_attention", "itm_pos_hidden", "itm_pos_attn", "itm_pos_cross",
                "itm_neg_hidden", "itm_neg_attn", "itm_neg_cross", "itm_logits")


def run_itr_kd_step(g, device, tol_parts, tol_total, tol_grad):
    """Shared body of the CPU (host logic) and GPU (product) ITR KD step checks against tests/golden/itr_kd_tiny.pt."""
    from efficientvlm_b200.distill import itr_loss
    student, teacher = (m.to(device) for m in itr_models(g))
    student.sample_itm_negatives = argmax_negatives(student)
    teacher.sample_itm_negatives = argmax_negatives(teacher)
    image, text_ids, text_atts, idx = (g[k].to(device) for k in ("image", "text_ids", "text_atts", "idx"))
    arm_eps(student.l0_module, g["eps"])
    so = student(image, text_ids, text_atts, idx=idx, output_attentions=True, output_hidden_states=True)
    with torch.no_grad():
        to = teacher(image, text_ids, text_atts, idx=idx, output_attentions=True, output_hidden_states=True)
    assert "loss" not in to                      # the teacher's KD branch returns no loss (models/model_retrieval.py:47-53)
    assert_close(to["logits_dict"]["itm_head_logits"], g["t_itm_logits"], tol_parts, "teacher itm logits")
    assert_close(to["cross_attention_dict"]["itm_neg_cross_attentions"][-1], g["t_neg_cross_last"], tol_parts, "teacher neg cross attention")
    total, parts = itr_loss(so, to, student.l0_module, g["step"], 1.0)
    for name in ITR_KD_TERMS:
        assert_close(parts["kd_" + name], g["parts"][name], tol_parts, "kd " + name)
    assert_close(parts["loss_itc"], g["parts"]["loss_itc"], tol_parts, "itc")
    assert_close(parts["loss_itm"], g["parts"]["loss_itm"], tol_parts, "itm")
    assert_close(parts["lagrangian_loss"], g["parts"]["lagrangian"], 1e-4, "lagrangian")
    assert_close(total, g["total"], tol_total, "total")
    sp = dict(student.named_parameters())
    grads = torch.autograd.grad(total, [sp[n] for n in g["grad_names"]])
    for n, x, y in zip(g["grad_names"], grads, g["grads"]):
        assert_close(x, y, 1e-4 if "lambda" in n else tol_grad, "grad " + n)


def nlvr_models(g):
    """(student EffXVLMForNLVR, teacher XVLMForNLVR) of tests/golden/nlvr_kd_tiny.pt, strict key check, tied cross K/V."""
    from efficientvlm_b200.nlvr import EffXVLMForNLVR, XVLMForNLVR
    out = []
    for cls, cfg, vis, spec_key in ((EffXVLMForNLVR, g["scfg"], g["vis"], "s_sd_spec"), (XVLMForNLVR, g["tcfg"], g["tvis"], "t_sd_spec")):
        cfg = dict(cfg, vision_config=dict(vis), text_encoder=None)
        m = build_with_tiny_bert(cls, cfg, g["bert"])
        sd = sd_from_spec(g[spec_key])
        # share_cross_attention ties layer L's key / value to layer L+1's; the state_dict lists the shared tensor under both names
        # and the fixture's deterministic init (load_state_dict in key order) left it with the values of the SECOND name
        n_text = m.num_text_layers
        for i in range(m.num_cross_layers):
            a, b = n_text + 2 * i, n_text + 2 * i + 1
            for kv in ("key", "value"):
                for wb in ("weight", "bias"):
                    sd["text_encoder.encoder.layer.%d.crossattention.self.%s.%s" % (a, kv, wb)] = \
                        sd["text_encoder.encoder.layer.%d.crossattention.self.%s.%s" % (b, kv, wb)]
        if cls is EffXVLMForNLVR:
            for k, v in g["l0_logas"].items():
                sd["l0_module." + L0_PARAM[k]] = v
            sd["l0_module.lambda_1"] = torch.tensor(g["lambda_1"])
            sd["l0_module.lambda_2"] = torch.tensor(g["lambda_2"])
        m.load_state_dict(sd, strict=True)
        out.append(m.eval())
    out[0].l0_module.set_lagrangian_warmup_steps(g["warmup"])
    return out


NLVR_KD_TERMS = ("text_hidden", "text_attention", "cross_hidden", "cross_self_attention", "cross_attention", "image_hidden", "image_attention", "logits")


def run_nlvr_kd_step(g, device, tol_parts, tol_total, tol_grad):
    """Shared body of the CPU (host logic) and GPU (product) NLVR2 KD step checks against tests/golden/nlvr_kd_tiny.pt."""
    from efficientvlm_b200.nlvr import nlvr_loss
    student, teacher = (m.to(device) for m in nlvr_models(g))
    assert [n for n, _ in student.named_parameters()] == g["s_param_names"]      # incl. the de-duplicated tied K / V projections
    image, text_ids, text_atts, targets = (g[k].to(device) for k in ("image", "text_ids", "text_atts", "targets"))
    arm_eps(student.l0_module, g["eps"])
    so = student(image, text_ids, text_atts, targets=targets, train=True, output_attentions=True, output_hidden_states=True)
    with torch.no_grad():
        to = teacher(image, text_ids, text_atts, targets=targets, train=True, output_attentions=True, output_hidden_states=True)
    c = g["counts"]
    assert [len(so["hidden_dict"]["text_hidden_states"]), len(so["attention_dict"]["text_attentions"]),
            len(so["cross_attention_dict"]["cross_attentions"])] == [c["s_text_h"], c["s_text_a"], c["s_cross_a"]]
    assert [len(to["hidden_dict"]["text_hidden_states"]), len(to["attention_dict"]["text_attentions"]),
            len(to["cross_attention_dict"]["cross_attentions"])] == [c["t_text_h"], c["t_text_a"], c["t_cross_a"]]
    assert_close(so["logits_dict"]["cls_head_logits"], g["s_logits"], tol_parts, "student logits")
    assert_close(to["logits_dict"]["cls_head_logits"], g["t_logits"], tol_parts, "teacher logits")
    assert_close(so["cross_attention_dict"]["cross_attentions"][-1], g["s_cross_last"], tol_parts, "last cross attention (second image)")
    total, parts = nlvr_loss(so, to, student.l0_module, g["step"], 1.0)
    for name in NLVR_KD_TERMS:
        assert_close(parts["kd_" + name], g["parts"][name], tol_parts, "kd " + name)
    assert_close(parts["loss_small"], g["parts"]["loss_small"], tol_parts, "task loss")
    assert_close(parts["loss_lagrangian"], g["parts"]["lagrangian"], 1e-4, "lagrangian")
    assert_close(total, g["total"], tol_total, "total")
    sp = dict(student.named_parameters())
    grads = torch.autograd.grad(total, [sp[n] for n in g["grad_names"]])
    for n, x, y in zip(g["grad_names"], grads, g["grads"]):
        assert_close(x, y, 1e-4 if "lambda" in n else tol_grad, "grad " + n)
    with torch.no_grad():
        pred = student(image, text_ids, text_atts, targets=targets, train=False)
    assert_close(pred, g["pred_eval"], tol_parts, "eval prediction (deterministic masks)")
    assert recording() or torch.equal(pred.argmax(1).cpu(), g["pred_eval"].argmax(1))


def caption_models(g):
    """(student EffXVLMForCaptioning, teacher XVLMForCaptioning) of tests/golden/caption_kd_tiny.pt with the fake tokenizer."""
    from oracle.fake_tokenizer import FakeTokenizer
    from efficientvlm_b200.captioning import EffXVLMForCaptioning, XVLMForCaptioning
    out = []
    for cls, cfg, vis, spec_key in ((EffXVLMForCaptioning, g["scfg"], g["vis"], "s_sd_spec"), (XVLMForCaptioning, g["tcfg"], g["tvis"], "t_sd_spec")):
        cfg = dict(cfg, vision_config=dict(vis), text_encoder=None)
        tok = FakeTokenizer(g["bert"]["vocab_size"])
        import efficientvlm_b200.eff_bert as eb
        orig = eb.BertConfig.__init__

        def patched(self, _o=orig, **kw):
            merged = dict(g["bert"])
            merged.update(kw)
            _o(self, **merged)
        eb.BertConfig.__init__ = patched
        try:
            m = cls(cfg, tokenizer=tok)
        finally:
            eb.BertConfig.__init__ = orig
        sd = sd_from_spec(g[spec_key])
        sd["text_decoder.cls.predictions.decoder.weight"] = sd["text_decoder.bert.embeddings.word_embeddings.weight"]
        if cls is EffXVLMForCaptioning:
            for k, v in g["l0_logas"].items():
                sd["l0_module." + L0_PARAM[k]] = v
            sd["l0_module.lambda_1"] = torch.tensor(g["lambda_1"])
            sd["l0_module.lambda_2"] = torch.tensor(g["lambda_2"])
        m.load_state_dict(sd, strict=True)
        out.append(m.eval())
    out[0].l0_module.set_lagrangian_warmup_steps(g["warmup"])
    return out


CAPTION_KD_TERMS = ("image_hidden", "image_attention", "decoder_hidden", "decoder_attention", "decoder_cross", "logits")


def run_caption_kd_step(g, device, tol_parts, tol_total, tol_grad, exact_decode):
    """Shared body of the CPU (host logic) and GPU (product) captioning KD step checks against tests/golden/caption_kd_tiny.pt."""
    from efficientvlm_b200.captioning import caption_loss
    student, teacher = (m.to(device) for m in caption_models(g))
    image = g["image"].to(device)
    assert student.prompt_length == g["prompt_length"]
    assert torch.equal(student.tokenizer(g["captions"], padding="longest", truncation=True, max_length=12, return_tensors="pt").input_ids, g["input_ids"])
    arm_eps(student.l0_module, g["eps"])
    so = student(image, g["captions"], output_attentions=True, output_hidden_states=True)
    with torch.no_grad():
        to = teacher(image, g["captions"], output_attentions=True, output_hidden_states=True)
    assert_close(so["logits_dict"]["logits"], g["s_logits"], tol_parts, "student logits")
    assert_close(to["logits_dict"]["logits"], g["t_logits"], tol_parts, "teacher logits")
    assert_close(so["cross_attention_dict"]["decoder_cross_attentions"][-1], g["s_dec_cross_last"], tol_parts, "decoder cross attention")
    total, parts = caption_loss(so, to, student.l0_module, g["step"], 1.0)
    for name in CAPTION_KD_TERMS:
        assert_close(parts["kd_" + name], g["parts"][name], tol_parts, "kd " + name)
    assert_close(parts["loss_small"], g["parts"]["loss_small"], tol_parts, "task loss (label smoothing 0.1, prompt masked)")
    assert_close(parts["loss_lagrangian"], g["parts"]["lagrangian"], 1e-4, "lagrangian")
    assert_close(total, g["total"], tol_total, "total")
    sp = dict(student.named_parameters())
    grads = torch.autograd.grad(total, [sp[n] for n in g["grad_names"]])
    for n, x, y in zip(g["grad_names"], grads, g["grads"]):
        assert_close(x, y, 1e-4 if "lambda" in n else tol_grad, "grad " + n)
    arm_eps(student.l0_module, g["eps"])
    assert_close(student(image, g["captions"]), g["loss_plain"], tol_parts, "task loss only (vision tower un-gated on this branch)")
    caps = student.generate(image, greedy=True, max_length=10)
    assert len(caps) == len(g["greedy_captions"]) and all(isinstance(c, str) for c in caps)
    # the sync-free loop (no per-token end-of-sequence check on the host) produces the same ids and captions
    caps2, ids2 = student.generate(image, greedy=True, max_length=10, return_ids=True, sync_free=True)
    _, ids1 = student.generate(image, greedy=True, max_length=10, return_ids=True)
    assert caps2 == caps and torch.equal(ids1, ids2)
    if exact_decode:
        assert caps == g["greedy_captions"]
        # repetition penalty (eff_bert.py:1497-1507) and the sampling branch (model_generation.py:455-469: one multinomial draw per step
        # over the batch, so the seeded CPU generator reproduces the reference's tokens and sequence log-probabilities)
        assert student.generate(image, greedy=True, max_length=12, repetition_penalty=1.3) == g["greedy_captions_rp13"]
        torch.manual_seed(g["sample_seed"])
        sampled, logprobs = student.generate(image, sample=True, max_length=12, repetition_penalty=1.1)
        assert sampled == g["sample_captions"]
        assert_close(logprobs, g["sample_logprobs"], 1e-4, "sampled sequence log-probabilities")
    elif not recording():
        _decode_on_device_vs_fp32_reference(g, student, image, caps)


class reference_ops_fp32:
    """Context: the oracle's operators (tests/ref_ops.py) stand in for the CUDA kernels, in fp32 — on CPU modules this is exactly the
    path the host-logic tests pin to the fixtures, so a `-m gpu` test can obtain the fixture's intermediate values (token ids, step
    logits) that the .pt file does not store."""

    def __enter__(self):
        from tests import ref_ops
        self.saved = []
        outer = self

        class MP:
            def setattr(self, obj, name, val):
                outer.saved.append((obj, name, getattr(obj, name)))
                setattr(obj, name, val)
        ref_ops.install(MP())
        return self

    def __exit__(self, *exc):
        for obj, name, val in reversed(self.saved):
            setattr(obj, name, val)
        return False


class _record_multinomial:
    """Records every `torch.multinomial` draw of a decode (one [B] tensor per step)."""

    def __enter__(self):
        self.orig, draws = torch.multinomial, []

        def recording_multinomial(*a, **kw):
            out = self.orig(*a, **kw)
            draws.append(out.reshape(-1).cpu())
            return out
        torch.multinomial = recording_multinomial
        return draws

    def __exit__(self, *exc):
        torch.multinomial = self.orig
        return False


def _penalised(logits, prefix_ids, repetition_penalty):
    """eff_bert.py:1497-1507 on one row of logits."""
    logits = logits.clone()
    if repetition_penalty != 1.0:
        for tok in set(prefix_ids.tolist()):
            logits[tok] = logits[tok] * repetition_penalty if logits[tok] < 0 else logits[tok] / repetition_penalty
    return logits


def _decode_on_device_vs_fp32_reference(g, student, image, caps):
    """Greedy, repetition-penalty and sampling decode ON THE DEVICE.  Token ids are index work: they must be
    `torch.equal` to the fixture's, except where the fp32 reference itself cannot decide — at the FIRST diverging position of a row the
    reference's own margin between its token and ours (after the repetition penalty) must be below 2 x 1e-2 x max|logit|, i.e. inside
    the stated logit tolerance (two logit vectors that differ by at most eps can only swap an argmax whose margin is < 2 eps).
    The fixture stores captions, not ids / step logits: those come from the fp32 oracle operators on a CPU copy of the same model,
    re-checked here against the fixture's captions."""
    ref_student, _ = caption_models(g)
    ref_student.tokenizer(g["captions"])                 # the whitespace tokenizer learns id -> word from the training captions
    kinds = {"greedy": dict(greedy=True, max_length=10), "rp13": dict(greedy=True, max_length=12, repetition_penalty=1.3)}
    gold = {"greedy": g["greedy_captions"], "rp13": g["greedy_captions_rp13"]}
    ref_ids = {}
    with reference_ops_fp32():
        for kind, kw in kinds.items():
            c, ids = ref_student.generate(g["image"], return_ids=True, **kw)
            assert c == gold[kind], "the on-box fp32 reference pass must reproduce the fixture"
            ref_ids[kind] = ids
        torch.manual_seed(g["sample_seed"])
        with _record_multinomial() as ref_draws:
            c, lp, ids = ref_student.generate(g["image"], sample=True, max_length=12, repetition_penalty=1.1, return_ids=True)
        assert c == g["sample_captions"]
        ref_ids["sample"], ref_lp = ids, lp
        zs = ref_student._zs(False)
        ref_img = ref_student.vision_encoder(g["image"], head_z=zs["vision_head_z"], mlp_z=zs["vision_intermediate_z"])[0]

        def ref_logits(prefix, row):
            d = ref_student.text_decoder
            inp = d.prepare_inputs_for_generation(prefix[None], past=None, encoder_hidden_states=ref_img[row:row + 1], encoder_attention_mask=None)
            return d(**inp, return_dict=True).logits[0, -1].float()
        ties, mismatched_rows = [], 0
        for kind, kw in kinds.items():
            c, ids = student.generate(image, return_ids=True, **kw)
            ids = ids.cpu()
            assert ids.shape == ref_ids[kind].shape
            for r in range(ids.shape[0]):
                if torch.equal(ids[r], ref_ids[kind][r]):
                    continue
                mismatched_rows += 1
                t = int((ids[r] != ref_ids[kind][r]).nonzero()[0])
                with torch.no_grad():
                    l = _penalised(ref_logits(ref_ids[kind][r, :t], r), ref_ids[kind][r, :t], kw.get("repetition_penalty", 1.0))
                a, b = int(ids[r, t]), int(ref_ids[kind][r, t])
                margin = float(l[b] - l[a])
                assert b == int(l.argmax()), "reference token is the reference argmax"
                assert 0 <= margin <= 2 * 1e-2 * float(l.abs().max()), \
                    "%s row %d step %d: ours %d vs reference %d with reference margin %.3e (max|logit| %.3e): not a tie" % (
                        kind, r, t, a, b, margin, float(l.abs().max()))
                ties.append((kind, r, t, margin))
    n_rows = sum(v.shape[0] for k, v in ref_ids.items() if k != "sample")
    assert mismatched_rows <= n_rows // 2, ties          # ties are the exception, not the rule
    log = os.environ.get("EVLM_CALIBRATE_LOG")
    if log:
        import json
        with open(log, "a") as f:
            f.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", ""), "decode_rows": n_rows, "tie_flips": ties}) + "\n")
    # sampling branch on the device (model_generation.py:455-469).  torch.multinomial on CUDA draws from the device generator, so no
    # implementation (the reference included) reproduces the CPU-seeded fixture tokens on a GPU; what is checked instead:
    # (i) the fixture's sampled sequences, teacher-forced through the device decoder, get the fixture's sequence log-probabilities;
    # (ii) the device sampler's own (tokens, log-probabilities) are self-consistent under the same teacher-forced re-score, every
    #      sequence ends with [SEP] / padding only after [SEP], and the repetition penalty is applied to the scores it returns.
    zs = student._zs(False)
    with torch.no_grad():
        img = student.vision_encoder(image, head_z=zs["vision_head_z"], mlp_z=zs["vision_intermediate_z"])[0]
    eos, pad, P = student.tokenizer.sep_token_id, student.tokenizer.pad_token_id, student.prompt_length

    def rescore(ids, draws):
        """Sequence log-probability (eff_bert.py:1520-1557: mean over the steps a row was unfinished) of the per-step draws, teacher
        forced through the DEVICE decoder; `draws[j][r]` is the token drawn for row r at step j (the last column of `ids` may have been
        overwritten with [SEP], :1551)."""
        ids = ids.to(image.device)
        d = student.text_decoder
        with torch.no_grad():
            inp = d.prepare_inputs_for_generation(ids, past=None, encoder_hidden_states=img, encoder_attention_mask=None)
            logits = d(**inp, return_dict=True).logits.float().cpu()
        ids = ids.cpu()
        out = []
        for r in range(ids.shape[0]):
            tot, n, alive = 0.0, 0, True
            for j, t in enumerate(range(P, P + len(draws))):
                if not alive:
                    break
                tok = int(draws[j][r])
                l = _penalised(logits[r, t - 1], ids[r, :t], 1.1)
                tot += float(torch.log_softmax(l, -1)[tok])
                n += 1
                alive = tok != eos
            out.append(tot / n)
        return torch.tensor(out)
    assert_close(rescore(ref_ids["sample"], ref_draws), ref_lp, 1e-2, "fixture's sampled sequences re-scored on the device")
    assert_close(ref_lp, g["sample_logprobs"], 1e-4, "on-box fp32 sampling pass vs fixture")
    torch.manual_seed(g["sample_seed"])
    with _record_multinomial() as dev_draws:
        sampled, logprobs, sids = student.generate(image, sample=True, max_length=12, repetition_penalty=1.1, return_ids=True)
    assert len(sampled) == len(g["sample_captions"]) and sids.shape == ref_ids["sample"].shape
    assert_close(rescore(sids, dev_draws), logprobs.cpu(), 2e-3, "device sampler: returned log-probabilities = teacher-forced re-score of its own draws")
    for row in sids.cpu().tolist():
        if eos in row[P:]:
            k = row.index(eos, P)
            assert all(x == pad for x in row[k + 1:]), row
        assert all(0 <= x < g["bert"]["vocab_size"] for x in row)


# ----------------------------------------------------------------------------------------------------------------------
# ITR re-rank evaluation (Eff_Retrieval.py:216-378) against tests/golden/itr_eval_tiny.pt
# ----------------------------------------------------------------------------------------------------------------------
class _EvalLoader:
    """The slice of the DataLoader surface `evaluation` reads: (image batch, ids) iteration + `.dataset.text`."""

    class _DS:
        pass

    def __init__(self, images, texts, bs):
        self.images, self.bs = images, bs
        self.dataset = self._DS()
        self.dataset.text, self.dataset.image = texts, list(range(len(images)))

    def __iter__(self):
        for i in range(0, len(self.images), self.bs):
            yield self.images[i:i + self.bs], torch.arange(i, min(len(self.images), i + self.bs))


def itr_eval_setup(g, device):
    """(student with the fixture's weights and log-alphas, data loader, tokenizer padding to max_tokens like the generator)."""
    from efficientvlm_b200.distill import EffXVLMforRetrieval
    from oracle.fake_tokenizer import FakeTokenizer
    cfg = dict(g["scfg"], vision_config=dict(g["vis"]), text_encoder=None)
    m = build_with_tiny_bert(EffXVLMforRetrieval, cfg, g["bert"])
    sd = sd_from_spec(g["s_sd_spec"])
    for k, v in g["l0_logas"].items():
        sd["l0_module." + L0_PARAM[k]] = v
    m.load_state_dict(sd, strict=True)
    m = m.eval().to(device)
    tok, L = FakeTokenizer(g["bert"]["vocab_size"]), g["config"]["max_tokens"]

    def tokenizer(text, **kw):
        enc = tok(text, **kw)
        ids = torch.zeros(len(text), L, dtype=torch.long)
        att = torch.zeros(len(text), L, dtype=torch.long)
        ids[:, :enc.input_ids.shape[1]] = enc.input_ids
        att[:, :enc.attention_mask.shape[1]] = enc.attention_mask
        return Tokens(ids, att)
    return m, _EvalLoader(g["images"], g["texts"], g["image_batch"]), tokenizer


def run_itr_eval(g, device, tol_sims, tol_scores, exact_candidates):
    """Shared body of the CPU (host logic) and GPU (product) checks of efficientvlm_b200.retrieval_eval."""
    import numpy as np
    from efficientvlm_b200 import retrieval_eval as RE
    model, loader, tokenizer = itr_eval_setup(g, device)
    enc = tokenizer(g["texts"], padding="max_length", truncation=True, max_length=g["config"]["max_tokens"], return_tensors="pt")
    assert torch.equal(enc.input_ids, g["text_ids"]) and torch.equal(enc.attention_mask, g["text_atts"])
    k = g["config"]["k_test"]
    details = {}
    s_i2t, s_t2i, sparsity = RE.evaluation(model, loader, tokenizer, device, g["config"], details=details)
    assert abs(float(sparsity) - g["sparsity"]) < 1e-6                   # deterministic masks: bit-exact kept counts
    sims = details["sims_matrix"]
    assert_close(sims, g["sims"], tol_sims, "similarity matrix")
    gi, gt = g["score_i2t"].numpy(), g["score_t2i"].numpy()
    assert s_i2t.shape == gi.shape and s_t2i.shape == gt.shape
    flips = []
    for ours, gold, sm in ((s_i2t, gi, sims), (s_t2i, gt, sims.t())):
        scored = ours != -100.0
        assert (scored.sum(1) == k).all()                                   # k_test candidates per query, -100 elsewhere
        own_topk = torch.zeros_like(sm, dtype=torch.bool).scatter_(1, sm.topk(k, dim=1)[1], True).cpu().numpy()
        assert (scored == own_topk).all()
        both = scored & (gold != -100.0)
        if exact_candidates:
            assert (scored == (gold != -100.0)).all()
        elif not recording():
            # candidate selection is index work.  On the device the similarities come out of bf16 encoders, so a candidate may swap with
            # another one ONLY where the fixture's fp32 similarities cannot separate them within the stated tolerance: every entry of the
            # symmetric difference sits within 2 x tol x max|sim| of the fixture's own k-th similarity of that row.
            gsm = (g["sims"] if sm is sims else g["sims"].t()).float()
            kth = gsm.topk(k, dim=1)[0][:, -1:]
            diff = torch.from_numpy(scored != (gold != -100.0))
            dist = ((gsm - kth).abs() * diff).max(1)[0]
            bar = 2 * tol_sims * gsm.abs().max(1)[0]
            assert bool((dist <= bar).all()), "candidate swap outside the similarity tolerance: %r vs %r" % (dist.tolist(), bar.tolist())
            flips.append(int(diff.sum()) // 2)
            if os.environ.get("EVLM_CALIBRATE_LOG"):
                import json
                with open(os.environ["EVLM_CALIBRATE_LOG"], "a") as f:
                    f.write(json.dumps({"test": os.environ.get("PYTEST_CURRENT_TEST", ""), "itr_candidate_swaps": flips[-1],
                                        "candidates": int(scored.sum())}) + "\n")
        assert_close(torch.from_numpy(ours[both]), torch.from_numpy(gold[both]), tol_scores, "ITM re-rank scores")
    if exact_candidates:
        assert RE.itm_eval(s_i2t, s_t2i, g["txt2img"], g["img2txt"]) == g["result"]
    # the re-rank loops on the GOLDEN similarity matrix (identical candidates on both sides), every batching mode and the
    # reference's `size // world + 1` row split
    gs = g["sims"].to(device)
    zs = details["zs"]
    args = (model, details["image_feats"], details["text_feats"], details["text_atts"], gs, k, zs["cross_head_z"], zs["cross_intermediate_z"])
    for kw in (dict(queries_per_pass=1, group_rows=1), dict(queries_per_pass=3, group_rows=2), dict(queries_per_pass=64, group_rows=16),
               dict(group_rows=5), dict(queries_per_pass=4, group_rows=3, kv_cache_bytes=0), dict(queries_per_pass=1, share_image_kv=False),
               dict(queries_per_pass=5, share_image_kv=False)):
        a, b = RE.rerank_scores(*args, rank=0, world=1, **kw)
        for ours, gold in ((a, g["score_i2t"]), (b, g["score_t2i"])):
            ours = ours.cpu()
            assert torch.equal(ours == -100.0, gold == -100.0), kw
            assert_close(ours, gold, tol_scores, "re-rank scores %r" % (kw,))
    total = [torch.zeros_like(g["score_i2t"]), torch.zeros_like(g["score_t2i"])]
    for r in range(2):
        a, b = RE.rerank_scores(*args, queries_per_pass=2, group_rows=3, rank=r, world=2)
        for j, (ours, gold) in enumerate(((a, g["per_rank"][r][0]), (b, g["per_rank"][r][1]))):
            assert torch.equal(ours.cpu() == -100.0, gold == -100.0)
            assert_close(ours.cpu(), gold, tol_scores, "rank %d scores" % r)
            total[j] += ours.cpu()
    # what the SUM all-reduce leaves: a scored pair carries its score - 100, an unscored one -200
    assert torch.equal(total[0] == -200.0, g["score_i2t"] == -100.0)
    assert np.isfinite(s_i2t).all() and np.isfinite(s_t2i).all()


# ----------------------------------------------------------------------------------------------------------------------
# general-distillation step (the headline workload) against tests/golden/gd_kd_tiny.pt
# ----------------------------------------------------------------------------------------------------------------------
GD_KD_TERMS = {"text_hidden": "text_hidden_loss", "text_attn": "text_attention_loss", "image_hidden": "image_hidden_loss",
               "image_attn": "image_attention_loss", "itm_pos_hidden": "itm_pos_hidden_loss", "itm_pos_attn": "itm_pos_attn_loss",
               "itm_neg_hidden": "itm_neg_hidden_loss", "itm_neg_attn": "itm_neg_attn_loss", "mlm_hidden": "mlm_hidden_loss",
               "mlm_attn": "mlm_attn_loss", "mlm_logits": "mlm_logits_loss", "itm_logits": "itm_logits_loss"}


def gd_models(g):
    """(student, teacher) `distill.XVLM` (models/model_pretrain.py::XVLM) of tests/golden/gd_kd_tiny.pt, strict key check."""
    from efficientvlm_b200.distill import XVLM
    out = []
    for cfg, vis, spec_key in ((g["scfg"], g["vis"], "s_sd_spec"), (g["tcfg"], g["tvis"], "t_sd_spec")):
        cfg = dict(cfg, vision_config=dict(vis), text_encoder=None)
        m = build_with_tiny_bert(XVLM, cfg, g["bert"])
        sd = sd_from_spec(g[spec_key])
        sd["text_encoder.cls.predictions.decoder.weight"] = sd["text_encoder.bert.embeddings.word_embeddings.weight"]
        m.load_state_dict(sd, strict=True)
        out.append(m.eval())
    return out


def run_gd_kd_step(g, device, tol_parts, tol_total, tol_grad, batch_passes=True):
    """Shared body of the GD-step checks against the fixture from the unmodified `models/model_pretrain.py::XVLM` + the reference's own
    train-loop statements (GeneralDistill.py:300-376): task losses, logits, every KD term, the 0.6 / 0.4 mix, 13 gradients."""
    from efficientvlm_b200.distill import gd_kd_losses, gd_loss
    student, teacher = (m.to(device) for m in gd_models(g))
    student.batch_passes = teacher.batch_passes = batch_passes
    student.sample_itm_negatives = argmax_negatives(student)
    teacher.sample_itm_negatives = argmax_negatives(teacher)
    b = {k: v.to(device) for k, v in g["batch"].items()}
    args = (b["image"], b["text_ids"], b["text_atts"])
    kw = dict(text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"], masked_ids=b["masked_ids"], output_attentions=True,
              output_hidden_states=True)
    so = student(*args, **kw)
    with torch.no_grad():
        to = teacher(*args, **kw)
    for k in ("loss_itc", "loss_itm", "loss_mlm"):
        assert_close(so["loss"][k], g["loss"][k], tol_parts, k)
    assert_close(so["logits_dict"]["itm_head_logits"], g["s_itm_logits"], tol_parts, "student itm logits")
    assert_close(so["logits_dict"]["mlm_logits"], g["s_mlm_logits"], tol_parts, "student mlm logits")
    assert_close(to["logits_dict"]["itm_head_logits"], g["t_itm_logits"], tol_parts, "teacher itm logits")
    assert_close(to["logits_dict"]["mlm_logits"], g["t_mlm_logits"], tol_parts, "teacher mlm logits")
    assert_close(so["hidden_dict"]["mlm_hidden_states"][-1], g["s_mlm_hidden_last"], tol_parts, "mlm hidden")
    assert_close(so["attention_dict"]["itm_neg_attentions"][-1], g["s_neg_attn_last"], tol_parts, "itm neg attention")
    for d in ("hidden_dict", "attention_dict", "cross_attention_dict"):
        for k, v in so[d].items():
            assert len(v) == g["counts"][k], k
    kd = gd_kd_losses(so, to, 1.0)
    for ours, theirs in GD_KD_TERMS.items():
        assert_close(kd[ours], g["parts"][theirs], tol_parts, theirs)
    total, parts = gd_loss(so, to, 1.0)
    for k in ("loss_small", "loss_text_kd", "loss_img_kd", "loss_cross_kd", "loss_kd"):
        assert_close(parts[k], g["parts"][k], tol_parts, k)
    assert_close(total, g["total"], tol_total, "loss_in_total")
    sp = dict(student.named_parameters())
    grads = torch.autograd.grad(total, [sp[n] for n in g["grad_names"]])
    for n, x, y in zip(g["grad_names"], grads, g["grads"]):
        assert_close(x, y, tol_grad, "grad " + n)


def run_gd_region_step(g, device, tol_parts, tol_total, tol_grad):
    """The region-batch half of a GD iteration (GeneralDistill.py:158-260: `ret_bbox_loss=True`, images replicated per region inside the
    local ViT layers under a patch-subset mask, bbox head, L1 + GIoU) against tests/golden/gd_region_tiny.pt — the unmodified reference
    model class + the reference's own statements of that branch."""
    from efficientvlm_b200.distill import gd_kd_losses, gd_loss
    student, teacher = (m.to(device) for m in gd_models(g))
    student.sample_itm_negatives = argmax_negatives(student)
    teacher.sample_itm_negatives = argmax_negatives(teacher)
    b = {k: v.to(device) for k, v in g["batch"].items()}
    kw = dict(text_ids_masked=b["text_ids_masked"], masked_pos=b["masked_pos"], masked_ids=b["masked_ids"], image_atts=b["image_atts"],
              idx_to_group_img=b["idx_to_group_img"], target_bbox=b["target_bbox"], is_image=b["is_image"], ret_bbox_loss=True,
              output_attentions=True, output_hidden_states=True)
    so = student(b["image"], b["text_ids"], b["text_atts"], **kw)
    with torch.no_grad():
        to = teacher(b["image"], b["text_ids"], b["text_atts"], **kw)
    for k in ("loss_itc", "loss_itm", "loss_mlm", "loss_bbox", "loss_giou"):
        assert_close(so["loss"][k], g["loss"][k], tol_parts, k)
    assert_close(so["logits_dict"]["itm_head_logits"], g["s_itm_logits"], tol_parts, "student itm logits")
    assert_close(to["logits_dict"]["itm_head_logits"], g["t_itm_logits"], tol_parts, "teacher itm logits")
    # the local layers run on [regions + images] rows, the others on [images] rows
    assert [tuple(h.shape) for h in so["hidden_dict"]["image_hidden_states"]] == g["s_image_hidden_shapes"]
    assert [tuple(a.shape) for a in so["attention_dict"]["image_attentions"]] == g["s_image_attn_shapes"]
    assert_close(so["hidden_dict"]["bbox_hidden_states"][-1], g["s_bbox_hidden_last"], tol_parts, "bbox fusion hidden")
    for d in ("hidden_dict", "attention_dict", "cross_attention_dict"):
        for k, v in so[d].items():
            assert len(v) == g["counts"][k], k
    kd = gd_kd_losses(so, to, 1.0)
    for ours, theirs in GD_KD_TERMS.items():
        assert_close(kd[ours], g["parts"][theirs], tol_parts, theirs)
    _, parts = gd_loss(so, to, 1.0)
    assert_close(parts["loss_kd"], g["parts"]["loss_kd"], tol_parts, "loss_kd")
    loss = so["loss"]
    loss_small = loss["loss_itc"] + loss["loss_itm"] + loss["loss_mlm"] + loss["loss_bbox"] + loss["loss_giou"]       # GeneralDistill.py:257
    assert_close(loss_small, g["parts"]["loss_small"], tol_parts, "loss_small")
    total = 0.6 * loss_small + 0.4 * parts["loss_kd"]                                                                 # GeneralDistill.py:259
    assert_close(total, g["total"], tol_total, "loss_in_total")
    sp = dict(student.named_parameters())
    grads = torch.autograd.grad(total, [sp[n] for n in g["grad_names"]])
    for n, x, y in zip(g["grad_names"], grads, g["grads"]):
        assert_close(x, y, tol_grad, "grad " + n)


def run_beam_search_vs_oracle(g, device, exact=True):
    """Beam search of the captioning evaluation (model_generation.py:471-483 with Captioning.yaml's num_beams 3 / min_length 5, and
    variants) on our captioning student against oracle/beam_search_oracle.py (transformers 4.12.5's published algorithm in plain
    Python; parity unpinned: third-party code).  The oracle's model is the SAME decoder run without a KV cache on the full prefixes,
    so the product's cache re-ordering, log-softmax / processors / top-k on the device and the hypothesis bookkeeping are all checked.
    exact=False (bf16 device arithmetic): a differing sequence must score within 1e-2 of the oracle's choice under the oracle's model."""
    from oracle.beam_search_oracle import beam_search
    student, _ = caption_models(g)
    student = student.to(device).eval()
    image = g["image"].to(device)
    tok = student.tokenizer
    vocab = g["bert"]["vocab_size"]
    zs = student._zs(False)
    gates = {}
    if zs is not None:
        gates = dict(head_z=torch.cat((zs["text_head_z"], zs["cross_head_z"]), dim=0),
                     mlp_z=torch.cat((zs["text_intermediate_z"], zs["cross_intermediate_z"]), dim=0))
    with torch.no_grad():
        enc = student.vision_encoder(image, head_z=zs["vision_head_z"] if zs else None, mlp_z=zs["vision_intermediate_z"] if zs else None)[0]
    prompt_ids = tok([student.prompt] * image.size(0), return_tensors="pt").input_ids[:, :-1]
    checked = 0
    for num_beams, max_length, min_length, rp in ((3, 12, 5, 1.0), (2, 9, 0, 1.0), (4, 14, 7, 1.3), (3, prompt_ids.shape[1] + 1, 0, 1.0)):
        caps, ids = student.generate(image, sample=False, num_beams=num_beams, max_length=max_length, min_length=min_length,
                                     repetition_penalty=rp, return_ids=True)
        enc_x = enc.repeat_interleave(num_beams, dim=0)

        def logp_of(rows):
            t = torch.tensor(rows, dtype=torch.long, device=device)
            with torch.no_grad():
                out = student.text_decoder(input_ids=t, attention_mask=torch.ones_like(t), encoder_hidden_states=enc_x[:t.shape[0]],
                                           encoder_attention_mask=None, is_decoder=True, return_dict=True, **gates)
            return torch.log_softmax(out.logits[:, -1, :].float(), dim=-1)

        want = beam_search(lambda rows: logp_of(rows).tolist(), [r for r in prompt_ids.tolist() for _ in range(num_beams)], num_beams,
                           max_length, min_length, tok.pad_token_id, tok.sep_token_id, vocab, repetition_penalty=rp)
        got = ids.tolist()
        assert len(got) == len(want) == image.size(0)
        for b, (x, y) in enumerate(zip(got, want)):
            if x == y:
                checked += 1
                continue
            assert not exact, (num_beams, max_length, min_length, rp, b, x, y)

            def seq_score(seq):                      # length-normalised sum of log-probabilities under the oracle's (cache-free) model
                seq = [t for t in seq]
                n = len(seq)
                while n > prompt_ids.shape[1] and seq[n - 1] == tok.pad_token_id:
                    n -= 1
                tot = 0.0
                for i in range(prompt_ids.shape[1], n):
                    tot += float(logp_of([seq[:i]])[0, seq[i]])
                return tot / max(1, n - 1)
            assert abs(seq_score(x) - seq_score(y)) <= 1e-2 * max(1.0, abs(seq_score(y))), (b, x, y, seq_score(x), seq_score(y))
        assert caps == [tok.decode(r, skip_special_tokens=True)[len(student.prompt):] for r in ids]
    if not exact:
        assert checked >= 1

