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


def assert_close(a, b, tol, what=""):
    e = rel_err(a, b)
    assert e <= tol, "%s: relative error %.3e > %.1e" % (what, e, tol)


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
