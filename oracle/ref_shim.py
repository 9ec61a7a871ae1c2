"""TEST INFRASTRUCTURE ONLY — import shim that lets the *unmodified* reference modules under
/root/reference run on this image (transformers 5.x, no timm/apex/pycocotools).

Used by oracle/make_golden.py (to generate tests/golden/*.pt) and by the CPU tests that validate the
oracle restatement against the real reference when /root/reference is present.  Nothing in the product
package imports this file.  Recipe follows SURVEY.md Appendix B.
"""
import json
import os
import sys
import tempfile
import types

import torch

REF_ROOT = os.environ.get("EVLM_REFERENCE_ROOT", "/root/reference")


def available():
    return os.path.isdir(os.path.join(REF_ROOT, "efficient_models"))


def _find_pruneable_heads_and_indices(heads, n_heads, head_size, already_pruned_heads):
    mask = torch.ones(n_heads, head_size)
    heads = set(heads) - already_pruned_heads
    for head in heads:
        head = head - sum(1 if h < head else 0 for h in already_pruned_heads)
        mask[head] = 0
    mask = mask.view(-1).contiguous().eq(1)
    index = torch.arange(len(mask))[mask].long()
    return heads, index


_installed = False


def install():
    """Patch the environment so `import efficient_models.*` / `import models.*` resolve to the reference."""
    global _installed
    if _installed:
        return
    if not available():
        raise RuntimeError("reference tree not found at %s" % REF_ROOT)
    import transformers.modeling_utils as mu
    import transformers.pytorch_utils as pu
    import transformers.file_utils as fu

    mu.apply_chunking_to_forward = pu.apply_chunking_to_forward
    mu.prune_linear_layer = pu.prune_linear_layer
    mu.find_pruneable_heads_and_indices = _find_pruneable_heads_and_indices
    fu.TF_RETURN_INTRODUCTION = ""
    for name in ("timm", "timm.models", "timm.models.vision_transformer", "timm.models.registry", "timm.models.layers"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["timm.models.vision_transformer"]._cfg = lambda **k: {}
    sys.modules["timm.models.vision_transformer"].PatchEmbed = torch.nn.Identity
    sys.modules["timm.models.registry"].register_model = lambda f: f
    L = sys.modules["timm.models.layers"]
    L.trunc_normal_ = torch.nn.init.trunc_normal_
    L.DropPath = torch.nn.Identity
    L.to_2tuple = lambda x: (x, x)
    for name in ("pycocotools", "pycocotools.coco", "pycocotools.cocoeval", "pycocoevalcap", "pycocoevalcap.eval"):
        sys.modules.setdefault(name, types.ModuleType(name))
    sys.modules["pycocotools.coco"].COCO = object
    sys.modules["pycocotools.cocoeval"].COCOeval = object
    sys.modules["pycocoevalcap.eval"].COCOEvalCap = object
    # the reference must win over any same-named drop-in package of ours
    if REF_ROOT in sys.path:
        sys.path.remove(REF_ROOT)
    sys.path.insert(0, REF_ROOT)
    for m in [k for k in sys.modules if k == "models" or k.startswith("models.") or k == "efficient_models"
              or k.startswith("efficient_models.") or k == "utils" or k.startswith("utils.")]:
        del sys.modules[m]
    import efficient_models.eff_bert as eb
    import models.xbert as xb

    for B in (eb, xb):
        B.BertPreTrainedModel.init_weights = lambda self: self.apply(self._init_weights)
        B.BertPreTrainedModel.get_head_mask = lambda self, head_mask, n, *a, **k: [None] * n
        # 4.12.5 semantics of ModuleUtilsMixin.invert_attention_mask for fp32: (1 - m) * -10000
        # (the installed 5.x uses finfo.min).  Identical results whenever the image mask is all ones.
    if not torch.distributed.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29533")
        torch.distributed.init_process_group("gloo", rank=0, world_size=1)
    _installed = True


def make_config_dir(vision_cfg, bert_cfg_dict):
    """Write a vision json + a text_encoder/config.json into a temp dir; returns (vision_json, text_dir)."""
    d = tempfile.mkdtemp(prefix="evlm_oracle_cfg_")
    vj = os.path.join(d, "vision.json")
    with open(vj, "w") as f:
        json.dump(vision_cfg, f)
    td = os.path.join(d, "text_encoder")
    os.makedirs(td)
    with open(os.path.join(td, "config.json"), "w") as f:
        json.dump(bert_cfg_dict, f)
    return vj, td
