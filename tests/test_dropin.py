"""Drop-in boundary test (CPU; needs the reference tree, skipped on the GPU box where /root/reference does not exist):
the reference's OWN, unmodified task-model files (`efficient_models/model_retrieval.py`, `models/model_pretrain.py`) are imported
on top of our `efficient_models.xvlm` / `models` packages (efficientvlm_b200/compat first on sys.path) and must reproduce the
reference-generated goldens.  Arithmetic comes from the test-only torch op backend (tests/ref_ops.py)."""
import os
import subprocess
import sys

import pytest

REF = os.environ.get("EVLM_REFERENCE_ROOT", "/root/reference")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
pytestmark = pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "efficient_models")), reason="reference tree not present")

SCRIPT = r'''
import os, sys, types
ROOT, REF = sys.argv[1], sys.argv[2]
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "efficientvlm_b200", "compat"))   # our drop-in packages win ...
sys.path.append(REF)                                                    # ... the reference supplies everything else
import torch
from tests import ref_ops
from tests.helpers import load_golden, sd_from_spec, rel_err
import efficientvlm_b200.ops as ops, efficientvlm_b200.kernels as K

class MP:
    def setattr(self, obj, name, val): setattr(obj, name, val)
ref_ops.install(MP())
# the reference's task model, unmodified, on top of OUR XVLMBase / L0 module
import efficient_models.model_retrieval as mr
assert mr.__file__.startswith(REF), mr.__file__
import efficient_models.xvlm as ex
assert "efficientvlm_b200" in ex.__file__, ex.__file__
import efficientvlm_b200.eff_bert as eb
g = load_golden("retrieval_tiny")
cfg = dict(g["cfg"]); cfg["vision_config"] = dict(g["vis"]); cfg["text_encoder"] = None
orig = eb.BertConfig.__init__
def patched(self, **kw):
    m = dict(g["bert"]); m.update(kw); orig(self, **m)
eb.BertConfig.__init__ = patched
model = mr.EffXVLMforRetrieval(cfg).eval()
eb.BertConfig.__init__ = orig
sd = sd_from_spec(g["sd_spec"])
names = {"vision_head": "vision_head_loga", "text_head": "text_head_loga", "cross_head": "cross_head_loga",
         "vision_intermediate": "vision_int_loga", "text_intermediate": "text_int_loga", "cross_intermediate": "cross_int_loga"}
for k, v in g["l0_logas"].items():
    sd["l0_module." + names[k]] = v
model.load_state_dict(sd, strict=True)
from oracle import xvlm_oracle as O
def sampler(image_feat, text_feat, idx=None):
    w_i2t, w_t2i = O.itm_negative_weights(image_feat.detach(), text_feat.detach(), model.temp.detach(), idx)
    return w_t2i.argmax(1), w_i2t.argmax(1)
model.sample_itm_negatives = sampler
loss_itc, loss_itm = model(g["image"], g["text_ids"], g["text_atts"], idx=g["idx"])
e1, e2 = rel_err(loss_itc, g["loss_itc"]), rel_err(loss_itm, g["loss_itm"])
print("retrieval via reference task model: itc err %.2e itm err %.2e" % (e1, e2))
assert e1 < 2e-5 and e2 < 1e-4
# models/model_pretrain.py (GD teacher / student) imports `from models import XVLMBase`
import models
assert "efficientvlm_b200" in models.xvlm.__file__
import models.model_pretrain as mp
assert mp.__file__.startswith(REF)
assert issubclass(mp.XVLM, models.XVLMBase)
print("OK")
'''


def test_reference_task_models_run_on_our_core():
    r = subprocess.run([sys.executable, "-c", SCRIPT, ROOT, REF], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "OK" in r.stdout
