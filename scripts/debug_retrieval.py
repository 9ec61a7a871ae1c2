"""GPU debugging aid: walks the tiny retrieval fixture through the CUDA path printing NaN / error checkpoints."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tests.helpers import load_golden, rel_err  # noqa: E402
from tests.test_host_logic import _argmax_negatives, _retrieval_model  # noqa: E402


def chk(name, t):
    t = t.detach().float()
    print("%-28s shape=%-22s nan=%d max|x|=%.4g" % (name, tuple(t.shape), int(torch.isnan(t).sum()), t.abs().max().item()), flush=True)


g = load_golden("retrieval_tiny")
model = _retrieval_model(g).cuda()
model.sample_itm_negatives = _argmax_negatives(model)
image, text_ids, text_atts, idx = (g[k].cuda() for k in ("image", "text_ids", "text_atts", "idx"))
zs = model.l0_module(training=False)
for k, v in zs.items():
    chk("z " + k, v)
ie, ia = model.get_vision_embeds(image, head_z=zs["vision_head_z"], mlp_z=zs["vision_intermediate_z"])
chk("image_embeds", ie)
te = model.get_text_embeds(text_ids, text_atts, head_z=zs["text_head_z"], mlp_z=zs["text_intermediate_z"])
chk("text_embeds", te)
f_i, f_t = model.get_features(ie, te)
chk("image_feat", f_i)
chk("text_feat", f_t)
chk("temp", model.temp)
from efficientvlm_b200 import ops  # noqa: E402
lg = ops.sim_over_temp(f_i, f_t, model.temp)
chk("logits", lg)
print("logits ref err", rel_err(lg, (f_i @ f_t.t() / model.temp)))
loss = model.get_contrastive_loss(f_i, f_t, idx=idx)
chk("itc(idx)", loss)
loss = model.get_contrastive_loss(f_i, f_t, idx=None)
chk("itc", loss)
print("golden itc", g["loss_itc"].item(), g["loss_itc_noidx"].item())
itm = model.get_matching_loss(ie, ia, f_i, te, text_atts, f_t, idx=idx, head_z=zs["cross_head_z"], mlp_z=zs["cross_intermediate_z"])
chk("itm", itm)
print("golden itm", g["loss_itm"].item())
