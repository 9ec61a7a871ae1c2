"""TEST INFRASTRUCTURE ONLY — generates tests/golden/itr_eval_tiny.pt: the ITR re-rank evaluation of `Eff_Retrieval.py:216-332`
(`evaluation`: deterministic L0 masks, text / image features, similarity top-k, k_test fusion passes per query, ITM score matrices)
and `itm_eval` (`:335-378`, recall@1/5/10), both extracted with `ast` from the UNMODIFIED driver and run on the UNMODIFIED reference
student `efficient_models/model_retrieval.py::EffXVLMforRetrieval` — once as a single process and once per rank of a 2-process job
(the reference's `size // num_tasks + 1` row split).

    python oracle/make_golden_itr_eval.py
"""
import ast
import datetime
import os
import sys
import time
import types

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle.det_init import det_init_module_  # noqa: E402
from oracle.fake_tokenizer import FakeTokenizer  # noqa: E402
from oracle.make_golden import BERT, VIS, cpu, save, spec  # noqa: E402

WORDS = "a the red blue small large dog cat bird horse runs sits jumps eats on under near grass table water sky two three people".split()


def make_texts(n, g):
    out = []
    for _ in range(n):
        m = int(torch.randint(3, 9, (1,), generator=g))
        out.append(" ".join(WORDS[int(i)] for i in torch.randint(0, len(WORDS), (m,), generator=g)))
    return out


class Dataset:
    def __init__(self, images, texts):
        self.image, self.text = list(range(len(images))), texts


class Loader:
    """The slice of the DataLoader surface `evaluation` touches: iteration over (image batch, ids) and `.dataset.{text,image}`."""

    def __init__(self, images, texts, bs):
        self.images, self.bs, self.dataset = images, bs, Dataset(images, texts)

    def __iter__(self):
        for i in range(0, len(self.images), self.bs):
            yield self.images[i:i + self.bs], torch.arange(i, min(len(self.images), i + self.bs))


class Logger:
    def __init__(self, delimiter=""):
        pass

    def log_every(self, it, freq, header=None):
        return it


def main():
    ref_shim.install()
    g = torch.Generator().manual_seed(77)
    vj, td = ref_shim.make_config_dir(dict(VIS, local_attn_depth=0), BERT)
    scfg = dict(text_encoder=td, vision_config=vj, patch_size=16, image_res=32, use_clip_vit=True, use_swin=False,
                text_num_hidden_layers=6, embed_dim=64, temp=0.07, sparsity=0.25)
    cwd = os.getcwd()
    os.chdir(ref_shim.REF_ROOT)
    from efficient_models.model_retrieval import EffXVLMforRetrieval
    torch.manual_seed(8)
    student = EffXVLMforRetrieval(scfg).eval()
    os.chdir(cwd)
    det_init_module_(student)
    with torch.no_grad():
        for k, la in student.l0_module.z_logas.items():
            la.copy_(torch.randn(la.shape, generator=g) * 1.5 + 1.0)
    n_img, per = 7, 3
    images = torch.randn(n_img, 3, 32, 32, generator=g)
    texts = make_texts(n_img * per, g)
    img2txt = {i: list(range(i * per, (i + 1) * per)) for i in range(n_img)}
    txt2img = {t: t // per for t in range(n_img * per)}
    config = dict(batch_size_test_text=8, max_tokens=9, k_test=4)
    tok = FakeTokenizer(BERT["vocab_size"])

    src = open(os.path.join(ref_shim.REF_ROOT, "Eff_Retrieval.py")).read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("evaluation", "itm_eval")]
    state = {"rank": 0, "world": 1}
    utils = types.SimpleNamespace(MetricLogger=Logger, get_world_size=lambda: state["world"], get_rank=lambda: state["rank"])
    ns = {"torch": torch, "np": np, "time": time, "datetime": datetime, "utils": utils, "dist": torch.distributed,
          "args": types.SimpleNamespace(distributed=False)}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "Eff_Retrieval.py", "exec"), ns)
    evaluation, itm_eval = ns["evaluation"], ns["itm_eval"]

    def tokenizer(text, **kw):   # the reference pads to max_length; the whitespace tokenizer pads to the longest row
        enc = tok(text, **kw)
        L = config["max_tokens"]
        ids = torch.zeros(len(text), L, dtype=torch.long)
        att = torch.zeros(len(text), L, dtype=torch.long)
        ids[:, :enc.input_ids.shape[1]] = enc.input_ids
        att[:, :enc.attention_mask.shape[1]] = enc.attention_mask
        return type(enc)(ids, att)

    loader = Loader(images, texts, 3)
    s_i2t, s_t2i, sparsity = evaluation(student, loader, tokenizer, "cpu", config)
    result = itm_eval(s_i2t, s_t2i, txt2img, img2txt)
    per_rank = []
    for r in range(2):
        state.update(rank=r, world=2)
        a, b, _ = evaluation(student, loader, tokenizer, "cpu", config)
        per_rank.append((torch.from_numpy(a), torch.from_numpy(b)))
    state.update(rank=0, world=1)
    # the similarity matrix the top-k runs on (not returned by the reference function): same calls, same order
    with torch.no_grad():
        zs = student.l0_module.forward(training=False)
        enc = tokenizer(texts, padding="max_length", truncation=True, max_length=config["max_tokens"], return_tensors="pt")
        tf = student.get_text_embeds(enc.input_ids, enc.attention_mask, head_z=zs["text_head_z"], head_layer_z=None, mlp_z=zs["text_intermediate_z"])
        vf, _ = student.get_vision_embeds(images, head_z=zs["vision_head_z"], head_layer_z=None, mlp_z=zs["vision_intermediate_z"])
        sims = student.get_features(image_embeds=vf) @ student.get_features(text_embeds=tf).t()
    save("itr_eval_tiny", dict(
        scfg=dict(scfg, text_encoder=None, vision_config=None), vis=dict(VIS, local_attn_depth=0), bert=BERT, s_sd_spec=spec(student),
        l0_logas={k: cpu(v) for k, v in student.l0_module.z_logas.items()}, images=images, texts=texts, text_ids=enc.input_ids,
        text_atts=enc.attention_mask, config=config, image_batch=3, img2txt=img2txt, txt2img=txt2img, sims=cpu(sims),
        score_i2t=torch.from_numpy(s_i2t), score_t2i=torch.from_numpy(s_t2i), sparsity=float(sparsity), result=result,
        per_rank=per_rank))
    print(result)


if __name__ == "__main__":
    main()
