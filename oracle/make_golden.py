"""TEST INFRASTRUCTURE ONLY — generates tests/golden/*.pt by running the UNMODIFIED reference modules
(/root/reference, through oracle/ref_shim.py) on seeded synthetic inputs with tiny random-init models.

    python oracle/make_golden.py            # rewrites tests/golden/

The fixtures hold the inputs and the reference outputs (+ a few gradients); weights are rebuilt on both sides by
oracle/det_init.py (integer-hash init keyed by state_dict name), so the
oracle re-statement, and through it the CUDA product, can be checked without /root/reference being present
(it does not exist on the GPU box).  head_dim is 64 everywhere (the only head size the hot path uses).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_shim  # noqa: E402
from oracle.det_init import det_init_module_  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")

VIS = dict(image_res=32, patch_size=16, vision_width=128, hidden_act="quick_gelu", num_attention_heads=2,
           attention_dropout=0.0, intermediate_size=256, num_hidden_layers=2, local_attn_depth=1)
BERT = dict(vocab_size=211, hidden_size=128, num_hidden_layers=6, num_attention_heads=2, intermediate_size=256,
            max_position_embeddings=40, type_vocab_size=2, hidden_act="gelu", hidden_dropout_prob=0.1,
            attention_probs_dropout_prob=0.1, layer_norm_eps=1e-12, initializer_range=0.02, pad_token_id=0)


def cpu(t):
    if isinstance(t, (tuple, list)):
        return [cpu(x) for x in t]
    return t.detach().clone() if torch.is_tensor(t) else t


def save(name, obj):
    path = os.path.join(OUT, name + ".pt")
    torch.save(obj, path)
    print("wrote %s (%.1f KiB)" % (path, os.path.getsize(path) / 1024))


def spec(module):
    """state_dict names -> (shape, dtype): lets the tests rebuild weights AND check key compatibility."""
    return {k: (tuple(v.shape), str(v.dtype)) for k, v in module.state_dict().items()}


def rand_gate(shape, g):
    z = torch.rand(shape, generator=g)
    z[z < 0.25] = 0.0  # exact zeros: pruned heads / columns
    z[z > 0.85] = 1.0
    return z


def main():
    ref_shim.install()
    os.makedirs(OUT, exist_ok=True)
    from transformers import BertConfig
    import efficient_models.eff_vit as ev
    import efficient_models.eff_bert as eb

    g = torch.Generator().manual_seed(1234)

    # ------------------------------------------------------------------ ViT
    torch.manual_seed(0)
    vit = ev.CLIPVisionTransformer(VIS["image_res"], 16, 128, "quick_gelu", 2, 0.0, 256, 2, local_attn_depth=1).eval()
    det_init_module_(vit)
    x = torch.randn(3, 3, 32, 32, generator=g)
    head_z = rand_gate((2, 1, 2, 1, 1), g).requires_grad_()
    mlp_z = rand_gate((2, 1, 1, 256), g).requires_grad_()
    out, hid, att = vit(x, output_attentions=True, output_hidden_states=True, head_z=head_z, mlp_z=mlp_z)
    loss = out.pow(2).mean() + sum(a.pow(2).mean() for a in att) * 3.0
    names = ["class_embedding", "patch_embed.weight", "pos_embed.weight", "encoder.layers.0.self_attn.q_proj.weight",
             "encoder.layers.0.self_attn.q_proj.bias", "encoder.layers.1.mlp.fc1.weight", "encoder.layers.1.mlp.fc2.bias",
             "encoder.layers.0.layer_norm1.weight", "encoder.layers.1.self_attn.out_proj.weight",
             "encoder.layers.0.self_attn.v_proj.weight", "encoder.layers.0.self_attn.k_proj.bias"]
    params = dict(vit.named_parameters())
    grads = torch.autograd.grad(loss, [params[n] for n in names] + [head_z, mlp_z])
    out_ng, hid_ng, att_ng = vit(x, output_attentions=True, output_hidden_states=True)  # no gates
    # region batch (local attention, eff_vit.py:332-357): 3 images, 4 regions
    idx_to_group = torch.tensor([0, 2, 2, 1])
    image_atts = torch.tensor([[1, 1, 0, 0, 1], [1, 0, 1, 1, 1], [1, 1, 1, 0, 0], [1, 0, 0, 0, 1]], dtype=torch.float32)
    reg = vit(x, idx_to_group_img=idx_to_group, image_atts=image_atts, output_attentions=True, output_hidden_states=True)
    save("vit_tiny", dict(cfg=VIS, sd_spec=spec(vit), x=x, head_z=cpu(head_z), mlp_z=cpu(mlp_z), out=cpu(out),
                          hidden=cpu(hid), attn=cpu(att), loss=cpu(loss), grad_names=names + ["head_z", "mlp_z"],
                          grads=cpu(grads), out_nogate=cpu(out_ng), hidden_nogate=cpu(hid_ng), attn_nogate=cpu(att_ng),
                          idx_to_group=idx_to_group, image_atts=image_atts, region_out=cpu(reg[0]),
                          region_hidden=cpu(reg[1]), region_attn=cpu(reg[2]), region_full=cpu(reg[3])))

    # ------------------------------------------------------------------ BERT encoder, 3 modes with gates
    torch.manual_seed(1)
    cfg = BertConfig(**BERT)
    cfg.fusion_layer = 3
    cfg.encoder_width = 128
    bert = eb.BertModel(cfg, add_pooling_layer=False).eval()
    det_init_module_(bert)
    B, L, N = 3, 9, 5
    ids = torch.randint(1, BERT["vocab_size"], (B, L), generator=g)
    atts = torch.ones(B, L, dtype=torch.long)
    atts[1, 6:] = 0
    atts[2, 4:] = 0
    img = out.detach()  # [3,5,128]
    img_atts = torch.ones(B, N, dtype=torch.long)
    text_head_z = rand_gate((3, 1, 2, 1, 1), g).requires_grad_()
    text_mlp_z = rand_gate((3, 1, 1, 256), g).requires_grad_()
    cross_head_z = rand_gate((6, 1, 2, 1, 1), g).requires_grad_()
    cross_mlp_z = rand_gate((3, 1, 1, 256), g).requires_grad_()
    o_text = bert(ids, attention_mask=atts, return_dict=True, mode="text", output_attentions=True, output_hidden_states=True,
                  head_z=text_head_z, mlp_z=text_mlp_z)
    o_fus = bert(encoder_embeds=o_text.last_hidden_state, attention_mask=atts, encoder_hidden_states=img,
                 encoder_attention_mask=img_atts, return_dict=True, mode="fusion", output_attentions=True,
                 output_hidden_states=True, head_z=cross_head_z, mlp_z=cross_mlp_z)
    loss = o_fus.last_hidden_state[:, 0].pow(2).mean() + sum(a.pow(2).mean() for a in o_fus.cross_attentions) \
        + sum(a.pow(2).mean() for a in o_text.attentions)
    bnames = ["embeddings.word_embeddings.weight", "embeddings.LayerNorm.weight", "encoder.layer.0.attention.self.query.weight",
              "encoder.layer.1.intermediate.dense.weight", "encoder.layer.3.crossattention.self.key.weight",
              "encoder.layer.4.crossattention.output.dense.bias", "encoder.layer.5.output.LayerNorm.bias",
              "encoder.layer.5.output.dense.weight", "encoder.layer.2.attention.output.LayerNorm.weight"]
    bparams = dict(bert.named_parameters())
    bgrads = torch.autograd.grad(loss, [bparams[n] for n in bnames] + [text_head_z, text_mlp_z, cross_head_z, cross_mlp_z])
    # multi_modal with concatenated gates (quirk Q1)
    mm_head_z = torch.cat([text_head_z, cross_head_z]).detach()
    mm_mlp_z = torch.cat([text_mlp_z, cross_mlp_z]).detach()
    o_mm = bert(ids, attention_mask=atts, encoder_hidden_states=img, encoder_attention_mask=img_atts, return_dict=True,
                mode="multi_modal", output_attentions=True, output_hidden_states=True, head_z=mm_head_z, mlp_z=mm_mlp_z)
    o_mm_ng = bert(ids, attention_mask=atts, encoder_hidden_states=img, encoder_attention_mask=img_atts, return_dict=True,
                   mode="multi_modal", output_attentions=True, output_hidden_states=True)
    # NLVR-style list of two image tensors
    img2 = torch.randn(B, N, 128, generator=g)
    o_list = bert(ids, attention_mask=atts, encoder_hidden_states=[img, img2], encoder_attention_mask=[img_atts, img_atts],
                  return_dict=True, mode="multi_modal")
    save("bert_tiny", dict(cfg=dict(BERT, fusion_layer=3, encoder_width=128), sd_spec=spec(bert), ids=ids, atts=atts,
                           img=img, img_atts=img_atts, img2=img2, text_head_z=cpu(text_head_z), text_mlp_z=cpu(text_mlp_z),
                           cross_head_z=cpu(cross_head_z), cross_mlp_z=cpu(cross_mlp_z),
                           text_last=cpu(o_text.last_hidden_state), text_hidden=cpu(o_text.hidden_states),
                           text_attn=cpu(o_text.attentions), fus_last=cpu(o_fus.last_hidden_state),
                           fus_hidden=cpu(o_fus.hidden_states), fus_attn=cpu(o_fus.attentions),
                           fus_cross=cpu(o_fus.cross_attentions), loss=cpu(loss),
                           grad_names=bnames + ["text_head_z", "text_mlp_z", "cross_head_z", "cross_mlp_z"], grads=cpu(bgrads),
                           mm_last=cpu(o_mm.last_hidden_state), mm_hidden=cpu(o_mm.hidden_states), mm_attn=cpu(o_mm.attentions),
                           mm_cross=cpu(o_mm.cross_attentions), mm_nogate_last=cpu(o_mm_ng.last_hidden_state),
                           list_last=cpu(o_list.last_hidden_state)))

    # ------------------------------------------------------------------ MLM head + LM (decoder) head
    torch.manual_seed(2)
    mlm = eb.BertForMaskedLM(cfg).eval()
    det_init_module_(mlm)
    mlm.cls.predictions.decoder.weight = mlm.bert.embeddings.word_embeddings.weight  # tie (4.12.5 init_weights did)
    masked_pos = torch.tensor([[1, 3, 5, 0], [2, 4, 0, 0], [1, 2, 3, 0]])
    labels = torch.tensor([[5, 17, 99, -100], [7, 8, -100, -100], [200, 3, 1, -100]])
    mo = mlm(ids, attention_mask=atts, encoder_hidden_states=img, encoder_attention_mask=img_atts, return_dict=True,
             labels=labels, masked_pos=masked_pos, output_attentions=True, output_hidden_states=True)
    mg = torch.autograd.grad(mo.loss, [mlm.bert.embeddings.word_embeddings.weight, mlm.cls.predictions.bias,
                                       mlm.cls.predictions.transform.dense.weight])
    cfg_dec = BertConfig(**BERT)
    cfg_dec.fusion_layer = 3
    cfg_dec.encoder_width = 128
    dec = eb.BertLMHeadModel(cfg_dec, label_smoothing=0.1).eval()
    det_init_module_(dec)
    dec.cls.predictions.decoder.weight = dec.bert.embeddings.word_embeddings.weight
    dlabels = ids.masked_fill(atts == 0, -100)
    dlabels[:, :2] = -100
    dec_head_z = torch.cat([text_head_z, cross_head_z]).detach()
    dec_mlp_z = torch.cat([text_mlp_z, cross_mlp_z]).detach()
    do = dec(ids, attention_mask=atts, encoder_hidden_states=img, encoder_attention_mask=img_atts, labels=dlabels,
             return_dict=True, reduction="none", head_z=dec_head_z, mlp_z=dec_mlp_z)
    dec0 = eb.BertLMHeadModel(cfg_dec, label_smoothing=0.0).eval()
    dec0.load_state_dict(dec.state_dict())
    dec0.cls.predictions.decoder.weight = dec0.bert.embeddings.word_embeddings.weight
    do0 = dec0(ids, attention_mask=atts, encoder_hidden_states=img, encoder_attention_mask=img_atts, labels=dlabels,
               return_dict=True, reduction="mean")
    # incremental decode with KV cache: feed first 4 tokens, then token 5 with past
    step1 = dec0(ids[:, :4], attention_mask=torch.ones(B, 4, dtype=torch.long), encoder_hidden_states=img,
                 encoder_attention_mask=img_atts, return_dict=True, use_cache=True)
    step2 = dec0(ids[:, 4:5], attention_mask=torch.ones(B, 5, dtype=torch.long), encoder_hidden_states=img,
                 encoder_attention_mask=img_atts, return_dict=True, use_cache=True, past_key_values=step1.past_key_values)
    full5 = dec0(ids[:, :5], attention_mask=torch.ones(B, 5, dtype=torch.long), encoder_hidden_states=img,
                 encoder_attention_mask=img_atts, return_dict=True)
    save("heads_tiny", dict(mlm_sd_spec=spec(mlm), dec_sd_spec=spec(dec), masked_pos=masked_pos, labels=labels, mlm_loss=cpu(mo.loss),
                            mlm_logits=cpu(mo.logits), mlm_hidden=cpu(mo.hidden_states), mlm_grads=cpu(mg),
                            dlabels=dlabels, dec_head_z=dec_head_z, dec_mlp_z=dec_mlp_z,
                            dec_loss_none_ls=cpu(do.loss), dec_logits=cpu(do.logits), dec_loss_mean=cpu(do0.loss),
                            dec_logits_nogate=cpu(do0.logits), step2_logits=cpu(step2.logits), full5_logits=cpu(full5.logits)))

    # ------------------------------------------------------------------ L0 module
    torch.manual_seed(3)
    vj, td = ref_shim.make_config_dir(dict(VIS, patch_size=16), BERT)
    rcfg = dict(text_encoder=td, vision_config=vj, patch_size=16, image_res=32, use_clip_vit=True, use_swin=False,
                text_num_hidden_layers=6, embed_dim=64, temp=0.07, sparsity=0.3)
    from efficient_models.xvlm_l0_module import XVLML0Module
    l0 = XVLML0Module(rcfg, target_sparsity=0.3)
    l0.set_lagrangian_warmup_steps(100)
    with torch.no_grad():  # spread loga so that deterministic masks are non-trivial, keep some exact ties
        for k, la in l0.z_logas.items():
            la.copy_(torch.randn(la.shape, generator=g) * 2.0)
        l0.vision_int_loga[0, :8] = 0.25
        l0.lambda_1.fill_(0.7)
        l0.lambda_2.fill_(-0.3)
    eps = {k: torch.rand(la.shape, generator=g).clamp(1e-6, 1 - 1e-6) for k, la in l0.z_logas.items()}
    it = iter([eps[k] for k in l0.types])
    l0.get_eps = lambda size: next(it)
    zs_train = l0.forward(training=True)
    zs_eval = l0.forward(training=False)
    lag, es, ts = l0.lagrangian_regularization(37)
    ltot = lag + sum((z * torch.arange(z.numel()).view(z.shape) / z.numel()).sum() for z in zs_train.values())
    lgrads = torch.autograd.grad(ltot, [l0.z_logas[k] for k in l0.types] + [l0.lambda_1, l0.lambda_2])
    save("l0_tiny", dict(types=list(l0.types), logas={k: cpu(v) for k, v in l0.z_logas.items()}, eps=eps,
                         shapes={k: list(v) for k, v in l0.shapes.items()}, sizes=dict(l0.sizes),
                         params_per_dim=dict(l0.parameters_per_dim), prunable_model_size=l0.prunable_model_size,
                         zs_train={k: cpu(v) for k, v in zs_train.items()}, zs_eval={k: cpu(v) for k, v in zs_eval.items()},
                         zs_order=list(zs_train.keys()), lagrangian=cpu(lag), expected_sparsity=cpu(es), target_sparsity=ts,
                         step=37, warmup=100, lambda_1=0.7, lambda_2=-0.3, target=0.3, grads=cpu(lgrads),
                         model_size=l0.calculate_model_size(zs_eval)))

    # ------------------------------------------------------------------ retrieval model (config 1 shape, tiny) + KD losses
    torch.manual_seed(4)
    from efficient_models.model_retrieval import EffXVLMforRetrieval
    cwd = os.getcwd()
    os.chdir(ref_shim.REF_ROOT)
    model = EffXVLMforRetrieval(rcfg).eval()
    det_init_module_(model)
    os.chdir(cwd)
    with torch.no_grad():
        for k, la in model.l0_module.z_logas.items():
            la.copy_(torch.randn(la.shape, generator=g) * 1.5 + 1.0)
    B = 4
    image = torch.randn(B, 3, 32, 32, generator=g)
    text_ids = torch.randint(1, BERT["vocab_size"], (B, 9), generator=g)
    text_atts = torch.ones(B, 9, dtype=torch.long)
    text_atts[3, 5:] = 0
    idx = torch.tensor([10, 11, 10, 13])
    orig_multinomial = torch.multinomial
    torch.multinomial = lambda w, n, *a, **k: torch.argmax(w, dim=-1, keepdim=True)  # deterministic hard negative
    try:
        loss_itc, loss_itm = model(image, text_ids, text_atts, idx=idx)
        loss_itc0, loss_itm0 = model(image, text_ids, text_atts, idx=None)
        eps = {k: torch.rand(la.shape, generator=g).clamp(1e-6, 1 - 1e-6) for k, la in model.l0_module.z_logas.items()}
        it2 = iter([eps[k] for k in model.l0_module.types])
        model.l0_module.get_eps = lambda size: next(it2)
        res = model(image, text_ids, text_atts, idx=idx, output_attentions=True, output_hidden_states=True)
    finally:
        torch.multinomial = orig_multinomial
    tot = res["loss"]["loss_itc"] + res["loss"]["loss_itm"]
    rn = ["vision_encoder.encoder.layers.0.self_attn.q_proj.weight", "text_encoder.encoder.layer.4.crossattention.self.value.weight",
          "vision_proj.weight", "itm_head.0.weight", "temp", "l0_module.vision_head_loga", "l0_module.cross_int_loga"]
    rp = dict(model.named_parameters())
    rg = torch.autograd.grad(tot, [rp[n] for n in rn])
    import ast
    src = open(os.path.join(ref_shim.REF_ROOT, "GeneralDistill.py")).read()
    tree = ast.parse(src)
    fns = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("get_kd_loss", "soft_cross_entropy", "get_cor_teacher")]
    ns = {"torch": torch, "KLDivLoss": torch.nn.KLDivLoss}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "GeneralDistill.py", "exec"), ns)  # the reference's own code, unmodified
    get_kd_loss, soft_cross_entropy, get_cor_teacher = ns["get_kd_loss"], ns["soft_cross_entropy"], ns["get_cor_teacher"]
    mse = torch.nn.MSELoss()
    t_hidden = [torch.randn(2, 5, 16, generator=g) for _ in range(13)]
    s_hidden = [torch.randn(2, 5, 16, generator=g) for _ in range(7)]
    t_att = [torch.softmax(torch.randn(2, 2, 5, 5, generator=g), -1) for _ in range(12)]
    s_att = [torch.softmax(torch.randn(2, 2, 5, 5, generator=g), -1) for _ in range(6)]
    th = get_cor_teacher(t_hidden, s_hidden)
    ta = get_cor_teacher(t_att, s_att, is_attn=True)
    kd = dict(t_hidden=t_hidden, s_hidden=s_hidden, t_att=t_att, s_att=s_att,
              hid=cpu(get_kd_loss(s_hidden, th, False, mse, "cpu")), hid_img=cpu(get_kd_loss(s_hidden, th, False, mse, "cpu", is_img=True)),
              att=cpu(get_kd_loss(s_att, ta, True, mse, "cpu")))
    sl, tl = torch.randn(6, 50, generator=g) * 3, torch.randn(6, 50, generator=g) * 3
    kd.update(s_logits=sl, t_logits=tl, kl=cpu(soft_cross_entropy(sl / 2.0, tl / 2.0)))
    save("retrieval_tiny", dict(cfg=dict(rcfg, text_encoder=None, vision_config=None), sd_spec=spec(model), vis=VIS, bert=BERT,
                                l0_logas={k: cpu(v) for k, v in model.l0_module.z_logas.items()}, image=image, text_ids=text_ids, text_atts=text_atts, idx=idx,
                                loss_itc=cpu(loss_itc), loss_itm=cpu(loss_itm), loss_itc_noidx=cpu(loss_itc0),
                                loss_itm_noidx=cpu(loss_itm0), eps=eps, kd_loss_itc=cpu(res["loss"]["loss_itc"]),
                                kd_loss_itm=cpu(res["loss"]["loss_itm"]), kd_itm_logits=cpu(res["logits_dict"]["itm_head_logits"]),
                                kd_image_hidden=cpu(res["hidden_dict"]["image_hidden_states"]),
                                kd_text_attn=cpu(res["attention_dict"]["text_attentions"]),
                                kd_neg_cross=cpu(res["cross_attention_dict"]["itm_neg_cross_attentions"]),
                                kd_neg_hidden=cpu(res["hidden_dict"]["itm_neg_hidden_states"]),
                                grad_names=rn, grads=cpu(rg), kd=kd))
    print("done")


if __name__ == "__main__":
    main()
