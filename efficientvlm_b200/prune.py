"""Mask materialisation: turn deterministic L0 gates into physically smaller weights (ragged head counts / FFN widths per layer).

Drop-in for `update_params` + `prune_model_with_z` of the reference's `utils/xvlm_utils.py:37-245` (retrieval / NLVR / grounding
models), `utils/vqa_utils.py:37-293` and `utils/caption_utils.py` (models with a `text_decoder`): same two-step recipe —
  1. `update_params`: fold the gate VALUES into the weights the gate multiplies (value projection rows + bias for head gates,
     FFN down-projection columns for intermediate gates), so that non-binary gates keep their scale;
  2. `prune_model_with_z`: `prune_heads` every head whose gate is exactly 0 and slice away the FFN columns whose gate is 0.
The layer counts come from the gate tensors instead of the literals 6 / 3 / 3 / 3 of the reference files.

The pruned modules run on the same kernels (tests/test_host_logic.py::test_materialised_vqa_model_matches_gated_networks): the
GEMMs accept any N / K that is a multiple of 8 elements after padding the bf16 shadow pitch, attention takes `num_heads` from the
module.  Like the reference, a layer whose heads or FFN columns are ALL pruned is not representable by the forward code
(the reference sets the Linear to None and then fails in forward); this module refuses to produce such a layer.
"""
import torch

from . import ops
from .eff_vit import prune_linear_layer

DIMS_PER_HEAD = 64


def _bert_of(model, attr):
    m = getattr(model, attr, None)
    if m is None:
        return None
    return m.bert if hasattr(m, "bert") else m


def _rows(z):
    """[layers, ...] gate tensor -> list of flat per-layer CPU vectors."""
    return [z[i].detach().cpu().reshape(-1).clone() for i in range(z.shape[0])]


def _scale_value(att_self, head_z):
    hz = torch.repeat_interleave(head_z, DIMS_PER_HEAD).to(att_self.value.weight.device)
    att_self.value.weight.data = att_self.value.weight.data.mul(hz[:, None])
    att_self.value.bias.data = att_self.value.bias.data.mul(hz)


@torch.no_grad()
def update_params(model, zs):
    """vqa_utils.py:37-104 / xvlm_utils.py:37-85: gate values folded into value projections and FFN down-projections."""
    vision = model.vision_encoder
    text = _bert_of(model, "text_encoder")
    decoder = _bert_of(model, "text_decoder")
    n_text = zs["text_head_z"].shape[0] if "text_head_z" in zs else (zs["text_intermediate_z"].shape[0] if "text_intermediate_z" in zs else 0)
    if "vision_intermediate_z" in zs:
        for layer, z in enumerate(_rows(zs["vision_intermediate_z"])):
            fc2 = vision.encoder.layers[layer].mlp.fc2
            fc2.weight.data = fc2.weight.data.mul(z.to(fc2.weight.device))
    if "vision_head_z" in zs:
        for layer, z in enumerate(_rows(zs["vision_head_z"])):
            att = vision.encoder.layers[layer].self_attn
            hz = torch.repeat_interleave(z, DIMS_PER_HEAD).to(att.v_proj.weight.device)
            att.v_proj.weight.data = att.v_proj.weight.data.mul(hz[:, None])
            att.v_proj.bias.data = att.v_proj.bias.data.mul(hz)
    if "text_intermediate_z" in zs:
        for layer, z in enumerate(_rows(zs["text_intermediate_z"])):
            d = text.encoder.layer[layer].output.dense
            d.weight.data = d.weight.data.mul(z.to(d.weight.device))
    if "text_head_z" in zs:
        for layer, z in enumerate(_rows(zs["text_head_z"])):
            _scale_value(text.encoder.layer[layer].attention.self, z)
    if "cross_intermediate_z" in zs:
        for layer, z in enumerate(_rows(zs["cross_intermediate_z"])):
            d = text.encoder.layer[n_text + layer].output.dense
            d.weight.data = d.weight.data.mul(z.to(d.weight.device))
    if "cross_head_z" in zs:
        rows = _rows(zs["cross_head_z"])
        for layer in range(len(rows) // 2):
            blk = text.encoder.layer[n_text + layer]
            _scale_value(blk.attention.self, rows[2 * layer])
            _scale_value(blk.crossattention.self, rows[2 * layer + 1])
    if decoder is not None and "decoder_intermediate_z" in zs:
        for layer, z in enumerate(_rows(zs["decoder_intermediate_z"])):
            d = decoder.encoder.layer[layer].output.dense
            d.weight.data = d.weight.data.mul(z.to(d.weight.device))
    if decoder is not None and "decoder_head_z" in zs:
        rows = _rows(zs["decoder_head_z"])
        for layer in range(len(rows) // 2):
            blk = decoder.encoder.layer[layer]
            _scale_value(blk.attention.self, rows[2 * layer])
            _scale_value(blk.crossattention.self, rows[2 * layer + 1])
    ops.invalidate_weight_cache()


def _heads_to_prune(z, what):
    out = {}
    for layer, row in enumerate(_rows(z)):
        idx = torch.where(row == 0)[0].tolist()
        if len(idx) == row.numel():
            raise NotImplementedError("%s layer %d: every head is pruned; the forward code (reference and B200) cannot run a layer "
                                      "without heads" % (what, layer))
        out[layer] = idx
    return out


def _kept_dims(z, what):
    out = {}
    for layer, row in enumerate(_rows(z)):
        keep = row.nonzero().reshape(-1).tolist()
        if not keep:
            raise NotImplementedError("%s layer %d: every FFN column is pruned; the forward code cannot run a layer without an FFN" % (what, layer))
        out[layer] = keep
    return out


def prune_intermediate_layers(bert, keep_dims, device=None):
    """vqa_utils.py:295-303."""
    for layer, keep in keep_dims.items():
        blk = bert.encoder.layer[layer]
        index = torch.LongTensor(keep).to(blk.intermediate.dense.weight.device)
        blk.intermediate.dense = prune_linear_layer(blk.intermediate.dense, index=index, dim=0)
        blk.output.dense = prune_linear_layer(blk.output.dense, index=index, dim=1)


def prune_vision_intermediate_layers(vision_encoder, keep_dims, device=None):
    """vqa_utils.py:305-313."""
    for layer, keep in keep_dims.items():
        mlp = vision_encoder.encoder.layers[layer].mlp
        index = torch.LongTensor(keep).to(mlp.fc1.weight.device)
        mlp.fc1 = prune_linear_layer(mlp.fc1, index=index, dim=0)
        mlp.fc2 = prune_linear_layer(mlp.fc2, index=index, dim=1)


@torch.no_grad()
def prune_model_with_z(zs, model):
    """vqa_utils.py:107-293 / xvlm_utils.py:88-245: physical head and FFN-column pruning from exact zeros in the gates."""
    if zs is None:
        return None, None
    vision = model.vision_encoder
    text = _bert_of(model, "text_encoder")
    decoder = _bert_of(model, "text_decoder")
    if "vision_head_z" in zs:
        vision.prune_heads(_heads_to_prune(zs["vision_head_z"], "vision"))
    if "text_head_z" in zs:
        text.prune_heads(_heads_to_prune(zs["text_head_z"], "text"))
    if "cross_head_z" in zs:
        text.prune_heads(_heads_to_prune(zs["cross_head_z"], "cross"), is_cross="cross")
    if decoder is not None and "decoder_head_z" in zs:
        decoder.prune_heads(_heads_to_prune(zs["decoder_head_z"], "decoder"), is_cross="decoder")
    if "vision_intermediate_z" in zs:
        prune_vision_intermediate_layers(vision, _kept_dims(zs["vision_intermediate_z"], "vision"))
    if "text_intermediate_z" in zs and "cross_intermediate_z" in zs:
        both = torch.cat((zs["text_intermediate_z"], zs["cross_intermediate_z"]), dim=0)
        prune_intermediate_layers(text, _kept_dims(both, "text/cross"))
    if decoder is not None and "decoder_intermediate_z" in zs:
        prune_intermediate_layers(decoder, _kept_dims(zs["decoder_intermediate_z"], "decoder"))
    ops.invalidate_weight_cache()
    return None, None


def materialize(model, zs=None):
    """update_params + prune_model_with_z with the model's own deterministic masks (l0_module.forward(training=False)) by
    default.  Returns the per-module sizes `l0_module.calculate_model_size` reports.  The pruned model is then run WITHOUT gates
    (the reference's `fake_forward`, model_generation.py:214-230): see `efficientvlm_b200.vqa.XVLMForVQA.forward(..., use_gates=False)`."""
    if zs is None:
        with torch.no_grad():
            zs = model.l0_module.forward(training=False)
    update_params(model, zs)
    prune_model_with_z(zs, model)
    return model.l0_module.calculate_model_size(zs) if hasattr(model, "l0_module") else None
