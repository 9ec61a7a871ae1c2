"""Checkpoint key surgery of the task models, as data: every task model turns a pre-training checkpoint (a
`models/model_pretrain.py::XVLM` state_dict, `text_encoder.bert.encoder.layer.N...`) into its own layout with ONE renaming rule, a pure
function `key -> None | (new names, drop the original?)`, applied by `remap_keys`; `load_into` then does what every reference
`load_pretrained` method ends with (`load_state_dict(strict=False)` + the three report lines).

Rules (same results as the reference methods — `tests/test_dropin.py`, case `load_pretrained`, loads the same checkpoint through the
reference classes' own methods and through these and compares the resulting state_dicts):
  * `vqa_decoder_rule`      efficient_models/model_generation.py:52-96 — the answer decoder starts from the fusion layers of the
                            pre-trained text encoder (layer N -> decoder layer N - num_text_layers); the text encoder itself is loaded
                            from `bert.`-less copies of the same tensors;
  * `nlvr_two_image_rule`   efficient_models/model_nlvr.py:150-185 — every fusion layer initialises both of its per-image copies;
  * `caption_decoder_rule`  efficient_models/model_generation.py:322-343 — the whole text encoder becomes the caption decoder.
"""


def remap_keys(state_dict, rule):
    """Applies `rule` to every key present when the call starts (entries created by the remapping are not revisited)."""
    for key in list(state_dict.keys()):
        decision = rule(key)
        if decision is None:
            continue
        new_names, drop = decision
        tensor = state_dict[key]
        for name in new_names:
            state_dict[name] = tensor
        if drop:
            del state_dict[key]
    return state_dict


def load_into(model, state_dict, ckpt_rpath):
    msg = model.load_state_dict(state_dict, strict=False)
    print("load checkpoint from %s" % ckpt_rpath)
    print("missing_keys: ", [p for p in msg.missing_keys if "vision_encoder" not in p])
    print("unexpected_keys: ", msg.unexpected_keys)
    return msg


def _layer_index(key, position):
    """Layer number of a `...layer.N...` key, read at the fixed dotted position the reference reads it from."""
    return int(key.split(".")[position])


def _with_layer(key, position, index):
    parts = key.split(".")
    parts[position] = str(index)
    return ".".join(parts)


def vqa_decoder_rule(num_text_layers, drop_cross_kv):
    """drop_cross_kv: the decoder's cross-attention K/V read question states of another width than the image tokens the pre-trained
    K/V were trained on (those projections then start from scratch: `init_params`)."""
    def rule(key):
        names = [key.replace("bert.", "")] if "bert." in key else []
        if "text_encoder." not in key:
            return (names, False) if names else None
        if "layer." in key:
            n = _layer_index(key, 4)                                   # text_encoder.bert.encoder.layer.N....
            is_cross_kv = ("crossattention.self.key" in key) or ("crossattention.self.value" in key)
            if n < num_text_layers or (drop_cross_kv and is_cross_kv):
                return names, True                                     # text-only layers / re-initialised projections: no decoder copy
            source = _with_layer(key, 4, n - num_text_layers)
        else:
            source = key                                               # embeddings, MLM head
        return names + [source.replace("text_encoder", "text_decoder")], True
    return rule


def nlvr_two_image_rule(num_text_layers):
    def rule(key):
        if "text_encoder." not in key or not (("bert." in key) or ("roberta." in key)):
            return None
        plain = key.replace("bert.", "").replace("roberta.", "")
        if "layer." in plain:
            n = _layer_index(plain, 3)                                 # text_encoder.encoder.layer.N....
            if n >= num_text_layers:
                first = (n - num_text_layers) * 2 + num_text_layers
                return [_with_layer(plain, 3, first), _with_layer(plain, 3, first + 1)], True
        return [plain], True
    return rule


def caption_decoder_rule():
    def rule(key):
        if not key.startswith("text_encoder."):
            return None
        return [key.replace("text_encoder.", "text_decoder.")], True
    return rule
