"""Minimal stand-ins for the transformers `ModelOutput` dataclasses the reference returns
(BaseModelOutputWithPoolingAndCrossAttentions, MaskedLMOutput, CausalLMOutputWithCrossAttentions; eff_bert.py:688-694,
1155-1162,1436-1443,1708-1714): attribute access, `out["name"]`, and integer / slice indexing over the non-None fields."""


class ModelOutput:
    _fields = ()

    def __init__(self, **kw):
        for f in self._fields:
            setattr(self, f, kw.pop(f, None))
        if kw:
            raise TypeError("unexpected fields %s" % sorted(kw))

    def to_tuple(self):
        return tuple(getattr(self, f) for f in self._fields if getattr(self, f) is not None)

    def __getitem__(self, k):
        if isinstance(k, str):
            return getattr(self, k)
        return self.to_tuple()[k]

    def __iter__(self):
        return iter(self.to_tuple())

    def __len__(self):
        return len(self.to_tuple())

    def keys(self):
        return [f for f in self._fields if getattr(self, f) is not None]


class BaseModelOutputWithPastAndCrossAttentions(ModelOutput):
    _fields = ("last_hidden_state", "past_key_values", "hidden_states", "attentions", "cross_attentions")


class BaseModelOutputWithPoolingAndCrossAttentions(ModelOutput):
    _fields = ("last_hidden_state", "pooler_output", "past_key_values", "hidden_states", "attentions", "cross_attentions")


class MaskedLMOutput(ModelOutput):
    _fields = ("loss", "logits", "hidden_states", "attentions", "cross_attentions")


class CausalLMOutputWithCrossAttentions(ModelOutput):
    _fields = ("loss", "logits", "past_key_values", "hidden_states", "attentions", "cross_attentions")
