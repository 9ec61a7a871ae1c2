"""TEST INFRASTRUCTURE ONLY — platform-independent deterministic parameter initialisation.

Golden fixtures do not store model weights (that would be megabytes per fixture); instead both
oracle/make_golden.py (which feeds the reference) and the tests (which feed the oracle / the CUDA product)
rebuild the same weights from this integer-hash generator, keyed by the parameter's state_dict name.
"""
import zlib

import numpy as np
import torch


def det_uniform(name, shape, lo=-1.0, hi=1.0):
    n = int(np.prod(shape)) if len(shape) else 1
    h = np.uint64(zlib.crc32(name.encode()))
    i = np.arange(n, dtype=np.uint64)
    x = (i * np.uint64(2654435761) + h * np.uint64(40503) + np.uint64(12345)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    x = (x * np.uint64(0x85EBCA6B)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(13)
    x = (x * np.uint64(0xC2B2AE35)) & np.uint64(0xFFFFFFFF)
    x ^= x >> np.uint64(16)
    u = x.astype(np.float64) / 4294967296.0
    return torch.from_numpy((lo + (hi - lo) * u).astype(np.float32)).reshape(tuple(shape))


def det_state_dict(sd):
    """Returns a new dict with every floating-point entry of `sd` (name -> tensor) re-initialised."""
    out = {}
    for orig_name, t in sd.items():
        name = orig_name
        if not torch.is_floating_point(t) or name.endswith("temp") or "loga" in name or "lambda" in name:
            out[name] = t.clone()
            continue
        shape = tuple(t.shape)
        # `cls.predictions.decoder.bias` is the same Parameter as `cls.predictions.bias` (eff_bert.py:738-741)
        name = name.replace("predictions.decoder.bias", "predictions.bias")
        if name.endswith("bias"):
            v = det_uniform(name, shape, -0.1, 0.1)
        elif t.dim() == 1 and name.endswith("weight"):        # LayerNorm gains
            v = det_uniform(name, shape, 0.8, 1.2)
        elif t.dim() <= 1:                                     # class_embedding
            v = det_uniform(name, shape, -0.5, 0.5)
        elif "embeddings" in name or "pos_embed" in name:      # embedding tables
            v = det_uniform(name, shape, -0.5, 0.5)
        else:                                                  # linear / conv weights: std ~ 1/sqrt(fan_in)
            fan_in = int(np.prod(shape[1:]))
            s = 1.7 / np.sqrt(fan_in)
            v = det_uniform(name, shape, -s, s)
        out[orig_name] = v.to(t.dtype)
    return out


def det_init_module_(module):
    sd = det_state_dict(module.state_dict())
    module.load_state_dict(sd)
    return module
