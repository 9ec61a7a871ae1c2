"""models.clip_vit -> B200 implementation (gates default to None)."""
from efficientvlm_b200.eff_vit import (CLIPAttention, CLIPEncoder, CLIPEncoderLayer, CLIPMLP, CLIPVisionTransformer)  # noqa: F401
