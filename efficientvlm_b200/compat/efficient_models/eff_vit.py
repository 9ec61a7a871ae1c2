"""efficient_models.eff_vit -> B200 implementation."""
from efficientvlm_b200.eff_vit import (CLIPAttention, CLIPEncoder, CLIPEncoderLayer, CLIPMLP, CLIPVisionTransformer,  # noqa: F401
                                       find_pruneable_heads_and_indices, prune_linear_layer)
