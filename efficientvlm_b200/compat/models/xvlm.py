"""models.xvlm -> B200 implementation (un-gated flavour: get_vision_embeds always returns the KD 4-tuple)."""
from efficientvlm_b200.distill import XVLMBaseUngated as XVLMBase  # noqa: F401
from efficientvlm_b200.xvlm import (AllGather, allgather, build_mlp, build_text_encoder, build_vision_encoder,  # noqa: F401
                                    interpolate_pos_embed, load_params_choose_layers, load_pretrained, read_json)
