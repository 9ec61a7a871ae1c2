"""efficient_models.xvlm -> B200 implementation (gated flavour)."""
from efficientvlm_b200.xvlm import (AllGather, XVLMBase, allgather, build_mlp, build_text_encoder, build_vision_encoder,  # noqa: F401
                                    interpolate_pos_embed, load_params_choose_layers, load_pretrained, read_json)
