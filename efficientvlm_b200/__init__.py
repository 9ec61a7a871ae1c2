"""efficientvlm_b200 — B200-native (sm_100a) implementation of EfficientVLM's data-parallel hot path:
the L0-masked X-VLM transformer step behind the reference's own Python module API.

    eff_vit      CLIPVisionTransformer ...            (reference: efficient_models/eff_vit.py, models/clip_vit.py)
    eff_bert     BertModel / BertForMaskedLM / ...    (reference: efficient_models/eff_bert.py, models/xbert.py)
    xvlm         XVLMBase, AllGather, build_mlp ...   (reference: efficient_models/xvlm.py, models/xvlm.py)
    l0_module    XVLML0Module / VQAL0Module / NLVR... (reference: efficient_models/*_l0_module.py)
    distill      KD losses, XVLM (pretrain), EffXVLMforRetrieval, GD loss mix
    optim / ddp  flat-arena AdamW + gradient allreduce (reference: optim.py, accelerators/apex_ddp_accelerator.py)
    ops/kernels  autograd functions and ctypes bindings over libevlm_b200.so (include/evlm.h)

No CPU fallback exists: running any forward without the built library and a CUDA device raises.
"""
__version__ = "0.1.0"
