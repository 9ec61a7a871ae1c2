"""efficient_models.nlvr_l0_module -> B200 implementation."""
from efficientvlm_b200.l0_module import NLVRL0Module, epsilon, limit_a, limit_b  # noqa: F401
