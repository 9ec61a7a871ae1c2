"""efficient_models.generation_l0_module -> B200 implementation."""
from efficientvlm_b200.l0_module import VQAL0Module, epsilon, limit_a, limit_b  # noqa: F401
