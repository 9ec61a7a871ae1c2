"""`optim` drop-in (reference: optim.py:4-69): same two functions, same name-based parameter groups, HF-AdamW update rule, on
`efficientvlm_b200.optim.FlatAdamW` (flat arenas, fused allreduce + global-norm clip + AdamW)."""
from efficientvlm_b200.optim import create_L0_optimizer, create_optimizer  # noqa: F401
