"""`accelerators` drop-in (reference: accelerators/{accelerator,apex_ddp_accelerator}.py).  apex AMP O1 + apex DDP are replaced by the
bf16 tensor-core kernels and the flat-arena NCCL gradient mean-allreduce of `efficientvlm_b200.optim.FlatAdamW`."""
