"""`ApexDDPAccelerator` with the reference's interface (accelerators/apex_ddp_accelerator.py:30-101) and none of apex:

  set_up          same seeding, device binding, NCCL process group and rank-0 parameter broadcast (one broadcast per flat parameter arena
                  instead of one per tensor); the model is returned as is — there is no DDP wrapper, the gradient mean-allreduce
                  (`delay_allreduce=True` semantics: once, after the whole backward) is issued by `FlatAdamW.step()`, one NCCL call per arena;
  backward_step   `loss.backward()` — no loss scaling: the kernels compute in bf16 with fp32 accumulation and fp32 master weights,
                  bf16 has fp32's exponent range (AMP O1's dynamic loss scale exists for fp16);
  optimizer_step  arms the global-norm clip for the `optimizer.step()` that follows it in the drivers (GeneralDistill.py:263-266,383-386):
                  `FlatAdamW.step()` clips AFTER its allreduce, which is the reference's order (apex reduces at the end of backward, then
                  `clip_grad_norm_`, then `step`).  Returns the norm of the previous step's gradients as a device scalar (no host sync;
                  the drivers ignore the return value).
"""
import os
import random
import sys

import numpy as np
import torch
import torch.distributed as distributed

from .accelerator import Accelerator


class ApexDDPAccelerator(Accelerator):
    def __init__(self, cfg, logger):
        super().__init__(cfg, logger)
        self.accelerator_rng_seed = self.cfg.RNG_SEED
        self.accelerator_syncbn = self.cfg.SYNCBN
        self.accelerator_fp16_opt_level = self.cfg.FP16_OPT_LEVEL
        self.accelerator_fp16_loss_scale = self.cfg.FP16_LOSS_SCALE

    def set_up(self, model, optimizer, lr_scheduler, local_rank, world_size, rank):
        if not torch.cuda.is_available():
            raise RuntimeError("efficientvlm_b200 has no CPU path: ApexDDPAccelerator.set_up needs a CUDA device")
        torch.backends.cudnn.benchmark = False
        random.seed(self.accelerator_rng_seed)
        np.random.seed(self.accelerator_rng_seed)
        torch.random.manual_seed(self.accelerator_rng_seed)
        torch.cuda.manual_seed_all(self.accelerator_rng_seed)
        master_address = os.environ.get("MASTER_ADDR", "127.0.0.1")
        master_port = int(os.environ.get("MASTER_PORT", 34171))
        torch.cuda.set_device(local_rank)
        model = model.cuda()
        if not distributed.is_initialized():
            distributed.init_process_group(backend="nccl", init_method="tcp://{}:{}".format(master_address, master_port),
                                           world_size=world_size, rank=rank)
            print(f"ApexDDPAccelerator distributed, size: {world_size}, rank: {rank}, local rank: {local_rank}")
            sys.stdout.flush()
        self.broadcast(model, optimizer)
        self.ddp_model = model
        return model, optimizer, lr_scheduler

    def broadcast(self, model, optimizer=None, src=0):
        owned = set()
        if optimizer is not None and hasattr(optimizer, "broadcast_parameters"):
            optimizer.broadcast_parameters(src)                          # the optimizer's parameters: one broadcast per arena
            owned = {p.data_ptr() for g in optimizer.param_groups for p in g["params"]}
        if distributed.is_initialized() and distributed.get_world_size() > 1:
            for v in model.state_dict().values():                       # buffers and anything the optimizer does not own
                if v.data_ptr() not in owned:
                    distributed.broadcast(v, src)

    def backward_step(self, loss, optimizer=None):
        loss.backward()

    def optimizer_step(self, optimizer, model, grad_norm):
        if hasattr(optimizer, "clip_grad_norm"):
            optimizer.clip_grad_norm = float(grad_norm)
            return optimizer.grad_norm()
        return float(torch.nn.utils.clip_grad_norm_(model.parameters(), grad_norm))
