"""accelerators/accelerator.py:15-33 — the base class the drivers' accelerator objects derive from."""
import torch


class Accelerator:
    def __init__(self, cfg, logger):
        self.cfg = cfg
        self.logger = logger

    def set_up(self, model):
        raise NotImplementedError("Set Up method not implement in Accelerator, please check! ")

    def broadcast(self):
        raise NotImplementedError("Broadcast method not implement in Accelerator, please check! ")

    def backward_step(self, loss):
        loss.backward()

    def optimizer_step(self, optimizer, model, grad_norm):
        return float(torch.nn.utils.clip_grad_norm_(model.parameters(), grad_norm))
