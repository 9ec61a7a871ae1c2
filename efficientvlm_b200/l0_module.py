"""Hard-concrete L0 gate modules — drop-in for the reference's `efficient_models/xvlm_l0_module.py` (XVLML0Module),
`generation_l0_module.py` (VQAL0Module, adds decoder_* gates) and `nlvr_l0_module.py` (NLVRL0Module, 2x cross layers):
same constructor arguments, parameter names (`*_loga`, `lambda_1`, `lambda_2`), `forward(training)` dict (same key order),
`lagrangian_regularization`, `constrain_parameters`, `calculate_model_size`.

Sampling, the deterministic top-k mask and the expected-size reduction run in the sm_100a kernels (efficientvlm_b200.ops /
kernels); the per-layer Python loop with `.item()` of the reference's eval path (xvlm_l0_module.py:253-271,330-339) becomes
one launch per gate type.
"""
import math
import os

import numpy as np
import torch
from torch.nn import Module
from torch.nn.parameter import Parameter

from . import kernels as K
from . import ops
from .eff_bert import BertConfig
from .xvlm import read_json

limit_a, limit_b, epsilon = -.1, 1.1, 1e-6


class _L0ModuleBase(Module):
    variant = "xvlm"

    def __init__(self, config, droprate_init=0.5, temperature=2. / 3., lagrangian_warmup=0, start_sparsity=0.0, target_sparsity=0.0,
                 pruning_type="structured_heads+structured_mlp", magical_number=0.8):
        super().__init__()
        cfg_path = os.path.join(config["text_encoder"], "config.json") if config.get("text_encoder") else None
        text_config = BertConfig.from_json_file(cfg_path) if cfg_path and os.path.exists(cfg_path) else BertConfig()
        text_config.num_hidden_layers = config["text_num_hidden_layers"] if "text_num_hidden_layers" in config else 12
        assert text_config.num_hidden_layers in [6, 12], "param initialization not implemented"
        text_config.fusion_layer = text_config.num_hidden_layers // 2
        if self.variant == "nlvr":                                             # nlvr_l0_module.py:36-39
            num_text_layers = text_config.fusion_layer
            num_cross_layers = text_config.num_hidden_layers - text_config.fusion_layer
            text_config.num_hidden_layers = num_text_layers + 2 * num_cross_layers
        vision_config = read_json(config["vision_config"])
        assert config["patch_size"] == vision_config["patch_size"]
        self.all_types = ["vision_intermediate_z", "vision_head_z", "text_intermediate_z", "text_head_z", "cross_intermediate_z",
                          "cross_head_z"]
        if self.variant == "vqa":
            self.all_types += ["decoder_head_z", "decoder_intermediate_z"]
        self.pruning_type = pruning_type
        self.hidden_size = text_config.hidden_size
        self.intermediate_size = text_config.intermediate_size
        self.num_attention_heads = text_config.num_attention_heads
        self.dim_per_head = self.hidden_size // self.num_attention_heads
        self.vision_num_hidden_layers = vision_config["num_hidden_layers"]
        self.text_num_hidden_layers = text_config.fusion_layer
        self.cross_num_hidden_layers = text_config.num_hidden_layers - text_config.fusion_layer
        if self.variant == "vqa":
            self.decoder_num_hidden_layers = self.cross_num_hidden_layers
        self.mlp_num_per_layer = 1
        self.params_per_head_layer = self.hidden_size * self.hidden_size * 4 + self.hidden_size * 4
        self.params_per_head = self.params_per_head_layer // self.num_attention_heads
        self.params_per_mlp_layer = self.hidden_size * self.intermediate_size * 2 + self.hidden_size + self.hidden_size * 4
        self.params_per_intermediate_dim = self.params_per_mlp_layer // self.intermediate_size   # integer floor (quirk Q7)
        self.full_model_size = (self.params_per_head_layer + self.params_per_mlp_layer) * self.vision_num_hidden_layers + \
                               (self.params_per_head_layer + self.params_per_mlp_layer) * self.text_num_hidden_layers + \
                               (self.params_per_head_layer * 2 + self.params_per_mlp_layer) * self.cross_num_hidden_layers
        if self.variant == "vqa":
            self.full_model_size += (self.params_per_head_layer * 2 + self.params_per_mlp_layer) * self.decoder_num_hidden_layers
        self.prunable_model_size = 0
        self.temperature = temperature
        self.droprate_init = droprate_init if droprate_init != 0. else 0.5
        self.types = []
        self.z_logas = {}
        self.parameters_per_dim = {}
        self.sizes = {}
        self.shapes = {}
        self.hidden_loga = None
        self.hidden_type = None
        types = self.pruning_type.split("+")
        for type in types:
            if type != "layer":
                self.initialize_one_module(type)
        if "layer" in types:
            self.initialize_one_module("layer")
        self.magical_number = magical_number
        self.lambda_1 = torch.nn.Parameter(torch.tensor(0.0))
        self.lambda_2 = torch.nn.Parameter(torch.tensor(0.0))
        self.lagrangian_warmup = lagrangian_warmup
        self.start_sparsity = start_sparsity
        self.target_sparsity = target_sparsity

    def set_lagrangian_warmup_steps(self, lagrangian_warmup):
        self.lagrangian_warmup = lagrangian_warmup

    def initialize_one_module(self, module_name):
        if module_name == "structured_heads":
            self.initialize_structured_head()
        elif module_name == "structured_mlp":
            self.initialize_structured_mlp()

    def add_one_module(self, z_loga, type, parameter_per_dim, size, shape):
        self.types.append(type)
        self.z_logas[type] = z_loga
        self.parameters_per_dim[type] = parameter_per_dim
        self.sizes[type] = size
        self.shapes[type] = shape

    def initialize_parameters(self, size, num_layer=None):
        if num_layer is not None:
            return Parameter(torch.Tensor(num_layer, size))
        return Parameter(torch.Tensor(size))

    def _modalities(self):
        mods = [("vision", self.vision_num_hidden_layers, 1), ("text", self.text_num_hidden_layers, 1),
                ("cross", self.cross_num_hidden_layers, 2)]
        if self.variant == "vqa":
            mods.append(("decoder", self.decoder_num_hidden_layers, 2))
        return mods

    def initialize_structured_head(self, add_prunable_model_size=True):
        for name, layers, mult in self._modalities():
            loga = self.initialize_parameters(self.num_attention_heads, layers * mult)
            setattr(self, name + "_head_loga", loga)
        for name, layers, mult in self._modalities():
            self.reset_loga(getattr(self, name + "_head_loga"), mean=10)
        for name, layers, mult in self._modalities():
            self.add_one_module(getattr(self, name + "_head_loga"), type=name + "_head", parameter_per_dim=self.params_per_head,
                                size=self.num_attention_heads, shape=[layers * mult, 1, self.num_attention_heads, 1, 1])
        if add_prunable_model_size:
            for name, layers, mult in self._modalities():
                self.prunable_model_size += self.params_per_head * layers * mult * self.num_attention_heads

    def initialize_structured_mlp(self):
        for name, layers, _ in self._modalities():
            setattr(self, name + "_int_loga", self.initialize_parameters(self.intermediate_size, layers))
        for name, layers, _ in self._modalities():
            self.add_one_module(getattr(self, name + "_int_loga"), type=name + "_intermediate",
                                parameter_per_dim=self.params_per_intermediate_dim, size=self.intermediate_size,
                                shape=[layers, 1, 1, self.intermediate_size])
            self.prunable_model_size += self.params_per_mlp_layer * layers
        for name, layers, _ in self._modalities():
            self.reset_loga(getattr(self, name + "_int_loga"))

    def reset_loga(self, tensor, mean=None):
        if mean is None:
            mean = math.log(1 - self.droprate_init) - math.log(self.droprate_init)
        tensor.data.normal_(mean, 1e-2)

    def constrain_parameters(self):
        for key in self.z_logas:
            t = self.z_logas[key].data
            if t.is_cuda:
                K.clamp_(t, math.log(1e-2), math.log(1e2))
            else:
                t.clamp_(min=math.log(1e-2), max=math.log(1e2))
        ops.invalidate_weight_cache(list(self.z_logas.values()))     # the in-place kernel does not bump autograd's version counter

    # -------------------------------------------------------------------------------------------- Lagrangian
    def get_num_parameters_and_constraint(self):
        """sum_type sum(1 - cdf_qz(0, loga)) * params_per_dim  (heads first, then intermediates: the reference's order)."""
        order = [t for t in self.types if t.endswith("_head")] + [t for t in self.types if t.endswith("_intermediate")]
        return ops.l0_expected_size([self.z_logas[t] for t in order], [float(self.parameters_per_dim[t]) for t in order],
                                    self.temperature)

    def get_target_sparsity(self, pruned_steps):
        if torch.is_tensor(pruned_steps):   # extension: a device-resident step counter (a captured step graph advances it itself)
            frac = torch.clamp(pruned_steps.to(torch.float32) / self.lagrangian_warmup, max=1.0)
            return (self.target_sparsity - self.start_sparsity) * frac + self.start_sparsity
        return (self.target_sparsity - self.start_sparsity) * min(1, pruned_steps / self.lagrangian_warmup) + self.start_sparsity

    def lagrangian_regularization(self, pruned_steps):
        target_sparsity = self.target_sparsity
        expected_size = self.get_num_parameters_and_constraint()
        expected_sparsity = 1 - expected_size / self.prunable_model_size
        if self.lagrangian_warmup > 0:
            target_sparsity = self.get_target_sparsity(pruned_steps)
        lagrangian_loss = (self.lambda_1 * (expected_sparsity - target_sparsity)
                           + self.lambda_2 * (expected_sparsity - target_sparsity) ** 2)
        return lagrangian_loss, expected_sparsity, target_sparsity

    # -------------------------------------------------------------------------------------------- gates
    def get_eps(self, size):
        """Uniform noise for the concrete distribution, drawn on the CPU generator like the reference (quirk Q5)."""
        return torch.FloatTensor(size).uniform_(epsilon, 1 - epsilon)

    def _sample_z(self, loga):
        eps = self.get_eps(torch.FloatTensor(*loga.shape))
        if eps.device != loga.device:       # a get_eps override may hand out device-resident noise (static graph inputs)
            eps = eps.to(loga.device, non_blocking=True)
        return ops.l0_sample(loga, eps, self.temperature)

    def _deterministic_all_layers(self, loga):
        """Deterministic masks are a function of the log-alphas alone: they are computed once per log-alpha VERSION and re-used by
        every evaluation batch (they were 13 % of the VQA inference kernel time when recomputed per batch).  The key follows the
        same rules as the bf16 weight shadows (ops.weight_bf16): autograd version, storage, and the per-parameter epoch that every
        optimizer step / `constrain_parameters` bumps."""
        key = (loga._version, loga.data_ptr(), ops._pepoch.get(id(loga), 0), ops._epoch[0], self.temperature, self.magical_number)
        cache = self.__dict__.setdefault("_det_cache", {})
        ent = cache.get(id(loga))
        if ent is not None and ent[0] == key:
            return ent[1]
        mask, _ = K.l0_deterministic(loga.detach().contiguous(), self.temperature, self.magical_number)
        cache[id(loga)] = (key, mask)
        ops.register_static_gate(mask, key)
        return mask

    def get_z_from_zs(self, zs):
        numpified_zs = {}
        for type in self.all_types:
            name = type[:-2]
            z = zs.get(type, np.ones(self.shapes[name]))
            if torch.is_tensor(z):
                new_z = z.squeeze().detach().cpu().numpy() > 0
            numpified_zs[name] = new_z
        return numpified_zs

    def calculate_model_size(self, zs):
        nz = self.get_z_from_zs(zs)
        results = {}
        head_nums = 0
        intermediate_nums = 0
        for name, layers, mult in self._modalities():
            inter = nz[name + "_intermediate"].reshape(layers, self.intermediate_size).sum(-1).tolist()
            heads = nz[name + "_head"].reshape(layers * mult, self.num_attention_heads).sum(-1).tolist()
            results[name + "_intermediate_dims"] = inter
            results[name + "_head_nums"] = heads
            head_nums += sum(heads)
            intermediate_nums += sum(inter)
        remaining_model_size = head_nums * self.params_per_head + intermediate_nums * 2 * self.hidden_size
        pruned_model_size = self.prunable_model_size - remaining_model_size
        results["pruned_params"] = pruned_model_size
        results["remaining_params"] = remaining_model_size
        results["pruned_model_sparsity"] = pruned_model_size / self.prunable_model_size
        return results

    def forward(self, training=True):
        zs = {f"{type}_z": [] for type in self.types}
        if training:
            for type in self.types:
                loga = self.z_logas[type]
                zs[f"{type}_z"] = self._sample_z(loga).reshape(self.shapes[type])
        else:
            for type in self.types:
                zs[f"{type}_z"] = self._deterministic_all_layers(self.z_logas[type]).reshape(self.shapes[type])
        return zs


class XVLML0Module(_L0ModuleBase):
    variant = "xvlm"


class VQAL0Module(_L0ModuleBase):
    variant = "vqa"


class NLVRL0Module(_L0ModuleBase):
    variant = "nlvr"
