"""FlowMatchEulerDiscreteScheduler with the surface the reference pipelines use
(``RepText/pipeline_flux_controlnet.py:154, :954-957, :995, :1109``): ``config``, ``set_timesteps(sigmas=,
device=, mu=)``, ``timesteps``, ``sigmas``, ``order``, ``step(model_output, t, sample, return_dict=False)``.

Semantics follow diffusers 0.36.0 ``scheduling_flow_match_euler_discrete.py`` with FLUX.1-dev's
``scheduler_config.json`` (dynamic exponential shifting).  ``step`` runs the hand-written Euler kernel
(``rt_euler_step``).  diffusers keeps ``sigmas`` on the DEVICE and indexes it per step (plus a device -> host sync to
find the step index); here a host copy of the table feeds the kernel's scalar arguments, so a step never syncs, and the
kernel reproduces the rounding of the device-tensor form (``dt`` rounded to the model dtype; csrc/elementwise.cu).
"""
from __future__ import annotations

import math
from typing import List, Optional, Tuple, Union

import numpy as np
import torch

from . import ops
from .config import SCHEDULER
from .models import FrozenConfig


class FlowMatchEulerDiscreteScheduler:
    order = 1

    def __init__(self, **config):
        cfg = dict(SCHEDULER)
        cfg.update(config)
        self.config = FrozenConfig(**cfg)
        n = cfg["num_train_timesteps"]
        ts = np.linspace(1, n, n, dtype=np.float32)[::-1].copy()
        sig = ts / n
        if not cfg["use_dynamic_shifting"]:
            sig = cfg["shift"] * sig / (1 + (cfg["shift"] - 1) * sig)
        self.sigmas = torch.from_numpy(sig)
        self.timesteps = self.sigmas * n
        self._host_timesteps: List[float] = self.timesteps.tolist()
        self._host_sigmas: List[float] = self.sigmas.tolist() + [0.0]
        self._step_index: Optional[int] = None
        self._begin_index: Optional[int] = None
        self.num_inference_steps: Optional[int] = None

    @classmethod
    def from_pretrained(cls, pretrained_model_name_or_path, subfolder: Optional[str] = None, **unused):
        """``<dir>[/subfolder]/scheduler_config.json`` (``black-forest-labs/FLUX.1-dev``'s ``scheduler/``), local only."""
        from . import checkpoint as ck
        d = ck.resolve_dir(pretrained_model_name_or_path, subfolder)
        raw, stored_cls = ck.read_config(d, ("scheduler_config.json", "config.json"))
        if stored_cls not in (None, "FlowMatchEulerDiscreteScheduler"):
            raise ValueError(f"{d!r} holds a {stored_cls}; the RepText pipelines step with FlowMatchEulerDiscreteScheduler")
        # diffusers >= 0.31 writes every constructor default when it re-saves a scheduler: keys that sit at their diffusers
        # default (an option that is OFF, or the formula this class implements) are accepted and ignored
        at_default = dict(time_shift_type="exponential", stochastic_sampling=False, invert_sigmas=False,
                          use_karras_sigmas=False, use_exponential_sigmas=False, use_beta_sigmas=False, shift_terminal=None)
        unknown = [k for k in raw if k not in SCHEDULER and raw[k] not in (None, False)
                   and not (k in at_default and raw[k] == at_default[k])]
        if unknown:
            raise ValueError(f"scheduler options {unknown} are not implemented (FLUX.1-dev's scheduler_config.json uses none)")
        return cls(**{k: v for k, v in raw.items() if k in SCHEDULER})

    def save_pretrained(self, save_directory) -> None:
        from . import checkpoint as ck
        ck.write_config(save_directory, {k: getattr(self.config, k) for k in SCHEDULER},
                        "FlowMatchEulerDiscreteScheduler", name="scheduler_config.json")

    @property
    def step_index(self):
        return self._step_index

    @property
    def begin_index(self):
        return self._begin_index

    def set_begin_index(self, begin_index: int = 0):
        self._begin_index = begin_index

    def time_shift(self, mu: float, sigma: float, t: np.ndarray) -> np.ndarray:
        return math.exp(mu) / (math.exp(mu) + (1 / t - 1) ** sigma)

    def set_timesteps(self, num_inference_steps: Optional[int] = None, device: Union[str, torch.device, None] = None,
                      sigmas: Optional[List[float]] = None, mu: Optional[float] = None,
                      timesteps: Optional[List[float]] = None):
        c = self.config
        if c.use_dynamic_shifting and mu is None:
            raise ValueError("`mu` must be passed when `use_dynamic_shifting` is set to be `True`")
        if timesteps is not None:
            raise ValueError("custom `timesteps` are not used by the RepText pipelines; pass `sigmas`")
        if sigmas is None:
            if num_inference_steps is None:
                raise ValueError("pass `num_inference_steps` or `sigmas`")
            t = np.linspace(float(c.num_train_timesteps), 1.0, num_inference_steps)
            sigmas = t / c.num_train_timesteps
        sigmas = np.array(sigmas).astype(np.float32)
        self.num_inference_steps = len(sigmas)
        if c.use_dynamic_shifting:
            sigmas = self.time_shift(mu, 1.0, sigmas)
        else:
            sigmas = c.shift * sigmas / (1 + (c.shift - 1) * sigmas)
        sig = torch.from_numpy(np.asarray(sigmas)).to(dtype=torch.float32)
        ts = sig * c.num_train_timesteps
        self._host_timesteps = ts.tolist()
        self._host_sigmas = torch.cat([sig, torch.zeros(1)]).tolist()
        self.timesteps = ts.to(device=device)
        self.sigmas = torch.cat([sig, torch.zeros(1)])      # host copy, fp32 (see the module docstring)
        self._step_index = None
        self._begin_index = None

    def index_for_timestep(self, timestep) -> int:
        t = float(timestep)  # one device->host read on the first step of a run, like diffusers
        hits = [i for i, v in enumerate(self._host_timesteps) if v == t]
        if not hits:
            raise ValueError(f"timestep {t} is not in the schedule")
        return hits[1] if len(hits) > 1 else hits[0]

    def _init_step_index(self, timestep):
        self._step_index = self.index_for_timestep(timestep) if self._begin_index is None else self._begin_index

    def _sigma_pair(self, timestep) -> Tuple[float, float]:
        if self._step_index is None:
            self._init_step_index(timestep)
        i = self._step_index
        return self._host_sigmas[i], self._host_sigmas[i + 1]

    def step(self, model_output: torch.Tensor, timestep, sample: torch.Tensor, return_dict: bool = True, **kwargs):
        """``prev = float(sample) + (sigma_next - sigma) * model_output``, cast to ``model_output.dtype``."""
        s, s_next = self._sigma_pair(timestep)
        if sample.dtype != model_output.dtype:
            sample = sample.to(model_output.dtype)
        prev = ops.euler_step(model_output.contiguous(), sample.contiguous(), s, s_next)
        self._step_index += 1
        if not return_dict:
            return (prev,)
        return FrozenConfig(prev_sample=prev)

    def step_cfg(self, model_output_2: torch.Tensor, timestep, sample: torch.Tensor, true_guidance_scale: float,
                 zero_pred: bool) -> torch.Tensor:
        """True-CFG combine (pipeline_flux_controlnet_inpaint.py:1264-1270) fused with the Euler step."""
        s, s_next = self._sigma_pair(timestep)
        prev = ops.cfg_euler_step(model_output_2.contiguous(), sample.contiguous(), true_guidance_scale, zero_pred,
                                  s, s_next)
        self._step_index += 1
        return prev
