"""scvi.module.base.{BaseModuleClass, LossOutput, auto_move_data} of scvi-tools 0.20.0,
restated (TEST INFRASTRUCTURE ONLY).  Used by the reference at
src/spVIPES/module/spVIPESmodule.py:9,18,425,720,895.
"""
from dataclasses import dataclass, field
from functools import wraps
from typing import Any, Optional

import torch
from torch import nn


def _move(obj, device):
    if isinstance(obj, torch.Tensor):
        return obj.to(device)
    if isinstance(obj, dict):
        return {k: _move(v, device) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_move(v, device) for v in obj)
    return obj


def auto_move_data(fn):
    @wraps(fn)
    def wrapper(self, *args, **kwargs):
        if not isinstance(self, nn.Module):
            return fn(self, *args, **kwargs)
        device = list({p.device for p in self.parameters()})
        if len(device) > 1:
            raise RuntimeError("Module tensors on multiple devices.")
        device = device[0]
        return fn(self, *_move(args, device), **_move(kwargs, device))

    return wrapper


@dataclass
class LossOutput:
    loss: Any
    reconstruction_loss: Optional[Any] = None
    kl_local: Optional[Any] = None
    kl_global: Optional[Any] = None
    extra_metrics: Optional[dict] = field(default_factory=dict)
    n_obs_minibatch: Optional[int] = None

    def __post_init__(self):
        if self.n_obs_minibatch is None and self.reconstruction_loss is not None:
            rec = self.reconstruction_loss
            first = next(iter(rec.values())) if isinstance(rec, dict) else rec
            self.n_obs_minibatch = first.shape[0]

    @staticmethod
    def _sum(d):
        if d is None:
            return 0.0
        if isinstance(d, dict):
            return sum(torch.sum(v) for v in d.values())
        return torch.sum(d)

    @property
    def reconstruction_loss_sum(self):
        return self._sum(self.reconstruction_loss)

    @property
    def kl_local_sum(self):
        return self._sum(self.kl_local)


class BaseModuleClass(nn.Module):
    @property
    def device(self):
        device = list({p.device for p in self.parameters()})
        if len(device) > 1:
            raise RuntimeError("Module tensors on multiple devices.")
        return device[0]

    def forward(
        self,
        tensors,
        get_inference_input_kwargs=None,
        get_generative_input_kwargs=None,
        inference_kwargs=None,
        generative_kwargs=None,
        loss_kwargs=None,
        compute_loss=True,
    ):
        inference_kwargs = inference_kwargs or {}
        generative_kwargs = generative_kwargs or {}
        loss_kwargs = loss_kwargs or {}
        get_inference_input_kwargs = get_inference_input_kwargs or {}
        get_generative_input_kwargs = get_generative_input_kwargs or {}
        inference_inputs = self._get_inference_input(tensors, **get_inference_input_kwargs)
        inference_outputs = self.inference(**inference_inputs, **inference_kwargs)
        generative_inputs = self._get_generative_input(tensors, inference_outputs, **get_generative_input_kwargs)
        generative_outputs = self.generative(**generative_inputs, **generative_kwargs)
        if compute_loss:
            losses = self.loss(tensors, inference_outputs, generative_outputs, **loss_kwargs)
            return inference_outputs, generative_outputs, losses
        return inference_outputs, generative_outputs
