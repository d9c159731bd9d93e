"""Optimiser for the drop-in module: torch.optim.Optimizer interface, ONE launch of the fused Adam kernel over the engine's flat
parameter / gradient buffers per step (torch.optim.Adam semantics: L2 weight decay folded into the gradient, bias correction).

scvi's TrainingPlan builds `torch.optim.Adam(params, lr=1e-3, eps=0.01, weight_decay=1e-6)` (reference
model/base/training_mixin.py:93-111); its multi-tensor implementation costs ~3.8 ms of host time per step on 66 parameter
tensors (tools/profile_plugin.py), ten times the GPU time of a whole C2 step.  scvi 0.20's TrainingPlan accepts
`optimizer="Custom", optimizer_creator=...`: pass `lambda params: FlatAdam(module)`.

The kernel also refreshes the 16-bit tensor-core operand copies of the large weights (spv_adam staging segments), so the next
forward does not start with conversion launches."""
from __future__ import annotations

import torch


class FlatAdam(torch.optim.Optimizer):
    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=0.01, weight_decay=1e-6):
        self.module = module
        super().__init__(list(module.parameters()), dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay))
        eng = module.engine
        if eng.bf16 and max(eng.d.genes) * max(eng.d.KMIX, 2 * eng.d.n_hidden) < (1 << 24):
            eng.stage_in_adam = True  # this optimiser owns the parameter update: it keeps the 16-bit operand copies current

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        g = self.param_groups[0]
        eng = self.module.engine
        eng.adam_step(lr=g["lr"], betas=g["betas"], eps=g["eps"], weight_decay=g["weight_decay"])
        if eng.stage_in_adam:
            eng._staged_version = eng.params.flat._version
        return loss

    def zero_grad(self, set_to_none: bool = True):
        # the backward kernels overwrite the flat gradient buffer; nothing to clear
        if set_to_none:
            return
        self.module.engine.grads.zero_()
