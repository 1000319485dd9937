"""Fused multi-tensor Adam with torch.optim.Adam semantics (the reference builds
torch.optim.Adam(lr, betas=(beta1, 0.999)) for G and D: models/cycle_gan_model.py:107-110).

It is a regular ``torch.optim.Optimizer`` (param_groups / state_dict / LR schedulers work), but
``step()`` is one kernel sweep (``mra_adam_multi``) over every parameter that also refreshes the
bf16 shadow copies the tensor-core convolutions read.
"""
import math

import torch

from . import ops
from .networks3D import _WeightsEpoch


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("the reference uses plain Adam (no weight decay, no amsgrad)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    # -- device-resident hyper-parameters (CUDA-graph replay) -----------------------------------------
    def _hyper_buffers(self, gi, device):
        bufs = self.__dict__.setdefault("_hyper", {})
        if gi not in bufs:
            bufs[gi] = (torch.zeros(8, dtype=torch.float32).pin_memory(), torch.zeros(8, dtype=torch.float32, device=device))
        return bufs[gi]

    @staticmethod
    def _fill_hyper(host, group, step_no):
        b1, b2 = group["betas"]
        lr = float(group["lr"])
        host[0], host[1], host[2], host[3] = lr, b1, b2, group["eps"]
        host[4] = lr / (1.0 - b1 ** step_no)
        host[5] = math.sqrt(1.0 - b2 ** step_no)

    def advance_host_state(self):
        """Bookkeeping of one optimiser step WITHOUT launching anything: bump the step counters and refresh the
        pinned hyper-parameter buffers (a captured graph of ``step()`` copies them to the device when replayed)."""
        for gi, group in enumerate(self.param_groups):
            step_no = None
            for p in group["params"]:
                st = self.state.get(p)
                if st:
                    st["step"] += 1
                    step_no = st["step"]
                    self._mark(p)                      # the replayed Adam kernel rewrites weight and shadow
            if step_no is not None and gi in self.__dict__.get("_hyper", {}):
                self._fill_hyper(self._hyper[gi][0], group, step_no)
        _WeightsEpoch.value += 1

    @torch.no_grad()
    def step(self, closure=None):
        loss = closure() if closure is not None else None
        I = ops.impl()
        use_dev = hasattr(I, "adam_step_dev")
        for gi, group in enumerate(self.param_groups):
            ps, gs, ms, vs, shs = [], [], [], [], []
            step_no = None
            for p in group["params"]:
                if p.grad is None:
                    continue
                st = self.state[p]
                if not st:
                    st["step"] = 0
                    st["exp_avg"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                    st["exp_avg_sq"] = torch.zeros_like(p, memory_format=torch.preserve_format)
                st["step"] += 1
                if step_no is None:
                    step_no = st["step"]
                elif step_no != st["step"]:
                    # parameters that joined late get their own sweep below
                    I.adam_step([p], [self._grad(p)], [st["exp_avg"]], [st["exp_avg_sq"]], [self._shadow(p)],
                                group["lr"], group["betas"][0], group["betas"][1], group["eps"], st["step"])
                    self._mark(p)
                    continue
                ps.append(p); gs.append(self._grad(p)); ms.append(st["exp_avg"]); vs.append(st["exp_avg_sq"])
                shs.append(self._shadow(p))
            if ps:
                if use_dev:
                    host, dev = self._hyper_buffers(gi, ps[0].device)
                    self._fill_hyper(host, group, step_no)
                    dev.copy_(host, non_blocking=True)        # a memcpy node when the step is being captured
                    I.adam_step_dev(ps, gs, ms, vs, shs, dev)
                else:
                    I.adam_step(ps, gs, ms, vs, shs, group["lr"], group["betas"][0], group["betas"][1], group["eps"],
                                step_no)
                for p in ps:
                    self._mark(p)
        _WeightsEpoch.value += 1
        return loss

    @staticmethod
    def _grad(p):
        g = p.grad
        if g.stride() != p.stride():
            g = torch.empty_like(p, memory_format=torch.preserve_format).copy_(g)
        return g

    @staticmethod
    def _shadow(p):
        return getattr(p, "_mra_shadow", None)

    @staticmethod
    def _mark(p):
        """The weight behind ``p`` changed (and its bf16 shadow with it): bump the parameter's epoch and stamp the
        shadow with the tag _ConvNd._tag() computes from now on."""
        p._mra_epoch = getattr(p, "_mra_epoch", 0) + 1
        if getattr(p, "_mra_shadow", None) is not None:
            p._mra_shadow_tag = (p._version, p._mra_epoch, p.data_ptr())
