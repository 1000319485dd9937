"""Fused multi-tensor Adam with torch.optim.Adam semantics (the reference builds
torch.optim.Adam(lr, betas=(beta1, 0.999)) for G and D: models/cycle_gan_model.py:107-110).

It is a regular ``torch.optim.Optimizer`` (param_groups / state_dict / LR schedulers work), but
``step()`` is one kernel sweep (``mra_adam_multi``) over every parameter that also refreshes the
bf16 shadow copies the tensor-core convolutions read.
"""
import torch

from . import ops
from .networks3D import _WeightsEpoch


class FusedAdam(torch.optim.Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0, amsgrad=False):
        if weight_decay != 0 or amsgrad:
            raise NotImplementedError("the reference uses plain Adam (no weight decay, no amsgrad)")
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps))

    # -- device-resident hyper-parameters and step counter -------------------------------------------
    # Per param group: ``state`` = {lr, beta1, beta2, eps, t} as doubles and ``hyper`` = the float block the sweep
    # reads, both in DEVICE memory.  mra_adam_advance (launched right before the sweep, hence part of a captured
    # graph) bumps t and derives the bias corrections on the device, so replaying the step never reads anything
    # the host may already have overwritten for a later step (the pinned staging buffer this replaces was rewritten
    # by the host while earlier replays' copies were still queued).  The host uploads ``state`` only when a rate
    # changes (LR scheduler) or its own step count disagrees with what the device will hold (first step, loaded
    # checkpoint); that copy comes from pageable memory, i.e. it is staged before the call returns and ordered on
    # the current stream like every launch.
    def _dev_state(self, gi, group, device, t_before):
        recs = self.__dict__.setdefault("_dev", {})
        rec = recs.get(gi)
        if rec is None:
            rec = recs[gi] = {"state": torch.zeros(8, dtype=torch.float64, device=device),
                              "hyper": torch.zeros(8, dtype=torch.float32, device=device), "key": None, "t": None}
        key = (float(group["lr"]), float(group["betas"][0]), float(group["betas"][1]), float(group["eps"]))
        if rec["key"] != key or rec["t"] != t_before:
            if torch.cuda.is_current_stream_capturing():
                raise RuntimeError("FusedAdam: the device step counter must be in sync before a CUDA-graph capture "
                                   "(run one eager step or call sync_device_state() first)")
            rec["state"].copy_(torch.tensor(list(key) + [float(t_before), 0.0, 0.0, 0.0], dtype=torch.float64))
            rec["key"] = key
        rec["t"] = t_before + 1
        return rec

    def _group_step(self, group):
        for p in group["params"]:
            st = self.state.get(p)
            if st:
                return int(st["step"])
        return None

    def sync_device_state(self):
        """Upload rates / step counts the device does not have yet (call before capturing ``step()``)."""
        for gi, group in enumerate(self.param_groups):
            t = self._group_step(group)
            if t is None:
                continue
            dev = next(p.device for p in group["params"])
            rec = self._dev_state(gi, group, dev, t)
            rec["t"] = t                                  # nothing was launched: the device still holds t

    def advance_host_state(self):
        """Bookkeeping of one optimiser step WITHOUT launching the sweep (a captured graph of ``step()`` is about to
        be replayed): bump the host's step counters, and re-upload the device state if a rate changed meanwhile."""
        for gi, group in enumerate(self.param_groups):
            step_no = None
            for p in group["params"]:
                st = self.state.get(p)
                if st:
                    st["step"] += 1
                    step_no = st["step"]
                    self._mark(p)                      # the replayed Adam kernel rewrites weight and shadow
            if step_no is not None and gi in self.__dict__.get("_dev", {}):
                self._dev_state(gi, group, self._dev[gi]["state"].device, step_no - 1)
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
                    rec = self._dev_state(gi, group, ps[0].device, step_no - 1)
                    I.adam_advance(rec["state"], rec["hyper"])       # t += 1, bias corrections of step t (on device)
                    I.adam_step_dev(ps, gs, ms, vs, shs, rec["hyper"])
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
