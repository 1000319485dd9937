"""BaseModel: the orchestration surface train.py / test.py use (reference models/base_model.py:7-171).

Method names, attribute names and the on-disk checkpoint format
(``<checkpoints_dir>/<name>/<epoch>_net_<X>.pth`` holding ``net.state_dict()`` with the reference key
layout, tensors on CPU in torch's standard layout) are kept.
"""
import os
from collections import OrderedDict

import torch

from .. import networks3D

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class BaseModel:
    @staticmethod
    def modify_commandline_options(parser, is_train):
        return parser

    def name(self):
        return "BaseModel"

    def initialize(self, opt):
        self.opt = opt
        self.gpu_ids = opt.gpu_ids
        self.isTrain = opt.isTrain
        # the reference resolves opt.gpu_ids=0 to CPU and then sprinkles .to(cuda) everywhere
        # (base_model.py:22, cycle_gan_model.py:131-135); here the model simply lives on the GPU.
        self.device = device
        self.save_dir = os.path.join(opt.checkpoints_dir, opt.name)
        self.loss_names, self.model_names, self.visual_names, self.image_paths = [], [], [], []

    def set_input(self, input):
        self.input = input

    def forward(self):
        pass

    def optimize_parameters(self):
        pass

    def _nets(self):
        return [(n, getattr(self, "net" + n)) for n in self.model_names if isinstance(n, str)]

    def setup(self, opt, parser=None):
        if self.isTrain:
            self.schedulers = [networks3D.get_scheduler(o, opt) for o in self.optimizers]
        if not self.isTrain or opt.continue_train:
            self.load_networks(opt.which_epoch)
            if self.isTrain:
                self.load_optimizers(opt.which_epoch)      # no-op for checkpoints written by the reference
        self.print_networks(opt.verbose)

    def eval(self):
        for _, net in self._nets():
            net.eval()

    def test(self):
        with torch.no_grad():
            self.forward()

    def get_image_paths(self):
        return self.image_paths

    def update_learning_rate(self):
        for s in self.schedulers:
            s.step()
        print("learning rate = %.7f" % self.optimizers[0].param_groups[0]["lr"])

    def get_current_visuals(self):
        return OrderedDict((n, getattr(self, n)) for n in self.visual_names if isinstance(n, str))

    def get_current_losses(self):
        """Host values of the step's losses: ONE device -> host read for all of them (print time only).  The read
        synchronises anyway, so the tensor-core kernels' time-out flag (a bounded mbarrier wait that expired leaves
        partial output behind) is checked here too and raised instead of training on garbage."""
        names = [n for n in self.loss_names if isinstance(n, str)]
        vals = [getattr(self, "loss_" + n) for n in names]
        dev = [v.detach().float().reshape(()) for v in vals if torch.is_tensor(v)]
        host = torch.stack(dev).cpu().tolist() if dev else []
        out, it = OrderedDict(), iter(host)
        for n, v in zip(names, vals):
            out[n] = next(it) if torch.is_tensor(v) else float(v)
        if dev and dev[0].is_cuda:
            from .. import ops
            I = ops.impl()
            if getattr(I, "name", "") == "cuda":
                code = I.tc_error()
                if code:
                    raise RuntimeError("mra_gan_b200: a tensor-core kernel timed out waiting on a barrier "
                                       "(code %d); the step's results are invalid" % code)
        return out

    # -- checkpoints ---------------------------------------------------------------------------
    def save_networks(self, which_epoch):
        os.makedirs(self.save_dir, exist_ok=True)
        for n, net in self._nets():
            path = os.path.join(self.save_dir, "%s_net_%s.pth" % (which_epoch, n))
            sd = OrderedDict((k, v.detach().to("cpu").contiguous()) for k, v in net.state_dict().items())
            torch.save(sd, path)
        if self.isTrain:
            self.save_optimizers(which_epoch)              # extra file; the reference-format files are unchanged

    # the reference's resume drops the Adam moments (base_model.py:89-148 saves the networks
    # only).  These two write / restore them next to the reference-format files, keyed by the parameters'
    # state_dict names and stored in torch's logical layout, so the extra file is independent of the packed memory
    # layout of this build.  Rank 0 saves under data parallelism (every rank holds the same averaged state).
    def save_optimizers(self, which_epoch):
        if not getattr(self, "optimizers", None):
            return None
        os.makedirs(self.save_dir, exist_ok=True)
        names = {}
        for n, net in self._nets():
            for k, p in net.named_parameters():
                names[p] = "%s.%s" % (n, k)
        out = OrderedDict()
        for oi, opt in enumerate(self.optimizers):
            entry = OrderedDict(param_groups=[{k: v for k, v in g.items() if k != "params"} for g in opt.param_groups],
                                state=OrderedDict())
            for p, st in opt.state.items():
                if p in names and st:
                    entry["state"][names[p]] = OrderedDict(
                        (k, (v.detach().to("cpu").contiguous() if torch.is_tensor(v) else v)) for k, v in st.items())
            out["optimizer_%d" % oi] = entry
        path = os.path.join(self.save_dir, "%s_optim.pth" % which_epoch)
        torch.save(out, path)
        return path

    def load_optimizers(self, which_epoch):
        """Restore the Adam moments and step counters written by save_optimizers(); returns False when the checkpoint
        has none (a checkpoint of the reference), in which case training resumes like the reference does."""
        path = os.path.join(self.save_dir, "%s_optim.pth" % which_epoch)
        if not getattr(self, "optimizers", None) or not os.path.exists(path):
            return False
        blob = torch.load(path, map_location="cpu")
        params = {}
        for n, net in self._nets():
            for k, p in net.named_parameters():
                params["%s.%s" % (n, k)] = p
        for oi, opt in enumerate(self.optimizers):
            entry = blob.get("optimizer_%d" % oi)
            if entry is None:
                continue
            for g, saved in zip(opt.param_groups, entry["param_groups"]):
                for k, v in saved.items():
                    if k != "lr":                      # the schedulers own the learning rate (setup() re-creates them)
                        g[k] = v
            for name, st in entry["state"].items():
                p = params.get(name)
                if p is None:
                    continue
                dst = opt.state[p]
                for k, v in st.items():
                    if torch.is_tensor(v):
                        t = torch.empty_like(p, memory_format=torch.preserve_format) if v.shape == p.shape else \
                            torch.empty(v.shape, dtype=v.dtype, device=p.device)
                        t.copy_(v.to(p.device))        # logical-layout copy into the parameter's (packed) strides
                        dst[k] = t
                    else:
                        dst[k] = v
        return True

    def load_networks(self, which_epoch):
        for n, net in self._nets():
            path = os.path.join(self.save_dir, "%s_net_%s.pth" % (which_epoch, n))
            print("loading the model from %s" % path)
            sd = torch.load(path, map_location="cpu")
            if hasattr(sd, "_metadata"):
                del sd._metadata
            own = net.state_dict()
            for k in list(sd.keys()):
                # pre-0.4 InstanceNorm checkpoints / DataParallel prefixes (base_model.py:114-148, utils.py:24-32)
                if k.startswith("module."):
                    sd[k[7:]] = sd.pop(k)
            for k in list(sd.keys()):
                if k.endswith("num_batches_tracked") and k not in own:
                    sd.pop(k)
            missing = [k for k in own if k not in sd and k.endswith("num_batches_tracked")]
            for k in missing:
                sd[k] = own[k].detach().cpu().clone()
            net.load_state_dict(sd)

    def print_networks(self, verbose):
        print("---------- Networks initialized -------------")
        for n, net in self._nets():
            if verbose:
                print(net)
            print("[Network %s] Total number of parameters : %.3f M" % (n, sum(p.numel() for p in net.parameters()) / 1e6))
        print("-----------------------------------------------")

    def set_requires_grad(self, nets, requires_grad=False):
        for net in nets if isinstance(nets, list) else [nets]:
            if net is not None:
                for p in net.parameters():
                    p.requires_grad = requires_grad
