"""CycleGANModel: the training step (reference models/cycle_gan_model.py:38-240) on the fused kernels.

Per ``optimize_parameters()`` and sample: 6 generator forwards, 6 generator backwards, 6 discriminator
forwards and 6 backwards (2 of them dgrad-only), two fused-Adam sweeps.  Loss names, visual names,
the ImagePool policy and the host RNG call order follow the reference.
"""
import gc
import itertools
import random

import torch

from .. import networks3D
from ..optim import FusedAdam
from .base_model import BaseModel


class ImagePool:
    """History buffer of generated volumes (cycle_gan_model.py:8-35): fill to ``pool_size``; afterwards
    each query returns, with probability 1/2, a random stored volume (which the new one replaces)."""

    def __init__(self, pool_size):
        self.pool_size = pool_size
        if pool_size > 0:
            self.num_imgs, self.images = 0, []

    def query(self, images):
        if self.pool_size == 0:
            return images
        picked = []
        for img in images:
            img = img.detach().unsqueeze(0).clone()      # own storage: the generator output may be a reused (graph) buffer
            if self.num_imgs < self.pool_size:
                self.num_imgs += 1
                self.images.append(img)
                picked.append(img)
            elif random.uniform(0, 1) > 0.5:
                idx = random.randint(0, self.pool_size - 1)
                old = self.images[idx].clone()
                self.images[idx] = img
                picked.append(old)
            else:
                picked.append(img)
        return torch.cat(picked, 0)


class CycleGANModel(BaseModel):
    def name(self):
        return "CycleGANModel"

    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        parser.set_defaults(no_dropout=True)
        if is_train:
            parser.add_argument("--lambda_A", type=float, default=10.0, help="weight for cycle loss (A -> B -> A)")
            parser.add_argument("--lambda_B", type=float, default=10.0, help="weight for cycle loss (B -> A -> B)")
            parser.add_argument("--lambda_identity", type=float, default=0.5,
                                help="identity-mapping weight relative to the cycle weight (0 disables it)")
            parser.add_argument("--lambda_co_A", type=float, default=2, help="weight for correlation coefficient loss (A -> B)")
            parser.add_argument("--lambda_co_B", type=float, default=2, help="weight for correlation coefficient loss (B -> A)")
        return parser

    def initialize(self, opt):
        BaseModel.initialize(self, opt)
        self.loss_names = ["D_A", "G_A", "cycle_A", "idt_A", "D_B", "G_B", "cycle_B", "idt_B"]
        vis_A, vis_B = ["real_A", "fake_B", "rec_A"], ["real_B", "fake_A", "rec_B"]
        if self.isTrain and opt.lambda_identity > 0.0:
            vis_A.append("idt_A")
            vis_B.append("idt_B")
        self.visual_names = vis_A + vis_B
        self.model_names = ["G_A", "G_B", "D_A", "D_B"] if self.isTrain else ["G_A", "G_B"]

        mk_G = lambda i, o: networks3D.define_G(i, o, opt.ngf, opt.netG, opt.norm, not opt.no_dropout,
                                                opt.init_type, opt.init_gain, self.gpu_ids)
        self.netG_A = mk_G(opt.input_nc, opt.output_nc)
        self.netG_B = mk_G(opt.output_nc, opt.input_nc)
        if self.isTrain:
            mk_D = lambda i: networks3D.define_D(i, opt.ndf, opt.netD, opt.n_layers_D, opt.norm, opt.no_lsgan,
                                                 opt.init_type, opt.init_gain, self.gpu_ids)
            self.netD_A = mk_D(opt.output_nc)
            self.netD_B = mk_D(opt.input_nc)
            self.fake_A_pool, self.fake_B_pool = ImagePool(opt.pool_size), ImagePool(opt.pool_size)
            self.criterionGAN = networks3D.GANLoss(use_lsgan=not opt.no_lsgan).to(self.device)
            self.criterionCycle = networks3D.L1Loss()
            self.criterionIdt = networks3D.L1Loss()
            self.optimizer_G = FusedAdam(itertools.chain(self.netG_A.parameters(), self.netG_B.parameters()),
                                         lr=opt.lr, betas=(opt.beta1, 0.999))
            self.optimizer_D = FusedAdam(itertools.chain(self.netD_A.parameters(), self.netD_B.parameters()),
                                         lr=opt.lr, betas=(opt.beta1, 0.999))
            self.optimizers = [self.optimizer_G, self.optimizer_D]
            for net in (self.netG_A, self.netG_B, self.netD_A, self.netD_B):
                for m in net.conv_modules():
                    m.make_shadow(net.compute_dtype)
                networks3D.set_fused_wgrad(net, True)       # parallel.attach() keeps it unless MRA_DP_FUSED_WGRAD=0
        self.grad_sync = None        # set by parallel.attach()
        self._graphs = None          # CUDA-graph replay of the step (enable_cuda_graphs)

    def set_input(self, input):
        AtoB = self.opt.which_direction == "AtoB"
        self.real_A = input[0 if AtoB else 1].to(self.device, non_blocking=True)
        self.real_B = input[1 if AtoB else 0].to(self.device, non_blocking=True)

    def forward(self):
        self.fake_B = self.netG_A(self.real_A)
        self.rec_A = self.netG_B(self.fake_B)
        self.fake_A = self.netG_B(self.real_B)
        self.rec_B = self.netG_A(self.fake_A)

    def backward_D_basic(self, netD, real, fake):
        loss_D_real = self.criterionGAN(netD(real), True)
        loss_D_fake = self.criterionGAN(netD(fake.detach()), False)
        loss_D = (loss_D_real + loss_D_fake) * 0.5
        loss_D.backward()
        return loss_D

    def backward_D_A(self):
        self.loss_D_A = self.backward_D_basic(self.netD_A, self.real_B, self.fake_B_pool.query(self.fake_B))

    def backward_D_B(self):
        self.loss_D_B = self.backward_D_basic(self.netD_B, self.real_A, self.fake_A_pool.query(self.fake_A))

    def backward_G(self):
        o = self.opt
        lam_idt, lam_A, lam_B = o.lambda_identity, o.lambda_A, o.lambda_B
        if lam_idt > 0:
            self.idt_A = self.netG_A(self.real_B)
            self.loss_idt_A = self.criterionIdt(self.idt_A, self.real_B) * lam_B * lam_idt
            self.idt_B = self.netG_B(self.real_A)
            self.loss_idt_B = self.criterionIdt(self.idt_B, self.real_A) * lam_A * lam_idt
        else:
            self.loss_idt_A = self.loss_idt_B = 0
        self.loss_G_A = self.criterionGAN(self.netD_A(self.fake_B), True)
        self.loss_G_B = self.criterionGAN(self.netD_B(self.fake_A), True)
        self.loss_cycle_A = self.criterionCycle(self.rec_A, self.real_A) * lam_A
        self.loss_cycle_B = self.criterionCycle(self.rec_B, self.real_B) * lam_B
        # evaluated every step by the reference but never added to loss_G (cycle_gan_model.py:217-223)
        self.loss_cor_coe_GA = networks3D.Cor_CoeLoss(self.fake_B, self.real_A) * o.lambda_co_A
        self.loss_cor_coe_GB = networks3D.Cor_CoeLoss(self.fake_A, self.real_B) * o.lambda_co_B
        self.loss_G = (self.loss_G_A + self.loss_G_B + self.loss_cycle_A + self.loss_cycle_B +
                       self.loss_idt_A + self.loss_idt_B)
        self.loss_G.backward()

    # -- CUDA-graph replay of the step ----------------------------------------------------------------
    def enable_cuda_graphs(self, warmup_steps=2):
        """Replay the training step as two CUDA graphs (generator phase, discriminator phase) instead of ~2400
        eager launches.  The image pools stay eager between the two (their control flow depends on the host RNG,
        cycle_gan_model.py:8-35).  The next ``warmup_steps`` calls of optimize_parameters() still run eagerly (they
        warm every kernel / allocation up), the one after is captured.
        Under data parallelism (parallel.attach) the bucketed NCCL all-reduces are issued from the gradient hooks
        while the backward pass is being captured, so they become nodes of the same graphs (the capture runs in
        thread-local error mode: NCCL's watchdog thread must not trip it)."""
        self._graphs = {"warm": int(warmup_steps), "shape": None}

    def _phase_G(self):
        sync = self.grad_sync
        self.forward()
        self.set_requires_grad([self.netD_A, self.netD_B], False)
        if sync:
            sync.zero("G")                         # gradients are views into the flat buckets: one fill per bucket
            sync.begin("G")
        else:
            self.optimizer_G.zero_grad(set_to_none=True)
        self.backward_G()
        if sync:
            sync.finish("G")
        self.optimizer_G.step()

    def _phase_D(self, pooled_B, pooled_A):
        sync = self.grad_sync
        self.set_requires_grad([self.netD_A, self.netD_B], True)
        if sync:
            sync.zero("D")
            sync.begin("D")
        else:
            self.optimizer_D.zero_grad(set_to_none=True)
        self.loss_D_A = self.backward_D_basic(self.netD_A, self.real_B, pooled_B)
        self.loss_D_B = self.backward_D_basic(self.netD_B, self.real_A, pooled_A)
        if sync:
            sync.finish("D")
        self.optimizer_D.step()

    def _optimize_graphed(self):
        G = self._graphs
        A, B = self.real_A, self.real_B
        shape = (tuple(A.shape), tuple(B.shape))
        if G.get("gG") is None or G["shape"] != shape:
            if G["warm"] > 0:
                G["warm"] -= 1
                return False
            # capture: static inputs, both phases in one memory pool
            G["shape"] = shape
            G["sA"], G["sB"] = A.clone(), B.clone()
            self.real_A, self.real_B = G["sA"], G["sB"]
            # drop every reference to the eager steps' autograd graphs: their AccumulateGrad nodes live on the default
            # stream and must not be reused inside the capture
            for n in self.loss_names:
                setattr(self, "loss_" + n, None)
            self.loss_G = self.loss_cor_coe_GA = self.loss_cor_coe_GB = None
            for n in ("fake_A", "fake_B", "rec_A", "rec_B", "idt_A", "idt_B"):
                setattr(self, n, None)
            if self.grad_sync is None:
                self.optimizer_G.zero_grad(set_to_none=True)
                self.optimizer_D.zero_grad(set_to_none=True)
            gc.collect()
            torch.cuda.synchronize()
            # capture only records: each graph is replayed right after its capture to execute this step (the host
            # side of optimizer.step() -- the step counters -- already ran during capture; Adam's own counter is bumped on the
            # device by the captured mra_adam_advance)
            from .. import ops
            n0 = ops.impl().launch_count() if hasattr(ops.impl(), "launch_count") else 0
            mode = dict(capture_error_mode="thread_local") if self.grad_sync is not None else {}
            G["gG"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(G["gG"], **mode):
                self._phase_G()
            G["gG"].replay()
            G["pB"] = torch.empty_like(self.fake_B)
            G["pA"] = torch.empty_like(self.fake_A)
            G["pB"].copy_(self.fake_B_pool.query(self.fake_B))
            G["pA"].copy_(self.fake_A_pool.query(self.fake_A))
            G["gD"] = torch.cuda.CUDAGraph()
            with torch.cuda.graph(G["gD"], pool=G["gG"].pool(), **mode):
                self._phase_D(G["pB"], G["pA"])
            G["gD"].replay()
            G["launches"] = (ops.impl().launch_count() - n0) if hasattr(ops.impl(), "launch_count") else 0   # kernels per replayed step
            return True
        G["sA"].copy_(A, non_blocking=True)
        G["sB"].copy_(B, non_blocking=True)
        self.real_A, self.real_B = G["sA"], G["sB"]
        self.optimizer_G.advance_host_state()
        G["gG"].replay()
        G["pB"].copy_(self.fake_B_pool.query(self.fake_B))
        G["pA"].copy_(self.fake_A_pool.query(self.fake_A))
        self.optimizer_D.advance_host_state()
        G["gD"].replay()
        return True

    def optimize_parameters(self):
        if self._graphs is not None and self._optimize_graphed():
            return
        sync = self.grad_sync
        self.forward()
        self.set_requires_grad([self.netD_A, self.netD_B], False)
        if sync:
            sync.zero("G")
            sync.begin("G")
        else:
            self.optimizer_G.zero_grad(set_to_none=True)
        self.backward_G()
        if sync:
            sync.finish("G")
        self.optimizer_G.step()
        self.set_requires_grad([self.netD_A, self.netD_B], True)
        if sync:
            sync.zero("D")
            sync.begin("D")
        else:
            self.optimizer_D.zero_grad(set_to_none=True)
        self.backward_D_A()
        self.backward_D_B()
        if sync:
            sync.finish("D")
        self.optimizer_D.step()
