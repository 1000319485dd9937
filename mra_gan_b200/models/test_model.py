"""TestModel: single-generator inference entry used by test.py (reference models/test_model.py:7-48)."""
import torch

from .. import networks3D
from .base_model import BaseModel
from .cycle_gan_model import CycleGANModel


class TestModel(BaseModel):
    def name(self):
        return "TestModel"

    @staticmethod
    def modify_commandline_options(parser, is_train=True):
        assert not is_train, "TestModel cannot be used in train mode"
        parser = CycleGANModel.modify_commandline_options(parser, is_train=False)
        parser.set_defaults(dataset_mode="single")
        parser.add_argument("--model_suffix", type=str, default="",
                            help="[which_epoch]_net_G[model_suffix].pth in checkpoints_dir is loaded as the generator")
        return parser

    def initialize(self, opt):
        assert not opt.isTrain
        BaseModel.initialize(self, opt)
        self.loss_names = []
        self.visual_names = ["real_A", "fake_B"]
        self.model_names = ["G" + opt.model_suffix]
        self.netG = networks3D.define_G(opt.input_nc, opt.output_nc, opt.ngf, opt.netG, opt.norm, not opt.no_dropout,
                                        opt.init_type, opt.init_gain, self.gpu_ids)
        setattr(self, "netG" + opt.model_suffix, self.netG)     # so load_networks finds it under its file name

    def set_input(self, input):
        self.real_A = input.to(self.device)

    def forward(self):
        # NB: like the reference, the generator stays in train mode (instance statistics per window)
        self.fake_B = self.netG(self.real_A)
