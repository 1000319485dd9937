"""mra_gan_b200 -- B200-native (sm_100a) 3D CycleGAN training / inference hot path.

Drop-in for pedrob37/MRA-GAN's ``models/networks3D.py`` + ``models/cycle_gan_model.py`` surface
(define_G / define_D, CycleGANModel.set_input / optimize_parameters / test, create_model), with all
arithmetic in hand-written CUDA behind the C ABI of include/mra_gan_b200.h.
"""
__version__ = "0.1.0"
