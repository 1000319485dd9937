"""Import the UNMODIFIED reference from /root/reference (build container only).

The reference is pure Python on top of torch; two third-party imports are missing in this
image and are stubbed because the hot path never touches them:
  * ``monai``      -- models/networks3D.py:6, used only inside ``Dynet()`` (networks3D.py:346-377)
  * ``SimpleITK``  -- utils/NiftiDataset.py:1, only needed to import test.py
``BaseOptions.parse()`` calls ``torch.cuda.set_device(0)`` (options/base_options.py:123) and so
cannot run without a GPU; ``make_opt`` builds the same Namespace by hand.

Nothing on the GPU box may call this module (the reference tree does not travel).
"""
import argparse
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("MRA_REFERENCE_ROOT", "/root/reference")


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "models", "networks3D.py"))


def import_reference():
    """Returns (networks3D, cycle_gan_model, test_model, models_pkg) of the reference."""
    if not reference_available():
        raise RuntimeError("reference tree not present at %s" % REFERENCE_ROOT)
    sys.dont_write_bytecode = True  # the tree is read-only
    if "monai" not in sys.modules:
        sys.modules["monai"] = types.ModuleType("monai")
    if "SimpleITK" not in sys.modules:
        sitk = types.ModuleType("SimpleITK")
        sitk.sitkLinear = 2
        sys.modules["SimpleITK"] = sitk
    # The reference's top-level package is literally called ``models`` -- keep it isolated
    # from anything of ours by importing it under its own sys.path entry.
    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    import importlib
    models_pkg = importlib.import_module("models")
    networks3D = importlib.import_module("models.networks3D")
    cycle = importlib.import_module("models.cycle_gan_model")
    testm = importlib.import_module("models.test_model")
    return networks3D, cycle, testm, models_pkg


def make_opt(**overrides):
    """The Namespace ``TrainOptions().parse()`` would give for BASELINE config 1
    (options/base_options.py:12-57, options/train_options.py:5-26,
    models/cycle_gan_model.py:42-62), with LSGAN selected (no_lsgan=False)."""
    opt = argparse.Namespace(
        gpu_ids=0, isTrain=True, checkpoints_dir="/tmp/mra_oracle_ckpt", name="oracle",
        input_nc=1, output_nc=1, ngf=64, ndf=64, netG="resnet_9blocks", netD="n_layers",
        n_layers_D=3, norm="instance", no_dropout=True, init_type="normal", init_gain=0.02,
        no_lsgan=False, pool_size=50, lr=2e-4, beta1=0.5, lambda_A=10.0, lambda_B=10.0,
        lambda_identity=0.5, lambda_co_A=2, lambda_co_B=2, which_direction="AtoB",
        lr_policy="lambda", epoch_count=1, niter=500, niter_decay=100, lr_decay_iters=50,
        continue_train=False, which_epoch="latest", verbose=False, model="cycle_gan",
        model_suffix="", batch_size=1)
    for k, v in overrides.items():
        setattr(opt, k, v)
    return opt
