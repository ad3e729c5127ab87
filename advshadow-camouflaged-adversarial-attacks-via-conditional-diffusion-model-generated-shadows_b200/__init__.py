"""advshadow_b200 -- B200-native (sm_100a) implementation of AdvShadow's diffusion shadow sampler.

Public surface:
  diff_model / diff_model2   drop-in modules for the reference's `from diff_model import *`
  shadow                     batched shadow-mask / compositing kernels
  ops                        tensor-level wrappers of single C-ABI entry points
  attack                     data-parallel attack-loop harness (success flags + ASR counts)
  datasets                   on-disk formats: image / mask_<name> / image_labels.json / id2label readers, writer
  metrics                    evaluation metrics of the reference's scripts: FID, SSIM, PSNR (host side)
  iddm                       IDDM class-conditional UNet + CFG DDIM sampler (model/networks/unet.py, model/samples/ddim.py)
The CUDA kernels live in csrc/ and are reached only through the C ABI in include/advshadow_b200.h.
"""
from . import _capi  # noqa: F401

__all__ = ["diff_model", "diff_model2", "shadow", "ops", "attack", "iddm", "datasets", "metrics", "sampler", "plan", "engine"]
__version__ = "0.1.0"


def __getattr__(name):
    if name in __all__:
        import importlib
        return importlib.import_module("." + name, __name__)
    raise AttributeError(name)
