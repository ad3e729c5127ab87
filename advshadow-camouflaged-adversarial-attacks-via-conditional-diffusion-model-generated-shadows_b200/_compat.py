"""Names the reference modules leak through `from diff_model import *` (dm1:1-14, dm2:1-16).
main.py / ddim2/main2.py rely on them (os, Dataset, DataLoader, Image, transforms, torch, tqdm,
plt, np, F ...), so the drop-in modules re-export the same set.  Optional packages that are absent
become lazy stubs that raise on first use instead of at import."""
import math  # noqa: F401
import os  # noqa: F401
from abc import abstractmethod  # noqa: F401

import numpy as np  # noqa: F401
import torch  # noqa: F401
import torch.nn as nn  # noqa: F401
import torch.nn.functional as F  # noqa: F401
from torch.utils.data import DataLoader, Dataset  # noqa: F401
from tqdm import tqdm  # noqa: F401


class _Missing:
    def __init__(self, name):
        self.__dict__["_name"] = name

    def __getattr__(self, item):
        raise ImportError(f"optional dependency '{self._name}' is not installed (needed for .{item})")


def _optional(modname, attr=None):
    try:
        mod = __import__(modname, fromlist=["*"])
        return getattr(mod, attr) if attr else mod
    except Exception:
        return _Missing(modname)


Image = _optional("PIL.Image")
requests = _optional("requests")
models = _optional("torchvision.models")
datasets = _optional("torchvision.datasets")
transforms = _optional("torchvision.transforms")
plt = _optional("matplotlib.pyplot")
PILImage = _optional("fastai.vision.core", "PILImage")

EXPORTS = ["os", "math", "abstractmethod", "Dataset", "DataLoader", "Image", "requests", "np", "torch", "nn", "F",
           "models", "datasets", "transforms", "tqdm", "plt"]
