"""Put this directory first on sys.path and the reference's `main.py` (`from diff_model import *`)
runs unchanged on the B200 kernels."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import advshadow_b200  # noqa: E402,F401
from advshadow_b200.diff_model import *  # noqa: E402,F401,F403
from advshadow_b200.diff_model import GaussianDiffusion, UNetModel, timestep_embedding, norm_layer  # noqa: E402,F401
