"""Put this directory first on sys.path and the reference's `ddim2/main2.py`
(`from diff_model2 import *`) resolves to the B200 implementation."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import advshadow_b200  # noqa: E402,F401
from advshadow_b200.diff_model2 import *  # noqa: E402,F401,F403
from advshadow_b200.diff_model2 import (GaussianDiffusion, PretrainedResNet50, UNetModel,  # noqa: E402,F401
                                        timestep_embedding, norm_layer)
