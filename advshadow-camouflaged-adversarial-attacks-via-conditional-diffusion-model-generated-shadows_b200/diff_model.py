"""Drop-in replacement of the reference's `diff_model.py` (`from diff_model import *`, main.py:6).

Same public names, constructor signatures, defaults, state_dict keys and return types as
diff_model.py:16-484; the denoiser and the sampling loops execute on the sm_100a kernels.
"""
from ._compat import *  # noqa: F401,F403  (re-exported on purpose, see _compat.EXPORTS)
from ._compat import math, torch
from ._diffusion import (GaussianDiffusionBase, cosine_beta_schedule, ddim_timestep_tables,  # noqa: F401
                         linear_beta_schedule)
from ._model import UNetModelBase


def timestep_embedding(timesteps, dim, max_period=10000):
    """Sinusoidal embedding [cos | sin] (dm1:16-33).  CUDA tensors use the library kernel."""
    from . import ops
    return ops.timestep_embedding(timesteps, dim, max_period)


def norm_layer(channels):
    return torch.nn.GroupNorm(32, channels)


class UNetModel(UNetModelBase):
    def __init__(self, in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
                 attention_resolutions=(8, 16), dropout=0, channel_mult=(1, 2, 2, 2), conv_resample=True,
                 num_heads=4):
        super().__init__(in_channels=in_channels, model_channels=model_channels, out_channels=out_channels,
                         num_res_blocks=num_res_blocks, attention_resolutions=attention_resolutions,
                         dropout=dropout, channel_mult=channel_mult, conv_resample=conv_resample,
                         num_heads=num_heads)


class GaussianDiffusion(GaussianDiffusionBase):
    _DEFAULT_SCHEDULE = 'cosine'          # dm1:290

    def __init__(self, timesteps=1000, beta_schedule='cosine'):
        super().__init__(timesteps, beta_schedule)
