"""UNetModel with the reference's constructor, attributes and state_dict, executed by UNetEngine.

Reference: diff_model.py:157-267 (and the ddim2/diff_model2.py:195-305 copy with other defaults).
The module tree below exists to (a) own the parameters under the reference's exact state_dict key
names / OIHW fp32 shapes, so `model.load_state_dict(torch.load(path))` (main.py:115) and
`torch.save(model.state_dict())` (dm1:574) interoperate, and (b) consume torch's RNG in the
reference's construction order so `torch.manual_seed(s); UNetModel()` yields identical weights.
None of these leaf modules is ever *called*: forward() hands the parameters to the CUDA engine.
"""
import warnings

import torch
import torch.nn as nn

from .plan import UNetSpec


class _Holder(nn.Module):
    """Numbered child container (state_dict-compatible with nn.Sequential); not callable."""

    def __init__(self, *mods):
        super().__init__()
        for i, m in enumerate(mods):
            self.add_module(str(i), m)

    def forward(self, *a, **k):
        raise RuntimeError("advshadow_b200 sub-modules are parameter holders; call UNetModel.forward")


def _norm(ch):
    return nn.GroupNorm(32, ch)          # dm1:62-63


class _Residual(nn.Module):               # parameter layout of ResidualBlock, dm1:67-92
    def __init__(self, cin, cout, tdim, dropout):
        super().__init__()
        self.conv1 = _Holder(_norm(cin), nn.SiLU(), nn.Conv2d(cin, cout, kernel_size=3, padding=1))
        self.time_emb = _Holder(nn.SiLU(), nn.Linear(tdim, cout))
        self.conv2 = _Holder(_norm(cout), nn.SiLU(), nn.Dropout(p=dropout),
                             nn.Conv2d(cout, cout, kernel_size=3, padding=1))
        self.shortcut = nn.Conv2d(cin, cout, kernel_size=1) if cin != cout else nn.Identity()


class _Attention(nn.Module):              # AttentionBlock, dm1:107-115
    def __init__(self, ch, num_heads):
        super().__init__()
        self.num_heads = num_heads
        assert ch % num_heads == 0
        self.norm = _norm(ch)
        self.qkv = nn.Conv2d(ch, ch * 3, kernel_size=1, bias=False)
        self.proj = nn.Conv2d(ch, ch, kernel_size=1)


class _Upsample(nn.Module):               # dm1:129-134
    def __init__(self, ch, use_conv):
        super().__init__()
        self.use_conv = use_conv
        if use_conv:
            self.conv = nn.Conv2d(ch, ch, kernel_size=3, padding=1)


class _Downsample(nn.Module):             # dm1:143-150
    def __init__(self, ch, use_conv):
        super().__init__()
        self.use_conv = use_conv
        if use_conv:
            self.op = nn.Conv2d(ch, ch, kernel_size=3, stride=2, padding=1)
        else:
            self.op = nn.AvgPool2d(stride=2)   # raises TypeError exactly like the reference (dm1:150)


class UNetModelBase(nn.Module):
    """Drop-in UNetModel.  Inference only: forward() needs CUDA tensors and returns eps computed by
    the sm_100a kernels.  `precision` ('bf16' | 'fp32') selects the throughput or the <=1e-4 mode."""

    _DEFAULTS = dict(in_channels=3, model_channels=128, out_channels=3, num_res_blocks=2,
                     attention_resolutions=(8, 16), dropout=0, channel_mult=(1, 2, 2, 2),
                     conv_resample=True, num_heads=4)

    def __init__(self, **kw):
        super().__init__()
        cfg = dict(self._DEFAULTS)
        cfg.update(kw)
        (in_channels, model_channels, out_channels, num_res_blocks, attention_resolutions, dropout,
         channel_mult, conv_resample, num_heads) = (cfg[k] for k in (
             "in_channels", "model_channels", "out_channels", "num_res_blocks", "attention_resolutions",
             "dropout", "channel_mult", "conv_resample", "num_heads"))
        self.in_channels = in_channels
        self.model_channels = model_channels
        self.out_channels = out_channels
        self.num_res_blocks = num_res_blocks
        self.attention_resolutions = attention_resolutions
        self.dropout = dropout
        self.channel_mult = channel_mult
        self.conv_resample = conv_resample
        self.num_heads = num_heads

        tdim = model_channels * 4
        self.time_embed = _Holder(nn.Linear(model_channels, tdim), nn.SiLU(), nn.Linear(tdim, tdim))
        self.down_blocks = nn.ModuleList([_Holder(nn.Conv2d(in_channels, model_channels, kernel_size=3, padding=1))])
        chans = [model_channels]
        ch, ds = model_channels, 1
        for level, mult in enumerate(channel_mult):
            for _ in range(num_res_blocks):
                layers = [_Residual(ch, mult * model_channels, tdim, dropout)]
                ch = mult * model_channels
                if ds in attention_resolutions:
                    layers.append(_Attention(ch, num_heads))
                self.down_blocks.append(_Holder(*layers))
                chans.append(ch)
            if level != len(channel_mult) - 1:
                self.down_blocks.append(_Holder(_Downsample(ch, conv_resample)))
                chans.append(ch)
                ds *= 2
        self.middle_block = _Holder(_Residual(ch, ch, tdim, dropout), _Attention(ch, num_heads),
                                    _Residual(ch, ch, tdim, dropout))
        self.up_blocks = nn.ModuleList([])
        for level, mult in list(enumerate(channel_mult))[::-1]:
            for i in range(num_res_blocks + 1):
                layers = [_Residual(ch + chans.pop(), model_channels * mult, tdim, dropout)]
                ch = model_channels * mult
                if ds in attention_resolutions:
                    layers.append(_Attention(ch, num_heads))
                if level and i == num_res_blocks:
                    layers.append(_Upsample(ch, conv_resample))
                    ds //= 2
                self.up_blocks.append(_Holder(*layers))
        self.out = _Holder(_norm(ch), nn.SiLU(), nn.Conv2d(model_channels, out_channels, kernel_size=3, padding=1))

        # ---- engine state (not part of state_dict) ----
        self.precision = "bf16"
        self._engines = {}
        self._warned_dropout = False

    # ---- B200 execution ----
    def spec(self) -> UNetSpec:
        return UNetSpec(in_channels=self.in_channels, model_channels=self.model_channels,
                        out_channels=self.out_channels, num_res_blocks=self.num_res_blocks,
                        attention_resolutions=tuple(self.attention_resolutions),
                        channel_mult=tuple(self.channel_mult), conv_resample=bool(self.conv_resample),
                        num_heads=self.num_heads)

    def set_precision(self, precision: str):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision
        return self

    def _weights_token(self):
        return tuple((p.data_ptr(), p._version) for p in self.parameters())

    def engine(self, B, H, W, precision=None, conv_impl="auto", attn_impl="auto", fuse_gn_stats=True,
               fuse_upsample=True, instance=0, wide_prenorm=2, gemm_operands="fp16", fp16_levels=2):
        """The (cached) UNetEngine for this input geometry; repacks weights if parameters changed.
        `instance` distinguishes independent engines of the same geometry (own arenas, for multi-stream use)."""
        from .engine import UNetEngine
        precision = precision or self.precision
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("advshadow_b200.UNetModel runs on CUDA only (there is no CPU path); "
                               "call model.to('cuda') first")
        if self.training and self.dropout and not self._warned_dropout:
            warnings.warn("UNetModel is in train() mode with dropout>0: the B200 inference path treats "
                          "Dropout as identity (call model.eval(), as main.py:116 does)")
            self._warned_dropout = True
        key = (B, H, W, precision, conv_impl, attn_impl, fuse_gn_stats, fuse_upsample, dev.index, instance, wide_prenorm,
               gemm_operands, fp16_levels)
        params = dict(self.named_parameters())
        tok = self._weights_token()
        hit = self._engines.get(key)
        if hit is None:
            eng = UNetEngine(self.spec(), params, B, H, W, precision, conv_impl, attn_impl, fuse_gn_stats, fuse_upsample,
                             wide_prenorm, gemm_operands, fp16_levels)
            self._engines[key] = [eng, tok]
            return eng
        if hit[1] != tok:
            hit[0].load_weights(params)
            hit[1] = tok
        return hit[0]

    def release_engines(self):
        self._engines.clear()

    def _apply(self, fn, *a, **k):   # .to()/.cuda()/.half(): cached engines point at stale storage
        self._engines = {}
        return super()._apply(fn, *a, **k)

    def forward(self, x, timesteps):
        """eps = model(x[N,C,H,W], timesteps[N])  (dm1:245-267).  CUDA, inference only."""
        if not x.is_cuda:
            raise RuntimeError("advshadow_b200.UNetModel.forward needs CUDA tensors (no CPU path)")
        if torch.is_grad_enabled() and self.training and any(p.requires_grad for p in self.parameters()):
            raise NotImplementedError("advshadow_b200.UNetModel is inference-only: wrap the call in "
                                      "torch.no_grad() or call model.eval(); training is out of scope")
        B, _, H, W = x.shape
        eng = self.engine(B, H, W)
        return eng.forward(x.float(), timesteps)
