"""IDDM class-conditional UNet + classifier-free-guidance DDIM sampler on the B200 kernels.

Reference: model/networks/unet.py:17-128 (`UNet`), model/networks/base.py:12-68 (`BaseNet`: label embedding, sin|cos
position encoding), model/modules/{conv,block,attention}.py, model/samples/{base,ddim}.py (`DDIMDiffusion.sample`).
Same constructor arguments, attributes and state_dict keys as the reference classes; forward / sample run on the C-ABI
kernels (implicit-GEMM convs and token linears, attention, GroupNorm(1,C), LayerNorm, MaxPool, bilinear upsample, CFG lerp).
In 16-bit mode every conv / token linear with 64-aligned channels and every attention block run on the tcgen05 kernels:
the 4-head nn.MultiheadAttention has head dims 16 / 32 / 64, which run as dh = 64 with zero-padded q / k / v^T
(advs_attention_sm100_plan_ex); conv outputs that a GroupNorm reads carry the int8 mantissa extension, and the bounded
tensors (GroupNorm+activation and LayerNorm outputs) and their weights are fp16 GEMM operands, as on the diff_model path.
Classifier-free guidance evaluates the conditional and the unconditional prediction as ONE forward over 2n rows.
`save_ckpt` / `load_ckpt` read and write the reference's checkpoint dict (utils/checkpoint.py:21-157).
CUDA only, inference only -- no fallback.
"""
import ctypes as C
import os
import math

import torch
import torch.nn as nn
from tqdm import tqdm

from . import _capi as capi

_ACT = {"silu": 1, "gelu": 2}


# ADVS_IDDM_GRAPH=0 runs the sampling loop eagerly (debugging)
_USE_GRAPH = os.environ.get("ADVS_IDDM_GRAPH", "1") != "0"


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


# ---------------------------------------------------------------------------------------------------------------
# parameter holders with the reference's module tree (=> identical state_dict keys and seeded initialisation)
# ---------------------------------------------------------------------------------------------------------------
class _Holder(nn.Module):
    def forward(self, *a, **k):
        raise RuntimeError("advshadow_b200.iddm sub-modules are parameter holders; call UNet.forward")


def _act_module(name):
    return {"relu": nn.ReLU(), "relu6": nn.ReLU6(), "silu": nn.SiLU(), "lrelu": nn.LeakyReLU(0.1), "gelu": nn.GELU()}.get(
        name, nn.SiLU())


class DoubleConv(_Holder):                      # conv.py:20-66
    def __init__(self, in_channels, out_channels, mid_channels=None, residual=False, act="silu"):
        super().__init__()
        self.residual = residual
        mid_channels = mid_channels or out_channels
        self.act = act
        self.double_conv = nn.Sequential(
            nn.Conv2d(in_channels, mid_channels, kernel_size=3, padding=1, bias=False), nn.GroupNorm(1, mid_channels),
            _act_module(act),
            nn.Conv2d(mid_channels, out_channels, kernel_size=3, padding=1, bias=False), nn.GroupNorm(1, out_channels))


class DownBlock(_Holder):                       # block.py:15-46
    def __init__(self, in_channels, out_channels, emb_channels=256, act="silu"):
        super().__init__()
        self.maxpool_conv = nn.Sequential(nn.MaxPool2d(2), DoubleConv(in_channels, in_channels, residual=True, act=act),
                                          DoubleConv(in_channels, out_channels, act=act))
        self.emb_layer = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, out_channels))


class UpBlock(_Holder):                         # block.py:49-78
    def __init__(self, in_channels, out_channels, emb_channels=256, act="silu"):
        super().__init__()
        self.up = nn.Upsample(scale_factor=2, mode="bilinear", align_corners=True)
        self.conv = nn.Sequential(DoubleConv(in_channels, in_channels, residual=True, act=act),
                                  DoubleConv(in_channels, out_channels, mid_channels=in_channels // 2, act=act))
        self.emb_layer = nn.Sequential(nn.SiLU(), nn.Linear(emb_channels, out_channels))


class SelfAttention(_Holder):                   # attention.py:12-53
    def __init__(self, channels, size, act="silu"):
        super().__init__()
        self.channels, self.size = channels, size
        self.mha = nn.MultiheadAttention(embed_dim=channels, num_heads=4, batch_first=True)
        self.ln = nn.LayerNorm([channels])
        self.ff_self = nn.Sequential(nn.LayerNorm([channels]), nn.Linear(channels, channels), _act_module(act),
                                     nn.Linear(channels, channels))


class UNet(nn.Module):
    """Drop-in for model/networks/unet.py `UNet` (label-conditional when num_classes is given)."""

    def __init__(self, in_channel=3, out_channel=3, channel=None, time_channel=256, num_classes=None, image_size=64,
                 device="cpu", act="silu"):
        super().__init__()
        self.in_channel, self.out_channel = in_channel, out_channel
        self.channel = channel if channel is not None else [32, 64, 128, 256, 512, 1024]   # base.py:44-53
        self.time_channel, self.num_classes, self.image_size, self.device, self.act = (time_channel, num_classes,
                                                                                       image_size, device, act)
        if num_classes is not None:
            self.label_emb = nn.Embedding(num_embeddings=num_classes, embedding_dim=time_channel)
        ch = self.channel
        self.inc = DoubleConv(in_channel, ch[1], act=act)
        self.down1 = DownBlock(ch[1], ch[2], act=act)
        self.sa1 = SelfAttention(ch[2], int(image_size / 2), act=act)
        self.down2 = DownBlock(ch[2], ch[3], act=act)
        self.sa2 = SelfAttention(ch[3], int(image_size / 4), act=act)
        self.down3 = DownBlock(ch[3], ch[3], act=act)
        self.sa3 = SelfAttention(ch[3], int(image_size / 8), act=act)
        self.bot1 = DoubleConv(ch[3], ch[4], act=act)
        self.bot2 = DoubleConv(ch[4], ch[4], act=act)
        self.bot3 = DoubleConv(ch[4], ch[3], act=act)
        self.up1 = UpBlock(ch[4], ch[2], act=act)
        self.sa4 = SelfAttention(ch[2], int(image_size / 4), act=act)
        self.up2 = UpBlock(ch[3], ch[1], act=act)
        self.sa5 = SelfAttention(ch[1], int(image_size / 2), act=act)
        self.up3 = UpBlock(ch[2], ch[1], act=act)
        self.sa6 = SelfAttention(ch[1], int(image_size), act=act)
        self.outc = nn.Conv2d(ch[1], out_channel, kernel_size=1)
        self.precision = "fp32"     # parity-first default on this path; "bf16" is opt-in
        self._engines = {}

    def set_precision(self, precision):
        if precision not in ("bf16", "fp32"):
            raise ValueError("precision must be 'bf16' or 'fp32'")
        self.precision = precision
        return self

    def _apply(self, fn, *a, **k):
        self._engines = {}
        return super()._apply(fn, *a, **k)

    def release_engines(self):
        self._engines.clear()

    def engine(self, B, precision=None):
        precision = precision or self.precision
        dev = next(self.parameters()).device
        if dev.type != "cuda":
            raise RuntimeError("advshadow_b200.iddm.UNet runs on CUDA only (there is no CPU path)")
        if self.act not in _ACT:
            raise NotImplementedError(f"activation {self.act!r}: the B200 path implements 'silu' and 'gelu'")
        tok = tuple((p.data_ptr(), p._version) for p in self.parameters())
        key = (B, precision, dev.index)
        hit = self._engines.get(key)
        if hit is None or hit[1] != tok:
            self._engines[key] = (_IddmEngine(self, B, precision), tok)
        return self._engines[key][0]

    def forward(self, x, time, y=None):
        """eps = UNet(x[N,3,S,S], time[N], y[N] | None)  (unet.py:95-128).  CUDA, inference only."""
        if not x.is_cuda:
            raise RuntimeError("advshadow_b200.iddm.UNet.forward needs CUDA tensors (no CPU path)")
        if x.shape[-1] != self.image_size or x.shape[-2] != self.image_size:
            raise ValueError(f"SelfAttention hard-codes image_size={self.image_size} (attention.py:46); got {tuple(x.shape)}")
        return self.engine(x.shape[0]).forward(x.float(), time, y)


# ---------------------------------------------------------------------------------------------------------------
class _IddmEngine:
    """Static buffers + launch list for one (batch, precision); built by walking the reference's forward."""

    def __init__(self, net: UNet, B, precision):
        self.lib = capi.lib()
        self.net, self.B, self.precision = net, B, precision
        self.dev = next(net.parameters()).device
        self.dt = capi.BF16 if precision == "bf16" else capi.F32
        self.tdt = torch.bfloat16 if precision == "bf16" else torch.float32
        with torch.cuda.device(self.dev):
            self.sm100 = precision == "bf16" and bool(self.lib.advs_device_is_sm100())
        self.act = _ACT[net.act]
        self.keep, self.L, self.attn_kinds = [], [], []
        S, tc = net.image_size, net.time_channel
        dev = self.dev
        self.x = torch.zeros(B, net.in_channel, S, S, dtype=torch.float32, device=dev)
        self.eps = torch.zeros(B, net.out_channel, S, S, dtype=torch.float32, device=dev)
        self.t = torch.zeros(B, dtype=torch.int64, device=dev)
        self.y = torch.zeros(B, dtype=torch.int64, device=dev)
        self.time = torch.zeros(B, tc, dtype=torch.float32, device=dev)
        # base.py:63: 1 / 10000^(arange(0, C, 2) / C), evaluated like the reference and uploaded
        self.inv_freq = (1.0 / (10000 ** (torch.arange(start=0, end=tc, step=2).float() / tc))).to(dev)
        self.label_w = net.label_emb.weight.detach().float().contiguous() if net.num_classes is not None else None
        with torch.cuda.device(dev):
            self._build()

    # ---- helpers -------------------------------------------------------------------------------------------
    def buf(self, *shape, dtype=None):
        t = torch.empty(*shape, dtype=dtype or self.tdt, device=self.dev)
        self.keep.append(t)
        return t

    def add(self, fn_name, *args):
        self.L.append((getattr(self.lib, fn_name), args, fn_name))

    def pack(self, w4d, f16=False):
        """OIHW fp32 parameter -> [O][taps][I] in the engine dtype (fp16 when it multiplies an fp16 operand)."""
        w = w4d.detach().float().contiguous()
        O, I, kh, kw = w.shape
        dst = self.buf(O, kh * kw, I, dtype=torch.float16 if f16 else None)
        capi.call("advs_pack_conv_weight", w.data_ptr(), dst.data_ptr(), O, I, kh, kw, capi.F16 if f16 else self.dt, _st())
        self.keep.append(w)
        return dst

    def f32(self, p):
        t = p.detach().float().contiguous()
        self.keep.append(t)
        return t

    def on_sm100(self, cin, cout, dh=None):
        return self.sm100 and cin % 64 == 0 and cout % 64 == 0 and (dh is None or dh % 32 == 0 or dh == 16)

    def conv(self, x, B, H, W, cin, w_packed, cout, bias=None, residual=None, qkv_heads=None, nchw_out=None, cout_valid=None,
             want_lo=False):
        """Returns the output (or (q, k, vt)); with want_lo also the int8 mantissa extension (16-bit mode) or None."""
        cp = capi.ConvParams()
        taps = w_packed.shape[1]
        cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, w_packed.shape[0], 1, 1
        cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = x.data_ptr(), w_packed.data_ptr(), cin, taps
        cp.bias = bias.data_ptr() if bias is not None else None
        cp.residual = residual.data_ptr() if residual is not None else None
        cp.dtype = self.dt
        if w_packed.dtype == torch.float16:      # fp16 x fp16 (the activations were written as fp16 by the producer)
            cp.operand_f16 = 0b0101
        out, lo, dh = None, None, None
        if qkv_heads:
            dh = cout // (3 * qkv_heads)
            T = H * W
            # head dims below 64 are stored 64 wide (zero beyond dh) for the tcgen05 attention kernel
            pad = 64 if (dh < 64 and self.on_sm100(cin, w_packed.shape[0], dh) and T % 8 == 0) else 0
            dhs = max(dh, pad)
            q, k, vt = (torch.zeros(B, qkv_heads, T, dhs, dtype=self.tdt, device=self.dev),
                        torch.zeros(B, qkv_heads, T, dhs, dtype=self.tdt, device=self.dev),
                        torch.zeros(B, qkv_heads, dhs, T, dtype=self.tdt, device=self.dev))
            self.keep += [q, k, vt]
            cp.out_mode, cp.q, cp.k, cp.vt, cp.heads, cp.qkv_dh_pad = 1, q.data_ptr(), k.data_ptr(), vt.data_ptr(), qkv_heads, pad
            cp.qk_scale = 1.0 / math.sqrt(math.sqrt(dh))     # MHA scales q by dh^-1/2; split evenly over q and k
            out = (q, k, vt)
        elif nchw_out is not None:
            cp.out_mode, cp.y, cp.cout_valid = 2, nchw_out.data_ptr(), cout_valid
            out = nchw_out
        else:
            out = self.buf(B, H, W, cout)
            cp.out_mode, cp.y = 0, out.data_ptr()
            if want_lo and self.precision == "bf16" and cout % 32 == 0:
                lo = self.buf(B, H, W, cout, dtype=torch.int8)
                cp.y_lo = lo.data_ptr()
        self.keep.append(cp)
        if self.on_sm100(cin, w_packed.shape[0], dh):
            pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
            capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
            self.keep.append(pb)
            self.add("advs_conv_sm100_launch", pb.ptr)
        else:
            assert w_packed.dtype != torch.float16, "fp16 operands exist on the tcgen05 path only"
            self.add("advs_conv_simt", C.byref(cp))
        return (out, lo) if want_lo else out

    def gn(self, x, B, HW, Cc, gnmod, act=0, residual=None, emb=None, x_lo=None, out_f16=False):
        ss = self.buf(B, Cc, 2, dtype=torch.float32)
        wsb = int(self.lib.advs_groupnorm_workspace_bytes(B, HW, Cc))
        ws = self.buf(max(wsb, 4), dtype=torch.uint8)
        g, b = self.f32(gnmod.weight), self.f32(gnmod.bias)
        self.add("advs_groupnorm_stats", x.data_ptr(), Cc, None, 0, B, HW, 1, float(gnmod.eps), g.data_ptr(), b.data_ptr(),
                 ss.data_ptr(), ws.data_ptr(), wsb, self.dt)
        y = self.buf(*x.shape, dtype=torch.float16 if out_f16 else None)
        rp = residual.data_ptr() if residual is not None else None
        ep = emb.data_ptr() if emb is not None else None
        if self.precision == "bf16":
            self.add("advs_groupnorm_apply_ex16", x.data_ptr(), x_lo.data_ptr() if x_lo is not None else None, B, HW, Cc,
                     ss.data_ptr(), rp, ep, Cc if emb is not None else 0, act, y.data_ptr(), capi.F16 if out_f16 else capi.BF16)
        else:
            self.add("advs_groupnorm_apply_ex", x.data_ptr(), B, HW, Cc, ss.data_ptr(), rp, ep, Cc if emb is not None else 0,
                     act, y.data_ptr(), self.dt)
        return y

    def double_conv(self, x, B, H, W, cin, dc: DoubleConv, emb=None, first_is_stem=False, out_f16=False):
        """`out_f16`: the caller guarantees that the result's only consumer is the first conv of the next DoubleConv
        (Down/UpBlock's residual DoubleConv, bot1, bot2): it is then written as an fp16 operand.  An fp16 input `x`
        (such a result) makes conv1 an fp16 x fp16 GEMM."""
        c1, g1, _, c2, g2 = dc.double_conv
        mid, cout = c1.out_channels, c2.out_channels
        lo = None
        x_f16 = x.dtype == torch.float16
        assert not (x_f16 and dc.residual), "the residual add reads bf16"
        if first_is_stem:      # Cin = 3: fp32 NCHW input, direct kernel
            w = self.f32(c1.weight)
            wp = self.buf(mid, 9, cin, dtype=torch.float32)
            capi.call("advs_pack_conv_weight", w.data_ptr(), wp.data_ptr(), mid, cin, 3, 3, capi.F32, _st())
            h = self.buf(B, H, W, mid)
            self.add("advs_conv3x3_stem", x.data_ptr(), wp.data_ptr(), None, h.data_ptr(), B, H, W, cin, mid, self.dt)
        else:
            h, lo = self.conv(x, B, H, W, cin, self.pack(c1.weight, f16=x_f16), mid, want_lo=True)
        # GroupNorm + activation is bounded and has one consumer (conv2): an fp16 operand on the tcgen05 path
        f16 = self.on_sm100(mid, cout)
        h = self.gn(h, B, H * W, mid, g1, act=self.act, x_lo=lo, out_f16=f16)
        h, lo = self.conv(h, B, H, W, mid, self.pack(c2.weight, f16=f16), cout, want_lo=True)
        if dc.residual:        # act(x + double_conv(x))  (conv.py:49-63)
            return self.gn(h, B, H * W, cout, g2, act=self.act, residual=x, x_lo=lo, out_f16=out_f16)
        return self.gn(h, B, H * W, cout, g2, act=0, emb=emb, x_lo=lo, out_f16=out_f16)

    def emb_proj(self, block):
        lin = block.emb_layer[1]
        out = self.buf(self.B, lin.out_features, dtype=torch.float32)
        w, b = self.f32(lin.weight), self.f32(lin.bias)
        self.add("advs_linear_f32", self.time.data_ptr(), w.data_ptr(), b.data_ptr(), out.data_ptr(), self.B, lin.in_features,
                 lin.out_features, 1, 0)
        return out

    def down(self, x, H, W, cin, blk: DownBlock):
        B = self.B
        p = self.buf(B, H // 2, W // 2, cin)
        self.add("advs_maxpool2x2", x.data_ptr(), p.data_ptr(), B, H, W, cin, self.dt)
        emb = self.emb_proj(blk)
        nxt = blk.maxpool_conv[2].double_conv[0]
        h = self.double_conv(p, B, H // 2, W // 2, cin, blk.maxpool_conv[1], out_f16=self.on_sm100(cin, nxt.out_channels))
        return self.double_conv(h, B, H // 2, W // 2, cin, blk.maxpool_conv[2], emb=emb)

    def up(self, x, skip, H, W, cx, cs, blk: UpBlock):
        """x [B,H,W,cx] is upsampled to 2H x 2W and concatenated AFTER skip [B,2H,2W,cs] (block.py:72-73)."""
        B = self.B
        ct = cs + cx
        cat = self.buf(B, 2 * H, 2 * W, ct)
        self.add("advs_copy_channels", skip.data_ptr(), cat.data_ptr(), B * 4 * H * W, cs, ct, 0, self.dt)
        self.add("advs_upsample_bilinear2x", x.data_ptr(), cat.data_ptr(), B, H, W, cx, ct, cs, self.dt)
        emb = self.emb_proj(blk)
        nxt = blk.conv[1].double_conv[0]
        h = self.double_conv(cat, B, 2 * H, 2 * W, ct, blk.conv[0], out_f16=self.on_sm100(ct, nxt.out_channels))
        return self.double_conv(h, B, 2 * H, 2 * W, ct, blk.conv[1], emb=emb)

    def layernorm(self, x, rows, Cc, ln, out_f16):
        y = self.buf(*x.shape, dtype=torch.float16 if out_f16 else None)
        g, b = self.f32(ln.weight), self.f32(ln.bias)
        if out_f16:
            self.add("advs_layernorm_f16out", x.data_ptr(), g.data_ptr(), b.data_ptr(), y.data_ptr(), rows, Cc, float(ln.eps))
        else:
            self.add("advs_layernorm", x.data_ptr(), g.data_ptr(), b.data_ptr(), y.data_ptr(), rows, Cc, float(ln.eps), self.dt)
        return y

    def sa(self, x, H, W, Cc, m: SelfAttention):
        B, T, heads = self.B, H * W, 4
        dh = Cc // heads
        f16 = self.on_sm100(Cc, 3 * Cc, dh)        # LayerNorm outputs are bounded: fp16 operands for the token linears
        ln = self.layernorm(x, B * T, Cc, m.ln, f16)
        # in_proj rows are [q | k | v] blocks; the qkv epilogue wants per-head [q_h | k_h | v_h]
        idx = torch.arange(3 * Cc, device=self.dev).view(3, heads, dh).permute(1, 0, 2).reshape(-1)
        w_in = m.mha.in_proj_weight.detach().float()[idx].contiguous().view(3 * Cc, Cc, 1, 1)
        b_in = m.mha.in_proj_bias.detach().float()[idx].contiguous()
        self.keep.append(b_in)
        q, k, vt = self.conv(ln, B, H, W, Cc, self.pack(w_in, f16=f16), 3 * Cc, bias=b_in, qkv_heads=heads)
        o = self.buf(B, T, Cc)
        dhs = q.shape[3]                            # storage head dim (64 when dh was padded)
        if self.sm100 and dhs in (64, 128, 256) and T % 8 == 0:
            pb = capi.PlanBuffer(capi.ATTN_PLAN_BYTES)
            capi.call("advs_attention_sm100_plan_ex", q.data_ptr(), k.data_ptr(), vt.data_ptr(), o.data_ptr(), B, heads, T, dhs,
                      dh, pb.ptr)
            self.keep.append(pb)
            self.add("advs_attention_sm100_launch", pb.ptr)
            self.attn_kinds.append("sm100")
        else:
            wsb = int(self.lib.advs_attention_simt_workspace_bytes(B, heads, T))
            ws = self.buf(wsb, dtype=torch.uint8)
            self.add("advs_attention_simt", q.data_ptr(), k.data_ptr(), vt.data_ptr(), o.data_ptr(), B, heads, T, dh,
                     ws.data_ptr(), wsb, self.dt)
            self.attn_kinds.append("simt")
        wo = m.mha.out_proj.weight.detach().float().view(Cc, Cc, 1, 1)
        av = self.conv(o, B, H, W, Cc, self.pack(wo), Cc, bias=self.f32(m.mha.out_proj.bias), residual=x)
        f16 = self.on_sm100(Cc, Cc)
        f = self.layernorm(av, B * T, Cc, m.ff_self[0], f16)
        h = self.conv(f, B, H, W, Cc, self.pack(m.ff_self[1].weight.detach().float().view(Cc, Cc, 1, 1), f16=f16), Cc,
                      bias=self.f32(m.ff_self[1].bias))
        ha = self.buf(B, H, W, Cc)
        self.add("advs_activation", h.data_ptr(), ha.data_ptr(), B * T * Cc, self.act, self.dt)
        return self.conv(ha, B, H, W, Cc, self.pack(m.ff_self[3].weight.detach().float().view(Cc, Cc, 1, 1)), Cc,
                         bias=self.f32(m.ff_self[3].bias), residual=av)

    def _build(self):
        n, B, S = self.net, self.B, self.net.image_size
        ch = n.channel
        x1 = self.double_conv(self.x, B, S, S, n.in_channel, n.inc, first_is_stem=True)
        x2 = self.sa(self.down(x1, S, S, ch[1], n.down1), S // 2, S // 2, ch[2], n.sa1)
        x3 = self.sa(self.down(x2, S // 2, S // 2, ch[2], n.down2), S // 4, S // 4, ch[3], n.sa2)
        x4 = self.sa(self.down(x3, S // 4, S // 4, ch[3], n.down3), S // 8, S // 8, ch[3], n.sa3)
        b = self.double_conv(x4, B, S // 8, S // 8, ch[3], n.bot1, out_f16=self.on_sm100(ch[4], ch[4]))
        b = self.double_conv(b, B, S // 8, S // 8, ch[4], n.bot2, out_f16=self.on_sm100(ch[4], ch[3]))
        b = self.double_conv(b, B, S // 8, S // 8, ch[4], n.bot3)
        u = self.sa(self.up(b, x3, S // 8, S // 8, ch[3], ch[3], n.up1), S // 4, S // 4, ch[2], n.sa4)
        u = self.sa(self.up(u, x2, S // 4, S // 4, ch[2], ch[2], n.up2), S // 2, S // 2, ch[1], n.sa5)
        u = self.sa(self.up(u, x1, S // 2, S // 2, ch[1], ch[1], n.up3), S, S, ch[1], n.sa6)
        # outc: 1x1 conv -> fp32 NCHW (zero-padded to a 64-row weight tile on the tcgen05 path)
        w = n.outc.weight.detach().float().contiguous()
        pad = 64 if (self.sm100 and ch[1] % 64 == 0) else n.out_channel
        wp = torch.zeros(pad, 1, ch[1], dtype=self.tdt, device=self.dev)
        wp[:n.out_channel] = self.pack(w)
        bp = torch.zeros(pad, dtype=torch.float32, device=self.dev)
        bp[:n.out_channel] = n.outc.bias.detach().float()
        self.keep += [wp, bp]
        self.conv(u, B, S, S, ch[1], wp, pad, bias=bp, nchw_out=self.eps, cout_valid=n.out_channel)

    # ---- execution -----------------------------------------------------------------------------------------
    def run(self, use_labels, n_labeled=None):
        """`n_labeled`: only the first n_labeled rows get their label embedding (the rest run unconditionally)."""
        st = _st()
        capi.call("advs_pos_encoding_ex", self.t.data_ptr(), self.B, self.inv_freq.data_ptr(), self.net.time_channel // 2,
                  self.y.data_ptr() if use_labels else None, self.label_w.data_ptr() if use_labels else None,
                  self.B if n_labeled is None else n_labeled, self.time.data_ptr(), st)
        err = self.lib.advs_last_error
        for fn, args, name in self.L:
            rc = fn(*args, st)
            if rc:
                raise capi.AdvsError(f"{name} failed (rc={rc}): {err().decode()}")

    def forward(self, x, time, y=None):
        with torch.cuda.device(self.dev):
            self.x.copy_(x)
            self.t.copy_(time.reshape(-1).to(torch.int64))
            if y is not None:
                if self.label_w is None:
                    raise ValueError("labels given but the network was built without num_classes")
                self.y.copy_(y.reshape(-1).to(torch.int64))
            self.run(y is not None)
            return self.eps.clone()


# ---------------------------------------------------------------------------------------------------------------
class DDIMDiffusion:
    """model/samples/ddim.py `DDIMDiffusion` (+ BaseDiffusion's linear schedule, base.py:18-49)."""

    def __init__(self, noise_steps=1000, sample_steps=500, beta_start=1e-4, beta_end=2e-2, img_size=64, device="cpu"):
        self.noise_steps, self.beta_start, self.beta_end, self.img_size, self.device = (noise_steps, beta_start, beta_end,
                                                                                        img_size, device)
        self.beta = torch.linspace(start=beta_start, end=beta_end, steps=noise_steps).to(device)
        self.alpha = 1. - self.beta
        self.alpha_hat = torch.cumprod(input=self.alpha, dim=0)
        self.sample_steps, self.eta = sample_steps, 0
        ts = torch.arange(0, noise_steps, (noise_steps // sample_steps)).long() + 1
        ts = reversed(torch.cat((torch.tensor([0], dtype=torch.long), ts)))
        self.time_step = list(zip(ts[:-1], ts[1:]))

    def _coefficients(self):
        """Per step [sqrt(1-a_t), sqrt(a_t), sqrt(a_prev), sqrt((1-a_prev) - c1^2), c1, 0,0,0] with the reference's fp32
        op order (ddim.py:70-94); eta = 0 => c1 = 0."""
        ah = self.alpha_hat.detach().float().cpu()
        rows = []
        for i, p_i in self.time_step:
            a_t, a_p = ah[int(i)], ah[int(p_i)]
            c1 = self.eta * torch.sqrt((1 - a_t / a_p) * (1 - a_p) / (1 - a_t))
            c2 = torch.sqrt((1 - a_p) - c1 ** 2)
            z = torch.zeros(())
            rows.append(torch.stack([torch.sqrt(1 - a_t), torch.sqrt(a_t), torch.sqrt(a_p), c2, c1.float(), z, z, z]))
        return torch.stack(rows).float().contiguous()

    def sample(self, model, n, labels=None, cfg_scale=None, *, x_T=None, return_float=False):
        """DDIM sampling with optional classifier-free guidance; returns uint8 [n,3,S,S] on the device like the
        reference (ddim.py:48-100).  `x_T` / `return_float` are extension kwargs for parity tests."""
        if not isinstance(model, UNet):
            raise TypeError("advshadow_b200.iddm.DDIMDiffusion.sample needs an advshadow_b200.iddm.UNet")
        model.eval()
        dev = next(model.parameters()).device
        S = self.img_size
        # ddim.py:75-84: unconditional if neither labels nor a guidance scale is given; otherwise the (possibly
        # label-free) prediction is mixed with the unconditional one when cfg_scale > 0 (None > 0 raises there too)
        use_labels = labels is not None
        guided_run = False
        if not (labels is None and cfg_scale is None):
            if cfg_scale is None:
                raise TypeError("'>' not supported between instances of 'NoneType' and 'int' (cfg_scale is required with labels)")
            guided_run = use_labels and cfg_scale > 0      # without labels both predictions coincide: lerp(a, a, w) = a
        # guided: rows [0, n) carry the labels, rows [n, 2n) are the same images without -- one forward gives both
        eng = model.engine(2 * n if guided_run else n)
        with torch.no_grad(), torch.cuda.device(dev):
            # the sampler state IS the engine's input buffer; every per-step quantity (timestep, coefficients, step
            # counter) is read from device memory, so one step is a fixed launch sequence: captured once per
            # (engine, guidance mode) in a CUDA graph and replayed -- the 64x64 network is launch-latency bound
            x = eng.x[:n]
            x.copy_((torch.randn((n, 3, S, S)) if x_T is None else x_T.float()).to(dev))
            if guided_run:
                eng.x[n:].copy_(x)
            if use_labels:
                if eng.label_w is None:
                    raise ValueError("labels given but the network was built without num_classes")
                eng.y[:n].copy_(labels.reshape(-1).to(device=dev, dtype=torch.int64))
            n_steps = len(self.time_step)
            cache = eng.__dict__.setdefault("_step_graphs", {})
            gkey = (use_labels, guided_run)
            st8 = cache.get(gkey)
            if st8 is None or st8["cap"] < n_steps:
                cap = max(n_steps, 64)
                st8 = dict(cap=cap, coef=torch.zeros(cap, 8, dtype=torch.float32, device=dev),
                           t=torch.zeros(cap, dtype=torch.int64, device=dev), step=torch.zeros(1, dtype=torch.int32, device=dev),
                           guided=torch.empty_like(x), graph=None)
                cache[gkey] = st8
            st8["coef"][:n_steps].copy_(self._coefficients())
            st8["t"][:n_steps].copy_(torch.tensor([int(i) for i, _ in self.time_step], dtype=torch.int64))
            coef, step, guided = st8["coef"], st8["step"], st8["guided"]
            t_rows = st8["t"].view(torch.float32)          # [cap, 2] floats: the int64 timesteps, copied bit for bit
            step.zero_()

            def one_step():
                st = _st()
                capi.call("advs_select_row", t_rows.data_ptr(), 2, step.data_ptr(), eng.t.data_ptr(), eng.B, st)
                eng.run(use_labels, n_labeled=n)
                eps = eng.eps
                if guided_run:      # lerp(uncond, cond, w) of the two halves (ddim.py:89)
                    capi.call("advs_cfg_lerp", eng.eps[n:].data_ptr(), eng.eps[:n].data_ptr(), C.c_float(float(cfg_scale)),
                              guided.data_ptr(), x.numel(), st)
                    eps = guided
                capi.call("advs_ddim_step", x.data_ptr(), eps.data_ptr(), None, x.data_ptr(), x.numel(), coef.data_ptr(),
                          step.data_ptr(), 1, 1, st)
                if guided_run:
                    eng.x[n:].copy_(x)

            graph = None
            if _USE_GRAPH and n_steps > 2:
                # the guidance scale is baked into the captured launch: one graph per scale
                if st8["graph"] is None or st8.get("graph_scale") != (float(cfg_scale) if guided_run else None):
                    x0 = x.clone()
                    s = torch.cuda.Stream(device=dev)
                    s.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(s):
                        one_step()                 # eager once: lazy one-time kernel attribute setup
                    torch.cuda.current_stream().wait_stream(s)
                    x.copy_(x0)
                    if guided_run:
                        eng.x[n:].copy_(x0)
                    step.zero_()
                    g = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(g):
                        one_step()
                    st8["graph"], st8["graph_scale"] = g, (float(cfg_scale) if guided_run else None)
                    x.copy_(x0)
                    if guided_run:
                        eng.x[n:].copy_(x0)
                    step.zero_()
                graph = st8["graph"]
            for _ in tqdm(range(n_steps), disable=None):
                if graph is not None:
                    graph.replay()
                else:
                    one_step()
            x = x.clone()
            model.train()          # the reference leaves the model in train() mode (ddim.py:95)
            if return_float:
                return x
            out = torch.empty(x.shape, dtype=torch.uint8, device=dev)
            capi.call("advs_to_uint8", x.data_ptr(), out.data_ptr(), x.numel(), _st())
            return out


# ---------------------------------------------------------------------------------------------------------------
# The reference's on-disk checkpoint (utils/checkpoint.py:21-157): a dict
#   {start_epoch, model, ema_model, optimizer, num_classes, classes_name, conditional, image_size, sample, network, act}
# written as <results_dir>/ckpt_last.pt (+ an optional per-epoch copy).  Host-side file I/O, no kernels.
# ---------------------------------------------------------------------------------------------------------------
CKPT_KEYS = ("start_epoch", "model", "ema_model", "optimizer", "num_classes", "classes_name", "conditional", "image_size",
             "sample", "network", "act")


def save_ckpt(epoch, save_name, ckpt_model, ckpt_ema_model, ckpt_optimizer, results_dir, save_model_interval,
              start_model_interval, num_classes=None, conditional=None, image_size=None, sample=None, network=None,
              act=None, classes_name=None, **kwargs):
    """utils/checkpoint.py:118-156, same arguments and the same files: ckpt_last.pt, plus `<save_name>.pt` when
    save_model_interval is set and epoch > start_model_interval."""
    import shutil
    state = {"start_epoch": epoch, "model": ckpt_model, "ema_model": ckpt_ema_model, "optimizer": ckpt_optimizer,
             "num_classes": num_classes if conditional else 1, "classes_name": classes_name, "conditional": conditional,
             "image_size": image_size, "sample": sample, "network": network, "act": act}
    last = os.path.join(results_dir, "ckpt_last.pt")
    torch.save(obj=state, f=last)
    if save_model_interval and epoch > start_model_interval:
        shutil.copyfile(last, os.path.join(results_dir, f"{save_name}.pt"))


def load_model_ckpt(model, model_ckpt, is_train=True, is_pretrain=False, is_distributed=False):
    """utils/checkpoint.py:71-115: strip (inference / single-process pretrain) or add (distributed pretrain) the
    DistributedDataParallel `module.` prefix, drop the label embedding of a pretrain checkpoint (its class count
    differs), keep only tensors whose shape matches the model, load the rest."""
    from collections import OrderedDict
    import numpy as np
    model_dict = model.state_dict()
    weights = dict(model_ckpt)
    if not is_train or (is_train and is_pretrain and not is_distributed):
        weights = {(k[len("module."):] if k.startswith("module.") else k): v for k, v in weights.items()}
    if is_train and is_pretrain:
        if is_distributed:
            weights = {(k if k.startswith("module.") else "module." + k): v for k, v in weights.items()}
            weights["module.label_emb.weight"] = None
        else:
            weights["label_emb.weight"] = None
    weights = {k: v for k, v in weights.items() if np.shape(model_dict[k]) == np.shape(v)}
    model_dict.update(weights)
    model.load_state_dict(state_dict=OrderedDict(model_dict))


def load_ckpt(ckpt_path, model, device, optimizer=None, is_train=True, is_pretrain=False, is_distributed=False,
              is_use_ema=False):
    """utils/checkpoint.py:21-68: 'model' is the default source, 'ema_model' when it is the only one present or when
    is_use_ema; in resumed training also the optimiser state, returning the next epoch."""
    state = torch.load(f=ckpt_path, map_location=device, weights_only=False)
    assert state["model"] is not None or state["ema_model"] is not None, \
        "Error!! Checkpoint model and ema_model are not None. Please check checkpoint's structure."
    if state["model"] is None:
        ckpt_model = state["ema_model"]
    else:
        ckpt_model = state["ema_model"] if is_use_ema else state["model"]
    load_model_ckpt(model=model, model_ckpt=ckpt_model, is_train=is_train, is_pretrain=is_pretrain,
                    is_distributed=is_distributed)
    if is_train and not is_pretrain:
        optimizer.load_state_dict(state_dict=state["optimizer"])
        return state["start_epoch"] + 1


def unet_from_ckpt(ckpt_path, device="cuda", is_use_ema=False, in_channel=3, out_channel=3):
    """tools/generate.py:57-73 in one call: build the UNet the checkpoint describes (num_classes when conditional,
    image_size, act) and load its weights for sampling."""
    state = torch.load(f=ckpt_path, map_location="cpu", weights_only=False)
    net = UNet(in_channel=in_channel, out_channel=out_channel,
               num_classes=state["num_classes"] if state.get("conditional") else None,
               image_size=state["image_size"], device=device, act=state.get("act") or "silu").to(device)
    load_ckpt(ckpt_path, net, device, is_train=False, is_use_ema=is_use_ema)
    return net.eval(), state
