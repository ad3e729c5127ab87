"""Tensor-level wrappers of single C-ABI entry points (used by the drop-in modules and the tests).
Every function takes CUDA tensors and launches on torch's current stream; CPU tensors raise."""
import ctypes as C
import math

import torch

from . import _capi as capi


def _st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("advshadow_b200 ops need CUDA tensors (there is no CPU path)")


def _dt(t):
    if t.dtype == torch.float32:
        return capi.F32
    if t.dtype == torch.bfloat16:
        return capi.BF16
    if t.dtype == torch.float16:
        return capi.F16
    raise TypeError(f"unsupported dtype {t.dtype}")


def _p(t):
    return t.data_ptr() if t is not None else None


def timestep_embedding(timesteps, dim, max_period=10000):
    _need_cuda(timesteps)
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(start=0, end=half, dtype=torch.float32) / half)
    freqs = freqs.to(timesteps.device)
    t = timesteps.to(torch.int64).contiguous()
    out = torch.zeros(t.numel(), dim, dtype=torch.float32, device=t.device)
    with torch.cuda.device(t.device):
        tmp = torch.empty(t.numel(), 2 * half, dtype=torch.float32, device=t.device)
        capi.call("advs_timestep_embedding", _p(t), t.numel(), _p(freqs), half, _p(tmp), _st())
    out[:, :2 * half] = tmp
    return out


def pack_conv_weight(w, dtype):
    _need_cuda(w)
    O, I, kh, kw = w.shape
    dst = torch.empty(O, kh * kw, I, dtype=dtype, device=w.device)
    w = w.float().contiguous()
    with torch.cuda.device(w.device):
        capi.call("advs_pack_conv_weight", _p(w), _p(dst), O, I, kh, kw, _dt(dst), _st())
    return dst


def groupnorm(x0, x1, gamma, beta, groups=32, eps=1e-5, silu=False):
    """NHWC GroupNorm(+SiLU) over cat([x0, x1], channel)."""
    _need_cuda(x0, x1, gamma, beta)
    B, H, W, c0 = x0.shape
    c1 = x1.shape[3] if x1 is not None else 0
    y = torch.empty(B, H, W, c0 + c1, dtype=x0.dtype, device=x0.device)
    ss = torch.empty(B, c0 + c1, 2, dtype=torch.float32, device=x0.device)
    lib = capi.lib()
    wsb = int(lib.advs_groupnorm_workspace_bytes(B, H * W, c0 + c1))
    ws = torch.empty(max(wsb, 4), dtype=torch.uint8, device=x0.device)
    with torch.cuda.device(x0.device):
        capi.call("advs_groupnorm_stats", _p(x0), c0, _p(x1), c1, B, H * W, groups, eps, _p(gamma), _p(beta), _p(ss),
                  _p(ws), wsb, _dt(x0), _st())
        capi.call("advs_groupnorm_apply", _p(x0), c0, _p(x1), c1, B, H * W, _p(ss), 1 if silu else 0, _p(y), _dt(x0),
                  _st())
    return y


def conv(segs, B, H, W, cout, stride=1, bias=None, temb=None, residual=None, qkv_heads=None, impl="simt", qkv_dh_pad=0):
    """segs: list of (x NHWC, packed weight [Cout][taps][C]).  Returns y NHWC, or (q, k, vt) (zero-padded to
    `qkv_dh_pad` columns / rows per head when that is larger than the head dim)."""
    x0 = segs[0][0]
    _need_cuda(x0)
    dev, dtype = x0.device, x0.dtype
    cp = capi.ConvParams()
    cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, cout, stride, len(segs)
    for i, (x, w) in enumerate(segs):
        cp.seg[i].x, cp.seg[i].w, cp.seg[i].C, cp.seg[i].taps = _p(x), _p(w), x.shape[3], w.shape[1]
    cp.bias = _p(bias)
    if temb is not None:
        cp.temb = _p(temb)
        cp.temb_stride = temb.shape[1] if temb.shape[0] > 1 else 0
    cp.residual = _p(residual)
    cp.dtype = _dt(x0)
    if qkv_heads:
        dh = cout // (3 * qkv_heads)
        T = H * W
        dhs = max(dh, qkv_dh_pad)
        q = torch.zeros(B, qkv_heads, T, dhs, dtype=dtype, device=dev)
        k = torch.zeros(B, qkv_heads, T, dhs, dtype=dtype, device=dev)
        vt = torch.zeros(B, qkv_heads, dhs, T, dtype=dtype, device=dev)
        cp.out_mode, cp.q, cp.k, cp.vt, cp.heads, cp.qkv_dh_pad = 1, _p(q), _p(k), _p(vt), qkv_heads, qkv_dh_pad
        cp.qk_scale = 1.0 / math.sqrt(math.sqrt(dh))
        ret = (q, k, vt)
    else:
        y = torch.empty(B, H, W, cout, dtype=dtype, device=dev)
        cp.out_mode, cp.y = 0, _p(y)
        ret = y
    with torch.cuda.device(dev):
        if impl == "simt":
            capi.call("advs_conv_simt", C.byref(cp), _st())
        elif impl == "sm100":
            pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
            capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
            capi.call("advs_conv_sm100_launch", pb.ptr, _st())
        else:
            raise ValueError(impl)
    return ret


def attention(q, k, vt, impl="simt", dh_valid=None):
    """`dh_valid` (sm100 only): q / k / vt are zero-padded from that head dim to dh = 64; the output is packed."""
    _need_cuda(q, k, vt)
    B, heads, T, dh = q.shape
    o = torch.empty(B, T, heads * (dh_valid or dh), dtype=q.dtype, device=q.device)
    with torch.cuda.device(q.device):
        if impl == "simt":
            wsb = int(capi.lib().advs_attention_simt_workspace_bytes(B, heads, T))
            ws = torch.empty(wsb, dtype=torch.uint8, device=q.device)
            capi.call("advs_attention_simt", _p(q), _p(k), _p(vt), _p(o), B, heads, T, dh, _p(ws), wsb, _dt(q), _st())
        elif impl == "sm100":
            pb = capi.PlanBuffer(capi.ATTN_PLAN_BYTES)
            capi.call("advs_attention_sm100_plan_ex", _p(q), _p(k), _p(vt), _p(o), B, heads, T, dh, dh_valid or dh, pb.ptr)
            capi.call("advs_attention_sm100_launch", pb.ptr, _st())
        else:
            raise ValueError(impl)
    return o


def upsample_nearest2x(x):
    _need_cuda(x)
    B, H, W, Cc = x.shape
    y = torch.empty(B, 2 * H, 2 * W, Cc, dtype=x.dtype, device=x.device)
    with torch.cuda.device(x.device):
        capi.call("advs_upsample_nearest2x", _p(x), _p(y), B, H, W, Cc, _dt(x), _st())
    return y


def success_flags(logits, labels):
    """flags[b] = argmax(logits[b]) != labels[b]; returns (flags uint8 [B], counts int64 [2])."""
    _need_cuda(logits, labels)
    logits = logits.float().contiguous()
    labels = labels.to(torch.int64).contiguous()
    B, ncls = logits.shape
    flags = torch.empty(B, dtype=torch.uint8, device=logits.device)
    counts = torch.zeros(2, dtype=torch.int64, device=logits.device)
    with torch.cuda.device(logits.device):
        capi.call("advs_success_flags", _p(logits), _p(labels), B, ncls, _p(flags), _p(counts), _st())
    return flags, counts
