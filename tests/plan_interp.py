"""TEST INFRASTRUCTURE: replays a `plan.Plan` with plain PyTorch ops on the CPU.

Purpose: prove on a GPU-less box that the op list / buffer wiring / fused-epilogue bookkeeping the
CUDA engine executes is the reference's UNet (diff_model.py:245-267).  `round_bf16=True` rounds every
activation buffer to bf16 at the points the bf16 engine does, which gives a CPU estimate of the
bf16-mode error.  Never imported by the product.
"""
import math

import torch
import torch.nn.functional as F


def _nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def _nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def temb_table(plan, params, timesteps):
    mc = plan.spec.model_channels
    half = mc // 2
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    args = timesteps[:, None].float() * freqs[None]
    e = torch.cat([torch.cos(args), torch.sin(args)], -1)
    e = F.linear(e, params["time_embed.0.weight"], params["time_embed.0.bias"])
    e = F.linear(F.silu(e), params["time_embed.2.weight"], params["time_embed.2.bias"])
    cols = [F.linear(F.silu(e), params[s.weight + ".weight"], params[s.weight + ".bias"]) for s in plan.temb_slots]
    return torch.cat(cols, 1)


def _sm100_ok(plan, a):
    """engine._conv_sm100_ok restated: which convs run on the tcgen05 path (and may take fp16 operands)"""
    if a["cout"] % 64 or any(plan.shape(src)[3] % 64 for (src, _, _, _) in a["segs"]):
        return False
    return not (a["qkv"] is not None and (a["cout"] // (3 * a["heads"])) % 32)


def run_plan(plan, params, x, timesteps, round_bf16=False, wide_prenorm=2, gemm_operands="fp16", mixed_mma=False,
             exact_w=(), fp16_levels=2):
    """`wide_prenorm` (with round_bf16): conv outputs of that many top-resolution levels keep an unrounded copy that
    only GroupNorm reads (the engine's bf16 + int8 mantissa-extension storage, exact to 2^-15 relative).
    `gemm_operands` = "fp16": GroupNorm outputs consumed by a tcgen05 conv, and the weights multiplying them, are
    rounded to fp16 instead of bf16 (engine.UNetEngine(gemm_operands=...)); `mixed_mma`: every other tcgen05
    conv's weights are fp16 as well (their activations stay bf16) -- the hardware rejects mixed operand formats
    (tools/gpu/probe_mixed_mma.py), so the engine never does this; kept for the error study.  `exact_w`: weight
    categories left unrounded ("stem", "shortcut", "upconv", "down", "proj"), same purpose."""
    rnd = (lambda t: t.to(torch.bfloat16).float()) if round_bf16 else (lambda t: t)
    r16 = (lambda t: t.to(torch.float16).float()) if (round_bf16 and gemm_operands == "fp16") else rnd
    # `fp16_levels`: fp16 operands only on that many top-resolution levels (None = all), engine.UNetEngine(fp16_levels=)
    top = lambda buf: fp16_levels is None or (plan.shape(buf)[1] << fp16_levels) > plan.H
    gn_dst = {op.args["dst"] for op in plan.ops if op.kind == "gn" and top(op.args["dst"])}
    gn_f16 = set()
    for op in plan.ops:
        a = op.args
        if op.kind == "conv" and _sm100_ok(plan, a) and a["segs"][0][0] in gn_dst:
            gn_f16.add(a["segs"][0][0])
        elif op.kind == "head" and a["cin"] % 64 == 0 and a["src"] in gn_dst:
            gn_f16.add(a["src"])
    wide = {}

    def store(dst, t_nhwc):
        bufs[dst] = rnd(t_nhwc)
        if round_bf16 and (t_nhwc.shape[1] << wide_prenorm) > plan.H and t_nhwc.shape[3] % 32 == 0:
            wide[dst] = t_nhwc

    wr = (lambda t: t.to(torch.bfloat16).float()) if round_bf16 else (lambda t: t)   # weight rounding
    table = temb_table(plan, params, timesteps)
    if table.shape[0] == 1:
        table = table.expand(x.shape[0], -1)
    bufs = {}
    eps = None
    for op in plan.ops:
        a = op.args
        if op.kind == "stem":
            # (the engine feeds x as a bf16 hi/lo pair: exact to 2^-17; only the weights are rounded)
            # fp16 mode: fp16 hi/lo rows of x (exact to 2^-22) x fp16 weights
            wq = params[a["weight"] + ".weight"] if "stem" in exact_w else r16(params[a["weight"] + ".weight"])
            y = F.conv2d(x, wq, params[a["weight"] + ".bias"], padding=1)
            store(a["dst"], _nhwc(y))
        elif op.kind == "gn":
            xin = torch.cat([wide.get(s, bufs[s]) for s in a["srcs"]], 3)
            y = F.group_norm(_nchw(xin), a["groups"], params[a["weight"] + ".weight"], params[a["weight"] + ".bias"], 1e-5)
            if a["silu"]:
                y = F.silu(y)
            bufs[a["dst"]] = (r16 if a["dst"] in gn_f16 else rnd)(_nhwc(y))
        elif op.kind == "conv":
            acc = None
            for i, (src, wname, taps, sl) in enumerate(a["segs"]):
                w = params[wname + ".weight"]
                if sl is not None:
                    w = w[:, sl[0]:sl[1]]
                xi = _nchw(bufs[src])
                kind = ("shortcut" if i > 0 else "down" if a["stride"] == 2 else "proj" if (taps == 1 and a["residual"]) else "other")
                wq = (r16 if (src in gn_f16 or (mixed_mma and _sm100_ok(plan, a))) else wr)(w)
                if kind in exact_w:
                    wq = w
                if taps == 9:
                    y = F.conv2d(xi, wq, None, stride=a["stride"] if i == 0 else 1, padding=1)
                else:
                    y = F.conv2d(xi, wq, None)
                acc = y if acc is None else acc + y
            for bn in a["bias"]:
                acc = acc + params[bn + ".bias"][None, :, None, None]
            if a["temb"] is not None:
                acc = acc + table[:, a["temb"]:a["temb"] + a["cout"]][:, :, None, None]
            if a["residual"]:
                acc = acc + _nchw(bufs[a["residual"]])
            if a["qkv"] is not None:
                B, C3, H, W = acc.shape
                heads = a["heads"]
                dh = C3 // (3 * heads)
                scale = 1.0 / math.sqrt(math.sqrt(dh))
                qkv = acc.reshape(B, heads, 3, dh, H * W)
                bufs[a["qkv"][0]] = rnd((qkv[:, :, 0] * scale).transpose(2, 3).contiguous())   # [B,h,T,dh]
                bufs[a["qkv"][1]] = rnd((qkv[:, :, 1] * scale).transpose(2, 3).contiguous())
                bufs[a["qkv"][2]] = rnd(qkv[:, :, 2].contiguous())                              # [B,h,dh,T]
            else:
                store(a["dst"], _nhwc(acc))
        elif op.kind == "attn":
            q, k, vt = bufs[a["q"]], bufs[a["k"]], bufs[a["vt"]]
            s = torch.einsum("bhtd,bhsd->bhts", q, k).softmax(-1)
            o = torch.einsum("bhts,bhds->bthd", s, vt)                                          # [B,T,h,dh]
            B, T = o.shape[0], o.shape[1]
            H, W = plan.bufs[a["dst"]].shape[1:3]
            bufs[a["dst"]] = rnd(o.reshape(B, H, W, -1))
        elif op.kind == "up":
            y = F.interpolate(_nchw(bufs[a["src"]]), scale_factor=2, mode="nearest")
            bufs[a["dst"]] = _nhwc(y)
        elif op.kind == "upconv":
            y = F.interpolate(_nchw(bufs[a["src"]]), scale_factor=2, mode="nearest")
            wq = params[a["weight"] + ".weight"] if "upconv" in exact_w else (r16 if mixed_mma else wr)(params[a["weight"] + ".weight"])
            y = F.conv2d(y, wq, params[a["weight"] + ".bias"], padding=1)
            store(a["dst"], _nhwc(y))
        elif op.kind == "head":
            wq = (r16 if (a["src"] in gn_f16 or mixed_mma) else wr)(params[a["weight"] + ".weight"])
            eps = F.conv2d(_nchw(bufs[a["src"]]), wq, params[a["weight"] + ".bias"], padding=1)
        else:
            raise AssertionError(op.kind)
    return eps
