"""Binds a `plan.Plan` to device memory and the C ABI: packs weights, carves the activation arena,
builds the TMA launch plans and replays the op list on torch's current CUDA stream.

torch is used here for plumbing only (device allocation, streams, parameter storage); every
arithmetic op on the hot path is one of the library's own kernels.  There is no fallback: a
non-CUDA device or a missing library raises.
"""
import ctypes as C
import math
from typing import Dict, List

import torch

from . import _capi as capi
from .plan import Plan, UNetSpec, build_unet_plan

_DT = {"fp32": capi.F32, "bf16": capi.BF16}
_TORCH_DT = {"fp32": torch.float32, "bf16": torch.bfloat16}
_CAPI_DT = {torch.float32: capi.F32, torch.bfloat16: capi.BF16, torch.float16: capi.F16}


def _stream_ptr():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


class UNetEngine:
    """One (spec, batch, height, width, precision) instance of the denoiser on one GPU."""

    def __init__(self, spec: UNetSpec, params: Dict[str, torch.Tensor], B: int, H: int, W: int,
                 precision: str = "bf16", conv_impl: str = "auto", attn_impl: str = "auto", fuse_gn_stats: bool = True,
                 fuse_upsample: bool = True, wide_prenorm: int = 2, gemm_operands: str = "fp16", fp16_levels: int = 2):
        """`wide_prenorm`: bf16 mode only -- tensors of the `wide_prenorm` highest-resolution levels that a GroupNorm
        reads (residual stream, conv1 outputs, skips) are stored as bf16 + an int8 mantissa extension
        (advs_conv_params.y_lo), so the value entering the normalisation carries 16 mantissa bits like the
        reference's fp32 tensor (dm1:71-72, 83-84) and the GEMM operand is rounded once instead of twice.  These
        levels have the fewest channels per dot product, hence the least averaging of rounding noise: the error
        study in DESIGN.md attributes 90 % of the bf16-mode output variance to them.  0 disables.
        `gemm_operands`: "fp16" (default) | "bf16" -- the 16-bit format of the GEMM operands that are bounded by
        construction: GroupNorm outputs (|y| <= sqrt(group size)*|gamma| + |beta|) and the weights multiplying them.
        fp16 gives them 11 mantissa bits instead of 8 at the same tcgen05 rate (kind::f16 takes either format);
        the stem's im2col rows of the (bounded) sampler state are fp16 too.  Activations whose range is not bounded
        -- conv outputs, the residual stream, q/k/v, attention output -- stay bf16, and so do the weights multiplying
        them: the hardware rejects an MMA whose two operands differ in format (tools/gpu/probe_mixed_mma.py:
        illegal instruction).  "bf16" makes every operand bf16.
        `fp16_levels`: fp16 operands are used on that many top-resolution levels only.  The deeper levels hold most
        of the FLOPs and contribute almost nothing to the output error (their dot products are 9 x 512 ... 9 x 2048
        terms long and their errors are averaged again on the way up), while fp16 MMAs draw measurably more power
        than bf16 ones: under the 1 kW cap the all-fp16 variant clocked 3 % lower (profiles/ab_r02.md)."""
        if precision not in _DT:
            raise ValueError(f"precision must be 'bf16' or 'fp32', got {precision!r}")
        any_p = next(iter(params.values()))
        if not any_p.is_cuda:
            raise RuntimeError("advshadow_b200: the denoiser runs on CUDA only (no CPU path); move the model to a GPU")
        self.lib = capi.lib()
        self.device = any_p.device
        self.spec, self.B, self.H, self.W = spec, B, H, W
        self.precision = precision
        self.dt = _DT[precision]
        self.act_bytes = 2 if precision == "bf16" else 4
        with torch.cuda.device(self.device):
            is_sm100 = bool(self.lib.advs_device_is_sm100())
        if conv_impl == "auto":
            conv_impl = "sm100" if (precision == "bf16" and is_sm100) else "simt"
        if attn_impl == "auto":
            attn_impl = "sm100" if (precision == "bf16" and is_sm100) else "simt"
        if (conv_impl == "sm100" or attn_impl == "sm100") and precision != "bf16":
            raise ValueError("the tcgen05 kernels are bf16-only; use precision='bf16'")
        self.conv_impl, self.attn_impl = conv_impl, attn_impl
        if gemm_operands not in ("fp16", "bf16"):
            raise ValueError("gemm_operands must be 'fp16' or 'bf16'")
        self.gemm_operands = gemm_operands if (precision == "bf16" and conv_impl == "sm100") else "bf16"

        # Upsample + conv3x3 as four low-res 2x2 phase convs (tcgen05 path only; fp32 mode keeps the
        # reference's op order: nearest upsample, then the 3x3 conv)
        fuse_up = conv_impl == "sm100" and spec.model_channels % 64 == 0 and fuse_upsample
        # only the attention blocks the tcgen05 kernel cannot take get the [B*heads, T, T] SIMT score buffer
        plan = build_unet_plan(spec, B, H, W, fuse_upsample=fuse_up,
                               attn_scores_ws=lambda T, dh: not self._attn_sm100_ok(dict(T=T, dh=dh)))
        self.plan: Plan = plan
        for b in plan.bufs.values():
            if b.shape and b.shape[0] == "gn_ws":
                _, bb, hw, c = b.shape
                b.elems = int(self.lib.advs_groupnorm_workspace_bytes(bb, hw, c)) // 4
        # GroupNorm statistics for free: a tcgen05 conv can emit per-tile channel sums of the tensor it
        # writes; attach such a partial buffer (same lifetime as the tensor) to every conv output a
        # GroupNorm will read
        self._stat_buf = {}
        gn_inputs = {s for op in plan.ops if op.kind == "gn" for s in op.args["srcs"]}
        # one {sum, sumsq} pair per 4 channels instead of per channel wherever every GroupNorm that reads the
        # tensor has 4 | channels-per-group and 4 | the tensor's channel offset inside the concat
        quad_ok = {}
        for op in plan.ops:
            if op.kind != "gn":
                continue
            ctot, off = sum(plan.shape(s)[3] for s in op.args["srcs"]), 0
            for sname in op.args["srcs"]:
                ok = (ctot // op.args["groups"]) % 4 == 0 and off % 4 == 0 and plan.shape(sname)[3] % 4 == 0
                quad_ok[sname] = quad_ok.get(sname, True) and ok
                off += plan.shape(sname)[3]
        if fuse_gn_stats:
            for idx, op in enumerate(plan.ops):
                a = op.args
                if op.kind == "upconv":
                    parts = 4 * int(self.lib.advs_conv_sm100_stats_parts(B, a["H"], a["W"]))
                elif op.kind == "conv" and a["qkv"] is None and self._conv_sm100_ok(a):
                    parts = int(self.lib.advs_conv_sm100_stats_parts(B, a["H"], a["W"]))
                elif op.kind == "stem" and self._stem_sm100_ok(a):
                    parts = int(self.lib.advs_conv_sm100_stats_parts(B, a["H"], a["W"]))
                else:
                    continue
                if parts <= 0 or a["dst"] not in gn_inputs:
                    continue
                gran = 4 if quad_ok.get(a["dst"], False) else 1
                name = plan.new_buf("gnpart", (B, parts, a["cout"] // gran, 2), "f32")
                plan.bufs[name].first, plan.bufs[name].last = idx, plan.bufs[a["dst"]].last
                self._stat_buf[a["dst"]] = (name, parts, gran)
        # fp16 GEMM operands: a GroupNorm output is fp16 when its (single) consumer is a tcgen05 conv / the head
        self._gn_f16 = set()
        self.fp16_levels = fp16_levels
        if self.gemm_operands == "fp16":
            gn_dst = {op.args["dst"] for op in plan.ops
                      if op.kind == "gn" and (plan.shape(op.args["dst"])[1] << fp16_levels) > H}
            for op in plan.ops:
                a = op.args
                if op.kind == "conv" and self._conv_sm100_ok(a) and a["segs"][0][0] in gn_dst:
                    self._gn_f16.add(a["segs"][0][0])
                elif op.kind == "head" and self._head_sm100_ok(a) and a["src"] in gn_dst:
                    self._gn_f16.add(a["src"])
        # "wide" pre-norm storage: an int8 companion (same lifetime) for every conv output of the top levels
        # that a GroupNorm reads
        self._lo_buf = {}
        self.wide_prenorm = wide_prenorm if precision == "bf16" else 0
        if self.wide_prenorm > 0:
            for idx, op in enumerate(plan.ops):
                a = op.args
                if op.kind == "stem":
                    if not self._stem_sm100_ok(a):
                        continue
                elif op.kind == "upconv":
                    pass
                elif not (op.kind == "conv" and a["qkv"] is None):
                    continue
                dst = a["dst"]
                shp = plan.shape(dst)
                if dst not in gn_inputs or shp[3] % 32 or shp[1] << self.wide_prenorm <= H:
                    continue
                name = plan.new_buf("lo8", shp, "u8")
                plan.bufs[name].first, plan.bufs[name].last = idx, plan.bufs[dst].last
                self._lo_buf[dst] = name
        plan.assign_offsets(self.act_bytes)

        dev = self.device
        self.arena = torch.empty(max(plan.arena_bytes, 1024), dtype=torch.uint8, device=dev)
        self.x = torch.zeros(B, spec.in_channels, H, W, dtype=torch.float32, device=dev)
        self.eps = torch.zeros(B, spec.out_channels, H, W, dtype=torch.float32, device=dev)
        self.temb_cur = torch.zeros(B, plan.temb_total, dtype=torch.float32, device=dev)
        half = spec.model_channels // 2
        # the reference evaluates this table on the host in fp32 (dm1:25-28) and uploads it
        self.freqs = torch.exp(-math.log(10000) * torch.arange(start=0, end=half, dtype=torch.float32) / half).to(dev)
        self._keep: List[object] = []     # packed weights, biases, plan blobs
        self._launches = []
        self._weights_loaded = False
        self._packed: Dict[tuple, torch.Tensor] = {}
        self._bias: Dict[tuple, torch.Tensor] = {}
        self._build_static()
        self.load_weights(params)
        self._build_launches()

    # ---- helpers ----
    def _attn_sm100_ok(self, a):
        # (a partial last key block is masked in-kernel; T % 8: TMA row pitch of v^T)
        return self.attn_impl == "sm100" and a["T"] % 8 == 0 and a["dh"] in (64, 128, 256)

    def _conv_sm100_ok(self, a):
        if self.conv_impl != "sm100":
            return False
        if a["cout"] % 64:
            return False
        for (src, _, _, _) in a["segs"]:
            if self.plan.shape(src)[3] % 64:
                return False
        if a["qkv"] is not None and (a["cout"] // (3 * a["heads"])) % 32:
            return False
        return True

    def _stem_sm100_ok(self, a):
        # Cin = 3 stem as a 64-wide 1x1 conv on the tensor cores (im2col rows, csrc/elementwise.cu)
        return self.conv_impl == "sm100" and 9 * a["cin"] <= 64 and a["cout"] % 64 == 0

    def _head_sm100_ok(self, a):
        return self.conv_impl == "sm100" and a["cin"] % 64 == 0 and a["cout"] <= 64

    def _ptr(self, buf):
        return self.arena.data_ptr() + self.plan.bufs[buf].offset

    def _w_dtype(self, src, on_sm100=True):
        """storage format of the packed weights that multiply activation buffer `src` on the tcgen05 path"""
        if self.gemm_operands == "fp16" and on_sm100 and (src in self._gn_f16 or src == "stem_col"):
            return torch.float16
        return _TORCH_DT[self.precision]

    def _operand_bits(self, srcs, dtypes):
        """advs_conv_params.operand_f16 from the activation buffers / packed-weight dtypes of segment 0 and 1.."""
        bits = 0
        f16 = lambda s_: s_ in self._gn_f16 or (s_ == "stem_col" and self.gemm_operands == "fp16")
        if f16(srcs[0]):
            bits |= 1
        if any(f16(s_) for s_ in srcs[1:]):
            bits |= 2
        if dtypes[0] == torch.float16:
            bits |= 4
        if any(d == torch.float16 for d in dtypes[1:]):
            bits |= 8
        return bits

    def buffer_view(self, buf):
        """torch view of an arena buffer (debug / tests)."""
        b = self.plan.bufs[buf]
        dt = {"act": _TORCH_DT[self.precision], "f32": torch.float32, "u8": torch.int8}[b.kind]
        if buf in self._gn_f16:
            dt = torch.float16
        n = b.elems * {"act": self.act_bytes, "f32": 4, "u8": 1}[b.kind]
        return self.arena[b.offset:b.offset + n].view(dt).view(*[s for s in b.shape if not isinstance(s, str)])

    # ---- static allocations that depend only on shapes ----
    def _build_static(self):
        dev, tdt = self.device, _TORCH_DT[self.precision]
        spec = self.spec
        for op in self.plan.ops:
            a = op.args
            if op.kind == "stem":
                if self._stem_sm100_ok(a):
                    self._packed[(a["weight"], "stem64")] = torch.empty(a["cout"], 1, 64, dtype=self._w_dtype("stem_col"), device=dev)
                    self._stem_col = torch.empty(self.B, a["H"], a["W"], 64, dtype=tdt, device=dev)
                else:
                    self._packed[(a["weight"], None)] = torch.empty(a["cout"], 9, a["cin"], dtype=torch.float32, device=dev)
                self._bias[(a["weight"],)] = torch.empty(a["cout"], dtype=torch.float32, device=dev)
            elif op.kind == "head":
                # the head runs through the implicit-GEMM conv (fp32 NCHW epilogue); on the tcgen05 path
                # its 3 output channels are zero-padded to one 64-row weight tile
                pad = 64 if self._head_sm100_ok(a) else a["cout"]
                wdt = self._w_dtype(a["src"], self._head_sm100_ok(a))
                self._packed[(a["weight"], None)] = torch.zeros(pad, 9, a["cin"], dtype=wdt, device=dev)
                self._bias[(a["weight"],)] = torch.zeros(pad, dtype=torch.float32, device=dev)
            elif op.kind == "conv":
                for (src, wname, taps, sl) in a["segs"]:
                    c = self.plan.shape(src)[3]
                    wdt = self._w_dtype(src, self._conv_sm100_ok(a))
                    self._packed[(wname, sl)] = torch.empty(a["cout"], taps, c, dtype=wdt, device=dev)
                if a["bias"]:
                    self._bias[tuple(a["bias"])] = torch.empty(a["cout"], dtype=torch.float32, device=dev)
            elif op.kind == "upconv":
                self._packed[(a["weight"], "up4")] = torch.empty(4, a["cout"], 4, a["C"], dtype=self._w_dtype(a["src"]), device=dev)
                self._bias[(a["weight"],)] = torch.empty(a["cout"], dtype=torch.float32, device=dev)
        ted = spec.time_embed_dim
        self.temb_w = torch.empty(self.plan.temb_total, ted, dtype=torch.float32, device=dev)
        self.temb_b = torch.empty(self.plan.temb_total, dtype=torch.float32, device=dev)
        mc = spec.model_channels
        self.te0_w = torch.empty(ted, mc, dtype=torch.float32, device=dev)
        self.te0_b = torch.empty(ted, dtype=torch.float32, device=dev)
        self.te2_w = torch.empty(ted, ted, dtype=torch.float32, device=dev)
        self.te2_b = torch.empty(ted, dtype=torch.float32, device=dev)
        # GroupNorm affine parameters live in engine-owned storage too: their addresses are baked into the launch
        # list and into captured CUDA graphs, which therefore stay valid across load_weights()
        self._gn = {}
        for op in self.plan.ops:
            if op.kind == "gn" and op.args["weight"] not in self._gn:
                c = op.args["C"]
                self._gn[op.args["weight"]] = (torch.empty(c, dtype=torch.float32, device=dev),
                                               torch.empty(c, dtype=torch.float32, device=dev))
        self.weights_version = 0

    # ---- weights ----
    def load_weights(self, params: Dict[str, torch.Tensor]):
        """(Re)pack every parameter into engine-owned buffers in the kernels' layouts.  No kernel argument ever
        points at a parameter tensor, so launch plans and captured CUDA graphs remain valid after a reload -- even
        when the parameters' storage moved (`load_state_dict(assign=True)`, `p.data = ...`).  Temporaries made here
        are consumed on the current stream and released by torch's stream-ordered allocator."""
        st = _stream_ptr()
        src = {k: v.detach() for k, v in params.items()}

        def p32(name):
            t = src[name]
            if t.dtype != torch.float32 or not t.is_contiguous():
                t = t.float().contiguous()
            return t

        with torch.cuda.device(self.device):
            for (wname, sl), dst in self._packed.items():
                w = p32(wname + ".weight")
                if sl == "up4":
                    capi.call("advs_pack_upconv_weight", w.data_ptr(), dst.data_ptr(), w.shape[0], w.shape[1],
                              _CAPI_DT[dst.dtype], st)
                    continue
                if sl == "stem64":
                    capi.call("advs_pack_stem_weight_ex", w.data_ptr(), dst.data_ptr(), w.shape[0], w.shape[1],
                              _CAPI_DT[dst.dtype], st)
                    continue
                if sl is not None:
                    w = w[:, sl[0]:sl[1]].contiguous()
                O, I, kh, kw = w.shape          # dst may have more (zero) rows than O: the padded head
                capi.call("advs_pack_conv_weight", w.data_ptr(), dst.data_ptr(), O, I, kh, kw, _CAPI_DT[dst.dtype], st)
            for names, dst in self._bias.items():
                acc = p32(names[0] + ".bias").clone()
                for n in names[1:]:
                    acc += p32(n + ".bias")
                dst[:acc.numel()].copy_(acc)          # (head bias may be zero-padded)
            off = 0
            for slot in self.plan.temb_slots:
                self.temb_w[off:off + slot.cout].copy_(p32(slot.weight + ".weight"))
                self.temb_b[off:off + slot.cout].copy_(p32(slot.weight + ".bias"))
                off += slot.cout
            self.te0_w.copy_(p32("time_embed.0.weight"))
            self.te0_b.copy_(p32("time_embed.0.bias"))
            self.te2_w.copy_(p32("time_embed.2.weight"))
            self.te2_b.copy_(p32("time_embed.2.bias"))
            for n, (g, bt) in self._gn.items():
                g.copy_(p32(n + ".weight"))
                bt.copy_(p32(n + ".bias"))
        self._weights_loaded = True
        self.weights_version += 1

    # ---- launch list ----
    def _build_launches(self):
        L = []
        lib, plan, dt = self.lib, self.plan, self.dt
        B = self.B
        self._plans = []
        self.n_kernels = 0
        for op in plan.ops:
            a = op.args
            if op.kind == "stem" and self._stem_sm100_ok(a):
                w, b = self._packed[(a["weight"], "stem64")], self._bias[(a["weight"],)]
                L.append((lib.advs_stem_im2col_ex, (self.x.data_ptr(), self._stem_col.data_ptr(), B, a["H"], a["W"], a["cin"],
                                                    _CAPI_DT[w.dtype]), "stem_im2col"))
                cp = capi.ConvParams()
                cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, a["H"], a["W"], a["cout"], 1, 1
                cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = self._stem_col.data_ptr(), w.data_ptr(), 64, 1
                cp.bias, cp.out_mode, cp.y, cp.dtype = b.data_ptr(), 0, self._ptr(a["dst"]), dt
                cp.operand_f16 = self._operand_bits(["stem_col"], [w.dtype])
                if a["dst"] in self._stat_buf:
                    cp.stats_partial = self._ptr(self._stat_buf[a["dst"]][0])
                    cp.stats_gran = self._stat_buf[a["dst"]][2]
                if a["dst"] in self._lo_buf:
                    cp.y_lo = self._ptr(self._lo_buf[a["dst"]])
                self._keep.append(cp)
                pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
                with torch.cuda.device(self.device):
                    capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
                self._plans.append(pb)
                L.append((lib.advs_conv_sm100_launch, (pb.ptr,), "stem_sm100"))
                self.n_kernels += 2
            elif op.kind == "stem":
                w, b = self._packed[(a["weight"], None)], self._bias[(a["weight"],)]
                L.append((lib.advs_conv3x3_stem, (self.x.data_ptr(), w.data_ptr(), b.data_ptr(), self._ptr(a["dst"]),
                                                  B, a["H"], a["W"], a["cin"], a["cout"], dt), "stem"))
                self.n_kernels += 1
            elif op.kind == "head":
                w, b = self._packed[(a["weight"], None)], self._bias[(a["weight"],)]
                cp = capi.ConvParams()
                cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, a["H"], a["W"], w.shape[0], 1, 1
                cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = self._ptr(a["src"]), w.data_ptr(), a["cin"], 9
                cp.bias = b.data_ptr()
                cp.out_mode, cp.y, cp.cout_valid, cp.dtype = 2, self.eps.data_ptr(), a["cout"], dt
                cp.operand_f16 = self._operand_bits([a["src"]], [w.dtype])
                self._keep.append(cp)
                if self._head_sm100_ok(a):
                    pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
                    with torch.cuda.device(self.device):
                        capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
                    self._plans.append(pb)
                    L.append((lib.advs_conv_sm100_launch, (pb.ptr,), "head_sm100"))
                else:
                    L.append((lib.advs_conv_simt, (C.byref(cp),), "head_simt"))
                self.n_kernels += 1
            elif op.kind == "gn":
                srcs = a["srcs"]
                x0, c0 = self._ptr(srcs[0]), plan.shape(srcs[0])[3]
                x1, c1 = (self._ptr(srcs[1]), plan.shape(srcs[1])[3]) if len(srcs) > 1 else (None, 0)
                g, bt = self._gn[a["weight"]]
                chunks = int(lib.advs_groupnorm_partial_parts(B, a["HW"]))
                parts, ws_off = [], 0
                for sname in srcs:
                    cs = plan.shape(sname)[3]
                    if sname in self._stat_buf:          # statistics came out of the producing conv's epilogue
                        pbuf, np_, gran = self._stat_buf[sname]
                        parts.append((self._ptr(pbuf), np_, gran))
                    else:
                        ptr = self._ptr(a["ws"]) + ws_off
                        ws_off += B * chunks * cs * 2 * 4
                        L.append((lib.advs_groupnorm_partial, (self._ptr(sname), cs, B, a["HW"], ptr, dt), "gn_stats"))
                        parts.append((ptr, chunks, 1))
                        self.n_kernels += 1
                p1, n1, g1 = parts[1] if len(parts) > 1 else (None, 0, 1)
                L.append((lib.advs_groupnorm_finalize_ex, (parts[0][0], c0, parts[0][1], parts[0][2], p1, c1, n1, g1, B, a["HW"],
                                                           a["groups"], 1e-5, g.data_ptr(), bt.data_ptr(), self._ptr(a["ss"])),
                          "gn_finalize"))
                los = [self._ptr(self._lo_buf[sn]) if sn in self._lo_buf else None for sn in srcs] + [None]
                if los[0] or los[1] or a["dst"] in self._gn_f16:
                    L.append((lib.advs_groupnorm_apply_wide, (x0, los[0], c0, x1, los[1], c1, B, a["HW"], self._ptr(a["ss"]),
                                                              1 if a["silu"] else 0, self._ptr(a["dst"]),
                                                              capi.F16 if a["dst"] in self._gn_f16 else capi.BF16), "gn_apply"))
                else:
                    L.append((lib.advs_groupnorm_apply, (x0, c0, x1, c1, B, a["HW"], self._ptr(a["ss"]),
                                                         1 if a["silu"] else 0, self._ptr(a["dst"]), dt), "gn_apply"))
                self.n_kernels += 2
            elif op.kind == "conv":
                cp = capi.ConvParams()
                cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, a["H"], a["W"], a["cout"], a["stride"], len(a["segs"])
                for i, (src, wname, taps, sl) in enumerate(a["segs"]):
                    cp.seg[i].x = self._ptr(src)
                    cp.seg[i].w = self._packed[(wname, sl)].data_ptr()
                    cp.seg[i].C = plan.shape(src)[3]
                    cp.seg[i].taps = taps
                cp.bias = self._bias[tuple(a["bias"])].data_ptr() if a["bias"] else None
                if a["temb"] is not None:
                    cp.temb = self.temb_cur.data_ptr() + 4 * a["temb"]
                    cp.temb_stride = plan.temb_total
                cp.residual = self._ptr(a["residual"]) if a["residual"] else None
                if a["qkv"] is not None:
                    cp.out_mode = 1
                    cp.q, cp.k, cp.vt = (self._ptr(x) for x in a["qkv"])
                    cp.heads = a["heads"]
                    cp.qk_scale = 1.0 / math.sqrt(math.sqrt(a["cout"] // (3 * a["heads"])))
                else:
                    cp.out_mode = 0
                    cp.y = self._ptr(a["dst"])
                cp.dtype = dt
                if a["dst"] in self._stat_buf:
                    cp.stats_partial = self._ptr(self._stat_buf[a["dst"]][0])
                    cp.stats_gran = self._stat_buf[a["dst"]][2]
                if a["dst"] in self._lo_buf:
                    cp.y_lo = self._ptr(self._lo_buf[a["dst"]])
                cp.operand_f16 = self._operand_bits([sg[0] for sg in a["segs"]],
                                                    [self._packed[(sg[1], sg[3])].dtype for sg in a["segs"]])
                self._keep.append(cp)
                if self._conv_sm100_ok(a):
                    pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
                    with torch.cuda.device(self.device):
                        capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
                    self._plans.append(pb)
                    L.append((lib.advs_conv_sm100_launch, (pb.ptr,), "conv_sm100"))
                else:
                    L.append((lib.advs_conv_simt, (C.byref(cp),), "conv_simt"))
                self.n_kernels += 1
            elif op.kind == "attn":
                if self._attn_sm100_ok(a):
                    pb = capi.PlanBuffer(capi.ATTN_PLAN_BYTES)
                    with torch.cuda.device(self.device):
                        capi.call("advs_attention_sm100_plan", self._ptr(a["q"]), self._ptr(a["k"]), self._ptr(a["vt"]),
                                  self._ptr(a["dst"]), B, a["heads"], a["T"], a["dh"], pb.ptr)
                    self._plans.append(pb)
                    L.append((lib.advs_attention_sm100_launch, (pb.ptr,), "attn_sm100"))
                    self.n_kernels += 1
                else:
                    wsb = plan.bufs[a["ws"]]
                    L.append((lib.advs_attention_simt, (self._ptr(a["q"]), self._ptr(a["k"]), self._ptr(a["vt"]),
                                                        self._ptr(a["dst"]), B, a["heads"], a["T"], a["dh"],
                                                        self._ptr(a["ws"]), wsb.elems * 4, dt), "attn_simt"))
                    self.n_kernels += 3
            elif op.kind == "upconv":
                w4, b = self._packed[(a["weight"], "up4")], self._bias[(a["weight"],)]
                for ph in range(4):
                    cp = capi.ConvParams()
                    cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, a["H"], a["W"], a["cout"], 1, 1
                    cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = self._ptr(a["src"]), w4[ph].data_ptr(), a["C"], 4
                    cp.bias, cp.out_mode, cp.y, cp.dtype, cp.up_phase = b.data_ptr(), 0, self._ptr(a["dst"]), dt, ph + 1
                    cp.operand_f16 = self._operand_bits([a["src"]], [w4.dtype])
                    if a["dst"] in self._stat_buf:
                        cp.stats_partial = self._ptr(self._stat_buf[a["dst"]][0])
                        cp.stats_gran = self._stat_buf[a["dst"]][2]
                    if a["dst"] in self._lo_buf:
                        cp.y_lo = self._ptr(self._lo_buf[a["dst"]])
                    self._keep.append(cp)
                    pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
                    with torch.cuda.device(self.device):
                        capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
                    self._plans.append(pb)
                    L.append((lib.advs_conv_sm100_launch, (pb.ptr,), "conv_sm100"))
                    self.n_kernels += 1
            elif op.kind == "up":
                L.append((lib.advs_upsample_nearest2x, (self._ptr(a["src"]), self._ptr(a["dst"]), B, a["H"], a["W"],
                                                        a["C"], dt), "upsample"))
                self.n_kernels += 1
            else:
                raise AssertionError(op.kind)
        self._launches = L

    # ---- execution ----
    def run(self):
        """eps = UNet(self.x; self.temb_cur) -- enqueue the whole forward on the current stream."""
        st = _stream_ptr()
        last_error = self.lib.advs_last_error
        for fn, args, name in self._launches:
            rc = fn(*args, st)
            if rc:
                raise capi.AdvsError(f"{name} failed (rc={rc}): {last_error().decode()}")

    def temb_table(self, timesteps: torch.Tensor) -> torch.Tensor:
        """[Nt] int64 timesteps -> [Nt, temb_total] stacked per-block projections
        Linear(SiLU(time_embed(timestep_embedding(t))))  (dm1:254, dm1:77-80,101)."""
        t = timesteps.to(device=self.device, dtype=torch.int64).contiguous()
        nt = t.numel()
        mc, ted = self.spec.model_channels, self.spec.time_embed_dim
        dev = self.device
        e0 = torch.empty(nt, mc, dtype=torch.float32, device=dev)
        if mc % 2:
            raise ValueError("model_channels must be even")
        e1 = torch.empty(nt, ted, dtype=torch.float32, device=dev)
        e2 = torch.empty(nt, ted, dtype=torch.float32, device=dev)
        out = torch.empty(nt, self.plan.temb_total, dtype=torch.float32, device=dev)
        st = _stream_ptr()
        with torch.cuda.device(dev):
            capi.call("advs_timestep_embedding", t.data_ptr(), nt, self.freqs.data_ptr(), mc // 2, e0.data_ptr(), st)
            capi.call("advs_linear_f32", e0.data_ptr(), self.te0_w.data_ptr(), self.te0_b.data_ptr(), e1.data_ptr(),
                      nt, mc, ted, 0, 1, st)
            capi.call("advs_linear_f32", e1.data_ptr(), self.te2_w.data_ptr(), self.te2_b.data_ptr(), e2.data_ptr(),
                      nt, ted, ted, 0, 0, st)
            capi.call("advs_linear_f32", e2.data_ptr(), self.temb_w.data_ptr(), self.temb_b.data_ptr(), out.data_ptr(),
                      nt, ted, self.plan.temb_total, 1, 0, st)
        self._last_emb = e2
        return out

    def forward(self, x: torch.Tensor, timesteps: torch.Tensor) -> torch.Tensor:
        if tuple(x.shape) != tuple(self.x.shape):
            raise ValueError(f"engine built for input {tuple(self.x.shape)}, got {tuple(x.shape)}")
        with torch.cuda.device(self.device):
            self.x.copy_(x)
            table = self.temb_table(timesteps.reshape(-1))
            if table.shape[0] == 1 and self.B > 1:
                table = table.expand(self.B, -1)
            if table.shape[0] != self.B:
                raise ValueError("timesteps must have one entry per batch element")
            self.temb_cur.copy_(table)
            self.run()
            return self.eps.clone()

    # ---- measurement helpers (bench.py) ----
    def launch_costs(self):
        """Per launch-list entry: (name, algorithmic flops, algorithmic bytes) -- DESIGN.md section 'rooflines'."""
        out = []
        plan, ab = self.plan, self.act_bytes
        it = iter(self._launches)
        for op in plan.ops:
            a = op.args
            if op.kind == "gn":
                n = self.B * a["HW"] * a["C"]
                for sname in a["srcs"]:
                    if sname not in self._stat_buf:
                        out.append(("gn_stats", 0, self.B * a["HW"] * plan.shape(sname)[3] * ab))   # one read
                out.append(("gn_finalize", 0, 0))
                wide = sum(self.B * a["HW"] * plan.shape(sn)[3] for sn in a["srcs"] if sn in self._lo_buf)
                out.append(("gn_apply", 0, 2 * n * ab + wide))       # one read + one write (+ the int8 extensions)
            elif op.kind == "conv":
                k = sum(plan.shape(s)[3] * taps for (s, _, taps, _) in a["segs"])
                m = self.B * a["H"] * a["W"]
                byts = sum(self.B * a["H"] * a["W"] * (a["stride"] ** 2 if i == 0 else 1) * plan.shape(s)[3]
                           for i, (s, _, _, _) in enumerate(a["segs"])) * ab
                byts += k * a["cout"] * ab + m * a["cout"] * ab * (2 if a["residual"] else 1)
                byts += m * a["cout"] if a.get("dst") in self._lo_buf else 0
                name = "conv_sm100" if self._conv_sm100_ok(a) else "conv_simt"
                out.append((name, 2 * m * k * a["cout"], byts))
            elif op.kind == "attn":
                c = a["heads"] * a["dh"]
                name = "attn_sm100" if self._attn_sm100_ok(a) else "attn_simt"
                out.append((name, 4 * self.B * a["T"] * a["T"] * c, 4 * self.B * a["T"] * c * ab))
            elif op.kind == "upconv":
                m = self.B * a["H"] * a["W"]
                for _ in range(4):     # executed work: 4 taps per output pixel instead of 9
                    out.append(("conv_sm100", 2 * m * 4 * a["C"] * a["cout"],
                                (m * a["C"] + 4 * a["C"] * a["cout"] + m * a["cout"]) * ab))
            elif op.kind == "up":
                n = self.B * a["H"] * a["W"] * a["C"]
                out.append(("upsample", 0, 5 * n * ab))
            elif op.kind == "stem" and self._stem_sm100_ok(a):
                m = self.B * a["H"] * a["W"]
                out.append(("stem_im2col", 0, m * (a["cin"] * 4 + 64 * ab)))
                out.append(("stem_sm100", 2 * m * 9 * a["cin"] * a["cout"], m * (64 + a["cout"]) * ab))
            elif op.kind == "stem":
                m = self.B * a["H"] * a["W"]
                out.append(("stem", 2 * m * 9 * a["cin"] * a["cout"], m * (a["cin"] * 4 + a["cout"] * ab)))
            elif op.kind == "head":
                m = self.B * a["H"] * a["W"]
                out.append(("head_sm100" if self._head_sm100_ok(a) else "head_simt",
                            2 * m * 9 * a["cin"] * a["cout"], m * (a["cin"] * ab + a["cout"] * 4)))
        assert len(out) == len(self._launches)
        return out

    def profile_forward(self, repeats=1, warmup=0, per_repeat=None):
        """Run the forward eagerly with a CUDA-event pair around every launch (on the launching stream); launches are
        queued back to back, the host only synchronises once per repeat.  Returns {name: dict(ms, flops, bytes,
        launches)} summed over the forward and averaged over `repeats` (after `warmup` untimed repeats); with
        `per_repeat=<name>` also the list of that class's summed ms in every repeat (for min / median)."""
        costs = self.launch_costs()
        st = _stream_ptr()
        agg, series = {}, []
        with torch.cuda.device(self.device):
            for _ in range(warmup):
                self.run()
            for _ in range(repeats):
                evs = []
                for (fn, args, name) in self._launches:
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    rc = fn(*args, st)
                    e1.record()
                    if rc:
                        raise capi.AdvsError(f"{name} failed: {self.lib.advs_last_error().decode()}")
                    evs.append((e0, e1))
                torch.cuda.synchronize(self.device)
                tot = 0.0
                for (e0, e1), (name, fl, by) in zip(evs, costs):
                    d = agg.setdefault(name, dict(ms=0.0, flops=0, bytes=0, launches=0))
                    ms = e0.elapsed_time(e1)
                    if name == per_repeat:
                        tot += ms
                    d["ms"] += ms / repeats
                    d["flops"] += fl / repeats
                    d["bytes"] += by / repeats
                    d["launches"] += 1.0 / repeats
                series.append(tot)
        return (agg, series) if per_repeat is not None else agg
