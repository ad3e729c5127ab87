"""Static execution plan of one UNet forward: an ordered list of kernel-level ops over arena buffers.

The graph wiring follows the reference's UNetModel.__init__/forward (diff_model.py:158-267; the
diff_model2.py copy differs only in its default arguments) but is re-expressed for the B200 path:

* activations are NHWC; `torch.cat([h, skip], 1)` (dm1:265) never materialises on its own -- the
  GroupNorm that consumes it reads both sources and writes the normalised concat, and the 1x1
  shortcut conv reads the two sources as two K-segments;
* ResidualBlock (dm1:67-103) = GN+SiLU -> conv3x3(+bias +time-embedding) -> GN+SiLU ->
  conv3x3(+bias, + identity residual | + fused 1x1 shortcut K-segments);
* AttentionBlock (dm1:107-127) = GN -> 1x1 qkv conv with a q/k/v^T splitting epilogue ->
  fused attention -> 1x1 proj conv (+bias +residual);
* all per-block time-embedding projections (dm1:77-80) are one stacked linear, evaluated once per
  timestep outside this plan; each ResidualBlock owns a column range of that table.

Pure Python, no torch: the plan is data.  `engine.py` binds it to device memory and the C ABI;
`tests/plan_interp.py` replays it with PyTorch ops to check the wiring on CPU.
"""
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Tuple


@dataclass(frozen=True)
class UNetSpec:
    in_channels: int = 3
    model_channels: int = 128
    out_channels: int = 3
    num_res_blocks: int = 2
    attention_resolutions: Tuple[int, ...] = (8, 16)
    channel_mult: Tuple[int, ...] = (1, 2, 2, 2)
    conv_resample: bool = True
    num_heads: int = 4
    groups: int = 32

    @property
    def time_embed_dim(self):
        return self.model_channels * 4


@dataclass
class Buf:
    name: str
    shape: Tuple[int, ...]      # logical shape (NHWC for activations)
    kind: str                   # "act" (engine dtype) | "f32" | "u8"
    elems: int = 0
    first: int = -1             # first op index that touches it
    last: int = -1              # last op index that touches it
    offset: int = -1            # byte offset in the arena (set by assign_offsets)
    nbytes: int = 0


@dataclass
class Op:
    kind: str                   # stem | gn | conv | attn | up | upconv | head
    args: dict = field(default_factory=dict)


@dataclass
class TembSlot:
    weight: str                 # "<block>.time_emb.1"
    offset: int                 # column offset in the stacked projection table
    cout: int


class Plan:
    def __init__(self, spec: UNetSpec, B: int, H: int, W: int):
        self.spec, self.B, self.H, self.W = spec, B, H, W
        self.ops: List[Op] = []
        self.bufs: Dict[str, Buf] = {}
        self.temb_slots: List[TembSlot] = []
        self.temb_total = 0
        self.arena_bytes = 0
        self._uid = 0
        self.flops = 0              # 2*MAC of convs + attention, for the roofline
        self.gn_elems = 0           # elements normalised by GroupNorm (K5 traffic)

    # ---- construction helpers ----
    def new_buf(self, tag, shape, kind="act"):
        self._uid += 1
        name = f"{tag}#{self._uid}"
        n = 1
        for s in shape:
            n *= s
        self.bufs[name] = Buf(name, tuple(shape), kind, n)
        return name

    def emit(self, kind, reads, writes, **args):
        idx = len(self.ops)
        for b in list(reads) + list(writes):
            buf = self.bufs[b]
            if buf.first < 0:
                buf.first = idx
            buf.last = idx
        args["reads"], args["writes"] = list(reads), list(writes)
        self.ops.append(Op(kind, args))

    def shape(self, b):
        return self.bufs[b].shape

    # ---- memory planning: first-fit with liveness (ops run in order on one stream) ----
    def assign_offsets(self, act_bytes: int, align: int = 1024):
        free: List[List[int]] = []   # [offset, size], sorted by offset
        end = 0
        by_first: Dict[int, List[Buf]] = {}
        by_last: Dict[int, List[Buf]] = {}
        for b in self.bufs.values():
            b.nbytes = b.elems * {"act": act_bytes, "f32": 4, "u8": 1}[b.kind]
            b.nbytes = (b.nbytes + align - 1) // align * align
            by_first.setdefault(b.first, []).append(b)
            by_last.setdefault(b.last, []).append(b)
        for i in range(len(self.ops)):
            for b in by_first.get(i, []):
                placed = False
                for slot in free:
                    if slot[1] >= b.nbytes:
                        b.offset = slot[0]
                        slot[0] += b.nbytes
                        slot[1] -= b.nbytes
                        placed = True
                        break
                free[:] = [s for s in free if s[1] > 0]
                if not placed:
                    # grow: merge with a trailing free block if there is one
                    if free and free[-1][0] + free[-1][1] == end:
                        b.offset = free[-1][0]
                        end = b.offset + b.nbytes
                        free.pop()
                    else:
                        b.offset = end
                        end += b.nbytes
            for b in by_last.get(i, []):
                free.append([b.offset, b.nbytes])
                free.sort()
                merged: List[List[int]] = []
                for s in free:
                    if merged and merged[-1][0] + merged[-1][1] == s[0]:
                        merged[-1][1] += s[1]
                    else:
                        merged.append(s)
                free[:] = merged
        self.arena_bytes = end
        return end


def _conv_flops(B, H, W, cin, cout, taps):
    return 2 * B * H * W * cin * cout * taps


def build_unet_plan(spec: UNetSpec, B: int, H: int, W: int, attn_scores_ws: bool = False,
                    fuse_upsample: bool = False) -> Plan:
    """Op list for eps = UNetModel(x, t).  `attn_scores_ws`: reserve the [B*heads, T, T] fp32 score
    buffer the SIMT attention needs (the tcgen05 flash kernel needs none); a bool, or a predicate (T, dh) -> bool
    deciding per attention block.  `fuse_upsample`: emit
    Upsample's nearest-2x + conv3x3 (dm1:137-139) as one "upconv" op (four 2x2 phase convolutions on the
    low-res tensor, 2.25x fewer MACs, no upsampled intermediate) instead of "up" + "conv"."""
    nlev = len(spec.channel_mult)
    if H % (1 << (nlev - 1)) or W % (1 << (nlev - 1)):
        raise ValueError(f"image size {H}x{W} is not divisible by 2^{nlev - 1} (one stride-2 conv per level)")
    p = Plan(spec, B, H, W)
    mc, G = spec.model_channels, spec.groups

    def gn(srcs, wname, silu, h, w):
        c = sum(p.shape(s)[3] for s in srcs)
        dst = p.new_buf("gn", (B, h, w, c))
        ss = p.new_buf("gn_ss", (B, c, 2), "f32")
        ws = p.new_buf("gn_ws", (0,), "f32")     # sized by the engine (advs_groupnorm_workspace_bytes)
        p.bufs[ws].shape = ("gn_ws", B, h * w, c)
        p.emit("gn", srcs, [dst, ss, ws], srcs=list(srcs), dst=dst, ss=ss, ws=ws, weight=wname, silu=silu,
               B=B, HW=h * w, C=c, groups=G)
        p.gn_elems += B * h * w * c
        return dst

    def res_block(prefix, srcs, cout, h, w):
        cin = sum(p.shape(s)[3] for s in srcs)
        a = gn(srcs, prefix + ".conv1.0", True, h, w)
        slot = TembSlot(prefix + ".time_emb.1", p.temb_total, cout)
        p.temb_slots.append(slot)
        p.temb_total += cout
        hbuf = p.new_buf("res_h", (B, h, w, cout))
        p.emit("conv", [a], [hbuf], segs=[(a, prefix + ".conv1.2", 9, None)], stride=1, H=h, W=w, cout=cout,
               bias=[prefix + ".conv1.2"], temb=slot.offset, residual=None, dst=hbuf, qkv=None)
        p.flops += _conv_flops(B, h, w, cin, cout, 9)
        bb = gn([hbuf], prefix + ".conv2.0", True, h, w)
        out = p.new_buf("res_out", (B, h, w, cout))
        segs = [(bb, prefix + ".conv2.3", 9, None)]
        bias = [prefix + ".conv2.3"]
        residual = None
        if cin != cout:
            # 1x1 shortcut conv fused as extra K-segments; (lo, hi) = input-channel slice of its weight
            lo = 0
            for s in srcs:
                c = p.shape(s)[3]
                segs.append((s, prefix + ".shortcut", 1, (lo, lo + c)))
                lo += c
            bias.append(prefix + ".shortcut")
            p.flops += _conv_flops(B, h, w, cin, cout, 1)
        else:
            assert len(srcs) == 1
            residual = srcs[0]
        reads = [bb] + [s for s in srcs]
        p.emit("conv", reads, [out], segs=segs, stride=1, H=h, W=w, cout=cout, bias=bias, temb=None,
               residual=residual, dst=out, qkv=None)
        p.flops += _conv_flops(B, h, w, cout, cout, 9)
        return out

    def attn_block(prefix, x, h, w):
        c = p.shape(x)[3]
        heads = spec.num_heads
        assert c % heads == 0, "channels must be divisible by num_heads"   # dm1:111
        dh, T = c // heads, h * w
        a = gn([x], prefix + ".norm", False, h, w)
        q = p.new_buf("attn_q", (B, heads, T, dh))
        k = p.new_buf("attn_k", (B, heads, T, dh))
        vt = p.new_buf("attn_vt", (B, heads, dh, T))
        p.emit("conv", [a], [q, k, vt], segs=[(a, prefix + ".qkv", 1, None)], stride=1, H=h, W=w, cout=3 * c,
               bias=[], temb=None, residual=None, dst=None, qkv=(q, k, vt), heads=heads)
        p.flops += _conv_flops(B, h, w, c, 3 * c, 1)
        o = p.new_buf("attn_o", (B, h, w, c))
        writes = [o]
        ws = None
        if attn_scores_ws(T, dh) if callable(attn_scores_ws) else attn_scores_ws:
            ws = p.new_buf("attn_scores", (B * heads, T, T), "f32")
            writes.append(ws)
        p.emit("attn", [q, k, vt], writes, q=q, k=k, vt=vt, dst=o, ws=ws, B=B, heads=heads, T=T, dh=dh)
        p.flops += 4 * B * T * T * c
        out = p.new_buf("attn_out", (B, h, w, c))
        p.emit("conv", [o, x], [out], segs=[(o, prefix + ".proj", 1, None)], stride=1, H=h, W=w, cout=c,
               bias=[prefix + ".proj"], temb=None, residual=x, dst=out, qkv=None)
        p.flops += _conv_flops(B, h, w, c, c, 1)
        return out

    # ---- down path (dm1:190-214) ----
    h, w = H, W
    cur = p.new_buf("stem", (B, h, w, mc))
    p.emit("stem", [], [cur], dst=cur, weight="down_blocks.0.0", cin=spec.in_channels, cout=mc, H=h, W=w)
    p.flops += _conv_flops(B, h, w, spec.in_channels, mc, 9)
    hs = [cur]
    ch, ds, idx = mc, 1, 1
    for level, mult in enumerate(spec.channel_mult):
        for _ in range(spec.num_res_blocks):
            cur = res_block(f"down_blocks.{idx}.0", [cur], mult * mc, h, w)
            ch = mult * mc
            if ds in spec.attention_resolutions:
                cur = attn_block(f"down_blocks.{idx}.1", cur, h, w)
            hs.append(cur)
            idx += 1
        if level != nlev - 1:
            if not spec.conv_resample:
                # the reference builds nn.AvgPool2d(stride=2) here, which raises TypeError (dm1:150)
                raise TypeError("AvgPool2d.__init__() missing 1 required positional argument: 'kernel_size'")
            dst = p.new_buf("down", (B, h // 2, w // 2, ch))
            p.emit("conv", [cur], [dst], segs=[(cur, f"down_blocks.{idx}.0.op", 9, None)], stride=2, H=h // 2,
                   W=w // 2, cout=ch, bias=[f"down_blocks.{idx}.0.op"], temb=None, residual=None, dst=dst, qkv=None)
            p.flops += _conv_flops(B, h // 2, w // 2, ch, ch, 9)
            cur = dst
            h, w = h // 2, w // 2
            hs.append(cur)
            ds *= 2
            idx += 1

    # ---- middle (dm1:217-221) ----
    cur = res_block("middle_block.0", [cur], ch, h, w)
    cur = attn_block("middle_block.1", cur, h, w)
    cur = res_block("middle_block.2", [cur], ch, h, w)

    # ---- up path (dm1:224-238, forward dm1:264-266) ----
    uidx = 0
    for level, mult in list(enumerate(spec.channel_mult))[::-1]:
        for i in range(spec.num_res_blocks + 1):
            skip = hs.pop()
            cur = res_block(f"up_blocks.{uidx}.0", [cur, skip], mc * mult, h, w)
            ch = mc * mult
            sub = 1
            if ds in spec.attention_resolutions:
                cur = attn_block(f"up_blocks.{uidx}.{sub}", cur, h, w)
                sub += 1
            if level and i == spec.num_res_blocks and spec.conv_resample and fuse_upsample:
                dst = p.new_buf("upconv", (B, 2 * h, 2 * w, ch))
                p.emit("upconv", [cur], [dst], src=cur, dst=dst, weight=f"up_blocks.{uidx}.{sub}.conv", H=h, W=w, C=ch,
                       cout=ch)
                p.flops += _conv_flops(B, 2 * h, 2 * w, ch, ch, 9)      # the reference's algorithmic count
                h, w = 2 * h, 2 * w
                cur = dst
                ds //= 2
            elif level and i == spec.num_res_blocks:
                up = p.new_buf("up", (B, 2 * h, 2 * w, ch))
                p.emit("up", [cur], [up], src=cur, dst=up, H=h, W=w, C=ch)
                h, w = 2 * h, 2 * w
                cur = up
                if spec.conv_resample:
                    dst = p.new_buf("upconv", (B, h, w, ch))
                    p.emit("conv", [cur], [dst], segs=[(cur, f"up_blocks.{uidx}.{sub}.conv", 9, None)], stride=1,
                           H=h, W=w, cout=ch, bias=[f"up_blocks.{uidx}.{sub}.conv"], temb=None, residual=None,
                           dst=dst, qkv=None)
                    p.flops += _conv_flops(B, h, w, ch, ch, 9)
                    cur = dst
                ds //= 2
            uidx += 1

    # ---- head (dm1:240-243): GN -> SiLU -> conv3x3(model_channels -> out_channels) ----
    if ch != mc:
        raise ValueError("UNetModel.out expects model_channels inputs: channel_mult[0] must be 1 (dm1:240-242)")
    a = gn([cur], "out.0", True, h, w)
    p.emit("head", [a], [], src=a, weight="out.2", cin=mc, cout=spec.out_channels, H=h, W=w)
    p.flops += _conv_flops(B, h, w, mc, spec.out_channels, 9)
    return p


def parameter_shapes(spec: UNetSpec) -> Dict[str, Tuple[int, ...]]:
    """state_dict key -> shape, in the reference's registration order (dm1:183-243)."""
    out: Dict[str, Tuple[int, ...]] = {}
    mc, ted = spec.model_channels, spec.time_embed_dim

    def conv(name, cin, cout, k, bias=True):
        out[name + ".weight"] = (cout, cin, k, k)
        if bias:
            out[name + ".bias"] = (cout,)

    def norm(name, c):
        out[name + ".weight"] = (c,)
        out[name + ".bias"] = (c,)

    def linear(name, i, o):
        out[name + ".weight"] = (o, i)
        out[name + ".bias"] = (o,)

    def res(prefix, cin, cout):
        norm(prefix + ".conv1.0", cin)
        conv(prefix + ".conv1.2", cin, cout, 3)
        linear(prefix + ".time_emb.1", ted, cout)
        norm(prefix + ".conv2.0", cout)
        conv(prefix + ".conv2.3", cout, cout, 3)
        if cin != cout:
            conv(prefix + ".shortcut", cin, cout, 1)

    def attn(prefix, c):
        norm(prefix + ".norm", c)
        conv(prefix + ".qkv", c, 3 * c, 1, bias=False)
        conv(prefix + ".proj", c, c, 1)

    linear("time_embed.0", mc, ted)
    linear("time_embed.2", ted, ted)
    conv("down_blocks.0.0", spec.in_channels, mc, 3)
    chans = [mc]
    ch, ds, idx = mc, 1, 1
    nlev = len(spec.channel_mult)
    for level, mult in enumerate(spec.channel_mult):
        for _ in range(spec.num_res_blocks):
            res(f"down_blocks.{idx}.0", ch, mult * mc)
            ch = mult * mc
            if ds in spec.attention_resolutions:
                attn(f"down_blocks.{idx}.1", ch)
            chans.append(ch)
            idx += 1
        if level != nlev - 1:
            if spec.conv_resample:
                conv(f"down_blocks.{idx}.0.op", ch, ch, 3)
            chans.append(ch)
            ds *= 2
            idx += 1
    res("middle_block.0", ch, ch)
    attn("middle_block.1", ch)
    res("middle_block.2", ch, ch)
    uidx = 0
    for level, mult in list(enumerate(spec.channel_mult))[::-1]:
        for i in range(spec.num_res_blocks + 1):
            res(f"up_blocks.{uidx}.0", ch + chans.pop(), mc * mult)
            ch = mc * mult
            sub = 1
            if ds in spec.attention_resolutions:
                attn(f"up_blocks.{uidx}.{sub}", ch)
                sub += 1
            if level and i == spec.num_res_blocks:
                if spec.conv_resample:
                    conv(f"up_blocks.{uidx}.{sub}.conv", ch, ch, 3)
                ds //= 2
            uidx += 1
    norm("out.0", ch)
    conv("out.2", mc, spec.out_channels, 3)
    return out
