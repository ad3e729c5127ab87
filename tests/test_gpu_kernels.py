"""Per-kernel numerics on the B200, through the C ABI (ops.py -> ctypes -> libadvshadow_b200.so).
Each kernel is compared with a plain PyTorch fp32 evaluation of the same reference op
(nn.Conv2d / GroupNorm / softmax-einsum: dm1:62-127) on the same inputs."""
import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ops():
    import advshadow_b200
    from advshadow_b200 import ops as o
    assert torch.cuda.is_available()
    # the PyTorch side is the fp32 reference: keep cuDNN / cuBLAS from silently using TF32
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return o


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    return ((a.float() - b.float()).abs().max() / b.float().abs().max().clamp_min(1e-6)).item()


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 1.5e-2)])
@pytest.mark.parametrize("c0,c1,silu", [(128, 0, True), (256, 128, True), (384, 0, False), (1024, 512, True)])
def test_groupnorm(ops, dtype, tol, c0, c1, silu):
    torch.manual_seed(0)
    B, H, W = 3, 16, 24
    x0 = (torch.randn(B, H, W, c0, device="cuda") * 2 + 0.5).to(dtype)
    x1 = (torch.randn(B, H, W, c1, device="cuda") - 1).to(dtype) if c1 else None
    g = torch.randn(c0 + c1, device="cuda")
    b = torch.randn(c0 + c1, device="cuda")
    y = ops.groupnorm(x0, x1, g, b, silu=silu)
    xin = torch.cat([x0, x1], 3) if c1 else x0
    ref = F.group_norm(nchw(xin.float()), 32, g, b, 1e-5)
    if silu:
        ref = F.silu(ref)
    assert rel_err(nchw(y), ref) < tol


CONV_CASES = [
    # B, H, W, cin, cout, taps, stride
    (2, 16, 16, 128, 128, 9, 1),
    (1, 32, 32, 64, 256, 9, 1),
    (3, 8, 8, 256, 128, 9, 1),      # tile spans images (tn > 1)
    (2, 16, 16, 128, 192, 1, 1),    # 1x1, Cout not a multiple of 128
    (2, 16, 16, 128, 128, 9, 2),    # stride 2 (output 16x16 from 32x32)
    (1, 28, 28, 64, 64, 9, 1),      # non power-of-two width (224/8)
    (1, 4, 256, 64, 128, 9, 1),     # W > 128: row segments (halo-reuse kernel)
    (2, 3, 128, 128, 256, 9, 1),    # W = 128, BN = 256 halo kernel, odd number of pixel tiles per image
    (1, 5, 128, 64, 64, 1, 1),      # 1x1 at W = 128 (no halo), odd tile count -> out-of-range peer tile
]


def conv_ref(x_nhwc, w, stride, bias=None):
    pad = 1 if w.shape[-1] == 3 else 0
    return F.conv2d(nchw(x_nhwc.float()), w.float(), bias, stride=stride, padding=pad)


@pytest.mark.parametrize("impl,dtype,tol", [("simt", torch.float32, 2e-5), ("simt", torch.bfloat16, 1e-2),
                                            ("sm100", torch.bfloat16, 1e-2)])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_plain(ops, impl, dtype, tol, case):
    B, H, W, cin, cout, taps, stride = case
    k = 3 if taps == 9 else 1
    torch.manual_seed(1)
    x = torch.randn(B, H * stride, W * stride, cin, device="cuda").to(dtype)
    w = (torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(cin * taps)).to(dtype).float()
    bias = torch.randn(cout, device="cuda")
    wp = ops.pack_conv_weight(w, dtype)
    y = ops.conv([(x, wp)], B, H, W, cout, stride=stride, bias=bias, impl=impl)
    torch.cuda.synchronize()
    ref = conv_ref(x, w, stride, bias)
    assert rel_err(nchw(y), ref) < tol


@pytest.mark.parametrize("H,W", [(16, 16), (3, 128)])
@pytest.mark.parametrize("impl,dtype,tol", [("simt", torch.float32, 2e-5), ("sm100", torch.bfloat16, 1e-2)])
def test_conv_fused_resblock_tail(ops, impl, dtype, tol, H, W):
    """conv2 of a ResidualBlock whose input is a virtual concat: 3x3 conv + two 1x1 shortcut
    K-segments + bias, and conv1 with the per-image time-embedding bias (dm1:94-103)."""
    torch.manual_seed(2)
    B, c0, c1, cout = 2, 128, 64, 128
    h = torch.randn(B, H, W, cout, device="cuda").to(dtype)
    xa = torch.randn(B, H, W, c0, device="cuda").to(dtype)
    xb = torch.randn(B, H, W, c1, device="cuda").to(dtype)
    w2 = (torch.randn(cout, cout, 3, 3, device="cuda") / 34).to(dtype).float()
    ws = (torch.randn(cout, c0 + c1, 1, 1, device="cuda") / 14).to(dtype).float()
    bias = torch.randn(cout, device="cuda")
    segs = [(h, ops.pack_conv_weight(w2, dtype)), (xa, ops.pack_conv_weight(ws[:, :c0].contiguous(), dtype)),
            (xb, ops.pack_conv_weight(ws[:, c0:].contiguous(), dtype))]
    y = ops.conv(segs, B, H, W, cout, bias=bias, impl=impl)
    ref = conv_ref(h, w2, 1) + conv_ref(torch.cat([xa, xb], 3), ws, 1) + bias[None, :, None, None]
    assert rel_err(nchw(y), ref) < tol
    # conv1: + temb[b, :] ; identity residual
    temb = torch.randn(B, 512, device="cuda")
    y = ops.conv([(h, segs[0][1])], B, H, W, cout, bias=bias, temb=temb[:, 128:256].contiguous(), residual=xa, impl=impl)
    ref = conv_ref(h, w2, 1, bias) + temb[:, 128:256, None, None] + nchw(xa.float())
    assert rel_err(nchw(y), ref) < tol


@pytest.mark.parametrize("impl,dtype,tol", [("simt", torch.float32, 2e-5), ("sm100", torch.bfloat16, 1e-2)])
def test_conv_qkv_split(ops, impl, dtype, tol):
    torch.manual_seed(3)
    B, H, W, c, heads = 2, 16, 8, 256, 4
    dh = c // heads
    x = torch.randn(B, H, W, c, device="cuda").to(dtype)
    w = (torch.randn(3 * c, c, 1, 1, device="cuda") / 16).to(dtype).float()
    q, k, vt = ops.conv([(x, ops.pack_conv_weight(w, dtype))], B, H, W, 3 * c, qkv_heads=heads, impl=impl)
    ref = conv_ref(x, w, 1).reshape(B, heads, 3, dh, H * W)       # dm1:119-120 per-head [q|k|v] interleave
    s = 1 / math.sqrt(math.sqrt(dh))
    assert rel_err(q, (ref[:, :, 0] * s).transpose(2, 3)) < tol
    assert rel_err(k, (ref[:, :, 1] * s).transpose(2, 3)) < tol
    assert rel_err(vt, ref[:, :, 2]) < tol


@pytest.mark.parametrize("impl,dtype,tol", [("simt", torch.float32, 2e-5), ("simt", torch.bfloat16, 1e-2),
                                            ("sm100", torch.bfloat16, 1.5e-2)])
@pytest.mark.parametrize("T,dh", [(128, 64), (256, 128), (512, 256), (1024, 128),
                                  (4096, 128), (1024, 256)])      # the two shapes of the headline config (dm2 at 256x256)
def test_attention(ops, impl, dtype, tol, T, dh):
    torch.manual_seed(4)
    B, heads = 2, 2
    q = (torch.randn(B, heads, T, dh, device="cuda") * 0.7).to(dtype)
    k = (torch.randn(B, heads, T, dh, device="cuda") * 0.7).to(dtype)
    vt = torch.randn(B, heads, dh, T, device="cuda").to(dtype)
    o = ops.attention(q, k, vt, impl=impl)
    torch.cuda.synchronize()
    p = torch.einsum("bhtd,bhsd->bhts", q.float(), k.float()).softmax(-1)
    ref = torch.einsum("bhts,bhds->bthd", p, vt.float()).reshape(B, T, heads * dh)
    assert rel_err(o, ref) < tol


def test_attention_lazy_rescale_path(ops):
    """keys whose scores grow block after block force the running-max rescale of the O accumulator."""
    torch.manual_seed(5)
    B, heads, T, dh = 1, 1, 1024, 128
    q = torch.ones(B, heads, T, dh, device="cuda").to(torch.bfloat16) * 0.25
    ramp = torch.linspace(0, 3.0, T, device="cuda")[None, None, :, None]
    k = (torch.ones(B, heads, T, dh, device="cuda") * ramp * 0.25).to(torch.bfloat16)
    vt = torch.randn(B, heads, dh, T, device="cuda").to(torch.bfloat16)
    o = ops.attention(q, k, vt, impl="sm100")
    p = torch.einsum("bhtd,bhsd->bhts", q.float(), k.float()).softmax(-1)
    ref = torch.einsum("bhts,bhds->bthd", p, vt.float()).reshape(B, T, heads * dh)
    assert rel_err(o, ref) < 2e-2


@pytest.mark.parametrize("T,dh", [(384, 64), (384, 128), (640, 128), (384, 256), (640, 256)])
def test_attention_odd_block_counts(ops, T, dh):
    """3 and 5 key blocks: the even-block softmax set runs one block more than the odd-block set."""
    torch.manual_seed(14)
    B, heads = 1, 3
    q = (torch.randn(B, heads, T, dh, device="cuda") * 0.8).to(torch.bfloat16)
    k = (torch.randn(B, heads, T, dh, device="cuda") * 0.8).to(torch.bfloat16)
    vt = torch.randn(B, heads, dh, T, device="cuda").to(torch.bfloat16)
    o = ops.attention(q, k, vt, impl="sm100")
    p = torch.einsum("bhtd,bhsd->bhts", q.float(), k.float()).softmax(-1)
    ref = torch.einsum("bhts,bhds->bthd", p, vt.float()).reshape(B, T, heads * dh)
    assert rel_err(o, ref) < 1.5e-2


@pytest.mark.parametrize("T,dh,top,descending", [(640, 128, 8.0, False), (1024, 64, 8.0, False), (640, 256, 6.0, False),
                                                 (640, 128, 8.0, True)])
def test_attention_running_max_handover(ops, T, dh, top, descending):
    """Row maxima that grow by more than the lazy threshold in EVERY key block (both softmax sets rescale O and
    correct each other's row sums), and the mirror case where block 0 holds the maximum and nothing ever grows."""
    torch.manual_seed(15)
    B, heads = 1, 2
    q = torch.ones(B, heads, T, dh, device="cuda").to(torch.bfloat16) * 0.25
    ramp = torch.linspace(0, top, T, device="cuda")
    if descending:
        ramp = ramp.flip(0)
    k = (torch.ones(B, heads, T, dh, device="cuda") * ramp[None, None, :, None] * (32.0 / dh)).to(torch.bfloat16)
    vt = torch.randn(B, heads, dh, T, device="cuda").to(torch.bfloat16)
    o = ops.attention(q, k, vt, impl="sm100")
    p = torch.einsum("bhtd,bhsd->bhts", q.float(), k.float()).softmax(-1)
    ref = torch.einsum("bhts,bhds->bthd", p, vt.float()).reshape(B, T, heads * dh)
    assert torch.isfinite(o.float()).all()
    assert rel_err(o, ref) < 2e-2


def test_attention_repeatable_over_shapes(ops):
    """Race detector for the warp-specialised attention kernel: random shapes, every launch repeated on the same
    inputs must be bit-identical (the kernel has no atomics; the only cross-warp traffic goes through mbarriers),
    and every result must match the fp32 reference."""
    gen = torch.Generator(device="cuda").manual_seed(99)
    cpu = torch.Generator().manual_seed(99)
    for it in range(12):
        T = 128 * int(torch.randint(1, 17, (1,), generator=cpu))
        dh = (64, 128, 256)[int(torch.randint(0, 3, (1,), generator=cpu))]
        B, heads = int(torch.randint(1, 5, (1,), generator=cpu)), int(torch.randint(1, 7, (1,), generator=cpu))
        scale = (0.3, 0.8, 1.5)[it % 3]
        q = (torch.randn(B, heads, T, dh, device="cuda", generator=gen) * scale).to(torch.bfloat16)
        k = (torch.randn(B, heads, T, dh, device="cuda", generator=gen) * scale).to(torch.bfloat16)
        vt = torch.randn(B, heads, dh, T, device="cuda", generator=gen).to(torch.bfloat16)
        outs = [ops.attention(q, k, vt, impl="sm100").clone() for _ in range(4)]
        for o in outs[1:]:
            assert torch.equal(o, outs[0]), f"non-deterministic result at T={T} dh={dh} B={B} heads={heads}"
        p = torch.einsum("bhtd,bhsd->bhts", q.float(), k.float()).softmax(-1)
        ref = torch.einsum("bhts,bhds->bthd", p, vt.float()).reshape(B, T, heads * dh)
        assert rel_err(outs[0], ref) < 2e-2, (T, dh, B, heads, scale)


def test_conv_statistics_repeatable(ops):
    """Same for the conv epilogue with fused statistics (halo and plain CTA-pair kernels, both granularities)."""
    import ctypes as C
    from advshadow_b200 import _capi as capi
    lib = capi.lib()
    torch.manual_seed(31)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for (B, H, W, cin, cout, gran) in [(2, 8, 256, 128, 128, 4), (2, 8, 256, 128, 128, 1), (3, 32, 32, 256, 256, 4),
                                       (1, 16, 128, 64, 256, 4)]:
        x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
        res = torch.randn(B, H, W, cout, device="cuda").to(torch.bfloat16)
        wp = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda") / 20, torch.bfloat16)
        bias = torch.randn(cout, device="cuda")
        parts = lib.advs_conv_sm100_stats_parts(B, H, W)
        results = []
        for _ in range(3):
            y = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device="cuda")
            part = torch.full((B, parts, cout // gran, 2), float("nan"), device="cuda")
            cp = capi.ConvParams()
            cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, cout, 1, 1
            cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = x.data_ptr(), wp.data_ptr(), cin, 9
            cp.bias, cp.out_mode, cp.y, cp.dtype, cp.residual = bias.data_ptr(), 0, y.data_ptr(), capi.BF16, res.data_ptr()
            cp.stats_partial, cp.stats_gran = part.data_ptr(), gran
            pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
            capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
            capi.call("advs_conv_sm100_launch", pb.ptr, st)
            torch.cuda.synchronize()
            results.append((y, part))
        for (y, part) in results[1:]:
            assert torch.equal(y, results[0][0]) and torch.equal(part, results[0][1])
        ref = F.conv2d(nchw(x.float()), wp.float().reshape(cout, 3, 3, cin).permute(0, 3, 1, 2), bias, padding=1) + nchw(res.float())
        assert rel_err(nchw(results[0][0]), ref) < 1e-2


def test_upsample_and_edges(ops):
    torch.manual_seed(6)
    x = torch.randn(2, 5, 7, 64, device="cuda").to(torch.bfloat16)
    y = ops.upsample_nearest2x(x)
    ref = F.interpolate(nchw(x.float()), scale_factor=2, mode="nearest")
    assert torch.equal(nchw(y.float()), ref)


def test_timestep_embedding(ops):
    t = torch.tensor([0, 1, 17, 501, 981, 999], device="cuda")
    e = ops.timestep_embedding(t, 128)
    half = 64
    freqs = torch.exp(-math.log(10000) * torch.arange(0, half, dtype=torch.float32) / half)
    args = t.cpu()[:, None].float() * freqs[None]
    ref = torch.cat([torch.cos(args), torch.sin(args)], -1)
    assert (e.cpu() - ref).abs().max().item() < 2e-6


def test_success_flags(ops):
    torch.manual_seed(7)
    logits = torch.randn(300, 37, device="cuda")
    labels = torch.randint(0, 37, (300,), device="cuda")
    labels[:100] = logits[:100].argmax(1)
    flags, counts = ops.success_flags(logits, labels)
    ref = (logits.argmax(1) != labels)
    assert torch.equal(flags.bool(), ref)
    assert counts.tolist() == [int(ref.sum()), 300]


@pytest.mark.parametrize("dtype,tol", [(torch.float32, 2e-5), (torch.bfloat16, 6e-3)])
def test_stem_conv(ops, dtype, tol):
    """Cin=3 stem (dm1:192) through advs_conv3x3_stem: fp32 NCHW in, NHWC out."""
    import ctypes as C
    from advshadow_b200 import _capi as capi
    torch.manual_seed(8)
    B, H, W, cin, cout = 2, 19, 23, 3, 128
    x = torch.randn(B, cin, H, W, device="cuda")
    w = torch.randn(cout, cin, 3, 3, device="cuda") / 5
    b = torch.randn(cout, device="cuda")
    wp = ops.pack_conv_weight(w, torch.float32)
    y = torch.empty(B, H, W, cout, dtype=dtype, device="cuda")
    capi.call("advs_conv3x3_stem", x.data_ptr(), wp.data_ptr(), b.data_ptr(), y.data_ptr(), B, H, W, cin, cout,
              capi.F32 if dtype == torch.float32 else capi.BF16, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    ref = F.conv2d(x, w, b, padding=1)
    assert rel_err(nchw(y), ref) < tol


@pytest.mark.parametrize("impl,dtype,tol", [("simt", torch.float32, 2e-5), ("sm100", torch.bfloat16, 1e-2)])
def test_head_conv_nchw_f32_output(ops, impl, dtype, tol):
    """UNet head (dm1:240-243): 128 -> 3 channels through the implicit GEMM with the fp32 NCHW epilogue;
    on the tcgen05 path the 3 output channels are zero-padded to a 64-row weight tile."""
    import ctypes as C
    from advshadow_b200 import _capi as capi
    torch.manual_seed(9)
    B, H, W, cin, cout = 2, 16, 32, 128, 3
    x = torch.randn(B, H, W, cin, device="cuda").to(dtype)
    w = (torch.randn(cout, cin, 3, 3, device="cuda") / 30).to(dtype).float()
    bias = torch.randn(cout, device="cuda")
    pad = 64 if impl == "sm100" else cout
    wp = torch.zeros(pad, 9, cin, dtype=dtype, device="cuda")
    wp[:cout] = ops.pack_conv_weight(w, dtype)
    bp = torch.zeros(pad, device="cuda")
    bp[:cout] = bias
    y = torch.full((B, cout, H, W), float("nan"), device="cuda")
    cp = capi.ConvParams()
    cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, pad, 1, 1
    cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = x.data_ptr(), wp.data_ptr(), cin, 9
    cp.bias, cp.out_mode, cp.y, cp.cout_valid = bp.data_ptr(), 2, y.data_ptr(), cout
    cp.dtype = capi.F32 if dtype == torch.float32 else capi.BF16
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    if impl == "simt":
        capi.call("advs_conv_simt", C.byref(cp), st)
    else:
        pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
        capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
        capi.call("advs_conv_sm100_launch", pb.ptr, st)
    ref = conv_ref(x, w, 1, bias)
    assert rel_err(y, ref) < tol
    # the dedicated bandwidth kernel of the C ABI computes the same thing
    y2 = torch.empty_like(y)
    w32 = ops.pack_conv_weight(w, torch.float32)
    capi.call("advs_conv3x3_head", x.data_ptr(), w32.data_ptr(), bias.data_ptr(), y2.data_ptr(), B, H, W, cin, cout,
              cp.dtype, st)
    assert rel_err(y2, ref) < tol


@pytest.mark.parametrize("case", [(2, 16, 16, 3, 128), (1, 9, 130, 3, 64), (1, 32, 32, 6, 128)])
def test_stem_on_tensor_cores(ops, case):
    """Cin = 3 stem (dm1:192) as im2col rows + a 64-wide 1x1 tcgen05 conv.  With the hi/lo split of the fp32 input
    (18*Cin <= 64) the only rounding left is the bf16 weights and the bf16 output."""
    import ctypes as C
    from advshadow_b200 import _capi as capi
    B, H, W, cin, cout = case
    torch.manual_seed(12)
    x = torch.randn(B, cin, H, W, device="cuda") * 1.5
    w = torch.randn(cout, cin, 3, 3, device="cuda") / 5
    bias = torch.randn(cout, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    col = torch.full((B, H, W, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    w64 = torch.full((cout, 1, 64), float("nan"), dtype=torch.bfloat16, device="cuda")
    capi.call("advs_stem_im2col", x.data_ptr(), col.data_ptr(), B, H, W, cin, st)
    capi.call("advs_pack_stem_weight", w.data_ptr(), w64.data_ptr(), cout, cin, st)
    assert torch.isfinite(col.float()).all() and torch.isfinite(w64.float()).all()
    y = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device="cuda")
    cp = capi.ConvParams()
    cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, cout, 1, 1
    cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = col.data_ptr(), w64.data_ptr(), 64, 1
    cp.bias, cp.out_mode, cp.y, cp.dtype = bias.data_ptr(), 0, y.data_ptr(), capi.BF16
    pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
    capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
    capi.call("advs_conv_sm100_launch", pb.ptr, st)
    wq = w.to(torch.bfloat16).float()
    ref = F.conv2d(x, wq, bias, padding=1)
    # vs the same bf16 weights: input exact to ~2^-17 when the lo part fits, else bf16 inputs as well
    assert rel_err(nchw(y), ref) < (3e-3 if 18 * cin <= 64 else 6e-3)
    if 18 * cin <= 64:
        # before the output rounding the GEMM reproduces the fp32-input convolution: check through the row sums
        got = torch.einsum("bhwk,ok->bhwo", col.float(), w64[:, 0].float()) + bias
        assert (got - ref.permute(0, 2, 3, 1)).abs().max() < 2e-4 * ref.abs().max()


@pytest.mark.parametrize("case", [(2, 16, 16, 128, 256), (3, 8, 32, 64, 192), (1, 4, 256, 128, 128)])
@pytest.mark.parametrize("gran", [1, 4])
def test_conv_epilogue_groupnorm_statistics(ops, case, gran):
    """The tcgen05 conv epilogue can emit per-tile channel sums of the tensor it stores; finalised, they
    must equal the stand-alone GroupNorm statistics kernel run on that tensor (K5 fused into K1)."""
    import ctypes as C
    from advshadow_b200 import _capi as capi
    B, H, W, cin, cout = case
    torch.manual_seed(10)
    lib = capi.lib()
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, device="cuda") / 20).to(torch.bfloat16).float()
    bias = torch.randn(cout, device="cuda")
    wp = ops.pack_conv_weight(w, torch.bfloat16)
    parts = lib.advs_conv_sm100_stats_parts(B, H, W)
    assert parts > 0
    part = torch.full((B, parts, cout // gran, 2), float("nan"), device="cuda")
    y = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device="cuda")
    cp = capi.ConvParams()
    cp.stats_gran = gran
    cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, cout, 1, 1
    cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = x.data_ptr(), wp.data_ptr(), cin, 9
    cp.bias, cp.out_mode, cp.y, cp.dtype, cp.stats_partial = bias.data_ptr(), 0, y.data_ptr(), capi.BF16, part.data_ptr()
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
    capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
    capi.call("advs_conv_sm100_launch", pb.ptr, st)
    # per-channel sums straight from the stored tensor (the epilogue sums the fp32 values before the bf16
    # rounding: zero-mean differences of relative size 2^-9 per element)
    yf = y.float()
    ref_sum = yf.sum(dim=(1, 2)).reshape(B, cout // gran, gran).sum(-1)      # gran 4: one pair per channel quad
    ref_sq = (yf * yf).sum(dim=(1, 2)).reshape(B, cout // gran, gran).sum(-1)
    got = part.sum(dim=1)
    assert torch.isfinite(part).all()
    assert (got[..., 0] - ref_sum).abs().max() <= 4e-3 * ref_sq.sqrt().max()
    assert ((got[..., 1] - ref_sq).abs() / ref_sq).max() < 2e-3
    # finalised scale/shift == the stand-alone statistics path
    g, bt = torch.randn(cout, device="cuda"), torch.randn(cout, device="cuda")
    ss_f = torch.empty(B, cout, 2, device="cuda")
    if gran == 4 and (cout // 32) % 4:
        # quads would straddle group boundaries: the finaliser refuses instead of mixing groups
        with pytest.raises(capi.AdvsError):
            capi.call("advs_groupnorm_finalize_ex", part.data_ptr(), cout, parts, gran, None, 0, 0, 1, B, H * W, 32, 1e-5,
                      g.data_ptr(), bt.data_ptr(), ss_f.data_ptr(), st)
        return
    capi.call("advs_groupnorm_finalize_ex", part.data_ptr(), cout, parts, gran, None, 0, 0, 1, B, H * W, 32, 1e-5,
              g.data_ptr(), bt.data_ptr(), ss_f.data_ptr(), st)
    ss_s = torch.empty(B, cout, 2, device="cuda")
    wsb = lib.advs_groupnorm_workspace_bytes(B, H * W, cout)
    ws = torch.empty(wsb, dtype=torch.uint8, device="cuda")
    capi.call("advs_groupnorm_stats", y.data_ptr(), cout, None, 0, B, H * W, 32, 1e-5, g.data_ptr(), bt.data_ptr(),
              ss_s.data_ptr(), ws.data_ptr(), wsb, capi.BF16, st)
    assert (ss_f - ss_s).abs().max() <= 2e-3 * ss_s.abs().max()


@pytest.mark.parametrize("case", [(2, 8, 16, 128, 128), (1, 32, 32, 64, 192), (1, 2, 128, 64, 64)])
@pytest.mark.parametrize("gran", [1, 4])
def test_upsample_conv_as_four_phase_convs(ops, case, gran):
    """Upsample (dm1:129-140): nearest 2x + conv3x3 == four 2x2 convolutions on the low-res tensor with
    pre-summed weights (advs_pack_upconv_weight), each writing one parity class of the output pixels;
    the fused GroupNorm partial statistics cover the whole high-res tensor."""
    import ctypes as C
    from advshadow_b200 import _capi as capi
    B, H, W, cin, cout = case
    torch.manual_seed(11)
    lib = capi.lib()
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, device="cuda") / 20).float()
    bias = torch.randn(cout, device="cuda")
    w4 = torch.empty(4, cout, 4, cin, dtype=torch.bfloat16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.call("advs_pack_upconv_weight", w.data_ptr(), w4.data_ptr(), cout, cin, capi.BF16, st)
    y = torch.full((B, 2 * H, 2 * W, cout), float("nan"), dtype=torch.bfloat16, device="cuda")
    parts = lib.advs_conv_sm100_stats_parts(B, H, W)
    part = torch.full((B, 4 * parts, cout // gran, 2), float("nan"), device="cuda") if parts else None
    keep = []
    for ph in range(4):
        cp = capi.ConvParams()
        cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, cout, 1, 1
        cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = x.data_ptr(), w4[ph].data_ptr(), cin, 4
        cp.bias, cp.out_mode, cp.y, cp.dtype, cp.up_phase = bias.data_ptr(), 0, y.data_ptr(), capi.BF16, ph + 1
        if part is not None:
            cp.stats_partial, cp.stats_gran = part.data_ptr(), gran
        pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
        capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
        capi.call("advs_conv_sm100_launch", pb.ptr, st)
        keep.append((cp, pb))
    ref = F.conv2d(F.interpolate(nchw(x.float()), scale_factor=2, mode="nearest"), w, bias, padding=1)
    assert rel_err(nchw(y), ref) < 1e-2
    if part is not None:
        yf = y.float()
        assert torch.isfinite(part).all()
        ref_sq = (yf * yf).sum(dim=(1, 2)).reshape(B, cout // gran, gran).sum(-1)
        assert ((part.sum(1)[..., 1] - ref_sq).abs() / ref_sq).max() < 2e-3


# ---- "wide" pre-norm storage: bf16 + int8 mantissa extension (advs_conv_params.y_lo) -------------------------
from wide_format import wide_decode, wide_encode  # noqa: E402  (tests/wide_format.py)


WIDE_CASES = [
    # B, H, W, cin, cout, taps, residual  -> kernel variant
    (2, 16, 16, 128, 128, 9, True),     # CTA pair, per-tap
    (1, 4, 256, 64, 128, 9, False),     # halo kernel, BN = 128
    (2, 3, 128, 128, 256, 9, True),     # halo kernel, BN = 256, odd tile count
    (1, 8, 8, 64, 64, 9, False),        # a single pixel tile: single-CTA kernel
    (2, 16, 16, 128, 192, 1, False),    # 1x1
]


@pytest.mark.parametrize("impl", ["sm100", "simt"])
@pytest.mark.parametrize("case", WIDE_CASES)
def test_conv_wide_prenorm_storage(ops, impl, case):
    """y stays the round-to-nearest bf16 tensor; (y, y_lo) together carry the fp32 accumulator to 2^-15 (2^-16 on average)."""
    import ctypes as C
    from advshadow_b200 import _capi as capi
    B, H, W, cin, cout, taps, with_res = case
    k = 3 if taps == 9 else 1
    torch.manual_seed(41)
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(cin * taps)).to(torch.bfloat16).float()
    bias = torch.randn(cout, device="cuda") * 3
    res = (torch.randn(B, H, W, cout, device="cuda") * 4).to(torch.bfloat16) if with_res else None
    wp = ops.pack_conv_weight(w, torch.bfloat16)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    outs = []
    for use_lo in (False, True):
        y = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device="cuda")
        lo = torch.full((B, H, W, cout), 77, dtype=torch.int8, device="cuda")
        cp = capi.ConvParams()
        cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, cout, 1, 1
        cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = x.data_ptr(), wp.data_ptr(), cin, taps
        cp.bias, cp.out_mode, cp.y, cp.dtype = bias.data_ptr(), 0, y.data_ptr(), capi.BF16
        cp.residual = res.data_ptr() if with_res else None
        cp.y_lo = lo.data_ptr() if use_lo else None
        if impl == "simt":
            capi.call("advs_conv_simt", C.byref(cp), st)
        else:
            pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
            capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
            capi.call("advs_conv_sm100_launch", pb.ptr, st)
        torch.cuda.synchronize()
        outs.append((y, lo))
    (y0, _), (y1, lo) = outs
    # y is the nearest bf16 either way; the wide encoder breaks exact ties away from zero instead of to even
    differ = (y0 != y1)
    assert differ.float().mean().item() < 1e-3 and ((y0.float() - y1.float()).abs() <= y0.float().abs() * 2 ** -7).all()
    ref = conv_ref(x, w, 1, bias)
    if with_res:
        ref = ref + nchw(res.float())
    scale = ref.abs().max().item()
    err_hi = (nchw(y1.float()) - ref).abs().max().item() / scale
    err_wide = (nchw(wide_decode(y1, lo)) - ref).abs().max().item() / scale
    print(f"{impl} {case}: bf16 alone {err_hi:.2e}, bf16 + int8 extension {err_wide:.2e} (relative to max|y|)")
    assert err_wide < 4e-5 and err_wide < err_hi / 20      # what is left is the fp32 summation order of the reference
    # and the decode never moves a value by more than half a bf16 ulp
    assert (wide_decode(y1, lo) - y1.float()).abs().max() <= (y1.float().abs().max() * 2 ** -8)


def test_upsample_conv_wide_prenorm_storage(ops):
    """the four phase convolutions of an upsample-conv scatter y_lo to the same strided pixels as y"""
    import ctypes as C
    from advshadow_b200 import _capi as capi
    B, H, W, cin, cout = 1, 8, 16, 128, 128
    torch.manual_seed(42)
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    w = (torch.randn(cout, cin, 3, 3, device="cuda") / 30).float()
    bias = torch.randn(cout, device="cuda")
    w4 = torch.empty(4, cout, 4, cin, dtype=torch.bfloat16, device="cuda")
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    capi.call("advs_pack_upconv_weight", w.data_ptr(), w4.data_ptr(), cout, cin, capi.BF16, st)
    y = torch.empty(B, 2 * H, 2 * W, cout, dtype=torch.bfloat16, device="cuda")
    lo = torch.full((B, 2 * H, 2 * W, cout), 77, dtype=torch.int8, device="cuda")
    keep = []
    for ph in range(4):
        cp = capi.ConvParams()
        cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, cout, 1, 1
        cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = x.data_ptr(), w4[ph].data_ptr(), cin, 4
        cp.bias, cp.out_mode, cp.y, cp.dtype, cp.up_phase, cp.y_lo = bias.data_ptr(), 0, y.data_ptr(), capi.BF16, ph + 1, lo.data_ptr()
        pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
        capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
        capi.call("advs_conv_sm100_launch", pb.ptr, st)
        keep.append((cp, pb))
    torch.cuda.synchronize()
    # reference with the same (bf16, pre-summed) phase weights: conv of the bf16 weights the kernel used
    ref = F.conv2d(F.interpolate(nchw(x.float()), scale_factor=2, mode="nearest"), w, bias, padding=1)
    scale = ref.abs().max().item()
    err_hi = (nchw(y.float()) - ref).abs().max().item() / scale
    dec = wide_decode(y, lo)
    # the pre-summed phase weights are rounded to bf16 AFTER summation, so the fp32 reference is only bf16-weight
    # accurate; what must hold exactly is the relation between y and its extension
    assert (dec - y.float()).abs().max() <= y.float().abs().max() * 2 ** -8
    assert err_hi < 1e-2


@pytest.mark.parametrize("c0,c1,silu,lo_mask", [(128, 0, True, (True, False)), (256, 128, True, (True, True)),
                                                (128, 256, False, (False, True)), (64, 64, False, (True, True))])
def test_groupnorm_apply_wide(ops, c0, c1, silu, lo_mask):
    """GroupNorm apply over (bf16, int8 extension) sources == apply over the fp32 tensor they encode."""
    import ctypes as C
    from advshadow_b200 import _capi as capi
    torch.manual_seed(43)
    B, H, W = 2, 12, 20
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    xs = [torch.randn(B, H, W, c, device="cuda") * 3 + 1.5 for c in (c0, c1) if c]
    enc = [wide_encode(x) for x in xs]
    # what each source encodes: fp32 to 2^-15 where the extension is passed, the bf16 value where it is not
    seen = [wide_decode(h, l) if use else h.float() for (h, l), use in zip(enc, lo_mask)]
    for x, (h, l) in zip(xs, enc):
        assert ((wide_decode(h, l) - x).abs() <= x.abs() * 2 ** -15).all()      # < 2^-16 by construction
    Ct = c0 + c1
    xin = torch.cat(seen, 3)
    g, bt = torch.randn(Ct, device="cuda"), torch.randn(Ct, device="cuda")
    xg = nchw(xin).reshape(B, 32, -1)
    mean, var = xg.mean(-1), xg.var(-1, unbiased=False)
    rstd = (var + 1e-5).rsqrt()
    cpg = Ct // 32
    sc = rstd.repeat_interleave(cpg, 1) * g[None]
    sh = bt[None] - mean.repeat_interleave(cpg, 1) * sc
    ss = torch.stack([sc, sh], -1).contiguous()
    x0h, x0l = enc[0]
    x1h, x1l = enc[1] if c1 else (None, None)
    ref = xin * sc[:, None, None, :] + sh[:, None, None, :]
    if silu:
        ref = F.silu(ref)
    for y_dtype, tdt in ((capi.BF16, torch.bfloat16), (capi.F16, torch.float16)):
        y = torch.empty(B, H, W, Ct, dtype=tdt, device="cuda")
        capi.call("advs_groupnorm_apply_wide", x0h.data_ptr(), x0l.data_ptr() if lo_mask[0] else None, c0,
                  x1h.data_ptr() if c1 else None, x1l.data_ptr() if (c1 and lo_mask[1]) else None, c1, B, H * W,
                  ss.data_ptr(), 1 if silu else 0, y.data_ptr(), y_dtype, st)
        torch.cuda.synchronize()
        refb = ref.to(tdt)
        same = (y == refb).float().mean().item()
        ulp = (y.float() - refb.float()).abs().max().item() / ref.abs().max().item()
        print(f"wide GN apply c0={c0} c1={c1} silu={silu} lo={lo_mask} -> {tdt}: {same:.5f} of the outputs bit-equal, "
              f"max diff {ulp:.2e} of max|y|")
        if tdt == torch.bfloat16:      # SiLU: tanh.approx moves a few results by one bf16 ulp
            assert same > (0.97 if silu else 0.999) and ulp < 8e-3
        else:                          # fp16 output resolves the 2^-11 error of tanh.approx itself
            assert (same > 0.999 or silu) and ulp < 1e-3


def test_ddim_step_composite_equals_two_kernels(ops):
    """The fused tail (last DDIM update + generated-shadow composite, hard and blurred mask) is bit-identical to
    advs_ddim_step followed by disk mask [+ blur] + composite."""
    import ctypes as C
    from advshadow_b200 import _capi as capi, shadow as sh
    torch.manual_seed(44)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    for (B, Cc, H, W, Cm) in [(3, 3, 40, 64, 1), (2, 3, 33, 31, 3), (2, 1, 16, 20, 1)]:      # vector and scalar paths
        x = torch.randn(B, Cc, H, W, device="cuda")
        eps = torch.randn(B, Cc, H, W, device="cuda")
        img = torch.rand(B, Cc, H, W, device="cuda")
        fm = torch.rand(B, Cm, H, W, device="cuda")
        cen = torch.rand(B, 2, device="cuda") * torch.tensor([W, H], device="cuda")
        cen[0] = torch.tensor([1.0, 0.5], device="cuda")          # a disk cut by the border: reflect-101 in the blur
        rad = torch.rand(B, device="cuda") * 10 + 4
        coef = torch.tensor([[0.6, 0.8, 0.9, 0.43, 0.0, 0, 0, 0], [0.2, 0.97, 0.999, 0.04, 0.0, 0, 0, 0]], device="cuda")
        for blur in (0, 1):
            step = torch.ones(1, dtype=torch.int32, device="cuda")
            x_ref = torch.empty_like(x)
            capi.call("advs_ddim_step", x.data_ptr(), eps.data_ptr(), None, x_ref.data_ptr(), x.numel(), coef.data_ptr(),
                      step.data_ptr(), 0, 1, st)
            m = sh.disk_mask(cen, rad, H, W)
            if blur:
                m = sh.gaussian_blur5(m)
            _, out_ref = sh.composite(img, m, fm, 0.33, adv=x_ref.clamp(0, 1))
            xs = x.clone()
            out = torch.full_like(x, float("nan"))
            capi.call("advs_ddim_step_composite", xs.data_ptr(), eps.data_ptr(), xs.data_ptr(), coef.data_ptr(),
                      step.data_ptr(), 1, 1, img.data_ptr(), cen.data_ptr(), rad.data_ptr(), fm.data_ptr(), Cm, blur,
                      out.data_ptr(), B, Cc, H, W, st)
            torch.cuda.synchronize()
            assert int(step) == 2
            assert torch.equal(xs, x_ref), "in-place state update differs from advs_ddim_step"
            assert torch.equal(out, out_ref), f"fused composite differs (blur={blur}, shape {(B, Cc, H, W, Cm)})"
            assert torch.equal(sh.composite_generated(img, x_ref, cen, rad, fm, blur=bool(blur)), out_ref)


def test_success_flags_nan_like_torch_max(ops):
    """torch.max propagates NaN (the first NaN logit is the arg-max), ASR_fast.py:117"""
    logits = torch.tensor([[0.1, float("nan"), 3.0], [2.0, 1.0, float("nan")], [0.0, 5.0, 1.0]], device="cuda")
    labels = torch.tensor([1, 0, 1], device="cuda")
    flags, counts = ops.success_flags(logits, labels)
    ref = torch.max(logits, 1)[1] != labels
    assert torch.equal(flags.bool(), ref) and counts.tolist() == [int(ref.sum()), 3]


@pytest.mark.parametrize("f16", [False, True])
@pytest.mark.parametrize("case", [(2, 16, 16, 128, 128, 9), (1, 4, 256, 64, 128, 9), (1, 8, 8, 64, 64, 1)])
def test_conv_fp16_operands(ops, case, f16):
    """advs_conv_params.operand_f16: segment 0 (activations and weights) in fp16 or bf16, the shortcut segment always
    bf16 x bf16 -- per-segment instruction descriptors on the CTA-pair, halo and single-CTA kernels.  (An MMA whose A and B formats differ is an illegal instruction on this hardware:
    tools/gpu/probe_mixed_mma.py; the engine never issues one and the planner refuses it.)"""
    import ctypes as C
    from advshadow_b200 import _capi as capi
    B, H, W, cin, cout, taps = case
    k = 3 if taps == 9 else 1
    torch.manual_seed(51)
    dt0 = torch.float16 if f16 else torch.bfloat16
    dt1 = torch.bfloat16
    x = torch.randn(B, H, W, cin, device="cuda").to(dt0)
    xs = torch.randn(B, H, W, 64, device="cuda").to(dt1)
    w = (torch.randn(cout, cin, k, k, device="cuda") / math.sqrt(cin * taps)).to(dt0).float()
    wsc = (torch.randn(cout, 64, 1, 1, device="cuda") / 8).to(dt1).float()
    wp, wscp = ops.pack_conv_weight(w, dt0), ops.pack_conv_weight(wsc, dt1)
    assert wp.dtype == dt0 and wscp.dtype == dt1
    y = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device="cuda")
    cp = capi.ConvParams()
    cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, cout, 1, 2
    cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = x.data_ptr(), wp.data_ptr(), cin, taps
    cp.seg[1].x, cp.seg[1].w, cp.seg[1].C, cp.seg[1].taps = xs.data_ptr(), wscp.data_ptr(), 64, 1
    cp.out_mode, cp.y, cp.dtype = 0, y.data_ptr(), capi.BF16
    cp.operand_f16 = 0b0101 if f16 else 0
    pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
    capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
    capi.call("advs_conv_sm100_launch", pb.ptr, C.c_void_p(torch.cuda.current_stream().cuda_stream))
    torch.cuda.synchronize()
    ref = conv_ref(x, w, 1) + conv_ref(xs, wsc, 1)
    err = rel_err(nchw(y), ref)
    print(f"conv {case} segment 0 {dt0}, shortcut {dt1}: rel err {err:.2e}")
    assert err < 6e-3
    for bad in (0b0001, 0b0100, 0b1010, 0b1111):      # mixed A / B formats, or fp16 shortcut segments
        cp.operand_f16 = bad
        with pytest.raises(capi.AdvsError):
            capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)


def attn_ref(q, k, vt):
    B, heads, T, dh = q.shape
    p = torch.einsum("bhtd,bhsd->bhts", q.float(), k.float()).softmax(-1)
    return torch.einsum("bhts,bhds->bthd", p, vt.float()).reshape(B, T, heads * dh)


@pytest.mark.parametrize("T", [8, 64, 200, 784, 3136])       # 28x28 and 56x56 maps of a 224x224 input, and tiny maps
@pytest.mark.parametrize("dh", [64, 128, 256])
def test_attention_partial_last_key_block(ops, T, dh):
    """T % 128 != 0 stays on the tcgen05 kernel: out-of-range keys of the last block are masked to -inf, query rows
    past T are not stored.  Several (image, head) pairs so that the rows a partial block over-reads belong to the
    NEXT head (finite garbage) for all but the last one (TMA zero fill)."""
    torch.manual_seed(61)
    B, heads = 2, 3
    q = (torch.randn(B, heads, T, dh, device="cuda") * 0.8).to(torch.bfloat16)
    k = (torch.randn(B, heads, T, dh, device="cuda") * 0.8).to(torch.bfloat16)
    vt = torch.randn(B, heads, dh, T, device="cuda").to(torch.bfloat16)
    guard = torch.full((B, T + 16, heads * dh), 7.0, dtype=torch.bfloat16, device="cuda")      # catches rows stored past T
    o = ops.attention(q, k, vt, impl="sm100")
    assert torch.isfinite(o.float()).all()
    assert rel_err(o, attn_ref(q, k, vt)) < 1.5e-2
    assert torch.equal(o, ops.attention(q, k, vt, impl="sm100"))
    del guard


@pytest.mark.parametrize("dhv", [16, 32])
@pytest.mark.parametrize("T", [64, 256, 1000, 4096])
def test_attention_padded_head_dim(ops, dhv, T):
    """IDDM's 4-head nn.MultiheadAttention has head dims 16 / 32 (model/modules/attention.py:27): run as dh = 64 with
    zero-padded q / k / v^T; the output holds the dh_valid real columns of every head, packed."""
    torch.manual_seed(62)
    B, heads = 2, 4
    q = torch.zeros(B, heads, T, 64, dtype=torch.bfloat16, device="cuda")
    k = torch.zeros_like(q)
    vt = torch.zeros(B, heads, 64, T, dtype=torch.bfloat16, device="cuda")
    q[..., :dhv] = (torch.randn(B, heads, T, dhv, device="cuda") * 0.9).to(torch.bfloat16)
    k[..., :dhv] = (torch.randn(B, heads, T, dhv, device="cuda") * 0.9).to(torch.bfloat16)
    vt[:, :, :dhv] = torch.randn(B, heads, dhv, T, device="cuda").to(torch.bfloat16)
    o = ops.attention(q, k, vt, impl="sm100", dh_valid=dhv)
    assert o.shape == (B, T, heads * dhv)
    ref = attn_ref(q[..., :dhv], k[..., :dhv], vt[:, :, :dhv])
    assert rel_err(o, ref) < 1.5e-2


@pytest.mark.parametrize("c,heads,pad", [(64, 4, 64), (128, 4, 64), (128, 4, 0)])
def test_conv_qkv_split_small_heads(ops, c, heads, pad):
    """qkv epilogue for head dims 16 / 32, written into 64-wide zero-padded q / k / v^T (advs_conv_params.qkv_dh_pad)."""
    torch.manual_seed(63)
    B, H, W = 2, 16, 8
    dh = c // heads
    x = torch.randn(B, H, W, c, device="cuda").to(torch.bfloat16)
    w = (torch.randn(3 * c, c, 1, 1, device="cuda") / 8).to(torch.bfloat16).float()
    bias = torch.randn(3 * c, device="cuda")
    q, k, vt = ops.conv([(x, ops.pack_conv_weight(w, torch.bfloat16))], B, H, W, 3 * c, bias=bias, qkv_heads=heads, impl="sm100",
                        qkv_dh_pad=pad)
    ref = conv_ref(x, w, 1, bias).reshape(B, heads, 3, dh, H * W)
    s_ = 1 / math.sqrt(math.sqrt(dh))
    dhs = max(dh, pad)
    assert q.shape == (B, heads, H * W, dhs) and vt.shape == (B, heads, dhs, H * W)
    assert rel_err(q[..., :dh], (ref[:, :, 0] * s_).transpose(2, 3)) < 1e-2
    assert rel_err(k[..., :dh], (ref[:, :, 1] * s_).transpose(2, 3)) < 1e-2
    assert rel_err(vt[:, :, :dh], ref[:, :, 2]) < 1e-2
    if dhs > dh:
        assert float(q[..., dh:].abs().max()) == 0 and float(k[..., dh:].abs().max()) == 0 and float(vt[:, :, dh:].abs().max()) == 0
