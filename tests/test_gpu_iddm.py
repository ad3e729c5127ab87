"""IDDM class-conditional UNet + CFG DDIM on the B200 vs the reference's own
model/networks/unet.py + model/samples/ddim.py outputs (tests/golden/iddm.pt)."""
import ctypes as C

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def iddm():
    import advshadow_b200
    from advshadow_b200 import iddm as m
    assert torch.cuda.is_available()
    return m


def st():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def build(iddm, size):
    torch.manual_seed(0)
    net = iddm.UNet(num_classes=37, image_size=size)
    chk = float(sum(p.detach().double().abs().sum() for p in net.parameters()))
    return net.eval().cuda(), chk


@pytest.mark.parametrize("size", [32, 64])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_iddm_unet_forward(iddm, golden, size, precision, tol):
    g = golden("iddm.pt")[f"fwd_{size}"]
    net, chk = build(iddm, size)
    assert abs(chk - g["checksum"]) <= 1e-6 * chk
    net.set_precision(precision)
    x, t, y = g["x"].cuda(), g["t"].cuda(), g["y"].cuda()
    if precision == "bf16":      # every attention block (head dims 16 / 32 / 64, T down to 16) is on the tcgen05 kernel
        eng = net.engine(x.shape[0])
        assert eng.attn_kinds == ["sm100"] * 6, eng.attn_kinds
        names = [n for (_, _, n) in eng.L]
        assert names.count("advs_conv_simt") == 0 and names.count("advs_conv_sm100_launch") >= 40
    e_c = net(x, t, y).cpu()
    e_u = net(x, t).cpu()
    err_c = (e_c - g["eps_cond"]).abs().max().item()
    err_u = (e_u - g["eps_uncond"]).abs().max().item()
    print(f"IDDM UNet {size}x{size} {precision}: max|eps err| cond {err_c:.3e} uncond {err_u:.3e} "
          f"(|eps|max {g['eps_cond'].abs().max():.2f})")
    assert err_c <= tol and err_u <= tol
    net.release_engines()


def test_iddm_cfg_ddim_sample(iddm, golden):
    """DDIMDiffusion.sample(model, n, labels, cfg_scale=3), 5 steps: every denoiser call of the reference run
    is replayed teacher-forced (fp32 mode), and the free-running uint8 result is compared."""
    g = golden("iddm.pt")["sample"]
    net, _ = build(iddm, 32)
    net.set_precision("fp32")
    eng = net.engine(2)
    worst = 0.0
    for i in range(g["trace_x"].shape[0]):
        y = g["labels"].cuda() if bool(g["trace_has_y"][i]) else None
        e = eng.forward(g["trace_x"][i].cuda(), g["trace_t"][i].cuda(), y).cpu()
        worst = max(worst, (e - g["trace_eps"][i]).abs().max().item())
    print(f"IDDM CFG-DDIM teacher-forced: worst max|eps err| over {g['trace_x'].shape[0]} denoiser calls = {worst:.3e}")
    assert worst <= 1e-4
    ddim = iddm.DDIMDiffusion(noise_steps=1000, sample_steps=g["sample_steps"], img_size=32, device="cpu")
    torch.manual_seed(5)     # same global CPU generator draw as the reference (ddim.py:61)
    img = ddim.sample(net, 2, labels=g["labels"].cuda(), cfg_scale=g["cfg_scale"])
    assert img.dtype == torch.uint8 and tuple(img.shape) == (2, 3, 32, 32) and img.is_cuda
    assert net.training            # the reference leaves the model in train() mode (ddim.py:95)
    diff = (img.cpu().int() - g["image"].int()).abs()
    diff = torch.minimum(diff, 256 - diff)     # the unclamped cast wraps (ddim.py:97-99)
    frac_exact = (diff == 0).float().mean().item()
    print(f"IDDM CFG-DDIM 5 steps: uint8 image identical on {100 * frac_exact:.2f}% of values, max LSB diff {int(diff.max())}")
    assert int(diff.max()) <= 1 and frac_exact >= 0.995
    net.release_engines()


def test_iddm_cfg_one_forward_equals_two(iddm, golden):
    """Classifier-free guidance as ONE forward over 2n rows (labels on the first n) == the reference's two forwards."""
    g = golden("iddm.pt")["sample"]
    net, _ = build(iddm, 32)
    for precision in ("fp32", "bf16"):
        net.set_precision(precision)
        x, t, y = g["trace_x"][0].cuda(), g["trace_t"][0].cuda(), g["labels"].cuda()
        n = x.shape[0]
        e_c, e_u = net.engine(n).forward(x, t, y), net.engine(n).forward(x, t)
        eng2 = net.engine(2 * n)
        eng2.x.copy_(torch.cat([x, x])); eng2.t.copy_(torch.cat([t, t])); eng2.y[:n].copy_(y)
        eng2.run(True, n_labeled=n)
        torch.cuda.synchronize()
        assert torch.equal(eng2.eps[:n], e_c) and torch.equal(eng2.eps[n:], e_u), precision
    # and the 16-bit sampler end to end: 5 CFG steps stay close to the reference's fp32 image
    ddim = iddm.DDIMDiffusion(noise_steps=1000, sample_steps=g["sample_steps"], img_size=32, device="cpu")
    torch.manual_seed(5)
    img = ddim.sample(net, 2, labels=g["labels"].cuda(), cfg_scale=g["cfg_scale"])
    diff = (img.cpu().int() - g["image"].int()).abs()
    diff = torch.minimum(diff, 256 - diff)
    print(f"IDDM CFG-DDIM 5 steps, 16-bit mode: uint8 image identical on {100 * (diff == 0).float().mean().item():.2f}% of values, "
          f"max LSB diff {int(diff.max())}, mean {diff.float().mean().item():.3f}")
    assert diff.float().mean().item() < 10.0      # free-running, guidance scale 3: a sanity bound, not a parity claim
    net.release_engines()


def test_iddm_bandwidth_kernels(iddm):
    from advshadow_b200 import _capi as capi
    torch.manual_seed(1)
    B, H, W, Cc = 2, 6, 10, 64
    x = torch.randn(B, H, W, Cc, device="cuda")
    nchw = lambda t: t.permute(0, 3, 1, 2)
    # MaxPool2d(2)
    y = torch.empty(B, H // 2, W // 2, Cc, device="cuda")
    capi.call("advs_maxpool2x2", x.data_ptr(), y.data_ptr(), B, H, W, Cc, capi.F32, st())
    assert torch.equal(nchw(y), F.max_pool2d(nchw(x), 2))
    # bilinear x2 align_corners into a concat slice + skip copy
    skip = torch.randn(B, 2 * H, 2 * W, 32, device="cuda")
    cat = torch.full((B, 2 * H, 2 * W, 32 + Cc), float("nan"), device="cuda")
    capi.call("advs_copy_channels", skip.data_ptr(), cat.data_ptr(), B * 4 * H * W, 32, 32 + Cc, 0, capi.F32, st())
    capi.call("advs_upsample_bilinear2x", x.data_ptr(), cat.data_ptr(), B, H, W, Cc, 32 + Cc, 32, capi.F32, st())
    ref = torch.cat([nchw(skip), F.interpolate(nchw(x), scale_factor=2, mode="bilinear", align_corners=True)], 1)
    assert (nchw(cat) - ref).abs().max().item() < 1e-5
    # LayerNorm over channels
    g, b = torch.randn(Cc, device="cuda"), torch.randn(Cc, device="cuda")
    y = torch.empty_like(x)
    capi.call("advs_layernorm", x.data_ptr(), g.data_ptr(), b.data_ptr(), y.data_ptr(), B * H * W, Cc, 1e-5, capi.F32, st())
    assert (y - F.layer_norm(x, [Cc], g, b, 1e-5)).abs().max().item() < 1e-5
    # GroupNorm(1, C) apply with residual + GELU + embedding
    ss = torch.randn(B, Cc, 2, device="cuda")
    res, emb = torch.randn_like(x), torch.randn(B, Cc, device="cuda")
    capi.call("advs_groupnorm_apply_ex", x.data_ptr(), B, H * W, Cc, ss.data_ptr(), res.data_ptr(), emb.data_ptr(), Cc, 2,
              y.data_ptr(), capi.F32, st())
    ref = F.gelu(x * ss[:, None, None, :, 0] + ss[:, None, None, :, 1] + res) + emb[:, None, None, :]
    assert (y - ref).abs().max().item() < 1e-5
    # torch.lerp, both branches, bit-exact; unclamped uint8 cast
    u, c = torch.randn(1000, device="cuda"), torch.randn(1000, device="cuda")
    out = torch.empty_like(u)
    for w in (3.0, 0.3):
        capi.call("advs_cfg_lerp", u.data_ptr(), c.data_ptr(), C.c_float(w), out.data_ptr(), 1000, st())
        assert torch.equal(out.cpu(), torch.lerp(u.cpu(), c.cpu(), w))   # the oracle is the reference on the CPU
    v = torch.linspace(-1.0, 1.0, 1001, device="cuda")
    o8 = torch.empty(1001, dtype=torch.uint8, device="cuda")
    capi.call("advs_to_uint8", v.data_ptr(), o8.data_ptr(), 1001, st())
    assert torch.equal(o8.cpu(), (((v.cpu() + 1) * 0.5) * 255).type(torch.uint8))
