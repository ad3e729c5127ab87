"""Parity of the CUDA path with the reference, through the public drop-in API, on the B200.
Oracle = tests/golden/*.pt, minted from the unmodified reference modules by oracle/make_golden.py
(the reference ships no tests/golden vectors of its own).  Tolerances are BASELINE.json's:
<=1e-4 max-abs in fp32 mode, <=2e-2 in bf16 mode; masks / composites / DDIM update bit-exact."""
import os

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL_FP32 = 1e-4
TOL_BF16 = 2e-2


def checksum(model):
    return float(sum(p.detach().double().abs().sum() for p in model.parameters()))


@pytest.fixture(scope="module")
def pkg():
    import advshadow_b200
    from advshadow_b200 import diff_model, diff_model2, shadow
    assert torch.cuda.is_available()
    return dict(dm1=diff_model, dm2=diff_model2, shadow=shadow)


_models = {}


def get_model(pkg, flavour):
    if flavour not in _models:
        torch.manual_seed(0)
        if flavour == "dm1":
            m = pkg["dm1"].UNetModel()
        elif flavour == "main":
            m = pkg["dm1"].UNetModel(channel_mult=(1, 2, 2, 2), attention_resolutions=(2,), dropout=0.1)
        else:
            m = pkg["dm2"].UNetModel()
        _models[flavour] = (m.eval().cuda(), checksum(m))
    return _models[flavour]


FORWARD_CASES = [("dm1", "dm1_32", "dm1_checksum"), ("dm1", "dm1_64", "dm1_checksum"), ("main", "main_32", None),
                 ("dm2", "dm2_64", "dm2_checksum"), ("dm2", "dm2_128", "dm2_checksum")]


@pytest.mark.parametrize("flavour,key,ck", FORWARD_CASES)
@pytest.mark.parametrize("precision,conv,attn,fuse,tol",
                         [("fp32", "simt", "simt", True, TOL_FP32), ("bf16", "simt", "simt", True, TOL_BF16),
                          ("bf16", "sm100", "simt", True, TOL_BF16), ("bf16", "sm100", "sm100", False, TOL_BF16),
                          ("bf16", "sm100", "sm100", "no_upfuse", TOL_BF16),
                          ("bf16", "sm100", "sm100", "narrow", TOL_BF16),
                          ("bf16", "sm100", "sm100", "pure_bf16", TOL_BF16),
                          ("bf16", "sm100", "sm100", True, TOL_BF16)])
def test_unet_forward_matches_reference(pkg, golden, flavour, key, ck, precision, conv, attn, fuse, tol):
    g = golden("forwards.pt")
    case = g[key]
    model, chk = get_model(pkg, flavour)
    want = g[ck] if ck else case["checksum"]
    assert abs(chk - want) <= 1e-6 * want, "seeded weights differ from the fixture"
    x, t = case["x"].cuda(), case["t"].cuda()
    # variants: "narrow" = no int8 mantissa extension on pre-norm tensors; "pure_bf16" = every GEMM operand bf16
    eng = model.engine(x.shape[0], x.shape[2], x.shape[3], precision=precision, conv_impl=conv, attn_impl=attn,
                       fuse_gn_stats=bool(fuse), fuse_upsample=(fuse != "no_upfuse"),
                       wide_prenorm=(0 if fuse == "narrow" else 2),
                       gemm_operands=("bf16" if fuse == "pure_bf16" else "fp16"))
    eps = eng.forward(x, t)
    torch.cuda.synchronize()
    err = (eps.cpu() - case["eps"]).abs().max().item()
    print(f"{key} {precision}/{conv}/{attn}/fused_gn_stats={fuse}: max|eps err| = {err:.3e} "
          f"(|eps|max {case['eps'].abs().max():.2f})")
    assert err <= tol
    model.release_engines()


def test_config1_ddim10_fp32_final_image(pkg, golden):
    """BASELINE.json configs[0]: free-running 10-step DDIM at 64x64 against the reference's output."""
    g = golden("config1.pt")
    model, chk = get_model(pkg, "dm1")
    assert abs(chk - g["weight_checksum"]) <= 1e-6 * g["weight_checksum"]
    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000)
    model.set_precision("fp32")
    img = gd.ddim_sample(model, 64, batch_size=1, channels=3, ddim_timesteps=10, x_T=g["x_T"])
    assert isinstance(img, np.ndarray) and img.shape == (1, 3, 64, 64) and img.dtype == np.float32
    err = np.abs(img - g["final"].numpy()).max()
    print(f"config1 fp32 final image max|err| = {err:.3e}")
    assert err <= TOL_FP32
    # graph replay and eager launches must agree bit for bit
    gd.use_cuda_graph = False
    img2 = gd.ddim_sample(model, 64, batch_size=1, channels=3, ddim_timesteps=10, x_T=g["x_T"])
    assert np.array_equal(img, img2)
    model.set_precision("bf16")
    model.release_engines()


@pytest.mark.parametrize("precision,tol", [("fp32", TOL_FP32), ("bf16", TOL_BF16)])
def test_config1_teacher_forced_steps(pkg, golden, precision, tol):
    """Feed the oracle's x_t into the CUDA path at every step: eps within tolerance, and the DDIM
    update applied to the oracle's eps reproduces the oracle's x_{t-1} bit for bit."""
    from advshadow_b200 import _capi as capi
    import ctypes as C
    g = golden("config1.pt")
    model, _ = get_model(pkg, "dm1")
    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000)
    n = g["trace_x"].shape[0]
    eng = model.engine(1, 64, 64, precision=precision)
    seq, prev = pkg["dm1"].ddim_timestep_tables(1000, n, "uniform")
    coef = gd.ddim_coefficients(seq, prev, n, 0.0).cuda()
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    worst, errs = 0.0, []
    for i in range(n):
        x = g["trace_x"][i].cuda()
        eps = eng.forward(x, g["trace_t"][i].cuda())
        errs.append((eps.cpu() - g["trace_eps"][i]).abs().max().item())
        worst = max(errs)
        want_next = g["trace_x"][i + 1] if i + 1 < n else g["final"]
        e_or = g["trace_eps"][i].cuda().contiguous()
        out = torch.empty_like(x)
        step.fill_(i)
        capi.call("advs_ddim_step", x.data_ptr(), e_or.data_ptr(), None, out.data_ptr(), x.numel(), coef.data_ptr(),
                  step.data_ptr(), 0, 1, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert torch.equal(out.cpu(), want_next), f"DDIM update differs from the reference at step {i}"
    errs.sort()
    print(f"teacher-forced {precision}: max|eps err| per step: median {errs[n // 2]:.3e}, worst {worst:.3e}")
    assert worst <= tol                     # every step, both modes (BASELINE.json: 1e-4 fp32, 2e-2 bf16)
    model.release_engines()


def test_shadow_masks_and_composites_bit_exact(pkg, golden):
    sh = pkg["shadow"]
    for c in golden("shadow.pt"):
        img, fm, adv = c["img"].cuda(), c["fm"].cuda(), c["adv"].cuda()
        H, W = img.shape[1:]
        m = sh.disk_mask(c["center"][None].cuda(), c["radius"][None].cuda(), H, W)
        assert torch.equal(m[0].cpu(), c["mask"])
        shadowed, out = sh.composite(img[None], m, fm[None], c["intensity"], adv=adv)
        assert torch.equal(shadowed[0].cpu(), c["shadowed"])
        assert torch.equal(out.cpu(), c["out"])


def test_config1_composite_bit_exact(pkg, golden):
    g = golden("config1.pt")
    sh = pkg["shadow"]
    gd2 = pkg["dm2"].GaussianDiffusion()
    m = gd2.create_shadow_mask((3, 64, 64), g["center"].cuda(), g["radius"].cuda(), "cuda")
    assert torch.equal(m.cpu(), g["shadow_mask"])
    out = sh.composite_generated(g["clean"][None].cuda(), g["final"].cuda(), g["center"][None].cuda(),
                                 g["radius"][None].cuda(), g["feature_mask"][None].cuda())
    assert torch.equal(out.cpu(), g["composite"])


def test_gaussian_blur_matches_cv2_table(pkg):
    """cv2.GaussianBlur(mask,(5,5),0) = separable [1,4,6,4,1]/16 with BORDER_REFLECT_101 (ts:147-153);
    on {0,1} masks every partial sum is a dyadic rational, so any evaluation order is exact."""
    sh = pkg["shadow"]
    torch.manual_seed(0)
    m = (torch.rand(3, 37, 53, device="cuda") > 0.6).float()
    out = sh.gaussian_blur5(m)
    k = torch.tensor([1., 4., 6., 4., 1.], device="cuda") / 16
    p = F.pad(m[:, None], (2, 2, 2, 2), mode="reflect")
    ref = F.conv2d(F.conv2d(p, k.view(1, 1, 1, 5)), k.view(1, 1, 5, 1))[:, 0]
    assert torch.equal(out, ref)
    try:
        import cv2
        ref2 = cv2.GaussianBlur(m[0].cpu().numpy(), (5, 5), 0)
        assert np.array_equal(out[0].cpu().numpy(), ref2)
    except ImportError:
        pass


@pytest.mark.parametrize("eta", [0.0, 0.5, 1.0])
def test_ddim_update_bit_exact_vs_reference_formula(pkg, eta):
    """dm1:457-472 evaluated with torch ops on the GPU vs the fused kernel, same inputs."""
    from advshadow_b200 import _capi as capi
    import ctypes as C
    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000)
    seq, prev = pkg["dm1"].ddim_timestep_tables(1000, 50, "uniform")
    coef = gd.ddim_coefficients(seq, prev, 50, eta).cuda()
    torch.manual_seed(0)
    x, e, z = (torch.randn(2, 3, 33, 31, device="cuda") for _ in range(3))
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    for row, i in ((0, 49), (20, 29), (49, 0)):
        t = torch.full((2,), int(seq[i]), device="cuda")
        pt = torch.full((2,), int(prev[i]), device="cuda")
        a_t = gd._extract(gd.alphas_cumprod, t, x.shape)
        a_p = gd._extract(gd.alphas_cumprod, pt, x.shape)
        x0 = torch.clamp((x - torch.sqrt(1. - a_t) * e) / torch.sqrt(a_t), -1., 1.)
        sig = eta * torch.sqrt((1 - a_p) / (1 - a_t) * (1 - a_t / a_p))
        ref = torch.sqrt(a_p) * x0 + torch.sqrt(1 - a_p - sig ** 2) * e + sig * z
        out = torch.empty_like(x)
        step.fill_(row)
        capi.call("advs_ddim_step", x.data_ptr(), e.data_ptr(), z.data_ptr(), out.data_ptr(), x.numel(), coef.data_ptr(),
                  step.data_ptr(), 0, 1, C.c_void_p(torch.cuda.current_stream().cuda_stream))
        assert torch.equal(out, ref)


def test_foreign_callable_model_uses_fused_update(pkg, golden):
    """GaussianDiffusion accepts any callable (x,t)->eps with parameters (dm1:442,454)."""
    class Tiny(torch.nn.Module):
        def __init__(self):
            super().__init__()
            self.w = torch.nn.Parameter(torch.tensor(0.1))

        def forward(self, x, t):
            return x * self.w

    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000)
    m = Tiny().cuda()
    torch.manual_seed(0)
    xT = torch.randn(2, 3, 8, 8)
    out = gd.ddim_sample(m, 8, batch_size=2, channels=3, ddim_timesteps=10, x_T=xT)
    x = xT.cuda()
    seq, prev = pkg["dm1"].ddim_timestep_tables(1000, 10, "uniform")
    for i in reversed(range(10)):
        a_t = gd.alphas_cumprod[int(seq[i])].float().cuda()
        a_p = gd.alphas_cumprod[int(prev[i])].float().cuda()
        e = x * m.w.detach()
        x0 = torch.clamp((x - torch.sqrt(1. - a_t) * e) / torch.sqrt(a_t), -1., 1.)
        x = torch.sqrt(a_p) * x0 + torch.sqrt(1 - a_p) * e
    assert np.allclose(out, x.cpu().numpy(), atol=1e-6)


def test_ddpm_sample_matches_reference(pkg, golden):
    """§8(f) row 1: the DDPM ancestral loop main.py:124 calls (dm1:356-413), noise injected, fp32 mode.
    Return type: list of T numpy arrays like the reference; every step is compared."""
    g = golden("stochastic.pt")["ddpm"]
    model, _ = get_model(pkg, "dm1")
    model.set_precision("fp32")
    gd = pkg["dm1"].GaussianDiffusion(timesteps=g["T"])
    imgs = gd.sample(model, 32, batch_size=1, channels=3, x_T=g["x_T"], noise=list(g["noise"]))
    assert isinstance(imgs, list) and len(imgs) == g["T"] and imgs[-1].shape == (1, 3, 32, 32)
    err = max(float(np.abs(imgs[i] - g["traj"][i].numpy()).max()) for i in range(g["T"]))
    print(f"DDPM T={g['T']} trajectory max|err| = {err:.3e}")
    assert err <= TOL_FP32
    last = gd.sample(model, 32, batch_size=1, channels=3, x_T=g["x_T"], noise=list(g["noise"]), keep="last")
    assert np.array_equal(last[-1], imgs[-1])
    with pytest.raises(IndexError):
        last[0]
    model.set_precision("bf16")
    model.release_engines()


def test_ddim_eta_matches_reference(pkg, golden):
    g = golden("stochastic.pt")["ddim_eta"]
    model, _ = get_model(pkg, "dm1")
    model.set_precision("fp32")
    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000)
    img = gd.ddim_sample(model, 32, batch_size=2, channels=3, ddim_timesteps=g["n"], ddim_eta=g["eta"],
                         x_T=g["x_T"], noise=list(g["noise"]))
    err = float(np.abs(img - g["final"].numpy()).max())
    print(f"DDIM eta={g['eta']} final max|err| = {err:.3e}")
    assert err <= TOL_FP32
    model.set_precision("bf16")
    model.release_engines()


def test_attack_loop_single_rank(pkg):
    """Sampler -> composite -> PyTorch victim -> success flags, K=2 candidates per image (world size 1)."""
    from advshadow_b200.attack import AttackLoop
    from advshadow_b200.sampler import ShadowSampler
    model, _ = get_model(pkg, "dm1")
    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000)
    B_img, K, S = 3, 2, 32
    sampler = ShadowSampler(model, gd, B_img * K, S, ddim_timesteps=4)
    torch.manual_seed(0)
    victim = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, padding=1), torch.nn.AdaptiveAvgPool2d(1), torch.nn.Flatten(),
                                 torch.nn.Linear(8, 37)).cuda().eval()
    g = torch.Generator().manual_seed(1)
    x_T = torch.randn(B_img * K, 3, S, S, generator=g)
    clean = torch.rand(B_img, 3, S, S, generator=g).repeat_interleave(K, 0)
    fmask = torch.ones(B_img * K, 1, S, S)
    centers = torch.full((B_img * K, 2), S / 2.0)
    radii = torch.tensor([6.0, 12.0] * B_img)
    labels = torch.randint(0, 37, (B_img,), generator=g).repeat_interleave(K, 0)
    loop = AttackLoop(sampler, victim, candidates=K, victim_size=S)
    res = loop.step(x_T, clean, fmask, centers, radii, labels)
    with torch.no_grad():
        ref = (victim(res["shadowed"]).argmax(1).cpu() != labels).view(B_img, K).any(1)
    assert torch.equal(res["flags"].bool().cpu(), ref)
    assert res["total"] == B_img and res["successes"] == int(ref.sum())
    # the composite only touches the disk: outside it the clean image is returned untouched
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    outside = ((xx - S / 2) ** 2 + (yy - S / 2) ** 2).sqrt() > 12.0
    assert torch.equal(res["shadowed"].cpu()[:, :, outside], clean[:, :, outside])
    model.release_engines()


def test_attack_decisions_agree_with_reference_sampler(pkg):
    """North-star end-to-end criterion: attack-success decisions on images sampled by the 16-bit CUDA path vs by the
    reference algorithm from the same noise must agree on >= 99 % of the images -- all images, no margin carve-out.
      * 1024 images, dm1 UNet, 32x32, free-running DDIM-10, dm2-flavour composite;
      * victim: ResNet-50-shaped (torchvision resnet50, 37 classes, random init, eval) on Resize((224,224)) inputs
        in [0,1] without normalisation (ASR_fast.py:90-97).  A random-init ResNet-50 predicts one class for every
        input; its fc bias is re-centred on the clean images (bias -= mean clean logit) so decisions spread over
        the classes.  "True" label = the victim's decision on the clean image; success = the shadow flips it;
      * reference = oracle/torch_port.py (plain PyTorch fp32, TF32 off) run on the GPU for all 1024 images, itself
        pinned here against the same oracle on the CPU for the first 32 (the CPU needs minutes for 1024)."""
    import torchvision
    from oracle import torch_port as P
    from advshadow_b200 import ops
    from advshadow_b200.attack import victim_preprocess
    from advshadow_b200.sampler import ShadowSampler
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    model, _ = get_model(pkg, "dm1")
    model.set_precision("bf16")
    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000)
    N, chunk, S, n = 1024, 256, 32, 10
    g = torch.Generator().manual_seed(21)
    x_T = torch.randn(N, 3, S, S, generator=g)
    clean = torch.rand(N, 3, S, S, generator=g)
    fmask = (torch.rand(N, 1, S, S, generator=g) > 0.3).float()
    centers = torch.rand(N, 2, generator=g) * S
    radii = torch.rand(N, generator=g) * 8 + 6
    torch.manual_seed(3)
    victim = torchvision.models.resnet50(num_classes=37).eval().cuda()

    @torch.no_grad()
    def logits_of(imgs):
        return torch.cat([victim(victim_preprocess(imgs[i:i + 128].cuda(), 224)).float() for i in range(0, len(imgs), 128)])

    with torch.no_grad():
        victim.fc.bias -= logits_of(clean).mean(0)
    labels = logits_of(clean).argmax(1)
    # ---- CUDA path ----
    sampler = ShadowSampler(model, gd, chunk, S, ddim_timesteps=n)
    out = torch.cat([sampler(x_T[i:i + chunk], clean[i:i + chunk], fmask[i:i + chunk], centers[i:i + chunk],
                             radii[i:i + chunk]) for i in range(0, N, chunk)])
    logits_gpu = logits_of(out)
    flags_gpu, counts = ops.success_flags(logits_gpu, labels)
    # ---- reference algorithm (fp32) ----
    acp = P.cosine_alphas_cumprod()
    params = {k: v.detach() for k, v in model.state_dict().items()}              # on the GPU

    def reference_images(p, dev, lo, hi):
        with torch.no_grad():
            x0 = P.ddim_sample(p, P.DM1_CFG, acp, x_T[lo:hi].to(dev), n)
            return torch.stack([P.apply_shadow(clean[i].to(dev), centers[i].to(dev), radii[i].to(dev), fmask[i].to(dev), 0.33,
                                               perturb=lambda s_, j=i - lo: x0[j].clamp(0, 1)[None])[0][0]
                                for i in range(lo, hi)])

    ref_imgs = torch.cat([reference_images(params, "cuda", i, i + chunk) for i in range(0, N, chunk)]).cpu()
    cpu_imgs = reference_images({k: v.cpu() for k, v in params.items()}, "cpu", 0, 32)
    pin = (ref_imgs[:32] - cpu_imgs).abs().max().item()
    logits_ref = logits_of(ref_imgs)
    pred_ref, pred_gpu = logits_ref.argmax(1), logits_gpu.argmax(1)
    flags_ref = pred_ref != labels
    assert pin <= 2e-3 and torch.equal(logits_of(cpu_imgs).argmax(1), pred_ref[:32]), "oracle on GPU drifted from the oracle on CPU"
    agree = (flags_gpu.bool() == flags_ref).float().mean().item()
    same_class = (pred_gpu == pred_ref).float().mean().item()
    img_err = (out - ref_imgs).abs()
    top2 = logits_ref.topk(2, dim=1).values
    margin = (top2[:, 0] - top2[:, 1]).cpu()
    miss = (pred_gpu != pred_ref).nonzero().flatten().cpu()
    print(f"16-bit DDIM-{n} + composite vs reference, {N} images, ResNet-50-shaped victim at 224x224: decision agreement = {agree:.4f}, "
          f"same predicted class = {same_class:.4f}, successes {int(counts[0])}/{int(counts[1])} (reference {int(flags_ref.sum())}), "
          f"{pred_ref.unique().numel()} distinct predicted classes; image err max {img_err.max():.3e} mean {img_err.mean():.3e}; "
          f"oracle GPU-vs-CPU pin {pin:.2e}; reference margins of the images whose class differs: "
          f"{[round(margin[i].item(), 4) for i in miss]} (median margin {margin.median().item():.4f})")
    assert pred_ref.unique().numel() >= 8 and 0.05 * N < int(flags_ref.sum()) < 0.95 * N, "degenerate victim: the check would be vacuous"
    assert agree >= 0.99 and same_class >= 0.99
    model.release_engines()


def test_shadow_sampler_multi_stream_matches_single_stream(pkg):
    """ShadowSampler(streams=2) runs two half-batches on two CUDA streams (HBM-bound kernels of one overlap the
    tensor-bound kernels of the other); images are independent trajectories, so results must agree."""
    from advshadow_b200.sampler import ShadowSampler
    model, _ = get_model(pkg, "dm1")
    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000)
    B, S, n = 4, 32, 6
    g = torch.Generator().manual_seed(2)
    args = (torch.randn(B, 3, S, S, generator=g), torch.rand(B, 3, S, S, generator=g), torch.ones(B, 1, S, S),
            torch.full((B, 2), S / 2.0), torch.tensor([5.0, 8.0, 11.0, 14.0]))
    one = ShadowSampler(model, gd, B, S, ddim_timesteps=n)(*args).clone()
    two = ShadowSampler(model, gd, B, S, ddim_timesteps=n, streams=2)(*args).clone()
    # No reduction in the path depends on the batch size (GroupNorm partial sums are chunked per image size only,
    # conv / attention tiles never mix images in a sum), so a trajectory is bit-reproducible in any (sub-)batch.
    # (Round 1 chunked the statistics by batch size: fp32 summation-order noise, amplified by 1/sqrt(a_t) over the
    # steps, showed up as a 5e-2 difference here.)
    assert torch.equal(one, two)
    solo = ShadowSampler(model, gd, 1, S, ddim_timesteps=n)(*(a[2:3] for a in args)).clone()
    assert torch.equal(solo[0], one[2])
    model.release_engines()


def test_ddim_sample_graph_reuse_across_calls(pkg):
    """The captured per-step graph is cached on the engine; later calls with another step count / schedule
    only refill the device tables.  Graph replays must equal eager launches bit for bit."""
    model, _ = get_model(pkg, "dm1")
    model.set_precision("bf16")
    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000)
    torch.manual_seed(0)
    xT = torch.randn(2, 3, 32, 32)
    outs = {}
    for n, method in ((6, "uniform"), (4, "quad"), (6, "uniform")):
        gd.use_cuda_graph = True
        a = gd.ddim_sample(model, 32, batch_size=2, channels=3, ddim_timesteps=n, ddim_discr_method=method, x_T=xT)
        gd.use_cuda_graph = False
        b = gd.ddim_sample(model, 32, batch_size=2, channels=3, ddim_timesteps=n, ddim_discr_method=method, x_T=xT)
        assert np.array_equal(a, b), (n, method)
        outs.setdefault((n, method), a)
        assert np.array_equal(outs[(n, method)], a)
    gd.use_cuda_graph = True
    model.release_engines()


@pytest.mark.parametrize("cfg", [
    dict(model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,), num_heads=2, num_res_blocks=1, H=32, W=48),
    dict(model_channels=32, channel_mult=(1, 2, 2), attention_resolutions=(1, 4), num_heads=4, num_res_blocks=2, H=24, W=24),
])
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-4), ("bf16", 2e-2)])
def test_custom_unet_configs_vs_oracle_port(pkg, cfg, precision, tol):
    """Non-default UNet shapes (non-square input, channel counts that force the SIMT kernels in bf16 mode,
    attention at full resolution) against the CPU oracle port with the same seeded weights."""
    from oracle import torch_port as P
    cfg = dict(cfg)
    H, W = cfg.pop("H"), cfg.pop("W")
    torch.manual_seed(0)
    m = pkg["dm1"].UNetModel(**cfg).eval()
    params = {k: v.detach().clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(4)
    x, t = torch.randn(2, 3, H, W, generator=g), torch.tensor([999, 3])
    with torch.no_grad():
        ref = P.unet_forward(params, dict(model_channels=cfg["model_channels"], num_res_blocks=cfg["num_res_blocks"],
                                          attention_resolutions=cfg["attention_resolutions"], channel_mult=cfg["channel_mult"],
                                          num_heads=cfg["num_heads"]), x, t)
    m = m.cuda().set_precision(precision)
    eps = m(x.cuda(), t.cuda()).cpu()
    err = (eps - ref).abs().max().item()
    print(f"custom UNet {cfg} {H}x{W} {precision}: max|eps err| = {err:.3e} (|eps|max {ref.abs().max():.2f})")
    assert err <= tol


def test_batched_shadow_optimisation_equals_per_image_loop(pkg):
    """optimize_shadow_position_batched == the reference-shaped single-image method applied to each
    image (dm2:457-550), with a small PyTorch victim."""
    gd = pkg["dm2"].GaussianDiffusion()
    torch.manual_seed(0)

    class Victim:
        model = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, stride=2, padding=1), torch.nn.ReLU(), torch.nn.Flatten(),
                                    torch.nn.Linear(8 * 16 * 16, 37)).cuda().eval()

    g = torch.Generator().manual_seed(8)
    B, S = 3, 32
    imgs = torch.rand(B, 3, S, S, generator=g).cuda()
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    masks = torch.stack([(((xx - cx) ** 2 + (yy - cy) ** 2) <= 100).float()[None] for cx, cy in ((10, 12), (20, 16), (15, 22))]).cuda()
    target = torch.tensor([3, 7, 11]).cuda()
    cb, rb, ob = gd.optimize_shadow_position_batched(Victim, imgs, masks, target, "cuda", iterations=6)
    for i in range(B):
        c, r, o = gd.optimize_shadow_position(Victim, imgs[i], masks[i], target[i:i + 1], "cuda", iterations=6)
        assert torch.allclose(cb[i].cpu(), c.cpu(), atol=1e-5) and abs(float(rb[i]) - float(r)) < 1e-5
        assert torch.allclose(ob[i], o[0], atol=1e-6)
    assert float(rb.max()) < 20.0        # Adam on the regulariser shrank the radii, as in the reference


def test_dm2_at_256_forward_and_teacher_forced_steps(pkg, golden):
    """The headline shape itself (BASELINE.json configs[1]: dm2.UNetModel() at 256x256): attention at T=4096/dh=128
    and T=1024/dh=256, the 128-channel full-resolution halo convs at W=256, halo<256> at W=128.  One forward and
    steps 0, 1, 49 of the reference's DDIM-50 loop (linear schedule), teacher-forced; fp32 <= 1e-4, bf16 <= 2e-2."""
    from advshadow_b200 import _capi as capi
    import ctypes as C
    g = golden("dm2_256.pt")
    model, chk = get_model(pkg, "dm2")
    assert abs(chk - g["dm2_checksum"]) <= 1e-6 * chk
    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000, beta_schedule="linear")
    d = g["ddim"]
    torch.manual_seed(d["x_T_seed"])
    x_T = torch.randn(1, 3, 256, 256)
    seq, prev = pkg["dm1"].ddim_timestep_tables(1000, d["n"], "uniform")
    coef = gd.ddim_coefficients(seq, prev, d["n"], 0.0).cuda()
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    for precision, tol in (("bf16", TOL_BF16), ("bf16-pure", None), ("fp32", TOL_FP32)):
        if precision == "bf16-pure":
            # informative: every GEMM operand bf16 (gemm_operands="bf16").  At this resolution the error has heavy
            # tails (a few pixels with large activations), so single steps exceed 2e-2 -- why fp16 operands are the default
            eng = model.engine(1, 256, 256, precision="bf16", gemm_operands="bf16")
            for j, i in enumerate(d["steps"]):
                x = (x_T if d["x"][j] is None else d["x"][j]).cuda()
                err = (eng.forward(x, d["t"][j].cuda()).cpu() - d["eps"][j]).abs().max().item()
                print(f"dm2 256x256 all-bf16 operands (informative): DDIM-50 step {i} max|eps err| = {err:.3e}")
            model.release_engines()
            continue
        eng = model.engine(1, 256, 256, precision=precision)
        if precision == "bf16":
            names = [n for (_, _, n) in eng._launches]
            assert "attn_sm100" in names and "conv_sm100" in names and "attn_simt" not in names and "conv_simt" not in names
        e = eng.forward(g["fwd"]["x"].cuda(), g["fwd"]["t"].cuda())
        err = (e.cpu() - g["fwd"]["eps"]).abs().max().item()
        print(f"dm2 256x256 {precision}: forward max|eps err| = {err:.3e} (|eps|max {g['fwd']['eps'].abs().max():.2f})")
        assert err <= tol
        for j, i in enumerate(d["steps"]):
            x = (x_T if d["x"][j] is None else d["x"][j]).cuda()
            e = eng.forward(x, d["t"][j].cuda())
            err = (e.cpu() - d["eps"][j]).abs().max().item()
            print(f"dm2 256x256 {precision}: DDIM-50 step {i} (t={int(d['t'][j])}) max|eps err| = {err:.3e}")
            assert err <= tol
            if j >= 1:      # the update applied to the oracle's eps reproduces the oracle's next state bit for bit
                out = torch.empty_like(x)
                step.fill_(i)
                capi.call("advs_ddim_step", x.data_ptr(), d["eps"][j].cuda().data_ptr(), None, out.data_ptr(), x.numel(),
                          coef.data_ptr(), step.data_ptr(), 0, 1, C.c_void_p(torch.cuda.current_stream().cuda_stream))
                assert torch.equal(out.cpu(), d["x_next"][j - 1])
        model.release_engines()


def test_blur_flavour_composites_bit_exact(pkg, golden):
    """apply_shadow with the 5x5-blurred mask (tools/train_shadow.py:224-266, I=0.43; ddim2/test.py:830-871, I=0.051)
    against the reference's own function bodies (tests/golden/shadow_blur.pt)."""
    sh = pkg["shadow"]
    for c in golden("shadow_blur.pt"):
        img, fm, adv = c["img"].cuda(), c["fm"].cuda(), c["adv"].cuda()
        H, W = img.shape[1:]
        cen, rad = c["center"][None].cuda(), c["radius"][None].cuda()
        mb = sh.gaussian_blur5(sh.disk_mask(cen, rad, H, W))
        assert torch.equal(mb[0].cpu(), c["blurred"])
        shadowed, out = sh.composite(img[None], mb, fm[None], c["intensity"], adv=adv[None])
        assert torch.equal(shadowed[0].cpu(), c["shadowed"]) and torch.equal(out[0].cpu(), c["out"])
        # the sampler's fused tail: mask + blur built in-kernel, the generated image (already in [0,1]) injected
        fused = sh.composite_generated(img[None], adv[None], cen, rad, fm[None], blur=True)
        assert torch.equal(fused[0].cpu(), c["out"])


@pytest.mark.parametrize("flavour,blur", [("dm2", False), ("ts", True), ("dt", True)])
def test_shadow_sampler_tail_flavours_equal_oracle_chain(pkg, flavour, blur):
    """ShadowSampler's fused last-step kernel == oracle port: apply_shadow(blur) on the sampler's own final x_0."""
    from oracle import torch_port as P
    from advshadow_b200.sampler import ShadowSampler
    model, _ = get_model(pkg, "dm1")
    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000)
    B, S, n = 3, 32, 3
    g = torch.Generator().manual_seed(5)
    x_T, clean = torch.randn(B, 3, S, S, generator=g), torch.rand(B, 3, S, S, generator=g)
    fmask = (torch.rand(B, 1, S, S, generator=g) > 0.2).float()
    centers = torch.tensor([[16.0, 14.5], [1.0, 30.0], [25.2, 8.1]])
    radii = torch.tensor([9.0, 12.5, 6.0])
    smp = ShadowSampler(model, gd, B, S, ddim_timesteps=n, shadow_flavour=flavour)
    out = smp(x_T, clean, fmask, centers, radii)
    x0 = smp.eng.x.cpu()
    for i in range(B):
        ref = P.apply_shadow(clean[i], centers[i], radii[i], fmask[i], 0.43 if flavour == "ts" else 0.33,
                             perturb=lambda s_, i=i: x0[i].clamp(0, 1)[None], blur=blur)[0][0]
        assert torch.equal(out[i], ref), (flavour, i)
    model.release_engines()


def test_whole_trajectory_graph_equals_per_step_graphs_and_eager(pkg):
    """One CUDA graph for the whole trajectory (n steps + fused composite) == n replays of a one-step graph == eager."""
    from advshadow_b200.sampler import ShadowSampler
    model, _ = get_model(pkg, "dm1")
    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000)
    B, S, n = 2, 32, 5
    g = torch.Generator().manual_seed(6)
    args = (torch.randn(B, 3, S, S, generator=g), torch.rand(B, 3, S, S, generator=g), torch.ones(B, 1, S, S),
            torch.full((B, 2), S / 2.0), torch.tensor([7.0, 12.0]))
    outs = [ShadowSampler(model, gd, B, S, ddim_timesteps=n, graph_scope=scope, use_graph=use)(*args).clone()
            for scope, use in (("trajectory", True), ("step", True), ("trajectory", False))]
    assert torch.equal(outs[0], outs[1]) and torch.equal(outs[0], outs[2])
    smp = ShadowSampler(model, gd, B, S, ddim_timesteps=n)
    a = smp(*args).clone()
    b = smp(*args).clone()            # the graph resets its own step counter: replays are idempotent
    assert torch.equal(a, b) and torch.equal(a, outs[0])
    model.release_engines()


def test_sampler_follows_weight_updates(pkg):
    """weights changed after the sampler (and its CUDA graph) was built -- load_state_dict, in-place edits,
    re-assigned parameter storage -- must be picked up by the next batch; captured graphs stay valid because the
    engine owns every buffer the kernels read."""
    from advshadow_b200.sampler import ShadowSampler
    torch.manual_seed(0)
    model = pkg["dm1"].UNetModel(model_channels=64, channel_mult=(1, 2), attention_resolutions=(2,), num_res_blocks=1).eval().cuda()
    gd = pkg["dm1"].GaussianDiffusion(timesteps=1000)
    B, S, n = 2, 32, 3
    g = torch.Generator().manual_seed(7)
    args = (torch.randn(B, 3, S, S, generator=g), torch.rand(B, 3, S, S, generator=g), torch.ones(B, 1, S, S),
            torch.full((B, 2), S / 2.0), torch.tensor([9.0, 13.0]))
    smp = ShadowSampler(model, gd, B, S, ddim_timesteps=n)
    before = smp(*args).clone()
    sd = {k: v.clone() for k, v in model.state_dict().items()}
    torch.manual_seed(123)
    new_sd = {k: (v + 0.05 * torch.randn_like(v)) for k, v in sd.items()}
    model.load_state_dict(new_sd)                                   # in-place copy into the same storage
    after = smp(*args).clone()
    fresh = ShadowSampler(model, gd, B, S, ddim_timesteps=n, use_graph=False)(*args).clone()
    assert not torch.equal(before, after) and torch.equal(after, fresh)
    model.load_state_dict(sd, assign=True)                          # parameter storage replaced
    again = smp(*args).clone()
    assert torch.equal(again, before)
    with torch.no_grad():
        getattr(model.out, "0").weight.mul_(1.5)                               # GroupNorm affine edited in place
    edited = smp(*args).clone()
    fresh = ShadowSampler(model, gd, B, S, ddim_timesteps=n, use_graph=False)(*args).clone()
    assert torch.equal(edited, fresh) and not torch.equal(edited, before)
    model.release_engines()


def test_two_devices_in_one_process(pkg, golden):
    """kernel attributes (dynamic shared memory) are per device; a second GPU in the same process must work."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    g = golden("forwards.pt")["dm1_32"]
    outs = []
    for dev in (0, 1):
        torch.manual_seed(0)
        m = pkg["dm1"].UNetModel().eval().to(f"cuda:{dev}")
        with torch.cuda.device(dev):
            outs.append(m(g["x"].to(f"cuda:{dev}"), g["t"].to(f"cuda:{dev}")).cpu())
    assert torch.equal(outs[0], outs[1]) and (outs[0] - g["eps"]).abs().max().item() <= TOL_BF16


def test_shadow_optimisation_matches_reference_golden(pkg, golden):
    """SURVEY 8f row 3 against the REFERENCE: optimize_shadow_position (dm2:457-550, with apply_shadow dm2:615-654 and
    the FGSM step dm2:572-613), 10 iterations, tiny seeded victim -- single-image and batched methods vs the centre /
    radius / final image the reference's own code produced (tests/golden/shadow_opt.pt).  The victim runs in PyTorch
    on the GPU here and on the CPU in the golden: FGSM takes the SIGN of its gradient, so an element whose gradient is
    within float noise of zero may flip and move that pixel by 2*epsilon."""
    from oracle import torch_port as P
    g = golden("shadow_opt.pt")
    gd = pkg["dm2"].GaussianDiffusion()

    class Victim:
        model = P.TinyVictim().eval()

    Victim.model.load_state_dict(g["victim_state"])
    Victim.model.cuda()
    imgs = torch.stack([c["img"] for c in g["cases"]]).cuda()
    masks = torch.stack([c["mask"] for c in g["cases"]]).cuda()
    labels = torch.cat([c["label"] for c in g["cases"]]).cuda()

    def check(center, radius, out, c, who):
        assert torch.allclose(center.cpu(), c["center"], atol=1e-5), who
        assert abs(float(radius) - float(c["radius"])) < 1e-5, who
        diff = (out.cpu().reshape(c["out"].shape) - c["out"]).abs()
        flipped = (diff > 1e-6).float().mean().item()
        print(f"{who}: centre {center.tolist()}, radius {float(radius):.4f}; {flipped:.5f} of the pixels differ, max {diff.max():.4f}")
        assert flipped <= 2e-3 and diff.max().item() <= 0.0201

    for i, c in enumerate(g["cases"]):
        ctr, rad, out = gd.optimize_shadow_position(Victim, imgs[i], masks[i], labels[i:i + 1], "cuda", iterations=g["iterations"])
        check(ctr, rad, out, c, f"single image {i}")
    cb, rb, ob = gd.optimize_shadow_position_batched(Victim, imgs, masks, labels, "cuda", iterations=g["iterations"])
    for i, c in enumerate(g["cases"]):
        check(cb[i], rb[i], ob[i], c, f"batched image {i}")


def test_compute_asr_equals_reference_function(pkg, golden, tmp_path):
    """attack.compute_asr (batched victim call, decisions by advs_success_flags) on the files of tests/golden/asr.pt
    gives the ASR the reference's own compute_asr (ASR_fast.py:101-126) printed for them; the top-2 logit margins of
    that golden are >= 4.8e-2, far above any fp32 / TF32 difference between the CPU and GPU victim."""
    from PIL import Image
    from advshadow_b200 import attack
    g = golden("asr.pt")
    for n, px in zip(g["names"], g["pixels"]):
        Image.fromarray(px.numpy()).save(str(tmp_path / n))
    (tmp_path / "notes.txt").write_text("not an image")
    victim = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 5, stride=4), torch.nn.Tanh(), torch.nn.AdaptiveAvgPool2d(3),
                                 torch.nn.Flatten(), torch.nn.Linear(36, 37)).eval()
    victim.load_state_dict(g["victim_state"])
    victim.cuda()
    int_to_label = {int(i): l for i, l in g["id2label"].items()}
    assert attack.compute_asr(str(tmp_path), victim, int_to_label, batch_size=3) == g["asr"]     # ragged last batch
    assert attack.compute_asr(str(tmp_path), victim, int_to_label) == g["successes"] / g["total"]
