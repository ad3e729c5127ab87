"""TEST INFRASTRUCTURE ONLY -- mints tests/golden/*.pt from the UNMODIFIED reference modules.

Run in the dev container (needs /root/reference):  python oracle/make_golden.py
The reference has no tests or golden vectors of its own, so parity is pinned to
"the reference's own code, run by torch 2.11 CPU fp32, on these seeded inputs".  Weights are not
stored: `torch.manual_seed(seed); UNetModel()` reproduces them (a checksum is stored to detect drift).
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import ref_loader as R  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def weight_checksum(model):
    return float(sum(p.detach().double().abs().sum() for p in model.parameters()))


def patched_randn(mod, tensors):
    """Make the reference's torch.randn / randn_like return our tensors in order (dm1:444,470)."""
    it = iter(tensors)
    real = mod.torch

    class T:
        def __getattr__(self, k):
            return getattr(real, k)

        def randn(self, *a, **k):
            return next(it).clone()

        def randn_like(self, x, *a, **k):
            return torch.zeros_like(x)   # multiplied by sigma = 0 at eta = 0

    return T()


def config1():
    """BASELINE.json configs[0]: dm1.UNetModel(), one 64x64 image + binary mask, 10 DDIM steps on CPU."""
    dm1, dm2 = R.dm1(), R.dm2()
    torch.manual_seed(0)
    model = dm1.UNetModel().eval()
    gd = dm1.GaussianDiffusion(timesteps=1000)
    torch.manual_seed(1234)
    x_T = torch.randn(1, 3, 64, 64)
    # teacher-forced trace: record (x_t, t, eps) at every step by wrapping the model
    trace = []

    class Wrap(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, t):
            e = self.m(x, t)
            trace.append((x.clone(), t.clone(), e.clone()))
            return e

    saved = dm1.torch
    dm1.torch = patched_randn(dm1, [x_T])
    try:
        final = gd.ddim_sample(Wrap(model), 64, batch_size=1, channels=3, ddim_timesteps=10)
    finally:
        dm1.torch = saved
    # shadow compositing with the reference's own apply_shadow (dm2:615-654); the victim-dependent
    # perturbation is injected: adv := clip(generated, 0, 1)  (main.py:135 mapping)
    g = torch.Generator().manual_seed(7)
    clean = torch.rand(3, 64, 64, generator=g)
    yy, xx = torch.meshgrid(torch.arange(64), torch.arange(64), indexing="ij")
    fmask = (((yy - 32) ** 2 + (xx - 32) ** 2) <= 20 ** 2).float()[None]   # [1,64,64] binary disk
    gen01 = torch.from_numpy(np.clip(final, 0, 1))[0]

    class GD2(dm2.GaussianDiffusion):
        def apply_adversarial_perturbation(self, classifier, image, target_label, device, epsilon=0.00001):
            self.seen_shadowed = image.clone()
            return gen01[None]

    gd2 = GD2()
    center = torch.nonzero(fmask).float().mean(0)[1:]
    radius = torch.tensor(20.0)
    out = gd2.apply_shadow(clean, center, radius, fmask, None, None, "cpu")
    smask = gd2.create_shadow_mask((3, 64, 64), center, radius, "cpu")
    return dict(
        weight_checksum=weight_checksum(model), x_T=x_T, final=torch.from_numpy(final),
        trace_x=torch.stack([t[0] for t in trace]), trace_t=torch.stack([t[1] for t in trace]),
        trace_eps=torch.stack([t[2] for t in trace]),
        clean=clean, feature_mask=fmask, center=center, radius=radius, shadow_mask=smask,
        shadowed=gd2.seen_shadowed, composite=out)


def forwards():
    """Single UNet forwards of both model flavours at small sizes."""
    dm1, dm2 = R.dm1(), R.dm2()
    out = {}
    torch.manual_seed(0)
    m1 = dm1.UNetModel().eval()
    out["dm1_checksum"] = weight_checksum(m1)
    g = torch.Generator().manual_seed(11)
    with torch.no_grad():
        for size, B in ((32, 2), (64, 1)):
            x = torch.randn(B, 3, size, size, generator=g)
            t = torch.tensor([981, 1][:B])
            out[f"dm1_{size}"] = dict(x=x, t=t, eps=m1(x, t))
        # main.py:71-77 configuration (attention at the 2x-downsampled level: T = (S/2)^2)
        torch.manual_seed(0)
        m3 = dm1.UNetModel(channel_mult=(1, 2, 2, 2), attention_resolutions=(2,), dropout=0.1).eval()
        x = torch.randn(1, 3, 32, 32, generator=g)
        t = torch.tensor([501])
        out["main_32"] = dict(x=x, t=t, eps=m3(x, t), checksum=weight_checksum(m3))
        torch.manual_seed(0)
        m2 = dm2.UNetModel().eval()
        out["dm2_checksum"] = weight_checksum(m2)
        for size in (64, 128):
            x = torch.randn(1, 3, size, size, generator=g)
            t = torch.tensor([741])
            out[f"dm2_{size}"] = dict(x=x, t=t, eps=m2(x, t))
    return out


def shadow_cases():
    """create_shadow_mask / apply_shadow (dm2:552-570, 615-654) on a few geometries, adv injected."""
    dm2 = R.dm2()
    cases = []
    g = torch.Generator().manual_seed(3)
    for (H, W, Cm, cx, cy, r, inten) in [(64, 64, 1, 31.5, 30.25, 20.0, 0.33), (48, 80, 3, 10.0, 70.5, 15.0, 0.43),
                                         (224, 224, 1, 100.3, 120.7, 55.5, 0.051), (32, 32, 1, 0.0, 0.0, 16.0, 0.33)]:
        img = torch.rand(3, H, W, generator=g)
        fm = torch.rand(Cm, H, W, generator=g)          # soft masks (bilinear-resized 'L' masks are not binary)
        adv = torch.rand(1, 3, H, W, generator=g)

        class GD2(dm2.GaussianDiffusion):
            def apply_adversarial_perturbation(self, classifier, image, target_label, device, epsilon=0.00001):
                self.seen = image.clone()
                return adv

        gd = GD2()
        c, rr = torch.tensor([cx, cy]), torch.tensor(r)
        out = gd.apply_shadow(img, c, rr, fm, None, None, "cpu", shadow_intensity=inten)
        cases.append(dict(img=img, fm=fm, adv=adv, center=c, radius=rr, intensity=inten,
                          mask=gd.create_shadow_mask((3, H, W), c, rr, "cpu"), shadowed=gd.seen, out=out))
    return cases


def stochastic_cases():
    """DDPM ancestral `sample` (dm1:356-413, what main.py:124 calls) and DDIM with eta>0, with the
    reference's torch.randn / randn_like draws replaced by recorded tensors."""
    dm1 = R.dm1()
    torch.manual_seed(0)
    model = dm1.UNetModel().eval()
    g = torch.Generator().manual_seed(99)
    out = {}

    class Feed:
        def __init__(self, real, tensors):
            self.real, self.it = real, iter(tensors)

        def __getattr__(self, k):
            return getattr(self.real, k)

        def randn(self, *a, **k):
            return next(self.it).clone()

        def randn_like(self, x, *a, **k):
            return next(self.it).clone()

    # DDPM, T = 20
    T = 20
    gd = dm1.GaussianDiffusion(timesteps=T)
    draws = [torch.randn(1, 3, 32, 32, generator=g) for _ in range(T + 1)]     # x_T, then one z per step
    saved = dm1.torch
    dm1.torch = Feed(saved, draws)
    try:
        imgs = gd.sample(model, 32, batch_size=1, channels=3)
    finally:
        dm1.torch = saved
    out["ddpm"] = dict(T=T, x_T=draws[0], noise=torch.stack(draws[1:]), traj=torch.from_numpy(np.stack(imgs)))
    # DDIM eta = 0.5, 5 steps of T = 1000
    gd = dm1.GaussianDiffusion(timesteps=1000)
    draws = [torch.randn(2, 3, 32, 32, generator=g) for _ in range(6)]
    dm1.torch = Feed(saved, draws)
    try:
        img = gd.ddim_sample(model, 32, batch_size=2, channels=3, ddim_timesteps=5, ddim_eta=0.5)
    finally:
        dm1.torch = saved
    out["ddim_eta"] = dict(x_T=draws[0], noise=torch.stack(draws[1:]), final=torch.from_numpy(img), eta=0.5, n=5)
    return out


def iddm_cases():
    """IDDM class-conditional UNet (unet.py:17-128) and DDIMDiffusion.sample with CFG (ddim.py:48-100)."""
    UNet, DDIM = R.iddm()
    out = {}
    g = torch.Generator().manual_seed(31)
    for size, B in ((32, 2), (64, 1)):
        torch.manual_seed(0)
        net = UNet(num_classes=37, image_size=size, device="cpu").eval()
        x = torch.randn(B, 3, size, size, generator=g)
        t = torch.tensor([777, 12][:B])
        y = torch.tensor([3, 36][:B])
        with torch.no_grad():
            out[f"fwd_{size}"] = dict(x=x, t=t, y=y, eps_cond=net(x, t, y), eps_uncond=net(x, t, None),
                                      checksum=weight_checksum(net))
    # CFG sampling, 5 steps of 1000, 32x32, with a per-call trace of the denoiser
    torch.manual_seed(0)
    net = UNet(num_classes=37, image_size=32, device="cpu").eval()
    trace = []
    real_forward = net.forward

    def traced(x, time, y=None):
        e = real_forward(x, time, y)
        trace.append((x.clone(), time.clone(), None if y is None else y.clone(), e.clone()))
        return e

    net.forward = traced
    ddim = DDIM(noise_steps=1000, sample_steps=5, img_size=32, device="cpu")
    labels = torch.tensor([5, 20])
    torch.manual_seed(5)
    x_T = torch.randn((2, 3, 32, 32))
    torch.manual_seed(5)                       # the reference draws x_T with the global CPU generator (ddim.py:61)
    img = ddim.sample(net, 2, labels=labels, cfg_scale=3)
    out["sample"] = dict(x_T=x_T, labels=labels, cfg_scale=3, sample_steps=5, image=img,
                         trace_x=torch.stack([t[0] for t in trace]), trace_t=torch.stack([t[1] for t in trace]),
                         trace_has_y=torch.tensor([t[2] is not None for t in trace]),
                         trace_eps=torch.stack([t[3] for t in trace]))
    assert torch.equal(trace[0][0], x_T)
    return out


def dm2_256():
    """BASELINE.json configs[1] at its own shape: dm2.UNetModel() at 256x256 (attention T=4096/dh=128 and
    T=1024/dh=256, 128-channel full-resolution convs), B=1.  One stand-alone forward plus the teacher-forced trace
    of steps 0, 1 and 49 of the reference's 50-step DDIM loop (dm1:416-474 with dm2's linear schedule, R1 in
    SURVEY.md).  x_T is reproducible from its seed and is not stored."""
    dm1, dm2 = R.dm1(), R.dm2()
    torch.manual_seed(0)
    model = dm2.UNetModel().eval()
    out = {"dm2_checksum": weight_checksum(model)}
    g = torch.Generator().manual_seed(256)
    x = torch.randn(1, 3, 256, 256, generator=g)
    t = torch.tensor([741])
    with torch.no_grad():
        out["fwd"] = dict(x=x, t=t, eps=model(x, t))
    gd = dm1.GaussianDiffusion(timesteps=1000, beta_schedule="linear")
    torch.manual_seed(1234)
    x_T = torch.randn(1, 3, 256, 256)
    trace = []

    class Wrap(torch.nn.Module):
        def __init__(self, m):
            super().__init__()
            self.m = m

        def forward(self, x, t):
            e = self.m(x, t)
            trace.append((x.clone(), t.clone(), e.clone()))
            return e

    saved = dm1.torch
    dm1.torch = patched_randn(dm1, [x_T])
    try:
        final = gd.ddim_sample(Wrap(model), 256, batch_size=1, channels=3, ddim_timesteps=50)
    finally:
        dm1.torch = saved
    assert len(trace) == 50 and torch.equal(trace[0][0], x_T)
    keep = (0, 1, 49)
    out["ddim"] = dict(x_T_seed=1234, steps=keep, n=50,
                       x=[None if i == 0 else trace[i][0] for i in keep],          # step 0's input is x_T
                       t=torch.stack([trace[i][1] for i in keep]),
                       eps=torch.stack([trace[i][2] for i in keep]),
                       x_next=[trace[2][0], torch.from_numpy(final)])              # after step 1, after step 49
    return out


def _ts_functions():
    """The blur-flavour compositing functions of tools/train_shadow.py (ts:147-174, 224-266), taken from the
    reference SOURCE TEXT and executed unmodified: the module itself cannot be imported (fastai, config.choices)."""
    import ast
    import cv2
    import torch.nn.functional as F
    path = os.path.join(R.REF_ROOT, "tools", "train_shadow.py")
    src = open(path, encoding="utf-8").read()
    ns = {"torch": torch, "cv2": cv2, "F": F, "np": np}
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name in ("apply_gaussian_blur", "create_shadow_mask", "apply_shadow"):
            exec(compile(ast.Module([node], []), path, "exec"), ns)
    return ns


def shadow_blur_cases():
    """apply_shadow with the 5x5 Gaussian-blurred mask: ts flavour (intensity 0.43, ts:224-266) and the
    ddim2/test.py flavour (same arithmetic, intensity 0.051, dt:830-871), adversarial image injected."""
    ns = _ts_functions()
    cases = []
    g = torch.Generator().manual_seed(13)
    for (H, W, Cm, cx, cy, r, inten, soft) in [(64, 64, 1, 31.5, 30.25, 20.0, 0.43, False), (128, 96, 3, 40.0, 70.5, 25.0, 0.43, True),
                                               (120, 160, 1, 100.3, 60.7, 45.5, 0.051, False), (32, 32, 1, 2.0, 1.0, 12.0, 0.051, True)]:
        img = torch.rand(3, H, W, generator=g)
        fm = torch.rand(Cm, H, W, generator=g)
        if not soft:
            fm = (fm > 0.4).float()
        adv = torch.rand(3, H, W, generator=g)
        seen = {}

        def perturb(classifier, shadowed, label, device, mask, epsilon, adv=adv, seen=seen):
            seen["shadowed"], seen["mask"] = shadowed.clone(), mask.clone()
            return adv

        ns["apply_adversarial_perturbation"] = perturb
        c, rr = torch.tensor([cx, cy]), torch.tensor(r)
        out = ns["apply_shadow"](img, c, rr, fm, None, None, "cpu", shadow_intensity=inten)
        cases.append(dict(img=img, fm=fm, adv=adv, center=c, radius=rr, intensity=inten,
                          blurred=ns["apply_gaussian_blur"](ns["create_shadow_mask"]((3, H, W), c, rr, "cpu")),
                          combined=seen["mask"], shadowed=seen["shadowed"], out=out))
    return cases


class _TinyVictim(torch.nn.Module):
    """Seeded stand-in for the fastai learner's `.model` (the victim stays PyTorch on both sides)."""

    def __init__(self):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, stride=2, padding=1), torch.nn.Tanh(),
                                       torch.nn.Conv2d(8, 8, 3, stride=2, padding=1), torch.nn.Tanh(),
                                       torch.nn.AdaptiveAvgPool2d(4), torch.nn.Flatten(), torch.nn.Linear(128, 37))

    def forward(self, x):
        return self.net(x)


def shadow_opt_cases():
    """The reference's own optimize_shadow_position (dm2:457-550, with apply_shadow dm2:615-654 and the FGSM
    step dm2:572-613) run for 10 iterations against a tiny seeded victim, per image."""
    dm2 = R.dm2()

    class _NoPlot:      # the reference plots at iteration 0 (dm2:510-536); matplotlib is a stub here
        def __getattr__(self, k):
            return lambda *a, **kw: (_NoPlot(), [_NoPlot()] * 8) if k == "subplots" else _NoPlot()

        def __call__(self, *a, **kw):
            return _NoPlot()

    dm2.plt = _NoPlot()
    torch.manual_seed(77)
    victim = _TinyVictim().eval()

    class Classifier:
        model = victim

    gd = dm2.GaussianDiffusion()
    g = torch.Generator().manual_seed(8)
    S = 32
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    cases = []
    for (cx, cy, lab) in ((10, 12, 3), (20, 16, 7), (15, 22, 11)):
        img = torch.rand(3, S, S, generator=g)
        mask = (((xx - cx) ** 2 + (yy - cy) ** 2) <= 100).float()[None]
        label = torch.tensor([lab])
        import contextlib
        import io
        with contextlib.redirect_stdout(io.StringIO()):
            c, r, out = gd.optimize_shadow_position(Classifier, img.clone(), mask.clone(), label, "cpu", iterations=10)
        cases.append(dict(img=img, mask=mask, label=label, center=c, radius=r, out=out.detach()))
    return dict(cases=cases, victim_state={k: v.clone() for k, v in victim.state_dict().items()}, iterations=10)


def schedules():
    dm1, dm2 = R.dm1(), R.dm2()
    out = {}
    for name, gd in (("cosine", dm1.GaussianDiffusion()), ("linear", dm2.GaussianDiffusion())):
        out[name] = {k: v for k, v in vars(gd).items() if torch.is_tensor(v)}
    return out


def iddm_ckpt_cases():
    """utils/checkpoint.py (save_ckpt / load_model_ckpt / load_ckpt, :21-157) run on a toy module: the dict the
    reference writes, and the state_dict its loader produces for every (is_train, is_pretrain, is_distributed) mode
    from checkpoints with and without the DistributedDataParallel `module.` prefix."""
    import tempfile
    R.install_stubs()
    if R.REF_ROOT not in sys.path:
        sys.path.insert(0, R.REF_ROOT)
    from utils import checkpoint as ck

    def toy(classes=5, prefix=False):
        torch.manual_seed(classes)
        m = torch.nn.Module()
        m.label_emb = torch.nn.Embedding(classes, 8)
        m.inc = torch.nn.Conv2d(3, 4, 3)
        if prefix:
            w = torch.nn.Module()
            w.module = m
            return w
        return m

    out = {"cases": []}
    with tempfile.TemporaryDirectory() as d:
        src = toy(5)
        opt = torch.optim.SGD(src.parameters(), lr=0.1, momentum=0.9)
        ck.save_ckpt(epoch=7, save_name="ckpt_7", ckpt_model=src.state_dict(), ckpt_ema_model=None,
                     ckpt_optimizer=opt.state_dict(), results_dir=d, save_model_interval=True, start_model_interval=3,
                     num_classes=5, conditional=True, image_size=64, sample="ddim", network="unet", act="gelu",
                     classes_name=["a", "b"])
        out["files"] = sorted(os.listdir(d))
        out["saved"] = torch.load(os.path.join(d, "ckpt_last.pt"), weights_only=False)
        for ckpt_prefix in (False, True):
            ckpt_sd = toy(5, ckpt_prefix).state_dict()
            for (is_train, is_pretrain, is_dist) in ((False, False, False), (True, True, False), (True, True, True), (True, False, False)):
                for classes in (5, 9):                       # 9: the label embedding has another shape and is filtered out
                    want_prefix = is_train and ((is_pretrain and is_dist) or (not is_pretrain and ckpt_prefix))
                    model = toy(classes, want_prefix)
                    try:
                        ck.load_model_ckpt(model, dict(ckpt_sd), is_train=is_train, is_pretrain=is_pretrain, is_distributed=is_dist)
                        res = {k: v.clone() for k, v in model.state_dict().items()}
                    except Exception as e:                     # e.g. KeyError when the prefixes cannot be reconciled
                        res = type(e).__name__
                    out["cases"].append(dict(ckpt_prefix=ckpt_prefix, is_train=is_train, is_pretrain=is_pretrain,
                                             is_distributed=is_dist, classes=classes, model_prefix=want_prefix, result=res))
    return out


def state_dict_layouts():
    """state_dict key names, shapes and dtypes of the reference networks (what `model.load_state_dict(torch.load(p))`
    at main.py:115 / gen.py:557 and the IDDM checkpoint loader rely on), per configuration, plus the sum of |w| after
    `torch.manual_seed(0)` construction (same RNG consumption order => same initial weights)."""
    dm1, dm2 = R.dm1(), R.dm2()
    UNet, _ = R.iddm()
    out = {}

    def layout(make):
        torch.manual_seed(0)
        m = make()
        return dict(entries=[(k, tuple(v.shape), str(v.dtype)) for k, v in m.state_dict().items()],
                    checksum=float(sum(v.double().abs().sum() for v in m.state_dict().values() if v.is_floating_point())))

    out["dm1_default"] = layout(lambda: dm1.UNetModel())
    out["dm1_main"] = layout(lambda: dm1.UNetModel(channel_mult=(1, 2, 2, 2), attention_resolutions=(2,), dropout=0.1))
    out["dm1_small"] = layout(lambda: dm1.UNetModel(model_channels=64, num_res_blocks=1, channel_mult=(1, 2), num_heads=2,
                                                    attention_resolutions=(1, 2)))
    out["dm2_default"] = layout(lambda: dm2.UNetModel())
    out["iddm_cond64"] = layout(lambda: UNet(num_classes=10, image_size=64, device="cpu"))
    out["iddm_uncond32_gelu"] = layout(lambda: UNet(image_size=32, device="cpu", act="gelu"))
    return out


def diffusion_helper_cases():
    """The small tensor helpers of GaussianDiffusion (dm1:334-395, 475-484; dm2:656-680) on seeded inputs, for both
    module flavours (cosine / linear default schedule): _extract, q_sample, q_mean_variance,
    q_posterior_mean_variance, predict_start_from_noise, p_mean_variance, p_sample and train_losses with the
    closed-form denoiser; the reference's randn_like draws are replaced by a recorded tensor."""
    out = {}
    g = torch.Generator().manual_seed(33)
    x0 = torch.rand(4, 3, 8, 8, generator=g) * 2 - 1
    xt = torch.randn(4, 3, 8, 8, generator=g)
    z = torch.randn(4, 3, 8, 8, generator=g)
    t = torch.tensor([0, 1, 500, 999])
    for key, mod in (("dm1", R.dm1()), ("dm2", R.dm2())):
        gd = mod.GaussianDiffusion(timesteps=1000)
        model = _CheapEps(1000)
        saved = mod.torch

        class Feed:
            def __getattr__(self, k):
                return getattr(saved, k)

            def randn_like(self, x, *a, **k):
                return z.clone()

        mod.torch = Feed()
        try:
            with torch.no_grad():
                c = dict(extract=gd._extract(gd.sqrt_recip_alphas_cumprod, t, x0.shape),
                         q_sample=gd.q_sample(x0, t, noise=z), q_sample_drawn=gd.q_sample(x0, t),
                         q_mean_variance=gd.q_mean_variance(x0, t),
                         q_posterior_mean_variance=gd.q_posterior_mean_variance(x0, xt, t),
                         predict_start_from_noise=gd.predict_start_from_noise(xt, t, z),
                         p_mean_variance=gd.p_mean_variance(model, xt, t),
                         p_mean_variance_noclip=gd.p_mean_variance(model, xt, t, clip_denoised=False),
                         p_sample=gd.p_sample(model, xt, t),
                         train_losses=(gd.train_losses(model, x0, t) if key == "dm1" else gd.train_losses(model, x0, t, "cpu")))
        finally:
            mod.torch = saved
        out[key] = c
    out.update(x0=x0, xt=xt, z=z, t=t, eps_a=0.3, eps_b=0.05)
    return out


def dataset_cases():
    """main.py:9-29 and ddim2/main2.py:30-66 `CustomDataset`, taken from the reference SOURCE TEXT and executed
    unmodified on a synthetic folder: RGB PNGs, `mask_<name>` 0/255 'L' masks (mask_for_dataset.py:29,76-80), one
    missing mask and one corrupt image (main2's loader moves on to the next index), the reference's
    Resize + ToTensor transform (bilinear: mask edges come out soft), and the image_labels.json unpacking of
    main.py:47-50."""
    import ast
    import contextlib
    import io
    import json
    import tempfile
    from PIL import Image, UnidentifiedImageError
    from torch.utils.data import Dataset
    from torchvision import transforms
    classes = {}
    for key, rel in (("main", "main.py"), ("main2", "ddim2/main2.py")):
        tree = ast.parse(open(os.path.join(R.REF_ROOT, rel)).read())
        body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CustomDataset"]
        ns = {"Dataset": Dataset, "os": os, "Image": Image, "UnidentifiedImageError": UnidentifiedImageError}
        exec(compile(ast.Module(body=body, type_ignores=[]), rel, "exec"), ns)
        classes[key] = ns["CustomDataset"]
    g = torch.Generator().manual_seed(12)
    names = ["american_bulldog_12.png", "Abyssinian_3.png", "pug_1.png", "Bengal_9.png", "Birman_4.png"]
    pixels = [(torch.rand(30 + 4 * i, 44 - 3 * i, 3, generator=g) * 255).to(torch.uint8) for i in range(len(names))]
    masks = []
    for px in pixels:
        H, W = px.shape[:2]
        yy, xx = torch.meshgrid(torch.arange(H), torch.arange(W), indexing="ij")
        masks.append((((yy - H / 2) ** 2 + (xx - W / 3) ** 2) <= (H / 3) ** 2).to(torch.uint8) * 255)
    labels_json = {n: n.rsplit("_", 1)[0] for n in names}
    tf = transforms.Compose([transforms.Resize((32, 32)), transforms.ToTensor()])
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "images")), os.makedirs(os.path.join(d, "images_mask"))
        for i, (n, px, m) in enumerate(zip(names, pixels, masks)):
            if i == 3:
                open(os.path.join(d, "images", n), "wb").write(b"this is not a PNG file")     # corrupt image
            else:
                Image.fromarray(px.numpy()).save(os.path.join(d, "images", n))
            if i != 2:                                                                          # pug_1: mask missing
                Image.fromarray(m.numpy()).save(os.path.join(d, "images_mask", "mask_" + n))
        json.dump(labels_json, open(os.path.join(d, "image_labels.json"), "w"))
        image_labels = json.load(open(os.path.join(d, "image_labels.json")))
        files, labels = zip(*[(k, v) for k, v in image_labels.items()])                         # main.py:47-50
        ds1 = classes["main"](os.path.join(d, "images"), files, labels, transform=tf)
        ds2 = classes["main2"](os.path.join(d, "images"), os.path.join(d, "images_mask"), files, labels, transform=tf)
        items1 = [ds1[i] for i in (0, 1, 2, 4)]                                                 # (index 3 raises: corrupt file)
        with contextlib.redirect_stdout(io.StringIO()):
            items2 = [ds2[i] for i in range(len(names))]
    # the IDDM trainer's dataset (utils/utils_shadow.py:252-276), from the source text: RGB masks, same transform, path
    from torchvision.datasets.folder import default_loader
    tree = ast.parse(open(os.path.join(R.REF_ROOT, "utils/utils_shadow.py")).read())
    body = [n for n in tree.body if isinstance(n, ast.ClassDef) and n.name == "CustomDataset"]
    ns = {"torch": torch, "os": os, "default_loader": default_loader}
    exec(compile(ast.Module(body=body, type_ignores=[]), "utils/utils_shadow.py", "exec"), ns)
    tf3 = transforms.Compose([transforms.Resize((24, 24)), transforms.ToTensor(),
                              transforms.Normalize(mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5))])
    with tempfile.TemporaryDirectory() as d:
        os.makedirs(os.path.join(d, "images")), os.makedirs(os.path.join(d, "images_mask"))
        for n, px, m in zip(names[:2], pixels[:2], masks[:2]):
            Image.fromarray(px.numpy()).save(os.path.join(d, "images", n))
            Image.fromarray(m.numpy()).save(os.path.join(d, "images_mask", "mask_" + n))
        ds3 = ns["CustomDataset"](os.path.join(d, "images"), os.path.join(d, "images_mask"), names[:2], [5, 9], transform=tf3)
        iddm_items = [dict(zip(("image", "mask", "label", "path"), ds3[i])) for i in range(2)]

    # the IDDM generator's writers (utils/utils.py:51-89), from the source text, on a uint8 batch; read back as pixels
    import logging
    import torchvision
    tree = ast.parse(open(os.path.join(R.REF_ROOT, "utils/utils.py")).read())
    body = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name in ("save_images", "save_one_image_in_images")]
    ns = {"torchvision": torchvision, "Image": Image, "os": os, "logger": logging.getLogger("ref")}
    exec(compile(ast.Module(body=body, type_ignores=[]), "utils/utils.py", "exec"), ns)
    batch = (torch.rand(11, 3, 16, 16, generator=g) * 255).to(torch.uint8)
    with tempfile.TemporaryDirectory() as d:
        ns["save_images"](images=batch, path=os.path.join(d, "df.png"))
        ns["save_one_image_in_images"](images=batch, path=d, generate_name="df", image_format="png")
        written = sorted(os.listdir(d))
        grid = torch.from_numpy(np.array(Image.open(os.path.join(d, "df.png"))))
        singles = torch.stack([torch.from_numpy(np.array(Image.open(os.path.join(d, f"df_{i}.png")))) for i in range(len(batch))])
    writers = dict(batch=batch, written=written, grid=grid, singles=singles)
    return dict(names=names, pixels=pixels, masks=masks, files=list(files), labels=list(labels), corrupt=3, no_mask=2, writers=writers, iddm_items=iddm_items,
                main_items=[dict(image=a, label=b) for a, b in items1], main_index=[0, 1, 2, 4],
                main2_items=[dict(image=a, mask=b, label=c) for a, b, c in items2], len=(len(ds1), len(ds2)))


def asr_cases():
    """ASR_fast.py:90-126 `preprocess_image` + `compute_asr`, taken from the reference SOURCE TEXT and executed
    unmodified (the script itself loads fastai pickles from D:\\ at import) on a synthetic folder: lossless PNGs named
    by the `<label>_<n>.<ext>` rule (incl. a label with underscores and one the id2label map does not know), a
    non-image file, the reference's own config.json id2label map, a tiny seeded victim."""
    import ast
    import contextlib
    import io
    import json
    import tempfile
    from PIL import Image
    from torchvision import transforms
    src = open(os.path.join(R.REF_ROOT, "ASR_fast.py")).read()
    fns = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name in ("preprocess_image", "compute_asr")]
    ns = {"os": os, "Image": Image, "transforms": transforms, "torch": torch, "device": torch.device("cpu")}
    exec(compile(ast.Module(body=fns, type_ignores=[]), "ASR_fast.py", "exec"), ns)
    id2label = json.load(open(os.path.join(R.REF_ROOT, "config.json")))["id2label"]
    label_to_int = {label: int(i) for i, label in id2label.items()}                      # ASR_fast.py:71-75,99
    int_to_label = {v: k for k, v in label_to_int.items()}
    torch.manual_seed(4)
    victim = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 5, stride=4), torch.nn.Tanh(), torch.nn.AdaptiveAvgPool2d(3),
                                 torch.nn.Flatten(), torch.nn.Linear(36, 37)).eval()
    g = torch.Generator().manual_seed(8)
    names = ["Abyssinian_1.png", "american_bulldog_12.png", "Bengal_3.png", "great_pyrenees_7.png", "Birman_2.png",
             "not_a_pet_1.png", "Bombay_44.png", "yorkshire_terrier_5.png"]
    pixels = [(torch.rand(40 + 7 * i, 50 + 3 * i, 3, generator=g) * 255).to(torch.uint8) for i in range(len(names))]
    # make the victim right on some images: name them after what it predicts (through the same preprocessing)
    seen = []
    victim.register_forward_hook(lambda m, a, out: seen.append(out.detach().clone()))
    with tempfile.TemporaryDirectory() as d, torch.no_grad():
        for i, (n, px) in enumerate(zip(names, pixels)):
            Image.fromarray(px.numpy()).save(os.path.join(d, n))
        for i in (0, 2, 6):
            pred = int(victim(ns["preprocess_image"](os.path.join(d, names[i]))).argmax(1))
            new = f"{int_to_label[pred]}_{i}.png"
            os.rename(os.path.join(d, names[i]), os.path.join(d, new))
            names[i] = new
        open(os.path.join(d, "notes.txt"), "w").write("not an image")
        seen.clear()
        order = [f for f in os.listdir(d) if f.lower().endswith(("png", "jpg", "jpeg", "bmp", "gif"))]
        pre = {n: ns["preprocess_image"](os.path.join(d, n))[0] for n in names}
        seen.clear()
        buf = io.StringIO()
        with contextlib.redirect_stdout(buf):
            asr = ns["compute_asr"](d, victim, int_to_label)
        total, successes = [int(v) for v in buf.getvalue().split()]
    logits = {n: l for n, l in zip(order, torch.cat(seen))}          # compute_asr walks os.listdir order
    return dict(names=names, pixels=pixels, id2label=id2label, victim_state=victim.state_dict(),
                logits=torch.stack([logits[n] for n in names]),
                pre_probe=torch.stack([pre[n][:, ::16, ::16] for n in names]),
                pre_sum=torch.stack([pre[n].double().sum() for n in names]), asr=asr, total=total, successes=successes)


class _CheapEps(torch.nn.Module):
    """A closed-form stand-in denoiser eps = a*x + b*(t/T) (elementwise, so bit-reproducible anywhere): lets the
    reference's full sampling loops run in milliseconds for many schedule / discretisation settings."""

    def __init__(self, T):
        super().__init__()
        self.T = T
        self.dummy = torch.nn.Parameter(torch.zeros(1))      # ddim_sample / p_sample_loop ask for next(model.parameters())

    def forward(self, x, t):
        return x * 0.3 + (t.float() / self.T).view(-1, 1, 1, 1) * 0.05


def sampler_loop_cases():
    """The reference's complete ddim_sample (dm1:416-474) and sample (dm1:398-413) loops with a closed-form denoiser,
    over the settings that shape the timestep tables and the per-step coefficients: both beta schedules, uniform and
    quadratic discretisation, T % n != 0, eta > 0 (recorded noise draws), clip on / off.  Records the timesteps the
    loop hands to the model and its output."""
    dm1, dm2 = R.dm1(), R.dm2()
    g = torch.Generator().manual_seed(21)
    saved = dm1.torch

    class Feed:
        def __init__(self, tensors):
            self.it = iter(tensors)

        def __getattr__(self, k):
            return getattr(saved, k)

        def randn(self, *a, **k):
            return next(self.it).clone()

        def randn_like(self, x, *a, **k):
            return next(self.it).clone()

    ddim = []
    for (sched, T, n, method, eta, clip) in [("cosine", 1000, 10, "uniform", 0.0, True), ("linear", 1000, 50, "uniform", 0.0, True),
                                              ("linear", 1000, 30, "uniform", 0.0, True), ("cosine", 1000, 100, "uniform", 0.0, False),
                                              ("linear", 1000, 20, "quad", 0.0, True), ("cosine", 1000, 50, "quad", 0.3, True),
                                              ("linear", 500, 7, "uniform", 1.0, True), ("cosine", 200, 100, "uniform", 0.0, True)]:
        gd = dm1.GaussianDiffusion(timesteps=T, beta_schedule=sched)
        model, seen = _CheapEps(T), []
        model.register_forward_pre_hook(lambda m, args: seen.append(args[1].clone()))
        draws = [torch.randn(2, 3, 8, 8, generator=g) for _ in range(n + 1)]          # x_T, then one z per step
        dm1.torch = Feed(draws)
        try:
            out = gd.ddim_sample(model, 8, batch_size=2, channels=3, ddim_timesteps=n, ddim_discr_method=method,
                                 ddim_eta=eta, clip_denoised=clip)
        finally:
            dm1.torch = saved
        ddim.append(dict(schedule=sched, T=T, n=n, method=method, eta=eta, clip=clip, x_T=draws[0], noise=torch.stack(draws[1:]),
                         t=torch.stack(seen), out=torch.from_numpy(out)))
    ddpm = []
    for (sched, T) in [("cosine", 40), ("linear", 25)]:
        gd = dm1.GaussianDiffusion(timesteps=T, beta_schedule=sched)
        model = _CheapEps(T)
        draws = [torch.randn(2, 3, 8, 8, generator=g) for _ in range(T + 1)]
        dm1.torch = Feed(draws)
        try:
            imgs = gd.sample(model, 8, batch_size=2, channels=3)
        finally:
            dm1.torch = saved
        ddpm.append(dict(schedule=sched, T=T, x_T=draws[0], noise=torch.stack(draws[1:]), traj=torch.from_numpy(np.stack(imgs))))
    # IDDM DDIMDiffusion.sample (model/samples/ddim.py:48-100): fp32 linear schedule, (t, prev) pairs, classifier-free
    # guidance through torch.lerp (both branches of its formula: |w| < 0.5 and >= 0.5), uint8 cast without clamp
    _, DDIM = R.iddm()
    ddim_mod = sys.modules["model.samples.ddim"]

    class CheapCond(torch.nn.Module):
        def __init__(self, T):
            super().__init__()
            self.T = T
            self.dummy = torch.nn.Parameter(torch.zeros(1))

        def forward(self, x, t, y=None):
            e = x * 0.3 + (t.float() / self.T).view(-1, 1, 1, 1) * 0.05
            return e if y is None else e + (y.float() + 1).view(-1, 1, 1, 1) * 0.02

    iddm = []
    real = ddim_mod.torch
    for (T, n_steps, labelled, w) in [(1000, 50, False, None), (1000, 500, False, None), (1000, 20, True, 3.0), (1000, 20, True, 0.3),
                                      (300, 7, True, 0.0)]:
        diff = DDIM(noise_steps=T, sample_steps=n_steps, img_size=8, device="cpu")
        x_T = torch.randn(3, 3, 8, 8, generator=g)
        labels = torch.tensor([0, 4, 2]) if labelled else None

        class FeedI:
            def __getattr__(self, k):
                return getattr(real, k)

            def randn(self, *a, **k):
                return x_T.clone()

        ddim_mod.torch = FeedI()
        try:
            out = diff.sample(CheapCond(T), 3, labels, w)
        finally:
            ddim_mod.torch = real
        iddm.append(dict(T=T, sample_steps=n_steps, labels=labels, cfg_scale=w, x_T=x_T, out=out,
                         pairs=torch.tensor([[int(a), int(b)] for a, b in diff.time_step])))
    return dict(ddim=ddim, ddpm=ddpm, iddm=iddm, eps_a=0.3, eps_b=0.05, eps_label=0.02)


def metrics_cases():
    """fid_fast.py:30-45 `calculate_fid`, taken from the reference SOURCE TEXT and executed unmodified (the script
    itself downloads Inception weights and reads Windows paths at import), on recorded activation sets: a
    well-conditioned pair, a shifted pair, and a rank-deficient pair (fewer samples than features: sqrtm turns
    complex and the reference drops the imaginary part)."""
    import ast
    from scipy import linalg
    src = open(os.path.join(R.REF_ROOT, "fid_fast.py")).read()
    fn = [n for n in ast.parse(src).body if isinstance(n, ast.FunctionDef) and n.name == "calculate_fid"]
    ns = {"np": np, "linalg": linalg}
    exec(compile(ast.Module(body=fn, type_ignores=[]), "fid_fast.py", "exec"), ns)
    rng = np.random.RandomState(5)
    cases = []
    for n1, n2, d, shift in ((64, 80, 24, 0.0), (200, 200, 16, 0.7), (10, 12, 16, 0.1)):
        a1 = rng.randn(n1, d).astype(np.float32) * (1 + rng.rand(d).astype(np.float32))
        a2 = rng.randn(n2, d).astype(np.float32) * (1 + rng.rand(d).astype(np.float32)) + np.float32(shift)
        cases.append(dict(act1=torch.from_numpy(a1), act2=torch.from_numpy(a2), fid=float(ns["calculate_fid"](a1, a2))))
    return dict(fid=cases)


def _signature(fn):
    import inspect
    return [(q.name, q.kind.name, None if q.default is inspect.Parameter.empty else repr(q.default))
            for q in inspect.signature(fn).parameters.values()]


def _star_imported_names(script, module):
    """Names a script reads without binding them itself, that the star-imported module supplies."""
    import ast
    import builtins
    tree = ast.parse(open(os.path.join(R.REF_ROOT, script)).read())
    bound = set(dir(builtins))
    for n in ast.walk(tree):
        if isinstance(n, (ast.FunctionDef, ast.ClassDef, ast.AsyncFunctionDef)):
            bound.add(n.name)
        elif isinstance(n, ast.arg):
            bound.add(n.arg)
        elif isinstance(n, ast.Name) and isinstance(n.ctx, (ast.Store, ast.Del)):
            bound.add(n.id)
        elif isinstance(n, (ast.Import, ast.ImportFrom)):
            bound.update((a.asname or a.name).split(".")[0] for a in n.names if a.name != "*")
        elif isinstance(n, ast.ExceptHandler) and n.name:
            bound.add(n.name)
    used = {n.id for n in ast.walk(tree) if isinstance(n, ast.Name) and isinstance(n.ctx, ast.Load)}
    return sorted(u for u in used - bound if hasattr(module, u))


def api_surface():
    """The Python surface the drop-in modules must reproduce: every public name of the reference's
    diff_model / ddim2/diff_model2 namespaces (what `from diff_model import *` hands to main.py:6 / main2.py:6), the
    names those two scripts actually take from the star import, and the signatures (parameter names, order, defaults)
    of UNetModel / GaussianDiffusion / PretrainedResNet50 methods and the module-level functions."""
    import inspect
    out = {}
    for key, mod, script in (("dm1", R.dm1(), "main.py"), ("dm2", R.dm2(), "ddim2/main2.py")):
        names = sorted(n for n in vars(mod) if not n.startswith("_"))
        classes, functions = {}, {}
        for n in names:
            obj = getattr(mod, n)
            if getattr(obj, "__module__", None) != mod.__name__:
                continue
            if inspect.isclass(obj):
                classes[n] = {m: _signature(f) for m, f in vars(obj).items() if inspect.isfunction(f) and
                              (not m.startswith("__") or m == "__init__")}
            elif inspect.isfunction(obj):
                functions[n] = _signature(obj)
        out[key] = dict(names=names, classes=classes, functions=functions, script=script,
                        star_names_used=_star_imported_names(script, mod))
    # the IDDM generator: model/networks/unet.py:17-128, model/samples/ddim.py:25-100,
    # utils/checkpoint.py:21-157
    import importlib
    UNet, DDIM = R.iddm()
    ck = importlib.import_module("utils.checkpoint")
    out["iddm"] = {"UNet.__init__": _signature(UNet.__init__), "UNet.forward": _signature(UNet.forward),
                   "DDIMDiffusion.__init__": _signature(DDIM.__init__), "DDIMDiffusion.sample": _signature(DDIM.sample),
                   "save_ckpt": _signature(ck.save_ckpt), "load_model_ckpt": _signature(ck.load_model_ckpt),
                   "load_ckpt": _signature(ck.load_ckpt)}
    return out


MINTERS = dict(api_surface=api_surface, state_dicts=state_dict_layouts, diffusion_helpers=diffusion_helper_cases, datasets=dataset_cases, asr=asr_cases, metrics=metrics_cases, sampler_loops=sampler_loop_cases, iddm_ckpt=iddm_ckpt_cases, config1=config1, forwards=forwards, shadow=shadow_cases, schedules=schedules, stochastic=stochastic_cases,
               iddm=iddm_cases, dm2_256=dm2_256, shadow_blur=shadow_blur_cases, shadow_opt=shadow_opt_cases)

if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    for name in (sys.argv[1:] or list(MINTERS)):      # `python oracle/make_golden.py [name ...]`
        torch.save(MINTERS[name](), os.path.join(OUT, name + ".pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
