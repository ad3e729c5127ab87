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


def schedules():
    dm1, dm2 = R.dm1(), R.dm2()
    out = {}
    for name, gd in (("cosine", dm1.GaussianDiffusion()), ("linear", dm2.GaussianDiffusion())):
        out[name] = {k: v for k, v in vars(gd).items() if torch.is_tensor(v)}
    return out


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(8)
    torch.save(config1(), os.path.join(OUT, "config1.pt"))
    torch.save(forwards(), os.path.join(OUT, "forwards.pt"))
    torch.save(shadow_cases(), os.path.join(OUT, "shadow.pt"))
    torch.save(schedules(), os.path.join(OUT, "schedules.pt"))
    torch.save(stochastic_cases(), os.path.join(OUT, "stochastic.pt"))
    torch.save(iddm_cases(), os.path.join(OUT, "iddm.pt"))
    for f in sorted(os.listdir(OUT)):
        print(f, os.path.getsize(os.path.join(OUT, f)))
