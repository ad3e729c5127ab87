"""TEST INFRASTRUCTURE ONLY -- CPU restatement (plain PyTorch fp32) of the reference's hot path.

The reference is Python; its heavy arithmetic lives in PyTorch (unpinned; effective pin = torch 2.11
CPU fp32 in this image).  /root/reference cannot travel to the GPU box, so this file restates the
algorithm function by function, citing the reference lines it follows.  It is pinned by
tests/test_oracle_cpu.py against tests/golden/*.pt, which were produced by the UNMODIFIED reference
(oracle/make_golden.py).  Only tests/, __graft_entry__.smoke() and the CPU / PyTorch baseline legs of the
benchmarks (bench.py's cpu_baseline, `--impl reference`, `--impl torch-gpu`; the `--cpu` column of tools/sweep.py)
may import it -- always as the thing compared against, never on the measured or shipped path; the package never does.

Parameters are addressed by the reference's state_dict keys.
"""
import math

import numpy as np
import torch
import torch.nn.functional as F


def timestep_embedding(t, dim, max_period=10000):                       # diff_model.py:16-33
    half = dim // 2
    freqs = torch.exp(-math.log(max_period) * torch.arange(0, half, dtype=torch.float32) / half).to(t.device)
    args = t[:, None].float() * freqs[None]
    return torch.cat([torch.cos(args), torch.sin(args)], dim=-1)


def _gn(p, name, x):                                                    # diff_model.py:62-63
    return F.group_norm(x, 32, p[name + ".weight"], p[name + ".bias"], 1e-5)


def _conv(p, name, x, stride=1, padding=1):
    return F.conv2d(x, p[name + ".weight"], p.get(name + ".bias"), stride=stride, padding=padding)


def residual_block(p, pre, x, emb):                                     # diff_model.py:67-103
    h = _conv(p, pre + ".conv1.2", F.silu(_gn(p, pre + ".conv1.0", x)))
    h = h + F.linear(F.silu(emb), p[pre + ".time_emb.1.weight"], p[pre + ".time_emb.1.bias"])[:, :, None, None]
    h = _conv(p, pre + ".conv2.3", F.silu(_gn(p, pre + ".conv2.0", h)))     # dropout = identity in eval()
    sc = _conv(p, pre + ".shortcut", x, padding=0) if (pre + ".shortcut.weight") in p else x
    return h + sc


def attention_block(p, pre, x, heads):                                  # diff_model.py:107-127
    B, C, H, W = x.shape
    qkv = _conv(p, pre + ".qkv", _gn(p, pre + ".norm", x), padding=0)
    q, k, v = qkv.reshape(B * heads, -1, H * W).chunk(3, dim=1)
    scale = 1. / math.sqrt(math.sqrt(C // heads))
    attn = torch.einsum("bct,bcs->bts", q * scale, k * scale).softmax(dim=-1)
    h = torch.einsum("bts,bcs->bct", attn, v).reshape(B, -1, H, W)
    return _conv(p, pre + ".proj", h, padding=0) + x


def unet_forward(p, cfg, x, t):
    """cfg: dict(model_channels, num_res_blocks, attention_resolutions, channel_mult, num_heads).
    Walks the module tree of UNetModel.__init__ / forward (diff_model.py:183-267)."""
    mc, heads = cfg["model_channels"], cfg["num_heads"]
    emb = timestep_embedding(t, mc)
    emb = F.linear(emb, p["time_embed.0.weight"], p["time_embed.0.bias"])
    emb = F.linear(F.silu(emb), p["time_embed.2.weight"], p["time_embed.2.bias"])
    hs = []
    h = _conv(p, "down_blocks.0.0", x)
    hs.append(h)
    ds, idx = 1, 1
    nlev = len(cfg["channel_mult"])
    for level in range(nlev):
        for _ in range(cfg["num_res_blocks"]):
            h = residual_block(p, f"down_blocks.{idx}.0", h, emb)
            if ds in cfg["attention_resolutions"]:
                h = attention_block(p, f"down_blocks.{idx}.1", h, heads)
            hs.append(h)
            idx += 1
        if level != nlev - 1:
            h = _conv(p, f"down_blocks.{idx}.0.op", h, stride=2)       # Downsample, diff_model.py:143-153
            hs.append(h)
            ds *= 2
            idx += 1
    h = residual_block(p, "middle_block.0", h, emb)
    h = attention_block(p, "middle_block.1", h, heads)
    h = residual_block(p, "middle_block.2", h, emb)
    uidx = 0
    for level in reversed(range(nlev)):
        for i in range(cfg["num_res_blocks"] + 1):
            h = residual_block(p, f"up_blocks.{uidx}.0", torch.cat([h, hs.pop()], dim=1), emb)
            sub = 1
            if ds in cfg["attention_resolutions"]:
                h = attention_block(p, f"up_blocks.{uidx}.{sub}", h, heads)
                sub += 1
            if level and i == cfg["num_res_blocks"]:
                h = F.interpolate(h, scale_factor=2, mode="nearest")    # Upsample, diff_model.py:129-140
                h = _conv(p, f"up_blocks.{uidx}.{sub}.conv", h)
                ds //= 2
            uidx += 1
    return _conv(p, "out.2", F.silu(_gn(p, "out.0", h)))


DM1_CFG = dict(model_channels=128, num_res_blocks=2, attention_resolutions=(8, 16), channel_mult=(1, 2, 2, 2), num_heads=4)
DM2_CFG = dict(model_channels=128, num_res_blocks=3, attention_resolutions=(4, 8, 16, 32), channel_mult=(1, 2, 4, 8),
               num_heads=4)


def cosine_alphas_cumprod(T=1000, s=0.008):                              # diff_model.py:275-285, 300-303
    x = torch.linspace(0, T, T + 1, dtype=torch.float64)
    acp = torch.cos(((x / T) + s) / (1 + s) * math.pi * 0.5) ** 2
    acp = acp / acp[0]
    betas = torch.clip(1 - (acp[1:] / acp[:-1]), 0, 0.999)
    return torch.cumprod(1. - betas, 0)


def linear_alphas_cumprod(T=1000):                                       # diff_model.py:269-273
    scale = 1000 / T
    betas = torch.linspace(scale * 0.0001, scale * 0.02, T, dtype=torch.float64)
    return torch.cumprod(1. - betas, 0)


def ddim_update(x, eps, a_t, a_p, eta=0.0, z=None, clip=True):           # diff_model.py:457-472
    x0 = (x - torch.sqrt(1. - a_t) * eps) / torch.sqrt(a_t)
    if clip:
        x0 = torch.clamp(x0, min=-1., max=1.)
    sig = eta * torch.sqrt((1 - a_p) / (1 - a_t) * (1 - a_t / a_p))
    out = torch.sqrt(a_p) * x0 + torch.sqrt(1 - a_p - sig ** 2) * eps
    return out + sig * (z if z is not None else torch.zeros_like(x))


@torch.no_grad()
def ddim_sample(p, cfg, acp, x_T, n, T=1000, eta=0.0, max_steps=None):   # diff_model.py:416-474
    c = T // n
    seq = np.asarray(list(range(0, T, c))) + 1
    prev = np.append(np.array([0]), seq[:-1])
    x = x_T.clone()
    B = x.shape[0]
    done = 0
    for i in reversed(range(n)):
        t = torch.full((B,), int(seq[i]), dtype=torch.long, device=x.device)
        a_t = acp[int(seq[i])].float().reshape(1, 1, 1, 1).to(x.device)
        a_p = acp[int(prev[i])].float().reshape(1, 1, 1, 1).to(x.device)
        x = ddim_update(x, unet_forward(p, cfg, x, t).float(), a_t, a_p, eta)
        done += 1
        if max_steps and done >= max_steps:
            break
    return x


def gaussian_blur5(mask):                                                # tools/train_shadow.py:147-153
    """cv2.GaussianBlur(mask,(5,5),0): fixed [1,4,6,4,1]/16 table, BORDER_REFLECT_101."""
    k = (torch.tensor([1., 4., 6., 4., 1.]) / 16).to(mask.device)
    m = F.pad(mask[None, None], (2, 2, 2, 2), mode="reflect")
    return F.conv2d(F.conv2d(m, k.view(1, 1, 1, 5)), k.view(1, 1, 5, 1))[0, 0]


def create_shadow_mask(H, W, center, radius):                            # ddim2/diff_model2.py:552-570
    Y, X = torch.meshgrid(torch.arange(H), torch.arange(W), indexing='ij')
    if torch.is_tensor(center):
        Y, X = Y.to(center.device), X.to(center.device)
    return (torch.sqrt((X - center[0]) ** 2 + (Y - center[1]) ** 2) <= radius).float()


def apply_shadow(img, center, radius, fmask, intensity=0.33, perturb=None, blur=False):
    """ddim2/diff_model2.py:615-654 (blur=False, I=0.33) / tools/train_shadow.py:224-266 (blur=True, I=0.43)."""
    _, H, W = img.shape
    m = create_shadow_mask(H, W, center, radius)
    if blur:
        m = gaussian_blur5(m)
    m = m * fmask
    shadowed = img * (1 - m) + m * (img * (1 - intensity))
    adv = perturb(shadowed) if perturb is not None else shadowed[None]
    return torch.clamp(img * (1 - m) + adv * m, 0, 1), shadowed, m


def fgsm_perturb(victim, image, target_label, epsilon):                  # ddim2/diff_model2.py:572-613
    x = image.detach().clone()[None].requires_grad_(True)
    with torch.enable_grad():
        loss = F.cross_entropy(victim(x), target_label)
        victim.zero_grad()
        loss.backward()
    return torch.clamp(x + epsilon * x.grad.data.sign(), 0, 1).detach()


def optimize_shadow_position(victim, img, mask, target_label, lr=1e-1, iterations=10, intensity=0.33, epsilon=0.01):
    """ddim2/diff_model2.py:457-550: Adam on (centre, radius); the hard disk mask passes no gradient, so only the
    regulariser moves them.  Returns (centre, radius, last shadowed image [1,C,H,W])."""
    mask_center = torch.nonzero(mask).float().mean(0)[1:]
    center = torch.nn.Parameter(mask_center.clone(), requires_grad=True)
    radius = torch.nn.Parameter(torch.tensor(20.0), requires_grad=True)
    opt = torch.optim.Adam([center, radius], lr=lr)
    out = None
    for _ in range(iterations):
        opt.zero_grad()
        out, _, _ = apply_shadow(img, center, radius, mask, intensity,
                                 perturb=lambda s: fgsm_perturb(victim, s, target_label, epsilon))
        x = out.squeeze(0) if out.dim() == 4 else out
        adv = -F.cross_entropy(victim(x.unsqueeze(0)), target_label)
        nat = F.mse_loss(x, img)
        reg = (center - mask_center).pow(2).sum() + radius.pow(2)
        (adv + nat + 0.1 * reg).backward()
        if center.grad is not None and radius.grad is not None:
            opt.step()
        with torch.no_grad():
            center.clamp_(min=0, max=img.size(2))
            radius.clamp_(min=0, max=min(img.size(1), img.size(2)) / 2)
    return center.detach(), radius.detach(), out


class TinyVictim(torch.nn.Module):
    """The seeded stand-in victim of tests/golden/shadow_opt.pt (same layers as oracle/make_golden._TinyVictim)."""

    def __init__(self):
        super().__init__()
        self.net = torch.nn.Sequential(torch.nn.Conv2d(3, 8, 3, stride=2, padding=1), torch.nn.Tanh(),
                                       torch.nn.Conv2d(8, 8, 3, stride=2, padding=1), torch.nn.Tanh(),
                                       torch.nn.AdaptiveAvgPool2d(4), torch.nn.Flatten(), torch.nn.Linear(128, 37))

    def forward(self, x):
        return self.net(x)
