"""Shadow mask construction and masked compositing on the GPU (K9).

Reference: create_shadow_mask dm2:552-570 (= ts:156-174), apply_gaussian_blur ts:147-153 (which
round-trips GPU -> CPU -> cv2 -> GPU per image), apply_shadow dm2:615-654 / ts:224-266.
All functions are batched over B images; the reference-compatible single-image wrappers live in
diff_model2.GaussianDiffusion.
"""
import ctypes as C

import torch

from . import _capi as capi
from .ops import _need_cuda, _p, _st


def mask_center(mask):
    """`torch.nonzero(mask).float().mean(0)[1:]` (dm2:476-477): the mean (row, col) of the non-zero
    pixels of a [1,H,W] mask -- which the reference then uses as (x, y); that quirk is preserved by
    passing the result straight to `disk_mask`."""
    return torch.nonzero(mask).float().mean(0)[1:]


def disk_mask(centers, radii, H, W):
    """[B,H,W] fp32 {0,1}: sqrt((X - c[0])**2 + (Y - c[1])**2) <= r, bit-exact with dm2:567-569."""
    _need_cuda(centers, radii)
    centers = centers.detach().float().reshape(-1, 2).contiguous()
    radii = radii.detach().float().reshape(-1).contiguous()
    B = centers.shape[0]
    out = torch.empty(B, H, W, dtype=torch.float32, device=centers.device)
    with torch.cuda.device(centers.device):
        capi.call("advs_shadow_disk_mask", _p(centers), _p(radii), B, H, W, _p(out), _st())
    return out


def gaussian_blur5(mask):
    """cv2.GaussianBlur(mask, (5,5), 0) on a [B,H,W] fp32 tensor, without leaving the GPU (ts:147-153)."""
    _need_cuda(mask)
    m = mask.float().contiguous()
    B, H, W = m.shape
    out = torch.empty_like(m)
    with torch.cuda.device(m.device):
        capi.call("advs_gaussian_blur5", _p(m), _p(out), B, H, W, _st())
    return out


def composite(img, shadow_mask, feature_mask, intensity, adv=None, want_shadowed=True, want_out=True):
    """img [B,C,H,W] in [0,1]; shadow_mask [B,H,W]; feature_mask [B,1|C,H,W].
    Returns (shadowed, out): shadowed = img*(1-m) + m*(img*(1-intensity)),
    out = clamp(img*(1-m) + adv*m, 0, 1) with adv defaulting to `shadowed` (dm2:642-653)."""
    _need_cuda(img, shadow_mask, feature_mask, adv)
    img = img.float().contiguous()
    sm = shadow_mask.float().contiguous()
    fm = feature_mask.float().contiguous()
    B, Cc, H, W = img.shape
    Cm = fm.shape[1]
    shadowed = torch.empty_like(img) if want_shadowed else None
    out = torch.empty_like(img) if want_out else None
    advc = adv.float().contiguous() if adv is not None else None
    one_minus = float(1 - intensity)   # Python evaluates (1 - shadow_intensity) in double, torch casts to fp32
    with torch.cuda.device(img.device):
        capi.call("advs_shadow_composite", _p(img), _p(sm), _p(fm), Cm, _p(advc), C.c_float(one_minus), _p(shadowed),
                  _p(out), B, Cc, H, W, _st())
    return shadowed, out


def composite_generated(img, x_final, centers, radii, feature_mask, blur=False):
    """Fused tail of the shadow sampler: out = clamp(img*(1-m) + clip(x_final,0,1)*m, 0, 1) with
    m = disk(centers, radii) * feature_mask built in-kernel (dm2:634-653); blur=True passes the disk through the
    5x5 Gaussian first (tools/train_shadow.py:244-247, ddim2/test.py:851-854)."""
    _need_cuda(img, x_final, centers, radii, feature_mask)
    img = img.float().contiguous()
    xf = x_final.float().contiguous()
    fm = feature_mask.float().contiguous()
    centers = centers.detach().float().reshape(-1, 2).contiguous()
    radii = radii.detach().float().reshape(-1).contiguous()
    B, Cc, H, W = img.shape
    out = torch.empty_like(img)
    with torch.cuda.device(img.device):
        capi.call("advs_shadow_composite_generated_ex", _p(img), _p(xf), _p(centers), _p(radii), _p(fm), fm.shape[1],
                  1 if blur else 0, _p(out), B, Cc, H, W, _st())
    return out
