"""CPU tests of the evaluation metrics (second half; advshadow_b200/metrics.py)."""
import os

import numpy as np
import pytest
import torch


def test_fid_equals_reference_function(golden):
    """calculate_fid against the reference's own function body (fid_fast.py:30-45) on recorded activations, incl. the
    rank-deficient case where scipy's sqrtm turns complex."""
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import metrics
    for c in golden("metrics.pt")["fid"]:
        got = metrics.calculate_fid(c["act1"].numpy(), c["act2"].numpy())
        assert abs(got - c["fid"]) <= 1e-9 * abs(c["fid"]), (got, c["fid"])
    a = np.random.RandomState(0).randn(50, 8)
    assert abs(metrics.calculate_fid(a, a)) < 1e-6          # identical sets (sqrtm's own round-off is ~1e-7 here)


def _ssim_scipy(a, b, win_size):
    """Independent float64 evaluation of the definition metrics.structural_similarity restates (Wang et al. 2004 with
    scikit-image's conventions), one HWC pair, via scipy.ndimage.gaussian_filter."""
    from scipy.ndimage import gaussian_filter
    a, b = a.astype(np.float64), b.astype(np.float64)
    R = a.max() - a.min()
    flt = lambda t: gaussian_filter(t, sigma=1.5, truncate=3.5, mode="reflect")
    NP = win_size ** 2
    cn = NP / (NP - 1)
    vals = []
    for ch in range(a.shape[2]):
        x, y = a[..., ch], b[..., ch]
        ux, uy = flt(x), flt(y)
        vx, vy, vxy = cn * (flt(x * x) - ux * ux), cn * (flt(y * y) - uy * uy), cn * (flt(x * y) - ux * uy)
        C1, C2 = (0.01 * R) ** 2, (0.03 * R) ** 2
        S = ((2 * ux * uy + C1) * (2 * vxy + C2)) / ((ux ** 2 + uy ** 2 + C1) * (vx + vy + C2))
        p = (win_size - 1) // 2
        vals.append(S[p:-p, p:-p].mean())
    return float(np.mean(vals))


@pytest.mark.parametrize("win_size", [7, 11])
def test_ssim_psnr_follow_the_published_definition(win_size):
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import metrics
    g = torch.Generator().manual_seed(3)
    yy, xx = torch.meshgrid(torch.linspace(0, 1, 64), torch.linspace(0, 1, 48), indexing="ij")
    base = torch.stack([0.5 + 0.4 * torch.sin(9 * xx + 3 * yy), yy * xx, 0.3 + 0.5 * torch.cos(7 * yy)])      # smooth CHW image
    imgs1 = torch.stack([base, torch.rand(3, 64, 48, generator=g)])
    imgs2 = torch.stack([(base + 0.05 * torch.randn(3, 64, 48, generator=g)).clamp(0, 1),
                         torch.rand(3, 64, 48, generator=g)])
    s = metrics.structural_similarity(imgs1, imgs2, win_size=win_size)
    p = metrics.peak_signal_noise_ratio(imgs1, imgs2)
    for i in range(2):
        a, b = imgs1[i].permute(1, 2, 0).numpy(), imgs2[i].permute(1, 2, 0).numpy()
        assert abs(float(s[i]) - _ssim_scipy(a, b, win_size)) < 2e-5
        R = float(a.max() - a.min())
        assert abs(float(p[i]) - 10 * np.log10(R ** 2 / np.mean((a.astype(np.float64) - b.astype(np.float64)) ** 2))) < 1e-9
        one = metrics.calculate_ssim_psnr(imgs1[i].numpy(), imgs2[i].numpy(), win_size=win_size)      # the reference's CHW helper
        assert abs(one[0] - float(s[i])) < 1e-12 and abs(one[1] - float(p[i])) < 1e-12
    assert float(s[0]) > 0.5 > float(s[1])                  # a lightly perturbed image vs two unrelated noise images
    assert abs(float(metrics.structural_similarity(imgs1, imgs1, win_size=win_size)[0]) - 1.0) < 1e-6
    assert torch.isinf(metrics.peak_signal_noise_ratio(imgs1, imgs1)).all()
    with pytest.raises(ValueError):
        metrics.structural_similarity(imgs1, imgs2, win_size=8)


def test_compare_folders_and_fid_preprocess(tmp_path):
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import datasets, metrics
    g = torch.Generator().manual_seed(1)
    a = torch.rand(3, 3, 80, 80, generator=g)
    b = (a + 0.1 * torch.randn(3, 3, 80, 80, generator=g)).clamp(0, 1)
    names = [f"Abyssinian_{i}.png" for i in range(3)]
    datasets.save_images(a, str(tmp_path / "a"), names)
    datasets.save_images(b, str(tmp_path / "b"), names)
    (tmp_path / "a" / "notes.txt").write_text("not an image")
    ssim, psnr = metrics.compare_folders(str(tmp_path / "a"), str(tmp_path / "b"))
    la, lb = metrics.load_images_from_folder(str(tmp_path / "a")), metrics.load_images_from_folder(str(tmp_path / "b"))
    assert la.shape == (3, 3, 64, 64) and os.listdir(str(tmp_path / "a"))
    assert abs(ssim - float(metrics.structural_similarity(la, lb, win_size=7).mean())) < 1e-12
    assert abs(psnr - float(metrics.peak_signal_noise_ratio(la, lb).mean())) < 1e-12 and 10 < psnr < 40
    datasets.save_images(a[:2], str(tmp_path / "c"), names[:2])
    with pytest.raises(ValueError):
        metrics.compare_folders(str(tmp_path / "a"), str(tmp_path / "c"))
    x = metrics.fid_preprocess(a)
    assert x.shape == (3, 3, 299, 299)
    # a stand-in feature extractor: get_activations batches, preprocesses and returns [N, D] numpy features
    feat = torch.nn.Sequential(torch.nn.AdaptiveAvgPool2d(2), torch.nn.Flatten(), torch.nn.Linear(12, 5))
    acts = metrics.get_activations(a, feat, batch_size=2)
    assert acts.shape == (3, 5) and np.isfinite(metrics.calculate_fid(acts, metrics.get_activations(b, feat)))
