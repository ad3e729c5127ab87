"""On-disk formats on CPU with temporary files."""
import json
import os

import numpy as np
import pytest
import torch


def test_image_mask_label_formats(tmp_path):
    import advshadow_b200
    from advshadow_b200 import datasets as D
    PIL = pytest.importorskip("PIL.Image")
    tv = pytest.importorskip("torchvision.transforms")
    img_dir, mask_dir = tmp_path / "images", tmp_path / "masks"
    img_dir.mkdir(), mask_dir.mkdir()
    names = ["american_bulldog_12.png", "Abyssinian_3.png", "pug_1.png"]
    rng = np.random.RandomState(0)
    for n in names:
        PIL.fromarray(rng.randint(0, 255, (20, 24, 3), dtype=np.uint8)).save(img_dir / n)
    for n in names[:2]:      # the third mask is missing -> sample skipped like ddim2/main2.py:63-65
        m = np.zeros((20, 24), dtype=np.uint8)
        m[5:15, 6:18] = 255
        PIL.fromarray(m).save(mask_dir / D.mask_name(n))
    (tmp_path / "image_labels.json").write_text(json.dumps({n: D.label_from_filename(n) for n in names}))
    files, labels = D.load_image_labels(str(tmp_path / "image_labels.json"))
    assert files == names and labels == ["american_bulldog", "Abyssinian", "pug"]
    assert D.list_images(str(img_dir)) == sorted(names)
    tf = tv.Compose([tv.Resize((16, 16)), tv.ToTensor()])
    ds = D.ImageMaskLabelDataset(str(img_dir), str(mask_dir), files, labels, transform=tf)
    img, mask, lab = ds[0]
    assert img.shape == (3, 16, 16) and mask.shape == (1, 16, 16) and lab == "american_bulldog"
    assert 0 < mask.min() + 1 and mask.max() <= 1 and ((mask > 0) & (mask < 1)).any()   # soft edges are kept
    _, _, lab2 = ds[2]                      # missing mask -> next readable sample
    assert lab2 == "american_bulldog"
    ds2 = D.ImageLabelDataset(str(img_dir), files, labels, transform=tf)
    assert ds2[1][1] == "Abyssinian" and len(ds2) == 3
    out = tmp_path / "shadowed_images"
    D.save_images(torch.rand(3, 3, 8, 8), str(out), names)
    assert sorted(os.listdir(out)) == sorted(names)
