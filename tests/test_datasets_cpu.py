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


def test_iddm_checkpoint_dict_matches_reference(golden, tmp_path):
    """the IDDM on-disk checkpoint (utils/checkpoint.py:21-157).  Golden = what the reference's own
    save_ckpt wrote and what its load_model_ckpt produced for every mode, with and without the DDP `module.` prefix."""
    import torch
    import advshadow_b200
    from advshadow_b200 import iddm
    g = golden("iddm_ckpt.pt")

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

    src = toy(5)
    opt = torch.optim.SGD(src.parameters(), lr=0.1, momentum=0.9)
    iddm.save_ckpt(epoch=7, save_name="ckpt_7", ckpt_model=src.state_dict(), ckpt_ema_model=None,
                   ckpt_optimizer=opt.state_dict(), results_dir=str(tmp_path), save_model_interval=True, start_model_interval=3,
                   num_classes=5, conditional=True, image_size=64, sample="ddim", network="unet", act="gelu",
                   classes_name=["a", "b"])
    import os
    assert sorted(os.listdir(tmp_path)) == g["files"]
    mine = torch.load(tmp_path / "ckpt_last.pt", weights_only=False)
    assert tuple(mine.keys()) == tuple(g["saved"].keys()) == iddm.CKPT_KEYS
    for k in mine:
        if k == "model":
            assert all(torch.equal(mine[k][n], g["saved"][k][n]) for n in g["saved"][k]) and mine[k].keys() == g["saved"][k].keys()
        elif k != "optimizer":
            assert mine[k] == g["saved"][k], k
    for c in g["cases"]:
        ckpt_sd = toy(5, c["ckpt_prefix"]).state_dict()
        model = toy(c["classes"], c["model_prefix"])
        try:
            iddm.load_model_ckpt(model, dict(ckpt_sd), is_train=c["is_train"], is_pretrain=c["is_pretrain"],
                                 is_distributed=c["is_distributed"])
            res = model.state_dict()
        except Exception as e:
            res = type(e).__name__
        if isinstance(c["result"], str):
            assert res == c["result"], c
        else:
            assert res.keys() == c["result"].keys() and all(torch.equal(res[k], v) for k, v in c["result"].items()), c
    # load_ckpt: 'model' by default, 'ema_model' when it is the only one or when asked for; resumed training returns epoch + 1
    tgt = toy(5)
    assert iddm.load_ckpt(str(tmp_path / "ckpt_last.pt"), tgt, "cpu", is_train=False) is None
    assert all(torch.equal(a, b) for a, b in zip(tgt.state_dict().values(), src.state_dict().values()))
    opt2 = torch.optim.SGD(toy(5).parameters(), lr=0.1, momentum=0.9)
    assert iddm.load_ckpt(str(tmp_path / "ckpt_last.pt"), toy(5), "cpu", optimizer=opt2, is_train=True) == 8


def test_datasets_equal_the_reference_loaders(golden, tmp_path):
    """datasets.ImageLabelDataset / ImageMaskLabelDataset against the reference's OWN `CustomDataset` classes
    (main.py:9-29, ddim2/main2.py:30-66, executed from the source text by oracle/make_golden.py::dataset_cases) on the
    same files: tensors bit-identical, labels identical, the sample with the missing mask and the corrupt image are
    skipped to the same successor, image_labels.json unpacks to the same lists."""
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import datasets as D
    from PIL import Image
    from torchvision import transforms
    g = golden("datasets.pt")
    img_dir, mask_dir = tmp_path / "images", tmp_path / "images_mask"
    img_dir.mkdir(), mask_dir.mkdir()
    for i, (n, px, m) in enumerate(zip(g["names"], g["pixels"], g["masks"])):
        if i == g["corrupt"]:
            (img_dir / n).write_bytes(b"this is not a PNG file")
        else:
            Image.fromarray(px.numpy()).save(str(img_dir / n))
        if i != g["no_mask"]:
            Image.fromarray(m.numpy()).save(str(mask_dir / D.mask_name(n)))
    (tmp_path / "image_labels.json").write_text(json.dumps({n: D.label_from_filename(n) for n in g["names"]}))
    files, labels = D.load_image_labels(str(tmp_path / "image_labels.json"))
    assert files == g["files"] and labels == g["labels"]
    tf = transforms.Compose([transforms.Resize((32, 32)), transforms.ToTensor()])
    ds1 = D.ImageLabelDataset(str(img_dir), files, labels, transform=tf)
    ds2 = D.ImageMaskLabelDataset(str(img_dir), str(mask_dir), files, labels, transform=tf)
    assert (len(ds1), len(ds2)) == tuple(g["len"])
    for idx, want in zip(g["main_index"], g["main_items"]):
        img, lab = ds1[idx]
        assert torch.equal(img, want["image"]) and lab == want["label"]
    with pytest.raises(OSError):
        ds1[g["corrupt"]]                       # main.py's loader does not skip: PIL's error surfaces, as in the reference
    for idx, want in enumerate(g["main2_items"]):
        img, mask, lab = ds2[idx]
        assert torch.equal(img, want["image"]) and torch.equal(mask, want["mask"]) and lab == want["label"], idx
    assert ds2[g["no_mask"]][2] == ds2[g["corrupt"]][2] == g["labels"][4]     # both fall through to the last sample


def test_iddm_image_writers_equal_reference(golden, tmp_path):
    """The IDDM generator's output files (utils/utils.py:51-89 via tools/generate.py:79-80): one grid image and one
    file per sample from the uint8 batch -- same file names and, decoded, the same pixels as the reference's own
    functions wrote (tests/golden/datasets.pt['writers']); the resized copies carry the size in their name."""
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import datasets as D
    from PIL import Image
    w = golden("datasets.pt")["writers"]
    D.save_image_grid(w["batch"], str(tmp_path / "df.png"))
    D.save_one_image_in_images(w["batch"], str(tmp_path), "df", image_format="png")
    assert sorted(os.listdir(tmp_path)) == w["written"]
    assert torch.equal(torch.from_numpy(np.array(Image.open(tmp_path / "df.png"))), w["grid"])
    got = torch.stack([torch.from_numpy(np.array(Image.open(tmp_path / f"df_{i}.png"))) for i in range(len(w["batch"]))])
    assert torch.equal(got, w["singles"]) and torch.equal(got, w["batch"].permute(0, 2, 3, 1))
    D.save_one_image_in_images(w["batch"][:2], str(tmp_path / "big"), "df", image_size=32, image_format="png")
    assert sorted(os.listdir(tmp_path / "big")) == ["df_0.png", "df_1.png", "df_32_0.png", "df_32_1.png"]
    assert Image.open(tmp_path / "big" / "df_32_1.png").size == (32, 32)


def test_write_image_labels_rule(tmp_path):
    """label_json.py:12-15 names the category by the text before the FIRST underscore."""
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import datasets as D
    d = tmp_path / "figure"
    d.mkdir()
    for n in ("american_bulldog_1.jpg", "Abyssinian_7.png", "pug_3.jpg"):
        (d / n).write_bytes(b"")
    out = D.write_image_labels(str(d), str(tmp_path / "image_labels.json"))
    assert out == {"american_bulldog_1.jpg": "american", "Abyssinian_7.png": "Abyssinian", "pug_3.jpg": "pug"}
    text = (tmp_path / "image_labels.json").read_text()
    assert json.loads(text) == out and text.startswith('{\n    "')        # indent=4 like the reference writes it
    files, labels = D.load_image_labels(str(tmp_path / "image_labels.json"))
    assert dict(zip(files, labels)) == out
    assert D.label_from_filename("american_bulldog_1.jpg") == "american_bulldog"     # the decision rule differs on purpose


def test_iddm_trainer_dataset_equals_reference(golden, tmp_path):
    """datasets.ImageMaskLabelPathDataset against the reference's own utils/utils_shadow.py:252-276 class (run from the
    source text): RGB masks through the same Normalize(0.5) transform -> 3-channel masks in [-1, 1], 4-tuples with path."""
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import datasets as D
    from PIL import Image
    from torchvision import transforms
    g = golden("datasets.pt")
    (tmp_path / "images").mkdir(), (tmp_path / "images_mask").mkdir()
    for n, px, m in zip(g["names"][:2], g["pixels"][:2], g["masks"][:2]):
        Image.fromarray(px.numpy()).save(str(tmp_path / "images" / n))
        Image.fromarray(m.numpy()).save(str(tmp_path / "images_mask" / D.mask_name(n)))
    tf = transforms.Compose([transforms.Resize((24, 24)), transforms.ToTensor(),
                             transforms.Normalize(mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5))])
    ds = D.ImageMaskLabelPathDataset(str(tmp_path / "images"), str(tmp_path / "images_mask"), g["names"][:2], [5, 9], transform=tf)
    assert len(ds) == 2
    for i, want in enumerate(g["iddm_items"]):
        img, mask, lab, path = ds[i]
        assert torch.equal(img, want["image"]) and torch.equal(mask, want["mask"]) and lab == want["label"] and path == want["path"]
        assert mask.shape[0] == 3 and float(mask.min()) == -1.0 and float(mask.max()) == 1.0
    with pytest.raises(OSError):
        D.ImageMaskLabelPathDataset(str(tmp_path / "images"), str(tmp_path / "images_mask"), ["missing_1.png"], [0])[0]
