"""On-disk formats either side of the sampler: the image / mask / label files the reference's
scripts read and write.  Host-side I/O only (PIL + JSON); nothing here touches the GPU.

  images/<label>_<n>.<ext>                  true label = text before the LAST '_' (ASR_fast.py:109)
  masks/mask_<image file name>              single-channel 0/255 PNG feature mask (mask_for_dataset.py:29,76-80)
  image_labels.json                         {"<file>": "<class name>", ...} (classifer_model.py:36-60, main.py:43-50)
  config*.json                              {"id2label": {"0": "Abyssinian", ...}} per victim (ASR_fast.py:67-75)
  <generate_name>.<fmt>, <generate_name>_<i>.<fmt>   the IDDM generator's outputs: one grid image of the batch and one
                                            file per sample, from uint8 [N,3,S,S] tensors (utils/utils.py:51-89,
                                            tools/generate.py:79-80)
"""
import json
import os
from typing import Dict, List, Sequence, Tuple

from torch.utils.data import Dataset

from ._compat import Image
from .attack import label_from_filename, load_id2label  # noqa: F401  (re-exported: same rules)

IMAGE_EXTENSIONS = ('png', 'jpg', 'jpeg', 'bmp', 'gif')      # ASR_fast.py:105


def load_image_labels(path: str) -> Tuple[List[str], List[str]]:
    """image_labels.json -> (files, labels) in file order, as main.py:47-50 unpacks it."""
    with open(path, 'r') as f:
        image_labels = json.load(f)
    files, labels = zip(*[(k, v) for k, v in image_labels.items()]) if image_labels else ((), ())
    return list(files), list(labels)


def write_image_labels(folder: str, json_path: str) -> Dict[str, str]:
    """label_json.py:7-21: image_labels.json for a folder -- every file name of os.listdir(folder) mapped to the text
    before its FIRST '_' (this file's category rule; decisions use the last-'_' rule of ASR_fast.py:109, and the two
    differ for labels with underscores such as american_bulldog_1.jpg -> "american"), indent 4."""
    image_labels = {name: name.split('_')[0] for name in os.listdir(folder)}
    with open(json_path, 'w') as f:
        json.dump(image_labels, f, indent=4)
    return image_labels


def list_images(folder: str) -> List[str]:
    """The files compute_asr walks (ASR_fast.py:104-105), sorted for reproducible sharding."""
    return sorted(f for f in os.listdir(folder) if f.lower().endswith(IMAGE_EXTENSIONS))


def mask_name(image_name: str) -> str:
    return 'mask_' + image_name                                  # ddim2/main2.py:48, mask_for_dataset.py:29


class ImageLabelDataset(Dataset):
    """main.py:9-29 `CustomDataset`: RGB image + label from parallel lists."""

    def __init__(self, image_dir, image_files: Sequence[str], labels: Sequence, transform=None):
        self.image_dir, self.image_files, self.labels, self.transform = image_dir, list(image_files), list(labels), transform

    def __len__(self):
        return len(self.image_files)

    def __getitem__(self, idx):
        image = Image.open(os.path.join(self.image_dir, self.image_files[idx])).convert('RGB')
        if self.transform:
            image = self.transform(image)
        return image, self.labels[idx]


class ImageMaskLabelDataset(Dataset):
    """ddim2/main2.py:30-66 `CustomDataset`: RGB image, 'L' feature mask `mask_<name>` run through the SAME transform
    (so a bilinear Resize leaves soft mask edges -- they are not re-binarised), label; unreadable samples are skipped
    by moving to the next index, like the reference does."""

    def __init__(self, image_dir, mask_dir, image_files: Sequence[str], labels: Sequence, transform=None):
        self.image_dir, self.mask_dir, self.transform = image_dir, mask_dir, transform
        self.image_files, self.labels = list(image_files), list(labels)

    def __len__(self):
        return len(self.image_files)

    def __getitem__(self, idx):
        for _ in range(len(self.image_files)):
            try:
                name = self.image_files[idx]
                image = Image.open(os.path.join(self.image_dir, name)).convert('RGB')
                mask = Image.open(os.path.join(self.mask_dir, mask_name(name))).convert('L')
                if self.transform:
                    image, mask = self.transform(image), self.transform(mask)
                return image, mask, self.labels[idx]
            except (OSError, FileNotFoundError):
                idx = (idx + 1) % len(self.image_files)
        raise FileNotFoundError("no readable (image, mask) pair in the dataset")


class ImageMaskLabelPathDataset(Dataset):
    """utils/utils_shadow.py:252-276 `CustomDataset` (the IDDM trainer's flavour, ts:417-435): image AND mask are loaded
    as RGB (torchvision's default_loader), both go through the same transform -- with the trainer's
    Normalize(0.5, 0.5) the mask therefore arrives as a 3-channel tensor in [-1, 1], which is the `[3,H,W]`
    feature-mask form apply_shadow accepts -- and the item carries the relative path: (image, mask, label, path).
    No skipping: a missing file raises, as there."""

    def __init__(self, image_dir, mask_dir, image_paths: Sequence[str], labels: Sequence, transform=None):
        self.image_dir, self.mask_dir, self.transform = image_dir, mask_dir, transform
        self.image_paths, self.labels = list(image_paths), list(labels)

    def __len__(self):
        return len(self.image_paths)

    def __getitem__(self, idx):
        rel = self.image_paths[idx]
        image = Image.open(os.path.join(self.image_dir, rel)).convert('RGB')
        mask = Image.open(os.path.join(self.mask_dir, mask_name(rel))).convert('RGB')
        if self.transform:
            image, mask = self.transform(image), self.transform(mask)
        return image, mask, self.labels[idx], rel


def save_images(images, folder: str, names: Sequence[str]):
    """Write [N,3,H,W] tensors in [0,1] as PNG/JPG files named like the inputs (`<label>_<n>.<ext>`), the layout
    compute_asr and classifer_model.py read back."""
    import numpy as np
    os.makedirs(folder, exist_ok=True)
    arr = (images.detach().float().clamp(0, 1).mul(255).round().byte().permute(0, 2, 3, 1).cpu().numpy())
    for a, n in zip(arr, names):
        Image.fromarray(np.ascontiguousarray(a)).save(os.path.join(folder, n))


def save_image_grid(images, path: str, **kwargs):
    """uint8 [N,3,H,W] sampler output -> one grid image (torchvision.utils.make_grid defaults: 8 per row, 2-pixel
    padding) written to `path` (utils/utils.py:51-62)."""
    import numpy as np
    import torchvision
    grid = torchvision.utils.make_grid(tensor=images.cpu(), **kwargs)
    Image.fromarray(np.ascontiguousarray(grid.permute(1, 2, 0).numpy())).save(path)


def save_one_image_in_images(images, path: str, generate_name: str, image_size=None, image_format: str = "jpg", **kwargs):
    """uint8 [N,3,H,W] -> `<generate_name>_<i>.<fmt>` per sample, plus `<generate_name>_<size>_<i>.<fmt>` resized
    copies when `image_size` is given (utils/utils.py:65-89; the reference asks Pillow for `Image.ANTIALIAS`, which
    current Pillow spells LANCZOS -- same filter)."""
    import numpy as np
    import torchvision
    os.makedirs(path, exist_ok=True)
    for count, one in enumerate(images.cpu()):
        grid = torchvision.utils.make_grid(tensor=one, **kwargs)
        im = Image.fromarray(np.ascontiguousarray(grid.permute(1, 2, 0).numpy()))
        im.save(os.path.join(path, f"{generate_name}_{count}.{image_format}"))
        if image_size is not None:
            im.resize(size=(image_size, image_size), resample=getattr(Image, "LANCZOS", 1)).save(
                os.path.join(path, f"{generate_name}_{image_size}_{count}.{image_format}"))
