"""Data-parallel attack-loop harness: shard images (and shadow candidates) over the GPUs of one box,
sample + composite on each shard, let the PyTorch victim judge, exchange per-image success flags and
ASR counts.

The reference has no such loop in one place; the pieces are compute_asr's decision rule
(ASR_fast.py:101-126: argmax -> int_to_label != filename.rsplit('_', 1)[0]), the id2label JSON maps
(config*.json) and the per-image Python loops of ddim2/main2.py:159-168.  Images and candidates are
independent trajectories, so the path shards with NO collective inside the step loop; the only
exchange is one all_gather(uint8 flags) + one all_reduce(int64[2] counts) per batch.
"""
import json
from typing import Dict, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist
import torch.nn.functional as F


# ---- label plumbing (bit-exact string handling) ----
def label_from_filename(filename: str) -> str:
    """True label = text before the LAST underscore (ASR_fast.py:109), e.g. american_bulldog_12.jpg."""
    return filename.rsplit('_', 1)[0]


def load_id2label(path: str) -> Tuple[Dict[int, str], Dict[str, int]]:
    """config.json / configvit.json format: {"id2label": {"0": "Abyssinian", ...}} (ASR_fast.py:67-75)."""
    with open(path, 'r') as f:
        data = json.load(f)
    id2label = data['id2label']
    label_to_int = {label: int(i) for i, label in id2label.items()}
    int_to_label = {v: k for k, v in label_to_int.items()}
    return int_to_label, label_to_int


def filenames_to_label_ids(filenames: Sequence[str], label_to_int: Dict[str, int]) -> List[int]:
    """Unknown labels map to -1: argmax can never equal it, so such an image always counts as a success,
    exactly like `predicted_label != true_label` does in the reference."""
    return [label_to_int.get(label_from_filename(f), -1) for f in filenames]


# ---- file-based evaluation (ASR_fast.py:90-126) ----
def preprocess_image(image_path: str, size: int = 224) -> torch.Tensor:
    """RGB -> Resize((224, 224)) -> ToTensor, no normalisation: [1,3,224,224] in [0,1] (ASR_fast.py:90-97)."""
    from ._compat import Image, transforms
    image = Image.open(image_path).convert('RGB')
    return transforms.Compose([transforms.Resize((size, size)), transforms.ToTensor()])(image).unsqueeze(0)


@torch.no_grad()
def compute_asr(folder_path: str, model, int_to_label: Dict[int, str], device="cuda", batch_size: int = 64) -> float:
    """Attack success rate of the images in a folder (ASR_fast.py:101-126): an image counts as a success when the
    victim's argmax label differs from the text before the last '_' of its file name.  Same rule and preprocessing
    as the reference, but the victim sees `batch_size` images per call and the decisions are taken on the GPU by
    advs_success_flags.  `model`: any PyTorch classifier returning [N, classes] logits."""
    import os
    from . import ops
    files = sorted(f for f in os.listdir(folder_path) if f.lower().endswith(('png', 'jpg', 'jpeg', 'bmp', 'gif')))
    label_to_int = {v: k for k, v in int_to_label.items()}
    ids = filenames_to_label_ids(files, label_to_int)
    successes = total = 0
    for lo in range(0, len(files), batch_size):
        batch = torch.cat([preprocess_image(os.path.join(folder_path, f)) for f in files[lo:lo + batch_size]]).to(device)
        logits = model(batch).float()
        _, counts = ops.success_flags(logits, torch.tensor(ids[lo:lo + batch_size], device=logits.device))
        successes, total = successes + int(counts[0]), total + int(counts[1])
    return successes / total


# ---- sharding ----
def shard_bounds(n_items: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous slice [lo, hi) of rank `rank`; the first n_items % world ranks get one extra item."""
    base, extra = divmod(n_items, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def fold_candidates(flags: torch.Tensor, candidates: int) -> torch.Tensor:
    """[B*K] per-candidate flags -> [B] per-image flags: an image is broken if any candidate breaks it."""
    if candidates == 1:
        return flags
    return flags.view(-1, candidates).amax(dim=1)


def exchange_success(flags_local: torch.Tensor, group=None, pad_to: Optional[int] = None, counts_local=None):
    """The path's only collective: ONE all_gather of fixed-width uint8 rows + ONE all_reduce(sum) of int64[2]
    {successes, total} per batch, no host synchronisation.  flags_local: uint8 [B_local] on any device
    (NCCL on GPUs, gloo on CPU in the tests).  Returns (flags_all uint8 [sum B_local], counts int64 [2]).
    Shards may be ragged: every rank pads its row to `pad_to` with 255 (default: ceil(B_total / world) as given by
    `shard_bounds`, i.e. the widest shard -- pass it explicitly for any other partition) and the padding is dropped
    after the gather; 255 never is a flag value.  `counts_local`: the int64[2] that advs_success_flags already
    produced on the device (ops.success_flags), used as is instead of being recomputed."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    if counts_local is None:
        counts = torch.stack([flags_local.sum(dtype=torch.int64),
                              torch.tensor(flags_local.numel(), dtype=torch.int64, device=flags_local.device)])
    else:
        counts = counts_local.to(torch.int64).clone()
    if world == 1:
        return flags_local.clone(), counts
    if pad_to is None:
        raise ValueError("exchange_success: pad_to (the widest shard, e.g. ceil(B_total / world)) is required when "
                         "world_size > 1 -- it keeps the exchange at one all_gather without a size handshake")
    if flags_local.numel() > pad_to:
        raise ValueError(f"exchange_success: shard of {flags_local.numel()} flags does not fit pad_to={pad_to}")
    padded = torch.full((pad_to,), 255, dtype=torch.uint8, device=flags_local.device)
    padded[:flags_local.numel()] = flags_local
    gathered = torch.empty(world * pad_to, dtype=torch.uint8, device=flags_local.device)
    dist.all_gather_into_tensor(gathered, padded, group=group)
    dist.all_reduce(counts, group=group)
    return gathered[gathered != 255], counts


def victim_preprocess(images: torch.Tensor, size: int = 224) -> torch.Tensor:
    """Resize((224,224)) + ToTensor with NO normalisation (ASR_fast.py:90-97): [0,1] tensors in, out."""
    if images.shape[-1] == size and images.shape[-2] == size:
        return images
    return F.interpolate(images, size=(size, size), mode="bilinear", antialias=True, align_corners=False)


class AttackLoop:
    """One rank of the sharded attack loop.  `sampler` is a ShadowSampler for this rank's local batch
    (B_local * candidates trajectories); `victim` is any PyTorch classifier returning [N, classes] logits
    (HF models: pass `lambda x: model(x).logits`)."""

    def __init__(self, sampler, victim, candidates: int = 1, victim_size: int = 224, group=None, pad_to: Optional[int] = None):
        """`pad_to`: the widest per-rank image count (needed at world_size > 1 when shards are ragged; defaults to
        this rank's own image count, i.e. even shards)."""
        self.sampler, self.victim, self.K, self.victim_size, self.group = sampler, victim, candidates, victim_size, group
        self.pad_to = pad_to

    @torch.no_grad()
    def step(self, x_T, clean, fmask, centers, radii, labels):
        """All inputs are this rank's shard, already expanded to B_local*K candidate rows
        (candidate = (noise seed, shadow radius) pair).  labels: int64 [B_local*K]."""
        from . import ops
        self.sampler.set_inputs(x_T, clean, fmask, centers, radii)
        shadowed = self.sampler.run_device()
        logits = self.victim(victim_preprocess(shadowed, self.victim_size)).float()
        flags, counts = ops.success_flags(logits, labels.to(logits.device))
        if self.K > 1:                     # an image is broken if any of its candidates breaks it
            flags = fold_candidates(flags, self.K)
            counts = None
        flags_all, counts = exchange_success(flags, self.group, pad_to=self.pad_to or flags.numel(), counts_local=counts)
        return {"shadowed": shadowed, "flags_local": flags, "flags": flags_all, "successes": int(counts[0]),
                "total": int(counts[1]), "asr": float(counts[0]) / max(int(counts[1]), 1)}
