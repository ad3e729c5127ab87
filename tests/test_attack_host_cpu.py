"""Host logic of the data-parallel attack loop on CPU: label rules, sharding, and the world_size-2
exchange over gloo (the N>1 path of SURVEY 8e)."""
import json
import os

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def test_label_rules(tmp_path):
    import advshadow_b200
    from advshadow_b200 import attack
    assert attack.label_from_filename("american_bulldog_12.jpg") == "american_bulldog"   # last '_' (ASR_fast.py:109)
    assert attack.label_from_filename("Abyssinian_7.png") == "Abyssinian"
    p = tmp_path / "config.json"
    p.write_text(json.dumps({"id2label": {"0": "Abyssinian", "1": "american_bulldog", "2": "Bengal"}}))
    i2l, l2i = attack.load_id2label(str(p))
    assert i2l == {0: "Abyssinian", 1: "american_bulldog", 2: "Bengal"} and l2i["Bengal"] == 2
    assert attack.filenames_to_label_ids(["Bengal_1.jpg", "american_bulldog_3.jpg", "pug_1.jpg"], l2i) == [2, 1, -1]


def test_shard_bounds_cover_everything():
    import advshadow_b200
    from advshadow_b200.attack import shard_bounds
    for n in (0, 1, 7, 256, 257):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_fold_candidates():
    import advshadow_b200
    from advshadow_b200.attack import fold_candidates
    f = torch.tensor([0, 0, 1, 0, 0, 0, 1, 1], dtype=torch.uint8)
    assert fold_candidates(f, 1) is f
    assert fold_candidates(f, 4).tolist() == [1, 1]
    assert fold_candidates(f, 2).tolist() == [0, 1, 0, 1]


def _worker(rank, world, port, n_items, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import advshadow_b200
    from advshadow_b200.attack import exchange_success, shard_bounds
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(n_items, 37, generator=g)
    labels = torch.randint(0, 37, (n_items,), generator=g)
    labels[::3] = logits[::3].argmax(1)
    lo, hi = shard_bounds(n_items, world, rank)
    flags = (logits[lo:hi].argmax(1) != labels[lo:hi]).to(torch.uint8)       # decision rule, ASR_fast.py:117-121
    pad_to = -(-n_items // world)                      # the widest shard
    allf, counts = exchange_success(flags, pad_to=pad_to)
    ref = (logits.argmax(1) != labels).to(torch.uint8)
    ok = bool(torch.equal(allf, ref)) and counts.tolist() == [int(ref.sum()), n_items]
    # counts handed over from the device kernel (here: computed on the host) are used as they are
    mine = torch.tensor([int(flags.sum()), flags.numel()], dtype=torch.int64)
    allf2, counts2 = exchange_success(flags, pad_to=pad_to, counts_local=mine)
    ok = ok and bool(torch.equal(allf2, ref)) and counts2.tolist() == counts.tolist() and mine.tolist() == [int(flags.sum()), flags.numel()]
    try:
        exchange_success(flags)                        # no size handshake: the width must be given
        ok = False
    except ValueError:
        pass
    q.put((rank, ok))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n_items", [64, 37])        # even and ragged shards
def test_exchange_world_size_2_gloo(n_items):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + n_items
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n_items, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
    assert sorted(res) == [(0, True), (1, True)]


def test_asr_rule_and_preprocessing_equal_reference_compute_asr(golden, tmp_path):
    """SURVEY a17 against the reference's OWN compute_asr / preprocess_image (ASR_fast.py:90-126, executed from the
    source text by oracle/make_golden.py::asr_cases): the same PNG files are written again, attack.preprocess_image
    gives the reference's tensors, and the label plumbing (file-name rule, the reference's config.json id2label map,
    unknown labels) turns the reference's logits into the reference's success count.  The GPU half of the rule
    (advs_success_flags == torch.max) is tests/test_gpu_kernels.py's."""
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import attack
    from PIL import Image
    g = golden("asr.pt")
    for n, px in zip(g["names"], g["pixels"]):
        Image.fromarray(px.numpy()).save(str(tmp_path / n))
    (tmp_path / "notes.txt").write_text("not an image")
    pre = [attack.preprocess_image(str(tmp_path / n)) for n in g["names"]]
    assert all(p.shape == (1, 3, 224, 224) for p in pre)
    assert torch.equal(torch.stack([p[0][:, ::16, ::16] for p in pre]), g["pre_probe"])
    assert torch.equal(torch.stack([p[0].double().sum() for p in pre]), g["pre_sum"])
    victim = torch.nn.Sequential(torch.nn.Conv2d(3, 4, 5, stride=4), torch.nn.Tanh(), torch.nn.AdaptiveAvgPool2d(3),
                                 torch.nn.Flatten(), torch.nn.Linear(36, 37)).eval()
    victim.load_state_dict(g["victim_state"])
    with torch.no_grad():
        logits = victim(torch.cat(pre))
    assert torch.allclose(logits, g["logits"], atol=1e-6)             # batched vs one image per call
    cfg = tmp_path / "config.json"
    cfg.write_text(json.dumps({"id2label": g["id2label"]}))
    int_to_label, label_to_int = attack.load_id2label(str(cfg))
    ids = torch.tensor(attack.filenames_to_label_ids(g["names"], label_to_int))
    assert int(ids[g["names"].index("not_a_pet_1.png")]) == -1
    flags = logits.argmax(1) != ids                                   # what advs_success_flags computes on the device
    assert int(flags.sum()) == g["successes"] and len(g["names"]) == g["total"] == 8
    assert int(flags.sum()) / len(g["names"]) == g["asr"]
    with pytest.raises(RuntimeError, match="CUDA"):
        attack.compute_asr(str(tmp_path), victim, int_to_label, device="cpu")     # decisions are a GPU kernel: no CPU path
