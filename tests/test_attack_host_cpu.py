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
