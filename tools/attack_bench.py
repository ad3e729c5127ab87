"""BASELINE.json configs[2] / configs[3]: the sharded attack loop at 224x224 against a PyTorch victim.

  torchrun --nproc-per-node N tools/attack_bench.py --victim resnet50 --images 256 [--candidates 8]

UNet = main.py:71-77 configuration, synthetic "ImageNet-shaped" clean images / disk masks / labels, victim =
torchvision resnet50 or vit_b_16 with a 37-class head, random init (no weights are shipped with the reference).
Each rank samples its slice of the images (x K candidates), the victim judges, and the only collective is
attack.exchange_success (uint8 flags + int64[2] counts)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--victim", default="resnet50", choices=["resnet50", "vit_b_16"])
    ap.add_argument("--images", type=int, default=256)
    ap.add_argument("--candidates", type=int, default=1)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--reps", type=int, default=1)
    args = ap.parse_args()
    import torchvision
    import advshadow_b200  # noqa
    from advshadow_b200 import diff_model
    from advshadow_b200.attack import AttackLoop, shard_bounds
    from advshadow_b200.sampler import ShadowSampler
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lo, hi = shard_bounds(args.images, world, rank)
    n_img, K, S = hi - lo, args.candidates, args.size
    torch.manual_seed(0)
    model = diff_model.UNetModel(channel_mult=(1, 2, 2, 2), attention_resolutions=(2,), dropout=0.1).eval().to(dev)
    gd = diff_model.GaussianDiffusion(timesteps=1000)
    victim = getattr(torchvision.models, args.victim)(num_classes=37).eval().to(dev)
    sampler = ShadowSampler(model, gd, n_img * K, S, ddim_timesteps=args.ddim_steps)
    kinds = sorted({nm for (_, _, nm) in sampler.eng._launches})
    loop = AttackLoop(sampler, victim, candidates=K, victim_size=224, pad_to=-(-args.images // world))
    g = torch.Generator().manual_seed(7)          # same global data on every rank, sliced by shard
    clean = torch.rand(args.images, 3, S, S, generator=g)[lo:hi].repeat_interleave(K, 0)
    labels = torch.randint(0, 37, (args.images,), generator=g)[lo:hi].repeat_interleave(K, 0)
    cen = (torch.rand(args.images, 2, generator=g) * (S / 2) + S / 4)[lo:hi].repeat_interleave(K, 0)
    rad_img = (torch.rand(args.images, generator=g) * 50 + 30)[lo:hi]
    rad = (rad_img[:, None] * torch.linspace(0.6, 1.4, K)[None]).reshape(-1) if K > 1 else rad_img   # candidate radii
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    fmask = (((xx[None] - cen[:, 0, None, None]) ** 2 + (yy[None] - cen[:, 1, None, None]) ** 2)
             <= (1.2 * rad.reshape(-1, 1, 1)) ** 2).float()[:, None]
    x_T = torch.randn(n_img * K, 3, S, S, generator=torch.Generator().manual_seed(1234 + rank))
    res = loop.step(x_T, clean, fmask, cen, rad, labels)       # warm-up (graph capture, cuDNN autotune)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.reps):
        res = loop.step(x_T, clean, fmask, cen, rad, labels)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    with torch.no_grad():   # decision rule check against plain torch on this rank's shard
        ref = (victim(res["shadowed"]).argmax(1).cpu() != labels).view(n_img, K).any(1)
    ok = bool(torch.equal(res["flags_local"].bool().cpu(), ref))
    if rank == 0:
        print(json.dumps({"victim": args.victim, "gpus": world, "images": args.images, "candidates": K, "size": S,
                          "ddim_steps": args.ddim_steps, "trajectories_per_s": round(args.images * K * args.reps / (float(ms) / 1e3), 2),
                          "images_per_s": round(args.images * args.reps / (float(ms) / 1e3), 2), "asr": res["asr"],
                          "flags_gathered": int(res["flags"].numel()), "decisions_match_torch": ok,
                          "ms_per_batch": round(float(ms) / args.reps, 1), "kernel_classes": kinds,
                          "arena_gb": round(sampler.eng.plan.arena_bytes / 2 ** 30, 1)}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
