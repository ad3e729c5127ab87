"""Throughput sweep (BASELINE.json configs[4]): batch x resolution x DDIM steps, one or more GPUs
(launch with torchrun for N > 1; weak scaling = `--batches` is per GPU).  One JSON line per point.
Timing: CUDA events around `reps` trajectories after one warm-up trajectory, max over ranks."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="dm2", choices=["dm1", "dm2", "main"])
    ap.add_argument("--batches", default="1,8,64")
    ap.add_argument("--sizes", default="64,128,256")
    ap.add_argument("--steps", default="10,50")
    ap.add_argument("--reps", type=int, default=1)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--out", default="gpurun_out/sweep.jsonl")
    args = ap.parse_args()
    import advshadow_b200  # noqa
    from advshadow_b200 import diff_model, diff_model2
    from advshadow_b200.sampler import ShadowSampler
    world, rank, local = (int(os.environ.get(k, d)) for k, d in (("WORLD_SIZE", "1"), ("RANK", "0"), ("LOCAL_RANK", "0")))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    torch.manual_seed(0)
    if args.model == "dm2":
        model, gd = diff_model2.UNetModel(), diff_model2.GaussianDiffusion()
    elif args.model == "dm1":
        model, gd = diff_model.UNetModel(), diff_model.GaussianDiffusion()
    else:
        model = diff_model.UNetModel(channel_mult=(1, 2, 2, 2), attention_resolutions=(2,), dropout=0.1)
        gd = diff_model.GaussianDiffusion()
    model = model.eval().to(dev)
    os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
    fout = open(args.out, "a") if rank == 0 else None
    for S in map(int, args.sizes.split(",")):
        for B in map(int, args.batches.split(",")):
            for n in map(int, args.steps.split(",")):
                try:
                    sampler = ShadowSampler(model, gd, B, S, ddim_timesteps=n, precision=args.precision)
                except torch.cuda.OutOfMemoryError:
                    model.release_engines()
                    torch.cuda.empty_cache()
                    continue
                g = torch.Generator().manual_seed(1234 + rank)
                sampler.set_inputs(torch.randn(B, 3, S, S, generator=g), torch.rand(B, 3, S, S, generator=g),
                                   torch.ones(B, 1, S, S), torch.full((B, 2), S / 2.0), torch.full((B,), S / 4.0))
                x_T = sampler.eng.x.clone()
                sampler.run_device()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps):
                    sampler.eng.x.copy_(x_T)
                    sampler.run_device()
                e1.record()
                torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                if rank == 0:
                    ips = world * B * args.reps / (float(ms) / 1e3)
                    flops = sampler.eng.plan.flops / B * n
                    rec = {"model": args.model, "gpus": world, "batch_per_gpu": B, "size": S, "ddim_steps": n,
                           "precision": args.precision, "images_per_s": round(ips, 3), "ms_per_trajectory_batch": round(float(ms) / args.reps, 2),
                           "tflops_per_gpu": round(ips / world * flops / 1e12, 1), "arena_gb": round(sampler.eng.plan.arena_bytes / 2 ** 30, 2)}
                    print(json.dumps(rec), flush=True)
                    fout.write(json.dumps(rec) + "\n")
                    fout.flush()
                del sampler
                model.release_engines()
                torch.cuda.empty_cache()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
