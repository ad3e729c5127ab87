"""Throughput sweep (BASELINE.json configs[4]): batch x resolution x DDIM steps, one or more GPUs
(launch with torchrun for N > 1; weak scaling = `--batches` is per GPU; a strong-scaling point is the same global
batch divided over the ranks, e.g. `--batches 8` on 8 GPUs against `--batches 64` on one).  One JSON line per point.
Batches above `--engine-batch` (64) are processed as consecutive engine-sized chunks through ONE sampler, like a
caller streaming a large job through the fixed-shape CUDA graph.  `--cpu` adds the host-CPU reference column
(oracle/torch_port.py, fp32, all threads, 1 image x 2 steps per resolution, extrapolated linearly in the step count
-- flagged as such).  `--budget-s` skips points whose estimated time exceeds it (estimated from the points already
measured at that resolution; skipped points are listed).
Timing: CUDA events around `reps` passes after one warm-up pass, max over ranks."""
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
    ap.add_argument("--engine-batch", type=int, default=64)
    ap.add_argument("--streams", type=int, default=1)
    ap.add_argument("--cpu", action="store_true")
    ap.add_argument("--budget-s", type=float, default=90.0)
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
    cpu_rate = {}
    if args.cpu and rank == 0:
        import time
        from oracle import torch_port as P
        cfg = {"dm2": P.DM2_CFG, "dm1": P.DM1_CFG, "main": dict(P.DM1_CFG, attention_resolutions=(2,))}[args.model]
        acp = P.linear_alphas_cumprod() if args.model == "dm2" else P.cosine_alphas_cumprod()
        pcpu = {k: v.detach().cpu() for k, v in model.state_dict().items()}
        torch.set_num_threads(os.cpu_count())
        for S in map(int, args.sizes.split(",")):
            x = torch.randn(1, 3, S, S)
            P.ddim_sample(pcpu, cfg, acp, x, 50, max_steps=1)                 # warm-up
            t0 = time.perf_counter()
            P.ddim_sample(pcpu, cfg, acp, x, 50, max_steps=2)
            cpu_rate[S] = (time.perf_counter() - t0) / 2                      # seconds per image per step
        del pcpu
    sec_per_img_step = {}                                                     # measured, per resolution (largest batch so far)
    for S in map(int, args.sizes.split(",")):
        for B in map(int, args.batches.split(",")):
            for n in map(int, args.steps.split(",")):
                Be = min(B, args.engine_batch)
                chunks = -(-B // Be)
                est = sec_per_img_step.get(S, 0.0) * B * n * (1 + args.reps)
                if est > args.budget_s:
                    if rank == 0:
                        rec = {"model": args.model, "gpus": world, "batch_per_gpu": B, "size": S, "ddim_steps": n, "skipped":
                               f"estimated {est:.0f} s > budget {args.budget_s:.0f} s (chunked: {chunks} x the batch-{Be} time)"}
                        print(json.dumps(rec), flush=True)
                        fout.write(json.dumps(rec) + "\n")
                    continue
                try:
                    sampler = ShadowSampler(model, gd, Be, S, ddim_timesteps=n, precision=args.precision,
                                            streams=args.streams if Be % args.streams == 0 and Be >= 2 * args.streams else 1)
                except torch.cuda.OutOfMemoryError:
                    model.release_engines()
                    torch.cuda.empty_cache()
                    continue
                g = torch.Generator().manual_seed(1234 + rank)
                sampler.set_inputs(torch.randn(Be, 3, S, S, generator=g), torch.rand(Be, 3, S, S, generator=g),
                                   torch.ones(Be, 1, S, S), torch.full((Be, 2), S / 2.0), torch.full((Be,), S / 4.0))
                x_T = torch.randn(Be, 3, S, S, generator=g).to(dev)
                sampler.load_x_T(x_T)
                sampler.run_device()
                torch.cuda.synchronize()
                if world > 1:
                    dist.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.reps * chunks):
                    sampler.load_x_T(x_T)
                    sampler.run_device()
                e1.record()
                torch.cuda.synchronize()
                ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
                if world > 1:
                    dist.all_reduce(ms, op=dist.ReduceOp.MAX)
                if rank == 0:
                    ips = world * Be * chunks * args.reps / (float(ms) / 1e3)
                    flops = sampler.eng.plan.flops / sampler.eng.B * n
                    sec_per_img_step[S] = float(ms) / 1e3 / (Be * chunks * args.reps * n)
                    rec = {"model": args.model, "gpus": world, "batch_per_gpu": B, "engine_batch": Be, "size": S, "ddim_steps": n,
                           "precision": args.precision, "images_per_s": round(ips, 3), "ms_per_trajectory_batch": round(float(ms) / args.reps, 2),
                           "tflops_per_gpu": round(ips / world * flops / 1e12, 1)}
                    if S in cpu_rate:
                        rec["cpu_images_per_s"] = round(1.0 / (cpu_rate[S] * n), 5)
                        rec["cpu_note"] = f"{os.cpu_count()} threads, 1 image x 2 steps, extrapolated x{n}"
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
