#!/usr/bin/env python
"""bench.py -- shadowed images/sec (DDIM-50, 256x256, batch 64, bf16) on N B200s of one node.

  python bench.py [--gpus N --steps K --warmup W]            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                       the reference's own CPU path (oracle port)

A "step" = one full batch trajectory: 64 noise tensors -> 50 UNet evaluations + fused DDIM updates
(CUDA-graph replays) -> fused shadow composite.  Workload = BASELINE.json configs[1]
(ddim2/diff_model2.py UNet defaults, random init seed 0, synthetic data).  See DESIGN.md section 'Measurement'.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

# profiles/r01c_summary.md / prof_r01c_key_metrics.csv: mean dram read+write of the 21 conv launches of the --set full
# capture (32-image sub-batch, deep UNet levels; e.g. the captured 256->256 3x3 @128^2 launch moved 782 MB against
# 805 MB algorithmic incl. its residual) -- no re-reads beyond the algorithmic traffic
NCU_CONV_TRAFFIC_BYTES_PER_LAUNCH = 0.312e9
METRIC = "shadowed images/sec (DDIM-50, 256x256)"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("ADVS_BENCH_STREAMS", "2")),
                    help="independent sub-batches on separate CUDA streams inside ShadowSampler")
    return ap.parse_args()


def workload(args):
    return {"workload": f"configs[1]: diff_model2.UNetModel() (330M params, random init seed 0), "
                        f"{args.size}x{args.size}, DDIM-{args.ddim_steps} eta=0, batch {args.batch}/GPU, "
                        f"+ fused shadow composite",
            "batch_per_gpu": args.batch, "image_size": args.size, "ddim_steps": args.ddim_steps,
            "precision": args.precision, "streams": args.streams, "l2": "per-forward activations (GBs) exceed the 126 MB L2",
            "parallelism": f"dp{args.gpus} (images sharded, no collective in the step loop)"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, gpu):
        self.lines, self.proc = [], None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(gpu)], stdout=subprocess.PIPE, text=True)
            self.th = threading.Thread(target=self._pump, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for ln in self.proc.stdout:
            self.lines.append(ln.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        self.th.join(timeout=2)
        sm, mx, reasons = [], [], set()
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 6:
                continue
            try:
                sm.append(float(f[0]))
                mx.append(float(f[1]))
            except ValueError:
                continue
            for name, v in zip(self.NAMES, f[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        sm.sort()
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_reference_rate(args, n_forwards, warm):
    """The reference's CPU path (oracle/torch_port.py: plain PyTorch fp32, all host threads) on a bounded
    sample: 1 image, a few of the 50 DDIM steps at full resolution; per-step cost is t-independent, so
    images/s = 1 / (ddim_steps * seconds per step)."""
    import torch
    from oracle import torch_port as P
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import diff_model2
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    model = diff_model2.UNetModel().eval()        # parameter container only (same seeded weights)
    p = {k: v.detach() for k, v in model.state_dict().items()}
    acp = P.linear_alphas_cumprod()
    torch.manual_seed(1234)
    x = torch.randn(1, 3, args.size, args.size)
    times = []
    with torch.no_grad():
        for i in range(warm + n_forwards):
            t0 = time.perf_counter()
            x = P.ddim_sample(p, P.DM2_CFG, acp, x, args.ddim_steps, max_steps=1)
            dt = time.perf_counter() - t0
            if i >= warm:
                times.append(dt)
    per_step = sum(times) / len(times)
    return 1.0 / (args.ddim_steps * per_step), per_step, torch.get_num_threads()


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, per_step, cores = cpu_reference_rate(args, max(args.steps, 1), max(args.warmup, 1))
    sample = (f"1 image x 1 DDIM step (UNet fwd + update) at {args.size}x{args.size} per bench step, "
              f"{args.steps} timed; images/s = 1/({args.ddim_steps} x {per_step:.2f} s)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": workload(args),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    import torch
    import torch.distributed as dist
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import diff_model2, ops
    from advshadow_b200.sampler import ShadowSampler
    from advshadow_b200.attack import exchange_success

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    B, S, n = args.batch, args.size, args.ddim_steps

    torch.manual_seed(0)
    model = diff_model2.UNetModel().eval().to(dev)
    gd = diff_model2.GaussianDiffusion(timesteps=1000)
    sampler = ShadowSampler(model, gd, B, S, ddim_timesteps=n, precision=args.precision, streams=args.streams)

    # synthetic batch (pinned host copies for the end-to-end leg)
    g = torch.Generator().manual_seed(1234 + rank)
    x_T = torch.randn(B, 3, S, S, generator=g).pin_memory()
    clean = torch.rand(B, 3, S, S, generator=g).pin_memory()
    yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
    cen = torch.rand(B, 2, generator=g) * (S / 2) + S / 4
    rad = torch.rand(B, generator=g) * (S / 8) + S / 8
    fmask = (((xx[None] - cen[:, 0, None, None]) ** 2 + (yy[None] - cen[:, 1, None, None]) ** 2)
             <= (1.5 * rad[:, None, None]) ** 2).float()[:, None].contiguous().pin_memory()
    cen, rad = cen.pin_memory(), rad.pin_memory()
    out_host = torch.empty(B, 3, S, S).pin_memory()
    labels = torch.randint(0, 37, (B,), generator=g).to(dev)
    logits = torch.randn(B, 37, generator=g).to(dev)
    x_T_dev = x_T.to(dev)
    sampler.set_inputs(x_T_dev, clean, fmask, cen, rad)
    torch.cuda.synchronize()

    def exchange():
        # the path's only collective: per-image success flags + ASR counts, once per batch
        # (synthetic victim logits: the victim network itself stays PyTorch and is not part of this metric)
        flags, _ = ops.success_flags(logits, labels)
        return exchange_success(flags, pad_to=B)[1]

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def device_step():
        sampler.load_x_T(x_T_dev)
        sampler.run_device()
        exchange()

    def e2e_step():
        sampler(x_T, clean, fmask, cen, rad, out_host=out_host)
        exchange()

    def timed(fn, k):
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        device_step()
    clocks = ClockSampler(local)
    ms = timed(device_step, args.steps)
    clk = clocks.stop()
    value = world * B * args.steps / (ms / 1e3)

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in (x_T, clean, fmask, cen, rad))
    d2h = out_host.numel() * 4

    out = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        peak_bw = peaks.get("hbm_gbs", 6650.0)
        prof = sampler.eng.profile_forward(repeats=2)
        dom = max(prof.items(), key=lambda kv: kv[1]["ms"])[0]
        conv = prof.get("conv_sm100") or prof[dom]
        ach = conv["flops"] / (conv["ms"] * 1e-3) / 1e12
        fwd_ms = sum(d["ms"] for d in prof.values())
        breakdown = {k: {"ms": round(d["ms"], 3), "share": round(d["ms"] / fwd_ms, 4),
                         "tflops": round(d["flops"] / (d["ms"] * 1e-3) / 1e12, 1) if d["flops"] else None,
                         "gbs": round(d["bytes"] / (d["ms"] * 1e-3) / 1e9, 1) if d["bytes"] else None,
                         "launches": round(d["launches"])}
                     for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        flops_per_img = sampler.eng.plan.flops / sampler.eng.B * n
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic", "config": workload(args),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": sampler.launches_per_trajectory * args.steps + args.steps,
            "roofline": {"kernel": "k_conv_sm100_2cta[_halo] (tcgen05 cta_group::2 implicit-GEMM conv, all launches of one UNet forward)",
                         "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                         # mean dram__bytes_read+write per captured conv launch, profiles/prof_r01c_key_metrics.csv
                         # (the capture covers 21 deep-level launches; "algorithmic_bytes_per_launch" averages all 105)
                         "traffic": NCU_CONV_TRAFFIC_BYTES_PER_LAUNCH,
                         "algorithmic_bytes_per_launch": conv["bytes"] / max(conv["launches"], 1),
                         "peak_source": peak_src},
            "whole_path_tensor_frac": value / world * flops_per_img / 1e12 / peak_tf,
            "forward_breakdown": breakdown, "hbm_peak_gbs": peak_bw,
        }
        if world == 1 and not args.no_cpu_baseline:
            rate, per_step, cores = cpu_reference_rate(args, 2, 1)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"1 image x 2 of {n} DDIM steps at {S}x{S} (oracle/torch_port.py, fp32), "
                                             f"{per_step:.2f} s/step, extrapolated x{n}"}
        print(json.dumps(out))
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
