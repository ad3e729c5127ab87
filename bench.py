#!/usr/bin/env python
"""bench.py -- shadowed images/sec (DDIM-50, 256x256, batch 64, bf16) on N B200s of one node.

  python bench.py [--gpus N --steps K --warmup W]            (N>1: launched by torch.distributed.run)
  python bench.py --impl reference ...                       the reference's own CPU path (oracle port)
  python bench.py --impl torch-gpu ...                       informative: the reference algorithm run by PyTorch
                                                             eager (cuDNN/cuBLAS, bf16 autocast) on the same GPU

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

# ncu dram__bytes_read+write per conv launch against the algorithmic bytes of THE SAME launch: written by
# tools/ncu_conv_traffic.py from a capture of one 32-image forward (every conv launch, in launch order)
CONV_TRAFFIC_TABLE = os.path.join(ROOT, "profiles", "conv_traffic_r02.json")
METRIC = "shadowed images/sec (DDIM-50, 256x256)"
UNIT = "images/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "torch-gpu"])
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--size", type=int, default=256)
    ap.add_argument("--ddim-steps", type=int, default=50)
    ap.add_argument("--precision", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-torch-gpu", action="store_true", help="skip the informative PyTorch-eager-on-GPU sample")
    ap.add_argument("--profile-repeats", type=int, default=20)
    ap.add_argument("--graph-scope", default="trajectory", choices=["trajectory", "step"])
    ap.add_argument("--wide-prenorm", type=int, default=2, help="A/B: levels stored with the int8 mantissa extension (0 = off)")
    ap.add_argument("--gemm-operands", default="fp16", choices=["fp16", "bf16"], help="A/B: format of the bounded GEMM operands")
    ap.add_argument("--fp16-levels", type=int, default=2, help="A/B: top-resolution levels whose bounded operands are fp16")
    ap.add_argument("--streams", type=int, default=int(os.environ.get("ADVS_BENCH_STREAMS", "2")),
                    help="independent sub-batches on separate CUDA streams inside ShadowSampler")
    return ap.parse_args()


def workload(args):
    return {"workload": f"configs[1]: diff_model2.UNetModel() (330M params, random init seed 0), "
                        f"{args.size}x{args.size}, DDIM-{args.ddim_steps} eta=0, batch {args.batch}/GPU, "
                        f"+ fused shadow composite",
            "batch_per_gpu": args.batch, "image_size": args.size, "ddim_steps": args.ddim_steps,
            "precision": (f"16-bit: bf16 storage + int8 mantissa extension on the {args.wide_prenorm} top-resolution levels' pre-norm "
                          f"tensors, {args.gemm_operands} bounded GEMM operands on the {args.fp16_levels} top levels, fp32 accumulate") if args.precision == "bf16" else args.precision,
            "streams": args.streams, "graph": f"one CUDA graph per {args.graph_scope}",
            "l2": "per-forward activations (GBs) exceed the 126 MB L2",
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


def reference_workload(args, what, batch, precision, device):
    """The config block of a reference-algorithm arm says what THAT arm ran (a bounded sample of configs[1])."""
    return {"workload": f"configs[1] model and shape (diff_model2.UNetModel(), {args.size}x{args.size}, DDIM-{args.ddim_steps} eta=0) "
                        f"through oracle/torch_port.py on {device}: {what}",
            "batch_per_gpu": batch, "image_size": args.size, "ddim_steps": args.ddim_steps, "precision": precision,
            "streams": 1, "graph": "none (PyTorch eager)", "parallelism": "single process",
            "extrapolation": f"images/s = batch / ({args.ddim_steps} x seconds per timed DDIM step); the per-step cost does not depend on t"}


def sampler_kernel_rates(sampler, repeats=20):
    """GB/s of advs_ddim_step and of the fused last-step kernel (update + mask + composite) on the sampler's own
    buffers, CUDA events on the launching stream; algorithmic bytes: x, eps read + x written = 12 B/elem, the fused
    tail adds the clean image read, the output write and the feature mask (4/3 B/elem)."""
    import ctypes as C
    import torch
    from advshadow_b200 import _capi as capi
    smp = sampler.children[0] if sampler.children else sampler
    eng, st = smp.eng, C.c_void_p(torch.cuda.current_stream().cuda_stream)
    n = smp.n_elems
    step = torch.zeros(1, dtype=torch.int32, device=smp.device)
    out = {}

    def timed(fn):
        for _ in range(3):
            fn()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(repeats):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / repeats

    x_keep = eng.x.clone()
    ms = timed(lambda: capi.call("advs_ddim_step", eng.x.data_ptr(), eng.eps.data_ptr(), None, eng.x.data_ptr(), n,
                                 smp.coef.data_ptr(), step.data_ptr(), 0, smp.clip, st))
    out["ddim_step"] = {"ms": round(ms, 4), "gbs": round(12 * n / (ms * 1e-3) / 1e9, 1), "launches": 1,
                        "note": f"{smp.B} images; working set {12 * n / 2 ** 20:.0f} MiB (L2-resident below 126 MiB)"}
    ms = timed(lambda: capi.call("advs_ddim_step_composite", eng.x.data_ptr(), eng.eps.data_ptr(), eng.x.data_ptr(),
                                 smp.coef.data_ptr(), step.data_ptr(), 0, smp.clip, smp.clean.data_ptr(),
                                 smp.centers.data_ptr(), smp.radii.data_ptr(), smp.fmask.data_ptr(), smp.fmask.shape[1],
                                 smp.blur, smp.out.data_ptr(), smp.B, smp.C, smp.S, smp.S, st))
    byts = 20 * n + 4 * smp.fmask.numel()
    out["ddim_step_composite"] = {"ms": round(ms, 4), "gbs": round(byts / (ms * 1e-3) / 1e9, 1), "launches": 1}
    eng.x.copy_(x_keep)
    return out


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    rate, per_step, cores = cpu_reference_rate(args, max(args.steps, 1), max(args.warmup, 1))
    sample = (f"1 image x 1 DDIM step (UNet fwd + update) at {args.size}x{args.size} per bench step, fp32, "
              f"{args.steps} timed; images/s = 1/({args.ddim_steps} x {per_step:.2f} s)")
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": reference_workload(args, sample, 1, "fp32", f"the host CPU ({cores} threads)"),
        "cpu_baseline": {"value": rate, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def torch_gpu_rate(args, device, batch=16, n_steps=2, warm=1):
    """INFORMATIVE arm ("the bar to beat on B200 is PyTorch-eager cuDNN/cuBLAS running the reference
    modules"): the reference algorithm (oracle/torch_port.py -- nn.functional conv2d / group_norm / einsum-softmax
    attention with the materialised T x T scores) executed by PyTorch on the GPU with bf16 autocast and
    channels_last weights.  None of this repository's kernels is on that path."""
    import torch
    from oracle import torch_port as P
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import diff_model2
    torch.manual_seed(0)
    model = diff_model2.UNetModel().eval()        # parameter container only (same seeded weights)
    p = {}
    for k, v in model.state_dict().items():
        v = v.detach().to(device)
        p[k] = v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v
    del model
    acp = P.linear_alphas_cumprod()
    torch.manual_seed(1234)
    x = torch.randn(batch, 3, args.size, args.size, device=device).contiguous(memory_format=torch.channels_last)
    torch.backends.cudnn.benchmark = True
    times = []
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16):
        for i in range(warm + n_steps):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            x = P.ddim_sample(p, P.DM2_CFG, acp, x, args.ddim_steps, max_steps=1)
            e1.record()
            torch.cuda.synchronize(device)
            if i >= warm:
                times.append(e0.elapsed_time(e1) * 1e-3)
    per_step = sum(times) / len(times)
    peak = torch.cuda.max_memory_allocated(device) / 2 ** 30
    del p, x
    torch.cuda.empty_cache()
    return batch / (args.ddim_steps * per_step), per_step, batch, peak


def run_torch_gpu(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import torch
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    torch.cuda.set_device(dev)
    rate, per_step, batch, peak = torch_gpu_rate(args, dev, n_steps=max(args.steps, 1), warm=max(args.warmup, 1))
    sample = (f"{batch} images x 1 DDIM step per bench step, bf16 autocast + channels_last, {args.steps} timed "
              f"({per_step:.3f} s/step, peak {peak:.1f} GiB); images/s = {batch}/({args.ddim_steps} x {per_step:.3f} s)")
    print(json.dumps({
        "impl": "torch-gpu", "informative": True, "metric": METRIC, "value": rate, "unit": UNIT, "n_gpus": 1,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": per_step * 1e3, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": reference_workload(args, sample, batch, "bf16 autocast", "one B200 (PyTorch eager, cuDNN/cuBLAS)")}))


def main():
    args = parse()
    if args.impl == "reference":
        return run_reference(args)
    if args.impl == "torch-gpu":
        return run_torch_gpu(args)
    import torch
    import torch.distributed as dist
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import diff_model2, ops
    from advshadow_b200.sampler import ShadowSampler
    from advshadow_b200.attack import exchange_success

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    t_start = time.perf_counter()

    def note(what):
        # progress breadcrumbs on stderr (stdout carries the one JSON line): if a multi-rank run ever stalls, the log
        # says in which phase and on which rank
        print(f"[bench rank {rank}/{world} +{time.perf_counter() - t_start:6.1f}s] {what}", file=sys.stderr, flush=True)

    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        # every collective of this run is a few bytes after minutes of equal work per rank; a rank that waits five
        # minutes for its peers is stuck, so fail then rather than after the default ten
        import datetime
        dist.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=300))
        note("process group up")
    B, S, n = args.batch, args.size, args.ddim_steps

    torch.manual_seed(0)
    model = diff_model2.UNetModel().eval().to(dev)
    gd = diff_model2.GaussianDiffusion(timesteps=1000)
    sampler = ShadowSampler(model, gd, B, S, ddim_timesteps=n, precision=args.precision, streams=args.streams,
                            graph_scope=args.graph_scope,
                            engine_options=dict(wide_prenorm=args.wide_prenorm, gemm_operands=args.gemm_operands,
                                                fp16_levels=args.fp16_levels))

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
    note(f"sampler built and captured (pdl={sampler.pdl}), inputs resident")

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

    for i in range(args.warmup):
        device_step()
        if i == 0:
            torch.cuda.synchronize()
            note("first trajectory + exchange done")
    clocks = ClockSampler(local)
    ms = timed(device_step, args.steps)
    clk = clocks.stop()
    value = world * B * args.steps / (ms / 1e3)
    note(f"device-resident leg timed: {ms / args.steps:.1f} ms/step")

    e2e_step()
    ms_e2e = timed(e2e_step, args.steps)
    note(f"end-to-end leg timed: {ms_e2e / args.steps:.1f} ms/step")
    e2e_value = world * B * args.steps / (ms_e2e / 1e3)
    h2d = sum(t.numel() * t.element_size() for t in (x_T, clean, fmask, cen, rad))
    d2h = out_host.numel() * 4
    if world > 1:
        # every collective of the run is behind us: leave the process group NOW, on all ranks together.  What follows
        # is rank 0's per-kernel profiling (seconds of collective-free work); ranks that waited for it inside an NCCL
        # barrier kept their GPUs spinning, and a stall there would also have swallowed the (then unflushed) JSON line.
        dist.barrier()
        torch.cuda.synchronize()
        dist.destroy_process_group()
        note("left the process group")

    out = None
    if rank == 0:
        # safety net: the timed numbers exist from here on.  If the (untimed) profiling / baseline legs below ever stall,
        # print the line with what is known instead of losing the run.
        done = threading.Event()
        n_launches = sampler.launches_per_trajectory * args.steps + args.steps

        def fallback():
            if not done.is_set():
                print(json.dumps({
                    "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                    "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                    "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic", "config": workload(args),
                    "clocks": clk, "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                                           "ms_per_step": ms_e2e / args.steps},
                    "gpu_launches": n_launches,
                    "roofline": None, "note": "per-kernel profiling / baseline legs did not finish within 600 s; timed legs only"}),
                    flush=True)
                os._exit(0)

        guard = threading.Timer(600.0, fallback)
        guard.daemon = True
        guard.start()
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak_tf = peaks.get("bf16_tflops_sustained", 1400.0)
        peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1.4 PFLOP/s sustained (B200_PROFILING.md)"
        peak_bw = peaks.get("hbm_gbs", 6650.0)
        # Per-kernel-class timing of one forward: CUDA events around every launch of the 32-image engine, eagerly
        # queued back to back (no host sync inside a repeat) for `profile_repeats` repeats after 5 warm-up repeats, so
        # the GPU sits at the same power-capped clocks as in the timed trajectories; clocks sampled alongside.
        pclk = ClockSampler(local)
        prof, per_repeat = sampler.eng.profile_forward(repeats=args.profile_repeats, warmup=5, per_repeat="conv_sm100")
        pclk = pclk.stop()
        dom = max(prof.items(), key=lambda kv: kv[1]["ms"])[0]
        conv = prof.get("conv_sm100") or prof[dom]
        conv_tf = sorted(conv["flops"] / (ms * 1e-3) / 1e12 for ms in per_repeat) if per_repeat else []
        ach = conv_tf[len(conv_tf) // 2] if conv_tf else conv["flops"] / (conv["ms"] * 1e-3) / 1e12
        fwd_ms = sum(d["ms"] for d in prof.values())
        breakdown = {k: {"ms": round(d["ms"], 3), "share": round(d["ms"] / fwd_ms, 4),
                         "tflops": round(d["flops"] / (d["ms"] * 1e-3) / 1e12, 1) if d["flops"] else None,
                         "gbs": round(d["bytes"] / (d["ms"] * 1e-3) / 1e9, 1) if d["bytes"] else None,
                         "launches": round(d["launches"])}
                     for k, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])}
        # the two sampler kernels outside the UNet (north star: ">= 70 % of HBM peak in the fused norm/DDIM kernels")
        breakdown.update(sampler_kernel_rates(sampler))
        traffic, traffic_alg, traffic_note = None, None, "no ncu capture table found"
        try:
            tt = json.load(open(CONV_TRAFFIC_TABLE))
            traffic = tt["dram_bytes_per_launch"]
            traffic_alg = tt["algorithmic_bytes_per_launch"]
            traffic_note = tt["note"]
        except Exception:
            pass
        flops_per_img = sampler.eng.plan.flops / sampler.eng.B * n
        out = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16" if args.precision == "bf16" else "f32", "data": "synthetic", "config": workload(args),
            "clocks": clk,
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "ms_per_step": ms_e2e / args.steps},
            "gpu_launches": n_launches,
            "roofline": {"kernel": "k_conv_sm100_2cta[_halo] (tcgen05 cta_group::2 implicit-GEMM conv, all launches of one UNet forward)",
                         "bound": "tensor", "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
                         "achieved_min_median_max": [round(conv_tf[0], 1), round(ach, 1), round(conv_tf[-1], 1)] if conv_tf else None,
                         "repeats": args.profile_repeats, "clocks_during_profile": pclk,
                         # ncu dram bytes and algorithmic bytes averaged over the same captured launches
                         "traffic": traffic, "algorithmic_bytes_per_launch": traffic_alg, "traffic_source": traffic_note,
                         "peak_source": peak_src},
            "whole_path_tensor_frac": value / world * flops_per_img / 1e12 / peak_tf,
            "forward_breakdown": breakdown, "hbm_peak_gbs": peak_bw,
        }
        if world == 1 and not args.no_torch_gpu:
            # informative only: the reference algorithm under PyTorch eager on this GPU (bounded sample)
            del sampler
            model.release_engines()
            torch.cuda.empty_cache()
            try:
                rate, per_step, tb, peak = torch_gpu_rate(args, dev)
                out["torch_gpu_baseline"] = {"value": rate, "unit": UNIT, "kind": "oracle/torch_port.py on cuda, bf16 autocast, channels_last",
                                             "sample": f"{tb} images x 2 of {n} DDIM steps, {per_step:.3f} s/step, peak {peak:.1f} GiB, extrapolated x{n}"}
            except Exception as exc:      # informative arm: never take the bench line down
                out["torch_gpu_baseline"] = {"error": repr(exc)[:200]}
        if world == 1 and not args.no_cpu_baseline:
            rate, per_step, cores = cpu_reference_rate(args, 2, 1)
            out["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                   "sample": f"1 image x 2 of {n} DDIM steps at {S}x{S} (oracle/torch_port.py, fp32), "
                                             f"{per_step:.2f} s/step, extrapolated x{n}"}
        done.set()
        guard.cancel()
        note("untimed profiling / baseline legs done")
        print(json.dumps(out), flush=True)


if __name__ == "__main__":
    main()
