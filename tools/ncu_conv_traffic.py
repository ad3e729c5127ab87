"""DRAM traffic of every conv launch of one UNet forward against the algorithmic bytes of THE SAME launch.

  on the GPU box (tools/gpu/ncu_r02.sh):
      python tools/ncu_conv_traffic.py run                      # plain run first (must exit 0), then the same under
      ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed \
          --clock-control none -k regex:k_conv_sm100 -s <2 x launches> -c <launches> --csv --log-file gpurun_out/conv_traffic_r02.csv \
          python tools/ncu_conv_traffic.py run
  here:
      python tools/ncu_conv_traffic.py join                     # -> profiles/conv_traffic_r02.{json,md}

`run` builds the 32-image dm2 engine at 256x256 (one sub-batch of the bench), runs three forwards and writes the
launch-order list of the tcgen05 conv launches with their algorithmic FLOPs / bytes (engine.launch_costs)."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
COSTS = os.path.join(ROOT, "gpurun_out", "conv_costs_r02.json")
CSV = os.path.join(ROOT, "gpurun_out", "conv_traffic_r02.csv")


def run():
    import torch
    import advshadow_b200  # noqa: F401
    from advshadow_b200 import diff_model2
    torch.manual_seed(0)
    model = diff_model2.UNetModel().eval().cuda()
    eng = model.engine(32, 256, 256, precision="bf16")
    costs = [(n, f, b) for (n, f, b) in eng.launch_costs() if n in ("conv_sm100", "stem_sm100", "head_sm100")]
    ops = [op for op in eng.plan.ops if op.kind in ("conv", "upconv", "stem", "head")]
    os.makedirs(os.path.dirname(COSTS), exist_ok=True)
    json.dump({"launches": costs, "n": len(costs)}, open(COSTS, "w"))
    x = torch.randn(32, 3, 256, 256, device="cuda")
    t = torch.full((32,), 500, device="cuda")
    for _ in range(3):            # ncu skips the first two (-s 2n) and captures the third (-c n)
        eng.forward(x, t)
    torch.cuda.synchronize()
    print("conv launches per forward:", len(costs))


def join():
    costs = json.load(open(COSTS))["launches"]
    rows = [r for r in csv.reader(l for l in open(CSV) if not l.startswith("==")) if len(r) > 5]
    hdr = rows[0]
    kn, mn, mv, idc = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value"), hdr.index("ID")
    per = {}
    for r in rows[1:]:
        d = per.setdefault(int(r[idc]), {"kernel": r[kn].split("(")[0].replace("void advs::", "")})
        d[r[mn]] = float(r[mv].replace(",", ""))
    launches = [per[k] for k in sorted(per)]
    assert len(launches) == len(costs), (len(launches), len(costs))
    out, tot_d, tot_a = [], 0.0, 0.0
    for (name, fl, by), m in zip(costs, launches):
        dram = m["dram__bytes_read.sum"] + m["dram__bytes_write.sum"]
        us = m["gpu__time_duration.sum"] / 1e3
        out.append({"launch": name, "kernel": m["kernel"], "us": round(us, 1), "tflops": round(fl / us / 1e6, 1),
                    "tensor_pipe_pct": round(m.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", float("nan")), 1),
                    "dram_mb": round(dram / 1e6, 1), "algorithmic_mb": round(by / 1e6, 1), "ratio": round(dram / by, 3)})
        tot_d += dram
        tot_a += by
    res = {"dram_bytes_per_launch": tot_d / len(out), "algorithmic_bytes_per_launch": tot_a / len(out), "ratio": tot_d / tot_a,
           "n_launches": len(out),
           "note": "ncu dram__bytes_read+write of all %d tcgen05 conv launches of one 32-image dm2 forward at 256x256 "
                   "(profiles/conv_traffic_r02.md), against engine.launch_costs() bytes of the same launches" % len(out),
           "launches": out}
    json.dump(res, open(os.path.join(ROOT, "profiles", "conv_traffic_r02.json"), "w"), indent=1)
    with open(os.path.join(ROOT, "profiles", "conv_traffic_r02.md"), "w") as f:
        f.write("# DRAM traffic vs algorithmic bytes, every tcgen05 conv launch of one 32-image forward (dm2, 256x256)\n\n")
        f.write("ncu `--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active` "
                "`--clock-control none` (cold cache, serialised: un-throttled clocks, so the TFLOP/s here are not the sustained ones).\n\n")
        f.write(f"Sum over {len(out)} launches: DRAM {tot_d / 1e9:.2f} GB vs algorithmic {tot_a / 1e9:.2f} GB -> ratio {tot_d / tot_a:.3f}\n\n")
        f.write("| # | launch | kernel | us | TFLOP/s | tensor pipe % | DRAM MB | algorithmic MB | ratio |\n|---|---|---|---|---|---|---|---|---|\n")
        for i, o in enumerate(out):
            f.write(f"| {i} | {o['launch']} | `{o['kernel']}` | {o['us']} | {o['tflops']} | {o['tensor_pipe_pct']} | {o['dram_mb']} | {o['algorithmic_mb']} | {o['ratio']} |\n")
    print(f"{len(out)} launches, DRAM / algorithmic = {tot_d / tot_a:.3f}")


if __name__ == "__main__":
    {"run": run, "join": join}[sys.argv[1]]()
