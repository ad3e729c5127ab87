#!/usr/bin/env python
"""16-bit error study on the CPU (no GPU needed): replays the engine's op list with tests/plan_interp.py -- rounding
every buffer where the CUDA engine rounds it -- against the reference's goldens, for the storage / operand policies the
engine can be built with (UNetEngine(wide_prenorm=, gemm_operands=, fp16_levels=)) and for two that the hardware does
not allow (mixed-format MMAs), kept as the bound they give.

    python tools/error_study.py > profiles/error_study_r02.md

Cases: dm2 UNet forward at 64x64 and 128x128 (tests/golden/forwards.pt) and the three teacher-forced steps of the
reference's DDIM-50 run at 256x256 (tests/golden/dm2_256.pt: steps 0, 1, 49 = t 981, 961, 1).
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

import advshadow_b200  # noqa: E402,F401
import plan_interp  # noqa: E402
from advshadow_b200 import diff_model2  # noqa: E402
from advshadow_b200.plan import build_unet_plan  # noqa: E402

VARIANTS = [
    ("round-2 default: int8 extension on levels 0-1, fp16 bounded operands on levels 0-1", dict()),
    ("round-1 numerics: plain bf16 storage and operands", dict(wide_prenorm=0, gemm_operands="bf16")),
    ("extension only (bf16 operands)", dict(wide_prenorm=2, gemm_operands="bf16")),
    ("fp16 operands only (no extension)", dict(wide_prenorm=0, gemm_operands="fp16")),
    ("extension on level 0 only", dict(wide_prenorm=1)),
    ("extension on levels 0-2", dict(wide_prenorm=3)),
    ("fp16 operands on level 0 only", dict(fp16_levels=1)),
    ("fp16 operands on every level", dict(fp16_levels=None)),
    ("NOT EXECUTABLE (mixed-format MMA): every tcgen05 conv's weights fp16", dict(mixed_mma=True, fp16_levels=None)),
    ("fp32 storage and operands (the op list itself)", None),
]


def main():
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    m = diff_model2.UNetModel()
    p = {k: v.detach() for k, v in m.state_dict().items()}
    fw = torch.load(os.path.join(ROOT, "tests", "golden", "forwards.pt"), weights_only=False)
    d = torch.load(os.path.join(ROOT, "tests", "golden", "dm2_256.pt"), weights_only=False)["ddim"]
    torch.manual_seed(d["x_T_seed"])
    x_T = torch.randn(1, 3, 256, 256)
    cases = [("64², t=741", fw["dm2_64"]["x"], fw["dm2_64"]["t"], fw["dm2_64"]["eps"]),
             ("128², t=741", fw["dm2_128"]["x"], fw["dm2_128"]["t"], fw["dm2_128"]["eps"])]
    for j, i in enumerate(d["steps"]):
        x = x_T if d["x"][j] is None else d["x"][j]
        cases.append((f"256², DDIM-50 step {i} (t={int(d['t'][j][0])})", x, d["t"][j], d["eps"][j]))
    print("# 16-bit error study (CPU emulation of the engine's rounding points; `tools/error_study.py`)\n")
    print("`max |eps - eps_reference|` (and the error's standard deviation) of one dm2 UNet forward, reference = the unmodified")
    print("reference modules in fp32 (tests/golden).  The north star's tolerance is 2e-2.  The GPU measures 1.07e-2 / 1.49e-2 at")
    print("steps 1 / 49 with the default policy and 5.1 - 6.8e-3 on the small cases (DESIGN.md section 5).\n")
    print("| policy | " + " | ".join(c[0] for c in cases) + " |")
    print("|---|" + "---|" * len(cases))
    with torch.no_grad():
        for name, kw in VARIANTS:
            cells = []
            for _, x, t, want in cases:
                plan = build_unet_plan(m.spec(), 1, x.shape[2], x.shape[3])
                e = plan_interp.run_plan(plan, p, x, t, **(dict(round_bf16=True, **kw) if kw is not None else {}))
                err = e - want
                cells.append(f"{float(err.abs().max()):.2e} ({float(err.std()):.1e})")
            print(f"| {name} | " + " | ".join(cells) + " |", flush=True)


if __name__ == "__main__":
    main()
