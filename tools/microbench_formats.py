"""Sustained throughput of the level-0 / level-1 conv shapes by operand format and epilogue options (round 2):
bf16 vs fp16 operands, with / without the int8 mantissa-extension store, with GroupNorm statistics (the engine's
configuration), each looped ~2.5 s under the power cap.  Also the cuBLAS bf16 vs fp16 GEMM for reference."""
import ctypes as C
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import advshadow_b200  # noqa
from advshadow_b200 import _capi as capi, ops
from tools.microbench_conv import Smi, loop


def conv_case(B, H, W, cin, cout, f16, lo, stats=4, residual=False):
    dt = torch.float16 if f16 else torch.bfloat16
    x = torch.randn(B, H, W, cin, device="cuda").to(dt)
    w = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda") / 30, dt)
    y = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device="cuda")
    bias = torch.zeros(cout, device="cuda")
    cp = capi.ConvParams()
    cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, cout, 1, 1
    cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = x.data_ptr(), w.data_ptr(), cin, 9
    cp.bias, cp.out_mode, cp.y, cp.dtype = bias.data_ptr(), 0, y.data_ptr(), capi.BF16
    cp.operand_f16 = 0b0101 if f16 else 0
    keep = [x, w, y, bias]
    if lo:
        l = torch.empty(B, H, W, cout, dtype=torch.int8, device="cuda")
        cp.y_lo = l.data_ptr()
        keep.append(l)
    if residual:
        r = torch.randn(B, H, W, cout, device="cuda").to(torch.bfloat16)
        cp.residual = r.data_ptr()
        keep.append(r)
    if stats:
        parts = capi.lib().advs_conv_sm100_stats_parts(B, H, W)
        part = torch.empty(B, parts, cout // stats, 2, device="cuda")
        cp.stats_partial, cp.stats_gran = part.data_ptr(), stats
        keep.append(part)
    pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
    capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib = capi.lib()
    return (lambda: lib.advs_conv_sm100_launch(pb.ptr, st)), 2.0 * B * H * W * cin * 9 * cout, keep + [pb, cp]


def main():
    smi = Smi()
    out = []
    for dt in (torch.bfloat16, torch.float16):
        a = torch.randn(8192, 8192, device="cuda", dtype=dt)
        b = torch.randn(8192, 8192, device="cuda", dtype=dt)
        per, t0, t1 = loop(lambda: torch.matmul(a, b), 2.5)
        clk, pw = smi.window(t0, t1)
        out.append(dict(case=f"cublas {dt} 8192^3", tflops=round(2 * 8192 ** 3 / per / 1e12, 1), sm_mhz=clk, watts=pw))
        print(out[-1], flush=True)
    for name, shape in (("128->128 3x3 @256^2 B32", (32, 256, 256, 128, 128)), ("256->256 3x3 @128^2 B32", (32, 128, 128, 256, 256)),
                        ("1024->1024 3x3 @32^2 B32", (32, 32, 32, 1024, 1024))):
        for f16 in (False, True):
            for lo in (False, True):
                for residual in (False, True):
                    fn, flops, keep = conv_case(*shape, f16=f16, lo=lo, residual=residual)
                    per, t0, t1 = loop(fn, 2.5)
                    clk, pw = smi.window(t0, t1)
                    out.append(dict(case=name, operands="fp16" if f16 else "bf16", y_lo=lo, residual=residual,
                                    tflops=round(flops / per / 1e12, 1), us=round(per * 1e6, 1), sm_mhz=clk, watts=pw))
                    print(out[-1], flush=True)
                    del keep
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/microbench_formats.json", "w"), indent=1)


if __name__ == "__main__":
    main()
