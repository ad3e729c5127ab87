"""Sustained (power-capped) throughput of single kernels on the B200: each case loops for ~3 s.
Reports TFLOP/s next to a cuBLAS bf16 GEMM looped the same way, plus median SM clock / power."""
import ctypes as C
import json
import subprocess
import sys
import threading
import time
import os

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import advshadow_b200  # noqa
from advshadow_b200 import _capi as capi, ops


class Smi:
    def __init__(self):
        self.lines = []
        self.p = subprocess.Popen(["nvidia-smi", "--query-gpu=clocks.sm,power.draw", "--format=csv,noheader,nounits",
                                   "-lms", "100", "-i", "0"], stdout=subprocess.PIPE, text=True)
        threading.Thread(target=self._pump, daemon=True).start()

    def _pump(self):
        for ln in self.p.stdout:
            self.lines.append((time.time(), ln.strip()))

    def window(self, t0, t1):
        v = [ln.split(",") for (t, ln) in self.lines if t0 + 0.5 <= t <= t1]
        if not v:
            return None, None
        clk = sorted(float(x[0]) for x in v)
        pw = sorted(float(x[1]) for x in v)
        return clk[len(clk) // 2], pw[len(pw) // 2]


def loop(fn, seconds=3.0):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    # calibrate
    e0.record()
    for _ in range(5):
        fn()
    e1.record()
    torch.cuda.synchronize()
    per = e0.elapsed_time(e1) / 5 / 1e3
    n = max(10, int(seconds / per))
    t0 = time.time()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t1 = time.time()
    return e0.elapsed_time(e1) / n / 1e3, t0, t1


def conv_case(B, H, W, cin, cout, taps, stats=0):
    x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
    k = 3 if taps == 9 else 1
    w = ops.pack_conv_weight(torch.randn(cout, cin, k, k, device="cuda") / 30, torch.bfloat16)
    y = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device="cuda")
    bias = torch.zeros(cout, device="cuda")
    cp = capi.ConvParams()
    cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, cout, 1, 1
    cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = x.data_ptr(), w.data_ptr(), cin, taps
    cp.bias, cp.out_mode, cp.y, cp.dtype = bias.data_ptr(), 0, y.data_ptr(), capi.BF16
    keep = [x, w, y, bias]
    if stats:
        parts = capi.lib().advs_conv_sm100_stats_parts(B, H, W)
        part = torch.empty(B, parts, cout // stats, 2, device="cuda")
        cp.stats_partial, cp.stats_gran = part.data_ptr(), stats
        keep.append(part)
    pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
    capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    lib = capi.lib()
    flops = 2.0 * B * H * W * cin * taps * cout
    return (lambda: lib.advs_conv_sm100_launch(pb.ptr, st)), flops, keep


def main():
    smi = Smi()
    out = []
    a = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    b = torch.randn(8192, 8192, device="cuda", dtype=torch.bfloat16)
    per, t0, t1 = loop(lambda: torch.matmul(a, b))
    clk, pw = smi.window(t0, t1)
    out.append(dict(case="cublas bf16 8192^3", tflops=2 * 8192 ** 3 / per / 1e12, sm_mhz=clk, watts=pw))
    print(out[-1], flush=True)
    cases = [("128->128 3x3 @256^2 B64", (64, 256, 256, 128, 128, 9)), ("256->256 3x3 @128^2 B64", (64, 128, 128, 256, 256, 9)),
             ("512->512 3x3 @64^2 B64", (64, 64, 64, 512, 512, 9)), ("1024->1024 3x3 @32^2 B64", (64, 32, 32, 1024, 1024, 9)),
             ("512->512 1x1 @64^2 B64", (64, 64, 64, 512, 512, 1))]
    for name, c in cases:
        for stats in (0, 1, 4):
            fn, flops, keep = conv_case(*c, stats=stats)
            per, t0, t1 = loop(fn)
            clk, pw = smi.window(t0, t1)
            out.append(dict(case=name + ({0: "", 1: " +gn-stats per channel", 4: " +gn-stats per 4 channels"}[stats]), tflops=flops / per / 1e12, ms=per * 1e3, sm_mhz=clk, watts=pw))
            print(out[-1], flush=True)
            del keep
    for (T, dh) in ((4096, 128), (1024, 256)):
        q = torch.randn(64, 4, T, dh, device="cuda").to(torch.bfloat16) * 0.3
        k = torch.randn(64, 4, T, dh, device="cuda").to(torch.bfloat16) * 0.3
        vt = torch.randn(64, 4, dh, T, device="cuda").to(torch.bfloat16)
        o = torch.empty(64, T, 4 * dh, dtype=torch.bfloat16, device="cuda")
        pb = capi.PlanBuffer(capi.ATTN_PLAN_BYTES)
        capi.call("advs_attention_sm100_plan", q.data_ptr(), k.data_ptr(), vt.data_ptr(), o.data_ptr(), 64, 4, T, dh, pb.ptr)
        st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
        lib = capi.lib()
        per, t0, t1 = loop(lambda: lib.advs_attention_sm100_launch(pb.ptr, st))
        clk, pw = smi.window(t0, t1)
        out.append(dict(case=f"attention T={T} dh={dh} B64 h4", tflops=4.0 * 64 * T * T * 4 * dh / per / 1e12, ms=per * 1e3, sm_mhz=clk, watts=pw))
        print(out[-1], flush=True)
    json.dump(out, open("gpurun_out/microbench_conv.json", "w"), indent=1)


if __name__ == "__main__":
    main()
