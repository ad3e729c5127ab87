"""Does running the level-0 chain  conv(+stats) -> GN finalize -> GN apply -> conv  in image chunks small enough
for the 126 MB L2 beat one pass over the whole sub-batch?  (Every op on the path is per-image independent.)"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import advshadow_b200  # noqa
from advshadow_b200 import _capi as capi, ops


def build(Bt, chunk, H, W, Cc):
    lib = capi.lib()
    x = torch.randn(Bt, H, W, Cc, device="cuda").to(torch.bfloat16)
    y1, y2, y3 = (torch.empty_like(x) for _ in range(3))
    w = ops.pack_conv_weight(torch.randn(Cc, Cc, 3, 3, device="cuda") / 30, torch.bfloat16)
    bias = torch.zeros(Cc, device="cuda")
    gamma, beta = torch.ones(Cc, device="cuda"), torch.zeros(Cc, device="cuda")
    parts = int(lib.advs_conv_sm100_stats_parts(chunk, H, W))
    part = torch.empty(Bt, parts, Cc, 2, device="cuda")
    ss = torch.empty(Bt, Cc, 2, device="cuda")
    keep = [x, y1, y2, y3, w, bias, gamma, beta, part, ss]
    L = []
    img = H * W * Cc * 2
    for c0 in range(0, Bt, chunk):
        def conv(src, dst, stats):
            cp = capi.ConvParams()
            cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = chunk, H, W, Cc, 1, 1
            cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = src.data_ptr() + c0 * img, w.data_ptr(), Cc, 9
            cp.bias, cp.out_mode, cp.y, cp.dtype = bias.data_ptr(), 0, dst.data_ptr() + c0 * img, capi.BF16
            if stats:
                cp.stats_partial = part.data_ptr() + c0 * parts * Cc * 8
            pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
            capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
            keep.extend([cp, pb])
            return (lib.advs_conv_sm100_launch, (pb.ptr,))
        L.append(conv(x, y1, True))
        L.append((lib.advs_groupnorm_finalize, (part.data_ptr() + c0 * parts * Cc * 8, Cc, parts, None, 0, 0, chunk, H * W, 32,
                                                1e-5, gamma.data_ptr(), beta.data_ptr(), ss.data_ptr() + c0 * Cc * 8)))
        L.append((lib.advs_groupnorm_apply, (y1.data_ptr() + c0 * img, Cc, None, 0, chunk, H * W, ss.data_ptr() + c0 * Cc * 8,
                                             1, y2.data_ptr() + c0 * img, capi.BF16)))
        L.append(conv(y2, y3, False))
    return L, keep


def main():
    H = W = 256
    Cc = 128
    Bt = 32
    for chunk in (32, 8, 4, 2, 1):
        L, keep = build(Bt, chunk, H, W, Cc)

        def run():
            st = C.c_void_p(torch.cuda.current_stream().cuda_stream)     # the capture stream inside torch.cuda.graph
            for fn, args in L:
                rc = fn(*args, st)
                assert rc == 0, capi.lib().advs_last_error()
        run()
        torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            run()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(20):
            g.replay()
        n = 400
        e0.record()
        for _ in range(n):
            g.replay()
        e1.record()
        torch.cuda.synchronize()
        print(f"chunk {chunk:3d}: {e0.elapsed_time(e1) / n:.3f} ms per {Bt} images", flush=True)
        del g, L, keep


if __name__ == "__main__":
    main()
