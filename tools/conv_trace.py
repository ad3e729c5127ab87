"""Debug aid: clock64 stamps of the first epilogue warp and the MMA warp of CTA 0 of the halo conv kernel
(needs tools/libadvs_trace.so: `make trace` in csrc)."""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, advshadow_b200
from advshadow_b200 import _capi as capi
capi.LIB_PATH = os.path.join(ROOT, "tools", "libadvs_trace.so")
lib = capi.lib()
from advshadow_b200 import ops
B, H, W, cin, cout, gran = 16, 256, 256, int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
x = torch.randn(B, H, W, cin, device="cuda").to(torch.bfloat16)
w = ops.pack_conv_weight(torch.randn(cout, cin, 3, 3, device="cuda") / 30, torch.bfloat16)
y = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device="cuda")
res = torch.randn(B, H, W, cout, device="cuda").to(torch.bfloat16)
bias = torch.zeros(cout, device="cuda")
cp = capi.ConvParams()
cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, cout, 1, 1
cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = x.data_ptr(), w.data_ptr(), cin, 9
cp.bias, cp.out_mode, cp.y, cp.dtype = bias.data_ptr(), 0, y.data_ptr(), capi.BF16
if len(sys.argv) > 4 and sys.argv[4] == "res":
    cp.residual = res.data_ptr()
if gran:
    parts = lib.advs_conv_sm100_stats_parts(B, H, W)
    part = torch.empty(B, parts, cout // gran, 2, device="cuda")
    cp.stats_partial, cp.stats_gran = part.data_ptr(), gran
pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    capi.call("advs_conv_sm100_launch", pb.ptr, st)
torch.cuda.synchronize()
buf = (C.c_longlong * (64 * 16))()
lib.advs_debug_conv_trace.restype = C.c_int
assert lib.advs_debug_conv_trace(buf) == 0
tr = [[buf[j * 16 + s] for s in range(16)] for j in range(64)]
print(f"conv {cin}->{cout} 3x3 @256^2, stats gran {gran}: epilogue warp 2 of CTA 0, cycles per phase, tiles 6..13")
names = ["wait accumulator", "tcgen05.ld chunk0", "bias/temb/residual", "stores", "stats", "chunk 1 (all)", "fence+release", "stats barrier", "stats flush"]
for j in range(6, 14):
    t = tr[j]
    ph = [t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4], t[6] - t[5], t[7] - t[6], t[8] - t[7], t[9] - t[8]]
    print(j, dict(zip(names, ph)), "busy", t[9] - t[1], "period", tr[j + 1][0] - t[0])
print("MMA warp: wait for a free accumulator, issue one tile")
for j in range(6, 14):
    t = tr[j]
    print(j, {"wait": t[13] - t[12], "issue": t[14] - t[13], "period": tr[j + 1][12] - t[12]})
