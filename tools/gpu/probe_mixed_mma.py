"""Hardware probe: which (A format, B format) pairs does tcgen05.mma kind::f16 accept on this GPU?
Each case runs in its own process (an illegal instruction kills the CUDA context)."""
import math
import subprocess
import sys

CASE = r'''
import sys, math, ctypes as C, torch, torch.nn.functional as F
sys.path.insert(0, ".")
import advshadow_b200
from advshadow_b200 import _capi as capi, ops
a_f16, w_f16 = int(sys.argv[1]), int(sys.argv[2])
B, H, W, cin, cout = 2, 16, 16, 128, 128
torch.manual_seed(0)
adt = torch.float16 if a_f16 else torch.bfloat16
wdt = torch.float16 if w_f16 else torch.bfloat16
x = torch.randn(B, H, W, cin, device="cuda").to(adt)
w = (torch.randn(cout, cin, 3, 3, device="cuda") / math.sqrt(cin * 9)).to(wdt).float()
wp = ops.pack_conv_weight(w, wdt)
y = torch.empty(B, H, W, cout, dtype=torch.bfloat16, device="cuda")
cp = capi.ConvParams()
cp.B, cp.H, cp.W, cp.Cout, cp.stride, cp.nseg = B, H, W, cout, 1, 1
cp.seg[0].x, cp.seg[0].w, cp.seg[0].C, cp.seg[0].taps = x.data_ptr(), wp.data_ptr(), cin, 9
cp.out_mode, cp.y, cp.dtype = 0, y.data_ptr(), capi.BF16
cp.operand_f16 = (1 if a_f16 else 0) | (4 if w_f16 else 0)
pb = capi.PlanBuffer(capi.CONV_PLAN_BYTES)
capi.call("advs_conv_sm100_plan", C.byref(cp), pb.ptr)
capi.call("advs_conv_sm100_launch", pb.ptr, C.c_void_p(torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
ref = F.conv2d(x.float().permute(0, 3, 1, 2), w, padding=1)
err = ((y.float().permute(0, 3, 1, 2) - ref).abs().max() / ref.abs().max()).item()
print(f"A={'fp16' if a_f16 else 'bf16'} W={'fp16' if w_f16 else 'bf16'}: rel err {err:.3e}")
'''

for a, w in ((0, 0), (1, 1), (0, 1), (1, 0)):
    r = subprocess.run([sys.executable, "-c", CASE, str(a), str(w)], capture_output=True, text=True)
    tail = (r.stdout.strip().splitlines() or [""])[-1]
    err = [l for l in r.stderr.splitlines() if "rror" in l][-1:] if r.returncode else []
    print(f"case A_f16={a} W_f16={w}: rc={r.returncode} {tail} {err}")
