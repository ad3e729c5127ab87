"""Sustained throughput of the attention kernel from a given build of the library (compile-time variants)."""
import ctypes as C, sys, os, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, advshadow_b200
from advshadow_b200 import _capi as capi
capi.LIB_PATH = os.path.abspath(sys.argv[1])
lib = capi.lib()
T, dh, B, H = int(sys.argv[2]), int(sys.argv[3]), 64, 4
q = torch.randn(B, H, T, dh, device="cuda").to(torch.bfloat16) * 0.3
k = torch.randn(B, H, T, dh, device="cuda").to(torch.bfloat16) * 0.3
vt = torch.randn(B, H, dh, T, device="cuda").to(torch.bfloat16)
o = torch.empty(B, T, H * dh, dtype=torch.bfloat16, device="cuda")
pb = capi.PlanBuffer(capi.ATTN_PLAN_BYTES)
capi.call("advs_attention_sm100_plan", q.data_ptr(), k.data_ptr(), vt.data_ptr(), o.data_ptr(), B, H, T, dh, pb.ptr)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(20):
    lib.advs_attention_sm100_launch(pb.ptr, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
n = int(float(sys.argv[4]) if len(sys.argv) > 4 else 600)
e0.record()
for _ in range(n):
    lib.advs_attention_sm100_launch(pb.ptr, st)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / n
print(f"{os.path.basename(sys.argv[1])} T={T} dh={dh}: {ms:.3f} ms  {4.0 * B * T * T * H * dh / ms / 1e9:.0f} TFLOP/s", flush=True)
