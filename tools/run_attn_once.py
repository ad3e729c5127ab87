import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, advshadow_b200
from advshadow_b200 import _capi as capi
T, dh, B, H = int(sys.argv[1]), int(sys.argv[2]), 16, 4
q = torch.randn(B, H, T, dh, device="cuda").to(torch.bfloat16) * 0.3
k = torch.randn(B, H, T, dh, device="cuda").to(torch.bfloat16) * 0.3
vt = torch.randn(B, H, dh, T, device="cuda").to(torch.bfloat16)
o = torch.empty(B, T, H * dh, dtype=torch.bfloat16, device="cuda")
pb = capi.PlanBuffer(capi.ATTN_PLAN_BYTES)
capi.call("advs_attention_sm100_plan", q.data_ptr(), k.data_ptr(), vt.data_ptr(), o.data_ptr(), B, H, T, dh, pb.ptr)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(3):
    capi.call("advs_attention_sm100_launch", pb.ptr, st)
torch.cuda.synchronize()
print("ok")
