"""Debug aid: per-phase clock64 timestamps of one softmax warp and the MMA warp of CTA (0,0) of the attention kernel.
Needs tools/libadvs_trace.so (csrc built with -DADVS_ATTN_TRACE; see DESIGN.md 'Flash attention')."""
import ctypes as C, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, advshadow_b200
from advshadow_b200 import _capi as capi
capi.LIB_PATH = os.path.join(ROOT, "tools", "libadvs_trace.so")
lib = capi.lib()
T, dh, B, H = int(sys.argv[1]), int(sys.argv[2]), 16, 4
q = torch.randn(B, H, T, dh, device="cuda").to(torch.bfloat16) * 0.3
k = torch.randn(B, H, T, dh, device="cuda").to(torch.bfloat16) * 0.3
vt = torch.randn(B, H, dh, T, device="cuda").to(torch.bfloat16)
o = torch.empty(B, T, H * dh, dtype=torch.bfloat16, device="cuda")
pb = capi.PlanBuffer(capi.ATTN_PLAN_BYTES)
capi.call("advs_attention_sm100_plan", q.data_ptr(), k.data_ptr(), vt.data_ptr(), o.data_ptr(), B, H, T, dh, pb.ptr)
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for _ in range(2):
    capi.call("advs_attention_sm100_launch", pb.ptr, st)
torch.cuda.synchronize()
buf = (C.c_longlong * (64 * 16))()
lib.advs_debug_attn_trace.restype = C.c_int
assert lib.advs_debug_attn_trace(buf) == 0
tr = [[buf[j * 16 + s] for s in range(16)] for j in range(64)]
names = ["wait S_j + ldtm + release S buffer", "row max + running-max handover", "128 x ex2 + sums, packs, P stores",
         "wait PV_{j-1} / rescale O", "P visible + arrive"]
print("first softmax warp of the even-block set: cycles per phase, own blocks 8, 10, ..; period = two key blocks")
for j in range(8, 22, 2):
    t = tr[j]
    ph = [t[1] - t[0], t[2] - t[1], t[3] - t[2], t[4] - t[3], t[5] - t[4]]
    print(j, dict(zip(names, ph)), "busy", t[5] - t[0], "period(2 blocks)", tr[j + 2][0] - t[0])
print("S-issuing warp, per key block")
for j in range(8, 16):
    t = tr[j]
    print(j, {"wait K_j and S buffer free": t[9] - t[8], "issue 8 MMAs + commits": t[10] - t[9], "period": tr[j + 1][8] - t[8]})
