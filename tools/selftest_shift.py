import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, advshadow_b200
from advshadow_b200 import _capi as capi
st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
for mode in (0, 1):
    for r0 in (0, 1, 2, 3, 7, 8, 9, 130, 131, 132):
        out = torch.zeros(128, 64, device="cuda")
        capi.call("advs_selftest_umma_row_shift", r0, mode, out.data_ptr(), st)
        torch.cuda.synchronize()
        rows = torch.arange(128, device="cuda").float()[:, None] + r0
        ref = (rows % 256) + torch.arange(64, device="cuda").float()[None] / 64
        ref = ref.to(torch.bfloat16).float()
        ok = torch.equal(out, ref)
        print(f"base_offset_mode={mode} r0={r0}: {'OK' if ok else 'MISMATCH'}  row0 got {out[0,:4].tolist()} want {ref[0,:4].tolist()}")
