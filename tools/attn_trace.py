"""Phase timing of the persistent attention kernel from its clock64 trace (build with B2D_NVCC_DEFINES=-DAT2_TRACE).
    B2D_NVCC_DEFINES=-DAT2_TRACE python -m diffusionmodelscustom_b200.build --force && python tools/attn_trace.py"""
import ctypes as C
import sys

import numpy as np
import torch

sys.path.insert(0, ".")
from diffusionmodelscustom_b200 import _native as N

B, L, Cc, h = 32, 4096, 64, 4
qkv = torch.randn(B, L, 3 * Cc, device="cuda").half()
o = torch.empty(B, L, Cc, device="cuda", dtype=torch.float16)
for _ in range(2):
    N.check(N.lib().b2d_op_attention(qkv.data_ptr(), o.data_ptr(), B, L, Cc, h, torch.cuda.current_stream().cuda_stream))
torch.cuda.synchronize()
buf = np.zeros((256, 12), dtype=np.int64)
L_ = N.lib()
L_.b2d_debug_attn_trace.argtypes = [C.c_void_p, C.c_int32]
L_.b2d_debug_attn_trace(buf.ctypes.data, buf.size)
t = buf[20:200].astype(np.float64)
names = ["start->ld issued(s_full wait)", "max", "decide/fix", "exps", "p_empty wait", "st P + arrive", "ld wait + s_empty arrive"]
print("softmax warp (tile 0), clk per phase, mean over blocks 20..200:")
for k in range(6):
    print(f"  {names[k]:32s} {np.mean(t[:, k + 1] - t[:, k]):8.1f}")
print(f"  {'block total':32s} {np.mean(t[1:, 0] - t[:-1, 0]):8.1f}")
print("MMA issuer (tile 0):")
print(f"  issue_s(n+2)                     {np.mean(t[:, 9] - t[:, 8]):8.1f}")
print(f"  issue_pv waits (v_full, p_full)  {np.mean(t[:, 10] - t[:, 9]):8.1f}")
print(f"  issue_pv MMAs + commits          {np.mean(t[:, 11] - t[:, 10]):8.1f}")
print(f"  iteration total                  {np.mean(t[1:, 8] - t[:-1, 8]):8.1f}")
print("p_full arrive (softmax) -> issuer saw p_full:", np.mean(t[:, 10] - t[:, 5]))
print("issuer PV issued -> softmax passes p_empty of next block:", np.mean(t[1:, 4] - t[:-1, 11]))
