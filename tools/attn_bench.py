"""Times b2d_op_attention alone (CUDA events) for the shapes that dominate the step.  B2D_NO_TC_ATTN=1 selects the mma.sync kernel; B2D_ATTN_V=1|2|3|4|6|7|8 (default 8) and B2D_ATTN_POLY=0..5 select the tcgen05 variants."""
import sys, torch
sys.path.insert(0, ".")
from diffusionmodelscustom_b200 import _native as N
shapes = [(64, 1024, 64, 4), (32, 4096, 64, 4)] if len(sys.argv) < 2 else [tuple(int(v) for v in sys.argv[1:5])]
for B, L, C, h in shapes:
    qkv = torch.randn(B, L, 3 * C, device="cuda").half()
    o = torch.empty(B, L, C, device="cuda", dtype=torch.float16)
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        N.check(N.lib().b2d_op_attention(qkv.data_ptr(), o.data_ptr(), B, L, C, h, s))
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10):
        N.check(N.lib().b2d_op_attention(qkv.data_ptr(), o.data_ptr(), B, L, C, h, s))
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 100
    print(f"attention B={B} L={L} C={C} h={h}: {us:.1f} us  {4.0*L*L*C*B/us/1e6:.1f} TFLOP/s  ({L*L*h*B/us/1e6:.2f} Texp/s)")
