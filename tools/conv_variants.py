"""One layer shape through every conv kernel variant (impl 4 = one tile per CTA, 5 = persistent, 3 = persistent + slab tiling),
with and without the residual / activation / time-projection epilogue.  CUDA events over 30 back-to-back launches.
    python tools/conv_variants.py [B H Cin Cout]"""
import math, sys, torch
sys.path.insert(0, ".")
from diffusionmodelscustom_b200 import _native as N
shapes = [(256, 32, 64, 64), (64, 16, 512, 512), (64, 64, 128, 128), (64, 32, 256, 256)] if len(sys.argv) < 5 else [tuple(int(v) for v in sys.argv[1:5])]
s = torch.cuda.current_stream().cuda_stream
for B, H, Cin, Cout in shapes:
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    K = 9 * Cin
    w = (torch.randn(Cout, K, device="cuda") / math.sqrt(K)).half()
    bias = torch.zeros(Cout, device="cuda")
    res = torch.randn(B, H, H, Cout, device="cuda").half()
    vec = torch.randn(B, Cout, device="cuda")
    out = torch.empty(B, H, H, Cout, device="cuda", dtype=torch.float16)
    for full in (0, 1):
        for impl, name in ((4, "one-tile"), (5, "persistent"), (3, "persistent+slab")):
            def run():
                N.check(N.lib().b2d_op_conv2d(x.data_ptr(), w.data_ptr(), bias.data_ptr(), res.data_ptr() if full else None,
                                              vec.data_ptr() if full else None, Cout, out.data_ptr(), B, H, H, Cin, Cout, 3, 3, 1, 1, 0,
                                              1 if full else 0, impl, s))
            try:
                for _ in range(3): run()
            except Exception as e:
                print(f"B={B} {H}x{H} {Cin}->{Cout} epi={full} {name:16s} n/a ({str(e)[:60]})")
                continue
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(30): run()
            e1.record(); torch.cuda.synchronize()
            us = e0.elapsed_time(e1) / 30 * 1e3
            fl = 2.0 * B * H * H * Cout * K
            print(f"B={B} {H}x{H} {Cin}->{Cout} epi={full} {name:16s} {us:8.2f} us  {fl/us/1e6:8.1f} TFLOP/s  {(B*H*H*(Cin+Cout*(2 if full else 1))*2)/us/1e3:7.1f} GB/s")
