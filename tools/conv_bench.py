"""Back-to-back timing of single conv_tc launches (CUDA events over 50 launches) for representative layer shapes."""
import math, sys, torch
sys.path.insert(0, ".")
from tests import gpu_util as G
from diffusionmodelscustom_b200 import _native as N
shapes = [  # name, B, H, Cin, Cout, R, stride, pad, convt
    ("l2b0.ds 1x1s2 tiny", 64, 16, 64, 128, 1, 2, 0, 0),
    ("l1 3x3 c64 16px", 64, 16, 64, 64, 3, 1, 1, 0),
    ("l4 3x3 c512 2px", 64, 2, 512, 512, 3, 1, 1, 0),
    ("qkv L=1024 c64->192", 64, 32, 64, 192, 1, 1, 0, 0),
    ("out L=1024 c64->64", 64, 32, 64, 64, 1, 1, 0, 0),
    ("conv2 8x8s2", 64, 32, 64, 64, 8, 2, 3, 0),
    ("final.up convT 32->64px", 64, 32, 64, 64, 1, 1, 0, 1),
    ("d3.conv 3x3 c64 32px", 64, 32, 64, 64, 3, 1, 1, 0),
]
s = torch.cuda.current_stream().cuda_stream
for name, B, H, Cin, Cout, R, st, pad, convt in shapes:
    x = torch.randn(B, H, H, Cin, device="cuda").half()
    K = R * R * Cin
    w = (torch.randn((4 * Cout if convt else Cout), K, device="cuda") / math.sqrt(K)).half()
    bias = torch.zeros(Cout, device="cuda")
    Ho = 2 * H if convt else (H + 2 * pad - R) // st + 1
    out = torch.empty(B, Ho, Ho, Cout, device="cuda", dtype=torch.float16)
    def run():
        N.check(N.lib().b2d_op_conv2d(x.data_ptr(), w.data_ptr(), bias.data_ptr(), None, None, 0, out.data_ptr(), B, H, H, Cin, Cout, R, R, st, pad, convt, 0, 0, s))
    for _ in range(5): run()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(50): run()
    e1.record(); torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 20
    M = B * (H if convt else Ho) ** 2; Nn = 4 * Cout if convt else Cout
    print(f"{name:28s} {us:7.2f} us/launch   {2.0*M*Nn*K/us/1e6:8.1f} TFLOP/s")
