import math, sys, torch
sys.path.insert(0, ".")
from diffusionmodelscustom_b200 import _native as N
B, H, Cin, Cout, R, st, pad = 64, 16, 64, 64, 3, 1, 1
x = torch.randn(B, H, H, Cin, device="cuda").half(); K = R * R * Cin
w = (torch.randn(Cout, K, device="cuda") / math.sqrt(K)).half(); bias = torch.zeros(Cout, device="cuda")
out = torch.empty(B, H, H, Cout, device="cuda", dtype=torch.float16)
s = torch.cuda.current_stream().cuda_stream
for _ in range(6):
    N.check(N.lib().b2d_op_conv2d(x.data_ptr(), w.data_ptr(), bias.data_ptr(), None, None, 0, out.data_ptr(), B, H, H, Cin, Cout, R, R, st, pad, 0, 0, 0, s))
torch.cuda.synchronize(); print("ok")
