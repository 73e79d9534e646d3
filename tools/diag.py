import math, sys, torch, torch.nn.functional as F
sys.path.insert(0, ".")
from tests import gpu_util as G
from diffusionmodelscustom_b200 import _native as N
torch.manual_seed(0)
def run(B,H,Cin,Cout,R,stride,pad,impl):
    g = torch.Generator().manual_seed(1)
    x = torch.randn(B, Cin, H, H, generator=g).half().float()
    w = (torch.randn(Cout, Cin, R, R, generator=g) / math.sqrt(Cin*R*R)).half().float()
    ref = F.conv2d(x, w, None, stride, pad)
    xd = G.nhwc_f16(x); wd = G.pack_conv_weight(w)
    Ho = (H + 2*pad - R)//stride + 1
    out = torch.full((B, Ho, Ho, Cout), float('nan'), dtype=torch.float16, device='cuda')
    N.check(N.lib().b2d_op_conv2d(xd.data_ptr(), wd.data_ptr(), None, None, None, 0, out.data_ptr(), B, H, H, Cin, Cout, R, R, stride, pad, 0, 0, impl, G.stream()))
    torch.cuda.synchronize()
    o = out.float().permute(0,3,1,2).cpu()
    return o, ref, x, w
for impl in (1,0):
    o, ref, x, w = run(2,16,64,64,1,1,0,impl)
    print("impl",impl,"nan frac",float(torch.isnan(o).float().mean()),"rel",G.rel_l2(torch.nan_to_num(o),ref))
    if impl==0:
        d=(torch.nan_to_num(o)-ref).abs()
        print("per-channel err", d.mean(dim=(0,2,3))[:16])
        print("per-row(h) err", d.mean(dim=(0,1,3)))
        print("per-sample err", d.mean(dim=(1,2,3)))
        # partial K hypotheses: each 16-wide K chunk
        for k in range(4):
            wk=w.clone(); wk[:, :k*16]=0; wk[:, (k+1)*16:]=0
            pk=F.conv2d(x,wk)
            print("chunk",k,"corr with out-residual", G.rel_l2(torch.nan_to_num(o), pk))
        for ks in ([0],[0,1],[0,1,2],[1,2,3],[0,2],[1,3]):
            wk=torch.zeros_like(w)
            for k in ks: wk[:,k*16:(k+1)*16]=w[:,k*16:(k+1)*16]
            print("chunks",ks, G.rel_l2(torch.nan_to_num(o), F.conv2d(x,wk)))
        print("ratio o/ref median", float((torch.nan_to_num(o)/ref).median()))
        print(o[0,:4,0,:4]); print(ref[0,:4,0,:4])
