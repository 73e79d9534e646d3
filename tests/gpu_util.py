"""Thin torch-tensor wrappers over the C-ABI single-operator entry points (tests only)."""
import torch

from diffusionmodelscustom_b200 import _native as N


def stream():
    return torch.cuda.current_stream().cuda_stream


def rel_l2(a, b):
    a = torch.as_tensor(a).double().cpu()
    b = torch.as_tensor(b).double().cpu()
    r = float((a - b).norm() / b.norm().clamp_min(1e-30))
    return r if r == r else float("inf")   # NaN (unwritten / poisoned output) is an infinite error


def nhwc_f16(x_nchw):
    return x_nchw.permute(0, 2, 3, 1).contiguous().to(torch.float16).cuda()


def pack_conv_weight(w):           # [Cout,Cin,R,S] -> f16 [Cout][(r*S+s)*Cin+ci]
    return w.permute(0, 2, 3, 1).reshape(w.shape[0], -1).contiguous().to(torch.float16).cuda()


def pack_convt_weight(w):          # [Cin,Cout,2,2] -> f16 [(a*2+b)*Cout+co][ci]
    return w.permute(2, 3, 1, 0).reshape(4 * w.shape[1], w.shape[0]).contiguous().to(torch.float16).cuda()


def conv2d(x_nhwc, w_packed, bias, residual, post_add, B, Hi, Wi, Cin, Cout, R, stride, pad, convt=False, act=0, impl=0):
    if convt:
        out = torch.full((B, 2 * Hi, 2 * Wi, Cout), float("nan"), dtype=torch.float16, device="cuda")
    else:
        Ho = (Hi + 2 * pad - R) // stride + 1
        out = torch.full((B, Ho, Ho, Cout), float("nan"), dtype=torch.float16, device="cuda")  # poison: stale memory must not pass
    N.check(N.lib().b2d_op_conv2d(x_nhwc.data_ptr(), w_packed.data_ptr(), N.ptr(bias), N.ptr(residual), N.ptr(post_add),
                                  0 if post_add is None else post_add.shape[1], out.data_ptr(), B, Hi, Wi, Cin, Cout, R, R,
                                  stride, pad, int(convt), act, impl, stream()))
    torch.cuda.synchronize()
    return out


def to_nchw_f32(y_nhwc):
    return y_nhwc.float().permute(0, 3, 1, 2).contiguous().cpu()
