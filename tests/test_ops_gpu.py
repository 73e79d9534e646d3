"""GPU unit tests of every kernel class through the C ABI, against plain torch FP32 on the CPU.

Inputs are rounded to f16 first so the only differences are accumulation order and the f16 rounding of the output
(<= 2^-9 relative per element): tolerance 6e-3 relative L2 for f16-output kernels, exact for the posterior update."""
import math

import pytest
import torch
import torch.nn.functional as F

from diffusionmodelscustom_b200 import _native as N
from tests import gpu_util as G

pytestmark = pytest.mark.gpu
TOL = 1.5e-3


def _bf(x):
    return x.to(torch.float16).float()


CONV_CASES = [
    # name, B, H, Cin, Cout, R, stride, pad
    ("3x3_s1_c64_32px", 2, 32, 64, 64, 3, 1, 1),
    ("3x3_s1_c128_16px", 3, 16, 128, 128, 3, 1, 1),
    ("3x3_s1_c512_2px_tn32", 5, 2, 512, 512, 3, 1, 1),
    ("3x3_s1_c256_1px", 3, 1, 256, 256, 3, 1, 1),
    ("3x3_s2_c64_to128", 2, 16, 64, 128, 3, 2, 1),
    ("3x3_s2_c256_to512_4px", 4, 4, 256, 512, 3, 2, 1),
    ("1x1_s2_downsample", 2, 16, 64, 128, 1, 2, 0),
    ("8x8_s2_conv2", 2, 32, 64, 64, 8, 2, 3),
    ("8x8_s2_conv2_64px", 1, 64, 64, 64, 8, 2, 3),
    ("1x1_linear_qkv", 2, 16, 128, 384, 1, 1, 0),
    ("3x3_s1_c64_64px_rowtile", 1, 64, 64, 64, 3, 1, 1),
    ("3x3_s1_c64_128px", 1, 128, 64, 64, 3, 1, 1),
]


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("case", CONV_CASES, ids=[c[0] for c in CONV_CASES])
def test_conv_matches_torch(case, impl):
    _, B, H, Cin, Cout, R, stride, pad = case
    g = torch.Generator().manual_seed(sum(case[0].encode()))
    x = _bf(torch.randn(B, Cin, H, H, generator=g))
    w = _bf(torch.randn(Cout, Cin, R, R, generator=g) / math.sqrt(Cin * R * R))
    bias = torch.randn(Cout, generator=g)
    ref = F.conv2d(x, w, bias, stride, pad)
    out = G.conv2d(G.nhwc_f16(x), G.pack_conv_weight(w), bias.cuda(), None, None, B, H, H, Cin, Cout, R, stride, pad, impl=impl)
    assert G.rel_l2(G.to_nchw_f32(out), ref) < TOL


@pytest.mark.parametrize("case", [("3x3_big", 2, 32, 64, 64, 3, 1, 1), ("1x1_stream", 70, 32, 64, 192, 1, 1, 0)], ids=["conv", "gemm_stream"])
def test_saturation_counter_fires_in_the_gemm_epilogues(case):
    """Outputs beyond +-65504 are clamped by the converter AND counted (per thread and tile since the epilogues track a running
    maximum instead of testing every pair): zero for an in-range tensor, non-zero once a value leaves the fp16 range."""
    _, B, H, Cin, Cout, R, stride, pad = case
    g = torch.Generator().manual_seed(7)
    x = _bf(torch.randn(B, Cin, H, H, generator=g))
    w = _bf(torch.randn(Cout, Cin, R, R, generator=g) / math.sqrt(Cin * R * R))
    for scale, want_clamp in ((1.0, False), (1.0e5, True)):
        bias = torch.randn(Cout, generator=g) * scale
        N.lib().b2d_saturation_count(1)
        out = G.conv2d(G.nhwc_f16(x), G.pack_conv_weight(w), bias.cuda(), None, None, B, H, H, Cin, Cout, R, stride, pad, impl=0)
        torch.cuda.synchronize()
        cnt = int(N.lib().b2d_saturation_count(0))
        assert (cnt > 0) == want_clamp, (scale, cnt)
        assert torch.isfinite(out.float()).all()
    N.lib().b2d_saturation_count(1)


STREAM_CASES = [  # B, H, Cin, Cout, convt
    (2, 16, 64, 192, False), (3, 16, 128, 384, False), (5, 8, 64, 64, False), (2, 32, 128, 128, False),
    (1, 19 * 0 + 16, 64, 256, False), (3, 8, 64, 64, True), (2, 16, 128, 128, True), (70, 32, 64, 192, False)]


@pytest.mark.parametrize("case", STREAM_CASES, ids=[f"B{c[0]}_H{c[1]}_K{c[2]}_N{c[3]}_{'convT' if c[4] else 'lin'}" for c in STREAM_CASES])
def test_streaming_gemm_matches_torch(case):
    """Persistent streaming GEMM (gemm_stream.cuh) incl. ragged last M tile, two K blocks, several N blocks, ConvT store,
    residual + ReLU + per-sample vector epilogue."""
    B, H, Cin, Cout, convt = case
    g = torch.Generator().manual_seed(B * 1000 + Cin + Cout)
    x = _bf(torch.randn(B, Cin, H, H, generator=g))
    bias = torch.randn(Cout, generator=g)
    if convt:
        w = _bf(torch.randn(Cin, Cout, 2, 2, generator=g) / math.sqrt(Cin))
        ref = F.conv_transpose2d(x, w, bias, stride=2)
        out = G.conv2d(G.nhwc_f16(x), G.pack_convt_weight(w), bias.cuda(), None, None, B, H, H, Cin, Cout, 1, 1, 0, convt=True, impl=2)
    else:
        w = _bf(torch.randn(Cout, Cin, 1, 1, generator=g) / math.sqrt(Cin))
        res = _bf(torch.randn(B, Cout, H, H, generator=g))
        vec = torch.randn(B, Cout, generator=g)
        ref = F.relu(F.conv2d(x, w, bias) + res) + vec[:, :, None, None]
        out = G.conv2d(G.nhwc_f16(x), G.pack_conv_weight(w), bias.cuda(), G.nhwc_f16(res), vec.cuda(), B, H, H, Cin, Cout, 1, 1, 0,
                       act=1, impl=2)
    assert G.rel_l2(G.to_nchw_f32(out), ref) < TOL


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
def test_conv_full_epilogue(impl):
    """bias -> +residual -> ReLU -> +per-sample channel vector (encoder stage tail) and GELU variant."""
    B, H, C = 3, 8, 128
    g = torch.Generator().manual_seed(5)
    x = _bf(torch.randn(B, C, H, H, generator=g))
    w = _bf(torch.randn(C, C, 3, 3, generator=g) / math.sqrt(9 * C))
    bias = torch.randn(C, generator=g)
    res = _bf(torch.randn(B, C, H, H, generator=g))
    vec = torch.randn(B, 200, generator=g)   # stride 200, first C used
    ref = F.relu(F.conv2d(x, w, bias, 1, 1) + res) + vec[:, :C, None, None]
    out = G.conv2d(G.nhwc_f16(x), G.pack_conv_weight(w), bias.cuda(), G.nhwc_f16(res), vec.cuda(), B, H, H, C, C, 3, 1, 1,
                   act=1, impl=impl)
    assert G.rel_l2(G.to_nchw_f32(out), ref) < TOL
    ref2 = F.gelu(F.conv2d(x, w, bias, 1, 1))
    out2 = G.conv2d(G.nhwc_f16(x), G.pack_conv_weight(w), bias.cuda(), None, None, B, H, H, C, C, 3, 1, 1, act=2, impl=impl)
    assert G.rel_l2(G.to_nchw_f32(out2), ref2) < TOL


@pytest.mark.parametrize("impl", [1, 0], ids=["simt", "tcgen05"])
@pytest.mark.parametrize("shape", [(2, 4, 512), (3, 16, 64), (1, 64, 64), (5, 2, 256)])
def test_conv_transpose_matches_torch(shape, impl):
    B, H, C = shape
    g = torch.Generator().manual_seed(11)
    x = _bf(torch.randn(B, C, H, H, generator=g))
    w = _bf(torch.randn(C, C, 2, 2, generator=g) / math.sqrt(C))
    bias = torch.randn(C, generator=g)
    ref = F.conv_transpose2d(x, w, bias, stride=2)
    out = G.conv2d(G.nhwc_f16(x), G.pack_convt_weight(w), bias.cuda(), None, None, B, H, H, C, C, 1, 1, 0, convt=True, impl=impl)
    assert G.rel_l2(G.to_nchw_f32(out), ref) < TOL


# Persistent two-accumulator convolution (conv_persist.cuh, impl=3 forces it; needs an un-split plan: >= 75 tiles): several tiles
# per CTA, both N tilings, stride 2, the 8x8 stem geometry, ragged last tile (TN = 2 images per tile, odd batch), ConvTranspose.
PERSIST_CASES = [
    ("p_3x3_c64_32px_b10", 10, 32, 64, 64, 3, 1, 1),
    ("p_3x3_c128_to256_64px_bn128", 4, 64, 128, 256, 3, 1, 1),
    ("p_3x3_s2_c64_to128_64px", 6, 64, 64, 128, 3, 2, 1),
    ("p_8x8_s2_128px", 5, 128, 64, 64, 8, 2, 3),
    ("p_ragged_8px_b161", 161, 8, 64, 64, 3, 1, 1),
    ("p_3x3_c64_64px_b40_many_tiles", 40, 64, 64, 64, 3, 1, 1),          # slab tiling (16 x 8 pixel tiles), BN = 64
    ("p_slab_c128_to256_32px_b24", 24, 32, 128, 256, 3, 1, 1),           # slab tiling, BN = 128
    ("p_slab_c256_16px_b160_deepK", 160, 16, 256, 256, 3, 1, 1),         # slab tiling, 12 ring stages per tile
    ("p_slab_c64_128px_b3", 3, 128, 64, 64, 3, 1, 1),
]


@pytest.mark.parametrize("case", PERSIST_CASES, ids=[c[0] for c in PERSIST_CASES])
def test_persistent_conv_matches_torch_and_one_tile_kernel(case):
    _, B, H, Cin, Cout, R, stride, pad = case
    g = torch.Generator().manual_seed(sum(case[0].encode()))
    x = _bf(torch.randn(B, Cin, H, H, generator=g))
    w = _bf(torch.randn(Cout, Cin, R, R, generator=g) / math.sqrt(Cin * R * R))
    bias = torch.randn(Cout, generator=g)
    Ho = (H + 2 * pad - R) // stride + 1
    res = _bf(torch.randn(B, Cout, Ho, Ho, generator=g))
    vec = torch.randn(B, Cout + 8, generator=g)
    ref = F.relu(F.conv2d(x, w, bias, stride, pad) + res) + vec[:, :Cout, None, None]
    args = (G.nhwc_f16(x), G.pack_conv_weight(w), bias.cuda(), G.nhwc_f16(res), vec.cuda(), B, H, H, Cin, Cout, R, stride, pad)
    out = G.conv2d(*args, act=1, impl=3)             # persistent kernel, slab tiling where the plan chooses it
    assert G.rel_l2(G.to_nchw_f32(out), ref) < TOL
    one = G.conv2d(*args, act=1, impl=4)             # the one-tile-per-CTA kernel
    noslab = G.conv2d(*args, act=1, impl=5)          # persistent kernel without the slab tiling: same K order as the one-tile kernel
    assert torch.equal(noslab, one)
    assert G.rel_l2(out.float(), one.float()) < 1e-3  # the slab walks K in (channel block, column, row) order


@pytest.mark.parametrize("shape", [(40, 16, 256), (10, 64, 64)])
def test_persistent_conv_transpose(shape):
    B, H, C = shape
    g = torch.Generator().manual_seed(13)
    x = _bf(torch.randn(B, C, H, H, generator=g))
    w = _bf(torch.randn(C, C, 2, 2, generator=g) / math.sqrt(C))
    bias = torch.randn(C, generator=g)
    ref = F.conv_transpose2d(x, w, bias, stride=2)
    out = G.conv2d(G.nhwc_f16(x), G.pack_convt_weight(w), bias.cuda(), None, None, B, H, H, C, C, 1, 1, 0, convt=True, impl=3)
    assert G.rel_l2(G.to_nchw_f32(out), ref) < TOL


@pytest.mark.parametrize("C", [64, 128, 256, 512])
def test_layernorm(C):
    g = torch.Generator().manual_seed(C)
    rows = 333
    x = _bf(torch.randn(rows, C, generator=g) * 3 + 1.5)
    gamma, beta = torch.randn(C, generator=g), torch.randn(C, generator=g)
    y = torch.full((rows, C), float("nan"), dtype=torch.float16, device="cuda")
    xd, gd, bd = x.to(torch.float16).cuda(), gamma.cuda(), beta.cuda()   # keep device copies alive across the call
    N.check(N.lib().b2d_op_layernorm(xd.data_ptr(), gd.data_ptr(), bd.data_ptr(), y.data_ptr(), rows, C, G.stream()))
    torch.cuda.synchronize()
    assert G.rel_l2(y.float().cpu(), F.layer_norm(x, (C,), gamma, beta, 1e-5)) < TOL


ATTN_CASES = [(2, 1024, 64, 4), (2, 256, 64, 4), (3, 64, 128, 4), (3, 16, 256, 4), (5, 4, 512, 4), (2, 1, 512, 4),
              (2, 100, 64, 4), (2, 256, 64, 8), (1, 64, 128, 32), (2, 4096, 64, 4), (2, 16, 128, 1),
              # n_heads = 1 (launcher default of the clean application): head_dim = C up to 512
              (3, 16, 256, 1), (3, 64, 256, 1), (5, 4, 512, 1), (2, 16, 512, 1), (2, 1024, 64, 1), (2, 256, 64, 1),
              # head_dim 32 on tcgen05 (attention_tc3d32.cuh): C = 128 with 4 heads, 256 with 8
              (2, 1024, 128, 4), (3, 256, 128, 4), (1, 128, 256, 8), (9, 512, 128, 4),
              # head_dim 16 with the fewest key blocks (L = 128: two blocks, no ring refill) and a three-block case
              (3, 128, 64, 4), (2, 384, 64, 4)]


@pytest.mark.parametrize("case", ATTN_CASES, ids=[f"B{b}_L{l}_C{c}_h{h}" for b, l, c, h in ATTN_CASES])
def test_attention_matches_torch(case):
    B, L, C, heads = case
    d = C // heads
    g = torch.Generator().manual_seed(L + C)
    qkv = _bf(torch.randn(B, L, 3 * C, generator=g))
    q, k, v = (t.reshape(B, L, heads, d).permute(0, 2, 1, 3) for t in qkv.split(C, dim=-1))
    ref = (torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), -1) @ v).permute(0, 2, 1, 3).reshape(B, L, C)
    o = torch.full((B, L, C), float("nan"), dtype=torch.float16, device="cuda")
    qd = qkv.to(torch.float16).cuda()
    N.check(N.lib().b2d_op_attention(qd.data_ptr(), o.data_ptr(), B, L, C, heads, G.stream()))
    torch.cuda.synchronize()
    assert G.rel_l2(o.float().cpu(), ref) < 3e-3     # P is rounded to fp16 before P.V


# Persistent head_dim-16 kernel (attention_tc2.cuh): several work items per CTA (pipeline running across item boundaries, O
# read-out / re-initialisation), and PEAKY score distributions (q, k scaled up: row maxima that keep growing across key blocks
# => lazy-reference moves with O rescaled in TMEM; scores far below the maximum => the packed-half polynomial path must flush
# to zero exactly like the SFU path).
ATTN_TC2_CASES = [(37, 1024, 64, 4, 1.0), (3, 1024, 64, 4, 3.0), (2, 4096, 64, 4, 2.5), (1, 256, 64, 4, 4.0), (150, 256, 64, 4, 1.5),
                  (3, 1024, 128, 4, 2.5), (2, 256, 128, 4, 4.0)]      # head_dim 32, peaky


@pytest.mark.parametrize("case", ATTN_TC2_CASES, ids=[f"B{b}_L{l}_C{c}_h{h}_x{s}" for b, l, c, h, s in ATTN_TC2_CASES])
def test_persistent_attention_items_and_peaky_scores(case):
    B, L, C, heads, qs = case
    d = C // heads
    g = torch.Generator().manual_seed(L + B)
    qkv = torch.randn(B, L, 3 * C, generator=g)
    qkv[..., : 2 * C] *= qs
    qkv = _bf(qkv)
    qd = qkv.to(torch.float16).cuda()
    o = torch.full((B, L, C), float("nan"), dtype=torch.float16, device="cuda")
    N.check(N.lib().b2d_op_attention(qd.data_ptr(), o.data_ptr(), B, L, C, heads, G.stream()))
    torch.cuda.synchronize()
    q, k, v = (t.cuda().double().reshape(B, L, heads, d).permute(0, 2, 1, 3) for t in qkv.split(C, dim=-1))
    errs = []
    for b0 in range(0, B, 8):          # reference on the GPU in fp64, a few samples at a time (B*h*L*L scores)
        sl = slice(b0, min(B, b0 + 8))
        ref = (torch.softmax(q[sl] @ k[sl].transpose(-1, -2) / math.sqrt(d), -1) @ v[sl]).permute(0, 2, 1, 3).reshape(-1, L, C)
        errs.append(G.rel_l2(o[sl].double(), ref))
    assert max(errs) < 3e-3, errs


ATTN_BLOCK_CASES = [(64, 64, 128, 4), (64, 16, 256, 4), (64, 4, 512, 4), (5, 4, 512, 4), (3, 16, 256, 4), (3, 64, 128, 4),
                    (7, 1, 512, 4), (4, 16, 256, 8), (2, 64, 256, 8), (9, 2, 128, 4), (3, 32, 512, 8)]


@pytest.mark.parametrize("case", ATTN_BLOCK_CASES, ids=[f"B{b}_L{l}_C{c}_h{h}" for b, l, c, h in ATTN_BLOCK_CASES])
def test_fused_lowres_attention_block(case):
    """LayerNorm + in-projection + softmax(QK^T)V in one launch vs torch fp32 (nn.LayerNorm + nn.MultiheadAttention internals,
    modules_DANRA_conditional.py:100-107) on the same fp16-rounded inputs."""
    B, L, C, heads = case
    d = C // heads
    g = torch.Generator().manual_seed(7 * L + C + heads)
    x = _bf(torch.randn(B, L, C, generator=g) * 1.7 + 0.4)
    gamma, beta = 1 + 0.3 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    W = torch.randn(3 * C, C, generator=g) / math.sqrt(C)
    bias = 0.1 * torch.randn(3 * C, generator=g)
    xn = F.layer_norm(x, (C,), gamma, beta, 1e-5)
    qkv = xn @ W.t() + bias
    q, k, v = (t.reshape(B, L, heads, d).permute(0, 2, 1, 3) for t in qkv.split(C, dim=-1))
    ref = (torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(d), -1) @ v).permute(0, 2, 1, 3).reshape(B, L, C)
    wf = (W * gamma[None, :]).to(torch.float16)
    c1 = wf.float().sum(1)
    bf = bias + W @ beta
    o = torch.full((B, L, C), float("nan"), dtype=torch.float16, device="cuda")
    xd, wd, cd, bd = x.to(torch.float16).cuda(), wf.cuda(), c1.cuda(), bf.cuda()
    N.check(N.lib().b2d_op_attn_block(xd.data_ptr(), wd.data_ptr(), cd.data_ptr(), bd.data_ptr(), o.data_ptr(), B, L, C, heads,
                                      G.stream()))
    torch.cuda.synchronize()
    assert G.rel_l2(o.float().cpu(), ref) < 4e-3     # fp16 weights, Q/K/V and P rounded to fp16


ATTN_BLOCK_OUT_CASES = [(64, 64, 128, 4, 0), (64, 16, 256, 4, 1), (64, 4, 512, 4, 0), (5, 4, 512, 4, 1), (3, 64, 128, 4, 1),
                        (7, 1, 512, 4, 0), (4, 16, 256, 8, 0), (2, 64, 256, 8, 1), (9, 2, 128, 4, 0), (3, 32, 512, 8, 1)]


@pytest.mark.parametrize("case", ATTN_BLOCK_OUT_CASES, ids=[f"B{b}_L{l}_C{c}_h{h}_act{a}" for b, l, c, h, a in ATTN_BLOCK_OUT_CASES])
def test_fused_lowres_attention_block_with_out_projection(case):
    """The whole ImageSelfAttention block (LN, MHA incl. out_proj, + x, optional ReLU) in one launch vs nn.MultiheadAttention."""
    B, L, C, heads, act = case
    g = torch.Generator().manual_seed(11 * L + C + heads)
    x = _bf(torch.randn(B, L, C, generator=g) * 1.7 + 0.4)
    gamma, beta = 1 + 0.3 * torch.randn(C, generator=g), 0.2 * torch.randn(C, generator=g)
    mha = torch.nn.MultiheadAttention(C, heads, batch_first=True)
    with torch.no_grad():
        mha.in_proj_weight.copy_(torch.randn(3 * C, C, generator=g) / math.sqrt(C))
        mha.in_proj_bias.copy_(0.1 * torch.randn(3 * C, generator=g))
        mha.out_proj.weight.copy_(_bf(torch.randn(C, C, generator=g) / math.sqrt(C)))
        mha.out_proj.bias.copy_(0.1 * torch.randn(C, generator=g))
        xn = F.layer_norm(x, (C,), gamma, beta, 1e-5)
        ref = mha(xn, xn, xn)[0] + x
        if act:
            ref = torch.relu(ref)
    W, bias = mha.in_proj_weight.detach(), mha.in_proj_bias.detach()
    wf = (W * gamma[None, :]).to(torch.float16)
    c1 = wf.float().sum(1)
    bf = bias + W @ beta
    y = torch.full((B, L, C), float("nan"), dtype=torch.float16, device="cuda")
    xd, wd, cd, bd = x.to(torch.float16).cuda(), wf.cuda(), c1.cuda(), bf.cuda()
    wo, bo = mha.out_proj.weight.detach().to(torch.float16).cuda(), mha.out_proj.bias.detach().cuda()
    N.check(N.lib().b2d_op_attn_block_out(xd.data_ptr(), wd.data_ptr(), cd.data_ptr(), bd.data_ptr(), wo.data_ptr(), bo.data_ptr(),
                                          y.data_ptr(), B, L, C, heads, act, G.stream()))
    torch.cuda.synchronize()
    assert G.rel_l2(y.float().cpu(), ref) < 4e-3


@pytest.mark.parametrize("shape", [(2, 16, 512), (3, 1024, 64), (1, 16384, 64), (4, 4, 256)])
def test_instance_norm_with_skip_and_vector(shape):
    B, HW, C = shape
    g = torch.Generator().manual_seed(HW)
    x = _bf(torch.randn(B, HW, C, generator=g) * 2 + 0.7)
    skip = _bf(torch.randn(B, HW, C, generator=g))
    vec = torch.randn(B, C + 8, generator=g)
    mean = x.mean(1, keepdim=True)
    var = x.var(1, unbiased=False, keepdim=True)
    ref = (x - mean) / torch.sqrt(var + 1e-5) + skip + vec[:, None, :C]
    y = torch.full((B, HW, C), float("nan"), dtype=torch.float16, device="cuda")
    ws = torch.empty(B * C * 2, dtype=torch.float32, device="cuda")
    xd, sd, vd = x.to(torch.float16).cuda(), skip.to(torch.float16).cuda(), vec.cuda()
    N.check(N.lib().b2d_op_instnorm(xd.data_ptr(), sd.data_ptr(), vd.data_ptr(), C + 8, y.data_ptr(), ws.data_ptr(), B, HW, C,
                                    G.stream()))
    torch.cuda.synchronize()
    assert G.rel_l2(y.float().cpu(), ref) < TOL


# Decoder.final_layer after its ConvTranspose: InstanceNorm2d -> Conv3x3(64 -> 1) (modules_DANRA_conditional.py:503-509).  The tcgen05
# kernel (taps as the N dimension + shifted sum) against torch in fp32, and against the mma.sync kernel it replaces; shapes cover one
# tile per image (16x8 = 128 pixels), several rows per tile (W = 32, 64), one row per tile (W = 128), bands with and without halo
# tiles above / below, and a non-zero mean (the mean term is dropped for taps outside the image).
@pytest.mark.parametrize("shape", [(3, 8, 16), (2, 32, 32), (5, 64, 64), (2, 128, 128), (1, 256, 64)])
def test_final_layer_instance_norm_conv_matches_torch(shape):
    B, H, W = shape
    g = torch.Generator().manual_seed(H * 1000 + W)
    x = _bf(torch.randn(B, H, W, 64, generator=g) * 1.5 + torch.randn(1, 1, 1, 64, generator=g))
    wt = torch.randn(1, 64, 3, 3, generator=g) * 0.05
    bias = torch.randn(1, generator=g)
    xn = F.instance_norm(x.permute(0, 3, 1, 2), eps=1e-5)
    ref = F.conv2d(xn, wt, bias, padding=1)
    wk = wt[0].permute(1, 2, 0).reshape(-1).contiguous()                # [tap * 64 + c]
    xd, wd, bd = x.to(torch.float16).cuda().contiguous(), wk.cuda(), bias.cuda()
    outs = []
    for legacy in (0, 1):
        out = torch.full((B, 1, H, W), float("nan"), dtype=torch.float32, device="cuda")
        N.check(N.lib().b2d_op_final_layer(xd.data_ptr(), wd.data_ptr(), bd.data_ptr(), out.data_ptr(), B, H, W, legacy, G.stream()))
        torch.cuda.synchronize()
        assert G.rel_l2(out.cpu(), ref) < TOL, legacy
        outs.append(out)
    assert G.rel_l2(outs[0].cpu(), outs[1].cpu()) < 1e-3


def test_posterior_update_bit_exact_with_host_noise():
    """Same op order as diffusion_DANRA_conditional.py:155-157, no FMA contraction => identical bits."""
    from oracle import ddpm_oracle as O
    betas, alphas, ahat = O.schedule_tables(1000, 1e-4, 0.02)
    g = torch.Generator().manual_seed(3)
    bd, ad, hd = betas.cuda(), alphas.cuda(), ahat.cuda()
    for i in (999, 500, 2, 1):
        x = torch.randn(3, 1, 64, 64, generator=g) * 50
        eps = torch.randn(3, 1, 64, 64, generator=g)
        z = torch.randn(3, 1, 64, 64, generator=g) if i > 1 else torch.zeros(3, 1, 64, 64)
        ref = O.posterior_update(x, eps, z, i, betas, alphas, ahat)
        xd, ed, zd = x.clone().cuda(), eps.cuda(), z.cuda()
        N.check(N.lib().b2d_op_posterior_update(xd.data_ptr(), ed.data_ptr(), zd.data_ptr(), bd.data_ptr(), ad.data_ptr(),
                                                hd.data_ptr(), i, 3, 64 * 64, 0, 0, 1.0, G.stream()))
        assert torch.equal(xd.cpu(), ref), i


def test_posterior_update_philox_noise_is_standard_normal_and_shard_invariant():
    from oracle import ddpm_oracle as O
    betas, alphas, ahat = O.schedule_tables(1000, 1e-4, 0.02)
    dev = [t.cuda() for t in (betas, alphas, ahat)]
    B, per = 8, 64 * 64
    zero = torch.zeros(B, 1, 64, 64, device="cuda")

    def draw(batch, offset, i=700):
        x = torch.zeros(batch, 1, 64, 64, device="cuda")
        N.check(N.lib().b2d_op_posterior_update(x.data_ptr(), zero.data_ptr(), None, dev[0].data_ptr(), dev[1].data_ptr(),
                                                dev[2].data_ptr(), i, batch, per, 1234, offset, 1.0, G.stream()))
        return (x / torch.sqrt(betas[i]).item()).cpu()     # x = sqrt(beta) * z

    z = draw(B, 0)
    assert abs(float(z.mean())) < 0.02 and abs(float(z.std()) - 1.0) < 0.02
    assert abs(float((z ** 4).mean()) - 3.0) < 0.15
    # sample k of a shard starting at offset o draws the same z as global sample o+k
    z_shard = draw(4, 4)
    assert torch.equal(z_shard, z[4:])
    assert not torch.equal(draw(B, 0, i=699), z)
