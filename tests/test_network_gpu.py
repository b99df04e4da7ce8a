"""GPU parity of the tcgen05 implicit-GEMM, attention and pointwise kernels against plain PyTorch
fp32 on the CPU (same op, same inputs).  Tolerances: one-pass bf16 operands ("bf16" mode) ~1e-2
relative to the output scale; 3-pass split ("fp32" mode) ~2e-4."""
import math

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu
DEV = "cuda:0"
MODES = [(False, 1.5e-2), (True, 3e-4)]


def rel_err(got, want):
    return ((got - want).abs().max() / want.abs().max().clamp_min(1e-6)).item()


def _act(x, x3, cs=None):
    from patchrefinerv2_b200.nn import Act
    return Act.from_nchw(x.to(DEV), x3, cs)


@pytest.mark.parametrize("x3,tol", MODES)
@pytest.mark.parametrize("M,K,N", [(128, 64, 64), (1025, 384, 1152), (2050, 1024, 4096), (300, 588, 384), (4100, 1536, 384)])
def test_linear_bias(x3, tol, M, K, N):
    from patchrefinerv2_b200.nn import Act, GemmLayer
    g = torch.Generator().manual_seed(M + K)
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g)
    want = F.linear(a, w, b)
    A = _act(a.t().reshape(1, K, 1, M), x3)                          # [N=1, C=K, H=1, W=M]
    lay = GemmLayer([(0, 0, 0, w)], 1, N, x3, DEV, bias=b)
    out = Act.empty(1, 1, M, N, x3, DEV)
    lay([A], out=out)
    got = out.to_nchw()[0, :, 0, :].t().cpu()
    assert rel_err(got, want) < tol


@pytest.mark.parametrize("x3,tol", MODES)
def test_linear_epilogues_gelu_residual_f32(x3, tol):
    from patchrefinerv2_b200 import _lib
    from patchrefinerv2_b200.nn import Act, GemmLayer
    g = torch.Generator().manual_seed(11)
    M, K, N = 1025, 384, 384
    a = torch.randn(M, K, generator=g)
    w = torch.randn(N, K, generator=g) / math.sqrt(K)
    b, gam = torch.randn(N, generator=g), torch.rand(N, generator=g)
    x = torch.randn(M, N, generator=g)
    A = _act(a.t().reshape(1, K, 1, M), x3)
    out = Act.empty(1, 1, M, N, x3, DEV)
    GemmLayer([(0, 0, 0, w)], 1, N, x3, DEV, bias=b, act=_lib.ACT_GELU)([A], out=out)
    assert rel_err(out.to_nchw()[0, :, 0, :].t().cpu(), F.gelu(F.linear(a, w, b))) < tol
    xd = x.clone().to(DEV)
    GemmLayer([(0, 0, 0, w)], 1, N, x3, DEV, epi=_lib.EPI_RESID_F32, bias=b, gamma=gam)([A], out_f32=xd, out_f32_ld=N)
    assert rel_err(xd.cpu(), x + gam * F.linear(a, w, b)) < tol
    of = torch.zeros(M, N, device=DEV)
    GemmLayer([(0, 0, 0, w)], 1, N, x3, DEV, epi=_lib.EPI_F32, bias=b)([A], out_f32=of, out_f32_ld=N)
    assert rel_err(of.cpu(), F.linear(a, w, b)) < tol


@pytest.mark.parametrize("x3,tol", MODES)
@pytest.mark.parametrize("B,H,W,splits,Cout", [(1, 16, 16, [64], 64), (2, 32, 32, [64, 64], 128), (1, 64, 64, [64, 66], 130), (2, 8, 8, [32], 32),
                                               (1, 224, 224, [32, 34], 66), (1, 28, 20, [24], 48)])
def test_conv3x3_virtual_concat(x3, tol, B, H, W, splits, Cout):
    """3x3 pad-1 conv over the channel concatenation of several sources, bias + two residuals + relu copy."""
    from patchrefinerv2_b200.nn import Act, GemmLayer, conv_segments
    g = torch.Generator().manual_seed(H * 7 + Cout)
    xs = [torch.randn(B, c, H, W, generator=g) for c in splits]
    cin = sum(splits)
    w = torch.randn(Cout, cin, 3, 3, generator=g) / math.sqrt(cin * 9)
    b = torch.randn(Cout, generator=g)
    r1, r2 = torch.randn(B, Cout, H, W, generator=g), torch.randn(B, Cout, H, W, generator=g)
    want = F.conv2d(torch.cat(xs, 1), w, b, padding=1) + r1 + r2
    srcs = [_act(x, x3, cs=(c + 7) // 8 * 8 + 8) for x, c in zip(xs, splits)]           # pitch > C on purpose
    lay = GemmLayer(conv_segments(w, splits), len(splits), Cout, x3, DEV, bias=b)
    out, rl = Act.empty(B, H, W, Cout, x3, DEV), Act.empty(B, H, W, Cout, x3, DEV)
    lay(srcs, out=out, relu_out=rl, res=_act(r1, x3), res2=_act(r2, x3))
    assert rel_err(out.to_nchw().cpu(), want) < tol
    assert rel_err(rl.to_nchw().cpu(), F.relu(want)) < tol


@pytest.mark.parametrize("x3,tol", MODES)
def test_conv_ln_gelu_epilogue(x3, tol):
    """SingleConvCNNLN (convs.py:64-75): conv -> LayerNorm over channels -> GELU, strided output tensor."""
    from patchrefinerv2_b200 import _lib
    from patchrefinerv2_b200.nn import Act, GemmLayer, conv_segments
    g = torch.Generator().manual_seed(21)
    B, H, W, C, Cout = 2, 32, 32, 66, 64
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(Cout, C, 3, 3, generator=g) / math.sqrt(C * 9)
    gam, bet = 1 + 0.1 * torch.randn(Cout, generator=g), 0.1 * torch.randn(Cout, generator=g)
    y = F.conv2d(x, w, padding=1)
    u = y.mean(1, keepdim=True)
    s = (y - u).pow(2).mean(1, keepdim=True)
    want = F.gelu(gam[:, None, None] * ((y - u) / torch.sqrt(s + 1e-6)) + bet[:, None, None])
    lay = GemmLayer(conv_segments(w, [C]), 1, Cout, x3, DEV, epi=_lib.EPI_LN_GELU, gamma=gam, beta=bet, eps=1e-6)
    out = Act.empty(B, H, W, Cout, x3, DEV, cs=Cout + 8)
    lay([_act(x, x3, cs=72)], out=out)
    assert rel_err(out.to_nchw().cpu(), want) < tol * 2


@pytest.mark.parametrize("x3,tol", MODES)
@pytest.mark.parametrize("k", [2, 4])
def test_conv_transpose_shuffle(x3, tol, k):
    from patchrefinerv2_b200 import _lib
    from patchrefinerv2_b200.nn import Act, GemmLayer
    g = torch.Generator().manual_seed(31 + k)
    B, H, W, C = 2, 16, 16, 48
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(C, C, k, k, generator=g) / math.sqrt(C)
    b = torch.randn(C, generator=g)
    want = F.conv_transpose2d(x, w, b, stride=k)
    wg = w.permute(2, 3, 1, 0).reshape(k * k * C, C)
    lay = GemmLayer([(0, 0, 0, wg)], 1, k * k * C, x3, DEV, epi=_lib.EPI_SHUFFLE, bias=b, shuffle_k=k)
    out = Act.empty(B, H * k, W * k, C, x3, DEV)
    lay([_act(x, x3)], out=out)
    assert rel_err(out.to_nchw().cpu(), want) < tol


@pytest.mark.parametrize("x3,tol", MODES)
def test_stride2_conv_via_phase_split(x3, tol):
    from patchrefinerv2_b200 import ops
    from patchrefinerv2_b200.nn import Act, GemmLayer
    g = torch.Generator().manual_seed(41)
    B, H, W, C = 2, 16, 16, 96
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(C, C, 3, 3, generator=g) / math.sqrt(C * 9)
    b = torch.randn(C, generator=g)
    want = F.conv2d(x, w, b, stride=2, padding=1)
    segs = []
    for r in range(3):
        for s in range(3):
            segs.append((((r - 1) & 1) * 2 + ((s - 1) & 1), -1 if r == 0 else 0, -1 if s == 0 else 0, w[:, :, r, s]))
    lay = GemmLayer(segs, 4, C, x3, DEV, bias=b)
    ph4 = Act.empty(4 * B, H // 2, W // 2, C, x3, DEV)
    ops.phase_split(_act(x, x3), ph4)
    out = Act.empty(B, H // 2, W // 2, C, x3, DEV)
    lay([ph4.batch_slice(slice(i * B, (i + 1) * B)) for i in range(4)], out=out)
    assert rel_err(out.to_nchw().cpu(), want) < tol


@pytest.mark.parametrize("x3,tol", MODES)
def test_head_epilogue(x3, tol):
    """output_conv2 (dpt.py:109-114) + * max_depth (dpt.py:190)."""
    from patchrefinerv2_b200 import _lib
    from patchrefinerv2_b200.nn import GemmLayer, conv_segments
    g = torch.Generator().manual_seed(51)
    B, H, W, C = 2, 56, 56, 32
    x = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(32, C, 3, 3, generator=g) / math.sqrt(C * 9)
    b = torch.randn(32, generator=g)
    w2, b2 = torch.randn(1, 32, 1, 1, generator=g) / 4, torch.randn(1, generator=g)
    want = torch.sigmoid(F.conv2d(F.relu(F.conv2d(x, w, b, padding=1)), w2, b2)) * 80.0
    lay = GemmLayer(conv_segments(w, [C]), 1, 32, x3, DEV, epi=_lib.EPI_HEAD, bias=b, gamma=w2.reshape(32), beta=b2, head_scale=80.0)
    out = torch.zeros(B, 1, H, W, device=DEV)
    lay([_act(x, x3)], out_f32=out, out_f32_ld=1)
    assert rel_err(out.cpu(), want) < tol


@pytest.mark.parametrize("x3,tol", [(False, 1.5e-2), (True, 5e-4)])
@pytest.mark.parametrize("B,T,heads,scale", [(1, 257, 6, 1.0), (2, 1025, 6, 1.0), (1, 128, 2, 1.0), (1, 1025, 16, 1.0), (3, 1025, 16, 1.0),
                                             (1, 144, 2, 1.0), (1, 129, 2, 1.0), (1, 300, 2, 1.0), (2, 256, 3, 1.0), (2, 640, 3, 1.0), (1, 77, 2, 1.0),
                                             (1, 1025, 4, 3.0), (2, 513, 2, 5.0)])
def test_attention(x3, tol, B, T, heads, scale):
    """softmax(q k^T / 8) v (attention.py:49-62).  Shapes cover: 1025 tokens (7 x 128 keys + a 129-wide last chunk, one leftover
    query row in the tail kernel), exact multiples of 128, a single ragged chunk, narrow last chunks (300 = 2 x 128 + 44), one
    CTA with a single query tile; scale > 1 makes the row maxima grow by more than 2^8 between chunks (lazy O rescale path)."""
    from patchrefinerv2_b200 import ops
    from patchrefinerv2_b200.nn import Act
    g = torch.Generator().manual_seed(T + heads)
    D = heads * 64
    qkv = torch.randn(B, T, 3 * D, generator=g)
    qkv[..., :2 * D] *= scale
    q, k, v = qkv.reshape(B, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    attn = ((q * 64 ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1)
    want = (attn @ v).transpose(1, 2).reshape(B, T, D)
    A = _act(qkv.reshape(B * T, 3 * D).t().reshape(1, 3 * D, 1, B * T), x3)
    if not x3:                                               # the kernel sees bf16-rounded operands: compare against the same inputs
        qb = qkv.to(torch.bfloat16).float()
        q, k, v = qb.reshape(B, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
        want = (((q * 64 ** -0.5) @ k.transpose(-2, -1)).softmax(dim=-1) @ v).transpose(1, 2).reshape(B, T, D)
    out = Act.empty(1, 1, B * T, D, x3, DEV)
    ops.attention(A, B, T, heads, out)
    got = out.to_nchw()[0, :, 0, :].t().reshape(B, T, D).cpu()
    assert torch.isfinite(got).all()
    assert rel_err(got, want) < tol


@pytest.mark.parametrize("x3,tol", [(False, 1e-2), (True, 1e-4)])
def test_layernorm_and_token_kernels(x3, tol):
    from patchrefinerv2_b200 import ops
    from patchrefinerv2_b200.nn import Act
    g = torch.Generator().manual_seed(61)
    B, T, D = 2, 256, 384
    x = torch.randn(B * (T + 1), D, generator=g) * 3 + 1
    w, b = 1 + 0.1 * torch.randn(D, generator=g), 0.1 * torch.randn(D, generator=g)
    want = F.layer_norm(x, (D,), w, b, 1e-6)
    out = Act.empty(1, 1, B * (T + 1), D, x3, DEV)
    ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV), 1e-6, out)
    assert rel_err(out.to_nchw()[0, :, 0, :].t().cpu(), want) < tol
    out2 = Act.empty(B, 16, 16, D, x3, DEV)
    ops.layernorm(x.to(DEV), w.to(DEV), b.to(DEV), 1e-6, out2, drop_period=T + 1)
    want2 = want.reshape(B, T + 1, D)[:, 1:].reshape(B, 16, 16, D).permute(0, 3, 1, 2)
    assert rel_err(out2.to_nchw().cpu(), want2) < tol
    # patch-embed im2col + token assembly
    img = torch.rand(B, 3, 224, 224, generator=g)
    cols = Act.empty(1, 1, B * 256, 588, x3, DEV, cs=592)
    ops.patchify(img.to(DEV), cols)
    mean, std = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1), torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    wantc = F.unfold((img - mean) / std, 14, stride=14).transpose(1, 2).reshape(B * 256, 588)
    assert rel_err(cols.to_nchw()[0, :, 0, :].t().cpu(), wantc) < tol
    emb, cls, pos = torch.randn(B * T, D, generator=g), torch.randn(D, generator=g), torch.randn(T + 1, D, generator=g)
    xo = torch.zeros(B * (T + 1), D, device=DEV)
    ops.assemble_tokens(emb.to(DEV), cls.to(DEV), pos.to(DEV), B, T, D, xo)
    wantx = torch.cat([cls.expand(B, 1, D), emb.reshape(B, T, D)], 1) + pos
    assert torch.equal(xo.cpu().reshape(B, T + 1, D), wantx)


@pytest.mark.parametrize("x3,tol", [(False, 1e-2), (True, 1e-4)])
def test_resize_depth_taps_final_conv(x3, tol):
    from patchrefinerv2_b200 import ops
    from patchrefinerv2_b200.nn import Act
    g = torch.Generator().manual_seed(71)
    x = torch.randn(2, 64, 16, 16, generator=g)
    for size in [(32, 32), (28, 36), (16, 16)]:
        out = Act.empty(2, size[0], size[1], 64, x3, DEV)
        ops.resize_bilinear(_act(x, x3), out)
        assert rel_err(out.to_nchw().cpu(), F.interpolate(x, size, mode="bilinear", align_corners=True)) < tol
    p1, p2 = torch.rand(2, 1, 224, 224, generator=g) * 80, torch.rand(2, 1, 224, 224, generator=g) * 80
    for size in [(224, 224), (64, 64), (8, 8)]:
        # depth maps as an im2col source: channel (r*3+s)*2+d = resized(pred_d) shifted by tap (r,s), zero padded
        t = Act.empty(2, size[0], size[1], 18, x3, DEV, cs=24)
        t.hi.fill_(1.0)
        ops.depth_taps(p1.to(DEV), p2.to(DEV), t)
        got = t.to_nchw().cpu()
        rs = torch.cat([F.interpolate(p1, size, mode="bilinear", align_corners=True), F.interpolate(p2, size, mode="bilinear", align_corners=True)], 1)
        want = F.unfold(rs, 3, padding=1).reshape(2, 2, 9, size[0], size[1]).permute(0, 2, 1, 3, 4).reshape(2, 18, size[0], size[1])
        assert rel_err(got, want) < tol
        assert torch.all(t.hi[..., 18:24] == 0)
    f = torch.randn(2, 32, 56, 56, generator=g)
    w = torch.randn(1, 32, 3, 3, generator=g) / 17
    base = torch.rand(2, 1, 56, 56, generator=g) - 0.3
    want = torch.clamp(base + F.conv2d(f, w, padding=1), min=0)
    from patchrefinerv2_b200 import _lib
    from patchrefinerv2_b200.nn import GemmLayer
    taps = torch.zeros(2, 56, 56, 16, device=DEV)
    GemmLayer([(0, 0, 0, w[0].permute(1, 2, 0).reshape(9, 32))], 1, 9, x3, DEV, epi=_lib.EPI_F32)([_act(f, x3)], out_f32=taps, out_f32_ld=16)
    out = torch.zeros(2, 1, 56, 56, device=DEV)
    ops.tap_stencil(taps, base.to(DEV), out)
    assert rel_err(out.cpu(), want) < tol


@pytest.mark.parametrize("x3,tol", [(False, 1e-2), (True, 1e-4)])
@pytest.mark.parametrize("B,C,H,W", [(2, 32, 56, 56), (3, 128, 70, 45), (1, 8, 5, 3)])
def test_fused_final_conv3x3(x3, tol, B, C, H, W):
    """prv2_final_conv3x3 (one pass) == clamp(base + conv2d(feat, w, padding=1), 0); ragged tiles, halo at every border."""
    from patchrefinerv2_b200 import ops
    g = torch.Generator().manual_seed(5 + C)
    f = torch.randn(B, C, H, W, generator=g)
    w = torch.randn(1, C, 3, 3, generator=g) / (3 * C ** 0.5)
    base = torch.rand(B, 1, H, W, generator=g) - 0.3
    w9c = w[0].permute(1, 2, 0).reshape(9, C).contiguous().to(DEV)
    out = torch.zeros(B, 1, H, W, device=DEV)
    ops.final_conv3x3(_act(f, x3), w9c, base.to(DEV), out)
    assert rel_err(out.cpu(), torch.clamp(base + F.conv2d(f, w, padding=1), min=0)) < tol
    ops.final_conv3x3(_act(f, x3), w9c, None, out)                        # no base: raw offset, no clamp
    assert rel_err(out.cpu(), F.conv2d(f, w, padding=1)) < tol
