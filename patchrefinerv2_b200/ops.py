"""Torch-tensor wrappers over the geometry / blend / pointwise entry points of the C ABI."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from . import _lib
from ._lib import GridStage
from .nn import Act, ptr, stream_ptr


def _chk(t: torch.Tensor, dtype, name: str):
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise ValueError(f"{name}: expected a contiguous CUDA {dtype} tensor, got {t.dtype} on {t.device}")


def crop_resize(image_hr: torch.Tensor, bboxs: torch.Tensor, ph: int, pw: int, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """image_hr [3,H,W] fp32, bboxs [P,4] int32 (device) -> [P,3,ph,pw] fp32 (baseline_pretrain.py:272-280)."""
    _chk(image_hr, torch.float32, "image_hr"); _chk(bboxs, torch.int32, "bboxs")
    _, H, W = image_hr.shape
    P = bboxs.shape[0]
    if out is None:
        out = torch.empty((P, 3, ph, pw), dtype=torch.float32, device=image_hr.device)
    else:
        _chk(out, torch.float32, "out")
        assert tuple(out.shape) == (P, 3, ph, pw)
    if P == 0:
        return out
    _lib.call("prv2_crop_resize", ptr(image_hr), H, W, ptr(bboxs), P, ptr(out), ph, pw, stream_ptr(), work=("byte", 12.0 * P * ph * pw))
    return out


def roi_gather_f32(feat_hwc: torch.Tensor, rois: torch.Tensor, spatial_scale: float, out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """feat [h,w,C] fp32, rois [P,4] fp32 -> [P,h,w,C] fp32 (patchrefiner.py:199-217, bit-exact form)."""
    _chk(feat_hwc, torch.float32, "feat"); _chk(rois, torch.float32, "rois")
    h, w, Cc = feat_hwc.shape
    P = rois.shape[0]
    if out is None:
        out = torch.empty((P, h, w, Cc), dtype=torch.float32, device=feat_hwc.device)
    else:
        _chk(out, torch.float32, "out")
        assert tuple(out.shape) == (P, h, w, Cc)
    _lib.call("prv2_roi_gather_f32", ptr(feat_hwc), h, w, Cc, ptr(rois), P, C.c_float(spatial_scale), ptr(out), stream_ptr(),
              work=("byte", 4.0 * P * h * w * Cc))
    return out


def roi_gather_act(feat: Act, rois: torch.Tensor, spatial_scale: float, out: Act) -> Act:
    """feat: ONE image (N==1) channels-last act; out [P,h,w,C] act (pre-allocated)."""
    assert feat.N == 1 and (out.H, out.W, out.C) == (feat.H, feat.W, feat.C) and out.N == rois.shape[0]
    _chk(rois, torch.float32, "rois")
    planes = 2 if out.lo is not None else 1
    _lib.call("prv2_roi_gather_act", ptr(feat.hi), ptr(feat.lo), feat.H, feat.W, feat.C, feat.cs, ptr(rois), out.N,
              C.c_float(spatial_scale), ptr(out.hi), ptr(out.lo), out.cs, stream_ptr(),
              work=("byte", 2.0 * planes * out.N * out.H * out.W * out.C))
    return out


def _stages(stages: Sequence[tuple]):
    arr = (GridStage * len(stages))()
    for i, (oh, ow, nh, nw, first) in enumerate(stages):
        arr[i].off_h, arr[i].off_w, arr[i].n_h, arr[i].n_w, arr[i].first = oh, ow, nh, nw, first
    return arr


def blend_canvas(preds: torch.Tensor, mask: torch.Tensor, stages: Sequence[tuple], Hc: int, Wc: int, want_count: bool = True):
    """Sequential-exact process-canvas blend.  preds [n,ph,pw] fp32; stages = [(off_h, off_w, n_h, n_w, first)]."""
    _chk(preds, torch.float32, "preds"); _chk(mask, torch.float32, "mask")
    ph, pw = mask.shape
    avg = torch.empty((Hc, Wc), dtype=torch.float32, device=preds.device)
    cnt = torch.empty((Hc, Wc), dtype=torch.float32, device=preds.device) if want_count else None
    nbytes = 4.0 * (preds.numel() + mask.numel() + (2 if want_count else 1) * Hc * Wc)      # algorithmic bytes
    _lib.call("prv2_blend_canvas", ptr(preds), ptr(mask), ph, pw, _stages(stages), len(stages), Hc, Wc, ptr(avg), ptr(cnt), stream_ptr(),
              work=("byte", nbytes))
    return avg, cnt


def blend_raw_prepare(rmask: torch.Tensor, pw: int) -> torch.Tensor:
    """One-time preparation of the random-patch weight map [rh, rw] for predictions of width ``pw`` (prv2_blend_raw_prepare):
    returns the opaque device buffer the *_raw blend calls accept as ``prep``."""
    _chk(rmask, torch.float32, "rmask")
    rh, rw = rmask.shape
    nbytes = int(_lib.load().prv2_blend_raw_prep_bytes(rh, rw))
    prep = torch.empty(nbytes, dtype=torch.uint8, device=rmask.device)
    _lib.call("prv2_blend_raw_prepare", ptr(rmask), rh, rw, pw, ptr(prep), stream_ptr())
    return prep


def blend_raw(avg_c: torch.Tensor, cnt_c: torch.Tensor, preds: Optional[torch.Tensor], starts: Optional[torch.Tensor],
              rmask: Optional[torch.Tensor], ph: int, pw: int, rh: int, rw: int, H: int, W: int, want_count: bool = True, prep=None):
    """rN stage: resize canvas to raw resolution and fold the random patches in, in draw order."""
    _chk(avg_c, torch.float32, "avg_c"); _chk(cnt_c, torch.float32, "cnt_c")
    Hc, Wc = avg_c.shape
    n = 0 if preds is None else preds.shape[0]
    if n:
        _chk(preds, torch.float32, "preds"); _chk(starts, torch.int32, "starts"); _chk(rmask, torch.float32, "rmask")
    out = torch.empty((H, W), dtype=torch.float32, device=avg_c.device)
    cnt = torch.empty((H, W), dtype=torch.float32, device=avg_c.device) if want_count else None
    nbytes = 4.0 * (2 * Hc * Wc + n * ph * pw + (rh * rw if n else 0) + (2 if want_count else 1) * H * W)
    _lib.call("prv2_blend_raw", ptr(avg_c), ptr(cnt_c), Hc, Wc, ptr(preds), ptr(starts), n, ph, pw, ptr(rmask), rh, rw, H, W,
              ptr(out), ptr(cnt), ptr(prep), stream_ptr(), work=("byte", nbytes))
    return out, cnt


def blend_partial_canvas(preds, own, mask, stages, Hc, Wc, num_c, m1):
    ph, pw = mask.shape
    _lib.call("prv2_blend_partial_canvas", ptr(preds), ptr(own), ptr(mask), ph, pw, _stages(stages), len(stages), Hc, Wc,
              ptr(num_c), ptr(m1), stream_ptr())


def blend_partial_raw(preds, own, starts, rmask, ph, pw, H, W, num_r, prep=None):
    rh, rw = rmask.shape
    _lib.call("prv2_blend_partial_raw", ptr(preds), ptr(own), ptr(starts), preds.shape[0], ph, pw, ptr(rmask), rh, rw, H, W,
              ptr(num_r), ptr(prep), stream_ptr())


def blend_finalize_canvas(num_c, m1, mask, stages, Hc, Wc):
    ph, pw = mask.shape
    avg = torch.empty((Hc, Wc), dtype=torch.float32, device=num_c.device)
    cnt = torch.empty((Hc, Wc), dtype=torch.float32, device=num_c.device)
    _lib.call("prv2_blend_finalize_canvas", ptr(num_c), ptr(m1), ptr(mask), ph, pw, _stages(stages), len(stages), Hc, Wc,
              ptr(avg), ptr(cnt), stream_ptr())
    return avg, cnt


def blend_finalize_raw(avg_c, cnt_c, num_r, starts, rmask, rh, rw, H, W, prep=None):
    Hc, Wc = avg_c.shape
    n = 0 if starts is None else starts.shape[0]
    out = torch.empty((H, W), dtype=torch.float32, device=avg_c.device)
    cnt = torch.empty((H, W), dtype=torch.float32, device=avg_c.device)
    _lib.call("prv2_blend_finalize_raw", ptr(avg_c), ptr(cnt_c), Hc, Wc, ptr(num_r), ptr(starts), n, ptr(rmask), rh, rw, H, W,
              ptr(out), ptr(cnt), ptr(prep), stream_ptr())
    return out, cnt


def layernorm(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float, out: Act, drop_period: int = 0) -> Act:
    """x fp32 [rows, D] -> act rows (optionally dropping the class-token rows)."""
    rows, D = x.shape
    planes = 2 if out.lo is not None else 1
    _lib.call("prv2_layernorm", ptr(x), rows, D, ptr(w), ptr(b), C.c_float(eps), drop_period, ptr(out.hi), ptr(out.lo), out.cs, stream_ptr(),
              work=("byte", rows * D * (4.0 + 2.0 * planes)))
    return out


def layernorm_gelu(x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, eps: float, out: Act) -> Act:
    """fp32 conv rows [pixels, D] -> channels LayerNorm -> exact GELU -> act (convs.py:64-75 for D > 256)."""
    rows, D = x.shape
    planes = 2 if out.lo is not None else 1
    _lib.call("prv2_layernorm_gelu", ptr(x), rows, D, ptr(w), ptr(b), C.c_float(eps), ptr(out.hi), ptr(out.lo), out.cs, stream_ptr(),
              work=("byte", rows * D * (4.0 + 2.0 * planes)))
    return out


def patchify(crops: torch.Tensor, out: Act) -> Act:
    B, _, H, W = crops.shape
    _lib.call("prv2_patchify", ptr(crops), B, H, W, ptr(out.hi), ptr(out.lo), out.cs, stream_ptr(),
              work=("byte", B * 3.0 * H * W * 4 + B * (H // 14) * (W // 14) * out.cs * 2.0 * (2 if out.lo is not None else 1)))
    return out


def assemble_tokens(emb: torch.Tensor, cls: torch.Tensor, pos: torch.Tensor, B: int, T: int, D: int, x: torch.Tensor):
    _lib.call("prv2_assemble_tokens", ptr(emb), ptr(cls), ptr(pos), B, T, D, ptr(x), stream_ptr(), work=("byte", 8.0 * B * (T + 1) * D))


def resize_bilinear(a: Act, out: Act, relu: bool = False) -> Act:
    assert a.C == out.C and a.N == out.N
    planes = 2 if out.lo is not None else 1
    _lib.call("prv2_resize_bilinear_act", ptr(a.hi), ptr(a.lo), a.N, a.H, a.W, a.C, a.cs, ptr(out.hi), ptr(out.lo), out.H, out.W, out.cs,
              1 if relu else 0, stream_ptr(), work=("byte", 2.0 * planes * a.N * a.C * (a.H * a.W + out.H * out.W)))
    return out


def depth_taps(pred1: torch.Tensor, pred2: torch.Tensor, out: Act):
    """pred1/pred2 [N,1,H,W] fp32 -> ``out`` [N,oh,ow,24]: 3x3 neighbourhoods of both maps resized to (oh, ow) as
    18 im2col channels ((r*3+s)*2+d), zero outside the map, channels 18..23 zero."""
    N, _, H, W = pred1.shape
    assert out.N == N and out.cs >= 24
    _lib.call("prv2_depth_taps", ptr(pred1), ptr(pred2), N, H, W, ptr(out.hi), ptr(out.lo), out.H, out.W, out.cs, stream_ptr(),
              work=("byte", N * (8.0 * H * W + 48.0 * out.H * out.W * (2 if out.lo is not None else 1))))


def tap_stencil(taps: torch.Tensor, base: Optional[torch.Tensor], out: torch.Tensor):
    """taps fp32 [N,H,W,ld] (9 tap responses per pixel) -> out[N,1,H,W] = clamp(base + 3x3 gather, 0)."""
    N, H, W, ld = taps.shape
    _lib.call("prv2_tap_stencil", ptr(taps), N, H, W, ld, ptr(base), ptr(out), stream_ptr(), work=("byte", N * H * W * (4.0 * 9 + 8.0)))


def final_conv3x3(feat: Act, w9c: torch.Tensor, base: Optional[torch.Tensor], out: torch.Tensor):
    """feat [N,H,W,C] act, w9c fp32 [9,C] -> out[N,1,H,W] = clamp(base + conv3x3(feat), 0)  (fusion_model.py:113-118), one pass."""
    planes = 2 if feat.lo is not None else 1
    _lib.call("prv2_final_conv3x3", ptr(feat.hi), ptr(feat.lo), feat.N, feat.H, feat.W, feat.C, feat.cs, ptr(w9c), ptr(base), ptr(out), stream_ptr(),
              work=("byte", feat.N * feat.H * feat.W * (2.0 * planes * feat.C + 8.0)))


def phase_split(a: Act, out: Act):
    """out is [4*N, ceil(H/2), ceil(W/2), C] (phase-major; zero where an odd H / W leaves a phase one row / column short)."""
    _lib.call("prv2_phase_split", ptr(a.hi), ptr(a.lo), a.N, a.H, a.W, a.C, a.cs, ptr(out.hi), ptr(out.lo), out.cs, stream_ptr(),
              work=("byte", 4.0 * a.N * a.H * a.W * a.C * (2 if a.lo is not None else 1)))


def attention(qkv: Act, B: int, T: int, heads: int, out: Act):
    _lib.call("prv2_attention", ptr(qkv.hi), ptr(qkv.lo), B, T, heads, ptr(out.hi), ptr(out.lo), stream_ptr(),
              work=("flop", 4.0 * B * heads * T * T * 64))


def zoe_attractor(a_raw: torch.Tensor, n_attractors: int, b_prev: torch.Tensor, prev_is_raw: bool, alpha: float, mean: bool) -> torch.Tensor:
    """a_raw fp32 [B,h,w,ld] (attractor MLP output, pre-softplus), b_prev fp32 [B,hp,wp,K] -> new bin centres fp32 [B,h,w,K]
    (external/zoedepth/models/layers/attractor.py:139-208)."""
    _chk(a_raw, torch.float32, "a_raw"); _chk(b_prev, torch.float32, "b_prev")
    B, h, w, ld = a_raw.shape
    _, hp, wp, K = b_prev.shape
    out = torch.empty((B, h, w, K), dtype=torch.float32, device=a_raw.device)
    _lib.call("prv2_zoe_attractor", ptr(a_raw), ld, n_attractors, ptr(b_prev), hp, wp, 1 if prev_is_raw else 0, ptr(out), B, h, w, K,
              C.c_float(alpha), 1 if mean else 0, stream_ptr(), work=("byte", 4.0 * B * h * w * (2 * K + ld)))
    return out


def zoe_logbinomial_depth(pt_raw: torch.Tensor, centers: torch.Tensor, min_temp: float, max_temp: float) -> torch.Tensor:
    """pt_raw fp32 [B,H,W,ld>=4] (conditional MLP output, pre-softplus), centers fp32 [B,hb,wb,K] -> metric depth fp32 [B,1,H,W]
    (layers/dist_layers.py:29-122, zoedepth_v1.py:212-219)."""
    _chk(pt_raw, torch.float32, "pt_raw"); _chk(centers, torch.float32, "centers")
    B, H, W, ld = pt_raw.shape
    _, hb, wb, K = centers.shape
    out = torch.empty((B, 1, H, W), dtype=torch.float32, device=pt_raw.device)
    _lib.call("prv2_zoe_logbinomial_depth", ptr(pt_raw), ld, ptr(centers), hb, wb, ptr(out), B, H, W, K, C.c_float(min_temp), C.c_float(max_temp),
              stream_ptr(), work=("byte", 4.0 * B * H * W * (ld + 1 + K)))
    return out
