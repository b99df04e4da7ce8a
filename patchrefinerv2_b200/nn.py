"""Host-side plumbing for the per-patch network: channels-last activations, a named workspace,
packed weights and the descriptor builder for ``prv2_umma_gemm``.  PyTorch is used only for
device memory and streams; every FLOP runs in the hand-written kernels behind the C ABI."""
from __future__ import annotations

import ctypes as C
import math
import os
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _lib
from ._lib import GemmDesc

BF16 = torch.bfloat16


def stream_ptr() -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def ptr(t: Optional[torch.Tensor]) -> C.c_void_p:
    return C.c_void_p(0 if t is None else t.data_ptr())


def ceil_to(x: int, m: int) -> int:
    return (x + m - 1) // m * m


class Act:
    """Channels-last activation [N,H,W,C] stored with channel pitch ``cs``: ONE bf16 plane ``hi`` (one-pass mode) or a pair of
    FP16 planes ``hi``, ``lo`` with value = hi + lo (x3 / fp32-class mode).  The torch dtype of the planes is just 16-bit storage
    (allocated as bfloat16 in both cases); only the kernels interpret them."""

    __slots__ = ("N", "H", "W", "C", "cs", "hi", "lo")

    def __init__(self, N, H, W, C, cs, hi, lo):
        self.N, self.H, self.W, self.C, self.cs, self.hi, self.lo = N, H, W, C, cs, hi, lo

    @staticmethod
    def empty(N, H, W, C, x3: bool, device, cs: Optional[int] = None, zero: bool = False) -> "Act":
        cs = ceil_to(C, 8) if cs is None else cs
        assert cs % 8 == 0 and cs >= C
        mk = torch.zeros if zero else torch.empty
        hi = mk((N, H, W, cs), dtype=BF16, device=device)
        lo = mk((N, H, W, cs), dtype=BF16, device=device) if x3 else None
        return Act(N, H, W, C, cs, hi, lo)

    def view_channels(self, C: int) -> "Act":
        """Same storage, different logical channel count (e.g. features + depth slots)."""
        assert C <= self.cs
        return Act(self.N, self.H, self.W, C, self.cs, self.hi, self.lo)

    def batch_slice(self, s: slice) -> "Act":
        hi = self.hi[s]
        return Act(hi.shape[0], self.H, self.W, self.C, self.cs, hi, None if self.lo is None else self.lo[s])

    def to_nchw(self) -> torch.Tensor:
        """fp32 [N,C,H,W] copy (API edge / tests) through prv2_act_to_nchw_f32."""
        out = torch.empty((self.N, self.C, self.H, self.W), dtype=torch.float32, device=self.hi.device)
        _lib.call("prv2_act_to_nchw_f32", ptr(self.hi), ptr(self.lo), self.N, self.C, self.H, self.W, self.cs, ptr(out), stream_ptr())
        return out

    @staticmethod
    def from_nchw(x: torch.Tensor, x3: bool, cs: Optional[int] = None) -> "Act":
        x = x.contiguous().float()
        N, Cc, H, W = x.shape
        a = Act.empty(N, H, W, Cc, x3, x.device, cs)
        _lib.call("prv2_nchw_f32_to_act", ptr(x), N, Cc, H, W, ptr(a.hi), ptr(a.lo), a.cs, stream_ptr())
        return a


class Workspace:
    """Named persistent device buffers (stable pointers across forwards of the same shape)."""

    def __init__(self, device, x3: bool):
        self.device, self.x3 = device, x3
        self._t: Dict[tuple, object] = {}

    def act(self, name: str, N, H, W, Cc, cs: Optional[int] = None, zero: bool = False) -> Act:
        cs = ceil_to(Cc, 8) if cs is None else cs
        key = ("act", name, N, H, W, Cc, cs)
        a = self._t.get(key)
        if a is None:
            a = Act.empty(N, H, W, Cc, self.x3, self.device, cs, zero=True)
            self._t[key] = a
        return a

    def f32(self, name: str, *shape) -> torch.Tensor:
        key = ("f32", name, shape)
        t = self._t.get(key)
        if t is None:
            t = torch.zeros(shape, dtype=torch.float32, device=self.device)
            self._t[key] = t
        return t

    def nbytes(self) -> int:
        n = 0
        for v in self._t.values():
            if isinstance(v, Act):
                n += v.hi.numel() * 2 * (2 if v.lo is not None else 1)
            else:
                n += v.numel() * v.element_size()
        return n


def pick_block_n(cout: int) -> Tuple[int, int]:
    """(block_n, Cout_pad): fewest N tiles with block_n <= 256, multiple of 16."""
    n_tiles = (cout + 255) // 256
    block_n = ceil_to((cout + n_tiles - 1) // n_tiles, 16)
    return block_n, block_n * n_tiles


def pick_tile(H: int, W: int) -> Tuple[int, int]:
    """(tile_w, tile_h) with tile_w*tile_h == 128: 16x8 pixel blocks for images, 128x1 for matrices."""
    th = 1
    while th < 8 and th < H:
        th *= 2
    return 128 // th, th


class GemmLayer:
    """One dense layer (linear / conv / deconv) with its weights packed for ``prv2_umma_gemm``.

    ``segs``: list of (source index, dh, dw, W[Cout, C_src] fp32) in K order.  In x3 mode every
    segment expands to (A_hi*W_hi, A_hi*W_lo, A_lo*W_hi); the sources passed at call time expand to
    (hi, lo) pairs accordingly."""

    def __init__(self, segs: Sequence[Tuple[int, int, int, torch.Tensor]], n_src: int, cout: int, x3: bool, device,
                 epi: int = _lib.EPI_STORE, act: int = _lib.ACT_NONE, bias: Optional[torch.Tensor] = None,
                 gamma: Optional[torch.Tensor] = None, beta: Optional[torch.Tensor] = None, eps: float = 1e-6,
                 head_scale: float = 1.0, shuffle_k: int = 0, name: str = "gemm"):
        self.x3, self.n_src, self.device, self.name = x3, n_src, device, name
        if epi == _lib.EPI_LN_GELU and act == _lib.ACT_NONE:
            act = _lib.ACT_GELU
        if act == _lib.ACT_GELU and not x3:
            # one-pass bf16 mode: tanh-form GELU on MUFU.TANH (|diff| <= 1e-3 abs vs the erf form, below the bf16
            # rounding of the stored activation); the fp32-class mode keeps the erf form (abs err 1.5e-7)
            act = _lib.ACT_GELU_TANH
        cout_real = cout
        if epi == _lib.EPI_STORE:
            cout = ceil_to(cout, 8)            # 16-byte channel groups are stored whole; extra rows are zero weights
        self.cout = cout
        self.block_n, self.cout_pad = pick_block_n(cout)
        if bias is not None and cout != cout_real:
            bias = torch.cat([bias.detach().float().reshape(-1), torch.zeros(cout - cout_real)])
        blocks, table, src_c = [], [], {}
        corr_blocks, corr_table = [], []          # x3: the two small cross passes of every segment, issued BEFORE all main passes
        # fp32-class mode: (hi, lo) FP16 weight planes.  FP16's exponent is narrow, so the whole layer is scaled by a power of two
        # that puts its largest weight near 2^14: hi keeps 11 bits, lo = w - hi stays a NORMAL fp16 number for everything within
        # 2^-13 of the maximum (~22 bits in all), and the kernel multiplies the accumulators by the inverse (exact).
        self.w_scale = 1.0
        if x3:
            wmax = max(float(x.detach().abs().max()) for seg in segs for x in (seg[3] if isinstance(seg[3], (list, tuple)) else [seg[3]]))
            if wmax > 0 and math.isfinite(wmax):
                self.w_scale = 2.0 ** math.floor(math.log2(16384.0 / wmax))
        for seg in segs:
            si, dh, dw, w = seg[:4]
            grouped = isinstance(w, (list, tuple))          # vertical tap group: weights of taps (dh-1, dh, dh+1)
            ws = [x.detach().to(device="cpu", dtype=torch.float32) for x in (w if grouped else [w])]   # packing runs on the HOST: one H2D copy per layer, no device kernels at setup
            assert all(x.shape[0] == cout_real for x in ws) and (not grouped or len(ws) == 3)
            c = ws[0].shape[1]
            assert src_c.setdefault(si, c) == c, "a source must present the same channel count in every segment"
            cp = ceil_to(c, 64)
            wp = torch.zeros((len(ws), self.cout_pad, cp), dtype=torch.float32)
            for r, x in enumerate(ws):
                wp[r, :cout_real, :c] = x
            # group: columns interleaved per 64-channel block [block 0: tap -1 | tap 0 | tap +1][block 1: ...] (include/prv2_b200.h)
            wp = wp.reshape(len(ws), self.cout_pad, cp // 64, 64).permute(1, 2, 0, 3).reshape(self.cout_pad, len(ws) * cp)
            taps = 3 if grouped else 1
            if x3:
                wp = wp * self.w_scale
                wh = wp.to(torch.float16)
                wl = (wp - wh.float()).to(torch.float16)
                if os.environ.get("PRV2_X3_ORDER", "main_last") == "main_first":      # diagnostics (scripts/diag_accum.py)
                    blocks += [wh, wl, wh]
                    table += [(2 * si, dh, dw, taps), (2 * si, dh, dw, taps), (2 * si + 1, dh, dw, taps)]
                else:
                    corr_blocks += [wl, wh]
                    corr_table += [(2 * si, dh, dw, taps), (2 * si + 1, dh, dw, taps)]
                    blocks.append(wh)
                    table.append((2 * si, dh, dw, taps))
            else:
                blocks.append(wp.to(BF16))
                table.append((si, dh, dw, taps))
        # The tensor core adds every MMA into its fp32 accumulator with round-toward-zero: a bias of ~2^-24 of the running sum
        # per MMA.  With the cross terms (2^-11 of the result) accumulated first, only the main pass's MMAs see a full-size
        # running sum -- a third of the bias of the interleaved order (scripts/diag_accum.py).
        blocks, table = corr_blocks + blocks, corr_table + table
        assert len(table) <= _lib.MAX_SEG and n_src * (2 if x3 else 1) <= _lib.MAX_SRC
        self.src_c = [src_c[i] for i in range(n_src)]
        self.weight = (blocks[0] if len(blocks) == 1 else torch.cat(blocks, dim=1)).contiguous().to(device)
        self.ktot = self.weight.shape[1]
        f32 = lambda t: None if t is None else t.detach().to(device="cpu", dtype=torch.float32).contiguous().to(device)
        self.bias, self.gamma, self.beta = f32(bias), f32(gamma), f32(beta)
        d = GemmDesc()
        d.Cout, d.block_n, d.Cout_pad, d.Ktot = cout, self.block_n, self.cout_pad, self.ktot
        d.n_src, d.n_seg = n_src * (2 if x3 else 1), len(table)
        for i, (si, dh, dw, taps) in enumerate(table):
            d.seg[i].src, d.seg[i].dh, d.seg[i].dw, d.seg[i].taps_h = si, dh, dw, taps
        d.weight = self.weight.data_ptr()
        d.epi, d.act = epi, act
        d.bias = 0 if self.bias is None else self.bias.data_ptr()
        d.gamma = 0 if self.gamma is None else self.gamma.data_ptr()
        d.beta = 0 if self.beta is None else self.beta.data_ptr()
        d.eps, d.head_scale, d.shuffle_k = eps, head_scale, shuffle_k
        d.acc_scale, d.f16 = 1.0 / self.w_scale, 1 if x3 else 0
        self.desc = d
        self.flops_per_pixel = 2 * cout * sum((3 * w[0].shape[1]) if isinstance(w, (list, tuple)) else w.shape[1] for _, _, _, w in segs)

    def __call__(self, srcs: Sequence[Act], out: Optional[Act] = None, relu_out: Optional[Act] = None, res: Optional[Act] = None,
                 res2: Optional[Act] = None, out_f32: Optional[torch.Tensor] = None, out_f32_ld: int = 0,
                 row_map: Tuple[int, int, int] = (0, 0, 0)) -> None:
        d = self.desc
        a0 = srcs[0]
        d.N, d.H, d.W = a0.N, a0.H, a0.W
        d.tile_w, d.tile_h = pick_tile(a0.H, a0.W)
        assert len(srcs) == self.n_src
        for i, a in enumerate(srcs):
            assert (a.N, a.H, a.W) == (a0.N, a0.H, a0.W) and a.C == self.src_c[i], (i, a.C, self.src_c[i])
            if self.x3:
                assert a.lo is not None, "x3 layer needs (hi, lo) sources"
                d.src[2 * i].ptr, d.src[2 * i].C, d.src[2 * i].cs = a.hi.data_ptr(), a.C, a.cs
                d.src[2 * i + 1].ptr, d.src[2 * i + 1].C, d.src[2 * i + 1].cs = a.lo.data_ptr(), a.C, a.cs
            else:
                d.src[i].ptr, d.src[i].C, d.src[i].cs = a.hi.data_ptr(), a.C, a.cs

        def put(prefix, a):
            if a is None:
                setattr(d, prefix + "_hi", 0); setattr(d, prefix + "_lo", 0); setattr(d, prefix + "_cs", 0)
            else:
                setattr(d, prefix + "_hi", a.hi.data_ptr())
                setattr(d, prefix + "_lo", 0 if a.lo is None else a.lo.data_ptr())
                setattr(d, prefix + "_cs", a.cs)
        put("out", out); put("relu", relu_out); put("res", res); put("res2", res2)
        d.out_f32 = 0 if out_f32 is None else out_f32.data_ptr()
        d.out_f32_ld = out_f32_ld
        d.row_map_period, d.row_map_extra, d.row_map_offset = row_map
        _lib.call("prv2_umma_gemm", C.byref(d), stream_ptr(), work=("flop", float(self.flops_per_pixel) * a0.N * a0.H * a0.W, self.name))


def conv_segments(w: torch.Tensor, splits: Sequence[int], pad: int = 1, group: bool = True) -> List[Tuple[int, int, int, object]]:
    """Segment list of a stride-1 conv with OIHW weight ``w`` whose input channels are the
    concatenation of sources with ``splits`` channels (virtual concat).

    3x3 / pad 1 convs are emitted as vertical tap GROUPS (one per column offset and source): the kernel
    fetches one activation tile with a one-row halo and runs the three row taps from it, which cuts
    the activation traffic of the conv by 2.4x.  Otherwise: one segment per tap, taps outer, sources inner."""
    co, ci, R, S = w.shape
    assert sum(splits) == ci
    segs = []
    if group and R == 3 and S == 3 and pad == 1:
        for s in range(3):
            o = 0
            for si, c in enumerate(splits):
                segs.append((si, 0, s - 1, [w[:, o:o + c, r, s] for r in range(3)]))
                o += c
        return segs
    for r in range(R):
        for s in range(S):
            o = 0
            for si, c in enumerate(splits):
                segs.append((si, r - pad, s - pad, w[:, o:o + c, r, s]))
                o += c
    return segs
