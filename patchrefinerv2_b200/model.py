"""Drop-in ``PatchRefiner`` estimator model: the reference's infer forward
(estimator/models/patchrefiner.py:286-401 with BaselinePretrain.regular_tile / random_tile,
estimator/models/baseline_pretrain.py:149-375) re-built on the sm_100a kernels.

Kept from the reference API: construction from ``config`` (keys of configs/patchrefiner_dav2/
pr_u4k.py:10-53), ``forward(mode='infer', image_lr, image_hr, tile_cfg, cai_mode, process_num)``
returning ``(depth, log_dict)`` with ``depth`` fp32 ``[1,1,ph*Sh,pw*Sw]`` (m1/m2) or ``[1,1,H,W]``
(rN), ``tile_cfg`` / ``resizer`` attributes, ``load_dict`` / ``get_save_dict`` and state-dict keys
``coarse_branch.* / refiner_fine_branch.* / refiner_fusion_model.*``.

Changed by design: patches are cropped, gathered, refined and blended on the device in batches
(no per-patch host round trips), and can be sharded over ranks with one NCCL sum-reduce.
"""
from __future__ import annotations

import os
import random as _random
from collections import OrderedDict
from typing import Dict, List, Optional

import numpy as np
import torch
import torch.nn as nn

from . import _lib, masks, ops, tiling
from .dav2 import VIT_CFG, DepthAnythingV2B200, PATCH
from .fusion import FusionUnetB200
from .nn import Act, Workspace, ptr, stream_ptr
from .registry import MODELS


def _get(cfg, key, default=None):
    if isinstance(cfg, dict):
        return cfg.get(key, default)
    return getattr(cfg, key, default)


def dav2_weight_spec(encoder: str, features: int, out_channels) -> "OrderedDict[str, tuple]":
    """Names and shapes of a DepthAnythingV2 state dict (external/depth_anything_v2/dpt.py:153-181)."""
    c = VIT_CFG[encoder]
    D, depth = c["dim"], c["depth"]
    s: "OrderedDict[str, tuple]" = OrderedDict()
    p = "pretrained."
    s[p + "cls_token"] = (1, 1, D)
    s[p + "pos_embed"] = (1, 37 * 37 + 1, D)
    s[p + "mask_token"] = (1, D)
    s[p + "patch_embed.proj.weight"] = (D, 3, PATCH, PATCH)
    s[p + "patch_embed.proj.bias"] = (D,)
    for i in range(depth):
        b = f"{p}blocks.{i}."
        for n, shp in (("norm1.weight", (D,)), ("norm1.bias", (D,)), ("attn.qkv.weight", (3 * D, D)), ("attn.qkv.bias", (3 * D,)),
                       ("attn.proj.weight", (D, D)), ("attn.proj.bias", (D,)), ("ls1.gamma", (D,)), ("norm2.weight", (D,)),
                       ("norm2.bias", (D,)), ("mlp.fc1.weight", (4 * D, D)), ("mlp.fc1.bias", (4 * D,)),
                       ("mlp.fc2.weight", (D, 4 * D)), ("mlp.fc2.bias", (D,)), ("ls2.gamma", (D,))):
            s[b + n] = shp
    s[p + "norm.weight"] = (D,)
    s[p + "norm.bias"] = (D,)
    h = "depth_head."
    oc, Fe = list(out_channels), features
    for i, o in enumerate(oc):
        s[f"{h}projects.{i}.weight"] = (o, D, 1, 1)
        s[f"{h}projects.{i}.bias"] = (o,)
    s[h + "resize_layers.0.weight"] = (oc[0], oc[0], 4, 4)
    s[h + "resize_layers.0.bias"] = (oc[0],)
    s[h + "resize_layers.1.weight"] = (oc[1], oc[1], 2, 2)
    s[h + "resize_layers.1.bias"] = (oc[1],)
    s[h + "resize_layers.3.weight"] = (oc[3], oc[3], 3, 3)
    s[h + "resize_layers.3.bias"] = (oc[3],)
    for i, o in enumerate(oc):
        s[f"{h}scratch.layer{i + 1}_rn.weight"] = (Fe, o, 3, 3)
    for r in (1, 2, 3, 4):
        q = f"{h}scratch.refinenet{r}."
        s[q + "out_conv.weight"] = (Fe, Fe, 1, 1)
        s[q + "out_conv.bias"] = (Fe,)
        for u in (1, 2):
            for cv in (1, 2):
                s[f"{q}resConfUnit{u}.conv{cv}.weight"] = (Fe, Fe, 3, 3)
                s[f"{q}resConfUnit{u}.conv{cv}.bias"] = (Fe,)
    s[h + "scratch.output_conv1.weight"] = (Fe // 2, Fe, 3, 3)
    s[h + "scratch.output_conv1.bias"] = (Fe // 2,)
    s[h + "scratch.output_conv2.0.weight"] = (32, Fe // 2, 3, 3)
    s[h + "scratch.output_conv2.0.bias"] = (32,)
    s[h + "scratch.output_conv2.2.weight"] = (1, 32, 1, 1)
    s[h + "scratch.output_conv2.2.bias"] = (1,)
    return s


def fusion_weight_spec(input_chl, temp_chl, dec_chl) -> "OrderedDict[str, tuple]":
    """FusionUnet state dict (estimator/models/blocks/fusion_model.py:52-82)."""
    s: "OrderedDict[str, tuple]" = OrderedDict()
    for idx, (ic, tc) in enumerate(zip(input_chl, temp_chl)):
        s[f"encoder_layers_1.{idx}.single_conv.0.weight"] = (tc, ic, 3, 3)
        s[f"encoder_layers_1.{idx}.single_conv.1.weight"] = (tc,)
        s[f"encoder_layers_1.{idx}.single_conv.1.bias"] = (tc,)
    for idx, (ic, tc) in enumerate(zip(input_chl, temp_chl)):
        s[f"encoder_layers_2.{idx}.single_conv.0.weight"] = (tc, tc + 2, 3, 3)
        s[f"encoder_layers_2.{idx}.single_conv.1.weight"] = (tc,)
        s[f"encoder_layers_2.{idx}.single_conv.1.bias"] = (tc,)
    rev = list(temp_chl)[::-1]
    chl = rev[0]
    for i, (tc, dc) in enumerate(zip(rev[1:], dec_chl)):
        cin = tc + chl + 2
        s[f"decoder_layers.{i}.conv.double_conv.0.weight"] = (cin, cin, 3, 3)
        s[f"decoder_layers.{i}.conv.double_conv.2.weight"] = (dc, cin, 3, 3)
        chl = dc
    last = dec_chl[-1] if len(dec_chl) else chl
    s["final_conv.weight"] = (1, last, 3, 3)
    return s


class _Resizer:
    """Stand-in for ``model.resizer`` (external/depth_anything/transform.py:127-129): bilinear,
    align_corners=True, to the process shape rounded to multiples of 14.  Runs the crop kernel."""

    def __init__(self, patch_process_shape):
        self.size = tiling.resizer_size(patch_process_shape)

    def __call__(self, x: torch.Tensor) -> torch.Tensor:
        lead = x.shape[:-3]
        img = x.reshape(-1, *x.shape[-3:])
        outs = []
        for im in img:
            assert im.shape[0] == 3, "resizer expects RGB"
            H, W = im.shape[-2:]
            bb = torch.tensor([[0, 0, W, H]], dtype=torch.int32, device=im.device)
            outs.append(ops.crop_resize(im.contiguous().float(), bb, self.size[0], self.size[1]))
        return torch.cat(outs, 0).reshape(*lead, 3, *self.size)


@MODELS.register_module()
class PatchRefiner(nn.Module):
    """B200-native PatchRefiner (DA2 coarse + DA2 refiner + FusionUnet family)."""

    def __init__(self, config, precision: str = "fp32", patch_batch: int = 8, output_device: str = "cpu"):
        super().__init__()
        if hasattr(config, "to_dict"):
            config = config.to_dict()
        self.config = config
        self.min_depth = _get(config, "min_depth")
        self.max_depth = float(_get(config, "max_depth"))
        self.patch_process_shape = tuple(_get(config, "patch_process_shape"))
        self.tile_cfg = self.prepare_tile_cfg(_get(config, "image_raw_shape"), _get(config, "patch_split_num"))
        cb, rf = _get(config, "coarse_branch"), _get(config, "refiner")
        fb, fu = _get(rf, "fine_branch"), _get(rf, "fusion_model")
        for name, br in (("coarse_branch", cb), ("refiner.fine_branch", fb)):
            if _get(br, "type") != "DA2":
                raise NotImplementedError(f"{name}.type={_get(br, 'type')!r}: only the DA2 (DepthAnythingV2) branch is implemented")
        if _get(fu, "type") != "FusionUnet":
            raise NotImplementedError(f"fusion_model.type={_get(fu, 'type')!r}: only FusionUnet is implemented")
        self._cb_cfg, self._fb_cfg, self._fu_cfg = dict(_get(cb, "model_cfg")), dict(_get(fb, "model_cfg")), fu
        self.fusion_feat_level = int(_get(config, "fusion_feat_level"))
        self.strategy_refiner_target = _get(config, "strategy_refiner_target")
        self.pre_norm_bbox = _get(config, "pre_norm_bbox", True)
        self.resizer = _Resizer(self.patch_process_shape)
        if tuple(self.resizer.size) != tuple(self.patch_process_shape):
            raise NotImplementedError("patch_process_shape must be a multiple of 14 for the DA2 branches")
        assert precision in ("bf16", "fp32"), "precision is 'bf16' (one tcgen05 pass) or 'fp32' (3-pass bf16 split, fp32-class)"
        self.precision, self.patch_batch, self.output_device = precision, int(patch_batch), output_device

        # weights: flat dict with the reference's key names (zeros until load_dict / pretrained files)
        self._weights: "OrderedDict[str, torch.Tensor]" = OrderedDict()
        spec_of = lambda c: dav2_weight_spec(c["encoder"], c["features"], c["out_channels"])
        for pre, spec in (("coarse_branch.", spec_of(self._cb_cfg)), ("refiner_fine_branch.", spec_of(self._fb_cfg)),
                          ("refiner_fusion_model.", fusion_weight_spec(_get(fu, "input_chl"), _get(fu, "temp_chl"), _get(fu, "dec_chl")))):
            for k, shp in spec.items():
                self._weights[pre + k] = torch.zeros(shp)
        # checkpoints named by the config, in the reference's order and strictness (patchrefiner.py:94-98, 118-121, 129-147): the branch
        # files are loaded strict=True over their branch (a wrong-keyed file raises, as DepthAnythingV2.load_state_dict would),
        # `pretrained` is strict=False and skips coarse_branch.* unless `load_whole`
        for pre, br in (("coarse_branch.", cb), ("refiner_fine_branch.", fb)):
            path = _get(br, "pretrained")
            if path:
                self._load_branch(pre, torch.load(path, map_location="cpu"))
        for key in ("pretrain_coarse_model", "pretrain_fine_model"):
            path = _get(config, key)
            if path:
                self._load_branch("coarse_branch." if "coarse" in key else "refiner_fine_branch.", torch.load(path, map_location="cpu")["model_state_dict"])
        path = _get(config, "pretrained")
        if path:
            sd = torch.load(path, map_location="cpu")["model_state_dict"]
            if not _get(config, "load_whole", False):
                sd = {k: v for k, v in sd.items() if "coarse_branch" not in k}
            self._load(sd, strict=False)
        self._engine = None
        self._device = torch.device("cpu")
        self.last_stats: dict = {}

    # -- reference-compatible surface ---------------------------------------------------------
    def prepare_tile_cfg(self, image_raw_shape, patch_split_num):
        return tiling.prepare_tile_cfg(self.patch_process_shape, image_raw_shape, patch_split_num)

    def _load(self, sd, strict):
        missing = [k for k in self._weights if k not in sd]
        unexpected = [k for k in sd if k not in self._weights]
        if not hasattr(self, "_loaded_keys"):
            self._loaded_keys = set()
        for k, v in sd.items():
            if k in self._weights:
                if tuple(v.shape) != tuple(self._weights[k].shape):
                    raise RuntimeError(f"size mismatch for {k}: {tuple(v.shape)} vs {tuple(self._weights[k].shape)}")
                self._weights[k] = v.detach().float().cpu().clone()
                self._loaded_keys.add(k)
        if strict and (missing or unexpected):
            raise RuntimeError(f"missing keys {missing[:4]}..., unexpected keys {unexpected[:4]}...")
        self._engine = None
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def _load_branch(self, pre: str, sd):
        """``self.<branch>.load_state_dict(sd)`` of the reference (strict=True): every tensor of the branch, nothing else."""
        want = {k[len(pre):] for k in self._weights if k.startswith(pre)}
        missing, unexpected = sorted(want - set(sd)), sorted(set(sd) - want)
        if missing or unexpected:
            raise RuntimeError(f"Error(s) in loading state_dict for {pre[:-1]}: missing keys {missing[:4]}{'...' if len(missing) > 4 else ''}, "
                               f"unexpected keys {unexpected[:4]}{'...' if len(unexpected) > 4 else ''}")
        return self._load({pre + k: v for k, v in sd.items()}, strict=False)

    def load_state_dict(self, state_dict, strict: bool = True, assign: bool = False):
        return self._load(state_dict, strict)

    def state_dict(self, *args, destination=None, prefix: str = "", keep_vars: bool = False):
        out = OrderedDict() if destination is None else destination
        for k, v in self._weights.items():
            out[prefix + k] = v.clone()
        return out

    _native_comm = None

    def use_native_reduce(self):
        """Route the sharded path's sum-reduce through ``prv2_reduce_canvas`` on a communicator of this model's own (comm.py)."""
        from .comm import NcclComm
        if self._native_comm is None:
            self._native_comm = NcclComm(self._device if self._device.type == "cuda" else torch.device("cuda", torch.cuda.current_device()))
        return self

    def load_dict(self, dict):                      # patchrefiner.py:155-156
        return self.load_state_dict(dict, strict=False)

    def get_save_dict(self):                        # patchrefiner.py:158-166
        return {k: v for k, v in self.state_dict().items() if "coarse_branch." not in k}

    def _apply(self, fn, recurse=True):
        probe = fn(torch.zeros(1))
        if probe.device != self._device:
            self._device = probe.device
            self._engine = None
        return super()._apply(fn, recurse)

    # -- device engine -------------------------------------------------------------------------
    def _warn_unloaded(self):
        """load_dict / the checkpoint keys of the config are strict=False like the reference's (patchrefiner.py:155-156): a wrong-keyed or
        absent checkpoint must not run silently on the zero weights this class starts from."""
        never = [k for k in self._weights if k not in getattr(self, "_loaded_keys", ()) and not k.endswith(("mask_token", "num_batches_tracked"))]
        if never:
            import warnings
            warnings.warn(f"{type(self).__name__}: {len(never)} of {len(self._weights)} weight tensors were never loaded and are still at their "
                          f"initial values (first: {never[0]}); check the checkpoint's key names", RuntimeWarning, stacklevel=3)

    def _build_engine(self, device):
        if device.type != "cuda":
            raise RuntimeError("patchrefinerv2_b200 runs on CUDA (sm_100a) only; there is no CPU path. Call .cuda() first.")
        self._warn_unloaded()
        _lib.load()
        x3 = self.precision == "fp32"
        sd = self._weights
        fu = self._fu_cfg
        eng = dict(
            coarse=DepthAnythingV2B200(sd, "coarse_branch.", self._cb_cfg["encoder"], self._cb_cfg["features"], self._cb_cfg["out_channels"], self.max_depth, x3, device),
            fine=DepthAnythingV2B200(sd, "refiner_fine_branch.", self._fb_cfg["encoder"], self._fb_cfg["features"], self._fb_cfg["out_channels"], self.max_depth, x3, device),
            fusion=FusionUnetB200(sd, "refiner_fusion_model.", _get(fu, "input_chl"), _get(fu, "temp_chl"), _get(fu, "dec_chl"), x3, device),
            ws=Workspace(device, x3), device=device, masks={})
        return eng

    def _mask_dev(self, eng, kind, size):
        key = (kind, tuple(size))
        if key not in eng["masks"]:
            m = masks.generatemask(size, border=0.15) if kind == "p" else masks.random_patch_mask(size, border=0.15)
            eng["masks"][key] = torch.from_numpy(np.array(m, copy=True)).to(eng["device"])
        return eng["masks"][key]

    def coarse_forward(self, image_lr):             # patchrefiner.py:168-185
        eng = self._engine
        depth, feats = eng["coarse"].forward(image_lr.float().contiguous())
        return feats, depth

    def _coarse_forward_sharded(self, eng, image_lr, rank: int, world: int):
        """Coarse pass of frames [rank * F/world, (rank + 1) * F/world) on this rank, then all_gather of the six feature maps
        (16-bit planes) and the coarse depth: every rank ends with the whole batch's coarse results in frame order."""
        dist = torch.distributed
        F_ = image_lr.shape[0]
        fl = F_ // world
        feats, depth = self.coarse_forward(image_lr[rank * fl:(rank + 1) * fl])
        ws = eng["ws"]
        out = []
        for li, f in enumerate(feats):
            g = ws.act(f"cg{li}_{F_}", F_, f.H, f.W, f.C, cs=f.cs)
            dist.all_gather_into_tensor(g.hi, f.hi.contiguous())
            if f.lo is not None:
                dist.all_gather_into_tensor(g.lo, f.lo.contiguous())
            out.append(g)
        gd = ws.f32(f"cgd_{F_}", F_, 1, depth.shape[-2], depth.shape[-1])
        dist.all_gather_into_tensor(gd, depth.contiguous())
        return out, gd

    def _gather_batch(self, eng, image_hr, bboxs_np, rois_np, coarse_feats, coarse_depth, idx: np.ndarray, P: int):
        """Crops and coarse ROIs of the work items ``idx`` (global indices frame * P + patch, ascending) into one batch: every
        frame present in the batch contributes a contiguous run of rows (baseline_pretrain.py:272-296, patchrefiner.py:199-217)."""
        dev = eng["device"]
        ph, pw = self.patch_process_shape
        ws = eng["ws"]
        pb = len(idx)
        crops = ws.f32(f"crops{pb}", pb, 3, ph, pw)
        d_roi = ws.f32(f"droi{pb}", pb, 1, ph, pw)
        c_roi = [ws.act(f"roi{li}", pb, f.H, f.W, f.C) for li, f in enumerate(coarse_feats)]
        tiling.check_roi_regime(rois_np[idx], [(f.H, f.W, f.H / ph) for f in coarse_feats] + [(ph, pw, 1.0)])
        bb = torch.from_numpy(np.ascontiguousarray(bboxs_np[idx])).to(dev, non_blocking=True)
        rois = torch.from_numpy(np.ascontiguousarray(rois_np[idx])).to(dev, non_blocking=True)
        frames = idx // P
        a = 0
        while a < pb:
            f = int(frames[a])
            b = a + int(np.searchsorted(frames[a:], f, side="right"))
            ops.crop_resize(image_hr[f], bb[a:b], ph, pw, out=crops[a:b])
            for li, feat in enumerate(coarse_feats):
                ops.roi_gather_act(feat.batch_slice(slice(f, f + 1)), rois[a:b], feat.H / ph, c_roi[li].batch_slice(slice(a, b)))
            ops.roi_gather_f32(coarse_depth[f].reshape(ph, pw, 1), rois[a:b], 1.0, out=d_roi[a:b].reshape(b - a, ph, pw, 1))
            a = b
        return crops, c_roi, d_roi

    def refine_patches(self, eng, image_hr, bboxs_np, rois_np, coarse_feats, coarse_depth, sel: np.ndarray, preds: torch.Tensor, P: int, trace=None):
        """crop -> ROI gather -> fine branch -> fusion for the work items ``sel`` (global indices frame * P + patch into the
        flattened schedule of a batch of frames), ``patch_batch`` at a time; results land in ``preds[sel]``.  ``image_hr`` is
        [F,3,H,W], ``coarse_feats`` acts with N = F, ``coarse_depth`` [F,1,ph,pw]."""
        dev = eng["device"]
        ph, pw = self.patch_process_shape
        level = self.fusion_feat_level
        for s in range(0, len(sel), self.patch_batch):
            idx = sel[s:s + self.patch_batch]
            pb = len(idx)
            crops, c_roi, d_roi = self._gather_batch(eng, image_hr, bboxs_np, rois_np, coarse_feats, coarse_depth, idx, P)
            r_depth, r_feats = eng["fine"].forward(crops, trace)                             # patchrefiner.py:219-232
            if self.strategy_refiner_target == "offset_fine":
                base = r_depth
            elif self.strategy_refiner_target == "offset_coarse":
                base = d_roi
            else:
                base = None
            c_list = c_roi[-level:][::-1]                                                   # patchrefiner.py:245-251
            f_list = r_feats[-level:][::-1]
            pred = eng["fusion"].forward(c_list, f_list, d_roi, r_depth, base, trace)        # fusion_model.py:84-122
            if self.strategy_refiner_target == "direct":                                    # patchrefiner.py:280-281 (no shipped config; base is None above)
                pred = torch.sigmoid(pred) * self.max_depth
            preds[torch.from_numpy(idx).to(dev, non_blocking=True)] = pred.reshape(pb, ph, pw)
            if trace is not None:
                trace.setdefault("crops", crops.clone()); trace.setdefault("roi_depth", d_roi.clone())
                trace.setdefault("roi_feats", [a.to_nchw() for a in c_roi]); trace.setdefault("fine_depth", r_depth.clone())
                trace.setdefault("fine_feats", [a.to_nchw() for a in r_feats])
                trace = None if "fine_depth" in trace else trace

    @torch.no_grad()
    def predict_patches(self, image_lr, image_hr, bboxs, tile_cfg=None, trace: Optional[dict] = None):
        """The front half the reference shares between regular_tile and random_tile (baseline_pretrain.py:158-206 / :249-345):
        coarse pass, then crop -> roi_align -> fine branch -> fusion for the raw-image boxes ``bboxs`` [P,4] int (x0,y0,x1,y1).
        Returns (preds [P,ph,pw] fp32 on the device, coarse_prediction).  Used by the parity checks (bench.py, tests)."""
        tile_cfg = self.tile_cfg if tile_cfg is None else self.prepare_tile_cfg(tile_cfg["image_raw_shape"], tile_cfg["patch_split_num"])
        dev = image_hr.device
        if self._engine is None or self._engine["device"] != dev:
            self._engine = self._build_engine(dev)
        eng = self._engine
        ph, pw = self.patch_process_shape
        H, W = tile_cfg["image_raw_shape"]
        bboxs_np = np.ascontiguousarray(np.asarray(bboxs.cpu() if torch.is_tensor(bboxs) else bboxs, dtype=np.int32).reshape(-1, 4))
        rois_np = tiling.bboxs_to_feat(bboxs_np, (H, W), (ph, pw))[:, 1:]
        coarse_feats, coarse_depth = self.coarse_forward(image_lr)
        preds = torch.empty((bboxs_np.shape[0], ph, pw), dtype=torch.float32, device=dev)
        self.refine_patches(eng, image_hr[:1].float().contiguous(), bboxs_np, rois_np, coarse_feats, coarse_depth,
                            np.arange(bboxs_np.shape[0]), preds, bboxs_np.shape[0], trace)
        return preds, coarse_depth

    @torch.no_grad()
    def forward(self, mode=None, image_lr=None, image_hr=None, crops_image_hr=None, depth_gt=None, crop_depths=None, bboxs=None,
                tile_cfg=None, cai_mode="m1", process_num=4, shard: bool = False, trace: Optional[dict] = None, **kwargs):
        if mode == "train":
            raise NotImplementedError("training is out of scope of patchrefinerv2_b200 (inference hot path only)")
        if tile_cfg is None:
            tile_cfg = self.tile_cfg
        else:
            tile_cfg = self.prepare_tile_cfg(tile_cfg["image_raw_shape"], tile_cfg["patch_split_num"])
        # The reference asserts batch 1 (patchrefiner.py:348) and its Tester feeds frames one by one (tester.py:62-69).  Here a
        # batch of F frames is ONE work list of F x P patches (BASELINE config 5): the schedules are drawn frame by
        # frame in order (so the global `random` stream is consumed exactly as F successive reference calls would), the coarse
        # pass runs once on the whole batch, the flattened patches are refined in mixed-frame batches (and split over the ranks:
        # round-robin for one frame, contiguous blocks for a batch), every frame keeps its own canvases and one sum-reduce
        # combines the whole batch.
        F_ = int(image_hr.shape[0])
        dev = image_hr.device if image_hr.is_cuda else self._device
        if self._engine is None or self._engine["device"] != dev:
            self._engine = self._build_engine(dev)
        eng = self._engine
        hr_ready = hr_host = None
        hr_shape = tuple(image_hr.shape)
        if not image_hr.is_cuda:
            # frame ingest: host frames are uploaded by the model -- image_lr first on the compute stream,
            # the 100 MB image_hr on a copy stream UNDER the coarse pass (only the crop kernel needs it); pinned memory makes
            # both copies asynchronous
            if "copy_stream" not in eng:
                eng["copy_stream"] = torch.cuda.Stream(device=dev)
            image_lr = image_lr.to(dev, non_blocking=True)
            hr_host, image_hr = image_hr, None                     # uploaded below, once this rank's share of the work list is known
        elif not image_lr.is_cuda:
            image_lr = image_lr.to(dev, non_blocking=True)
        ph, pw = self.patch_process_shape
        rh, rw = tile_cfg["patch_raw_shape"]
        H, W = tile_cfg["image_raw_shape"]
        Hc, Wc = tile_cfg["patch_reensemble_shape"]
        if hr_shape[-2:] != (H, W):
            raise ValueError(f"image_hr is {hr_shape[-2:]} but tile_cfg.image_raw_shape is {(H, W)}")
        if image_lr.shape[0] != F_:
            raise ValueError(f"image_lr holds {image_lr.shape[0]} frames, image_hr {F_}")

        sched = [tiling.schedule(tile_cfg, self.patch_process_shape, cai_mode, process_num) for _ in range(F_)]   # consumes `random` like F reference calls
        stages = sched[0]
        P = sum(s.bboxs.shape[0] for s in stages)
        n_regular = sum(s.bboxs.shape[0] for s in stages if s.kind == "regular")
        n_random = P - n_regular
        bboxs_np = np.concatenate([s.bboxs for st in sched for s in st], axis=0)                 # [F*P, 4], frame-major

        world, rank = 1, 0
        if shard and torch.distributed.is_available() and torch.distributed.is_initialized():
            world, rank = torch.distributed.get_world_size(), torch.distributed.get_rank()
        if world > 1 and n_random:
            # every rank must blend the SAME random patches: rank 0's draw wins (ranks seeded differently -- seed+rank is
            # common practice -- would otherwise add num_r for different bboxes and finalize against their own starts)
            bboxs_np = _broadcast_bboxs(bboxs_np, dev)
        rois_np = tiling.bboxs_to_feat(bboxs_np, (H, W), (ph, pw))[:, 1:]

        own_np = tiling.shard_patches(F_ * P, rank, world, frames=F_)
        sel = np.nonzero(own_np)[0]
        if hr_host is not None:
            # the frames this rank's patches are cut from -- all of them on one GPU, this rank's own frame when a batch holds one
            # frame per rank -- go up on a copy stream UNDER the coarse pass (only the crop kernel needs them)
            cs = eng["copy_stream"]
            cs.wait_stream(torch.cuda.current_stream(dev))
            image_hr = eng["ws"].f32("hr_in", *hr_shape) if hr_host.dtype == torch.float32 else torch.empty(hr_shape, dtype=hr_host.dtype, device=dev)
            with torch.cuda.stream(cs):
                for f in np.unique(sel // P):
                    image_hr[int(f)].copy_(hr_host[int(f)], non_blocking=True)
                hr_ready = torch.cuda.Event()
                hr_ready.record(cs)

        if world > 1 and F_ % world == 0:
            # the coarse passes of a batch are sharded too (a contiguous block of frames per rank) and all-gathered: ~100 MB of
            # features per frame over NVLink instead of F replicated 0.94-TFLOP passes on every rank
            coarse_feats, coarse_depth = self._coarse_forward_sharded(eng, image_lr, rank, world)
        else:
            coarse_feats, coarse_depth = self.coarse_forward(image_lr)
        if hr_ready is not None:
            torch.cuda.current_stream(dev).wait_event(hr_ready)
            image_hr.record_stream(torch.cuda.current_stream(dev))
        hr = image_hr.float().contiguous()
        preds = eng["ws"].f32("preds", F_ * P, ph, pw)
        self.refine_patches(eng, hr, bboxs_np, rois_np, coarse_feats, coarse_depth, sel, preds, P, trace)

        grid_stages, first = [], 0
        for s in stages:
            if s.kind == "regular":
                grid_stages.append((s.off_process[0], s.off_process[1], s.grid[0], s.grid[1], first))
                first += s.bboxs.shape[0]
        mask = self._mask_dev(eng, "p", (ph, pw))
        rmask = rprep = None
        if n_random:
            rmask = self._mask_dev(eng, "r", (rh, rw))
            key = ("rprep", rh, rw, pw)
            if key not in eng["masks"]:
                eng["masks"][key] = ops.blend_raw_prepare(rmask, pw)                               # once per geometry
            rprep = eng["masks"][key]
            starts_all = torch.from_numpy(np.ascontiguousarray(bboxs_np.reshape(F_, P, 4)[:, n_regular:, [1, 0]])).to(dev)   # [F, n_random, (y0, x0)]
        is_r = cai_mode[0] == "r"
        n_c = Hc * Wc
        per_frame = 2 * n_c + (H * W if is_r else 0)
        packed = own = None
        if shard:
            packed = eng["ws"].f32("packed", F_, per_frame)
            packed.zero_()
            own = torch.from_numpy(own_np).to(dev)
            for f in range(F_):
                pf, of, pk = preds[f * P:(f + 1) * P], own[f * P:(f + 1) * P], packed[f]
                ops.blend_partial_canvas(pf[:n_regular], of[:n_regular].contiguous(), mask, grid_stages, Hc, Wc, pk[:n_c].view(Hc, Wc), pk[n_c:2 * n_c].view(Hc, Wc))
                if is_r and n_random:
                    ops.blend_partial_raw(pf[n_regular:], of[n_regular:].contiguous(), starts_all[f], rmask, ph, pw, H, W, pk[2 * n_c:].view(H, W), prep=rprep)
            if world > 1:                                                                    # ONE sum-reduce of the packed partial canvases of the whole batch
                if self._native_comm is None and os.environ.get("PRV2_NATIVE_REDUCE") == "1":
                    self.use_native_reduce()
                if self._native_comm is not None:                                            # issued by the library on its own ncclComm_t (prv2_reduce_canvas)
                    _lib.call("prv2_reduce_canvas", self._native_comm.handle, ptr(packed), packed.numel(), stream_ptr())
                else:
                    torch.distributed.all_reduce(packed)
        depths, cnts = [], []
        for f in range(F_):
            pf = preds[f * P:(f + 1) * P]
            starts = starts_all[f] if n_random else None
            if not shard:
                avg_c, cnt_c = ops.blend_canvas(pf[:n_regular], mask, grid_stages, Hc, Wc, want_count=True)
                if is_r:
                    depth, cnt = ops.blend_raw(avg_c, cnt_c, pf[n_regular:] if n_random else None, starts, rmask, ph, pw, rh, rw, H, W, prep=rprep)
                else:
                    depth, cnt = avg_c, cnt_c
            else:
                pk = packed[f]
                avg_c, cnt_c = ops.blend_finalize_canvas(pk[:n_c].view(Hc, Wc), pk[n_c:2 * n_c].view(Hc, Wc), mask, grid_stages, Hc, Wc)
                if is_r:
                    depth, cnt = ops.blend_finalize_raw(avg_c, cnt_c, pk[2 * n_c:].view(H, W), starts, rmask, rh, rw, H, W, prep=rprep)
                else:
                    depth, cnt = avg_c, cnt_c
            depths.append(depth)
            cnts.append(cnt)
        depth = depths[0].unsqueeze(0).unsqueeze(0) if F_ == 1 else torch.stack(depths).unsqueeze(1)
        self.last_stats = dict(patches=P, frames=F_, patches_local=int(len(sel)), count_map=cnts[0] if F_ == 1 else torch.stack(cnts),
                               n_regular=n_regular, n_random=n_random)
        if self.output_device == "cpu":
            depth = depth.cpu()
        return depth, {"rgb": image_lr, "depth_pred": depth, "depth_gt": depth_gt, "coarse_prediction": coarse_depth}


def _broadcast_bboxs(bboxs_np: np.ndarray, dev) -> np.ndarray:
    """Rank 0's patch list to every rank (tiny: P x 4 int32), over whatever backend the default group runs on."""
    dist = torch.distributed
    on_dev = dist.get_backend() == "nccl"
    t = torch.from_numpy(np.ascontiguousarray(bboxs_np, dtype=np.int32))
    t = t.to(dev) if on_dev else t.clone()
    dist.broadcast(t, src=0)
    return t.cpu().numpy()


def _default_fine_encoder(encoder_name: str, in_chans: int):
    """timm.create_model(..., features_only=True) like lightweight_refiner.py:259-262 -- only where timm exists.  The
    encoder's arithmetic lives in timm (pinned ``timm==0.9.2`` in the reference's environment.yml:27), which is not
    installable offline: parity of the ENCODER is unpinned; pass ``fine_encoder=`` to supply one."""
    try:
        import timm
    except Exception as e:
        raise NotImplementedError(
            f"refiner.fine_branch encoder {encoder_name!r} is a timm model and timm is not installed: construct PatchRefinerPlus with "
            "fine_encoder=<torch module returning the 5 feature maps> (see INTEGRATION.md)") from e
    return timm.create_model(encoder_name, pretrained=False, features_only=True, in_chans=in_chans)


@MODELS.register_module()
class PatchRefinerPlus(PatchRefiner):
    """B200-native PatchRefinerPlus (the V2 family: DA2 coarse branch + LightWeightRefiner + BiDirectionalFusion),
    infer path only (estimator/models/patchrefinerplus.py:367-533).  Tiling, cropping, ROI gather, blending and the
    sharded multi-GPU path are PatchRefiner's; ``refine_patches`` runs the light-weight encoder (a PyTorch module the
    caller supplies or timm builds -- library code, not one of this package's kernels) and hands its features to
    ``BiDirectionalFusionB200``, where 99 % of the V2 refiner's FLOPs are."""

    def __init__(self, config, precision: str = "fp32", patch_batch: int = 8, output_device: str = "cpu", fine_encoder: Optional[nn.Module] = None):
        nn.Module.__init__(self)
        from .bifusion import C2F_TYPES, bifusion_weight_spec
        if hasattr(config, "to_dict"):
            config = config.to_dict()
        self.config = config
        if _get(config, "pretrain_stage", False):
            raise NotImplementedError("pretrain_stage=True (training-time hack path, patchrefinerplus.py:383-425) is out of scope")
        self.min_depth = _get(config, "min_depth")
        self.max_depth = float(_get(config, "max_depth"))
        self.patch_process_shape = tuple(_get(config, "patch_process_shape"))
        self.tile_cfg = self.prepare_tile_cfg(_get(config, "image_raw_shape"), _get(config, "patch_split_num"))
        cb, rf = _get(config, "coarse_branch"), _get(config, "refiner")
        fb, fu = _get(rf, "fine_branch"), _get(rf, "fusion_model")
        if _get(cb, "type") != "DA2":
            raise NotImplementedError(f"coarse_branch.type={_get(cb, 'type')!r}: only the DA2 (DepthAnythingV2) coarse branch is implemented")
        if _get(fb, "type") != "LightWeightRefiner" or _get(fb, "with_decoder", False):
            raise NotImplementedError("refiner.fine_branch must be LightWeightRefiner(with_decoder=False)")
        if _get(fu, "type") not in ("BiDirectionalFusion", "BiDirectionalFusionHeavy") or _get(fu, "coarse2fine_type") not in C2F_TYPES or _get(fu, "glb_att", False):
            raise NotImplementedError(f"fusion_model {_get(fu, 'type')!r}/{_get(fu, 'coarse2fine_type')!r}: BiDirectionalFusion[Heavy] with coarse2fine_type in {sorted(C2F_TYPES)}")
        self._fu_heavy = _get(fu, "type") == "BiDirectionalFusionHeavy"
        self._cb_cfg, self._fu_cfg = dict(_get(cb, "model_cfg")), fu
        self.coarse_condition = bool(_get(fb, "coarse_condition", True))
        self.fusion_feat_level = int(_get(config, "fusion_feat_level"))
        self.strategy_refiner_target = _get(config, "strategy_refiner_target")
        self.pre_norm_bbox = _get(config, "pre_norm_bbox", True)
        self.resizer = _Resizer(self.patch_process_shape)
        if tuple(self.resizer.size) != tuple(self.patch_process_shape):
            raise NotImplementedError("patch_process_shape must be a multiple of 14 for the DA2 coarse branch")
        assert precision in ("bf16", "fp32")
        self.precision, self.patch_batch, self.output_device = precision, int(patch_batch), output_device
        enc_name, in_chans = str(_get(fb, "encoder_name", "")), 4 if self.coarse_condition else 3
        # convnext encoders return four maps (strides 4..32); LightWeightRefiner adds the stride-2 level with its own
        # ``upsample_convx`` = ConvTranspose2d(encoder_channels[1] -> encoder_channels[0], k = s = 2) + ReLU (lightweight_refiner.py:276-283, 307-314)
        self._convx = "convnext" in enc_name
        self._enc_chl = list(_get(fb, "encoder_channels", None) or [16, 24, 40, 112, 960])
        # mobilenetv4_conv_small (configs/patchrefinerv2_dav2/plus_mobile_*): this package's own kernels (mnv4.py).  Any other
        # encoder: a PyTorch module the caller supplies, or timm where it is installed (library code, parity unpinned either way).
        self._native_enc = fine_encoder is None and enc_name.startswith("mobilenetv4_conv_small")
        self._enc_in_chans = in_chans
        self._weights = OrderedDict()
        if self._native_enc:
            from .mnv4 import DEFAULT_MEAN, DEFAULT_STD, mnv4_conv_small_spec
            self.refiner_fine_encoder = None
            self._enc_mean, self._enc_std = DEFAULT_MEAN, DEFAULT_STD
            for k, shp in mnv4_conv_small_spec(in_chans).items():           # a fresh BatchNorm: weight 1, running_var 1, the rest 0
                self._weights[self.ENC_PREFIX + k] = torch.ones(shp) if k.endswith(("bn.weight", "bn1.weight", "running_var")) else torch.zeros(shp)
        else:
            self.refiner_fine_encoder = fine_encoder if fine_encoder is not None else _default_fine_encoder(enc_name, in_chans)
            self.refiner_fine_encoder.eval()
            cfgd = getattr(self.refiner_fine_encoder, "default_cfg", None) or {}
            self._enc_mean = tuple(cfgd.get("mean", (0.485, 0.456, 0.406)))
            self._enc_std = tuple(cfgd.get("std", (0.229, 0.224, 0.225)))
        if self._convx:
            self._weights["refiner_fine_branch.upsample_convx.0.weight"] = torch.zeros(self._enc_chl[1], self._enc_chl[0], 2, 2)
            self._weights["refiner_fine_branch.upsample_convx.0.bias"] = torch.zeros(self._enc_chl[0])
        for k, shp in dav2_weight_spec(self._cb_cfg["encoder"], self._cb_cfg["features"], self._cb_cfg["out_channels"]).items():
            self._weights["coarse_branch." + k] = torch.zeros(shp)
        keys = ("coarse_chl", "fine_chl", "fine_chl_after_coarse2fine", "temp_chl", "dec_chl")
        for k, shp in bifusion_weight_spec(*[_get(fu, x) for x in keys], coarse2fine_type=_get(fu, "coarse2fine_type"), heavy=self._fu_heavy).items():
            self._weights["refiner_fusion_model." + k] = torch.zeros(shp)
        # checkpoints named by the config, in the reference's order and strictness (patchrefinerplus.py:121-126, 138-141, 202-205):
        # the coarse-branch files strict=True over the branch, `pretrained` / `whole_pretrained` strict=False
        path = _get(cb, "pretrained")
        if path:
            self._load_branch("coarse_branch.", torch.load(path, map_location="cpu"))
        path = _get(config, "pretrain_coarse_model")
        if path:
            self._load_branch("coarse_branch.", torch.load(path, map_location="cpu")["model_state_dict"])
        for key in ("pretrained", "whole_pretrained"):
            path = _get(config, key)
            if path:
                self._load(torch.load(path, map_location="cpu")["model_state_dict"], strict=False)
        self._engine = None
        self._device = torch.device("cpu")
        self.last_stats = {}

    ENC_PREFIX = "refiner_fine_branch.refiner_encoder."

    def _load(self, sd, strict):
        if self._native_enc:
            return super()._load(sd, strict)
        enc = {k[len(self.ENC_PREFIX):]: v for k, v in sd.items() if k.startswith(self.ENC_PREFIX)}
        rest = {k: v for k, v in sd.items() if not k.startswith(self.ENC_PREFIX)}
        res = super()._load(rest, strict)
        missing, unexpected = list(res.missing_keys), list(res.unexpected_keys)
        if enc or strict:
            r2 = self.refiner_fine_encoder.load_state_dict(enc, strict=strict)
            missing += [self.ENC_PREFIX + k for k in r2.missing_keys]
            unexpected += [self.ENC_PREFIX + k for k in r2.unexpected_keys]
        return torch.nn.modules.module._IncompatibleKeys(missing, unexpected)

    def state_dict(self, *args, **kwargs):
        sd = OrderedDict((k, v.clone()) for k, v in self._weights.items())
        if not self._native_enc:
            for k, v in self.refiner_fine_encoder.state_dict().items():
                sd[self.ENC_PREFIX + k] = v.detach().cpu().clone()
        return sd

    def get_save_dict(self):                        # patchrefinerplus.py:215-216 (keeps the coarse branch)
        return self.state_dict()

    def _build_engine(self, device):
        if device.type != "cuda":
            raise RuntimeError("patchrefinerv2_b200 runs on CUDA (sm_100a) only; there is no CPU path. Call .cuda() first.")
        from .bifusion import BiDirectionalFusionB200
        self._warn_unloaded()
        _lib.load()
        x3 = self.precision == "fp32"
        sd, fu = self._weights, self._fu_cfg
        keys = ("coarse_chl", "fine_chl", "fine_chl_after_coarse2fine", "temp_chl", "dec_chl")
        enc = None
        if self._native_enc:
            from .mnv4 import MobileNetV4ConvSmallB200
            enc = MobileNetV4ConvSmallB200(sd, self.ENC_PREFIX, self._enc_in_chans, x3, device, self._enc_mean, self._enc_std)
        else:
            self.refiner_fine_encoder.to(device)
        up_convx = None
        if self._convx:
            from .nn import GemmLayer
            wt, c0 = sd["refiner_fine_branch.upsample_convx.0.weight"].detach().float(), self._enc_chl[0]          # [Cin, C0, 2, 2]
            up_convx = GemmLayer([(0, 0, 0, wt.permute(2, 3, 1, 0).reshape(4 * c0, wt.shape[0]))], 1, 4 * c0, x3, device, epi=_lib.EPI_SHUFFLE,
                                 act=_lib.ACT_RELU, bias=sd["refiner_fine_branch.upsample_convx.0.bias"].detach().float(), shuffle_k=2,
                                 name="lwr.upsample_convx")
        return dict(
            encoder=enc, up_convx=up_convx,
            coarse=DepthAnythingV2B200(sd, "coarse_branch.", self._cb_cfg["encoder"], self._cb_cfg["features"], self._cb_cfg["out_channels"], self.max_depth, x3, device),
            fusion=BiDirectionalFusionB200(sd, "refiner_fusion_model.", *[_get(fu, k) for k in keys], coarse2fine_type=_get(fu, "coarse2fine_type"),
                                           x3=x3, device=device, heavy=self._fu_heavy),
            ws=Workspace(device, x3), device=device, masks={},
            enc_mean=torch.tensor(self._enc_mean, device=device).view(1, -1, 1, 1), enc_std=torch.tensor(self._enc_std, device=device).view(1, -1, 1, 1))

    def refine_patches(self, eng, image_hr, bboxs_np, rois_np, coarse_feats, coarse_depth, sel: np.ndarray, preds: torch.Tensor, P: int, trace=None):
        """patchrefinerplus.py:330-365 for the work items ``sel`` (see PatchRefiner.refine_patches), ``patch_batch`` at a time."""
        dev = eng["device"]
        ph, pw = self.patch_process_shape
        level = self.fusion_feat_level
        x3 = self.precision == "fp32"
        ws = eng["ws"]
        for s in range(0, len(sel), self.patch_batch):
            idx = sel[s:s + self.patch_batch]
            pb = len(idx)
            crops, c_roi, d_roi = self._gather_batch(eng, image_hr, bboxs_np, rois_np, coarse_feats, coarse_depth, idx, P)
            # LightWeightRefiner.forward (lightweight_refiner.py:285-322): normalise, condition on the coarse depth, encode
            if eng["encoder"] is not None:
                f_acts = eng["encoder"].forward(crops, d_roi if self.coarse_condition else None, ws)
            else:
                x = (crops - eng["enc_mean"]) / eng["enc_std"]
                feats = list(self.refiner_fine_encoder(torch.cat([x, d_roi], dim=1) if self.coarse_condition else x))
                f_acts = [Act.from_nchw(f, x3) for f in feats]
            if self._convx:                                                                        # :307-314
                top = ws.act("enc_up_convx", pb, f_acts[0].H * 2, f_acts[0].W * 2, self._enc_chl[0])
                eng["up_convx"]([f_acts[0]], out=top)                                              # ConvTranspose2d(k = s = 2) + bias -> ReLU
                f_acts = [top] + f_acts
            top = f_acts[0]
            up = ops.resize_bilinear(top, ws.act("enc_up", pb, top.H * 2, top.W * 2, top.C))       # :316-318 / :312-313
            r_feats = ([up] + f_acts)[::-1]                                                        # :320, coarsest first
            if self.strategy_refiner_target == "offset_fine":
                base = torch.zeros_like(d_roi)                                                     # :321: the refiner depth is zeros
            elif self.strategy_refiner_target == "offset_coarse":
                base = d_roi
            else:
                base = None
            c_list = c_roi[-level:][::-1]
            f_list = r_feats[-level:][::-1]
            pred = eng["fusion"].forward(c_list, f_list, d_roi, None, base, trace)
            if self.strategy_refiner_target == "direct":                                           # patchrefinerplus.py:362-363
                pred = torch.sigmoid(pred) * self.max_depth
            preds[torch.from_numpy(idx).to(dev, non_blocking=True)] = pred.reshape(pb, ph, pw)
            if trace is not None:
                trace.setdefault("crops", crops.clone()); trace.setdefault("roi_depth", d_roi.clone())
                trace.setdefault("fine_feats", [a.to_nchw() for a in r_feats])
                trace = None
