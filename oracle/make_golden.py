"""Generate tests/golden/*.npz by running the UNMODIFIED reference (read-only /root/reference,
through oracle/ref_shim.py) on seeded synthetic inputs.  Run here (no GPU needed):

    python -m oracle.make_golden

Two families of fixtures:
  tiny_<mode>.npz  -- reference PatchRefiner (DAv2 ViT-S coarse + ViT-S refiner + FusionUnet),
                      224x224 patches, 432x768 frame, 2x2 split, modes m1 / m2 / r4: bboxes,
                      bboxs_feat, per-patch predictions, count map and final depth (full fp32).
  geom_<ph>x<pw>_<mode>.npz -- reference tiling + blend at BASELINE sizes (2160x3840, 4x4,
                      448x448 and 384x512 patches, m1 / m2 / r32) with the networks replaced by
                      oracle.pr_oracle.fake_prediction: bboxes, bboxs_feat, and sha256 + strided
                      subsample of the count map and the blended depth.
The weights / frames are NOT stored; they are regenerated from seeds by oracle.pr_oracle and
their sha256 is stored so drift in the generators is detected.
"""
from __future__ import annotations

import os
import random
import sys
import tempfile

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)

from oracle import pr_oracle as O  # noqa: E402
from oracle import ref_shim  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")
TINY = dict(encoder="vits", patch_process_shape=(224, 224), image_raw_shape=(432, 768), patch_split_num=(2, 2))
TINY_MODES = (("m1", 2), ("m2", 2), ("r4", 2))
GEOM_MODES = (("m1", 4), ("m2", 4), ("r32", 4))
SUB = 8


def build_reference(cfg, sd):
    d = tempfile.mkdtemp()
    cp, fp = os.path.join(d, "c.pth"), os.path.join(d, "f.pth")
    torch.save({k[len("coarse_branch."):]: v for k, v in sd.items() if k.startswith("coarse_branch.")}, cp)
    torch.save({k[len("refiner_fine_branch."):]: v for k, v in sd.items() if k.startswith("refiner_fine_branch.")}, fp)
    ref = ref_shim.build_reference_patchrefiner(cfg, cp, fp)
    res = ref.load_state_dict(sd, strict=False)
    assert not res.missing_keys and not res.unexpected_keys, res
    return ref


def sd_digest(sd):
    import hashlib
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].numpy()).tobytes())
    return h.hexdigest()


def tiny():
    cfg = O.make_config(**TINY)
    sd = O.init_patchrefiner_state_dict(cfg, 0)
    ref = build_reference(cfg, sd)
    lr, hr = O.synthetic_frame(cfg, 1)
    rec = {}
    orig_post = ref.coarse_postprocess_test
    orig_inf = ref.infer_forward

    def post(coarse_prediction, coarse_features, bboxs, bboxs_feat):
        rec["bboxs"].append(bboxs.clone())
        rec["bboxs_feat"].append(bboxs_feat.clone())
        out = orig_post(coarse_prediction=coarse_prediction, coarse_features=coarse_features, bboxs=bboxs, bboxs_feat=bboxs_feat)
        if "roi_depth" not in rec:
            rec["roi_depth"] = out["coarse_depth_roi"][:2].clone()
            rec["roi_feat0"] = out["coarse_feats_roi"][0][:2].clone()
            rec["roi_feat5"] = out["coarse_feats_roi"][5][:2].clone()
        return out

    def inf(imgs_crop, bbox_feat_forward, tile_temp, coarse_temp_dict):
        p = orig_inf(imgs_crop, bbox_feat_forward, tile_temp, coarse_temp_dict)
        rec["preds"].append(p.clone())
        if "crop0" not in rec:
            rec["crop0"] = imgs_crop[:1].clone()
        return p

    ref.coarse_postprocess_test = post
    ref.infer_forward = inf
    for mode, pn in TINY_MODES:
        rec.clear()
        rec.update(bboxs=[], bboxs_feat=[], preds=[])
        random.seed(1)
        with torch.no_grad():
            depth, log = ref(mode="infer", image_lr=lr, image_hr=hr, cai_mode=mode, process_num=pn, tile_cfg=None)
        np.savez_compressed(
            os.path.join(OUT, f"tiny_{mode}.npz"),
            mode=mode, process_num=pn, sd_sha=sd_digest(sd), frame_sha=O.sha256_f32(hr.numpy()),
            bboxs=torch.cat(rec["bboxs"]).numpy(), bboxs_feat=torch.cat(rec["bboxs_feat"]).numpy(),
            preds=torch.cat(rec["preds"]).numpy(), depth=depth.numpy(),
            coarse=log["coarse_prediction"].numpy(),
            roi_depth=rec["roi_depth"].numpy(), roi_feat0=rec["roi_feat0"].numpy(), roi_feat5_sha=O.sha256_f32(rec["roi_feat5"].numpy()),
            crop0=rec["crop0"].numpy())
        print("tiny", mode, tuple(depth.shape), float(depth.min()), float(depth.max()))


def geom(ph, pw):
    cfg = O.make_config("vits", (ph, pw), (2160, 3840), (4, 4))
    # a reference PatchRefiner instance is needed only for its tiling / blend methods; build it
    # at 224x224 weights-wise (weights are never used: the three network entry points are patched)
    sd = O.init_patchrefiner_state_dict(cfg, 0)
    ref = build_reference(cfg, sd)
    _, hr = O.synthetic_frame(cfg, 1)
    lr = torch.zeros(1, 3, ph, pw)
    rec = {}
    ref.coarse_forward = lambda image_lr: ([torch.zeros(1, 1, 8, 8)], torch.zeros(1, 1, ph, pw))

    def post(coarse_prediction, coarse_features, bboxs, bboxs_feat):
        rec["bboxs"].append(bboxs.clone())
        rec["bboxs_feat"].append(bboxs_feat.clone())
        rec["cursor"] = 0
        P = bboxs.shape[0]
        return {"coarse_depth_roi": torch.zeros(P, 1, 1, 1), "coarse_feats_roi": [torch.zeros(P, 1, 1, 1)]}

    def inf(imgs_crop, bbox_feat_forward, tile_temp, coarse_temp_dict):
        n = imgs_crop.shape[0]
        rows = rec["bboxs"][-1][rec["cursor"]:rec["cursor"] + n].tolist()
        rec["cursor"] += n
        return torch.stack([O.fake_prediction(r, ph, pw) for r in rows], dim=0)

    ref.coarse_postprocess_test = post
    ref.infer_forward = inf
    for mode, pn in GEOM_MODES:
        rec.clear()
        rec.update(bboxs=[], bboxs_feat=[])
        random.seed(1)
        with torch.no_grad():
            depth, log = ref(mode="infer", image_lr=lr, image_hr=hr, cai_mode=mode, process_num=pn, tile_cfg=None)
        # the RunningAverageMap is not returned by forward(); rebuild the count map with the
        # oracle (bit-identical to the reference, asserted by tests/test_oracle_vs_reference.py)
        go = O.GeometryOracle((ph, pw), (2160, 3840), (4, 4))
        random.seed(1)
        d2, _, avg = go.infer(lr, hr, None, mode, pn)
        assert torch.equal(d2, depth), "oracle blend differs from the reference"
        cnt = avg.count_map.numpy()
        dn = depth[0, 0].numpy()
        np.savez_compressed(
            os.path.join(OUT, f"geom_{ph}x{pw}_{mode}.npz"),
            mode=mode, process_num=pn,
            bboxs=torch.cat(rec["bboxs"]).numpy(), bboxs_feat=torch.cat(rec["bboxs_feat"]).numpy(),
            depth_sha=O.sha256_f32(dn), depth_sub=dn[::SUB, ::SUB].copy(), depth_shape=np.array(dn.shape),
            count_sha=O.sha256_f32(cnt), count_sub=cnt[::SUB, ::SUB].copy())
        print("geom", ph, pw, mode, dn.shape, float(dn.min()), float(dn.max()))


#: BiDirectionalFusion fixture (V2 family, SURVEY.md row a9'): the reference module itself on seeded stand-in features
BIFUSION = dict(coarse_chl=[32, 256, 256, 256, 256, 256], fine_chl=[32, 32, 64, 96, 960],
                fine_chl_after_coarse2fine=[32, 256, 256, 256, 256, 256], temp_chl=[32, 64, 64, 128, 256, 512], dec_chl=[512, 256, 128, 64, 32])
BIFUSION_SIZES_C = [(64, 64), (36, 36), (18, 18), (9, 9), (5, 5), (3, 3)]      # coarse ROI maps (resized to the fine sizes inside)
BIFUSION_SIZES_F = [(64, 64), (32, 32), (16, 16), (8, 8), (4, 4), (2, 2)]
BIFUSION_TYPES = ("coarse-gated", "coarse-fusion", "self-agg")


def bifusion():
    """bifusion_<type>.npz: output of the reference's BiDirectionalFusion (bi_directional_fusion_model.py) on
    oracle.synthetic_fusion_inputs(seed 3), weights init_bidirectional_fusion_state_dict(seed 5)."""
    ref_shim.install()
    from estimator.models.blocks.bi_directional_fusion_model import BiDirectionalFusion
    for t in BIFUSION_TYPES:
        m = BiDirectionalFusion(encoder_name="golden", coarse2fine_type=t, **{k: list(v) for k, v in BIFUSION.items()}).eval()
        sd = O.init_bidirectional_fusion_state_dict(seed=5, coarse2fine_type=t, **BIFUSION)
        m.load_state_dict(sd, strict=True)
        c, f, p1, p2 = O.synthetic_fusion_inputs(BIFUSION["coarse_chl"], BIFUSION["fine_chl"], BIFUSION_SIZES_C, BIFUSION_SIZES_F, 2, 3)
        with torch.no_grad():
            out = m(c_feat=[x.clone() for x in c], f_feat=[x.clone() for x in f], pred1=p1, pred2=p2, update_base=p1)
            off = m(c_feat=[x.clone() for x in c], f_feat=[x.clone() for x in f], pred1=p1, pred2=p2, update_base=None)
        np.savez_compressed(os.path.join(OUT, f"bifusion_{t}.npz"), coarse2fine_type=t, depth=out.numpy(), offset=off.numpy(),
                            sd_sha=sd_digest(sd), pred1_sha=O.sha256_f32(p1.numpy()), keys=np.array(sorted(m.state_dict().keys())))
        print("bifusion", t, tuple(out.shape), float(out.mean()), float(off.std()))


ZOE_HEAD_SIZES = [(12, 16), (12, 16), (24, 32), (48, 64), (96, 128)]      # btlnck, then the four decoder blocks coarse -> fine


def zoe_head_inputs(seed=2, B=2):
    g = torch.Generator().manual_seed(seed)
    btl = torch.randn(B, 256, *ZOE_HEAD_SIZES[0], generator=g)
    xb = [torch.randn(B, 256, *s, generator=g) for s in ZOE_HEAD_SIZES[1:]]
    outc = torch.relu(torch.randn(B, 32, 192, 256, generator=g))
    rel = torch.rand(B, 192, 256, generator=g) * 5
    return rel, btl, xb, outc


def zoe_head():
    """zoe_head.npz: the reference's ZoeDepth metric-bins head (zoedepth_v1.py:173-233) with a dummy core, fed through hack_feature."""
    ref_shim.install()
    from zoedepth.models.zoedepth.zoedepth_v1 import ZoeDepth

    class DummyCore(torch.nn.Module):
        output_channels = [256] * 5

        def freeze_encoder(self, *a, **k):
            pass

    c = O.ZOE_HEAD_CFG
    m = ZoeDepth(DummyCore(), n_bins=c["n_bins"], bin_centers_type=c["bin_centers_type"], bin_embedding_dim=c["bin_embedding_dim"], min_depth=1e-3,
                 max_depth=80, n_attractors=list(c["n_attractors"]), attractor_alpha=c["attractor_alpha"], attractor_gamma=c["attractor_gamma"],
                 attractor_kind=c["attractor_kind"], attractor_type=c["attractor_type"], min_temp=c["min_temp"], max_temp=c["max_temp"]).eval()
    sd = O.init_zoe_head_state_dict([256] * 5, 7)
    m.load_state_dict(sd, strict=False)
    rel, btl, xb, outc = zoe_head_inputs()
    with torch.no_grad():
        r = m(None, hack_feature=[rel, [btl] + xb + [outc]], return_final_centers=True)
    np.savez_compressed(os.path.join(OUT, "zoe_head.npz"), depth=r["metric_depth"].numpy(), centers_sub=r["bin_centers"][:, :, ::16, ::16].numpy(),
                        sd_sha=sd_digest(sd), rel_sha=O.sha256_f32(rel.numpy()))
    print("zoe_head", tuple(r["metric_depth"].shape), float(r["metric_depth"].mean()))


PLUS_MODES = (("m1", 2), ("r2", 2))


def plus():
    """plus_<mode>.npz: the reference PatchRefinerPlus (patchrefinerplus.py) end to end -- DA2 ViT-S coarse branch with 256 decoder
    features, LightWeightRefiner around oracle.ToyFineEncoder (timm stand-in), BiDirectionalFusion coarse-gated -- on the tiny frame."""
    cfg = O.make_plus_config()
    sd = O.init_patchrefinerplus_state_dict(cfg, 0)
    d = tempfile.mkdtemp()
    cp = os.path.join(d, "c.pth")
    torch.save({k[len("coarse_branch."):]: v for k, v in sd.items() if k.startswith("coarse_branch.")}, cp)
    ref = ref_shim.build_reference_patchrefinerplus(cfg, cp, lambda: O.ToyFineEncoder(3))
    res = ref.load_state_dict(sd, strict=False)
    assert not res.missing_keys and not res.unexpected_keys, res
    lr, hr = O.synthetic_frame(cfg, 1)
    for mode, pn in PLUS_MODES:
        random.seed(1)
        with torch.no_grad():
            depth, log = ref(mode="infer", image_lr=lr, image_hr=hr, cai_mode=mode, process_num=pn, tile_cfg=None)
        np.savez_compressed(os.path.join(OUT, f"plus_{mode}.npz"), mode=mode, process_num=pn, depth=depth.numpy(),
                            coarse=log["coarse_prediction"].numpy(), sd_sha=sd_digest(sd), frame_sha=O.sha256_f32(hr.numpy()),
                            keys=np.array(sorted(ref.state_dict().keys())))
        print("plus", mode, tuple(depth.shape), float(depth.mean()))


def plus_convx():
    """plus_convx_r2.npz: the same with a convnext-named encoder (oracle.ToyConvNeXtEncoder: four maps) -- the reference then runs its
    4-channel ``stem_0`` surgery (patchrefinerplus.py:194-200) and LightWeightRefiner's ``upsample_convx`` stage (:276-283, 307-314)."""
    cfg = O.make_plus_config(convnext=True)
    sd = O.init_patchrefinerplus_state_dict(cfg, 0)
    d = tempfile.mkdtemp()
    cp = os.path.join(d, "c.pth")
    torch.save({k[len("coarse_branch."):]: v for k, v in sd.items() if k.startswith("coarse_branch.")}, cp)
    ref = ref_shim.build_reference_patchrefinerplus(cfg, cp, lambda: O.ToyConvNeXtEncoder(3))
    res = ref.load_state_dict(sd, strict=False)
    assert not res.missing_keys and not res.unexpected_keys, res
    lr, hr = O.synthetic_frame(cfg, 1)
    random.seed(1)
    with torch.no_grad():
        depth, log = ref(mode="infer", image_lr=lr, image_hr=hr, cai_mode="r2", process_num=2, tile_cfg=None)
    np.savez_compressed(os.path.join(OUT, "plus_convx_r2.npz"), mode="r2", process_num=2, depth=depth.numpy(), coarse=log["coarse_prediction"].numpy(),
                        sd_sha=sd_digest(sd), frame_sha=O.sha256_f32(hr.numpy()), keys=np.array(sorted(ref.state_dict().keys())))
    print("plus_convx r2", tuple(depth.shape), float(depth.mean()))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    torch.set_num_threads(os.cpu_count() or 1)
    if "--plus-convx-only" in sys.argv:
        plus_convx()
        sys.exit(0)
    if "--bifusion-only" in sys.argv:
        bifusion()
        sys.exit(0)
    if "--zoe-head-only" in sys.argv:
        zoe_head()
        sys.exit(0)
    if "--plus-only" in sys.argv:
        plus()
        sys.exit(0)
    plus()
    plus_convx()
    zoe_head()
    bifusion()
    tiny()
    geom(448, 448)
    geom(384, 512)
