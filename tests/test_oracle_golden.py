"""The restated CPU oracle against the fixtures generated from the reference itself
(oracle/make_golden.py).  Runs without /root/reference."""
import hashlib
import os
import random

import numpy as np
import pytest
import torch

from oracle import pr_oracle as O

TINY_MODES = [("m1", 2), ("m2", 2), ("r4", 2)]


def _sd_digest(sd):
    h = hashlib.sha256()
    for k in sorted(sd):
        h.update(k.encode())
        h.update(np.ascontiguousarray(sd[k].numpy()).tobytes())
    return h.hexdigest()


@pytest.mark.parametrize("mode,pn", TINY_MODES)
def test_tiny_end_to_end_matches_reference_golden(golden_dir, tiny_setup, mode, pn):
    cfg, sd, lr, hr = tiny_setup
    g = np.load(os.path.join(golden_dir, f"tiny_{mode}.npz"))
    assert str(g["sd_sha"]) == _sd_digest(sd), "weight generator drifted from the golden run"
    assert str(g["frame_sha"]) == O.sha256_f32(hr.numpy()), "frame generator drifted from the golden run"
    orc = O.PatchRefinerOracle(cfg, sd)
    rec = {}
    random.seed(1)
    depth, coarse, avg = orc.infer(lr, hr, None, mode, pn, record=rec)
    assert np.array_equal(torch.cat(rec["bboxs"]).numpy(), g["bboxs"])                       # bit-exact ints
    assert np.array_equal(torch.cat(rec["bboxs_feat"]).numpy(), g["bboxs_feat"])             # bit-exact float32
    # the reference ran on this same CPU build, so the float paths agree to the last bit here; on a
    # different host ISA allow the 1e-3 relative budget of the north star
    np.testing.assert_allclose(coarse.numpy(), g["coarse"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(torch.cat(rec["preds"]).numpy(), g["preds"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(depth.numpy(), g["depth"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(rec["roi_first"]["depth"].numpy(), g["roi_depth"], rtol=1e-5, atol=1e-5)
    np.testing.assert_allclose(rec["roi_first"]["crop"][:1].numpy(), g["crop0"], rtol=0, atol=1e-6)


@pytest.mark.parametrize("shape", [(448, 448), (384, 512)])
@pytest.mark.parametrize("mode,pn", [("m1", 4), ("m2", 4), ("r32", 4)])
def test_geometry_and_blend_at_baseline_size(golden_dir, shape, mode, pn):
    ph, pw = shape
    g = np.load(os.path.join(golden_dir, f"geom_{ph}x{pw}_{mode}.npz"))
    cfg = O.make_config("vits", shape, (2160, 3840), (4, 4))
    _, hr = O.synthetic_frame(cfg, 1)
    go = O.GeometryOracle(shape, (2160, 3840), (4, 4))
    rec = {"bboxs": [], "bboxs_feat": [], "preds": []}
    random.seed(1)
    depth, _, avg = go.infer(torch.zeros(1, 3, ph, pw), hr, None, mode, pn, record=rec)
    assert np.array_equal(torch.cat(rec["bboxs"]).numpy(), g["bboxs"])
    assert np.array_equal(torch.cat(rec["bboxs_feat"]).numpy(), g["bboxs_feat"])
    d = depth[0, 0].numpy()
    assert tuple(d.shape) == tuple(g["depth_shape"])
    assert O.sha256_f32(d) == str(g["depth_sha"])                   # blend is pure fp32 element-wise: bit-exact everywhere
    assert O.sha256_f32(avg.count_map.numpy()) == str(g["count_sha"])
    assert np.array_equal(d[::8, ::8], g["depth_sub"])


def test_patch_counts():
    """m1 = Sh*Sw, m2 = 16+12+12+9 = 49 at 4x4, r32 adds 32."""
    from patchrefinerv2_b200 import tiling
    tc = tiling.prepare_tile_cfg((448, 448), (2160, 3840), (4, 4))
    n = lambda m: sum(s.bboxs.shape[0] for s in tiling.schedule(tc, (448, 448), m, 4, random.Random(0)))
    assert (n("m1"), n("m2"), n("r32")) == (16, 49, 81)
    tc8 = tiling.prepare_tile_cfg((448, 448), (4320, 7680), (8, 8))
    assert sum(s.bboxs.shape[0] for s in tiling.schedule(tc8, (448, 448), "r128", 4, random.Random(0))) == 353


def test_index_restatements_bit_exact_vs_torch():
    import torch.nn.functional as F
    from torchvision.ops import roi_align
    g = torch.Generator().manual_seed(3)
    for hi, wi, ho, wo in [(540, 960, 448, 448), (216, 384, 224, 224), (448, 448, 432, 768), (896, 896, 1080, 1920)]:
        x = torch.rand(2, hi, wi, generator=g)
        ref = F.interpolate(x[None], (ho, wo), mode="bilinear", align_corners=True)[0].numpy()
        assert np.array_equal(ref, O.np_bilinear_ac(x.numpy(), ho, wo))
    for ni, no in [(448, 540), (448, 960), (1792, 2160), (1792, 3840), (224, 216), (384, 540), (512, 960)]:
        x = torch.arange(ni, dtype=torch.float32)[None, None, None, :]
        ref = F.interpolate(x, (1, no))[0, 0, 0].long().numpy()
        assert np.array_equal(ref, O.np_nearest_index(ni, no))
    for H, W, ph, pw, h, w in [(2160, 3840, 448, 448, 448, 448), (2160, 3840, 448, 448, 16, 16), (432, 768, 224, 224, 8, 8),
                               (2160, 3840, 384, 512, 96, 128)]:
        feat = torch.rand(1, 3, h, w, generator=g)
        bb = torch.tensor([[0, 0, W // 4, H // 4], [1234 % (W - W // 4), 777 % (H - H // 4), 1234 % (W - W // 4) + W // 4, 777 % (H - H // 4) + H // 4],
                           [W - W // 4 - 1, H - H // 4 - 1, W - 1, H - 1]]).int()
        bf = O.bboxs_to_feat(bb, (H, W), (ph, pw))
        rois = bf.clone()
        rois[:, 0] = 0
        ref = roi_align(feat, rois, (h, w), h / ph, aligned=True).numpy()
        for i in range(bb.shape[0]):
            assert np.array_equal(ref[i], O.np_roi_align_1s(feat[0].numpy(), bf[i, 1:].numpy(), h / ph, h, w))


def test_zoe_bins_head_matches_reference_golden(golden_dir):
    """ZoeDepth metric-bins head (oracle only so far) against the output the reference's ZoeDepth produced (oracle/make_golden.py)."""
    from oracle.make_golden import sd_digest, zoe_head_inputs
    g = np.load(os.path.join(golden_dir, "zoe_head.npz"))
    sd = O.init_zoe_head_state_dict([256] * 5, 7)
    rel, btl, xb, outc = zoe_head_inputs()
    assert str(g["sd_sha"]) == sd_digest(sd) and str(g["rel_sha"]) == O.sha256_f32(rel.numpy())
    tr = {}
    with torch.no_grad():
        d, _ = O.zoe_bins_head(sd, "", rel, btl, xb, outc, O.ZOE_HEAD_CFG, tr)
    np.testing.assert_allclose(d.numpy(), g["depth"], rtol=1e-3, atol=1e-4)
    np.testing.assert_allclose(tr["bin_centers"][:, :, ::16, ::16].numpy(), g["centers_sub"], rtol=1e-3, atol=1e-4)
    assert float(tr["probs"].sum(1).sub(1).abs().max()) < 1e-5
