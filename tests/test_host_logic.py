"""Host-side product logic (tiling, masks, registry, C-ABI surface) on CPU."""
import os
import random
import re

import numpy as np
import pytest
import torch

from oracle import pr_oracle as O
from patchrefinerv2_b200 import _lib, masks, tiling
from patchrefinerv2_b200.registry import MODELS, build_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("shape", [(448, 448), (384, 512)])
@pytest.mark.parametrize("mode", ["m1", "m2", "r32"])
def test_schedule_matches_reference_golden_bit_exact(golden_dir, shape, mode):
    g = np.load(os.path.join(golden_dir, f"geom_{shape[0]}x{shape[1]}_{mode}.npz"))
    tc = tiling.prepare_tile_cfg(shape, (2160, 3840), (4, 4))
    random.seed(1)
    stages = tiling.schedule(tc, shape, mode, int(g["process_num"]))
    bb = np.concatenate([s.bboxs for s in stages])
    assert bb.dtype == np.int32 and np.array_equal(bb, g["bboxs"])
    bf = np.concatenate([tiling.bboxs_to_feat(s.bboxs, (2160, 3840), shape) for s in stages])
    assert bf.dtype == np.float32 and np.array_equal(bf, g["bboxs_feat"])     # inexact float32 factors replayed exactly


def test_bbox_factor_is_inexact_like_the_reference():
    """2160 -> 448 gives y2 = 447.99997, not 448."""
    bf = tiling.bboxs_to_feat(np.array([[2880, 1620, 3840, 2160]], np.int32), (2160, 3840), (448, 448))
    assert bf[0, 4] != np.float32(448.0) and abs(float(bf[0, 4]) - 448.0) < 1e-4
    t = O.bboxs_to_feat(torch.tensor([[2880, 1620, 3840, 2160]]).int(), (2160, 3840), (448, 448)).numpy()
    assert np.array_equal(bf, t)


def test_random_stream_consumption_matches_reference_order():
    tc = tiling.prepare_tile_cfg((448, 448), (2160, 3840), (4, 4))
    random.seed(7)
    st = tiling.random_stage(tc, 4)
    random.seed(7)
    hs = [random.randint(0, 2160 - 540 - 1) for _ in range(4)]
    w0 = random.randint(0, 3840 - 960 - 1)
    assert st.bboxs[:, 1].tolist() == hs and set(st.bboxs[:, 0].tolist()) == {w0}
    # r-count is floor(N / process_num) * process_num (patchrefiner.py:389)
    random.seed(1)
    n = sum(s.bboxs.shape[0] for s in tiling.schedule(tc, (448, 448), "r10", 4) if s.kind == "random")
    assert n == 8


def test_tile_cfg_and_edge_cases():
    tc = tiling.prepare_tile_cfg((448, 448), (2160, 3840), (4, 4))
    ref = O.prepare_tile_cfg((448, 448), (2160, 3840), (4, 4))
    assert tc == ref
    assert tiling.resizer_size((448, 448)) == (448, 448) and tiling.resizer_size((384, 512)) == (378, 518)
    with pytest.raises(NotImplementedError):
        tiling.parse_cai_mode("p16")
    with pytest.raises(AssertionError):
        tiling.regular_stage(tc, (448, 448), (-1, 0), (0, 0), False)
    one = tiling.prepare_tile_cfg((448, 448), (1080, 1920), (1, 1))
    random.seed(0)
    st = tiling.schedule(one, (448, 448), "m2", 4)
    assert [s.bboxs.shape[0] for s in st] == [1, 0, 0, 0]            # shifted grids are empty for a 1x1 split


def test_shard_partition():
    for P, G in [(81, 8), (49, 4), (16, 2), (5, 8)]:
        owns = [tiling.shard_patches(P, r, G) for r in range(G)]
        assert np.array_equal(np.sum(owns, axis=0), np.ones(P, np.uint8))
        assert max(int(o.sum()) for o in owns) - min(int(o.sum()) for o in owns) <= 1


def test_shard_partition_batch_of_frames():
    """A batch of frames is split into contiguous blocks: same balance, and a rank touches the fewest frames (its own frame when
    the batch holds one frame per rank), so only those frames have to be uploaded to it."""
    for F, P, G in [(8, 81, 8), (16, 353, 8), (2, 9, 2), (3, 49, 4), (2, 5, 8)]:
        owns = [tiling.shard_patches(F * P, r, G, frames=F) for r in range(G)]
        assert np.array_equal(np.sum(owns, axis=0), np.ones(F * P, np.uint8))
        assert max(int(o.sum()) for o in owns) - min(int(o.sum()) for o in owns) <= 1
        for r, o in enumerate(owns):
            idx = np.nonzero(o)[0]
            if len(idx):
                assert np.array_equal(idx, np.arange(idx[0], idx[-1] + 1))              # contiguous
                assert len(np.unique(idx // P)) <= -(-len(idx) // P) + 1
            if F == G:
                assert np.array_equal(np.unique(idx // P), [r])


@pytest.mark.parametrize("size,border", [((448, 448), 0.15), ((224, 224), 0.15), ((540, 960), 0.15), ((384, 512), 0.1)])
def test_masks_bit_identical_to_oracle(size, border):
    assert np.array_equal(masks.generatemask(size, border), O.generatemask(size, border))
    assert masks.generatemask(size, border) is masks.generatemask(size, border)        # cached
    assert np.array_equal(masks.random_patch_mask(size, border), O.generatemask(size, border) + 1e-3)


def test_registry_builds_by_type_name():
    cfg = O.make_config("vits", (224, 224), (432, 768), (2, 2))
    m = build_model(dict(type="PatchRefiner", config=cfg))
    assert type(m).__name__ == "PatchRefiner" and MODELS.get("PatchRefiner") is type(m)
    assert m.tile_cfg["patch_raw_shape"] == (216, 384) and m.tile_cfg["patch_reensemble_shape"] == (448, 448)
    with pytest.raises(KeyError):
        build_model(dict(type="Nope"))
    bad = O.make_config("vits", (224, 224), (432, 768), (2, 2))
    bad["coarse_branch"]["type"] = "ZoeDepth"
    with pytest.raises(NotImplementedError):
        build_model(dict(type="PatchRefiner", config=bad))


def test_model_refuses_to_run_without_cuda():
    cfg = O.make_config("vits", (224, 224), (432, 768), (2, 2))
    m = build_model(dict(type="PatchRefiner", config=cfg))
    lr, hr = O.synthetic_frame(cfg, 1)
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            m(mode="infer", image_lr=lr, image_hr=hr, cai_mode="m1")
    with pytest.raises(NotImplementedError):
        m(mode="train")


def test_c_abi_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "prv2_b200.h")).read()
    declared = set(re.findall(r"\b(prv2_[a-z0-9_]+)\s*\(", header))
    declared -= {"prv2_stream_t"}
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    lib = _lib.load()                                   # loads; no compute call without a GPU
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.prv2_version() == _lib.ABI_VERSION == int(re.search(r"#define PRV2_ABI_VERSION (\d+)", header).group(1))
    # the library is stamped with the digest of the sources it was compiled from; a stale .so is refused at load time
    from patchrefinerv2_b200 import build as B
    assert lib.prv2_build_digest().decode() == B.source_digest() == B.built_digest()


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "patchrefinerv2_b200")
    for fn in os.listdir(pkg):
        if fn.endswith(".py"):
            src = open(os.path.join(pkg, fn)).read()
            assert "oracle" not in src.replace("the oracle", "").replace("oracle/", ""), fn


def test_schedule_at_config5_geometry_matches_oracle_draws():
    """BASELINE config 5 (4320x7680, 8x8, r128): bboxes and roi coordinates of the whole schedule equal the oracle's
    (= the reference's) regular grids and random draws, stage by stage, without running any network."""
    shape, raw, split, pn = (448, 448), (4320, 7680), (8, 8), 4
    tc = tiling.prepare_tile_cfg(shape, raw, split)
    assert tc == O.prepare_tile_cfg(shape, raw, split) and tuple(tc["patch_reensemble_shape"]) == (3584, 3584)
    rh, rw = tc["patch_raw_shape"]
    random.seed(3)
    stages = tiling.schedule(tc, shape, "r128", pn)
    assert [s.bboxs.shape[0] for s in stages if s.kind == "regular"] == [64, 56, 56, 49]
    assert sum(s.bboxs.shape[0] for s in stages if s.kind == "random") == 128
    random.seed(3)
    want = []
    for off in ([0, 0], [0, rw // 2], [rh // 2, 0], [rh // 2, rw // 2]):
        hs, ws = O.regular_bboxes(tc, off)
        want.append(O.make_bboxs(hs, ws, rh, rw))
    for _ in range(128 // pn):                                   # baseline_pretrain.py:160-161: pn rows, then ONE column
        hs = [random.randint(0, raw[0] - rh - 1) for _ in range(pn)]
        ws = [random.randint(0, raw[1] - rw - 1)]
        want.append(O.make_bboxs(hs, ws, rh, rw))
    want = torch.cat(want)
    got = np.concatenate([s.bboxs for s in stages])
    assert np.array_equal(got, want.numpy())
    bf = np.concatenate([tiling.bboxs_to_feat(s.bboxs, raw, shape) for s in stages])
    assert np.array_equal(bf[:, 1:], O.bboxs_to_feat(want, raw, shape).numpy()[:, 1:])


def test_tiling_matches_oracle_on_random_geometries():
    """Property check over ragged geometries (frame sizes not divisible by the split, odd patch sizes, every CAI mode):
    tile config, every stage's bboxes and the float32 roi coordinates equal the oracle's restatement of the reference."""
    rng = np.random.default_rng(11)
    for case in range(25):
        sh, sw = int(rng.integers(1, 5)), int(rng.integers(1, 5))
        ph, pw = int(rng.integers(2, 9)) * 14, int(rng.integers(2, 9)) * 14
        H, W = int(rng.integers(sh * 24, sh * 200)), int(rng.integers(sw * 24, sw * 200))
        mode = ["m1", "m2", f"r{int(rng.integers(1, 5)) * 2}"][case % 3]
        pn = 2
        tc = tiling.prepare_tile_cfg((ph, pw), (H, W), (sh, sw))
        assert tc == O.prepare_tile_cfg((ph, pw), (H, W), (sh, sw)), (H, W, sh, sw)
        rh, rw = tc["patch_raw_shape"]
        if H - rh - 1 < 0 or W - rw - 1 < 0:
            continue                                              # random.randint would raise in the reference too (1x1 split)
        random.seed(case)
        stages = tiling.schedule(tc, (ph, pw), mode, pn)
        random.seed(case)
        want = []
        offs = [[0, 0]] + ([[0, rw // 2], [rh // 2, 0], [rh // 2, rw // 2]] if mode != "m1" else [])
        for off in offs:
            hs, ws = O.regular_bboxes(tc, off)
            if hs and ws:
                want.append(O.make_bboxs(hs, ws, rh, rw))
        if mode[0] == "r":
            for _ in range(int(mode[1:]) // pn):
                hs = [random.randint(0, H - rh - 1) for _ in range(pn)]
                ws = [random.randint(0, W - rw - 1)]
                want.append(O.make_bboxs(hs, ws, rh, rw))
        want = torch.cat(want)
        got = np.concatenate([s.bboxs for s in stages if s.bboxs.shape[0]])
        assert np.array_equal(got, want.numpy()), (case, H, W, sh, sw, mode)
        bf = np.concatenate([tiling.bboxs_to_feat(s.bboxs, (H, W), (ph, pw)) for s in stages if s.bboxs.shape[0]])
        assert np.array_equal(bf[:, 1:], O.bboxs_to_feat(want, (H, W), (ph, pw)).numpy()[:, 1:])
        assert (got[:, 2] <= W).all() and (got[:, 3] <= H).all() and (got[:, :2] >= 0).all()


def test_integration_doc_lists_every_c_abi_symbol():
    """INTEGRATION.md section 2 is the binding table a maintainer reads: every entry point the header declares must appear in it."""
    hdr = open(os.path.join(ROOT, "include", "prv2_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    syms = sorted(set(re.findall(r"\b(prv2_[a-z0-9_]+)\s*\(", hdr)))
    assert len(syms) >= 25
    # the doc groups some families as `prv2_blend_partial_{canvas,raw}`: expand those
    grouped = set()
    for stem, alts in re.findall(r"`(prv2_[a-z0-9_]*)\{([a-z0-9_,]+)\}`", doc):
        grouped.update(stem + a for a in alts.split(","))
    missing = [s for s in syms if s not in doc and s not in grouped]
    assert not missing, missing


def test_roi_regime_check_accepts_baseline_geometries_and_refuses_a_two_sample_grid():
    """roi_align's adaptive sampling grid must be 1x1 (the regime the gather kernels implement): true for every BASELINE geometry,
    refused loudly otherwise."""
    import random
    import numpy as np
    import pytest
    from patchrefinerv2_b200 import tiling
    for raw, split, mode in (((1080, 1920), (2, 2), "r8"), ((2160, 3840), (4, 4), "r8"), ((4320, 7680), (8, 8), "r8"), ((432, 768), (1, 1), "m1")):
        tc = tiling.prepare_tile_cfg((448, 448), raw, split)
        bb = np.concatenate([s.bboxs for s in tiling.schedule(tc, (448, 448), mode, 4, random.Random(0))])
        rois = tiling.bboxs_to_feat(bb, raw, (448, 448))[:, 1:]
        tiling.check_roi_regime(rois, [(32, 32, 32 / 448), (256, 256, 256 / 448), (448, 448, 1.0)])
    with pytest.raises(NotImplementedError):
        tiling.check_roi_regime(np.array([[0, 0, 449, 448]], np.float32), [(448, 448, 1.0)])


def test_unloaded_weights_are_reported_loudly():
    """A model whose checkpoint never populated some tensors must say so when its engine is built (strict=False loading)."""
    import warnings
    from oracle import pr_oracle as O
    from patchrefinerv2_b200 import build_model
    cfg = O.make_config("vits", (224, 224), (432, 768), (2, 2))
    m = build_model(dict(type="PatchRefiner", config=cfg))
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m._warn_unloaded()
    assert len(w) == 1 and "never loaded" in str(w[0].message)
    sd = O.init_patchrefiner_state_dict(cfg, 0)
    m.load_dict({k: v for k, v in sd.items() if not k.startswith("refiner_fusion_model.")})     # a checkpoint without the fusion model
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m._warn_unloaded()
    assert len(w) == 1 and "refiner_fusion_model." in str(w[0].message)
    m.load_dict(sd)
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        m._warn_unloaded()
    assert not w


def test_checkpoint_keys_of_the_config_follow_the_reference(tmp_path):
    """patchrefiner.py:94-98, 118-121, 129-147: branch files are strict over their branch (a wrong-keyed file raises), `pretrained` is
    strict=False and skips coarse_branch.* unless `load_whole`; state_dict() honours prefix / destination like nn.Module's."""
    import torch
    from patchrefinerv2_b200.model import PatchRefiner
    cfg = O.make_config("vits", (224, 224), (432, 768), (2, 2))
    sd = O.init_patchrefiner_state_dict(cfg, 0)
    coarse = {k[len("coarse_branch."):]: v for k, v in sd.items() if k.startswith("coarse_branch.")}
    fine = {k[len("refiner_fine_branch."):]: v for k, v in sd.items() if k.startswith("refiner_fine_branch.")}
    cp, fp, wp = str(tmp_path / "c.pth"), str(tmp_path / "f.pth"), str(tmp_path / "w.pth")
    torch.save(coarse, cp); torch.save(fine, fp)
    whole = {k: v + 1.0 for k, v in sd.items()}
    torch.save({"model_state_dict": whole}, wp)
    c2 = dict(cfg, coarse_branch=dict(cfg["coarse_branch"], pretrained=cp), refiner=dict(cfg["refiner"], fine_branch=dict(cfg["refiner"]["fine_branch"], pretrained=fp)))
    m = PatchRefiner(c2)
    got = m.state_dict()
    k_c, k_f, k_u = "coarse_branch.pretrained.cls_token", "refiner_fine_branch.pretrained.cls_token", "refiner_fusion_model.final_conv.weight"
    assert torch.equal(got[k_c], sd[k_c]) and torch.equal(got[k_f], sd[k_f]) and not got[k_u].any()
    # `pretrained` without load_whole: everything but the coarse branch
    m = PatchRefiner(dict(c2, pretrained=wp, load_whole=False))
    got = m.state_dict()
    assert torch.equal(got[k_c], sd[k_c]) and torch.equal(got[k_f], whole[k_f]) and torch.equal(got[k_u], whole[k_u])
    m = PatchRefiner(dict(c2, pretrained=wp, load_whole=True))
    assert torch.equal(m.state_dict()[k_c], whole[k_c])
    # a wrong-keyed branch file raises instead of leaving the branch at its initial values
    bad = str(tmp_path / "bad.pth")
    torch.save({"module." + k: v for k, v in coarse.items()}, bad)
    with pytest.raises(RuntimeError, match="coarse_branch"):
        PatchRefiner(dict(c2, coarse_branch=dict(cfg["coarse_branch"], pretrained=bad)))
    # nn.Module.state_dict signature
    dest = {}
    out = m.state_dict(destination=dest, prefix="module.")
    assert out is dest and set(dest) == {"module." + k for k in sd}
