"""Pins the restated oracle against the reference's own code, imported read-only from
/root/reference through oracle/ref_shim.py.  Skipped where the reference tree is absent (GPU box)."""
import os
import random
import tempfile

import pytest
import torch

from oracle import pr_oracle as O
from oracle import ref_shim

pytestmark = pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")


@pytest.fixture(scope="module")
def pair():
    cfg = O.make_config("vits", (224, 224), (432, 768), (2, 2))
    sd = O.init_patchrefiner_state_dict(cfg, 0)
    d = tempfile.mkdtemp()
    cp, fp = os.path.join(d, "c.pth"), os.path.join(d, "f.pth")
    torch.save({k[len("coarse_branch."):]: v for k, v in sd.items() if k.startswith("coarse_branch.")}, cp)
    torch.save({k[len("refiner_fine_branch."):]: v for k, v in sd.items() if k.startswith("refiner_fine_branch.")}, fp)
    cwd = os.getcwd()
    ref = ref_shim.build_reference_patchrefiner(cfg, cp, fp)
    res = ref.load_state_dict(sd, strict=False)
    assert not res.missing_keys and not res.unexpected_keys
    yield cfg, sd, ref
    os.chdir(cwd)


@pytest.mark.parametrize("mode", ["m1", "m2", "r4"])
def test_oracle_is_bit_identical_to_reference(pair, mode):
    cfg, sd, ref = pair
    lr, hr = O.synthetic_frame(cfg, 1)
    random.seed(1)
    with torch.no_grad():
        dref, log = ref(mode="infer", image_lr=lr, image_hr=hr, cai_mode=mode, process_num=2, tile_cfg=None)
    random.seed(1)
    dor, coarse, _ = O.PatchRefinerOracle(cfg, sd).infer(lr, hr, None, mode, 2)
    assert dref.shape == dor.shape
    assert torch.equal(dref, dor)
    assert torch.equal(log["coarse_prediction"], coarse)


@pytest.mark.parametrize("target", ["offset_fine", "direct"])
def test_oracle_strategy_refiner_targets_bit_identical_to_reference(pair, target):
    """strategy_refiner_target other than the shipped 'offset_coarse' (patchrefiner.py:270-283): the refiner depth as update base,
    and 'direct' (no base, sigmoid * max_depth)."""
    cfg, sd, ref = pair
    lr, hr = O.synthetic_frame(cfg, 1)
    old = ref.strategy_refiner_target
    try:
        ref.strategy_refiner_target = target
        random.seed(1)
        with torch.no_grad():
            dref, _ = ref(mode="infer", image_lr=lr, image_hr=hr, cai_mode="m1", process_num=2, tile_cfg=None)
    finally:
        ref.strategy_refiner_target = old
    random.seed(1)
    dor, _, _ = O.PatchRefinerOracle(dict(cfg, strategy_refiner_target=target), sd).infer(lr, hr, None, "m1", 2)
    assert torch.equal(dref, dor)


def test_state_dict_keys_match_reference(pair):
    cfg, sd, ref = pair
    assert set(ref.state_dict().keys()) == set(sd.keys())
    from patchrefinerv2_b200.model import PatchRefiner
    mine = PatchRefiner(cfg)
    assert set(mine.state_dict().keys()) == set(sd.keys())
    for k, v in ref.state_dict().items():
        assert tuple(mine.state_dict()[k].shape) == tuple(v.shape), k
    assert set(mine.get_save_dict().keys()) == set(ref.get_save_dict().keys())


def test_zoe_bins_head_oracle_is_bit_identical_to_reference():
    """ZoeDepth metric-bins head, oracle only so far: the reference's ZoeDepth with a dummy core and
    ``hack_feature`` inputs (zoedepth_v1.py:164-171) against oracle.zoe_bins_head -- depth, probabilities, bin centres."""
    ref_shim.install()
    from zoedepth.models.zoedepth.zoedepth_v1 import ZoeDepth

    class DummyCore(torch.nn.Module):
        output_channels = [256] * 5                      # DPT_BEiT_L_384 (base_models/midas.py:377-380)

        def freeze_encoder(self, *a, **k):
            pass

    c = O.ZOE_HEAD_CFG
    m = ZoeDepth(DummyCore(), n_bins=c["n_bins"], bin_centers_type=c["bin_centers_type"], bin_embedding_dim=c["bin_embedding_dim"], min_depth=1e-3,
                 max_depth=80, n_attractors=list(c["n_attractors"]), attractor_alpha=c["attractor_alpha"], attractor_gamma=c["attractor_gamma"],
                 attractor_kind=c["attractor_kind"], attractor_type=c["attractor_type"], min_temp=c["min_temp"], max_temp=c["max_temp"]).eval()
    sd = O.init_zoe_head_state_dict([256] * 5, 7)
    res = m.load_state_dict(sd, strict=False)
    assert not res.unexpected_keys and all("log_binomial_transform" in k for k in res.missing_keys)      # only the two index buffers
    g = torch.Generator().manual_seed(2)
    sizes = [(12, 16), (12, 16), (24, 32), (48, 64), (96, 128)]                                          # btlnck, then coarse -> fine
    btl = torch.randn(2, 256, *sizes[0], generator=g)
    xb = [torch.randn(2, 256, *s, generator=g) for s in sizes[1:]]
    outc = torch.relu(torch.randn(2, 32, 192, 256, generator=g))
    rel = torch.rand(2, 192, 256, generator=g) * 5
    with torch.no_grad():
        r = m(None, hack_feature=[rel, [btl] + xb + [outc]], return_final_centers=True, return_probs=True)
        tr = {}
        d, feats = O.zoe_bins_head(sd, "", rel, btl, xb, outc, c, tr)
    assert torch.equal(r["metric_depth"], d) and torch.equal(r["probs"], tr["probs"]) and torch.equal(r["bin_centers"], tr["bin_centers"])
    assert set(feats) == {"x_d0", "x_blocks_feat_0", "x_blocks_feat_1", "x_blocks_feat_2", "x_blocks_feat_3", "midas_final_feat"}
    assert torch.equal(r["temp_features"]["x_d0"], feats["x_d0"])
