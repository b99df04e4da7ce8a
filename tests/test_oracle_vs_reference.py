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


def test_state_dict_keys_match_reference(pair):
    cfg, sd, ref = pair
    assert set(ref.state_dict().keys()) == set(sd.keys())
    from patchrefinerv2_b200.model import PatchRefiner
    mine = PatchRefiner(cfg)
    assert set(mine.state_dict().keys()) == set(sd.keys())
    for k, v in ref.state_dict().items():
        assert tuple(mine.state_dict()[k].shape) == tuple(v.shape), k
    assert set(mine.get_save_dict().keys()) == set(ref.get_save_dict().keys())
