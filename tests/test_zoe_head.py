"""ZoeDepth metric-bins head on the B200 kernels: against the reference-made golden (tests/golden/zoe_head.npz,
written by the reference's own ZoeDepth with a dummy core and ``hack_feature`` inputs) and against the oracle, which
tests/test_oracle_vs_reference.py pins bit-identical to the reference.  fp32 mode: 1e-3; bf16 mode: its own tolerance."""
import os

import numpy as np
import pytest
import torch

from oracle import pr_oracle as O
from oracle.make_golden import zoe_head_inputs


def test_zoe_head_registry_and_state_dict_surface():
    from patchrefinerv2_b200 import MODELS, build_model
    from patchrefinerv2_b200.zoe import ZoeDepthBinsHead
    m = build_model(dict(type="ZoeDepthBinsHead", output_channels=[256] * 5))
    assert isinstance(m, ZoeDepthBinsHead) and MODELS.get("ZoeDepthBinsHead") is ZoeDepthBinsHead
    sd = O.init_zoe_head_state_dict([256] * 5, 7)
    assert set(m.state_dict()) == set(sd) and all(tuple(v.shape) == tuple(sd[k].shape) for k, v in m.state_dict().items())
    res = m.load_state_dict(sd, strict=True)
    assert not res.missing_keys and not res.unexpected_keys
    with pytest.raises(RuntimeError):                                  # no CPU path: fail loudly
        m(torch.zeros(1, 8, 8), torch.zeros(1, 256, 2, 2), [torch.zeros(1, 256, 2, 2)] * 4, torch.zeros(1, 32, 8, 8))


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 5e-2)])
def test_b200_zoe_head_matches_reference_golden_and_oracle(golden_dir, precision, tol):
    from patchrefinerv2_b200 import build_model
    dev = torch.device("cuda:0")
    g = np.load(os.path.join(golden_dir, "zoe_head.npz"))
    sd = O.init_zoe_head_state_dict([256] * 5, 7)
    rel, btl, xb, outc = zoe_head_inputs()
    m = build_model(dict(type="ZoeDepthBinsHead", output_channels=[256] * 5, precision=precision))
    m.load_state_dict(sd)
    tr, otr = {}, {}
    depth = m(rel.to(dev), btl.to(dev), [x.to(dev) for x in xb], outc.to(dev), trace=tr).cpu()
    with torch.no_grad():
        want, _ = O.zoe_bins_head(sd, "", rel, btl, xb, outc, O.ZOE_HEAD_CFG, otr)
    ref = torch.from_numpy(g["depth"])
    assert depth.shape == ref.shape == want.shape
    rel_err = ((depth - ref).abs() / ref.abs().clamp_min(1e-3)).max().item()
    assert rel_err < tol, rel_err                                     # per pixel, against what the reference itself wrote
    assert ((depth - want).abs() / want.abs().clamp_min(1e-3)).max().item() < tol
    # bin centres at the last decoder level ([B,h,w,K] here, [B,K,h,w] upsampled to the output size in the oracle trace)
    c = torch.nn.functional.interpolate(tr["bin_centers_last"].permute(0, 3, 1, 2).cpu(), otr["bin_centers"].shape[-2:], mode="bilinear", align_corners=True)
    assert ((c - otr["bin_centers"]).abs().max() / otr["bin_centers"].abs().max()).item() < tol
