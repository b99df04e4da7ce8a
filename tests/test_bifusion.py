"""BiDirectionalFusion (V2 family fusion model, SURVEY.md row a9'): oracle vs the reference-made goldens and vs the
reference module itself on CPU; the sm_100a implementation vs both on the GPU."""
import os

import numpy as np
import pytest
import torch

from oracle import pr_oracle as O
from oracle import ref_shim
from oracle.make_golden import BIFUSION, BIFUSION_SIZES_C, BIFUSION_SIZES_F, BIFUSION_TYPES, sd_digest


ABLATIONS = [("only-gate", False), ("coarse-gated", True), ("self-agg", True)]     # (coarse2fine_type, heavy)


def _case(t, seed_w=5, seed_x=3, B=2):
    sd = O.init_bidirectional_fusion_state_dict(seed=seed_w, coarse2fine_type=t, **BIFUSION)
    c, f, p1, p2 = O.synthetic_fusion_inputs(BIFUSION["coarse_chl"], BIFUSION["fine_chl"], BIFUSION_SIZES_C, BIFUSION_SIZES_F, B, seed_x)
    return sd, c, f, p1, p2


@pytest.mark.parametrize("t", BIFUSION_TYPES)
def test_oracle_matches_reference_golden(golden_dir, t):
    g = np.load(os.path.join(golden_dir, f"bifusion_{t}.npz"))
    sd, c, f, p1, p2 = _case(t)
    assert str(g["sd_sha"]) == sd_digest(sd) and str(g["pred1_sha"]) == O.sha256_f32(p1.numpy()), "generators drifted from the golden run"
    assert sorted(sd.keys()) == list(g["keys"])                                   # the reference module's own state-dict keys
    with torch.no_grad():
        depth = O.bidirectional_fusion(sd, "", c, f, p1, p2, p1, t)
        off = O.bidirectional_fusion(sd, "", c, f, p1, p2, None, t)
    # the goldens were made on this CPU build (bit-identical there); 1e-3 relative is the north-star budget elsewhere
    np.testing.assert_allclose(depth.numpy(), g["depth"], rtol=1e-3, atol=1e-3)
    np.testing.assert_allclose(off.numpy(), g["offset"], rtol=1e-3, atol=1e-3)
    assert depth.min() >= 0                                                       # clamp(min=0) (:441)


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("t", BIFUSION_TYPES)
def test_oracle_is_bit_identical_to_reference_module(t):
    ref_shim.install()
    from estimator.models.blocks.bi_directional_fusion_model import BiDirectionalFusion
    m = BiDirectionalFusion(encoder_name="x", coarse2fine_type=t, **{k: list(v) for k, v in BIFUSION.items()}).eval()
    sd, c, f, p1, p2 = _case(t, seed_w=9, seed_x=4)
    m.load_state_dict(sd, strict=True)
    with torch.no_grad():
        r = m(c_feat=[x.clone() for x in c], f_feat=[x.clone() for x in f], pred1=p1, pred2=p2, update_base=p1)
        o = O.bidirectional_fusion(sd, "", c, f, p1, p2, p1, t)
    assert torch.equal(r, o)


def test_product_state_dict_and_registry_surface():
    from patchrefinerv2_b200 import MODELS, build_model
    from patchrefinerv2_b200.bifusion import BiDirectionalFusion
    for t in BIFUSION_TYPES:
        m = build_model(dict(type="BiDirectionalFusion", encoder_name="x", coarse2fine_type=t, **{k: list(v) for k, v in BIFUSION.items()}))
        assert isinstance(m, BiDirectionalFusion) and MODELS.get("BiDirectionalFusion") is BiDirectionalFusion
        sd = O.init_bidirectional_fusion_state_dict(seed=5, coarse2fine_type=t, **BIFUSION)
        assert set(m.state_dict().keys()) == set(sd.keys())
        for k, v in m.state_dict().items():
            assert tuple(v.shape) == tuple(sd[k].shape), k
        res = m.load_state_dict(sd, strict=True)
        assert not res.missing_keys and not res.unexpected_keys
        with pytest.raises(RuntimeError):
            m.load_state_dict({"nope": torch.zeros(1)}, strict=True)
    for t, heavy in ABLATIONS:                                                    # the four ablation configs' fusion models
        m = build_model(dict(type="BiDirectionalFusionHeavy" if heavy else "BiDirectionalFusion", encoder_name="x", coarse2fine_type=t,
                             **{k: list(v) for k, v in BIFUSION.items()}))
        sd = O.init_bidirectional_fusion_state_dict(seed=5, coarse2fine_type=t, heavy=heavy, **BIFUSION)
        assert set(m.state_dict().keys()) == set(sd.keys())
        assert all(tuple(v.shape) == tuple(sd[k].shape) for k, v in m.state_dict().items())
    with pytest.raises(NotImplementedError):
        build_model(dict(type="BiDirectionalFusion", coarse2fine_type="no-such-type"))
    with pytest.raises(NotImplementedError):
        build_model(dict(type="BiDirectionalFusion", glb_att=True))
    m = build_model(dict(type="BiDirectionalFusion", coarse2fine_type="coarse-gated"))
    with pytest.raises(RuntimeError):                                             # no CPU path: fail loudly
        m(c_feat=[torch.zeros(1, 32, 8, 8)] * 6, f_feat=[torch.zeros(1, 32, 8, 8)] * 6, pred1=torch.zeros(1, 1, 8, 8))


@pytest.fixture(scope="module")
def dev():
    return torch.device("cuda:0")


def _rel(a, b):
    return float((a - b).abs().max() / b.abs().max())


@pytest.mark.gpu
@pytest.mark.parametrize("t", BIFUSION_TYPES)
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 6e-2)])
def test_b200_bifusion_matches_reference_golden_and_oracle(dev, golden_dir, t, precision, tol):
    """fp32 mode: within 1e-3 (relative to the offset's range) of the reference-made golden; bf16 mode: its own tolerance.
    The intermediate C2F depth and features are checked against the oracle trace as well."""
    from patchrefinerv2_b200 import build_model
    g = np.load(os.path.join(golden_dir, f"bifusion_{t}.npz"))
    sd, c, f, p1, p2 = _case(t)
    m = build_model(dict(type="BiDirectionalFusion", encoder_name="x", coarse2fine_type=t, precision=precision, **{k: list(v) for k, v in BIFUSION.items()}))
    m.load_state_dict(sd)
    m = m.cuda()
    tr = {}
    cd, fd = [x.to(dev) for x in c], [x.to(dev) for x in f]
    off = m(c_feat=cd, f_feat=fd, pred1=p1.to(dev), pred2=p2.to(dev), update_base=None, trace=tr).cpu()
    depth = m(c_feat=cd, f_feat=fd, pred1=p1.to(dev), pred2=p2.to(dev), update_base=p1.to(dev)).cpu()
    otr = {}
    with torch.no_grad():
        O.bidirectional_fusion(sd, "", c, f, p1, p2, None, t, otr)
    assert _rel(tr["c2f_depth"].cpu(), otr["c2f_depth"]) < tol
    for a, b in zip(tr["c2f_feats"], otr["c2f_feats"]):
        assert a.shape == b.shape and _rel(a.cpu(), b) < tol
    ref_off, ref_depth = torch.from_numpy(g["offset"]), torch.from_numpy(g["depth"])
    assert off.shape == ref_off.shape
    assert _rel(off, ref_off) < tol
    assert float((depth - ref_depth).abs().max()) < tol * float(ref_off.abs().max()) and depth.min() >= 0


@pytest.mark.gpu
@pytest.mark.parametrize("t,heavy", ABLATIONS)
@pytest.mark.parametrize("precision,tol", [("fp32", 1e-3), ("bf16", 6e-2)])
def test_b200_ablation_variants_match_oracle(dev, t, heavy, precision, tol):
    """C2FNOENCModule ('only-gate': transposed-conv level 0, two gated units per level, no top-down path) and
    BiDirectionalFusionHeavy (conv-LN-conv-LN-conv-GELU encoder blocks, five-conv decoder blocks) on the kernels, against the
    oracle that tests/test_bifusion.py::test_oracle_covers_the_ablation_variants_bit_identically pins to the reference modules."""
    from patchrefinerv2_b200 import build_model
    sd = O.init_bidirectional_fusion_state_dict(seed=5, coarse2fine_type=t, heavy=heavy, **BIFUSION)
    c, f, p1, p2 = O.synthetic_fusion_inputs(BIFUSION["coarse_chl"], BIFUSION["fine_chl"], BIFUSION_SIZES_C, BIFUSION_SIZES_F, 2, 3)
    m = build_model(dict(type="BiDirectionalFusionHeavy" if heavy else "BiDirectionalFusion", encoder_name="x", coarse2fine_type=t, precision=precision,
                         **{k: list(v) for k, v in BIFUSION.items()}))
    m.load_state_dict(sd)
    m = m.cuda()
    tr, otr = {}, {}
    cd, fd = [x.to(dev) for x in c], [x.to(dev) for x in f]
    off = m(c_feat=cd, f_feat=fd, pred1=p1.to(dev), pred2=p2.to(dev), update_base=None, trace=tr).cpu()
    depth = m(c_feat=cd, f_feat=fd, pred1=p1.to(dev), pred2=p2.to(dev), update_base=p1.to(dev)).cpu()
    with torch.no_grad():
        ref_off = O.bidirectional_fusion(sd, "", c, f, p1, p2, None, t, otr, heavy=heavy)
        ref_depth = O.bidirectional_fusion(sd, "", c, f, p1, p2, p1, t, heavy=heavy)
    assert _rel(tr["c2f_depth"].cpu(), otr["c2f_depth"]) < tol
    for a, b in zip(tr["c2f_feats"], otr["c2f_feats"]):
        assert a.shape == b.shape and _rel(a.cpu(), b) < tol
    for a, b in zip(tr["fusion_enc"], otr["fusion_enc"]):
        assert _rel(a.cpu(), b) < tol * 3
    assert off.shape == ref_off.shape and _rel(off, ref_off) < tol * (3 if heavy else 1)
    assert float((depth - ref_depth).abs().max()) < tol * (3 if heavy else 1) * float(ref_off.abs().max()) and depth.min() >= 0


@pytest.mark.gpu
def test_b200_gate_and_ln_relu_epilogues_vs_torch(dev):
    """The two epilogues BiDirectionalFusion adds to prv2_umma_gemm, each against a plain PyTorch fp32 reference:
    LN(acc + bias) -> ReLU over a virtual concat, and res * sigmoid(acc) (+ res2, + ReLU copy)."""
    import torch.nn.functional as F
    from patchrefinerv2_b200 import _lib
    from patchrefinerv2_b200.nn import Act, GemmLayer, conv_segments
    torch.manual_seed(0)
    B, H, W, Fe = 2, 24, 20, 256
    a, c = torch.randn(B, Fe, H, W), torch.randn(B, Fe, H, W)
    w = torch.randn(Fe, 2 * Fe, 3, 3) / (2 * Fe * 9) ** 0.5
    bias, gamma, beta = torch.randn(Fe) * 0.5, 1 + 0.1 * torch.randn(Fe), 0.1 * torch.randn(Fe)
    y = F.conv2d(torch.cat([a, c], 1), w, bias, padding=1)
    u = y.mean(1, keepdim=True); s = (y - u).pow(2).mean(1, keepdim=True)
    want = F.relu(gamma[:, None, None] * ((y - u) / torch.sqrt(s + 1e-6)) + beta[:, None, None])
    for x3, tol in ((True, 2e-3), (False, 5e-2)):
        lay = GemmLayer(conv_segments(w, [Fe, Fe]), 2, Fe, x3, dev, epi=_lib.EPI_LN_GELU, act=_lib.ACT_RELU, bias=bias, gamma=gamma, beta=beta, eps=1e-6)
        out = Act.empty(B, H, W, Fe, x3, dev)
        lay([Act.from_nchw(a.to(dev), x3), Act.from_nchw(c.to(dev), x3)], out=out)
        assert float((out.to_nchw().cpu() - want).abs().max()) < tol * float(want.abs().max())
    w1 = torch.randn(Fe, Fe) / Fe ** 0.5
    fz, skip = torch.randn(B, Fe, H, W), torch.randn(B, Fe, H, W)
    gated = a * torch.sigmoid(F.conv2d(fz, w1[:, :, None, None])) + skip
    for x3, tol in ((True, 2e-3), (False, 3e-2)):
        lay = GemmLayer([(0, 0, 0, w1)], 1, Fe, x3, dev, act=_lib.ACT_SIGMOID_GATE)
        out, out_relu = Act.empty(B, H, W, Fe, x3, dev), Act.empty(B, H, W, Fe, x3, dev)
        lay([Act.from_nchw(fz.to(dev), x3)], out=out, relu_out=out_relu, res=Act.from_nchw(a.to(dev), x3), res2=Act.from_nchw(skip.to(dev), x3))
        assert float((out.to_nchw().cpu() - gated).abs().max()) < tol * float(gated.abs().max())
        assert float((out_relu.to_nchw().cpu() - F.relu(gated)).abs().max()) < tol * float(gated.abs().max())


@pytest.mark.skipif(not ref_shim.reference_available(), reason="/root/reference not present")
@pytest.mark.parametrize("t,heavy", [("only-gate", False), ("coarse-gated", True), ("self-agg", True)])
def test_oracle_covers_the_ablation_variants_bit_identically(t, heavy):
    """The four ablation configs' fusion models (C2FNOENCModule 'only-gate', BiDirectionalFusionHeavy): the oracle against the reference -- the product
    raises NotImplementedError for them -- pinned against the reference modules so the next step has a checker."""
    ref_shim.install()
    from estimator.models.blocks import bi_directional_fusion_model as ref
    cls = ref.BiDirectionalFusionHeavy if heavy else ref.BiDirectionalFusion
    m = cls(encoder_name="x", coarse2fine_type=t, **{k: list(v) for k, v in BIFUSION.items()}).eval()
    sd = O.init_bidirectional_fusion_state_dict(seed=5, coarse2fine_type=t, heavy=heavy, **BIFUSION)
    assert set(m.state_dict().keys()) == set(sd.keys())
    m.load_state_dict(sd, strict=True)
    c, f, p1, p2 = O.synthetic_fusion_inputs(BIFUSION["coarse_chl"], BIFUSION["fine_chl"], BIFUSION_SIZES_C, BIFUSION_SIZES_F, 2, 3)
    with torch.no_grad():
        r = m(c_feat=[x.clone() for x in c], f_feat=[x.clone() for x in f], pred1=p1, pred2=p2, update_base=p1)
        o = O.bidirectional_fusion(sd, "", c, f, p1, p2, p1, t, heavy=heavy)
    assert torch.equal(r, o)
