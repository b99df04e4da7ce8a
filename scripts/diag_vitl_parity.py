"""Where does the fp32-mode error at ViT-L come from?  Per-intermediate error of the B200 model (fp32 and bf16 modes) against the
CPU oracle, next to the error of the SAME oracle code run as eager PyTorch on the GPU (strict fp32: the reference's own
backend-to-backend noise floor).  Run on the GPU box: python scripts/diag_vitl_parity.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
import test_vitl_gpu as T
from oracle import pr_oracle as O

c = T.vitl_case.__wrapped__()
t = c["trace"]
DEV = "cuda:0"


def rel_max(a, b):
    return ((a - b).abs().max() / b.abs().max().clamp_min(1e-6)).item()


def stats(name, got, want):
    err = (got - want).abs()
    rel = err / want.abs().clamp_min(1e-3)
    k = rel.flatten().argmax().item()
    print(f"  {name:14s} max_abs {err.max():.3e}  rel_to_max {rel_max(got, want):.3e}  px_rel max {rel.max():.3e} (want {want.flatten()[k]:.4f} got {got.flatten()[k]:.4f})  "
          f"p99 {rel.flatten().kthvalue(int(rel.numel() * 0.99)).values:.3e}  p99.99 {rel.flatten().kthvalue(int(rel.numel() * 0.9999)).values:.3e}  mean {rel.mean():.3e}  n(rel>1e-3) {(rel > 1e-3).sum().item()}/{rel.numel()}")


# (1) the oracle itself on the GPU (eager fp32)
sd_g = {k: v.to(DEV) for k, v in c["sd"].items()}
og = O.PatchRefinerOracle(c["cfg"], sd_g)
og.trace = {}
with torch.no_grad():
    feats, coarse = og.coarse_forward(c["lr"].to(DEV))
    tt = {"coarse_prediction": coarse, "coarse_features": feats}
    pg = og._predict(c["hr"][0].to(DEV), c["bb"], og.tile_cfg, tt, 1, None)
tg = og.trace["first_patch"]
print("oracle on GPU (eager fp32, TF32 off) vs oracle on CPU:")
stats("coarse", coarse.cpu(), c["coarse"])
for k in ("tokens0", "block0"):
    stats(k, tg[k].cpu(), t[k])
for i, (a, b) in enumerate(zip(tg["taps"], t["taps"])):
    stats(f"tap{i}", a.cpu(), b)
for i, (a, b) in enumerate(zip(tg["fine_feats"], t["fine_feats"])):
    stats(f"fine_feat{i}", a.cpu(), b)
stats("fine_depth", tg["fine_depth"].cpu(), t["fine_depth"])
for i, (a, b) in enumerate(zip(tg["fusion_enc"], t["fusion_enc"])):
    stats(f"fusion_enc{i}", a.cpu(), b)
stats("fusion_dec", tg["fusion_dec"].cpu(), t["fusion_dec"])
stats("pred", pg.cpu()[:, 0], c["preds"][:, 0])
del og, sd_g
torch.cuda.empty_cache()

from patchrefinerv2_b200 import build_model
for prec in ("fp32", "bf16"):
    m = build_model(dict(type="PatchRefiner", config=c["cfg"], precision=prec, patch_batch=2, output_device="cuda"))
    m.load_dict(c["sd"])
    m = m.cuda().eval()
    tr = {}
    got, coarse = m.predict_patches(c["lr"].to(DEV), c["hr"].to(DEV), c["bb"], trace=tr)
    print(f"B200 model, {prec} mode vs oracle on CPU:")
    stats("coarse", coarse.cpu(), c["coarse"])
    stats("roi_depth", tr["roi_depth"].cpu(), c["rec"]["roi_first"]["depth"])
    for k in ("tokens0", "block0"):
        stats(k, tr[k].cpu()[:1], t[k])
    for i, (a, b) in enumerate(zip(tr["taps"], t["taps"])):
        stats(f"tap{i}", a.cpu()[:1], b)
    for i, (a, b) in enumerate(zip(tr["roi_feats"], c["rec"]["roi_first"]["feats"])):
        stats(f"roi_feat{i}", a.cpu(), b)
    for i, (a, b) in enumerate(zip(tr["fine_feats"], t["fine_feats"])):
        stats(f"fine_feat{i}", a.cpu()[:1], b)
    stats("fine_depth", tr["fine_depth"].cpu()[:1], t["fine_depth"])
    for i, (a, b) in enumerate(zip(tr["fusion_enc"], t["fusion_enc"])):
        stats(f"fusion_enc{i}", a.cpu()[:1], b)
    stats("fusion_dec", tr["fusion_dec"].cpu()[:1], t["fusion_dec"])
    stats("pred", got.cpu(), c["preds"][:, 0])
    del m
    torch.cuda.empty_cache()
