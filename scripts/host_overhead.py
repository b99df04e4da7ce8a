"""How much of a frame is host enqueue time?  Times (a) the host-side duration of model.forward without synchronising and
(b) the device time of the same frames, for the patch batches of N = 1 (27) and N = 8 (11 per rank).
    python scripts/host_overhead.py"""
import os, random, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from patchrefinerv2_b200 import build_model

enc, pshape, raw, split, cai_mode, pn = bench.WORKLOADS["dav2_vitl_2160x3840_4x4_r32"]
cfg = bench.make_config(enc, pshape, raw, split)
dev = torch.device("cuda:0")
hr = bench.synthetic_frame(raw, 1).to(dev)
for pb in (27, 11):
    m = build_model(dict(type="PatchRefiner", config=cfg, precision="bf16", patch_batch=pb, output_device="cuda"))
    m.load_dict(bench.random_state_dict({k: tuple(v.shape) for k, v in m.state_dict().items()}, 0))
    m = m.cuda().eval()
    lr = m.resizer(hr)
    for _ in range(3):
        random.seed(1); m(mode="infer", image_lr=lr, image_hr=hr, cai_mode=cai_mode, process_num=pn)
    torch.cuda.synchronize()
    host, devt = [], []
    for _ in range(5):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        random.seed(1)
        t0 = time.perf_counter(); e0.record()
        m(mode="infer", image_lr=lr, image_hr=hr, cai_mode=cai_mode, process_num=pn)
        e1.record(); t1 = time.perf_counter()
        torch.cuda.synchronize()
        host.append((t1 - t0) * 1e3); devt.append(e0.elapsed_time(e1))
    print(f"patch_batch {pb}: host enqueue {sorted(host)[2]:.1f} ms/frame, device {sorted(devt)[2]:.1f} ms/frame  (81 patches on one GPU)")
    del m
    torch.cuda.empty_cache()
