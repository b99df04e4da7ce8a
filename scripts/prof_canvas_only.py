import os, sys, random
import numpy as np, torch
sys.path.insert(0, os.getcwd())
from patchrefinerv2_b200 import masks, ops, tiling
dev = torch.device("cuda")
tc = tiling.prepare_tile_cfg((448, 448), (2160, 3840), (4, 4))
st = tiling.schedule(tc, (448, 448), "m2", 4, random.Random(1))
bb = np.concatenate([s.bboxs for s in st])
stages, first = [], 0
for s_ in st:
    stages.append((s_.off_process[0], s_.off_process[1], s_.grid[0], s_.grid[1], first)); first += s_.bboxs.shape[0]
preds = torch.rand(bb.shape[0], 448, 448, device=dev) * 10
mask = torch.from_numpy(masks.generatemask((448, 448), 0.15).copy()).to(dev)
flush = torch.empty(64 * 1024 * 1024, device=dev)
for _ in range(6):
    flush.fill_(1.0)
    ops.blend_canvas(preds[:first], mask, stages, 1792, 1792)
torch.cuda.synchronize()
