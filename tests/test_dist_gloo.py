"""World-size-2 gloo test (CPU) of the patch-sharding host logic: every rank
replays the same schedule (same `random` seed), owns patches i = rank (mod world), builds packed
partial canvases, ONE all_reduce(sum) combines them and the closed-form finalize reproduces the
reference's sequential blend (depth within 1e-3 relative, count map bit-exact because it is
recomputed locally in reference order).  The partial / finalize arithmetic below mirrors the CUDA
kernels (csrc/geometry.cu MODE 1 / MODE 2) in plain torch."""
import os
import random
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp
import torch.nn.functional as F

from oracle import pr_oracle as O
from patchrefinerv2_b200 import masks, tiling

SHAPE, RAW, SPLIT, MODE, PN = (224, 224), (432, 768), (2, 2), "r4", 2


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _schedule():
    tc = tiling.prepare_tile_cfg(SHAPE, RAW, SPLIT)
    random.seed(1)
    stages = tiling.schedule(tc, SHAPE, MODE, PN)
    return tc, stages


def _schedule_batch(frames):
    """Schedules of a batch of frames, drawn frame by frame from ONE `random` stream (as `frames` successive reference calls do)."""
    tc = tiling.prepare_tile_cfg(SHAPE, RAW, SPLIT)
    random.seed(1)
    return tc, [tiling.schedule(tc, SHAPE, MODE, PN) for _ in range(frames)]


def _partials(rank, world, stages=None, own=None):
    tc, st1 = _schedule()
    stages = st1 if stages is None else stages
    ph, pw = SHAPE
    rh, rw = tc["patch_raw_shape"]
    H, W = tc["image_raw_shape"]
    Hc, Wc = tc["patch_reensemble_shape"]
    bb = np.concatenate([s.bboxs for s in stages])
    own = tiling.shard_patches(bb.shape[0], rank, world) if own is None else own
    mask = torch.from_numpy(masks.generatemask(SHAPE, 0.15).copy())
    rmask = torch.from_numpy(masks.random_patch_mask((rh, rw), 0.15).copy())
    num_c, m1, num_r = torch.zeros(Hc, Wc), torch.zeros(Hc, Wc), torch.zeros(H, W)
    k = 0
    for s in stages:
        for i, b in enumerate(s.bboxs.tolist()):
            if own[k]:
                pred = O.fake_prediction(b, ph, pw)[0]            # only owned patches are ever evaluated
                if s.kind == "regular":
                    r, c = divmod(i, s.grid[1])
                    y0, x0 = s.off_process[0] + r * ph, s.off_process[1] + c * pw
                    if s.init:
                        m1[y0:y0 + ph, x0:x0 + pw] = pred
                    else:
                        num_c[y0:y0 + ph, x0:x0 + pw] += pred * mask
                else:
                    pr = F.interpolate(pred[None, None], (rh, rw))[0, 0]
                    num_r[b[1]:b[1] + rh, b[0]:b[0] + rw] += pr * rmask
            k += 1
    return torch.cat([num_c.flatten(), m1.flatten(), num_r.flatten()]), (tc, stages, mask, rmask)


def _finalize(packed, ctx):
    tc, stages, mask, rmask = ctx
    ph, pw = SHAPE
    rh, rw = tc["patch_raw_shape"]
    H, W = tc["image_raw_shape"]
    Hc, Wc = tc["patch_reensemble_shape"]
    num_c, m1, num_r = packed[:Hc * Wc].view(Hc, Wc), packed[Hc * Wc:2 * Hc * Wc].view(Hc, Wc), packed[2 * Hc * Wc:].view(H, W)
    cnt = torch.zeros(Hc, Wc)
    cnt0 = torch.zeros(Hc, Wc)
    for s in stages:                                            # count map: local, reference order
        if s.kind != "regular":
            continue
        for i in range(s.bboxs.shape[0]):
            r, c = divmod(i, s.grid[1])
            y0, x0 = s.off_process[0] + r * ph, s.off_process[1] + c * pw
            if s.init:
                cnt[y0:y0 + ph, x0:x0 + pw] = mask
                cnt0 = cnt.clone()
            else:
                sl = cnt[y0:y0 + ph, x0:x0 + pw]
                sl[mask > 0] = sl[mask > 0] + mask[mask > 0]
    avg = torch.where(cnt > cnt0, (m1 * cnt0 + num_c) / cnt, m1)
    a0 = F.interpolate(avg[None, None], (H, W))[0, 0]
    c0 = F.interpolate(cnt[None, None], (H, W), mode="bilinear", align_corners=True)[0, 0]
    cr = c0.clone()
    for s in stages:
        if s.kind == "random":
            for b in s.bboxs.tolist():
                cr[b[1]:b[1] + rh, b[0]:b[0] + rw] += rmask
    out = torch.where(cr > c0, (a0 * c0 + num_r) / cr, a0)
    return out, cr


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    packed, ctx = _partials(rank, world)
    dist.all_reduce(packed)                                     # the single sum-reduce of the packed partial canvases
    out, cnt = _finalize(packed, ctx)
    if rank == 0:
        q.put((out.numpy(), cnt.numpy()))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_sharded_blend_matches_sequential_reference():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    out, cnt = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    go = O.GeometryOracle(SHAPE, RAW, SPLIT)
    cfg = O.make_config("vits", SHAPE, RAW, SPLIT)
    _, hr = O.synthetic_frame(cfg, 1)
    random.seed(1)
    want, _, avg = go.infer(torch.zeros(1, 3, *SHAPE), hr, None, MODE, PN)
    assert np.array_equal(cnt, avg.count_map.numpy())                       # bit-exact count map regardless of world size
    rel = np.abs(out - want[0, 0].numpy()) / np.maximum(np.abs(want[0, 0].numpy()), 1e-6)
    assert rel.max() < 1e-3, rel.max()


def _batch_worker(rank, world, port, q, frames):
    """A batch of frames: the flattened F x P work list is split into one contiguous block per rank (model.py / tiling.shard_patches
    with frames > 1), every frame keeps its own packed canvases, ONE all_reduce combines the whole batch."""
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    tc, sched = _schedule_batch(frames)
    P = sum(s.bboxs.shape[0] for s in sched[0])
    own = tiling.shard_patches(frames * P, rank, world, frames=frames)
    parts, ctxs = [], []
    for f in range(frames):
        packed, ctx = _partials(rank, world, stages=sched[f], own=own[f * P:(f + 1) * P])
        parts.append(packed)
        ctxs.append((ctx[0], sched[f], ctx[2], ctx[3]))
    batch = torch.stack(parts)
    dist.all_reduce(batch)                                      # the single sum-reduce of the batch
    outs = [_finalize(batch[f], ctxs[f]) for f in range(frames)]
    if rank == 0:
        q.put(([o[0].numpy() for o in outs], [o[1].numpy() for o in outs], [int(x) for x in np.unique(np.nonzero(own)[0] // P)]))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_batch_of_frames_block_split():
    world, frames = 2, 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_batch_worker, args=(r, world, port, q, frames)) for r in range(world)]
    for p in procs:
        p.start()
    outs, cnts, frames_rank0 = q.get(timeout=180)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert frames_rank0 == [0]                                   # one frame per rank: rank 0 only ever touches (and uploads) frame 0
    go = O.GeometryOracle(SHAPE, RAW, SPLIT)
    cfg = O.make_config("vits", SHAPE, RAW, SPLIT)
    _, hr = O.synthetic_frame(cfg, 1)
    random.seed(1)                                               # ... against two successive single-frame reference runs on one `random` stream
    for f in range(frames):
        want, _, avg = go.infer(torch.zeros(1, 3, *SHAPE), hr, None, MODE, PN)
        assert np.array_equal(cnts[f], avg.count_map.numpy()), f
        rel = np.abs(outs[f] - want[0, 0].numpy()) / np.maximum(np.abs(want[0, 0].numpy()), 1e-6)
        assert rel.max() < 1e-3, (f, rel.max())


def test_every_rank_draws_the_same_schedule():
    a = [s.bboxs.copy() for s in _schedule()[1]]
    b = [s.bboxs.copy() for s in _schedule()[1]]
    assert all(np.array_equal(x, y) for x, y in zip(a, b))
    owns = [tiling.shard_patches(sum(x.shape[0] for x in a), r, 2) for r in range(2)]
    assert np.array_equal(owns[0] + owns[1], np.ones_like(owns[0]))


def _bcast_worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from patchrefinerv2_b200.model import _broadcast_bboxs
    tc = tiling.prepare_tile_cfg(SHAPE, RAW, SPLIT)
    random.seed(100 + rank)                                    # the bad practice the broadcast guards against: seed + rank
    bb = np.concatenate([s.bboxs for s in tiling.schedule(tc, SHAPE, MODE, PN)])
    got = _broadcast_bboxs(bb, torch.device("cpu"))
    q.put((rank, bb, got))
    dist.barrier()
    dist.destroy_process_group()


def test_ranks_seeded_differently_still_blend_rank0s_patches():
    """the sharded forward must not assume identical `random` state on every rank -- rank 0's draw is broadcast."""
    port = _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_bcast_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = dict()
    for _ in range(2):
        r, own, got = q.get(timeout=120)
        res[r] = (own, got)
    for p in procs:
        p.join(60)
    assert not np.array_equal(res[0][0], res[1][0])            # the two ranks really drew different random patches
    assert np.array_equal(res[0][1], res[0][0]) and np.array_equal(res[1][1], res[0][0])
