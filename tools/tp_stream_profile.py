"""Per-phase stamps of one decode step of the streaming kernel under tensor parallelism.
Usage: torchrun --nproc-per-node N tools/tp_stream_profile.py [8b|1b]   (or plain python for one GPU)"""
import os
import pathlib
import sys

import numpy as np

sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import bench  # noqa: E402
from metalchat_b200 import capi, tp  # noqa: E402

shape = bench.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "8b"]
rank, world, local = tp.env_rank_world()
dist = None
if world > 1:
    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local)
    dist.init_process_group("gloo")
dev = capi.Device(local)
m = tp.create(dev, **shape, max_seq_len=1024) if world > 1 else capi.Llama(dev, capi.llama_config(**shape, max_seq_len=1024))
m.init_random(0x5EED)
m.finalize()
m.prefill(np.arange(512, dtype=np.int32) % shape["vocab"])
m.decode_loop([1], [512], 8)
L = shape["n_layers"]
names = ["qkv", "attn", "wo", "w13", "w2"] * L + ["head"]
P = len(names)
reps = []
for rep in range(4):
    raw = m.profile_step(1)
    t = raw[:-512].reshape(-1, P, 4)  # [cta][phase][entry, epilogue done, staged, tiles done] in us
    m.decode_loop([1], [512], 2)
    if rep:
        reps.append(t)
t = np.mean(reps, axis=0)
if rank == 0:
    G = t.shape[0]
    print(f"world {world}: {G} CTAs, step = {t[:, -1, 3].max() - t[:, 0, 0].min():.1f} us")
    print("phase   | first entry -> last done | stage(wait+stage) med/max | tiles med/max | done skew (max-min)")
    tot = 0.0
    for kind in ["qkv", "attn", "wo", "w13", "w2", "head"]:
        idx = [i for i, n in enumerate(names) if n == kind]
        span = np.mean([t[:, i, 3].max() - t[:, i, 0].min() for i in idx])
        stage = t[:, idx, 2] - t[:, idx, 0]
        tiles = t[:, idx, 3] - t[:, idx, 2]
        skew = np.mean([t[:, i, 3].max() - t[:, i, 3].min() for i in idx])
        # time the phase adds to the critical path: from the previous phase's median tiles-done to this phase's median tiles-done
        adv = np.mean([np.median(t[:, i, 3]) - np.median(t[:, i - 1, 3]) for i in idx if i > 0])
        tot += adv * len(idx)
        print(f"{kind:6s}  | {span:7.2f} | {np.median(stage):6.2f} / {stage.max(axis=0).mean():6.2f} | {np.median(tiles):6.2f} / {tiles.max(axis=0).mean():6.2f} | {skew:6.2f} | advance {adv:6.2f} us x {len(idx)}")
    for kind, nxt in [("qkv", "attn"), ("wo", "w13"), ("w13", "w2"), ("w2", "qkv")]:
        idx = [i for i, n in enumerate(names) if n == kind and i + 1 < P and names[i + 1] == nxt]
        lag = np.array([t[:, i, 1] - t[:, i, 3] for i in idx])
        prop = np.array([t[:, i + 1, 2] - t[:, i, 1].max() for i in idx])
        print(f"{kind}->{nxt}: epilogue lag med {np.median(lag):.2f} max {lag.max(axis=1).mean():.2f} us; staged after last epilogue: med {np.median(prop):.2f} max {prop.max(axis=1).mean():.2f} us")
if dist is not None:
    dist.barrier()
    dist.destroy_process_group()
