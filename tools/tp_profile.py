"""Per-launch device times of one tensor-parallel decode step (run under torchrun, one rank per GPU): rank 0 prints the mean time of
every kernel kind next to the single-GPU per-op numbers.  Usage: torchrun --nproc-per-node N tools/tp_profile.py [1b|8b]"""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import bench  # noqa: E402
from metalchat_b200 import capi, tp  # noqa: E402

rank, world, local = tp.env_rank_world()
torch.cuda.set_device(local)
dist.init_process_group("gloo")
shape = bench.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "8b"]
dev = capi.Device(local)
L = shape["n_layers"]
names = ["embed"] + ["qkv", "attn", "wo", "w13", "w2"] * L + ["head", "argmax1", "argmax2"]


def profile(m, label):
    m.prefill(np.arange(64, dtype=np.int32) % shape["vocab"])
    m.decode_loop([1], [64], 4)
    acc = {}
    for rep in range(4):
        us = m.profile_step(1)
        dist.barrier()
        if rep == 0:
            continue
        assert len(us) == len(names), (len(us), len(names))
        for n, t in zip(names, us):
            acc.setdefault(n, []).append(float(t))
    if rank == 0:
        tot = 0.0
        for n, v in acc.items():
            per, cnt = np.mean(v), len(v) / 3
            tot += per * cnt
            print(f"{label:8s} {n:8s} avg {per:8.2f} us x {cnt:4.0f} = {per * cnt:9.1f} us", flush=True)
        print(f"{label:8s} sum {tot:.1f} us", flush=True)


m = tp.create(dev, **shape, max_seq_len=256)
m.init_random(0x5EED)
m.finalize()
profile(m, f"tp{world}")
m.close()
if rank == 0:
    pass
single = capi.Llama(dev, capi.llama_config(**shape, max_seq_len=256, flags=capi.LLAMA_NO_STREAM))
single.init_random(0x5EED)
single.finalize()
profile(single, "single")
dist.barrier()
dist.destroy_process_group()
