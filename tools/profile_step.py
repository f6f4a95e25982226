"""Per-launch (per-op path) or per-phase (megakernel) device times of one decode step.
Usage: python tools_profile_step.py [1b|8b] [mega|ops]"""
import sys
import numpy as np
import pathlib as _p, sys as _s
_s.path.insert(0, str(_p.Path(__file__).resolve().parent.parent))
from metalchat_b200 import capi
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import bench

shape = bench.SHAPES[sys.argv[1] if len(sys.argv) > 1 else "1b"]
mode = sys.argv[2] if len(sys.argv) > 2 else "stream"
dev = capi.Device(0)
FLAGS = {"mega": capi.LLAMA_MEGAKERNEL, "ops": capi.LLAMA_NO_STREAM, "stream": 0}
m = capi.Llama(dev, capi.llama_config(**shape, max_seq_len=1024, flags=FLAGS[mode]))
m.init_random(0x5EED)
m.finalize()
m.prefill(np.arange(512, dtype=np.int32) % shape["vocab"])
m.decode_loop([1], [512], 8)
L = shape["n_layers"]
if mode == "stream":
    names = ["qkv", "attn", "wo", "w13", "w2"] * L + ["head"]
    acc = {}
    for rep in range(4):
        us = m.profile_step(1).reshape(-1, 4)
        m.decode_loop([1], [512], 2)
        if rep == 0:
            continue
        for n, t in zip(names, us):
            acc.setdefault(n, []).append(t)
    tot = 0
    for n, v in acc.items():
        v = np.array(v)
        per = v.mean(axis=0)
        cnt = len(v) / 3
        tot += per.sum() * cnt
        print(f"{n:6s} wait {per[0]:7.2f}  stage {per[1]:7.2f}  tiles {per[2]:7.2f}  gap {per[3]:7.2f} us  x{cnt:3.0f} = {per.sum() * cnt:8.1f} us")
    print("sum", tot)
elif mode == "mega":
    names = ["qkv", "attn", "wo", "w13", "w2"] * L + ["head"]
    acc = {}
    for rep in range(4):
        us = m.profile_step(1).reshape(-1, 3)
        m.decode_loop([1], [512], 2)
        if rep == 0:
            continue
        for n, t in zip(names, us):
            acc.setdefault(n, []).append(t)
    tot = 0
    for n, v in acc.items():
        v = np.array(v)
        per = v.mean(axis=0)
        cnt = len(v) / 3
        tot += per.sum() * cnt
        print(f"{n:6s} wait {per[0]:7.2f}  work {per[1]:7.2f}  gap {per[2]:7.2f} us  x{cnt:3.0f} = {per.sum() * cnt:8.1f} us")
    print("sum", tot)
else:
    names = ["embed"] + ["qkv", "attn", "wo", "w13", "w2"] * L + ["head", "argmax1", "argmax2"]
    acc = {}
    for rep in range(5):
        us = m.profile_step(1)
        assert len(us) == len(names), (len(us), len(names))
        if rep == 0:
            continue
        for n, t in zip(names, us):
            acc.setdefault(n, []).append(float(t))
    tot = 0
    for n, v in acc.items():
        per = np.mean(v)
        cnt = len(v) / 4
        tot += per * cnt
        print(f"{n:8s} avg {per:7.2f} us x {cnt:4.0f} = {per * cnt:8.1f} us")
    print("sum", tot)
