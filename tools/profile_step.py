"""Per-launch (per-op path) or per-phase (streaming kernel) device times of one decode step.
Usage: python tools/profile_step.py [1b|8b] [stream|ops] [quant]"""
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
quant = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dev = capi.Device(0)
FLAGS = {"ops": capi.LLAMA_NO_STREAM, "stream": 0}
m = capi.Llama(dev, capi.llama_config(**shape, max_seq_len=1024, flags=FLAGS[mode], quant=quant))
m.init_random(0x5EED)
m.finalize()
m.prefill(np.arange(512, dtype=np.int32) % shape["vocab"])
m.decode_loop([1], [512], 8)
L = shape["n_layers"]
if mode == "stream":
    names = ["qkv", "attn", "wo", "w13", "w2"] * L + ["head"]
    P = len(names)
    reps = []
    for rep in range(4):
        raw = m.profile_step(1)
        dbg = raw[-512:].reshape(64, 8)
        t = raw[:-512].reshape(-1, P, 4)  # [cta][phase][entry, -, staged, done] in us
        m.decode_loop([1], [512], 2)
        if rep:
            reps.append(t)
    t = np.mean(reps, axis=0)
    G = t.shape[0]
    print(f"{G} CTAs, step = {t[:, -1, 3].max() - t[:, 0, 0].min():.1f} us")
    print("phase   | first entry -> last done | stage(wait+stage) med/max | tiles med/max | done skew (max-min) | slowest CTA")
    for kind in ["qkv", "attn", "wo", "w13", "w2", "head"]:
        idx = [i for i, n in enumerate(names) if n == kind]
        span = np.mean([t[:, i, 3].max() - t[:, i, 0].min() for i in idx])
        stage = t[:, idx, 2] - t[:, idx, 0]
        tiles = t[:, idx, 3] - t[:, idx, 2]
        skew = np.mean([t[:, i, 3].max() - t[:, i, 3].min() for i in idx])
        slow = np.bincount(np.concatenate([[int(t[:, i, 3].argmax())] for i in idx]), minlength=G).argmax()
        print(f"{kind:6s}  | {span:7.2f} | {np.median(stage):6.2f} / {stage.max(axis=0).mean():6.2f} | {np.median(tiles):6.2f} / {tiles.max(axis=0).mean():6.2f} | {skew:6.2f} | {slow}")
    # epilogue lag: last epilogue store of a phase vs the mma warps leaving it; propagation: next phase staged vs the last epilogue store anywhere
    for kind, nxt in [("qkv", "attn"), ("wo", "w13"), ("w13", "w2"), ("w2", "qkv")]:
        idx = [i for i, n in enumerate(names) if n == kind and i + 1 < P and names[i + 1] == nxt]
        lag = np.array([t[:, i, 1] - t[:, i, 3] for i in idx])
        prop = np.array([t[:, i + 1, 2] - t[:, i, 1].max() for i in idx])
        print(f"{kind}->{nxt}: epilogue lag med {np.median(lag):.2f} max {lag.max(axis=1).mean():.2f} us; staged after last epilogue: med {np.median(prop):.2f} max {prop.max(axis=1).mean():.2f} us")
    done = t[:, :, 3]
    lag = (done - done.min(axis=0, keepdims=True)).mean(axis=1)
    order = np.argsort(lag)
    print("mean lag behind the first finisher per CTA: min", lag[order[:5]].round(2), order[:5], "max", lag[order[-8:]].round(2), order[-8:])
    np.save("gpurun_out/stream_stamps.npy", t)
    if dbg[:, 0].any():
        d = dbg[:40]
        print("head blocks of CTA 0 (us): wait-full, mma, hand-over | epilogue: ready after mma-done, epilogue time")
        for i in range(0, 40, 4):
            print(i, " ".join(f"[{d[j,1]-d[j,0]:.2f} {d[j,2]-d[j,1]:.2f} {d[j,3]-d[j,2]:.2f} | {d[j,4]-d[j,3]:.2f} {d[j,5]-d[j,4]:.2f}]" for j in range(i, i + 4)))
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
