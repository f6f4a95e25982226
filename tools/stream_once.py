"""A few decode steps through the streaming kernel (for ncu captures). Usage: stream_once.py [steps_per_launch] [launches] [1b|8b] [kv_len]"""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
from metalchat_b200 import capi
import bench

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 1
launches = int(sys.argv[2]) if len(sys.argv) > 2 else 2
shape = bench.SHAPES[sys.argv[3] if len(sys.argv) > 3 else "1b"]
kv = int(sys.argv[4]) if len(sys.argv) > 4 else 512
quant = int(sys.argv[5]) if len(sys.argv) > 5 else 0
dev = capi.Device(0)
m = capi.Llama(dev, capi.llama_config(**shape, max_seq_len=1024, quant=quant))
m.init_random(0x5EED)
m.finalize()
m.prefill(np.arange(kv, dtype=np.int32) % shape["vocab"])
pos = kv
for i in range(launches):
    toks, ms = m.decode_loop([1], [pos], steps)
    pos += steps
    print("launch", i, "steps", steps, "ms/step", ms / steps, "tokens", toks[:4, 0].tolist())
