"""Debug: one decode step through the streaming kernel vs the per-op path, cache / logits diffs layer by layer."""
import sys, pathlib
sys.path.insert(0, str(pathlib.Path(__file__).resolve().parent.parent))
import numpy as np
from metalchat_b200 import capi

def unbf(a):
    return (a.astype(np.uint32) << 16).view(np.float32)

SMALL = dict(dim=512, n_layers=3, n_heads=8, n_kv_heads=2, head_dim=64, ffn_dim=1024, vocab=2000, max_seq_len=96)
import sys as _sys
Q = int(_sys.argv[1]) if len(_sys.argv) > 1 else 0
dev = capi.Device(0)
ms = []
for flags in (0, 16):
    m = capi.Llama(dev, capi.llama_config(**SMALL, flags=flags, quant=Q))
    m.init_random(0x5EED)
    m.finalize()
    ids = [3, 77, 512, 999, 0, 41, 41, 7, 1500, 2]
    m.prefill(ids)
    out = m.decode([292], [len(ids)])
    print("flags", flags, "token", out)
    ms.append(m)
a, b = ms
n = 11
for layer in range(SMALL["n_layers"]):
    for which in (0, 1):
        ca, cb = unbf(a.cache(0, layer, which, n)), unbf(b.cache(0, layer, which, n))
        d = np.abs(ca - cb)
        print("layer", layer, "KV"[which], "max diff old rows", d[:10].max(), "new row", d[10].max(), "ref max", np.abs(cb[10]).max())
la, lb = unbf(a.logits()), unbf(b.logits())
print("logits diff", np.abs(la - lb).max(), np.abs(lb).max(), "argmax", la.argmax(), lb.argmax())
