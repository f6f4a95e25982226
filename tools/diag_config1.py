"""Prints the config-1 (one 1B layer, 127 cached + 1 new position) agreement metrics of the two prompt paths against the oracles."""
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from metalchat_b200 import capi  # noqa: E402
from oracle import orc  # noqa: E402
from oracle.orc import BF16, F32  # noqa: E402

cfgd = dict(dim=2048, n_layers=1, n_heads=32, n_kv_heads=8, head_dim=64, ffn_dim=8192, vocab=128256, max_seq_len=256)
unbf = orc.bf16_to_f32
o = orc.Llama(orc.make_cfg(**cfgd), BF16)
o.init_random(0x5EED)
of = orc.Llama(orc.make_cfg(**cfgd), F32)
of.init_random(0x5EED)
ids = [int(orc.lib().orc_hash_int(0x5EED, 0xFFFF, i, 0, cfgd["vocab"])) for i in range(128)]
o.forward(ids[:127], 0)
of.forward(ids[:127], 0)
lb, hb = o.forward(ids[127:], 127, want_hidden=True)
lf, hf = of.forward(ids[127:], 127, want_hidden=True)
dev = capi.Device(0)
for name, flags in (("tensor-core prompt path", 0), ("4-row GEMV prompt path", capi.LLAMA_NO_TC_PREFILL)):
    m = capi.Llama(dev, capi.llama_config(**cfgd, flags=flags))
    m.init_random(0x5EED)
    m.finalize()
    m.prefill(ids[:127])
    for which in (0, 1):
        got = unbf(m.cache(0, 0, which, 127).reshape(-1))
        want = unbf(o.cache(0, 0, which)[: got.size])
        print(name, "cache", which, "exact", float(np.mean(got == want)), "max_rel", float(np.abs(got - want).max() / np.abs(want).max()))
    m.prefill(ids[127:], start_pos=127)
    h = m.hidden()
    ulps = np.abs(h.astype(np.int32) - hb[-1].astype(np.int32))
    print(name, "hidden vs f32 max_rel", float(np.abs(unbf(h) - hf[-1]).max() / np.abs(hf[-1]).max()),
          "| vs bf16 oracle: median ulps", float(np.median(ulps)), "mean rel", float(np.abs(unbf(h) - unbf(hb[-1])).mean() / np.abs(unbf(hb[-1])).mean()),
          "max_rel", float(np.abs(unbf(h) - unbf(hb[-1])).max() / np.abs(unbf(hb[-1])).max()), "exact", float(np.mean(h == hb[-1])))
    print(name, "mean rel of f32-oracle vs bf16-oracle", float(np.abs(hf[-1] - unbf(hb[-1])).mean() / np.abs(unbf(hb[-1])).mean()))
    m.close()
