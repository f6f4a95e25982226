"""Isolated decode GEMV kernels (batch 1) over the linear shapes of the 1B / 8B / 70B (TP 8 shard) models: achieved HBM GB/s on the
weight stream, bf16 (`gemv_bf16_kernel`) and packed int4 (`gemv_q_kernel`), CUDA events on the engine stream.  Every shape runs over
a ring of distinct weight buffers larger than L2 in aggregate, so no launch re-reads cached weights.  Prints one JSON document."""
import json
import sys
from pathlib import Path

import numpy as np

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from metalchat_b200 import capi  # noqa: E402

import torch  # noqa: E402

PEAK = json.loads((Path(__file__).resolve().parent.parent / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]
MODELS = {
    "1b": [("wqkv", 3072, 2048), ("wo", 2048, 2048), ("w13", 16384, 2048), ("w2", 2048, 8192), ("head", 128256, 2048)],
    "8b": [("wqkv", 6144, 4096), ("wo", 4096, 4096), ("w13", 28672, 4096), ("w2", 4096, 14336)],
    "70b-tp8": [("wqkv", 1280, 8192), ("wo", 8192, 1024), ("w13", 7168, 8192), ("w2", 8192, 3584), ("head", 16032, 8192)],
}
dev = capi.Device(0)
st = torch.cuda.ExternalStream(dev.stream())
out = {"peak_gbs": PEAK, "peak_source": "MEASURED_PEAKS.json hbm_gbs", "rows": []}
for model, shapes in MODELS.items():
    for name, N, K in shapes:
        for fmt in ("bf16", "w4"):
            if fmt == "w4" and name == "head":
                continue  # the quantised head is int8 per row (quantization/linear.h), not part of this sweep
            if fmt == "bf16":
                wbytes = N * K * 2
            else:
                wb, sb = capi.w4_sizes(N, K)
                wbytes = wb + sb
            ring = max(2, min(12, int(300e6 // wbytes) + 1))
            ws = []
            for _ in range(ring):
                if fmt == "bf16":
                    w = dev.alloc(wbytes)
                    capi.check(capi.lib().mc_memset(dev.h, w.h, 0, 0x3c, wbytes))
                    ws.append((w, None))
                else:
                    w, s = dev.alloc(wb), dev.alloc(sb)
                    capi.check(capi.lib().mc_memset(dev.h, w.h, 0, 0x77, wb))
                    capi.check(capi.lib().mc_memset(dev.h, s.h, 0, 0x3c, sb))
                    ws.append((w, s))
            x = dev.upload(np.full(K, 0x3c00, np.uint16))
            y = dev.alloc(N * 2)

            def run(i):
                w, s = ws[i % ring]
                if fmt == "bf16":
                    capi.linear_bf16(dev, y, x, w, 1, N, K)
                else:
                    capi.linear_w4(dev, y, x, w, s, 1, N, K)

            for i in range(ring):
                run(i)
            dev.synchronize()
            reps = max(ring * 3, 24)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(st)
            for i in range(reps):
                run(i)
            e1.record(st)
            dev.synchronize()
            us = e0.elapsed_time(e1) / reps * 1e3
            gbs = (wbytes + K * 2 + N * 2) / us / 1e3
            row = {"model": model, "linear": name, "N": N, "K": K, "format": fmt, "weight_MB": wbytes / 1e6, "us": us, "GBps": gbs, "frac_of_peak": gbs / PEAK}
            out["rows"].append(row)
            print(f"{model:8s} {name:5s} {fmt:4s} N={N:6d} K={K:5d} {wbytes / 1e6:8.1f} MB {us:8.1f} us {gbs:8.1f} GB/s {gbs / PEAK:5.2f}", file=sys.stderr, flush=True)
            for w, s in ws:
                w.release()
                if s is not None:
                    s.release()
            x.release(), y.release()
print(json.dumps(out))
