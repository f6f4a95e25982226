"""Times mc_gemm_bf16 (the tcgen05 GEMM) over the linear shapes of a model for a few row counts: CUDA events, 20 launches."""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from metalchat_b200 import capi  # noqa: E402

dev = capi.Device(0)
shapes = [("wqkv", 3072, 2048), ("wo", 2048, 2048), ("w13", 16384, 2048), ("w2", 2048, 8192), ("head", 128256, 2048)]
for M in [int(a) for a in sys.argv[1:]] or [32, 2048]:
    for name, N, K in shapes:
        if M > 256 and name == "head":
            continue
        x, w, y = dev.alloc(M * K * 2), dev.alloc(N * K * 2), dev.alloc(M * N * 2)
        for b, n in ((x, M * K * 2), (w, N * K * 2)):
            capi.check(capi.lib().mc_memset(dev.h, b.h, 0, 0x3c, n))
        capi.gemm_bf16(dev, y, x, w, M, N, K, iters=3)
        ms = capi.gemm_bf16(dev, y, x, w, M, N, K, iters=20) / 20
        print(f"M={M:5d} {name:5s} N={N:6d} K={K:5d}  {ms * 1e3:8.1f} us  {2.0 * M * N * K / ms / 1e9:8.1f} TFLOP/s  {N * K * 2 / ms / 1e6:8.1f} GB/s (weights)", flush=True)
        for b in (x, w, y):
            b.release()
