"""Shared helpers of the GPU parity tests."""
import numpy as np
import pytest


def require_gpu():
    import torch

    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


_gpu = None


def accelerator():
    """One hardware_accelerator per test process (the product raises if the extension is missing)."""
    global _gpu
    require_gpu()
    if _gpu is None:
        from metalchat_b200 import ops

        _gpu = ops.Accelerator(0)
    return _gpu


def bf(a):
    from oracle import orc

    return orc.f32_to_bf16(np.asarray(a, dtype=np.float32))


def unbf(a):
    from oracle import orc

    return orc.bf16_to_f32(a)
