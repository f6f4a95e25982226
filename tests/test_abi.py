"""CPU-side checks of the drop-in boundary: libmc_cuda.so builds, loads and exports every symbol
include/mc_cuda.h declares; the kernel registry answers all 71 host names of metalchat.metallib
(SURVEY.md appendix A); without a GPU the product fails loudly instead of falling back."""
import ctypes as C
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    from metalchat_b200 import build, capi

    build.build()
    return capi.lib()


def declared_symbols():
    text = (ROOT / "include" / "mc_cuda.h").read_text()
    return sorted(set(re.findall(r"MC_API\s+[\w\s\*]+?\b(mc_\w+)\s*\(", text)))


def test_header_symbols_are_exported(lib):
    from metalchat_b200 import capi

    names = declared_symbols()
    assert len(names) >= 55
    for n in names:
        assert hasattr(lib, n), f"{n} declared in include/mc_cuda.h but not exported"
    assert set(names) == set(capi._SIGNATURES), set(names) ^ set(capi._SIGNATURES)


def reference_kernel_names():
    both = ["silu", "gelu", "add", "add_broadcast", "sub", "div", "bmm_8", "scatter", "embedding", "gt", "le", "hadamard",
            "scalar_mul", "multinomial", "rmsnorm", "roll", "rope", "softmax", "sort", "sum"]
    names = [f"{k}_{t}" for k in both for t in ("bfloat", "float")]
    names += [f"copy_{t}" for t in ("bfloat", "float", "int32_t")]
    names += [f"gather_{t}" for t in ("bfloat", "float", "int32_t")]
    names += [f"cumsum_{b}_{t}" for b in (2, 4, 8, 16, 32, 64, 128, 256, 512, 1024) for t in ("bfloat", "float")]
    names += [f"hadamard_broadcast_{o}_int8_t_{s}" for o in ("bfloat", "float") for s in ("bfloat", "float")]
    names += ["rope_freqs_float"]
    return names


def test_registry_has_all_reference_kernels(lib):
    n = C.c_int()
    assert lib.mc_kernel_count(C.byref(n)) == 0
    have = set()
    for i in range(n.value):
        p = C.c_char_p()
        assert lib.mc_kernel_name_at(i, C.byref(p)) == 0
        have.add(p.value.decode())
    want = reference_kernel_names()
    assert len(want) == 71
    assert set(want) == have


def test_no_cpu_fallback(lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    from metalchat_b200 import capi

    with pytest.raises(capi.McError, match="no CUDA device"):
        capi.Device(0)


def test_product_does_not_import_oracle():
    for p in (ROOT / "metalchat_b200").rglob("*"):
        if p.suffix in (".py", ".cu", ".cuh", ".h"):
            text = p.read_text()
            assert "from oracle" not in text and "import oracle" not in text and "liborc" not in text, p
