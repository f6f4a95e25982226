"""The reference's own unit tests (test/*.cc, unmodified) on the B200: the same objects as tests/test_ref_suite_cpu.py, linked with
libmc_cuda.so instead of the oracle backend.  This is the drop-in claim of north_star made executable: reference-style C++ user code
(hardware_accelerator, kernel::*, tensors, futures, allocators) compiled against the reference's headers runs on the CUDA backend
through the five façade translation units (metalchat_b200/facade) and the C ABI (include/mc_cuda.h).

The binaries are built in the authoring container (the reference's sources do not travel) and shipped in oracle/_ref/cuda."""
import pytest

from tests import ref_tests
from tests.gpu_util import require_gpu

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not (ref_tests.REF_BIN / "cuda").exists(), reason="oracle/_ref/cuda was not shipped")]


@pytest.mark.parametrize("name", ref_tests.TESTS)
def test_reference_unit_test_passes_on_b200(name):
    require_gpu()
    r = ref_tests.run("cuda", name)
    assert r["failed"] == 0 and r["rc"] == 0, r["out"][-3000:]
    assert r["trapped"] == ref_tests.EXPECTED_TRAPS.get(name, 0), r["out"][-2000:]
    assert r["passed"] >= 1
