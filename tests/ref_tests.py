"""Shared helper: runs the prebuilt binaries of the reference's OWN unit tests (oracle/_ref/{cpu,cuda}/test_*, built by
oracle/ref/Makefile from /root/reference/test/*.cc with the façade of metalchat_b200/facade) and parses their reports."""
import re
import subprocess
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
REF_BIN = ROOT / "oracle" / "_ref"

# the reference's unit tests that build without its absent third-party dependencies (oracle/ref/Makefile TESTS)
TESTS = ["test_accelerator", "test_allocator", "test_concatenate", "test_functional", "test_indexing", "test_iterator",
         "test_kernel_activation", "test_kernel_arithmetic", "test_kernel_bmm", "test_kernel_copy", "test_kernel_embedding",
         "test_kernel_logical", "test_kernel_mul", "test_kernel_multinomial", "test_kernel_rmsnorm", "test_kernel_roll",
         "test_kernel_softmax", "test_kernel_sort", "test_kernel_sum", "test_kernel_thread", "test_layer", "test_tensor", "test_triu"]
# test cases that divide an integer by zero: 0 on arm64 (the reference's only target), SIGFPE on x86-64 -- a host-CPU difference
EXPECTED_TRAPS = {"test_tensor": 1}
MIN_CASES = 62  # non-integration, non-benchmark TEST_CASEs in those files


def run(backend: str, name: str, timeout: int = 900):
    exe = REF_BIN / backend / name
    res = subprocess.run([str(exe)], capture_output=True, text=True, timeout=timeout, cwd=ROOT)
    m = re.search(r"== (\d+) passed, (\d+) failed, (\d+) skipped, (\d+) trapped", res.stdout)
    assert m, f"{name} ({backend}) did not finish: rc {res.returncode}\n{res.stdout[-2000:]}\n{res.stderr[-2000:]}"
    passed, failed, skipped, trapped = map(int, m.groups())
    return dict(passed=passed, failed=failed, skipped=skipped, trapped=trapped, rc=res.returncode, out=res.stdout)
