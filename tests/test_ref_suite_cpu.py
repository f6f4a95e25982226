"""The reference's own unit tests (test/*.cc, unmodified, compiled where they lie) run against the façade over the CPU backend of the
C ABI, whose kernels are the oracle's op functions (oracle/ref/orc_mc_abi.cc).  This is what pins the op-level oracle with the
reference's own known-answer tests AS EXECUTED BY THE REFERENCE'S OWN WRAPPER CODE (launch shapes, block sizes, padding, views).
The binaries exist only where /root/reference does (this container); they are built by __graft_entry__.build()."""
import pytest

from tests import ref_tests

pytestmark = pytest.mark.skipif(not (ref_tests.REF_BIN / "cpu").exists(), reason="oracle/_ref/cpu not built (needs /root/reference; run __graft_entry__.build())")


@pytest.mark.parametrize("name", ref_tests.TESTS)
def test_reference_unit_test_passes_on_the_oracle_backend(name):
    r = ref_tests.run("cpu", name)
    assert r["failed"] == 0 and r["rc"] == 0, r["out"][-3000:]
    assert r["trapped"] == ref_tests.EXPECTED_TRAPS.get(name, 0), r["out"][-2000:]
    assert r["passed"] >= 1


def test_reference_unit_tests_cover_the_expected_number_of_cases():
    total = sum(ref_tests.run("cpu", n)["passed"] for n in ref_tests.TESTS)
    assert total >= ref_tests.MIN_CASES, total
