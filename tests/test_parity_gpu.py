"""GPU parity tests (-m gpu): the CUDA path, called through the C ABI (ctypes), against
(1) the CPU oracle on the same inputs and (2) the committed golden vectors generated from the
reference's own headers.  Tolerance = BASELINE.json north_star: 1e-8 abs + 1e-10 rel (fp64)."""
import numpy as np
import pytest

from tests import cases as CASES
from tests.conftest import assert_parity, golden_expect

pytestmark = pytest.mark.gpu

CAT = CASES.catalogue()
IDS = [c["name"] for c in CAT]


@pytest.fixture(scope="module")
def engine():
    import fftoptionpricing_b200 as F
    eng = F.Engine(0)
    yield eng
    eng.close()


@pytest.mark.parametrize("case", CAT, ids=IDS)
def test_case_matches_oracle_and_golden(case, engine, best_oracle, golden):
    got = CASES.run_case(engine, case)
    ref = CASES.run_case(best_oracle, case)
    assert_parity(got, ref, what=f"gpu vs {best_oracle.flavor} oracle: {case['name']}")
    g, e = golden_expect(golden[case["name"]], got)
    assert_parity(g, e, what=f"gpu vs golden: {case['name']}")


@pytest.mark.parametrize("case", [c for c in CAT if c.get("kat")],
                         ids=[c["name"] for c in CAT if c.get("kat")])
def test_reference_known_answers(case, engine):
    """The constants the reference's own test.cpp asserts, through the CUDA path."""
    out = CASES.run_case(engine, case)
    for idx, expected, eps in case["kat"]:
        assert abs(out[idx] - expected) < eps * (1.0 + max(abs(out[idx]), abs(expected)))
