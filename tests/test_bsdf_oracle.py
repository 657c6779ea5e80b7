"""Pins the oracle's BSDF layer with the reference's own property suite (tests/test_bsdf.cpp restated in bsdf_cases.py):
every TEST_F / TEST_P instantiation whose BSDF is on the GltfBSDF tree, same sample count (2^18), same XorShift32
seed (0x1234567), same tolerances."""
import pytest

import bsdf_cases as bc

CASES = bc.cases(include_albedo=True)


@pytest.mark.parametrize("name,params,checks,extra", CASES, ids=[c[0] for c in CASES])
def test_reference_bsdf_suite_on_oracle(oracle_lib, name, params, checks, extra):
    from oracle.oracle_py import bsdf_check
    bc.run_case(bsdf_check, params, checks, extra)
