"""The oracle is the unmodified reference compiled into oracle/_ref (recipe:
oracle/Makefile).  Pin it against the golden values recorded from the
reference in SURVEY.md section 8(c), and check the committed fixtures are
what a fresh reference run produces."""
import os
import re
import subprocess

import numpy as np
import pytest

import cases
import refdump

REF = os.path.join(refdump.ROOT, "oracle", "_ref")
needs_ref = pytest.mark.skipif(not os.path.exists(os.path.join(REF, "SWTest2")),
                               reason="oracle/_ref not built")


def checksums(exe, flags):
    out = subprocess.run([os.path.join(REF, exe)] + flags + ["--output_none"],
                         stdout=subprocess.PIPE, stderr=subprocess.STDOUT,
                         text=True, timeout=1800).stdout
    final = out[out.rindex("(Final)"):]
    return {m.group(1): float(m.group(2))
            for m in re.finditer(r"Checksum \((\w+)\): ([-+0-9.eE]+)", final)}


@needs_ref
def test_sw2_golden_checksums():
    # SWTest2 --resolution 20 --order 4, one step
    c = checksums("SWTest2", ["--resolution", "20", "--order", "4"])
    assert abs(c["U"] - 7.114413176809416e+22) <= 1e-12 * 7.2e22
    assert abs(c["H"] - 1.205365996298435e+18) <= 1e-12 * 1.3e18
    assert abs(c["V"] - 1.651200000000000e+06) <= 1e-12 * 7.2e22


@needs_ref
def test_jw_mini_conserved_checksums():
    # mass and rho-theta after 3 steps.  (U, V, W of
    # this case move with the last bit of the build - DESIGN.md section 4 - so
    # only the conserved sums are pinned; measured with -O3 here.)
    c = checksums("BaroclinicWaveJWTest",
                  ["--resolution", "8", "--levels", "10", "--dt", "200s",
                   "--endtime", "600s", "--pert", "Exp"])
    assert abs(c["Rho"] - 3.782896362711638e+18) <= 1e-12 * 3.8e18
    assert abs(c["RhoTheta"] - 1.172078685972657e+21) <= 1e-11 * 1.2e21


@needs_ref
@pytest.mark.skipif(not refdump.have_ref_dump(), reason="ref_dump not built")
def test_committed_fixture_matches_fresh_reference_run(tmp_path):
    name = "sw2_ne2"
    c = cases.CASES[name]
    fresh = refdump.run_ref_dump(str(tmp_path / "x.bin"), c["case"], c["script"], c["flags"])
    gold = cases.load_case(name)
    # (ref_dump may have gained records since the fixture was written)
    assert set(gold) <= set(fresh), set(gold) - set(fresh)
    for k in gold:
        assert np.array_equal(fresh[k], gold[k], equal_nan=True), k


@needs_ref
@pytest.mark.skipif(not refdump.have_ref_dump(), reason="ref_dump not built")
def test_committed_l30_fixture_matches_fresh_reference_run(tmp_path):
    """The L = 30 Strang fixture (run records only; geometry is shared)."""
    name = "jw_ne2_l30_strang"
    c = cases.CASES[name]
    fresh = refdump.run_ref_dump(str(tmp_path / "x.bin"), c["case"], c["script"], c["flags"])
    with np.load(cases.golden_path(name)) as z:
        for k in z.files:
            assert np.array_equal(fresh[k], z[k], equal_nan=True), k


@needs_ref
@pytest.mark.skipif(not refdump.have_ref_dump(), reason="ref_dump not built")
@pytest.mark.parametrize("key", ["jw_ne8_l10_strang_3steps", "jwtr_ne2_l6_ars343_2steps"])
def test_reference_sensitivity_is_what_the_json_records(key):
    """tests/golden/sensitivity.json (the reference's spread against itself
    under 1e-15 perturbations, which bounds the loosened parity tolerances)
    re-measured on the reference built here."""
    import make_sensitivity as ms
    gold = ms.load()[key]
    fresh = ms.ENTRIES[key]()
    if "checksum" in gold:
        assert np.allclose(fresh["checksum"], gold["checksum"], rtol=1e-13, atol=0)
        pairs = zip(fresh["abs_spread"], gold["abs_spread"])
    else:
        assert np.allclose(fresh["mass"], gold["mass"], rtol=1e-13, atol=0)
        pairs = zip(fresh["field_rel_spread"] + fresh["mass_rel_spread"],
                    gold["field_rel_spread"] + gold["mass_rel_spread"])
    for a, b in pairs:
        assert a == b, (fresh, gold)     # same binary, same input: deterministic
