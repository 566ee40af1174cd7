"""Pins the CPU oracle (oracle/mpm_oracle.cpp) against committed outputs of the reference's own GPU kernels
(tests/golden/*.npz; generated on a B200 by scripts/make_golden.py from the PTX the reference ships).  Runs without a GPU."""
import pytest

from tests import util


@pytest.mark.parametrize("case", util.golden_cases())
def test_oracle_matches_reference_fixture(case, oracle):
    def make(sc):
        dim2 = sc.particles.position.dtype["point"].shape[0] == 2  # the dim2 fixtures come from the reference's sparkl2d kernels
        e = (oracle.OracleEngine2D if dim2 else oracle.OracleEngine)(sc.cell_width)
        e.models_upload(sc.models); e.colliders_upload(sc.colliders); e.upload(sc.particles)
        return e
    worst = util.run_against_golden(make, case, base=5e-5 if case in ("fluids3", "fluids2") else 2e-5, dt_rtol=1e9 if case == "fluids2" else 1e-5)  # fluids2: dt not compared, see tests/test_dim2_gpu.py
    print(case, "worst relative deviations of the oracle vs the reference fixture:", {k: f"{v:.2e}" for k, v in worst.items()})


def test_fixtures_present():
    assert set(util.golden_cases()) >= {"sand3", "fluids3", "neo_hookean_block", "cube_through_sand3"}
