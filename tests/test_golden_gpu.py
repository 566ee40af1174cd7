"""The CUDA product, through the C ABI, against the committed outputs of the reference's own GPU kernels
(tests/golden/*.npz) — the same check as tests/test_reference_gpu.py but independent of oracle/_ref being present."""
import pytest

from tests import util

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("case", util.golden_cases())
def test_product_matches_reference_fixture(case):
    from sparkl_b200.pipeline import context_from_scene
    worst = util.run_against_golden(lambda sc: context_from_scene(sc), case, base=5e-5 if case in ("fluids3", "fluids2") else 2e-5, dt_rtol=1e9 if case == "fluids2" else 1e-5)  # fluids2: dt not compared, see tests/test_dim2_gpu.py
    print(case, "worst relative deviations of the product vs the reference fixture:", {k: f"{v:.2e}" for k, v in worst.items()})
