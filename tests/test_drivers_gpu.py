"""GPU: the reference's unmodified example/inverse/{macro,micro,hybrid}.py drivers run on the B200 kernels through
the drop-in packages and reproduce the curves of the reference's own CPU lanes (BASELINE.json configs[0..2])."""
import pytest

from drivers_cases import run_driver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("problem", ["macro", "micro", "hybrid"])
def test_unmodified_inverse_driver_on_gpu(problem, tmp_path, dev):
    line = run_driver(problem, tmp_path)
    print(problem, "s/episode", line["s_per_episode"])
