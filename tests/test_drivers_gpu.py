"""GPU: the reference's unmodified example/inverse/{macro,micro,hybrid}.py drivers run on the B200 kernels through
the drop-in packages and reproduce the curves of the reference's own CPU lanes (BASELINE.json configs[0..2]), in
both precision tiers, with deferred stepping (the default) and with one launch per lane and step."""
import pytest

from drivers_cases import run_driver

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tier", ["fp64", "fp32"])
@pytest.mark.parametrize("problem", ["macro", "micro", "hybrid"])
def test_unmodified_inverse_driver_on_gpu(problem, tier, tmp_path, dev):
    line = run_driver(problem, tmp_path, tier=tier)
    print(problem, tier, "s/episode", line["s_per_episode"])


@pytest.mark.parametrize("problem", ["macro", "hybrid"])
def test_unmodified_inverse_driver_on_gpu_immediate_stepping(problem, tmp_path, dev):
    run_driver(problem, tmp_path, tier="fp32", defer=False)
