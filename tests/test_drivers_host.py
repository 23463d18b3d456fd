"""Host-side check (no GPU): the reference's unmodified example/inverse drivers on top of the drop-in packages, the
kernel calls swapped for the oracle-backed stand-ins of tests/cpu_standin.py -- with the deferred stepping of
dropin/deferred.py (500 queued RoadNetwork.forward calls run as one rollout) and with immediate stepping."""
import pytest

from drivers_cases import run_driver


@pytest.mark.parametrize("problem,tier,defer", [("macro", "fp64", True), ("micro", "fp64", True), ("macro", "fp32", False)])
def test_unmodified_driver_runs_on_the_dropin_host_logic(problem, tier, defer, tmp_path):
    run_driver(problem, tmp_path, tier=tier, defer=defer, extra=["--cpu-standin"])
