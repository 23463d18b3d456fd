"""Host-side check (no GPU): the reference's unmodified example/inverse/macro.py driver on top of the drop-in
packages, the four kernel calls swapped for the oracle-backed stand-ins of tests/cpu_standin.py."""
from drivers_cases import run_driver


def test_unmodified_macro_driver_runs_on_the_dropin_host_logic(tmp_path):
    run_driver("macro", tmp_path, extra=["--cpu-standin"])
