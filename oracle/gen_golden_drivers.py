"""TEST INFRASTRUCTURE: freeze the first gradient-descent episodes of the reference's UNMODIFIED inverse-problem
drivers (example/inverse/{macro,micro,hybrid}.py on the reference's own CPU lanes) into tests/golden/drivers_fp32.npz
(the reference as shipped: fp32 state) and drivers_fp64.npz (the same files run in float64 through the no-edit dtype
rebinding of SURVEY App. C).  Run in THIS container (needs baseline/_ref, see baseline/install_ref.py):

    python oracle/gen_golden_drivers.py

tests/test_drivers_*.py run the same drivers, same seed, on top of the drop-in packages and compare the curves.
"""
import json
import os
import subprocess
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EPISODES, SEED = 3, 20221008


def main():
    for tier, extra in (("fp32", []), ("fp64", ["--fp64"])):
        gen(tier, extra)


def gen(tier, extra):
    out = {}
    for prob in ("macro", "micro", "hybrid"):
        with tempfile.TemporaryDirectory() as d:
            f = os.path.join(d, "o.npz")
            r = subprocess.run([sys.executable, os.path.join(ROOT, "baseline", "run_drivers.py"), "--impl", "reference",
                                "--problem", prob, "--episodes", str(EPISODES), "--seed", str(SEED), "--out", f] + extra,
                               capture_output=True, text=True, check=True)
            line = json.loads(r.stdout.strip().splitlines()[-1])
            assert line["core_packages_from"].endswith("baseline/_ref"), line
            z = np.load(f)
            for k in z.files:
                out[prob + "_" + k] = z[k]
            print(prob, line["end_errors"], "%.2f s/episode" % line["s_per_episode"])
    np.savez(os.path.join(ROOT, "tests", "golden", "drivers_%s.npz" % tier), episodes=EPISODES, seed=SEED, **out)


if __name__ == "__main__":
    main()
