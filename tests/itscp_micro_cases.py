"""Shared pieces of the ITSCP MICRO-mode tests: episodes frozen from the live reference (oracle/gen_golden_micro.py:
run_itscp_micro.sh's configuration and a 2 x 2 grid) and envs configured the way run.py configures the reference's."""
import numpy as np

from conftest import golden

CASES = {"m": dict(num_intersection=1, num_lane=3, lane_length=30.0, policy_length=10, signal_length=2, problem=1),
         "n": dict(num_intersection=2, num_lane=1, lane_length=20.0, policy_length=8, signal_length=2, problem=2)}


def micro_fixture(tag):
    g = golden("itscp_micro_fp64")
    return {k[2:]: g[k] for k in g.files if k.startswith(tag + "_")}


def micro_env(tag, G, device):
    """example/control/itscp/run.py:48-59 with run_itscp_micro.sh's arguments, seeded like the fixture's generator: the
    env's own reset must then reproduce the reference's schedule, waiting routes and np.random state."""
    from dhts_b200 import itscp_env as E
    c = CASES[tag]
    env = E.ItscpEnv(device=device)
    env.schedule_callback = {1: E.problem_1, 2: E.problem_2, 3: E.problem_3}[c["problem"]]
    env.config.update(num_intersection=c["num_intersection"], lane_length=c["lane_length"], num_lane=c["num_lane"], render=False,
                      policy_length=c["policy_length"], signal_length=c["signal_length"], mode="micro", speed_limit=60.0,
                      random_seed=int(G["seed"]))
    env.reset()
    return env
