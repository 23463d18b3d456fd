"""``python -m dhts_b200.run_itscp``: the reference's ``example/control/itscp/run.py`` (same arguments, same result
tree ``./result/control/itscp/<mode>_<time>/trial_<k>/``) over the headless env and the fused kernels.

    python -m dhts_b200.run_itscp --mode=hybrid --problem=1 --n_trial=1 --n_intersection=3 --n_lane=1 --lane_length=5 \
        --speed_limit=60 --simulation_length=20 --signal_length=4 --n_episode=100 --lr=1e-4        # run_itscp_hybrid.sh, line 1

Extra arguments: ``--episodes_per_epoch`` (the reference hard-codes 1, run.py:70), ``--seed``, ``--out``.  Under
torchrun the episodes of an epoch are split over the ranks (one process per GPU) and the controller gradients
all-reduced.
"""
import argparse
import os
from time import time

import numpy as np
import torch

from .control import Trainer
from .itscp_env import ItscpEnv, problem_1, problem_2, problem_3


def main(argv=None):
    parser = argparse.ArgumentParser("Script to solve intersection signal control problem")
    parser.add_argument("--mode", type=str, choices=["macro", "micro", "hybrid"], default="macro")
    parser.add_argument("--problem", type=int, choices=[1, 2, 3], default=1)
    parser.add_argument("--n_trial", type=int, default=5)
    parser.add_argument("--n_intersection", type=int, default=1)
    parser.add_argument("--n_lane", type=int, default=3)
    parser.add_argument("--lane_length", type=float, default=20.)
    parser.add_argument("--speed_limit", type=float, default=60.)
    parser.add_argument("--simulation_length", type=int, default=10)
    parser.add_argument("--signal_length", type=int, default=2)
    parser.add_argument("--n_episode", type=int, default=200)
    parser.add_argument("--lr", type=float, default=1e-3)
    parser.add_argument("--episodes_per_epoch", type=int, default=1)
    parser.add_argument("--seed", type=int, default=0)
    parser.add_argument("--out", type=str, default=None)
    args = parser.parse_args(argv)

    ws = int(os.environ.get("WORLD_SIZE", "1"))
    if ws > 1 and not torch.distributed.is_initialized():
        torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", "0")))
        torch.distributed.init_process_group("nccl")
    rank = int(os.environ.get("RANK", "0"))
    # Every rank must solve the SAME problem (inflow schedule, per-frame macro routes: both drawn from np.random in
    # env.reset()) and write into the SAME result tree, or the all-reduced controller gradients / averaged evaluation
    # rewards would mix different problems.  With --seed=0 (the reference default: unseeded) rank 0 draws a seed and a
    # run name and broadcasts them.
    seed, stamp = args.seed, int(time())
    if ws > 1:
        box = torch.tensor([seed if seed > 0 else int(np.random.randint(1, 2 ** 31 - 1)), stamp], dtype=torch.int64,
                           device="cuda")
        torch.distributed.broadcast(box, src=0)
        seed, stamp = int(box[0]), int(box[1])
    if seed > 0:
        np.random.seed(seed); torch.manual_seed(seed)
    problem = {1: problem_1, 2: problem_2, 3: problem_3}[args.problem]
    run_name = args.out or "./result/control/itscp/{}_{}".format(args.mode, stamp)
    if rank == 0:
        os.makedirs(run_name, exist_ok=True)
    if ws > 1:
        torch.distributed.barrier()

    env = ItscpEnv()
    env.schedule_callback = problem
    env.config["num_intersection"] = args.n_intersection
    env.config["lane_length"] = args.lane_length
    env.config["num_lane"] = args.n_lane
    env.config["render"] = False
    env.config["policy_length"] = args.simulation_length
    env.config["signal_length"] = args.signal_length
    env.config["mode"] = args.mode
    env.config["speed_limit"] = args.speed_limit
    env.config["random_seed"] = seed          # > 0 under torchrun: env.reset() re-seeds np.random identically on every rank
    env.reset()
    curves = []
    for it in range(args.n_trial):
        log_path = run_name + "/trial_{}".format(it)
        trainer = Trainer(env, lr=args.lr)
        curves.append(trainer.train(args.episodes_per_epoch, args.n_episode + 1, max(args.n_episode // 10, 1), 1, log_path,
                                    progress=True))
    return curves


if __name__ == "__main__":
    main()
