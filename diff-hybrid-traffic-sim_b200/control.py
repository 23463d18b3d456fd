"""Signal-controller training loop over the headless ITSCP environment.

``Controller`` and ``Trainer`` keep the reference's names, constructor arguments, method names, optimiser, loss and
on-disk formats (example/control/controller.py:3-35, example/control/trainer.py:14-226):

  * ``<log_path>/eval.txt``: one ``"{:08f}\\n"`` line of minus the mean evaluation reward per evaluation (trainer.py:127-128)
  * ``<log_path>/model.zip`` after every epoch and ``<log_path>/best/model.zip`` on a new best evaluation: ``torch.save`` of
    ``{'controller_state_dict', 'optimizer_state_dict'}`` (trainer.py:90,130-138,204-215)
  * tensorboard scalars ``loss/train`` and ``loss/eval`` (trainer.py:82,125) when tensorboard is importable; the same
    scalars always go to ``<log_path>/scalars.jsonl``

What differs: the episodes of an epoch are ONE batched launch (``ItscpEnv.rollout`` with R rows) instead of R sequential
deep-copied environments, no images are written (headless), and under ``torch.distributed`` every rank runs its own
share of the epoch's episodes and the controller gradients are summed with one all-reduce (``dist.reduce_shared_grads``,
SURVEY 8e: connected networks do not shard, replicas do).
"""
from __future__ import annotations

import gc
import json
import os
from typing import List, Optional

import torch as th

from . import dist
from .itscp_env import Box, ItscpEnv


class Controller(th.nn.Module):
    """MLP emitting one value per signal phase and intersection (controller.py:3-35): Linear-Tanh per hidden layer."""

    def __init__(self, input_size: int, output_size: int, network_size=[256, 256]):
        super().__init__()
        num_layer = len(network_size)
        assert num_layer > 0, ""
        layer = [th.nn.Linear(input_size, network_size[0]), th.nn.Tanh()]
        for i in range(num_layer - 1):
            layer.append(th.nn.Linear(network_size[i], network_size[i + 1]))
            layer.append(th.nn.Tanh())
        layer.append(th.nn.Linear(network_size[-1], output_size))
        self.network = th.nn.Sequential(*layer)

    def forward(self, obs: th.Tensor):
        return self.network(obs)


class _ScalarLog:
    """``add_scalar(tag, value, step)`` into scalars.jsonl and, when available, a tensorboard SummaryWriter."""

    def __init__(self, log_path: str, tensorboard: bool = True):
        self.path = os.path.join(log_path, "scalars.jsonl")
        self.tb = None
        if tensorboard:
            try:
                from torch.utils.tensorboard import SummaryWriter
                self.tb = SummaryWriter(log_path)
            except Exception:      # tensorboard is optional on the GPU box
                self.tb = None

    def add_scalar(self, tag: str, value, step: int):
        value = float(value)
        with open(self.path, "a") as f:
            f.write(json.dumps({"tag": tag, "value": value, "step": int(step)}) + "\n")
        if self.tb is not None:
            self.tb.add_scalar(tag, value, step)

    def close(self):
        if self.tb is not None:
            self.tb.close()


class Trainer:
    """Gradient-based training of a controller on a traffic-control env: minimise minus the sum of episode rewards with
    Adam (trainer.py:14-226)."""

    def __init__(self, env: ItscpEnv, network_size=[256, 256], lr=1e-3, tensorboard: bool = True):
        self.env = env
        input_size = self.env.observation_space.shape
        output_size = self.env.action_space.shape
        assert len(input_size) == 1 and len(output_size) == 1, ""
        self.device = env.device
        self.controller = Controller(input_size[0], output_size[0], network_size).to(self.device)
        self.optimizer = th.optim.Adam(self.controller.parameters(), lr)
        self.best_eval_result = -float("inf")
        self.writer: Optional[_ScalarLog] = None
        self.tensorboard = tensorboard
        self.rank, self.world_size = dist.world()
        if self.world_size > 1:      # every rank starts from rank 0's weights
            for p in self.controller.parameters():
                th.distributed.broadcast(p.data, 0)

    # ------------------------------------------------------------------ loop (trainer.py:39-90)
    def train(self, num_episode_per_epoch: int, num_epoch: int, num_eval_epoch: int, num_eval_episode: int, log_path: str,
              progress: bool = False) -> List[float]:
        if self.rank == 0:
            os.makedirs(log_path, exist_ok=True)
            self.writer = _ScalarLog(log_path, self.tensorboard)
        self.best_eval_result = -float("inf")
        losses = []
        for epoch in range(num_epoch):
            if epoch % max(num_eval_epoch, 1) == 0:
                self.evaluate(epoch, num_eval_episode, log_path)
            self.controller.train(True)
            loss = self.train_epoch(num_episode_per_epoch)
            losses.append(float(loss))
            if self.rank == 0:
                self.writer.add_scalar("loss/train", loss, epoch)
                if progress:
                    print("epoch {}: Loss: {:.6f}".format(epoch, float(loss)), flush=True)
                self.save(log_path + "/model.zip")
            gc.collect()
        if self.writer is not None:
            self.writer.close()
        return losses

    def evaluate(self, epoch: int, num_episode: int, log_path: str):
        """trainer.py:92-138 (hard signals, hard queue test; no rendering)."""
        self.controller.train(False)
        with th.no_grad():
            reward, _, _ = self.run_episodes(max(num_episode, 1), False)
            avg_reward = float(reward.mean())
        if self.world_size > 1:
            avg_reward = float(dist.reduce_losses(th.tensor(avg_reward, device=self.device))[0]) / self.world_size
        if self.rank != 0:
            return avg_reward
        self.writer.add_scalar("loss/eval", -avg_reward, epoch)
        with open(log_path + "/eval.txt", "a") as f:
            f.write("{:08f}\n".format(-avg_reward))
        if avg_reward > self.best_eval_result:
            self.best_eval_result = avg_reward
            os.makedirs(log_path + "/best", exist_ok=True)
            self.save(log_path + "/best/model.zip")
        return avg_reward

    def train_epoch(self, num_episode: int):
        """trainer.py:140-162: loss = -(sum of episode rewards) / num_episode, one Adam step."""
        lo, hi = dist.shard_range(num_episode, self.rank, self.world_size)
        mine = hi - lo
        self.optimizer.zero_grad()
        total = th.zeros((), device=self.device, dtype=self.env.dtype)
        if mine > 0:
            reward, action, _ = self.run_episodes(mine, True)
            total = reward.sum()
            ((-total) / num_episode).backward()
        if self.world_size > 1:
            for p in self.controller.parameters():      # a rank without episodes still takes part in the all-reduce
                if p.grad is None:
                    p.grad = th.zeros_like(p)
            dist.reduce_shared_grads(self.controller.parameters())
            total = dist.reduce_losses(total)[0]
        if mine > 0:
            self.env.flags.check(quiet_collisions=True)
        self.optimizer.step()
        return (-total.detach()) / num_episode

    def policy(self, obs) -> th.Tensor:
        """Action of the current controller for an observation (trainer.py:176-186): sigmoid-squashed into the action box."""
        action = self.controller(th.as_tensor(obs, device=self.device))
        space: Box = self.env.action_space
        low = th.tensor(space.low, device=self.device)
        high = th.tensor(space.high, device=self.device)
        return low + (high - low) * th.sigmoid(action)

    def run_episodes(self, num_episode: int, differentiable: bool):
        """`num_episode` episodes of the env in one batched rollout: same schedule (the reference deep-copies ONE env per
        episode, trainer.py:172), a fresh draw of the spawned vehicles' routes per episode.  Returns (reward [R], action, info)."""
        env = self.env
        action = self.policy(env.observe())
        env.resample_spawn_routes(num_episode)
        reward = env.rollout(action.unsqueeze(0).expand(num_episode, -1), differentiable)
        return reward, action, {"img": [None] * env.num_timestep}

    def run_episode(self, differentiable: bool):
        """trainer.py:164-200: (episode_reward, action, info)."""
        reward, action, info = self.run_episodes(1, differentiable)
        return reward[0], action, info

    # ------------------------------------------------------------------ checkpoints (trainer.py:203-226)
    def save(self, path: str):
        th.save({"controller_state_dict": self.controller.state_dict(),
                 "optimizer_state_dict": self.optimizer.state_dict()}, path)

    def load(self, path: str):
        checkpoint = th.load(path, map_location=self.device)
        self.controller.load_state_dict(checkpoint["controller_state_dict"])
        self.optimizer.load_state_dict(checkpoint["optimizer_state_dict"])
