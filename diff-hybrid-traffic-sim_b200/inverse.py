"""Population-parallel inverse problems: the three problems of ``example/inverse/`` with every
trial -- and every candidate of a gradient-free optimiser's population -- evaluated as one batch on the fused rollouts.

The reference runs ``num_trial`` trials one after another, each a loop of ``num_episode`` simulations of ``num_timestep``
steps (example/inverse/_inverse.py:99-172,187-242), and evaluates CMA-ES populations one candidate at a time
(:268-292).  Trials and candidates never interact, so here they are rows of ONE rollout:

  * ``MacroInverseBatch``   example/inverse/macro.py   -> ``functional.arz_rollout``      (P lanes)
  * ``MicroInverseBatch``   example/inverse/micro.py   -> ``functional.idm_rollout``      (P lanes, CSR)
  * ``HybridInverseBatch``  example/inverse/hybrid.py  -> ``hybrid_network.hybrid_rollout`` (P replicas, plain mode)

Problem data is drawn exactly as the reference draws it (CPU fp32 ``torch.rand`` in the same order per trial), the error
is the reference's (sum of squared differences of the two state vectors), Adam + projection onto the bounds is
``solve_gd`` (:187-242) applied to the whole batch -- Adam is element-wise, so a batched run equals P independent ones --
and results are written in the reference's two-column ``trial_<k>.txt`` format (``log_error``, :504-514).
``evaluate_vector_states`` is the batched form of ``evaluate_vector_state`` (:440-458) for ask/tell optimisers (``cma`` is
not a dependency here).
"""
from __future__ import annotations

import os
from typing import List, Optional, Sequence, Tuple

import torch as th

from . import _lib
from . import functional as F
from .hybrid_network import HybridNetTopology, default_vehicle_params, hybrid_rollout
from .network import MODE_PLAIN

State = Tuple[th.Tensor, th.Tensor]


def log_error(path: str, beg_errors: Sequence[float], end_errors: Sequence[float]):
    """_inverse.py:504-514: one ``"<beg> <end>"`` line per episode."""
    with open(path, "w") as f:
        for b, e in zip(beg_errors, end_errors):
            f.write("{} {}\n".format(b, e))


class InverseBatch:
    """Estimate the initial traffic state that ends in a given final state (_inverse.py:16-30), P trials at once.
    States are pairs of [P, n] tensors on the device."""

    method_dir = {"gd": "gd", "cma": "cma-es", "nm": "nelder-mead", "slsqp": "slsqp"}      # _inverse.py:124-160

    def __init__(self, num_trial: int, num_timestep: int, num_episode: int, delta_time: float, speed_limit: float,
                 run_name: str, device=None, dtype=th.float64, log_root: str = "result/inverse"):
        self.num_trial, self.num_timestep, self.num_episode = int(num_trial), int(num_timestep), int(num_episode)
        self.delta_time, self.speed_limit = float(delta_time), float(speed_limit)
        if device is None:
            if not th.cuda.is_available():
                raise RuntimeError("the inverse problems step CUDA kernels only (no CPU fallback)")
            device = th.device("cuda", th.cuda.current_device())
        self.device, self.dtype = th.device(device), dtype
        self.beg_state: Optional[State] = None
        self.end_state: Optional[State] = None
        self.gd_lr = 1e-3                       # _inverse.py:56
        self.log_dir = os.path.join(log_root, run_name)
        self.flags = _lib.Flags(self.device) if self.device.type == "cuda" else None

    # -- problem-specific ------------------------------------------------------------------------------------------
    def init_network(self, trial: int):
        raise NotImplementedError

    def random_initial_state(self, trial: int) -> State:
        raise NotImplementedError

    def simulate(self, state: State, differentiable: bool, rows: Optional[th.Tensor] = None) -> State:
        """End state of `num_timestep` steps from `state`; ``rows[i]`` names the trial whose network row i is simulated on
        (default: row i = trial i)."""
        raise NotImplementedError

    def bounds(self) -> Tuple[State, State]:
        raise NotImplementedError

    # -- shared ----------------------------------------------------------------------------------------------------
    def _dev(self, x: th.Tensor) -> th.Tensor:
        return x.to(device=self.device, dtype=self.dtype)

    def initialize(self):
        """_inverse.py:68-88 for every trial, in trial order (same RNG stream as the reference's sequential trials when
        only the GD solver runs): network, true initial state, estimate; then ONE batched simulation for the targets."""
        beg, est = [], []
        # Quirk kept: the reference never resets ``beg_state`` between trials, so from trial 1 on ``random_initial_state``
        # perturbs the PREVIOUS trial's truth instead of drawing afresh (macro.py:79-96 with _inverse.py:80-84,113-121).
        self._trial_beg = None
        for k in range(self.num_trial):
            self.init_network(k)
            b = self.random_initial_state(k)
            self._trial_beg = b
            beg.append(b)
            est.append(self.random_initial_state(k))
        self.beg_state = tuple(self._dev(th.stack([s[i] for s in beg])) for i in range(2))
        self.initial_estimate = tuple(self._dev(th.stack([s[i] for s in est])) for i in range(2))
        self._finish_networks()
        with th.no_grad():
            self.end_state = tuple(x.detach() for x in self.simulate(self.beg_state, False))
        return self.initial_estimate

    def _finish_networks(self):
        pass

    @staticmethod
    def compute_error(sa: State, sb: State) -> th.Tensor:
        """macro.py:206-241 / micro.py:214-232 per trial: [P]."""
        return th.pow(sa[0] - sb[0], 2.0).sum(-1) + th.pow(sa[1] - sb[1], 2.0).sum(-1)

    def solve_gd(self, est: Optional[State] = None, lr: Optional[float] = None):
        """_inverse.py:187-242 on all trials at once.  Returns (beg_errors, end_errors), each [num_episode][P] lists."""
        est = self.initial_estimate if est is None else est
        lr = self.gd_lr if lr is None else lr
        x = tuple(self._dev(s).detach().clone().requires_grad_() for s in est)
        opt = th.optim.Adam(x, lr=lr)
        lb, ub = self.bounds()
        beg_errors, end_errors = [], []
        for _ in range(self.num_episode):
            self.flags.reset()
            end = self.simulate(x, True)
            beg_error = self.compute_error(self.beg_state, x)
            end_error = self.compute_error(self.end_state, end)
            beg_errors.append(beg_error.detach().cpu().tolist())
            end_errors.append(end_error.detach().cpu().tolist())
            opt.zero_grad()
            end_error.sum().backward()           # trials are independent: d(sum)/d(row p) = d(error_p)/d(row p)
            self.flags.check(quiet_collisions=True)
            opt.step()
            with th.no_grad():
                for i in range(2):
                    x[i].copy_(th.min(th.max(x[i], lb[i]), ub[i]))
        self.estimate = tuple(s.detach() for s in x)
        return beg_errors, end_errors

    def vectorize(self, state: State) -> th.Tensor:
        return th.cat([state[0], state[1]], dim=-1)

    def unvectorize(self, v: th.Tensor) -> State:
        n = v.shape[-1] // 2
        return v[..., :n], v[..., n:]

    def evaluate_vector_states(self, vstates: th.Tensor, trial) -> Tuple[th.Tensor, th.Tensor]:
        """Batched _inverse.py:440-458: M candidate vectors [M, 2n] (a CMA-ES population, a simplex, finite-difference
        probes), each evaluated on the network of trial ``trial`` (an int or an [M] tensor).  Returns (beg_error, end_error) [M]."""
        v = self._dev(th.as_tensor(vstates))
        M = v.shape[0]
        rows = th.full((M,), int(trial), dtype=th.long, device=self.device) if isinstance(trial, int) else \
            th.as_tensor(trial, dtype=th.long, device=self.device)
        st = self.unvectorize(v)
        with th.no_grad():
            self.flags.reset()
            end = self.simulate(st, False, rows)
            self.flags.check(quiet_collisions=True)
            pick = lambda s: tuple(x[rows] for x in s)
            return self.compute_error(pick(self.beg_state), st), self.compute_error(pick(self.end_state), end)

    def solve_scipy(self, est_row: State, method: str, trial: int):
        """_inverse.py:301-352 for ONE trial: Nelder-Mead / SLSQP from scipy over ``evaluate_vector_states``; every objective
        call is logged, the lists are padded / cut to ``num_episode`` as the reference does.  SLSQP's finite-difference
        gradient (2n probes) is evaluated as ONE batched launch instead of 2n sequential simulations."""
        import numpy as np
        import scipy.optimize
        lb, ub = self.bounds()
        bounds = scipy.optimize.Bounds(th.cat([lb[0], lb[1]]).cpu().numpy(), th.cat([ub[0], ub[1]]).cpu().numpy())
        x0 = self.vectorize(tuple(self._dev(s) for s in est_row)).cpu().numpy()
        beg_errors: List[float] = []
        end_errors: List[float] = []

        def fun(v):
            b, e = self.evaluate_vector_states(th.as_tensor(np.asarray(v)[None]), trial)
            beg_errors.append(float(b[0])); end_errors.append(float(e[0]))
            return float(e[0])

        def jac(v):      # forward differences with scipy's default relative step, all probes in one batch
            v = np.asarray(v, dtype=np.float64)
            h = np.sqrt(np.finfo(np.float64).eps) * np.maximum(1.0, np.abs(v))
            probes = np.concatenate([v[None], v[None] + np.diag(h)])
            _, e = self.evaluate_vector_states(th.as_tensor(probes), trial)
            e = e.cpu().numpy()
            return (e[1:] - e[0]) / h

        kw = dict(fun=fun, x0=x0, bounds=bounds, options={"maxiter": self.num_episode + 1}, method=method)
        if method == "SLSQP":
            kw["jac"] = jac
        scipy.optimize.minimize(**kw)
        while len(beg_errors) < self.num_episode:
            beg_errors.append(beg_errors[-1]); end_errors.append(end_errors[-1])
        return beg_errors[:self.num_episode], end_errors[:self.num_episode]

    def solve_cma(self, est_row: State, sigma: float, trial: int):
        """_inverse.py:245-299 for ONE trial: ask / tell CMA-ES (the optional ``cma`` package) with every population
        evaluated as one batch."""
        from math import ceil
        import cma      # not a dependency of this repository: pip install cma
        lb, ub = self.bounds()
        opts = cma.CMAOptions(); opts.set("bounds", [th.cat([lb[0], lb[1]]).cpu().tolist(), th.cat([ub[0], ub[1]]).cpu().tolist()])
        opts.set("verbose", -1)
        es = cma.CMAEvolutionStrategy(self.vectorize(tuple(self._dev(s) for s in est_row)).cpu().numpy(), sigma, opts)
        beg_errors: List[float] = []
        end_errors: List[float] = []
        for _ in range(ceil(self.num_episode / es.popsize)):
            sol = es.ask()
            b, e = self.evaluate_vector_states(th.as_tensor(sol), trial)
            beg_errors.extend(b.cpu().tolist()); end_errors.extend(e.cpu().tolist())
            es.tell(sol, e.cpu().tolist())
        return beg_errors[:self.num_episode], end_errors[:self.num_episode]

    def write_trials(self, method: str, beg_errors, end_errors) -> List[str]:
        """``<log_dir>/<method dir>/trial_<k>.txt`` for every trial (_inverse.py:162-166)."""
        d = os.path.join(self.log_dir, self.method_dir.get(method, method))
        os.makedirs(d, exist_ok=True)
        paths = []
        for k in range(self.num_trial):
            p = os.path.join(d, "trial_{}.txt".format(k))
            log_error(p, [row[k] for row in beg_errors], [row[k] for row in end_errors])
            paths.append(p)
        return paths


class MacroInverseBatch(InverseBatch):
    """example/inverse/macro.py: one dMacroLane of `num_cell` cells with random static ghost cells per trial."""

    def __init__(self, num_trial, num_timestep, num_episode, delta_time, speed_limit, run_name, num_cell: int,
                 cell_length: float, **kw):
        super().__init__(num_trial, num_timestep, num_episode, delta_time, speed_limit, run_name, **kw)
        self.num_cell, self.cell_length = int(num_cell), float(cell_length)
        self.draw_dtype = th.float32                     # macro.py:33
        self._bd, self._bs = [], []

    def init_network(self, trial):                       # macro.py:35-66
        self._bd.append(th.rand((2,), dtype=self.draw_dtype))
        self._bs.append(th.rand((2,), dtype=self.draw_dtype) * self.speed_limit)
        self.random_initial_state(trial)                 # init_density / init_speed, overwritten by initialize()

    def random_initial_state(self, trial):               # macro.py:68-98
        n, dt = self.num_cell, self.draw_dtype
        if self._trial_beg is None:
            sl = th.tensor([self.speed_limit], dtype=dt)
            return th.rand((n,), dtype=dt), th.rand((n,), dtype=dt) * sl
        r = self._trial_beg[0] + th.randn((n,), dtype=dt) * 1e-2
        u = self._trial_beg[1] + th.randn((n,), dtype=dt) * 1e-2
        return th.clamp(r, min=0., max=1.), th.clamp(u, min=0., max=self.speed_limit)

    def _finish_networks(self):
        self.ghost_r = self._dev(th.stack(self._bd)); self.ghost_u = self._dev(th.stack(self._bs))

    def simulate(self, state, differentiable, rows=None):
        gr, gu = (self.ghost_r, self.ghost_u) if rows is None else (self.ghost_r[rows], self.ghost_u[rows])
        rT, _, uT = F.arz_rollout(state[0], state[1], gr, gu, self.cell_length, self.speed_limit, self.delta_time,
                                  self.num_timestep, ckpt_every=1 if differentiable else 32, flags=self.flags)
        return rT, uT                                    # get_state: density and speed (macro.py:109-121)

    def bounds(self):                                    # macro.py:182-204
        z = th.zeros((self.num_cell,), dtype=self.dtype, device=self.device)
        return (z, z), (z + 1.0, z + self.speed_limit)


class MicroInverseBatch(InverseBatch):
    """example/inverse/micro.py: `num_vehicle` default vehicles on one (endless) dMicroLane per trial."""

    def __init__(self, num_trial, num_timestep, num_episode, delta_time, speed_limit, run_name, num_vehicle: int,
                 vehicle_length: float, **kw):
        super().__init__(num_trial, num_timestep, num_episode, delta_time, speed_limit, run_name, **kw)
        self.num_vehicle, self.vehicle_length = int(num_vehicle), float(vehicle_length)
        self.draw_dtype = th.float32
        self.gd_lr = 1e-3

    def init_network(self, trial):                       # micro.py:37-58
        self.random_initial_state(trial)

    def random_initial_state(self, trial):               # micro.py:60-92
        n, L, dt = self.num_vehicle, self.vehicle_length, self.draw_dtype
        sl = th.tensor([self.speed_limit], dtype=dt)
        if self._trial_beg is None:
            start = th.arange(0, n) * 4.0 * L
            p = start + th.rand((n,), dtype=dt) * 2.0 * L
            v = th.lerp(0.3 * sl, 0.7 * sl, th.rand((n,), dtype=dt))
            return p, v
        p = self._trial_beg[0] + th.randn((n,), dtype=dt) * 0.1 * L
        v = self._trial_beg[1] + th.randn((n,), dtype=dt) * 1e-2 * sl
        lb, ub = self._cpu_bounds()
        return th.clamp(p, min=lb[0], max=ub[0]), th.clamp(v, min=lb[1], max=ub[1])

    def _cpu_bounds(self):                               # micro.py:196-212
        n, L = self.num_vehicle, self.vehicle_length
        plb = (th.arange(0, n) * 4.0 * L).to(self.draw_dtype)
        return (plb, th.zeros(n)), (plb + 2.0 * L, th.ones(n) * self.speed_limit)

    def bounds(self):
        lb, ub = self._cpu_bounds()
        return tuple(self._dev(x) for x in lb), tuple(self._dev(x) for x in ub)

    def simulate(self, state, differentiable, rows=None):
        P, n = state[0].shape
        par = th.tensor(default_vehicle_params(self.speed_limit, self.vehicle_length), dtype=self.dtype, device=self.device)
        params = par[:, None].expand(6, P * n).contiguous()                      # default_micro_vehicle, micro_vehicle.py:30-72
        off = th.arange(P + 1, dtype=th.int32, device=self.device) * n
        head = th.tensor([1000.0, 0.0], dtype=self.dtype, device=self.device).expand(P, 2).contiguous()   # _micro_lane.py:14-15
        pT, vT = F.idm_rollout(state[0].reshape(-1), state[1].reshape(-1), params, off, head, self.delta_time,
                               self.num_timestep, flags=self.flags, max_lane=n)
        return pT.reshape(P, n), vT.reshape(P, n)


class HybridInverseBatch(MacroInverseBatch):
    """example/inverse/hybrid.py: macro lane -> micro lane -> macro lane; the state estimated and the error measured are
    lane 0's (:108-131), lanes 1 and 2 start empty."""

    def init_network(self, trial):                       # hybrid.py:37-74
        self._bd.append(th.rand((4,), dtype=self.draw_dtype))
        self._bs.append(th.rand((4,), dtype=self.draw_dtype) * self.speed_limit)
        self.random_initial_state(trial)

    def _finish_networks(self):
        N, dx = self.num_cell, self.cell_length
        L = N * dx
        self.topo = HybridNetTopology([0, 1, 0], [N, 0, N], [dx, 1.0, dx], [L, L, L], [(0, 1), (1, 2)], self.device,
                                      MODE_PLAIN, veh_cap=max(4, int(L // 5) + 4))
        bd, bs = th.stack(self._bd), th.stack(self._bs)                          # [P, 4]: l0 left, l0 right, l2 left, l2 right
        own = th.stack([th.stack([bd[:, i], bs[:, i]], -1) for i in (0, 2, 1, 3)], 1)      # own slots: left sides, then right sides
        self.own0 = self._dev(own)
        T = self.num_timestep
        self.route = th.tensor([[[-1, 0, -1], [1, -1, -1]]] * T, dtype=th.int32, device=self.device)   # the only MacroRoute: 0 -> 1
        self.spawn = th.zeros((1, 4 * (int(self.speed_limit * T * self.delta_time / 5.0) + 2)), dtype=th.int32, device=self.device)

    def simulate(self, state, differentiable, rows=None):
        P, N = state[0].shape
        own0 = self.own0 if rows is None else self.own0[rows]
        z = th.zeros((P, N), dtype=self.dtype, device=self.device)
        r0 = th.cat([state[0], z], 1); u0 = th.cat([state[1], z + self.speed_limit], 1)     # lane 2 cleared: (0, u_max)
        st = hybrid_rollout(self.topo, r0, u0, self.speed_limit, self.delta_time, self.num_timestep, route=self.route,
                            spawn_route=self.spawn, own0=own0.contiguous(), soft=True, flags=self.flags)
        c = st.cells[self.num_timestep]
        return c[:, 0, :N], c[:, 2, :N]
