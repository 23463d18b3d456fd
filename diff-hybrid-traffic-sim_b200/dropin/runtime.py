"""Device / precision context of the drop-in object API and the scalar <-> vector marshalling.

The reference keeps every cell / vehicle quantity as a Python float or a 0-dim tensor
(road/lane/_macro_lane.py:30-44, road/vehicle/vehicle.py:1-17) and rebuilds per-step
vectors element by element (dmacro_lane.py:134-158, dmicro_lane.py:130-153).  Here the
vectors are the primary storage (on the GPU); the objects hold 0-dim VIEWS of them, and a
lane re-uses its vector whenever nobody replaced those views (identity check), falling
back to an autograd-preserving gather otherwise.
"""
from __future__ import annotations

import os

import torch

_PRECISIONS = ("mixed", "float64", "float32")
_cfg = {"precision": os.environ.get("DHTS_PRECISION", "mixed"), "device": os.environ.get("DHTS_DEVICE"),
        "defer": os.environ.get("DHTS_DEFER", "1") != "0", "defer_hyb": os.environ.get("DHTS_DEFER_HYB", "1") != "0"}
_flags = {}


def configure(precision=None, device=None, defer=None, defer_hyb=None):
    """defer: queue RoadNetwork.forward calls and run them as fused rollouts (dropin/deferred.py); defer_hyb: also for
    connected macro / micro networks."""
    if defer is not None:
        _cfg["defer"] = bool(defer)
    if defer_hyb is not None:
        _cfg["defer_hyb"] = bool(defer_hyb)
    if precision is not None:
        if precision not in _PRECISIONS:
            raise ValueError("precision must be one of %s" % (_PRECISIONS,))
        _cfg["precision"] = precision
    if device is not None:
        _cfg["device"] = str(device)


def defer_enabled(kind=None) -> bool:
    return bool(_cfg["defer"] and (kind != "hyb" or _cfg["defer_hyb"]))


def device() -> torch.device:
    """The CUDA device the lanes live on.  Fails loudly without one: the step has no CPU path."""
    if not _cfg.get("cuda_ok"):                 # asked once per process: the call is on the per-step path
        if not torch.cuda.is_available():
            raise RuntimeError("dhts drop-in lanes need a CUDA device: the simulation step runs only in the sm_100a "
                               "kernels of libdhts_b200.so (no CPU fallback)")
        _cfg["cuda_ok"] = True
    d = _cfg["device"]
    return torch.device(d) if d else torch.device("cuda", torch.cuda.current_device())


def store_dtype() -> torch.dtype:
    """dtype of the state kept between steps (what get_state_vector returns)."""
    return torch.float64 if _cfg["precision"] == "float64" else torch.float32


def step_dtype() -> torch.dtype:
    """dtype the step kernels run in."""
    return torch.float32 if _cfg["precision"] == "float32" else torch.float64


def flags():
    from dhts_b200 import _lib
    dev = device()
    if dev not in _flags:
        _flags[dev] = _lib.Flags(dev)
    return _flags[dev]


def check_flags():
    """Maps the device-side conditions to the reference's conventions (CFL / NaN-gradient
    AssertionError, collision print-and-continue) and clears them."""
    for f in _flags.values():
        try:
            f.check()
        finally:
            f.reset()


def is_tensor(x) -> bool:
    return isinstance(x, torch.Tensor)


def scalar(x, dtype=None):
    """float / 0-dim tensor (any device) -> 0-dim tensor on the lane device, autograd edge kept."""
    dtype = dtype or store_dtype()
    if is_tensor(x):
        return x.reshape(()).to(device=device(), dtype=dtype)
    return torch.tensor(float(x), dtype=dtype, device=device())


def vector(x, dtype=None):
    """sequence / tensor -> 1-D tensor on the lane device (autograd edge kept for tensors)."""
    dtype = dtype or store_dtype()
    if is_tensor(x):
        return x.reshape(-1).to(device=device(), dtype=dtype)
    return gather(list(x), dtype)


def gather(vals, dtype=None):
    """List of floats / 0-dim tensors on any device -> 1-D device tensor, keeping autograd edges.
    Floats travel in one host->device copy, tensors in one stack per source device."""
    dtype = dtype or store_dtype()
    dev = device()
    n = len(vals)
    groups = {}
    fidx, fval = [], []
    for i, v in enumerate(vals):
        if is_tensor(v):
            groups.setdefault(v.device, ([], []))
            groups[v.device][0].append(i)
            groups[v.device][1].append(v.reshape(()))
        else:
            fidx.append(i)
            fval.append(float(v))
    if len(fidx) == n:
        return torch.tensor(fval, dtype=dtype).to(dev)
    if len(groups) == 1 and not fidx:
        (idx, ts), = groups.values()
        return torch.stack(ts).to(device=dev, dtype=dtype)
    out = torch.zeros(n, dtype=dtype, device=dev)
    if fidx:
        out = out.index_copy(0, torch.tensor(fidx, device=dev), torch.tensor(fval, dtype=dtype).to(dev))
    for idx, ts in groups.values():
        out = out.index_copy(0, torch.tensor(idx, device=dev), torch.stack(ts).to(device=dev, dtype=dtype))
    return out


class Views:
    """A device vector together with the 0-dim views handed out to the objects."""
    __slots__ = ("vec", "items")

    def __init__(self, vec: torch.Tensor):
        self.vec = vec
        self.items = vec.unbind(0) if vec.numel() else ()

    def still(self, current) -> bool:
        """True iff `current` (an iterable of the objects' attribute values) are exactly our views."""
        items = self.items
        n = 0
        for a in current:
            if n >= len(items) or a is not items[n]:
                return False
            n += 1
        return n == len(items)
