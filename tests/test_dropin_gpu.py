"""Drop-in object API on the B200: the cases of tests/dropin_cases.py with the REAL kernels (libdhts_b200.so)."""
import pytest

import dropin_cases as C

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("tier,precision,dtype,tol_s,tol_g", C.TIERS)
@pytest.mark.parametrize("mode", ["macro", "hybrid"])
def test_inverse_macro_and_hybrid_loss_curves(dev, tier, precision, dtype, tol_s, tol_g, mode):
    C.case_inverse_macro_and_hybrid_loss_curves(tier, precision, dtype, tol_s, tol_g, mode)


@pytest.mark.parametrize("tier,precision,dtype,tol_s,tol_g", C.TIERS)
def test_inverse_micro_loss_curve(dev, tier, precision, dtype, tol_s, tol_g):
    C.case_inverse_micro_loss_curve(tier, precision, dtype, tol_s, tol_g)


@pytest.mark.parametrize("tier,precision,dtype,tol_s,tol_g", C.TIERS)
def test_hybrid_chain_spawn_absorb_and_gradients(dev, tier, precision, dtype, tol_s, tol_g):
    C.case_hybrid_chain_spawn_absorb_and_gradients(tier, precision, dtype, tol_s, tol_g)


@pytest.mark.parametrize("tier,precision,dtype,tol_s,tol_g", C.TIERS)
def test_macro_lane_rollout_and_ghost_gradients(dev, tier, precision, dtype, tol_s, tol_g):
    C.case_macro_lane_rollout_and_ghost_gradients(tier, precision, dtype, tol_s, tol_g)


def test_object_surface_on_device(dev):
    C.case_object_surface_on_device()


def test_cfl_violation_raises_like_the_reference(dev):
    C.case_cfl_violation_raises_like_the_reference()


def test_all_float32_precision_runs(dev):
    """precision="float32": fp32 kernels end to end; vs the fp32 reference only to fp32-arithmetic accuracy."""
    import torch as th
    C.case_inverse_macro_and_hybrid_loss_curves("fp32", "float32", th.float32, 2e-3, 5e-2, "macro")
