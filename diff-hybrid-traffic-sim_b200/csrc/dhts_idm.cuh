// IDM device math shared by the lane kernels (idm_kernels.cu) and the hybrid network rollout (net_hybrid.cu).
//
// Behaviour restated from the reference (file:line relative to its checkout):
//   model/micro/_idm.py:5-51            acceleration with the two clips
//   road/lane/_micro_lane.py:149-168    collision reset and the 1e-5 floor of the gap
//   model/micro/didm.py:12-102          ego / leader 2x2 Jacobians
#pragma once
#include "dhts_arz.cuh"

namespace dhts {

template <typename T> struct IdmPar { T a_max, v_t_inv, s0, tp, len, sab2_inv; };
//  v_t_inv = 1/target_speed, sab2_inv = 1/(2 sqrt(a_max a_pref)): per-vehicle constants of _idm.py:31-40

template <typename T> __device__ __forceinline__ IdmPar<T> load_par(const T* __restrict__ params, size_t V, size_t i) {
    IdmPar<T> k;
    T a_max = params[i], a_pref = params[V + i], v_t = params[2 * V + i];
    k.a_max = a_max; k.v_t_inv = T(1) / v_t; k.s0 = params[3 * V + i]; k.tp = params[4 * V + i];
    k.len = params[5 * V + i];
    k.sab2_inv = T(1) / (T(2) * t_sqrt(a_max * a_pref));
    return k;
}

template <typename T> struct IdmEval { T acc, s; bool clip_acc, clip_s, col; };

// _micro_lane.py:149-168 (collision reset, 1e-5 floor) + _idm.py:31-51
template <typename T>
__device__ __forceinline__ IdmEval<T> idm_eval(T v, const IdmPar<T>& k, T dp_raw, T dv_raw, T inv_dt) {
    IdmEval<T> e;
    e.col = dp_raw < T(0);
    T dp = e.col ? T(0) : dp_raw, dv = e.col ? T(0) : dv_raw;
    dp = t_max(dp, T(1e-5));
    T s = k.s0 + v * k.tp + (v * dv) * k.sab2_inv;
    e.clip_s = s < T(0);
    s = t_max(s, T(0));
    T q = v * k.v_t_inv; q = q * q;
    T sr = s * f_rcp(dp);              // dp >= 1e-5: branch-free reciprocal (<= 1 ulp) instead of the IEEE division's slow-path call
    T acc = k.a_max * (T(1) - q * q - sr * sr);
    T lim = -v * inv_dt;
    e.clip_acc = acc < lim;
    e.acc = e.clip_acc ? lim : acc;
    e.s = s;
    return e;
}

// didm.py:12-102 with the RAW deltas (dmicro_lane.py:97).  Returns the second
// rows (E10, E11) and (L10, L11); first rows are [1, dt] and [0, 0].
template <typename T>
__device__ __forceinline__ void idm_jac(T v, const IdmPar<T>& k, T dp_raw, T dv_raw, const IdmEval<T>& e, T dt,
                                        T& E10, T& E11, T& L10, T& L11) {
    // Branch-free: the two clips differ per vehicle, so branches here diverge in every warp.  With s* clipped e.s is 0, hence
    // s / dp^2 = 0 and the general expressions reduce to the clipped ones of didm.py:38-56,84-102 (E11 = 1 + dt a_max t1, L11 = 0);
    // the acceleration clip zeroes both second rows.
    T idp = f_rcp(dp_raw);             // raw gap (dmicro_lane.py:97); a zero gap gives a non-finite Jacobian (the reference raises), flagged as NaN gradient
    T sd2 = e.s * idp * idp;          // s / dp^2
    T sd3 = e.s * sd2 * idp;          // s^2 / dp^3
    T l10 = dt * (T(2) * k.a_max * sd3);
    T vt2 = k.v_t_inv * k.v_t_inv;
    T t1 = T(-4) * (v * v * v) * (vt2 * vt2);
    T e11 = T(1) + dt * k.a_max * (t1 - T(2) * sd2 * (k.tp + (v + dv_raw) * k.sab2_inv));
    T l11 = dt * k.a_max * (T(-2) * sd2 * (-v * k.sab2_inv));
    L10 = e.clip_acc ? T(0) : l10;
    E10 = e.clip_acc ? T(0) : -l10;
    E11 = e.clip_acc ? T(0) : e11;
    L11 = e.clip_acc ? T(0) : l11;
}

}  // namespace dhts
