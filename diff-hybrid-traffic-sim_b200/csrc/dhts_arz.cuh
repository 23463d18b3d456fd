// ARZ (Aw-Rascle-Zhang) device math for the B200 path: per-cell derived record,
// per-interface Riemann flux, per-interface adjoint in flux-difference form.
//
// Behaviour restated from the reference (file:line relative to its checkout):
//   model/macro/_arz.py:121-149      u_eq, u_eq', compute_u, compute_y
//   model/macro/_arz.py:155-199      Q_L / Q_C / Q_M candidate states
//   model/macro/_arz.py:212-332      riemann_solve case tree (order matters)
//   model/macro/darz.py:12-233       dL / dM / dC Jacobians and flux_prime
//   road/lane/_macro_lane.py:83-146  Godunov update and CFL assert
//   road/lane/dmacro_lane.py:96-132, 277-310  Jacobian band and its VJP
// gamma is the module constant 0.5 (no caller overrides it), so r^gamma =
// sqrt(r), r^(gamma-1) = 1/sqrt(r), x^(1/gamma) = x*x.
//
// The adjoint is NOT a stored Jacobian band: with A_i = F'(Q0_i) dQ0_i/dQ_left
// and B_i = F'(Q0_i) dQ0_i/dQ_right per interface i, and w_i = g_{i+1} - g_i,
//   gbar_j = g_j + c (A_j^T w_j + B_{j-1}^T w_{j-1})           
// which is algebraically the band-transpose product of dmacro_lane.py:293-303.
#pragma once
#include <cuda_runtime.h>

namespace dhts {

constexpr int FLAG_CFL = 1;        // _macro_lane.py:141-146 assert would fire
constexpr int FLAG_NAN_GRAD = 2;   // dmacro_lane.py:308 assert would fire
constexpr int FLAG_COLLISION = 4;  // _micro_lane.py:151-162 print-and-continue

template <typename T> __device__ __forceinline__ T t_sqrt(T x);
template <> __device__ __forceinline__ double t_sqrt<double>(double x) { return sqrt(x); }
template <> __device__ __forceinline__ float t_sqrt<float>(float x) { return sqrtf(x); }
template <typename T> __device__ __forceinline__ T t_abs(T x) { return x < T(0) ? -x : x; }
template <typename T> __device__ __forceinline__ T t_max(T a, T b) { return a > b ? a : b; }
template <typename T> __device__ __forceinline__ bool t_isnan(T x) { return !(x == x); }

#define DHTS_EPS (T(1e-5))
#define DHTS_RSQRT_EPS (T(316.22776601683796))   // 1/sqrt(1e-5)

// Per-cell record.  The reference STORES u and u_eq on each cell instead of
// recomputing them at use:
//   us = stored speed (drives the case tree, Q_M and Q_C),
//   uc = compute_u(r, y) (what a Q_L outcome re-derives, _arz.py:155-165),
//   w  = u_max + us - u_eq_stored (every vacuum / rarefaction test, :237,283,308),
//   uf = u_eq(r) freshly evaluated (what flux_prime sees on Q_L, darz.py:217-233),
//   sq = sqrt(max(r,eps)), rs = 1/sq, ri = 1/max(r,eps)  (rs, ri: adjoint only).
template <typename T> struct Cell {
    T r, y, us, uc, w, sq, rs, ri, uf;
};

// u_eq(r) = u_max (1 - (max(r,0)+eps)^gamma)                    _arz.py:133-138
template <typename T> __device__ __forceinline__ T u_eq(T r, T umax) {
    return umax * (T(1) - t_sqrt(t_max(r, T(0)) + DHTS_EPS));
}
// compute_u(r, y)                                               _arz.py:126-131
template <typename T> __device__ __forceinline__ T compute_u(T r, T y, T umax) {
    T rc = t_max(r, DHTS_EPS);
    return y / rc + umax * (T(1) - t_sqrt(rc + DHTS_EPS));
}

// Record of a cell whose (u, u_eq) follow set_r_y (_arz.py:88-92), i.e. every
// interior cell after a step.
template <typename T, bool ADJ> __device__ __forceinline__ Cell<T> derive_cell(T r, T y, T umax) {
    Cell<T> c;
    T rc = t_max(r, DHTS_EPS);
    T ueq_c = umax * (T(1) - t_sqrt(rc + DHTS_EPS));   // compute_u evaluates u_eq at the clamped r
    c.r = r; c.y = y;
    c.sq = t_sqrt(rc);
    if (ADJ) {
        c.ri = T(1) / rc; c.rs = c.sq * c.ri;          // 1/sqrt(rc) = sqrt(rc)/rc
        c.uc = y * c.ri + ueq_c;
    } else {
        c.ri = T(0); c.rs = T(0);
        c.uc = y / rc + ueq_c;
    }
    c.us = c.uc;
    c.uf = (r >= DHTS_EPS) ? ueq_c : u_eq(r, umax);    // same value when r >= eps
    c.w = umax + c.us - c.uf;
    return c;
}

// Record of a cell with an explicitly stored speed: ghosts built by from_r_u
// (_arz.py:74-80), initial cells set by set_r_u (:82-86), cells rewritten by
// micro_to_macro (conversion.py:157-167; their stored u_eq is stale).
template <typename T, bool ADJ>
__device__ __forceinline__ Cell<T> derive_cell_stored(T r, T y, T us, T ueq_stored, bool have_ueq, T umax) {
    Cell<T> c = derive_cell<T, ADJ>(r, y, umax);
    c.us = us;
    c.w = umax + us - (have_ueq ? ueq_stored : c.uf);
    return c;
}

template <typename T> struct Riem {
    int cas;        // 0 = Q_L, 1 = Q_M, 2 = Q_C                 _arz.py:209
    T r0, y0, u0, ueq0;
    T rootr;        // sqrt(r0) without re-evaluating a square root (r0 = b^2 for Q_M / Q_C)
    bool cfl_bad;
};

// Riemann solve at one interface.  Case tree of _arz.py:225-322, selected state
// per :324-336.  Where the reference only consumes the SIGN of a wave speed the
// division is skipped; the CFL test dt < dx / max(|s|,1e-5) is evaluated as
// dt * max(|s|,1e-5) < dx.
template <typename T>
__device__ __forceinline__ Riem<T> riemann(const Cell<T>& L, const Cell<T>& R, T umax, T inv_umax, T inv15, T dt,
                                           T dx) {
    Riem<T> o;
    T s0 = T(0), s1;
    bool shock = false; T fd = T(0), den = T(1), b = T(0);
    if (L.r < DHTS_EPS) {                                          // :225
        o.cas = 0; s1 = L.us;
    } else {
        T lam0l = L.us - T(0.5) * umax * L.sq;                     // u + r u_eq'(r), r >= eps   :103
        if (R.r < DHTS_EPS) {                                      // :235
            s0 = (lam0l + L.w) * T(0.5); s1 = s0;
            o.cas = (lam0l >= T(0)) ? 0 : 2;
        } else if (t_abs(L.us - R.us) < DHTS_EPS) {                // :256
            o.cas = 0; s1 = R.us;
        } else if (L.us > R.us) {                                  // :265 shock
            b = L.sq + (L.us - R.us) * inv_umax;
            T rm = b * b;
            fd = rm * R.us - L.r * L.us;
            den = t_max(rm - L.r, DHTS_EPS);
            shock = true; s1 = R.us;
            o.cas = (fd >= T(0)) ? 0 : 1;
        } else if (L.w > R.us) {                                   // :283 rarefaction
            b = L.sq + (L.us - R.us) * inv_umax;
            T rm = b * b;
            T lam0m = R.us - T(0.5) * umax * ((rm >= DHTS_EPS) ? t_abs(b) : rm * DHTS_RSQRT_EPS);
            s0 = (lam0l + lam0m) * T(0.5); s1 = R.us;
            o.cas = (lam0l >= T(0)) ? 0 : ((lam0m <= T(0)) ? 1 : 2);
        } else {                                                   // :306 vacuum forms
            s0 = (lam0l + L.w) * T(0.5); s1 = R.us;
            o.cas = (lam0l >= T(0)) ? 0 : 2;
        }
    }
    // CFL (_macro_lane.py:137-146)
    bool ok1 = dt * t_max(t_abs(s1), T(1e-5)) < dx;
    bool ok0 = shock ? (dt * t_max(t_abs(fd), T(1e-5) * den) < dx * den) : (dt * t_max(t_abs(s0), T(1e-5)) < dx);
    o.cfl_bad = !(ok0 && ok1);
    if (o.cas == 0) {                                              // compute_Ql :155-165
        o.r0 = L.r; o.y0 = L.y; o.u0 = L.uc; o.ueq0 = L.uf; o.rootr = L.sq;
    } else if (o.cas == 1) {                                       // compute_Qm :186-199
        o.r0 = b * b; o.u0 = R.us;
        o.ueq0 = umax * (T(1) - t_sqrt(o.r0 + DHTS_EPS));
        o.y0 = o.r0 * (o.u0 - o.ueq0);
        o.rootr = t_abs(b);
    } else {                                                       // compute_Qc :168-183
        T sc = L.us + umax * L.sq;
        T q = sc * inv15;
        o.r0 = q * q; o.u0 = (T(0.5) / T(1.5)) * sc;
        o.ueq0 = umax * (T(1) - t_sqrt(o.r0 + DHTS_EPS));
        o.y0 = o.r0 * (o.u0 - o.ueq0);
        o.rootr = t_abs(q);
    }
    return o;
}

// Adjoint of one interface flux: given w = g_{right} - g_{left} (2-vector),
// returns pa = A^T w (goes to the left cell) and pb = B^T w (right cell).
// L and R must carry rs / ri (derive_cell<T, true>).
template <typename T>
__device__ __forceinline__ void riemann_adj(const Cell<T>& L, const Cell<T>& R, const Riem<T>& s, T umax, T inv_umax,
                                            T inv15, T wr, T wy, T& par, T& pay, T& pbr, T& pby) {
    // flux_prime at Q0 (darz.py:217-233), transposed and applied to w
    T rr, inv_sq, inv_rr;
    if (s.cas == 0) { rr = t_max(s.r0, DHTS_EPS); inv_sq = L.rs; inv_rr = L.ri; }
    else {
        bool big = s.r0 >= DHTS_EPS;
        rr = big ? s.r0 : DHTS_EPS;
        inv_sq = big ? T(1) / s.rootr : DHTS_RSQRT_EPS;
        inv_rr = inv_sq * inv_sq;
    }
    T ueqp0 = T(-0.5) * umax * inv_sq;                 // u_eq'(max(r0,eps))  _arz.py:146-149
    T yr = s.y0 * inv_rr;
    T f00 = s.ueq0 + rr * ueqp0;
    T f10 = s.y0 * ueqp0 - yr * yr;
    T f11 = T(2) * yr + s.ueq0;
    T z0 = f00 * wr + f10 * wy;
    T z1 = wr + f11 * wy;
    if (s.cas == 0) {                                  // dL = I, dR = 0      darz.py:12-33
        par = z0; pay = z1; pbr = T(0); pby = T(0);
        return;
    }
    T ueqpL = T(-0.5) * umax * L.rs;
    T duL_drL = -L.y * L.ri * L.ri + ueqpL;
    T duL_dyL = L.ri;
    if (s.cas == 1) {                                  // compute_dM          darz.py:35-122
        T ueqpR = T(-0.5) * umax * R.rs;
        T duR_drR = -R.y * R.ri * R.ri + ueqpR;
        T duR_dyR = R.ri;
        T a = T(2) * s.rootr;                          // (1/gamma) rM^(1-gamma)
        T drM_drL = a * (T(0.5) * L.rs + duL_drL * inv_umax);
        T drM_dyL = a * (duL_dyL * inv_umax);
        T k = (s.u0 - s.ueq0) - s.r0 * ueqp0;          // e - rM u_eq'(rM)
        T dyM_drL = drM_drL * k;
        T dyM_dyL = drM_dyL * k;
        T drM_drR = -a * (duR_drR * inv_umax);
        T drM_dyR = -a * (duR_dyR * inv_umax);
        T dyM_drR = drM_drR * k + s.r0 * duR_drR;
        T dyM_dyR = drM_dyR * k + s.r0 * duR_dyR;
        par = drM_drL * z0 + dyM_drL * z1; pay = drM_dyL * z0 + dyM_dyL * z1;
        pbr = drM_drR * z0 + dyM_drR * z1; pby = drM_dyR * z0 + dyM_dyR * z1;
    } else {                                           // compute_dC          darz.py:124-192
        const T g13 = T(0.5) / T(1.5);
        T f = umax * T(0.5) * L.rs;
        T duC_drL = g13 * (duL_drL + f);
        T duC_dyL = g13 * duL_dyL;
        T e = T(2) * s.rootr * inv15;                  // rC^(1-gamma) / gamma / ((gamma+1) u_max)
        T drC_drL = e * (duL_drL + f);
        T drC_dyL = e * duL_dyL;
        T g = s.u0 - s.ueq0;
        T dyC_drL = drC_drL * g + s.r0 * (duC_drL - ueqp0 * drC_drL);
        T dyC_dyL = drC_dyL * g + s.r0 * (duC_dyL - ueqp0 * drC_dyL);
        par = drC_drL * z0 + dyC_drL * z1; pay = drC_dyL * z0 + dyC_dyL * z1;
        pbr = T(0); pby = T(0);
    }
}

// d compute_u / d(r, y) as autograd differentiates set_r_y OUTSIDE the
// reference's Function (_arz.py:88-92,126-138): the TRUE derivative of
// y/max(r,eps) + u_max (1 - (max(r,eps)+eps)^gamma).
template <typename T> __device__ __forceinline__ void du_dry(T r, T y, T umax, T& du_dr, T& du_dy) {
    if (r >= DHTS_EPS) {
        T ri = T(1) / r;
        du_dy = ri;
        du_dr = -y * ri * ri - T(0.5) * umax / t_sqrt(r + DHTS_EPS);
    } else {
        du_dy = T(1) / DHTS_EPS;
        du_dr = T(0);
    }
}

}  // namespace dhts
