// ARZ macroscopic lane kernels for sm_100a (K1 forward, K2 adjoint).
//
//  * arz_step_{fwd,bwd}_kernel   one Godunov step over B lanes x N cells, tiled:
//    each CTA stages a tile of cells plus a one-cell halo per side in shared
//    memory (coalesced loads), solves the tile's interfaces once, then updates.
//    This is the operator behind the reference's per-step autograd contract
//    dMacroForwardLayer (road/lane/dmacro_lane.py:234-310).
//  * arz_rollout_{fwd,bwd}_kernel  T fused steps with the whole lane resident in
//    shared memory (a 1024-cell fp64 lane is 16 KB of state), static ghost
//    cells, state checkpoints every K steps in HBM; the adjoint walks segments
//    backwards, recomputes the segment forward into an L2-resident per-CTA
//    scratch, then applies the flux-difference adjoint step by step.
//    Replaces T x (RoadNetwork.forward -> dMacroLane.forward) plus the autograd
//    chain of example/inverse/_inverse.py:91-99,227.
//
// No tensor cores: nothing here is a contraction (BASELINE.json north_star).
#include "dhts_arz.cuh"
#include "dhts_api.h"

namespace dhts {

constexpr int STEP_TILE = 256;

template <typename T> struct SM {
    T *r, *y, *us, *uc, *w, *sq, *rs, *ri, *uf;   // per padded cell
    T *f0, *f1, *f2, *f3;                         // per interface (flux, or A^T w / B^T w)
    T *gr, *gy;                                   // per padded cell (adjoint)
};

template <typename T, bool ADJ> __device__ __forceinline__ SM<T> carve(unsigned char* raw, int ncell, int nif) {
    SM<T> s;
    T* p = reinterpret_cast<T*>(raw);
    s.r = p; p += ncell; s.y = p; p += ncell; s.us = p; p += ncell; s.uc = p; p += ncell;
    s.w = p; p += ncell; s.sq = p; p += ncell;
    s.f0 = p; p += nif; s.f1 = p; p += nif;
    if (ADJ) {
        s.rs = p; p += ncell; s.ri = p; p += ncell; s.uf = p; p += ncell;
        s.f2 = p; p += nif; s.f3 = p; p += nif;
        s.gr = p; p += ncell; s.gy = p; p += ncell;
    } else {
        s.rs = s.ri = s.uf = s.f2 = s.f3 = s.gr = s.gy = nullptr;
    }
    return s;
}
template <typename T, bool ADJ> __host__ __device__ inline size_t smem_bytes(int ncell, int nif) {
    return sizeof(T) * (size_t)(ADJ ? (11 * ncell + 4 * nif) : (6 * ncell + 2 * nif));
}

template <typename T, bool ADJ> __device__ __forceinline__ void st_cell(const SM<T>& s, int q, const Cell<T>& c) {
    s.r[q] = c.r; s.y[q] = c.y; s.us[q] = c.us; s.uc[q] = c.uc; s.w[q] = c.w; s.sq[q] = c.sq;
    if (ADJ) { s.rs[q] = c.rs; s.ri[q] = c.ri; s.uf[q] = c.uf; }
}
template <typename T, bool ADJ> __device__ __forceinline__ Cell<T> ld_cell(const SM<T>& s, int q) {
    Cell<T> c;
    c.r = s.r[q]; c.y = s.y[q]; c.us = s.us[q]; c.uc = s.uc[q]; c.w = s.w[q]; c.sq = s.sq[q];
    if (ADJ) { c.rs = s.rs[q]; c.ri = s.ri[q]; c.uf = s.uf[q]; }
    else { c.rs = T(0); c.ri = T(0); c.uf = T(0); }
    return c;
}

// ------------------------------------------------------------------ single step, tiled

template <typename T>
__global__ void __launch_bounds__(STEP_TILE)
arz_step_fwd_kernel(const T* __restrict__ r_pad, const T* __restrict__ y_pad, const T* __restrict__ u_pad,
                    const T* __restrict__ ueq_pad, const T* __restrict__ dx, const T* __restrict__ umax_, T dt, int B,
                    int N, int ntile, T* __restrict__ nr, T* __restrict__ ny, T* __restrict__ nu,
                    int* __restrict__ case_out, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char raw[];
    SM<T> s = carve<T, false>(raw, STEP_TILE + 2, STEP_TILE + 1);
    const int b = blockIdx.x / ntile, tile = blockIdx.x % ntile;
    const int c0 = tile * STEP_TILE;
    const int nc = min(STEP_TILE, N - c0);
    const int P = N + 2;
    const T umax = umax_[b], inv_umax = T(1) / umax, inv15 = T(1) / (T(1.5) * umax), dxb = dx[b], cc = dt / dxb;
    const size_t base = (size_t)b * P + c0;
    for (int k = threadIdx.x; k < nc + 2; k += STEP_TILE) {
        T ue = ueq_pad ? ueq_pad[base + k] : T(0);
        st_cell<T, false>(s, k, derive_cell_stored<T, false>(r_pad[base + k], y_pad[base + k], u_pad[base + k], ue,
                                                              ueq_pad != nullptr, umax));
    }
    __syncthreads();
    bool bad = false;
    for (int i = threadIdx.x; i < nc + 1; i += STEP_TILE) {
        Riem<T> o = riemann(ld_cell<T, false>(s, i), ld_cell<T, false>(s, i + 1), umax, inv_umax, inv15, dt, dxb);
        s.f0[i] = o.r0 * o.u0; s.f1[i] = o.y0 * o.u0;
        bad |= o.cfl_bad;
        if (case_out) case_out[(size_t)b * (N + 1) + c0 + i] = o.cas;
    }
    __syncthreads();
    for (int j = threadIdx.x; j < nc; j += STEP_TILE) {
        T r = s.r[j + 1] + (s.f0[j] - s.f0[j + 1]) * cc;
        T y = s.y[j + 1] + (s.f1[j] - s.f1[j + 1]) * cc;
        size_t o = (size_t)b * N + c0 + j;
        nr[o] = r; ny[o] = y; nu[o] = compute_u(r, y, umax);
    }
    if (bad) atomicOr(flags, FLAG_CFL);
}

template <typename T>
__global__ void __launch_bounds__(STEP_TILE)
arz_step_bwd_kernel(const T* __restrict__ r_pad, const T* __restrict__ y_pad, const T* __restrict__ u_pad,
                    const T* __restrict__ ueq_pad, const T* __restrict__ dx, const T* __restrict__ umax_, T dt, int B,
                    int N, int ntile, const T* __restrict__ nr, const T* __restrict__ ny,
                    const T* __restrict__ g_nr, const T* __restrict__ g_ny, const T* __restrict__ g_nu,
                    T* __restrict__ g_r_pad, T* __restrict__ g_y_pad, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char raw[];
    SM<T> s = carve<T, true>(raw, STEP_TILE + 2, STEP_TILE + 1);
    const int b = blockIdx.x / ntile, tile = blockIdx.x % ntile;
    const int c0 = tile * STEP_TILE;
    const int nc = min(STEP_TILE, N - c0);
    const int P = N + 2;
    const T umax = umax_[b], inv_umax = T(1) / umax, inv15 = T(1) / (T(1.5) * umax), dxb = dx[b], cc = dt / dxb;
    const size_t base = (size_t)b * P + c0;
    for (int k = threadIdx.x; k < nc + 2; k += STEP_TILE) {
        T ue = ueq_pad ? ueq_pad[base + k] : T(0);
        st_cell<T, true>(s, k, derive_cell_stored<T, true>(r_pad[base + k], y_pad[base + k], u_pad[base + k], ue,
                                                            ueq_pad != nullptr, umax));
        int j = c0 + k;   // padded index; updated cells are 1..N
        T gr = T(0), gy = T(0);
        if (j >= 1 && j <= N) {
            size_t o = (size_t)b * N + j - 1;
            gr = g_nr[o]; gy = g_ny[o];
            if (g_nu) {   // nu = compute_u(nr, ny) is produced inside the operator
                T dr, dy; du_dry(nr[o], ny[o], umax, dr, dy);
                T gu = g_nu[o]; gr += gu * dr; gy += gu * dy;
            }
        }
        s.gr[k] = gr; s.gy[k] = gy;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < nc + 1; i += STEP_TILE) {
        Cell<T> L = ld_cell<T, true>(s, i), R = ld_cell<T, true>(s, i + 1);
        Riem<T> o = riemann(L, R, umax, inv_umax, inv15, dt, dxb);
        riemann_adj(L, R, o, umax, inv_umax, inv15, s.gr[i + 1] - s.gr[i], s.gy[i + 1] - s.gy[i], s.f0[i], s.f1[i],
                    s.f2[i], s.f3[i]);
    }
    __syncthreads();
    bool bad = false;
    for (int j = threadIdx.x; j < nc; j += STEP_TILE) {
        int k = j + 1;
        T gr = s.gr[k] + cc * (s.f0[k] + s.f2[k - 1]);
        T gy = s.gy[k] + cc * (s.f1[k] + s.f3[k - 1]);
        bad |= t_isnan(gr) || t_isnan(gy);
        g_r_pad[base + k] = gr; g_y_pad[base + k] = gy;
    }
    if (threadIdx.x == 0) {
        if (tile == 0) {
            T gr = cc * s.f0[0], gy = cc * s.f1[0];
            bad |= t_isnan(gr) || t_isnan(gy);
            g_r_pad[(size_t)b * P] = gr; g_y_pad[(size_t)b * P] = gy;
        }
        if (tile == ntile - 1) {
            T gr = cc * s.f2[nc], gy = cc * s.f3[nc];
            bad |= t_isnan(gr) || t_isnan(gy);
            g_r_pad[(size_t)b * P + N + 1] = gr; g_y_pad[(size_t)b * P + N + 1] = gy;
        }
    }
    if (bad) atomicOr(flags, FLAG_NAN_GRAD);
}

// ------------------------------------------------------------------ fused rollouts, lane resident in smem

template <typename T> struct LaneConst { T umax, inv_umax, inv15, dx, cc; };

template <typename T, bool ADJ>
__device__ __forceinline__ void load_group(const SM<T>& s, LaneConst<T>* lc, int lane0, int nl, int N,
                                           const T* __restrict__ ghost, const T* __restrict__ dx,
                                           const T* __restrict__ umax_, T dt) {
    const int P = N + 2;
    for (int l = threadIdx.x; l < nl; l += blockDim.x) {
        T um = umax_[lane0 + l], d = dx[lane0 + l];
        lc[l].umax = um; lc[l].inv_umax = T(1) / um; lc[l].inv15 = T(1) / (T(1.5) * um); lc[l].dx = d; lc[l].cc = dt / d;
    }
    // static ghost cells (from_r_u records): padded index 0 and N+1 of every lane
    for (int e = threadIdx.x; e < nl * 2; e += blockDim.x) {
        int l = e >> 1, side = e & 1;
        const T* g = ghost + ((size_t)(lane0 + l) * 2 + side) * 3;
        st_cell<T, ADJ>(s, l * P + (side ? N + 1 : 0),
                        derive_cell_stored<T, ADJ>(g[0], g[1], g[2], T(0), false, umax_[lane0 + l]));
        if (ADJ) { s.gr[l * P + (side ? N + 1 : 0)] = T(0); s.gy[l * P + (side ? N + 1 : 0)] = T(0); }
    }
}

// interior cells of a group from global (r, y[, stored u]) arrays laid out [B][N]
template <typename T, bool ADJ>
__device__ __forceinline__ void load_cells(const SM<T>& s, const LaneConst<T>* lc, int nl, int N,
                                           const T* __restrict__ r, const T* __restrict__ y,
                                           const T* __restrict__ u_stored) {
    const int P = N + 2;
    for (int c = threadIdx.x; c < nl * N; c += blockDim.x) {
        int l = c / N, j = c - l * N;
        Cell<T> cl = u_stored ? derive_cell_stored<T, ADJ>(r[c], y[c], u_stored[c], T(0), false, lc[l].umax)
                              : derive_cell<T, ADJ>(r[c], y[c], lc[l].umax);
        st_cell<T, ADJ>(s, l * P + j + 1, cl);
    }
}

// one forward step on the smem-resident group; returns the CFL verdict of this thread's interfaces
template <typename T, bool ADJ>
__device__ __forceinline__ bool fwd_step(const SM<T>& s, const LaneConst<T>* lc, int nl, int N, T dt) {
    const int P = N + 2, NI = N + 1;
    bool bad = false;
    for (int i = threadIdx.x; i < nl * NI; i += blockDim.x) {
        int l = i / NI, k = i - l * NI, q = l * P + k;
        Riem<T> o = riemann(ld_cell<T, false>(s, q), ld_cell<T, false>(s, q + 1), lc[l].umax, lc[l].inv_umax,
                            lc[l].inv15, dt, lc[l].dx);
        s.f0[i] = o.r0 * o.u0; s.f1[i] = o.y0 * o.u0;
        bad |= o.cfl_bad;
    }
    __syncthreads();
    for (int c = threadIdx.x; c < nl * N; c += blockDim.x) {
        int l = c / N, j = c - l * N, q = l * P + j + 1, i = l * NI + j;
        T r = s.r[q] + (s.f0[i] - s.f0[i + 1]) * lc[l].cc;
        T y = s.y[q] + (s.f1[i] - s.f1[i + 1]) * lc[l].cc;
        st_cell<T, false>(s, q, derive_cell<T, false>(r, y, lc[l].umax));
    }
    __syncthreads();
    return bad;
}

template <typename T>
__global__ void arz_rollout_fwd_kernel(const T* __restrict__ r0, const T* __restrict__ y0, const T* __restrict__ u0,
                                       const T* __restrict__ ghost, const T* __restrict__ dx,
                                       const T* __restrict__ umax_, T dt, int B, int N, int steps, int K, int lpc,
                                       T* __restrict__ ckpt, T* __restrict__ rT, T* __restrict__ yT,
                                       T* __restrict__ uT, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int P = N + 2;
    SM<T> s = carve<T, false>(raw, lpc * P, lpc * (N + 1));
    LaneConst<T>* lc = reinterpret_cast<LaneConst<T>*>(raw + smem_bytes<T, false>(lpc * P, lpc * (N + 1)));
    const int ngroup = (B + lpc - 1) / lpc;
    bool bad = false;
    for (int grp = blockIdx.x; grp < ngroup; grp += gridDim.x) {
        const int lane0 = grp * lpc, nl = min(lpc, B - lane0);
        const size_t goff = (size_t)lane0 * N;
        __syncthreads();
        load_group<T, false>(s, lc, lane0, nl, N, ghost, dx, umax_, dt);
        __syncthreads();
        load_cells<T, false>(s, lc, nl, N, r0 + goff, y0 + goff, u0 ? u0 + goff : nullptr);
        __syncthreads();
        for (int t = 0; t < steps; t++) {
            if (ckpt && t % K == 0) {
                T* cr = ckpt + ((size_t)(t / K) * 2) * B * N + goff;
                T* cy = cr + (size_t)B * N;
                for (int c = threadIdx.x; c < nl * N; c += blockDim.x) {
                    int l = c / N, j = c - l * N, q = l * P + j + 1;
                    cr[c] = s.r[q]; cy[c] = s.y[q];
                }
            }
            bad |= fwd_step<T, false>(s, lc, nl, N, dt);
        }
        for (int c = threadIdx.x; c < nl * N; c += blockDim.x) {
            int l = c / N, j = c - l * N, q = l * P + j + 1;
            rT[goff + c] = s.r[q]; yT[goff + c] = s.y[q]; uT[goff + c] = s.uc[q];
        }
    }
    if (bad) atomicOr(flags, FLAG_CFL);
}

template <typename T>
__global__ void arz_rollout_bwd_kernel(const T* __restrict__ ckpt, const T* __restrict__ u0,
                                       const T* __restrict__ ghost, const T* __restrict__ dx,
                                       const T* __restrict__ umax_, T dt, int B, int N, int steps, int K, int lpc,
                                       const T* __restrict__ rT, const T* __restrict__ yT,
                                       const T* __restrict__ g_rT, const T* __restrict__ g_yT,
                                       const T* __restrict__ g_uT, T* __restrict__ scratch, T* __restrict__ g_r0,
                                       T* __restrict__ g_y0, T* __restrict__ g_ghost, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int P = N + 2, NI = N + 1;
    SM<T> s = carve<T, true>(raw, lpc * P, lpc * NI);
    unsigned char* tail = raw + smem_bytes<T, true>(lpc * P, lpc * NI);
    LaneConst<T>* lc = reinterpret_cast<LaneConst<T>*>(tail);
    T* gacc = reinterpret_cast<T*>(tail + sizeof(LaneConst<T>) * lpc);   // [lpc][4] ghost adjoints (r,y) x (left,right)
    const int ngroup = (B + lpc - 1) / lpc;
    const int S = (steps + K - 1) / K;
    // per-CTA scratch: [K][2][lpc*N]
    T* scr = scratch + (size_t)blockIdx.x * K * 2 * lpc * N;
    const size_t sstride = (size_t)2 * lpc * N;
    bool bad = false;
    for (int grp = blockIdx.x; grp < ngroup; grp += gridDim.x) {
        const int lane0 = grp * lpc, nl = min(lpc, B - lane0);
        const size_t goff = (size_t)lane0 * N;
        __syncthreads();
        load_group<T, true>(s, lc, lane0, nl, N, ghost, dx, umax_, dt);
        for (int e = threadIdx.x; e < nl * 4; e += blockDim.x) gacc[e] = T(0);
        __syncthreads();
        // terminal adjoint; uT = compute_u(rT, yT) is produced inside the operator
        for (int c = threadIdx.x; c < nl * N; c += blockDim.x) {
            int l = c / N, j = c - l * N, q = l * P + j + 1;
            T gr = g_rT ? g_rT[goff + c] : T(0), gy = g_yT ? g_yT[goff + c] : T(0);
            if (g_uT) {
                T dr, dy; du_dry(rT[goff + c], yT[goff + c], lc[l].umax, dr, dy);
                T gu = g_uT[goff + c]; gr += gu * dr; gy += gu * dy;
            }
            s.gr[q] = gr; s.gy[q] = gy;
        }
        for (int seg = S - 1; seg >= 0; seg--) {
            const int t0 = seg * K, ks = min(K, steps - t0);
            const T* cr = ckpt + ((size_t)seg * 2) * B * N + goff;
            const T* cy = cr + (size_t)B * N;
            __syncthreads();
            load_cells<T, false>(s, lc, nl, N, cr, cy, (seg == 0 && u0) ? u0 + goff : nullptr);
            __syncthreads();
            // recompute the segment, stashing every state in the CTA's scratch (stays in L2)
            for (int k = 0; k < ks; k++) {
                T* sr = scr + (size_t)k * sstride; T* sy = sr + (size_t)lpc * N;
                for (int c = threadIdx.x; c < nl * N; c += blockDim.x) {
                    int l = c / N, j = c - l * N, q = l * P + j + 1;
                    sr[c] = s.r[q]; sy[c] = s.y[q];
                }
                if (k + 1 < ks) fwd_step<T, true>(s, lc, nl, N, dt);
            }
            // adjoint steps, last to first
            for (int k = ks - 1; k >= 0; k--) {
                const T* sr = scr + (size_t)k * sstride; const T* sy = sr + (size_t)lpc * N;
                __syncthreads();
                load_cells<T, true>(s, lc, nl, N, sr, sy, (t0 + k == 0 && u0) ? u0 + goff : nullptr);
                __syncthreads();
                for (int i = threadIdx.x; i < nl * NI; i += blockDim.x) {
                    int l = i / NI, kk = i - l * NI, q = l * P + kk;
                    Cell<T> L = ld_cell<T, true>(s, q), R = ld_cell<T, true>(s, q + 1);
                    Riem<T> o = riemann(L, R, lc[l].umax, lc[l].inv_umax, lc[l].inv15, dt, lc[l].dx);
                    riemann_adj(L, R, o, lc[l].umax, lc[l].inv_umax, lc[l].inv15, s.gr[q + 1] - s.gr[q],
                                s.gy[q + 1] - s.gy[q], s.f0[i], s.f1[i], s.f2[i], s.f3[i]);
                }
                __syncthreads();
                for (int c = threadIdx.x; c < nl * N; c += blockDim.x) {
                    int l = c / N, j = c - l * N, q = l * P + j + 1, i = l * NI + j + 1;
                    T cc = lc[l].cc;
                    T gr = s.gr[q] + cc * (s.f0[i] + s.f2[i - 1]);
                    T gy = s.gy[q] + cc * (s.f1[i] + s.f3[i - 1]);
                    bad |= t_isnan(gr) || t_isnan(gy);
                    s.gr[q] = gr; s.gy[q] = gy;
                    if (j == 0) { gacc[l * 4 + 0] += cc * s.f0[i - 1]; gacc[l * 4 + 1] += cc * s.f1[i - 1]; }
                    if (j == N - 1) { gacc[l * 4 + 2] += cc * s.f2[i]; gacc[l * 4 + 3] += cc * s.f3[i]; }
                }
            }
        }
        __syncthreads();
        for (int c = threadIdx.x; c < nl * N; c += blockDim.x) {
            int l = c / N, j = c - l * N, q = l * P + j + 1;
            g_r0[goff + c] = s.gr[q]; g_y0[goff + c] = s.gy[q];
        }
        if (g_ghost)
            for (int e = threadIdx.x; e < nl * 4; e += blockDim.x) {
                bad |= t_isnan(gacc[e]);
                g_ghost[(size_t)lane0 * 4 + e] = gacc[e];
            }
    }
    if (bad) atomicOr(flags, FLAG_NAN_GRAD);
}

// ------------------------------------------------------------------ host-side launch planning

struct RolloutPlan { int lpc, threads, grid; size_t smem; };

static int round_up32(int x) { return (x + 31) / 32 * 32; }

template <typename T, bool ADJ> static int plan_rollout(int B, int N, RolloutPlan* p) {
    const size_t cap = 200 * 1024;   // leave headroom below the 227 KB opt-in limit
    const int P = N + 2;
    size_t per_lane = smem_bytes<T, ADJ>(P, N + 1) + sizeof(LaneConst<T>) + 4 * sizeof(T);
    if (per_lane > cap) return DHTS_ERR_UNSUPPORTED;
    // pack several short lanes per CTA so that a CTA has >= 128 interfaces to solve
    int lpc = 1;
    if (N + 1 < 128) lpc = (128 + N) / (N + 1);
    while (lpc > 1 && per_lane * lpc > cap) lpc--;
    if (lpc > B) lpc = B;
    int items = lpc * (N + 1);
    int max_threads = ADJ ? 512 : 512;
    int m = (items + max_threads - 1) / max_threads;      // work items per thread
    int threads = round_up32((items + m - 1) / m);
    if (threads < 32) threads = 32;
    p->lpc = lpc; p->threads = threads; p->smem = per_lane * lpc + 16;
    p->grid = (B + lpc - 1) / lpc;
    return DHTS_OK;
}

static int g_sm_count = 0;
static int sm_count() {
    if (!g_sm_count) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&g_sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (g_sm_count <= 0) g_sm_count = 148;
    }
    return g_sm_count;
}

template <typename K> static int ensure_smem(K kernel, size_t bytes) {
    if (bytes > 48 * 1024) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes) != cudaSuccess) {
            cudaGetLastError();
            return DHTS_ERR_CUDA;
        }
    }
    return DHTS_OK;
}

static int last_status() { return cudaGetLastError() == cudaSuccess ? DHTS_OK : DHTS_ERR_CUDA; }

template <typename T>
static int arz_step_fwd(const T* r_pad, const T* y_pad, const T* u_pad, const T* ueq_pad, const T* dx, const T* umax,
                        T dt, int B, int N, T* nr, T* ny, T* nu, int* case_out, int* flags, cudaStream_t st) {
    if (!r_pad || !y_pad || !u_pad || !dx || !umax || !nr || !ny || !nu || !flags || B < 0 || N < 1) return DHTS_ERR_INVALID;
    if (B == 0) return DHTS_OK;
    int ntile = (N + STEP_TILE - 1) / STEP_TILE;
    if ((long long)B * ntile > 2147483647LL) return DHTS_ERR_UNSUPPORTED;
    size_t sm = smem_bytes<T, false>(STEP_TILE + 2, STEP_TILE + 1);
    arz_step_fwd_kernel<T><<<B * ntile, STEP_TILE, sm, st>>>(r_pad, y_pad, u_pad, ueq_pad, dx, umax, dt, B, N, ntile, nr,
                                                              ny, nu, case_out, flags);
    return last_status();
}

template <typename T>
static int arz_step_bwd(const T* r_pad, const T* y_pad, const T* u_pad, const T* ueq_pad, const T* dx, const T* umax,
                        T dt, int B, int N, const T* nr, const T* ny, const T* g_nr, const T* g_ny, const T* g_nu,
                        T* g_r_pad, T* g_y_pad, int* flags, cudaStream_t st) {
    if (!r_pad || !y_pad || !u_pad || !dx || !umax || !g_nr || !g_ny || !g_r_pad || !g_y_pad || !flags || B < 0 || N < 1)
        return DHTS_ERR_INVALID;
    if (g_nu && (!nr || !ny)) return DHTS_ERR_INVALID;
    if (B == 0) return DHTS_OK;
    int ntile = (N + STEP_TILE - 1) / STEP_TILE;
    if ((long long)B * ntile > 2147483647LL) return DHTS_ERR_UNSUPPORTED;
    size_t sm = smem_bytes<T, true>(STEP_TILE + 2, STEP_TILE + 1);
    int rc = ensure_smem(arz_step_bwd_kernel<T>, sm);
    if (rc) return rc;
    arz_step_bwd_kernel<T><<<B * ntile, STEP_TILE, sm, st>>>(r_pad, y_pad, u_pad, ueq_pad, dx, umax, dt, B, N, ntile, nr,
                                                              ny, g_nr, g_ny, g_nu, g_r_pad, g_y_pad, flags);
    return last_status();
}

template <typename T>
static int arz_rollout_fwd(const T* r0, const T* y0, const T* u0, const T* ghost, const T* dx, const T* umax, T dt,
                           int B, int N, int steps, int K, T* ckpt, T* rT, T* yT, T* uT, int* flags, cudaStream_t st) {
    if (!r0 || !y0 || !ghost || !dx || !umax || !rT || !yT || !uT || !flags || B < 0 || N < 1 || steps < 0)
        return DHTS_ERR_INVALID;
    if (ckpt && K < 1) return DHTS_ERR_INVALID;
    if (B == 0) return DHTS_OK;
    RolloutPlan p;
    int rc = plan_rollout<T, false>(B, N, &p);
    if (rc) return rc;
    rc = ensure_smem(arz_rollout_fwd_kernel<T>, p.smem);
    if (rc) return rc;
    arz_rollout_fwd_kernel<T><<<p.grid, p.threads, p.smem, st>>>(r0, y0, u0, ghost, dx, umax, dt, B, N, steps,
                                                                  K < 1 ? 1 : K, p.lpc, ckpt, rT, yT, uT, flags);
    return last_status();
}

template <typename T> static int bwd_grid(const RolloutPlan& p) {
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, arz_rollout_bwd_kernel<T>, p.threads, p.smem);
    if (occ < 1) occ = 1;
    int g = sm_count() * occ;
    return g < p.grid ? g : p.grid;
}

template <typename T> static long long arz_rollout_scratch_elems(int B, int N, int K) {
    RolloutPlan p;
    if (B <= 0) return 0;
    if (plan_rollout<T, true>(B, N, &p)) return -1;
    if (ensure_smem(arz_rollout_bwd_kernel<T>, p.smem)) return -1;
    return (long long)bwd_grid<T>(p) * K * 2 * p.lpc * N;
}

template <typename T>
static int arz_rollout_bwd(const T* ckpt, const T* u0, const T* ghost, const T* dx, const T* umax, T dt, int B, int N,
                           int steps, int K, const T* rT, const T* yT, const T* g_rT, const T* g_yT, const T* g_uT,
                           T* scratch, long long scratch_elems, T* g_r0, T* g_y0, T* g_ghost, int* flags,
                           cudaStream_t st) {
    if (!ckpt || !ghost || !dx || !umax || !g_r0 || !g_y0 || !flags || !scratch || B < 0 || N < 1 || steps < 0 || K < 1)
        return DHTS_ERR_INVALID;
    if (g_uT && (!rT || !yT)) return DHTS_ERR_INVALID;
    if (B == 0) return DHTS_OK;
    RolloutPlan p;
    int rc = plan_rollout<T, true>(B, N, &p);
    if (rc) return rc;
    rc = ensure_smem(arz_rollout_bwd_kernel<T>, p.smem);
    if (rc) return rc;
    int grid = bwd_grid<T>(p);
    if ((long long)grid * K * 2 * p.lpc * N > scratch_elems) return DHTS_ERR_INVALID;
    arz_rollout_bwd_kernel<T><<<grid, p.threads, p.smem, st>>>(ckpt, u0, ghost, dx, umax, dt, B, N, steps, K, p.lpc, rT,
                                                                yT, g_rT, g_yT, g_uT, scratch, g_r0, g_y0, g_ghost,
                                                                flags);
    return last_status();
}

}  // namespace dhts

// ------------------------------------------------------------------ C ABI (include/dhts.h)

#define DHTS_ARZ_API(SUF, T)                                                                                           \
    DHTS_EXPORT int dhts_arz_step_fwd_##SUF(const T* r_pad, const T* y_pad, const T* u_pad, const T* ueq_pad,          \
                                            const T* dx, const T* umax, T dt, int B, int N, T* nr, T* ny, T* nu,       \
                                            int* case_out, int* flags, void* stream) {                                 \
        return dhts::arz_step_fwd<T>(r_pad, y_pad, u_pad, ueq_pad, dx, umax, dt, B, N, nr, ny, nu, case_out, flags,    \
                                     (cudaStream_t)stream);                                                            \
    }                                                                                                                  \
    DHTS_EXPORT int dhts_arz_step_bwd_##SUF(const T* r_pad, const T* y_pad, const T* u_pad, const T* ueq_pad,          \
                                            const T* dx, const T* umax, T dt, int B, int N, const T* nr, const T* ny,  \
                                            const T* g_nr, const T* g_ny, const T* g_nu, T* g_r_pad, T* g_y_pad,       \
                                            int* flags, void* stream) {                                                \
        return dhts::arz_step_bwd<T>(r_pad, y_pad, u_pad, ueq_pad, dx, umax, dt, B, N, nr, ny, g_nr, g_ny, g_nu,       \
                                     g_r_pad, g_y_pad, flags, (cudaStream_t)stream);                                   \
    }                                                                                                                  \
    DHTS_EXPORT int dhts_arz_rollout_fwd_##SUF(const T* r0, const T* y0, const T* u0, const T* ghost, const T* dx,     \
                                               const T* umax, T dt, int B, int N, int steps, int ckpt_every, T* ckpt,  \
                                               T* rT, T* yT, T* uT, int* flags, void* stream) {                        \
        return dhts::arz_rollout_fwd<T>(r0, y0, u0, ghost, dx, umax, dt, B, N, steps, ckpt_every, ckpt, rT, yT, uT,    \
                                        flags, (cudaStream_t)stream);                                                  \
    }                                                                                                                  \
    DHTS_EXPORT long long dhts_arz_rollout_scratch_elems_##SUF(int B, int N, int ckpt_every) {                         \
        return dhts::arz_rollout_scratch_elems<T>(B, N, ckpt_every);                                                   \
    }                                                                                                                  \
    DHTS_EXPORT int dhts_arz_rollout_bwd_##SUF(const T* ckpt, const T* u0, const T* ghost, const T* dx, const T* umax, \
                                               T dt, int B, int N, int steps, int ckpt_every, const T* rT,             \
                                               const T* yT, const T* g_rT, const T* g_yT, const T* g_uT, T* scratch,   \
                                               long long scratch_elems, T* g_r0, T* g_y0, T* g_ghost, int* flags,      \
                                               void* stream) {                                                         \
        return dhts::arz_rollout_bwd<T>(ckpt, u0, ghost, dx, umax, dt, B, N, steps, ckpt_every, rT, yT, g_rT, g_yT,    \
                                        g_uT, scratch, scratch_elems, g_r0, g_y0, g_ghost, flags,                      \
                                        (cudaStream_t)stream);                                                         \
    }

// C ABI
DHTS_ARZ_API(f64, double)
DHTS_ARZ_API(f32, float)
// end C ABI
