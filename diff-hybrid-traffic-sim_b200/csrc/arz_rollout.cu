// Fused T-step ARZ rollouts for sm_100a, lanes REGISTER-resident.
//
// Each thread owns C consecutive cells of one lane (state, and in the adjoint
// the running adjoint, live in its registers for the whole rollout).  Per step a
// thread derives its cells' records once, solves the C interfaces to the left of
// its cells in one left-to-right sweep, and needs exactly two neighbour items:
//   - the record of the cell left of its chunk  (from thread t-1, before the sweep)
//   - the flux / A^T w at its right edge         (from thread t+1, after the sweep)
// Both travel by warp shuffle; only the two lanes at a warp edge go through a
// small double-buffered shared-memory mailbox (one __syncthreads per exchange).
// Ghost cells are static per lane (road/network/road_network.py:299-387 with no
// neighbouring macro lane) and live in shared memory.
//
// Forward: state checkpoints every K steps to HBM.  Adjoint: segments walked
// backwards; each is recomputed from its checkpoint while stashing every state
// in a per-CTA scratch (L2-resident), then the flux-difference adjoint of
// dhts_arz.cuh is applied step by step, last to first.
//
// Replaces T x (RoadNetwork.forward -> dMacroLane.forward -> dMacroForwardLayer)
// and the autograd chain through them (example/inverse/_inverse.py:91-99,227;
// road/lane/dmacro_lane.py:68-132,234-310).
#include <cstdint>
#include <cstdlib>
#include "dhts_arz.cuh"
#include "dhts_api.h"

namespace dhts {

constexpr unsigned FULL = 0xffffffffu;
constexpr int RF_FWD = 6;    // r, y, us, uc, w, sq
constexpr int RF_ADJ = 11;   // + rs, ri, uf, gr, gy

template <typename T> struct LaneK { T umax, inv_umax, inv15, dx, cc; };

template <typename T> __device__ __forceinline__ void pack_fwd(const Cell<T>& c, T* a) {
    a[0] = c.r; a[1] = c.y; a[2] = c.us; a[3] = c.uc; a[4] = c.w; a[5] = c.sq;
}
template <typename T> __device__ __forceinline__ Cell<T> unpack_fwd(const T* a) {
    Cell<T> c; c.r = a[0]; c.y = a[1]; c.us = a[2]; c.uc = a[3]; c.w = a[4]; c.sq = a[5];
    c.rs = T(0); c.ri = T(0); c.uf = T(0);
    return c;
}
template <typename T> __device__ __forceinline__ void pack_adj(const Cell<T>& c, T gr, T gy, T* a) {
    a[0] = c.r; a[1] = c.y; a[2] = c.us; a[3] = c.uc; a[4] = c.w; a[5] = c.sq; a[6] = c.rs; a[7] = c.ri; a[8] = c.uf;
    a[9] = gr; a[10] = gy;
}
template <typename T> __device__ __forceinline__ Cell<T> unpack_adj(const T* a) {
    Cell<T> c; c.r = a[0]; c.y = a[1]; c.us = a[2]; c.uc = a[3]; c.w = a[4]; c.sq = a[5]; c.rs = a[6]; c.ri = a[7];
    c.uf = a[8];
    return c;
}

// value held by thread t-1 -> thread t (shuffle; warp edges through the mailbox). All threads must call.
template <typename T, int NF>
__device__ __forceinline__ void from_left(const T* mine, T* out, T* box, int warp, unsigned lane) {
#pragma unroll
    for (int f = 0; f < NF; f++) out[f] = __shfl_up_sync(FULL, mine[f], 1);
    if (lane == 31) {
#pragma unroll
        for (int f = 0; f < NF; f++) box[warp * NF + f] = mine[f];
    }
    __syncthreads();
    if (lane == 0 && warp > 0) {
#pragma unroll
        for (int f = 0; f < NF; f++) out[f] = box[(warp - 1) * NF + f];
    }
}
// value held by thread t+1 -> thread t
template <typename T, int NF>
__device__ __forceinline__ void from_right(const T* mine, T* out, T* box, int warp, int nwarp, unsigned lane) {
#pragma unroll
    for (int f = 0; f < NF; f++) out[f] = __shfl_down_sync(FULL, mine[f], 1);
    if (lane == 0) {
#pragma unroll
        for (int f = 0; f < NF; f++) box[warp * NF + f] = mine[f];
    }
    __syncthreads();
    if (lane == 31 && warp + 1 < nwarp) {
#pragma unroll
        for (int f = 0; f < NF; f++) out[f] = box[(warp + 1) * NF + f];
    }
}

template <typename T, int C> __device__ __forceinline__ void load_chunk(const T* __restrict__ p, T* v) {
#pragma unroll
    for (int c = 0; c < C; c++) v[c] = p[c];
}
template <> __device__ __forceinline__ void load_chunk<double, 4>(const double* __restrict__ p, double* v) {
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        double2 a = reinterpret_cast<const double2*>(p)[0], b = reinterpret_cast<const double2*>(p)[1];
        v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y;
    } else { v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3]; }
}
template <> __device__ __forceinline__ void load_chunk<float, 4>(const float* __restrict__ p, float* v) {
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        float4 a = reinterpret_cast<const float4*>(p)[0];
        v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    } else { v[0] = p[0]; v[1] = p[1]; v[2] = p[2]; v[3] = p[3]; }
}
template <typename T, int C> __device__ __forceinline__ void store_chunk(T* __restrict__ p, const T* v) {
#pragma unroll
    for (int c = 0; c < C; c++) p[c] = v[c];
}
template <> __device__ __forceinline__ void store_chunk<double, 4>(double* __restrict__ p, const double* v) {
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) {
        reinterpret_cast<double2*>(p)[0] = make_double2(v[0], v[1]);
        reinterpret_cast<double2*>(p)[1] = make_double2(v[2], v[3]);
    } else { p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; p[3] = v[3]; }
}
template <> __device__ __forceinline__ void store_chunk<float, 4>(float* __restrict__ p, const float* v) {
    if ((reinterpret_cast<uintptr_t>(p) & 15) == 0) reinterpret_cast<float4*>(p)[0] = make_float4(v[0], v[1], v[2], v[3]);
    else { p[0] = v[0]; p[1] = v[1]; p[2] = v[2]; p[3] = v[3]; }
}

// Shared-memory layout of a CTA: per-lane constants, per-lane ghost records, two mailboxes x two parities.
template <typename T> struct Shm {
    LaneK<T>* lk;     // [lpc]
    T* ghost;         // [lpc][2][9]   left / right ghost records (ADJ fields included)
    T* gacc;          // [lpc][4]      ghost adjoints (r,y) x (left,right), adjoint kernel only
    T* boxL;          // [2][nwarp][RF_ADJ]
    T* boxR;          // [2][nwarp][4]
};
template <typename T> __host__ __device__ inline size_t shm_bytes(int lpc, int nwarp) {
    return sizeof(LaneK<T>) * lpc + sizeof(T) * ((size_t)lpc * (18 + 4) + (size_t)2 * nwarp * (RF_ADJ + 4)) + 16;
}
template <typename T> __device__ __forceinline__ Shm<T> carve_shm(unsigned char* raw, int lpc, int nwarp) {
    Shm<T> s;
    s.lk = reinterpret_cast<LaneK<T>*>(raw);
    T* p = reinterpret_cast<T*>(raw + sizeof(LaneK<T>) * lpc);
    s.ghost = p; p += lpc * 18; s.gacc = p; p += lpc * 4;
    s.boxL = p; p += 2 * nwarp * RF_ADJ; s.boxR = p;
    return s;
}

template <typename T>
__device__ __forceinline__ void setup_group(const Shm<T>& s, int lane0, int nl, const T* __restrict__ ghost,
                                            const T* __restrict__ dx, const T* __restrict__ umax_, T dt) {
    for (int l = threadIdx.x; l < nl; l += blockDim.x) {
        T um = umax_[lane0 + l], d = dx[lane0 + l];
        LaneK<T> k; k.umax = um; k.inv_umax = T(1) / um; k.inv15 = T(1) / (T(1.5) * um); k.dx = d; k.cc = dt / d;
        s.lk[l] = k;
    }
    for (int e = threadIdx.x; e < nl * 2; e += blockDim.x) {
        int l = e >> 1, side = e & 1;
        const T* g = ghost + ((size_t)(lane0 + l) * 2 + side) * 3;
        Cell<T> c = derive_cell_stored<T, true>(g[0], g[1], g[2], T(0), false, umax_[lane0 + l]);   // from_r_u record
        T* o = s.ghost + (l * 2 + side) * 9;
        o[0] = c.r; o[1] = c.y; o[2] = c.us; o[3] = c.uc; o[4] = c.w; o[5] = c.sq; o[6] = c.rs; o[7] = c.ri; o[8] = c.uf;
    }
    for (int e = threadIdx.x; e < nl * 4; e += blockDim.x) s.gacc[e] = T(0);
}

// One forward step of this thread's chunk (Godunov update, _macro_lane.py:83-114).
// u_first != nullptr only for a step whose cells carry an explicitly stored speed (set_r_u, step 0).
template <typename T, int C>
__device__ __forceinline__ bool chunk_fwd_step(T* r, T* y, const T* u_first, bool active, bool first_chunk,
                                               bool last_chunk, const LaneK<T>& k, const T* ghostL, const T* ghostR,
                                               T dt, T* boxL, T* boxR, int warp, int nwarp, unsigned lane) {
    Cell<T> rec[C];
#pragma unroll
    for (int c = 0; c < C; c++)
        rec[c] = u_first ? derive_cell_stored<T, false>(r[c], y[c], u_first[c], T(0), false, k.umax)
                         : derive_cell<T, false>(r[c], y[c], k.umax);
    T mine[RF_FWD], left[RF_FWD];
    pack_fwd(rec[C - 1], mine);
    from_left<T, RF_FWD>(mine, left, boxL, warp, lane);
    Cell<T> L = first_chunk ? unpack_adj(ghostL) : unpack_fwd(left);
    T fr[C + 1], fy[C + 1];
    bool bad = false;
#pragma unroll
    for (int c = 0; c < C; c++) {
        Riem<T> o = riemann(L, rec[c], k.umax, k.inv_umax, k.inv15, dt, k.dx);
        fr[c] = o.r0 * o.u0; fy[c] = o.y0 * o.u0;
        bad |= o.cfl_bad;
        L = rec[c];
    }
    T f0[2] = {fr[0], fy[0]}, fR[2];
    from_right<T, 2>(f0, fR, boxR, warp, nwarp, lane);
    if (last_chunk) {   // interface with the right ghost cell
        Riem<T> o = riemann(L, unpack_adj(ghostR), k.umax, k.inv_umax, k.inv15, dt, k.dx);
        fR[0] = o.r0 * o.u0; fR[1] = o.y0 * o.u0;
        bad |= o.cfl_bad;
    }
    fr[C] = fR[0]; fy[C] = fR[1];
#pragma unroll
    for (int c = 0; c < C; c++) {
        r[c] = r[c] + (fr[c] - fr[c + 1]) * k.cc;
        y[c] = y[c] + (fy[c] - fy[c + 1]) * k.cc;
    }
    return bad && active;
}

template <typename T, int C, int MB>
__global__ void __launch_bounds__(1024 / C, MB) arz_rollout_fwd_reg_kernel(const T* __restrict__ r0, const T* __restrict__ y0,
                                           const T* __restrict__ u0, const T* __restrict__ ghost,
                                           const T* __restrict__ dx, const T* __restrict__ umax_, T dt, int B, int N,
                                           int steps, int K, int lpc, T* __restrict__ ckpt, T* __restrict__ rT,
                                           T* __restrict__ yT, T* __restrict__ uT, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31;
    Shm<T> s = carve_shm<T>(raw, lpc, nwarp);
    const int tpl = N / C;                        // threads per lane
    const int l = threadIdx.x / tpl, kc = threadIdx.x - l * tpl;
    const int ngroup = (B + lpc - 1) / lpc;
    bool bad = false;
    for (int grp = blockIdx.x; grp < ngroup; grp += gridDim.x) {
        const int lane0 = grp * lpc, nl = min(lpc, B - lane0);
        const bool active = l < nl;
        const int ll = active ? l : 0;
        __syncthreads();
        setup_group(s, lane0, nl, ghost, dx, umax_, dt);
        __syncthreads();
        const LaneK<T> k = s.lk[ll];
        const T* gL = s.ghost + (ll * 2) * 9; const T* gR = gL + 9;
        const size_t off = (size_t)(lane0 + ll) * N + (size_t)kc * C;
        T r[C], y[C], us[C];
        load_chunk<T, C>(r0 + off, r); load_chunk<T, C>(y0 + off, y);
        if (u0) load_chunk<T, C>(u0 + off, us);
        for (int t = 0; t < steps; t++) {
            if (ckpt && t % K == 0 && active) {
                T* cr = ckpt + ((size_t)(t / K) * 2) * B * N + off;
                store_chunk<T, C>(cr, r); store_chunk<T, C>(cr + (size_t)B * N, y);
            }
            const int par = t & 1;
            bad |= chunk_fwd_step<T, C>(r, y, (t == 0 && u0) ? us : nullptr, active, kc == 0, kc == tpl - 1, k, gL, gR,
                                        dt, s.boxL + par * nwarp * RF_ADJ, s.boxR + par * nwarp * 4, warp, nwarp, lane);
        }
        if (active) {
            T u[C];
#pragma unroll
            for (int c = 0; c < C; c++) u[c] = (steps == 0 && u0) ? us[c] : compute_u(r[c], y[c], k.umax);
            store_chunk<T, C>(rT + off, r); store_chunk<T, C>(yT + off, y); store_chunk<T, C>(uT + off, u);
        }
    }
    if (bad) atomicOr(flags, FLAG_CFL);
}

template <typename T, int C, int MB>
__global__ void __launch_bounds__(1024 / C, MB) arz_rollout_bwd_reg_kernel(const T* __restrict__ ckpt, const T* __restrict__ u0,
                                           const T* __restrict__ ghost, const T* __restrict__ dx,
                                           const T* __restrict__ umax_, T dt, int B, int N, int steps, int K, int lpc,
                                           const T* __restrict__ rT, const T* __restrict__ yT,
                                           const T* __restrict__ g_rT, const T* __restrict__ g_yT,
                                           const T* __restrict__ g_uT, T* __restrict__ scratch,
                                           T* __restrict__ g_r0, T* __restrict__ g_y0, T* __restrict__ g_ghost,
                                           int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char raw[];
    const int nwarp = blockDim.x >> 5, warp = threadIdx.x >> 5;
    const unsigned lane = threadIdx.x & 31;
    Shm<T> s = carve_shm<T>(raw, lpc, nwarp);
    const int tpl = N / C;
    const int l = threadIdx.x / tpl, kc = threadIdx.x - l * tpl;
    const int ngroup = (B + lpc - 1) / lpc;
    const int S = (steps + K - 1) / K;
    const size_t sstride = (size_t)2 * lpc * N;               // one stashed state of the group
    T* scr = scratch + (size_t)blockIdx.x * K * sstride;
    bool bad = false;
    for (int grp = blockIdx.x; grp < ngroup; grp += gridDim.x) {
        const int lane0 = grp * lpc, nl = min(lpc, B - lane0);
        const bool active = l < nl;
        const int ll = active ? l : 0;
        const bool first_chunk = kc == 0, last_chunk = kc == tpl - 1;
        __syncthreads();
        setup_group(s, lane0, nl, ghost, dx, umax_, dt);
        __syncthreads();
        const LaneK<T> k = s.lk[ll];
        const T* gL = s.ghost + (ll * 2) * 9; const T* gR = gL + 9;
        const size_t off = (size_t)(lane0 + ll) * N + (size_t)kc * C;
        const size_t soff = (size_t)ll * N + (size_t)kc * C;
        T gr[C], gy[C], us[C];
        T accL[2] = {T(0), T(0)}, accR[2] = {T(0), T(0)};     // ghost adjoints of this lane (first / last chunk)
        // terminal adjoint; uT = compute_u(rT, yT) is produced inside the operator
#pragma unroll
        for (int c = 0; c < C; c++) {
            gr[c] = g_rT ? g_rT[off + c] : T(0); gy[c] = g_yT ? g_yT[off + c] : T(0);
            if (g_uT) {
                T dr, dy; du_dry(rT[off + c], yT[off + c], k.umax, dr, dy);
                T gu = g_uT[off + c]; gr[c] += gu * dr; gy[c] += gu * dy;
            }
        }
        if (u0) load_chunk<T, C>(u0 + off, us);
        for (int seg = S - 1; seg >= 0; seg--) {
            const int t0 = seg * K, ks = min(K, steps - t0);
            T r[C], y[C];
            {
                const T* cr = ckpt + ((size_t)seg * 2) * B * N + off;
                load_chunk<T, C>(cr, r); load_chunk<T, C>(cr + (size_t)B * N, y);
            }
            // recompute the segment, stashing every state (the stash stays in L2)
            for (int kk = 0; kk < ks; kk++) {
                T* sr = scr + (size_t)kk * sstride + soff;
                if (active) { store_chunk<T, C>(sr, r); store_chunk<T, C>(sr + (size_t)lpc * N, y); }
                if (kk + 1 < ks) {
                    const int par = kk & 1;
                    chunk_fwd_step<T, C>(r, y, (t0 + kk == 0 && u0) ? us : nullptr, active, first_chunk, last_chunk, k,
                                         gL, gR, dt, s.boxL + par * nwarp * RF_ADJ, s.boxR + par * nwarp * 4, warp,
                                         nwarp, lane);
                }
            }
            // adjoint steps, last to first
            for (int kk = ks - 1; kk >= 0; kk--) {
                const int par = kk & 1;
                T* boxL = s.boxL + par * nwarp * RF_ADJ; T* boxR = s.boxR + par * nwarp * 4;
                if (kk != ks - 1) {
                    const T* sr = scr + (size_t)kk * sstride + soff;
                    load_chunk<T, C>(sr, r); load_chunk<T, C>(sr + (size_t)lpc * N, y);
                }
                const bool stored = (t0 + kk == 0 && u0);
                Cell<T> rec[C];
#pragma unroll
                for (int c = 0; c < C; c++)
                    rec[c] = stored ? derive_cell_stored<T, true>(r[c], y[c], us[c], T(0), false, k.umax)
                                    : derive_cell<T, true>(r[c], y[c], k.umax);
                T mine[RF_ADJ], left[RF_ADJ];
                pack_adj(rec[C - 1], gr[C - 1], gy[C - 1], mine);
                from_left<T, RF_ADJ>(mine, left, boxL, warp, lane);
                Cell<T> L = first_chunk ? unpack_adj(gL) : unpack_adj(left);
                T gLr = first_chunk ? T(0) : left[9], gLy = first_chunk ? T(0) : left[10];
                // sweep my C left-interfaces: pa -> cell on the left, pb -> my cell
                T ar[C + 1], ay[C + 1], br[C], by[C];
#pragma unroll
                for (int c = 0; c < C; c++) {
                    Riem<T> o = riemann(L, rec[c], k.umax, k.inv_umax, k.inv15, dt, k.dx);
                    riemann_adj(L, rec[c], o, k.umax, k.inv_umax, k.inv15, gr[c] - gLr, gy[c] - gLy, ar[c], ay[c],
                                br[c], by[c]);
                    L = rec[c]; gLr = gr[c]; gLy = gy[c];
                }
                T a0[2] = {ar[0], ay[0]}, aR[2];
                from_right<T, 2>(a0, aR, boxR, warp, nwarp, lane);
                if (last_chunk) {   // interface with the right ghost (its updated-state adjoint is zero)
                    Cell<T> G = unpack_adj(gR);
                    Riem<T> o = riemann(L, G, k.umax, k.inv_umax, k.inv15, dt, k.dx);
                    T pbr, pby;
                    riemann_adj(L, G, o, k.umax, k.inv_umax, k.inv15, -gLr, -gLy, aR[0], aR[1], pbr, pby);
                    accR[0] += k.cc * pbr; accR[1] += k.cc * pby;
                }
                if (first_chunk) { accL[0] += k.cc * ar[0]; accL[1] += k.cc * ay[0]; }
                ar[C] = aR[0]; ay[C] = aR[1];
#pragma unroll
                for (int c = 0; c < C; c++) {
                    gr[c] = gr[c] + k.cc * (ar[c + 1] + br[c]);
                    gy[c] = gy[c] + k.cc * (ay[c + 1] + by[c]);
                    bad |= active && (t_isnan(gr[c]) || t_isnan(gy[c]));
                }
            }
        }
        if (active) {
            store_chunk<T, C>(g_r0 + off, gr); store_chunk<T, C>(g_y0 + off, gy);
            if (g_ghost) {
                T* gg = g_ghost + (size_t)(lane0 + ll) * 4;
                if (first_chunk) { gg[0] = accL[0]; gg[1] = accL[1]; bad |= t_isnan(accL[0]) || t_isnan(accL[1]); }
                if (last_chunk) { gg[2] = accR[0]; gg[3] = accR[1]; bad |= t_isnan(accR[0]) || t_isnan(accR[1]); }
            }
        }
    }
    if (bad) atomicOr(flags, FLAG_NAN_GRAD);
}

// ------------------------------------------------------------------ host-side planning and launch

struct RegPlan { int C, lpc, threads, grid, mb; size_t smem; };

static int round32(int x) { return (x + 31) / 32 * 32; }

static int sm_count_r() {
    static int n = 0;
    if (!n) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}

// Measured on B200 (profiles/r1c): forward is fastest with 4 cells per thread capped at 128 registers
// (2 CTAs of 256 threads per SM); the adjoint holds more live state and is fastest with 2 cells per thread.
template <typename T> static int plan_reg(int B, int N, bool adj, RegPlan* p) {
    int want = adj ? 2 : 4;
    if (const char* e = getenv("DHTS_ARZ_C")) {      // tuning knob: cells per thread (must divide N)
        int c = atoi(e);
        if (c == 1 || c == 2 || c == 4) want = c;
    }
    int C = (want >= 4 && N % 4 == 0) ? 4 : ((want >= 2 && N % 2 == 0) ? 2 : 1);
    int tpl = N / C;
    if (tpl > 1024 / C && C > 1) { C = (N % 4 == 0) ? 4 : C; tpl = N / C; }   // long lanes: widest chunk that divides N
    if (tpl > 1024 / C) return DHTS_ERR_UNSUPPORTED;
    const int target = 256;                       // threads per CTA when lanes are short
    int lpc = tpl >= target ? 1 : target / tpl;
    if (lpc > B) lpc = B;
    if (lpc < 1) lpc = 1;
    p->C = C; p->lpc = lpc; p->threads = round32(lpc * tpl);
    p->smem = shm_bytes<T>(lpc, p->threads / 32);
    p->grid = (B + lpc - 1) / lpc;
    p->mb = adj ? 1 : 2;
    if (const char* e = getenv("DHTS_ARZ_MB")) p->mb = atoi(e) == 2 ? 2 : 1;
    return DHTS_OK;
}

#define DHTS_C_DISPATCH(P, CALL)                                                                     \
    if ((P).mb == 2) {                                                                               \
        if ((P).C == 4) { CALL(4, 2) } else if ((P).C == 2) { CALL(2, 2) } else { CALL(1, 1) }       \
    } else {                                                                                         \
        if ((P).C == 4) { CALL(4, 1) } else if ((P).C == 2) { CALL(2, 1) } else { CALL(1, 1) }       \
    }

static int status_r() { return cudaGetLastError() == cudaSuccess ? DHTS_OK : DHTS_ERR_CUDA; }

template <typename T>
static int rollout_fwd(const T* r0, const T* y0, const T* u0, const T* ghost, const T* dx, const T* umax, T dt, int B,
                       int N, int steps, int K, T* ckpt, T* rT, T* yT, T* uT, int* flags, cudaStream_t st) {
    if (!r0 || !y0 || !ghost || !dx || !umax || !rT || !yT || !uT || !flags || B < 0 || N < 1 || steps < 0)
        return DHTS_ERR_INVALID;
    if (ckpt && K < 1) return DHTS_ERR_INVALID;
    if (B == 0) return DHTS_OK;
    RegPlan p;
    int rc = plan_reg<T>(B, N, false, &p);
    if (rc) return rc;
    if (K < 1) K = 1;
    int grid = p.grid;
#define CALL(CC, MB) arz_rollout_fwd_reg_kernel<T, CC, MB><<<grid, p.threads, p.smem, st>>>(r0, y0, u0, ghost, dx, umax, dt, B, N, steps, K, p.lpc, ckpt, rT, yT, uT, flags);
    DHTS_C_DISPATCH(p, CALL)
#undef CALL
    return status_r();
}

template <typename T> static int bwd_grid_r(const RegPlan& p) {
    int occ = 0;
#define CALL(CC, MB) cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, arz_rollout_bwd_reg_kernel<T, CC, MB>, p.threads, p.smem);
    DHTS_C_DISPATCH(p, CALL)
#undef CALL
    if (occ < 1) occ = 1;
    long long g = (long long)sm_count_r() * occ;
    return (int)(g < p.grid ? g : p.grid);
}

template <typename T> static long long rollout_scratch_elems(int B, int N, int K) {
    RegPlan p;
    if (B <= 0) return 0;
    if (K < 1 || plan_reg<T>(B, N, true, &p)) return -1;
    return (long long)bwd_grid_r<T>(p) * K * 2 * p.lpc * N;
}

template <typename T>
static int rollout_bwd(const T* ckpt, const T* u0, const T* ghost, const T* dx, const T* umax, T dt, int B, int N,
                       int steps, int K, const T* rT, const T* yT, const T* g_rT, const T* g_yT, const T* g_uT,
                       T* scratch, long long scratch_elems, T* g_r0, T* g_y0, T* g_ghost, int* flags,
                       cudaStream_t st) {
    if (!ckpt || !ghost || !dx || !umax || !g_r0 || !g_y0 || !flags || !scratch || B < 0 || N < 1 || steps < 0 || K < 1)
        return DHTS_ERR_INVALID;
    if (g_uT && (!rT || !yT)) return DHTS_ERR_INVALID;
    if (B == 0) return DHTS_OK;
    RegPlan p;
    int rc = plan_reg<T>(B, N, true, &p);
    if (rc) return rc;
    int grid = bwd_grid_r<T>(p);
    if ((long long)grid * K * 2 * p.lpc * N > scratch_elems) return DHTS_ERR_INVALID;
#define CALL(CC, MB) arz_rollout_bwd_reg_kernel<T, CC, MB><<<grid, p.threads, p.smem, st>>>(ckpt, u0, ghost, dx, umax, dt, B, N, steps, K, p.lpc, rT, yT, g_rT, g_yT, g_uT, scratch, g_r0, g_y0, g_ghost, flags);
    DHTS_C_DISPATCH(p, CALL)
#undef CALL
    return status_r();
}

}  // namespace dhts

#define DHTS_ARZ_ROLLOUT_API(SUF, T)                                                                                   \
    DHTS_EXPORT int dhts_arz_rollout_fwd_##SUF(const T* r0, const T* y0, const T* u0, const T* ghost, const T* dx,     \
                                               const T* umax, T dt, int B, int N, int steps, int ckpt_every, T* ckpt,  \
                                               T* rT, T* yT, T* uT, int* flags, void* stream) {                        \
        return dhts::rollout_fwd<T>(r0, y0, u0, ghost, dx, umax, dt, B, N, steps, ckpt_every, ckpt, rT, yT, uT, flags, \
                                    (cudaStream_t)stream);                                                             \
    }                                                                                                                  \
    DHTS_EXPORT long long dhts_arz_rollout_scratch_elems_##SUF(int B, int N, int ckpt_every) {                         \
        return dhts::rollout_scratch_elems<T>(B, N, ckpt_every);                                                       \
    }                                                                                                                  \
    DHTS_EXPORT int dhts_arz_rollout_bwd_##SUF(const T* ckpt, const T* u0, const T* ghost, const T* dx, const T* umax, \
                                               T dt, int B, int N, int steps, int ckpt_every, const T* rT,             \
                                               const T* yT, const T* g_rT, const T* g_yT, const T* g_uT, T* scratch,   \
                                               long long scratch_elems, T* g_r0, T* g_y0, T* g_ghost, int* flags,      \
                                               void* stream) {                                                         \
        return dhts::rollout_bwd<T>(ckpt, u0, ghost, dx, umax, dt, B, N, steps, ckpt_every, rT, yT, g_rT, g_yT, g_uT,  \
                                    scratch, scratch_elems, g_r0, g_y0, g_ghost, flags, (cudaStream_t)stream);         \
    }

DHTS_ARZ_ROLLOUT_API(f64, double)
DHTS_ARZ_ROLLOUT_API(f32, float)
