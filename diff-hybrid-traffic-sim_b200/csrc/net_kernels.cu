// Fused T-step rollout of a CONNECTED macro network (many replicas per launch), forward and adjoint.
//
// What it replaces in the reference, per simulation step and per replica:
//   RoadNetwork.forward                       road/network/road_network.py:79-111   (boundaries -> forward -> update, Jacobi)
//   RoadNetwork.get_macro_boundary            road/network/road_network.py:299-362  (ghost source: the only neighbour, or the
//                                                                                   lane the MacroRoute of this step selects,
//                                                                                   else the lane's own ghost cell)
//   ItscpRoadNetwork.setup_macro_boundary     example/control/itscp/_simulator.py:56-142  (signal blend of the ghost cells)
//   MacroLane.set_{left,right}most_cell       road/lane/_macro_lane.py:156-162 -> ARZ.FullQ.from_r_u (_arz.py:74-80)
//   dMacroLane.forward / dMacroForwardLayer   road/lane/dmacro_lane.py:68-132,234-310
//   queue-length reward (optional)            example/control/itscp/_env.py:662-742,770-797
// and the autograd chain through all of it.
//
// Shape of the work: a network is L lanes of a few cells each (ITSCP: 144 lanes of 1-4 cells), coupled every step
// through ghost cells; replicas (scenarios, candidate signal plans) are independent.  One CTA steps one replica:
// the whole network state lives in shared memory for the T steps, one thread per lane (its cells are swept
// serially with the same per-interface math as the lane kernels, dhts_arz.cuh), one block barrier per step.
// Long lanes belong to the lane-batched rollout (arz_rollout.cu); this kernel is for many short coupled lanes.
//
// Ghost cells.  side 0 = left (upstream), 1 = right.  green = (r, u) of the source; the ITSCP blend is
//   left : final = green * s + (0, u_max) * (1 - s),  s = 1 when the lane has no predecessor (green = (incoming,
//          u_eq(incoming))), 0 when the step's MacroRoute gives it no predecessor, else the RAW signal of that lane
//   right: final = s * green + (1 - s) * (1, 0),      s = sigmoid(clamp(32 (sig - 0.5), -16, 16))  [soft]
//                                                     s = sig > 0.5                                 [hard]
// and the lane's own ghost record is overwritten with `final` every step, so a side without a neighbour feeds on
// its own previous value (a recurrence carried in `own`).  mode 0 (plain RoadNetwork): final = green.
//
// Adjoint.  Every state is stored by the forward pass ([T+1][R][3][NC]); the adjoint walks the steps backwards:
// compute_u's true derivative folds the adjoint of the stored speed into (r, y); each lane thread rebuilds its two
// ghosts, runs the flux-difference adjoint over its cells, pulls the ghost adjoint back through from_r_u and the
// blend, and publishes (d green, d signal) per side in shared memory; after one barrier each lane GATHERS what its
// neighbours published for its edge cells and for its signal (fixed order: deterministic, no atomics).
#include <cuda_pipeline.h>
#include "dhts_net.cuh"

namespace dhts {

// Thread -> lane assignment: lanes sorted by their number of cells (stable), so that the threads of a warp sweep lanes
// of equal length (an ITSCP grid mixes 1-, 2- and 4-cell lanes: in lane-id order a warp ran at 56 % thread efficiency,
// profiles/r1n_net_ncu_summary.json).  order[rank] = lane; every thread calls it; ends with a block barrier.
template <typename T> __device__ __forceinline__ void lane_order(const NetArgs<T>& a, int* order) {
    for (int l = threadIdx.x; l < a.L; l += blockDim.x) {
        const int n = a.cell_off[l + 1] - a.cell_off[l];
        int rank = 0;
        for (int j = 0; j < a.L; j++) {
            const int m = a.cell_off[j + 1] - a.cell_off[j];
            rank += (m > n) || (m == n && j < l);
        }
        order[rank] = l;
    }
    __syncthreads();
}

// ------------------------------------------------------------------------------------------------ forward
// hist  [T+1][R][3][NC]  state (r, y, u) before step t (t = 0..T-1) and after the last step
// ownh  [T+1][R][n_own][2] carried own-ghost records, same indexing
template <typename T>
__global__ void __launch_bounds__(256) net_rollout_fwd_kernel(NetArgs<T> a, const T* __restrict__ r0, const T* __restrict__ y0,
                                                               const T* __restrict__ u0, const T* __restrict__ own0,
                                                               T* __restrict__ hist, T* __restrict__ ownh,
                                                               T* __restrict__ reward, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char raw[];
    T* sm = reinterpret_cast<T*>(raw);
    const int NC = a.NC, L = a.L;
    T* buf[2] = {sm, sm + 3 * NC};                         // (r, y, u) x NC, double buffered
    T* own[2] = {sm + 6 * NC, sm + 6 * NC + 2 * a.n_own};
    T* red = sm + 6 * NC + 4 * a.n_own;                    // [blockDim] reward reduction
    int* order = reinterpret_cast<int*>(red + blockDim.x);  // [L] lanes by decreasing number of cells
    const T inv_umax = T(1) / a.umax, inv15 = T(1) / (T(1.5) * a.umax);
    lane_order(a, order);
    for (int b = blockIdx.x; b < a.R; b += gridDim.x) {
        __syncthreads();
        for (int c = threadIdx.x; c < NC; c += blockDim.x) {
            buf[0][c] = r0[(size_t)b * NC + c]; buf[0][NC + c] = y0[(size_t)b * NC + c]; buf[0][2 * NC + c] = u0[(size_t)b * NC + c];
        }
        for (int c = threadIdx.x; c < 2 * a.n_own; c += blockDim.x) own[0][c] = own0[(size_t)b * 2 * a.n_own + c];
        __syncthreads();
        bool bad = false, bad_route = false;
        T rew = T(0);
        int p = 0;
        for (int t = 0; t <= a.T_steps; t++) {
            const T* cr = buf[p]; const T* cy = cr + NC; const T* cu = cy + NC;
            {   // store the state before step t (coalesced)
                T* h = hist + ((size_t)t * a.R + b) * 3 * NC;
                for (int c = threadIdx.x; c < 3 * NC; c += blockDim.x) h[c] = cr[c];
                T* oh = ownh + ((size_t)t * a.R + b) * 2 * a.n_own;
                for (int c = threadIdx.x; c < 2 * a.n_own; c += blockDim.x) oh[c] = own[p][c];
            }
            if (t == a.T_steps) break;
            T* nr = buf[p ^ 1]; T* ny = nr + NC; T* nu = ny + NC;
            const int* rt = a.route ? a.route + (size_t)b * a.route_stride + (size_t)t * 2 * L : nullptr;
            const T* sig_t = a.sig ? a.sig + ((size_t)b * a.T_steps + t) * L : nullptr;
            const T* inc_t = a.incoming ? a.incoming + ((size_t)b * a.T_steps + t) * L : nullptr;
            for (int li = threadIdx.x; li < L; li += blockDim.x) {
                const int l = order[li];
                const int c0 = a.cell_off[l], N = a.cell_off[l + 1] - c0;
                const T dxl = a.dx[l], cc = a.dt / dxl;
                const Side<T> sl = resolve_side(a, l, 0, rt, cr, cu, own[p], sig_t, inc_t, bad_route);
                const Side<T> sr = resolve_side(a, l, 1, rt, cr, cu, own[p], sig_t, inc_t, bad_route);
                const int osl = a.own_slot[l], osr = a.own_slot[L + l];
                if (osl >= 0) { own[p ^ 1][2 * osl] = sl.fr; own[p ^ 1][2 * osl + 1] = sl.fu; }
                if (osr >= 0) { own[p ^ 1][2 * osr] = sr.fr; own[p ^ 1][2 * osr + 1] = sr.fu; }
                Cell<T> Lc = ghost_cell<T, false>(sl.fr, sl.fu, a.umax);
                T fpr = T(0), fpy = T(0), q = T(0);
                for (int i = 0; i <= N; i++) {
                    const Cell<T> Rc = (i < N) ? derive_cell_stored<T, false>(cr[c0 + i], cy[c0 + i], cu[c0 + i], T(0), false, a.umax)
                                               : ghost_cell<T, false>(sr.fr, sr.fu, a.umax);
                    const Riem<T> o = riemann(Lc, Rc, a.umax, inv_umax, inv15, a.dt, dxl);
                    bad |= o.cfl_bad;
                    const T fr = o.r0 * o.u0, fy = o.y0 * o.u0;
                    if (i > 0) {
                        const T r = cr[c0 + i - 1] + (fpr - fr) * cc;
                        const T y = cy[c0 + i - 1] + (fpy - fy) * cc;
                        const T u = compute_u(r, y, a.umax);
                        nr[c0 + i - 1] = r; ny[c0 + i - 1] = y; nu[c0 + i - 1] = u;
                        if (a.qk) {
                            T z = (a.static_speed - u) * a.qk[t];
                            z = z < T(-16) ? T(-16) : (z > T(16) ? T(16) : z);
                            q += sigm(z) * (r * dxl / a.veh_len);
                        }
                    }
                    fpr = fr; fpy = fy; Lc = Rc;
                }
                if (a.qk) rew -= q * q * a.dt;
            }
            __syncthreads();
            p ^= 1;
        }
        if (reward) {   // fixed-order tree reduction of the per-thread partial sums
            red[threadIdx.x] = rew;
            __syncthreads();
            for (int s = 1; s < (int)blockDim.x; s <<= 1) {
                if ((threadIdx.x & (2 * s - 1)) == 0 && threadIdx.x + s < blockDim.x) red[threadIdx.x] += red[threadIdx.x + s];
                __syncthreads();
            }
            if (threadIdx.x == 0) reward[b] = red[0];
        }
        if (bad) atomicOr(flags, FLAG_CFL);
        if (bad_route) atomicOr(flags, FLAG_ROUTE);
    }
}

// ------------------------------------------------------------------------------------------------ adjoint
// g_states [T][R][3][NC]  optional: dLoss/d(r, y, u) of the state AFTER step t (the last one is the terminal adjoint)
// g_reward [R]            optional: dLoss/d reward (fused queue reward)
// outputs: g_r0, g_y0, g_u0 [R][NC]; g_own0 [R][n_own][2]; g_sig, g_inc [R][T][L]
template <typename T>
__global__ void __launch_bounds__(256, 2) net_rollout_bwd_kernel(NetArgs<T> a, const T* __restrict__ hist, const T* __restrict__ ownh,
                                                               const T* __restrict__ g_states, const T* __restrict__ g_reward,
                                                               T* __restrict__ g_r0, T* __restrict__ g_y0, T* __restrict__ g_u0,
                                                               T* __restrict__ g_own0, T* __restrict__ g_sig,
                                                               T* __restrict__ g_inc, int* __restrict__ flags) {
    extern __shared__ __align__(16) unsigned char raw[];
    T* sm = reinterpret_cast<T*>(raw);
    const int NC = a.NC, L = a.L;
    // three rotating state buffers: state t (cur), state t + 1 (nxt) and the row being prefetched for the next step
    // (pre); the stored rows stream in with cp.async (LDGSTS) one step ahead of the arithmetic
    T* cur = sm;                        // state t      (r, y, u)
    T* nxt = sm + 3 * NC;               // state t + 1  (r, y, u)
    T* pre = sm + 6 * NC;               // state t - 1, in flight
    T* G = sm + 9 * NC;                 // adjoint of state t+1 on entry of a step, of state t on exit: (gr, gy, gu)
    T* own = sm + 12 * NC;              // own records at step t
    T* own_pre = own + 2 * a.n_own;     // own records at step t - 1, in flight
    T* GO = own_pre + 2 * a.n_own;      // adjoint of the own records
    T* pub = GO + 2 * a.n_own;          // [L][2 sides][3] published (d green r, d green u, d signal)
    int* order = reinterpret_cast<int*>(pub + (size_t)6 * L);      // [L] lanes by decreasing number of cells
    const T inv_umax = T(1) / a.umax, inv15 = T(1) / (T(1.5) * a.umax);
    lane_order(a, order);
    for (int b = blockIdx.x; b < a.R; b += gridDim.x) {
        __syncthreads();
        {   // terminal state and adjoint
            const T* h = hist + ((size_t)a.T_steps * a.R + b) * 3 * NC;
            for (int c = threadIdx.x; c < 3 * NC; c += blockDim.x) {
                nxt[c] = h[c];
                G[c] = (g_states && a.T_steps > 0) ? g_states[((size_t)(a.T_steps - 1) * a.R + b) * 3 * NC + c] : T(0);
            }
            for (int c = threadIdx.x; c < 2 * a.n_own; c += blockDim.x) GO[c] = T(0);
        }
        const T grew = g_reward ? g_reward[b] : T(0);
        bool nan = false;
#define DHTS_NET_PREFETCH(TT, DST, ODST)                                                                               \
        {                                                                                                              \
            const T* h_ = hist + ((size_t)(TT) * a.R + b) * 3 * NC;                                                    \
            for (int c = threadIdx.x; c < 3 * NC; c += blockDim.x) __pipeline_memcpy_async((DST) + c, h_ + c, sizeof(T)); \
            const T* oh_ = ownh + ((size_t)(TT) * a.R + b) * 2 * a.n_own;                                              \
            for (int c = threadIdx.x; c < 2 * a.n_own; c += blockDim.x) __pipeline_memcpy_async((ODST) + c, oh_ + c, sizeof(T)); \
            __pipeline_commit();                                                                                       \
        }
        if (a.T_steps > 0) DHTS_NET_PREFETCH(a.T_steps - 1, cur, own)
        for (int t = a.T_steps - 1; t >= 0; t--) {
            __pipeline_wait_prior(0);
            __syncthreads();          // state t has landed; the previous step's gathers are complete
            if (t > 0) DHTS_NET_PREFETCH(t - 1, pre, own_pre)          // pre was state t + 2: nobody reads it any more
            const T* cr = cur; const T* cy = cur + NC; const T* cu = cur + 2 * NC;
            const int* rt = a.route ? a.route + (size_t)b * a.route_stride + (size_t)t * 2 * L : nullptr;
            const T* sig_t = a.sig ? a.sig + ((size_t)b * a.T_steps + t) * L : nullptr;
            const T* inc_t = a.incoming ? a.incoming + ((size_t)b * a.T_steps + t) * L : nullptr;
            bool dummy = false;
            for (int li = threadIdx.x; li < L; li += blockDim.x) {
                const int l = order[li];
                const int c0 = a.cell_off[l], N = a.cell_off[l + 1] - c0;
                const T dxl = a.dx[l], cc = a.dt / dxl;
                // (1) fused queue reward of state t+1 and the stored speed's adjoint folded into (r, y)
                if (a.qk && g_reward) {
                    T q = T(0);
                    for (int i = 0; i < N; i++) {
                        T z = (a.static_speed - nxt[2 * NC + c0 + i]) * a.qk[t];
                        z = z < T(-16) ? T(-16) : (z > T(16) ? T(16) : z);
                        q += sigm(z) * (nxt[c0 + i] * dxl / a.veh_len);
                    }
                    const T gq = -T(2) * q * a.dt * grew;
                    for (int i = 0; i < N; i++) {
                        const T zr = (a.static_speed - nxt[2 * NC + c0 + i]) * a.qk[t];
                        const bool in = zr >= T(-16) && zr <= T(16);
                        const T z = zr < T(-16) ? T(-16) : (zr > T(16) ? T(16) : zr);
                        const T sg = sigm(z), w = dxl / a.veh_len;
                        G[c0 + i] += gq * sg * w;
                        if (in) G[2 * NC + c0 + i] += gq * nxt[c0 + i] * w * sg * (T(1) - sg) * (-a.qk[t]);
                    }
                }
                for (int i = 0; i < N; i++) {
                    T dr, dy; du_dry(nxt[c0 + i], nxt[NC + c0 + i], a.umax, dr, dy);
                    const T gu = G[2 * NC + c0 + i];
                    G[c0 + i] += gu * dr; G[NC + c0 + i] += gu * dy;
                }
                // (2) ghosts of step t and the flux-difference adjoint over the lane's interfaces
                const Side<T> sl = resolve_side(a, l, 0, rt, cr, cu, own, sig_t, inc_t, dummy);
                const Side<T> sr = resolve_side(a, l, 1, rt, cr, cu, own, sig_t, inc_t, dummy);
                const Cell<T> gL = ghost_cell<T, true>(sl.fr, sl.fu, a.umax);
                const Cell<T> gR = ghost_cell<T, true>(sr.fr, sr.fu, a.umax);
                Cell<T> Lc = gL;
                T gLr = T(0), gLy = T(0);            // old adjoint of the cell left of the interface (ghost: 0)
                T pbr = T(0), pby = T(0);            // B^T w of the previous interface
                T ggl_r = T(0), ggl_y = T(0), ggr_r = T(0), ggr_y = T(0);
                for (int i = 0; i <= N; i++) {
                    const Cell<T> Rc = (i < N) ? derive_cell_stored<T, true>(cr[c0 + i], cy[c0 + i], cu[c0 + i], T(0), false, a.umax) : gR;
                    const T gRr = (i < N) ? G[c0 + i] : T(0), gRy = (i < N) ? G[NC + c0 + i] : T(0);
                    const Riem<T> o = riemann(Lc, Rc, a.umax, inv_umax, inv15, a.dt, dxl);
                    T par, pay, qbr, qby;
                    riemann_adj(Lc, Rc, o, a.umax, inv_umax, inv15, gRr - gLr, gRy - gLy, par, pay, qbr, qby);
                    if (i == 0) { ggl_r = cc * par; ggl_y = cc * pay; }
                    else {
                        const T nr_ = gLr + cc * (par + pbr), ny_ = gLy + cc * (pay + pby);
                        nan |= t_isnan(nr_) || t_isnan(ny_);
                        G[c0 + i - 1] = nr_; G[NC + c0 + i - 1] = ny_;
                    }
                    if (i == N) { ggr_r = cc * qbr; ggr_y = cc * qby; }
                    pbr = qbr; pby = qby; gLr = gRr; gLy = gRy; Lc = Rc;
                }
                for (int i = 0; i < N; i++) G[2 * NC + c0 + i] = T(0);      // the stored speed of state t: filled by the gathers
                // (3) ghost adjoint -> from_r_u -> blend -> (d green, d signal) per side
                for (int side = 0; side < 2; side++) {
                    const Side<T>& s = side == 0 ? sl : sr;
                    const T g_r = side == 0 ? ggl_r : ggr_r, g_y = side == 0 ? ggl_y : ggr_y;
                    const T ue = u_eq(s.fr, a.umax);
                    T gfr = g_r + g_y * (s.fu - ue - s.fr * u_eq_true_prime(s.fr, a.umax));
                    T gfu = g_y * s.fr;
                    const int os = a.own_slot[side * L + l];
                    if (os >= 0) { gfr += GO[2 * os]; gfu += GO[2 * os + 1]; }       // own_{t+1} = final_t
                    T ggr = gfr, ggu = gfu, gs = T(0);
                    if (a.mode == 1) {
                        const T red_r = side == 0 ? T(0) : T(1), red_u = side == 0 ? a.umax : T(0);
                        ggr = gfr * s.s; ggu = gfu * s.s;
                        gs = gfr * (s.gr_ - red_r) + gfu * (s.gu_ - red_u);
                        if (side == 1) gs = a.soft ? gs * T(32) * s.s * (T(1) - s.s) : T(0);
                        if (s.sig_lane < 0) gs = T(0);
                    }
                    if (os >= 0) {
                        const bool own_src = s.src == -1;
                        GO[2 * os] = own_src ? ggr : T(0); GO[2 * os + 1] = own_src ? ggu : T(0);
                    }
                    if (s.src == -2 && g_inc)
                        g_inc[((size_t)b * a.T_steps + t) * L + l] = ggr + ggu * u_eq_true_prime(s.gr_, a.umax);
                    T* pb = pub + ((size_t)l * 2 + side) * 3;
                    pb[0] = s.src >= 0 ? ggr : T(0); pb[1] = s.src >= 0 ? ggu : T(0); pb[2] = gs;
                }
                if (g_inc && !(a.mode == 1 && a.nadj[l] == 0)) g_inc[((size_t)b * a.T_steps + t) * L + l] = T(0);
            }
            __syncthreads();
            // (4) gathers: what the neighbours took from this lane's edge cells and from its signal
            for (int li = threadIdx.x; li < L; li += blockDim.x) {
                const int l = order[li];
                const int c0 = a.cell_off[l], N = a.cell_off[l + 1] - c0;
                T gsig = pub[((size_t)l * 2 + 1) * 3 + 2];
                // successors whose LEFT ghost came from this lane's last cell / was blended by this lane's signal
                for (int e = a.adj_off[(L + 1) + l]; e < a.adj_off[(L + 1) + l + 1]; e++) {
                    const int n = a.adj[e];
                    const int cnt = a.nadj[n];
                    const int sel = rt ? rt[n] : -1;
                    const int srcn = cnt == 1 ? a.one_adj[n] : (cnt > 1 ? sel : -1);
                    const T* pb = pub + ((size_t)n * 2 + 0) * 3;
                    if (srcn == l) { G[c0 + N - 1] += pb[0]; G[2 * NC + c0 + N - 1] += pb[1]; }
                    if (a.mode == 1 && sel == l) gsig += pb[2];
                }
                // predecessors whose RIGHT ghost came from this lane's first cell
                for (int e = a.adj_off[l]; e < a.adj_off[l + 1]; e++) {
                    const int pl = a.adj[e];
                    const int cnt = a.nadj[L + pl];
                    const int sel = rt ? rt[L + pl] : -1;
                    const int srcp = cnt == 1 ? a.one_adj[L + pl] : (cnt > 1 ? sel : -1);
                    const T* pb = pub + ((size_t)pl * 2 + 1) * 3;
                    if (srcp == l) { G[c0] += pb[0]; G[2 * NC + c0] += pb[1]; }
                }
                if (g_sig) g_sig[((size_t)b * a.T_steps + t) * L + l] = gsig;
                nan |= t_isnan(gsig);
                // (5) injected adjoint of state t (it is the state after step t - 1)
                if (g_states && t > 0) {
                    const T* gs_ = g_states + ((size_t)(t - 1) * a.R + b) * 3 * NC;
                    for (int i = 0; i < N; i++) {
                        G[c0 + i] += gs_[c0 + i]; G[NC + c0 + i] += gs_[NC + c0 + i]; G[2 * NC + c0 + i] += gs_[2 * NC + c0 + i];
                    }
                }
            }
            { T* x_ = nxt; nxt = cur; cur = pre; pre = x_; x_ = own; own = own_pre; own_pre = x_; }     // rotate: no copy, no barrier
        }
#undef DHTS_NET_PREFETCH
        __syncthreads();
        for (int c = threadIdx.x; c < NC; c += blockDim.x) {
            g_r0[(size_t)b * NC + c] = G[c]; g_y0[(size_t)b * NC + c] = G[NC + c]; g_u0[(size_t)b * NC + c] = G[2 * NC + c];
            nan |= t_isnan(G[c]) || t_isnan(G[NC + c]) || t_isnan(G[2 * NC + c]);
        }
        if (g_own0)
            for (int c = threadIdx.x; c < 2 * a.n_own; c += blockDim.x) g_own0[(size_t)b * 2 * a.n_own + c] = GO[c];
        if (nan) atomicOr(flags, FLAG_NAN_GRAD);
    }
}

// ------------------------------------------------------------------------------------------------ host side
static int net_sm_count() {
    static int n = 0;
    if (!n) {
        int dev = 0; cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        if (n <= 0) n = 148;
    }
    return n;
}
static int net_threads(int L) { int t = (L + 31) / 32 * 32; return t > 256 ? 256 : (t < 32 ? 32 : t); }

template <typename T> static size_t net_smem(int L, int NC, int n_own, bool adj, int threads) {
    return sizeof(T) * (adj ? ((size_t)12 * NC + 6 * n_own + (size_t)6 * L) : ((size_t)6 * NC + 4 * n_own + threads)) + sizeof(int) * (size_t)L + 16;
}

template <typename T>
static int net_check(const NetArgs<T>& a) {
    if (a.L < 1 || a.NC < a.L || a.n_own < 0 || a.T_steps < 0 || a.R < 0 || !a.cell_off || !a.dx || !a.nadj || !a.one_adj ||
        !a.adj_off || !a.own_slot)
        return DHTS_ERR_INVALID;
    if (a.mode == 1 && (!a.sig || !a.incoming)) return DHTS_ERR_INVALID;
    if (a.mode != 0 && a.mode != 1) return DHTS_ERR_INVALID;
    return DHTS_OK;
}

template <typename T, typename K> static int net_launch_cfg(K kernel, size_t smem, int threads, int R, int* grid) {
    if (smem > 227 * 1024) return DHTS_ERR_UNSUPPORTED;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) {
        cudaGetLastError();
        return DHTS_ERR_CUDA;
    }
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, threads, smem);
    if (occ < 1) occ = 1;
    long long g = (long long)net_sm_count() * occ;
    *grid = (int)(g < R ? g : R);
    return DHTS_OK;
}

template <typename T>
static int net_fwd(const NetArgs<T>& a, const T* r0, const T* y0, const T* u0, const T* own0, T* hist, T* ownh, T* reward,
                   int* flags, cudaStream_t st) {
    int rc = net_check(a);
    if (rc) return rc;
    if (!r0 || !y0 || !u0 || !hist || !flags || (a.n_own > 0 && (!own0 || !ownh))) return DHTS_ERR_INVALID;
    if (a.R == 0) return DHTS_OK;
    const int threads = net_threads(a.L);
    const size_t smem = net_smem<T>(a.L, a.NC, a.n_own, false, threads);
    int grid = 1;
    rc = net_launch_cfg<T>(net_rollout_fwd_kernel<T>, smem, threads, a.R, &grid);
    if (rc) return rc;
    net_rollout_fwd_kernel<T><<<grid, threads, smem, st>>>(a, r0, y0, u0, own0, hist, ownh, reward, flags);
    return cudaGetLastError() == cudaSuccess ? DHTS_OK : DHTS_ERR_CUDA;
}

template <typename T>
static int net_bwd(const NetArgs<T>& a, const T* hist, const T* ownh, const T* g_states, const T* g_reward, T* g_r0, T* g_y0,
                   T* g_u0, T* g_own0, T* g_sig, T* g_inc, int* flags, cudaStream_t st) {
    int rc = net_check(a);
    if (rc) return rc;
    if (!hist || !g_r0 || !g_y0 || !g_u0 || !flags || (a.n_own > 0 && !ownh)) return DHTS_ERR_INVALID;
    if (a.R == 0) return DHTS_OK;
    const int threads = net_threads(a.L);
    const size_t smem = net_smem<T>(a.L, a.NC, a.n_own, true, threads);
    int grid = 1;
    rc = net_launch_cfg<T>(net_rollout_bwd_kernel<T>, smem, threads, a.R, &grid);
    if (rc) return rc;
    net_rollout_bwd_kernel<T><<<grid, threads, smem, st>>>(a, hist, ownh, g_states, g_reward, g_r0, g_y0, g_u0, g_own0, g_sig,
                                                           g_inc, flags);
    return cudaGetLastError() == cudaSuccess ? DHTS_OK : DHTS_ERR_CUDA;
}

template <typename T> static NetArgs<T> net_args(const dhts_net_topology* tp, const int* route, int route_per_replica, T umax, T dt,
                                                 int steps, int R, int mode, int soft, const T* sig, const T* incoming,
                                                 const T* dx, const T* qk, T veh_len, T static_speed) {
    NetArgs<T> a;
    a.L = tp->L; a.NC = tp->NC; a.n_own = tp->n_own; a.T_steps = steps; a.R = R; a.mode = mode; a.soft = soft;
    a.cell_off = tp->cell_off; a.dx = dx; a.nadj = tp->nadj; a.one_adj = tp->one_adj; a.adj_off = tp->adj_off; a.adj = tp->adj;
    a.own_slot = tp->own_slot; a.route = route; a.route_stride = route_per_replica ? (long long)steps * 2 * tp->L : 0;
    a.umax = umax; a.dt = dt; a.veh_len = veh_len; a.static_speed = static_speed;
    a.sig = sig; a.incoming = incoming; a.qk = qk; a.kind = nullptr;
    return a;
}

}  // namespace dhts

#define DHTS_NET_API(SUF, T)                                                                                           \
    DHTS_EXPORT int dhts_net_rollout_fwd_##SUF(const dhts_net_topology* topo, const T* dx, const int* route,           \
                                               int route_per_replica, const T* sig, const T* incoming, const T* qk,    \
                                               T umax, T dt, T veh_len, T static_speed, int steps, int R, int mode,    \
                                               int soft, const T* r0, const T* y0, const T* u0, const T* own0,         \
                                               T* hist, T* own_hist, T* reward, int* flags, void* stream) {            \
        if (!topo) return DHTS_ERR_INVALID;                                                                            \
        return dhts::net_fwd<T>(dhts::net_args<T>(topo, route, route_per_replica, umax, dt, steps, R, mode, soft, sig, \
                                                  incoming, dx, qk, veh_len, static_speed),                            \
                                r0, y0, u0, own0, hist, own_hist, reward, flags, (cudaStream_t)stream);                \
    }                                                                                                                  \
    DHTS_EXPORT int dhts_net_rollout_bwd_##SUF(const dhts_net_topology* topo, const T* dx, const int* route,           \
                                               int route_per_replica, const T* sig, const T* incoming, const T* qk,    \
                                               T umax, T dt, T veh_len, T static_speed, int steps, int R, int mode,    \
                                               int soft, const T* hist, const T* own_hist, const T* g_states,          \
                                               const T* g_reward, T* g_r0, T* g_y0, T* g_u0, T* g_own0, T* g_sig,      \
                                               T* g_incoming, int* flags, void* stream) {                              \
        if (!topo) return DHTS_ERR_INVALID;                                                                            \
        return dhts::net_bwd<T>(dhts::net_args<T>(topo, route, route_per_replica, umax, dt, steps, R, mode, soft, sig, \
                                                  incoming, dx, qk, veh_len, static_speed),                            \
                                hist, own_hist, g_states, g_reward, g_r0, g_y0, g_u0, g_own0, g_sig, g_incoming, flags, \
                                (cudaStream_t)stream);                                                                 \
    }

DHTS_NET_API(f64, double)
DHTS_NET_API(f32, float)
