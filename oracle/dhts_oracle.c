/*
 * dhts_oracle.c -- TEST INFRASTRUCTURE ONLY.  NOT PART OF THE PRODUCT PATH.
 *
 * A plain-C, scalar, CPU restatement of the simulation step of
 * SonSang/diff-hybrid-traffic-sim (reference @ c2ea7b9), used ONLY as the
 * checker in tests/, in __graft_entry__.smoke() and as the `cpu_baseline` /
 * `--impl reference` leg of bench.py.  Nothing under the product package may
 * import, link or call it.
 *
 * Parity pin: every function here is checked in tests/test_oracle_golden.py
 * against outputs of the LIVE reference (fp32-native and dtype-proxied fp64,
 * SURVEY.md section 8c tiers 1-2) frozen under tests/golden/ by
 * oracle/gen_golden.py.
 *
 * Each function cites the reference file:line it restates (paths relative to
 * the reference checkout).  The structure follows the reference (per-lane
 * loops, dqs Jacobian band, band-transpose VJP); the CUDA product path uses a
 * different (flux-difference) formulation, so agreement is a real check.
 *
 * Numerics: all math in double.  When `f32 != 0` the oracle mimics the
 * reference's native mixed precision: state vectors crossing the step
 * boundary are rounded to float (road/lane/_macro_lane.py:271-273,
 * road/lane/_micro_lane.py:247-248 under float32 default) and Jacobian
 * factors / band products are float (model/macro/darz.py:28-31,110-116,
 * road/lane/dmacro_lane.py:56,126-129, road/lane/dmicro_lane.py:54).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* Thread count of the lane-parallel rollouts (bench.py CPU arms); returns what is in effect (1 without OpenMP). */
int orc_set_threads(int n) {
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
    return omp_get_max_threads();
#else
    (void)n;
    return 1;
#endif
}

#define GAMMA 0.5      /* model/macro/_arz.py:1 */
#define EPSILON 1e-5   /* model/macro/_arz.py:2 */
#define IDM_DELTA 4.0  /* model/micro/_idm.py:1 */
#define POSITION_DELTA_EPS 1e-5 /* road/lane/_micro_lane.py:17 */

typedef struct { double r, y, u, ueq; } fullq_t; /* model/macro/_arz.py:50-62 */

static double rnd(double x, int f32) { return f32 ? (double)(float)x : x; }

/* model/macro/_arz.py:133-138 */
double orc_u_eq(double r, double umax) {
    r = r > 0.0 ? r : 0.0;
    return umax * (1.0 - pow(r + EPSILON, GAMMA));
}
/* model/macro/_arz.py:146-149 */
double orc_u_eq_prime(double r, double umax) {
    r = r > EPSILON ? r : EPSILON;
    return -umax * GAMMA * pow(r, GAMMA - 1.0);
}
/* model/macro/_arz.py:126-131 */
double orc_compute_u(double r, double y, double umax) {
    r = r > EPSILON ? r : EPSILON;
    return y / r + orc_u_eq(r, umax);
}
/* model/macro/_arz.py:121-124 */
double orc_compute_y(double r, double u, double umax) {
    return r * (u - orc_u_eq(r, umax));
}
/* model/macro/_arz.py:88-92 */
static fullq_t set_r_y(double r, double y, double umax) {
    fullq_t q; q.r = r; q.y = y;
    q.u = orc_compute_u(r, y, umax);
    q.ueq = orc_u_eq(r, umax);
    return q;
}
/* model/macro/_arz.py:74-80 */
static fullq_t from_r_u(double r, double u, double umax) {
    fullq_t q; q.r = r; q.u = u;
    q.y = orc_compute_y(r, u, umax);
    q.ueq = orc_u_eq(r, umax);
    return q;
}
/* model/macro/_arz.py:103-104 */
static double lambda0(const fullq_t *q, double umax) {
    return q->u + q->r * orc_u_eq_prime(q->r, umax);
}
/* model/macro/_arz.py:186-199 */
static fullq_t compute_Qm(const fullq_t *L, const fullq_t *R, double umax) {
    fullq_t m;
    m.r = pow(pow(L->r, GAMMA) + (L->u - R->u) / umax, 1.0 / GAMMA);
    m.u = R->u;
    m.y = orc_compute_y(m.r, m.u, umax);
    m.ueq = orc_u_eq(m.r, umax);
    return m;
}
/* model/macro/_arz.py:168-183 */
static fullq_t compute_Qc(const fullq_t *L, double umax) {
    fullq_t c;
    c.r = pow((L->u + umax * pow(L->r, GAMMA)) / ((GAMMA + 1.0) * umax), 1.0 / GAMMA);
    c.u = (GAMMA / (GAMMA + 1.0)) * (L->u + umax * pow(L->r, GAMMA));
    c.y = orc_compute_y(c.r, c.u, umax);
    c.ueq = orc_u_eq(c.r, umax);
    return c;
}

typedef struct { fullq_t q0; int case_ind; double speed0, speed1; } riemann_t;

/* model/macro/_arz.py:212-332 */
static riemann_t riemann_solve(const fullq_t *L, const fullq_t *R, double umax) {
    riemann_t s; s.case_ind = -1; s.speed0 = 0; s.speed1 = 0;
    if (L->r < EPSILON) {                       /* :225-232 */
        s.speed0 = 0.0; s.speed1 = L->u; s.case_ind = 0;
    } else if (R->r < EPSILON) {                /* :235-253 */
        fullq_t m = from_r_u(0.0, umax + L->u - L->ueq, umax);
        double l0l = lambda0(L, umax), l0m = m.u;
        s.speed0 = (l0l + l0m) * 0.5; s.speed1 = s.speed0;
        s.case_ind = (l0l >= 0.0) ? 0 : 2;
    } else if (fabs(L->u - R->u) < EPSILON) {   /* :256-262 */
        s.speed0 = 0.0; s.speed1 = R->u; s.case_ind = 0;
    } else if (L->u > R->u) {                   /* :265-280 */
        fullq_t m = compute_Qm(L, R, umax);
        double fd = m.r * m.u - L->r * L->u;
        double den = m.r - L->r; den = den > EPSILON ? den : EPSILON;
        s.speed0 = fd / den; s.speed1 = R->u;
        s.case_ind = (s.speed0 >= 0.0) ? 0 : 1;
    } else if (umax + L->u - L->ueq > R->u) {   /* :283-303 */
        fullq_t m = compute_Qm(L, R, umax);
        double l0l = lambda0(L, umax), l0m = lambda0(&m, umax);
        s.speed0 = (l0l + l0m) * 0.5; s.speed1 = R->u;
        if (l0l >= 0) s.case_ind = 0; else if (l0m <= 0) s.case_ind = 1; else s.case_ind = 2;
    } else {                                    /* :306-322 */
        fullq_t m = from_r_u(0.0, umax + L->u - L->ueq, umax);
        double l0l = lambda0(L, umax), l0m = m.u;
        s.speed0 = (l0l + l0m) * 0.5; s.speed1 = R->u;
        s.case_ind = (l0l >= 0.0) ? 0 : 2;
    }
    if (s.case_ind == 0) s.q0 = set_r_y(L->r, L->y, umax);       /* :155-165 */
    else if (s.case_ind == 1) s.q0 = compute_Qm(L, R, umax);
    else s.q0 = compute_Qc(L, umax);
    return s;
}

/* 2x2 helpers; when f32, factors are float and products are float
 * (np.matmul of float32 operands). */
static void mm2(const double *a, const double *b, double *c, int f32) {
    if (f32) {
        float fa[4], fb[4];
        for (int i = 0; i < 4; i++) { fa[i] = (float)a[i]; fb[i] = (float)b[i]; }
        c[0] = (double)(float)(fa[0] * fb[0] + fa[1] * fb[2]);
        c[1] = (double)(float)(fa[0] * fb[1] + fa[1] * fb[3]);
        c[2] = (double)(float)(fa[2] * fb[0] + fa[3] * fb[2]);
        c[3] = (double)(float)(fa[2] * fb[1] + fa[3] * fb[3]);
    } else {
        c[0] = a[0] * b[0] + a[1] * b[2];
        c[1] = a[0] * b[1] + a[1] * b[3];
        c[2] = a[2] * b[0] + a[3] * b[2];
        c[3] = a[2] * b[1] + a[3] * b[3];
    }
}

/* model/macro/darz.py:35-122 */
static void compute_dM(const fullq_t *M, const fullq_t *L, const fullq_t *R, double umax,
                       double *dL, double *dR) {
    double rL = L->r > EPSILON ? L->r : EPSILON;
    double rR = R->r > EPSILON ? R->r : EPSILON;
    double yL = L->y, yR = R->y;
    double rM = M->r, uM = M->u, ueqM = M->ueq;
    double ueqpM = orc_u_eq_prime(rM, umax);
    double duL_drL = -yL / (rL * rL) + orc_u_eq_prime(rL, umax);
    double duL_dyL = 1.0 / rL;
    double duR_drR = -yR / (rR * rR) + orc_u_eq_prime(rR, umax);
    double duR_dyR = 1.0 / rR;
    double a = (1.0 / GAMMA) * pow(rM, 1.0 - GAMMA);
    double b = GAMMA * pow(rL, GAMMA - 1.0);
    double c = (1.0 / umax) * duL_drL;
    double drM_drL = a * (b + c);
    double d = (1.0 / umax) * duL_dyL;
    double drM_dyL = a * d;
    double e = uM - ueqM;
    double dyM_drL = drM_drL * e + rM * (-ueqpM * drM_drL);
    double dyM_dyL = drM_dyL * e + rM * (-ueqpM * drM_dyL);
    double f = (-1.0 / umax) * duR_drR;
    double drM_drR = a * f;
    double g = (-1.0 / umax) * duR_dyR;
    double drM_dyR = a * g;
    double dyM_drR = drM_drR * e + rM * (duR_drR - ueqpM * drM_drR);
    double dyM_dyR = drM_dyR * e + rM * (duR_dyR - ueqpM * drM_dyR);
    dL[0] = drM_drL; dL[1] = drM_dyL; dL[2] = dyM_drL; dL[3] = dyM_dyL;
    dR[0] = drM_drR; dR[1] = drM_dyR; dR[2] = dyM_drR; dR[3] = dyM_dyR;
}

/* model/macro/darz.py:124-192 */
static void compute_dC(const fullq_t *C, const fullq_t *L, double umax, double *dL, double *dR) {
    double rL = L->r > EPSILON ? L->r : EPSILON;
    double yL = L->y;
    double ueqpL = orc_u_eq_prime(rL, umax);
    double rC = C->r, uC = C->u, ueqC = C->ueq;
    double ueqpC = orc_u_eq_prime(rC, umax);
    double duL_drL = -yL / (rL * rL) + ueqpL;
    double duL_dyL = 1.0 / rL;
    double f = umax * GAMMA * pow(rL, GAMMA - 1.0);
    double duC_drL = (GAMMA / (GAMMA + 1.0)) * (duL_drL + f);
    double duC_dyL = (GAMMA / (GAMMA + 1.0)) * duL_dyL;
    double b = (GAMMA + 1.0) * umax;
    double c = pow(rC, 1.0 - GAMMA);
    double d = c / GAMMA;
    double e = d / b;
    double drC_drL = e * (duL_drL + f);
    double drC_dyL = e * duL_dyL;
    double g = uC - ueqC;
    double dyC_drL = drC_drL * g + rC * (duC_drL - ueqpC * drC_drL);
    double dyC_dyL = drC_dyL * g + rC * (duC_dyL - ueqpC * drC_dyL);
    dL[0] = drC_drL; dL[1] = drC_dyL; dL[2] = dyC_drL; dL[3] = dyC_dyL;
    dR[0] = dR[1] = dR[2] = dR[3] = 0.0;
}

/* model/macro/darz.py:194-215 (dispatch), :12-33 (dL) */
static void compute_dLdR(const riemann_t *rs, const fullq_t *L, const fullq_t *R, double umax,
                         double *dL, double *dR) {
    if (rs->case_ind == 0) {
        dL[0] = 1; dL[1] = 0; dL[2] = 0; dL[3] = 1;
        dR[0] = dR[1] = dR[2] = dR[3] = 0;
    } else if (rs->case_ind == 1) compute_dM(&rs->q0, L, R, umax, dL, dR);
    else compute_dC(&rs->q0, L, umax, dL, dR);
}

/* model/macro/darz.py:217-233 */
static void flux_prime(const fullq_t *q, double umax, double *fp) {
    double r = q->r > EPSILON ? q->r : EPSILON;
    double y = q->y, ueq = q->ueq;
    double ueqp = orc_u_eq_prime(r, umax);
    fp[0] = ueq + r * ueqp;
    fp[1] = 1.0;
    fp[2] = y * ueqp - (y / r) * (y / r);
    fp[3] = (2.0 * y) / r + ueq;
}

/*
 * One lane, one step: road/lane/_macro_lane.py:83-146 (Godunov update + CFL
 * check) and road/lane/dmacro_lane.py:96-132 (Jacobian band).
 *
 * pad_* are the N+2 padded cell records (index 0 / N+1 = ghosts), each a
 * full (r, y, u, ueq) record because the reference STORES u and u_eq on the
 * cell rather than recomputing them at use.
 * Outputs: nr, ny, nu [N] (nu = compute_u via set_r_y, _macro_lane.py:112);
 * case_ind [N+1]; speeds [N+1][2]; dqs [N][3][2][2] (may be NULL);
 * returns 1 when the CFL assert (_macro_lane.py:141-146) would fire.
 */
int orc_arz_step(const double *pad_r, const double *pad_y, const double *pad_u, const double *pad_ueq,
                 int N, double dx, double umax, double dt, int f32,
                 double *nr, double *ny, double *nu, int *case_ind, double *speeds, double *dqs) {
    riemann_t *rs = (riemann_t *)malloc(sizeof(riemann_t) * (size_t)(N + 1));
    fullq_t *cell = (fullq_t *)malloc(sizeof(fullq_t) * (size_t)(N + 2));
    int cfl = 0;
    for (int i = 0; i < N + 2; i++) {
        cell[i].r = pad_r[i]; cell[i].y = pad_y[i]; cell[i].u = pad_u[i]; cell[i].ueq = pad_ueq[i];
    }
    for (int i = 0; i <= N; i++) {
        rs[i] = riemann_solve(&cell[i], &cell[i + 1], umax);
        double s0 = fabs(rs[i].speed0), s1 = fabs(rs[i].speed1);
        s0 = s0 > 1e-5 ? s0 : 1e-5; s1 = s1 > 1e-5 ? s1 : 1e-5;
        if (!(dt < dx / s0 && dt < dx / s1)) cfl = 1;
        if (case_ind) case_ind[i] = rs[i].case_ind;
        if (speeds) { speeds[2 * i] = rs[i].speed0; speeds[2 * i + 1] = rs[i].speed1; }
    }
    double c = dt / dx;
    for (int j = 0; j < N; j++) {
        const fullq_t *a = &rs[j].q0, *b = &rs[j + 1].q0;
        double r = cell[j + 1].r + (a->r * a->u - b->r * b->u) * c;
        double y = cell[j + 1].y + (a->y * a->u - b->y * b->u) * c;
        /* get_next_state_vector -> float32 tensors (_macro_lane.py:327-336) */
        nr[j] = rnd(r, f32); ny[j] = rnd(y, f32);
        if (nu) nu[j] = orc_compute_u(nr[j], ny[j], umax);
    }
    if (dqs) {
        for (int j = 0; j < N; j++) {
            double LdL[4], LdR[4], RdL[4], RdR[4], fpL[4], fpR[4], m[4], m2[4];
            compute_dLdR(&rs[j], &cell[j], &cell[j + 1], umax, LdL, LdR);
            compute_dLdR(&rs[j + 1], &cell[j + 1], &cell[j + 2], umax, RdL, RdR);
            flux_prime(&rs[j].q0, umax, fpL);
            flux_prime(&rs[j + 1].q0, umax, fpR);
            double *o = dqs + (size_t)j * 12;
            mm2(fpL, LdL, m, f32);
            for (int k = 0; k < 4; k++) o[k] = rnd(-c * (-m[k]), f32);
            mm2(fpR, RdR, m, f32);
            for (int k = 0; k < 4; k++) o[8 + k] = rnd(-c * m[k], f32);
            mm2(fpR, RdL, m, f32);
            mm2(fpL, LdR, m2, f32);
            for (int k = 0; k < 4; k++) {
                double eye = (k == 0 || k == 3) ? 1.0 : 0.0;
                o[4 + k] = rnd(eye - rnd(c * rnd(m[k] - m2[k], f32), f32), f32);
            }
        }
    }
    free(rs); free(cell);
    return cfl;
}

/* road/lane/dmacro_lane.py:277-310: band-transpose VJP.
 * g_nr, g_ny [N] -> g_r, g_y [N+2] (ghost entries included). */
void orc_arz_vjp(const double *dqs, int N, const double *g_nr, const double *g_ny, int f32,
                 double *g_r, double *g_y) {
    double *gc = (double *)malloc(sizeof(double) * (size_t)N * 6); /* [N][3][2] */
    for (int j = 0; j < N; j++)
        for (int k = 0; k < 3; k++) {
            const double *d = dqs + (size_t)j * 12 + k * 4;
            /* transpose(d) @ g */
            double a = d[0] * g_nr[j] + d[2] * g_ny[j];
            double b = d[1] * g_nr[j] + d[3] * g_ny[j];
            gc[j * 6 + k * 2] = rnd(a, f32); gc[j * 6 + k * 2 + 1] = rnd(b, f32);
        }
    for (int i = 0; i < N + 2; i++) { g_r[i] = 0; g_y[i] = 0; }
    for (int j = 0; j < N; j++) { g_r[j + 1] = gc[j * 6 + 2]; g_y[j + 1] = gc[j * 6 + 3]; }
    for (int j = 0; j < N - 1; j++) {   /* grad_ry[2:-1] += grad_cell[:-1, 2] */
        g_r[j + 2] = rnd(g_r[j + 2] + gc[j * 6 + 4], f32); g_y[j + 2] = rnd(g_y[j + 2] + gc[j * 6 + 5], f32);
    }
    for (int j = 1; j < N; j++) {       /* grad_ry[1:-2] += grad_cell[1:, 0] */
        g_r[j] = rnd(g_r[j] + gc[j * 6 + 0], f32); g_y[j] = rnd(g_y[j] + gc[j * 6 + 1], f32);
    }
    g_r[0] = gc[0]; g_y[0] = gc[1];
    g_r[N + 1] = gc[(N - 1) * 6 + 4]; g_y[N + 1] = gc[(N - 1) * 6 + 5];
    free(gc);
}

/* d compute_u / d(r, y) as torch autograd sees it outside the Function
 * (set_r_y, model/macro/_arz.py:88-92,126-138): TRUE derivative of
 * u = y/max(r,eps) + umax(1-(max(r,eps)+eps)^gamma). */
void orc_du_dry(double r, double y, double umax, double *du_dr, double *du_dy) {
    if (r >= EPSILON) {   /* python max(r, EPS) returns r unless EPS > r */
        *du_dr = -y / (r * r) - umax * GAMMA * pow(r + EPSILON, GAMMA - 1.0);
        *du_dy = 1.0 / r;
    } else {
        *du_dr = 0.0; *du_dy = 1.0 / EPSILON;
    }
}
/* d compute_y / d(r, u) as autograd sees it (model/macro/_arz.py:121-124). */
void orc_dy_dru(double r, double u, double umax, double *dy_dr, double *dy_du) {
    double ueqp = (r >= 0.0) ? -umax * GAMMA * pow(r + EPSILON, GAMMA - 1.0) : 0.0;
    *dy_dr = (u - orc_u_eq(r, umax)) - r * ueqp;
    *dy_du = r;
}

/*
 * T-step rollout of B independent lanes with static ghosts, forward and
 * adjoint; restates the loop of example/inverse/_inverse.py:91-99 over
 * RoadNetwork.forward (road/network/road_network.py:79-111) for a network
 * of disconnected dMacroLanes, i.e. per step: ghosts re-derived from their
 * own (r,u) (road_network.py:364-387, _macro_lane.py:156-163), lane.forward,
 * update_state.
 *
 * r0,u0 [B][N]; ghost_ru [B][2][2] = (r,u) left, (r,u) right.
 * Forward outputs rT,yT,uT [B][N].  If g_rT != NULL also runs the adjoint:
 * terminal adjoints g_rT,g_yT,g_uT [B][N] -> g_r0,g_u0 [B][N] and
 * g_ghost [B][2][2] (wrt ghost r,u, summed over steps).
 * hist (optional) [T+1][B][N][2] receives every (r,y) state.
 * Returns number of lanes that tripped the CFL assert.
 */
/* ghost_t  optional [T][B][2][2]: (r, u) of the two ghost cells PER STEP (a lane inside a network gets new ghosts every
 *          step, road_network.py:364-387); overrides ghost_ru.
 * g_hist   optional [T][B][N][2]: dLoss/d(r, y) of the state BEFORE step t, i.e. a loss that reads the lane at every
 *          step (what autograd through dmacro_lane.py:277-310 accumulates for such a loss).
 * g_ghost  [B][2][2] summed over the steps, or with ghost_t: [T][B][2][2] per step. */
int orc_arz_rollout_ex(const double *r0, const double *u0, const double *ghost_ru, const double *ghost_t,
                       int B, int N, const double *dx, const double *umax, double dt, int T, int f32,
                       double *rT, double *yT, double *uT,
                       const double *g_rT, const double *g_yT, const double *g_uT, const double *g_hist,
                       double *g_r0, double *g_u0, double *g_ghost, double *hist);

int orc_arz_rollout(const double *r0, const double *u0, const double *ghost_ru,
                    int B, int N, const double *dx, const double *umax, double dt, int T, int f32,
                    double *rT, double *yT, double *uT,
                    const double *g_rT, const double *g_yT, const double *g_uT,
                    double *g_r0, double *g_u0, double *g_ghost, double *hist) {
    return orc_arz_rollout_ex(r0, u0, ghost_ru, NULL, B, N, dx, umax, dt, T, f32, rT, yT, uT, g_rT, g_yT, g_uT, NULL,
                              g_r0, g_u0, g_ghost, hist);
}

int orc_arz_rollout_ex(const double *r0, const double *u0, const double *ghost_ru, const double *ghost_t,
                       int B, int N, const double *dx, const double *umax, double dt, int T, int f32,
                       double *rT, double *yT, double *uT,
                       const double *g_rT, const double *g_yT, const double *g_uT, const double *g_hist,
                       double *g_r0, double *g_u0, double *g_ghost, double *hist) {
    int ncfl = 0;
    int P = N + 2;
#pragma omp parallel for schedule(dynamic) reduction(+ : ncfl)
    for (int b = 0; b < B; b++) {
        double um = umax[b], dxb = dx[b];
        double *sr = (double *)malloc(sizeof(double) * (size_t)(T + 1) * P * 4);
        double *dq = g_rT ? (double *)malloc(sizeof(double) * (size_t)T * N * 12) : NULL;
        /* padded records per step: [t][4][P] */
        #define REC(t, k) (sr + ((size_t)(t) * 4 + (k)) * P)
        double *pr = REC(0, 0), *py = REC(0, 1), *pu = REC(0, 2), *pe = REC(0, 3);
        for (int j = 0; j < N; j++) {   /* set_state_vector_u -> set_r_u (_arz.py:82-86) */
            double r = rnd(r0[(size_t)b * N + j], f32), u = rnd(u0[(size_t)b * N + j], f32);
            fullq_t q = from_r_u(r, u, um);
            /* fp32-native: y and u_eq are 0-dim float32 tensor results */
            pr[j + 1] = q.r; py[j + 1] = rnd(q.y, f32); pu[j + 1] = q.u; pe[j + 1] = rnd(q.ueq, f32);
        }
        int cfl = 0;
        for (int t = 0; t < T; t++) {
            pr = REC(t, 0); py = REC(t, 1); pu = REC(t, 2); pe = REC(t, 3);
            const double *gh = ghost_t ? ghost_t + ((size_t)t * B + b) * 4 : ghost_ru + (size_t)b * 4;
            for (int s = 0; s < 2; s++) {
                fullq_t q = from_r_u(gh[s * 2], gh[s * 2 + 1], um);
                int i = s ? N + 1 : 0;
                pr[i] = q.r; py[i] = rnd(q.y, f32); pu[i] = q.u; pe[i] = rnd(q.ueq, f32);
            }
            double *nr = REC(t + 1, 0) + 1, *ny = REC(t + 1, 1) + 1, *nu = REC(t + 1, 2) + 1, *ne = REC(t + 1, 3) + 1;
            cfl |= orc_arz_step(pr, py, pu, pe, N, dxb, um, dt, f32, nr, ny, nu, NULL, NULL,
                                dq ? dq + (size_t)t * N * 12 : NULL);
            /* set_next_state_vector_y -> set_r_y on float32 tensors (_macro_lane.py:282-299) */
            for (int j = 0; j < N; j++) { nu[j] = rnd(nu[j], f32); ne[j] = rnd(orc_u_eq(nr[j], um), f32); }
            if (hist)
                for (int j = 0; j < N; j++) {
                    hist[(((size_t)t * B + b) * N + j) * 2] = pr[j + 1];
                    hist[(((size_t)t * B + b) * N + j) * 2 + 1] = py[j + 1];
                }
        }
        pr = REC(T, 0); py = REC(T, 1); pu = REC(T, 2);
        for (int j = 0; j < N; j++) {
            rT[(size_t)b * N + j] = pr[j + 1]; yT[(size_t)b * N + j] = py[j + 1]; uT[(size_t)b * N + j] = pu[j + 1];
            if (hist) {
                hist[(((size_t)T * B + b) * N + j) * 2] = pr[j + 1];
                hist[(((size_t)T * B + b) * N + j) * 2 + 1] = py[j + 1];
            }
        }
        ncfl += cfl;
        if (g_rT) {
            double *gr = (double *)malloc(sizeof(double) * P * 4);
            double *gy = gr + P, *hr = gy + P, *hy = hr + P;
            double gg[4] = {0, 0, 0, 0};
            for (int j = 0; j < N; j++) {
                double dr, dy;
                orc_du_dry(pr[j + 1], py[j + 1], um, &dr, &dy);
                double gu = g_uT ? g_uT[(size_t)b * N + j] : 0.0;
                gr[j] = g_rT[(size_t)b * N + j] + gu * dr;
                gy[j] = (g_yT ? g_yT[(size_t)b * N + j] : 0.0) + gu * dy;
            }
            for (int t = T - 1; t >= 0; t--) {
                orc_arz_vjp(dq + (size_t)t * N * 12, N, gr, gy, f32, hr, hy);
                const double *gh = ghost_t ? ghost_t + ((size_t)t * B + b) * 4 : ghost_ru + (size_t)b * 4;
                for (int s = 0; s < 2; s++) {
                    int i = s ? N + 1 : 0;
                    double dyr, dyu;
                    orc_dy_dru(gh[s * 2], gh[s * 2 + 1], um, &dyr, &dyu);
                    if (ghost_t && g_ghost) {
                        g_ghost[((size_t)t * B + b) * 4 + s * 2] = hr[i] + hy[i] * dyr;
                        g_ghost[((size_t)t * B + b) * 4 + s * 2 + 1] = hy[i] * dyu;
                    } else {
                        gg[s * 2] += hr[i] + hy[i] * dyr;
                        gg[s * 2 + 1] += hy[i] * dyu;
                    }
                }
                for (int j = 0; j < N; j++) { gr[j] = hr[j + 1]; gy[j] = hy[j + 1]; }
                if (g_hist)
                    for (int j = 0; j < N; j++) {
                        gr[j] += g_hist[(((size_t)t * B + b) * N + j) * 2];
                        gy[j] += g_hist[(((size_t)t * B + b) * N + j) * 2 + 1];
                    }
            }
            for (int j = 0; j < N; j++) {
                double dyr, dyu;
                double r = rnd(r0[(size_t)b * N + j], f32), u = rnd(u0[(size_t)b * N + j], f32);
                orc_dy_dru(r, u, um, &dyr, &dyu);
                g_r0[(size_t)b * N + j] = gr[j] + gy[j] * dyr;
                g_u0[(size_t)b * N + j] = gy[j] * dyu;
            }
            if (g_ghost && !ghost_t) for (int k = 0; k < 4; k++) g_ghost[(size_t)b * 4 + k] = gg[k];
            free(gr);
        }
        #undef REC
        free(sr); if (dq) free(dq);
    }
    return ncfl;
}

/* ------------------------------------------------------------------ IDM */

/* model/micro/_idm.py:5-51 */
static double idm_acc(double a_max, double a_pref, double v, double v_t, double pos_delta, double vel_delta,
                      double min_space, double time_pref, double dt, double *opt_spacing, int *clip_acc,
                      int *clip_s) {
    double s = min_space + v * time_pref + (v * vel_delta) / (2 * pow(a_max * a_pref, 0.5));
    *clip_s = s < 0.0;
    s = s > 0 ? s : 0;
    double acc = a_max * (1.0 - pow(v / v_t, IDM_DELTA) - pow(s / pos_delta, 2.0));
    *clip_acc = acc < -v / dt;
    if (*clip_acc) acc = -v / dt;
    *opt_spacing = s;
    return acc;
}

/*
 * One lane, one step: road/lane/_micro_lane.py:131-214 (forward, leader
 * lookup, collision handling) + road/lane/dmicro_lane.py:87-127 and
 * model/micro/didm.py:12-102 (Jacobian band).
 *
 * p, v [n]; params [6][n] = a_max, a_pref, v_target, min_space, time_pref,
 * length; head_dp/head_dv: ghost leader deltas (_micro_lane.py:199-202).
 * Outputs np_, nv_ [n]; flags [n] bit0 = clipped_acc, bit1 = clipped_s*,
 * bit2 = collision; dqs [n][2][2][2] (may be NULL).
 * Returns number of collisions (reference prints and continues).
 */
int orc_idm_step(const double *p, const double *v, const double *params, int n, double head_dp, double head_dv,
                 double dt, int f32, double *np_, double *nv_, int *flags, double *dqs) {
    const double *a_max = params, *a_pref = params + n, *v_t = params + 2 * n, *s0 = params + 3 * n,
                 *tp = params + 4 * n, *len = params + 5 * n;
    int ncol = 0;
    for (int i = 0; i < n; i++) {
        double dp, dv;
        if (i == n - 1) { dp = head_dp; dv = head_dv; }
        else { dp = fabs(p[i + 1] - p[i]) - (len[i + 1] + len[i]) * 0.5; dv = v[i] - v[i + 1]; }
        double dp_raw = dp, dv_raw = dv;
        int col = 0;
        if (dp < 0) { col = 1; ncol++; dp = 0; dv = 0; }           /* :151-162 */
        dp = dp > POSITION_DELTA_EPS ? dp : POSITION_DELTA_EPS;   /* :168 */
        double s; int ca, cs;
        double acc = idm_acc(a_max[i], a_pref[i], v[i], v_t[i], dp, dv, s0[i], tp[i], dt, &s, &ca, &cs);
        np_[i] = rnd(p[i] + dt * v[i], f32);
        nv_[i] = rnd(v[i] + dt * acc, f32);
        if (flags) flags[i] = ca | (cs << 1) | (col << 2);
        if (dqs) {
            /* dmicro_lane.py:97 re-derives the RAW deltas (no floor, no collision reset) */
            double *E = dqs + (size_t)i * 8, *L = E + 4;
            double sab = sqrt(a_max[i] * a_pref[i]);
            E[0] = 1; E[1] = dt; E[2] = 0; E[3] = 0;
            L[0] = 0; L[1] = 0; L[2] = 0; L[3] = 0;
            if (!ca) {
                E[2] = dt * (-2 * a_max[i] * (pow(s, 2) / pow(dp_raw, 3)));
                L[2] = dt * (2 * a_max[i] * (pow(s, 2) / pow(dp_raw, 3)));
                double t1 = -IDM_DELTA * (pow(v[i], IDM_DELTA - 1) / pow(v_t[i], IDM_DELTA));
                if (cs) {
                    E[3] = 1 + dt * a_max[i] * t1;
                    L[3] = dt * a_max[i] * (-2 * (s / pow(dp_raw, 2)));
                } else {
                    E[3] = 1 + dt * a_max[i] * (t1 - 2 * (s / pow(dp_raw, 2)) * (tp[i] + ((v[i] + dv_raw) / (2 * sab))));
                    L[3] = dt * a_max[i] * (-2 * (s / pow(dp_raw, 2)) * (-v[i] / (2 * sab)));
                }
            }
            if (f32) for (int k = 0; k < 8; k++) E[k] = (double)(float)E[k];
        }
    }
    return ncol;
}

/* road/lane/dmicro_lane.py:271-297: g_np,g_ns [n] -> g_p,g_s [n+1]
 * (entry n = ghost leader). */
void orc_idm_vjp(const double *dqs, int n, const double *g_np, const double *g_ns, int f32, double *g_p,
                 double *g_s) {
    for (int i = 0; i <= n; i++) { g_p[i] = 0; g_s[i] = 0; }
    for (int i = 0; i < n; i++) {
        const double *E = dqs + (size_t)i * 8;
        g_p[i] = rnd(E[0] * g_np[i] + E[2] * g_ns[i], f32);
        g_s[i] = rnd(E[1] * g_np[i] + E[3] * g_ns[i], f32);
    }
    for (int i = 0; i < n; i++) {
        const double *L = dqs + (size_t)i * 8 + 4;
        g_p[i + 1] = rnd(g_p[i + 1] + rnd(L[0] * g_np[i] + L[2] * g_ns[i], f32), f32);
        g_s[i + 1] = rnd(g_s[i + 1] + rnd(L[1] * g_np[i] + L[3] * g_ns[i], f32), f32);
    }
}

/*
 * T-step rollout of independent micro lanes (CSR offsets) with the default
 * ghost leader, forward and adjoint.  Restates example/inverse/_inverse.py:91-99
 * over RoadNetwork.forward for single-lane routes: setup_micro_boundary gives
 * (head_dp, head_dv) constants (road_network.py:548-553), dMicroLane.forward
 * appends the ghost leader p_head+dp, v_head-dv (dmicro_lane.py:130-153) so
 * the ghost adjoint flows back to the head vehicle (and to head_dp / head_dv).
 *
 * p0, v0 [V]; params [6][V]; lane_off [L+1]; head [L][2] = (dp, dv).
 * Outputs pT, vT [V]; adjoint (if g_pT): g_p0, g_v0 [V], g_head [L][2].
 * hist (optional) [T+1][V][2].  Returns total collisions.
 */
/* head_t  optional [T][L][2]: head deltas PER STEP (inside a network RoadNetwork.setup_micro_boundary rewrites them before
 *         every step, road_network.py:429-580); overrides head;  g_head is then [T][L][2], per step.
 * g_hist  optional [T][V][2]: dLoss/d(p, v) of the state BEFORE step t (a loss that reads the lane at every step). */
int orc_idm_rollout_ex(const double *p0, const double *v0, const double *params, int V, const int *lane_off, int L,
                       const double *head, const double *head_t, double dt, int T, int f32, double *pT, double *vT,
                       const double *g_pT, const double *g_vT, const double *g_hist, double *g_p0, double *g_v0,
                       double *g_head, double *hist);

int orc_idm_rollout(const double *p0, const double *v0, const double *params, int V, const int *lane_off, int L,
                    const double *head, double dt, int T, int f32, double *pT, double *vT, const double *g_pT,
                    const double *g_vT, double *g_p0, double *g_v0, double *g_head, double *hist) {
    return orc_idm_rollout_ex(p0, v0, params, V, lane_off, L, head, NULL, dt, T, f32, pT, vT, g_pT, g_vT, NULL, g_p0, g_v0,
                              g_head, hist);
}

int orc_idm_rollout_ex(const double *p0, const double *v0, const double *params, int V, const int *lane_off, int L,
                       const double *head, const double *head_t, double dt, int T, int f32, double *pT, double *vT,
                       const double *g_pT, const double *g_vT, const double *g_hist, double *g_p0, double *g_v0,
                       double *g_head, double *hist) {
    int ncol = 0;
#pragma omp parallel for schedule(dynamic) reduction(+ : ncol)
    for (int l = 0; l < L; l++) {
        int o = lane_off[l], n = lane_off[l + 1] - o;
        if (n <= 0) {
            if (g_head && !head_t) { g_head[2 * l] = 0; g_head[2 * l + 1] = 0; }
            for (int t = 0; g_head && head_t && t < T; t++) { g_head[((size_t)t * L + l) * 2] = 0; g_head[((size_t)t * L + l) * 2 + 1] = 0; }
            continue;
        }
        double *st = (double *)malloc(sizeof(double) * (size_t)(T + 1) * n * 2);
        double *dq = g_pT ? (double *)malloc(sizeof(double) * (size_t)T * n * 8) : NULL;
        double *par = (double *)malloc(sizeof(double) * 6 * n);
        for (int k = 0; k < 6; k++) for (int i = 0; i < n; i++) par[k * n + i] = params[(size_t)k * V + o + i];
        for (int i = 0; i < n; i++) { st[i] = rnd(p0[o + i], f32); st[n + i] = rnd(v0[o + i], f32); }
        for (int t = 0; t < T; t++) {
            double *cp = st + (size_t)t * 2 * n, *cv = cp + n, *np_ = cp + 2 * n, *nv_ = np_ + n;
            const double *hd = head_t ? head_t + ((size_t)t * L + l) * 2 : head + 2 * l;
            ncol += orc_idm_step(cp, cv, par, n, hd[0], hd[1], dt, f32, np_, nv_, NULL,
                                 dq ? dq + (size_t)t * n * 8 : NULL);
        }
        for (int t = 0; hist && t <= T; t++)
            for (int i = 0; i < n; i++) {
                hist[((size_t)t * V + o + i) * 2] = st[(size_t)t * 2 * n + i];
                hist[((size_t)t * V + o + i) * 2 + 1] = st[(size_t)t * 2 * n + n + i];
            }
        for (int i = 0; i < n; i++) { pT[o + i] = st[(size_t)T * 2 * n + i]; vT[o + i] = st[(size_t)T * 2 * n + n + i]; }
        if (g_pT) {
            double *gp = (double *)malloc(sizeof(double) * (size_t)(n + 1) * 4);
            double *gv = gp + n + 1, *hp = gv + n + 1, *hv = hp + n + 1;
            double ghd = 0, ghv = 0;
            for (int i = 0; i < n; i++) { gp[i] = g_pT[o + i]; gv[i] = g_vT[o + i]; }
            for (int t = T - 1; t >= 0; t--) {
                orc_idm_vjp(dq + (size_t)t * n * 8, n, gp, gv, f32, hp, hv);
                /* ghost = p_head + dp ; v_head - dv (dmicro_lane.py:144-151) */
                hp[n - 1] += hp[n]; ghd += hp[n];
                hv[n - 1] += hv[n]; ghv -= hv[n];
                if (head_t && g_head) {
                    g_head[((size_t)t * L + l) * 2] = ghd; g_head[((size_t)t * L + l) * 2 + 1] = ghv;
                    ghd = 0; ghv = 0;
                }
                for (int i = 0; i < n; i++) { gp[i] = hp[i]; gv[i] = hv[i]; }
                if (g_hist)
                    for (int i = 0; i < n; i++) {
                        gp[i] += g_hist[((size_t)t * V + o + i) * 2];
                        gv[i] += g_hist[((size_t)t * V + o + i) * 2 + 1];
                    }
            }
            for (int i = 0; i < n; i++) { g_p0[o + i] = gp[i]; g_v0[o + i] = gv[i]; }
            if (g_head && !head_t) { g_head[2 * l] = ghd; g_head[2 * l + 1] = ghv; }
            free(gp);
        }
        free(st); free(par); if (dq) free(dq);
    }
    return ncol;
}
