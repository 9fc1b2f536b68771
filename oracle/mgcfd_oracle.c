/* oracle/mgcfd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * A plain-C, single-threaded restatement of the MG-CFD per-cycle solver loop, used only as the checker in
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg.  The product (libmgcfd_b200.so) never links,
 * imports or calls anything in this directory.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py checks every function below against the unmodified
 * reference compiled from /root/reference (oracle/_ref/libmgcfd_ref.so, built by oracle/Makefile `ref`) on seeded
 * inputs, and tests/golden/ holds outputs of the reference itself for the GPU box where /root/reference is absent.
 * The reference tree ships no golden vectors of its own.
 *
 * Layouts are the reference's: node arrays AoS double[nel*5] = (rho, mx, my, mz, rhoE) (src/Base/const.h:19-26),
 * edges `edge_neighbour` {long a,b; double x,y,z} (src/Base/definitions.h:83), coords double3, MG map long[].
 * Built with -ffp-contract=off: every operation rounds once, in the order the reference writes it.
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

#define NV 5
#define GAMMA_ 1.4

typedef struct { long a, b; double x, y, z; } orc_edge;

/* src/Base/common.h:24 -- a float literal widened to double: 0.20000000298023224 */
static const double SMOOTH = (double)0.2f;

typedef struct { double rho, m[3], re, v[3], sq, speed, p, c; } pt;

/* compute_velocity, compute_speed_sqd, compute_pressure, compute_speed_of_sound: src/Kernels/cfd_loops.h:121-148 */
static pt point(const double* q) {
    pt s;
    s.rho = q[0]; s.m[0] = q[1]; s.m[1] = q[2]; s.m[2] = q[3]; s.re = q[4];
    for (int d = 0; d < 3; d++) s.v[d] = s.m[d] / s.rho;
    s.sq = s.v[0] * s.v[0] + s.v[1] * s.v[1] + s.v[2] * s.v[2];
    s.speed = sqrt(s.sq);
    s.p = (GAMMA_ - 1.0) * (s.re - 0.5 * s.rho * s.sq);
    s.c = sqrt(GAMMA_ * s.p / s.rho);
    return s;
}

/* compute_flux_contribution: src/Kernels/cfd_loops.h:57-83.  fc[k][d]: k = mx,my,mz,energy ; d = x,y,z */
static void contribution(const pt* s, double fc[4][3]) {
    /* the reference computes the upper triangle as velocity.row * momentum.col and mirrors it */
    fc[0][0] = s->v[0] * s->m[0] + s->p; fc[0][1] = s->v[0] * s->m[1]; fc[0][2] = s->v[0] * s->m[2];
    fc[1][0] = fc[0][1];                 fc[1][1] = s->v[1] * s->m[1] + s->p; fc[1][2] = s->v[1] * s->m[2];
    fc[2][0] = fc[0][2];                 fc[2][1] = fc[1][2];                 fc[2][2] = s->v[2] * s->m[2] + s->p;
    const double de_p = s->re + s->p;
    for (int d = 0; d < 3; d++) fc[3][d] = s->v[d] * de_p;
}

/* initialize_far_field_conditions: src/Kernels/cfd_loops.h:85-119 */
void orc_far_field(double ffv[5], double ffc[12]) {
    const double aoa = (3.1415926535897931 / 180.0) * 0.0;
    ffv[0] = 1.4;
    const double p = 1.0;
    const double c = sqrt(GAMMA_ * p / ffv[0]);
    const double speed = 1.2 * c;
    pt s;
    s.v[0] = speed * cos(aoa); s.v[1] = speed * sin(aoa); s.v[2] = 0.0;
    for (int d = 0; d < 3; d++) ffv[1 + d] = ffv[0] * s.v[d];
    ffv[4] = ffv[0] * (0.5 * (speed * speed)) + (p / (GAMMA_ - 1.0));
    s.rho = ffv[0]; s.re = ffv[4]; s.p = p;
    for (int d = 0; d < 3; d++) s.m[d] = ffv[1 + d];
    double fc[4][3];
    contribution(&s, fc);
    for (int k = 0; k < 4; k++) for (int d = 0; d < 3; d++) ffc[3 * k + d] = fc[k][d];
}

/* adjust_ewt + dampen_ewt: src/Kernels/validation.cpp:28-75, selected per variant as src/euler3d_cpu_double.cpp:337-352 */
void orc_adjust_dampen(int variant, const double* coords, long ne, orc_edge* e) {
    double damp = 0.0;
    if (variant == 2) damp = 5e-8; else if (variant == 3) damp = 1e-7; else if (variant == 4) damp = 2e-7;
    if (damp == 0.0) return;
    for (long i = 0; i < ne; i++) {
        if (e[i].a >= 0 && e[i].b >= 0) {
            /* `dist += d*d` three times; the reference's build contracts the 2nd and 3rd into fma (gcc -ffp-contract=fast on
             * an FMA host, Makefile:95-108) -- written explicitly because this file is compiled with -ffp-contract=off */
            double dist = 0.0;
            for (int d = 0; d < 3; d++) { const double t = coords[3 * e[i].b + d] - coords[3 * e[i].a + d]; dist = d ? fma(t, t, dist) : t * t; }
            dist = sqrt(dist);
            e[i].x /= dist; e[i].y /= dist; e[i].z /= dist;
        }
    }
    for (long i = 0; i < ne; i++) { e[i].x *= damp; e[i].y *= damp; e[i].z *= damp; }
}

/* compute_step_factor_legacy: src/Kernels/cfd_loops.cpp:13-73 */
void orc_step_factor_legacy(long nel, const double* var, const double* vol, double* sf) {
    for (long i = 0; i < nel; i++) {
        const pt s = point(var + NV * i);
        sf[i] = 0.5 / (sqrt(vol[i]) * (sqrt(s.sq) + s.c));
    }
}
/* compute_step_factor: src/Kernels/cfd_loops.cpp:76-157 (local dt, global min, divide by volume) */
void orc_step_factor(long nel, const double* var, const double* vol, double* sf) {
    for (long i = 0; i < nel; i++) {
        const pt s = point(var + NV * i);
        const double dt = cbrt(vol[i]) / (sqrt(s.sq) + s.c);
        sf[i] = 0.5 * dt;
    }
    double mn = sf[0];
    for (long i = 0; i < nel; i++) if (sf[i] < mn) mn = sf[i];
    for (long i = 0; i < nel; i++) sf[i] = mn / vol[i];
}

/* compute_flux_edge: src/Kernels/flux_loops.cpp:78-153, body src/Kernels/flux_kernel.elemfunc.c:18-228 (default build:
 * no FLUX_REUSE_*, no precomputed weights).  End b is evaluated before end a; a's increments are applied first. */
void orc_flux_edge(long first, long n, const orc_edge* e, const double* var, double* flux) {
    for (long i = first; i < first + n; i++) {
        const long a = e[i].a, b = e[i].b;
        const double w[3] = {e[i].x, e[i].y, e[i].z};
        const double ewt = sqrt(w[0] * w[0] + w[1] * w[1] + w[2] * w[2]);
        const pt B = point(var + NV * b);
        double fb[4][3]; contribution(&B, fb);
        const pt A = point(var + NV * a);
        double fa[4][3]; contribution(&A, fa);
        const double factor_a = -ewt * SMOOTH * 0.5 * (A.speed + B.speed + A.c + B.c);
        const double factor_b = -ewt * SMOOTH * 0.5 * (A.speed + B.speed + A.c + B.c);
        const double f[3] = {-0.5 * w[0], -0.5 * w[1], -0.5 * w[2]};
        double av[NV], bv[NV];
        av[0] = factor_a * (A.rho - B.rho) + f[0] * (A.m[0] + B.m[0]) + f[1] * (A.m[1] + B.m[1]) + f[2] * (A.m[2] + B.m[2]);
        av[4] = factor_a * (A.re - B.re) + f[0] * (fa[3][0] + fb[3][0]) + f[1] * (fa[3][1] + fb[3][1]) + f[2] * (fa[3][2] + fb[3][2]);
        for (int k = 0; k < 3; k++)
            av[1 + k] = factor_a * (A.m[k] - B.m[k]) + f[0] * (fa[k][0] + fb[k][0]) + f[1] * (fa[k][1] + fb[k][1]) + f[2] * (fa[k][2] + fb[k][2]);
        bv[0] = factor_b * (B.rho - A.rho) - f[0] * (A.m[0] + B.m[0]) - f[1] * (A.m[1] + B.m[1]) - f[2] * (A.m[2] + B.m[2]);
        bv[4] = factor_b * (B.re - A.re) - f[0] * (fa[3][0] + fb[3][0]) - f[1] * (fa[3][1] + fb[3][1]) - f[2] * (fa[3][2] + fb[3][2]);
        for (int k = 0; k < 3; k++)
            bv[1 + k] = factor_b * (B.m[k] - A.m[k]) - f[0] * (fa[k][0] + fb[k][0]) - f[1] * (fa[k][1] + fb[k][1]) - f[2] * (fa[k][2] + fb[k][2]);
        for (int k = 0; k < NV; k++) flux[NV * a + k] += av[k];
        for (int k = 0; k < NV; k++) flux[NV * b + k] += bv[k];
    }
}

/* compute_boundary_flux_edge: src/Kernels/flux_loops.cpp:10-42, body flux_boundary_kernel.elemfunc.c:15-65 */
void orc_boundary_flux_edge(long first, long n, const orc_edge* e, const double* var, double* flux) {
    for (long i = first; i < first + n; i++) {
        const long b = e[i].b;
        const pt B = point(var + NV * b);
        flux[NV * b + 0] += 0.0;
        flux[NV * b + 1] += e[i].x * B.p;
        flux[NV * b + 2] += e[i].y * B.p;
        flux[NV * b + 3] += e[i].z * B.p;
        flux[NV * b + 4] += 0.0;
    }
}

/* compute_wall_flux_edge: src/Kernels/flux_loops.cpp:44-76, body flux_wall_kernel.elemfunc.c:15-89 */
void orc_wall_flux_edge(long first, long n, const orc_edge* e, const double* var, double* flux, const double ffv[5], const double ffc[12]) {
    for (long i = first; i < first + n; i++) {
        const long b = e[i].b;
        const pt B = point(var + NV * b);
        double fb[4][3]; contribution(&B, fb);
        const double f[3] = {0.5 * e[i].x, 0.5 * e[i].y, 0.5 * e[i].z};
        flux[NV * b + 0] += f[0] * (ffv[1] + B.m[0]) + f[1] * (ffv[2] + B.m[1]) + f[2] * (ffv[3] + B.m[2]);
        for (int k = 0; k < 3; k++)
            flux[NV * b + 1 + k] += f[0] * (ffc[3 * k] + fb[k][0]) + f[1] * (ffc[3 * k + 1] + fb[k][1]) + f[2] * (ffc[3 * k + 2] + fb[k][2]);
        flux[NV * b + 4] += f[0] * (ffc[9] + fb[3][0]) + f[1] * (ffc[10] + fb[3][1]) + f[2] * (ffc[11] + fb[3][2]);
    }
}

/* indirect_rw: src/Kernels/indirect_rw_loop.cpp:11-78, body indirect_rw_kernel.elemfunc.c */
void orc_indirect_rw(long first, long n, const orc_edge* e, const double* var, double* flux) {
    for (long i = first; i < first + n; i++) {
        const double* qa = var + NV * e[i].a; const double* qb = var + NV * e[i].b;
        double* fa = flux + NV * e[i].a; double* fb = flux + NV * e[i].b;
        fa[0] += qb[0] + e[i].x; fa[1] += qb[1] + e[i].z; fa[2] += qb[2]; fa[3] += qb[3]; fa[4] += qb[4] + e[i].y;
        for (int k = 0; k < NV; k++) fb[k] += qa[k];
    }
}

/* time_step: src/Kernels/cfd_loops.cpp:215-280 */
void orc_time_step(int j, long nel, const double* sf, double* flux, const double* old, double* var) {
    for (long i = 0; i < nel; i++) {
        const double factor = sf[i] / (double)(3 + 1 - j);
        for (int k = 0; k < NV; k++) { var[NV * i + k] = old[NV * i + k] + factor * flux[NV * i + k]; flux[NV * i + k] = 0.0; }
    }
}

/* residual + calc_rms: src/Kernels/validation.cpp:77-105 */
void orc_residual(long nel, const double* old, const double* var, double* res) {
    for (long i = 0; i < nel * NV; i++) res[i] = var[i] - old[i];
}
double orc_calc_rms(long nel, const double* res) {
    double rms = 0.0;
    for (long i = 0; i < nel * NV; i++) rms += pow(res[i], 2);
    return sqrt(rms / (double)nel);
}
/* per-variable RMS: not in the reference; sqrt(sum_i r_iv^2 / N), index order */
void orc_rms_per_var(long nel, const double* res, double out[5]) {
    for (int k = 0; k < NV; k++) {
        double acc = 0.0;
        for (long i = 0; i < nel; i++) acc += res[NV * i + k] * res[NV * i + k];
        out[k] = sqrt(acc / (double)nel);
    }
}

/* check_for_invalid_variables: src/Kernels/validation.cpp:107-138; returns -1 if clean else the first bad cell, reason in *why */
long orc_check_invalid(const double* var, long n, int* why) {
    for (long i = 0; i < n; i++) {
        for (int k = 0; k < NV; k++) if (isnan(var[NV * i + k]) || isinf(var[NV * i + k])) { *why = 1; return i; }
        if (var[NV * i] < 0.0) { *why = 2; return i; }
        if (var[NV * i + 4] < 0.0) { *why = 3; return i; }
    }
    *why = 0;
    return -1;
}

/* mg_restrict: src/Kernels/mg_loops.cpp:30-202 (zero mapped coarse nodes, count, accumulate, scale by 1/count) */
void orc_mg_restrict(const double* var1, double* var2, long nel2, const long* map, long* scratch, long mgc) {
    for (long i = 0; i < mgc; i++) for (int k = 0; k < NV; k++) var2[NV * map[i] + k] = 0.0;
    for (long i = 0; i < nel2; i++) scratch[i] = 0;
    for (long i = 0; i < mgc; i++) {
        for (int k = 0; k < NV; k++) var2[NV * map[i] + k] += var1[NV * i + k];
        scratch[map[i]]++;
    }
    for (long i = 0; i < nel2; i++) {
        const double average = scratch[i] == 0 ? 1.0 : 1.0 / (double)scratch[i];
        for (int k = 0; k < NV; k++) var2[NV * i + k] *= average;
    }
}

/* prolong_residuals_interpolate_proper: src/Kernels/mg_loops.cpp:678-864.  Level 1 = coarse, level 2 = fine.
 * Keeps the reference's behaviour at :804-810 (the a1->b2 term multiplies residuals1[b1]). */
static double inv_dist(const double* p, const double* q) {
    const double dx = p[0] - q[0], dy = p[1] - q[1], dz = p[2] - q[2];
    return 1.0 / sqrt(dx * dx + dy * dy + dz * dz);
}
void orc_prolong(const orc_edge* e, long ne, const double* res1, const double* res2, double* var2, long nel2, const long* map,
                 const double* coords1, const double* coords2) {
    double* wsum = (double*)calloc((size_t)nel2, sizeof(double));
    double* wavg = (double*)calloc((size_t)nel2 * NV, sizeof(double));
    for (long i = 0; i < ne; i++) {
        const long n2[2] = {e[i].a, e[i].b};
        const long n1[2] = {map[n2[0]], map[n2[1]]};
        for (int side = 0; side < 2; side++) {
            const long f = n2[side], own = n1[side], other = n1[1 - side];
            const double* cf = coords2 + 3 * f; const double* co = coords1 + 3 * own;
            if (cf[0] - co[0] == 0.0 && cf[1] - co[1] == 0.0 && cf[2] - co[2] == 0.0) {
                for (int k = 0; k < NV; k++) wavg[NV * f + k] = res1[NV * own + k];
                wsum[f] = 1.0;
            } else {
                const double w_own = inv_dist(cf, co);
                for (int k = 0; k < NV; k++) wavg[NV * f + k] += w_own * res1[NV * own + k];
                wsum[f] += w_own;
                const double w_other = inv_dist(coords1 + 3 * other, cf);
                /* side 0 (a2): source is b1 = `other`; side 1 (b2): the reference reads residuals1[b1] = `own` */
                const long src = (side == 0) ? other : own;
                for (int k = 0; k < NV; k++) wavg[NV * f + k] += w_other * res1[NV * src + k];
                wsum[f] += w_other;
            }
        }
    }
    for (long i = 0; i < nel2; i++)
        for (int k = 0; k < NV; k++) {
            wavg[NV * i + k] /= wsum[i];
            var2[NV * i + k] += res2[NV * i + k] - wavg[NV * i + k];
        }
    free(wsum); free(wavg);
}

/* The V-cycle loop of main(): src/euler3d_cpu_double.cpp:371-694 (without the results-neutral indirect_rw probe).
 * All per-level arrays are caller-owned, reference layout.  rms_all[c] and rms_var[c*5+v] are level-0 histories.
 * Returns 0, or 1+cell if check_for_invalid_variables would have aborted. */
typedef struct {
    long nel, nI, nB, nW;
    const double* vol; const orc_edge* edges; const double* coords; const long* map;
    double *var, *old, *res, *flux, *sf;
} orc_level;

static long smooth(orc_level* L, int legacy, const double* ffv, const double* ffc) {
    memcpy(L->old, L->var, sizeof(double) * NV * (size_t)L->nel);
    if (legacy) orc_step_factor_legacy(L->nel, L->var, L->vol, L->sf); else orc_step_factor(L->nel, L->var, L->vol, L->sf);
    for (int j = 0; j < 3; j++) {
        orc_flux_edge(0, L->nI, L->edges, L->var, L->flux);
        orc_boundary_flux_edge(L->nI, L->nB, L->edges, L->var, L->flux);
        orc_wall_flux_edge(L->nI + L->nB, L->nW, L->edges, L->var, L->flux, ffv, ffc);
        orc_time_step(j, L->nel, L->sf, L->flux, L->old, L->var);
        int why; const long bad = orc_check_invalid(L->var, L->nel, &why);
        if (bad >= 0) return 1 + bad;
    }
    orc_residual(L->nel, L->old, L->var, L->res);
    return 0;
}

long orc_run_cycles(int levels, int variant, orc_level* L, int cycles, double* rms_all, double* rms_var) {
    double ffv[5], ffc[12];
    orc_far_field(ffv, ffc);
    const int legacy = (variant == 0);
    long* scratch = (long*)malloc(sizeof(long) * (size_t)L[0].nel);
    int lev = 0, up = 1;
    long rc = 0;
    for (int c = 0; c < cycles && rc == 0;) {
        rc = smooth(&L[lev], legacy, ffv, ffc);
        if (rc) break;
        if (lev == 0) {
            if (rms_all) rms_all[c] = orc_calc_rms(L[0].nel, L[0].res);
            if (rms_var) orc_rms_per_var(L[0].nel, L[0].res, rms_var + 5 * c);
        }
        if (levels <= 1) { c++; continue; }
        if (up) {
            lev++;
            orc_mg_restrict(L[lev - 1].var, L[lev].var, L[lev].nel, L[lev - 1].map, scratch, L[lev - 1].nel);
            if (lev == levels - 1) up = 0;
        } else {
            lev--;
            orc_prolong(L[lev].edges, L[lev].nI, L[lev + 1].res, L[lev].res, L[lev].var, L[lev].nel, L[lev].map, L[lev + 1].coords, L[lev].coords);
            if (lev == 0) { up = 1; c++; }
        }
    }
    free(scratch);
    return rc;
}
