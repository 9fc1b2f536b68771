// assess_kernels.cuh -- the reference's assess-compute protocol on the GPU: the flux kernel's arithmetic in the
// forms the reference's compile-time toggles select (src/Kernels/flux_kernel.elemfunc.c), one thread per internal edge in original
// edge order with RED.ADD.F64 scatters -- identical memory traffic for every variant (the five variables of both end points, the
// edge, ten atomics), so that differences are arithmetic only.  Benchmark kernels: the solver itself never launches them
// (mgcfd_flux_variant / mgcfd_time_kernel selectors 16..31 do).
//   bit 0  FLUX_REUSE_DIV               one reciprocal of the density instead of three (+1) divisions per end point (:46-71)
//   bit 1  FLUX_REUSE_FACTOR + _FLUX    the b-side factor and increments are the negated a-side ones (:132-190)
//   bit 2  FLUX_PRECOMPUTE_EDGE_WEIGHTS |e| read from an array instead of sqrt(e.e) per edge (:24-28)
// Variant 0 is the reference's default build, evaluated the way it is written there (velocity by three divisions, both sides
// computed independently); the production kernels correspond to all three bits plus per-node storage of the derived quantities.
#pragma once
#include "kernels.cuh"

namespace mgcfd {

struct APoint { double rho, mx, my, mz, re, vx, vy, vz, sq, speed, p, c; };

// SOA: the node state as five planes of `stride` doubles (the layout BASELINE.json's north star names) instead of 64-byte records --
// the A/B behind the choice of records: same arithmetic, same edges, same scatters, only the gathers differ
template <bool REUSE_DIV, bool SOA = false>
__host__ __device__ __forceinline__ APoint assess_point(const double* __restrict__ recs, long i, long stride = 0) {
    APoint s;
    if (SOA) {
        s.rho = recs[i]; s.mx = recs[stride + i]; s.my = recs[2 * stride + i]; s.mz = recs[3 * stride + i]; s.re = recs[4 * stride + i];
    } else {
        const double2* q = reinterpret_cast<const double2*>(recs + 8 * i);
        const double2 a = q[0], b = q[1];
        s.rho = a.x; s.mx = a.y; s.my = b.x; s.mz = b.y; s.re = recs[8 * i + 4];
    }
    if (REUSE_DIV) {
        const double r = 1.0 / s.rho;                       // compute_velocity_reciprocal / compute_speed_of_sound_reciprocal
        s.vx = s.mx * r; s.vy = s.my * r; s.vz = s.mz * r;
        s.sq = s.vx * s.vx + s.vy * s.vy + s.vz * s.vz;
        s.speed = sqrt(s.sq);
        s.p = (double(MG_GAMMA) - 1.0) * (s.re - 0.5 * s.rho * s.sq);
        s.c = sqrt(double(MG_GAMMA) * s.p * r);
    } else {
        s.vx = s.mx / s.rho; s.vy = s.my / s.rho; s.vz = s.mz / s.rho;      // compute_velocity, cfd_loops.h:121-128
        s.sq = s.vx * s.vx + s.vy * s.vy + s.vz * s.vz;
        s.speed = sqrt(s.sq);
        s.p = (double(MG_GAMMA) - 1.0) * (s.re - 0.5 * s.rho * s.sq);
        s.c = sqrt(double(MG_GAMMA) * s.p / s.rho);
    }
    return s;
}
// compute_flux_contribution (cfd_loops.h:57-83): fc[k][d], k = momentum x, y, z, density-energy
__host__ __device__ __forceinline__ void assess_contribution(const APoint& s, double fc[4][3]) {
    fc[0][0] = s.vx * s.mx + s.p; fc[0][1] = s.vx * s.my;       fc[0][2] = s.vx * s.mz;
    fc[1][0] = fc[0][1];          fc[1][1] = s.vy * s.my + s.p; fc[1][2] = s.vy * s.mz;
    fc[2][0] = fc[0][2];          fc[2][1] = fc[1][2];          fc[2][2] = s.vz * s.mz + s.p;
    const double de_p = s.re + s.p;
    fc[3][0] = s.vx * de_p; fc[3][1] = s.vy * de_p; fc[3][2] = s.vz * de_p;
}

// the increments of one internal edge for its end a (av) and its end b (bv), flux_kernel.elemfunc.c:18-190; ewt = |e| (computed
// by the caller or read from the precomputed array)
template <bool REUSE_DIV, bool REUSE_FLUX, bool SOA = false>
__host__ __device__ __forceinline__ void assess_edge(const double* __restrict__ recs, long a, long b, double ex, double ey, double ez, double ewt,
                                                     double smoothing, double av[5], double bv[5], long stride = 0) {
    const APoint B = assess_point<REUSE_DIV, SOA>(recs, b, stride);
    double fb[4][3]; assess_contribution(B, fb);
    const APoint A = assess_point<REUSE_DIV, SOA>(recs, a, stride);
    double fa[4][3]; assess_contribution(A, fa);
    const double factor_a = -ewt * smoothing * 0.5 * (A.speed + B.speed + A.c + B.c);
    const double fx = -0.5 * ex, fy = -0.5 * ey, fz = -0.5 * ez;
    const double am[3] = {A.mx, A.my, A.mz}, bm[3] = {B.mx, B.my, B.mz};
    av[0] = factor_a * (A.rho - B.rho) + fx * (A.mx + B.mx) + fy * (A.my + B.my) + fz * (A.mz + B.mz);
    av[4] = factor_a * (A.re - B.re) + fx * (fa[3][0] + fb[3][0]) + fy * (fa[3][1] + fb[3][1]) + fz * (fa[3][2] + fb[3][2]);
    for (int k = 0; k < 3; k++) av[1 + k] = factor_a * (am[k] - bm[k]) + fx * (fa[k][0] + fb[k][0]) + fy * (fa[k][1] + fb[k][1]) + fz * (fa[k][2] + fb[k][2]);
    if (REUSE_FLUX) {
        for (int k = 0; k < 5; k++) bv[k] = -av[k];
    } else {
        const double factor_b = -ewt * smoothing * 0.5 * (A.speed + B.speed + A.c + B.c);
        bv[0] = factor_b * (B.rho - A.rho) - fx * (A.mx + B.mx) - fy * (A.my + B.my) - fz * (A.mz + B.mz);
        bv[4] = factor_b * (B.re - A.re) - fx * (fa[3][0] + fb[3][0]) - fy * (fa[3][1] + fb[3][1]) - fz * (fa[3][2] + fb[3][2]);
        for (int k = 0; k < 3; k++) bv[1 + k] = factor_b * (bm[k] - am[k]) - fx * (fa[k][0] + fb[k][0]) - fy * (fa[k][1] + fb[k][1]) - fz * (fa[k][2] + fb[k][2]);
    }
}

template <bool REUSE_DIV, bool REUSE_FLUX, bool PRE_EW, bool SOA = false>
__global__ void k_flux_assess(long ne, const int* __restrict__ ea, const int* __restrict__ eb, const double* __restrict__ ew,
                              const double* __restrict__ ewt_pre, const double* __restrict__ recs, long stride, double* __restrict__ flux,
                              double smoothing) {
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const int a = ea[e], b = eb[e];
    const double ex = ew[e], ey = ew[ne + e], ez = ew[2 * ne + e];
    const double ewt = PRE_EW ? ewt_pre[e] : sqrt(ex * ex + ey * ey + ez * ez);
    double av[5], bv[5];
    assess_edge<REUSE_DIV, REUSE_FLUX, SOA>(recs, a, b, ex, ey, ez, ewt, smoothing, av, bv, stride);
#pragma unroll
    for (int k = 0; k < 5; k++) atomicAdd(&flux[k * stride + a], av[k]);
#pragma unroll
    for (int k = 0; k < 5; k++) atomicAdd(&flux[k * stride + b], bv[k]);
}

// the five conserved variables of every node as planes (for the SOA variant)
__global__ void k_records_to_planes(long n, const double* __restrict__ recs, long stride, double* __restrict__ planes) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 5; k++) planes[k * stride + i] = recs[8 * i + k];
}

// |e| of every internal edge (FLUX_PRECOMPUTE_EDGE_WEIGHTS: computed once, as the reference does at start-up)
__global__ void k_edge_weights(long ne, const double* __restrict__ ew, double* __restrict__ out) {
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e < ne) out[e] = sqrt(ew[e] * ew[e] + ew[ne + e] * ew[ne + e] + ew[2 * ne + e] * ew[2 * ne + e]);
}

}  // namespace mgcfd
