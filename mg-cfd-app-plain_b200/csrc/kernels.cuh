// kernels.cuh -- hand-written fp64 CUDA kernels (sm_100a) for the MG-CFD per-cycle solver loop.
// Node state is SoA: plane v of a level lives at base + v*stride (stride = padded node count).
// No tensor cores: the path is gather/scatter and FP64/memory bound.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mgcfd {

#define MG_GAMMA 1.4
// smoothing_coefficient = double(0.2f) (src/Base/common.h:24); -ewt*smoothing*0.5 == ewt*MG_KDISS exactly
// (scaling by 0.5 and negation are exact)
__device__ __constant__ double c_ff[5];      // ff_variable
__device__ __constant__ double c_ffc[12];    // ff_flux_contribution_{momentum_x,y,z,density_energy}

struct NodeVals { double rho, mx, my, mz, re, vx, vy, vz, p, s; };

// compute_velocity / speed_sqd / pressure / speed_of_sound (src/Kernels/cfd_loops.h:121-148), true divisions kept
__device__ __forceinline__ void derive(double rho, double mx, double my, double mz, double re,
                                       double& vx, double& vy, double& vz, double& p, double& speed, double& sos) {
    vx = mx / rho; vy = my / rho; vz = mz / rho;
    const double sq = vx * vx + vy * vy + vz * vz;
    speed = sqrt(sq);
    p = (double(MG_GAMMA) - double(1.0)) * (re - double(0.5) * rho * sq);
    sos = sqrt(double(MG_GAMMA) * p / rho);
}

// One internal edge seen from end A (the thread's node) towards B, (wx,wy,wz) = stored edge vector oriented A->B.
// Returns A's five flux increments (flux_kernel.elemfunc.c:130-162); B's are their exact negation (:170-189 is the
// same expression with every term negated, which IEEE arithmetic rounds symmetrically).
struct Flux5 { double r, mx, my, mz, e; };
__device__ __forceinline__ Flux5 edge_flux(const NodeVals& A, const NodeVals& B, double wx, double wy, double wz, double kdiss) {
    const double ewt = sqrt(wx * wx + wy * wy + wz * wz);
    const double factor = (ewt * kdiss) * (A.s + B.s);
    const double fx = -0.5 * wx, fy = -0.5 * wy, fz = -0.5 * wz;
    // flux contributions of both ends (cfd_loops.h:57-83); A's could be hoisted by the caller, the compiler does it
    const double axx = A.vx * A.mx + A.p, axy = A.vx * A.my, axz = A.vx * A.mz;
    const double ayy = A.vy * A.my + A.p, ayz = A.vy * A.mz, azz = A.vz * A.mz + A.p;
    const double adp = A.re + A.p;
    const double bxx = B.vx * B.mx + B.p, bxy = B.vx * B.my, bxz = B.vx * B.mz;
    const double byy = B.vy * B.my + B.p, byz = B.vy * B.mz, bzz = B.vz * B.mz + B.p;
    const double bdp = B.re + B.p;
    Flux5 f;
    f.r  = factor * (A.rho - B.rho) + fx * (A.mx + B.mx) + fy * (A.my + B.my) + fz * (A.mz + B.mz);
    f.e  = factor * (A.re - B.re) + fx * (A.vx * adp + B.vx * bdp) + fy * (A.vy * adp + B.vy * bdp) + fz * (A.vz * adp + B.vz * bdp);
    f.mx = factor * (A.mx - B.mx) + fx * (axx + bxx) + fy * (axy + bxy) + fz * (axz + bxz);
    f.my = factor * (A.my - B.my) + fx * (axy + bxy) + fy * (ayy + byy) + fz * (ayz + byz);
    f.mz = factor * (A.mz - B.mz) + fx * (axz + bxz) + fy * (ayz + byz) + fz * (azz + bzz);
    return f;
}

// boundary edge (neighbour -1): flux_boundary_kernel.elemfunc.c:33-45
__device__ __forceinline__ Flux5 boundary_flux(const NodeVals& B, double x, double y, double z) {
    Flux5 f; f.r = 0.0; f.e = 0.0;
    f.mx = x * B.p; f.my = y * B.p; f.mz = z * B.p;
    return f;
}
// wall edge (neighbour -2, far field): flux_wall_kernel.elemfunc.c:47-69
__device__ __forceinline__ Flux5 wall_flux(const NodeVals& B, double x, double y, double z) {
    const double fx = 0.5 * x, fy = 0.5 * y, fz = 0.5 * z;
    const double bxx = B.vx * B.mx + B.p, bxy = B.vx * B.my, bxz = B.vx * B.mz;
    const double byy = B.vy * B.my + B.p, byz = B.vy * B.mz, bzz = B.vz * B.mz + B.p;
    const double bdp = B.re + B.p;
    Flux5 f;
    f.r  = fx * (c_ff[1] + B.mx) + fy * (c_ff[2] + B.my) + fz * (c_ff[3] + B.mz);
    f.e  = fx * (c_ffc[9] + B.vx * bdp) + fy * (c_ffc[10] + B.vy * bdp) + fz * (c_ffc[11] + B.vz * bdp);
    f.mx = fx * (c_ffc[0] + bxx) + fy * (c_ffc[1] + bxy) + fz * (c_ffc[2] + bxz);
    f.my = fx * (c_ffc[3] + bxy) + fy * (c_ffc[4] + byy) + fz * (c_ffc[5] + byz);
    f.mz = fx * (c_ffc[6] + bxz) + fy * (c_ffc[7] + byz) + fz * (c_ffc[8] + bzz);
    return f;
}

__device__ __forceinline__ NodeVals load_node(const double* __restrict__ v, long stride, long i) {
    NodeVals n;
    n.rho = v[i]; n.mx = v[stride + i]; n.my = v[2 * stride + i]; n.mz = v[3 * stride + i]; n.re = v[4 * stride + i];
    double speed, sos;
    derive(n.rho, n.mx, n.my, n.mz, n.re, n.vx, n.vy, n.vz, n.p, speed, sos);
    n.s = speed + sos;
    return n;
}

// ------------------------------------------------------------------------------------------------------
// Tiled, coloured flux kernel.  One CTA = one tile of TN owned nodes (thread t <-> node tile*TN+t).
//   phase 1: owned + halo node state is staged in shared memory together with the derived quantities
//            (velocity, pressure, |v|+c), computed ONCE per node per tile instead of twice per edge;
//   phase 2: colour rounds.  In round r thread t evaluates the edge in its slot r, keeps its own end's
//            increment in registers and scatters the other end's (the exact negation) into a shared
//            accumulator.  The host colouring guarantees that within one round no two threads of the CTA
//            target the same node, so plain shared-memory read-modify-writes suffice: no atomics, fixed
//            summation order, bit-reproducible.  Edges cut by a tile boundary are evaluated by both tiles
//            (owner-computes), so nothing is ever scattered outside the tile;
//   phase 3: FUSED: the Runge-Kutta update of time_step (cfd_loops.cpp:215-280) is applied directly,
//            the flux never touches HBM; on the last stage residual, RMS partials and the validity check
//            (validation.cpp:77-138) ride along.  !FUSED: fluxes[node] += total (granular API).
// ------------------------------------------------------------------------------------------------------
struct TileArgs {
    const double* vin;     // stage input state
    const double* vold;    // old_variables (FUSED)
    double* vout;          // stage output state (FUSED) or fluxes (+=) (!FUSED)
    double* res;           // residuals (last stage) or nullptr
    const double* sf;      // step factors
    long stride;
    const long* halo_off; const int* halo_ids;
    const long* slot_off; const int* tile_rounds; const uint16_t* slot_other; const double* slot_w; long nslots;
    const long* bslot_off; const int* tile_brounds; const uint8_t* bslot_kind; const double* bslot_w; long nbslots;
    double rk_div;         // double(RK+1-j)
    double kdiss;
    double* rms_partial;   // [ntiles][5] or nullptr
    unsigned long long* bad_key;  // invalid-state key or nullptr
    const int* old_of_new;
    unsigned long long stage_seq;
    int mask;              // bit0 internal, bit1 boundary, bit2 wall
    int smem_nodes;        // plane length of the staged state (TN + max halo, padded)
};

template <int TN, bool FUSED>
__global__ void __launch_bounds__(TN, (TN <= 256 ? 2 : 1))
k_tile_flux(const TileArgs a) {
    extern __shared__ double sm[];
    const int NL = a.smem_nodes;
    double* st = sm;                 // [10][NL]
    double* acc = sm + 10 * NL;      // [5][TN]
    const int t = threadIdx.x;
    const long tile = blockIdx.x;
    const long gid = tile * TN + t;
    const long h0 = a.halo_off[tile];
    const int nh = int(a.halo_off[tile + 1] - h0);

    NodeVals me = load_node(a.vin, a.stride, gid);
    st[0 * NL + t] = me.rho; st[1 * NL + t] = me.mx; st[2 * NL + t] = me.my; st[3 * NL + t] = me.mz; st[4 * NL + t] = me.re;
    st[5 * NL + t] = me.vx;  st[6 * NL + t] = me.vy; st[7 * NL + t] = me.vz; st[8 * NL + t] = me.p;  st[9 * NL + t] = me.s;
    for (int h = t; h < nh; h += TN) {
        const NodeVals o = load_node(a.vin, a.stride, a.halo_ids[h0 + h]);
        const int l = TN + h;
        st[0 * NL + l] = o.rho; st[1 * NL + l] = o.mx; st[2 * NL + l] = o.my; st[3 * NL + l] = o.mz; st[4 * NL + l] = o.re;
        st[5 * NL + l] = o.vx;  st[6 * NL + l] = o.vy; st[7 * NL + l] = o.vz; st[8 * NL + l] = o.p;  st[9 * NL + l] = o.s;
    }
#pragma unroll
    for (int k = 0; k < 5; k++) acc[k * TN + t] = 0.0;
    __syncthreads();

    double fr = 0.0, fmx = 0.0, fmy = 0.0, fmz = 0.0, fe = 0.0;
    if (a.mask & 1) {
        const int rounds = a.tile_rounds[tile];
        const long s0 = a.slot_off[tile] + t;
        for (int r = 0; r < rounds; r++) {
            const long si = s0 + long(r) * TN;
            const unsigned o = a.slot_other[si];
            if (o != 0xFFFFu) {
                const double wx = a.slot_w[si], wy = a.slot_w[a.nslots + si], wz = a.slot_w[2 * a.nslots + si];
                NodeVals B;
                B.rho = st[0 * NL + o]; B.mx = st[1 * NL + o]; B.my = st[2 * NL + o]; B.mz = st[3 * NL + o]; B.re = st[4 * NL + o];
                B.vx = st[5 * NL + o];  B.vy = st[6 * NL + o]; B.vz = st[7 * NL + o]; B.p = st[8 * NL + o];  B.s = st[9 * NL + o];
                const Flux5 f = edge_flux(me, B, wx, wy, wz, a.kdiss);
                fr += f.r; fmx += f.mx; fmy += f.my; fmz += f.mz; fe += f.e;
                if (o < (unsigned)TN) {   // owned by this tile: conflict-free by colouring
                    acc[0 * TN + o] -= f.r; acc[1 * TN + o] -= f.mx; acc[2 * TN + o] -= f.my; acc[3 * TN + o] -= f.mz; acc[4 * TN + o] -= f.e;
                }
            }
            __syncthreads();
        }
    }
    if (a.mask & 6) {
        const int br = a.tile_brounds[tile];
        const long b0 = a.bslot_off[tile] + t;
        for (int r = 0; r < br; r++) {
            const long si = b0 + long(r) * TN;
            const int kind = a.bslot_kind[si];
            if (kind == 0 || !((a.mask >> kind) & 1)) continue;
            const double x = a.bslot_w[si], y = a.bslot_w[a.nbslots + si], z = a.bslot_w[2 * a.nbslots + si];
            const Flux5 f = (kind == 1) ? boundary_flux(me, x, y, z) : wall_flux(me, x, y, z);
            fr += f.r; fmx += f.mx; fmy += f.my; fmz += f.mz; fe += f.e;
        }
    }
    fr += acc[0 * TN + t]; fmx += acc[1 * TN + t]; fmy += acc[2 * TN + t]; fmz += acc[3 * TN + t]; fe += acc[4 * TN + t];

    const long S = a.stride;
    if (!FUSED) {
        a.vout[gid] += fr; a.vout[S + gid] += fmx; a.vout[2 * S + gid] += fmy; a.vout[3 * S + gid] += fmz; a.vout[4 * S + gid] += fe;
        return;
    } else {
        const double factor = a.sf[gid] / a.rk_div;     // a true divide, cfd_loops.cpp:243
        const double o0 = a.vold[gid], o1 = a.vold[S + gid], o2 = a.vold[2 * S + gid], o3 = a.vold[3 * S + gid], o4 = a.vold[4 * S + gid];
        const double n0 = o0 + factor * fr, n1 = o1 + factor * fmx, n2 = o2 + factor * fmy, n3 = o3 + factor * fmz, n4 = o4 + factor * fe;
        a.vout[gid] = n0; a.vout[S + gid] = n1; a.vout[2 * S + gid] = n2; a.vout[3 * S + gid] = n3; a.vout[4 * S + gid] = n4;
        if (a.bad_key) {
            // check_for_invalid_variables (validation.cpp:107-138): first offending cell of the first offending stage
            int reason = 0;
            if (!(isfinite(n0) && isfinite(n1) && isfinite(n2) && isfinite(n3) && isfinite(n4))) reason = 1;
            else if (n0 < 0.0) reason = 2;
            else if (n4 < 0.0) reason = 3;
            if (reason) {
                const int oi = a.old_of_new[gid];
                if (oi >= 0) atomicMin(a.bad_key, (a.stage_seq << 40) | ((unsigned long long)oi << 2) | (unsigned long long)reason);
            }
        }
        if (a.res) {
            const double r0 = n0 - o0, r1 = n1 - o1, r2 = n2 - o2, r3 = n3 - o3, r4 = n4 - o4;   // residual(), validation.cpp:77-89
            a.res[gid] = r0; a.res[S + gid] = r1; a.res[2 * S + gid] = r2; a.res[3 * S + gid] = r3; a.res[4 * S + gid] = r4;
            if (a.rms_partial) {
                // deterministic block reduction of r^2 per variable: warp shuffles, then warp 0 over the warp sums
                __syncthreads();   // acc is free again
                double q[5] = {r0 * r0, r1 * r1, r2 * r2, r3 * r3, r4 * r4};
#pragma unroll
                for (int k = 0; k < 5; k++) {
#pragma unroll
                    for (int d = 16; d > 0; d >>= 1) q[k] += __shfl_down_sync(0xffffffffu, q[k], d);
                }
                if ((t & 31) == 0) {
#pragma unroll
                    for (int k = 0; k < 5; k++) acc[k * 32 + (t >> 5)] = q[k];
                }
                __syncthreads();
                if (t < 5) {
                    double s = 0.0;
                    for (int w = 0; w < TN / 32; w++) s += acc[t * 32 + w];
                    a.rms_partial[tile * 5 + t] = s;
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// node kernels
// ------------------------------------------------------------------------------------------------------
// compute_step_factor (cfd_loops.cpp:76-157) part 1 / compute_step_factor_legacy (:13-73)
template <bool LEGACY>
__global__ void k_step_factor(const double* __restrict__ v, long stride, long n, const double* __restrict__ vol_root,
                              double* __restrict__ sf, unsigned long long* __restrict__ min_bits) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    double val = __longlong_as_double(0x7F7F7F7F7F7F7F7FLL);
    if (i < n) {
        double vx, vy, vz, p, speed, sos;
        derive(v[i], v[stride + i], v[2 * stride + i], v[3 * stride + i], v[4 * stride + i], vx, vy, vz, p, speed, sos);
        if (LEGACY) {
            sf[i] = double(0.5) / (sqrt(vol_root[i]) * (speed + sos));   // vol_root = volumes here; IEEE sqrt, cfd_loops.cpp:60
        } else {
            const double dt = vol_root[i] / (speed + sos);           // vol_root = cbrt(volume) (host glibc), :123
            val = 0.5 * dt;
        }
    }
    if (!LEGACY) {
        // min over the grid: positive doubles order like their bit patterns
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) val = fmin(val, __shfl_xor_sync(0xffffffffu, val, d));
        __shared__ double wmin[32];
        if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = val;
        __syncthreads();
        if (threadIdx.x < 32) {
            val = (threadIdx.x < (blockDim.x >> 5)) ? wmin[threadIdx.x] : __longlong_as_double(0x7F7F7F7F7F7F7F7FLL);
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) val = fmin(val, __shfl_xor_sync(0xffffffffu, val, d));
            if (threadIdx.x == 0) atomicMin(min_bits, (unsigned long long)__double_as_longlong(val));
        }
    }
}
// step_factors[i] = min_dt / volumes[i] (cfd_loops.cpp:146-156)
__global__ void k_apply_min_dt(const unsigned long long* __restrict__ min_bits, const double* __restrict__ vol, double* __restrict__ sf, long n) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) sf[i] = __longlong_as_double((long long)*min_bits) / vol[i];
}

// time_step (cfd_loops.cpp:215-280), granular API
__global__ void k_time_step(double rk_div, long n, long stride, const double* __restrict__ sf, double* __restrict__ flux,
                            const double* __restrict__ vold, double* __restrict__ v) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double factor = sf[i] / rk_div;
#pragma unroll
    for (int k = 0; k < 5; k++) {
        v[k * stride + i] = vold[k * stride + i] + factor * flux[k * stride + i];
        flux[k * stride + i] = 0.0;
    }
}
__global__ void k_fill(double* __restrict__ p, long n, double val) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = val;
}
__global__ void k_fill_state(double* __restrict__ p, long stride, long n) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 5; k++) p[k * stride + i] = c_ff[k];
}
__global__ void k_copy(double* __restrict__ dst, const double* __restrict__ src, long n) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
// residual (validation.cpp:77-89)
__global__ void k_residual(long n, const double* __restrict__ vold, const double* __restrict__ v, double* __restrict__ r) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) r[i] = v[i] - vold[i];
}
// calc_rms (validation.cpp:91-105) stage 1: per-block sums of r^2 per variable (fixed order => deterministic)
__global__ void k_rms_partial(const double* __restrict__ r, long stride, long n, double* __restrict__ partial) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    __shared__ double ws[5][32];
    double q[5];
#pragma unroll
    for (int k = 0; k < 5; k++) { const double x = (i < n) ? r[k * stride + i] : 0.0; q[k] = x * x; }
#pragma unroll
    for (int k = 0; k < 5; k++)
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) q[k] += __shfl_down_sync(0xffffffffu, q[k], d);
    if ((threadIdx.x & 31) == 0)
#pragma unroll
        for (int k = 0; k < 5; k++) ws[k][threadIdx.x >> 5] = q[k];
    __syncthreads();
    if (threadIdx.x < 5) {
        double s = 0.0;
        for (int w = 0; w < (int)(blockDim.x >> 5); w++) s += ws[threadIdx.x][w];
        partial[blockIdx.x * 5 + threadIdx.x] = s;
    }
}
// stage 2: one block sums the partials in index order; out[slot*6] = {rms_all, rms_var[5]}; bumps *counter if given
__global__ void k_rms_final(const double* __restrict__ partial, long nparts, double nel, double* __restrict__ out, int* counter, int cap) {
    __shared__ double ws[5][256];
    const int t = threadIdx.x;
    double s[5] = {0, 0, 0, 0, 0};
    for (long p = t; p < nparts; p += 256)
#pragma unroll
        for (int k = 0; k < 5; k++) s[k] += partial[p * 5 + k];
#pragma unroll
    for (int k = 0; k < 5; k++) ws[k][t] = s[k];
    __syncthreads();
    if (t == 0) {
        int slot = 0;
        if (counter) { slot = *counter; *counter = slot + 1; if (slot >= cap) slot = cap - 1; }
        double tot = 0.0;
        for (int k = 0; k < 5; k++) {
            double acc = 0.0;
            for (int j = 0; j < 256; j++) acc += ws[k][j];
            out[slot * 6 + 1 + k] = sqrt(acc / nel);
            tot += acc;
        }
        out[slot * 6] = sqrt(tot / nel);
    }
}
// check_for_invalid_variables (validation.cpp:107-138): lowest offending cell in reference order + reason
__global__ void k_check_invalid(const double* __restrict__ v, long stride, long n, const int* __restrict__ old_of_new,
                                unsigned long long* __restrict__ key) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int oi = old_of_new[i];
    if (oi < 0) return;
    const double a0 = v[i], a1 = v[stride + i], a2 = v[2 * stride + i], a3 = v[3 * stride + i], a4 = v[4 * stride + i];
    int reason = 0;
    if (!(isfinite(a0) && isfinite(a1) && isfinite(a2) && isfinite(a3) && isfinite(a4))) reason = 1;
    else if (a0 < 0.0) reason = 2;
    else if (a4 < 0.0) reason = 3;
    if (reason) atomicMin(key, ((unsigned long long)oi << 2) | (unsigned long long)reason);
}

// ------------------------------------------------------------------------------------------------------
// multigrid transfers
// ------------------------------------------------------------------------------------------------------
// mg_restrict (mg_loops.cpp:30-202) as a gather: children summed in ascending original fine index (the reference's
// accumulation order, bit for bit), then multiplied by 1.0/count; coarse nodes without children keep their value.
__global__ void k_restrict(const double* __restrict__ vf, long sfine, double* __restrict__ vc, long scoarse, long ncoarse,
                           const long* __restrict__ child_off, const int* __restrict__ child_ids) {
    const long c = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (c >= ncoarse) return;
    const long k0 = child_off[c], k1 = child_off[c + 1];
    if (k1 == k0) return;
    double s[5] = {0, 0, 0, 0, 0};
    for (long k = k0; k < k1; k++) {
        const long f = child_ids[k];
#pragma unroll
        for (int j = 0; j < 5; j++) s[j] += vf[j * sfine + f];
    }
    const double average = 1.0 / (double)(k1 - k0);
#pragma unroll
    for (int j = 0; j < 5; j++) vc[j * scoarse + c] = s[j] * average;
}
// prolong_residuals_interpolate_proper (mg_loops.cpp:678-864) as a gather over each fine node's incident internal
// edges in original edge order: per edge the own-parent term then the neighbour-parent term (whose source is the own
// parent on the `b` side -- the reference's quirk at :804-810, baked into ent_src by the host).
__global__ void k_prolong(long nfine, long sfine, long scoarse, const int* __restrict__ parent, const double* __restrict__ idist_own,
                          const long* __restrict__ ent_off, const int* __restrict__ ent_src, const double* __restrict__ ent_w,
                          const double* __restrict__ res_c, const double* __restrict__ res_f, double* __restrict__ var_f) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nfine) return;
    const int p = parent[i];
    if (p < 0) return;
    double rp[5];
#pragma unroll
    for (int j = 0; j < 5; j++) rp[j] = res_c[j * scoarse + p];
    const double w0 = idist_own[i];
    double acc[5] = {0, 0, 0, 0, 0};
    double wsum = 0.0;
    const long k0 = ent_off[i], k1 = ent_off[i + 1];
    if (w0 < 0.0) {
        // coincident with its parent: assignment, w_sums = 1 (only if the node has an internal edge at all)
        if (k1 > k0) {
#pragma unroll
            for (int j = 0; j < 5; j++) acc[j] = rp[j];
            wsum = 1.0;
        }
    } else {
        for (long k = k0; k < k1; k++) {
            const int q = ent_src[k];
            const double w = ent_w[k];
#pragma unroll
            for (int j = 0; j < 5; j++) {
                acc[j] += w0 * rp[j];
                acc[j] += w * res_c[j * scoarse + q];
            }
            wsum += w0;
            wsum += w;
        }
    }
#pragma unroll
    for (int j = 0; j < 5; j++) {
        const double avg = acc[j] / wsum;
        var_f[j * sfine + i] += res_f[j * sfine + i] - avg;
    }
}

// ------------------------------------------------------------------------------------------------------
// alternative flux modes and the bandwidth probe
// ------------------------------------------------------------------------------------------------------
// one thread per internal edge, fp64 atomics (RED.ADD.F64): baseline + node-ordering sweep kernel
__global__ void k_flux_atomic(long ne, const int* __restrict__ ea, const int* __restrict__ eb, const double* __restrict__ ew,
                              const double* __restrict__ v, long stride, double* __restrict__ flux, double kdiss) {
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const int a = ea[e], b = eb[e];
    const NodeVals B = load_node(v, stride, b);
    const NodeVals A = load_node(v, stride, a);
    const Flux5 f = edge_flux(A, B, ew[e], ew[ne + e], ew[2 * ne + e], kdiss);
    atomicAdd(&flux[a], f.r); atomicAdd(&flux[stride + a], f.mx); atomicAdd(&flux[2 * stride + a], f.my);
    atomicAdd(&flux[3 * stride + a], f.mz); atomicAdd(&flux[4 * stride + a], f.e);
    atomicAdd(&flux[b], -f.r); atomicAdd(&flux[stride + b], -f.mx); atomicAdd(&flux[2 * stride + b], -f.my);
    atomicAdd(&flux[3 * stride + b], -f.mz); atomicAdd(&flux[4 * stride + b], -f.e);
}
// boundary + wall edges, one thread per edge (a node can carry several, hence atomics)
__global__ void k_bflux_atomic(long nb, const int* __restrict__ bnode, const uint8_t* __restrict__ bkind, const double* __restrict__ bw,
                               const double* __restrict__ v, long stride, double* __restrict__ flux, int mask) {
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= nb) return;
    const int kind = bkind[e];
    if (!((mask >> kind) & 1)) return;
    const int b = bnode[e];
    const NodeVals B = load_node(v, stride, b);
    const Flux5 f = (kind == 1) ? boundary_flux(B, bw[e], bw[nb + e], bw[2 * nb + e]) : wall_flux(B, bw[e], bw[nb + e], bw[2 * nb + e]);
    atomicAdd(&flux[b], f.r); atomicAdd(&flux[stride + b], f.mx); atomicAdd(&flux[2 * stride + b], f.my);
    atomicAdd(&flux[3 * stride + b], f.mz); atomicAdd(&flux[4 * stride + b], f.e);
}
// deterministic sorted-segment mode: one thread per node walks its CSR segment (original edge order) and evaluates
// every incident edge from its own end; nothing is scattered (the reference's FLUX_FISSION + update_edges analogue,
// cfd_loops.cpp:159-213, without materialising per-edge values)
__global__ void k_flux_segment(long n, const long* __restrict__ adj_off, const int* __restrict__ adj_nbr, const double* __restrict__ adj_w,
                               long nadj, const double* __restrict__ v, long stride, double* __restrict__ flux, double kdiss) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const long k0 = adj_off[i], k1 = adj_off[i + 1];
    if (k1 == k0) return;
    const NodeVals A = load_node(v, stride, i);
    double fr = 0, fmx = 0, fmy = 0, fmz = 0, fe = 0;
    for (long k = k0; k < k1; k++) {
        const int raw = adj_nbr[k];
        const double sg = raw < 0 ? -1.0 : 1.0;
        const NodeVals B = load_node(v, stride, raw & 0x7fffffff);
        const Flux5 f = edge_flux(A, B, sg * adj_w[k], sg * adj_w[nadj + k], sg * adj_w[2 * nadj + k], kdiss);
        fr += f.r; fmx += f.mx; fmy += f.my; fmz += f.mz; fe += f.e;
    }
    flux[i] += fr; flux[stride + i] += fmx; flux[2 * stride + i] += fmy; flux[3 * stride + i] += fmz; flux[4 * stride + i] += fe;
}
// indirect_rw (indirect_rw_kernel.elemfunc.c): same gather/scatter as the flux kernel, no arithmetic to speak of
__global__ void k_indirect_rw(long ne, const int* __restrict__ ea, const int* __restrict__ eb, const double* __restrict__ ew,
                              const double* __restrict__ v, long stride, double* __restrict__ flux) {
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const int a = ea[e], b = eb[e];
    const double ex = ew[e], ey = ew[ne + e], ez = ew[2 * ne + e];
    const double ra = v[a], mxa = v[stride + a], mya = v[2 * stride + a], mza = v[3 * stride + a], ea_ = v[4 * stride + a];
    const double rb = v[b], mxb = v[stride + b], myb = v[2 * stride + b], mzb = v[3 * stride + b], eb_ = v[4 * stride + b];
    atomicAdd(&flux[a], rb + ex); atomicAdd(&flux[4 * stride + a], eb_ + ey); atomicAdd(&flux[stride + a], mxb + ez);
    atomicAdd(&flux[2 * stride + a], myb); atomicAdd(&flux[3 * stride + a], mzb);
    atomicAdd(&flux[b], ra); atomicAdd(&flux[4 * stride + b], ea_); atomicAdd(&flux[stride + b], mxa);
    atomicAdd(&flux[2 * stride + b], mya); atomicAdd(&flux[3 * stride + b], mza);
}

// ------------------------------------------------------------------------------------------------------
// layout conversion at the boundary: reference AoS (old order) <-> device SoA (new order, padded)
// ------------------------------------------------------------------------------------------------------
__global__ void k_export_aos(const double* __restrict__ soa, long stride, int ncomp, long nel, const int* __restrict__ new_of_old, double* __restrict__ aos) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nel) return;
    const long g = new_of_old[i];
    for (int k = 0; k < ncomp; k++) aos[i * ncomp + k] = soa[k * stride + g];
}
__global__ void k_import_aos(double* __restrict__ soa, long stride, int ncomp, long nel, const int* __restrict__ new_of_old, const double* __restrict__ aos) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nel) return;
    const long g = new_of_old[i];
    for (int k = 0; k < ncomp; k++) soa[k * stride + g] = aos[i * ncomp + k];
}

}  // namespace mgcfd
