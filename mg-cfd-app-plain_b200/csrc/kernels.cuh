// kernels.cuh -- hand-written fp64 CUDA kernels (sm_100a) for the MG-CFD per-cycle solver loop.
//
// Data layout in HBM (DESIGN.md "Data layout"):
//   node state  : one 64-byte RECORD per node, {rho, mx, my, mz, rhoE, 1/rho, p, |v|+c}: the five conserved variables of the
//                 reference (src/Base/const.h:19-26) plus the three per-node quantities every incident edge needs
//                 (compute_velocity / compute_pressure / compute_speed_of_sound, src/Kernels/cfd_loops.h:121-148), computed
//                 ONCE per node by whichever kernel writes the state instead of twice per edge.  A record is exactly two
//                 32-byte sectors, so the halo gather of a tile moves no padding.
//   residuals / fluxes (granular API) / step factors / volumes : SoA planes of length npad.
//   edge slots  : per tile, per round, one block [hx[TN] | hy[TN] | hz[TN] | other[TN](u16)], h = -0.5 * (edge vector
//                 oriented thread-node -> other); a tile's rounds are contiguous, so the edge stream is read once, coalesced.
// No tensor cores: the path is gather/scatter, FP64- and LSU-bound (profiles/).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace mgcfd {

#define MG_GAMMA 1.4
// the far-field state: ff_variable and ff_flux_contribution_{momentum_x,y,z,density_energy} (globals.h:10-14).  Travels by value in
// the kernel arguments (the parameter constant bank: as cheap as a __constant__ symbol), so that two contexts on one device can hold
// different far fields (the module-wide symbols of round 1 were shared by all contexts)
struct FarField { double v[5]; double c[12]; };

struct Rec { double rho, mx, my, mz, re, ir, p, s; };   // ir = 1/rho, s = |v| + speed of sound
struct Flux5 { double r, mx, my, mz, e; };

// sqrt for positive normal x: MUFU.RSQ64H seed (2^-22) + ONE coupled Newton step (2^-43) + one residual correction, branch-free.
// x is a sum of squares >= 1e-300 here (never 0, inf or denormal), so the special-case paths of sqrt() are dead weight; the
// result is within 1 ulp of the correctly rounded root.  (A second Newton step, as in round 1, adds three dependent FP64
// instructions for nothing: the correction step already squares the error.  FP64 dependent-issue latency is what bounds the
// per-node update, profiles/r02*_timeline_c2.jsonl.)
__device__ __forceinline__ double sqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    double g = __dmul_rn(x, y), h = __dmul_rn(0.5, y);
    const double r = __fma_rn(-h, g, 0.5);
    g = __fma_rn(g, r, g); h = __fma_rn(h, r, h);
    const double d = __fma_rn(-g, g, x);
    return __fma_rn(d, h, g);
}
// 1 / x for positive normal x: MUFU.RCP64H seed (2^-23) + two Newton steps, branch-free, within 1 ulp of the correctly rounded
// reciprocal (the IEEE division it replaces is a ~25-instruction sequence with a 15-deep dependent chain)
__device__ __forceinline__ double rcp_pos(double x) {
    double y;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    y = __fma_rn(__fma_rn(-x, y, 1.0), y, y);
    y = __fma_rn(__fma_rn(-x, y, 1.0), y, y);
    return y;
}

// per-node derived quantities (cfd_loops.h:121-148): velocity = momentum / rho, speed_sqd, pressure, speed of sound.
// The reference divides three times by rho; here one reciprocal (<= 1 ulp) and three products.
__device__ __forceinline__ Rec make_rec(double rho, double mx, double my, double mz, double re) {
    Rec n;
    n.rho = rho; n.mx = mx; n.my = my; n.mz = mz; n.re = re;
    n.ir = rcp_pos(rho);
    const double vx = mx * n.ir, vy = my * n.ir, vz = mz * n.ir;
    const double sq = vx * vx + vy * vy + vz * vz;
    n.p = (double(MG_GAMMA) - double(1.0)) * (re - double(0.5) * rho * sq);
    n.s = sqrt_pos(sq + 1e-300) + sqrt_pos(double(MG_GAMMA) * n.p * n.ir);   // + 1e-300: still fluid (sq == 0) stays finite; vanishes in the sum
    return n;
}

__device__ __forceinline__ Rec load_rec(const double* __restrict__ recs, long i) {
    const double2* p = reinterpret_cast<const double2*>(recs + 8 * i);
    const double2 a = p[0], b = p[1], c = p[2], d = p[3];
    Rec n; n.rho = a.x; n.mx = a.y; n.my = b.x; n.mz = b.y; n.re = c.x; n.ir = c.y; n.p = d.x; n.s = d.y;
    return n;
}
__device__ __forceinline__ void store_rec(double* __restrict__ recs, long i, const Rec& n) {
    double2* p = reinterpret_cast<double2*>(recs + 8 * i);
    p[0] = make_double2(n.rho, n.mx); p[1] = make_double2(n.my, n.mz); p[2] = make_double2(n.re, n.ir); p[3] = make_double2(n.p, n.s);
}
// shared-memory copy of a record: 16-byte chunk k of record o lives at chunk k ^ ((o >> 1) & 3) of its 64-byte row, the
// 64B swizzle pattern, so that the 8 lanes of a quarter-warp reading chunk k of 8 consecutive records hit 8 distinct bank groups
__device__ __forceinline__ void sm_store_rec(double2* sm, int o, const double2& a, const double2& b, const double2& c, const double2& d) {
    double2* row = sm + 4 * o;
    const int x = (o >> 1) & 3;
    row[0 ^ x] = a; row[1 ^ x] = b; row[2 ^ x] = c; row[3 ^ x] = d;
}
__device__ __forceinline__ Rec sm_load_rec(const double2* sm, int o) {
    const double2* row = sm + 4 * o;
    const int x = (o >> 1) & 3;
    const double2 a = row[0 ^ x], b = row[1 ^ x], c = row[2 ^ x], d = row[3 ^ x];
    Rec n; n.rho = a.x; n.mx = a.y; n.my = b.x; n.mz = b.y; n.re = c.x; n.ir = c.y; n.p = d.x; n.s = d.y;
    return n;
}

// ---- TMA bulk copy (cp.async.bulk, SASS UBLKCP) + mbarrier: one thread moves a tile's whole edge stream into shared memory
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");    // make the initialised barrier visible to the async (TMA) proxy
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
// The edge stream is read once per launch and is as large as everything else a level touches together: it is fetched with an
// L2 evict-first policy so that it does not push the node records (gathered ~3x per stage, rewritten every stage) out of L2.
__device__ __forceinline__ void bulk_g2s(void* dst_smem, const void* src_gmem, unsigned bytes, unsigned long long* bar) {
    unsigned long long policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(policy));
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(smem_u32(dst_smem)), "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar)), "l"(policy) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nMG_WAIT:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra MG_DONE;\nbra MG_WAIT;\nMG_DONE:\n}"
                 ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}

// One internal edge seen from end A towards B; (hx,hy,hz) = -0.5 * stored edge vector oriented A->B; k2 = 2*kdiss with
// kdiss = -0.5*smoothing_coefficient (src/Base/common.h:24), so that |e|*kdiss == |h|*k2 exactly.
// Adds A's five flux increments (flux_kernel.elemfunc.c:130-162) to acc; B's increments are their exact negation (:170-189).
// Algebra: sum_d h_d * flux_contribution_momentum_x[d] = mx*(h.v) + p*hx etc. (cfd_loops.h:57-83), with h.v = (h.m)/rho.
__device__ __forceinline__ double edge_weight(double hx, double hy, double hz) {
    return sqrt_pos(__fma_rn(hx, hx, __fma_rn(hy, hy, __fma_rn(hz, hz, 1e-300))));   // + 1e-300: exact no-op unless h == 0 (empty slot)
}
// Written with explicit round-to-nearest intrinsics: the compiler can neither contract nor reassociate them, so every
// instantiation (simple / pipelined kernel, scatter / segment rounds, atomic baseline) produces the same bits for the same edge.
__device__ __forceinline__ void edge_flux_acc_w(const Rec& A, double A_ep, const Rec& B, double hx, double hy, double hz, double ewt, double k2, Flux5& acc) {
    const double factor = __dmul_rn(__dmul_rn(ewt, k2), __dadd_rn(A.s, B.s));
    const double gA = __fma_rn(hz, A.mz, __fma_rn(hy, A.my, __dmul_rn(hx, A.mx)));
    const double gB = __fma_rn(hz, B.mz, __fma_rn(hy, B.my, __dmul_rn(hx, B.mx)));
    const double qA = __dmul_rn(gA, A.ir), qB = __dmul_rn(gB, B.ir);
    const double ps = __dadd_rn(A.p, B.p);
    acc.r  = __fma_rn(factor, __dsub_rn(A.rho, B.rho), __dadd_rn(gA, __dadd_rn(gB, acc.r)));
    acc.e  = __fma_rn(factor, __dsub_rn(A.re, B.re), __fma_rn(A_ep, qA, __fma_rn(__dadd_rn(B.re, B.p), qB, acc.e)));
    acc.mx = __fma_rn(factor, __dsub_rn(A.mx, B.mx), __fma_rn(A.mx, qA, __fma_rn(B.mx, qB, __fma_rn(ps, hx, acc.mx))));
    acc.my = __fma_rn(factor, __dsub_rn(A.my, B.my), __fma_rn(A.my, qA, __fma_rn(B.my, qB, __fma_rn(ps, hy, acc.my))));
    acc.mz = __fma_rn(factor, __dsub_rn(A.mz, B.mz), __fma_rn(A.mz, qA, __fma_rn(B.mz, qB, __fma_rn(ps, hz, acc.mz))));
}
__device__ __forceinline__ void edge_flux_acc(const Rec& A, double A_ep, const Rec& B, double hx, double hy, double hz, double k2, Flux5& acc) {
    edge_flux_acc_w(A, A_ep, B, hx, hy, hz, edge_weight(hx, hy, hz), k2, acc);
}
__device__ __forceinline__ Flux5 edge_flux(const Rec& A, const Rec& B, double hx, double hy, double hz, double k2) {
    Flux5 f = {0.0, 0.0, 0.0, 0.0, 0.0};
    edge_flux_acc(A, A.re + A.p, B, hx, hy, hz, k2, f);
    return f;
}
// boundary edge (neighbour -1): flux_boundary_kernel.elemfunc.c:33-45; (x,y,z) = stored edge vector
__device__ __forceinline__ void boundary_flux_acc(const Rec& B, double x, double y, double z, Flux5& acc) {
    acc.mx += x * B.p; acc.my += y * B.p; acc.mz += z * B.p;
}
// wall edge (neighbour -2, far field): flux_wall_kernel.elemfunc.c:47-69
__device__ __forceinline__ void wall_flux_acc(const FarField& F, const Rec& B, double x, double y, double z, Flux5& acc) {
    const double fx = 0.5 * x, fy = 0.5 * y, fz = 0.5 * z;
    const double g = fx * B.mx + fy * B.my + fz * B.mz;
    const double q = g * B.ir;
    acc.r  += (fx * F.v[1] + fy * F.v[2] + fz * F.v[3]) + g;
    acc.e  += (fx * F.c[9] + fy * F.c[10] + fz * F.c[11]) + (B.re + B.p) * q;
    acc.mx += (fx * F.c[0] + fy * F.c[1] + fz * F.c[2]) + (B.mx * q + B.p * fx);
    acc.my += (fx * F.c[3] + fy * F.c[4] + fz * F.c[5]) + (B.my * q + B.p * fy);
    acc.mz += (fx * F.c[6] + fy * F.c[7] + fz * F.c[8]) + (B.mz * q + B.p * fz);
}

// ------------------------------------------------------------------------------------------------------
// The stage kernel: one CTA = one tile of TN owned nodes (thread t <-> node tile*TN+t), one launch = one Runge-Kutta
// stage of one smoothing visit (compute_flux_edge + compute_boundary_flux_edge + compute_wall_flux_edge + time_step,
// euler3d_cpu_double.cpp:397-506), the flux never touching HBM.
//   phase 1: the records of the owned nodes and of the tile's halo (nodes of other tiles adjacent to it) are staged in
//            shared memory (swizzled rows);
//   phase 2: edge rounds.
//            SCATTER = false (sorted-segment): every node walks ALL its incident edges (its CSR segment, original edge
//              order), each edge is evaluated from both of its ends, nothing is ever written to another node: no colouring,
//              no barriers inside the loop, no atomics, bit-reproducible;
//            SCATTER = true (coloured): an edge inside the tile is evaluated by one of its ends, which adds the negated
//              increment to a shared accumulator of the other end; the host colouring guarantees that within one round no
//              two threads of the CTA target the same node (no atomics, fixed order, bit-reproducible); edges cut by a
//              tile boundary are evaluated by both tiles (owner computes);
//   phase 3: FUSED: time_step (cfd_loops.cpp:215-280) applied directly, the new record (state + derived quantities)
//            written once; on the last stage residual (validation.cpp:77-89), RMS partials (:91-105) and the validity
//            check (:107-138) ride along.  !FUSED (granular API): fluxes[node] += total.
// ------------------------------------------------------------------------------------------------------
// ---- peers of a level in a multi-GPU run (direct peer-to-peer data plane, see k_p2p_exchange below) ----
struct P2PPeer {                   // one entry per peer of a level
    int rank;
    long send0, nsend;             // slice of the level's send list
    long recv0, nrecv;             // slice of my ghost rows filled by this peer
    double* dst[2];                // peer staging (parity 0/1) at the offset where my rows land
    unsigned long long* flag;      // &peer_window.flags[my rank]
};
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
// A kernel launched as a programmatic dependent (cudaLaunchAttributeProgrammaticStreamSerialization) may be scheduled while its
// predecessor in the stream still runs: nothing the predecessor writes may be touched before this returns (it returns at once when
// the kernel was launched the ordinary way).
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Every cross-GPU wait is bounded: a flag that has not arrived after a few seconds (4 M polls of ~1 us: nanosleep + a system-scope
// load) means the ranks disagree on the protocol (or a rank has died) -- the kernel reports and traps, the host sees a CUDA error
// instead of a hung device.
__device__ __forceinline__ bool spin_expired(unsigned& n, const char* what) {
    __nanosleep(40);
    if (++n < 4000000u) return false;
    printf("mgcfd: a multi-GPU wait timed out (%s): block %d thread %d\n", what, (int)blockIdx.x, (int)threadIdx.x);
    __trap();
    return true;
}
// ---- multi-GPU: the kernel that PRODUCES a row also delivers it --------------------------------------------------
// where the copies other ranks hold of this rank's rows live (one entry per peer of the level)
struct PeerOut {
    double* rec[3];                // the peer's three record buffers (same rotation as ours)
    double* res;                   // the peer's residual planes
    long res_stride;               // the peer's npad
};
// (the protocol -- epochs, announcements, late waits -- is described at struct DistTail below)
// all-reduce plumbing of a rank (as k_p2p_allreduce): every rank's window has reduction slots [parity][source rank][8]
struct AllRed {
    int nranks, me;
    double* const* red_of_rank; unsigned long long* const* flag_of_rank;      // [nranks]: window reduction bases, &window.flags[me]
    const unsigned long long* my_flags; const double* my_red;
    unsigned int* red_counter;
};
// all-reduce over ALL ranks of n <= 8 doubles held by lane 0 in v[] (is_min: one positive double, compared as a bit pattern), run by
// one warp (all 32 lanes); deterministic rank order.  epoch = the number this synchronisation carries.
__device__ __forceinline__ void dist_allreduce(const AllRed& d, unsigned long long epoch, double* v, int n, bool is_min) {
    const int lane = threadIdx.x & 31;
    const int parity = int(*(volatile unsigned int*)d.red_counter & 1u);
    double mine[8];
#pragma unroll
    for (int j = 0; j < 8; j++) mine[j] = __shfl_sync(0xffffffffu, j < n ? v[j] : 0.0, 0);
    for (int p = lane; p < d.nranks; p += 32) {
        double* slot = d.red_of_rank[p] + ((size_t)parity * 64 + d.me) * 8;
        for (int j = 0; j < n; j++) slot[j] = mine[j];
        __threadfence_system();
        st_release_sys(d.flag_of_rank[p], epoch);
    }
    for (int p = lane; p < d.nranks; p += 32) {
        unsigned spins = 0;
        while (ld_acquire_sys(d.my_flags + p) < epoch) { if (spin_expired(spins, "all-reduce")) break; }
    }
    __syncwarp();
    if (lane == 0) {
        const double* base = d.my_red + (size_t)parity * 64 * 8;
        if (is_min) {
            unsigned long long m = ~0ull;
            for (int r = 0; r < d.nranks; r++) { const unsigned long long x = (unsigned long long)__double_as_longlong(__ldcg(base + r * 8)); m = x < m ? x : m; }
            v[0] = __longlong_as_double((long long)m);
        } else {
            for (int j = 0; j < n; j++) { double acc = 0.0; for (int r = 0; r < d.nranks; r++) acc += __ldcg(base + r * 8 + j); v[j] = acc; }
        }
        *d.red_counter += 1;
    }
    __syncwarp();
}
// Epochs.  Every kernel of a distributed run that produces rows other ranks hold copies of owns one epoch number; all ranks run the
// same kernel sequence, so the numbers agree.  epoch = *op_counter (a base word on the device, advanced by the host-enqueued
// k_epoch_advance once per V-cycle) + epoch_off (baked into the launch): a replayed CUDA graph needs no per-launch argument.
// A kernel with epoch e
//   * ANNOUNCES, at its start, that this rank's kernels up to e - 1 are complete: block 0 writes e into flags[me] of every
//     neighbour rank.  The completion of a grid makes its stores -- remote ones included -- visible before the next grid of the
//     stream starts, so nothing has to be fenced or counted inside the producing kernel (the first versions did: a system-scope
//     fence per CTA and a grid-wide ticket per kernel cost more than the NVLink latency they guarded, profiles/r02_dist_overhead_history.jsonl);
//   * WAITS until flags[p] >= e for the ranks p it reads ghost rows of -- as late as it can: the stage kernel takes the tiles that
//     read no ghost row first and checks the flags only before it prefetches its first ghost-reading tile, so the one-way flag
//     latency (3.2 us, profiles/r02q_p2p_pingpong.txt) hides behind interior tiles;
//   * stores the rows it produces straight into the peers' copies (dist_push_rec / dist_push_res).
struct DistTail {
    const P2PPeer* wait_peers; int nwait;      // whose rows this kernel reads
    const P2PPeer* sig_peers; int nsig;        // every neighbour rank of this rank (any level): who gets the announcement
    const P2PPeer* peers; int npeers;          // who holds copies of the rows it writes (index space of tgt_peer)
    const PeerOut* peer_out; int ib;           // the peers' record buffer that mirrors the output buffer
    const int* tgt_off; const int* tgt_peer; const int* tgt_row;      // row -> (peer index, row in the peer's arrays)
    const unsigned char* tile_sends;           // per tile of the kernel's granularity: any row with a target
    const int* order; int n_send;              // stage kernels: tiles in the order they are taken, the n_send ghost-reading (= delivering) tiles LAST
    const unsigned char* blk_wait;             // transfer kernels: per 128-row block, does it read a ghost row (only those blocks wait); nullptr: all wait
    int release_early;                         // transfer kernels (one GPU too): let the next kernel of the stream -- a stage kernel launched as a programmatic
                                               // dependent -- be scheduled as blocks of this one exit (its static prologue overlaps this kernel's tail)
    const unsigned long long* op_counter; int epoch_off; const unsigned long long* my_flags;
    // the minimum dt of the state a transfer kernel leaves behind: every block of the transfer kernel folds its minimum into ONE word
    // of the level (atomicMin on the bit pattern of a positive double: fire and forget, no fence, no ticket).  The first stage
    // kernel of the next smoothing visit (recv_min) has its block 0 send that word, tagged with the kernel's epoch, into every
    // rank's reduction slot at its very start; every CTA picks the ranks' values up just before its first update, one NVLink flag
    // latency later.  The second stage kernel of the visit (finish_min) puts the word back to +inf and counts the reduction.
    unsigned long long* minword;
    AllRed ar; int recv_min, finish_min;
    int dbg;       // measurement only (MGCFD_DIST_DEBUG, results become wrong): 1 no waits, 2 no remote stores, 8 no announcements
};
__device__ __forceinline__ void dist_wait(const DistTail& d, unsigned long long e0) {
    if ((int)threadIdx.x < d.nwait && !(d.dbg & 1)) {
        const unsigned long long* f = d.my_flags + d.wait_peers[threadIdx.x].rank;
        unsigned spins = 0;
        while (ld_acquire_sys(f) < e0) { if (spin_expired(spins, "peer epoch")) break; }
    }
    __syncthreads();
}
// start of a distributed kernel: the epoch, the announcement (block 0) and -- unless the caller waits later -- the wait
__device__ __forceinline__ unsigned long long dist_kernel_begin(const DistTail& d, bool wait_now) {
    unsigned long long e0 = 0;
    if (blockIdx.x == 0 || wait_now) e0 = *(volatile const unsigned long long*)d.op_counter + (unsigned long long)d.epoch_off;      // (block-uniform: other blocks never need it)
    if (blockIdx.x == 0 && (int)threadIdx.x < d.nsig && !(d.dbg & 8)) st_release_sys(d.sig_peers[threadIdx.x].flag, e0);
    if (wait_now) dist_wait(d, e0);
    return e0;
}
// the epoch of a stage kernel (every CTA needs it)
__device__ __forceinline__ unsigned long long dist_epoch(const DistTail& d) {
    return *(volatile const unsigned long long*)d.op_counter + (unsigned long long)d.epoch_off;
}
// first stage of a smoothing visit, block 0, first warp: this rank's minimum (left in d.minword by the transfer kernel that produced
// the state) goes to every rank, tagged with this kernel's epoch
__device__ __forceinline__ void dist_min_send(const DistTail& d, unsigned long long e0) {
    const int parity = int(*(volatile unsigned int*)d.ar.red_counter & 1u);
    const double v = __longlong_as_double((long long)*(volatile const unsigned long long*)d.minword);
    for (int p = threadIdx.x; p < d.ar.nranks; p += 32) {
        double* slot = d.ar.red_of_rank[p] + ((size_t)parity * 64 + d.ar.me) * 8;
        slot[0] = v;
        st_release_sys(reinterpret_cast<unsigned long long*>(slot + 1), e0);      // the tag: value delivered (same thread: ordered by the release)
    }
}
// ... and every CTA: the minimum over the ranks of the values tagged e0; every thread returns it.  The slot pair alternates with
// every reduction of the rank (red_counter, advanced by the visit's SECOND stage kernel: no CTA of this one may see it move)
__device__ __forceinline__ double dist_recv_min(const DistTail& d, unsigned long long e0) {
    __shared__ unsigned long long s_min;
    if (threadIdx.x < 32) {
        const int parity = int(*(volatile unsigned int*)d.ar.red_counter & 1u);
        unsigned long long m = ~0ull;
        for (int r = threadIdx.x; r < d.ar.nranks; r += 32) {
            const double* slot = d.ar.my_red + ((size_t)parity * 64 + r) * 8;
            unsigned spins = 0;
            while (!(d.dbg & 1) && ld_acquire_sys(reinterpret_cast<const unsigned long long*>(slot + 1)) != e0) {
                if (spins == 3999998u)
                    printf("mgcfd: rank %d waits for the minimum-dt tag %llu of rank %d: slot holds %llu, reductions so far %u, epoch base %llu off %d\n", d.ar.me, e0, r,
                           *reinterpret_cast<const volatile unsigned long long*>(slot + 1), *(volatile unsigned int*)d.ar.red_counter,
                           *(volatile const unsigned long long*)d.op_counter, d.epoch_off);
                if (spin_expired(spins, "minimum dt tag")) break;
            }
            const unsigned long long x = (unsigned long long)__double_as_longlong(__ldcg(slot));
            m = x < m ? x : m;
        }
#pragma unroll
        for (int dl = 16; dl > 0; dl >>= 1) { const unsigned long long o = __shfl_xor_sync(0xffffffffu, m, dl); m = o < m ? o : m; }
        if (threadIdx.x == 0) s_min = m;
    }
    __syncthreads();
    return __longlong_as_double((long long)s_min);
}
// advances the epoch base by the epochs the kernels enqueued since the last advance have used (one thread, enqueued by the host)
__global__ void k_epoch_advance(unsigned long long* op_counter, int n) { *op_counter += (unsigned long long)n; }
__device__ __forceinline__ void dist_push_rec(const DistTail& d, long row, const Rec& n) {
    if (d.dbg & 2) return;
    for (int k = d.tgt_off[row]; k < d.tgt_off[row + 1]; k++) store_rec(d.peer_out[d.tgt_peer[k]].rec[d.ib], d.tgt_row[k], n);
}
__device__ __forceinline__ void dist_push_res(const DistTail& d, long row, const double r[5]) {
    if (d.dbg & 2) return;
    for (int k = d.tgt_off[row]; k < d.tgt_off[row + 1]; k++) {
        const PeerOut& po = d.peer_out[d.tgt_peer[k]];
        double* pr = po.res + d.tgt_row[k];
        pr[0] = r[0]; pr[po.res_stride] = r[1]; pr[2 * po.res_stride] = r[2]; pr[3 * po.res_stride] = r[3]; pr[4 * po.res_stride] = r[4];
    }
}

// fixed-stride per-tile header (hdr_stride bytes per tile): what a CTA needs to know about a tile before it can fetch it
struct TileHdr {
    int rounds, nh, brounds, pad;
    long long slot_blk0, bslot_blk0;
    // followed by int halo_ids[hpad]
};

struct StageArgs {
    FarField ff;           // the context's far-field state (wall edges)
    const double* vin;     // records: stage input state
    const double* vold;    // records: old_variables (FUSED)
    double* vout;          // records: stage output state (FUSED)
    double* flux;          // SoA fluxes (+=) (!FUSED)
    double* res;           // SoA residuals (last stage) or nullptr
    double* sf;            // step factors: written by stage 0 of a fused smooth (kept for mgcfd_get_field / --output-step-factors)
    const double* vol;     // volumes
    const unsigned long long* min_bits;   // bit pattern of the global minimum dt of this smoothing visit (k_min_dt)
    int legacy;            // compute_step_factor_legacy (mesh_name = fvcorr): no global minimum
    long stride;           // npad
    const unsigned char* hdrs; int hdr_stride;    // TileHdr + halo ids, one per tile
    const unsigned char* slots;   // edge round blocks of TN*26 bytes
    const unsigned char* bslots;  // boundary/wall round blocks of TN*25 bytes
    int ntiles;
    int rec_rows;          // rows of one shared-memory record buffer (TN + padded max halo)
    int chunk_rounds;      // pipelined kernel: edge rounds per ring entry
    double rk_div;         // double(RK+1-j)
    double rk_rcp;         // 1.0 / rk_div (rounded): see div_rk
    double k2;             // 2 * kdiss
    double* rms_partial;   // [ntiles][5] or nullptr
    unsigned long long* bad_key;  // invalid-state key or nullptr
    const int* old_of_new;
    unsigned long long stage_seq;
    int mask;              // bit0 internal, bit1 boundary, bit2 wall
    const double* premin; int npremin;     // first stage: per-block minima of dt left by the transfer kernel that produced vin (or nullptr: *min_bits holds the minimum)
    unsigned long long* minword;           // one GPU, MGCFD_MINWORD=1: the level's minimum dt as ONE word (atomicMin of the transfer kernel's block minima)
    int min_from_word, reset_word;         //   first stage: the minimum is *minword; second stage: block 0 puts the word back to +inf
    DistTail d;            // DIST instantiation of k_stage_pipe: rows are delivered by the kernel itself
};

// A slot's `other` field is the byte offset of logical chunk 0 of the other endpoint's row in the shared record buffer,
// off0 = (o << 6) | (((o >> 1) & 3) << 4); chunk k lives at off0 ^ (k << 4) (the buffer is 64-byte aligned).
__device__ __forceinline__ Rec sm_load_rec_off(const unsigned char* recs, unsigned off0) {
    const double2 c0 = *reinterpret_cast<const double2*>(recs + off0);
    const double2 c1 = *reinterpret_cast<const double2*>(recs + (off0 ^ 16u));
    const double2 c2 = *reinterpret_cast<const double2*>(recs + (off0 ^ 32u));
    const double2 c3 = *reinterpret_cast<const double2*>(recs + (off0 ^ 48u));
    Rec n; n.rho = c0.x; n.mx = c0.y; n.my = c1.x; n.mz = c1.y; n.re = c2.x; n.ir = c2.y; n.p = c3.x; n.s = c3.y;
    return n;
}

// edge rounds of one block sequence (`nr` blocks of TN*26 bytes at `blk`, in shared or global memory)
template <int TN, bool SCATTER>
__device__ __forceinline__ void edge_rounds(const unsigned char* blk, int nr, const unsigned char* recs, double* acc, int t,
                                            const Rec& me, double me_ep, double k2, Flux5& f) {
    if (!SCATTER) {
        // software pipeline: the slot, the other endpoint's record and the edge weight (the sqrt chain) of round r+1 are fetched /
        // computed while round r's flux arithmetic runs; empty slots point at the node itself with h = 0 and add exact zeros
        if (nr <= 0) return;
        // two register sets used alternately (no rotation copies): set 0 holds even rounds, set 1 odd rounds
        const double* w = reinterpret_cast<const double*>(blk);
        double h0x = w[t], h0y = w[TN + t], h0z = w[2 * TN + t];
        Rec B0 = sm_load_rec_off(recs, reinterpret_cast<const unsigned short*>(blk + TN * 24)[t]);
        double e0 = edge_weight(h0x, h0y, h0z);
        int r = 1;
        for (; r + 1 < nr; r += 2) {
            blk += TN * 26;
            const double* w1 = reinterpret_cast<const double*>(blk);
            const double h1x = w1[t], h1y = w1[TN + t], h1z = w1[2 * TN + t];
            const Rec B1 = sm_load_rec_off(recs, reinterpret_cast<const unsigned short*>(blk + TN * 24)[t]);
            const double e1 = edge_weight(h1x, h1y, h1z);
            edge_flux_acc_w(me, me_ep, B0, h0x, h0y, h0z, e0, k2, f);
            blk += TN * 26;
            const double* w2 = reinterpret_cast<const double*>(blk);
            h0x = w2[t]; h0y = w2[TN + t]; h0z = w2[2 * TN + t];
            B0 = sm_load_rec_off(recs, reinterpret_cast<const unsigned short*>(blk + TN * 24)[t]);
            e0 = edge_weight(h0x, h0y, h0z);
            edge_flux_acc_w(me, me_ep, B1, h1x, h1y, h1z, e1, k2, f);
        }
        if (r < nr) {       // one round left to fetch
            blk += TN * 26;
            const double* w1 = reinterpret_cast<const double*>(blk);
            const double h1x = w1[t], h1y = w1[TN + t], h1z = w1[2 * TN + t];
            const Rec B1 = sm_load_rec_off(recs, reinterpret_cast<const unsigned short*>(blk + TN * 24)[t]);
            const double e1 = edge_weight(h1x, h1y, h1z);
            edge_flux_acc_w(me, me_ep, B0, h0x, h0y, h0z, e0, k2, f);
            edge_flux_acc_w(me, me_ep, B1, h1x, h1y, h1z, e1, k2, f);
        } else {
            edge_flux_acc_w(me, me_ep, B0, h0x, h0y, h0z, e0, k2, f);
        }
    } else {
        for (int r = 0; r < nr; r++, blk += TN * 26) {
            const double* w = reinterpret_cast<const double*>(blk);
            const unsigned off0 = reinterpret_cast<const unsigned short*>(blk + TN * 24)[t];
            if (off0 != 0xFFFFu) {
                const double hx = w[t], hy = w[TN + t], hz = w[2 * TN + t];
                const Rec B = sm_load_rec_off(recs, off0);
                Flux5 g = {0.0, 0.0, 0.0, 0.0, 0.0};
                edge_flux_acc(me, me_ep, B, hx, hy, hz, k2, g);
                f.r += g.r; f.mx += g.mx; f.my += g.my; f.mz += g.mz; f.e += g.e;
                const unsigned o = off0 >> 6;
                if (o < (unsigned)TN) {   // owned by this tile: conflict-free by colouring
                    acc[0 * TN + o] -= g.r; acc[1 * TN + o] -= g.mx; acc[2 * TN + o] -= g.my; acc[3 * TN + o] -= g.mz; acc[4 * TN + o] -= g.e;
                }
            }
            __syncthreads();
        }
    }
}
// boundary / wall rounds (global memory; only tiles touching the domain boundary have any).  The slot of the first round is
// fetched by the caller BEFORE the edge loop (bslot_fetch) and the slot of round r+1 while round r is evaluated, so that the two
// dependent global loads (kind, then weights) of a round never sit exposed between the edge loop and the node update.
struct BSlot { int kind; double x, y, z; };
template <int TN>
__device__ __forceinline__ BSlot bslot_fetch(const unsigned char* blk, int t) {
    BSlot b;
    const double* w = reinterpret_cast<const double*>(blk);
    b.kind = blk[TN * 24 + t];
    b.x = w[t]; b.y = w[TN + t]; b.z = w[2 * TN + t];
    return b;
}
template <int TN>
__device__ __forceinline__ void boundary_rounds(const FarField& F, const unsigned char* blk, int br, int t, int mask, const Rec& me, Flux5& f, BSlot cur) {
    for (int r = 0; r < br; r++) {
        BSlot nxt = {0, 0.0, 0.0, 0.0};
        if (r + 1 < br) nxt = bslot_fetch<TN>(blk + (size_t)(r + 1) * (TN * 25), t);
        if (cur.kind != 0 && ((mask >> cur.kind) & 1)) {
            if (cur.kind == 1) boundary_flux_acc(me, cur.x, cur.y, cur.z, f);
            else wall_flux_acc(F, me, cur.x, cur.y, cur.z, f);
        }
        cur = nxt;
    }
}
template <int TN>
__device__ __forceinline__ void boundary_rounds(const FarField& F, const unsigned char* blk, int br, int t, int mask, const Rec& me, Flux5& f) {
    if (br > 0) boundary_rounds<TN>(F, blk, br, t, mask, me, f, bslot_fetch<TN>(blk, t));
}
// step factor of one node (cfd_loops.cpp:146-156 / :60): min_dt / volume, or the legacy local form.  Evaluated by the FIRST stage
// of a smoothing visit (vold == vin), which stores it; the later stages read that value back (same bits, no second division).
__device__ __forceinline__ double step_factor_of(const StageArgs& a, double min_dt, double vol, double s_old) {
    if (a.legacy) return double(0.5) / (sqrt(vol) * s_old);
    return min_dt / vol;
}
// x / d for d = double(RK+1-j) in {4, 3, 2} with rd = RN(1/d): q = RN(x*rd), r = x - d*q (exact, fma), result RN(q + r*rd) -- the
// correctly rounded quotient for every normal x (Markstein; checked against `/` on 6e8 random operands), i.e. the reference's
// true divide (cfd_loops.cpp:243) bit for bit in three FP64 instructions instead of the ~25-instruction division sequence
__device__ __forceinline__ double div_rk(double x, double d, double rd) {
    const double q = __dmul_rn(x, rd);
    return __fma_rn(__fma_rn(-d, q, x), rd, q);
}
// phase 3 of a fused stage for node gid: time_step + record + validity + residual; returns the five squared residuals in q
template <bool DIST = false>
__device__ __forceinline__ void fused_update(const StageArgs& a, long gid, double sf, const double o[5], const Flux5& f, double q[5], bool sends = false) {
    const long S = a.stride;
    const double factor = div_rk(sf, a.rk_div, a.rk_rcp);     // == sf / rk_div, the true divide of cfd_loops.cpp:243
    const double n0 = o[0] + factor * f.r, n1 = o[1] + factor * f.mx, n2 = o[2] + factor * f.my, n3 = o[3] + factor * f.mz, n4 = o[4] + factor * f.e;
    const Rec nrec = make_rec(n0, n1, n2, n3, n4);
    store_rec(a.vout, gid, nrec);
    if (DIST && sends) dist_push_rec(a.d, gid, nrec);      // the same record into the copies other ranks hold of this node (over NVLink)
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
        const double r0 = n0 - o[0], r1 = n1 - o[1], r2 = n2 - o[2], r3 = n3 - o[3], r4 = n4 - o[4];   // residual(), validation.cpp:77-89
        a.res[gid] = r0; a.res[S + gid] = r1; a.res[2 * S + gid] = r2; a.res[3 * S + gid] = r3; a.res[4 * S + gid] = r4;
        if (DIST && sends) { const double rr[5] = {r0, r1, r2, r3, r4}; dist_push_res(a.d, gid, rr); }      // prolong on the other ranks reads them
        q[0] = r0 * r0; q[1] = r1 * r1; q[2] = r2 * r2; q[3] = r3 * r3; q[4] = r4 * r4;
    }
}
// deterministic block reduction of r^2 per variable (calc_rms partials): warp shuffles, then 5 threads over the warp sums
template <int TN>
__device__ __forceinline__ void rms_block(double q[5], double (*ws)[32], int t, double* out5) {
#pragma unroll
    for (int k = 0; k < 5; k++) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) q[k] += __shfl_down_sync(0xffffffffu, q[k], d);
    }
    if ((t & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 5; k++) ws[k][t >> 5] = q[k];
    }
    __syncthreads();
    if (t < 5) {
        double s = 0.0;
        for (int w = 0; w < TN / 32; w++) s += ws[t][w];
        out5[t] = s;
    }
}

// ---- simple form: one CTA per tile, loads issued where they are needed (granular API; reference point for the pipeline) ----
template <int TN, bool SCATTER, bool FUSED>
__global__ void __launch_bounds__(TN, (TN <= 128 ? 5 : (TN <= 256 ? 2 : 1)))
k_stage(const StageArgs a) {
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ double ws[5][32];
    const int t = threadIdx.x;
    const long tile = blockIdx.x;
    const long gid = tile * TN + t;
    const TileHdr* hdr = reinterpret_cast<const TileHdr*>(a.hdrs + tile * (long)a.hdr_stride);
    const int* ids = reinterpret_cast<const int*>(hdr + 1);
    const int nh = hdr->nh;
    double2* st = reinterpret_cast<double2*>(smraw);          // [rec_rows records][4 chunks]
    double* acc = reinterpret_cast<double*>(smraw + 64 * (size_t)a.rec_rows);   // SCATTER: [5][TN]

    Rec me;
    {
        const double2* p = reinterpret_cast<const double2*>(a.vin + 8 * gid);
        const double2 c0 = p[0], c1 = p[1], c2 = p[2], c3 = p[3];
        sm_store_rec(st, t, c0, c1, c2, c3);
        me.rho = c0.x; me.mx = c0.y; me.my = c1.x; me.mz = c1.y; me.re = c2.x; me.ir = c2.y; me.p = c3.x; me.s = c3.y;
    }
    for (int h = t; h < nh; h += TN) {
        const double2* p = reinterpret_cast<const double2*>(a.vin + 8 * (long)ids[h]);
        const double2 c0 = p[0], c1 = p[1], c2 = p[2], c3 = p[3];
        sm_store_rec(st, TN + h, c0, c1, c2, c3);
    }
    if (SCATTER) {
#pragma unroll
        for (int k = 0; k < 5; k++) acc[k * TN + t] = 0.0;
    }
    __syncthreads();

    Flux5 f = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (a.mask & 1) edge_rounds<TN, SCATTER>(a.slots + hdr->slot_blk0 * (long)(TN * 26), hdr->rounds, smraw, acc, t, me, me.re + me.p, a.k2, f);
    if (a.mask & 6) boundary_rounds<TN>(a.ff, a.bslots + hdr->bslot_blk0 * (long)(TN * 25), hdr->brounds, t, a.mask, me, f);
    if (SCATTER) { f.r += acc[0 * TN + t]; f.mx += acc[1 * TN + t]; f.my += acc[2 * TN + t]; f.mz += acc[3 * TN + t]; f.e += acc[4 * TN + t]; }

    if (!FUSED) {
        const long S = a.stride;
        a.flux[gid] += f.r; a.flux[S + gid] += f.mx; a.flux[2 * S + gid] += f.my; a.flux[3 * S + gid] += f.mz; a.flux[4 * S + gid] += f.e;
    } else {
        double o[5], s_old = me.s;
        if (a.vold == a.vin) { o[0] = me.rho; o[1] = me.mx; o[2] = me.my; o[3] = me.mz; o[4] = me.re; }
        else {
            const double2* p = reinterpret_cast<const double2*>(a.vold + 8 * gid);
            const double2 c0 = p[0], c1 = p[1];
            o[0] = c0.x; o[1] = c0.y; o[2] = c1.x; o[3] = c1.y; o[4] = a.vold[8 * gid + 4];
        }
        double sf;
        if (a.vold == a.vin) { sf = step_factor_of(a, a.legacy ? 0.0 : __longlong_as_double((long long)*a.min_bits), a.vol[gid], s_old); a.sf[gid] = sf; }
        else sf = a.sf[gid];
        double q[5] = {0, 0, 0, 0, 0};
        fused_update(a, gid, sf, o, f, q);
        if (a.res && a.rms_partial) rms_block<TN>(q, ws, t, a.rms_partial + tile * 5);
    }
}

// ---- pipelined form: persistent CTAs, every global->shared transfer asynchronous and issued ahead of its use ----
//   * tiles blockIdx.x, blockIdx.x + gridDim.x, ... ; iteration `it` computes tile T_it
//   * headers (+ halo ids) of T_{it+2}: cp.async (LDGSTS) at iteration it           (3 header buffers)
//   * records of T_{it+1} (owned rows + halo rows, swizzled): cp.async at iteration it, completion tracked by an
//     mbarrier through cp.async.mbarrier.arrive                                        (2 record buffers)
//   * the edge stream: a ring of RING entries of `chunk_rounds` round blocks, each filled by ONE TMA bulk copy
//     (cp.async.bulk, SASS UBLKCP) issued by thread 0 as soon as the entry is free, completion by mbarrier complete_tx
//   so the edge loop reads shared memory only and, in steady state, never waits on HBM or L2.
__device__ __forceinline__ void cp_async16(void* dst_smem, const void* src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_u32(dst_smem)), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_mbar_arrive(unsigned long long* bar) {
    asm volatile("cp.async.mbarrier.arrive.noinc.shared::cta.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

constexpr int RING = 2;     // ring entries (each `chunk_rounds` round blocks)

template <int TN, bool SCATTER, bool DIST = false>
__global__ void __launch_bounds__(TN, (TN <= 128 ? 4 : (TN <= 256 ? 2 : 1)))
k_stage_pipe(const StageArgs a) {
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ __align__(8) unsigned long long bar_ring[RING], bar_recs[2];
    __shared__ double ws[5][32];
    const int t = threadIdx.x;
    const int G = gridDim.x;
    const int my_count = (a.ntiles - (int)blockIdx.x + G - 1) / G;
    const int R = a.chunk_rounds;
    const unsigned ring_bytes = unsigned(R) * unsigned(TN * 26);
    unsigned char* ring = smraw;                                                // RING * ring_bytes (ring_bytes is a multiple of 128)
    unsigned char* recs = ring + RING * (size_t)ring_bytes;                     // 2 * rec_rows * 64
    unsigned char* hdrs = recs + 2 * 64 * (size_t)a.rec_rows;                   // 3 * hdr_stride
    double* acc = reinterpret_cast<double*>(hdrs + 3 * (size_t)a.hdr_stride);   // SCATTER: [5][TN]

    if (t == 0) {
        for (int i = 0; i < RING; i++) mbar_init(&bar_ring[i], 1);
        mbar_init(&bar_recs[0], TN); mbar_init(&bar_recs[1], TN);
    }
    __syncthreads();

    // DIST: tiles are taken in an order that puts the tiles that read ghost rows (= the tiles that deliver rows) LAST (DistTail::order):
    // the CTA checks the peers' flags only before it prefetches the records of its first such tile
    // (the CTA's slice of the order table is staged in shared memory by the static prologue: a global load per tile_of() call sat
    // on the header-prefetch chain of every tile, profiles/r02u_distperf.jsonl)
    constexpr int ORDER_CACHE = 96;
    __shared__ int s_order[DIST ? ORDER_CACHE : 1];
    if (DIST) {
        for (int k = t; k < my_count && k < ORDER_CACHE; k += TN) s_order[k] = a.d.order[(long)blockIdx.x + (long)k * G];
        __syncthreads();
    }
    auto tile_of = [&](int it) -> long {
        const long k = (long)blockIdx.x + (long)it * G;
        if (!DIST) return k;
        return it < ORDER_CACHE ? (long)s_order[it] : (long)a.d.order[k];
    };
    const long first_ghost_pos = DIST ? (long)a.ntiles - a.d.n_send : 0;
    bool waited = false;
    auto hdr_of = [&](int it) -> const TileHdr* { return reinterpret_cast<const TileHdr*>(hdrs + (it % 3) * (size_t)a.hdr_stride); };
    auto copy_hdr = [&](int it) {
        if (it >= my_count) return;
        const unsigned char* src = a.hdrs + tile_of(it) * (long)a.hdr_stride;
        unsigned char* dst = hdrs + (it % 3) * (size_t)a.hdr_stride;
        for (int c = t * 16; c < a.hdr_stride; c += TN * 16) cp_async16(dst + c, src + c);
    };
    auto copy_recs = [&](int it) {      // needs the header of T_it in shared memory
        if (it >= my_count) return;
        const TileHdr* hd = hdr_of(it);
        const int* ids = reinterpret_cast<const int*>(hd + 1);
        unsigned char* buf = recs + (it & 1) * 64 * (size_t)a.rec_rows;
        // four lanes per record (one 16-byte chunk each): a warp instruction moves 8 whole 64-byte records, i.e. full sectors
        const int k16 = (t & 3) << 4;
        {
            const unsigned char* src = reinterpret_cast<const unsigned char*>(a.vin + 8 * (tile_of(it) * TN));
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const int row = (t >> 2) + j * (TN / 4);
                cp_async16(buf + 64 * row + (k16 ^ (((row >> 1) & 3) << 4)), src + 64 * row + k16);
            }
        }
        const int nh = hd->nh;
        for (int i = t; i < 4 * nh; i += TN) {
            const int h = i >> 2;
            const unsigned char* src = reinterpret_cast<const unsigned char*>(a.vin + 8 * (long)ids[h]);
            const int o = TN + h;
            cp_async16(buf + 64 * o + (k16 ^ (((o >> 1) & 3) << 4)), src + k16);
        }
        cp_async_mbar_arrive(&bar_recs[it & 1]);
    };
    // edge-stream producer state (thread 0): next chunk to issue = chunk `p_chunk` of tile iteration `p_it`
    int p_it = 0, p_chunk = 0, issued = 0, consumed = 0;
    auto produce = [&](int it_visible) {          // headers of T_0..T_it_visible are in shared memory
        while (issued - consumed < RING && p_it <= it_visible && p_it < my_count) {
            const TileHdr* hd = hdr_of(p_it);
            const int rounds = (a.mask & 1) ? hd->rounds : 0;
            const int nchunks = (rounds + R - 1) / R;
            if (p_chunk >= nchunks) { p_it++; p_chunk = 0; continue; }
            const int nr = min(R, rounds - p_chunk * R);
            const unsigned bytes = unsigned(nr) * unsigned(TN * 26);
            const int e = issued % RING;
            mbar_expect_tx(&bar_ring[e], bytes);
            bulk_g2s(ring + e * (size_t)ring_bytes, a.slots + (hd->slot_blk0 + (long)p_chunk * R) * (long)(TN * 26), bytes, &bar_ring[e]);
            issued++; p_chunk++;
        }
    };

    // prologue: headers of T_0 and T_1 and the first edge chunks are static data: with programmatic dependent launch they are
    // fetched while the previous kernel of the stream is still draining; node state is touched only after griddepcontrol.wait
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    copy_hdr(0); copy_hdr(1);
    cp_async_wait_all();
    __syncthreads();
    if (t == 0) produce(1);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned long long e0 = 0;
    if (DIST) {
        e0 = dist_epoch(a.d);
        if (blockIdx.x == 0 && t < a.d.nsig && !(a.d.dbg & 8)) st_release_sys(a.d.sig_peers[t].flag, e0);      // the announcement
        if (blockIdx.x == 0 && t < 32 && a.d.recv_min) dist_min_send(a.d, e0);
        if (blockIdx.x == 0 && t == 0 && a.d.finish_min) { *a.d.minword = 0x7F7F7F7F7F7F7F7FULL; *a.d.ar.red_counter += 1; }
        if ((long)blockIdx.x >= first_ghost_pos) { dist_wait(a.d, e0); waited = true; }      // the very first tile already reads ghost rows
    }
    copy_recs(0);
    const bool first_stage = (a.vold == a.vin);
    // the visit's global minimum dt: *min_bits (k_min_dt, or -- multi-GPU -- the all-reduce at the end of the transfer kernel that
    // produced this state), or the per-block minima that transfer kernel left behind, reduced here by every CTA for itself
    double min_dt = 0.0;
    const bool recv_min_late = DIST && first_stage && !a.legacy && a.d.recv_min;      // picked up before the first update, behind the first edge rounds
    if (!DIST && a.reset_word && blockIdx.x == 0 && t == 0) *a.minword = 0x7F7F7F7F7F7F7F7FULL;      // consumed by the first stage (complete: we are past griddepcontrol.wait)
    if (first_stage && !a.legacy && !recv_min_late) {
        if (!DIST && a.min_from_word) min_dt = __longlong_as_double((long long)*(volatile const unsigned long long*)a.minword);
        else if (a.premin) {
            double v = __longlong_as_double(0x7F7F7F7F7F7F7F7FLL);
            for (int b = t; b < a.npremin; b += TN) v = fmin(v, __ldcg(a.premin + b));
#pragma unroll
            for (int dl = 16; dl > 0; dl >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, dl));
            if ((t & 31) == 0) ws[0][t >> 5] = v;
            __syncthreads();
            v = ((t & 31) < TN / 32) ? ws[0][t & 31] : __longlong_as_double(0x7F7F7F7F7F7F7F7FLL);
#pragma unroll
            for (int dl = 16; dl > 0; dl >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, dl));
            min_dt = v;
            __syncthreads();          // ws is reused by the RMS sums
        } else min_dt = __longlong_as_double((long long)*a.min_bits);
    }

    for (int it = 0; it < my_count; it++) {
        const long tile = tile_of(it);
        const long gid = tile * TN + t;
        const TileHdr* hd = hdr_of(it);
        copy_hdr(it + 2);
        mbar_wait(&bar_recs[it & 1], (it >> 1) & 1);       // records of T_it and the header of T_{it+1} have landed
        if (DIST && !waited && it + 1 < my_count && (long)blockIdx.x + (long)(it + 1) * G >= first_ghost_pos) { dist_wait(a.d, e0); waited = true; }
        copy_recs(it + 1);
        if (t == 0) produce(it + 1);
        // early loads for the epilogue: volume (first stage) or the stored step factor (later stages), the old state, and the
        // first boundary / wall slot of the tile -- all in flight while the edge rounds run
        const double vol_or_sf = first_stage ? a.vol[gid] : a.sf[gid];
        const int brounds = (a.mask & 6) ? hd->brounds : 0;
        const unsigned char* bblk = a.bslots + hd->bslot_blk0 * (long)(TN * 25);
        BSlot b0 = {0, 0.0, 0.0, 0.0};
        if (brounds > 0) b0 = bslot_fetch<TN>(bblk, t);
        double o[5];
        const unsigned char* buf = recs + (it & 1) * 64 * (size_t)a.rec_rows;
        const Rec me = sm_load_rec_off(buf, (unsigned(t) << 6) | (((unsigned(t) >> 1) & 3u) << 4));
        if (first_stage) { o[0] = me.rho; o[1] = me.mx; o[2] = me.my; o[3] = me.mz; o[4] = me.re; }
        else {
            const double2* p = reinterpret_cast<const double2*>(a.vold + 8 * gid);
            const double2 c0 = p[0], c1 = p[1];
            o[0] = c0.x; o[1] = c0.y; o[2] = c1.x; o[3] = c1.y; o[4] = a.vold[8 * gid + 4];
        }
        if (SCATTER) {
#pragma unroll
            for (int k = 0; k < 5; k++) acc[k * TN + t] = 0.0;
            __syncthreads();
        }
        Flux5 f = {0.0, 0.0, 0.0, 0.0, 0.0};
        const double me_ep = me.re + me.p;
        const int rounds = (a.mask & 1) ? hd->rounds : 0;
        for (int r0 = 0; r0 < rounds; r0 += R) {
            const int e = consumed % RING;
            mbar_wait(&bar_ring[e], (consumed / RING) & 1);
            edge_rounds<TN, SCATTER>(ring + e * (size_t)ring_bytes, min(R, rounds - r0), buf, acc, t, me, me_ep, a.k2, f);
            // hand the ring entry back.  Only the producer has to know that EVERY warp is done with it: warp 0 waits on a named
            // barrier, the other warps just arrive and run on (two barrier ids alternate so that a fast warp's next arrival
            // cannot be counted into a generation a slow warp has not reached; it cannot get further ahead than that because
            // the entry after next is only refilled once this hand-over has completed)
            if (SCATTER) __syncthreads();                   // the coloured rounds synchronise anyway
            else if (t < 32) asm volatile("barrier.cta.sync %0, %1;" ::"r"(1 + (consumed & 1)), "n"(TN) : "memory");
            else asm volatile("barrier.cta.arrive %0, %1;" ::"r"(1 + (consumed & 1)), "n"(TN) : "memory");
            consumed++;
            if (t == 0) produce(it + 1);
        }
        boundary_rounds<TN>(a.ff, bblk, brounds, t, a.mask, me, f, b0);
        if (SCATTER) { f.r += acc[0 * TN + t]; f.mx += acc[1 * TN + t]; f.my += acc[2 * TN + t]; f.mz += acc[3 * TN + t]; f.e += acc[4 * TN + t]; }
        if (DIST && recv_min_late && it == 0) min_dt = dist_recv_min(a.d, e0);       // the ranks' minima, sent by their last transfer kernel
        double sf = vol_or_sf;
        if (first_stage) { sf = step_factor_of(a, min_dt, vol_or_sf, me.s); a.sf[gid] = sf; }
        double q[5] = {0, 0, 0, 0, 0};
        fused_update<DIST>(a, gid, sf, o, f, q, DIST && (long)blockIdx.x + (long)it * G >= first_ghost_pos);      // (the delivering tiles are the last n_send of the order)
        if (a.res && a.rms_partial) rms_block<TN>(q, ws, t, a.rms_partial + tile * 5);
        __syncthreads();      // record buffer (it & 1), header buffer (it % 3), acc and ws are free again
    }
}

// ------------------------------------------------------------------------------------------------------
// node kernels
// ------------------------------------------------------------------------------------------------------
// compute_step_factor (cfd_loops.cpp:76-157) part 1 / compute_step_factor_legacy (:13-73); speed + speed_of_sound is the
// record's s
template <bool LEGACY>
__global__ void k_step_factor(const double* __restrict__ recs, long n, const double* __restrict__ vol_root,
                              double* __restrict__ sf, unsigned long long* __restrict__ min_bits) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    double val = __longlong_as_double(0x7F7F7F7F7F7F7F7FLL);
    if (i < n) {
        const double s = recs[8 * i + 7];
        if (LEGACY) {
            sf[i] = double(0.5) / (sqrt(vol_root[i]) * s);   // vol_root = volumes here; IEEE sqrt, cfd_loops.cpp:60
        } else {
            const double dt = vol_root[i] / s;               // vol_root = cbrt(volume) (host glibc), :123
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

// fused path: the global minimum of 0.5 * cbrt(vol) / (|v| + c) (cfd_loops.cpp:123-145) in ONE launch -- block minima, then the
// last block to finish (ticket counter) reduces them and publishes the bit pattern; nothing to reset between launches
__global__ void k_min_dt(const double* __restrict__ recs, long n, const double* __restrict__ vol_root, double* __restrict__ blockmins,
                         unsigned int* __restrict__ ticket, unsigned long long* __restrict__ min_bits) {
    const double BIG = __longlong_as_double(0x7F7F7F7F7F7F7F7FLL);
    __shared__ double wmin[32];
    __shared__ bool last;
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    double val = BIG;
    if (i < n) val = 0.5 * (vol_root[i] / recs[8 * i + 7]);
    auto block_min = [&](double v) {
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, d));
        if ((threadIdx.x & 31) == 0) wmin[threadIdx.x >> 5] = v;
        __syncthreads();
        v = (threadIdx.x < (blockDim.x >> 5)) ? wmin[threadIdx.x] : BIG;
        if (threadIdx.x < 32) {
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) v = fmin(v, __shfl_xor_sync(0xffffffffu, v, d));
        }
        __syncthreads();
        return v;      // valid in thread 0
    };
    val = block_min(val);
    if (threadIdx.x == 0) {
        blockmins[blockIdx.x] = val;
        __threadfence();
        last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);     // wraps back to 0 for the next launch
    }
    __syncthreads();
    if (!last) return;
    __threadfence();
    double v = BIG;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) v = fmin(v, ((volatile double*)blockmins)[b]);
    v = block_min(v);
    if (threadIdx.x == 0) *min_bits = (unsigned long long)__double_as_longlong(v);
}

// time_step (cfd_loops.cpp:215-280), granular API
__global__ void k_time_step(double rk_div, long n, long stride, const double* __restrict__ sf, double* __restrict__ flux,
                            const double* __restrict__ vold, double* __restrict__ v) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const double factor = sf[i] / rk_div;
    double nv[5];
#pragma unroll
    for (int k = 0; k < 5; k++) {
        nv[k] = vold[8 * i + k] + factor * flux[k * stride + i];
        flux[k * stride + i] = 0.0;
    }
    store_rec(v, i, make_rec(nv[0], nv[1], nv[2], nv[3], nv[4]));
}
__global__ void k_fill(double* __restrict__ p, long n, double val) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) p[i] = val;
}
__global__ void k_fill_state(double* __restrict__ recs, long n, const FarField F) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    store_rec(recs, i, make_rec(F.v[0], F.v[1], F.v[2], F.v[3], F.v[4]));
}
__global__ void k_copy(double* __restrict__ dst, const double* __restrict__ src, long n) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[i];
}
// residual (validation.cpp:77-89)
__global__ void k_residual(long n, long stride, const double* __restrict__ vold, const double* __restrict__ v, double* __restrict__ r) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
#pragma unroll
    for (int k = 0; k < 5; k++) r[k * stride + i] = v[8 * i + k] - vold[8 * i + k];
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
    pdl_wait();
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
__global__ void k_check_invalid(const double* __restrict__ recs, long n, const int* __restrict__ old_of_new,
                                unsigned long long* __restrict__ key) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int oi = old_of_new[i];
    if (oi < 0) return;
    const double a0 = recs[8 * i], a1 = recs[8 * i + 1], a2 = recs[8 * i + 2], a3 = recs[8 * i + 3], a4 = recs[8 * i + 4];
    int reason = 0;
    if (!(isfinite(a0) && isfinite(a1) && isfinite(a2) && isfinite(a3) && isfinite(a4))) reason = 1;
    else if (a0 < 0.0) reason = 2;
    else if (a4 < 0.0) reason = 3;
    if (reason) atomicMin(key, ((unsigned long long)oi << 2) | (unsigned long long)reason);
}

// ------------------------------------------------------------------------------------------------------
// multigrid transfers
// ------------------------------------------------------------------------------------------------------
// The transfer kernels leave the per-block minima of 0.5 * cbrt(vol) / (|v| + c) of the state they produce behind (blockDim.x = 128):
// the next smoothing visit of that level needs exactly this minimum (compute_step_factor, cfd_loops.cpp:123-145) and no longer has to
// pass over the nodes for it.  One shuffle reduction, one barrier and one store per block -- no atomics, no ticket.
__device__ __forceinline__ void block_min_store(double val, double* __restrict__ out) {
    __shared__ double wm[4];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) val = fmin(val, __shfl_xor_sync(0xffffffffu, val, d));
    if ((threadIdx.x & 31) == 0) wm[threadIdx.x >> 5] = val;
    __syncthreads();
    if (threadIdx.x == 0) out[blockIdx.x] = fmin(fmin(wm[0], wm[1]), fmin(wm[2], wm[3]));
}

// ... or (multi-GPU) folded into one word of the level: atomicMin on the bit pattern (positive doubles order like unsigned integers)
__device__ __forceinline__ void block_min_atomic(double val, unsigned long long* __restrict__ word) {
    __shared__ double wm2[4];
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) val = fmin(val, __shfl_xor_sync(0xffffffffu, val, d));
    if ((threadIdx.x & 31) == 0) wm2[threadIdx.x >> 5] = val;
    __syncthreads();
    if (threadIdx.x == 0) atomicMin(word, (unsigned long long)__double_as_longlong(fmin(fmin(wm2[0], wm2[1]), fmin(wm2[2], wm2[3]))));
}

// mg_restrict (mg_loops.cpp:30-202) as a gather: children summed in ascending original fine index (the reference's
// accumulation order, bit for bit), then multiplied by 1.0/count; coarse nodes without children keep their value.
template <bool DIST>
__global__ void k_restrict(const double* __restrict__ vf, double* __restrict__ vc, long ncoarse,
                           const long* __restrict__ child_off, const int* __restrict__ child_ids, const double* __restrict__ vol_root,
                           double* __restrict__ blockmins, const DistTail d) {
    pdl_wait();
    if (d.release_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (DIST) dist_kernel_begin(d, !d.blk_wait || d.blk_wait[blockIdx.x] != 0);      // only blocks with ghost children wait for the fine level's owners
    const long c = blockIdx.x * (long)blockDim.x + threadIdx.x;
    double s_new = 0.0;          // |v| + c of the node's state after this kernel
    const bool want_min = blockmins || d.minword;
    const double vroot = (want_min && c < ncoarse) ? vol_root[c] : 0.0;      // requested first: in flight behind the gather chain
    if (c < ncoarse) {
        const long k0 = child_off[c], k1 = child_off[c + 1];
        if (k1 > k0) {
            double s[5] = {0, 0, 0, 0, 0};
            for (long k = k0; k < k1; k++) {
                const double2* p = reinterpret_cast<const double2*>(vf + 8 * (long)child_ids[k]);
                const double2 c0 = p[0], c1 = p[1];
                const double e = vf[8 * (long)child_ids[k] + 4];
                s[0] += c0.x; s[1] += c0.y; s[2] += c1.x; s[3] += c1.y; s[4] += e;
            }
            const double average = 1.0 / (double)(k1 - k0);
            const Rec n = make_rec(s[0] * average, s[1] * average, s[2] * average, s[3] * average, s[4] * average);
            store_rec(vc, c, n);
            s_new = n.s;
            if (DIST) dist_push_rec(d, c, n);
        } else {
            if (want_min) s_new = vc[8 * c + 7];
            // a childless coarse node keeps its value; the copies other ranks hold of it must keep up with whatever the last
            // visit left in THIS buffer of theirs (their ghost rows are only ever written by the owner)
            if (DIST) { if (d.tgt_off[c + 1] > d.tgt_off[c]) dist_push_rec(d, c, load_rec(vc, c)); }
        }
    }
    if (blockmins) block_min_store(c < ncoarse ? 0.5 * (vroot / s_new) : __longlong_as_double(0x7F7F7F7F7F7F7F7FLL), blockmins);
    if (d.minword) block_min_atomic(c < ncoarse ? 0.5 * (vroot / s_new) : __longlong_as_double(0x7F7F7F7F7F7F7F7FLL), d.minword);
}
// prolong_residuals_interpolate_proper (mg_loops.cpp:678-864) as a gather over each fine node's incident internal
// edges in original edge order: per edge the own-parent term then the neighbour-parent term (whose source is the own
// parent on the `b` side -- the reference's quirk at :804-810, baked into ent_src by the host).
template <bool DIST>
__global__ void k_prolong(long nfine, long sfine, long scoarse, const int* __restrict__ parent, const double* __restrict__ idist_own,
                          const long* __restrict__ ent_off, const int* __restrict__ ent_src, const double* __restrict__ ent_w,
                          const double* __restrict__ res_c, const double* __restrict__ res_f, double* __restrict__ var_f,
                          const double* __restrict__ vol_root, double* __restrict__ blockmins, const DistTail d) {
    pdl_wait();
    if (d.release_early) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    if (DIST) dist_kernel_begin(d, !d.blk_wait || d.blk_wait[blockIdx.x] != 0);      // only blocks that read a ghost parent's residual wait
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    const int p = (i < nfine) ? parent[i] : -1;
    double dt_new = __longlong_as_double(0x7F7F7F7F7F7F7F7FLL);      // padding rows (no parent) never win the minimum
    if (p >= 0) {
        const bool want_min = blockmins || d.minword;
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
#pragma unroll 4
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
        // the row's own state and residual: all ten loads requested together, ahead of the five divisions (one exposed latency
        // instead of one per variable: the compiler kept each load next to its use, ncu source page of round 2)
        double vf0[5], rf0[5];
        {
            const double2* q = reinterpret_cast<const double2*>(var_f + 8 * i);
            const double2 a = q[0], b = q[1];
            vf0[0] = a.x; vf0[1] = a.y; vf0[2] = b.x; vf0[3] = b.y; vf0[4] = var_f[8 * i + 4];
        }
#pragma unroll
        for (int j = 0; j < 5; j++) rf0[j] = res_f[j * sfine + i];
        const double vroot = want_min ? vol_root[i] : 0.0;
        double nv[5];
#pragma unroll
        for (int j = 0; j < 5; j++) {
            const double avg = acc[j] / wsum;
            nv[j] = vf0[j] + (rf0[j] - avg);
        }
        const Rec n = make_rec(nv[0], nv[1], nv[2], nv[3], nv[4]);
        store_rec(var_f, i, n);
        if (want_min) dt_new = 0.5 * (vroot / n.s);
        if (DIST) dist_push_rec(d, i, n);
    }
    if (blockmins) block_min_store(dt_new, blockmins);
    if (d.minword) block_min_atomic(dt_new, d.minword);
}

// ------------------------------------------------------------------------------------------------------
// alternative flux path and the bandwidth probe (node-ordering sweeps, BASELINE.json config 5)
// ------------------------------------------------------------------------------------------------------
// one thread per internal edge in ORIGINAL edge order, fp64 atomics (RED.ADD.F64): baseline + ordering-sweep kernel
__global__ void k_flux_atomic(long ne, const int* __restrict__ ea, const int* __restrict__ eb, const double* __restrict__ ew,
                              const double* __restrict__ recs, long stride, double* __restrict__ flux, double k2) {
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const int a = ea[e], b = eb[e];
    const Rec B = load_rec(recs, b);
    const Rec A = load_rec(recs, a);
    const Flux5 f = edge_flux(A, B, -0.5 * ew[e], -0.5 * ew[ne + e], -0.5 * ew[2 * ne + e], k2);
    atomicAdd(&flux[a], f.r); atomicAdd(&flux[stride + a], f.mx); atomicAdd(&flux[2 * stride + a], f.my);
    atomicAdd(&flux[3 * stride + a], f.mz); atomicAdd(&flux[4 * stride + a], f.e);
    atomicAdd(&flux[b], -f.r); atomicAdd(&flux[stride + b], -f.mx); atomicAdd(&flux[2 * stride + b], -f.my);
    atomicAdd(&flux[3 * stride + b], -f.mz); atomicAdd(&flux[4 * stride + b], -f.e);
}
// boundary + wall edges, one thread per edge (a node can carry several, hence atomics)
__global__ void k_bflux_atomic(long nb, const int* __restrict__ bnode, const uint8_t* __restrict__ bkind, const double* __restrict__ bw,
                               const double* __restrict__ recs, long stride, double* __restrict__ flux, int mask, const FarField F) {
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= nb) return;
    const int kind = bkind[e];
    if (!((mask >> kind) & 1)) return;
    const int b = bnode[e];
    const Rec B = load_rec(recs, b);
    Flux5 f = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (kind == 1) boundary_flux_acc(B, bw[e], bw[nb + e], bw[2 * nb + e], f);
    else wall_flux_acc(F, B, bw[e], bw[nb + e], bw[2 * nb + e], f);
    atomicAdd(&flux[b], f.r); atomicAdd(&flux[stride + b], f.mx); atomicAdd(&flux[2 * stride + b], f.my);
    atomicAdd(&flux[3 * stride + b], f.mz); atomicAdd(&flux[4 * stride + b], f.e);
}
// indirect_rw (indirect_rw_kernel.elemfunc.c): same gather/scatter as the flux kernel, no arithmetic to speak of
__global__ void k_indirect_rw(long ne, const int* __restrict__ ea, const int* __restrict__ eb, const double* __restrict__ ew,
                              const double* __restrict__ recs, long stride, double* __restrict__ flux) {
    const long e = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (e >= ne) return;
    const int a = ea[e], b = eb[e];
    const double ex = ew[e], ey = ew[ne + e], ez = ew[2 * ne + e];
    const Rec A = load_rec(recs, a), B = load_rec(recs, b);
    atomicAdd(&flux[a], B.rho + ex); atomicAdd(&flux[4 * stride + a], B.re + ey); atomicAdd(&flux[stride + a], B.mx + ez);
    atomicAdd(&flux[2 * stride + a], B.my); atomicAdd(&flux[3 * stride + a], B.mz);
    atomicAdd(&flux[b], A.rho); atomicAdd(&flux[4 * stride + b], A.re); atomicAdd(&flux[stride + b], A.mx);
    atomicAdd(&flux[2 * stride + b], A.my); atomicAdd(&flux[3 * stride + b], A.mz);
}

// ------------------------------------------------------------------------------------------------------
// distributed runs: halo exchange staging and the split RMS reduction
// ------------------------------------------------------------------------------------------------------
// send buffer of node records: dst row k = record idx[k] (one thread per 16-byte chunk)
__global__ void k_pack_records(const double* __restrict__ recs, const int* __restrict__ idx, long n, double* __restrict__ dst) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= 4 * n) return;
    const long k = i >> 2; const int c = int(i & 3);
    reinterpret_cast<double2*>(dst)[4 * k + c] = reinterpret_cast<const double2*>(recs)[4 * (long)idx[k] + c];
}
// residuals (SoA planes) of the listed nodes -> packed rows of 5
__global__ void k_pack_soa5(const double* __restrict__ soa, long stride, const int* __restrict__ idx, long n, double* __restrict__ dst) {
    const long k = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (k >= n) return;
    const long g = idx[k];
#pragma unroll
    for (int j = 0; j < 5; j++) dst[5 * k + j] = soa[j * stride + g];
}
// packed rows of 5 -> SoA planes at rows row0 .. row0+n (the ghost rows)
__global__ void k_unpack_soa5(double* __restrict__ soa, long stride, long row0, long n, const double* __restrict__ src) {
    const long k = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (k >= n) return;
#pragma unroll
    for (int j = 0; j < 5; j++) soa[j * stride + row0 + k] = src[5 * k + j];
}
// ---- direct peer-to-peer halo exchange over NVLink (CUDA IPC windows), one kernel per exchange ----------------------------
// Every rank owns a WINDOW other ranks can write: flags[64] (latest operation number completed by each source), reduction slots
// and, per level, two staging buffers for incoming halo rows.  An operation number g grows by one per collective operation on
// every rank alike.  k_p2p_exchange, on every rank at once:
//   1. PUT   : gathers the rows each peer needs and stores them straight into that peer's staging buffer (peer pointers);
//   2. SIGNAL: the last block to finish (ticket) fences system-wide and writes g into flags[me] of every peer;
//   3. WAIT  : every block spins (ld.acquire.sys) until flags[src] >= g for every source rank of this level;
//   4. UNPACK: copies its share of the staging buffer into the ghost rows.
// The grid is capped to what is resident at once (spinning blocks must not keep unscheduled ones from running).
// width = doubles per node in the message (8: records, 5: residuals).  src_is_soa: residual planes (stride) instead of record rows.
template <int WIDTH, bool SOA>
__global__ void k_p2p_exchange(const double* __restrict__ src, long stride, const int* __restrict__ send_idx, const P2PPeer* __restrict__ peers, int npeers,
                               unsigned long long* op_counter, unsigned int* level_counter, unsigned int* ticket,
                               const unsigned long long* my_flags, const double* __restrict__ my_stage0, long stage_parity_stride,
                               double* __restrict__ dst, long ghost_row0) {
    const long tid = blockIdx.x * (long)blockDim.x + threadIdx.x, nthreads = (long)gridDim.x * blockDim.x;
    // operation number and staging parity live on the device (bumped by the last block of each operation), so the kernel's
    // arguments never change and a whole V-cycle, exchanges included, can be replayed as one CUDA graph
    const unsigned long long g = *(volatile unsigned long long*)op_counter + 1;
    const int parity = int(*(volatile unsigned int*)level_counter & 1u);
    const double* my_stage = my_stage0 + (size_t)parity * stage_parity_stride;
    // 1. put
    for (int p = 0; p < npeers; p++) {
        const P2PPeer pe = peers[p];
        double* out = pe.dst[parity];
        for (long k = tid; k < pe.nsend * WIDTH; k += nthreads) {
            const long row = k / WIDTH; const int j = int(k - row * WIDTH);
            const long node = send_idx[pe.send0 + row];
            out[k] = SOA ? src[j * stride + node] : src[8 * node + j];
        }
    }
    // 2. signal
    __shared__ bool last;
    __syncthreads();                         // the block's stores are ordered before thread 0's system-scope fence (cumulativity,
    if (threadIdx.x == 0) {                  // the pattern of a cooperative-groups grid barrier): one fence per block, not per thread
        __threadfence_system();
        last = (atomicInc(ticket, gridDim.x - 1) == gridDim.x - 1);
    }
    __syncthreads();
    if (last && threadIdx.x < npeers) { __threadfence_system(); st_release_sys(peers[threadIdx.x].flag, g); }
    if (last && threadIdx.x == 0) { *op_counter = g; *level_counter += 1; }     // every block has read both (it passed the ticket)
    // 3. wait for every source of this level
    if (threadIdx.x < npeers && peers[threadIdx.x].nrecv >= 0) {
        const unsigned long long* f = my_flags + peers[threadIdx.x].rank;
        unsigned spins = 0;
        while (ld_acquire_sys(f) < g) { if (spin_expired(spins, "halo exchange")) break; }
    }
    __syncthreads();
    // 4. unpack staging -> ghost rows (a source's rows start at 8 * recv0 doubles whatever the message width)
    for (int p = 0; p < npeers; p++) {
        const P2PPeer pe = peers[p];
        const double* in = my_stage + 8 * pe.recv0;
        for (long k = tid; k < pe.nrecv * WIDTH; k += nthreads) {
            const long row = k / WIDTH; const int j = int(k - row * WIDTH);
            const double v = __ldcg(in + k);
            if (SOA) dst[j * stride + ghost_row0 + pe.recv0 + row] = v; else dst[8 * (ghost_row0 + pe.recv0 + row) + j] = v;
        }
    }
}
// all-reduce of n (<= 8) doubles (sum) or of one uint64 (min) over all ranks, one kernel on every rank: store my contribution into
// slot [parity][me] of every window (mine included), signal, wait for everybody, combine in rank order (deterministic).
__global__ void k_p2p_allreduce(double* __restrict__ vals, int n, int is_min, int nranks, int me, double* const* __restrict__ red_of_rank /*[nranks] window red bases*/,
                                unsigned long long* const* __restrict__ flag_of_rank /*[nranks] &window.flags[me]*/, const unsigned long long* my_flags,
                                const double* __restrict__ my_red, unsigned long long* op_counter, unsigned int* red_counter) {
    const int t = threadIdx.x;
    const unsigned long long g = *(volatile unsigned long long*)op_counter + 1;
    const int parity = int(*(volatile unsigned int*)red_counter & 1u);
    __syncthreads();
    if (t < nranks) {
        double* slot = red_of_rank[t] + ((size_t)parity * 64 + me) * 8;
        for (int j = 0; j < n; j++) slot[j] = vals[j];
        __threadfence_system();
        st_release_sys(flag_of_rank[t], g);
        unsigned spins = 0;
        while (ld_acquire_sys(my_flags + t) < g) { if (spin_expired(spins, "all-reduce kernel")) break; }
    }
    __syncthreads();
    if (t < n) {
        const double* base = my_red + (size_t)parity * 64 * 8;
        if (is_min) {
            unsigned long long m = ~0ull;
            for (int r = 0; r < nranks; r++) { const unsigned long long v = (unsigned long long)__double_as_longlong(__ldcg(base + r * 8 + t)); m = v < m ? v : m; }
            vals[t] = __longlong_as_double((long long)m);
        } else {
            double acc = 0.0;
            for (int r = 0; r < nranks; r++) acc += __ldcg(base + r * 8 + t);
            vals[t] = acc;
        }
    }
    if (t == 0) { *op_counter = g; *red_counter += 1; }
}

// calc_rms of a distributed run in ONE kernel (one block): local sums of the per-tile partials in index order, all-reduce(sum) of the
// five sums over the ranks (k_p2p_allreduce's protocol; the kernel owns one epoch like every other: base + epoch_off), square roots
// over the global node count
__global__ void k_rms_dist(const double* __restrict__ partial, long nparts, double nel_global, double* __restrict__ out, int* counter, int cap,
                           const AllRed ar, const unsigned long long* op_counter, int epoch_off) {
    __shared__ double ws[5][256];
    __shared__ double sums[8];
    const int t = threadIdx.x;
    pdl_wait();
    double s[5] = {0, 0, 0, 0, 0};
    for (long p = t; p < nparts; p += 256)
#pragma unroll
        for (int k = 0; k < 5; k++) s[k] += partial[p * 5 + k];
#pragma unroll
    for (int k = 0; k < 5; k++) ws[k][t] = s[k];
    __syncthreads();
    if (t < 5) {
        double acc = 0.0;
        for (int j = 0; j < 256; j++) acc += ws[t][j];
        sums[t] = acc;
    }
    __syncthreads();
    if (t < 32) {
        const unsigned long long g = *(volatile const unsigned long long*)op_counter + (unsigned long long)epoch_off;
        double v[8];
#pragma unroll
        for (int k = 0; k < 5; k++) v[k] = sums[k];
        dist_allreduce(ar, g, v, 5, false);
        if (t == 0) {
            int slot = 0;
            if (counter) { slot = *counter; *counter = slot + 1; if (slot >= cap) slot = cap - 1; }
            double tot = 0.0;
            for (int k = 0; k < 5; k++) { out[slot * 6 + 1 + k] = sqrt(v[k] / nel_global); tot += v[k]; }
            out[slot * 6] = sqrt(tot / nel_global);
        }
    }
}

// calc_rms split around an all-reduce: local sums of squares, then the roots over the global node count
__global__ void k_rms_sums(const double* __restrict__ partial, long nparts, double* __restrict__ sums) {
    __shared__ double ws[5][256];
    const int t = threadIdx.x;
    double s[5] = {0, 0, 0, 0, 0};
    for (long p = t; p < nparts; p += 256)
#pragma unroll
        for (int k = 0; k < 5; k++) s[k] += partial[p * 5 + k];
#pragma unroll
    for (int k = 0; k < 5; k++) ws[k][t] = s[k];
    __syncthreads();
    if (t < 5) {
        double acc = 0.0;
        for (int j = 0; j < 256; j++) acc += ws[t][j];
        sums[t] = acc;
    }
}
__global__ void k_rms_finish(const double* __restrict__ sums, double nel, double* __restrict__ out, int* counter, int cap) {
    if (threadIdx.x != 0) return;
    int slot = 0;
    if (counter) { slot = *counter; *counter = slot + 1; if (slot >= cap) slot = cap - 1; }
    double tot = 0.0;
    for (int k = 0; k < 5; k++) { out[slot * 6 + 1 + k] = sqrt(sums[k] / nel); tot += sums[k]; }
    out[slot * 6] = sqrt(tot / nel);
}

// ------------------------------------------------------------------------------------------------------
// layout conversion at the boundary: reference AoS (old order) <-> device (new order, padded)
// ------------------------------------------------------------------------------------------------------
__global__ void k_export_soa(const double* __restrict__ soa, long stride, int ncomp, long nel, const int* __restrict__ new_of_old, double* __restrict__ aos) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nel) return;
    const long g = new_of_old[i];
    for (int k = 0; k < ncomp; k++) aos[i * ncomp + k] = soa[k * stride + g];
}
__global__ void k_import_soa(double* __restrict__ soa, long stride, int ncomp, long nel, const int* __restrict__ new_of_old, const double* __restrict__ aos) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nel) return;
    const long g = new_of_old[i];
    for (int k = 0; k < ncomp; k++) soa[k * stride + g] = aos[i * ncomp + k];
}
__global__ void k_export_recs(const double* __restrict__ recs, long nel, const int* __restrict__ new_of_old, double* __restrict__ aos) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nel) return;
    const long g = new_of_old[i];
    for (int k = 0; k < 5; k++) aos[i * 5 + k] = recs[8 * g + k];
}
__global__ void k_import_recs(double* __restrict__ recs, long nel, const int* __restrict__ new_of_old, const double* __restrict__ aos) {
    const long i = blockIdx.x * (long)blockDim.x + threadIdx.x;
    if (i >= nel) return;
    const long g = new_of_old[i];
    store_rec(recs, g, make_rec(aos[i * 5], aos[i * 5 + 1], aos[i * 5 + 2], aos[i * 5 + 3], aos[i * 5 + 4]));
}

}  // namespace mgcfd
