// visit_kernel.cuh -- ONE persistent kernel per smoothing visit of a level (euler3d_cpu_double.cpp:383-512):
//   compute_step_factor's global minimum (cfd_loops.cpp:123-145) + the three Runge-Kutta stages, each
//   compute_flux_edge + compute_boundary_flux_edge + compute_wall_flux_edge + time_step (flux_loops.cpp, cfd_loops.cpp:215-280),
//   + residual, the RMS sums (validation.cpp:77-105) and the validity check on the last stage -- separated by GRID BARRIERS
//   instead of kernel boundaries.
//
// Why (profiles/r01g_final_c2_launches.txt): on L2-sized levels the stage-per-launch design was latency bound -- 31 dependent
// launches per V-cycle, every 128/256-node tile paying its own header -> records -> edges -> update chain, halo rows ~1-2x the
// owned rows.  Here
//   * one CTA per SM (512 threads = 16 autonomous WARPS), all CTAs co-resident, the grid barrier a 32-bit word in L2;
//   * a CTA owns K consecutive SUPER-TILES (runs of consecutive 128-node tiles, plan.h VisitPlan); the records of a whole
//     super-tile (own rows + its halo) are staged in shared memory together by cp.async, so the halo is that of a ~600-2000 node
//     block, not of a 128-node tile;
//   * the unit of work is a WARP-TILE (32 rows); warp w takes the warp-tiles w, w+16, ... of the super-tile and reads its edge
//     rounds from its OWN ring, filled by TMA bulk copies (cp.async.bulk + mbarrier) that its lane 0 issues from a flat chunk list
//     the host has precomputed per warp (VisitPlan::desc) the moment the warp has finished an entry: no warp ever waits for
//     another one inside a super-tile.  Measured on the way here (profiles/r02*_timeline_c2.jsonl, r02g_visit_*): ring entries handed
//     back through a named barrier of four warps cost 2.5 of 4.7 us per tile; a lane walking the tile / chunk structure itself to
//     find the next refill cost 31 % of the warp's time; a producer warp leaves 15 consumers for super-tiles of 16 warp-tiles
//     (a third of all stall samples at the end-of-iteration barrier).  The edge stream is static: refills run ahead across stages;
//   * the flux arithmetic is split (edge_acc / edge_acc_finish): everything that depends only on the node itself is summed once per
//     node from precomputed per-row geometry, |h| k2 comes with the slot: 21 FP64 instructions per edge and no square root;
//   * K == 1 and enough shared memory ("resident"): the own rows never leave the SM during the visit -- a stage writes the new
//     record into the other own-row buffer (and to global memory for the neighbours), only the halo rows are re-read after a barrier;
//   * the minimum dt: the transfer kernel that produced the level's state (restrict / prolong) leaves per-block minima of
//     0.5 cbrt(vol) / (|v| + c) behind, every CTA reduces them itself -- no pass over the nodes, no barrier.  Only when the state
//     came from elsewhere (set_field, the first cycle) the kernel reduces over its own nodes and the last CTA to arrive at
//     "barrier 0" combines the CTA minima; the flux rounds of stage 0 do not need the value and run while that barrier completes;
//   * multi-GPU (DIST): the records (and, on the last stage, residuals) of nodes other ranks hold as ghosts are stored straight
//     into those ranks' arrays over NVLink from the update; the last CTA to arrive at a barrier fences system-wide, signals the
//     peers and waits for their signal before it releases the local barrier -- the halo exchange IS the grid barrier.
#pragma once
#include "kernels.cuh"

namespace mgcfd {

constexpr int VW = 32;             // rows of a warp-tile
constexpr int VNW_MAX = 16;        // warps per CTA: 16 (one CTA per SM) or 8 (two CTAs per SM, whose phases interleave); 128 registers per thread either way
constexpr int VSLOT = 34;          // bytes per slot: hx, hy, hz, wk (doubles) + code (u16)
constexpr int VT = 128;            // rows of a tile of the level plan (four warp-tiles)
constexpr int VRING_MAX = 4;       // most ring entries per warp

struct DistArgs {
    // peers of this level (halo exchange): signal targets + whom to wait for
    const P2PPeer* peers; int npeers;
    const PeerOut* peer_out;
    const int* tgt_off; const int* tgt_peer; const int* tgt_row;       // node -> (peer index, row in the peer's arrays)
    const unsigned char* tile_sends;                                    // per tile: any node with a target
    // start-of-kernel wait (may be another level's peers) and the announcement targets (kernels.cuh "Epochs")
    const P2PPeer* wait_peers; int nwait;
    const P2PPeer* sig_peers; int nsig;
    AllRed ar;                                                          // all-reduce plumbing (kernels.cuh)
    const unsigned long long* my_flags;
    const unsigned long long* op_counter; int epoch_off;                // epoch of the kernel's first synchronisation = *op_counter + epoch_off (kernels.cuh)
};

struct VisitArgs {
    FarField ff;                                  // the context's far-field state (wall edges)
    double* bufX; double* bufA; double* bufB;     // records: state at visit start (= old_variables), stage 0/2 output, stage 1 output
    int ibX, ibA, ibB;                            // their indices in the level's buffer triple (DIST: which peer buffer mirrors them)
    double* res; double* sf; const double* vol; const double* vol_root;
    const double* hsum; long hs_stride;           // [3][hs_stride]: per-row sums of the edge vectors h (VisitPlan::hsum)
    long stride;
    int legacy;
    const unsigned char* desc; int desc_stride, max_ent, hpad;
    const unsigned char* vslots; const unsigned char* bslots;
    const int* cta_rows;                          // [grid + 1] first row of every CTA's run of super-tiles
    int K, sr_max, resident;                      // super-tiles per CTA, rows of the largest one, own rows resident
    int max_chunk;
    double k2;
    unsigned long long* bad_key; const int* old_of_new; unsigned long long stage_seq0;
    unsigned int* bar;                            // grid barrier word: count in the low 16 bits, generation above
    double* cta_min; unsigned long long* min_bits;
    const double* premin; int npremin;            // per-block minima of dt left by the transfer kernel that produced bufX, or nullptr
    int gmin_ready;                               // *min_bits already holds the global minimum (multi-GPU: all-reduced at the end of that transfer kernel)
    double* cta_rms;                              // [grid][5] (level 0) or nullptr
    double* rms_out; int* rms_counter; int rms_cap; double nel_global;
    long long* dbg;                               // DBG instantiation: 64 clock stamps per CTA (thread 0), see tools/visit_timeline.py
    DistArgs d;
};

__device__ __forceinline__ unsigned ld_acquire_gpu_u32(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// generations completed since the kernel read gen0
__device__ __forceinline__ unsigned bar_done(const unsigned* bar, unsigned gen0) { return ((ld_acquire_gpu_u32(bar) >> 16) - gen0) & 0xFFFFu; }
// spins until `n` generations have completed: relaxed polls (an acquire load per poll invalidates the L1 every time, CCTL.IVALL),
// one acquire once the generation is there
__device__ __forceinline__ void bar_wait(const unsigned* bar, unsigned gen0, unsigned n) {
    unsigned spins = 0;
    while ((((*(volatile const unsigned*)bar) >> 16) - gen0 & 0xFFFFu) < n) { if (spin_expired(spins, "grid barrier")) break; }
    (void)ld_acquire_gpu_u32(bar);
}

// a record from the visit kernel's buffers: code as VisitPlan documents it
__device__ __forceinline__ Rec sm_load_rec_code(const unsigned char* own, const unsigned char* halo, unsigned code) {
    const unsigned char* base = (code & 0x8000u) ? halo : own;
    return sm_load_rec_off(base, (code & 0x7fffu) << 4);
}
__device__ __forceinline__ void sm_store_rec_row(unsigned char* base, int row, const Rec& n) {
    double2* r = reinterpret_cast<double2*>(base + 64 * (size_t)row);
    const int x = (row >> 1) & 3;
    r[0 ^ x] = make_double2(n.rho, n.mx); r[1 ^ x] = make_double2(n.my, n.mz); r[2 ^ x] = make_double2(n.re, n.ir); r[3 ^ x] = make_double2(n.p, n.s);
}

// What one internal edge contributes to node A's flux (flux_kernel.elemfunc.c:130-162, with the algebra of edge_flux_acc_w in
// kernels.cuh), split into the part that depends on the OTHER end B -- accumulated per edge below -- and the part that only needs A
// and sums of per-edge geometry, added once per node by edge_acc_finish:
//   sum_e [fac_e (A.x - B.x) + ...]  =  A.x * sum_e fac_e  -  sum_e fac_e B.x  + ...      fac_e = wk_e (A.s + B.s), wk_e = |h_e| k2 (precomputed)
//   sum_e h_e . m_A                  =  (sum_e h_e) . m_A                                  (sum_e h_e precomputed per row: VisitPlan::hsum)
// 21 FP64 instructions per edge and no square root, against 49 for the unsplit form; only the order of the additions differs.
struct EdgeAcc { double F, R, E, MX, MY, MZ; };
__device__ __forceinline__ void edge_acc(double As, const Rec& B, double hx, double hy, double hz, double wk, EdgeAcc& a) {
    const double gB = __fma_rn(hz, B.mz, __fma_rn(hy, B.my, __dmul_rn(hx, B.mx)));
    const double qB = __dmul_rn(gB, B.ir);
    const double fac = __dmul_rn(wk, __dadd_rn(As, B.s));
    a.F = __dadd_rn(a.F, fac);
    a.R = __dadd_rn(__fma_rn(-fac, B.rho, a.R), gB);
    a.E = __fma_rn(__dadd_rn(B.re, B.p), qB, __fma_rn(-fac, B.re, a.E));
    a.MX = __fma_rn(B.p, hx, __fma_rn(B.mx, qB, __fma_rn(-fac, B.mx, a.MX)));
    a.MY = __fma_rn(B.p, hy, __fma_rn(B.my, qB, __fma_rn(-fac, B.my, a.MY)));
    a.MZ = __fma_rn(B.p, hz, __fma_rn(B.mz, qB, __fma_rn(-fac, B.mz, a.MZ)));
}
__device__ __forceinline__ Flux5 edge_acc_finish(const Rec& A, const EdgeAcc& a, double hsx, double hsy, double hsz) {
    const double gA = __fma_rn(hsz, A.mz, __fma_rn(hsy, A.my, __dmul_rn(hsx, A.mx)));
    const double qA = __dmul_rn(gA, A.ir);
    Flux5 f;
    f.r  = __dadd_rn(__fma_rn(A.rho, a.F, a.R), gA);
    f.e  = __fma_rn(__dadd_rn(A.re, A.p), qA, __fma_rn(A.re, a.F, a.E));
    f.mx = __fma_rn(A.p, hsx, __fma_rn(A.mx, qA, __fma_rn(A.mx, a.F, a.MX)));
    f.my = __fma_rn(A.p, hsy, __fma_rn(A.my, qA, __fma_rn(A.my, a.F, a.MY)));
    f.mz = __fma_rn(A.p, hsz, __fma_rn(A.mz, qA, __fma_rn(A.mz, a.F, a.MZ)));
    return f;
}
// Shared memory through 32-bit window addresses: one register per address, offsets as immediates -- the generic-pointer forms
// of these loads cost ~3 integer instructions each in the edge loop (64-bit adds, the generic -> shared conversion).
__device__ __forceinline__ double2 lds128(unsigned addr) {
    double2 v;
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v.x), "=d"(v.y) : "r"(addr));
    return v;
}
__device__ __forceinline__ double lds64(unsigned addr) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds16(unsigned addr) {
    unsigned short v;
    asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr));
    return v;
}
__device__ __forceinline__ unsigned lds32(unsigned addr) {
    unsigned v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
    return v;
}
// record at row offset `off` (chunk 0 of the row under the 64B swizzle) of a 128-byte aligned buffer
__device__ __forceinline__ Rec lds_rec(unsigned off) {
    const double2 c0 = lds128(off), c1 = lds128(off ^ 16u), c2 = lds128(off ^ 32u), c3 = lds128(off ^ 48u);
    Rec n; n.rho = c0.x; n.mx = c0.y; n.my = c1.x; n.mz = c1.y; n.re = c2.x; n.ir = c2.y; n.p = c3.x; n.s = c3.y;
    return n;
}
// one ring chunk = exactly RC rounds of one warp-tile (the host pads with empty slots): fully unrolled, every load of the chunk
// can be in flight before the first flux instruction.  slot = shared address of hx[lane] of the chunk's first round.
template <int RC>
__device__ __forceinline__ void visit_chunk(unsigned slot, unsigned code_addr, unsigned own, unsigned halo, double As, EdgeAcc& acc) {
#pragma unroll
    for (int r = 0; r < RC; r++) {
        const unsigned s0 = slot + r * (VW * VSLOT);
        const double hx = lds64(s0), hy = lds64(s0 + VW * 8), hz = lds64(s0 + 2 * VW * 8), wk = lds64(s0 + 3 * VW * 8);
        const unsigned code = lds16(code_addr + r * (VW * VSLOT));
        const Rec B = lds_rec(((code & 0x8000u) ? halo : own) + ((code & 0x7fffu) << 4));
        edge_acc(As, B, hx, hy, hz, wk, acc);
    }
}

struct VEnt { int orow0, rounds, brounds, blane0; long long vblk0, bblk0; };      // 32 bytes, VisitPlan::desc
static_assert(sizeof(VEnt) == 32, "warp-tile entry layout (plan.cpp build_visit_streams)");

// Cross-rank part of a barrier, run by warp 0 of the last CTA to arrive (all 32 lanes): every remote store of this rank's CTAs
// is ordered before it (each CTA fenced system-wide before it arrived).  epoch = the number this synchronisation carries.
__device__ __forceinline__ void dist_signal_wait_peers(const DistArgs& d, unsigned long long epoch, bool wait) {
    const int lane = threadIdx.x & 31;
    __threadfence_system();
    for (int p = lane; p < d.npeers; p += 32) st_release_sys(d.peers[p].flag, epoch);
    if (wait)
        for (int p = lane; p < d.npeers; p += 32) {
            const unsigned long long* f = d.my_flags + d.peers[p].rank;
            unsigned spins = 0;
            while (ld_acquire_sys(f) < epoch) { if (spin_expired(spins, "visit kernel: stage barrier")) break; }
        }
    __syncwarp();
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
template <bool DIST, bool DBG, int RC, int VNW>
__global__ void __launch_bounds__(VNW * 32, 16 / VNW)
k_visit(const VisitArgs a) {
    constexpr int VNC = VNW * VW, VNT = VNC;
    auto consumer_sync = []() { __syncthreads(); };
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ __align__(8) unsigned long long bar_full[VNW][VRING_MAX], bar_recs[2];
    __shared__ double red[VNW][5];
    const int t = threadIdx.x, w = t >> 5, ln = t & 31;
    const int c = blockIdx.x, G = gridDim.x;
    const int K = a.K, Q = 3 * K;
    constexpr int D = 2;                                     // ring entries per warp
    constexpr unsigned ring_bytes = unsigned(RC) * unsigned(VW * VSLOT);
    unsigned char* recs = smraw + (((size_t)VNW * D * ring_bytes + 127) & ~size_t(127));
    const size_t own_bytes = 64 * (size_t)a.sr_max, halo_bytes = 64 * (size_t)a.hpad;
    unsigned char* descs = recs + (a.resident ? 2 * own_bytes + halo_bytes : 2 * (own_bytes + halo_bytes));
    auto own_base = [&](int q) -> unsigned char* { return recs + (q & 1) * (a.resident ? own_bytes : own_bytes + halo_bytes); };
    auto halo_base = [&](int q) -> unsigned char* { return a.resident ? recs + 2 * own_bytes : own_base(q) + own_bytes; };
    // K == 1: every iteration works on the same super-tile, one descriptor buffer; else four (q + 2 overwrites q - 2)
    auto desc_of = [&](int q) -> const unsigned char* { return descs + (K == 1 ? 0 : (q & 3)) * (size_t)a.desc_stride; };
    auto vin_of = [&](int q) -> const double* { const int j = q / K; return j == 0 ? a.bufX : (j == 1 ? a.bufA : a.bufB); };

    long long dbg_ring = 0, dbg_edge = 0, dbg_upd = 0, dbg_t = 0, dbg_pro = 0, dbg_t2 = 0;
    auto stamp = [&](int slot) { if (DBG && t == 0 && slot < 56) a.dbg[(size_t)c * 64 + slot] = clock64(); };
    stamp(0);
    if (ln == 0) for (int e = 0; e < 2; e++) mbar_init(&bar_full[w][e], 1);
    if (t == 0) { mbar_init(&bar_recs[0], VNC); mbar_init(&bar_recs[1], VNC); }
    __syncthreads();

    // descriptors of the first two iterations: static data, fetched (like the first edge chunks) while the predecessor drains
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    for (int q = 0; q < 2 && q < Q && (K > 1 || q == 0); q++) {
        const unsigned char* src = a.desc + ((size_t)c * K + (q % K)) * (size_t)a.desc_stride;
        unsigned char* dst = descs + (K == 1 ? 0 : (q & 3)) * (size_t)a.desc_stride;
        for (int b = t * 16; b < a.desc_stride; b += VNT * 16) cp_async16(dst + b, src + b);
    }
    cp_async_wait_all();
    __syncthreads();
    unsigned char* ring = smraw + (size_t)w * D * ring_bytes;                         // this warp's ring
    // Ring refills (lane 0): chunk n of the warp goes into entry n & 1; once the warp has finished chunk n the entry takes chunk
    // n + 2.  The chunks of an iteration are listed per warp in its descriptor (block indices; every chunk is RC rounds), so the
    // next refill is one 4-byte shared load away.
    int p_q = -1, issued = 0, consumed = 0;
    unsigned cl_ptr = 0, cl_end = 0;                          // shared addresses: next / end of this warp's chunk list of iteration p_q
    const unsigned ring32 = smem_u32(ring), full32 = smem_u32(&bar_full[w][0]);
    unsigned long long l2_policy;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(l2_policy));
    const int clist_off = 32 + 32 * a.max_ent;
    auto produce = [&](int q_visible) {
        while (issued - consumed < D) {
            if (cl_ptr == cl_end) {
                if (p_q >= q_visible) return;
                p_q++;
                const unsigned dd = smem_u32(desc_of(p_q)) + clist_off;
                const unsigned c01 = lds32(dd + 2 * (w & ~1));           // coff[w & ~1], coff[(w & ~1) + 1]
                const unsigned c2 = lds16(dd + 2 * (w + 1));
                const unsigned c0 = (w & 1) ? (c01 >> 16) : (c01 & 0xFFFFu);
                cl_ptr = dd + 48 + 4 * c0; cl_end = dd + 48 + 4 * c2;
                continue;
            }
            const unsigned vblk = lds32(cl_ptr);
            const unsigned e = issued & 1u;
            const unsigned bar = full32 + 8 * e;
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "n"(ring_bytes) : "memory");
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                         ::"r"(ring32 + e * ring_bytes), "l"(a.vslots + (size_t)vblk * (VW * VSLOT)), "n"(ring_bytes), "r"(bar), "l"(l2_policy) : "memory");
            issued++; cl_ptr += 4;
        }
    };
    auto copy_desc = [&](int q) {
        if (q >= Q || K == 1) return;
        const unsigned char* src = a.desc + ((size_t)c * K + (q % K)) * (size_t)a.desc_stride;
        unsigned char* dst = descs + (q & 3) * (size_t)a.desc_stride;      // q + 2 overwrites q - 2, which nobody can still be reading
        for (int b = t * 16; b < a.desc_stride; b += VNC * 16) cp_async16(dst + b, src + b);
    };
    // records of iteration q: the own rows of its super-tile (skipped when they are already resident) and its halo rows
    auto copy_recs = [&](int q, bool own_too) {
        if (q >= Q) return;
        const unsigned char* d = desc_of(q);
        const int* di = reinterpret_cast<const int*>(d);
        const long row0 = di[0]; const int nrows = di[1] * VT, nhalo = di[2];
        const int* ids = reinterpret_cast<const int*>(d + 32 + 32 * a.max_ent + 48 + 4 * a.max_chunk);
        const double* vin = vin_of(q);
        const int k16 = (t & 3) << 4;
        if (own_too) {
            unsigned char* ob = own_base(q);
            const unsigned char* src = reinterpret_cast<const unsigned char*>(vin + 8 * row0);
            for (int i = t; i < 4 * nrows; i += VNC) {
                const int row = i >> 2;
                cp_async16(ob + 64 * row + (k16 ^ (((row >> 1) & 3) << 4)), src + 64 * (size_t)row + k16);
            }
        }
        unsigned char* hb = halo_base(q);
        for (int i = t; i < 4 * nhalo; i += VNC) {
            const int h = i >> 2;
            cp_async16(hb + 64 * h + (k16 ^ (((h >> 1) & 3) << 4)), reinterpret_cast<const unsigned char*>(vin + 8 * (long)ids[h]) + k16);
        }
        cp_async_mbar_arrive(&bar_recs[q & 1]);
    };

    if (ln == 0) produce(Q > 1 && K > 1 ? 1 : 0);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned long long E0 = 0;
    if (DIST) {
        E0 = *(volatile const unsigned long long*)a.d.op_counter + (unsigned long long)a.d.epoch_off;
        if (c == 0 && t < a.d.nsig) st_release_sys(a.d.sig_peers[t].flag, E0);      // this rank's kernels up to E0 - 1 are complete
        if (t < a.d.nwait) {
            const unsigned long long* f = a.d.my_flags + a.d.wait_peers[t].rank;
            unsigned spins = 0;
            while (ld_acquire_sys(f) < E0) { if (spin_expired(spins, "visit kernel start: peer epoch")) break; }
        }
        consumer_sync();
    }
    stamp(1);
    const unsigned gen0 = ld_acquire_gpu_u32(a.bar) >> 16;
    unsigned nbar = 0;                   // grid barriers this CTA has arrived at
    unsigned long long nsync = 0;        // cross-rank synchronisations so far (DIST)
    copy_recs(0, true);

    // ---- the visit's global minimum dt (cfd_loops.cpp:123-145) ----
    const double BIG = __longlong_as_double(0x7F7F7F7F7F7F7F7FLL);
    const bool need_min = !a.legacy;
    double min_dt = 0.0;
    bool have_min = !need_min;
    if (need_min && a.gmin_ready) { min_dt = __longlong_as_double((long long)__ldcg(a.min_bits)); have_min = true; }
    else if (need_min || DIST) {
        double val = BIG;
        if (need_min) {
            if (a.premin) {           // per-block minima left by the restrict / prolong kernel that produced bufX: every CTA reduces them itself
                for (int b = t; b < a.npremin; b += VNC) val = fmin(val, __ldcg(a.premin + b));
            } else {                  // the state came from elsewhere: reduce over this CTA's own nodes
                const long r0 = a.cta_rows[c], r1 = a.cta_rows[c + 1];
                for (long r = r0 + t; r < r1; r += VNC) val = fmin(val, 0.5 * (a.vol_root[r] / a.bufX[8 * r + 7]));
            }
#pragma unroll
            for (int dlt = 16; dlt > 0; dlt >>= 1) val = fmin(val, __shfl_xor_sync(0xffffffffu, val, dlt));
            if (ln == 0) red[w][0] = val;
        }
        consumer_sync();
        if (need_min) {
            val = (ln < VNW) ? red[ln][0] : BIG;
#pragma unroll
            for (int dlt = 16; dlt > 0; dlt >>= 1) val = fmin(val, __shfl_xor_sync(0xffffffffu, val, dlt));      // every thread: this CTA's minimum
        }
        if (need_min && a.premin && !DIST) { min_dt = val; have_min = true; }
        else {
            // barrier 0: the CTA minima (or, DIST, the rank minima) are combined by the last CTA to arrive; the stage-0 flux rounds run meanwhile
            if (w == 0) {
                if (need_min && t == 0) a.cta_min[c] = val;
                unsigned old = 0;
                if (t == 0) { __threadfence(); old = atomicAdd(a.bar, 1u); }
                old = __shfl_sync(0xffffffffu, old, 0);
                if ((old & 0xFFFFu) == unsigned(G - 1)) {
                    __threadfence();
                    double m = BIG;
                    if (need_min) {
                        if (a.premin) m = val;
                        else {
                            for (int b = t; b < G; b += 32) m = fmin(m, __ldcg(a.cta_min + b));
#pragma unroll
                            for (int dlt = 16; dlt > 0; dlt >>= 1) m = fmin(m, __shfl_xor_sync(0xffffffffu, m, dlt));
                        }
                    }
                    if (DIST) dist_allreduce(a.d.ar, E0 + 1, &m, 1, true);
                    if (t == 0) {
                        *a.min_bits = (unsigned long long)__double_as_longlong(m);
                        __threadfence();
                        atomicAdd(a.bar, 0x10000u - unsigned(G));
                    }
                }
            }
            nbar = 1; nsync = 1;
        }
        consumer_sync();                 // red[][] is reused by the RMS sums
    }
    stamp(2);
    double rms_acc[5] = {0.0, 0.0, 0.0, 0.0, 0.0};      // per warp (lane 0): r^2 sums over the warp's nodes of the last stage

    for (int q = 0; q < Q; q++) {
        const int j = q / K, k = q - j * K;
        const bool first_stage = (j == 0), last_stage = (j == MGCFD_RK - 1);
        copy_desc(q + 2);
        stamp(3 + 4 * q);
        mbar_wait(&bar_recs[q & 1], (q >> 1) & 1);          // records of iteration q (and the descriptor of q + 1) have landed
        stamp(4 + 4 * q);
        if (k + 1 < K) copy_recs(q + 1, true);              // same stage: the other buffer is free (barrier at the end of q - 1)
        const int q_vis = (K == 1) ? Q - 1 : min(q + 1, Q - 1);    // K == 1: one descriptor serves every iteration
        if (ln == 0) produce(q_vis);
        __syncwarp();
        const unsigned char* d = desc_of(q);
        const int* di = reinterpret_cast<const int*>(d);
        const long row0 = di[0]; const int nent = di[4];
        const unsigned own32 = smem_u32(own_base(q)), halo32 = smem_u32(halo_base(q));
        const double* vold = a.bufX;
        double* vout = (j == 1) ? a.bufB : a.bufA;
        const double rk_div = double(MGCFD_RK + 1 - j), rk_rcp = 1.0 / rk_div;

        for (int ei = w; ei < nent; ei += VNW) {
            if (DBG) dbg_t2 = clock64();
            const VEnt* en = reinterpret_cast<const VEnt*>(d + 32) + ei;
            const int orow = en->orow0 + ln;                // this thread's row inside the super-tile
            const long gid = row0 + orow;
            // early loads for the update: in flight while the edge rounds run
            const double vol_or_sf = first_stage ? a.vol[gid] : a.sf[gid];
            const double hsx = a.hsum[gid], hsy = a.hsum[a.hs_stride + gid], hsz = a.hsum[2 * a.hs_stride + gid];
            const int brounds = en->brounds;
            const int bl = en->blane0 + ln;
            const unsigned char* bblk = a.bslots + en->bblk0 * (long)(VT * 25);
            BSlot b0 = {0, 0.0, 0.0, 0.0};
            if (brounds > 0) b0 = bslot_fetch<VT>(bblk, bl);
            const Rec me = lds_rec(own32 + ((unsigned(orow) << 6) | (((unsigned(orow) >> 1) & 3u) << 4)));
            double o[5];
            if (first_stage) { o[0] = me.rho; o[1] = me.mx; o[2] = me.my; o[3] = me.mz; o[4] = me.re; }
            else {
                const double2* p = reinterpret_cast<const double2*>(vold + 8 * gid);
                const double2 c0 = p[0], c1 = p[1];
                o[0] = c0.x; o[1] = c0.y; o[2] = c1.x; o[3] = c1.y; o[4] = vold[8 * gid + 4];
            }
            EdgeAcc acc = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
            const int nchunks = en->rounds / RC;             // whole chunks (VisitPlan pads the rounds)
            if (DBG) dbg_pro += clock64() - dbg_t2;
            for (int ch = 0; ch < nchunks; ch++) {
                const unsigned e = unsigned(consumed) & 1u;
                if (DBG) dbg_t = clock64();
                mbar_wait(&bar_full[w][e], (consumed >> 1) & 1);
                if (DBG) { const long long now = clock64(); dbg_ring += now - dbg_t; dbg_t = now; }
                const unsigned ent32 = ring32 + e * ring_bytes;
                visit_chunk<RC>(ent32 + 8 * ln, ent32 + VW * 32 + 2 * ln, own32, halo32, me.s, acc);
                if (DBG) { const long long now = clock64(); dbg_edge += now - dbg_t; }
                __syncwarp();                               // every lane is done with the entry: lane 0 refills it
                consumed++;
                if (ln == 0) produce(q_vis);
            }
            if (DBG) dbg_t = clock64();
            Flux5 f = edge_acc_finish(me, acc, hsx, hsy, hsz);
            boundary_rounds<VT>(a.ff, bblk, brounds, bl, 7, me, f, b0);
            double sfv = vol_or_sf;
            if (first_stage) {
                if (!have_min) {          // first update of the visit: barrier 0 must have completed
                    if (ln == 0) bar_wait(a.bar, gen0, 1u);
                    __syncwarp();
                    min_dt = __longlong_as_double((long long)__ldcg(a.min_bits));
                    have_min = true;
                }
                sfv = a.legacy ? double(0.5) / (sqrt(vol_or_sf) * me.s) : min_dt / vol_or_sf;      // cfd_loops.cpp:60 / :146-156
                a.sf[gid] = sfv;
            }
            // time_step (cfd_loops.cpp:215-280) + the new record
            const double factor = div_rk(sfv, rk_div, rk_rcp);
            const double n0 = o[0] + factor * f.r, n1 = o[1] + factor * f.mx, n2 = o[2] + factor * f.my, n3 = o[3] + factor * f.mz, n4 = o[4] + factor * f.e;
            const Rec nrec = make_rec(n0, n1, n2, n3, n4);
            store_rec(vout, gid, nrec);
            if (a.resident && !last_stage) sm_store_rec_row(own_base(q + 1), orow, nrec);
            const bool sends = DIST && a.d.tile_sends[di[3] + (en->orow0 >> 7)] != 0;
            if (DIST) {
                if (sends) {
                    const int ib = (j == 1) ? a.ibB : a.ibA;
                    for (int x = a.d.tgt_off[gid]; x < a.d.tgt_off[gid + 1]; x++) store_rec(a.d.peer_out[a.d.tgt_peer[x]].rec[ib], a.d.tgt_row[x], nrec);
                }
            }
            {   // check_for_invalid_variables (validation.cpp:107-138): first offending cell of the first offending stage
                int reason = 0;
                if (!(isfinite(n0) && isfinite(n1) && isfinite(n2) && isfinite(n3) && isfinite(n4))) reason = 1;
                else if (n0 < 0.0) reason = 2;
                else if (n4 < 0.0) reason = 3;
                if (reason) {
                    const int oi = a.old_of_new[gid];
                    if (oi >= 0) atomicMin(a.bad_key, (((a.stage_seq0 + j) & 0xFFFFFFull) << 40) | ((unsigned long long)oi << 2) | (unsigned long long)reason);
                }
            }
            if (last_stage) {
                const long S = a.stride;
                const double r0 = n0 - o[0], r1 = n1 - o[1], r2 = n2 - o[2], r3 = n3 - o[3], r4 = n4 - o[4];     // residual(), validation.cpp:77-89
                a.res[gid] = r0; a.res[S + gid] = r1; a.res[2 * S + gid] = r2; a.res[3 * S + gid] = r3; a.res[4 * S + gid] = r4;
                if (DIST) {
                    if (sends)
                        for (int x = a.d.tgt_off[gid]; x < a.d.tgt_off[gid + 1]; x++) {
                            const PeerOut& po = a.d.peer_out[a.d.tgt_peer[x]];
                            double* pr = po.res + a.d.tgt_row[x];
                            pr[0] = r0; pr[po.res_stride] = r1; pr[2 * po.res_stride] = r2; pr[3 * po.res_stride] = r3; pr[4 * po.res_stride] = r4;
                        }
                }
                if (a.cta_rms) {
                    double qq[5] = {r0 * r0, r1 * r1, r2 * r2, r3 * r3, r4 * r4};
#pragma unroll
                    for (int v = 0; v < 5; v++) {
#pragma unroll
                        for (int dlt = 16; dlt > 0; dlt >>= 1) qq[v] += __shfl_down_sync(0xffffffffu, qq[v], dlt);
                        rms_acc[v] += qq[v];
                    }
                }
            }
            if (DBG) dbg_upd += clock64() - dbg_t;
        }
        stamp(5 + 4 * q);
        consumer_sync();        // every warp is done with this iteration's record buffers
        stamp(6 + 4 * q);
        if (k == K - 1 && !last_stage) {
            // ---- stage barrier: all CTAs (all ranks' neighbours) have written the new state ----
            if (w == 0) {
                unsigned old = 0;
                // (the counter is shared by all barriers of the launch: never arrive at one before the previous one has completed --
                // only the legacy multi-GPU case can get here without having waited for barrier 0)
                if (t == 0) { bar_wait(a.bar, gen0, nbar); if (DIST) __threadfence_system(); else __threadfence(); old = atomicAdd(a.bar, 1u); }
                old = __shfl_sync(0xffffffffu, old, 0);
                if ((old & 0xFFFFu) == unsigned(G - 1)) {
                    if (DIST) dist_signal_wait_peers(a.d, E0 + nsync + 1, true);
                    if (t == 0) { __threadfence(); atomicAdd(a.bar, 0x10000u - unsigned(G)); }
                }
                if (t == 0) bar_wait(a.bar, gen0, nbar + 1u);
            }
            nbar++; nsync++;
            consumer_sync();
            copy_recs(q + 1, !a.resident);
        }
    }

    if (DBG && t == 0) {
        long long* o = a.dbg + (size_t)c * 64;
        o[56] = dbg_ring; o[57] = dbg_edge; o[58] = dbg_upd; o[59] = clock64(); o[62] = 0; o[63] = dbg_pro;
        unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt)); o[60] = (long long)gt;
        unsigned smid; asm volatile("mov.u32 %0, %%smid;" : "=r"(smid)); o[61] = smid;
    }
    // ---- epilogue: RMS sums of level 0 (calc_rms, validation.cpp:91-105) and, DIST, the end-of-kernel signal ----
    if (a.cta_rms || DIST) {
        if (a.cta_rms) {
            if (ln == 0) {
#pragma unroll
                for (int v = 0; v < 5; v++) red[w][v] = rms_acc[v];
            }
            consumer_sync();
            if (t < 5) {
                double s = 0.0;
                for (int ww = 0; ww < VNW; ww++) s += red[ww][t];
                a.cta_rms[c * 5 + t] = s;
            }
            consumer_sync();
        }
        if (w == 0) {
            unsigned old = 0;
            if (t == 0) { bar_wait(a.bar, gen0, nbar); if (DIST) __threadfence_system(); else __threadfence(); old = atomicAdd(a.bar, 1u); }
            old = __shfl_sync(0xffffffffu, old, 0);
            if ((old & 0xFFFFu) == unsigned(G - 1)) {
                __threadfence();
                if (a.cta_rms) {
                    // fixed order: CTA 0, 1, ... per variable (deterministic whichever CTA arrives last)
                    double sums = 0.0;
                    if (t < 5) { for (int b = 0; b < G; b++) sums += __ldcg(a.cta_rms + b * 5 + t); }
                    double v5[8];
#pragma unroll
                    for (int v = 0; v < 5; v++) v5[v] = __shfl_sync(0xffffffffu, sums, v);
                    if (DIST) dist_allreduce(a.d.ar, E0 + nsync + 1, v5, 5, false);       // also this kernel's end signal to every rank
                    if (t == 0) {
                        int slot = 0;
                        if (a.rms_counter) { slot = *a.rms_counter; *a.rms_counter = slot + 1; if (slot >= a.rms_cap) slot = a.rms_cap - 1; }
                        double tot = 0.0;
                        for (int v = 0; v < 5; v++) { a.rms_out[slot * 6 + 1 + v] = sqrt(v5[v] / a.nel_global); tot += v5[v]; }
                        a.rms_out[slot * 6] = sqrt(tot / a.nel_global);
                    }
                } else if (DIST) {
                    dist_signal_wait_peers(a.d, E0 + nsync + 1, false);
                }
                if (t == 0) {
                    __threadfence();
                    atomicAdd(a.bar, 0x10000u - unsigned(G));
                }
            }
        }
    }
}

}  // namespace mgcfd
