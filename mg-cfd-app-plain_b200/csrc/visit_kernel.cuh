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
//   * the unit of work is a WARP-TILE (32 rows); warp w takes the warp-tiles w, w+16, ... of the super-tile and feeds ITSELF: its
//     own ring of edge-round blocks, filled by TMA bulk copies (cp.async.bulk + mbarrier) that its lane 0 issues as soon as the
//     warp has finished an entry -- no warp ever waits for another one inside a super-tile (the first version handed ring entries
//     back through a named barrier of four warps: 2.5 of 4.7 us per tile went into that hand-over, profiles/r02b_timeline_c2.jsonl);
//     the edge stream is static, so the producers run ahead across stage boundaries;
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
constexpr int VNW = 16;            // warps per CTA
constexpr int VNT = VW * VNW;      // threads per CTA
constexpr int VT = 128;            // rows of a tile of the level plan (four warp-tiles)
constexpr int VRING_MAX = 4;       // most ring entries per warp

struct DistArgs {
    int nranks, me;
    // peers of this level (halo exchange): signal targets + whom to wait for
    const P2PPeer* peers; int npeers;
    const PeerOut* peer_out;
    const int* tgt_off; const int* tgt_peer; const int* tgt_row;       // node -> (peer index, row in the peer's arrays)
    const unsigned char* tile_sends;                                    // per tile: any node with a target
    // start-of-kernel wait (may be another level's peers)
    const P2PPeer* wait_peers; int nwait;
    // all-reduce plumbing (as k_p2p_allreduce)
    double* const* red_of_rank; unsigned long long* const* flag_of_rank;
    const unsigned long long* my_flags; const double* my_red;
    unsigned long long* op_counter; unsigned int* red_counter;
};

struct VisitArgs {
    double* bufX; double* bufA; double* bufB;     // records: state at visit start (= old_variables), stage 0/2 output, stage 1 output
    int ibX, ibA, ibB;                            // their indices in the level's buffer triple (DIST: which peer buffer mirrors them)
    double* res; double* sf; const double* vol; const double* vol_root;
    long stride;
    int legacy;
    const unsigned char* desc; int desc_stride, max_ent, hpad;
    const unsigned char* vslots; const unsigned char* bslots;
    const int* cta_rows;                          // [grid + 1] first row of every CTA's run of super-tiles
    int K, sr_max, resident, R, D;                // super-tiles per CTA, rows of the largest one, own rows resident, rounds per ring entry, ring entries
    double k2;
    unsigned long long* bad_key; const int* old_of_new; unsigned long long stage_seq0;
    unsigned int* bar;                            // grid barrier word: count in the low 16 bits, generation above
    double* cta_min; unsigned long long* min_bits;
    const double* premin; int npremin;            // per-block minima of dt left by the transfer kernel that produced bufX, or nullptr
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

// edge rounds of one ring entry, sorted-segment form (edge_rounds<.., false> of kernels.cuh with the visit kernel's addressing and
// 32-lane blocks): two register sets used alternately, the slot / record / edge weight of round r+1 in flight during round r
__device__ __forceinline__ void visit_edge_rounds(const unsigned char* blk, int nr, const unsigned char* own, const unsigned char* halo, int t,
                                                  const Rec& me, double me_ep, double k2, Flux5& f) {
    if (nr <= 0) return;
    const double* w = reinterpret_cast<const double*>(blk);
    double h0x = w[t], h0y = w[VW + t], h0z = w[2 * VW + t];
    Rec B0 = sm_load_rec_code(own, halo, reinterpret_cast<const unsigned short*>(blk + VW * 24)[t]);
    double e0 = edge_weight(h0x, h0y, h0z);
    int r = 1;
    for (; r + 1 < nr; r += 2) {
        blk += VW * 26;
        const double* w1 = reinterpret_cast<const double*>(blk);
        const double h1x = w1[t], h1y = w1[VW + t], h1z = w1[2 * VW + t];
        const Rec B1 = sm_load_rec_code(own, halo, reinterpret_cast<const unsigned short*>(blk + VW * 24)[t]);
        const double e1 = edge_weight(h1x, h1y, h1z);
        edge_flux_acc_w(me, me_ep, B0, h0x, h0y, h0z, e0, k2, f);
        blk += VW * 26;
        const double* w2 = reinterpret_cast<const double*>(blk);
        h0x = w2[t]; h0y = w2[VW + t]; h0z = w2[2 * VW + t];
        B0 = sm_load_rec_code(own, halo, reinterpret_cast<const unsigned short*>(blk + VW * 24)[t]);
        e0 = edge_weight(h0x, h0y, h0z);
        edge_flux_acc_w(me, me_ep, B1, h1x, h1y, h1z, e1, k2, f);
    }
    if (r < nr) {
        blk += VW * 26;
        const double* w1 = reinterpret_cast<const double*>(blk);
        const double h1x = w1[t], h1y = w1[VW + t], h1z = w1[2 * VW + t];
        const Rec B1 = sm_load_rec_code(own, halo, reinterpret_cast<const unsigned short*>(blk + VW * 24)[t]);
        const double e1 = edge_weight(h1x, h1y, h1z);
        edge_flux_acc_w(me, me_ep, B0, h0x, h0y, h0z, e0, k2, f);
        edge_flux_acc_w(me, me_ep, B1, h1x, h1y, h1z, e1, k2, f);
    } else {
        edge_flux_acc_w(me, me_ep, B0, h0x, h0y, h0z, e0, k2, f);
    }
}

struct VEnt { int orow0, rounds, brounds, blane0; long long vblk0, bblk0, pad; };      // 32 bytes, VisitPlan::desc

// Cross-rank part of a barrier, run by warp 0 of the last CTA to arrive (all 32 lanes): every remote store of this rank's CTAs
// is ordered before it (each CTA fenced system-wide before it arrived).  epoch = the number this synchronisation carries.
__device__ __forceinline__ void dist_signal_wait_peers(const DistArgs& d, unsigned long long epoch, bool wait) {
    const int lane = threadIdx.x & 31;
    __threadfence_system();
    for (int p = lane; p < d.npeers; p += 32) st_release_sys(d.peers[p].flag, epoch);
    if (wait)
        for (int p = lane; p < d.npeers; p += 32) {
            const unsigned long long* f = d.my_flags + d.peers[p].rank;
            while (ld_acquire_sys(f) < epoch) { __nanosleep(20); }
        }
    __syncwarp();
}
// all-reduce over ALL ranks of n <= 8 doubles held by lane 0 in v[] (is_min: one bit pattern), deterministic rank order
__device__ __forceinline__ void dist_allreduce(const DistArgs& d, unsigned long long epoch, double* v, int n, bool is_min) {
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
        while (ld_acquire_sys(d.my_flags + p) < epoch) { __nanosleep(20); }
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

template <bool DIST, bool DBG = false>
__global__ void __launch_bounds__(VNT, 1)
k_visit(const VisitArgs a) {
    extern __shared__ __align__(128) unsigned char smraw[];
    __shared__ __align__(8) unsigned long long bar_ring[VNW][VRING_MAX], bar_recs[2];
    __shared__ double red[VNW][5];
    const int t = threadIdx.x, w = t >> 5, ln = t & 31;
    const int c = blockIdx.x, G = gridDim.x;
    const int K = a.K, Q = 3 * K, R = a.R, D = a.D;
    const unsigned ring_bytes = unsigned(R) * unsigned(VW * 26);
    unsigned char* ring = smraw + (size_t)w * D * ring_bytes;                         // this warp's ring
    unsigned char* recs = smraw + (((size_t)VNW * D * ring_bytes + 127) & ~size_t(127));
    const size_t own_bytes = 64 * (size_t)a.sr_max, halo_bytes = 64 * (size_t)a.hpad;
    unsigned char* descs = recs + (a.resident ? 2 * own_bytes + halo_bytes : 2 * (own_bytes + halo_bytes));
    auto own_base = [&](int q) -> unsigned char* { return recs + (q & 1) * (a.resident ? own_bytes : own_bytes + halo_bytes); };
    auto halo_base = [&](int q) -> unsigned char* { return a.resident ? recs + 2 * own_bytes : own_base(q) + own_bytes; };
    // K == 1: every iteration works on the same super-tile, one descriptor buffer; else four (q + 2 overwrites q - 2)
    auto desc_of = [&](int q) -> const unsigned char* { return descs + (K == 1 ? 0 : (q & 3)) * (size_t)a.desc_stride; };
    auto vin_of = [&](int q) -> const double* { const int j = q / K; return j == 0 ? a.bufX : (j == 1 ? a.bufA : a.bufB); };

    long long dbg_ring = 0, dbg_edge = 0, dbg_upd = 0, dbg_t = 0;
    auto stamp = [&](int slot) { if (DBG && t == 0 && slot < 56) a.dbg[(size_t)c * 64 + slot] = clock64(); };
    stamp(0);
    if (ln == 0) for (int e = 0; e < D; e++) mbar_init(&bar_ring[w][e], 1);
    if (t == 0) { mbar_init(&bar_recs[0], VNT); mbar_init(&bar_recs[1], VNT); }
    __syncthreads();

    auto copy_desc = [&](int q) {
        if (q >= Q || (K == 1 && q > 0)) return;
        const unsigned char* src = a.desc + ((size_t)c * K + (q % K)) * (size_t)a.desc_stride;
        unsigned char* dst = descs + (K == 1 ? 0 : (q & 3)) * (size_t)a.desc_stride;      // q + 2 overwrites q - 2, which no producer can still be reading
        for (int b = t * 16; b < a.desc_stride; b += VNT * 16) cp_async16(dst + b, src + b);
    };
    // records of iteration q: the own rows of its super-tile (skipped when they are already resident) and its halo rows
    auto copy_recs = [&](int q, bool own_too) {
        if (q >= Q) return;
        const unsigned char* d = desc_of(q);
        const int* di = reinterpret_cast<const int*>(d);
        const long row0 = di[0]; const int nrows = di[1] * VT, nhalo = di[2];
        const int* ids = reinterpret_cast<const int*>(d + 32 + 32 * a.max_ent);
        const double* vin = vin_of(q);
        const int k16 = (t & 3) << 4;
        if (own_too) {
            unsigned char* ob = own_base(q);
            const unsigned char* src = reinterpret_cast<const unsigned char*>(vin + 8 * row0);
            for (int i = t; i < 4 * nrows; i += VNT) {
                const int row = i >> 2;
                cp_async16(ob + 64 * row + (k16 ^ (((row >> 1) & 3) << 4)), src + 64 * (size_t)row + k16);
            }
        }
        unsigned char* hb = halo_base(q);
        for (int i = t; i < 4 * nhalo; i += VNT) {
            const int h = i >> 2;
            cp_async16(hb + 64 * h + (k16 ^ (((h >> 1) & 3) << 4)), reinterpret_cast<const unsigned char*>(vin + 8 * (long)ids[h]) + k16);
        }
        cp_async_mbar_arrive(&bar_recs[q & 1]);
    };
    // this warp's edge-stream producer (its lane 0): next chunk = chunk p_chunk of warp-tile p_e of iteration p_q.  An entry is
    // refilled right after the warp itself has finished with it (__syncwarp orders the lanes' reads before the refill).
    int p_q = 0, p_e = w, p_chunk = 0, issued = 0, consumed = 0;
    auto produce = [&](int q_visible) {
        while (issued - consumed < D && p_q <= q_visible && p_q < Q) {
            const unsigned char* d = desc_of(p_q);
            if (p_e >= reinterpret_cast<const int*>(d)[4]) { p_q++; p_e = w; p_chunk = 0; continue; }
            const VEnt* en = reinterpret_cast<const VEnt*>(d + 32) + p_e;
            const int rounds = en->rounds;
            const int nchunks = (rounds + R - 1) / R;
            if (p_chunk >= nchunks) { p_e += VNW; p_chunk = 0; continue; }
            const int nr = min(R, rounds - p_chunk * R);
            const unsigned bytes = unsigned(nr) * unsigned(VW * 26);
            const int e = issued % D;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");      // the warp's reads of this entry (generic proxy) before the refill (async proxy)
            mbar_expect_tx(&bar_ring[w][e], bytes);
            bulk_g2s(ring + e * (size_t)ring_bytes, a.vslots + (en->vblk0 + (long)p_chunk * R) * (long)(VW * 26), bytes, &bar_ring[w][e]);
            issued++; p_chunk++;
        }
    };

    // ---- prologue: static data (descriptors, first edge chunks) may be fetched while the predecessor drains ----
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    copy_desc(0); copy_desc(1);
    cp_async_wait_all();
    __syncthreads();
    if (ln == 0) produce(1);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    unsigned long long E0 = 0;
    if (DIST) {
        E0 = *(volatile unsigned long long*)a.d.op_counter;
        if (t < a.d.nwait) {
            const unsigned long long* f = a.d.my_flags + a.d.wait_peers[t].rank;
            while (ld_acquire_sys(f) < E0) { __nanosleep(20); }
        }
        __syncthreads();
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
    if (need_min || DIST) {
        double val = BIG;
        if (need_min) {
            if (a.premin) {           // per-block minima left by the restrict / prolong kernel that produced bufX: every CTA reduces them itself
                for (int b = t; b < a.npremin; b += VNT) val = fmin(val, __ldcg(a.premin + b));
            } else {                  // the state came from elsewhere: reduce over this CTA's own nodes
                const long r0 = a.cta_rows[c], r1 = a.cta_rows[c + 1];
                for (long r = r0 + t; r < r1; r += VNT) val = fmin(val, 0.5 * (a.vol_root[r] / a.bufX[8 * r + 7]));
            }
#pragma unroll
            for (int dlt = 16; dlt > 0; dlt >>= 1) val = fmin(val, __shfl_xor_sync(0xffffffffu, val, dlt));
            if (ln == 0) red[w][0] = val;
        }
        __syncthreads();
        if (need_min) {
            val = red[ln & (VNW - 1)][0];
#pragma unroll
            for (int dlt = 8; dlt > 0; dlt >>= 1) val = fmin(val, __shfl_xor_sync(0xffffffffu, val, dlt));       // every thread: this CTA's minimum
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
                    if (DIST) dist_allreduce(a.d, E0 + 1, &m, 1, true);
                    if (t == 0) {
                        *a.min_bits = (unsigned long long)__double_as_longlong(m);
                        __threadfence();
                        atomicAdd(a.bar, 0x10000u - unsigned(G));
                    }
                }
            }
            nbar = 1; nsync = 1;
        }
        __syncthreads();                 // red[][] is reused by the RMS sums
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
        if (ln == 0) produce(q + 1);
        __syncwarp();
        const unsigned char* d = desc_of(q);
        const int* di = reinterpret_cast<const int*>(d);
        const long row0 = di[0]; const int nent = di[4];
        const unsigned char* own = own_base(q);
        const unsigned char* halo = halo_base(q);
        const double* vold = a.bufX;
        double* vout = (j == 1) ? a.bufB : a.bufA;
        const double rk_div = double(MGCFD_RK + 1 - j), rk_rcp = 1.0 / rk_div;

        for (int ei = w; ei < nent; ei += VNW) {
            const VEnt* en = reinterpret_cast<const VEnt*>(d + 32) + ei;
            const int orow = en->orow0 + ln;                // this thread's row inside the super-tile
            const long gid = row0 + orow;
            // early loads for the update: in flight while the edge rounds run
            const double vol_or_sf = first_stage ? a.vol[gid] : a.sf[gid];
            const int brounds = en->brounds;
            const int bl = en->blane0 + ln;
            const unsigned char* bblk = a.bslots + en->bblk0 * (long)(VT * 25);
            BSlot b0 = {0, 0.0, 0.0, 0.0};
            if (brounds > 0) b0 = bslot_fetch<VT>(bblk, bl);
            const Rec me = sm_load_rec_off(own, (unsigned(orow) << 6) | (((unsigned(orow) >> 1) & 3u) << 4));
            double o[5];
            if (first_stage) { o[0] = me.rho; o[1] = me.mx; o[2] = me.my; o[3] = me.mz; o[4] = me.re; }
            else {
                const double2* p = reinterpret_cast<const double2*>(vold + 8 * gid);
                const double2 c0 = p[0], c1 = p[1];
                o[0] = c0.x; o[1] = c0.y; o[2] = c1.x; o[3] = c1.y; o[4] = vold[8 * gid + 4];
            }
            Flux5 f = {0.0, 0.0, 0.0, 0.0, 0.0};
            const double me_ep = me.re + me.p;
            const int rounds = en->rounds;
            for (int r0 = 0; r0 < rounds; r0 += R) {
                const int e = consumed % D;
                if (DBG) dbg_t = clock64();
                mbar_wait(&bar_ring[w][e], (consumed / D) & 1);
                if (DBG) { const long long now = clock64(); dbg_ring += now - dbg_t; dbg_t = now; }
                visit_edge_rounds(ring + e * (size_t)ring_bytes, min(R, rounds - r0), own, halo, ln, me, me_ep, a.k2, f);
                if (DBG) { const long long now = clock64(); dbg_edge += now - dbg_t; }
                __syncwarp();                               // every lane is done with the entry: lane 0 may refill it
                consumed++;
                if (ln == 0) produce(q + 1);
            }
            if (DBG) dbg_t = clock64();
            boundary_rounds<VT>(bblk, brounds, bl, 7, me, f, b0);
            double sfv = vol_or_sf;
            if (first_stage) {
                if (!have_min) {          // first update of the visit: barrier 0 must have completed
                    if (ln == 0) { while (bar_done(a.bar, gen0) < 1u) { } }
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
        __syncthreads();        // every warp is done with this iteration's record buffers
        stamp(6 + 4 * q);
        if (k == K - 1 && !last_stage) {
            // ---- stage barrier: all CTAs (all ranks' neighbours) have written the new state ----
            if (w == 0) {
                unsigned old = 0;
                // (the counter is shared by all barriers of the launch: never arrive at one before the previous one has completed --
                // only the legacy multi-GPU case can get here without having waited for barrier 0)
                if (t == 0) { while (bar_done(a.bar, gen0) < nbar) { } if (DIST) __threadfence_system(); else __threadfence(); old = atomicAdd(a.bar, 1u); }
                old = __shfl_sync(0xffffffffu, old, 0);
                if ((old & 0xFFFFu) == unsigned(G - 1)) {
                    if (DIST) dist_signal_wait_peers(a.d, E0 + nsync + 1, true);
                    if (t == 0) { __threadfence(); atomicAdd(a.bar, 0x10000u - unsigned(G)); }
                }
                if (t == 0) { while (bar_done(a.bar, gen0) < nbar + 1u) { } }
            }
            nbar++; nsync++;
            __syncthreads();
            copy_recs(q + 1, !a.resident);
        }
    }

    if (DBG && t == 0) {
        long long* o = a.dbg + (size_t)c * 64;
        o[56] = dbg_ring; o[57] = dbg_edge; o[58] = dbg_upd; o[59] = clock64();
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
            __syncthreads();
            if (t < 5) {
                double s = 0.0;
                for (int ww = 0; ww < VNW; ww++) s += red[ww][t];
                a.cta_rms[c * 5 + t] = s;
            }
            __syncthreads();
        }
        if (w == 0) {
            unsigned old = 0;
            if (t == 0) { while (bar_done(a.bar, gen0) < nbar) { } if (DIST) __threadfence_system(); else __threadfence(); old = atomicAdd(a.bar, 1u); }
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
                    if (DIST) dist_allreduce(a.d, E0 + nsync + 1, v5, 5, false);       // also this kernel's end signal to every rank
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
                    if (DIST) *a.d.op_counter = E0 + nsync + 1;
                    __threadfence();
                    atomicAdd(a.bar, 0x10000u - unsigned(G));
                }
            }
        }
    }
}

}  // namespace mgcfd
