// context.cu -- device-resident level state, kernel launch sequencing (granular + fused/graph paths) and the
// extern "C" boundary declared in include/mgcfd_b200.h.  No CPU fallback: every compute entry point launches
// CUDA kernels; without a device mgcfd_create fails.
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <string>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <vector>

#include "../../include/mgcfd_b200.h"
#include "host_mesh.h"
#include "kernels.cuh"
#include "visit_kernel.cuh"
#include "assess_kernels.cuh"
#include "plan.h"
#include "partition.h"
#include "../../include/mgcfd_dist.h"

#include <dlfcn.h>
#include <nccl.h>

using namespace mgcfd;

namespace {

thread_local std::string g_err;

#define CK(call)                                                                                         \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) {                                                                         \
            g_err = std::string(#call) + ": " + cudaGetErrorString(e_) + " (" __FILE__ ":" + std::to_string(__LINE__) + ")"; \
            return MGCFD_ERR_CUDA;                                                                       \
        }                                                                                                \
    } while (0)
#define CKRC(call) do { int rc_ = (call); if (rc_ != MGCFD_OK) return rc_; } while (0)

// ---- guard zones (MGCFD_GUARD=1; a debugging aid, off by default) ---------------------------------------------------------------
// Every device allocation of this file gets GUARD bytes of a known pattern on both sides (and the pattern in the allocation itself
// until it is first written); mgcfd_guard_check() reads the zones back and reports the allocations whose zones were written to.
// A linear overrun of a per-level array -- the one class of bug compute-sanitizer's memcheck finds that parity tests do not --
// shows up as a damaged zone with the source line of the allocation.  (The peer-to-peer slab is exempt: its address travels as
// a CUDA IPC handle.)
const size_t GUARD_BYTES = 64 << 10;
struct GuardRec { char* base; size_t bytes, padded; int line; bool gap; };      // gap: a zone inside the slab (nothing to free)
std::mutex g_guard_mu;
std::map<void*, GuardRec> g_guard;
inline bool guard_on() { const char* e = getenv("MGCFD_GUARD"); return e && atoi(e) != 0; }
cudaError_t guarded_malloc(void** p, size_t bytes, int line) {
    if (!guard_on()) return cudaMalloc(p, bytes);
    const size_t padded = (bytes + 255) & ~size_t(255);
    char* base = nullptr;
    cudaError_t e = cudaMalloc((void**)&base, padded + 2 * GUARD_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaMemset(base, 0xA5, padded + 2 * GUARD_BYTES);
    if (e != cudaSuccess) return e;
    e = cudaDeviceSynchronize();      // the fill runs on the null stream, the uploads that follow on the context's non-blocking stream: order them
    if (e != cudaSuccess) return e;
    *p = base + GUARD_BYTES;
    std::lock_guard<std::mutex> lk(g_guard_mu);
    g_guard[*p] = GuardRec{base, bytes, padded, line, false};
    return cudaSuccess;
}
cudaError_t guarded_free(void* p) {
    if (p) {
        std::lock_guard<std::mutex> lk(g_guard_mu);
        auto it = g_guard.find(p);
        if (it != g_guard.end()) { char* base = it->second.base; g_guard.erase(it); return cudaFree(base); }
    }
    return cudaFree(p);
}
void guard_gap_add(char* at, int line) {
    std::lock_guard<std::mutex> lk(g_guard_mu);
    g_guard[at] = GuardRec{at, 0, 0, line, true};
}
void guard_gaps_drop(char* lo, size_t len) {
    std::lock_guard<std::mutex> lk(g_guard_mu);
    for (auto it = g_guard.begin(); it != g_guard.end();) {
        if (it->second.gap && it->second.base >= lo && it->second.base < lo + len) it = g_guard.erase(it); else ++it;
    }
}
// damaged zones, as text ("line L: first damaged byte at <offset relative to the allocation>"); returns their number
int guard_check(std::string& report) {
    cudaDeviceSynchronize();
    std::lock_guard<std::mutex> lk(g_guard_mu);
    std::vector<unsigned char> h(GUARD_BYTES + 256);
    int bad = 0;
    for (const auto& kv : g_guard) {
        const GuardRec& r = kv.second;
        struct Zone { const char* at; size_t len; long long rel; };      // rel: offset of the zone's first byte from the allocation's first
        std::vector<Zone> zones;
        if (r.gap) zones.push_back({r.base, GUARD_BYTES, 0});
        else {
            zones.push_back({r.base, GUARD_BYTES, -(long long)GUARD_BYTES});
            zones.push_back({r.base + GUARD_BYTES + r.bytes, r.padded - r.bytes + GUARD_BYTES, (long long)r.bytes});
        }
        for (const Zone& z : zones) {
            if (cudaMemcpy(h.data(), z.at, z.len, cudaMemcpyDeviceToHost) != cudaSuccess) { report += "guard zone unreadable; "; bad++; continue; }
            for (size_t i = 0; i < z.len; i++)
                if (h[i] != 0xA5) {
                    report += std::string(r.gap ? "slab gap after sub-buffer, line " : "allocation at line ") + std::to_string(r.line) + ": first damaged byte at offset " +
                              std::to_string(z.rel + (long long)i) + (r.gap ? " of the gap; " : " (size " + std::to_string(r.bytes) + "); ");
                    bad++;
                    break;
                }
        }
    }
    return bad;
}
#define cudaMalloc(p, n) guarded_malloc((void**)(p), (n), __LINE__)
#define cudaFree(p) guarded_free(p)

template <class T>
int dev_upload(T** dptr, const std::vector<T>& h, cudaStream_t s) {
    *dptr = nullptr;
    const size_t bytes = std::max<size_t>(h.size(), 1) * sizeof(T);
    CK(cudaMalloc((void**)dptr, bytes));
    if (!h.empty()) CK(cudaMemcpyAsync(*dptr, h.data(), h.size() * sizeof(T), cudaMemcpyHostToDevice, s));
    return MGCFD_OK;
}

enum KernelId { K_STEP = 0, K_FLUX = 1, K_UPDATE = 2, K_INDIRECT = 3, K_TIME = 4, K_RESTRICT = 5, K_PROLONG = 6, K_COUNT = 7 };

struct Level {
    HostLevel host;            // kept until finalize (transfer operators need mg + coords of both levels)
    LevelPlan plan;
    bool uploaded = false;
    long nel = 0, npad = 0, ntiles = 0, nI = 0, nB = 0, nW = 0;
    long ncomp = 0;            // rows this rank computes (the owned tiles, = ntiles * TN); rows [ncomp, npad) are ghosts
    long nel_global = 0;       // nodes of the level over all ranks (RMS normalisation)
    long nI_global = 0;        // internal edges of the level over all ranks (work accounting of distributed runs)
    int TN = 256, smem_nodes = 0;
    size_t smem_bytes = 0;        // simple stage kernel
    size_t pipe_smem = 0;         // pipelined stage kernel
    int pipe_grid = 0, pipe_grid_dist = 0, chunk_rounds = 0;
    bool pipe = false, untiled = false;   // untiled: the numbering has no locality (atomic baseline only)
    double* buf[3] = {nullptr, nullptr, nullptr};   // node records (8 doubles each), rotating roles
    int i_var = 0, i_old = 1, i_tmp = 2;
    double *res = nullptr, *flux = nullptr, *sf = nullptr, *vol = nullptr, *vol_root = nullptr;
    int *new_of_old = nullptr, *old_of_new = nullptr;
    unsigned char *hdrs = nullptr, *slots = nullptr, *bslots = nullptr;
    // flat + CSR (lazy)
    int *ea = nullptr, *eb = nullptr; double* ew = nullptr;
    int* bnode = nullptr; uint8_t* bkind = nullptr; double* bw = nullptr;
    bool flat_up = false;
    double* soa = nullptr;                         // the conserved variables as five planes (assess-compute SoA variant only)
    double* ewt_pre = nullptr;                     // |e| per internal edge (assess-compute variants with precomputed weights)
    // transfers (operators between this level and the next coarser one)
    long* child_off = nullptr; int* child_ids = nullptr;       // stored on the COARSE level (children in level-1)
    int* parent = nullptr; double* idist_own = nullptr; long* ent_off = nullptr; int* ent_src = nullptr; double* ent_w = nullptr;
    double* rms_partial = nullptr; long rms_parts = 0;
    double* blockmins = nullptr;
    double* io = nullptr;      // AoS staging for get/set_field
    // distributed runs: halo exchange lists (partition.h)
    std::vector<long> send_off, recv_off, gid, send_list;
    long nsend = 0, nghost = 0, n_owned = 0;
    int* d_send_idx = nullptr; double *sendbuf = nullptr, *recvtmp = nullptr;
    P2PPeer* d_peers = nullptr; int npeers = 0;    // p2p: peers of this level
    // multi-GPU, peer-to-peer plane: node -> (peer, remote row) CSR, per-tile flags, where the peers keep their copies
    std::vector<int> h_send_idx;                   // send list in device numbering (host copy)
    std::vector<P2PPeer> h_peers;                  // host copy of d_peers
    int *d_tgt_off = nullptr, *d_tgt_peer = nullptr, *d_tgt_row = nullptr;
    unsigned char* d_tile_sends = nullptr;
    PeerOut* d_peer_out = nullptr;
    int* d_order_tiles = nullptr;                  // tiles in the order the distributed stage kernel takes them (ghost-reading tiles last)
    unsigned char *d_rblk_wait = nullptr, *d_pblk_wait = nullptr;   // per 128-row block: restrict INTO this level reads a fine ghost row / prolong INTO this level reads a coarse ghost residual
    int n_send_tiles = 0;
    // the persistent visit kernel (visit_kernel.cuh): one launch per smoothing visit
    bool visit = false;
    int vK = 1, vG = 0, vR = 1, vD = 2, vW = 16, v_resident = 0, v_srmax = 0;
    unsigned long long* d_minword = nullptr;       // multi-GPU: the minimum dt the transfer kernels fold their blocks' minima into (kernels.cuh, DistTail::minword)
    int minword_state = 0;                         // 0: +inf (clean)  1: holds the minimum of the level's CURRENT state  2: holds something stale
    bool premin_valid = false;                     // blockmins holds the per-block minima of dt of the CURRENT state (left by restrict / prolong)
    size_t v_smem = 0;
    unsigned char *d_desc = nullptr, *d_vslots = nullptr;
    double* d_hsum = nullptr;
    int* d_cta_rows = nullptr;
    double* V(int i) const { return buf[i]; }
};

}  // namespace

// NCCL is resolved at run time (dlopen) so that single-GPU users of the library need no NCCL at all
struct NcclApi {
    void* handle = nullptr;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
    ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
    ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
    ncclResult_t (*GroupStart)() = nullptr;
    ncclResult_t (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
struct Dist {
    bool active = false;
    int rank = 0, nranks = 1;
    ncclComm_t comm = nullptr;
    long exchanges = 0;
    // direct peer-to-peer exchange (CUDA IPC windows; mgcfd_dist_p2p_prepare / _attach): replaces NCCL on the data path
    bool p2p = false, want_graph = false;
    unsigned char* win = nullptr;              // my window: flags[64] | red[2][64][8] | staging per level (2 parities); the start of the slab
    size_t win_bytes = 0;
    std::vector<size_t> buf_off;               // per level: byte offsets of buf[0..2] and res inside the slab (4 per level)
    std::vector<long> stage_off;               // per level: offset (in doubles, from the window's staging base) of parity 0; parity 1 follows
    std::vector<void*> peer_win;               // [nranks] mapped windows (mine at [rank])
    unsigned long long* d_op = nullptr;        // device: operation number (identical on every rank), bumped by each operation's kernel
    unsigned int* d_ctr = nullptr;             // device: [0] reductions so far, [1 + level] exchanges of that level so far (staging parities)
    double** d_red_of_rank = nullptr;          // device: [nranks] window reduction bases
    unsigned long long** d_flag_of_rank = nullptr;   // device: [nranks] &window.flags[me]
    unsigned int* d_ticket = nullptr;
    int off = 0;                               // epochs used by kernels enqueued since the last k_epoch_advance (kernels.cuh "Epochs")
    P2PPeer* d_sig = nullptr; int nsig = 0;    // every neighbour rank (any level): the announcement targets
};
constexpr size_t P2P_FLAGS_BYTES = 64 * 8, P2P_RED_BYTES = 2 * 64 * 8 * 8, P2P_HDR_BYTES = P2P_FLAGS_BYTES + P2P_RED_BYTES;

struct mgcfd_ctx {
    mgcfd_options opt;
    Dist dist;
    double* d_rms_sums = nullptr;
    unsigned char* slab = nullptr;             // ONE allocation: [peer-to-peer window |] record buffers and residual planes of every level
    size_t slab_bytes = 0;                     // (one CUDA IPC handle exposes everything other ranks write)
    unsigned int* d_bar = nullptr;             // grid barrier word of the visit kernel
    double* d_cta_min = nullptr;               // [num_sms] per-CTA minima, [num_sms * 5] per-CTA RMS sums
    double* d_cta_rms = nullptr;
    std::map<std::string, long> graph_launches;
    std::map<std::string, std::string> graph_flags_end;
    long long* d_visit_dbg = nullptr;          // MGCFD_VISIT_DEBUG=1: clock stamps of the last visit-kernel launch (64 per CTA)
    int levels = 0, variant = 2;
    bool finalized = false;
    std::vector<Level> L;
    cudaStream_t stream = nullptr;
    double ff[5], ffc[12];
    bool have_ff = false;
    unsigned long long* d_minbits = nullptr;
    unsigned int* d_ticket = nullptr;
    unsigned long long* d_badkey = nullptr;
    double* d_rms = nullptr;       // [cap][6]
    int* d_rms_counter = nullptr;
    int rms_cap = 0;
    long launches = 0;
    std::map<std::string, cudaGraphExec_t> graphs;
    // timing
    std::vector<double> t_ms;      // [K_COUNT][levels]
    std::vector<long> t_iters;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // per-kernel timing without host synchronisation: start/stop events are drawn from a pool and resolved lazily
    struct PendingTime { int kid, lev; long iters; cudaEvent_t e0, e1; };
    std::vector<cudaEvent_t> ev_pool;
    size_t ev_next = 0;
    std::vector<PendingTime> pending;
    bool capturing = false;
    unsigned long long stage_seq = 0;
    double kdiss;
    int num_sms = 148;
    int rms_pending = 0;
    long bad_cell = -1; int bad_reason = 0;
};

namespace {

inline long blocks_for(long n, int bs) { return (n + bs - 1) / bs; }
int env_int(const char* name, int dflt) { const char* e = getenv(name); return (e && *e) ? atoi(e) : dflt; }

// Launch of a many-block kernel (transfers, RMS), optionally (MGCFD_PDL_TRANSFERS=1) as a programmatic dependent of its predecessor
// in the stream: its blocks are then scheduled while the predecessor -- a stage kernel, which releases its dependents at its start --
// still runs, and sit in griddepcontrol.wait (pdl_wait) until it has completed; the kernel never releases ITS dependents early.
// Measured on B200 (profiles/r02H_pdl_transfers_ab.txt): 0.3695 vs 0.3494 ms per C2 cycle -- SLOWER, like round 1's variant that
// also released early: blocks parked in the wait hold warp slots and registers next to the persistent stage CTAs for the whole
// stage.  Hence off by default; kept as a measured switch.
template <class... KArgs, class... Args>
int launch_dependent(mgcfd_ctx* c, void (*kernel)(KArgs...), unsigned grid, unsigned block, Args&&... args) {
    static const int on = env_int("MGCFD_PDL_TRANSFERS", 0);
    if (!on || c->opt.no_pdl) {
        kernel<<<grid, block, 0, c->stream>>>(KArgs(args)...);
        return MGCFD_OK;
    }
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.dynamicSmemBytes = 0; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...));
    return MGCFD_OK;
}

// Tile size when the caller leaves it to the library (measured on B200, profiles/):
//  * multi-million-node levels: 128-node tiles (more CTAs in flight per SM; 202 vs 197 cycles/s on the 8 M-node mesh);
//  * smaller levels of low-degree meshes (hex-dual, <= 4 internal edges per node: a 128-node tile's whole edge stream fits the ring
//    at 3 CTAs per SM): whichever of 128 / 256 needs less time in WAVES of the persistent grid (3 resp. 2 CTAs per SM), a 256-node
//    tile costing 1.43x a 128-node one (C2 level 0: 6 waves of 128 = 23.2 us, 4 waves of 256 = 22.1 us).  On the M6-shaped mesh this
//    picks 256 for the 300 K-node level and 128 for the three coarser ones: 2631 -> ~2790 V-cycles/s;
//  * otherwise 256 (smaller halo share per tile).
inline int auto_tile_nodes(long n, long nI, int num_sms) {
    if (n >= 1000000) return 128;
    if (nI <= 4 * n) {
        const long w128 = blocks_for(blocks_for(n, 128), 3L * num_sms), w256 = blocks_for(blocks_for(n, 256), 2L * num_sms);
        return (143 * w256 < 100 * w128) ? 256 : 128;
    }
    return 256;
}

// folds every recorded (start, stop) pair into the per-kernel per-level totals; synchronises the stream once
void resolve_times(mgcfd_ctx* c) {
    if (c->pending.empty()) return;
    cudaStreamSynchronize(c->stream);
    for (const auto& p : c->pending) {
        float ms = 0;
        if (cudaEventElapsedTime(&ms, p.e0, p.e1) == cudaSuccess) {
            c->t_ms[p.kid * c->levels + p.lev] += ms;
            c->t_iters[p.kid * c->levels + p.lev] += p.iters;
        }
    }
    c->pending.clear();
    c->ev_next = 0;
}

// hands out the next pooled event; the caller has made sure (reserve_events) that the pool will not be recycled
// between the start and the stop event of one bracket
cudaEvent_t pool_event(mgcfd_ctx* c) {
    if (c->ev_next == c->ev_pool.size()) { cudaEvent_t e = nullptr; cudaEventCreate(&e); c->ev_pool.push_back(e); }
    return c->ev_pool[c->ev_next++];
}
void reserve_events(mgcfd_ctx* c, size_t n) {
    if (c->ev_next + n > 8192) resolve_times(c);      // recycle: one stream sync every ~4096 timed launches
}

// CUDA-event bracket around the launches of one reference kernel on one level (the analogue of the reference's
// start_timer()/stop_timer() pairs, src/Monitoring/timer.cpp:58-104), recorded on the context's own stream
struct Timed {
    mgcfd_ctx* c; int kid, lev; long iters; bool on; cudaEvent_t e0 = nullptr, e1 = nullptr;
    Timed(mgcfd_ctx* c_, int kid_, int lev_, long iters_) : c(c_), kid(kid_), lev(lev_), iters(iters_) {
        on = c->opt.timing && !c->capturing;
        if (!on) return;
        reserve_events(c, 2);
        e0 = pool_event(c); e1 = pool_event(c);
        cudaEventRecord(e0, c->stream);
    }
    ~Timed() {
        if (!on) return;
        cudaEventRecord(e1, c->stream);
        c->pending.push_back({kid, lev, iters, e0, e1});
    }
};

int post_launch(mgcfd_ctx* c) {
    c->launches++;
    CK(cudaGetLastError());
    return MGCFD_OK;
}

// ---- distributed runs: NCCL over NVLink ---------------------------------------------------
NcclApi g_nccl;
int nccl_load() {
    if (g_nccl.handle) return MGCFD_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    void* h = nullptr;
    for (const char* n : names) { h = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (h) break; }
    if (!h) { g_err = std::string("cannot load NCCL (libnccl.so.2): ") + dlerror(); return MGCFD_ERR_COMM; }
    bool ok = true;
    auto sym = [&](const char* n) { void* p = dlsym(h, n); if (!p) ok = false; return p; };
    g_nccl.GetUniqueId = (decltype(g_nccl.GetUniqueId))sym("ncclGetUniqueId");
    g_nccl.CommInitRank = (decltype(g_nccl.CommInitRank))sym("ncclCommInitRank");
    g_nccl.CommDestroy = (decltype(g_nccl.CommDestroy))sym("ncclCommDestroy");
    g_nccl.Send = (decltype(g_nccl.Send))sym("ncclSend");
    g_nccl.Recv = (decltype(g_nccl.Recv))sym("ncclRecv");
    g_nccl.AllReduce = (decltype(g_nccl.AllReduce))sym("ncclAllReduce");
    g_nccl.GroupStart = (decltype(g_nccl.GroupStart))sym("ncclGroupStart");
    g_nccl.GroupEnd = (decltype(g_nccl.GroupEnd))sym("ncclGroupEnd");
    g_nccl.GetErrorString = (decltype(g_nccl.GetErrorString))sym("ncclGetErrorString");
    if (!ok) { g_err = "libnccl is missing a required symbol"; dlclose(h); return MGCFD_ERR_COMM; }
    g_nccl.handle = h;
    return MGCFD_OK;
}
#define NK(call)                                                                                          \
    do {                                                                                                  \
        ncclResult_t r_ = (call);                                                                         \
        if (r_ != ncclSuccess) { g_err = std::string(#call) + ": " + g_nccl.GetErrorString(r_); return MGCFD_ERR_COMM; } \
    } while (0)

// multi-GPU with the peer-to-peer plane: the kernels deliver the rows they produce themselves
bool dist_inkernel(mgcfd_ctx* c) { return c->dist.active && c->dist.p2p; }

// the kernels enqueued since the last advance have used d.off epochs: move the base word on (kernels.cuh "Epochs")
int dist_flush_epoch(mgcfd_ctx* c) {
    Dist& d = c->dist;
    if (!d.p2p || d.off == 0) return MGCFD_OK;
    k_epoch_advance<<<1, 1, 0, c->stream>>>(d.d_op, d.off);
    d.off = 0;
    c->launches++;
    CK(cudaGetLastError());
    return MGCFD_OK;
}
// global minimum of the per-rank min-dt bit patterns (positive doubles order like their bits)
int p2p_allreduce(mgcfd_ctx* c, double* vals, int n, int is_min) {
    Dist& d = c->dist;
    CKRC(dist_flush_epoch(c));      // this kernel works with the absolute epoch
    k_p2p_allreduce<<<1, 64, 0, c->stream>>>(vals, n, is_min, d.nranks, d.rank, d.d_red_of_rank, d.d_flag_of_rank, (const unsigned long long*)d.win,
                                              (const double*)(d.win + P2P_FLAGS_BYTES), d.d_op, d.d_ctr);
    return post_launch(c);
}
int dist_allreduce_min(mgcfd_ctx* c) {
    if (!c->dist.active) return MGCFD_OK;
    if (c->dist.p2p) return p2p_allreduce(c, (double*)c->d_minbits, 1, 1);
    NK(g_nccl.AllReduce(c->d_minbits, c->d_minbits, 1, ncclUint64, ncclMin, c->dist.comm, c->stream));
    return MGCFD_OK;
}
int dist_allreduce_sum5(mgcfd_ctx* c) {
    if (!c->dist.active) return MGCFD_OK;
    if (c->dist.p2p) return p2p_allreduce(c, c->d_rms_sums, 5, 0);
    NK(g_nccl.AllReduce(c->d_rms_sums, c->d_rms_sums, 5, ncclDouble, ncclSum, c->dist.comm, c->stream));
    return MGCFD_OK;
}
// ghost records of level l in buffer `recs` <- their owners' records: pack the rows every peer needs, one grouped
// ncclSend/ncclRecv per peer, received straight into the ghost rows (contiguous per owner)
// one kernel per exchange on every rank: put into the peers' staging buffers, signal, wait for the peers, unpack (kernels.cuh)
template <int WIDTH, bool SOA>
int p2p_exchange(mgcfd_ctx* c, int l, const double* src, double* dst) {
    Dist& d = c->dist;
    Level& v = c->L[l];
    const double* stage = (const double*)(d.win + P2P_HDR_BYTES) + d.stage_off[l];
    CKRC(dist_flush_epoch(c));
    const long work = std::max(v.nsend, v.nghost) * WIDTH;
    const unsigned grid = (unsigned)std::max<long>(1, std::min<long>(blocks_for(work, 256), 2L * c->num_sms));   // resident at once
    k_p2p_exchange<WIDTH, SOA><<<grid, 256, 0, c->stream>>>(src, v.npad, v.d_send_idx, v.d_peers, v.npeers, d.d_op, d.d_ctr + 1 + l, d.d_ticket,
                                                            (const unsigned long long*)d.win, stage, 8 * std::max<long>(v.nghost, 1), dst, v.ncomp);
    d.exchanges++;
    return post_launch(c);
}
int dist_exchange_records(mgcfd_ctx* c, int l, double* recs) {
    if (!c->dist.active) return MGCFD_OK;
    Level& v = c->L[l];
    // peer-to-peer plane: the epoch number (Dist::d_op) must advance alike on EVERY rank, so a rank without halo at this level
    // still runs the (empty) operation -- skipping it would leave the rank one epoch behind for good
    if (c->dist.p2p) return p2p_exchange<8, false>(c, l, recs, recs);
    if (v.nsend == 0 && v.nghost == 0) return MGCFD_OK;
    if (v.nsend) { k_pack_records<<<(unsigned)blocks_for(4 * v.nsend, 256), 256, 0, c->stream>>>(recs, v.d_send_idx, v.nsend, v.sendbuf); CKRC(post_launch(c)); }
    NK(g_nccl.GroupStart());
    for (int p = 0; p < c->dist.nranks; p++) {
        const long ns = v.send_off[p + 1] - v.send_off[p], nr = v.recv_off[p + 1] - v.recv_off[p];
        if (ns) NK(g_nccl.Send(v.sendbuf + 8 * v.send_off[p], 8 * (size_t)ns, ncclDouble, p, c->dist.comm, c->stream));
        if (nr) NK(g_nccl.Recv(recs + 8 * (v.ncomp + v.recv_off[p]), 8 * (size_t)nr, ncclDouble, p, c->dist.comm, c->stream));
    }
    NK(g_nccl.GroupEnd());
    c->dist.exchanges++;
    return MGCFD_OK;
}
int dist_exchange_residuals(mgcfd_ctx* c, int l) {
    if (!c->dist.active) return MGCFD_OK;
    Level& v = c->L[l];
    if (c->dist.p2p) return p2p_exchange<5, true>(c, l, v.res, v.res);
    if (v.nsend == 0 && v.nghost == 0) return MGCFD_OK;
    if (v.nsend) { k_pack_soa5<<<(unsigned)blocks_for(v.nsend, 256), 256, 0, c->stream>>>(v.res, v.npad, v.d_send_idx, v.nsend, v.sendbuf); CKRC(post_launch(c)); }
    NK(g_nccl.GroupStart());
    for (int p = 0; p < c->dist.nranks; p++) {
        const long ns = v.send_off[p + 1] - v.send_off[p], nr = v.recv_off[p + 1] - v.recv_off[p];
        if (ns) NK(g_nccl.Send(v.sendbuf + 5 * v.send_off[p], 5 * (size_t)ns, ncclDouble, p, c->dist.comm, c->stream));
        if (nr) NK(g_nccl.Recv(v.recvtmp + 5 * v.recv_off[p], 5 * (size_t)nr, ncclDouble, p, c->dist.comm, c->stream));
    }
    NK(g_nccl.GroupEnd());
    if (v.nghost) { k_unpack_soa5<<<(unsigned)blocks_for(v.nghost, 256), 256, 0, c->stream>>>(v.res, v.npad, v.ncomp, v.nghost, v.recvtmp); CKRC(post_launch(c)); }
    c->dist.exchanges++;
    return MGCFD_OK;
}

template <int TN, bool SCATTER, bool FUSED>
int launch_stage_t(mgcfd_ctx* c, Level& v, const StageArgs& a) {
    static size_t attr_bytes = 0;
    if (v.smem_bytes > attr_bytes) {
        CK(cudaFuncSetAttribute(k_stage<TN, SCATTER, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.smem_bytes));
        attr_bytes = v.smem_bytes;
    }
    k_stage<TN, SCATTER, FUSED><<<(unsigned)v.ntiles, TN, v.smem_bytes, c->stream>>>(a);
    return post_launch(c);
}
template <int TN, bool SCATTER, bool DIST = false>
int launch_pipe_t(mgcfd_ctx* c, Level& v, const StageArgs& a) {
    const int grid = DIST ? v.pipe_grid_dist : v.pipe_grid;
    if (c->opt.no_pdl) {
        k_stage_pipe<TN, SCATTER, DIST><<<(unsigned)grid, TN, v.pipe_smem, c->stream>>>(a);
        return post_launch(c);
    }
    // programmatic dependent launch: this kernel may start (barrier set-up, header + edge-stream prefetch: static data) while
    // its predecessor in the stream drains; it synchronises with griddepcontrol.wait before reading node state
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)grid); cfg.blockDim = dim3(TN); cfg.dynamicSmemBytes = v.pipe_smem; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    CK(cudaLaunchKernelEx(&cfg, k_stage_pipe<TN, SCATTER, DIST>, a));
    return post_launch(c);
}
// in-kernel delivery: occupancy and shared-memory attribute of the DIST instantiation (its register count differs)
template <int TN, bool SCATTER>
int setup_pipe_dist_t(mgcfd_ctx* c, Level& v) {
    static size_t attr_bytes = 0;       // per instantiation: the attribute only ever grows (levels differ in shared memory)
    if (v.pipe_smem > attr_bytes) {
        CK(cudaFuncSetAttribute(k_stage_pipe<TN, SCATTER, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.pipe_smem));
        attr_bytes = v.pipe_smem;
    }
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stage_pipe<TN, SCATTER, true>, TN, v.pipe_smem));
    if (per_sm < 1) { v.pipe_grid_dist = 0; return MGCFD_OK; }
    v.pipe_grid_dist = (int)std::min<long>(v.ntiles, (long)per_sm * c->num_sms);
    return MGCFD_OK;
}
int setup_pipe_dist(mgcfd_ctx* c, Level& v) {
    const bool sc = v.plan.scatter;
    if (v.TN == 128) return sc ? setup_pipe_dist_t<128, true>(c, v) : setup_pipe_dist_t<128, false>(c, v);
    if (v.TN == 256) return sc ? setup_pipe_dist_t<256, true>(c, v) : setup_pipe_dist_t<256, false>(c, v);
    return sc ? setup_pipe_dist_t<512, true>(c, v) : setup_pipe_dist_t<512, false>(c, v);
}
// persistent grid of the pipelined kernel: as many CTAs as fit on the device at once (occupancy API), never more than tiles
template <int TN, bool SCATTER>
int setup_pipe_t(mgcfd_ctx* c, Level& v) {
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k_stage_pipe<TN, SCATTER>));
    if (v.pipe_smem + fa.sharedSizeBytes > 227 * 1024) { v.pipe = false; return MGCFD_OK; }
    static size_t attr_bytes = 0;       // per instantiation: the attribute only ever grows
    if (v.pipe_smem > attr_bytes) {
        CK(cudaFuncSetAttribute(k_stage_pipe<TN, SCATTER>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)v.pipe_smem));
        attr_bytes = v.pipe_smem;
    }
    int per_sm = 0;
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_stage_pipe<TN, SCATTER>, TN, v.pipe_smem));
    if (per_sm < 1) { v.pipe = false; return MGCFD_OK; }
    v.pipe_grid = (int)std::min<long>(v.ntiles, (long)per_sm * c->num_sms);
    v.pipe = true;
    return MGCFD_OK;
}
int setup_pipe(mgcfd_ctx* c, Level& v) {
    const bool sc = v.plan.scatter;
    if (v.TN == 128) return sc ? setup_pipe_t<128, true>(c, v) : setup_pipe_t<128, false>(c, v);
    if (v.TN == 256) return sc ? setup_pipe_t<256, true>(c, v) : setup_pipe_t<256, false>(c, v);
    return sc ? setup_pipe_t<512, true>(c, v) : setup_pipe_t<512, false>(c, v);
}
int launch_stage(mgcfd_ctx* c, Level& v, const StageArgs& a, bool fused) {
    if (v.untiled) { g_err = "this level cannot be tiled (numbering without locality): only the atomic flux kernels run on it"; return MGCFD_ERR_ARG; }
    const bool sc = v.plan.scatter;
    if (fused && v.pipe) {
        if (v.TN == 128) return sc ? launch_pipe_t<128, true>(c, v, a) : launch_pipe_t<128, false>(c, v, a);
        if (v.TN == 256) return sc ? launch_pipe_t<256, true>(c, v, a) : launch_pipe_t<256, false>(c, v, a);
        return sc ? launch_pipe_t<512, true>(c, v, a) : launch_pipe_t<512, false>(c, v, a);
    }
    if (v.TN == 128) return sc ? (fused ? launch_stage_t<128, true, true>(c, v, a) : launch_stage_t<128, true, false>(c, v, a))
                               : (fused ? launch_stage_t<128, false, true>(c, v, a) : launch_stage_t<128, false, false>(c, v, a));
    if (v.TN == 256) return sc ? (fused ? launch_stage_t<256, true, true>(c, v, a) : launch_stage_t<256, true, false>(c, v, a))
                               : (fused ? launch_stage_t<256, false, true>(c, v, a) : launch_stage_t<256, false, false>(c, v, a));
    return sc ? (fused ? launch_stage_t<512, true, true>(c, v, a) : launch_stage_t<512, true, false>(c, v, a))
              : (fused ? launch_stage_t<512, false, true>(c, v, a) : launch_stage_t<512, false, false>(c, v, a));
}

int launch_stage_dist(mgcfd_ctx* c, Level& v, const StageArgs& a) {
    const bool sc = v.plan.scatter;
    if (v.TN == 128) return sc ? launch_pipe_t<128, true, true>(c, v, a) : launch_pipe_t<128, false, true>(c, v, a);
    if (v.TN == 256) return sc ? launch_pipe_t<256, true, true>(c, v, a) : launch_pipe_t<256, false, true>(c, v, a);
    return sc ? launch_pipe_t<512, true, true>(c, v, a) : launch_pipe_t<512, false, true>(c, v, a);
}

FarField far_field_of(const mgcfd_ctx* c) {
    FarField F;
    memcpy(F.v, c->ff, sizeof(F.v)); memcpy(F.c, c->ffc, sizeof(F.c));
    return F;
}
StageArgs base_args(mgcfd_ctx* c, Level& v) {
    StageArgs a;
    memset(&a, 0, sizeof(a));
    a.ff = far_field_of(c);
    a.stride = v.npad;
    a.hdrs = v.hdrs; a.hdr_stride = v.plan.hdr_stride;
    a.slots = v.slots; a.bslots = v.bslots;
    a.ntiles = (int)v.ntiles;
    a.rec_rows = v.smem_nodes;
    a.chunk_rounds = v.chunk_rounds;
    a.k2 = 2.0 * c->kdiss;
    a.old_of_new = v.old_of_new;
    a.sf = v.sf; a.vol = v.vol; a.min_bits = c->d_minbits; a.legacy = (c->variant == MGCFD_MESH_FVCORR);
    return a;
}

int ensure_flux(mgcfd_ctx* c, Level& v) {
    if (v.flux) return MGCFD_OK;
    CK(cudaMalloc((void**)&v.flux, sizeof(double) * 5 * v.npad));
    CK(cudaMemsetAsync(v.flux, 0, sizeof(double) * 5 * v.npad, c->stream));
    return MGCFD_OK;
}
int ensure_flat(mgcfd_ctx* c, Level& v) {
    if (v.flat_up) return MGCFD_OK;
    CKRC(dev_upload(&v.ea, v.plan.ea, c->stream)); CKRC(dev_upload(&v.eb, v.plan.eb, c->stream)); CKRC(dev_upload(&v.ew, v.plan.ew, c->stream));
    CKRC(dev_upload(&v.bnode, v.plan.bnode, c->stream)); CKRC(dev_upload(&v.bkind, v.plan.bkind, c->stream)); CKRC(dev_upload(&v.bw, v.plan.bw, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    v.flat_up = true;
    return MGCFD_OK;
}
// flux into v.flux (+=) for the edge classes in mask, honouring the configured flux mode
int flux_granular(mgcfd_ctx* c, int l, int mask) {
    Level& v = c->L[l];
    CKRC(ensure_flux(c, v));
    if (c->opt.flux_mode != MGCFD_FLUX_ATOMIC) {
        StageArgs a = base_args(c, v);
        a.vin = v.V(v.i_var); a.flux = v.flux; a.mask = mask;
        return launch_stage(c, v, a, false);
    }
    CKRC(ensure_flat(c, v));
    if ((mask & 1) && v.nI) {
        k_flux_atomic<<<(unsigned)blocks_for(v.nI, 256), 256, 0, c->stream>>>(v.nI, v.ea, v.eb, v.ew, v.V(v.i_var), v.npad, v.flux, 2.0 * c->kdiss);
        CKRC(post_launch(c));
    }
    if ((mask & 6) && (v.nB + v.nW)) {
        k_bflux_atomic<<<(unsigned)blocks_for(v.nB + v.nW, 256), 256, 0, c->stream>>>(v.nB + v.nW, v.bnode, v.bkind, v.bw, v.V(v.i_var), v.npad, v.flux, mask, far_field_of(c));
        CKRC(post_launch(c));
    }
    return MGCFD_OK;
}

// assess-compute variants of the flux kernel (assess_kernels.cuh): bits = REUSE_DIV | REUSE_FACTOR+FLUX << 1 | PRECOMPUTE_EDGE_WEIGHTS << 2
int launch_flux_variant(mgcfd_ctx* c, Level& v, int bits, bool refresh_planes = true) {
    // bits 0..7: the reference's toggles on 64-byte node records; 8: all three toggles on an SoA copy of the node state
    if (bits < 0 || bits > 8) { g_err = "flux variant bits must be in 0..8"; return MGCFD_ERR_ARG; }
    CKRC(ensure_flux(c, v));
    CKRC(ensure_flat(c, v));
    if (!v.nI) return MGCFD_OK;
    const unsigned nb = (unsigned)blocks_for(v.nI, 256);
    if ((bits & 4 || bits == 8) && !v.ewt_pre) {
        CK(cudaMalloc((void**)&v.ewt_pre, sizeof(double) * v.nI));
        k_edge_weights<<<nb, 256, 0, c->stream>>>(v.nI, v.ew, v.ewt_pre);
        CKRC(post_launch(c));
    }
    const double smoothing = double(0.2f);      // smoothing_coefficient, src/Base/common.h:24
    if (bits == 8) {
        if (!v.soa) { CK(cudaMalloc((void**)&v.soa, sizeof(double) * 5 * v.npad)); refresh_planes = true; }
        if (refresh_planes) { k_records_to_planes<<<(unsigned)blocks_for(v.npad, 256), 256, 0, c->stream>>>(v.npad, v.V(v.i_var), v.npad, v.soa); CKRC(post_launch(c)); }
        k_flux_assess<true, true, true, true><<<nb, 256, 0, c->stream>>>(v.nI, v.ea, v.eb, v.ew, v.ewt_pre, v.soa, v.npad, v.flux, smoothing);
        return post_launch(c);
    }
#define MG_ASSESS(D, F, P) k_flux_assess<D, F, P><<<nb, 256, 0, c->stream>>>(v.nI, v.ea, v.eb, v.ew, v.ewt_pre, v.V(v.i_var), v.npad, v.flux, smoothing)
    switch (bits) {
        case 0: MG_ASSESS(false, false, false); break; case 1: MG_ASSESS(true, false, false); break;
        case 2: MG_ASSESS(false, true, false); break;  case 3: MG_ASSESS(true, true, false); break;
        case 4: MG_ASSESS(false, false, true); break;  case 5: MG_ASSESS(true, false, true); break;
        case 6: MG_ASSESS(false, true, true); break;   default: MG_ASSESS(true, true, true); break;
    }
#undef MG_ASSESS
    return post_launch(c);
}

int step_factor(mgcfd_ctx* c, int l, int legacy) {
    Level& v = c->L[l];
    Timed tm(c, K_STEP, l, v.nel);
    const unsigned nb = (unsigned)blocks_for(v.ncomp, 256);
    if (legacy) {
        k_step_factor<true><<<nb, 256, 0, c->stream>>>(v.V(v.i_var), v.ncomp, v.vol, v.sf, c->d_minbits);
        CKRC(post_launch(c));
    } else {
        CK(cudaMemsetAsync(c->d_minbits, 0x7F, sizeof(unsigned long long), c->stream));
        k_step_factor<false><<<nb, 256, 0, c->stream>>>(v.V(v.i_var), v.ncomp, v.vol_root, v.sf, c->d_minbits);
        CKRC(post_launch(c));
        CKRC(dist_allreduce_min(c));             // global min over all ranks (cfd_loops.cpp:138-145)
        k_apply_min_dt<<<nb, 256, 0, c->stream>>>(c->d_minbits, v.vol, v.sf, v.ncomp);
        CKRC(post_launch(c));
    }
    return MGCFD_OK;
}

AllRed allred_of(mgcfd_ctx* c);
int rms_final(mgcfd_ctx* c, Level& v, bool use_counter) {
    double* out = use_counter ? c->d_rms : c->d_rms + 6 * (c->rms_cap - 1);
    int* counter = use_counter ? c->d_rms_counter : nullptr;
    if (!c->dist.active) {
        CKRC(launch_dependent(c, k_rms_final, 1u, 256u, v.rms_partial, v.rms_parts, (double)v.nel, out, counter, c->rms_cap - 1));
        return post_launch(c);
    }
    if (c->dist.p2p) {      // one kernel: local sums, all-reduce over the ranks, square roots
        CKRC(launch_dependent(c, k_rms_dist, 1u, 256u, v.rms_partial, v.rms_parts, (double)v.nel_global, out, counter, c->rms_cap - 1, allred_of(c), c->dist.d_op, c->dist.off++));
        return post_launch(c);
    }
    // distributed: local sums of squares -> all-reduce(sum) of 5 doubles -> square roots over the global node count
    k_rms_sums<<<1, 256, 0, c->stream>>>(v.rms_partial, v.rms_parts, c->d_rms_sums);
    CKRC(post_launch(c));
    CKRC(dist_allreduce_sum5(c));
    k_rms_finish<<<1, 32, 0, c->stream>>>(c->d_rms_sums, (double)v.nel_global, out, counter, c->rms_cap - 1);
    return post_launch(c);
}

// one smoothing visit (euler3d_cpu_double.cpp:383-512) on the fused path
// fused path: only the global minimum is a kernel of its own (one launch); min/volume resp. the legacy local form are
// evaluated inside the stage kernels (step_factor_of)
int min_dt_fused(mgcfd_ctx* c, int l) {
    if (c->variant == MGCFD_MESH_FVCORR) {
        // no global minimum in the legacy form; with the in-kernel data plane the stage kernels still must not read ghost rows before
        // their owners' last kernel has delivered them: an (otherwise unused) all-reduce is that synchronisation point
        if (dist_inkernel(c)) return p2p_allreduce(c, c->d_rms_sums + 7, 1, 0);
        return MGCFD_OK;
    }
    Level& v = c->L[l];
    Timed tm(c, K_STEP, l, v.nel);
    k_min_dt<<<(unsigned)blocks_for(v.ncomp, 256), 256, 0, c->stream>>>(v.V(v.i_var), v.ncomp, v.vol_root, v.blockmins, c->d_ticket, c->d_minbits);
    CKRC(post_launch(c));
    return dist_allreduce_min(c);
}

// the whole visit as ONE persistent kernel (visit_kernel.cuh): minimum dt, three stages, residual, RMS sums; multi-GPU: the halo
// exchange rides on its grid barriers
AllRed allred_of(mgcfd_ctx* c) {
    AllRed ar;
    Dist& d = c->dist;
    ar.nranks = d.nranks; ar.me = d.rank; ar.red_of_rank = d.d_red_of_rank; ar.flag_of_rank = d.d_flag_of_rank;
    ar.my_flags = (const unsigned long long*)d.win; ar.my_red = (const double*)(d.win + P2P_FLAGS_BYTES); ar.red_counter = d.d_ctr;
    return ar;
}
DistArgs dist_args(mgcfd_ctx* c, Level& v, const P2PPeer* wait_peers, int nwait) {
    DistArgs da;
    memset(&da, 0, sizeof(da));
    Dist& d = c->dist;
    da.peers = v.d_peers; da.npeers = v.npeers; da.peer_out = v.d_peer_out;
    da.tgt_off = v.d_tgt_off; da.tgt_peer = v.d_tgt_peer; da.tgt_row = v.d_tgt_row; da.tile_sends = v.d_tile_sends;
    da.wait_peers = wait_peers; da.nwait = nwait;
    da.sig_peers = d.d_sig; da.nsig = d.nsig;
    da.ar = allred_of(c);
    da.my_flags = (const unsigned long long*)d.win;
    da.op_counter = d.d_op; da.epoch_off = d.off;
    return da;
}
// one epoch per kernel: the tail of a distributed stage / transfer kernel (kernels.cuh "Epochs")
DistTail dist_tail(mgcfd_ctx* c, Level& out_level, int ib, const P2PPeer* wait_peers, int nwait) {
    DistTail t;
    memset(&t, 0, sizeof(t));
    Dist& d = c->dist;
    t.wait_peers = wait_peers; t.nwait = nwait;
    t.sig_peers = d.d_sig; t.nsig = d.nsig;
    t.peers = out_level.d_peers; t.npeers = out_level.npeers; t.peer_out = out_level.d_peer_out; t.ib = ib;
    t.tgt_off = out_level.d_tgt_off; t.tgt_peer = out_level.d_tgt_peer; t.tgt_row = out_level.d_tgt_row; t.tile_sends = out_level.d_tile_sends;
    t.order = out_level.d_order_tiles; t.n_send = out_level.n_send_tiles;
    t.op_counter = d.d_op; t.epoch_off = d.off++; t.my_flags = (const unsigned long long*)d.win;
    t.ar = allred_of(c);
    t.dbg = env_int("MGCFD_DIST_DEBUG", 0);
    return t;
}

int smooth_visit(mgcfd_ctx* c, int l) {
    Level& v = c->L[l];
    const int X = v.i_var, A = v.i_tmp, B = v.i_old;
    Timed tm(c, K_FLUX, l, MGCFD_RK * v.nI);
    VisitArgs a;
    memset(&a, 0, sizeof(a));
    a.ff = far_field_of(c);
    a.bufX = v.V(X); a.bufA = v.V(A); a.bufB = v.V(B); a.ibX = X; a.ibA = A; a.ibB = B;
    a.res = v.res; a.sf = v.sf; a.vol = v.vol; a.vol_root = v.vol_root; a.stride = v.npad;
    a.hsum = v.d_hsum; a.hs_stride = v.ncomp;
    a.legacy = (c->variant == MGCFD_MESH_FVCORR);
    a.desc = v.d_desc; a.desc_stride = v.plan.visit.desc_stride; a.max_ent = v.plan.visit.max_ent; a.hpad = v.plan.visit.hpad;
    a.vslots = v.d_vslots; a.bslots = v.bslots; a.cta_rows = v.d_cta_rows;
    a.K = v.vK; a.sr_max = v.v_srmax; a.resident = v.v_resident; a.max_chunk = v.plan.visit.max_chunk;
    if (!c->dist.active && v.premin_valid) { a.premin = v.blockmins; a.npremin = (int)blocks_for(v.ncomp, 128); }
    if (v.minword_state == 1) v.minword_state = 2;
    a.k2 = 2.0 * c->kdiss;
    a.bad_key = c->d_badkey; a.old_of_new = v.old_of_new;
    a.stage_seq0 = c->stage_seq & 0xFFFFFFull; c->stage_seq += MGCFD_RK;
    a.bar = c->d_bar; a.cta_min = c->d_cta_min; a.min_bits = c->d_minbits;
    if (l == 0) {
        a.cta_rms = c->d_cta_rms; a.rms_out = c->d_rms; a.rms_counter = c->d_rms_counter; a.rms_cap = c->rms_cap - 1;
        a.nel_global = (double)(c->dist.active ? v.nel_global : v.nel);
    }
    const bool dist = dist_inkernel(c);
    if (dist) { a.d = dist_args(c, v, v.d_peers, v.npeers); c->dist.off += 4; }      // barrier 0, two stage barriers, the end signal
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3((unsigned)v.vG); cfg.blockDim = dim3(512); cfg.dynamicSmemBytes = v.v_smem; cfg.stream = c->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr; cfg.numAttrs = c->opt.no_pdl ? 0 : 1;
    // instantiations: warps per CTA (16: one CTA per SM, 8: two), multi-GPU delivery, debug stamps
    const bool dbg = c->d_visit_dbg && !dist;
    if (dbg) a.dbg = c->d_visit_dbg;
    cfg.blockDim = dim3(32 * v.vW);
    if (v.vW == 8) {
        if (dbg) CK(cudaLaunchKernelEx(&cfg, k_visit<false, true, 2, 8>, a));
        else if (dist) CK(cudaLaunchKernelEx(&cfg, k_visit<true, false, 2, 8>, a));
        else CK(cudaLaunchKernelEx(&cfg, k_visit<false, false, 2, 8>, a));
    } else {
        if (dbg) CK(cudaLaunchKernelEx(&cfg, k_visit<false, true, 2, 16>, a));
        else if (dist) CK(cudaLaunchKernelEx(&cfg, k_visit<true, false, 2, 16>, a));
        else CK(cudaLaunchKernelEx(&cfg, k_visit<false, false, 2, 16>, a));
    }
    CKRC(post_launch(c));
    if (dist) c->dist.exchanges += MGCFD_RK;
    v.i_old = X; v.i_var = A; v.i_tmp = B;
    v.premin_valid = false;
    if (v.minword_state == 1) v.minword_state = 2;
    return MGCFD_OK;
}

int launch_stage_dist(mgcfd_ctx* c, Level& v, const StageArgs& a);

int smooth_fused(mgcfd_ctx* c, int l) {
    Level& v = c->L[l];
    if (v.visit && (!c->dist.active || c->dist.p2p)) return smooth_visit(c, l);
    const bool legacy = (c->variant == MGCFD_MESH_FVCORR);
    // multi-GPU, peer-to-peer plane, pipelined stage kernel: the stage kernels deliver their rows themselves (start-wait on the
    // level's peers, remote stores from the update, end signal): no exchange kernel between the stages
    const bool deliver = dist_inkernel(c) && v.pipe;
    // the visit's minimum dt: already all-reduced by the transfer kernel that produced this state (multi-GPU), or reduced by every
    // CTA of the first stage from the per-block minima that kernel left behind (one GPU), or -- the state came from elsewhere --
    // by k_min_dt (+ all-reduce)
    bool use_premin = false, recv_min = false, use_word = false;
    if (!legacy) {
        if (c->dist.active) { if (deliver && v.minword_state == 1) recv_min = true; else CKRC(min_dt_fused(c, l)); }
        else if (v.minword_state == 1 && v.pipe) use_word = true;  // one GPU, MGCFD_MINWORD=1: the transfer kernel folded the minimum into one word
        else if (v.premin_valid && v.pipe) use_premin = true;      // (the simple one-CTA-per-tile kernel reads *min_bits only)
        else CKRC(min_dt_fused(c, l));
    } else if (!deliver) CKRC(min_dt_fused(c, l));
    const int X = v.i_var, A = v.i_tmp, B = v.i_old;   // the previous old_variables are dead once a smooth starts
    // timing: on one GPU the three stage launches of the visit share ONE event bracket (3 * nI edge updates), so that they run as
    // in the replayed graph -- back to back, with their programmatic-dependent-launch overlap -- instead of each paying an event
    // round trip; distributed runs bracket every stage on its own (the halo exchange between stages is not flux time)
    std::unique_ptr<Timed> visit_tm;
    if (!c->dist.active) visit_tm.reset(new Timed(c, K_FLUX, l, MGCFD_RK * v.nI));
    for (int j = 0; j < MGCFD_RK; j++) {
        std::unique_ptr<Timed> stage_tm;
        if (c->dist.active) stage_tm.reset(new Timed(c, K_FLUX, l, v.nI));
        StageArgs a = base_args(c, v);
        a.vold = v.V(X);
        a.vin = (j == 0) ? v.V(X) : (j == 1 ? v.V(A) : v.V(B));
        a.vout = (j == 1) ? v.V(B) : v.V(A);
        a.rk_div = double(MGCFD_RK + 1 - j); a.rk_rcp = 1.0 / a.rk_div;
        a.mask = 7;
        a.bad_key = c->d_badkey;
        a.stage_seq = (c->stage_seq++) & 0xFFFFFFull;
        if (j == MGCFD_RK - 1) {
            a.res = v.res;
            a.rms_partial = (l == 0) ? v.rms_partial : nullptr;
        }
        if (j == 0 && use_premin) { a.premin = v.blockmins; a.npremin = (int)blocks_for(v.ncomp, 128); }
        if (use_word) { a.minword = v.d_minword; a.min_from_word = (j == 0); a.reset_word = (j == 1); }
        if (deliver) {
            const int ib = (a.vout == v.buf[0]) ? 0 : (a.vout == v.buf[1] ? 1 : 2);
            a.d = dist_tail(c, v, ib, v.d_peers, v.npeers);
            a.d.minword = v.d_minword;
            a.d.recv_min = (j == 0 && recv_min) ? 1 : 0;          // block 0 sends the level's word, every CTA picks the ranks' minima up
            a.d.finish_min = (j == 1 && recv_min) ? 1 : 0;        // the word goes back to +inf, the reduction is counted
            CKRC(launch_stage_dist(c, v, a));
            c->dist.exchanges++;
            continue;
        }
        CKRC(launch_stage(c, v, a, true));
        stage_tm.reset();
        CKRC(dist_exchange_records(c, l, a.vout));
    }
    visit_tm.reset();
    v.i_old = X; v.i_var = A; v.i_tmp = B;
    v.premin_valid = false;
    if (recv_min || use_word) v.minword_state = 0; else if (v.minword_state == 1) v.minword_state = 2;
    if (l == 0) { v.rms_parts = v.ntiles; CKRC(rms_final(c, v, true)); }
    return MGCFD_OK;
}

// a transfer kernel is about to fold its blocks' minima of dt into the level's word: the word must be +inf (it is, unless the last
// minimum was never picked up -- the state was changed through the API in between)
int minword_arm(mgcfd_ctx* c, Level& v, DistTail& t) {
    if (v.minword_state != 0) CK(cudaMemsetAsync(v.d_minword, 0x7F, sizeof(unsigned long long), c->stream));
    t.minword = v.d_minword;
    v.minword_state = 1;
    return MGCFD_OK;
}

// a transfer kernel whose blocks are all resident at once (one wave) may release the stage kernel behind it early: the stage
// CTAs then only ever take the place of transfer blocks that have exited (MGCFD_EARLY_RELEASE=0 switches it off)
inline int release_early(mgcfd_ctx* c, unsigned nb) {
    static const int on = env_int("MGCFD_EARLY_RELEASE", 1), per_sm = env_int("MGCFD_EARLY_RELEASE_BLOCKS", 12);
    return (on && !c->opt.no_pdl && nb <= (unsigned)per_sm * (unsigned)c->num_sms) ? 1 : 0;
}

// one GPU: the minimum dt the transfer kernels leave behind as ONE word per level (what the multi-GPU path does) instead of one
// value per block that every CTA of the first stage kernel reduces again (2350 loads per CTA on C2's fine level: 8 % of the stage
// kernels' stall samples, profiles/r02I_final_c2_stage_tn256_level0_stalls.txt).  Measured 0.3187 vs 0.3265 ms per C2 cycle
// (profiles/r02U_minword_one_gpu.txt), bit-identical results; MGCFD_MINWORD=0 brings the per-block minima back.
inline bool minword_one_gpu() { static const int on = env_int("MGCFD_MINWORD", 1); return on != 0; }

int do_restrict(mgcfd_ctx* c, int lc) {
    Level& vc = c->L[lc]; Level& vf = c->L[lc - 1];
    Timed tm(c, K_RESTRICT, lc, vf.nel);
    const unsigned nb = (unsigned)blocks_for(vc.ncomp, 128);
    // the per-block minima of dt of the new state (never for the legacy step factor): the next smoothing visit of the level needs
    // their minimum; multi-GPU the kernel's last CTA all-reduces it over the ranks into d_minbits
    double* bm = (c->variant != MGCFD_MESH_FVCORR && env_int("MGCFD_PREMIN", 1)) ? vc.blockmins : nullptr;
    if (dist_inkernel(c)) {
        // reads the fine level's ghost rows (wait for the fine level's peers), delivers the coarse rows itself
        DistTail t = dist_tail(c, vc, vc.i_var, vf.d_peers, vf.npeers);
        t.blk_wait = vc.d_rblk_wait; t.release_early = release_early(c, nb);
        if (vc.visit || !vc.pipe) bm = nullptr;             // only a level whose stage kernels deliver their rows themselves picks the minimum up
        if (bm) CKRC(minword_arm(c, vc, t)); else if (vc.minword_state == 1) vc.minword_state = 2;
        CKRC(launch_dependent(c, k_restrict<true>, nb, 128u, vf.V(vf.i_var), vc.V(vc.i_var), vc.ncomp, vc.child_off, vc.child_ids, vc.vol_root, nullptr, t));
        c->dist.exchanges++;
        vc.premin_valid = false;
        return post_launch(c);
    }
    DistTail none;
    memset(&none, 0, sizeof(none));
    none.release_early = release_early(c, nb);
    if (bm && vc.pipe && !vc.visit && minword_one_gpu()) { CKRC(minword_arm(c, vc, none)); bm = nullptr; }
    else if (vc.minword_state == 1) vc.minword_state = 2;
    CKRC(launch_dependent(c, k_restrict<false>, nb, 128u, vf.V(vf.i_var), vc.V(vc.i_var), vc.ncomp, vc.child_off, vc.child_ids, vc.vol_root, bm, none));
    CKRC(post_launch(c));
    vc.premin_valid = (bm != nullptr);
    return dist_exchange_records(c, lc, vc.V(vc.i_var));
}
int do_prolong(mgcfd_ctx* c, int lf) {
    Level& vf = c->L[lf]; Level& vc = c->L[lf + 1];
    Timed tm(c, K_PROLONG, lf, vf.nI);
    const unsigned nb = (unsigned)blocks_for(vf.ncomp, 128);
    double* bm = (c->variant != MGCFD_MESH_FVCORR && env_int("MGCFD_PREMIN", 1)) ? vf.blockmins : nullptr;
    if (dist_inkernel(c)) {
        // the coarse residuals of ghost parents were delivered by the coarse visit's last stage when that level runs the visit
        // kernel; otherwise they are exchanged here.  The prolonged rows are delivered by the kernel itself.
        if (!vc.visit && !vc.pipe) CKRC(dist_exchange_residuals(c, lf + 1));
        DistTail t = dist_tail(c, vf, vf.i_var, vc.d_peers, vc.npeers);
        // (when the coarse level's residuals came through an exchange kernel just above, every block may run at once)
        t.blk_wait = vf.d_pblk_wait; t.release_early = release_early(c, nb);
        if (vf.visit || !vf.pipe) bm = nullptr;
        if (bm) CKRC(minword_arm(c, vf, t)); else if (vf.minword_state == 1) vf.minword_state = 2;
        CKRC(launch_dependent(c, k_prolong<true>, nb, 128u, vf.ncomp, vf.npad, vc.npad, vf.parent, vf.idist_own, vf.ent_off, vf.ent_src, vf.ent_w,
                              vc.res, vf.res, vf.V(vf.i_var), vf.vol_root, nullptr, t));
        c->dist.exchanges++;
        vf.premin_valid = false;
        return post_launch(c);
    }
    CKRC(dist_exchange_residuals(c, lf + 1));
    DistTail none;
    memset(&none, 0, sizeof(none));
    none.release_early = release_early(c, nb);
    if (bm && vf.pipe && !vf.visit && minword_one_gpu()) { CKRC(minword_arm(c, vf, none)); bm = nullptr; }
    else if (vf.minword_state == 1) vf.minword_state = 2;
    CKRC(launch_dependent(c, k_prolong<false>, nb, 128u, vf.ncomp, vf.npad, vc.npad, vf.parent, vf.idist_own, vf.ent_off, vf.ent_src, vf.ent_w,
                          vc.res, vf.res, vf.V(vf.i_var), vf.vol_root, bm, none));
    CKRC(post_launch(c));
    vf.premin_valid = (bm != nullptr);
    return dist_exchange_records(c, lf, vf.V(vf.i_var));
}

// one full iteration of main()'s loop body sequence for a V-cycle (euler3d_cpu_double.cpp:371-694)
int cycle_fused(mgcfd_ctx* c) {
    const int nl = c->levels;
    CKRC(smooth_fused(c, 0));
    if (nl == 1) return dist_flush_epoch(c);
    for (int l = 1; l < nl; l++) { CKRC(do_restrict(c, l)); CKRC(smooth_fused(c, l)); }
    for (int l = nl - 2; l >= 1; l--) { CKRC(do_prolong(c, l)); CKRC(smooth_fused(c, l)); }
    CKRC(do_prolong(c, 0));
    return dist_flush_epoch(c);
}

std::string role_key(mgcfd_ctx* c) {
    std::string k;
    for (auto& v : c->L) { k += char('0' + v.i_var); k += char('0' + v.i_old); }
    for (auto& v : c->L) k += v.premin_valid ? 'v' : '-';      // a captured kernel has the source of its minimum dt baked in
    for (auto& v : c->L) k += char('0' + v.minword_state);
    return k;
}

int field_ptr(mgcfd_ctx* c, Level& v, int field, double** p, int* ncomp, bool for_write) {
    *ncomp = 5;
    switch (field) {
        case MGCFD_FIELD_VARIABLES: *p = v.V(v.i_var); return MGCFD_OK;
        case MGCFD_FIELD_OLD_VARIABLES: *p = v.V(v.i_old); return MGCFD_OK;
        case MGCFD_FIELD_RESIDUALS: *p = v.res; return MGCFD_OK;
        case MGCFD_FIELD_FLUXES: CKRC(ensure_flux(c, v)); *p = v.flux; return MGCFD_OK;
        case MGCFD_FIELD_STEP_FACTORS: *p = v.sf; *ncomp = 1; return MGCFD_OK;
        case MGCFD_FIELD_VOLUMES: if (for_write) break; *p = v.vol; *ncomp = 1; return MGCFD_OK;
    }
    g_err = "unknown or read-only field";
    return MGCFD_ERR_ARG;
}

int check_level(mgcfd_ctx* c, int l, bool need_final = true) {
    if (!c) { g_err = "null context"; return MGCFD_ERR_ARG; }
    if (l < 0 || l >= c->levels) { g_err = "level out of range"; return MGCFD_ERR_ARG; }
    if (need_final && !c->finalized) { g_err = "mgcfd_finalize has not been called"; return MGCFD_ERR_ARG; }
    if (need_final) CK(cudaSetDevice(c->opt.device));     // every compute entry point runs on the context's device
    return MGCFD_OK;
}

void free_level(Level& v) {
    // buf[] and res live in the context's slab
    void* ptrs[] = {v.flux, v.sf, v.vol, v.vol_root, v.new_of_old, v.old_of_new, v.hdrs,
                    v.slots, v.bslots, v.ea, v.eb, v.ew, v.bnode, v.bkind, v.bw, v.child_off, v.child_ids, v.parent,
                    v.idist_own, v.ent_off, v.ent_src, v.ent_w, v.rms_partial, v.blockmins, v.io, v.d_send_idx, v.sendbuf, v.recvtmp, v.d_peers,
                    v.d_tgt_off, v.d_tgt_peer, v.d_tgt_row, v.d_tile_sends, v.d_peer_out, v.ewt_pre, v.soa, v.d_desc, v.d_vslots, v.d_cta_rows, v.d_hsum,
                    v.d_order_tiles, v.d_rblk_wait, v.d_pblk_wait, v.d_minword};
    for (void* p : ptrs) if (p) cudaFree(p);
}


// ---- the visit kernel's configuration of a level (visit_kernel.cuh) ------------------------------------------------------------
struct VisitCfg { bool ok = false; bool roomy = false; int K = 1, G = 0, R = 1, D = 2, W = 16, resident = 0, sr_max = 0; size_t smem = 0; };
// dynamic shared memory a CTA may use: one CTA per SM (16 warps) 227 KB minus the kernel's static shared memory; two CTAs per SM
// (8 warps each) half of the SM's 228 KB minus the 1 KB the system reserves per CTA and the static part
inline size_t visit_smem_limit(int warps) { return warps == 8 ? 114000 : 230000; }
// shared memory of k_visit for a plan built with G * K super-tiles: 16 per-warp rings (D entries x R rounds x 832 bytes) + record
// buffers (resident: two own-row buffers + one halo buffer; streaming: two buffers of own + halo rows) + descriptor buffers (one
// when K == 1, else four).  Ring shapes are tried from the roomiest down; `roomy` = at least four rounds buffered per warp.
VisitCfg visit_config(const LevelPlan& P, int G, int K) {
    VisitCfg cfg;
    const VisitPlan& V = P.visit;
    if (V.ns <= 0 || V.ns != G * K) return cfg;
    const size_t own = 64 * (size_t)VT * V.maxt, halo = 64 * (size_t)V.hpad, desc = (K == 1 ? 1 : 4) * (size_t)V.desc_stride;
    const bool resident = (K == 1) && env_int("MGCFD_VISIT_RESIDENT", 1) != 0;
    const size_t recs = resident ? 2 * own + halo : 2 * (own + halo);
    const int R = V.rounds_per_chunk;                     // the descriptors' chunk lists are built for it (PlanOptions::visit_rounds)
    const int D = 2;
    const int W = V.warps;
    const size_t ring = (((size_t)W * D * R * VW * VSLOT) + 127) & ~size_t(127);
    if (R == 2 && ring + recs + desc <= visit_smem_limit(W)) {
        cfg.ok = true; cfg.roomy = true; cfg.K = K; cfg.G = G; cfg.R = R; cfg.D = D; cfg.W = W; cfg.resident = resident ? 1 : 0;
        cfg.sr_max = VT * V.maxt; cfg.smem = ring + recs + desc;
    }
    return cfg;
}
// Builds the level plan for the visit kernel: K = 1 (own rows resident) when that fits with a ring of at least two rounds per
// entry, else the smallest K whose double-buffered super-tiles fit.  Returns false (plan untouched or rebuilt by the caller) when
// the level cannot run the visit kernel.
bool plan_for_visit(int num_sms, LevelPlan& plan_out, const HostLevel& H, PlanOptions po, VisitCfg& out) {
    const long n_own = H.n_owned >= 0 ? H.n_owned : H.nel;
    const long ntiles = std::max<long>(1, (n_own + VT - 1) / VT);
    po.visit_warps = (env_int("MGCFD_VISIT_WARPS", 16) == 8) ? 8 : 16;
    const size_t VISIT_SMEM_LIMIT = visit_smem_limit(po.visit_warps);
    const int G = (int)std::min<long>((long)num_sms * (16 / po.visit_warps), ntiles);
    po.tile_nodes = VT;
    po.visit_rounds = 2;
    const int forceK = env_int("MGCFD_VISIT_K", 0);
    auto good = [&](const VisitCfg& f, const LevelPlan&) { return f.ok && f.roomy; };
    VisitCfg best; LevelPlan bestP;
    int K = forceK > 0 ? forceK : 1;
    // the own rows of a super-tile alone (double-buffered, 13-bit row index) bound K from below
    if (forceK <= 0) while ((long)G * K < ntiles && (2 * 64 * (size_t)VT * ((ntiles + (long)G * K - 1) / ((long)G * K)) > VISIT_SMEM_LIMIT * 7 / 10 ||
                                                      VT * ((ntiles + (long)G * K - 1) / ((long)G * K)) >= 8192)) K++;
    for (int tries = 0; tries < 6; tries++) {
        if ((long)G * K > ntiles) break;
        po.supers = G * K;
        LevelPlan P;
        build_level_plan(H, po, P);
        const VisitCfg f = visit_config(P, G, K);
        if (f.ok && (!best.ok || f.D * f.R > best.D * best.R)) { best = f; bestP = P; }
        if (forceK > 0 || good(f, P)) break;
        if (K == 1) {
            // estimate the smallest K that fits from this build: halo rows scale like rows^(2/3)
            const double own1 = (double)VT * P.visit.maxt, h1 = (double)P.visit.hpad;
            int k = 2;
            for (; k < 64; k++) {
                const double rows = std::ceil(own1 / k / VT) * VT + h1 * std::pow((double)k, -2.0 / 3.0);
                if (2 * 64 * rows + (double)po.visit_warps * 4 * VW * VSLOT + 4 * (128.0 * std::ceil(own1 / k / VT) + 4 * h1 * std::pow((double)k, -2.0 / 3.0) + 32) <= (double)VISIT_SMEM_LIMIT) break;
            }
            K = k;
        } else K++;
    }
    if (!best.ok) return false;
    plan_out = std::move(bestP);
    out = best;
    return true;
}
}  // namespace

extern "C" {

void mgcfd_default_options(mgcfd_options* opt) {
    memset(opt, 0, sizeof(*opt));
    opt->device = 0;
    opt->flux_mode = MGCFD_FLUX_SORTED_SEGMENT;
    opt->ordering = MGCFD_ORDER_PARTITION_RCM;
    opt->tile_nodes = 0;     // auto
    opt->use_graph = 1;
    opt->timing = 0;
}
const char* mgcfd_last_error(void) { return g_err.c_str(); }
int mgcfd_guard_selftest(mgcfd_ctx* c) {
    if (!c) { g_err = "null context"; return MGCFD_ERR_ARG; }
    bool guarded;
    { std::lock_guard<std::mutex> lk(g_guard_mu); guarded = g_guard.count(c->d_rms_sums) != 0; }
    if (!guarded) { g_err = "the context was created without MGCFD_GUARD=1"; return MGCFD_ERR_ARG; }
    CK(cudaMemset((char*)c->d_rms_sums + sizeof(double) * 8 + 100, 7, 1));
    return MGCFD_OK;
}
int mgcfd_guard_check(char* report, int cap) {
    std::string r;
    const int bad = guard_check(r);
    if (report && cap > 0) { snprintf(report, (size_t)cap, "%s", r.c_str()); }
    return bad;
}
const char* mgcfd_version(void) { return "mgcfd-b200 0.1 (sm_100a, fp64)"; }

void mgcfd_far_field_conditions(double ffv[5], double ffc[12]) {
    // initialize_far_field_conditions (src/Kernels/cfd_loops.h:85-119), same expressions in the same order
    const double gamma = 1.4, ff_mach = 1.2, deg_aoa = 0.0;
    const double angle_of_attack = double(3.1415926535897931 / 180.0) * double(deg_aoa);
    ffv[0] = double(1.4);
    const double ff_pressure = double(1.0);
    const double ff_speed_of_sound = sqrt(gamma * ff_pressure / ffv[0]);
    const double ff_speed = double(ff_mach) * ff_speed_of_sound;
    const double vel[3] = {ff_speed * double(cos(angle_of_attack)), ff_speed * double(sin(angle_of_attack)), 0.0};
    ffv[1] = ffv[0] * vel[0]; ffv[2] = ffv[0] * vel[1]; ffv[3] = ffv[0] * vel[2];
    ffv[4] = ffv[0] * (double(0.5) * (ff_speed * ff_speed)) + (ff_pressure / double(gamma - 1.0));
    const double mom[3] = {ffv[1], ffv[2], ffv[3]};
    // compute_flux_contribution (cfd_loops.h:57-83)
    ffc[0] = vel[0] * mom[0] + ff_pressure; ffc[1] = vel[0] * mom[1]; ffc[2] = vel[0] * mom[2];
    ffc[3] = ffc[1]; ffc[4] = vel[1] * mom[1] + ff_pressure; ffc[5] = vel[1] * mom[2];
    ffc[6] = ffc[2]; ffc[7] = ffc[5]; ffc[8] = vel[2] * mom[2] + ff_pressure;
    const double de_p = ffv[4] + ff_pressure;
    ffc[9] = vel[0] * de_p; ffc[10] = vel[1] * de_p; ffc[11] = vel[2] * de_p;
}

int mgcfd_adjust_dampen_ewt(int mesh_variant, const double* coords_xyz, long num_edges, void* edges_aos40) {
    if (!edges_aos40 || (mesh_variant != MGCFD_MESH_FVCORR && !coords_xyz)) { g_err = "null argument"; return MGCFD_ERR_ARG; }
    apply_ewt(mesh_variant, coords_xyz, num_edges, (EdgeNb*)edges_aos40);
    return MGCFD_OK;
}

int mgcfd_create(int levels, int mesh_variant, const mgcfd_options* opt, mgcfd_ctx** out) {
    if (!out || levels < 1 || levels > 8) { g_err = "bad arguments"; return MGCFD_ERR_ARG; }
    *out = nullptr;
    mgcfd_options o;
    if (opt) o = *opt; else mgcfd_default_options(&o);
    if (o.tile_nodes != 0 && o.tile_nodes != 128 && o.tile_nodes != 256 && o.tile_nodes != 512) { g_err = "tile_nodes must be 0 (auto), 128, 256 or 512"; return MGCFD_ERR_ARG; }
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev <= 0 || o.device >= ndev) {
        g_err = std::string("no usable CUDA device (") + cudaGetErrorString(e) + "); libmgcfd_b200 has no CPU fallback";
        return MGCFD_ERR_NO_DEVICE;
    }
    CK(cudaSetDevice(o.device));
    mgcfd_ctx* c = new mgcfd_ctx();
    cudaDeviceGetAttribute(&c->num_sms, cudaDevAttrMultiProcessorCount, o.device);
    c->opt = o; c->levels = levels; c->variant = mesh_variant;
    c->L.resize(levels);
    c->kdiss = -0.5 * double(0.2f);
    c->t_ms.assign(K_COUNT * levels, 0.0);
    c->t_iters.assign(K_COUNT * levels, 0);
    CK(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
    CK(cudaEventCreate(&c->ev0)); CK(cudaEventCreate(&c->ev1));
    CK(cudaMalloc((void**)&c->d_minbits, 8)); CK(cudaMalloc((void**)&c->d_badkey, 8));
    CK(cudaMalloc((void**)&c->d_ticket, 4)); CK(cudaMemset(c->d_ticket, 0, 4));
    { const double one = 1.0; CK(cudaMemcpy(c->d_minbits, &one, 8, cudaMemcpyHostToDevice)); }   // a finite value until the first k_min_dt
    CK(cudaMemset(c->d_badkey, 0xFF, 8));
    c->rms_cap = 4096 + 1;
    CK(cudaMalloc((void**)&c->d_rms, sizeof(double) * 6 * c->rms_cap));
    CK(cudaMalloc((void**)&c->d_rms_sums, sizeof(double) * 8));
    CK(cudaMemset(c->d_rms_sums, 0, sizeof(double) * 8));
    CK(cudaMalloc((void**)&c->d_bar, sizeof(unsigned int))); CK(cudaMemset(c->d_bar, 0, sizeof(unsigned int)));
    CK(cudaMalloc((void**)&c->d_cta_min, sizeof(double) * 2 * c->num_sms));
    CK(cudaMalloc((void**)&c->d_cta_rms, sizeof(double) * 5 * 2 * c->num_sms));
    if (env_int("MGCFD_VISIT_DEBUG", 0)) { CK(cudaMalloc((void**)&c->d_visit_dbg, sizeof(long long) * 64 * 2 * c->num_sms)); CK(cudaMemset(c->d_visit_dbg, 0, sizeof(long long) * 64 * 2 * c->num_sms)); }
    CK(cudaMalloc((void**)&c->d_rms_counter, sizeof(int)));
    CK(cudaMemset(c->d_rms_counter, 0, sizeof(int)));
    mgcfd_far_field_conditions(c->ff, c->ffc);
    *out = c;
    return MGCFD_OK;
}

int mgcfd_destroy(mgcfd_ctx* c) {
    if (!c) return MGCFD_OK;
    cudaSetDevice(c->opt.device);
    cudaStreamSynchronize(c->stream);
    for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);
    for (auto& v : c->L) free_level(v);
    cudaFree(c->d_minbits); cudaFree(c->d_badkey); cudaFree(c->d_ticket); cudaFree(c->d_rms); cudaFree(c->d_rms_counter); cudaFree(c->d_rms_sums);
    if (c->dist.comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->dist.comm);
    for (int p = 0; p < (int)c->dist.peer_win.size(); p++) if (p != c->dist.rank && c->dist.peer_win[p]) cudaIpcCloseMemHandle(c->dist.peer_win[p]);
    cudaFree(c->d_visit_dbg); if (c->slab) guard_gaps_drop((char*)c->slab, c->slab_bytes); (cudaFree)(c->slab); cudaFree(c->d_bar); cudaFree(c->d_cta_min); cudaFree(c->d_cta_rms);
    cudaFree(c->dist.d_ticket); cudaFree(c->dist.d_sig); cudaFree(c->dist.d_op); cudaFree(c->dist.d_ctr); cudaFree(c->dist.d_red_of_rank); cudaFree(c->dist.d_flag_of_rank);
    cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1);
    for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
    cudaStreamDestroy(c->stream);
    delete c;
    return MGCFD_OK;
}

int mgcfd_set_farfield(mgcfd_ctx* c, const double ffv[5], const double ffc[12]) {
    if (!c || !ffv || !ffc) { g_err = "null argument"; return MGCFD_ERR_ARG; }
    memcpy(c->ff, ffv, sizeof(c->ff)); memcpy(c->ffc, ffc, sizeof(c->ffc));
    for (auto& kv : c->graphs) cudaGraphExecDestroy(kv.second);      // captured launches carry the far field in their arguments
    c->graphs.clear(); c->graph_launches.clear(); c->graph_flags_end.clear();
    return MGCFD_OK;
}

int mgcfd_upload_level(mgcfd_ctx* c, int l, long nel, const double* volumes, const double* coords, long nI, long nB, long nW,
                       const void* edges, const long* mg_map, long mgc) {
    CKRC(check_level(c, l, false));
    if (c->finalized) { g_err = "context already finalized"; return MGCFD_ERR_ARG; }
    if (nel <= 0 || !volumes || !edges || nI < 0 || nB < 0 || nW < 0) { g_err = "bad mesh arguments"; return MGCFD_ERR_ARG; }
    if (c->levels > 1 && !coords) { g_err = "coords are required when levels > 1"; return MGCFD_ERR_ARG; }
    if (l < c->levels - 1 && (!mg_map || mgc != nel)) { g_err = "mg_map of size nel is required on every level but the coarsest"; return MGCFD_ERR_ARG; }
    Level& v = c->L[l];
    HostLevel& H = v.host;
    H.nel = nel; H.nI = nI; H.nB = nB; H.nW = nW;
    H.volumes.assign(volumes, volumes + nel);
    const EdgeNb* e = (const EdgeNb*)edges;
    H.edges.assign(e, e + nI + nB + nW);
    for (long k = 0; k < nI + nB + nW; k++) {
        const bool internal = k < nI;
        if (H.edges[k].b < 0 || H.edges[k].b >= nel || (internal && (H.edges[k].a < 0 || H.edges[k].a >= nel))) {
            g_err = "edge endpoint out of range"; return MGCFD_ERR_ARG;
        }
    }
    if (coords) H.coords.assign(coords, coords + 3 * nel); else H.coords.clear();
    if (mg_map && l < c->levels - 1) H.mg.assign(mg_map, mg_map + mgc); else H.mg.clear();
    PlanOptions po; po.ordering = c->opt.ordering; po.scatter = (c->opt.flux_mode == MGCFD_FLUX_TILED_COLOURED);
    po.strict = (c->opt.flux_mode != MGCFD_FLUX_ATOMIC);     // the atomic baseline runs on any numbering, tiled or not
    po.tile_nodes = c->opt.tile_nodes ? c->opt.tile_nodes : auto_tile_nodes(H.n_owned >= 0 ? H.n_owned : nel, nI, c->num_sms);
    // on request (mgcfd_options::visit, MGCFD_VISIT=1) levels up to ~1.2 M nodes run the persistent visit kernel (one launch per
    // smoothing visit, visit_kernel.cuh): sorted-segment mode, the partitioning order, 128-node tiles grouped into super-tiles
    v.visit = false;
    const long n_own = H.n_owned >= 0 ? H.n_owned : nel;
    const bool want_visit = env_int("MGCFD_VISIT", c->opt.visit) != 0 && c->opt.flux_mode == MGCFD_FLUX_SORTED_SEGMENT &&
                            c->opt.ordering == MGCFD_ORDER_PARTITION_RCM && (c->opt.tile_nodes == 0 || c->opt.tile_nodes == VT) &&
                            !c->opt.no_pipeline && n_own <= (long)env_int("MGCFD_VISIT_MAX_NODES", 1200000);
    try {
        VisitCfg cfg;
        if (want_visit && plan_for_visit(c->num_sms, v.plan, H, po, cfg)) {
            v.visit = true; v.vK = cfg.K; v.vG = cfg.G; v.vR = cfg.R; v.vD = cfg.D; v.vW = cfg.W; v.v_resident = cfg.resident; v.v_srmax = cfg.sr_max; v.v_smem = cfg.smem;
        } else build_level_plan(H, po, v.plan);
    }
    catch (const std::exception& ex) { g_err = ex.what(); return MGCFD_ERR_ARG; }
    v.uploaded = true;
    return MGCFD_OK;
}

int mgcfd_finalize(mgcfd_ctx* c) {
    if (!c) { g_err = "null context"; return MGCFD_ERR_ARG; }
    if (c->finalized) return MGCFD_OK;
    for (auto& v : c->L) if (!v.uploaded) { g_err = "not every level has been uploaded"; return MGCFD_ERR_ARG; }
    CK(cudaSetDevice(c->opt.device));
    cudaStream_t s = c->stream;
    {   // one slab: [peer-to-peer window |] three record buffers + residual planes per level (what other ranks write into)
        auto up = [](size_t x) { return (x + 255) & ~size_t(255); };
        Dist& d = c->dist;
        size_t off = 0;
        if (d.active) {
            d.stage_off.assign(c->levels, 0);
            size_t doubles = 0;
            for (int l = 0; l < c->levels; l++) { const LevelPlan& P = c->L[l].plan; d.stage_off[l] = (long)doubles; doubles += 2 * 8 * (size_t)std::max<long>(P.nel - P.n_owned, 1); }
            d.win_bytes = P2P_HDR_BYTES + doubles * sizeof(double);
            off = up(d.win_bytes);
        }
        d.buf_off.assign(4 * (size_t)c->levels, 0);
        for (int l = 0; l < c->levels; l++) {
            const size_t npad = (size_t)c->L[l].plan.npad;
            const size_t gap = guard_on() ? GUARD_BYTES : 0;          // MGCFD_GUARD=1: a checked zone after every sub-buffer
            for (int b = 0; b < 3; b++) { d.buf_off[4 * l + b] = off; off += up(sizeof(double) * 8 * npad) + gap; }
            d.buf_off[4 * l + 3] = off; off += up(sizeof(double) * 5 * npad) + gap;
        }
        c->slab_bytes = off;
        CK((cudaMalloc)((void**)&c->slab, c->slab_bytes));      // (no guard zones: the address travels as an IPC handle)
        if (guard_on()) CK(cudaMemsetAsync(c->slab, 0xA5, c->slab_bytes, s));
        if (d.active) { d.win = c->slab; CK(cudaMemsetAsync(d.win, 0, up(d.win_bytes), s)); }
        for (int l = 0; guard_on() && l < c->levels; l++) {
            const size_t npad = (size_t)c->L[l].plan.npad;
            for (int b = 0; b < 3; b++) guard_gap_add((char*)c->slab + d.buf_off[4 * l + b] + up(sizeof(double) * 8 * npad), __LINE__);
            guard_gap_add((char*)c->slab + d.buf_off[4 * l + 3] + up(sizeof(double) * 5 * npad), __LINE__);
        }
        for (int l = 0; l < c->levels; l++) {
            for (int b = 0; b < 3; b++) c->L[l].buf[b] = (double*)(c->slab + d.buf_off[4 * l + b]);
            c->L[l].res = (double*)(c->slab + d.buf_off[4 * l + 3]);
        }
    }
    for (int l = 0; l < c->levels; l++) {
        Level& v = c->L[l];
        LevelPlan& P = v.plan;
        v.nel = P.nel; v.npad = P.npad; v.ncomp = P.npad_owned; v.ntiles = P.ntiles; v.nI = P.nI; v.nB = P.nB; v.nW = P.nW; v.TN = P.TN;
        v.smem_nodes = P.TN + P.hpad;
        v.smem_bytes = 64 * (size_t)v.smem_nodes + (P.scatter ? 40 * (size_t)P.TN : 0);
        // pipelined kernel: a ring of RING entries of chunk_rounds round blocks + two record buffers + three header buffers.
        // chunk_rounds as large as possible (fewer CTA-wide hand-overs), at most half the rounds of a tile rounded up (so that
        // one entry is being filled while the other is consumed), while `want_ctas` CTAs still fit the SM's 228 KB
        {
            const size_t fixed = 2 * 64 * (size_t)v.smem_nodes + 3 * (size_t)P.hdr_stride + (P.scatter ? 40 * (size_t)P.TN : 0);
            // 128-node tiles run 3 CTAs per SM when that does not shrink the ring below half a tile's rounds per entry (hex-dual
            // meshes: 2761 vs 2580 V-cycles/s on C2); with many rounds per tile (tets) larger ring entries at 2 CTAs win (237 vs 246 us)
            int want_ctas = P.TN <= 256 ? 2 : 1;
            if (P.TN <= 128 && fixed + (size_t)RING * ((P.max_rounds + 1) / 2) * P.TN * 26 <= (228 * 1024) / 3 - 1024 - 2048) want_ctas = 3;
            const size_t budget = (228 * 1024) / want_ctas - 1024 - 2048;
            int R = std::max(1, P.max_rounds > 8 ? (P.max_rounds + 1) / 2 : P.max_rounds);
            while (R > 1 && fixed + (size_t)RING * R * P.TN * 26 > budget) R--;
            const int nchunks = std::max(1, (P.max_rounds + R - 1) / R);
            v.chunk_rounds = std::max(1, (P.max_rounds + nchunks - 1) / nchunks);
            v.pipe_smem = fixed + (size_t)RING * v.chunk_rounds * P.TN * 26;
        }
        v.untiled = P.oversize || v.smem_bytes > 227 * 1024;
        if (v.untiled && c->opt.flux_mode != MGCFD_FLUX_ATOMIC) { g_err = "tile halo too large for shared memory; use a smaller tile_nodes or a locality-preserving ordering"; return MGCFD_ERR_ARG; }
        CK(cudaMalloc((void**)&v.sf, sizeof(double) * v.npad));
        // volumes in new order; padding: volume 1, root +inf so that padded nodes never win the min-dt reduction
        std::vector<double> vol(v.npad, 1.0), root(v.npad, INFINITY);
        for (long i = 0; i < v.nel; i++) {
            const long g = P.new_of_old[i];
            vol[g] = v.host.volumes[i];
            root[g] = cbrt(v.host.volumes[i]);   // host libm cbrt, as the reference (cfd_loops.cpp:123); cbrt is not correctly rounded anywhere
        }
        CKRC(dev_upload(&v.vol, vol, s)); CKRC(dev_upload(&v.vol_root, root, s));
        std::vector<int> n2o(P.old_of_new.begin(), P.old_of_new.end()), o2n(P.new_of_old.begin(), P.new_of_old.end());
        CKRC(dev_upload(&v.old_of_new, n2o, s)); CKRC(dev_upload(&v.new_of_old, o2n, s));
        CKRC(dev_upload(&v.hdrs, P.hdrs, s)); CKRC(dev_upload(&v.slots, P.slots, s)); CKRC(dev_upload(&v.bslots, P.bslots, s));
        v.pipe = false;
        if (!c->opt.no_pipeline && !v.untiled) CKRC(setup_pipe(c, v));
        if (v.visit) {
            // the visit kernel's streams; its grid must be co-resident (one CTA per SM)
            const VisitPlan& V = P.visit;
            CKRC(dev_upload(&v.d_desc, V.desc, s)); CKRC(dev_upload(&v.d_vslots, V.vslots, s)); CKRC(dev_upload(&v.d_hsum, V.hsum, s));
            std::vector<int> rows(v.vG + 1);
            for (int g = 0; g <= v.vG; g++) rows[g] = int(V.super_off[(size_t)std::min<long>((long)g * v.vK, V.ns)] * VT);
            CKRC(dev_upload(&v.d_cta_rows, rows, s));
            static size_t attr_bytes[2] = {0, 0};
            const int wi = (v.vW == 8) ? 1 : 0;
            if (v.v_smem > attr_bytes[wi]) {
                const int sm = (int)v.v_smem;
                if (wi) {
                    CK(cudaFuncSetAttribute(k_visit<false, false, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
                    CK(cudaFuncSetAttribute(k_visit<true, false, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
                    CK(cudaFuncSetAttribute(k_visit<false, true, 2, 8>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
                } else {
                    CK(cudaFuncSetAttribute(k_visit<false, false, 2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
                    CK(cudaFuncSetAttribute(k_visit<true, false, 2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
                    CK(cudaFuncSetAttribute(k_visit<false, true, 2, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, sm));
                }
                attr_bytes[wi] = v.v_smem;
            }
            int per_sm = 0;
            if (wi) CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_visit<true, false, 2, 8>, 256, v.v_smem));
            else CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_visit<true, false, 2, 16>, 512, v.v_smem));
            if (per_sm < 16 / v.vW || v.vG > per_sm * c->num_sms) { g_err = "the visit kernel does not fit an SM with this level's configuration"; return MGCFD_ERR_ARG; }
        }
        if (v.nel_global == 0) v.nel_global = v.nel;
        v.n_owned = P.n_owned;
        if (c->dist.active) {
            // halo exchange lists in device numbering
            v.nsend = (long)v.send_list.size(); v.nghost = v.nel - v.n_owned;
            std::vector<int> sidx(v.nsend);
            for (long k = 0; k < v.nsend; k++) sidx[k] = (int)P.new_of_old[v.send_list[k]];
            CKRC(dev_upload(&v.d_send_idx, sidx, s));
            v.h_send_idx = sidx;
            CK(cudaMalloc((void**)&v.sendbuf, sizeof(double) * 8 * std::max<long>(v.nsend, 1)));
            CK(cudaMalloc((void**)&v.recvtmp, sizeof(double) * 5 * std::max<long>(v.nghost, 1)));
        }
        const long parts = std::max<long>(v.ntiles, blocks_for(v.npad, 256));
        CK(cudaMalloc((void**)&v.rms_partial, sizeof(double) * 5 * parts));
        // per-block minima of dt: k_min_dt writes one per 256 rows, the transfer kernels one per 128 rows whatever the tile size
        CK(cudaMalloc((void**)&v.blockmins, sizeof(double) * std::max<long>(parts, blocks_for(v.npad, 128))));
        CK(cudaMalloc((void**)&v.d_minword, sizeof(unsigned long long))); CK(cudaMemsetAsync(v.d_minword, 0x7F, sizeof(unsigned long long), s));
        CK(cudaStreamSynchronize(s));
        // node state: every buffer starts at the far-field state (what initialize_variables leaves, cfd_loops.h:44-55);
        // padding nodes keep it forever (no edges, zero residual), which keeps them finite in every stage
        for (int b = 0; b < 3; b++) { k_fill_state<<<(unsigned)blocks_for(v.npad, 256), 256, 0, s>>>(v.buf[b], v.npad, far_field_of(c)); CKRC(post_launch(c)); }
        CK(cudaMemsetAsync(v.res, 0, sizeof(double) * 5 * v.npad, s));
        CK(cudaMemsetAsync(v.sf, 0, sizeof(double) * v.npad, s));
    }
    // multigrid operators
    for (int l = 0; l + 1 < c->levels; l++) {
        Level& vf = c->L[l]; Level& vc = c->L[l + 1];
        for (long i = 0; i < vf.nel; i++)
            if (vf.host.mg[i] >= vc.nel || (vf.host.mg[i] < 0 && !c->dist.active)) { g_err = "mg_map entry out of range"; return MGCFD_ERR_ARG; }
        TransferPlan T;
        try { build_transfer_plan(vf.host, vc.host, vf.plan, vc.plan, T); }
        catch (const std::exception& ex) { g_err = ex.what(); return MGCFD_ERR_ARG; }
        CKRC(dev_upload(&vc.child_off, T.child_off, s)); CKRC(dev_upload(&vc.child_ids, T.child_ids, s));
        CKRC(dev_upload(&vf.parent, T.parent, s)); CKRC(dev_upload(&vf.idist_own, T.idist_own, s));
        CKRC(dev_upload(&vf.ent_off, T.ent_off, s)); CKRC(dev_upload(&vf.ent_src, T.ent_src, s)); CKRC(dev_upload(&vf.ent_w, T.ent_w, s));
        if (c->dist.active) {
            // which 128-row blocks of the transfer kernels read rows other ranks own: only those wait for the owners' epoch
            std::vector<unsigned char> rw(std::max<long>(1, blocks_for(vc.ncomp, 128)), 0), pw(std::max<long>(1, blocks_for(vf.ncomp, 128)), 0);
            for (long cc = 0; cc < vc.ncomp; cc++)
                for (long k = T.child_off[cc]; k < T.child_off[cc + 1]; k++) if (T.child_ids[k] >= vf.ncomp) { rw[cc / 128] = 1; break; }
            for (long i = 0; i < vf.ncomp; i++) {
                bool g = T.parent[i] >= vc.ncomp;
                for (long k = T.ent_off[i]; k < T.ent_off[i + 1] && !g; k++) g = T.ent_src[k] >= vc.ncomp;
                if (g) pw[i / 128] = 1;
            }
            CKRC(dev_upload(&vc.d_rblk_wait, rw, s)); CKRC(dev_upload(&vf.d_pblk_wait, pw, s));
        }
        CK(cudaStreamSynchronize(s));
    }
    CK(cudaStreamSynchronize(s));
    // host copies are no longer needed, except the plan pieces used lazily (flat/CSR) and by introspection
    for (auto& v : c->L) {
        v.host = HostLevel();
        v.plan.visit.vslots.clear(); v.plan.visit.vslots.shrink_to_fit();
        v.plan.visit.hsum.clear(); v.plan.visit.hsum.shrink_to_fit();
        if (v.npad > 2000000) {   // big levels: drop the host copy of the edge stream (introspection needs it only on small meshes)
            v.plan.slots.clear(); v.plan.slots.shrink_to_fit();
            v.plan.bslots.clear(); v.plan.bslots.shrink_to_fit();
        }
    }
    c->finalized = true;
    return MGCFD_OK;
}

// ---- granular path -------------------------------------------------------------------------------------
int mgcfd_initialize_variables(mgcfd_ctx* c, int l) {
    CKRC(check_level(c, l));
    Level& v = c->L[l];
    v.premin_valid = false;
    if (v.minword_state == 1) v.minword_state = 2;
    k_fill_state<<<(unsigned)blocks_for(v.npad, 256), 256, 0, c->stream>>>(v.V(v.i_var), v.npad, far_field_of(c));
    return post_launch(c);
}
int mgcfd_copy_old_variables(mgcfd_ctx* c, int l) {
    CKRC(check_level(c, l));
    Level& v = c->L[l];
    k_copy<<<(unsigned)blocks_for(8 * v.npad, 256), 256, 0, c->stream>>>(v.V(v.i_old), v.V(v.i_var), 8 * v.npad);
    return post_launch(c);
}
int mgcfd_compute_step_factor(mgcfd_ctx* c, int l, int legacy) {
    CKRC(check_level(c, l));
    return step_factor(c, l, legacy);
}
int mgcfd_compute_flux_edge(mgcfd_ctx* c, int l) {
    CKRC(check_level(c, l));
    Timed tm(c, K_FLUX, l, c->L[l].nI);
    return flux_granular(c, l, 1);
}
int mgcfd_compute_boundary_flux_edge(mgcfd_ctx* c, int l) { CKRC(check_level(c, l)); return flux_granular(c, l, 2); }
int mgcfd_compute_wall_flux_edge(mgcfd_ctx* c, int l) { CKRC(check_level(c, l)); return flux_granular(c, l, 4); }
int mgcfd_time_step(mgcfd_ctx* c, int l, int j) {
    CKRC(check_level(c, l));
    if (j < 0 || j >= MGCFD_RK) { g_err = "rk_stage out of range"; return MGCFD_ERR_ARG; }
    Level& v = c->L[l];
    CKRC(ensure_flux(c, v));
    v.premin_valid = false;
    if (v.minword_state == 1) v.minword_state = 2;
    Timed tm(c, K_TIME, l, v.nel);
    k_time_step<<<(unsigned)blocks_for(v.ncomp, 256), 256, 0, c->stream>>>(double(MGCFD_RK + 1 - j), v.ncomp, v.npad, v.sf, v.flux, v.V(v.i_old), v.V(v.i_var));
    return post_launch(c);
}
int mgcfd_zero_fluxes(mgcfd_ctx* c, int l) {
    CKRC(check_level(c, l));
    Level& v = c->L[l];
    CKRC(ensure_flux(c, v));
    k_fill<<<(unsigned)blocks_for(5 * v.npad, 256), 256, 0, c->stream>>>(v.flux, 5 * v.npad, 0.0);
    return post_launch(c);
}
int mgcfd_indirect_rw(mgcfd_ctx* c, int l) {
    CKRC(check_level(c, l));
    Level& v = c->L[l];
    CKRC(ensure_flux(c, v)); CKRC(ensure_flat(c, v));
    Timed tm(c, K_INDIRECT, l, v.nI);
    if (v.nI) { k_indirect_rw<<<(unsigned)blocks_for(v.nI, 256), 256, 0, c->stream>>>(v.nI, v.ea, v.eb, v.ew, v.V(v.i_var), v.npad, v.flux); CKRC(post_launch(c)); }
    return MGCFD_OK;
}
int mgcfd_residual(mgcfd_ctx* c, int l) {
    CKRC(check_level(c, l));
    Level& v = c->L[l];
    k_residual<<<(unsigned)blocks_for(v.ncomp, 256), 256, 0, c->stream>>>(v.ncomp, v.npad, v.V(v.i_old), v.V(v.i_var), v.res);
    return post_launch(c);
}
int mgcfd_calc_rms(mgcfd_ctx* c, int l, double* rms_all, double rms_var[5]) {
    CKRC(check_level(c, l));
    Level& v = c->L[l];
    const long nb = blocks_for(v.ncomp, 256);
    k_rms_partial<<<(unsigned)nb, 256, 0, c->stream>>>(v.res, v.npad, v.ncomp, v.rms_partial);
    CKRC(post_launch(c));
    v.rms_parts = nb;
    CKRC(rms_final(c, v, false));
    double out[6];
    CK(cudaMemcpyAsync(out, c->d_rms + 6 * (c->rms_cap - 1), sizeof(out), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (rms_all) *rms_all = out[0];
    if (rms_var) for (int k = 0; k < 5; k++) rms_var[k] = out[1 + k];
    return MGCFD_OK;
}
int mgcfd_check_for_invalid_variables(mgcfd_ctx* c, int l, long* first_bad_cell, int* reason) {
    CKRC(check_level(c, l));
    Level& v = c->L[l];
    unsigned long long* key = c->d_minbits;   // scratch word; the step-factor kernel re-initialises it before use
    CK(cudaMemsetAsync(key, 0xFF, 8, c->stream));
    k_check_invalid<<<(unsigned)blocks_for(v.ncomp, 256), 256, 0, c->stream>>>(v.V(v.i_var), v.ncomp, v.old_of_new, key);
    CKRC(post_launch(c));
    unsigned long long h = 0;
    CK(cudaMemcpyAsync(&h, key, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    if (h == ~0ull) { if (first_bad_cell) *first_bad_cell = -1; if (reason) *reason = 0; return MGCFD_OK; }
    if (first_bad_cell) *first_bad_cell = (long)(h >> 2);
    if (reason) *reason = (int)(h & 3);
    g_err = "invalid variables detected";
    return MGCFD_ERR_INVALID_VARIABLES;
}
int mgcfd_mg_restrict(mgcfd_ctx* c, int lc) {
    CKRC(check_level(c, lc));
    if (lc < 1) { g_err = "coarse_level must be >= 1"; return MGCFD_ERR_ARG; }
    CKRC(do_restrict(c, lc));
    return dist_flush_epoch(c);
}
int mgcfd_prolong(mgcfd_ctx* c, int lf) {
    CKRC(check_level(c, lf));
    if (lf >= c->levels - 1) { g_err = "fine_level must have a coarser level"; return MGCFD_ERR_ARG; }
    CKRC(do_prolong(c, lf));
    return dist_flush_epoch(c);
}

// ---- fused path ----------------------------------------------------------------------------------------
namespace {
void advance_roles_one_cycle(mgcfd_ctx* c) {
    // exactly as cycle_fused does: one smooth on levels 0 and L-1, two on the others
    for (int l = 0; l < c->levels; l++) {
        const int visits = (c->levels == 1 || l == 0 || l == c->levels - 1) ? 1 : 2;
        for (int k = 0; k < visits; k++) { Level& v = c->L[l]; const int X = v.i_var, A = v.i_tmp, B = v.i_old; v.i_old = X; v.i_var = A; v.i_tmp = B; }
    }
}
// enqueues one V-cycle on the context's stream (a graph replay when enabled); no host synchronisation
int enqueue_one_cycle(mgcfd_ctx* c) {
    const bool graph = c->opt.use_graph && !c->opt.timing;
    if (!graph) return cycle_fused(c);
    const std::string key = role_key(c);
    auto it = c->graphs.find(key);
    if (it == c->graphs.end()) {
        cudaGraph_t g = nullptr;
        const long launches_before = c->launches;
        c->capturing = true;
        std::vector<char> flags0, mw0;
        for (auto& v : c->L) { flags0.push_back(v.premin_valid); mw0.push_back((char)v.minword_state); }
        CK(cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
        int rc = cycle_fused(c);
        cudaError_t ce = cudaStreamEndCapture(c->stream, &g);
        c->capturing = false;
        c->graph_launches[key] = c->launches - launches_before;
        {   // what the cycle leaves behind of the minimum-dt bookkeeping; rolled back like the buffer roles, re-applied by every replay
            std::string end;
            for (size_t l = 0; l < c->L.size(); l++) { end += c->L[l].premin_valid ? 'v' : '-'; c->L[l].premin_valid = flags0[l] != 0; }
            for (size_t l = 0; l < c->L.size(); l++) { end += char('0' + c->L[l].minword_state); c->L[l].minword_state = mw0[l]; }
            c->graph_flags_end[key] = end;
        }
        c->launches = launches_before;        // capture recorded the launches, it did not run them
        if (rc != MGCFD_OK) { if (g) cudaGraphDestroy(g); return rc; }
        CK(ce);
        cudaGraphExec_t ge = nullptr;
        CK(cudaGraphInstantiate(&ge, g, 0));
        CK(cudaGraphDestroy(g));
        c->graphs[key] = ge;
        // capture advanced the buffer roles: roll them back, the replay below advances them for real
        for (size_t l = 0; l < c->L.size(); l++) { c->L[l].i_var = key[2 * l] - '0'; c->L[l].i_old = key[2 * l + 1] - '0'; c->L[l].i_tmp = 3 - c->L[l].i_var - c->L[l].i_old; }
        it = c->graphs.find(key);
    }
    CK(cudaGraphLaunch(it->second, c->stream));
    advance_roles_one_cycle(c);
    c->launches += c->graph_launches[key];
    {
        const std::string& end = c->graph_flags_end[key];
        for (size_t l = 0; l < c->L.size() && l < end.size(); l++) c->L[l].premin_valid = (end[l] == 'v');
        for (size_t l = 0; l < c->L.size() && c->L.size() + l < end.size(); l++) c->L[l].minword_state = end[c->L.size() + l] - '0';
    }
    return MGCFD_OK;
}
}  // namespace

int mgcfd_enqueue_cycles(mgcfd_ctx* c, int ncycles) {
    if (!c || !c->finalized) { g_err = "context not finalized"; return MGCFD_ERR_ARG; }
    if (ncycles < 0 || c->rms_pending + ncycles > c->rms_cap - 1) { g_err = "too many cycles enqueued without mgcfd_collect (limit 4096)"; return MGCFD_ERR_ARG; }
    CK(cudaSetDevice(c->opt.device));
    for (int i = 0; i < ncycles; i++) CKRC(enqueue_one_cycle(c));
    c->rms_pending += ncycles;
    return MGCFD_OK;
}

int mgcfd_collect(mgcfd_ctx* c, double* rms_all, double* rms_var) {
    if (!c || !c->finalized) { g_err = "context not finalized"; return MGCFD_ERR_ARG; }
    const int n = c->rms_pending;
    std::vector<double> h(6 * (size_t)std::max(n, 1));
    unsigned long long key = 0;
    if (n) CK(cudaMemcpyAsync(h.data(), c->d_rms, sizeof(double) * 6 * n, cudaMemcpyDeviceToHost, c->stream));
    if (c->dist.active && c->dist.p2p) CKRC(p2p_allreduce(c, (double*)c->d_badkey, 1, 1));
    else if (c->dist.active) NK(g_nccl.AllReduce(c->d_badkey, c->d_badkey, 1, ncclUint64, ncclMin, c->dist.comm, c->stream));   // every rank learns of an invalid state anywhere
    CK(cudaMemcpyAsync(&key, c->d_badkey, 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemsetAsync(c->d_rms_counter, 0, sizeof(int), c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->rms_pending = 0;
    for (int i = 0; i < n; i++) {
        if (rms_all) rms_all[i] = h[6 * i];
        if (rms_var) for (int k = 0; k < 5; k++) rms_var[i * 5 + k] = h[6 * i + 1 + k];
    }
    if (key != ~0ull) {
        c->bad_cell = (long)((key >> 2) & 0x3FFFFFFFFFull); c->bad_reason = (int)(key & 3);
        g_err = "invalid variables detected (cell " + std::to_string(c->bad_cell) + ")";
        return MGCFD_ERR_INVALID_VARIABLES;
    }
    return MGCFD_OK;
}

int mgcfd_run_cycles(mgcfd_ctx* c, int ncycles, double* rms_all, double* rms_var) {
    if (!c || !c->finalized) { g_err = "context not finalized"; return MGCFD_ERR_ARG; }
    if (ncycles < 0) { g_err = "ncycles < 0"; return MGCFD_ERR_ARG; }
    if (c->rms_pending) { g_err = "cycles enqueued with mgcfd_enqueue_cycles are still pending: call mgcfd_collect first"; return MGCFD_ERR_ARG; }
    int done = 0, rc_final = MGCFD_OK;
    while (done < ncycles) {
        const int chunk = std::min(ncycles - done, c->rms_cap - 1);
        CKRC(mgcfd_enqueue_cycles(c, chunk));
        int rc = mgcfd_collect(c, rms_all ? rms_all + done : nullptr, rms_var ? rms_var + 5 * (size_t)done : nullptr);
        if (rc != MGCFD_OK && rc != MGCFD_ERR_INVALID_VARIABLES) return rc;
        if (rc != MGCFD_OK) rc_final = rc;
        done += chunk;
    }
    if (ncycles == 0) return mgcfd_collect(c, nullptr, nullptr);
    return rc_final;
}

int mgcfd_get_stream(mgcfd_ctx* c, void** stream) {
    if (!c || !stream) { g_err = "null argument"; return MGCFD_ERR_ARG; }
    *stream = (void*)c->stream;
    return MGCFD_OK;
}
int mgcfd_set_timing(mgcfd_ctx* c, int on) {
    if (!c) { g_err = "null context"; return MGCFD_ERR_ARG; }
    resolve_times(c);
    c->opt.timing = on ? 1 : 0;
    return MGCFD_OK;
}
int mgcfd_invalid_cell(mgcfd_ctx* c, long* cell, int* reason) {
    if (!c) { g_err = "null context"; return MGCFD_ERR_ARG; }
    if (cell) *cell = c->bad_cell;
    if (reason) *reason = c->bad_reason;
    return MGCFD_OK;
}

// ---- state access ----------------------------------------------------------------------------------------
int mgcfd_get_field(mgcfd_ctx* c, int l, int field, double* host_out) {
    CKRC(check_level(c, l));
    if (!host_out) { g_err = "null buffer"; return MGCFD_ERR_ARG; }
    Level& v = c->L[l];
    double* p; int nc;
    CKRC(field_ptr(c, v, field, &p, &nc, false));
    if (!v.io) CK(cudaMalloc((void**)&v.io, sizeof(double) * 5 * v.nel));
    const unsigned nb = (unsigned)blocks_for(v.nel, 256);
    if (field == MGCFD_FIELD_VARIABLES || field == MGCFD_FIELD_OLD_VARIABLES) k_export_recs<<<nb, 256, 0, c->stream>>>(p, v.nel, v.new_of_old, v.io);
    else k_export_soa<<<nb, 256, 0, c->stream>>>(p, v.npad, nc, v.nel, v.new_of_old, v.io);
    CKRC(post_launch(c));
    CK(cudaMemcpyAsync(host_out, v.io, sizeof(double) * nc * v.nel, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    return MGCFD_OK;
}
int mgcfd_set_field(mgcfd_ctx* c, int l, int field, const double* host_in) {
    CKRC(check_level(c, l));
    if (!host_in) { g_err = "null buffer"; return MGCFD_ERR_ARG; }
    Level& v = c->L[l];
    v.premin_valid = false;
    if (v.minword_state == 1) v.minword_state = 2;
    double* p; int nc;
    CKRC(field_ptr(c, v, field, &p, &nc, true));
    if (!v.io) CK(cudaMalloc((void**)&v.io, sizeof(double) * 5 * v.nel));
    CK(cudaMemcpyAsync(v.io, host_in, sizeof(double) * nc * v.nel, cudaMemcpyHostToDevice, c->stream));
    const unsigned nb = (unsigned)blocks_for(v.nel, 256);
    // state fields: the record's derived quantities (1/rho, p, |v|+c) are rebuilt from the five variables
    if (field == MGCFD_FIELD_VARIABLES || field == MGCFD_FIELD_OLD_VARIABLES) k_import_recs<<<nb, 256, 0, c->stream>>>(p, v.nel, v.new_of_old, v.io);
    else k_import_soa<<<nb, 256, 0, c->stream>>>(p, v.npad, nc, v.nel, v.new_of_old, v.io);
    CKRC(post_launch(c));
    CK(cudaStreamSynchronize(c->stream));
    return MGCFD_OK;
}
int mgcfd_synchronize(mgcfd_ctx* c) {
    if (!c) { g_err = "null context"; return MGCFD_ERR_ARG; }
    CK(cudaStreamSynchronize(c->stream));
    return MGCFD_OK;
}

// ---- introspection -----------------------------------------------------------------------------------------
int mgcfd_level_info(mgcfd_ctx* c, int l, long info[16]) {
    CKRC(check_level(c, l, false));
    const LevelPlan& P = c->L[l].plan;
    memset(info, 0, sizeof(long) * 16);
    info[0] = P.nel; info[1] = P.nI; info[2] = P.nB; info[3] = P.nW; info[4] = P.npad; info[5] = P.ntiles; info[6] = P.TN;
    info[7] = P.max_rounds; info[8] = P.slot_off.empty() ? 0 : P.slot_off[P.ntiles] * P.TN; info[9] = (long)P.halo_ids.size();
    info[10] = P.cut_edges / 2; info[11] = P.used_slots; info[12] = P.max_halo; info[13] = P.bslot_off.empty() ? 0 : P.bslot_off[P.ntiles] * P.TN;
    info[14] = (long)(c->L[l].pipe ? c->L[l].pipe_smem : c->L[l].smem_bytes); info[15] = c->L[l].pipe ? c->L[l].pipe_grid : 0;
    return MGCFD_OK;
}
int mgcfd_get_permutation(mgcfd_ctx* c, int l, long* new_of_old) {
    CKRC(check_level(c, l, false));
    const LevelPlan& P = c->L[l].plan;
    memcpy(new_of_old, P.new_of_old.data(), sizeof(long) * P.nel);
    return MGCFD_OK;
}
int mgcfd_visit_info(mgcfd_ctx* c, int l, long info[8]) {
    CKRC(check_level(c, l, false));
    const Level& v = c->L[l];
    memset(info, 0, sizeof(long) * 8);
    info[0] = v.visit ? 1 : 0;
    if (v.visit) {
        info[1] = v.vK; info[2] = v.vG; info[3] = v.vR | (v.vD << 8) | (v.vW << 16); info[4] = v.v_resident; info[5] = v.v_srmax; info[6] = (long)v.v_smem;
        info[7] = v.plan.visit.halo_total;
    }
    return MGCFD_OK;
}
// MGCFD_VISIT_DEBUG=1 (read at mgcfd_create): per-CTA clock stamps of the most recent visit-kernel launch, out[num_sms * 64]
int mgcfd_visit_debug(mgcfd_ctx* c, long long* out, long cap) {
    if (!c || !out) { g_err = "null argument"; return MGCFD_ERR_ARG; }
    if (!c->d_visit_dbg) { g_err = "MGCFD_VISIT_DEBUG was not set when the context was created"; return MGCFD_ERR_ARG; }
    if (cap < 128L * c->num_sms) { g_err = "buffer too small"; return MGCFD_ERR_ARG; }
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaMemcpy(out, c->d_visit_dbg, sizeof(long long) * 128 * c->num_sms, cudaMemcpyDeviceToHost));
    return 2 * c->num_sms;
}
long mgcfd_check_colouring(mgcfd_ctx* c, int l) {
    if (check_level(c, l, false) != MGCFD_OK) return -1;
    return check_colouring(c->L[l].plan);
}
int mgcfd_get_times(mgcfd_ctx* c, double* out_ms, long* out_iters) {
    if (!c) { g_err = "null context"; return MGCFD_ERR_ARG; }
    resolve_times(c);
    if (out_ms) memcpy(out_ms, c->t_ms.data(), sizeof(double) * c->t_ms.size());
    if (out_iters) memcpy(out_iters, c->t_iters.data(), sizeof(long) * c->t_iters.size());
    return MGCFD_OK;
}
int mgcfd_reset_times(mgcfd_ctx* c) {
    if (!c) { g_err = "null context"; return MGCFD_ERR_ARG; }
    resolve_times(c);
    std::fill(c->t_ms.begin(), c->t_ms.end(), 0.0);
    std::fill(c->t_iters.begin(), c->t_iters.end(), 0L);
    return MGCFD_OK;
}
long mgcfd_launch_count(mgcfd_ctx* c) { return c ? c->launches : -1; }

int mgcfd_time_kernel(mgcfd_ctx* c, int l, int which, int reps, double* ms_total) {
    CKRC(check_level(c, l));
    if (reps < 1 || !ms_total) { g_err = "bad arguments"; return MGCFD_ERR_ARG; }
    Level& v = c->L[l];
    CKRC(ensure_flux(c, v));
    if (which == 2 || which == 3) CKRC(ensure_flat(c, v));
    if ((which == 0 || which == 1 || which == 5) && c->opt.flux_mode == MGCFD_FLUX_ATOMIC) { g_err = "the stage kernel needs a tiled flux mode"; return MGCFD_ERR_ARG; }
    if (which == 24) CKRC(launch_flux_variant(c, v, 8, true));      // allocate + fill the SoA copy of the node state, untimed
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaEventRecord(c->ev0, c->stream));
    for (int r = 0; r < reps; r++) {
        if (which == 0 || which == 1 || which == 5) {
            StageArgs a = base_args(c, v);
            a.vin = v.V(v.i_var); a.vold = v.V(v.i_var); a.rk_div = 4.0; a.rk_rcp = 0.25;
            if (which == 0) { a.vout = v.V(v.i_tmp); a.mask = 7; CKRC(launch_stage(c, v, a, true)); }
            else if (which == 5) { a.vout = v.V(v.i_tmp); a.mask = 6; CKRC(launch_stage(c, v, a, true)); }   // no internal edges: staging + update only
            else { a.flux = v.flux; a.mask = 1; CKRC(launch_stage(c, v, a, false)); }
        } else if (which == 2) {
            k_indirect_rw<<<(unsigned)blocks_for(v.nI, 256), 256, 0, c->stream>>>(v.nI, v.ea, v.eb, v.ew, v.V(v.i_var), v.npad, v.flux); CKRC(post_launch(c));
        } else if (which == 3) {
            k_flux_atomic<<<(unsigned)blocks_for(v.nI, 256), 256, 0, c->stream>>>(v.nI, v.ea, v.eb, v.ew, v.V(v.i_var), v.npad, v.flux, 2.0 * c->kdiss); CKRC(post_launch(c));
        } else if (which >= 16 && which <= 24) {
            CKRC(launch_flux_variant(c, v, which - 16, false));      // (24: the planes were filled before the timed region)
        } else { g_err = "unknown kernel selector"; return MGCFD_ERR_ARG; }
    }
    CK(cudaEventRecord(c->ev1, c->stream));
    CK(cudaEventSynchronize(c->ev1));
    float ms = 0;
    CK(cudaEventElapsedTime(&ms, c->ev0, c->ev1));
    *ms_total = ms;
    CK(cudaMemsetAsync(v.flux, 0, sizeof(double) * 5 * v.npad, c->stream));
    return MGCFD_OK;
}

int mgcfd_flux_variant(mgcfd_ctx* c, int l, int bits) {
    CKRC(check_level(c, l));
    return launch_flux_variant(c, c->L[l], bits);
}

int mgcfd_plan_level(long nel, const double* coords, long nI, long nB, long nW, const void* edges, int ordering, int tile_nodes,
                     int flux_mode, long info[16], long* new_of_old, long* conflicts) {
    if (nel <= 0 || !edges || !info) { g_err = "bad arguments"; return MGCFD_ERR_ARG; }
    HostLevel H;
    H.nel = nel; H.nI = nI; H.nB = nB; H.nW = nW;
    H.volumes.assign(nel, 1.0);
    const EdgeNb* e = (const EdgeNb*)edges;
    H.edges.assign(e, e + nI + nB + nW);
    if (coords) H.coords.assign(coords, coords + 3 * nel);
    PlanOptions po; po.ordering = ordering; po.tile_nodes = tile_nodes ? tile_nodes : auto_tile_nodes(nel, nI, 148); po.scatter = (flux_mode == MGCFD_FLUX_TILED_COLOURED); po.strict = false;
    LevelPlan P;
    try { build_level_plan(H, po, P); }
    catch (const std::exception& ex) { g_err = ex.what(); return MGCFD_ERR_ARG; }
    memset(info, 0, sizeof(long) * 16);
    info[0] = P.nel; info[1] = P.nI; info[2] = P.nB; info[3] = P.nW; info[4] = P.npad; info[5] = P.ntiles; info[6] = P.TN;
    info[7] = P.max_rounds; info[8] = P.slot_off[P.ntiles] * P.TN; info[9] = (long)P.halo_ids.size();
    info[10] = P.cut_edges / 2; info[11] = P.used_slots; info[12] = P.max_halo; info[13] = P.bslot_off[P.ntiles] * P.TN;
    {   // info[15]: FNV-1a over everything the device receives of the level -- the plan must be a pure function of the mesh
        unsigned long long h = 1469598103934665603ull;
        auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
        mix(P.new_of_old.data(), P.new_of_old.size() * sizeof(long));
        mix(P.hdrs.data(), P.hdrs.size()); mix(P.slots.data(), P.slots.size()); mix(P.bslots.data(), P.bslots.size());
        mix(P.halo_ids.data(), P.halo_ids.size() * sizeof(int)); mix(P.tile_nown.data(), P.tile_nown.size() * sizeof(int));
        mix(P.adj_off.data(), P.adj_off.size() * sizeof(long)); mix(P.adj_nbr.data(), P.adj_nbr.size() * sizeof(int));
        const long tail[4] = {P.cut_edges, P.used_slots, P.max_halo, P.max_rounds};
        mix(tail, sizeof(tail));
        info[15] = (long)(h >> 1);
    }
    if (new_of_old) memcpy(new_of_old, P.new_of_old.data(), sizeof(long) * nel);
    if (conflicts) *conflicts = P.oversize ? -1 : check_colouring(P);
    return MGCFD_OK;
}

// host-only checking aids: the plan's byte streams walked on the host (plan.h: emulate_*)
static int host_level(HostLevel& H, long nel, const double* coords, long nI, long nB, long nW, const void* edges, const long* mg_map) {
    if (nel <= 0 || !edges || nI < 0 || nB < 0 || nW < 0) { g_err = "bad arguments"; return MGCFD_ERR_ARG; }
    H.nel = nel; H.nI = nI; H.nB = nB; H.nW = nW;
    H.volumes.assign(nel, 1.0);
    const EdgeNb* e = (const EdgeNb*)edges;
    H.edges.assign(e, e + nI + nB + nW);
    if (coords) H.coords.assign(coords, coords + 3 * nel);
    if (mg_map) H.mg.assign(mg_map, mg_map + nel);
    return MGCFD_OK;
}
int mgcfd_plan_emulate_flux(long nel, const double* coords, long nI, long nB, long nW, const void* edges, int ordering, int tile_nodes,
                            int flux_mode, const double* variables, int mask, double* fluxes) {
    if (!variables || !fluxes || flux_mode == MGCFD_FLUX_ATOMIC) { g_err = "bad arguments"; return MGCFD_ERR_ARG; }
    HostLevel H;
    CKRC(host_level(H, nel, coords, nI, nB, nW, edges, nullptr));
    PlanOptions po; po.ordering = ordering; po.tile_nodes = tile_nodes ? tile_nodes : auto_tile_nodes(nel, nI, 148); po.scatter = (flux_mode == MGCFD_FLUX_TILED_COLOURED);
    double ff[5], ffc[12];
    mgcfd_far_field_conditions(ff, ffc);
    try {
        LevelPlan P;
        build_level_plan(H, po, P);
        emulate_stage_flux(P, variables, mask, ff, ffc, 2.0 * (-0.5 * double(0.2f)), fluxes);
    } catch (const std::exception& ex) { g_err = ex.what(); return MGCFD_ERR_ARG; }
    return MGCFD_OK;
}
// the visit kernel's streams of the same level (plan.h VisitPlan), grouped into `supers` super-tiles; info[0..7] = super-tiles, most
// tiles per super-tile, most halo rows of a super-tile, halo rows in total, most rounds of a tile, tiles, padded rows, 0
int mgcfd_plan_emulate_visit_flux(long nel, const double* coords, long nI, long nB, long nW, const void* edges, int supers,
                                  const double* variables, int mask, double* fluxes, long info[8]) {
    if (!variables || !fluxes || supers < 1) { g_err = "bad arguments"; return MGCFD_ERR_ARG; }
    HostLevel H;
    CKRC(host_level(H, nel, coords, nI, nB, nW, edges, nullptr));
    PlanOptions po; po.ordering = MGCFD_ORDER_PARTITION_RCM; po.tile_nodes = VT; po.scatter = false; po.supers = supers;
    double ff[5], ffc[12];
    mgcfd_far_field_conditions(ff, ffc);
    try {
        LevelPlan P;
        build_level_plan(H, po, P);
        emulate_visit_flux(P, variables, mask, ff, ffc, 2.0 * (-0.5 * double(0.2f)), fluxes);
        if (info) {
            memset(info, 0, sizeof(long) * 8);
            info[0] = P.visit.ns; info[1] = P.visit.maxt; info[2] = P.visit.max_halo; info[3] = P.visit.halo_total; info[4] = P.visit.max_rounds;
            info[5] = P.ntiles; info[6] = P.npad; info[7] = P.visit.went_off[P.visit.ns];
        }
    } catch (const std::exception& ex) { g_err = ex.what(); return MGCFD_ERR_ARG; }
    return MGCFD_OK;
}
// host-only: the configuration the library would choose for the visit kernel on a device with `num_sms` SMs;
// info[0..7] = usable (0/1), super-tiles per CTA, CTAs, rounds per ring entry, resident, rows of the largest super-tile, shared memory, halo rows
int mgcfd_plan_visit_config(long nel, const double* coords, long nI, long nB, long nW, const void* edges, int num_sms, long info[8]) {
    if (!info || num_sms < 1) { g_err = "bad arguments"; return MGCFD_ERR_ARG; }
    HostLevel H;
    CKRC(host_level(H, nel, coords, nI, nB, nW, edges, nullptr));
    PlanOptions po; po.ordering = MGCFD_ORDER_PARTITION_RCM; po.scatter = false;
    memset(info, 0, sizeof(long) * 8);
    try {
        LevelPlan P; VisitCfg cfg;
        if (plan_for_visit(num_sms, P, H, po, cfg)) {
            info[0] = 1; info[1] = cfg.K; info[2] = cfg.G; info[3] = cfg.R | (cfg.D << 8) | (cfg.W << 16); info[4] = cfg.resident; info[5] = cfg.sr_max; info[6] = (long)cfg.smem;
            info[7] = P.visit.halo_total;
        }
    } catch (const std::exception& ex) { g_err = ex.what(); return MGCFD_ERR_ARG; }
    return MGCFD_OK;
}
int mgcfd_plan_emulate_transfers(long nel_f, const double* coords_f, long nI_f, long nB_f, long nW_f, const void* edges_f, const long* mg_map,
                                 long nel_c, const double* coords_c, long nI_c, long nB_c, long nW_c, const void* edges_c, int ordering,
                                 int tile_nodes, const double* var_f, const double* res_f, const double* res_c, double* var_c, double* var_f_out) {
    if (!coords_f || !coords_c || !mg_map || !var_f || !res_f || !res_c || !var_c || !var_f_out) { g_err = "null argument"; return MGCFD_ERR_ARG; }
    HostLevel F, C;
    CKRC(host_level(F, nel_f, coords_f, nI_f, nB_f, nW_f, edges_f, mg_map));
    CKRC(host_level(C, nel_c, coords_c, nI_c, nB_c, nW_c, edges_c, nullptr));
    for (long i = 0; i < nel_f; i++) if (mg_map[i] < 0 || mg_map[i] >= nel_c) { g_err = "mg_map entry out of range"; return MGCFD_ERR_ARG; }
    try {
        PlanOptions po; po.ordering = ordering;
        LevelPlan Pf, Pc;
        po.tile_nodes = tile_nodes ? tile_nodes : auto_tile_nodes(nel_f, nI_f, 148); build_level_plan(F, po, Pf);
        po.tile_nodes = tile_nodes ? tile_nodes : auto_tile_nodes(nel_c, nI_c, 148); build_level_plan(C, po, Pc);
        TransferPlan T;
        build_transfer_plan(F, C, Pf, Pc, T);
        emulate_restrict(Pf, Pc, T, var_f, var_c);
        memcpy(var_f_out, var_f, sizeof(double) * 5 * nel_f);
        emulate_prolong(Pf, Pc, T, res_c, res_f, var_f_out);
    } catch (const std::exception& ex) { g_err = ex.what(); return MGCFD_ERR_ARG; }
    return MGCFD_OK;
}

// ---- distributed runs (include/mgcfd_dist.h) -----------------------------------------------------------------
int mgcfd_dist_get_unique_id(char id[128]) {
    if (!id) { g_err = "null argument"; return MGCFD_ERR_ARG; }
    CKRC(nccl_load());
    ncclUniqueId u;
    NK(g_nccl.GetUniqueId(&u));
    memcpy(id, u.internal, 128);
    return MGCFD_OK;
}
int mgcfd_dist_init(mgcfd_ctx* c, int rank, int nranks, const char id[128]) {
    if (!c || !id || nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) { g_err = "bad arguments"; return MGCFD_ERR_ARG; }
    if (c->finalized || c->dist.active) { g_err = "mgcfd_dist_init must be called once, before any level is uploaded"; return MGCFD_ERR_ARG; }
    if (c->opt.flux_mode == MGCFD_FLUX_ATOMIC) { g_err = "distributed runs need a tiled flux mode"; return MGCFD_ERR_ARG; }
    CKRC(nccl_load());
    CK(cudaSetDevice(c->opt.device));
    ncclUniqueId u;
    memcpy(u.internal, id, 128);
    NK(g_nccl.CommInitRank(&c->dist.comm, nranks, u, rank));
    c->dist.active = true; c->dist.rank = rank; c->dist.nranks = nranks;
    // the cycle interleaves NCCL calls with kernels: launched eagerly unless MGCFD_DIST_GRAPH=1 asks for the captured form
    // (NCCL supports stream capture; every rank must then capture the same sequence)
    c->dist.want_graph = c->opt.use_graph != 0;
    { const char* e = getenv("MGCFD_DIST_GRAPH"); if (!(e && e[0] == '1')) c->opt.use_graph = 0; }
    return MGCFD_OK;
}
// uploads one level of a partition (partition.h) and records its exchange lists
static int upload_local_level(mgcfd_ctx* c, int l, const LocalLevel& LL, long nel_global) {
    const HostLevel& M = LL.mesh;
    int rc = mgcfd_upload_level(c, l, M.nel, M.volumes.data(), M.coords.empty() ? nullptr : M.coords.data(), M.nI, M.nB, M.nW, M.edges.data(),
                                M.mg.empty() ? nullptr : M.mg.data(), (long)M.mg.size());
    if (rc) return rc;
    Level& v = c->L[l];
    v.send_off = LL.send_off; v.recv_off = LL.recv_off; v.send_list = LL.send_idx; v.gid = LL.gid;
    v.host.gid = LL.gid;
    v.nel_global = nel_global;
    v.nI_global = LL.nI_global;
    return MGCFD_OK;
}
int mgcfd_dist_level_info(mgcfd_ctx* c, int l, long info[8]) {
    CKRC(check_level(c, l, false));
    const Level& v = c->L[l];
    memset(info, 0, sizeof(long) * 8);
    info[0] = v.plan.n_owned; info[1] = v.plan.nel - v.plan.n_owned; info[2] = (long)v.send_list.size(); info[3] = v.nel_global;
    info[4] = c->dist.rank; info[5] = c->dist.nranks; info[6] = c->dist.exchanges; info[7] = v.nI_global ? v.nI_global : v.nI;
    return MGCFD_OK;
}
int mgcfd_dist_global_ids(mgcfd_ctx* c, int l, long* gid) {
    CKRC(check_level(c, l, false));
    const Level& v = c->L[l];
    if (!gid) { g_err = "null buffer"; return MGCFD_ERR_ARG; }
    if (v.gid.empty()) for (long i = 0; i < v.plan.nel; i++) gid[i] = i;
    else memcpy(gid, v.gid.data(), sizeof(long) * v.gid.size());
    return MGCFD_OK;
}
int mgcfd_upload_partition(mgcfd_ctx* c, int levels, int mesh_variant, const void* host_mesh_opaque) {
    // host_mesh_opaque: const mgcfd::HostMesh* (called from mesh_api.cpp, which owns the mesh type)
    if (!c || !host_mesh_opaque) { g_err = "null argument"; return MGCFD_ERR_ARG; }
    if (!c->dist.active) { g_err = "mgcfd_dist_init has not been called"; return MGCFD_ERR_ARG; }
    const HostMesh& full = *(const HostMesh*)host_mesh_opaque;
    if ((int)full.levels.size() != c->levels || levels != c->levels || mesh_variant != c->variant) { g_err = "mesh does not match the context"; return MGCFD_ERR_ARG; }
    LocalMesh loc;
    try { partition_mesh(full, c->dist.nranks, c->dist.rank, loc); }
    catch (const std::exception& ex) { g_err = ex.what(); return MGCFD_ERR_ARG; }
    for (int l = 0; l < c->levels; l++) {
        // mgcfd_upload_level copies into Level::host; n_owned / gid ride along for the plan builder
        c->L[l].host.n_owned = loc.levels[l].n_owned;
        CKRC(upload_local_level(c, l, loc.levels[l], full.levels[l].nel));
    }
    return mgcfd_finalize(c);
}
// ---- direct peer-to-peer data path (CUDA IPC) -------------------------------------------------------------------------------
// table layout (longs): [0] = levels, then per level: stage_off (doubles), nghost, recv_off[0..nranks], then per level: first ghost
// row, npad, byte offsets of the three record buffers and of the residual planes inside the slab.  ONE CUDA IPC handle per rank
// exposes its slab (window + record buffers + residual planes of every level): that is everything other ranks ever write.
static const long SLAB_PER_LEVEL = 6;
long mgcfd_dist_p2p_table_len(mgcfd_ctx* c) {
    if (!c) return -1;
    return 1 + (long)c->levels * (2 + c->dist.nranks + 1) + (long)c->levels * SLAB_PER_LEVEL;
}
int mgcfd_dist_p2p_prepare(mgcfd_ctx* c, char handle[64], long* table, long table_cap) {
    if (!c || !handle || !table) { g_err = "null argument"; return MGCFD_ERR_ARG; }
    if (!c->dist.active || !c->finalized) { g_err = "mgcfd_dist_p2p_prepare needs a finalized distributed context"; return MGCFD_ERR_ARG; }
    Dist& d = c->dist;
    const long need = mgcfd_dist_p2p_table_len(c);
    if (table_cap < need) { g_err = "table too small"; return MGCFD_ERR_ARG; }
    CK(cudaSetDevice(c->opt.device));
    if (!d.d_op) {
        CK(cudaMalloc((void**)&d.d_ticket, 4 * 4096)); CK(cudaMemset(d.d_ticket, 0, 4 * 4096));      // [0] + one word per group of 32 blocks
        CK(cudaMalloc((void**)&d.d_op, 8)); CK(cudaMemset(d.d_op, 0, 8));
        CK(cudaMalloc((void**)&d.d_ctr, 4 * (c->levels + 1))); CK(cudaMemset(d.d_ctr, 0, 4 * (c->levels + 1)));
        CK(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    CK(cudaIpcGetMemHandle(&h, c->slab));
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    memcpy(handle, &h, 64);
    long k = 0;
    table[k++] = c->levels;
    for (int l = 0; l < c->levels; l++) {
        table[k++] = d.stage_off[l]; table[k++] = c->L[l].nghost;
        for (int p = 0; p <= d.nranks; p++) table[k++] = c->L[l].recv_off[p];
    }
    for (int l = 0; l < c->levels; l++) {
        table[k++] = c->L[l].ncomp; table[k++] = c->L[l].npad;
        for (int b = 0; b < 4; b++) table[k++] = (long)d.buf_off[4 * l + b];
    }
    return MGCFD_OK;
}
// handles: nranks x 64 bytes, tables: nranks x table_len longs, both in rank order (what every rank's prepare returned)
int mgcfd_dist_p2p_attach(mgcfd_ctx* c, const char* handles, const long* tables, long table_len) {
    if (!c || !handles || !tables) { g_err = "null argument"; return MGCFD_ERR_ARG; }
    Dist& d = c->dist;
    if (!d.active || !d.win || !d.d_op) { g_err = "mgcfd_dist_p2p_prepare has not been called"; return MGCFD_ERR_ARG; }
    d.off = 0;
    if (table_len != mgcfd_dist_p2p_table_len(c)) { g_err = "table length does not match this build of the library"; return MGCFD_ERR_COMM; }
    CK(cudaSetDevice(c->opt.device));
    d.peer_win.assign(d.nranks, nullptr);
    for (int p = 0; p < d.nranks; p++) {
        if (p == d.rank) { d.peer_win[p] = d.win; continue; }
        cudaIpcMemHandle_t h;
        memcpy(&h, handles + 64 * (size_t)p, 64);
        CK(cudaIpcOpenMemHandle(&d.peer_win[p], h, cudaIpcMemLazyEnablePeerAccess));
    }
    std::vector<double*> red(d.nranks);
    std::vector<unsigned long long*> flg(d.nranks);
    for (int p = 0; p < d.nranks; p++) {
        red[p] = (double*)((unsigned char*)d.peer_win[p] + P2P_FLAGS_BYTES);
        flg[p] = (unsigned long long*)d.peer_win[p] + d.rank;
    }
    CKRC(dev_upload(&d.d_red_of_rank, red, c->stream)); CKRC(dev_upload(&d.d_flag_of_rank, flg, c->stream));
    const long per_level = 2 + d.nranks + 1;
    const long base_len = 1 + (long)c->levels * per_level;
    for (int l = 0; l < c->levels; l++) {
        Level& v = c->L[l];
        std::vector<P2PPeer> peers;
        std::vector<PeerOut> pouts;
        std::vector<PeerSlice> slices;
        for (int p = 0; p < d.nranks; p++) {
            const long ns = v.send_off[p + 1] - v.send_off[p], nr = v.recv_off[p + 1] - v.recv_off[p];
            if (p == d.rank || (ns == 0 && nr == 0)) continue;
            const long* tp = tables + (size_t)p * table_len + 1 + (size_t)l * per_level;     // peer p's entry for level l
            const long* sp = tables + (size_t)p * table_len + base_len + (size_t)l * SLAB_PER_LEVEL;
            const long p_stage_off = tp[0], p_nghost = tp[1], p_recv_me = tp[2 + d.rank], p_recv_me_n = tp[2 + d.rank + 1] - tp[2 + d.rank];
            if (p_recv_me_n != ns) { g_err = "send / receive lists of two ranks do not match"; return MGCFD_ERR_COMM; }
            P2PPeer e;
            e.rank = p; e.send0 = v.send_off[p]; e.nsend = ns; e.recv0 = v.recv_off[p]; e.nrecv = nr;
            double* pst = (double*)((unsigned char*)d.peer_win[p] + P2P_HDR_BYTES) + p_stage_off;
            // the staging buffer of a level holds up to 8 doubles per ghost row; my rows start at the peer's recv_off[me]
            e.dst[0] = pst + 8 * p_recv_me;
            e.dst[1] = pst + 8 * (size_t)std::max<long>(p_nghost, 1) + 8 * p_recv_me;
            e.flag = (unsigned long long*)d.peer_win[p] + d.rank;
            peers.push_back(e);
            // where the peer keeps its copies of my rows: its record buffers / residual planes inside its slab
            PeerOut po;
            for (int b = 0; b < 3; b++) po.rec[b] = (double*)((unsigned char*)d.peer_win[p] + sp[2 + b]);
            po.res = (double*)((unsigned char*)d.peer_win[p] + sp[5]);
            po.res_stride = sp[1];
            pouts.push_back(po);
            PeerSlice sl;
            sl.send0 = e.send0; sl.nsend = ns; sl.first_ghost_row = sp[0]; sl.recv_off_me = p_recv_me;
            slices.push_back(sl);
        }
        v.npeers = (int)peers.size();
        if (v.npeers > 64) { g_err = "too many peers"; return MGCFD_ERR_ARG; }
        CKRC(dev_upload(&v.d_peers, peers, c->stream));
        CKRC(dev_upload(&v.d_peer_out, pouts, c->stream));
        v.h_peers = peers;
        // row -> (peer, row in the peer's arrays): build_send_targets (partition.h); every rank builds them, peers or not
        SendTargets st;
        try { build_send_targets(v.ncomp, v.TN, v.h_send_idx, slices, st); }
        catch (const std::exception& ex) { g_err = ex.what(); return MGCFD_ERR_ARG; }
        CKRC(dev_upload(&v.d_tgt_off, st.off, c->stream)); CKRC(dev_upload(&v.d_tgt_peer, st.peer, c->stream)); CKRC(dev_upload(&v.d_tgt_row, st.row, c->stream));
        CKRC(dev_upload(&v.d_tile_sends, st.tile_sends, c->stream));
        if (v.pipe) { CKRC(setup_pipe_dist(c, v)); if (v.pipe_grid_dist < 1) v.pipe = false; }
        // tiles in the order the stage kernel takes them: the tiles that own rows on send lists -- the only ones that read ghost
        // rows, flux halos being symmetric -- LAST (stable otherwise); build_tile_order (partition.h) also looks at the halo lists
        {
            std::vector<int> order;
            v.n_send_tiles = build_tile_order(v.ncomp, v.TN, st, v.plan.halo_off, v.plan.halo_ids, (long)v.plan.ntiles, order);
            CKRC(dev_upload(&v.d_order_tiles, order, c->stream));
        }
    }
    {   // every neighbour rank of this rank, whatever the level: who is told at the start of a kernel that the previous one is complete
        std::vector<P2PPeer> sig;
        for (int p = 0; p < d.nranks; p++) {
            if (p == d.rank) continue;
            bool nb = false;
            for (int l = 0; l < c->levels; l++) for (const P2PPeer& e : c->L[l].h_peers) if (e.rank == p) nb = true;
            if (!nb) continue;
            P2PPeer e;
            memset(&e, 0, sizeof(e));
            e.rank = p; e.flag = (unsigned long long*)d.peer_win[p] + d.rank;
            sig.push_back(e);
        }
        d.nsig = (int)sig.size();
        CKRC(dev_upload(&d.d_sig, sig, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    d.p2p = true;
    c->opt.use_graph = c->dist.want_graph;     // epoch numbers live on the device: the cycle, exchanges included, is graph-capturable
    return MGCFD_OK;
}

// FNV-1a over everything a rank holds of one level: two ways of computing a partition must agree on this bit for bit
static unsigned long long hash_local_level(const LocalLevel& LL) {
    unsigned long long h = 1469598103934665603ull;
    auto mix = [&](const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { h ^= b[i]; h *= 1099511628211ull; } };
    const HostLevel& M = LL.mesh;
    const long hdr[5] = {M.nel, M.nI, M.nB, M.nW, LL.n_owned};
    mix(hdr, sizeof(hdr));
    mix(&LL.nI_global, sizeof(long));
    mix(M.volumes.data(), M.volumes.size() * sizeof(double));
    mix(M.coords.data(), M.coords.size() * sizeof(double));
    mix(M.edges.data(), M.edges.size() * sizeof(EdgeNb));
    mix(M.mg.data(), M.mg.size() * sizeof(long));
    mix(LL.gid.data(), LL.gid.size() * sizeof(long));
    mix(LL.edge_gid.data(), LL.edge_gid.size() * sizeof(long));
    mix(LL.send_off.data(), LL.send_off.size() * sizeof(long));
    mix(LL.send_idx.data(), LL.send_idx.size() * sizeof(long));
    mix(LL.recv_off.data(), LL.recv_off.size() * sizeof(long));
    return h;
}
static void fill_plan_outputs(const LocalLevel& LL, long nel_global, int nranks, long info[8], long* gid, long* send_counts, long* recv_counts, long* send_gids) {
    memset(info, 0, sizeof(long) * 8);
    info[0] = LL.n_owned; info[1] = (long)LL.gid.size() - LL.n_owned; info[2] = (long)LL.send_idx.size(); info[3] = nel_global;
    info[4] = LL.mesh.nI; info[5] = LL.mesh.nB; info[6] = LL.mesh.nW; info[7] = (long)(hash_local_level(LL) >> 1);
    if (gid) memcpy(gid, LL.gid.data(), sizeof(long) * LL.gid.size());
    for (int p = 0; p < nranks; p++) {
        if (send_counts) send_counts[p] = LL.send_off[p + 1] - LL.send_off[p];
        if (recv_counts) recv_counts[p] = LL.recv_off[p + 1] - LL.recv_off[p];
    }
    if (send_gids) for (size_t k = 0; k < LL.send_idx.size(); k++) send_gids[k] = LL.gid[LL.send_idx[k]];
}
// host-only: the partition of one level for one rank (tests, tooling); any output pointer may be NULL
int mgcfd_partition_plan(int levels, const void* host_mesh_opaque, int nranks, int rank, int level, long info[8], long* gid,
                         long* send_counts, long* recv_counts, long* send_gids) {
    if (!host_mesh_opaque || !info) { g_err = "null argument"; return MGCFD_ERR_ARG; }
    const HostMesh& full = *(const HostMesh*)host_mesh_opaque;
    if (level < 0 || level >= (int)full.levels.size() || levels != (int)full.levels.size()) { g_err = "bad level"; return MGCFD_ERR_ARG; }
    LocalMesh loc;
    try { partition_mesh(full, nranks, rank, loc); }
    catch (const std::exception& ex) { g_err = ex.what(); return MGCFD_ERR_ARG; }
    fill_plan_outputs(loc.levels[level], full.levels[level].nel, nranks, info, gid, send_counts, recv_counts, send_gids);
    return MGCFD_OK;
}

// Host-only check of the in-kernel halo exchange's tables (no device is touched): the mesh is split over `nranks` ranks in this one
// process, every rank's levels are planned as mgcfd_finalize would (same numbering), every rank builds its row -> (peer, remote row)
// targets and its tile order as mgcfd_dist_p2p_attach does -- from what the OTHER ranks would publish (first ghost row, recv_off) --
// and the delivery is replayed on global node ids: a producer "stores" the id of each of its rows into the target rows.
//   out[0] rows delivered   out[1] ghost rows over all ranks and levels
//   out[2] errors: a row delivered to a ghost row that holds another node, a ghost row hit twice or never
//   out[3] tiles that read a ghost row in their halo but are not among the tiles taken last (the late wait would miss them)
//   out[4] transfer-kernel blocks that read a ghost row (children of owned coarse rows / parents and operator sources of owned fine rows)
int mgcfd_mesh_delivery_check_impl(const void* host_mesh_opaque, int nranks, int tile_nodes, long out[5]) {
    if (!host_mesh_opaque || !out || nranks < 1) { g_err = "bad arguments"; return MGCFD_ERR_ARG; }
    const HostMesh& full = *(const HostMesh*)host_mesh_opaque;
    const int nl = (int)full.levels.size();
    for (int k = 0; k < 5; k++) out[k] = 0;
    try {
        std::vector<LocalMesh> loc(nranks);
        for (int r = 0; r < nranks; r++) partition_mesh(full, nranks, r, loc[r]);
        std::vector<std::vector<LevelPlan>> plans(nranks, std::vector<LevelPlan>(nl));
        for (int l = 0; l < nl; l++) {
            std::vector<std::vector<long>> row_gid(nranks);      // device row -> global node id (-1: padding)
            for (int r = 0; r < nranks; r++) {
                const LocalLevel& LL = loc[r].levels[l];
                HostLevel H = LL.mesh;
                H.n_owned = LL.n_owned;
                PlanOptions po;
                po.ordering = MGCFD_ORDER_PARTITION_RCM; po.scatter = false; po.strict = false;
                po.tile_nodes = tile_nodes ? tile_nodes : auto_tile_nodes(LL.n_owned, H.nI, 148);
                build_level_plan(H, po, plans[r][l]);
                const LevelPlan& P = plans[r][l];
                row_gid[r].assign(P.npad, -1);
                for (long i = 0; i < P.nel; i++) row_gid[r][P.new_of_old[i]] = LL.gid[i];
                for (long j = 0; j < P.nel - P.n_owned; j++)
                    if (P.new_of_old[P.n_owned + j] != P.npad_owned + j) throw std::runtime_error("ghost rows do not follow the owned tiles in order");
                out[1] += P.nel - P.n_owned;
            }
            std::vector<std::vector<int>> hits(nranks);
            for (int r = 0; r < nranks; r++) hits[r].assign(plans[r][l].npad, 0);
            for (int r = 0; r < nranks; r++) {
                const LocalLevel& LL = loc[r].levels[l];
                const LevelPlan& P = plans[r][l];
                std::vector<int> send_rows(LL.send_idx.size());
                for (size_t k = 0; k < send_rows.size(); k++) send_rows[k] = (int)P.new_of_old[LL.send_idx[k]];
                std::vector<PeerSlice> slices;
                std::vector<int> peer_rank;
                for (int p = 0; p < nranks; p++) {
                    const long ns = LL.send_off[p + 1] - LL.send_off[p], nr = LL.recv_off[p + 1] - LL.recv_off[p];
                    if (p == r || (ns == 0 && nr == 0)) continue;
                    const LocalLevel& PL = loc[p].levels[l];
                    if (PL.recv_off[r + 1] - PL.recv_off[r] != ns) throw std::runtime_error("send / receive lists of two ranks do not match");
                    PeerSlice sl;
                    sl.send0 = LL.send_off[p]; sl.nsend = ns; sl.first_ghost_row = plans[p][l].npad_owned; sl.recv_off_me = PL.recv_off[r];
                    slices.push_back(sl); peer_rank.push_back(p);
                }
                SendTargets st;
                build_send_targets(P.npad_owned, P.TN, send_rows, slices, st);
                for (long row = 0; row < P.npad_owned; row++)
                    for (int k = st.off[row]; k < st.off[row + 1]; k++) {
                        const int p = peer_rank[st.peer[k]];
                        const long tr = st.row[k];
                        out[0]++;
                        if (tr < plans[p][l].npad_owned || tr >= plans[p][l].npad || row_gid[p][tr] != row_gid[r][row] || row_gid[r][row] < 0) out[2]++;
                        else hits[p][tr]++;
                    }
                std::vector<int> order;
                const int n_last = build_tile_order(P.npad_owned, P.TN, st, P.halo_off, P.halo_ids, (long)P.ntiles, order);
                std::vector<char> last(order.size(), 0);
                for (size_t k = order.size() - (size_t)n_last; k < order.size(); k++) last[order[k]] = 1;
                for (long t = 0; t < (long)P.ntiles; t++)
                    for (long k = P.halo_off[t]; k < P.halo_off[t + 1]; k++) if (P.halo_ids[k] >= P.npad_owned && !last[t]) { out[3]++; break; }
                std::vector<char> seen(order.size(), 0);
                for (int u : order) { if (u < 0 || u >= (int)order.size() || seen[u]) out[2]++; else seen[u] = 1; }
            }
            for (int r = 0; r < nranks; r++) {
                const LevelPlan& P = plans[r][l];
                for (long row = P.npad_owned; row < P.npad_owned + (P.nel - P.n_owned); row++) if (hits[r][row] != 1) out[2]++;
            }
        }
        // the transfer kernels' per-block wait flags (mgcfd_finalize): count the 128-row blocks that read a ghost row
        for (int r = 0; r < nranks; r++)
            for (int l = 0; l + 1 < nl; l++) {
                HostLevel Hf = loc[r].levels[l].mesh, Hc = loc[r].levels[l + 1].mesh;
                Hf.n_owned = loc[r].levels[l].n_owned; Hc.n_owned = loc[r].levels[l + 1].n_owned;
                TransferPlan T;
                build_transfer_plan(Hf, Hc, plans[r][l], plans[r][l + 1], T);
                const long ncf = plans[r][l].npad_owned, ncc = plans[r][l + 1].npad_owned;
                std::vector<char> rw(std::max<long>(1, (ncc + 127) / 128), 0), pw(std::max<long>(1, (ncf + 127) / 128), 0);
                for (long cc = 0; cc < ncc; cc++)
                    for (long k = T.child_off[cc]; k < T.child_off[cc + 1]; k++) if (T.child_ids[k] >= ncf) { rw[cc / 128] = 1; break; }
                for (long i = 0; i < ncf; i++) {
                    bool g = T.parent[i] >= ncc;
                    for (long k = T.ent_off[i]; k < T.ent_off[i + 1] && !g; k++) g = T.ent_src[k] >= ncc;
                    if (g) pw[i / 128] = 1;
                }
                for (char x : rw) out[4] += x;
                for (char x : pw) out[4] += x;
            }
    }
    catch (const std::exception& ex) { g_err = ex.what(); return MGCFD_ERR_ARG; }
    return MGCFD_OK;
}

// ---- rank-local generation: the part of a synthetic mesh this rank holds, without assembling the global mesh ----------------
static int make_spec(int kind, int levels, const long* dims, const double lengths[3], int mesh_variant, double tilt, MeshSpec& s) {
    if (!dims || levels < 1 || levels > 8) { g_err = "bad arguments"; return MGCFD_ERR_ARG; }
    s.kind = kind; s.levels = levels; s.mesh_variant = mesh_variant; s.ordering = 0; s.tilt = tilt;
    for (int l = 0; l < levels; l++) for (int k = 0; k < 3; k++) s.dims[l][k] = dims[3 * l + k];
    if (lengths) for (int k = 0; k < 3; k++) s.lengths[k] = lengths[k];
    return MGCFD_OK;
}
int mgcfd_generate_partition_plan(int kind, int levels, const long* dims, const double lengths[3], int mesh_variant, double tilt, int apply_ewt_too,
                                  int nranks, int rank, int level, long info[8], long* gid, long* send_counts, long* recv_counts, long* send_gids) {
    MeshSpec spec;
    CKRC(make_spec(kind, levels, dims, lengths, mesh_variant, tilt, spec));
    if (!info || level < 0 || level >= levels) { g_err = "bad level"; return MGCFD_ERR_ARG; }
    LocalMesh loc;
    std::string err;
    if (generate_partition(spec, nranks, rank, apply_ewt_too != 0, loc, err) != 0) { g_err = err; return MGCFD_ERR_ARG; }
    long nglob = spec.dims[level][0] * spec.dims[level][1] * spec.dims[level][2] * (kind == 2 ? 6 : 1);
    fill_plan_outputs(loc.levels[level], nglob, nranks, info, gid, send_counts, recv_counts, send_gids);
    return MGCFD_OK;
}
int mgcfd_generate_upload_partition(mgcfd_ctx* c, int kind, int levels, const long* dims, const double lengths[3], int mesh_variant, double tilt) {
    if (!c) { g_err = "null context"; return MGCFD_ERR_ARG; }
    if (!c->dist.active) { g_err = "mgcfd_dist_init has not been called"; return MGCFD_ERR_ARG; }
    if (levels != c->levels || mesh_variant != c->variant) { g_err = "mesh does not match the context"; return MGCFD_ERR_ARG; }
    MeshSpec spec;
    CKRC(make_spec(kind, levels, dims, lengths, mesh_variant, tilt, spec));
    LocalMesh loc;
    std::string err;
    if (generate_partition(spec, c->dist.nranks, c->dist.rank, true, loc, err) != 0) { g_err = err; return MGCFD_ERR_ARG; }
    for (int l = 0; l < c->levels; l++) {
        c->L[l].host.n_owned = loc.levels[l].n_owned;
        const long nglob = spec.dims[l][0] * spec.dims[l][1] * spec.dims[l][2] * (kind == 2 ? 6 : 1);
        CKRC(upload_local_level(c, l, loc.levels[l], nglob));
    }
    return mgcfd_finalize(c);
}

void mgcfd_free(void* p) { free(p); }

}  // extern "C"
