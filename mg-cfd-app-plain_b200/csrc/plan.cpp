// plan.cpp -- integer preprocessing (renumbering, tiling, colouring, MG operators). See plan.h.
#include "plan.h"

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <future>
#include <numeric>
#include <stdexcept>
#include <string>
#include <thread>

namespace mgcfd {

namespace {

struct Csr {
    std::vector<long> off;
    std::vector<int> nbr;
    std::vector<int> eid;    // edge index of every entry (ascending per node)
};

// host threads the preprocessing may use: hardware concurrency, capped by MGCFD_PLAN_THREADS (1 = serial; the result never depends on it)
unsigned plan_threads() {
    unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    if (const char* e = getenv("MGCFD_PLAN_THREADS")) { const int v = atoi(e); if (v >= 1) hw = std::min<unsigned>(hw, unsigned(v)); }
    return hw;
}

// f(i0, i1) over [0, n) in contiguous pieces on the preprocessing threads
template <class F>
void parallel_ranges(long n, long grain, F f) {
    const long pieces = std::max<long>(1, std::min<long>(plan_threads(), n / std::max<long>(grain, 1)));
    if (pieces <= 1) { f(0L, n); return; }
    std::vector<std::future<void>> pool;
    for (long k = 1; k < pieces; k++) pool.push_back(std::async(std::launch::async, [=] { f(n * k / pieces, n * (k + 1) / pieces); }));
    f(0L, n / pieces);
    for (auto& x : pool) x.get();
}

// MGCFD_PLAN_TIMING=1: wall time of the phases of build_level_plan on stderr (setup cost of large meshes)
struct PhaseClock {
    bool on = getenv("MGCFD_PLAN_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "  plan: %-28s %8.3f s\n", what, std::chrono::duration<double>(now - t).count());
        t = now;
    }
};

// undirected adjacency of the internal edges, in OLD ids; per node, entries in ascending edge index
Csr adjacency_old(const HostLevel& L) {
    Csr g;
    g.off.assign(L.nel + 1, 0);
    for (long e = 0; e < L.nI; e++) { g.off[L.edges[e].a + 1]++; g.off[L.edges[e].b + 1]++; }
    for (long i = 0; i < L.nel; i++) g.off[i + 1] += g.off[i];
    g.nbr.resize(g.off[L.nel]);
    g.eid.resize(g.off[L.nel]);
    std::vector<long> pos(g.off.begin(), g.off.end() - 1);
    for (long e = 0; e < L.nI; e++) {
        const long a = L.edges[e].a, b = L.edges[e].b;
        g.eid[pos[a]] = int(e); g.nbr[pos[a]++] = int(b);
        g.eid[pos[b]] = int(e); g.nbr[pos[b]++] = int(a);
    }
    return g;
}

// Breadth-first (Cuthill-McKee) order of the nodes for which member(i) holds, appended to `order`.
// Neighbours are visited by (degree, id). `visited` is a stamp array.
template <class Member>
void cm_order(const Csr& g, const std::vector<long>& nodes, Member member, std::vector<int>& stamp, int mark,
              std::vector<long>& order) {
    auto degree = [&](long i) { return g.off[i + 1] - g.off[i]; };
    std::vector<long> frontier_nbrs;
    // seeds by (degree, id): low-degree nodes first, as in classical CM
    std::vector<long> seeds(nodes);
    std::sort(seeds.begin(), seeds.end(), [&](long x, long y) {
        const long dx = degree(x), dy = degree(y);
        return dx != dy ? dx < dy : x < y;
    });
    for (long seed : seeds) {
        if (stamp[seed] == mark) continue;
        // pseudo-peripheral start: two BFS sweeps from the seed inside the member set
        long start = seed;
        for (int sweep = 0; sweep < 2; sweep++) {
            std::vector<long> q{start};
            std::vector<long> touched{start};
            stamp[start] = -mark;
            size_t head = 0;
            while (head < q.size()) {
                long u = q[head++];
                for (long k = g.off[u]; k < g.off[u + 1]; k++) {
                    long v = g.nbr[k];
                    if (!member(v) || stamp[v] == mark || stamp[v] == -mark) continue;
                    stamp[v] = -mark;
                    q.push_back(v);
                    touched.push_back(v);
                }
            }
            // last level: pick the lowest-degree node of the final BFS layer (approximate: last visited)
            start = q.back();
            for (long v : touched) stamp[v] = 0;
        }
        size_t head = order.size();
        order.push_back(start);
        stamp[start] = mark;
        while (head < order.size()) {
            long u = order[head++];
            frontier_nbrs.clear();
            for (long k = g.off[u]; k < g.off[u + 1]; k++) {
                long v = g.nbr[k];
                if (!member(v) || stamp[v] == mark) continue;
                stamp[v] = mark;
                frontier_nbrs.push_back(v);
            }
            std::sort(frontier_nbrs.begin(), frontier_nbrs.end(), [&](long x, long y) {
                const long dx = degree(x), dy = degree(y);
                return dx != dy ? dx < dy : x < y;
            });
            order.insert(order.end(), frontier_nbrs.begin(), frontier_nbrs.end());
        }
    }
}

struct Bisector {
    const HostLevel& L;
    const Csr& g;
    long TN;
    std::vector<long>& tile_of;
    std::vector<int> sub;     // subset stamp for the graph variant
    std::vector<int> stamp;
    int next_mark = 1;
    Bisector(const HostLevel& L_, const Csr& g_, long tn, std::vector<long>& t) : L(L_), g(g_), TN(tn), tile_of(t) {}

    static long left_count(long n, long k, long kl, long TN) {
        long nl = (n * kl + k / 2) / k;
        nl = std::min(nl, kl * TN);
        nl = std::max(nl, n - (k - kl) * TN);
        return nl;
    }
    // recursive coordinate bisection: split the widest extent at the proportional rank
    void rcb(long* idx, long n, long k, long tile0, int depth) {
        if (k == 1) { for (long i = 0; i < n; i++) tile_of[idx[i]] = tile0; return; }
        const long kl = k / 2, nl = left_count(n, k, kl, TN);
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (long i = 0; i < n; i++) for (int d = 0; d < 3; d++) {
            const double c = L.coords[3 * idx[i] + d];
            lo[d] = std::min(lo[d], c); hi[d] = std::max(hi[d], c);
        }
        int ax = 0;
        for (int d = 1; d < 3; d++) if (hi[d] - lo[d] > hi[ax] - lo[ax]) ax = d;
        const double* c = L.coords.data();
        std::nth_element(idx, idx + nl, idx + n, [c, ax](long x, long y) {
            const double cx = c[3 * x + ax], cy = c[3 * y + ax];
            return cx != cy ? cx < cy : x < y;
        });
        if (depth < 3 && n > 200000 && plan_threads() > 1) {
            auto fut = std::async(std::launch::async, [&] { rcb(idx, nl, kl, tile0, depth + 1); });
            rcb(idx + nl, n - nl, k - kl, tile0 + kl, depth + 1);
            fut.get();
        } else {
            rcb(idx, nl, kl, tile0, depth + 1);
            rcb(idx + nl, n - nl, k - kl, tile0 + kl, depth + 1);
        }
    }
    // ---- two-level bisection (PlanOptions::supers): first into super-tiles [s0, s0 + k) whose capacities are whole numbers of
    // tiles (cap[s] .. cap[s+1], prefix sums in tiles), then every super-tile into its own tiles -- a super-tile is a compact
    // block of consecutive tiles, which keeps its halo (rows outside the block that its edges touch) small
    const std::vector<long>* cap = nullptr;
    long left_count_cap(long n, long s0, long k, long kl) const {
        const long capL = ((*cap)[s0 + kl] - (*cap)[s0]) * TN, capT = ((*cap)[s0 + k] - (*cap)[s0]) * TN, capR = capT - capL;
        long nl = (long)(((__int128)n * capL + capT / 2) / capT);
        nl = std::min(nl, capL);
        nl = std::max(nl, n - capR);
        return nl;
    }
    void split_widest(long* idx, long n, long nl) {
        double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
        for (long i = 0; i < n; i++) for (int d = 0; d < 3; d++) {
            const double c = L.coords[3 * idx[i] + d];
            lo[d] = std::min(lo[d], c); hi[d] = std::max(hi[d], c);
        }
        int ax = 0;
        for (int d = 1; d < 3; d++) if (hi[d] - lo[d] > hi[ax] - lo[ax]) ax = d;
        const double* c = L.coords.data();
        std::nth_element(idx, idx + nl, idx + n, [c, ax](long x, long y) {
            const double cx = c[3 * x + ax], cy = c[3 * y + ax];
            return cx != cy ? cx < cy : x < y;
        });
    }
    void rcb_super(long* idx, long n, long s0, long k, int depth) {
        if (k == 1) { rcb(idx, n, (*cap)[s0 + 1] - (*cap)[s0], (*cap)[s0], depth); return; }
        const long kl = k / 2, nl = left_count_cap(n, s0, k, kl);
        split_widest(idx, n, nl);
        if (depth < 3 && n > 200000 && plan_threads() > 1) {
            auto fut = std::async(std::launch::async, [&] { rcb_super(idx, nl, s0, kl, depth + 1); });
            rcb_super(idx + nl, n - nl, s0 + kl, k - kl, depth + 1);
            fut.get();
        } else {
            rcb_super(idx, nl, s0, kl, depth + 1);
            rcb_super(idx + nl, n - nl, s0 + kl, k - kl, depth + 1);
        }
    }
    void gbis_super(long* idx, long n, long s0, long k) {
        if (k == 1) { gbis(idx, n, (*cap)[s0 + 1] - (*cap)[s0], (*cap)[s0]); return; }
        const long kl = k / 2, nl = left_count_cap(n, s0, k, kl);
        if (sub.empty()) { sub.assign(L.nel, 0); stamp.assign(L.nel, 0); }
        const int s = next_mark++;
        for (long i = 0; i < n; i++) sub[idx[i]] = s;
        std::vector<long> nodes(idx, idx + n), order;
        order.reserve(n);
        const int mark = next_mark++;
        cm_order(g, nodes, [&](long v) { return sub[v] == s; }, stamp, mark, order);
        std::copy(order.begin(), order.end(), idx);
        gbis_super(idx, nl, s0, kl);
        gbis_super(idx + nl, n - nl, s0 + kl, k - kl);
    }
    // graph bisection: BFS order of the subset from a pseudo-peripheral node, first nl nodes go left
    void gbis(long* idx, long n, long k, long tile0) {
        if (k == 1) { for (long i = 0; i < n; i++) tile_of[idx[i]] = tile0; return; }
        const long kl = k / 2, nl = left_count(n, k, kl, TN);
        if (sub.empty()) { sub.assign(L.nel, 0); stamp.assign(L.nel, 0); }
        const int s = next_mark++;
        for (long i = 0; i < n; i++) sub[idx[i]] = s;
        std::vector<long> nodes(idx, idx + n), order;
        order.reserve(n);
        const int mark = next_mark++;
        cm_order(g, nodes, [&](long v) { return sub[v] == s; }, stamp, mark, order);
        std::copy(order.begin(), order.end(), idx);
        gbis(idx, nl, kl, tile0);
        gbis(idx + nl, n - nl, k - kl, tile0 + kl);
    }
};

// byte offset of chunk 0 of row o in a tile's shared record buffer (64-byte rows, 64B swizzle)
inline uint16_t row_code(int o) { return uint16_t((o << 6) | (((o >> 1) & 3) << 4)); }
inline int row_of_code(uint16_t c) { return c >> 6; }

struct Mask256 {
    uint64_t w[4] = {0, 0, 0, 0};
    inline void set(int c) { w[c >> 6] |= (1ull << (c & 63)); }
    static inline int first_free(const Mask256& a, const Mask256& b) {
        for (int k = 0; k < 4; k++) {
            uint64_t m = ~(a.w[k] | b.w[k]);
            if (m) return 64 * k + __builtin_ctzll(m);
        }
        throw std::runtime_error("mgcfd: more than 256 colours needed in one tile");
    }
};

}  // namespace

static void build_visit_streams(const HostLevel& L, const PlanOptions& opt, LevelPlan& P, const std::vector<int>& adj_eid);

void build_level_plan(const HostLevel& L, const PlanOptions& opt, LevelPlan& P) {
    const long nall = L.nel;                                   // owned + ghosts
    const long n = L.n_owned >= 0 ? L.n_owned : L.nel;         // owned: only these are tiled, ordered and computed
    const long TN = opt.tile_nodes;
    P = LevelPlan();
    P.nel = nall; P.nI = L.nI; P.nB = L.nB; P.nW = L.nW; P.TN = int(TN);
    P.n_owned = n;
    P.ntiles = std::max<long>(1, (n + TN - 1) / TN);
    P.npad_owned = P.ntiles * TN;
    P.npad = P.npad_owned + (((nall - n) + 7) & ~7L);
    if (P.npad > 0x7fffffffL) throw std::runtime_error("mgcfd: level too large for 32-bit node ids");
    PhaseClock clk;
    const Csr g = adjacency_old(L);
    clk.lap("adjacency");

    // ---- 1. node -> tile, and order inside tiles --------------------------------------------------
    std::vector<long> tile_of(n, 0);
    std::vector<long> seq(n);    // global visiting sequence used to rank nodes inside a tile
    if (opt.ordering == 0) {
        for (long i = 0; i < n; i++) { tile_of[i] = i / TN; seq[i] = i; }
    } else if (opt.ordering == 1) {
        std::vector<long> nodes(n), order;
        std::iota(nodes.begin(), nodes.end(), 0L);
        std::vector<int> stamp(nall, 0);
        order.reserve(n);
        cm_order(g, nodes, [n](long v) { return v < n; }, stamp, 1, order);
        std::reverse(order.begin(), order.end());   // reverse Cuthill-McKee
        for (long r = 0; r < n; r++) { tile_of[order[r]] = r / TN; seq[order[r]] = r; }
    } else {
        std::vector<long> idx(n);
        std::iota(idx.begin(), idx.end(), 0L);
        Bisector B(L, g, TN, tile_of);
        std::vector<long> cap;
        if (opt.supers > 0 && !opt.scatter && TN == 128) {
            // super-tile s holds ntiles / ns tiles, the first ntiles % ns of them one more
            const long ns = std::min<long>(opt.supers, P.ntiles);
            cap.assign(ns + 1, 0);
            for (long s = 0; s < ns; s++) cap[s + 1] = cap[s] + P.ntiles / ns + (s < P.ntiles % ns ? 1 : 0);
            B.cap = &cap;
            if (!L.coords.empty()) B.rcb_super(idx.data(), n, 0, ns, 0);
            else B.gbis_super(idx.data(), n, 0, ns);
            P.visit.ns = int(ns);
            P.visit.super_off = cap;
        }
        else if (!L.coords.empty()) B.rcb(idx.data(), n, P.ntiles, 0, 0);
        else B.gbis(idx.data(), n, P.ntiles, 0);
        clk.lap("bisection into tiles");
        // Cuthill-McKee inside every tile
        std::vector<long> toff(P.ntiles + 1, 0);
        for (long i = 0; i < n; i++) toff[tile_of[i] + 1]++;
        for (long t = 0; t < P.ntiles; t++) toff[t + 1] += toff[t];
        std::vector<long> members(n), pos(toff.begin(), toff.end() - 1);
        for (long i = 0; i < n; i++) members[pos[tile_of[i]]++] = i;
        std::vector<int> stamp(nall, 0);
        // tiles are disjoint and cm_order only touches the stamps of its own tile's nodes (stamps are tile specific, t+1, so no
        // reset is needed): tiles are ordered concurrently; tile t fills the sequence numbers toff[t] .. toff[t+1]
        std::atomic<long> next_tile(0);
        auto worker = [&]() {
            std::vector<long> order, nodes;
            for (long t0 = next_tile.fetch_add(64); t0 < P.ntiles; t0 = next_tile.fetch_add(64))
                for (long t = t0; t < std::min(t0 + 64, P.ntiles); t++) {
                    nodes.assign(members.begin() + toff[t], members.begin() + toff[t + 1]);
                    order.clear();
                    cm_order(g, nodes, [&](long v) { return v < n && tile_of[v] == t; }, stamp, int(t % 1000000000) + 1, order);
                    long r = toff[t];
                    for (long v : order) seq[v] = r++;
                }
        };
        const unsigned nthreads = unsigned(std::max<long>(1, std::min<long>(plan_threads(), P.ntiles / 64)));
        std::vector<std::future<void>> pool;
        for (unsigned k = 1; k < nthreads; k++) pool.push_back(std::async(std::launch::async, worker));
        worker();
        for (auto& f : pool) f.get();
    }
    clk.lap("ordering inside tiles");
    // rank inside tile by seq
    P.new_of_old.assign(nall, -1);
    P.old_of_new.assign(P.npad, -1);
    for (long i = n; i < nall; i++) { P.new_of_old[i] = P.npad_owned + (i - n); P.old_of_new[P.npad_owned + (i - n)] = i; }
    P.tile_nown.assign(P.ntiles, 0);
    {
        std::vector<long> byseq(n);                  // seq is a permutation of 0..n-1: its inverse lists the nodes in sequence
        for (long i = 0; i < n; i++) byseq[seq[i]] = i;
        for (long i : byseq) {
            const long t = tile_of[i];
            const long id = t * TN + P.tile_nown[t]++;
            P.new_of_old[i] = id;
            P.old_of_new[id] = i;
        }
        for (long t = 0; t < P.ntiles; t++)
            if (P.tile_nown[t] > TN) throw std::runtime_error("mgcfd: internal error, tile overflow");
    }

    clk.lap("renumbering");
    // ---- 2. flat edge list + CSR by node in new ids (entries in ascending original edge index) ------
    P.ea.resize(L.nI); P.eb.resize(L.nI); P.ew.resize(3 * L.nI);
    parallel_ranges(L.nI, 1 << 16, [&](long e0, long e1) {
        for (long e = e0; e < e1; e++) {
            P.ea[e] = int(P.new_of_old[L.edges[e].a]);
            P.eb[e] = int(P.new_of_old[L.edges[e].b]);
            P.ew[e] = L.edges[e].x; P.ew[L.nI + e] = L.edges[e].y; P.ew[2 * L.nI + e] = L.edges[e].z;
        }
    });
    const long nbw = L.nB + L.nW;
    P.bnode.resize(nbw); P.bkind.resize(nbw); P.bw.resize(3 * nbw);
    for (long k = 0; k < nbw; k++) {
        const EdgeNb& e = L.edges[L.nI + k];
        P.bnode[k] = int(P.new_of_old[e.b]);
        P.bkind[k] = uint8_t(k < L.nB ? 1 : 2);
        P.bw[k] = e.x; P.bw[nbw + k] = e.y; P.bw[2 * nbw + k] = e.z;
    }
    // CSR by node in new ids: the old-id adjacency `g` (entries in ascending edge index) carried over node by node; bit 31 of an
    // entry marks that this node is the edge's `b` end
    P.adj_off.assign(P.npad + 1, 0);
    for (long i = 0; i < P.npad; i++) {
        const long on = P.old_of_new[i];
        P.adj_off[i + 1] = P.adj_off[i] + (on >= 0 ? g.off[on + 1] - g.off[on] : 0);
    }
    P.adj_nbr.resize(2 * L.nI);
    std::vector<int> adj_eid(2 * L.nI);
    parallel_ranges(P.npad, 1 << 14, [&](long i0, long i1) {
        for (long i = i0; i < i1; i++) {
            const long on = P.old_of_new[i];
            if (on < 0) continue;
            long o = P.adj_off[i];
            for (long k = g.off[on]; k < g.off[on + 1]; k++, o++) {
                const int e = g.eid[k];
                // the node is the edge's `b` end (a self-loop lists its `a` entry first)
                const bool is_b = L.edges[e].a != on || (k > g.off[on] && g.eid[k - 1] == e);
                const int nb = int(P.new_of_old[g.nbr[k]]);
                P.adj_nbr[o] = is_b ? int(uint32_t(nb) | 0x80000000u) : nb;
                adj_eid[o] = e;
            }
        }
    });
    // boundary/wall edges per node (CSR, original order)
    std::vector<long> bn_off(P.npad + 1, 0);
    for (long k = 0; k < nbw; k++) bn_off[P.bnode[k] + 1]++;
    for (long i = 0; i < P.npad; i++) bn_off[i + 1] += bn_off[i];
    std::vector<long> bn_idx(nbw);
    {
        std::vector<long> pos(bn_off.begin(), bn_off.end() - 1);
        for (long k = 0; k < nbw; k++) bn_idx[pos[P.bnode[k]]++] = k;
    }

    clk.lap("edge lists + CSR");
    // ---- 3. tiles: halo lists, edge rounds ------------------------------------------------------------
    // Tiles are independent: contiguous ranges of tiles are processed by a pool of host threads, each range into its own buffers,
    // which are then laid end to end in tile order -- byte for byte what a serial pass produces (tests compare the plan's hash).
    P.scatter = opt.scatter;
    P.halo_off.assign(P.ntiles + 1, 0);
    P.slot_off.assign(P.ntiles + 1, 0);
    P.bslot_off.assign(P.ntiles + 1, 0);
    struct Slot { int owner; int round; int other; long e; bool owner_is_a; };
    const size_t BLK = size_t(TN) * 26, BBLK = size_t(TN) * 25;
    struct TileRange {
        long t0 = 0, t1 = 0;
        std::vector<unsigned char> slots, bslots;      // this range's round blocks / boundary blocks, tile after tile
        std::vector<int> halo;                         // this range's halo ids, tile after tile
        std::vector<int> rounds, brounds, nhalo;       // per tile
        long cut_edges = 0, used_slots = 0;
        int max_halo = 0, max_rounds = 0;
        bool oversize = false;
        std::string error;
    };
    // segment rounds: the ROUND in which each edge of a node is taken (see process_range): greedy list schedule per group of 8
    // consecutive nodes so that the 8 lanes of a quarter-warp read rows with distinct (row mod 8)
    auto schedule_rounds = [](std::vector<Slot>& slots, int tile_rounds) {
                size_t s0 = 0;
                while (s0 < slots.size()) {
                    const int g = slots[s0].owner / 8;
                    size_t s1 = s0;
                    while (s1 < slots.size() && slots[s1].owner / 8 == g) s1++;
                    // remaining edges per lane
                    std::vector<size_t> lane_edges[8];
                    for (size_t k = s0; k < s1; k++) lane_edges[slots[k].owner % 8].push_back(k);
                    for (int r = 0; r < tile_rounds; r++) {
                        int used_by[8];                       // residue -> row using it in this round (-1 free)
                        for (int q = 0; q < 8; q++) used_by[q] = -1;
                        int order[8] = {0, 1, 2, 3, 4, 5, 6, 7};
                        std::stable_sort(order, order + 8, [&](int x, int y) { return lane_edges[x].size() > lane_edges[y].size(); });
                        const int rounds_left = tile_rounds - r;
                        for (int oi = 0; oi < 8; oi++) {
                            auto& le = lane_edges[order[oi]];
                            if (le.empty()) continue;
                            int pick = -1;
                            for (size_t c = 0; c < le.size(); c++) {
                                const int row = slots[le[c]].other;
                                if (used_by[row & 7] == -1 || used_by[row & 7] == row) { pick = int(c); break; }
                            }
                            if (pick < 0) {
                                if (int(le.size()) < rounds_left) continue;     // can wait: leave this round empty
                                pick = 0;                                        // must go now: accept the conflict
                            }
                            const size_t k = le[pick];
                            slots[k].round = r;
                            used_by[slots[k].other & 7] = slots[k].other;
                            le.erase(le.begin() + pick);
                        }
                    }
                    s0 = s1;
                }
    };
    auto process_range = [&](TileRange& C) {
    std::vector<Slot> slots;
    std::vector<int> halo;
    std::vector<Mask256> Lm(TN), Rm(TN);
    std::vector<int> nassigned(TN);
    long blocks_done = 0, bblocks_done = 0;
    for (long t = C.t0; t < C.t1; t++) {
        const long base = t * TN;
        const int nown = P.tile_nown[t];
        halo.clear();
        for (int lu = 0; lu < nown; lu++)
            for (long k = P.adj_off[base + lu]; k < P.adj_off[base + lu + 1]; k++) {
                const int v = P.adj_nbr[k] & 0x7fffffff;
                if (v < base || v >= base + TN) halo.push_back(v);
            }
        std::sort(halo.begin(), halo.end());
        halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
        if (long(TN) + long(halo.size()) >= 1020) {
            // the 16-bit row code addresses at most 1020 rows of 64 bytes; a tile this scattered could not be staged anyway
            if (opt.strict) throw std::runtime_error("mgcfd: tile halo too large (TN + halo must stay below 1020 rows): use a smaller tile_nodes or a locality-preserving ordering (MGCFD_ORDER_PARTITION_RCM)");
            C.oversize = true;
        }
        C.nhalo.push_back(int(halo.size()));
        C.halo.insert(C.halo.end(), halo.begin(), halo.end());
        C.max_halo = std::max(C.max_halo, int(halo.size()));
        auto local_of = [&](int v) -> int {
            if (v >= base && v < base + TN) return int(v - base);
            return int(TN) + int(std::lower_bound(halo.begin(), halo.end(), v) - halo.begin());
        };

        slots.clear();
        if (!opt.scatter) {
            // sorted-segment rounds: every node lists all its incident edges (ascending original edge index) ...
            int tile_rounds = 0;
            for (int lu = 0; lu < nown; lu++) {
                int r = 0;
                for (long k = P.adj_off[base + lu]; k < P.adj_off[base + lu + 1]; k++, r++) {
                    const int v = P.adj_nbr[k] & 0x7fffffff;
                    const bool cut = (v < base || v >= base + TN);
                    slots.push_back({lu, r, local_of(v), adj_eid[k], P.adj_nbr[k] >= 0});
                    if (cut) C.cut_edges++;
                }
                tile_rounds = std::max(tile_rounds, r);
            }
            // ... and the ROUND in which each edge of a node is taken is then chosen so that the 8 lanes of a quarter-warp
            // (which share one 128-bit shared-memory transaction) read rows with distinct (row mod 8), i.e. distinct bank
            // groups under the 64B swizzle: a greedy list schedule per group of 8 consecutive nodes.  Only the order of a
            // node's additions changes (still a fixed, reproducible order).
            if (opt.conflict_free_rounds) schedule_rounds(slots, tile_rounds);
        } else {
            for (int i = 0; i < nown; i++) { Lm[i] = Mask256(); Rm[i] = Mask256(); nassigned[i] = 0; }
            // internal-to-tile edges first (visited from their `a` end, ascending edge index per node)
            for (int lu = 0; lu < nown; lu++)
                for (long k = P.adj_off[base + lu]; k < P.adj_off[base + lu + 1]; k++) {
                    if (P.adj_nbr[k] < 0) continue;                   // this node is the `b` end: handled from `a`
                    const int v = P.adj_nbr[k];
                    if (v < base || v >= base + TN) continue;
                    const int lv = int(v - base);
                    const int c1 = Mask256::first_free(Lm[lu], Rm[lv]);   // lu computes, scatters into lv
                    const int c2 = Mask256::first_free(Lm[lv], Rm[lu]);
                    bool pick_u = c1 < c2 || (c1 == c2 && nassigned[lu] <= nassigned[lv]);
                    if (pick_u) { Lm[lu].set(c1); Rm[lv].set(c1); nassigned[lu]++; slots.push_back({lu, c1, lv, adj_eid[k], true}); }
                    else        { Lm[lv].set(c2); Rm[lu].set(c2); nassigned[lv]++; slots.push_back({lv, c2, lu, adj_eid[k], false}); }
                }
            // cut edges: computed by the owned end, the halo end is read-only
            static const Mask256 none;
            for (int lu = 0; lu < nown; lu++)
                for (long k = P.adj_off[base + lu]; k < P.adj_off[base + lu + 1]; k++) {
                    const int v = P.adj_nbr[k] & 0x7fffffff;
                    if (v >= base && v < base + TN) continue;
                    const int c = Mask256::first_free(Lm[lu], none);
                    Lm[lu].set(c); nassigned[lu]++;
                    slots.push_back({lu, c, local_of(v), adj_eid[k], P.adj_nbr[k] >= 0});
                    C.cut_edges++;
                }
        }
        int rounds = 0;
        for (const Slot& s : slots) rounds = std::max(rounds, s.round + 1);
        C.max_rounds = std::max(C.max_rounds, rounds);
        C.rounds.push_back(rounds);
        const long b0 = blocks_done;
        blocks_done += rounds;
        C.slots.resize(size_t(blocks_done) * BLK, 0);
        // empty slots
        for (int r = 0; r < rounds; r++) {
            uint16_t* oth = reinterpret_cast<uint16_t*>(C.slots.data() + size_t(b0 + r) * BLK + size_t(TN) * 24);
            for (int lu = 0; lu < TN; lu++) oth[lu] = opt.scatter ? uint16_t(0xFFFF) : row_code(lu);
        }
        for (const Slot& s : slots) {
            unsigned char* blk = C.slots.data() + size_t(b0 + s.round) * BLK;
            double* w = reinterpret_cast<double*>(blk);
            uint16_t* oth = reinterpret_cast<uint16_t*>(blk + size_t(TN) * 24);
            const double sg = s.owner_is_a ? -0.5 : 0.5;      // h = -0.5 * vector(thread-node -> other); stored vector is a -> b
            w[s.owner] = sg * P.ew[s.e];
            w[TN + s.owner] = sg * P.ew[L.nI + s.e];
            w[2 * TN + s.owner] = sg * P.ew[2 * L.nI + s.e];
            oth[s.owner] = row_code(s.other);
        }
        C.used_slots += long(slots.size());
        // boundary / wall rounds: round r of a node = its r-th boundary or wall edge in original order
        int br = 0;
        for (int lu = 0; lu < nown; lu++) br = std::max<int>(br, int(bn_off[base + lu + 1] - bn_off[base + lu]));
        C.brounds.push_back(br);
        const long bb0 = bblocks_done;
        bblocks_done += br;
        C.bslots.resize(size_t(bblocks_done) * BBLK, 0);
        for (int lu = 0; lu < nown; lu++) {
            int r = 0;
            for (long k = bn_off[base + lu]; k < bn_off[base + lu + 1]; k++, r++) {
                const long bi = bn_idx[k];
                unsigned char* blk = C.bslots.data() + size_t(bb0 + r) * BBLK;
                double* w = reinterpret_cast<double*>(blk);
                w[lu] = P.bw[bi]; w[TN + lu] = P.bw[nbw + bi]; w[2 * TN + lu] = P.bw[2 * nbw + bi];
                blk[size_t(TN) * 24 + lu] = P.bkind[bi];
            }
        }
    }
    };   // process_range
    {
        const unsigned hw = plan_threads();
        const long nranges = std::max<long>(1, std::min<long>(P.ntiles, P.ntiles < 64 ? 1 : 8L * hw));
        std::vector<TileRange> ranges(nranges);
        for (long r = 0; r < nranges; r++) { ranges[r].t0 = P.ntiles * r / nranges; ranges[r].t1 = P.ntiles * (r + 1) / nranges; }
        std::atomic<long> next(0);
        auto worker = [&]() {
            for (long r = next.fetch_add(1); r < nranges; r = next.fetch_add(1)) {
                try { process_range(ranges[r]); }
                catch (const std::exception& ex) { ranges[r].error = ex.what(); if (ranges[r].error.empty()) ranges[r].error = "error"; }
            }
        };
        const unsigned nthreads = unsigned(std::min<long>(hw, nranges));
        std::vector<std::future<void>> pool;
        for (unsigned k = 1; k < nthreads; k++) pool.push_back(std::async(std::launch::async, worker));
        worker();
        for (auto& f : pool) f.get();
        for (const TileRange& C : ranges) if (!C.error.empty()) throw std::runtime_error(C.error);     // the first failing range, as a serial pass would
        // lay the ranges end to end
        for (const TileRange& C : ranges)
            for (long t = C.t0; t < C.t1; t++) {
                P.halo_off[t + 1] = P.halo_off[t] + C.nhalo[t - C.t0];
                P.slot_off[t + 1] = P.slot_off[t] + C.rounds[t - C.t0];
                P.bslot_off[t + 1] = P.bslot_off[t] + C.brounds[t - C.t0];
            }
        P.halo_ids.resize(size_t(P.halo_off[P.ntiles]));
        P.slots.resize(size_t(P.slot_off[P.ntiles]) * BLK);
        P.bslots.resize(size_t(P.bslot_off[P.ntiles]) * BBLK);
        next.store(0);
        auto copier = [&]() {
            for (long r = next.fetch_add(1); r < nranges; r = next.fetch_add(1)) {
                TileRange& C = ranges[r];
                if (!C.halo.empty()) memcpy(P.halo_ids.data() + P.halo_off[C.t0], C.halo.data(), C.halo.size() * sizeof(int));
                if (!C.slots.empty()) memcpy(P.slots.data() + size_t(P.slot_off[C.t0]) * BLK, C.slots.data(), C.slots.size());
                if (!C.bslots.empty()) memcpy(P.bslots.data() + size_t(P.bslot_off[C.t0]) * BBLK, C.bslots.data(), C.bslots.size());
                std::vector<unsigned char>().swap(C.slots); std::vector<unsigned char>().swap(C.bslots); std::vector<int>().swap(C.halo);
            }
        };
        pool.clear();
        for (unsigned k = 1; k < nthreads; k++) pool.push_back(std::async(std::launch::async, copier));
        copier();
        for (auto& f : pool) f.get();
        for (const TileRange& C : ranges) {
            P.cut_edges += C.cut_edges; P.used_slots += C.used_slots;
            P.max_halo = std::max(P.max_halo, C.max_halo); P.max_rounds = std::max(P.max_rounds, C.max_rounds);
            P.oversize = P.oversize || C.oversize;
        }
    }
    clk.lap("tile rounds");
    // fixed-stride tile headers (+ halo ids) for the pipelined stage kernel
    P.hpad = (P.max_halo + 3) & ~3;
    P.hdr_stride = 32 + 4 * P.hpad;
    P.hdrs.assign(size_t(P.ntiles) * P.hdr_stride, 0);
    for (long t = 0; t < P.ntiles; t++) {
        unsigned char* h = P.hdrs.data() + size_t(t) * P.hdr_stride;
        int* hi = reinterpret_cast<int*>(h);
        long long* hl = reinterpret_cast<long long*>(h + 16);
        hi[0] = int(P.slot_off[t + 1] - P.slot_off[t]);
        hi[1] = int(P.halo_off[t + 1] - P.halo_off[t]);
        hi[2] = int(P.bslot_off[t + 1] - P.bslot_off[t]);
        hl[0] = P.slot_off[t]; hl[1] = P.bslot_off[t];
        int* ids = reinterpret_cast<int*>(h + 32);
        for (long k = P.halo_off[t]; k < P.halo_off[t + 1]; k++) ids[k - P.halo_off[t]] = P.halo_ids[k];
    }
    if (P.visit.ns > 0) { build_visit_streams(L, opt, P, adj_eid); clk.lap("visit streams"); }
}

namespace {
inline uint16_t visit_code(int idx, bool halo) { return uint16_t(((idx << 2) | ((idx >> 1) & 3)) | (halo ? 0x8000 : 0)); }
}

// The visit kernel's streams (VisitPlan): per super-tile the halo list and descriptor, per warp-tile (32 rows) the edge rounds with
// the other endpoint addressed inside the super-tile.  Called at the end of build_level_plan when PlanOptions::supers > 0.
static void build_visit_streams(const HostLevel& L, const PlanOptions& opt, LevelPlan& P, const std::vector<int>& adj_eid) {
    VisitPlan& V = P.visit;
    const long TN = P.TN;
    const long ns = V.ns;
    const int WT = 32;                              // rows of a warp-tile
    const size_t BLK = size_t(WT) * 34;
    const double k2 = 2.0 * (-0.5 * double(0.2f));      // 2 * kdiss, kdiss = -0.5 * smoothing_coefficient (src/Base/common.h:24)
    const int RC = std::max(1, opt.visit_rounds);       // rounds per ring chunk
    V.hsum.assign(3 * size_t(P.npad_owned), 0.0);
    struct VSlot { int owner; int round; int other; long e; bool owner_is_a; };     // other = idx | halo << 20
    struct Ent { int orow0, rounds, blane0; long tile; long blk0; };                // blk0 relative to the super-tile's first block
    struct SuperOut {
        std::vector<int> halo;
        std::vector<Ent> ents;
        std::vector<unsigned char> blocks;        // the super-tile's round blocks, warp-tile after warp-tile
        std::string error;
    };
    std::vector<SuperOut> out(ns);
    auto process_super = [&](long s) {
        SuperOut& O = out[s];
        const long t0 = V.super_off[s], t1 = V.super_off[s + 1];
        const long row0 = t0 * TN, row1 = t1 * TN;
        std::vector<int>& halo = O.halo;
        for (long t = t0; t < t1; t++)
            for (int lu = 0; lu < P.tile_nown[t]; lu++)
                for (long k = P.adj_off[t * TN + lu]; k < P.adj_off[t * TN + lu + 1]; k++) {
                    const int v = P.adj_nbr[k] & 0x7fffffff;
                    if (v < row0 || v >= row1) halo.push_back(v);
                }
        std::sort(halo.begin(), halo.end());
        halo.erase(std::unique(halo.begin(), halo.end()), halo.end());
        if (row1 - row0 >= 8192 || long(halo.size()) >= 8192) { O.error = "mgcfd: super-tile too large for the 13-bit row index of the visit kernel"; return; }
        std::vector<VSlot> slots;
        long blocks_done = 0;
        for (long t = t0; t < t1; t++)
            for (int w4 = 0; w4 < TN / WT; w4++) {
                const long base = t * TN + w4 * WT;
                const int nown = std::max(0, std::min(WT, P.tile_nown[t] - w4 * WT));
                if (nown == 0) continue;                  // nothing but padding rows: no work, no entry
                slots.clear();
                int wt_rounds = 0;
                for (int lu = 0; lu < nown; lu++) {
                    int r = 0;
                    for (long k = P.adj_off[base + lu]; k < P.adj_off[base + lu + 1]; k++, r++) {
                        const int v = P.adj_nbr[k] & 0x7fffffff;
                        const bool is_halo = (v < row0 || v >= row1);
                        const int idx = is_halo ? int(std::lower_bound(halo.begin(), halo.end(), v) - halo.begin()) : int(v - row0);
                        slots.push_back({lu, r, idx | (is_halo ? (1 << 20) : 0), adj_eid[k], P.adj_nbr[k] >= 0});
                    }
                    wt_rounds = std::max(wt_rounds, r);
                }
                if (opt.conflict_free_rounds) {
                    // the greedy list schedule of the stage kernel's rounds (build_level_plan) on the rows of the super-tile's buffers:
                    // the 8 lanes of a quarter-warp should read rows with distinct (index mod 8) -- own and halo rows both start at a
                    // 128-byte boundary, so (index mod 8) decides the bank group in either region
                    size_t s0 = 0;
                    while (s0 < slots.size()) {
                        const int g = slots[s0].owner / 8;
                        size_t s1 = s0;
                        while (s1 < slots.size() && slots[s1].owner / 8 == g) s1++;
                        std::vector<size_t> lane_edges[8];
                        for (size_t k = s0; k < s1; k++) lane_edges[slots[k].owner % 8].push_back(k);
                        for (int r = 0; r < wt_rounds; r++) {
                            int used_by[8];
                            for (int q = 0; q < 8; q++) used_by[q] = -1;
                            int order[8] = {0, 1, 2, 3, 4, 5, 6, 7};
                            std::stable_sort(order, order + 8, [&](int x, int y) { return lane_edges[x].size() > lane_edges[y].size(); });
                            const int rounds_left = wt_rounds - r;
                            for (int oi = 0; oi < 8; oi++) {
                                auto& le = lane_edges[order[oi]];
                                if (le.empty()) continue;
                                int pick = -1;
                                for (size_t c = 0; c < le.size(); c++) {
                                    const int row = slots[le[c]].other;
                                    if (used_by[row & 7] == -1 || used_by[row & 7] == row) { pick = int(c); break; }
                                }
                                if (pick < 0) {
                                    if (int(le.size()) < rounds_left) continue;
                                    pick = 0;
                                }
                                const size_t k = le[pick];
                                slots[k].round = r;
                                used_by[slots[k].other & 7] = slots[k].other;
                                le.erase(le.begin() + pick);
                            }
                        }
                        s0 = s1;
                    }
                }
                int rounds = 0;
                for (const VSlot& sl : slots) rounds = std::max(rounds, sl.round + 1);
                rounds = (rounds + RC - 1) / RC * RC;        // whole chunks: the kernel's chunk loop has a compile-time trip count
                const long b0 = blocks_done;
                blocks_done += rounds;
                O.blocks.resize(size_t(blocks_done) * BLK, 0);
                const int orow0 = int(base - row0);          // the warp-tile's first row inside the super-tile
                O.ents.push_back({orow0, rounds, w4 * WT, t, b0});
                for (int r = 0; r < rounds; r++) {
                    uint16_t* oth = reinterpret_cast<uint16_t*>(O.blocks.data() + size_t(b0 + r) * BLK + size_t(WT) * 32);
                    for (int lu = 0; lu < WT; lu++) oth[lu] = visit_code(orow0 + lu, false);      // empty slot: the node itself, h = 0, wk = 0
                }
                for (const VSlot& sl : slots) {
                    unsigned char* blk = O.blocks.data() + size_t(b0 + sl.round) * BLK;
                    double* w = reinterpret_cast<double*>(blk);
                    uint16_t* oth = reinterpret_cast<uint16_t*>(blk + size_t(WT) * 32);
                    const double sg = sl.owner_is_a ? -0.5 : 0.5;
                    const double hx = sg * P.ew[sl.e], hy = sg * P.ew[L.nI + sl.e], hz = sg * P.ew[2 * L.nI + sl.e];
                    w[sl.owner] = hx; w[WT + sl.owner] = hy; w[2 * WT + sl.owner] = hz;
                    w[3 * WT + sl.owner] = std::sqrt(std::fma(hx, hx, std::fma(hy, hy, hz * hz))) * k2;
                    oth[sl.owner] = visit_code(sl.other & 0xFFFFF, (sl.other >> 20) != 0);
                }
                // per-row sums of h in round order (rows of different super-tiles are disjoint: no two workers write the same entry)
                for (int r = 0; r < rounds; r++) {
                    const double* w = reinterpret_cast<const double*>(O.blocks.data() + size_t(b0 + r) * BLK);
                    for (int lu = 0; lu < WT; lu++) {
                        const size_t row = size_t(base + lu);
                        V.hsum[row] += w[lu]; V.hsum[P.npad_owned + row] += w[WT + lu]; V.hsum[2 * size_t(P.npad_owned) + row] += w[2 * WT + lu];
                    }
                }
            }
    };
    {
        std::atomic<long> next(0);
        auto worker = [&]() { for (long s = next.fetch_add(1); s < ns; s = next.fetch_add(1)) process_super(s); };
        const unsigned nthreads = unsigned(std::max<long>(1, std::min<long>(plan_threads(), ns / 4)));
        std::vector<std::future<void>> pool;
        for (unsigned k = 1; k < nthreads; k++) pool.push_back(std::async(std::launch::async, worker));
        worker();
        for (auto& f : pool) f.get();
    }
    for (const SuperOut& O : out) if (!O.error.empty()) throw std::runtime_error(O.error);
    V.maxt = 0; V.max_halo = 0; V.max_rounds = 0; V.halo_total = 0; V.max_ent = 0; V.vblocks = 0;
    V.went_off.assign(ns + 1, 0);
    std::vector<long> blk_off(ns + 1, 0);
    for (long s = 0; s < ns; s++) {
        V.maxt = std::max<int>(V.maxt, int(V.super_off[s + 1] - V.super_off[s]));
        V.max_halo = std::max<int>(V.max_halo, int(out[s].halo.size()));
        V.halo_total += long(out[s].halo.size());
        V.max_ent = std::max<int>(V.max_ent, int(out[s].ents.size()));
        V.went_off[s + 1] = V.went_off[s] + long(out[s].ents.size());
        for (const Ent& e : out[s].ents) V.max_rounds = std::max(V.max_rounds, e.rounds);
        blk_off[s + 1] = blk_off[s] + long(out[s].blocks.size() / BLK);
    }
    V.vblocks = blk_off[ns];
    if (V.vblocks >= (1L << 32)) throw std::runtime_error("mgcfd: too many edge round blocks for the visit kernel's 32-bit chunk index");
    V.vslots.resize(size_t(V.vblocks) * BLK);
    V.hpad = (V.max_halo + 3) & ~3;
    // chunk lists: warp w (of 16) takes the entries w, w+16, ... of a super-tile, each cut into chunks of R rounds
    const int NW = (opt.visit_warps == 8) ? 8 : 16, R = std::max(1, opt.visit_rounds);
    V.rounds_per_chunk = R; V.warps = NW;
    V.max_chunk = 0;
    for (long s = 0; s < ns; s++) {
        long n = 0;
        for (const Ent& e : out[s].ents) n += (e.rounds + R - 1) / R;
        V.max_chunk = std::max<int>(V.max_chunk, int(n));
    }
    if (V.max_chunk >= 65536) throw std::runtime_error("mgcfd: too many chunks in one super-tile");
    V.max_chunk = (V.max_chunk + 3) & ~3;
    V.desc_stride = (32 + 32 * V.max_ent + 48 + 4 * V.max_chunk + 4 * V.hpad + 15) & ~15;
    V.desc.assign(size_t(ns) * V.desc_stride, 0);
    for (long s = 0; s < ns; s++) {
        const long t0 = V.super_off[s], t1 = V.super_off[s + 1];
        if (!out[s].blocks.empty()) memcpy(V.vslots.data() + size_t(blk_off[s]) * BLK, out[s].blocks.data(), out[s].blocks.size());
        unsigned char* d = V.desc.data() + size_t(s) * V.desc_stride;
        int* di = reinterpret_cast<int*>(d);
        di[0] = int(t0 * TN); di[1] = int(t1 - t0); di[2] = int(out[s].halo.size()); di[3] = int(t0); di[4] = int(out[s].ents.size());
        for (size_t k = 0; k < out[s].ents.size(); k++) {
            const Ent& e = out[s].ents[k];
            unsigned char* eh = d + 32 + 32 * k;
            int* ei = reinterpret_cast<int*>(eh);
            ei[0] = e.orow0; ei[1] = e.rounds; ei[2] = int(P.bslot_off[e.tile + 1] - P.bslot_off[e.tile]); ei[3] = e.blane0;
            reinterpret_cast<long long*>(eh)[2] = blk_off[s] + e.blk0;
            reinterpret_cast<long long*>(eh)[3] = P.bslot_off[e.tile];
        }
        unsigned short* coff = reinterpret_cast<unsigned short*>(d + 32 + 32 * V.max_ent);
        unsigned* cl = reinterpret_cast<unsigned*>(d + 32 + 32 * V.max_ent + 48);
        int nch = 0;
        for (int w = 0; w < NW; w++) {
            coff[w] = (unsigned short)nch;
            for (size_t k = w; k < out[s].ents.size(); k += NW) {
                const Ent& e = out[s].ents[k];
                for (int r0 = 0; r0 < e.rounds; r0 += R, nch++) cl[nch] = unsigned(blk_off[s] + e.blk0 + r0);
            }
        }
        coff[NW] = (unsigned short)nch;
        di[5] = nch;
        int* ids = reinterpret_cast<int*>(d + 32 + 32 * V.max_ent + 48 + 4 * V.max_chunk);
        for (size_t k = 0; k < out[s].halo.size(); k++) ids[k] = out[s].halo[k];
        std::vector<unsigned char>().swap(out[s].blocks);
    }
}

long check_colouring(const LevelPlan& P) {
    long conflicts = 0;
    const long TN = P.TN;
    const size_t BLK = size_t(TN) * 26;
    std::vector<int> seen(TN);
    long stored = 0;
    for (long t = 0; t < P.ntiles; t++) {
        const long nhalo = P.halo_off[t + 1] - P.halo_off[t];
        for (long b = P.slot_off[t]; b < P.slot_off[t + 1]; b++) {
            const unsigned char* blk = P.slots.data() + size_t(b) * BLK;
            const double* w = reinterpret_cast<const double*>(blk);
            const uint16_t* oth = reinterpret_cast<const uint16_t*>(blk + size_t(TN) * 24);
            std::fill(seen.begin(), seen.end(), 0);
            for (long lu = 0; lu < TN; lu++) {
                const uint16_t code = oth[lu];
                const int o = row_of_code(code);
                const bool empty = P.scatter ? (code == 0xFFFF) : (o == lu && w[lu] == 0.0 && w[TN + lu] == 0.0 && w[2 * TN + lu] == 0.0);
                if (empty) continue;
                if (code != row_code(o)) conflicts++;
                stored++;
                if (lu >= P.tile_nown[t]) conflicts++;               // padding threads must own nothing
                if (o < TN) {
                    if (o >= P.tile_nown[t]) conflicts++;
                    if (P.scatter && seen[o]++) conflicts++;         // two writers of one node in one round
                } else if (o - TN >= nhalo) conflicts++;
            }
        }
    }
    // coverage: scatter mode stores an edge once if both ends share a tile and twice otherwise; segment mode always twice
    const long expect = P.scatter ? P.nI + P.cut_edges / 2 : 2 * P.nI;
    if (stored != expect) conflicts += 1000000;
    return conflicts;
}

void build_transfer_plan(const HostLevel& fine, const HostLevel& coarse, const LevelPlan& Pf, const LevelPlan& Pc, TransferPlan& T) {
    PhaseClock clk;
    T = TransferPlan();
    const long nf = fine.nel;
    const long nf_owned = fine.n_owned >= 0 ? fine.n_owned : fine.nel;
    const long nc_owned = coarse.n_owned >= 0 ? coarse.n_owned : coarse.nel;
    // fine nodes in ascending GLOBAL index: the reference's accumulation order in mg_restrict (mg_loops.cpp:101-154)
    std::vector<long> by_gid(nf);
    std::iota(by_gid.begin(), by_gid.end(), 0L);
    if (!fine.gid.empty()) std::sort(by_gid.begin(), by_gid.end(), [&](long x, long y) { return fine.gid[x] < fine.gid[y]; });
    // restrict: stable counting sort of the children of every OWNED coarse node
    auto restricts = [&](long i) { return fine.mg[i] >= 0 && fine.mg[i] < nc_owned; };
    T.child_off.assign(Pc.npad + 1, 0);
    long nchildren = 0;
    for (long i = 0; i < nf; i++) if (restricts(i)) { T.child_off[Pc.new_of_old[fine.mg[i]] + 1]++; nchildren++; }
    for (long i = 0; i < Pc.npad; i++) T.child_off[i + 1] += T.child_off[i];
    T.child_ids.resize(nchildren);
    {
        std::vector<long> pos(T.child_off.begin(), T.child_off.end() - 1);
        for (long i : by_gid) if (restricts(i)) T.child_ids[pos[Pc.new_of_old[fine.mg[i]]]++] = int(Pf.new_of_old[i]);
    }
    clk.lap("restrict operator");
    // prolong (owned fine nodes only)
    T.parent.assign(Pf.npad, -1);
    T.idist_own.assign(Pf.npad, 0.0);
    T.ent_off.assign(Pf.npad + 1, 0);
    for (long i = 0; i < Pf.npad; i++) {
        const long on = Pf.old_of_new[i];
        const long cnt = (on >= 0 && on < nf_owned) ? (Pf.adj_off[i + 1] - Pf.adj_off[i]) : 0;
        T.ent_off[i + 1] = T.ent_off[i] + cnt;
    }
    T.ent_src.resize(T.ent_off[Pf.npad]);
    T.ent_w.resize(T.ent_off[Pf.npad]);
    const double* cf = fine.coords.data();
    const double* cc = coarse.coords.data();
    auto idist = [](const double* x, const double* y) {
        const double dx = x[0] - y[0], dy = x[1] - y[1], dz = x[2] - y[2];
        return 1.0 / std::sqrt(dx * dx + dy * dy + dz * dz);
    };
    for (long id = 0; id < Pf.npad; id++) {
        const long on = Pf.old_of_new[id];
        if (on < 0 || on >= nf_owned) continue;
        const long p = fine.mg[on];
        if (p < 0) throw std::runtime_error("mgcfd: an owned fine node has no local coarse parent (partition closure violated)");
        T.parent[id] = int(Pc.new_of_old[p]);
        const double dx = cf[3 * on] - cc[3 * p], dy = cf[3 * on + 1] - cc[3 * p + 1], dz = cf[3 * on + 2] - cc[3 * p + 2];
        const bool coincident = (dx == 0.0 && dy == 0.0 && dz == 0.0);     // exact test, mg_loops.cpp:745,781
        T.idist_own[id] = coincident ? -1.0 : 1.0 / std::sqrt(dx * dx + dy * dy + dz * dz);
        for (long k = Pf.adj_off[id], o = T.ent_off[id]; k < Pf.adj_off[id + 1]; k++, o++) {
            const bool node_is_b = Pf.adj_nbr[k] < 0;
            const long om = Pf.old_of_new[Pf.adj_nbr[k] & 0x7fffffff];
            const long q = fine.mg[om];
            if (q < 0) throw std::runtime_error("mgcfd: the coarse parent of an edge neighbour is not local (partition closure violated)");
            T.ent_w[o] = idist(&cc[3 * q], &cf[3 * on]);
            // mg_loops.cpp:804-810: on the b side the neighbour-parent term multiplies residuals1[b1] (= own parent)
            T.ent_src[o] = int(Pc.new_of_old[node_is_b ? p : q]);
        }
    }
    clk.lap("prolong operator");
}


// ---- host walk of the device data structures ----------------------------------------------------------------------------------
// What a thread of the stage / transfer kernels does (kernels.cuh), restated on the host over the SAME byte streams the device
// receives (tile headers, round blocks, boundary blocks, transfer operators).  It exists so that the integer preprocessing can be
// checked against the oracle without a GPU (tests/test_host_mesh.py); it is not a compute path of the product (no entry point
// of the solver calls it) and uses plain libm arithmetic, so it agrees with the device to rounding, not bit for bit.
namespace {
struct HRec { double rho, mx, my, mz, re, ir, p, s; };
HRec host_rec(const double* v) {
    HRec n;
    n.rho = v[0]; n.mx = v[1]; n.my = v[2]; n.mz = v[3]; n.re = v[4];
    n.ir = 1.0 / n.rho;
    const double vx = n.mx * n.ir, vy = n.my * n.ir, vz = n.mz * n.ir;
    const double sq = vx * vx + vy * vy + vz * vz;
    n.p = (1.4 - 1.0) * (n.re - 0.5 * n.rho * sq);
    n.s = std::sqrt(sq) + std::sqrt(1.4 * n.p * n.ir);
    return n;
}
// edge_flux_acc_w (kernels.cuh): A's five increments for the edge A -> B, h = -0.5 * edge vector oriented A -> B
void host_edge_flux(const HRec& A, const HRec& B, double hx, double hy, double hz, double k2, double g[5]) {
    const double ewt = std::sqrt(hx * hx + hy * hy + hz * hz);
    const double factor = ewt * k2 * (A.s + B.s);
    const double gA = hx * A.mx + hy * A.my + hz * A.mz, gB = hx * B.mx + hy * B.my + hz * B.mz;
    const double qA = gA * A.ir, qB = gB * B.ir, ps = A.p + B.p;
    g[0] = factor * (A.rho - B.rho) + (gA + gB);
    g[4] = factor * (A.re - B.re) + ((A.re + A.p) * qA + (B.re + B.p) * qB);
    g[1] = factor * (A.mx - B.mx) + (A.mx * qA + B.mx * qB + ps * hx);
    g[2] = factor * (A.my - B.my) + (A.my * qA + B.my * qB + ps * hy);
    g[3] = factor * (A.mz - B.mz) + (A.mz * qA + B.mz * qB + ps * hz);
}
}  // namespace

void emulate_stage_flux(const LevelPlan& P, const double* var, int mask, const double ff[5], const double ffc[12], double k2, double* flux) {
    const long TN = P.TN;
    const size_t BLK = size_t(TN) * 26, BBLK = size_t(TN) * 25;
    std::vector<HRec> rec(P.npad);
    const double pad_state[5] = {ff[0], ff[1], ff[2], ff[3], ff[4]};
    for (long g = 0; g < P.npad; g++) rec[g] = host_rec(P.old_of_new[g] >= 0 ? var + 5 * P.old_of_new[g] : pad_state);
    std::vector<double> acc(5 * TN), f(5 * TN);
    for (long t = 0; t < P.ntiles; t++) {
        const unsigned char* h = P.hdrs.data() + size_t(t) * P.hdr_stride;
        const int* hi = reinterpret_cast<const int*>(h);
        const long long* hl = reinterpret_cast<const long long*>(h + 16);
        const int rounds = hi[0], nh = hi[1], brounds = hi[2];
        const int* ids = reinterpret_cast<const int*>(h + 32);
        auto row = [&](int o) -> const HRec& {
            if (o < TN) return rec[t * TN + o];
            if (o - TN >= nh) throw std::runtime_error("mgcfd: slot points past the tile's halo");
            return rec[ids[o - TN]];
        };
        std::fill(acc.begin(), acc.end(), 0.0);
        std::fill(f.begin(), f.end(), 0.0);
        if (mask & 1)
            for (int r = 0; r < rounds; r++) {
                const unsigned char* blk = P.slots.data() + size_t(hl[0] + r) * BLK;
                const double* w = reinterpret_cast<const double*>(blk);
                const uint16_t* oth = reinterpret_cast<const uint16_t*>(blk + size_t(TN) * 24);
                for (long lu = 0; lu < TN; lu++) {
                    const uint16_t code = oth[lu];
                    if (P.scatter && code == 0xFFFF) continue;
                    const int o = row_of_code(code);
                    if (code != row_code(o)) throw std::runtime_error("mgcfd: slot row code does not follow the swizzle");
                    double g[5];
                    host_edge_flux(rec[t * TN + lu], row(o), w[lu], w[TN + lu], w[2 * TN + lu], k2, g);
                    for (int k = 0; k < 5; k++) f[k * TN + lu] += g[k];
                    if (P.scatter && o < TN) for (int k = 0; k < 5; k++) acc[k * TN + o] -= g[k];
                }
            }
        if (mask & 6)
            for (int r = 0; r < brounds; r++) {
                const unsigned char* blk = P.bslots.data() + size_t(hl[1] + r) * BBLK;
                const double* w = reinterpret_cast<const double*>(blk);
                for (long lu = 0; lu < TN; lu++) {
                    const int kind = blk[size_t(TN) * 24 + lu];
                    if (kind == 0 || !((mask >> kind) & 1)) continue;
                    const HRec& B = rec[t * TN + lu];
                    const double x = w[lu], y = w[TN + lu], z = w[2 * TN + lu];
                    if (kind == 1) { f[1 * TN + lu] += x * B.p; f[2 * TN + lu] += y * B.p; f[3 * TN + lu] += z * B.p; }
                    else {            // flux_wall_kernel.elemfunc.c:47-69
                        const double fx = 0.5 * x, fy = 0.5 * y, fz = 0.5 * z;
                        const double g = fx * B.mx + fy * B.my + fz * B.mz, q = g * B.ir;
                        f[0 * TN + lu] += (fx * ff[1] + fy * ff[2] + fz * ff[3]) + g;
                        f[4 * TN + lu] += (fx * ffc[9] + fy * ffc[10] + fz * ffc[11]) + (B.re + B.p) * q;
                        f[1 * TN + lu] += (fx * ffc[0] + fy * ffc[1] + fz * ffc[2]) + (B.mx * q + B.p * fx);
                        f[2 * TN + lu] += (fx * ffc[3] + fy * ffc[4] + fz * ffc[5]) + (B.my * q + B.p * fy);
                        f[3 * TN + lu] += (fx * ffc[6] + fy * ffc[7] + fz * ffc[8]) + (B.mz * q + B.p * fz);
                    }
                }
            }
        for (long lu = 0; lu < TN; lu++) {
            const long on = P.old_of_new[t * TN + lu];
            if (on < 0) {            // padding threads hold empty slots only: exact zeros
                for (int k = 0; k < 5; k++) if (f[k * TN + lu] != 0.0) throw std::runtime_error("mgcfd: a padding thread accumulated flux");
                continue;
            }
            for (int k = 0; k < 5; k++) flux[5 * on + k] = f[k * TN + lu] + (P.scatter ? acc[k * TN + lu] : 0.0);
        }
    }
}

// the visit kernel's view of the same level: super-tile descriptors, their halo lists, warp-tile entries and re-addressed edge rounds
void emulate_visit_flux(const LevelPlan& P, const double* var, int mask, const double ff[5], const double ffc[12], double k2, double* flux) {
    const VisitPlan& V = P.visit;
    if (V.ns <= 0) throw std::runtime_error("mgcfd: the level has no visit plan");
    const long TN = P.TN;
    const int WT = 32;
    const size_t BLK = size_t(WT) * 34, BBLK = size_t(TN) * 25;
    if (V.hsum.size() != 3 * size_t(P.npad_owned)) throw std::runtime_error("mgcfd: the visit plan has no per-row edge-vector sums");
    std::vector<HRec> rec(P.npad);
    const double pad_state[5] = {ff[0], ff[1], ff[2], ff[3], ff[4]};
    for (long g = 0; g < P.npad; g++) rec[g] = host_rec(P.old_of_new[g] >= 0 ? var + 5 * P.old_of_new[g] : pad_state);
    std::vector<char> seen(P.npad, 0);
    long tiles_seen = 0;
    for (long s = 0; s < V.ns; s++) {
        const unsigned char* d = V.desc.data() + size_t(s) * V.desc_stride;
        const int* di = reinterpret_cast<const int*>(d);
        const long row0 = di[0]; const int ntile = di[1], nhalo = di[2]; const long tile0 = di[3]; const int nent = di[4];
        if (row0 != tile0 * TN || ntile < 1 || ntile > V.maxt || nhalo > V.hpad || nent > V.max_ent || nent > 4 * ntile) throw std::runtime_error("mgcfd: bad super-tile descriptor");
        const int* ids = reinterpret_cast<const int*>(d + 32 + 32 * V.max_ent + 48 + 4 * V.max_chunk);
        {   // the chunk lists: what warp w's lane 0 feeds the warp's ring with must be, in order, what the warp then consumes
            const unsigned short* coff = reinterpret_cast<const unsigned short*>(d + 32 + 32 * V.max_ent);
            const unsigned* cl = reinterpret_cast<const unsigned*>(d + 32 + 32 * V.max_ent + 48);
            const int R = V.rounds_per_chunk;
            const int NW = V.warps;
            if (coff[0] != 0 || coff[NW] != di[5] || di[5] > V.max_chunk) throw std::runtime_error("mgcfd: bad chunk offsets in a super-tile");
            for (int w = 0; w < NW; w++) {
                int pos = coff[w];
                for (int k = w; k < nent; k += NW) {
                    const unsigned char* eh = d + 32 + 32 * k;
                    const int rounds = reinterpret_cast<const int*>(eh)[1];
                    const long long vblk0 = reinterpret_cast<const long long*>(eh)[2];
                    if (rounds % R) throw std::runtime_error("mgcfd: a warp-tile's rounds are not whole chunks");
                    for (int r0 = 0; r0 < rounds; r0 += R, pos++)
                        if (pos >= coff[w + 1] || cl[pos] != unsigned(vblk0 + r0)) throw std::runtime_error("mgcfd: a warp's chunk list does not match its warp-tiles");
                }
                if (pos != coff[w + 1]) throw std::runtime_error("mgcfd: a warp's chunk list is longer than its warp-tiles");
            }
        }
        for (int k = 0; k < nhalo; k++) {
            if (ids[k] < 0 || ids[k] >= P.npad || (ids[k] >= row0 && ids[k] < row0 + ntile * TN)) throw std::runtime_error("mgcfd: bad halo id in a super-tile");
            if (k && ids[k] <= ids[k - 1]) throw std::runtime_error("mgcfd: halo ids of a super-tile are not strictly ascending");
        }
        tiles_seen += ntile;
        for (int k = 0; k < nent; k++) {
            const unsigned char* eh = d + 32 + 32 * k;
            const int* ei = reinterpret_cast<const int*>(eh);
            const int orow0 = ei[0], rounds = ei[1], brounds = ei[2], blane0 = ei[3];
            const long long vblk0 = reinterpret_cast<const long long*>(eh)[2], bblk0 = reinterpret_cast<const long long*>(eh)[3];
            if (orow0 < 0 || orow0 + WT > ntile * TN || (orow0 % WT) || blane0 != orow0 % TN || vblk0 < 0 || vblk0 + rounds > V.vblocks) throw std::runtime_error("mgcfd: bad warp-tile entry");
            for (int lu = 0; lu < WT; lu++) {
                const long gid = row0 + orow0 + lu;
                if (seen[gid]++) throw std::runtime_error("mgcfd: a row belongs to two warp-tiles");
                const HRec& me = rec[gid];
                double f[5] = {0, 0, 0, 0, 0};
                double hs[3] = {0, 0, 0};
                if (mask & 1)
                    for (int r = 0; r < rounds; r++) {
                        const unsigned char* blk = V.vslots.data() + size_t(vblk0 + r) * BLK;
                        const double* w = reinterpret_cast<const double*>(blk);
                        const uint16_t code = reinterpret_cast<const uint16_t*>(blk + size_t(WT) * 32)[lu];
                        const double hx = w[lu], hy = w[WT + lu], hz = w[2 * WT + lu], wk = w[3 * WT + lu];
                        // the stream's precomputed |h| * k2 against its definition (one rounding each: exact equality)
                        if (wk != std::sqrt(std::fma(hx, hx, std::fma(hy, hy, hz * hz))) * k2) throw std::runtime_error("mgcfd: visit slot weight is not |h| * k2");
                        hs[0] += hx; hs[1] += hy; hs[2] += hz;
                        const bool is_halo = (code & 0x8000) != 0;
                        const int idx = (code & 0x7fff) >> 2;
                        if ((code & 3) != ((idx >> 1) & 3)) throw std::runtime_error("mgcfd: visit slot code does not follow the swizzle");
                        if (is_halo ? idx >= nhalo : idx >= ntile * TN) throw std::runtime_error("mgcfd: visit slot points outside the super-tile");
                        const HRec& B = is_halo ? rec[ids[idx]] : rec[row0 + idx];
                        double g[5];
                        host_edge_flux(me, B, w[lu], w[WT + lu], w[2 * WT + lu], k2, g);
                        for (int v = 0; v < 5; v++) f[v] += g[v];
                    }
                if (mask & 6)
                    for (int r = 0; r < brounds; r++) {
                        const unsigned char* blk = P.bslots.data() + size_t(bblk0 + r) * BBLK;
                        const double* w = reinterpret_cast<const double*>(blk);
                        const int bl = blane0 + lu;
                        const int kind = blk[size_t(TN) * 24 + bl];
                        if (kind == 0 || !((mask >> kind) & 1)) continue;
                        const double x = w[bl], y = w[TN + bl], z = w[2 * TN + bl];
                        if (kind == 1) { f[1] += x * me.p; f[2] += y * me.p; f[3] += z * me.p; }
                        else {
                            const double fx = 0.5 * x, fy = 0.5 * y, fz = 0.5 * z;
                            const double g = fx * me.mx + fy * me.my + fz * me.mz, q = g * me.ir;
                            f[0] += (fx * ff[1] + fy * ff[2] + fz * ff[3]) + g;
                            f[4] += (fx * ffc[9] + fy * ffc[10] + fz * ffc[11]) + (me.re + me.p) * q;
                            f[1] += (fx * ffc[0] + fy * ffc[1] + fz * ffc[2]) + (me.mx * q + me.p * fx);
                            f[2] += (fx * ffc[3] + fy * ffc[4] + fz * ffc[5]) + (me.my * q + me.p * fy);
                            f[3] += (fx * ffc[6] + fy * ffc[7] + fz * ffc[8]) + (me.mz * q + me.p * fz);
                        }
                    }
                if (mask & 1)       // the hoisted A-side terms use these sums: same additions in the same (round) order
                    for (int v = 0; v < 3; v++) if (hs[v] != V.hsum[size_t(v) * P.npad_owned + gid]) throw std::runtime_error("mgcfd: per-row edge-vector sum does not match the slots");
                const long on = P.old_of_new[gid];
                if (on < 0) {
                    for (int v = 0; v < 5; v++) if (f[v] != 0.0) throw std::runtime_error("mgcfd: a padding thread accumulated flux");
                    continue;
                }
                for (int v = 0; v < 5; v++) flux[5 * on + v] = f[v];
            }
        }
    }
    if (tiles_seen != P.ntiles) throw std::runtime_error("mgcfd: the super-tiles do not cover the tiles");
    // every node must have been visited (rows never listed are padding)
    for (long g = 0; g < P.npad_owned; g++) if (!seen[g] && P.old_of_new[g] >= 0) throw std::runtime_error("mgcfd: a node belongs to no warp-tile");
}

void emulate_restrict(const LevelPlan& Pf, const LevelPlan& Pc, const TransferPlan& T, const double* var_f, double* var_c) {
    (void)Pf;
    for (long c = 0; c < Pc.npad; c++) {
        const long k0 = T.child_off[c], k1 = T.child_off[c + 1], oc = Pc.old_of_new[c];
        if (k1 == k0 || oc < 0) continue;
        double sum[5] = {0, 0, 0, 0, 0};
        for (long k = k0; k < k1; k++) {
            const long of = Pf.old_of_new[T.child_ids[k]];
            for (int j = 0; j < 5; j++) sum[j] += var_f[5 * of + j];
        }
        const double average = 1.0 / double(k1 - k0);
        for (int j = 0; j < 5; j++) var_c[5 * oc + j] = sum[j] * average;
    }
}

void emulate_prolong(const LevelPlan& Pf, const LevelPlan& Pc, const TransferPlan& T, const double* res_c, const double* res_f, double* var_f) {
    for (long i = 0; i < Pf.npad; i++) {
        const int p = T.parent[i];
        const long on = Pf.old_of_new[i];
        if (p < 0 || on < 0) continue;
        const double* rp = res_c + 5 * Pc.old_of_new[p];
        const double w0 = T.idist_own[i];
        double acc[5] = {0, 0, 0, 0, 0}, wsum = 0.0;
        const long k0 = T.ent_off[i], k1 = T.ent_off[i + 1];
        if (w0 < 0.0) {
            if (k1 > k0) { for (int j = 0; j < 5; j++) acc[j] = rp[j]; wsum = 1.0; }
        } else {
            for (long k = k0; k < k1; k++) {
                const double* rq = res_c + 5 * Pc.old_of_new[T.ent_src[k]];
                const double w = T.ent_w[k];
                for (int j = 0; j < 5; j++) { acc[j] += w0 * rp[j]; acc[j] += w * rq[j]; }
                wsum += w0; wsum += w;
            }
        }
        for (int j = 0; j < 5; j++) var_f[5 * on + j] = var_f[5 * on + j] + (res_f[5 * on + j] - acc[j] / wsum);
    }
}

}  // namespace mgcfd
