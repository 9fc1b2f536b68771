// partition.cpp -- domain decomposition of a multigrid mesh over ranks (see partition.h).
#include "partition.h"

#include <chrono>
#include <cstdio>
#include <cstdlib>

#include <algorithm>
#include <cstdint>
#include <numeric>
#include <stdexcept>

namespace mgcfd {

namespace {

const int MESH_FVCORR = 0;

// MGCFD_PLAN_TIMING=1: wall time of the phases of the rank-local partitioning on stderr
struct PhaseClock {
    bool on = getenv("MGCFD_PLAN_TIMING") != nullptr;
    std::chrono::steady_clock::time_point t = std::chrono::steady_clock::now();
    void lap(const char* what) {
        if (!on) return;
        const auto now = std::chrono::steady_clock::now();
        fprintf(stderr, "  partition: %-26s %8.3f s\n", what, std::chrono::duration<double>(now - t).count());
        t = now;
    }
};

struct Csr {
    std::vector<long> off;
    std::vector<int> idx;
};

// undirected adjacency over the internal edges (global ids)
Csr adjacency(const HostLevel& L) {
    Csr g;
    g.off.assign(L.nel + 1, 0);
    for (long e = 0; e < L.nI; e++) { g.off[L.edges[e].a + 1]++; g.off[L.edges[e].b + 1]++; }
    for (long i = 0; i < L.nel; i++) g.off[i + 1] += g.off[i];
    g.idx.resize(g.off[L.nel]);
    std::vector<long> pos(g.off.begin(), g.off.end() - 1);
    for (long e = 0; e < L.nI; e++) {
        g.idx[pos[L.edges[e].a]++] = int(L.edges[e].b);
        g.idx[pos[L.edges[e].b]++] = int(L.edges[e].a);
    }
    return g;
}
// children (fine ids) of every coarse node
Csr children(const HostLevel& fine, long ncoarse) {
    Csr c;
    c.off.assign(ncoarse + 1, 0);
    for (long i = 0; i < fine.nel; i++) c.off[fine.mg[i] + 1]++;
    for (long k = 0; k < ncoarse; k++) c.off[k + 1] += c.off[k];
    c.idx.resize(fine.nel);
    std::vector<long> pos(c.off.begin(), c.off.end() - 1);
    for (long i = 0; i < fine.nel; i++) c.idx[pos[fine.mg[i]]++] = int(i);
    return c;
}

void rcb(const double* coords, long* idx, long n, int k, int rank0, std::vector<int>& owner) {
    if (k == 1) { for (long i = 0; i < n; i++) owner[idx[i]] = rank0; return; }
    const int kl = k / 2;
    const long nl = (n * kl + k / 2) / k;
    double lo[3] = {1e300, 1e300, 1e300}, hi[3] = {-1e300, -1e300, -1e300};
    for (long i = 0; i < n; i++) for (int d = 0; d < 3; d++) {
        const double c = coords[3 * idx[i] + d];
        lo[d] = std::min(lo[d], c); hi[d] = std::max(hi[d], c);
    }
    int ax = 0;
    for (int d = 1; d < 3; d++) if (hi[d] - lo[d] > hi[ax] - lo[ax]) ax = d;
    std::nth_element(idx, idx + nl, idx + n, [coords, ax](long x, long y) {
        const double cx = coords[3 * x + ax], cy = coords[3 * y + ax];
        return cx != cy ? cx < cy : x < y;
    });
    rcb(coords, idx, nl, kl, rank0, owner);
    rcb(coords, idx + nl, n - nl, k - kl, rank0 + kl, owner);
}

}  // namespace

void rcb_owners(const HostLevel& L, int nranks, std::vector<int>& owner) {
    owner.assign(L.nel, 0);
    if (nranks <= 1) return;
    std::vector<long> idx(L.nel);
    std::iota(idx.begin(), idx.end(), 0L);
    if (L.coords.empty()) {      // no coordinates (single-level meshes may come without): contiguous blocks of node ids
        for (long i = 0; i < L.nel; i++) owner[i] = int((i * nranks) / L.nel);
        return;
    }
    rcb(L.coords.data(), idx.data(), L.nel, nranks, 0, owner);
}

void partition_mesh(const HostMesh& full, int nranks, int rank, LocalMesh& out) {
    if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) throw std::runtime_error("mgcfd: bad rank / nranks (1..64 ranks)");
    const int nl = int(full.levels.size());
    out = LocalMesh();
    out.mesh_variant = full.mesh_variant; out.rank = rank; out.nranks = nranks;
    out.levels.resize(nl);

    // the fine -> coarse maps index per-node arrays of the next level below: validate them before anything does
    for (int l = 0; l + 1 < nl; l++) {
        const HostLevel& F = full.levels[l];
        if (long(F.mg.size()) != F.nel) throw std::runtime_error("mgcfd: level " + std::to_string(l) + " has no fine -> coarse map of one entry per node");
        for (long v : F.mg) if (v < 0 || v >= full.levels[l + 1].nel) throw std::runtime_error("mgcfd: mg_map entry of level " + std::to_string(l) + " out of range");
    }
    std::vector<std::vector<int>> owner(nl);
    std::vector<Csr> adj(nl), kids(nl);         // kids[l]: children in level l-1 of the nodes of level l
    // a mesh duplicated m times (-m, mgcfd_mesh_duplicate: copy-major node numbering on every level) with m a multiple of the
    // number of ranks is dealt out copy by copy: independent replicas, no halo at all -- the reference's assess-memory protocol
    // (one mesh copy per thread) across GPUs.  Any other mesh is cut by recursive coordinate bisection.
    bool whole_copies = full.copies > 1 && full.copies % nranks == 0;
    for (int l = 0; l < nl && whole_copies; l++) whole_copies = full.levels[l].nel % full.copies == 0;
    for (int l = 0; l < nl; l++) {
        if (whole_copies) {
            const long per_copy = full.levels[l].nel / full.copies, copies_per_rank = full.copies / nranks;
            owner[l].resize(full.levels[l].nel);
            for (long i = 0; i < full.levels[l].nel; i++) owner[l][i] = int((i / per_copy) / copies_per_rank);
        } else rcb_owners(full.levels[l], nranks, owner[l]);
        adj[l] = adjacency(full.levels[l]);
        if (l > 0) kids[l] = children(full.levels[l - 1], full.levels[l].nel);
    }
    // flux_local[l][i]: owned by `rank` or edge neighbour of a node owned by `rank`
    std::vector<std::vector<char>> flux_local(nl);
    for (int l = 0; l < nl; l++) {
        const long n = full.levels[l].nel;
        flux_local[l].assign(n, 0);
        for (long i = 0; i < n; i++) {
            if (owner[l][i] != rank) continue;
            flux_local[l][i] = 1;
            for (long k = adj[l].off[i]; k < adj[l].off[i + 1]; k++) flux_local[l][adj[l].idx[k]] = 1;
        }
    }
    // local node sets + global -> local maps
    std::vector<std::vector<int>> g2l(nl);
    for (int l = 0; l < nl; l++) {
        const HostLevel& G = full.levels[l];
        const long n = G.nel;
        std::vector<char> need(n, 0);
        for (long i = 0; i < n; i++) if (flux_local[l][i]) need[i] = 1;
        if (l + 1 < nl) for (long i = 0; i < n; i++) if (owner[l + 1][G.mg[i]] == rank) need[i] = 1;            // restrict
        if (l >= 1) {                                                                                               // prolong
            const HostLevel& F = full.levels[l - 1];
            for (long j = 0; j < F.nel; j++) if (flux_local[l - 1][j]) need[F.mg[j]] = 1;
        }
        LocalLevel& LL = out.levels[l];
        std::vector<long> ghosts;
        for (long i = 0; i < n; i++) {
            if (owner[l][i] == rank) LL.gid.push_back(i);
            else if (need[i]) ghosts.push_back(i);
        }
        LL.n_owned = long(LL.gid.size());
        std::stable_sort(ghosts.begin(), ghosts.end(), [&](long x, long y) { return owner[l][x] < owner[l][y]; });   // ids stay ascending per owner
        LL.recv_off.assign(nranks + 1, 0);
        for (long g : ghosts) LL.recv_off[owner[l][g] + 1]++;
        for (int p = 0; p < nranks; p++) LL.recv_off[p + 1] += LL.recv_off[p];
        LL.gid.insert(LL.gid.end(), ghosts.begin(), ghosts.end());
        g2l[l].assign(n, -1);
        for (size_t k = 0; k < LL.gid.size(); k++) g2l[l][LL.gid[k]] = int(k);
        // what the peers need from this rank (the mirror image of the three rules above)
        std::vector<uint64_t> wanted(LL.n_owned, 0);
        for (long k = 0; k < LL.n_owned; k++) {
            const long i = LL.gid[k];
            uint64_t m = 0;
            for (long q = adj[l].off[i]; q < adj[l].off[i + 1]; q++) m |= 1ull << owner[l][adj[l].idx[q]];
            if (l + 1 < nl) m |= 1ull << owner[l + 1][G.mg[i]];
            if (l >= 1) {
                for (long c = kids[l].off[i]; c < kids[l].off[i + 1]; c++) {
                    const long j = kids[l].idx[c];
                    m |= 1ull << owner[l - 1][j];
                    for (long q = adj[l - 1].off[j]; q < adj[l - 1].off[j + 1]; q++) m |= 1ull << owner[l - 1][adj[l - 1].idx[q]];
                }
            }
            wanted[k] = m & ~(1ull << rank);
        }
        LL.send_off.assign(nranks + 1, 0);
        for (int p = 0; p < nranks; p++) {
            for (long k = 0; k < LL.n_owned; k++) if ((wanted[k] >> p) & 1) LL.send_idx.push_back(k);
            LL.send_off[p + 1] = long(LL.send_idx.size());
        }
    }
    // local meshes
    for (int l = 0; l < nl; l++) {
        const HostLevel& G = full.levels[l];
        LocalLevel& LL = out.levels[l];
        HostLevel& M = LL.mesh;
        const long nloc = long(LL.gid.size());
        M.nel = nloc; M.name = G.name;
        M.volumes.resize(nloc);
        if (!G.coords.empty()) M.coords.resize(3 * nloc);
        for (long k = 0; k < nloc; k++) {
            const long i = LL.gid[k];
            M.volumes[k] = G.volumes[i];
            if (!G.coords.empty()) for (int d = 0; d < 3; d++) M.coords[3 * k + d] = G.coords[3 * i + d];
        }
        for (long e = 0; e < G.nI; e++) {
            const EdgeNb& ed = G.edges[e];
            if (owner[l][ed.a] != rank && owner[l][ed.b] != rank) continue;
            M.edges.push_back({long(g2l[l][ed.a]), long(g2l[l][ed.b]), ed.x, ed.y, ed.z});
            LL.edge_gid.push_back(e);
        }
        M.nI = long(M.edges.size());
        LL.nI_global = G.nI;
        for (int cls = 0; cls < 2; cls++) {
            const long e0 = cls == 0 ? G.nI : G.nI + G.nB, e1 = cls == 0 ? G.nI + G.nB : G.nI + G.nB + G.nW;
            long cnt = 0;
            for (long e = e0; e < e1; e++) {
                const EdgeNb& ed = G.edges[e];
                if (owner[l][ed.b] != rank) continue;
                M.edges.push_back({ed.a, long(g2l[l][ed.b]), ed.x, ed.y, ed.z});
                cnt++;
            }
            (cls == 0 ? M.nB : M.nW) = cnt;
        }
        if (l + 1 < nl) {
            M.mg.resize(nloc);
            for (long k = 0; k < nloc; k++) M.mg[k] = g2l[l + 1][G.mg[LL.gid[k]]];
        }
    }
}

void partition_sources(const std::vector<LevelSource>& levels, int mesh_variant, int nranks, int rank, LocalMesh& out) {
    if (nranks < 1 || nranks > 64 || rank < 0 || rank >= nranks) throw std::runtime_error("mgcfd: bad rank / nranks (1..64 ranks)");
    const int nl = int(levels.size());
    out = LocalMesh();
    out.mesh_variant = mesh_variant; out.rank = rank; out.nranks = nranks;
    out.levels.resize(nl);
    std::vector<long> n(nl);
    std::vector<std::vector<int>> owner(nl);
    std::vector<Csr> kids(nl);
    PhaseClock clk;
    for (int l = 0; l < nl; l++) {
        const NodeSource& S = *levels[l].src;
        n[l] = S.nel();
        HostLevel tmp;                       // coordinates only, for the bisection
        tmp.nel = n[l];
        tmp.coords.resize(3 * n[l]);
        for (long i = 0; i < n[l]; i++) S.coords(i, &tmp.coords[3 * i]);
        rcb_owners(tmp, nranks, owner[l]);
        if (l > 0) {
            const std::vector<int>& mg = levels[l - 1].mg;
            if (long(mg.size()) != n[l - 1]) throw std::runtime_error("mgcfd: level " + std::to_string(l - 1) + " has no fine -> coarse map of one entry per node");
            for (int v : mg) if (v < 0 || v >= n[l]) throw std::runtime_error("mgcfd: mg_map entry of level " + std::to_string(l - 1) + " out of range");
            Csr& c = kids[l];
            c.off.assign(n[l] + 1, 0);
            for (long i = 0; i < n[l - 1]; i++) c.off[mg[i] + 1]++;
            for (long k = 0; k < n[l]; k++) c.off[k + 1] += c.off[k];
            c.idx.resize(n[l - 1]);
            std::vector<long> pos(c.off.begin(), c.off.end() - 1);
            for (long i = 0; i < n[l - 1]; i++) c.idx[pos[mg[i]]++] = int(i);
        }
    }
    clk.lap("coordinates + bisection");
    std::vector<std::vector<char>> flux_local(nl);
    for (int l = 0; l < nl; l++) {
        const NodeSource& S = *levels[l].src;
        flux_local[l].assign(n[l], 0);
        for (long i = 0; i < n[l]; i++) {
            if (owner[l][i] != rank) continue;
            flux_local[l][i] = 1;
            int deg = 0;
            const Entry* ent = S.listing(i, deg);
            for (int j = 0; j < deg; j++) if (ent[j].nbr >= 0) flux_local[l][ent[j].nbr] = 1;
        }
    }
    clk.lap("flux closure");
    std::vector<std::vector<int>> g2l(nl);
    for (int l = 0; l < nl; l++) {
        const NodeSource& S = *levels[l].src;
        std::vector<char> need(n[l], 0);
        for (long i = 0; i < n[l]; i++) if (flux_local[l][i]) need[i] = 1;
        if (l + 1 < nl) for (long i = 0; i < n[l]; i++) if (owner[l + 1][levels[l].mg[i]] == rank) need[i] = 1;
        if (l >= 1) for (long j = 0; j < n[l - 1]; j++) if (flux_local[l - 1][j]) need[levels[l - 1].mg[j]] = 1;
        LocalLevel& LL = out.levels[l];
        std::vector<long> ghosts;
        for (long i = 0; i < n[l]; i++) {
            if (owner[l][i] == rank) LL.gid.push_back(i);
            else if (need[i]) ghosts.push_back(i);
        }
        LL.n_owned = long(LL.gid.size());
        std::stable_sort(ghosts.begin(), ghosts.end(), [&](long x, long y) { return owner[l][x] < owner[l][y]; });
        LL.recv_off.assign(nranks + 1, 0);
        for (long g : ghosts) LL.recv_off[owner[l][g] + 1]++;
        for (int p = 0; p < nranks; p++) LL.recv_off[p + 1] += LL.recv_off[p];
        LL.gid.insert(LL.gid.end(), ghosts.begin(), ghosts.end());
        g2l[l].assign(n[l], -1);
        for (size_t k = 0; k < LL.gid.size(); k++) g2l[l][LL.gid[k]] = int(k);
        std::vector<uint64_t> wanted(LL.n_owned, 0);
        for (long k = 0; k < LL.n_owned; k++) {
            const long i = LL.gid[k];
            uint64_t m = 0;
            int deg = 0;
            const Entry* ent = S.listing(i, deg);
            for (int j = 0; j < deg; j++) if (ent[j].nbr >= 0) m |= 1ull << owner[l][ent[j].nbr];
            if (l + 1 < nl) m |= 1ull << owner[l + 1][levels[l].mg[i]];
            if (l >= 1) {
                const NodeSource& F = *levels[l - 1].src;
                for (long c = kids[l].off[i]; c < kids[l].off[i + 1]; c++) {
                    const long j = kids[l].idx[c];
                    m |= 1ull << owner[l - 1][j];
                    int d2 = 0;
                    const Entry* e2 = F.listing(j, d2);
                    for (int q = 0; q < d2; q++) if (e2[q].nbr >= 0) m |= 1ull << owner[l - 1][e2[q].nbr];
                }
            }
            wanted[k] = m & ~(1ull << rank);
        }
        LL.send_off.assign(nranks + 1, 0);
        for (int p = 0; p < nranks; p++) {
            for (long k = 0; k < LL.n_owned; k++) if ((wanted[k] >> p) & 1) LL.send_idx.push_back(k);
            LL.send_off[p + 1] = long(LL.send_idx.size());
        }
    }
    clk.lap("local sets + send lists");
    for (int l = 0; l < nl; l++) {
        const NodeSource& S = *levels[l].src;
        LocalLevel& LL = out.levels[l];
        HostLevel& M = LL.mesh;
        const long nloc = long(LL.gid.size());
        M.nel = nloc; M.name = levels[l].name;
        M.volumes.resize(nloc);
        M.coords.resize(3 * nloc);
        for (long k = 0; k < nloc; k++) { M.volumes[k] = S.volume(LL.gid[k]); S.coords(LL.gid[k], &M.coords[3 * k]); }
        // edges exactly as read_grid creates them (io.cpp:84-181; build_level_like_read_grid): an edge per entry with nbr < i, in
        // ascending i then listing order = the global edge order; the running count of internal entries is the global edge index
        std::vector<EdgeNb> bnd, wall;
        long global_edge = 0;
        for (long i = 0; i < n[l]; i++) {
            int deg = 0;
            const Entry* ent = S.listing(i, deg);
            const bool mine = owner[l][i] == rank;
            for (int j = 0; j < deg; j++) {
                const long i2 = ent[j].nbr;
                if (i2 >= i) continue;
                if (i2 >= 0) {
                    if (mine || owner[l][i2] == rank) {
                        EdgeNb e = {long(g2l[l][i2]), long(g2l[l][i]), -ent[j].w[0], -ent[j].w[1], -ent[j].w[2]};
                        M.edges.push_back(e);
                        LL.edge_gid.push_back(global_edge);
                    }
                    global_edge++;
                } else if (mine) {
                    EdgeNb e = {i2, long(g2l[l][i]), ent[j].w[0], ent[j].w[1], ent[j].w[2]};
                    if (mesh_variant == MESH_FVCORR) { e.x *= -1; e.y *= -1; e.z *= -1; }
                    (i2 == -1 ? bnd : wall).push_back(e);
                }
            }
        }
        LL.nI_global = global_edge;
        M.nI = long(M.edges.size()); M.nB = long(bnd.size()); M.nW = long(wall.size());
        M.edges.insert(M.edges.end(), bnd.begin(), bnd.end());
        M.edges.insert(M.edges.end(), wall.begin(), wall.end());
        if (l + 1 < nl) {
            M.mg.resize(nloc);
            for (long k = 0; k < nloc; k++) M.mg[k] = g2l[l + 1][levels[l].mg[LL.gid[k]]];
        }
    }
    clk.lap("local meshes");
}

int build_tile_order(long owned_rows, int tile_nodes, const SendTargets& st, const std::vector<long>& halo_off, const std::vector<int>& halo_ids,
                     long ntiles, std::vector<int>& order) {
    const long nu = (owned_rows + tile_nodes - 1) / tile_nodes;
    std::vector<char> sends(nu, 0);
    for (long r = 0; r < owned_rows; r++) if (st.off[r + 1] > st.off[r]) sends[r / tile_nodes] = 1;
    // a tile that reads a ghost row without owning a delivered row would break the late wait: look at the halo lists as well
    for (long t = 0; t < ntiles && t < nu; t++)
        for (long k = halo_off[t]; k < halo_off[t + 1]; k++) if (halo_ids[k] >= owned_rows) sends[t] = 1;
    order.clear();
    for (long u = 0; u < nu; u++) if (!sends[u]) order.push_back((int)u);
    const int n_last = (int)(nu - (long)order.size());
    for (long u = 0; u < nu; u++) if (sends[u]) order.push_back((int)u);
    return n_last;
}

void build_send_targets(long owned_rows, int tile_nodes, const std::vector<int>& send_rows, const std::vector<PeerSlice>& peers, SendTargets& out) {
    out = SendTargets();
    out.off.assign(owned_rows + 1, 0);
    out.tile_sends.assign((owned_rows + tile_nodes - 1) / tile_nodes, 0);
    for (const PeerSlice& p : peers)
        for (long k = 0; k < p.nsend; k++) {
            const int node = send_rows[p.send0 + k];
            if (node < 0 || node >= owned_rows) throw std::runtime_error("mgcfd: send list entry is not an owned row");
            out.off[node + 1]++;
        }
    for (long i = 0; i < owned_rows; i++) out.off[i + 1] += out.off[i];
    out.peer.resize(out.off[owned_rows]); out.row.resize(out.off[owned_rows]);
    std::vector<int> pos(out.off.begin(), out.off.end() - 1);
    for (size_t pi = 0; pi < peers.size(); pi++)
        for (long k = 0; k < peers[pi].nsend; k++) {
            const int node = send_rows[peers[pi].send0 + k];
            out.peer[pos[node]] = int(pi);
            out.row[pos[node]] = int(peers[pi].first_ghost_row + peers[pi].recv_off_me + k);
            pos[node]++;
            out.tile_sends[node / tile_nodes] = 1;
        }
}

}  // namespace mgcfd
