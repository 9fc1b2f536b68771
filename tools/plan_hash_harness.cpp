// plan_hash_harness.cpp -- host-only regression guard for the integer preprocessing of DISTRIBUTED levels (owned nodes + ghosts):
// generates the three synthetic mesh kinds, partitions them over 2 / 3 / 8 ranks, builds every rank's level plans (both flux
// modes) and transfer operators and prints one FNV-1a hash per mesh kind over everything the device would receive.
// tests/test_partition.py compiles it against csrc/ and compares with tests/golden/plan_hashes.txt; when the plan format is
// It also plays the in-kernel halo exchange's send-target tables (partition.h: build_send_targets) through: every ghost row of every
// rank must receive exactly its own node, exactly once.  When the plan format is
// changed ON PURPOSE, rerun it and update that file:
//   g++ -O2 -std=c++17 -pthread -ffp-contract=off -Img-cfd-app-plain_b200/csrc -o /tmp/plan_hash tools/plan_hash_harness.cpp \
//       mg-cfd-app-plain_b200/csrc/{plan,partition,mesh_gen,mesh_io}.cpp && /tmp/plan_hash
#include <cstdio>
#include <string>
#include "host_mesh.h"
#include "partition.h"
#include "plan.h"
using namespace mgcfd;
static unsigned long long H = 1469598103934665603ull;
static void mix(const void* p, size_t n) { const unsigned char* b = (const unsigned char*)p; for (size_t i = 0; i < n; i++) { H ^= b[i]; H *= 1099511628211ull; } }
template <class T> static void mixv(const std::vector<T>& v) { mix(v.data(), v.size() * sizeof(T)); }
int main() {
    for (int kind = 0; kind < 3; kind++) {
        MeshSpec spec; spec.kind = kind; spec.levels = kind == 2 ? 1 : 3; spec.mesh_variant = kind == 2 ? 0 : 2;
        long d[3][3] = {{33, 29, 25}, {17, 15, 13}, {9, 8, 7}};
        for (int l = 0; l < 3; l++) for (int k = 0; k < 3; k++) spec.dims[l][k] = kind == 2 ? d[l][k] / 3 : d[l][k];
        HostMesh full; std::string err;
        if (generate_mesh(spec, full, err)) { printf("gen error %s\n", err.c_str()); return 1; }
        for (int nr : {2, 3, 8}) {
            std::vector<LocalMesh> locs(nr);
            std::vector<std::vector<LevelPlan>> all_plans(nr);
            for (int r = 0; r < nr; r++) {
                LocalMesh& loc = locs[r]; partition_mesh(full, nr, r, loc);
                std::vector<LevelPlan>& plans = all_plans[r];
                plans.resize(loc.levels.size());
                for (size_t l = 0; l < loc.levels.size(); l++) {
                    HostLevel HL = loc.levels[l].mesh; HL.n_owned = loc.levels[l].n_owned; HL.gid = loc.levels[l].gid;
                    for (int sc = 0; sc < 2; sc++) {
                        PlanOptions po; po.tile_nodes = 128; po.scatter = sc; po.ordering = 2;
                        LevelPlan P; build_level_plan(HL, po, P);
                        mixv(P.new_of_old); mixv(P.old_of_new); mixv(P.hdrs); mixv(P.slots); mixv(P.bslots); mixv(P.halo_ids); mixv(P.adj_off); mixv(P.adj_nbr); mixv(P.ea); mixv(P.eb); mixv(P.ew); mixv(P.bnode); mixv(P.bw);
                        long tail[6] = {P.cut_edges, P.used_slots, P.max_halo, P.max_rounds, P.npad, P.ntiles}; mix(tail, sizeof(tail));
                        if (!sc) plans[l] = P;
                    }
                }
                for (size_t l = 0; l + 1 < loc.levels.size(); l++) {
                    HostLevel F = loc.levels[l].mesh; F.n_owned = loc.levels[l].n_owned; F.gid = loc.levels[l].gid;
                    HostLevel C = loc.levels[l + 1].mesh; C.n_owned = loc.levels[l + 1].n_owned; C.gid = loc.levels[l + 1].gid;
                    TransferPlan T; build_transfer_plan(F, C, plans[l], plans[l + 1], T);
                    mixv(T.child_off); mixv(T.child_ids); mixv(T.parent); mixv(T.idist_own); mixv(T.ent_off); mixv(T.ent_src); mixv(T.ent_w);
                }
            }
            // in-kernel halo exchange: every rank "stores" the global id of its send-list nodes through
            // build_send_targets into the peers' row space; every ghost row must receive exactly its own node, exactly once
            for (size_t l = 0; l < full.levels.size(); l++) {
                std::vector<std::vector<long>> rows(nr), writes(nr);
                for (int p = 0; p < nr; p++) { rows[p].assign(all_plans[p][l].npad, -1); writes[p].assign(all_plans[p][l].npad, 0); }
                for (int r = 0; r < nr; r++) {
                    const LocalLevel& LL = locs[r].levels[l];
                    const LevelPlan& P = all_plans[r][l];
                    std::vector<int> send_rows(LL.send_idx.size());
                    for (size_t k = 0; k < LL.send_idx.size(); k++) send_rows[k] = int(P.new_of_old[LL.send_idx[k]]);
                    std::vector<PeerSlice> peers; std::vector<int> peer_rank;
                    for (int p = 0; p < nr; p++) {
                        const long ns = LL.send_off[p + 1] - LL.send_off[p], nrv = LL.recv_off[p + 1] - LL.recv_off[p];
                        if (p == r || (ns == 0 && nrv == 0)) continue;
                        PeerSlice ps; ps.send0 = LL.send_off[p]; ps.nsend = ns; ps.first_ghost_row = all_plans[p][l].npad_owned; ps.recv_off_me = locs[p].levels[l].recv_off[r];
                        peers.push_back(ps); peer_rank.push_back(p);
                    }
                    SendTargets st; build_send_targets(P.npad_owned, P.TN, send_rows, peers, st);
                    for (long node = 0; node < P.npad_owned; node++)
                        for (int k = st.off[node]; k < st.off[node + 1]; k++) {
                            const int p = peer_rank[st.peer[k]];
                            if (!st.tile_sends[node / P.TN]) { printf("tile flag missing\n"); return 1; }
                            if (st.row[k] < all_plans[p][l].npad_owned || st.row[k] >= all_plans[p][l].npad) { printf("target row outside the ghost rows\n"); return 1; }
                            rows[p][st.row[k]] = LL.gid[P.old_of_new[node]];
                            writes[p][st.row[k]]++;
                        }
                }
                for (int p = 0; p < nr; p++) {
                    const LocalLevel& LL = locs[p].levels[l];
                    const LevelPlan& P = all_plans[p][l];
                    for (long g = LL.n_owned; g < long(LL.gid.size()); g++) {
                        const long row = P.new_of_old[g];
                        if (rows[p][row] != LL.gid[g] || writes[p][row] != 1) { printf("ghost row %ld of rank %d (level %zu, %d ranks) holds node %ld, wants %ld, %ld writes\n", row, p, l, nr, rows[p][row], LL.gid[g], writes[p][row]); return 1; }
                    }
                }
            }
        }
        printf("kind %d hash %016llx\n", kind, H);
    }
    return 0;
}
