// partition.h -- host-side domain decomposition of a multigrid mesh over the GPUs of one node.
// The reference is a single process (its only "scaling" mechanism is -m mesh duplication, io_enhanced.cpp:89-201), so
// this is new integer work: the host implementation below is the definition, tests check its invariants
// (owned sets partition every level, ghost closure for flux / restrict / prolong, matching send / receive lists).
//
// Owner computes: every level is split into `nranks` parts by recursive coordinate bisection; a rank holds its owned nodes
// plus GHOST copies of every foreign node its kernels read:
//   flux      : the other endpoint of every internal edge of an owned node;
//   restrict  : every child (level l) of an owned coarse node (level l+1)          (mg_restrict reads all children);
//   prolong   : the coarse parents (level l+1) of every owned fine node and of all its edge neighbours
//               (prolong_residuals_interpolate_proper reads their residuals).
// Ghost nodes are never computed locally: their records / residuals are received from the owner.
#pragma once
#include <string>
#include <vector>

#include "host_mesh.h"

namespace mgcfd {

struct LocalLevel {
    HostLevel mesh;                    // local numbering: owned nodes first (ascending global id), then ghosts grouped by
                                       // owner rank (ascending), ascending global id inside a group.  Internal edges: every
                                       // global internal edge with at least one owned endpoint, in GLOBAL edge order;
                                       // boundary / wall edges of owned nodes.  mesh.mg: local fine -> local coarse (-1 if the
                                       // parent is not held, only for ghosts that do not need it).
    long n_owned = 0;
    long nI_global = 0;                // internal edges of the level over all ranks (work accounting)
    std::vector<long> gid;             // local -> global node id
    std::vector<long> edge_gid;        // local internal edge -> global edge index (ascending)
    std::vector<long> send_off;        // nranks+1: send_idx[send_off[p]..send_off[p+1]) goes to rank p
    std::vector<long> send_idx;        // local (owned) node indices, ascending global id per peer
    std::vector<long> recv_off;        // nranks+1: ghosts [n_owned + recv_off[p], n_owned + recv_off[p+1]) come from rank p
};

struct LocalMesh {
    int mesh_variant = 2;
    int rank = 0, nranks = 1;
    std::vector<LocalLevel> levels;
};

// owner rank of every node of one level: recursive coordinate bisection (widest extent, proportional split, ties by id)
void rcb_owners(const HostLevel& L, int nranks, std::vector<int>& owner);

// the part of `full` that rank `rank` of `nranks` holds (edge weights are taken as they are: apply adjust/dampen first)
void partition_mesh(const HostMesh& full, int nranks, int rank, LocalMesh& out);

// The same partition computed WITHOUT materialising the global mesh: every level is described by a streaming NodeSource (what a
// text mesh file holds: per node volume, coordinates and neighbour listing, host_mesh.h) and the fine -> coarse map by a table of
// ints; only per-node arrays (owners, coordinates for the bisection, maps) are held globally, never the edges.  Produces exactly
// what partition_mesh() produces from the assembled mesh (tests/test_partition.py compares them), at ~1/8 of the host memory:
// the way a 64 M-node level is split over 8 ranks.  Edge weights are raw (apply_ewt afterwards on the local edges).
struct LevelSource {
    const NodeSource* src = nullptr;
    std::vector<int> mg;               // fine -> coarse (global ids); empty on the coarsest level
    std::string name;
};
void partition_sources(const std::vector<LevelSource>& levels, int mesh_variant, int nranks, int rank, LocalMesh& out);

// In-kernel halo exchange (kernels.cuh: k_stage_pipe<.., DIST>, k_restrict<DIST>, k_prolong<DIST>): where the records of this rank's send-list nodes
// go.  The k-th node this rank sends to peer p is p's ghost row  first_ghost_row_p + recv_off_p[this rank] + k  (ghost rows follow
// the owned tiles, grouped by owner in rank order: partition.h).  Pure host arithmetic, shared by mgcfd_dist_p2p_attach and the
// host regression harness.
struct PeerSlice {
    long send0 = 0, nsend = 0;         // slice of the level's send list (device rows)
    long first_ghost_row = 0;          // the peer's first ghost row (its owned tiles * tile size)
    long recv_off_me = 0;              // the peer's recv_off[this rank]
};
struct SendTargets {
    std::vector<int> off, peer, row;   // CSR over this rank's owned rows: node -> (index into the peer list, row in that peer's arrays)
    std::vector<unsigned char> tile_sends;   // per tile: any node with a target
};
// the order in which the distributed stage kernel takes its tiles: the tiles that deliver rows or read a ghost row in their halo
// LAST (stable otherwise); returns how many those are.  halo_off / halo_ids: the plan's per-tile halo lists (device rows; rows
// >= owned_rows are ghosts).
int build_tile_order(long owned_rows, int tile_nodes, const SendTargets& st, const std::vector<long>& halo_off, const std::vector<int>& halo_ids,
                     long ntiles, std::vector<int>& order);
// throws std::runtime_error when a send-list entry is not an owned row
void build_send_targets(long owned_rows, int tile_nodes, const std::vector<int>& send_rows, const std::vector<PeerSlice>& peers, SendTargets& out);

}  // namespace mgcfd
