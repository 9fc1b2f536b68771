// plan.h -- host-side integer preprocessing of one mesh level for the B200 kernels:
// node renumbering (partition into tiles + RCM inside), per-tile halo lists, conflict-free edge colouring
// (ELL rounds), CSR views for the deterministic sorted-segment mode, and the multigrid transfer operators.
// The reference contains no renumbering or colouring code (SURVEY.md 8c "Integer work"): this host
// implementation is the definition; tests check its invariants (bijection, tile coverage, colour validity)
// and that un-permuted results equal the oracle's.
#pragma once
#include <cstdint>
#include <vector>

#include "host_mesh.h"

namespace mgcfd {

struct PlanOptions {
    int ordering = 2;     // MGCFD_ORDER_*
    int tile_nodes = 256; // owned nodes per tile (= CTA size of the tiled kernel)
    bool strict = true;   // false (introspection only): tiles whose halo cannot be staged are flagged (LevelPlan::oversize) instead of rejected
    bool conflict_free_rounds = true;  // segment mode: schedule each node's edges over the rounds so that quarter-warps avoid shared-memory bank conflicts
    bool scatter = false; // true: coloured-scatter rounds (each in-tile edge once); false: sorted-segment rounds (every edge from both ends)
    int visit_rounds = 2; // edge rounds per ring chunk of the visit kernel (the chunk lists in the descriptors are built for it)
    int visit_warps = 16; // warps per CTA of the visit kernel (16: one CTA per SM; 8: two): warp w takes the warp-tiles w, w + warps, ...
    int supers = 0;       // > 0 (segment mode, tile_nodes 128): group the tiles into this many SUPER-TILES (two-level bisection: first into
                          // super-tiles, then each into its tiles) and build the visit kernel's streams (VisitPlan below)
};

// What the persistent visit kernel (kernels.cuh k_visit) reads of a level.  A super-tile is a run of consecutive tiles (= consecutive
// rows) whose records are staged in shared memory TOGETHER: its halo is the set of rows outside the run that its edges touch, far
// fewer per owned row than the tile-by-tile halos of the stage kernel.
struct VisitPlan {
    int ns = 0;                            // super-tiles (0: not built)
    int maxt = 0;                          // most tiles in one super-tile
    int max_halo = 0, hpad = 0;            // most halo rows of one super-tile, padded to a multiple of 4
    int max_rounds = 0;                    // most rounds of one warp-tile
    long halo_total = 0;
    std::vector<long> super_off;           // ns+1: super-tile s = tiles [super_off[s], super_off[s+1])
    // The unit of work is a WARP-TILE: 32 consecutive rows (a quarter of a 128-row tile) that hold at least one node.  Edge rounds
    // per warp-tile, blocks of 32*34 bytes [hx[32] | hy[32] | hz[32] | wk[32] | code[32] (u16)], h as in LevelPlan::slots,
    //   wk = RN(sqrt(hx^2 + hy^2 + hz^2) * k2) with the IEEE square root (|e| * kdiss of flux_kernel.elemfunc.c:130 up to one rounding):
    //   the visit kernel's edge loop has no square root;
    //   code = (idx << 2) | ((idx >> 1) & 3) | (halo ? 0x8000 : 0), idx = row of the other endpoint inside the super-tile's own rows
    //   (halo = 0) or inside its halo list (halo = 1); (code & 0x7fff) << 4 is the byte offset of chunk 0 of that 64-byte row under
    //   the 64B swizzle, relative to the own-row / halo-row base.  Empty slot: the thread's own row, h = 0.
    std::vector<long> went_off;            // ns+1: warp-tiles (entries) of super-tile s = [went_off[s], went_off[s+1])
    long vblocks = 0;                      // 32-lane round blocks in total
    int max_ent = 0;                       // most warp-tiles in one super-tile
    std::vector<unsigned char> vslots;
    std::vector<double> hsum;              // [3][npad_owned]: sum of h over a row's internal edges (the A-side terms of the flux are hoisted out of the edge loop)
    // fixed-stride super-tile descriptors:
    //   {int row0, ntile, nhalo, tile0, nent, nchunk, 0, 0;                                        (32 bytes)
    //    max_ent x {int orow0, rounds, brounds, blane0; long long vblk0, bblk0};                   (32 bytes each; orow0 = first row
    //                 inside the super-tile, blane0 = first lane inside the 128-wide boundary blocks of its tile)
    //    unsigned short coff[warps + 1 (padded to 48 bytes)]; chunks of warp w = clist[coff[w] .. coff[w+1])
    //    max_chunk x unsigned vblk;                           the ring refills of warp w in the order it consumes them: warp w takes the
    //                 warp-tiles w, w + warps, ...; each is cut into chunks of exactly `rounds_per_chunk` rounds (a warp-tile's rounds are
    //                 padded with empty slots to a whole number of chunks)
    //    int halo_ids[hpad]}
    int rounds_per_chunk = 2, max_chunk = 0, warps = 16;
    std::vector<unsigned char> desc;
    int desc_stride = 0;
};

struct LevelPlan {
    long nel = 0, nI = 0, nB = 0, nW = 0;
    int TN = 256;
    long ntiles = 0, npad = 0;             // npad = rows of every node array: owned tiles (ntiles*TN) + ghosts
    long n_owned = 0, npad_owned = 0;      // tiles cover the owned nodes only; ghost g sits in row npad_owned + g
    std::vector<long> new_of_old;         // nel -> padded id
    std::vector<long> old_of_new;         // npad -> old id or -1 (padding)
    // ---- tiled, coloured flux structure -------------------------------------------------------
    std::vector<int> tile_nown;           // owned nodes per tile
    std::vector<long> halo_off;           // ntiles+1
    std::vector<int> halo_ids;            // padded global ids, sorted per tile
    bool scatter = false, oversize = false;
    // edge rounds: per tile `rounds` blocks of TN*26 bytes, block = [hx[TN] | hy[TN] | hz[TN] | other[TN] (uint16)],
    //   h = -0.5 * (edge vector oriented thread-node -> other); other = the row of the other endpoint in the tile's shared
    //   record buffer (local index o: < TN owned, >= TN halo) encoded as the byte offset of its chunk 0 under the 64B
    //   swizzle, (o << 6) | (((o >> 1) & 3) << 4).  Empty slot: scatter mode 0xFFFF; segment mode the thread's own row, h = 0.
    std::vector<long> slot_off;           // ntiles+1, in blocks (prefix sum of rounds)
    std::vector<unsigned char> slots;
    // boundary/wall rounds: blocks of TN*25 bytes, block = [x[TN] | y[TN] | z[TN] | kind[TN] (uint8: 0 none, 1 boundary, 2 wall)]
    std::vector<long> bslot_off;          // ntiles+1, in blocks
    std::vector<unsigned char> bslots;
    // fixed-stride tile headers: {int rounds, nh, brounds, pad; long slot_blk0, bslot_blk0; int halo_ids[hpad]}
    std::vector<unsigned char> hdrs;
    int hdr_stride = 0, hpad = 0;
    int max_halo = 0, max_rounds = 0;
    long cut_edges = 0, used_slots = 0;
    // ---- flat edge list in new numbering (atomic mode, indirect_rw, ordering sweeps) -------------
    std::vector<int> ea, eb;              // internal edges, original edge order
    std::vector<double> ew;               // [3][nI]
    std::vector<int> bnode;               // boundary+wall edges: node, kind
    std::vector<uint8_t> bkind;
    std::vector<double> bw;               // [3][nB+nW]
    // ---- CSR by node, original edge order (segment rounds + prolong) -----------------------------
    std::vector<long> adj_off;            // npad+1
    std::vector<int> adj_nbr;             // neighbour padded id, bit31 set when this node is the edge's `b`
    VisitPlan visit;                      // PlanOptions::supers > 0
};

void build_level_plan(const HostLevel& L, const PlanOptions& opt, LevelPlan& P);
long check_colouring(const LevelPlan& P);   // number of write conflicts (must be 0)

// multigrid operators between level l (fine) and l+1 (coarse), in padded new ids of both
struct TransferPlan {
    // restrict: children of each coarse node in ascending ORIGINAL fine index (the reference's summation order)
    std::vector<long> child_off;          // npad_c+1
    std::vector<int> child_ids;           // fine padded ids
    // prolong: per fine node, entries in ORIGINAL edge order
    std::vector<int> parent;              // npad_f: own coarse parent (padded id), -1 for padding
    std::vector<double> idist_own;        // npad_f: 1/dist(parent, node); -1.0 if coincident (exact ==)
    std::vector<long> ent_off;            // npad_f+1
    std::vector<int> ent_src;             // coarse padded id whose residual is multiplied (the b-side quirk is baked in)
    std::vector<double> ent_w;            // 1/dist(parent of neighbour, node)
};
void build_transfer_plan(const HostLevel& fine, const HostLevel& coarse, const LevelPlan& Pf, const LevelPlan& Pc, TransferPlan& T);

// Host walk of the device data structures (checking aid for the preprocessing, not a compute path): what the stage kernel's threads
// accumulate from the tile headers / round blocks / boundary blocks of P for the state `var` (AoS, original node order) -> flux
// (AoS, original order; mask: bit0 internal, bit1 boundary, bit2 wall edges), and what k_restrict / k_prolong compute from T.
void emulate_stage_flux(const LevelPlan& P, const double* var, int mask, const double ff[5], const double ffc[12], double k2, double* flux);
void emulate_visit_flux(const LevelPlan& P, const double* var, int mask, const double ff[5], const double ffc[12], double k2, double* flux);
void emulate_restrict(const LevelPlan& Pf, const LevelPlan& Pc, const TransferPlan& T, const double* var_f, double* var_c);
void emulate_prolong(const LevelPlan& Pf, const LevelPlan& Pc, const TransferPlan& T, const double* res_c, const double* res_f, double* var_f);

}  // namespace mgcfd
