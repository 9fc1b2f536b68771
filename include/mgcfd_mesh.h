/* mgcfd_mesh.h -- host-side mesh utilities of libmgcfd_b200.so: the reference's mesh interchange formats
 * (text mesh + .coords + MG map + input.dat, src/Base/io.cpp:14-199, src/Base/io_enhanced.cpp:407-650; the .bin
 * cache, io_enhanced.cpp:203-405) and the synthetic generators for the BASELINE.json configs.  A mgcfd_mesh holds
 * every level exactly as read_grid / read_mg_connectivity would leave it in memory. */
#ifndef MGCFD_MESH_H
#define MGCFD_MESH_H
#include "mgcfd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

typedef struct mgcfd_mesh mgcfd_mesh;

#define MGCFD_GEN_HEX_BOX 0   /* vertex-centred hex dual, 6 neighbours (~3 edges/node: Onera-M6-like density) */
#define MGCFD_GEN_TET_BOX 1   /* Kuhn tetrahedra, 14 neighbours (7 edges/node) */
#define MGCFD_GEN_TET_CELLS 2 /* cell-centred tetrahedra, 4 faces per cell (fvcorr.domn.097K-like), single level */

/* dims = levels x 3 node counts (cube counts for TET_CELLS); ordering 0 = lexicographic, 1 = seeded random permutation */
int mgcfd_mesh_generate(int kind, int levels, const long* dims, const double lengths[3], int mesh_variant, int ordering,
                        unsigned long seed, double tilt, mgcfd_mesh** out);
/* read_input_dat + read_grid (+ .coords) + read_mg_connectivity; prefers <layer>.bin when present */
int mgcfd_mesh_load(const char* input_dat, const char* input_directory, mgcfd_mesh** out);
/* writes input.dat, one text mesh (+ .coords) per level and the MG maps; binary != 0 additionally writes <layer>.bin */
int mgcfd_mesh_write(const mgcfd_mesh* m, const char* directory, const char* input_dat_name, int binary);
int mgcfd_mesh_levels(const mgcfd_mesh* m);
int mgcfd_mesh_variant(const mgcfd_mesh* m);
/* out = nel, num_internal, num_boundary, num_wall, mgc */
int mgcfd_mesh_dims(const mgcfd_mesh* m, int level, long out[5]);
/* what: 0 volumes (double[nel]), 1 edges (edge_neighbour[nI+nB+nW]), 2 coords (double[3*nel]), 3 mg map (long[mgc]) */
const void* mgcfd_mesh_ptr(const mgcfd_mesh* m, int level, int what);
/* adjust_ewt + dampen_ewt on every level, once (euler3d_cpu_double.cpp:337-352) */
int mgcfd_mesh_apply_ewt(mgcfd_mesh* m);
/* mgcfd_upload_level for every level (applying the edge-weight adjustment first if not yet done) + mgcfd_finalize */
int mgcfd_mesh_upload(mgcfd_mesh* m, mgcfd_ctx* ctx);
/* distributed (mgcfd_dist.h): keeps only this rank's part of every level (owned nodes + ghosts), uploads it with its halo
 * exchange lists and finalizes; the context must have joined a communicator with mgcfd_dist_init */
int mgcfd_mesh_upload_partition(mgcfd_mesh* m, mgcfd_ctx* ctx);
/* host-only view of that partition for one rank and level (no device, no NCCL): info[0..7] = owned, ghosts, sent nodes, global
 * nodes, local internal / boundary / wall edges, a hash of everything the rank holds of the level; gid[owned+ghosts]; send_counts / recv_counts[nranks]; send_gids[sent nodes] =
 * global ids in send order. Any output pointer may be NULL. */
int mgcfd_mesh_partition_plan(mgcfd_mesh* m, int nranks, int rank, int level, long info[8], long* gid, long* send_counts,
                              long* recv_counts, long* send_gids);
/* Host-only check of the in-kernel halo exchange's tables for the whole partition at once (no device, one process): every rank's
 * levels are numbered as mgcfd_finalize numbers them, every rank's row -> (peer, remote row) targets and tile order are built as
 * mgcfd_dist_p2p_attach builds them, and the delivery is replayed on global node ids.  out[0] rows delivered, out[1] ghost rows over
 * all ranks and levels (equal when every ghost row has exactly one producer), out[2] errors (a row delivered to a ghost row holding
 * another node, a ghost row hit twice or never, a broken tile order), out[3] tiles that read a ghost row without being among the
 * tiles the stage kernel takes last, out[4] transfer-kernel blocks that wait for a peer.  tile_nodes = 0: the automatic choice. */
int mgcfd_mesh_delivery_check(mgcfd_mesh* m, int nranks, int tile_nodes, long out[5]);
/* -m / --mesh-duplicate-count: `count` independent copies of every level, laid out as duplicate_mesh (io_enhanced.cpp:89-201) */
int mgcfd_mesh_duplicate(mgcfd_mesh* m, int count);
void mgcfd_mesh_free(mgcfd_mesh* m);
const char* mgcfd_mesh_last_error(void);

#ifdef __cplusplus
}
#endif
#endif
