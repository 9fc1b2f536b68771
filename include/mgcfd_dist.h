/* mgcfd_dist.h -- multi-GPU runs of libmgcfd_b200.so: one process per GPU, the mesh split over the ranks by recursive
 * coordinate bisection, a halo exchange of node records after every Runge-Kutta stage / restriction / prolongation and of
 * coarse residuals before every prolongation (ncclSend/ncclRecv over NVLink, grouped per exchange), plus two scalar
 * all-reduces (min dt per smoothing visit: src/Kernels/cfd_loops.cpp:138-150; RMS sums per cycle: validation.cpp:91-105).
 * The reference has no distributed path at all (single process): nothing here replaces a reference
 * interface, it extends mgcfd_b200.h.  NCCL is loaded at run time (dlopen), so single-GPU users do not need it.
 *
 * Call order per rank:  mgcfd_create -> mgcfd_dist_init -> mgcfd_mesh_upload_partition (or your own partition through
 * mgcfd_upload_level + the lists) -> mgcfd_run_cycles / mgcfd_enqueue_cycles + mgcfd_collect (collective: every rank calls
 * them with the same arguments) -> mgcfd_get_field (local nodes: owned first, then ghosts; mgcfd_dist_global_ids maps them). */
#ifndef MGCFD_DIST_H
#define MGCFD_DIST_H
#include "mgcfd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* rank 0 creates the NCCL unique id; the host program broadcasts the 128 bytes to the other ranks (torch.distributed, MPI, a file) */
int mgcfd_dist_get_unique_id(char id[128]);
/* joins the communicator; after mgcfd_create, before any upload. CUDA graphs are switched off for distributed contexts. */
int mgcfd_dist_init(mgcfd_ctx* ctx, int rank, int nranks, const char id[128]);
/* info[0..7]: owned nodes, ghost nodes, nodes sent per exchange, nodes of the level over all ranks, rank, nranks, exchanges so far,
 * internal edges of the level over all ranks */
int mgcfd_dist_level_info(mgcfd_ctx* ctx, int level, long info[8]);
/* global node id of every local node (owned + ghost) */
int mgcfd_dist_global_ids(mgcfd_ctx* ctx, int level, long* gid);

/* Rank-local generation of the synthetic meshes (mgcfd_mesh.h: MGCFD_GEN_*): generates only the part of the mesh this rank holds --
 * its owned nodes, its ghosts, their edges and maps -- straight from the generator's node descriptions, never assembling the global
 * edge list, applies adjust_ewt / dampen_ewt, uploads and finalizes.  Same partition, bit for bit, as mgcfd_mesh_generate +
 * mgcfd_mesh_upload_partition (tests/test_partition.py), at a fraction of the host memory (what a 64 M-node mesh on 8 ranks needs).
 * dims = levels x 3 as for mgcfd_mesh_generate; lexicographic node numbering. */
int mgcfd_generate_upload_partition(mgcfd_ctx* ctx, int kind, int levels, const long* dims, const double lengths[3], int mesh_variant,
                                    double tilt);
/* host-only view of it (no device): the outputs of mgcfd_mesh_partition_plan; info[7] = a hash of everything the rank holds of the
 * level (nodes, edges with their weights -- adjusted and dampened iff apply_ewt != 0 --, maps, exchange lists) */
int mgcfd_generate_partition_plan(int kind, int levels, const long* dims, const double lengths[3], int mesh_variant, double tilt,
                                  int apply_ewt, int nranks, int rank, int level, long info[8], long* gid, long* send_counts,
                                  long* recv_counts, long* send_gids);

/* Direct peer-to-peer data path (optional, after mgcfd_mesh_upload_partition): every halo exchange and scalar all-reduce becomes
 * ONE kernel per rank that stores straight into the peers' memory over NVLink (CUDA IPC windows) and hand-shakes through
 * system-scope flags, instead of a packing kernel + NCCL calls.  prepare: allocates this rank's window, returns its 64-byte
 * cudaIpcMemHandle and a table of offsets (mgcfd_dist_p2p_table_len longs); the launcher all-gathers handles and tables in
 * rank order; attach: maps the peers' windows and switches the data path.  One process per GPU, all GPUs peer-accessible.
 * EXPERIMENTAL, off by default -- environment MGCFD_P2P_FUSED=1 on every rank: the Runge-Kutta stage kernels exchange their halo rows
 * themselves (stores into the peers' record buffers from the kernel that computes the rows, signal when the grid is done, wait
 * at the start of the next kernel that reads ghosts): no exchange kernel between the stages.  The offset table then also carries
 * the IPC handles of the record buffers (mgcfd_dist_p2p_table_len grows accordingly; nothing changes for the launcher).
 * Implemented and reviewed, not yet run on hardware. */
long mgcfd_dist_p2p_table_len(mgcfd_ctx* ctx);
int mgcfd_dist_p2p_prepare(mgcfd_ctx* ctx, char handle[64], long* table, long table_cap);
int mgcfd_dist_p2p_attach(mgcfd_ctx* ctx, const char* handles, const long* tables, long table_len);

#ifdef __cplusplus
}
#endif
#endif
