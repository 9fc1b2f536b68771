/* mgcfd_dist.h -- multi-GPU runs of libmgcfd_b200.so: one process per GPU, the mesh split over the ranks by recursive
 * coordinate bisection.  What crosses ranks: the node records of ghost nodes after every Runge-Kutta stage, restriction and
 * prolongation, the coarse residuals of ghost nodes before a prolongation, and two scalar all-reduces (min dt per smoothing
 * visit: src/Kernels/cfd_loops.cpp:138-150; RMS sums per cycle: validation.cpp:91-105).  Two data planes, same results:
 *   - peer-to-peer (mgcfd_dist_p2p_prepare / _attach, the default of bench.py and the driver): every kernel that PRODUCES a row
 *     stores it straight into the other ranks' copies over NVLink (one CUDA IPC slab per rank) and every kernel that READS ghost
 *     rows waits for its neighbours' epoch flags as late as it can -- no exchange kernel is left in a V-cycle; with
 *     the optional persistent visit kernel the grid barriers double as the halo exchange; the whole distributed V-cycle is one
 *     replayed CUDA graph;
 *   - NCCL (no attach): packing kernels + grouped ncclSend/ncclRecv + ncclAllReduce between the stage kernels, launched eagerly
 *     (MGCFD_DIST_GRAPH=1 captures them too).
 * The reference has no distributed path at all (single process): nothing here replaces a reference
 * interface, it extends mgcfd_b200.h.  NCCL is loaded at run time (dlopen), so single-GPU users do not need it.
 *
 * Call order per rank:  mgcfd_create -> mgcfd_dist_init -> mgcfd_mesh_upload_partition (or mgcfd_generate_upload_partition)
 * [-> mgcfd_dist_p2p_prepare, all-gather, mgcfd_dist_p2p_attach] -> mgcfd_run_cycles / mgcfd_enqueue_cycles + mgcfd_collect
 * (collective: every rank calls them with the same arguments) -> mgcfd_get_field (local nodes: owned first, then ghosts;
 * mgcfd_dist_global_ids maps them). */
#ifndef MGCFD_DIST_H
#define MGCFD_DIST_H
#include "mgcfd_b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* rank 0 creates the NCCL unique id; the host program broadcasts the 128 bytes to the other ranks (torch.distributed, MPI, a file) */
int mgcfd_dist_get_unique_id(char id[128]);
/* joins the communicator; after mgcfd_create, before any upload.  Until mgcfd_dist_p2p_attach the cycle is launched eagerly (NCCL
 * calls between kernels; MGCFD_DIST_GRAPH=1 captures them); after the attach it is replayed as a CUDA graph like a single-GPU run. */
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

/* Direct peer-to-peer data path (after the upload): prepare returns the 64-byte cudaIpcMemHandle of this rank's slab (window of
 * flags / reduction slots / staging + the record buffers and residual planes of every level) and a table of offsets
 * (mgcfd_dist_p2p_table_len longs); the launcher all-gathers handles and tables in rank order; attach maps the peers' slabs, builds
 * the row -> (peer, remote row) tables and switches the data plane.  From then on the stage kernels, restrict, prolong (and the visit
 * kernel) deliver the rows they produce themselves and synchronise through epoch numbers in system-scope flags (every rank runs the same kernel
 * sequence, so the epochs advance alike on all ranks -- also on a rank that has no halo at some level).  One process per GPU, all
 * GPUs peer-accessible (NVLink / NVSwitch).  MGCFD_NO_P2P=1 in the launchers keeps the NCCL plane. */
long mgcfd_dist_p2p_table_len(mgcfd_ctx* ctx);
int mgcfd_dist_p2p_prepare(mgcfd_ctx* ctx, char handle[64], long* table, long table_cap);
int mgcfd_dist_p2p_attach(mgcfd_ctx* ctx, const char* handles, const long* tables, long table_len);

#ifdef __cplusplus
}
#endif
#endif
