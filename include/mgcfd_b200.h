/* mgcfd_b200.h -- C ABI of libmgcfd_b200.so: the B200-native (sm_100a, fp64) replacement for the
 * per-cycle solver loop of MG-CFD (warwick-hpsc/MG-CFD-app-plain).
 *
 * The reference has no plugin/FFI layer: its hot path is a set of free C++ functions on raw host arrays,
 * called only from main()'s V-cycle loop (src/euler3d_cpu_double.cpp:371-694).  This header is the boundary
 * a maintainer binds instead: one entry point per reference function, same argument meaning, but on
 * DEVICE-RESIDENT level state owned by an opaque context (passing host arrays per call would make every
 * kernel PCIe-bound).  Each declaration cites the reference interface it replaces.  INTEGRATION.md shows
 * the main() a maintainer would write against it.
 *
 * Conventions
 *  - plain C, no torch / C++ types in any signature; every function returns an int status (0 = ok).
 *  - host arrays use the REFERENCE layout and node order: node arrays AoS double[nel*5] (rho, mx, my, mz,
 *    rhoE; src/Base/const.h:19-26), edges AoS `edge_neighbour` {long a,b; double x,y,z} (40 B,
 *    src/Base/definitions.h:83) stored [internal | boundary(-1) | wall(-2)] (src/Base/io.cpp:149-181),
 *    coords AoS double3, MG map long[nel_fine].  The library copies what it is given; callers keep ownership.
 *  - the library renumbers nodes/edges internally (partition + RCM, colouring); everything it returns
 *    is un-permuted back to the reference order.
 *  - not re-entrant per context; one host thread drives one context (as the reference's main() does).
 *  - there is NO CPU fallback: without a usable CUDA device mgcfd_create fails with MGCFD_ERR_NO_DEVICE.
 */
#ifndef MGCFD_B200_H
#define MGCFD_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define MGCFD_NVAR 5
#define MGCFD_RK 3

/* status codes */
#define MGCFD_OK 0
#define MGCFD_ERR_CUDA 1              /* a CUDA runtime call failed; see mgcfd_last_error() */
#define MGCFD_ERR_ARG 2               /* bad argument / call order */
#define MGCFD_ERR_INVALID_VARIABLES 3 /* NaN/Inf/negative density or energy (validation.cpp:107-138) */
#define MGCFD_ERR_NO_DEVICE 4
#define MGCFD_ERR_IO 5
#define MGCFD_ERR_COMM 6

/* mesh variants: values of the reference's MESH_* (src/Base/const.h:38-41) */
#define MGCFD_MESH_FVCORR 0
#define MGCFD_MESH_M6_WING 2
#define MGCFD_MESH_LA_CASCADE 3
#define MGCFD_MESH_ROTOR_37 4

/* device fields addressable by mgcfd_get_field / mgcfd_set_field */
#define MGCFD_FIELD_VARIABLES 0
#define MGCFD_FIELD_OLD_VARIABLES 1
#define MGCFD_FIELD_RESIDUALS 2
#define MGCFD_FIELD_FLUXES 3
#define MGCFD_FIELD_STEP_FACTORS 4 /* 1 double per node */
#define MGCFD_FIELD_VOLUMES 5      /* 1 double per node, read-only */

/* how the flux scatter is made race-free */
#define MGCFD_FLUX_TILED_COLOURED 0 /* node tiles staged in shared memory; an edge inside a tile is evaluated once and scattered to its other end through a shared accumulator, conflict-free by edge colouring; no atomics, deterministic */
#define MGCFD_FLUX_SORTED_SEGMENT 1 /* node tiles staged in shared memory; every node walks its CSR segment (original edge order), every edge is evaluated from both ends, nothing is scattered; no atomics, no barriers in the edge loop, deterministic */
#define MGCFD_FLUX_ATOMIC 2         /* one thread per edge in original edge order, fp64 atomics (baseline; the ordering-sweep kernel) */

/* node renumbering applied at upload */
#define MGCFD_ORDER_AS_GIVEN 0
#define MGCFD_ORDER_RCM 1
#define MGCFD_ORDER_PARTITION_RCM 2 /* default: recursive bisection into tiles, RCM inside each tile */

typedef struct mgcfd_ctx mgcfd_ctx;

typedef struct mgcfd_options {
    int device;     /* CUDA device ordinal (default 0) */
    int flux_mode;  /* MGCFD_FLUX_* (default MGCFD_FLUX_SORTED_SEGMENT: the faster of the two tiled modes on B200, profiles/) */
    int ordering;   /* MGCFD_ORDER_* */
    int tile_nodes; /* owned nodes per tile = threads per CTA of the stage kernels; 0 = auto (default: 128 for the visit kernel and for
                       multi-million-node levels, else 128 or 256 by a wave model, see auto_tile_nodes in csrc/context.cu); else 128, 256 or 512 */
    int use_graph;  /* run_cycles replays one captured CUDA graph per V-cycle (default 1) */
    int timing;     /* record CUDA-event times per kernel per level (forces use_graph=0) */
    int no_pipeline; /* 1: fused stages use the simple one-CTA-per-tile kernel instead of the persistent kernel whose transfers
                        (TMA bulk copies of the edge stream, cp.async gathers of node records) run one tile ahead (default 0) */
    int no_pdl;      /* 1: stage kernels are launched without programmatic dependent launch (default 0: the prologue of a stage
                        kernel -- barrier set-up, header and edge-stream prefetch -- overlaps the tail of its predecessor) */
    int visit;       /* 1: smoothing visits of levels up to ~1.2 M nodes run the persistent visit kernel (ONE launch per visit: minimum dt,
                        the three RK stages, residual and RMS sums separated by grid barriers, node records resident in shared memory
                        where they fit; multi-GPU: the grid barriers double as the halo exchange).  Default 0: one stage kernel per RK
                        stage -- measured faster on B200 for the BASELINE workloads (profiles/r02*).  The environment
                        variable MGCFD_VISIT=0/1 overrides it. */
    int reserved[7];
} mgcfd_options;

void mgcfd_default_options(mgcfd_options* opt);
const char* mgcfd_last_error(void);
const char* mgcfd_version(void);
/* Debugging aid (no reference counterpart).  With MGCFD_GUARD=1 in the environment when contexts are created, every device
   allocation of the library is bracketed by 64 KB zones of a known byte pattern; this call reads the zones of all live contexts
   back and returns how many were written to (0 = no kernel wrote past the end or before the start of an array), with a short
   description of each in report[0..cap).  Without MGCFD_GUARD it returns 0. */
int mgcfd_guard_check(char* report, int cap);
/* the check's own test: damages one guard zone of the context (one byte, 100 bytes past the end of a 64-byte array); an error
   unless the context was created under MGCFD_GUARD=1 */
int mgcfd_guard_selftest(mgcfd_ctx* ctx);

/* ---- lifetime --------------------------------------------------------------------------------- */
/* replaces the per-level array set-up in main() (euler3d_cpu_double.cpp:138-243) */
int mgcfd_create(int levels, int mesh_variant, const mgcfd_options* opt, mgcfd_ctx** out);
int mgcfd_destroy(mgcfd_ctx* ctx);

/* globals ff_variable[5] and ff_flux_contribution_{momentum_x,momentum_y,momentum_z,density_energy}
 * (src/Base/globals.h:10-14) made explicit; ffc = 4 x double3 in that order.  Per context (the values travel in the kernel
 * arguments): contexts on one device may hold different far fields.  mgcfd_create sets the reference's values
 * (mgcfd_far_field_conditions); call this before mgcfd_finalize if the initial node state is to start from other ones. */
int mgcfd_set_farfield(mgcfd_ctx* ctx, const double ff_variable[5], const double ff_flux_contribution[12]);
/* host evaluation of initialize_far_field_conditions (src/Kernels/cfd_loops.h:85-119) */
void mgcfd_far_field_conditions(double ff_variable[5], double ff_flux_contribution[12]);

/* Takes one level exactly as read_grid + read_mg_connectivity leave it (io.cpp:14-199,
 * io_enhanced.cpp:629-650), AFTER adjust_ewt/dampen_ewt (validation.cpp:28-75) if the mesh variant asks
 * for them (mgcfd_adjust_dampen_ewt does that on the host).  coords may be NULL iff levels == 1;
 * mg_map NULL on the coarsest level. */
int mgcfd_upload_level(mgcfd_ctx* ctx, int level, long nel, const double* volumes, const double* coords_xyz,
                       long num_internal, long num_boundary, long num_wall, const void* edges_aos40,
                       const long* mg_map, long mgc);
/* builds MG operators (needs all levels), allocates node state, zeroes it (euler3d_cpu_double.cpp:234-243) */
int mgcfd_finalize(mgcfd_ctx* ctx);
/* adjust_ewt + dampen_ewt in place on a host edge array, by mesh variant (euler3d_cpu_double.cpp:337-352) */
int mgcfd_adjust_dampen_ewt(int mesh_variant, const double* coords_xyz, long num_edges, void* edges_aos40);

/* ---- one entry point per reference function (granular path) -------------------------------------- */
int mgcfd_initialize_variables(mgcfd_ctx* ctx, int level);      /* cfd_loops.h:44-55 */
int mgcfd_copy_old_variables(mgcfd_ctx* ctx, int level);        /* copy<double>(old, variables), common.h:100-112 */
int mgcfd_compute_step_factor(mgcfd_ctx* ctx, int level, int legacy); /* cfd_loops.cpp:76-157 / :13-73 (legacy) */
int mgcfd_compute_flux_edge(mgcfd_ctx* ctx, int level);          /* flux_loops.cpp:78-153 */
int mgcfd_compute_boundary_flux_edge(mgcfd_ctx* ctx, int level); /* flux_loops.cpp:10-42 */
int mgcfd_compute_wall_flux_edge(mgcfd_ctx* ctx, int level);     /* flux_loops.cpp:44-76 */
int mgcfd_time_step(mgcfd_ctx* ctx, int level, int rk_stage);    /* cfd_loops.cpp:215-280 */
int mgcfd_zero_fluxes(mgcfd_ctx* ctx, int level);                /* cfd_loops.cpp:282-305 */
int mgcfd_indirect_rw(mgcfd_ctx* ctx, int level);                /* indirect_rw_loop.cpp:11-78 (bandwidth probe) */
int mgcfd_residual(mgcfd_ctx* ctx, int level);                   /* validation.cpp:77-89 */
/* calc_rms (validation.cpp:91-105) plus the per-variable RMS the north star asks for */
int mgcfd_calc_rms(mgcfd_ctx* ctx, int level, double* rms_all, double rms_var[5]);
/* check_for_invalid_variables (validation.cpp:107-138): returns MGCFD_ERR_INVALID_VARIABLES and the first
 * offending cell (reference node order) + reason (1 NaN/Inf, 2 negative density, 3 negative energy). */
int mgcfd_check_for_invalid_variables(mgcfd_ctx* ctx, int level, long* first_bad_cell, int* reason);
/* mg_restrict(variables[l-1] -> variables[l]) (mg_loops.cpp:30-202), l = coarse level */
int mgcfd_mg_restrict(mgcfd_ctx* ctx, int coarse_level);
/* prolong_residuals_interpolate_proper(level l+1 -> l) (mg_loops.cpp:678-864), l = fine level */
int mgcfd_prolong(mgcfd_ctx* ctx, int fine_level);

/* ---- fused fast path ------------------------------------------------------------------------------ */
/* Runs `ncycles` iterations of main()'s loop (euler3d_cpu_double.cpp:371-694) entirely on the device.
 * rms_all[c]   = the value main() prints as "(RMS = ...)" for cycle c  (may be NULL)
 * rms_var[c*5+v] = per-variable residual RMS of level 0               (may be NULL)
 * Returns MGCFD_ERR_INVALID_VARIABLES if a NaN/negative state appeared (checked once per call). */
int mgcfd_run_cycles(mgcfd_ctx* ctx, int ncycles, double* rms_all, double* rms_var);
/* The same loop split in two so that a caller can keep the device busy: enqueue (no host synchronisation; at most 4096
 * cycles may be outstanding) and collect (one synchronisation: RMS histories of the cycles enqueued since the last
 * collect + the invalid-state verdict). mgcfd_run_cycles == enqueue + collect. */
int mgcfd_enqueue_cycles(mgcfd_ctx* ctx, int ncycles);
int mgcfd_collect(mgcfd_ctx* ctx, double* rms_all, double* rms_var);
/* after MGCFD_ERR_INVALID_VARIABLES from run_cycles/collect: the first offending cell (reference node order) of the first
 * offending RK stage and the reason, i.e. what check_for_invalid_variables would have printed before exit(1) */
int mgcfd_invalid_cell(mgcfd_ctx* ctx, long* cell, int* reason);

/* ---- host <-> device state, reference layout ------------------------------------------------------ */
int mgcfd_get_field(mgcfd_ctx* ctx, int level, int field, double* host_out);
int mgcfd_set_field(mgcfd_ctx* ctx, int level, int field, const double* host_in);
int mgcfd_synchronize(mgcfd_ctx* ctx);
/* the cudaStream_t every kernel of this context is launched on (for callers that bracket calls with their own CUDA events) */
int mgcfd_get_stream(mgcfd_ctx* ctx, void** cuda_stream);

/* ---- introspection (tests, Times.csv, roofline) --------------------------------------------------- */
/* info[0..15]: nel, nI, nB, nW, padded nodes, tiles, tile_nodes, max rounds, slots allocated, halo entries, cut edges,
 * slots used, max halo of a tile, boundary slots allocated, shared memory per CTA (bytes), persistent grid (0 = simple kernel) */
int mgcfd_level_info(mgcfd_ctx* ctx, int level, long info[16]);
/* the persistent visit kernel's configuration of a level: info[0..7] = in use (0/1), super-tiles per CTA, CTAs, edge rounds per ring
 * entry, own rows resident in shared memory (0/1), rows of the largest super-tile, shared memory per CTA (bytes), halo rows in total */
int mgcfd_visit_info(mgcfd_ctx* ctx, int level, long info[8]);
/* development aid (MGCFD_VISIT_DEBUG=1 in the environment at mgcfd_create): clock stamps of the most recent visit-kernel launch, 64
 * per CTA (tools/visit_timeline.py decodes them); returns the number of CTAs' worth of stamps copied or a negative error */
int mgcfd_visit_debug(mgcfd_ctx* ctx, long long* out, long cap);
/* new_of_old[nel]: the node renumbering (a bijection onto [0,padded) minus padding) */
int mgcfd_get_permutation(mgcfd_ctx* ctx, int level, long* new_of_old);
/* verifies on the host that no two edges of one colour round of one tile write the same node; returns #conflicts */
long mgcfd_check_colouring(mgcfd_ctx* ctx, int level);
/* accumulated CUDA-event milliseconds, reference kernel ids (src/Base/const.h:30-37):
 * out[kernel*levels + level], kernels = compute_step, flux, update(0), indirect_rw, time_step, restrict, prolong */
int mgcfd_get_times(mgcfd_ctx* ctx, double* out_ms, long* out_iters);
int mgcfd_reset_times(mgcfd_ctx* ctx);
/* switches per-kernel CUDA-event timing on/off (on: every kernel is launched individually, graphs are bypassed) */
int mgcfd_set_timing(mgcfd_ctx* ctx, int on);
/* launches issued by this context since creation (all kernels are ours) */
long mgcfd_launch_count(mgcfd_ctx* ctx);
/* time (ms, CUDA events on the context's stream) of `reps` back-to-back launches of one kernel on `level`:
 * which = 0 fused stage kernel (flux + boundary + wall + time_step) of the configured tiled mode, 1 its flux-only form
 * (internal edges, granular API), 2 indirect_rw, 3 atomic flux, 5 the stage kernel without its edge rounds, 16 + bits the
 * assess-compute variant `bits` (mgcfd_flux_variant) */
int mgcfd_time_kernel(mgcfd_ctx* ctx, int level, int which, int reps, double* ms_total);

/* The reference's assess-compute protocol (run-inputs/assess-compute*.json): compute_flux_edge's arithmetic in the forms
 * its compile-time toggles select (src/Kernels/flux_kernel.elemfunc.c), as a benchmark kernel -- one thread per internal edge,
 * atomics, the same memory traffic for every variant.  bits: 1 = FLUX_REUSE_DIV (:46-71), 2 = FLUX_REUSE_FACTOR + FLUX_REUSE_FLUX
 * (:132-190), 4 = FLUX_PRECOMPUTE_EDGE_WEIGHTS (:24-28); 0 is the default build's arithmetic as written.  Accumulates into
 * `fluxes` like mgcfd_compute_flux_edge; mgcfd_time_kernel times variant `bits` as selector 16 + bits.
 * bits = 8: all three toggles with the node state gathered from an SoA copy (five planes) instead of the 64-byte node records --
 * the layout A/B of DESIGN.md 3 (no reference counterpart; same results as bits = 7, bit for bit). */
int mgcfd_flux_variant(mgcfd_ctx* ctx, int level, int bits);

/* Host-only run of the integer preprocessing (no device needed): renumbering, tiling and colouring of one level.
 * info[] as mgcfd_level_info, except info[15] = a hash of everything the device would receive of the level (the plan is a pure
 * function of the mesh: tests compare it across runs and thread counts); new_of_old may be NULL; *conflicts = result of the
 * colouring validity check. */
int mgcfd_plan_level(long nel, const double* coords_xyz, long num_internal, long num_boundary, long num_wall,
                     const void* edges_aos40, int ordering, int tile_nodes, int flux_mode, long info[16], long* new_of_old,
                     long* conflicts);
/* Host-only checking aids (no device; NOT a compute path -- nothing in the solver calls them): the byte streams the plan hands to
 * the device (tile headers, edge round blocks, boundary blocks; the restrict / prolong operators) walked on the host exactly as
 * the kernels' threads walk them, with plain libm arithmetic.  Arrays are AoS in the reference's node order.
 *   flux:      fluxes[nel*5] (overwritten) = what compute_flux_edge (mask bit 0), compute_boundary_flux_edge (bit 1) and
 *              compute_wall_flux_edge (bit 2) accumulate from zero for `variables`;
 *   transfers: var_c[nel_c*5] (in/out) = mg_restrict(var_f); var_f_out[nel_f*5] = prolong_residuals_interpolate_proper applied to
 *              var_f with res_c / res_f. */
int mgcfd_plan_emulate_flux(long nel, const double* coords_xyz, long num_internal, long num_boundary, long num_wall, const void* edges_aos40,
                            int ordering, int tile_nodes, int flux_mode, const double* variables, int mask, double* fluxes);
/* the same walk over the visit kernel's streams (super-tile descriptors, halo lists, re-addressed edge rounds): `supers` super-tiles
 * of 128-node tiles; info[0..7] = super-tiles, most tiles in one, most halo rows of one, halo rows in total, most rounds, tiles, rows */
int mgcfd_plan_emulate_visit_flux(long nel, const double* coords_xyz, long num_internal, long num_boundary, long num_wall, const void* edges_aos40,
                                  int supers, const double* variables, int mask, double* fluxes, long info[8]);
/* host-only: the visit-kernel configuration mgcfd_upload_level would choose for this level on a device with `num_sms` SMs; info[] as
 * mgcfd_visit_info */
int mgcfd_plan_visit_config(long nel, const double* coords_xyz, long num_internal, long num_boundary, long num_wall, const void* edges_aos40,
                            int num_sms, long info[8]);
int mgcfd_plan_emulate_transfers(long nel_f, const double* coords_f, long nI_f, long nB_f, long nW_f, const void* edges_f, const long* mg_map,
                                 long nel_c, const double* coords_c, long nI_c, long nB_c, long nW_c, const void* edges_c, int ordering,
                                 int tile_nodes, const double* var_f, const double* res_f, const double* res_c, double* var_c, double* var_f_out);
void mgcfd_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* MGCFD_B200_H */
