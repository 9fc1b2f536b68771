// oracle/ref_shim.cpp -- TEST INFRASTRUCTURE ONLY (never linked into the product library).
//
// A thin extern "C" veneer that is compiled TOGETHER WITH the unmodified reference sources where they
// lie under $(REF)/src (see oracle/Makefile, target `ref`).  It contains no arithmetic of its own: every
// number it returns is produced by the reference's functions (src/Kernels/*.cpp, src/Base/io*.cpp).
// It exists because the reference has no library/FFI surface: its kernels are free C++
// functions driven only by main() (src/euler3d_cpu_double.cpp:371-694) and depend on a handful of
// globals that main()'s translation unit defines (src/euler3d_cpu_double.cpp:42-50); the shim defines
// those globals instead of main() so the kernels can be called on in-memory arrays from the tests.
//
// The V-cycle sequencing in refs_run() follows src/euler3d_cpu_double.cpp:371-694 call for call
// (copy, step factor, RK x (flux, boundary, wall, time_step, invalid check), residual, rms, transfer);
// the results-neutral indirect_rw probe + zero_fluxes pair (:491-505) is optional.
#include <omp.h>
#include <string.h>
#include <string>
#include <vector>

#include "common.h"
#include "io.h"
#include "io_enhanced.h"
#include "flux_loops.h"
#include "indirect_rw_loop.h"
#include "cfd_loops.h"
#include "mg_loops.h"
#include "validation.h"

// globals normally owned by main()'s translation unit (src/euler3d_cpu_double.cpp:42-50)
int levels = 0;
int level = 0;
int current_kernel;
int mesh_variant;
double ff_variable[NVAR];
double3 ff_flux_contribution_momentum_x;
double3 ff_flux_contribution_momentum_y;
double3 ff_flux_contribution_momentum_z;
double3 ff_flux_contribution_density_energy;

struct ref_level {
    long nel, ne, nI, nB, nW, iS, bS, wS, mgc;
    double* volumes; edge_neighbour* edges; double3* coords; long* mg;
    double *variables, *old_variables, *residuals, *fluxes, *step_factors;
};
struct ref_session {
    std::vector<ref_level> L;
    long* up_scratch;
    double t_flux[8], t_step[8], t_time[8], t_restrict[8], t_prolong[8], t_total;
};

static void alloc_state(ref_level& v) {
    v.variables = alloc<double>(v.nel*NVAR);     zero_array(v.nel*NVAR, v.variables);
    v.residuals = alloc<double>(v.nel*NVAR);     zero_array(v.nel*NVAR, v.residuals);
    v.old_variables = alloc<double>(v.nel*NVAR); zero_array(v.nel*NVAR, v.old_variables);
    v.fluxes = alloc<double>(v.nel*NVAR);        zero_array(v.nel*NVAR, v.fluxes);
    v.step_factors = alloc<double>(v.nel);       zero_array(v.nel, v.step_factors);
}
static void free_state(ref_level& v) {
    dealloc(v.variables); dealloc(v.residuals); dealloc(v.old_variables); dealloc(v.fluxes); dealloc(v.step_factors);
}

extern "C" {

int refs_sizeof_edge() { return (int)sizeof(edge_neighbour); }
int refs_omp_threads() {
#ifdef OMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}
void refs_set_globals(int nlevels, int variant) { levels = nlevels; mesh_variant = variant; level = 0; }
void refs_set_level(int l) { level = l; }

void refs_far_field(double* ffv, double* ffc) {
    initialize_far_field_conditions();
    for (int v = 0; v < NVAR; v++) ffv[v] = ff_variable[v];
    const double3* s[4] = { &ff_flux_contribution_momentum_x, &ff_flux_contribution_momentum_y,
                            &ff_flux_contribution_momentum_z, &ff_flux_contribution_density_energy };
    for (int k = 0; k < 4; k++) { ffc[3*k] = s[k]->x; ffc[3*k+1] = s[k]->y; ffc[3*k+2] = s[k]->z; }
}

// ---- single-function wrappers on caller-owned arrays (kernel-level known-answer tests) ----
void refs_compute_step_factor(long nel, const double* variables, const double* volumes, double* sf, int legacy) {
    if (legacy) compute_step_factor_legacy(nel, variables, volumes, sf);
    else        compute_step_factor(nel, variables, volumes, sf);
}
void refs_compute_flux_edge(long first, long n, const void* edges, const double* variables, double* fluxes) {
    compute_flux_edge(first, n, (const edge_neighbour*)edges, variables, fluxes);
}
void refs_compute_boundary_flux_edge(long first, long n, const void* edges, const double* variables, double* fluxes) {
    compute_boundary_flux_edge(first, n, (const edge_neighbour*)edges, variables, fluxes);
}
void refs_compute_wall_flux_edge(long first, long n, const void* edges, const double* variables, double* fluxes) {
    compute_wall_flux_edge(first, n, (const edge_neighbour*)edges, variables, fluxes);
}
void refs_indirect_rw(long first, long n, const void* edges, const double* variables, double* fluxes) {
    indirect_rw(first, n, (const edge_neighbour*)edges, variables, fluxes);
}
void refs_time_step(int j, long nel, const double* sf, double* fluxes, const double* old_variables, double* variables) {
    time_step(j, nel, sf, fluxes, old_variables, variables);
}
void refs_residual(long nel, const double* old_variables, const double* variables, double* residuals) {
    residual(nel, old_variables, variables, residuals);
}
double refs_calc_rms(long nel, const double* residuals) { return calc_rms(nel, residuals); }
void refs_mg_restrict(double* v1, double* v2, long nel2, long* mapping, long* up_scratch, long mgc) {
    mg_restrict(v1, v2, nel2, mapping, up_scratch, mgc);
}
void refs_prolong(void* edges, long nI, double* res1, double* res2, double* vars2, long nel2, long* mapping,
                  void* coords1, void* coords2) {
    prolong_residuals_interpolate_proper((edge_neighbour*)edges, nI, res1, res2, vars2, nel2, mapping,
                                         (double3*)coords1, (double3*)coords2);
}
void refs_adjust_ewt(const void* coords, long ne, void* edges) { adjust_ewt((const double3*)coords, ne, (edge_neighbour*)edges); }
void refs_dampen_ewt(long ne, void* edges, double f) { dampen_ewt(ne, (edge_neighbour*)edges, f); }

// ---- file readers (the reference's own parsers; pin our writers/loaders against them) ----
// read_grid (src/Base/io.cpp:14-199). Returned arrays are malloc'ed by the reference; release with refs_free.
int refs_read_grid(const char* path, long* hdr /*nel,ne,nI,nB,nW,iS,bS,wS*/, double** volumes, void** edges, void** coords) {
    edge_neighbour* e; double3* c;
    read_grid(path, &hdr[0], volumes, &hdr[1], &hdr[2], &hdr[3], &hdr[4], &hdr[5], &hdr[6], &hdr[7], &e, &c);
    *edges = e; *coords = c;
    return 0;
}
int refs_read_mg(const char* path, long** mg, long* mgc) { read_mg_connectivity(path, mg, mgc); return 0; }
void refs_free(void* p) { free(p); }
// read_input_dat (src/Base/io_enhanced.cpp:407-579): returns levels/mesh_variant via globals.
int refs_read_input_dat(const char* path, int* size, int* nlevels, int* variant, char* names_out, int cap) {
    std::string* layers = NULL; std::string* mgs = NULL;
    read_input_dat(path, size, &layers, &mgs);
    *nlevels = levels; *variant = mesh_variant;
    std::string all;
    for (int l = 0; l < levels; l++) { all += layers[l]; all += "\n"; }
    for (int l = 0; l < levels-1; l++) { all += mgs[l]; all += "\n"; }
    if ((int)all.size() + 1 > cap) return -1;
    memcpy(names_out, all.c_str(), all.size() + 1);
    return 0;
}

// ---- a multi-level session: arrays owned by the reference's alloc<>, loop sequencing as main() ----
ref_session* refs_create(int nlevels, int variant) {
    ref_session* s = new ref_session();
    levels = nlevels; mesh_variant = variant; level = 0;
    s->L.resize(nlevels);
    memset(&s->L[0], 0, sizeof(ref_level)*nlevels);
    s->up_scratch = NULL;
    return s;
}
// copies caller arrays (as read_grid would have produced them: edges AoS 40 B, coords double3, mg long)
void refs_set_mesh(ref_session* s, int l, long nel, const double* volumes, const void* coords,
                   long nI, long nB, long nW, const void* edges, const long* mg, long mgc) {
    ref_level& v = s->L[l];
    v.nel = nel; v.nI = nI; v.nB = nB; v.nW = nW; v.ne = nI+nB+nW; v.iS = 0; v.bS = nI; v.wS = nI+nB; v.mgc = mgc;
    v.volumes = alloc<double>(nel); memcpy(v.volumes, volumes, sizeof(double)*nel);
    v.coords = alloc<double3>(nel);
    if (coords) memcpy(v.coords, coords, sizeof(double3)*nel); else memset(v.coords, 0, sizeof(double3)*nel);
    v.edges = alloc<edge_neighbour>(v.ne); memcpy(v.edges, edges, sizeof(edge_neighbour)*v.ne);
    v.mg = NULL;
    if (mg) { v.mg = alloc<long>(mgc); memcpy(v.mg, mg, sizeof(long)*mgc); }
    alloc_state(v);
}
// -m duplication through the reference's own duplicate_mesh (src/Base/io_enhanced.cpp:89-201)
void refs_duplicate(ref_session* s, int m) {
    conf.mesh_duplicate_count = m;
    const int nl = (int)s->L.size();
    for (int i = 0; i < nl; i++) {
        ref_level& v = s->L[i];
        if (i < nl-1) duplicate_mesh(&v.nel, &v.volumes, &v.coords, &v.ne, &v.nI, &v.nB, &v.nW, &v.bS, &v.wS, &v.edges,
                                     s->L[i+1].nel, &v.mg, &v.mgc);
        else          duplicate_mesh(&v.nel, &v.volumes, &v.coords, &v.ne, &v.nI, &v.nB, &v.nW, &v.bS, &v.wS, &v.edges,
                                     0, NULL, NULL);
        free_state(v); alloc_state(v);
    }
}
// initialise exactly as main() does (src/euler3d_cpu_double.cpp:321-352)
void refs_prepare(ref_session* s) {
    initialize_far_field_conditions();
    const int nl = (int)s->L.size();
    s->up_scratch = alloc<long>(s->L[0].nel);
    for (int i = 0; i < nl; i++) initialize_variables(s->L[i].nel, s->L[i].variables);
    for (int i = 0; i < nl; i++) { zero_array(NVAR*s->L[i].nel, s->L[i].fluxes); zero_array(NVAR*s->L[i].nel, s->L[i].residuals); }
    double damp = 0.0;
    if (mesh_variant == MESH_M6_WING) damp = 5e-8;
    else if (mesh_variant == MESH_LA_CASCADE) damp = 1e-7;
    else if (mesh_variant == MESH_ROTOR_37) damp = 2e-7;
    if (damp != 0.0) for (int l = 0; l < nl; l++) {
        adjust_ewt(s->L[l].coords, s->L[l].ne, s->L[l].edges);
        dampen_ewt(s->L[l].ne, s->L[l].edges, damp);
    }
}
// field: 0 variables 1 old 2 residuals 3 fluxes 4 step_factors 5 volumes 6 edges 7 coords 8 mg
void* refs_ptr(ref_session* s, int l, int field) {
    ref_level& v = s->L[l];
    switch (field) {
        case 0: return v.variables; case 1: return v.old_variables; case 2: return v.residuals;
        case 3: return v.fluxes; case 4: return v.step_factors; case 5: return v.volumes;
        case 6: return v.edges; case 7: return v.coords; case 8: return v.mg;
    }
    return NULL;
}
void refs_dims(ref_session* s, int l, long* out /*nel,nI,nB,nW,mgc*/) {
    ref_level& v = s->L[l]; out[0]=v.nel; out[1]=v.nI; out[2]=v.nB; out[3]=v.nW; out[4]=v.mgc;
}

// One smoothing visit of level l (src/euler3d_cpu_double.cpp:383-512), timed per call like -DTIME would.
static void smooth(ref_session* s, int l, int probe) {
    ref_level& v = s->L[l]; level = l;
    copy<double>(v.old_variables, v.variables, v.nel*NVAR);
    double t0 = omp_get_wtime();
    if (mesh_variant == MESH_FVCORR) compute_step_factor_legacy(v.nel, v.variables, v.volumes, v.step_factors);
    else                             compute_step_factor(v.nel, v.variables, v.volumes, v.step_factors);
    s->t_step[l] += omp_get_wtime() - t0;
    for (int j = 0; j < RK; j++) {
        t0 = omp_get_wtime();
        compute_flux_edge(v.iS, v.nI, v.edges, v.variables, v.fluxes);
        s->t_flux[l] += omp_get_wtime() - t0;
        compute_boundary_flux_edge(v.bS, v.nB, v.edges, v.variables, v.fluxes);
        compute_wall_flux_edge(v.wS, v.nW, v.edges, v.variables, v.fluxes);
        t0 = omp_get_wtime();
        time_step(j, v.nel, v.step_factors, v.fluxes, v.old_variables, v.variables);
        s->t_time[l] += omp_get_wtime() - t0;
        check_for_invalid_variables(v.variables, v.nel);
        if (probe) {
            indirect_rw(v.iS, v.nI, v.edges, v.variables, v.fluxes);
            zero_fluxes(v.nel, v.fluxes);
        }
    }
    residual(v.nel, v.old_variables, v.variables, v.residuals);
}

// Runs `cycles` V-cycles. rms_all[c] = calc_rms(level 0) as printed by main(); rms_var[c*5+v] is the
// per-variable RMS sqrt(sum_i r_iv^2 / N) computed here from the reference's residuals[0] (the reference
// itself has no per-variable RMS).  Returns loop wall time.
double refs_run(ref_session* s, int cycles, int probe, double* rms_all, double* rms_var) {
    const int nl = (int)s->L.size();
    for (int l = 0; l < 8; l++) s->t_flux[l] = s->t_step[l] = s->t_time[l] = s->t_restrict[l] = s->t_prolong[l] = 0.0;
    double t_begin = omp_get_wtime();
    int lev = 0; int dir = MG_RESTRICT;
    for (int i = 0; i < cycles;) {
        smooth(s, lev, probe);
        if (lev == 0) {
            ref_level& v = s->L[0];
            if (rms_all) rms_all[i] = calc_rms(v.nel, v.residuals);
            if (rms_var) for (int k = 0; k < NVAR; k++) {
                double acc = 0.0;
                for (long n = 0; n < v.nel; n++) { double r = v.residuals[n*NVAR+k]; acc += r*r; }
                rms_var[i*NVAR+k] = sqrt(acc / double(v.nel));
            }
        }
        if (nl <= 1) { i++; continue; }
        if (dir == MG_RESTRICT) {
            lev++; level = lev;
            double t0 = omp_get_wtime();
            mg_restrict(s->L[lev-1].variables, s->L[lev].variables, s->L[lev].nel, s->L[lev-1].mg, s->up_scratch, s->L[lev-1].mgc);
            s->t_restrict[lev] += omp_get_wtime() - t0;
            if (lev == nl-1) dir = MG_PROLONG;
        } else {
            lev--; level = lev;
            double t0 = omp_get_wtime();
            prolong_residuals_interpolate_proper(s->L[lev].edges, s->L[lev].nI, s->L[lev+1].residuals, s->L[lev].residuals,
                                                 s->L[lev].variables, s->L[lev].nel, s->L[lev].mg, s->L[lev+1].coords, s->L[lev].coords);
            s->t_prolong[lev] += omp_get_wtime() - t0;
            if (lev == 0) { dir = MG_RESTRICT; i++; }
        }
    }
    s->t_total = omp_get_wtime() - t_begin;
    return s->t_total;
}
void refs_times(ref_session* s, double* out /*5*8*/) {
    for (int l = 0; l < 8; l++) { out[l]=s->t_flux[l]; out[8+l]=s->t_step[l]; out[16+l]=s->t_time[l]; out[24+l]=s->t_restrict[l]; out[32+l]=s->t_prolong[l]; }
}
void refs_destroy(ref_session* s) {
    for (size_t l = 0; l < s->L.size(); l++) {
        ref_level& v = s->L[l];
        if (!v.volumes) continue;
        dealloc(v.volumes); dealloc(v.edges); dealloc(v.coords); if (v.mg) dealloc(v.mg);
        free_state(v);
    }
    if (s->up_scratch) dealloc(s->up_scratch);
    delete s;
}

} // extern "C"
