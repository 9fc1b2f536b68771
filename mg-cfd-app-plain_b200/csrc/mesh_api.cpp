// mesh_api.cpp -- extern "C" surface of the host mesh utilities (include/mgcfd_mesh.h)
#include <algorithm>
#include <cstring>
#include <string>
#include <vector>

#include "../../include/mgcfd_mesh.h"
#include "host_mesh.h"

using namespace mgcfd;

struct mgcfd_mesh { HostMesh m; };

namespace { thread_local std::string g_mesh_err; }
extern "C" const char* mgcfd_mesh_last_error(void) { return g_mesh_err.c_str(); }

extern "C" {

int mgcfd_mesh_generate(int kind, int levels, const long* dims, const double lengths[3], int mesh_variant, int ordering,
                        unsigned long seed, double tilt, mgcfd_mesh** out) {
    if (!out || !dims || levels < 1 || levels > 8) { g_mesh_err = "bad arguments"; return MGCFD_ERR_ARG; }
    MeshSpec s;
    s.kind = kind; s.levels = levels; s.mesh_variant = mesh_variant; s.ordering = ordering; s.seed = seed; s.tilt = tilt;
    for (int l = 0; l < levels; l++) for (int k = 0; k < 3; k++) s.dims[l][k] = dims[3 * l + k];
    if (lengths) for (int k = 0; k < 3; k++) s.lengths[k] = lengths[k];
    mgcfd_mesh* m = new mgcfd_mesh();
    int rc = generate_mesh(s, m->m, g_mesh_err);
    if (rc) { delete m; *out = nullptr; return rc; }
    *out = m;
    return MGCFD_OK;
}

int mgcfd_mesh_load(const char* input_dat, const char* dir, mgcfd_mesh** out) {
    if (!out || !input_dat) { g_mesh_err = "bad arguments"; return MGCFD_ERR_ARG; }
    mgcfd_mesh* m = new mgcfd_mesh();
    int rc;
    try { rc = load_mesh(input_dat, dir ? dir : "", m->m, g_mesh_err); }      // no exception may cross the C ABI
    catch (const std::exception& ex) { g_mesh_err = std::string("cannot load the mesh: ") + ex.what(); rc = MGCFD_ERR_IO; }
    if (rc) { delete m; *out = nullptr; return rc; }
    *out = m;
    return MGCFD_OK;
}

int mgcfd_mesh_write(const mgcfd_mesh* m, const char* dir, const char* name, int binary) {
    if (!m || !dir || !name) { g_mesh_err = "bad arguments"; return MGCFD_ERR_ARG; }
    if (m->m.ewt_applied) { g_mesh_err = "edge weights already adjusted: write the mesh before mgcfd_mesh_apply_ewt"; return MGCFD_ERR_ARG; }
    const std::string d(dir);
    const int nl = int(m->m.levels.size());
    for (int l = 0; l < nl; l++) {
        const HostLevel& L = m->m.levels[l];
        int rc = write_level_text(L, m->m.mesh_variant, d + "/" + L.name, true);
        if (rc) { g_mesh_err = "cannot write " + L.name; return rc; }
        if (l + 1 < nl) { rc = write_mg_text(L, d + "/" + L.name + ".mg"); if (rc) { g_mesh_err = "cannot write mg map"; return rc; } }
        if (binary) { rc = write_level_bin(L, d + "/" + L.name + ".bin"); if (rc) { g_mesh_err = "cannot write .bin"; return rc; } }
    }
    int rc = write_input_dat(m->m, d, name);
    if (rc) g_mesh_err = "cannot write input.dat";
    return rc;
}

int mgcfd_mesh_levels(const mgcfd_mesh* m) { return m ? int(m->m.levels.size()) : -1; }
int mgcfd_mesh_variant(const mgcfd_mesh* m) { return m ? m->m.mesh_variant : -1; }

int mgcfd_mesh_dims(const mgcfd_mesh* m, int l, long out[5]) {
    if (!m || l < 0 || l >= int(m->m.levels.size())) { g_mesh_err = "bad level"; return MGCFD_ERR_ARG; }
    const HostLevel& L = m->m.levels[l];
    out[0] = L.nel; out[1] = L.nI; out[2] = L.nB; out[3] = L.nW; out[4] = long(L.mg.size());
    return MGCFD_OK;
}

const void* mgcfd_mesh_ptr(const mgcfd_mesh* m, int l, int what) {
    if (!m || l < 0 || l >= int(m->m.levels.size())) return nullptr;
    const HostLevel& L = m->m.levels[l];
    switch (what) {
        case 0: return L.volumes.data();
        case 1: return L.edges.data();
        case 2: return L.coords.empty() ? nullptr : L.coords.data();
        case 3: return L.mg.empty() ? nullptr : L.mg.data();
    }
    return nullptr;
}

int mgcfd_mesh_apply_ewt(mgcfd_mesh* m) {
    if (!m) { g_mesh_err = "null mesh"; return MGCFD_ERR_ARG; }
    if (m->m.ewt_applied) return MGCFD_OK;
    for (auto& L : m->m.levels) {
        if (m->m.mesh_variant != MGCFD_MESH_FVCORR && L.coords.empty()) { g_mesh_err = "coords needed for adjust_ewt"; return MGCFD_ERR_ARG; }
        apply_ewt(m->m.mesh_variant, L.coords.data(), L.nI + L.nB + L.nW, L.edges.data());
    }
    m->m.ewt_applied = true;
    return MGCFD_OK;
}

int mgcfd_mesh_upload(mgcfd_mesh* m, mgcfd_ctx* ctx) {
    if (!m || !ctx) { g_mesh_err = "null argument"; return MGCFD_ERR_ARG; }
    int rc = mgcfd_mesh_apply_ewt(m);
    if (rc) return rc;
    const int nl = int(m->m.levels.size());
    for (int l = 0; l < nl; l++) {
        const HostLevel& L = m->m.levels[l];
        rc = mgcfd_upload_level(ctx, l, L.nel, L.volumes.data(), L.coords.empty() ? nullptr : L.coords.data(), L.nI, L.nB, L.nW,
                                L.edges.data(), L.mg.empty() ? nullptr : L.mg.data(), long(L.mg.size()));
        if (rc) { g_mesh_err = mgcfd_last_error(); return rc; }
    }
    rc = mgcfd_finalize(ctx);
    if (rc) g_mesh_err = mgcfd_last_error();
    return rc;
}

// implemented in context.cu (they need the context's internals); the mesh travels as an opaque pointer
int mgcfd_upload_partition(mgcfd_ctx* c, int levels, int mesh_variant, const void* host_mesh_opaque);
int mgcfd_partition_plan(int levels, const void* host_mesh_opaque, int nranks, int rank, int level, long info[8], long* gid,
                         long* send_counts, long* recv_counts, long* send_gids);

int mgcfd_mesh_delivery_check_impl(const void* host_mesh_opaque, int nranks, int tile_nodes, long out[5]);
int mgcfd_mesh_delivery_check(mgcfd_mesh* m, int nranks, int tile_nodes, long out[5]) {
    if (!m) { g_mesh_err = "null argument"; return MGCFD_ERR_ARG; }
    int rc = mgcfd_mesh_delivery_check_impl(&m->m, nranks, tile_nodes, out);
    if (rc) g_mesh_err = mgcfd_last_error();
    return rc;
}

int mgcfd_mesh_upload_partition(mgcfd_mesh* m, mgcfd_ctx* ctx) {
    if (!m || !ctx) { g_mesh_err = "null argument"; return MGCFD_ERR_ARG; }
    int rc = mgcfd_mesh_apply_ewt(m);
    if (rc) return rc;
    rc = mgcfd_upload_partition(ctx, int(m->m.levels.size()), m->m.mesh_variant, &m->m);
    if (rc) g_mesh_err = mgcfd_last_error();
    return rc;
}
int mgcfd_mesh_partition_plan(mgcfd_mesh* m, int nranks, int rank, int level, long info[8], long* gid, long* send_counts,
                              long* recv_counts, long* send_gids) {
    if (!m) { g_mesh_err = "null argument"; return MGCFD_ERR_ARG; }
    int rc = mgcfd_partition_plan(int(m->m.levels.size()), &m->m, nranks, rank, level, info, gid, send_counts, recv_counts, send_gids);
    if (rc) g_mesh_err = mgcfd_last_error();
    return rc;
}

// -m / --mesh-duplicate-count: m independent copies of every level, laid out as duplicate_mesh does (io_enhanced.cpp:89-201):
// nodes copy-major, each edge class copy-major inside its own range, MG maps shifted per copy
int mgcfd_mesh_duplicate(mgcfd_mesh* m, int count) {
    if (!m || count < 1) { g_mesh_err = "bad arguments"; return MGCFD_ERR_ARG; }
    if (count == 1) return MGCFD_OK;
    const int nl = int(m->m.levels.size());
    std::vector<long> nel0(nl);
    for (int l = 0; l < nl; l++) nel0[l] = m->m.levels[l].nel;
    for (int l = 0; l < nl; l++) {
        HostLevel& L = m->m.levels[l];
        const long n = L.nel, nI = L.nI, nB = L.nB, nW = L.nW;
        HostLevel D;
        D.name = L.name; D.nel = n * count; D.nI = nI * count; D.nB = nB * count; D.nW = nW * count;
        D.volumes.resize(D.nel);
        if (!L.coords.empty()) D.coords.resize(3 * D.nel);
        for (int c = 0; c < count; c++) {
            std::copy(L.volumes.begin(), L.volumes.end(), D.volumes.begin() + c * n);
            if (!L.coords.empty()) std::copy(L.coords.begin(), L.coords.end(), D.coords.begin() + 3 * c * n);
        }
        D.edges.reserve(D.nI + D.nB + D.nW);
        const long start[3] = {0, nI, nI + nB}, cnt[3] = {nI, nB, nW};
        for (int cls = 0; cls < 3; cls++)
            for (int c = 0; c < count; c++)
                for (long e = 0; e < cnt[cls]; e++) {
                    EdgeNb ed = L.edges[start[cls] + e];
                    if (ed.a >= 0) ed.a += c * n;
                    ed.b += c * n;
                    D.edges.push_back(ed);
                }
        if (!L.mg.empty()) {
            D.mg.resize(D.nel);
            for (int c = 0; c < count; c++) for (long i = 0; i < n; i++) D.mg[c * n + i] = L.mg[i] + c * nel0[l + 1];
        }
        L = std::move(D);
    }
    m->m.size *= count;
    m->m.copies *= count;
    return MGCFD_OK;
}

void mgcfd_mesh_free(mgcfd_mesh* m) { delete m; }

}  // extern "C"
