// mesh_gen.cpp -- synthetic multigrid meshes "of the named shapes" (BASELINE.json configs), produced as the
// node-centric listing a reference text mesh file holds and then passed through the same rules read_grid applies
// (src/Base/io.cpp:84-181), so that what we upload is what the reference would have loaded from the file.
// The release meshes (fvcorr.domn.097K, Onera M6) are not available offline.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <memory>
#include <numeric>

#include "host_mesh.h"
#include "partition.h"

namespace mgcfd {

namespace {

const int MESH_FVCORR = 0;

struct SplitMix64 {
    uint64_t s;
    explicit SplitMix64(uint64_t seed) : s(seed) {}
    uint64_t next() {
        uint64_t z = (s += 0x9E3779B97F4A7C15ull);
        z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
        z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
        return z ^ (z >> 31);
    }
};

// Structured box with a 6-neighbour (hex dual) or 14-neighbour (Kuhn tetrahedra) vertex stencil.
struct BoxSource : NodeSource {
    long nx, ny, nz;
    double L[3], h[3];
    int kind, variant;
    double tilt;
    BoxSource(const long* d, const double* len, int kind_, int variant_, double tilt_)
        : nx(d[0]), ny(d[1]), nz(d[2]), kind(kind_), variant(variant_), tilt(tilt_) {
        for (int k = 0; k < 3; k++) { L[k] = len[k]; h[k] = len[k] / double(d[k] - 1); }
    }
    long nel() const override { return nx * ny * nz; }
    inline void ijk(long i, long* q) const { q[0] = i % nx; q[1] = (i / nx) % ny; q[2] = i / (nx * ny); }
    inline long id(const long* q) const { return q[0] + nx * (q[1] + ny * q[2]); }
    inline double extent(int k, long q) const {
        const long n = (k == 0 ? nx : (k == 1 ? ny : nz));
        return (q == 0 || q == n - 1) ? 0.5 * h[k] : h[k];
    }
    double volume(long i) const override {
        long q[3]; ijk(i, q);
        return extent(0, q[0]) * extent(1, q[1]) * extent(2, q[2]);
    }
    void coords(long i, double* xyz) const override {
        long q[3]; ijk(i, q);
        const long n[3] = {nx, ny, nz};
        // same expression on every level so that coincident fine/coarse nodes compare equal bit for bit
        for (int k = 0; k < 3; k++) xyz[k] = L[k] * (double(q[k]) / double(n[k] - 1));
    }
    bool in_patch(long ix) const {
        const double xr = double(ix) / double(nx - 1);
        return xr > 0.3 && xr < 0.6;
    }
    mutable Entry buf[32];      // a node of these meshes lists at most 14 neighbours + 6 boundary faces
    const Entry* listing(long i, int& deg) const override { deg = fill(i, buf); return buf; }
    int fill(long i, Entry* out) const {
        long q[3]; ijk(i, q);
        const long n[3] = {nx, ny, nz};
        const double sgn_bnd = (variant == MESH_FVCORR) ? 1.0 : -1.0;  // non-fvcorr: boundary normals written inward
        int deg = 0;
        double closure[3] = {0, 0, 0};  // outward boundary normal = sum of the missing directions' normals
        if (kind == 0) {
            const double ext[3] = {extent(0, q[0]), extent(1, q[1]), extent(2, q[2])};
            const double area[3] = {ext[1] * ext[2], ext[0] * ext[2], ext[0] * ext[1]};
            for (int k = 0; k < 3; k++)
                for (int s = -1; s <= 1; s += 2) {
                    long p[3] = {q[0], q[1], q[2]};
                    p[k] += s;
                    if (p[k] < 0 || p[k] >= n[k]) { closure[k] += s * area[k]; continue; }
                    Entry& e = out[deg++];
                    e.nbr = id(p);
                    e.w[0] = e.w[1] = e.w[2] = 0.0;
                    e.w[k] = s * area[k];
                }
        } else {
            static const int D[7][3] = {{1,0,0},{0,1,0},{0,0,1},{1,1,0},{0,1,1},{1,0,1},{1,1,1}};
            static const double A[7] = {0.25, 0.25, 0.25, 0.1, 0.1, 0.1, 0.05};
            const double face[3] = {h[1] * h[2], h[0] * h[2], h[0] * h[1]};
            for (int d = 0; d < 7; d++)
                for (int s = -1; s <= 1; s += 2) {
                    long p[3] = {q[0] + s * D[d][0], q[1] + s * D[d][1], q[2] + s * D[d][2]};
                    double w[3];
                    for (int k = 0; k < 3; k++) w[k] = s * A[d] * D[d][k] * face[k];
                    bool inside = true;
                    for (int k = 0; k < 3; k++) if (p[k] < 0 || p[k] >= n[k]) inside = false;
                    if (!inside) { for (int k = 0; k < 3; k++) closure[k] += w[k]; continue; }
                    Entry& e = out[deg++];
                    e.nbr = id(p);
                    for (int k = 0; k < 3; k++) e.w[k] = w[k];
                }
        }
        // boundary faces this node lies on, in the order x-, x+, y-, y+, z-, z+
        int faces[6], nf = 0;
        for (int k = 0; k < 3; k++) {
            if (q[k] == 0) faces[nf++] = 2 * k;
            if (q[k] == n[k] - 1) faces[nf++] = 2 * k + 1;
        }
        if (nf == 0) return deg;
        double fvec[6][3];
        memset(fvec, 0, sizeof(fvec));
        for (int k = 0; k < 3; k++) {
            int target = -1;
            for (int f = 0; f < nf; f++) if (faces[f] / 2 == k) { target = f; break; }
            if (target < 0) target = 0;
            fvec[target][k] += closure[k];
        }
        for (int f = 0; f < nf; f++) {
            Entry& e = out[deg++];
            const bool zminus = (faces[f] == 4);
            e.nbr = zminus ? -1 : -2;  // z=0 face: "boundary" (slip wall physics), all others: "wall" (far field)
            for (int k = 0; k < 3; k++) e.w[k] = sgn_bnd * fvec[f][k];
            if (zminus && in_patch(q[0])) e.w[0] += tilt * std::fabs(fvec[f][2]);
        }
        return deg;
    }
};

// Cell-centred tetrahedra: every cube of a cx*cy*cz box split into 6 Kuhn tets; a cell has 4 faces,
// like the cell-centred fvcorr.domn.097K mesh of rodinia/cfd that the reference's `fvcorr` variant targets.
struct TetCellSource : NodeSource {
    long cx, cy, cz;
    double h[3], L[3];
    int variant;
    double tilt;
    static const int P[6][3];
    TetCellSource(const long* d, const double* len, int variant_, double tilt_) : cx(d[0]), cy(d[1]), cz(d[2]), variant(variant_), tilt(tilt_) {
        for (int k = 0; k < 3; k++) { L[k] = len[k]; h[k] = len[k] / double(d[k]); }
    }
    long nel() const override { return 6 * cx * cy * cz; }
    static int perm_index(int a, int b, int c) {
        for (int p = 0; p < 6; p++) if (P[p][0] == a && P[p][1] == b && P[p][2] == c) return p;
        return -1;
    }
    void verts(long i, double v[4][3]) const {
        const long cube = i / 6; const int p = int(i % 6);
        const long q[3] = {cube % cx, (cube / cx) % cy, cube / (cx * cy)};
        for (int k = 0; k < 3; k++) v[0][k] = h[k] * double(q[k]);
        for (int s = 1; s <= 3; s++) {
            for (int k = 0; k < 3; k++) v[s][k] = v[s - 1][k];
            v[s][P[p][s - 1]] += h[P[p][s - 1]];
        }
    }
    double volume(long) const override { return h[0] * h[1] * h[2] / 6.0; }
    void coords(long i, double* xyz) const override {
        double v[4][3]; verts(i, v);
        for (int k = 0; k < 3; k++) xyz[k] = 0.25 * (v[0][k] + v[1][k] + v[2][k] + v[3][k]);
    }
    mutable Entry buf[8];
    const Entry* listing(long i, int& deg) const override { deg = fill(i, buf); return buf; }
    int fill(long i, Entry* out) const {
        const long cube = i / 6; const int p = int(i % 6);
        const long q[3] = {cube % cx, (cube / cx) % cy, cube / (cx * cy)};
        const long n[3] = {cx, cy, cz};
        const int p1 = P[p][0], p2 = P[p][1], p3 = P[p][2];
        double v[4][3]; verts(i, v);
        // face f is opposite vertex opp[f]: F0 (x_p1 = 1) opp v0, F1 (x_p3 = 0) opp v3, F2 (x_p1 = x_p2) opp v1, F3 (x_p2 = x_p3) opp v2
        static const int opp[4] = {0, 3, 1, 2};
        static const int fv[4][3] = {{1, 2, 3}, {0, 1, 2}, {0, 2, 3}, {0, 1, 3}};
        long nbr[4];
        {
            long c[3] = {q[0], q[1], q[2]};
            c[p1] += 1;
            nbr[0] = (c[p1] >= n[p1]) ? -2 : 6 * (c[0] + cx * (c[1] + cy * c[2])) + perm_index(p2, p3, p1);
            long d[3] = {q[0], q[1], q[2]};
            d[p3] -= 1;
            if (d[p3] < 0) nbr[1] = (p3 == 2) ? -1 : -2;   // z = 0 face is the "boundary" (-1); the rest far field (-2)
            else nbr[1] = 6 * (d[0] + cx * (d[1] + cy * d[2])) + perm_index(p3, p1, p2);
            nbr[2] = 6 * cube + perm_index(p2, p1, p3);
            nbr[3] = 6 * cube + perm_index(p1, p3, p2);
        }
        const bool inward_bnd = (variant != MESH_FVCORR);
        for (int f = 0; f < 4; f++) {
            const double* a = v[fv[f][0]]; const double* b = v[fv[f][1]]; const double* c = v[fv[f][2]];
            const double u[3] = {b[0] - a[0], b[1] - a[1], b[2] - a[2]};
            const double w[3] = {c[0] - a[0], c[1] - a[1], c[2] - a[2]};
            double nrm[3] = {0.5 * (u[1] * w[2] - u[2] * w[1]), 0.5 * (u[2] * w[0] - u[0] * w[2]), 0.5 * (u[0] * w[1] - u[1] * w[0])};
            const double* o = v[opp[f]];
            const double dot = nrm[0] * (o[0] - a[0]) + nrm[1] * (o[1] - a[1]) + nrm[2] * (o[2] - a[2]);
            if (dot > 0) for (int k = 0; k < 3; k++) nrm[k] = -nrm[k];   // outward = away from the opposite vertex
            Entry& e = out[f];
            e.nbr = nbr[f];
            const double s = (nbr[f] < 0 && inward_bnd) ? -1.0 : 1.0;
            for (int k = 0; k < 3; k++) e.w[k] = s * nrm[k];
            if (nbr[f] == -1) {
                const double xr = (0.25 * (v[0][0] + v[1][0] + v[2][0] + v[3][0])) / L[0];
                if (xr > 0.3 && xr < 0.6) e.w[0] += tilt * std::fabs(nrm[2]);
            }
        }
        return 4;
    }
};
const int TetCellSource::P[6][3] = {{0,1,2},{0,2,1},{1,0,2},{1,2,0},{2,0,1},{2,1,0}};

// A node source seen through a renumbering (new id = new_of_old[old id]).
struct PermutedSource : NodeSource {
    const NodeSource& base;
    const std::vector<long>& new_of_old;
    std::vector<long> old_of_new;
    PermutedSource(const NodeSource& b, const std::vector<long>& p) : base(b), new_of_old(p), old_of_new(p.size()) {
        for (size_t i = 0; i < p.size(); i++) old_of_new[p[i]] = long(i);
    }
    long nel() const override { return base.nel(); }
    double volume(long i) const override { return base.volume(old_of_new[i]); }
    void coords(long i, double* xyz) const override { base.coords(old_of_new[i], xyz); }
    mutable std::vector<Entry> buf;
    const Entry* listing(long i, int& deg) const override {
        const Entry* src = base.listing(old_of_new[i], deg);
        buf.assign(src, src + deg);
        for (int k = 0; k < deg; k++) if (buf[k].nbr >= 0) buf[k].nbr = new_of_old[buf[k].nbr];
        return buf.data();
    }
};

std::vector<long> random_permutation(long n, uint64_t seed) {
    std::vector<long> p(n);
    std::iota(p.begin(), p.end(), 0L);
    SplitMix64 rng(seed);
    for (long i = n - 1; i > 0; i--) {
        long j = long(rng.next() % uint64_t(i + 1));
        std::swap(p[i], p[j]);
    }
    return p;
}

// nearest coarse index along one direction, ties -> lower index; exact integer arithmetic
inline long nearest_1d(long qf, long nf, long nc) {
    const long t = qf * (nc - 1);
    long k = t / (nf - 1);
    const long rem = t % (nf - 1);
    if (2 * rem > (nf - 1)) k++;
    return k;
}

}  // namespace

void build_level_like_read_grid(const NodeSource& src, int mesh_variant, bool want_coords, HostLevel& out) {
    const long n = src.nel();
    out.nel = n;
    out.volumes.resize(n);
    if (want_coords) out.coords.resize(3 * n); else out.coords.clear();
    std::vector<EdgeNb> bnd, wall;
    out.edges.clear();
    for (long i = 0; i < n; i++) {
        out.volumes[i] = src.volume(i);
        if (want_coords) src.coords(i, &out.coords[3 * i]);
        int deg = 0;
        const Entry* ent = src.listing(i, deg);
        for (int j = 0; j < deg; j++) {
            const long i2 = ent[j].nbr;
            if (i2 >= i) continue;
            EdgeNb e;
            e.a = i2; e.b = i; e.x = ent[j].w[0]; e.y = ent[j].w[1]; e.z = ent[j].w[2];
            if (mesh_variant == MESH_FVCORR || i2 >= 0) { e.x *= -1; e.y *= -1; e.z *= -1; }
            if (i2 == -1) bnd.push_back(e);
            else if (i2 == -2) wall.push_back(e);
            else out.edges.push_back(e);
        }
    }
    out.nI = long(out.edges.size()); out.nB = long(bnd.size()); out.nW = long(wall.size());
    out.edges.insert(out.edges.end(), bnd.begin(), bnd.end());
    out.edges.insert(out.edges.end(), wall.begin(), wall.end());
}

void apply_ewt(int mesh_variant, const double* coords, long ne, EdgeNb* edges) {
    double damp = 0.0;
    if (mesh_variant == 2) damp = 5e-8; else if (mesh_variant == 3) damp = 1e-7; else if (mesh_variant == 4) damp = 2e-7;
    if (damp == 0.0) return;
    for (long i = 0; i < ne; i++) {
        EdgeNb& e = edges[i];
        if (e.a >= 0 && e.b >= 0) {
            // same accumulation order as adjust_ewt (validation.cpp:41-48).  The reference's own build (gcc, default
            // -ffp-contract=fast, -march=native on any FMA host; Makefile:95-108) contracts `dist += d*d` into an fma:
            // spelled out here so that the weights every kernel consumes are bit-identical to the reference's
            double dist, d;
            d = coords[3 * e.b + 0] - coords[3 * e.a + 0]; dist = d * d;
            d = coords[3 * e.b + 1] - coords[3 * e.a + 1]; dist = std::fma(d, d, dist);
            d = coords[3 * e.b + 2] - coords[3 * e.a + 2]; dist = std::fma(d, d, dist);
            dist = std::sqrt(dist);
            e.x /= dist; e.y /= dist; e.z /= dist;
        }
        e.x *= damp; e.y *= damp; e.z *= damp;
    }
}

int generate_mesh(const MeshSpec& spec, HostMesh& out, std::string& err) {
    if (spec.levels < 1 || spec.levels > 8) { err = "levels must be in 1..8"; return 2; }
    if (spec.kind == 2 && spec.levels != 1) { err = "cell-centred tet meshes are single-level"; return 2; }
    out = HostMesh();
    out.mesh_variant = spec.mesh_variant;
    out.levels.resize(spec.levels);
    std::vector<std::vector<long>> perms(spec.levels);
    for (int l = 0; l < spec.levels; l++) {
        const long* d = spec.dims[l];
        if (spec.kind != 2 && (d[0] < 2 || d[1] < 2 || d[2] < 2)) { err = "box dims must be >= 2"; return 2; }
        if (spec.kind == 2 && (d[0] < 1 || d[1] < 1 || d[2] < 1)) { err = "cube dims must be >= 1"; return 2; }
        std::unique_ptr<NodeSource> src;
        if (spec.kind == 2) src.reset(new TetCellSource(d, spec.lengths, spec.mesh_variant, spec.tilt));
        else src.reset(new BoxSource(d, spec.lengths, spec.kind, spec.mesh_variant, spec.tilt));
        HostLevel& L = out.levels[l];
        const bool want_coords = true;
        if (spec.ordering == 1) {
            perms[l] = random_permutation(src->nel(), spec.seed + uint64_t(l));
            PermutedSource ps(*src, perms[l]);
            build_level_like_read_grid(ps, spec.mesh_variant, want_coords, L);
        } else {
            build_level_like_read_grid(*src, spec.mesh_variant, want_coords, L);
        }
        L.name = "synth.L" + std::to_string(l) + ".dat";
    }
    // fine -> coarse maps: nearest coarse node (ties to the lower index), so coincident nodes map to each other
    for (int l = 0; l + 1 < spec.levels; l++) {
        const long* df = spec.dims[l]; const long* dc = spec.dims[l + 1];
        HostLevel& L = out.levels[l];
        L.mg.resize(L.nel);
        for (long i = 0; i < L.nel; i++) {
            const long q[3] = {i % df[0], (i / df[0]) % df[1], i / (df[0] * df[1])};
            const long c = nearest_1d(q[0], df[0], dc[0]) + dc[0] * (nearest_1d(q[1], df[1], dc[1]) + dc[1] * nearest_1d(q[2], df[2], dc[2]));
            const long fi = perms[l].empty() ? i : perms[l][i];
            const long ci = perms[l + 1].empty() ? c : perms[l + 1][c];
            L.mg[fi] = ci;
        }
    }
    long total = 0;
    for (auto& L : out.levels) total += L.nel;
    out.size = int(std::min<long>(total, 2000000000L));
    return 0;
}

int generate_partition(const MeshSpec& spec, int nranks, int rank, bool apply_ewt_too, LocalMesh& out, std::string& err) {
    if (spec.levels < 1 || spec.levels > 8) { err = "levels must be in 1..8"; return 2; }
    if (spec.kind == 2 && spec.levels != 1) { err = "cell-centred tet meshes are single-level"; return 2; }
    if (spec.ordering != 0) { err = "rank-local generation supports the lexicographic node numbering only"; return 2; }
    std::vector<std::unique_ptr<NodeSource>> src(spec.levels);
    std::vector<LevelSource> lv(spec.levels);
    for (int l = 0; l < spec.levels; l++) {
        const long* d = spec.dims[l];
        if (spec.kind != 2 && (d[0] < 2 || d[1] < 2 || d[2] < 2)) { err = "box dims must be >= 2"; return 2; }
        if (spec.kind == 2 && (d[0] < 1 || d[1] < 1 || d[2] < 1)) { err = "cube dims must be >= 1"; return 2; }
        if (spec.kind == 2) src[l].reset(new TetCellSource(d, spec.lengths, spec.mesh_variant, spec.tilt));
        else src[l].reset(new BoxSource(d, spec.lengths, spec.kind, spec.mesh_variant, spec.tilt));
        lv[l].src = src[l].get();
        lv[l].name = "synth.L" + std::to_string(l) + ".dat";
    }
    for (int l = 0; l + 1 < spec.levels; l++) {        // the same nearest-coarse-node map as generate_mesh
        const long* df = spec.dims[l]; const long* dc = spec.dims[l + 1];
        const long nf = src[l]->nel();
        lv[l].mg.resize(nf);
        for (long i = 0; i < nf; i++) {
            const long q[3] = {i % df[0], (i / df[0]) % df[1], i / (df[0] * df[1])};
            lv[l].mg[i] = int(nearest_1d(q[0], df[0], dc[0]) + dc[0] * (nearest_1d(q[1], df[1], dc[1]) + dc[1] * nearest_1d(q[2], df[2], dc[2])));
        }
    }
    try { partition_sources(lv, spec.mesh_variant, nranks, rank, out); }
    catch (const std::exception& ex) { err = ex.what(); return 2; }
    if (apply_ewt_too)
        for (auto& LL : out.levels) apply_ewt(spec.mesh_variant, LL.mesh.coords.data(), LL.mesh.nI + LL.mesh.nB + LL.mesh.nW, LL.mesh.edges.data());
    return 0;
}

}  // namespace mgcfd
