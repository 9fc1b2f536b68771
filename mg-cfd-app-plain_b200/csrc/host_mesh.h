// host_mesh.h -- host-side multigrid mesh in the reference's in-memory layout (what read_grid /
// read_mg_connectivity leave behind; src/Base/io.cpp:14-199, src/Base/io_enhanced.cpp:629-650).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mgcfd {

// layout-compatible with the reference's `edge_neighbour` (src/Base/definitions.h:83): 40 bytes
struct EdgeNb { long a, b; double x, y, z; };
static_assert(sizeof(EdgeNb) == 40, "edge_neighbour layout");

struct HostLevel {
    long nel = 0, nI = 0, nB = 0, nW = 0;
    std::vector<double> volumes;   // nel
    std::vector<EdgeNb> edges;     // [internal | boundary(a=-1) | wall(a=-2)]
    std::vector<double> coords;    // 3*nel (xyz AoS) or empty
    std::vector<long> mg;          // fine->coarse map (size nel) or empty on the coarsest level
    std::string name;              // file name used by the writer
    // distributed runs (partition.h): nodes [0, n_owned) are owned by this rank, the rest are ghosts (never computed here);
    // gid = global node ids, used wherever the reference's ordering (ascending node / edge index) decides a summation order
    long n_owned = -1;             // -1: every node is owned
    std::vector<long> gid;
};

struct HostMesh {
    int mesh_variant = 2;
    int size = 1;                  // input.dat `size=`
    int copies = 1;                // mgcfd_mesh_duplicate: independent copies laid out copy-major on every level
    bool ewt_applied = false;
    std::vector<HostLevel> levels;
};

// one line of the text format: neighbour id (>=0 node, -1 boundary, -2 wall) + weight vector
struct Entry { long nbr; double w[3]; };

// Streaming node-centric description of one level (the content of a reference text mesh file).
struct NodeSource {
    virtual ~NodeSource() {}
    virtual long nel() const = 0;
    virtual double volume(long i) const = 0;
    virtual void coords(long i, double* xyz) const = 0;
    // the listing of node i (any degree): `deg` entries, valid until the next call on this source
    virtual const Entry* listing(long i, int& deg) const = 0;
};

// Applies read_grid's rules (io.cpp:84-181) to a node source: an edge for every entry with nbr < i,
// a = nbr, b = i, weight negated for internal edges (all edges for fvcorr), stored internal|boundary|wall.
void build_level_like_read_grid(const NodeSource& src, int mesh_variant, bool want_coords, HostLevel& out);

// adjust_ewt + dampen_ewt (src/Kernels/validation.cpp:28-75) by variant (euler3d_cpu_double.cpp:337-352)
void apply_ewt(int mesh_variant, const double* coords, long ne, EdgeNb* edges);

// text-format I/O (the interchange contract)
int write_level_text(const HostLevel& L, int mesh_variant, const std::string& path, bool with_coords);
int write_mg_text(const HostLevel& L, const std::string& path);
int write_input_dat(const HostMesh& m, const std::string& dir, const std::string& fname);
int read_level_text(const std::string& path, int mesh_variant, bool need_coords, HostLevel& out, std::string& err);
int read_mg_text(const std::string& path, std::vector<long>& mg, std::string& err);
int read_input_dat(const std::string& path, int& size, int& levels, int& variant, std::vector<std::string>& layers,
                   std::vector<std::string>& mgfiles, std::string& err);
int load_mesh(const std::string& input_dat, const std::string& dir, HostMesh& out, std::string& err);
// the reference's .bin cache layout (io_enhanced.cpp:384-400), reader bug (:341) not reproduced
int write_level_bin(const HostLevel& L, const std::string& path);
int read_level_bin(const std::string& path, HostLevel& out, std::string& err);

// synthetic generators
struct MeshSpec {
    int kind = 0;            // 0 hex-stencil box (6 nbrs), 1 Kuhn-tet box (14 nbrs), 2 cell-centred tets (4 faces, fvcorr-like)
    int levels = 1;
    long dims[8][3] = {};    // nodes per direction per level (kind 2: cubes per direction)
    double lengths[3] = {1, 1, 1};
    int mesh_variant = 2;
    int ordering = 0;        // 0 lexicographic, 1 seeded random permutation of node ids (Fisher-Yates)
    uint64_t seed = 12345;
    double tilt = 0.05;      // x-tilt of the z=0 wall normals on the patch 0.3 < x/Lx < 0.6
};
int generate_mesh(const MeshSpec& spec, HostMesh& out, std::string& err);
// rank-local generation for multi-GPU runs (partition.h: partition_sources): the part of the mesh `spec` describes that rank `rank`
// of `nranks` holds, without ever assembling the global edge list; apply_ewt != 0 applies adjust_ewt + dampen_ewt to the local edges
struct LocalMesh;
int generate_partition(const MeshSpec& spec, int nranks, int rank, bool apply_ewt_too, LocalMesh& out, std::string& err);

}  // namespace mgcfd
