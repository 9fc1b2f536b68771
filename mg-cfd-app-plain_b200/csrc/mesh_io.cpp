// mesh_io.cpp -- readers/writers for the reference's mesh interchange formats:
//   text mesh     `nel nedges` + per node `volume degree` + degree x (`nbr wx wy wz`)   src/Base/io.cpp:56-137
//   <mesh>.coords `x y z` per node (needed iff levels > 1)                               src/Base/io.cpp:49-54,77-81
//   MG map        `mgc` + mgc fine->coarse indices                                       src/Base/io_enhanced.cpp:629-650
//   input.dat     size= / num_levels= / mesh_name= / [levels] / [mg_mapping]             src/Base/io_enhanced.cpp:407-579
//   .bin cache    8 long header, volumes, edges, coords, mg_size, mg                     src/Base/io_enhanced.cpp:384-400
// Own implementation: files are memory-mapped and parsed in one pass with std::from_chars (correctly rounded, ~6x the
// throughput of the strtod / operator>> the reference uses), written with std::to_chars in the shortest form that reads back
// to the same double; the edge construction rules are shared with the generators through build_level_like_read_grid.
#include <fcntl.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <cerrno>
#include <charconv>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "host_mesh.h"

namespace mgcfd {

namespace {

// read-only mapping of a whole file (page-cache backed: a 48 GB text mesh costs no second copy in memory)
struct MappedFile {
    const char* data = nullptr;
    size_t size = 0;
    bool ok = false;
    explicit MappedFile(const std::string& path) {
        const int fd = open(path.c_str(), O_RDONLY);
        if (fd < 0) return;
        struct stat st;
        if (fstat(fd, &st) == 0 && S_ISREG(st.st_mode)) {
            size = size_t(st.st_size);
            if (size == 0) { ok = true; data = ""; }
            else {
                void* m = mmap(nullptr, size, PROT_READ, MAP_PRIVATE, fd, 0);
                if (m != MAP_FAILED) { data = static_cast<const char*>(m); ok = true; madvise(m, size, MADV_SEQUENTIAL); }
            }
        }
        close(fd);
    }
    ~MappedFile() { if (ok && size) munmap(const_cast<char*>(data), size); }
    MappedFile(const MappedFile&) = delete;
    MappedFile& operator=(const MappedFile&) = delete;
    const char* begin() const { return data; }
    const char* end() const { return data + size; }
};

// whitespace-separated numbers, as operator>> reads them (io.cpp:56-137); a malformed or missing token is an error here
struct Cursor {
    const char* p;
    const char* e;
    const char* what;
    void skip() { while (p < e && static_cast<unsigned char>(*p) <= ' ') p++; }
    [[noreturn]] void fail() const { throw std::runtime_error(std::string("Corruption detected in '") + what + "': malformed or missing number"); }
    long integer() {
        skip();
        if (p < e && *p == '+') p++;
        long v = 0;
        const auto r = std::from_chars(p, e, v);
        if (r.ec != std::errc()) fail();
        p = r.ptr;
        return v;
    }
    double real() {
        skip();
        if (p < e && *p == '+') p++;
        double v = 0;
        const auto r = std::from_chars(p, e, v);
        if (r.ec == std::errc::result_out_of_range) {           // denormal / overflow: let strtod decide, as the reference's stream would
            std::string tok(p, size_t(std::min<long>(64, e - p)));
            char* endp = nullptr;
            v = strtod(tok.c_str(), &endp);
            p += endp - tok.c_str();
            return v;
        }
        if (r.ec != std::errc()) fail();
        p = r.ptr;
        return v;
    }
};

// buffered writer of numbers in the shortest form that parses back to the same value
struct NumberWriter {
    FILE* f;
    std::vector<char> buf;
    size_t n = 0;
    explicit NumberWriter(FILE* f_) : f(f_), buf(1 << 20) {}
    ~NumberWriter() { flush(); }
    void flush() { if (n) fwrite(buf.data(), 1, n, f); n = 0; }
    void room() { if (n + 64 > buf.size()) flush(); }
    void put(char c) { room(); buf[n++] = c; }
    void put(long v) { room(); n = size_t(std::to_chars(buf.data() + n, buf.data() + buf.size(), v).ptr - buf.data()); }
    void put(double v) { room(); n = size_t(std::to_chars(buf.data() + n, buf.data() + buf.size(), v).ptr - buf.data()); }
};

std::string trim(const std::string& s) {
    size_t a = s.find_first_not_of(" \t\r\n");
    if (a == std::string::npos) return "";
    size_t b = s.find_last_not_of(" \t\r\n");
    return s.substr(a, b - a + 1);
}

const char* variant_name(int v) {
    switch (v) { case 0: return "fvcorr"; case 2: return "m6wing"; case 3: return "la_cascade"; case 4: return "rotor37"; }
    return "m6wing";
}

// node source over a mapped text file; parses lazily and sequentially (listing(i) must be called with i ascending)
struct TextSource : NodeSource {
    mutable Cursor m, c;    // mesh and coords cursors
    bool have_coords;
    long n = 0, ne_claimed = 0;
    mutable double vol = 0;
    mutable long cur = -1;
    mutable std::vector<Entry> ent;   // the current node's listing, whatever its degree
    mutable int deg = 0;
    mutable double xyz[3] = {0, 0, 0};
    TextSource(const MappedFile& mesh, const MappedFile* coords, const char* mesh_name, const char* coords_name)
        : m{mesh.begin(), mesh.end(), mesh_name}, c{coords ? coords->begin() : nullptr, coords ? coords->end() : nullptr, coords_name}, have_coords(coords != nullptr) {
        n = m.integer();
        ne_claimed = m.integer();
        if (n < 0 || ne_claimed < 0) m.fail();
    }
    void advance(long i) const {
        while (cur < i) {
            vol = m.real();
            const long d = m.integer();
            if (d < 0 || d > 1000000) m.fail();
            deg = int(d);
            ent.resize(deg);
            for (int j = 0; j < deg; j++) {
                Entry& e = ent[j];
                e.nbr = m.integer();
                e.w[0] = m.real(); e.w[1] = m.real(); e.w[2] = m.real();
            }
            if (have_coords) { xyz[0] = c.real(); xyz[1] = c.real(); xyz[2] = c.real(); }
            cur++;
        }
    }
    long nel() const override { return n; }
    double volume(long i) const override { advance(i); return vol; }
    void coords(long i, double* o) const override { advance(i); o[0] = xyz[0]; o[1] = xyz[1]; o[2] = xyz[2]; }
    const Entry* listing(long i, int& d) const override { advance(i); d = deg; return ent.data(); }
};

}  // namespace

int write_level_text(const HostLevel& L, int mesh_variant, const std::string& path, bool with_coords) {
    const long n = L.nel;
    const long ne = L.nI + L.nB + L.nW;
    // CSR of edges incident to each node: lower entries (b == i) keep edge order; upper entries (a == i) follow.
    std::vector<long> cnt(n + 1, 0);
    for (long e = 0; e < ne; e++) {
        cnt[L.edges[e].b + 1]++;
        if (L.edges[e].a >= 0) cnt[L.edges[e].a + 1]++;
    }
    for (long i = 0; i < n; i++) cnt[i + 1] += cnt[i];
    std::vector<long> pos(cnt.begin(), cnt.end() - 1), slot(cnt[n]);
    // order per node: internal lower, boundary, wall, then upper -- fill in that class order
    for (long e = 0; e < ne; e++) slot[pos[L.edges[e].b]++] = e;                              // b-side, edge order: internal, bnd, wall
    for (long e = 0; e < L.nI; e++) slot[pos[L.edges[e].a]++] = -(e + 1);                      // a-side (upper neighbour)
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return 5;
    const bool fv = (mesh_variant == 0);
    {
        NumberWriter w(f);
        auto entry = [&w](long nbr, double x, double y, double z) { w.put(nbr); w.put(' '); w.put(x); w.put(' '); w.put(y); w.put(' '); w.put(z); w.put('\n'); };
        w.put(n); w.put(' '); w.put(ne); w.put('\n');
        for (long i = 0; i < n; i++) {
            w.put(L.volumes[i]); w.put(' '); w.put(cnt[i + 1] - cnt[i]); w.put('\n');
            for (long k = cnt[i]; k < cnt[i + 1]; k++) {
                const long s = slot[k];
                if (s >= 0) {
                    const EdgeNb& e = L.edges[s];
                    const double sg = (fv || e.a >= 0) ? -1.0 : 1.0;   // undo the loader's flip (io.cpp:111-133)
                    entry(e.a, sg * e.x, sg * e.y, sg * e.z);
                } else {
                    const EdgeNb& e = L.edges[-s - 1];                // listed from a: ignored by the loader (nbr > i)
                    entry(e.b, e.x, e.y, e.z);
                }
            }
        }
    }
    fclose(f);
    if (with_coords && !L.coords.empty()) {
        FILE* c = fopen((path + ".coords").c_str(), "w");
        if (!c) return 5;
        {
            NumberWriter w(c);
            for (long i = 0; i < n; i++) { w.put(L.coords[3 * i]); w.put(' '); w.put(L.coords[3 * i + 1]); w.put(' '); w.put(L.coords[3 * i + 2]); w.put('\n'); }
        }
        fclose(c);
    }
    return 0;
}

int write_mg_text(const HostLevel& L, const std::string& path) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) return 5;
    {
        NumberWriter w(f);
        w.put(long(L.mg.size())); w.put('\n');
        for (long v : L.mg) { w.put(v); w.put('\n'); }
    }
    fclose(f);
    return 0;
}

int write_input_dat(const HostMesh& m, const std::string& dir, const std::string& fname) {
    FILE* f = fopen((dir + "/" + fname).c_str(), "w");
    if (!f) return 5;
    fprintf(f, "# synthetic mesh written by mgcfd-b200 in the MG-CFD input format\n");
    fprintf(f, "size = %d\nnum_levels = %d\nmesh_name = %s\n", m.size, int(m.levels.size()), variant_name(m.mesh_variant));
    fprintf(f, "[levels]\n");
    for (size_t l = 0; l < m.levels.size(); l++) fprintf(f, "%zu = %s\n", l, m.levels[l].name.c_str());
    if (m.levels.size() > 1) {
        fprintf(f, "[mg_mapping]\n");
        for (size_t l = 0; l + 1 < m.levels.size(); l++) fprintf(f, "%zu = %s.mg\n", l, m.levels[l].name.c_str());
    }
    fclose(f);
    return 0;
}

int read_level_text(const std::string& path, int mesh_variant, bool need_coords, HostLevel& out, std::string& err) {
    MappedFile mesh(path);
    if (!mesh.ok) { err = "could not open data file: '" + path + "'"; return 5; }
    const std::string cpath = path + ".coords";
    MappedFile coords(cpath);
    if (!coords.ok && need_coords) { err = "could not open coords file for: " + path; return 5; }
    try {
        TextSource src(mesh, coords.ok ? &coords : nullptr, path.c_str(), cpath.c_str());
        build_level_like_read_grid(src, mesh_variant, coords.ok, out);
        if (out.nI + out.nB + out.nW != src.ne_claimed)
            fprintf(stderr, "WARNING: Mesh claims to have %ld edges, actually has %ld\n", src.ne_claimed, out.nI + out.nB + out.nW);
    } catch (const std::exception& ex) { err = ex.what(); return 5; }
    return 0;
}

int read_mg_text(const std::string& path, std::vector<long>& mg, std::string& err) {
    MappedFile file(path);
    if (!file.ok) { err = "could not open mg file: '" + path + "'"; return 5; }
    try {
        Cursor c{file.begin(), file.end(), path.c_str()};
        const long n = c.integer();
        if (n < 0) c.fail();
        mg.resize(n);
        for (long i = 0; i < n; i++) mg[i] = c.integer();
    } catch (const std::exception& ex) { err = ex.what(); return 5; }
    return 0;
}

int read_input_dat(const std::string& path, int& size, int& levels, int& variant, std::vector<std::string>& layers,
                   std::vector<std::string>& mgfiles, std::string& err) {
    std::ifstream file(path.c_str());
    if (!file.is_open()) { err = "Error: Could not open input file '" + path + "'"; return 5; }
    bool have_size = false, have_levels = false, have_name = false, have_files = false;
    std::string line;
    levels = 0;
    auto read_block = [&](std::vector<std::string>& dst, int count, const char* what) -> bool {
        dst.assign(count, "");
        for (int i = 0; i < count; i++) {
            if (!std::getline(file, line)) { err = std::string("Error parsing ") + path + ": reached EOF before reading all " + what; return false; }
            size_t eq = line.find('=');
            if (eq == std::string::npos) { err = std::string("Error parsing '") + path + "': expected key-value pair in " + what; return false; }
            int idx = atoi(trim(line.substr(0, eq)).c_str());
            if (idx >= 0 && idx < count) dst[idx] = trim(line.substr(eq + 1));
        }
        return true;
    };
    while (std::getline(file, line)) {
        if (!line.empty() && line[0] == '#') continue;
        if (!line.empty() && line[0] == '[') {
            const std::string t = trim(line);
            if (t == "[levels]") {
                if (!have_levels) { err = "Error parsing " + path + ": Need to know number of levels before parsing level filenames"; return 5; }
                if (!read_block(layers, levels, "mesh filenames")) return 5;
                have_files = true;
            } else if (t == "[mg_mapping]") {
                if (!have_levels) { err = "Error parsing " + path + ": Need to know number of levels before parsing level filenames"; return 5; }
                if (!read_block(mgfiles, levels - 1, "MG filenames")) return 5;
            }
            continue;
        }
        size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        const std::string key = trim(line.substr(0, eq)), value = trim(line.substr(eq + 1));
        if (value.empty()) continue;
        if (key == "size") { size = atoi(value.c_str()); have_size = true; }
        else if (key == "num_levels") {
            levels = atoi(value.c_str()); have_levels = true;
            if (levels < 1 || levels > 8) { err = "Error parsing '" + path + "': num_levels must be between 1 and 8"; return 5; }
        }
        else if (key == "mesh_name") {
            if (value == "la_cascade") variant = 3;
            else if (value == "rotor37") variant = 4;
            else if (value == "fvcorr") variant = 0;
            else if (value == "m6wing") variant = 2;
            else { err = "Error parsing " + path + ": Unknown mesh_name '" + value + "'"; return 5; }
            have_name = true;
        }
    }
    if (!have_size) { err = "Error parsing '" + path + "': size not present"; return 5; }
    if (!have_levels) { err = "Error parsing '" + path + "': number of levels not present"; return 5; }
    if (levels < 1 || levels > 8) { err = "Error parsing '" + path + "': num_levels must be between 1 and 8"; return 5; }
    if (!have_name) { err = "Error parsing '" + path + "': mesh name not present"; return 5; }
    if (!have_files) { err = "Error parsing '" + path + "': mesh filenames not present"; return 5; }
    if (int(mgfiles.size()) != levels - 1) mgfiles.assign(levels > 0 ? levels - 1 : 0, "");
    return 0;
}

int write_level_bin(const HostLevel& L, const std::string& path) {
    FILE* f = fopen(path.c_str(), "wb");
    if (!f) return 5;
    const long ne = L.nI + L.nB + L.nW;
    const long hdr[8] = {L.nel, ne, L.nI, L.nB, L.nW, 0, L.nI, L.nI + L.nB};
    fwrite(hdr, sizeof(long), 8, f);
    fwrite(L.volumes.data(), sizeof(double), L.nel, f);
    fwrite(L.edges.data(), sizeof(EdgeNb), ne, f);
    if (L.coords.empty()) { std::vector<double> z(3 * L.nel, 0.0); fwrite(z.data(), sizeof(double), 3 * L.nel, f); }
    else fwrite(L.coords.data(), sizeof(double), 3 * L.nel, f);
    const long mgs = long(L.mg.size());
    fwrite(&mgs, sizeof(long), 1, f);
    if (mgs) fwrite(L.mg.data(), sizeof(long), mgs, f);
    fclose(f);
    return 0;
}

int read_level_bin(const std::string& path, HostLevel& out, std::string& err) {
    FILE* f = fopen(path.c_str(), "rb");
    if (!f) { err = "'" + path + "' binary file cannot be read"; return 5; }
    long hdr[8];
    auto bad = [&]() { fclose(f); err = "Corruption detected in '" + path + "'"; return 5; };
    // every size in the header is checked against what the file can still hold BEFORE anything is allocated or indexed: a corrupt
    // cache file must come back as an error, never as a crash (load_mesh tries the .bin cache first)
    if (fseek(f, 0, SEEK_END) != 0) return bad();
    const long fsize = ftell(f);
    if (fsize < 0 || fseek(f, 0, SEEK_SET) != 0) return bad();
    if (fread(hdr, sizeof(long), 8, f) != 8) return bad();
    for (int k = 0; k < 8; k++) if (hdr[k] < 0) return bad();
    const long nel = hdr[0], ne = hdr[1], nI = hdr[2], nB = hdr[3], nW = hdr[4];
    const long cap = fsize / 8;                              // nothing in the file can count more items than this
    if (nel > cap || ne > cap || nI > ne || nB > ne || nW > ne || nI + nB + nW > ne) return bad();
    if (hdr[5] > ne - nI || hdr[6] > ne - nB || hdr[7] > ne - nW) return bad();      // the three ranges lie inside the edge array
    const __int128 need = (__int128)64 + (__int128)8 * nel + (__int128)40 * ne + (__int128)24 * nel + 8;
    if (need > (__int128)fsize) return bad();
    try {
        out.nel = nel; out.nI = nI; out.nB = nB; out.nW = nW;
        out.volumes.resize(nel);
        if (fread(out.volumes.data(), sizeof(double), nel, f) != size_t(nel)) return bad();
        std::vector<EdgeNb> all(ne);
        if (fread(all.data(), sizeof(EdgeNb), ne, f) != size_t(ne)) return bad();
        // the header allows gaps between the three ranges; we store them contiguously
        out.edges.clear();
        out.edges.insert(out.edges.end(), all.begin() + hdr[5], all.begin() + hdr[5] + nI);
        out.edges.insert(out.edges.end(), all.begin() + hdr[6], all.begin() + hdr[6] + nB);
        out.edges.insert(out.edges.end(), all.begin() + hdr[7], all.begin() + hdr[7] + nW);
        out.coords.resize(3 * nel);
        if (fread(out.coords.data(), sizeof(double), 3 * nel, f) != size_t(3 * nel)) return bad();
        long mgs = 0;
        if (fread(&mgs, sizeof(long), 1, f) != 1 || mgs < 0 || mgs > cap) return bad();
        out.mg.resize(mgs);   // (the reference casts the POINTER mg_size here, io_enhanced.cpp:341; not reproduced)
        if (mgs && fread(out.mg.data(), sizeof(long), mgs, f) != size_t(mgs)) return bad();
    } catch (const std::exception&) { return bad(); }
    fclose(f);
    return 0;
}

int load_mesh(const std::string& input_dat, const std::string& dir, HostMesh& out, std::string& err) {
    std::string path = input_dat;
    if (!dir.empty()) path = dir + "/" + input_dat;
    int size = 0, levels = 0, variant = 2;
    std::vector<std::string> layers, mgs;
    int rc = read_input_dat(path, size, levels, variant, layers, mgs, err);
    if (rc) return rc;
    out = HostMesh();
    out.mesh_variant = variant; out.size = size;
    out.levels.resize(levels);
    for (int l = 0; l < levels; l++) {
        std::string lp = dir.empty() ? layers[l] : dir + "/" + layers[l];
        HostLevel& L = out.levels[l];
        std::string e2;
        if (read_level_bin(lp + ".bin", L, e2) != 0) {
            rc = read_level_text(lp, variant, levels > 1, L, err);
            if (rc) return rc;
            if (l != levels - 1) {
                std::string mp = dir.empty() ? mgs[l] : dir + "/" + mgs[l];
                rc = read_mg_text(mp, L.mg, err);
                if (rc) return rc;
            }
        }
        L.name = layers[l];
    }
    // fine -> coarse maps must point into the next level (partitioning indexes per-node arrays with them before anything else looks)
    for (int l = 0; l + 1 < levels; l++) {
        const HostLevel& L = out.levels[l];
        if (long(L.mg.size()) != L.nel) { err = "multigrid map of level " + std::to_string(l) + " does not have one entry per node"; return 5; }
        for (long v : L.mg) if (v < 0 || v >= out.levels[l + 1].nel) { err = "multigrid map of level " + std::to_string(l) + " points outside level " + std::to_string(l + 1); return 5; }
    }
    return 0;
}

}  // namespace mgcfd
