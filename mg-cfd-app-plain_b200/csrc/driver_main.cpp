// driver_main.cpp -- euler3d_b200: the reference's euler3d driver (src/euler3d_cpu_double.cpp:69-809) on top of the C ABI.
// Same command line (src/Base/config.cpp:32-47, :281-305), same key = value config file (:81-217), same input.dat / mesh files,
// same progress lines ("MG cycle i / n (RMS = ...)", "Total runtime = ..."), same dump files (variables / step_factors / fluxes,
// %.17e, io.cpp:201-233, io_enhanced.cpp:652-817), same -v validation rule (validation.cpp:140-199) and the same
// Times.csv / LoopNumIters.csv schema (timer.cpp:106-195, loop_stats.cpp:83-171) with CUDA-event times; CPU-specific
// identification columns carry the GPU equivalents.  Host code only: every number comes from libmgcfd_b200.so.
//
// --gpus=N (not in the reference, which is a single process): the driver forks one process per GPU after the mesh has been read
// (the children inherit it copy-on-write, nothing touches CUDA before the fork); rank r drives device r through mgcfd_dist.h --
// NCCL id, peer-to-peer window handles, timings and the gathered fine-level arrays travel through one anonymous shared mapping;
// rank 0 prints, validates and writes the files, with one Times.csv / LoopNumIters.csv row per GPU where the reference writes
// one per thread.
#include <getopt.h>
#include <sched.h>
#include <signal.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <unistd.h>

#include <atomic>
#include <charconv>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/mgcfd_b200.h"
#include "../../include/mgcfd_mesh.h"
#include "../../include/mgcfd_dist.h"

namespace {

struct Config {                                   // config.h:27-47
    std::string config_filepath, input_file, input_file_directory, papi_config_file, output_file_prefix;
    int mesh_duplicate_count = 1, num_cycles = 25, omp_num_threads = 1;
    bool validate_result = false, output_variables = false, output_fluxes = false, output_step_factors = false, output_volumes = false;
    bool output_old_variables = false, output_edge_fluxes = false;      // accepted like the reference accepts them (never dumped there either)
    // not in the reference: device selection, number of GPUs and timing granularity
    int device = 0, flux_mode = -1, tile_nodes = 0, gpus = 1;
    bool kernel_times = true, write_solution = false;
} conf;

void print_help() {                               // config.cpp:281-305, plus the three device options
    fprintf(stderr, "MG-CFD (B200) instructions\n\n");
    fprintf(stderr, "Usage: euler3d_b200 [OPTIONS] \n");
    fprintf(stderr, "  -h, --help    Print help\n");
    fprintf(stderr, "  -i, --input-file=FILEPATH\n        multigrid input grid (.dat file)\n");
    fprintf(stderr, "  -c, --config-filepath=FILEPATH\n        config file\n");
    fprintf(stderr, "  -d, --input-directory=DIRPATH\n        directory path to input files\n");
    fprintf(stderr, "  -p, --papi_config_file=FILEPATH\n        accepted and ignored (CPU hardware counters; use ncu on the GPU)\n");
    fprintf(stderr, "  -o, --output-file-prefix=STRING\n        string to prepend to output filenames\n");
    fprintf(stderr, "  -m, --mesh-duplicate-count=INT\n        number of times to duplicate mesh\n");
    fprintf(stderr, "  -g, --num-cycles=INT\n        number of multigrid V-cycles to perform\n");
    fprintf(stderr, "  -v, --validate-result\n        check final state against pre-calculated solution\n");
    fprintf(stderr, "  --output-variables\n        write Euler equation variable values to file\n");
    fprintf(stderr, "  --output-fluxes\n        write flux accumulations to file\n");
    fprintf(stderr, "  --output-step-factors\n        write time-step factors to file\n");
    fprintf(stderr, "  --device=INT  --flux-mode=INT  --tile-nodes=INT  --no-kernel-times  --write-solution\n        B200 build only\n");
    fprintf(stderr, "  --gpus=INT\n        B200 build only: split the mesh over INT GPUs (devices 0..INT-1, one process each)\n");
}

std::string trim(const std::string& s) {
    const size_t a = s.find_first_not_of(" \t\r\n"), b = s.find_last_not_of(" \t\r\n");
    return a == std::string::npos ? "" : s.substr(a, b - a + 1);
}
bool truthy(const std::string& v) { return v == "Y" || v == "y" || v == "1" || v == "true" || v == "yes"; }

void read_config_file(const std::string& path) {  // config.cpp:159-217
    std::ifstream f(path.c_str());
    if (!f) { fprintf(stderr, "ERROR: Failed to open config: '%s'\n", path.c_str()); exit(EXIT_FAILURE); }
    std::string line;
    while (std::getline(f, line)) {
        line = trim(line);
        if (line.empty() || line[0] == '#') continue;
        const size_t eq = line.find('=');
        if (eq == std::string::npos) continue;
        const std::string key = trim(line.substr(0, eq)), val = trim(line.substr(eq + 1));
        if (key == "input_file") conf.input_file = val;
        else if (key == "input_file_directory") conf.input_file_directory = val;
        else if (key == "papi_config_file") conf.papi_config_file = val;
        else if (key == "output_file_prefix") conf.output_file_prefix = val;
        else if (key == "mesh_duplicate_count") conf.mesh_duplicate_count = atoi(val.c_str());
        else if (key == "cycles") conf.num_cycles = atoi(val.c_str());
        else if (key == "omp_num_threads") conf.omp_num_threads = atoi(val.c_str());
        else if (key == "validate_result") conf.validate_result = truthy(val);
        else if (key == "output_variables") conf.output_variables = truthy(val);
        else if (key == "output_fluxes") conf.output_fluxes = truthy(val);
        else if (key == "output_step_factors") conf.output_step_factors = truthy(val);
        else if (key == "output_volumes") conf.output_volumes = truthy(val);
        else if (key == "output_old_variables") conf.output_old_variables = truthy(val);
        else if (key == "output_edge_fluxes") conf.output_edge_fluxes = truthy(val);
        else if (key == "gpus") conf.gpus = atoi(val.c_str());
        else printf("WARNING: Unknown key '%s' encountered during parsing of config file.\n", key.c_str());
    }
    // a relative input_file_directory is relative to the config file (config.cpp:192-216)
    const size_t slash = path.rfind('/');
    const std::string config_dir = slash == std::string::npos ? "" : path.substr(0, slash);
    if (!config_dir.empty() && (conf.input_file_directory.empty() || conf.input_file_directory[0] != '/'))
        conf.input_file_directory = (conf.input_file_directory == "./" || conf.input_file_directory.empty()) ? config_dir : config_dir + "/" + conf.input_file_directory;
}

void parse_arguments(int argc, char** argv) {     // config.cpp:219-259
    static struct option long_opts[] = {
        {"help", no_argument, nullptr, 'h'}, {"config-filepath", required_argument, nullptr, 'c'}, {"input-file", required_argument, nullptr, 'i'},
        {"input-directory", required_argument, nullptr, 'd'}, {"papi_config_file", required_argument, nullptr, 'p'},
        {"output-file-prefix", required_argument, nullptr, 'o'}, {"mesh-duplicate-count", required_argument, nullptr, 'm'},
        {"num-cycles", required_argument, nullptr, 'g'}, {"validate-result", no_argument, nullptr, 'v'},
        {"output-variables", no_argument, nullptr, 1001}, {"output-fluxes", no_argument, nullptr, 1002}, {"output-step-factors", no_argument, nullptr, 1003},
        {"device", required_argument, nullptr, 1004}, {"flux-mode", required_argument, nullptr, 1005}, {"tile-nodes", required_argument, nullptr, 1006},
        {"no-kernel-times", no_argument, nullptr, 1007}, {"write-solution", no_argument, nullptr, 1008}, {"gpus", required_argument, nullptr, 1009},
        {nullptr, 0, nullptr, 0}};
    // first pass: the config file (command line options override it, as in the reference)
    for (int i = 1; i < argc; i++) {
        const std::string a(argv[i]);
        if ((a == "-c" || a == "--config-filepath") && i + 1 < argc) read_config_file(argv[i + 1]);
        else if (a.rfind("--config-filepath=", 0) == 0) read_config_file(a.substr(18));
    }
    int opt;
    optind = 1;
    while ((opt = getopt_long(argc, argv, "hi:c:d:p:o:m:g:v", long_opts, nullptr)) != -1) {
        switch (opt) {
            case 'h': print_help(); exit(EXIT_SUCCESS);
            case 'i': conf.input_file = optarg; break;
            case 'c': conf.config_filepath = optarg; break;
            case 'd': conf.input_file_directory = optarg; break;
            case 'p': conf.papi_config_file = optarg; break;
            case 'o': conf.output_file_prefix = optarg; break;
            case 'm': conf.mesh_duplicate_count = atoi(optarg); break;
            case 'g': conf.num_cycles = atoi(optarg); break;
            case 'v': conf.validate_result = true; break;
            case 1001: conf.output_variables = true; break;
            case 1002: conf.output_fluxes = true; break;
            case 1003: conf.output_step_factors = true; break;
            case 1004: conf.device = atoi(optarg); break;
            case 1005: conf.flux_mode = atoi(optarg); break;
            case 1006: conf.tile_nodes = atoi(optarg); break;
            case 1007: conf.kernel_times = false; break;
            case 1008: conf.write_solution = true; break;
            case 1009: conf.gpus = atoi(optarg); break;
            default: print_help(); exit(EXIT_FAILURE);
        }
    }
    if (conf.input_file.empty()) { fprintf(stderr, "ERROR: Input file not specified\n"); print_help(); exit(EXIT_FAILURE); }
    if (conf.mesh_duplicate_count < 1 || conf.num_cycles < 0) { fprintf(stderr, "ERROR: bad -m / -g value\n"); exit(EXIT_FAILURE); }
    if (conf.gpus < 1 || conf.gpus > 64) { fprintf(stderr, "ERROR: --gpus must be in 1..64\n"); exit(EXIT_FAILURE); }
}

std::string suffix(int level) {                   // io_enhanced.cpp:26-34
    std::ostringstream s;
    s << "size=" << conf.mesh_duplicate_count << "x.cycles=" << conf.num_cycles;
    if (level >= 0) s << ".level=" << level;
    return s.str();
}
std::string output_path(const std::string& name, int level) {     // io_enhanced.cpp:36-52
    std::string p = conf.output_file_prefix;
    if (!p.empty() && p[p.size() - 1] != '/') p += ".";
    return p + name + "." + suffix(level);
}
std::string solution_path(const std::string& name, int level) {   // io_enhanced.cpp:54-74
    std::string p = conf.input_file_directory;
    if (!p.empty() && p[p.size() - 1] != '/') p += "/";
    return p + "solution." + name + "." + suffix(level);
}
std::string csv_path(const std::string& name) {                   // timer.cpp:111-115
    std::string p = conf.output_file_prefix;
    if (!p.empty() && p[p.size() - 1] != '/') p += ".";
    return p + name;
}

// the reference's dump format (io.cpp:201-233, io_enhanced.cpp:652-817): one node per line, "%.17e" per value -- written with
// std::to_chars(scientific, 17), which produces the same bytes as printf (checked on 2e7 values incl. 0, -0, inf, nan, denormals)
// at a fifth of the time
void dump_rows(const std::string& path, const double* v, long n, int ncomp, bool announce) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) { fprintf(stderr, "ERROR: Failed to open file for writing: '%s'\n", path.c_str()); exit(EXIT_FAILURE); }
    if (announce) printf("Dumping variables[] to file: %s\n", path.c_str());
    std::vector<char> buf(1 << 20);
    size_t used = 0;
    for (long i = 0; i < n; i++) {
        if (used + 40 * (size_t)ncomp + 8 > buf.size()) { fwrite(buf.data(), 1, used, f); used = 0; }
        for (int k = 0; k < ncomp; k++) {
            used = size_t(std::to_chars(buf.data() + used, buf.data() + buf.size(), v[(size_t)ncomp * i + k], std::chars_format::scientific, 17).ptr - buf.data());
            buf[used++] = (k + 1 < ncomp) ? ' ' : '\n';
        }
    }
    fwrite(buf.data(), 1, used, f);
    fclose(f);
}

// identify_differences (validation.cpp:140-199)
void identify_differences(const double* test, const double* master, long n, int mesh_variant) {
    const double rel = 10.0e-9;
    const double abs_floor = mesh_variant == MGCFD_MESH_FVCORR ? 1.0e-15 : 3.0e-19;
    for (long i = 0; i < n; i++)
        for (int v = 0; v < 5; v++) {
            const long idx = 5 * i + v;
            double ok = master[idx] * rel;
            if (ok < 0.0) ok = -ok;
            if (ok < abs_floor) ok = abs_floor;
            double d = test[idx] - master[idx];
            if (d < 0.0) d = -d;
            if (d > ok || d != d) {
                printf("ERROR: Unacceptable error detected at (i=%ld, v=%d)\n", i, v);
                printf("       - incorrect value = %.23f\n", test[idx]);
                printf("       - correct value =   %.23f\n", master[idx]);
                printf("       - diff          =   %.23f\n", d);
                exit(EXIT_FAILURE);
            }
        }
}

#define CHECK(call)                                                                                         \
    do {                                                                                                    \
        int rc_ = (call);                                                                                   \
        if (rc_ != MGCFD_OK) { fprintf(stderr, "ERROR: %s failed (%d): %s\n", #call, rc_, mgcfd_last_error()); exit(EXIT_FAILURE); } \
    } while (0)

const char* variant_name(int v) {
    switch (v) { case MGCFD_MESH_LA_CASCADE: return "la_cascade"; case MGCFD_MESH_ROTOR_37: return "rotor37"; case MGCFD_MESH_FVCORR: return "fvcorr";
                 case MGCFD_MESH_M6_WING: return "m6wing"; }
    return "unknown";
}

// identification columns of prepare_csv_identification (io_enhanced.cpp:858-1016); CPU items carry the GPU equivalents
void csv_identification(std::ostringstream& header, std::ostringstream& line, int size, int variant, int flux_mode) {
    header << "Size,Mesh,MG cycles,Flux variant,Flux options,CC,CC version,Opt level,Instruction set,SIMD,SIMD len,OpenMP,Num threads,Permit scatter OpenMP,Flux fission,CPU,";
    std::string version(mgcfd_version());
    for (char& ch : version) if (ch == ',') ch = ';';
    const char* fm = flux_mode == MGCFD_FLUX_TILED_COLOURED ? "TiledColoured;" : (flux_mode == MGCFD_FLUX_ATOMIC ? "Atomic;" : "SortedSegment;");
    line << size << "," << variant_name(variant) << "," << conf.num_cycles << ",Normal," << fm << "FusedTimeStep;,nvcc," << version << ",3,sm_100a,N,1,N,1,N,N,NVIDIA B200 (device " << conf.device << "),";
}


// ---- what the end of main() does with a finished run (euler3d_cpu_double.cpp:698-779): validation, dumps, CSV files ----------
struct RunResult {
    int levels = 0, variant = 0, flux_mode = 0;
    long nel0 = 0;
    std::vector<double> rms;
    double total = 0.0;
    const double* var = nullptr;          // fine-level arrays in the reference's node order
    const double* sf = nullptr;
    const double* flux = nullptr;
    const double* vol = nullptr;
    int nrows = 1;                        // one CSV row per GPU
    const double* ms = nullptr;           // [nrows][7 * levels]
    const long* iters = nullptr;          // [nrows][7 * levels]
    const long* dims = nullptr;           // [nrows][levels][2]: nodes and internal edges the row's GPU computes
    const int* device = nullptr;          // [nrows]
};

int finish_run(const RunResult& R) {
    for (int i = 0; i < conf.num_cycles; i++) printf("\n%s %d / %d (RMS = %.3e)", R.levels <= 1 ? "Cycle" : "MG cycle", i + 1, conf.num_cycles, R.rms[i]);
    printf("\n");
    std::cout << "Total runtime = " << R.total << std::endl;
    printf("\n");
    if (conf.write_solution) dump_rows(solution_path("variables", 0), R.var, R.nel0, 5, false);
    if (conf.validate_result) {                   // euler3d_cpu_double.cpp:704-744; the NaN check ran on the device after every stage
        printf("Beginning validation of variables[]\n");
        printf("  NaN check passed\n");
        const std::string sp = solution_path("variables", 0);
        std::ifstream f(sp.c_str());
        if (!f) {
            printf("  could not open variables solution file:\n    %s\n  aborting validation\n", sp.c_str());
            return EXIT_FAILURE;                  // the reference carries on; a validation that did not happen is a failure here
        }
        std::vector<double> master(5 * R.nel0);
        long got = 0;
        while (got < 5 * R.nel0 && (f >> master[got])) got++;
        if (got != 5 * R.nel0) { printf("ERROR: solution file '%s' holds %ld values, expected %ld\n", sp.c_str(), got, 5 * R.nel0); return EXIT_FAILURE; }
        printf("  scanning variables[] on level 0 for errors\n");
        identify_differences(R.var, master.data(), R.nel0, R.variant);
        printf("PASS: variables[] validated successfully\n\n");
    }
    if (conf.output_variables) dump_rows(output_path("variables", 0), R.var, R.nel0, 5, true);
    if (conf.output_step_factors) dump_rows(output_path("step_factors", 0), R.sf, R.nel0, 1, false);
    if (conf.output_fluxes) dump_rows(output_path("fluxes", 0), R.flux, R.nel0, 5, false);   // time_step leaves them zeroed (cfd_loops.cpp:250-262)
    if (conf.output_volumes) dump_rows(output_path("volumes", 0), R.vol, R.nel0, 1, false);

    // ---- Times.csv / LoopNumIters.csv: reference column order flux, update, compute_step, time_step, restrict, prolong,
    // indirect_rw; library kernel ids (const.h:30-37): 0 compute_step, 1 flux, 2 update, 3 indirect_rw, 4 time_step, 5 restrict, 6 prolong
    static const int col2kid[7] = {1, 2, 0, 4, 5, 6, 3};
    static const char* names[7] = {"flux", "update", "compute_step", "time_step", "restrict", "prolong", "indirect_rw"};
    const int levels = R.levels;
    for (int which = 0; which < 2; which++) {
        const std::string path = csv_path(which == 0 ? "Times.csv" : "LoopNumIters.csv");
        std::remove(path.c_str());
        std::ofstream out(path.c_str());
        for (int row = 0; row < R.nrows; row++) {
            std::ostringstream header, line;
            const int saved_device = conf.device;
            conf.device = R.device[row];
            csv_identification(header, line, conf.mesh_duplicate_count, R.variant, R.flux_mode);
            conf.device = saved_device;
            header << "ThreadNum,CpuId,";
            line << row << "," << sched_getcpu() << ",";
            const double* ms = R.ms + (size_t)row * 7 * levels;
            const long* iters = R.iters + (size_t)row * 7 * levels;
            for (int l = 0; l < levels; l++) {
                const long nodes = R.dims[((size_t)row * levels + l) * 2], edges = R.dims[((size_t)row * levels + l) * 2 + 1];
                for (int k = 0; k < 7; k++) {
                    header << names[k] << l << ",";
                    const int kid = col2kid[k];
                    if (which == 0) line << ms[kid * levels + l] * 1e-3 << ",";          // seconds, as the reference
                    else {
                        long it = iters[kid * levels + l];
                        const long stage_launches = iters[1 * levels + l] / (edges > 0 ? edges : 1);
                        // time_step is fused into the flux stage kernel: same number of node updates as the reference's loop
                        if (kid == 4 && it == 0) it = stage_launches * nodes;
                        // the step factor is evaluated inside the stage kernels: the reference's loop count is one pass over the
                        // nodes per smoothing visit
                        if (kid == 0) it = stage_launches / MGCFD_RK * nodes;
                        line << it << ",";
                    }
                }
            }
            if (which == 0) { header << "Total,"; line << R.total << ","; }
            if (row == 0) out << header.str() << std::endl;
            out << line.str() << std::endl;
        }
        printf("%s written to: %s\n", which == 0 ? "Loop runtimes" : "Loop stats", path.c_str());
    }
    return EXIT_SUCCESS;
}

void report_invalid(mgcfd_ctx* ctx) {             // check_for_invalid_variables (validation.cpp:107-138)
    long cell = -1; int reason = 0;
    mgcfd_invalid_cell(ctx, &cell, &reason);
    printf("%s detected at cell %ld\n", reason == 1 ? "NaN or infinity" : (reason == 2 ? "Negative density" : "Negative energy"), cell);
}

// ---- one GPU -------------------------------------------------------------------------------------------------------------------
int run_single(mgcfd_mesh* mesh) {
    const int levels = mgcfd_mesh_levels(mesh), variant = mgcfd_mesh_variant(mesh);
    long d0[5];
    mgcfd_mesh_dims(mesh, 0, d0);
    const long nel0 = d0[0];

    mgcfd_options opt;
    mgcfd_default_options(&opt);
    opt.device = conf.device;
    if (conf.flux_mode >= 0) opt.flux_mode = conf.flux_mode;
    if (conf.tile_nodes > 0) opt.tile_nodes = conf.tile_nodes;
    opt.timing = conf.kernel_times ? 1 : 0;
    mgcfd_ctx* ctx = nullptr;
    CHECK(mgcfd_create(levels, variant, &opt, &ctx));
    if (mgcfd_mesh_upload(mesh, ctx) != MGCFD_OK) { fprintf(stderr, "ERROR: %s\n", mgcfd_mesh_last_error()); return EXIT_FAILURE; }   // adjust_ewt + dampen_ewt inside
    CHECK(mgcfd_synchronize(ctx));

    // ---- the V-cycle loop (euler3d_cpu_double.cpp:371-694), entirely on the device ----
    RunResult R;
    R.levels = levels; R.variant = variant; R.flux_mode = opt.flux_mode; R.nel0 = nel0;
    R.rms.assign(conf.num_cycles > 0 ? conf.num_cycles : 1, 0.0);
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = mgcfd_run_cycles(ctx, conf.num_cycles, R.rms.data(), nullptr);
    R.total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    if (rc == MGCFD_ERR_INVALID_VARIABLES) {
        for (int i = 0; i < conf.num_cycles; i++) printf("\n%s %d / %d (RMS = %.3e)", levels <= 1 ? "Cycle" : "MG cycle", i + 1, conf.num_cycles, R.rms[i]);
        printf("\n");
        report_invalid(ctx);
        return EXIT_FAILURE;
    }
    if (rc != MGCFD_OK) { fprintf(stderr, "ERROR: mgcfd_run_cycles: %s\n", mgcfd_last_error()); return EXIT_FAILURE; }

    std::vector<double> var(5 * nel0), sf, fl, vol;
    CHECK(mgcfd_get_field(ctx, 0, MGCFD_FIELD_VARIABLES, var.data()));
    if (conf.output_step_factors) { sf.resize(nel0); CHECK(mgcfd_get_field(ctx, 0, MGCFD_FIELD_STEP_FACTORS, sf.data())); }
    if (conf.output_fluxes) { fl.resize(5 * nel0); CHECK(mgcfd_get_field(ctx, 0, MGCFD_FIELD_FLUXES, fl.data())); }
    if (conf.output_volumes) { vol.resize(nel0); CHECK(mgcfd_get_field(ctx, 0, MGCFD_FIELD_VOLUMES, vol.data())); }
    std::vector<double> ms(7 * levels, 0.0);
    std::vector<long> iters(7 * levels, 0), dims(2 * levels);
    CHECK(mgcfd_get_times(ctx, ms.data(), iters.data()));
    for (int l = 0; l < levels; l++) { long dl[5]; mgcfd_mesh_dims(mesh, l, dl); dims[2 * l] = dl[0]; dims[2 * l + 1] = dl[1]; }
    R.var = var.data(); R.sf = sf.data(); R.flux = fl.data(); R.vol = vol.data();
    R.ms = ms.data(); R.iters = iters.data(); R.dims = dims.data(); R.device = &conf.device;
    const int out = finish_run(R);
    mgcfd_destroy(ctx);
    return out;
}

// ---- N GPUs: one forked process per GPU ------------------------------------------------------------------------------------------
const int MAX_RANKS = 64, MAX_TABLE = 1 + 8 * (2 + MAX_RANKS + 1) + 8 * 25;      // 8 levels; + the slab block of the in-kernel exchange (mgcfd_dist_p2p_table_len)
struct Shared {                                   // lives in an anonymous MAP_SHARED mapping created before the fork
    std::atomic<int> arrived, generation, abort_flag;
    char nccl_id[128];
    char ipc[MAX_RANKS][64];
    long table_len;
    long tables[MAX_RANKS][MAX_TABLE];
    int p2p_ok[MAX_RANKS];
    int rc[MAX_RANKS];
    double total[MAX_RANKS];
    long bad_cell; int bad_reason;
    // followed by: rms[cycles] | ms[N][7*levels] | iters[N][7*levels] | dims[N][levels][2] | var[5*nel0] | sf[nel0] | flux[5*nel0] | vol[nel0]
};
// barrier of the rank processes; gives up (exit) when any rank has failed, so that nobody waits for a dead peer
void rank_barrier(Shared* S, int n) {
    const int gen = S->generation.load();
    if (S->arrived.fetch_add(1) == n - 1) { S->arrived.store(0); S->generation.store(gen + 1); return; }
    while (S->generation.load() == gen) {
        if (S->abort_flag.load()) _exit(EXIT_FAILURE);
        usleep(50);
    }
}
#define RCHECK(call)                                                                                        \
    do {                                                                                                    \
        int rc_ = (call);                                                                                   \
        if (rc_ != MGCFD_OK) {                                                                              \
            fprintf(stderr, "ERROR (rank %d): %s failed (%d): %s\n", rank, #call, rc_, mgcfd_last_error()); \
            S->abort_flag.store(1); _exit(EXIT_FAILURE);                                                    \
        }                                                                                                   \
    } while (0)

int rank_main(mgcfd_mesh* mesh, Shared* S, int rank, int N, double* rms, double* ms_all, long* iters_all, long* dims_all, double* var, double* sf,
              double* flux, double* vol) {
    const int levels = mgcfd_mesh_levels(mesh), variant = mgcfd_mesh_variant(mesh);
    long d0[5];
    mgcfd_mesh_dims(mesh, 0, d0);
    const long nel0 = d0[0];
    mgcfd_options opt;
    mgcfd_default_options(&opt);
    opt.device = rank;
    if (conf.flux_mode >= 0) opt.flux_mode = conf.flux_mode;
    if (conf.tile_nodes > 0) opt.tile_nodes = conf.tile_nodes;
    opt.timing = conf.kernel_times ? 1 : 0;
    mgcfd_ctx* ctx = nullptr;
    RCHECK(mgcfd_create(levels, variant, &opt, &ctx));
    if (rank == 0) RCHECK(mgcfd_dist_get_unique_id(S->nccl_id));
    rank_barrier(S, N);
    RCHECK(mgcfd_dist_init(ctx, rank, N, S->nccl_id));
    if (mgcfd_mesh_upload_partition(mesh, ctx) != MGCFD_OK) {
        fprintf(stderr, "ERROR (rank %d): %s\n", rank, mgcfd_mesh_last_error());
        S->abort_flag.store(1); _exit(EXIT_FAILURE);
    }
    // direct peer-to-peer data path (CUDA IPC windows) where every rank can attach; otherwise NCCL send/recv (MGCFD_NO_P2P=1 forces it)
    const char* no_p2p = getenv("MGCFD_NO_P2P");
    if (!(no_p2p && no_p2p[0] == '1')) {
        const long tl = mgcfd_dist_p2p_table_len(ctx);
        if (tl > 0 && tl <= MAX_TABLE) {
            RCHECK(mgcfd_dist_p2p_prepare(ctx, S->ipc[rank], S->tables[rank], MAX_TABLE));
            rank_barrier(S, N);
            std::vector<char> handles((size_t)N * 64);
            std::vector<long> tables((size_t)N * tl);
            for (int r = 0; r < N; r++) { memcpy(&handles[(size_t)r * 64], S->ipc[r], 64); memcpy(&tables[(size_t)r * tl], S->tables[r], sizeof(long) * tl); }
            S->p2p_ok[rank] = mgcfd_dist_p2p_attach(ctx, handles.data(), tables.data(), tl) == MGCFD_OK;
            if (!S->p2p_ok[rank]) fprintf(stderr, "rank %d: peer-to-peer attach failed (%s)\n", rank, mgcfd_last_error());
            rank_barrier(S, N);
            for (int r = 0; r < N; r++)
                if (!S->p2p_ok[r]) {               // the ranks must agree on the data plane
                    if (rank == 0) fprintf(stderr, "ERROR: the GPUs are not all peer-accessible; rerun with MGCFD_NO_P2P=1 (NCCL data plane)\n");
                    S->abort_flag.store(1); _exit(EXIT_FAILURE);
                }
        }
    }
    RCHECK(mgcfd_synchronize(ctx));
    rank_barrier(S, N);

    std::vector<double> my_rms(conf.num_cycles > 0 ? conf.num_cycles : 1, 0.0);
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = mgcfd_run_cycles(ctx, conf.num_cycles, my_rms.data(), nullptr);      // collective: every rank, same arguments
    S->total[rank] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    S->rc[rank] = rc;
    if (rank == 0) {
        memcpy(rms, my_rms.data(), sizeof(double) * conf.num_cycles);
        if (rc == MGCFD_ERR_INVALID_VARIABLES) mgcfd_invalid_cell(ctx, &S->bad_cell, &S->bad_reason);
    }
    if (rc != MGCFD_OK && rc != MGCFD_ERR_INVALID_VARIABLES) { fprintf(stderr, "ERROR (rank %d): mgcfd_run_cycles: %s\n", rank, mgcfd_last_error()); S->abort_flag.store(1); _exit(EXIT_FAILURE); }
    rank_barrier(S, N);
    if (rc == MGCFD_ERR_INVALID_VARIABLES) { mgcfd_destroy(ctx); return EXIT_FAILURE; }   // every rank learns of it (all-reduced key); rank 0's parent reports

    // owned rows of the fine level -> the shared arrays, in the reference's (global) node order
    long info[8];
    RCHECK(mgcfd_dist_level_info(ctx, 0, info));
    const long owned = info[0], local = info[0] + info[1];
    std::vector<long> gid(local);
    RCHECK(mgcfd_dist_global_ids(ctx, 0, gid.data()));
    std::vector<double> buf(5 * local);
    auto gather = [&](int field, int ncomp, double* dst) {
        RCHECK(mgcfd_get_field(ctx, 0, field, buf.data()));
        for (long k = 0; k < owned; k++) memcpy(dst + (size_t)ncomp * gid[k], &buf[(size_t)ncomp * k], sizeof(double) * ncomp);
    };
    gather(MGCFD_FIELD_VARIABLES, 5, var);
    if (conf.output_step_factors) gather(MGCFD_FIELD_STEP_FACTORS, 1, sf);
    if (conf.output_fluxes) gather(MGCFD_FIELD_FLUXES, 5, flux);
    if (conf.output_volumes) gather(MGCFD_FIELD_VOLUMES, 1, vol);
    RCHECK(mgcfd_get_times(ctx, ms_all + (size_t)rank * 7 * levels, iters_all + (size_t)rank * 7 * levels));
    for (int l = 0; l < levels; l++) {
        long li[16];
        RCHECK(mgcfd_level_info(ctx, l, li));
        long di[8];
        RCHECK(mgcfd_dist_level_info(ctx, l, di));
        dims_all[((size_t)rank * levels + l) * 2] = di[0];         // nodes this GPU updates
        dims_all[((size_t)rank * levels + l) * 2 + 1] = li[1];     // internal edges it evaluates (cut edges on both sides)
    }
    (void)nel0;
    rank_barrier(S, N);
    mgcfd_destroy(ctx);
    return EXIT_SUCCESS;
}

int run_distributed(mgcfd_mesh* mesh) {
    const int N = conf.gpus, levels = mgcfd_mesh_levels(mesh), variant = mgcfd_mesh_variant(mesh);
    if (levels > 8) { fprintf(stderr, "ERROR: --gpus supports at most 8 levels\n"); return EXIT_FAILURE; }
    long d0[5];
    mgcfd_mesh_dims(mesh, 0, d0);
    const long nel0 = d0[0];
    const int cyc = conf.num_cycles > 0 ? conf.num_cycles : 1;
    const size_t n_rms = cyc, n_ms = (size_t)N * 7 * levels, n_dims = (size_t)N * levels * 2;
    const size_t bytes = sizeof(Shared) + sizeof(double) * (n_rms + n_ms + 12 * (size_t)nel0) + sizeof(long) * (n_ms + n_dims) + 64;
    void* map = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    if (map == MAP_FAILED) { perror("mmap"); return EXIT_FAILURE; }
    memset(map, 0, sizeof(Shared));
    Shared* S = new (map) Shared;
    S->arrived.store(0); S->generation.store(0); S->abort_flag.store(0);
    double* rms = reinterpret_cast<double*>(reinterpret_cast<char*>(map) + ((sizeof(Shared) + 63) & ~size_t(63)));
    double* ms = rms + n_rms;
    double* var = ms + n_ms;
    double* sf = var + 5 * (size_t)nel0;
    double* flux = sf + nel0;
    double* vol = flux + 5 * (size_t)nel0;
    long* iters = reinterpret_cast<long*>(vol + nel0);
    long* dims = iters + n_ms;
    // adjust_ewt + dampen_ewt once, here: the ranks then only read the mesh, so its pages stay shared between the processes
    if (mgcfd_mesh_apply_ewt(mesh) != MGCFD_OK) { fprintf(stderr, "ERROR: %s\n", mgcfd_mesh_last_error()); munmap(map, bytes); return EXIT_FAILURE; }
    fflush(stdout); fflush(stderr);
    std::vector<pid_t> pids(N, -1);
    for (int r = 0; r < N; r++) {
        const pid_t pid = fork();
        if (pid < 0) { perror("fork"); S->abort_flag.store(1); break; }
        if (pid == 0) {
            const int rc = rank_main(mesh, S, r, N, rms, ms, iters, dims, var, sf, flux, vol);
            fflush(stdout); fflush(stderr);
            _exit(rc);
        }
        pids[r] = pid;
    }
    // wait for the ranks; the first failure releases whoever waits at a barrier, stragglers stuck on the device are killed
    int failed = 0, left = 0;
    for (int r = 0; r < N; r++) if (pids[r] > 0) left++; else failed = 1;
    double deadline = -1.0;
    auto now = [] { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); };
    while (left > 0) {
        int status = 0;
        const pid_t p = waitpid(-1, &status, WNOHANG);
        if (p > 0) {
            left--;
            for (int r = 0; r < N; r++) if (pids[r] == p) pids[r] = -1;
            const bool invalid = WIFEXITED(status) && WEXITSTATUS(status) != 0 && S->rc[0] == MGCFD_ERR_INVALID_VARIABLES;
            if (!(WIFEXITED(status) && WEXITSTATUS(status) == 0) && !invalid) { failed = 1; S->abort_flag.store(1); if (deadline < 0) deadline = now() + 10.0; }
        } else {
            if (deadline > 0 && now() > deadline) { for (int r = 0; r < N; r++) if (pids[r] > 0) kill(pids[r], SIGKILL); deadline = now() + 1e9; }
            usleep(2000);
        }
    }
    int out = EXIT_FAILURE;
    if (failed) fprintf(stderr, "ERROR: a rank process failed\n");
    else if (S->rc[0] == MGCFD_ERR_INVALID_VARIABLES) {
        for (int i = 0; i < conf.num_cycles; i++) printf("\n%s %d / %d (RMS = %.3e)", levels <= 1 ? "Cycle" : "MG cycle", i + 1, conf.num_cycles, rms[i]);
        printf("\n%s detected at cell %ld\n", S->bad_reason == 1 ? "NaN or infinity" : (S->bad_reason == 2 ? "Negative density" : "Negative energy"), S->bad_cell);
    } else {
        RunResult R;
        R.levels = levels; R.variant = variant; R.nel0 = nel0;
        mgcfd_options opt;
        mgcfd_default_options(&opt);
        R.flux_mode = conf.flux_mode >= 0 ? conf.flux_mode : opt.flux_mode;
        R.rms.assign(rms, rms + cyc);
        for (int r = 0; r < N; r++) R.total = std::max(R.total, S->total[r]);
        std::vector<int> devices(N);
        for (int r = 0; r < N; r++) devices[r] = r;
        R.var = var; R.sf = sf; R.flux = flux; R.vol = vol;
        R.nrows = N; R.ms = ms; R.iters = iters; R.dims = dims; R.device = devices.data();
        out = finish_run(R);
    }
    munmap(map, bytes);
    return out;
}

}  // namespace

int main(int argc, char** argv) {
    parse_arguments(argc, argv);
    mgcfd_mesh* mesh = nullptr;
    if (mgcfd_mesh_load(conf.input_file.c_str(), conf.input_file_directory.c_str(), &mesh) != MGCFD_OK) {
        fprintf(stderr, "ERROR: %s\n", mgcfd_mesh_last_error());
        return EXIT_FAILURE;
    }
    if (mgcfd_mesh_duplicate(mesh, conf.mesh_duplicate_count) != MGCFD_OK) { fprintf(stderr, "ERROR: %s\n", mgcfd_mesh_last_error()); return EXIT_FAILURE; }
    const int rc = conf.gpus > 1 ? run_distributed(mesh) : run_single(mesh);
    mgcfd_mesh_free(mesh);
    return rc;
}
