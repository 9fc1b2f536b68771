// driver_main.cpp -- euler3d_b200: the reference's euler3d driver (src/euler3d_cpu_double.cpp:69-809) on top of the C ABI.
// Same command line (src/Base/config.cpp:32-47, :281-305), same key = value config file (:81-217), same input.dat / mesh files,
// same progress lines ("MG cycle i / n (RMS = ...)", "Total runtime = ..."), same dump files (variables / step_factors / fluxes,
// %.17e, io.cpp:201-233, io_enhanced.cpp:652-817), same -v validation rule (validation.cpp:140-199) and the same
// Times.csv / LoopNumIters.csv schema (timer.cpp:106-195, loop_stats.cpp:83-171) with CUDA-event times; CPU-specific
// identification columns carry the GPU equivalents.  Host code only: every number comes from libmgcfd_b200.so.
#include <getopt.h>
#include <sched.h>

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

namespace {

struct Config {                                   // config.h:27-47
    std::string config_filepath, input_file, input_file_directory, papi_config_file, output_file_prefix;
    int mesh_duplicate_count = 1, num_cycles = 25, omp_num_threads = 1;
    bool validate_result = false, output_variables = false, output_fluxes = false, output_step_factors = false, output_volumes = false;
    // not in the reference: device selection and timing granularity
    int device = 0, flux_mode = -1, tile_nodes = 0;
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
        else printf("WARNING: Unknown key '%s' encountered during parsing of config file.\n", key.c_str());
    }
}

void parse_arguments(int argc, char** argv) {     // config.cpp:219-259
    static struct option long_opts[] = {
        {"help", no_argument, nullptr, 'h'}, {"config-filepath", required_argument, nullptr, 'c'}, {"input-file", required_argument, nullptr, 'i'},
        {"input-directory", required_argument, nullptr, 'd'}, {"papi_config_file", required_argument, nullptr, 'p'},
        {"output-file-prefix", required_argument, nullptr, 'o'}, {"mesh-duplicate-count", required_argument, nullptr, 'm'},
        {"num-cycles", required_argument, nullptr, 'g'}, {"validate-result", no_argument, nullptr, 'v'},
        {"output-variables", no_argument, nullptr, 1001}, {"output-fluxes", no_argument, nullptr, 1002}, {"output-step-factors", no_argument, nullptr, 1003},
        {"device", required_argument, nullptr, 1004}, {"flux-mode", required_argument, nullptr, 1005}, {"tile-nodes", required_argument, nullptr, 1006},
        {"no-kernel-times", no_argument, nullptr, 1007}, {"write-solution", no_argument, nullptr, 1008}, {nullptr, 0, nullptr, 0}};
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
            default: print_help(); exit(EXIT_FAILURE);
        }
    }
    if (conf.input_file.empty()) { fprintf(stderr, "ERROR: Input file not specified\n"); print_help(); exit(EXIT_FAILURE); }
    if (conf.mesh_duplicate_count < 1 || conf.num_cycles < 0) { fprintf(stderr, "ERROR: bad -m / -g value\n"); exit(EXIT_FAILURE); }
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

void dump_rows(const std::string& path, const double* v, long n, int ncomp, bool announce) {
    FILE* f = fopen(path.c_str(), "w");
    if (!f) { fprintf(stderr, "ERROR: Failed to open file for writing: '%s'\n", path.c_str()); exit(EXIT_FAILURE); }
    if (announce) printf("Dumping variables[] to file: %s\n", path.c_str());
    for (long i = 0; i < n; i++) {
        if (ncomp == 5) fprintf(f, "%.17e %.17e %.17e %.17e %.17e\n", v[5 * i], v[5 * i + 1], v[5 * i + 2], v[5 * i + 3], v[5 * i + 4]);
        else fprintf(f, "%.17e\n", v[i]);
    }
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

}  // namespace

int main(int argc, char** argv) {
    parse_arguments(argc, argv);
    mgcfd_mesh* mesh = nullptr;
    if (mgcfd_mesh_load(conf.input_file.c_str(), conf.input_file_directory.c_str(), &mesh) != MGCFD_OK) {
        fprintf(stderr, "ERROR: %s\n", mgcfd_mesh_last_error());
        return EXIT_FAILURE;
    }
    if (mgcfd_mesh_duplicate(mesh, conf.mesh_duplicate_count) != MGCFD_OK) { fprintf(stderr, "ERROR: %s\n", mgcfd_mesh_last_error()); return EXIT_FAILURE; }
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
    std::vector<double> rms(conf.num_cycles > 0 ? conf.num_cycles : 1);
    const auto t0 = std::chrono::steady_clock::now();
    const int rc = mgcfd_run_cycles(ctx, conf.num_cycles, rms.data(), nullptr);
    const double total = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    for (int i = 0; i < conf.num_cycles; i++) printf("\n%s %d / %d (RMS = %.3e)", levels <= 1 ? "Cycle" : "MG cycle", i + 1, conf.num_cycles, rms[i]);
    printf("\n");
    if (rc == MGCFD_ERR_INVALID_VARIABLES) {      // check_for_invalid_variables (validation.cpp:107-138)
        long cell = -1; int reason = 0;
        mgcfd_invalid_cell(ctx, &cell, &reason);
        printf("%s detected at cell %ld\n", reason == 1 ? "NaN or infinity" : (reason == 2 ? "Negative density" : "Negative energy"), cell);
        return EXIT_FAILURE;
    }
    if (rc != MGCFD_OK) { fprintf(stderr, "ERROR: mgcfd_run_cycles: %s\n", mgcfd_last_error()); return EXIT_FAILURE; }
    std::cout << "Total runtime = " << total << std::endl;

    std::vector<double> var(5 * nel0);
    CHECK(mgcfd_get_field(ctx, 0, MGCFD_FIELD_VARIABLES, var.data()));
    if (conf.write_solution) dump_rows(solution_path("variables", 0), var.data(), nel0, 5, false);
    if (conf.validate_result) {                   // euler3d_cpu_double.cpp:704-744
        const std::string sp = solution_path("variables", 0);
        std::ifstream f(sp.c_str());
        if (!f) { printf("ERROR: solution file not present: %s\n", sp.c_str()); return EXIT_FAILURE; }
        std::vector<double> master(5 * nel0);
        long got = 0;
        while (got < 5 * nel0 && (f >> master[got])) got++;
        if (got != 5 * nel0) { printf("ERROR: solution file '%s' holds %ld values, expected %ld\n", sp.c_str(), got, 5 * nel0); return EXIT_FAILURE; }
        identify_differences(var.data(), master.data(), nel0, variant);
        printf("PASS: No errors detected\n");
    }
    if (conf.output_variables) dump_rows(output_path("variables", 0), var.data(), nel0, 5, true);
    if (conf.output_step_factors) {
        std::vector<double> sf(nel0);
        CHECK(mgcfd_get_field(ctx, 0, MGCFD_FIELD_STEP_FACTORS, sf.data()));
        dump_rows(output_path("step_factors", 0), sf.data(), nel0, 1, false);
    }
    if (conf.output_fluxes) {                     // time_step leaves the fluxes zeroed (cfd_loops.cpp:250-262): so does this dump
        std::vector<double> fl(5 * nel0);
        CHECK(mgcfd_get_field(ctx, 0, MGCFD_FIELD_FLUXES, fl.data()));
        dump_rows(output_path("fluxes", 0), fl.data(), nel0, 5, false);
    }
    if (conf.output_volumes) {
        std::vector<double> vol(nel0);
        CHECK(mgcfd_get_field(ctx, 0, MGCFD_FIELD_VOLUMES, vol.data()));
        dump_rows(output_path("volumes", 0), vol.data(), nel0, 1, false);
    }

    // ---- Times.csv / LoopNumIters.csv ----
    std::vector<double> ms(7 * levels, 0.0);
    std::vector<long> iters(7 * levels, 0);
    CHECK(mgcfd_get_times(ctx, ms.data(), iters.data()));
    // reference column order: flux, update, compute_step, time_step, restrict, prolong, indirect_rw; library kernel ids
    // (const.h:30-37): 0 compute_step, 1 flux, 2 update, 3 indirect_rw, 4 time_step, 5 restrict, 6 prolong
    static const int col2kid[7] = {1, 2, 0, 4, 5, 6, 3};
    for (int which = 0; which < 2; which++) {
        const std::string path = csv_path(which == 0 ? "Times.csv" : "LoopNumIters.csv");
        std::remove(path.c_str());
        std::ostringstream header, line;
        csv_identification(header, line, conf.mesh_duplicate_count, variant, opt.flux_mode);
        header << "ThreadNum,CpuId,";
        line << 0 << "," << sched_getcpu() << ",";
        for (int l = 0; l < levels; l++) {
            static const char* names[7] = {"flux", "update", "compute_step", "time_step", "restrict", "prolong", "indirect_rw"};
            for (int k = 0; k < 7; k++) {
                header << names[k] << l << ",";
                const int kid = col2kid[k];
                if (which == 0) line << ms[kid * levels + l] * 1e-3 << ",";          // seconds, as the reference
                else {
                    long it = iters[kid * levels + l];
                    // time_step is fused into the flux stage kernel: same number of node updates as the reference's loop
                    if (kid == 4 && it == 0) { long dl[5]; mgcfd_mesh_dims(mesh, l, dl); it = (iters[1 * levels + l] / (dl[1] > 0 ? dl[1] : 1)) * dl[0]; }
                    // the step factor is evaluated inside the stage kernels and its global minimum inside the transfer kernels: the
                    // reference's loop count is one pass over the nodes per smoothing visit
                    if (kid == 0) { long dl[5]; mgcfd_mesh_dims(mesh, l, dl); it = (iters[1 * levels + l] / (dl[1] > 0 ? dl[1] : 1) / MGCFD_RK) * dl[0]; }
                    line << it << ",";
                }
            }
        }
        if (which == 0) { header << "Total,"; line << total << ","; }
        std::ofstream out(path.c_str());
        out << header.str() << std::endl << line.str() << std::endl;
        printf("%s written to: %s\n", which == 0 ? "Loop runtimes" : "Loop stats", path.c_str());
    }
    mgcfd_destroy(ctx);
    mgcfd_mesh_free(mesh);
    return EXIT_SUCCESS;
}
