// assess_host_check.cu -- CPU check (no GPU: the __host__ side of the same templates) that the assess-compute flux variants of
// csrc/assess_kernels.cuh reproduce the oracle's compute_flux_edge.  Built and run by tests/test_host_mesh.py:
//   nvcc -O2 -std=c++17 -Img-cfd-app-plain_b200/csrc -o /tmp/assess_check tools/assess_host_check.cu oracle/mgcfd_oracle.o
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "assess_kernels.cuh"

struct orc_edge { long a, b; double x, y, z; };
extern "C" void orc_flux_edge(long first, long n, const orc_edge* e, const double* var, double* flux);

template <bool D, bool F>
static double run(long nn, const std::vector<orc_edge>& E, const std::vector<double>& recs, const std::vector<double>& want) {
    std::vector<double> got(5 * nn, 0.0);
    for (const orc_edge& e : E) {
        double av[5], bv[5];
        mgcfd::assess_edge<D, F>(recs.data(), e.a, e.b, e.x, e.y, e.z, std::sqrt(e.x * e.x + e.y * e.y + e.z * e.z), double(0.2f), av, bv);
        for (int k = 0; k < 5; k++) { got[5 * e.a + k] += av[k]; got[5 * e.b + k] += bv[k]; }
    }
    double worst = 0.0;
    for (int k = 0; k < 5; k++) {
        double scale = 0.0, err = 0.0;
        for (long i = 0; i < nn; i++) { scale = std::fmax(scale, std::fabs(want[5 * i + k])); err = std::fmax(err, std::fabs(got[5 * i + k] - want[5 * i + k])); }
        worst = std::fmax(worst, err / scale);
    }
    return worst;
}

int main() {
    const long nn = 4000, ne = 15000;
    std::vector<double> var(5 * nn), recs(8 * nn, 0.0);
    unsigned long long s = 12345;
    auto rnd = [&]() { s ^= s << 13; s ^= s >> 7; s ^= s << 17; return double(s % 1000000) / 1e6; };
    for (long i = 0; i < nn; i++) {
        var[5 * i] = 1.4 * (0.9 + 0.2 * rnd()); var[5 * i + 1] = 1.68 * (0.9 + 0.2 * rnd()); var[5 * i + 2] = 0.1 * (rnd() - 0.5);
        var[5 * i + 3] = 0.1 * (rnd() - 0.5); var[5 * i + 4] = 3.508 * (0.95 + 0.1 * rnd());
        for (int k = 0; k < 5; k++) recs[8 * i + k] = var[5 * i + k];
    }
    std::vector<orc_edge> E(ne);
    for (long e = 0; e < ne; e++) {
        long a = long(rnd() * nn) % nn, b = long(rnd() * nn) % nn;
        if (a == b) b = (a + 1) % nn;
        E[e] = {a, b, 1e-3 * (rnd() - 0.5), 1e-3 * (rnd() - 0.5), 1e-3 * (rnd() - 0.5)};
    }
    std::vector<double> want(5 * nn, 0.0);
    orc_flux_edge(0, ne, E.data(), var.data(), want.data());
    const double w[4] = {run<false, false>(nn, E, recs, want), run<true, false>(nn, E, recs, want), run<false, true>(nn, E, recs, want), run<true, true>(nn, E, recs, want)};
    printf("max rel err: default %.2e reuse_div %.2e reuse_flux %.2e both %.2e\n", w[0], w[1], w[2], w[3]);
    for (double x : w) if (!(x < 1e-13)) { printf("FAIL\n"); return 1; }
    printf("PASS\n");
    return 0;
}
