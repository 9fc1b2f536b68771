// tools/microbench.cu -- the pipe rates the flux kernel design is budgeted against, measured on the box
// (MEASURED_PEAKS.json has HBM copy and bf16 GEMM only): FP64 FMA issue rate, fp64 divide / sqrt / rcp cost,
// shared-memory 64- and 128-bit load rate.  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mb tools/microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int ILP>
__global__ void k_dfma(double* out, int iters, double a, double b) {
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; i++) x[i] = threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < ILP; i++) x[i] = fma(x[i], a, b);
    }
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; i++) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
__global__ void k_special(double* out, int iters, double a) {
    double x[4];
#pragma unroll
    for (int i = 0; i < 4; i++) x[i] = 1.0 + threadIdx.x * 1e-3 + i;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int i = 0; i < 4; i++) {
            if (OP == 0) x[i] = a / x[i] + 1.0;
            else if (OP == 1) x[i] = sqrt(x[i]) + a;
            else if (OP == 2) x[i] = __drcp_rn(x[i]) + a;
            else x[i] = rsqrt(x[i]) + a;
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = x[0] + x[1] + x[2] + x[3];
}
template <int W>   // W = 1: LDS.64, 2: LDS.128
__global__ void k_lds(double* out, int iters, int stride) {
    extern __shared__ double sm[];
    for (int i = threadIdx.x; i < 4096; i += blockDim.x) sm[i] = i;
    __syncthreads();
    double s = 0;
    int idx = (threadIdx.x * stride * W) & 4095;
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (W == 1) s += sm[(idx + u * 64) & 4095];
            else { double2 v = *reinterpret_cast<double2*>(&sm[(idx + u * 64) & 4094]); s += v.x + v.y; }
        }
        idx = (idx + 1) & 4095;
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

int main() {
    cudaDeviceProp p;
    CK(cudaGetDeviceProperties(&p, 0));
    int clk = 0;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("device %s, %d SMs, L2 %d MB, max clock %d MHz, smem/SM %zu KB\n", p.name, p.multiProcessorCount, p.l2CacheSize >> 20, clk / 1000,
           p.sharedMemPerMultiprocessor >> 10);
    const int nsm = p.multiProcessorCount;
    double* out;
    CK(cudaMalloc(&out, sizeof(double) * nsm * 8 * 1024));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    float ms;
    const int blocks = nsm * 4, threads = 512;
    {
        const int iters = 1 << 15;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            k_dfma<8><<<blocks, threads>>>(out, iters, 1.0000001, 1e-9);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        const double n = double(blocks) * threads * iters * 8;
        printf("DFMA: %.2f T FMA/s = %.1f TFLOP/s fp64, %.1f lanes/clk/SM at %d MHz\n", n / ms / 1e9, 2 * n / ms / 1e9, n / (ms * 1e-3) / nsm / (clk * 1e3), clk / 1000);
    }
    const char* names[4] = {"div", "sqrt", "__drcp_rn", "rsqrt"};
    for (int op = 0; op < 4; op++) {
        const int iters = 1 << 12;
        for (int rep = 0; rep < 2; rep++) {
            cudaEventRecord(e0);
            if (op == 0) k_special<0><<<blocks, threads>>>(out, iters, 1.5);
            if (op == 1) k_special<1><<<blocks, threads>>>(out, iters, 1.5);
            if (op == 2) k_special<2><<<blocks, threads>>>(out, iters, 1.5);
            if (op == 3) k_special<3><<<blocks, threads>>>(out, iters, 1.5);
            cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
        }
        const double n = double(blocks) * threads * iters * 4;
        printf("fp64 %s (+1 add): %.3f T op/s, %.2f lanes/clk/SM  => ~%.1f DFMA-equivalents each\n", names[op], n / ms / 1e9,
               n / (ms * 1e-3) / nsm / (clk * 1e3), 64.0 / (n / (ms * 1e-3) / nsm / (clk * 1e3)));
    }
    for (int w = 1; w <= 2; w++)
        for (int stride = 1; stride <= 2; stride++) {
            const int iters = 1 << 12;
            for (int rep = 0; rep < 2; rep++) {
                cudaEventRecord(e0);
                if (w == 1) k_lds<1><<<blocks, threads, 4096 * 8>>>(out, iters, stride);
                else k_lds<2><<<blocks, threads, 4096 * 8>>>(out, iters, stride);
                cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&ms, e0, e1);
            }
            const double bytes = double(blocks) * threads * iters * 8 * 8 * w;
            printf("LDS.%d stride %d: %.1f TB/s, %.1f B/clk/SM\n", 64 * w, stride, bytes / ms / 1e9, bytes / (ms * 1e-3) / nsm / (clk * 1e3));
        }
    return 0;
}
