// p2p_pingpong.cu -- latency of the flag protocols the multi-GPU data plane could use, between two peer-accessible GPUs of one
// node (one process, two devices):   nvcc -arch=sm_100a -O3 -o /tmp/p2p_pingpong tools/p2p_pingpong.cu && /tmp/p2p_pingpong
// Two kernels (one per GPU) bounce a counter N times: each writes k into the OTHER GPU's flag and waits until its own flag reads k.
// Variants: how the flag is written (st.release.sys / plain volatile store after __threadfence_system) and polled (ld.acquire.sys /
// ld.relaxed.sys (volatile) + one fence), with and without a payload of remote stores before the signal, with and without nanosleep.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long* p) { unsigned long long v; asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long* p) { unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v; }
__device__ __forceinline__ void st_release_sys(unsigned long long* p, unsigned long long v) { asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
__device__ __forceinline__ void st_relaxed_sys(unsigned long long* p, unsigned long long v) { asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory"); }
// mode bit0: poll with relaxed loads + one acq_rel fence (else acquire loads); bit1: nanosleep(20) in the poll loop;
// bit2: signal = __threadfence_system + relaxed store (else st.release.sys); payload: doubles stored remotely before every signal
__global__ void k_bounce(unsigned long long* remote_flag, const unsigned long long* my_flag, double* remote_buf, int payload, int n, int first, int mode, long long* clocks) {
    const int t = threadIdx.x;
    long long t0 = clock64();
    for (int k = 1; k <= n; k++) {
        if (!first || k > 1) {      // wait for round k - (first ? 1 : 0)
            const unsigned long long want = first ? k - 1 : k;
            if (t == 0) {
                if (mode & 1) { while (ld_relaxed_sys(my_flag) < want) { if (mode & 2) __nanosleep(20); } asm volatile("fence.acq_rel.sys;" ::: "memory"); }
                else { while (ld_acquire_sys(my_flag) < want) { if (mode & 2) __nanosleep(20); } }
            }
            __syncthreads();
        }
        for (int i = t; i < payload; i += blockDim.x) remote_buf[i] = double(k);
        __syncthreads();
        if (t == 0) {
            if (mode & 4) { __threadfence_system(); st_relaxed_sys(remote_flag, k); }
            else st_release_sys(remote_flag, k);
        }
    }
    if (t == 0) clocks[0] = clock64() - t0;
}
int main() {
    int nd = 0; CK(cudaGetDeviceCount(&nd));
    if (nd < 2) { printf("needs 2 GPUs\n"); return 0; }
    int can = 0; CK(cudaDeviceCanAccessPeer(&can, 0, 1)); printf("peer access 0->1: %d\n", can);
    unsigned long long* flag[2]; double* buf[2]; long long* clk[2]; cudaStream_t st[2];
    for (int d = 0; d < 2; d++) {
        CK(cudaSetDevice(d)); CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&flag[d], 256)); CK(cudaMalloc(&buf[d], 8 << 20)); CK(cudaMalloc(&clk[d], 64)); CK(cudaStreamCreate(&st[d]));
    }
    const int n = 2000;
    for (int payload : {0, 4096, 65536})
        for (int mode = 0; mode < 8; mode++) {
            for (int d = 0; d < 2; d++) { CK(cudaSetDevice(d)); CK(cudaMemset(flag[d], 0, 256)); CK(cudaDeviceSynchronize()); }
            cudaEvent_t e0, e1; CK(cudaSetDevice(0)); CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
            CK(cudaEventRecord(e0, st[0]));
            k_bounce<<<1, 256, 0, st[0]>>>(flag[1], flag[0], buf[1], payload, n, 1, mode, clk[0]);
            CK(cudaEventRecord(e1, st[0]));
            CK(cudaSetDevice(1));
            k_bounce<<<1, 256, 0, st[1]>>>(flag[0], flag[1], buf[0], payload, n, 0, mode, clk[1]);
            CK(cudaSetDevice(0)); CK(cudaEventSynchronize(e1)); CK(cudaSetDevice(1)); CK(cudaDeviceSynchronize());
            float ms = 0; CK(cudaEventElapsedTime(&ms, e0, e1));
            printf("payload %6d doubles  poll=%s%s signal=%s : %.2f us per round trip (%.2f us one way)\n", payload, (mode & 1) ? "relaxed+fence" : "acquire", (mode & 2) ? "+nanosleep" : "",
                   (mode & 4) ? "fence.sys+relaxed st" : "st.release.sys", ms * 1e3 / n, ms * 1e3 / n / 2);
        }
    return 0;
}
