// Dependent-issue latency and per-SMSP throughput of the fp64 instructions the feature kernels are made of (B200, sm_100a).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
template <int ILP>
__global__ void k_dfma(double* out, long long* cyc, double a, double b, int iters) {
    double v[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) v[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int r = 0; r < 8; ++r)
#pragma unroll
            for (int i = 0; i < ILP; ++i) v[i] = fma(v[i], b, a);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += v[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_lds(double* out, long long* cyc, int iters) {
    __shared__ int nxt[1024];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) nxt[i] = (i + 33) & 1023;
    __syncthreads();
    int p = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters * 8; ++it) p = nxt[p];
    long long t1 = clock64();
    out[threadIdx.x] = p;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
__global__ void k_shfl(double* out, long long* cyc, int iters) {
    double v = threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters * 8; ++it) v = __shfl_up_sync(0xffffffffu, v, 1) + 1.0;
    long long t1 = clock64();
    out[threadIdx.x] = v;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 20); cudaMalloc(&cyc, 4096);
    long long h;
    const int iters = 2000;
#define RUN(ILP, thr) k_dfma<ILP><<<1, thr>>>(out, cyc, 1.0, 0.999, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
    printf("DFMA ILP=%d threads=%d: %.2f cycles per dependent step, %.2f cycles per warp-instruction per SMSP\n", ILP, thr, (double)h / (iters * 8), (double)h / (iters * 8 * ILP * ((thr + 127) / 128)));
    RUN(1, 32) RUN(2, 32) RUN(4, 32) RUN(8, 32) RUN(1, 128) RUN(4, 128) RUN(8, 128) RUN(4, 512) RUN(8, 512)
    k_lds<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("LDS dependent: %.2f cycles\n", (double)h / (iters * 8));
    k_shfl<<<1, 32>>>(out, cyc, iters); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); printf("SHFL.64 + DADD dependent: %.2f cycles\n", (double)h / (iters * 8));
    return 0;
}
