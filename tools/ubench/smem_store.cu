// Shared-memory store / load throughput of the access patterns k_windows uses (B200, sm_100a).
// nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o smem_store smem_store.cu && ./smem_store
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kRun = 11;
// mode 0: 5 x STS.64 + 1 x STS.32 per tick into six arrays (the ring layout), lane stride kRun
// mode 1: 3 x STS.128 per tick into three arrays of 16-byte pairs
// mode 2: LDS.64 with the same addressing (5 per tick)
template <int MODE>
__global__ void k(double* out, long long* cyc, int iters, int R) {
    extern __shared__ __align__(16) unsigned char dyn[];
    double* QA = reinterpret_cast<double*>(dyn);
    double* QS = QA + R; double* Q1 = QS + R; double* Q2 = Q1 + R; double* QG = Q2 + R;
    unsigned* QC = reinterpret_cast<unsigned*>(QG + R);
    double2* PA = reinterpret_cast<double2*>(dyn); double2* PB = PA + R; double2* PC = PB + R;
    const int tid = threadIdx.x;
    double a = tid, acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        const int slot = (it * 37 + kRun * tid) % (R - kRun);
#pragma unroll
        for (int i = 0; i < kRun; ++i) {
            a += 1.0;
            if (MODE == 0) { QA[slot + i] = a; QS[slot + i] = a; Q1[slot + i] = a; Q2[slot + i] = a; QG[slot + i] = a; QC[slot + i] = (unsigned)it; }
            if (MODE == 1) { PA[slot + i] = make_double2(a, a); PB[slot + i] = make_double2(a, a); PC[slot + i] = make_double2(a, a); }
            if (MODE == 2) { acc += QA[slot + i] + QS[slot + i] + Q1[slot + i] + Q2[slot + i] + QG[slot + i]; }
        }
        __syncthreads();
    }
    long long t1 = clock64();
    out[blockIdx.x * blockDim.x + tid] = a + acc + QA[tid];
    if (tid == 0) cyc[blockIdx.x] = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1 << 22); cudaMalloc(&cyc, 1 << 16);
    const int R = 3300, iters = 2000;
    const int smem = R * 48 + 64;
    cudaFuncSetAttribute(k<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(k<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    long long h;
    for (int thr : {32, 128, 256}) {
        k<0><<<148, thr, smem>>>(out, cyc, iters, R); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads %3d  5xSTS.64+STS.32 per tick: %.0f cycles per %d ticks  = %.1f bytes/cycle/SM\n", thr, (double)h / iters, thr * kRun, (double)thr * kRun * 44 / ((double)h / iters));
        k<1><<<148, thr, smem>>>(out, cyc, iters, R); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads %3d  3xSTS.128 per tick:        %.0f cycles per %d ticks  = %.1f bytes/cycle/SM\n", thr, (double)h / iters, thr * kRun, (double)thr * kRun * 48 / ((double)h / iters));
        k<2><<<148, thr, smem>>>(out, cyc, iters, R); cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("threads %3d  5xLDS.64 per tick:         %.0f cycles per %d ticks  = %.1f bytes/cycle/SM\n", thr, (double)h / iters, thr * kRun, (double)thr * kRun * 40 / ((double)h / iters));
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
