// kernels_state.cu — the rolling state FeatureEngine keeps between snapshots, derived on the device from per-event arrays
// (ref: data_ingestion/src/core/feature_engine.cpp:87-150):
//
//   rolling_buy_volume / rolling_sell_volume  (feature_engine.cpp:95-110)  sums of the sizes of the last ROLLING_WINDOW = 1500 trades,
//       kept by the reference as running doubles: += size on arrival (buy when direction > 0, sell when < 0), -= the size of the
//       trade that leaves the 1500-deep deque (from buy when ITS direction > 0, else from sell — a direction-0 trade is subtracted
//       from sell although it was never added).  Trade sizes are integers (market_event.hpp:13, order_engine.cpp:104), so every
//       partial sum is an exactly representable integer and the running doubles equal exact integer window sums:
//           buy(t)  = B[j_t] - B[j_t - 1500]          B[j]  = sum_{i<j} size_i [dir_i > 0]
//           sell(t) = S[j_t] - E[j_t - 1500]          S[j]  = sum_{i<j} size_i [dir_i < 0],  E[j] = sum_{i<j} size_i [dir_i <= 0]
//       with j_t the snapshot's trade cursor.  The prefixes are 64-bit integers (wrapping: only differences are used).
//   adds_since_last_snapshot  (feature_engine.cpp:119-122, :74-75)         the counter of 'A' events, reset at every snapshot:
//           adds(t) = add_cursor[t] - add_cursor[t-1]   (add_cursor = 'A' events of the stream before the snapshot; row 0: from 0)
//   tick directions  (feature_engine.cpp:134-141)                          already derived on the device from the midprice ticks
//       (kernels_window.cu), never staged.
//
// Three launches per stream for the volumes: block totals, a scan of the block totals, block-local prefixes + bases;
// then one thread per row.  Integer arithmetic only until the final conversion.
#include <algorithm>

#include "mrp_device.cuh"

namespace mrp {

constexpr int kVolThreads = 256, kVolPer = 8, kVolBlock = kVolThreads * kVolPer;   // trades per block

struct Vol3 {
    unsigned long long b, s, e;
};
__device__ __forceinline__ Vol3 vol_of(int d, unsigned long long sz) {
    Vol3 v;
    v.b = d > 0 ? sz : 0ull;
    v.s = d < 0 ? sz : 0ull;
    v.e = d <= 0 ? sz : 0ull;
    return v;
}

// block totals of (buy, sell-add, sell-evict) sizes
__global__ void __launch_bounds__(kVolThreads)
k_vol_totals(const int8_t* __restrict__ dir, const int32_t* __restrict__ size, const int64_t T, unsigned long long* __restrict__ part) {
    __shared__ unsigned long long sh[3][kVolThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t j0 = (int64_t)blockIdx.x * kVolBlock + (int64_t)tid * kVolPer;
    unsigned long long b = 0, s = 0, e = 0;
#pragma unroll
    for (int i = 0; i < kVolPer; ++i)
        if (j0 + i < T) {
            const Vol3 v = vol_of(dir[j0 + i], (unsigned long long)(long long)size[j0 + i]);
            b += v.b; s += v.s; e += v.e;
        }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        b += __shfl_xor_sync(0xffffffffu, b, d);
        s += __shfl_xor_sync(0xffffffffu, s, d);
        e += __shfl_xor_sync(0xffffffffu, e, d);
    }
    if (lane == 0) { sh[0][warp] = b; sh[1][warp] = s; sh[2][warp] = e; }
    __syncthreads();
    if (tid < 3) {
        unsigned long long t = 0;
        for (int w = 0; w < kVolThreads / 32; ++w) t += sh[tid][w];
        part[(int64_t)blockIdx.x * 3 + tid] = t;
    }
}

// exclusive scan of the block totals (one CTA; the table is T / 2048 entries long)
__global__ void __launch_bounds__(1024)
k_vol_scan(unsigned long long* __restrict__ part, const int nb) {
    __shared__ unsigned long long sw[3][32];
    __shared__ unsigned long long carry[3];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid < 3) carry[tid] = 0ull;
    __syncthreads();
    for (int base = 0; base < nb; base += 1024) {
        const int i = base + tid;
        unsigned long long v[3] = {0ull, 0ull, 0ull}, own[3];
        if (i < nb) { v[0] = part[(int64_t)i * 3]; v[1] = part[(int64_t)i * 3 + 1]; v[2] = part[(int64_t)i * 3 + 2]; }
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            own[q] = v[q];
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const unsigned long long u = __shfl_up_sync(0xffffffffu, v[q], d);
                if (lane >= d) v[q] += u;
            }
            if (lane == 31) sw[q][warp] = v[q];
        }
        __syncthreads();
        unsigned long long off[3];
#pragma unroll
        for (int q = 0; q < 3; ++q) {
            unsigned long long o = carry[q];
            for (int w = 0; w < warp; ++w) o += sw[q][w];
            off[q] = o;
        }
        if (i < nb) {
#pragma unroll
            for (int q = 0; q < 3; ++q) part[(int64_t)i * 3 + q] = off[q] + v[q] - own[q];   // exclusive
        }
        __syncthreads();
        if (tid == 1023) {
#pragma unroll
            for (int q = 0; q < 3; ++q) carry[q] = off[q] + v[q];
        }
        __syncthreads();
    }
}

// P[j] = exclusive prefix at trade j (P has T + 1 entries per quantity, quantity-major)
__global__ void __launch_bounds__(kVolThreads)
k_vol_prefix(const int8_t* __restrict__ dir, const int32_t* __restrict__ size, const int64_t T, const unsigned long long* __restrict__ part,
             unsigned long long* __restrict__ P) {
    __shared__ unsigned long long sw[3][kVolThreads / 32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int64_t j0 = (int64_t)blockIdx.x * kVolBlock + (int64_t)tid * kVolPer;
    Vol3 loc[kVolPer];
    unsigned long long b = 0, s = 0, e = 0;
#pragma unroll
    for (int i = 0; i < kVolPer; ++i) {
        Vol3 v{0ull, 0ull, 0ull};
        if (j0 + i < T) v = vol_of(dir[j0 + i], (unsigned long long)(long long)size[j0 + i]);
        loc[i].b = b; loc[i].s = s; loc[i].e = e;   // exclusive within the thread
        b += v.b; s += v.s; e += v.e;
    }
    unsigned long long ib = b, is = s, ie = e;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        const unsigned long long ub = __shfl_up_sync(0xffffffffu, ib, d), us = __shfl_up_sync(0xffffffffu, is, d), ue = __shfl_up_sync(0xffffffffu, ie, d);
        if (lane >= d) { ib += ub; is += us; ie += ue; }
    }
    if (lane == 31) { sw[0][warp] = ib; sw[1][warp] = is; sw[2][warp] = ie; }
    __syncthreads();
    unsigned long long ob = part[(int64_t)blockIdx.x * 3] + ib - b, os = part[(int64_t)blockIdx.x * 3 + 1] + is - s,
                       oe = part[(int64_t)blockIdx.x * 3 + 2] + ie - e;
    for (int w = 0; w < warp; ++w) { ob += sw[0][w]; os += sw[1][w]; oe += sw[2][w]; }
    const int64_t stride = T + 1;
#pragma unroll
    for (int i = 0; i < kVolPer; ++i)
        if (j0 + i < T) {
            P[j0 + i] = ob + loc[i].b;
            P[stride + j0 + i] = os + loc[i].s;
            P[2 * stride + j0 + i] = oe + loc[i].e;
        }
    if (j0 <= T - 1 && T - 1 < j0 + kVolPer) {   // the thread that owns the last trade also writes P[T]
        P[T] = ob + b;
        P[stride + T] = os + s;
        P[2 * stride + T] = oe + e;
    }
}

// one thread per row of the stream: the two window sums (and the adds counter)
__global__ void __launch_bounds__(256)
k_state_rows(const unsigned long long* __restrict__ P, const int64_t T, const uint32_t* __restrict__ jcur, const uint32_t* __restrict__ add_cursor,
             const int64_t n, const int window, double* __restrict__ buy, double* __restrict__ sell, int32_t* __restrict__ adds) {
    const int64_t stride = T + 1;
    for (int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; r < n; r += (int64_t)gridDim.x * blockDim.x) {
        if (P) {
            int64_t j = jcur[r];
            if (j > T) j = T;
            const int64_t lo = j > window ? j - window : 0;
            buy[r] = (double)(long long)(P[j] - P[lo]);                       // exact: integer window sums (feature_engine.cpp:95-110)
            sell[r] = (double)(long long)(P[stride + j] - P[2 * stride + lo]);
        }
        if (add_cursor) adds[r] = (int32_t)(add_cursor[r] - (r > 0 ? add_cursor[r - 1] : 0u));   // feature_engine.cpp:74-75,119-122
    }
}

size_t state_prefix_bytes(int64_t T) { return (size_t)(T + 1) * 3 * sizeof(unsigned long long); }
size_t state_partial_bytes(int64_t T) { return (size_t)((T + kVolBlock - 1) / kVolBlock + 1) * 3 * sizeof(unsigned long long); }

// One stream: trades [0, T) (dir, size), rows [0, n) (trade cursors, optional add cursors) -> buy / sell / adds of those rows.
cudaError_t launch_build_state(const int8_t* dir, const int32_t* size, int64_t T, const uint32_t* jcur, const uint32_t* add_cursor, int64_t n,
                               int window, unsigned long long* part, unsigned long long* P, double* buy, double* sell, int32_t* adds,
                               cudaStream_t st) {
    const bool vols = size != nullptr;
    if (vols) {
        const int nb = (int)((T + kVolBlock - 1) / kVolBlock);
        if (nb > 0) {
            k_vol_totals<<<nb, kVolThreads, 0, st>>>(dir, size, T, part);
            k_vol_scan<<<1, 1024, 0, st>>>(part, nb);
            k_vol_prefix<<<nb, kVolThreads, 0, st>>>(dir, size, T, part, P);
        } else {
            cudaError_t e = cudaMemsetAsync(P, 0, 3 * sizeof(unsigned long long), st);   // no trades: P[0] = 0 for the three quantities
            if (e != cudaSuccess) return e;
        }
    }
    if (n > 0) {
        const unsigned g = (unsigned)std::min<int64_t>((n + 255) / 256, 148 * 8);
        k_state_rows<<<g, 256, 0, st>>>(vols ? P : nullptr, T, jcur, add_cursor, n, window, buy, sell, adds);
    }
    return cudaGetLastError();
}

}  // namespace mrp
