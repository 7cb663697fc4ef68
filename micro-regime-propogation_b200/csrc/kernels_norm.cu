// kernels_norm.cu — the rolling z-score (FeatureNormalizer) and its fused write epilogue.
//
// Reference (feature_normalizer.cpp:32-83): running sum / sum-of-squares per feature, add the new row,
// evict the row 30 000 back, z = (x - sum/n) / sqrt(sum2/n - (sum/n)^2) (std = 1 when the variance is not
// positive).  Those running sums random-walk in rounding error as the stream grows.
//
// k_normalize evaluates the SAME window sums without any long-lived accumulator.  The row axis of a stream
// is cut into tiles of `tile_rows` rows, tile_rows dividing norm_window (30 000 = 30 x 1000).  For a row t
// in tile T (offset o in the stream), with l = t - norm_window in tile T-30:
//      window sum(t) = [totals of tiles T-30 .. T-1] + prefix_T(t) - prefix_{T-30}(l)
// so a CTA needs 30 small totals of its predecessors plus two in-tile scans.  Tiles are handed out by an
// atomic ticket in time order; each CTA first publishes its own totals (release flag), only then waits for
// its predecessors' flags — predecessors hold smaller tickets, hence are already resident: no deadlock,
// no serial carry chain.  The lagged tile was streamed by a CTA ~30 tickets earlier and is still in the
// 126 MB L2, so HBM sees each raw value once (read) and each z-score once (write).
//
// k_normalize_faithful re-enacts the reference's add-then-evict chain in its exact rounding order
// (one thread per stream x feature; serial, for validation only).
#include "mrp_device.cuh"

namespace mrp {

constexpr int kNormWarps = 10;
constexpr int kNormThreads = kNormWarps * 32;
constexpr int kNF = MRP_FG_NUM_FEATURES;
constexpr int kRowsPerLane = 4;

__device__ __forceinline__ int ld_flag(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

__global__ void __launch_bounds__(kNormThreads, 2)
k_normalize(const DevOutputs out, const int64_t N, const DevTile* __restrict__ tiles, double* __restrict__ tile_tot,
            int* __restrict__ flags, unsigned int* __restrict__ ticket, const DevParams P) {
    __shared__ int sId;
    __shared__ double sHead[2 * kNF];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sId = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int id = sId;
    const DevTile T = tiles[id];
    const int n = T.n;

    // ---- pass 1: this tile's column totals, published before anything is awaited
    for (int c = warp; c < kNF; c += kNormWarps) {
        const double* __restrict__ x = out.raw + (int64_t)c * N + T.row;
        double s = 0.0, s2 = 0.0;
        for (int i = lane; i < n; i += 32) {
            const double v = x[i];
            s += v;
            s2 += v * v;
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, d);
            s2 += __shfl_xor_sync(0xffffffffu, s2, d);
        }
        if (lane == 0) {
            tile_tot[(int64_t)id * 2 * kNF + c] = s;
            tile_tot[(int64_t)id * 2 * kNF + kNF + c] = s2;
        }
    }
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        atomicExch(flags + id, 1);
    }

    // ---- totals of the (up to) tiles_back preceding tiles of the same stream, oldest first
    const int nb = T.index < P.tiles_back ? T.index : P.tiles_back;
    if (tid < 2 * kNF) {
        double s = 0.0;
        for (int j = id - nb; j < id; ++j) {
            while (ld_flag(flags + j) == 0) { __nanosleep(40); }
            __threadfence();
            s += __ldcg(tile_tot + (int64_t)j * 2 * kNF + tid);
        }
        sHead[tid] = s;
    }
    __syncthreads();

    // ---- pass 2: one warp per feature column; in-tile scans of the current and the lagged tile
    const bool has_lag = T.index >= P.tiles_back;
    const int64_t tl0 = (int64_t)T.index * P.tile_rows;  // stream-local index of the tile's first row
    for (int c = warp; c < kNF; c += kNormWarps) {
        const double* __restrict__ x = out.raw + (int64_t)c * N + T.row;
        const double* __restrict__ xl = x - P.norm_window;
        double* __restrict__ z = out.norm + (int64_t)c * N + T.row;
        const double H1 = sHead[c], H2 = sHead[kNF + c];
        double run1 = 0.0, run2 = 0.0, lrun1 = 0.0, lrun2 = 0.0;  // warp-uniform running prefixes
        for (int i0 = 0; i0 < n; i0 += 32 * kRowsPerLane) {
            const int ib = i0 + lane * kRowsPerLane;
            double v[kRowsPerLane], p1[kRowsPerLane], p2[kRowsPerLane], q1[kRowsPerLane], q2[kRowsPerLane];
            double a1 = 0.0, a2 = 0.0, b1 = 0.0, b2 = 0.0;
#pragma unroll
            for (int k = 0; k < kRowsPerLane; ++k) {
                const bool in_t = ib + k < n;
                v[k] = in_t ? x[ib + k] : 0.0;
                const double lv = (in_t && has_lag) ? xl[ib + k] : 0.0;
                a1 += v[k];
                a2 += v[k] * v[k];
                b1 += lv;
                b2 += lv * lv;
                p1[k] = a1; p2[k] = a2; q1[k] = b1; q2[k] = b2;
            }
            double s1 = a1, s2 = a2, t1 = b1, t2 = b2;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double u1 = shfl_up_d(s1, d), u2 = shfl_up_d(s2, d), w1 = shfl_up_d(t1, d), w2 = shfl_up_d(t2, d);
                if (lane >= d) { s1 += u1; s2 += u2; t1 += w1; t2 += w2; }
            }
            const double e1 = run1 + (s1 - a1), e2 = run2 + (s2 - a2), f1 = lrun1 + (t1 - b1), f2 = lrun2 + (t2 - b2);
#pragma unroll
            for (int k = 0; k < kRowsPerLane; ++k) {
                if (ib + k < n) {
                    const int64_t tl = tl0 + ib + k;
                    const int64_t nwin = tl + 1 < P.norm_window ? tl + 1 : P.norm_window;  // window.size(), :71
                    const double sum = (H1 + (e1 + p1[k])) - (f1 + q1[k]);
                    const double sum2 = (H2 + (e2 + p2[k])) - (f2 + q2[k]);
                    const double dn = (double)nwin;
                    const double mean = sum / dn;                                   // :62
                    const double variance = (sum2 / dn) - (mean * mean);            // :63
                    const double stddev = (variance > 0.0) ? sqrt(variance) : 1.0;  // :67
                    z[ib + k] = (v[k] - mean) / stddev;                             // :68
                    if (c == 0 && out.counts) out.counts[(int64_t)MRP_C_NORM_N * N + T.row + ib + k] = (int32_t)nwin;
                }
            }
            run1 += __shfl_sync(0xffffffffu, s1, 31);
            run2 += __shfl_sync(0xffffffffu, s2, 31);
            lrun1 += __shfl_sync(0xffffffffu, t1, 31);
            lrun2 += __shfl_sync(0xffffffffu, t2, 31);
        }
    }
}

// One thread per (stream, feature): the reference's own chain, feature_normalizer.cpp:37-51,61-69.
__global__ void k_normalize_faithful(const DevOutputs out, const int64_t N, const int64_t* __restrict__ row_off,
                                     const int n_streams, const DevParams P) {
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;
    if (idx >= n_streams * kNF) return;
    const int s = idx / kNF, c = idx % kNF;
    const int64_t r0 = row_off[s], r1 = row_off[s + 1];
    const double* __restrict__ x = out.raw + (int64_t)c * N;
    double* __restrict__ z = out.norm + (int64_t)c * N;
    const int64_t W = P.norm_window;
    double sum = 0.0, sum2 = 0.0;
    for (int64_t t = r0; t < r1; ++t) {
        const double val = x[t];
        sum += val;
        sum2 += val * val;
        int64_t size = t - r0 + 1;
        if (size > W) {
            const double old = x[t - W];
            sum += -old;
            sum2 += -(old * old);
            size = W;
        }
        const double mean = sum / (double)size;
        const double variance = (sum2 / (double)size) - (mean * mean);
        const double stddev = (variance > 0.0) ? sqrt(variance) : 1.0;
        z[t] = (val - mean) / stddev;
    }
}

cudaError_t launch_normalize(const DevOutputs& out, int64_t N, const DevTile* tiles, int ntiles, double* tile_tot,
                             int* flags, const DevParams& P, cudaStream_t st) {
    // flags[0..ntiles) + the ticket counter at flags[ntiles]
    cudaError_t e = cudaMemsetAsync(flags, 0, sizeof(int) * ((size_t)ntiles + 1), st);
    if (e != cudaSuccess) return e;
    k_normalize<<<ntiles, kNormThreads, 0, st>>>(out, N, tiles, tile_tot, flags,
                                                 reinterpret_cast<unsigned int*>(flags + ntiles), P);
    return cudaGetLastError();
}

cudaError_t launch_normalize_faithful(const DevOutputs& out, int64_t N, const int64_t* row_off, int n_streams,
                                      const DevParams& P, cudaStream_t st) {
    const int total = n_streams * kNF;
    k_normalize_faithful<<<(total + 31) / 32, 32, 0, st>>>(out, N, row_off, n_streams, P);
    return cudaGetLastError();
}

}  // namespace mrp
