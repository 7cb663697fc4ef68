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
constexpr int kStep = 32 * kRowsPerLane;

__device__ __forceinline__ int ld_flag(const int* p) { return *reinterpret_cast<const volatile int*>(p); }

// four consecutive doubles of a lane: two 128-bit loads when the column is 16-byte aligned
__device__ __forceinline__ void load4(const double* __restrict__ p, int i, int n, bool vec, double (&o)[kRowsPerLane]) {
    if (vec && i + kRowsPerLane <= n) {
        const double2 a = *reinterpret_cast<const double2*>(p + i), b = *reinterpret_cast<const double2*>(p + i + 2);
        o[0] = a.x; o[1] = a.y; o[2] = b.x; o[3] = b.y;
    } else {
#pragma unroll
        for (int k = 0; k < kRowsPerLane; ++k) o[k] = i + k < n ? p[i + k] : 0.0;
    }
}

__global__ void __launch_bounds__(kNormThreads, 2)
k_normalize(const DevOutputs out, const int64_t N, const DevTile* __restrict__ tiles, double* __restrict__ tile_tot,
            int* __restrict__ flags, unsigned int* __restrict__ ticket, const DevParams P) {
    __shared__ int sId;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    if (tid == 0) sId = (int)atomicAdd(ticket, 1u);
    __syncthreads();
    const int id = sId;
    const DevTile T = tiles[id];
    const int n = T.n;
    // After this point the ten warps of the CTA never meet again: every warp owns whole feature columns.

    // ---- pass 1: column totals of this tile, published (release) before anything is awaited.
    //      flags[id] counts the columns published so far; a tile is complete at 19.
    for (int c = warp; c < kNF; c += kNormWarps) {
        const double* __restrict__ x = out.raw + (int64_t)c * N + T.row;
        const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0);
        double s = 0.0, s2 = 0.0, t1 = 0.0, t2 = 0.0;
        for (int i0 = 0; i0 < n; i0 += 2 * kStep) {
            double a[kRowsPerLane], b[kRowsPerLane];
            load4(x, i0 + lane * kRowsPerLane, n, vec, a);
            load4(x, i0 + kStep + lane * kRowsPerLane, n, vec, b);
#pragma unroll
            for (int k = 0; k < kRowsPerLane; ++k) {
                s += a[k];
                s2 += a[k] * a[k];
                t1 += b[k];
                t2 += b[k] * b[k];
            }
        }
        s += t1;
        s2 += t2;
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            s += __shfl_xor_sync(0xffffffffu, s, d);
            s2 += __shfl_xor_sync(0xffffffffu, s2, d);
        }
        if (lane == 0) {
            tile_tot[(int64_t)id * 2 * kNF + c] = s;
            tile_tot[(int64_t)id * 2 * kNF + kNF + c] = s2;
            __threadfence();
            atomicAdd(flags + id, 1);
        }
    }

    // ---- pass 2: the window sum slides along the tile:
    //        sum(t) = [totals of the previous tiles] + sum_{i<=t in tile} (x_i - x_{i-W})
    //      so one in-tile scan of (x - x_lag, x^2 - x_lag^2) gives every row's window sums.
    const int nb = T.index < P.tiles_back ? T.index : P.tiles_back;  // preceding tiles inside the window
    const bool has_lag = T.index >= P.tiles_back;
    const int64_t tl0 = (int64_t)T.index * P.tile_rows;  // stream-local index of the tile's first row
    const double2 invW = __ldg(P.inv_n + P.norm_window);
    for (int c = warp; c < kNF; c += kNormWarps) {
        const double* __restrict__ x = out.raw + (int64_t)c * N + T.row;
        const double* __restrict__ xl = x - P.norm_window;
        double* __restrict__ z = out.norm + (int64_t)c * N + T.row;
        const bool vec = ((reinterpret_cast<uintptr_t>(x) & 15) == 0) && ((P.norm_window & 1) == 0) &&
                         ((reinterpret_cast<uintptr_t>(z) & 15) == 0);
        // first step's loads are in flight while the predecessors' totals are awaited
        double v[kRowsPerLane], lv[kRowsPerLane];
        load4(x, lane * kRowsPerLane, n, vec, v);
        if (has_lag) load4(xl, lane * kRowsPerLane, n, vec, lv);
        else {
#pragma unroll
            for (int k = 0; k < kRowsPerLane; ++k) lv[k] = 0.0;
        }
        // totals of the (up to tiles_back) preceding tiles, one lane per predecessor, fixed reduction tree
        double run1 = 0.0, run2 = 0.0;
        for (int j0 = 0; j0 < nb; j0 += 32) {
            const int j = id - nb + j0 + lane;
            double h1 = 0.0, h2 = 0.0;
            if (j0 + lane < nb) {
                while (ld_flag(flags + j) < kNF) { __nanosleep(32); }
                __threadfence();
                h1 = __ldcg(tile_tot + (int64_t)j * 2 * kNF + c);
                h2 = __ldcg(tile_tot + (int64_t)j * 2 * kNF + kNF + c);
            }
#pragma unroll
            for (int d = 16; d > 0; d >>= 1) {
                h1 += __shfl_xor_sync(0xffffffffu, h1, d);
                h2 += __shfl_xor_sync(0xffffffffu, h2, d);
            }
            run1 += h1;
            run2 += h2;
        }
        for (int i0 = 0; i0 < n; i0 += kStep) {
            const int ib = i0 + lane * kRowsPerLane;
            // software pipeline: next step's values are requested before this step's arithmetic
            double vn[kRowsPerLane], ln[kRowsPerLane];
            const bool more = i0 + kStep < n;
            if (more) {
                load4(x, ib + kStep, n, vec, vn);
                if (has_lag) load4(xl, ib + kStep, n, vec, ln);
            }
            double d1[kRowsPerLane], d2[kRowsPerLane];
            double a1 = 0.0, a2 = 0.0;
#pragma unroll
            for (int k = 0; k < kRowsPerLane; ++k) {
                d1[k] = v[k] - lv[k];
                d2[k] = v[k] * v[k] - lv[k] * lv[k];
                a1 += d1[k];
                a2 += d2[k];
            }
            double s1 = a1, s2 = a2;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double u1 = shfl_up_d(s1, d), u2 = shfl_up_d(s2, d);
                if (lane >= d) { s1 += u1; s2 += u2; }
            }
            double e1 = run1 + (s1 - a1), e2 = run2 + (s2 - a2);
            double zz[kRowsPerLane];
#pragma unroll
            for (int k = 0; k < kRowsPerLane; ++k) {
                e1 += d1[k];
                e2 += d2[k];
                const int64_t tl = tl0 + ib + k;
                const bool fullw = tl + 1 >= P.norm_window;
                const int nwin = fullw ? P.norm_window : (int)(tl + 1);  // window.size(), feature_normalizer.cpp:71
                // sum / n as sum * (1/n) with 1/n in double-double (host table): rounds like the division
                const double2 inv = (fullw || ib + k >= n) ? invW : __ldg(P.inv_n + nwin);
                const double m0 = e1 * inv.x;
                const double mean = m0 + (fma(e1, inv.x, -m0) + e1 * inv.y);      // :62
                const double q0 = e2 * inv.x;
                const double ex2 = q0 + (fma(e2, inv.x, -q0) + e2 * inv.y);
                const double variance = ex2 - (mean * mean);                    // :63
                const double rstd = (variance > 0.0) ? rsqrt(variance) : 1.0;   // :67 — 1/sqrt, then multiply
                zz[k] = (v[k] - mean) * rstd;                                   // :68
                if (c == 0 && out.counts && ib + k < n) out.counts[(int64_t)MRP_C_NORM_N * N + T.row + ib + k] = nwin;
            }
            if (vec && ib + kRowsPerLane <= n) {
                *reinterpret_cast<double2*>(z + ib) = make_double2(zz[0], zz[1]);
                *reinterpret_cast<double2*>(z + ib + 2) = make_double2(zz[2], zz[3]);
            } else {
#pragma unroll
                for (int k = 0; k < kRowsPerLane; ++k)
                    if (ib + k < n) z[ib + k] = zz[k];
            }
            run1 += __shfl_sync(0xffffffffu, s1, 31);
            run2 += __shfl_sync(0xffffffffu, s2, 31);
            if (more) {
#pragma unroll
                for (int k = 0; k < kRowsPerLane; ++k) {
                    v[k] = vn[k];
                    lv[k] = has_lag ? ln[k] : 0.0;
                }
            }
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
