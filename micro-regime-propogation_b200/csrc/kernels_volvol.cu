// kernels_volvol.cu — rolling mean / variance of ONE column over several row horizons ("vol-of-vol").
//
// EXTENSION WITHOUT A REFERENCE COUNTERPART (BASELINE.json configs[3]): the reference has no
// vol-of-vol feature.  Definition used here (include/mrp_fg.h; the test suite's CPU checker restates it with direct window loops):
//     vov_h(t)  = population standard deviation of `ewm_volatility` over the rows t-n+1 .. t of the same stream,
//                 n = min(t + 1, h)                                  (0 when the variance is not positive)
//     z_h(t)    = rolling z-score of vov_h over norm_window rows, arithmetic of feature_normalizer.cpp:61-69
//                 (std = 1 when the variance is not positive)
// A non-finite value makes the rows whose window contains it non-finite, plus the rest of the tile in which it leaves the
// window (that tile's sliding sums carry it: at most tile_rows - 1 further rows); the reference's "poisoned for ever" running
// sums are a property of ITS normaliser, not of this extension.
//
// Same sliding-window algebra as the z-score normaliser (kernels_norm.cu): tiles of `tile_rows` rows, tile_rows dividing every
// window;  window sum(t) = [totals of the preceding window / tile_rows tiles] + sum_{i <= t in tile} (x_i - x_{i - window}).
// Two launches instead of a ticket protocol (the column is 8 bytes per row: a second read costs nothing next to the feature
// pass): k_roll_totals publishes the tile totals, k_roll_apply consumes them.  One warp per (tile, job); a lane owns four
// consecutive rows of a 128-row step.
#include "mrp_device.cuh"

namespace mrp {

struct RollTile {
    int64_t row;     // first row of the tile (index into the column)
    int32_t n;       // rows in the tile
    int32_t index;   // tile index inside its stream
};

constexpr int kRollWarps = 4;
constexpr int kRollRPL = 4;
constexpr int kRollStep = 32 * kRollRPL;
enum { MRP_ROLL_STD = 0, MRP_ROLL_Z = 1 };   // output of k_roll_apply: the window's standard deviation, or the value's z-score

__global__ void __launch_bounds__(kRollWarps * 32)
k_roll_totals(const double* __restrict__ x_base, const int64_t x_stride, const RollTile* __restrict__ tiles, const int ntiles,
              double* __restrict__ tot) {
    const int lane = threadIdx.x & 31, t = blockIdx.x * kRollWarps + (threadIdx.x >> 5), job = blockIdx.y;
    if (t >= ntiles) return;
    const RollTile T = tiles[t];
    const double* __restrict__ x = x_base + (int64_t)job * x_stride + T.row;
    double s = 0.0, s2 = 0.0;
    for (int i = lane; i < T.n; i += 32) {
        const double v = x[i];
        s += v;
        s2 = fma(v, v, s2);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
        s += __shfl_xor_sync(0xffffffffu, s, d);
        s2 += __shfl_xor_sync(0xffffffffu, s2, d);
    }
    if (lane == 0) {
        double* o = tot + ((int64_t)job * ntiles + t) * 2;
        o[0] = s;
        o[1] = s2;
    }
}

// one tile of one job; LAG = the tile lies a full window into its stream
template <bool LAG>
__device__ __forceinline__ void roll_tile(const double* __restrict__ x, double* __restrict__ out, const int n, const int W, const int64_t tl0,
                                          double run1, double run2, const int mode, const int lane) {
    const double invW = 1.0 / (double)W;
    for (int i0 = 0; i0 < n; i0 += kRollStep) {
        const int ib = i0 + lane * kRollRPL;
        double v[kRollRPL], p1[kRollRPL], p2[kRollRPL];
        double a1 = 0.0, a2 = 0.0;
#pragma unroll
        for (int k = 0; k < kRollRPL; ++k) {
            const bool ok = ib + k < n;
            v[k] = ok ? x[ib + k] : 0.0;
            if (LAG) {
                const double l = ok ? x[ib + k - W] : 0.0;
                a1 += v[k] - l;
                a2 += fma(v[k], v[k], -(l * l));
            } else {
                a1 += v[k];
                a2 = fma(v[k], v[k], a2);
            }
            p1[k] = a1;
            p2[k] = a2;
        }
        double s1 = a1, s2 = a2;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            const double u1 = shfl_up_d(s1, d), u2 = shfl_up_d(s2, d);
            if (lane >= d) { s1 += u1; s2 += u2; }
        }
        // exclusive offsets by shuffle, not by subtraction: a NaN in this lane's rows must not reach the rows before it
        const double x1 = shfl_up_d(s1, 1), x2 = shfl_up_d(s2, 1);
        const double b1 = run1 + (lane ? x1 : 0.0), b2 = run2 + (lane ? x2 : 0.0);
#pragma unroll
        for (int k = 0; k < kRollRPL; ++k) {
            if (ib + k >= n) continue;
            // a tile without lag lies inside the first window of its stream: n = rows so far
            const double inv = LAG ? invW : 1.0 / (double)(tl0 + ib + k + 1);
            const double mean = (b1 + p1[k]) * inv;
            const double var = (b2 + p2[k]) * inv - mean * mean;
            double r;
            if (rsqrt_pos_ok(var)) {
                const double rs = rsqrt_pos(var);
                r = mode == MRP_ROLL_STD ? var * rs : (v[k] - mean) * rs;
            } else if (mode == MRP_ROLL_STD) {
                r = var > 0.0 ? sqrt(var) : (var == var ? 0.0 : var);
            } else {
                r = (v[k] - mean) / (var > 0.0 ? sqrt(var) : 1.0);
            }
            out[ib + k] = r;
        }
        run1 += __shfl_sync(0xffffffffu, s1, 31);
        run2 += __shfl_sync(0xffffffffu, s2, 31);
    }
}

__global__ void __launch_bounds__(kRollWarps * 32)
k_roll_apply(const double* __restrict__ x_base, const int64_t x_stride, double* __restrict__ out_base, const int64_t out_stride,
             const RollTile* __restrict__ tiles, const int ntiles, const double* __restrict__ tot, const int32_t* __restrict__ windows,
             const int tile_rows, const int mode) {
    const int lane = threadIdx.x & 31, t = blockIdx.x * kRollWarps + (threadIdx.x >> 5), job = blockIdx.y;
    if (t >= ntiles) return;
    const RollTile T = tiles[t];
    const int W = windows[job];
    const int tb = W / tile_rows;                         // tiles per window
    const int nb = T.index < tb ? T.index : tb;           // preceding tiles of this stream inside the window
    const double* __restrict__ x = x_base + (int64_t)job * x_stride + T.row;
    double* __restrict__ out = out_base + (int64_t)job * out_stride + T.row;
    const int64_t tl0 = (int64_t)T.index * tile_rows;     // stream-local index of the tile's first row
    double h1 = 0.0, h2 = 0.0;
    {
        const double* __restrict__ tp = tot + ((int64_t)job * ntiles + (t - nb)) * 2;
        for (int j = lane; j < nb; j += 32) {
            h1 += tp[(int64_t)j * 2];
            h2 += tp[(int64_t)j * 2 + 1];
        }
#pragma unroll
        for (int d = 16; d > 0; d >>= 1) {
            h1 += __shfl_xor_sync(0xffffffffu, h1, d);
            h2 += __shfl_xor_sync(0xffffffffu, h2, d);
        }
    }
    if (T.index >= tb) roll_tile<true>(x, out, T.n, W, tl0, h1, h2, mode, lane);    // rows W back exist (and belong to the same stream)
    else roll_tile<false>(x, out, T.n, W, tl0, h1, h2, mode, lane);
}

// x: njobs columns (x_stride apart; 0 = every job reads the same column), out: njobs columns, tot: scratch [njobs][ntiles][2]
cudaError_t launch_rolling(const double* x, int64_t x_stride, double* out, int64_t out_stride, const void* tiles, int ntiles, double* tot,
                           const int32_t* windows_dev, int njobs, int tile_rows, int mode, cudaStream_t st) {
    if (ntiles == 0 || njobs == 0) return cudaSuccess;
    const dim3 grid((unsigned)((ntiles + kRollWarps - 1) / kRollWarps), (unsigned)njobs);
    k_roll_totals<<<grid, kRollWarps * 32, 0, st>>>(x, x_stride, static_cast<const RollTile*>(tiles), ntiles, tot);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return e;
    k_roll_apply<<<grid, kRollWarps * 32, 0, st>>>(x, x_stride, out, out_stride, static_cast<const RollTile*>(tiles), ntiles, tot, windows_dev,
                                                   tile_rows, mode);
    return cudaGetLastError();
}

}  // namespace mrp
