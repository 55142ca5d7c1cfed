// Counting voxeliser: the reference's ground-truth generator OccupancyProcessor.transform_points_to_occupancy_grid_vect
// (SOccDPT/datasets/bdd_helper.py:289-362) on the device -- points with INTEGER class ids are counted per (voxel, class)
// (np.add.at), then thresholded: a cell is listed as a point when count >= threshold and set in the boolean grid when
// count > threshold (the reference's asymmetry, :340 vs :357).  BASELINE.json's north_star also names a per-voxel semantic
// argmax: label = 1 + argmax_c count (first maximum), 0 for empty cells.
//
//   voxel_count_kernel<F64>   one point per thread; voxel index with numpy's promotion rules restated exactly
//                             (float64 points: p / double(occ_f32) * G in fp64; float32 points: fp32 IEEE division, the product
//                             in fp64; truncation towards zero; strict 0 < ijk < G); warp-aggregated atomicAdd: lanes that hit
//                             the same (voxel, class) are found with __match_any_sync, the lowest adds the population count.
//                             Neighbouring pixels of a depth map fall into the same cell most of the time, so a warp issues a
//                             handful of atomics instead of 32.
//   voxel_count_finish_kernel one voxel per thread: bool grid (count > thr), bit-packed mask of count >= thr (the layout
//                             soccdpt_occupancy_points_fwd consumes: the points list), argmax labels.
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

struct CountGeom {
    int G0, G1, G2, C;
    float occ[3];
};

__device__ __forceinline__ long long trunc_ll(double v) { return (long long)v; }   // cvt.rzi.s64.f64 (saturating)

template <bool F64>
__global__ void __launch_bounds__(kThreads)
voxel_count_kernel(const void *__restrict__ points, const long long *__restrict__ sem64, const int *__restrict__ sem32,
                   long long n, CountGeom g, int *__restrict__ counts, unsigned long long *__restrict__ bad_class) {
    const long long stride = (long long)gridDim.x * kThreads;
    const long long n_round = (n + 31) / 32 * 32;               // whole warps stay in the loop (match_any needs them)
    for (long long i = (long long)blockIdx.x * kThreads + threadIdx.x; i < n_round; i += stride) {
        long long key = -1;
        if (i < n) {
            double x, y, z;
            bool finite;
            long long ii, jj, kk;
            if (F64) {
                const double *p = static_cast<const double *>(points) + i * 3;
                x = p[0]; y = p[1]; z = p[2];
                finite = isfinite(x) && isfinite(y) && isfinite(z);
                ii = trunc_ll(__dmul_rn(__ddiv_rn(x, (double)g.occ[0]), (double)g.G0));
                jj = trunc_ll(__dmul_rn(__ddiv_rn(y, (double)g.occ[1]), (double)g.G1));
                kk = trunc_ll(__dmul_rn(__ddiv_rn(z, (double)g.occ[2]), (double)g.G2));
            } else {
                const float *p = static_cast<const float *>(points) + i * 3;
                const float xf = p[0], yf = p[1], zf = p[2];
                finite = isfinite(xf) && isfinite(yf) && isfinite(zf);
                ii = trunc_ll(__dmul_rn((double)__fdiv_rn(xf, g.occ[0]), (double)g.G0));
                jj = trunc_ll(__dmul_rn((double)__fdiv_rn(yf, g.occ[1]), (double)g.G1));
                kk = trunc_ll(__dmul_rn((double)__fdiv_rn(zf, g.occ[2]), (double)g.G2));
            }
            if (finite && 0 < ii && ii < g.G0 && 0 < jj && jj < g.G1 && 0 < kk && kk < g.G2) {
                long long c = sem64 ? sem64[i] : (long long)sem32[i];
                if (c < 0) c += g.C;                              // numpy: negative ids index from the end
                if (c >= 0 && c < g.C) key = ((ii * g.G1 + jj) * g.G2 + kk) * g.C + c;
                else atomicAdd(bad_class, 1ull);                  // numpy raises IndexError; reported by the host wrapper
            }
        }
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (key >= 0 && (int)(threadIdx.x & 31) == __ffs(peers) - 1) atomicAdd(counts + key, __popc(peers));
    }
}

__global__ void __launch_bounds__(kThreads)
voxel_count_finish_kernel(const int *__restrict__ counts, long long nvox, int C, float threshold, unsigned char *__restrict__ grid_gt,
                          unsigned *__restrict__ mask_ge, unsigned char *__restrict__ labels) {
    const long long words = (nvox + 7) / 8;
    for (long long w = (long long)blockIdx.x * kThreads + threadIdx.x; w < words; w += (long long)gridDim.x * kThreads) {
        unsigned m = 0u;
        for (int v = 0; v < 8; ++v) {
            const long long vox = w * 8 + v;
            if (vox >= nvox) break;
            int best = 0, best_c = 0, total = 0;
            for (int c = 0; c < C; ++c) {
                const int n = counts[vox * C + c];
                const float f = (float)n;                        // the reference counts in a float32 grid (exact below 2^24)
                if (grid_gt) grid_gt[vox * C + c] = f > threshold ? 1 : 0;
                if (f >= threshold) m |= 1u << (v * 4 + c);
                if (n > best) { best = n; best_c = c; }
                total += n;
            }
            if (labels) labels[vox] = total > 0 ? (unsigned char)(1 + best_c) : 0;
        }
        if (mask_ge) mask_ge[w] = m;
    }
}

unsigned grid_for(long long items) {
    long long b = (items + kThreads - 1) / kThreads;
    const long long cap = (long long)soccdpt::sm_count() * 16;
    return (unsigned)(b < 1 ? 1 : (b > cap ? cap : b));
}

}  // namespace

extern "C" {

int soccdpt_voxel_count_fwd(const void *points, int points_f64, const void *semantics, int semantics_i64, long long n,
                            const int grid[3], const float occ_shape[3], int num_classes, int32_t *counts,
                            unsigned long long *bad_class, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(grid && occ_shape && counts && bad_class, "voxel_count: NULL pointer");
    SOCCDPT_REQUIRE(n >= 0 && (n == 0 || (points && semantics)), "voxel_count: NULL points / semantics");
    SOCCDPT_REQUIRE(num_classes >= 1 && num_classes <= 4, "voxel_count: num_classes must be 1..4 (got %d)", num_classes);
    SOCCDPT_REQUIRE(grid[0] >= 1 && grid[1] >= 1 && grid[2] >= 1 &&
                    (long long)grid[0] * grid[1] * grid[2] * num_classes < (1ll << 31), "voxel_count: bad grid");
    if (n == 0) return SOCCDPT_OK;
    CountGeom g{grid[0], grid[1], grid[2], num_classes, {occ_shape[0], occ_shape[1], occ_shape[2]}};
    const long long *s64 = semantics_i64 ? static_cast<const long long *>(semantics) : nullptr;
    const int *s32 = semantics_i64 ? nullptr : static_cast<const int *>(semantics);
    cudaStream_t st = soccdpt::as_stream(stream);
    if (points_f64)
        voxel_count_kernel<true><<<grid_for(n), kThreads, 0, st>>>(points, s64, s32, n, g, counts, bad_class);
    else
        voxel_count_kernel<false><<<grid_for(n), kThreads, 0, st>>>(points, s64, s32, n, g, counts, bad_class);
    return soccdpt::check_launch("voxel_count_kernel");
}

int soccdpt_voxel_count_finish_fwd(const int32_t *counts, const int grid[3], int num_classes, float threshold,
                                   uint8_t *grid_gt, uint32_t *mask_ge, uint8_t *labels, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(counts && grid, "voxel_count_finish: NULL pointer");
    SOCCDPT_REQUIRE(num_classes >= 1 && num_classes <= 4, "voxel_count_finish: num_classes must be 1..4 (got %d)", num_classes);
    const long long nvox = (long long)grid[0] * grid[1] * grid[2];
    SOCCDPT_REQUIRE(nvox >= 1 && nvox * num_classes < (1ll << 31), "voxel_count_finish: bad grid");
    voxel_count_finish_kernel<<<grid_for((nvox + 7) / 8), kThreads, 0, soccdpt::as_stream(stream)>>>(
        counts, nvox, num_classes, threshold, grid_gt, mask_ge, labels);
    return soccdpt::check_launch("voxel_count_finish_kernel");
}

}  // extern "C"
