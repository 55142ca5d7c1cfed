// Sparse occupancy outputs: the step right after the hot path.
//
// The voxeliser accumulates the occupancy in a bit-packed mask (4 class bits per voxel, 1 MB for 256x256x32); the dense fp32
// grid the reference returns (B x 100 MB) is only an expansion of it.  This file turns the mask -- or a dense grid produced by
// anybody -- into the list the reference's `occupancy_grid_to_points` builds on the CPU (SOccDPT/utils/__init__.py:532-568):
//     rows (x, y, z, class) as float64, x = float32(i / G0 * occupancy_shape[0]) ..., all cells with grid >= 0.5,
//     ordered by class, then by (i, j, k) -- the order of np.argwhere filtered per class.
//   grid_pack_kernel     dense fp32 (G0,G1,G2,C) -> mask (threshold >= 0.5, like the reference's `occupancy_grid >= 0.5`)
//   mask_count_kernel    per 256-word block and class: number of set cells
//   mask_scan_kernel     exclusive scan of the block counts in (class, block) order (one CTA; 4 x 1024 entries for 256x256x32)
//   mask_emit_kernel     order-preserving compaction: block-local prefix per class + the scanned block offset
#include "common.cuh"

namespace {

constexpr int kThreads = 256;

template <int C>
__global__ void __launch_bounds__(kThreads)
grid_pack_kernel(const float *__restrict__ grid, unsigned *__restrict__ mask, long long nvox) {
    const long long words = (nvox + 7) / 8;
    for (long long w = (long long)blockIdx.x * kThreads + threadIdx.x; w < words; w += (long long)gridDim.x * kThreads) {
        unsigned m = 0u;
#pragma unroll
        for (int v = 0; v < 8; ++v) {
            const long long vox = w * 8 + v;
            if (vox < nvox) {
#pragma unroll
                for (int c = 0; c < C; ++c) m |= (__ldg(grid + vox * C + c) >= 0.5f) ? (1u << (v * 4 + c)) : 0u;
            }
        }
        mask[w] = m;
    }
}

__device__ __forceinline__ unsigned class_bits(unsigned word, int c) { return word & (0x11111111u << c); }

// counts[c * nblocks + blk]
__global__ void __launch_bounds__(kThreads)
mask_count_kernel(const unsigned *__restrict__ mask, long long words, int C, unsigned *__restrict__ counts, int nblocks) {
    __shared__ unsigned s[4][kThreads / 32];
    const long long w = (long long)blockIdx.x * kThreads + threadIdx.x;
    const unsigned word = w < words ? mask[w] : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int c = 0; c < C; ++c) {
        const unsigned n = __reduce_add_sync(0xffffffffu, (unsigned)__popc(class_bits(word, c)));
        if (lane == 0) s[c][warp] = n;
    }
    __syncthreads();
    if (threadIdx.x < C) {
        unsigned t = 0;
        for (int i = 0; i < kThreads / 32; ++i) t += s[threadIdx.x][i];
        counts[threadIdx.x * nblocks + blockIdx.x] = t;
    }
}

// in-place exclusive scan of n = C * nblocks counts by ONE CTA of 1024 threads; total -> *count
__global__ void __launch_bounds__(1024)
mask_scan_kernel(unsigned *__restrict__ counts, int n, long long *__restrict__ count) {
    __shared__ unsigned warp_sum[32];
    __shared__ unsigned carry_s;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry_s = 0u;
    __syncthreads();
    for (int base = 0; base < n; base += 1024) {
        const int i = base + threadIdx.x;
        const unsigned v = i < n ? counts[i] : 0u;
        unsigned x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_sum[warp] = x;
        __syncthreads();
        if (warp == 0) {
            unsigned ws = warp_sum[lane], t = ws;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const unsigned y = __shfl_up_sync(0xffffffffu, t, o);
                if (lane >= o) t += y;
            }
            warp_sum[lane] = t - ws;       // exclusive prefix of the warp totals
        }
        __syncthreads();
        const unsigned carry = carry_s;
        if (i < n) counts[i] = carry + warp_sum[warp] + x - v;
        __syncthreads();
        if (threadIdx.x == 1023) carry_s = carry + warp_sum[31] + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) *count = (long long)carry_s;
}

struct Cell2Point {
    int G0, G1, G2;
    float occ[3];
};

__global__ void __launch_bounds__(kThreads)
mask_emit_kernel(const unsigned *__restrict__ mask, long long words, int C, const unsigned *__restrict__ offsets, int nblocks,
                 const Cell2Point g, double *__restrict__ out, long long cap) {
    __shared__ unsigned warp_tot[4][kThreads / 32];
    const long long w = (long long)blockIdx.x * kThreads + threadIdx.x;
    const unsigned word = w < words ? mask[w] : 0u;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned pre[4] = {0, 0, 0, 0};
    for (int c = 0; c < C; ++c) {
        const unsigned v = (unsigned)__popc(class_bits(word, c));
        unsigned x = v;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const unsigned y = __shfl_up_sync(0xffffffffu, x, o);
            if (lane >= o) x += y;
        }
        if (lane == 31) warp_tot[c][warp] = x;
        pre[c] = x - v;
    }
    __syncthreads();
    if (word == 0u) return;
    for (int c = 0; c < C; ++c) {
        unsigned bits = class_bits(word, c);
        if (bits == 0u) continue;
        unsigned pos = offsets[c * nblocks + blockIdx.x] + pre[c];
        for (int i = 0; i < warp; ++i) pos += warp_tot[c][i];
        while (bits) {
            const int b = __ffs(bits) - 1;
            bits &= bits - 1;
            const long long vox = w * 8 + (b >> 2);
            const int k = (int)(vox % g.G2);
            const long long ij = vox / g.G2;
            const int j = (int)(ij % g.G1), i = (int)(ij / g.G1);
            if ((long long)pos < cap) {
                // class_indices / grid_size * occupancy_shape: int64 / int64 -> float64, * float32 -> float64, .astype(float32)
                double *o = out + (long long)pos * 4;
                o[0] = (double)(float)((double)i / (double)g.G0 * (double)g.occ[0]);
                o[1] = (double)(float)((double)j / (double)g.G1 * (double)g.occ[1]);
                o[2] = (double)(float)((double)k / (double)g.G2 * (double)g.occ[2]);
                o[3] = (double)c;
            }
            ++pos;
        }
    }
}

int geometry_ok(const int grid[3], int C) {
    SOCCDPT_REQUIRE(grid && grid[0] >= 1 && grid[1] >= 1 && grid[2] >= 1, "occupancy: bad grid size");
    SOCCDPT_REQUIRE(C >= 1 && C <= 4, "occupancy: num_classes must be in [1,4] (got %d)", C);
    SOCCDPT_REQUIRE((long long)grid[0] * grid[1] * grid[2] < (1ll << 31), "occupancy: grid too large");
    return SOCCDPT_OK;
}
inline long long words_of(const int grid[3]) { return ((long long)grid[0] * grid[1] * grid[2] + 7) / 8; }
inline int blocks_of(long long words) { return (int)((words + kThreads - 1) / kThreads); }

}  // namespace

extern "C" {

size_t soccdpt_occupancy_mask_bytes(const int grid[3]) { return grid ? (size_t)words_of(grid) * sizeof(unsigned) : 0; }

size_t soccdpt_occupancy_points_workspace_bytes(const int grid[3], int num_classes) {
    if (!grid || num_classes < 1) return 0;
    return (size_t)blocks_of(words_of(grid)) * (size_t)num_classes * sizeof(unsigned) + 16;
}

int soccdpt_grid_pack_fwd(const float *grid_dense, const int grid[3], int num_classes, uint32_t *mask, soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(grid_dense && mask, "grid_pack: NULL pointer");
    int rc = geometry_ok(grid, num_classes);
    if (rc) return rc;
    const long long nvox = (long long)grid[0] * grid[1] * grid[2], words = words_of(grid);
    const int blocks = (int)min((long long)blocks_of(words), (long long)soccdpt::sm_count() * 8);
    cudaStream_t st = soccdpt::as_stream(stream);
    switch (num_classes) {
        case 1: grid_pack_kernel<1><<<blocks, kThreads, 0, st>>>(grid_dense, mask, nvox); break;
        case 2: grid_pack_kernel<2><<<blocks, kThreads, 0, st>>>(grid_dense, mask, nvox); break;
        case 3: grid_pack_kernel<3><<<blocks, kThreads, 0, st>>>(grid_dense, mask, nvox); break;
        default: grid_pack_kernel<4><<<blocks, kThreads, 0, st>>>(grid_dense, mask, nvox); break;
    }
    return soccdpt::check_launch("grid_pack_kernel");
}

int soccdpt_occupancy_points_fwd(const uint32_t *mask, const int grid[3], const float occ_shape[3], int num_classes,
                                 double *points, long long capacity, long long *count, void *workspace, size_t workspace_bytes,
                                 soccdpt_stream_t stream) {
    SOCCDPT_REQUIRE(mask && count && workspace && occ_shape, "occupancy_points: NULL pointer");
    int rc = geometry_ok(grid, num_classes);
    if (rc) return rc;
    SOCCDPT_REQUIRE(workspace_bytes >= soccdpt_occupancy_points_workspace_bytes(grid, num_classes), "occupancy_points: workspace too small");
    SOCCDPT_REQUIRE(points != nullptr || capacity == 0, "occupancy_points: capacity without an output buffer");
    const long long words = words_of(grid);
    const int nblocks = blocks_of(words);
    unsigned *counts = static_cast<unsigned *>(workspace);
    cudaStream_t st = soccdpt::as_stream(stream);
    mask_count_kernel<<<nblocks, kThreads, 0, st>>>(mask, words, num_classes, counts, nblocks);
    rc = soccdpt::check_launch("mask_count_kernel");
    if (rc) return rc;
    mask_scan_kernel<<<1, 1024, 0, st>>>(counts, nblocks * num_classes, count);
    rc = soccdpt::check_launch("mask_scan_kernel");
    if (rc || points == nullptr) return rc;          // count-only call (size the output, then call again)
    const Cell2Point g{grid[0], grid[1], grid[2], {occ_shape[0], occ_shape[1], occ_shape[2]}};
    mask_emit_kernel<<<nblocks, kThreads, 0, st>>>(mask, words, num_classes, counts, nblocks, g, points, capacity);
    return soccdpt::check_launch("mask_emit_kernel");
}
}
