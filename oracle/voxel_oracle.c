/*
 * TEST INFRASTRUCTURE ONLY -- CPU restatement (plain C, scalar fp32, no FMA
 * contraction) of the reference's depth->points->occupancy post-processing.
 * Nothing in the product package links or calls this file; only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs do.
 *
 * Follows, line by line (all paths under /root/reference/):
 *   SOccDPT/model/SOccDPT.py:288-293   clamp inv_depth >= 1e-8 in place, 1/x, inf/nan -> +inf
 *   SOccDPT/model/SOccDPT.py:301-316   X=(V-cx)*d/fx, Y=(U-cy)*d/fy, Z=d  (fp32, one rounding per op)
 *   SOccDPT/model/SOccDPT.py:343-353   points viewed (B,N,3); points_3D[:,k] = p*pc_scale[k]+pc_shift[k]
 *                                      indexes the POINT axis -> only pixels 0,1,2 of every frame change
 *   SOccDPT/model/SOccDPT.py:60-130    rotate_points: three chained p @ R (einsum -> bmm)
 *   SOccDPT/model/SOccDPT.py:374-463   points_to_occupancy_grid: finite mask, trunc(p/shape*grid),
 *                                      strict 0<ijk<G, nonzero(semantics), grid[:, i,j,k,c] = 1
 *
 * Pinned against the reference's own outputs: tests/golden/voxel_*.npz (made by
 * oracle/make_golden.py importing the unmodified reference in the build container)
 * and live in tests/test_oracle_vs_reference.py when /root/reference is present.
 *
 * Conventions verified on the reference's CPU path:
 *   - unproject: sub, mul, TRUE division, each rounded to fp32, host scalars cast to fp32;
 *   - bmm with K=3: acc = p0*R[0][j]; acc = fma(p1,R[1][j],acc); acc = fma(p2,R[2][j],acc);
 *   - voxel index: trunc((p / shape) * grid) with both ops rounded to fp32.
 * Build: see oracle/Makefile (-O2 -ffp-contract=off; fmaf is called explicitly).
 */
#include <math.h>
#include <stdint.h>
#include <pthread.h>
#include <string.h>
#include <unistd.h>

static inline void rot3(const float *R, float *p)
{
    float q[3];
    for (int j = 0; j < 3; ++j) {
        float acc = p[0] * R[0 * 3 + j];
        acc = fmaf(p[1], R[1 * 3 + j], acc);
        acc = fmaf(p[2], R[2 * 3 + j], acc);
        q[j] = acc;
    }
    p[0] = q[0]; p[1] = q[1]; p[2] = q[2];
}

typedef struct {
    float *inv_depth; const float *seg; int B, H, W, C;
    float fx, fy, cx, cy;
    const float *pc_scale, *pc_shift, *rot, *occ_shape; const int *gsz;
    float *points, *grid; int per_frame;
    long long begin, end, stores;
} job_t;

static void *worker(void *arg)
{
    job_t *J = (job_t *)arg;
    const int W = J->W, C = J->C; const int *gsz = J->gsz;
    const long long N = (long long)J->H * W;
    const long long cells = (long long)gsz[0] * gsz[1] * gsz[2] * C;
    const float eps = 1e-8f;
    const float g0 = (float)gsz[0], g1 = (float)gsz[1], g2 = (float)gsz[2];
    const float fx = J->fx, fy = J->fy, cx = J->cx, cy = J->cy;
    long long stores = 0;
    for (long long bn = J->begin; bn < J->end; ++bn) {
        const int b = (int)(bn / N);
        const long long n = bn % N;
        const int u = (int)(n / W), v = (int)(n % W);
        float inv = J->inv_depth[bn];
        if (inv < eps) inv = eps;           /* NaN compares false and stays NaN */
        J->inv_depth[bn] = inv;
        float d = 1.0f / inv;
        if (isinf(d) || isnan(d)) d = INFINITY;
        float p[3];
        p[0] = (((float)v - cx) * d) / fx;
        p[1] = (((float)u - cy) * d) / fy;
        p[2] = d;
        if (n < 3) {                        /* the dim-1 indexing quirk, SOccDPT.py:351-353 */
            const float s = J->pc_scale[n], t = J->pc_shift[n];
            p[0] = p[0] * s + t; p[1] = p[1] * s + t; p[2] = p[2] * s + t;
        }
        J->points[bn * 3 + 0] = p[0]; J->points[bn * 3 + 1] = p[1]; J->points[bn * 3 + 2] = p[2];
        if (!J->grid) continue;
        rot3(J->rot, p); rot3(J->rot + 9, p); rot3(J->rot + 18, p);
        if (!(isfinite(p[0]) && isfinite(p[1]) && isfinite(p[2]))) continue;
        const float fi = (p[0] / J->occ_shape[0]) * g0;
        const float fj = (p[1] / J->occ_shape[1]) * g1;
        const float fk = (p[2] / J->occ_shape[2]) * g2;
        /* 0 < trunc(f) < G  <=>  1 <= f < G  (also rejects NaN / int64-overflowing values) */
        if (!(fi >= 1.0f && fi < g0 && fj >= 1.0f && fj < g1 && fk >= 1.0f && fk < g2)) continue;
        const long long vox = ((long long)(int)fi * gsz[1] + (int)fj) * gsz[2] + (int)fk;
        for (int c = 0; c < C; ++c) {
            const float s = J->seg[((long long)b * C + c) * N + n];
            if (s != 0.0f) {                /* torch.nonzero: NaN counts, -0.0 does not */
                ++stores;
                /* idempotent store: racing threads all write the same 1.0f */
                J->grid[(J->per_frame ? (long long)b * cells : 0) + vox * C + c] = 1.0f;
            }
        }
    }
    J->stores = stores;
    return 0;
}

/*
 * inv_depth : (B,H,W)   in/out  -- clamped in place like the reference (SOccDPT.py:289)
 * seg       : (B,C,H,W) in
 * rot       : 3 row-major 3x3 matrices Ra,Rb,Rc (27 floats) built by the host exactly as
 *             SOccDPT.py:82-111 does (fp32 cos/sin of deg2rad(angle))
 * points    : (B,H,W,3) out     -- un-rotated points incl. the three altered pixels
 * grid      : (B,G0,G1,G2,C) out or NULL (compute_occ=False)
 * per_frame : 0 = reference semantics (union over the batch written to every b),
 *             1 = extension: each frame's own voxels
 * threads   : worker threads (<=0: all online cores)
 * returns number of (voxel,class) stores issued (with multiplicity)
 */
long long soccdpt_oracle_voxelize(float *inv_depth, const float *seg, int B, int H, int W, int C,
                                  float fx, float fy, float cx, float cy,
                                  const float *pc_scale, const float *pc_shift, const float *rot,
                                  const float *occ_shape, const int *gsz,
                                  float *points, float *grid, int per_frame, int threads)
{
    const long long total = (long long)B * H * W;
    const long long cells = (long long)gsz[0] * gsz[1] * gsz[2] * C;
    if (threads <= 0) threads = (int)sysconf(_SC_NPROCESSORS_ONLN);
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    if (grid) memset(grid, 0, sizeof(float) * (size_t)cells * B);
    job_t jobs[256]; pthread_t tid[256];
    for (int t = 0; t < threads; ++t) {
        job_t j = { inv_depth, seg, B, H, W, C, fx, fy, cx, cy, pc_scale, pc_shift, rot, occ_shape, gsz,
                    points, grid, per_frame, total * t / threads, total * (t + 1) / threads, 0 };
        jobs[t] = j;
        if (threads == 1) worker(&jobs[t]); else pthread_create(&tid[t], 0, worker, &jobs[t]);
    }
    long long stores = 0;
    for (int t = 0; t < threads; ++t) {
        if (threads > 1) pthread_join(tid[t], 0);
        stores += jobs[t].stores;
    }
    if (grid && !per_frame)                 /* occupancy_grid[:, ...] : every b gets the union */
        for (int b = 1; b < B; ++b) memcpy(grid + (size_t)b * cells, grid, sizeof(float) * (size_t)cells);
    return stores;
}
