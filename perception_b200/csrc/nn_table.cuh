// nn_table.cuh — exact nearest-neighbour candidate table of an ICP template (icp.cpp:170-178: the correspondence search of
// pcl::IterativeClosestPoint).
//
// Once ICP has roughly aligned the clouds almost every query lies within a few millimetres of the template, where the set of
// template points that can be ITS nearest neighbour is tiny: the points whose Voronoi cells reach the query's neighbourhood.
// The table stores that set for every cell of a dense grid around the template (cell side h, default 1 mm): a 64-byte record
// (two L2 sectors; cells with at most 15 candidates are answered from the first) with up to 31 kd-ordered template positions,
// sorted by ORIGINAL template index. A query whose record is
// valid scans those few points with the un-fused float distance and strict '<' in that order — exactly what a brute-force scan
// of the whole template in original order with strict '<' returns (the canonical tie rule). Queries outside the grid, or in
// cells that would need more than 31 candidates (far from the template), are misses and go to the BVH search (icp.cuh).
//
// Why the candidate set is complete (no float-minimiser can be missing). Let V be the cell, widened by delta on every side so
// that it contains every query the run-time index computation can map to it (that computation's rounding error is below
// 1e-4 cells; delta = 1e-3 cells). For a sub-box s of V and a template point t let near(t, s) / far(t, s) be the smallest /
// largest distance from t to s (exact, double). If near(t, s)^2 > U (1 + 1e-5) with U = min over t' of far(t', s)^2, then for
// every query q in s: d2(q, t) >= near^2 > U (1 + 1e-5) >= d2(q, t') (1 + 1e-5) for the t' attaining U. The run-time float
// evaluation of d2 is within 3 ulp (2e-7 relative) of the exact value, so t is STRICTLY farther than t' in float as well: t can
// neither win nor tie. The record holds the union over the 4 x 4 x 4 sub-boxes of the cell of the points that survive this
// test (the minimum U is taken over a first-level superset computed the same way on the whole cell; a larger U only keeps
// more points). Ties between surviving points are resolved at run time on the float distances themselves.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace cuboid {

constexpr int NNT_K = 31;         // candidates per record (64 bytes: two L2 sectors; the second one is only read for cells with more than 15)
constexpr int NNT_C1MAX = 160;    // first-level candidates per cell; more = the cell is too far from the template
constexpr int NNT_SUB = 4;        // sub-boxes per axis of the second level
constexpr int NNT_THREADS = 128;
constexpr double NNT_BAND = 25.0;   // cells of margin around the template's bounding box = how far from the template a query can still be answered

struct NnTableView {              // what k_icp needs at run time
    const uint4* rec;             // [nz][ny][nx][4]: halfword 0 = n (0xffff: miss), halfwords 1..31 = kd-ordered positions
    float org[3];
    float inv_h;
    int nx, ny, nz;
    // seed grid for the queries the table cannot answer (far from the template): per coarse cell the kd position of the template point
    // nearest to the cell's centre. ANY template position is a valid seed of the exact BVH search; a near one makes its bound tight.
    const unsigned short* seed;   // [snz][sny][snx], NULL = none
    float sorg[3];
    float sinv_h;
    int snx, sny, snz;
};

struct NnTableGeom {              // build-time geometry, all in double of the exact float values the run time uses
    double org[3], h, delta, band2;
    int nx, ny, nz;
};

__device__ __forceinline__ void nnt_near_far(double tx, double ty, double tz, const double lo[3], const double hi[3], double& near2, double& far2) {
    const double t[3] = {tx, ty, tz};
    near2 = 0.0; far2 = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double dl = lo[a] - t[a], dh = t[a] - hi[a];
        const double nr = fmax(fmax(dl, dh), 0.0);
        const double fr = fmax(fabs(dl), fabs(dh));
        near2 += nr * nr;
        far2 += fr * fr;
    }
}

// One thread per cell. tmpl: kd-ordered SoA leaves [nleaf][3][LEAF] (the array k_icp stages), n real points (sentinel
// padding beyond n is ignored), orig: original index per kd position.
template <int LEAF>
__global__ void __launch_bounds__(NNT_THREADS) k_nn_table_build(const float* __restrict__ tmpl, const int* __restrict__ orig, int n, NnTableGeom g,
                                                                uint4* __restrict__ rec) {
    constexpr int TILE = 512;
    __shared__ float s_x[TILE], s_y[TILE], s_z[TILE];
    const long long nvox = (long long)g.nx * g.ny * g.nz;
    const long long v = (long long)blockIdx.x * NNT_THREADS + threadIdx.x;
    const bool live = v < nvox;
    const int ix = live ? (int)(v % g.nx) : 0, iy = live ? (int)((v / g.nx) % g.ny) : 0, iz = live ? (int)(v / ((long long)g.nx * g.ny)) : 0;
    const int iv[3] = {ix, iy, iz};
    double lo[3], hi[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) { lo[a] = g.org[a] + iv[a] * g.h - g.delta; hi[a] = g.org[a] + (iv[a] + 1) * g.h + g.delta; }
    // ---- level 1, pass 1: U = min far^2 over the whole template; m = min near^2 (band test) ----
    double U = 1.0e300, m = 1.0e300;
    for (int t0 = 0; t0 < n; t0 += TILE) {
        __syncthreads();
        for (int j = threadIdx.x; j < TILE; j += NNT_THREADS) {
            const int p = t0 + j;
            if (p < n) {
                const float* lf = tmpl + (size_t)(p / LEAF) * (3 * LEAF) + (p % LEAF);
                s_x[j] = lf[0]; s_y[j] = lf[LEAF]; s_z[j] = lf[2 * LEAF];
            }
        }
        __syncthreads();
        const int cnt = min(TILE, n - t0);
        for (int j = 0; j < cnt; ++j) {
            double nr, fr;
            nnt_near_far((double)s_x[j], (double)s_y[j], (double)s_z[j], lo, hi, nr, fr);
            U = fmin(U, fr);
            m = fmin(m, nr);
        }
    }
    bool far_cell = !live || m > g.band2;
    // ---- level 1, pass 2: first-level candidates ----
    unsigned short c1[NNT_C1MAX];
    int n1 = 0;
    const double thr1 = U * (1.0 + 1.0e-5) + 1.0e-30;
    for (int t0 = 0; t0 < n; t0 += TILE) {
        __syncthreads();
        for (int j = threadIdx.x; j < TILE; j += NNT_THREADS) {
            const int p = t0 + j;
            if (p < n) {
                const float* lf = tmpl + (size_t)(p / LEAF) * (3 * LEAF) + (p % LEAF);
                s_x[j] = lf[0]; s_y[j] = lf[LEAF]; s_z[j] = lf[2 * LEAF];
            }
        }
        __syncthreads();
        if (far_cell) continue;
        const int cnt = min(TILE, n - t0);
        for (int j = 0; j < cnt; ++j) {
            double nr, fr;
            nnt_near_far((double)s_x[j], (double)s_y[j], (double)s_z[j], lo, hi, nr, fr);
            if (nr <= thr1) {
                if (n1 < NNT_C1MAX) c1[n1] = (unsigned short)(t0 + j);
                ++n1;
            }
        }
    }
    if (n1 > NNT_C1MAX) far_cell = true;
    if (!live) return;
    uint4* out = rec + 4 * v;
    if (far_cell) {
        out[0] = make_uint4(0xffffu, 0u, 0u, 0u);
        out[1] = out[2] = out[3] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    // ---- level 2: union over the sub-boxes of the points that can be a nearest neighbour inside the sub-box ----
    unsigned int mark[(NNT_C1MAX + 31) / 32] = {};
    const double hs = g.h / NNT_SUB;
    for (int sz = 0; sz < NNT_SUB; ++sz)
        for (int sy = 0; sy < NNT_SUB; ++sy)
            for (int sx = 0; sx < NNT_SUB; ++sx) {
                const int si[3] = {sx, sy, sz};
                double slo[3], shi[3];
#pragma unroll
                for (int a = 0; a < 3; ++a) {
                    slo[a] = g.org[a] + iv[a] * g.h + si[a] * hs - g.delta;
                    shi[a] = g.org[a] + iv[a] * g.h + (si[a] + 1) * hs + g.delta;
                }
                double Us = 1.0e300;
                for (int c = 0; c < n1; ++c) {
                    const int p = c1[c];
                    const float* lf = tmpl + (size_t)(p / LEAF) * (3 * LEAF) + (p % LEAF);
                    double nr, fr;
                    nnt_near_far((double)lf[0], (double)lf[LEAF], (double)lf[2 * LEAF], slo, shi, nr, fr);
                    Us = fmin(Us, fr);
                }
                const double thr = Us * (1.0 + 1.0e-5) + 1.0e-30;
                for (int c = 0; c < n1; ++c) {
                    const int p = c1[c];
                    const float* lf = tmpl + (size_t)(p / LEAF) * (3 * LEAF) + (p % LEAF);
                    double nr, fr;
                    nnt_near_far((double)lf[0], (double)lf[LEAF], (double)lf[2 * LEAF], slo, shi, nr, fr);
                    if (nr <= thr) mark[c >> 5] |= 1u << (c & 31);
                }
            }
    int pos[NNT_K], org_[NNT_K];
    int k = 0;
    bool over = false;
    for (int c = 0; c < n1; ++c) {
        if (!((mark[c >> 5] >> (c & 31)) & 1u)) continue;
        if (k == NNT_K) { over = true; break; }
        const int p = c1[c], o = orig[p];
        int j = k++;
        while (j > 0 && org_[j - 1] > o) { org_[j] = org_[j - 1]; pos[j] = pos[j - 1]; --j; }   // ascending ORIGINAL index
        org_[j] = o; pos[j] = p;
    }
    if (over) {
        out[0] = make_uint4(0xffffu, 0u, 0u, 0u);
        out[1] = out[2] = out[3] = make_uint4(0u, 0u, 0u, 0u);
        return;
    }
    unsigned int w[16];
#pragma unroll
    for (int q = 0; q < 16; ++q) w[q] = 0u;
    w[0] = (unsigned int)k;
    for (int c = 0; c < k; ++c) {
        const int hw = c + 1;
        w[hw >> 1] |= (unsigned int)pos[c] << ((hw & 1) * 16);
    }
    out[0] = make_uint4(w[0], w[1], w[2], w[3]);
    out[1] = make_uint4(w[4], w[5], w[6], w[7]);
    out[2] = make_uint4(w[8], w[9], w[10], w[11]);
    out[3] = make_uint4(w[12], w[13], w[14], w[15]);
}


// seed grid: one thread per coarse cell, brute-force nearest template point of the cell centre (float: it is only a seed)
template <int LEAF>
__global__ void __launch_bounds__(NNT_THREADS) k_nn_seed_build(const float* __restrict__ tmpl, int n, float ox, float oy, float oz, float h, int nx, int ny, int nz,
                                                               unsigned short* __restrict__ seed) {
    constexpr int TILE = 512;
    __shared__ float s_x[TILE], s_y[TILE], s_z[TILE];
    const long long nc = (long long)nx * ny * nz;
    const long long v = (long long)blockIdx.x * NNT_THREADS + threadIdx.x;
    const bool live = v < nc;
    const int ix = live ? (int)(v % nx) : 0, iy = live ? (int)((v / nx) % ny) : 0, iz = live ? (int)(v / ((long long)nx * ny)) : 0;
    const float cx = ox + ((float)ix + 0.5f) * h, cy = oy + ((float)iy + 0.5f) * h, cz = oz + ((float)iz + 0.5f) * h;
    float best = 3.0e38f;
    int bpos = 0;
    for (int t0 = 0; t0 < n; t0 += TILE) {
        __syncthreads();
        for (int j = threadIdx.x; j < TILE; j += NNT_THREADS) {
            const int p = t0 + j;
            if (p < n) {
                const float* lf = tmpl + (size_t)(p / LEAF) * (3 * LEAF) + (p % LEAF);
                s_x[j] = lf[0]; s_y[j] = lf[LEAF]; s_z[j] = lf[2 * LEAF];
            }
        }
        __syncthreads();
        const int cnt = min(TILE, n - t0);
        for (int j = 0; j < cnt; ++j) {
            const float dx = cx - s_x[j], dy = cy - s_y[j], dz = cz - s_z[j];
            const float d = dx * dx + dy * dy + dz * dz;
            if (d < best) { best = d; bpos = t0 + j; }
        }
    }
    if (live) seed[v] = (unsigned short)bpos;
}

}  // namespace cuboid
