// cluster.cuh — stage 3: pcl::search::KdTree + pcl::EuclideanClusterExtraction (opd.cpp:346-362).
//
// A Euclidean cluster is a connected component of the graph {(i,j) : d2(i,j) < (float)(tol*tol)} with
// d2 = ((dx*dx)+dy*dy)+dz*dz in float (FLANN L2_Simple, strict '<'), so the BFS order of the reference
// does not matter. One CTA per frame: tiled brute-force pair test from shared memory + lock-free
// union-find (larger root hooks under smaller root, so a component's root is its smallest member),
// then the size filter, the canonical ordering (size descending, ties by smallest member) and a
// stable scatter of the member indices (ascending inside each cluster).
//
// Roofline: FP32/latency bound, ops = 8 * M^2/2 (d-bar = M/2 candidates per point examined); bytes = 20*M.
#pragma once
#include "common.cuh"

namespace cuboid {

struct CluArgs {
    const float4* remain;   // [F][P]
    int* parent;            // [F][M]
    int* csize;             // [F][M]
    int* crank;             // [F][M]
    int* idx_sorted;        // [F][M]
    int* offsets;           // [F][KC+1]
    int* roots;             // [F][KC]
    cuboid_frame_result* res;
    int P, M, KC;
    float r2;
    int min_size, max_size, use_cluster;
};

constexpr int CLU_THREADS = 1024;

__device__ __forceinline__ int uf_find(int* parent, int x) {
    volatile int* p = parent;
    int px = p[x];
    while (px != x) {
        const int gp = p[px];
        if (gp != px) p[x] = gp;   // path halving: only ever replaces a parent by an ancestor
        x = px;
        px = gp;
    }
    return x;
}
__device__ __forceinline__ void uf_unite(int* parent, int a, int b) {
    while (true) {
        a = uf_find(parent, a);
        b = uf_find(parent, b);
        if (a == b) return;
        if (a < b) { const int t = a; a = b; b = t; }
        if (atomicCAS(&parent[a], a, b) == a) return;
    }
}

__global__ void __launch_bounds__(CLU_THREADS) k_cluster(const CluArgs a) {
    __shared__ float4 s_tile[CLU_THREADS];
    __shared__ int s_w[CLU_THREADS / 32 + 1];
    __shared__ int s_cur[1024];
    __shared__ unsigned long long s_h[CLU_THREADS / 32];
    const int f = blockIdx.x;
    cuboid_frame_result& R = a.res[f];
    const int n = R.n_remain;
    const float4* pts = a.remain + (size_t)f * a.P;
    int* parent = a.parent + (size_t)f * a.M;
    int* csize = a.csize + (size_t)f * a.M;
    int* crank = a.crank + (size_t)f * a.M;
    int* idx_sorted = a.idx_sorted + (size_t)f * a.M;
    int* offsets = a.offsets + (size_t)f * (a.KC + 1);
    int* roots = a.roots + (size_t)f * a.KC;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    if (!a.use_cluster) {   // icp.cpp:156-172: the whole non-plane cloud is the ICP source
        unsigned long long h = 0;
        for (int i = threadIdx.x; i < n; i += CLU_THREADS) { idx_sorted[i] = i; h += hash_index((unsigned int)i, i); }
        h = warp_sum_u64(h);
        if (lane == 0) s_h[wid] = h;
        __syncthreads();
        if (threadIdx.x == 0) {
            unsigned long long t = 0;
            for (int k = 0; k < CLU_THREADS / 32; ++k) t += s_h[k];
            const int ncl = n > 0 ? 1 : 0;
            offsets[0] = 0; offsets[1] = n;
            R.n_clusters = ncl;
            R.cluster[0].size = n;
            R.cluster_hash = t + splitmix64((unsigned long long)ncl);
        }
        return;
    }

    for (int i = threadIdx.x; i < n; i += CLU_THREADS) { parent[i] = i; csize[i] = 0; crank[i] = -1; }
    __syncthreads();

    // pair test: i-block ib against j-tiles jb >= ib
    const int nb = (n + CLU_THREADS - 1) / CLU_THREADS;
    for (int ib = 0; ib < nb; ++ib) {
        const int i = ib * CLU_THREADS + threadIdx.x;
        const bool iv = i < n;
        const float4 pi = iv ? pts[i] : make_float4(0.f, 0.f, 0.f, 0.f);
        for (int jb = ib; jb < nb; ++jb) {
            const int jj = jb * CLU_THREADS + threadIdx.x;
            __syncthreads();
            s_tile[threadIdx.x] = (jj < n) ? pts[jj] : make_float4(0.f, 0.f, 0.f, 0.f);
            __syncthreads();
            if (!iv) continue;
            const int jn = min(CLU_THREADS, n - jb * CLU_THREADS);
            const int j0 = (jb == ib) ? threadIdx.x + 1 : 0;
            for (int j = j0; j < jn; ++j) {
                const float4 pj = s_tile[j];
                const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                const float d2 = ((dx * dx) + dy * dy) + dz * dz;
                if (d2 < a.r2) uf_unite(parent, i, jb * CLU_THREADS + j);
            }
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += CLU_THREADS) {
        const int r = uf_find(parent, i);
        parent[i] = r;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += CLU_THREADS) atomicAdd(&csize[parent[i]], 1);
    __syncthreads();

    // kept roots, ascending (ordered compaction)
    int K = 0;
    for (int start = 0; start < n; start += CLU_THREADS) {
        const int i = start + threadIdx.x;
        bool keep = false;
        if (i < n && parent[i] == i) { const int s = csize[i]; keep = s >= a.min_size && s <= a.max_size; }
        int total;
        const int pos = K + block_excl_scan<CLU_THREADS>(keep ? 1 : 0, s_w, &total);
        if (keep && pos < a.KC) roots[pos] = i;
        K += total;
        __syncthreads();
    }
    const int Kc = min(K, min(a.KC, 1024));
    // canonical order: size descending, ties by smallest member (= root)
    for (int k = threadIdx.x; k < Kc; k += CLU_THREADS) {
        const int rk = roots[k], sk = csize[rk];
        int rank = 0, off = 0;
        for (int m = 0; m < Kc; ++m) {
            const int rm = roots[m], sm = csize[rm];
            if (sm > sk || (sm == sk && rm < rk)) { ++rank; off += sm; }
        }
        crank[rk] = rank;
        offsets[rank] = off;
        s_cur[rank] = off;
        if (rank < CUBOID_MAX_CLUSTERS) R.cluster[rank].size = sk;
        if (rank == Kc - 1) offsets[Kc] = off + sk;
    }
    if (Kc == 0 && threadIdx.x == 0) offsets[0] = 0;
    __syncthreads();

    // stable scatter by one warp: members in ascending index order inside each cluster
    if (wid == 0) {
        unsigned long long h = 0;
        for (int i0 = 0; i0 < n; i0 += 32) {
            const int i = i0 + lane;
            const int r = (i < n) ? crank[parent[i]] : -1;
            const unsigned int peers = __match_any_sync(FULL_MASK, r);
            const int leader = __ffs(peers) - 1;
            int before = 0;
            if (r >= 0 && lane == leader) { before = s_cur[r]; s_cur[r] = before + __popc(peers); }
            before = __shfl_sync(FULL_MASK, before, leader);
            if (r >= 0) {
                const int pos = before + __popc(peers & ((1u << lane) - 1u));
                idx_sorted[pos] = i;
                h += hash_index((unsigned int)pos, i);
            }
            __syncwarp();
        }
        h = warp_sum_u64(h);
        if (lane == 0) {
            R.n_clusters = Kc;
            R.cluster_hash = h + splitmix64((unsigned long long)Kc);
            if (K > Kc || Kc > CUBOID_MAX_CLUSTERS) atomicOr(&R.status, CUBOID_W_CLUSTERS_TRUNCATED);
        }
    }
}

}  // namespace cuboid
