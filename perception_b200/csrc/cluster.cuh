// cluster.cuh — stage 3: pcl::search::KdTree + pcl::EuclideanClusterExtraction (opd.cpp:346-362).
//
// A Euclidean cluster is a connected component of the graph {(i,j) : d2(i,j) < (float)(tol*tol)} with
// d2 = ((dx*dx)+dy*dy)+dz*dz in float (FLANN L2_Simple, strict '<'), so the BFS order of the reference
// does not matter, and neither does which edges are used as long as the components come out the same.
// One CTA per frame: points are binned into a hashed uniform grid of FINE cells (side 0.52 * tol). Two points
// of one cell are always within tol (3 * 0.52^2 = 0.81 < 1), so a cell is united without distance tests;
// two points within tol are at most 2 cells apart per axis (1 / 0.52 = 1.92 < 2), so cell pairs of the
// 5x5x5 neighbourhood are examined, each only until ONE pair of points within tol is found (exact float test)
// or not at all when both cells already hang under the same root. Both margins (19 % and 4 %) dwarf the float
// rounding of the cell coordinate for |coordinate / cell| < 5e5, where it is below 0.03 cells (beyond that the frame is flagged CUBOID_W_CLUSTER_RANGE). Edges feed a lock-free union-find (larger
// root hooks under smaller root, so a component's root is its smallest member). Then the size filter, the
// canonical ordering (size descending, ties by smallest member) and a stable scatter of the member indices
// (ascending inside each cluster).
//
// Roofline: latency bound; work = cells * 62 probes + (cell pairs examined) * (points per cell)^2 tests; bytes = 20*M.
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
    int* cell_start;        // [F][cell_stride] hashed grid: end of every bucket in cell_pts (cell_stride = pow2 >= 2*M, + 1)
    float4* cell_pts;       // [F][M]     points grouped by bucket, .w = point index (bits)
    cuboid_frame_result* res;
    int P, M, KC;
    int cell_stride;
    float r2;
    float inv_cell;         // 1 / (0.52 * tol)
    int min_size, max_size, use_cluster;
};

constexpr int CLU_THREADS = 1024;   // two CTAs per SM (the shared-memory tiers below are sized for that): 1.86 ms vs 2.15 ms / 1024 frames with one

__device__ __forceinline__ unsigned int cell_hash(int cx, int cy, int cz) {
    return ((unsigned int)cx * 73856093u) ^ ((unsigned int)cy * 19349663u) ^ ((unsigned int)cz * 83492791u);
}
// find with full path compression: parent links only ever point to smaller indices (a root is hooked under a
// smaller root), so there are no cycles and rewriting a node's parent to any ancestor is safe under races.
__device__ __forceinline__ int uf_find(int* parent, int x) {
    volatile int* p = parent;
    int r = x, q;
    while ((q = p[r]) != r) r = q;
    while ((q = p[x]) > r) { p[x] = r; x = q; }   // '>' not '!=': a racing thread may already have lifted x above r
    return r;
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

// MODE 0: n <= CLU_SMEM_ALL: union-find forest, bucket table and bucketed points all in shared memory
// MODE 1: n <= CLU_SMEM_UF : forest in shared memory, grid in global memory (k_cluster, two CTAs per SM);
//         n <= CLU_SMEM_UF_BIG: the same in k_cluster_big (one CTA per SM, 128 KB forest), launched next to k_cluster: every CTA
//         of either kernel looks at its frame's n_remain and leaves at once when the frame belongs to the other kernel
// MODE 2: everything in global memory
constexpr int CLU_SMEM_ALL = 2048;
constexpr int CLU_SMEM_UF = 12288;
constexpr int CLU_SMEM_UF_BIG = 32768;
// static shared memory of k_cluster, declared once in the kernel (not per MODE instantiation)
struct CluShared {
    int s_w[CLU_THREADS / 32 + 1];
    int s_cur[1024];
    unsigned long long s_h[CLU_THREADS / 32];
    int s_nrep, s_next;
};
template <int MODE>
__device__ __forceinline__ void cluster_body(const CluArgs& a, CluShared& cs) {
    int* s_w = cs.s_w;
    int* s_cur = cs.s_cur;
    unsigned long long* s_h = cs.s_h;
    const int f = blockIdx.x;
    cuboid_frame_result& R = a.res[f];
    const int n = R.n_remain;
    const float4* pts = a.remain + (size_t)f * a.P;
    // union-find forest in shared memory when the frame's remainder fits (pointer chasing at smem latency), else in global.
    // The two cases are separate instantiations so the compiler emits LDS/ATOMS, not generic accesses.
    extern __shared__ __align__(16) int s_dyn[];
    int* parent = MODE <= 1 ? s_dyn : a.parent + (size_t)f * a.M;
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

    // hashed uniform grid of fine cells: table size = power of two >= 2n; points are counting-sorted by bucket so
    // that the points of a cell are (part of) one contiguous run
    // MODE 0 shared layout: parent[4096] | cpts float4[4096] | cend[hs+1 <= 4097]
    float4* cpts = MODE == 0 ? reinterpret_cast<float4*>(s_dyn + CLU_SMEM_ALL) : a.cell_pts + (size_t)f * a.M;
    int* cend = MODE == 0 ? s_dyn + CLU_SMEM_ALL + 4 * CLU_SMEM_ALL : a.cell_start + (size_t)f * a.cell_stride;
    int hs = 64;
    while (hs < (MODE == 0 ? n : 2 * n)) hs <<= 1;
    const unsigned int hmask = (unsigned int)hs - 1u;
    for (int i = threadIdx.x; i <= hs; i += CLU_THREADS) cend[i] = 0;
    __syncthreads();
    const float inv_cell = a.inv_cell;
    bool far = false;
    for (int i = threadIdx.x; i < n; i += CLU_THREADS) {
        const float4 p = pts[i];
        const int cx = (int)floorf(p.x * inv_cell), cy = (int)floorf(p.y * inv_cell), cz = (int)floorf(p.z * inv_cell);
        atomicAdd(&cend[cell_hash(cx, cy, cz) & hmask], 1);
        far |= !(fmaxf(fmaxf(fabsf(p.x), fabsf(p.y)), fabsf(p.z)) * inv_cell <= 5.0e5f);   // also catches NaN / inf
    }
    if (far) atomicOr(&R.status, CUBOID_W_CLUSTER_RANGE);   // the cell coordinate's rounding is no longer small against the margins
    __syncthreads();
    {   // exclusive scan of the hs bucket counts (running carry over 1024-wide slices): cend[b] = start of bucket b
        int carry = 0;
        for (int b0 = 0; b0 < hs; b0 += CLU_THREADS) {
            const int b = b0 + threadIdx.x;
            const int c = b < hs ? cend[b] : 0;
            int total;
            const int ex = carry + block_excl_scan<CLU_THREADS>(c, s_w, &total);
            if (b < hs) cend[b] = ex;
            carry += total;
            __syncthreads();
        }
    }
    // scatter: the cursor of bucket b walks from its start to its end, so afterwards cend[b] = end of bucket b
    // and bucket b occupies [b ? cend[b-1] : 0, cend[b])
    for (int i = threadIdx.x; i < n; i += CLU_THREADS) {
        const float4 p = pts[i];
        const int cx = (int)floorf(p.x * inv_cell), cy = (int)floorf(p.y * inv_cell), cz = (int)floorf(p.z * inv_cell);
        const int pos = atomicAdd(&cend[cell_hash(cx, cy, cz) & hmask], 1);
        cpts[pos] = make_float4(p.x, p.y, p.z, __int_as_float(i));
    }
    __syncthreads();
    // (S) Points of one fine cell are within tol of each other (cell diagonal^2 = 3 * 0.52^2 tol^2 < r2), so every
    // point is united with the first point of its cell in bucket order without a distance test; that first point
    // is the cell's representative. One thread per bucketed position.
    int* reps = idx_sorted;     // scratch until the final scatter writes the member lists
    if (threadIdx.x == 0) { cs.s_nrep = 0; cs.s_next = 0; }
    __syncthreads();
    const bool edges = a.r2 > 0.f;
    for (int pos = threadIdx.x; edges && pos < n; pos += CLU_THREADS) {
        const float4 p = cpts[pos];
        const int cx = (int)floorf(p.x * inv_cell), cy = (int)floorf(p.y * inv_cell), cz = (int)floorf(p.z * inv_cell);
        const unsigned int b = cell_hash(cx, cy, cz) & hmask;
        int q = b ? cend[b - 1] : 0;
        float4 pq = p;
        for (; q < pos; ++q) {
            pq = cpts[q];
            if ((int)floorf(pq.x * inv_cell) == cx && (int)floorf(pq.y * inv_cell) == cy && (int)floorf(pq.z * inv_cell) == cz) break;
        }
        if (q == pos) reps[atomicAdd(&cs.s_nrep, 1)] = pos;
        else uf_unite(parent, __float_as_int(p.w), __float_as_int(pq.w));
    }
    __syncthreads();
    // (P) Cell pairs, one warp per representative: two points closer than tol sit in cells at most 2 apart per axis
    // (tol / cell = 1.92), and each unordered pair of cells is visited once, from the cell whose offset to the other
    // is in the upper half of the 5x5x5 neighbourhood (62 offsets, probed by the lanes in two rounds). For every
    // non-empty neighbour not already under the same root the lanes test point pairs with the exact float distance
    // until the first one within tol; that single edge joins the two cells.
    // Representatives are claimed dynamically (cells differ a lot in cost); the edges found in a round are applied
    // together afterwards, one per lane, instead of one at a time while the other lanes wait.
    const int nrep = cs.s_nrep;
    while (true) {
        int r = 0;
        if (lane == 0) r = atomicAdd(&cs.s_next, 1);
        r = __shfl_sync(FULL_MASK, r, 0);
        if (r >= nrep) break;
        const float4 pa0 = cpts[reps[r]];
        const int cx = (int)floorf(pa0.x * inv_cell), cy = (int)floorf(pa0.y * inv_cell), cz = (int)floorf(pa0.z * inv_cell);
        const unsigned int bA = cell_hash(cx, cy, cz) & hmask;
        const int sA = bA ? cend[bA - 1] : 0, lenA = cend[bA] - sA;
        const int ia0 = __float_as_int(pa0.w);
        for (int round = 0; round < 2; ++round) {
            const int k = round * 32 + lane;
            const int rootA = uf_find(parent, ia0);
            int sB = 0, lenB = 0;
            if (k < 62) {
                const int code = k + 63;
                const int tx = cx + code % 5 - 2, ty = cy + (code / 5) % 5 - 2, tz = cz + code / 25 - 2;
                const unsigned int b = cell_hash(tx, ty, tz) & hmask;
                sB = b ? cend[b - 1] : 0;
                lenB = cend[b] - sB;
                if (lenB > 0) {     // cheap skip when the run starts with a point of the target cell that already shares A's root
                    const float4 pb0 = cpts[sB];
                    if ((int)floorf(pb0.x * inv_cell) == tx && (int)floorf(pb0.y * inv_cell) == ty && (int)floorf(pb0.z * inv_cell) == tz &&
                        uf_find(parent, __float_as_int(pb0.w)) == rootA)
                        lenB = 0;
                }
            }
            unsigned int todo = __ballot_sync(FULL_MASK, lenB > 0);
            int ua = -1, ub = -1;       // the edge found for this lane's neighbour
            while (todo) {
                const int src = __ffs(todo) - 1;
                todo &= todo - 1;
                const int s2 = __shfl_sync(FULL_MASK, sB, src), l2 = __shfl_sync(FULL_MASK, lenB, src);
                const int code = round * 32 + src + 63;
                const int tx = cx + code % 5 - 2, ty = cy + (code / 5) % 5 - 2, tz = cz + code / 25 - 2;
                // pairs (qa, qb) of run A x run B: the lanes form rows of width 2^sh >= l2 (no division); a run of
                // more than 32 points is swept 32 at a time for every point of A
                const int sh = l2 <= 1 ? 0 : 32 - __clz(l2 - 1);
                const bool wide = sh > 5;
                const int rows = wide ? 1 : 32 >> sh;
                const int row = wide ? 0 : lane >> sh, col = wide ? lane : lane & ((1 << sh) - 1);
                bool found = false;
                for (int qa0 = 0; qa0 < lenA && !found; qa0 += rows) {
                    const int qa = qa0 + row;
                    for (int qb0 = 0; qb0 < l2; qb0 += 32) {
                        const int qb = qb0 + col;
                        bool hit = false;
                        int ia = 0, ib = 0;
                        if (qa < lenA && qb < l2) {
                            const float4 pa = cpts[sA + qa], pb = cpts[s2 + qb];
                            const float ddx = pa.x - pb.x, ddy = pa.y - pb.y, ddz = pa.z - pb.z;
                            const float d2 = ((ddx * ddx) + ddy * ddy) + ddz * ddz;
                            if (d2 < a.r2 &&        // buckets can hold foreign cells: both points must be in the cells of this pair
                                (int)floorf(pa.x * inv_cell) == cx && (int)floorf(pa.y * inv_cell) == cy && (int)floorf(pa.z * inv_cell) == cz &&
                                (int)floorf(pb.x * inv_cell) == tx && (int)floorf(pb.y * inv_cell) == ty && (int)floorf(pb.z * inv_cell) == tz) {
                                hit = true; ia = __float_as_int(pa.w); ib = __float_as_int(pb.w);
                            }
                        }
                        const unsigned int hm = __ballot_sync(FULL_MASK, hit);
                        if (hm) {
                            const int hl = __ffs(hm) - 1;
                            const int ja = __shfl_sync(FULL_MASK, ia, hl), jb = __shfl_sync(FULL_MASK, ib, hl);
                            if (lane == src) { ua = ja; ub = jb; }
                            found = true;
                            break;
                        }
                        if (!wide) break;   // one sweep covers the whole run
                    }
                }
            }
            if (ua >= 0) uf_unite(parent, ua, ub);
            __syncwarp();
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < n; i += CLU_THREADS) {
        const int r = uf_find(parent, i);
        parent[i] = r;
    }
    __syncthreads();
    for (int i0 = 0; i0 < n; i0 += CLU_THREADS) {   // warp-aggregated: one atomic per distinct root per warp
        const int i = i0 + threadIdx.x;
        const int r = i < n ? parent[i] : -1;
        const unsigned int peers = __match_any_sync(FULL_MASK, r);
        if (r >= 0 && lane == __ffs(peers) - 1) atomicAdd(&csize[r], __popc(peers));
    }
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

__global__ void __launch_bounds__(CLU_THREADS) k_cluster(const CluArgs a) {
    __shared__ CluShared cs;
    const int n = a.res[blockIdx.x].n_remain;
    if (n <= CLU_SMEM_ALL) cluster_body<0>(a, cs);
    else if (n <= CLU_SMEM_UF) cluster_body<1>(a, cs);
    else if (n <= CLU_SMEM_UF_BIG && a.use_cluster) return;   // k_cluster_big's frame
    else cluster_body<2>(a, cs);
}
// frames with CLU_SMEM_UF < n_remain <= CLU_SMEM_UF_BIG (multi-object scenes): the union-find forest still fits shared memory when a
// CTA has an SM to itself
__global__ void __launch_bounds__(CLU_THREADS, 1) k_cluster_big(const CluArgs a) {
    __shared__ CluShared cs;
    const int n = a.res[blockIdx.x].n_remain;
    if (n > CLU_SMEM_UF && n <= CLU_SMEM_UF_BIG && a.use_cluster) cluster_body<1>(a, cs);
}
constexpr size_t CLU_DYN_SMEM_BIG = (size_t)CLU_SMEM_UF_BIG * 4;
constexpr size_t CLU_DYN_SMEM = (size_t)CLU_SMEM_ALL * 4 + (size_t)CLU_SMEM_ALL * 16 + (size_t)(CLU_SMEM_ALL + 4) * 4;   // 49 168 B >= 12288*4

}  // namespace cuboid
