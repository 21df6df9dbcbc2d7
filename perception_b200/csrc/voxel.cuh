// voxel.cuh — stage 1b: pcl::VoxelGrid<PCLPointCloud2> (gps.cpp:69-73, opd.cpp:294-298).
//
//   k_voxel_keys     idx = ijk0 + ijk1*dx + ijk2*dx*dy per point (the bit-exact "voxel key"), packed as the
//                    64-bit sort record (sortkey << 32 | point#)
//   k_sort_hist/scan/scatter   segmented (per frame) stable LSD radix sort, 8-bit digits, only the
//                    significant key bits; stable => points of a voxel stay in ascending point order,
//                    the canonical replacement for PCL's unstable std::sort (SURVEY.md A.2 hazard)
//   k_voxel_reduce   segment heads -> voxel ordinal (decoupled look-back), sequential float centroid
//
// Roofline: HBM. Algorithmic bytes per frame = 16*N + 16*V (sort scratch is overhead, not counted).
#pragma once
#include "common.cuh"

namespace cuboid {

struct VoxArgs {
    const float4* pts;            // [F][P]
    unsigned long long* keysA;    // [F][P]
    unsigned long long* keysB;    // [F][P]
    int* kpp;                     // [F][P] voxel idx per point (parity tap) or NULL
    unsigned int* hist;           // [F][256][tilesP]
    float4* vox;                  // [F][P]
    int* vcount;                  // [F][P] points per voxel or NULL
    cuboid_frame_result* res;
    FrameScratch* scr;
    unsigned long long* desc;     // [F][tilesV] look-back descriptors (zeroed)
    unsigned int* ticket;         // [F] zeroed
    int P, tilesP, tilesV, n_frames;
    float inv_leaf;
};

constexpr int SORT_THREADS = 256;
constexpr int SORT_ITEMS = 8;
constexpr int SORT_TILE = SORT_THREADS * SORT_ITEMS;  // 2048

__global__ void __launch_bounds__(256) k_voxel_keys(const VoxArgs a) {
    const int f = blockIdx.y;
    const int N = a.res[f].n_points;
    const int i = blockIdx.x * 256 + threadIdx.x;
    if (blockIdx.x * 256 >= N) return;
    const VoxelGeom g = voxel_geom(a.scr[f], a.inv_leaf);
    unsigned long long h = 0;
    if (i < N) {
        const float4 p = a.pts[(size_t)f * a.P + i];
        const int idx = voxel_index(g, p.x, p.y, p.z);
        const unsigned int sk = g.overflow_mode ? ((unsigned int)idx ^ 0x80000000u) : (unsigned int)idx;
        a.keysA[(size_t)f * a.P + i] = ((unsigned long long)sk << 32) | (unsigned int)i;
        if (a.kpp) a.kpp[(size_t)f * a.P + i] = idx;
        h = hash_index((unsigned int)i, idx);
    }
    h = warp_sum_u64(h);
    __shared__ unsigned long long s_h[8];
    if ((threadIdx.x & 31) == 0) s_h[threadIdx.x >> 5] = h;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long t = 0;
        for (int k = 0; k < 8; ++k) t += s_h[k];
        atomic_add_u64(&a.res[f].voxel_key_hash, t);
        if (blockIdx.x == 0) {
            a.scr[f].sort_bits = g.sort_bits;
            a.scr[f].overflow_mode = g.overflow_mode;
            for (int c = 0; c < 3; ++c) { a.res[f].min_b[c] = g.min_b[c]; a.res[f].div_b[c] = g.div_b[c]; }
            if (g.pcl_overflow) atomicOr(&a.res[f].status, CUBOID_W_VOXEL_OVERFLOW);
        }
    }
}

// frames with no surviving point never reach k_voxel_keys' block 0; give them defined geometry
__global__ void k_voxel_empty(const VoxArgs a) {
    const int f = blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= a.n_frames) return;
    if (a.res[f].n_points == 0) { a.scr[f].sort_bits = 0; a.res[f].n_voxels = 0; }
}

__device__ __forceinline__ const unsigned long long* sort_src(const VoxArgs& a, int f, int pass) {
    return ((pass & 1) ? a.keysB : a.keysA) + (size_t)f * a.P;
}
__device__ __forceinline__ unsigned long long* sort_dst(const VoxArgs& a, int f, int pass) {
    return ((pass & 1) ? a.keysA : a.keysB) + (size_t)f * a.P;
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_hist(const VoxArgs a, int pass) {
    const int f = blockIdx.y, t = blockIdx.x;
    const int N = a.res[f].n_points;
    if (t * SORT_TILE >= N || pass * 8 >= a.scr[f].sort_bits) return;
    __shared__ unsigned int s_h[256];
    s_h[threadIdx.x] = 0;
    __syncthreads();
    const unsigned long long* src = sort_src(a, f, pass);
    const int shift = 32 + pass * 8;
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const int i = t * SORT_TILE + k * SORT_THREADS + threadIdx.x;
        if (i < N) atomicAdd(&s_h[(unsigned int)(src[i] >> shift) & 255u], 1u);
    }
    __syncthreads();
    a.hist[((size_t)f * 256 + threadIdx.x) * a.tilesP + t] = s_h[threadIdx.x];
}

// per frame: exclusive scan of hist[digit][tile] in (digit, tile) order, in place
__global__ void __launch_bounds__(256) k_sort_scan(const VoxArgs a, int pass) {
    const int f = blockIdx.x;
    const int N = a.res[f].n_points;
    if (N == 0 || pass * 8 >= a.scr[f].sort_bits) return;
    const int nt = (N + SORT_TILE - 1) / SORT_TILE;
    unsigned int* row = a.hist + ((size_t)f * 256 + threadIdx.x) * a.tilesP;
    unsigned int tot = 0;
    for (int t = 0; t < nt; ++t) tot += row[t];
    __shared__ int s_w[9];
    int total;
    unsigned int run = (unsigned int)block_excl_scan256((int)tot, s_w, &total);
    for (int t = 0; t < nt; ++t) { const unsigned int c = row[t]; row[t] = run; run += c; }
}

__global__ void __launch_bounds__(SORT_THREADS) k_sort_scatter(const VoxArgs a, int pass) {
    const int f = blockIdx.y, t = blockIdx.x;
    const int N = a.res[f].n_points;
    if (t * SORT_TILE >= N || pass * 8 >= a.scr[f].sort_bits) return;
    __shared__ unsigned int s_cnt[8][256];   // per-warp digit counters -> per-warp bases
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int d = lane; d < 256; d += 32) s_cnt[w][d] = 0;
    __syncwarp();
    const unsigned long long* src = sort_src(a, f, pass);
    unsigned long long* dst = sort_dst(a, f, pass);
    const int shift = 32 + pass * 8;
    // warp w owns the contiguous run [t*2048 + w*256, +256): item k, lane l <-> element w*256 + k*32 + l,
    // so (warp, item, lane) order == ascending input order and the in-warp ranks below are stable.
    unsigned long long key[SORT_ITEMS];
    unsigned int rank[SORT_ITEMS];
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const int i = t * SORT_TILE + w * (32 * SORT_ITEMS) + k * 32 + lane;
        const bool valid = i < N;
        key[k] = valid ? src[i] : ~0ull;
        const unsigned int d = valid ? ((unsigned int)(key[k] >> shift) & 255u) : 256u;
        const unsigned int peers = __match_any_sync(FULL_MASK, d);
        const int leader = __ffs(peers) - 1;
        unsigned int before = 0;
        if (valid && lane == leader) { before = s_cnt[w][d]; s_cnt[w][d] = before + __popc(peers); }
        before = __shfl_sync(FULL_MASK, before, leader);
        rank[k] = before + __popc(peers & ((1u << lane) - 1u));
        __syncwarp();
    }
    __syncthreads();
    {   // digit d = threadIdx.x: exclusive scan over the 8 warps, plus this tile's global base for the digit
        const int d = threadIdx.x;
        unsigned int run = a.hist[((size_t)f * 256 + d) * a.tilesP + t];
#pragma unroll
        for (int ww = 0; ww < 8; ++ww) { const unsigned int c = s_cnt[ww][d]; s_cnt[ww][d] = run; run += c; }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < SORT_ITEMS; ++k) {
        const int i = t * SORT_TILE + w * (32 * SORT_ITEMS) + k * 32 + lane;
        if (i < N) {
            const unsigned int d = (unsigned int)(key[k] >> shift) & 255u;
            dst[s_cnt[w][d] + rank[k]] = key[k];
        }
    }
}

constexpr int VR_THREADS = 256;
constexpr int VR_ITEMS = 4;
constexpr int VR_TILE = VR_THREADS * VR_ITEMS;  // 1024

// Segment heads -> voxel ordinal (decoupled look-back over the frame's tiles) and the sequential float centroid.
// The tile's keys and (gathered) points are staged in shared memory with fully parallel loads first; the
// serial per-voxel sums then run out of shared memory. A voxel that runs past the tile (e.g. the origin voxel
// that collects every zero-depth pixel) is finished by warp 0 with lane-parallel loads and an in-order
// shuffle accumulation, so the float sum order stays exactly "ascending point index".
__global__ void __launch_bounds__(VR_THREADS) k_voxel_reduce(const VoxArgs a) {
    __shared__ int s_w[9];
    __shared__ int s_tile, s_base;
    __shared__ unsigned long long s_h[8];
    __shared__ unsigned int s_key[VR_TILE];
    __shared__ float s_x[VR_TILE], s_y[VR_TILE], s_z[VR_TILE];
    __shared__ float s_open[3];
    __shared__ int s_open_cnt, s_open_pos, s_open_flag;
    __shared__ unsigned int s_open_key;
    const int f = blockIdx.x / a.tilesV;   // per-frame ticket, see k_preprocess
    if (f >= a.n_frames) return;
    if (threadIdx.x == 0) { s_tile = (int)atomicAdd(a.ticket + f, 1u); s_open_flag = 0; }
    __syncthreads();
    const int t = s_tile;
    const int N = a.res[f].n_points;
    const int ntile = (N + VR_TILE - 1) / VR_TILE;
    if (t >= ntile) return;   // tiles past the data are never looked at by live tiles (they only look back)
    const int npass = (a.scr[f].sort_bits + 7) / 8;
    const unsigned long long* keys = ((npass & 1) ? a.keysB : a.keysA) + (size_t)f * a.P;
    const float4* pts = a.pts + (size_t)f * a.P;
    const int base = t * VR_TILE;
    const int n_here = min(VR_TILE, N - base);
#pragma unroll
    for (int k = 0; k < VR_ITEMS; ++k) {
        const int l = k * VR_THREADS + threadIdx.x;
        if (l < n_here) {
            const unsigned long long r = keys[base + l];
            const float4 p = pts[(unsigned int)r];
            s_key[l] = (unsigned int)(r >> 32);
            s_x[l] = p.x; s_y[l] = p.y; s_z[l] = p.z;
        }
    }
    const unsigned int prev_key = base > 0 ? (unsigned int)(keys[base - 1] >> 32) : 0u;
    __syncthreads();
    const int first = threadIdx.x * VR_ITEMS;
    unsigned int heads = 0;
#pragma unroll
    for (int k = 0; k < VR_ITEMS; ++k) {
        const int lp = first + k;
        if (lp < n_here && (base + lp == 0 || s_key[lp] != (lp > 0 ? s_key[lp - 1] : prev_key))) heads |= 1u << k;
    }
    int total;
    int pos = block_excl_scan256(__popc(heads), s_w, &total);
    if (threadIdx.x < 32) {
        const int ex = lookback_exclusive_warp(a.desc + (size_t)f * a.tilesV, t, total);
        if (threadIdx.x == 0) s_base = ex;
    }
    __syncthreads();
    pos += s_base;
    float4* vox = a.vox + (size_t)f * a.P;
    unsigned long long h = 0;
#pragma unroll
    for (int k = 0; k < VR_ITEMS; ++k) {
        if (!(heads & (1u << k))) continue;
        const int lp = first + k;
        const unsigned int mykey = s_key[lp];
        float sx = s_x[lp], sy = s_y[lp], sz = s_z[lp];   // centroid starts as the first point, then += in sorted order
        int q = lp + 1;
        while (q < n_here && s_key[q] == mykey) { sx += s_x[q]; sy += s_y[q]; sz += s_z[q]; ++q; }
        if (q == n_here && base + n_here < N) {   // may continue in the next tile(s): at most one such voxel per tile
            s_open[0] = sx; s_open[1] = sy; s_open[2] = sz;
            s_open_cnt = q - lp; s_open_pos = pos; s_open_key = mykey; s_open_flag = 1;
        } else {
            const float cnt = (float)(q - lp);
            const float cx = sx / cnt, cy = sy / cnt, cz = sz / cnt;
            vox[pos] = make_float4(cx, cy, cz, 1.0f);
            if (a.vcount) a.vcount[(size_t)f * a.P + pos] = q - lp;
            h += hash_point((unsigned int)pos, cx, cy, cz);
        }
        ++pos;
    }
    __syncthreads();
    if (s_open_flag && threadIdx.x < 32) {
        const int lane = threadIdx.x;
        const unsigned int mykey = s_open_key;
        float sx = s_open[0], sy = s_open[1], sz = s_open[2];
        int cnt = s_open_cnt;
        int q = base + n_here;
        while (q < N) {
            const int i = q + lane;
            unsigned int kk = ~mykey;
            float px = 0.f, py = 0.f, pz = 0.f;
            if (i < N) {
                const unsigned long long r = keys[i];
                kk = (unsigned int)(r >> 32);
                if (kk == mykey) { const float4 p = pts[(unsigned int)r]; px = p.x; py = p.y; pz = p.z; }
            }
            const unsigned int miss = __ballot_sync(FULL_MASK, kk != mykey);
            const int nmatch = miss ? (__ffs(miss) - 1) : 32;
            for (int j = 0; j < nmatch; ++j) {
                sx += __shfl_sync(FULL_MASK, px, j);
                sy += __shfl_sync(FULL_MASK, py, j);
                sz += __shfl_sync(FULL_MASK, pz, j);
            }
            cnt += nmatch;
            if (nmatch < 32) break;
            q += 32;
        }
        if (lane == 0) {
            const float c = (float)cnt;
            const float cx = sx / c, cy = sy / c, cz = sz / c;
            const int vp = s_open_pos;
            vox[vp] = make_float4(cx, cy, cz, 1.0f);
            if (a.vcount) a.vcount[(size_t)f * a.P + vp] = cnt;
            h += hash_point((unsigned int)vp, cx, cy, cz);
        }
    }
    h = warp_sum_u64(h);
    if ((threadIdx.x & 31) == 0) s_h[threadIdx.x >> 5] = h;
    __syncthreads();
    if (threadIdx.x == 0) {
        unsigned long long tsum = 0;
        for (int k = 0; k < 8; ++k) tsum += s_h[k];
        if (tsum) atomic_add_u64(&a.res[f].voxel_hash, tsum);
        if (t == ntile - 1) a.res[f].n_voxels = s_base + total;
    }
}

}  // namespace cuboid
