// preprocess.cuh — stage 1: depth -> cloud unprojection (or PointCloud2 blob read) fused with both
// PassThrough filters, order-preserving compaction, min/max reduction and the parity hash.
//
// Replaces: realsense2_camera deprojection (intrinsics README.md:78) +
//           pcl::PassThrough z [0,0.9] and x [-0.2,0.2]  (gps.cpp:53-65, opd.cpp:275-289) +
//           pcl::getMinMax3D inside VoxelGrid            (gps.cpp:69-73).
// Roofline: HBM. Algorithmic bytes per frame = 2*P + 16*N (depth) or point_step*P + 16*N (cloud).
// One pass: tiles of 2048 inputs, decoupled look-back per frame for the ordered output offset, kept
// points staged in shared memory so the float4 stores are fully coalesced.
#pragma once
#include "common.cuh"

namespace cuboid {

struct PreArgs {
    const uint16_t* depth;       // [F][P] (SRC 0)
    const unsigned char* blob;   // [F][P*point_step] (SRC 1)
    int point_step, xoff, yoff, zoff;
    int rgboff;                  // offset of the packed rgb / rgba field inside a record, or -1: carried in .w of every point (fused front end)
    const int* n_in;             // per-frame input count (SRC 1) or NULL -> P
    int w, h;                    // image size (SRC 0)
    unsigned int w_magic;        // ceil(2^32 / w) when i / w == __umulhi(i, w_magic) for every input index (P*w <= 2^32), else 0
    int P;                       // inputs per frame (input stride)
    int Pout;                    // output stride of pts (the handle's max_points)
    float fx, fy, cx, cy, depth_scale;
    const float* xr;             // [w] ((float)u - cx) / fx, precomputed with the same float ops (SRC 0)
    const float* yr;             // [h] ((float)v - cy) / fy
    float z_lo, z_hi, x_lo, x_hi; // float-exact equivalents of the double limits (see host: limit_lo/limit_hi)
    float4* pts;                 // [F][P]
    cuboid_frame_result* res;    // [F]
    FrameScratch* scr;           // [F]
    int* tile_count;             // [F][tiles] survivors per tile (written by k_pre_count)
    int tiles;                   // tiles per frame
    int n_frames;
};

constexpr int PRE_THREADS = 256;
constexpr int PRE_ITEMS = 8;
constexpr int PRE_TILE = PRE_THREADS * PRE_ITEMS;

// keep iff finite and inside both ranges; `(double)v > max || (double)v < min` is evaluated with the
// exactly-equivalent float limits computed on the host.
__device__ __forceinline__ bool pass_keep(const PreArgs& a, float x, float y, float z) {
    if (!finite_f32(x) || !finite_f32(y) || !finite_f32(z)) return false;
    if (z > a.z_hi || z < a.z_lo) return false;
    if (x > a.x_hi || x < a.x_lo) return false;
    return true;
}

// image row of input i: an exact multiply-high instead of an integer division when the host proved it exact
__device__ __forceinline__ int pre_row(const PreArgs& a, int i) {
    return a.w_magic ? (int)__umulhi((unsigned int)i, a.w_magic) : i / a.w;
}

// the 8 inputs of one thread -> points and keep mask (shared by the counting and the writing kernel)
template <int SRC>
__device__ __forceinline__ unsigned int pre_points(const PreArgs& a, int f, int first, int n_in, float (&px)[PRE_ITEMS], float (&py)[PRE_ITEMS],
                                                   float (&pz)[PRE_ITEMS]) {
    unsigned int keep = 0;
    if (SRC == 0) {
        const uint16_t* d = a.depth + (size_t)f * a.P;
        uint16_t dv[PRE_ITEMS];
        const bool full = first + PRE_ITEMS <= n_in;
        if (full && ((((size_t)f * a.P + first) & 7) == 0)) {
            const uint4 raw = *reinterpret_cast<const uint4*>(d + first);
            dv[0] = raw.x & 0xffff; dv[1] = raw.x >> 16; dv[2] = raw.y & 0xffff; dv[3] = raw.y >> 16;
            dv[4] = raw.z & 0xffff; dv[5] = raw.z >> 16; dv[6] = raw.w & 0xffff; dv[7] = raw.w >> 16;
        } else {
#pragma unroll
            for (int k = 0; k < PRE_ITEMS; ++k) dv[k] = (first + k < n_in) ? d[first + k] : 0;
        }
        int v = pre_row(a, first), u = first - v * a.w;
        // z = d*scale; x = z*((u-cx)/fx); y = z*((v-cy)/fy), all float, no contraction;
        // the two quotients come from tables filled with exactly these float operations.
        // depth-derived points are always finite: only the range tests remain
        if (full && u + PRE_ITEMS <= a.w) {   // the usual case: 8 pixels of one image row, all inside the input
            const float yrv = a.yr[v];
#pragma unroll
            for (int k = 0; k < PRE_ITEMS; ++k) {
                const float z = (float)dv[k] * a.depth_scale;
                px[k] = z * a.xr[u + k];
                py[k] = z * yrv;
                pz[k] = z;
                if (!(z > a.z_hi || z < a.z_lo) && !(px[k] > a.x_hi || px[k] < a.x_lo)) keep |= 1u << k;
            }
        } else {
            float yrv = a.yr[min(v, a.h - 1)];
#pragma unroll
            for (int k = 0; k < PRE_ITEMS; ++k) {
                const float z = (float)dv[k] * a.depth_scale;
                px[k] = z * a.xr[u];
                py[k] = z * yrv;
                pz[k] = z;
                if (first + k < n_in && !(z > a.z_hi || z < a.z_lo) && !(px[k] > a.x_hi || px[k] < a.x_lo)) keep |= 1u << k;
                if (++u == a.w) { u = 0; ++v; yrv = a.yr[min(v, a.h - 1)]; }
            }
        }
    } else {
        const unsigned char* b = a.blob + (size_t)f * a.P * a.point_step;
#pragma unroll
        for (int k = 0; k < PRE_ITEMS; ++k) {
            const int i = first + k;
            if (i < n_in) {
                const unsigned char* rec = b + (size_t)i * a.point_step;
                px[k] = *reinterpret_cast<const float*>(rec + a.xoff);
                py[k] = *reinterpret_cast<const float*>(rec + a.yoff);
                pz[k] = *reinterpret_cast<const float*>(rec + a.zoff);
                if (pass_keep(a, px[k], py[k], pz[k])) keep |= 1u << k;
            } else {
                px[k] = py[k] = pz[k] = 0.f;
            }
        }
    }
    return keep;
}

// pass 1: survivors per tile. A tile's output offset is then a plain sum over the preceding tile counts of its
// frame (no decoupled look-back, no spinning): depth is read twice (2 B/pixel), which is cheap next to the 16 B/point writes.
template <int SRC>
__global__ void __launch_bounds__(PRE_THREADS) k_pre_count(const PreArgs a) {
    const int f = blockIdx.y, t = blockIdx.x;
    const int n_in = (SRC == 1 && a.n_in) ? a.n_in[f] : a.P;
    const int first = t * PRE_TILE + threadIdx.x * PRE_ITEMS;
    float px[PRE_ITEMS], py[PRE_ITEMS], pz[PRE_ITEMS];
    const unsigned int keep = pre_points<SRC>(a, f, first, n_in, px, py, pz);
    int c = __popc(keep);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) c += __shfl_xor_sync(FULL_MASK, c, o);
    __shared__ int s_c[8];
    if ((threadIdx.x & 31) == 0) s_c[threadIdx.x >> 5] = c;
    __syncthreads();
    if (threadIdx.x == 0) {
        int tot = 0;
        for (int k = 0; k < 8; ++k) tot += s_c[k];
        a.tile_count[(size_t)f * a.tiles + t] = tot;
    }
}

template <int SRC>
__global__ void __launch_bounds__(PRE_THREADS) k_preprocess(const PreArgs a) {
    __shared__ float4 s_pts[PRE_TILE];
    __shared__ int s_w[9];
    __shared__ int s_base;
    __shared__ unsigned long long s_hash[8];
    __shared__ float s_mm[8][6];

    const int f = blockIdx.y, t = blockIdx.x;
    const int n_in = (SRC == 1 && a.n_in) ? a.n_in[f] : a.P;
    const int first = t * PRE_TILE + threadIdx.x * PRE_ITEMS;
    if (threadIdx.x < 32) {   // output offset of this tile = sum of the counts of the tiles before it
        int b = 0;
        for (int k = threadIdx.x; k < t; k += 32) b += a.tile_count[(size_t)f * a.tiles + k];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) b += __shfl_xor_sync(FULL_MASK, b, o);
        if (threadIdx.x == 0) s_base = b;
    }
    float px[PRE_ITEMS], py[PRE_ITEMS], pz[PRE_ITEMS];
    const unsigned int keep = pre_points<SRC>(a, f, first, n_in, px, py, pz);
    const int cnt = __popc(keep);
    int total;
    int pos = block_excl_scan256(cnt, s_w, &total);
#pragma unroll
    for (int k = 0; k < PRE_ITEMS; ++k)
        if (keep & (1u << k)) s_pts[pos++] = make_float4(px[k], py[k], pz[k], 1.0f);
    __syncthreads();
    const int gbase = s_base;

    float4* out = a.pts + (size_t)f * a.Pout;
    unsigned long long hsum = 0;
    float mn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f};
    float mx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
    for (int q = threadIdx.x; q < total; q += PRE_THREADS) {
        const float4 p = s_pts[q];
        out[gbase + q] = p;
        hsum += hash_point((unsigned int)(gbase + q), p.x, p.y, p.z);
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
    hsum = warp_sum_u64(hsum);
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) {
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(FULL_MASK, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(FULL_MASK, mx[c], o));
        }
    }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0) {
        s_hash[wid] = hsum;
#pragma unroll
        for (int c = 0; c < 3; ++c) { s_mm[wid][c] = mn[c]; s_mm[wid][3 + c] = mx[c]; }
    }
    __syncthreads();
    if (threadIdx.x == 0 && total > 0) {
        unsigned long long hs = 0;
        for (int i = 0; i < 8; ++i) hs += s_hash[i];
        atomic_add_u64(&a.res[f].points_hash, hs);
    }
    if (threadIdx.x < 6 && total > 0) {
        const int c = threadIdx.x;
        float v = s_mm[0][c];
        for (int i = 1; i < 8; ++i) v = (c < 3) ? fminf(v, s_mm[i][c]) : fmaxf(v, s_mm[i][c]);
        if (c < 3) atomicMin(&a.scr[f].mm[c], enc_f32(v)); else atomicMax(&a.scr[f].mm[c], enc_f32(v));
    }
    if (threadIdx.x == 0 && t == a.tiles - 1) a.res[f].n_points = gbase + total;
}

// all-points unprojection for cuboid_unproject (no filtering): out[i] = (x,y,z,1)
__global__ void k_unproject_all(const uint16_t* depth, int w, int n, float fx, float fy, float cx, float cy,
                                float scale, float4* out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int v = i / w, u = i - v * w;
    const float z = (float)depth[i] * scale;
    out[i] = make_float4(z * (((float)u - cx) / fx), z * (((float)v - cy) / fy), z, 1.0f);
}

__global__ void k_init_scratch(FrameScratch* scr, int n) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    FrameScratch s;
    s.mm[0] = s.mm[1] = s.mm[2] = 0xffffffffu;
    s.mm[3] = s.mm[4] = s.mm[5] = 0u;
    s.sort_bits = 0; s.overflow_mode = 0; s.best_count = 0; s.pad = 0;
    scr[i] = s;
}

}  // namespace cuboid
