// common.cuh — shared device helpers for libcuboid_cuda (sm_100a).
//
// The whole library is compiled with -fmad=false -prec-div=true -prec-sqrt=true -ftz=false: every float
// result that feeds a comparison, a floor or an arg-min must round exactly like the SSE2 (no-FMA) CPU
// reference path does. Do not add __fmaf_rn / fast-math intrinsics on such values.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "cuboid_cuda.h"

#define CUBOID_WARP 32
#define FULL_MASK 0xffffffffu

namespace cuboid {

// ---- per-frame device bookkeeping that never leaves the device ---------------------------------
struct FrameScratch {
    unsigned int mm[6];      // order-preserving encodings of min x,y,z / max x,y,z over passthrough survivors
    int sort_bits;           // significant bits of the voxel sort key
    int overflow_mode;       // 1: idx may be negative (dx*dy*dz > INT32_MAX) -> 32-bit biased sort
    int best_count;          // RANSAC best inlier count
    int pad;
};

struct Caps {
    int P;        // max points per frame (pixels or cloud points)
    int M;        // max remaining (non-plane) points per frame that clustering / ICP accept
    int tilesP;   // ceil(P / 2048)
    int tilesV;   // ceil(P / 1024)
};

__device__ __forceinline__ unsigned long long splitmix64(unsigned long long x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
__device__ __forceinline__ unsigned long long hash_point(unsigned int i, float x, float y, float z) {
    return splitmix64(splitmix64(((unsigned long long)i << 32) | __float_as_uint(x)) ^
                      (((unsigned long long)__float_as_uint(y) << 32) | __float_as_uint(z)));
}
__device__ __forceinline__ unsigned long long hash_index(unsigned int i, int v) {
    return splitmix64(((unsigned long long)i << 32) | (unsigned int)v);
}

__device__ __forceinline__ void atomic_add_u64(uint64_t* p, unsigned long long v) {
    atomicAdd(reinterpret_cast<unsigned long long*>(p), v);
}

// order-preserving float <-> uint (for atomicMin / atomicMax)
__device__ __forceinline__ unsigned int enc_f32(float f) {
    const unsigned int u = __float_as_uint(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float dec_f32(unsigned int e) {
    const unsigned int u = (e & 0x80000000u) ? (e & 0x7fffffffu) : ~e;
    return __uint_as_float(u);
}
__device__ __forceinline__ bool finite_f32(float v) { return (__float_as_uint(v) & 0x7f800000u) != 0x7f800000u; }

// Eigen SSE2 4-float dot: (a0b0 + a2b2) + (a1b1 + a3b3)   
__device__ __forceinline__ float dot4_sse(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3) {
    return (a0 * b0 + a2 * b2) + (a1 * b1 + a3 * b3);
}
__device__ __forceinline__ float plane_abs_dist(const float c[4], float x, float y, float z) {
    return fabsf(dot4_sse(c[0], c[1], c[2], c[3], x, y, z, 1.0f));
}

// ---- warp / block primitives ---------------------------------------------------------------------
__device__ __forceinline__ int warp_incl_scan(int v, int lane) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const int n = __shfl_up_sync(FULL_MASK, v, o);
        if (lane >= o) v += n;
    }
    return v;
}
// exclusive scan of one int per thread over a 256-thread block; returns exclusive prefix, *total = block sum.
// s_w must hold 9 ints. Contains two __syncthreads.
__device__ __forceinline__ int block_excl_scan256(int v, int* s_w, int* total) {
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int inc = warp_incl_scan(v, lane);
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (threadIdx.x == 0) {
        int run = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) { const int t = s_w[i]; s_w[i] = run; run += t; }
        s_w[8] = run;
    }
    __syncthreads();
    *total = s_w[8];
    return s_w[w] + inc - v;
}
// same for NT threads (NT/32 warps, up to 32); s_w must hold NT/32 + 1 ints
template <int NT>
__device__ __forceinline__ int block_excl_scan(int v, int* s_w, int* total) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int inc = warp_incl_scan(v, lane);
    if (lane == 31) s_w[w] = inc;
    __syncthreads();
    if (w == 0) {
        const int t = lane < NW ? s_w[lane] : 0;
        const int ti = warp_incl_scan(t, lane);
        if (lane < NW) s_w[lane] = ti - t;
        if (lane == 31) s_w[NW] = ti;
    }
    __syncthreads();
    *total = s_w[NW];
    return s_w[w] + inc - v;
}
__device__ __forceinline__ unsigned long long warp_sum_u64(unsigned long long v) {
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1) v += __shfl_xor_sync(FULL_MASK, v, o);
    return v;
}

// Decoupled look-back over the tiles of ONE frame. desc[] (zeroed before launch) holds
// [63:62] state (1 = tile aggregate, 2 = inclusive prefix) and [31:0] value. Called by one thread.
// Tiles must be claimed through an atomic ticket so that every predecessor is already running.
__device__ __forceinline__ int lookback_exclusive(unsigned long long* desc, int tile, int aggregate) {
    volatile unsigned long long* d = desc;
    if (tile == 0) {
        d[0] = (2ull << 62) | (unsigned int)aggregate;
        return 0;
    }
    d[tile] = (1ull << 62) | (unsigned int)aggregate;
    int excl = 0;
    int j = tile - 1;
    while (true) {
        const unsigned long long v = d[j];
        const unsigned int st = (unsigned int)(v >> 62);
        if (st == 0) continue;
        excl += (int)(unsigned int)v;
        if (st == 2) break;
        --j;
    }
    d[tile] = (2ull << 62) | (unsigned int)(excl + aggregate);
    return excl;
}

// Warp-parallel variant: called by all 32 lanes of ONE warp; each step inspects 32 predecessors at once, so a tile
// that starts while none of its predecessors has an inclusive prefix yet needs tile/32 round trips instead of `tile`.
__device__ __forceinline__ int lookback_exclusive_warp(unsigned long long* desc, int tile, int aggregate) {
    volatile unsigned long long* d = desc;
    const int lane = threadIdx.x & 31;
    if (tile == 0) {
        if (lane == 0) d[0] = (2ull << 62) | (unsigned int)aggregate;
        return 0;
    }
    if (lane == 0) d[tile] = (1ull << 62) | (unsigned int)aggregate;
    int excl = 0;
    int j = tile - 1;
    while (true) {
        const int idx = j - lane;
        const unsigned long long v = idx >= 0 ? d[idx] : (2ull << 62);   // virtual tile -1: inclusive prefix 0
        const unsigned int st = (unsigned int)(v >> 62);
        const unsigned int incl = __ballot_sync(FULL_MASK, st == 2);
        const unsigned int notready = __ballot_sync(FULL_MASK, st == 0);
        const int first = incl ? (__ffs(incl) - 1) : 31;
        const unsigned int upto = first == 31 ? 0xffffffffu : ((2u << first) - 1u);
        if (notready & upto) continue;   // somebody in the window has not published yet: look again
        int val = lane <= first ? (int)(unsigned int)v : 0;
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) val += __shfl_xor_sync(FULL_MASK, val, o);
        excl += val;
        if (incl) break;
        j -= 32;
    }
    if (lane == 0) d[tile] = (2ull << 62) | (unsigned int)(excl + aggregate);
    return excl;
}

// VoxelGrid geometry of one frame, recomputed by whoever needs it from the encoded min/max
// (pcl::VoxelGrid<PCLPointCloud2>::applyFilter: min_b_, div_b_, divb_mul_  — SURVEY.md A.2)
struct VoxelGeom {
    int min_b[3];
    int div_b[3];
    unsigned int mul1, mul2;
    float inv;
    int pcl_overflow;   // PCL's own dx*dy*dz > INT32_MAX warning test
    int sort_bits, overflow_mode;
};
__device__ __forceinline__ VoxelGeom voxel_geom(const FrameScratch& s, float inv) {
    VoxelGeom g;
    g.inv = inv;
    long long d64 = 1, dv = 1;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const float mn = dec_f32(s.mm[a]), mx = dec_f32(s.mm[3 + a]);
        g.min_b[a] = (int)floorf(mn * inv);
        const int max_b = (int)floorf(mx * inv);
        g.div_b[a] = max_b - g.min_b[a] + 1;
        d64 *= (long long)((mx - mn) * inv) + 1;
        dv *= (long long)g.div_b[a];
    }
    g.mul1 = (unsigned int)g.div_b[0];
    g.mul2 = (unsigned int)g.div_b[0] * (unsigned int)g.div_b[1];
    g.pcl_overflow = d64 > 2147483647ll;
    g.overflow_mode = dv > 2147483647ll;
    g.sort_bits = g.overflow_mode ? 32 : (dv <= 1 ? 0 : 64 - __clzll(dv - 1));
    return g;
}
__device__ __forceinline__ int voxel_index(const VoxelGeom& g, float x, float y, float z) {
    const int i0 = (int)(floorf(x * g.inv) - (float)g.min_b[0]);
    const int i1 = (int)(floorf(y * g.inv) - (float)g.min_b[1]);
    const int i2 = (int)(floorf(z * g.inv) - (float)g.min_b[2]);
    return (int)((unsigned int)i0 + (unsigned int)i1 * g.mul1 + (unsigned int)i2 * g.mul2);
}

}  // namespace cuboid
