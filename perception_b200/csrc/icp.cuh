// icp.cuh — stage 4: pcl::IterativeClosestPoint<PointXYZ,PointXYZ>::align + getFitnessScore
// (icp.cpp:170-182, opd.cpp:220-235). The whole ICP loop — nearest neighbours, Umeyama, 3x3 Jacobi SVD,
// incremental transform, DefaultConvergenceCriteria — runs on the device, no host round trip.
//
//   execution           k_icp_init builds every (frame, cluster, guess) problem; k_icp is a PERSISTENT kernel, one
//                       1024-thread CTA per SM hosting 1024 / SUB sub-workers (SUB = 256, 512 or 1024 threads, named
//                       barriers) around ONE shared-memory copy of the template, its BVH, the sibling chains and the
//                       original-index table. A sub-worker pops a problem from a ring queue, runs a time slice of
//                       iterations and either finishes it or pushes it back.
//   nearest neighbour   exact search over a bounding-volume hierarchy of the template (kd split on the host into 16-point
//                       leaves stored SoA, nodes in depth-first order, fp16 boxes rounded OUTWARD). A subtree is skipped
//                       only if its box's lower bound — evaluated with the same un-fused rounding sequence as the distance
//                       itself, hence a true bit-level bound — is strictly above the running minimum. The search runs
//                       OUTWARD from the previous iteration's correspondent: its leaf first, then the sibling subtrees of
//                       that leaf's path to the root are box-tested by all lanes in lockstep. What survives is NOT walked
//                       lane by lane: (query, subtree) items go on a per-warp work list in shared memory that the warp
//                       drains 32 items at a time — node tests and leaf scans always run with full warps, whichever
//                       queries they belong to (icp_nn_pass, QUEUED). Leaves are scanned brute force,
//                       d2 = ((dx*dx)+dy*dy)+dz*dz; ties resolve to the LOWEST ORIGINAL template index, i.e. exactly what a
//                       brute-force scan in template order with strict '<' returns (the canonical tie rule).
//                       A work list that would overflow (queries far from the template) falls back to the per-lane
//                       stackless walk, which is also the path for templates too large for shared memory.
//   Umeyama             canonical 256-lane strided partial sums + xor-butterfly + 8 warp partials left to right
//                       (identical in oracle/cuboid_oracle.cpp: canon_reduce), Eigen JacobiSVD restated
//   convergence         max iterations | transform epsilon | |dMSE| < 1e-12 | rel dMSE < icp_fitness_score
//
// Roofline: FP32 pipe (un-fused). Brute force is 8*S*T ops per pass; the culled kernel reports both the
// pairs it actually evaluated and the brute-force equivalent (IcpArgs::work).
#pragma once
#include <type_traits>
#include <cuda_fp16.h>

#include "common.cuh"
#include "nn_table.cuh"
#include "ransac.cuh"   // TMA bulk-copy helpers

namespace cuboid {

struct IcpOut {
    float T[16];
    double fitness;
    int converged, iters, state, pad;
    unsigned long long corr_hash;
};

struct IcpArgs {
    const float4* remain;    // [F][P]
    const int* idx_sorted;   // [F][M]
    const int* offsets;      // [F][KC+1]
    const float* tmpl;       // [nleaf][3][16] kd-ordered 16-point leaves, SoA per leaf (x[16] y[16] z[16]); far sentinels pad the tail
    const int* tmpl_orig;    // [Tpad] original template index of every kd-ordered position (sentinels: INT_MAX)
    const uint4* nodes;      // [nnodes] depth-first BVH, 16 B each: fp16 AABB rounded OUTWARD (lo down, hi up) + link
                             //          (link >= 0: inner node, index of the first node after its subtree; link < 0: leaf ~link)
    int T, Tpad, nleaf, nnodes;
    // Sibling chains (resident templates, staged after the template when they fit): for every leaf the roots of the subtrees
    // hanging off its path to the root, deepest first, 0xffff-padded to sib_max entries. Leaf + these subtrees = the whole
    // tree, so a search can start at the seed's leaf and work outwards instead of descending from the root.
    const unsigned short* sib;     // [nleaf][sib_max]
    int sib_on, sib_max, sib_bytes;
    const unsigned short* orig16;  // [Tpad] original template index of every kd-ordered position as u16 (queued search; Tpad <= 65536)
    NnTableView tab;               // nearest-neighbour candidate table of the template (nn_table.cuh); used when tmode = 1
    int tmode;
    int* miss;                     // [F][G][M] queries the table could not answer, per problem (visited by the BVH search afterwards)
    unsigned long long* stats;     // developer counters (only written when compiled with -DCUBOID_ICP_STATS), else ignored
    int local_cap;                 // > 0: a sub-worker keeps the working set of a problem with at most this many points in shared memory (one sub-worker per CTA)
    int qmode;                     // 1: queued outward search (needs resident template, sibling chains, orig16 and the per-warp scratch)
    const float* guesses;    // n_guess * (16 | 9) or NULL
    int n_guess, guess_mode;
    float4* cur;             // [F][G][M]
    int* corr;               // [F][G][M]  correspondences as positions in the kd-ordered template
    float* cd;               // [F][G][M]
    int* order;              // [F][G][M]  Morton order of the source points (nearest-neighbour visiting order only)
    IcpOut* out;             // [F][MAXC][G]
    cuboid_frame_result* res;
    int P, M, KC;
    int max_iter;
    double rot_thr, trans_thr, rel_mse, abs_thr;
    int resident;            // 1: the whole template sits in shared memory; 0: chunks are read through L1/L2
    int cull;                // 1: skip subtrees whose AABB lower bound exceeds the running minimum (exact)
    unsigned long long* work;   // [2] += (source-template pairs evaluated, brute-force pairs S*T per pass) or NULL
    int* corr_trace; float* T_trace; int cap_trace;   // debug taps for problem (0,0,0)
    // persistent, time-sliced execution (see k_icp)
    struct IcpState* pstate;   // [problem slots of the launch]
    struct IcpQueue* queue; struct IcpSlot* ring; int n_slots;
    int slice_iters;           // iterations per time slice
    int crew;                  // worker CTAs of k_icp
    int nsub;                  // sub-workers per CTA (1 or 2)
    int init_smem;             // dynamic shared memory of k_icp_init (Morton sort window)
    int hashes;                // 1: accumulate corr_hash (parity tap); 0: leave it 0
    double max_d2;             // setMaxCorrespondenceDistance squared (k_icp MODE 4 only: pairs beyond it are dropped, icp.cpp:175)
};

constexpr int ICP_THREADS = 512;   // k_icp_init's CTA size
constexpr int ICP_NT = 1024;       // k_icp's CTA size: one CTA per SM, 1024 / SUB sub-workers
constexpr int ICP_LANES = 256;
// A CTA of k_icp hosts ICP_NT / SUB independent sub-workers of SUB threads (template parameter: 256, 512 or 1024). A sub-worker
// serves its own problem and synchronises on its own named barrier; all share the CTA's one copy of the template and tree.
// SUB = 256: four problems per SM - while one sits in its serial SVD or in a reduction the others keep the SM busy.
// SUB = 512 / 1024: more threads per problem, for launches with fewer problems than sub-workers (sub-chunks, single frames),
// where latency per problem is what counts. (128-thread sub-workers were measured slower: the reductions take twice the rounds.)
template <int SUB>
__device__ __forceinline__ void sub_sync(int sub) {   // literal barrier ids: a register id would reserve all 16 barriers
    switch (sub) {
        case 0: asm volatile("bar.sync 1, %0;" ::"n"(SUB) : "memory"); break;
        case 1: asm volatile("bar.sync 2, %0;" ::"n"(SUB) : "memory"); break;
        case 2: asm volatile("bar.sync 3, %0;" ::"n"(SUB) : "memory"); break;
        case 3: asm volatile("bar.sync 4, %0;" ::"n"(SUB) : "memory"); break;
        case 4: asm volatile("bar.sync 5, %0;" ::"n"(SUB) : "memory"); break;
        case 5: asm volatile("bar.sync 6, %0;" ::"n"(SUB) : "memory"); break;
        case 6: asm volatile("bar.sync 7, %0;" ::"n"(SUB) : "memory"); break;
        default: asm volatile("bar.sync 8, %0;" ::"n"(SUB) : "memory"); break;
    }
}
constexpr int ICP_LEAF = 16;       // template points per BVH leaf


struct M3f { float a[3][3]; };
struct Rotf { float c, s; };

__device__ __forceinline__ void rot_rows(M3f& m, int p, int q, Rotf j) {
    if (j.c == 1.f && j.s == 0.f) return;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float xi = m.a[p][i], yi = m.a[q][i];
        m.a[p][i] = j.c * xi + j.s * yi;
        m.a[q][i] = -j.s * xi + j.c * yi;
    }
}
__device__ __forceinline__ void rot_cols(M3f& m, int p, int q, Rotf j) {
    const Rotf t{j.c, -j.s};
    if (t.c == 1.f && t.s == 0.f) return;
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float xi = m.a[i][p], yi = m.a[i][q];
        m.a[i][p] = t.c * xi + t.s * yi;
        m.a[i][q] = -t.s * xi + t.c * yi;
    }
}
__device__ __forceinline__ Rotf make_jacobi(float x, float y, float z) {
    const float deno = 2.0f * fabsf(y);
    if (deno < 1.17549435e-38f) return Rotf{1.f, 0.f};
    const float tau = (x - z) / deno;
    const float w = sqrtf(tau * tau + 1.0f);
    float t;
    if (tau > 0.f) t = 1.0f / (tau + w); else t = 1.0f / (tau - w);
    const float sign_t = t > 0.f ? 1.0f : -1.0f;
    const float n = 1.0f / sqrtf(t * t + 1.0f);
    Rotf r;
    r.s = -sign_t * (y / fabsf(y)) * fabsf(t) * n;
    r.c = n;
    return r;
}
__device__ __forceinline__ void jacobi_2x2(const M3f& w, int p, int q, Rotf* jl, Rotf* jr) {
    float m00 = w.a[p][p], m01 = w.a[p][q], m10 = w.a[q][p], m11 = w.a[q][q];
    Rotf rot1;
    const float t = m00 + m11;
    const float d = m10 - m01;
    if (fabsf(d) < 1.17549435e-38f) { rot1.s = 0.f; rot1.c = 1.f; }
    else {
        const float u = t / d;
        const float tmp = sqrtf(1.0f + u * u);
        rot1.s = 1.0f / tmp;
        rot1.c = u / tmp;
    }
    if (!(rot1.c == 1.f && rot1.s == 0.f)) {
        const float x0 = m00, x1 = m01, y0 = m10, y1 = m11;
        m00 = rot1.c * x0 + rot1.s * y0; m01 = rot1.c * x1 + rot1.s * y1;
        m10 = -rot1.s * x0 + rot1.c * y0; m11 = -rot1.s * x1 + rot1.c * y1;
    }
    *jr = make_jacobi(m00, m01, m11);
    const Rotf jt{jr->c, -jr->s};
    jl->c = rot1.c * jt.c - rot1.s * jt.s;
    jl->s = rot1.c * jt.s + rot1.s * jt.c;
}
__device__ __forceinline__ float det3(const M3f& m) {
    const float h0 = m.a[0][0] * (m.a[1][1] * m.a[2][2] - m.a[1][2] * m.a[2][1]);
    const float h1 = m.a[0][1] * (m.a[1][0] * m.a[2][2] - m.a[1][2] * m.a[2][0]);
    const float h2 = m.a[0][2] * (m.a[1][0] * m.a[2][1] - m.a[1][1] * m.a[2][0]);
    return h0 - h1 + h2;
}
// Eigen::JacobiSVD<Matrix3f>(sigma, ComputeFullU | ComputeFullV)
__device__ __forceinline__ void jacobi_svd3(const M3f& in, M3f& U, M3f& V) {
    const float precision = 2.0f * 1.1920928955078125e-07f;
    const float consider_zero = 1.17549435e-38f;
    float scale = 0.f;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) scale = fmaxf(scale, fabsf(in.a[i][j]));
    if (scale == 0.f) scale = 1.f;
    M3f W;
#pragma unroll
    for (int i = 0; i < 3; ++i)
#pragma unroll
        for (int j = 0; j < 3; ++j) { W.a[i][j] = in.a[i][j] / scale; U.a[i][j] = (i == j) ? 1.f : 0.f; V.a[i][j] = (i == j) ? 1.f : 0.f; }
    float max_diag = fmaxf(fabsf(W.a[0][0]), fmaxf(fabsf(W.a[1][1]), fabsf(W.a[2][2])));
    bool finished = false;
    int guard = 0;
    while (!finished && guard++ < 1000) {
        finished = true;
#pragma unroll
        for (int p = 1; p < 3; ++p) {
#pragma unroll
            for (int q = 0; q < 2; ++q) {
                if (q >= p) continue;
                const float thr = fmaxf(consider_zero, precision * max_diag);
                if (fabsf(W.a[p][q]) > thr || fabsf(W.a[q][p]) > thr) {
                    finished = false;
                    Rotf jl, jr;
                    jacobi_2x2(W, p, q, &jl, &jr);
                    rot_rows(W, p, q, jl);
                    rot_cols(U, p, q, Rotf{jl.c, -jl.s});
                    rot_cols(W, p, q, jr);
                    rot_cols(V, p, q, jr);
                    max_diag = fmaxf(max_diag, fmaxf(fabsf(W.a[p][p]), fabsf(W.a[q][q])));
                }
            }
        }
    }
    float sv[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        const float d = W.a[i][i];
        sv[i] = fabsf(d);
        if (d < 0.f) { U.a[0][i] = -U.a[0][i]; U.a[1][i] = -U.a[1][i]; U.a[2][i] = -U.a[2][i]; }
    }
#pragma unroll
    for (int i = 0; i < 3; ++i) sv[i] = sv[i] * scale;
    // Selection sort of the singular values, largest first (first maximum on ties; stops at a zero maximum), columns of U and V
    // swapped along. Written with static indices only: a run-time column index would put U, V and W in local memory for the whole
    // routine, and every rotation above would go through it (the SVD is the serial part of an ICP iteration).
    auto swap_cols = [&](auto I, auto J) {
        constexpr int i = decltype(I)::value, j = decltype(J)::value;
        float t = sv[i]; sv[i] = sv[j]; sv[j] = t;
#pragma unroll
        for (int r = 0; r < 3; ++r) {
            t = U.a[r][i]; U.a[r][i] = U.a[r][j]; U.a[r][j] = t;
            t = V.a[r][i]; V.a[r][i] = V.a[r][j]; V.a[r][j] = t;
        }
    };
    using I0 = std::integral_constant<int, 0>; using I1 = std::integral_constant<int, 1>; using I2 = std::integral_constant<int, 2>;
    {
        const bool p1 = sv[1] > sv[0];
        const float m01 = p1 ? sv[1] : sv[0];
        const bool p2 = sv[2] > m01;
        if ((p2 ? sv[2] : m01) == 0.f) return;
        if (p2) swap_cols(I0{}, I2{}); else if (p1) swap_cols(I0{}, I1{});
    }
    {
        const bool p2 = sv[2] > sv[1];
        if ((p2 ? sv[2] : sv[1]) == 0.f) return;
        if (p2) swap_cols(I1{}, I2{});
    }
}

// tr * (x,y,z,1) = ((m0*x + m1*y) + m2*z) + m3   (Eigen 4x4 * 4x1)
__device__ __forceinline__ float4 xform(const float* m, const float4 p) {
    float4 o;
    o.x = ((m[0] * p.x + m[1] * p.y) + m[2] * p.z) + m[3];
    o.y = ((m[4] * p.x + m[5] * p.y) + m[6] * p.z) + m[7];
    o.z = ((m[8] * p.x + m[9] * p.y) + m[10] * p.z) + m[11];
    o.w = 1.0f;
    return o;
}

// canonical block reduction of NQ quantities held one-per-lane by the first 256 threads (lane partials are
// already the sequential strided sums). Result for quantity q lands in s_out[q]. Contains two __syncthreads.
template <typename Tq, int NQ>
__device__ __forceinline__ void canon_block_reduce(Tq (&v)[NQ], Tq* s_part /* [NQ][8] */, Tq* s_out) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        Tq x = v[q];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) x = x + __shfl_xor_sync(FULL_MASK, x, o);
        if (lane == 0 && wid < 8) s_part[q * 8 + wid] = x;
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        const Tq* p = s_part + threadIdx.x * 8;
        Tq s = p[0];
#pragma unroll
        for (int g = 1; g < 8; ++g) s = s + p[g];
        s_out[threadIdx.x] = s;
    }
    __syncthreads();
}


// the same for one sub-worker of SUB threads: canon_sub_partial after each of its lane sets (v = the partial sums of canonical
// lane tid + set * SUB; threads beyond lane 255 carry zeros and are ignored), then canon_sub_finish
template <typename Tq, int NQ, int SUB>
__device__ __forceinline__ void canon_sub_partial(Tq (&v)[NQ], Tq* s_part /* [NQ][8] */, int tid, int set) {
    const int lane = tid & 31, wid = tid >> 5;
    if (SUB > ICP_LANES && wid >= ICP_LANES / 32) return;   // (warp-uniform) only the canonical warps hold anything: the others would shuffle zeros
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        Tq x = v[q];
#pragma unroll
        for (int o = 16; o >= 1; o >>= 1) x = x + __shfl_xor_sync(FULL_MASK, x, o);
        if (lane == 0 && (SUB <= ICP_LANES || wid < 8)) s_part[q * 8 + set * (SUB / 32) + wid] = x;   // canonical warp = lanes [32w, 32w + 32)
    }
}
template <typename Tq, int NQ, int SUB>
__device__ __forceinline__ void canon_sub_finish(Tq* s_part /* [NQ][8] */, Tq* s_out, int tid, int sub) {
    sub_sync<SUB>(sub);
    if (tid < NQ) {
        const Tq* p = s_part + tid * 8;
        Tq s = p[0];
#pragma unroll
        for (int g = 1; g < 8; ++g) s = s + p[g];
        s_out[tid] = s;
    }
    sub_sync<SUB>(sub);
}

// both at once (the means / MSE step): one barrier pair instead of two
template <int NQ, int SUB>
__device__ __forceinline__ void canon_sub_finish_fd(float* s_part_f, float* s_out_f, double* s_part_d, double* s_out_d, int tid, int sub) {
    sub_sync<SUB>(sub);
    if (tid < NQ) {
        const float* p = s_part_f + tid * 8;
        float s = p[0];
#pragma unroll
        for (int g = 1; g < 8; ++g) s = s + p[g];
        s_out_f[tid] = s;
    } else if (tid == 32) {
        double s = s_part_d[0];
#pragma unroll
        for (int g = 1; g < 8; ++g) s = s + s_part_d[g];
        s_out_d[0] = s;
    }
    sub_sync<SUB>(sub);
}

struct IcpShared {
    unsigned long long bar;
    float part_f[16 * 8];
    double part_d[8];
    float red_f[16];
    double red_d[2];
    float Tm[16];       // this iteration's transformation_
    float fin[16];      // final_transformation_
    float guess[16];
    int done, converged, state, iters;
    int task;           // dynamic task counter of the nearest-neighbour pass
    int nmiss;          // queries of the current pass the candidate table could not answer
    int pend;           // 1: Tm has not been applied to the working cloud yet (the next table pass does it while it reads the points)
    double prev_mse;
};

__device__ __forceinline__ float dist2(float sx, float sy, float sz, float tx, float ty, float tz) {
    const float dx = sx - tx, dy = sy - ty, dz = sz - tz;
    return ((dx * dx) + dy * dy) + dz * dz;   // FLANN L2_Simple order, un-fused
}
constexpr int ICP_LEAF_FLOATS = 3 * ICP_LEAF;   // one SoA leaf
__device__ __forceinline__ float3 tmpl_point(const float* soa, int pos) {
    const float* lf = soa + (size_t)(pos / ICP_LEAF) * ICP_LEAF_FLOATS + (pos % ICP_LEAF);
    return make_float3(lf[0], lf[ICP_LEAF], lf[2 * ICP_LEAF]);
}
// Lower bound of dist2(s, t) over every t inside the box [lo, hi], evaluated with the same rounding sequence:
// float subtraction, squaring and addition are monotone, so lb <= dist2(s,t) holds bit-for-bit.
__device__ __forceinline__ float box_lb(float sx, float sy, float sz, const float4 lo, const float4 hi) {
    const float dx = fmaxf(fmaxf(lo.x - sx, sx - hi.x), 0.f);
    const float dy = fmaxf(fmaxf(lo.y - sy, sy - hi.y), 0.f);
    const float dz = fmaxf(fmaxf(lo.z - sz, sz - hi.z), 0.f);
    return ((dx * dx) + dy * dy) + dz * dz;
}

// fp16 box -> float lower bound. The box was widened outward when it was rounded to fp16, so it still contains every
// point of the subtree and box_lb stays a true lower bound.
__device__ __forceinline__ float node_lb(float sx, float sy, float sz, const uint4 nd) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&nd.x));
    const float2 b = __half22float2(*reinterpret_cast<const __half2*>(&nd.y));
    const float2 c = __half22float2(*reinterpret_cast<const __half2*>(&nd.z));
    return box_lb(sx, sy, sz, make_float4(a.x, a.y, b.x, 0.f), make_float4(b.y, c.x, c.y, 0.f));
}

// One nearest-neighbour pass over cur[0..S): writes corr[] (kd-ordered template position) and cd[].
// On entry corr[] holds a valid template position per point (the previous pass's answer, or 0): its distance
// seeds the running minimum so the culling is tight from the first node.
//
// Exactness: a subtree is skipped only when lb > best (strict), so no subtree holding a minimiser or a tie is
// ever skipped; among equal distances the LOWEST ORIGINAL template index wins — the answer of a brute-force
// scan in original order with strict '<'.
struct IcpBest { float d; int pos; int orig; };   // orig < 0: original index of pos not fetched yet
// brute-force scan of one 16-point leaf; among equal distances the LOWEST ORIGINAL template index wins
__device__ __forceinline__ void icp_scan_leaf(const IcpArgs& a, const float* tp, int leaf, float sx, float sy, float sz, IcpBest& b) {
    const float* lf = tp + (size_t)leaf * ICP_LEAF_FLOATS;
    const int pbase = leaf * ICP_LEAF;
#pragma unroll
    for (int jj = 0; jj < ICP_LEAF; jj += 4) {
        const float4 X = *reinterpret_cast<const float4*>(lf + jj);
        const float4 Y = *reinterpret_cast<const float4*>(lf + ICP_LEAF + jj);
        const float4 Z = *reinterpret_cast<const float4*>(lf + 2 * ICP_LEAF + jj);
        const float d0 = dist2(sx, sy, sz, X.x, Y.x, Z.x);
        const float d1 = dist2(sx, sy, sz, X.y, Y.y, Z.y);
        const float d2 = dist2(sx, sy, sz, X.z, Y.z, Z.z);
        const float d3 = dist2(sx, sy, sz, X.w, Y.w, Z.w);
        if (fminf(fminf(d0, d1), fminf(d2, d3)) <= b.d) {   // rare once the seed is good
            const float dd[4] = {d0, d1, d2, d3};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const int pos = pbase + jj + q;
                if (dd[q] < b.d) { b.d = dd[q]; b.pos = pos; b.orig = -1; }
                else if (dd[q] == b.d && pos != b.pos) {
                    // an exact tie between two different points: only now are the original indices needed (global memory)
                    if (b.orig < 0) b.orig = a.tmpl_orig[b.pos];
                    const int o = a.tmpl_orig[pos];
                    if (o < b.orig) { b.orig = o; b.pos = pos; }
                }
            }
        }
    }
}

// Working-set loads of the ICP loop. In global memory cur / corr / cd travel between SMs from one time slice to the next, so they are
// read at L2 (__ldcg: this SM's L1 may hold lines from an earlier slice). LOCAL = the slice keeps them in shared memory (launches
// with one sub-worker per CTA and a cluster small enough: single-frame latency), where a plain load is right.
template <bool LOCAL, typename T>
__device__ __forceinline__ T icp_ld(const T* p) { if (LOCAL) return *p; else return __ldcg(p); }

// ---- queued outward search: per-warp scratch in shared memory -----------------------------------------------------------
constexpr int ICP_QCAP = 224;      // node work list entries per warp
constexpr int ICP_QLEAF = 64;      // leaf work list entries per warp (a leaf round runs as soon as 32 are waiting)
struct IcpWarpScr {
    unsigned long long key[32];    // per query of the task: (bits of the best d2) << 32 | original index << 16 | kd position
    float px[32], py[32], pz[32];  // the queries
    unsigned int nodes[ICP_QCAP];  // (owner lane << 16) | node index: subtrees whose box may still hold a closer point
    unsigned int leaves[ICP_QLEAF];// (owner lane << 16) | leaf index: leaves whose box survived, waiting for a full-warp scan
};
static_assert(sizeof(IcpWarpScr) % 16 == 0, "per-warp scratch keeps 16-byte alignment");

// Brute-force scan of one 16-point leaf, branch-free: a tournament over the 16 distances carries the arg-min along.
// Returns (bits of min d2) << 32 | original index << 16 | kd position. d2 >= +0, so its bit pattern orders like its value and
// one 64-bit unsigned minimum over such keys is "lowest distance, then lowest ORIGINAL template index" - the canonical tie rule.
// Two equal distances inside the leaf are rare (exact float ties); any equal pair met by the tournament sends the lane through
// the exact loop below (false positives only cost time).
__device__ __forceinline__ unsigned long long icp_leaf_key(const float* tp, const unsigned short* s_orig, int leaf, float sx, float sy, float sz) {
    const float* lf = tp + (size_t)leaf * ICP_LEAF_FLOATS;
    float d[ICP_LEAF];
#pragma unroll
    for (int jj = 0; jj < ICP_LEAF; jj += 4) {
        const float4 X = *reinterpret_cast<const float4*>(lf + jj);
        const float4 Y = *reinterpret_cast<const float4*>(lf + ICP_LEAF + jj);
        const float4 Z = *reinterpret_cast<const float4*>(lf + 2 * ICP_LEAF + jj);
        d[jj + 0] = dist2(sx, sy, sz, X.x, Y.x, Z.x);
        d[jj + 1] = dist2(sx, sy, sz, X.y, Y.y, Z.y);
        d[jj + 2] = dist2(sx, sy, sz, X.z, Y.z, Z.z);
        d[jj + 3] = dist2(sx, sy, sz, X.w, Y.w, Z.w);
    }
    float m[ICP_LEAF];
    int ix[ICP_LEAF];
    bool tie = false;
#pragma unroll
    for (int k = 0; k < ICP_LEAF; ++k) { m[k] = d[k]; ix[k] = k; }
#pragma unroll
    for (int w = ICP_LEAF / 2; w >= 1; w >>= 1)
#pragma unroll
        for (int k = 0; k < w; ++k) {      // merge (2k, 2k+1): strict '<' keeps the lower kd position on equality
            const float x = m[2 * k], y = m[2 * k + 1];
            const bool lt = y < x;
            tie = tie || (y == x);
            m[k] = lt ? y : x;
            ix[k] = lt ? ix[2 * k + 1] : ix[2 * k];
        }
    const float best = m[0];
    int j = ix[0];
    const int pbase = leaf * ICP_LEAF;
    unsigned int o = s_orig[pbase + j];
    if (tie) {   // exact: lowest original index among the points of this leaf at the minimal distance
#pragma unroll
        for (int q = 0; q < ICP_LEAF; ++q)
            if (d[q] == best) {
                const unsigned int oq = s_orig[pbase + q];
                if (oq < o) { o = oq; j = q; }
            }
    }
    return ((unsigned long long)__float_as_uint(best) << 32) | (unsigned long long)((o << 16) | (unsigned int)(pbase + j));
}

// Per-lane stackless walk of the depth-first node array [node, end) plus the sibling subtrees in `cand` (bit e = entry e of the
// sibling chain sl), deepest first. Lanes reconverge before every leaf scan so the scans issue together.
template <bool RESIDENT>
__device__ __forceinline__ void icp_walk(const IcpArgs& a, const float* tp, const uint4* s_nodes, const unsigned short* sl, unsigned int cand,
                                         int node, int end, bool cull, float sx, float sy, float sz, IcpBest& b, unsigned int& nleaf_eval) {
    while (true) {
        // every lane walks to its next surviving leaf
        int leaf = -1;
        while (true) {
            if (node >= end) {
                if (!cand) break;
                const int e = __ffs(cand) - 1;
                cand &= cand - 1u;
                const int sn = sl[e];
                const uint4 nd = s_nodes[sn];
                if (node_lb(sx, sy, sz, nd) > b.d) continue;   // the minimum may have tightened since the lockstep test
                const int link = (int)nd.w;
                if (link < 0) { leaf = ~link; break; }           // the sibling is a leaf
                node = sn + 1; end = link;
                continue;
            }
            const uint4 nd = s_nodes[node];
            const int link = (int)nd.w;
            if (cull && node_lb(sx, sy, sz, nd) > b.d) { node = link >= 0 ? link : node + 1; continue; }
            ++node;
            if (link < 0) { leaf = ~link; break; }
        }
        // reconverge, so the leaf scans of all lanes issue together
        if (!__any_sync(FULL_MASK, leaf >= 0)) break;
        if (leaf >= 0) { ++nleaf_eval; icp_scan_leaf(a, tp, leaf, sx, sy, sz, b); }
        __syncwarp();
    }
}

// One nearest-neighbour pass over cur[0..S): writes corr[] (kd-ordered template position) and cd[].
// On entry corr[] holds a valid template position per point (the previous pass's answer, or any position): the search
// starts there, so the running minimum is tight from the first box test.
//
// Exactness: a subtree is skipped only when lb > best (strict), so no subtree holding a minimiser or a tie is
// ever skipped; among equal distances the LOWEST ORIGINAL template index wins - the answer of a brute-force
// scan in original order with strict '<'.
//
// With sibling chains the search runs OUTWARDS from the seed: the seed's own leaf is scanned first, then the subtrees
// hanging off the leaf's path to the root are box-tested by all lanes in lockstep (a fixed, divergence-free loop of
// <= sib_max tests, about 9). Leaf + those subtrees = the whole tree, so nothing is left out. Once ICP has roughly aligned
// the clouds almost every sibling test fails.
//   QUEUED  the survivors of all 32 queries go on the warp's work list (IcpWarpScr) as (owner lane, node) items. The warp
//           pops up to 32 items at a time: each lane box-tests ONE node against its owner's current minimum; surviving inner
//           nodes push their two children, surviving leaves go on the leaf list, which is scanned 32 leaves at a time, each
//           result merged into the owner's key by a 64-bit shared-memory atomicMin. Every instruction of the drain runs with
//           (nearly) all lanes, however unevenly the survivors are spread over the queries - the per-lane walk below ran at
//           4 of 32 lanes. A list that would overflow (queries still far from the template: many boxes survive) abandons the
//           lists and finishes the task with the per-lane walk from the root, seeded with the best found so far.
//   else    every lane walks its own survivors (icp_walk).
template <bool RESIDENT, bool QUEUED, bool LOCAL>
__device__ __forceinline__ void icp_nn_pass(const IcpArgs& a, IcpShared& sh, const float* s_tmpl, const uint4* s_nodes, const unsigned short* s_sib,
                                            const unsigned short* s_orig, IcpWarpScr* wscr, const float4* cur, int S, const int* order, int* corr,
                                            float* cd, unsigned long long& evaluated) {
    const int lane = threadIdx.x & 31;
    const float* tp = RESIDENT ? s_tmpl : a.tmpl;
    const int ntask = (S + 31) / 32;
    const bool cull = a.cull != 0;
    const bool outward = RESIDENT && cull && a.sib_on;
    const int nnodes = a.nnodes;
    while (true) {
        int task = 0;
        if (lane == 0) task = atomicAdd(&sh.task, 1);
        task = __shfl_sync(FULL_MASK, task, 0);
        if (task >= ntask) break;
        const int k = task * 32 + lane;
        const bool valid = k < S;
        const int i = order[valid ? k : S - 1];   // 32 lanes = 32 spatial neighbours: coherent tree walks, broadcast loads
        const float4 p = icp_ld<LOCAL>(cur + i);
        const float sx = p.x, sy = p.y, sz = p.z;
        int pos0 = icp_ld<LOCAL>(corr + i);
        unsigned int nleaf_eval = 0;
        if (QUEUED) {
            IcpWarpScr& ws = *wscr;
            if (a.tmode && a.tab.seed) {
                // these are the queries the candidate table could not answer: far from the template, typically because the cloud has
                // just moved a lot and the previous correspondent is a poor seed. The seed grid names a template point near the
                // query's coarse cell (clamped to the grid); start from whichever of the two is closer. Any position is a valid seed.
                const NnTableView& T = a.tab;
                const int gx = min(max((int)floorf((sx - T.sorg[0]) * T.sinv_h), 0), T.snx - 1);
                const int gy = min(max((int)floorf((sy - T.sorg[1]) * T.sinv_h), 0), T.sny - 1);
                const int gz = min(max((int)floorf((sz - T.sorg[2]) * T.sinv_h), 0), T.snz - 1);
                const int pos1 = (int)__ldg(T.seed + ((size_t)gz * T.sny + gy) * T.snx + gx);
                const float3 t0 = tmpl_point(tp, pos0), t1 = tmpl_point(tp, pos1);
                if (dist2(sx, sy, sz, t1.x, t1.y, t1.z) < dist2(sx, sy, sz, t0.x, t0.y, t0.z)) pos0 = pos1;
            }
            const int L = pos0 / ICP_LEAF;
            const unsigned short* sl = s_sib + L * a.sib_max;
            ++nleaf_eval;
            unsigned long long key = icp_leaf_key(tp, s_orig, L, sx, sy, sz);
            unsigned int cand = 0u;
            {
                const float bd = __uint_as_float((unsigned int)(key >> 32));
                for (int e = 0; e < a.sib_max; ++e) {
                    const unsigned int sn = sl[e];
                    if (sn != 0xffffu && !(node_lb(sx, sy, sz, s_nodes[sn]) > bd)) cand |= 1u << e;
                }
            }
            if (!valid) cand = 0u;   // padding lanes repeat the last query: no work-list items for them
#ifdef CUBOID_ICP_STATS
            unsigned int st_nr = 0, st_ni = 0, st_lr = 0, st_li = 0, st_q = 0, st_fb = 0, st_init = 0;
#endif
            if (__any_sync(FULL_MASK, cand != 0u)) {
                const unsigned int lt_mask = (1u << lane) - 1u;
                ws.key[lane] = key; ws.px[lane] = sx; ws.py[lane] = sy; ws.pz[lane] = sz;
                const int cnt = __popc(cand);
                const int incl = warp_incl_scan(cnt, lane);
                const int total = __shfl_sync(FULL_MASK, incl, 31);
                bool fallback = total > ICP_QCAP;
#ifdef CUBOID_ICP_STATS
                st_q = 1; st_init = (unsigned int)total;
#endif
                if (!fallback) {
                    int off = incl - cnt;
                    for (unsigned int c2 = cand; c2; c2 &= c2 - 1u) ws.nodes[off++] = ((unsigned int)lane << 16) | (unsigned int)sl[__ffs(c2) - 1];
                    __syncwarp();
                    int n = total, nl = 0;
                    while (true) {
                        if (nl >= 32 || (n == 0 && nl > 0)) {
                            // leaf round: up to 32 waiting leaves, one per lane, whoever owns them
                            const int take = min(nl, 32);
                            const bool has = lane < take;
#ifdef CUBOID_ICP_STATS
                            ++st_lr; st_li += take;
#endif
                            const unsigned int item = ws.leaves[nl - take + (has ? lane : 0)];
                            nl -= take;
                            const int owner = (int)(item >> 16), leaf = (int)(item & 0xffffu);
                            const unsigned long long k2 = icp_leaf_key(tp, s_orig, leaf, ws.px[owner], ws.py[owner], ws.pz[owner]);
                            if (has) {
                                ++nleaf_eval;
                                if (k2 < ws.key[owner]) atomicMin(&ws.key[owner], k2);
                            }
                            __syncwarp();
                            continue;
                        }
                        if (n == 0) break;
                        // node round: up to 32 waiting subtrees, one box test per lane against the owner's current minimum
                        const int take = min(n, 32);
                        const bool has = lane < take;
#ifdef CUBOID_ICP_STATS
                        ++st_nr; st_ni += take;
#endif
                        const unsigned int item = ws.nodes[n - take + (has ? lane : 0)];
                        n -= take;
                        __syncwarp();   // every lane holds its item before the pushes below reuse the popped entries
                        const int owner = (int)(item >> 16), node = (int)(item & 0xffffu);
                        const uint4 nd = s_nodes[node];
                        const float bd = __uint_as_float((unsigned int)(ws.key[owner] >> 32));
                        const bool alive = has && !(node_lb(ws.px[owner], ws.py[owner], ws.pz[owner], nd) > bd);
                        const int link = (int)nd.w;
                        const bool isleaf = alive && link < 0, inner = alive && link >= 0;
                        const unsigned int lm = __ballot_sync(FULL_MASK, isleaf), im = __ballot_sync(FULL_MASK, inner);
                        if (isleaf) ws.leaves[nl + __popc(lm & lt_mask)] = ((unsigned int)owner << 16) | (unsigned int)(~link);
                        nl += __popc(lm);
                        const int ni = __popc(im);
                        if (n + 2 * ni > ICP_QCAP) { fallback = true; break; }   // warp-uniform
                        if (inner) {
                            // depth-first layout: left child = node + 1, right child = the first node after the left subtree
                            const int lw = (int)s_nodes[node + 1].w;
                            const int right = lw >= 0 ? lw : node + 2;
                            const int r = n + 2 * __popc(im & lt_mask);
                            ws.nodes[r] = ((unsigned int)owner << 16) | (unsigned int)right;
                            ws.nodes[r + 1] = ((unsigned int)owner << 16) | (unsigned int)(node + 1);
                        }
                        n += 2 * ni;
                        __syncwarp();
                    }
                }
                __syncwarp();
                key = ws.key[lane];
                __syncwarp();   // the scratch is rewritten by the next task
#ifdef CUBOID_ICP_STATS
                st_fb = fallback ? 1 : 0;
#endif
                if (fallback) {
#ifdef CUBOID_ICP_STATS
                    if (lane == 0 && a.stats) { atomicAdd(&a.stats[0], 1ull); atomicAdd(&a.stats[1], 1ull); atomicAdd(&a.stats[2], 1ull); atomicAdd(&a.stats[3], st_nr);
                        atomicAdd(&a.stats[4], st_ni); atomicAdd(&a.stats[5], st_lr); atomicAdd(&a.stats[6], st_li); atomicAdd(&a.stats[7], st_init); }
#endif
                    // queries still far from the template: finish with the per-lane walk over the whole tree, seeded with the best so far
                    IcpBest b;
                    b.d = __uint_as_float((unsigned int)(key >> 32));
                    b.pos = (int)((unsigned int)key & 0xffffu);
                    b.orig = (int)(((unsigned int)key >> 16) & 0xffffu);
                    icp_walk<RESIDENT>(a, tp, s_nodes, s_sib, 0u, 0, nnodes, true, sx, sy, sz, b, nleaf_eval);
                    if (valid) { corr[i] = b.pos; cd[i] = b.d; }
                    evaluated += (unsigned long long)nleaf_eval * ICP_LEAF;
                    continue;
                }
            }
#ifdef CUBOID_ICP_STATS
            if (lane == 0 && a.stats) { atomicAdd(&a.stats[0], 1ull); atomicAdd(&a.stats[1], st_q); atomicAdd(&a.stats[3], st_nr);
                atomicAdd(&a.stats[4], st_ni); atomicAdd(&a.stats[5], st_lr); atomicAdd(&a.stats[6], st_li); atomicAdd(&a.stats[7], st_init);
                atomicAdd(&a.stats[8 + min(st_nr, 21u)], 1ull); }
            (void)st_fb;
#endif
            evaluated += (unsigned long long)nleaf_eval * ICP_LEAF;
            if (valid) { corr[i] = (int)((unsigned int)key & 0xffffu); cd[i] = __uint_as_float((unsigned int)(key >> 32)); }
            continue;
        }
        IcpBest b;
        b.pos = pos0;
        b.orig = -1;
        {
            const float3 t = tmpl_point(tp, b.pos);
            b.d = dist2(sx, sy, sz, t.x, t.y, t.z);
        }
        int node = 0, end = nnodes;           // [node, end): the part of the depth-first node array still to walk
        unsigned int cand = 0u;               // sibling subtrees whose box survived the lockstep test, not walked yet
        const unsigned short* sl = s_sib;
        if (outward) {
            const int L = b.pos / ICP_LEAF;
            sl = s_sib + L * a.sib_max;
            ++nleaf_eval;
            icp_scan_leaf(a, tp, L, sx, sy, sz, b);
            for (int e = 0; e < a.sib_max; ++e) {
                const unsigned int sn = sl[e];
                if (sn != 0xffffu && !(node_lb(sx, sy, sz, s_nodes[sn]) > b.d)) cand |= 1u << e;
            }
            node = 0; end = 0;                // nothing to walk until a surviving sibling is opened
        }
        icp_walk<RESIDENT>(a, tp, s_nodes, sl, cand, node, end, cull, sx, sy, sz, b, nleaf_eval);
        evaluated += (unsigned long long)nleaf_eval * ICP_LEAF;
        if (valid) { corr[i] = b.pos; cd[i] = b.d; }
    }
}

// Table pass of the nearest-neighbour search (nn_table.cuh): every query whose grid cell has a valid record scans the record's
// <= 15 candidates (sorted by original index, strict '<': the canonical tie rule) and is done; the others are appended to the
// problem's miss list, which the BVH search (icp_nn_pass) then visits instead of `order`.
template <bool LOCAL>
__device__ __forceinline__ void icp_nn_table_pass(const IcpArgs& a, IcpShared& sh, const float* tp, float4* cur, int S, const int* order,
                                                  int* corr, float* cd, int* miss, unsigned long long& evaluated, bool pend, int wid, int nw) {
    const int lane = threadIdx.x & 31;
    const int ntask = (S + 31) / 32;
    const NnTableView& T = a.tab;
    const unsigned int lt_mask = (1u << lane) - 1u;
    // Tasks (32 consecutive points) are dealt to the sub-worker's warps round-robin: a table lookup costs every task about the same,
    // so nothing is gained by claiming them from a counter (a shared-memory atomic round trip in front of every task), and with the
    // next task known its points are requested while this task's record is on its way. (Also requesting the next task's RECORD one
    // task ahead was measured: 4.71 against 4.36 ms - sixteen more live registers in a loop that already spills at 64.)
    float4 p_next = make_float4(0.f, 0.f, 0.f, 0.f);
    if (wid < ntask) p_next = icp_ld<LOCAL>(cur + min(wid * 32 + lane, S - 1));
    for (int task = wid; task < ntask; task += nw) {
        const int k = task * 32 + lane;
        const bool valid = k < S;
        // points in their own order (no visiting-order indirection: coalesced reads; a table lookup gains nothing from spatially
        // coherent lanes). The BVH pass over the misses keeps the Morton order when everything missed (early iterations).
        const int i = valid ? k : S - 1;
        float4 p = p_next;
        if (task + nw < ntask) p_next = icp_ld<LOCAL>(cur + min((task + nw) * 32 + lane, S - 1));
        if (pend && valid) {   // transformCloud(input_transformed, transformation_) of the previous iteration, folded into this read
            p = xform(sh.Tm, p);
            cur[i] = p;
        }
        const float sx = p.x, sy = p.y, sz = p.z;
        const float fx = (sx - T.org[0]) * T.inv_h, fy = (sy - T.org[1]) * T.inv_h, fz = (sz - T.org[2]) * T.inv_h;
        const bool inside = valid && fx >= 0.f && fx < (float)T.nx && fy >= 0.f && fy < (float)T.ny && fz >= 0.f && fz < (float)T.nz;
        unsigned int R[8] = {0xffffu, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
        size_t cell = 0;
        if (inside) {
            cell = ((size_t)(int)fz * T.ny + (size_t)(int)fy) * T.nx + (size_t)(int)fx;   // fx >= 0: truncation = floor
            const uint4 r0 = __ldg(T.rec + 4 * cell), r1 = __ldg(T.rec + 4 * cell + 1);
            R[0] = r0.x; R[1] = r0.y; R[2] = r0.z; R[3] = r0.w; R[4] = r1.x; R[5] = r1.y; R[6] = r1.z; R[7] = r1.w;
        }
        const unsigned int nrec = R[0] & 0xffffu;
        const bool hit = nrec != 0xffffu;
        const int n = hit ? (int)nrec : 0;
        float bd = __uint_as_float(0x7f800000u);
        int bpos = 0;
#pragma unroll
        for (int c = 0; c < 15; ++c) {
            if (c > 0 && (c & 3) == 0 && !__any_sync(FULL_MASK, c < n)) break;
            const unsigned int word = R[(c + 1) >> 1];
            const int pos = (int)(((c + 1) & 1) ? (word >> 16) : (word & 0xffffu));
            if (c < n) {
                const float3 t = tmpl_point(tp, pos);
                const float d = dist2(sx, sy, sz, t.x, t.y, t.z);
                if (d < bd) { bd = d; bpos = pos; }
            }
        }
        if (__any_sync(FULL_MASK, n > 15)) {   // cells further from the template: the second half of the record (candidates 15..30)
            if (n > 15) {
                const uint4 r2 = __ldg(T.rec + 4 * cell + 2), r3 = __ldg(T.rec + 4 * cell + 3);
                R[0] = r2.x; R[1] = r2.y; R[2] = r2.z; R[3] = r2.w; R[4] = r3.x; R[5] = r3.y; R[6] = r3.z; R[7] = r3.w;
            }
#pragma unroll
            for (int c = 15; c < NNT_K; ++c) {
                if (c > 15 && ((c - 15) & 3) == 0 && !__any_sync(FULL_MASK, c < n)) break;
                const unsigned int word = R[(c - 15) >> 1];
                const int pos = (int)(((c - 15) & 1) ? (word >> 16) : (word & 0xffffu));
                if (c < n) {
                    const float3 t = tmpl_point(tp, pos);
                    const float d = dist2(sx, sy, sz, t.x, t.y, t.z);
                    if (d < bd) { bd = d; bpos = pos; }
                }
            }
        }
        evaluated += (unsigned long long)n;
#ifdef CUBOID_ICP_STATS
        {
            const unsigned int hm = __ballot_sync(FULL_MASK, hit), vm = __ballot_sync(FULL_MASK, valid);
            if (lane == 0 && a.stats) { atomicAdd(&a.stats[30], (unsigned long long)__popc(hm)); atomicAdd(&a.stats[31], (unsigned long long)__popc(vm)); }
        }
#endif
        if (hit) { corr[i] = bpos; cd[i] = bd; }
        const unsigned int mm = __ballot_sync(FULL_MASK, valid && !hit);
        if (mm) {
            int base = 0;
            if (lane == 0) base = atomicAdd(&sh.nmiss, __popc(mm));
            base = __shfl_sync(FULL_MASK, base, 0);
            if (valid && !hit) miss[base + __popc(mm & lt_mask)] = i;
        }
    }
}

// Morton order of the source points (visiting order of the nearest-neighbour pass only; results do not depend
// on it). Sorted in the dynamic shared memory window BEFORE the template is staged there. S > cap: identity.
__device__ __forceinline__ unsigned int morton_spread10(unsigned int v) {
    v = (v | (v << 16)) & 0x030000FFu;
    v = (v | (v << 8)) & 0x0300F00Fu;
    v = (v | (v << 4)) & 0x030C30C3u;
    v = (v | (v << 2)) & 0x09249249u;
    return v;
}
__device__ void icp_morton_order(const float4* src, const int* idx, int S, int* order, unsigned long long* s_keys, int cap, float* s_mm /*[6*8]*/) {
    int n2 = 1;
    while (n2 < S) n2 <<= 1;
    if (S < 64 || n2 > cap) {
        for (int i = threadIdx.x; i < S; i += ICP_THREADS) order[i] = i;
        return;
    }
    float mn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f}, mx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
    for (int i = threadIdx.x; i < S; i += ICP_THREADS) {
        const float4 p = src[idx[i]];
        mn[0] = fminf(mn[0], p.x); mx[0] = fmaxf(mx[0], p.x);
        mn[1] = fminf(mn[1], p.y); mx[1] = fmaxf(mx[1], p.y);
        mn[2] = fminf(mn[2], p.z); mx[2] = fmaxf(mx[2], p.z);
    }
#pragma unroll
    for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            mn[c] = fminf(mn[c], __shfl_xor_sync(FULL_MASK, mn[c], o));
            mx[c] = fmaxf(mx[c], __shfl_xor_sync(FULL_MASK, mx[c], o));
        }
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    if (lane == 0)
        for (int c = 0; c < 3; ++c) { s_mm[wid * 6 + c] = mn[c]; s_mm[wid * 6 + 3 + c] = mx[c]; }
    __syncthreads();
    float ext = 0.f;
    for (int c = 0; c < 3; ++c) {
        float lo = s_mm[c], hi = s_mm[3 + c];
        for (int w = 1; w < ICP_THREADS / 32; ++w) { lo = fminf(lo, s_mm[w * 6 + c]); hi = fmaxf(hi, s_mm[w * 6 + 3 + c]); }
        mn[c] = lo;
        ext = fmaxf(ext, hi - lo);
    }
    const float scale = ext > 0.f ? 1023.0f / ext : 0.f;
    for (int i = threadIdx.x; i < n2; i += ICP_THREADS) {
        unsigned long long key = ~0ull;
        if (i < S) {
            const float4 p = src[idx[i]];
            const unsigned int qx = min(1023u, (unsigned int)((p.x - mn[0]) * scale));
            const unsigned int qy = min(1023u, (unsigned int)((p.y - mn[1]) * scale));
            const unsigned int qz = min(1023u, (unsigned int)((p.z - mn[2]) * scale));
            const unsigned int code = morton_spread10(qx) | (morton_spread10(qy) << 1) | (morton_spread10(qz) << 2);
            key = ((unsigned long long)code << 32) | (unsigned int)i;
        }
        s_keys[i] = key;
    }
    __syncthreads();
    for (int k = 2; k <= n2; k <<= 1) {
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < n2; t += ICP_THREADS) {
                const int u = t ^ j;
                if (u > t) {
                    const unsigned long long x = s_keys[t], y = s_keys[u];
                    const bool up = (t & k) == 0;
                    if ((x > y) == up) { s_keys[t] = y; s_keys[u] = x; }
                }
            }
            __syncthreads();
        }
    }
    for (int i = threadIdx.x; i < S; i += ICP_THREADS) order[i] = (int)(unsigned int)s_keys[i];
    __syncthreads();
}

// ---- problem state that survives between time slices (global memory) --------------------------------------------------
struct IcpState {
    float fin[16];             // final_transformation_
    double prev_mse;
    int it, passes, done, converged, state, pad;
    unsigned long long chash;  // sum of the per-iteration correspondence hashes so far
    unsigned long long evaluated;
};
// Work queue of one launch: a ring with one slot per problem of the launch. A problem is in exactly one place (one slot, or
// one CTA's hands), so at most n_slots entries are ever unread and a slot is never overwritten before it was read.
struct IcpQueue {
    unsigned int head, tail;   // tickets handed to poppers / pushers
    int n_done, n_total;       // problems finished / problems of this launch (set by k_icp_init)
    int error;
    int alive;                 // worker CTAs that have not left yet; a worker leaves when fewer problems than workers remain
    int pad[2];
};
struct IcpSlot { int prob; unsigned int seq; };   // seq == ticket + 1 once the slot holds the entry of that ticket

__device__ __forceinline__ void icp_queue_push(IcpQueue* q, IcpSlot* ring, int n_slots, int prob) {
    const unsigned int t = atomicAdd(&q->tail, 1u);
    IcpSlot* sl = ring + (t % (unsigned int)n_slots);
    sl->prob = prob;
    __threadfence();
    ((volatile IcpSlot*)sl)->seq = t + 1u;
}
// called by ONE thread; returns the next problem, or -1 when every problem of the launch is finished
__device__ __forceinline__ int icp_queue_pop(IcpQueue* q, IcpSlot* ring, int n_slots) {
    volatile IcpQueue* vq = q;
    const int remaining = vq->n_total - vq->n_done;
    if (remaining <= 0) return -1;
    // more workers than unfinished problems: this one is not needed any more (checked BEFORE a ticket is claimed, a claimed
    // ticket is always waited for). `remaining` only shrinks, so a stale value errs on the side of staying.
    if (vq->alive > remaining) {
        if (atomicSub(&q->alive, 1) - 1 >= remaining) return -1;
        atomicAdd(&q->alive, 1);
    }
    const unsigned int t = atomicAdd(&q->head, 1u);
    volatile IcpSlot* sl = ring + (t % (unsigned int)n_slots);
    for (unsigned int spin = 0;; ++spin) {
        if (sl->seq == t + 1u) { __threadfence(); return sl->prob; }
        if (vq->n_done >= vq->n_total) return -1;
        if (spin > (1u << 26)) { atomicExch(&q->error, 1); return -1; }   // ~10 s of sleeping: never hang the device
        __nanosleep(128);
    }
}

// k_icp_init: one CTA per (guess, cluster, frame) slot. Empty slots finish at once; the others get their Morton visiting
// order, the guess-transformed working cloud, seeds, a fresh state, and are pushed on the queue.
__global__ void __launch_bounds__(ICP_THREADS) k_icp_init(const IcpArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    __shared__ float s_part[16 * 8];
    __shared__ float s_red[16];
    __shared__ float s_guess[16];
    const int g = blockIdx.x, c = blockIdx.y, f = blockIdx.z;
    const cuboid_frame_result& R = a.res[f];
    if (blockIdx.x == 0 && blockIdx.y == 0 && blockIdx.z == 0 && threadIdx.x == 0) a.queue->alive = a.crew * a.nsub;   // poppers = sub-workers
    if (c >= min(R.n_clusters, CUBOID_MAX_CLUSTERS)) return;   // not a problem: n_total counts only real ones
    const int prob = (f * CUBOID_MAX_CLUSTERS + c) * a.n_guess + g;
    const int* offsets = a.offsets + (size_t)f * (a.KC + 1);
    const int o0 = offsets[c], S = offsets[c + 1] - o0;
    const int* idx = a.idx_sorted + (size_t)f * a.M + o0;
    const float4* src = a.remain + (size_t)f * a.P;
    const size_t pbase = ((size_t)f * a.n_guess + g) * a.M + o0;
    float4* cur = a.cur + pbase;
    int* corr = a.corr + pbase;
    int* order = a.order + pbase;
    const bool lane_thread = threadIdx.x < ICP_LANES;
    icp_morton_order(src, idx, S, order, reinterpret_cast<unsigned long long*>(smem_raw), a.init_smem / 8, s_part);
    // ---- guess: final = guess; src_t = (guess == I) ? src : guess * src ----
    if (threadIdx.x < 16) {
        float v = (threadIdx.x % 5 == 0) ? 1.f : 0.f;
        if (a.guesses && a.guess_mode == 0) v = a.guesses[(size_t)g * 16 + threadIdx.x];
        s_guess[threadIdx.x] = v;
    }
    __syncthreads();
    if (a.guesses && a.guess_mode == 1) {
        // rotation about the cluster centroid: G = T(c) R T(-c), centroid by the canonical reduction
        float v3[3] = {0.f, 0.f, 0.f};
        if (lane_thread)
            for (int i = threadIdx.x; i < S; i += ICP_LANES) {
                const float4 p = src[idx[i]];
                v3[0] = v3[0] + p.x; v3[1] = v3[1] + p.y; v3[2] = v3[2] + p.z;
            }
        canon_block_reduce<float, 3>(v3, s_part, s_red);
        if (threadIdx.x == 0) {
            const float cn = (float)S;
            const float cx = s_red[0] / cn, cy = s_red[1] / cn, cz = s_red[2] / cn;
            const float* Rm = a.guesses + (size_t)g * 9;
            const float cc[3] = {cx, cy, cz};
            for (int i = 0; i < 3; ++i) {
                for (int j = 0; j < 3; ++j) s_guess[4 * i + j] = Rm[3 * i + j];
                s_guess[4 * i + 3] = cc[i] - ((Rm[3 * i] * cx + Rm[3 * i + 1] * cy) + Rm[3 * i + 2] * cz);
            }
        }
        __syncthreads();
    }
    bool guess_identity = true;
#pragma unroll
    for (int k = 0; k < 16; ++k) guess_identity = guess_identity && (s_guess[k] == ((k % 5 == 0) ? 1.f : 0.f));
    for (int i = threadIdx.x; i < S; i += ICP_THREADS) {
        const float4 p = src[idx[i]];
        cur[i] = guess_identity ? make_float4(p.x, p.y, p.z, 1.0f) : xform(s_guess, p);
        corr[i] = 0;   // seed of the first nearest-neighbour pass: any valid template position
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        IcpState st;
        for (int k = 0; k < 16; ++k) st.fin[k] = s_guess[k];
        st.prev_mse = 1.7976931348623157e308;
        st.it = 0; st.passes = 0; st.converged = 0; st.pad = 0;
        st.done = S < 3 ? 1 : 0;   // "Not enough correspondences found" -> converged_ = false, no iteration
        st.state = S < 3 ? CUBOID_ICP_NO_CORRESPONDENCES : CUBOID_ICP_NOT_CONVERGED;
        st.chash = 0ull; st.evaluated = 0ull;
        a.pstate[prob] = st;
        __threadfence();
        atomicAdd(&a.queue->n_total, 1);
        icp_queue_push(a.queue, a.ring, a.n_slots, prob);
    }
}

// One time slice of one problem: up to a.slice_iters iterations of the ICP loop, then either the closing fitness pass
// (problem finished) or a state save (problem goes back on the queue).
template <bool RESIDENT, bool QUEUED, bool TABLE, bool LOCAL, int SUB, bool REJECT = false>
__device__ __forceinline__ bool icp_slice(const IcpArgs& a, IcpShared& sh, const float* s_tmpl, const uint4* s_nodes, const unsigned short* s_sib,
                                          const unsigned short* s_orig, IcpWarpScr* wscr, unsigned char* s_local, int prob, int tid, int sub,
                                          unsigned long long* s_hh, unsigned long long* s_ev) {
    const int g = prob % a.n_guess, c = (prob / a.n_guess) % CUBOID_MAX_CLUSTERS, f = prob / (a.n_guess * CUBOID_MAX_CLUSTERS);
    const int* offsets = a.offsets + (size_t)f * (a.KC + 1);
    const int o0 = offsets[c], S = offsets[c + 1] - o0;
    const int* idx = a.idx_sorted + (size_t)f * a.M + o0;
    const float4* src = a.remain + (size_t)f * a.P;
    const size_t pbase = ((size_t)f * a.n_guess + g) * a.M + o0;
    float4* const g_cur = a.cur + pbase;
    int* const g_corr = a.corr + pbase;
    // LOCAL: the working cloud, its correspondences and distances live in shared memory for the whole slice
    float4* cur = LOCAL ? reinterpret_cast<float4*>(s_local) : g_cur;
    int* corr = LOCAL ? reinterpret_cast<int*>(s_local + (size_t)S * 16) : g_corr;
    float* cd = LOCAL ? reinterpret_cast<float*>(s_local + (size_t)S * 20) : a.cd + pbase;
    const int* order = a.order + pbase;
    int* miss = TABLE ? a.miss + pbase : nullptr;
    IcpState& ps = a.pstate[prob];
    if (LOCAL) {
        for (int i = tid; i < S; i += SUB) { cur[i] = __ldcg(g_cur + i); corr[i] = __ldcg(g_corr + i); }
    }
    const bool trace = a.corr_trace && f == 0 && c == 0 && g == 0;
    const float* tp = RESIDENT ? s_tmpl : a.tmpl;
    constexpr int LPT = SUB >= ICP_LANES ? 1 : ICP_LANES / SUB;   // canonical lanes per thread
    const bool canon = SUB <= ICP_LANES || tid < ICP_LANES;     // SUB = 512: the upper half only helps outside the reductions

    if (tid < 16) sh.fin[tid] = __ldcg(&ps.fin[tid]);
    if (tid == 0) {
        sh.done = __ldcg(&ps.done); sh.converged = __ldcg(&ps.converged); sh.state = __ldcg(&ps.state); sh.iters = __ldcg(&ps.it);
        sh.prev_mse = __ldcg(&ps.prev_mse);
        sh.task = 0; sh.nmiss = 0;
    }
    sub_sync<SUB>(sub);
    unsigned long long chash = 0, evaluated = 0;
    const float one_over_n = 1.0f / (float)S;
    int it = sh.iters, passes = 0;
    const int it_end = it + a.slice_iters;
    if (it == 0 && !sh.done && S > 64 && a.cull) {
        // First pass of a problem: every seed is template position 0, i.e. no bound at all, and the cloud is typically far from
        // the template, so each query would walk most of the tree. Search the first warp's worth of queries properly, then hand
        // the answer of one of them to everybody as the seed: any valid template position is a valid seed, results are unchanged.
        icp_nn_pass<RESIDENT, QUEUED, LOCAL>(a, sh, s_tmpl, s_nodes, s_sib, s_orig, wscr, cur, 32, order, corr, cd, evaluated);
        sub_sync<SUB>(sub);
        if (tid == 0) sh.task = 0;
        const int seed = icp_ld<LOCAL>(corr + order[0]);
        sub_sync<SUB>(sub);
        for (int i = tid; i < S; i += SUB) corr[i] = seed;
        sub_sync<SUB>(sub);
    }
    bool pend = false;   // TABLE: this iteration's transformation_ is applied by the next table pass (or below, when the slice ends)
    while (!sh.done && it < it_end) {
        // 1. correspondences: the candidate table answers the queries close to the template, the BVH search the rest
        if (TABLE) {
            icp_nn_table_pass<LOCAL>(a, sh, tp, cur, S, order, corr, cd, miss, evaluated, pend, tid >> 5, SUB / 32);
            pend = false;
            sub_sync<SUB>(sub);
            const int nm = sh.nmiss;
            if (nm > 0) {   // (uniform) the usual late iteration has no misses and goes on behind the one barrier above
                if (tid == 0) sh.task = 0;
                sub_sync<SUB>(sub);
                if (tid == 0) sh.nmiss = 0;
                icp_nn_pass<RESIDENT, QUEUED, LOCAL>(a, sh, s_tmpl, s_nodes, s_sib, s_orig, wscr, cur, nm, nm == S ? order : miss, corr, cd, evaluated);
                sub_sync<SUB>(sub);
            }
        } else {
            icp_nn_pass<RESIDENT, QUEUED, LOCAL>(a, sh, s_tmpl, s_nodes, s_sib, s_orig, wscr, cur, S, order, corr, cd, evaluated);
            sub_sync<SUB>(sub);
        }
        ++passes;
        if (tid == 0) sh.task = 0;
        // REJECT (a finite setMaxCorrespondenceDistance): CorrespondenceEstimation drops the pairs beyond the distance and keeps the
        // order of the rest; everything below then runs over that compacted list (its canonical lanes are positions IN the list).
        int NC = S;
        const int* list = nullptr;
        if constexpr (REJECT) {
            int* wl = a.miss + pbase;                                  // the table's miss list is free in this mode
            int* s_cnt = reinterpret_cast<int*>(sh.part_f);            // [SUB / 32]
            const int lane = tid & 31, wid = tid >> 5;
            int base = 0;
            for (int c0 = 0; c0 < S; c0 += SUB) {
                const int i = c0 + tid;
                const bool acc = i < S && !((double)icp_ld<LOCAL>(cd + i) > a.max_d2);
                const unsigned int bal = __ballot_sync(FULL_MASK, acc);
                if (lane == 0) s_cnt[wid] = __popc(bal);
                sub_sync<SUB>(sub);
                int before = 0, tot = 0;
                for (int w = 0; w < SUB / 32; ++w) { const int c = s_cnt[w]; before += w < wid ? c : 0; tot += c; }
                if (acc) wl[base + before + __popc(bal & ((1u << lane) - 1u))] = i;
                base += tot;
                sub_sync<SUB>(sub);
            }
            NC = base;
            list = wl;
            if (NC < 3) {              // "Not enough correspondences found": not converged, iteration not counted
                if (tid == 0) { sh.done = 1; sh.converged = 0; sh.state = CUBOID_ICP_NO_CORRESPONDENCES; }
                sub_sync<SUB>(sub);
                break;
            }
            if (a.hashes || trace) {   // the parity taps see every source point; a dropped pair reads -1
                for (int i = tid; i < S; i += SUB) {
                    const int j = ((double)icp_ld<LOCAL>(cd + i) > a.max_d2) ? -1 : a.tmpl_orig[icp_ld<LOCAL>(corr + i)];
                    chash += splitmix64((((unsigned long long)it * (unsigned long long)S + (unsigned long long)i) << 32) | (unsigned int)j);
                    if (trace && it < a.cap_trace) a.corr_trace[(size_t)it * S + i] = j;
                }
            }
        }
        const float oon = REJECT ? 1.0f / (float)NC : one_over_n;
        // 2. means + MSE: the first 256 threads are the 256 canonical lanes
        for (int set = 0; set < LPT; ++set) {
            float q6[6] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            double qd[1] = {0.0};
            for (int ci = canon ? tid + set * SUB : NC; ci < NC; ci += ICP_LANES) {
                const int i = REJECT ? __ldcg(list + ci) : ci;
                const float4 p = icp_ld<LOCAL>(cur + i);
                const int pos = icp_ld<LOCAL>(corr + i);
                const float3 t = tmpl_point(tp, pos);
                q6[0] = q6[0] + p.x; q6[1] = q6[1] + p.y; q6[2] = q6[2] + p.z;
                q6[3] = q6[3] + t.x; q6[4] = q6[4] + t.y; q6[5] = q6[5] + t.z;
                qd[0] = qd[0] + (double)icp_ld<LOCAL>(cd + i);
                if (!REJECT && (a.hashes || trace)) {   // parity taps: the correspondence hash and the per-iteration trace
                    const int j = a.tmpl_orig[pos];  // original template index
                    chash += splitmix64((((unsigned long long)it * (unsigned long long)S + (unsigned long long)i) << 32) | (unsigned int)j);
                    if (trace && it < a.cap_trace) a.corr_trace[(size_t)it * S + i] = j;
                }
            }
            canon_sub_partial<float, 6, SUB>(q6, sh.part_f, tid, set);
            canon_sub_partial<double, 1, SUB>(qd, sh.part_d, tid, set);
        }
        canon_sub_finish_fd<6, SUB>(sh.part_f, sh.red_f, sh.part_d, sh.red_d, tid, sub);
        float sm[3], dm[3];
#pragma unroll
        for (int k = 0; k < 3; ++k) { sm[k] = sh.red_f[k] * oon; dm[k] = sh.red_f[3 + k] * oon; }
        const double mse_sum = sh.red_d[0];
        // (no barrier here: red_f / red_d are next written behind the first barrier of the covariance step's finish, part_f was read
        // in front of the second barrier of the finish above)
        // 3. sigma = one_over_n * dst_demean * src_demean^T
        for (int set = 0; set < LPT; ++set) {
            float q9[9];
#pragma unroll
            for (int k = 0; k < 9; ++k) q9[k] = 0.f;
            for (int ci = canon ? tid + set * SUB : NC; ci < NC; ci += ICP_LANES) {
                const int i = REJECT ? __ldcg(list + ci) : ci;
                const float4 p = icp_ld<LOCAL>(cur + i);
                const float3 t = tmpl_point(tp, icp_ld<LOCAL>(corr + i));
                const float sd[3] = {p.x - sm[0], p.y - sm[1], p.z - sm[2]};
                const float dd[3] = {t.x - dm[0], t.y - dm[1], t.z - dm[2]};
#pragma unroll
                for (int r = 0; r < 3; ++r)
#pragma unroll
                    for (int cc = 0; cc < 3; ++cc) q9[3 * r + cc] = q9[3 * r + cc] + dd[r] * sd[cc];
            }
            canon_sub_partial<float, 9, SUB>(q9, sh.part_f, tid, set);
        }
        canon_sub_finish<float, 9, SUB>(sh.part_f, sh.red_f, tid, sub);
        // 4. thread 0: SVD, R, t, final, convergence
        if (tid == 0) {
            M3f sigma, U, V;
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int cc = 0; cc < 3; ++cc) sigma.a[r][cc] = oon * sh.red_f[3 * r + cc];
            jacobi_svd3(sigma, U, V);
            float Sg[3] = {1.f, 1.f, 1.f};
            if (det3(U) * det3(V) < 0.f) Sg[2] = -1.f;
            float Tm[16];
#pragma unroll
            for (int i = 0; i < 3; ++i) {
#pragma unroll
                for (int j = 0; j < 3; ++j)
                    Tm[4 * i + j] = ((U.a[i][0] * Sg[0]) * V.a[j][0] + (U.a[i][1] * Sg[1]) * V.a[j][1]) + (U.a[i][2] * Sg[2]) * V.a[j][2];
            }
#pragma unroll
            for (int i = 0; i < 3; ++i) Tm[4 * i + 3] = dm[i] - ((Tm[4 * i] * sm[0] + Tm[4 * i + 1] * sm[1]) + Tm[4 * i + 2] * sm[2]);
            Tm[12] = 0.f; Tm[13] = 0.f; Tm[14] = 0.f; Tm[15] = 1.f;
            float nf[16];
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j)
                    nf[4 * i + j] = ((Tm[4 * i] * sh.fin[j] + Tm[4 * i + 1] * sh.fin[4 + j]) + Tm[4 * i + 2] * sh.fin[8 + j]) + Tm[4 * i + 3] * sh.fin[12 + j];
#pragma unroll
            for (int k = 0; k < 16; ++k) { sh.fin[k] = nf[k]; sh.Tm[k] = Tm[k]; }
            if (trace && it < a.cap_trace) for (int k = 0; k < 16; ++k) a.T_trace[(size_t)it * 16 + k] = Tm[k];
            const int itn = it + 1;
            sh.iters = itn;
            // DefaultConvergenceCriteria::hasConverged
            int done = 0;
            if (itn >= a.max_iter) { done = 1; sh.converged = 1; sh.state = CUBOID_ICP_ITERATIONS; }
            if (!done) {
                const double cos_angle = 0.5 * (double)(Tm[0] + Tm[5] + Tm[10] - 1.0f);
                const double tsq = (double)(Tm[3] * Tm[3] + Tm[7] * Tm[7] + Tm[11] * Tm[11]);
                if (cos_angle >= a.rot_thr && tsq <= a.trans_thr) { done = 1; sh.converged = 1; sh.state = CUBOID_ICP_TRANSFORM; }
            }
            if (!done) {
                const double mse = mse_sum / (double)NC;
                if (fabs(mse - sh.prev_mse) < a.abs_thr) { done = 1; sh.converged = 1; sh.state = CUBOID_ICP_ABS_MSE; }
                else if (fabs(mse - sh.prev_mse) / sh.prev_mse < a.rel_mse) { done = 1; sh.converged = 1; sh.state = CUBOID_ICP_REL_MSE; }
                else sh.prev_mse = mse;
            }
            sh.done = done;
        }
        sub_sync<SUB>(sub);
        // 5. transformCloud(input_transformed, input_transformed, transformation_): incremental, in place. With the candidate table
        //    the next pass over the cloud is the table pass of the next iteration, which applies it as it reads (one pass and one
        //    barrier less); when ICP has converged the cloud is rebuilt from the source with the final transformation anyway.
        ++it;
        if (TABLE) {
            pend = true;
        } else {
            for (int i = tid; i < S; i += SUB) cur[i] = xform(sh.Tm, icp_ld<LOCAL>(cur + i));
            sub_sync<SUB>(sub);
        }
    }
    if (TABLE && pend && !sh.done) {   // the slice ends between two iterations: the saved cloud must be the transformed one
        for (int i = tid; i < S; i += SUB) cur[i] = xform(sh.Tm, icp_ld<LOCAL>(cur + i));
        sub_sync<SUB>(sub);
    }
    const bool finished = sh.done != 0;
    double fitness = 1.7976931348623157e308;
    if (finished) {
        // ---- output = transformCloud(src, final); getFitnessScore(): one more nearest-neighbour pass ----
        for (int i = tid; i < S; i += SUB) cur[i] = xform(sh.fin, src[idx[i]]);
        sub_sync<SUB>(sub);
        if (S > 0) {
            if (TABLE) {
                icp_nn_table_pass<LOCAL>(a, sh, tp, cur, S, order, corr, cd, miss, evaluated, false, tid >> 5, SUB / 32);
                sub_sync<SUB>(sub);
                const int nm = sh.nmiss;
                if (tid == 0) sh.task = 0;
                sub_sync<SUB>(sub);
                if (tid == 0) sh.nmiss = 0;
                if (nm > 0) icp_nn_pass<RESIDENT, QUEUED, LOCAL>(a, sh, s_tmpl, s_nodes, s_sib, s_orig, wscr, cur, nm, nm == S ? order : miss, corr, cd, evaluated);
            } else {
                icp_nn_pass<RESIDENT, QUEUED, LOCAL>(a, sh, s_tmpl, s_nodes, s_sib, s_orig, wscr, cur, S, order, corr, cd, evaluated);
            }
            ++passes;
            sub_sync<SUB>(sub);
            for (int set = 0; set < LPT; ++set) {
                double qd[1] = {0.0};
                for (int i = canon ? tid + set * SUB : S; i < S; i += ICP_LANES) qd[0] = qd[0] + (double)icp_ld<LOCAL>(cd + i);
                canon_sub_partial<double, 1, SUB>(qd, sh.part_d, tid, set);
            }
            canon_sub_finish<double, 1, SUB>(sh.part_d, sh.red_d, tid, sub);
            fitness = sh.red_d[0] / (double)S;
        }
    }
    if (LOCAL) {   // hand the working cloud back: the next slice (or k_icp_aligned) reads it from global memory
        for (int i = tid; i < S; i += SUB) { g_cur[i] = cur[i]; g_corr[i] = corr[i]; }
    }
    // reduce this slice's share of the correspondence hash and of the work counters
    chash = warp_sum_u64(chash);
    evaluated = warp_sum_u64(evaluated);   // pairs evaluated by the 32 lanes
    if ((tid & 31) == 0) { s_hh[tid >> 5] = chash; s_ev[tid >> 5] = evaluated; }
    sub_sync<SUB>(sub);
    if (tid == 0) {
        unsigned long long t = __ldcg(&ps.chash), ev = __ldcg(&ps.evaluated);
        for (int k = 0; k < SUB / 32; ++k) { t += s_hh[k]; ev += s_ev[k]; }
        const int all_passes = __ldcg(&ps.passes) + passes;
        if (finished) {
            IcpOut& out = a.out[prob];
            for (int k = 0; k < 16; ++k) out.T[k] = sh.fin[k];
            out.fitness = fitness;
            out.converged = sh.converged;
            out.iters = sh.iters;
            out.state = sh.state;
            out.corr_hash = t;
            if (a.work) {
                atomicAdd(&a.work[0], ev);
                atomicAdd(&a.work[1], (unsigned long long)all_passes * (unsigned long long)S * (unsigned long long)a.T);
            }
        } else {
            for (int k = 0; k < 16; ++k) ps.fin[k] = sh.fin[k];
            ps.prev_mse = sh.prev_mse; ps.it = sh.iters; ps.passes = all_passes; ps.done = 0; ps.converged = sh.converged; ps.state = sh.state;
            ps.chash = t; ps.evaluated = ev;
        }
    }
    sub_sync<SUB>(sub);
    return finished;
}

// Persistent worker: stages the BVH (and, if they fit, the template, the sibling chains and the original-index table) in shared
// memory ONCE, then serves time slices of whatever problem is next on the queue until every problem of the launch is finished.
// Slicing bounds the tail: without it the kernel ends when the problem with the most iterations (40 .. 150 here) ends, with
// most SMs idle by then. MODE 0: nodes only in shared memory (template through L1/L2); 1: resident template, per-lane walk;
// 2: resident template + queued outward search (icp_nn_pass); 3: 2 + the nearest-neighbour candidate table in front of it;
// 4: 1 + a finite maximum correspondence distance (pairs beyond it dropped: the compacted-list path of icp_slice).
template <int SUB, int MODE, int NT = ICP_NT>
__global__ void __launch_bounds__(NT, 1) k_icp(const IcpArgs a) {
    constexpr int NSUB = NT / SUB;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ IcpShared shs[NSUB];
    __shared__ int s_prob[NSUB];
    __shared__ unsigned long long s_hh[NSUB][SUB / 32], s_ev[NSUB][SUB / 32];
    // dynamic shared memory: [BVH nodes (nnodes x 16 B)] [template SoA leaves (Tpad*3 floats)] [sibling chains] [orig16 (Tpad u16)]
    // [per-warp scratch x 32]   (everything after the nodes: resident modes; the last two: MODE 2)
    uint4* s_nodes = reinterpret_cast<uint4*>(smem_raw);
    float* s_tmpl = reinterpret_cast<float*>(s_nodes + a.nnodes);
    const unsigned int tb = MODE >= 1 ? (unsigned int)a.Tpad * 12u : 0u;
    const unsigned int bb = (unsigned int)a.nnodes * 16u;
    const unsigned int sb = (MODE >= 1 && a.sib_on) ? (unsigned int)a.sib_bytes : 0u;
    const unsigned int ob = (MODE == 2 || MODE == 3) ? (unsigned int)a.Tpad * 2u : 0u;
    const unsigned short* s_sib = reinterpret_cast<const unsigned short*>(reinterpret_cast<unsigned char*>(s_tmpl) + tb);
    const unsigned short* s_orig = reinterpret_cast<const unsigned short*>(reinterpret_cast<unsigned char*>(s_tmpl) + tb + sb);
    IcpWarpScr* s_wscr = reinterpret_cast<IcpWarpScr*>(reinterpret_cast<unsigned char*>(s_tmpl) + tb + sb + ob);
    if (threadIdx.x == 0) {
        mbar_init(&shs[0].bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        mbar_expect_tx(&shs[0].bar, tb + bb + sb + ob);
        unsigned char* t8 = reinterpret_cast<unsigned char*>(s_tmpl);
        for (unsigned int off = 0; off < ob; off += 32768u)
            tma_bulk_g2s(t8 + tb + sb + off, reinterpret_cast<const unsigned char*>(a.orig16) + off, min(32768u, ob - off), &shs[0].bar);
        for (unsigned int off = 0; off < sb; off += 32768u)
            tma_bulk_g2s(t8 + tb + off, reinterpret_cast<const unsigned char*>(a.sib) + off, min(32768u, sb - off), &shs[0].bar);
        for (unsigned int off = 0; off < tb; off += 32768u)
            tma_bulk_g2s(t8 + off, reinterpret_cast<const unsigned char*>(a.tmpl) + off, min(32768u, tb - off), &shs[0].bar);
        for (unsigned int off = 0; off < bb; off += 32768u)
            tma_bulk_g2s(reinterpret_cast<unsigned char*>(s_nodes) + off, reinterpret_cast<const unsigned char*>(a.nodes) + off, min(32768u, bb - off), &shs[0].bar);
    }
    mbar_wait(&shs[0].bar, 0);
    __syncthreads();
    // from here on the sub-workers go their own ways
    const int sub = threadIdx.x / SUB, tid = threadIdx.x % SUB;
    IcpShared& sh = shs[sub];
    IcpWarpScr* wscr = s_wscr + (threadIdx.x >> 5);
    while (true) {
        if (tid == 0) s_prob[sub] = icp_queue_pop(a.queue, a.ring, a.n_slots);
        sub_sync<SUB>(sub);
        const int prob = s_prob[sub];
        if (prob < 0) break;
        bool finished;
        if (MODE == 3 && SUB == NT) {   // one sub-worker per CTA: room for the problem's working set next to the template
            const int pc = (prob / a.n_guess) % CUBOID_MAX_CLUSTERS, pf = prob / (a.n_guess * CUBOID_MAX_CLUSTERS);
            const int* po = a.offsets + (size_t)pf * (a.KC + 1);
            const int pS = po[pc + 1] - po[pc];
            unsigned char* s_local = reinterpret_cast<unsigned char*>(s_wscr + NT / 32);
            if (pS <= a.local_cap) finished = icp_slice<true, true, true, true, SUB>(a, sh, s_tmpl, s_nodes, s_sib, s_orig, wscr, s_local, prob, tid, sub, s_hh[sub], s_ev[sub]);
            else finished = icp_slice<true, true, true, false, SUB>(a, sh, s_tmpl, s_nodes, s_sib, s_orig, wscr, nullptr, prob, tid, sub, s_hh[sub], s_ev[sub]);
        } else {
            finished = icp_slice<(MODE >= 1), (MODE == 2 || MODE == 3), (MODE == 3), false, SUB, (MODE == 4)>(a, sh, s_tmpl, s_nodes, s_sib, s_orig, wscr, nullptr, prob, tid, sub, s_hh[sub], s_ev[sub]);
        }
        if (tid == 0) {
            __threadfence();   // state / outputs before the hand-over
            if (finished) atomicAdd(&a.queue->n_done, 1);
            else icp_queue_push(a.queue, a.ring, a.n_slots, prob);
        }
        sub_sync<SUB>(sub);
    }
}

// best guess per (frame, cluster): lowest fitness, ties -> lowest guess id; fills cuboid_cluster_result
__global__ void k_icp_select(const IcpOut* out, cuboid_frame_result* res, int n_frames, int n_guess, double gate, int guess_offset) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = t / CUBOID_MAX_CLUSTERS, c = t % CUBOID_MAX_CLUSTERS;
    if (f >= n_frames) return;
    cuboid_frame_result& R = res[f];
    if (c >= min(R.n_clusters, CUBOID_MAX_CLUSTERS)) return;
    const IcpOut* o = out + ((size_t)f * CUBOID_MAX_CLUSTERS + c) * n_guess;
    int bg = 0;
    for (int g = 1; g < n_guess; ++g) if (o[g].fitness < o[bg].fitness) bg = g;
    cuboid_cluster_result& C = R.cluster[c];
    C.converged = o[bg].converged;
    C.iterations = o[bg].iters;
    C.best_guess = bg + guess_offset;   // global hypothesis id when the hypotheses are split over GPUs
    C.state = o[bg].state;
    C.fitness = o[bg].fitness;
    C.accepted = (o[bg].converged && o[bg].fitness < gate) ? 1 : 0;
    for (int k = 0; k < 16; ++k) C.T[k] = o[bg].T[k];
    C.corr_hash = o[bg].corr_hash;
}
// single-problem API (cuboid_icp): hand back the aligned cloud of the winning guess of (frame 0, cluster 0)
__global__ void k_icp_aligned(const cuboid_frame_result* res, const float4* cur, int M, const int* offsets, float4* aligned_out) {
    if (res[0].n_clusters < 1) return;
    const int bg = res[0].cluster[0].best_guess;
    const int o0 = offsets[0], S = offsets[1] - o0;
    const float4* src = cur + (size_t)bg * M + o0;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < S; i += gridDim.x * blockDim.x) aligned_out[i] = src[i];
}

}  // namespace cuboid
