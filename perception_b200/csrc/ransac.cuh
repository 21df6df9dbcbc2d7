// ransac.cuh — stage 2: pcl::SACSegmentation<PointXYZ> SACMODEL_PLANE / SAC_RANSAC / optimize, followed by
// pcl::ExtractIndices (gps.cpp:76-101, opd.cpp:301-326) and the optional second
// PassThrough z (opd.cpp:331-336). One CTA per frame, the whole adaptive loop runs on the device:
//
//   draw      thread 0 replays SampleConsensusModel::drawIndexSample on the persistent shuffled index array
//             with the precomputed boost::mt19937(seed) >> 1 stream (same stream for every frame)
//   score     one warp per hypothesis: |dot4(coef,(x,y,z,1))| < thr counted with __ballot_sync/__popc over
//             point tiles staged in shared memory by a TMA bulk copy (cp.async.bulk + mbarrier), double buffered
//   accept    thread 0 replays the sequential acceptance / adaptive-k rule over the scored batch, so the
//             stopping iteration is exactly the serial algorithm's
//   refine    computeMeanAndCovarianceMatrix with PCL's sequential float accumulation order (one lane per
//             accumulator), pcl::eigen33 closed form, then the inlier re-selection and the extraction
//
// Roofline: scoring is FP32-pipe bound (8*H*V ops), refine/reselect/extract are HBM/L2 bound
// (16*I + 16*V + 4*I' and 16*V + 16*M bytes per frame).
#pragma once
#include "common.cuh"

namespace cuboid {

struct SacArgs {
    const float4* vox;          // [F][P]
    int* shuffled;              // [F][P] scratch: the persistent shuffled_indices_
    const int* rng;             // mt19937(seed)() >> 1 stream
    int rng_len;
    const int* triplets;        // optional explicit samples (frame 0 only), else NULL
    int n_triplets;
    int* inl_pre;               // [F][P]
    int* inl;                   // [F][P]
    float4* remain;             // [F][P]
    cuboid_frame_result* res;
    FrameScratch* scr;
    int P;
    float thr_f;                // smallest float >= (double)threshold: (double)|d| < thr  <=>  |d| < thr_f
    int max_iter;
    double log_prob;            // log(1 - prob), computed on the host like PCL does
    int refine, negative;
    int use_z2; float z2_lo, z2_hi;
    int cap_remain;             // frames whose remainder exceeds this are flagged and truncated
    // SACMODEL_PERPENDICULAR_PLANE (1) / SACMODEL_PARALLEL_PLANE (2) of surface_normal_estimation.cpp:118-123; 0 = SACMODEL_PLANE
    int model_type;
    float axis[3];              // seg.setAxis
    double cos_eps, sin_eps;    // cos / |sin| of seg.setEpsAngle, evaluated on the host in double; eps <= 0 switches the test off
    int use_bbox;               // bbox_filter.cpp as a fused predicate of the extraction (after PassThrough z2)
    double bbP[12];             // CameraInfo P, row-major 3x4
    int bb[4];                  // x1, y1, x2, y2
};

constexpr int SAC_THREADS = 256;
constexpr int SAC_TILE = 1024;      // points per shared-memory tile (16 KB), two buffers
constexpr int SAC_PROD = 512;       // inliers per refine chunk; the product rows live in the (then idle) tile buffers
constexpr int SAC_HB = 32;          // hypotheses scored per round (first round: 8)

struct SacHyp { float c[4]; int valid; int kind; int ok; };   // kind: 0 normal, 1 sampler gave up (selection.empty()); ok: isModelValid()

// ---- TMA bulk copy helpers (SASS: UBLKCP / SYNCS) -------------------------------------------------
__device__ __forceinline__ unsigned int smem_u32(const void* p) { return (unsigned int)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned int bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned int phase) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(phase) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst_smem, const void* src_gmem, unsigned int bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst_smem)),
                 "l"(src_gmem), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// bbox_filter.cpp:30-51 (within_bbox): u, v, w accumulate in double (the matrix is vector<double>) and are stored to
// float; the divide and the strict comparisons against the int rectangle run in float.
__device__ __forceinline__ bool bbox_within(const double* P, const int* bb, float xf, float yf, float zf) {
    const double x = (double)xf, y = (double)yf, z = (double)zf;
    const float u = (float)((((P[0] * x) + (P[1] * y)) + (P[2] * z)) + P[3]);
    const float v = (float)((((P[4] * x) + (P[5] * y)) + (P[6] * z)) + P[7]);
    const float w = (float)((((P[8] * x) + (P[9] * y)) + (P[10] * z)) + P[11]);
    const float un = u / w, vn = v / w;
    return ((float)bb[0] < un && un < (float)bb[2]) && ((float)bb[1] < vn && vn < (float)bb[3]);
}

// SampleConsensusModel{Perpendicular,Parallel}Plane::isModelValid, canonical form shared with oracle/cuboid_oracle.cpp
// (model_valid): coeff[3] = 0, coeff.normalize(), then |axis . coeff| > sin(eps) -> invalid (parallel model) or
// |cos(angle(axis, coeff))| < cos(eps) -> invalid (perpendicular model).
__device__ __forceinline__ bool sac_model_valid(const SacArgs& a, const float c[4]) {
    if (a.model_type == 0 || !(a.cos_eps < 1.0)) return true;
    const float n2 = dot4_sse(c[0], c[1], c[2], 0.0f, c[0], c[1], c[2], 0.0f);
    const float nrm = sqrtf(n2);
    const float q0 = c[0] / nrm, q1 = c[1] / nrm, q2 = c[2] / nrm;
    const float d = dot4_sse(a.axis[0], a.axis[1], a.axis[2], 0.0f, q0, q1, q2, 0.0f);
    if (a.model_type == 2) return !((double)fabsf(d) > a.sin_eps);
    const float nn = dot4_sse(a.axis[0], a.axis[1], a.axis[2], 0.0f, a.axis[0], a.axis[1], a.axis[2], 0.0f) * dot4_sse(q0, q1, q2, 0.0f, q0, q1, q2, 0.0f);
    double rad = (double)(d / sqrtf(nn));
    if (rad < -1.0) rad = -1.0; else if (rad > 1.0) rad = 1.0;
    return !(fabs(rad) < a.cos_eps);
}

// SampleConsensusModelPlane::isSampleGood
__device__ __forceinline__ bool sac_sample_good(const float4& p0, const float4& p1, const float4& p2) {
    const float r0 = (p1.x - p0.x) / (p2.x - p0.x);
    const float r1 = (p1.y - p0.y) / (p2.y - p0.y);
    const float r2 = (p1.z - p0.z) / (p2.z - p0.z);
    return (r0 != r1) || (r2 != r1);
}
// SampleConsensusModelPlane::computeModelCoefficients
__device__ __forceinline__ bool sac_plane_from_sample(const float4& p0, const float4& p1, const float4& p2, float c[4]) {
    const float ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
    const float bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
    const float r0 = ax / bx, r1 = ay / by, r2 = az / bz;
    if ((r0 == r1) && (r2 == r1)) return false;
    float c0 = ay * bz - az * by;
    float c1 = az * bx - ax * bz;
    float c2 = ax * by - ay * bx;
    const float sq = (c0 * c0 + c2 * c2) + (c1 * c1 + 0.0f * 0.0f);
    if (sq > 0.0f) {
        const float nrm = sqrtf(sq);
        c0 = c0 / nrm; c1 = c1 / nrm; c2 = c2 / nrm;
    }
    c[0] = c0; c[1] = c1; c[2] = c2;
    c[3] = -1.0f * dot4_sse(c0, c1, c2, 0.0f, p0.x, p0.y, p0.z, 1.0f);
    return true;
}

// pcl::computeRoots -> roots(0); float trig evaluated as correctly-rounded (double evaluation, one rounding),
// the canonical choice shared with the oracle (glibc 2.23's float libm is not reproducible offline).
__device__ float sac_smallest_root(const float m[3][3]) {
    const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] -
                     m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
    const float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] +
                     m[1][1] * m[2][2] - m[1][2] * m[1][2];
    const float c2 = m[0][0] + m[1][1] + m[2][2];
    if (fabsf(c0) < 1.1920928955078125e-07f) return 0.0f;
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = sqrtf(3.0f);
    const float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    const float rho = sqrtf(-a_over_3);
    const float theta = (float)atan2((double)sqrtf(-q), (double)half_b) * s_inv3;
    const float cos_t = (float)cos((double)theta);
    const float sin_t = (float)sin((double)theta);
    float r0 = c2_over_3 + 2.0f * rho * cos_t;
    float r1 = c2_over_3 - rho * (cos_t + s_sqrt3 * sin_t);
    float r2 = c2_over_3 - rho * (cos_t - s_sqrt3 * sin_t);
    float tmp;
    if (r0 >= r1) { tmp = r0; r0 = r1; r1 = tmp; }
    if (r1 >= r2) {
        tmp = r1; r1 = r2; r2 = tmp;
        if (r0 >= r1) { tmp = r0; r0 = r1; r1 = tmp; }
    }
    if (r0 <= 0.0f) return 0.0f;
    return r0;
}

// tail of optimizeModelCoefficients: accu[9] (already the raw sums) + count -> refined coefficients
__device__ void sac_refine_from_sums(const float* acc_in, int count, float cout[4]) {
    float acc[9];
    const float cnt = (float)count;
    for (int i = 0; i < 9; ++i) acc[i] = acc_in[i] / cnt;
    float cov[3][3];
    cov[0][0] = acc[0] - acc[6] * acc[6];
    cov[0][1] = acc[1] - acc[6] * acc[7];
    cov[0][2] = acc[2] - acc[6] * acc[8];
    cov[1][1] = acc[3] - acc[7] * acc[7];
    cov[1][2] = acc[4] - acc[7] * acc[8];
    cov[2][2] = acc[5] - acc[8] * acc[8];
    cov[1][0] = cov[0][1]; cov[2][0] = cov[0][2]; cov[2][1] = cov[1][2];
    float scale = 0.0f;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = fmaxf(scale, fabsf(cov[i][j]));
    if (scale <= 1.17549435e-38f) scale = 1.0f;
    float sm[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sm[i][j] = cov[i][j] / scale;
    const float ev = sac_smallest_root(sm);
    sm[0][0] -= ev; sm[1][1] -= ev; sm[2][2] -= ev;
    float v[3][3];
    const int ra[3] = {0, 0, 1}, rb[3] = {1, 2, 2};
    float len[3];
    for (int k = 0; k < 3; ++k) {
        const float* x = sm[ra[k]]; const float* y = sm[rb[k]];
        v[k][0] = x[1] * y[2] - x[2] * y[1];
        v[k][1] = x[2] * y[0] - x[0] * y[2];
        v[k][2] = x[0] * y[1] - x[1] * y[0];
        len[k] = v[k][0] * v[k][0] + (v[k][1] * v[k][1] + v[k][2] * v[k][2]);
    }
    int b;
    if (len[0] >= len[1] && len[0] >= len[2]) b = 0;
    else if (len[1] >= len[0] && len[1] >= len[2]) b = 1;
    else b = 2;
    const float sl = sqrtf(len[b]);
    cout[0] = v[b][0] / sl; cout[1] = v[b][1] / sl; cout[2] = v[b][2] / sl;
    cout[3] = -1.0f * dot4_sse(cout[0], cout[1], cout[2], 0.0f, acc[6], acc[7], acc[8], 1.0f);
}

struct SacShared {
    float4 tile[2][SAC_TILE];          // 32 KB; after the scoring loop the same bytes hold float prod[9][SAC_PROD + 1] (18 KB)
    unsigned long long bar[2];
    SacHyp hyp[SAC_HB];
    int counts[SAC_HB];
    int s_w[33];
    float sums[9];
    float best_c[4], coeff[4];
    int n_hyp, done, have_model, rp, n_inl_pre, n_inl, n_rem;
    int iterations, skipped, draws, best, status;
    double k;
};

// ordered compaction of {i : pred(i)} over [0,V) into out_idx / out_pts (either may be NULL); returns count.
template <int NT = SAC_THREADS, typename Pred>
__device__ int sac_compact(int V, Pred pred, int* out_idx, float4* out_pts, const float4* vox, int cap,
                           uint64_t* hash_idx, uint64_t* hash_pts, int* s_w) {
    constexpr int PER = 4;   // consecutive elements per thread and round: a quarter of the block scans of one-per-thread
    int base = 0;
    unsigned long long hi = 0, hp = 0;
    for (int start = 0; start < V; start += NT * PER) {
        const int i0 = start + threadIdx.x * PER;
        float4 p[PER];
        unsigned int keep = 0;
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            p[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (i0 + k < V) { p[k] = vox[i0 + k]; if (pred(i0 + k, p[k])) keep |= 1u << k; }
        }
        int total;
        int pos = base + (NT == 256 ? block_excl_scan256(__popc(keep), s_w, &total) : block_excl_scan<NT>(__popc(keep), s_w, &total));
#pragma unroll
        for (int k = 0; k < PER; ++k) {
            if (!(keep & (1u << k))) continue;
            if (pos < cap) {
                if (out_idx) { out_idx[pos] = i0 + k; hi += hash_index((unsigned int)pos, i0 + k); }
                if (out_pts) { out_pts[pos] = p[k]; hp += hash_point((unsigned int)pos, p[k].x, p[k].y, p[k].z); }
            }
            ++pos;
        }
        base += total;
        __syncthreads();
    }
    if (hash_idx) { hi = warp_sum_u64(hi); if ((threadIdx.x & 31) == 0 && hi) atomic_add_u64(hash_idx, hi); }
    if (hash_pts) { hp = warp_sum_u64(hp); if ((threadIdx.x & 31) == 0 && hp) atomic_add_u64(hash_pts, hp); }
    return base;
}

// NT = 256: one CTA per frame, four per SM (throughput). NT = 1024: launches with so few frames that most SMs would idle (the
// single-frame ROS callback, small 720p batches): the 32 warps split every point tile four ways per hypothesis group, compactions
// and the product staging of the refinement run four times as wide; the sequential parts (draw, accept, the 9 refine chains) are
// the same. Results do not depend on NT.
template <int NT>
__global__ void __launch_bounds__(NT, 1) k_sac_plane(const SacArgs a) {
    constexpr int NSLICE = NT / 256;    // warps per hypothesis group
    extern __shared__ __align__(128) unsigned char smem_raw[];
    SacShared& S = *reinterpret_cast<SacShared*>(smem_raw);
    const int f = blockIdx.x;
    const int V = a.res[f].n_voxels;
    const float4* vox = a.vox + (size_t)f * a.P;
    int* shuffled = a.shuffled + (size_t)f * a.P;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;

    if (threadIdx.x == 0) {
        S.done = (V < 3) ? 1 : 0; S.have_model = 0; S.rp = 0; S.iterations = 0; S.skipped = 0; S.draws = 0;
        S.best = -2147483647; S.k = 1.0; S.status = 0; S.n_hyp = 0;
        mbar_init(&S.bar[0], 1); mbar_init(&S.bar[1], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (!a.triplets) for (int i = threadIdx.x; i < V; i += NT) shuffled[i] = i;
    __syncthreads();

    unsigned int phase0 = 0, phase1 = 0;
    int round = 0;
    while (!S.done) {
        // ---- draw: sequential, thread 0 (SampleConsensusModel::getSamples / drawIndexSample) ----
        if (threadIdx.x == 0) {
            const int hb = (round == 0) ? 8 : SAC_HB;
            int nh = 0;
            int rp = S.rp;
            // draws already replayed count against the table / explicit list
            int next_trip = S.draws + 0;
            while (nh < hb) {
                int s0, s1, s2;
                bool got = false;
                if (a.triplets) {
                    const int d = next_trip + nh;
                    if (d < a.n_triplets) { s0 = a.triplets[3 * d]; s1 = a.triplets[3 * d + 1]; s2 = a.triplets[3 * d + 2]; got = true; }
                } else {
                    for (int chk = 0; chk < 1000 && !got; ++chk) {
                        if (rp + 3 > a.rng_len) { S.status |= CUBOID_W_RNG_EXHAUSTED; break; }
                        for (int i = 0; i < 3; ++i) {
                            const int r = a.rng[rp++];
                            const int j = i + (r % (V - i));
                            const int t0 = shuffled[i]; shuffled[i] = shuffled[j]; shuffled[j] = t0;
                        }
                        s0 = shuffled[0]; s1 = shuffled[1]; s2 = shuffled[2];
                        got = sac_sample_good(vox[s0], vox[s1], vox[s2]);
                    }
                }
                SacHyp& hy = S.hyp[nh];
                if (!got) { hy.kind = 1; hy.valid = 0; hy.ok = 0; ++nh; break; }
                hy.kind = 0;
                hy.valid = sac_plane_from_sample(vox[s0], vox[s1], vox[s2], hy.c) ? 1 : 0;
                hy.ok = (hy.valid && sac_model_valid(a, hy.c)) ? 1 : 0;   // countWithinDistance returns 0 for an invalid model
                ++nh;
            }
            S.rp = rp;
            S.n_hyp = nh;
        }
        __syncthreads();
        const int nh = S.n_hyp;

        // ---- score: warp w owns hypotheses (w % 8), (w % 8) + 8, ... on slice w / 8 of every tile; tiles arrive by TMA bulk copy ----
        float hc[4][4];
        int hcnt[4] = {0, 0, 0, 0};
        bool hact[4];
        const int hgrp = wid & 7, slice = wid >> 3;
        if (NSLICE > 1) { if (threadIdx.x < SAC_HB) S.counts[threadIdx.x] = 0; }
#pragma unroll
        for (int s = 0; s < 4; ++s) {
            const int h = hgrp + 8 * s;
            hact[s] = h < nh && S.hyp[h].valid && S.hyp[h].ok;
#pragma unroll
            for (int c = 0; c < 4; ++c) hc[s][c] = hact[s] ? S.hyp[h].c[c] : 0.f;
        }
        const int ntile = (V + SAC_TILE - 1) / SAC_TILE;
        if (threadIdx.x == 0 && ntile > 0) {
            const unsigned int bytes = (unsigned int)min(SAC_TILE, V) * 16u;
            mbar_expect_tx(&S.bar[0], bytes);
            tma_bulk_g2s(S.tile[0], vox, bytes, &S.bar[0]);
        }
        for (int tl = 0; tl < ntile; ++tl) {
            const int buf = tl & 1;
            if (threadIdx.x == 0 && tl + 1 < ntile) {
                const int nb = buf ^ 1;
                const unsigned int bytes = (unsigned int)min(SAC_TILE, V - (tl + 1) * SAC_TILE) * 16u;
                mbar_expect_tx(&S.bar[nb], bytes);
                tma_bulk_g2s(S.tile[nb], vox + (size_t)(tl + 1) * SAC_TILE, bytes, &S.bar[nb]);
            }
            if (buf == 0) { mbar_wait(&S.bar[0], phase0); phase0 ^= 1; } else { mbar_wait(&S.bar[1], phase1); phase1 ^= 1; }
            const int cntp = min(SAC_TILE, V - tl * SAC_TILE);
            for (int j0 = slice * (SAC_TILE / NSLICE); j0 < min(cntp, (slice + 1) * (SAC_TILE / NSLICE)); j0 += 32) {
                const int j = j0 + lane;
                const bool in = j < cntp;
                const float4 p = in ? S.tile[buf][j] : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int s = 0; s < 4; ++s) {
                    if (!hact[s]) continue;
                    const bool hit = in && (plane_abs_dist(hc[s], p.x, p.y, p.z) < a.thr_f);
                    hcnt[s] += __popc(__ballot_sync(FULL_MASK, hit));
                }
            }
            __syncthreads();   // everyone is done with buf before it is refilled two tiles later
        }
        if (lane == 0) {
#pragma unroll
            for (int s = 0; s < 4; ++s) {
                const int h = hgrp + 8 * s;
                if (h < SAC_HB) { if (NSLICE > 1) atomicAdd(&S.counts[h], hcnt[s]); else S.counts[h] = hcnt[s]; }
            }
        }
        __syncthreads();

        // ---- accept: sequential replay of RandomSampleConsensus::computeModel over this batch ----
        if (threadIdx.x == 0) {
            const double one_over = 1.0 / (double)V;
            const unsigned int max_skip = (unsigned int)a.max_iter * 10u;
            int done = 0;
            for (int h = 0; h < nh; ++h) {
                if (!((double)S.iterations < S.k && (unsigned int)S.skipped < max_skip)) { done = 1; break; }
                if (S.hyp[h].kind == 1) { done = 1; break; }           // selection.empty(): break
                ++S.draws;
                if (!S.hyp[h].valid) { ++S.skipped; continue; }
                const int cnt = S.counts[h];
                if (cnt > S.best) {
                    S.best = cnt;
                    for (int c = 0; c < 4; ++c) S.best_c[c] = S.hyp[h].c[c];
                    S.have_model = 1;
                    const double w = (double)cnt * one_over;
                    double p_no = 1.0 - pow(w, 3.0);
                    p_no = fmax(2.220446049250313e-16, p_no);
                    p_no = fmin(1.0 - 2.220446049250313e-16, p_no);
                    S.k = a.log_prob / log(p_no);
                }
                ++S.iterations;
                if (S.iterations > a.max_iter) { done = 1; break; }
            }
            if (!done && !((double)S.iterations < S.k && (unsigned int)S.skipped < max_skip)) done = 1;
            if (a.triplets && S.draws >= a.n_triplets) done = 1;
            S.done = done;
        }
        __syncthreads();
        ++round;
    }

    // ---- SACSegmentation::segment tail: inliers, refine, reselect, extract ----
    cuboid_frame_result& R = a.res[f];
    int* inl_pre = a.inl_pre + (size_t)f * a.P;
    int* inl = a.inl + (size_t)f * a.P;
    float4* remain = a.remain + (size_t)f * a.P;
    const int have = S.have_model;
    float c[4];
    int n_pre = 0;
    if (have) {
#pragma unroll
        for (int k = 0; k < 4; ++k) c[k] = S.best_c[k];
        const float thr = a.thr_f;
        const bool ok_pre = sac_model_valid(a, c);   // selectWithinDistance clears the inliers of an invalid model
        n_pre = sac_compact<NT>(V, [&](int, const float4& p) { return ok_pre && plane_abs_dist(c, p.x, p.y, p.z) < thr; }, inl_pre, nullptr, vox,
                            a.P, nullptr, nullptr, S.s_w);
        if (a.refine && n_pre >= 4) {
            // PCL accumulates the 9 sums sequentially in float over the inliers in index order.
            if (threadIdx.x < 9) S.sums[threadIdx.x] = 0.0f;
            __syncthreads();
            float (*prod)[SAC_PROD + 1] = reinterpret_cast<float (*)[SAC_PROD + 1]>(&S.tile[0][0]);   // padded rows: conflict-free lanes
            for (int start = 0; start < n_pre; start += SAC_PROD) {
                const int m = min(SAC_PROD, n_pre - start);
                for (int j = threadIdx.x; j < m; j += NT) {
                    const float4 p = vox[inl_pre[start + j]];
                    prod[0][j] = p.x * p.x; prod[1][j] = p.x * p.y; prod[2][j] = p.x * p.z;
                    prod[3][j] = p.y * p.y; prod[4][j] = p.y * p.z; prod[5][j] = p.z * p.z;
                    prod[6][j] = p.x; prod[7][j] = p.y; prod[8][j] = p.z;
                }
                __syncthreads();
                if (threadIdx.x < 9) {
                    float acc = S.sums[threadIdx.x];
                    const float* row = prod[threadIdx.x];
#pragma unroll 8
                    for (int j = 0; j < m; ++j) acc += row[j];
                    S.sums[threadIdx.x] = acc;
                }
                __syncthreads();
            }
            if (threadIdx.x == 0) sac_refine_from_sums(S.sums, n_pre, S.coeff);
            __syncthreads();
#pragma unroll
            for (int k = 0; k < 4; ++k) c[k] = S.coeff[k];
        }
    }
    int n_inl = 0;
    if (have) {
        const float thr = a.thr_f;
        const bool same = !(a.refine && n_pre >= 4);
        (void)same;
        const bool ok_fin = sac_model_valid(a, c);
        n_inl = sac_compact<NT>(V, [&](int, const float4& p) { return ok_fin && plane_abs_dist(c, p.x, p.y, p.z) < thr; }, inl, nullptr, vox, a.P,
                            &R.inlier_hash, nullptr, S.s_w);
    }
    // ExtractIndices (+ PassThrough z2): negative -> everything that is not an inlier, ascending
    const float thr = a.thr_f;
    const int neg = a.negative, uz2 = a.use_z2;
    const bool ok_ext = have && sac_model_valid(a, c);
    const float z2lo = a.z2_lo, z2hi = a.z2_hi;
    const int n_rem = sac_compact<NT>(
        V,
        [&](int, const float4& p) {
            const bool is_in = have && ok_ext && (plane_abs_dist(c, p.x, p.y, p.z) < thr);
            bool keep = neg ? !is_in : is_in;
            if (keep && uz2) keep = finite_f32(p.z) && !(p.z > z2hi || p.z < z2lo);
            if (keep && a.use_bbox) keep = bbox_within(a.bbP, a.bb, p.x, p.y, p.z);
            return keep;
        },
        nullptr, remain, vox, a.cap_remain, nullptr, &R.remain_hash, S.s_w);
    if (threadIdx.x == 0) {
        R.plane_found = have;
        for (int k = 0; k < 4; ++k) R.plane_coeff[k] = have ? c[k] : 0.f;
        R.n_inliers_pre = n_pre;
        R.n_inliers = n_inl;
        R.sac_iterations = S.iterations;
        R.sac_draws = S.draws;
        R.n_remain = min(n_rem, a.cap_remain);
        int st = S.status;
        if (n_rem > a.cap_remain) st |= CUBOID_W_CLUSTERS_TRUNCATED;
        if (st) atomicOr(&R.status, st);
        a.scr[f].best_count = S.best;
    }
}

// stand-alone bbox_filter node (cuboid_detection/src/bbox_filter.cpp:89-103): ordered compaction of one cloud
struct BboxArgs { const float4* pts; int n; double P[12]; int bb[4]; int* idx_out; float4* pts_out; int* n_out; };
__global__ void __launch_bounds__(SAC_THREADS) k_bbox_filter(const BboxArgs a) {
    __shared__ int s_w[9];
    const int n = sac_compact(a.n, [&](int, const float4& p) { return bbox_within(a.P, a.bb, p.x, p.y, p.z); }, a.idx_out, a.pts_out, a.pts,
                              a.n, nullptr, nullptr, s_w);
    if (threadIdx.x == 0) *a.n_out = n;
}

}  // namespace cuboid
