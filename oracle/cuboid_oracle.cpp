/*
 * cuboid_oracle.cpp — CPU restatement of the reference hot path. See cuboid_oracle.h.
 *
 * TEST INFRASTRUCTURE ONLY — never linked into libcuboid_cuda. PARITY UNPINNED (see header).
 * Build: g++ -O2 -std=c++17 -ffp-contract=off -fPIC -shared (oracle/Makefile). No FMA, no fast-math:
 * x86-64 PCL binaries of the reference's era are SSE2.
 *
 * Abbreviations for reference call sites:
 *   gps.cpp = /root/reference/cuboid_detection/src/ground_plane_segmentation.cpp
 *   icp.cpp = /root/reference/cuboid_detection/src/iterative_closest_point.cpp
 *   opd.cpp = /root/reference/object_detection/src/object_pose_detection.cpp
 */
#include "cuboid_oracle.h"

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <limits>
#include <numeric>
#include <unordered_map>
#include <vector>

namespace {

struct P4 { float x, y, z, w; };

inline uint64_t splitmix64(uint64_t x) {
    x += 0x9E3779B97F4A7C15ull;
    x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
    x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
    return x ^ (x >> 31);
}
inline uint32_t fbits(float f) { uint32_t u; std::memcpy(&u, &f, 4); return u; }

/* Eigen SSE2 4-float dot: predux = (a0+a2)+(a1+a3) */
inline float dot4(float a0, float a1, float a2, float a3, float b0, float b1, float b2, float b3) {
    return (a0 * b0 + a2 * b2) + (a1 * b1 + a3 * b3);
}

/* ------------------------------------------------------------------ canonical reductions ----
 * 256 strided lanes (lane l sums elements l, l+256, ... in ascending order from 0.0), xor-butterfly
 * 16,8,4,2,1 inside each group of 32 lanes, then the 8 group results added left to right. This is
 * exactly what one 256-thread CTA does with __shfl_xor_sync + an 8-entry shared array. */
template <typename T, typename F>
T canon_reduce(int n, F elem) {
    T lane[256];
    for (int l = 0; l < 256; ++l) lane[l] = T(0);
    for (int i = 0; i < n; ++i) lane[i & 255] = lane[i & 255] + elem(i);
    for (int off = 16; off >= 1; off >>= 1) {
        T nxt[256];
        for (int l = 0; l < 256; ++l) nxt[l] = lane[l] + lane[l ^ off];
        for (int l = 0; l < 256; ++l) lane[l] = nxt[l];
    }
    T s = lane[0];
    for (int g = 1; g < 8; ++g) s = s + lane[32 * g];
    return s;
}
template <typename T, typename F>
T seq_reduce(int n, F elem) {
    T s = T(0);
    for (int i = 0; i < n; ++i) s = s + elem(i);
    return s;
}
template <typename T, typename F>
T mode_reduce(int mode, int n, F elem) {
    return mode == ORC_LITERAL ? seq_reduce<T>(n, elem) : canon_reduce<T>(n, elem);
}

/* ------------------------------------------------------------------ mt19937 (boost::mt19937) */
struct MT19937 {
    uint32_t mt[624];
    int idx;
    explicit MT19937(uint32_t seed) {
        mt[0] = seed;
        for (int i = 1; i < 624; ++i) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
        idx = 624;
    }
    uint32_t next() {
        if (idx >= 624) {
            for (int i = 0; i < 624; ++i) {
                uint32_t y = (mt[i] & 0x80000000u) | (mt[(i + 1) % 624] & 0x7fffffffu);
                mt[i] = mt[(i + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
            }
            idx = 0;
        }
        uint32_t y = mt[idx++];
        y ^= y >> 11;
        y ^= (y << 7) & 0x9d2c5680u;
        y ^= (y << 15) & 0xefc60000u;
        y ^= y >> 18;
        return y;
    }
};

/* ------------------------------------------------------------------ plane model (A.3) */
/* SampleConsensusModelPlane::isSampleGood */
bool sample_good(const P4* p, const int s[3]) {
    const P4 &p0 = p[s[0]], &p1 = p[s[1]], &p2 = p[s[2]];
    const float r0 = (p1.x - p0.x) / (p2.x - p0.x);
    const float r1 = (p1.y - p0.y) / (p2.y - p0.y);
    const float r2 = (p1.z - p0.z) / (p2.z - p0.z);
    return (r0 != r1) || (r2 != r1);
}
/* SampleConsensusModelPlane::computeModelCoefficients */
bool plane_from_sample(const P4* p, const int s[3], float c[4]) {
    const P4 &p0 = p[s[0]], &p1 = p[s[1]], &p2 = p[s[2]];
    const float ax = p1.x - p0.x, ay = p1.y - p0.y, az = p1.z - p0.z;
    const float bx = p2.x - p0.x, by = p2.y - p0.y, bz = p2.z - p0.z;
    const float r0 = ax / bx, r1 = ay / by, r2 = az / bz;
    if ((r0 == r1) && (r2 == r1)) return false;
    float c0 = ay * bz - az * by;
    float c1 = az * bx - ax * bz;
    float c2 = ax * by - ay * bx;
    /* VectorXf::normalize(): z = squaredNorm (SSE2 predux order, 4th lane 0); if (z > 0) v /= sqrt(z) */
    const float sq = (c0 * c0 + c2 * c2) + (c1 * c1 + 0.0f * 0.0f);
    if (sq > 0.0f) {
        const float nrm = std::sqrt(sq);
        c0 = c0 / nrm; c1 = c1 / nrm; c2 = c2 / nrm;
    }
    c[0] = c0; c[1] = c1; c[2] = c2;
    c[3] = -1.0f * dot4(c0, c1, c2, 0.0f, p0.x, p0.y, p0.z, 1.0f);
    return true;
}
inline float plane_dist(const float c[4], const P4& q) {
    return std::fabs(dot4(c[0], c[1], c[2], c[3], q.x, q.y, q.z, 1.0f));
}

/* correctly-rounded float trig for the canonical mode: evaluate in double, round once.
 * (glibc 2.23's atan2f/sinf/cosf — what the reference ran on — are not reproducible offline.) */
inline float cr_atan2f(float y, float x) { return (float)std::atan2((double)y, (double)x); }
inline float cr_cosf(float a) { return (float)std::cos((double)a); }
inline float cr_sinf(float a) { return (float)std::sin((double)a); }

/* pcl::computeRoots restricted to what pcl::eigen33(mat, eigenvalue, eigenvector) consumes: roots(0) */
float smallest_root(const float m[3][3], int mode) {
    const float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] -
                     m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
    const float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] +
                     m[1][1] * m[2][2] - m[1][2] * m[1][2];
    const float c2 = m[0][0] + m[1][1] + m[2][2];
    if (std::fabs(c0) < FLT_EPSILON) return 0.0f; /* computeRoots2: roots(0) = 0 */
    const float s_inv3 = (float)(1.0 / 3.0);
    const float s_sqrt3 = std::sqrt(3.0f);
    const float c2_over_3 = c2 * s_inv3;
    float a_over_3 = (c1 - c2 * c2_over_3) * s_inv3;
    if (a_over_3 > 0.0f) a_over_3 = 0.0f;
    const float half_b = 0.5f * (c0 + c2_over_3 * (2.0f * c2_over_3 * c2_over_3 - c1));
    float q = half_b * half_b + a_over_3 * a_over_3 * a_over_3;
    if (q > 0.0f) q = 0.0f;
    const float rho = std::sqrt(-a_over_3);
    float theta, cos_t, sin_t;
    if (mode == ORC_LITERAL) {
        theta = std::atan2(std::sqrt(-q), half_b) * s_inv3;
        cos_t = std::cos(theta);
        sin_t = std::sin(theta);
    } else {
        theta = cr_atan2f(std::sqrt(-q), half_b) * s_inv3;
        cos_t = cr_cosf(theta);
        sin_t = cr_sinf(theta);
    }
    float r0 = c2_over_3 + 2.0f * rho * cos_t;
    float r1 = c2_over_3 - rho * (cos_t + s_sqrt3 * sin_t);
    float r2 = c2_over_3 - rho * (cos_t - s_sqrt3 * sin_t);
    if (r0 >= r1) std::swap(r0, r1);
    if (r1 >= r2) {
        std::swap(r1, r2);
        if (r0 >= r1) std::swap(r0, r1);
    }
    if (r0 <= 0.0f) return 0.0f;
    return r0;
}

/* SampleConsensusModelPlane::optimizeModelCoefficients = computeMeanAndCovarianceMatrix + pcl::eigen33 */
void plane_refine(const P4* p, const std::vector<int32_t>& inl, const float cin[4], float cout[4], int mode) {
    if (inl.size() < 4) { std::memcpy(cout, cin, 16); return; }
    float acc[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    for (int32_t id : inl) { /* sequential float sums, PCL's own loop order */
        const P4& q = p[id];
        acc[0] += q.x * q.x; acc[1] += q.x * q.y; acc[2] += q.x * q.z;
        acc[3] += q.y * q.y; acc[4] += q.y * q.z; acc[5] += q.z * q.z;
        acc[6] += q.x; acc[7] += q.y; acc[8] += q.z;
    }
    const float cnt = (float)inl.size();
    for (float& a : acc) a = a / cnt;
    float cov[3][3];
    cov[0][0] = acc[0] - acc[6] * acc[6];
    cov[0][1] = acc[1] - acc[6] * acc[7];
    cov[0][2] = acc[2] - acc[6] * acc[8];
    cov[1][1] = acc[3] - acc[7] * acc[7];
    cov[1][2] = acc[4] - acc[7] * acc[8];
    cov[2][2] = acc[5] - acc[8] * acc[8];
    cov[1][0] = cov[0][1]; cov[2][0] = cov[0][2]; cov[2][1] = cov[1][2];
    /* pcl::eigen33 */
    float scale = 0.0f;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(cov[i][j]));
    if (scale <= FLT_MIN) scale = 1.0f;
    float sm[3][3];
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) sm[i][j] = cov[i][j] / scale;
    const float ev = smallest_root(sm, mode);
    sm[0][0] -= ev; sm[1][1] -= ev; sm[2][2] -= ev;
    auto cross = [](const float* a, const float* b, float* o) {
        o[0] = a[1] * b[2] - a[2] * b[1];
        o[1] = a[2] * b[0] - a[0] * b[2];
        o[2] = a[0] * b[1] - a[1] * b[0];
    };
    float v1[3], v2[3], v3[3];
    cross(sm[0], sm[1], v1); cross(sm[0], sm[2], v2); cross(sm[1], sm[2], v3);
    auto sqn = [](const float* v) { return v[0] * v[0] + (v[1] * v[1] + v[2] * v[2]); }; /* Eigen size-3 redux tree */
    const float l1 = sqn(v1), l2 = sqn(v2), l3 = sqn(v3);
    const float* best; float len;
    if (l1 >= l2 && l1 >= l3) { best = v1; len = l1; }
    else if (l2 >= l1 && l2 >= l3) { best = v2; len = l2; }
    else { best = v3; len = l3; }
    const float sl = std::sqrt(len);
    cout[0] = best[0] / sl; cout[1] = best[1] / sl; cout[2] = best[2] / sl;
    cout[3] = -1.0f * dot4(cout[0], cout[1], cout[2], 0.0f, acc[6], acc[7], acc[8], 1.0f);
}

/* ------------------------------------------------------------------ exact NN (A.6) ----------
 * KD-tree over the template, exact, ties -> lowest index, so it equals a brute-force scan with a
 * strict '<' update in index order. Pruning uses fl((q_d - split)^2), a true lower bound of the
 * float distance ((dx*dx)+dy*dy)+dz*dz because float rounding is monotone. */
struct KdTree {
    struct Node { int lo, hi, dim; float split; int left, right; };
    std::vector<Node> nodes;
    std::vector<int> perm;
    const P4* pts = nullptr;
    static float coord(const P4& p, int d) { return d == 0 ? p.x : (d == 1 ? p.y : p.z); }
    int build(int lo, int hi) {
        Node nd{lo, hi, -1, 0.f, -1, -1};
        const int id = (int)nodes.size();
        nodes.push_back(nd);
        if (hi - lo > 12) {
            float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
            for (int i = lo; i < hi; ++i) for (int d = 0; d < 3; ++d) {
                const float c = coord(pts[perm[i]], d);
                mn[d] = std::min(mn[d], c); mx[d] = std::max(mx[d], c);
            }
            int dim = 0;
            for (int d = 1; d < 3; ++d) if (mx[d] - mn[d] > mx[dim] - mn[dim]) dim = d;
            if (mx[dim] > mn[dim]) {
                const int mid = (lo + hi) / 2;
                std::nth_element(perm.begin() + lo, perm.begin() + mid, perm.begin() + hi,
                                 [&](int a, int b) { return coord(pts[a], dim) < coord(pts[b], dim); });
                const float split = coord(pts[perm[mid]], dim);
                const int l = build(lo, mid);
                const int r = build(mid, hi);
                nodes[id].dim = dim; nodes[id].split = split; nodes[id].left = l; nodes[id].right = r;
            }
        }
        return id;
    }
    void init(const P4* p, int n) {
        pts = p; perm.resize(n); std::iota(perm.begin(), perm.end(), 0);
        nodes.clear(); nodes.reserve(n / 4 + 8);
        if (n > 0) build(0, n);
    }
    static float d2(const P4& a, const P4& b) {
        const float dx = a.x - b.x, dy = a.y - b.y, dz = a.z - b.z;
        return ((dx * dx) + dy * dy) + dz * dz;
    }
    void search(int id, const P4& q, float& best, int& bi) const {
        const Node& nd = nodes[id];
        if (nd.dim < 0) {
            for (int i = nd.lo; i < nd.hi; ++i) {
                const int j = perm[i];
                const float d = d2(q, pts[j]);
                if (d < best || (d == best && j < bi)) { best = d; bi = j; }
            }
            return;
        }
        const float diff = coord(q, nd.dim) - nd.split;
        const int near = diff < 0.f ? nd.left : nd.right;
        const int far = diff < 0.f ? nd.right : nd.left;
        search(near, q, best, bi);
        if (!(diff * diff > best)) search(far, q, best, bi);
    }
    int nearest(const P4& q, float& dist) const {
        float best = std::numeric_limits<float>::infinity();
        int bi = INT_MAX;
        if (!nodes.empty()) search(0, q, best, bi);
        dist = best;
        return bi == INT_MAX ? -1 : bi;
    }
};

/* ------------------------------------------------------------------ small fixed-size linear algebra */
struct M3 { float a[3][3]; };
struct M4 { float a[4][4]; };

inline M4 m4_identity() { M4 m{}; for (int i = 0; i < 4; ++i) m.a[i][i] = 1.f; return m; }
/* Eigen fixed 4x4 * 4x4 float: column-wise ((l0*r0 + l1*r1) + l2*r2) + l3*r3 */
inline M4 m4_mul(const M4& l, const M4& r) {
    M4 o;
    for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j)
        o.a[i][j] = ((l.a[i][0] * r.a[0][j] + l.a[i][1] * r.a[1][j]) + l.a[i][2] * r.a[2][j]) + l.a[i][3] * r.a[3][j];
    return o;
}
/* tr * (x,y,z,1): ((m0*x + m1*y) + m2*z) + m3 */
inline P4 m4_apply(const M4& m, const P4& p) {
    P4 o;
    o.x = ((m.a[0][0] * p.x + m.a[0][1] * p.y) + m.a[0][2] * p.z) + m.a[0][3];
    o.y = ((m.a[1][0] * p.x + m.a[1][1] * p.y) + m.a[1][2] * p.z) + m.a[1][3];
    o.z = ((m.a[2][0] * p.x + m.a[2][1] * p.y) + m.a[2][2] * p.z) + m.a[2][3];
    o.w = 1.0f;
    return o;
}
inline float det3(const M3& m) {
    auto h = [&](int a, int b, int c) { return m.a[0][a] * (m.a[1][b] * m.a[2][c] - m.a[1][c] * m.a[2][b]); };
    return h(0, 1, 2) - h(1, 0, 2) + h(2, 0, 1);
}

struct Rot { float c, s; };
/* Eigen internal::apply_rotation_in_the_plane on two strided 3-vectors */
inline void rot_rows(M3& m, int p, int q, Rot j) {
    if (j.c == 1.f && j.s == 0.f) return;
    for (int i = 0; i < 3; ++i) {
        const float xi = m.a[p][i], yi = m.a[q][i];
        m.a[p][i] = j.c * xi + j.s * yi;
        m.a[q][i] = -j.s * xi + j.c * yi;
    }
}
inline void rot_cols(M3& m, int p, int q, Rot j) { /* applyOnTheRight(p,q,j) = rotation with j.transpose() */
    const Rot t{j.c, -j.s};
    if (t.c == 1.f && t.s == 0.f) return;
    for (int i = 0; i < 3; ++i) {
        const float xi = m.a[i][p], yi = m.a[i][q];
        m.a[i][p] = t.c * xi + t.s * yi;
        m.a[i][q] = -t.s * xi + t.c * yi;
    }
}
/* JacobiRotation::makeJacobi(x, y, z) */
inline Rot make_jacobi(float x, float y, float z) {
    const float deno = 2.0f * std::fabs(y);
    if (deno < FLT_MIN) return Rot{1.f, 0.f};
    const float tau = (x - z) / deno;
    const float w = std::sqrt(tau * tau + 1.0f);
    float t;
    if (tau > 0.f) t = 1.0f / (tau + w); else t = 1.0f / (tau - w);
    const float sign_t = t > 0.f ? 1.0f : -1.0f;
    const float n = 1.0f / std::sqrt(t * t + 1.0f);
    Rot r;
    r.s = -sign_t * (y / std::fabs(y)) * std::fabs(t) * n;
    r.c = n;
    return r;
}
/* Eigen internal::real_2x2_jacobi_svd */
inline void jacobi_2x2(const M3& w, int p, int q, Rot* jl, Rot* jr) {
    float m00 = w.a[p][p], m01 = w.a[p][q], m10 = w.a[q][p], m11 = w.a[q][q];
    Rot rot1;
    const float t = m00 + m11;
    const float d = m10 - m01;
    if (std::fabs(d) < FLT_MIN) { rot1.s = 0.f; rot1.c = 1.f; }
    else {
        const float u = t / d;
        const float tmp = std::sqrt(1.0f + u * u);
        rot1.s = 1.0f / tmp;
        rot1.c = u / tmp;
    }
    /* m.applyOnTheLeft(0,1,rot1) */
    if (!(rot1.c == 1.f && rot1.s == 0.f)) {
        const float x0 = m00, x1 = m01, y0 = m10, y1 = m11;
        m00 = rot1.c * x0 + rot1.s * y0; m01 = rot1.c * x1 + rot1.s * y1;
        m10 = -rot1.s * x0 + rot1.c * y0; m11 = -rot1.s * x1 + rot1.c * y1;
    }
    *jr = make_jacobi(m00, m01, m11);
    /* *j_left = rot1 * j_right->transpose() */
    const Rot jt{jr->c, -jr->s};
    jl->c = rot1.c * jt.c - rot1.s * jt.s;
    jl->s = rot1.c * jt.s + rot1.s * jt.c;
}
/* Eigen::JacobiSVD<Matrix3f>(sigma, ComputeFullU | ComputeFullV) */
void jacobi_svd3(const M3& in, M3& U, M3& V, float sv[3]) {
    const float precision = 2.0f * FLT_EPSILON;
    const float consider_zero = FLT_MIN;
    float scale = 0.f;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) scale = std::max(scale, std::fabs(in.a[i][j]));
    if (scale == 0.f) scale = 1.f;
    M3 W;
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) W.a[i][j] = in.a[i][j] / scale;
    U = M3{}; V = M3{};
    for (int i = 0; i < 3; ++i) { U.a[i][i] = 1.f; V.a[i][i] = 1.f; }
    float max_diag = std::max(std::fabs(W.a[0][0]), std::max(std::fabs(W.a[1][1]), std::fabs(W.a[2][2])));
    bool finished = false;
    int guard = 0;
    while (!finished && guard++ < 1000) {
        finished = true;
        for (int p = 1; p < 3; ++p) for (int q = 0; q < p; ++q) {
            const float thr = std::max(consider_zero, precision * max_diag);
            if (std::fabs(W.a[p][q]) > thr || std::fabs(W.a[q][p]) > thr) {
                finished = false;
                Rot jl, jr;
                jacobi_2x2(W, p, q, &jl, &jr);
                rot_rows(W, p, q, jl);
                rot_cols(U, p, q, Rot{jl.c, -jl.s}); /* U.applyOnTheRight(p,q,j_left.transpose()) */
                rot_cols(W, p, q, jr);
                rot_cols(V, p, q, jr);
                max_diag = std::max(max_diag, std::max(std::fabs(W.a[p][p]), std::fabs(W.a[q][q])));
            }
        }
    }
    for (int i = 0; i < 3; ++i) {
        const float a = W.a[i][i];
        sv[i] = std::fabs(a);
        if (a < 0.f) for (int r = 0; r < 3; ++r) U.a[r][i] = -U.a[r][i];
    }
    for (int i = 0; i < 3; ++i) sv[i] = sv[i] * scale;
    for (int i = 0; i < 3; ++i) {
        int pos = i;
        for (int k = i + 1; k < 3; ++k) if (sv[k] > sv[pos]) pos = k;
        if (sv[pos] == 0.f) break;
        if (pos != i) {
            std::swap(sv[i], sv[pos]);
            for (int r = 0; r < 3; ++r) { std::swap(U.a[r][i], U.a[r][pos]); std::swap(V.a[r][i], V.a[r][pos]); }
        }
    }
}

/* pcl::umeyama(src, dst, false) as called by TransformationEstimationSVD (A.6) */
M4 umeyama(const std::vector<P4>& src, const P4* tgt, const std::vector<int32_t>& corr, int mode) {
    const int n = (int)src.size();
    const float one_over_n = 1.0f / (float)n;
    float sm[3], dm[3];
    sm[0] = mode_reduce<float>(mode, n, [&](int i) { return src[i].x; }) * one_over_n;
    sm[1] = mode_reduce<float>(mode, n, [&](int i) { return src[i].y; }) * one_over_n;
    sm[2] = mode_reduce<float>(mode, n, [&](int i) { return src[i].z; }) * one_over_n;
    dm[0] = mode_reduce<float>(mode, n, [&](int i) { return tgt[corr[i]].x; }) * one_over_n;
    dm[1] = mode_reduce<float>(mode, n, [&](int i) { return tgt[corr[i]].y; }) * one_over_n;
    dm[2] = mode_reduce<float>(mode, n, [&](int i) { return tgt[corr[i]].z; }) * one_over_n;
    auto sc = [&](int i, int c) { const P4& p = src[i]; return (c == 0 ? p.x : (c == 1 ? p.y : p.z)) - sm[c]; };
    auto dc = [&](int i, int c) { const P4& p = tgt[corr[i]]; return (c == 0 ? p.x : (c == 1 ? p.y : p.z)) - dm[c]; };
    M3 sigma;
    for (int r = 0; r < 3; ++r) for (int c = 0; c < 3; ++c)
        sigma.a[r][c] = one_over_n * mode_reduce<float>(mode, n, [&](int i) { return dc(i, r) * sc(i, c); });
    M3 U, V; float sv[3];
    jacobi_svd3(sigma, U, V, sv);
    float S[3] = {1.f, 1.f, 1.f};
    if (det3(U) * det3(V) < 0.f) S[2] = -1.f;
    M4 Rt = m4_identity();
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
        Rt.a[i][j] = ((U.a[i][0] * S[0]) * V.a[j][0] + (U.a[i][1] * S[1]) * V.a[j][1]) + (U.a[i][2] * S[2]) * V.a[j][2];
    for (int i = 0; i < 3; ++i)
        Rt.a[i][3] = dm[i] - ((Rt.a[i][0] * sm[0] + Rt.a[i][1] * sm[1]) + Rt.a[i][2] * sm[2]);
    return Rt;
}

/* ------------------------------------------------------------------ stage implementations */
int voxel_grid(const P4* p, int n, float leaf, int mode, std::vector<P4>& out, int32_t* key_per_point,
               std::vector<int32_t>* vkeys, std::vector<int32_t>* vcounts, int32_t min_b[3], int32_t div_b[3],
               int* overflow) {
    out.clear();
    if (overflow) *overflow = 0;
    for (int a = 0; a < 3; ++a) { min_b[a] = 0; div_b[a] = 0; }
    if (n <= 0) return 0;
    const float inv = 1.0f / leaf;
    float mn[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, mx[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int n_finite = 0;
    for (int i = 0; i < n; ++i) {
        if (!std::isfinite(p[i].x) || !std::isfinite(p[i].y) || !std::isfinite(p[i].z)) continue;
        ++n_finite;
        mn[0] = std::min(mn[0], p[i].x); mx[0] = std::max(mx[0], p[i].x);
        mn[1] = std::min(mn[1], p[i].y); mx[1] = std::max(mx[1], p[i].y);
        mn[2] = std::min(mn[2], p[i].z); mx[2] = std::max(mx[2], p[i].z);
    }
    if (n_finite == 0) return 0;
    int64_t d64[3];
    int32_t max_b[3];
    for (int a = 0; a < 3; ++a) {
        d64[a] = (int64_t)((mx[a] - mn[a]) * inv) + 1;
        min_b[a] = (int32_t)std::floor(mn[a] * inv);
        max_b[a] = (int32_t)std::floor(mx[a] * inv);
        div_b[a] = max_b[a] - min_b[a] + 1;
    }
    if (overflow && d64[0] * d64[1] * d64[2] > (int64_t)INT32_MAX) *overflow = 1;
    const int32_t mul1 = div_b[0], mul2 = (int32_t)((uint32_t)div_b[0] * (uint32_t)div_b[1]);
    struct Rec { int32_t idx; int32_t cp; };
    std::vector<Rec> rec;
    rec.reserve(n);
    for (int i = 0; i < n; ++i) {
        if (!std::isfinite(p[i].x) || !std::isfinite(p[i].y) || !std::isfinite(p[i].z)) {
            if (key_per_point) key_per_point[i] = -1;
            continue;
        }
        const int32_t i0 = (int32_t)(std::floor(p[i].x * inv) - (float)min_b[0]);
        const int32_t i1 = (int32_t)(std::floor(p[i].y * inv) - (float)min_b[1]);
        const int32_t i2 = (int32_t)(std::floor(p[i].z * inv) - (float)min_b[2]);
        const int32_t idx = (int32_t)((uint32_t)i0 + (uint32_t)i1 * (uint32_t)mul1 + (uint32_t)i2 * (uint32_t)mul2);
        if (key_per_point) key_per_point[i] = idx;
        rec.push_back(Rec{idx, i});
    }
    if (mode == ORC_LITERAL)
        std::sort(rec.begin(), rec.end(), [](const Rec& a, const Rec& b) { return a.idx < b.idx; });
    else
        std::stable_sort(rec.begin(), rec.end(), [](const Rec& a, const Rec& b) { return a.idx < b.idx; });
    size_t cp = 0;
    while (cp < rec.size()) {
        float cx = p[rec[cp].cp].x, cy = p[rec[cp].cp].y, cz = p[rec[cp].cp].z;
        size_t i = cp + 1;
        while (i < rec.size() && rec[i].idx == rec[cp].idx) {
            cx += p[rec[i].cp].x; cy += p[rec[i].cp].y; cz += p[rec[i].cp].z;
            ++i;
        }
        const float cnt = (float)(i - cp);
        out.push_back(P4{cx / cnt, cy / cnt, cz / cnt, 1.0f});
        if (vkeys) vkeys->push_back(rec[cp].idx);
        if (vcounts) vcounts->push_back((int32_t)(i - cp));
        cp = i;
    }
    return (int)out.size();
}

struct SacOut {
    bool found = false;
    float coeff[4] = {0, 0, 0, 0}, coeff_pre[4] = {0, 0, 0, 0};
    std::vector<int32_t> inliers, inliers_pre;
    int iters = 0, draws = 0;
    double k_margin = 1e300;
    std::vector<int32_t> triplets;
};

/* SACMODEL_PERPENDICULAR_PLANE / SACMODEL_PARALLEL_PLANE (surface_normal_estimation.cpp:118-123): the plane model plus
 * isModelValid(). type 0 = SACMODEL_PLANE, 1 = perpendicular plane (normal within eps of the axis, either sense),
 * 2 = parallel plane (normal within eps of perpendicular to the axis). [PCL-recall]: coeff[3] = 0; coeff.normalize();
 * perpendicular: min(angle, pi - angle) > eps_angle_ -> invalid; parallel: |axis . coeff| > sin_angle_ -> invalid.
 * Canonical form of the perpendicular test (shared with the CUDA path): |cos(angle)| < cos(eps) in double, with
 * cos(angle) = (float)(axis . coeff) / sqrtf(|axis|^2 * |coeff|^2); it differs from acos() + compare only when the angle
 * is within an ulp of eps. */
struct SacModel { int type = 0; float axis[3] = {0, 0, 0}; double eps = 0.0; };
bool model_valid(const SacModel* m, const float c[4]) {
    if (!m || m->type == 0 || !(m->eps > 0.0)) return true;
    const float n2 = dot4(c[0], c[1], c[2], 0.0f, c[0], c[1], c[2], 0.0f);
    const float nrm = std::sqrt(n2);
    const float q0 = c[0] / nrm, q1 = c[1] / nrm, q2 = c[2] / nrm;
    const float d = dot4(m->axis[0], m->axis[1], m->axis[2], 0.0f, q0, q1, q2, 0.0f);
    if (m->type == 2) return !((double)std::fabs(d) > std::fabs(std::sin(m->eps)));
    const float nn = dot4(m->axis[0], m->axis[1], m->axis[2], 0.0f, m->axis[0], m->axis[1], m->axis[2], 0.0f) * dot4(q0, q1, q2, 0.0f, q0, q1, q2, 0.0f);
    double rad = (double)(d / std::sqrt(nn));
    if (rad < -1.0) rad = -1.0; else if (rad > 1.0) rad = 1.0;
    return !(std::fabs(rad) < std::cos(m->eps));
}

void select_within(const P4* p, int n, const float c[4], double thr, std::vector<int32_t>& out, const SacModel* m = nullptr) {
    out.clear();
    if (!model_valid(m, c)) return;   /* SampleConsensusModel{Perpendicular,Parallel}Plane::selectWithinDistance */
    for (int i = 0; i < n; ++i) if ((double)plane_dist(c, p[i]) < thr) out.push_back(i);
}

/* pcl::RandomSampleConsensus::computeModel + SACSegmentation::segment tail (A.3) */
void sac_plane(const P4* p, int n, double thr, int max_iter, double prob, uint32_t seed, int refine, int mode,
               const int32_t* trip_in, int n_trip_in, SacOut& o, const SacModel* model = nullptr) {
    o = SacOut();
    if (n < 3) return;
    MT19937 rng(seed);
    std::vector<int32_t> shuffled(n);
    std::iota(shuffled.begin(), shuffled.end(), 0);
    int iterations = 0, best = -INT_MAX;
    double k = 1.0;
    const double log_probability = std::log(1.0 - prob);
    const double one_over_indices = 1.0 / (double)n;
    unsigned skipped = 0;
    const unsigned max_skip = (unsigned)max_iter * 10u;
    float best_c[4] = {0, 0, 0, 0};
    bool have_model = false;
    int draw = 0;
    while ((double)iterations < k && skipped < max_skip) {
        int s[3];
        bool got = false;
        if (trip_in) {
            if (draw >= n_trip_in) break;
            s[0] = trip_in[3 * draw]; s[1] = trip_in[3 * draw + 1]; s[2] = trip_in[3 * draw + 2];
            got = true;
        } else {
            for (int chk = 0; chk < 1000 && !got; ++chk) { /* getSamples: max_sample_checks_ = 1000 */
                for (int i = 0; i < 3; ++i) {
                    const int r = (int)(rng.next() >> 1); /* boost::uniform_int<>(0, INT_MAX) over mt19937 */
                    std::swap(shuffled[i], shuffled[i + (r % (n - i))]);
                }
                s[0] = shuffled[0]; s[1] = shuffled[1]; s[2] = shuffled[2];
                got = sample_good(p, s);
            }
            if (!got) break;
        }
        ++draw;
        o.triplets.push_back(s[0]); o.triplets.push_back(s[1]); o.triplets.push_back(s[2]);
        float c[4];
        if (!plane_from_sample(p, s, c)) { ++skipped; continue; }
        int cnt = 0;   /* countWithinDistance: 0 for a model that fails isModelValid, still an iteration */
        if (model_valid(model, c))
            for (int i = 0; i < n; ++i) if ((double)plane_dist(c, p[i]) < thr) ++cnt;
        if (cnt > best) {
            best = cnt;
            std::memcpy(best_c, c, 16);
            have_model = true;
            const double w = (double)best * one_over_indices;
            double p_no = 1.0 - std::pow(w, 3.0);
            p_no = std::max(std::numeric_limits<double>::epsilon(), p_no);
            p_no = std::min(1.0 - std::numeric_limits<double>::epsilon(), p_no);
            k = log_probability / std::log(p_no);
        }
        ++iterations;
        /* how close the loop test came to flipping on a last-ulp difference of pow/log */
        o.k_margin = std::min(o.k_margin, std::fabs(k - (double)iterations));
        if (iterations > max_iter) break;
    }
    o.iters = iterations;
    o.draws = draw;
    if (!have_model) return;
    o.found = true;
    std::memcpy(o.coeff_pre, best_c, 16);
    select_within(p, n, best_c, thr, o.inliers_pre, model);
    if (refine) {
        plane_refine(p, o.inliers_pre, best_c, o.coeff, mode);
        select_within(p, n, o.coeff, thr, o.inliers, model);
    } else {
        std::memcpy(o.coeff, best_c, 16);
        o.inliers = o.inliers_pre;
    }
}

/* pcl::extractEuclideanClusters -> connected components of {d2 < (float)(tol*tol)} (A.5) */
int cluster(const P4* p, int n, double tol, int min_size, int max_size, std::vector<int32_t>& idx_sorted,
            std::vector<int32_t>& offsets) {
    idx_sorted.clear(); offsets.clear(); offsets.push_back(0);
    if (n <= 0) return 0;
    const float r2 = (float)(tol * tol);
    const double cell = tol > 0 ? tol : 1.0;
    auto ck = [&](float v) { return (int64_t)std::floor((double)v / cell); };
    auto hkey = [](int64_t a, int64_t b, int64_t c) {
        return (uint64_t)(a + (1 << 20)) | ((uint64_t)(b + (1 << 20)) << 21) | ((uint64_t)(c + (1 << 20)) << 42);
    };
    std::unordered_map<uint64_t, std::vector<int32_t>> grid;
    grid.reserve((size_t)n * 2);
    for (int i = 0; i < n; ++i) grid[hkey(ck(p[i].x), ck(p[i].y), ck(p[i].z))].push_back(i);
    std::vector<int32_t> parent(n);
    std::iota(parent.begin(), parent.end(), 0);
    auto find = [&](int a) { while (parent[a] != a) { parent[a] = parent[parent[a]]; a = parent[a]; } return a; };
    for (int i = 0; i < n; ++i) {
        const int64_t cx = ck(p[i].x), cy = ck(p[i].y), cz = ck(p[i].z);
        for (int64_t dx = -1; dx <= 1; ++dx) for (int64_t dy = -1; dy <= 1; ++dy) for (int64_t dz = -1; dz <= 1; ++dz) {
            auto it = grid.find(hkey(cx + dx, cy + dy, cz + dz));
            if (it == grid.end()) continue;
            for (int32_t j : it->second) {
                if (j <= i) continue;
                if (KdTree::d2(p[i], p[j]) < r2) {
                    int a = find(i), b = find(j);
                    if (a != b) { if (a < b) parent[b] = a; else parent[a] = b; }
                }
            }
        }
    }
    std::vector<int32_t> root(n), size(n, 0);
    for (int i = 0; i < n; ++i) { root[i] = find(i); ++size[root[i]]; }
    std::vector<int32_t> kept;
    for (int i = 0; i < n; ++i) if (root[i] == i && size[i] >= min_size && size[i] <= max_size) kept.push_back(i);
    /* canonical order: size descending, ties by smallest member index (root = smallest member) */
    std::sort(kept.begin(), kept.end(), [&](int a, int b) { return size[a] != size[b] ? size[a] > size[b] : a < b; });
    std::vector<int32_t> rank(n, -1);
    for (size_t k = 0; k < kept.size(); ++k) rank[kept[k]] = (int32_t)k;
    std::vector<std::vector<int32_t>> members(kept.size());
    for (int i = 0; i < n; ++i) if (rank[root[i]] >= 0) members[rank[root[i]]].push_back(i);
    for (auto& m : members) { idx_sorted.insert(idx_sorted.end(), m.begin(), m.end()); offsets.push_back((int32_t)idx_sorted.size()); }
    return (int)kept.size();
}

struct IcpOut {
    M4 T = m4_identity();
    double fitness = DBL_MAX;
    int converged = 0, iters = 0, state = ORC_ICP_NOT_CONVERGED;
    uint64_t corr_hash = 0;
    std::vector<P4> aligned;
};

/* pcl::IterativeClosestPoint::computeTransformation + DefaultConvergenceCriteria + getFitnessScore (A.6) */
void icp(const P4* src, int n_src, const P4* tgt, int n_tgt, const KdTree& tree, const M4* guess, int max_iter,
         double tf_eps, double rel_mse, double max_corr_dist, int mode, IcpOut& o, int32_t* corr_trace,
         float* T_trace, int cap_iters) {
    o = IcpOut();
    std::vector<P4> cur(src, src + n_src);
    M4 fin = m4_identity();
    bool guess_is_identity = true;
    if (guess) {
        fin = *guess;
        const M4 id = m4_identity(); /* Eigen `guess != Matrix4::Identity()`: coefficient-wise float compare */
        for (int i = 0; i < 4; ++i) for (int j = 0; j < 4; ++j) if (guess->a[i][j] != id.a[i][j]) guess_is_identity = false;
    }
    if (!guess_is_identity) for (int i = 0; i < n_src; ++i) cur[i] = m4_apply(fin, src[i]);
    const double max_d2 = max_corr_dist * max_corr_dist;
    const double rot_thr = 1.0 - tf_eps, trans_thr = tf_eps, abs_thr = 1e-12;
    double prev_mse = DBL_MAX;
    std::vector<int32_t> corr(n_src);
    int it = 0;
    bool converged = false;
    int state = ORC_ICP_NOT_CONVERGED;
    if (n_tgt > 0) {
        do {
            /* CorrespondenceEstimation::determineCorrespondences, then no rejectors */
            std::vector<P4> ks; std::vector<int32_t> kc; std::vector<float> kd;
            ks.reserve(n_src); kc.reserve(n_src); kd.reserve(n_src);
            for (int i = 0; i < n_src; ++i) {
                float d;
                const int j = tree.nearest(cur[i], d);
                corr[i] = -1;
                if ((double)d > max_d2) continue; /* never at the default sqrt(DBL_MAX) */
                corr[i] = j;
                ks.push_back(cur[i]); kc.push_back(j); kd.push_back(d);
            }
            const int ncorr = (int)kc.size();
            if (ncorr < 3) { state = ORC_ICP_NO_CORRESPONDENCES; converged = false; break; }
            for (int i = 0; i < n_src; ++i)
                o.corr_hash += splitmix64((((uint64_t)it * (uint64_t)n_src + (uint64_t)i) << 32) | (uint32_t)corr[i]);
            if (corr_trace && it < cap_iters) std::memcpy(corr_trace + (size_t)it * n_src, corr.data(), 4 * (size_t)n_src);
            const M4 T = umeyama(ks, tgt, kc, mode);
            for (int i = 0; i < n_src; ++i) cur[i] = m4_apply(T, cur[i]); /* incremental, in place */
            fin = m4_mul(T, fin);
            if (T_trace && it < cap_iters) std::memcpy(T_trace + 16 * it, &T, 64);
            ++it;
            /* DefaultConvergenceCriteria::hasConverged */
            if (it >= max_iter) { converged = true; state = ORC_ICP_ITERATIONS; break; }
            const double cos_angle = 0.5 * (double)(T.a[0][0] + T.a[1][1] + T.a[2][2] - 1.0f);
            const double tsq = (double)(T.a[0][3] * T.a[0][3] + T.a[1][3] * T.a[1][3] + T.a[2][3] * T.a[2][3]);
            if (cos_angle >= rot_thr && tsq <= trans_thr) { converged = true; state = ORC_ICP_TRANSFORM; break; }
            const double mse = mode_reduce<double>(mode, ncorr, [&](int i) { return (double)kd[i]; }) / (double)ncorr;
            if (std::fabs(mse - prev_mse) < abs_thr) { converged = true; state = ORC_ICP_ABS_MSE; break; }
            if (std::fabs(mse - prev_mse) / prev_mse < rel_mse) { converged = true; state = ORC_ICP_REL_MSE; break; }
            prev_mse = mse;
        } while (!converged);
    }
    o.T = fin; o.iters = it; o.converged = converged ? 1 : 0; o.state = state;
    /* output = transformCloud(src, final); getFitnessScore() */
    o.aligned.resize(n_src);
    for (int i = 0; i < n_src; ++i) o.aligned[i] = m4_apply(fin, src[i]);
    if (n_src > 0 && n_tgt > 0) {
        std::vector<float> fd(n_src);
        for (int i = 0; i < n_src; ++i) tree.nearest(o.aligned[i], fd[i]);
        o.fitness = mode_reduce<double>(mode, n_src, [&](int i) { return (double)fd[i]; }) / (double)n_src;
    }
}

uint64_t hash_i32(const int32_t* v, int n) {
    uint64_t h = 0;
    for (int i = 0; i < n; ++i) h += splitmix64(((uint64_t)i << 32) | (uint32_t)v[i]);
    return h;
}
uint64_t hash_pts(const P4* p, int n) {
    uint64_t h = 0;
    for (int i = 0; i < n; ++i)
        h += splitmix64(splitmix64(((uint64_t)i << 32) | fbits(p[i].x)) ^ (((uint64_t)fbits(p[i].y) << 32) | fbits(p[i].z)));
    return h;
}

void guess_about_centroid(const P4* s, int n, const float R[9], int mode, M4& G) {
    const float cn = (float)n;
    float c[3];
    c[0] = mode_reduce<float>(mode, n, [&](int i) { return s[i].x; }) / cn;
    c[1] = mode_reduce<float>(mode, n, [&](int i) { return s[i].y; }) / cn;
    c[2] = mode_reduce<float>(mode, n, [&](int i) { return s[i].z; }) / cn;
    G = m4_identity();
    for (int i = 0; i < 3; ++i) {
        for (int j = 0; j < 3; ++j) G.a[i][j] = R[3 * i + j];
        G.a[i][3] = c[i] - ((R[3 * i] * c[0] + R[3 * i + 1] * c[1]) + R[3 * i + 2] * c[2]);
    }
}

/* stages after the cloud exists (shared by the depth and the PointCloud2 entry) */
int process_points(const orc_params* pr, std::vector<P4>& pts, const P4* tmpl, int n_tmpl, const float* guesses,
                   int mode, orc_frame_result* out) {
    out->n_points = (int)pts.size();
    out->points_hash = hash_pts(pts.data(), (int)pts.size());
    std::vector<P4> vox;
    std::vector<int32_t> kpp(pts.size());
    int ovf = 0;
    voxel_grid(pts.data(), (int)pts.size(), pr->leaf, mode, vox, kpp.data(), nullptr, nullptr, out->min_b, out->div_b, &ovf);
    if (ovf) out->status |= 1;
    out->n_voxels = (int)vox.size();
    out->voxel_key_hash = hash_i32(kpp.data(), (int)kpp.size());
    out->voxel_hash = hash_pts(vox.data(), (int)vox.size());
    SacOut so;
    sac_plane(vox.data(), (int)vox.size(), pr->sac_threshold, pr->sac_max_iter, pr->sac_prob, pr->sac_seed,
              pr->sac_refine, mode, nullptr, 0, so);
    out->plane_found = so.found ? 1 : 0;
    std::memcpy(out->plane_coeff, so.coeff, 16);
    out->n_inliers_pre = (int)so.inliers_pre.size();
    out->n_inliers = (int)so.inliers.size();
    out->sac_iterations = so.iters;
    out->sac_draws = so.draws;
    out->inlier_hash = hash_i32(so.inliers.data(), (int)so.inliers.size());
    /* ExtractIndices (gps.cpp:96-101) */
    std::vector<P4> rem;
    {
        std::vector<char> is_in(vox.size(), 0);
        for (int32_t i : so.inliers) is_in[i] = 1;
        for (size_t i = 0; i < vox.size(); ++i) if ((is_in[i] != 0) != (pr->extract_negative != 0)) rem.push_back(vox[i]);
    }
    if (pr->use_pass_z2) { /* opd.cpp:331-336 */
        std::vector<P4> r2;
        for (const P4& q : rem) if (!((double)q.z > pr->pass_z2_max || (double)q.z < pr->pass_z2_min)) r2.push_back(q);
        rem.swap(r2);
    }
    out->n_remain = (int)rem.size();
    out->remain_hash = hash_pts(rem.data(), (int)rem.size());
    /* clusters (opd.cpp:346-362) or the whole remaining cloud (icp.cpp:156,170) */
    std::vector<int32_t> cidx, coff;
    int ncl;
    if (pr->use_cluster) {
        ncl = cluster(rem.data(), (int)rem.size(), pr->cluster_tol, pr->cluster_min, pr->cluster_max, cidx, coff);
    } else {
        cidx.resize(rem.size()); std::iota(cidx.begin(), cidx.end(), 0);
        coff = {0, (int32_t)rem.size()};
        ncl = rem.empty() ? 0 : 1;
    }
    out->n_clusters = ncl;
    out->cluster_hash = hash_i32(cidx.data(), (int)cidx.size()) + splitmix64((uint64_t)ncl);
    if (n_tmpl <= 0 || !tmpl) return 0;
    KdTree tree;
    tree.init(tmpl, n_tmpl);
    for (int c = 0; c < ncl && c < ORC_MAX_CLUSTERS; ++c) {
        std::vector<P4> src;
        for (int k = coff[c]; k < coff[c + 1]; ++k) src.push_back(rem[cidx[k]]);
        orc_cluster_result& cr = out->cluster[c];
        cr.size = (int)src.size();
        const int ng = std::max(1, (int)pr->n_guess);
        bool have = false;
        for (int g = 0; g < ng; ++g) {
            M4 G = m4_identity();
            if (guesses) {
                if (pr->guess_mode == 1) guess_about_centroid(src.data(), (int)src.size(), guesses + 9 * g, mode, G);
                else std::memcpy(&G, guesses + 16 * g, 64);
            }
            IcpOut io;
            icp(src.data(), (int)src.size(), tmpl, n_tmpl, tree, &G, pr->icp_max_iter, pr->icp_tf_eps, pr->icp_rel_mse,
                pr->icp_max_corr_dist, mode, io, nullptr, nullptr, 0);
            /* best = lowest fitness, ties -> lowest guess id */
            if (!have || io.fitness < cr.fitness) {
                have = true;
                cr.fitness = io.fitness; cr.converged = io.converged; cr.iterations = io.iters; cr.best_guess = g;
                cr.state = io.state; cr.corr_hash = io.corr_hash;
                std::memcpy(cr.T, &io.T, 64);
                cr.accepted = (io.converged && io.fitness < pr->icp_fitness_gate) ? 1 : 0;
            }
        }
    }
    return 0;
}

}  // namespace

/* ================================================================== C API */
extern "C" {

int orc_unproject(const uint16_t* depth, int w, int h, float fx, float fy, float cx, float cy, float depth_scale,
                  float* out) {
    P4* o = reinterpret_cast<P4*>(out);
    for (int v = 0; v < h; ++v) for (int u = 0; u < w; ++u) {
        const size_t i = (size_t)v * w + u;
        const float z = (float)depth[i] * depth_scale;
        o[i].x = z * (((float)u - cx) / fx);
        o[i].y = z * (((float)v - cy) / fy);
        o[i].z = z;
        o[i].w = 1.0f;
    }
    return w * h;
}

int orc_passthrough(const float* xyzw, int n, int field, double lo, double hi, float* out, int32_t* src_index) {
    const P4* p = reinterpret_cast<const P4*>(xyzw);
    P4* o = reinterpret_cast<P4*>(out);
    int k = 0;
    for (int i = 0; i < n; ++i) {
        const float v = field == 0 ? p[i].x : (field == 1 ? p[i].y : p[i].z);
        if (!std::isfinite(v)) continue;
        if ((double)v > hi || (double)v < lo) continue;
        if (!std::isfinite(p[i].x) || !std::isfinite(p[i].y) || !std::isfinite(p[i].z)) continue;
        o[k] = p[i];
        if (src_index) src_index[k] = i;
        ++k;
    }
    return k;
}

int orc_voxel_grid(const float* xyzw, int n, float leaf, int mode, float* vox_out, int32_t* key_per_point,
                   int32_t* voxel_key, int32_t* voxel_count, int32_t min_b[3], int32_t div_b[3], int* overflow) {
    std::vector<P4> out;
    std::vector<int32_t> vk, vc;
    const int V = voxel_grid(reinterpret_cast<const P4*>(xyzw), n, leaf, mode, out, key_per_point, &vk, &vc, min_b,
                             div_b, overflow);
    if (vox_out) std::memcpy(vox_out, out.data(), sizeof(P4) * out.size());
    if (voxel_key) std::memcpy(voxel_key, vk.data(), 4 * vk.size());
    if (voxel_count) std::memcpy(voxel_count, vc.data(), 4 * vc.size());
    return V;
}

int orc_sac_plane(const float* xyzw, int n, double thr, int max_iter, double prob, uint32_t seed, int refine,
                  int mode, const int32_t* triplets_in, int n_triplets_in, float coeff_out[4], int32_t* inliers_out,
                  int* n_inliers, float coeff_pre[4], int32_t* inliers_pre, int* n_inliers_pre, int* iters_run,
                  int32_t* triplets_used, int cap_triplets, int* n_draws, double* k_margin) {
    SacOut o;
    sac_plane(reinterpret_cast<const P4*>(xyzw), n, thr, max_iter, prob, seed, refine, mode, triplets_in,
              n_triplets_in, o);
    if (coeff_out) std::memcpy(coeff_out, o.coeff, 16);
    if (coeff_pre) std::memcpy(coeff_pre, o.coeff_pre, 16);
    if (inliers_out) std::memcpy(inliers_out, o.inliers.data(), 4 * o.inliers.size());
    if (inliers_pre) std::memcpy(inliers_pre, o.inliers_pre.data(), 4 * o.inliers_pre.size());
    if (n_inliers) *n_inliers = (int)o.inliers.size();
    if (n_inliers_pre) *n_inliers_pre = (int)o.inliers_pre.size();
    if (iters_run) *iters_run = o.iters;
    if (n_draws) *n_draws = o.draws;
    if (k_margin) *k_margin = o.k_margin;
    if (triplets_used) {
        const int m = std::min(cap_triplets, o.draws);
        std::memcpy(triplets_used, o.triplets.data(), 12 * (size_t)m);
    }
    return o.found ? 1 : 0;
}

int orc_sac_plane_model(const float* xyzw, int n, int model_type, const float axis[3], double eps_angle, double thr, int max_iter,
                        double prob, uint32_t seed, int refine, int mode, float coeff_out[4], int32_t* inliers_out, int* n_inliers,
                        int32_t* inliers_pre, int* n_inliers_pre, int* iters_run) {
    SacOut o;
    SacModel m;
    m.type = model_type; m.eps = eps_angle;
    for (int k = 0; k < 3; ++k) m.axis[k] = axis ? axis[k] : 0.0f;
    sac_plane(reinterpret_cast<const P4*>(xyzw), n, thr, max_iter, prob, seed, refine, mode, nullptr, 0, o, &m);
    if (coeff_out) std::memcpy(coeff_out, o.coeff, 16);
    if (inliers_out) std::memcpy(inliers_out, o.inliers.data(), 4 * o.inliers.size());
    if (inliers_pre) std::memcpy(inliers_pre, o.inliers_pre.data(), 4 * o.inliers_pre.size());
    if (n_inliers) *n_inliers = (int)o.inliers.size();
    if (n_inliers_pre) *n_inliers_pre = (int)o.inliers_pre.size();
    if (iters_run) *iters_run = o.iters;
    return o.found ? 1 : 0;
}

/* surface_normal_estimation.cpp:182-233 (callback) with getNormal (:105-165): three constrained-plane RANSACs on the
 * shrinking cloud, pcl::compute3DCentroid of each plane (sequential float sums, / (float)n), the count-descending
 * exchange sort of :207-221, the handedness fix of :223-226 and the pose matrix of :228-237. */
int orc_surface_normals(const float* xyzw, int n, const float axis[3], double eps_angle, double thr, int max_iter, double prob,
                        uint32_t seed, int mode, orc_surface_result* out) {
    std::vector<P4> cloud(reinterpret_cast<const P4*>(xyzw), reinterpret_cast<const P4*>(xyzw) + n);
    struct Pl { float c[4]; float mid[3]; int n; int found; };
    Pl pl[3];
    for (int i = 0; i < 3; ++i) {
        SacOut o;
        SacModel m;
        m.type = (i == 0) ? 1 : 2; m.eps = eps_angle;
        for (int k = 0; k < 3; ++k) m.axis[k] = axis[k];
        sac_plane(cloud.data(), (int)cloud.size(), thr, max_iter, prob, seed, 1, mode, nullptr, 0, o, &m);
        Pl& P = pl[i];
        P.found = o.found ? 1 : 0;
        for (int k = 0; k < 4; ++k) P.c[k] = o.found ? o.coeff[k] : 0.0f;
        P.n = (int)o.inliers.size();
        float sx = 0.f, sy = 0.f, sz = 0.f;
        for (int32_t id : o.inliers) { sx += cloud[id].x; sy += cloud[id].y; sz += cloud[id].z; }
        const float cn = (float)P.n;
        P.mid[0] = P.n ? sx / cn : 0.f; P.mid[1] = P.n ? sy / cn : 0.f; P.mid[2] = P.n ? sz / cn : 0.f;
        std::vector<char> in(cloud.size(), 0);
        for (int32_t id : o.inliers) in[id] = 1;
        std::vector<P4> rest;
        for (size_t k = 0; k < cloud.size(); ++k) if (!in[k]) rest.push_back(cloud[k]);
        out->n_in[i] = (int)cloud.size();
        cloud.swap(rest);
    }
    for (int i = 0; i < 3; ++i) {
        for (int k = 0; k < 4; ++k) out->coeff[i][k] = pl[i].c[k];
        for (int k = 0; k < 3; ++k) out->midpoint[i][k] = pl[i].mid[k];
        out->n_plane[i] = pl[i].n; out->found[i] = pl[i].found;
    }
    out->n_left = (int)cloud.size();
    orc_surface_pose(&out->coeff[0][0], &out->midpoint[0][0], out->n_plane, out->Rt, out->order);
    return pl[0].found && pl[1].found && pl[2].found;
}

/* surface_normal_estimation.cpp:207-237: order planes by point count (largest first) with the node's own exchange loop,
 * flip normal 2 if the triad is left-handed, project the first midpoint; Rt columns = (n2, n1, n0, centroid), row-major. */
void orc_surface_pose(const float coeff[12], const float midpoint[9], const int32_t n_plane[3], float Rt[16], int32_t order[3]) {
    float nr[3][3], mid[3][3];
    int cnt[3], ord[3] = {0, 1, 2};
    for (int i = 0; i < 3; ++i) { for (int k = 0; k < 3; ++k) { nr[i][k] = coeff[4 * i + k]; mid[i][k] = midpoint[3 * i + k]; } cnt[i] = n_plane[i]; }
    for (int i = 0; i < 3; ++i)
        for (int j = i; j < 3; ++j)
            if (cnt[i] < cnt[j]) {
                std::swap(cnt[i], cnt[j]); std::swap(ord[i], ord[j]);
                for (int k = 0; k < 3; ++k) { std::swap(nr[i][k], nr[j][k]); std::swap(mid[i][k], mid[j][k]); }
            }
    /* normals[2].dot(normals[1].cross(normals[0])) < 0 : Eigen Vector3f, cross then the size-3 redux a0 + (a1 + a2) */
    const float cr[3] = {nr[1][1] * nr[0][2] - nr[1][2] * nr[0][1], nr[1][2] * nr[0][0] - nr[1][0] * nr[0][2], nr[1][0] * nr[0][1] - nr[1][1] * nr[0][0]};
    const float triple = nr[2][0] * cr[0] + (nr[2][1] * cr[1] + nr[2][2] * cr[2]);
    if (triple < 0.0f) for (int k = 0; k < 3; ++k) nr[2][k] = -nr[2][k];
    const float df[3] = {mid[0][0] - mid[1][0], mid[0][1] - mid[1][1], mid[0][2] - mid[1][2]};
    const float proj = nr[0][0] * df[0] + (nr[0][1] * df[1] + nr[0][2] * df[2]);
    float cen[3];
    for (int k = 0; k < 3; ++k) cen[k] = mid[0][k] - proj * nr[0][k];
    for (int r = 0; r < 3; ++r) { Rt[4 * r + 0] = nr[2][r]; Rt[4 * r + 1] = nr[1][r]; Rt[4 * r + 2] = nr[0][r]; Rt[4 * r + 3] = cen[r]; }
    Rt[12] = 0.f; Rt[13] = 0.f; Rt[14] = 0.f; Rt[15] = 1.f;
    for (int i = 0; i < 3; ++i) order[i] = ord[i];
}

int orc_select_object(const int32_t* sizes, const int32_t* converged, const double* fitness, int n_clusters, int template_points,
                      double icp_fitness_score, int32_t* argmin_out, int32_t* reference_cluster, int32_t* attempts) {
    std::vector<int> icp_transforms;      /* which cluster each pushed transform belongs to */
    std::vector<double> diff_scores;
    for (int c = 0; c < n_clusters; ++c) {
        int max_iter = 0;
        while (true) {                     /* icp_registration: the same deterministic ICP again and again */
            icp_transforms.push_back(c);
            ++max_iter;
            if ((converged[c] && fitness[c] < icp_fitness_score) || max_iter > 10) break;
        }
        if (attempts) attempts[c] = max_iter;
        diff_scores.push_back((double)std::abs((int)sizes[c] - template_points));
    }
    double min_score = 1000;
    int argmin = -1;
    for (size_t i = 0; i < diff_scores.size(); ++i)
        if (diff_scores[i] < min_score) { argmin = (int)i; min_score = diff_scores[i]; }
    *argmin_out = argmin;
    *reference_cluster = (argmin >= 0 && argmin < (int)icp_transforms.size()) ? icp_transforms[argmin] : -1;
    return (argmin >= 0 && min_score < 250) ? 1 : 0;
}

// bbox_filter.cpp:30-51: u, v, w are accumulated in double (the matrix is vector<double>) and stored to float; the
// perspective divide and the comparisons against the int rectangle run in float; a point is kept only strictly inside.
int orc_bbox_filter(const float* xyzw, int n, const double P[12], const int32_t bbox[4], float* out_xyzw, int32_t* idx_out) {
    const P4* p = reinterpret_cast<const P4*>(xyzw);
    P4* o = reinterpret_cast<P4*>(out_xyzw);
    int k = 0;
    for (int i = 0; i < n; ++i) {
        const double x = (double)p[i].x, y = (double)p[i].y, z = (double)p[i].z;
        volatile float u = (float)((((P[0] * x) + (P[1] * y)) + (P[2] * z)) + P[3]);
        volatile float v = (float)((((P[4] * x) + (P[5] * y)) + (P[6] * z)) + P[7]);
        volatile float w = (float)((((P[8] * x) + (P[9] * y)) + (P[10] * z)) + P[11]);
        const float un = u / w, vn = v / w;
        const bool in = ((float)bbox[0] < un && un < (float)bbox[2]) && ((float)bbox[1] < vn && vn < (float)bbox[3]);
        if (in) { if (o) o[k] = p[i]; if (idx_out) idx_out[k] = i; ++k; }
    }
    return k;
}

int orc_extract(const float* xyzw, int n, const int32_t* idx, int n_idx, int negative, float* out, int32_t* src_index) {
    const P4* p = reinterpret_cast<const P4*>(xyzw);
    P4* o = reinterpret_cast<P4*>(out);
    int k = 0;
    if (negative) {
        std::vector<char> in(n, 0);
        for (int i = 0; i < n_idx; ++i) if (idx[i] >= 0 && idx[i] < n) in[idx[i]] = 1;
        for (int i = 0; i < n; ++i) if (!in[i]) { o[k] = p[i]; if (src_index) src_index[k] = i; ++k; }
    } else {
        for (int i = 0; i < n_idx; ++i) { o[k] = p[idx[i]]; if (src_index) src_index[k] = idx[i]; ++k; }
    }
    return k;
}

int orc_cluster(const float* xyzw, int n, double tol, int min_size, int max_size, int32_t* idx_sorted_out,
                int32_t* offsets_out) {
    std::vector<int32_t> idx, off;
    const int k = cluster(reinterpret_cast<const P4*>(xyzw), n, tol, min_size, max_size, idx, off);
    std::memcpy(idx_sorted_out, idx.data(), 4 * idx.size());
    std::memcpy(offsets_out, off.data(), 4 * off.size());
    return k;
}

int orc_icp(const float* src_xyzw, int n_src, const float* tgt_xyzw, int n_tgt, const float* guess16, int max_iter,
            double tf_eps, double rel_mse, double max_corr_dist, int mode, float T_out[16], double* fitness,
            int* converged, int* iters, int* state, float* aligned_xyzw, int32_t* corr_trace, float* T_trace,
            int cap_iters, uint64_t* corr_hash) {
    const P4* tgt = reinterpret_cast<const P4*>(tgt_xyzw);
    KdTree tree;
    tree.init(tgt, n_tgt);
    M4 G = m4_identity();
    if (guess16) std::memcpy(&G, guess16, 64);
    IcpOut o;
    icp(reinterpret_cast<const P4*>(src_xyzw), n_src, tgt, n_tgt, tree, &G, max_iter, tf_eps, rel_mse, max_corr_dist,
        mode, o, corr_trace, T_trace, cap_iters);
    if (T_out) std::memcpy(T_out, &o.T, 64);
    if (fitness) *fitness = o.fitness;
    if (converged) *converged = o.converged;
    if (iters) *iters = o.iters;
    if (state) *state = o.state;
    if (aligned_xyzw) std::memcpy(aligned_xyzw, o.aligned.data(), sizeof(P4) * o.aligned.size());
    if (corr_hash) *corr_hash = o.corr_hash;
    return 0;
}

/* Eigen general 4x4 inverse (cofactor expansion) in double, then tf::Matrix3x3::getRotation */
void orc_pose_from_transform(const float T[16], double H[16], double pose7[7]) {
    double m[16], inv[16];
    for (int i = 0; i < 16; ++i) m[i] = (double)T[i];
    inv[0] = m[5] * m[10] * m[15] - m[5] * m[11] * m[14] - m[9] * m[6] * m[15] + m[9] * m[7] * m[14] + m[13] * m[6] * m[11] - m[13] * m[7] * m[10];
    inv[4] = -m[4] * m[10] * m[15] + m[4] * m[11] * m[14] + m[8] * m[6] * m[15] - m[8] * m[7] * m[14] - m[12] * m[6] * m[11] + m[12] * m[7] * m[10];
    inv[8] = m[4] * m[9] * m[15] - m[4] * m[11] * m[13] - m[8] * m[5] * m[15] + m[8] * m[7] * m[13] + m[12] * m[5] * m[11] - m[12] * m[7] * m[9];
    inv[12] = -m[4] * m[9] * m[14] + m[4] * m[10] * m[13] + m[8] * m[5] * m[14] - m[8] * m[6] * m[13] - m[12] * m[5] * m[10] + m[12] * m[6] * m[9];
    inv[1] = -m[1] * m[10] * m[15] + m[1] * m[11] * m[14] + m[9] * m[2] * m[15] - m[9] * m[3] * m[14] - m[13] * m[2] * m[11] + m[13] * m[3] * m[10];
    inv[5] = m[0] * m[10] * m[15] - m[0] * m[11] * m[14] - m[8] * m[2] * m[15] + m[8] * m[3] * m[14] + m[12] * m[2] * m[11] - m[12] * m[3] * m[10];
    inv[9] = -m[0] * m[9] * m[15] + m[0] * m[11] * m[13] + m[8] * m[1] * m[15] - m[8] * m[3] * m[13] - m[12] * m[1] * m[11] + m[12] * m[3] * m[9];
    inv[13] = m[0] * m[9] * m[14] - m[0] * m[10] * m[13] - m[8] * m[1] * m[14] + m[8] * m[2] * m[13] + m[12] * m[1] * m[10] - m[12] * m[2] * m[9];
    inv[2] = m[1] * m[6] * m[15] - m[1] * m[7] * m[14] - m[5] * m[2] * m[15] + m[5] * m[3] * m[14] + m[13] * m[2] * m[7] - m[13] * m[3] * m[6];
    inv[6] = -m[0] * m[6] * m[15] + m[0] * m[7] * m[14] + m[4] * m[2] * m[15] - m[4] * m[3] * m[14] - m[12] * m[2] * m[7] + m[12] * m[3] * m[6];
    inv[10] = m[0] * m[5] * m[15] - m[0] * m[7] * m[13] - m[4] * m[1] * m[15] + m[4] * m[3] * m[13] + m[12] * m[1] * m[7] - m[12] * m[3] * m[5];
    inv[14] = -m[0] * m[5] * m[14] + m[0] * m[6] * m[13] + m[4] * m[1] * m[14] - m[4] * m[2] * m[13] - m[12] * m[1] * m[6] + m[12] * m[2] * m[5];
    inv[3] = -m[1] * m[6] * m[11] + m[1] * m[7] * m[10] + m[5] * m[2] * m[11] - m[5] * m[3] * m[10] - m[9] * m[2] * m[7] + m[9] * m[3] * m[6];
    inv[7] = m[0] * m[6] * m[11] - m[0] * m[7] * m[10] - m[4] * m[2] * m[11] + m[4] * m[3] * m[10] + m[8] * m[2] * m[7] - m[8] * m[3] * m[6];
    inv[11] = -m[0] * m[5] * m[11] + m[0] * m[7] * m[9] + m[4] * m[1] * m[11] - m[4] * m[3] * m[9] - m[8] * m[1] * m[7] + m[8] * m[3] * m[5];
    inv[15] = m[0] * m[5] * m[10] - m[0] * m[6] * m[9] - m[4] * m[1] * m[10] + m[4] * m[2] * m[9] + m[8] * m[1] * m[6] - m[8] * m[2] * m[5];
    const double det = m[0] * inv[0] + m[1] * inv[4] + m[2] * inv[8] + m[3] * inv[12];
    for (int i = 0; i < 16; ++i) H[i] = inv[i] / det;
    /* tf::Matrix3x3::getRotation */
    const double xx = H[0], yy = H[5], zz = H[10];
    const double trace = xx + yy + zz;
    double q[4];
    auto R = [&](int r, int c) { return H[4 * r + c]; };
    if (trace > 0.0) {
        double s = std::sqrt(trace + 1.0);
        q[3] = s * 0.5; s = 0.5 / s;
        q[0] = (R(2, 1) - R(1, 2)) * s; q[1] = (R(0, 2) - R(2, 0)) * s; q[2] = (R(1, 0) - R(0, 1)) * s;
    } else {
        const int i = xx < yy ? (yy < zz ? 2 : 1) : (xx < zz ? 2 : 0);
        const int j = (i + 1) % 3, k = (i + 2) % 3;
        double s = std::sqrt(R(i, i) - R(j, j) - R(k, k) + 1.0);
        q[i] = s * 0.5; s = 0.5 / s;
        q[3] = (R(k, j) - R(j, k)) * s; q[j] = (R(j, i) + R(i, j)) * s; q[k] = (R(k, i) + R(i, k)) * s;
    }
    pose7[0] = H[3]; pose7[1] = H[7]; pose7[2] = H[11];
    pose7[3] = q[0]; pose7[4] = q[1]; pose7[5] = q[2]; pose7[6] = q[3];
}

/* publish_bounding_box (icp.cpp:94-128): 8 corners through H.cast<float>() with pcl::transformPointCloud */
void orc_bbox_corners(const double H[16], double l, double w, double h, float out[32]) {
    M4 Hf;
    for (int i = 0; i < 16; ++i) Hf.a[i / 4][i % 4] = (float)H[i];
    int k = 0;
    for (int sx = -1; sx <= 1; sx += 2) for (int sy = -1; sy <= 1; sy += 2) for (int sz = -1; sz <= 1; sz += 2) {
        const P4 c{(float)(sx * l / 2), (float)(sy * w / 2), (float)(sz * h / 2), 1.0f};
        const P4 t = m4_apply(Hf, c);
        out[4 * k] = t.x; out[4 * k + 1] = t.y; out[4 * k + 2] = t.z; out[4 * k + 3] = 1.0f;
        ++k;
    }
}

void orc_guess_about_centroid(const float* src_xyzw, int n, const float R9[9], float G16[16]) {
    M4 G;
    guess_about_centroid(reinterpret_cast<const P4*>(src_xyzw), n, R9, ORC_CANONICAL, G);
    std::memcpy(G16, &G, 64);
}

int orc_process_frame(const orc_params* pr, const uint16_t* depth, int w, int h, const float* tmpl, int n_tmpl,
                      const float* guesses, int mode, orc_frame_result* out) {
    std::memset(out, 0, sizeof(*out));
    std::vector<P4> all((size_t)w * h);
    orc_unproject(depth, w, h, pr->fx, pr->fy, pr->cx, pr->cy, pr->depth_scale, reinterpret_cast<float*>(all.data()));
    std::vector<P4> pz(all.size()), px(all.size());
    const int nz = orc_passthrough(reinterpret_cast<float*>(all.data()), (int)all.size(), 2, pr->pass_z_min, pr->pass_z_max,
                                   reinterpret_cast<float*>(pz.data()), nullptr);
    const int nx = orc_passthrough(reinterpret_cast<float*>(pz.data()), nz, 0, pr->pass_x_min, pr->pass_x_max,
                                   reinterpret_cast<float*>(px.data()), nullptr);
    px.resize(nx);
    return process_points(pr, px, reinterpret_cast<const P4*>(tmpl), n_tmpl, guesses, mode, out);
}

int orc_process_cloud(const orc_params* pr, const void* blob, int point_step, int xoff, int yoff, int zoff, int n,
                      const float* tmpl, int n_tmpl, const float* guesses, int mode, orc_frame_result* out) {
    std::memset(out, 0, sizeof(*out));
    std::vector<P4> all(n);
    const uint8_t* b = static_cast<const uint8_t*>(blob);
    for (int i = 0; i < n; ++i) {
        std::memcpy(&all[i].x, b + (size_t)i * point_step + xoff, 4);
        std::memcpy(&all[i].y, b + (size_t)i * point_step + yoff, 4);
        std::memcpy(&all[i].z, b + (size_t)i * point_step + zoff, 4);
        all[i].w = 1.0f;
    }
    std::vector<P4> pz(all.size()), px(all.size());
    const int nz = orc_passthrough(reinterpret_cast<float*>(all.data()), n, 2, pr->pass_z_min, pr->pass_z_max,
                                   reinterpret_cast<float*>(pz.data()), nullptr);
    const int nx = orc_passthrough(reinterpret_cast<float*>(pz.data()), nz, 0, pr->pass_x_min, pr->pass_x_max,
                                   reinterpret_cast<float*>(px.data()), nullptr);
    px.resize(nx);
    return process_points(pr, px, reinterpret_cast<const P4*>(tmpl), n_tmpl, guesses, mode, out);
}

/* pcl::io::loadPCDFile for the ASCII v0.7 files make_cuboid.py writes (make_cuboid.py:24-35,66) */
int orc_load_pcd(const char* path, float* xyzw_out, int cap) {
    FILE* f = std::fopen(path, "r");
    if (!f) return -1;
    char line[512];
    int points = -1;
    bool data = false;
    while (std::fgets(line, sizeof line, f)) {
        if (std::strncmp(line, "POINTS", 6) == 0) points = std::atoi(line + 6);
        if (std::strncmp(line, "DATA", 4) == 0) { data = std::strstr(line, "ascii") != nullptr; break; }
    }
    if (points < 0 || !data) { std::fclose(f); return -2; }
    int k = 0;
    while (k < points && std::fgets(line, sizeof line, f)) {
        char* e = line;
        const float x = std::strtof(e, &e), y = std::strtof(e, &e), z = std::strtof(e, &e);
        if (xyzw_out && k < cap) { xyzw_out[4 * k] = x; xyzw_out[4 * k + 1] = y; xyzw_out[4 * k + 2] = z; xyzw_out[4 * k + 3] = 1.0f; }
        ++k;
    }
    std::fclose(f);
    return k;
}

uint32_t orc_mt19937_nth(uint32_t seed, int n) {
    MT19937 r(seed);
    uint32_t v = 0;
    for (int i = 0; i < n; ++i) v = r.next();
    return v;
}
uint64_t orc_hash_i32(const int32_t* v, int n) { return hash_i32(v, n); }
uint64_t orc_hash_f32x3(const float* xyzw, int n) { return hash_pts(reinterpret_cast<const P4*>(xyzw), n); }
int orc_params_size(void) { return (int)sizeof(orc_params); }
int orc_frame_result_size(void) { return (int)sizeof(orc_frame_result); }

}  // extern "C"
