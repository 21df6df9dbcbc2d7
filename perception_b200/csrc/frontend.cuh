// frontend.cuh — stages 1a + 1b of one frame in ONE kernel: depth -> cloud unprojection (or PointCloud2 blob read),
// both PassThrough filters, ordered compaction, getMinMax3D, the VoxelGrid key, the stable radix sort and the
// sequential float centroids (gps.cpp:53-73, opd.cpp:275-298).
//
// One thread-block CLUSTER owns one frame at a time (persistent grid, clusters stride over the frames of the chunk); a throughput
// launch uses clusters of ONE CTA, two CTAs per SM (the SOLO instance below), a launch with a handful of frames spreads each frame
// over an 8-CTA cluster. The CTAs of a cluster split the frame's pixels, then its sort records, then its sorted records into
// contiguous ranges; what they must agree on (survivor counts, min/max, digit histograms, voxel counts) is exchanged through
// distributed shared memory between cluster barriers. The radix ping-pong buffers belong to the cluster SLOT, not to the frame, so
// the same few megabytes are rewritten frame after frame. With 296 frames in flight they do not stay in the 126 MB L2: measured
// DRAM traffic is 2.7x the algorithmic bytes (profiles/README.md), and the kernel is bound by instruction issue, not by that traffic
// (two variants that move fewer bytes - the voxel-hash path and run records, both below, both opt-in - are slower).
//
// Results are bit-identical to the unfused kernels in preprocess.cuh / voxel.cuh (kept as the CUBOID_OPT_FRONTEND=0
// path and compared byte for byte in tests/test_gpu_parity.py): same point order, same keys, same stable order inside
// a voxel, same sequential float sums.
//
// Roofline: HBM. Algorithmic bytes per frame = 2*P + 16*N (a0+a1) + 16*N + 16*V (a2).
#pragma once
#include <cooperative_groups.h>
#include <type_traits>

#include "common.cuh"
#include "preprocess.cuh"
#include "voxel.cuh"

namespace cuboid {
namespace cg = cooperative_groups;

struct FrontArgs {
    PreArgs pre;                 // input description, limits, pts, res, scr (tile_count / tiles are unused here)
    unsigned long long* keys;    // [n_slots][keys_stride] scratch of the frame a cluster slot is working on: the radix ping-pong [2][P], or the
                                 // voxel-hash path's records / offsets / slot ids / grouped point indices (FehScratch)
    size_t keys_stride;          // u64 per slot
    // One-pass mode (depth input, one CTA per frame, no parity taps): the sort key of a point is its (k, j, i) voxel coordinates relative to
    // STATIC lower bounds the host derives from the pass-through limits and the intrinsics, k in the top bits. Its order is the order
    // of PCL's idx = i + j*dx + k*dx*dy for any actual min / max, so the first pass over the inputs (survivor count, min / max: A1)
    // is not needed before the keys can be written: min / max are reduced during the one pass and the geometry is published after it.
    int st_on, st_min[3], st_b0, st_b1, st_bits;
    // Run records (one-pass mode only): consecutive survivors of one warp step that fall into the same voxel become ONE sort record
    // (key << 32 | first point << 5 | length - 1). A raster run stays contiguous in the stable order, so the radix passes move ~2x fewer
    // records and the centroid pass expands them again in order: same sums, bit for bit.
    int runs;
    int hash;                    // 1: try the voxel-hash path first (needs NT = 1024, cluster size 1), fall back to the radix path per frame
    int* kpp;                    // [F][P] voxel idx per point (parity tap) or NULL
    float4* vox;                 // [F][P]
    int* vcount;                 // [F][P] points per voxel (parity tap) or NULL
    float inv_leaf;
    int rgb;                     // 1: .w of every point is a packed rgb(a) word: VoxelGrid<PCLPointCloud2> averages r, g, b per voxel (gps.cpp:69-73)
    int hashes;                  // 1: accumulate points_hash / voxel_key_hash / voxel_hash (parity taps); 0: leave them 0
    int P;                       // per-frame stride of pts / vox / kpp / vcount / keys
    int n_frames;
    int arena;                   // dynamic shared memory of this launch (bytes)
};

constexpr int FE_ITEMS = 8;      // inputs / sort records per thread and tile
constexpr int FE_RITEMS = 4;     // sorted records per thread and reduce tile
constexpr int FE_LONGRUN = 96;   // voxels with more points than this are summed by a whole warp (lane-parallel loads)
constexpr int FE_MAXDEF = 64;    // >= 512*4/96 + 1 and >= 1024*4/96 + 1
#ifndef CUBOID_FE_RREC
#define CUBOID_FE_RREC 4
#endif
#ifndef CUBOID_FE_RCAPT
#define CUBOID_FE_RCAPT 8
#endif
constexpr int FE_RREC = CUBOID_FE_RREC;     // run-record path: sorted records staged per thread and reduce tile
constexpr int FE_RCAPT = CUBOID_FE_RCAPT;   // ... and the points they may expand to, per thread
constexpr int FE_LOOK = 128;     // sorted records staged past a reduce tile so that the tile's last voxel can finish in shared memory

struct FeXchg {          // what a CTA publishes to its cluster peers
    int count;           // survivors of its input slice
    float mn[3], mx[3];  // their min / max
    int heads;           // voxels that start inside its range of the sorted records
};
struct FeFrame {         // cluster-wide facts of the current frame, replicated per CTA by thread 0
    int base, N;         // first output position of this CTA's survivors, survivors of the frame
    int vbase, V;
    VoxelGeom g;
};

// dynamic shared memory: the three phases reuse one arena
//   A1/A2   u16 tile-local input index of every survivor of the tile            NT*8*2  bytes
//           + the tile's depth values (u16), + the keep masks A1 found, so A2       NT*8*2 + FE_MASK_BYTES
//             does not unproject and filter again
//   B       per-warp digit counters [NT/32][256] u32 + one prefetched key tile    NT*32 + NT*8*8 bytes
//   C2      key,x,y,z,w of the staged sorted records (+ look-ahead) + u16 heads (NT*4+128)*20 + NT*4*2 bytes
constexpr int FE_MASK_BYTES = 40960;   // keep masks of one input slice (1 byte per 8 inputs): a VGA frame on one CTA needs 38 400
constexpr int fe_max(int a, int b) { return a > b ? a : b; }
// ---- voxel-hash path (fe_hash_frame, opt-in with CUBOID_FE_HASH=1): the distinct voxels of a frame are found with a hash table in
// shared memory, only THEY are sorted, and the points are grouped by one counting scatter. Same outputs as the radix path, byte for
// byte (tests), and 1.6x instead of 2.8x the algorithmic DRAM traffic - but on B200 it executes as many instructions as the radix
// path (hashing and the unproject / filter work done twice replace the saved radix passes) with more barrier stalls: 7.1 ms against
// 4.75 ms per 1024 VGA frames (profiles/README.md). Kept as the measured alternative, not the default.
constexpr int FEH_CAP = 32768;        // table slots (keys u32 + packed u16 counts = 192 KB)
constexpr int FEH_PROBES = 64;        // an insert that does not find its slot within this many probes sends the frame to the radix path
constexpr int FEH_RCAP = 4096;        // points staged per reduce chunk
constexpr int FEH_LONG = 64;          // runs longer than this are sorted and summed by a whole warp
constexpr int FEH_MAXRUN = 2048;      // longest run (points of one voxel) the path handles; longer: radix path
constexpr int FEH_SMEM = FEH_CAP * 4 + FEH_CAP * 2 + 1024 * FE_ITEMS * 2;   // table + counts + selection tile = 212 992 B
// What a launch needs depends on its mode (the carve-out left to L1 matters: 4.7 -> 5.2 ms per 1024 VGA frames when two 97 KB CTAs
// pushed it from 196 to 228 KB): keep masks only when A1 runs, the hash table only on the hash path, run staging only with run records.
template <int NT>
constexpr int fe_dyn_smem(bool masks = true, bool hash = true, bool runs = true) {
    // A: selection + depth tile (+ masks | + run staging); B: digit counters + one prefetched key tile; C2: staging + heads
    return fe_max(fe_max(fe_max(NT * 32 + (masks ? FE_MASK_BYTES : (runs ? NT * FE_ITEMS * 8 + 2 * NT : 0)), NT * 32 + NT * FE_ITEMS * 8),
                         runs && !masks ? NT * FE_RREC * 14 + 4 + NT * FE_RCAPT * 12 : (NT * FE_RITEMS + FE_LOOK) * 20 + NT * FE_RITEMS * 2),
                  (hash && NT == 1024) ? FEH_SMEM : 0);
}
// global scratch of one slot on the hash path (inside FrontArgs::keys): two record buffers, voxel start offsets, slot id per
// point, point indices grouped by voxel
struct FehScratch {
    unsigned long long* recA; unsigned long long* recB; unsigned int* vstart; unsigned short* slotid; unsigned int* sidx;
};
__host__ __device__ inline size_t feh_scratch_u64(int P) {   // u64 needed per slot
    return (size_t)2 * FEH_CAP + (size_t)(FEH_CAP + 8) / 2 + ((size_t)P * 2 + 7) / 8 + ((size_t)P * 4 + 7) / 8 + 8;
}
__device__ __forceinline__ FehScratch feh_scratch(unsigned long long* base, int P) {
    FehScratch s;
    s.recA = base; s.recB = base + FEH_CAP;
    s.vstart = reinterpret_cast<unsigned int*>(base + 2 * FEH_CAP);
    unsigned long long* q = base + 2 * FEH_CAP + (FEH_CAP + 8) / 2;
    s.slotid = reinterpret_cast<unsigned short*>(q);
    s.sidx = reinterpret_cast<unsigned int*>(q + ((size_t)P * 2 + 7) / 8);
    return s;
}
// asynchronous global -> shared copies (LDGSTS): the next tile of sort records is on its way while this one is ranked
__device__ __forceinline__ void fe_cp_async16(void* smem, const void* gmem) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((unsigned int)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void fe_cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void fe_cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

template <int NT, int NQ>
__device__ __forceinline__ void fe_block_sum_u64(unsigned long long (&v)[NQ], unsigned long long* s_part /*[NQ][NT/32]*/) {
    constexpr int NW = NT / 32;
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
        const unsigned long long x = warp_sum_u64(v[q]);
        if (lane == 0) s_part[q * NW + wid] = x;
    }
    __syncthreads();
    if (threadIdx.x < NQ) {
        unsigned long long s = 0;
        for (int w = 0; w < NW; ++w) s += s_part[threadIdx.x * NW + w];
        s_part[threadIdx.x * NW] = s;
    }
    __syncthreads();
#pragma unroll
    for (int q = 0; q < NQ; ++q) v[q] = s_part[q * NW];
    __syncthreads();
}

// one input -> point, with exactly the float operations of pre_points (preprocess.cuh)
template <int SRC>
__device__ __forceinline__ float4 fe_point_at(const PreArgs& a, int f, int i) {
    if (SRC == 0) {
        const int v = pre_row(a, i), u = i - v * a.w;
        const float z = (float)a.depth[(size_t)f * a.P + i] * a.depth_scale;
        return make_float4(z * a.xr[u], z * a.yr[v], z, 1.0f);
    } else {
        const unsigned char* rec = a.blob + ((size_t)f * a.P + i) * a.point_step;
        // .w = pcl::PointXYZ's padding (1.0f), or the record's packed rgb(a) word when the cloud carries one: it then travels with
        // the point through every float4 copy of the pipeline (nothing downstream reads .w as a coordinate)
        return make_float4(*reinterpret_cast<const float*>(rec + a.xoff), *reinterpret_cast<const float*>(rec + a.yoff),
                           *reinterpret_cast<const float*>(rec + a.zoff), a.rgboff >= 0 ? *reinterpret_cast<const float*>(rec + a.rgboff) : 1.0f);
    }
}

// the same from a depth value already staged in shared memory (SRC 0 only)
__device__ __forceinline__ float4 fe_point_from_depth(const PreArgs& a, int i, unsigned short dv) {
    const int v = pre_row(a, i), u = i - v * a.w;
    const float z = (float)dv * a.depth_scale;
    return make_float4(z * a.xr[u], z * a.yr[v], z, 1.0f);
}

// s += (a run of +-0 values), folded: the sum only changes when it is -0 and a +0 is added (IEEE round-to-nearest)
__device__ __forceinline__ float fe_add_zeros(float s, bool any_plus_zero) {
    return (__float_as_uint(s) == 0x80000000u && any_plus_zero) ? 0.0f : s;
}

// lanes of the warp holding the same 8-bit digit (256 = "no record": never equal to a real digit's peers that matter).
// Eight votes: MATCH.ANY costs ~2 cycles per DISTINCT value per SM (measured, tools/microbench/match_bench.cu), which
// is several times more than this for the scattered digits of the later radix passes.
template <bool ALL_VALID = false>   // ALL_VALID: every lane holds a record (a full tile): no vote on that
__device__ __forceinline__ unsigned int fe_digit_peers(unsigned int d) {
    unsigned int peers = FULL_MASK;
    if (!ALL_VALID) {
        peers = __ballot_sync(FULL_MASK, d < 256u);
        if (d >= 256u) peers = ~peers;
    }
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const unsigned int bit = (d >> b) & 1u;
        const unsigned int bal = __ballot_sync(FULL_MASK, bit != 0u);
        peers &= bit ? bal : ~bal;
    }
    return peers;
}

// One 32-record step of a voxel's run, called by a whole warp: lane j holds record j (match = it belongs to the voxel,
// a = its point). Adds the leading matching records to the running sums IN ORDER and returns how many there were
// (32 = the run may continue). An axis whose values are all +-0 folds without the 32-deep dependent chain.
struct FeRun { float sx, sy, sz; int cnt; unsigned int sr, sg, sb; };
// VoxelGrid<PCLPointCloud2>'s rgb special case (PCL voxel_grid.cpp, [PCL-recall]): r, g, b of pcl::RGB {b, g, r, a} are summed as
// floats (exact integers below 2^24, so the order does not matter), divided by the float count like every centroid field, truncated to
// int and repacked as (r << 16) | (g << 8) | b.
__device__ __forceinline__ float fe_rgb_avg(unsigned int sr, unsigned int sg, unsigned int sb, int cnt) {
    const float c = (float)cnt;
    const int r = (int)((float)sr / c), g = (int)((float)sg / c), b = (int)((float)sb / c);
    return __int_as_float((r << 16) | (g << 8) | b);
}
template <bool RGB>
__device__ __forceinline__ int fe_run_step(FeRun& r, bool match, float ax, float ay, float az, unsigned int aw = 0u) {
    const unsigned int miss = __ballot_sync(FULL_MASK, !match);
    const int nmatch = miss ? (__ffs(miss) - 1) : 32;
    const unsigned int in = nmatch == 32 ? FULL_MASK : ((1u << nmatch) - 1u);
    const unsigned int nzx = __ballot_sync(FULL_MASK, ax != 0.0f) & in, nzy = __ballot_sync(FULL_MASK, ay != 0.0f) & in,
                       nzz = __ballot_sync(FULL_MASK, az != 0.0f) & in;
    const unsigned int pzx = __ballot_sync(FULL_MASK, __float_as_uint(ax) == 0u) & in, pzy = __ballot_sync(FULL_MASK, __float_as_uint(ay) == 0u) & in,
                       pzz = __ballot_sync(FULL_MASK, __float_as_uint(az) == 0u) & in;
    if (nzx) { for (int j = 0; j < nmatch; ++j) r.sx += __shfl_sync(FULL_MASK, ax, j); } else r.sx = fe_add_zeros(r.sx, pzx != 0u);
    if (nzy) { for (int j = 0; j < nmatch; ++j) r.sy += __shfl_sync(FULL_MASK, ay, j); } else r.sy = fe_add_zeros(r.sy, pzy != 0u);
    if (nzz) { for (int j = 0; j < nmatch; ++j) r.sz += __shfl_sync(FULL_MASK, az, j); } else r.sz = fe_add_zeros(r.sz, pzz != 0u);
    if (RGB) {
        const bool mine = (int)(threadIdx.x & 31) < nmatch;
        r.sr += __reduce_add_sync(FULL_MASK, mine ? ((aw >> 16) & 255u) : 0u);
        r.sg += __reduce_add_sync(FULL_MASK, mine ? ((aw >> 8) & 255u) : 0u);
        r.sb += __reduce_add_sync(FULL_MASK, mine ? (aw & 255u) : 0u);
    }
    r.cnt += nmatch;
    return nmatch;
}

// static shared memory of k_frontend the hash path borrows
struct FehStatic {
    int* s_w;                       // [NT/32 + 1] block scan scratch
    unsigned int (*s_histA)[256];   // [4][256] digit histograms of the record sort
    unsigned int* s_base;           // [256]
    unsigned long long* s_h64;      // [2 * NT/32]
    int* s_ndef; int* s_def_lp; int* s_def_pos;   // deferred (long) runs of a reduce chunk
    int* s_misc;                    // [8]: distinct voxels, abort flag, max count, chunk end
};

__device__ __forceinline__ int fe_bitlen(int v) { return v <= 0 ? 0 : 32 - __clz(v); }

// Voxel-hash path for ONE frame (NT = 1024 threads, cluster size 1). On entry A1 has run: s_f.N survivors, geometry g.
// Returns false (nothing usable written except pts / kpp, which the radix path rewrites identically) when the frame does not fit
// the path: too many distinct voxels, a voxel with more than FEH_MAXRUN points, or more than 30 key bits.
//   A2h  one pass over the inputs: unproject / filter (recomputed: no room for A1's keep masks), ordered compaction, point store,
//        then per survivor the voxel's (i, j, k) packed into a key whose order is the order of PCL's idx = i + j*dx + k*dx*dy
//        (k in the top bits), find-or-insert in the table, count += 1, slot id -> global
//   Bh   occupied slots -> (key, slot, count) records; stable LSD radix sort of the V records (not of the N points)
//   Ch   prefix sum of the counts in key order: vstart[r]; cursor[slot] = vstart[r]
//   Dh   every point index to cursor[slot]++: the points of a voxel become one contiguous run (in no particular order)
//   Eh   chunks of whole runs staged in shared memory: each run is sorted ascending (= the stable order of the radix path), the
//        points are gathered, ONE THREAD PER VOXEL sums its run sequentially (long runs: a whole warp), centroids out
template <int SRC, int NT, bool RGB>
__device__ __forceinline__ bool fe_hash_frame(const FrontArgs& a, int f, int n_in, int N, const VoxelGeom& g, unsigned long long* slot_scratch,
                                              unsigned char* fe_dyn, const FehStatic& st) {
    static_assert(NT == 1024, "the hash path is laid out for one 1024-thread CTA per SM");
    constexpr int NW = NT / 32;
    const PreArgs& p = a.pre;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int b0 = fe_bitlen(g.div_b[0] - 1), b1 = fe_bitlen(g.div_b[1] - 1), b2 = fe_bitlen(g.div_b[2] - 1);
    const int kbits = b0 + b1 + b2;
    auto why = [&](int code) { if (a.hash == 2 && threadIdx.x == 0) atomicOr(&a.pre.res[f].status, code << 8); };   // developer: CUBOID_FE_HASH=2
    if (kbits > 30 || g.overflow_mode) { why(1); return false; }
    unsigned int* s_tab = reinterpret_cast<unsigned int*>(fe_dyn);
    unsigned int* s_cnt = s_tab + FEH_CAP;                       // two u16 counts per word
    unsigned short* s_sel = reinterpret_cast<unsigned short*>(s_cnt + FEH_CAP / 2);
    const FehScratch gs = feh_scratch(slot_scratch, a.P);
    int* s_misc = st.s_misc;
    for (int i = tid; i < FEH_CAP; i += NT) s_tab[i] = 0xffffffffu;
    for (int i = tid; i < FEH_CAP / 2; i += NT) s_cnt[i] = 0u;
    if (tid < 8) s_misc[tid] = 0;
    __syncthreads();

    // ---- A2h ----
    {
        constexpr int HT = NT * FE_ITEMS;   // inputs per tile
        float4* out = p.pts + (size_t)f * p.Pout;
        int* kpp = a.kpp ? a.kpp + (size_t)f * a.P : nullptr;
        int run = 0;
        unsigned long long hh[2] = {0ull, 0ull};
        for (int t0 = 0; t0 < n_in; t0 += HT) {
            const int first = t0 + tid * FE_ITEMS;
            unsigned int keep = 0;
            if (first < n_in) {
                float px[FE_ITEMS], py[FE_ITEMS], pz[FE_ITEMS];
                keep = pre_points<SRC>(p, f, first, n_in, px, py, pz);
            }
            int total;
            int lp = block_excl_scan<NT>(__popc(keep), st.s_w, &total);
#pragma unroll
            for (int k = 0; k < FE_ITEMS; ++k)
                if (keep & (1u << k)) s_sel[lp++] = (unsigned short)(tid * FE_ITEMS + k);
            __syncthreads();
            for (int q = tid; q < total; q += NT) {
                const int sel = (int)s_sel[q];
                // the depth value was read a moment ago by pre_points: it comes from L1 / L2
                const float4 pt = SRC == 0 ? fe_point_from_depth(p, t0 + sel, __ldg(p.depth + (size_t)f * p.P + t0 + sel)) : fe_point_at<SRC>(p, f, t0 + sel);
                const int pos = run + q;
                out[pos] = pt;
                // the three terms of voxel_index (common.cuh), kept apart
                const int i0 = (int)(floorf(pt.x * g.inv) - (float)g.min_b[0]);
                const int i1 = (int)(floorf(pt.y * g.inv) - (float)g.min_b[1]);
                const int i2 = (int)(floorf(pt.z * g.inv) - (float)g.min_b[2]);
                const unsigned int key = (unsigned int)i0 | ((unsigned int)i1 << b0) | ((unsigned int)i2 << (b0 + b1));
                unsigned int h = (key * 2654435761u) >> 17;     // 15 bits
                bool placed = false;
                const bool aborted = *reinterpret_cast<volatile int*>(&s_misc[1]) != 0;   // the frame is going to the radix path anyway: skip the probing
                for (int probe = 0; probe < FEH_PROBES && !aborted; ++probe) {
                    const unsigned int cur = s_tab[h];
                    if (cur == key) { placed = true; break; }
                    if (cur == 0xffffffffu) {
                        const unsigned int old = atomicCAS(&s_tab[h], 0xffffffffu, key);
                        if (old == 0xffffffffu || old == key) { placed = true; break; }
                    }
                    h = (h + 1) & (FEH_CAP - 1);
                }
                if (!placed) { s_misc[1] = 1; h = 0; }       // table (nearly) full: more distinct voxels than the path is built for
                const unsigned int sh16 = (h & 1u) * 16u;
                const unsigned int oldc = atomicAdd(&s_cnt[h >> 1], 1u << sh16);
                if (((oldc >> sh16) & 0xffffu) == 0xffffu) s_misc[1] = 1;   // the packed count overflowed into its neighbour
                gs.slotid[pos] = (unsigned short)h;
                if (kpp || a.hashes) {
                    const int idx = (int)((unsigned int)i0 + (unsigned int)i1 * g.mul1 + (unsigned int)i2 * g.mul2);
                    if (kpp) kpp[pos] = idx;
                    if (a.hashes) {
                        hh[0] += hash_point((unsigned int)pos, pt.x, pt.y, pt.z);
                        hh[1] += hash_index((unsigned int)pos, idx);
                    }
                }
            }
            run += total;
            __syncthreads();
        }
        if (s_misc[1]) { why(2); return false; }    // block-uniform: nobody writes the flag after the last barrier of the loop
        // the parity hashes are committed only when the path completes (the radix path would add them again otherwise)
        fe_block_sum_u64<NT, 2>(hh, st.s_h64);
        __syncthreads();
        if (tid == 0) { st.s_h64[0] = hh[0]; st.s_h64[1] = hh[1]; }
        __syncthreads();
    }
    const int npass = (kbits + 7) / 8;

    // ---- Bh: records of the occupied slots, digit histograms of every pass, longest run ----
    {
        for (int k = tid; k < 4 * 256; k += NT) (&st.s_histA[0][0])[k] = 0u;
        __syncthreads();
        constexpr int PER = FEH_CAP / NT;   // 32 consecutive slots per thread
        int occ = 0, mx = 0;
#pragma unroll 4
        for (int j = 0; j < PER; ++j) occ += (s_tab[tid * PER + j] != 0xffffffffu) ? 1 : 0;
        int total;
        int wp = block_excl_scan<NT>(occ, st.s_w, &total);
        for (int j = 0; j < PER; ++j) {
            const int sl = tid * PER + j;
            const unsigned int key = s_tab[sl];
            if (key == 0xffffffffu) continue;
            const unsigned int c = (s_cnt[sl >> 1] >> ((sl & 1) * 16)) & 0xffffu;
            mx = max(mx, (int)c);
            __stcg(gs.recA + wp, ((unsigned long long)key << 32) | ((unsigned long long)sl << 16) | (unsigned long long)c);
            ++wp;
            for (int ps = 0; ps < npass; ++ps) atomicAdd(&st.s_histA[ps][(key >> (8 * ps)) & 255u], 1u);
        }
        atomicMax(&s_misc[2], mx);
        if (tid == 0) s_misc[0] = total;
        __syncthreads();
        if (s_misc[2] > FEH_MAXRUN) { why(4); return false; }
    }
    const int V = s_misc[0];
    // stable LSD radix sort of the V records by key (the radix path's tile code, one CTA, global ping-pong recA <-> recB)
    {
        constexpr int TILE = NT * FE_ITEMS;
        unsigned int (*s_cntw)[256] = reinterpret_cast<unsigned int (*)[256]>(fe_dyn);   // the table is dead from here on
        for (int pass = 0; pass < npass; ++pass) {
            const unsigned long long* src = (pass & 1) ? gs.recB : gs.recA;
            unsigned long long* dst = (pass & 1) ? gs.recA : gs.recB;
            const int shift = 32 + pass * 8;
            {
                const unsigned int tot = tid < 256 ? st.s_histA[pass][tid] : 0u;
                int total;
                const int ex = block_excl_scan<NT>((int)tot, st.s_w, &total);
                if (tid < 256) st.s_base[tid] = (unsigned int)ex;
                __syncthreads();
            }
            for (int t0 = 0; t0 < V; t0 += TILE) {
                for (int d = lane; d < 256; d += 32) s_cntw[wid][d] = 0;
                __syncwarp();
                unsigned long long key[FE_ITEMS];
                unsigned int rank[FE_ITEMS];
#pragma unroll
                for (int k = 0; k < FE_ITEMS; ++k) {
                    const int i = t0 + wid * (32 * FE_ITEMS) + k * 32 + lane;
                    key[k] = i < V ? __ldcg(src + i) : ~0ull;
                }
#pragma unroll
                for (int k = 0; k < FE_ITEMS; ++k) {
                    const int i = t0 + wid * (32 * FE_ITEMS) + k * 32 + lane;
                    const bool valid = i < V;
                    const unsigned int d = valid ? ((unsigned int)(key[k] >> shift) & 255u) : 256u;
                    const unsigned int peers = fe_digit_peers(d);
                    const int leader = __ffs(peers) - 1;
                    unsigned int before = 0;
                    if (valid && lane == leader) { before = s_cntw[wid][d]; s_cntw[wid][d] = before + __popc(peers); }
                    before = __shfl_sync(FULL_MASK, before, leader);
                    rank[k] = before + __popc(peers & ((1u << lane) - 1u));
                    __syncwarp();
                }
                __syncthreads();
                if (tid < 256) {
                    unsigned int runb = st.s_base[tid];
#pragma unroll 8
                    for (int ww = 0; ww < NW; ++ww) { const unsigned int c = s_cntw[ww][tid]; s_cntw[ww][tid] = runb; runb += c; }
                    st.s_base[tid] = runb;
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < FE_ITEMS; ++k) {
                    const int i = t0 + wid * (32 * FE_ITEMS) + k * 32 + lane;
                    if (i < V) {
                        const unsigned int d = (unsigned int)(key[k] >> shift) & 255u;
                        __stcg(dst + s_cntw[wid][d] + rank[k], key[k]);
                    }
                }
                __syncthreads();
            }
        }
    }
    const unsigned long long* recs = (npass & 1) ? gs.recB : gs.recA;

    // ---- Ch: vstart[r] = first grouped position of voxel r (key order); cursor[slot] = the same ----
    unsigned int* s_cursor = reinterpret_cast<unsigned int*>(fe_dyn);
    {
        int carry = 0;
        for (int r0 = 0; r0 < V; r0 += NT) {
            const int r = r0 + tid;
            const unsigned long long rec = r < V ? __ldcg(recs + r) : 0ull;
            const int c = (int)(rec & 0xffffull);
            int total;
            const int ex = carry + block_excl_scan<NT>(c, st.s_w, &total);
            if (r < V) {
                __stcg(gs.vstart + r, (unsigned int)ex);
                s_cursor[(unsigned int)(rec >> 16) & 0x7fffu] = (unsigned int)ex;
            }
            carry += total;
            __syncthreads();
        }
        if (tid == 0) __stcg(gs.vstart + V, (unsigned int)carry);
        if (carry != N) { why(16); return false; }   // cannot happen; never publish a frame whose counts do not add up
    }
    __syncthreads();
    // ---- Dh: group the point indices by voxel. Every warp takes a contiguous range of the points and walks it in order; inside a
    //      warp a run of consecutive points of one voxel (the usual case in raster order) claims its places with ONE atomic, so the
    //      points of a voxel arrive in ascending order except where two warps' ranges meet: Eh's sort has next to nothing to do ----
    {
        const int per = (((N + NW - 1) / NW) + 31) & ~31;
        const int w0 = min(N, wid * per), w1 = min(N, wid * per + per);
        for (int b = w0; b < w1; b += 32) {
            const int pos = b + lane;
            const bool v = pos < w1;
            const unsigned int sl = v ? (unsigned int)__ldcg(gs.slotid + pos) : 0xffffffffu;
            const unsigned int prev = __shfl_up_sync(FULL_MASK, sl, 1);
            const bool head = v && (lane == 0 || sl != prev);
            const unsigned int hm = __ballot_sync(FULL_MASK, head), vm = __ballot_sync(FULL_MASK, v);
            const int hl = max(0, 31 - __clz(hm & (0xffffffffu >> (31 - lane))));   // the head of this lane's run
            unsigned int dest = 0;
            if (head) {
                const unsigned int above = hm & ~((2u << lane) - 1u);
                const int nxt = above ? (__ffs(above) - 1) : __popc(vm);            // valid lanes are a prefix of the warp
                dest = atomicAdd(&s_cursor[sl], (unsigned int)(nxt - lane));
            }
            dest = __shfl_sync(FULL_MASK, dest, hl);
            if (v) __stcg(gs.sidx + dest + (unsigned int)(lane - hl), (unsigned int)pos);
        }
    }
    __syncthreads();

    // ---- Eh: sorted runs -> sequential float centroids ----
    {
        unsigned int* s_idx = reinterpret_cast<unsigned int*>(fe_dyn);            // the cursors are dead from here on
        float* s_px = reinterpret_cast<float*>(s_idx + FEH_RCAP);
        float* s_py = s_px + FEH_RCAP;
        float* s_pz = s_py + FEH_RCAP;
        unsigned int* s_pw = reinterpret_cast<unsigned int*>(s_pz + FEH_RCAP);
        unsigned int* s_big = s_pw + FEH_RCAP;                                    // [FEH_MAXRUN] bitonic scratch of one long run per warp round
        const float4* pts = p.pts + (size_t)f * p.Pout;
        float4* vox = a.vox + (size_t)f * a.P;
        int* vcount = a.vcount ? a.vcount + (size_t)f * a.P : nullptr;
        unsigned long long hv[1] = {0ull};
        int r0 = 0;
        while (r0 < V) {
            const unsigned int base = __ldcg(gs.vstart + r0);
            const int r = r0 + tid;
            unsigned int s = 0, e = 0;
            bool fits = false;
            if (r < V) { s = __ldcg(gs.vstart + r); e = __ldcg(gs.vstart + r + 1); fits = (e - base) <= (unsigned int)FEH_RCAP; }
            const int nv = __syncthreads_count(fits ? 1 : 0);     // runs are contiguous: the fitting voxels are a prefix
            if (tid == nv - 1) s_misc[3] = (int)(e - base);
            if (tid == 0) *st.s_ndef = 0;
            __syncthreads();
            const int npts = s_misc[3];
            for (int l = tid; l < npts; l += NT) s_idx[l] = __ldcg(gs.sidx + base + l);
            __syncthreads();
            const int lp = (int)(s - base), c = (int)(e - s);
            bool deferred = false;
            if (tid < nv) {
                if (c <= FEH_LONG) {
                    for (int i = lp + 1; i < lp + c; ++i) {      // insertion sort: ascending point index
                        const unsigned int v = s_idx[i];
                        int j = i - 1;
                        while (j >= lp && s_idx[j] > v) { s_idx[j + 1] = s_idx[j]; --j; }
                        s_idx[j + 1] = v;
                    }
                } else {
                    deferred = true;
                    const int d = atomicAdd(st.s_ndef, 1);
                    if (d < FE_MAXDEF) { st.s_def_lp[d] = lp; st.s_def_pos[d] = tid; }
                }
            }
            __syncthreads();
            const int ndef = min(*st.s_ndef, FE_MAXDEF);          // <= RCAP / LONG = 64 = FE_MAXDEF long runs fit a chunk
            // long runs arrive sorted unless several warps contributed to them (Dh): a warp checks each, and only the unsorted ones
            // (the origin voxel that collects the zero-depth pixels of the whole frame, voxels where two ranges meet) are sorted,
            // one after the other, by a block-wide bitonic network in the scratch (padded with ~0 to a power of two)
            for (int d = wid; d < ndef; d += NW) {
                const int dl = st.s_def_lp[d];
                const int vt = st.s_def_pos[d] & 0xffff;
                const int dc = (int)(__ldcg(gs.vstart + r0 + vt + 1) - __ldcg(gs.vstart + r0 + vt));
                bool bad = false;
                for (int i = lane; i + 1 < dc; i += 32) bad = bad || (s_idx[dl + i] > s_idx[dl + i + 1]);
                bad = __any_sync(FULL_MASK, bad);
                if (lane == 0) st.s_def_pos[d] = vt | (bad ? 0x10000 : 0);
            }
            __syncthreads();
            for (int d = 0; d < ndef; ++d) {
                if (!(st.s_def_pos[d] & 0x10000)) continue;        // block-uniform
                const int dl = st.s_def_lp[d];
                const int vt = st.s_def_pos[d] & 0xffff;
                const int dc = (int)(__ldcg(gs.vstart + r0 + vt + 1) - __ldcg(gs.vstart + r0 + vt));
                int n2 = 64;
                while (n2 < dc) n2 <<= 1;
                for (int i = tid; i < n2; i += NT) s_big[i] = i < dc ? s_idx[dl + i] : 0xffffffffu;
                __syncthreads();
                for (int k = 2; k <= n2; k <<= 1)
                    for (int j = k >> 1; j > 0; j >>= 1) {
                        for (int t = tid; t < n2; t += NT) {
                            const int u = t ^ j;
                            if (u > t) {
                                const unsigned int x = s_big[t], y = s_big[u];
                                const bool up = (t & k) == 0;
                                if ((x > y) == up) { s_big[t] = y; s_big[u] = x; }
                            }
                        }
                        __syncthreads();
                    }
                for (int i = tid; i < dc; i += NT) s_idx[dl + i] = s_big[i];
                __syncthreads();
            }
            for (int l = tid; l < npts; l += NT) {
                const float4 pt = __ldcg(pts + s_idx[l]);
                s_px[l] = pt.x; s_py[l] = pt.y; s_pz[l] = pt.z;
                if (RGB) s_pw[l] = __float_as_uint(pt.w);
            }
            __syncthreads();
            if (tid < nv && !deferred) {
                float sx = s_px[lp], sy = s_py[lp], sz = s_pz[lp];   // centroid starts as the first point, then += in ascending point order
                for (int q = lp + 1; q < lp + c; ++q) { sx += s_px[q]; sy += s_py[q]; sz += s_pz[q]; }
                const float cnt = (float)c;
                const float cx = sx / cnt, cy = sy / cnt, cz = sz / cnt;
                float cw = 1.0f;
                if (RGB) {
                    unsigned int sr = 0, sg = 0, sb = 0;
                    for (int q = lp; q < lp + c; ++q) { const unsigned int w = s_pw[q]; sr += (w >> 16) & 255u; sg += (w >> 8) & 255u; sb += w & 255u; }
                    cw = fe_rgb_avg(sr, sg, sb, c);
                }
                vox[r] = make_float4(cx, cy, cz, cw);
                if (vcount) vcount[r] = c;
                if (a.hashes) hv[0] += hash_point((unsigned int)r, cx, cy, cz);
            }
            for (int d = wid; d < ndef; d += NW) {   // long runs: a whole warp, lane-parallel loads, in-order sums (fe_run_step)
                const int dl = st.s_def_lp[d];
                const int vt = st.s_def_pos[d] & 0xffff;
                const int dc = (int)(__ldcg(gs.vstart + r0 + vt + 1) - __ldcg(gs.vstart + r0 + vt));
                FeRun acc;
                acc.sx = s_px[dl]; acc.sy = s_py[dl]; acc.sz = s_pz[dl]; acc.cnt = 1;
                {
                    const unsigned int w0 = RGB ? s_pw[dl] : 0u;
                    acc.sr = (w0 >> 16) & 255u; acc.sg = (w0 >> 8) & 255u; acc.sb = w0 & 255u;
                }
                for (int l = dl + 1; l < dl + dc; l += 32) {
                    const int i = l + lane;
                    const bool m = i < dl + dc;
                    fe_run_step<RGB>(acc, m, m ? s_px[i] : 0.f, m ? s_py[i] : 0.f, m ? s_pz[i] : 0.f, (m && RGB) ? s_pw[i] : 0u);
                }
                if (lane == 0) {
                    const float cf = (float)acc.cnt;
                    const float cx = acc.sx / cf, cy = acc.sy / cf, cz = acc.sz / cf;
                    const int vp = r0 + vt;
                    vox[vp] = make_float4(cx, cy, cz, RGB ? fe_rgb_avg(acc.sr, acc.sg, acc.sb, acc.cnt) : 1.0f);
                    if (vcount) vcount[vp] = acc.cnt;
                    if (a.hashes) hv[0] += hash_point((unsigned int)vp, cx, cy, cz);
                }
            }
            __syncthreads();
            r0 += nv;
        }
        // commit: voxel count and parity hashes
        const unsigned long long h0 = st.s_h64[0], h1 = st.s_h64[1];
        __syncthreads();
        fe_block_sum_u64<NT, 1>(hv, st.s_h64);
        if (tid == 0) {
            p.res[f].n_voxels = V;
            if (a.hash == 2) atomicOr(&p.res[f].status, 32 << 8);
            if (h0) atomic_add_u64(&p.res[f].points_hash, h0);
            if (h1) atomic_add_u64(&p.res[f].voxel_key_hash, h1);
            if (hv[0]) atomic_add_u64(&p.res[f].voxel_hash, hv[0]);
        }
    }
    return true;
}

// RGB (PointCloud2 inputs with a packed rgb field, FrontArgs::rgb) is a compile-time variant: the colour bookkeeping costs the
// depth-frame path 4 % when it is a run-time test.
// RUNS (run records, depth input only) likewise: compiled into the common kernel it costs the default path registers (spills).
// SOLO: depth input, one CTA per frame, one-pass mode, no parity taps - what a throughput launch runs. As its own instance the cluster
// exchange, the A1 pass and the taps are compiled out (registers, and a CTA barrier where the general kernel has a cluster barrier).
template <int SRC, int NT, bool RGB, bool RUNS = false, bool SOLO = false>
__global__ void __launch_bounds__(NT, NT <= 256 ? 4 : (NT <= 512 ? 2 : 1)) k_frontend(const FrontArgs a) {
    constexpr int NW = NT / 32;
    constexpr int TILE = NT * FE_ITEMS;
    constexpr int RTILE = NT * FE_RITEMS;
    extern __shared__ __align__(16) unsigned char fe_dyn[];
    __shared__ FeXchg s_x;
    __shared__ FeFrame s_f;
    __shared__ unsigned int s_hist[256];
    __shared__ unsigned int s_base[256];
    __shared__ unsigned int s_histA[4][256];   // C == 1: digit histograms of all passes, counted while the keys are generated
    __shared__ int s_w[NW + 1];
    __shared__ float s_mm[NW][6];
    __shared__ int s_wc[NW];
    __shared__ unsigned long long s_h64[2 * NW];
    __shared__ int s_misc[8];              // hash path: distinct voxels, abort flag, longest run, chunk size
    __shared__ int s_ndef;                 // voxels of the current reduce tile deferred to the warp-cooperative routine
    __shared__ int s_def_lp[FE_MAXDEF], s_def_pos[FE_MAXDEF];

    cg::cluster_group cluster = cg::this_cluster();
    static_assert(!SOLO || (SRC == 0 && !RGB), "SOLO is the depth-frame instance");
    static_assert(!RUNS || SOLO, "run records need the one-pass mode");
    const int C = SOLO ? 1 : (int)cluster.num_blocks(), r = SOLO ? 0 : (int)cluster.block_rank();
    auto csync = [&]() { if (SOLO) __syncthreads(); else cluster.sync(); };
    const bool tap_hashes = !SOLO && a.hashes != 0;
    const int slot = blockIdx.x / C, nslots = gridDim.x / C;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const PreArgs& p = a.pre;
    unsigned long long* bufA = a.keys + (size_t)slot * a.keys_stride;
    unsigned long long* bufB = bufA + a.P;

    for (int f = slot; f < a.n_frames; f += nslots) {
        const int n_in = (SRC == 1 && p.n_in) ? p.n_in[f] : p.P;
        int slice = (n_in + C - 1) / C;
        slice = (slice + 7) & ~7;
        const int s0 = min(n_in, r * slice), s1 = min(n_in, r * slice + slice);
        unsigned char* s_mask = fe_dyn + NT * 32;
        const bool hash_try = !SOLO && NT == 1024 && C == 1 && a.hash != 0;    // the hash path recomputes the masks (its table takes their room)
        const bool onepass = SOLO || (SRC == 0 && C == 1 && a.st_on != 0 && !a.kpp && !a.hashes && !hash_try);   // kernel-uniform
        const bool use_mask = !hash_try && !onepass && ((slice + TILE - 1) / TILE) * NT <= FE_MASK_BYTES && NT * 32 + FE_MASK_BYTES <= a.arena;
        const bool runs = RUNS && a.runs != 0;
        int NR = 0;                      // sort records of the frame: its points, or its runs

        // ---- A1: survivors and min/max of this CTA's input slice ----
        if (!onepass) {
            int cnt = 0;
            float mn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f};
            float mx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
            int it = 0;
            for (int first = s0 + tid * FE_ITEMS; first < s1; first += TILE, ++it) {
                float px[FE_ITEMS], py[FE_ITEMS], pz[FE_ITEMS];
                const unsigned int keep = pre_points<SRC>(p, f, first, s1, px, py, pz);
                if (use_mask) s_mask[it * NT + tid] = (unsigned char)keep;
                cnt += __popc(keep);
#pragma unroll
                for (int k = 0; k < FE_ITEMS; ++k)
                    if (keep & (1u << k)) {
                        mn[0] = fminf(mn[0], px[k]); mx[0] = fmaxf(mx[0], px[k]);
                        mn[1] = fminf(mn[1], py[k]); mx[1] = fmaxf(mx[1], py[k]);
                        mn[2] = fminf(mn[2], pz[k]); mx[2] = fmaxf(mx[2], pz[k]);
                    }
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) {
                cnt += __shfl_xor_sync(FULL_MASK, cnt, o);
#pragma unroll
                for (int c = 0; c < 3; ++c) {
                    mn[c] = fminf(mn[c], __shfl_xor_sync(FULL_MASK, mn[c], o));
                    mx[c] = fmaxf(mx[c], __shfl_xor_sync(FULL_MASK, mx[c], o));
                }
            }
            if (lane == 0) {
                s_wc[wid] = cnt;
#pragma unroll
                for (int c = 0; c < 3; ++c) { s_mm[wid][c] = mn[c]; s_mm[wid][3 + c] = mx[c]; }
            }
            __syncthreads();
            if (tid == 0) {
                int count = 0;
                float xmn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f};
                float xmx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
                for (int w = 0; w < NW; ++w) {
                    count += s_wc[w];
                    for (int c = 0; c < 3; ++c) { xmn[c] = fminf(xmn[c], s_mm[w][c]); xmx[c] = fmaxf(xmx[c], s_mm[w][3 + c]); }
                }
                s_x.count = count;
                for (int c = 0; c < 3; ++c) { s_x.mn[c] = xmn[c]; s_x.mx[c] = xmx[c]; }
            }
        }
        if (!onepass) csync();
        if (tid == 0 && !onepass) {
            int base = 0, N = 0;
            float mn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f};
            float mx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
            for (int rr = 0; rr < C; ++rr) {
                const FeXchg* q = cluster.map_shared_rank(&s_x, rr);
                const int c = q->count;
                if (rr < r) base += c;
                N += c;
                if (c > 0)
                    for (int k = 0; k < 3; ++k) { mn[k] = fminf(mn[k], q->mn[k]); mx[k] = fmaxf(mx[k], q->mx[k]); }
            }
            FrameScratch sc;
            sc.mm[0] = sc.mm[1] = sc.mm[2] = 0xffffffffu;
            sc.mm[3] = sc.mm[4] = sc.mm[5] = 0u;
            sc.sort_bits = 0; sc.overflow_mode = 0; sc.best_count = 0; sc.pad = 0;
            VoxelGeom g;
            g.inv = a.inv_leaf; g.sort_bits = 0; g.overflow_mode = 0; g.pcl_overflow = 0; g.mul1 = g.mul2 = 0;
            for (int k = 0; k < 3; ++k) { g.min_b[k] = 0; g.div_b[k] = 0; }
            if (N > 0) {
                for (int k = 0; k < 3; ++k) { sc.mm[k] = enc_f32(mn[k]); sc.mm[3 + k] = enc_f32(mx[k]); }
                g = voxel_geom(sc, a.inv_leaf);
                sc.sort_bits = g.sort_bits; sc.overflow_mode = g.overflow_mode;
            }
            s_f.base = base; s_f.N = N; s_f.g = g;
            if (r == 0) {
                p.scr[f] = sc;
                cuboid_frame_result& R = p.res[f];
                R.n_points = N;
                if (N > 0) {
                    for (int k = 0; k < 3; ++k) { R.min_b[k] = g.min_b[k]; R.div_b[k] = g.div_b[k]; }
                    if (g.pcl_overflow) atomicOr(&R.status, CUBOID_W_VOXEL_OVERFLOW);
                } else {
                    R.n_voxels = 0;
                }
            }
        }
        __syncthreads();
        int N = onepass ? 1 : s_f.N;     // one-pass mode learns N at the end of A2
        if (N == 0) {          // cluster-uniform; the barrier keeps a fast CTA from overwriting s_x while a peer still reads it
            csync();
            continue;
        }
        VoxelGeom g = s_f.g;             // one-pass mode: stale, not used before it is recomputed after A2
        if (onepass) { g.inv = a.inv_leaf; g.overflow_mode = 0; }
        const int npass = onepass ? (a.st_bits + 7) / 8 : (g.sort_bits + 7) / 8;
        if constexpr (NT == 1024) {
            if (hash_try) {
                FehStatic st;
                st.s_w = s_w; st.s_histA = s_histA; st.s_base = s_base; st.s_h64 = s_h64; st.s_ndef = &s_ndef; st.s_def_lp = s_def_lp;
                st.s_def_pos = s_def_pos; st.s_misc = s_misc;
                const bool done = fe_hash_frame<SRC, NT, RGB>(a, f, n_in, N, g, bufA, fe_dyn, st);
                __syncthreads();
                if (done) { csync(); continue; }
            }
        }
        if (C == 1) {
            for (int k = tid; k < 4 * 256; k += NT) (&s_histA[0][0])[k] = 0u;
            __syncthreads();
        }

        // ---- A2: ordered compaction of the slice. The divergent part only records WHICH inputs survive (u16 tile-local
        //      index, in order); the per-point work then runs dense, one survivor per thread, with coalesced stores ----
        {
            unsigned short* s_sel = reinterpret_cast<unsigned short*>(fe_dyn);
            unsigned short* s_dep = reinterpret_cast<unsigned short*>(fe_dyn + NT * 16);
            float4* out = p.pts + (size_t)f * p.Pout;
            int* kpp = (!SOLO && a.kpp) ? a.kpp + (size_t)f * a.P : nullptr;
            int run = onepass ? 0 : s_f.base;
            int rrun = 0;                                                                                     // run records written so far
            unsigned long long* s_run = reinterpret_cast<unsigned long long*>(fe_dyn + NT * 32);            // [TILE / 32][32] records of a tile, per warp step
            int* s_rc = reinterpret_cast<int*>(fe_dyn + NT * 32 + TILE * 8);                                  // [2][TILE / 32] their counts, then offsets
            unsigned long long hh[2] = {0ull, 0ull};
            float omn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f};      // one-pass mode: min / max of the survivors
            float omx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
            int it = 0;
            for (int t0 = s0; t0 < s1; t0 += TILE, ++it) {
                const int first = t0 + tid * FE_ITEMS;
                unsigned int keep = 0;
                if (SRC == 0 && first < s1) {   // the tile's depth values -> shared memory: the dense loop below gathers from there
                    const uint16_t* dsrc = p.depth + (size_t)f * p.P;
                    if (first + FE_ITEMS <= s1 && ((((size_t)f * p.P + first) & 7) == 0))
                        *reinterpret_cast<uint4*>(s_dep + tid * FE_ITEMS) = *reinterpret_cast<const uint4*>(dsrc + first);
                    else
                        for (int k = 0; k < FE_ITEMS; ++k) s_dep[tid * FE_ITEMS + k] = (first + k < s1) ? dsrc[first + k] : (uint16_t)0;
                }
                if (first < s1) {
                    if (use_mask) keep = s_mask[it * NT + tid];
                    else {
                        float px[FE_ITEMS], py[FE_ITEMS], pz[FE_ITEMS];
                        keep = pre_points<SRC>(p, f, first, s1, px, py, pz);
                        if (onepass) {
#pragma unroll
                            for (int k = 0; k < FE_ITEMS; ++k)
                                if (keep & (1u << k)) {
                                    omn[0] = fminf(omn[0], px[k]); omx[0] = fmaxf(omx[0], px[k]);
                                    omn[1] = fminf(omn[1], py[k]); omx[1] = fmaxf(omx[1], py[k]);
                                    omn[2] = fminf(omn[2], pz[k]); omx[2] = fmaxf(omx[2], pz[k]);
                                }
                        }
                    }
                }
                int total;
                int lp = block_excl_scan<NT>(__popc(keep), s_w, &total);
#pragma unroll
                for (int k = 0; k < FE_ITEMS; ++k)
                    if (keep & (1u << k)) s_sel[lp++] = (unsigned short)(tid * FE_ITEMS + k);
                __syncthreads();
                for (int qb = 0; qb < total; qb += NT) {   // uniform trip count: the warp votes below need every lane
                    const int q = qb + tid;
                    const bool valid = q < total;
                    unsigned int sk = 0u;
                    if (valid) {
                        const int sel = (int)s_sel[q];
                        const float4 pt = SRC == 0 ? fe_point_from_depth(p, t0 + sel, s_dep[sel]) : fe_point_at<SRC>(p, f, t0 + sel);
                        const int pos = run + q;
                        out[pos] = pt;
                        int idx = 0;
                        if (onepass) {   // (k, j, i) relative to the static bounds: orders like PCL's idx whatever the frame's own min / max turn out to be
                            const int i0 = (int)floorf(pt.x * g.inv) - a.st_min[0], i1 = (int)floorf(pt.y * g.inv) - a.st_min[1],
                                      i2 = (int)floorf(pt.z * g.inv) - a.st_min[2];
                            sk = (unsigned int)i0 | ((unsigned int)i1 << a.st_b0) | ((unsigned int)i2 << a.st_b1);
                        } else {
                            idx = voxel_index(g, pt.x, pt.y, pt.z);
                            sk = g.overflow_mode ? ((unsigned int)idx ^ 0x80000000u) : (unsigned int)idx;
                        }
                        if (!runs) __stcg(bufA + pos, ((unsigned long long)sk << 32) | (unsigned int)pos);
                        if (kpp) kpp[pos] = idx;
                        if (tap_hashes) {
                            hh[0] += hash_point((unsigned int)pos, pt.x, pt.y, pt.z);
                            hh[1] += hash_index((unsigned int)pos, idx);
                        }
                    }
                    bool count = C == 1 && valid;   // one CTA owns the frame: every pass's digit histogram is order-independent, count it here
                    if (runs) {   // heads of the runs of equal keys among this warp step's 32 consecutive survivors -> the step's slots of s_run, in order
                        const unsigned int prev = __shfl_up_sync(FULL_MASK, sk, 1);
                        const bool head = valid && (lane == 0 || sk != prev);
                        const unsigned int hb = __ballot_sync(FULL_MASK, head), vb = __ballot_sync(FULL_MASK, valid);
                        const int wb = (qb / NT) * NW + wid;
                        if (head) {
                            const unsigned int above = hb & ~((2u << lane) - 1u);
                            const int len = (above ? (__ffs(above) - 1) : __popc(vb)) - lane;   // the valid lanes are a prefix
                            s_run[wb * 32 + __popc(hb & ((1u << lane) - 1u))] =
                                ((unsigned long long)sk << 32) | ((unsigned long long)(unsigned int)(run + q) << 5) | (unsigned int)(len - 1);
                        }
                        if (lane == 0) s_rc[wb] = __popc(hb);
                        count = false;   // digit histograms: per record, in the dense copy below
                    }
                    if (count) {
#pragma unroll
                        for (int ps = 0; ps < 4; ++ps)
                            if (ps < npass) atomicAdd(&s_histA[ps][(sk >> (8 * ps)) & 255u], 1u);
                    }
                }
                if (runs) {   // the tile's run records, dense and in order, to the sort buffer
                    __syncthreads();
                    const int nwb = ((total + NT - 1) / NT) * NW;
                    int tot_r;
                    const int ex = block_excl_scan<NT>(tid < nwb ? s_rc[tid] : 0, s_w, &tot_r);
                    if (tid < nwb) s_rc[TILE / 32 + tid] = ex;
                    __syncthreads();
                    for (int wb = wid; wb < nwb; wb += NW)
                        if (lane < s_rc[wb]) {
                            const unsigned long long rec = s_run[wb * 32 + lane];
                            __stcg(bufA + rrun + s_rc[TILE / 32 + wb] + lane, rec);
                            const unsigned int rk = (unsigned int)(rec >> 32);
#pragma unroll
                            for (int ps = 0; ps < 4; ++ps)
                                if (ps < npass) atomicAdd(&s_histA[ps][(rk >> (8 * ps)) & 255u], 1u);
                        }
                    rrun += tot_r;
                }
                run += total;
                __syncthreads();
            }
            fe_block_sum_u64<NT, 2>(hh, s_h64);
            if (tid == 0) {
                if (hh[0]) atomic_add_u64(&p.res[f].points_hash, hh[0]);
                if (hh[1]) atomic_add_u64(&p.res[f].voxel_key_hash, hh[1]);
            }
            NR = runs ? rrun : 0;
            if (onepass) {   // what A1 and the frame set-up would have published: survivor count, min / max -> min_b / div_b
                N = run;
#pragma unroll
                for (int o = 16; o >= 1; o >>= 1)
#pragma unroll
                    for (int c = 0; c < 3; ++c) {
                        omn[c] = fminf(omn[c], __shfl_xor_sync(FULL_MASK, omn[c], o));
                        omx[c] = fmaxf(omx[c], __shfl_xor_sync(FULL_MASK, omx[c], o));
                    }
                if (lane == 0)
#pragma unroll
                    for (int c = 0; c < 3; ++c) { s_mm[wid][c] = omn[c]; s_mm[wid][3 + c] = omx[c]; }
                __syncthreads();
                if (tid == 0) {
                    float xmn[3] = {3.402823466e38f, 3.402823466e38f, 3.402823466e38f};
                    float xmx[3] = {-3.402823466e38f, -3.402823466e38f, -3.402823466e38f};
                    for (int w = 0; w < NW; ++w)
                        for (int c = 0; c < 3; ++c) { xmn[c] = fminf(xmn[c], s_mm[w][c]); xmx[c] = fmaxf(xmx[c], s_mm[w][3 + c]); }
                    FrameScratch sc;
                    sc.mm[0] = sc.mm[1] = sc.mm[2] = 0xffffffffu;
                    sc.mm[3] = sc.mm[4] = sc.mm[5] = 0u;
                    sc.sort_bits = 0; sc.overflow_mode = 0; sc.best_count = 0; sc.pad = 0;
                    cuboid_frame_result& R = p.res[f];
                    R.n_points = N;
                    if (N > 0) {
                        for (int k = 0; k < 3; ++k) { sc.mm[k] = enc_f32(xmn[k]); sc.mm[3 + k] = enc_f32(xmx[k]); }
                        const VoxelGeom gg = voxel_geom(sc, a.inv_leaf);
                        sc.sort_bits = gg.sort_bits; sc.overflow_mode = gg.overflow_mode;
                        for (int k = 0; k < 3; ++k) { R.min_b[k] = gg.min_b[k]; R.div_b[k] = gg.div_b[k]; }
                        if (gg.pcl_overflow) atomicOr(&R.status, CUBOID_W_VOXEL_OVERFLOW);
                    } else {
                        R.n_voxels = 0;
                    }
                    p.scr[f] = sc;
                    s_ndef = 0;
                }
                __syncthreads();
                if (N == 0) continue;   // block-uniform (C == 1: no cluster peers to keep in step)
            }
        }
        __threadfence();
        csync();

        // ---- B: stable LSD radix sort of the frame's N records, 8-bit digits, significant key bits only ----
        if (!runs) NR = N;
        int per = (NR + C - 1) / C;
        per = (per + 255) & ~255;       // whole warp runs (8 items x 32 lanes): tiles of a range stay warp-aligned
        const int q0 = min(NR, r * per), q1 = min(NR, r * per + per);
        unsigned int (*s_cnt)[256] = reinterpret_cast<unsigned int (*)[256]>(fe_dyn);
        for (int pass = 0; pass < npass; ++pass) {
            const unsigned long long* src = (pass & 1) ? bufB : bufA;
            unsigned long long* dst = (pass & 1) ? bufA : bufB;
            const int shift = 32 + pass * 8;
            if (C > 1) {   // digit histogram of this CTA's range
                if (tid < 256) s_hist[tid] = 0;
                __syncthreads();
                for (int i0 = q0 + tid; i0 < q1; i0 += 4 * NT) {
                    unsigned long long k4[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) k4[j] = (i0 + j * NT < q1) ? __ldcg(src + i0 + j * NT) : 0ull;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        if (i0 + j * NT < q1) atomicAdd(&s_hist[(unsigned int)(k4[j] >> shift) & 255u], 1u);
                }
                __syncthreads();
                csync();
            }
            {   // digit d = tid: records of digit d in front of this CTA's = all smaller digits + digit d of lower ranks
                unsigned int tot = 0, before = 0;
                if (tid < 256) {
                    if (C == 1) tot = s_histA[pass][tid];
                    else
                        for (int rr = 0; rr < C; ++rr) {
                            const unsigned int v = cluster.map_shared_rank(s_hist, rr)[tid];
                            tot += v;
                            if (rr < r) before += v;
                        }
                }
                int total;
                const int ex = block_excl_scan<NT>((int)tot, s_w, &total);
                if (tid < 256) s_base[tid] = (unsigned int)ex + before;
                __syncthreads();
            }
            // Full tiles are prefetched one tile ahead with 16-byte asynchronous copies (L2 only, like __ldcg): every warp fetches
            // its own 256 records (2 KB), lane l the 16-byte chunks l, l + 32, l + 64, l + 96, so a __syncwarp publishes them.
            unsigned long long* s_pre = reinterpret_cast<unsigned long long*>(fe_dyn + NT * 32);
            auto prefetch = [&](int t0, int buf) {
                if (t0 < q1 && t0 + TILE <= q1) {
                    const unsigned long long* g = src + t0 + wid * (32 * FE_ITEMS);
                    unsigned long long* sdst = s_pre + buf * TILE + wid * (32 * FE_ITEMS);
#pragma unroll
                    for (int j = 0; j < FE_ITEMS / 2; ++j) fe_cp_async16(sdst + 2 * (j * 32 + lane), g + 2 * (j * 32 + lane));
                }
                fe_cp_async_commit();
            };
            prefetch(q0, 0);
            constexpr int buf = 0;   // ONE staging buffer: a warp has its tile in registers before it requests the next one
            for (int t0 = q0; t0 < q1; t0 += TILE) {
                for (int d = lane; d < 256; d += 32) s_cnt[wid][d] = 0;
                fe_cp_async_wait<0>();
                __syncwarp();
                // warp w owns the contiguous run [t0 + w*256, +256): (warp, item, lane) order == ascending input order
                unsigned long long key[FE_ITEMS];
                unsigned int rank[FE_ITEMS];
                const bool full = t0 + TILE <= q1;
#pragma unroll
                for (int k = 0; k < FE_ITEMS; ++k) {
                    const int i = t0 + wid * (32 * FE_ITEMS) + k * 32 + lane;
                    if (full) key[k] = s_pre[buf * TILE + wid * (32 * FE_ITEMS) + k * 32 + lane];
                    else key[k] = i < q1 ? __ldcg(src + i) : ~0ull;
                }
                __syncwarp();
                prefetch(t0 + TILE, 0);   // in flight while this tile is ranked and scattered
                auto rank_tile = [&](auto all_valid) {   // all but the last tile of a range are full: their lanes need no validity vote
                    constexpr bool ALL = decltype(all_valid)::value;
#pragma unroll
                    for (int k = 0; k < FE_ITEMS; ++k) {
                        const int i = t0 + wid * (32 * FE_ITEMS) + k * 32 + lane;
                        const bool valid = ALL || i < q1;
                        const unsigned int d = valid ? ((unsigned int)(key[k] >> shift) & 255u) : 256u;
                        const unsigned int peers = fe_digit_peers<ALL>(d);
                        const int leader = __ffs(peers) - 1;
                        unsigned int before = 0;
                        if (valid && lane == leader) { before = s_cnt[wid][d]; s_cnt[wid][d] = before + __popc(peers); }
                        before = __shfl_sync(FULL_MASK, before, leader);
                        rank[k] = before + __popc(peers & ((1u << lane) - 1u));
                        __syncwarp();
                    }
                };
                if (full) rank_tile(std::true_type{}); else rank_tile(std::false_type{});
                __syncthreads();
                if (tid < 256) {
                    unsigned int run = s_base[tid];
#pragma unroll 8
                    for (int ww = 0; ww < NW; ++ww) { const unsigned int c = s_cnt[ww][tid]; s_cnt[ww][tid] = run; run += c; }
                    s_base[tid] = run;
                }
                __syncthreads();
#pragma unroll
                for (int k = 0; k < FE_ITEMS; ++k) {
                    const int i = t0 + wid * (32 * FE_ITEMS) + k * 32 + lane;
                    if (i < q1) {
                        const unsigned int d = (unsigned int)(key[k] >> shift) & 255u;
                        __stcg(dst + s_cnt[wid][d] + rank[k], key[k]);
                    }
                }
                __syncthreads();
            }
            fe_cp_async_wait<0>();
            __threadfence();
            csync();
        }
        const unsigned long long* keys = (npass & 1) ? bufB : bufA;

        // ---- C1: voxels that start inside this CTA's range -> voxel ordinal base (nothing to exchange when C == 1) ----
        if (C > 1) {
            int heads = 0;
            for (int i = q0 + tid; i < q1; i += NT) {
                const unsigned int k = (unsigned int)(__ldcg(keys + i) >> 32);
                heads += (i == 0 || k != (unsigned int)(__ldcg(keys + i - 1) >> 32)) ? 1 : 0;
            }
#pragma unroll
            for (int o = 16; o >= 1; o >>= 1) heads += __shfl_xor_sync(FULL_MASK, heads, o);
            if (lane == 0) s_wc[wid] = heads;
            __syncthreads();
            if (tid == 0) {
                int t = 0;
                for (int w = 0; w < NW; ++w) t += s_wc[w];
                s_x.heads = t;
            }
            csync();
        }
        if (tid == 0) {
            int vbase = 0;
            for (int rr = 0; rr < r; ++rr) vbase += cluster.map_shared_rank(&s_x, rr)->heads;
            s_f.vbase = vbase;
            s_ndef = 0;
        }
        __syncthreads();

        // ---- C2: sequential float centroid of every voxel that starts in this range. A tile of sorted records (key and
        //      gathered point) is staged with parallel loads; heads are compacted; then ONE THREAD PER VOXEL sums its run
        //      out of shared memory in sorted (= ascending point index) order and the centroid stores are coalesced ----
        if (runs) {
            // Run records: a tile stages up to RS sorted records (key, first point, length), expands them IN ORDER into at most FE_RCAP
            // points in shared memory (the points of a voxel are then contiguous, as in the point-record path below), and one thread per
            // voxel sums its points. Tiles advance by whole voxels; the records past n_here are look-ahead for the tile's last voxels.
            constexpr int RS = NT * FE_RREC;
            constexpr int FE_RCAP = NT * FE_RCAPT;
            unsigned int* s_key = reinterpret_cast<unsigned int*>(fe_dyn);
            unsigned int* s_meta = s_key + RS;                          // first point << 5 | length - 1
            unsigned int* s_off = s_meta + RS;                          // [RS + 1] first staged point of a record
            float* s_px = reinterpret_cast<float*>(s_off + RS + 1);
            float* s_py = s_px + FE_RCAP;
            float* s_pz = s_py + FE_RCAP;
            unsigned short* s_head = reinterpret_cast<unsigned short*>(s_pz + FE_RCAP);   // [RS]
            const float4* pts = p.pts + (size_t)f * p.Pout;
            float4* vox = a.vox + (size_t)f * a.P;
            int vrun = 0;
            for (int t0 = 0; t0 < NR;) {
                const int n_avail = min(RS, NR - t0);
                for (int l = tid; l < n_avail; l += NT) {
                    const unsigned long long rec = __ldcg(keys + t0 + l);
                    s_key[l] = (unsigned int)(rec >> 32);
                    s_meta[l] = (unsigned int)rec;
                }
                const unsigned int prev_key = t0 > 0 ? (unsigned int)(__ldcg(keys + t0 - 1) >> 32) : 0u;
                if (tid == 0) { s_misc[0] = n_avail; s_misc[1] = 0; }
                __syncthreads();
                // thread t owns the records [t * FE_RREC, + FE_RREC): point offsets and voxel heads in one packed scan
                const int first = tid * FE_RREC;
                unsigned int meta[FE_RREC];
                unsigned int heads = 0;
                int npt = 0;
#pragma unroll
                for (int k = 0; k < FE_RREC; ++k) {
                    const int l = first + k;
                    meta[k] = 0u;
                    if (l < n_avail) {
                        meta[k] = s_meta[l];
                        npt += (int)(meta[k] & 31u) + 1;
                        if (t0 + l == 0 || s_key[l] != (l > 0 ? s_key[l - 1] : prev_key)) heads |= 1u << k;
                    }
                }
                int total;
                const int ex = block_excl_scan<NT>(npt | (__popc(heads) << 18), s_w, &total);
                int off = ex & 0x3ffff, hp = ex >> 18;
                const int nheads = total >> 18;
#pragma unroll
                for (int k = 0; k < FE_RREC; ++k) {
                    const int l = first + k;
                    if (l < n_avail) {
                        const int len = (int)(meta[k] & 31u) + 1;
                        s_off[l] = (unsigned int)off;
                        if (heads & (1u << k)) s_head[hp++] = (unsigned short)l;
                        if (off <= FE_RCAP && off + len > FE_RCAP) s_misc[0] = l;   // the first record that does not fit: staging stops here
                        off += len;
                        if (l == n_avail - 1) s_off[n_avail] = (unsigned int)off;
                    }
                }
                __syncthreads();
                const int n_stage = s_misc[0];                                      // records whose points are staged
                const bool last = t0 + n_stage == NR;
                const int n_here = last ? n_stage : max(n_stage - FE_LOOK, min(n_stage, 32));
                // expansion, in record order: the points of a voxel become contiguous in shared memory
#pragma unroll
                for (int k = 0; k < FE_RREC; ++k)
                    if (first + k < n_stage) {
                        const int len = (int)(meta[k] & 31u) + 1, o = (int)s_off[first + k];
                        const float4* src = pts + (meta[k] >> 5);
                        for (int q = 0; q < len; ++q) {
                            const float4 pt = src[q];
                            s_px[o + q] = pt.x; s_py[o + q] = pt.y; s_pz[o + q] = pt.z;
                        }
                    }
                __syncthreads();
                for (int v = tid; v < nheads; v += NT) {
                    const int lp = s_head[v];
                    if (lp >= n_here) continue;                                     // look-ahead: belongs to the next tile
                    const int end = (v + 1 < nheads) ? min((int)s_head[v + 1], n_stage) : n_stage;
                    const bool open = !last && (v + 1 == nheads || (int)s_head[v + 1] > n_stage);   // may continue past the staged records
                    const int pa = (int)s_off[lp], pb = (int)s_off[end];
                    const int pos = vrun + v;
                    if (open || pb - pa > FE_LONGRUN) {
                        const int d = atomicAdd(&s_ndef, 1);
                        s_def_lp[d] = lp | (open ? 0x10000 : 0) | (end << 17); s_def_pos[d] = pos;
                        continue;
                    }
                    float sx = s_px[pa], sy = s_py[pa], sz = s_pz[pa];
                    for (int q = pa + 1; q < pb; ++q) { sx += s_px[q]; sy += s_py[q]; sz += s_pz[q]; }
                    const float cnt = (float)(pb - pa);
                    vox[pos] = make_float4(sx / cnt, sy / cnt, sz / cnt, 1.0f);
                }
                // voxels of this tile = heads in front of n_here
                for (int v = tid; v < nheads; v += NT)
                    if ((int)s_head[v] < n_here && (v + 1 == nheads || (int)s_head[v + 1] >= n_here)) s_misc[1] = v + 1;
                __syncthreads();
                for (int d = wid; d < s_ndef; d += NW) {   // long or open voxels: a whole warp, lane-parallel loads, in-order sums
                    const int lp = s_def_lp[d] & 0xffff, end = s_def_lp[d] >> 17;
                    const bool open = (s_def_lp[d] & 0x10000) != 0;
                    const unsigned int mykey = s_key[lp];
                    const int pa = (int)s_off[lp], pb = (int)s_off[end];
                    FeRun acc;
                    acc.sx = s_px[pa]; acc.sy = s_py[pa]; acc.sz = s_pz[pa]; acc.cnt = 1; acc.sr = acc.sg = acc.sb = 0u;
                    for (int l = pa + 1; l < pb; l += 32) {
                        const int i = l + lane;
                        const bool m = i < pb;
                        fe_run_step<false>(acc, m, m ? s_px[i] : 0.f, m ? s_py[i] : 0.f, m ? s_pz[i] : 0.f);
                    }
                    bool more = open;
                    for (int q = t0 + n_stage; more && q < NR; q += 32) {   // the rest of the voxel from the sorted records, 32 records per round
                        const int i = q + lane;
                        const unsigned long long rec = i < NR ? __ldcg(keys + i) : 0ull;
                        const bool mine = i < NR && (unsigned int)(rec >> 32) == mykey;
                        const unsigned int mb = __ballot_sync(FULL_MASK, !mine);
                        const int nrec = mb ? (__ffs(mb) - 1) : 32;                 // leading records of this voxel
                        more = nrec == 32;
                        const int len = lane < nrec ? (int)((unsigned int)rec & 31u) + 1 : 0;
                        const int inc = warp_incl_scan(len, lane);
                        const int tot = __shfl_sync(FULL_MASK, inc, 31);
                        const int exl = inc - len;
                        const unsigned int pfirst = ((unsigned int)rec >> 5);
                        for (int c0 = 0; c0 < tot; c0 += 32) {                      // 32 points per step: point c0 + lane belongs to the record whose range holds it
                            const int want = c0 + lane;
                            int lo = 0;                                             // last record with exl <= want (binary search over the lanes)
#pragma unroll
                            for (int sft = 16; sft >= 1; sft >>= 1) {
                                const int cand = lo + sft;
                                const int ce = __shfl_sync(FULL_MASK, exl, cand & 31);
                                const int cl = __shfl_sync(FULL_MASK, len, cand & 31);
                                if (cand < 32 && cl > 0 && ce <= want) lo = cand;
                            }
                            const int re = __shfl_sync(FULL_MASK, exl, lo);
                            const unsigned int rp = __shfl_sync(FULL_MASK, pfirst, lo);
                            const bool m = want < tot;
                            float4 pt = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (m) pt = pts[rp + (unsigned int)(want - re)];
                            fe_run_step<false>(acc, m, pt.x, pt.y, pt.z);
                        }
                    }
                    if (lane == 0) {
                        const float c = (float)acc.cnt;
                        vox[s_def_pos[d]] = make_float4(acc.sx / c, acc.sy / c, acc.sz / c, 1.0f);
                    }
                }
                __syncthreads();
                vrun += s_misc[1];
                t0 += n_here;
                __syncthreads();
                if (tid == 0) s_ndef = 0;
            }
            if (tid == 0) p.res[f].n_voxels = vrun;
        } else {
            unsigned int* s_key = reinterpret_cast<unsigned int*>(fe_dyn);
            constexpr int RSTAGE = RTILE + FE_LOOK;
            float* s_px = reinterpret_cast<float*>(fe_dyn) + RSTAGE;
            float* s_py = s_px + RSTAGE;
            float* s_pz = s_py + RSTAGE;
            unsigned short* s_head = reinterpret_cast<unsigned short*>(s_pz + RSTAGE);
            unsigned int* s_pw = reinterpret_cast<unsigned int*>(s_head + RTILE);   // packed rgb(a) of the staged points (RGB only)
            const float4* pts = p.pts + (size_t)f * p.Pout;
            float4* vox = a.vox + (size_t)f * a.P;
            int* vcount = (!SOLO && a.vcount) ? a.vcount + (size_t)f * a.P : nullptr;
            int vrun = s_f.vbase;
            unsigned long long hv[1] = {0ull};
            for (int t0 = q0; t0 < q1; t0 += RTILE) {
                const int n_here = min(RTILE, q1 - t0);
                const int n_stage = min(n_here + FE_LOOK, N - t0);   // look-ahead records: the tile's last voxel usually ends in them
                {
                    unsigned long long rec[FE_RITEMS + 1];
#pragma unroll
                    for (int k = 0; k <= FE_RITEMS; ++k) {
                        const int l = k * NT + tid;
                        rec[k] = (l < n_stage && (k < FE_RITEMS || tid < FE_LOOK)) ? __ldcg(keys + t0 + l) : 0ull;
                    }
#pragma unroll
                    for (int k = 0; k <= FE_RITEMS; ++k) {
                        const int l = k * NT + tid;
                        if (l < n_stage && (k < FE_RITEMS || tid < FE_LOOK)) {
                            const float4 pt = __ldcg(pts + (unsigned int)rec[k]);
                            s_key[l] = (unsigned int)(rec[k] >> 32);
                            s_px[l] = pt.x; s_py[l] = pt.y; s_pz[l] = pt.z;
                            if (RGB) s_pw[l] = __float_as_uint(pt.w);
                        }
                    }
                }
                const unsigned int prev_key = t0 > 0 ? (unsigned int)(__ldcg(keys + t0 - 1) >> 32) : 0u;
                __syncthreads();
                const int first = tid * FE_RITEMS;
                unsigned int heads = 0;
#pragma unroll
                for (int k = 0; k < FE_RITEMS; ++k) {
                    const int lp = first + k;
                    if (lp < n_here && (t0 + lp == 0 || s_key[lp] != (lp > 0 ? s_key[lp - 1] : prev_key))) heads |= 1u << k;
                }
                int total;
                int hp = block_excl_scan<NT>(__popc(heads), s_w, &total);
#pragma unroll
                for (int k = 0; k < FE_RITEMS; ++k)
                    if (heads & (1u << k)) s_head[hp++] = (unsigned short)(first + k);
                __syncthreads();
                for (int v = tid; v < total; v += NT) {
                    const int lp = s_head[v];
                    int end = n_here;
                    if (v + 1 < total) end = (int)s_head[v + 1];
                    else { const unsigned int mk = s_key[lp]; while (end < n_stage && s_key[end] == mk) ++end; }
                    const int pos = vrun + v;
                    if ((end == n_stage && t0 + n_stage < N) || end - lp > FE_LONGRUN) {
                        // may continue past the tile (only the last voxel can), or long (e.g. the origin voxel that collects
                        // every zero-depth pixel): a whole warp takes it
                        const int d = atomicAdd(&s_ndef, 1);
                        s_def_lp[d] = lp; s_def_pos[d] = pos;
                        continue;
                    }
                    float sx = s_px[lp], sy = s_py[lp], sz = s_pz[lp];   // centroid starts as the first point, then += in sorted order
                    for (int q = lp + 1; q < end; ++q) { sx += s_px[q]; sy += s_py[q]; sz += s_pz[q]; }
                    const float cnt = (float)(end - lp);
                    const float cx = sx / cnt, cy = sy / cnt, cz = sz / cnt;
                    float cw = 1.0f;
                    if (RGB) {
                        unsigned int sr = 0, sg = 0, sb = 0;
                        for (int q = lp; q < end; ++q) { const unsigned int w = s_pw[q]; sr += (w >> 16) & 255u; sg += (w >> 8) & 255u; sb += w & 255u; }
                        cw = fe_rgb_avg(sr, sg, sb, end - lp);
                    }
                    vox[pos] = make_float4(cx, cy, cz, cw);
                    if (vcount) vcount[pos] = end - lp;
                    if (tap_hashes) hv[0] += hash_point((unsigned int)pos, cx, cy, cz);
                }
                vrun += total;
                __syncthreads();
                for (int d = wid; d < s_ndef; d += NW) {   // a whole warp per deferred voxel: lane-parallel loads, in-order sums
                    const int lp = s_def_lp[d];
                    const unsigned int mykey = s_key[lp];
                    FeRun acc;
                    acc.sx = s_px[lp]; acc.sy = s_py[lp]; acc.sz = s_pz[lp]; acc.cnt = 1;
                    {
                        const unsigned int w0 = RGB ? s_pw[lp] : 0u;
                        acc.sr = (w0 >> 16) & 255u; acc.sg = (w0 >> 8) & 255u; acc.sb = w0 & 255u;
                    }
                    bool open = true;
                    for (int l = lp + 1; open && l < n_stage; l += 32) {   // the part of the run that is staged in shared memory
                        const int i = l + lane;
                        const bool m = i < n_stage && s_key[i] == mykey;
                        open = fe_run_step<RGB>(acc, m, m ? s_px[i] : 0.f, m ? s_py[i] : 0.f, m ? s_pz[i] : 0.f, (m && RGB) ? s_pw[i] : 0u) == min(32, n_stage - l);
                    }
                    for (int q = t0 + n_stage; open && q < N; q += 128) {    // the rest from the sorted records, 128 per round trip
                        unsigned long long rec[4];
                        bool m[4];
                        float4 pt[4];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const int i = q + j * 32 + lane;
                            rec[j] = i < N ? __ldcg(keys + i) : 0ull;
                            m[j] = i < N && (unsigned int)(rec[j] >> 32) == mykey;
                        }
#pragma unroll
                        for (int j = 0; j < 4; ++j) pt[j] = m[j] ? __ldcg(pts + (unsigned int)rec[j]) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                        for (int j = 0; j < 4; ++j)
                            if (open) open = fe_run_step<RGB>(acc, m[j], pt[j].x, pt[j].y, pt[j].z, RGB ? __float_as_uint(pt[j].w) : 0u) == 32;
                    }
                    if (lane == 0) {
                        const float c = (float)acc.cnt;
                        const float cx = acc.sx / c, cy = acc.sy / c, cz = acc.sz / c;
                        const int vp = s_def_pos[d];
                        vox[vp] = make_float4(cx, cy, cz, RGB ? fe_rgb_avg(acc.sr, acc.sg, acc.sb, acc.cnt) : 1.0f);
                        if (vcount) vcount[vp] = acc.cnt;
                        if (tap_hashes) hv[0] += hash_point((unsigned int)vp, cx, cy, cz);
                    }
                }
                __syncthreads();
                if (tid == 0) s_ndef = 0;
            }
            if (tid == 0 && q1 == N && q0 < q1) p.res[f].n_voxels = vrun;   // the CTA that owns the tail knows the total
            fe_block_sum_u64<NT, 1>(hv, s_h64);
            if (tid == 0 && hv[0]) atomic_add_u64(&p.res[f].voxel_hash, hv[0]);
        }
    }
    if (!SOLO) cluster.sync();   // no CTA may exit while a peer can still read its shared memory
}

}  // namespace cuboid
