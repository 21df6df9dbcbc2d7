// Developer micro-benchmark: __match_any_sync vs 8-ballot peer masks vs shared atomics (cycles per warp instruction).
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(int mode, unsigned int* out, long long* cyc, int iters, int spread) {
    __shared__ unsigned int cnt[16][256];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 16 * 256; i += blockDim.x) (&cnt[0][0])[i] = 0;
    __syncthreads();
    unsigned int d = (threadIdx.x * 2654435761u >> 8) % spread;
    unsigned int acc = 0;
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        unsigned int peers;
        if (mode == 0) peers = __match_any_sync(0xffffffffu, d);
        else if (mode == 1) {
            peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < 8; ++b) { const unsigned int bal = __ballot_sync(0xffffffffu, (d >> b) & 1u); peers &= ((d >> b) & 1u) ? bal : ~bal; }
        } else if (mode == 2) { atomicAdd(&cnt[w & 15][d & 255], 1u); peers = d; }
        else if (mode == 3) { atomicAdd(&cnt[0][d & 255], 1u); peers = d; }
        else {      // match through a warp-private mask table: OR the lane bit in, read the mask back, leader clears it
            atomicOr(&cnt[w & 15][d & 255], 1u << lane);
            __syncwarp();
            peers = ((volatile unsigned int*)cnt[w & 15])[d & 255];
            __syncwarp();
            if (lane == __ffs(peers) - 1) cnt[w & 15][d & 255] = 0u;
            __syncwarp();
        }
        acc += __popc(peers);
        d = (d + (acc & 1) + 1) % spread;
    }
    long long t1 = clock64();
    if (lane == 0) cyc[blockIdx.x * (blockDim.x / 32) + w] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc + cnt[w & 15][lane];
}
int main() {
    unsigned int* out; long long* cyc;
    cudaMalloc(&out, 4 * 148 * 2 * 512); cudaMalloc(&cyc, 8 * 148 * 2 * 16);
    const char* names[] = {"match_any", "8 ballots", "atoms warp-private", "atoms shared row", "atomicOr match"};
    for (int spread : {1, 4, 32, 256})
        for (int mode = 0; mode < 5; ++mode) {
            const int iters = 2000;
            k<<<296, 512>>>(mode, out, cyc, iters, spread);
            cudaDeviceSynchronize();
            long long h[16];
            cudaMemcpy(h, cyc, sizeof h, cudaMemcpyDeviceToHost);
            printf("spread %3d  %-20s %.1f cycles per warp-iteration (16 warps/CTA, 2 CTAs/SM)\n", spread, names[mode], (double)h[0] / iters);
        }
    return 0;
}
