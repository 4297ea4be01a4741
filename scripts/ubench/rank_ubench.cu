// Micro-benchmark: warp-level stable ranking of 8-bit digits (the core of a radix pass), keys in registers.
//   V0  match.any + leader shared atomicAdd + shuffle broadcast      (the round-1 kernel)
//   V1  match.any + leader plain LDS/STS (no atomic)
//   V2  8 ballots + LOP3 (no MATCH instruction) + leader plain LDS/STS
//   V3  match.any only (no histogram)      V4  8 ballots only
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
typedef unsigned int u32;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s line %d\n", #x, cudaGetErrorString(e), __LINE__); exit(1); } } while (0)
__device__ __forceinline__ u32 hash32(u32 x) { x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x; }
__device__ __forceinline__ u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }
constexpr int ITEMS = 16;

__device__ __forceinline__ u32 ballot_match(u32 d) {
    u32 peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const u32 v = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? v : ~v;
    }
    return peers;
}

__device__ __forceinline__ u32 redux_match(u32 d, int lane) {
    u32 peers = 0xffffffffu;
#pragma unroll
    for (int b = 0; b < 8; ++b) {
        const u32 v = __reduce_or_sync(0xffffffffu, ((d >> b) & 1u) << lane);
        peers &= ((d >> b) & 1u) ? v : ~v;
    }
    return peers;
}

template <int V>
__global__ void __launch_bounds__(256, 3) rank_kernel(u32 *out, int iters, u32 seed) {
    __shared__ u32 hist[8][256];
    __shared__ u32 tbl[8][256];
    if (V == 5) { for (int i = threadIdx.x; i < 8 * 256; i += 256) (&tbl[0][0])[i] = 0; __syncthreads(); }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const u32 lt = lanemask_lt();
    u32 acc = 0;
    u32 key[ITEMS];
#pragma unroll
    for (int i = 0; i < ITEMS; ++i) key[i] = hash32(seed + (blockIdx.x * 256 + threadIdx.x) * ITEMS + i);
    for (int it = 0; it < iters; ++it) {
        for (int i = lane; i < 256; i += 32) hist[warp][i] = 0;
        __syncwarp();
        u32 rank[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const u32 d = (key[i] >> ((it & 3) * 8)) & 255u;
            if (V == 5) {
                // peer mask through a warp-private table: OR the lane bit in, read the word back, leader clears it
                atomicOr(&tbl[warp][d], 1u << lane);
                __syncwarp();
                const u32 peers = tbl[warp][d];
                __syncwarp();
                if ((peers & lt) == 0u) tbl[warp][d] = 0;
                __syncwarp();
                rank[i] = peers;
            } else if (V == 6) {
                rank[i] = redux_match(d, lane);
            } else {
                rank[i] = (V == 2 || V == 4) ? ballot_match(d) : __match_any_sync(0xffffffffu, d);
            }
        }
        if (V <= 2 || V == 5) {
#pragma unroll
            for (int i = 0; i < ITEMS; ++i) {
                const u32 d = (key[i] >> ((it & 3) * 8)) & 255u;
                const u32 peers = rank[i];
                const u32 below = peers & lt;
                u32 old = 0;
                if (V == 0) {
                    if (below == 0u) old = atomicAdd(&hist[warp][d], (u32)__popc(peers));
                    old = __shfl_sync(0xffffffffu, old, __ffs(peers) - 1);
                } else {
                    old = hist[warp][d];          // every peer reads the same word (broadcast)
                    __syncwarp();
                    if (below == 0u) hist[warp][d] = old + __popc(peers);
                    __syncwarp();
                }
                rank[i] = old + __popc(below);
            }
        }
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) { acc += rank[i]; key[i] = key[i] * 1664525u + 1013904223u + rank[i]; }
    }
    out[blockIdx.x * 256 + threadIdx.x] = acc;
}

template <int V>
void run(const char *name, u32 *out, int sms) {
    const int iters = 400, grid = sms * 3;
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int r = 0; r < 3; ++r) {
        CK(cudaEventRecord(a));
        rank_kernel<V><<<grid, 256>>>(out, iters, 17u);
        CK(cudaEventRecord(b)); CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (r && ms < best) best = ms;
    }
    const double keys = (double)grid * 256 * ITEMS * iters;
    printf("%-52s %8.3f ms  %7.1f Gkeys/s  -> 2.56 G keys in %.2f ms\n", name, best, keys / best * 1e-6, 2.56e9 / (keys / best));
}

int main() {
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    u32 *out; CK(cudaMalloc(&out, (size_t)p.multiProcessorCount * 3 * 256 * 4));
    run<0>("V0 match.any + shared atomicAdd + shfl", out, p.multiProcessorCount);
    run<1>("V1 match.any + plain LDS/STS", out, p.multiProcessorCount);
    run<2>("V2 8 ballots + plain LDS/STS", out, p.multiProcessorCount);
    run<3>("V3 match.any only", out, p.multiProcessorCount);
    run<4>("V4 8 ballots only", out, p.multiProcessorCount);
    run<5>("V5 atomicOr table match + plain LDS/STS count", out, p.multiProcessorCount);
    run<6>("V6 8 REDUX.OR only", out, p.multiProcessorCount);
    return 0;
}
