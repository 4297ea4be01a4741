// Micro-benchmark: where should the partition's side bits live?
// One warp moves a 512-position tile of (row id, y) to a (left run, right run) layout; the side of a
// row is bit `row` of an N-bit mask held in (0) global memory, (1) this CTA's shared memory,
// (2) the distributed shared memory of an 8-CTA cluster.  No look-back: only the data movement
// and the look-up are timed.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o /tmp/pl partition_lookup.cu
#include <cooperative_groups.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
namespace cg = cooperative_groups;
typedef unsigned int u32;
typedef unsigned long long u64;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s line %d\n", #x, cudaGetErrorString(e), __LINE__); exit(1); } } while (0)

__device__ __forceinline__ u32 hash32(u32 x) {
    x ^= x >> 16; x *= 0x7feb352dU; x ^= x >> 15; x *= 0x846ca68bU; x ^= x >> 16; return x;
}
__global__ void fill_kernel(int *rows, double *y, int64_t n, u32 N, u32 seed) {
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const u32 h = hash32((u32)i * 2654435761u + seed + (u32)(i >> 32));
        rows[i] = (int)(((u64)h * N) >> 32);
        y[i] = (double)h;
    }
}
__global__ void fill_mask(u32 *mask, int words, u32 seed) {
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) mask[i] = hash32(i + seed);
}
__device__ __forceinline__ int ld_stream_i32(const int *p) {
    int v; asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p)); return v;
}
__device__ __forceinline__ u32 lanemask_lt() { u32 m; asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m)); return m; }

constexpr int ITEMS = 16, TILE = 512;

// SCHED 0: consecutive tiles; G > 0: 256 columns, tile-major in groups of G consecutive tiles per column;
// outputs of a column go to two streams (left / right child) like the real partition
template <int MODE, int THREADS, int SCHED>
__global__ void __launch_bounds__(THREADS) part_kernel(const int *__restrict__ in, int *__restrict__ out,
                                                       const double *__restrict__ yin, double *__restrict__ yout,
                                                       int64_t n_tiles, const u32 *__restrict__ mask, int words,
                                                       int words_per_cta) {
    extern __shared__ u32 smask[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    constexpr int WARPS = THREADS / 32;
    u32 my_rank = 0;
    if (MODE == 1) {
        for (int i = threadIdx.x; i < words; i += THREADS) smask[i] = mask[i];
        __syncthreads();
    }
    if (MODE == 2) {
        cg::cluster_group cl = cg::this_cluster();
        my_rank = cl.block_rank();
        const int w0 = my_rank * words_per_cta;
        for (int i = threadIdx.x; i < words_per_cta && w0 + i < words; i += THREADS) smask[i] = mask[w0 + i];
        cl.sync();
    }
    const u32 lt = lanemask_lt();
    u32 sbase = (u32)__cvta_generic_to_shared(smask);
    for (int64_t t = (int64_t)blockIdx.x * WARPS + warp; t < n_tiles; t += (int64_t)gridDim.x * WARPS) {
        int64_t pt = t;          // physical tile
        int64_t obase = t * TILE, rbase = 0;
        if (SCHED > 0) {
            const int64_t tpc = n_tiles / 256;           // tiles per column
            const int64_t g = t / SCHED;
            const int w = (int)(t % SCHED);
            const int64_t f = g % 256, ti = (g / 256) * SCHED + w;
            if (ti >= tpc) continue;
            pt = f * tpc + ti;
            obase = f * tpc * 640 + ti * 320;            // left stream of the column
            rbase = tpc * 320;                           // right stream starts here
        }
        const int *col = in + pt * TILE;
        int rows[ITEMS];
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) rows[i] = ld_stream_i32(col + i * 32 + lane);
        u32 leftbits = 0;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) {
            const int row = rows[i];
            u32 w;
            if (MODE == 0) w = __ldg(mask + (row >> 5));
            else if (MODE == 1) w = smask[row >> 5];
            else {
                const u32 wi = (u32)row >> 5;
                const u32 cta = wi / (u32)words_per_cta;
                const u32 local = sbase + (wi - cta * words_per_cta) * 4;
                u32 remote;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local), "r"(cta));
                asm volatile("ld.shared::cluster.u32 %0, [%1];" : "=r"(w) : "r"(remote));
            }
            leftbits |= ((w >> (row & 31)) & 1u) << i;
        }
        int nleft = 0;
#pragma unroll
        for (int i = 0; i < ITEMS; ++i) nleft += __popc(__ballot_sync(0xffffffffu, (leftbits >> i) & 1u));
        int lb_row = 0;
        int *ocol = out + obase;
        const double *ycol = yin + pt * TILE;
        double *oy = yout + obase;
#pragma unroll
        for (int i0 = 0; i0 < ITEMS; i0 += 4) {
            double yv[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) yv[u] = __ldcs(ycol + (i0 + u) * 32 + lane);
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const int i = i0 + u;
                const u32 b = __ballot_sync(0xffffffffu, (leftbits >> i) & 1u);
                const int e = i * 32 + lane;
                const int lb = lb_row + __popc(b & lt);
                const int dest = ((leftbits >> i) & 1u) ? lb : (SCHED > 0 ? (int)rbase : nleft) + (e - lb);
                __stcs(ocol + dest, rows[i]);
                __stcs(oy + dest, yv[u]);
                lb_row += __popc(b);
            }
        }
    }
    if (MODE == 2) cg::this_cluster().sync();
}

template <int MODE, int THREADS, int SCHED = 0>
float run(const char *name, int grid, size_t smem, int cluster, const int *in, int *out, const double *yin, double *yout,
          int64_t n_tiles, const u32 *mask, int words, int wpc) {
    auto kern = part_kernel<MODE, THREADS, SCHED>;
    CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (cluster > 1) CK(cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(grid); cfg.blockDim = dim3(THREADS); cfg.dynamicSmemBytes = smem;
    cudaLaunchAttribute at[1];
    int na = 0;
    if (cluster > 1) { at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = cluster; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1; na = 1; }
    cfg.attrs = at; cfg.numAttrs = na;
    if (cluster > 1) {
        int nc = 0;
        cudaError_t e = cudaOccupancyMaxActiveClusters(&nc, kern, &cfg);
        printf("  [%s] max active clusters of %d: %d (%s)\n", name, cluster, nc, cudaGetErrorString(e));
        if (e != cudaSuccess || nc < 1) { cudaGetLastError(); return -1.f; }
        cfg.gridDim = dim3(nc * cluster);
        grid = nc * cluster;
    }
    cudaEvent_t a, b; CK(cudaEventCreate(&a)); CK(cudaEventCreate(&b));
    float best = 1e30f;
    for (int it = 0; it < 4; ++it) {
        CK(cudaEventRecord(a));
        CK(cudaLaunchKernelEx(&cfg, kern, in, out, yin, yout, n_tiles, mask, words, wpc));
        CK(cudaEventRecord(b));
        CK(cudaEventSynchronize(b));
        float ms; CK(cudaEventElapsedTime(&ms, a, b));
        if (it > 0 && ms < best) best = ms;
    }
    const double el = (double)n_tiles * TILE;
    printf("%-44s grid %4d smem %6zu: %8.3f ms  %.2f Gelem/s  %.0f GB/s (24 B/elem)\n", name, grid, smem, best, el / best * 1e-6, el * 24 / best * 1e-6);
    return best;
}

int main(int argc, char **argv) {
    const int64_t n_elem = argc > 1 ? atoll(argv[1]) : 1280000000ll;   // half a level of C4
    const int64_t n_tiles = n_elem / TILE;
    int *in, *out; double *yin, *yout; u32 *mask;
    CK(cudaMalloc(&in, n_elem * 4)); CK(cudaMalloc(&out, n_elem * 5 + 4096));
    CK(cudaMalloc(&yin, n_elem * 8)); CK(cudaMalloc(&yout, n_elem * 10 + 4096));
    const int max_words = 10000000 / 32 + 1;
    CK(cudaMalloc(&mask, max_words * 4 + 1024));
    fill_mask<<<256, 256>>>(mask, max_words, 7u);
    cudaDeviceProp p; CK(cudaGetDeviceProperties(&p, 0));
    const int sms = p.multiProcessorCount;
    printf("device %s, %d SMs, n_elem %lld\n", p.name, sms, (long long)n_elem);
    for (u32 N : {10000000u, 1250000u}) {
        fill_kernel<<<sms * 8, 256>>>(in, yin, n_elem, N, N);
        CK(cudaDeviceSynchronize());
        const int words = (N + 31) / 32;
        printf("--- rows in the window: %u (mask %d KB)\n", N, words * 4 / 1024);
        run<0, 256, 0>("consecutive tiles, 256 thr x 4/SM", sms * 4, 0, 1, in, out, yin, yout, n_tiles, mask, words, 0);
        run<0, 256, 1>("256 cols tile-major (today), 256 thr x 4", sms * 4, 0, 1, in, out, yin, yout, n_tiles, mask, words, 0);
        run<0, 256, 8>("256 cols groups of 8, 256 thr x 4", sms * 4, 0, 1, in, out, yin, yout, n_tiles, mask, words, 0);
        run<0, 256, 32>("256 cols groups of 32, 256 thr x 4", sms * 4, 0, 1, in, out, yin, yout, n_tiles, mask, words, 0);
        run<0, 1024, 32>("256 cols groups of 32, 1024 thr x 1", sms, 0, 1, in, out, yin, yout, n_tiles, mask, words, 0);
        run<0, 1024, 32>("256 cols groups of 32, 1024 thr x 2", sms * 2, 0, 1, in, out, yin, yout, n_tiles, mask, words, 0);
        run<0, 1024, 128>("256 cols groups of 128, 1024 thr x 2", sms * 2, 0, 1, in, out, yin, yout, n_tiles, mask, words, 0);
    }
    return 0;
}
