// Dev check (run on the GPU box): div_counts(a, b) == IEEE a / b for the operand ranges of classes.cu.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/check_div scripts/check_div_counts.cu && /tmp/check_div
#include <cstdio>
#include <cstdint>
#include "../np_decision_tree_b200/csrc/class_div.cuh"

__device__ unsigned long long mix(unsigned long long x) {
    x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
    return x;
}

__global__ void check(unsigned long long seed, unsigned long long per_thread, unsigned long long *bad, double *ex) {
    unsigned long long id = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    unsigned long long nb = 0;
    for (unsigned long long i = 0; i < per_thread; ++i) {
        unsigned long long h = mix(seed + id * per_thread + i), g = mix(h + 0x9e3779b97f4a7c15ULL);
        // b: count in [1, 2^27); a: sum of squares, anywhere up to b * 2^26 (and some tiny / exact cases)
        const int bbits = 1 + (int)(h % 27);
        const double b = (double)(1 + (g % ((1ull << bbits) - 0)) % ((1ull << 27) - 1));
        const int abits = (int)((h >> 8) % 54);
        const double a = (double)(abits ? (mix(g) & ((1ull << abits) - 1)) : 0ull);
        const double want = __ddiv_rn(a, b), got = div_counts(a, b);
        if (__double_as_longlong(want) != __double_as_longlong(got)) {
            if (nb == 0 && atomicAdd(bad, 0ull) == 0) { ex[0] = a; ex[1] = b; }
            ++nb;
        }
    }
    if (nb) atomicAdd(bad, nb);
}

int main() {
    unsigned long long *bad; double *ex;
    cudaMallocManaged(&bad, 8); cudaMallocManaged(&ex, 16);
    *bad = 0;
    const unsigned long long per_thread = 4096, threads = 148ull * 32 * 256;
    for (int round = 0; round < 4; ++round) check<<<148 * 32, 256>>>(1234567ull * (round + 1), per_thread, bad, ex);
    cudaError_t e = cudaDeviceSynchronize();
    printf("checked %llu pairs, mismatches %llu (%s)", 4 * per_thread * threads, *bad, cudaGetErrorString(e));
    if (*bad) printf("  e.g. a=%.17g b=%.17g", ex[0], ex[1]);
    printf("\n");
    return *bad != 0;
}
