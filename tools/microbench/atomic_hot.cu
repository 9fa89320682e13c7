// Microbenchmark: cost of returning atomics on hot addresses (one per warp, result used).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o atomic_hot atomic_hot.cu
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(unsigned *ctr, unsigned n_addr, unsigned *out, int per_warp, int use_result) {
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    unsigned acc = 0;
    for (int i = 0; i < per_warp; ++i) {
        unsigned a = (warp * 2654435761u + i * 40503u) % n_addr;
        unsigned v = 0;
        if (lane == 0) { if (use_result) v = atomicAdd(&ctr[a * 32], 7u); else atomicAdd(&ctr[a * 32], 7u); }
        v = __shfl_sync(0xffffffffu, v, 0);
        acc += v;
    }
    if (acc == 0xdeadbeef) out[0] = acc;
}
int main() {
    unsigned *ctr, *out; cudaMalloc(&ctr, 32 * 4 * 65536); cudaMalloc(&out, 4); cudaMemset(ctr, 0, 32 * 4 * 65536);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int use : {1, 0}) for (unsigned n_addr : {1u, 40u, 240u, 2000u, 8160u, 65536u}) for (int per_warp : {1, 4}) {
        const int warps = 92000 / per_warp, blocks = (warps * 32 + 255) / 256;
        k<<<blocks, 256>>>(ctr, n_addr, out, per_warp, use); cudaDeviceSynchronize();
        cudaEventRecord(e0); k<<<blocks, 256>>>(ctr, n_addr, out, per_warp, use); cudaEventRecord(e1); cudaDeviceSynchronize();
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("use_result=%d addresses=%6u atomics/warp=%d total=92000: %7.1f us  (%.1f ns per atomic)\n", use, n_addr, per_warp, ms * 1e3, ms * 1e6 / 92000);
    }
    return 0;
}
