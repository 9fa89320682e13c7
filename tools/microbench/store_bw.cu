// Microbenchmark: how fast can 373 MB of frame bytes be written on a B200, by mechanism?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o store_bw store_bw.cu && ./store_bw
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1); } } while (0)

__device__ __forceinline__ void bulk_store(void *g, const void *s, unsigned bytes) {
    unsigned sa = (unsigned)__cvta_generic_to_shared(s);
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(g), "r"(sa), "r"(bytes) : "memory");
}
__device__ __forceinline__ void commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void wait_all() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void wait_all_w() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

__global__ void stg128(uint4 *out, size_t n16) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (size_t)gridDim.x * blockDim.x) out[i] = z;
}

// each CTA writes a contiguous chunk list; issuing warps = nw (lane 0 of each), op size = op bytes
__global__ void bulk(unsigned char *out, size_t total, unsigned op, int nw, int waitmode) {
    extern __shared__ __align__(128) unsigned char sm[];
    for (unsigned i = threadIdx.x * 16; i < op; i += blockDim.x * 16) *reinterpret_cast<uint4 *>(sm + i) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane != 0 || warp >= nw) return;
    const size_t n_ops = total / op;
    const size_t issuers = (size_t)gridDim.x * nw;
    const size_t me = (size_t)blockIdx.x * nw + warp;
    int k = 0;
    for (size_t i = me; i < n_ops; i += issuers) {
        bulk_store(out + i * op, sm, op);
        commit();
        if (waitmode == 1 && (++k & 3) == 0) asm volatile("cp.async.bulk.wait_group.read 3;" ::: "memory");
    }
    if (waitmode == 2) wait_all_w(); else wait_all();
}

int main() {
    const size_t total = 373248000;
    unsigned char *buf;
    CK(cudaMalloc(&buf, total));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto timeit = [&](const char *name, auto fn) {
        for (int i = 0; i < 3; ++i) fn();
        CK(cudaDeviceSynchronize());
        cudaEventRecord(e0);
        const int reps = 20;
        for (int i = 0; i < reps; ++i) fn();
        cudaEventRecord(e1);
        CK(cudaDeviceSynchronize());
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        printf("%-48s %8.1f us  %7.1f GB/s\n", name, ms / reps * 1e3, total / (ms / reps * 1e-3) / 1e9);
    };
    timeit("cudaMemsetAsync", [&] { cudaMemsetAsync(buf, 0, total); });
    for (int g : {148 * 4, 148 * 8, 148 * 16})
        for (int b : {256, 512}) {
            char nm[64]; snprintf(nm, 64, "STG.128 grid=%d block=%d", g, b);
            timeit(nm, [&] { stg128<<<g, b>>>((uint4 *)buf, total / 16); });
        }
    CK(cudaFuncSetAttribute(bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
    for (unsigned op : {768u, 2880u, 11520u, 46080u, 92160u})
        for (int nw : {1, 4, 8})
            for (int ctas : {148, 296, 592}) {
                if ((size_t)op * 1 > 100 * 1024) continue;
                if (ctas == 592 && op > 46080u) continue;
                char nm[96]; snprintf(nm, 96, "bulk op=%u B issuers/CTA=%d CTAs=%d", op, nw, ctas);
                timeit(nm, [&] { bulk<<<ctas, 256, op>>>(buf, total, op, nw, 0); });
            }
    timeit("bulk op=2880 nw=8 CTAs=296 wait.read every 4", [&] { bulk<<<296, 256, 2880>>>(buf, total, 2880, 8, 1); });
    timeit("bulk op=768 nw=8 CTAs=296 wait.read every 4", [&] { bulk<<<296, 256, 768>>>(buf, total, 768, 8, 1); });
    timeit("bulk op=11520 nw=1 CTAs=296 full wait at end", [&] { bulk<<<296, 256, 11520>>>(buf, total, 11520, 1, 2); });
    return 0;
}
