// Micro-probe: why is staging 75 KB into shared memory slow inside k_refine?  Each CTA repeatedly picks a block of `cols`
// columns (nrP floats each) at a pseudo-random position of its own region and stages it, by cp.async or by plain loads.
// Variants: region size per CTA (TLB reach), access method.  Prints cycles per staging for issue and wait.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define THREADS 512
__device__ __forceinline__ void cp16(void *s, const void *g) { uint32_t d = (uint32_t) __cvta_generic_to_shared(s); asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g) : "memory"); }
__global__ void __launch_bounds__(THREADS, 1) k(float *X, uint64_t regionFloats, uint32_t nrP, uint32_t cols, int iters, int mode, unsigned long long *out) {
    extern __shared__ __align__(16) float tile[];
    const uint32_t tid = threadIdx.x, nq = nrP / 4;
    float *base = X + (uint64_t) blockIdx.x * regionFloats;
    float *baseB = X + (uint64_t) (gridDim.x + blockIdx.x) * regionFloats;
    uint32_t rng = 12345u + blockIdx.x * 7919u;
    unsigned long long tIssue = 0, tWait = 0; float acc = 0; uint64_t prevP = 0;
    const uint64_t positions = regionFloats / nrP - cols;
    for (int it = 0; it < iters; it++) {
        rng = rng * 1664525u + 1013904223u;
        uint64_t p0 = (uint64_t) (rng >> 4) % positions;
        if (mode >= 2 && (it & 3)) p0 = prevP;                       /* child split: read what the parent just wrote */
        prevP = p0;
        const float *src = ((mode >= 2 && (it & 1)) ? baseB : base) + p0 * nrP;
        float *dstG = ((mode >= 2 && (it & 1)) ? base : baseB) + p0 * nrP;
        __syncthreads();
        long long t0 = clock64();
        if (mode == 0 || mode >= 2) {
            for (uint32_t i = tid; i < cols * nq; i += THREADS) { const uint32_t c = i / nq, q = i - c * nq; cp16(tile + c * nrP + 4 * (q ^ (c & 7u)), src + (size_t) c * nrP + 4 * q); }
            asm volatile("cp.async.commit_group;" ::: "memory");
        } else {
            for (uint32_t i = tid; i < cols * nq; i += THREADS) { const float4 v = *reinterpret_cast<const float4 *>(src + 4 * (size_t) i); *reinterpret_cast<float4 *>(tile + 4 * i) = v; }
        }
        long long t1 = clock64();
        if (mode == 0 || mode >= 2) asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncthreads();
        long long t2 = clock64();
        if (mode == 2) for (uint32_t i = tid; i < cols * nq; i += THREADS) *reinterpret_cast<float4 *>(dstG + 4 * (size_t) i) = *reinterpret_cast<const float4 *>(tile + 4 * i);
        if (mode == 3) for (uint32_t i = tid; i < cols * nq; i += THREADS) __stcg(reinterpret_cast<float4 *>(dstG + 4 * (size_t) i), *reinterpret_cast<const float4 *>(tile + 4 * i));
        acc += tile[(tid * 33) % (cols * nrP)];
        if (tid == 0) { tIssue += t1 - t0; tWait += t2 - t1; }
    }
    if (tid == 0) { out[blockIdx.x * 2] = tIssue; out[blockIdx.x * 2 + 1] = tWait; }
    if (acc == 123.456f) out[0] = 0;
}
int main() {
    const int ctas = 100, iters = 2000; const uint32_t nrP = 160, cols = 118;
    unsigned long long *out; cudaMallocManaged(&out, ctas * 2 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    for (int smemKB : {96})
    for (uint64_t regionMB : {64ull}) {
        const uint64_t regionFloats = regionMB * 1024 * 1024 / 4;
        float *X; if (cudaMalloc(&X, regionFloats * 4 * ctas * 2) != cudaSuccess) { printf("alloc failed\n"); return 1; }
        cudaMemset(X, 0, regionFloats * 4 * ctas * 2);
        for (int mode = 0; mode < 4; mode++) {
            k<<<ctas, THREADS, smemKB * 1024>>>(X, regionFloats, nrP, cols, iters, mode, out);
            if (cudaDeviceSynchronize() != cudaSuccess) { printf("kernel failed: %s\n", cudaGetErrorString(cudaGetLastError())); return 1; }
            double a = 0, b = 0; for (int i = 0; i < ctas; i++) { a += out[2 * i]; b += out[2 * i + 1]; }
            printf("smem %d KB region %4llu MB/CTA (total %.1f GB) %s: issue %.0f cycles, wait+sync %.0f cycles per staging of %u KB\n", smemKB, (unsigned long long) regionMB,
                   regionMB * ctas / 1024.0, mode == 0 ? "cp.async" : mode == 1 ? "ld.global+st.shared" : mode == 2 ? "cp.async, ping-pong write-back st" : "cp.async, ping-pong write-back st.cg", a / ctas / iters, b / ctas / iters, cols * nrP * 4 / 1024);
        }
        cudaFree(X);
    }
    return 0;
}
