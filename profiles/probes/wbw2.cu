// write-bandwidth probe 2 (scratch, not product): bulk asynchronous copies out of per-warp staging buffers that are RE-FILLED
// between copies, as the env kernel does -- how much smem must be in flight, and how much "think time" per buffer is free?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o wbw2 wbw2.cu && ./wbw2
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

// each warp owns DEPTH staging buffers of `chunk` bytes; per chunk: wait until the buffer was read, refill it (lanes, 16 B each),
// spin `think` cycles (stands for the step logic), fence, one bulk copy
template <int DEPTH>
__global__ void probe(unsigned char *out, size_t nchunks, int chunk, int think, int fill) {
    extern __shared__ __align__(128) unsigned char sm[];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    unsigned char *my = sm + (size_t)warp * DEPTH * chunk;
    size_t c = (size_t)blockIdx.x * nw + warp, stride = (size_t)gridDim.x * nw;
    int b = 0;
    for (; c < nchunks; c += stride, b = (b + 1 == DEPTH) ? 0 : b + 1) {
        if (lane == 0) asm volatile("cp.async.bulk.wait_group.read %0;" :: "n"(DEPTH - 1) : "memory");
        __syncwarp();
        unsigned char *buf = my + b * chunk;
        if (fill)
            for (int i = lane * 16; i < chunk; i += 512) *reinterpret_cast<uint4 *>(buf + i) = make_uint4(1, 2, 3, (unsigned)c);
        if (think) { long long t0 = clock64(); while (clock64() - t0 < think) {} }
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        __syncwarp();
        if (lane == 0) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;"
                         :: "l"(out + c * (size_t)chunk), "r"((uint32_t)__cvta_generic_to_shared(buf)), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
    }
    if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}

int main() {
    const size_t MB = 1 << 20;
    const int S = 6;
    unsigned char *buf[S];
    for (int i = 0; i < S; ++i) cudaMalloc(&buf[i], 280 * MB);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaFuncSetAttribute(probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    cudaFuncSetAttribute(probe<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024);
    const int chunk = 4160;
    printf("chunk %d B; columns: size_MB depth warps/CTA CTAs/SM think fill -> us GB/s (smem staged per SM KB)\n", chunk);
    for (size_t sz : {68 * MB, 272 * MB}) {
        const size_t nchunks = sz / chunk;
        for (int depth : {1, 2})
            for (int ctas : {4, 7, 8, 12})
                for (int think : {0, 2000, 6000})
                    for (int fill : {1}) {
                        const int warps = 4;
                        const int smem = warps * depth * chunk;
                        if ((size_t)smem * ctas > 220 * 1024) continue;
                        auto run = [&](int i) {
                            unsigned char *b = buf[i % S];
                            if (depth == 1) probe<1><<<148 * ctas, warps * 32, smem>>>(b, nchunks, chunk, think, fill);
                            else probe<2><<<148 * ctas, warps * 32, smem>>>(b, nchunks, chunk, think, fill);
                        };
                        for (int i = 0; i < 10; ++i) run(i);
                        cudaDeviceSynchronize();
                        cudaEventRecord(e0);
                        const int IT = 100;
                        for (int i = 0; i < IT; ++i) run(i);
                        cudaEventRecord(e1); cudaEventSynchronize(e1);
                        float ms; cudaEventElapsedTime(&ms, e0, e1);
                        printf("%3zu MB depth %d warps %d ctas %2d think %5d fill %d: %7.2f us %6.0f GB/s (%d KB) %s\n", sz / MB, depth, warps,
                               ctas, think, fill, ms / IT * 1e3, sz / (ms / IT * 1e-3) / 1e9, smem * ctas / 1024,
                               cudaGetErrorString(cudaGetLastError()));
                    }
    }
    return 0;
}
