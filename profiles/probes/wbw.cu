// write-bandwidth probe: plain v4 stores vs bulk async copies from shared memory (scratch, not product)
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void fill_v4(uint4* p, size_t n16) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, st = (size_t)gridDim.x * blockDim.x;
    uint4 v = make_uint4(1, 2, 3, 4);
    for (; i < n16; i += st) p[i] = v;
}
template <int CHUNK>
__global__ void fill_bulk(unsigned char* p, size_t nbytes) {
    extern __shared__ __align__(128) unsigned char sm[];
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nw = blockDim.x >> 5;
    unsigned char* my = sm + (size_t)warp * CHUNK;
    for (int i = lane * 16; i < CHUNK; i += 512) *reinterpret_cast<uint4*>(my + i) = make_uint4(1, 2, 3, 4);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    size_t chunk = (size_t)blockIdx.x * nw + warp, nchunks = nbytes / CHUNK, stride = (size_t)gridDim.x * nw;
    if (lane == 0) {
        for (; chunk < nchunks; chunk += stride) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" :: "l"(p + chunk * CHUNK), "r"((uint32_t)__cvta_generic_to_shared(my)), "r"(CHUNK) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory");
        }
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
}
int main() {
    const size_t MB = 1 << 20; size_t sizes[2] = {68 * MB, 272 * MB};
    const int S = 6; unsigned char* buf[S];
    for (int i = 0; i < S; ++i) cudaMalloc(&buf[i], 272 * MB);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (size_t sz : sizes) {
        for (int variant = 0; variant < 4; ++variant) {
            auto run = [&](int i) {
                unsigned char* b = buf[i % S];
                if (variant == 0) fill_v4<<<148 * 8, 256>>>((uint4*)b, sz / 16);
                else if (variant == 1) fill_v4<<<148 * 16, 512>>>((uint4*)b, sz / 16);
                else if (variant == 2) fill_bulk<4096><<<148 * 4, 256, 8 * 4096>>>(b, sz);
                else fill_bulk<8192><<<148 * 4, 128, 4 * 8192>>>(b, sz);
            };
            for (int i = 0; i < 20; ++i) run(i);
            cudaDeviceSynchronize();
            cudaEventRecord(e0);
            const int IT = 200;
            for (int i = 0; i < IT; ++i) run(i);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            printf("size %zu MB variant %d: %.2f us  %.0f GB/s  err=%s\n", sz / MB, variant, ms / IT * 1e3, sz / (ms / IT * 1e-3) / 1e9, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
