// DRAM fetch granularity probe: every thread reads 8 bytes at a pseudo-random offset aligned to `align` bytes inside a 2 GiB
// buffer; ncu's dram__bytes_read.sum / (threads) tells how many bytes one isolated 32-byte sector request costs.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__global__ void probe(const uint2 *buf, uint64_t nSlots, uint32_t stride8, unsigned long long *sink, uint32_t salt)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t h = (i + salt) * 0x9E3779B97F4A7C15ull; h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    uint64_t slot = h % nSlots;
    uint2 v = __ldg(buf + slot * stride8);
    if (v.x == 0x12345678u && v.y == 0x9abcdef0u) atomicAdd(sink, 1ull);
}
int main(int argc, char **argv)
{
    const size_t bytes = 2ull << 30;
    uint2 *buf; unsigned long long *sink;
    cudaMalloc(&buf, bytes); cudaMemset(buf, 0, bytes); cudaMalloc(&sink, 8);
    const uint32_t n = 4u << 20;            // 4 M requests
    for (int align = 32; align <= 256; align *= 2) {
        uint32_t stride8 = align / 8;
        probe<<<n / 256, 256>>>(buf, bytes / align, stride8, sink, align);
        cudaDeviceSynchronize();
    }
    printf("done %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
