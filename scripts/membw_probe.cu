// Calibration probe (not product code): what this B200 sustains for pure streaming reads, pure
// streaming writes and a copy, with simple 128-bit grid-stride kernels.  Gives the practical
// ceiling the pooling kernels are measured against besides MEASURED_PEAKS.json's copy number.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a scripts/membw_probe.cu -o /tmp/membw_probe && /tmp/membw_probe
#include <cstdio>
#include <cuda_runtime.h>

template <int U>
__global__ void read_k(const uint4 *__restrict__ p, size_t n, uint4 *out)
{
    uint4 acc = make_uint4(0, 0, 0, 0);
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    for (; i + (U - 1) * stride < n; i += U * stride) {
        uint4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u)
            asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                         : "=r"(v[u].x), "=r"(v[u].y), "=r"(v[u].z), "=r"(v[u].w)
                         : "l"(p + i + u * stride));
#pragma unroll
        for (int u = 0; u < U; ++u) { acc.x ^= v[u].x; acc.y ^= v[u].y; acc.z ^= v[u].z; acc.w ^= v[u].w; }
    }
    if (acc.x == 0x12345678u) out[0] = acc;  // never true for the test data; keeps the loads alive
}

__global__ void write_k(uint4 *__restrict__ p, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    const uint4 v = make_uint4(1, 2, 3, 4);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride)
        asm volatile("st.global.cs.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p + i), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void copy_k(const uint4 *__restrict__ a, uint4 *__restrict__ b, size_t n)
{
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) b[i] = a[i];
}

template <typename F>
static float time_ms(F f, int iters)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    for (int i = 0; i < 3; ++i) f();
    cudaEventRecord(e0);
    for (int i = 0; i < iters; ++i) f();
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    return ms / iters;
}

int main()
{
    const size_t bytes = 1ull << 30;  // 1 GiB per buffer (>> 126 MB L2)
    const size_t n = bytes / 16;
    uint4 *a, *b, *out;
    cudaMalloc(&a, bytes);
    cudaMalloc(&b, bytes);
    cudaMalloc(&out, 64);
    cudaMemset(a, 1, bytes);
    cudaMemset(b, 2, bytes);
    for (int ctas_per_sm : {4, 8, 16}) {
        const int grid = 148 * ctas_per_sm;
        float r4 = time_ms([&] { read_k<4><<<grid, 256>>>(a, n, out); }, 20);
        float r8 = time_ms([&] { read_k<8><<<grid, 256>>>(a, n, out); }, 20);
        float w = time_ms([&] { write_k<<<grid, 256>>>(b, n); }, 20);
        float c = time_ms([&] { copy_k<<<grid, 256>>>(a, b, n); }, 20);
        printf("grid=148x%-2d  read(U4) %.0f GB/s  read(U8) %.0f GB/s  write %.0f GB/s  copy(r+w) %.0f GB/s\n", ctas_per_sm,
               bytes / r4 / 1e6, bytes / r8 / 1e6, bytes / w / 1e6, 2.0 * bytes / c / 1e6);
    }
    float m = time_ms([&] { cudaMemcpyAsync(b, a, bytes, cudaMemcpyDeviceToDevice); }, 20);
    printf("cudaMemcpy D2D (r+w) %.0f GB/s\n", 2.0 * bytes / m / 1e6);
    printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
