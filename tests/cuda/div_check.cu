// Test-only: checks gvcnn::div_by_rcp (csrc/common.cuh) against __fdiv_rn for EVERY float32
// dividend and the divisors the kernels use (G + V, group sizes, tie counts).
// Built and run by tests/test_gpu_div.py; prints "mismatches <n>".
#include <cstdio>
#include <cstdlib>

#include "common.cuh"

__global__ void check(const float b, const unsigned stride, unsigned long long *bad, unsigned *first_bad)
{
    const float rcp = __frcp_rn(b);
    const unsigned long long n = (0x100000000ull + stride - 1) / stride;
    for (unsigned long long i = blockIdx.x * (unsigned long long)blockDim.x + threadIdx.x; i < n;
         i += (unsigned long long)gridDim.x * blockDim.x) {
        const unsigned bits = (unsigned)(i * stride);
        const float a = __uint_as_float(bits);
        const float want = __fdiv_rn(a, b);
        const float got = gvcnn::div_by_rcp(a, b, rcp);
        const bool same = (__float_as_uint(want) == __float_as_uint(got)) || (isnan(want) && isnan(got));
        if (!same) {
            if (atomicAdd(bad, 1ull) == 0) *first_bad = bits;
        }
    }
}

int main(int argc, char **argv)
{
    unsigned long long *bad;
    unsigned *first;
    cudaMallocManaged(&bad, sizeof(*bad));
    cudaMallocManaged(&first, sizeof(*first));
    *bad = 0;
    *first = 0;
    // every dividend for the divisors of the shipped configurations, every 61st for the rest
    const int full[] = {1, 2, 3, 5, 7, 12, 13, 14, 16, 18, 20, 22, 24, 28, 36, 84, 96, 4224};
    for (int b : full) check<<<148 * 8, 256>>>((float)b, 1u, bad, first);
    for (int b = 1; b <= 300; ++b) check<<<148 * 8, 256>>>((float)b, 61u, bad, first);
    if (cudaDeviceSynchronize() != cudaSuccess) {
        printf("cuda error %s\n", cudaGetErrorString(cudaGetLastError()));
        return 2;
    }
    printf("mismatches %llu first_bits 0x%08x\n", *bad, *first);
    return *bad ? 1 : 0;
}
