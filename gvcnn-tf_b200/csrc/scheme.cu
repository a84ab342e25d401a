// Scheme / weight glue: the reference's host-side NumPy functions
// group_scheme (nets/model.py:16-25) and group_weight (:28-41) as tiny device
// kernels, so the scheme never has to leave the GPU (the reference round-trips
// it through the host every step, train.py:270-288).  O(rows*G*V) integers.
#include "common.cuh"

namespace gvcnn {

// scheme[row, g, v] = (bins[row, v] == g)           nets/model.py:21-23
__global__ void __launch_bounds__(256) bins_to_scheme_kernel(const int32_t *__restrict__ bins,
                                                             int32_t *__restrict__ scheme,
                                                             const int64_t total, const int V, const int G)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int v = (int)(i % V);
    const int64_t rg = i / V;
    const int g = (int)(rg % G);
    const int64_t row = rg / G;
    scheme[i] = (bins[row * V + v] == g) ? 1 : 0;
}

// bins[row, v] = the g with scheme[row, g, v] != 0; columns with 0 or >1
// nonzeros are counted in status[GVCNN_STATUS_BAD_SCHEME] and get bin 0.
__global__ void __launch_bounds__(256) scheme_to_bins_kernel(const int32_t *__restrict__ scheme,
                                                             int32_t *__restrict__ bins, int32_t *status,
                                                             const int64_t total, const int V, const int G)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int v = (int)(i % V);
    const int64_t row = i / V;
    int found = 0, bin = 0;
    for (int g = 0; g < G; ++g)
        if (scheme[(row * G + g) * V + v] != 0) {
            if (!found) bin = g;
            ++found;
        }
    if (found != 1) {
        if (status) atomicAdd(status + GVCNN_STATUS_BAD_SCHEME, 1);
        bin = 0;
    }
    bins[i] = bin;
}

// weights[row, g] = 1 + #{v: bins[row, v] == g}      nets/model.py:33-39
__global__ void __launch_bounds__(256) group_weight_kernel(const int32_t *__restrict__ bins,
                                                           float *__restrict__ weights, const int64_t total,
                                                           const int V, const int G)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int g = (int)(i % G);
    const int64_t row = i / G;
    int sum = 1;
    for (int v = 0; v < V; ++v) sum += (bins[row * V + v] == g);
    weights[i] = (float)sum;
}

int launch_bins_to_scheme(const int32_t *bins, int32_t *scheme, int rows, int V, int G, cudaStream_t st)
{
    const int64_t total = (int64_t)rows * G * V;
    bins_to_scheme_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(bins, scheme, total, V, G);
    return (int)cudaGetLastError();
}
int launch_scheme_to_bins(const int32_t *scheme, int32_t *bins, int32_t *status, int rows, int V, int G,
                          cudaStream_t st)
{
    const int64_t total = (int64_t)rows * V;
    scheme_to_bins_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(scheme, bins, status, total, V, G);
    return (int)cudaGetLastError();
}
int launch_group_weight(const int32_t *bins, float *weights, int rows, int V, int G, cudaStream_t st)
{
    const int64_t total = (int64_t)rows * G;
    group_weight_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(bins, weights, total, V, G);
    return (int)cudaGetLastError();
}

}  // namespace gvcnn
