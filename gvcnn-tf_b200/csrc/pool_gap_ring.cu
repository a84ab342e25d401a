// Pooling + fusion forward with the global average pooling that follows it folded in
// (nets/model.py:154-163: view_pooling -> group_fusion -> GlobalAveragePooling2D).
//
// The real descriptors are spatial maps [N, h, w, C] per view (D = h*w*C, channel-last, nets/model.py:149) and
// the only consumer of the fused map S is the GAP in front of the classifier.  This kernel never writes S:
// a work unit is (shape b, channel block cb, position split ps); it streams that channel block of its
// positions through the same TMA ring as pool_fwd_ring.cu (same producer, same ring_consume_tile, so the
// per-position arithmetic is bit-identical), adds the per-position results in position order in registers and
// writes one float32 partial row; gap_finish_kernel adds the splits in ascending order and divides by h*w.
// Saves the write of S ([N, h*w*C], 1/(V+1) of the forward traffic) and, with gvcnn_pool_fuse_gap_bwd, the
// read of dS in the backward.  Max pooling does not commute with the GAP, so the pooling itself is unchanged.
#include "ring_common.cuh"

namespace gvcnn {

template <typename T, int POOL, bool MASK, int V, int NCONS, int MINB>
__global__ void __launch_bounds__(NCONS + kRingProducerThreads, MINB)
pool_fuse_gap_fwd_kernel(const ViewPtrs fp, const int64_t f_sb, const int32_t *__restrict__ bins,
                         const int64_t bin_sb, float *__restrict__ partial, uint8_t *__restrict__ mask,
                         int32_t *status, const int B, const int C, const int HW, const int G, const float fill,
                         const int CB, const int NSPLIT, const int PPS, const int num_units, const int stages)
{
    constexpr int E = Elem<T>::kVec;
    constexpr int kMaxStages = 8;
    constexpr int TD = NCONS * E;
    constexpr uint32_t kRowStride = (uint32_t)NCONS * 16u;
    constexpr uint32_t kStageBytes = kRowStride * (uint32_t)V;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ RingPlan plans[kMaxStages];
    __shared__ int32_t sorted_bin[32];
    __shared__ __align__(8) uint64_t full_bar[kMaxStages];
    __shared__ __align__(8) uint64_t empty_bar[kMaxStages];

    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], (uint32_t)(NCONS >> 5));
        }
        mbar_fence_init();
    }
    __syncthreads();
    pdl_wait();
    pdl_launch_dependents();

    const int64_t D = (int64_t)HW * C;
    int s = 0;
    uint32_t ph = 0;

    if ((int)threadIdx.x >= NCONS) {
        // ------------------------------------------------------------------ producer warp
        const int lane = threadIdx.x & 31;
        for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
            const int ps = u % NSPLIT;
            const int cb = (u / NSPLIT) % CB;
            const int b = u / (NSPLIT * CB);
            int bin = 0x7fffffff;
            if (lane < V) {
                bin = __ldg(bins + (int64_t)b * bin_sb + lane);
                if (bin < 0 || bin >= G) {
                    if (status && cb == 0 && ps == 0) atomicAdd(status + GVCNN_STATUS_BIN_RANGE, 1);
                    bin = bin < 0 ? 0 : G - 1;
                }
            }
            int below = 0, same_before = 0;
#pragma unroll
            for (int q = 0; q < V; ++q) {
                const int bu = __shfl_sync(0xffffffffu, bin, q);
                below += (bu < bin);
                same_before += (bu == bin) & (q < lane);
            }
            const int k = below + same_before;
            const bool first = lane < V && same_before == 0;
            const uint32_t fm = __reduce_or_sync(0xffffffffu, first ? (1u << k) : 0u);
            __syncwarp();
            if (lane < V) sorted_bin[k] = bin;
            __syncwarp();
            const int prev = (lane < V && k > 0) ? sorted_bin[k - 1] : -1;
            const int last_bin = sorted_bin[V - 1];
            const int p1 = min(HW, (ps + 1) * PPS);
            for (int p = ps * PPS; p < p1; ++p) {
                const int64_t d0 = (int64_t)p * C + (int64_t)cb * TD;
                if (lane == 0) mbar_wait(&empty_bar[s], ph ^ 1u);
                __syncwarp();
                if (lane < V) plans[s].skip[k] = (uint8_t)(first ? bin - prev - 1 : 0);
                if (lane == 0) {
                    plans[s].first_mask = fm;
                    plans[s].tail_skip = (uint32_t)(G - 1 - last_bin);
                }
                __syncwarp();
                if (lane == 0) mbar_expect_tx(&full_bar[s], kRowStride * (uint32_t)V);
                __syncwarp();
                if (lane < V)
                    bulk_g2s(smem_raw + (size_t)s * kStageBytes + (size_t)k * kRowStride,
                             fp.p[lane] + ((int64_t)b * f_sb + d0) * (int64_t)sizeof(T), kRowStride, &full_bar[s]);
                if (++s == stages) { s = 0; ph ^= 1u; }
            }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int e0 = threadIdx.x * E;
    const float sumw = (float)(G + V);
    const float rcp_sumw = __frcp_rn(sumw);
    for (int u = blockIdx.x; u < num_units; u += gridDim.x) {
        const int ps = u % NSPLIT;
        const int cb = (u / NSPLIT) % CB;
        const int b = u / (NSPLIT * CB);
        float gacc[E];
#pragma unroll
        for (int e = 0; e < E; ++e) gacc[e] = 0.0f;
        const int p1 = min(HW, (ps + 1) * PPS);
        for (int p = ps * PPS; p < p1; ++p) {
            const int64_t out_off = (int64_t)b * D + (int64_t)p * C + (int64_t)cb * TD + e0;
            const unsigned char *col = smem_raw + (size_t)s * kStageBytes + (size_t)threadIdx.x * 16;
            mbar_wait(&full_bar[s], ph);
            float acc[E];
            ring_consume_tile<T, POOL, MASK, V, kRowStride>(col, plans[s], fill, true, mask, B, D, out_off, acc);
            __syncwarp();
            if ((threadIdx.x & 31) == 0) mbar_arrive(&empty_bar[s]);
            div_vec_by_rcp(acc, sumw, rcp_sumw);  // S at this position, then the running sum over positions
#pragma unroll
            for (int e = 0; e < E; ++e) gacc[e] = __fadd_rn(gacc[e], acc[e]);
            if (++s == stages) { s = 0; ph ^= 1u; }
        }
        float *dst = partial + ((int64_t)b * NSPLIT + ps) * C + (int64_t)cb * TD + e0;
#pragma unroll
        for (int q = 0; q < E; q += 4)
            *reinterpret_cast<float4 *>(dst + q) = make_float4(gacc[q], gacc[q + 1], gacc[q + 2], gacc[q + 3]);
    }
}

// out[b, c] = (sum over splits, ascending, of partial[b, ps, c]) / HW
template <typename T>
__global__ void __launch_bounds__(256) gap_finish_kernel(const float *__restrict__ partial, T *__restrict__ out,
                                                        const int64_t total, const int C, const int NSPLIT,
                                                        const float hw)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (i >= total) return;
    const int64_t b = i / C;
    const int c = (int)(i - b * C);
    float sum = partial[(b * NSPLIT) * C + c];
    for (int ps = 1; ps < NSPLIT; ++ps) sum = __fadd_rn(sum, partial[(b * NSPLIT + ps) * C + c]);
    out[i] = Elem<T>::from_float(__fdiv_rn(sum, hw));
}

static int gap_sm_count()
{
    static int cached = 0;
    if (cached == 0) {
        int dev = 0, n = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) {
            cudaGetLastError();
            return 148;
        }
        cached = n;
    }
    return cached;
}

// how the positions are split: enough units to fill the machine twice, never more splits than positions
void gap_split(int B, int CB, int HW, int &nsplit, int &pps)
{
    const int64_t base = (int64_t)B * CB;
    int64_t want = (4LL * gap_sm_count() + base - 1) / base;
    if (want < 1) want = 1;
    if (want > HW) want = HW;
    pps = (int)((HW + want - 1) / want);
    nsplit = (HW + pps - 1) / pps;
}

template <typename T, int V, int MINB>
static int launch_gap_v(const ViewPtrs &fp, int64_t f_sb, const int32_t *bins, int64_t bin_sb, void *out,
                        uint8_t *mask, int32_t *status, float *partial, int B, int HW, int C, int G, int pool,
                        float fill, cudaStream_t st)
{
    constexpr int NCONS = 256;
    constexpr int E = Elem<T>::kVec;
    constexpr int TD = NCONS * E;
    constexpr size_t stage_bytes = (size_t)V * NCONS * 16;
    if (C % TD != 0) return -1000;
    int stages = (int)(((226 * 1024) / MINB - 3 * 1024) / stage_bytes);
    if (stages > 8) stages = 8;
    if (stages < 2) return -1000;
    const int CB = C / TD;
    int nsplit, pps;
    gap_split(B, CB, HW, nsplit, pps);
    const int64_t units = (int64_t)B * CB * nsplit;
    if (units > 0x7fffffffLL) return GVCNN_E_BAD_ARG;
    const size_t smem = stage_bytes * stages;
    const int64_t max_grid = (int64_t)gap_sm_count() * MINB;
    const int grid = (int)(units < max_grid ? units : max_grid);
    const bool want_mask = (mask != nullptr) && pool == GVCNN_POOL_MAX;
    cudaError_t err = cudaSuccess;
#define GVCNN_LAUNCH_GAP(POOL_, MASK_)                                                                        \
    do {                                                                                                      \
        auto kern = pool_fuse_gap_fwd_kernel<T, POOL_, MASK_, V, NCONS, MINB>;                                \
        err = ensure_dyn_smem<pool_fuse_gap_fwd_kernel<T, POOL_, MASK_, V, NCONS, MINB>>((int)smem);          \
        if (err == cudaSuccess)                                                                               \
            err = launch_pdl(kern, dim3(grid), dim3(NCONS + kRingProducerThreads), smem, st, fp, f_sb, bins,  \
                             bin_sb, partial, mask, status, B, C, HW, G, fill, CB, nsplit, pps, (int)units,   \
                             stages);                                                                         \
    } while (0)
    if (pool == GVCNN_POOL_MAX) {
        if (want_mask) GVCNN_LAUNCH_GAP(GVCNN_POOL_MAX, true); else GVCNN_LAUNCH_GAP(GVCNN_POOL_MAX, false);
    } else {
        GVCNN_LAUNCH_GAP(GVCNN_POOL_MEAN, false);
    }
#undef GVCNN_LAUNCH_GAP
    if (err != cudaSuccess) return (int)err;
    const int64_t total = (int64_t)B * C;
    gap_finish_kernel<T><<<(unsigned)((total + 255) / 256), 256, 0, st>>>(partial, static_cast<T *>(out), total, C,
                                                                         nsplit, (float)HW);
    return (int)cudaGetLastError();
}

size_t gap_workspace_bytes(int B, int C, int HW, int dtype)
{
    const int td = 256 * (dtype == GVCNN_F32 ? 4 : 8);
    if (C % td != 0) return 0;
    int nsplit, pps;
    gap_split(B, C / td, HW, nsplit, pps);
    return (size_t)B * nsplit * C * sizeof(float);
}

// returns -1000 when the shapes are not supported (V not instantiated, C not a multiple of the tile, G > 255)
int launch_pool_fuse_gap_fwd(const ViewPtrs &fp, int64_t f_sb, const int32_t *bins, int64_t bin_sb, void *out,
                             uint8_t *mask, int32_t *status, float *partial, int B, int V, int HW, int C, int G,
                             int pool, float fill, int dtype, cudaStream_t st)
{
    if (G > 255) return -1000;
#define GVCNN_GAP_CASE(T_)                                                                                    \
    switch (V) {                                                                                              \
    case 4: return launch_gap_v<T_, 4, 2>(fp, f_sb, bins, bin_sb, out, mask, status, partial, B, HW, C, G, pool, fill, st); \
    case 6: return launch_gap_v<T_, 6, 2>(fp, f_sb, bins, bin_sb, out, mask, status, partial, B, HW, C, G, pool, fill, st); \
    case 8: return launch_gap_v<T_, 8, 2>(fp, f_sb, bins, bin_sb, out, mask, status, partial, B, HW, C, G, pool, fill, st); \
    case 12: return launch_gap_v<T_, 12, 2>(fp, f_sb, bins, bin_sb, out, mask, status, partial, B, HW, C, G, pool, fill, st); \
    case 16: return launch_gap_v<T_, 16, 1>(fp, f_sb, bins, bin_sb, out, mask, status, partial, B, HW, C, G, pool, fill, st); \
    case 20: return launch_gap_v<T_, 20, 1>(fp, f_sb, bins, bin_sb, out, mask, status, partial, B, HW, C, G, pool, fill, st); \
    default: return -1000;                                                                                    \
    }
    if (dtype == GVCNN_F32) { GVCNN_GAP_CASE(float) }
    GVCNN_GAP_CASE(__nv_bfloat16)
#undef GVCNN_GAP_CASE
}

}  // namespace gvcnn
