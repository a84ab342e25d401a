// Pooling + fusion forward, fast path: persistent, warp-specialised, TMA-fed ring.
//
// Same arithmetic and op order as pool_fwd.cu (nets/model.py:28-102; see there),
// different plumbing.  One CTA per SM stays resident and walks tiles
// t = blockIdx.x, blockIdx.x + gridDim.x, ...  (tile = TD consecutive descriptor
// elements of one shape, all V views).  Roles:
//   * producer warp: prefetches the shape's V bins, ranks the views by
//     (bin, view) with warp shuffles, waits for a free ring slot, publishes the
//     per-tile plan, and then lanes 0..V-1 each fire ONE 1-D bulk async copy
//     (cp.async.bulk, the TMA engine) that lands "their" view's row in the slot
//     at its SORTED position - so consumers read rows k = 0..V-1 in bin order
//     at fixed offsets, with no index indirection;
//   * 8 consumer warps: wait on the slot's full barrier, reduce their 16-byte
//     column group by group in float32 registers, release the slot, then do
//     the division and the streaming store.
// The ring (4 slots x 48 KB at V = 12, fp32) keeps ~150-190 KB of loads in
// flight per SM continuously; there is no per-tile prologue bubble and no
// wave tail.  Used when rows are 16-byte aligned, V <= 32, the reference's own
// group weights are wanted and no per-group descriptors are requested; every
// other case takes the generic kernel in pool_fwd.cu.
#include <cstdlib>

#include "common.cuh"

namespace gvcnn {

constexpr int kRingProducerThreads = 32;

// Per-slot plan written by the producer warp, read by the consumers.
struct __align__(16) RingPlan {
    uint32_t first_mask;  // bit k: the k-th sorted view starts a group
    uint32_t tail_skip;   // empty groups after the last non-empty one
    uint32_t pad[2];
    uint8_t skip[32];     // at a group start k: empty groups between the previous group and this one
};

__device__ __forceinline__ uint32_t bf16x2_max(uint32_t a, uint32_t b)
{
    const __nv_bfloat162 r = __hmax2(*reinterpret_cast<const __nv_bfloat162 *>(&a), *reinterpret_cast<const __nv_bfloat162 *>(&b));
    return *reinterpret_cast<const uint32_t *>(&r);
}
// 0xFFFF in each half where the bf16 values compare equal (IEEE: -0 == +0, NaN != NaN)
__device__ __forceinline__ uint32_t bf16x2_eq_mask(uint32_t a, uint32_t b)
{
    return __heq2_mask(*reinterpret_cast<const __nv_bfloat162 *>(&a), *reinterpret_cast<const __nv_bfloat162 *>(&b));
}

__device__ __forceinline__ void mbar_arrive(uint64_t *bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// V and the consumer count are compile-time: rows sit at immediate offsets, every loop over views is
// fully unrolled, and the only data-dependent control flow left is one uniform branch per view
// ("does this view start a group?").  MINB = CTAs per SM the register budget is sized for.
template <typename T, int POOL, bool MASK, int V, int NCONS, int MINB>
__global__ void __launch_bounds__(NCONS + kRingProducerThreads, MINB)
pool_fuse_fwd_ring_kernel(const ViewPtrs fp, const int64_t f_sb, const int32_t *__restrict__ bins,
                          const int64_t bin_sb, T *__restrict__ S, uint8_t *__restrict__ mask, int32_t *status,
                          const int B, const int64_t D, const int G, const float fill,
                          const int tiles_per_shape, const int num_tiles, const int stages)
{
    constexpr int E = Elem<T>::kVec;
    constexpr int NW = (E + 3) / 4;
    constexpr int P = (V + 7) / 8;
    constexpr int kMaxStages = 8;
    constexpr int TD = NCONS * E;
    constexpr bool kPackedMax = (E == 8) && (POOL == GVCNN_POOL_MAX);  // bf16 max pooling
    constexpr uint32_t kRowStride = (uint32_t)NCONS * 16u;   // bytes between sorted rows in a slot
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
    pdl_wait();                // nothing above touches global memory
    pdl_launch_dependents();

    // tile walk shared by both roles: t = blockIdx.x + it * gridDim.x, kept as (shape b, tile in shape)
    const int step_b = (int)gridDim.x / tiles_per_shape;
    const int step_t = (int)gridDim.x - step_b * tiles_per_shape;
    int b = (int)blockIdx.x / tiles_per_shape;
    int tile = (int)blockIdx.x - b * tiles_per_shape;
    int s = 0;
    uint32_t ph = 0;

    if ((int)threadIdx.x >= NCONS) {
        // ------------------------------------------------------------------ producer warp
        const int lane = threadIdx.x & 31;
        int nb = 0x7fffffff;
        if ((int)blockIdx.x < num_tiles && lane < V) nb = __ldg(bins + (int64_t)b * bin_sb + lane);
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const int64_t d0 = (int64_t)tile * TD;
            const uint32_t row_bytes = (uint32_t)min((int64_t)TD, D - d0) * sizeof(T);
            int bin = nb;
            // next tile's coordinates; prefetch its bins behind this tile's work
            int b_n = b + step_b, tile_n = tile + step_t;
            if (tile_n >= tiles_per_shape) { tile_n -= tiles_per_shape; ++b_n; }
            if (t + (int)gridDim.x < num_tiles && lane < V) nb = __ldg(bins + (int64_t)b_n * bin_sb + lane);
            if (lane < V && (bin < 0 || bin >= G)) {
                if (status && tile == 0) atomicAdd(status + GVCNN_STATUS_BIN_RANGE, 1);
                bin = bin < 0 ? 0 : G - 1;
            }
            int below = 0, same_before = 0;
#pragma unroll
            for (int u = 0; u < V; ++u) {
                const int bu = __shfl_sync(0xffffffffu, bin, u);
                below += (bu < bin);
                same_before += (bu == bin) & (u < lane);
            }
            const int k = below + same_before;
            const bool first = lane < V && same_before == 0;
            const uint32_t fm = __reduce_or_sync(0xffffffffu, first ? (1u << k) : 0u);
            if (lane < V) sorted_bin[k] = bin;
            __syncwarp();
            const int prev = (lane < V && k > 0) ? sorted_bin[k - 1] : -1;
            const int last_bin = sorted_bin[V - 1];
            if (lane == 0) mbar_wait(&empty_bar[s], ph ^ 1u);  // slot drained by all consumer warps
            __syncwarp();
            if (lane < V) plans[s].skip[k] = (uint8_t)(first ? bin - prev - 1 : 0);
            if (lane == 0) {
                plans[s].first_mask = fm;
                plans[s].tail_skip = (uint32_t)(G - 1 - last_bin);
            }
            __syncwarp();
            if (lane == 0) mbar_expect_tx(&full_bar[s], row_bytes * (uint32_t)V);
            __syncwarp();
            if (lane < V)
                bulk_g2s(smem_raw + (size_t)s * kStageBytes + (size_t)k * kRowStride,
                         fp.p[lane] + ((int64_t)b * f_sb + d0) * (int64_t)sizeof(T), row_bytes, &full_bar[s]);
            b = b_n;
            tile = tile_n;
            if (++s == stages) { s = 0; ph ^= 1u; }
        }
        return;
    }

    // ---------------------------------------------------------------------- consumer warps
    const int e0 = threadIdx.x * E;
    const float sumw = (float)(G + V);  // sum_g (1 + n_g): exact in float32 in any order
    const float rcp_sumw = __frcp_rn(sumw);
    for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
        const int64_t d0 = (int64_t)tile * TD;
        const bool active = (int64_t)e0 < D - d0;
        const int64_t out_off = (int64_t)b * D + d0 + e0;
        const unsigned char *col = smem_raw + (size_t)s * kStageBytes + (size_t)threadIdx.x * 16;

        mbar_wait(&full_bar[s], ph);

        float acc[E];
#pragma unroll
        for (int e = 0; e < E; ++e) acc[e] = 0.0f;
        // the plan is the same for every thread: pass it through a warp reduction so the compiler keeps it
        // in uniform registers and the per-view branches below are uniform branches
        const uint32_t fm = __reduce_or_sync(0xffffffffu, plans[s].first_mask);
        const uint32_t tail_skip = __reduce_or_sync(0xffffffffu, plans[s].tail_skip);
        uint32_t skw[(V + 3) / 4];
#pragma unroll
        for (int i = 0; i < (V + 3) / 4; ++i)
            skw[i] = __reduce_or_sync(0xffffffffu, reinterpret_cast<const uint32_t *>(plans[s].skip)[i]);
        if (active && kPackedMax) {
            // bf16 max pooling: the max of bf16 values is exact in bf16, so the running group max stays
            // packed (2 elements per register, max.bf16x2) and is widened to float32 only when a group
            // closes.  Tie bits of an element pair share a register: bits 0..15 / 16..31 = sorted views
            // 0..15 of the even / odd element (a second register set covers views 16..31).
            uint4 raw[V];
#pragma unroll
            for (int k = 0; k < V; ++k) raw[k] = *reinterpret_cast<const uint4 *>(col + k * kRowStride);
            uint32_t m2[4], me2[4], me2b[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) me2[i] = me2b[i] = 0u;
            int cnt = 0;
#pragma unroll
            for (int k = 0; k <= V; ++k) {
                if (k == V || k == 0 || ((fm >> k) & 1u)) {
                    if (k > 0) {
                        if constexpr (MASK) {
#pragma unroll
                            for (int j = 1; j <= k; ++j) {
                                if (j > cnt) break;
                                const uint32_t xw[4] = {raw[k - j].x, raw[k - j].y, raw[k - j].z, raw[k - j].w};
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const uint32_t eq = bf16x2_eq_mask(xw[i], m2[i]);
                                    if (k - j < 16) me2[i] |= eq & (0x00010001u << ((k - j) & 15));
                                    else me2b[i] |= eq & (0x00010001u << ((k - j) & 15));
                                }
                            }
                        }
                        float m[E];
                        Elem<T>::unpack(make_uint4(m2[0], m2[1], m2[2], m2[3]), m);
                        const float w = (float)(1 + cnt);
#pragma unroll
                        for (int e = 0; e < E; ++e) acc[e] = __fadd_rn(acc[e], __fmul_rn(w, m[e]));
                    }
                    if (fill != 0.0f) {
                        const uint32_t nskip = (k == V) ? tail_skip : ((skw[(k < V ? k : 0) >> 2] >> (8 * (k & 3))) & 0xffu);
#pragma unroll 1
                        for (uint32_t q = 0; q < nskip; ++q) {
#pragma unroll
                            for (int e = 0; e < E; ++e) acc[e] = __fadd_rn(acc[e], fill);
                        }
                    }
                    if (k < V) {
                        m2[0] = raw[k].x; m2[1] = raw[k].y; m2[2] = raw[k].z; m2[3] = raw[k].w;
                        cnt = 1;
                    }
                } else {
                    const uint4 r = raw[k < V ? k : 0];
                    m2[0] = bf16x2_max(m2[0], r.x); m2[1] = bf16x2_max(m2[1], r.y);
                    m2[2] = bf16x2_max(m2[2], r.z); m2[3] = bf16x2_max(m2[3], r.w);
                    ++cnt;
                }
            }
            if constexpr (MASK) {
                // byte planes: plane p, element e -> bits 8(p&1).. of the half of me2/me2b[e >> 1]
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const uint32_t *src = (p < 2) ? me2 : me2b;
                    const uint32_t sel = (p & 1) ? 0x7531u : 0x6420u;
                    const uint32_t w0 = __byte_perm(src[0], src[1], sel);
                    const uint32_t w1 = __byte_perm(src[2], src[3], sel);
                    *reinterpret_cast<uint2 *>(mask + ((int64_t)p * B) * D + out_off) = make_uint2(w0, w1);
                }
            }
        } else if (active) {
            // all V rows of this thread's column, in bin order, fetched in one batch
            uint4 raw[V];
#pragma unroll
            for (int k = 0; k < V; ++k) raw[k] = *reinterpret_cast<const uint4 *>(col + k * kRowStride);

            float m[E];
            uint32_t me[E];  // tie bits per element: bit k <=> sorted view k attains its group's max
#pragma unroll
            for (int e = 0; e < E; ++e) me[e] = 0u;
            int cnt = 0;
#pragma unroll
            for (int k = 0; k <= V; ++k) {
                if (k == V || k == 0 || ((fm >> k) & 1u)) {  // uniform: a group ends / starts here
                    if (k > 0) {                             // close the previous group
                        if constexpr (MASK && POOL == GVCNN_POOL_MAX) {
                            // its members are the cnt rows before k: compare each with the group max
#pragma unroll
                            for (int j = 1; j <= k; ++j) {
                                if (j > cnt) break;
                                float x[E];
                                Elem<T>::unpack(raw[k - j], x);
#pragma unroll
                                for (int e = 0; e < E; ++e)
                                    if (x[e] == m[e]) me[e] |= 1u << (k - j);
                            }
                        }
                        const float w = (float)(1 + cnt);    // acc += w_g * P_g
#pragma unroll
                        for (int e = 0; e < E; ++e) {
                            if (POOL == GVCNN_POOL_MEAN) m[e] = __fdiv_rn(m[e], (float)cnt);
                            acc[e] = __fadd_rn(acc[e], __fmul_rn(w, m[e]));
                        }
                    }
                    if (fill != 0.0f) {  // empty groups in between / after: w = 1, P = fill
                        const uint32_t nskip = (k == V) ? tail_skip : ((skw[(k < V ? k : 0) >> 2] >> (8 * (k & 3))) & 0xffu);
#pragma unroll 1
                        for (uint32_t q = 0; q < nskip; ++q) {
#pragma unroll
                            for (int e = 0; e < E; ++e) acc[e] = __fadd_rn(acc[e], fill);
                        }
                    }
                    if (k < V) {
                        Elem<T>::unpack(raw[k], m);
                        cnt = 1;
                    }
                } else {
                    float x[E];
                    Elem<T>::unpack(raw[k < V ? k : 0], x);
#pragma unroll
                    for (int e = 0; e < E; ++e)
                        m[e] = (POOL == GVCNN_POOL_MAX) ? fmaxf(m[e], x[e]) : __fadd_rn(m[e], x[e]);
                    ++cnt;
                }
            }
            if constexpr (MASK && POOL == GVCNN_POOL_MAX) {
                // transpose to byte planes: byte e of plane word p = bits 8p..8p+7 of me[e]
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    uint32_t wd[NW];
#pragma unroll
                    for (int i = 0; i < NW; ++i) {
                        wd[i] = 0u;
#pragma unroll
                        for (int e = 4 * i; e < 4 * i + 4; ++e) wd[i] |= ((me[e] >> (8 * p)) & 0xffu) << (8 * (e & 3));
                    }
                    uint8_t *mp = mask + ((int64_t)p * B) * D + out_off;
                    if constexpr (E == 8) *reinterpret_cast<uint2 *>(mp) = make_uint2(wd[0], wd[1]);
                    else *reinterpret_cast<uint32_t *>(mp) = wd[0];
                }
            }
        }
        // every lane of the warp is done reading the slot: hand it back to the producer
        __syncwarp();
        if ((threadIdx.x & 31) == 0) mbar_arrive(&empty_bar[s]);
        if (active) {
#pragma unroll
            for (int e = 0; e < E; ++e) acc[e] = div_by_rcp(acc[e], sumw, rcp_sumw);
            stg_stream_16(S + out_off, Elem<T>::pack(acc));
        }
        b += step_b;
        tile += step_t;
        if (tile >= tiles_per_shape) { tile -= tiles_per_shape; ++b; }
        if (++s == stages) { s = 0; ph ^= 1u; }
    }
}

static int ring_sm_count()
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

template <typename T, int V, int NCONS, int MINB>
static int launch_ring_v(const ViewPtrs &fp, int64_t f_sb, const int32_t *bins, int64_t bin_sb, void *S,
                         uint8_t *mask, int32_t *status, int B, int64_t D, int G, int pool, float fill,
                         cudaStream_t st)
{
    constexpr int E = Elem<T>::kVec;
    constexpr size_t stage_bytes = (size_t)V * NCONS * 16;
    // shared memory an SM can give to MINB co-resident CTAs of this kernel (1 KB reserved + ~1.3 KB static each)
    int stages = (int)(((226 * 1024) / MINB - 3 * 1024) / stage_bytes);
    if (stages > 8) stages = 8;
    if (const char *env = getenv("GVCNN_RING_STAGES")) {  // tuning knob for A/B runs
        const int want = atoi(env);
        if (want >= 2 && want <= stages) stages = want;
    }
    if (stages < 2) return -1000;
    const int64_t td = (int64_t)NCONS * E;
    const int64_t tps = (D + td - 1) / td;
    const int64_t tiles = (int64_t)B * tps;
    if (tiles > 0x7fffffffLL) return GVCNN_E_BAD_ARG;
    const size_t smem = stage_bytes * stages;
    const int64_t max_grid = (int64_t)ring_sm_count() * MINB;
    const int grid = (int)(tiles < max_grid ? tiles : max_grid);
    const bool want_mask = (mask != nullptr) && pool == GVCNN_POOL_MAX;
    cudaError_t err = cudaSuccess;
#define GVCNN_LAUNCH_RING(POOL_, MASK_)                                                                      \
    do {                                                                                                     \
        auto kern = pool_fuse_fwd_ring_kernel<T, POOL_, MASK_, V, NCONS, MINB>;                              \
        err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);            \
        if (err == cudaSuccess)                                                                              \
            err = launch_pdl(kern, dim3(grid), dim3(NCONS + kRingProducerThreads), smem, st, fp, f_sb, bins,  \
                             bin_sb, static_cast<T *>(S), mask, status, B, D, G, fill, (int)tps,             \
                             (int)tiles, stages);                                                            \
    } while (0)
    if (pool == GVCNN_POOL_MAX) {
        if (want_mask) GVCNN_LAUNCH_RING(GVCNN_POOL_MAX, true); else GVCNN_LAUNCH_RING(GVCNN_POOL_MAX, false);
    } else {
        GVCNN_LAUNCH_RING(GVCNN_POOL_MEAN, false);
    }
#undef GVCNN_LAUNCH_RING
    if (err != cudaSuccess) return (int)err;
    return (int)cudaGetLastError();
}

template <typename T>
static int launch_ring_t(const ViewPtrs &fp, int64_t f_sb, const int32_t *bins, int64_t bin_sb, void *S,
                         uint8_t *mask, int32_t *status, int B, int V, int64_t D, int G, int pool, float fill,
                         cudaStream_t st)
{
    // the view counts of the reference's configurations (train.py:96 default 6; BASELINE sweep 6/12/20)
    // plus the other common multi-view rigs (4, 8, 16)
    if (D < 256 * Elem<T>::kVec) return -1000;  // tiles narrower than one consumer row: generic kernel
    switch (V) {
    case 4: return launch_ring_v<T, 4, 256, 2>(fp, f_sb, bins, bin_sb, S, mask, status, B, D, G, pool, fill, st);
    case 6: return launch_ring_v<T, 6, 256, 2>(fp, f_sb, bins, bin_sb, S, mask, status, B, D, G, pool, fill, st);
    case 8: return launch_ring_v<T, 8, 256, 2>(fp, f_sb, bins, bin_sb, S, mask, status, B, D, G, pool, fill, st);  // GVCNN paper: 8 / 12 views
    case 12: return launch_ring_v<T, 12, 256, 2>(fp, f_sb, bins, bin_sb, S, mask, status, B, D, G, pool, fill, st);
    case 16: return launch_ring_v<T, 16, 256, 1>(fp, f_sb, bins, bin_sb, S, mask, status, B, D, G, pool, fill, st);
    case 20: return launch_ring_v<T, 20, 256, 1>(fp, f_sb, bins, bin_sb, S, mask, status, B, D, G, pool, fill, st);
    default: return -1000;
    }
}

// returns -1000 when this fast path does not apply
int launch_pool_fuse_fwd_ring(const ViewPtrs &fp, int64_t f_sb, const int32_t *bins, int64_t bin_sb, void *S,
                              uint8_t *mask, int32_t *status, int B, int V, int64_t D, int G, int pool,
                              float fill, int dtype, cudaStream_t st)
{
    if (V > 32 || G > 255) return -1000;
    if (dtype == GVCNN_F32) {
        if (D % 4) return -1000;
        return launch_ring_t<float>(fp, f_sb, bins, bin_sb, S, mask, status, B, V, D, G, pool, fill, st);
    }
    if (D % 8) return -1000;
    return launch_ring_t<__nv_bfloat16>(fp, f_sb, bins, bin_sb, S, mask, status, B, V, D, G, pool, fill, st);
}

}  // namespace gvcnn
