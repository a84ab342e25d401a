// Pooling + fusion backward, fast path: V compile-time, views walked in bin order with one uniform
// branch per view, tie bits held as one 32-bit word per descriptor element.
//
// Same arithmetic and op order as pool_bwd.cu (TF autodiff of nets/model.py:62-100; see there):
//     g0 = dS / (G + V);  g1 = g0 * w_g;  max: dF_v = (1 / num_selected) * g1 | 0;  mean: dF_v = g1 / n_g.
// Plumbing: one CTA per tile of NT*E descriptor elements of one shape.  Warp 0 ranks the views with
// shuffles and publishes, per sorted position k, the destination row pointer of that view and, per
// group start, the group's member mask; meanwhile every thread has its 16 bytes of dS and its
// ceil(V/8) tie-mask words in flight.  The per-group work (num_selected = popc(ties & members), the
// reciprocal from a shared table, two multiplies) is shared by the group's members; per view only a bit
// test + select per element and one 128-bit streaming store remain.  Used when rows are 16-byte
// aligned, V is one of the instantiated view counts and the reference's own group weights are wanted.
#include "common.cuh"

namespace gvcnn {

struct __align__(16) BwdPlan {
    unsigned long long rowptr[32];  // k -> byte address of row (b, view order[k]) at this tile's d0
    uint32_t seg[32];               // at a group start k: bitmask of the group's sorted positions
    float gw[32];                   // caller-supplied weights only: weight of the group of sorted view k
    float sumw;                     // caller-supplied weights only: sum of all G weights (left to right)
    uint32_t first_mask;            // bit k: sorted view k starts a group
};

// prmt.b32, generic mode: a selector nibble with bit 3 set replicates the top bit of the selected byte over the
// result byte (PTX ISA, prmt) - 0x00 or 0xff per byte
__device__ __forceinline__ uint32_t prmt_sign(const uint32_t a, const uint32_t sel)
{
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(0u), "r"(sel));
    return d;
}

// GAP: dS is the gradient of the global average pool that follows the fusion, [B, C] (D = HW * C,
// channel-last): dS[b, p, c] = dOut[b, c] / HW (tf.reduce_mean's gradient), never materialised.
template <typename T, int POOL, int V, int NT, bool GAP, bool WTS, int UPC>
__global__ void __launch_bounds__(NT)
pool_fuse_bwd_fast_kernel(const T *__restrict__ dS, const int32_t *__restrict__ bins, const int64_t bin_sb,
                          const uint8_t *__restrict__ mask, const float *__restrict__ weights, const int64_t w_sb,
                          const ViewPtrs gp, const int64_t g_sb, int32_t *status,
                          const int B, const int64_t D, const int G, const int tiles_per_shape, const int C,
                          const int HW, const int num_units)
{
    constexpr int E = Elem<T>::kVec;
    constexpr int NW = (E + 3) / 4;
    constexpr int P = (V + 7) / 8;
    constexpr int TD = NT * E;
    static_assert(UPC * 32 <= NT, "one planning warp per unit");
    // A CTA handles UPC consecutive (shape, tile) units: their loads are all in flight before the one barrier, and
    // warp i plans unit i, so the per-CTA costs (launch slot, barrier, reciprocal table) are paid once per UPC tiles.
    __shared__ BwdPlan plans[UPC];
    __shared__ float rcp_tab[V + 1];  // rcp_tab[n] = 1 / n, IEEE division

    const int u0 = blockIdx.x * UPC;
    const int e0 = threadIdx.x * E;

    pdl_wait();
    pdl_launch_dependents();
    // ---- loads first: dS and the tie-mask planes of this thread's elements
    uint4 raws[UPC];
    uint32_t pwds[UPC][P][NW];
    bool actives[UPC];
#pragma unroll
    for (int i = 0; i < UPC; ++i) {
        const int u = u0 + i;
        const int b = u / tiles_per_shape;
        const int64_t d0 = (int64_t)(u - b * tiles_per_shape) * TD;
        actives[i] = u < num_units && (int64_t)e0 < D - d0;
        const int64_t off = (int64_t)b * D + d0 + e0;
        raws[i] = make_uint4(0u, 0u, 0u, 0u);
        if (actives[i]) {
            if constexpr (GAP) raws[i] = *reinterpret_cast<const uint4 *>(dS + (int64_t)b * C + (int)((d0 + e0) % C));
            else raws[i] = ldg_stream_16(dS + off);
            if constexpr (POOL == GVCNN_POOL_MAX) {
#pragma unroll
                for (int p = 0; p < P; ++p) {
                    const uint8_t *mp = mask + ((int64_t)p * B) * D + off;
                    if constexpr (E == 8) {
                        const uint2 w2 = *reinterpret_cast<const uint2 *>(mp);
                        pwds[i][p][0] = w2.x;
                        pwds[i][p][NW - 1] = w2.y;
                    } else {
                        pwds[i][p][0] = *reinterpret_cast<const uint32_t *>(mp);
                    }
                }
            }
        }
    }
    if (threadIdx.x <= V && threadIdx.x > 0) rcp_tab[threadIdx.x] = __fdiv_rn(1.0f, (float)threadIdx.x);

    // ---- plans: warp i ranks the views of unit i by (bin, view)
    if ((int)(threadIdx.x >> 5) < UPC && u0 + (int)(threadIdx.x >> 5) < num_units) {
        BwdPlan &plan = plans[threadIdx.x >> 5];
        const int un = u0 + (int)(threadIdx.x >> 5);
        const int b = un / tiles_per_shape;
        const int tile = un - b * tiles_per_shape;
        const int64_t d0 = (int64_t)tile * TD;
        const int lane = threadIdx.x & 31;
        int bin = 0x7fffffff;
        if (lane < V) {
            bin = __ldg(bins + (int64_t)b * bin_sb + lane);
            if (bin < 0 || bin >= G) {
                if (status && tile == 0) atomicAdd(status + GVCNN_STATUS_BIN_RANGE, 1);
                bin = bin < 0 ? 0 : G - 1;
            }
        }
        int below = 0, same_before = 0;
        uint32_t peers = 0u;  // lanes (views) in my group
#pragma unroll
        for (int u = 0; u < V; ++u) {
            const int bu = __shfl_sync(0xffffffffu, bin, u);
            below += (bu < bin);
            same_before += (bu == bin) & (u < lane);
            peers |= (bu == bin) ? (1u << u) : 0u;
        }
        const int k = below + same_before;
        const bool first = lane < V && same_before == 0;
        const uint32_t fm = __reduce_or_sync(0xffffffffu, first ? (1u << k) : 0u);
        if (lane < V) {
            // the group occupies sorted positions k - same_before .. + popc(peers) - 1
            const int n = __popc(peers);
            const int start = k - same_before;
            plan.seg[k] = (n >= 32 ? 0xffffffffu : ((1u << n) - 1u)) << start;
            plan.rowptr[k] = (unsigned long long)(gp.p[lane] + ((int64_t)b * g_sb + d0) * (int64_t)sizeof(T));
            if constexpr (WTS) plan.gw[k] = __ldg(weights + (int64_t)b * w_sb + bin);
        }
        if (lane == 0) {
            plan.first_mask = fm;
            if constexpr (WTS) {
                float sw = 0.0f;
                for (int g = 0; g < G; ++g) sw = __fadd_rn(sw, __ldg(weights + (int64_t)b * w_sb + g));
                plan.sumw = sw;
            }
        }
    }
    __syncthreads();

    const uint32_t thread_off = (uint32_t)e0 * (uint32_t)sizeof(T);
#pragma unroll
    for (int ui = 0; ui < UPC; ++ui) {
    if (!actives[ui]) continue;
    const BwdPlan &plan = plans[ui];
    const uint4 raw = raws[ui];
    uint32_t (&pwd)[P][NW] = pwds[ui];
    constexpr bool wts = WTS;  // caller-supplied group weights: compiled out of the default instantiation
    const float sumw = wts ? plan.sumw : (float)(G + V);
    const float rcp_sumw = __frcp_rn((float)(G + V));
    float t[E];
    Elem<T>::unpack(raw, t);
    if constexpr (GAP) {
#pragma unroll
        for (int e = 0; e < E; ++e) t[e] = __fdiv_rn(t[e], (float)HW);  // gradient of the mean over positions
    }
    if constexpr (wts) {  // g0 = dS / sum_w
#pragma unroll
        for (int e = 0; e < E; ++e) t[e] = __fdiv_rn(t[e], sumw);
    } else {
        div_vec_by_rcp(t, sumw, rcp_sumw);  // one range test for the vector, IEEE fall-back out of line
    }

    const uint32_t fm = plan.first_mask;

    if constexpr (E == 8 && POOL == GVCNN_POOL_MAX) {
        // bf16: tie bits of an element pair share a register (bits 0..15 / 16..31 = sorted views 0..15
        // of the even / odd element; me2b covers views 16..31); the group's value is rounded to bf16 and
        // packed once per group, each view then only masks the packed pairs.
        uint32_t me2[4], me2b[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t sel = (i & 1) ? 0x7362u : 0x5140u;
            me2[i] = __byte_perm(pwd[0][i >> 1], P > 1 ? pwd[P > 1 ? 1 : 0][i >> 1] : 0u, sel);
            me2b[i] = P > 2 ? __byte_perm(pwd[P > 2 ? 2 : 0][i >> 1], P > 3 ? pwd[P > 3 ? 3 : 0][i >> 1] : 0u, sel) : 0u;
        }
        uint32_t val2[4];
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (k == 0 || ((fm >> k) & 1u)) {  // uniform: a group starts here
                const uint32_t seg = plan.seg[k];
                const int n = __popc(seg);
                const float w = wts ? plan.gw[k] : (float)(1 + n);
                float val[E];
                if (n == 1) {  // uniform: a group of one view - num_selected is 1 wherever the value is kept
#pragma unroll
                    for (int e = 0; e < E; ++e) val[e] = __fmul_rn(t[e], w);  // (1 / 1) * g1 = g1 exactly
                } else {
#pragma unroll
                    for (int e = 0; e < E; ++e) {
                        const uint32_t half = (e & 1) ? 0xffff0000u : 0x0000ffffu;
                        const uint32_t lo = (seg & 0xffffu) * 0x00010001u, hi = (seg >> 16) * 0x00010001u;
                        const int nsel = __popc(me2[e >> 1] & lo & half) + (V > 16 ? __popc(me2b[e >> 1] & hi & half) : 0);
                        val[e] = __fmul_rn(rcp_tab[nsel], __fmul_rn(t[e], w));
                    }
                }
                const uint4 pk = Elem<T>::pack(val);
                val2[0] = pk.x; val2[1] = pk.y; val2[2] = pk.z; val2[3] = pk.w;
            }
            // view k keeps the value where its tie bit is set: bit (k & 7) of element e's byte in plane k >> 3.  Shifted
            // to the byte's top bit, one PRMT in sign-replicate mode turns the two bytes of an element pair into the
            // 0xffff / 0x0000 halves that mask the packed pair (shift per word + PRMT + AND per pair).
            uint4 o;
            uint32_t *ow = reinterpret_cast<uint32_t *>(&o);
            const uint32_t sh0 = pwd[k >> 3][0] << (7 - (k & 7)), sh1 = pwd[k >> 3][NW - 1] << (7 - (k & 7));
#pragma unroll
            for (int i = 0; i < 4; ++i)
                ow[i] = val2[i] & prmt_sign((i >> 1) ? sh1 : sh0, (i & 1) ? 0xBBAAu : 0x9988u);
            stg_stream_16(reinterpret_cast<char *>(plan.rowptr[k]) + thread_off, o);
        }
    } else if constexpr (POOL == GVCNN_POOL_MEAN) {
        // mean: every member of a group receives g1 / n - one value per group, one store per view.  Nothing here needs
        // a compile-time k, so the walk is a rolled loop (unrolled it was V copies of the group code: 6100 SASS
        // instructions at V = 20).  g1 / n: n is uniform; for a power of two 1/n is exact and the product is the
        // correctly rounded quotient (n == 1 included), otherwise the exact division by the reciprocal.
        uint4 packed = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll 1
        for (int k = 0; k < V; ++k) {
            if ((fm >> k) & 1u) {  // uniform: a group starts here (bit 0 is always set)
                const int n = __popc(plan.seg[k]);
                const float w = wts ? plan.gw[k] : (float)(1 + n);
                const float rn = rcp_tab[n];
                float val[E];
#pragma unroll
                for (int e = 0; e < E; ++e) val[e] = __fmul_rn(t[e], w);
                if ((n & (n - 1)) == 0) {  // uniform branch
#pragma unroll
                    for (int e = 0; e < E; ++e) val[e] = __fmul_rn(val[e], rn);
                } else {
                    div_vec_by_rcp(val, (float)n, rn);
                }
                packed = Elem<T>::pack(val);
            }
            stg_stream_16(reinterpret_cast<char *>(plan.rowptr[k]) + thread_off, packed);
        }
    } else {
        // max pooling, float32: tie bits per element: bit k <=> sorted view k attains its group's max
        uint32_t me[E];
#pragma unroll
        for (int e = 0; e < E; ++e) {
            me[e] = 0u;
#pragma unroll
            for (int p = 0; p < P; ++p) me[e] |= ((pwd[p][e >> 2] >> (8 * (e & 3))) & 0xffu) << (8 * p);
        }
        float val[E];
        uint4 packed = make_uint4(0u, 0u, 0u, 0u);
#pragma unroll
        for (int k = 0; k < V; ++k) {
            if (k == 0 || ((fm >> k) & 1u)) {  // uniform: a group starts here
                const uint32_t seg = plan.seg[k];
                const int n = __popc(seg);
                const float w = wts ? plan.gw[k] : (float)(1 + n);
#pragma unroll
                for (int e = 0; e < E; ++e) {
                    const int nsel = __popc(me[e] & seg);
                    val[e] = __fmul_rn(rcp_tab[nsel], __fmul_rn(t[e], w));  // (1 / num_selected) * g1; nsel == 0 only for NaN
                }
            }
            float o[E];
#pragma unroll
            for (int e = 0; e < E; ++e) o[e] = (me[e] & (1u << k)) ? val[e] : 0.0f;
            packed = Elem<T>::pack(o);
            stg_stream_16(reinterpret_cast<char *>(plan.rowptr[k]) + thread_off, packed);
        }
    }
    }  // units
}

template <typename T, int V, int NT = 256>
static int launch_bwd_fast_v(const void *dS, const int32_t *bins, int64_t bin_sb, const uint8_t *mask,
                             const float *weights, int64_t w_sb, const ViewPtrs &gp, int64_t g_sb, int32_t *status, int B, int64_t D, int G, int pool,
                             cudaStream_t st, int gapC = 0, int gapHW = 0)
{
    constexpr int E = Elem<T>::kVec;
    const int64_t td = (int64_t)NT * E;
    const int64_t tiles = (D + td - 1) / td;
    if ((int64_t)B * tiles > 0x7fffffffLL) return GVCNN_E_BAD_ARG;
    const int units = (int)(B * tiles);
    // one unit per CTA: two (the kernel's UPC parameter) measured equal or slower at every size - 70.5 vs 71.7 us at
    // configs[1], 10.3 vs 12.8 us at 512 shapes (profiles/r03_bwd_units_per_cta_ab.jsonl) - so only UPC = 1 is built
    constexpr int upc = 1;
    const unsigned grid = (unsigned)((units + upc - 1) / upc);
    cudaError_t err;
#define GVCNN_LAUNCH_BF(POOL_, GAP_, WTS_, UPC_)                                                              \
    err = launch_pdl(pool_fuse_bwd_fast_kernel<T, POOL_, V, NT, GAP_, WTS_, UPC_>, dim3(grid), dim3(NT), 0, st, \
                     static_cast<const T *>(dS), bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G,   \
                     (int)tiles, gapC, gapHW, units)
#define GVCNN_LAUNCH_BF_P(GAP_, WTS_)                                                                         \
    do {                                                                                                      \
        if (pool == GVCNN_POOL_MAX) GVCNN_LAUNCH_BF(GVCNN_POOL_MAX, GAP_, WTS_, upc);                         \
        else GVCNN_LAUNCH_BF(GVCNN_POOL_MEAN, GAP_, WTS_, upc);                                               \
    } while (0)
    if (gapC > 0) {
        if (weights) return GVCNN_E_UNSUPPORTED;  // the GAP-folded backward always uses the reference's weights
        GVCNN_LAUNCH_BF_P(true, false);
    } else if (weights) {
        GVCNN_LAUNCH_BF_P(false, true);
    } else {
        GVCNN_LAUNCH_BF_P(false, false);
    }
#undef GVCNN_LAUNCH_BF_P
#undef GVCNN_LAUNCH_BF
    if (err != cudaSuccess) return (int)err;
    return (int)cudaGetLastError();
}

template <typename T>
static int launch_bwd_fast_t(const void *dS, const int32_t *bins, int64_t bin_sb, const uint8_t *mask,
                             const float *weights, int64_t w_sb, const ViewPtrs &gp, int64_t g_sb, int32_t *status, int B, int V, int64_t D, int G,
                             int pool, cudaStream_t st)
{
    if (D < 256 * Elem<T>::kVec) {
        // bf16 at D = 1024: one 128-thread tile per shape (see pool_fwd_ring.cu); anything shorter: generic kernel
        if constexpr (Elem<T>::kVec == 8) {
            if (D < 128 * Elem<T>::kVec) return -1000;
            switch (V) {
            case 4: return launch_bwd_fast_v<T, 4, 128>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G, pool, st);
            case 6: return launch_bwd_fast_v<T, 6, 128>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G, pool, st);
            case 8: return launch_bwd_fast_v<T, 8, 128>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G, pool, st);
            case 12: return launch_bwd_fast_v<T, 12, 128>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G, pool, st);
            case 16: return launch_bwd_fast_v<T, 16, 128>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G, pool, st);
            case 20: return launch_bwd_fast_v<T, 20, 128>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G, pool, st);
            default: return -1000;
            }
        }
        return -1000;
    }
    switch (V) {
    case 4: return launch_bwd_fast_v<T, 4>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G, pool, st);
    case 6: return launch_bwd_fast_v<T, 6>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G, pool, st);
    case 8: return launch_bwd_fast_v<T, 8>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G, pool, st);
    case 12: return launch_bwd_fast_v<T, 12>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G, pool, st);
    case 16: return launch_bwd_fast_v<T, 16>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G, pool, st);
    case 20: return launch_bwd_fast_v<T, 20>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, D, G, pool, st);
    default: return -1000;
    }
}

// GAP-folded backward: dOut [B, C] -> dF; same applicability as the forward GAP kernel plus the fast-kernel V set
int launch_pool_fuse_gap_bwd(const void *dOut, const int32_t *bins, int64_t bin_sb, const uint8_t *mask,
                             const ViewPtrs &gp, int64_t g_sb, int32_t *status, int B, int V, int HW, int C, int G,
                             int pool, int dtype, cudaStream_t st)
{
    const int64_t D = (int64_t)HW * C;
    const int td = 256 * (dtype == GVCNN_F32 ? 4 : 8);
    if (C % td != 0) return -1000;
#define GVCNN_GAPB_CASE(T_)                                                                                   \
    switch (V) {                                                                                              \
    case 4: return launch_bwd_fast_v<T_, 4>(dOut, bins, bin_sb, mask, nullptr, 0, gp, g_sb, status, B, D, G, pool, st, C, HW);   \
    case 6: return launch_bwd_fast_v<T_, 6>(dOut, bins, bin_sb, mask, nullptr, 0, gp, g_sb, status, B, D, G, pool, st, C, HW);   \
    case 8: return launch_bwd_fast_v<T_, 8>(dOut, bins, bin_sb, mask, nullptr, 0, gp, g_sb, status, B, D, G, pool, st, C, HW);   \
    case 12: return launch_bwd_fast_v<T_, 12>(dOut, bins, bin_sb, mask, nullptr, 0, gp, g_sb, status, B, D, G, pool, st, C, HW); \
    case 16: return launch_bwd_fast_v<T_, 16>(dOut, bins, bin_sb, mask, nullptr, 0, gp, g_sb, status, B, D, G, pool, st, C, HW); \
    case 20: return launch_bwd_fast_v<T_, 20>(dOut, bins, bin_sb, mask, nullptr, 0, gp, g_sb, status, B, D, G, pool, st, C, HW); \
    default: return -1000;                                                                                    \
    }
    if (dtype == GVCNN_F32) { GVCNN_GAPB_CASE(float) }
    GVCNN_GAPB_CASE(__nv_bfloat16)
#undef GVCNN_GAPB_CASE
}

// returns -1000 when this fast path does not apply
int launch_pool_fuse_bwd_fast(const void *dS, const int32_t *bins, int64_t bin_sb, const uint8_t *mask,
                              const float *weights, int64_t w_sb, const ViewPtrs &gp, int64_t g_sb, int32_t *status, int B, int V, int64_t D, int G,
                              int pool, int dtype, cudaStream_t st)
{
    if (dtype == GVCNN_F32) {
        if (D % 4) return -1000;
        return launch_bwd_fast_t<float>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, V, D, G, pool, st);
    }
    if (D % 8) return -1000;
    return launch_bwd_fast_t<__nv_bfloat16>(dS, bins, bin_sb, mask, weights, w_sb, gp, g_sb, status, B, V, D, G, pool, st);
}

}  // namespace gvcnn
