// Score + binning kernels: nets/model.py:144-147 (per-view Dense(1) on the raw
// descriptor, sigmoid(log|x|)) and :23 (bin = int(score * G)).
//
// Shape of the work: B*V independent dot products of length C (1024) against
// V weight rows - a batch of GEMVs with 0.5 flop/byte, so HBM-bound; one warp
// per (shape, view) row, 128-bit streaming loads of R, W served from L1/L2
// (V*C*4 = 48 KB total), xor-butterfly reduction, everything after the dot
// product fused into lane 0's epilogue.  The summation order is fixed and is
// restated in oracle/gvcnn_oracle.c (oracle_view_score_x_kernel_order):
//   lane l accumulates chunks (i*32 + l) of E consecutive elements with one
//   fmaf chain, lanes combine with offsets 16,8,4,2,1, bias is added last.
#include "common.cuh"

namespace gvcnn {

constexpr int kScoreWarps = 8;
constexpr int kScoreUnroll = 8;  // 16-byte loads in flight per lane per pass

// One warp per (b, v) row.  VEC: 16-byte loads (E = 16/sizeof(T) elements per
// chunk) when rows are 16-byte aligned and C % E == 0, else E = 1.
template <typename T, bool VEC, bool FUSE_BIN>
__global__ void __launch_bounds__(kScoreWarps * 32)
view_score_kernel(const ViewPtrs rp, const int64_t r_sb, const float *__restrict__ W,
                  const float *__restrict__ bias, float *__restrict__ x_out, float *__restrict__ scores,
                  int32_t *__restrict__ bins, int32_t *__restrict__ flag_out, int32_t *status, const int B,
                  const int V, const int C, const int G, const int edge_ulps, const int clamp)
{
    constexpr int E = VEC ? Elem<T>::kVec : 1;
    const int lane = threadIdx.x & 31;
    const int64_t row = (int64_t)blockIdx.x * kScoreWarps + (threadIdx.x >> 5);
    pdl_wait();
    pdl_launch_dependents();
    if (row >= (int64_t)B * V) return;
    const int b = (int)(row / V);
    const int v = (int)(row - (int64_t)b * V);
    const T *__restrict__ r = reinterpret_cast<const T *>(rp.p[v]) + (int64_t)b * r_sb;
    const float *__restrict__ w = W + (int64_t)v * C;

    float acc = 0.0f;
    if constexpr (VEC) {
        for (int base0 = lane * E; base0 < C; base0 += 32 * E * kScoreUnroll) {
            uint4 raw[kScoreUnroll];
#pragma unroll
            for (int u = 0; u < kScoreUnroll; ++u) {
                const int base = base0 + u * 32 * E;
                if (base < C) raw[u] = ldg_stream_16(r + base);
            }
#pragma unroll
            for (int u = 0; u < kScoreUnroll; ++u) {
                const int base = base0 + u * 32 * E;
                if (base < C) {
                    float f[E];
                    Elem<T>::unpack(raw[u], f);
#pragma unroll
                    for (int j = 0; j < E; j += 4) {
                        const float4 wv = __ldg(reinterpret_cast<const float4 *>(w + base + j));
                        acc = fmaf(f[j + 0], wv.x, acc);
                        acc = fmaf(f[j + 1], wv.y, acc);
                        acc = fmaf(f[j + 2], wv.z, acc);
                        acc = fmaf(f[j + 3], wv.w, acc);
                    }
                }
            }
        }
    } else {
        for (int c = lane; c < C; c += 32) acc = fmaf(Elem<T>::to_float(r[c]), __ldg(w + c), acc);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));

    if (lane == 0) {
        const float x = __fadd_rn(acc, __ldg(bias + v));
        if (x_out) x_out[row] = x;
        if constexpr (FUSE_BIN) {
            float s;
            int bin;
            const int flags = score_and_bin(x, 1.0f, G, edge_ulps, clamp, s, bin);
            scores[row] = s;
            bins[row] = bin;
            publish(flags, flag_out ? flag_out + row : nullptr, status);
        }
    }
}

// Fast path: one warp reduces 4 rows of the same view (4 consecutive shapes), one 4 KB row at a
// time with all 32 lanes (so every burst of loads is one contiguous row, like the generic kernel),
// the next row's loads in flight while the current one is multiplied, and ONE transposing butterfly
// for the 4 partial sums (6 shuffles instead of 20) followed by one epilogue pass in 4 lanes.  The
// arithmetic is exactly the generic kernel's: lane l walks chunks (i*32 + l), butterfly offsets
// 16,8,4,2,1 (a + b is commutative, so which lane holds which row's partial sum does not matter),
// bias last - oracle_view_score_x_kernel_order with LPR = 32.  C is a compile-time multiple of the
// 32-lane x 16-byte x NB batch: no bounds checks.
constexpr int kFastRows = 4;  // rows (shapes) per warp

template <typename T, int NB, bool FUSE_BIN>  // NB = 16-byte loads per lane per row = C / (32 * E)
__global__ void __launch_bounds__(kScoreWarps * 32)
view_score_fast_kernel(const ViewPtrs rp, const int64_t r_sb, const float *__restrict__ W,
                       const float *__restrict__ bias, float *__restrict__ x_out, float *__restrict__ scores,
                       int32_t *__restrict__ bins, int32_t *__restrict__ flag_out, int32_t *status, const int B,
                       const int V, const int G, const int edge_ulps, const int clamp, const int64_t items)
{
    constexpr int E = Elem<T>::kVec;
    constexpr int C = NB * 32 * E;
    const int lane = threadIdx.x & 31;
    const int64_t item = (int64_t)blockIdx.x * kScoreWarps + (threadIdx.x >> 5);
    pdl_wait();
    pdl_launch_dependents();
    if (item >= items) return;
    const int v = (int)(item % V);
    const int b0 = (int)(item / V) * kFastRows;
    const T *__restrict__ r0 = reinterpret_cast<const T *>(rp.p[v]) + lane * E;
    const float *__restrict__ w = W + (int64_t)v * C + lane * E;

    uint4 buf[2][NB];
    float acc[kFastRows];
#pragma unroll
    for (int u = 0; u < NB; ++u) buf[0][u] = ldg_stream_16(r0 + (int64_t)min(b0, B - 1) * r_sb + u * 32 * E);
#pragma unroll
    for (int j = 0; j < kFastRows; ++j) {
        if (j + 1 < kFastRows) {
            const T *rn = r0 + (int64_t)min(b0 + j + 1, B - 1) * r_sb;
#pragma unroll
            for (int u = 0; u < NB; ++u) buf[(j + 1) & 1][u] = ldg_stream_16(rn + u * 32 * E);
        }
        float a = 0.0f;
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            float f[E];
            Elem<T>::unpack(buf[j & 1][u], f);
#pragma unroll
            for (int q = 0; q < E; q += 4) {
                const float4 wv = __ldg(reinterpret_cast<const float4 *>(w + u * 32 * E + q));
                a = fmaf(f[q + 0], wv.x, a);
                a = fmaf(f[q + 1], wv.y, a);
                a = fmaf(f[q + 2], wv.z, a);
                a = fmaf(f[q + 3], wv.w, a);
            }
        }
        acc[j] = a;
    }
    // transposing butterfly: after offsets 16 and 8, lane group (lane >> 3) holds row (lane >> 3)
    const bool hi = (lane & 16) != 0;
    float k0 = hi ? acc[2] : acc[0], k1 = hi ? acc[3] : acc[1];
    const float s0 = hi ? acc[0] : acc[2], s1 = hi ? acc[1] : acc[3];
    k0 = __fadd_rn(k0, __shfl_xor_sync(0xffffffffu, s0, 16));
    k1 = __fadd_rn(k1, __shfl_xor_sync(0xffffffffu, s1, 16));
    const bool h8 = (lane & 8) != 0;
    float k = h8 ? k1 : k0;
    const float sd = h8 ? k0 : k1;
    k = __fadd_rn(k, __shfl_xor_sync(0xffffffffu, sd, 8));
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) k = __fadd_rn(k, __shfl_xor_sync(0xffffffffu, k, off));

    const int b = b0 + (lane >> 3);
    if ((lane & 7) == 0 && b < B) {
        const int64_t row = (int64_t)b * V + v;
        const float x = __fadd_rn(k, __ldg(bias + v));
        if (x_out) x_out[row] = x;
        if constexpr (FUSE_BIN) {
            float s;
            int bin;
            const int flags = score_and_bin(x, 1.0f, G, edge_ulps, clamp, s, bin);
            scores[row] = s;
            bins[row] = bin;
            publish(flags, flag_out ? flag_out + row : nullptr, status);
        }
    }
}

// xsum[v] = sum_b x[b, v], fixed order: thread t adds b = t, t+256, ... in
// sequence, warps butterfly, thread 0 adds the 8 warp sums in warp order.
__global__ void __launch_bounds__(256) batch_sum_x_kernel(const float *__restrict__ x,
                                                          float *__restrict__ xsum, const int B, const int V)
{
    __shared__ float warp_sum[8];
    const int v = blockIdx.x;
    pdl_wait();
    pdl_launch_dependents();
    float acc = 0.0f;
    for (int b = threadIdx.x; b < B; b += 256) acc = __fadd_rn(acc, x[(int64_t)b * V + v]);
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) acc = __fadd_rn(acc, __shfl_xor_sync(0xffffffffu, acc, off));
    if ((threadIdx.x & 31) == 0) warp_sum[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = warp_sum[0];
        for (int i = 1; i < 8; ++i) t = __fadd_rn(t, warp_sum[i]);
        xsum[v] = t;
    }
}

__global__ void __launch_bounds__(256) score_bin_kernel(const float *__restrict__ x, const float denom,
                                                        float *__restrict__ x_mean, float *__restrict__ scores,
                                                        int32_t *__restrict__ bins,
                                                        int32_t *__restrict__ flag_out, int32_t *status,
                                                        const int64_t n, const int G, const int mult,
                                                        const int edge_ulps, const int clamp,
                                                        const bool x_is_score)
{
    const int64_t i = (int64_t)blockIdx.x * 256 + threadIdx.x;
    pdl_wait();
    pdl_launch_dependents();
    if (i >= n) return;
    float s;
    int bin;
    const int flags = score_and_bin(x[i], denom, G, edge_ulps, clamp, s, bin, x_is_score, mult);
    if (x_mean) x_mean[i] = __fdiv_rn(x[i], denom);  // tf.reduce_mean(raw), nets/model.py:146
    if (scores) scores[i] = s;
    bins[i] = bin;
    publish(flags, flag_out ? flag_out + i : nullptr, status);
}

template <typename T>
static int launch_view_score_t(const ViewPtrs &rp, int64_t r_sb, const float *W, const float *bias, float *x,
                               float *scores, int32_t *bins, int32_t *flags, int32_t *status, int B, int V,
                               int C, int G, bool vec, bool fuse_bin, int edge_ulps, int clamp,
                               cudaStream_t st)
{
    const int64_t rows = (int64_t)B * V;
    cudaError_t err = cudaSuccess;
    if (vec && (C == 8 * 32 * Elem<T>::kVec || C == 4 * 32 * Elem<T>::kVec)) {
        // C_raw = 1024 (block3 of ResNet-v2-50, nets/resnet_v2.py:242) is the reference's only width
        const int64_t items = (int64_t)V * ((B + kFastRows - 1) / kFastRows);
        const dim3 fgrid((unsigned)((items + kScoreWarps - 1) / kScoreWarps)), fblock(kScoreWarps * 32);
#define GVCNN_LAUNCH_FAST(NB_, FUSE_)                                                                        \
    err = launch_pdl(view_score_fast_kernel<T, NB_, FUSE_>, fgrid, fblock, 0, st, rp, r_sb, W, bias, x, scores, \
                     bins, flags, status, B, V, G, edge_ulps, clamp, items)
        if (C == 8 * 32 * Elem<T>::kVec) {
            if (fuse_bin) GVCNN_LAUNCH_FAST(8, true); else GVCNN_LAUNCH_FAST(8, false);
        } else {
            if (fuse_bin) GVCNN_LAUNCH_FAST(4, true); else GVCNN_LAUNCH_FAST(4, false);
        }
#undef GVCNN_LAUNCH_FAST
        if (err != cudaSuccess) return (int)err;
        return (int)cudaGetLastError();
    }
    const dim3 grid((unsigned)((rows + kScoreWarps - 1) / kScoreWarps)), block(kScoreWarps * 32);
#define GVCNN_LAUNCH_SCORE(VEC_, FUSE_)                                                               \
    err = launch_pdl(view_score_kernel<T, VEC_, FUSE_>, grid, block, 0, st, rp, r_sb, W, bias, x, scores, bins, \
                     flags, status, B, V, C, G, edge_ulps, clamp)
    if (vec) {
        if (fuse_bin) GVCNN_LAUNCH_SCORE(true, true); else GVCNN_LAUNCH_SCORE(true, false);
    } else {
        if (fuse_bin) GVCNN_LAUNCH_SCORE(false, true); else GVCNN_LAUNCH_SCORE(false, false);
    }
#undef GVCNN_LAUNCH_SCORE
    if (err != cudaSuccess) return (int)err;
    return (int)cudaGetLastError();
}

int launch_view_score(const ViewPtrs &rp, int64_t r_sb, const float *W, const float *bias, float *x,
                      float *scores, int32_t *bins, int32_t *flags, int32_t *status, int B, int V, int C,
                      int G, int dtype, bool aligned16, bool fuse_bin, int edge_ulps, int clamp,
                      cudaStream_t st)
{
    if (dtype == GVCNN_F32) {
        const bool vec = aligned16 && (C % 4 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
        return launch_view_score_t<float>(rp, r_sb, W, bias, x, scores, bins, flags, status, B, V, C, G, vec,
                                          fuse_bin, edge_ulps, clamp, st);
    }
    const bool vec = aligned16 && (C % 8 == 0) && ((reinterpret_cast<uintptr_t>(W) & 15) == 0);
    return launch_view_score_t<__nv_bfloat16>(rp, r_sb, W, bias, x, scores, bins, flags, status, B, V, C, G,
                                              vec, fuse_bin, edge_ulps, clamp, st);
}

int launch_batch_sum_x(const float *x, float *xsum, int B, int V, cudaStream_t st)
{
    const cudaError_t err = launch_pdl(batch_sum_x_kernel, dim3(V), dim3(256), 0, st, x, xsum, B, V);
    if (err != cudaSuccess) return (int)err;
    return (int)cudaGetLastError();
}

int launch_score_bin(const float *x, float denom, float *x_mean, float *scores, int32_t *bins, int32_t *flags,
                     int32_t *status, int64_t n, int G, int multiplier, int edge_ulps, int clamp, bool x_is_score,
                     cudaStream_t st)
{
    const unsigned grid = (unsigned)((n + 255) / 256);
    const cudaError_t err = launch_pdl(score_bin_kernel, dim3(grid), dim3(256), 0, st, x, denom, x_mean, scores, bins,
                                       flags, status, n, G, multiplier, edge_ulps, clamp, x_is_score);
    if (err != cudaSuccess) return (int)err;
    return (int)cudaGetLastError();
}

}  // namespace gvcnn
