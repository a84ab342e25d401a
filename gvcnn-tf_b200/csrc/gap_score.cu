// Global average pooling of the raw view maps folded into the score kernel (first half):
//   nets/model.py:144-145   raw = GlobalAveragePooling2D()(end_points['resnet_v2_50/block3'])   [N, h, w, 1024] -> [N, 1024]
//                           raw = Dense(1)(raw)                                                  one FC per view
// At the reference's geometry the block3 map of one (shape, view) is 10 x 10 x 1024 floats = 409,600 B - a third of
// the head's compulsory bytes - and the only thing the path needs from it is one scalar.  This kernel streams the map
// once and writes 4-16 bytes: no [N, V, 1024] intermediate, no separate mean op.
//
// Shape of the work: B*V items of HW*C contiguous elements (channel-last), each reduced over positions per channel,
// then dotted with the view's weight row.  HBM-bound (0.25 flop/byte).  One CTA of 8 warps per item: warp s streams a
// contiguous slice of positions (ceil/floor(HW/8) of them), every lane owning the same channel chunks the score
// kernel's lanes own (chunk u*32 + lane of E elements), accumulating per channel in float32 registers, two positions
// of loads (2 x NB x 16 B per lane) in flight; the 8 partial channel sums meet in shared memory and warp 0 finishes:
//   R[c]  = (p_0[c] + p_1[c] + ... + p_7[c]) / HW        slices ascending, then ONE IEEE division
//           (tf.reduce_mean: sum, then divide by the count; the order of the sum over positions is this kernel's -
//            TF's own is not reproducible without TF - and is restated in oracle/gvcnn_oracle.py gap_mean_kernel_order)
//   x     = sum_c R[c] * W[v, c] + bias[v]                 exactly view_score_kernel's order (fmaf chain per lane over
//           its chunks, butterfly 16,8,4,2,1, bias last), so x equals gvcnn_view_score_fwd on the same R bit for bit
// followed by the same epilogue (score, bin, flags).
#include "common.cuh"

namespace gvcnn {

constexpr int kGapWarps = 8;

template <typename T, int NB, bool FUSE_BIN>  // NB = 16-byte chunks per lane per position = C / (32 * E)
__global__ void __launch_bounds__(kGapWarps * 32, 2)
gap_score_kernel(const ViewPtrs mp, const int64_t m_sb, const float *__restrict__ W, const float *__restrict__ bias,
                 float *__restrict__ R_out, float *__restrict__ x_out, float *__restrict__ scores,
                 int32_t *__restrict__ bins, int32_t *__restrict__ flag_out, int32_t *status, const int B, const int V,
                 const int HW, const int G, const int edge_ulps, const int clamp)
{
    constexpr int E = Elem<T>::kVec;
    constexpr int C = NB * 32 * E;
    constexpr int NA = NB * E;  // channels (accumulators) per lane
    extern __shared__ __align__(16) float part_raw[];  // [kGapWarps][C] partial channel sums (32 / 64 KB)
    float (*part)[C] = reinterpret_cast<float (*)[C]>(part_raw);

    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int item = blockIdx.x;
    const int b = item / V;
    const int v = item - b * V;
    // slice of positions of this warp: the first HW % 8 slices have one more
    const int q = HW / kGapWarps, r = HW % kGapWarps;
    const int p0 = warp * q + min(warp, r);
    const int np = q + (warp < r ? 1 : 0);

    pdl_wait();
    pdl_launch_dependents();
    const T *__restrict__ src = reinterpret_cast<const T *>(mp.p[v]) + (int64_t)b * m_sb + (int64_t)p0 * C + lane * E;

    float acc[NB][E];
#pragma unroll
    for (int i = 0; i < NA; ++i) acc[i / E][i % E] = 0.0f;
    uint4 buf[2][NB];
    if (np > 0) {
#pragma unroll
        for (int u = 0; u < NB; ++u) buf[0][u] = ldg_stream_16(src + u * 32 * E);
    }
    for (int j = 0; j < np; j += 2) {
        if (j + 1 < np) {
#pragma unroll
            for (int u = 0; u < NB; ++u) buf[1][u] = ldg_stream_16(src + (int64_t)(j + 1) * C + u * 32 * E);
        }
#pragma unroll
        for (int u = 0; u < NB; ++u) {
            Elem<T>::add_to(acc[u], buf[0][u]);
        }
        if (j + 1 < np) {
            if (j + 2 < np) {
#pragma unroll
                for (int u = 0; u < NB; ++u) buf[0][u] = ldg_stream_16(src + (int64_t)(j + 2) * C + u * 32 * E);
            }
#pragma unroll
            for (int u = 0; u < NB; ++u) {
                Elem<T>::add_to(acc[u], buf[1][u]);
            }
        }
    }
    // partial channel sums of this slice -> shared memory (lane's chunks, 16-byte stores, conflict-free)
#pragma unroll
    for (int u = 0; u < NB; ++u)
#pragma unroll
        for (int e = 0; e < E; e += 4)
            *reinterpret_cast<float4 *>(&part[warp][(u * 32 + lane) * E + e]) =
                make_float4(acc[u][e], acc[u][e + 1], acc[u][e + 2], acc[u][e + 3]);
    __syncthreads();
    if (warp != 0) return;

    const float hw = (float)HW;
    const float *__restrict__ w = W + (int64_t)v * C + lane * E;
    float a = 0.0f;
#pragma unroll
    for (int u = 0; u < NB; ++u) {
        float m[E];
#pragma unroll
        for (int e = 0; e < E; e += 4) {
            float4 s4 = *reinterpret_cast<const float4 *>(&part[0][(u * 32 + lane) * E + e]);
#pragma unroll
            for (int s = 1; s < kGapWarps; ++s) {  // slices in ascending order; empty slices hold exact zeros
                const float4 t = *reinterpret_cast<const float4 *>(&part[s][(u * 32 + lane) * E + e]);
                s4.x = __fadd_rn(s4.x, t.x); s4.y = __fadd_rn(s4.y, t.y);
                s4.z = __fadd_rn(s4.z, t.z); s4.w = __fadd_rn(s4.w, t.w);
            }
            m[e] = __fdiv_rn(s4.x, hw); m[e + 1] = __fdiv_rn(s4.y, hw);
            m[e + 2] = __fdiv_rn(s4.z, hw); m[e + 3] = __fdiv_rn(s4.w, hw);
        }
        if (R_out) {
            float *ro = R_out + ((int64_t)b * V + v) * C + (u * 32 + lane) * E;
#pragma unroll
            for (int e = 0; e < E; e += 4) *reinterpret_cast<float4 *>(ro + e) = make_float4(m[e], m[e + 1], m[e + 2], m[e + 3]);
        }
#pragma unroll
        for (int e = 0; e < E; e += 4) {
            const float4 wv = __ldg(reinterpret_cast<const float4 *>(w + u * 32 * E + e));
            a = fmaf(m[e + 0], wv.x, a);
            a = fmaf(m[e + 1], wv.y, a);
            a = fmaf(m[e + 2], wv.z, a);
            a = fmaf(m[e + 3], wv.w, a);
        }
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) a = __fadd_rn(a, __shfl_xor_sync(0xffffffffu, a, off));
    if (lane == 0) {
        const int64_t row = (int64_t)b * V + v;
        const float x = __fadd_rn(a, __ldg(bias + v));
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

template <typename T>
static int launch_gap_score_t(const ViewPtrs &mp, int64_t m_sb, const float *W, const float *bias, float *R_out, float *x,
                              float *scores, int32_t *bins, int32_t *flags, int32_t *status, int B, int V, int HW, int C,
                              int G, bool fuse_bin, int edge_ulps, int clamp, cudaStream_t st)
{
    constexpr int E = Elem<T>::kVec;
    const int64_t items = (int64_t)B * V;
    if (items > 0x7fffffffLL) return GVCNN_E_BAD_ARG;
    const dim3 grid((unsigned)items), block(kGapWarps * 32);
    cudaError_t err = cudaSuccess;
#define GVCNN_LAUNCH_GAP_SCORE(NB_, FUSE_)                                                                        \
    do {                                                                                                          \
        const size_t smem = (size_t)kGapWarps * NB_ * 32 * E * sizeof(float);                                     \
        err = ensure_dyn_smem<gap_score_kernel<T, NB_, FUSE_>>((int)smem);                                        \
        if (err == cudaSuccess)                                                                                   \
            err = launch_pdl(gap_score_kernel<T, NB_, FUSE_>, grid, block, smem, st, mp, m_sb, W, bias, R_out, x, \
                             scores, bins, flags, status, B, V, HW, G, edge_ulps, clamp);                         \
    } while (0)
    if (C == 8 * 32 * E) {
        if (fuse_bin) GVCNN_LAUNCH_GAP_SCORE(8, true); else GVCNN_LAUNCH_GAP_SCORE(8, false);
    } else if (C == 4 * 32 * E) {
        if (fuse_bin) GVCNN_LAUNCH_GAP_SCORE(4, true); else GVCNN_LAUNCH_GAP_SCORE(4, false);
    } else {
        return -1000;
    }
#undef GVCNN_LAUNCH_GAP_SCORE
    if (err != cudaSuccess) return (int)err;
    return (int)cudaGetLastError();
}

// returns -1000 when the channel count is not one of the instantiated ones
int launch_gap_score(const ViewPtrs &mp, int64_t m_sb, const float *W, const float *bias, float *R_out, float *x,
                     float *scores, int32_t *bins, int32_t *flags, int32_t *status, int B, int V, int HW, int C, int G,
                     int dtype, bool fuse_bin, int edge_ulps, int clamp, cudaStream_t st)
{
    if (dtype == GVCNN_F32)
        return launch_gap_score_t<float>(mp, m_sb, W, bias, R_out, x, scores, bins, flags, status, B, V, HW, C, G, fuse_bin,
                                         edge_ulps, clamp, st);
    return launch_gap_score_t<__nv_bfloat16>(mp, m_sb, W, bias, R_out, x, scores, bins, flags, status, B, V, HW, C, G,
                                             fuse_bin, edge_ulps, clamp, st);
}

}  // namespace gvcnn
