/*
 * C restatement of the GVCNN grouping + fusion path.  TEST INFRASTRUCTURE ONLY.
 *
 * Second, independent oracle beside oracle/gvcnn_oracle.py (NumPy).  It exists
 * so that (1) the two restatements can be required to agree bit for bit,
 * (2) full BASELINE-size batches (B=4096, V=12, D=2048) can be checked in
 * seconds, and (3) the float32 summation order of the CUDA score kernel can be
 * reproduced exactly with fmaf() (oracle_view_score_x_kernel_order).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load
 * this library; the product (gvcnn-tf_b200/) never does.
 *
 * Build (see oracle/Makefile):  gcc -O2 -ffp-contract=off -fno-fast-math
 * -fopenmp -shared -fPIC.  -ffp-contract=off matters: every float32 product
 * and sum below must round once, like one TF op per arithmetic op.
 *
 * Parity unpinned against live TensorFlow (not installable in this image);
 * pinned by tests/golden/ (KATs from unit_test.py:18-19 + the reference's own
 * graph code run over a NumPy stand-in) and by agreement with the NumPy oracle.
 *
 * Reference lines followed (all in /root/reference/nets/model.py):
 *   :23       bin = int(float32(score) * float32(G))     -> oracle_bins
 *   :28-41    w_g = 1 + count_g                          -> inside pool_fuse
 *   :62-72    per group: gather | ones dummy, reduce_max -> oracle_pool_fuse_fwd
 *   :94-100   sum_g w_g*P_g (left to right) / sum_g w_g  -> oracle_pool_fuse_fwd
 *   :144-147  x = R.W + b ; s = sigmoid(log|x|)          -> oracle_view_score_*
 *   TF autodiff of :62-100               -> oracle_pool_fuse_bwd
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define POOL_MAX 0
#define POOL_MEAN 1

/* ---- score ------------------------------------------------------------ */

/* x[b,v] = sum_c R[b,v,c] * W[v,c] + bias[v] in float64 (the value of the
 * mathematics; model.py:144-145).  R addressed as R[b*rsb + v*rsv + c]. */
void oracle_view_score_x_f64(const float *R, const float *W, const float *bias,
                             double *x, int B, int V, int C, int64_t rsb, int64_t rsv)
{
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int v = 0; v < V; ++v) {
            const float *r = R + b * rsb + v * rsv;
            const float *w = W + (int64_t)v * C;
            double acc = 0.0;
            for (int c = 0; c < C; ++c) acc += (double)r[c] * (double)w[c];
            x[(int64_t)b * V + v] = acc + (double)bias[v];
        }
}

/* The CUDA score kernels' float32 order, restated with fmaf: a row is reduced by LPR lanes
 * (32 in the generic kernel, 8 in the fast one: 4 rows per warp); lane l walks chunks
 * (i*LPR + l) of E consecutive elements with one fused multiply-add chain, the LPR lane sums
 * are combined by an xor butterfly (offsets LPR/2 ... 1), the bias is added last.  E = vector
 * width in elements (4 for float32 rows, 8 for bf16 rows, 1 for the unaligned fallback). */
void oracle_view_score_x_kernel_order(const float *R, const float *W, const float *bias,
                                      float *x, int B, int V, int C, int64_t rsb, int64_t rsv, int E, int LPR)
{
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b)
        for (int v = 0; v < V; ++v) {
            const float *r = R + b * rsb + v * rsv;
            const float *w = W + (int64_t)v * C;
            float lane[32], nxt[32];
            for (int l = 0; l < LPR; ++l) {
                float acc = 0.0f;
                for (int64_t base = (int64_t)l * E; base < C; base += LPR * (int64_t)E)
                    for (int j = 0; j < E && base + j < C; ++j)
                        acc = fmaf(r[base + j], w[base + j], acc);
                lane[l] = acc;
            }
            for (int off = LPR / 2; off >= 1; off >>= 1) {
                for (int l = 0; l < LPR; ++l) nxt[l] = lane[l] + lane[l ^ off];
                memcpy(lane, nxt, sizeof(float) * LPR);
            }
            x[(int64_t)b * V + v] = lane[0] + bias[v];
        }
}

/* s = |x| / (1 + |x|)  ==  sigmoid(log|x|) (model.py:147) as one IEEE
 * division; x = 0 -> 0, |x| = inf -> 1, NaN -> NaN. */
float oracle_score_f32(float x)
{
    float ax = fabsf(x);
    if (isinf(ax)) return 1.0f;
    return ax / (1.0f + ax);
}

/* the literal composition in float32 with libm, for the edge report */
float oracle_score_f32_literal(float x)
{
    float y = logf(fabsf(x));
    return 1.0f / (1.0f + expf(-y));
}

/* bin = (int)(float32(s) * float32(G)), truncation toward zero (model.py:23).
 * NaN -> INT32_MIN (the reference raises ValueError); no range check here:
 * bin == G is the reference's IndexError case. */
void oracle_bins(const float *s, int32_t *bins, int64_t n, int G)
{
    for (int64_t i = 0; i < n; ++i) {
        volatile float t = s[i] * (float)G;
        bins[i] = isnan(t) ? INT32_MIN : (int32_t)t;
    }
}

/* ---- pooling + fusion forward ----------------------------------------- */

/* F addressed as F[b*fsb + v*fsv + d] (covers [B,V,D], [V,B,D]); bins as
 * bins[b*bin_sb + v] (bin_sb = 0 shares one scheme across the batch, the
 * literal score_reduce='batch' mode).  mask (nullable): uint8 planes
 * [ceil(V/8)][B][D], bit k%8 of plane k/8 set iff the k-th view in (bin, view)
 * order attains its group's max (max mode only).  Returns 0, or -1 if a bin is
 * outside [0, G). */
int oracle_pool_fuse_fwd(const float *F, const int32_t *bins, float *S, uint8_t *mask,
                         int B, int V, int D, int G, int pool, float fill,
                         int64_t fsb, int64_t fsv, int64_t bin_sb)
{
    int bad = 0;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const int32_t *bn = bins + b * bin_sb;
        int *cnt = (int *)calloc((size_t)G, sizeof(int));
        int *rank = (int *)malloc(sizeof(int) * (size_t)V);   /* sorted position of view v */
        int ok = 1;
        for (int v = 0; v < V; ++v) {
            if (bn[v] < 0 || bn[v] >= G) { ok = 0; break; }
            cnt[bn[v]]++;
        }
        if (!ok) { bad = 1; free(cnt); free(rank); continue; }
        {
            int pos = 0;
            for (int g = 0; g < G; ++g)
                for (int v = 0; v < V; ++v)
                    if (bn[v] == g) rank[v] = pos++;
        }
        float sumw = 0.0f;                                     /* tf.reduce_sum(weights) */
        for (int g = 0; g < G; ++g) sumw = sumw + (float)(1 + cnt[g]);
        const float *Fb = F + b * fsb;
        for (int d = 0; d < D; ++d) {
            float acc = 0.0f;
            int first = 1;
            for (int g = 0; g < G; ++g) {
                float P;
                if (cnt[g] == 0) {
                    P = fill;                                  /* reduce over the dummy */
                } else {
                    int seen = 0;
                    P = 0.0f;
                    for (int v = 0; v < V; ++v) {
                        if (bn[v] != g) continue;
                        float f = Fb[v * fsv + d];
                        if (!seen) { P = f; seen = 1; }
                        else if (pool == POOL_MAX) { P = (f > P) ? f : P; }
                        else { P = P + f; }
                    }
                    if (pool == POOL_MEAN) P = P / (float)cnt[g];
                    if (mask && pool == POOL_MAX)
                        for (int v = 0; v < V; ++v)
                            if (bn[v] == g && Fb[v * fsv + d] == P) {
                                int k = rank[v];
                                mask[((int64_t)(k >> 3) * B + b) * D + d] |= (uint8_t)(1u << (k & 7));
                            }
                }
                float term = (float)(1 + cnt[g]) * P;           /* tf.multiply(w_g, P_g) */
                if (first) { acc = term; first = 0; } else { acc = acc + term; }
            }
            S[(int64_t)b * D + d] = acc / sumw;                /* tf.div */
        }
        free(cnt); free(rank);
    }
    return bad ? -1 : 0;
}

/* ---- pooling + fusion backward ---------------------------------------- */

/* dF[b*gsb + v*gsv + d]; TF op order: g0 = dS / sumw; g1 = g0 * w_g;
 * max: dF = (1 / num_selected) * g1 for views attaining the max, 0 otherwise;
 * mean: dF = g1 / n_g. */
int oracle_pool_fuse_bwd(const float *dS, const float *F, const int32_t *bins, float *dF,
                         int B, int V, int D, int G, int pool,
                         int64_t fsb, int64_t fsv, int64_t gsb, int64_t gsv, int64_t bin_sb)
{
    int bad = 0;
#pragma omp parallel for schedule(static)
    for (int b = 0; b < B; ++b) {
        const int32_t *bn = bins + b * bin_sb;
        int *cnt = (int *)calloc((size_t)G, sizeof(int));
        int ok = 1;
        for (int v = 0; v < V; ++v) {
            if (bn[v] < 0 || bn[v] >= G) { ok = 0; break; }
            cnt[bn[v]]++;
        }
        if (!ok) { bad = 1; free(cnt); continue; }
        float sumw = 0.0f;
        for (int g = 0; g < G; ++g) sumw = sumw + (float)(1 + cnt[g]);
        const float *Fb = F ? F + b * fsb : NULL;
        float *Gb = dF + b * gsb;
        for (int d = 0; d < D; ++d) {
            float g0 = dS[(int64_t)b * D + d] / sumw;
            for (int g = 0; g < G; ++g) {
                if (cnt[g] == 0) continue;
                float g1 = g0 * (float)(1 + cnt[g]);
                if (pool == POOL_MAX) {
                    float P = 0.0f; int seen = 0, nsel = 0;
                    for (int v = 0; v < V; ++v)
                        if (bn[v] == g) {
                            float f = Fb[v * fsv + d];
                            if (!seen) { P = f; seen = 1; } else { P = (f > P) ? f : P; }
                        }
                    for (int v = 0; v < V; ++v)
                        if (bn[v] == g && Fb[v * fsv + d] == P) nsel++;
                    float share = 1.0f / (float)nsel;
                    for (int v = 0; v < V; ++v)
                        if (bn[v] == g)
                            Gb[v * gsv + d] = (Fb[v * fsv + d] == P) ? share * g1 : 0.0f;
                } else {
                    float val = g1 / (float)cnt[g];
                    for (int v = 0; v < V; ++v)
                        if (bn[v] == g) Gb[v * gsv + d] = val;
                }
            }
        }
        free(cnt);
    }
    return bad ? -1 : 0;
}

int oracle_version(void) { return 1; }
