// Timeline of the persistent ring kernel (pool+fuse forward): where do the ~10 us that do not scale with
// the batch go?  Compiles the product kernel with its trace marks enabled and prints, per launch, the
// spread of CTA start times, the time to the first landed tile, and the spread of CTA end times.
//   nvcc -std=c++17 -O3 -gencode arch=compute_100a,code=sm_100a -fmad=false -DGVCNN_RING_TRACE \
//        -I include -o gpurun_out/ring_trace_probe scripts/ring_trace_probe.cu
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../gvcnn-tf_b200/csrc/pool_fwd_ring.cu"

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e_)); return 1; } } while (0)

static double pct(std::vector<double> &v, double q) { return v[(size_t)(q * (v.size() - 1))]; }

__global__ void null_kernel(int *p) { if (p && threadIdx.x == 9999) *p = 1; }

static int null_launch_overhead()
{
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    CK(cudaFuncSetAttribute(null_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304));
    for (int mode = 0; mode < 4; ++mode) {
        // 0: <<<1,32>>>   1: <<<296,288, 96 KB>>>   2: same through cudaLaunchKernelEx + PDL attribute   3: 1,32 + PDL
        std::vector<double> us;
        for (int it = 0; it < 40; ++it) {
            CK(cudaEventRecord(e0));
            if (mode == 0) null_kernel<<<1, 32>>>(nullptr);
            else if (mode == 1) null_kernel<<<296, 288, 98304>>>(nullptr);
            else if (mode == 2) CK(gvcnn::launch_pdl(null_kernel, dim3(296), dim3(288), 98304, 0, (int *)nullptr));
            else CK(gvcnn::launch_pdl(null_kernel, dim3(1), dim3(32), 0, 0, (int *)nullptr));
            CK(cudaEventRecord(e1));
            CK(cudaDeviceSynchronize());
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (it >= 8) us.push_back(ms * 1e3);
        }
        std::sort(us.begin(), us.end());
        printf("null kernel, mode %d: event-to-event p10 %.2f p50 %.2f p90 %.2f us\n", mode, pct(us, .1), pct(us, .5), pct(us, .9));
    }
    return 0;
}

int main(int argc, char **argv)
{
    if (null_launch_overhead()) return 1;
    const int B = argc > 1 ? atoi(argv[1]) : 4096, V = 12, G = 8;
    const int64_t D = 2048;
    const int NSETS = 3;
    float *F[NSETS], *S;
    int32_t *bins;
    for (int i = 0; i < NSETS; ++i) {
        CK(cudaMalloc(&F[i], (size_t)B * V * D * 4));
        CK(cudaMemset(F[i], 0x3c, (size_t)B * V * D * 4));
    }
    CK(cudaMalloc(&S, (size_t)B * D * 4));
    std::vector<int32_t> hb((size_t)B * V);
    srand(1);
    for (auto &x : hb) x = rand() % G;
    CK(cudaMalloc(&bins, hb.size() * 4));
    CK(cudaMemcpy(bins, hb.data(), hb.size() * 4, cudaMemcpyHostToDevice));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    for (int it = 0; it < 8; ++it) {
        gvcnn::ViewPtrs fp = {};
        for (int v = 0; v < V; ++v) fp.p[v] = reinterpret_cast<char *>(F[it % NSETS]) + (size_t)v * D * 4;
        CK(cudaEventRecord(e0));
        int rc = gvcnn::launch_pool_fuse_fwd_ring(fp, (int64_t)V * D, bins, V, nullptr, 0, S, nullptr, nullptr, B, V, D, G,
                                                  GVCNN_POOL_MAX, 1.0f, GVCNN_F32, 0);
        CK(cudaEventRecord(e1));
        if (rc) { printf("launch rc %d\n", rc); return 1; }
        CK(cudaDeviceSynchronize());
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it < 4) continue;
        static unsigned long long h[8 * 1024];
        CK(cudaMemcpyFromSymbol(h, gvcnn::g_ring_trace, sizeof(h)));
        const int n = 296;
        unsigned long long t0 = ~0ull, tend = 0;
        for (int c = 0; c < n; ++c) { t0 = std::min(t0, h[8 * c]); tend = std::max(tend, h[8 * c + 2]); }
        static std::vector<double> prev_idle;
        std::vector<double> idle;
        for (int c = 0; c < n; ++c) idle.push_back((tend - h[8 * c + 2]) * 1e-3);
        if (!prev_idle.empty()) {   // is the end-time spread systematic (same CTAs late every launch)?
            double ma = 0, mb = 0, sab = 0, saa = 0, sbb = 0;
            for (int c = 0; c < n; ++c) { ma += idle[c]; mb += prev_idle[c]; }
            ma /= n; mb /= n;
            for (int c = 0; c < n; ++c) { sab += (idle[c] - ma) * (prev_idle[c] - mb); saa += (idle[c] - ma) * (idle[c] - ma); sbb += (prev_idle[c] - mb) * (prev_idle[c] - mb); }
            double s0 = 0, s1 = 0;  // the two co-resident rounds of CTAs: blockIdx < 148 and >= 148
            for (int c = 0; c < n; ++c) (c < n / 2 ? s0 : s1) += idle[c];
            printf("  end-idle correlation with the previous launch: %.2f; mean idle CTAs 0..147 %.2f us, 148..295 %.2f us\n",
                   sab / sqrt(saa * sbb + 1e-30), s0 / (n / 2), s1 / (n / 2));
        }
        prev_idle = idle;
        std::vector<double> st, ff, en, pd, m4, m5, m6;
        for (int c = 0; c < n; ++c) {
            st.push_back((h[8 * c] - t0) * 1e-3);
            ff.push_back((h[8 * c + 1] - h[8 * c]) * 1e-3);
            en.push_back((tend - h[8 * c + 2]) * 1e-3);
            pd.push_back((tend - h[8 * c + 3]) * 1e-3);
            m4.push_back((h[8 * c + 4] - h[8 * c]) * 1e-3);
            m5.push_back((h[8 * c + 5] - h[8 * c]) * 1e-3);
            m6.push_back((h[8 * c + 6] - h[8 * c]) * 1e-3);
        }
        std::sort(st.begin(), st.end()); std::sort(ff.begin(), ff.end()); std::sort(en.begin(), en.end()); std::sort(pd.begin(), pd.end());
        std::sort(m4.begin(), m4.end()); std::sort(m5.begin(), m5.end()); std::sort(m6.begin(), m6.end());
        printf("B=%d launch %d: event %.1f us, first start -> last end %.1f us\n", B, it, ms * 1e3, (tend - t0) * 1e-3);
        printf("  CTA start after first start   : p0 %.2f p50 %.2f p90 %.2f p100 %.2f us\n", pct(st, 0), pct(st, .5), pct(st, .9), pct(st, 1));
        printf("  entry -> first tile landed    : p0 %.2f p50 %.2f p90 %.2f p100 %.2f us\n", pct(ff, 0), pct(ff, .5), pct(ff, .9), pct(ff, 1));
        printf("  entry -> past dependency wait : p0 %.2f p50 %.2f p90 %.2f p100 %.2f us\n", pct(m4, 0), pct(m4, .5), pct(m4, .9), pct(m4, 1));
        printf("  entry -> first bins in regs   : p0 %.2f p50 %.2f p90 %.2f p100 %.2f us\n", pct(m5, 0), pct(m5, .5), pct(m5, .9), pct(m5, 1));
        printf("  entry -> first bulk issued    : p0 %.2f p50 %.2f p90 %.2f p100 %.2f us\n", pct(m6, 0), pct(m6, .5), pct(m6, .9), pct(m6, 1));
        printf("  CTA idle before kernel end    : p0 %.2f p50 %.2f p90 %.2f p100 %.2f us\n", pct(en, 0), pct(en, .5), pct(en, .9), pct(en, 1));
        printf("  producer idle before kernel end: p0 %.2f p50 %.2f p90 %.2f p100 %.2f us\n", pct(pd, 0), pct(pd, .5), pct(pd, .9), pct(pd, 1));
    }
    return 0;
}
