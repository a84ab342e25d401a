// One-shot all-reduce over NVLink peer memory for the two tiny exchanges of the path:
//   (1) the flat parameter-gradient bucket of the training step, V * (C_raw + 1) floats = 49 KB at V = 12 - the
//       B200-native form of `nccl_ops.all_sum(grads)` + `* 1/K` at utils/_train_helper.py:17-31;
//   (2) the V per-view partial sums of the literal batch mean before binning (nets/model.py:146 on a sharded batch).
// Both are latency-bound (an NCCL all-reduce of this size costs 20-33 us on 2-8 B200s against a 100-180 us step), so
// they get ONE kernel per rank and no rendezvous on the host:
//   push   every rank stores its vector into its own slot of EVERY peer's receive buffer (plain stores to
//          cudaIpc-mapped peer memory: NVLink writes), then a system-scope release store of the call's sequence
//          number into its flag at every peer;
//   wait   spin (acquire loads of LOCAL memory) until all ranks' flags show this sequence number;
//   reduce add the slots in rank order 0..K-1 - the same order on every rank, so all ranks get bit-identical sums
//          (every rank must derive the same bins) - scale, write in place.
// Channels: the vector is cut into fixed 2048-float channels, one CTA each, with private slots, flags and sequence
// counters, so CTAs never synchronise with each other.  Slots are double-buffered by sequence parity: a rank can
// only push call s+2 after it has finished call s+1, i.e. after every peer has pushed s+1, i.e. after every peer has
// finished READING call s - no second barrier.  The sequence counters live in device memory and are advanced by
// the kernel, so a captured launch replays correctly from a CUDA graph.
// A rank that never arrives would hang its peers: the wait gives up after kCommTimeoutNs and raises an error word the
// host can read (gvcnn_comm_error), instead of hanging the GPU.
#include <cstring>
#include <new>

#include "common.cuh"

using namespace gvcnn;

namespace {
constexpr int kCommMaxWorld = GVCNN_COMM_MAX_WORLD;
constexpr int kChanFloats = 2048;                                    // one CTA's share
constexpr int kCommMaxChan = GVCNN_COMM_MAX_FLOATS / kChanFloats;    // 8
constexpr int kCommThreads = 512;                                    // one float4 each per pass
constexpr unsigned long long kCommTimeoutNs = 4000000000ull;         // 4 s

// device view of one rank's receive buffer
struct CommBuf {
    float data[2][kCommMaxWorld][GVCNN_COMM_MAX_FLOATS];    // [phase][source rank][float]
    uint32_t flag[2][kCommMaxWorld][kCommMaxChan][8];       // [phase][source rank][channel], padded to 32 B
    uint32_t seq[kCommMaxChan][8];                          // local: last sequence number used per channel
    uint32_t error;                                         // local: set when a wait timed out
};

struct CommPeers {
    CommBuf *buf[kCommMaxWorld];
};

__device__ __forceinline__ void st_release_sys(uint32_t *p, uint32_t v)
{
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t *p)
{
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ unsigned long long global_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

__global__ void __launch_bounds__(kCommThreads)
allreduce_oneshot_kernel(const CommPeers peers, const int rank, const int world, float *__restrict__ data, const int n,
                         const float scale)
{
    const int c = blockIdx.x;
    const int lo = c * kChanFloats;
    const int cnt = min(n - lo, kChanFloats);            // floats of this channel (> 0 by the grid size)
    CommBuf *mine = peers.buf[rank];
    const uint32_t seq = mine->seq[c][0] + 1u;           // every thread reads it before thread 0 advances it below
    const int phase = (int)(seq & 1u);
    const int i = threadIdx.x * 4;                       // this thread's float4 of the channel
    const bool vec_ok = (i + 3 < cnt) && ((reinterpret_cast<uintptr_t>(data) & 15) == 0);
    float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (vec_ok) {
        const float4 t = *reinterpret_cast<const float4 *>(data + lo + i);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        for (int j = 0; j < 4; ++j)
            if (i + j < cnt) v[j] = data[lo + i + j];
    }
    // ---- push: my slot at every peer (own buffer included), farthest-first rotation to spread the links
    if (i < cnt) {
        for (int d = 0; d < world; ++d) {
            const int p = (rank + 1 + d) % world;
            float *dst = &peers.buf[p]->data[phase][rank][lo + i];
            *reinterpret_cast<float4 *>(dst) = make_float4(v[0], v[1], v[2], v[3]);  // slots are 16-byte aligned, padded
        }
    }
    __threadfence_system();
    __syncthreads();                                     // all of this CTA's stores are ordered before the flags
    if ((int)threadIdx.x < world) st_release_sys(&peers.buf[threadIdx.x]->flag[phase][rank][c][0], seq);
    if (threadIdx.x == 0) mine->seq[c][0] = seq;
    // ---- wait: every source rank's flag for this channel and phase
    if ((int)threadIdx.x < world) {
        const uint32_t *f = &mine->flag[phase][threadIdx.x][c][0];
        const unsigned long long t0 = global_timer_ns();
        while (ld_acquire_sys(f) != seq) {
            if (global_timer_ns() - t0 > kCommTimeoutNs) {
                atomicExch(&mine->error, 1u);
                break;
            }
        }
    }
    __syncthreads();
    // ---- reduce in rank order (identical on every rank), scale, store in place
    if (i < cnt) {
        float acc[4];
        {
            const float4 t = __ldcg(reinterpret_cast<const float4 *>(&mine->data[phase][0][lo + i]));
            acc[0] = t.x; acc[1] = t.y; acc[2] = t.z; acc[3] = t.w;
        }
        for (int r = 1; r < world; ++r) {
            const float4 t = __ldcg(reinterpret_cast<const float4 *>(&mine->data[phase][r][lo + i]));
            acc[0] = __fadd_rn(acc[0], t.x); acc[1] = __fadd_rn(acc[1], t.y);
            acc[2] = __fadd_rn(acc[2], t.z); acc[3] = __fadd_rn(acc[3], t.w);
        }
        if (scale != 1.0f)
            for (int j = 0; j < 4; ++j) acc[j] = __fmul_rn(acc[j], scale);
        if (vec_ok) {
            *reinterpret_cast<float4 *>(data + lo + i) = make_float4(acc[0], acc[1], acc[2], acc[3]);
        } else {
            for (int j = 0; j < 4; ++j)
                if (i + j < cnt) data[lo + i + j] = acc[j];
        }
    }
}
}  // namespace

struct gvcnn_comm {
    int rank, world, device;
    bool connected;
    CommBuf *local;
    CommPeers peers;
};

extern "C" {

int gvcnn_comm_create(gvcnn_comm **out, int rank, int world, void *handle_out)
{
    static_assert(GVCNN_COMM_HANDLE_BYTES >= sizeof(cudaIpcMemHandle_t), "handle size");
    static_assert(kChanFloats == kCommThreads * 4, "one float4 per thread per channel");
    if (!out || !handle_out || world < 1 || world > kCommMaxWorld || rank < 0 || rank >= world) return GVCNN_E_BAD_ARG;
    *out = nullptr;
    int rc = gvcnn_check_device();
    if (rc) return rc;
    gvcnn_comm *c = new (std::nothrow) gvcnn_comm();
    if (!c) return (int)cudaErrorMemoryAllocation;
    c->rank = rank;
    c->world = world;
    c->connected = false;
    c->local = nullptr;
    for (int i = 0; i < kCommMaxWorld; ++i) c->peers.buf[i] = nullptr;
    cudaError_t err = cudaGetDevice(&c->device);
    if (err == cudaSuccess) err = cudaMalloc(reinterpret_cast<void **>(&c->local), sizeof(CommBuf));
    if (err == cudaSuccess) err = cudaMemset(c->local, 0, sizeof(CommBuf));
    cudaIpcMemHandle_t h;
    if (err == cudaSuccess) err = cudaIpcGetMemHandle(&h, c->local);
    if (err != cudaSuccess) {
        if (c->local) cudaFree(c->local);
        delete c;
        cudaGetLastError();
        return (int)err;
    }
    memset(handle_out, 0, GVCNN_COMM_HANDLE_BYTES);
    memcpy(handle_out, &h, sizeof(h));
    c->peers.buf[rank] = c->local;
    *out = c;
    return 0;
}

int gvcnn_comm_connect(gvcnn_comm *c, const void *all_handles)
{
    if (!c || !all_handles) return GVCNN_E_BAD_ARG;
    if (c->connected) return 0;
    const char *hs = static_cast<const char *>(all_handles);
    for (int r = 0; r < c->world; ++r) {
        if (r == c->rank) continue;
        cudaIpcMemHandle_t h;
        memcpy(&h, hs + (size_t)r * GVCNN_COMM_HANDLE_BYTES, sizeof(h));
        void *p = nullptr;
        const cudaError_t err = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (err != cudaSuccess) {
            cudaGetLastError();
            return (int)err;
        }
        c->peers.buf[r] = static_cast<CommBuf *>(p);
    }
    c->connected = true;
    return 0;
}

int gvcnn_comm_allreduce_scaled_f32(void *comm, float *data_dev, int n, float scale, void *stream)
{
    gvcnn_comm *c = static_cast<gvcnn_comm *>(comm);
    if (!c || !data_dev || n <= 0 || n > GVCNN_COMM_MAX_FLOATS) return GVCNN_E_BAD_ARG;
    if (!c->connected && c->world > 1) return GVCNN_E_BAD_ARG;
    if (reinterpret_cast<uintptr_t>(data_dev) % 4) return GVCNN_E_MISALIGNED;
    const int nchan = (n + kChanFloats - 1) / kChanFloats;
    allreduce_oneshot_kernel<<<nchan, kCommThreads, 0, static_cast<cudaStream_t>(stream)>>>(c->peers, c->rank, c->world,
                                                                                           data_dev, n, scale);
    return (int)cudaGetLastError();
}

int gvcnn_comm_allreduce_f32(void *comm, float *data_dev, int n, void *stream)
{
    return gvcnn_comm_allreduce_scaled_f32(comm, data_dev, n, 1.0f, stream);
}

int gvcnn_comm_error(gvcnn_comm *c)
{
    if (!c || !c->local) return GVCNN_E_BAD_ARG;
    uint32_t e = 0;
    const cudaError_t err = cudaMemcpy(&e, &c->local->error, sizeof(e), cudaMemcpyDeviceToHost);
    if (err != cudaSuccess) return (int)err;
    return e ? GVCNN_E_COMM_TIMEOUT : 0;
}

int gvcnn_comm_destroy(gvcnn_comm *c)
{
    if (!c) return 0;
    for (int r = 0; r < c->world; ++r)
        if (r != c->rank && c->peers.buf[r]) cudaIpcCloseMemHandle(c->peers.buf[r]);
    if (c->local) cudaFree(c->local);
    delete c;
    return 0;
}

}  // extern "C"
