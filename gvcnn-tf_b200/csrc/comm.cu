// One-shot all-reduce over NVLink peer memory for the two tiny exchanges of the path:
//   (1) the flat parameter-gradient bucket of the training step, V * (C_raw + 1) floats = 49 KB at V = 12 - the
//       B200-native form of `nccl_ops.all_sum(grads)` + `* 1/K` at utils/_train_helper.py:17-31;
//   (2) the V per-view partial sums of the literal batch mean before binning (nets/model.py:146 on a sharded batch).
// Both are latency-bound (an NCCL all-reduce of this size costs 18-33 us on 2-8 B200s against a 100-190 us step), so
// they get ONE kernel per rank, no rendezvous on the host, and a protocol with a single NVLink hop:
//   push   every rank stores its vector into its own slot of EVERY peer's receive buffer (cudaIpc-mapped peer memory:
//          NVLink stores) as low-latency elements {value, sequence number} - one 8-byte single-copy-atomic store per
//          float (comm_dev.cuh), so no fence and no separate flag follow;
//   wait   poll the LOCAL slots until every element carries this call's sequence number;
//   reduce add the K slots in rank order 0..K-1 - the same order on every rank, so all ranks get bit-identical sums
//          (every rank must derive the same bins) - scale, write in place.
// (Round 2's first version pushed plain data, fenced with __threadfence_system() and then released a flag: 10.5 us per
// 12-float exchange between 2 GPUs against 19.2 us for NCCL; the fence + flag cost a second NVLink round trip.)
// Channels: the vector is cut into fixed 2048-float channels, one CTA each, with private slots and sequence counters,
// so CTAs never synchronise with each other.  Slots are double-buffered by sequence parity: a rank can only push call
// s+2 after it has finished call s+1, i.e. after every peer has pushed s+1, i.e. after every peer has finished READING
// call s - no second barrier.  The sequence counters live in device memory and are advanced by the kernel, so a
// captured launch replays correctly from a CUDA graph.  The fused batch-mean kernel of score.cu speaks the same
// protocol on channel 0 (it IS an all-reduce of V floats, issued from inside the kernel that needs the result).
// A rank that never arrives would hang its peers: the wait gives up after kCommTimeoutNs and raises an error word the
// host can read (gvcnn_comm_error), instead of hanging the GPU.
#include <cstring>
#include <new>

#include "comm_dev.cuh"

using namespace gvcnn;

namespace {
constexpr int kCommThreads = 512;  // one float4 (4 LL elements) each

__global__ void __launch_bounds__(kCommThreads)
allreduce_oneshot_kernel(const CommPeers peers, const int rank, const int world, float *__restrict__ data, const int n,
                         const float scale)
{
    const int c = blockIdx.x;
    const int lo = c * kCommChanFloats;
    const int cnt = min(n - lo, kCommChanFloats);        // floats of this channel (> 0 by the grid size)
    CommBuf *mine = peers.buf[rank];
    const uint32_t seq = mine->seq[c][0] + 1u;           // every thread reads it before thread 0 advances it below
    const int phase = (int)(seq & 1u);
    const int i = threadIdx.x * 4;                       // this thread's 4 floats of the channel
    const bool vec_ok = (i + 3 < cnt) && ((reinterpret_cast<uintptr_t>(data) & 15) == 0);
    float v[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    if (vec_ok) {
        const float4 t = *reinterpret_cast<const float4 *>(data + lo + i);
        v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
        for (int j = 0; j < 4; ++j)
            if (i + j < cnt) v[j] = data[lo + i + j];
    }
    __syncthreads();                                     // all reads of seq[c] are done
    if (threadIdx.x == 0) mine->seq[c][0] = seq;
    if (i < cnt) {
        // ---- push: my slot at every peer (own buffer included), rotated so the ranks do not all hit one link first
        for (int d = 0; d < world; ++d) {
            const int p = (rank + 1 + d) % world;
            uint2 *dst = &peers.buf[p]->ll[phase][rank][lo + i];  // 32-byte aligned; slots are padded to 4 elements
            ll_store2(dst, v[0], v[1], seq);
            ll_store2(dst + 2, v[2], v[3], seq);
        }
        // ---- wait + reduce in rank order (identical on every rank), scale, store in place.  All K slots are polled
        // in ONE pass of independent loads per round (the first version polled rank after rank: 2 K dependent L2 round
        // trips, 6 us of the 14.6 us a 49 KB exchange took between 8 GPUs)
        const unsigned long long t0 = comm_timer_ns();
        unsigned long long w[kCommMaxWorld][4];
        bool ok = true;
        for (;;) {
            bool all = true;
#pragma unroll
            for (int r = 0; r < kCommMaxWorld; ++r) {
                if (r < world) {
                    const uint2 *src = &mine->ll[phase][r][lo + i];
                    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[r][0]), "=l"(w[r][1]) : "l"(src) : "memory");
                    asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w[r][2]), "=l"(w[r][3]) : "l"(src + 2) : "memory");
                }
            }
#pragma unroll
            for (int r = 0; r < kCommMaxWorld; ++r) {
                if (r < world) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) all = all && ((uint32_t)(w[r][j] >> 32) == seq);
                }
            }
            if (all) break;
            if (comm_timer_ns() - t0 > kCommTimeoutNs) {
                ok = false;
                break;
            }
        }
        float acc[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[j] = __uint_as_float((uint32_t)w[0][j]);
#pragma unroll
        for (int r = 1; r < kCommMaxWorld; ++r) {
            if (r < world) {
#pragma unroll
                for (int j = 0; j < 4; ++j) acc[j] = __fadd_rn(acc[j], __uint_as_float((uint32_t)w[r][j]));
            }
        }
        if (!ok) atomicExch(&mine->error, 1u);
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
    static_assert(kCommChanFloats == kCommThreads * 4, "one float4 per thread per channel");
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
    const int nchan = (n + kCommChanFloats - 1) / kCommChanFloats;
    allreduce_oneshot_kernel<<<nchan, kCommThreads, 0, static_cast<cudaStream_t>(stream)>>>(c->peers, c->rank, c->world,
                                                                                           data_dev, n, scale);
    return (int)cudaGetLastError();
}

int gvcnn_comm_allreduce_f32(void *comm, float *data_dev, int n, void *stream)
{
    return gvcnn_comm_allreduce_scaled_f32(comm, data_dev, n, 1.0f, stream);
}

}  // extern "C"

namespace gvcnn {
// for the fused kernels of other translation units: the device view of a connected communicator
bool comm_device_view(void *comm, const CommPeers **peers, int *rank, int *world)
{
    gvcnn_comm *c = static_cast<gvcnn_comm *>(comm);
    if (!c || (!c->connected && c->world > 1)) return false;
    *peers = &c->peers;
    *rank = c->rank;
    *world = c->world;
    return true;
}
}  // namespace gvcnn

extern "C" {

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
