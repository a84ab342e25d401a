// Device side of gvcnn_comm (comm.cu): the receive buffer every rank exposes to its peers over NVLink, and the
// low-latency ("LL") exchange primitives shared by the stand-alone all-reduce kernel and the fused
// batch-mean + exchange + bin kernel (score.cu).
//
// LL protocol: a float travels as ONE 8-byte store {value, sequence number}.  8-byte stores are single-copy atomic
// all the way through NVLink, so the receiver needs no fence and no separate flag: it polls the slot itself until
// the sequence number is the one it expects, and the value it read alongside is the matching one.  Compared with
// "store data, __threadfence_system(), store flag" this removes one full NVLink round trip from every exchange.
// Slots are double-buffered by sequence parity (see comm.cu for why that needs no second barrier).
#pragma once

#include "common.cuh"

namespace gvcnn {

constexpr int kCommMaxWorld = GVCNN_COMM_MAX_WORLD;
constexpr int kCommChanFloats = 2048;                                       // one CTA's share of a vector
constexpr int kCommMaxChan = GVCNN_COMM_MAX_FLOATS / kCommChanFloats;       // 8
constexpr unsigned long long kCommTimeoutNs = 4000000000ull;                // 4 s: a peer that never arrives

struct CommBuf {
    uint2 ll[2][kCommMaxWorld][GVCNN_COMM_MAX_FLOATS];   // [phase][source rank][element] = {float bits, sequence}
    uint2 llv[2][kCommMaxWorld][GVCNN_MAX_VIEWS];        // the fused batch-mean kernel's own lane: one element per view
    uint32_t seq[kCommMaxChan][8];                       // local: last sequence number used per channel (32-byte pads)
    uint32_t seqv[GVCNN_MAX_VIEWS];                      // local: last sequence number used per view (fused kernel)
    uint32_t error;                                      // local: set when a wait timed out
};

struct CommPeers {
    CommBuf *buf[kCommMaxWorld];
};

__device__ __forceinline__ unsigned long long comm_timer_ns()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}

// One LL element = a 64-bit word {sequence << 32 | float bits}; written and read as 64-bit scalars (or as the two
// 64-bit elements of a 16-byte vector access - each element is then still one single-copy-atomic 8-byte access, and
// tearing BETWEEN the two elements is harmless because each carries its own sequence number).
__device__ __forceinline__ unsigned long long ll_pack(float a, uint32_t seq)
{
    return ((unsigned long long)seq << 32) | (unsigned long long)__float_as_uint(a);
}
__device__ __forceinline__ void ll_store2(uint2 *dst, float a, float b, uint32_t seq)
{
    asm volatile("st.volatile.global.v2.u64 [%0], {%1, %2};" ::"l"(dst), "l"(ll_pack(a, seq)), "l"(ll_pack(b, seq)) : "memory");
}
__device__ __forceinline__ void ll_store1(uint2 *dst, float a, uint32_t seq)
{
    asm volatile("st.volatile.global.u64 [%0], %1;" ::"l"(dst), "l"(ll_pack(a, seq)) : "memory");
}
// Polls two adjacent LL elements until both carry `seq`; false on timeout.
__device__ __forceinline__ bool ll_wait2(const uint2 *src, uint32_t seq, float &a, float &b, unsigned long long t0)
{
    unsigned long long w0, w1;
    for (;;) {
        asm volatile("ld.volatile.global.v2.u64 {%0, %1}, [%2];" : "=l"(w0), "=l"(w1) : "l"(src) : "memory");
        if ((uint32_t)(w0 >> 32) == seq && (uint32_t)(w1 >> 32) == seq) break;
        if (comm_timer_ns() - t0 > kCommTimeoutNs) return false;
    }
    a = __uint_as_float((uint32_t)w0);
    b = __uint_as_float((uint32_t)w1);
    return true;
}
__device__ __forceinline__ bool ll_wait1(const uint2 *src, uint32_t seq, float &a, unsigned long long t0)
{
    unsigned long long w0;
    for (;;) {
        asm volatile("ld.volatile.global.u64 %0, [%1];" : "=l"(w0) : "l"(src) : "memory");
        if ((uint32_t)(w0 >> 32) == seq) break;
        if (comm_timer_ns() - t0 > kCommTimeoutNs) return false;
    }
    a = __uint_as_float((uint32_t)w0);
    return true;
}

bool comm_device_view(void *comm, const CommPeers **peers, int *rank, int *world);

// launcher of the fused literal-mode tail (score.cu): column sums of x -> [exchange over `peers`] -> mean -> score -> bin
int launch_batch_mean_bin_fused(const float *x, float *xsum, float *x_mean, float *scores, int32_t *bins, int32_t *flags,
                                int32_t *status, int B, int V, int G, int multiplier, int edge_ulps, int clamp,
                                float denom, const CommPeers *peers, int rank, int world, cudaStream_t st);

int batch_score_tail(const float *x, float *xsum, float *x_mean, float *scores, int32_t *bins, int32_t *flags,
                     int32_t *status, int B, int V, int G, int multiplier, int edge_ulps, int clamp,
                     int64_t global_count, gvcnn_exchange_fn exchange, void *exchange_user, cudaStream_t st);

}  // namespace gvcnn
