// Device helpers shared by the kernel translation units (check node, variable node, everything else).
// Internal linkage: every translation unit gets its own copy.
#pragma once
#include <type_traits>

#include "ldpc_device.cuh"
#include "ldpc_internal.h"

namespace ldpc {

namespace {

constexpr int kThreads = 256;

// variables per group in vn_item_small (loads of a whole group are issued before the first use)
#ifndef LDPC_VN_UQ_LO
#define LDPC_VN_UQ_LO 4
#endif
#ifndef LDPC_VN_UQ_MID
#define LDPC_VN_UQ_MID 2
#endif
#ifndef LDPC_VN_UQ_HI
#define LDPC_VN_UQ_HI 1
#endif
#ifndef LDPC_VN_UF_LO
#define LDPC_VN_UF_LO 2
#endif
// resident CTAs per SM the float32 variable-node kernels are compiled for (4 -> at most 64 registers)
#ifndef LDPC_VN_F32_MINCTAS
#define LDPC_VN_F32_MINCTAS 4
#endif
#ifndef LDPC_VN_UF_MID
#define LDPC_VN_UF_MID 1
#endif

template <int V>
__device__ __forceinline__ uint32_t load_done_mask(const uint8_t* __restrict__ done, int64_t f0) {
    uint32_t m = 0;
    if constexpr (V == 4) {
        uint32_t w = __ldg(reinterpret_cast<const uint32_t*>(done + f0));
#pragma unroll
        for (int v = 0; v < 4; ++v) m |= ((w >> (8 * v)) & 0xffu) ? (1u << v) : 0u;
    } else {
        uint16_t w = __ldg(reinterpret_cast<const uint16_t*>(done + f0));
#pragma unroll
        for (int v = 0; v < 2; ++v) m |= ((w >> (8 * v)) & 0xffu) ? (1u << v) : 0u;
    }
    return m;
}

// Row addressing: `base` already points at this lane's first frame of row 0; one IMAD.WIDE.U32 per row
// (row index and row stride in bytes both fit 32 bits: the host caps Bp at 2^28 frames).
template <typename T>
__device__ __forceinline__ T* row_at(T* base, uint32_t row, uint32_t stride_bytes) {
    using C = typename std::conditional<std::is_const<T>::value, const char, char>::type;
    return reinterpret_cast<T*>(reinterpret_cast<C*>(base) + (uint64_t)row * stride_bytes);
}

// Store V frames of one row; frames whose bit is set in `keep` hold on to their old (frozen) value: the row
// segment is read, merged and written back whole, so the access stays one vector transaction per lane.
template <typename T, int V>
__device__ __forceinline__ void store_masked(T* __restrict__ rowptr, Pack<T, V> val, uint32_t keep) {
    if (keep == (1u << V) - 1u) return;   // nothing of this lane changes
    if (keep != 0) {
        const Pack<T, V> old = *reinterpret_cast<const Pack<T, V>*>(rowptr);
#pragma unroll
        for (int v = 0; v < V; ++v)
            if ((keep >> v) & 1u) val.v[v] = old.v[v];
    }
    st_stream<Pack<T, V>>(rowptr, val);
}

// Frames of this lane a variable-node pass works on, as a "leave alone" mask.  Normal passes: the stopped frames.
// Posterior-on-stop pass (post_iter > 0): everything except the frames that stopped exactly at iteration post_iter.
template <int V>
__device__ __forceinline__ uint32_t vn_frame_mask(const uint8_t* __restrict__ done, const int32_t* __restrict__ iters,
                                                  int post_iter, int64_t f0) {
    uint32_t dmask = load_done_mask<V>(done, f0);
    if (post_iter > 0) {
        uint32_t sel = 0;
#pragma unroll
        for (int v = 0; v < V; ++v)
            if (((dmask >> v) & 1u) && __ldg(iters + f0 + v) == post_iter) sel |= 1u << v;
        dmask = ~sel & ((1u << V) - 1u);
    }
    return dmask;
}

// message type of the check->variable array
template <typename Real, bool QUANT>
struct CnOut { using type = Real; };
template <typename Real>
struct CnOut<Real, true> { using type = uint8_t; };

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// One 32-bit word per (variable, V-th frame of 32 lanes): bit = lane.  `keepw` (lanes < V: the word of this
// lane's frame slot) marks frames that have stopped: their decisions of the iteration they stopped at stay in
// place (read-modify-write), so no later pass has to recompute them.
template <typename Real, int V>
__device__ __forceinline__ void write_hard(uint32_t* __restrict__ hardw, int64_t Wn, int64_t j, int64_t wbase,
                                           const bool (&bit)[V], uint32_t keepw) {
    uint32_t words[V];
#pragma unroll
    for (int v = 0; v < V; ++v) words[v] = __ballot_sync(0xffffffffu, bit[v]);
    const int lane = threadIdx.x & 31;
    if (lane < V) {
        uint32_t w = words[0];
#pragma unroll
        for (int v = 1; v < V; ++v)
            if (lane == v) w = words[v];
        uint32_t* ptr = hardw + j * Wn + wbase + lane;
        if (keepw) w = (w & ~keepw) | (*ptr & keepw);
        *ptr = w;
    }
}

// keep-word of this lane (see write_hard) from the per-lane done masks of the warp
template <int V>
__device__ __forceinline__ uint32_t keep_word(uint32_t dmask) {
    uint32_t kw = 0;
    const int lane = threadIdx.x & 31;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        const uint32_t b = __ballot_sync(0xffffffffu, (dmask >> v) & 1u);
        if (lane == v) kw = b;
    }
    return kw;
}

__host__ __device__ __forceinline__ void frame_to_wordbit(int64_t f, int V, int64_t& w, int& bit) {
    int64_t g = f / (32 * V);
    int r = (int)(f % (32 * V));
    bit = r / V;
    w = g * V + (r % V);
}
__host__ __device__ __forceinline__ int64_t wordbit_to_frame(int64_t w, int bit, int V) {
    return (w / V) * (32 * V) + (int64_t)bit * V + (w % V);
}

inline int threads_for(int64_t Bp, int V) {
    int64_t lanes = Bp / V;
    int t = (int)(lanes < kThreads ? lanes : kThreads);
    t = (t + 31) / 32 * 32;
    return t < 32 ? 32 : t;
}

}  // namespace

}  // namespace ldpc
