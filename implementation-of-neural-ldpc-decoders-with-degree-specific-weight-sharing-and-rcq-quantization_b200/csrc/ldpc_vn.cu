// Variable-node half iteration + posterior + hard decision: vn_kernel (degree <= 8 unrolled, > 64 generic),
// vn_wide_kernel (degree 9..64, inputs staged once in shared memory) and hard_kernel (layered schedule).
#include <algorithm>

#include "ldpc_kernel_common.cuh"

namespace ldpc {

namespace {

// Hard decisions of every frame from a posterior array [n][Bp], ballot-packed like vn_kernel's.
template <typename Real>
__global__ void __launch_bounds__(kThreads) hard_kernel(const Real* __restrict__ P, uint32_t* __restrict__ hardw,
                                                         int64_t Wn, int32_t n, int64_t Bp, int nfb) {
    constexpr int V = FramesPerLane<Real>::value;
    const int fb = blockIdx.x % nfb;
    const int chunk = blockIdx.x / nfb;
    const int64_t f0 = ((int64_t)fb * blockDim.x + threadIdx.x) * V;
    if (f0 >= Bp) return;
    const int64_t warp_f0 = f0 - (int64_t)(threadIdx.x & 31) * V;
    const int64_t wbase = (warp_f0 / (32 * V)) * V;
    for (int32_t j = chunk * 16; j < min(n, (chunk + 1) * 16); ++j) {
        Pack<Real, V> x = *reinterpret_cast<const Pack<Real, V>*>(P + (int64_t)j * Bp + f0);
        bool bit[V];
#pragma unroll
        for (int v = 0; v < V; ++v) bit[v] = x.v[v] < Real(0);
        write_hard<Real, V>(hardw, Wn, j, wbase, bit, 0u);
    }
}

// ---------------------------------------------------------------------------------------------
// Variable node + posterior + hard decision (ldpc_decoder.py:123-140; neural_2d_decoder.py:194-212)
//   v2c_d = fl(llr + fl(alpha * S(c2v of the other checks, ascending check index)))
//   post  = fl(llr + S(all c2v))        -- never alpha-weighted
//   bit   = post < 0, ballot-packed: one 32-bit word holds the same variable of 32 frames.
// ---------------------------------------------------------------------------------------------
template <typename Real, bool QUANT>
__device__ __forceinline__ Real c2v_value(const void* __restrict__ c2v, int64_t idx, const float* s_lut,
                                          int lutbase, uint32_t lutmask) {
    if constexpr (QUANT) {
        return (Real)s_lut[lutbase + (static_cast<const uint8_t*>(c2v)[idx] & lutmask)];
    } else {
        return static_cast<const Real*>(c2v)[idx];
    }
}

// U consecutive variables of degree DV at once: all loads of the group are issued before the first use,
// which is what keeps enough bytes in flight for the low-degree classes (a degree-2 variable on its own
// has only 3 loads to overlap; RCQ code rows are just 128 bytes per warp).
template <typename Real, bool QUANT, bool FINAL, bool POST, int DV, int U>
__device__ __forceinline__ void vn_node_small(const VnLaunch& p, int32_t vpos, int64_t lbase, int64_t f0,
                                              uint32_t dmask, uint32_t keepw, int64_t wbase,
                                              const float* s_lut, const int (&lutbase)[FramesPerLane<Real>::value]) {
    constexpr int V = FramesPerLane<Real>::value;
    constexpr int D1 = DV > 0 ? DV : 1;
    using InT = typename CnOut<Real, QUANT>::type;
    const InT* __restrict__ c2v = static_cast<const InT*>(p.c2v);
    Real* __restrict__ v2c = static_cast<Real*>(p.v2c);
    const uint32_t lutmask = (1u << p.bc) - 1u;  // pad frames hold unwritten codes: keep the LUT index in range
    uint32_t j[U];
    uint32_t slot[U][D1];
    Pack<InT, V> cin[U][D1];
    Pack<Real, V> L[U];
    const uint32_t in_stride = (uint32_t)p.Bp * (uint32_t)sizeof(InT), real_stride = (uint32_t)p.Bp * (uint32_t)sizeof(Real);
    c2v += f0;
    v2c += f0;
    const Real* __restrict__ llr0 = static_cast<const Real*>(p.llrT) + f0;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        j[u] = (uint32_t)__ldg(p.vpos_var + vpos + u);
#pragma unroll
        for (int d = 0; d < DV; ++d) slot[u][d] = (uint32_t)__ldg(p.vslots + lbase + u * DV + d);
    }
    // a lane whose frames are all left alone (stopped, or not selected by the posterior-on-stop pass) moves no data
    const bool lane_idle = FINAL && POST && dmask == ((1u << V) - 1u);
#pragma unroll
    for (int u = 0; u < U; ++u) {
        if (!lane_idle) {
#pragma unroll
            for (int d = 0; d < DV; ++d) cin[u][d] = ld_stream<Pack<InT, V>>(row_at(c2v, slot[u][d], in_stride));
            L[u] = ld_stream<Pack<Real, V>>(row_at(llr0, j[u], real_stride));
        } else {
#pragma unroll
            for (int d = 0; d < DV; ++d) cin[u][d] = Pack<InT, V>{};
            L[u] = Pack<Real, V>{};
        }
    }
    const bool has_alpha = (p.alpha_t != nullptr) && !FINAL;
#pragma unroll
    for (int u = 0; u < U; ++u) {
        Real alpha = Real(1);
        if (has_alpha) {
            int col = p.aidx ? __ldg(p.aidx + vpos + u) : 0;
            alpha = __ldg(static_cast<const Real*>(p.alpha_t) + col);
        }
        Pack<Real, V> out[D1];
        Pack<Real, V> post;
        bool bit[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            Real c[D1];
#pragma unroll
            for (int d = 0; d < DV; ++d) {
                if constexpr (QUANT) c[d] = (Real)s_lut[lutbase[v] + (cin[u][d].v[v] & lutmask)];
                else c[d] = cin[u][d].v[v];
            }
            if constexpr (!FINAL) {
#pragma unroll
                for (int d = 0; d < DV; ++d) {
                    Real s = LibSum<Real>::template stat<(DV > 0 ? DV - 1 : 0)>([&](int i) { return c[i < d ? i : i + 1]; });
                    if (has_alpha) s = Arith<Real>::mul(alpha, s);
                    out[d].v[v] = Arith<Real>::add(L[u].v[v], s);
                }
            }
            Real tot = LibSum<Real>::template stat<DV>([&](int i) { return c[i]; });
            Real pv = (DV > 0) ? Arith<Real>::add(L[u].v[v], tot) : L[u].v[v];
            post.v[v] = pv;
            bit[v] = (pv < Real(0)) && !((dmask >> v) & 1u);
        }
        if constexpr (!FINAL) {
#pragma unroll
            for (int d = 0; d < DV; ++d) st_stream<Pack<Real, V>>(row_at(v2c, slot[u][d], real_stride), out[d]);
        }
        // forward(): the posterior of a frame is the one of the iteration it stops at -- running frames refresh
        // their entry every iteration, stopped frames keep theirs
        if constexpr (POST) store_masked<Real, V>(row_at(static_cast<Real*>(p.postT) + f0, j[u], real_stride), post, dmask);
        // (the posterior-on-stop pass leaves the decisions alone: they are in place since the frame's last iteration)
        if (!(FINAL && POST) || p.post_iter == 0) write_hard<Real, V>(p.hardw, p.Wn, j[u], wbase, bit, keepw);
    }
}

// All variables of one work item, in groups of U (remainder one by one).
template <typename Real, bool QUANT, bool FINAL, bool POST, int DV>
__device__ __forceinline__ void vn_item_small(const VnLaunch& p, const WorkItem& it, int64_t f0, uint32_t dmask,
                                              uint32_t keepw, int64_t wbase, const float* s_lut,
                                              const int (&lutbase)[FramesPerLane<Real>::value]) {
    // byte-wide RCQ code rows need more rows in flight than 16-byte float rows
    constexpr int U = QUANT ? ((DV <= 2) ? LDPC_VN_UQ_LO : ((DV <= 4) ? LDPC_VN_UQ_MID : LDPC_VN_UQ_HI))
                            : ((DV <= 2) ? LDPC_VN_UF_LO : ((DV <= 4) ? LDPC_VN_UF_MID : 1));
    int64_t lbase = it.first_slot;
    int32_t vpos = it.first_node;
    int c = 0;
    if constexpr (U > 1) {
        for (; c + U <= it.count; c += U, lbase += U * DV, vpos += U)
            vn_node_small<Real, QUANT, FINAL, POST, DV, U>(p, vpos, lbase, f0, dmask, keepw, wbase, s_lut, lutbase);
    }
    for (; c < it.count; ++c, lbase += DV, ++vpos)
        vn_node_small<Real, QUANT, FINAL, POST, DV, 1>(p, vpos, lbase, f0, dmask, keepw, wbase, s_lut, lutbase);
}

template <typename Real, bool QUANT, bool FINAL, bool POST>
__device__ void vn_node_generic(const VnLaunch& p, int32_t vpos, int64_t lbase, int dv, int64_t f0, uint32_t dmask,
                                uint32_t keepw, int64_t wbase, const float* s_lut, const int (&lutbase)[FramesPerLane<Real>::value]) {
    constexpr int V = FramesPerLane<Real>::value;
    Real* __restrict__ v2c = static_cast<Real*>(p.v2c);
    const int64_t j = __ldg(p.vpos_var + vpos);
    const uint32_t lutmask = (1u << p.bc) - 1u;
    const Pack<Real, V> L = *reinterpret_cast<const Pack<Real, V>*>(static_cast<const Real*>(p.llrT) + j * p.Bp + f0);
    const bool has_alpha = (p.alpha_t != nullptr) && !FINAL;
    Real alpha = Real(1);
    if (has_alpha) {
        int col = p.aidx ? __ldg(p.aidx + vpos) : 0;
        alpha = __ldg(static_cast<const Real*>(p.alpha_t) + col);
    }
    bool bit[V];
    Pack<Real, V> post;
#pragma unroll
    for (int v = 0; v < V; ++v) {
        auto elem = [&](int i) -> Real {
            int64_t s = __ldg(p.vslots + lbase + i);
            return c2v_value<Real, QUANT>(p.c2v, s * p.Bp + f0 + v, s_lut, lutbase[v], lutmask);
        };
        Real tot = LibSum<Real>::dyn(elem, dv);
        Real pv = dv > 0 ? Arith<Real>::add(L.v[v], tot) : L.v[v];
        post.v[v] = pv;
        bit[v] = (pv < Real(0)) && !((dmask >> v) & 1u);
    }
    if constexpr (!FINAL) {
        // all sums are formed from c2v before any v2c of this variable is written (separate arrays)
        for (int d = 0; d < dv; ++d) {
            Pack<Real, V> out;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                auto others = [&](int i) -> Real {
                    int64_t s = __ldg(p.vslots + lbase + (i < d ? i : i + 1));
                    return c2v_value<Real, QUANT>(p.c2v, s * p.Bp + f0 + v, s_lut, lutbase[v], lutmask);
                };
                Real s = LibSum<Real>::dyn(others, dv - 1);
                if (has_alpha) s = Arith<Real>::mul(alpha, s);
                out.v[v] = Arith<Real>::add(L.v[v], s);
            }
            int64_t sd = __ldg(p.vslots + lbase + d);
            st_stream<Pack<Real, V>>(v2c + sd * p.Bp + f0, out);
        }
    }
    if constexpr (POST) store_masked<Real, V>(static_cast<Real*>(p.postT) + j * p.Bp + f0, post, dmask);
    if (!(FINAL && POST) || p.post_iter == 0) write_hard<Real, V>(p.hardw, p.Wn, j, wbase, bit, keepw);
}

// FINAL: iteration T-1 (the dead v2c update is not written).  POST: forward()'s posterior output -- running
// frames write their posterior row entries every iteration, so a stopped frame's entry is the posterior of the
// iteration it stopped at; messages and decisions of stopped frames are never touched again either way.
template <typename Real, bool QUANT, bool FINAL, bool POST>
__global__ void __launch_bounds__(kThreads, sizeof(Real) == 4 ? LDPC_VN_F32_MINCTAS : 3) vn_kernel(const VnLaunch p, const int nfb,
                                                                                                   const int item0, const int item1,
                                                                                                   const int item_stride) {
    constexpr int V = FramesPerLane<Real>::value;
    extern __shared__ float s_lut[];
    if (QUANT) {
        const int nl = p.n_quant << p.bc;
        for (int i = threadIdx.x; i < nl; i += blockDim.x) s_lut[i] = p.lut[i];
        __syncthreads();
    }
    const int fb = blockIdx.x % nfb;
    const int64_t f0 = ((int64_t)fb * blockDim.x + threadIdx.x) * V;
    if (f0 >= p.Bp) return;
    const uint32_t dmask = vn_frame_mask<V>(p.done, p.iters, (FINAL && POST) ? p.post_iter : 0, f0);
    if (__all_sync(0xffffffffu, dmask == ((1u << V) - 1u))) return;
    const uint32_t keepw = keep_word<V>(dmask);
    int lutbase[V];
#pragma unroll
    for (int v = 0; v < V; ++v) lutbase[v] = QUANT ? (p.q_now << p.bc) : 0;
    const int64_t warp_f0 = f0 - (int64_t)(threadIdx.x & 31) * V;
    const int64_t wbase = (warp_f0 / (32 * V)) * V;
    // one work item per CTA (item_stride == number of items of the launch), except in the posterior-on-stop pass,
    // whose few CTAs per frame block walk the items (most of its warps have left above)
    for (int item_id = item0 + blockIdx.x / nfb; item_id < item1; item_id += item_stride) {
    const WorkItem it = p.items[item_id];
    int64_t lbase = it.first_slot;
    int32_t vpos = it.first_node;
#define LDPC_VN_CASE(D)                                                                               \
    case D:                                                                                           \
        vn_item_small<Real, QUANT, FINAL, POST, D>(p, it, f0, dmask, keepw, wbase, s_lut, lutbase);   \
        break;
    switch (it.deg) {
        LDPC_VN_CASE(0)
        LDPC_VN_CASE(1)
        LDPC_VN_CASE(2)
        LDPC_VN_CASE(3)
        LDPC_VN_CASE(4)
        LDPC_VN_CASE(5)
        LDPC_VN_CASE(6)
        LDPC_VN_CASE(7)
        LDPC_VN_CASE(8)
        default:
            for (int c = 0; c < it.count; ++c, lbase += it.deg, ++vpos)
                vn_node_generic<Real, QUANT, FINAL, POST>(p, vpos, lbase, it.deg, f0, dmask, keepw, wbase, s_lut, lutbase);
    }
#undef LDPC_VN_CASE
    }
}

// ---------------------------------------------------------------------------------------------
// Variables of degree 9..64 (`vn_wide_kernel`).  Every leave-one-out sum has its own summation order (the
// library orders depend on the positions), so a node of degree dv needs dv * (dv - 1) element reads; taking
// them from global memory again (the generic path) ran at 0.09 of the HBM roofline.  Here a thread copies
// the dv message segments of its frames ONCE, with per-thread async copies (cp.async, no register staging:
// all dv loads are in flight together), into its own column of a shared-memory stage, and the sums read that
// column (LDS.128 per element for the four frames).  A column belongs to one thread: no barriers.
// ---------------------------------------------------------------------------------------------
constexpr int kVnWideThreads = 128;
constexpr int kPostOnlyGroups = 32;   // CTAs per frame block in the posterior-on-stop pass

template <int BYTES>
__device__ __forceinline__ void cp_async_own(void* smem_dst, const void* gmem_src) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async size");
    if constexpr (BYTES == 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem_dst)), "l"(gmem_src) : "memory");
    else
        asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "n"(BYTES) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

// Sums of one staged variable: DVS > 0 = compile-time degree (unrolled, constant stage offsets), 0 = run-time.
template <typename Real, bool FINAL, int DVS>
__device__ __forceinline__ void vn_wide_sums(const VnLaunch& p, const Pack<Real, FramesPerLane<Real>::value>* __restrict__ s_val,
                                             int dv, const Pack<Real, FramesPerLane<Real>::value>& L, Real alpha, bool has_alpha,
                                             int64_t lbase, Real* __restrict__ v2c, uint32_t real_stride,
                                             uint32_t dmask, Pack<Real, FramesPerLane<Real>::value>& post,
                                             bool (&bit)[FramesPerLane<Real>::value]) {
    constexpr int V = FramesPerLane<Real>::value;
    using PackR = Pack<Real, V>;
    auto elem = [&](int i) -> PackR { return s_val[(size_t)i * kVnWideThreads]; };
    PackR tot;
    if constexpr (DVS > 0) tot = LibSum<Real>::template stat_pack<DVS, V>(elem);
    else tot = LibSum<Real>::template dyn_pack<V>(elem, dv);
#pragma unroll
    for (int v = 0; v < V; ++v) {
        post.v[v] = Arith<Real>::add(L.v[v], tot.v[v]);
        bit[v] = (post.v[v] < Real(0)) && !((dmask >> v) & 1u);
    }
    if constexpr (!FINAL) {
        auto emit = [&](int d, const PackR& sum) {
            PackR out;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                Real sv = sum.v[v];
                if (has_alpha) sv = Arith<Real>::mul(alpha, sv);
                out.v[v] = Arith<Real>::add(L.v[v], sv);
            }
            const uint32_t sd = (uint32_t)__ldg(p.vslots + lbase + d);
            st_stream<PackR>(row_at(v2c, sd, real_stride), out);
        };
        if constexpr (DVS > 0) {
#pragma unroll
            for (int d = 0; d < DVS; ++d) {
                auto others = [&](int i) -> PackR { return s_val[(size_t)(i < d ? i : i + 1) * kVnWideThreads]; };
                emit(d, LibSum<Real>::template stat_pack<DVS - 1, V>(others));
            }
        } else {
            for (int d = 0; d < dv; ++d) {
                auto others = [&](int i) -> PackR { return s_val[(size_t)(i < d ? i : i + 1) * kVnWideThreads]; };
                emit(d, LibSum<Real>::template dyn_pack<V>(others, dv - 1));
            }
        }
    }
}

// resident CTAs per SM the staged kernel is compiled for (RCQ 6 -> 80 registers: 0.71 -> 0.85 of the
// roofline on a dv-12 code; float 5 -> 96 registers: 0.91 -> 0.97)
#ifndef LDPC_VNW_Q_MINCTAS
#define LDPC_VNW_Q_MINCTAS 6
#endif
#ifndef LDPC_VNW_F_MINCTAS
#define LDPC_VNW_F_MINCTAS 5
#endif
template <typename Real, bool QUANT, bool FINAL, bool POST>
__global__ void __launch_bounds__(kVnWideThreads, QUANT ? LDPC_VNW_Q_MINCTAS : LDPC_VNW_F_MINCTAS) vn_wide_kernel(const VnLaunch p, const int nfb, const int item0,
                                                                  const int item1, const int item_stride, const int stage_rows) {
    constexpr int V = FramesPerLane<Real>::value;
    using InT = typename CnOut<Real, QUANT>::type;
    using PackR = Pack<Real, V>;
    using PackIn = Pack<InT, V>;
    extern __shared__ __align__(16) unsigned char vn_stage[];
    // stage: values [stage_rows][threads] PackR, then (QUANT) raw codes [stage_rows][threads] PackIn, then the LUT
    PackR* __restrict__ s_val = reinterpret_cast<PackR*>(vn_stage) + threadIdx.x;
    PackIn* __restrict__ s_code = reinterpret_cast<PackIn*>(vn_stage + (size_t)stage_rows * kVnWideThreads * sizeof(PackR)) + threadIdx.x;
    float* s_lut = reinterpret_cast<float*>(vn_stage + (size_t)stage_rows * kVnWideThreads * (sizeof(PackR) + (QUANT ? sizeof(PackIn) : 0)));
    if (QUANT) {
        const int nl = p.n_quant << p.bc;
        for (int i = threadIdx.x; i < nl; i += blockDim.x) s_lut[i] = p.lut[i];
        __syncthreads();
    }
    const int fb = blockIdx.x % nfb;
    const int64_t f0 = ((int64_t)fb * kVnWideThreads + threadIdx.x) * V;
    if (f0 >= p.Bp) return;   // whole warps
    const uint32_t dmask = vn_frame_mask<V>(p.done, p.iters, (FINAL && POST) ? p.post_iter : 0, f0);
    if (__all_sync(0xffffffffu, dmask == ((1u << V) - 1u))) return;
    const uint32_t keepw = keep_word<V>(dmask);
    int lutbase[V];
#pragma unroll
    for (int v = 0; v < V; ++v) lutbase[v] = QUANT ? (p.q_now << p.bc) : 0;
    const uint32_t lutmask = (1u << p.bc) - 1u;
    const int64_t warp_f0 = f0 - (int64_t)(threadIdx.x & 31) * V;
    const int64_t wbase = (warp_f0 / (32 * V)) * V;
    const uint32_t in_stride = (uint32_t)p.Bp * (uint32_t)sizeof(InT), real_stride = (uint32_t)p.Bp * (uint32_t)sizeof(Real);
    const InT* __restrict__ c2v = static_cast<const InT*>(p.c2v) + f0;
    Real* __restrict__ v2c = static_cast<Real*>(p.v2c) + f0;
    const Real* __restrict__ llr0 = static_cast<const Real*>(p.llrT) + f0;
    const bool has_alpha = (p.alpha_t != nullptr) && !FINAL;
    for (int item_id = item0 + blockIdx.x / nfb; item_id < item1; item_id += item_stride) {
    const WorkItem it = p.items[item_id];
    const int dv = it.deg;
    int64_t lbase = it.first_slot;
    int32_t vpos = it.first_node;
    for (int c = 0; c < it.count; ++c, lbase += dv, ++vpos) {
        const uint32_t j = (uint32_t)__ldg(p.vpos_var + vpos);
        for (int i = 0; i < dv; ++i) {
            if (FINAL && POST && dmask == ((1u << V) - 1u)) break;   // idle lane: nothing it computes is stored
            const uint32_t slot = (uint32_t)__ldg(p.vslots + lbase + i);
            if constexpr (QUANT) cp_async_own<sizeof(PackIn)>(s_code + (size_t)i * kVnWideThreads, row_at(c2v, slot, in_stride));
            else cp_async_own<sizeof(PackR)>(s_val + (size_t)i * kVnWideThreads, row_at(c2v, slot, in_stride));
        }
        const PackR L = ld_stream<PackR>(row_at(llr0, j, real_stride));
        Real alpha = Real(1);
        if (has_alpha) alpha = __ldg(static_cast<const Real*>(p.alpha_t) + (p.aidx ? __ldg(p.aidx + vpos) : 0));
        cp_async_wait_all();
        if constexpr (QUANT) {   // decode the codes once
            for (int i = 0; i < dv; ++i) {
                const PackIn code = s_code[(size_t)i * kVnWideThreads];
                PackR val;
#pragma unroll
                for (int v = 0; v < V; ++v) val.v[v] = (Real)s_lut[lutbase[v] + (code.v[v] & lutmask)];
                s_val[(size_t)i * kVnWideThreads] = val;
            }
        }
        PackR post;
        bool bit[V];
        bool handled = false;
#define LDPC_VNW_CASE(D)                                                                                              \
    case D:                                                                                                           \
        vn_wide_sums<Real, FINAL, D>(p, s_val, D, L, alpha, has_alpha, lbase, v2c, real_stride, dmask, post, bit); \
        handled = true;                                                                                               \
        break;
        switch (dv) {
            LDPC_VNW_CASE(9)
            LDPC_VNW_CASE(10)
            LDPC_VNW_CASE(11)
            LDPC_VNW_CASE(12)
            LDPC_VNW_CASE(13)
            LDPC_VNW_CASE(14)
            LDPC_VNW_CASE(15)
            LDPC_VNW_CASE(16)
            default: break;
        }
#undef LDPC_VNW_CASE
        if (!handled) vn_wide_sums<Real, FINAL, 0>(p, s_val, dv, L, alpha, has_alpha, lbase, v2c, real_stride, dmask, post, bit);
        if constexpr (POST) store_masked<Real, V>(row_at(static_cast<Real*>(p.postT) + f0, j, real_stride), post, dmask);
        if (!(FINAL && POST) || p.post_iter == 0) write_hard<Real, V>(p.hardw, p.Wn, j, wbase, bit, keepw);
    }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------
cudaError_t launch_hard(int dtype, const void* P, uint32_t* hardw, int64_t Wn, int32_t n, int64_t Bp, cudaStream_t stream) {
    const int V = dtype == 0 ? 4 : 2;
    const int threads = threads_for(Bp, V);
    const int64_t nfb = (Bp / V + threads - 1) / threads;
    const int64_t grid = nfb * ((n + 15) / 16);
    if (dtype == 0) hard_kernel<float><<<(unsigned)grid, threads, 0, stream>>>(static_cast<const float*>(P), hardw, Wn, n, Bp, (int)nfb);
    else hard_kernel<double><<<(unsigned)grid, threads, 0, stream>>>(static_cast<const double*>(P), hardw, Wn, n, Bp, (int)nfb);
    return cudaGetLastError();
}

namespace {

template <typename Real, bool QUANT>
cudaError_t launch_vn_range(const VnLaunch& p, int item0, int item1, bool wide, cudaStream_t stream) {
    if (item1 <= item0) return cudaSuccess;
    constexpr int V = FramesPerLane<Real>::value;
    // posterior-on-stop pass: nearly all of its warps leave at once, so a few CTAs per frame block walk the items
    const bool post_only = p.final_pass && p.postT && p.post_iter > 0;
    if (wide) {
        using InT = typename CnOut<Real, QUANT>::type;
        const int rows = p.wide_max_deg;
        const size_t smem = (size_t)rows * kVnWideThreads * (sizeof(Pack<Real, V>) + (QUANT ? sizeof(Pack<InT, V>) : 0)) +
                            (QUANT ? sizeof(float) * ((size_t)p.n_quant << p.bc) : 0);
        const int64_t nfb = (p.Bp + (int64_t)kVnWideThreads * V - 1) / ((int64_t)kVnWideThreads * V);
        const int groups = post_only ? std::min(item1 - item0, kPostOnlyGroups) : item1 - item0;
        const int64_t grid = nfb * groups;
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
#define LDPC_VNW(FINAL, POST)                                                                                          \
    do {                                                                                                               \
        cudaError_t e_ = cudaFuncSetAttribute(vn_wide_kernel<Real, QUANT, FINAL, POST>,                                \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
        if (e_ != cudaSuccess) return e_;                                                                              \
        vn_wide_kernel<Real, QUANT, FINAL, POST><<<(unsigned)grid, kVnWideThreads, smem, stream>>>(p, (int)nfb, item0, item1, groups, rows);   \
    } while (0)
        if (p.final_pass) { if (p.postT) LDPC_VNW(true, true); else LDPC_VNW(true, false); }
        else { if (p.postT) LDPC_VNW(false, true); else LDPC_VNW(false, false); }
#undef LDPC_VNW
    } else {
        const int threads = threads_for(p.Bp, V);
        const int64_t nfb = (p.Bp / V + threads - 1) / threads;
        const int groups = post_only ? std::min(item1 - item0, kPostOnlyGroups) : item1 - item0;
        const int64_t grid = nfb * groups;
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        const size_t smem = p.bc ? sizeof(float) * ((size_t)p.n_quant << p.bc) : 0;
        const unsigned g = (unsigned)grid;
        const int nf = (int)nfb;
        if (p.final_pass) {
            if (p.postT) vn_kernel<Real, QUANT, true, true><<<g, threads, smem, stream>>>(p, nf, item0, item1, groups);
            else vn_kernel<Real, QUANT, true, false><<<g, threads, smem, stream>>>(p, nf, item0, item1, groups);
        } else {
            if (p.postT) vn_kernel<Real, QUANT, false, true><<<g, threads, smem, stream>>>(p, nf, item0, item1, groups);
            else vn_kernel<Real, QUANT, false, false><<<g, threads, smem, stream>>>(p, nf, item0, item1, groups);
        }
    }
    return cudaGetLastError();
}

template <typename Real, bool QUANT>
cudaError_t launch_vn_all(const VnLaunch& p, cudaStream_t stream) {
    // items are sorted by degree: [0, wide0) degree <= 8, [wide0, wide1) degree 9..64 (shared-memory stage), rest > 64
    const int wide0 = p.wide_stage ? p.items_wide_begin : p.n_items, wide1 = p.wide_stage ? p.items_wide_end : p.n_items;
    cudaError_t e = launch_vn_range<Real, QUANT>(p, 0, wide0, false, stream);
    if (e == cudaSuccess) e = launch_vn_range<Real, QUANT>(p, wide0, wide1, true, stream);
    if (e == cudaSuccess) e = launch_vn_range<Real, QUANT>(p, wide1, p.n_items, false, stream);
    return e;
}

}  // namespace

cudaError_t launch_vn(int dtype, const VnLaunch& p, cudaStream_t stream) {
    if (p.n_items == 0) return cudaSuccess;
    if (dtype == 0) return p.bc ? launch_vn_all<float, true>(p, stream) : launch_vn_all<float, false>(p, stream);
    return launch_vn_all<double, false>(p, stream);
}

}  // namespace ldpc
