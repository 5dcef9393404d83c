// Check-node half iteration: cn_kernel (degree <= 8 unrolled in registers, > 64 streaming), cn_wide_kernel
// (degree 9..64, bulk-async shared-memory row ring), cn_offset_kernel (offset min-sum rule) and the layered RCQ
// schedule.  Layout and conventions: ldpc_kernel_common.cuh / DESIGN.md sections 3-4.
#include <algorithm>

#include "ldpc_cn_common.cuh"

namespace ldpc {

namespace {

// ---------------------------------------------------------------------------------------------
// Check node (ldpc_decoder.py:91-120; neural_2d_decoder.py:161-191; rcq_decoder.py:211-246, :526-563)
//
// Per frame: m1 = min |x|, k0 = its first index, m2 = min over the others; for edge k
//   raw = (k == k0) ? m2 : m1,   c2v = fl(beta_k * raw) with the sign of prod_{k' != k} sign(x_k').
// The product of the other signs is applied as an XOR of IEEE sign bits: whenever an input is +-0 the
// magnitudes force every affected output to +-0 (appendix A2), so three-valued sign() never shows.
// The same expression covers (beta*raw)*sp [N-MS] and (beta*sp)*raw [W-RCQ]: sp = +-1 is exact.
// ---------------------------------------------------------------------------------------------

// ---------------------------------------------------------------------------------------------
// Byte-parallel (SWAR) assembly of RCQ codes: one 32-bit word holds the codes of a lane's four frames, so
// selecting between the check's two magnitudes, applying the sign and packing the store cost a few
// LOP3/PRMT per EDGE instead of per edge and frame.  prmt's sign-replicate mode (selector nibble | 8) turns
// the top bit of a byte into 0x00 / 0xFF.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}
// byte v of the result = low byte of w[v]
__device__ __forceinline__ uint32_t pack_low_bytes(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    return prmt(prmt(w0, w1, 0x0040u), prmt(w2, w3, 0x0040u), 0x5410u);
}
// byte v of the result = byte `b` (0..3) of w[v]
__device__ __forceinline__ uint32_t pack_bytes_at(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3, uint32_t b) {
    const uint32_t sel = b | ((b + 4u) << 4);
    return prmt(prmt(w0, w1, sel), prmt(w2, w3, sel), 0x5410u);
}
// byte v of the result = 0xFF if bit 31 of w[v] is set, else 0x00
__device__ __forceinline__ uint32_t pack_sign_masks(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t w3) {
    return prmt(prmt(w0, w1, 0x00FBu), prmt(w2, w3, 0x00FBu), 0x5410u);
}
// every byte -> 0xFF if its top bit is set, else 0x00
__device__ __forceinline__ uint32_t spread_byte_signs(uint32_t w) { return prmt(w, 0u, 0xBA98u); }

// The four frames' CheckOut of one check, byte-packed (RCQ, one beta per check).
struct CheckOut4 {
    uint32_t CA, CB;   // magnitude index of the non-minimum / minimum edges
    uint32_t MA, MB;   // sign-bit mask (1 << (bc-1), or 0 where the value is +-0)
    uint32_t PAR;      // 0xFF where the product of ALL input signs (and beta's) is negative
    __device__ __forceinline__ void pack(const CheckOut<float, true> (&co)[4]) {
        CA = pack_low_bytes(co[0].ia, co[1].ia, co[2].ia, co[3].ia);
        CB = pack_low_bytes(co[0].ib, co[1].ib, co[2].ib, co[3].ib);
        MA = pack_low_bytes(co[0].ma, co[1].ma, co[2].ma, co[3].ma);
        MB = pack_low_bytes(co[0].mb, co[1].mb, co[2].mb, co[3].mb);
        PAR = pack_sign_masks(co[0].par, co[1].par, co[2].par, co[3].par);
    }
    // min4: 0xFF where this edge carries the frame's minimum; neg4: 0xFF where the output is negative
    __device__ __forceinline__ uint32_t emit(uint32_t min4, uint32_t neg4) const {
        const uint32_t idx = (CB & min4) | (CA & ~min4);
        const uint32_t msk = (MB & min4) | (MA & ~min4);
        return idx | (neg4 & msk);
    }
};

template <typename Real, bool QUANT, int NTH, int DC>
__device__ __forceinline__ void cn_check_small(const CnLaunch& p, int64_t slot0, int64_t f0,
                                               const Quantizer<NTH>& qz) {
    constexpr int V = FramesPerLane<Real>::value;
    using OutT = typename CnOut<Real, QUANT>::type;
    const Real* __restrict__ src = static_cast<const Real*>(p.src);
    OutT* __restrict__ dst = static_cast<OutT*>(p.dst);
    Pack<Real, V> x[DC];
    const bool has_beta = p.beta_t != nullptr;
    const uint32_t in_stride = (uint32_t)p.Bp * (uint32_t)sizeof(Real), out_stride = (uint32_t)p.Bp * (uint32_t)sizeof(OutT);
    src += f0;
    dst += f0;
#pragma unroll
    for (int k = 0; k < DC; ++k) {
        const uint32_t row = p.row_map ? (uint32_t)__ldg(p.row_map + slot0 + k) : (uint32_t)(slot0 + k);
        x[k] = ld_stream<Pack<Real, V>>(row_at(src, row, in_stride));
    }
    Pack<OutT, V> out[DC];
    if constexpr (QUANT && V == 4) {
        if (!p.beta_per_edge) {
            // byte-parallel output phase (see CheckOut4): per edge a few LOP3/PRMT for all four frames
            float beta = 1.f;
            if (has_beta) beta = __ldg(static_cast<const float*>(p.beta_t) + (p.bidx ? __ldg(p.bidx + slot0) : 0));
            CheckOut<float, true> co[4];
            uint32_t nm1[4];
#pragma unroll
            for (int v = 0; v < 4; ++v) {
                MinState<float, false> st;
                st.init();
#pragma unroll
                for (int k = 0; k < DC; ++k) st.push(x[k].v[v], k);
                if (DC == 1) st.m2 = st.m1;
                co[v].prepare(st.m1, st.m2, st.par, beta, has_beta, qz, p.bc);
                nm1[v] = ~__float_as_uint(st.m1);
            }
            CheckOut4 c4;
            c4.pack(co);
#pragma unroll
            for (int k = 0; k < DC; ++k) {
                uint32_t xb[4], eq[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) {
                    xb[v] = __float_as_uint(x[k].v[v]);
                    eq[v] = (xb[v] & 0x7fffffffu) + nm1[v];   // |x| - m1 - 1: negative iff |x| == m1 (|x| >= m1)
                }
                const uint32_t min4 = pack_sign_masks(eq[0], eq[1], eq[2], eq[3]);
                const uint32_t neg4 = pack_sign_masks(xb[0], xb[1], xb[2], xb[3]) ^ c4.PAR;
                *reinterpret_cast<uint32_t*>(&out[k]) = c4.emit(min4, neg4);
            }
#pragma unroll
            for (int k = 0; k < DC; ++k) st_stream<Pack<OutT, V>>(row_at(dst, (uint32_t)(slot0 + k), out_stride), out[k]);
            return;
        }
    }
    if (!p.beta_per_edge) {
        Real beta = Real(1);
        if (has_beta) beta = __ldg(static_cast<const Real*>(p.beta_t) + (p.bidx ? __ldg(p.bidx + slot0) : 0));
#pragma unroll
        for (int v = 0; v < V; ++v) {
            MinState<Real, false> st;
            st.init();
#pragma unroll
            for (int k = 0; k < DC; ++k) st.push(x[k].v[v], k);
            if (DC == 1) st.m2 = st.m1;  // ldpc_decoder.py:112-113
            CheckOut<Real, QUANT> co;
            co.prepare(st.m1, st.m2, st.par, beta, has_beta, qz, p.bc);
#pragma unroll
            for (int k = 0; k < DC; ++k)
                out[k].v[v] = co.emit(Arith<Real>::abs(x[k].v[v]) == st.m1, Arith<Real>::hi(x[k].v[v]));
        }
    } else {
        Real beta[DC];
#pragma unroll
        for (int k = 0; k < DC; ++k) beta[k] = __ldg(static_cast<const Real*>(p.beta_t) + __ldg(p.bidx + slot0 + k));
#pragma unroll
        for (int v = 0; v < V; ++v) {
            MinState<Real, false> st;
            st.init();
#pragma unroll
            for (int k = 0; k < DC; ++k) st.push(x[k].v[v], k);
            if (DC == 1) st.m2 = st.m1;
#pragma unroll
            for (int k = 0; k < DC; ++k) {
                Real raw = (Arith<Real>::abs(x[k].v[v]) == st.m1) ? st.m2 : st.m1;
                out[k].v[v] = cn_emit<Real, QUANT, NTH>(raw, beta[k], st.par ^ Arith<Real>::hi(x[k].v[v]), qz, p.bc);
            }
        }
    }
#pragma unroll
    for (int k = 0; k < DC; ++k) st_stream<Pack<OutT, V>>(row_at(dst, (uint32_t)(slot0 + k), out_stride), out[k]);
}

// Checks of degree 9..32: stream the inputs once, keeping min1/min2/first-argmin/parity and one sign
// bit per edge in a 32-bit shift register (funnel shift: one instruction per edge and frame).
template <typename Real, bool QUANT, int NTH>
__device__ void cn_check_mask32(const CnLaunch& p, int64_t slot0, int dc, int64_t f0,
                                const Quantizer<NTH>& qz) {
    constexpr int V = FramesPerLane<Real>::value;
    using OutT = typename CnOut<Real, QUANT>::type;
    const Real* __restrict__ src = static_cast<const Real*>(p.src);
    OutT* __restrict__ dst = static_cast<OutT*>(p.dst);
    const bool has_beta = p.beta_t != nullptr;
    MinState<Real, true> st[V];
    uint32_t neg[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        st[v].init();
        neg[v] = 0;
    }
#pragma unroll 6
    for (int k = 0; k < dc; ++k) {
        int64_t row = p.row_map ? (int64_t)__ldg(p.row_map + slot0 + k) : slot0 + k;
        Pack<Real, V> x = ld_stream<Pack<Real, V>>(src + row * p.Bp + f0);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            st[v].push(x.v[v], k);
            neg[v] = __funnelshift_l(Arith<Real>::hi(x.v[v]), neg[v], 1);  // (neg << 1) | sign(x)
        }
    }
    CheckOut<Real, QUANT> co[V];
    if (!p.beta_per_edge) {
        Real beta = Real(1);
        if (has_beta) beta = __ldg(static_cast<const Real*>(p.beta_t) + (p.bidx ? __ldg(p.bidx + slot0) : 0));
#pragma unroll
        for (int v = 0; v < V; ++v) co[v].prepare(st[v].m1, st[v].m2, st[v].par, beta, has_beta, qz, p.bc);
    }
#pragma unroll
    for (int v = 0; v < V; ++v) neg[v] <<= (32 - dc);  // bit 31 = sign of edge 0
#pragma unroll 6
    for (int k = 0; k < dc; ++k) {
        Real beta = Real(1);
        if (p.beta_per_edge) beta = __ldg(static_cast<const Real*>(p.beta_t) + __ldg(p.bidx + slot0 + k));
        Pack<OutT, V> out;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint32_t sb = neg[v] & 0x80000000u;
            neg[v] <<= 1;
            if (!p.beta_per_edge) {
                out.v[v] = co[v].emit(k == st[v].k0, sb);
            } else {
                Real raw = (k == st[v].k0) ? st[v].m2 : st[v].m1;
                out.v[v] = cn_emit<Real, QUANT, NTH>(raw, beta, st[v].par ^ sb, qz, p.bc);
            }
        }
        st_stream<Pack<OutT, V>>(dst + (slot0 + k) * p.Bp + f0, out);
    }
}

// Checks of degree > 32: same streaming pass, then the inputs are read again (they were just fetched)
// for their signs and for the "is the minimum" test.
template <typename Real, bool QUANT, int NTH>
__device__ void cn_check_reread(const CnLaunch& p, int64_t slot0, int dc, int64_t f0,
                                const Quantizer<NTH>& qz) {
    constexpr int V = FramesPerLane<Real>::value;
    using OutT = typename CnOut<Real, QUANT>::type;
    const Real* __restrict__ src = static_cast<const Real*>(p.src);
    OutT* __restrict__ dst = static_cast<OutT*>(p.dst);
    const bool has_beta = p.beta_t != nullptr;
    MinState<Real, false> st[V];
#pragma unroll
    for (int v = 0; v < V; ++v) st[v].init();
#pragma unroll 4
    for (int k = 0; k < dc; ++k) {
        int64_t row = p.row_map ? (int64_t)__ldg(p.row_map + slot0 + k) : slot0 + k;
        Pack<Real, V> x = ld_stream<Pack<Real, V>>(src + row * p.Bp + f0);
#pragma unroll
        for (int v = 0; v < V; ++v) st[v].push(x.v[v], k);
    }
    CheckOut<Real, QUANT> co[V];
    if (!p.beta_per_edge) {
        Real beta = Real(1);
        if (has_beta) beta = __ldg(static_cast<const Real*>(p.beta_t) + (p.bidx ? __ldg(p.bidx + slot0) : 0));
#pragma unroll
        for (int v = 0; v < V; ++v) co[v].prepare(st[v].m1, st[v].m2, st[v].par, beta, has_beta, qz, p.bc);
    }
#pragma unroll 4
    for (int k = 0; k < dc; ++k) {
        int64_t row = p.row_map ? (int64_t)__ldg(p.row_map + slot0 + k) : slot0 + k;
        Pack<Real, V> x = *reinterpret_cast<const Pack<Real, V>*>(src + row * p.Bp + f0);
        Real beta = Real(1);
        if (p.beta_per_edge) beta = __ldg(static_cast<const Real*>(p.beta_t) + __ldg(p.bidx + slot0 + k));
        Pack<OutT, V> out;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const bool is_min = Arith<Real>::abs(x.v[v]) == st[v].m1;
            const uint32_t sb = Arith<Real>::hi(x.v[v]);
            if (!p.beta_per_edge) {
                out.v[v] = co[v].emit(is_min, sb);
            } else {
                Real raw = is_min ? st[v].m2 : st[v].m1;
                out.v[v] = cn_emit<Real, QUANT, NTH>(raw, beta, st[v].par ^ sb, qz, p.bc);
            }
        }
        st_stream<Pack<OutT, V>>(dst + (slot0 + k) * p.Bp + f0, out);
    }
}

// resident CTAs per SM the check-node kernel is compiled for: the RCQ variant is issue-bound and gains from
// 4 (64 registers); float32 / float64 stream at the HBM roofline with 3 (80 registers)
#ifndef LDPC_CN_F64_MINCTAS
#define LDPC_CN_F64_MINCTAS 3
#endif
#ifndef LDPC_CN_Q_MINCTAS
#define LDPC_CN_Q_MINCTAS 4
#endif
#define LDPC_CN_BOUNDS __launch_bounds__(kThreads, QUANT ? LDPC_CN_Q_MINCTAS : (sizeof(Real) == 4 ? 3 : LDPC_CN_F64_MINCTAS))
// Messages of stopped frames are never read again (their decisions and posteriors were delivered at the
// iteration they stopped at), so the stores carry no mask code; warps whose frames have all stopped exit.
template <typename Real, bool QUANT, int NTH>
__global__ void LDPC_CN_BOUNDS cn_kernel(const CnLaunch p, const int nfb, const int item0) {
    constexpr int V = FramesPerLane<Real>::value;
    __shared__ float s_thr[kMaxQuantLevels];
    Quantizer<NTH> qz;
    if (QUANT) {
        for (int i = threadIdx.x; i < p.nth; i += blockDim.x) s_thr[i] = p.thr[i];
        __syncthreads();
        qz.load(s_thr, p.nth, p.mono != 0);
    }
    const int fb = blockIdx.x % nfb;
    const int item_id = item0 + blockIdx.x / nfb;
    const int64_t f0 = ((int64_t)fb * blockDim.x + threadIdx.x) * V;
    if (f0 >= p.Bp) return;  // whole warps: Bp is a multiple of 32*V
    const uint32_t done_mask = load_done_mask<V>(p.done, f0);
    if (__all_sync(0xffffffffu, done_mask == ((1u << V) - 1u))) return;
    const WorkItem it = p.items[item_id];
    int64_t slot = it.first_slot;
#define LDPC_CN_CASE(D)                                                                   \
    case D:                                                                               \
        for (int c = 0; c < it.count; ++c, slot += D)                                     \
            cn_check_small<Real, QUANT, NTH, D>(p, slot, f0, qz);                  \
        break;
    switch (it.deg) {
        LDPC_CN_CASE(1)
        LDPC_CN_CASE(2)
        LDPC_CN_CASE(3)
        LDPC_CN_CASE(4)
        LDPC_CN_CASE(5)
        LDPC_CN_CASE(6)
        LDPC_CN_CASE(7)
        LDPC_CN_CASE(8)
        default:
            if (it.deg <= 32) {
                for (int c = 0; c < it.count; ++c, slot += it.deg)
                    cn_check_mask32<Real, QUANT, NTH>(p, slot, it.deg, f0, qz);
            } else {
                for (int c = 0; c < it.count; ++c, slot += it.deg)
                    cn_check_reread<Real, QUANT, NTH>(p, slot, it.deg, f0, qz);
            }
    }
#undef LDPC_CN_CASE
}

// ---------------------------------------------------------------------------------------------
// Wide checks (degree 9..64): bulk-async row ring.
//
// A register-staged streaming loop keeps only a handful of 16-byte loads per thread in flight and none at
// all while a check's outputs are written, so wide checks ran latency-bound (long_scoreboard, ~0.6 of the
// HBM roofline).  Here the rows of a work item -- `count * deg` CONSECUTIVE message rows, one contiguous
// segment of `blockDim.x * 16` bytes each for this CTA's frames -- are streamed into a shared-memory ring
// by the copy engine (cp.async.bulk + mbarrier transaction counts, SASS UBLKCP), kWideSlabs - 1 slabs of
// kWideRows rows ahead of the arithmetic, across check boundaries and across the output phase.  A thread
// reads its own 16-byte column of each row (conflict-free LDS.128), so registers hold only the running
// min1 / min2 / first-argmin / parity and one sign bit per edge.
// ---------------------------------------------------------------------------------------------
constexpr int kWideThreads = 128;
#ifndef LDPC_WIDE_ROWS
#define LDPC_WIDE_ROWS 4
#endif
#ifndef LDPC_WIDE_SLABS
#define LDPC_WIDE_SLABS 4
#endif
#ifndef LDPC_WIDE_Q_MINCTAS
#define LDPC_WIDE_Q_MINCTAS 6
#endif
#ifndef LDPC_WIDE_MINCTAS
#define LDPC_WIDE_MINCTAS 6
#endif
constexpr int kWideRows = LDPC_WIDE_ROWS;     // rows per slab
constexpr int kWideSlabs = LDPC_WIDE_SLABS;   // slabs in the ring
constexpr int kWideMinCtas = LDPC_WIDE_MINCTAS;
constexpr int kWideRowBytes = kWideThreads * 16;
constexpr size_t kWideSmem = (size_t)kWideSlabs * kWideRows * kWideRowBytes;

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_addr(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_addr(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t a = smem_addr(bar);
    uint32_t ok;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok)
            : "r"(a), "r"(parity)
            : "memory");
    } while (!ok);
}
__device__ __forceinline__ uint64_t l2_evict_first_policy() {
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
// global -> shared bulk copy (16-byte aligned, size a multiple of 16); completes `bytes` on `bar`
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t policy) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
            smem_addr(dst)),
        "l"(src), "r"(bytes), "r"(smem_addr(bar)), "l"(policy)
        : "memory");
}

template <typename MaskT> struct SignMask;
template <> struct SignMask<uint32_t> {
    static __device__ __forceinline__ uint32_t push(uint32_t m, uint32_t signword) { return __funnelshift_l(signword, m, 1); }
    static __device__ __forceinline__ uint32_t align(uint32_t m, int dc) { return m << (32 - dc); }
    static __device__ __forceinline__ uint32_t top(uint32_t m) { return m & 0x80000000u; }
    static __device__ __forceinline__ uint32_t word(uint32_t m, int) { return m; }   // 32 edges per word, edge 0 on top
};
template <> struct SignMask<uint64_t> {
    static __device__ __forceinline__ uint64_t push(uint64_t m, uint32_t signword) { return (m << 1) | (uint64_t)(signword >> 31); }
    static __device__ __forceinline__ uint64_t align(uint64_t m, int dc) { return m << (64 - dc); }
    static __device__ __forceinline__ uint32_t top(uint64_t m) { return (uint32_t)(m >> 32) & 0x80000000u; }
    static __device__ __forceinline__ uint32_t word(uint64_t m, int i) { return i == 0 ? (uint32_t)(m >> 32) : (uint32_t)m; }
};

// The ring as seen by one thread.  A check occupies ceil(deg / kWideRows) consecutive slabs (its last slab
// may hold fewer rows), so slab boundaries never fall inside the unrolled row loop.  All threads of the CTA
// walk the slabs in lockstep: acquire(), read rows, release().
template <typename Real>
struct RowRing {
    unsigned char* smem;
    uint64_t* bars;
    const Real* src;
    const int32_t* row_map;
    int64_t Bp, cta_f0;
    int64_t first_row;
    int deg, slabs_per_check, slabs_total;
    uint32_t row_bytes;
    uint64_t policy;
    int g;  // slabs consumed so far

    __device__ __forceinline__ int rows_in(int s) const { return min(kWideRows, deg - s * kWideRows); }
    // warp 0: arm the barrier, then one lane per row issues its copy
    __device__ __forceinline__ void issue(int slab) {
        if (threadIdx.x < 32) {
            const int buf = slab % kWideSlabs;
            const int c = slab / slabs_per_check, s = slab - c * slabs_per_check;
            const int nrows = rows_in(s);
            if (threadIdx.x == 0) mbar_arrive_expect_tx(bars + buf, (uint32_t)nrows * row_bytes);
            __syncwarp();
            if ((int)threadIdx.x < nrows) {
                const int64_t slot = first_row + (int64_t)c * deg + s * kWideRows + threadIdx.x;
                const int64_t row = row_map ? (int64_t)__ldg(row_map + slot) : slot;
                bulk_g2s(smem + (size_t)(buf * kWideRows + threadIdx.x) * kWideRowBytes, src + row * Bp + cta_f0, row_bytes,
                         bars + buf, policy);
            }
        }
    }
    __device__ __forceinline__ void start() {
        g = 0;
        for (int s = 0; s < kWideSlabs && s < slabs_total; ++s) issue(s);
    }
    // wait until the current slab has landed; returns this thread's column of its first row
    __device__ __forceinline__ const unsigned char* acquire(bool active) {
        const int buf = g % kWideSlabs;
        if (active) mbar_wait(bars + buf, (uint32_t)(g / kWideSlabs) & 1u);
        return smem + (size_t)buf * kWideRows * kWideRowBytes + threadIdx.x * 16;
    }
    // every warp is through with the current slab: refill its buffer with the slab kWideSlabs ahead
    __device__ __forceinline__ void release() {
        __syncthreads();
        if (g + kWideSlabs < slabs_total) issue(g + kWideSlabs);
        ++g;
    }
};

// offset min-sum rule (defined with cn_offset_kernel below)
template <typename Real>
__device__ __forceinline__ void offset_weights(const CnLaunch& p, int64_t slot, Real beta_check, Real& beta, Real& alpha);

template <typename Real, bool QUANT, int NTH, typename MaskT, bool OFFSET>
__device__ __forceinline__ void cn_wide_check(const CnLaunch& p, RowRing<Real>& ring, int64_t slot0, int dc, int64_t f0,
                                              bool active, const Quantizer<NTH>& qz) {
    constexpr int V = FramesPerLane<Real>::value;
    using OutT = typename CnOut<Real, QUANT>::type;
    const bool has_beta = p.beta_t != nullptr;
    MinState<Real, true> st[V];
    MaskT neg[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        st[v].init();
        neg[v] = 0;
    }
    for (int s = 0, k0 = 0; s < ring.slabs_per_check; ++s, k0 += kWideRows) {
        const unsigned char* col = ring.acquire(active);
        if (active) {
            if (k0 + kWideRows <= dc) {
#pragma unroll
                for (int q = 0; q < kWideRows; ++q) {
                    const Pack<Real, V> x = *reinterpret_cast<const Pack<Real, V>*>(col + q * kWideRowBytes);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        st[v].push(x.v[v], k0 + q);
                        neg[v] = SignMask<MaskT>::push(neg[v], Arith<Real>::hi(x.v[v]));
                    }
                }
            } else {
                for (int q = 0; q < dc - k0; ++q) {
                    const Pack<Real, V> x = *reinterpret_cast<const Pack<Real, V>*>(col + q * kWideRowBytes);
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        st[v].push(x.v[v], k0 + q);
                        neg[v] = SignMask<MaskT>::push(neg[v], Arith<Real>::hi(x.v[v]));
                    }
                }
            }
        }
        ring.release();
    }
    if (!active) return;
    if constexpr (OFFSET) {
        // c2v = prod(other signs) * (relu(raw - beta) - alpha[variable]); a zero among the OTHER inputs gives 0
#pragma unroll
        for (int v = 0; v < V; ++v) neg[v] = SignMask<MaskT>::align(neg[v], dc);
        Real beta_check = Real(0);
        if (p.beta_t && !p.beta_per_edge) beta_check = __ldg(static_cast<const Real*>(p.beta_t) + (p.bidx ? __ldg(p.bidx + slot0) : 0));
        Real* __restrict__ orow = static_cast<Real*>(p.dst) + slot0 * p.Bp + f0;
        // (unrolled: the weight lookups of an edge are two dependent loads; four edges' worth go out together)
#pragma unroll 4
        for (int k = 0; k < dc; ++k, orow += p.Bp) {
            Real beta, alpha;
            offset_weights<Real>(p, slot0 + k, beta_check, beta, alpha);
            Pack<Real, V> out;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const uint32_t sb = SignMask<MaskT>::top(neg[v]);
                neg[v] <<= 1;
                const bool is_min = (k == st[v].k0);
                const bool zero_others = st[v].m2 == Real(0) || (st[v].m1 == Real(0) && !is_min);
                out.v[v] = offset_value<Real>(is_min ? st[v].m2 : st[v].m1, beta, p.beta_t != nullptr, alpha,
                                              p.alpha_t != nullptr, st[v].par ^ sb, zero_others);
            }
            st_stream<Pack<Real, V>>(orow, out);
        }
        return;
    }
    CheckOut<Real, QUANT> co[V];
    if (!p.beta_per_edge) {
        Real beta = Real(1);
        if (has_beta) beta = __ldg(static_cast<const Real*>(p.beta_t) + (p.bidx ? __ldg(p.bidx + slot0) : 0));
#pragma unroll
        for (int v = 0; v < V; ++v) co[v].prepare(st[v].m1, st[v].m2, st[v].par, beta, has_beta, qz, p.bc);
    }
#pragma unroll
    for (int v = 0; v < V; ++v) neg[v] = SignMask<MaskT>::align(neg[v], dc);  // top bit = sign of edge 0
    OutT* __restrict__ out_row = static_cast<OutT*>(p.dst) + slot0 * p.Bp + f0;
    if constexpr (QUANT && V == 4) {
        if (!p.beta_per_edge) {
            // byte-parallel output phase: edges in groups of eight, whose input signs sit in one byte per frame
            CheckOut4 c4;
            c4.pack(co);
            const uint32_t K0 = pack_low_bytes((uint32_t)st[0].k0, (uint32_t)st[1].k0, (uint32_t)st[2].k0, (uint32_t)st[3].k0);
            for (int j = 0; j * 8 < dc; ++j) {
                uint32_t w[4];
#pragma unroll
                for (int v = 0; v < 4; ++v) w[v] = SignMask<MaskT>::word(neg[v], j >> 2);
                // byte v = input signs of edges 8j..8j+7 of frame v (edge 8j in bit 7), times the total parity
                const uint32_t S = pack_bytes_at(w[0], w[1], w[2], w[3], 3u - (uint32_t)(j & 3)) ^ c4.PAR;
                const int kend = min(8, dc - j * 8);
#pragma unroll
                for (int i = 0; i < 8; ++i) {
                    if (i < kend) {
                        const uint32_t kk = (uint32_t)(j * 8 + i) * 0x01010101u;
                        const uint32_t min4 = spread_byte_signs(0x80808080u - (K0 ^ kk));   // k, k0 < 128
                        const uint32_t neg4 = spread_byte_signs(S << i);
                        Pack<OutT, V> out;
                        *reinterpret_cast<uint32_t*>(&out) = c4.emit(min4, neg4);
                        st_stream<Pack<OutT, V>>(out_row, out);
                        out_row += p.Bp;
                    }
                }
            }
            return;
        }
    }
#pragma unroll 4
    for (int k = 0; k < dc; ++k, out_row += p.Bp) {
        Real beta = Real(1);
        if (p.beta_per_edge) beta = __ldg(static_cast<const Real*>(p.beta_t) + __ldg(p.bidx + slot0 + k));
        Pack<OutT, V> out;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const uint32_t sb = SignMask<MaskT>::top(neg[v]);
            neg[v] <<= 1;
            if (!p.beta_per_edge) {
                out.v[v] = co[v].emit(k == st[v].k0, sb);
            } else {
                Real raw = (k == st[v].k0) ? st[v].m2 : st[v].m1;
                out.v[v] = cn_emit<Real, QUANT, NTH>(raw, beta, st[v].par ^ sb, qz, p.bc);
            }
        }
        st_stream<Pack<OutT, V>>(out_row, out);
    }
}

template <typename Real, bool QUANT, int NTH, bool OFFSET = false>
__global__ void __launch_bounds__(kWideThreads, QUANT ? LDPC_WIDE_Q_MINCTAS : kWideMinCtas) cn_wide_kernel(const CnLaunch p, const int nfb, const int item0) {
    constexpr int V = FramesPerLane<Real>::value;
    extern __shared__ __align__(128) unsigned char wide_smem[];
    __shared__ __align__(8) uint64_t bars[kWideSlabs];
    __shared__ float s_thr[kMaxQuantLevels];
    Quantizer<NTH> qz;
    if (QUANT) {
        for (int i = threadIdx.x; i < p.nth; i += blockDim.x) s_thr[i] = p.thr[i];
    }
    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < kWideSlabs; ++s) mbar_init(bars + s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    const int fb = blockIdx.x % nfb;
    const int item_id = item0 + blockIdx.x / nfb;
    const int64_t cta_f0 = (int64_t)fb * kWideThreads * V;
    const int64_t f0 = cta_f0 + (int64_t)threadIdx.x * V;
    uint32_t done_mask = (1u << V) - 1u;
    if (f0 < p.Bp) done_mask = load_done_mask<V>(p.done, f0);
    const bool active = !__all_sync(0xffffffffu, done_mask == ((1u << V) - 1u));   // warp-uniform
    if (!__syncthreads_or(active ? 1 : 0)) return;   // also publishes the barriers and s_thr
    if (QUANT) qz.load(s_thr, p.nth, p.mono != 0);
    const WorkItem it = p.items[item_id];
    RowRing<Real> ring;
    ring.smem = wide_smem;
    ring.bars = bars;
    ring.src = static_cast<const Real*>(p.src);
    ring.row_map = p.row_map;
    ring.Bp = p.Bp;
    ring.cta_f0 = cta_f0;
    ring.first_row = it.first_slot;
    ring.deg = it.deg;
    ring.slabs_per_check = (it.deg + kWideRows - 1) / kWideRows;
    ring.slabs_total = ring.slabs_per_check * it.count;
    const int64_t cta_frames = min((int64_t)kWideThreads * V, p.Bp - cta_f0);
    ring.row_bytes = (uint32_t)(cta_frames * (int64_t)sizeof(Real));
    ring.policy = l2_evict_first_policy();
    ring.start();
    int64_t slot = it.first_slot;
    if (it.deg <= 32) {
        for (int c = 0; c < it.count; ++c, slot += it.deg)
            cn_wide_check<Real, QUANT, NTH, uint32_t, OFFSET>(p, ring, slot, it.deg, f0, active, qz);
    } else {
        for (int c = 0; c < it.count; ++c, slot += it.deg)
            cn_wide_check<Real, QUANT, NTH, uint64_t, OFFSET>(p, ring, slot, it.deg, f0, active, qz);
    }
}

// ---------------------------------------------------------------------------------------------
// Offset min-sum check node (neural_minsum_decoder.py:236-253, neural_2d_decoder.py:383-401):
//   c2v = prod(other signs) * (relu(raw - beta) - alpha),   alpha indexed by the edge's VARIABLE degree
// Here the reference's three-valued sign() is visible (relu(0 - beta) - alpha need not be 0), so a zero
// among the OTHER inputs forces the output to 0: that is the case iff m2 == 0 (two zeros), or m1 == 0
// and this edge is not the zero one.  A degree-1 check has an empty product (= 1).
// ---------------------------------------------------------------------------------------------
template <typename Real>
__device__ __forceinline__ void offset_weights(const CnLaunch& p, int64_t slot, Real beta_check, Real& beta, Real& alpha) {
    beta = beta_check;
    if (p.beta_t && p.beta_per_edge) beta = __ldg(static_cast<const Real*>(p.beta_t) + __ldg(p.bidx + slot));
    alpha = Real(0);
    if (p.alpha_t) alpha = __ldg(static_cast<const Real*>(p.alpha_t) + (p.aidx_slot ? __ldg(p.aidx_slot + slot) : 0));
}

template <typename Real, int DC>
__device__ __forceinline__ void cn_offset_small(const CnLaunch& p, int64_t slot0, int64_t f0) {
    constexpr int V = FramesPerLane<Real>::value;
    const Real* __restrict__ src = static_cast<const Real*>(p.src);
    Real* __restrict__ dst = static_cast<Real*>(p.dst);
    Pack<Real, V> x[DC];
    Real beta[DC], alpha[DC];
    Real beta_check = Real(0);
    if (p.beta_t && !p.beta_per_edge) beta_check = __ldg(static_cast<const Real*>(p.beta_t) + (p.bidx ? __ldg(p.bidx + slot0) : 0));
    const uint32_t stride = (uint32_t)p.Bp * (uint32_t)sizeof(Real);
    src += f0;
    dst += f0;
#pragma unroll
    for (int k = 0; k < DC; ++k) {
        const uint32_t row = p.row_map ? (uint32_t)__ldg(p.row_map + slot0 + k) : (uint32_t)(slot0 + k);
        x[k] = ld_stream<Pack<Real, V>>(row_at(src, row, stride));
    }
#pragma unroll
    for (int k = 0; k < DC; ++k) offset_weights<Real>(p, slot0 + k, beta_check, beta[k], alpha[k]);
    Pack<Real, V> out[DC];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        MinState<Real, false> st;
        st.init();
#pragma unroll
        for (int k = 0; k < DC; ++k) st.push(x[k].v[v], k);
        if (DC == 1) st.m2 = st.m1;
#pragma unroll
        for (int k = 0; k < DC; ++k) {
            const bool is_min = Arith<Real>::abs(x[k].v[v]) == st.m1;
            const bool zero_others = (DC > 1) && (st.m2 == Real(0) || (st.m1 == Real(0) && !is_min));
            out[k].v[v] = offset_value<Real>(is_min ? st.m2 : st.m1, beta[k], p.beta_t != nullptr, alpha[k],
                                             p.alpha_t != nullptr, st.par ^ Arith<Real>::hi(x[k].v[v]), zero_others);
        }
    }
#pragma unroll
    for (int k = 0; k < DC; ++k) st_stream<Pack<Real, V>>(row_at(dst, (uint32_t)(slot0 + k), stride), out[k]);
}

template <typename Real>
__device__ void cn_offset_wide(const CnLaunch& p, int64_t slot0, int dc, int64_t f0) {
    constexpr int V = FramesPerLane<Real>::value;
    const Real* __restrict__ src = static_cast<const Real*>(p.src);
    Real* __restrict__ dst = static_cast<Real*>(p.dst);
    MinState<Real, false> st[V];
#pragma unroll
    for (int v = 0; v < V; ++v) st[v].init();
    for (int k = 0; k < dc; ++k) {
        int64_t row = p.row_map ? (int64_t)__ldg(p.row_map + slot0 + k) : slot0 + k;
        Pack<Real, V> x = ld_stream<Pack<Real, V>>(src + row * p.Bp + f0);
#pragma unroll
        for (int v = 0; v < V; ++v) st[v].push(x.v[v], k);
    }
    Real beta_check = Real(0);
    if (p.beta_t && !p.beta_per_edge) beta_check = __ldg(static_cast<const Real*>(p.beta_t) + (p.bidx ? __ldg(p.bidx + slot0) : 0));
    for (int k = 0; k < dc; ++k) {
        int64_t row = p.row_map ? (int64_t)__ldg(p.row_map + slot0 + k) : slot0 + k;
        Pack<Real, V> x = *reinterpret_cast<const Pack<Real, V>*>(src + row * p.Bp + f0);
        Real beta, alpha;
        offset_weights<Real>(p, slot0 + k, beta_check, beta, alpha);
        Pack<Real, V> out;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const bool is_min = Arith<Real>::abs(x.v[v]) == st[v].m1;
            const bool zero_others = st[v].m2 == Real(0) || (st[v].m1 == Real(0) && !is_min);
            out.v[v] = offset_value<Real>(is_min ? st[v].m2 : st[v].m1, beta, p.beta_t != nullptr, alpha,
                                          p.alpha_t != nullptr, st[v].par ^ Arith<Real>::hi(x.v[v]), zero_others);
        }
        st_stream<Pack<Real, V>>(dst + (slot0 + k) * p.Bp + f0, out);
    }
}

template <typename Real>
__global__ void __launch_bounds__(kThreads, sizeof(Real) == 4 ? 3 : 2) cn_offset_kernel(const CnLaunch p, const int nfb, const int item0) {
    constexpr int V = FramesPerLane<Real>::value;
    const int fb = blockIdx.x % nfb;
    const int item_id = item0 + blockIdx.x / nfb;
    const int64_t f0 = ((int64_t)fb * blockDim.x + threadIdx.x) * V;
    if (f0 >= p.Bp) return;
    const uint32_t done_mask = load_done_mask<V>(p.done, f0);
    if (__all_sync(0xffffffffu, done_mask == ((1u << V) - 1u))) return;
    const WorkItem it = p.items[item_id];
    int64_t slot = it.first_slot;
#define LDPC_CNO_CASE(D)                                                \
    case D:                                                             \
        for (int c = 0; c < it.count; ++c, slot += D)                   \
            cn_offset_small<Real, D>(p, slot, f0);               \
        break;
    switch (it.deg) {
        LDPC_CNO_CASE(1)
        LDPC_CNO_CASE(2)
        LDPC_CNO_CASE(3)
        LDPC_CNO_CASE(4)
        LDPC_CNO_CASE(5)
        LDPC_CNO_CASE(6)
        LDPC_CNO_CASE(7)
        LDPC_CNO_CASE(8)
        default:
            for (int c = 0; c < it.count; ++c, slot += it.deg) cn_offset_wide<Real>(p, slot, it.deg, f0);
    }
#undef LDPC_CNO_CASE
}

// ---------------------------------------------------------------------------------------------
// Layered RCQ schedule as the reference executes it (rcq_decoder.py:281-350):
// posteriors start at the LLRs, checks are visited in INDEX order, each visit reads the current
// posteriors of its variables, quantises sp*raw and ADDS the reconstruction to those posteriors in place
// (the "subtract the previous C2V" step subtracts 0 on any graph with more than one non-empty check).
// One thread = one frame walks all checks of one iteration; frames are independent.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kThreads) layered_iter_kernel(float* __restrict__ P, const int64_t* __restrict__ chk_ptr,
                                                                 const int32_t* __restrict__ chk_var, int32_t m,
                                                                 const float* __restrict__ thr, int nth, int bc, int mono,
                                                                 const uint8_t* __restrict__ done, int64_t Bp) {
    __shared__ float s_thr[kMaxQuantLevels];
    for (int i = threadIdx.x; i < nth; i += blockDim.x) s_thr[i] = thr[i];
    __syncthreads();
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= Bp || done[f]) return;
    Quantizer<0> qz;
    qz.load(s_thr, nth, mono != 0);
    for (int32_t i = 0; i < m; ++i) {
        const int64_t e0 = __ldg(chk_ptr + i), e1 = __ldg(chk_ptr + i + 1);
        const int dc = (int)(e1 - e0);
        if (dc == 0) continue;
        MinState<float, false> st;
        st.init();
        for (int k = 0; k < dc; ++k) st.push(P[(int64_t)__ldg(chk_var + e0 + k) * Bp + f], k);
        if (dc == 1) st.m2 = st.m1;
        const uint32_t ia = qz.index(st.m1), ib = qz.index(st.m2);
        const float va = s_thr[ia], vb = s_thr[ib];
        for (int k = 0; k < dc; ++k) {
            float* ptr = P + (int64_t)__ldg(chk_var + e0 + k) * Bp + f;
            const float x = *ptr;
            const bool is_min = fabsf(x) == st.m1;
            const float raw = is_min ? st.m2 : st.m1;
            const float mag = is_min ? vb : va;
            // code sign bit = (sp * raw < 0): a negative product of the other signs AND a non-zero magnitude
            const bool neg = (((st.par ^ __float_as_uint(x)) >> 31) != 0u) && (raw != 0.f);
            *ptr = __fadd_rn(x, neg ? -mag : mag);
        }
    }
}

// The same sequential walk, software-pipelined.  A thread still owns its frame(s) and visits the checks in index
// order, but the walk no longer waits for DRAM once per check: every input of check s comes out of the thread's
// own column of a shared-memory ring of kLayerDepth check slots.  The host classifies every edge (s, k) of the
// walk (ldpc_graph_create, `LayerRec`):
//   * no check in (s - kLayerDepth, s) touches its variable  -> the posterior is copied global -> ring by a
//     per-thread cp.async issued kLayerDepth checks ahead (nothing writes the value in between);
//   * otherwise the latest such check j FORWARDS the value it writes to the posterior row into ring slot s as well
//     (descriptor of its own edge: distance s - j and position k).
// So a dual-diagonal parity chain (every check shares a variable with its predecessor) runs off registers and
// shared memory, and the frame's other posteriors arrive kLayerDepth checks early.  The arithmetic per check is
// the plain kernel's, on the same values in the same order.  Needs check degrees <= kLayerMaxDeg.
#ifndef LDPC_LAYER_DEPTH
#define LDPC_LAYER_DEPTH 4
#endif
#ifndef LDPC_LAYER_CTA_SYNC
#define LDPC_LAYER_CTA_SYNC 0
#endif
#ifndef LDPC_LAYER_MIN_CTAS
#define LDPC_LAYER_MIN_CTAS 1
#endif
constexpr int kLayerDepth = LDPC_LAYER_DEPTH;
constexpr int kLayerThreads = 128;
static_assert(kLayerDepth >= 2 && kLayerDepth <= 16, "descriptor holds a 4-bit distance");
static_assert(kLayerMaxDeg == 8, "descriptor holds a 3-bit position");

// One check of the walk as the kernel reads it: three broadcast 16-byte loads from the warp's record ring in
// shared memory (fixed, short latency -- a record read from global memory sits on every step's critical path
// and an L1 sector miss costs more than the whole step).
struct LayerRecRegs {
    uint4 a, b, c;   // var[0..3] | var[4..7] | desc[0..3], desc[4..7], dc + (copy-ahead mask << 8), pad
    __device__ __forceinline__ void load(const LayerRec* rec) {
        const uint4* r = reinterpret_cast<const uint4*>(rec);
        a = r[0];
        b = r[1];
        c = r[2];
    }
    __device__ __forceinline__ uint32_t var(int k) const {
        return k == 0 ? a.x : k == 1 ? a.y : k == 2 ? a.z : k == 3 ? a.w : k == 4 ? b.x : k == 5 ? b.y : k == 6 ? b.z : b.w;
    }
    __device__ __forceinline__ uint32_t fwd(int k) const { return ((k < 4 ? c.x : c.y) >> (8 * (k & 3))) & 0x7fu; }
    __device__ __forceinline__ int dc() const { return (int)(c.z & 0xffu); }
    __device__ __forceinline__ uint32_t ahead_mask() const { return c.z >> 8; }
};

// reconstruction value of a magnitude: thr[last j >= 1 with mag >= thr[j]], thr[0] if none (rcq_decoder.py:76-84,
// 100-121).  NTH > 0: the table sits in registers and "the last one that passes" is a chain of selects.
template <int NTH>
struct LayerQuant {
    float t[NTH > 0 ? NTH : 1];
    const float* s_thr;
    int nth;
    bool mono;
    __device__ __forceinline__ void load(const float* s_thr_, int nth_, bool mono_) {
        s_thr = s_thr_;
        nth = nth_;
        mono = mono_;
        if constexpr (NTH > 0) {
#pragma unroll
            for (int j = 0; j < NTH; ++j) t[j] = (j < nth_) ? s_thr_[j] : __int_as_float(0x7fc00000);   // NaN never passes
        }
    }
    __device__ __forceinline__ float value(float mag) const {
        if constexpr (NTH > 0) {
            float v = t[0];
#pragma unroll
            for (int j = 1; j < NTH; ++j) v = (mag >= t[j]) ? t[j] : v;
            return v;
        } else {
            return s_thr[quant_index(mag, s_thr, nth, mono)];
        }
    }
};

constexpr uint32_t kLayerPosBytes = kLayerThreads * sizeof(float);          // one ring position, all threads
constexpr uint32_t kLayerSlotBytes = kLayerMaxDeg * kLayerPosBytes;         // one check
constexpr uint32_t kLayerRingBytes = kLayerDepth * kLayerSlotBytes;
static_assert((kLayerDepth & (kLayerDepth - 1)) == 0, "ring offsets wrap with a mask");
static_assert(kLayerSlotBytes == (8u << 9) && kLayerPosBytes == (1u << 9), "descriptor (distance << 3 | position) << 9 is a ring offset");

// One check of degree DC for the V frames of a thread: `col` is the thread's column of the ring (V adjacent
// floats per position), `sbase` the byte offset of the check's slot, bit v of `live` says frame v still runs.
template <int DC, int NTH, int V>
__device__ __forceinline__ void layer_step(const LayerRecRegs& r, float* __restrict__ Pf, uint32_t live, uint32_t stride, char* col,
                                           uint32_t sbase, const LayerQuant<NTH>& qz) {
    Pack<float, V> x[DC];
#pragma unroll
    for (int k = 0; k < DC; ++k) x[k] = *reinterpret_cast<const Pack<float, V>*>(col + sbase + k * kLayerPosBytes);
    MinState<float, false> st[V];
    uint32_t a1[V], a2[V], g1[V], g2[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        st[v].init();
#pragma unroll
        for (int k = 0; k < DC; ++k) st[v].push(x[k].v[v], k);
        if (DC == 1) st[v].m2 = st[v].m1;
        // code sign bit = (sp * raw < 0): a negative product of the other signs AND a non-zero magnitude; the
        // reconstruction is then -value, i.e. the value with its sign bit flipped
        a1[v] = __float_as_uint(qz.value(st[v].m1));
        a2[v] = __float_as_uint(qz.value(st[v].m2));
        g1[v] = st[v].m1 != 0.f ? 0x80000000u : 0u;
        g2[v] = st[v].m2 != 0.f ? 0x80000000u : 0u;
    }
#pragma unroll
    for (int k = 0; k < DC; ++k) {
        Pack<float, V> out;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const bool is_min = fabsf(x[k].v[v]) == st[v].m1;
            const uint32_t rec = (is_min ? a2[v] : a1[v]) ^ ((st[v].par ^ __float_as_uint(x[k].v[v])) & (is_min ? g2[v] : g1[v]));
            // stopped frames keep their posteriors: the value they came in with goes back out (V > 1: one vector store
            // per edge whatever the mix of running and stopped frames in the lane, no per-frame store branches)
            out.v[v] = (V == 1 || ((live >> v) & 1u)) ? __fadd_rn(x[k].v[v], __uint_as_float(rec)) : x[k].v[v];
        }
        float* row = row_at(Pf, r.var(k), stride);
        if (live) *reinterpret_cast<Pack<float, V>*>(row) = out;
        const uint32_t fwd = r.fwd(k);   // next reader within the ring: (distance << 3) | position
        if (fwd) *reinterpret_cast<Pack<float, V>*>(col + ((sbase + (fwd << 9)) & (kLayerRingBytes - 1))) = out;
    }
}

constexpr int kLayerRecAhead = 3 * kLayerDepth;   // a record is staged this many steps before its check is visited
constexpr int kLayerRecRing = 4 * kLayerDepth;    // records per warp ring
static_assert(kLayerRecAhead >= 2 * kLayerDepth + 1, "a staged record must have landed kLayerDepth steps before its check");
static_assert(kLayerRecRing > kLayerRecAhead && (kLayerRecRing & (kLayerRecRing - 1)) == 0, "record ring");

__device__ __forceinline__ void cp_async_16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16;" ::"r"(smem_addr(smem_dst)), "l"(gmem_src) : "memory");
}

template <int BYTES>
__device__ __forceinline__ void cp_async_small(void* smem_dst, const void* gmem_src) {
    static_assert(BYTES == 4 || BYTES == 8 || BYTES == 16, "cp.async.ca size");
    asm volatile("cp.async.ca.shared.global [%0], [%1], %2;" ::"r"(smem_addr(smem_dst)), "l"(gmem_src), "n"(BYTES) : "memory");
}

// V frames per thread (kLayerThreads / V threads per CTA: a CTA always owns kLayerThreads frames, so the ring
// geometry does not depend on V).  V = 2 amortises the warp-uniform work of a step (records, addresses, predicates)
// over two frames and gives every thread two independent dependency chains; it pays once the batch is large
// enough to fill the machine with half the warps.
template <int NTH, int V>
__global__ void __launch_bounds__(kLayerThreads / V, LDPC_LAYER_MIN_CTAS) layered_pipe_kernel(float* __restrict__ P, const LayerRec* __restrict__ recs,
                                                                      int n_checks, const float* __restrict__ thr, int nth,
                                                                      int mono, const uint8_t* __restrict__ done, int64_t Bp) {
    __shared__ float s_thr[kMaxQuantLevels];
    __shared__ __align__(16) float s_ring[kLayerDepth][kLayerMaxDeg][kLayerThreads];
    __shared__ __align__(16) LayerRec s_recs[kLayerThreads / V / 32][kLayerRecRing];
    for (int i = threadIdx.x; i < nth; i += blockDim.x) s_thr[i] = thr[i];
    __syncthreads();
    // A warp walks as long as one of its frames runs; lanes of stopped frames walk along (the warp stages its
    // records cooperatively) but never write a posterior.
    int64_t f = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * V;
    uint32_t live = 0;
#pragma unroll
    for (int v = 0; v < V; ++v)
        if (f + v < Bp && !done[f + v]) live |= 1u << v;
#if LDPC_LAYER_CTA_SYNC
    if (!__syncthreads_or(live != 0)) return;
#else
    if (!__any_sync(0xffffffffu, live != 0)) return;
#endif
    if (f + V > Bp) f = Bp - V;   // (Bp is a multiple of the CTA's frame count; such a lane has live == 0)
    const int lane = threadIdx.x & 31;
    LayerRec* const wrecs = s_recs[threadIdx.x >> 5];
    LayerQuant<NTH> qz;
    qz.load(s_thr, nth, mono != 0);
    float* const Pf = P + f;
    const uint32_t stride = (uint32_t)Bp * (uint32_t)sizeof(float);
    char* const col = reinterpret_cast<char*>(&s_ring[0][0][threadIdx.x * V]);
    // the copies of a check that may be issued ahead of time, into the slot at byte offset sbase
    auto fetch = [&](const LayerRecRegs& r, uint32_t sbase) {
        const uint32_t mask = r.ahead_mask();
#pragma unroll
        for (int k = 0; k < kLayerMaxDeg; ++k)
            if (mask & (1u << k))
                cp_async_small<4 * V>(col + sbase + k * kLayerPosBytes, row_at(Pf, r.var(k), stride));
    };
    // records 0 .. kLayerRecAhead-1 synchronously, 16 bytes per lane and turn
    {
        const int chunks = min(kLayerRecAhead, n_checks) * 3;
        const uint4* src = reinterpret_cast<const uint4*>(recs);
        uint4* dst = reinterpret_cast<uint4*>(wrecs);
        for (int c = lane; c < chunks; c += 32) dst[c] = __ldg(src + c);
        __syncwarp();
    }
    for (int s = 0; s < kLayerDepth; ++s) {
        if (s < n_checks) {
            LayerRecRegs r;
            r.load(wrecs + s);
            fetch(r, (uint32_t)s * kLayerSlotBytes);
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
    for (int s = 0; s < n_checks; ++s) {
        // lanes 0..2 stage the record of check s + kLayerRecAhead; it belongs to this step's copy group, which has
        // landed (wait_group below) 2 * kLayerDepth steps before anybody reads it, and the per-step __syncwarp
        // publishes it to the other lanes
        if (lane < 3 && s + kLayerRecAhead < n_checks)
            cp_async_16(reinterpret_cast<uint4*>(wrecs + ((s + kLayerRecAhead) % kLayerRecRing)) + lane,
                        reinterpret_cast<const uint4*>(recs + s + kLayerRecAhead) + lane);
        LayerRecRegs cur, ahead;
        cur.load(wrecs + (s % kLayerRecRing));
        ahead.load(wrecs + ((s + kLayerDepth) % kLayerRecRing));   // used at the end of this step (if that check exists)
        asm volatile("cp.async.wait_group %0;" ::"n"(kLayerDepth - 1) : "memory");   // the copies of check s have landed
#if LDPC_LAYER_CTA_SYNC
        __syncthreads();   // tuning build: the warps of a CTA visit a check together (their row segments are adjacent)
#else
        __syncwarp();
#endif
        const uint32_t sbase = ((uint32_t)s % kLayerDepth) * kLayerSlotBytes;
        switch (cur.dc()) {   // warp-uniform
            case 1: layer_step<1, NTH, V>(cur, Pf, live, stride, col, sbase, qz); break;
            case 2: layer_step<2, NTH, V>(cur, Pf, live, stride, col, sbase, qz); break;
            case 3: layer_step<3, NTH, V>(cur, Pf, live, stride, col, sbase, qz); break;
            case 4: layer_step<4, NTH, V>(cur, Pf, live, stride, col, sbase, qz); break;
            case 5: layer_step<5, NTH, V>(cur, Pf, live, stride, col, sbase, qz); break;
            case 6: layer_step<6, NTH, V>(cur, Pf, live, stride, col, sbase, qz); break;
            case 7: layer_step<7, NTH, V>(cur, Pf, live, stride, col, sbase, qz); break;
            default: layer_step<8, NTH, V>(cur, Pf, live, stride, col, sbase, qz); break;
        }
        if (s + kLayerDepth < n_checks) fetch(ahead, sbase);   // slot s is free again: check s + kLayerDepth moves in
        asm volatile("cp.async.commit_group;" ::: "memory");
    }
}

// The same schedule, level-parallel.  Checks are grouped into dependency LEVELS on the host (a check's level is one
// above the highest level among the earlier checks it shares a variable with), so checks of one level touch
// disjoint variables and every posterior still receives its updates in check-index order: running a level's
// checks concurrently gives exactly the sequential result.  Quasi-cyclic codes have a handful of levels (one
// per block row), chain-structured codes (dual-diagonal parity) have as many levels as checks and keep the
// sequential kernel.  A lane owns four frames; one CTA handles kLayerChunk checks of the level.
constexpr int kLayerChunk = 4;
__global__ void __launch_bounds__(kThreads) layered_level_kernel(float* __restrict__ P, const int64_t* __restrict__ chk_ptr,
                                                                  const int32_t* __restrict__ chk_var,
                                                                  const int32_t* __restrict__ level_chk, int n_checks,
                                                                  const float* __restrict__ thr, int nth, int mono,
                                                                  const uint8_t* __restrict__ done, int64_t Bp, int nfb) {
    constexpr int V = 4;
    __shared__ float s_thr[kMaxQuantLevels];
    for (int i = threadIdx.x; i < nth; i += blockDim.x) s_thr[i] = thr[i];
    __syncthreads();
    const int fb = blockIdx.x % nfb;
    const int group = blockIdx.x / nfb;
    const int64_t f0 = ((int64_t)fb * blockDim.x + threadIdx.x) * V;
    if (f0 >= Bp) return;
    const uint32_t dmask = load_done_mask<V>(done, f0);
    if (__all_sync(0xffffffffu, dmask == ((1u << V) - 1u))) return;
    Quantizer<0> qz;
    qz.load(s_thr, nth, mono != 0);
    const uint32_t stride = (uint32_t)Bp * (uint32_t)sizeof(float);
    float* __restrict__ P0 = P + f0;
    for (int c = group * kLayerChunk; c < min(n_checks, (group + 1) * kLayerChunk); ++c) {
        const int32_t i = __ldg(level_chk + c);
        const int64_t e0 = __ldg(chk_ptr + i), e1 = __ldg(chk_ptr + i + 1);
        const int dc = (int)(e1 - e0);
        MinState<float, false> st[V];
#pragma unroll
        for (int v = 0; v < V; ++v) st[v].init();
        // eight rows in flight at a time (a row at a time leaves the check latency-bound: dc is 29/30 on the QC shape)
        for (int k0 = 0; k0 < dc; k0 += 8) {
            Pack<float, V> x[8];
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (k0 + u < dc)
                    x[u] = *reinterpret_cast<const Pack<float, V>*>(row_at(P0, (uint32_t)__ldg(chk_var + e0 + k0 + u), stride));
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (k0 + u < dc) {
#pragma unroll
                    for (int v = 0; v < V; ++v) st[v].push(x[u].v[v], k0 + u);
                }
        }
        float va[V], vb[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (dc == 1) st[v].m2 = st[v].m1;
            va[v] = s_thr[qz.index(st[v].m1)];
            vb[v] = s_thr[qz.index(st[v].m2)];
        }
        for (int k0 = 0; k0 < dc; k0 += 4) {
            float* ptr[4];
            Pack<float, V> x[4];
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k0 + u < dc) {
                    ptr[u] = row_at(P0, (uint32_t)__ldg(chk_var + e0 + k0 + u), stride);
                    x[u] = *reinterpret_cast<const Pack<float, V>*>(ptr[u]);
                }
#pragma unroll
            for (int u = 0; u < 4; ++u)
                if (k0 + u < dc) {
                    Pack<float, V> out;
#pragma unroll
                    for (int v = 0; v < V; ++v) {
                        const bool is_min = fabsf(x[u].v[v]) == st[v].m1;
                        const float raw = is_min ? st[v].m2 : st[v].m1;
                        const float mag = is_min ? vb[v] : va[v];
                        // code sign bit = (sp * raw < 0): a negative product of the other signs AND a non-zero magnitude
                        const bool neg = (((st[v].par ^ __float_as_uint(x[u].v[v])) >> 31) != 0u) && (raw != 0.f);
                        out.v[v] = __fadd_rn(x[u].v[v], neg ? -mag : mag);
                    }
                    store_masked<float, V>(ptr[u], out, dmask);   // stopped frames keep their posteriors
                }
        }
    }
}


// The level-parallel kernel with the check's posteriors STAGED in shared memory: a thread copies the dc row segments of
// its four frames with per-thread cp.async (16 bytes each, all dc in flight at once, no register staging) into its own
// column of the stage -- no barrier: a column belongs to one thread -- and both passes (min / parity, then the outputs)
// read shared memory, so every posterior is read from global memory ONCE and written once: the `8 * E` bytes per
// frame-iteration the schedule needs.  (The kernel above reads every row twice, eight and four rows in flight: 0.56
// of that roofline on the (9472,8192)-shaped QC code, dc = 29 / 30; this one 0.65 at 32 768 frames and 0.73 at 131 072.)
// Check degrees up to kLevelStageMaxDeg.  (Packing the hard decisions of the written posteriors here as well -- to save the
// pass over all n rows per iteration -- was tried and lost, 1.41 -> 1.17 M frames/s: four 16-byte partial-sector stores
// per edge and warp, dv times per variable, cost more than one 4n-byte pass.  Tried again with a host-made flag on the ONE
// edge per variable that enters its last level, so each decision word is written once per iteration and only iteration 0
// keeps the separate pass: 1.59 -> 1.46 M frames/s at 131 072 frames, 1.49 -> 1.29 M at 32 768 -- the ballots need whole
// warps and the extra predicates / registers slow the unfused path of the same kernel by 6 % as well.)
constexpr int kLevelStageThreads = 128;
constexpr int kLevelStageMaxDeg = 64;
#ifndef LDPC_LEVEL_CHUNK
#define LDPC_LEVEL_CHUNK 4      // checks of a level per CTA of the staged kernel (8 and 16 measured slower, with or without the prefetch)
#endif
#ifndef LDPC_LEVEL_PREFETCH
#define LDPC_LEVEL_PREFETCH 1   // the rows of the CTA's next check are asked into L2 while the current one is staged (+2-6 %)
#endif
constexpr int kLevelStageChunk = LDPC_LEVEL_CHUNK;
template <int V>
__global__ void __launch_bounds__(kLevelStageThreads) layered_level_stage_kernel(float* __restrict__ P, const int64_t* __restrict__ chk_ptr,
                                                                               const int32_t* __restrict__ chk_var,
                                                                               const int32_t* __restrict__ level_chk, int n_checks,
                                                                               const float* __restrict__ thr, int nth, int mono,
                                                                               const uint8_t* __restrict__ done, int64_t Bp, int nfb) {
    static_assert(V == 2 || V == 4, "frames per lane");
    constexpr int SEG = 4 * V;                                     // bytes of a lane's row segment
    extern __shared__ __align__(16) unsigned char level_stage[];   // [max dc][threads] segments
    __shared__ float s_thr[kMaxQuantLevels];
    for (int i = threadIdx.x; i < nth; i += blockDim.x) s_thr[i] = thr[i];
    __syncthreads();
    const int fb = blockIdx.x % nfb;
    const int group = blockIdx.x / nfb;
    const int64_t f0 = ((int64_t)fb * blockDim.x + threadIdx.x) * V;
    if (f0 >= Bp) return;
    const uint32_t dmask = load_done_mask<V>(done, f0);
    if (dmask == ((1u << V) - 1u)) return;   // (per lane: the stage needs no warp or block agreement)
    Quantizer<0> qz;
    qz.load(s_thr, nth, mono != 0);
    const uint32_t stride = (uint32_t)Bp * (uint32_t)sizeof(float);
    float* __restrict__ P0 = P + f0;
    unsigned char* const col = level_stage + threadIdx.x * SEG;
    constexpr uint32_t kRow = kLevelStageThreads * SEG;
    const int c_end = min(n_checks, (group + 1) * kLevelStageChunk);
    for (int c = group * kLevelStageChunk; c < c_end; ++c) {
        const int32_t i = __ldg(level_chk + c);
        const int64_t e0 = __ldg(chk_ptr + i), e1 = __ldg(chk_ptr + i + 1);
        const int dc = (int)(e1 - e0);
        for (int k = 0; k < dc; ++k) {
            const float* src = row_at(P0, (uint32_t)__ldg(chk_var + e0 + k), stride);
            if constexpr (V == 4)
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(smem_addr(col + k * kRow)), "l"(src) : "memory");
            else
                asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(smem_addr(col + k * kRow)), "l"(src) : "memory");
        }
#if LDPC_LEVEL_PREFETCH
        if (c + 1 < c_end) {   // (checks of a level touch disjoint variables: nothing writes these rows in between)
            const int32_t i2 = __ldg(level_chk + c + 1);
            const int64_t n0 = __ldg(chk_ptr + i2), n1 = __ldg(chk_ptr + i2 + 1);
            for (int64_t e = n0; e < n1; ++e)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(row_at(P0, (uint32_t)__ldg(chk_var + e), stride)));
        }
#endif
        asm volatile("cp.async.wait_all;" ::: "memory");
        MinState<float, false> st[V];
#pragma unroll
        for (int v = 0; v < V; ++v) st[v].init();
#pragma unroll 4
        for (int k = 0; k < dc; ++k) {
            const Pack<float, V> x = *reinterpret_cast<const Pack<float, V>*>(col + k * kRow);
#pragma unroll
            for (int v = 0; v < V; ++v) st[v].push(x.v[v], k);
        }
        float va[V], vb[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            if (dc == 1) st[v].m2 = st[v].m1;
            va[v] = s_thr[qz.index(st[v].m1)];
            vb[v] = s_thr[qz.index(st[v].m2)];
        }
#pragma unroll 4
        for (int k = 0; k < dc; ++k) {
            const Pack<float, V> x = *reinterpret_cast<const Pack<float, V>*>(col + k * kRow);
            Pack<float, V> out;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const bool is_min = fabsf(x.v[v]) == st[v].m1;
                const float raw = is_min ? st[v].m2 : st[v].m1;
                const float mag = is_min ? vb[v] : va[v];
                // code sign bit = (sp * raw < 0): a negative product of the other signs AND a non-zero magnitude
                const bool neg = (((st[v].par ^ __float_as_uint(x.v[v])) >> 31) != 0u) && (raw != 0.f);
                // stopped frames keep their posteriors: the staged value is written back unchanged
                out.v[v] = ((dmask >> v) & 1u) ? x.v[v] : __fadd_rn(x.v[v], neg ? -mag : mag);
            }
            *reinterpret_cast<Pack<float, V>*>(row_at(P0, (uint32_t)__ldg(chk_var + e0 + k), stride)) = out;
        }
    }
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------
namespace {

template <typename Real, bool QUANT, int NTH>
cudaError_t launch_cn_range(const CnLaunch& p, int item0, int item1, bool wide, cudaStream_t stream) {
    if (item1 <= item0) return cudaSuccess;
    constexpr int V = FramesPerLane<Real>::value;
    if (wide) {
        // per device and cheap, so simply repeated on every launch
        cudaError_t e = cudaFuncSetAttribute(cn_wide_kernel<Real, QUANT, NTH>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kWideSmem);
        if (e != cudaSuccess) return e;
        const int64_t nfb = (p.Bp + (int64_t)kWideThreads * V - 1) / ((int64_t)kWideThreads * V);
        const int64_t grid = nfb * (item1 - item0);
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        cn_wide_kernel<Real, QUANT, NTH><<<(unsigned)grid, kWideThreads, kWideSmem, stream>>>(p, (int)nfb, item0);
    } else {
        const int threads = threads_for(p.Bp, V);
        const int64_t nfb = (p.Bp / V + threads - 1) / threads;
        const int64_t grid = nfb * (item1 - item0);
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        cn_kernel<Real, QUANT, NTH><<<(unsigned)grid, threads, 0, stream>>>(p, (int)nfb, item0);
    }
    return cudaGetLastError();
}

template <typename Real, bool QUANT, int NTH>
cudaError_t launch_cn_all(const CnLaunch& p, cudaStream_t stream) {
    // items are sorted by degree: [0, wide0) degree <= 8, [wide0, wide1) degree 9..64 (row ring), rest > 64
    const int wide0 = p.wide_ring ? p.items_wide_begin : p.n_items, wide1 = p.wide_ring ? p.items_wide_end : p.n_items;
    cudaError_t e = launch_cn_range<Real, QUANT, NTH>(p, 0, wide0, false, stream);
    if (e == cudaSuccess) e = launch_cn_range<Real, QUANT, NTH>(p, wide0, wide1, true, stream);
    if (e == cudaSuccess) e = launch_cn_range<Real, QUANT, NTH>(p, wide1, p.n_items, false, stream);
    return e;
}

}  // namespace

cudaError_t launch_cn(int dtype, const CnLaunch& p, cudaStream_t stream) {
    if (p.n_items == 0) return cudaSuccess;
    if (dtype == 0) {
        if (p.nth > 0) {
            // register-resident thresholds need a non-decreasing table (count == last index reached)
            if (p.mono && p.nth <= 4) return launch_cn_all<float, true, 4>(p, stream);
            if (p.mono && p.nth <= 8) return launch_cn_all<float, true, 8>(p, stream);
            return launch_cn_all<float, true, 0>(p, stream);
        }
        return launch_cn_all<float, false, 0>(p, stream);
    }
    return launch_cn_all<double, false, 0>(p, stream);
}

namespace {

template <typename Real>
cudaError_t launch_cn_offset_range(const CnLaunch& p, int item0, int item1, bool wide, cudaStream_t stream) {
    if (item1 <= item0) return cudaSuccess;
    constexpr int V = FramesPerLane<Real>::value;
    if (wide) {
        cudaError_t e = cudaFuncSetAttribute(cn_wide_kernel<Real, false, 0, true>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kWideSmem);
        if (e != cudaSuccess) return e;
        const int64_t nfb = (p.Bp + (int64_t)kWideThreads * V - 1) / ((int64_t)kWideThreads * V);
        const int64_t grid = nfb * (item1 - item0);
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        cn_wide_kernel<Real, false, 0, true><<<(unsigned)grid, kWideThreads, kWideSmem, stream>>>(p, (int)nfb, item0);
    } else {
        const int threads = threads_for(p.Bp, V);
        const int64_t nfb = (p.Bp / V + threads - 1) / threads;
        const int64_t grid = nfb * (item1 - item0);
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        cn_offset_kernel<Real><<<(unsigned)grid, threads, 0, stream>>>(p, (int)nfb, item0);
    }
    return cudaGetLastError();
}

template <typename Real>
cudaError_t launch_cn_offset_all(const CnLaunch& p, cudaStream_t stream) {
    const int wide0 = p.wide_ring ? p.items_wide_begin : p.n_items, wide1 = p.wide_ring ? p.items_wide_end : p.n_items;
    cudaError_t e = launch_cn_offset_range<Real>(p, 0, wide0, false, stream);
    if (e == cudaSuccess) e = launch_cn_offset_range<Real>(p, wide0, wide1, true, stream);
    if (e == cudaSuccess) e = launch_cn_offset_range<Real>(p, wide1, p.n_items, false, stream);
    return e;
}

}  // namespace

cudaError_t launch_cn_offset(int dtype, const CnLaunch& p, cudaStream_t stream) {
    if (p.n_items == 0) return cudaSuccess;
    return dtype == 0 ? launch_cn_offset_all<float>(p, stream) : launch_cn_offset_all<double>(p, stream);
}

cudaError_t launch_layered_level(float* P, const int64_t* chk_ptr, const int32_t* chk_var, const int32_t* level_chk,
                                 int n_checks, const float* thr, int nth, int mono, const uint8_t* done, int64_t Bp,
                                 int max_dc, int staged, cudaStream_t stream) {
    if (n_checks <= 0) return cudaSuccess;
    if (staged && max_dc <= kLevelStageMaxDeg) {
        // staged == 2 (LDPC_LAYERED_STAGE=2, tuning): two frames per lane -- half the stage per warp, twice the warps per SM;
        // measured slower (131 072 frames: 1.40 M frames/s against 1.55 M), so occupancy is not what bounds the kernel
        const int V = staged == 2 ? 2 : 4;
        const size_t smem = (size_t)std::max(max_dc, 1) * kLevelStageThreads * 4 * V;
        auto kern = V == 2 ? layered_level_stage_kernel<2> : layered_level_stage_kernel<4>;
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        const int64_t nfb = (Bp / V + kLevelStageThreads - 1) / kLevelStageThreads;
        const int64_t grid = nfb * ((n_checks + kLevelStageChunk - 1) / kLevelStageChunk);
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        kern<<<(unsigned)grid, kLevelStageThreads, smem, stream>>>(P, chk_ptr, chk_var, level_chk, n_checks, thr, nth, mono, done, Bp, (int)nfb);
        return cudaGetLastError();
    }
    const int threads = threads_for(Bp, 4);
    const int64_t nfb = (Bp / 4 + threads - 1) / threads;
    const int64_t grid = nfb * ((n_checks + kLayerChunk - 1) / kLayerChunk);
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    layered_level_kernel<<<(unsigned)grid, threads, 0, stream>>>(P, chk_ptr, chk_var, level_chk, n_checks, thr, nth, mono, done,
                                                                  Bp, (int)nfb);
    return cudaGetLastError();
}

namespace {
template <int V>
cudaError_t launch_layered_pipe_v(float* P, const LayerRec* recs, int n_checks, const float* thr, int nth, int mono,
                                  const uint8_t* done, int64_t Bp, cudaStream_t stream) {
    const unsigned grid = (unsigned)((Bp + kLayerThreads - 1) / kLayerThreads);
    const int threads = kLayerThreads / V;
    if (nth <= 4) layered_pipe_kernel<4, V><<<grid, threads, 0, stream>>>(P, recs, n_checks, thr, nth, mono, done, Bp);
    else if (nth <= 8) layered_pipe_kernel<8, V><<<grid, threads, 0, stream>>>(P, recs, n_checks, thr, nth, mono, done, Bp);
    else layered_pipe_kernel<0, V><<<grid, threads, 0, stream>>>(P, recs, n_checks, thr, nth, mono, done, Bp);
    return cudaGetLastError();
}
}  // namespace

// frames_per_thread: 1, 2 or 4 (Bp must be a multiple of kLayerThreads, which the workspace padding guarantees)
cudaError_t launch_layered_pipe(float* P, const LayerRec* recs, int n_checks, const float* thr, int nth, int mono,
                                const uint8_t* done, int64_t Bp, int frames_per_thread, cudaStream_t stream) {
    if (n_checks <= 0) return cudaSuccess;
    if (Bp % kLayerThreads != 0) return cudaErrorInvalidValue;
    if (frames_per_thread == 4) return launch_layered_pipe_v<4>(P, recs, n_checks, thr, nth, mono, done, Bp, stream);
    return frames_per_thread == 2 ? launch_layered_pipe_v<2>(P, recs, n_checks, thr, nth, mono, done, Bp, stream)
                                  : launch_layered_pipe_v<1>(P, recs, n_checks, thr, nth, mono, done, Bp, stream);
}

int layered_pipe_depth() { return kLayerDepth; }

cudaError_t launch_layered_iter(float* P, const int64_t* chk_ptr, const int32_t* chk_var, int32_t m, const float* thr,
                                int nth, int bc, int mono, const uint8_t* done, int64_t Bp, cudaStream_t stream) {
    const int threads = (int)(Bp < 128 ? Bp : 128);
    layered_iter_kernel<<<(unsigned)((Bp + threads - 1) / threads), threads, 0, stream>>>(P, chk_ptr, chk_var, m, thr, nth,
                                                                                            bc, mono, done, Bp);
    return cudaGetLastError();
}

}  // namespace ldpc
