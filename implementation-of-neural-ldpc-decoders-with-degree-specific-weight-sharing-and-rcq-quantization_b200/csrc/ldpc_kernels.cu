// Hand-written sm_100a kernels of the flooding min-sum / RCQ hot path.
//
// Data layout: every per-edge message array is [slot][Bp] with the FRAME index
// fastest, slots numbered check-major over degree-sorted checks.  One lane owns V consecutive frames
// (V = 4 for float, 2 for double), so a warp moves 512 contiguous bytes per message row and the
// graph indices / weights are warp-uniform broadcast loads.  The work is gather / min / add on
// streaming data: HBM-bound, no tensor cores (it is not a contraction).
//
//   cn_kernel   check-node half iteration      reads v2c [E][Bp]     writes c2v [E][Bp] (float or code)
//   vn_kernel   variable-node half iteration   reads c2v, llrT       writes v2c, packed hard decisions
//   syn_kernel  parity checks on the bit-packed hard decisions (32 frames per word)
//   commit      per-frame early stop bookkeeping
//   pack/unpack row-major user buffers <-> interleaved layout, Philox AWGN, error counting
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
    if (keep != 0) {
        const Pack<T, V> old = *reinterpret_cast<const Pack<T, V>*>(rowptr);
#pragma unroll
        for (int v = 0; v < V; ++v)
            if ((keep >> v) & 1u) val.v[v] = old.v[v];
    }
    st_stream<Pack<T, V>>(rowptr, val);
}

// ---------------------------------------------------------------------------------------------
// Check node (ldpc_decoder.py:91-120; neural_2d_decoder.py:161-191; rcq_decoder.py:211-246, :526-563)
//
// Per frame: m1 = min |x|, k0 = its first index, m2 = min over the others; for edge k
//   raw = (k == k0) ? m2 : m1,   c2v = fl(beta_k * raw) with the sign of prod_{k' != k} sign(x_k').
// The product of the other signs is applied as an XOR of IEEE sign bits: whenever an input is +-0 the
// magnitudes force every affected output to +-0 (appendix A2), so three-valued sign() never shows.
// The same expression covers (beta*raw)*sp [N-MS] and (beta*sp)*raw [W-RCQ]: sp = +-1 is exact.
// ---------------------------------------------------------------------------------------------
template <typename Real, bool QUANT>
struct CnOut { using type = Real; };
template <typename Real>
struct CnOut<Real, true> { using type = uint8_t; };

// Running min1 / min2 / sign parity of one check for one frame.  TRACK_K0 also keeps the first index of
// the minimum; it is only needed where the inputs are not kept (sign-mask path of wide checks) -- where
// they are, "|x_k| == m1" selects the same outputs: with a tie m2 == m1, so both choices coincide.
template <typename Real, bool TRACK_K0>
struct MinState {
    Real m1, m2;
    int k0;
    uint32_t par;
    __device__ __forceinline__ void init() {
        m1 = Arith<Real>::inf();
        m2 = Arith<Real>::inf();
        k0 = 0;
        par = 0;
    }
    __device__ __forceinline__ void push(Real x, int k) {
        Real a = Arith<Real>::abs(x);
        m2 = Arith<Real>::fmin_(m2, Arith<Real>::fmax_(m1, a));
        if (TRACK_K0) {
            if (a < m1) k0 = k;  // strict: first index wins ties
        }
        m1 = Arith<Real>::fmin_(m1, a);
        par ^= Arith<Real>::hi(x);
    }
};

// RCQ magnitude index (rcq_decoder.py:76-84).  NTH > 0: non-decreasing thresholds held in registers,
// index = number of thresholds j >= 1 that the magnitude reaches (== "last j reached").  NTH == 0:
// any table, read from shared memory.
template <int NTH>
struct Quantizer {
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
            for (int j = 0; j < NTH; ++j) t[j] = (j < nth_) ? s_thr_[j] : __int_as_float(0x7f800000);
        }
    }
    __device__ __forceinline__ uint32_t index(float mag) const {
        if constexpr (NTH > 0) {
            uint32_t idx = 0;
#pragma unroll
            for (int j = 1; j < NTH; ++j) idx += (mag >= t[j]) ? 1u : 0u;
            return idx;
        } else {
            return (uint32_t)quant_index(mag, s_thr, nth, mono);
        }
    }
};

// One edge whose beta is its own (type-1 weights over mixed variable degrees, per-edge N-NMS weights).
template <typename Real, bool QUANT, int NTH>
__device__ __forceinline__ typename CnOut<Real, QUANT>::type cn_emit(Real raw, Real beta, uint32_t signbits,
                                                                      const Quantizer<NTH>& qz, int bc) {
    Real val = Arith<Real>::flip(Arith<Real>::mul(beta, raw), signbits);
    if constexpr (QUANT) {
        float x = (float)val;
        uint32_t code = ((x < 0.f) ? (1u << (bc - 1)) : 0u) | qz.index(fabsf(x));
        return (uint8_t)code;
    } else {
        return val;
    }
}

// When every edge of the check shares one beta (Basic, RCQ, N-2D types 2-4, type 1 where a check sees
// one variable degree) a check has only TWO output magnitudes per frame, A = fl(beta*m1) for the edges
// other than the minimum and B = fl(beta*m2) for the minimum edge, so the multiply -- and for RCQ the
// threshold search -- runs twice per check instead of once per edge.
template <typename Real, bool QUANT>
struct CheckOut {
    using OutT = typename CnOut<Real, QUANT>::type;
    Real A, B;
    uint32_t ia, ib, ma, mb;  // RCQ: magnitude indices and sign-bit masks (0 when the value is +-0)
    uint32_t par;             // XOR of the input sign words (RCQ: also of beta's sign)
    int sh;
    template <int NTH>
    __device__ __forceinline__ void prepare(Real m1, Real m2, uint32_t par_, Real beta, bool has_beta,
                                            const Quantizer<NTH>& qz, int bc) {
        A = has_beta ? Arith<Real>::mul(beta, m1) : m1;
        B = has_beta ? Arith<Real>::mul(beta, m2) : m2;
        par = par_;
        if constexpr (QUANT) {
            const float a = (float)A, b = (float)B;
            const uint32_t S = 1u << (bc - 1);
            ia = qz.index(fabsf(a));
            ib = qz.index(fabsf(b));
            // code sign bit = (x < 0) needs a non-zero magnitude (rcq_decoder.py:87); A and B carry
            // beta's sign whenever they are non-zero, and B == 0 implies A == 0
            ma = (a != 0.f) ? S : 0u;
            mb = (b != 0.f) ? S : 0u;
            par = par_ ^ __float_as_uint(b);
            sh = 32 - bc;
        }
    }
    // is_min: this edge carries the check's minimum magnitude; sx: sign word of its own input
    __device__ __forceinline__ OutT emit(bool is_min, uint32_t sx) const {
        if constexpr (QUANT) {
            const uint32_t idx = is_min ? ib : ia, mask = is_min ? mb : ma;
            return (uint8_t)(idx | (((par ^ sx) >> sh) & mask));
        } else {
            return Arith<Real>::flip(is_min ? B : A, par ^ sx);
        }
    }
};

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
__device__ __forceinline__ void cn_check_small(const CnLaunch& p, int64_t slot0, int64_t f0, uint32_t dmask,
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
            for (int k = 0; k < DC; ++k) store_masked<OutT, V>(row_at(dst, (uint32_t)(slot0 + k), out_stride), out[k], dmask);
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
    for (int k = 0; k < DC; ++k) store_masked<OutT, V>(row_at(dst, (uint32_t)(slot0 + k), out_stride), out[k], dmask);
}

// Checks of degree 9..32: stream the inputs once, keeping min1/min2/first-argmin/parity and one sign
// bit per edge in a 32-bit shift register (funnel shift: one instruction per edge and frame).
template <typename Real, bool QUANT, int NTH>
__device__ void cn_check_mask32(const CnLaunch& p, int64_t slot0, int dc, int64_t f0, uint32_t dmask,
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
        store_masked<OutT, V>(dst + (slot0 + k) * p.Bp + f0, out, dmask);
    }
}

// Checks of degree > 32: same streaming pass, then the inputs are read again (they were just fetched)
// for their signs and for the "is the minimum" test.
template <typename Real, bool QUANT, int NTH>
__device__ void cn_check_reread(const CnLaunch& p, int64_t slot0, int dc, int64_t f0, uint32_t dmask,
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
        store_masked<OutT, V>(dst + (slot0 + k) * p.Bp + f0, out, dmask);
    }
}

// resident CTAs per SM the check-node kernel is compiled for: the RCQ variant is issue-bound and gains from
// 4 (64 registers); float32 / float64 stream at the HBM roofline with 3 / 2
#define LDPC_CN_BOUNDS __launch_bounds__(kThreads, QUANT ? 4 : (sizeof(Real) == 4 ? 3 : 2))
// FREEZE: stopped frames keep their c2v (forward()'s posterior output); otherwise the stores carry no mask code.
template <typename Real, bool QUANT, int NTH, bool FREEZE>
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
    const uint32_t dmask = FREEZE ? done_mask : 0u;   // otherwise nobody reads the messages of stopped frames
    const WorkItem it = p.items[item_id];
    int64_t slot = it.first_slot;
#define LDPC_CN_CASE(D)                                                                   \
    case D:                                                                               \
        for (int c = 0; c < it.count; ++c, slot += D)                                     \
            cn_check_small<Real, QUANT, NTH, D>(p, slot, f0, dmask, qz);                  \
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
                    cn_check_mask32<Real, QUANT, NTH>(p, slot, it.deg, f0, dmask, qz);
            } else {
                for (int c = 0; c < it.count; ++c, slot += it.deg)
                    cn_check_reread<Real, QUANT, NTH>(p, slot, it.deg, f0, dmask, qz);
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
#ifndef LDPC_WIDE_MINCTAS
#define LDPC_WIDE_MINCTAS 6
#endif
constexpr int kWideRows = LDPC_WIDE_ROWS;     // rows per slab
constexpr int kWideSlabs = LDPC_WIDE_SLABS;   // slabs in the ring
constexpr int kWideMinCtas = LDPC_WIDE_MINCTAS;
constexpr int kWideRowBytes = kWideThreads * 16;
constexpr size_t kWideSmem = (size_t)kWideSlabs * kWideRows * kWideRowBytes;

__device__ __forceinline__ uint32_t smem_addr(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
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

template <typename Real, bool QUANT, int NTH, typename MaskT>
__device__ __forceinline__ void cn_wide_check(const CnLaunch& p, RowRing<Real>& ring, int64_t slot0, int dc, int64_t f0,
                                              uint32_t dmask, bool active, const Quantizer<NTH>& qz) {
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
                        store_masked<OutT, V>(out_row, out, dmask);
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
        store_masked<OutT, V>(out_row, out, dmask);
    }
}

template <typename Real, bool QUANT, int NTH, bool FREEZE>
__global__ void __launch_bounds__(kWideThreads, kWideMinCtas) cn_wide_kernel(const CnLaunch p, const int nfb, const int item0) {
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
    const uint32_t dmask = FREEZE ? done_mask : 0u;
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
            cn_wide_check<Real, QUANT, NTH, uint32_t>(p, ring, slot, it.deg, f0, dmask, active, qz);
    } else {
        for (int c = 0; c < it.count; ++c, slot += it.deg)
            cn_wide_check<Real, QUANT, NTH, uint64_t>(p, ring, slot, it.deg, f0, dmask, active, qz);
    }
}

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

// ---------------------------------------------------------------------------------------------
// Offset min-sum check node (neural_minsum_decoder.py:236-253, neural_2d_decoder.py:383-401):
//   c2v = prod(other signs) * (relu(raw - beta) - alpha),   alpha indexed by the edge's VARIABLE degree
// Here the reference's three-valued sign() is visible (relu(0 - beta) - alpha need not be 0), so a zero
// among the OTHER inputs forces the output to 0: that is the case iff m2 == 0 (two zeros), or m1 == 0
// and this edge is not the zero one.  A degree-1 check has an empty product (= 1).
// ---------------------------------------------------------------------------------------------
template <typename Real>
__device__ __forceinline__ Real offset_value(Real raw, Real beta, bool has_beta, Real alpha, bool has_alpha,
                                             uint32_t signbits, bool zero_others) {
    Real t = has_beta ? Arith<Real>::add(raw, -beta) : raw;
    t = Arith<Real>::fmax_(t, Real(0));
    if (has_alpha) t = Arith<Real>::add(t, -alpha);
    t = Arith<Real>::flip(t, signbits);
    return zero_others ? Real(0) : t;
}

template <typename Real>
__device__ __forceinline__ void offset_weights(const CnLaunch& p, int64_t slot, Real beta_check, Real& beta, Real& alpha) {
    beta = beta_check;
    if (p.beta_t && p.beta_per_edge) beta = __ldg(static_cast<const Real*>(p.beta_t) + __ldg(p.bidx + slot));
    alpha = Real(0);
    if (p.alpha_t) alpha = __ldg(static_cast<const Real*>(p.alpha_t) + (p.aidx_slot ? __ldg(p.aidx_slot + slot) : 0));
}

template <typename Real, int DC>
__device__ __forceinline__ void cn_offset_small(const CnLaunch& p, int64_t slot0, int64_t f0, uint32_t dmask) {
    constexpr int V = FramesPerLane<Real>::value;
    const Real* __restrict__ src = static_cast<const Real*>(p.src);
    Real* __restrict__ dst = static_cast<Real*>(p.dst);
    Pack<Real, V> x[DC];
    Real beta[DC], alpha[DC];
    Real beta_check = Real(0);
    if (p.beta_t && !p.beta_per_edge) beta_check = __ldg(static_cast<const Real*>(p.beta_t) + (p.bidx ? __ldg(p.bidx + slot0) : 0));
#pragma unroll
    for (int k = 0; k < DC; ++k) {
        int64_t row = p.row_map ? (int64_t)__ldg(p.row_map + slot0 + k) : slot0 + k;
        x[k] = ld_stream<Pack<Real, V>>(src + row * p.Bp + f0);
        offset_weights<Real>(p, slot0 + k, beta_check, beta[k], alpha[k]);
    }
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
    for (int k = 0; k < DC; ++k) store_masked<Real, V>(dst + (slot0 + k) * p.Bp + f0, out[k], dmask);
}

template <typename Real>
__device__ void cn_offset_wide(const CnLaunch& p, int64_t slot0, int dc, int64_t f0, uint32_t dmask) {
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
        store_masked<Real, V>(dst + (slot0 + k) * p.Bp + f0, out, dmask);
    }
}

template <typename Real, bool FREEZE>
__global__ void __launch_bounds__(kThreads) cn_offset_kernel(const CnLaunch p, const int nfb) {
    constexpr int V = FramesPerLane<Real>::value;
    const int fb = blockIdx.x % nfb;
    const int item_id = blockIdx.x / nfb;
    const int64_t f0 = ((int64_t)fb * blockDim.x + threadIdx.x) * V;
    if (f0 >= p.Bp) return;
    const uint32_t done_mask = load_done_mask<V>(p.done, f0);
    if (__all_sync(0xffffffffu, done_mask == ((1u << V) - 1u))) return;
    const uint32_t dmask = FREEZE ? done_mask : 0u;
    const WorkItem it = p.items[item_id];
    int64_t slot = it.first_slot;
#define LDPC_CNO_CASE(D)                                                \
    case D:                                                             \
        for (int c = 0; c < it.count; ++c, slot += D)                   \
            cn_offset_small<Real, D>(p, slot, f0, dmask);               \
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
            for (int c = 0; c < it.count; ++c, slot += it.deg) cn_offset_wide<Real>(p, slot, it.deg, f0, dmask);
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
template <typename Real, bool QUANT, bool FINAL, int DV, int U>
__device__ __forceinline__ void vn_node_small(const VnLaunch& p, int32_t vpos, int64_t lbase, int64_t f0,
                                              uint32_t dmask, uint32_t smask, uint32_t keepw, int64_t wbase,
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
#pragma unroll
    for (int u = 0; u < U; ++u) {
#pragma unroll
        for (int d = 0; d < DV; ++d) cin[u][d] = ld_stream<Pack<InT, V>>(row_at(c2v, slot[u][d], in_stride));
        L[u] = ld_stream<Pack<Real, V>>(row_at(llr0, j[u], real_stride));
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
            for (int d = 0; d < DV; ++d) store_masked<Real, V>(row_at(v2c, slot[u][d], real_stride), out[d], smask);
        } else {
            if (p.postT) st_stream<Pack<Real, V>>(row_at(static_cast<Real*>(p.postT) + f0, j[u], real_stride), post);
        }
        write_hard<Real, V>(p.hardw, p.Wn, j[u], wbase, bit, keepw);
    }
}

// All variables of one work item, in groups of U (remainder one by one).
template <typename Real, bool QUANT, bool FINAL, int DV>
__device__ __forceinline__ void vn_item_small(const VnLaunch& p, const WorkItem& it, int64_t f0, uint32_t dmask,
                                              uint32_t smask, uint32_t keepw, int64_t wbase, const float* s_lut,
                                              const int (&lutbase)[FramesPerLane<Real>::value]) {
    // byte-wide RCQ code rows need more rows in flight than 16-byte float rows
    constexpr int U = QUANT ? ((DV <= 2) ? LDPC_VN_UQ_LO : ((DV <= 4) ? LDPC_VN_UQ_MID : LDPC_VN_UQ_HI))
                            : ((DV <= 2) ? LDPC_VN_UF_LO : ((DV <= 4) ? LDPC_VN_UF_MID : 1));
    int64_t lbase = it.first_slot;
    int32_t vpos = it.first_node;
    int c = 0;
    if constexpr (U > 1) {
        for (; c + U <= it.count; c += U, lbase += U * DV, vpos += U)
            vn_node_small<Real, QUANT, FINAL, DV, U>(p, vpos, lbase, f0, dmask, smask, keepw, wbase, s_lut, lutbase);
    }
    for (; c < it.count; ++c, lbase += DV, ++vpos)
        vn_node_small<Real, QUANT, FINAL, DV, 1>(p, vpos, lbase, f0, dmask, smask, keepw, wbase, s_lut, lutbase);
}

template <typename Real, bool QUANT, bool FINAL>
__device__ void vn_node_generic(const VnLaunch& p, int32_t vpos, int64_t lbase, int dv, int64_t f0, uint32_t dmask,
                                uint32_t smask, uint32_t keepw, int64_t wbase, const float* s_lut, const int (&lutbase)[FramesPerLane<Real>::value]) {
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
            store_masked<Real, V>(v2c + sd * p.Bp + f0, out, smask);
        }
    } else {
        if (p.postT) *reinterpret_cast<Pack<Real, V>*>(static_cast<Real*>(p.postT) + j * p.Bp + f0) = post;
    }
    write_hard<Real, V>(p.hardw, p.Wn, j, wbase, bit, keepw);
}

// FREEZE: stopped frames keep their v2c (forward()'s posterior output); otherwise their stores are plain.
template <typename Real, bool QUANT, bool FINAL, bool FREEZE>
__global__ void __launch_bounds__(kThreads, sizeof(Real) == 4 ? LDPC_VN_F32_MINCTAS : 3) vn_kernel(const VnLaunch p, const int nfb,
                                                                                                   const int item0) {
    constexpr int V = FramesPerLane<Real>::value;
    extern __shared__ float s_lut[];
    if (QUANT) {
        const int nl = p.n_quant << p.bc;
        for (int i = threadIdx.x; i < nl; i += blockDim.x) s_lut[i] = p.lut[i];
        __syncthreads();
    }
    const int fb = blockIdx.x % nfb;
    const int item_id = item0 + blockIdx.x / nfb;
    const int64_t f0 = ((int64_t)fb * blockDim.x + threadIdx.x) * V;
    if (f0 >= p.Bp) return;
    // Stopped frames keep their messages and their packed decisions in place; only forward()'s posterior
    // output makes the final pass recompute them (posterior of the iteration a frame stopped at, from its
    // frozen c2v).
    uint32_t dmask = 0;
    if (!FINAL || p.postT == nullptr) {
        dmask = load_done_mask<V>(p.done, f0);
        if (__all_sync(0xffffffffu, dmask == ((1u << V) - 1u))) return;
    }
    const uint32_t keepw = keep_word<V>(dmask);
    const uint32_t smask = FREEZE ? dmask : 0u;
    int lutbase[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        lutbase[v] = 0;
        if (QUANT) {
            int q = p.q_now;
            if (FINAL) {
                int itv = __ldg(p.iters + f0 + v);
                q = __ldg(p.q_of_iter + (itv > 0 ? itv - 1 : 0));
            }
            lutbase[v] = q << p.bc;
        }
    }
    const int64_t warp_f0 = f0 - (int64_t)(threadIdx.x & 31) * V;
    const int64_t wbase = (warp_f0 / (32 * V)) * V;
    const WorkItem it = p.items[item_id];
    int64_t lbase = it.first_slot;
    int32_t vpos = it.first_node;
#define LDPC_VN_CASE(D)                                                                               \
    case D:                                                                                           \
        vn_item_small<Real, QUANT, FINAL, D>(p, it, f0, dmask, smask, keepw, wbase, s_lut, lutbase);  \
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
                vn_node_generic<Real, QUANT, FINAL>(p, vpos, lbase, it.deg, f0, dmask, smask, keepw, wbase, s_lut, lutbase);
    }
#undef LDPC_VN_CASE
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
constexpr int kVnWideMaxDeg = 64;

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
                                             int64_t lbase, Real* __restrict__ v2c, uint32_t real_stride, uint32_t smask,
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
            store_masked<Real, V>(row_at(v2c, sd, real_stride), out, smask);
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

template <typename Real, bool QUANT, bool FINAL, bool FREEZE>
__global__ void __launch_bounds__(kVnWideThreads) vn_wide_kernel(const VnLaunch p, const int nfb, const int item0,
                                                                  const int stage_rows) {
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
    const int item_id = item0 + blockIdx.x / nfb;
    const int64_t f0 = ((int64_t)fb * kVnWideThreads + threadIdx.x) * V;
    if (f0 >= p.Bp) return;   // whole warps
    uint32_t dmask = 0;
    if (!FINAL || p.postT == nullptr) {
        dmask = load_done_mask<V>(p.done, f0);
        if (__all_sync(0xffffffffu, dmask == ((1u << V) - 1u))) return;
    }
    const uint32_t keepw = keep_word<V>(dmask);
    const uint32_t smask = FREEZE ? dmask : 0u;
    int lutbase[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        lutbase[v] = 0;
        if (QUANT) {
            int q = p.q_now;
            if (FINAL) {
                const int itv = __ldg(p.iters + f0 + v);
                q = __ldg(p.q_of_iter + (itv > 0 ? itv - 1 : 0));
            }
            lutbase[v] = q << p.bc;
        }
    }
    const uint32_t lutmask = (1u << p.bc) - 1u;
    const int64_t warp_f0 = f0 - (int64_t)(threadIdx.x & 31) * V;
    const int64_t wbase = (warp_f0 / (32 * V)) * V;
    const WorkItem it = p.items[item_id];
    const int dv = it.deg;
    const uint32_t in_stride = (uint32_t)p.Bp * (uint32_t)sizeof(InT), real_stride = (uint32_t)p.Bp * (uint32_t)sizeof(Real);
    const InT* __restrict__ c2v = static_cast<const InT*>(p.c2v) + f0;
    Real* __restrict__ v2c = static_cast<Real*>(p.v2c) + f0;
    const Real* __restrict__ llr0 = static_cast<const Real*>(p.llrT) + f0;
    const bool has_alpha = (p.alpha_t != nullptr) && !FINAL;
    int64_t lbase = it.first_slot;
    int32_t vpos = it.first_node;
    for (int c = 0; c < it.count; ++c, lbase += dv, ++vpos) {
        const uint32_t j = (uint32_t)__ldg(p.vpos_var + vpos);
        for (int i = 0; i < dv; ++i) {
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
        vn_wide_sums<Real, FINAL, D>(p, s_val, D, L, alpha, has_alpha, lbase, v2c, real_stride, smask, dmask, post, bit); \
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
        if (!handled) vn_wide_sums<Real, FINAL, 0>(p, s_val, dv, L, alpha, has_alpha, lbase, v2c, real_stride, smask, dmask, post, bit);
        if constexpr (FINAL) {
            if (p.postT) st_stream<PackR>(row_at(static_cast<Real*>(p.postT) + f0, j, real_stride), post);
        }
        write_hard<Real, V>(p.hardw, p.Wn, j, wbase, bit, keepw);
    }
}

// ---------------------------------------------------------------------------------------------
// Syndrome on packed hard decisions (ldpc_decoder.py:141): one thread = one 32-frame word,
// XOR over a check's variables, OR over the item's checks, atomicOr into unsat[w].
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) syn_kernel(const SynLaunch p, const int nwb) {
    const int wb = blockIdx.x % nwb;
    const int item_id = blockIdx.x / nwb;
    const int64_t w = (int64_t)wb * blockDim.x + threadIdx.x;
    if (w >= p.Wn) return;
    const WorkItem it = p.items[item_id];
    int64_t slot = it.first_slot;
    uint32_t acc = 0;
    for (int c = 0; c < it.count; ++c, slot += it.deg) {
        uint32_t syn = 0;
        for (int k = 0; k < it.deg; ++k) {
            int64_t j = __ldg(p.slot_var + slot + k);
            syn ^= __ldg(p.hardw + j * p.Wn + w);
        }
        acc |= syn;
    }
    if (acc) atomicOr(p.unsat + w, acc);
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

// ldpc_decoder.py:143-144: first iteration whose syndrome is all-zero ends the frame.
__global__ void commit_kernel(int V, const uint32_t* __restrict__ unsat, uint32_t* __restrict__ unsat_next,
                              uint8_t* __restrict__ done, int32_t* __restrict__ iters,
                              uint8_t* __restrict__ success, int32_t t1, int64_t Bp) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= Bp) return;
    int64_t w;
    int bit;
    frame_to_wordbit(f, V, w, bit);
    if (!done[f]) {
        if (!((unsat[w] >> bit) & 1u)) {
            done[f] = 1;
            iters[f] = t1;
            success[f] = 1;
        }
    }
    if (unsat_next && bit == 0) unsat_next[w] = 0;
}

__global__ void reset_kernel(uint8_t* __restrict__ done, int32_t* __restrict__ iters, uint8_t* __restrict__ success,
                             uint32_t* __restrict__ unsat2, int64_t B, int64_t Bp, int32_t T) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (f >= Bp) return;
    done[f] = (f >= B) ? 1 : 0;  // pad frames never run
    iters[f] = T;
    success[f] = 0;
    if (f < 2 * (Bp / 32)) unsat2[f] = 0;
}

// ---------------------------------------------------------------------------------------------
// Layout conversion: user [B][n] row-major <-> interleaved [n][Bp]
// ---------------------------------------------------------------------------------------------
// 64 frames x 64 variables per CTA; both the row-major reads and the interleaved writes move two elements per
// lane (256- / 512-byte segments per warp instead of 128 / 256).  PAIR needs an even n (aligned row starts).
template <typename Real, bool PAIR>
__global__ void __launch_bounds__(256) pack_kernel(const Real* __restrict__ llr, Real* __restrict__ llrT, int64_t B,
                                                    int64_t Bp, int32_t n) {
    __shared__ Real tile[64][65];   // [variable][frame]
    const int64_t f_base = (int64_t)blockIdx.x * 64;
    const int32_t j_base = blockIdx.y * 64;
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    for (int r = wy; r < 64; r += 8) {
        const int64_t f = f_base + r;
        const int32_t j = j_base + 2 * lane;
        Real a = Real(0), b = Real(0);
        if (f < B) {
            if (PAIR) {
                if (j < n) {   // n even: j + 1 < n as well
                    const Pack<Real, 2> v = *reinterpret_cast<const Pack<Real, 2>*>(llr + f * n + j);
                    a = v.v[0];
                    b = v.v[1];
                }
            } else {
                if (j < n) a = llr[f * n + j];
                if (j + 1 < n) b = llr[f * n + j + 1];
            }
        }
        tile[2 * lane][r] = a;
        tile[2 * lane + 1][r] = b;
    }
    __syncthreads();
    for (int c = wy; c < 64; c += 8) {
        const int32_t j = j_base + c;
        const int64_t f = f_base + 2 * lane;   // Bp is a multiple of 128: f + 1 < Bp whenever f < Bp
        if (j < n && f < Bp) {
            Pack<Real, 2> v;
            v.v[0] = tile[c][2 * lane];
            v.v[1] = tile[c][2 * lane + 1];
            *reinterpret_cast<Pack<Real, 2>*>(llrT + (int64_t)j * Bp + f) = v;
        }
    }
}

template <typename Real>
__global__ void unpack_post_kernel(const Real* __restrict__ postT, Real* __restrict__ post, int64_t B, int64_t Bp,
                                   int32_t n, const int32_t* __restrict__ map) {
    __shared__ Real tile[32][33];
    const int64_t f_base = (int64_t)blockIdx.x * 32;
    const int32_t j_base = blockIdx.y * 32;
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int32_t j = j_base + r;
        int64_t f = f_base + threadIdx.x;
        tile[r][threadIdx.x] = (j < n && f < Bp) ? postT[(int64_t)j * Bp + f] : Real(0);
    }
    __syncthreads();
    for (int r = threadIdx.y; r < 32; r += blockDim.y) {
        int64_t f = f_base + r;
        int32_t j = j_base + threadIdx.x;
        if (f < B && j < n) post[(map ? (int64_t)map[f] : f) * n + j] = tile[threadIdx.x][r];
    }
}

// One warp: one hard word (32 frames) x 128 variables; a lane holds the words of 4 consecutive variables and
// writes 4 bytes per frame (128-byte rows per warp store).  n % 4 != 0 falls back to byte stores at the tail.
__global__ void unpack_bits_kernel(int V, const uint32_t* __restrict__ hardw, int64_t Wn, uint8_t* __restrict__ bits,
                                   int64_t B, int32_t n, const int32_t* __restrict__ map) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= Wn) return;
    const int32_t j = blockIdx.x * 128 + lane * 4;
    uint32_t word[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) word[i] = (j + i < n) ? __ldg(hardw + (int64_t)(j + i) * Wn + w) : 0u;
    const bool vec = (n % 4 == 0) && (j + 3 < n) && ((reinterpret_cast<uintptr_t>(bits) & 3u) == 0);
    for (int b = 0; b < 32; ++b) {
        const int64_t f = wordbit_to_frame(w, b, V);
        if (f >= B) continue;
        uint8_t* row = bits + (map ? (int64_t)map[f] : f) * n + j;
        if (vec) {
            const uint32_t v = ((word[0] >> b) & 1u) | (((word[1] >> b) & 1u) << 8) | (((word[2] >> b) & 1u) << 16) |
                               (((word[3] >> b) & 1u) << 24);
            *reinterpret_cast<uint32_t*>(row) = v;
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (j + i < n) row[i] = (uint8_t)((word[i] >> b) & 1u);
        }
    }
}

// ---------------------------------------------------------------------------------------------
// AWGN channel (ldpc_decoder.py:286-302) with counter-based Philox4x32-10 noise.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                                              uint32_t k1, uint32_t (&out)[4]) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        uint32_t n0 = hi1 ^ c1 ^ k0, n1 = lo1, n2 = hi0 ^ c3 ^ k1, n3 = lo0;
        c0 = n0; c1 = n1; c2 = n2; c3 = n3;
        k0 += 0x9E3779B9u;
        k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

__device__ __forceinline__ float u01(uint32_t x) { return (float)x * 2.3283064365386963e-10f + 1.1641532182693481e-10f; }

template <typename Real, bool ROW_MAJOR>
__global__ void awgn_kernel(void* __restrict__ out_, int32_t n, int64_t B, int64_t Bp, uint64_t frame0, uint64_t seed,
                            float sigma, float inv_sigma2_x2, float llr_sign, const uint8_t* __restrict__ codeword) {
    const int64_t f0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4;
    const int32_t jg = blockIdx.y;
    const int64_t Flim = ROW_MAJOR ? B : Bp;
    if (f0 >= Flim) return;
    float val[4][4];  // [variable in group][frame]
#pragma unroll
    for (int v = 0; v < 4; ++v) {
        uint64_t gf = frame0 + (uint64_t)(f0 + v);
        uint32_t r[4];
        philox4x32_10((uint32_t)jg, (uint32_t)gf, (uint32_t)(gf >> 32), 0x4c445043u, (uint32_t)seed,
                      (uint32_t)(seed >> 32), r);
        float z[4];
        {
            float rad = sqrtf(-2.f * logf(u01(r[0])));
            float s, c;
            sincospif(2.f * u01(r[1]), &s, &c);
            z[0] = rad * c;
            z[1] = rad * s;
            rad = sqrtf(-2.f * logf(u01(r[2])));
            sincospif(2.f * u01(r[3]), &s, &c);
            z[2] = rad * c;
            z[3] = rad * s;
        }
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            int32_t j = 4 * jg + i;
            float cw = (codeword && j < n) ? (float)codeword[j] : 0.f;
            float sym = llr_sign * (1.f - 2.f * cw);  // ldpc_decoder.py:289 with the chosen convention
            float y = __fadd_rn(sym, __fmul_rn(sigma, z[i]));
            val[i][v] = __fmul_rn(y, inv_sigma2_x2);
            if (!ROW_MAJOR && (f0 + v) >= B) val[i][v] = 0.f;  // pad frames
        }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        int32_t j = 4 * jg + i;
        if (j >= n) break;
        if (ROW_MAJOR) {
            float* out = static_cast<float*>(out_);
#pragma unroll
            for (int v = 0; v < 4; ++v)
                if (f0 + v < B) out[(f0 + v) * n + j] = val[i][v];
        } else {
            Real* out = static_cast<Real*>(out_) + (int64_t)j * Bp + f0;
#pragma unroll
            for (int v = 0; v < 4; ++v) out[v] = (Real)val[i][v];
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Error counting (simulation_framework.py:125-131)
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void accumulate_counters(int64_t* counters, int ferr, int berr, int iters, int valid) {
    unsigned fe = __reduce_add_sync(0xffffffffu, (unsigned)ferr);
    unsigned be = __reduce_add_sync(0xffffffffu, (unsigned)berr);
    unsigned it = __reduce_add_sync(0xffffffffu, (unsigned)iters);
    unsigned nf = __reduce_add_sync(0xffffffffu, (unsigned)valid);
    if ((threadIdx.x & 31) == 0) {
        unsigned long long* c = reinterpret_cast<unsigned long long*>(counters);
        if (fe) atomicAdd(c + 0, (unsigned long long)fe);
        if (be) atomicAdd(c + 1, (unsigned long long)be);
        if (it) atomicAdd(c + 2, (unsigned long long)it);
        if (nf) atomicAdd(c + 3, (unsigned long long)nf);
    }
}

// Bit errors per frame from the packed decisions.  Pass 1: a thread owns one 32-frame word and a chunk of
// variables (coalesced word loads), counts per bit in registers and adds the non-zero counts to
// frame_cnt[frame]; pass 2: one thread per frame folds them into the counters / per-frame outputs.
constexpr int kCountVarChunk = 128;

__global__ void __launch_bounds__(128) count_partial_kernel(int V, const uint32_t* __restrict__ hardw, int64_t Wn, int32_t n,
                                                             const uint8_t* __restrict__ codeword,
                                                             int32_t* __restrict__ frame_cnt) {
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= Wn) return;
    const int32_t j0 = blockIdx.y * kCountVarChunk, j1 = min(n, j0 + kCountVarChunk);
    // vertical (bit-sliced) counters: plane[k] holds bit k of the 32 per-frame counts (kCountVarChunk < 256)
    uint32_t plane[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) plane[k] = 0;
    for (int32_t j = j0; j < j1; ++j) {
        uint32_t carry = __ldg(hardw + (int64_t)j * Wn + w);
        if (codeword && __ldg(codeword + j)) carry = ~carry;
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const uint32_t t = plane[k] & carry;
            plane[k] ^= carry;
            carry = t;
        }
    }
    uint32_t any = 0;
#pragma unroll
    for (int k = 0; k < 8; ++k) any |= plane[k];
    while (any) {
        const int b = __ffs(any) - 1;
        any &= any - 1;
        int cnt = 0;
#pragma unroll
        for (int k = 0; k < 8; ++k) cnt |= (int)((plane[k] >> b) & 1u) << k;
        atomicAdd(frame_cnt + wordbit_to_frame(w, b, V), cnt);
    }
}

__global__ void count_final_kernel(const int32_t* __restrict__ frame_cnt, int64_t B, const int32_t* __restrict__ iters,
                                   int64_t* counters, int32_t* __restrict__ frame_bit_errors,
                                   int32_t* __restrict__ frame_iters, const int32_t* __restrict__ map,
                                   const uint8_t* __restrict__ only_done) {
    const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    const int valid = f < B && !(only_done && !only_done[f]);   // frames handed on to a compacted level are counted there
    const int cnt = valid ? frame_cnt[f] : 0;
    const int it = valid ? iters[f] : 0;
    if (valid) {
        const int64_t fo = map ? (int64_t)map[f] : f;
        if (frame_bit_errors) frame_bit_errors[fo] = cnt;
        if (frame_iters) frame_iters[fo] = it;
    }
    accumulate_counters(counters, valid && cnt > 0, cnt, it, valid);
}

__global__ void count_bits_kernel(const uint8_t* __restrict__ bits, int32_t n, int64_t B,
                                  const uint8_t* __restrict__ codeword, const int32_t* __restrict__ iters,
                                  int64_t* counters, int32_t* __restrict__ frame_bit_errors) {
    const int lane = threadIdx.x & 31;
    const int64_t f = (int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (f >= B) return;
    int cnt = 0;
    for (int32_t j = lane; j < n; j += 32) {
        uint8_t b = bits[f * n + j];
        uint8_t c = codeword ? codeword[j] : 0;
        cnt += (b != c);
    }
    cnt = (int)__reduce_add_sync(0xffffffffu, (unsigned)cnt);
    if (lane == 0) {
        if (frame_bit_errors) frame_bit_errors[f] = cnt;
        unsigned long long* c = reinterpret_cast<unsigned long long*>(counters);
        if (cnt) {
            atomicAdd(c + 0, 1ull);
            atomicAdd(c + 1, (unsigned long long)cnt);
        }
        if (iters) atomicAdd(c + 2, (unsigned long long)iters[f]);
        atomicAdd(c + 3, 1ull);
    }
}

// ---------------------------------------------------------------------------------------------
// Frame compaction ("early stop at scale").  Lanes own fixed frames, so a batch in which
// most frames have stopped still streams every message row of every warp that holds one running frame.  At
// checkpoints the frames still running are gathered (LLRs and V2C state, column by column) into a smaller
// dense batch that carries on from the same iteration; these kernels do the bookkeeping.
// ---------------------------------------------------------------------------------------------
constexpr int kScanBlock = 1024;

// counts[b] = frames of block b (kScanBlock frames) that are still running
__global__ void __launch_bounds__(kScanBlock) pending_count_kernel(const uint8_t* __restrict__ done, int64_t Bp,
                                                                    int32_t* __restrict__ counts) {
    __shared__ int32_t s_warp[kScanBlock / 32];
    const int64_t f = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    const int pend = (f < Bp) ? (done[f] == 0) : 0;
    const unsigned c = __reduce_add_sync(0xffffffffu, (unsigned)pend);
    if ((threadIdx.x & 31) == 0) s_warp[threadIdx.x >> 5] = (int32_t)c;
    __syncthreads();
    if (threadIdx.x < 32) {
        const unsigned t = __reduce_add_sync(0xffffffffu, (unsigned)s_warp[threadIdx.x]);
        if (threadIdx.x == 0) counts[blockIdx.x] = (int32_t)t;
    }
}

// in-place exclusive scan of counts[nb] by one block; total[0] = sum
__global__ void __launch_bounds__(kScanBlock) pending_scan_kernel(int32_t* __restrict__ counts, int nb,
                                                                   int32_t* __restrict__ total) {
    __shared__ int32_t s_warp[kScanBlock / 32];
    __shared__ int32_t s_carry;
    if (threadIdx.x == 0) s_carry = 0;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int base = 0; base < nb; base += kScanBlock) {
        const int i = base + threadIdx.x;
        const int32_t x = (i < nb) ? counts[i] : 0;
        int32_t incl = x;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t y = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += y;
        }
        if (lane == 31) s_warp[warp] = incl;
        __syncthreads();
        if (warp == 0) {
            int32_t w = s_warp[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
                if (lane >= o) w += y;
            }
            s_warp[lane] = w;   // inclusive over warps
        }
        __syncthreads();
        const int32_t carry = s_carry;
        const int32_t before = carry + (warp ? s_warp[warp - 1] : 0) + incl - x;
        if (i < nb) counts[i] = before;
        __syncthreads();
        if (threadIdx.x == kScanBlock - 1) s_carry = before + x;
        __syncthreads();
    }
    if (threadIdx.x == 0) total[0] = s_carry;
}

// idx[offsets[b] + r] = r-th running frame of block b (ascending frame order overall)
__global__ void __launch_bounds__(kScanBlock) pending_index_kernel(const uint8_t* __restrict__ done, int64_t Bp,
                                                                    const int32_t* __restrict__ offsets,
                                                                    int32_t* __restrict__ idx) {
    __shared__ int32_t s_warp[kScanBlock / 32];
    const int64_t f = (int64_t)blockIdx.x * kScanBlock + threadIdx.x;
    const int pend = (f < Bp) ? (done[f] == 0) : 0;
    const unsigned bal = __ballot_sync(0xffffffffu, pend);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (lane == 0) s_warp[warp] = __popc(bal);
    __syncthreads();
    if (warp == 0) {
        int32_t w = s_warp[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int32_t y = __shfl_up_sync(0xffffffffu, w, o);
            if (lane >= o) w += y;
        }
        s_warp[lane] = w;
    }
    __syncthreads();
    if (pend) {
        const int32_t pos = offsets[blockIdx.x] + (warp ? s_warp[warp - 1] : 0) + __popc(bal & ((1u << lane) - 1u));
        idx[pos] = (int32_t)f;
    }
}

// dst[j][i] = src[j][idx[i]] for i < count, 0 for the pad frames; rows j stride over gridDim.y
template <typename Real>
__global__ void gather_cols_kernel(const Real* __restrict__ src, int64_t Bp_src, Real* __restrict__ dst, int64_t Bp_dst,
                                   const int32_t* __restrict__ idx, int64_t count, int64_t rows) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Bp_dst) return;
    const int64_t col = (i < count) ? (int64_t)__ldg(idx + i) : -1;
    for (int64_t j = blockIdx.y; j < rows; j += gridDim.y)
        dst[j * Bp_dst + i] = (col >= 0) ? __ldg(src + j * Bp_src + col) : Real(0);
}

__global__ void scatter_frames_kernel(const int32_t* __restrict__ iters_src, const uint8_t* __restrict__ succ_src,
                                      int32_t* __restrict__ iters_dst, uint8_t* __restrict__ succ_dst,
                                      const int32_t* __restrict__ map, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const int64_t f = map ? (int64_t)map[i] : i;
    if (iters_dst) iters_dst[f] = iters_src[i];
    if (succ_dst) succ_dst[f] = succ_src[i];
}

__global__ void compose_map_kernel(const int32_t* __restrict__ idx, const int32_t* __restrict__ parent_map,
                                   int32_t* __restrict__ out, int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) out[i] = parent_map ? parent_map[idx[i]] : idx[i];
}

inline int threads_for(int64_t Bp, int V) {
    int64_t lanes = Bp / V;
    int t = (int)(lanes < kThreads ? lanes : kThreads);
    t = (t + 31) / 32 * 32;
    return t < 32 ? 32 : t;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------
namespace {

template <typename Real, bool QUANT, int NTH, bool FREEZE>
cudaError_t launch_cn_range(const CnLaunch& p, int item0, int item1, bool wide, cudaStream_t stream) {
    if (item1 <= item0) return cudaSuccess;
    constexpr int V = FramesPerLane<Real>::value;
    if (wide) {
        // per device and cheap, so simply repeated on every launch
        cudaError_t e = cudaFuncSetAttribute(cn_wide_kernel<Real, QUANT, NTH, FREEZE>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)kWideSmem);
        if (e != cudaSuccess) return e;
        const int64_t nfb = (p.Bp + (int64_t)kWideThreads * V - 1) / ((int64_t)kWideThreads * V);
        const int64_t grid = nfb * (item1 - item0);
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        cn_wide_kernel<Real, QUANT, NTH, FREEZE><<<(unsigned)grid, kWideThreads, kWideSmem, stream>>>(p, (int)nfb, item0);
    } else {
        const int threads = threads_for(p.Bp, V);
        const int64_t nfb = (p.Bp / V + threads - 1) / threads;
        const int64_t grid = nfb * (item1 - item0);
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        cn_kernel<Real, QUANT, NTH, FREEZE><<<(unsigned)grid, threads, 0, stream>>>(p, (int)nfb, item0);
    }
    return cudaGetLastError();
}

template <typename Real, bool QUANT, int NTH, bool FREEZE>
cudaError_t launch_cn_frz(const CnLaunch& p, cudaStream_t stream) {
    // items are sorted by degree: [0, wide0) degree <= 8, [wide0, wide1) degree 9..64 (row ring), rest > 64
    const int wide0 = p.wide_ring ? p.items_wide_begin : p.n_items, wide1 = p.wide_ring ? p.items_wide_end : p.n_items;
    cudaError_t e = launch_cn_range<Real, QUANT, NTH, FREEZE>(p, 0, wide0, false, stream);
    if (e == cudaSuccess) e = launch_cn_range<Real, QUANT, NTH, FREEZE>(p, wide0, wide1, true, stream);
    if (e == cudaSuccess) e = launch_cn_range<Real, QUANT, NTH, FREEZE>(p, wide1, p.n_items, false, stream);
    return e;
}

template <typename Real, bool QUANT, int NTH>
cudaError_t launch_cn_all(const CnLaunch& p, cudaStream_t stream) {
    return p.freeze ? launch_cn_frz<Real, QUANT, NTH, true>(p, stream) : launch_cn_frz<Real, QUANT, NTH, false>(p, stream);
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

cudaError_t launch_cn_offset(int dtype, const CnLaunch& p, cudaStream_t stream) {
    if (p.n_items == 0) return cudaSuccess;
    const int V = dtype == 0 ? 4 : 2;
    const int threads = threads_for(p.Bp, V);
    const int64_t nfb = (p.Bp / V + threads - 1) / threads;
    const int64_t grid = nfb * p.n_items;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    if (dtype == 0) {
        if (p.freeze) cn_offset_kernel<float, true><<<(unsigned)grid, threads, 0, stream>>>(p, (int)nfb);
        else cn_offset_kernel<float, false><<<(unsigned)grid, threads, 0, stream>>>(p, (int)nfb);
    } else {
        if (p.freeze) cn_offset_kernel<double, true><<<(unsigned)grid, threads, 0, stream>>>(p, (int)nfb);
        else cn_offset_kernel<double, false><<<(unsigned)grid, threads, 0, stream>>>(p, (int)nfb);
    }
    return cudaGetLastError();
}

cudaError_t launch_layered_iter(float* P, const int64_t* chk_ptr, const int32_t* chk_var, int32_t m, const float* thr,
                                int nth, int bc, int mono, const uint8_t* done, int64_t Bp, cudaStream_t stream) {
    const int threads = (int)(Bp < 128 ? Bp : 128);
    layered_iter_kernel<<<(unsigned)((Bp + threads - 1) / threads), threads, 0, stream>>>(P, chk_ptr, chk_var, m, thr, nth,
                                                                                            bc, mono, done, Bp);
    return cudaGetLastError();
}

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
    if (wide) {
        using InT = typename CnOut<Real, QUANT>::type;
        const int rows = p.wide_max_deg;
        const size_t smem = (size_t)rows * kVnWideThreads * (sizeof(Pack<Real, V>) + (QUANT ? sizeof(Pack<InT, V>) : 0)) +
                            (QUANT ? sizeof(float) * ((size_t)p.n_quant << p.bc) : 0);
        const int64_t nfb = (p.Bp + (int64_t)kVnWideThreads * V - 1) / ((int64_t)kVnWideThreads * V);
        const int64_t grid = nfb * (item1 - item0);
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
#define LDPC_VNW(FINAL, FREEZE)                                                                                        \
    do {                                                                                                               \
        cudaError_t e_ = cudaFuncSetAttribute(vn_wide_kernel<Real, QUANT, FINAL, FREEZE>,                              \
                                              cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                 \
        if (e_ != cudaSuccess) return e_;                                                                              \
        vn_wide_kernel<Real, QUANT, FINAL, FREEZE><<<(unsigned)grid, kVnWideThreads, smem, stream>>>(p, (int)nfb, item0, rows); \
    } while (0)
        if (p.final_pass) LDPC_VNW(true, false);
        else if (p.freeze) LDPC_VNW(false, true);
        else LDPC_VNW(false, false);
#undef LDPC_VNW
    } else {
        const int threads = threads_for(p.Bp, V);
        const int64_t nfb = (p.Bp / V + threads - 1) / threads;
        const int64_t grid = nfb * (item1 - item0);
        if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
        const size_t smem = p.bc ? sizeof(float) * ((size_t)p.n_quant << p.bc) : 0;
        const unsigned g = (unsigned)grid;
        const int nf = (int)nfb;
        if (p.final_pass) vn_kernel<Real, QUANT, true, false><<<g, threads, smem, stream>>>(p, nf, item0);
        else if (p.freeze) vn_kernel<Real, QUANT, false, true><<<g, threads, smem, stream>>>(p, nf, item0);
        else vn_kernel<Real, QUANT, false, false><<<g, threads, smem, stream>>>(p, nf, item0);
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

cudaError_t launch_syndrome(const SynLaunch& p, cudaStream_t stream) {
    if (p.n_items == 0) return cudaSuccess;
    const int threads = 128;
    const int64_t nwb = (p.Wn + threads - 1) / threads;
    const int64_t grid = nwb * p.n_items;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    syn_kernel<<<(unsigned)grid, threads, 0, stream>>>(p, (int)nwb);
    return cudaGetLastError();
}

cudaError_t launch_commit(int V, const uint32_t* unsat, uint32_t* unsat_next, uint8_t* done, int32_t* iters,
                          uint8_t* success, int32_t t1, int64_t Bp, cudaStream_t stream) {
    const int threads = 256;
    commit_kernel<<<(unsigned)((Bp + threads - 1) / threads), threads, 0, stream>>>(V, unsat, unsat_next, done, iters,
                                                                                      success, t1, Bp);
    return cudaGetLastError();
}

cudaError_t launch_reset_state(uint8_t* done, int32_t* iters, uint8_t* success, uint32_t* unsat2, int64_t B,
                               int64_t Bp, int32_t T, cudaStream_t stream) {
    const int threads = 256;
    reset_kernel<<<(unsigned)((Bp + threads - 1) / threads), threads, 0, stream>>>(done, iters, success, unsat2, B, Bp, T);
    return cudaGetLastError();
}

cudaError_t launch_pack(int dtype, const void* llr, void* llrT, int64_t B, int64_t Bp, int32_t n, uint8_t* done,
                        int32_t* iters, uint8_t* success, int32_t T, cudaStream_t stream) {
    (void)done; (void)iters; (void)success; (void)T;
    dim3 grid((unsigned)((Bp + 63) / 64), (unsigned)((n + 63) / 64));
    const bool pair = (n % 2 == 0) && (reinterpret_cast<uintptr_t>(llr) % (dtype == 0 ? 8 : 16) == 0);
    if (dtype == 0) {
        if (pair) pack_kernel<float, true><<<grid, 256, 0, stream>>>(static_cast<const float*>(llr), static_cast<float*>(llrT), B, Bp, n);
        else pack_kernel<float, false><<<grid, 256, 0, stream>>>(static_cast<const float*>(llr), static_cast<float*>(llrT), B, Bp, n);
    } else {
        if (pair) pack_kernel<double, true><<<grid, 256, 0, stream>>>(static_cast<const double*>(llr), static_cast<double*>(llrT), B, Bp, n);
        else pack_kernel<double, false><<<grid, 256, 0, stream>>>(static_cast<const double*>(llr), static_cast<double*>(llrT), B, Bp, n);
    }
    return cudaGetLastError();
}

cudaError_t launch_unpack_bits(int V, const uint32_t* hardw, int64_t Wn, uint8_t* bits, int64_t B, int32_t n,
                               const int32_t* map, cudaStream_t stream) {
    const int warps = 8;
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)((Wn + warps - 1) / warps));
    unpack_bits_kernel<<<grid, warps * 32, 0, stream>>>(V, hardw, Wn, bits, B, n, map);
    return cudaGetLastError();
}

cudaError_t launch_unpack_post(int dtype, const void* postT, void* post, int64_t B, int64_t Bp, int32_t n,
                               const int32_t* map, cudaStream_t stream) {
    dim3 block(32, 8);
    dim3 grid((unsigned)((B + 31) / 32), (unsigned)((n + 31) / 32));
    if (dtype == 0) unpack_post_kernel<float><<<grid, block, 0, stream>>>(static_cast<const float*>(postT), static_cast<float*>(post), B, Bp, n, map);
    else unpack_post_kernel<double><<<grid, block, 0, stream>>>(static_cast<const double*>(postT), static_cast<double*>(post), B, Bp, n, map);
    return cudaGetLastError();
}

cudaError_t launch_awgn(int dtype, int row_major, void* out, int32_t n, int64_t B, int64_t Bp, uint64_t frame0,
                        uint64_t seed, float snr_db, int32_t llr_sign, const uint8_t* codeword, cudaStream_t stream) {
    // ldpc_decoder.py:292-300: noise_power = 1 / 10^(snr/10); llr = 2 * received / noise_power
    const double sigma2 = 1.0 / pow(10.0, (double)snr_db / 10.0);
    const float sigma = (float)sqrt(sigma2);
    const float k = (float)(2.0 / sigma2);
    const float sgn = llr_sign >= 0 ? 1.f : -1.f;
    const int64_t F = row_major ? B : Bp;
    const int threads = 128;
    dim3 grid((unsigned)((F / 4 + (F % 4 != 0) + threads - 1) / threads), (unsigned)((n + 3) / 4));
    if (row_major) awgn_kernel<float, true><<<grid, threads, 0, stream>>>(out, n, B, Bp, frame0, seed, sigma, k, sgn, codeword);
    else if (dtype == 0) awgn_kernel<float, false><<<grid, threads, 0, stream>>>(out, n, B, Bp, frame0, seed, sigma, k, sgn, codeword);
    else awgn_kernel<double, false><<<grid, threads, 0, stream>>>(out, n, B, Bp, frame0, seed, sigma, k, sgn, codeword);
    return cudaGetLastError();
}

cudaError_t launch_count_packed(int V, const uint32_t* hardw, int64_t Wn, int32_t n, int64_t B, const uint8_t* codeword,
                                const int32_t* iters, int64_t* counters, int32_t* frame_bit_errors,
                                int32_t* frame_iters, const int32_t* map, const uint8_t* only_done, int32_t* frame_cnt,
                                cudaStream_t stream) {
    cudaError_t e = cudaMemsetAsync(frame_cnt, 0, (size_t)Wn * 32 * sizeof(int32_t), stream);
    if (e != cudaSuccess) return e;
    dim3 grid((unsigned)((Wn + 127) / 128), (unsigned)((n + kCountVarChunk - 1) / kCountVarChunk));
    count_partial_kernel<<<grid, 128, 0, stream>>>(V, hardw, Wn, n, codeword, frame_cnt);
    count_final_kernel<<<(unsigned)((B + 255) / 256), 256, 0, stream>>>(frame_cnt, B, iters, counters, frame_bit_errors,
                                                                          frame_iters, map, only_done);
    return cudaGetLastError();
}

cudaError_t launch_count_bits(const uint8_t* bits, int32_t n, int64_t B, const uint8_t* codeword, const int32_t* iters,
                              int64_t* counters, int32_t* frame_bit_errors, cudaStream_t stream) {
    const int warps = 8;
    count_bits_kernel<<<(unsigned)((B + warps - 1) / warps), warps * 32, 0, stream>>>(bits, n, B, codeword, iters,
                                                                                       counters, frame_bit_errors);
    return cudaGetLastError();
}

cudaError_t launch_pending_scan(const uint8_t* done, int64_t Bp, int32_t* counts, int32_t* total, cudaStream_t stream) {
    const int nb = (int)((Bp + kScanBlock - 1) / kScanBlock);
    pending_count_kernel<<<nb, kScanBlock, 0, stream>>>(done, Bp, counts);
    pending_scan_kernel<<<1, kScanBlock, 0, stream>>>(counts, nb, total);
    return cudaGetLastError();
}

cudaError_t launch_pending_indices(const uint8_t* done, int64_t Bp, const int32_t* offsets, int32_t* idx, cudaStream_t stream) {
    const int nb = (int)((Bp + kScanBlock - 1) / kScanBlock);
    pending_index_kernel<<<nb, kScanBlock, 0, stream>>>(done, Bp, offsets, idx);
    return cudaGetLastError();
}

cudaError_t launch_gather_cols(int dtype, const void* src, int64_t Bp_src, void* dst, int64_t Bp_dst, const int32_t* idx,
                               int64_t count, int64_t rows, cudaStream_t stream) {
    if (rows <= 0) return cudaSuccess;
    dim3 grid((unsigned)((Bp_dst + 255) / 256), (unsigned)(rows < 4096 ? rows : 4096));
    if (dtype == 0) gather_cols_kernel<float><<<grid, 256, 0, stream>>>(static_cast<const float*>(src), Bp_src, static_cast<float*>(dst), Bp_dst, idx, count, rows);
    else gather_cols_kernel<double><<<grid, 256, 0, stream>>>(static_cast<const double*>(src), Bp_src, static_cast<double*>(dst), Bp_dst, idx, count, rows);
    return cudaGetLastError();
}

cudaError_t launch_scatter_frames(const int32_t* iters_src, const uint8_t* succ_src, int32_t* iters_dst, uint8_t* succ_dst,
                                  const int32_t* map, int64_t count, cudaStream_t stream) {
    scatter_frames_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(iters_src, succ_src, iters_dst, succ_dst, map, count);
    return cudaGetLastError();
}

cudaError_t launch_compose_map(const int32_t* idx, const int32_t* parent_map, int32_t* out, int64_t count, cudaStream_t stream) {
    compose_map_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(idx, parent_map, out, count);
    return cudaGetLastError();
}

}  // namespace ldpc
