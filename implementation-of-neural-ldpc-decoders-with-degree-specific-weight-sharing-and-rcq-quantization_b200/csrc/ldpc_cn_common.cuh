// Check-node arithmetic shared by the per-iteration kernels (ldpc_cn.cu) and the on-chip decode (ldpc_small.cu):
// running min1 / min2 / parity, the RCQ quantiser, the two-magnitudes-per-check output rule and the offset rule.
// Internal linkage: every translation unit gets its own copy.
#pragma once
#include "ldpc_kernel_common.cuh"

namespace ldpc {

namespace {

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

// Offset min-sum output (neural_minsum_decoder.py:236-253, neural_2d_decoder.py:383-401):
//   c2v = prod(other signs) * (relu(raw - beta) - alpha); a zero among the OTHER inputs forces 0 (see cn_offset_kernel)
template <typename Real>
__device__ __forceinline__ Real offset_value(Real raw, Real beta, bool has_beta, Real alpha, bool has_alpha,
                                             uint32_t signbits, bool zero_others) {
    Real t = has_beta ? Arith<Real>::add(raw, -beta) : raw;
    t = Arith<Real>::fmax_(t, Real(0));
    if (has_alpha) t = Arith<Real>::add(t, -alpha);
    t = Arith<Real>::flip(t, signbits);
    return zero_others ? Real(0) : t;
}

}  // namespace

}  // namespace ldpc
