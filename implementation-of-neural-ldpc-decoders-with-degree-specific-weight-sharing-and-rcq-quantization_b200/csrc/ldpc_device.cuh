// Device-side building blocks shared by the kernels: packed frame vectors, streaming loads/stores,
// round-to-nearest arithmetic without FMA contraction, and the library reduction orders that the
// reference's hard decisions depend on.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldpc {

// ---------------------------------------------------------------------------------------------
// V consecutive frames of one message row, held by one lane.  A warp therefore touches 32*V
// consecutive frames = 512 contiguous bytes (float x4, double x2) per message row.
// ---------------------------------------------------------------------------------------------
template <typename T, int V>
struct alignas(sizeof(T) * V) Pack {
    T v[V];
};

template <typename T> struct FramesPerLane;
template <> struct FramesPerLane<float> { static constexpr int value = 4; };
template <> struct FramesPerLane<double> { static constexpr int value = 2; };

// Streaming (evict-first) accesses: every message byte is touched once per half-iteration.
template <typename P>
__device__ __forceinline__ P ld_stream(const void* ptr) {
    static_assert(sizeof(P) == 16 || sizeof(P) == 8 || sizeof(P) == 4 || sizeof(P) == 2, "pack size");
    P out;
    if constexpr (sizeof(P) == 16) {
        float4 t = __ldcs(reinterpret_cast<const float4*>(ptr));
        out = *reinterpret_cast<P*>(&t);
    } else if constexpr (sizeof(P) == 8) {
        float2 t = __ldcs(reinterpret_cast<const float2*>(ptr));
        out = *reinterpret_cast<P*>(&t);
    } else if constexpr (sizeof(P) == 4) {
        unsigned t = __ldcs(reinterpret_cast<const unsigned*>(ptr));
        out = *reinterpret_cast<P*>(&t);
    } else {
        unsigned short t = __ldcs(reinterpret_cast<const unsigned short*>(ptr));
        out = *reinterpret_cast<P*>(&t);
    }
    return out;
}

template <typename P>
__device__ __forceinline__ void st_stream(void* ptr, const P& val) {
    if constexpr (sizeof(P) == 16) {
        __stcs(reinterpret_cast<float4*>(ptr), *reinterpret_cast<const float4*>(&val));
    } else if constexpr (sizeof(P) == 8) {
        __stcs(reinterpret_cast<float2*>(ptr), *reinterpret_cast<const float2*>(&val));
    } else if constexpr (sizeof(P) == 4) {
        __stcs(reinterpret_cast<unsigned*>(ptr), *reinterpret_cast<const unsigned*>(&val));
    } else {
        __stcs(reinterpret_cast<unsigned short*>(ptr), *reinterpret_cast<const unsigned short*>(&val));
    }
}

// ---------------------------------------------------------------------------------------------
// IEEE round-to-nearest ops that the compiler may not fuse (llr + alpha*s must round the product).
// ---------------------------------------------------------------------------------------------
template <typename T> struct Arith;
template <> struct Arith<float> {
    static __device__ __forceinline__ float add(float a, float b) { return __fadd_rn(a, b); }
    static __device__ __forceinline__ float mul(float a, float b) { return __fmul_rn(a, b); }
    static __device__ __forceinline__ float abs(float a) { return fabsf(a); }
    static __device__ __forceinline__ float inf() { return __int_as_float(0x7f800000); }
    static __device__ __forceinline__ float fmin_(float a, float b) { return fminf(a, b); }
    static __device__ __forceinline__ float fmax_(float a, float b) { return fmaxf(a, b); }
    // top 32 bits of the representation (sign bit = bit 31)
    static __device__ __forceinline__ uint32_t hi(float a) { return __float_as_uint(a); }
    static __device__ __forceinline__ float flip(float a, uint32_t signmask) {
        return __uint_as_float(__float_as_uint(a) ^ (signmask & 0x80000000u));
    }
};
template <> struct Arith<double> {
    static __device__ __forceinline__ double add(double a, double b) { return __dadd_rn(a, b); }
    static __device__ __forceinline__ double mul(double a, double b) { return __dmul_rn(a, b); }
    static __device__ __forceinline__ double abs(double a) { return fabs(a); }
    static __device__ __forceinline__ double inf() { return __longlong_as_double(0x7ff0000000000000LL); }
    static __device__ __forceinline__ double fmin_(double a, double b) { return fmin(a, b); }
    static __device__ __forceinline__ double fmax_(double a, double b) { return fmax(a, b); }
    static __device__ __forceinline__ uint32_t hi(double a) { return (uint32_t)__double2hiint(a); }
    static __device__ __forceinline__ double flip(double a, uint32_t signmask) {
        int h = __double2hiint(a) ^ (int)(signmask & 0x80000000u);
        return __hiloint2double(h, __double2loint(a));
    }
};

// ---------------------------------------------------------------------------------------------
// Reduction orders.  `get(i)` returns element i of the k-vector being summed.
//
// An accumulator that is still empty takes the first value unchanged: the reference starts its
// accumulators at +0 and 0 + x == x for every x except that it turns -0 into +0, and signed zeros
// never change a later non-zero value, a comparison `< 0`, or a quantiser index.
// ---------------------------------------------------------------------------------------------
template <typename T>
struct OptAcc {
    T v;
    bool has;
    __device__ __forceinline__ OptAcc() : v(T(0)), has(false) {}
    __device__ __forceinline__ void add(T x) {
        if (has) {
            v = Arith<T>::add(v, x);
        } else {
            v = x;
            has = true;
        }
    }
    __device__ __forceinline__ void add(const OptAcc& o) {
        if (o.has) add(o.v);
    }
};

// torch.sum over a contiguous float32 k-vector.  K is a compile-time count so that everything
// unrolls into registers.
//   k <= 7: acc[q] += x[4i+q] over full groups of four, leftovers into acc[0], ((a0+a1)+a2)+a3
//   k >= 8: k/8 eight-lane vectors combined lane-wise by the same scheme; r = tail scalars in
//           order, then r += lane_0 .. lane_7
template <int K, typename Get>
__device__ __forceinline__ float torch_sum_static(Get get) {
    if constexpr (K == 0) {
        return 0.f;
    } else if constexpr (K < 8) {
        OptAcc<float> a[4];
        constexpr int g = K / 4;
#pragma unroll
        for (int i = 0; i < g; ++i) {
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q].add(get(4 * i + q));
        }
#pragma unroll
        for (int r = 4 * g; r < K; ++r) a[0].add(get(r));
        a[0].add(a[1]);
        a[0].add(a[2]);
        a[0].add(a[3]);
        return a[0].v;
    } else {
        constexpr int nv = K / 8;
        OptAcc<float> r;
#pragma unroll
        for (int t = 8 * nv; t < K; ++t) r.add(get(t));
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            OptAcc<float> a[4];
            constexpr int g = nv / 4;
#pragma unroll
            for (int i = 0; i < g; ++i) {
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q].add(get(8 * (4 * i + q) + l));
            }
#pragma unroll
            for (int v = 4 * g; v < nv; ++v) a[0].add(get(8 * v + l));
            a[0].add(a[1]);
            a[0].add(a[2]);
            a[0].add(a[3]);
            r.add(a[0]);
        }
        return r.v;
    }
}

template <typename Get>
__device__ float torch_sum_dynamic(Get get, int k) {
    if (k == 0) return 0.f;
    if (k < 8) {
        OptAcc<float> a[4];
        int g = k / 4;
        for (int i = 0; i < g; ++i)
            for (int q = 0; q < 4; ++q) a[q].add(get(4 * i + q));
        for (int r = 4 * g; r < k; ++r) a[0].add(get(r));
        a[0].add(a[1]);
        a[0].add(a[2]);
        a[0].add(a[3]);
        return a[0].v;
    }
    int nv = k / 8;
    OptAcc<float> r;
    for (int t = 8 * nv; t < k; ++t) r.add(get(t));
    for (int l = 0; l < 8; ++l) {
        OptAcc<float> a[4];
        int g = nv / 4;
        for (int i = 0; i < g; ++i)
            for (int q = 0; q < 4; ++q) a[q].add(get(8 * (4 * i + q) + l));
        for (int v = 4 * g; v < nv; ++v) a[0].add(get(8 * v + l));
        a[0].add(a[1]);
        a[0].add(a[2]);
        a[0].add(a[3]);
        r.add(a[0]);
    }
    return r.v;
}

// np.sum over a contiguous float64 k-vector (numpy pairwise summation, k <= 128):
//   k < 8 : left to right;  else r[q] = x[q], r[q] += x[8i+q], ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)),
//   then the k%8 tail in order.
template <int K, typename Get>
__device__ __forceinline__ double np_sum_static(Get get) {
    if constexpr (K == 0) {
        return 0.0;
    } else if constexpr (K < 8) {
        double r = get(0);
#pragma unroll
        for (int i = 1; i < K; ++i) r = __dadd_rn(r, get(i));
        return r;
    } else {
        double r[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q] = get(q);
        constexpr int full = K - (K % 8);
#pragma unroll
        for (int i = 8; i < full; i += 8) {
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q] = __dadd_rn(r[q], get(i + q));
        }
        double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                               __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
#pragma unroll
        for (int i = full; i < K; ++i) res = __dadd_rn(res, get(i));
        return res;
    }
}

template <typename Get>
__device__ double np_sum_dynamic(Get get, int k) {
    if (k == 0) return 0.0;
    if (k < 8) {
        double r = get(0);
        for (int i = 1; i < k; ++i) r = __dadd_rn(r, get(i));
        return r;
    }
    // Blocks beyond 128 terms recurse in numpy; variable degrees that large are rejected on the host.
    double r[8];
    for (int q = 0; q < 8; ++q) r[q] = get(q);
    int full = k - (k % 8);
    for (int i = 8; i < full; i += 8)
        for (int q = 0; q < 8; ++q) r[q] = __dadd_rn(r[q], get(i + q));
    double res = __dadd_rn(__dadd_rn(__dadd_rn(r[0], r[1]), __dadd_rn(r[2], r[3])),
                           __dadd_rn(__dadd_rn(r[4], r[5]), __dadd_rn(r[6], r[7])));
    for (int i = full; i < k; ++i) res = __dadd_rn(res, get(i));
    return res;
}

// ---------------------------------------------------------------------------------------------
// The same two reduction orders on whole frame packs (all V frames of a lane at once), with a run-time
// length: used where a node's inputs are staged in shared memory (variable degrees above 8).
// ---------------------------------------------------------------------------------------------
template <typename T, int V>
struct PackAcc {
    Pack<T, V> v;
    bool has;
    __device__ __forceinline__ PackAcc() : has(false) {
#pragma unroll
        for (int i = 0; i < V; ++i) v.v[i] = T(0);
    }
    __device__ __forceinline__ void add(const Pack<T, V>& x) {
        if (has) {
#pragma unroll
            for (int i = 0; i < V; ++i) v.v[i] = Arith<T>::add(v.v[i], x.v[i]);
        } else {
            v = x;
            has = true;
        }
    }
    __device__ __forceinline__ void add(const PackAcc& o) {
        if (o.has) add(o.v);
    }
};

template <int V, typename Get>
__device__ __forceinline__ Pack<float, V> torch_sum_dynamic_pack(Get get, int k) {
    if (k < 8) {
        PackAcc<float, V> a[4];
        const int g = k / 4;
        for (int i = 0; i < g; ++i) {
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q].add(get(4 * i + q));
        }
        for (int r = 4 * g; r < k; ++r) a[0].add(get(r));
        a[0].add(a[1]);
        a[0].add(a[2]);
        a[0].add(a[3]);
        return a[0].v;
    }
    const int nv = k / 8;
    PackAcc<float, V> r;
    for (int t = 8 * nv; t < k; ++t) r.add(get(t));
    const int g = nv / 4;
    for (int l = 0; l < 8; ++l) {
        PackAcc<float, V> a[4];
        for (int i = 0; i < g; ++i) {
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q].add(get(8 * (4 * i + q) + l));
        }
        for (int v = 4 * g; v < nv; ++v) a[0].add(get(8 * v + l));
        a[0].add(a[1]);
        a[0].add(a[2]);
        a[0].add(a[3]);
        r.add(a[0]);
    }
    return r.v;
}

template <int V, typename Get>
__device__ __forceinline__ Pack<double, V> np_sum_dynamic_pack(Get get, int k) {
    PackAcc<double, V> res;
    if (k < 8) {
        for (int i = 0; i < k; ++i) res.add(get(i));
        return res.v;
    }
    PackAcc<double, V> r[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) r[q].add(get(q));
    const int full = k - (k % 8);
    for (int i = 8; i < full; i += 8) {
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q].add(get(i + q));
    }
    r[0].add(r[1]);
    r[2].add(r[3]);
    r[4].add(r[5]);
    r[6].add(r[7]);
    r[0].add(r[2]);
    r[4].add(r[6]);
    r[0].add(r[4]);
    for (int i = full; i < k; ++i) r[0].add(get(i));
    return r[0].v;
}

// Compile-time lengths: everything unrolls, the `has` flags fold away and `get(i)` sees constant indices.
template <int K, int V, typename Get>
__device__ __forceinline__ Pack<float, V> torch_sum_static_pack(Get get) {
    if constexpr (K < 8) {
        PackAcc<float, V> a[4];
        constexpr int g = K / 4;
#pragma unroll
        for (int i = 0; i < g; ++i) {
#pragma unroll
            for (int q = 0; q < 4; ++q) a[q].add(get(4 * i + q));
        }
#pragma unroll
        for (int r = 4 * g; r < K; ++r) a[0].add(get(r));
        a[0].add(a[1]);
        a[0].add(a[2]);
        a[0].add(a[3]);
        return a[0].v;
    } else {
        constexpr int nv = K / 8;
        constexpr int g = nv / 4;
        PackAcc<float, V> r;
#pragma unroll
        for (int t = 8 * nv; t < K; ++t) r.add(get(t));
#pragma unroll
        for (int l = 0; l < 8; ++l) {
            PackAcc<float, V> a[4];
#pragma unroll
            for (int i = 0; i < g; ++i) {
#pragma unroll
                for (int q = 0; q < 4; ++q) a[q].add(get(8 * (4 * i + q) + l));
            }
#pragma unroll
            for (int v = 4 * g; v < nv; ++v) a[0].add(get(8 * v + l));
            a[0].add(a[1]);
            a[0].add(a[2]);
            a[0].add(a[3]);
            r.add(a[0]);
        }
        return r.v;
    }
}

template <int K, int V, typename Get>
__device__ __forceinline__ Pack<double, V> np_sum_static_pack(Get get) {
    if constexpr (K < 8) {
        PackAcc<double, V> res;
#pragma unroll
        for (int i = 0; i < K; ++i) res.add(get(i));
        return res.v;
    } else {
        PackAcc<double, V> r[8];
#pragma unroll
        for (int q = 0; q < 8; ++q) r[q].add(get(q));
        constexpr int full = K - (K % 8);
#pragma unroll
        for (int i = 8; i < full; i += 8) {
#pragma unroll
            for (int q = 0; q < 8; ++q) r[q].add(get(i + q));
        }
        r[0].add(r[1]);
        r[2].add(r[3]);
        r[4].add(r[5]);
        r[6].add(r[7]);
        r[0].add(r[2]);
        r[4].add(r[6]);
        r[0].add(r[4]);
#pragma unroll
        for (int i = full; i < K; ++i) r[0].add(get(i));
        return r[0].v;
    }
}

template <typename T> struct LibSum;
template <> struct LibSum<float> {
    template <int K, typename Get> static __device__ __forceinline__ float stat(Get g) { return torch_sum_static<K>(g); }
    template <typename Get> static __device__ __forceinline__ float dyn(Get g, int k) { return torch_sum_dynamic(g, k); }
    template <int V, typename Get> static __device__ __forceinline__ Pack<float, V> dyn_pack(Get g, int k) { return torch_sum_dynamic_pack<V>(g, k); }
    template <int K, int V, typename Get> static __device__ __forceinline__ Pack<float, V> stat_pack(Get g) { return torch_sum_static_pack<K, V>(g); }
};
template <> struct LibSum<double> {
    template <int K, typename Get> static __device__ __forceinline__ double stat(Get g) { return np_sum_static<K>(g); }
    template <typename Get> static __device__ __forceinline__ double dyn(Get g, int k) { return np_sum_dynamic(g, k); }
    template <int V, typename Get> static __device__ __forceinline__ Pack<double, V> dyn_pack(Get g, int k) { return np_sum_dynamic_pack<V>(g, k); }
    template <int K, int V, typename Get> static __device__ __forceinline__ Pack<double, V> stat_pack(Get g) { return np_sum_static_pack<K, V>(g); }
};

// ---------------------------------------------------------------------------------------------
// Quantiser index (rcq_decoder.py:76-84): last j with mag >= thr[j]; j = 0 never changes the
// initial 0.  `mono` says the thresholds are non-decreasing, so "last" == "count" and a binary
// search is valid for wide quantisers.
// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ int quant_index(float mag, const float* __restrict__ thr, int nth, bool mono) {
    int idx = 0;
    if (nth <= 8 || !mono) {
        for (int j = 1; j < nth; ++j)
            if (mag >= thr[j]) idx = j;
    } else {
        int lo = 0, hi = nth;  // invariant: thr[lo] passes (or lo == 0), thr[hi] fails (or hi == nth)
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (mag >= thr[mid]) lo = mid; else hi = mid;
        }
        idx = lo;
    }
    return idx;
}

}  // namespace ldpc
