// Backward pass of posterior training (training_framework.py:37-295 with the two repairs recorded in
// oracle/reference_training_repairs.patch): gradients of a loss on the final posteriors with respect to the
// degree-shared / per-edge weights beta[t][.] and alpha[t][.] of the normalised neural min-sum decoders
// (neural_2d_decoder.py:133-225, neural_minsum_decoder.py:58-150), batched over frames.
//
// The forward pass is the inference schedule itself (ldpc_cn.cu / ldpc_vn.cu) run on per-iteration message slices, so
// the history the backward pass needs -- v2c_t (inputs of the check nodes of iteration t) and c2v_t (their outputs) --
// is simply what those kernels wrote.  What autograd differentiates in the reference, per frame whose decode stopped
// after iteration t* (0-based):
//     post_j      = llr_j + sum_e c2v_{t*}[e]                          -> g c2v_{t*}[e] = g post_{var(e)}
//     c2v_t[e]    = beta_t[e] * raw_e * sp_e,   raw_e = m2 if e == k0 else m1      (sp_e: product of the other signs)
//                   g beta_t[col(e)] += g c2v_t[e] * raw_e * sp_e
//                   g |x_{k0}| += sum_{e != k0} g c2v_t[e] * beta_t[e] * sp_e       (m1 = |x_{k0}|, first minimum)
//                   g |x_k|    += g c2v_t[k0] * beta_t[k0] * sp_{k0} / cnt          for the cnt edges k != k0 with |x_k| == m2
//                   g x = sign(x) * g |x|   (torch.abs; torch.sign and the argmin carry no gradient)
//     v2c_{t+1}[e] = llr_j + alpha_t[j] * sum_{e' != e} c2v_t[e']
//                   g alpha_t[col(j)] += g v2c_{t+1}[e] * sum_{e' != e} c2v_t[e']
//                   g c2v_t[e']       += alpha_t[j] * (G - g v2c_{t+1}[e']),  G = sum_e g v2c_{t+1}[e]
// Lanes over frames like every other kernel here; weight gradients are reduced over the warp, accumulated over the nodes
// of a work item that share a column, and added atomically to one of `parts` spread copies (folded at the end).
// (Degree-templated bodies were tried twice -- rows of a node held in registers: 96-128 registers, 22 -> 33 ms; rows re-read
// from L1 in every pass, unrolled, with and without a register cap: 30-35 ms.  ncu on the run-time-degree kernels below: the
// variable side runs at the HBM roofline (6.5 TB/s), the check side at 5.9 TB/s since its three-pass rewrite.)
#include "ldpc_cn_common.cuh"

namespace ldpc {

namespace {

constexpr int kTrainThreads = 128;
#ifndef LDPC_TRAIN_UNROLL
#define LDPC_TRAIN_UNROLL 4   // rows of a check in flight per pass of the check-side backward kernel
#endif
constexpr int kTrainUnroll = LDPC_TRAIN_UNROLL;
#ifndef LDPC_TRAIN_CN_MINCTAS
#define LDPC_TRAIN_CN_MINCTAS 6
#endif
#ifndef LDPC_TRAIN_PREFETCH
#define LDPC_TRAIN_PREFETCH 1   // the next check's rows are asked into L2 while the current one is worked on
#endif

__device__ __forceinline__ void prefetch_l2(const void* ptr) { asm volatile("prefetch.global.L2 [%0];" ::"l"(ptr)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Variable side of iteration t: g_c2v_t from the posterior gradient (frames that stopped at t) or from g_v2c_{t+1}
// (frames that ran on), zero for frames that had stopped earlier; alpha gradients.
__global__ void __launch_bounds__(kTrainThreads) train_bwd_vn_kernel(const TrainBwd p, const int t, const int nfb) {
    constexpr int V = 4;
    const int fb = blockIdx.x % nfb;
    const int item_id = blockIdx.x / nfb;
    const int64_t f0 = ((int64_t)fb * kTrainThreads + threadIdx.x) * V;
    if (f0 >= p.Bp) return;   // whole warps
    int tstar[V];
#pragma unroll
    for (int v = 0; v < V; ++v) tstar[v] = (f0 + v < p.B) ? p.iters[f0 + v] - 1 : -1;
    const WorkItem it = p.vn_items[item_id];
    const int dv = it.deg;
    const float* alpha_t = p.alpha ? p.alpha + (size_t)t * p.n_alpha : nullptr;
    const int lane = threadIdx.x & 31;
    float* const g_alpha = p.g_alpha ? p.g_alpha + ((size_t)(blockIdx.x & (unsigned)(p.alpha_parts - 1)) * p.T + t) * p.n_alpha : nullptr;
    float acc = 0.f;
    int acc_col = -1;
    auto flush_alpha = [&]() {
        if (acc_col < 0) return;
        const float s = warp_sum(acc);
        if (lane == 0 && s != 0.f) atomicAdd(g_alpha + acc_col, s);
        acc = 0.f;
    };
    for (int c = 0; c < it.count; ++c) {
        const int pos = it.first_node + c;
        const int64_t lbase = (int64_t)it.first_slot + (int64_t)c * dv;
        const int64_t j = p.vpos_var[pos];
        const Pack<float, V> gp = *reinterpret_cast<const Pack<float, V>*>(p.g_post + j * p.Bp + f0);
        float alpha = 1.f;
        const int acol = p.aidx ? p.aidx[pos] : 0;
        if (alpha_t) alpha = alpha_t[acol];
        // sums over the variable's edges: G = sum g_v2c_{t+1}, C = sum c2v_t, GC = sum g_v2c_{t+1} * c2v_t
        float G[V] = {0.f, 0.f, 0.f, 0.f}, C[V] = {0.f, 0.f, 0.f, 0.f}, GC[V] = {0.f, 0.f, 0.f, 0.f};
        for (int i = 0; i < dv; ++i) {
            const int64_t s = p.vslots[lbase + i];
            const Pack<float, V> g = *reinterpret_cast<const Pack<float, V>*>(p.g_v2c + s * p.Bp + f0);
            const Pack<float, V> cv = *reinterpret_cast<const Pack<float, V>*>(p.c2v_t + s * p.Bp + f0);
#pragma unroll
            for (int v = 0; v < V; ++v) {
                const bool on = t < tstar[v];   // g_v2c_{t+1} exists only for frames that ran iteration t+1; the
                const float gv = on ? g.v[v] : 0.f;   // slices of a frame that stopped earlier may hold anything
                const float cc = on ? cv.v[v] : 0.f;
                G[v] += gv;
                C[v] += cc;
                GC[v] += gv * cc;
            }
        }
        for (int i = 0; i < dv; ++i) {
            const int64_t s = p.vslots[lbase + i];
            const Pack<float, V> g = *reinterpret_cast<const Pack<float, V>*>(p.g_v2c + s * p.Bp + f0);
            Pack<float, V> out;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                if (t == tstar[v]) out.v[v] = gp.v[v];
                else if (t < tstar[v]) out.v[v] = alpha * (G[v] - g.v[v]);
                else out.v[v] = 0.f;
            }
            *reinterpret_cast<Pack<float, V>*>(p.g_c2v + s * p.Bp + f0) = out;
        }
        if (p.g_alpha && alpha_t) {
            // sum_e g_e * (C - c_e) = G * C - GC
            float ga = 0.f;
#pragma unroll
            for (int v = 0; v < V; ++v) ga += (t < tstar[v]) ? (G[v] * C[v] - GC[v]) : 0.f;
            // consecutive variables of an item mostly share their column (one degree class): one atomic per run
            if (acol != acc_col) {
                flush_alpha();
                acc_col = acol;
            }
            acc += ga;
        }
    }
    flush_alpha();
}

// Check side of iteration t: g_v2c_t and the beta gradients from g_c2v_t.
//
// General form of one check (any degree, zero-valued inputs with torch.sign's three values, ties): four passes over
// the check's rows.  Taken for degree-1 checks and for the rare check that sees an exact zero in one of the warp's
// running frames; returns the frames' summed beta gradient of a one-beta-per-check table (per-edge tables are added
// here).  Warp-uniform: every lane of the warp is inside.
__device__ __noinline__ float cn_bwd_check_general(const TrainBwd& p, const int t, const float* __restrict__ src, const int64_t slot0,
                                                   const int dc, const int64_t f0, const uint32_t on_mask, const float* beta_t,
                                                   float* g_beta, const int lane) {
    constexpr int V = 4;
    bool on[V];
#pragma unroll
    for (int v = 0; v < V; ++v) on[v] = (on_mask >> v) & 1u;
    auto row = [&](int k) -> int64_t { return (t == 0) ? (int64_t)p.slot_var[slot0 + k] : slot0 + k; };
    // forward statistics of the check, per frame
    float m1[V], m2[V];
    int k0[V], cnt2[V], zeros[V];
    uint32_t par[V];
#pragma unroll
    for (int v = 0; v < V; ++v) {
        m1[v] = m2[v] = __int_as_float(0x7f800000);
        k0[v] = 0;
        cnt2[v] = 0;
        zeros[v] = 0;
        par[v] = 0;
    }
    for (int k = 0; k < dc; ++k) {
        const Pack<float, V> x = *reinterpret_cast<const Pack<float, V>*>(src + row(k) * p.Bp + f0);
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const float a = fabsf(x.v[v]);
            if (a < m1[v]) {          // strict: the first minimum keeps the argmin (torch.argmin)
                m2[v] = m1[v];
                m1[v] = a;
                k0[v] = k;
            } else if (a < m2[v]) {
                m2[v] = a;
            }
            par[v] ^= __float_as_uint(x.v[v]);
            zeros[v] += (x.v[v] == 0.f);
        }
    }
    for (int k = 0; k < dc; ++k) {   // ties of the second minimum share its gradient (torch.min backward)
        const Pack<float, V> x = *reinterpret_cast<const Pack<float, V>*>(src + row(k) * p.Bp + f0);
#pragma unroll
        for (int v = 0; v < V; ++v) cnt2[v] += (k != k0[v] && fabsf(x.v[v]) == m2[v]);
    }
    // pass 1: beta gradients and the two sums that flow to the minima
    float to_min[V] = {0.f, 0.f, 0.f, 0.f}, to_min2[V] = {0.f, 0.f, 0.f, 0.f};
    float gb_check = 0.f;
    for (int k = 0; k < dc; ++k) {
        const Pack<float, V> x = *reinterpret_cast<const Pack<float, V>*>(src + row(k) * p.Bp + f0);
        const Pack<float, V> g = *reinterpret_cast<const Pack<float, V>*>(p.g_c2v + (slot0 + k) * p.Bp + f0);
        const int bcol = p.bidx ? p.bidx[slot0 + k] : 0;
        const float beta = beta_t ? beta_t[bcol] : p.beta_const;
        float gb = 0.f;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            const bool is_min = (k == k0[v]);
            // product of the OTHER signs: 0 if one of them is zero (three-valued torch.sign)
            const int other_zeros = zeros[v] - (x.v[v] == 0.f);
            float sp = ((par[v] ^ __float_as_uint(x.v[v])) >> 31) ? -1.f : 1.f;
            if (other_zeros > 0) sp = 0.f;
            if (dc == 1) sp = 1.f;
            const float raw = is_min ? (dc == 1 ? m1[v] : m2[v]) : m1[v];
            const float gv = on[v] ? g.v[v] : 0.f;
            gb += on[v] ? gv * raw * sp : 0.f;
            const float graw = gv * beta * sp;
            if (is_min) to_min2[v] += graw;     // raw = m2 (or m1 itself for a degree-1 check)
            else to_min[v] += graw;             // raw = m1
        }
        if (g_beta && beta_t) {
            if (p.beta_per_edge) {
                gb = warp_sum(gb);
                if (lane == 0 && gb != 0.f) atomicAdd(g_beta + bcol, gb);
            } else {
                gb_check += gb;
            }
        }
    }
    if (t == 0) return gb_check;   // g_v2c_0 would be the gradient with respect to the LLRs: not needed
    // pass 2: g_v2c_t
    for (int k = 0; k < dc; ++k) {
        const Pack<float, V> x = *reinterpret_cast<const Pack<float, V>*>(src + row(k) * p.Bp + f0);
        Pack<float, V> out;
#pragma unroll
        for (int v = 0; v < V; ++v) {
            float gabs = 0.f;
            if (k == k0[v]) gabs = to_min[v] + (dc == 1 ? to_min2[v] : 0.f);
            else if (fabsf(x.v[v]) == m2[v]) gabs = to_min2[v] / (float)cnt2[v];
            const float sg = x.v[v] > 0.f ? 1.f : (x.v[v] < 0.f ? -1.f : 0.f);
            out.v[v] = on[v] ? sg * gabs : 0.f;
        }
        *reinterpret_cast<Pack<float, V>*>(p.g_v2c + (slot0 + k) * p.Bp + f0) = out;
    }
    return gb_check;
}

// The kernel proper runs the common case -- degree >= 2, no exact zero among the inputs of the warp's running frames --
// in three short passes over the check's rows (the first from HBM, the others from L1):
//   statistics   m1, m2, first argmin k0 and the xor of the sign bits; an input is zero iff m1 == 0, which sends the
//                warp to the general form above, so below every sign is +-1 and is applied by flipping a sign bit;
//   gradients    g*sp per edge (two LOP3), the sums flowing to the two minima, the count of ties at m2; the beta
//                gradient of a one-beta-per-check table needs nothing per edge: it is m1 * to_min + m2 * to_min2;
//   outputs      g_v2c_t: to_min at k0, to_min2 / ties at the ties of m2, zero elsewhere, times sign(x).
// Selections are written as selp (the compiler turned the nested conditionals into divergent branches: 12 % of the
// issued instructions were BSSY / BSYNC / BRA); rows are walked by pointer increments (a quarter of the instructions
// of the first version were 64-bit row-address multiplies); while a check is worked on, the rows of the next one are
// asked into L2.  Frames that did not execute iteration t are computed on whatever their slices hold and masked per
// check.  (ncu: four-pass form 1.55 ms per launch at 8192 frames, 74 % of issue slots; this form: see profiles/.)
__device__ __forceinline__ float selp(const float a, const float b, const bool p) {   // p ? a : b, never a branch
    float d;
    asm("{ .reg .pred q; setp.ne.u32 q, %3, 0; selp.f32 %0, %1, %2, q; }" : "=f"(d) : "f"(a), "f"(b), "r"((uint32_t)p));
    return d;
}
__device__ __forceinline__ int selp(const int a, const int b, const bool p) {
    int d;
    asm("{ .reg .pred q; setp.ne.u32 q, %3, 0; selp.s32 %0, %1, %2, q; }" : "=r"(d) : "r"(a), "r"(b), "r"((uint32_t)p));
    return d;
}

template <bool PER_EDGE, bool T0>
__global__ void __launch_bounds__(kTrainThreads, LDPC_TRAIN_CN_MINCTAS) train_bwd_cn_kernel(const __grid_constant__ TrainBwd p, const int t, const int nfb) {
    constexpr int V = 4;
    using Row = Pack<float, V>;
    const int fb = blockIdx.x % nfb;
    const int item_id = blockIdx.x / nfb;
    const int64_t f0 = ((int64_t)fb * kTrainThreads + threadIdx.x) * V;
    if (f0 >= p.Bp) return;
    bool on[V];   // frames that executed iteration t (the message slices of the others may hold anything)
#pragma unroll
    for (int v = 0; v < V; ++v) on[v] = (f0 + v < p.B) && t <= p.iters[f0 + v] - 1;
    const uint32_t on_mask = (uint32_t)on[0] | ((uint32_t)on[1] << 1) | ((uint32_t)on[2] << 2) | ((uint32_t)on[3] << 3);
    const WorkItem it = p.cn_items[item_id];
    const int dc = it.deg;
    const float* beta_t = p.beta ? p.beta + (size_t)t * p.n_beta : nullptr;
    const float* src = T0 ? p.llrT : p.v2c_t;
    const int lane = threadIdx.x & 31;
    float* const g_beta = p.g_beta ? p.g_beta + ((size_t)(blockIdx.x & (unsigned)(p.beta_parts - 1)) * p.T + t) * p.n_beta : nullptr;
    const bool want_gb = g_beta && beta_t;
    float acc = 0.f;      // one-beta-per-check weights: the checks of an item mostly share their column
    int acc_col = -1;
    auto flush_beta = [&]() {
        if (acc_col < 0) return;
        const float s = warp_sum(acc);
        if (lane == 0 && s != 0.f) atomicAdd(g_beta + acc_col, s);
        acc = 0.f;
    };
    auto add_check = [&](const int64_t slot0, const float gb_check) {
        if (PER_EDGE || !want_gb) return;
        const int bcol = p.bidx ? p.bidx[slot0] : 0;
        if (bcol != acc_col) {
            flush_beta();
            acc_col = bcol;
        }
        acc += gb_check;
    };
    const float inf = __int_as_float(0x7f800000);
    const uint32_t row_bytes = (uint32_t)p.Bp * 4u;
    const size_t row_step = (size_t)p.Bp;                       // floats between consecutive rows
    const float* const src_f = src + f0;                        // this lane's frames of row 0
    // rows of the item's checks are consecutive slots: one running offset serves x (t > 0), g_c2v and g_v2c
    size_t off0 = (size_t)it.first_slot * row_step + (size_t)f0;
    const size_t check_step = (size_t)dc * row_step;
    for (int c = 0; c < it.count; ++c, off0 += check_step) {
        const int64_t slot0 = (int64_t)it.first_slot + (int64_t)c * dc;
        if (dc < 2 || p.force_general) {
            add_check(slot0, cn_bwd_check_general(p, t, src, slot0, dc, f0, on_mask, beta_t, g_beta, lane));
            continue;
        }
        const int32_t* const vars = p.slot_var + slot0;         // T0: the check's inputs are LLR rows
        auto xrow = [&](const int k, const float* walk) -> const Row* {
            if (T0) return reinterpret_cast<const Row*>(row_at(src_f, (uint32_t)vars[k], row_bytes));
            return reinterpret_cast<const Row*>(walk);
        };
#if LDPC_TRAIN_PREFETCH
        if (c + 1 < it.count) {
            const float* nx = p.v2c_t + off0 + check_step;
            const float* ng = p.g_c2v + off0 + check_step;
            for (int k = 0; k < dc; ++k, nx += row_step, ng += row_step) {
                prefetch_l2(T0 ? (const void*)row_at(src_f, (uint32_t)vars[dc + k], row_bytes) : (const void*)nx);
                prefetch_l2(ng);
            }
        }
#endif
        float m1[V], m2[V];
        int k0[V];
        uint32_t par[V];
#pragma unroll
        for (int v = 0; v < V; ++v) {
            m1[v] = m2[v] = inf;
            k0[v] = 0;
            par[v] = 0;
        }
        {
            const float* xp = p.v2c_t + off0;
#pragma unroll kTrainUnroll
            for (int k = 0; k < dc; ++k, xp += row_step) {
                const Row x = *xrow(k, xp);
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float a = fabsf(x.v[v]);
                    k0[v] = selp(k, k0[v], a < m1[v]);      // strict: the first minimum keeps the argmin (torch.argmin)
                    m2[v] = fminf(m2[v], fmaxf(a, m1[v]));
                    m1[v] = fminf(m1[v], a);
                    par[v] ^= __float_as_uint(x.v[v]);
                }
            }
        }
        bool zero_in = false;
#pragma unroll
        for (int v = 0; v < V; ++v) zero_in |= on[v] && !(m1[v] > 0.f);   // a zero (or a NaN) among a running frame's inputs
        if (__any_sync(0xffffffffu, zero_in)) {
            add_check(slot0, cn_bwd_check_general(p, t, src, slot0, dc, f0, on_mask, beta_t, g_beta, lane));
            continue;
        }
        float to_min[V] = {0.f, 0.f, 0.f, 0.f}, to_min2[V] = {0.f, 0.f, 0.f, 0.f};
        int cnt2[V] = {0, 0, 0, 0};
        float rm1[V], rm2[V];      // the two raw magnitudes, zero for frames that did not run (PER_EDGE: masks the beta sums)
#pragma unroll
        for (int v = 0; v < V; ++v) {
            par[v] &= 0x80000000u;
            rm1[v] = on[v] ? m1[v] : 0.f;
            rm2[v] = on[v] ? m2[v] : 0.f;
        }
        {
            const float* xp = p.v2c_t + off0;
            const float* gp = p.g_c2v + off0;
#pragma unroll kTrainUnroll
            for (int k = 0; k < dc; ++k, xp += row_step, gp += row_step) {
                const Row x = *xrow(k, xp);
                const Row g = ld_stream<Row>(gp);
                int bcol = 0;
                float beta = p.beta_const;
                if (PER_EDGE) {
                    bcol = p.bidx ? p.bidx[slot0 + k] : 0;
                    if (beta_t) beta = beta_t[bcol];
                }
                float gbe = 0.f;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const uint32_t xb = __float_as_uint(x.v[v]);
                    const bool is_min = (k == k0[v]);
                    // g * (product of the other signs): the sign bit of g flipped by parity ^ own sign
                    float gs = __uint_as_float(__float_as_uint(g.v[v]) ^ ((xb & 0x80000000u) ^ par[v]));
                    if (PER_EDGE) {
                        gbe = fmaf(gs, selp(rm2[v], rm1[v], is_min), gbe);
                        gs *= beta;
                    }                                        // (one beta per check: applied once after the loop)
                    to_min2[v] = selp(gs, to_min2[v], is_min);
                    to_min[v] += selp(0.f, gs, is_min);
                    cnt2[v] += (int)(!is_min && fabsf(x.v[v]) == m2[v]);
                }
                if (PER_EDGE && want_gb) {
                    gbe = warp_sum(gbe);
                    if (lane == 0 && gbe != 0.f) atomicAdd(g_beta + bcol, gbe);
                }
            }
        }
        if (!PER_EDGE && want_gb) {
            float gb_check = 0.f;     // sum over the check's edges of g * sp * raw
#pragma unroll
            for (int v = 0; v < V; ++v) gb_check += fmaf(rm1[v], to_min[v], rm2[v] * to_min2[v]);
            add_check(slot0, gb_check);
        }
        if (T0) continue;   // g_v2c_0 would be the gradient with respect to the LLRs: not needed
        float a_out[V], b_out[V];
        {
            float beta_check = 1.f;
            if (!PER_EDGE) {
                const int bcol = p.bidx ? p.bidx[slot0] : 0;
                beta_check = beta_t ? beta_t[bcol] : p.beta_const;
            }
            bool ties = false;
#pragma unroll
            for (int v = 0; v < V; ++v) {
                a_out[v] = on[v] ? (PER_EDGE ? to_min[v] : to_min[v] * beta_check) : 0.f;
                b_out[v] = on[v] ? (PER_EDGE ? to_min2[v] : to_min2[v] * beta_check) : 0.f;
                ties |= cnt2[v] > 1;
            }
            if (ties) {      // ties of the second minimum share its gradient (torch.min backward)
#pragma unroll
                for (int v = 0; v < V; ++v)
                    if (cnt2[v] > 1) b_out[v] = b_out[v] / (float)cnt2[v];
            }
        }
        {
            const float* xp = p.v2c_t + off0;
            float* op = p.g_v2c + off0;
#pragma unroll kTrainUnroll
            for (int k = 0; k < dc; ++k, xp += row_step, op += row_step) {
                const Row x = *reinterpret_cast<const Row*>(xp);
                Row out;
#pragma unroll
                for (int v = 0; v < V; ++v) {
                    const float gabs = selp(a_out[v], selp(b_out[v], 0.f, fabsf(x.v[v]) == m2[v]), k == k0[v]);
                    out.v[v] = __uint_as_float(__float_as_uint(gabs) ^ (__float_as_uint(x.v[v]) & 0x80000000u));
                }
                *reinterpret_cast<Row*>(op) = out;
            }
        }
    }
    if (!PER_EDGE && want_gb) flush_beta();
}

// out[i] = sum over the copies of parts[copy][i]
__global__ void train_fold_kernel(const float* __restrict__ parts, float* __restrict__ out, const int n_parts, const int64_t count) {
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    float s = 0.f;
    for (int k = 0; k < n_parts; ++k) s += parts[(size_t)k * count + i];
    out[i] = s;
}

}  // namespace

cudaError_t launch_train_fold(const float* parts, float* out, int n_parts, int64_t count, cudaStream_t stream) {
    if (count < 1) return cudaSuccess;
    train_fold_kernel<<<(unsigned)((count + 255) / 256), 256, 0, stream>>>(parts, out, n_parts, count);
    return cudaGetLastError();
}

cudaError_t launch_train_bwd_vn(const TrainBwd& p, int t, cudaStream_t stream) {
    if (p.n_vn_items == 0) return cudaSuccess;
    const int64_t nfb = (p.Bp / 4 + kTrainThreads - 1) / kTrainThreads;
    const int64_t grid = nfb * p.n_vn_items;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    train_bwd_vn_kernel<<<(unsigned)grid, kTrainThreads, 0, stream>>>(p, t, (int)nfb);
    return cudaGetLastError();
}

cudaError_t launch_train_bwd_cn(const TrainBwd& p, int t, cudaStream_t stream) {
    if (p.n_cn_items == 0) return cudaSuccess;
    const int64_t nfb = (p.Bp / 4 + kTrainThreads - 1) / kTrainThreads;
    const int64_t grid = nfb * p.n_cn_items;
    if (grid > 0x7fffffffLL) return cudaErrorInvalidConfiguration;
    const unsigned g = (unsigned)grid;
    if (p.beta_per_edge) {
        if (t == 0) train_bwd_cn_kernel<true, true><<<g, kTrainThreads, 0, stream>>>(p, t, (int)nfb);
        else train_bwd_cn_kernel<true, false><<<g, kTrainThreads, 0, stream>>>(p, t, (int)nfb);
    } else {
        if (t == 0) train_bwd_cn_kernel<false, true><<<g, kTrainThreads, 0, stream>>>(p, t, (int)nfb);
        else train_bwd_cn_kernel<false, false><<<g, kTrainThreads, 0, stream>>>(p, t, (int)nfb);
    }
    return cudaGetLastError();
}

}  // namespace ldpc
