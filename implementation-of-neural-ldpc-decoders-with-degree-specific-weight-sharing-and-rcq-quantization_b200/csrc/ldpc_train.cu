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
// variable side runs at the HBM roofline (6.5 TB/s), the check side is issue-bound (74 % of issue slots, 3.1 TB/s).)
#include "ldpc_cn_common.cuh"

namespace ldpc {

namespace {

constexpr int kTrainThreads = 128;

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
__global__ void __launch_bounds__(kTrainThreads) train_bwd_cn_kernel(const TrainBwd p, const int t, const int nfb) {
    constexpr int V = 4;
    const int fb = blockIdx.x % nfb;
    const int item_id = blockIdx.x / nfb;
    const int64_t f0 = ((int64_t)fb * kTrainThreads + threadIdx.x) * V;
    if (f0 >= p.Bp) return;
    bool on[V];   // frames that executed iteration t (the message slices of the others may hold anything)
#pragma unroll
    for (int v = 0; v < V; ++v) on[v] = (f0 + v < p.B) && t <= p.iters[f0 + v] - 1;
    const WorkItem it = p.cn_items[item_id];
    const int dc = it.deg;
    const float* beta_t = p.beta ? p.beta + (size_t)t * p.n_beta : nullptr;
    const float* src = (t == 0) ? p.llrT : p.v2c_t;
    const int lane = threadIdx.x & 31;
    float* const g_beta = p.g_beta ? p.g_beta + ((size_t)(blockIdx.x & (unsigned)(p.beta_parts - 1)) * p.T + t) * p.n_beta : nullptr;
    float acc = 0.f;      // one-beta-per-check weights: the checks of an item mostly share their column
    int acc_col = -1;
    auto flush_beta = [&]() {
        if (acc_col < 0) return;
        const float s = warp_sum(acc);
        if (lane == 0 && s != 0.f) atomicAdd(g_beta + acc_col, s);
        acc = 0.f;
    };
    for (int c = 0; c < it.count; ++c) {
        const int64_t slot0 = (int64_t)it.first_slot + (int64_t)c * dc;
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
            if (p.g_beta && beta_t) {
                if (p.beta_per_edge) {
                    gb = warp_sum(gb);
                    if (lane == 0 && gb != 0.f) atomicAdd(g_beta + bcol, gb);
                } else {
                    gb_check += gb;
                }
            }
        }
        if (p.g_beta && beta_t && !p.beta_per_edge) {
            const int bcol = p.bidx ? p.bidx[slot0] : 0;
            if (bcol != acc_col) {
                flush_beta();
                acc_col = bcol;
            }
            acc += gb_check;
        }
        if (t == 0) continue;   // g_v2c_0 would be the gradient with respect to the LLRs: not needed
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
    }
    flush_beta();
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
    train_bwd_cn_kernel<<<(unsigned)grid, kTrainThreads, 0, stream>>>(p, t, (int)nfb);
    return cudaGetLastError();
}

}  // namespace ldpc
