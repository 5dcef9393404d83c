// CTA-resident decode: one thread block decodes one frame at a time with the frame's messages held in ITS shared
// memory for all T iterations -- the fused kernel of the design brief for codes whose per-frame state fits one SM
// (4 bytes per edge, in place: 194 KB for the (16200,7200)-shaped code, E = 48 599).  Threads of the block run over the
// frame's check nodes, then over its variable nodes ("lanes over nodes"), with a block barrier between the phases:
//     check node    reads its dc inputs from the message array and overwrites them with its dc outputs
//     variable node reads its dv inputs, forms posterior / hard decision / the dv leave-one-out sums, overwrites them
//     syndrome      XOR of the hard decisions of every check; __syncthreads_or gives the early stop
// so a message slot is v2c before the check-node phase and c2v after it; HBM sees a frame's LLRs (once per iteration,
// from L2) and its results, never its messages: 4n (+4n with posteriors) instead of 16E + 4n bytes per frame-iteration.
// Persistent grid: as many blocks as fit the device, each taking frames blockIdx.x, += gridDim.x -- a frame that stops
// early frees its block at once, so early stop needs no frame compaction here.
//
// Layout (second version; the first one walked one WorkItem per node with 32-bit tables and check-major slots: 71
// instructions per edge-iteration, 12 warps per issue waiting on index loads, 2.5 shared-memory wavefronts per access).
// Nodes are visited by DEGREE CLASS (all checks / variables of one degree: the class loop is block-uniform, the node
// loop inside it is straight-line code for that degree), and the messages of a class are stored edge-major:
//     edge k of the c-th check of a class of `count` checks at  first_slot + k * count + c
// so the lanes of a warp -- consecutive checks -- read and write consecutive shared-memory words (no bank conflicts
// on the check side) and consecutive entries of every index table (coalesced 16-bit loads: E and n are below 65 536
// for anything that fits an SM).  The variable side gathers through its slot lists, stored transposed the same way.
// Low-degree classes take two nodes per thread at a time so that the index loads of both are in flight together.
//
// Same arithmetic contract as the per-iteration kernels (helpers shared through ldpc_cn_common.cuh / ldpc_device.cuh):
// first-argmin / min2 rule, fl(beta*raw) with the sign product as an XOR of sign bits, float32 threshold compares and
// lower-bin-edge reconstruction (the reconstructed VALUE is what is stored), library summation orders, posterior
// without alpha, stop on the first zero syndrome.  float32 decoders, variable degree <= 64.
// Reference: neural_2d_decoder.py:133-225, neural_minsum_decoder.py:58-150, rcq_decoder.py:190-279 / :495-597.
#include <type_traits>

#include "ldpc_cn_common.cuh"

namespace ldpc {

namespace {

enum { RES_NORMALIZED = 0, RES_QUANT = 1, RES_OFFSET = 2 };
constexpr int kResMaxDv = 64;
// nodes of degree <= these take two nodes per thread at a time (tuning macros).  Off: at 1024 threads a thread has 64
// registers and the paired variants spill ((16200,7200)-shaped, 148 frames: 273 us unpaired, 288 us with pairs for
// variable degrees <= 3, 326 us with pairs up to degree 4 on both sides).
#ifndef LDPC_RES_CN_U2_MAX
#define LDPC_RES_CN_U2_MAX 0
#endif
#ifndef LDPC_RES_VN_U2_MAX
#define LDPC_RES_VN_U2_MAX -1
#endif

template <int N>
using IntC = std::integral_constant<int, N>;

template <int KIND, int NTH>
__global__ void __launch_bounds__(1024, 1) resident_decode_kernel(const ResidentLaunch p) {
    constexpr bool QUANT = KIND == RES_QUANT;
    extern __shared__ __align__(16) unsigned char res_smem[];
    __shared__ WorkItem s_cls[2 * kResMaxClasses];
    float* const msg = reinterpret_cast<float*>(res_smem);                             // [E] messages, in place
    uint8_t* const hbit = res_smem + (size_t)p.E * sizeof(float);                      // [n] hard decisions
    float* const s_thr = reinterpret_cast<float*>(res_smem + (size_t)p.E * sizeof(float) + (((size_t)p.n + 15) & ~(size_t)15));
    const float* const s_lut = s_thr + p.n_quant * p.nth;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int E = p.E, n = p.n;
    const int ncc = p.n_cclass, nvc = p.n_vclass;
    for (int i = tid; i < ncc + nvc; i += nthr) s_cls[i] = p.classes[i];
    if (QUANT) {
        const int nthr_q = p.n_quant * p.nth, nlut = p.n_quant << p.bc;
        for (int i = tid; i < nthr_q; i += nthr) s_thr[i] = p.thr[i];
        for (int i = tid; i < nlut; i += nthr) s_thr[nthr_q + i] = p.lut[i];
    }
    const bool has_beta = p.beta != nullptr;
    const bool has_alpha = p.alpha != nullptr && KIND != RES_OFFSET;

    for (int64_t f = blockIdx.x; f < p.B; f += gridDim.x) {
        const float* __restrict__ llr = p.llr_rows + f * n;
        float* __restrict__ post = p.post_rows ? p.post_rows + f * n : nullptr;
        __syncthreads();   // the previous frame is delivered; (first frame: the class / quantiser tables are in place)
        for (int s = tid; s < E; s += nthr) msg[s] = __ldg(llr + __ldg(p.slot_var + s));   // ldpc_decoder.py:84-87
        __syncthreads();
        int it_done = p.T;
        int ok = 0;
        for (int t = 0; t < p.T; ++t) {
            const float* __restrict__ beta_t = has_beta ? p.beta + (size_t)t * p.n_beta : nullptr;
            const float* __restrict__ alpha_t = p.alpha ? p.alpha + (size_t)t * p.n_alpha : nullptr;
            Quantizer<NTH> qz;
            const float* lutq = s_lut;
            if (QUANT) {
                const int q = __ldg(p.q_of_iter + t);
                qz.load(s_thr + q * p.nth, p.nth, __ldg(p.mono + q) != 0);
                lutq = s_lut + (q << p.bc);
            }
            const bool last = t == p.T - 1;

            // ---- check nodes of one degree class.  DC > 0: inputs in registers, U checks per thread at a time;
            //      DC == 0: run-time degree, inputs read again for the output phase ----
            auto cn_class = [&](auto dc_tag, auto u_tag, const WorkItem cl) {
                constexpr int DC = decltype(dc_tag)::value;
                constexpr int U = decltype(u_tag)::value;
                constexpr int X = DC > 0 ? DC : 1;
                const int D = DC > 0 ? DC : cl.deg;
                const int m = cl.count;
                for (int c0 = tid; c0 < m; c0 += nthr * U) {
                    float x[U][X];
                    float betac[U];
                    bool live[U];
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        const int c = c0 + u * nthr;
                        live[u] = c < m;
                        betac[u] = 1.f;
                        if (live[u]) {
                            if constexpr (DC > 0) {
#pragma unroll
                                for (int k = 0; k < DC; ++k) x[u][k] = msg[cl.first_slot + k * m + c];
                            }
                            if (has_beta && !p.beta_per_edge)
                                betac[u] = __ldg(beta_t + (p.bidx ? __ldg(p.bidx + cl.first_slot + c) : 0));
                        }
                    }
#pragma unroll
                    for (int u = 0; u < U; ++u) {
                        if (!live[u]) continue;
                        const int c = c0 + u * nthr;
                        float* const io = msg + cl.first_slot + c;        // edge k at io[k * m]
                        const int s0 = cl.first_slot + c;                 // physical slot of edge k: s0 + k * m
                        auto input = [&](int k) -> float {
                            if constexpr (DC > 0) return x[u][k];
                            else return io[k * m];
                        };
                        MinState<float, false> st;
                        st.init();
#pragma unroll
                        for (int k = 0; k < D; ++k) st.push(input(k), k);
                        if (D == 1) st.m2 = st.m1;   // ldpc_decoder.py:112-113
                        if constexpr (KIND == RES_OFFSET) {
                            const float beta_check = (has_beta && !p.beta_per_edge) ? betac[u] : 0.f;
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                const float xk = input(k);
                                const bool is_min = fabsf(xk) == st.m1;
                                const bool zero_others = D > 1 && (st.m2 == 0.f || (st.m1 == 0.f && !is_min));
                                const float beta = (has_beta && p.beta_per_edge) ? __ldg(beta_t + __ldg(p.bidx + s0 + k * m)) : beta_check;
                                const float alpha = alpha_t ? __ldg(alpha_t + (p.aidx_slot ? __ldg(p.aidx_slot + s0 + k * m) : 0)) : 0.f;
                                io[k * m] = offset_value<float>(is_min ? st.m2 : st.m1, beta, has_beta, alpha, alpha_t != nullptr,
                                                                st.par ^ __float_as_uint(xk), zero_others);
                            }
                        } else if (!p.beta_per_edge) {
                            CheckOut<float, QUANT> co;
                            co.prepare(st.m1, st.m2, st.par, betac[u], has_beta, qz, p.bc);
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                const float xk = input(k);
                                const auto o = co.emit(fabsf(xk) == st.m1, __float_as_uint(xk));
                                if constexpr (QUANT) io[k * m] = lutq[o];
                                else io[k * m] = o;
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < D; ++k) {
                                const float xk = input(k);
                                const float raw = (fabsf(xk) == st.m1) ? st.m2 : st.m1;
                                const auto o = cn_emit<float, QUANT, NTH>(raw, __ldg(beta_t + __ldg(p.bidx + s0 + k * m)),
                                                                          st.par ^ __float_as_uint(xk), qz, p.bc);
                                if constexpr (QUANT) io[k * m] = lutq[o];
                                else io[k * m] = o;
                            }
                        }
                    }
                }
            };
            for (int ci = 0; ci < ncc; ++ci) {
                const WorkItem cl = s_cls[ci];
                switch (cl.deg) {   // block-uniform
                    case 1: cn_class(IntC<1>{}, IntC<(1 <= LDPC_RES_CN_U2_MAX ? 2 : 1)>{}, cl); break;
                    case 2: cn_class(IntC<2>{}, IntC<(2 <= LDPC_RES_CN_U2_MAX ? 2 : 1)>{}, cl); break;
                    case 3: cn_class(IntC<3>{}, IntC<(3 <= LDPC_RES_CN_U2_MAX ? 2 : 1)>{}, cl); break;
                    case 4: cn_class(IntC<4>{}, IntC<(4 <= LDPC_RES_CN_U2_MAX ? 2 : 1)>{}, cl); break;
                    case 5: cn_class(IntC<5>{}, IntC<1>{}, cl); break;
                    case 6: cn_class(IntC<6>{}, IntC<1>{}, cl); break;
                    case 7: cn_class(IntC<7>{}, IntC<1>{}, cl); break;
                    case 8: cn_class(IntC<8>{}, IntC<1>{}, cl); break;
                    default: cn_class(IntC<0>{}, IntC<1>{}, cl); break;
                }
            }
            __syncthreads();

            // ---- variable nodes of one degree class: inputs in registers (DV <= 8, U variables per thread at a time) or
            //      a local array ----
            auto vn_class = [&](auto dv_tag, auto u_tag, const WorkItem cl) {
                constexpr int DV = decltype(dv_tag)::value;
                constexpr int U = decltype(u_tag)::value;
                const int m = cl.count;
                const uint16_t* __restrict__ lists = p.vslots + cl.first_slot;   // entry d of variable i at lists[d * m + i]
                if constexpr (DV >= 0) {
                    constexpr int D1 = DV > 0 ? DV : 1;
                    for (int i0 = tid; i0 < m; i0 += nthr * U) {
                        int sl[U][D1];
                        int j[U];
                        float c[U][D1];
                        float L[U], alpha[U];
                        bool live[U];
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            const int i = i0 + u * nthr;
                            live[u] = i < m;
                            j[u] = 0;
                            if (live[u]) {
                                j[u] = __ldg(p.vpos_var + cl.first_node + i);
#pragma unroll
                                for (int d = 0; d < DV; ++d) sl[u][d] = __ldg(lists + d * m + i);
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            L[u] = 0.f;
                            alpha[u] = 1.f;
                            if (live[u]) {
                                L[u] = __ldg(llr + j[u]);
                                if (has_alpha && !last) alpha[u] = __ldg(alpha_t + (p.aidx ? __ldg(p.aidx + cl.first_node + i0 + u * nthr) : 0));
#pragma unroll
                                for (int d = 0; d < DV; ++d) c[u][d] = msg[sl[u][d]];
                            }
                        }
#pragma unroll
                        for (int u = 0; u < U; ++u) {
                            if (!live[u]) continue;
                            const float tot = LibSum<float>::template stat<DV>([&](int i) { return c[u][i]; });
                            const float pv = DV > 0 ? __fadd_rn(L[u], tot) : L[u];
                            if (!last) {   // the v2c update of iteration T-1 is dead
#pragma unroll
                                for (int d = 0; d < DV; ++d) {
                                    float sd = LibSum<float>::template stat<(DV > 0 ? DV - 1 : 0)>([&](int i) { return c[u][i < d ? i : i + 1]; });
                                    if (has_alpha) sd = __fmul_rn(alpha[u], sd);
                                    msg[sl[u][d]] = __fadd_rn(L[u], sd);
                                }
                            }
                            hbit[j[u]] = pv < 0.f ? 1 : 0;
                            if (post) post[j[u]] = pv;   // refreshed every iteration: the row holds the posterior of the stop iteration
                        }
                    }
                } else {
                    const int dvr = cl.deg;
                    for (int i = tid; i < m; i += nthr) {
                        const int j = __ldg(p.vpos_var + cl.first_node + i);
                        const float L = __ldg(llr + j);
                        float alpha = 1.f;
                        if (has_alpha && !last) alpha = __ldg(alpha_t + (p.aidx ? __ldg(p.aidx + cl.first_node + i) : 0));
                        int sl[kResMaxDv];
                        float c[kResMaxDv];
                        for (int d = 0; d < dvr; ++d) {
                            sl[d] = __ldg(lists + d * m + i);
                            c[d] = msg[sl[d]];
                        }
                        const float tot = LibSum<float>::dyn([&](int k) { return c[k]; }, dvr);
                        const float pv = __fadd_rn(L, tot);
                        if (!last) {
                            for (int d = 0; d < dvr; ++d) {
                                float sd = LibSum<float>::dyn([&](int k) { return c[k < d ? k : k + 1]; }, dvr - 1);
                                if (has_alpha) sd = __fmul_rn(alpha, sd);
                                msg[sl[d]] = __fadd_rn(L, sd);
                            }
                        }
                        hbit[j] = pv < 0.f ? 1 : 0;
                        if (post) post[j] = pv;
                    }
                }
            };
            for (int vi = 0; vi < nvc; ++vi) {
                const WorkItem cl = s_cls[ncc + vi];
                switch (cl.deg) {   // block-uniform
                    case 0: vn_class(IntC<0>{}, IntC<(0 <= LDPC_RES_VN_U2_MAX ? 2 : 1)>{}, cl); break;
                    case 1: vn_class(IntC<1>{}, IntC<(1 <= LDPC_RES_VN_U2_MAX ? 2 : 1)>{}, cl); break;
                    case 2: vn_class(IntC<2>{}, IntC<(2 <= LDPC_RES_VN_U2_MAX ? 2 : 1)>{}, cl); break;
                    case 3: vn_class(IntC<3>{}, IntC<(3 <= LDPC_RES_VN_U2_MAX ? 2 : 1)>{}, cl); break;
                    case 4: vn_class(IntC<4>{}, IntC<(4 <= LDPC_RES_VN_U2_MAX ? 2 : 1)>{}, cl); break;
                    case 5: vn_class(IntC<5>{}, IntC<1>{}, cl); break;
                    case 6: vn_class(IntC<6>{}, IntC<1>{}, cl); break;
                    case 7: vn_class(IntC<7>{}, IntC<1>{}, cl); break;
                    case 8: vn_class(IntC<8>{}, IntC<1>{}, cl); break;
                    default: vn_class(IntC<-1>{}, IntC<1>{}, cl); break;
                }
            }
            __syncthreads();

            // ---- syndrome and early stop (ldpc_decoder.py:141-144) ----
            if (p.early_stop || last) {
                int unsat = 0;
                for (int ci = 0; ci < ncc; ++ci) {
                    const WorkItem cl = s_cls[ci];
                    const uint16_t* __restrict__ sv = p.slot_var + cl.first_slot;
                    for (int c = tid; c < cl.count; c += nthr) {
                        uint32_t par = 0;
                        for (int k = 0; k < cl.deg; ++k) par ^= hbit[__ldg(sv + k * cl.count + c)];
                        unsat |= (int)par;
                    }
                }
                if (!__syncthreads_or(unsat)) {   // block-uniform
                    it_done = t + 1;
                    ok = 1;
                    break;
                }
            }
        }
        // ---- deliver the frame ----
        if (p.bits_rows) {
            uint8_t* __restrict__ dst = p.bits_rows + f * n;
            for (int j = tid; j < n; j += nthr) dst[j] = hbit[j];
        }
        if (p.packed_rows) {
            const int nw = (n + 31) >> 5;
            uint32_t* __restrict__ dst = p.packed_rows + f * nw;
            for (int w = tid; w < nw; w += nthr) {
                uint32_t word = 0;
                const int j0 = w << 5, j1 = min(n, j0 + 32);
                for (int j = j0; j < j1; ++j) word |= (uint32_t)hbit[j] << (j - j0);
                dst[w] = word;
            }
        }
        if (tid == 0) {
            if (p.iters) p.iters[f] = it_done;
            if (p.success) p.success[f] = (uint8_t)ok;
        }
    }
}

size_t resident_smem_bytes(const ResidentLaunch& p) {
    size_t b = (size_t)p.E * sizeof(float) + (((size_t)p.n + 15) & ~(size_t)15);
    if (p.bc) b += sizeof(float) * ((size_t)p.n_quant * p.nth + ((size_t)p.n_quant << p.bc));
    return b;
}

int resident_threads(const ResidentLaunch& p) {
    // enough threads that a phase is a handful of nodes per thread, whole warps
    int t = 64;
    while (t < 1024 && t * 8 < p.n) t *= 2;
    return t;
}

// thread blocks of this kernel that one SM holds at a time (0: the launch is impossible)
template <int KIND, int NTH>
cudaError_t resident_blocks_per_sm(const ResidentLaunch& p, int* per_sm) {
    const size_t smem = resident_smem_bytes(p);
    cudaError_t e = cudaFuncSetAttribute(resident_decode_kernel<KIND, NTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    return cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, resident_decode_kernel<KIND, NTH>, resident_threads(p), smem);
}

template <int KIND, int NTH>
cudaError_t launch_resident_t(const ResidentLaunch& p, cudaStream_t stream) {
    int per_sm = 0;
    cudaError_t e = resident_blocks_per_sm<KIND, NTH>(p, &per_sm);
    if (e != cudaSuccess) return e;
    if (per_sm < 1) return cudaErrorInvalidConfiguration;
    int64_t grid = (int64_t)per_sm * p.sm_count;
    if (grid > p.B) grid = p.B;
    resident_decode_kernel<KIND, NTH><<<(unsigned)grid, resident_threads(p), resident_smem_bytes(p), stream>>>(p);
    return cudaGetLastError();
}

// the kernel variant of a decoder
#define LDPC_RESIDENT_DISPATCH(p, CALL)                                      \
    do {                                                                     \
        if ((p).check_rule == 1) return CALL(RES_OFFSET, 0);                 \
        if ((p).bc) {                                                        \
            if ((p).all_mono && (p).nth <= 4) return CALL(RES_QUANT, 4);     \
            if ((p).all_mono && (p).nth <= 8) return CALL(RES_QUANT, 8);     \
            return CALL(RES_QUANT, 0);                                       \
        }                                                                    \
        return CALL(RES_NORMALIZED, 0);                                      \
    } while (0)

cudaError_t resident_blocks_per_sm_any(const ResidentLaunch& p, int* per_sm) {
#define LDPC_RES_CALL(K, N) resident_blocks_per_sm<K, N>(p, per_sm)
    LDPC_RESIDENT_DISPATCH(p, LDPC_RES_CALL);
#undef LDPC_RES_CALL
}

}  // namespace

bool resident_decode_fits(const ResidentLaunch& p) {
    if (p.E <= 0 || p.n <= 0 || p.max_dv > kResMaxDv || !p.classes) return false;
    if (p.n_cclass > kResMaxClasses || p.n_vclass > kResMaxClasses) return false;
    return resident_smem_bytes(p) <= (size_t)220 * 1024;
}

cudaError_t launch_resident_decode(const ResidentLaunch& p, cudaStream_t stream) {
#define LDPC_RES_CALL(K, N) launch_resident_t<K, N>(p, stream)
    LDPC_RESIDENT_DISPATCH(p, LDPC_RES_CALL);
#undef LDPC_RES_CALL
}

// frames that decode at the same time: one thread block each, as many blocks as the device holds (0 on error)
int64_t resident_wave_frames(const ResidentLaunch& p) {
    int per_sm = 0;
    if (resident_blocks_per_sm_any(p, &per_sm) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return (int64_t)per_sm * p.sm_count;
}

}  // namespace ldpc
