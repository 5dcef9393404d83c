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

template <int KIND, int NTH>
__global__ void __launch_bounds__(1024, 1) resident_decode_kernel(const ResidentLaunch p) {
    constexpr bool QUANT = KIND == RES_QUANT;
    extern __shared__ __align__(16) unsigned char res_smem[];
    float* const msg = reinterpret_cast<float*>(res_smem);                             // [E] messages, in place
    uint8_t* const hbit = res_smem + (size_t)p.E * sizeof(float);                      // [n] hard decisions
    float* const s_thr = reinterpret_cast<float*>(res_smem + (size_t)p.E * sizeof(float) + (((size_t)p.n + 15) & ~(size_t)15));
    const float* const s_lut = s_thr + p.n_quant * p.nth;
    const int tid = threadIdx.x, nthr = blockDim.x;
    const int E = p.E, n = p.n, nc = p.n_checks;
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
        __syncthreads();   // the previous frame is delivered; (first frame: the quantiser tables are in place)
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
            // ---- check nodes: one thread per check.  DC > 0: inputs in registers; DC == 0: run-time degree, inputs
            //      read again for the output phase (an output depends on its own input and the check's summary) ----
            auto cn_node = [&](auto dc_tag, const int deg, const int s0) {
                constexpr int DC = decltype(dc_tag)::value;
                const int D = DC > 0 ? DC : deg;
                float* const io = msg + s0;
                float x[DC > 0 ? DC : 1];
                MinState<float, false> st;
                st.init();
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const float xk = io[k];
                    if constexpr (DC > 0) x[k] = xk;
                    st.push(xk, k);
                }
                if (D == 1) st.m2 = st.m1;   // ldpc_decoder.py:112-113
                auto input = [&](int k) -> float {
                    if constexpr (DC > 0) return x[k];
                    else return io[k];
                };
                if constexpr (KIND == RES_OFFSET) {
                    float beta_check = 0.f;
                    if (has_beta && !p.beta_per_edge) beta_check = __ldg(beta_t + (p.bidx ? __ldg(p.bidx + s0) : 0));
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const float xk = input(k);
                        const bool is_min = fabsf(xk) == st.m1;
                        const bool zero_others = D > 1 && (st.m2 == 0.f || (st.m1 == 0.f && !is_min));
                        const float beta = (has_beta && p.beta_per_edge) ? __ldg(beta_t + __ldg(p.bidx + s0 + k)) : beta_check;
                        const float alpha = alpha_t ? __ldg(alpha_t + (p.aidx_slot ? __ldg(p.aidx_slot + s0 + k) : 0)) : 0.f;
                        io[k] = offset_value<float>(is_min ? st.m2 : st.m1, beta, has_beta, alpha, alpha_t != nullptr,
                                                    st.par ^ __float_as_uint(xk), zero_others);
                    }
                } else if (!p.beta_per_edge) {
                    const float beta = has_beta ? __ldg(beta_t + (p.bidx ? __ldg(p.bidx + s0) : 0)) : 1.f;
                    CheckOut<float, QUANT> co;
                    co.prepare(st.m1, st.m2, st.par, beta, has_beta, qz, p.bc);
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const float xk = input(k);
                        const auto o = co.emit(fabsf(xk) == st.m1, __float_as_uint(xk));
                        if constexpr (QUANT) io[k] = lutq[o];
                        else io[k] = o;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const float xk = input(k);
                        const float raw = (fabsf(xk) == st.m1) ? st.m2 : st.m1;
                        const auto o = cn_emit<float, QUANT, NTH>(raw, __ldg(beta_t + __ldg(p.bidx + s0 + k)), st.par ^ __float_as_uint(xk), qz, p.bc);
                        if constexpr (QUANT) io[k] = lutq[o];
                        else io[k] = o;
                    }
                }
            };
            for (int c = tid; c < nc; c += nthr) {
                const WorkItem it = p.cn_items[c];
                switch (it.deg) {
                    case 1: cn_node(std::integral_constant<int, 1>{}, 1, it.first_slot); break;
                    case 2: cn_node(std::integral_constant<int, 2>{}, 2, it.first_slot); break;
                    case 3: cn_node(std::integral_constant<int, 3>{}, 3, it.first_slot); break;
                    case 4: cn_node(std::integral_constant<int, 4>{}, 4, it.first_slot); break;
                    case 5: cn_node(std::integral_constant<int, 5>{}, 5, it.first_slot); break;
                    case 6: cn_node(std::integral_constant<int, 6>{}, 6, it.first_slot); break;
                    case 7: cn_node(std::integral_constant<int, 7>{}, 7, it.first_slot); break;
                    case 8: cn_node(std::integral_constant<int, 8>{}, 8, it.first_slot); break;
                    default: cn_node(std::integral_constant<int, 0>{}, it.deg, it.first_slot); break;
                }
            }
            __syncthreads();
            // ---- variable nodes: one thread per variable; inputs in registers (DV <= 8) or a local array ----
            auto vn_node = [&](auto dv_tag, const int dvr, const int lb, const int pos) {
                constexpr int DV = decltype(dv_tag)::value;
                const int j = __ldg(p.vpos_var + pos);
                const float L = __ldg(llr + j);
                float pv;
                float alpha = 1.f;
                if (has_alpha && !last) alpha = __ldg(alpha_t + (p.aidx ? __ldg(p.aidx + pos) : 0));
                if constexpr (DV >= 0) {
                    constexpr int D1 = DV > 0 ? DV : 1;
                    int sl[D1];
                    float c[D1];
#pragma unroll
                    for (int i = 0; i < DV; ++i) {
                        sl[i] = __ldg(p.vslots + lb + i);
                        c[i] = msg[sl[i]];
                    }
                    const float tot = LibSum<float>::template stat<DV>([&](int i) { return c[i]; });
                    pv = DV > 0 ? __fadd_rn(L, tot) : L;
                    if (!last) {   // the v2c update of iteration T-1 is dead
#pragma unroll
                        for (int d = 0; d < DV; ++d) {
                            float sd = LibSum<float>::template stat<(DV > 0 ? DV - 1 : 0)>([&](int i) { return c[i < d ? i : i + 1]; });
                            if (has_alpha) sd = __fmul_rn(alpha, sd);
                            msg[sl[d]] = __fadd_rn(L, sd);
                        }
                    }
                } else {
                    int sl[kResMaxDv];
                    float c[kResMaxDv];
                    for (int i = 0; i < dvr; ++i) {
                        sl[i] = __ldg(p.vslots + lb + i);
                        c[i] = msg[sl[i]];
                    }
                    const float tot = LibSum<float>::dyn([&](int i) { return c[i]; }, dvr);
                    pv = __fadd_rn(L, tot);
                    if (!last) {
                        for (int d = 0; d < dvr; ++d) {
                            float sd = LibSum<float>::dyn([&](int i) { return c[i < d ? i : i + 1]; }, dvr - 1);
                            if (has_alpha) sd = __fmul_rn(alpha, sd);
                            msg[sl[d]] = __fadd_rn(L, sd);
                        }
                    }
                }
                hbit[j] = pv < 0.f ? 1 : 0;
                if (post) post[j] = pv;   // refreshed every iteration: the row holds the posterior of the stop iteration
            };
            for (int pos = tid; pos < n; pos += nthr) {
                const WorkItem it = p.vn_items[pos];
                switch (it.deg) {
                    case 0: vn_node(std::integral_constant<int, 0>{}, 0, it.first_slot, pos); break;
                    case 1: vn_node(std::integral_constant<int, 1>{}, 1, it.first_slot, pos); break;
                    case 2: vn_node(std::integral_constant<int, 2>{}, 2, it.first_slot, pos); break;
                    case 3: vn_node(std::integral_constant<int, 3>{}, 3, it.first_slot, pos); break;
                    case 4: vn_node(std::integral_constant<int, 4>{}, 4, it.first_slot, pos); break;
                    case 5: vn_node(std::integral_constant<int, 5>{}, 5, it.first_slot, pos); break;
                    case 6: vn_node(std::integral_constant<int, 6>{}, 6, it.first_slot, pos); break;
                    case 7: vn_node(std::integral_constant<int, 7>{}, 7, it.first_slot, pos); break;
                    case 8: vn_node(std::integral_constant<int, 8>{}, 8, it.first_slot, pos); break;
                    default: vn_node(std::integral_constant<int, -1>{}, it.deg, it.first_slot, pos); break;
                }
            }
            __syncthreads();
            // ---- syndrome and early stop (ldpc_decoder.py:141-144) ----
            if (p.early_stop || last) {
                int unsat = 0;
                for (int c = tid; c < nc; c += nthr) {
                    const WorkItem it = p.cn_items[c];
                    uint32_t par = 0;
                    for (int k = 0; k < it.deg; ++k) par ^= hbit[__ldg(p.slot_var + it.first_slot + k)];
                    unsat |= (int)par;
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
    const int nodes = p.n > p.n_checks ? p.n : p.n_checks;
    int t = 64;
    while (t < 1024 && t * 8 < nodes) t *= 2;
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
    if (p.E <= 0 || p.n <= 0 || p.max_dv > kResMaxDv) return false;
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
