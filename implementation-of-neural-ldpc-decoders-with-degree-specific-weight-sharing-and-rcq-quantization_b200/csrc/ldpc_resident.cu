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
// loop inside it is straight-line code for that degree), and the messages of a class are stored in tiles of 32 checks,
// edge-major inside a tile:
//     edge k of the c-th check of a class at  first_slot + (c / 32) * 32 * deg + k * 32 + c % 32
// so the lanes of a warp -- consecutive checks -- read and write consecutive shared-memory words (no bank conflicts
// on the check side) and consecutive entries of every index table (coalesced 16-bit loads: E and n are below 65 536
// for anything that fits an SM), and the edges of a node are a COMPILE-TIME stride apart: one address per node, the
// rest are immediate offsets (with a run-time stride -- the class size -- address arithmetic was a third of all
// instructions).  The variable side gathers through its slot lists, stored in the same tiled order.
// A thread's nodes of a class are visited in a software-pipelined loop: the index-table entries (and the LLR / weight
// they lead to) of the thread's NEXT node are requested before its current node is worked on, so the L2 round trips
// of the tables -- 194 KB of messages leave no room for them in an SM -- overlap the shared-memory work.
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
// variable degrees up to this run the software-pipelined loop (tuning macro): with 1024 threads a thread has 64
// registers, and the pipeline registers of higher degrees spill ((16200,7200)-shaped, 148 frames: 235 us with 2,
// 241 us with 3, 245 us with 4; unpipelined 273 us)
#ifndef LDPC_RES_PIPE_MAX_DV
#define LDPC_RES_PIPE_MAX_DV 2
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

            // ---- check nodes of one degree class.  DC > 0: inputs in registers; DC == 0: run-time degree, inputs read
            //      again for the output phase.  The beta of the thread's next check is fetched while this one is worked on
            //      (column index two checks ahead, value one ahead). ----
            auto cn_class = [&](auto dc_tag, const WorkItem cl) {
                constexpr int DC = decltype(dc_tag)::value;
                constexpr int X = DC > 0 ? DC : 1;
                const int D = DC > 0 ? DC : cl.deg;
                const int m = cl.count;
                const int lane = tid & 31;
                auto node_off = [&](int c) { return (c - lane) * D + lane; };   // first edge of check c (c % 32 == lane)
                const bool one_beta = has_beta && !p.beta_per_edge;
                const int32_t* __restrict__ bx = (one_beta && p.bidx) ? p.bidx + cl.first_slot : nullptr;   // column of edge 0
                float beta = 1.f, beta_next = 1.f;
                int col_next = 0;
                if (one_beta) {
                    if (!bx) {
                        beta = beta_next = __ldg(beta_t);
                    } else {
                        if (tid < m) beta = __ldg(beta_t + __ldg(bx + node_off(tid)));
                        if (tid + nthr < m) col_next = __ldg(bx + node_off(tid + nthr));
                    }
                }
                for (int c = tid; c < m; c += nthr) {
                    int col_next2 = 0;
                    if (bx) {
                        if (c + nthr < m) beta_next = __ldg(beta_t + col_next);
                        if (c + 2 * nthr < m) col_next2 = __ldg(bx + node_off(c + 2 * nthr));
                    }
                    const int s0 = cl.first_slot + node_off(c);       // physical slot of edge k: s0 + k * 32
                    float* const io = msg + s0;                       // edge k at io[k * 32]
                    float x[X];
                    if constexpr (DC > 0) {
#pragma unroll
                        for (int k = 0; k < DC; ++k) x[k] = io[k * 32];
                    }
                    auto input = [&](int k) -> float {
                        if constexpr (DC > 0) return x[k];
                        else return io[k * 32];
                    };
                    MinState<float, false> st;
                    st.init();
#pragma unroll
                    for (int k = 0; k < D; ++k) st.push(input(k), k);
                    if (D == 1) st.m2 = st.m1;   // ldpc_decoder.py:112-113
                    if constexpr (KIND == RES_OFFSET) {
                        const float beta_check = one_beta ? beta : 0.f;
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const float xk = input(k);
                            const bool is_min = fabsf(xk) == st.m1;
                            const bool zero_others = D > 1 && (st.m2 == 0.f || (st.m1 == 0.f && !is_min));
                            const float b = (has_beta && p.beta_per_edge) ? __ldg(beta_t + __ldg(p.bidx + s0 + k * 32)) : beta_check;
                            const float alpha = alpha_t ? __ldg(alpha_t + (p.aidx_slot ? __ldg(p.aidx_slot + s0 + k * 32) : 0)) : 0.f;
                            io[k * 32] = offset_value<float>(is_min ? st.m2 : st.m1, b, has_beta, alpha, alpha_t != nullptr,
                                                            st.par ^ __float_as_uint(xk), zero_others);
                        }
                    } else if (!p.beta_per_edge) {
                        CheckOut<float, QUANT> co;
                        co.prepare(st.m1, st.m2, st.par, beta, has_beta, qz, p.bc);
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const float xk = input(k);
                            const auto o = co.emit(fabsf(xk) == st.m1, __float_as_uint(xk));
                            if constexpr (QUANT) io[k * 32] = lutq[o];
                            else io[k * 32] = o;
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < D; ++k) {
                            const float xk = input(k);
                            const float raw = (fabsf(xk) == st.m1) ? st.m2 : st.m1;
                            const auto o = cn_emit<float, QUANT, NTH>(raw, __ldg(beta_t + __ldg(p.bidx + s0 + k * 32)),
                                                                      st.par ^ __float_as_uint(xk), qz, p.bc);
                            if constexpr (QUANT) io[k * 32] = lutq[o];
                            else io[k * 32] = o;
                        }
                    }
                    beta = beta_next;
                    col_next = col_next2;
                }
            };
            for (int ci = 0; ci < ncc; ++ci) {
                const WorkItem cl = s_cls[ci];
                switch (cl.deg) {   // block-uniform
                    case 1: cn_class(IntC<1>{}, cl); break;
                    case 2: cn_class(IntC<2>{}, cl); break;
                    case 3: cn_class(IntC<3>{}, cl); break;
                    case 4: cn_class(IntC<4>{}, cl); break;
                    case 5: cn_class(IntC<5>{}, cl); break;
                    case 6: cn_class(IntC<6>{}, cl); break;
                    case 7: cn_class(IntC<7>{}, cl); break;
                    case 8: cn_class(IntC<8>{}, cl); break;
                    default: cn_class(IntC<0>{}, cl); break;
                }
            }
            __syncthreads();

            // ---- variable nodes of one degree class: inputs in registers (DV <= 8) or a local array.  Software pipeline
            //      over the thread's variables: while variable i is worked on, the slot list, LLR and alpha of i + nthr and
            //      the variable id / alpha column of i + 2 nthr are on their way. ----
            auto vn_class = [&](auto dv_tag, const WorkItem cl) {
                constexpr int DV = decltype(dv_tag)::value;
                const int m = cl.count;
                const int lane = tid & 31;
                const int dstride = DV >= 0 ? DV : cl.deg;
                // entry d of variable i (i % 32 == lane) at lists[(i - lane) * deg + lane + d * 32]
                const uint16_t* __restrict__ lists = p.vslots + cl.first_slot + lane;
                const uint16_t* __restrict__ vp = p.vpos_var + cl.first_node;
                const bool use_alpha = has_alpha && !last;
                const int32_t* __restrict__ ax = (use_alpha && p.aidx) ? p.aidx + cl.first_node : nullptr;
                const float alpha_one = (use_alpha && !ax) ? __ldg(alpha_t) : 1.f;
                if constexpr (DV >= 0 && DV <= LDPC_RES_PIPE_MAX_DV) {
                    constexpr int D1 = DV > 0 ? DV : 1;
                    int sl[D1];
                    int j = 0, j_next = 0, acol_next = 0;
                    float L = 0.f, alpha = alpha_one;
                    if (tid < m) {
                        j = __ldg(vp + tid);
#pragma unroll
                        for (int d = 0; d < DV; ++d) sl[d] = __ldg(lists + (tid - lane) * dstride + d * 32);
                    }
                    if (tid + nthr < m) {
                        j_next = __ldg(vp + tid + nthr);
                        if (ax) acol_next = __ldg(ax + tid + nthr);
                    }
                    if (tid < m) {
                        L = __ldg(llr + j);
                        if (ax) alpha = __ldg(alpha_t + __ldg(ax + tid));
                    }
                    for (int i = tid; i < m; i += nthr) {
                        int sl_next[D1];
                        int j_next2 = 0, acol_next2 = 0;
                        float L_next = 0.f, alpha_next = alpha_one;
                        if (i + nthr < m) {
#pragma unroll
                            for (int d = 0; d < DV; ++d) sl_next[d] = __ldg(lists + (i + nthr - lane) * dstride + d * 32);
                            L_next = __ldg(llr + j_next);
                            if (ax) alpha_next = __ldg(alpha_t + acol_next);
                        }
                        if (i + 2 * nthr < m) {
                            j_next2 = __ldg(vp + i + 2 * nthr);
                            if (ax) acol_next2 = __ldg(ax + i + 2 * nthr);
                        }
                        float c[D1];
#pragma unroll
                        for (int d = 0; d < DV; ++d) c[d] = msg[sl[d]];
                        const float tot = LibSum<float>::template stat<DV>([&](int k) { return c[k]; });
                        const float pv = DV > 0 ? __fadd_rn(L, tot) : L;
                        if (!last) {   // the v2c update of iteration T-1 is dead
#pragma unroll
                            for (int d = 0; d < DV; ++d) {
                                float sd = LibSum<float>::template stat<(DV > 0 ? DV - 1 : 0)>([&](int k) { return c[k < d ? k : k + 1]; });
                                if (has_alpha) sd = __fmul_rn(alpha, sd);
                                msg[sl[d]] = __fadd_rn(L, sd);
                            }
                        }
                        hbit[j] = pv < 0.f ? 1 : 0;
                        if (post) post[j] = pv;   // refreshed every iteration: the row holds the posterior of the stop iteration
#pragma unroll
                        for (int d = 0; d < DV; ++d) sl[d] = sl_next[d];
                        j = j_next;
                        j_next = j_next2;
                        L = L_next;
                        alpha = alpha_next;
                        acol_next = acol_next2;
                    }
                } else if constexpr (DV >= 0) {   // higher degrees: the pipeline registers would spill (64 per thread)
                    for (int i = tid; i < m; i += nthr) {
                        const int j = __ldg(vp + i);
                        int sl[DV];
#pragma unroll
                        for (int d = 0; d < DV; ++d) sl[d] = __ldg(lists + (i - lane) * dstride + d * 32);
                        const float L = __ldg(llr + j);
                        const float alpha = ax ? __ldg(alpha_t + __ldg(ax + i)) : alpha_one;
                        float c[DV];
#pragma unroll
                        for (int d = 0; d < DV; ++d) c[d] = msg[sl[d]];
                        const float tot = LibSum<float>::template stat<DV>([&](int k) { return c[k]; });
                        const float pv = __fadd_rn(L, tot);
                        if (!last) {
#pragma unroll
                            for (int d = 0; d < DV; ++d) {
                                float sd = LibSum<float>::template stat<DV - 1>([&](int k) { return c[k < d ? k : k + 1]; });
                                if (has_alpha) sd = __fmul_rn(alpha, sd);
                                msg[sl[d]] = __fadd_rn(L, sd);
                            }
                        }
                        hbit[j] = pv < 0.f ? 1 : 0;
                        if (post) post[j] = pv;
                    }
                } else {
                    const int dvr = cl.deg;
                    for (int i = tid; i < m; i += nthr) {
                        const int j = __ldg(vp + i);
                        const float L = __ldg(llr + j);
                        const float alpha = ax ? __ldg(alpha_t + __ldg(ax + i)) : alpha_one;
                        int sl[kResMaxDv];
                        float c[kResMaxDv];
                        for (int d = 0; d < dvr; ++d) {
                            sl[d] = __ldg(lists + (i - lane) * dstride + d * 32);
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
                    case 0: vn_class(IntC<0>{}, cl); break;
                    case 1: vn_class(IntC<1>{}, cl); break;
                    case 2: vn_class(IntC<2>{}, cl); break;
                    case 3: vn_class(IntC<3>{}, cl); break;
                    case 4: vn_class(IntC<4>{}, cl); break;
                    case 5: vn_class(IntC<5>{}, cl); break;
                    case 6: vn_class(IntC<6>{}, cl); break;
                    case 7: vn_class(IntC<7>{}, cl); break;
                    case 8: vn_class(IntC<8>{}, cl); break;
                    default: vn_class(IntC<-1>{}, cl); break;
                }
            }
            __syncthreads();

            // ---- syndrome and early stop (ldpc_decoder.py:141-144) ----
            if (p.early_stop || last) {
                int unsat = 0;
                // two checks per thread at a time, all their variable ids requested before the first decision is read
                auto syn_class = [&](auto dc_tag, const WorkItem cl) {
                    constexpr int DC = decltype(dc_tag)::value;
                    constexpr int X = DC > 0 ? DC : 1;
                    const int m = cl.count;
                    const int lane = tid & 31;
                    const int dstride = DC > 0 ? DC : cl.deg;
                    const uint16_t* __restrict__ sv = p.slot_var + cl.first_slot + lane;   // edge k of check c at sv[(c - lane) * deg + k * 32]
                    if constexpr (DC > 0) {
                        for (int c = tid; c < m; c += 2 * nthr) {
                            const bool two = c + nthr < m;
                            int v0[X], v1[X];
#pragma unroll
                            for (int k = 0; k < DC; ++k) {
                                v0[k] = __ldg(sv + (c - lane) * dstride + k * 32);
                                v1[k] = two ? __ldg(sv + (c + nthr - lane) * dstride + k * 32) : 0;
                            }
                            uint32_t par0 = 0, par1 = 0;
#pragma unroll
                            for (int k = 0; k < DC; ++k) {
                                par0 ^= hbit[v0[k]];
                                par1 ^= hbit[v1[k]];
                            }
                            unsat |= (int)(par0 | (two ? par1 : 0u));
                        }
                    } else {
                        for (int c = tid; c < m; c += nthr) {
                            uint32_t par = 0;
                            for (int k = 0; k < cl.deg; ++k) par ^= hbit[__ldg(sv + (c - lane) * dstride + k * 32)];
                            unsat |= (int)par;
                        }
                    }
                };
                for (int ci = 0; ci < ncc; ++ci) {
                    const WorkItem cl = s_cls[ci];
                    switch (cl.deg) {   // block-uniform
                        case 1: syn_class(IntC<1>{}, cl); break;
                        case 2: syn_class(IntC<2>{}, cl); break;
                        case 3: syn_class(IntC<3>{}, cl); break;
                        case 4: syn_class(IntC<4>{}, cl); break;
                        case 5: syn_class(IntC<5>{}, cl); break;
                        case 6: syn_class(IntC<6>{}, cl); break;
                        case 7: syn_class(IntC<7>{}, cl); break;
                        case 8: syn_class(IntC<8>{}, cl); break;
                        default: syn_class(IntC<0>{}, cl); break;
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
