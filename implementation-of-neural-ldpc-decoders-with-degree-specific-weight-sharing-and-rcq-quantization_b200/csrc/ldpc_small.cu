// On-chip decode for small codes: ONE launch runs all T flooding iterations of a frame -- check nodes, weights, RCQ
// quantise / reconstruct, variable nodes, posterior, hard decision, syndrome and early stop -- with the frame's messages
// held in shared memory, never in HBM.  One thread owns one frame ("lanes over frames"): its messages are its own
// column of the shared-memory arrays (conflict-free), the Tanner graph tables are copied to shared memory once per CTA
// and read with warp-uniform (broadcast) accesses.
//
// The per-iteration kernels (ldpc_cn.cu / ldpc_vn.cu) move 16E + 4n bytes per frame-iteration through HBM and need
// ~4 launches per iteration; for a code whose whole state is a few hundred bytes per frame both costs vanish here:
// HBM sees the LLRs once and the results once.  Same arithmetic contract as those kernels (the helper code is shared):
// first-argmin / min2 rule, fl(beta*raw) with the sign product as an XOR of sign bits, float32 threshold compares,
// lower-bin-edge reconstruction, library summation orders (torch.sum / np.sum), posterior without alpha, stop on the
// first zero syndrome.  Reference: ldpc_decoder.py:80-153, neural_2d_decoder.py:133-225, neural_minsum_decoder.py:58-150,
// rcq_decoder.py:190-279 / :495-597.
//
// Two I/O modes.  Row mode (ldpc_decode_device / ldpc_decode_host): the kernel reads the caller's row-major LLRs and
// writes the caller's row-major decisions / posteriors / iterations / success itself (coalesced through a shared-memory
// transpose), so a decode is exactly one launch and no workspace exists.  Workspace mode (ldpc_mc_round): LLRs from
// the interleaved llrT the AWGN kernel fills, results in the workspace layout (packed decisions, per-frame iterations)
// that the error-counting kernels read.
#include <type_traits>

#include "ldpc_cn_common.cuh"

namespace ldpc {

namespace {

constexpr int kSmallThreads = 128;
constexpr int kSmallMaxDv = 64;   // run-time-degree variable nodes gather their inputs into a local array

enum { SMALL_NORMALIZED = 0, SMALL_QUANT = 1, SMALL_OFFSET = 2 };

// library reduction of k elements: unrolled add chains for the degrees that occur, the run-time loop beyond
template <typename Real, typename Get>
__device__ __forceinline__ Real lib_sum(Get get, int k) {
    switch (k) {
        case 0: return Real(0);
        case 1: return LibSum<Real>::template stat<1>(get);
        case 2: return LibSum<Real>::template stat<2>(get);
        case 3: return LibSum<Real>::template stat<3>(get);
        case 4: return LibSum<Real>::template stat<4>(get);
        case 5: return LibSum<Real>::template stat<5>(get);
        case 6: return LibSum<Real>::template stat<6>(get);
        case 7: return LibSum<Real>::template stat<7>(get);
        case 8: return LibSum<Real>::template stat<8>(get);
        default: return LibSum<Real>::dyn(get, k);
    }
}

// int tables in shared memory, in this order
struct SmallTables {
    const int32_t* cdeg;    // [nc] degree of each non-empty check (slot order)
    const int32_t* cslot;   // [nc] its first message slot
    const int32_t* svar;    // [E]  variable of each slot
    const int32_t* vdeg;    // [n]  degree of each variable (degree-sorted position order)
    const int32_t* vbase;   // [n]  offset of its slot list
    const int32_t* vid;     // [n]  variable index
    const int32_t* vslot;   // [E]  slot lists, ascending check index inside a variable
    const int32_t* bidx;    // [E]  beta column per slot (or nullptr)
    const int32_t* aidx;    // [n]  alpha column per position (or nullptr)
    const int32_t* aslot;   // [E]  alpha column per slot (offset rule; or nullptr)
};

template <typename Real, int KIND, int NTH>
__global__ void __launch_bounds__(kSmallThreads) small_decode_kernel(const SmallLaunch p) {
    constexpr int V = FramesPerLane<Real>::value;
    constexpr bool QUANT = KIND == SMALL_QUANT;
    extern __shared__ __align__(16) unsigned char small_smem[];
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int E = p.E, n = p.n, nc = p.n_checks, nw = (n + 31) >> 5;
    // ---- shared-memory carve-up: per-frame columns first, then the tables ----
    // messages in place: a slot holds v2c before the check-node pass and c2v after it (half the columns of a
    // two-array layout: (7,4) float64 frames 268 -> 164 bytes, 6 -> 10 resident CTAs per SM)
    const bool want_post = p.post_rows != nullptr || p.postT != nullptr;
    const int ncol = E + n + (want_post ? n : 0);
    Real* const msg = reinterpret_cast<Real*>(small_smem) + tid;                 // [E][threads]
    Real* const llr = msg + (size_t)E * kSmallThreads;                           // [n][threads]
    Real* const pcol = llr + (size_t)n * kSmallThreads;                          // [n][threads] posteriors (only if wanted)
    uint32_t* const hb = reinterpret_cast<uint32_t*>(reinterpret_cast<Real*>(small_smem) + (size_t)ncol * kSmallThreads) + tid;   // [nw][threads]
    int32_t* const tab = reinterpret_cast<int32_t*>(reinterpret_cast<uint32_t*>(reinterpret_cast<Real*>(small_smem) + (size_t)ncol * kSmallThreads) +
                                                    (size_t)nw * kSmallThreads);
    SmallTables g;
    {
        int32_t* w = tab;
        int32_t* cdeg = w; w += nc;
        int32_t* cslot = w; w += nc;
        int32_t* svar = w; w += E;
        int32_t* vdeg = w; w += n;
        int32_t* vbase = w; w += n;
        int32_t* vid = w; w += n;
        int32_t* vslot = w; w += E;
        int32_t* bidx = p.bidx ? w : nullptr; w += p.bidx ? E : 0;
        int32_t* aidx = p.aidx ? w : nullptr; w += p.aidx ? n : 0;
        int32_t* aslot = p.aidx_slot ? w : nullptr; w += p.aidx_slot ? E : 0;
        for (int i = tid; i < nc; i += kSmallThreads) {
            cdeg[i] = p.cn_items[i].deg;
            cslot[i] = p.cn_items[i].first_slot;
        }
        for (int i = tid; i < n; i += kSmallThreads) {
            const WorkItem it = p.vn_items[i];
            vdeg[i] = it.deg;
            vbase[i] = it.first_slot;
            vid[i] = p.vpos_var[it.first_node];
            if (aidx) aidx[i] = p.aidx[it.first_node];
        }
        for (int i = tid; i < E; i += kSmallThreads) {
            svar[i] = p.slot_var[i];
            vslot[i] = p.vslots[i];
            if (bidx) bidx[i] = p.bidx[i];
            if (aslot) aslot[i] = p.aidx_slot[i];
        }
        g = SmallTables{cdeg, cslot, svar, vdeg, vbase, vid, vslot, bidx, aidx, aslot};
        // quantiser tables behind the int tables
        float* s_thr = reinterpret_cast<float*>(w);
        if (QUANT) {
            const int nthr = p.n_quant * p.nth, nlut = p.n_quant << p.bc;
            for (int i = tid; i < nthr; i += kSmallThreads) s_thr[i] = p.thr[i];
            for (int i = tid; i < nlut; i += kSmallThreads) s_thr[nthr + i] = p.lut[i];
        }
    }
    const float* const s_thr = reinterpret_cast<const float*>(tab + 2 * nc + 2 * E + 3 * n + (p.bidx ? E : 0) + (p.aidx ? n : 0) + (p.aidx_slot ? E : 0));
    const float* const s_lut = s_thr + p.n_quant * p.nth;
    __syncthreads();

    // ---- frame of this thread ----
    // workspace mode (llrT in, packed decisions out): frames in the order of the packed decision words, so that a
    // warp holds bit `lane` of word `word`;  row mode (the caller's row-major buffers in and out): thread t of the
    // CTA owns frame cta_base + t
    const bool rows = p.llr_rows != nullptr;
    const int group = wid / V, v = wid % V;
    const int64_t cta_base = (int64_t)blockIdx.x * kSmallThreads;
    const int64_t gbase = cta_base + (int64_t)group * 32 * V;
    const int64_t f = rows ? cta_base + tid : gbase + (int64_t)lane * V + v;
    const int64_t word = (gbase / (32 * V)) * V + v;
    const bool active = f < p.B;
    constexpr int S = kSmallThreads;   // column stride
    const int cta_frames = (int)min((int64_t)kSmallThreads, p.B - cta_base);   // row mode: frames of this CTA (> 0)
    const int q128 = kSmallThreads / n, r128 = kSmallThreads % n;               // element e -> (frame e / n, variable e % n)

    if (rows) {
        // the CTA's LLR rows are one contiguous block: coalesced loads, transposed into the per-frame columns
        const Real* __restrict__ src = static_cast<const Real*>(p.llr_rows) + cta_base * n;
        const int total = cta_frames * n;
        int fl = tid / n, j = tid % n;
        for (int e = tid; e < total; e += kSmallThreads) {
            (llr - tid)[j * S + fl] = src[e];
            j += r128;
            fl += q128;
            if (j >= n) { j -= n; ++fl; }
        }
        __syncthreads();
    }

    int it_done = p.T;
    bool ok = false;
    uint64_t bits64_out = 0;
    if (active) {
        if (!rows) {
            const Real* __restrict__ gl = static_cast<const Real*>(p.llrT) + f;
            for (int j = 0; j < n; ++j) llr[j * S] = gl[(int64_t)j * p.Bp];
        }
        for (int s = 0; s < E; ++s) msg[s * S] = llr[g.svar[s] * S];   // ldpc_decoder.py:84-87
        const bool has_beta = p.beta != nullptr;
        const bool has_alpha = p.alpha != nullptr && KIND != SMALL_OFFSET;
        // hard decisions of the frame: in two registers while n <= 64, else in the frame's shared-memory words
        const bool reg_bits = nw <= 2;
        uint64_t bits64 = 0;
        for (int t = 0; t < p.T; ++t) {
            const Real* __restrict__ beta_t = has_beta ? static_cast<const Real*>(p.beta) + (size_t)t * p.n_beta : nullptr;
            const Real* __restrict__ alpha_t = p.alpha ? static_cast<const Real*>(p.alpha) + (size_t)t * p.n_alpha : nullptr;
            Quantizer<NTH> qz;
            const float* lutq = s_lut;
            if (QUANT) {
                const int q = __ldg(p.q_of_iter + t);
                qz.load(s_thr + q * p.nth, p.nth, __ldg(p.mono + q) != 0);
                lutq = s_lut + (q << p.bc);
            }
            const bool last = t == p.T - 1;
            // ---- one check node.  DC > 0: compile-time degree, inputs held in registers (read once);
            //      DC == 0: run-time degree, inputs read again for the output phase ----
            auto cn_node = [&](auto dc_tag, const int deg, const int s0) {
                constexpr int DC = decltype(dc_tag)::value;
                const int D = DC > 0 ? DC : deg;
                const Real* const in = msg + s0 * S;
                Real* const out = msg + s0 * S;    // in place: an output is written after its own input was read
                Real x[DC > 0 ? DC : 1];
                MinState<Real, false> st;
                st.init();
#pragma unroll
                for (int k = 0; k < D; ++k) {
                    const Real xk = in[k * S];
                    if constexpr (DC > 0) x[k] = xk;
                    st.push(xk, k);
                }
                if (D == 1) st.m2 = st.m1;   // ldpc_decoder.py:112-113
                auto input = [&](int k) -> Real {
                    if constexpr (DC > 0) return x[k];
                    else return in[k * S];
                };
                if constexpr (KIND == SMALL_OFFSET) {
                    Real beta_check = Real(0);
                    if (has_beta && !p.beta_per_edge) beta_check = beta_t[g.bidx ? g.bidx[s0] : 0];
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const Real xk = input(k);
                        const bool is_min = Arith<Real>::abs(xk) == st.m1;
                        const bool zero_others = D > 1 && (st.m2 == Real(0) || (st.m1 == Real(0) && !is_min));
                        const Real beta = (has_beta && p.beta_per_edge) ? beta_t[g.bidx[s0 + k]] : beta_check;
                        const Real alpha = alpha_t ? alpha_t[g.aslot ? g.aslot[s0 + k] : 0] : Real(0);
                        out[k * S] = offset_value<Real>(is_min ? st.m2 : st.m1, beta, has_beta, alpha, alpha_t != nullptr,
                                                        st.par ^ Arith<Real>::hi(xk), zero_others);
                    }
                } else if (!p.beta_per_edge) {
                    const Real beta = has_beta ? beta_t[g.bidx ? g.bidx[s0] : 0] : Real(1);
                    CheckOut<Real, QUANT> co;
                    co.prepare(st.m1, st.m2, st.par, beta, has_beta, qz, p.bc);
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const Real xk = input(k);
                        const auto o = co.emit(Arith<Real>::abs(xk) == st.m1, Arith<Real>::hi(xk));
                        if constexpr (QUANT) out[k * S] = (Real)lutq[o];
                        else out[k * S] = o;
                    }
                } else {
#pragma unroll
                    for (int k = 0; k < D; ++k) {
                        const Real xk = input(k);
                        const Real raw = (Arith<Real>::abs(xk) == st.m1) ? st.m2 : st.m1;
                        const auto o = cn_emit<Real, QUANT, NTH>(raw, beta_t[g.bidx[s0 + k]], st.par ^ Arith<Real>::hi(xk), qz, p.bc);
                        if constexpr (QUANT) out[k * S] = (Real)lutq[o];
                        else out[k * S] = o;
                    }
                }
            };
            for (int c = 0; c < nc; ++c) {
                const int deg = g.cdeg[c], s0 = g.cslot[c];
                switch (deg) {
                    case 1: cn_node(std::integral_constant<int, 1>{}, 1, s0); break;
                    case 2: cn_node(std::integral_constant<int, 2>{}, 2, s0); break;
                    case 3: cn_node(std::integral_constant<int, 3>{}, 3, s0); break;
                    case 4: cn_node(std::integral_constant<int, 4>{}, 4, s0); break;
                    case 5: cn_node(std::integral_constant<int, 5>{}, 5, s0); break;
                    case 6: cn_node(std::integral_constant<int, 6>{}, 6, s0); break;
                    case 7: cn_node(std::integral_constant<int, 7>{}, 7, s0); break;
                    case 8: cn_node(std::integral_constant<int, 8>{}, 8, s0); break;
                    default: cn_node(std::integral_constant<int, 0>{}, deg, s0); break;
                }
            }
            // ---- one variable node: posterior, hard decision, outgoing messages.  DV > 0: inputs in registers ----
            if (reg_bits) bits64 = 0;
            else
                for (int w = 0; w < nw; ++w) hb[w * S] = 0u;
            auto vn_node = [&](auto dv_tag, const int dvr, const int lb, const int pos) {
                constexpr int DV = decltype(dv_tag)::value;
                const int j = g.vid[pos];
                const Real L = llr[j * S];
                Real post;
                if constexpr (DV >= 0) {
                    constexpr int D1 = DV > 0 ? DV : 1;
                    int off[D1];
                    Real c[D1];
#pragma unroll
                    for (int i = 0; i < DV; ++i) {
                        off[i] = g.vslot[lb + i] * S;
                        c[i] = msg[off[i]];
                    }
                    const Real tot = LibSum<Real>::template stat<DV>([&](int i) { return c[i]; });
                    post = DV > 0 ? Arith<Real>::add(L, tot) : L;
                    if (!last) {   // the v2c update of iteration T-1 is dead
                        Real alpha = Real(1);
                        if (has_alpha) alpha = alpha_t[g.aidx ? g.aidx[pos] : 0];
#pragma unroll
                        for (int d = 0; d < DV; ++d) {
                            Real sd = LibSum<Real>::template stat<(DV > 0 ? DV - 1 : 0)>([&](int i) { return c[i < d ? i : i + 1]; });
                            if (has_alpha) sd = Arith<Real>::mul(alpha, sd);
                            msg[off[d]] = Arith<Real>::add(L, sd);
                        }
                    }
                } else {
                    Real c[kSmallMaxDv];   // in place: every input is read before the first output is written
                    for (int i = 0; i < dvr; ++i) c[i] = msg[g.vslot[lb + i] * S];
                    const Real tot = LibSum<Real>::dyn([&](int i) { return c[i]; }, dvr);
                    post = Arith<Real>::add(L, tot);
                    if (!last) {
                        Real alpha = Real(1);
                        if (has_alpha) alpha = alpha_t[g.aidx ? g.aidx[pos] : 0];
                        for (int d = 0; d < dvr; ++d) {
                            Real sd = LibSum<Real>::dyn([&](int i) { return c[i < d ? i : i + 1]; }, dvr - 1);
                            if (has_alpha) sd = Arith<Real>::mul(alpha, sd);
                            msg[g.vslot[lb + d] * S] = Arith<Real>::add(L, sd);
                        }
                    }
                }
                if (want_post) pcol[j * S] = post;   // a frame that stops keeps the posterior of its last iteration
                if (post < Real(0)) {
                    if (reg_bits) bits64 |= 1ull << j;
                    else hb[(j >> 5) * S] |= 1u << (j & 31);
                }
            };
            for (int pos = 0; pos < n; ++pos) {
                const int dv = g.vdeg[pos], lb = g.vbase[pos];
                switch (dv) {
                    case 0: vn_node(std::integral_constant<int, 0>{}, 0, lb, pos); break;
                    case 1: vn_node(std::integral_constant<int, 1>{}, 1, lb, pos); break;
                    case 2: vn_node(std::integral_constant<int, 2>{}, 2, lb, pos); break;
                    case 3: vn_node(std::integral_constant<int, 3>{}, 3, lb, pos); break;
                    case 4: vn_node(std::integral_constant<int, 4>{}, 4, lb, pos); break;
                    case 5: vn_node(std::integral_constant<int, 5>{}, 5, lb, pos); break;
                    case 6: vn_node(std::integral_constant<int, 6>{}, 6, lb, pos); break;
                    case 7: vn_node(std::integral_constant<int, 7>{}, 7, lb, pos); break;
                    case 8: vn_node(std::integral_constant<int, 8>{}, 8, lb, pos); break;
                    default: vn_node(std::integral_constant<int, -1>{}, dv, lb, pos); break;
                }
            }
            // ---- syndrome and early stop (ldpc_decoder.py:141-144) ----
            if (p.early_stop || last) {
                uint32_t unsat = 0;
                for (int c = 0; c < nc; ++c) {
                    const int deg = g.cdeg[c], s0 = g.cslot[c];
                    uint32_t par = 0;
                    if (reg_bits) {
                        for (int k = 0; k < deg; ++k) par ^= (uint32_t)(bits64 >> g.svar[s0 + k]);
                    } else {
                        for (int k = 0; k < deg; ++k) {
                            const int j = g.svar[s0 + k];
                            par ^= hb[(j >> 5) * S] >> (j & 31);
                        }
                    }
                    unsat |= par & 1u;
                }
                if (!unsat) {
                    it_done = t + 1;
                    ok = true;
                    break;
                }
            }
        }
        bits64_out = bits64;
    }
    if (active && nw <= 2) {   // the output code reads the decision words from shared memory
        hb[0] = (uint32_t)bits64_out;
        if (nw > 1) hb[S] = (uint32_t)(bits64_out >> 32);
    }
    if (rows) {
        // ---- results straight into the caller's row-major buffers, staged through shared memory so that the
        // CTA's contiguous output blocks leave with coalesced stores ----
        if (p.iters && active) p.iters[f] = it_done;
        if (p.success && active) p.success[f] = ok ? 1 : 0;
        if (p.packed_rows && active)   // the frame's decision words ARE its packed row
            for (int w = 0; w < nw; ++w) p.packed_rows[f * nw + w] = hb[w * S];
        __syncthreads();   // every frame's message columns are dead: the region becomes the byte stage of the decisions
        uint8_t* const stage = small_smem;
        if (p.bits_rows && active)
            for (int j = 0; j < n; ++j) stage[tid * n + j] = (uint8_t)((hb[(j >> 5) * S] >> (j & 31)) & 1u);
        __syncthreads();
        const int total = cta_frames * n;
        if (p.bits_rows) {
            uint8_t* __restrict__ dst = p.bits_rows + cta_base * n;   // cta_base * n is a multiple of 128
            const int words = total >> 2;
            for (int e = tid; e < words; e += kSmallThreads)
                reinterpret_cast<uint32_t*>(dst)[e] = reinterpret_cast<const uint32_t*>(stage)[e];
            for (int e = (words << 2) + tid; e < total; e += kSmallThreads) dst[e] = stage[e];
        }
        if (p.post_rows) {
            Real* __restrict__ dst = static_cast<Real*>(p.post_rows) + cta_base * n;
            int fl = tid / n, j = tid % n;
            for (int e = tid; e < total; e += kSmallThreads) {
                dst[e] = (pcol - tid)[j * S + fl];
                j += r128;
                fl += q128;
                if (j >= n) { j -= n; ++fl; }
            }
        }
        return;
    }
    __syncwarp();
    // ---- results in the workspace layout ----
    for (int j = 0; j < n; ++j) {
        const bool bit = active && ((hb[(j >> 5) * S] >> (j & 31)) & 1u);
        const uint32_t w = __ballot_sync(0xffffffffu, bit);
        if (lane == 0) p.hardw[(int64_t)j * p.Wn + word] = w;
    }
    p.iters[f] = it_done;
    p.success[f] = ok ? 1 : 0;
    p.done[f] = 1;
    if (p.postT && active) {
        Real* __restrict__ gp = static_cast<Real*>(p.postT) + f;
        for (int j = 0; j < n; ++j) gp[(int64_t)j * p.Bp] = pcol[j * S];
    }
}

size_t small_smem_bytes(int dtype, const SmallLaunch& p) {
    const size_t rsz = dtype == 0 ? 4 : 8;
    const size_t nw = (size_t)(p.n + 31) / 32;
    const bool want_post = p.post_rows != nullptr || p.postT != nullptr;
    size_t b = ((size_t)(p.E + p.n + (want_post ? p.n : 0)) * rsz + nw * 4) * kSmallThreads;
    b += 4 * ((size_t)2 * p.n_checks + 2 * (size_t)p.E + 3 * (size_t)p.n + (p.bidx ? p.E : 0) + (p.aidx ? p.n : 0) + (p.aidx_slot ? p.E : 0));
    if (p.bc) b += 4 * ((size_t)p.n_quant * p.nth + ((size_t)p.n_quant << p.bc));
    return b;
}

template <typename Real, int KIND, int NTH>
cudaError_t launch_small_t(const SmallLaunch& p, size_t smem, cudaStream_t stream) {
    cudaError_t e = cudaFuncSetAttribute(small_decode_kernel<Real, KIND, NTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    small_decode_kernel<Real, KIND, NTH><<<(unsigned)(p.Bp / kSmallThreads), kSmallThreads, smem, stream>>>(p);
    return cudaGetLastError();
}

}  // namespace

// Per-frame shared-memory footprint decides: the on-chip decode pays while at least ~256 frames fit one SM.
bool small_decode_fits(int dtype, const SmallLaunch& p) {
    if (p.E <= 0 || p.n <= 0) return false;
    const size_t rsz = dtype == 0 ? 4 : 8;
    // (the posterior columns are counted whether or not a call wants them: which decodes go on chip must not depend on it)
    const size_t per_frame = (size_t)(p.E + 2 * p.n) * rsz + 4 * (size_t)((p.n + 31) / 32);
    // (row mode stages the decisions as n bytes per frame in the region of the message columns)
    return per_frame <= 896 && p.max_dv <= kSmallMaxDv && (size_t)p.E * rsz >= (size_t)p.n &&
           small_smem_bytes(dtype, p) <= (size_t)160 * 1024;
}

cudaError_t launch_small_decode(int dtype, const SmallLaunch& p, cudaStream_t stream) {
    if (p.Bp % kSmallThreads != 0) return cudaErrorInvalidValue;
    const size_t smem = small_smem_bytes(dtype, p);
    if (p.check_rule == 1) {
        return dtype == 0 ? launch_small_t<float, SMALL_OFFSET, 0>(p, smem, stream) : launch_small_t<double, SMALL_OFFSET, 0>(p, smem, stream);
    }
    if (p.bc) {
        if (dtype != 0) return cudaErrorInvalidValue;
        // register-resident thresholds need every quantiser's table to be non-decreasing
        if (p.all_mono && p.nth <= 4) return launch_small_t<float, SMALL_QUANT, 4>(p, smem, stream);
        if (p.all_mono && p.nth <= 8) return launch_small_t<float, SMALL_QUANT, 8>(p, smem, stream);
        return launch_small_t<float, SMALL_QUANT, 0>(p, smem, stream);
    }
    return dtype == 0 ? launch_small_t<float, SMALL_NORMALIZED, 0>(p, smem, stream) : launch_small_t<double, SMALL_NORMALIZED, 0>(p, smem, stream);
}

}  // namespace ldpc
