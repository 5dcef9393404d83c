// Everything around the two half iterations: syndrome on packed decisions, early-stop bookkeeping, layout
// conversion, Philox AWGN, error counting, frame-compaction bookkeeping.
#include "ldpc_kernel_common.cuh"

namespace ldpc {

namespace {

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

// postT [n][Bp] -> post [B][n] (row map[f] when a frame map is given): the mirror image of pack_kernel,
// 64 variables x 64 frames per CTA, two elements per lane on both sides.
template <typename Real, bool PAIR>
__global__ void __launch_bounds__(256) unpack_post_kernel(const Real* __restrict__ postT, Real* __restrict__ post, int64_t B,
                                                           int64_t Bp, int32_t n, const int32_t* __restrict__ map) {
    __shared__ Real tile[64][65];   // [frame][variable]
    const int64_t f_base = (int64_t)blockIdx.x * 64;
    const int32_t j_base = blockIdx.y * 64;
    const int lane = threadIdx.x & 31, wy = threadIdx.x >> 5;
    for (int c = wy; c < 64; c += 8) {
        const int32_t j = j_base + c;
        const int64_t f = f_base + 2 * lane;   // Bp is a multiple of 128
        Real a = Real(0), b = Real(0);
        if (j < n && f < Bp) {
            const Pack<Real, 2> v = *reinterpret_cast<const Pack<Real, 2>*>(postT + (int64_t)j * Bp + f);
            a = v.v[0];
            b = v.v[1];
        }
        tile[2 * lane][c] = a;
        tile[2 * lane + 1][c] = b;
    }
    __syncthreads();
    for (int r = wy; r < 64; r += 8) {
        const int64_t f = f_base + r;
        if (f >= B) continue;
        const int32_t j = j_base + 2 * lane;
        Real* row = post + (map ? (int64_t)map[f] : f) * n + j;
        if (PAIR) {
            if (j < n) {
                Pack<Real, 2> v;
                v.v[0] = tile[r][2 * lane];
                v.v[1] = tile[r][2 * lane + 1];
                *reinterpret_cast<Pack<Real, 2>*>(row) = v;
            }
        } else {
            if (j < n) row[0] = tile[r][2 * lane];
            if (j + 1 < n) row[1] = tile[r][2 * lane + 1];
        }
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

// hardw -> packed decision rows [B][row_words] u32: bit (j & 31) of word (j >> 5) of row f = decision of variable j
// of frame f (one eighth of the bytes of the one-byte-per-bit rows: what goes back over PCIe in the packed host
// entry point).  One warp: one hard word (32 frames) x 128 variables.  Lane l loads the words of variables
// 32 i + l (i < 4); a ballot over the lanes of bit b of those words is word i of frame b's row, kept by lane b.
__global__ void pack_rows_kernel(int V, const uint32_t* __restrict__ hardw, int64_t Wn, uint32_t* __restrict__ rows,
                                 int64_t B, int32_t n, int32_t row_words, const int32_t* __restrict__ map) {
    const int lane = threadIdx.x & 31;
    const int64_t w = (int64_t)blockIdx.y * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= Wn) return;
    const int32_t j0 = blockIdx.x * 128;
    uint32_t word[4], mine[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const int32_t j = j0 + 32 * i + lane;
        word[i] = (j < n) ? __ldg(hardw + (int64_t)j * Wn + w) : 0u;
        mine[i] = 0u;
    }
#pragma unroll 4
    for (int b = 0; b < 32; ++b) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const uint32_t bal = __ballot_sync(0xffffffffu, (word[i] >> b) & 1u);
            if (lane == b) mine[i] = bal;
        }
    }
    const int64_t f = wordbit_to_frame(w, lane, V);
    if (f >= B) return;
    uint32_t* row = rows + (map ? (int64_t)map[f] : f) * row_words + blockIdx.x * 4;
#pragma unroll
    for (int i = 0; i < 4; ++i)
        if (blockIdx.x * 4 + i < row_words) row[i] = mine[i];
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

// four standard normal samples of (variable group jg, global frame gf): Box-Muller on one Philox block
__device__ __forceinline__ void awgn_normals(uint32_t jg, uint64_t gf, uint64_t seed, float (&z)[4]) {
    uint32_t r[4];
    philox4x32_10(jg, (uint32_t)gf, (uint32_t)(gf >> 32), 0x4c445043u, (uint32_t)seed, (uint32_t)(seed >> 32), r);
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

// Row-major output [B][n] (ldpc_awgn_llr): consecutive lanes own consecutive variable groups, so a warp writes
// 512 contiguous bytes of one frame's row; frames stride over gridDim.y.
__global__ void __launch_bounds__(128) awgn_rows_kernel(float* __restrict__ out, int32_t n, int64_t B, uint64_t frame0,
                                                        uint64_t seed, float sigma, float inv_sigma2_x2, float llr_sign,
                                                        const uint8_t* __restrict__ codeword) {
    const int32_t jg = blockIdx.x * blockDim.x + threadIdx.x;
    const int32_t j0 = 4 * jg;
    if (j0 >= n) return;
    float sym[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const float cw = (codeword && j0 + i < n) ? (float)codeword[j0 + i] : 0.f;
        sym[i] = llr_sign * (1.f - 2.f * cw);
    }
    const bool vec = (n % 4 == 0) && ((reinterpret_cast<uintptr_t>(out) & 15u) == 0);
    for (int64_t f = blockIdx.y; f < B; f += gridDim.y) {
        float z[4];
        awgn_normals((uint32_t)jg, frame0 + (uint64_t)f, seed, z);
        float val[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) val[i] = __fmul_rn(__fadd_rn(sym[i], __fmul_rn(sigma, z[i])), inv_sigma2_x2);
        float* row = out + f * n + j0;
        if (vec) {
            *reinterpret_cast<float4*>(row) = make_float4(val[0], val[1], val[2], val[3]);
        } else {
#pragma unroll
            for (int i = 0; i < 4; ++i)
                if (j0 + i < n) row[i] = val[i];
        }
    }
}

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
        float z[4];
        awgn_normals((uint32_t)jg, frame0 + (uint64_t)(f0 + v), seed, z);
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
                                                                   int32_t* __restrict__ total,
                                                                   volatile int32_t* __restrict__ total_host) {
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
    if (threadIdx.x == 0) {
        total[0] = s_carry;
        if (total_host) {   // mapped pinned host word: the checkpoint count reaches the host without a copy-engine hop
            *total_host = s_carry;
            __threadfence_system();
        }
    }
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

}  // namespace

// ---------------------------------------------------------------------------------------------
// Launchers
// ---------------------------------------------------------------------------------------------
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

cudaError_t launch_pack_rows(int V, const uint32_t* hardw, int64_t Wn, uint32_t* rows, int64_t B, int32_t n,
                             const int32_t* map, cudaStream_t stream) {
    const int warps = 8;
    dim3 grid((unsigned)((n + 127) / 128), (unsigned)((Wn + warps - 1) / warps));
    pack_rows_kernel<<<grid, warps * 32, 0, stream>>>(V, hardw, Wn, rows, B, n, (n + 31) / 32, map);
    return cudaGetLastError();
}

cudaError_t launch_unpack_post(int dtype, const void* postT, void* post, int64_t B, int64_t Bp, int32_t n,
                               const int32_t* map, cudaStream_t stream) {
    dim3 grid((unsigned)((B + 63) / 64), (unsigned)((n + 63) / 64));
    const bool pair = (n % 2 == 0) && (reinterpret_cast<uintptr_t>(post) % (dtype == 0 ? 8 : 16) == 0);
    if (dtype == 0) {
        if (pair) unpack_post_kernel<float, true><<<grid, 256, 0, stream>>>(static_cast<const float*>(postT), static_cast<float*>(post), B, Bp, n, map);
        else unpack_post_kernel<float, false><<<grid, 256, 0, stream>>>(static_cast<const float*>(postT), static_cast<float*>(post), B, Bp, n, map);
    } else {
        if (pair) unpack_post_kernel<double, true><<<grid, 256, 0, stream>>>(static_cast<const double*>(postT), static_cast<double*>(post), B, Bp, n, map);
        else unpack_post_kernel<double, false><<<grid, 256, 0, stream>>>(static_cast<const double*>(postT), static_cast<double*>(post), B, Bp, n, map);
    }
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
    if (row_major) {
        const int ng = (n + 3) / 4;
        dim3 rgrid((unsigned)((ng + 127) / 128), (unsigned)(B < 32768 ? B : 32768));
        awgn_rows_kernel<<<rgrid, 128, 0, stream>>>(static_cast<float*>(out), n, B, frame0, seed, sigma, k, sgn, codeword);
    } else if (dtype == 0) awgn_kernel<float, false><<<grid, threads, 0, stream>>>(out, n, B, Bp, frame0, seed, sigma, k, sgn, codeword);
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

cudaError_t launch_pending_scan(const uint8_t* done, int64_t Bp, int32_t* counts, int32_t* total, int32_t* total_host,
                                cudaStream_t stream) {
    const int nb = (int)((Bp + kScanBlock - 1) / kScanBlock);
    pending_count_kernel<<<nb, kScanBlock, 0, stream>>>(done, Bp, counts);
    pending_scan_kernel<<<1, kScanBlock, 0, stream>>>(counts, nb, total, total_host);
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
