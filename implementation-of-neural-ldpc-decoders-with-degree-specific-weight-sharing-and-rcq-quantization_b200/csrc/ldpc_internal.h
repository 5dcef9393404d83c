// Host-visible declarations shared between the kernel translation unit and the C-ABI layer.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace ldpc {

// One unit of grid work: `count` consecutive nodes of one degree class.
//   check side : `first_slot` = message slot of the first check's first edge; the next check
//                starts `deg` slots later (degree-sorted, check-major slot layout).
//   variable side: `first_node` = position (vpos) of the first variable in degree-sorted order;
//                `first_slot` = offset of its slot list inside vslots; next variable += deg.
struct WorkItem {
    int32_t deg;
    int32_t count;
    int32_t first_node;
    int32_t first_slot;
};

constexpr int kMaxQuantLevels = 128;   // 2^(bc-1) for bc <= 8
constexpr int kMaxLutFloats = 4096;    // Q * 2^bc

struct CnLaunch {
    const void* src;            // v2c [E][Bp], or llrT [n][Bp] for iteration 0
    void* dst;                  // c2v: Real [E][Bp] or uint8 codes [E][Bp]
    const int32_t* row_map;     // iteration 0: variable of each slot; else nullptr
    const int32_t* bidx;        // per-slot column of beta, or nullptr (column 0)
    int beta_per_edge;          // 0: every edge of a check uses the column of its first slot; 1: per edge
    const void* beta_t;         // beta row of this iteration, or nullptr (beta == 1; offset rule: beta == 0)
    const int32_t* aidx_slot;   // offset rule only: per-slot column of alpha (by the edge's variable), or nullptr
    const void* alpha_t;        // offset rule only: alpha row of this iteration, or nullptr (alpha == 0)
    const float* thr;           // device thresholds of this iteration's quantiser [nth]
    int nth;                    // 2^(bc-1), 0 = float messages
    int bc;
    int mono;                   // thresholds non-decreasing
    const uint8_t* done;        // [Bp]
    const WorkItem* items;
    int n_items;
    int items_wide_begin;       // items [begin, end) hold checks of degree 9..64 (sorted by degree)
    int items_wide_end;
    int wide_ring;              // 1: those items run in the bulk-async row-ring kernel
    int64_t Bp;
};

struct VnLaunch {
    const void* c2v;            // Real [E][Bp] or uint8 codes
    void* v2c;                  // Real [E][Bp] (ignored by the final pass)
    const void* llrT;           // Real [n][Bp]
    void* postT;                // Real [n][Bp] or nullptr: posterior rows, refreshed by running frames every iteration
    const int32_t* vslots;      // slot lists, ascending check index inside a variable
    const int32_t* vpos_var;    // variable id at each degree-sorted position
    const int32_t* aidx;        // per-position column of alpha, or nullptr (column 0)
    const void* alpha_t;        // alpha row of this iteration, or nullptr
    const float* lut;           // device LUT [Q][2^bc] (value of every code), or nullptr
    int bc;
    int n_quant;
    int q_now;                  // quantiser of this iteration (non-final)
    uint32_t* hardw;            // [n][Wn] packed hard decisions
    int64_t Wn;                 // Bp / 32
    const uint8_t* done;
    const WorkItem* items;
    int n_items;
    int64_t Bp;
    int final_pass;
    int items_wide_begin;       // items [begin, end) hold variables of degree 9..64 (sorted by degree)
    int items_wide_end;
    int wide_max_deg;           // largest degree among them (rows of the shared-memory stage)
    int wide_stage;             // 1: those items run in vn_wide_kernel
    // Posterior-on-stop pass (final_pass && postT only): post_iter > 0 restricts the pass to the frames that
    // stopped exactly at iteration post_iter (done && iters == post_iter); it writes their posterior row entries
    // from the check->variable messages of that iteration and nothing else.
    const int32_t* iters;       // [Bp]
    int post_iter;
};

struct SynLaunch {
    const uint32_t* hardw;
    int64_t Wn;
    const int32_t* slot_var;
    const WorkItem* items;      // check-side items
    int n_items;
    uint32_t* unsat;            // [Wn] OR-accumulated syndrome words
};

// On-chip decode of a whole frame in one launch (ldpc_small.cu): inputs llrT [n][Bp], results in the workspace layout.
struct SmallLaunch {
    // row mode (llr_rows != nullptr): the caller's row-major buffers, no workspace involved
    const void* llr_rows;       // Real [B][n]
    uint8_t* bits_rows;         // [B][n] or nullptr
    uint32_t* packed_rows;      // [B][ceil(n/32)] or nullptr (bit j & 31 of word j >> 5 = decision of variable j)
    void* post_rows;            // Real [B][n] or nullptr
    // workspace mode
    const void* llrT;           // Real [n][Bp]
    void* postT;                // Real [n][Bp] or nullptr
    uint32_t* hardw;            // [n][Wn]
    int64_t Wn;
    uint8_t* done;              // [Bp] (set to 1 for every frame)
    int32_t* iters;             // [Bp]; row mode: [B] or nullptr
    uint8_t* success;           // [Bp]; row mode: [B] or nullptr
    int64_t B, Bp;              // Bp: B rounded up to a multiple of 128
    int T, early_stop;
    int n, E, n_checks;         // n_checks: non-empty checks = entries of cn_items
    int max_dv;
    const WorkItem* cn_items;   // one check per item (the graph's fine list)
    const WorkItem* vn_items;   // one variable per item, all n of them
    const int32_t* slot_var;
    const int32_t* vslots;
    const int32_t* vpos_var;
    const int32_t* bidx;        // per slot, or nullptr
    int beta_per_edge;
    const void* beta;           // [T][n_beta] or nullptr
    int n_beta;
    const int32_t* aidx;        // per degree-sorted position, or nullptr
    const int32_t* aidx_slot;   // per slot (offset rule), or nullptr
    const void* alpha;          // [T][n_alpha] or nullptr
    int n_alpha;
    int check_rule;             // 0 normalised, 1 offset
    int bc, nth, n_quant;
    const float* thr;           // [Q][nth]
    const float* lut;           // [Q][2^bc]
    const int32_t* q_of_iter;   // device [T]
    const int32_t* mono;        // device [Q]: thresholds of quantiser q are non-decreasing
    int all_mono;
};
bool small_decode_fits(int dtype, const SmallLaunch& p);
cudaError_t launch_small_decode(int dtype, const SmallLaunch& p, cudaStream_t stream);

// CTA-resident decode (ldpc_resident.cu): one thread block per frame, messages in shared memory, row-major I/O.
// Index tables in the kernel's own ("physical") slot order -- see ldpc_graph::Resident in ldpc_api.cu.
constexpr int kResMaxClasses = 24;   // degree classes per side
struct ResidentLaunch {
    const float* llr_rows;      // [B][n]
    uint8_t* bits_rows;         // [B][n] or nullptr
    uint32_t* packed_rows;      // [B][ceil(n/32)] or nullptr
    float* post_rows;           // [B][n] or nullptr
    int32_t* iters;             // [B] or nullptr
    uint8_t* success;           // [B] or nullptr
    int64_t B;
    int T, early_stop;
    int n, E, max_dv;           // E: PHYSICAL message slots (classes padded to whole 32-node tiles)
    int n_cclass, n_vclass;
    const WorkItem* classes;    // [n_cclass] check classes (first_slot: first physical slot), then [n_vclass] variable
                                // classes (first_node: first position, first_slot: offset of the class's lists in vslots)
    const uint16_t* slot_var;   // [E] physical slot -> variable.  Physical slot of edge k of the c-th check of a class:
                                //     first_slot + (c / 32) * 32 * deg + k * 32 + c % 32
    const uint16_t* vslots;     // entry d of the i-th variable of a class (a physical slot) at
                                //     first_slot + (i / 32) * 32 * deg + d * 32 + i % 32
    const uint16_t* vpos_var;   // [n] position -> variable
    const int32_t* bidx;        // [E] beta column per physical slot, or nullptr
    int beta_per_edge;
    const float* beta;
    int n_beta;
    const int32_t* aidx;        // [n] alpha column per position, or nullptr
    const int32_t* aidx_slot;   // [E] alpha column per physical slot (offset rule), or nullptr
    const float* alpha;
    int n_alpha;
    int check_rule;
    int bc, nth, n_quant;
    const float* thr;
    const float* lut;
    const int32_t* q_of_iter;
    const int32_t* mono;
    int all_mono;
    int sm_count;
};
bool resident_decode_fits(const ResidentLaunch& p);
cudaError_t launch_resident_decode(const ResidentLaunch& p, cudaStream_t stream);
int64_t resident_wave_frames(const ResidentLaunch& p);

// One iteration of the training backward pass (ldpc_train.cu); float32, normalised rule, frames in the [rows][Bp] layout.
struct TrainBwd {
    int64_t B, Bp;
    const int32_t* iters;       // [Bp] iterations each frame executed (forward pass)
    const WorkItem* cn_items;
    int n_cn_items;
    const WorkItem* vn_items;
    int n_vn_items;
    const int32_t* slot_var;
    const int32_t* vslots;
    const int32_t* vpos_var;
    const int32_t* bidx;        // per slot or nullptr
    int beta_per_edge;
    const float* beta;          // [T][n_beta] or nullptr
    int n_beta;
    float beta_const;           // beta when there is no table
    const int32_t* aidx;        // per position or nullptr
    const float* alpha;         // [T][n_alpha] or nullptr
    int n_alpha;
    const float* llrT;          // [n][Bp]: inputs of the check nodes of iteration 0 (through slot_var)
    const float* v2c_t;         // [E][Bp] inputs of the check nodes of iteration t (t > 0)
    const float* c2v_t;         // [E][Bp] their outputs
    const float* g_post;        // [n][Bp] gradient of the loss with respect to the posteriors
    float* g_v2c;               // [E][Bp] in: g v2c_{t+1};  out (check side): g v2c_t
    float* g_c2v;               // [E][Bp] g c2v_t
    float* g_beta;              // [beta_parts][T][n_beta] accumulated, or nullptr
    float* g_alpha;             // [alpha_parts][T][n_alpha] accumulated, or nullptr
    // Degree-shared weights have a handful of columns: every warp of the grid adding to the same few addresses
    // serialises in L2 (2.1 TB/s instead of 4.3 for the backward pass of a type-2 decoder).  The sums are therefore
    // spread over `parts` copies (a power of two; block b adds to copy b & (parts - 1)) that a last kernel folds.
    int beta_parts, alpha_parts;
    int T;
    int force_general;          // tuning / test knob: the check side runs its general four-pass form for every check
};
cudaError_t launch_train_fold(const float* parts, float* out, int n_parts, int64_t count, cudaStream_t stream);
cudaError_t launch_train_bwd_vn(const TrainBwd& p, int t, cudaStream_t stream);
cudaError_t launch_train_bwd_cn(const TrainBwd& p, int t, cudaStream_t stream);

// All launchers enqueue on `stream` and return the CUDA launch status.
cudaError_t launch_cn(int dtype, const CnLaunch& p, cudaStream_t stream);
cudaError_t launch_vn(int dtype, const VnLaunch& p, cudaStream_t stream);
// offset min-sum check rule: c2v = sp * (relu(raw - beta) - alpha)
cudaError_t launch_cn_offset(int dtype, const CnLaunch& p, cudaStream_t stream);
// one layered-RCQ iteration over all checks in index order, in place on the posteriors P [n][Bp]
cudaError_t launch_layered_iter(float* P, const int64_t* chk_ptr, const int32_t* chk_var, int32_t m, const float* thr,
                                int nth, int bc, int mono, const uint8_t* done, int64_t Bp, cudaStream_t stream);
// the checks level_chk[0..n_checks) of ONE dependency level of that schedule, concurrently (they share no variable)
cudaError_t launch_layered_level(float* P, const int64_t* chk_ptr, const int32_t* chk_var, const int32_t* level_chk,
                                 int n_checks, const float* thr, int nth, int mono, const uint8_t* done, int64_t Bp,
                                 int max_dc, int staged, cudaStream_t stream);
// the whole iteration in index order, software-pipelined (see layered_pipe_kernel): one record per NON-EMPTY check.
// desc[k] bits 6..0: (distance << 3 | position) of the next reader this edge's new value is forwarded to, 0 = none;
// ahead_mask bit k: input k is copied ahead from its posterior row (nobody forwards it).  Positions >= dc hold 0.
constexpr int kLayerMaxDeg = 8;
struct alignas(16) LayerRec {
    int32_t var[kLayerMaxDeg];
    uint8_t desc[kLayerMaxDeg];
    uint8_t dc;
    uint8_t ahead_mask;
    uint8_t pad[6];
};
static_assert(sizeof(LayerRec) == 48, "three 16-byte loads per check");
cudaError_t launch_layered_pipe(float* P, const LayerRec* recs, int n_checks, const float* thr, int nth, int mono,
                                const uint8_t* done, int64_t Bp, int frames_per_thread, cudaStream_t stream);
int layered_pipe_depth();   // ring depth the descriptors must be built for
// hard decisions (P < 0) of every frame, bit-packed
cudaError_t launch_hard(int dtype, const void* P, uint32_t* hardw, int64_t Wn, int32_t n, int64_t Bp, cudaStream_t stream);
cudaError_t launch_syndrome(const SynLaunch& p, cudaStream_t stream);
// frames with done == 0 whose syndrome word bit is clear become done with iterations = t1;
// also clears `unsat_next`.  V = frames per lane of the decoder's dtype.
cudaError_t launch_commit(int V, const uint32_t* unsat, uint32_t* unsat_next, uint8_t* done,
                          int32_t* iters, uint8_t* success, int32_t t1, int64_t Bp, cudaStream_t stream);
// llr [B][n] row-major -> llrT [n][Bp]; also resets done / iters / success for a new decode.
cudaError_t launch_pack(int dtype, const void* llr, void* llrT, int64_t B, int64_t Bp, int32_t n,
                        uint8_t* done, int32_t* iters, uint8_t* success, int32_t T, cudaStream_t stream);
cudaError_t launch_reset_state(uint8_t* done, int32_t* iters, uint8_t* success, uint32_t* unsat2,
                               int64_t B, int64_t Bp, int32_t T, cudaStream_t stream);
// hardw -> bits [B][n] uint8
// `map` (may be nullptr): output row of local frame f is map[f] (frames of a compacted level)
cudaError_t launch_unpack_bits(int V, const uint32_t* hardw, int64_t Wn, uint8_t* bits, int64_t B,
                               int32_t n, const int32_t* map, cudaStream_t stream);
// hardw -> packed rows [B][ceil(n/32)] u32 (bit j & 31 of word j >> 5 = decision of variable j)
cudaError_t launch_pack_rows(int V, const uint32_t* hardw, int64_t Wn, uint32_t* rows, int64_t B, int32_t n,
                             const int32_t* map, cudaStream_t stream);
// postT [n][Bp] -> post [B][n]
cudaError_t launch_unpack_post(int dtype, const void* postT, void* post, int64_t B, int64_t Bp,
                               int32_t n, const int32_t* map, cudaStream_t stream);
cudaError_t launch_copy_frames(const int32_t* iters_src, const uint8_t* succ_src, int32_t* iters_dst,
                               uint8_t* succ_dst, int64_t B, cudaStream_t stream);
// AWGN LLRs.  row_major != 0: out is float [B][n]; else out is Real [n][Bp] of `dtype`.
cudaError_t launch_awgn(int dtype, int row_major, void* out, int32_t n, int64_t B, int64_t Bp,
                        uint64_t frame0, uint64_t seed, float snr_db, int32_t llr_sign,
                        const uint8_t* codeword, cudaStream_t stream);
// counters += {frame_errors, bit_errors, total_iterations, total_frames}
// `map` (may be nullptr): per-frame outputs of local frame f go to row map[f]; `only_done` (may be nullptr):
// frames with only_done[f] == 0 are left out (they were handed on to a compacted level and are counted there)
cudaError_t launch_count_packed(int V, const uint32_t* hardw, int64_t Wn, int32_t n, int64_t B,
                                const uint8_t* codeword, const int32_t* iters, int64_t* counters,
                                int32_t* frame_bit_errors, int32_t* frame_iters, const int32_t* map,
                                const uint8_t* only_done, int32_t* frame_cnt /* scratch [Wn*32] */, cudaStream_t stream);
// frame compaction bookkeeping: counts[ceil(Bp/1024)] becomes the exclusive scan of running frames per
// 1024-frame block and total[0] their number; then idx[0..total) = the running frames in ascending order
// total_host (may be nullptr): device-visible address of a mapped pinned host word that receives the total as well
cudaError_t launch_pending_scan(const uint8_t* done, int64_t Bp, int32_t* counts, int32_t* total, int32_t* total_host,
                                cudaStream_t stream);
cudaError_t launch_pending_indices(const uint8_t* done, int64_t Bp, const int32_t* offsets, int32_t* idx, cudaStream_t stream);
// dst [rows][Bp_dst] column i = src [rows][Bp_src] column idx[i] (i < count), zero for the pad columns
cudaError_t launch_gather_cols(int dtype, const void* src, int64_t Bp_src, void* dst, int64_t Bp_dst, const int32_t* idx,
                               int64_t count, int64_t rows, cudaStream_t stream);
// iters_dst[map[i]] = iters_src[i] (same for success); map == nullptr: identity
cudaError_t launch_scatter_frames(const int32_t* iters_src, const uint8_t* succ_src, int32_t* iters_dst, uint8_t* succ_dst,
                                  const int32_t* map, int64_t count, cudaStream_t stream);
// out[i] = parent_map ? parent_map[idx[i]] : idx[i]
cudaError_t launch_compose_map(const int32_t* idx, const int32_t* parent_map, int32_t* out, int64_t count, cudaStream_t stream);
cudaError_t launch_count_bits(const uint8_t* bits, int32_t n, int64_t B, const uint8_t* codeword,
                              const int32_t* iters, int64_t* counters, int32_t* frame_bit_errors,
                              cudaStream_t stream);

}  // namespace ldpc
