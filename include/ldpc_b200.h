/*
 * ldpc_b200.h -- C ABI of the B200-native batched LDPC flooding min-sum / RCQ decoding engine.
 *
 * The reference (Lalwaniamisha789/Implementation-of-Neural-LDPC-Decoders-...) is pure Python and
 * has NO FFI / plugin boundary of its own: its boundary is the Python
 * constructor + decode()/forward() surface of five classes.  This header is therefore the boundary
 * a maintainer would bind *underneath* those classes (ctypes stub in INTEGRATION.md).  Each entry
 * point names the reference code it replaces (file:line, relative to the reference root).
 *
 * Conventions: plain C types only (no torch / CUDA types in signatures; a CUDA stream travels as
 * void*), every function returns 0 on success or an LDPC_ERR_* code and never throws;
 * ldpc_last_error() returns a thread-local description of the last failure.  All buffers are
 * caller-allocated.  Handles are independent objects with no shared mutable state, so different
 * handles may be driven from different host threads concurrently (the reference runs one decoder
 * object per thread, simulation_framework.py:192-198).  There is no CPU fallback: without a CUDA
 * device every compute entry point fails with LDPC_ERR_CUDA.
 */
#ifndef LDPC_B200_H
#define LDPC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LDPC_B200_VERSION 104

enum {
    LDPC_OK = 0,
    LDPC_ERR_INVALID = 1,     /* bad argument                                  */
    LDPC_ERR_CUDA = 2,        /* CUDA runtime failure (incl. no device)        */
    LDPC_ERR_NOMEM = 3,       /* host or device allocation failed              */
    LDPC_ERR_UNSUPPORTED = 4  /* valid request this build does not implement   */
};

enum { LDPC_F32 = 0, LDPC_F64 = 1 };
enum { LDPC_RULE_NORMALIZED = 0, LDPC_RULE_OFFSET = 1 };
enum { LDPC_SCHEDULE_FLOODING = 0, LDPC_SCHEDULE_LAYERED = 1 };

typedef struct ldpc_graph ldpc_graph;     /* Tanner graph re-laid-out on one device           */
typedef struct ldpc_decoder ldpc_decoder; /* graph + weights + quantisers + device workspace  */

int ldpc_version(void);
const char *ldpc_last_error(void);
int ldpc_device_count(int *count);

/* Pinned host memory for the host-buffer entry points (full PCIe rate needs page-locked memory). */
int ldpc_host_alloc(void **ptr, int64_t bytes);
int ldpc_host_free(void *ptr);

/* ---------------------------------------------------------------------------------------------
 * Graph.  Replaces LDPCCode's dense H and the repeated np.where / torch.where neighbour scans
 * (ldpc_decoder.py:26-54, :85, :92, :124, :136; neural_2d_decoder.py:155, :162, :195).
 *   check_ptr[m+1], check_var[E]: CSR by check; inside a check the variables MUST be in ascending
 *   index order (that is the order np.where yields and the order the min-sum argmin tie-break and
 *   all sums follow).  Only H == 1 entries are edges.  Checks/variables of degree 0 are allowed.
 * The library derives the variable-side adjacency (ascending check index), the degree classes and
 * the degree-sorted slot layout itself and uploads them to `device`.
 * ------------------------------------------------------------------------------------------- */
int ldpc_graph_create(int device, int32_t n, int32_t m, const int64_t *check_ptr,
                      const int32_t *check_var, ldpc_graph **out);
int ldpc_graph_destroy(ldpc_graph *g);

enum {
    LDPC_GRAPH_N = 0,
    LDPC_GRAPH_M = 1,
    LDPC_GRAPH_E = 2,
    LDPC_GRAPH_CHECK_CLASSES = 3, /* number of distinct non-zero check degrees     */
    LDPC_GRAPH_VAR_CLASSES = 4,   /* number of distinct variable degrees (incl. 0) */
    LDPC_GRAPH_MAX_DC = 5,
    LDPC_GRAPH_MAX_DV = 6,
    LDPC_GRAPH_DEVICE = 7,
    LDPC_GRAPH_LAYER_LEVELS = 8,  /* dependency levels of the checks in index order (layered schedule)        */
    LDPC_GRAPH_LAYER_PIPED = 9    /* 1: the sequential layered walk runs software-pipelined (check degree <= 8) */
};
int ldpc_graph_query(const ldpc_graph *g, int what, int64_t *value);
/* Test hook: message-slot id of every edge (check-major edge order in, slot out). */
int ldpc_graph_slot_of_edge(const ldpc_graph *g, int32_t *slot_of_edge /* [E] */);

/* ---------------------------------------------------------------------------------------------
 * Decoder.  One configuration struct covers the five reference decoders:
 *
 *   BasicMinSumDecoder        (ldpc_decoder.py:56-153)        dtype F64, beta = {factor}, T = code.max_iterations
 *   NeuralMinSumDecoder       (neural_minsum_decoder.py:19-150) beta[T][E], beta_index[e] = e
 *   Neural2DMinSumDecoder     (neural_2d_decoder.py:16-225)    types 1-4 through beta_index / alpha_index
 *   RCQMinSumDecoder flooding (rcq_decoder.py:123-279)         bc > 0, no beta / alpha
 *   WeightedRCQDecoder        (rcq_decoder.py:352-597)         bc > 0 plus beta / alpha
 *   NeuralOffsetMinSumDecoder   (neural_minsum_decoder.py:152-286) check_rule OFFSET, beta[T][E]
 *   Neural2DOffsetMinSumDecoder (neural_2d_decoder.py:227-434)     check_rule OFFSET, types 1-4
 *   RCQMinSumDecoder layered    (rcq_decoder.py:281-350)           schedule LAYERED: posteriors updated in
 *        place, checks in index order, exactly as the reference executes it (its "subtract the previous
 *        C2V" step subtracts zero on every graph with more than one non-empty check; a graph with a
 *        single non-empty check is rejected with LDPC_ERR_UNSUPPORTED).  No beta / alpha.
 *
 * Arithmetic (verified against the reference):
 *   check node   raw_k = min2 if k == first-argmin else min1;  c2v = fl(fl(beta*raw) * sp)
 *   quantised    code = (x < 0) << (bc-1) | last j with |x| >= float(thr_j);  c2v = +-float(thr_idx)
 *   variable     v2c = fl(llr + fl(alpha * S(others)));  posterior = fl(llr + S(all))  (no alpha)
 *   S()          torch.sum order for F32, np.sum (pairwise) order for F64
 *   stop         bits = posterior < 0; all parity checks zero -> iterations = t+1, success
 * ------------------------------------------------------------------------------------------- */
typedef struct ldpc_decoder_config {
    int32_t struct_size;     /* = sizeof(ldpc_decoder_config)                                    */
    int32_t dtype;           /* LDPC_F32 | LDPC_F64 (F64 only without quantiser)                 */
    int32_t max_iterations;  /* T >= 1                                                           */
    int32_t early_stop;      /* 1: reference behaviour; 0: always run T iterations (benchmarks)  */
    int32_t n_beta;          /* columns of beta; 0 -> beta == 1                                  */
    int32_t n_alpha;         /* columns of alpha; 0 -> no alpha multiply (alpha == 1)            */
    int32_t bc;              /* 0: float c2v; 2..8: c2v stored as bc-bit sign-magnitude codes    */
    int32_t n_quantizers;    /* Q                                                                */
    int32_t check_rule;      /* LDPC_RULE_NORMALIZED: c2v = beta*raw*sp;  LDPC_RULE_OFFSET:          */
                             /*   c2v = sp*(relu(raw - beta) - alpha[variable]) and unweighted VN sums */
    int32_t schedule;        /* LDPC_SCHEDULE_FLOODING | LDPC_SCHEDULE_LAYERED (RCQ only, see below) */
    const int32_t *beta_index;        /* [E] column of beta per edge (check-major); NULL -> 0    */
    const void *beta;                 /* [T][n_beta] of dtype                                    */
    const int32_t *alpha_index;       /* [n] column of alpha per variable; NULL -> 0             */
    const void *alpha;                /* [T][n_alpha] of dtype                                   */
    const float *thresholds;          /* [Q][2^(bc-1)] thresholds rounded to float32             */
    const int32_t *quantizer_of_iter; /* [T] quantiser used by iteration t (rcq_decoder.py:156-167) */
} ldpc_decoder_config;

int ldpc_decoder_create(ldpc_graph *g, const ldpc_decoder_config *cfg, ldpc_decoder **out);
/* Replace the weight tables (same shapes as at creation); host pointers. */
int ldpc_decoder_set_weights(ldpc_decoder *d, const void *beta, const void *alpha);
int ldpc_decoder_destroy(ldpc_decoder *d);
/* Pre-size the device workspace for batches of up to `frames` frames (otherwise grown on demand). */
int ldpc_decoder_reserve(ldpc_decoder *d, int64_t frames);

/*
 * Decode B frames.  Replaces decode()/forward() of the classes above, batched:
 *   llr      [B][n] of dtype, row-major (one reference call per row)
 *   bits     [B][n] uint8   hard decisions (posterior < 0)
 *   posterior[B][n] of dtype, may be NULL  (forward() returns it, decode() does not)
 *   iterations[B] int32, success[B] uint8 (either may be NULL)
 * _device: all pointers are device pointers on the graph's device; work is enqueued on `stream`
 *          (a cudaStream_t, NULL = default stream).  With early_stop the call synchronises the stream at
          a few checkpoints (to read the number of running frames: an all-stopped batch ends at once,
          a mostly-stopped one is compacted); the tail of the work is left enqueued, NOT synchronised.
 * _host:   all pointers are host pointers (pinned for full speed); the call copies in, decodes and
 *          copies out through an internal chunked double-buffered pipeline and returns when the
 *          outputs are valid.
 * Which kernels run is the library's business and never changes a result: codes of a few hundred bytes of state per
 * frame decode in ONE launch with their messages in shared memory (ldpc_small.cu); float32 batches of at most a few
 * waves of thread blocks decode one block per frame, messages in that block's shared memory (ldpc_resident.cu) -- both
 * only enqueue, with or without early_stop; everything else runs one kernel per half iteration with the messages in
 * HBM (ldpc_cn.cu / ldpc_vn.cu).  ldpc_decoder_profile_read says which path the calls took.
 */
int ldpc_decode_device(ldpc_decoder *d, const void *llr, int64_t B, uint8_t *bits, void *posterior,
                       int32_t *iterations, uint8_t *success, void *stream);
int ldpc_decode_host(ldpc_decoder *d, const void *llr, int64_t B, uint8_t *bits, void *posterior,
                     int32_t *iterations, uint8_t *success);

/*
 * The same two calls with the hard decisions as PACKED rows (additive: the reference returns one integer per bit):
 *   bits_packed [B][ceil(n/32)] uint32 -- bit (j & 31) of word (j >> 5) of row f is the decision of variable j of
 *   frame f (numpy: np.unpackbits(rows.view(np.uint8), axis=1, bitorder="little")[:, :n]).  One eighth of the bytes of
 *   `bits`: the decisions leave the device as the bit-packed words the kernels already hold, which is what matters
 *   where the host link carries every output (ldpc_decode_host_packed).
 */
int ldpc_decode_device_packed(ldpc_decoder *d, const void *llr, int64_t B, uint32_t *bits_packed, void *posterior,
                              int32_t *iterations, uint8_t *success, void *stream);
int ldpc_decode_host_packed(ldpc_decoder *d, const void *llr, int64_t B, uint32_t *bits_packed, void *posterior,
                            int32_t *iterations, uint8_t *success);

/* Test hook: the chunk plan ldpc_decode_host uses for a batch of B frames (frames per chunk, in order).
 * chunk <= 0 selects the default of calls without posteriors (4096; with posteriors the default is 2048); frames_per_lane is 4 for F32 decoders, 2 for F64. */
int ldpc_host_chunk_plan(int64_t B, int64_t chunk, int32_t frames_per_lane, int64_t *frames_out,
                         int32_t max_chunks, int32_t *n_chunks);

/* ---------------------------------------------------------------------------------------------
 * Posterior training.  Replaces the forward + autograd backward of PosteriorJointTrainer.train_epoch
 * (training_framework.py:108-165, which as shipped crashes -- SURVEY appendix C3; the two repairs are recorded in
 * oracle/reference_training_repairs.patch) for the float32 normalised neural min-sum decoders
 * (NeuralMinSumDecoder, Neural2DMinSumDecoder types 1-4), batched over frames.
 *
 * ldpc_train_forward : ldpc_decode_device with every iteration's messages kept in the handle (2*T*E floats per frame).
 * ldpc_train_backward: given d loss / d posterior [B][n] (device, of the frames of the last forward call) writes
 *                      d loss / d beta [T][n_beta] and d loss / d alpha [T][n_alpha] (device float32, either may be
 *                      NULL), i.e. what torch.autograd accumulates into the reference's ParameterDict entries:
 *                      the posterior of a frame is the one of the iteration it stopped at, min1 sends its gradient to
 *                      the first argmin, min2 shares it among ties, sign() and the quantities derived from it carry
 *                      none.  Gradients are summed over the frames with float32 atomics (order not fixed: compare
 *                      with a tolerance).
 * ------------------------------------------------------------------------------------------- */
int ldpc_train_forward(ldpc_decoder *d, const float *llr, int64_t B, uint8_t *bits, float *posterior,
                       int32_t *iterations, uint8_t *success, void *stream);
int ldpc_train_backward(ldpc_decoder *d, const float *grad_posterior, float *grad_beta, float *grad_alpha,
                        void *stream);

/* ---------------------------------------------------------------------------------------------
 * Monte-Carlo leg.  Replaces simulate_awgn_channel (ldpc_decoder.py:286-302) and the per-frame
 * body of LDPSimulator.simulate_single_snr (simulation_framework.py:110-131).
 *
 * Noise: z(frame, variable) is a pure function of (seed, global frame index, variable index):
 * Philox4x32-10, key = seed, counter = (variable/4, frame_lo, frame_hi, 0x4c445043), Box-Muller on
 * the four outputs.  y = s + sigma*z, sigma^2 = 10^(-snr_db/10), llr = 2*y/sigma^2 in float32
 * (then widened for F64 decoders).  s = +1 for codeword bit 0 when llr_sign = +1 (decisions then
 * converge to the all-zero word) and -1 when llr_sign = -1 (the reference's own convention,
 * ldpc_decoder.py:289 -- see SURVEY appendix C1).
 * ------------------------------------------------------------------------------------------- */
/* llr_out: device [B][n] float32 row-major.  codeword: device uint8 [n] or NULL (all-zero). */
int ldpc_awgn_llr(int device, int32_t n, int64_t B, uint64_t frame0, uint64_t seed, float snr_db,
                  int32_t llr_sign, const uint8_t *codeword, float *llr_out, void *stream);

/*
 * One Monte-Carlo round on frames [frame0, frame0+B): generate LLRs straight into the decoder's
 * interleaved layout, decode, compare with `codeword` (device uint8[n] or NULL = all-zero) and
 * ACCUMULATE into counters (device int64[4]):
 *   {frame_errors, bit_errors (of erroneous frames), total_iterations, total_frames}
 * i.e. the four sums of simulation_framework.py:125-131.  Optional per-frame outputs (device,
 * may be NULL): frame_bit_errors[B] int32, frame_iterations[B] int32 -- used for the exact
 * sequential max_errors stop rule.  Enqueued on `stream`, not synchronised.
 */
int ldpc_mc_round(ldpc_decoder *d, float snr_db, int32_t llr_sign, uint64_t seed, uint64_t frame0,
                  int64_t B, const uint8_t *codeword, int64_t *counters, int32_t *frame_bit_errors,
                  int32_t *frame_iterations, void *stream);

/* Stand-alone error counting on decoded bits (device uint8 [B][n]), same counter semantics. */
int ldpc_count_errors(int device, int32_t n, int64_t B, const uint8_t *bits, const uint8_t *codeword,
                      const int32_t *iterations, int64_t *counters, int32_t *frame_bit_errors,
                      void *stream);

/* ---------------------------------------------------------------------------------------------
 * Instrumentation (bench.py / tests).
 * ------------------------------------------------------------------------------------------- */
typedef struct ldpc_profile {
    int64_t launches;         /* kernels of this library launched by the handle since reset        */
    int64_t cn_launches;      /* check-node kernel launches                                        */
    int64_t vn_launches;      /* variable-node kernel launches (incl. the final posterior pass)    */
    double cn_ms;             /* device time of those launches, CUDA events (profiling mode only)  */
    double vn_ms;
    double other_ms;          /* pack / unpack / syndrome / commit / awgn / count                  */
    int64_t frames_padded;    /* Bp of the last call                                               */
    int64_t compactions;      /* times the running frames were gathered into a smaller dense batch */
    int64_t early_exits;      /* decodes that ended before T because every frame had stopped       */
    int64_t graph_replays;    /* launch-bound decodes replayed from a captured CUDA graph          */
    int64_t small_decodes;    /* decodes that ran as ONE on-chip launch (small codes, messages in shared memory) */
    int64_t resident_decodes; /* decodes that ran CTA-resident (one thread block per frame, messages in shared memory) */
} ldpc_profile;
/* mode 0: count launches only (no overhead); mode 1: also bracket every kernel with CUDA events. */
int ldpc_decoder_profile_mode(ldpc_decoder *d, int32_t mode);
int ldpc_decoder_profile_read(ldpc_decoder *d, ldpc_profile *out, int32_t reset);

#ifdef __cplusplus
}
#endif
#endif /* LDPC_B200_H */
