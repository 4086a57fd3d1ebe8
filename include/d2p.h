/* libd2p - C ABI of the B200-native demo2program hot path.
 *
 * The reference (shaohua0116/demo2program) has no native/FFI layer: its seam
 * is the Python `Model` class driven by trainer.py / evaler.py, and all
 * arithmetic is delegated to TensorFlow-1.3 ops.  Each entry point below
 * replaces the TF op sequence at the cited reference call site; the Python
 * host (demo2program_b200/) binds them with ctypes.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer owned by the caller (the PyTorch
 *    allocator); the library never allocates or frees user-visible memory;
 *  - every call is asynchronous on `stream` (a cudaStream_t passed as void*),
 *    performs no host synchronisation and is CUDA-graph capturable;
 *  - return value 0 = success, negative = d2p_status; d2p_last_error() returns
 *    a thread-local message; nothing throws or aborts;
 *  - sequences are time-major inside the library: [T, R, C] with R = B*k rows,
 *    r = b*k + i (i = demonstration index);
 *  - "saved" buffers carry forward activations to the matching *_bwd call,
 *    "ws" buffers are scratch; query sizes with the *_floats / *_bytes calls;
 *  - gradient outputs of parameters ACCUMULATE (+=) into the flat gradient
 *    buffer, which the caller zeroes once per step.
 */
#ifndef D2P_H_
#define D2P_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    D2P_OK = 0,
    D2P_ERR_ARG = -1,   /* bad argument / shape / workspace size */
    D2P_ERR_CUDA = -2   /* CUDA runtime error (message has the details) */
} d2p_status;

enum { D2P_U8 = 0, D2P_F32 = 1 };
enum { D2P_MAX_CONV_LAYERS = 5 };

const char* d2p_last_error(void);
int d2p_version(void);
/* kernels launched (or captured into a CUDA graph) by this library so far */
long long d2p_launch_count(void);

/* ---- K1: State_Encoder CNN --------------------------------------------------
 * reference models/model_full.py:216-231 (State_Encoder) via ops.conv2d,
 * models/ops.py:27-33: slim.conv2d(3x3, stride 2, SAME, bias) -> lrelu(0.2)
 * -> contrib.layers.batch_norm, x3 (Karel) / x5 (ViZDoom); train-mode batch
 * statistics are per demonstration index (model_full.py:373-376). */
typedef struct {
    const float* w;        /* [3,3,cin,cout] HWIO  (<scope>/Conv/weights) */
    const float* b;        /* [cout]               (<scope>/Conv/biases) */
    const float* gamma;    /* [cout] */
    const float* beta;     /* [cout] */
    float* moving_mean;    /* [cout] in/out */
    float* moving_var;     /* [cout] in/out */
    float* dw;             /* grads, accumulate; may be NULL for fwd-only */
    float* db;
    float* dgamma;
    float* dbeta;
    int cout;
    int _pad;
} d2p_conv_layer;

typedef struct {
    int B, k, T, h, w, d;  /* frames [B,k,T,h,w,d] */
    int frames_dtype;      /* D2P_U8 (as stored) or D2P_F32 (as the reference feeds) */
    int n_layers;
    d2p_conv_layer layers[D2P_MAX_CONV_LAYERS];
} d2p_conv_desc;

size_t d2p_conv_encoder_saved_floats(const d2p_conv_desc* d);
size_t d2p_conv_encoder_ws_bytes(const d2p_conv_desc* d);
int d2p_conv_encoder_feature_dim(const d2p_conv_desc* d);
/* feat: [T, R, F] time-major. training != 0: batch statistics + moving update.
 * The RGB input layer of u8 frames (d = 3, first layer 16 channels: ViZDoom) stages its weights in ONE
 * __constant__ symbol of the library in front of its kernel: calls for such descriptors must not overlap on
 * different streams of the same process (one process per GPU, as everywhere here, is unaffected). */
int d2p_conv_encoder_fwd(const d2p_conv_desc* d, const void* frames, float* feat, float* saved,
                         int training, void* ws, size_t ws_bytes, void* stream);
int d2p_conv_encoder_bwd(const d2p_conv_desc* d, const void* frames, const float* dfeat,
                         const float* saved, int training, void* ws, size_t ws_bytes, void* stream);
/* 1 (default): the Karel geometry (8x8x16 -> 16/32/48 channels) runs as ONE cooperative kernel
 * (conv -> lrelu -> BatchNorm x3, activations in shared memory, per-slice statistics exchanged
 * among the CTAs of a slice); 0: one kernel sequence per layer (always used for ViZDoom). */
int d2p_conv_set_fused(int mode);
/* Per-layer path, layers with 16/32/48 input channels (ViZDoom conv2-5, Karel conv2-3; replaces the
 * same slim.conv2d call, models/ops.py:27-33, and its gradients): bit 0 = forward, bit 1 = input
 * gradient, bit 2 = weight gradient run as tcgen05 implicit-GEMM kernels (bf16x3 split, no im2col
 * buffer) whenever the tensor-core arena is configured (d2p_tc_configure); bit 2 also runs the weight
 * gradient of the u8 RGB input layer (CIN = 3) on the tensor cores.  Bit 4 SET switches the direct
 * RGB-layer forward kernel (fused BatchNorm partial sums) off; bit 5 SET computes the input gradient of
 * 16- / 32-channel inputs per parity class instead of the quad form (one 2x2-tap stride-1 product over
 * the dZ grid that yields the four classes of a 2x2 input block at once).  Default 7; returns the previous
 * mode.  16 = the fp32 CUDA-core per-layer kernels everywhere (the A/B reference of the parity tests). */
int d2p_conv_set_tc(int mode);

/* ---- K2/K3: LSTM over a sequence -------------------------------------------
 * reference models/model_full.py:244-258 (Demo_Encoder), :265-277
 * (SecondPathEncoder), :465-471 (decoder cell loop): BasicLSTMCell, gate order
 * i,j,f,o, forget_bias, dynamic_rnn length masking.
 * X [T,R,In]; W [(In+H),4H] (TF kernel layout); Y [T,R,H] zero past len;
 * gates [T,R,4H] and cells [T,R,H] are saved for the backward. */
enum { D2P_LSTM_INPUT = 1,      /* gates = X*Wx + b for all steps (independent of h0/c0) */
       D2P_LSTM_RECUR = 2,      /* the recurrence over T steps */
       D2P_LSTM_BWD_RECUR = 1,  /* BPTT recurrence + dX, dh0, dc0 */
       D2P_LSTM_BWD_PARAMS = 2, /* dW, db from the dZ left in `gates` by the recurrence phase */
       D2P_LSTM_COMPACT = 8,    /* (fwd and bwd recurrence) 32-CTA grid: every CTA walks all row tiles, so that
                                 * independent recurrences (action / perception / program decoders) share the GPU */
       D2P_LSTM_WIDE = 16,      /* (fwd and bwd recurrence, more than 128 rows) this recurrence has the GPU to itself:
                                 * its rows are split evenly over 4 row tiles x 32 column CTAs (R = 320: 4 x 80 rows,
                                 * 128 CTAs) instead of 128-row tiles (96 CTAs), which shortens the per-step operand
                                 * stream of every CTA; not for recurrences that run beside another persistent kernel */
       D2P_LSTM_BWD_NO_DWX = 4  /* with BWD_PARAMS: leave the input-weight rows dW[0:In] alone (X is not read).
                                 * For a teacher-forced token decoder (X = embedding rows, reference
                                 * models/model_full.py:440-471) the caller forms dWx = E^T * S and
                                 * dE = S * Wx^T from the per-token sums S[v] = sum of dZ rows fed token v
                                 * (d2p_embed_shifted_bwd on dZ): a [V+1]-row product instead of a [T*R]-row one */ };
/* `phases` selects which parts run (3 = all); the parts may be issued on different
 * streams as long as the stream order of the data dependencies is kept. */
int d2p_lstm_seq_fwd(const float* X, int T, int R, int In, int H, const int* len, const float* h0,
                     const float* c0, const float* W, const float* b, float forget_bias, float* Y,
                     float* hT, float* cT, float* gates, float* cells, int phases, void* stream);
size_t d2p_lstm_seq_bwd_ws_bytes(int T, int R, int H);
/* gates is consumed (holds dZ on return). dY/dhT/dcT/dX may be NULL.
 * dh0/dc0 [R,H] are required (they double as the running state grads). */
int d2p_lstm_seq_bwd(const float* X, int T, int R, int In, int H, const int* len, const float* h0,
                     const float* c0, const float* W, const float* Y, float* gates,
                     const float* cells, const float* dY, const float* dhT, const float* dcT,
                     float* dX, float* dW, float* db, float* dh0, float* dc0, void* ws,
                     size_t ws_bytes, int phases, void* stream);

/* ---- decoder inputs / losses -------------------------------------------------
 * reference models/model_full.py:282-296 (Token_Embedding), :446-450 (<s>
 * shift; the out-of-range start id yields a zero row, TF GPU gather),
 * :620-657 (Sequence_Loss). tokens [R,L] int32; X [L,R,E]. */
int d2p_embed_shifted(const float* table, int vocab_rows, int E, const int* tokens, int R, int L,
                      int start_id, float* X, void* stream);
/* ---- scheduled sampling (reference models/model_full.py:59-67, 414-423; trainer.py:278-281) ----
 * seq2seq.ScheduledEmbeddingTrainingHelper: after decoder step t a row feeds, with probability
 * p = 1 - polynomial_decay(1.0 -> 0.1 over decay_steps)(global step), a token drawn from
 * Categorical(logits_t) to step t+1 instead of the ground-truth token.  The decoder therefore runs
 * step by step: d2p_embed_shifted_step (input row of step t from the tokens fed so far), one
 * d2p_lstm_seq_fwd step with the 0/1 lengths of d2p_step_lens, the projection, then
 * d2p_sched_sample_step, which writes fed_tokens[r, t] (and sampled[r, t] = 1 where it drew).
 * The reference's draws are unseeded; here they are d2p_sched_hash(seed, global step, decoder, t, row,
 * 0 | 1) / 2^32 (Bernoulli | inverse-CDF categorical), reproducible and restated by the oracle.
 * adam_state[0] is the device-resident global step; p_override >= 0 replaces the schedule (tests). */
int d2p_step_lens(const int* runlen, int R, int L, int* out /*[L,R]: t < runlen[r]*/, void* stream);
int d2p_embed_shifted_step(const float* table, int vocab_rows, int E, const int* tokens, int R, int L, int t,
                           int start_id, float* out, void* stream);
int d2p_sched_sample_step(const float* logits, int R, int V, const int* gt_tokens, int L, int t,
                          const double* adam_state, int decay_steps, float p_override, unsigned seed,
                          int decoder, int* fed_tokens, int* sampled, void* stream);
unsigned d2p_sched_hash(unsigned seed, unsigned step, unsigned decoder, unsigned t, unsigned r, unsigned which);
size_t d2p_embed_shifted_bwd_ws_bytes(int vocab_rows, int E, int R, int L);
int d2p_embed_shifted_bwd(const float* dX, int vocab_rows, int E, const int* tokens, int R, int L,
                          int start_id, float* dTable, void* ws, size_t ws_bytes, void* stream);
/* w[r] = coef / sum(len over rows of the same decoder instance r % nsl);
 * runlen[r] = max len over that instance (TrainingHelper runs max_b len steps). */
int d2p_seq_weights(const int* len, int R, int nsl, float coef, int max_len, float* w, int* runlen,
                    void* stream);
/* logits [T,R,V] (rows t >= runlen forced to 0); labels [R,T] int32;
 * loss[0] (+)= sum w[r]*ce over t < len[r]; dlogits optional. rowloss: scratch [T*R]. */
int d2p_softmax_ce(float* logits, int T, int R, int V, const int* labels, const int* len,
                   const int* runlen, const float* w, float* rowloss, float* dlogits, float* loss,
                   int accumulate, void* stream);
/* labels [R,T,P] float; ce = mean_P sigmoid_cross_entropy (model_full.py:651-653). */
int d2p_sigmoid_ce(float* logits, int T, int R, int P, const float* labels, const int* len,
                   const int* runlen, const float* w, float* rowloss, float* dlogits, float* loss,
                   int accumulate, void* stream);

/* ---- K4: greedy decode -------------------------------------------------------
 * dynamic_decode(BasicDecoder(cell, GreedyEmbeddingHelper(embed, start, end), (c,h),
 * Dense(V))) with maximum_iterations = max_len, reference models/model_full.py:424-435,
 * 513-521.  logits [max_len, R, V] time-major, zero past the executed steps; tokens
 * [max_len, R]; lengths [R] = first end-token step + 1, or max_len. */
size_t d2p_greedy_ws_bytes(int R, int H, int E);
int d2p_lstm_decoder_greedy(const float* table, int vocab_rows, int E, const float* W, const float* b,
                            const float* proj, int R, int H, int V, int start_id, int end_id,
                            int max_len, int nsl /* decoder instances: rows r share r % nsl */,
                            const float* h0, const float* c0, float* logits, int* tokens,
                            int* lengths, void* ws, size_t ws_bytes, void* stream);

/* Arg-max margin guard for a greedy decode that ran on the tensor-core (bf16x3) engine: *count =
 * number of executed positions (t < lengths[row]) of logits [Tdec, R, V] whose top-2 gap is
 * <= rel_tol * max(1, max|logit|).  North_star asks for bit-exact greedy token ids: when the count
 * is not zero the host repeats the decode on the exact fp32 engine (models/model_full.py:424-435
 * argmax semantics, lowest index wins). */
int d2p_greedy_near_ties(const float* logits, int Tdec, int R, int V, const int* lengths,
                         float rel_tol, int* count, void* stream);

/* ---- K6: pooled Luong attention + induction decoder --------------------------------
 * reference models/baselines/model_induction.py:25-53, 107-182, 638-709.
 * keys/values [T, R, H] time-major (R = B*k, keys = values * W_mem), mem_len [R];
 * queries q [B*test_k, H]; ctx = mean over the k memories of the attention contexts. */
size_t d2p_luong_pool_attention_ws_bytes(int B, int k, int tk, int H);
/* q rows have stride ldq floats, ctx rows stride ldc (>= H, multiples of 4; 16-byte aligned
 * buffers); ws holds the per-memory contexts [B, k, test_k, H] that are averaged in order. */
int d2p_luong_pool_attention(const float* q, int ldq, const float* keys, const float* values,
                             const int* mem_len, int B, int k, int tk, int T, int H, float* ctx,
                             int ldc, void* ws, size_t ws_bytes, void* stream);
size_t d2p_induction_decode_ws_bytes(int B, int k, int tk, int H);
/* tokens [B*test_k, Tdec] int32 = teacher forcing, NULL = greedy (start A, end A-1).
 * logits [Tdec, B*test_k, A]; greedy also fills out_tokens [Tdec, B*test_k], lengths.
 * keys = values * memory_layer precomputed by the caller, or keys == NULL and memory_layer [H,H]
 * (LuongAttention's Dense(num_units, no bias), model_induction.py:647-649) given: the layer is
 * folded into the query, each step then reads `values` only. */
int d2p_induction_decode(const float* keys, const float* memory_layer, const float* values,
                         const int* mem_len, int B, int k,
                         int tk, int T, int H, const float* h_sum, const float* c_sum,
                         const float* table, int A, const float* Wcell, const float* bcell,
                         const float* Wa, const float* proj, const int* tokens, int Tdec,
                         float* logits, int* out_tokens, int* lengths, void* ws, size_t ws_bytes,
                         void* stream);
/* out[row] = [A[row, :F1] ; B[row, :F2]]  (conv features ++ perception vector,
 * model_induction.py:399-424) */
/* Training forms of the pooled Luong attention (the induction baseline is trained by the reference's
 * trainer.py:102-109 through tf.gradients of models/baselines/model_induction.py:25-53, 107-182):
 * _train_fwd also returns the attention weights alpha [B,k,tk,T] (zero past the memory length);
 * _train_bwd takes dctx (gradient of the MEAN context, rows b*tk+j, stride ldd) and ACCUMULATES into
 * dq (stride lddq), dkeys and dvalues [T, B*k, H]: one decoder step per call, every memory row owned by
 * one CTA, fixed summation order. */
size_t d2p_luong_pool_attention_train_ws_bytes(int B, int k, int tk, int H);
int d2p_luong_pool_attention_train_fwd(const float* q, int ldq, const float* keys, const float* values,
                                       const int* mem_len, int B, int k, int tk, int T, int H, float* ctx,
                                       int ldc, float* alpha, void* ws, size_t ws_bytes, void* stream);
int d2p_luong_pool_attention_train_bwd(const float* q, int ldq, const float* keys, const float* values,
                                       const int* mem_len, const float* alpha, const float* dctx, int ldd,
                                       int B, int k, int tk, int T, int H, float* dq, int lddq, float* dkeys,
                                       float* dvalues, void* ws, size_t ws_bytes, void* stream);
/* dst[row, :F1] = src[row, c0 : c0+F1] (rows of width F): the conv-feature part of the gradient of the
 * [features ; perception] encoder input (model_induction.py:399-474; the perception vector is data) */
int d2p_split_cols(const float* src, int F, int c0, int F1, long long rows, float* dst, void* stream);
int d2p_concat_cols(const float* A, int F1, const float* Bm, int F2, long long rows, float* out,
                    void* stream);

/* ---- Karel DSL on the host: parser, interpreter, evaluation metrics --------
 * HOST pointers, CPU code.  Token ids follow the reference vocabulary (dsl_prob.py:13-28, INT
 * expanded).  Replaces the per-step Python py_funcs of models/model_full.py:602-616
 * (check_correct_syntax -> karel_env/dsl/dsl_parse.py:252-265), 747-787
 * (generate_program_output_karel -> dsl_parse.py rule closures + karel_env/karel.py:33-185),
 * 712-727 (exact_program_compare_karel -> karel_env/dsl/dsl_enum_program.py), 870-897
 * (CompareDemoAndExecution). */
int d2p_karel_check_syntax(const int* tokens, int len);          /* 1 parses, 0 does not */
/* state0 [h, w, 16] u8; s_h receives min(n_states, max_states) states (may be NULL).
 * returns 1 = ran to completion, 0 = run-time failure / time-out, -1 = does not parse */
int d2p_karel_execute(const int* tokens, int len, const unsigned char* state0, int h, int w,
                      int make_error, int max_states, unsigned char* s_h, int* n_states);
/* canonical-form comparison: 1 equal, 0 different, -1 either side is not a complete program */
int d2p_karel_programs_equal(const int* a, int la, const int* b, int lb);
/* tokens [B, L], lens [B], is_same_seq [B] u8, demos [B, k, T, h, w, 16] u8, demo_len [B, k];
 * out: is_correct_syntax [B], is_correct_execution [B, k], num_correct_execution [B].
 * nthreads <= 0: all host threads. */
int d2p_karel_eval_batch(const int* tokens, const int* lens, const unsigned char* is_same_seq, int B,
                         int L, const unsigned char* demos, const int* demo_len, int k, int T, int h,
                         int w, int make_error, float* is_correct_syntax,
                         float* is_correct_execution, float* num_correct_execution, int nthreads);

/* ---- fc -> (lrelu) -> BN : ops.fc, reference models/ops.py:149-155 ------------
 * used by Per_Encoder (model_full.py:308-316; act = 0, per-demo slices) */
typedef struct {
    const float* w;      /* [cin, cout] */
    const float* b;
    const float* gamma;
    const float* beta;
    float* moving_mean;
    float* moving_var;
    float* dw;
    float* db;
    float* dgamma;
    float* dbeta;
} d2p_fc_bn;

size_t d2p_fc_bn_saved_floats(long long rows, int cout, int nsl);
size_t d2p_fc_bn_ws_bytes(long long rows, int cout, int nsl);
/* slice(row) = (row / seg) % nsl selects the BN statistics group. */
int d2p_fc_bn_fwd(const float* X, long long rows, int cin, int cout, int seg, int nsl, int act,
                  const d2p_fc_bn* p, float* Y, float* saved, int training, void* ws,
                  size_t ws_bytes, void* stream);
int d2p_fc_bn_bwd(const float* X, long long rows, int cin, int cout, int seg, int nsl, int act,
                  const d2p_fc_bn* p, const float* dY, const float* saved, float* dX, int training,
                  void* ws, size_t ws_bytes, void* stream);

/* ---- K5: rn_pool, reference models/model_full.py:333-349 ---------------------
 * F [B,k,H] -> pooled [B,H]; fc1->w is [2H,H], fc2->w is [H,H]. */
size_t d2p_rn_pool_saved_floats(int B, int k, int H);
size_t d2p_rn_pool_ws_bytes(int B, int k, int H);
int d2p_rn_pool_fwd(const float* F, int B, int k, int H, const d2p_fc_bn* fc1,
                    const d2p_fc_bn* fc2, float* pooled, float* saved, int training, void* ws,
                    size_t ws_bytes, void* stream);
/* phases / hold: hold == NULL and phases == 3: everything in one call.  With a caller-owned hold
 * buffer (d2p_rn_pool_bwd_hold_floats floats) the call may be split: phases & 1 = data path (dF and
 * the BatchNorm scale/shift gradients), phases & 2 = the fc weight / bias gradients, which only read
 * `hold` and may run later on another stream (with that stream's own workspace). */
size_t d2p_rn_pool_bwd_hold_floats(int B, int k, int H);
int d2p_rn_pool_bwd(const float* F, int B, int k, int H, const d2p_fc_bn* fc1,
                    const d2p_fc_bn* fc2, const float* dpooled, const float* saved, float* dF,
                    int training, void* ws, size_t ws_bytes, int phases, float* hold, void* stream);

/* ---- small [B,k,H] reductions (SummarizeFeature avgpool, model_full.py:351-362) */
int d2p_group_sum(const float* F, int B, int k, int H, float alpha, float* out, int accumulate,
                  void* stream);
int d2p_group_bcast(const float* S, int B, int k, int H, float alpha, float* out, int accumulate,
                    void* stream);
/* demo_aggregation == 'maxpool' of synthesis_baseline / induction_baseline (tf.layers.max_pooling1d over the k
 * demonstrations, reference models/baselines/model_synthesis.py:344-356): out[b,u] = max_i F[b,i,u]; the winning
 * demonstration index (lowest among equal maxima) is kept for the backward pass, which routes the whole
 * gradient to it like MaxPoolGrad. */
int d2p_group_max(const float* F, int B, int k, int H, float* out, int* arg, void* stream);
int d2p_group_max_bwd(const float* dout, const int* arg, int B, int k, int H, float* dF, void* stream);
int d2p_axpby(const float* x, float alpha, float* y, float beta, size_t n, void* stream);
/* y = (a + b) + c: the sum of the three gradient contributions to the per-demonstration summary state
 * (action decoder, perception decoder, summary pools; reference models/model_full.py:918-1079 adds the three
 * losses, so the state's gradient is the sum of the three branches) in one pass. */
int d2p_add3(const float* a, const float* b, const float* c, float* y, size_t n, void* stream);
/* layout helpers at the facade boundary */
int d2p_logits_to_bvl(const float* X, int T, int R, int V, float* Y, void* stream);
int d2p_rtp_to_trp(const float* X, int R, int T, int P, float* Y, void* stream);
int d2p_len_to_int(const float* x, int* y, int n, void* stream);

/* ---- K7: clip_by_global_norm(20) + Adam, reference trainer.py:102-109 --------
 * state: 8 device doubles, zeroed before the first step ([0] = step count,
 * [3] = last global norm). grad_scale = 1/world_size after the NCCL sum. */
size_t d2p_adam_ws_bytes(void);
int d2p_clip_adam_step(float* params, const float* grads, float* m, float* v, size_t n, float lr,
                       float b1, float b2, float eps, float clip_norm, float grad_scale,
                       int staircase_decay_steps, double* state, void* ws, size_t ws_bytes,
                       void* stream);

/* ---- GEMM engine (exposed for tests): row-major
 * C[M,N] = alpha*op(A)*op(B) + beta*C (+ bias[N]) */
int d2p_gemm(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
             const float* B, int ldb, float beta, float* C, int ldc, const float* bias,
             void* stream);

/* Tensor-core engine (tcgen05 + TMEM, bf16x3 split = fp32-equivalent products).
 * Same contract as d2p_gemm; ws holds the bf16 hi/lo operand splits. */
size_t d2p_gemm_tc_ws_bytes(int M, int N, int K);
int d2p_gemm_tc(int transA, int transB, int M, int N, int K, float alpha, const float* A, int lda,
                const float* B, int ldb, float beta, float* C, int ldc, const float* bias, void* ws,
                size_t ws_bytes, void* stream);

/* Packed-operand interface of the tensor-core engine: Op[mn,k] (MN x K) is split into
 * bf16 hi/lo and stored as UMMA core matrices (d2p_packed_bytes bytes).  k_contig: the
 * fp32 source is S[mn*ld + k] (1) or S[k*ld + mn] (0).  d2p_gemm_tc_packed computes
 * C = alpha*A*B^T-form product of two packed operands (A: M x K, B: N x K); with
 * ksplit > 1 (or C == NULL) raw partial sums go to partials[ksplit][M][N]. */
size_t d2p_packed_bytes(int MN, int K);
int d2p_pack_bf16(const float* S, int MN, int K, int ld, int k_contig, void* out, void* stream);
int d2p_gemm_tc_packed(const void* Apk, const void* Bpk, int M, int N, int K, float alpha, float beta,
                       float* C, int ldc, const float* bias, int ksplit, float* partials,
                       void* stream);

/* Arena for the tensor-core engine: `scratch` holds packed activations (reused by
 * every GEMM on the stream), `cache` holds packed weights until d2p_tc_new_step()
 * (call it whenever parameters changed).  With no arena configured (or
 * enabled = 0) every contraction runs on the exact-fp32 SIMT engine. */
int d2p_tc_configure(void* scratch, size_t scratch_bytes, void* cache, size_t cache_bytes,
                     int enabled);
int d2p_tc_new_step(void);
/* per-stream scratch arena for concurrent branches (call after d2p_tc_configure) */
int d2p_tc_bind_stream(void* stream, void* scratch, size_t scratch_bytes);
/* Recurrence scheduling of d2p_lstm_seq_fwd / d2p_lstm_seq_bwd (the while_loop of
 * tf.nn.dynamic_rnn / dynamic_decode, reference models/model_full.py:254,274,469):
 * 1 (default) = ONE persistent cooperative kernel per sequence where supported
 * (H = 512, <= 512 rows, tensor-core arena configured): recurrent weight resident in
 * shared memory, cell state in registers, steps separated by a release/acquire barrier
 * among the CTAs of a row tile; 0 = one launch per time step; 2 = persistent kernels launched
 * without the cooperative attribute (profiling under ncu).  Bit 2 (value 4, added to the mode): the
 * producer lane's reader-side proxy fence covers all state spaces instead of shared memory only
 * (A/B switch; see proxy_fence_reader in csrc/lstm_persist.cu).  Bit 3 (value 8): 32-row recurrences
 * (program decoder at B = 32) keep the interleaved hi/lo operand and two MMAs per k16 instead of the
 * stacked operand (rows 0-31 = bf16 hi, rows 32-63 = lo of the same rows: one MMA per k16). */
int d2p_lstm_set_persistent(int mode);
/* Synchronises the device and reports (then clears) whether a step barrier of one of the persistent
 * cooperative kernels ran into its ~2 s spin limit since the last call (0 = none, bit 0 = LSTM
 * recurrence, bit 1 = fused conv encoder); the outputs of such a launch are invalid.  The engine
 * checks it whenever it hands a loss to the caller. */
int d2p_device_error(int* flags);
/* Stream-ordered, non-clearing form for the training loop: dst[0] = LSTM recurrence word, dst[1] =
 * fused conv encoder word (device memory; capturable into a CUDA graph).  d2p_clip_adam_step reads
 * the same words and skips the update (parameters, slots and step counter untouched) when one is
 * set, so a failed step never reaches the weights. */
int d2p_device_error_async(unsigned* dst, void* stream);
/* test hook: set the sticky words as a timed-out barrier would (bit 0 / bit 1 as above) */
int d2p_debug_inject_device_error(int flags);
/* developer tool: record SM-clock stamps of CTA (0,0,0) of the tensor-core kernels
 * into buf (>= 128 int64 on the device; the persistent recurrence kernels use
 * slots 64..127); NULL disables. */
int d2p_debug_set_probe(long long* buf);
/* developer tool: one-thread kernel that writes %globaltimer (ns) into buf[slot] in stream
 * order; called between ops during graph capture it yields the timeline of a replay. */
int d2p_debug_stamp(unsigned long long* buf, int slot, void* stream);

/* Scheduling of large tensor-core products (more 128 x 128 output tiles than SMs, no split-K), e.g.
 * the hoisted LSTM gate GEMM [T*R, In] x [In, 4H] (x*Wx of BasicLSTMCell, reference
 * models/model_full.py:244-258): 1 (default) = persistent kernel, one CTA per SM pulling tiles from
 * an atomic counter, accumulator double-buffered in TMEM so the epilogue of tile i overlaps the
 * k-loop of tile i+1; 0 = one CTA per tile. */
int d2p_gemm_set_persistent(int mode);
/* Host-side CRC-32C (Castagnoli), running form: pass 0 (or the value returned for the bytes so
 * far).  Used by demo2program_b200/tf_checkpoint.py for the block and tensor checksums of
 * TensorFlow checkpoint files (tf.train.Saver at reference trainer.py:114,145,182, evaler.py:82-99). */
unsigned int d2p_crc32c(const void* data, size_t n, unsigned int crc);

#ifdef __cplusplus
}
#endif
#endif /* D2P_H_ */
