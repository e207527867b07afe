/* chiron_b200 -- C ABI of the B200-native basecalling hot path.
 *
 * The reference (haotianteng/Chiron) has no FFI: its operator boundary is the pair of TensorFlow session calls
 *   sess.run(logits_enqueue, {x[B,L] f32, seq_length[B] i32, training=False})     chiron/chiron_eval.py:335-342
 *   sess.run([decoded_fname, decode_idx, decode_predict, decode_prob])            chiron/chiron_eval.py:408-409
 * restated as a SavedModel signature {x, seq_len} -> {indices, values, dense_shape, logits, prob_logits}
 * in chiron/export_test.py:103-113.  Every entry point below names the reference code it replaces.
 *
 * Conventions: plain C, no torch types.  All functions return 0 on success or a negative CB_ERR_* code;
 * cb_last_error() gives the message of the last failure on the calling thread.  Unless a function says "host",
 * pointers are DEVICE pointers on the handle's GPU and work is enqueued asynchronously on `stream`
 * (a cudaStream_t passed as void*; NULL = the legacy default stream).  One handle per GPU; a handle is not
 * re-entrant (the reference has a single in-flight forward pass: one feeder thread, chiron_eval.py:369-372), with one
 * exception: cb_assemble_host / cb_assemble keep their own staging and workspace, so ONE other thread may run them while
 * the forward/decode entry points (cb_forward, cb_decode_*, cb_basecall_*) run on the first -- the reference likewise
 * assembles on its main thread while the feeder thread keeps the session busy (chiron_eval.py:369-372,446-457).
 */
#ifndef CHIRON_B200_H
#define CHIRON_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct cb_handle cb_handle;

enum {
    CB_OK = 0,
    CB_ERR_ARG = -1,      /* bad argument / unsupported shape */
    CB_ERR_BLOB = -2,     /* malformed weight blob */
    CB_ERR_CUDA = -3,     /* CUDA runtime failure (message has the cudaError string) */
    CB_ERR_NOMEM = -4,    /* workspace allocation failed */
    CB_ERR_RANGE = -5     /* an activation left the fp16 range of the tensor-core path; rerun with CB_PREC_FP32 */
};

/* Arithmetic of the dense contractions (convolutions, LSTM gate GEMMs).  Accumulation is always fp32. */
enum {
    CB_PREC_FP32 = 0,     /* fp32 FFMA SIMT kernels (reference-grade; slow path kept for A/B checks) */
    CB_PREC_TC_SPLIT = 1  /* tcgen05 fp16 MMAs on hi/lo-split operands (3 MMAs per product), short-K partial sums added
                           * in fp32 registers (the tensor core's accumulator truncates): fp32-class error.  The
                           * production mode: `chiron call`, bench.py and the parity tests run it. */
};

/* BatchNorm statistics of the residual conv stack.  The shipped checkpoints hold pop_mean/pop_var and their graph uses
 * them at inference (batchnorm(), chiron/cnn.py:125-163): CB_BN_POPULATION, folded into the weights by cb_create.
 * HEAD's conv_layer calls simple_global_bn (chiron/cnn.py:65-68,166-188), which normalises with the moments of the
 * CURRENT batch (tf.nn.moments over every frame of every window) even at inference: CB_BN_BATCH, for models trained at
 * HEAD.  In that mode a window's result depends on the batch it is in, exactly as in the reference. */
enum { CB_BN_POPULATION = 0, CB_BN_BATCH = 1 };

/* Recurrent cell of the model (CBW1 header field cell_type; model.json "cell_type", chiron/rnn.py:47-53,126-131).
 * GRU models run on CB_PREC_FP32 handles only (cb_create fails with CB_ERR_ARG otherwise); BNLSTM is not supported. */
enum { CB_CELL_LSTM = 0, CB_CELL_GRU = 1 };

/* Assembly kernels, chiron/chiron_eval.py:138-150 (get_assembler_kernal). */
enum { CB_ASM_SIMPLE = 0, CB_ASM_GLUE = 1, CB_ASM_STICK = 2 };

/* -- lifecycle --------------------------------------------------------------------------------------------------- */

/* Upload a CBW1 weight blob (chiron_b200/model.py), fold population BatchNorm into the conv weights, build the
 * packed operand images.  Replaces build_eval_graph + tf.train.Saver.restore (chiron_eval.py:244-276).
 * `blob` is a HOST pointer. */
int cb_create(const void* blob, size_t nbytes, int device, int precision, cb_handle** out);
int cb_destroy(cb_handle* h);
const char* cb_last_error(void);
const char* cb_version(void);

/* Override the BatchNorm mode the blob header asks for (CBW1 header field bn_mode).  CB_BN_BATCH is available on
 * CB_PREC_FP32 handles only (CB_ERR_ARG otherwise). */
int cb_set_bn_mode(cb_handle* h, int bn_mode);
int cb_bn_mode(const cb_handle* h);

/* Model facts read from the blob header. */
int cb_out_len(const cb_handle* h, int L);            /* CNN output frames T for an L-sample window (ratio = L/T) */
int cb_n_class(const cb_handle* h);                   /* 5: A C G T blank */
int cb_precision(const cb_handle* h);
size_t cb_workspace_bytes(const cb_handle* h);        /* current device workspace (grown on demand) */

/* -- the per-window hot path (device pointers, async) --------------------------------------------------------------- */

/* seq_len_out[b] = round_half_even(seq_len_in[b] / ratio)  -- chiron_eval.py:337.  */
int cb_seq_len_out(cb_handle* h, const int32_t* seq_len_in, int B, int L, int32_t* seq_len_out, void* stream);

/* chiron_model.inference (chiron_model.py:134-172: getcnnfeature -> rnn_layers -> logits) + path_prob
 * (chiron_eval.py:116-136).  x[B,L] normalised signal windows (zero padded); seq_len_out[B] frames per window
 * (already divided by ratio); logits[B,T,n_class]; path_prob[B] (may be NULL). */
int cb_forward(cb_handle* h, const float* x, const int32_t* seq_len_out, int B, int L,
               float* logits, float* path_prob, void* stream);

/* Synchronise `stream` and report deferred device-side errors of the asynchronous calls (CB_ERR_RANGE when an activation
 * left the fp16 range of the tensor-core path in cb_forward; CB_ERR_NOMEM when cb_decode_beam ran out of fallback
 * workspaces).  The flags are sticky until reported.  cb_basecall_host and cb_basecall_collect call this themselves. */
int cb_check_status(cb_handle* h, void* stream);

/* tf.nn.ctc_greedy_decoder(merge_repeated=True) (chiron_eval.py:486-487).  Dense padded output replaces the
 * SparseTensor: bases[B,T] int8 (0..3), n_bases[B]; rows with n_bases == 0 are the rows sparse2dense drops. */
int cb_decode_greedy(cb_handle* h, const float* logits, const int32_t* seq_len_out, int B, int T,
                     int8_t* bases, int32_t* n_bases, void* stream);

/* tf.nn.ctc_beam_search_decoder(merge_repeated=False, beam_width, top_paths=1) (chiron_eval.py:489-492).  Asynchronous like
 * the other entry points: three passes are enqueued back to back (warp per window with a shared-memory trie; the windows
 * whose trie outgrew it alone, one per CTA; what is still left over global workspaces that cannot overflow) and nothing is
 * read back -- a failure of the last resort is reported by cb_check_status / cb_basecall_collect. */
int cb_decode_beam(cb_handle* h, const float* logits, const int32_t* seq_len_out, int B, int T, int beam_width,
                   int8_t* bases, int32_t* n_bases, void* stream);

/* The same search, also returning the decoder's second output: log_prob[B] = the log probability TopPaths() reports for the
 * decoded path (newp.total of the best beam; per-frame max-subtracted logits, TF 1.15) -- the `log_prob` tensor of the
 * serving signature {x, seq_len} -> {indices, values, dense_shape, logits, prob_logits, log_prob} (chiron/export_test.py:36-40,
 * 103-113).  The reference holds no fixture for it: pinned to the oracle's restatement only. */
int cb_decode_beam_scored(cb_handle* h, const float* logits, const int32_t* seq_len_out, int B, int T, int beam_width,
                          int8_t* bases, int32_t* n_bases, float* log_prob, void* stream);

/* simple_assembly(_qs) + argmax + qs() for ONE read (easy_assembler.py:302-335,393-442; chiron_eval.py:152-174,457).
 * bases[n_windows,T] / n_bases[n_windows] / path_prob[n_windows] in TRUE window order (empty windows are skipped like
 * sparse2dense does).  Outputs: consensus[max_len] int8 base indices, qual[max_len] phred+33 chars (may be NULL),
 * pos[n_windows] window start coordinates (-1 for skipped windows), *out_len consensus length (device int32). */
int cb_assemble(cb_handle* h, const int8_t* bases, const int32_t* n_bases, const float* path_prob,
                int n_windows, int T, int jump, int L, int kernel,
                int8_t* consensus, char* qual, int32_t* pos, int32_t* out_len, int max_len, void* stream);

/* -- host-buffer convenience (what a Python/ctypes or cgo caller binds) ---------------------------------------------- */

/* One batch through forward + decode with HOST buffers: copies x/seq_len_in to the GPU, runs cb_seq_len_out,
 * cb_forward, cb_decode_greedy (beam_width == 0) or cb_decode_beam, copies bases/n_bases/path_prob (and logits when
 * non-NULL) back and synchronises.  Replaces one _worker_fn feed + one decode dequeue (chiron_eval.py:335-342,
 * 408-409). */
int cb_basecall_host(cb_handle* h, const float* x, const int32_t* seq_len_in, int B, int L, int beam_width,
                     int8_t* bases, int32_t* n_bases, float* path_prob, float* logits);

/* Two-slot asynchronous form of cb_basecall_host (replaces the feeder thread + FIFO queue + decode
 * threads of chiron_eval.py:262-268,495-521).  cb_basecall_submit copies the batch into the slot's PINNED staging buffer
 * (the caller's arrays may be reused as soon as it returns), then enqueues H2D on a copy-in stream, seq_len + forward +
 * decode on the compute stream and D2H into pinned memory on a copy-out stream, chained by events: the copies of batch
 * i+1 and the host work of the caller overlap the kernels of batch i.  cb_basecall_collect waits for that slot, copies
 * bases[B,T] / n_bases[B] / path_prob[B] (may be NULL) out and reports deferred errors.  slot is 0 or 1; a slot must be
 * collected before it is submitted again; batches complete in submission order. */
int cb_basecall_submit(cb_handle* h, int slot, const float* x, const int32_t* seq_len_in, int B, int L, int beam_width);
int cb_basecall_collect(cb_handle* h, int slot, int8_t* bases, int32_t* n_bases, float* path_prob);

/* Host-buffer form of cb_assemble for one read (synchronous). */
int cb_assemble_host(cb_handle* h, const int8_t* bases, const int32_t* n_bases, const float* path_prob,
                     int n_windows, int T, int jump, int L, int kernel,
                     int8_t* consensus, char* qual, int32_t* pos, int32_t* out_len, int max_len);

/* -- host-side signal preparation (no GPU work; rows a1-a2 of the path) ------------------------------------------------ */

/* The token loop of read_signal (chiron/chiron_input.py:527-532): ASCII-whitespace-separated numbers -> float32 samples
 * (parsed as double, rounded once, like numpy's string -> float32 cast).  Returns the number of samples, or CB_ERR_ARG for
 * a token that is not a number or when more than `cap` samples are present.  out == NULL only counts. */
long long cb_host_parse_signal(const char* text, size_t nbytes, float* out, size_t cap);

/* Signal normalisation (chiron/chiron_input.py:535-538, 548-554): mode = the CBW1 header's sig_norm --
 * 0 none, 1 (s - median(unique(s))) / mad(unique(s)), 2 (s - median(s)) / mad(s), mad = median(|x - median|) / 0.6745
 * (statsmodels.robust.mad); float64 statistics, result rounded once to float32.  in and out may alias. */
int cb_host_normalize(const float* in, size_t n, int mode, float* out);

/* read_data_for_eval + padding (chiron/chiron_input.py:253-292, 681-692): windows of L samples starting at 0, jump,
 * 2*jump ... < n; tails zero padded, lens[w] = true length.  Returns the window count ceil(n / jump); with x == NULL or
 * lens == NULL it only counts. */
long long cb_host_windows(const float* sig, size_t n, int jump, int L, float* x, int32_t* lens, size_t cap_windows);

/* The segment records write_output puts into segments/<name>.<ext> (chiron/chiron_eval.py:211-214): for every window with
 * n_bases > 0 (sparse2dense drops the others, :56-66), in window order, ">" name idx "\n" ACGT... "\n" with idx counting
 * the kept windows.  Returns the bytes written (out == NULL: the bytes needed). */
long long cb_host_format_segments(const char* name, const int8_t* bases, const int32_t* n_bases, int n_windows, int T,
                                  char* out, size_t cap);

/* Leave `n` SMs (0..16, rounded up to even) out of the grids of the persistent tensor-core contractions of this handle.
 * For callers that run other small kernels next to the forward pass -- evaluation() assembles finished reads on a second
 * stream while the next batches run: a persistent grid that fills every SM makes each of those kernels wait for a kernel
 * boundary and then delays one CTA of the next contraction by its run time (measured on files -> fastq, 1 x B200:
 * 61.7 Msamples/s with n = 0, 64.8 / 66.0 / 65.2 with n = 2 / 4 / 8; the resident step itself does not slow down, it runs
 * at the power cap).  Default 0 (CB_RESERVE_SMS overrides).  No reference counterpart: TF's executor owns the device there. */
int cb_reserve_sms(cb_handle* h, int n);

/* -- introspection for tests / benchmarks ----------------------------------------------------------------------------- */

/* Number of kernels this library has launched on the handle since creation (bench.py's gpu_launches). */
long long cb_launch_count(const cb_handle* h);

/* Phase times of the last cb_forward (CUDA events on its stream; call after synchronising it): fills ms[0..n) with
 * {conv stack, BiLSTM stack, head + path_prob, total}; returns how many entries were written. */
int cb_last_forward_ms(const cb_handle* h, float* ms, int n);
void cb_enable_timing(cb_handle* h, int on);

/* Per-kernel-category device time of the last cb_forward (timing on; CUDA event pairs around every launch on the
 * forward's stream): category 0 = conv-stack contractions, 1 = LSTM input-projection contractions, 2 = LSTM recurrence,
 * 3 = logit head.  ms[c] = summed duration, count[c] = launches.  Returns the number of categories written. */
int cb_last_forward_profile(const cb_handle* h, float* ms, int* count, int n);

/* Copy an intermediate activation of the last cb_forward to HOST memory (tests only).
 * what: 0 = CNN feature [B*T,C]; n_layers-1 / n_layers = output [B*T,2H] of the last two LSTM layers (earlier ones are
 * overwritten by the ping-pong buffers).  Returns floats copied or <0. */
long long cb_debug_fetch(cb_handle* h, int what, float* dst, size_t max_floats);

/* Pinned host memory for callers that want full-rate host<->device copies in cb_basecall_host. */
void* cb_host_alloc(size_t bytes);
void cb_host_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* CHIRON_B200_H */
