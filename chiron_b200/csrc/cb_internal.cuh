// Internal declarations shared by the translation units of libchiron_b200.so.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string>
#include <vector>

#include "../../include/chiron_b200.h"
#include "cb_tc_common.cuh"

#include "cb_simt_types.h"

#define CB_PROF_MAX 96
enum { CB_CAT_CONV = 0, CB_CAT_LSTM_IN = 1, CB_CAT_LSTM_REC = 2, CB_CAT_HEAD = 3, CB_CAT_COUNT = 4 };

void cb_set_error(const char* fmt, ...);

#define CB_CUDA(call)                                                                          \
    do {                                                                                       \
        cudaError_t e__ = (call);                                                              \
        if (e__ != cudaSuccess) {                                                              \
            cb_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
            return CB_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

#define CB_CHECK_LAUNCH()                                                                      \
    do {                                                                                       \
        cudaError_t e__ = cudaGetLastError();                                                  \
        if (e__ != cudaSuccess) {                                                              \
            cb_set_error("%s:%d kernel launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e__)); \
            return CB_ERR_CUDA;                                                                \
        }                                                                                      \
    } while (0)

// ---- one tensor-core contraction (cb_tc.cu) ----------------------------------------------------------------------------
// Row space: m = to*Bp + b (output frame to < T, window b < B; Bp = B rounded up to 128), so a 128-row tile is 128
// consecutive windows of one frame.  K is image a0 read at frames to*stride + j - left (j < taps) followed by image a1
// read at frame to (the appended 1x1 branch input).
struct TcGemm {
    int layer_id;             // prepared weight image
    int T, B, Bp;
    CbImg a0, a1;
    int a0_plane0, a1_plane0;                 // first k-group plane of each source
    int a0_chunks_per_tap, taps, left, stride, a1_chunks;     // K = (taps*a0_chunks_per_tap + a1_chunks) * 32
    int a1_stride;            // frame stride of the appended 1x1 branch input (0 = 1: the strided branch1 of a strided block)
    int N; const float* shift; int relu;
    int res; const float* xT; int res_stride; const float *rw, *rinv, *rsh;   // + (xT[to*res_stride][b]*rw)*rinv + rsh
    int out_mode;             // 1 fp32 time-major out[to][ldo][Bp]; 2 operand image o
    float* out; int ldo;
    CbImg o; int o_plane0;
};

struct cb_handle {
    int device;
    int precision;
    CbConfig cfg;
    int sm_count;
    // device weights (fp32, derived)
    float* d_weights;                 // one allocation holding everything below
    size_t weights_floats;
    CbConvW conv2a[CB_MAX_BLOCKS], conv2b[CB_MAX_BLOCKS], convc[CB_MAX_BLOCKS];
    // batch-statistics BN mode: the convolutions as they are in the checkpoint (nothing folded) + BN scale/offset
    int bn_mode;
    CbRawConv raw1[CB_MAX_BLOCKS], raw2a[CB_MAX_BLOCKS], raw2b[CB_MAX_BLOCKS], raw2c[CB_MAX_BLOCKS];
    const float* zeros;                   // [max(C, 8H)] zero shift vector
    CbStem stem;                          // stem convolution (cfg.stem_k > 0)
    double* bn_part;                      // [CB_BN_MAX_PART][2][C] partial sums
    float* bn_vec;                        // [CB_BN_VECS][C]
    const float *g_w, *g_inv, *g_sh;  // block-1 conv2a rank-1 generator
    const float *r_w, *r_inv, *r_sh;  // block-1 branch1 rank-1 residual
    const float* wx[CB_MAX_LAYERS][2];   // LSTM input kernels  [in,4H] per direction
    const float* bias[CB_MAX_LAYERS][2]; // [4H]
    const float* wxcat[CB_MAX_LAYERS];   // [in, 8H] fw||bw (stacked-bidirectional layout)
    const float* bcat[CB_MAX_LAYERS];    // [8H]
    const float* whh[CB_MAX_LAYERS][2];  // [H,4H] (LSTM) or [H,2H] recurrent gate kernel (GRU)
    const float* gru_wc[CB_MAX_LAYERS][2];  // [H,H] recurrent candidate kernel (GRU)
    const float *head_w, *head_b, *head_wc, *head_bc;
    // workspace
    void* ws; size_t ws_bytes;
    float *act[3]; float* pre; float* lstm_out[2];
    int ws_B, ws_L;
    void* stage; size_t stage_bytes;   // device staging of the host-buffer API
    struct PipeSlot {                  // cb_basecall_submit / cb_basecall_collect
        void* pin; size_t pin_bytes;   // pinned host staging: x, seq_len | bases, n_bases, path_prob
        void* dev; size_t dev_bytes;   // device: x, seq_len_in, seq_len_out, n_bases, path_prob, logits, bases
        cudaEvent_t h2d_done, compute_done, d2h_done;
        int B, L, T, busy;
    } pipe[2];
    cudaStream_t pipe_in, pipe_compute, pipe_out;
    cudaStream_t asm_stream;          // cb_assemble_host's own non-blocking stream (created on first use)
    void* asm_stage; size_t asm_stage_bytes;   // device staging of cb_assemble_host (grow-only: no cudaFree in the read loop)
    const float* fea;                  // CNN feature of the last forward (debug fetch)
    void* tc;                          // tensor-core path state (cb_tc.cu)
    void* lstm_tc;                     // tensor-core recurrence state (cb_lstm_tc.cu)
    void* tc_ws; size_t ws_bytes_tc;   // tensor-core workspace: operand images, pre, out (cb_forward_tc.cu)
    int last_Bp, last_tmajor;
    void* beam_ws; size_t beam_ws_bytes;
    void* asm_ws; size_t asm_ws_bytes;
    int* d_flag;
    long long launches;
    int reserve_sms;                  // SMs left out of the persistent GEMM grids (cb_reserve_sms)
    int timing; cudaEvent_t ev[8]; int have_ms;
    // per-launch profile of the last cb_forward (timing on): event pairs tagged with a kernel category
    cudaEvent_t prof_ev[CB_PROF_MAX][2]; int prof_cat[CB_PROF_MAX]; int prof_n; int prof_ready;
    int last_B, last_T;
};

// ---- launchers (each returns CB_OK or an error code; they bump h->launches) -----------------------------------------
int cb_launch_gemm_simt(cb_handle* h, const GemmProblem& p, cudaStream_t s);
int cb_launch_lstm_simt(cb_handle* h, const LstmProblem& p, cudaStream_t s);
int cb_launch_gru_simt(cb_handle* h, const GruProblem& p, cudaStream_t s);
int cb_launch_stem(cb_handle* h, const StemProblem& p, cudaStream_t s);
int cb_launch_bn_rank1(cb_handle* h, const float* x, int B, int t_in, int stride, int t_out, const float* w,
                       const float* scale, const float* offset, float* inv, float* shift, cudaStream_t s);
int cb_launch_bn_stats(cb_handle* h, const float* X, long long M, const float* scale, const float* offset, float* inv,
                       float* shift, cudaStream_t s);
int cb_launch_bn_apply(cb_handle* h, const BnApplyArgs& a, cudaStream_t s);
int cb_launch_gemm_tc(cb_handle* h, const TcGemm& g, cudaStream_t s);
const float* cb_tc_lstm_bias(cb_handle* h, int layer, int d);   // biases in unit-major gate-column order
double cb_tc_trunc_c();                    // mean relative shrink per truncating tensor-core accumulator add
#define CB_LSTM_TC_CHAIN 21                // MMAs the recurrence chains on top of the pre-loaded input projection
int cb_launch_transpose_x(cb_handle* h, const float* x, int B, int L, int Bp, float* xT, cudaStream_t s);
int cb_launch_gen_conv2a(cb_handle* h, const float* xT, int B, int Bp, int L, const CbImg& o, cudaStream_t s);
int cb_launch_stem_image(cb_handle* h, const float* xT, int B, int Bp, int L, int t_out, int left, const CbImg& o, cudaStream_t s);
int cb_launch_lstm_tc(cb_handle* h, const LstmProblem& p, const CbImg* o_img, int write_f32, cudaStream_t s);
int cb_tc_prepare(cb_handle* h, const float* host_weights);   // build fp16 hi/lo operand images from d_weights layout
void cb_tc_release(cb_handle* h);
// Flags of the handle's 64-byte device status block (h->d_flag): scratch of the beam passes, and the two sticky error flags
// cb_check_deferred reports (beam search out of fallback workspaces; an activation left the fp16 range of the tc path).
enum { CB_FLAG_BEAM_MARKED = 0, CB_FLAG_BEAM_SLOTS = 1, CB_FLAG_BEAM_ERROR = 2, CB_FLAG_TC_RANGE = 3 };
int cb_check_deferred(cb_handle* h, cudaStream_t s);   // synchronises s; CB_ERR_RANGE / CB_ERR_NOMEM for a raised flag
int cb_forward_tc(cb_handle* h, const float* x, const int32_t* seq_len_out, int B, int L, float* logits,
                  float* path_prob, cudaStream_t s);
void cb_forward_tc_release(cb_handle* h);
long long cb_debug_fetch_tc(cb_handle* h, int what, float* dst, size_t max_floats);
int cb_prof_begin(cb_handle* h, int cat, cudaStream_t s);      // event pair around a launch when timing is on
void cb_prof_end(cb_handle* h, int i, cudaStream_t s);
int cb_lstm_tc_prepare(cb_handle* h, const float* host_weights);
void cb_lstm_tc_release(cb_handle* h);
bool cb_lstm_tc_available(const cb_handle* h);
int cb_launch_head_tmajor(cb_handle* h, const float* out_t, int B, int Bp, int T, float* logits, cudaStream_t s);
int cb_launch_head(cb_handle* h, const float* lasth, int M, float* logits, cudaStream_t s);
int cb_launch_path_prob(cb_handle* h, const float* logits, int B, int T, float* prob, cudaStream_t s);
int cb_launch_seq_len(cb_handle* h, const int32_t* in, int B, int L, int T, int32_t* out, cudaStream_t s);
int cb_launch_greedy(cb_handle* h, const float* logits, const int32_t* lens, int B, int T, int8_t* bases,
                     int32_t* n_bases, cudaStream_t s);
int cb_launch_beam(cb_handle* h, const float* logits, const int32_t* lens, int B, int T, int W, int8_t* bases,
                   int32_t* n_bases, float* scores, cudaStream_t s);      // scores: top-path log probability [B] or null
int cb_launch_assemble(cb_handle* h, const int8_t* bases, const int32_t* n_bases, const float* path_prob, int n_windows,
                       int T, int jump, int L, int kernel, int8_t* consensus, char* qual, int32_t* pos,
                       int32_t* out_len, int max_len, cudaStream_t s);
