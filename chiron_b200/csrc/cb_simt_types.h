// Plain C++ types of the fp32 (SIMT) path: model geometry, contraction / BatchNorm problem descriptors and the weights
// a residual block reads.  No CUDA headers: tests/cuda_emu compiles the kernels and the conv-stack orchestration that use
// these types for the host.
#pragma once
#include <stdint.h>

#define CB_MAX_BLOCKS 8
#define CB_MAX_LAYERS 8
#define CB_BN_MAX_PART 1024       // per-CTA partial sums of a batch-statistics BN reduction (cb_bn_kernels.cuh)
#define CB_BN_VECS 8              // [C]-float scratch vectors holding inv/shift pairs of the BNs in flight

// ---- model (host copy of the CBW1 header) ----------------------------------------------------------------------
struct CbConfig {
    int n_blocks, channels, hidden, n_layers, n_class, rnn_layout, branch1_bn_mask;
    int k[CB_MAX_BLOCKS], stride[CB_MAX_BLOCKS];
    int sig_norm, reverse_signal;
    int cell_type;            // 0 LSTMCell, 1 GRUCell
    int stem_k, stem_stride;  // 0, 0 = none; else a strided 1 x stem_k conv + BN + ReLU of the raw signal before block 1
};

// ---- one dense contraction  out[M,N] = act(A_gather[M,K] @ W[K,N] + shift[N] (+ rank-1 residual)) -----------------
// A row m = output frame (b = m / t_out, to = m % t_out).  K is the concatenation of
//   part 0: `taps` taps of c0 channels: frame ti = to*stride0 + j - left (zero outside [0,t_in0)), read from src0 with
//           row stride lda0, or generated on the fly from the raw signal (gen != 0, block-1 conv2a, cnn.py:254):
//           a = relu((x*gw[c])*ginv[c] + gsh[c]);
//   part 1: c1 channels of src1 at frame to*stride1 of a t_in1-frame window (the 1x1 branch1 conv input,
//           cnn.py:251), row stride lda1.
struct GemmProblem {
    int M, N, K;              // K = taps*c0 + c1
    int t_out;
    int t_in0, stride0, taps, left, c0;
    int t_in1, stride1, c1;
    const float* src0; int lda0;
    const float* src1; int lda1;
    int gen;                  // 1: part 0 generated from x (rank-1 conv + BN + ReLU)
    const float* x;           // raw window samples: [B*t_in0] for gen, [B*t_inr] for the residual
    const float *gw, *ginv, *gsh;
    const float* W;           // [K,N] fp32, BN scale folded in (SIMT path)
    const float* shift;       // [N]
    int relu;
    int res, t_inr, strider;  // res=1: add rank-1 residual (x[to*strider]*rw[n])*rinv[n] + rsh[n] before the ReLU
    const float *rw, *rinv, *rsh;
    float* out; int ldo;
    int layer_id;             // which prepared tensor-core weight image belongs to this contraction
};

// ---- batch-statistics BatchNorm (cb_bn.cu): out = act(a*a_inv + a_sh [+ b*b_inv + b_sh | + b] [+ rank-1 branch]) --------
struct BnApplyArgs {
    const float *a, *a_inv, *a_sh;
    const float *b, *b_inv, *b_sh;          // b_inv == nullptr: b is added as it is (branch1 without BN)
    const float *x, *rw, *rinv, *rsh;       // rank-1 branch of the raw signal: (x[win*t_inr + to*strider]*rw)*rinv + rsh
    int t_out, t_inr, strider;
    int relu;
    float* out;                             // may alias a (every element is read and written by the same thread)
    long long M;
};

// ---- recurrences (fp32 path): both directions of one layer (grid.y = direction) -------------------------------------------------
struct LstmProblem {
    int B, T, H;
    const float* pre;         // [B*T, ld_pre] hoisted input projection + bias; direction d uses columns [d*4H, (d+1)*4H)
    int ld_pre;
    const float* whh[2];      // [H,4H] fp32 recurrent kernels (fw, bw)
    const int32_t* lens;      // [B]
    float* out; int ldo;      // h of direction d written to out[(b*T+t)*ldo + d*H + u]; zeros for t >= len
    int layer;
};

struct GruProblem {
    int B, T, H;
    const float* pre;         // [B*T, ld_pre]; direction d uses columns [d*3H, (d+1)*3H) = r | u | candidate
    int ld_pre;
    const float* wg[2];       // [H,2H] recurrent gate kernels (fw, bw), columns r | u
    const float* wc[2];       // [H,H] recurrent candidate kernels
    const int32_t* lens;      // [B]
    float* out; int ldo;      // as LstmProblem
    int layer;
};

// ---- stem convolution of the raw signal (RNA_model2 / RNA_model3, chiron/cnn.py:454-476; cb_stem_kernel.cuh) -----------------
// out[(b*t_out + to)*C + c] = act((sum_j x[b*t_in + to*stride + j - left] * w[j*C + c]) * inv[c] + shift[c]), 'SAME' padding.
struct StemProblem {
    const float* x; int B, t_in, t_out, k, stride, left, C;
    const float* w;                     // [k][C]
    const float *inv, *shift;           // nullptr: raw convolution output (batch-statistics BN normalises it afterwards)
    int relu;
    float* out;
};
struct CbStem { const float *w, *inv, *shift, *scale, *offset; };   // inv/shift: population BN folded; scale/offset: raw

// ---- weights of the residual blocks --------------------------------------------------------------------------------------
struct CbConvW { const float *W, *shift; };                 // population BN folded: W[K,N] * inv[n], shift[N]
struct CbRawConv { const float *W, *scale, *offset; };      // as in the checkpoint; scale == nullptr: no BN on this conv
