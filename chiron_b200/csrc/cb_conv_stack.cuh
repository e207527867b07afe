// Orchestration of the residual conv stack on the fp32 path (chiron/cnn.py:234-262 residual_layer, 380-389 DNA_model1,
// 555-566 rna_test): which contraction / BatchNorm pass runs on which ping-pong buffer, in the two BatchNorm modes.
// The work is issued through an `Ops` adapter -- CUDA launches in cb_api.cu, the host emulator in tests/cuda_emu -- so that
// the same source is exercised on a machine without a GPU.
#pragma once
#include <string.h>

#include "../../include/chiron_b200.h"
#include "cb_simt_types.h"

// Geometry of the stem convolution for an L-sample window ('SAME').
struct CbStemGeom { int t_out, left; };
inline CbStemGeom cb_stem_geom(const CbConfig& c, int L) {
    CbStemGeom g;
    g.t_out = (L + c.stem_stride - 1) / c.stem_stride;
    int pad = (g.t_out - 1) * c.stem_stride + c.stem_k - L; if (pad < 0) pad = 0;
    g.left = pad / 2;
    return g;
}

struct CbConvStackBufs {
    float* act[3];                  // ping-pong activations, each [B*L, C] fp32
    float* vec[CB_BN_VECS];         // [C] scratch vectors (batch-statistics mode)
    const float* zeros;             // [C] zero shift (batch-statistics mode)
};

// Residual conv stack with batch-statistics BN (HEAD's conv_layer -> simple_global_bn, chiron/cnn.py:65-68,166-188;
// residual_layer cnn.py:234-262): every convolution is run raw (nothing folded, no shift, no ReLU), its output is reduced
// to per-channel batch moments, and the normalisation + activation (+ the residual sum) is a separate pass.
// On return *feat is the block stack's output [B*T,C] (one of act[]) and *t_feat its frame count.
template <class Ops>
int cb_conv_stack_batch_bn(Ops& ops, const CbConfig& c, const CbRawConv* raw1, const CbRawConv* raw2a, const CbRawConv* raw2b,
                           const CbRawConv* raw2c, const CbStem* stem, const CbConvStackBufs& buf, const float* x, int B,
                           int L, const float** feat, int* t_feat) {
    const int C = c.channels;
    int rc;
    float* const* vec = buf.vec;
    int t_in = L;
    const float* X = nullptr;
    int xi = -1;
    const bool has_stem = c.stem_k > 0;
    if (has_stem) {          // stem conv raw -> act[2]; batch moments; BN + ReLU in place; the blocks then all read C channels
        const CbStemGeom sg = cb_stem_geom(c, L);
        StemProblem sp;
        memset(&sp, 0, sizeof(sp));
        sp.x = x; sp.B = B; sp.t_in = L; sp.t_out = sg.t_out; sp.k = c.stem_k; sp.stride = c.stem_stride; sp.left = sg.left;
        sp.C = C; sp.w = stem->w; sp.out = buf.act[2];
        if ((rc = ops.stem(sp)) != CB_OK) return rc;
        const long long M_s = (long long)B * sg.t_out;
        if ((rc = ops.bn_stats(buf.act[2], M_s, stem->scale, stem->offset, vec[0], vec[1])) != CB_OK) return rc;
        BnApplyArgs sa;
        memset(&sa, 0, sizeof(sa));
        sa.a = buf.act[2]; sa.a_inv = vec[0]; sa.a_sh = vec[1]; sa.relu = 1; sa.out = buf.act[2]; sa.M = M_s;
        if ((rc = ops.bn_apply(sa)) != CB_OK) return rc;
        X = buf.act[2]; xi = 2; t_in = sg.t_out;
    }
    auto raw_gemm = [&](GemmProblem& g, const float* W, float* out) {
        g.N = C; g.W = W; g.shift = buf.zeros; g.relu = 0; g.out = out; g.ldo = C;
        return ops.gemm(g);
    };
    for (int b = 0; b < c.n_blocks; ++b) {
        const int st = c.stride[b], k = c.k[b];
        const int t_out = (t_in + st - 1) / st;
        int pad = (t_out - 1) * st + k - t_in; if (pad < 0) pad = 0;     // TF 'SAME'
        const int left = pad / 2;
        int ia = (xi + 1) % 3, ib = (xi + 2) % 3;
        if (xi < 0) { ia = 0; ib = 1; }
        const long long M_in = (long long)B * t_in, M_out = (long long)B * t_out;
        GemmProblem g;
        BnApplyArgs ap;
        const bool rank1 = b == 0 && !has_stem;      // block 1 reads the one-channel signal itself
        // conv2a 1x1 + BN + ReLU
        if (rank1) {        // rank-1 in the raw signal: statistics from the samples, tensor generated inside conv2b's loader
            if ((rc = ops.bn_rank1(x, B, t_in, 1, t_in, raw2a[b].W, raw2a[b].scale, raw2a[b].offset,
                                   vec[0], vec[1])) != CB_OK) return rc;
        } else {
            memset(&g, 0, sizeof(g));
            g.M = (int)M_in; g.K = C; g.t_out = t_in; g.t_in0 = t_in; g.stride0 = 1; g.taps = 1; g.c0 = C; g.src0 = X; g.lda0 = C;
            if ((rc = raw_gemm(g, raw2a[b].W, buf.act[ia])) != CB_OK) return rc;
            if ((rc = ops.bn_stats(buf.act[ia], M_in, raw2a[b].scale, raw2a[b].offset, vec[0], vec[1])) != CB_OK) return rc;
            memset(&ap, 0, sizeof(ap));
            ap.a = buf.act[ia]; ap.a_inv = vec[0]; ap.a_sh = vec[1]; ap.relu = 1; ap.out = buf.act[ia]; ap.M = M_in;
            if ((rc = ops.bn_apply(ap)) != CB_OK) return rc;
        }
        // conv2b 1xk (stride) + BN + ReLU -> act[ib]
        memset(&g, 0, sizeof(g));
        g.M = (int)M_out; g.K = k * C; g.t_out = t_out; g.t_in0 = t_in; g.stride0 = st; g.taps = k; g.left = left; g.c0 = C;
        if (rank1) { g.gen = 1; g.x = x; g.gw = raw2a[b].W; g.ginv = vec[0]; g.gsh = vec[1]; }
        else { g.src0 = buf.act[ia]; g.lda0 = C; }
        if ((rc = raw_gemm(g, raw2b[b].W, buf.act[ib])) != CB_OK) return rc;
        if ((rc = ops.bn_stats(buf.act[ib], M_out, raw2b[b].scale, raw2b[b].offset, vec[2], vec[3])) != CB_OK) return rc;
        memset(&ap, 0, sizeof(ap));
        ap.a = buf.act[ib]; ap.a_inv = vec[2]; ap.a_sh = vec[3]; ap.relu = 1; ap.out = buf.act[ib]; ap.M = M_out;
        if ((rc = ops.bn_apply(ap)) != CB_OK) return rc;
        // conv2c 1x1 + BN -> act[ia] (raw), its inv/shift in vec[4], vec[5]
        memset(&g, 0, sizeof(g));
        g.M = (int)M_out; g.K = C; g.t_out = t_out; g.t_in0 = t_out; g.stride0 = 1; g.taps = 1; g.c0 = C; g.src0 = buf.act[ib]; g.lda0 = C;
        if ((rc = raw_gemm(g, raw2c[b].W, buf.act[ia])) != CB_OK) return rc;
        if ((rc = ops.bn_stats(buf.act[ia], M_out, raw2c[b].scale, raw2c[b].offset, vec[4], vec[5])) != CB_OK) return rc;
        // branch1: 1x1 conv (stride) of the block input (+ BN), then relu(branch1 + conv2c)
        memset(&ap, 0, sizeof(ap));
        ap.a = buf.act[ia]; ap.a_inv = vec[4]; ap.a_sh = vec[5]; ap.relu = 1; ap.out = buf.act[ia]; ap.M = M_out;
        if (rank1) {
            if ((rc = ops.bn_rank1(x, B, t_in, st, t_out, raw1[b].W, raw1[b].scale, raw1[b].offset,
                                   vec[6], vec[7])) != CB_OK) return rc;
            ap.x = x; ap.rw = raw1[b].W; ap.rinv = vec[6]; ap.rsh = vec[7]; ap.t_out = t_out; ap.t_inr = t_in; ap.strider = st;
        } else {
            memset(&g, 0, sizeof(g));     // conv2b's output (act[ib]) has been consumed: reuse it for the raw branch
            g.M = (int)M_out; g.K = C; g.t_out = t_out; g.t_in0 = t_in; g.stride0 = st; g.taps = 1; g.left = 0; g.c0 = C; g.src0 = X; g.lda0 = C;
            if ((rc = raw_gemm(g, raw1[b].W, buf.act[ib])) != CB_OK) return rc;
            ap.b = buf.act[ib];
            if (raw1[b].scale) {
                if ((rc = ops.bn_stats(buf.act[ib], M_out, raw1[b].scale, raw1[b].offset, vec[6], vec[7])) != CB_OK) return rc;
                ap.b_inv = vec[6]; ap.b_sh = vec[7];
            }
        }
        if ((rc = ops.bn_apply(ap)) != CB_OK) return rc;
        X = buf.act[ia]; xi = ia; t_in = t_out;
    }
    *feat = X; *t_feat = t_in;
    return CB_OK;
}

// Residual conv stack with population BN folded into the weights (the shipped checkpoints' graph: cnn.py:234-262,
// 380-389 with batchnorm() cnn.py:125-163).  On return *feat is the stack's output [B*T,C] and *t_feat its frame count.
template <class Ops>
int cb_conv_stack_folded(Ops& ops, const CbConfig& c, const CbConvW* conv2a, const CbConvW* conv2b, const CbConvW* convc,
                         const float* g_w, const float* g_inv, const float* g_sh, const float* r_w, const float* r_inv,
                         const float* r_sh, const CbStem* stem, const CbConvStackBufs& buf, const float* x, int B, int L,
                         const float** feat, int* t_feat) {
    const int C = c.channels;
    int rc;
    int t_in = L;
    const float* X = nullptr;          // block input (nullptr = raw signal for block 1)
    int xi = -1;                       // which act[] buffer holds X
    const bool has_stem = c.stem_k > 0;
    if (has_stem) {                    // stem conv + folded BN + ReLU -> act[2]; the blocks then all read C channels
        const CbStemGeom sg = cb_stem_geom(c, L);
        StemProblem sp;
        memset(&sp, 0, sizeof(sp));
        sp.x = x; sp.B = B; sp.t_in = L; sp.t_out = sg.t_out; sp.k = c.stem_k; sp.stride = c.stem_stride; sp.left = sg.left;
        sp.C = C; sp.w = stem->w; sp.inv = stem->inv; sp.shift = stem->shift; sp.relu = 1; sp.out = buf.act[2];
        if ((rc = ops.stem(sp)) != CB_OK) return rc;
        X = buf.act[2]; xi = 2; t_in = sg.t_out;
    }
    for (int b = 0; b < c.n_blocks; ++b) {
        const bool rank1 = b == 0 && !has_stem;       // block 1 reads the one-channel signal itself
        const int st = c.stride[b], k = c.k[b];
        const int t_out = (t_in + st - 1) / st;
        int pad = (t_out - 1) * st + k - t_in; if (pad < 0) pad = 0;     // TF 'SAME'
        const int left = pad / 2;
        int ia = (xi + 1) % 3, ib = (xi + 2) % 3;     // scratch buffers that are not X
        if (xi < 0) { ia = 0; ib = 1; }
        GemmProblem g;
        if (!rank1) {                  // conv2a 1x1 + BN + ReLU  -> act[ia]   (a rank-1 block 1 generates it on the fly)
            memset(&g, 0, sizeof(g));
            g.M = B * t_in; g.N = C; g.K = C; g.t_out = t_in;
            g.t_in0 = t_in; g.stride0 = 1; g.taps = 1; g.left = 0; g.c0 = C; g.src0 = X; g.lda0 = C;
            g.W = conv2a[b].W; g.shift = conv2a[b].shift; g.relu = 1; g.out = buf.act[ia]; g.ldo = C;
            if ((rc = ops.gemm(g)) != CB_OK) return rc;
        }
        // conv2b 1xk (stride) + BN + ReLU -> act[ib]
        memset(&g, 0, sizeof(g));
        g.M = B * t_out; g.N = C; g.K = k * C; g.t_out = t_out;
        g.t_in0 = t_in; g.stride0 = st; g.taps = k; g.left = left; g.c0 = C;
        if (rank1) { g.gen = 1; g.x = x; g.gw = g_w; g.ginv = g_inv; g.gsh = g_sh; }
        else { g.src0 = buf.act[ia]; g.lda0 = C; }
        g.W = conv2b[b].W; g.shift = conv2b[b].shift; g.relu = 1; g.out = buf.act[ib]; g.ldo = C;
        if ((rc = ops.gemm(g)) != CB_OK) return rc;
        // conv2c 1x1 + BN, + branch1 (1x1 conv of the block input, stride st), ReLU -> act[ia]
        memset(&g, 0, sizeof(g));
        g.M = B * t_out; g.N = C; g.t_out = t_out;
        g.t_in0 = t_out; g.stride0 = 1; g.taps = 1; g.left = 0; g.c0 = C; g.src0 = buf.act[ib]; g.lda0 = C;
        if (rank1) {
            g.K = C; g.res = 1; g.x = x; g.t_inr = t_in; g.strider = st; g.rw = r_w; g.rinv = r_inv; g.rsh = r_sh;
        } else {
            g.K = 2 * C; g.c1 = C; g.src1 = X; g.lda1 = C; g.t_in1 = t_in; g.stride1 = st;
        }
        g.W = convc[b].W; g.shift = convc[b].shift; g.relu = 1; g.out = buf.act[ia]; g.ldo = C;
        if ((rc = ops.gemm(g)) != CB_OK) return rc;
        X = buf.act[ia]; xi = ia; t_in = t_out;
    }
    *feat = X; *t_feat = t_in;
    return CB_OK;
}

