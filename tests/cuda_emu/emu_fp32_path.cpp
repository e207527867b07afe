// Host build of the fp32 (SIMT) path for tests/test_cuda_emu.py -- TEST INFRASTRUCTURE ONLY.
// Compiles the SAME kernel and orchestration sources the CUDA library is built from (cb_gemm_simt_kernel.cuh,
// cb_bn_kernels.cuh, cb_conv_stack.cuh, cb_lstm_simt_kernel.cuh, cb_gru_simt_kernel.cuh) against the host emulation in
// cuda_emu.h.
#define CB_HOST_EMU 1
#include "cuda_emu.h"

#include "../../include/chiron_b200.h"
#include "../../chiron_b200/csrc/cb_simt_types.h"
#include "../../chiron_b200/csrc/cb_gemm_simt_kernel.cuh"
#include "../../chiron_b200/csrc/cb_bn_kernels.cuh"
#include "../../chiron_b200/csrc/cb_stem_kernel.cuh"
#include "../../chiron_b200/csrc/cb_conv_stack.cuh"
#include "../../chiron_b200/csrc/cb_gru_simt_kernel.cuh"
#include "../../chiron_b200/csrc/cb_lstm_simt_kernel.cuh"
#include "../../chiron_b200/csrc/cb_head_decode_kernels.cuh"
#include "../../chiron_b200/csrc/cb_seq_kernels.cuh"

namespace {

struct EmuOps {          // mirrors the launchers of cb_gemm_simt.cu / cb_bn.cu: same checks, same grid plans
    int sm_count, C;
    std::vector<double> part;
    EmuOps(int sms, int c) : sm_count(sms), C(c), part((size_t)CB_BN_MAX_PART * 2 * c) {}
    int gemm(const GemmProblem& p) {
        if (p.M <= 0) return CB_OK;
        if ((p.c0 & 3) || (p.c1 & 3) || (p.N & 3) || (p.lda0 & 3) || (p.lda1 & 3) || (p.ldo & 3)) return CB_ERR_ARG;
        const long long blocks = (long long)((p.N + cb_simt::BN - 1) / cb_simt::BN) * ((p.M + cb_simt::BM - 1) / cb_simt::BM);
        emu::launch((unsigned)blocks, cb_simt::NT, [&] { cb_simt::gemm_simt_kernel(p); });
        return CB_OK;
    }
    int bn_rank1(const float* x, int B, int t_in, int stride, int t_out, const float* w, const float* scale,
                 const float* offset, float* inv, float* shift) {
        int n_part = 0;
        if (scale) {
            n_part = cb_bn::bn_grid(sm_count, (long long)B * t_out, cb_bn::BN_THREADS, cb_bn::BN_STATS_CTAS);
            emu::launch(n_part, cb_bn::BN_THREADS, [&] { cb_bn::bn_x_stats_kernel(x, B, t_in, stride, t_out, part.data()); });
        }
        emu::launch((C + 31) / 32, cb_bn::BN_FIN_THREADS, [&] {
            cb_bn::bn_finalize_kernel(part.data(), n_part, C, (double)B * t_out, w, scale, offset, inv, shift);
        });
        return CB_OK;
    }
    int bn_stats(const float* X, long long M, const float* scale, const float* offset, float* inv, float* shift) {
        if ((C & 3) || C > 1024) return CB_ERR_ARG;
        const int n_part = cb_bn::bn_grid(sm_count, M, cb_bn::BN_THREADS / (C >> 2), cb_bn::BN_STATS_CTAS);
        emu::launch(n_part, cb_bn::BN_THREADS, [&] { cb_bn::bn_col_stats_kernel(X, M, C, part.data()); });
        emu::launch((C + 31) / 32, cb_bn::BN_FIN_THREADS, [&] {
            cb_bn::bn_finalize_kernel(part.data(), n_part, C, (double)M, nullptr, scale, offset, inv, shift);
        });
        return CB_OK;
    }
    int stem(const StemProblem& p) {
        if (p.B <= 0 || p.t_out <= 0) return CB_OK;
        if ((p.C & 3) || p.k < 1 || p.stride < 1) return CB_ERR_ARG;
        const int grid = cb_stem::stem_grid(sm_count, (long long)p.B * p.t_out * (p.C >> 2));
        emu::launch(grid, cb_stem::STEM_THREADS, [&] { cb_stem::stem_conv_kernel(p); });
        return CB_OK;
    }
    int bn_apply(const BnApplyArgs& a) {
        if (a.M <= 0) return CB_OK;
        const cb_bn::BnApply p = cb_bn::bn_apply_params(a, C);
        const int grid = cb_bn::bn_grid(sm_count, (p.M * (p.C >> 2) + cb_bn::BN_APPLY_U - 1) / cb_bn::BN_APPLY_U, cb_bn::BN_THREADS, cb_bn::BN_APPLY_CTAS);
        emu::launch(grid, cb_bn::BN_THREADS, [&] { cb_bn::bn_apply_kernel(p); });
        return CB_OK;
    }
};

}  // namespace

// geom: n_blocks, channels, branch1_bn_mask, k[8], stride[8], stem_k, stem_stride.
// stem[0..4] = W[k,C], folded inv, folded shift, scale, offset of the stem convolution (stem_k > 0).
// bn_mode CB_BN_BATCH:      w[(b*4 + i)*3 + {0,1,2}] = W, scale, offset of block b's conv i (branch1, conv2a, conv2b, conv2c).
// bn_mode CB_BN_POPULATION: w[(b*3 + i)*2 + {0,1}] = folded W, shift of conv2a / conv2b / conv2c(++branch1);
//                           rank1[0..5] = g_w, g_inv, g_sh, r_w, r_inv, r_sh of block 1.
// out[B*T*C] receives the stack's output; returns T (>0) or a negative CB_ERR_* code.
extern "C" int emu_conv_stack(int bn_mode, const int* geom, const float* const* w, const float* const* rank1,
                              const float* const* stem, const float* x, int B, int L, int sm_count, float* out,
                              long long* launches) {
    CbConfig c;
    memset(&c, 0, sizeof(c));
    c.n_blocks = geom[0]; c.channels = geom[1]; c.branch1_bn_mask = geom[2];
    for (int i = 0; i < CB_MAX_BLOCKS; ++i) { c.k[i] = geom[3 + i]; c.stride[i] = geom[3 + CB_MAX_BLOCKS + i]; }
    c.stem_k = geom[3 + 2 * CB_MAX_BLOCKS]; c.stem_stride = geom[4 + 2 * CB_MAX_BLOCKS];
    CbStem st;
    memset(&st, 0, sizeof(st));
    if (c.stem_k > 0) st = CbStem{stem[0], stem[1], stem[2], stem[3], stem[4]};
    const int C = c.channels;
    std::vector<float> act[3], vec((size_t)CB_BN_VECS * C), zeros((size_t)C, 0.f);
    CbConvStackBufs bufs;
    for (int i = 0; i < 3; ++i) { act[i].assign((size_t)B * L * C, -777.f); bufs.act[i] = act[i].data(); }
    for (int i = 0; i < CB_BN_VECS; ++i) bufs.vec[i] = vec.data() + (size_t)i * C;
    bufs.zeros = zeros.data();
    EmuOps ops(sm_count, C);
    const float* feat = nullptr;
    int T = 0, rc;
    const long long l0 = emu::launches;
    if (bn_mode == CB_BN_BATCH) {
        CbRawConv raw[4][CB_MAX_BLOCKS];
        for (int b = 0; b < c.n_blocks; ++b)
            for (int i = 0; i < 4; ++i) raw[i][b] = CbRawConv{w[(b * 4 + i) * 3], w[(b * 4 + i) * 3 + 1], w[(b * 4 + i) * 3 + 2]};
        rc = cb_conv_stack_batch_bn(ops, c, raw[0], raw[1], raw[2], raw[3], &st, bufs, x, B, L, &feat, &T);
    } else {
        CbConvW cw[3][CB_MAX_BLOCKS];
        for (int b = 0; b < c.n_blocks; ++b)
            for (int i = 0; i < 3; ++i) cw[i][b] = CbConvW{w[(b * 3 + i) * 2], w[(b * 3 + i) * 2 + 1]};
        rc = cb_conv_stack_folded(ops, c, cw[0], cw[1], cw[2], rank1[0], rank1[1], rank1[2], rank1[3], rank1[4], rank1[5],
                                  &st, bufs, x, B, L, &feat, &T);
    }
    if (rc != CB_OK) return rc;
    memcpy(out, feat, sizeof(float) * (size_t)B * T * C);
    if (launches) *launches = emu::launches - l0;
    return T;
}

// The reduction kernels on their own (odd channel counts, more partials than rows, strided sample walks).
extern "C" int emu_bn_stats(const float* X, long long M, int C, int sm_count, const float* scale, const float* offset,
                            float* inv, float* shift) {
    EmuOps ops(sm_count, C);
    return ops.bn_stats(X, M, scale, offset, inv, shift);
}

extern "C" int emu_bn_rank1(const float* x, int B, int t_in, int stride, int t_out, int C, int sm_count, const float* w,
                            const float* scale, const float* offset, float* inv, float* shift) {
    EmuOps ops(sm_count, C);
    return ops.bn_rank1(x, B, t_in, stride, t_out, w, scale, offset, inv, shift);
}

// The GRU recurrence (cb_gru_simt_kernel.cuh) for both directions of one layer, RG = 1 / 2 / 4 row groups per CTA.
extern "C" int emu_gru(int rg, int B, int T, int H, const float* pre, int ld_pre, const float* wg_fw, const float* wg_bw,
                       const float* wc_fw, const float* wc_bw, const int32_t* lens, float* out, int ldo) {
    GruProblem p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.T = T; p.H = H; p.pre = pre; p.ld_pre = ld_pre; p.wg[0] = wg_fw; p.wg[1] = wg_bw; p.wc[0] = wc_fw; p.wc[1] = wc_bw;
    p.lens = lens; p.out = out; p.ldo = ldo;
    if (H > 128 || (H & 3)) return CB_ERR_ARG;
    const int R = cb_gru::RPT * rg;
    const size_t smem = cb_gru::gru_smem_bytes(H, rg);
    const unsigned gx = (B + R - 1) / R;
    if (rg == 1) emu::launch2d(gx, 2, 128, smem, [&] { cb_gru::gru_simt_kernel<1>(p); });
    else if (rg == 2) emu::launch2d(gx, 2, 256, smem, [&] { cb_gru::gru_simt_kernel<2>(p); });
    else if (rg == 4) emu::launch2d(gx, 2, 512, smem, [&] { cb_gru::gru_simt_kernel<4>(p); });
    else return CB_ERR_ARG;
    return CB_OK;
}

// The LSTM recurrence (cb_lstm_simt_kernel.cuh) for both directions of one layer.
extern "C" int emu_lstm(int rg, int B, int T, int H, const float* pre, int ld_pre, const float* whh_fw, const float* whh_bw,
                        const int32_t* lens, float* out, int ldo) {
    LstmProblem p;
    memset(&p, 0, sizeof(p));
    p.B = B; p.T = T; p.H = H; p.pre = pre; p.ld_pre = ld_pre; p.whh[0] = whh_fw; p.whh[1] = whh_bw;
    p.lens = lens; p.out = out; p.ldo = ldo;
    if (H > 128 || (H & 3)) return CB_ERR_ARG;
    const int R = cb_lstm::RPT * rg;
    const size_t smem = cb_lstm::lstm_smem_bytes(H, rg);
    const unsigned gx = (B + R - 1) / R;
    if (rg == 1) emu::launch2d(gx, 2, 128, smem, [&] { cb_lstm::lstm_simt_kernel<1>(p); });
    else if (rg == 2) emu::launch2d(gx, 2, 256, smem, [&] { cb_lstm::lstm_simt_kernel<2>(p); });
    else if (rg == 4) emu::launch2d(gx, 2, 512, smem, [&] { cb_lstm::lstm_simt_kernel<4>(p); });
    else return CB_ERR_ARG;
    return CB_OK;
}

// Logit head (row-major and time-major forms), path_prob, seq_len scaling and greedy CTC (cb_head_decode_kernels.cuh), with
// the grids of their launchers in cb_head_decode.cu.
extern "C" int emu_head(const float* lasth, long long M, int H, int C, const float* w, const float* bias, const float* wc,
                        const float* bc, float* logits, int blocks) {
    emu::launch(blocks, 256, [&] { cb_hd::head_kernel(lasth, M, H, C, w, bias, wc, bc, logits); });
    return CB_OK;
}
extern "C" int emu_head_tmajor(const float* lasth, int B, int Bp, int T, int H, int C, const float* w, const float* bias,
                               const float* wc, const float* bc, float* logits) {
    emu::launch2d((B + 127) / 128, T, 128, 0, [&] { cb_hd::head_tmajor_kernel(lasth, B, Bp, T, H, C, w, bias, wc, bc, logits); });
    return CB_OK;
}
extern "C" int emu_path_prob(const float* logits, int B, int T, int C, float* prob) {
    emu::launch((B + 3) / 4, 128, [&] { cb_hd::path_prob_kernel(logits, B, T, C, prob); });
    return CB_OK;
}
extern "C" int emu_seq_len(const int32_t* in, int B, int L, int T, int32_t* out) {
    emu::launch((B + 255) / 256, 256, [&] { cb_hd::seq_len_kernel(in, B, L, T, out); });
    return CB_OK;
}
extern "C" int emu_greedy(const float* logits, const int32_t* lens, int B, int T, int C, int8_t* bases, int32_t* n_bases) {
    emu::launch((B + 3) / 4, 128, [&] { cb_hd::greedy_kernel(logits, lens, B, T, C, bases, n_bases); });
    return CB_OK;
}

// CTC beam search: the warp-cooperative shared-memory kernel (warp = 1: logits from global memory, 2: staged in shared
// memory; 16-bit trie nodes) or the thread-per-window kernel over global workspaces (warp = 0).  Returns 1 when a window
// outgrew the pool (the cooperative kernels then report it as n_bases = -1, beam_kernel as an empty read), 0 when every
// window decoded, or a negative CB_ERR_* code.
static int emu_beam_impl(int warp, const float* logits, const int32_t* lens, int B, int T, int C, int W, int pool,
                         int8_t* bases, int32_t* n_bases, float* scores) {
    int overflow = 0;
    if (pool < 2 * W + 2) return CB_ERR_ARG;
    if (warp) {
        const bool staged = warp == 2;
        const size_t stride = cb_seq::beam_warp_stride(T, C, W, pool, staged);
        const unsigned grid = (B + cb_seq::BEAM_WARPS - 1) / cb_seq::BEAM_WARPS;
        if (staged)
            emu::launch2d(grid, 1, cb_seq::BEAM_WARPS * 32, stride * cb_seq::BEAM_WARPS, [&] {
                cb_seq::beam_warp_kernel<true>(logits, lens, B, T, C, W, pool, (int)stride, bases, n_bases, &overflow, scores);
            });
        else
            emu::launch2d(grid, 1, cb_seq::BEAM_WARPS * 32, stride * cb_seq::BEAM_WARPS, [&] {
                cb_seq::beam_warp_kernel<false>(logits, lens, B, T, C, W, pool, (int)stride, bases, n_bases, &overflow, scores);
            });
    } else {
        const size_t stride = cb_seq::align_up(cb_beam_work_bytes(W, pool), 16);
        std::vector<char> ws(stride * (size_t)B + 64);
        emu::launch((B + 63) / 64, 64, [&] {
            cb_seq::beam_kernel(logits, lens, B, T, C, W, pool, ws.data(), stride, bases, n_bases, &overflow, nullptr, 0, scores);
        });
    }
    return overflow;
}

extern "C" int emu_beam(int warp, const float* logits, const int32_t* lens, int B, int T, int C, int W, int pool,
                        int8_t* bases, int32_t* n_bases) {
    return emu_beam_impl(warp, logits, lens, B, T, C, W, pool, bases, n_bases, nullptr);
}
// ... also returning the top path's log probability per window (cb_decode_beam_scored)
extern "C" int emu_beam_scored(int warp, const float* logits, const int32_t* lens, int B, int T, int C, int W, int pool,
                               int8_t* bases, int32_t* n_bases, float* scores) {
    return emu_beam_impl(warp, logits, lens, B, T, C, W, pool, bases, n_bases, scores);
}

// Overlap assembly of one read: the five kernels of cb_launch_assemble with its grids and workspace plan.
extern "C" int emu_assemble(const int8_t* bases, const int32_t* n_bases, const float* path_prob, int n_windows, int T, int jump,
                            int L, int kernel, int8_t* consensus, char* qual, int32_t* pos, int32_t* out_len, int max_len) {
    if (n_windows == 0) { *out_len = 0; return CB_OK; }
    const cb_seq::AsmPlan plan = cb_seq::asm_plan(n_windows, T, kernel, max_len);
    std::vector<char> buf(plan.total + 256, 0);
    char* base = buf.data() + (256 - reinterpret_cast<uintptr_t>(buf.data()) % 256) % 256;
    cb_seq::AsmWork w = cb_seq::asm_work(base, plan);
    const int ml = (int)plan.ml;
    emu::launch(1, 1024, [&] { cb_seq::asm_compact_kernel(n_bases, n_windows, T, w, pos); });
    const double jsr = (double)jump / (double)L;
    emu::launch((n_windows + 127) / 128, 128, [&] { cb_seq::asm_disp_kernel(bases, n_bases, T, kernel, jsr, w); });
    emu::launch(1, 1024, [&] { cb_seq::asm_scan_kernel(n_bases, w, pos, out_len, max_len); });
    emu::launch(n_windows, 128, [&] { cb_seq::asm_vote_kernel(bases, n_bases, path_prob, T, w, ml); });
    if (max_len > 0) emu::launch((max_len + 255) / 256, 256, [&] { cb_seq::asm_finish_kernel(w, ml, consensus, qual); });
    return CB_OK;
}

// One plain contraction out[M,N] = act(A[M,K] @ W[K,N] + shift[N]) through gemm_simt_kernel (the hoisted LSTM input projection).
extern "C" int emu_gemm(const float* A, int lda, const float* W, const float* shift, int M, int N, int K, int relu, float* out,
                        int ldo) {
    GemmProblem g;
    memset(&g, 0, sizeof(g));
    g.M = M; g.N = N; g.K = K; g.t_out = M; g.t_in0 = M; g.stride0 = 1; g.taps = 1; g.c0 = K; g.src0 = A; g.lda0 = lda;
    g.W = W; g.shift = shift; g.relu = relu; g.out = out; g.ldo = ldo;
    EmuOps ops(1, 4);
    return ops.gemm(g);
}

// The node pool cb_launch_beam gives the shared-memory beam search (cb_seq_kernels.cuh: beam_small_pool).
extern "C" long long emu_beam_small_pool(int T, int W) { return cb_seq::beam_small_pool(T, W); }

// The three passes of cb_launch_beam, exactly as it enqueues them: beam_warp_kernel with a first-pass pool of `pool_first`
// nodes marks the windows that outgrow it (n_bases = -1), beam_retry_kernel redoes those alone with a pool of `pool_retry`
// nodes (0 = the launcher's, what a CTA's shared memory holds), beam_kernel redoes what is still marked over `n_slots`
// atomically claimed global workspaces that cannot overflow.  marked[0..2] receive the number of windows marked after each
// pass (marked[2] > 0 only if the slots ran out).  Returns the sticky error flag of the last pass.
extern "C" int emu_beam_passes(const float* logits, const int32_t* lens, int B, int T, int C, int W, int pool_first,
                               int pool_retry, int n_slots, int8_t* bases, int32_t* n_bases, int* marked) {
    int flag = 0, error = 0, slots = 0;
    auto count = [&]() { int m = 0; for (int b = 0; b < B; ++b) m += n_bases[b] == -1; return m; };
    const size_t stride = cb_seq::beam_warp_stride(T, C, W, pool_first, false);
    emu::launch2d((B + cb_seq::BEAM_WARPS - 1) / cb_seq::BEAM_WARPS, 1, cb_seq::BEAM_WARPS * 32, stride * cb_seq::BEAM_WARPS, [&] {
        cb_seq::beam_warp_kernel<false>(logits, lens, B, T, C, W, pool_first, (int)stride, bases, n_bases, &flag, nullptr);
    });
    marked[0] = count();
    if ((marked[0] > 0) != (flag != 0)) return -100;          // marks without the flag (or the reverse) would be a bug
    const long long pool_r = pool_retry > 0 ? pool_retry : cb_seq::beam_retry_pool(T, W);
    const size_t smem_r = cb_seq::align_up(cb_beam_work_bytes<cb_seq::BeamIdx>(W, (int)pool_r), 16);
    emu::launch2d(B, 1, 32, smem_r, [&] {
        cb_seq::beam_retry_kernel(logits, lens, B, T, C, W, (int)pool_r, bases, n_bases, &flag, nullptr);
    });
    marked[1] = count();
    const int cap = 2 * W * (T + 1) + 2;
    const size_t stride_g = cb_seq::align_up(cb_beam_work_bytes(W, cap), 16);
    std::vector<char> ws(stride_g * (size_t)(n_slots > 0 ? n_slots : 1) + 64);
    emu::launch((B + 63) / 64, 64, [&] {
        cb_seq::beam_kernel(logits, lens, B, T, C, W, cap, ws.data(), stride_g, bases, n_bases, &error, &slots, n_slots, nullptr);
    });
    marked[2] = count();
    return error;
}
