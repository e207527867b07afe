// Forward pass in tensor-core mode: orchestration of the tcgen05 kernels over operand images.
// (chiron_model.inference, chiron/chiron_model.py:134-172: getcnnfeature -> rnn_layers -> logits.)
//
// Data flow (all activations as fp16 hi/lo k-group-plane images, see cb_tc_common.cuh):
//   x --transpose--> xT --gen_conv2a--> A0 --conv2b(1) (k taps, stride)--> P0 --conv2c(1)+rank-1 branch--> P1 = block-1 output
//   block n>=2:  X --conv2a--> A --conv2b (3 taps = 3 frame-shifted bulk loads)--> Bt --conv2c ++ branch1(X)--> X'
//   every image is time-major (row = frame*Bp + window), so the last block's output is the LSTM's A operand as is
//   LSTM layer l: image --input projection--> pre[T][8H][Bp] fp32 --recurrence--> h image (or fp32 out for the last layer)
//   head (time-major) -> logits[B][T][n_class]; path_prob.
#include <string.h>

#include <vector>

#include "cb_internal.cuh"

namespace {

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

struct TcWorkspace {
    void* base; size_t bytes;
    int B, L;                     // geometry the images were zeroed for
    CbImg a0;                     // block-1 conv2a output: L frames + the 'SAME' padding frames of conv2b
    CbImg conv[3];                // block activations: T frames + the zero frames the widest later block's 'SAME' padding
                                  // reads in front of and behind them (one each for width 3)
    CbImg himg;                   // LSTM layer output, planes [fw 13][bw 13]
    float* xT;                    // [L][Bp] transposed raw windows
    float* pre;                   // [T][8H][Bp]
    float* out;                   // [T][2H][Bp] (last layer)
    int fea_idx;                  // which conv image holds the CNN feature of the last forward
};

int ensure_ws(cb_handle* h, int B, int L, int T, int Bp, int pad0, int left0, cudaStream_t s) {
    TcWorkspace* w = (TcWorkspace*)h->tc_ws;
    if (!w) { w = new TcWorkspace(); memset(w, 0, sizeof(*w)); h->tc_ws = w; }
    const CbConfig& c = h->cfg;
    const int planes = c.channels / 8;
    const long long rows_a0 = c.stem_k > 0 ? Bp : (long long)(L + pad0) * Bp;      // (stem models have no rank-1 block)
    int pad_l = 1, pad_r = 1;                                 // TF 'SAME' of the blocks after the first (the stem): frames they
    for (int b = c.stem_k > 0 ? 0 : 1, t = T; b < c.n_blocks; ++b) {      // read in front of / behind their input's data
        const int to = (t + c.stride[b] - 1) / c.stride[b];
        int pad = (to - 1) * c.stride[b] + c.k[b] - t; if (pad < 0) pad = 0;
        const int l = pad / 2, r = pad - l;
        pad_l = l > pad_l ? l : pad_l; pad_r = r > pad_r ? r : pad_r;
        t = to;
    }
    const long long rows_c = (long long)(T + pad_l + pad_r) * Bp;
    const long long rows_h = (long long)T * Bp;
    const int hplanes = 32;       // 26 real k-group planes (2 x 13) + zero planes the K padding reads
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    size_t o_a0[2], o_conv[3][2], o_h[2];
    for (int j = 0; j < 2; ++j) o_a0[j] = carve(cb_img_halfs(rows_a0, planes) * 2);
    for (int i = 0; i < 3; ++i) for (int j = 0; j < 2; ++j) o_conv[i][j] = carve(cb_img_halfs(rows_c, planes) * 2);
    for (int j = 0; j < 2; ++j) o_h[j] = carve(cb_img_halfs(rows_h, hplanes) * 2);
    const size_t img_bytes = off;
    const size_t o_xt = carve((size_t)L * Bp * sizeof(float));
    const size_t o_pre = carve((size_t)T * 8 * c.hidden * Bp * sizeof(float));
    const size_t o_out = carve((size_t)T * 2 * c.hidden * Bp * sizeof(float));
    bool rezero = w->B != B || w->L != L;
    if (off > w->bytes) {
        if (w->base) { cudaFree(w->base); w->base = nullptr; w->bytes = 0; }
        cudaError_t e = cudaMalloc(&w->base, off);
        if (e != cudaSuccess) { cb_set_error("tensor-core workspace cudaMalloc(%zu bytes): %s", off, cudaGetErrorString(e)); return CB_ERR_NOMEM; }
        w->bytes = off;
        rezero = true;
    }
    char* base = (char*)w->base;
    w->a0.hi = (__half*)(base + o_a0[0]); w->a0.lo = (__half*)(base + o_a0[1]);
    w->a0.plane_rows = rows_a0; w->a0.row0 = (long long)left0 * Bp; w->a0.planes = planes;
    for (int i = 0; i < 3; ++i) {
        w->conv[i].hi = (__half*)(base + o_conv[i][0]); w->conv[i].lo = (__half*)(base + o_conv[i][1]);
        w->conv[i].plane_rows = rows_c; w->conv[i].row0 = (long long)pad_l * Bp; w->conv[i].planes = planes;
    }
    w->himg.hi = (__half*)(base + o_h[0]); w->himg.lo = (__half*)(base + o_h[1]);
    w->himg.plane_rows = rows_h; w->himg.row0 = 0; w->himg.planes = hplanes;
    w->xT = (float*)(base + o_xt); w->pre = (float*)(base + o_pre); w->out = (float*)(base + o_out);
    if (rezero) {     // padding frames, rows of windows >= B and unused planes must read as zero; the rest is rewritten
        CB_CUDA(cudaMemsetAsync(base, 0, img_bytes, s));
        w->B = B; w->L = L;
    }
    h->ws_bytes_tc = w->bytes;
    return CB_OK;
}

int timed_gemm(cb_handle* h, const TcGemm& g, cudaStream_t s, int cat) {
    const int pi = cb_prof_begin(h, cat, s);
    const int rc = cb_launch_gemm_tc(h, g, s);
    cb_prof_end(h, pi, s);
    return rc;
}

}  // namespace

void cb_forward_tc_release(cb_handle* h) {
    TcWorkspace* w = (TcWorkspace*)h->tc_ws;
    if (!w) return;
    if (w->base) cudaFree(w->base);
    delete w;
    h->tc_ws = nullptr;
}

int cb_forward_tc(cb_handle* h, const float* x, const int32_t* seq_len_out, int B, int L, float* logits,
                  float* path_prob, cudaStream_t s) {
    const CbConfig& c = h->cfg;
    const int C = c.channels, H = c.hidden;
    const bool stem = c.stem_k > 0;                           // RNA_model2/3: a strided stem conv, then stride-1 blocks only
    const int st0 = stem ? c.stem_stride : c.stride[0], k0 = stem ? c.stem_k : c.k[0];
    const int T = (L + st0 - 1) / st0;                        // frames after block 1 (the stem); later blocks have stride 1
    int pad0 = (T - 1) * st0 + k0 - L; if (pad0 < 0) pad0 = 0;   // TF 'SAME'
    const int left0 = pad0 / 2;
    const int Bp = (B + 127) / 128 * 128;
    int rc = ensure_ws(h, B, L, T, Bp, pad0, left0, s);
    if (rc != CB_OK) return rc;
    TcWorkspace* w = (TcWorkspace*)h->tc_ws;
    h->prof_n = 0;
    if (h->timing) CB_CUDA(cudaEventRecord(h->ev[0], s));

    const int cpt = C / 32;                                   // 32-channel k-chunks per tap
    auto base_gemm = [&](int layer_id) {
        TcGemm g;
        memset(&g, 0, sizeof(g));
        g.layer_id = layer_id; g.T = T; g.B = B; g.Bp = Bp; g.taps = 1; g.stride = 1; g.a0_chunks_per_tap = cpt;
        g.N = C; g.relu = 1; g.out_mode = 2;
        return g;
    };
    // ---- stem models: the stem convolution writes the operand image every block (the first included) then reads ----------
    int xi = 1;
    if (stem) {
        int pi = cb_prof_begin(h, CB_CAT_CONV, s);
        rc = cb_launch_transpose_x(h, x, B, L, Bp, w->xT, s);
        if (rc == CB_OK) rc = cb_launch_stem_image(h, w->xT, B, Bp, L, T, left0, w->conv[xi], s);
        cb_prof_end(h, pi, s);
        if (rc != CB_OK) return rc;
    } else
    // ---- block 1 (cnn.py:383-384): conv2a is a rank-1 function of the raw signal ----------------------------------------
    {
        int pi = cb_prof_begin(h, CB_CAT_CONV, s);
        rc = cb_launch_transpose_x(h, x, B, L, Bp, w->xT, s);
        if (rc == CB_OK) rc = cb_launch_gen_conv2a(h, w->xT, B, Bp, L, w->a0, s);
        cb_prof_end(h, pi, s);
        if (rc != CB_OK) return rc;
        TcGemm g = base_gemm(1);                                  // conv2b: k0 taps, stride st0
        g.a0 = w->a0; g.taps = k0; g.left = left0; g.stride = st0; g.shift = h->conv2b[0].shift; g.o = w->conv[0];
        if ((rc = timed_gemm(h, g, s, CB_CAT_CONV)) != CB_OK) return rc;
        g = base_gemm(2);                                         // conv2c + rank-1 branch1 of the raw signal
        g.a0 = w->conv[0]; g.shift = h->convc[0].shift;
        g.res = 1; g.xT = w->xT; g.res_stride = st0; g.rw = h->r_w; g.rinv = h->r_inv; g.rsh = h->r_sh;
        g.o = w->conv[1];
        if ((rc = timed_gemm(h, g, s, CB_CAT_CONV)) != CB_OK) return rc;
    }
    int Tc = T;                                                   // frames of the current block input
    for (int b = stem ? 0 : 1; b < c.n_blocks; ++b) {
        const int ai = (xi + 1) % 3, bi = (xi + 2) % 3;
        const int sb = c.stride[b], kb = c.k[b];
        const int To = (Tc + sb - 1) / sb;                        // TF 'SAME': ceil(T / stride) output frames
        int pad = (To - 1) * sb + kb - Tc; if (pad < 0) pad = 0;
        const int left = pad / 2, right = pad - left;
        TcGemm g = base_gemm(b * 4 + 0);                          // conv2a 1x1
        g.T = Tc; g.a0 = w->conv[xi]; g.shift = h->conv2a[b].shift; g.o = w->conv[ai];
        if ((rc = timed_gemm(h, g, s, CB_CAT_CONV)) != CB_OK) return rc;
        if (Tc < T && right > 0) {
            // After a strided block the images hold fewer frames than they did before: the frames conv2b's padding reads
            // behind the data are leftovers of an earlier, longer tensor.  Zero them (one strided memset per hi / lo).
            const CbImg& im = w->conv[ai];
            const size_t pitch = (size_t)im.plane_rows * 16, width = (size_t)right * Bp * 16;
            CB_CUDA(cudaMemset2DAsync((char*)im.hi + ((size_t)im.row0 + (size_t)Tc * Bp) * 16, pitch, 0, width, im.planes, s));
            CB_CUDA(cudaMemset2DAsync((char*)im.lo + ((size_t)im.row0 + (size_t)Tc * Bp) * 16, pitch, 0, width, im.planes, s));
        }
        g = base_gemm(b * 4 + 1);                                 // conv2b 1xk: frame-shifted (and strided) views of one image
        g.T = To; g.a0 = w->conv[ai]; g.taps = kb; g.left = left; g.stride = sb; g.shift = h->conv2b[b].shift; g.o = w->conv[bi];
        if ((rc = timed_gemm(h, g, s, CB_CAT_CONV)) != CB_OK) return rc;
        g = base_gemm(b * 4 + 2);                                 // conv2c ++ branch1(X) (1x1, the block's stride), ReLU
        g.T = To; g.a0 = w->conv[bi]; g.a1 = w->conv[xi]; g.a1_chunks = cpt; g.a1_stride = sb; g.shift = h->convc[b].shift;
        g.o = w->conv[ai];
        if ((rc = timed_gemm(h, g, s, CB_CAT_CONV)) != CB_OK) return rc;
        xi = ai;
        Tc = To;
    }
    const int Tf = Tc;                                            // frames the LSTM stack and the head see (= cb_out_len(L))
    w->fea_idx = xi;
    if (h->timing) CB_CUDA(cudaEventRecord(h->ev[1], s));

    // ---- BiLSTM stack -------------------------------------------------------------------------------------------------------
    for (int l = 0; l < c.n_layers; ++l) {
        const bool last = l == c.n_layers - 1;
        const int n_gemm = (l == 0 || c.rnn_layout == 0) ? 1 : 2;
        for (int d = 0; d < n_gemm; ++d) {
            TcGemm g;
            memset(&g, 0, sizeof(g));
            g.layer_id = 32 + l * 2 + d; g.T = Tf; g.B = B; g.Bp = Bp; g.taps = 1; g.stride = 1;
            if (l == 0) { g.a0 = w->conv[xi]; g.a0_chunks_per_tap = cpt; }
            else if (n_gemm == 1) { g.a0 = w->himg; g.a0_chunks_per_tap = 7; }                 // K' = 208 -> 7 chunks
            else { g.a0 = w->himg; g.a0_plane0 = d * 13; g.a0_chunks_per_tap = 4; }            // K' = 104 -> 4 chunks
            g.N = n_gemm == 1 ? 8 * H : 4 * H;
            g.shift = cb_tc_lstm_bias(h, l, d);                    // gate columns in unit-major order
            g.out_mode = 1; g.out = w->pre + (size_t)d * 4 * H * Bp; g.ldo = 8 * H;    // pre[T][8H/4][Bp][4]
            if ((rc = timed_gemm(h, g, s, CB_CAT_LSTM_IN)) != CB_OK) return rc;
        }
        LstmProblem lp;
        memset(&lp, 0, sizeof(lp));
        lp.B = B; lp.T = Tf; lp.H = H; lp.pre = w->pre; lp.ld_pre = Bp; lp.lens = seq_len_out; lp.out = w->out; lp.ldo = 2 * H;
        lp.layer = l;
        const int pi = cb_prof_begin(h, CB_CAT_LSTM_REC, s);
        rc = cb_launch_lstm_tc(h, lp, last ? nullptr : &w->himg, last ? 1 : 0, s);
        cb_prof_end(h, pi, s);
        if (rc != CB_OK) return rc;
    }
    if (h->timing) CB_CUDA(cudaEventRecord(h->ev[2], s));

    {
        const int pi = cb_prof_begin(h, CB_CAT_HEAD, s);
        rc = cb_launch_head_tmajor(h, w->out, B, Bp, Tf, logits, s);
        cb_prof_end(h, pi, s);
        if (rc != CB_OK) return rc;
    }
    if (path_prob && (rc = cb_launch_path_prob(h, logits, B, Tf, path_prob, s)) != CB_OK) return rc;
    if (h->timing) { CB_CUDA(cudaEventRecord(h->ev[3], s)); h->have_ms = 1; }
    h->last_B = B; h->last_T = Tf; h->last_Bp = Bp; h->last_tmajor = 1;
    return CB_OK;
}

// Reconstruct fp32 tensors from the operand images for tests: what 0 = CNN feature [B][T][C]; n_layers = last LSTM
// output [B][T][2H] (fp32 time-major buffer); n_layers-1 = previous layer (h image).
long long cb_debug_fetch_tc(cb_handle* h, int what, float* dst, size_t max_floats) {
    TcWorkspace* w = (TcWorkspace*)h->tc_ws;
    if (!w || !h->last_B) { cb_set_error("cb_debug_fetch: nothing to fetch"); return CB_ERR_ARG; }
    const size_t B = h->last_B, T = h->last_T, Bp = h->last_Bp, C = h->cfg.channels, H = h->cfg.hidden;
    CB_CUDA(cudaDeviceSynchronize());
    auto fetch_img = [&](const CbImg& img, size_t width, auto plane_of, float* out) -> int {
        const size_t n = cb_img_halfs(img.plane_rows, img.planes);
        std::vector<__half> hi(n), lo(n);
        CB_CUDA(cudaMemcpy(hi.data(), img.hi, n * sizeof(__half), cudaMemcpyDeviceToHost));
        CB_CUDA(cudaMemcpy(lo.data(), img.lo, n * sizeof(__half), cudaMemcpyDeviceToHost));
        for (size_t b = 0; b < B; ++b)
            for (size_t t = 0; t < T; ++t)
                for (size_t ch = 0; ch < width; ++ch) {
                    size_t plane, e;
                    plane_of(ch, plane, e);
                    const size_t i = (plane * img.plane_rows + img.row0 + t * Bp + b) * 8 + e;
                    out[(b * T + t) * width + ch] = __half2float(hi[i]) + __half2float(lo[i]);
                }
        return CB_OK;
    };
    if (what == 0) {
        if (B * T * C > max_floats) { cb_set_error("cb_debug_fetch: destination too small"); return CB_ERR_ARG; }
        int rc = fetch_img(w->conv[w->fea_idx], C, [](size_t ch, size_t& pl, size_t& e) { pl = ch / 8; e = ch % 8; }, dst);
        return rc == CB_OK ? (long long)(B * T * C) : rc;
    }
    if (B * T * 2 * H > max_floats) { cb_set_error("cb_debug_fetch: destination too small"); return CB_ERR_ARG; }
    if (what == h->cfg.n_layers) {
        std::vector<float> tmp(T * 2 * H * Bp);
        CB_CUDA(cudaMemcpy(tmp.data(), w->out, tmp.size() * sizeof(float), cudaMemcpyDeviceToHost));
        for (size_t b = 0; b < B; ++b)
            for (size_t t = 0; t < T; ++t)
                for (size_t u = 0; u < 2 * H; ++u) {               // out[T][2*25][Bp][4]
                    const size_t d = u / H, uu = u % H;
                    dst[(b * T + t) * 2 * H + u] = tmp[(((t * 2 + d) * (H / 4) + uu / 4) * Bp + b) * 4 + (uu & 3)];
                }
        return (long long)(B * T * 2 * H);
    }
    if (what == h->cfg.n_layers - 1 && what >= 1) {
        int rc = fetch_img(w->himg, 2 * H, [H](size_t ch, size_t& pl, size_t& e) {
            const size_t d = ch / H, u = ch % H; pl = d * 13 + u / 8; e = u % 8; }, dst);
        return rc == CB_OK ? (long long)(B * T * 2 * H) : rc;
    }
    cb_set_error("cb_debug_fetch: tensor %d is not retained in tensor-core mode", what);
    return CB_ERR_ARG;
}
