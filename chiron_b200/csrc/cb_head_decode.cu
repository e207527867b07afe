// Launchers of the logit head, path_prob, seq_len and greedy CTC kernels (cb_head_decode_kernels.cuh).
#include "cb_internal.cuh"
#include "cb_head_decode_kernels.cuh"

using namespace cb_hd;


int cb_launch_head(cb_handle* h, const float* lasth, int M, float* logits, cudaStream_t s) {
    if (M <= 0) return CB_OK;
    if (h->cfg.n_class > MAX_CLASS) { cb_set_error("head: n_class > %d", MAX_CLASS); return CB_ERR_ARG; }
    long long blocks = ((long long)M + 7) / 8;
    const long long cap = (long long)h->sm_count * 32;
    if (blocks > cap) blocks = cap;
    head_kernel<<<(unsigned)blocks, 256, 0, s>>>(lasth, M, h->cfg.hidden, h->cfg.n_class, h->head_w, h->head_b,
                                                 h->head_wc, h->head_bc, logits);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}

int cb_launch_head_tmajor(cb_handle* h, const float* out_t, int B, int Bp, int T, float* logits, cudaStream_t s) {
    if (B <= 0 || T <= 0) return CB_OK;
    if (h->cfg.n_class > MAX_CLASS || T > 65535) { cb_set_error("head: unsupported shape"); return CB_ERR_ARG; }
    head_tmajor_kernel<<<dim3((B + 127) / 128, T), 128, 0, s>>>(out_t, B, Bp, T, h->cfg.hidden, h->cfg.n_class, h->head_w,
                                                               h->head_b, h->head_wc, h->head_bc, logits);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}

int cb_launch_path_prob(cb_handle* h, const float* logits, int B, int T, float* prob, cudaStream_t s) {
    if (B <= 0) return CB_OK;
    path_prob_kernel<<<(B + 3) / 4, 128, 0, s>>>(logits, B, T, h->cfg.n_class, prob);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}

int cb_launch_seq_len(cb_handle* h, const int32_t* in, int B, int L, int T, int32_t* out, cudaStream_t s) {
    if (B <= 0) return CB_OK;
    seq_len_kernel<<<(B + 255) / 256, 256, 0, s>>>(in, B, L, T, out);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}

int cb_launch_greedy(cb_handle* h, const float* logits, const int32_t* lens, int B, int T, int8_t* bases,
                     int32_t* n_bases, cudaStream_t s) {
    if (B <= 0) return CB_OK;
    greedy_kernel<<<(B + 3) / 4, 128, 0, s>>>(logits, lens, B, T, h->cfg.n_class, bases, n_bases);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}
