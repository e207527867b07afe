// Launchers of the batch-statistics BatchNorm kernels (cb_bn_kernels.cuh; CB_BN_BATCH, fp32 path only).
#include "cb_internal.cuh"
#include "cb_bn_kernels.cuh"

using namespace cb_bn;

// inv/shift of BN(w[c] * x[win*t_in + to*stride]) over win < B, to < t_out (scale == nullptr: identity).
int cb_launch_bn_rank1(cb_handle* h, const float* x, int B, int t_in, int stride, int t_out, const float* w,
                       const float* scale, const float* offset, float* inv, float* shift, cudaStream_t s) {
    const int C = h->cfg.channels;
    int n_part = 0;
    if (scale) {
        n_part = bn_grid(h->sm_count, (long long)B * t_out, BN_THREADS, BN_STATS_CTAS);
        bn_x_stats_kernel<<<n_part, BN_THREADS, 0, s>>>(x, B, t_in, stride, t_out, h->bn_part);
        CB_CHECK_LAUNCH();
        h->launches++;
    }
    bn_finalize_kernel<<<(C + 31) / 32, BN_FIN_THREADS, 0, s>>>(h->bn_part, n_part, C, (double)B * t_out, w, scale, offset, inv, shift);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}

// inv/shift of BN(X[M,C]) with the statistics of X itself.
int cb_launch_bn_stats(cb_handle* h, const float* X, long long M, const float* scale, const float* offset, float* inv,
                       float* shift, cudaStream_t s) {
    const int C = h->cfg.channels;
    if ((C & 3) || C > 1024) { cb_set_error("batch-statistics BatchNorm needs channels %% 4 == 0 and <= 1024"); return CB_ERR_ARG; }
    const int rpp = BN_THREADS / (C >> 2);
    const int n_part = bn_grid(h->sm_count, M, rpp, BN_STATS_CTAS);
    bn_col_stats_kernel<<<n_part, BN_THREADS, 0, s>>>(X, M, C, h->bn_part);
    CB_CHECK_LAUNCH();
    h->launches++;
    bn_finalize_kernel<<<(C + 31) / 32, BN_FIN_THREADS, 0, s>>>(h->bn_part, n_part, C, (double)M, nullptr, scale, offset, inv, shift);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}

int cb_launch_bn_apply(cb_handle* h, const BnApplyArgs& a, cudaStream_t s) {
    const BnApply p = bn_apply_params(a, h->cfg.channels);
    if (p.M <= 0) return CB_OK;
    const int grid = bn_grid(h->sm_count, (p.M * (p.C >> 2) + BN_APPLY_U - 1) / BN_APPLY_U, BN_THREADS, BN_APPLY_CTAS);
    bn_apply_kernel<<<grid, BN_THREADS, 0, s>>>(p);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}
