// Launcher of the stem convolution kernel (cb_stem_kernel.cuh).
#include "cb_internal.cuh"
#include "cb_stem_kernel.cuh"

int cb_launch_stem(cb_handle* h, const StemProblem& p, cudaStream_t s) {
    if (p.B <= 0 || p.t_out <= 0) return CB_OK;
    if ((p.C & 3) || p.k < 1 || p.stride < 1) { cb_set_error("stem: bad geometry"); return CB_ERR_ARG; }
    const int grid = cb_stem::stem_grid(h->sm_count, (long long)p.B * p.t_out * (p.C >> 2));
    cb_stem::stem_conv_kernel<<<grid, cb_stem::STEM_THREADS, 0, s>>>(p);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}
