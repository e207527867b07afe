// tcgen05 tensor-core path (CB_PREC_TC_SPLIT / CB_PREC_TC_FAST) -- placeholder until the kernels land.
#include "cb_internal.cuh"

int cb_tc_prepare(cb_handle*, const float*) { cb_set_error("tensor-core path not built yet"); return CB_ERR_ARG; }
void cb_tc_release(cb_handle*) {}
int cb_launch_gemm_tc(cb_handle*, const GemmProblem&, cudaStream_t) { cb_set_error("tensor-core path not built yet"); return CB_ERR_ARG; }
int cb_launch_lstm_tc(cb_handle*, const LstmProblem&, cudaStream_t) { cb_set_error("tensor-core path not built yet"); return CB_ERR_ARG; }
