// beam search + assembly kernels -- placeholder
#include "cb_internal.cuh"
int cb_launch_beam(cb_handle*, const float*, const int32_t*, int, int, int, int8_t*, int32_t*, cudaStream_t) { cb_set_error("beam search kernel not built yet"); return CB_ERR_ARG; }
int cb_launch_assemble(cb_handle*, const int8_t*, const int32_t*, const float*, int, int, int, int, int, int8_t*, char*, int32_t*, int32_t*, int, cudaStream_t) { cb_set_error("assembly kernel not built yet"); return CB_ERR_ARG; }
