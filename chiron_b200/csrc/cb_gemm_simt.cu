// Launcher of the fp32 FFMA implicit-GEMM kernel (cb_gemm_simt_kernel.cuh).
#include "cb_internal.cuh"
#include "cb_gemm_simt_kernel.cuh"

using namespace cb_simt;


int cb_launch_gemm_simt(cb_handle* h, const GemmProblem& p, cudaStream_t s) {
    if (p.M <= 0) return CB_OK;
    if ((p.c0 & 3) || (p.c1 & 3) || (p.N & 3) || (p.lda0 & 3) || (p.lda1 & 3) || (p.ldo & 3)) {
        cb_set_error("gemm: channel counts / leading dimensions must be multiples of 4");
        return CB_ERR_ARG;
    }
    const long long blocks = (long long)((p.N + BN - 1) / BN) * ((p.M + BM - 1) / BM);
    if (blocks > 0x7fffffffLL) { cb_set_error("gemm: problem too large"); return CB_ERR_ARG; }
    gemm_simt_kernel<<<(unsigned)blocks, NT, 0, s>>>(p);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}
