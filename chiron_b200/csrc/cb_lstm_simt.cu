// Launcher of the fp32 LSTM recurrence (cb_lstm_simt_kernel.cuh).
#include "cb_internal.cuh"
#include "cb_lstm_simt_kernel.cuh"

using namespace cb_lstm;

namespace {

template <int RG>
int launch_rg(cb_handle* h, const LstmProblem& p, cudaStream_t s) {
    constexpr int R = RPT * RG;
    const size_t smem = lstm_smem_bytes(p.H, RG);
    if (smem > 227 * 1024) { cb_set_error("lstm: hidden size %d does not fit shared memory", p.H); return CB_ERR_ARG; }
    CB_CUDA(cudaFuncSetAttribute(lstm_simt_kernel<RG>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    lstm_simt_kernel<RG><<<dim3((p.B + R - 1) / R, 2), 128 * RG, smem, s>>>(p);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}

}  // namespace

int cb_launch_lstm_simt(cb_handle* h, const LstmProblem& p, cudaStream_t s) {
    if (p.B <= 0 || p.T <= 0) return CB_OK;
    if (p.H > 128 || (p.H & 3)) { cb_set_error("lstm: hidden size must be a multiple of 4 and <= 128"); return CB_ERR_ARG; }
    // rows per CTA: the largest slice that still gives (nearly) every SM a CTA; grid.y = the two directions
    const int sms = h->sm_count > 0 ? h->sm_count : 148;
    if (2 * ((p.B + 63) / 64) >= (sms * 3) / 4) return launch_rg<4>(h, p, s);
    if (2 * ((p.B + 31) / 32) >= (sms * 3) / 4) return launch_rg<2>(h, p, s);
    return launch_rg<1>(h, p, s);
}
