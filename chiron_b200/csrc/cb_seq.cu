// Launchers of the CTC beam search and overlap assembly kernels (cb_seq_kernels.cuh; algorithms in cb_seq_algos.cuh), and the
// host-compiled self-test hooks of the sequential routines.
#include "cb_internal.cuh"
#include "cb_seq_kernels.cuh"
#include "../../include/chiron_b200_selftest.h"

#include <stdlib.h>
#include <string.h>

using namespace cb_seq;


namespace {

int ensure_buf(void** buf, size_t* cur, size_t need, const char* what) {
    if (need <= *cur) return CB_OK;
    if (*buf) { cudaFree(*buf); *buf = nullptr; *cur = 0; }
    cudaError_t e = cudaMalloc(buf, need);
    if (e != cudaSuccess) { cb_set_error("%s cudaMalloc(%zu bytes): %s", what, need, cudaGetErrorString(e)); return CB_ERR_NOMEM; }
    *cur = need;
    return CB_OK;
}

}  // namespace

int cb_launch_beam(cb_handle* h, const float* logits, const int32_t* lens, int B, int T, int W, int8_t* bases,
                   int32_t* n_bases, cudaStream_t s) {
    if (B <= 0) return CB_OK;
    if (W > 4096) { cb_set_error("beam width %d too large", W); return CB_ERR_ARG; }
    const int C = h->cfg.n_class;
    // Pool: children are only materialised when they enter the beam and the pool is compacted when full, so a few
    // thousand nodes cover T*W-sized tries.  2*W*(T+1)+2 nodes can never overflow (every live node is an ancestor of
    // a leaf or of a current branch); that size is used for a retry if the small pool ever proves too small.
    const long long cap = 2LL * W * (T + 1) + 2;
    {   // fast path: warp per window over shared memory with a small, frequently compacted pool
        // CB_BEAM_RETRY=1 (experimental, not yet run on a GPU): a small first-pass pool at high occupancy + a one-window-per-CTA
        // second pass for the windows that overflow it, instead of a first-pass pool sized for the tail.
        const bool retry = getenv("CB_BEAM_RETRY") && atoi(getenv("CB_BEAM_RETRY")) != 0;
        long long pool_first = beam_small_pool(T, W);
        if (retry && 8LL * W >= 64 && 8LL * W < pool_first) pool_first = 8LL * W;
        const long long pool_s = pool_first;
        const int smem_env = getenv("CB_BEAM_SMEM") ? atoi(getenv("CB_BEAM_SMEM")) : 1;    // 0: force the fallback kernel (tests)
        // Two variants of the same search (bit-identical outputs).  Staging a window's logits in shared memory saves the
        // per-frame global row read (3-5 % when every window is resident anyway) but enlarges the footprint; reading the
        // rows from global memory (prefetched a frame ahead) keeps more CTAs resident per SM and wins as soon as the windows
        // do not fit in one wave.  So: stage iff the staged launch fits the budget and is a single wave.
        // CB_BEAM_STAGE_LOGITS=0/1 forces a variant (tests, A/B).
        const size_t stride_staged = beam_warp_stride(T, C, W, (int)pool_s, true);
        const bool staged_fits = (long long)(stride_staged * BEAM_WARPS) <= BEAM_SMEM_BUDGET + 4096;
        const long long n_ctas = (B + BEAM_WARPS - 1) / BEAM_WARPS;
        const long long staged_ctas_per_sm = (227LL * 1024) / (long long)(stride_staged * BEAM_WARPS + 1024);
        const bool one_wave = n_ctas <= (long long)(h->sm_count > 0 ? h->sm_count : 148) * staged_ctas_per_sm;
        const bool want_staged = getenv("CB_BEAM_STAGE_LOGITS") ? atoi(getenv("CB_BEAM_STAGE_LOGITS")) != 0 : one_wave;
        const bool staged = want_staged && staged_fits;
        const size_t stride = beam_warp_stride(T, C, W, (int)pool_s, staged);
        if (smem_env && pool_s >= 2LL * W + 2 && stride * BEAM_WARPS <= 200 * 1024) {
            // per launch, like the recurrence launchers: the attribute belongs to the current device's copy of the kernel, so
            // a process-wide "already set" flag would miss every GPU but the first in a one-process multi-GPU host
            const dim3 grid((B + BEAM_WARPS - 1) / BEAM_WARPS);
            const size_t smem = stride * BEAM_WARPS;
            const int attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
            CB_CUDA(cudaMemsetAsync(h->d_flag, 0, sizeof(int), s));
            // A refused shared-memory opt-in or launch configuration is not an error of the decode: the thread-per-window
            // kernel below needs no shared memory and gives the same result.
            cudaError_t fe;
            if (retry) {            // first pass marks the windows that overflow its pool (n_bases = -1)
                if (staged) {
                    fe = cudaFuncSetAttribute(beam_warp_kernel<true, true>, (cudaFuncAttribute)attr, 200 * 1024);
                    if (fe == cudaSuccess)
                        beam_warp_kernel<true, true><<<grid, BEAM_WARPS * 32, smem, s>>>(
                            logits, lens, B, T, C, W, (int)pool_s, (int)stride, bases, n_bases, h->d_flag);
                } else {
                    fe = cudaFuncSetAttribute(beam_warp_kernel<false, true>, (cudaFuncAttribute)attr, 200 * 1024);
                    if (fe == cudaSuccess)
                        beam_warp_kernel<false, true><<<grid, BEAM_WARPS * 32, smem, s>>>(
                            logits, lens, B, T, C, W, (int)pool_s, (int)stride, bases, n_bases, h->d_flag);
                }
            } else if (staged) {
                fe = cudaFuncSetAttribute(beam_warp_kernel<true>, (cudaFuncAttribute)attr, 200 * 1024);
                if (fe == cudaSuccess)
                    beam_warp_kernel<true><<<grid, BEAM_WARPS * 32, smem, s>>>(
                        logits, lens, B, T, C, W, (int)pool_s, (int)stride, bases, n_bases, h->d_flag);
            } else {
                fe = cudaFuncSetAttribute(beam_warp_kernel<false>, (cudaFuncAttribute)attr, 200 * 1024);
                if (fe == cudaSuccess)
                    beam_warp_kernel<false><<<grid, BEAM_WARPS * 32, smem, s>>>(
                        logits, lens, B, T, C, W, (int)pool_s, (int)stride, bases, n_bases, h->d_flag);
            }
            if (fe == cudaSuccess) fe = cudaGetLastError();
            int flag = 1;              // 1 = take the fallback below
            if (fe == cudaSuccess) {
                h->launches++;
                // the beam decoder is synchronous (like the reference's decode dequeue)
                CB_CUDA(cudaMemcpyAsync(&flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
                CB_CUDA(cudaStreamSynchronize(s));
                if (!flag) return CB_OK;
            } else {
                static bool told = false;          // say so once: it is a tenfold slowdown of the decode, not an error
                if (!told) {
                    fprintf(stderr, "chiron_b200: shared-memory beam search refused (%s, %zu bytes per CTA); using the "
                                    "thread-per-window kernel\n", cudaGetErrorString(fe), smem);
                    told = true;
                }
                (void)cudaGetLastError();          // launch-configuration errors are not sticky: clear and fall back
            }
            if (retry && fe == cudaSuccess) {   // second pass: the marked windows alone, one per CTA with a pool of ~130 W nodes
                const long long pool_r = beam_retry_pool(T, W);
                const size_t smem_r = align_up(cb_beam_work_bytes(W, (int)pool_r), 16);
                CB_CUDA(cudaFuncSetAttribute(beam_retry_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
                CB_CUDA(cudaMemsetAsync(h->d_flag, 0, sizeof(int), s));
                beam_retry_kernel<<<B, 32, smem_r, s>>>(logits, lens, B, T, C, W, (int)pool_r, bases, n_bases, h->d_flag);
                CB_CHECK_LAUNCH();
                h->launches++;
                CB_CUDA(cudaMemcpyAsync(&flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
                CB_CUDA(cudaStreamSynchronize(s));
                if (!flag) return CB_OK;
            }
        }
    }
    long long pool = 4LL * W + 4096;
    for (int attempt = 0; attempt < 2; ++attempt) {
        if (pool > cap || attempt == 1) pool = cap;
        if (pool < 2LL * W + 2) pool = 2LL * W + 2;
        const size_t stride = align_up(cb_beam_work_bytes(W, (int)pool), 16);
        int rc = ensure_buf(&h->beam_ws, &h->beam_ws_bytes, stride * (size_t)B, "beam workspace");
        if (rc != CB_OK) return rc;
        CB_CUDA(cudaMemsetAsync(h->d_flag, 0, sizeof(int), s));
        beam_kernel<<<(B + 63) / 64, 64, 0, s>>>(logits, lens, B, T, C, W, (int)pool, (char*)h->beam_ws, stride, bases,
                                                 n_bases, h->d_flag);
        CB_CHECK_LAUNCH();
        h->launches++;
        int flag = 0;                  // the beam decoder is synchronous (like the reference's decode dequeue)
        CB_CUDA(cudaMemcpyAsync(&flag, h->d_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
        CB_CUDA(cudaStreamSynchronize(s));
        if (!flag) return CB_OK;
        if (pool == cap) break;
    }
    cb_set_error("beam search node pool exhausted (beam_width %d, T %d)", W, T);
    return CB_ERR_NOMEM;
}

int cb_launch_assemble(cb_handle* h, const int8_t* bases, const int32_t* n_bases, const float* path_prob, int n_windows,
                       int T, int jump, int L, int kernel, int8_t* consensus, char* qual, int32_t* pos,
                       int32_t* out_len, int max_len, cudaStream_t s) {
    if (n_windows == 0) { CB_CUDA(cudaMemsetAsync(out_len, 0, sizeof(int32_t), s)); return CB_OK; }
    const AsmPlan plan = asm_plan(n_windows, T, kernel, max_len);
    const size_t ml = plan.ml;
    int rc = ensure_buf(&h->asm_ws, &h->asm_ws_bytes, plan.total, "assembly workspace");
    if (rc != CB_OK) return rc;
    char* base = (char*)h->asm_ws;
    AsmWork w = asm_work(base, plan);
    CB_CUDA(cudaMemsetAsync(base + plan.counts, 0, plan.logfact - plan.counts, s));       // counts + qsum
    asm_compact_kernel<<<1, 1024, 0, s>>>(n_bases, n_windows, T, w, pos);
    CB_CHECK_LAUNCH();
    const double jsr = (double)jump / (double)L;                                // FLAGS.jump / FLAGS.segment_len
    asm_disp_kernel<<<(n_windows + 127) / 128, 128, 0, s>>>(bases, n_bases, T, kernel, jsr, w);
    CB_CHECK_LAUNCH();
    asm_scan_kernel<<<1, 1024, 0, s>>>(n_bases, w, pos, out_len, max_len);
    CB_CHECK_LAUNCH();
    asm_vote_kernel<<<n_windows, 128, 0, s>>>(bases, n_bases, path_prob, T, w, (int)ml);
    CB_CHECK_LAUNCH();
    if (max_len > 0) {
        asm_finish_kernel<<<(max_len + 255) / 256, 256, 0, s>>>(w, (int)ml, consensus, qual);
        CB_CHECK_LAUNCH();
    }
    h->launches += 5;
    return CB_OK;
}

// ---- host-compiled instantiations of the same routines (unit tests without a GPU; never on the product path) -------
extern "C" int cb_selftest_beam(const float* logits, int len, int n_class, int beam_width, int pool, int8_t* out) {
    if (!logits || !out || n_class < 2 || n_class > 8 || beam_width < 1 || pool < 2 * beam_width + 2) return -1;
    void* mem = malloc(cb_beam_work_bytes(beam_width, pool));
    if (!mem) return -1;
    CbBeamWork k = cb_beam_work_carve(mem, beam_width, pool);
    const int n = cb_beam_decode_one(logits, len, n_class, beam_width, k, out);
    free(mem);
    return n;
}

extern "C" int cb_selftest_disp(const int8_t* cur, int la, const int8_t* prev, int lb, int kernel, int jump, int L) {
    if (kernel == CB_ASM_STICK) return cb_disp_stick(la, lb);
    if (kernel == CB_ASM_GLUE) return cb_disp_glue(cur, la, prev, lb);
    const int mx = la > lb ? la : lb;
    double* lf = (double*)malloc(sizeof(double) * (mx + 2));
    int* scratch = (int*)malloc(sizeof(int) * cb_simple_scratch_ints(la, lb));
    double acc = 0.0;
    lf[0] = 0.0;
    for (int k = 1; k <= mx + 1; ++k) { acc += log((double)k); lf[k] = acc; }
    const int d = cb_disp_simple(cur, la, prev, lb, (double)jump / (double)L, lf, scratch);
    free(lf); free(scratch);
    return d;
}
