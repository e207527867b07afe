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

__global__ void fill_i32_kernel(int32_t* p, int n, int32_t v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

}  // namespace

// tf.nn.ctc_beam_search_decoder for a batch, fully asynchronous on `s`: three passes enqueued back to back, each a no-op
// unless the one before it marked windows (n_bases = -1) whose trie outgrew its node pool:
//   1. beam_warp_kernel   warp per window, 16W+ nodes in shared memory, 4-5 CTAs per SM     (every window)
//   2. beam_retry_kernel  one window per CTA, ~260W nodes (the CTA's whole shared memory)    (marked windows: <1 % on real logits)
//   3. beam_kernel        thread per window over global workspaces sized so the pool cannot overflow (2W(T+1)+2 nodes),
//                         claimed atomically from a small set                                (practically never)
// If pass 3 runs out of workspaces the sticky CB_FLAG_BEAM_ERROR is raised; cb_check_status / cb_basecall_collect /
// cb_basecall_host report it as CB_ERR_NOMEM.  No host synchronisation here.
// CB_BEAM_SMEM=0 skips passes 1-2 and runs beam_kernel on every window (tests, A/B).
int cb_launch_beam(cb_handle* h, const float* logits, const int32_t* lens, int B, int T, int W, int8_t* bases,
                   int32_t* n_bases, float* scores, cudaStream_t s) {
    if (B <= 0) return CB_OK;
    if (W < 1 || W > 4096) { cb_set_error("beam width %d out of range", W); return CB_ERR_ARG; }
    const int C = h->cfg.n_class;
    const long long cap = 2LL * W * (T + 1) + 2;        // every live node is an ancestor of a leaf or of a current branch
    const int smem_env = getenv("CB_BEAM_SMEM") ? atoi(getenv("CB_BEAM_SMEM")) : 1;
    const size_t stride_g = align_up(cb_beam_work_bytes(W, (int)cap), 16);
    if (!smem_env || T > 32767) {                        // the reference search on every window
        int rc = ensure_buf(&h->beam_ws, &h->beam_ws_bytes, stride_g * (size_t)B, "beam workspace");
        if (rc != CB_OK) return rc;
        beam_kernel<<<(B + 63) / 64, 64, 0, s>>>(logits, lens, B, T, C, W, (int)cap, (char*)h->beam_ws, stride_g, bases, n_bases,
                                                 h->d_flag + CB_FLAG_BEAM_ERROR, nullptr, 0, scores);
        CB_CHECK_LAUNCH();
        h->launches++;
        return CB_OK;
    }
    // workspaces of pass 3: up to 32 windows, at most 256 MB, at least one (allocated before anything is enqueued: a
    // cudaMalloc in the middle of the decode would synchronise the device)
    long long n_slots = (256LL << 20) / (long long)stride_g;
    n_slots = n_slots > 32 ? 32 : (n_slots < 1 ? 1 : n_slots);
    if (n_slots > B) n_slots = B;
    int rc = ensure_buf(&h->beam_ws, &h->beam_ws_bytes, stride_g * (size_t)n_slots, "beam workspace");
    if (rc != CB_OK) return rc;
    CB_CUDA(cudaMemsetAsync(h->d_flag + CB_FLAG_BEAM_MARKED, 0, 2 * sizeof(int), s));     // marked flag + slot counter

    // ---- pass 1.  Two variants of the same search (bit-identical outputs): staging a window's logits in shared memory saves
    // the per-frame global row read (3-5 % when every window is resident anyway) but enlarges the footprint; reading the rows
    // from global memory (prefetched a frame ahead) keeps more CTAs resident per SM and wins as soon as the windows do not
    // fit in one wave.  So: stage iff the staged launch fits the budget and is a single wave (CB_BEAM_STAGE_LOGITS=0/1 forces).
    // 16W nodes at 4-5 CTAs per SM, or 24W nodes (no window of the surveyed real logits outgrows that) at 2: in units of one
    // window's search time a launch costs its number of waves, plus one more for the retry pass the smaller pool needs for its
    // tail -- so a batch that fits one wave at 24W takes 24W (measured, 1024 x 512, W=30: 4.8 ms against 9.4).
    long long pool_s = beam_small_pool(T, W);
    {
        const long long pool_hi = beam_small_pool(T, W, 24);
        auto waves = [&](long long pool) {
            const long long cta = (long long)beam_warp_stride(T, C, W, (int)pool, false) * BEAM_WARPS + 1024;
            const long long per_sm = (227LL * 1024) / cta;
            const long long n = (B + BEAM_WARPS - 1) / BEAM_WARPS, slots = (long long)(h->sm_count > 0 ? h->sm_count : 148) * (per_sm > 0 ? per_sm : 1);
            return (n + slots - 1) / slots;
        };
        if (pool_hi > pool_s && waves(pool_hi) < waves(pool_s) + 1) pool_s = pool_hi;
    }
    if (getenv("CB_BEAM_POOL") && atoll(getenv("CB_BEAM_POOL")) >= 2LL * W + 2) pool_s = atoll(getenv("CB_BEAM_POOL"));   // A/B
    if (pool_s > 32767) pool_s = 32767;
    const size_t stride_staged = beam_warp_stride(T, C, W, (int)pool_s, true);
    const bool staged_fits = (long long)(stride_staged * BEAM_WARPS) <= BEAM_SMEM_BUDGET + 4096;
    const long long n_ctas = (B + BEAM_WARPS - 1) / BEAM_WARPS;
    const long long staged_ctas_per_sm = (227LL * 1024) / (long long)(stride_staged * BEAM_WARPS + 1024);
    const bool one_wave = n_ctas <= (long long)(h->sm_count > 0 ? h->sm_count : 148) * staged_ctas_per_sm;
    const bool want_staged = getenv("CB_BEAM_STAGE_LOGITS") ? atoi(getenv("CB_BEAM_STAGE_LOGITS")) != 0 : one_wave;
    const bool staged = want_staged && staged_fits;
    const size_t stride = beam_warp_stride(T, C, W, (int)pool_s, staged);
    const int attr = cudaFuncAttributeMaxDynamicSharedMemorySize;
    bool pass1 = false;
    if (pool_s >= 2LL * W + 2 && stride * BEAM_WARPS <= 200 * 1024) {
        // the attribute is set per launch: it belongs to the current device's copy of the kernel.  A refused opt-in or
        // launch configuration is not an error of the decode: the later passes give the same result.
        const dim3 grid((unsigned)n_ctas);
        const size_t smem = stride * BEAM_WARPS;
        cudaError_t fe;
        if (staged) {
            fe = cudaFuncSetAttribute(beam_warp_kernel<true>, (cudaFuncAttribute)attr, 200 * 1024);
            if (fe == cudaSuccess)
                beam_warp_kernel<true><<<grid, BEAM_WARPS * 32, smem, s>>>(logits, lens, B, T, C, W, (int)pool_s, (int)stride, bases,
                                                                          n_bases, h->d_flag + CB_FLAG_BEAM_MARKED, scores);
        } else {
            fe = cudaFuncSetAttribute(beam_warp_kernel<false>, (cudaFuncAttribute)attr, 200 * 1024);
            if (fe == cudaSuccess)
                beam_warp_kernel<false><<<grid, BEAM_WARPS * 32, smem, s>>>(logits, lens, B, T, C, W, (int)pool_s, (int)stride, bases,
                                                                           n_bases, h->d_flag + CB_FLAG_BEAM_MARKED, scores);
        }
        if (fe == cudaSuccess) fe = cudaGetLastError();
        if (fe == cudaSuccess) { pass1 = true; h->launches++; }
        else (void)cudaGetLastError();                  // launch-configuration errors are not sticky: clear and go on
    }
    if (!pass1) {                                        // nothing ran: mark every window for the passes below
        fill_i32_kernel<<<(B + 255) / 256, 256, 0, s>>>(n_bases, B, -1);
        CB_CHECK_LAUNCH();
        h->launches++;
    }
    // ---- pass 2: marked windows, one per CTA
    const long long pool_r = beam_retry_pool(T, W);
    if (pool_r >= 2LL * W + 2 && pool_r > pool_s) {
        const size_t smem_r = align_up(cb_beam_work_bytes<BeamIdx>(W, (int)pool_r), 16);
        const cudaError_t fe = cudaFuncSetAttribute(beam_retry_kernel, (cudaFuncAttribute)attr, 200 * 1024);
        if (fe == cudaSuccess) {
            beam_retry_kernel<<<B, 32, smem_r, s>>>(logits, lens, B, T, C, W, (int)pool_r, bases, n_bases,
                                                    h->d_flag + CB_FLAG_BEAM_MARKED, scores);
            if (cudaGetLastError() == cudaSuccess) h->launches++;
        } else {
            (void)cudaGetLastError();
        }
    }
    // ---- pass 3: windows still marked, over global workspaces that cannot overflow
    beam_kernel<<<(B + 63) / 64, 64, 0, s>>>(logits, lens, B, T, C, W, (int)cap, (char*)h->beam_ws, stride_g, bases, n_bases,
                                             h->d_flag + CB_FLAG_BEAM_ERROR, h->d_flag + CB_FLAG_BEAM_SLOTS, (int)n_slots, scores);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}

// Deferred device-side errors of the asynchronous calls: synchronises `s`, reads the sticky flags of the status block.
int cb_check_deferred(cb_handle* h, cudaStream_t s) {
    int flags[2] = {0, 0};
    CB_CUDA(cudaMemcpyAsync(flags, h->d_flag + CB_FLAG_BEAM_ERROR, sizeof(flags), cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    if (flags[0] || flags[1]) CB_CUDA(cudaMemsetAsync(h->d_flag + CB_FLAG_BEAM_ERROR, 0, sizeof(flags), s));
    if (flags[1]) {
        cb_set_error("an activation exceeded the fp16 range of the tensor-core path; rerun with precision fp32");
        return CB_ERR_RANGE;
    }
    if (flags[0]) {
        cb_set_error("beam search: more windows than fallback workspaces outgrew the shared-memory node pools");
        return CB_ERR_NOMEM;
    }
    return CB_OK;
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

// the same search over the 16-bit trie of the shared-memory kernels (pool, len and beam_width below 32768)
extern "C" int cb_selftest_beam16(const float* logits, int len, int n_class, int beam_width, int pool, int8_t* out) {
    if (!logits || !out || n_class < 2 || n_class > 8 || beam_width < 1 || pool < 2 * beam_width + 2 || pool > 32767 ||
        len > 32767)
        return -1;
    void* mem = malloc(cb_beam_work_bytes<BeamIdx>(beam_width, pool));
    if (!mem) return -1;
    BeamWorkS k = cb_beam_work_carve<BeamIdx>(mem, beam_width, pool);
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
