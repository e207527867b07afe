// tcgen05 persistent LSTM recurrence (CB_PREC_TC_SPLIT / CB_PREC_TC_FAST).
//
// Replaces the tf.while_loop of dynamic_rnn around LSTMCell (chiron/rnn.py:49-50,64,140-143), like cb_lstm_simt.cu, but
// the per-step contraction h[128 rows,100] x W_hh[100,400] runs on the tensor core:
//   * one CTA owns 128 batch rows of one direction for all T steps (rows are independent: no grid synchronisation);
//   * W_hh is resident in shared memory for the whole kernel as fp16 hi/lo K-major core-matrix images (166 KB), the
//     hidden state h is re-written every step by the gate warps as fp16 hi/lo (56 KB);
//   * per step one thread issues 42 tcgen05.mma (7 K-steps x {h_hi*W_lo, h_lo*W_hi first, then h_hi*W_hi} x 2 N-halves
//     of 208/192 columns) into a 400-column fp32 TMEM accumulator -- low-order products are accumulated first so the
//     tensor core's truncating accumulator adds the big terms last;
//   * 15 gate warps (3-4 per TMEM lane quadrant, each owning 5-10 half-groups of 4 hidden units of its 32 rows) read the
//     accumulator with tcgen05.ld, add the hoisted input projection (software-prefetched one half-group ahead),
//     evaluate the cell with MUFU ex2/rcp (7 per cell), keep c in registers, and write h back as one 16-byte
//     core-matrix row per (row, 8-unit K-group).
// Global layouts are time-major with the batch innermost -- pre[T][8H][Bp], out[T][2H][Bp] -- so that the 32 rows of a
// warp read and write consecutive floats (128-byte coalesced) at every step.
#include <cuda_fp16.h>
#include <string.h>

#include <vector>

#include "cb_internal.cuh"
#include "cb_tc_common.cuh"

namespace {

constexpr int RM = 128;            // rows per CTA = UMMA M
constexpr int H = 100, H4 = 400;   // this kernel is specialised for the shipped hidden size
constexpr int KG = 13;             // 16-byte K-groups that hold real data (13*8 = 104 >= 100)
constexpr int KG_A = 14;           // K-groups of the h operand (7 K-steps of 16)
constexpr int N0 = 208, N1 = 192;  // N split of the 400 gate columns (multiples of 16)
constexpr int C_COL = 400;         // TMEM columns 400..499 hold the cell state c[row][unit]
constexpr int GATE_WARPS = 15;   // 16 warps in total: 128 registers per thread (the register file is per SM quarter)
constexpr int NTHREADS = (1 + GATE_WARPS) * 32;
constexpr uint32_t W_BYTES = KG * H4 * 16;        // one of hi / lo: 83,200
constexpr uint32_t HS_BYTES = KG_A * RM * 16;     // one of hi / lo: 28,672
constexpr size_t SMEM_BYTES = 2 * (size_t)W_BYTES + 2 * (size_t)HS_BYTES + 64;

struct LstmTcParams {
    int B, Bp, T;
    const float* pre;          // [T][8H][Bp]
    const __half* wimg[2];     // per direction: hi image then lo image, [KG][400][8] halfs each
    const int32_t* lens;       // [B]
    float* out;                // [T][2H][Bp] fp32 (written when write_f32: the last layer, read by the logit head)
    CbImg o_img;               // hi/lo operand image of h for the next layer's input projection (when write_img):
                               //   plane dir*13 + kg, row row0 + t*Bp + b
    int write_f32, write_img;
    int passes;
};

// h(t) of one (row, 8-unit K-group) as one 16-byte core-matrix row per hi / lo image: into shared memory for the next
// step's MMA and (optionally) into the global operand image the next layer's input projection bulk-loads.
__device__ __forceinline__ void store_h_row(uint8_t* h_hi, uint8_t* h_lo, int kg, int row, __half* g_hi, __half* g_lo,
                                            float v0, float v1, float v2, float v3, float v4, float v5, float v6, float v7) {
    const float v[8] = {v0, v1, v2, v3, v4, v5, v6, v7};
    uint4 hi, lo;
    split8(v, hi, lo);
    *reinterpret_cast<uint4*>(h_hi + (size_t)kg * (RM * 16) + row * 16) = hi;
    *reinterpret_cast<uint4*>(h_lo + (size_t)kg * (RM * 16) + row * 16) = lo;
    if (g_hi) {
        *reinterpret_cast<uint4*>(g_hi) = hi;
        *reinterpret_cast<uint4*>(g_lo) = lo;
    }
}

__device__ __forceinline__ float ex2f(float x) { float y; asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }
__device__ __forceinline__ float rcpf(float x) { float y; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x)); return y; }

// One LSTM cell (TF LSTMCell, forget_bias 1.0):  c' = sigmoid(f+1)*c + sigmoid(i)*tanh(j);  h' = sigmoid(o)*tanh(c').
// sigmoid(x) = 1/(1+e^-x), tanh(x) = (1-e^-2x)/(1+e^-2x); the quotients are merged so a cell costs 5 ex2 + 2 rcp.
__device__ __forceinline__ void lstm_cell(float gi, float gj, float gf, float go, float& c, float& h) {
    constexpr float L2E = 1.4426950408889634f;
    // e^-x only overflows for very negative x, so a one-sided clamp keeps every denominator finite (25: e^25 ~ 7e10,
    // three of them multiply to < FLT_MAX); for large positive x, e^-x -> 0 and the quotients saturate by themselves.
    const float ei = ex2f(-L2E * fmaxf(gi, -25.f));
    const float ej = ex2f(-2.f * L2E * fmaxf(gj, -12.5f));
    const float ef = ex2f(-L2E * fmaxf(gf + 1.0f, -25.f));
    const float di = 1.f + ei, dj = 1.f + ej, df = 1.f + ef;
    const float dij = di * dj;
    // c' = c/df + (1-ej)/(di*dj) = (c*dij + (1-ej)*df) / (df*dij)
    const float cn = fmaf(c, dij, (1.f - ej) * df) * rcpf(df * dij);
    const float eo = ex2f(-L2E * fmaxf(go, -25.f));
    const float ec = ex2f(-2.f * L2E * fmaxf(cn, -12.5f));
    h = (1.f - ec) * rcpf((1.f + eo) * (1.f + ec));
    c = cn;
}

// The gate loop of one warp: `nh` half-groups (4 hidden units each) starting at half-group `hg0`, for the warp's 32
// rows.  The loop is deliberately NOT unrolled (a fully unrolled version is ~300 KB of SASS and thrashes the instruction
// cache); the cell state c therefore lives in the 100 spare TMEM columns next to the accumulator instead of registers.
__device__ __forceinline__ void gate_loop(const LstmTcParams& q, uint8_t* h_hi, uint8_t* h_lo, uint64_t* h_ready,
                                          uint64_t* acc_ready, uint32_t tmem_base, int warp, int lane, int dir, int b0,
                                          int hg0, int nh) {
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int b = b0 + row;
    int len = 0;
    if (b < q.B) { len = q.lens[b]; len = len < 0 ? 0 : (len > q.T ? q.T : len); }
    const size_t Bp = (size_t)q.Bp;
    const float* pre_b = q.pre + (size_t)dir * H4 * Bp + b;           // + (t*8H + col)*Bp
    float* out_b = q.out + (size_t)dir * H * Bp + b;                  // + (t*2H + u)*Bp
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t t_cell = t_lane + C_COL;                           // c[row][u] at column C_COL + u

    for (int i = 0; i < nh; ++i) tmem_st4(t_cell + (hg0 + i) * 4, 0.f, 0.f, 0.f, 0.f);
    tmem_st_wait();

    // frame this row works on at step s (inactive rows: frame s, where they write zeros)
    auto frame_of = [&](int s) { return s < len ? (dir ? len - 1 - s : s) : s; };
    // the 16 hoisted input-projection values (4 gates x 4 units) of one half-group; coalesced over the warp's rows
    auto load_pre = [&](float (&dst)[16], int t, int hg) {
        const float* src = pre_b + ((size_t)t * (8 * H) + hg * 4) * Bp;
#pragma unroll
        for (int g = 0; g < 4; ++g)
#pragma unroll
            for (int e = 0; e < 4; ++e) dst[g * 4 + e] = __ldg(src + (size_t)(g * H + e) * Bp);
    };

    float pre_cur[16], pre_nxt[16];
    load_pre(pre_cur, frame_of(0), hg0);
    for (int s = 0; s < q.T; ++s) {
        const bool active = s < len;
        const int t = frame_of(s);
        const int t_next = s + 1 < q.T ? frame_of(s + 1) : t;
        float* out_t = out_b + (size_t)t * (2 * H) * Bp;
        const size_t img_row = (size_t)q.o_img.row0 + (size_t)t * Bp + b;
        mbar_wait(acc_ready, s & 1);
        tc_fence_after();
        float hlow[4] = {0.f, 0.f, 0.f, 0.f};       // h of the even half-group, carried to the odd one (same K-group)
#pragma unroll 1
        for (int i = 0; i < nh; ++i) {
            const int hg = hg0 + i, u0 = hg * 4;
            // prefetch the next half-group (or the first one of the next step) while this one is evaluated
            if (i + 1 < nh) load_pre(pre_nxt, t, hg + 1);
            else load_pre(pre_nxt, t_next, hg0);
            float z[16], c[4];
            tmem_ld4(t_cell + u0, c);
            if (s > 0) {
#pragma unroll
                for (int g = 0; g < 4; ++g) tmem_ld4(t_lane + g * H + u0, *reinterpret_cast<float(*)[4]>(&z[g * 4]));
            } else {
#pragma unroll
                for (int e = 0; e < 16; ++e) z[e] = 0.f;
            }
            tmem_ld_wait();
            float hv[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                hv[e] = 0.f;
                if (active)
                    lstm_cell(z[e] + pre_cur[e], z[4 + e] + pre_cur[4 + e], z[8 + e] + pre_cur[8 + e],
                              z[12 + e] + pre_cur[12 + e], c[e], hv[e]);
                if (q.write_f32) out_t[(size_t)(u0 + e) * Bp] = hv[e];
            }
            tmem_st4(t_cell + u0, c[0], c[1], c[2], c[3]);
            __half *g_hi = nullptr, *g_lo = nullptr;
            if (q.write_img) {
                const size_t off = ((size_t)(dir * KG + (hg >> 1)) * q.o_img.plane_rows + img_row) * 8;
                g_hi = q.o_img.hi + off; g_lo = q.o_img.lo + off;
            }
            if (hg & 1) {
                store_h_row(h_hi, h_lo, hg >> 1, row, g_hi, g_lo, hlow[0], hlow[1], hlow[2], hlow[3], hv[0], hv[1], hv[2], hv[3]);
            } else if (hg == 24) {                   // units 100..103 do not exist: upper half of the last K-group is zero
                store_h_row(h_hi, h_lo, 12, row, g_hi, g_lo, hv[0], hv[1], hv[2], hv[3], 0.f, 0.f, 0.f, 0.f);
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e) hlow[e] = hv[e];
            }
#pragma unroll
            for (int e = 0; e < 16; ++e) pre_cur[e] = pre_nxt[e];
        }
        tmem_st_wait();
        fence_proxy_async();
        tc_fence_before();
        mbar_arrive(h_ready);
    }
}

__global__ void __launch_bounds__(NTHREADS, 1) lstm_tc_kernel(const LstmTcParams q) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_hi = smem;
    uint8_t* w_lo = smem + W_BYTES;
    uint8_t* h_hi = smem + 2 * (size_t)W_BYTES;
    uint8_t* h_lo = h_hi + HS_BYTES;
    uint64_t* bars = reinterpret_cast<uint64_t*>(h_lo + HS_BYTES);
    uint64_t* w_bar = bars;            // weights landed
    uint64_t* h_ready = bars + 1;      // gate warps wrote h(t) and released the accumulator
    uint64_t* acc_ready = bars + 2;    // MMAs of the step retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 3);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const int b0 = blockIdx.x * RM;

    // zero the h operand (h(0) = 0; padding K-groups stay zero for the whole kernel)
    for (uint32_t i = threadIdx.x * 16; i < 2 * HS_BYTES; i += NTHREADS * 16)
        *reinterpret_cast<uint4*>(h_hi + i) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(w_bar, 1);
        mbar_init(h_ready, GATE_WARPS * 32);
        mbar_init(acc_ready, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();               // the zero fill must be visible to the tensor core
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================ MMA issuer =====================================================================
        if (lane == 0) {
            mbar_arrive_expect_tx(w_bar, 2 * W_BYTES);
            bulk_g2s(w_hi, q.wimg[dir], 2 * W_BYTES, w_bar);
            mbar_wait(w_bar, 0);
            const uint32_t idesc0 = (1u << 4) | ((uint32_t)(N0 >> 3) << 17) | ((uint32_t)(RM >> 4) << 24);
            const uint32_t idesc1 = (1u << 4) | ((uint32_t)(N1 >> 3) << 17) | ((uint32_t)(RM >> 4) << 24);
            // descriptors differ only in the 14-bit start-address field (units of 16 B): advance it with integer adds
            const uint64_t da_hi = make_desc(smem_u32(h_hi), RM * 16, 128), da_lo = make_desc(smem_u32(h_lo), RM * 16, 128);
            const uint64_t db_hi = make_desc(smem_u32(w_hi), H4 * 16, 128), db_lo = make_desc(smem_u32(w_lo), H4 * 16, 128);
            constexpr uint32_t A_STEP = 2 * RM, B_STEP = 2 * H4;       // two K-groups per UMMA K-step, in 16 B units
            for (int s = 0; s < q.T; ++s) {
                if (s > 0) {                                   // h(0) = 0: the first step has no recurrent term
                    mbar_wait(h_ready, (s - 1) & 1);
                    tc_fence_after();
#pragma unroll
                    for (int half = 0; half < 2; ++half) {
                        const uint32_t d = tmem_base + (half ? N0 : 0);
                        const uint32_t idesc = half ? idesc1 : idesc0;
                        const uint32_t brow = half ? N0 : 0;   // 16 B units
                        uint32_t acc = 0;
                        if (q.passes == 3) {
#pragma unroll
                            for (int ks = 0; ks < KG_A / 2; ++ks) {   // low-order products first
                                umma_f16(d, da_hi + ks * A_STEP, db_lo + (ks * B_STEP + brow), idesc, acc);
                                acc = 1;
                                umma_f16(d, da_lo + ks * A_STEP, db_hi + (ks * B_STEP + brow), idesc, 1);
                            }
                        }
#pragma unroll
                        for (int ks = 0; ks < KG_A / 2; ++ks) {
                            umma_f16(d, da_hi + ks * A_STEP, db_hi + (ks * B_STEP + brow), idesc, acc);
                            acc = 1;
                        }
                    }
                }
                umma_commit(acc_ready);
            }
        }
    } else {
        // ============================ gate warps ======================================================================
        // Warp w may only touch TMEM lanes 32*(w%4)..+31.  Quadrant 0 shares its SM quarter with the MMA warp and has
        // three gate warps (half-groups [0,10) [10,18) [18,25)); quadrants 1-3 have four ([0,8) [8,14) [14,20) [20,25)).
        const int idx = warp >> 2;
        int hg0, nh;
        if ((warp & 3) == 0) { hg0 = idx == 1 ? 0 : (idx == 2 ? 10 : 18); nh = idx == 1 ? 10 : (idx == 2 ? 8 : 7); }
        else { hg0 = idx == 0 ? 0 : 2 + 6 * idx; nh = idx == 0 ? 8 : (idx == 3 ? 5 : 6); }
        gate_loop(q, h_hi, h_lo, h_ready, acc_ready, tmem_base, warp, lane, dir, b0, hg0, nh);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

struct LstmTcState {
    __half* wimg[CB_MAX_LAYERS][2];
};

}  // namespace

int cb_lstm_tc_prepare(cb_handle* h, const float* hw) {
    if (h->cfg.hidden != H) {
        cb_set_error("tensor-core path is specialised for hidden=%d (model has %d); use precision fp32", H, h->cfg.hidden);
        return CB_ERR_ARG;
    }
    LstmTcState* st = new LstmTcState();
    memset(st, 0, sizeof(*st));
    h->lstm_tc = st;
    for (int l = 0; l < h->cfg.n_layers; ++l)
        for (int d = 0; d < 2; ++d) {
            const float* W = hw + (h->whh[l][d] - h->d_weights);      // [H][4H]
            std::vector<__half> img(2 * (size_t)KG * H4 * 8);
            for (int g = 0; g < KG; ++g)
                for (int n = 0; n < H4; ++n)
                    for (int e = 0; e < 8; ++e) {
                        const int k = g * 8 + e;
                        const float w = k < H ? W[(size_t)k * H4 + n] : 0.f;
                        const __half hi = __float2half_rn(w);
                        const __half lo = __float2half_rn(w - __half2float(hi));
                        img[((size_t)g * H4 + n) * 8 + e] = hi;
                        img[(size_t)KG * H4 * 8 + ((size_t)g * H4 + n) * 8 + e] = lo;
                    }
            CB_CUDA(cudaMalloc(&st->wimg[l][d], img.size() * sizeof(__half)));
            CB_CUDA(cudaMemcpy(st->wimg[l][d], img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
        }
    CB_CUDA(cudaFuncSetAttribute(lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    return CB_OK;
}

void cb_lstm_tc_release(cb_handle* h) {
    LstmTcState* st = (LstmTcState*)h->lstm_tc;
    if (!st) return;
    for (int l = 0; l < CB_MAX_LAYERS; ++l)
        for (int d = 0; d < 2; ++d) if (st->wimg[l][d]) cudaFree(st->wimg[l][d]);
    delete st;
    h->lstm_tc = nullptr;
}

bool cb_lstm_tc_available(const cb_handle* h) { return h->lstm_tc != nullptr; }

// pre: [T][8H][Bp], out: [T][2H][Bp] (time-major, batch innermost), Bp a multiple of 128.
int cb_launch_lstm_tc(cb_handle* h, const LstmProblem& p, const CbImg* o_img, int write_f32, cudaStream_t s) {
    LstmTcState* st = (LstmTcState*)h->lstm_tc;
    if (!st) { cb_set_error("lstm tensor-core path unavailable for hidden=%d", h->cfg.hidden); return CB_ERR_ARG; }
    if (p.B <= 0 || p.T <= 0) return CB_OK;
    LstmTcParams q;
    memset(&q, 0, sizeof(q));
    q.B = p.B; q.Bp = p.ld_pre; q.T = p.T; q.pre = p.pre; q.lens = p.lens; q.out = p.out;
    q.wimg[0] = st->wimg[p.layer][0]; q.wimg[1] = st->wimg[p.layer][1];
    q.passes = h->precision == CB_PREC_TC_FAST ? 1 : 3;
    q.write_f32 = write_f32;
    if (o_img) { q.o_img = *o_img; q.write_img = 1; }
    if (q.Bp % RM) { cb_set_error("lstm tensor-core path: padded batch %d not a multiple of %d", q.Bp, RM); return CB_ERR_ARG; }
    lstm_tc_kernel<<<dim3(q.Bp / RM, 2), NTHREADS, SMEM_BYTES, s>>>(q);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}
