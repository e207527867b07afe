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
// Global layouts are time-major with the batch (almost) innermost -- pre[T][50][Bp][16], out[T][50][Bp][4] -- so that the
// 32 rows of a warp read and write contiguous, 128-bit vectorised segments at every step.
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
constexpr int N0 = 192, N1 = 208;  // N split of the 400 gate columns: K-groups 0-5 | 6-12 in unit-major column order
constexpr int NHG = 25;            // half-groups (4 hidden units = 16 gate columns) per direction
constexpr int C_COL = 400;         // TMEM columns 400..499 hold the cell state c[row][unit]
constexpr int GATE_WARPS = 15;   // 16 warps in total: 128 registers per thread (the register file is per SM quarter)
constexpr int NTHREADS = (1 + GATE_WARPS) * 32;
constexpr uint32_t W_BYTES = KG * H4 * 16;        // one of hi / lo: 83,200
constexpr uint32_t HS_BYTES = KG_A * RM * 16;     // one of hi / lo: 28,672
constexpr size_t SMEM_BYTES = 2 * (size_t)W_BYTES + 2 * (size_t)HS_BYTES + 64;

struct LstmTcParams {
    int B, Bp, T;
    const float* pre;          // [T][2*25 half-groups][Bp][16]: per half-group i0..3 j0..3 f0..3 o0..3 (unit-major)
    const __half* wimg[2];     // per direction: hi image then lo image, [KG][400][8] halfs each
    const int32_t* lens;       // [B]
    float* out;                // [T][2*25][Bp][4] fp32 (written when write_f32: the last layer, read by the logit head)
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
// The gate pre-activations arrive PRE-SCALED (the host folds -log2(e) into the i/f/o columns of W_ih, W_hh and the bias,
// -2*log2(e) into the j columns, and the forget bias into the bias), so yi = -log2e*i etc. feed ex2 directly.
// e^-x only overflows for very negative x: a one-sided clamp (y <= 25*log2e, i.e. e^-x <= e^25 ~ 7e10) keeps the product
// of three denominators below FLT_MAX; for large positive x, e^-x -> 0 and the quotients saturate by themselves.
__device__ __forceinline__ void lstm_cell(float yi, float yj, float yf, float yo, bool active, float& c, float& h) {
    constexpr float L2E = 1.4426950408889634f, YMAX = 25.f * L2E;
    const float ei = ex2f(fminf(yi, YMAX));
    const float ej = ex2f(fminf(yj, YMAX));
    const float ef = ex2f(fminf(yf, YMAX));
    const float di = 1.f + ei, dj = 1.f + ej, df = 1.f + ef;
    const float dij = di * dj;
    // c' = c/df + (1-ej)/(di*dj) = (c*dij + (1-ej)*df) / (df*dij)
    const float cn = fmaf(c, dij, (1.f - ej) * df) * rcpf(df * dij);
    const float eo = ex2f(fminf(yo, YMAX));
    const float ec = ex2f(fminf(-2.f * L2E * cn, YMAX));
    const float hn = (1.f - ec) * rcpf((1.f + eo) * (1.f + ec));
    c = active ? cn : c;                 // dynamic_rnn: state frozen and output zero past sequence_length
    h = active ? hn : 0.f;
}

// The gate loop of one warp.  Gate columns are in UNIT-MAJOR order: half-group hg (hidden units 4hg..4hg+3) owns the
// 16 consecutive columns [i0..i3 j0..j3 f0..f3 o0..o3] -- one tcgen05.ld.x16 from the accumulator, four 128-bit loads
// of the hoisted input projection pre[t][dir*25+hg][Bp][16], one 128-bit store of the fp32 output.  The MMA is issued
// in two N-halves (K-groups 0-5 = 192 columns, K-groups 6-12 = 208 columns); a warp first handles its K-groups of the
// first half (while the tensor core is still working on the second) and then those of the second half.
// The loop is deliberately NOT unrolled over K-groups (a fully unrolled version is ~300 KB of SASS and thrashes the
// instruction cache); the cell state c therefore lives in the 100 spare TMEM columns next to the accumulator.
struct GateCtx {
    const LstmTcParams* q;
    uint8_t *h_hi, *h_lo;
    uint32_t t_lane, t_cell;
    int row, dir;
    size_t hg_stride4;                  // float4 elements between consecutive half-groups of pre: Bp*4
    size_t Bp;
};

// One half-group: returns h of its 4 units.  `pre` holds the 16 pre-scaled input-projection values.
__device__ __forceinline__ void half_group(const GateCtx& g, int hg, bool first_step, bool active, const float4 (&pre)[4],
                                           float4* out_t, float (&hv)[4]) {
    uint32_t z[16];
    float c[4];
    tmem_ld4(g.t_cell + hg * 4, c);
    if (!first_step) {
        tmem_ld16(g.t_lane + hg * 16, z);
    } else {
#pragma unroll
        for (int e = 0; e < 16; ++e) z[e] = 0u;
    }
    tmem_ld_wait();
    const float pi[4] = {pre[0].x, pre[0].y, pre[0].z, pre[0].w};
    const float pj[4] = {pre[1].x, pre[1].y, pre[1].z, pre[1].w};
    const float pf[4] = {pre[2].x, pre[2].y, pre[2].z, pre[2].w};
    const float po[4] = {pre[3].x, pre[3].y, pre[3].z, pre[3].w};
#pragma unroll
    for (int e = 0; e < 4; ++e)
        lstm_cell(__uint_as_float(z[e]) + pi[e], __uint_as_float(z[4 + e]) + pj[e], __uint_as_float(z[8 + e]) + pf[e],
                  __uint_as_float(z[12 + e]) + po[e], active, c[e], hv[e]);
    tmem_st4(g.t_cell + hg * 4, c[0], c[1], c[2], c[3]);
    if (g.q->write_f32) out_t[(size_t)hg * g.Bp] = make_float4(hv[0], hv[1], hv[2], hv[3]);
}

__device__ __forceinline__ void gate_loop(const LstmTcParams& q, uint8_t* h_hi, uint8_t* h_lo, uint64_t* h_ready,
                                          uint64_t* acc_ready, uint32_t tmem_base, int warp, int lane, int dir, int b0,
                                          int kgA0, int nA, int kgB0, int nB) {
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    GateCtx g;
    g.q = &q; g.h_hi = h_hi; g.h_lo = h_lo; g.dir = dir;
    g.row = quad * 32 + lane;
    const int b = b0 + g.row;
    int len = 0;
    if (b < q.B) { len = q.lens[b]; len = len < 0 ? 0 : (len > q.T ? q.T : len); }
    const size_t Bp = (size_t)q.Bp;
    g.Bp = Bp; g.hg_stride4 = Bp * 4;
    g.t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    g.t_cell = g.t_lane + C_COL;                    // c[row][u] at column C_COL + u
    const float4* pre_b = reinterpret_cast<const float4*>(q.pre) + ((size_t)dir * NHG * Bp + b) * 4;   // + ((t*50 + hg)*Bp)*4
    float4* out_b = reinterpret_cast<float4*>(q.out) + (size_t)dir * NHG * Bp + b;                     // + (t*50 + hg)*Bp
    const size_t t_stride4 = (size_t)(2 * NHG) * Bp * 4;              // float4 elements between frames of pre
    const int hgA = 2 * kgA0, hgB = 2 * kgB0;
    const bool tailB = (kgB0 + nB == KG);            // this warp owns K-group 12, whose upper half (units 100..103) is void

    for (int k = 0; k < nA; ++k) { tmem_st4(g.t_cell + (hgA + 2 * k) * 4, 0.f, 0.f, 0.f, 0.f); tmem_st4(g.t_cell + (hgA + 2 * k + 1) * 4, 0.f, 0.f, 0.f, 0.f); }
    for (int k = 0; k < nB; ++k) { tmem_st4(g.t_cell + (hgB + 2 * k) * 4, 0.f, 0.f, 0.f, 0.f); if (!(tailB && k == nB - 1)) tmem_st4(g.t_cell + (hgB + 2 * k + 1) * 4, 0.f, 0.f, 0.f, 0.f); }
    tmem_st_wait();

    // frame this row works on at step s (inactive rows: frame s, where they write zeros)
    auto frame_of = [&](int s) { return s < len ? (dir ? len - 1 - s : s) : s; };
    auto load_pre = [&](float4 (&dst)[4], const float4* src) {
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[e] = __ldg(src + e);
    };
    // The input projection is streamed from HBM exactly once (1.6 KB per row and step).  Register prefetch keeps only
    // ~30 KB in flight per SM, too little for ~1 us of DRAM latency, so every half-group also pulls its lines for the
    // step after next into L2 (one prefetch instruction per warp and 2 KB).
    auto prefetch_l2 = [&](const float4* src) { asm volatile("prefetch.global.L2 [%0];" ::"l"(src)); };

    float4 pre_e[4], pre_o[4];                       // even / odd half-group of the K-group being processed
    const float4* p_cur = pre_b + (size_t)frame_of(0) * t_stride4;
    load_pre(pre_e, p_cur + (size_t)hgA * g.hg_stride4);
    load_pre(pre_o, p_cur + (size_t)(hgA + 1) * g.hg_stride4);
    for (int s = 0; s < q.T; ++s) {
        const bool active = s < len;
        const bool first = s == 0;
        const int t = frame_of(s);
        const float4* p_nxt = pre_b + (size_t)(s + 1 < q.T ? frame_of(s + 1) : t) * t_stride4;
        const float4* p_pf = pre_b + (size_t)(s + 2 < q.T ? frame_of(s + 2) : t) * t_stride4;
        float4* out_t = out_b + (size_t)t * (2 * NHG) * Bp;
        const size_t img_row = (size_t)q.o_img.row0 + (size_t)t * Bp + b;
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            mbar_wait(&acc_ready[phase], s & 1);
            tc_fence_after();
            const int kg0 = phase ? kgB0 : kgA0, nk = phase ? nB : nA;
#pragma unroll 1
            for (int k = 0; k < nk; ++k) {
                const int kg = kg0 + k, hg = 2 * kg;
                const bool void_odd = tailB && phase == 1 && k == nk - 1;
                float hv[8];
                half_group(g, hg, first, active, pre_e, out_t, *reinterpret_cast<float(*)[4]>(&hv[0]));
                // refill the even slot: next K-group of this phase, first K-group of the other phase, or of the next step
                {
                    const float4* nx;
                    if (k + 1 < nk) nx = p_cur + (size_t)(hg + 2) * g.hg_stride4;
                    else if (phase == 0) nx = p_cur + (size_t)hgB * g.hg_stride4;
                    else nx = p_nxt + (size_t)hgA * g.hg_stride4;
                    load_pre(pre_e, nx);
                    prefetch_l2(p_pf + (size_t)hg * g.hg_stride4);
                }
                if (!void_odd) {
                    half_group(g, hg + 1, first, active, pre_o, out_t, *reinterpret_cast<float(*)[4]>(&hv[4]));
                    prefetch_l2(p_pf + (size_t)(hg + 1) * g.hg_stride4);
                } else {
                    hv[4] = 0.f; hv[5] = 0.f; hv[6] = 0.f; hv[7] = 0.f;
                }
                {
                    const float4* nx;
                    if (k + 1 < nk) nx = p_cur + (size_t)(hg + 3) * g.hg_stride4;
                    else if (phase == 0) nx = p_cur + (size_t)(hgB + 1) * g.hg_stride4;
                    else nx = p_nxt + (size_t)(hgA + 1) * g.hg_stride4;
                    load_pre(pre_o, nx);             // (for the void upper half of K-group 12 the slot is simply unused)
                }
                __half *g_hi = nullptr, *g_lo = nullptr;
                if (q.write_img) {
                    const size_t off = ((size_t)(dir * KG + kg) * q.o_img.plane_rows + img_row) * 8;
                    g_hi = q.o_img.hi + off; g_lo = q.o_img.lo + off;
                }
                store_h_row(h_hi, h_lo, kg, g.row, g_hi, g_lo, hv[0], hv[1], hv[2], hv[3], hv[4], hv[5], hv[6], hv[7]);
            }
        }
        p_cur = p_nxt;
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
    uint64_t* acc_ready = bars + 2;    // [2] MMAs of the first / second N-half of the step retired
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const int b0 = blockIdx.x * RM;

    // zero the h operand (h(0) = 0; padding K-groups stay zero for the whole kernel)
    for (uint32_t i = threadIdx.x * 16; i < 2 * HS_BYTES; i += NTHREADS * 16)
        *reinterpret_cast<uint4*>(h_hi + i) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(w_bar, 1);
        mbar_init(h_ready, GATE_WARPS * 32);
        mbar_init(&acc_ready[0], 1);
        mbar_init(&acc_ready[1], 1);
        fence_barrier_init();
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
            const uint32_t idesc0 = make_idesc_f16(RM, N0), idesc1 = make_idesc_f16(RM, N1);
            // descriptors differ only in the 14-bit start-address field (units of 16 B): advance it with integer adds
            const uint64_t da_hi = make_desc(smem_u32(h_hi), RM * 16, 128), da_lo = make_desc(smem_u32(h_lo), RM * 16, 128);
            const uint64_t db_hi = make_desc(smem_u32(w_hi), H4 * 16, 128), db_lo = make_desc(smem_u32(w_lo), H4 * 16, 128);
            constexpr uint32_t A_STEP = 2 * RM, B_STEP = 2 * H4;       // two K-groups per UMMA K-step, in 16 B units
            for (int s = 0; s < q.T; ++s) {
                if (s > 0) {                                   // h(0) = 0: the first step has no recurrent term
                    mbar_wait(h_ready, (s - 1) & 1);
                    tc_fence_after();
                }
#pragma unroll
                for (int half = 0; half < 2; ++half) {
                    if (s > 0) {
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
                    umma_commit(&acc_ready[half]);
                }
            }
        }
    } else {
        // ============================ gate warps ======================================================================
        // Warp w may only touch TMEM lanes 32*(w%4)..+31.  Quadrant 0 shares its SM quarter with the MMA warp and has
        // three gate warps, quadrants 1-3 have four.  K-groups (8 hidden units) per warp: first N-half {0..5}, second {6..12}.
        const int idx = warp >> 2;
        int kgA0, nA, kgB0, nB;
        if ((warp & 3) == 0) {                 // idx 1..3
            kgA0 = 2 * (idx - 1); nA = 2;
            kgB0 = idx == 1 ? 6 : (idx == 2 ? 8 : 10); nB = idx == 3 ? 3 : 2;
        } else {                               // idx 0..3
            kgA0 = idx == 0 ? 0 : (idx == 1 ? 2 : (idx == 2 ? 4 : 5)); nA = idx < 2 ? 2 : 1;
            kgB0 = idx == 0 ? 6 : (idx == 1 ? 7 : (idx == 2 ? 9 : 11)); nB = idx == 0 ? 1 : 2;
        }
        gate_loop(q, h_hi, h_lo, h_ready, acc_ready, tmem_base, warp, lane, dir, b0, kgA0, nA, kgB0, nB);
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
                        // image row n is unit-major: n = (u/4)*16 + gate*4 + u%4  <-  TF column gate*H + u
                        const int u = (n / 16) * 4 + (n & 3), gate = (n >> 2) & 3;
                        // -log2(e) folded into the i/f/o columns, -2*log2(e) into the j columns (see lstm_cell)
                        const float sc = gate == 1 ? -2.f * 1.4426950408889634f : -1.4426950408889634f;
                        const float w = k < H ? W[(size_t)k * H4 + gate * H + u] * sc : 0.f;
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
