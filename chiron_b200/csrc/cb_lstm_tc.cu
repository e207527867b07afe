// tcgen05 persistent LSTM recurrence (CB_PREC_TC_SPLIT).
//
// Replaces the tf.while_loop of dynamic_rnn around LSTMCell (chiron/rnn.py:49-50,64,140-143), like cb_lstm_simt.cu, but
// the per-step contraction h[128 rows,100] x W_hh[100,400] runs on the tensor core, and the 400 gate columns of one
// 128-row group are SPLIT OVER A 2-CTA CLUSTER so that a 4096-window batch keeps 128 SMs busy instead of 64:
//   * gate columns are in unit-major order (half-group hg = hidden units 4hg..4hg+3 = 16 consecutive columns
//     [i0..3 j0..3 f0..3 o0..3]); CTA rank 0 owns half-groups 0-11 (192 columns, units 0-47), rank 1 owns 12-24
//     (208 columns, units 48-99).  Batch rows are independent: no grid-wide synchronisation.
//   * each CTA keeps its slice of W_hh resident in shared memory as fp16 hi/lo K-major core-matrix images (<= 93 KB)
//     and a DOUBLE-BUFFERED copy of the full h operand (2 x 56 KB).  Every step the gate warps write the h they
//     produce into the CTA's own buffer and one thread ships the finished K-groups to the peer CTA with
//     cp.async.bulk shared::cta -> shared::cluster, byte-counted on the peer's mbarrier.
//   * the fp32 accumulator is DOUBLE-BUFFERED in TMEM by step parity and PRE-LOADED with the hoisted, pre-scaled input
//     projection: while the tensor core works on step s the gate warps tcgen05.st pre(s+1) into the other buffer, and
//     the MMAs of step s+1 accumulate h*W_hh on top of it.  The gate math therefore reads finished pre-activations
//     straight from TMEM -- no global-load latency and no extra registers inside the cell loop.
//   * per step one thread issues 21 tcgen05.mma (7 K-steps x {h_hi*W_lo, h_lo*W_hi first, then h_hi*W_hi}) per
//     column phase; the columns are issued in two phases so the gate math of the first overlaps the MMAs of the second.
//   * 24 gate warps (6 per TMEM lane quadrant = per SM sub-partition; <= 80 registers each) take one half-group of
//     each phase (4 cells per thread, evaluated with MUFU ex2/rcp, 7 per cell -- the MUFU pipe is the bound), keep c in
//     spare TMEM columns, and write h back as 8-byte halves of the 16-byte core-matrix rows.
// Global layouts are time-major with the batch (almost) innermost -- pre[T][200][Bp][4], out[T][50][Bp][4] -- so that the
// 32 rows of a warp read and write contiguous, vectorised segments at every step.
#include <cuda_fp16.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "cb_internal.cuh"
#include "cb_tc_common.cuh"

#ifndef CB_LSTM_DEV
#define CB_LSTM_DEV 0          // 1: timeline probe (CB_LSTM_PROBE) and ablation switches (CB_LSTM_DBG) compiled in
#endif

namespace {

constexpr int RM = 128;            // rows per CTA = UMMA M
constexpr int H = 100, H4 = 400;   // this kernel is specialised for the shipped hidden size
constexpr int KG = 13;             // 16-byte K-groups that hold real data (13*8 = 104 >= 100)
constexpr int KG_A = 14;           // K-groups of the h operand (7 K-steps of 16)
constexpr int NHG = 25;            // half-groups (4 hidden units = 16 gate columns) per direction
constexpr int NMAX = 208;          // widest column slice of a CTA
constexpr int ACC1_COL = 224;      // TMEM: accumulator of even steps at columns [0,208), of odd steps at [224,432)
constexpr int C_COL = 448;         // TMEM columns 448..499 hold the cell state c[row][local unit]
constexpr int OWN_KS0 = 3;         // K-steps [0, 3) multiply rank 0's hidden units (K-groups 0-5 = units 0-47), [3, 7) rank 1's
constexpr int GATE_WARPS = 24;
constexpr int SLOTS = GATE_WARPS / 4;   // gate warps per TMEM lane quadrant
constexpr int NTHREADS = (1 + GATE_WARPS) * 32;
constexpr uint32_t W_BYTES = KG_A * NMAX * 16;    // one of hi / lo (14th K-group = zeros the K padding multiplies with)
constexpr uint32_t HS_BYTES = KG_A * RM * 16;     // one of hi / lo of one h buffer: 28,672
constexpr size_t SMEM_BYTES = 2 * (size_t)W_BYTES + 4 * (size_t)HS_BYTES + 128;

struct LstmTcParams {
    int B, Bp, T;
    const float* pre;          // [T][2*25 half-groups][4 gates i,j,f,o][Bp][4 units]
    const __half* wimg[2];     // per direction: hi image then lo image, [KG][400][8] halfs each
    const int32_t* lens;       // [B]
    float* out;                // [T][2*25][Bp][4] fp32 (written when write_f32: the last layer, read by the logit head)
    CbImg o_img;               // hi/lo operand image of h for the next layer's input projection (when write_img):
                               //   plane dir*13 + kg, row row0 + t*Bp + b
    int write_f32, write_img;
    long long* dbg;            // optional timeline probe (development): clock64 stamps of a few steps of CTA (0,0)
    int dbg_flags;             // development experiments: 1 = no pre loads, 2 = no image stores, 4 = no L2 prefetch
};

// shared::cta -> peer CTA's shared memory, completion counted in bytes on the peer's mbarrier (all cluster addresses)
__device__ __forceinline__ void bulk_s2s_cluster(uint32_t dst_cluster, const void* src, uint32_t bytes, uint32_t mbar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
                 "r"(smem_u32(src)), "r"(bytes), "r"(mbar_cluster)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }   // incl. shared::cluster
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

__device__ __forceinline__ void tmem_st16(uint32_t taddr, const float4 (&v)[4]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(__float_as_uint(v[0].x)), "r"(__float_as_uint(v[0].y)), "r"(__float_as_uint(v[0].z)), "r"(__float_as_uint(v[0].w)),
        "r"(__float_as_uint(v[1].x)), "r"(__float_as_uint(v[1].y)), "r"(__float_as_uint(v[1].z)), "r"(__float_as_uint(v[1].w)),
        "r"(__float_as_uint(v[2].x)), "r"(__float_as_uint(v[2].y)), "r"(__float_as_uint(v[2].z)), "r"(__float_as_uint(v[2].w)),
        "r"(__float_as_uint(v[3].x)), "r"(__float_as_uint(v[3].y)), "r"(__float_as_uint(v[3].z)), "r"(__float_as_uint(v[3].w))
        : "memory");
}

// 4 fp32 -> the 8-byte half of a core-matrix row of the hi image and of the lo image
__device__ __forceinline__ void split4(const float (&v)[4], uint2& hi, uint2& lo) {
    uint32_t ph[2], pl[2];
#pragma unroll
    for (int e = 0; e < 2; ++e) {
        const __half2 hh = __floats2half2_rn(v[2 * e], v[2 * e + 1]);
        const float2 hf = __half22float2(hh);
        const __half2 ll = __floats2half2_rn(v[2 * e] - hf.x, v[2 * e + 1] - hf.y);
        ph[e] = *reinterpret_cast<const uint32_t*>(&hh);
        pl[e] = *reinterpret_cast<const uint32_t*>(&ll);
    }
    hi = make_uint2(ph[0], ph[1]);
    lo = make_uint2(pl[0], pl[1]);
}

// Gate loop of one warp.  A warp owns up to three half-groups ("items") of its 32 rows: hlA in the first column phase,
// hlB and (one warp per quadrant of rank 1) hlX in the second; a negative index marks an unused item (warp-uniform).
// Written for <= 72 registers (25 warps per SM): per-step addresses are rebuilt from 32-bit element indices (every
// tensor has < 2^32 16-byte elements, checked by the launcher) instead of being carried as 64-bit pointers, and the
// items are processed one after the other (4 cells of instruction-level parallelism, 6 warps per scheduler).
// REMOTE (CB_LSTM_EXCH=remote, see lstm_tc_kernel): every h value is ALSO stored straight into the peer CTA's buffer
// (st.shared::cluster at hbuf_peer), and a warp signals once per step -- after its last item -- on h_full[s & 1] of BOTH
// CTAs; local_done is unused.
template <bool REMOTE>
__device__ __forceinline__ void gate_loop(const LstmTcParams& q, uint32_t hbuf_s, uint64_t* local_done, uint64_t* acc_ready,
                                          uint64_t* pre_done, uint32_t tmem_base, int warp, int lane, int dir, int b0,
                                          int hg_base, int hlA, int hlB, int hlX, uint32_t hbuf_peer, uint64_t* h_full,
                                          uint32_t h_full_peer) {
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int b = b0 + row;
    int len = 0;
    if (b < q.B) { len = q.lens[b]; len = len < 0 ? 0 : (len > q.T ? q.T : len); }
    const uint32_t Bp = (uint32_t)q.Bp;
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t dhg = (uint32_t)(dir * NHG + hg_base);            // first half-group of this CTA in a frame of pre / out
    const float4* pre4 = reinterpret_cast<const float4*>(q.pre);     // element ((t*50 + hg)*4 + gate)*Bp + b
    float4* out4 = reinterpret_cast<float4*>(q.out);                 // element (t*50 + hg)*Bp + b
    auto item_hl = [&](int i) { return i == 0 ? hlA : (i == 1 ? hlB : hlX); };

#pragma unroll 1
    for (int i = 0; i < 3; ++i)
        if (item_hl(i) >= 0) tmem_st4(t_lane + C_COL + item_hl(i) * 4, 0.f, 0.f, 0.f, 0.f);

    // frame this row works on at step s (inactive rows: frame s, where they write zeros)
    auto frame_of = [&](int s) { return s < len ? (dir ? len - 1 - s : s) : s; };
    // The input projection is streamed from HBM exactly once (1.6 KB per row and step): its lines are pulled into L2
    // four steps ahead (one prefetch instruction per warp and 2 KB); pre(s+2) is requested into registers at the end of
    // step s -- when the registers of the cell math are dead -- and parked in the other accumulator buffer during the
    // idle head of step s+1, while the tensor core is busy with that step's recurrent product.
    // (The seventh half-group of rank 1's second phase, hlX, is not carried in registers -- 72 is all a thread gets -- but
    // loaded inside park_pre.)
    float4 pr[2][4];
#pragma unroll
    for (int i = 0; i < 2; ++i)
#pragma unroll
        for (int e = 0; e < 4; ++e) pr[i][e] = make_float4(0.f, 0.f, 0.f, 0.f);
    auto load_pre = [&](int s) {
        if (CB_LSTM_DEV && (q.dbg_flags & 1)) return;
        const uint32_t f50 = (uint32_t)frame_of(s) * (2 * NHG) + dhg;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float4* p = pre4 + (size_t)((f50 + item_hl(i)) * 4u * Bp + b);
#pragma unroll
            for (int e = 0; e < 4; ++e) pr[i][e] = __ldg(p + (size_t)(e * Bp));
        }
    };
    auto park_pre = [&](int s) {                     // registers -> accumulator buffer of step s
        const uint32_t t_acc = t_lane + ((s & 1) ? ACC1_COL : 0);
        if (hlX >= 0 && !(CB_LSTM_DEV && (q.dbg_flags & 1))) {
            const float4* p = pre4 + (size_t)(((uint32_t)frame_of(s) * (2 * NHG) + dhg + hlX) * 4u * Bp + b);
            float4 px[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) px[e] = __ldg(p + (size_t)(e * Bp));
            tmem_st16(t_acc + hlX * 16, px);
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) tmem_st16(t_acc + item_hl(i) * 16, pr[i]);
        tmem_st_wait();                              // (also covers the c stores of the previous step)
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&pre_done[s & 1]);
    };
    auto prefetch_l2 = [&](int s) {
        if (CB_LSTM_DEV && (q.dbg_flags & 4)) return;
        const uint32_t f50 = (uint32_t)frame_of(s) * (2 * NHG) + dhg;
#pragma unroll
        for (int i = 0; i < 3; ++i)
            if (item_hl(i) >= 0)       // 4 gate segments of 512 B per item: lanes 0..15 take one 128-byte line each
                asm volatile("prefetch.global.L2 [%0];" ::"l"(pre4 + (size_t)(((f50 + item_hl(i)) * 4u + ((lane >> 2) & 3)) * Bp +
                                                                             (b - lane) + (lane & 3) * 8)));
    };

    load_pre(0);
    park_pre(0);                                     // (also orders the c = 0 stores above before anything else)
    if (q.T > 1) load_pre(1);
    for (int s = 2; s < 4 && s < q.T; ++s) prefetch_l2(s);

#pragma unroll 1
    for (int s = 0; s < q.T; ++s) {
        const bool active = s < len;
        const int t = frame_of(s);
        // h(s) goes into buffer s&1 (the MMAs of step s+1 read it while h(s+1) fills the other buffer)
        const uint32_t h_hi = hbuf_s + (uint32_t)(s & 1) * (2 * HS_BYTES) + row * 16;
        const uint32_t t_acc = t_lane + ((s & 1) ? ACC1_COL : 0);
        const bool probe = CB_LSTM_DEV && q.dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && s >= 100 && s < 104;
        if (CB_LSTM_DEV) { if (probe) q.dbg[((s - 100) * 32 + warp) * 8 + 7] = clock64(); __syncwarp(); }
        // idle head of the step (the tensor core is busy with step s): park pre(s+1) in the other accumulator buffer.
        // This warp read its columns of that buffer for the last time in step s-1.
        if (s + 1 < q.T) park_pre(s + 1);
        else { __syncwarp(); if (lane == 0) mbar_arrive(&pre_done[(s + 1) & 1]); }
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            if (CB_LSTM_DEV) { if (probe) q.dbg[((s - 100) * 32 + warp) * 8 + 3 * phase + 0] = clock64(); __syncwarp(); }
            mbar_wait(&acc_ready[phase], s & 1);
            tc_fence_after();
            if (CB_LSTM_DEV) { if (probe) q.dbg[((s - 100) * 32 + warp) * 8 + 3 * phase + 1] = clock64(); __syncwarp(); }
#pragma unroll 1
            for (int i = phase; i < 1 + 2 * phase; ++i) {      // first phase: item 0; second phase: items 1 and 2
                const int hl = item_hl(i);
                if (hl < 0) continue;
                const int hg = hg_base + hl, kg = hg >> 1;
                uint32_t z[16];
                float c[4], hv[4];
                tmem_ld16(t_acc + hl * 16, z);
                tmem_ld4(t_lane + C_COL + hl * 4, c);
                tmem_ld_wait();
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    lstm_cell(__uint_as_float(z[e]), __uint_as_float(z[4 + e]), __uint_as_float(z[8 + e]),
                              __uint_as_float(z[12 + e]), active, c[e], hv[e]);
                tmem_st4(t_lane + C_COL + hl * 4, c[0], c[1], c[2], c[3]);
                if (q.write_f32) out4[(size_t)(((uint32_t)t * (2 * NHG) + dir * NHG + hg) * Bp + b)] = make_float4(hv[0], hv[1], hv[2], hv[3]);
                // h(t) of this (row, half K-group) as the 8-byte half of one core-matrix row per hi / lo image: the
                // CTA's buffer (the exchange thread ships it to the peer) and, for layers that feed another layer,
                // the global operand image
                uint2 hi, lo;
                split4(hv, hi, lo);
                const uint32_t off = h_hi + (uint32_t)kg * (RM * 16) + (hg & 1) * 8;
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(off), "r"(hi.x), "r"(hi.y) : "memory");
                asm volatile("st.shared.v2.b32 [%0], {%1, %2};" ::"r"(off + HS_BYTES), "r"(lo.x), "r"(lo.y) : "memory");
                if constexpr (REMOTE) {
                    const uint32_t offp = off - hbuf_s + hbuf_peer;        // the same place in the peer's buffer
                    asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(offp), "r"(hi.x), "r"(hi.y) : "memory");
                    asm volatile("st.shared::cluster.v2.b32 [%0], {%1, %2};" ::"r"(offp + HS_BYTES), "r"(lo.x), "r"(lo.y) : "memory");
                }
                if (q.write_img && !(CB_LSTM_DEV && (q.dbg_flags & 2))) {
                    // uint2 element ((dir*13 + kg)*plane_rows + row0 + t*Bp + b)*2 + (hg&1)
                    const uint32_t g2 = ((uint32_t)(dir * KG + kg) * (uint32_t)q.o_img.plane_rows + (uint32_t)q.o_img.row0 +
                                         (uint32_t)t * Bp + b) * 2u + (hg & 1);
                    reinterpret_cast<uint2*>(q.o_img.hi)[g2] = hi;
                    reinterpret_cast<uint2*>(q.o_img.lo)[g2] = lo;
                }
            }
            if (CB_LSTM_DEV) { if (probe) q.dbg[((s - 100) * 32 + warp) * 8 + 3 * phase + 2] = clock64(); __syncwarp(); }
            if constexpr (REMOTE) {
                if (phase == 1) {
                    // this warp's h values of the step sit in both CTAs' buffers: generic-proxy writes (shared::cta and
                    // shared::cluster) -> visible to the async proxy (the tensor cores of both CTAs), then one arrival per CTA
                    fence_proxy_async_all();
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) {
                        mbar_arrive(&h_full[s & 1]);
                        mbar_arrive_cluster(h_full_peer + (uint32_t)(s & 1) * 8u);
                    }
                }
            } else {
                // this warp's h rows of the phase are in the CTA's buffer: let the exchange thread ship them to the peer
                fence_proxy_async();       // generic-proxy h writes -> visible to the async proxy (bulk copy, tensor core)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&local_done[phase]);
            }
        }
        if (s + 2 < q.T) load_pre(s + 2);
        if (s + 4 < q.T) prefetch_l2(s + 4);
        if (CB_LSTM_DEV) { if (probe) q.dbg[((s - 100) * 32 + warp) * 8 + 6] = clock64(); __syncwarp(); }
    }
    tmem_st_wait();
}

// REMOTE = false (default): warp 0 waits for the phase's gate warps and ships the finished K-groups with two
// cp.async.bulk shared::cta -> shared::cluster per phase.  REMOTE = true (CB_LSTM_EXCH=remote, kept as a measured negative
// result): the gate threads store every h value into both CTAs' buffers themselves (st.shared / st.shared::cluster) and
// each warp arrives once per step on h_full[s & 1] of both CTAs -- bit-identical, but 3.5 instead of 2.4 ms per launch at
// 4096 x 512 (profiles/r02_s19_lstm_exchange_ab.txt): 8-byte remote stores move far fewer bytes per clock over the
// SM-to-SM network than bulk copies (DSMEM: ~17 B/clk per SM pair), and the proxy fence waits for their acknowledgements.
template <bool REMOTE, bool SPLIT_K>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) lstm_tc_kernel(const LstmTcParams q) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint8_t* w_hi = smem;
    uint8_t* w_lo = smem + W_BYTES;
    uint8_t* hbuf = smem + 2 * (size_t)W_BYTES;      // [2 buffers][hi, lo][14 K-groups][128 rows][16 B]
    uint64_t* bars = reinterpret_cast<uint64_t*>(hbuf + 4 * (size_t)HS_BYTES);
    uint64_t* w_bar = bars;            // weights landed
    uint64_t* h_ready = bars + 1;      // the peer's half of h(t) has landed in this CTA's buffer (byte-counted bulk copies)
    uint64_t* acc_ready = bars + 2;    // [2] MMAs of the first / second column phase of the step retired
    uint64_t* local_done = bars + 4;   // [2] this CTA's gate warps finished the phase (h rows written, accumulator free)
    uint64_t* pre_done = bars + 6;     // [2] by step parity: the gate warps parked the step's input projection in its
                                       //     accumulator buffer (two barriers: a warp arrives for step s+1 without
                                       //     having waited on anything since its arrival for step s)
    uint64_t* h_full = bars + 8;       // [2] REMOTE: by step parity -- every gate warp of BOTH CTAs stored its h(s) here
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const uint32_t rank = cluster_ctarank();         // which column slice of the row group this CTA owns
    const int b0 = (blockIdx.x >> 1) * RM;
    const int hg_base = rank ? 12 : 0;               // half-groups [0,12) | [12,25)
    const int ncol = rank ? 208 : 192;
    const int col0 = rank ? 192 : 0;
    // column phases (MMA issue order) in K-groups of 32 columns: rank 0: {0,1,2}{3,4,5}; rank 1: {6,7,8}{9,10,11,12}
    // (= local half-groups [0,6)[6,12) and [0,6)[6,13))
    const int kgP0 = rank ? 6 : 0, nP0 = 3, kgP1 = rank ? 9 : 3, nP1 = rank ? 4 : 3;
    const int NA = nP0 * 32, NB = ncol - NA;         // 96|96 or 96|112 columns

    // zero both h buffers (h(0) = 0; padding K-groups stay zero) and the weight region (its 14th K-group and unused
    // rows must read as finite zeros: the K padding of the last MMA K-step multiplies them with zero)
    for (uint32_t i = threadIdx.x * 16; i < 2 * W_BYTES + 4 * HS_BYTES; i += NTHREADS * 16)
        *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) {
        mbar_init(w_bar, 1);
        mbar_init(h_ready, 1);
        mbar_init(&local_done[0], GATE_WARPS);
        mbar_init(&local_done[1], GATE_WARPS);
        mbar_init(&acc_ready[0], 1);
        mbar_init(&acc_ready[1], 1);
        mbar_init(&pre_done[0], GATE_WARPS);
        mbar_init(&pre_done[1], GATE_WARPS);
        mbar_init(&h_full[0], 2 * GATE_WARPS);
        mbar_init(&h_full[1], 2 * GATE_WARPS);
        fence_barrier_init();
    }
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();               // the zero fill must be visible to the tensor core / bulk copies
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    cluster_sync_all();                // the peer's barriers and buffers exist before anybody writes to them
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ============================ MMA issuer / exchange (whole warp runs the loop, one elected lane issues) ========
        const bool leader = elect_one();
        if (leader) {
            // my rows [col0, col0+ncol) of every K-group of the hi and lo weight images
            mbar_arrive_expect_tx(w_bar, 2 * KG * (uint32_t)ncol * 16);
            for (int kg = 0; kg < KG; ++kg) {
                bulk_g2s(w_hi + (size_t)kg * ncol * 16, q.wimg[dir] + ((size_t)kg * H4 + col0) * 8, ncol * 16, w_bar);
                bulk_g2s(w_lo + (size_t)kg * ncol * 16, q.wimg[dir] + ((size_t)(KG + kg) * H4 + col0) * 8, ncol * 16, w_bar);
            }
            mbar_wait(w_bar, 0);
        }
        __syncwarp();
        const uint32_t idescA = make_idesc_f16(RM, NA), idescB = make_idesc_f16(RM, NB);
        const uint64_t db_hi = make_desc(smem_u32(w_hi), ncol * 16, 128), db_lo = make_desc(smem_u32(w_lo), ncol * 16, 128);
        constexpr uint32_t A_STEP = 2 * RM;                         // two K-groups per UMMA K-step, in 16 B units
        const uint32_t B_STEP = 2 * (uint32_t)ncol;
        const uint32_t peer = rank ^ 1u;
        const uint32_t hbuf_peer = mapa_shared(smem_u32(hbuf), peer), h_ready_peer = mapa_shared(smem_u32(h_ready), peer);
        // K-groups are 2 KB per image: my phases' chunks, and how many bytes the peer sends me per step
        const uint32_t offA = (uint32_t)kgP0 * (RM * 16), lenA = (uint32_t)nP0 * (RM * 16);
        const uint32_t offB = (uint32_t)kgP1 * (RM * 16), lenB = (uint32_t)nP1 * (RM * 16);
        const uint32_t peer_bytes = 2 * (uint32_t)(rank ? 6 : 7) * (RM * 16);
        for (int s = 0; s < q.T; ++s) {
            if (leader) {
                mbar_wait(&pre_done[s & 1], (s >> 1) & 1);     // pre(s) sits in accumulator buffer s&1
                if (s > 0) {
                    if constexpr (REMOTE) mbar_wait_cluster(&h_full[(s - 1) & 1], ((s - 1) >> 1) & 1);   // all of h(s-1), both halves
                    else if constexpr (!SPLIT_K) mbar_wait_cluster(h_ready, (s - 1) & 1);   // peer's half of h(s-1) landed (own half: local_done, below)
                }
            }
            __syncwarp();
            tc_fence_after();
            const uint32_t hb = smem_u32(hbuf) + (uint32_t)((s + 1) & 1) * (2 * HS_BYTES);     // h(s-1) lives in buffer (s-1)&1
            const uint64_t da_hi = make_desc(hb, RM * 16, 128), da_lo = make_desc(hb + HS_BYTES, RM * 16, 128);
            // K-steps [k0, k1) of one column phase: low-order products first, then hi*hi
            auto issue = [&](int phase, int k0, int k1) {
                const uint32_t d = tmem_base + ((s & 1) ? ACC1_COL : 0) + (phase ? NA : 0);
                const uint32_t idesc = phase ? idescB : idescA;
                const uint32_t brow = phase ? NA : 0;          // 16 B units
                for (int ks = k0; ks < k1; ++ks) {
                    umma_f16(d, da_hi + ks * A_STEP, db_lo + (ks * B_STEP + brow), idesc, 1);
                    umma_f16(d, da_lo + ks * A_STEP, db_hi + (ks * B_STEP + brow), idesc, 1);
                }
                for (int ks = k0; ks < k1; ++ks) umma_f16(d, da_hi + ks * A_STEP, db_hi + (ks * B_STEP + brow), idesc, 1);
            };
            if constexpr (REMOTE || !SPLIT_K) {
#pragma unroll
                for (int phase = 0; phase < 2; ++phase)
                    if (leader) {
                        if (s > 0) issue(phase, 0, KG_A / 2);  // h(0) = 0: the first step has no recurrent term; later steps
                                                               // accumulate on top of the parked projection
                        umma_commit(&acc_ready[phase]);
                    }
            } else {
                // OWN HALF FIRST.  The K-steps split exactly between the CTAs (rank 0's units = K-steps 0-2, rank 1's = 3-6),
                // and this CTA's own half of h(s-1) is complete as soon as its gate warps are (local_done of step s-1, waited
                // for before the copies below were issued).  Its products -- for both column phases -- are issued while the
                // peer's half is still on the wire; only the peer's K-steps of the first column phase (9-12 MMAs instead of
                // 21) remain between the arrival of h and the first gate math of the step.
                const int own0 = rank ? OWN_KS0 : 0, own1 = rank ? KG_A / 2 : OWN_KS0;
                const int peer0 = rank ? 0 : OWN_KS0, peer1 = rank ? OWN_KS0 : KG_A / 2;
                if (leader && s > 0) { issue(0, own0, own1); issue(1, own0, own1); }
                if (leader && s > 0) mbar_wait_cluster(h_ready, (s - 1) & 1);     // peer's half of h(s-1) landed
                __syncwarp();
                tc_fence_after();
#pragma unroll
                for (int phase = 0; phase < 2; ++phase)
                    if (leader) {
                        if (s > 0) issue(phase, peer0, peer1);
                        umma_commit(&acc_ready[phase]);
                    }
            }
            if constexpr (!REMOTE) {
                // arm this step's receive barrier BEFORE shipping my half (the peer can only send h(s+1) after it got
                // my h(s), so bytes of different steps never meet in one barrier phase)
                if (leader) mbar_arrive_expect_tx(h_ready, peer_bytes);
                const uint32_t buf_off = (uint32_t)(s & 1) * (2 * HS_BYTES);
#pragma unroll
                for (int phase = 0; phase < 2; ++phase) {
                    const uint32_t off = buf_off + (phase ? offB : offA), len = phase ? lenB : lenA;
                    if (leader) {
                        mbar_wait(&local_done[phase], s & 1);      // my gate warps wrote the phase's K-groups of h(s)
                        bulk_s2s_cluster(hbuf_peer + off, hbuf + off, len, h_ready_peer);                        // hi image rows
                        bulk_s2s_cluster(hbuf_peer + off + HS_BYTES, hbuf + off + HS_BYTES, len, h_ready_peer);  // lo image rows
                    }
                }
            }
            __syncwarp();
            tc_fence_after();                              // gate warps released the accumulator (local_done[1])
        }
        // the peer's last writes into my shared memory have landed
        if (leader) {
            if constexpr (REMOTE) mbar_wait_cluster(&h_full[(q.T - 1) & 1], ((q.T - 1) >> 1) & 1);
            else mbar_wait_cluster(h_ready, (q.T - 1) & 1);
        }
        __syncwarp();
    } else {
        // ============================ gate warps ======================================================================
        // Warp w may only touch TMEM lanes 32*(w%4)..+31; warps 1..24 give every quadrant six warps (slots 0..5).
        // Slot j takes half-group j of the first phase and half-group j of the second; rank 1's second phase has a
        // seventh half-group, which goes to slot 0.
        const int slot = (warp - 1) >> 2;
        const int hlA = slot, hlB = SLOTS + slot, hlX = (rank && slot == 0) ? 2 * SLOTS : -1;
        const uint32_t peer = rank ^ 1u;
        gate_loop<REMOTE>(q, smem_u32(hbuf), local_done, acc_ready, pre_done, tmem_base, warp, lane, dir, b0, hg_base, hlA, hlB, hlX,
                          mapa_shared(smem_u32(hbuf), peer), h_full, mapa_shared(smem_u32(h_full), peer));
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                // no CTA may exit while its peer can still write into its shared memory
    if (warp == 0) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

}  // namespace

// CB_LSTM_SPLITK=0: one sweep over K per column phase after the whole h arrived (the round-1 issue order; A/B)
bool cb_lstm_tc_split_k() {
    static const bool on = getenv("CB_LSTM_SPLITK") && atoi(getenv("CB_LSTM_SPLITK")) != 0;
    return on;
}

namespace {

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
                        // truncation compensation (cb_tc_build_layer): truncating adds from this K-step's hi*hi product
                        // to the end of the step's chain.  One sweep over K: it is the (g/2 + 1)-th of the 7 last MMAs.
                        // Own half first (the default): a column of rank r sees [own low-order, own hi*hi, peer low-order,
                        // peer hi*hi]; rank 0 owns K-steps 0-2 and columns 0-191.
                        int after = KG_A / 2 - g / 2;
                        if (cb_lstm_tc_split_k()) {
                            const int ks = g / 2, col_rank = n < 192 ? 0 : 1;
                            const bool own = (ks < OWN_KS0) == (col_rank == 0);
                            const int n_own = col_rank ? KG_A / 2 - OWN_KS0 : OWN_KS0, n_peer = KG_A / 2 - n_own;
                            const int j = ks < OWN_KS0 ? ks : ks - OWN_KS0;          // position inside its half
                            after = own ? (n_own - j) + 3 * n_peer : n_peer - j;
                        }
                        const double comp = 1.0 + cb_tc_trunc_c() * after;
                        const float w = k < H ? (float)((double)W[(size_t)k * H4 + gate * H + u] * sc * comp) : 0.f;
                        const __half hi = __float2half_rn(w);
                        const __half lo = __float2half_rn(w - __half2float(hi));
                        img[((size_t)g * H4 + n) * 8 + e] = hi;
                        img[(size_t)KG * H4 * 8 + ((size_t)g * H4 + n) * 8 + e] = lo;
                    }
            CB_CUDA(cudaMalloc(&st->wimg[l][d], img.size() * sizeof(__half)));
            CB_CUDA(cudaMemcpy(st->wimg[l][d], img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
        }
    CB_CUDA(cudaFuncSetAttribute(lstm_tc_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    CB_CUDA(cudaFuncSetAttribute(lstm_tc_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
    CB_CUDA(cudaFuncSetAttribute(lstm_tc_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_BYTES));
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
    q.write_f32 = write_f32;
    if (o_img) { q.o_img = *o_img; q.write_img = 1; }
    if (q.Bp % RM) { cb_set_error("lstm tensor-core path: padded batch %d not a multiple of %d", q.Bp, RM); return CB_ERR_ARG; }
    // the gate warps address pre / out / the operand image with 32-bit element indices
    if ((unsigned long long)q.T * (2 * NHG) * q.Bp * 4ull >= (1ull << 32) ||
        (o_img && (unsigned long long)o_img->planes * o_img->plane_rows * 2ull >= (1ull << 32))) {
        cb_set_error("lstm tensor-core path: batch %d x %d frames exceeds the 32-bit element index range; split the batch", q.B, q.T);
        return CB_ERR_ARG;
    }
    static long long* d_dbg = nullptr;
    const char* probe = getenv("CB_LSTM_PROBE");
    if (probe && !d_dbg) { cudaMalloc(&d_dbg, 4 * 32 * 8 * sizeof(long long)); cudaMemset(d_dbg, 0, 4 * 32 * 8 * sizeof(long long)); }
    q.dbg = probe ? d_dbg : nullptr;
    q.dbg_flags = getenv("CB_LSTM_DBG") ? atoi(getenv("CB_LSTM_DBG")) : 0;
    static const bool remote_exchange = getenv("CB_LSTM_EXCH") && !strcmp(getenv("CB_LSTM_EXCH"), "remote");      // A/B
    const dim3 grid(2 * (q.Bp / RM), 2);                 // clusters of 2 along x
    if (remote_exchange) lstm_tc_kernel<true, false><<<grid, NTHREADS, SMEM_BYTES, s>>>(q);
    else if (cb_lstm_tc_split_k()) lstm_tc_kernel<false, true><<<grid, NTHREADS, SMEM_BYTES, s>>>(q);
    else lstm_tc_kernel<false, false><<<grid, NTHREADS, SMEM_BYTES, s>>>(q);
    CB_CHECK_LAUNCH();
    h->launches++;
    if (q.dbg && p.layer == 0) {          // development probe: print the timeline of steps 100..103 of CTA (0,0)
        long long hbuf[4 * 32 * 8];
        cudaStreamSynchronize(s);
        cudaMemcpy(hbuf, q.dbg, sizeof(hbuf), cudaMemcpyDeviceToHost);
        const long long t0 = hbuf[1 * 8 + 7];
        for (int st = 0; st < 4; ++st) {
            const long long* m = hbuf + (st * 32) * 8;
            fprintf(stderr, "step %d (gate warps only; t0 = first stamp of warp 1)\n", 100 + st); (void)m;
            for (int w = 1; w <= GATE_WARPS; ++w) {
                const long long* g = hbuf + (st * 32 + w) * 8;
                fprintf(stderr, "   warp %2d: begin %lld waitA %lld gotA %lld cellsA %lld | waitB %lld gotB %lld cellsB %lld | end %lld\n", w, g[7] - t0, g[0] - t0, g[1] - t0, g[2] - t0, g[3] - t0, g[4] - t0, g[5] - t0, g[6] - t0);
            }
        }
    }
    return CB_OK;
}
