// tcgen05 persistent LSTM recurrence (CB_PREC_TC_SPLIT / CB_PREC_TC_FAST).
//
// Replaces the tf.while_loop of dynamic_rnn around LSTMCell (chiron/rnn.py:49-50,64,140-143), like cb_lstm_simt.cu, but
// the per-step contraction h[128 rows,100] x W_hh[100,400] runs on the tensor core, and the 400 gate columns of one
// 128-row group are SPLIT OVER A 2-CTA CLUSTER so that a 4096-window batch keeps 128 SMs busy instead of 64:
//   * gate columns are in unit-major order (half-group hg = hidden units 4hg..4hg+3 = 16 consecutive columns
//     [i0..3 j0..3 f0..3 o0..3]); CTA rank 0 owns half-groups 0-11 (192 columns, units 0-47), rank 1 owns 12-24
//     (208 columns, units 48-99).  Batch rows are independent: no grid-wide synchronisation.
//   * each CTA keeps its slice of W_hh resident in shared memory as fp16 hi/lo K-major core-matrix images (<= 86 KB)
//     and a DOUBLE-BUFFERED copy of the full h operand (2 x 56 KB).  Every step the gate warps write the h K-groups
//     they produce into their own buffer AND, through distributed shared memory (st.shared::cluster), into the peer
//     CTA's buffer, then arrive on both CTAs' mbarriers (release/acquire at cluster scope).
//   * per step one thread issues 21 tcgen05.mma (7 K-steps x {h_hi*W_lo, h_lo*W_hi first, then h_hi*W_hi}) per
//     N-phase into a <= 208-column fp32 TMEM accumulator; the columns are issued in two phases so the gate math of
//     the first phase overlaps the MMAs of the second.
//   * 12 gate warps (3 per TMEM lane quadrant, 2 K-groups of 8 units each) read the accumulator with tcgen05.ld.x16,
//     add the hoisted, pre-scaled input projection (register prefetch one half-group ahead + L2 prefetch two steps
//     ahead), evaluate the cell with MUFU ex2/rcp (7 per cell), keep c in spare TMEM columns, and write h back as
//     16-byte core-matrix rows.
// Global layouts are time-major with the batch (almost) innermost -- pre[T][50][Bp][16], out[T][50][Bp][4] -- so that the
// 32 rows of a warp read and write contiguous, 128-bit vectorised segments at every step.
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
constexpr int C_COL = 256;         // TMEM columns 256.. hold the cell state c[row][local unit]
constexpr int GATE_WARPS = 12;
constexpr int NTHREADS = (1 + GATE_WARPS) * 32;
constexpr uint32_t W_BYTES = KG_A * NMAX * 16;    // one of hi / lo (14th K-group = zeros the K padding multiplies with)
constexpr uint32_t HS_BYTES = KG_A * RM * 16;     // one of hi / lo of one h buffer: 28,672
constexpr size_t SMEM_BYTES = 2 * (size_t)W_BYTES + 4 * (size_t)HS_BYTES + 64;

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
    long long* dbg;            // optional timeline probe (development): clock64 stamps of a few steps of CTA (0,0)
    int dbg_flags;             // development experiments: 1 = no pre loads, 2 = no image stores, 4 = no L2 prefetch
};

__device__ __forceinline__ uint32_t cluster_ctarank() { uint32_t r; asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r)); return r; }
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r; asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank)); return r;
}
__device__ __forceinline__ void st_cluster_v4(uint32_t addr, const uint4& v) {
    asm volatile("st.shared::cluster.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
// shared::cta -> peer CTA's shared memory, completion counted in bytes on the peer's mbarrier (all cluster addresses)
__device__ __forceinline__ void bulk_s2s_cluster(uint32_t dst_cluster, const void* src, uint32_t bytes, uint32_t mbar_cluster) {
    asm volatile("cp.async.bulk.shared::cluster.shared::cta.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst_cluster),
                 "r"(smem_u32(src)), "r"(bytes), "r"(mbar_cluster)
                 : "memory");
}
__device__ __forceinline__ void fence_proxy_async_all() { asm volatile("fence.proxy.async;" ::: "memory"); }   // incl. shared::cluster
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
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

// One half-group (4 hidden units, 16 accumulator columns): returns h of its 4 units.
__device__ __forceinline__ void half_group(uint32_t t_lane, uint32_t t_cell, int hl, bool first_step, bool active,
                                           const float4 (&pre)[4], float (&hv)[4]) {
    uint32_t z[16];
    float c[4];
    tmem_ld4(t_cell + hl * 4, c);
    if (!first_step) {
        tmem_ld16(t_lane + hl * 16, z);
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
    tmem_st4(t_cell + hl * 4, c[0], c[1], c[2], c[3]);
}

// Gate loop of one warp: K-groups [kgA0, kgA0+nA) after the first MMA phase, [kgB0, kgB0+nB) after the second.
// Not unrolled over K-groups on purpose (instruction-cache footprint); c lives in TMEM; the input projection of the
// next half-group is prefetched into registers while the current one is evaluated.
__device__ __forceinline__ void gate_loop(const LstmTcParams& q, uint8_t* hbuf, uint64_t* local_done,
                                          uint64_t* acc_ready, uint32_t tmem_base, int warp, int lane,
                                          int dir, int b0, int hg_base, int kgA0, int nA, int kgB0, int nB) {
    const int quad = warp & 3;                      // TMEM lane quadrant this warp may access
    const int row = quad * 32 + lane;
    const int b = b0 + row;
    int len = 0;
    if (b < q.B) { len = q.lens[b]; len = len < 0 ? 0 : (len > q.T ? q.T : len); }
    const size_t Bp = (size_t)q.Bp;
    const size_t hg_stride4 = Bp * 4;               // float4 elements between consecutive half-groups of pre
    const uint32_t t_lane = tmem_base + ((uint32_t)(quad * 32) << 16);
    const uint32_t t_cell = t_lane + C_COL;
    const float4* pre_b = reinterpret_cast<const float4*>(q.pre) + ((size_t)dir * NHG * Bp + b) * 4;   // + ((t*50 + hg)*Bp)*4
    float4* out_b = reinterpret_cast<float4*>(q.out) + (size_t)dir * NHG * Bp + b;                     // + (t*50 + hg)*Bp
    const size_t t_stride4 = (size_t)(2 * NHG) * Bp * 4;              // float4 elements between frames of pre
    const int hgA = 2 * kgA0, hgB = 2 * kgB0;
    const bool tailB = (kgB0 + nB == KG);            // this warp owns K-group 12, whose upper half (units 100..103) is void

    for (int k = 0; k < 2 * nA; ++k) tmem_st4(t_cell + (hgA + k - hg_base) * 4, 0.f, 0.f, 0.f, 0.f);
    for (int k = 0; k < 2 * nB; ++k) tmem_st4(t_cell + (hgB + k - hg_base) * 4, 0.f, 0.f, 0.f, 0.f);
    tmem_st_wait();

    // frame this row works on at step s (inactive rows: frame s, where they write zeros)
    auto frame_of = [&](int s) { return s < len ? (dir ? len - 1 - s : s) : s; };
    auto load_pre = [&](float4 (&dst)[4], const float4* src) {
        if (CB_LSTM_DEV && (q.dbg_flags & 1)) return;
#pragma unroll
        for (int e = 0; e < 4; ++e) dst[e] = __ldg(src + e);
    };
    // The input projection is streamed from HBM exactly once (1.6 KB per row and step).  Register prefetch alone keeps
    // too little in flight for ~1 us of DRAM latency, so every half-group also pulls its lines for the step after next
    // into L2 (one prefetch instruction per warp and 2 KB).
    auto prefetch_l2 = [&](const float4* src) { if (!(CB_LSTM_DEV && (q.dbg_flags & 4))) asm volatile("prefetch.global.L2 [%0];" ::"l"(src)); };

    float4 pre_e[4], pre_o[4];                       // even / odd half-group of the K-group being processed
#pragma unroll
    for (int e = 0; e < 4; ++e) { pre_e[e] = make_float4(0.f, 0.f, 0.f, 0.f); pre_o[e] = pre_e[e]; }
    const float4* p_cur = pre_b + (size_t)frame_of(0) * t_stride4;
    load_pre(pre_e, p_cur + (size_t)hgA * hg_stride4);
    load_pre(pre_o, p_cur + (size_t)(hgA + 1) * hg_stride4);
    for (int s = 0; s < q.T; ++s) {
        const bool active = s < len;
        const bool first = s == 0;
        const int t = frame_of(s);
        const float4* p_nxt = pre_b + (size_t)(s + 1 < q.T ? frame_of(s + 1) : t) * t_stride4;
        const float4* p_pf = pre_b + (size_t)(s + 2 < q.T ? frame_of(s + 2) : t) * t_stride4;
        float4* out_t = out_b + (size_t)t * (2 * NHG) * Bp;
        const size_t img_row = (size_t)q.o_img.row0 + (size_t)t * Bp + b;
        // h(s) goes into buffer s&1 (the MMAs of step s+1 read it while h(s+1) fills the other buffer)
        uint8_t* h_hi = hbuf + (size_t)(s & 1) * (2 * HS_BYTES);
        uint8_t* h_lo = h_hi + HS_BYTES;
        const bool probe = CB_LSTM_DEV && q.dbg && blockIdx.x == 0 && blockIdx.y == 0 && lane == 0 && s >= 100 && s < 104;
#pragma unroll 1
        for (int phase = 0; phase < 2; ++phase) {
            if (probe && phase == 0) q.dbg[((s - 100) * 16 + warp) * 8 + 0] = clock64();
            mbar_wait(&acc_ready[phase], s & 1);
            tc_fence_after();
            if (probe && phase == 0) q.dbg[((s - 100) * 16 + warp) * 8 + 1] = clock64();
            const int kg0 = phase ? kgB0 : kgA0, nk = phase ? nB : nA;
#pragma unroll 1
            for (int k = 0; k < nk; ++k) {
                const int kg = kg0 + k, hg = 2 * kg, hl = hg - hg_base;
                const bool void_odd = tailB && phase == 1 && k == nk - 1;
                float hv[8];
                half_group(t_lane, t_cell, hl, first, active, pre_e, *reinterpret_cast<float(*)[4]>(&hv[0]));
                {   // refill the even slot: next K-group of this phase, first K-group of the other phase / the next step
                    const float4* nx;
                    if (k + 1 < nk) nx = p_cur + (size_t)(hg + 2) * hg_stride4;
                    else if (phase == 0) nx = p_cur + (size_t)hgB * hg_stride4;
                    else nx = p_nxt + (size_t)hgA * hg_stride4;
                    load_pre(pre_e, nx);
                    prefetch_l2(p_pf + (size_t)hg * hg_stride4);
                }
                half_group(t_lane, t_cell, hl + 1, first, active && !void_odd, pre_o, *reinterpret_cast<float(*)[4]>(&hv[4]));
                {
                    const float4* nx;
                    if (k + 1 < nk) nx = p_cur + (size_t)(hg + 3) * hg_stride4;
                    else if (phase == 0) nx = p_cur + (size_t)(hgB + 1) * hg_stride4;
                    else nx = p_nxt + (size_t)(hgA + 1) * hg_stride4;
                    load_pre(pre_o, nx);             // (for the void upper half of K-group 12 the values are simply unused)
                    if (!void_odd) prefetch_l2(p_pf + (size_t)(hg + 1) * hg_stride4);
                }
                if (q.write_f32) {
                    out_t[(size_t)hg * Bp] = make_float4(hv[0], hv[1], hv[2], hv[3]);
                    if (!void_odd) out_t[(size_t)(hg + 1) * Bp] = make_float4(hv[4], hv[5], hv[6], hv[7]);
                }
                // h(t) of this (row, K-group) as one 16-byte core-matrix row per hi / lo image: the CTA's buffer (the
                // exchange thread ships it to the peer) and, for layers that feed another layer, the global operand image
                uint4 hi, lo;
                split8(hv, hi, lo);
                const uint32_t off = (uint32_t)kg * (RM * 16) + row * 16;
                *reinterpret_cast<uint4*>(h_hi + off) = hi;
                *reinterpret_cast<uint4*>(h_lo + off) = lo;
                if (q.write_img && !(CB_LSTM_DEV && (q.dbg_flags & 2))) {
                    const size_t goff = ((size_t)(dir * KG + kg) * q.o_img.plane_rows + img_row) * 8;
                    *reinterpret_cast<uint4*>(q.o_img.hi + goff) = hi;
                    *reinterpret_cast<uint4*>(q.o_img.lo + goff) = lo;
                }
            }
            // this warp's h rows of the phase are in the CTA's buffer: let the exchange thread ship them to the peer
            if (phase == 1) tmem_st_wait();
            fence_proxy_async();           // generic-proxy h writes -> visible to the async proxy (bulk copy, tensor core)
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&local_done[phase]);
        }
        if (probe) q.dbg[((s - 100) * 16 + warp) * 8 + 6] = clock64();
        p_cur = p_nxt;
    }
}

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
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 6);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int dir = blockIdx.y;
    const uint32_t rank = cluster_ctarank();         // which column slice of the row group this CTA owns
    const int b0 = (blockIdx.x >> 1) * RM;
    const int hg_base = rank ? 12 : 0;               // half-groups [0,12) | [12,25)
    const int ncol = rank ? 208 : 192;
    const int col0 = rank ? 192 : 0;
    // column phases (MMA issue order) in K-groups of 32 columns: rank 0: {0,1,2}{3,4,5}; rank 1: {6,7,8,9}{10,11,12}
    const int kgP0 = rank ? 6 : 0, nP0 = rank ? 4 : 3, kgP1 = rank ? 10 : 3, nP1 = 3;
    const int NA = nP0 * 32, NB = ncol - NA;         // 96|96 or 128|80 columns

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
        // ============================ MMA issuer =====================================================================
        if (lane == 0) {
            // my rows [col0, col0+ncol) of every K-group of the hi and lo weight images
            mbar_arrive_expect_tx(w_bar, 2 * KG * (uint32_t)ncol * 16);
            for (int kg = 0; kg < KG; ++kg) {
                bulk_g2s(w_hi + (size_t)kg * ncol * 16, q.wimg[dir] + ((size_t)kg * H4 + col0) * 8, ncol * 16, w_bar);
                bulk_g2s(w_lo + (size_t)kg * ncol * 16, q.wimg[dir] + ((size_t)(KG + kg) * H4 + col0) * 8, ncol * 16, w_bar);
            }
            mbar_wait(w_bar, 0);
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
                uint64_t da_hi = 0, da_lo = 0;
                const bool probe = CB_LSTM_DEV && q.dbg && blockIdx.x == 0 && blockIdx.y == 0 && s >= 100 && s < 104;
                if (probe) q.dbg[((s - 100) * 16 + 0) * 8 + 0] = clock64();
                if (s > 0) {                                   // h(0) = 0: the first step has no recurrent term
                    mbar_wait_cluster(h_ready, (s - 1) & 1);  // peer's half of h(s-1) landed (own half: local_done, below)
                    tc_fence_after();
                    const uint32_t hb = smem_u32(hbuf) + (uint32_t)((s - 1) & 1) * (2 * HS_BYTES);
                    da_hi = make_desc(hb, RM * 16, 128); da_lo = make_desc(hb + HS_BYTES, RM * 16, 128);
                }
                if (probe) q.dbg[((s - 100) * 16 + 0) * 8 + 1] = clock64();
#pragma unroll
                for (int phase = 0; phase < 2; ++phase) {
                    if (s > 0) {
                        const uint32_t d = tmem_base + (phase ? NA : 0);
                        const uint32_t idesc = phase ? idescB : idescA;
                        const uint32_t brow = phase ? NA : 0;   // 16 B units
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
                    umma_commit(&acc_ready[phase]);
                    if (probe) q.dbg[((s - 100) * 16 + 0) * 8 + 2 + phase] = clock64();
                }
                // arm this step's receive barrier BEFORE shipping my half (the peer can only send h(s+1) after it got
                // my h(s), so bytes of different steps never meet in one barrier phase)
                mbar_arrive_expect_tx(h_ready, peer_bytes);
                const uint32_t buf_off = (uint32_t)(s & 1) * (2 * HS_BYTES);
#pragma unroll
                for (int phase = 0; phase < 2; ++phase) {
                    mbar_wait(&local_done[phase], s & 1);      // my gate warps wrote the phase's K-groups of h(s)
                    const uint32_t off = buf_off + (phase ? offB : offA), len = phase ? lenB : lenA;
                    bulk_s2s_cluster(hbuf_peer + off, hbuf + off, len, h_ready_peer);                        // hi image rows
                    bulk_s2s_cluster(hbuf_peer + off + HS_BYTES, hbuf + off + HS_BYTES, len, h_ready_peer);  // lo image rows
                }
                tc_fence_after();                              // gate warps released the accumulator (local_done[1])
            }
            mbar_wait_cluster(h_ready, (q.T - 1) & 1);         // the peer's last copies into my shared memory have landed
        }
    } else {
        // ============================ gate warps ======================================================================
        // Warp w may only touch TMEM lanes 32*(w%4)..+31; warps 1..12 give every quadrant three warps (slots 0..2).
        // Slot j takes K-group j of the first column phase and K-group j of the second (rank 1's first phase has four:
        // slot 2 takes two, and its second phase ends with the half-empty K-group 12).
        const int slot = (warp - 1) >> 2;
        int kgA0, nA, kgB0, nB;
        if (rank == 0) { kgA0 = kgP0 + slot; nA = 1; kgB0 = kgP1 + slot; nB = 1; }
        else { kgA0 = kgP0 + slot; nA = slot == 2 ? 2 : 1; kgB0 = kgP1 + slot; nB = 1; }
        gate_loop(q, hbuf, local_done, acc_ready, tmem_base, warp, lane, dir, b0, hg_base, kgA0, nA, kgB0, nB);
    }

    tc_fence_before();
    __syncthreads();
    cluster_sync_all();                // no CTA may exit while its peer can still write into its shared memory
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
    static long long* d_dbg = nullptr;
    const char* probe = getenv("CB_LSTM_PROBE");
    if (probe && !d_dbg) { cudaMalloc(&d_dbg, 4 * 16 * 8 * sizeof(long long)); cudaMemset(d_dbg, 0, 4 * 16 * 8 * sizeof(long long)); }
    q.dbg = probe ? d_dbg : nullptr;
    q.dbg_flags = getenv("CB_LSTM_DBG") ? atoi(getenv("CB_LSTM_DBG")) : 0;
    lstm_tc_kernel<<<dim3(2 * (q.Bp / RM), 2), NTHREADS, SMEM_BYTES, s>>>(q);     // clusters of 2 along x
    CB_CHECK_LAUNCH();
    h->launches++;
    if (q.dbg && p.layer == 0) {          // development probe: print the timeline of steps 100..103 of CTA (0,0)
        long long hbuf[4 * 16 * 8];
        cudaStreamSynchronize(s);
        cudaMemcpy(hbuf, q.dbg, sizeof(hbuf), cudaMemcpyDeviceToHost);
        const long long t0 = hbuf[1];
        for (int st = 0; st < 4; ++st) {
            const long long* m = hbuf + (st * 16) * 8;
            fprintf(stderr, "step %d mma: wait_begin %lld h_ready %lld commitA %lld commitB %lld\n", 100 + st, m[0] - t0, m[1] - t0, m[2] - t0, m[3] - t0);
            for (int w = 1; w <= GATE_WARPS; ++w) {
                const long long* g = hbuf + (st * 16 + w) * 8;
                fprintf(stderr, "   warp %2d: waitA %lld gotA %lld | kg0: tmem_ld %lld sums+ldg %lld cells %lld stores %lld | done %lld arrived %lld\n", w, g[0] - t0, g[1] - t0, g[2] - t0, g[3] - t0, g[4] - t0, g[5] - t0, g[6] - t0, g[7] - t0);
            }
        }
    }
    return CB_OK;
}
