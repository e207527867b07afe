// tcgen05 tensor-core contractions for the conv stack and the hoisted LSTM input projection (CB_PREC_TC_SPLIT).
// Same maths as cb_gemm_simt.cu (chiron/cnn.py:60-82,251-261; chiron/rnn.py:49-50,64), different machine:
//
//   * operands are fp16 hi/lo splits (a = hi + lo, |lo| <= 2^-11 |a|): D = Ah*Wl + Al*Wh + Ah*Wh, three
//     tcgen05.mma.kind::f16 per K-step with fp32 accumulation in TMEM.
//   * SHORT-K PARTIAL SUMS.  The tensor core's fp32 accumulator TRUNCATES on every add, a bias of ~0.4 ulp per MMA
//     that grows with the number of MMAs chained into one accumulator (K/16 x 3 of them; measured: it, not the
//     operand split, sets the logit error of this path).  A TMEM accumulator therefore only ever holds the sum over
//     `cpp` K-chunks (cpp*32 input channels; low-order products issued first, while the accumulator is small), and the
//     epilogue warps add the partial sums in fp32 REGISTERS with round-to-nearest.  The two TMEM buffers ping-pong
//     per partial, so the tensor core fills one while the other is drained.
//   * activations travel between layers as time-major operand images (cb_tc_common.cuh): the epilogue of the producing
//     kernel writes the hi/lo k-group planes; a conv tap is the same plane shifted by one frame = Bp rows, a strided conv
//     multiplies the frame index, the appended 1x1 branch input is a second image.  Block-1 conv2a (a rank-1 function of
//     the raw signal, cnn.py:254) is written as an image by gen_conv2a_kernel.
//   * the A side of a stage (hi and lo, 4 k-group planes x 128 rows x 16 B each) is ONE cp.async.bulk.tensor: the image
//     is described to the TMA unit as a 3-D tensor {256 x 8-byte elements = 128 rows, plane, hi|lo} whose {256,4,2} box
//     lands in shared memory as [hi|lo][4 k-groups][128 rows][8 halfs] -- the UMMA K-major no-swizzle core-matrix order.
//   * persistent CTAs, warp-specialised: 8 epilogue warps (TMEM -> scale/shift/residual/ReLU -> hi/lo image or fp32),
//     one MMA warp, one loader warp (both run warp-convergent; an elected lane issues); shared-memory full/empty ring
//     (q.stages deep); the epilogue of tile i (from registers) overlaps the first two partial sums of tile i+1.
//   * CTA-PAIR form (gemm_tc_pair_kernel, the convolutions): tcgen05 cta_group::2, M = 256 over two m-tiles; each CTA
//     stages its own A tile and half of the B tile, every byte counted on the leader's barrier; multicast commits.
//   * RESIDENT-WEIGHT mode (contractions whose whole n-tile of W fits beside the A ring -- the N = 8H LSTM input
//     projection): a CTA is bound to ONE n-tile, loads its hi/lo weight image once and then streams only A boxes.
// What bounds them: the convolutions run at the board's power cap (0.87 of the sustained cuBLAS bf16
// figure), the input projection at the HBM write bandwidth of its 6.65 GB fp32 output.
#include <cuda.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "cb_internal.cuh"
#include "cb_tc_common.cuh"

#ifndef CB_TC_DEV
#define CB_TC_DEV 0            // 1: timeline probe of the partial-sum hand-over (CB_TC_PROBE=<layer id>) compiled in
#endif

namespace {

constexpr int BM = 128;          // rows (windows of one frame) per tile = UMMA M
constexpr int BK = 32;           // K elements per pipeline stage (2 UMMA K-steps of 16)
constexpr int MIN_STAGES = 4, MAX_STAGES = 6;   // depth of the operand ring (as many as fit in shared memory)
constexpr int TWO_PASS_MAX = 4;  // partial sums of up to this many chunks (<= MIN_STAGES: they are all resident at once) issue the
                                 // low-order products of every chunk first, then the hi*hi ones
constexpr int ACC_COLS = 128;    // accumulator columns an epilogue thread sums in registers (half of the widest tile)
constexpr int N_EPI_WARPS = 8;   // two per TMEM lane quadrant (each takes half of the tile's columns)
// 12 warps = 3 warpgroups: two of epilogue warps, one holding the MMA warp, the loader warp and two idle warps.  The kernel
// is compiled for 168 registers per thread (65536 / 384); setmaxnreg then moves registers from the third warpgroup to the
// epilogue warps, which keep 128 fp32 partial-sum accumulators AND 64 columns of tcgen05.ld results in flight per thread.
// (The single-CTA kernel -- the LSTM input projection, one accumulator per tile, HBM-write-bound -- keeps the plain 10-warp
// layout: measured 2.75 instead of 2.22 ms per launch with the 12-warp layout.)
constexpr int NTHREADS = (N_EPI_WARPS + 4) * 32;   // 384 (CTA-pair kernel)
constexpr int NTHREADS1 = (N_EPI_WARPS + 2) * 32;  // 320 (single-CTA kernel)
constexpr int EPI_REGS = 224, AUX_REGS = 56;       // 256 * 224 + 128 * 56 = 64512 <= 65536

struct TcLayer {                 // one prepared weight image
    __half* img;                 // [n_tiles][k_chunks][2 (hi,lo)][4 k-groups][BN rows][8]
    __half* img2;                // CTA-pair form: [n_tiles][k_chunks][2 (cta)][2 (hi,lo)][4 k-groups][BN/2 rows][8]
    CUtensorMap tm_w2;           // img2 as 2 KB rows (TMA view: a CTA's chunk of a stage = BN/32 rows)
    int K, Kpad, N, BN, n_tiles, k_chunks;
    int cpp;                     // k-chunks per partial sum (the weight image's truncation compensation is built for it)
    float out_scale;             // 2^-s, undoes the power-of-two prescale of the weights
};

struct TcState {
    std::vector<TcLayer> layers; // indexed by layer id
    int* d_range_flag;
    float* d_lstm_bias;          // LSTM biases in the unit-major gate-column order of the recurrence kernel
    size_t lstm_bias_off[CB_MAX_LAYERS][2];
};

struct TcParams {
    CUtensorMap tm_a0, tm_a1;    // TMA views of the operand images a0 / a1 (see make_img_map)
    CUtensorMap tm_w;            // CTA-pair kernel: TMA view of the weight image
    TcGemm g;
    const __half* img;
    int BN, n_tiles, k_chunks, m_tiles;
    float out_scale;
    int stages;                  // depth of the shared-memory operand ring
    int cpp;                     // k-chunks per partial sum (see header): one TMEM accumulator never chains more than
                                 // cpp*2 full-magnitude MMAs; the partial sums are added in registers
    int resident;                // 1: the CTA keeps its n-tile of W in shared memory (see header)
    int* range_flag;
    long long* dbg;              // development probe (CB_TC_DEV): clock64 stamps of the first partial sums of CTA 0
    uint32_t vec_off;            // byte offset of the cached epilogue vectors in shared memory (0 = read them with LDG)
    int early_release;           // several partial sums per tile: give the TMEM buffer back before the output tile is written
    int stagger;                 // units start this many fractions of a tile time apart (0 = together)
};

// Which (m-unit, n-tile) a scheduling unit (a CTA, or a CTA pair) works on in its i-th iteration.  Streaming mode: work
// items round-robin over the units.  Resident mode: n-tile = idx % n_tiles for the whole launch, m-units round-robin over
// the units sharing it.  An m-unit is one 128-row tile (one CTA) or two consecutive ones (CTA pair, M = 256).
struct TileIter {
    int resident, n_tiles, m_units, total, first, step, nt_fixed;
    __device__ TileIter(const TcParams& q, int idx, int cnt, int m_units_) {
        resident = q.resident; n_tiles = q.n_tiles; m_units = m_units_; total = m_units * q.n_tiles;
        if (resident) {
            nt_fixed = idx % n_tiles;
            first = idx / n_tiles;
            step = (cnt - nt_fixed + n_tiles - 1) / n_tiles;
        } else { nt_fixed = 0; first = idx; step = cnt; }
    }
    __device__ __forceinline__ bool get(int i, int& mu, int& nt) const {
        const int v = first + i * step;
        if (resident) { mu = v; nt = nt_fixed; return v < m_units; }
        mu = v / n_tiles; nt = v - mu * n_tiles; return v < total;
    }
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
// 16-byte store of a finished output element group.  CB_TC_STCS=1: streaming (evict-first) stores -- the output image is
// consumed by the NEXT launch, 2 GB later; it has no business staying in L2 (A/B switch).
#ifndef CB_TC_STCS
#define CB_TC_STCS 0
#endif
__device__ __forceinline__ void st_out(uint4* p, const uint4& v) {
#if CB_TC_STCS
    __stcs(p, v);
#else
    *p = v;
#endif
}

__device__ __forceinline__ float4 lds4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// Lean emit of 16 finished sums (tile columns n..n+15 of this thread's row) of a convolution: x 2^-s + shift (+ rank-1
// residual) -> ReLU -> hi/lo split -> two 16-byte stores per k-group plane.  vec_s: shared-memory address of the cached
// epilogue vectors [shift | rw | rinv | rsh][BN]; hi_row / lo_row: this row in plane o_plane0 of the output image (uint4
// units); mx: running maximum of the outputs (the fp16 range check, once per kernel instead of once per element).
template <bool RES>
__device__ __forceinline__ void emit16_image(const float (&sum)[16], int n, uint32_t vec_s, int BN, float out_scale, float xr,
                                             uint4* hi_row, uint4* lo_row, size_t plane_rows, float& mx) {
    float o[16];
#pragma unroll
    for (int q4 = 0; q4 < 4; ++q4) {
        const uint32_t a = vec_s + (uint32_t)(n + q4 * 4) * 4u;
        const float4 sh = lds4(a);
        o[q4 * 4 + 0] = fmaf(sum[q4 * 4 + 0], out_scale, sh.x);
        o[q4 * 4 + 1] = fmaf(sum[q4 * 4 + 1], out_scale, sh.y);
        o[q4 * 4 + 2] = fmaf(sum[q4 * 4 + 2], out_scale, sh.z);
        o[q4 * 4 + 3] = fmaf(sum[q4 * 4 + 3], out_scale, sh.w);
        if constexpr (RES) {
            const float4 w = lds4(a + (uint32_t)BN * 4u), iv = lds4(a + (uint32_t)BN * 8u), rs = lds4(a + (uint32_t)BN * 12u);
            o[q4 * 4 + 0] += fmaf(xr * w.x, iv.x, rs.x);
            o[q4 * 4 + 1] += fmaf(xr * w.y, iv.y, rs.y);
            o[q4 * 4 + 2] += fmaf(xr * w.z, iv.z, rs.z);
            o[q4 * 4 + 3] += fmaf(xr * w.w, iv.w, rs.w);
        }
    }
#pragma unroll
    for (int e = 0; e < 16; ++e) { o[e] = fmaxf(o[e], 0.f); mx = fmaxf(mx, o[e]); }
#pragma unroll
    for (int h8 = 0; h8 < 2; ++h8) {
        float v8[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v8[e] = o[h8 * 8 + e];
        uint4 hi, lo;
        split8(v8, hi, lo);
        const size_t off = (size_t)((n >> 3) + h8) * plane_rows;
        st_out(hi_row + off, hi);
        st_out(lo_row + off, lo);
    }
}

// ---- PTX of the two flavours: NC = 1 (one CTA per tile) and NC = 2 (CTA pair, tcgen05 cta_group::2) -------------------------
template <int NC>
__device__ __forceinline__ void umma_f16_nc(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    if constexpr (NC == 1) {
        umma_f16(d_tmem, a_desc, b_desc, idesc, acc);
    } else {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "setp.ne.b32 p, %4, 0;\n\t"
            "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}"
            ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
            : "memory");
    }
}
// MMA completion -> mbarrier; the pair form arrives on the barrier at the same offset in BOTH CTAs
template <int NC>
__device__ __forceinline__ void umma_commit_nc(uint64_t* bar) {
    if constexpr (NC == 1) {
        umma_commit(bar);
    } else {
        asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                         smem_u32(bar)),
                     "h"((uint16_t)3)
                     : "memory");
    }
}

// one box {256 x u64, 4 planes, hi|lo} = 16 KB of the image -> [hi|lo][4][128 rows][16 B] in shared memory
__device__ __forceinline__ void tma_img_g2s(void* dst, const CUtensorMap* tm, int row, int plane, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(row * 2), "r"(plane), "r"(0), "r"(smem_u32(bar))
        : "memory");
}
// CTA-pair forms: data into this CTA's shared memory, completion bytes onto the LEADER CTA's barrier (cluster address)
__device__ __forceinline__ void tma_img_g2s_pair(void* dst, const CUtensorMap* tm, int row, int plane, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(row * 2), "r"(plane), "r"(0), "r"(bar_cluster)
        : "memory");
}
__device__ __forceinline__ void tma_w_g2s_pair(void* dst, const CUtensorMap* tm, int row, uint32_t bar_cluster) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(
            smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(tm)), "r"(0), "r"(row), "r"(bar_cluster)
        : "memory");
}

// shared memory: q.stages x { A_hi[4][128][8], A_lo, B_hi[4][BNL][8], B_lo } halfs (BNL = BN / NC rows of the B tile live in
// this CTA), then the barriers; resident mode: k_chunks x {B_hi, B_lo} first, stages hold A only.
// MULTI: several partial sums per tile (q.cpp < q.k_chunks); the single-accumulator instantiation carries none of the
// register-accumulation code (with it compiled in, the HBM-write-bound input projection ran 2.7 instead of 2.2 ms).
// EPI: epilogue code path.  0 = generic (every TcGemm flag read at run time).  1 / 2 = the convolutions' path (ReLU, operand
// image out, one n-tile, epilogue vectors cached in shared memory; 2 = with the rank-1 residual of block-1 conv2c): the same
// arithmetic in the same order, compiled without the run-time flag tests, the generic-to-shared address conversions and the
// per-store 64-bit address products of the generic lambda.  ncu (profiles/r02_e1_ncu_gemm_source.csv.gz) showed the generic
// emit as 383 SASS instructions per 16 columns = 3,064 per output tile and warp: with two epilogue warps per scheduler the
// emit is ISSUE-bound (10.6 k clocks per tile, in which the tensor pipe can only run two partial sums ahead).
template <int NC, bool MULTI, int EPI>
__device__ __forceinline__ void gemm_tc_body(const TcParams& q) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const TcGemm& g = q.g;
    const int BN = q.BN, BNL = q.BN / NC;
    const int STAGES = q.stages;
    constexpr uint32_t a_bytes = BM * BK * 2;             // one of hi / lo
    const uint32_t b_bytes = (uint32_t)BNL * BK * 2;
    const uint32_t stage_bytes = q.resident ? 2 * a_bytes : 2 * a_bytes + 2 * b_bytes;
    const uint32_t res_bytes = q.resident ? (uint32_t)q.k_chunks * 2 * b_bytes : 0;
    uint8_t* ring = smem + res_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)STAGES * stage_bytes);
    uint64_t* full_bar = bars;                            // [STAGES]  operands landed (pair: in BOTH CTAs; leader's barrier)
    uint64_t* empty_bar = bars + MAX_STAGES;              // [STAGES]  MMAs that read the stage retired
    uint64_t* acc_full = bars + 2 * MAX_STAGES;           // [2]       partial sum ready for the epilogue warps
    uint64_t* acc_empty = bars + 2 * MAX_STAGES + 2;      // [2]       partial sum read (pair: by both CTAs; leader's)
    uint64_t* w_bar = bars + 2 * MAX_STAGES + 4;          //           resident weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = NC == 2 ? cluster_ctarank() : 0;            // pair: rank 0 = leader (issues the MMAs)
    const TileIter tiles(q, (int)blockIdx.x / NC, (int)gridDim.x / NC, q.m_tiles / NC);
    const int tiles_per_frame = g.Bp / BM;
    const int n_part = (q.k_chunks + q.cpp - 1) / q.cpp;  // partial sums per tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], NC * N_EPI_WARPS); }
        mbar_init(w_bar, 1);
        fence_barrier_init();
        if (q.resident) {                                 // this CTA's part of its n-tile of W: one contiguous image
            const uint32_t chunk = 2 * b_bytes;
            mbar_arrive_expect_tx(w_bar, (uint32_t)q.k_chunks * chunk);
            const int nt = tiles.nt_fixed;
            for (int kc = 0; kc < q.k_chunks; ++kc)
                bulk_g2s(smem + (size_t)kc * chunk, q.img + (((size_t)nt * q.k_chunks + kc) * NC + rank) * (chunk / 2), chunk, w_bar);
            mbar_wait(w_bar, 0);
        }
    }
    if (warp == N_EPI_WARPS) {                            // MMA warp owns the TMEM allocation (all 512 columns)
        if constexpr (NC == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        }
    }
    if (q.vec_off) {
        float* v = reinterpret_cast<float*>(smem + q.vec_off);
        for (int i = threadIdx.x; i < BN; i += (int)blockDim.x) {
            v[i] = __ldg(g.shift + i);
            if (g.res) {                                  // (the launcher reserves the three residual vectors only then)
                v[BN + i] = __ldg(g.rw + i);
                v[2 * BN + i] = __ldg(g.rinv + i);
                v[3 * BN + i] = __ldg(g.rsh + i);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if constexpr (NC == 2) cluster_sync_all();            // both CTAs' barriers, weights and TMEM exist before any remote signal
    const uint32_t tmem_base = *tmem_slot;
    // epilogue vectors (BN shift, rank-1 residual terms) of a contraction whose N is one tile: read from shared memory in the
    // emit (the per-16-column LDG round trips were the largest stall of the epilogue warps: ncu r02_s11)
    float* epi_vec = reinterpret_cast<float*>(smem + q.vec_off);
    const bool vec_cached = q.vec_off != 0;

    if constexpr (NC == 2 && MULTI) {
        if (warp < N_EPI_WARPS) asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(EPI_REGS));
        else asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(AUX_REGS));
    }

    if (warp < N_EPI_WARPS) {
        // ============================ epilogue: partial sums TMEM -> registers (+=) -> scale/shift/residual/ReLU -> HBM =====
        const int quad = warp & 3, chalf = warp >> 2;     // TMEM lane quadrant, which half of the tile's columns
        const int first = ((BN / 16 + 1) / 2) * 16;       // BN is a multiple of 16; the two column halves are 16-col aligned
        const int cbeg = chalf ? first : 0, cend = chalf ? BN : first;      // cend - cbeg <= ACC_COLS
        uint32_t acc_empty_leader[2];
        for (int b = 0; b < 2; ++b)
            acc_empty_leader[b] = NC == 2 ? mapa_shared(smem_u32(&acc_empty[b]), 0) : smem_u32(&acc_empty[b]);
        uint32_t it = 0, pit = 0;                          // tiles / partial sums this unit has worked on
        bool overflow = false;
        float out_max = 0.f;                               // EPI > 0: largest output written (>= 0 after the ReLU)
        const uint32_t vec_s = smem_u32(smem + q.vec_off);
        for (int mu, nt; tiles.get((int)it, mu, nt); ++it) {
            const int mt = mu * NC + (int)rank;
            const int to = mt / tiles_per_frame;
            const int b = (mt - to * tiles_per_frame) * BM + quad * 32 + lane;
            const bool row_ok = b < g.B;
            float xr = 0.f;
            if (g.res && row_ok) xr = __ldg(g.xT + (size_t)to * g.res_stride * g.Bp + b);
            const size_t orow = (size_t)g.o.row0 + (size_t)to * g.Bp + b;
            uint4* const hi_row = reinterpret_cast<uint4*>(g.o.hi) + (size_t)g.o_plane0 * g.o.plane_rows + orow;
            uint4* const lo_row = reinterpret_cast<uint4*>(g.o.lo) + (size_t)g.o_plane0 * g.o.plane_rows + orow;
            // 16 finished sums (columns n0..n0+15 of this thread's row) -> scale/shift/residual/ReLU -> HBM
            auto emit = [&](int n0, const float (&sum)[16]) {
                if constexpr (EPI > 0) {                  // (one n-tile: n0 is the tile column)
                    emit16_image<EPI == 2>(sum, n0, vec_s, BN, q.out_scale, xr, hi_row, lo_row, (size_t)g.o.plane_rows, out_max);
                    return;
                }
                float o[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int n = n0 + q4 * 4;
                    float4 sh;
                    if (vec_cached) sh = *reinterpret_cast<const float4*>(epi_vec + n); else sh = ldg4(g.shift + n);
                    o[q4 * 4 + 0] = fmaf(sum[q4 * 4 + 0], q.out_scale, sh.x);
                    o[q4 * 4 + 1] = fmaf(sum[q4 * 4 + 1], q.out_scale, sh.y);
                    o[q4 * 4 + 2] = fmaf(sum[q4 * 4 + 2], q.out_scale, sh.z);
                    o[q4 * 4 + 3] = fmaf(sum[q4 * 4 + 3], q.out_scale, sh.w);
                    if (g.res) {
                        float4 w, iv, rs;
                        if (vec_cached) {
                            w = *reinterpret_cast<const float4*>(epi_vec + BN + n);
                            iv = *reinterpret_cast<const float4*>(epi_vec + 2 * BN + n);
                            rs = *reinterpret_cast<const float4*>(epi_vec + 3 * BN + n);
                        } else {
                            w = ldg4(g.rw + n); iv = ldg4(g.rinv + n); rs = ldg4(g.rsh + n);
                        }
                        o[q4 * 4 + 0] += fmaf(xr * w.x, iv.x, rs.x);
                        o[q4 * 4 + 1] += fmaf(xr * w.y, iv.y, rs.y);
                        o[q4 * 4 + 2] += fmaf(xr * w.z, iv.z, rs.z);
                        o[q4 * 4 + 3] += fmaf(xr * w.w, iv.w, rs.w);
                    }
                }
                if (g.relu) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) o[e] = fmaxf(o[e], 0.f);
                }
                if (g.out_mode == 2) {                    // hi/lo operand image for the next contraction
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8) {
                        float v8[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            v8[e] = o[h8 * 8 + e];
                            overflow |= !(fabsf(v8[e]) <= 65504.f);
                        }
                        uint4 hi, lo;
                        split8(v8, hi, lo);
                        const size_t off = ((size_t)(g.o_plane0 + (n0 >> 3) + h8) * g.o.plane_rows + orow) * 8;
                        st_out(reinterpret_cast<uint4*>(g.o.hi + off), hi);
                        st_out(reinterpret_cast<uint4*>(g.o.lo + off), lo);
                    }
                } else {                                  // fp32 [to][n/4][Bp][4]: every store instruction of a warp writes
                                                          // 512 contiguous bytes (16 full sectors)
                    float4* dst = reinterpret_cast<float4*>(g.out) + ((size_t)to * (g.ldo >> 2) + (n0 >> 2)) * g.Bp + b;
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        dst[(size_t)q4 * g.Bp] = make_float4(o[q4 * 4], o[q4 * 4 + 1], o[q4 * 4 + 2], o[q4 * 4 + 3]);
                }
            };
            // This warp has read the buffer (pair: arrival on the leader's barrier).  RELAXED arrival: the default
            // .release semantics make the warp wait until everything it stored before -- the whole output tile -- is
            // visible at CTA / cluster scope (measured: 1,100 clocks per arrival, 2 us after an emit), and nothing of that
            // needs ordering: the only hazard is the tensor core overwriting TMEM columns this warp still reads, and its
            // tcgen05.ld results are complete (tcgen05.wait::ld) before the arrival is even issued.
            auto release = [&](uint32_t buf) {
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (NC == 1)
                        mbar_arrive(&acc_empty[buf]);
                    else
                        asm volatile("mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [%0];" ::"r"(acc_empty_leader[buf]) : "memory");
                }
            };
            if constexpr (!MULTI) {
                // one accumulator per tile: stream it out 16 columns at a time (TMEM -> registers costs tensor-pipe time:
                // measured ~64 B/clk per SM and not overlapped with the MMAs, so every column is read exactly once)
                const uint32_t buf = pit & 1, par = (pit >> 1) & 1;
                ++pit;
                mbar_wait(&acc_full[buf], par);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (uint32_t)BN;
                for (int c0 = cbeg; c0 < cend; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c0, v);
                    tmem_ld_wait();
                    if (!row_ok) continue;
                    float sum[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) sum[e] = __uint_as_float(v[e]);
                    emit(nt * BN + c0, sum);
                }
                release(buf);
            } else {
                // several partial sums per tile: all but the last are added up in registers, the last is streamed out
                float acc[ACC_COLS];
#pragma unroll
                for (int e = 0; e < ACC_COLS; ++e) acc[e] = 0.f;
                for (int p = 0; p < n_part; ++p, ++pit) {
                    const uint32_t buf = pit & 1, par = (pit >> 1) & 1;
                    const bool last = p == n_part - 1;
                    const bool probe = CB_TC_DEV && q.dbg && blockIdx.x == 0 && warp == 0 && lane == 0 && pit >= 24 && pit < 56;
                    if (CB_TC_DEV) { if (probe) q.dbg[(pit - 24) * 8 + 0] = clock64(); __syncwarp(); }
                    mbar_wait(&acc_full[buf], par);
                    tc_fence_after();
                    if (CB_TC_DEV) { if (probe) q.dbg[(pit - 24) * 8 + 1] = clock64(); __syncwarp(); }
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (uint32_t)BN + (uint32_t)cbeg;
                    if (!last || q.early_release) {
                        // 64 columns in flight per wait (a tcgen05.ld round trip is ~190 clocks whatever its width); the
                        // single-CTA kernel has no registers to spare for that (no setmaxnreg): 16 columns
                                constexpr int DEPTH = NC == 2 ? 4 : 1;
#pragma unroll
                        for (int j = 0; j < ACC_COLS / 16; j += DEPTH) {
                            if (cbeg + j * 16 < cend) {            // (warp-uniform)
                                uint32_t v[DEPTH][16];
#pragma unroll
                                for (int d = 0; d < DEPTH; ++d)
                                    if (cbeg + (j + d) * 16 < cend) tmem_ld16(taddr + (j + d) * 16, v[d]);
                                tmem_ld_wait();
#pragma unroll
                                for (int d = 0; d < DEPTH; ++d)
                                    if (cbeg + (j + d) * 16 < cend) {
#pragma unroll
                                        for (int e = 0; e < 16; ++e) acc[(j + d) * 16 + e] += __uint_as_float(v[d][e]);
                                    }
                            }
                        }
                    } else if (!q.early_release) {
#pragma unroll
                        for (int j = 0; j < ACC_COLS / 16; ++j) {
                            if (cbeg + j * 16 < cend) {            // (warp-uniform)
                                uint32_t v[16];
                                tmem_ld16(taddr + j * 16, v);
                                tmem_ld_wait();
                                if (row_ok) {
                                    float sum[16];
#pragma unroll
                                    for (int e = 0; e < 16; ++e) sum[e] = acc[j * 16 + e] + __uint_as_float(v[e]);
                                    emit(nt * BN + cbeg + j * 16, sum);
                                }
                            }
                        }
                    }
                    if (last && q.early_release) {
                        // The output tile leaves from the registers AFTER the TMEM buffer went back to the tensor core: while
                        // this warp is busy with scale/shift/ReLU/split and ~4 KB of stores per thread, the MMAs of the next
                        // tile's second partial sum already run (with the buffer held until the last store was issued, the
                        // tensor pipe idled for most of every emit: measured 7-9 k clocks per tile at 2 chunks per partial sum).
                        if (CB_TC_DEV) { if (probe) q.dbg[(pit - 24) * 8 + 2] = clock64(); __syncwarp(); }
                        release(buf);
                        if (CB_TC_DEV) { if (probe) q.dbg[(pit - 24) * 8 + 3] = clock64(); __syncwarp(); }
                        if (row_ok) {
#pragma unroll
                            for (int j = 0; j < ACC_COLS / 16; ++j) {
                                if (cbeg + j * 16 < cend) {
                                    float sum[16];
#pragma unroll
                                    for (int e = 0; e < 16; ++e) sum[e] = acc[j * 16 + e];
                                    emit(nt * BN + cbeg + j * 16, sum);
                                }
                            }
                        }
                        continue;
                    }
                    if (CB_TC_DEV) { if (probe) q.dbg[(pit - 24) * 8 + 2] = clock64(); __syncwarp(); }
                    release(buf);
                    if (CB_TC_DEV) { if (probe) q.dbg[(pit - 24) * 8 + 3] = clock64(); __syncwarp(); }
                }
            }
        }
        if (overflow || !(out_max <= 65504.f)) atomicExch(q.range_flag, 1);
    } else if (warp == N_EPI_WARPS) {
        // ============================ MMA issuer (whole warp runs the loop, one elected lane issues; pair: leader CTA) ======
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_f16(BM * NC, BN);
        constexpr uint32_t A_STEP = 2 * BM;               // two k-groups per UMMA K-step, in 16-byte units
        const uint32_t B_STEP = 2 * (uint32_t)BNL;
        // a partial sum's chunks are all resident at once (low-order products of every chunk first, then hi*hi) when the
        // ring can hold them and still load ahead; longer partial sums order the products chunk by chunk
        const bool two_pass = q.cpp <= TWO_PASS_MAX;
        uint32_t kit = 0, it = 0, pit = 0;
        auto stage_descs = [&](uint32_t k, int kc, uint64_t& dah, uint64_t& dal, uint64_t& dbh, uint64_t& dbl) {
            const uint32_t s = k % (uint32_t)STAGES;
            const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
            const uint32_t sb = q.resident ? smem_u32(smem) + (uint32_t)kc * 2 * b_bytes : sa + 2 * a_bytes;
            dah = make_desc(sa, BM * 16, 128); dal = make_desc(sa + a_bytes, BM * 16, 128);
            dbh = make_desc(sb, BNL * 16, 128); dbl = make_desc(sb + b_bytes, BNL * 16, 128);
            return s;
        };
        // (Tried and dropped: testing the next partial sum's barriers -- its TMEM buffer, its first operand stage -- while the
        // current one still issues MMAs, to spare the barrier round trips at the boundary: neutral at 4 chunks per partial
        // sum, 7-9 % slower at 2-3, where the extra cluster-scope tests sit between the two issue passes.)
        if (rank == 0) {
            for (int mu, nt; tiles.get((int)it, mu, nt); ++it) {
                for (int p = 0; p < n_part; ++p, ++pit) {
                    const uint32_t buf = pit & 1, par = (pit >> 1) & 1;
                    const bool probe = CB_TC_DEV && q.dbg && blockIdx.x == 0 && leader && pit >= 24 && pit < 56;
                    if (CB_TC_DEV) { if (probe) q.dbg[(pit - 24) * 8 + 4] = clock64(); }
                    if (leader) {                         // the epilogue warps have read the previous contents of this buffer
                        if constexpr (NC == 1) mbar_wait(&acc_empty[buf], par ^ 1); else mbar_wait_cluster(&acc_empty[buf], par ^ 1);
                    }
                    __syncwarp();
                    tc_fence_after();
                    if (CB_TC_DEV) { if (probe) q.dbg[(pit - 24) * 8 + 5] = clock64(); }
                    const uint32_t d_tmem = tmem_base + buf * (uint32_t)BN;
                    const int kc0 = p * q.cpp;
                    const int nck = q.k_chunks - kc0 < q.cpp ? q.k_chunks - kc0 : q.cpp;
                    uint64_t dah, dal, dbh, dbl;
                    if (two_pass) {
                        for (int j = 0; j < nck; ++j) {   // low-order products, while the accumulator is small
                            const uint32_t s = stage_descs(kit + j, kc0 + j, dah, dal, dbh, dbl);
                            if (leader) mbar_wait(&full_bar[s], ((kit + j) / (uint32_t)STAGES) & 1);
                            __syncwarp();
                            tc_fence_after();
                            if (leader) {
#pragma unroll
                                for (int ks = 0; ks < BK / 16; ++ks) {
                                    umma_f16_nc<NC>(d_tmem, dah + ks * A_STEP, dbl + ks * B_STEP, idesc, (j | ks) != 0);
                                    umma_f16_nc<NC>(d_tmem, dal + ks * A_STEP, dbh + ks * B_STEP, idesc, 1);
                                }
                            }
                        }
                        for (int j = 0; j < nck; ++j) {
                            const uint32_t s = stage_descs(kit + j, kc0 + j, dah, dal, dbh, dbl);
                            if (leader) {
#pragma unroll
                                for (int ks = 0; ks < BK / 16; ++ks)
                                    umma_f16_nc<NC>(d_tmem, dah + ks * A_STEP, dbh + ks * B_STEP, idesc, 1);
                                umma_commit_nc<NC>(&empty_bar[s]);    // frees the stage (in both CTAs) once these MMAs retire
                            }
                        }
                    } else {
                        for (int j = 0; j < nck; ++j) {
                            const uint32_t s = stage_descs(kit + j, kc0 + j, dah, dal, dbh, dbl);
                            if (leader) mbar_wait(&full_bar[s], ((kit + j) / (uint32_t)STAGES) & 1);
                            __syncwarp();
                            tc_fence_after();
                            if (leader) {
#pragma unroll
                                for (int ks = 0; ks < BK / 16; ++ks) {
                                    umma_f16_nc<NC>(d_tmem, dah + ks * A_STEP, dbl + ks * B_STEP, idesc, (j | ks) != 0);
                                    umma_f16_nc<NC>(d_tmem, dal + ks * A_STEP, dbh + ks * B_STEP, idesc, 1);
                                    umma_f16_nc<NC>(d_tmem, dah + ks * A_STEP, dbh + ks * B_STEP, idesc, 1);
                                }
                                umma_commit_nc<NC>(&empty_bar[s]);
                            }
                        }
                    }
                    if (leader) umma_commit_nc<NC>(&acc_full[buf]);
                    if (CB_TC_DEV) { if (probe) q.dbg[(pit - 24) * 8 + 6] = clock64(); }
                    kit += (uint32_t)nck;
                }
            }
        }
        __syncwarp();
    } else if (warp == N_EPI_WARPS + 1) {
        // ============================ loader: weight + activation images (TMA) ================================================
        const bool leader = elect_one();
        uint32_t kit = 0;
        const int n0c = g.taps * g.a0_chunks_per_tap;     // k-chunks served by image a0
        if (q.stagger > 0) {
            // Every unit needs the same time per tile, so all SMs would write their output tiles in the same instant, in a
            // burst bounded by the HBM write bandwidth (measured: ~12 k clocks per emit, in which the tensor pipe can run
            // ahead by two partial sums only).  The units therefore start q.stagger-th fractions of a tile time apart and
            // stay out of phase: the emits of different SMs interleave with the loads of the others.
            const long long wait_clk = (long long)(((int)blockIdx.x / NC) % q.stagger) * q.k_chunks * 768 / q.stagger;
            const long long t_begin = clock64();
            while (clock64() - t_begin < wait_clk) { }
        }
        for (int it = 0, mu, nt; tiles.get(it, mu, nt); ++it) {
            const int mt = mu * NC + (int)rank;
            const int to = mt / tiles_per_frame;
            const long long b0 = (long long)(mt - to * tiles_per_frame) * BM;
            for (int kc = 0; kc < q.k_chunks; ++kc, ++kit) {
                const uint32_t s = kit % (uint32_t)STAGES, ph = (kit / (uint32_t)STAGES) & 1;
                // A-side box of k-chunk kc of this CTA's tile
                const CUtensorMap* tm; int row, plane;
                if (kc < n0c) {
                    const int j = kc / g.a0_chunks_per_tap, cc = kc - j * g.a0_chunks_per_tap;
                    row = (int)(g.a0.row0 + ((long long)to * g.stride + j - g.left) * g.Bp + b0);
                    plane = g.a0_plane0 + cc * 4; tm = &q.tm_a0;
                } else {
                    row = (int)(g.a1.row0 + (long long)to * (g.a1_stride > 1 ? g.a1_stride : 1) * g.Bp + b0);
                    plane = g.a1_plane0 + (kc - n0c) * 4; tm = &q.tm_a1;
                }
                uint8_t* st = ring + (size_t)s * stage_bytes;
                if constexpr (NC == 1) {
                    const __half* wsrc = q.img + ((size_t)nt * q.k_chunks + kc) * (2 * (size_t)BN * BK);
                    if (leader) {
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
                        tma_img_g2s(st, tm, row, plane, &full_bar[s]);
                        if (!q.resident) bulk_g2s(st + 2 * a_bytes, wsrc, 2 * b_bytes, &full_bar[s]);
                    }
                } else {
                    // both CTAs fill their own stage; every byte is counted on the leader's barrier
                    const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[s]), 0);
                    const int wrow = (((nt * q.k_chunks + kc) * 2 + (int)rank) * BN) / 32;      // 2 KB rows of the weight image
                    if (leader) {
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * stage_bytes);
                        tma_img_g2s_pair(st, tm, row, plane, full_leader);
                        if (!q.resident) tma_w_g2s_pair(st + 2 * a_bytes, &q.tm_w, wrow, full_leader);
                    }
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (NC == 2) cluster_sync_all();            // the leader's MMAs read the peer's shared memory and TMEM
    if (warp == N_EPI_WARPS) {
        tc_fence_after();
        if constexpr (NC == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

// ---- the single-CTA kernel (LSTM input projection): the body as it was before the CTA-pair kernel grew its register
// accumulation machinery.  Kept verbatim: every later variant of the shared body ran this HBM-write-bound contraction at
// 2.5-2.7 instead of 2.2 ms per launch (it reads its A operand twice from DRAM: the five CTAs that share an m-tile fall out
// of step and lose the L2 / TMA request merging; ncu r02).
// shared memory: q.stages x { A_hi[4][128][8], A_lo, B_hi[4][BNL][8], B_lo } halfs (BNL = BN / NC rows of the B tile live in
// this CTA), then the barriers; resident mode: k_chunks x {B_hi, B_lo} first, stages hold A only.
template <int NC>
__device__ __forceinline__ void gemm_tc_body_v1(const TcParams& q) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const TcGemm& g = q.g;
    const int BN = q.BN, BNL = q.BN / NC;
    const int STAGES = q.stages;
    constexpr uint32_t a_bytes = BM * BK * 2;             // one of hi / lo
    const uint32_t b_bytes = (uint32_t)BNL * BK * 2;
    const uint32_t stage_bytes = q.resident ? 2 * a_bytes : 2 * a_bytes + 2 * b_bytes;
    const uint32_t res_bytes = q.resident ? (uint32_t)q.k_chunks * 2 * b_bytes : 0;
    uint8_t* ring = smem + res_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(ring + (size_t)STAGES * stage_bytes);
    uint64_t* full_bar = bars;                            // [STAGES]  operands landed (pair: in BOTH CTAs; leader's barrier)
    uint64_t* empty_bar = bars + MAX_STAGES;              // [STAGES]  MMAs that read the stage retired
    uint64_t* acc_full = bars + 2 * MAX_STAGES;           // [2]       partial sum ready for the epilogue warps
    uint64_t* acc_empty = bars + 2 * MAX_STAGES + 2;      // [2]       partial sum read (pair: by both CTAs; leader's)
    uint64_t* w_bar = bars + 2 * MAX_STAGES + 4;          //           resident weights landed
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * MAX_STAGES + 5);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t rank = NC == 2 ? cluster_ctarank() : 0;            // pair: rank 0 = leader (issues the MMAs)
    const TileIter tiles(q, (int)blockIdx.x / NC, (int)gridDim.x / NC, q.m_tiles / NC);
    const int tiles_per_frame = g.Bp / BM;
    const int n_part = (q.k_chunks + q.cpp - 1) / q.cpp;  // partial sums per tile

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], NC * N_EPI_WARPS); }
        mbar_init(w_bar, 1);
        fence_barrier_init();
        if (q.resident) {                                 // this CTA's part of its n-tile of W: one contiguous image
            const uint32_t chunk = 2 * b_bytes;
            mbar_arrive_expect_tx(w_bar, (uint32_t)q.k_chunks * chunk);
            const int nt = tiles.nt_fixed;
            for (int kc = 0; kc < q.k_chunks; ++kc)
                bulk_g2s(smem + (size_t)kc * chunk, q.img + (((size_t)nt * q.k_chunks + kc) * NC + rank) * (chunk / 2), chunk, w_bar);
            mbar_wait(w_bar, 0);
        }
    }
    if (warp == N_EPI_WARPS) {                            // MMA warp owns the TMEM allocation (all 512 columns)
        if constexpr (NC == 1) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        }
    }
    // Resident mode binds the CTA to ONE n-tile: its BN entries of the shift vector (the LSTM bias) are read once into shared
    // memory.  Read with LDG per 16 columns they were the largest stall of this kernel -- source-level ncu: 34 % of all
    // samples on the FFMAs that wait for them (profiles/r02_s31_proj_source_sass.csv.gz), i.e. the epilogue, not the HBM
    // write, set the tile time.
    const uint32_t vec_s = smem_u32(smem + q.vec_off);
    const int vec_n0 = tiles.nt_fixed * BN;
    if (q.vec_off) {
        float* v = reinterpret_cast<float*>(smem + q.vec_off);
        for (int i = threadIdx.x; i < BN; i += (int)blockDim.x) v[i] = __ldg(g.shift + vec_n0 + i);
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    if constexpr (NC == 2) cluster_sync_all();            // both CTAs' barriers, weights and TMEM exist before any remote signal
    const uint32_t tmem_base = *tmem_slot;

    if (warp < N_EPI_WARPS) {
        // ============================ epilogue: partial sums TMEM -> registers (+=) -> scale/shift/residual/ReLU -> HBM =====
        const int quad = warp & 3, chalf = warp >> 2;     // TMEM lane quadrant, which half of the tile's columns
        const int first = ((BN / 16 + 1) / 2) * 16;       // BN is a multiple of 16; the two column halves are 16-col aligned
        const int cbeg = chalf ? first : 0, cend = chalf ? BN : first;      // cend - cbeg <= ACC_COLS
        uint32_t acc_empty_leader[2];
        for (int b = 0; b < 2; ++b)
            acc_empty_leader[b] = NC == 2 ? mapa_shared(smem_u32(&acc_empty[b]), 0) : smem_u32(&acc_empty[b]);
        uint32_t it = 0, pit = 0;                          // tiles / partial sums this unit has worked on
        bool overflow = false;
        for (int mu, nt; tiles.get((int)it, mu, nt); ++it) {
            const int mt = mu * NC + (int)rank;
            const int to = mt / tiles_per_frame;
            const int b = (mt - to * tiles_per_frame) * BM + quad * 32 + lane;
            const bool row_ok = b < g.B;
            float xr = 0.f;
            if (g.res && row_ok) xr = __ldg(g.xT + (size_t)to * g.res_stride * g.Bp + b);
            const size_t orow = (size_t)g.o.row0 + (size_t)to * g.Bp + b;
            // 16 finished sums (columns n0..n0+15 of this thread's row) -> scale/shift/residual/ReLU -> HBM
            auto emit = [&](int n0, const float (&sum)[16]) {
                float o[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int n = n0 + q4 * 4;
                    const float4 sh = q.vec_off ? lds4(vec_s + (uint32_t)(n - vec_n0) * 4u) : ldg4(g.shift + n);
                    o[q4 * 4 + 0] = fmaf(sum[q4 * 4 + 0], q.out_scale, sh.x);
                    o[q4 * 4 + 1] = fmaf(sum[q4 * 4 + 1], q.out_scale, sh.y);
                    o[q4 * 4 + 2] = fmaf(sum[q4 * 4 + 2], q.out_scale, sh.z);
                    o[q4 * 4 + 3] = fmaf(sum[q4 * 4 + 3], q.out_scale, sh.w);
                    if (g.res) {
                        const float4 w = ldg4(g.rw + n), iv = ldg4(g.rinv + n), rs = ldg4(g.rsh + n);
                        o[q4 * 4 + 0] += fmaf(xr * w.x, iv.x, rs.x);
                        o[q4 * 4 + 1] += fmaf(xr * w.y, iv.y, rs.y);
                        o[q4 * 4 + 2] += fmaf(xr * w.z, iv.z, rs.z);
                        o[q4 * 4 + 3] += fmaf(xr * w.w, iv.w, rs.w);
                    }
                }
                if (g.relu) {
#pragma unroll
                    for (int e = 0; e < 16; ++e) o[e] = fmaxf(o[e], 0.f);
                }
                if (g.out_mode == 2) {                    // hi/lo operand image for the next contraction
#pragma unroll
                    for (int h8 = 0; h8 < 2; ++h8) {
                        float v8[8];
#pragma unroll
                        for (int e = 0; e < 8; ++e) {
                            v8[e] = o[h8 * 8 + e];
                            overflow |= !(fabsf(v8[e]) <= 65504.f);
                        }
                        uint4 hi, lo;
                        split8(v8, hi, lo);
                        const size_t off = ((size_t)(g.o_plane0 + (n0 >> 3) + h8) * g.o.plane_rows + orow) * 8;
                        *reinterpret_cast<uint4*>(g.o.hi + off) = hi;
                        *reinterpret_cast<uint4*>(g.o.lo + off) = lo;
                    }
                } else {                                  // fp32 [to][n/4][Bp][4]: every store instruction of a warp writes
                                                          // 512 contiguous bytes (16 full sectors)
                    float4* dst = reinterpret_cast<float4*>(g.out) + ((size_t)to * (g.ldo >> 2) + (n0 >> 2)) * g.Bp + b;
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        dst[(size_t)q4 * g.Bp] = make_float4(o[q4 * 4], o[q4 * 4 + 1], o[q4 * 4 + 2], o[q4 * 4 + 3]);
                }
            };
            auto release = [&](uint32_t buf) {            // this warp has read the buffer (pair: on the leader's barrier)
                tc_fence_before();
                __syncwarp();
                if (lane == 0) {
                    if constexpr (NC == 1) mbar_arrive(&acc_empty[buf]); else mbar_arrive_cluster(acc_empty_leader[buf]);
                }
            };
            if (n_part == 1) {
                // one accumulator per tile: stream it out 16 columns at a time (TMEM -> registers costs tensor-pipe time:
                // measured ~64 B/clk per SM and not overlapped with the MMAs, so every column is read exactly once)
                const uint32_t buf = pit & 1, par = (pit >> 1) & 1;
                ++pit;
                mbar_wait(&acc_full[buf], par);
                tc_fence_after();
                const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (uint32_t)BN;
                for (int c0 = cbeg; c0 < cend; c0 += 16) {
                    uint32_t v[16];
                    tmem_ld16(taddr + c0, v);
                    tmem_ld_wait();
                    if (!row_ok) continue;
                    float sum[16];
#pragma unroll
                    for (int e = 0; e < 16; ++e) sum[e] = __uint_as_float(v[e]);
                    emit(nt * BN + c0, sum);
                }
                release(buf);
            } else {
                // several partial sums per tile: all but the last are added up in registers, the last is streamed out
                float acc[ACC_COLS];
#pragma unroll
                for (int e = 0; e < ACC_COLS; ++e) acc[e] = 0.f;
                for (int p = 0; p < n_part; ++p, ++pit) {
                    const uint32_t buf = pit & 1, par = (pit >> 1) & 1;
                    const bool last = p == n_part - 1;
                    mbar_wait(&acc_full[buf], par);
                    tc_fence_after();
                    const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + buf * (uint32_t)BN + (uint32_t)cbeg;
#pragma unroll
                    for (int j = 0; j < ACC_COLS / 16; ++j) {
                        if (cbeg + j * 16 < cend) {                // (warp-uniform)
                            uint32_t v[16];
                            tmem_ld16(taddr + j * 16, v);
                            tmem_ld_wait();
#pragma unroll
                            for (int e = 0; e < 16; ++e) acc[j * 16 + e] += __uint_as_float(v[e]);
                            if (last && row_ok) {
                                float sum[16];
#pragma unroll
                                for (int e = 0; e < 16; ++e) sum[e] = acc[j * 16 + e];
                                emit(nt * BN + cbeg + j * 16, sum);
                            }
                        }
                    }
                    release(buf);
                }
            }
        }
        if (overflow) atomicExch(q.range_flag, 1);
    } else if (warp == N_EPI_WARPS) {
        // ============================ MMA issuer (whole warp runs the loop, one elected lane issues; pair: leader CTA) ======
        const bool leader = elect_one();
        const uint32_t idesc = make_idesc_f16(BM * NC, BN);
        constexpr uint32_t A_STEP = 2 * BM;               // two k-groups per UMMA K-step, in 16-byte units
        const uint32_t B_STEP = 2 * (uint32_t)BNL;
        // a partial sum's chunks are all resident at once (low-order products of every chunk first, then hi*hi) when the
        // ring can hold them and still load ahead; longer partial sums order the products chunk by chunk
        const bool two_pass = q.cpp <= TWO_PASS_MAX;
        uint32_t kit = 0, it = 0, pit = 0;
        auto stage_descs = [&](uint32_t k, int kc, uint64_t& dah, uint64_t& dal, uint64_t& dbh, uint64_t& dbl) {
            const uint32_t s = k % (uint32_t)STAGES;
            const uint32_t sa = smem_u32(ring + (size_t)s * stage_bytes);
            const uint32_t sb = q.resident ? smem_u32(smem) + (uint32_t)kc * 2 * b_bytes : sa + 2 * a_bytes;
            dah = make_desc(sa, BM * 16, 128); dal = make_desc(sa + a_bytes, BM * 16, 128);
            dbh = make_desc(sb, BNL * 16, 128); dbl = make_desc(sb + b_bytes, BNL * 16, 128);
            return s;
        };
        if (rank == 0) {
            for (int mu, nt; tiles.get((int)it, mu, nt); ++it) {
                for (int p = 0; p < n_part; ++p, ++pit) {
                    const uint32_t buf = pit & 1, par = (pit >> 1) & 1;
                    if (leader) {                         // the epilogue warps have read the previous contents of this buffer
                        if constexpr (NC == 1) mbar_wait(&acc_empty[buf], par ^ 1); else mbar_wait_cluster(&acc_empty[buf], par ^ 1);
                    }
                    __syncwarp();
                    tc_fence_after();
                    const uint32_t d_tmem = tmem_base + buf * (uint32_t)BN;
                    const int kc0 = p * q.cpp;
                    const int nck = q.k_chunks - kc0 < q.cpp ? q.k_chunks - kc0 : q.cpp;
                    uint64_t dah, dal, dbh, dbl;
                    if (two_pass) {
                        for (int j = 0; j < nck; ++j) {   // low-order products, while the accumulator is small
                            const uint32_t s = stage_descs(kit + j, kc0 + j, dah, dal, dbh, dbl);
                            if (leader) mbar_wait(&full_bar[s], ((kit + j) / (uint32_t)STAGES) & 1);
                            __syncwarp();
                            tc_fence_after();
                            if (leader) {
#pragma unroll
                                for (int ks = 0; ks < BK / 16; ++ks) {
                                    umma_f16_nc<NC>(d_tmem, dah + ks * A_STEP, dbl + ks * B_STEP, idesc, (j | ks) != 0);
                                    umma_f16_nc<NC>(d_tmem, dal + ks * A_STEP, dbh + ks * B_STEP, idesc, 1);
                                }
                            }
                        }
                        for (int j = 0; j < nck; ++j) {
                            const uint32_t s = stage_descs(kit + j, kc0 + j, dah, dal, dbh, dbl);
                            if (leader) {
#pragma unroll
                                for (int ks = 0; ks < BK / 16; ++ks)
                                    umma_f16_nc<NC>(d_tmem, dah + ks * A_STEP, dbh + ks * B_STEP, idesc, 1);
                                umma_commit_nc<NC>(&empty_bar[s]);    // frees the stage (in both CTAs) once these MMAs retire
                            }
                        }
                    } else {
                        for (int j = 0; j < nck; ++j) {
                            const uint32_t s = stage_descs(kit + j, kc0 + j, dah, dal, dbh, dbl);
                            if (leader) mbar_wait(&full_bar[s], ((kit + j) / (uint32_t)STAGES) & 1);
                            __syncwarp();
                            tc_fence_after();
                            if (leader) {
#pragma unroll
                                for (int ks = 0; ks < BK / 16; ++ks) {
                                    umma_f16_nc<NC>(d_tmem, dah + ks * A_STEP, dbl + ks * B_STEP, idesc, (j | ks) != 0);
                                    umma_f16_nc<NC>(d_tmem, dal + ks * A_STEP, dbh + ks * B_STEP, idesc, 1);
                                    umma_f16_nc<NC>(d_tmem, dah + ks * A_STEP, dbh + ks * B_STEP, idesc, 1);
                                }
                                umma_commit_nc<NC>(&empty_bar[s]);
                            }
                        }
                    }
                    if (leader) umma_commit_nc<NC>(&acc_full[buf]);
                    kit += (uint32_t)nck;
                }
            }
        }
        __syncwarp();
    } else {
        // ============================ loader: weight + activation images (TMA) ================================================
        const bool leader = elect_one();
        uint32_t kit = 0;
        const int n0c = g.taps * g.a0_chunks_per_tap;     // k-chunks served by image a0
        for (int it = 0, mu, nt; tiles.get(it, mu, nt); ++it) {
            const int mt = mu * NC + (int)rank;
            const int to = mt / tiles_per_frame;
            const long long b0 = (long long)(mt - to * tiles_per_frame) * BM;
            for (int kc = 0; kc < q.k_chunks; ++kc, ++kit) {
                const uint32_t s = kit % (uint32_t)STAGES, ph = (kit / (uint32_t)STAGES) & 1;
                // A-side box of k-chunk kc of this CTA's tile
                const CUtensorMap* tm; int row, plane;
                if (kc < n0c) {
                    const int j = kc / g.a0_chunks_per_tap, cc = kc - j * g.a0_chunks_per_tap;
                    row = (int)(g.a0.row0 + ((long long)to * g.stride + j - g.left) * g.Bp + b0);
                    plane = g.a0_plane0 + cc * 4; tm = &q.tm_a0;
                } else {
                    row = (int)(g.a1.row0 + (long long)to * (g.a1_stride > 1 ? g.a1_stride : 1) * g.Bp + b0);
                    plane = g.a1_plane0 + (kc - n0c) * 4; tm = &q.tm_a1;
                }
                uint8_t* st = ring + (size_t)s * stage_bytes;
                if constexpr (NC == 1) {
                    const __half* wsrc = q.img + ((size_t)nt * q.k_chunks + kc) * (2 * (size_t)BN * BK);
                    if (leader) {
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        mbar_arrive_expect_tx(&full_bar[s], stage_bytes);
                        tma_img_g2s(st, tm, row, plane, &full_bar[s]);
                        if (!q.resident) bulk_g2s(st + 2 * a_bytes, wsrc, 2 * b_bytes, &full_bar[s]);
                    }
                } else {
                    // both CTAs fill their own stage; every byte is counted on the leader's barrier
                    const uint32_t full_leader = mapa_shared(smem_u32(&full_bar[s]), 0);
                    const int wrow = (((nt * q.k_chunks + kc) * 2 + (int)rank) * BN) / 32;      // 2 KB rows of the weight image
                    if (leader) {
                        mbar_wait(&empty_bar[s], ph ^ 1);
                        if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * stage_bytes);
                        tma_img_g2s_pair(st, tm, row, plane, full_leader);
                        if (!q.resident) tma_w_g2s_pair(st + 2 * a_bytes, &q.tm_w, wrow, full_leader);
                    }
                }
                __syncwarp();
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if constexpr (NC == 2) cluster_sync_all();            // the leader's MMAs read the peer's shared memory and TMEM
    if (warp == N_EPI_WARPS) {
        tc_fence_after();
        if constexpr (NC == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
        else asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

template <bool MULTI>
__global__ void __launch_bounds__(NTHREADS1, 1) gemm_tc_kernel(const __grid_constant__ TcParams q) {
    if constexpr (MULTI) gemm_tc_body<1, true, 0>(q); else gemm_tc_body_v1<1>(q);
}
template <bool MULTI, int EPI>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NTHREADS, 1) gemm_tc_pair_kernel(const __grid_constant__ TcParams q) {
    gemm_tc_body<2, MULTI, EPI>(q);
}

// x[B][L] -> xT[L][Bp]  (so that everything downstream reads the raw signal coalesced over windows)
__global__ void __launch_bounds__(256) transpose_x_kernel(const float* __restrict__ x, int B, int L, int Bp,
                                                          float* __restrict__ xT) {
    __shared__ float tile[32][33];
    const int b0 = blockIdx.x * 32, f0 = blockIdx.y * 32;
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int b = b0 + i, f = f0 + threadIdx.x;
        tile[i][threadIdx.x] = (b < B && f < L) ? x[(size_t)b * L + f] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += 8) {
        const int f = f0 + i, b = b0 + threadIdx.x;
        if (f < L && b < Bp) xT[(size_t)f * Bp + b] = tile[threadIdx.x][i];
    }
}

// Block-1 conv2a (cnn.py:254, C_in = 1): a = relu((x*w[c])*inv[c] + shift[c]) written straight into an operand image.
// One thread per (frame, window, k-group); consecutive threads = consecutive windows (16-byte coalesced stores).
__global__ void __launch_bounds__(256) gen_conv2a_kernel(const float* __restrict__ xT, int B, int Bp, int L, int planes,
                                                         const float* __restrict__ gw, const float* __restrict__ ginv,
                                                         const float* __restrict__ gsh, CbImg o) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int f = blockIdx.y;
    if (b >= B) return;
    const float xv = __ldg(xT + (size_t)f * Bp + b);
    for (int kg = 0; kg < planes; ++kg) {
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) {
            const int c = kg * 8 + e;
            v[e] = fmaxf(fmaf(xv * __ldg(gw + c), __ldg(ginv + c), __ldg(gsh + c)), 0.f);
        }
        uint4 hi, lo;
        split8(v, hi, lo);
        const size_t off = ((size_t)kg * o.plane_rows + o.row0 + (size_t)f * Bp + b) * 8;
        *reinterpret_cast<uint4*>(o.hi + off) = hi;
        *reinterpret_cast<uint4*>(o.lo + off) = lo;
    }
}

// Stem convolution of RNA_model2 / RNA_model3 (conv_layer(net, [1, k, 1, C], strides = s) + BN + ReLU, cnn.py:454-476) written
// straight into an operand image: the arithmetic of stem_conv_kernel (cb_stem_kernel.cuh: taps in order, fmaf, then the folded
// BN as one fmaf), one thread per (output frame, window, k-group), consecutive threads = consecutive windows.
__global__ void __launch_bounds__(256) stem_image_kernel(const float* __restrict__ xT, int B, int Bp, int L, int t_out, int k,
                                                         int stride, int left, int planes, const float* __restrict__ w,
                                                         const float* __restrict__ inv, const float* __restrict__ shift, CbImg o) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int to = blockIdx.y;
    if (b >= B || to >= t_out) return;
    for (int kg = 0; kg < planes; ++kg) {
        float acc[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) acc[e] = 0.f;
        for (int j = 0; j < k; ++j) {
            const int ti = to * stride + j - left;
            if (ti < 0 || ti >= L) continue;
            const float xv = __ldg(xT + (size_t)ti * Bp + b);
            const float* wj = w + (size_t)j * planes * 8 + kg * 8;
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = fmaf(xv, __ldg(wj + e), acc[e]);
        }
        float v[8];
#pragma unroll
        for (int e = 0; e < 8; ++e) v[e] = fmaxf(fmaf(acc[e], __ldg(inv + kg * 8 + e), __ldg(shift + kg * 8 + e)), 0.f);
        uint4 hi, lo;
        split8(v, hi, lo);
        const size_t off = ((size_t)kg * o.plane_rows + o.row0 + (size_t)to * Bp + b) * 8;
        *reinterpret_cast<uint4*>(o.hi + off) = hi;
        *reinterpret_cast<uint4*>(o.lo + off) = lo;
    }
}

// TMA view of an operand image: dim0 = 8-byte elements along the rows of a plane (2 per 16-byte row), dim1 = plane,
// dim2 = hi | lo (the two allocations of an image; lo lies behind hi in the workspace).
int make_img_map(const CbImg& img, CUtensorMap* tm) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    static EncodeFn encode = nullptr;
    if (!encode) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
            cb_set_error("cuTensorMapEncodeTiled is not available from this driver");
            return CB_ERR_CUDA;
        }
        encode = (EncodeFn)fn;
    }
    const long long hl = (const char*)img.lo - (const char*)img.hi;
    if (hl <= 0 || (hl & 15)) { cb_set_error("operand image: lo plane set must lie behind hi, 16-byte aligned"); return CB_ERR_ARG; }
    const cuuint64_t dims[3] = {(cuuint64_t)img.plane_rows * 2, (cuuint64_t)img.planes, 2};
    const cuuint64_t strides[2] = {(cuuint64_t)img.plane_rows * 16, (cuuint64_t)hl};
    const cuuint32_t box[3] = {256, 4, 2}, estr[3] = {1, 1, 1};
    const CUresult r = encode(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 3, (void*)img.hi, dims, strides, box, estr,
                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { cb_set_error("cuTensorMapEncodeTiled failed (%d)", (int)r); return CB_ERR_CUDA; }
    return CB_OK;
}

// TMA view of a CTA-pair weight image as 2 KB rows: one CTA's chunk of a pipeline stage is BN/32 consecutive rows.
int make_w_map(const __half* img2, size_t halfs, int BN, CUtensorMap* tm) {
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                 CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) != cudaSuccess || !fn) {
        cb_set_error("cuTensorMapEncodeTiled is not available from this driver");
        return CB_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {256, (cuuint64_t)(halfs * 2 / 2048)};
    const cuuint64_t strides[1] = {2048};
    const cuuint32_t box[2] = {256, (cuuint32_t)(BN / 32)}, estr[2] = {1, 1};
    const CUresult r = ((EncodeFn)fn)(tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*)img2, dims, strides, box, estr,
                                      CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                      CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) { cb_set_error("cuTensorMapEncodeTiled(weights) failed (%d)", (int)r); return CB_ERR_CUDA; }
    return CB_OK;
}

size_t smem_bytes_for(int BN, int stages) { return (size_t)stages * (2 * BM * BK * 2 + 2 * (size_t)BN * BK * 2) + 256; }
size_t smem_bytes_resident(int BN, int k_chunks, int stages) {
    return (size_t)k_chunks * 2 * BN * BK * 2 + (size_t)stages * (2 * BM * BK * 2) + 256;
}
// (the CTA-pair kernel holds BN/2 rows of B per CTA: pass BN/2)
constexpr size_t SMEM_MAX = 232448;      // 227 KB of dynamic shared memory per CTA

int pick_bn(int N) {            // widest tile <= 256 that divides N (UMMA N must be a multiple of 16 at M = 128)
    for (int bn = 256; bn >= 16; bn -= 16)
        if (N % bn == 0) return bn;
    return 0;
}

}  // namespace

// ---- host: weight images -------------------------------------------------------------------------------------------------
// W is [K][N] fp32 (row k = input channel in the order the A operand presents it).
// `chain_after`: truncating accumulator adds the OUTPUT of this contraction still goes through downstream (the LSTM
// recurrence chains 21 MMAs on top of the input projection it is pre-loaded with); 0 for the convolutions.
int cb_tc_build_layer(cb_handle* h, int layer_id, const float* W, int K, int N, int chain_after) {
    TcState* st = (TcState*)h->tc;
    TcLayer L;
    memset(&L, 0, sizeof(L));
    L.K = K; L.N = N; L.BN = pick_bn(N);
    if (L.BN == 0) { cb_set_error("tensor-core path: N=%d is not a multiple of 16", N); return CB_ERR_ARG; }
    L.n_tiles = N / L.BN;
    L.k_chunks = (K + BK - 1) / BK;
    L.Kpad = L.k_chunks * BK;
    // k-chunks (of 32 input channels) per partial sum = how many MMAs are chained into one TMEM accumulator (6 per chunk).
    // The hand-over of a partial sum to the epilogue warps is not free (the MMA warp's issue loop, and the output tile of
    // the previous tile still being written while only two partial sums fit in TMEM), so partial sums are as long as the
    // error budget allows.  Measured on the 4096 x 512 bench batch (conv stack ms sustained / logit rms error against the
    // float64 oracle; the fp32 FFMA kernels: 1.1e-5; profiles/r02_*):
    //   convolutions:  1 accumulator per tile 9.5 ms / 1.1e-4    8 chunks (products interleaved) 9.7 / 4.1e-5
    //                  4 chunks 12.0 / 1.8e-5   <- default         3 chunks 12.7 / 1.7e-5      2 chunks 13.7 / 1.7e-5
    //   (largest logit error on 262,144 frames: 1.6e-2 / 6.1e-3 / 5.4e-3 / 3.7e-3 at 8 / 4 / 3 / 2 chunks; fp32 kernels 4.1e-3;
    //    greedy bases identical to the fp32 kernels' on all 4096 windows from 8 chunks down)
    //   LSTM input projections (HBM-write-bound, K <= 256): one accumulator (2.2 ms; 2.9 ms with two partial sums)
    // CB_TC_CPP / CB_TC_CPP_PROJ override (n <= 4: low-order products of the whole partial sum first; n >= k_chunks: one
    // accumulator).
    static const int cpp_conv = getenv("CB_TC_CPP") ? atoi(getenv("CB_TC_CPP")) : 4;
    static const int cpp_proj = getenv("CB_TC_CPP_PROJ") ? atoi(getenv("CB_TC_CPP_PROJ")) : 8;
    // CB_TC_CPP_K256 / CB_TC_CPP_K512: the same for the convolutions with K <= 256 / K = 512 only (the HBM-bound 1x1 layers)
    static const int cpp_k256 = getenv("CB_TC_CPP_K256") ? atoi(getenv("CB_TC_CPP_K256")) : cpp_conv;
    static const int cpp_k512 = getenv("CB_TC_CPP_K512") ? atoi(getenv("CB_TC_CPP_K512")) : cpp_conv;
    const int cpp_env = layer_id >= 32 ? cpp_proj : (L.k_chunks <= 8 ? cpp_k256 : (L.k_chunks <= 16 ? cpp_k512 : cpp_conv));
    L.cpp = cpp_env > 0 && cpp_env < L.k_chunks ? cpp_env : L.k_chunks;
    // TRUNCATION COMPENSATION.  The tensor core's fp32 accumulator rounds toward zero on every MMA, so an accumulated value
    // shrinks by a small relative amount c per chained MMA (round-to-zero of a 24-bit significand loses half an ulp on
    // average = 2^-24 * E[1/m] = 0.72 * 2^-24 for log-uniform significands m; CALIBRATED on the float64 oracle the constant is
    // 0.42 * 2^-24: CNN-feature rms error 3.4e-6 uncompensated, 2.6e-7 at 0.4-0.45 -- profiles/r02_s2_diag, r02_s3_diag).
    // A product that enters the chain at MMA number t of n is truncated (n - t + 1) times, so its weight is inflated by
    // (1 + c*(n - t + 1)): first order, exact in expectation; what remains is the zero-mean part of the rounding errors, as in
    // any fp32 summation.
    const double comp_c = cb_tc_trunc_c();
    const bool two_pass = L.cpp <= TWO_PASS_MAX;         // issue order of a partial sum (see gemm_tc_body)
    auto comp = [&](int kc, int ks) {                   // inflation of the hi*hi products of K-step ks of chunk kc
        const int p = kc / L.cpp, j = kc - p * L.cpp;
        const int nck = (p + 1) * L.cpp <= L.k_chunks ? L.cpp : L.k_chunks - p * L.cpp;
        const int steps = BK / 16;
        // MMAs of the partial sum issued after and including this K-step's hi*hi product
        const int after = two_pass ? (nck * steps - (j * steps + ks)) : 3 * (nck * steps - (j * steps + ks)) - 2;
        return 1.0 + comp_c * (after + chain_after);
    };
    float mx = 0.f;
    for (size_t i = 0; i < (size_t)K * N; ++i) mx = fmaxf(mx, fabsf(W[i]));
    int s = 0;
    if (mx > 0.f) { s = (int)floorf(log2f(8192.0f / mx)); if (s > 24) s = 24; if (s < -8) s = -8; }
    const float scale = ldexpf(1.0f, s);
    L.out_scale = ldexpf(1.0f, -s);
    const size_t per_chunk = 2 * (size_t)L.BN * BK;       // hi + lo
    std::vector<__half> img((size_t)L.n_tiles * L.k_chunks * per_chunk);
    for (int nt = 0; nt < L.n_tiles; ++nt)
        for (int kc = 0; kc < L.k_chunks; ++kc) {
            __half* base = img.data() + ((size_t)nt * L.k_chunks + kc) * per_chunk;
            for (int gq = 0; gq < 4; ++gq)
                for (int n = 0; n < L.BN; ++n)
                    for (int e = 0; e < 8; ++e) {
                        const int k = kc * BK + gq * 8 + e;
                        const float w = k < K ? (float)((double)W[(size_t)k * N + nt * L.BN + n] * scale * comp(kc, gq / 2)) : 0.f;
                        const __half hi = __float2half_rn(w);
                        const __half lo = __float2half_rn(w - __half2float(hi));
                        base[(size_t)gq * L.BN * 8 + n * 8 + e] = hi;
                        base[(size_t)L.BN * BK + (size_t)gq * L.BN * 8 + n * 8 + e] = lo;
                    }
        }
    CB_CUDA(cudaMalloc(&L.img, img.size() * sizeof(__half)));
    CB_CUDA(cudaMemcpy(L.img, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
    if (L.BN % 32 == 0) {      // CTA-pair form: CTA r of the pair holds rows [r*BN/2, (r+1)*BN/2) of every B tile
        const int BH = L.BN / 2;
        std::vector<__half> img2(img.size());
        for (int nt = 0; nt < L.n_tiles; ++nt)
            for (int kc = 0; kc < L.k_chunks; ++kc) {
                const __half* src = img.data() + ((size_t)nt * L.k_chunks + kc) * per_chunk;
                for (int r = 0; r < 2; ++r)
                    for (int hl = 0; hl < 2; ++hl)
                        for (int gq = 0; gq < 4; ++gq)
                            for (int n = 0; n < BH; ++n)
                                memcpy(&img2[(((((size_t)nt * L.k_chunks + kc) * 2 + r) * 2 + hl) * 4 + gq) * BH * 8 + (size_t)n * 8],
                                       src + (size_t)hl * L.BN * BK + (size_t)gq * L.BN * 8 + (size_t)(r * BH + n) * 8, 8 * sizeof(__half));
            }
        CB_CUDA(cudaMalloc(&L.img2, img2.size() * sizeof(__half)));
        CB_CUDA(cudaMemcpy(L.img2, img2.data(), img2.size() * sizeof(__half), cudaMemcpyHostToDevice));
        const int rc = make_w_map(L.img2, img2.size(), L.BN, &L.tm_w2);
        if (rc != CB_OK) return rc;
    }
    if ((int)st->layers.size() <= layer_id) st->layers.resize(layer_id + 1, TcLayer{});
    st->layers[layer_id] = L;
    return CB_OK;
}

int cb_tc_prepare(cb_handle* h, const float* hw) {
    TcState* st = new TcState();
    st->d_range_flag = nullptr;
    st->d_lstm_bias = nullptr;
    h->tc = st;
    const CbConfig& c = h->cfg;
    const int C = c.channels, H = c.hidden;
    for (int b = 1; b < c.n_blocks; ++b)
        if (c.stride[b] < 1 || c.k[b] < 1) {       // any width and stride: a tap is a frame shift of the same operand image,
            cb_set_error("bad block geometry");    // a stride multiplies the frame index (conv2b and the 1x1 branch input)
            return CB_ERR_ARG;
        }
    if (C % 32) { cb_set_error("tensor-core path needs channels %% 32 == 0"); return CB_ERR_ARG; }
    auto host = [&](const float* dev) { return hw + (dev - h->d_weights); };
    int rc;
    if (c.stem_k > 0 && c.stride[0] != 1) {
        cb_set_error("tensor-core path: a stem convolution needs a stride-1 first block; use precision fp32");
        return CB_ERR_ARG;
    }
    for (int b = 0; b < c.n_blocks; ++b) {
        const bool rank1 = b == 0 && c.stem_k == 0;         // block 1 of a model without a stem reads the raw signal
        if (!rank1 && (rc = cb_tc_build_layer(h, b * 4 + 0, host(h->conv2a[b].W), C, C, 0)) != CB_OK) return rc;
        if ((rc = cb_tc_build_layer(h, b * 4 + 1, host(h->conv2b[b].W), c.k[b] * C, C, 0)) != CB_OK) return rc;
        if ((rc = cb_tc_build_layer(h, b * 4 + 2, host(h->convc[b].W), rank1 ? C : 2 * C, C, 0)) != CB_OK) return rc;
    }
    // LSTM input projections.  Layer 0 reads the CNN feature image (K = C).  Later layers read the h image written by
    // the recurrence, whose planes are [fw: 13 k-groups (104 ch, 100 real)][bw: 13 k-groups]: K' = 208 with zero rows.
    // Output (gate) columns are permuted to the unit-major order of the recurrence kernel:
    //   TF column gate*H + u  ->  (u/4)*16 + gate*4 + u%4   (per direction)
    const int HP = (H + 7) / 8 * 8;                         // 104
    auto perm = [H](int n) { const int gate = n / H, u = n % H; return (u / 4) * 16 + gate * 4 + (u & 3); };
    // the recurrence evaluates sigmoid/tanh through ex2: -log2(e) is folded into the i/f/o columns and -2*log2(e) into
    // the j columns of W_ih (here), W_hh (cb_lstm_tc.cu) and the bias, so the gate sums feed ex2 directly
    auto gate_scale = [H](int n) { return n / H == 1 ? -2.f * 1.4426950408889634f : -1.4426950408889634f; };
    std::vector<float> bias_all;
    for (int l = 0; l < c.n_layers; ++l) {
        const int n_gemm = (l == 0 || c.rnn_layout == 0) ? 1 : 2;
        for (int d = 0; d < n_gemm; ++d) {
            const int ndir = n_gemm == 1 ? 2 : 1;           // directions covered by this contraction
            const int N = ndir * 4 * H;
            const int Kin = l == 0 ? C : (c.rnn_layout == 0 ? 2 * H : H);
            const int Kimg = l == 0 ? C : (c.rnn_layout == 0 ? 2 * HP : HP);
            const float* src = host(n_gemm == 1 ? h->wxcat[l] : h->wx[l][d]);          // [Kin][N]
            const float* bsrc = host(n_gemm == 1 ? h->bcat[l] : h->bias[l][d]);        // [N]
            std::vector<float> W((size_t)Kimg * N, 0.f), bperm(N);
            for (int k = 0; k < Kin; ++k) {
                const int kk = l == 0 ? k : (k / H) * HP + (k % H);                     // skip the 4 padding channels per direction
                for (int n = 0; n < N; ++n) {
                    const int dd = n / (4 * H), nn = n % (4 * H);
                    W[(size_t)kk * N + dd * 4 * H + perm(nn)] = src[(size_t)k * N + n] * gate_scale(nn);
                }
            }
            for (int n = 0; n < N; ++n) {        // forget_bias 1.0 of TF's LSTMCell folded into the bias
                const int nn = n % (4 * H);
                // (the bias rides through the recurrence's accumulator chain like the projection itself: same inflation)
                bperm[(n / (4 * H)) * 4 * H + perm(nn)] =
                    (float)((double)(bsrc[n] + (nn / H == 2 ? 1.0f : 0.0f)) * gate_scale(nn) * (1.0 + cb_tc_trunc_c() * CB_LSTM_TC_CHAIN));
            }
            if ((rc = cb_tc_build_layer(h, 32 + l * 2 + d, W.data(), Kimg, N, CB_LSTM_TC_CHAIN)) != CB_OK) return rc;
            while (bias_all.size() & 3) bias_all.push_back(0.f);
            st->lstm_bias_off[l][d] = bias_all.size();
            bias_all.insert(bias_all.end(), bperm.begin(), bperm.end());
        }
    }
    CB_CUDA(cudaMalloc(&st->d_lstm_bias, bias_all.size() * sizeof(float)));
    CB_CUDA(cudaMemcpy(st->d_lstm_bias, bias_all.data(), bias_all.size() * sizeof(float), cudaMemcpyHostToDevice));
    st->d_range_flag = h->d_flag + CB_FLAG_TC_RANGE;     // sticky, reported by cb_check_deferred
    CB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    CB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    CB_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<false, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    CB_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    CB_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<false, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    CB_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<true, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    CB_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<false, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    CB_CUDA(cudaFuncSetAttribute(gemm_tc_pair_kernel<true, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SMEM_MAX));
    return cb_lstm_tc_prepare(h, hw);
}

void cb_tc_release(cb_handle* h) {
    cb_lstm_tc_release(h);
    TcState* st = (TcState*)h->tc;
    if (!st) return;
    for (auto& L : st->layers) { if (L.img) cudaFree(L.img); if (L.img2) cudaFree(L.img2); }
    if (st->d_lstm_bias) cudaFree(st->d_lstm_bias);
    delete st;
    h->tc = nullptr;
}

// Mean relative shrink per truncating accumulator add (see cb_tc_build_layer).  CB_TC_BIAS overrides the factor E[1/m]
// (0 switches the compensation off) for calibration runs.
double cb_tc_trunc_c() {
    static const double f = getenv("CB_TC_BIAS") ? atof(getenv("CB_TC_BIAS")) : 0.42;
    return f * 5.9604644775390625e-08;      // * 2^-24
}

int* cb_tc_range_flag(cb_handle* h) { return h->tc ? ((TcState*)h->tc)->d_range_flag : nullptr; }

const float* cb_tc_lstm_bias(cb_handle* h, int layer, int d) {
    TcState* st = (TcState*)h->tc;
    return st->d_lstm_bias + st->lstm_bias_off[layer][d];
}

int cb_launch_gemm_tc(cb_handle* h, const TcGemm& g, cudaStream_t s) {
    TcState* st = (TcState*)h->tc;
    if (!st || g.layer_id < 0 || g.layer_id >= (int)st->layers.size() || !st->layers[g.layer_id].img) {
        cb_set_error("tensor-core path: no weight image for layer %d", g.layer_id);
        return CB_ERR_ARG;
    }
    if (g.T <= 0 || g.B <= 0) return CB_OK;
    const TcLayer& L = st->layers[g.layer_id];
    if (L.N != g.N) { cb_set_error("tensor-core path: layer %d N mismatch", g.layer_id); return CB_ERR_ARG; }
    if (g.taps * g.a0_chunks_per_tap + g.a1_chunks != L.k_chunks) {
        cb_set_error("tensor-core path: layer %d K mismatch (%d chunks vs %d)", g.layer_id,
                     g.taps * g.a0_chunks_per_tap + g.a1_chunks, L.k_chunks);
        return CB_ERR_ARG;
    }
    if (g.Bp % BM) { cb_set_error("tensor-core path: padded batch must be a multiple of 128"); return CB_ERR_ARG; }
    TcParams q;
    int rc = make_img_map(g.a0, &q.tm_a0);
    if (rc == CB_OK) rc = make_img_map(g.a1_chunks ? g.a1 : g.a0, &q.tm_a1);
    if (rc != CB_OK) return rc;
    q.g = g; q.img = L.img; q.BN = L.BN; q.n_tiles = L.n_tiles; q.k_chunks = L.k_chunks;
    q.m_tiles = g.T * (g.Bp / BM); q.out_scale = L.out_scale;
    q.range_flag = st->d_range_flag;
    q.cpp = L.cpp;
    // CTA pairs (tcgen05 cta_group::2: M = 256 over two m-tiles, each CTA stages half of the B tile -- the single-CTA
    // kernel is bound by shared-memory bandwidth: three MMA passes re-read A and B, 96 B/clk + 62 B/clk of TMA writes
    // against the SM's 128 B/clk) whenever the m-tiles pair up; CB_TC_PAIR=0 forces the single-CTA kernel (A/B timing).
    static const int pair_env = getenv("CB_TC_PAIR") ? atoi(getenv("CB_TC_PAIR")) : 1;
    const bool pair = pair_env && L.img2 && q.m_tiles % 2 == 0 && h->sm_count % 2 == 0 && (L.n_tiles == 1 || pair_env == 2);
    const int nc = pair ? 2 : 1;
    static const int sms_env = getenv("CB_TC_GEMM_SMS") ? atoi(getenv("CB_TC_GEMM_SMS")) : 0;   // experiment: cap the grid
    const int sms = sms_env > 0 && sms_env < h->sm_count ? sms_env : h->sm_count - h->reserve_sms;     // (cb_reserve_sms)
    const int units = sms / nc;                           // scheduling units: CTAs or CTA pairs
    const int m_units = q.m_tiles / nc;
    // resident weights when the n-tile's image fits beside the A ring, the units can be dealt evenly over the n-tiles
    // and every unit gets enough m-tiles to amortise the weight load
    const int per_nt = units / L.n_tiles;
    q.resident = L.n_tiles > 1 && smem_bytes_resident(L.BN / nc, L.k_chunks, MIN_STAGES) <= SMEM_MAX && per_nt >= 1 &&
                 m_units >= 8 * per_nt;
    const long long work = (long long)m_units * q.n_tiles;
    int grid = (int)(work < units ? work : units) * nc;
    // as deep a ring as shared memory allows (the two-pass issue order holds cpp stages until their hi*hi products are out)
    static const int stages_env = getenv("CB_TC_STAGES") ? atoi(getenv("CB_TC_STAGES")) : 0;
    auto smem_for = [&](int stages) {
        return q.resident ? smem_bytes_resident(L.BN / nc, L.k_chunks, stages) : smem_bytes_for(L.BN / nc, stages);
    };
    // epilogue vectors of a one-n-tile contraction cached behind the ring and the barriers: shift, and the three rank-1
    // residual vectors when there is a residual (a 7-stage ring, which the plain convolutions then fit, measured neutral: r02_s20)
    //  -- and the bound n-tile's slice of the shift vector in resident mode (single-CTA, single-accumulator kernel)
    static const int force_multi = getenv("CB_TC_FORCE_MULTI") ? atoi(getenv("CB_TC_FORCE_MULTI")) : 0;   // A/B
    const bool multi = q.cpp < q.k_chunks || force_multi;
    static const int proj_vec_env = getenv("CB_TC_PROJ_VEC") ? atoi(getenv("CB_TC_PROJ_VEC")) : 1;        // A/B
    const size_t vec_bytes = L.n_tiles == 1 ? (g.res ? 4 : 1) * (size_t)L.BN * sizeof(float)
                                            : (q.resident && !pair && !multi && !g.res && proj_vec_env ? (size_t)L.BN * sizeof(float) : 0);
    q.stages = MAX_STAGES;
    while (q.stages > MIN_STAGES && smem_for(q.stages) + vec_bytes > SMEM_MAX) --q.stages;
    if (stages_env >= 2 && stages_env <= q.stages) q.stages = stages_env;
    size_t smem = smem_for(q.stages);
    q.vec_off = 0;
    if (vec_bytes && smem + vec_bytes <= SMEM_MAX) {
        q.vec_off = (uint32_t)smem;
        smem += vec_bytes;
    }
    if (smem > SMEM_MAX) { cb_set_error("tensor-core path: layer %d needs %zu bytes of shared memory", g.layer_id, smem); return CB_ERR_ARG; }
    q.dbg = nullptr;
    static const int early_env = getenv("CB_TC_EARLY") ? atoi(getenv("CB_TC_EARLY")) : 1;
    static const int stagger_env = getenv("CB_TC_STAGGER") ? atoi(getenv("CB_TC_STAGGER")) : 4;
    q.early_release = early_env;
    q.stagger = q.k_chunks > q.cpp ? stagger_env : 0;
#if CB_TC_DEV
    static long long* d_dbg = nullptr;
    const int probe_layer = getenv("CB_TC_PROBE") ? atoi(getenv("CB_TC_PROBE")) : -1;
    if (probe_layer == g.layer_id) {
        if (!d_dbg) cudaMalloc(&d_dbg, 32 * 8 * sizeof(long long));
        cudaMemset(d_dbg, 0, 32 * 8 * sizeof(long long));
        q.dbg = d_dbg;
    }
#endif
    if (pair) {
        q.img = L.img2; q.tm_w = L.tm_w2;
        // the convolutions' specialised epilogue (see gemm_tc_body): ReLU'd operand image out of one n-tile, vectors cached
        static const int lean_env = getenv("CB_TC_LEAN_EPI") ? atoi(getenv("CB_TC_LEAN_EPI")) : 1;      // 0: generic (A/B)
        const int epi = (lean_env && q.vec_off && L.n_tiles == 1 && g.relu && g.out_mode == 2) ? (g.res ? 2 : 1) : 0;
        if (multi) {
            if (epi == 2) gemm_tc_pair_kernel<true, 2><<<grid, NTHREADS, smem, s>>>(q);
            else if (epi == 1) gemm_tc_pair_kernel<true, 1><<<grid, NTHREADS, smem, s>>>(q);
            else gemm_tc_pair_kernel<true, 0><<<grid, NTHREADS, smem, s>>>(q);
        } else {
            if (epi == 2) gemm_tc_pair_kernel<false, 2><<<grid, NTHREADS, smem, s>>>(q);
            else if (epi == 1) gemm_tc_pair_kernel<false, 1><<<grid, NTHREADS, smem, s>>>(q);
            else gemm_tc_pair_kernel<false, 0><<<grid, NTHREADS, smem, s>>>(q);
        }
    } else {
        if (multi) gemm_tc_kernel<true><<<grid, NTHREADS1, smem, s>>>(q);
        else gemm_tc_kernel<false><<<grid, NTHREADS1, smem, s>>>(q);
    }
    CB_CHECK_LAUNCH();
    h->launches++;
#if CB_TC_DEV
    if (q.dbg) {          // timeline of partial sums 24..55 of CTA 0: epilogue warp 0 and the MMA warp (clocks from the first stamp)
        long long hb[32 * 8];
        cudaStreamSynchronize(s);
        cudaMemcpy(hb, q.dbg, sizeof(hb), cudaMemcpyDeviceToHost);
        const long long t0 = hb[4];
        fprintf(stderr, "layer %d: k_chunks %d cpp %d stages %d\n", g.layer_id, q.k_chunks, q.cpp, q.stages);
        for (int i = 0; i < 32; ++i)
            fprintf(stderr, "  partial %2d | mma: wait_empty %7lld got %7lld issued %7lld | epi: wait_full %7lld got %7lld drained %7lld released %7lld\n",
                    24 + i, hb[i * 8 + 4] - t0, hb[i * 8 + 5] - t0, hb[i * 8 + 6] - t0, hb[i * 8 + 0] - t0, hb[i * 8 + 1] - t0,
                    hb[i * 8 + 2] - t0, hb[i * 8 + 3] - t0);
    }
#endif
    return CB_OK;
}

int cb_launch_transpose_x(cb_handle* h, const float* x, int B, int L, int Bp, float* xT, cudaStream_t s) {
    transpose_x_kernel<<<dim3((Bp + 31) / 32, (L + 31) / 32), dim3(32, 8), 0, s>>>(x, B, L, Bp, xT);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}

int cb_launch_stem_image(cb_handle* h, const float* xT, int B, int Bp, int L, int t_out, int left, const CbImg& o, cudaStream_t s) {
    if (t_out > 65535) { cb_set_error("segment_len too large for the stem grid"); return CB_ERR_ARG; }
    stem_image_kernel<<<dim3((B + 255) / 256, t_out), 256, 0, s>>>(xT, B, Bp, L, t_out, h->cfg.stem_k, h->cfg.stem_stride, left,
                                                                  h->cfg.channels / 8, h->stem.w, h->stem.inv, h->stem.shift, o);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}

int cb_launch_gen_conv2a(cb_handle* h, const float* xT, int B, int Bp, int L, const CbImg& o, cudaStream_t s) {
    if (L > 65535) { cb_set_error("segment_len too large for the generator grid"); return CB_ERR_ARG; }
    gen_conv2a_kernel<<<dim3((B + 255) / 256, L), 256, 0, s>>>(xT, B, Bp, L, h->cfg.channels / 8, h->g_w, h->g_inv, h->g_sh, o);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}
