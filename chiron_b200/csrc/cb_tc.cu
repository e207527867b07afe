// tcgen05 tensor-core contractions for the conv stack and the hoisted LSTM input projection
// (CB_PREC_TC_SPLIT / CB_PREC_TC_FAST).  Same maths as cb_gemm_simt.cu (chiron/cnn.py:60-82,251-261;
// chiron/rnn.py:49-50,64), different machine:
//
//   * operands are fp16 hi/lo splits (a = hi + lo, |lo| <= 2^-11 |a|): D = Ah*Wh + Ah*Wl + Al*Wh, three
//     tcgen05.mma.kind::f16 per K-step with fp32 accumulation in TMEM.  CB_PREC_TC_FAST issues only Ah*Wh.
//   * activations travel between layers as operand images (cb_tc_common.cuh): the epilogue of the producing kernel
//     writes the hi/lo k-group planes, so the A side of a pipeline stage is 8 cp.async.bulk copies of 2 KB (a conv tap
//     is the same plane shifted by one row; the appended 1x1 branch input is a second image).  Only the first conv2b
//     (block-1 conv2a is a rank-1 function of the raw signal, cnn.py:254) still uses SIMT producer warps, which
//     generate relu((x*w)*inv+shift) on the fly, split it and store the core-matrix image.
//   * persistent CTAs (one per SM), warp-specialised: 4 epilogue warps (TMEM -> scale/shift/residual/ReLU -> fp32 or
//     hi/lo image), 1 MMA-issuing thread, 1 loader thread (weight images + activation images, mbarrier expect-tx),
//     8 producer warps (generator mode only); smem full/empty ring (STAGES deep) and a double-buffered TMEM accumulator
//     so the epilogue of tile i overlaps the MMAs of tile i+1.
#include <math.h>
#include <string.h>

#include <vector>

#include "cb_internal.cuh"
#include "cb_tc_common.cuh"

namespace {

constexpr int BM = 128;          // rows (frames) per tile = UMMA M
constexpr int BK = 32;           // K elements per pipeline stage (2 UMMA K-steps of 16)
constexpr int STAGES = 4;
constexpr int N_PROD_WARPS = 8;  // A-operand producers (generator mode)
constexpr int N_EPI_WARPS = 4;   // one per TMEM lane quadrant
constexpr int NTHREADS = (N_EPI_WARPS + 2 + N_PROD_WARPS) * 32;   // 448

struct TcLayer {                 // one prepared weight image
    __half* img;                 // [n_tiles][k_chunks][2 (hi,lo)][4 k-groups][BN rows][8]
    int K, Kpad, N, BN, n_tiles, k_chunks;
    float out_scale;             // 2^-s, undoes the power-of-two prescale of the weights
};

struct TcState {
    std::vector<TcLayer> layers; // indexed by layer id
    int* d_range_flag;
};

struct TcParams {
    TcGemm g;
    const __half* img;
    int BN, n_tiles, k_chunks, m_tiles;
    float out_scale;
    int passes;                  // 3 = hi/lo split, 1 = fast
    int* range_flag;
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

// fp32 x4 -> fp16 hi x4 (packed in uint2) and fp16 lo x4
__device__ __forceinline__ void split4(const float4& a, uint2& hi, uint2& lo, bool& overflow) {
    const __half2 h01 = __floats2half2_rn(a.x, a.y), h23 = __floats2half2_rn(a.z, a.w);
    const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
    const __half2 l01 = __floats2half2_rn(a.x - f01.x, a.y - f01.y), l23 = __floats2half2_rn(a.z - f23.x, a.w - f23.y);
    hi.x = *reinterpret_cast<const uint32_t*>(&h01); hi.y = *reinterpret_cast<const uint32_t*>(&h23);
    lo.x = *reinterpret_cast<const uint32_t*>(&l01); lo.y = *reinterpret_cast<const uint32_t*>(&l23);
    overflow |= !(fabsf(a.x) <= 65504.f && fabsf(a.y) <= 65504.f && fabsf(a.z) <= 65504.f && fabsf(a.w) <= 65504.f);
}

// Row m of the kernel's row space -> (window b, output frame to); false for padding rows.
__device__ __forceinline__ bool decode_row(const TcGemm& g, long long m, int& b, int& to) {
    if (m >= g.M) return false;
    if (g.row_mode == 0) { b = (int)(m / g.t_out); to = (int)(m - (long long)b * g.t_out); return true; }
    if (g.row_mode == 1) {                       // every window carries one zero row before and after its frames
        const int W = g.t_out + 2;
        b = (int)(m / W); to = (int)(m - (long long)b * W) - 1;
        return to >= 0 && to < g.t_out && b < g.B;
    }
    to = (int)(m / g.Bp); b = (int)(m - (long long)to * g.Bp);     // time-major: m = t*Bp + b
    return b < g.B;
}

// shared memory: STAGES x { A_hi[4][128][8], A_lo, B_hi[4][BN][8], B_lo } halfs, then the barriers.
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(const TcParams q) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const TcGemm& g = q.g;
    const GemmProblem& p = g.p;
    const int BN = q.BN;
    constexpr uint32_t a_bytes = BM * BK * 2;             // one of hi / lo
    const uint32_t b_bytes = (uint32_t)BN * BK * 2;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * stage_bytes);
    uint64_t* full_bar = bars;                            // [STAGES]  operands landed
    uint64_t* empty_bar = bars + STAGES;                  // [STAGES]  MMAs that read the stage retired
    uint64_t* acc_full = bars + 2 * STAGES;               // [2]       accumulator ready for the epilogue
    uint64_t* acc_empty = bars + 2 * STAGES + 2;          // [2]       accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = q.m_tiles * q.n_tiles;

    if (threadIdx.x == 0) {
        const uint32_t full_count = g.a_mode == 1 ? 1 : N_PROD_WARPS * 32 + 1;
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], full_count); mbar_init(&empty_bar[s], 1); }
        for (int b = 0; b < 2; ++b) { mbar_init(&acc_full[b], 1); mbar_init(&acc_empty[b], N_EPI_WARPS * 32); }
        fence_barrier_init();
    }
    if (warp == N_EPI_WARPS) {                            // MMA warp owns the TMEM allocation (all 512 columns)
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp < N_EPI_WARPS) {
        // ============================ epilogue: TMEM -> registers -> scale/shift/residual/ReLU -> HBM =====================
        uint32_t it = 0;
        bool overflow = false;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int mt = tile / q.n_tiles, nt = tile - mt * q.n_tiles;
            const uint32_t buf = it & 1, par = (it >> 1) & 1;
            mbar_wait(&acc_full[buf], par);
            tc_fence_after();
            const long long m = (long long)mt * BM + warp * 32 + lane;
            int b = 0, to = 0;
            const bool row_ok = decode_row(g, m, b, to);
            float xr = 0.f;
            if (p.res && row_ok) xr = __ldg(p.x + (long long)b * p.t_inr + (long long)to * p.strider);
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * (uint32_t)BN;
            const long long orow = g.o_tmajor ? (long long)to * g.Bp + b : m;
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);
                tmem_ld_wait();
                const int n0 = nt * BN + c0;
                if (!row_ok) continue;
                float o[16];
#pragma unroll
                for (int q4 = 0; q4 < 4; ++q4) {
                    const int n = n0 + q4 * 4;
                    const float4 sh = ldg4(p.shift + n);
                    o[q4 * 4 + 0] = fmaf(__uint_as_float(v[q4 * 4 + 0]), q.out_scale, sh.x);
                    o[q4 * 4 + 1] = fmaf(__uint_as_float(v[q4 * 4 + 1]), q.out_scale, sh.y);
                    o[q4 * 4 + 2] = fmaf(__uint_as_float(v[q4 * 4 + 2]), q.out_scale, sh.z);
                    o[q4 * 4 + 3] = fmaf(__uint_as_float(v[q4 * 4 + 3]), q.out_scale, sh.w);
                    if (p.res) {
                        const float4 w = ldg4(p.rw + n), iv = ldg4(p.rinv + n), rs = ldg4(p.rsh + n);
                        o[q4 * 4 + 0] += fmaf(xr * w.x, iv.x, rs.x);
                        o[q4 * 4 + 1] += fmaf(xr * w.y, iv.y, rs.y);
                        o[q4 * 4 + 2] += fmaf(xr * w.z, iv.z, rs.z);
                        o[q4 * 4 + 3] += fmaf(xr * w.w, iv.w, rs.w);
                    }
                }
                if (p.relu) {
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
                        const size_t off = ((size_t)(g.o_plane0 + (n0 >> 3) + h8) * g.o.plane_rows + CB_IMG_GUARD + orow) * 8;
                        *reinterpret_cast<uint4*>(g.o.hi + off) = hi;
                        *reinterpret_cast<uint4*>(g.o.lo + off) = lo;
                    }
                } else if (g.out_mode == 1) {             // fp32, time-major [t][ldo][Bp]: coalesced over the warp's rows
                    float* dst = p.out + ((size_t)to * p.ldo + n0) * (size_t)g.Bp + b;
#pragma unroll
                    for (int e = 0; e < 16; ++e) dst[(size_t)e * g.Bp] = o[e];
                } else {                                  // fp32 row-major
                    float* dst = p.out + m * p.ldo + n0;
#pragma unroll
                    for (int q4 = 0; q4 < 4; ++q4)
                        *reinterpret_cast<float4*>(dst + q4 * 4) = make_float4(o[q4 * 4], o[q4 * 4 + 1], o[q4 * 4 + 2], o[q4 * 4 + 3]);
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
        }
        if (overflow) atomicExch(q.range_flag, 1);
    } else if (warp == N_EPI_WARPS) {
        // ============================ MMA issuer (one elected thread) =======================================================
        if (lane == 0) {
            const uint32_t idesc = make_idesc_f16(BM, BN);
            constexpr uint32_t A_STEP = 2 * BM;               // two k-groups per UMMA K-step, in 16-byte units
            const uint32_t B_STEP = 2 * (uint32_t)BN;
            uint32_t kit = 0, it = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
                const uint32_t buf = it & 1, par = (it >> 1) & 1;
                mbar_wait(&acc_empty[buf], par ^ 1);          // epilogue has drained this accumulator
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + buf * (uint32_t)BN;
                for (int kc = 0; kc < q.k_chunks; ++kc, ++kit) {
                    const uint32_t s = kit % STAGES, ph = (kit / STAGES) & 1;
                    mbar_wait(&full_bar[s], ph);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + (size_t)s * stage_bytes);
                    const uint64_t dah = make_desc(sa, BM * 16, 128), dal = make_desc(sa + a_bytes, BM * 16, 128);
                    const uint64_t dbh = make_desc(sa + 2 * a_bytes, BN * 16, 128);
                    const uint64_t dbl = make_desc(sa + 2 * a_bytes + b_bytes, BN * 16, 128);
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        if (q.passes == 3) {                  // low-order products first (truncating accumulator)
                            umma_f16(d_tmem, dah + ks * A_STEP, dbl + ks * B_STEP, idesc, (kc | ks) != 0);
                            umma_f16(d_tmem, dal + ks * A_STEP, dbh + ks * B_STEP, idesc, 1);
                            umma_f16(d_tmem, dah + ks * A_STEP, dbh + ks * B_STEP, idesc, 1);
                        } else {
                            umma_f16(d_tmem, dah + ks * A_STEP, dbh + ks * B_STEP, idesc, (kc | ks) != 0);
                        }
                    }
                    umma_commit(&empty_bar[s]);               // frees the stage once these MMAs retire
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else if (warp == N_EPI_WARPS + 1) {
        // ============================ loader: weight images (+ activation images) via cp.async.bulk ==========================
        if (lane == 0) {
            uint32_t kit = 0;
            const int n0c = g.taps * g.a0_chunks_per_tap;     // k-chunks served by image 0
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int mt = tile / q.n_tiles, nt = tile - mt * q.n_tiles;
                const __half* wsrc = q.img + (size_t)nt * q.k_chunks * (2 * (size_t)BN * BK);
                const long long r0 = (long long)mt * BM + CB_IMG_GUARD;
                for (int kc = 0; kc < q.k_chunks; ++kc, ++kit) {
                    const uint32_t s = kit % STAGES, ph = (kit / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    uint8_t* st = smem + (size_t)s * stage_bytes;
                    if (g.a_mode == 1) {
                        mbar_arrive_expect_tx(&full_bar[s], 2 * a_bytes + 2 * b_bytes);
                        const __half *hi, *lo; long long row; int plane;
                        if (kc < n0c) {
                            const int j = kc / g.a0_chunks_per_tap, cc = kc - j * g.a0_chunks_per_tap;
                            hi = g.a0.hi; lo = g.a0.lo; row = r0 + j - g.left; plane = g.a0_plane0 + cc * 4;
                            hi += ((size_t)plane * g.a0.plane_rows + row) * 8; lo += ((size_t)plane * g.a0.plane_rows + row) * 8;
#pragma unroll
                            for (int kg = 0; kg < 4; ++kg) {
                                bulk_g2s(st + kg * (BM * 16), hi + (size_t)kg * g.a0.plane_rows * 8, BM * 16, &full_bar[s]);
                                bulk_g2s(st + a_bytes + kg * (BM * 16), lo + (size_t)kg * g.a0.plane_rows * 8, BM * 16, &full_bar[s]);
                            }
                        } else {
                            const int cc = kc - n0c;
                            plane = g.a1_plane0 + cc * 4;
                            hi = g.a1.hi + ((size_t)plane * g.a1.plane_rows + r0) * 8;
                            lo = g.a1.lo + ((size_t)plane * g.a1.plane_rows + r0) * 8;
#pragma unroll
                            for (int kg = 0; kg < 4; ++kg) {
                                bulk_g2s(st + kg * (BM * 16), hi + (size_t)kg * g.a1.plane_rows * 8, BM * 16, &full_bar[s]);
                                bulk_g2s(st + a_bytes + kg * (BM * 16), lo + (size_t)kg * g.a1.plane_rows * 8, BM * 16, &full_bar[s]);
                            }
                        }
                    } else {
                        mbar_arrive_expect_tx(&full_bar[s], 2 * b_bytes);
                    }
                    bulk_g2s(st + 2 * a_bytes, wsrc + (size_t)kc * (2 * (size_t)BN * BK), 2 * b_bytes, &full_bar[s]);
                }
            }
        }
    } else if (g.a_mode == 0) {
        // ============================ A producers (generator / gather mode) ==================================================
        const int pt = threadIdx.x - (N_EPI_WARPS + 2) * 32;      // 0..255
        const int r = pt & 127, hsel = pt >> 7;                   // row of the tile, which 16-wide half of the 32-wide chunk
        const int K0 = p.taps * p.c0;
        bool overflow = false;
        uint32_t kit = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / q.n_tiles;
            const long long m = (long long)mt * BM + r;
            int b = 0, to = 0;
            const bool row_ok = decode_row(g, m, b, to);
            const long long f0 = (long long)b * p.t_in0;
            const int tbase = to * p.stride0 - p.left;
            for (int kc = 0; kc < q.k_chunks; ++kc, ++kit) {
                const uint32_t s = kit % STAGES, ph = (kit / STAGES) & 1;
                float4 a[4];
                const int kk = kc * BK + hsel * 16;
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) {
                    const int kq = kk + gq * 4;
                    a[gq] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!row_ok || kq >= p.K) continue;
                    if (kq < K0) {
                        const int j = kq / p.c0, c = kq - j * p.c0;
                        const int ti = tbase + j;
                        if (ti < 0 || ti >= p.t_in0) continue;
                        if (p.gen) {
                            const float xv = __ldg(p.x + f0 + ti);
                            const float4 w = ldg4(p.gw + c), iv = ldg4(p.ginv + c), sh = ldg4(p.gsh + c);
                            a[gq].x = fmaxf(fmaf(xv * w.x, iv.x, sh.x), 0.f);
                            a[gq].y = fmaxf(fmaf(xv * w.y, iv.y, sh.y), 0.f);
                            a[gq].z = fmaxf(fmaf(xv * w.z, iv.z, sh.z), 0.f);
                            a[gq].w = fmaxf(fmaf(xv * w.w, iv.w, sh.w), 0.f);
                        } else {
                            a[gq] = ldg4(p.src0 + (f0 + ti) * p.lda0 + c);
                        }
                    } else {
                        a[gq] = ldg4(p.src1 + ((long long)b * p.t_in1 + (long long)to * p.stride1) * p.lda1 + (kq - K0));
                    }
                }
                uint2 hi[4], lo[4];
#pragma unroll
                for (int gq = 0; gq < 4; ++gq) split4(a[gq], hi[gq], lo[gq], overflow);
                mbar_wait(&empty_bar[s], ph ^ 1);
                uint8_t* st = smem + (size_t)s * stage_bytes;
                // k-group (8 halfs = 16 B) index within the stage: hsel*2 + {0,1}; row r at +r*16
                uint4* ah = reinterpret_cast<uint4*>(st + (size_t)(hsel * 2) * (BM * 16) + r * 16);
                uint4* al = reinterpret_cast<uint4*>(st + a_bytes + (size_t)(hsel * 2) * (BM * 16) + r * 16);
                ah[0] = make_uint4(hi[0].x, hi[0].y, hi[1].x, hi[1].y);
                ah[BM] = make_uint4(hi[2].x, hi[2].y, hi[3].x, hi[3].y);          // next k-group: + BM*16 bytes
                al[0] = make_uint4(lo[0].x, lo[0].y, lo[1].x, lo[1].y);
                al[BM] = make_uint4(lo[2].x, lo[2].y, lo[3].x, lo[3].y);
                fence_proxy_async();                           // generic-proxy stores -> visible to the tensor core
                mbar_arrive(&full_bar[s]);
            }
        }
        if (overflow) atomicExch(q.range_flag, 1);
    }

    tc_fence_before();
    __syncthreads();
    if (warp == N_EPI_WARPS) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

size_t smem_bytes_for(int BN) { return (size_t)STAGES * (2 * BM * BK * 2 + 2 * (size_t)BN * BK * 2) + 256; }

int pick_bn(int N) {
    if (N % 256 == 0) return 256;
    for (int bn = 256; bn >= 16; bn -= 16)
        if (N % bn == 0) return bn;
    return 0;
}

}  // namespace

// ---- host: weight images -------------------------------------------------------------------------------------------------
// W is [K][N] fp32 (row k = input channel in the order the A operand presents it).
int cb_tc_build_layer(cb_handle* h, int layer_id, const float* W, int K, int N) {
    TcState* st = (TcState*)h->tc;
    TcLayer L;
    memset(&L, 0, sizeof(L));
    L.K = K; L.N = N; L.BN = pick_bn(N);
    if (L.BN == 0) { cb_set_error("tensor-core path: N=%d is not a multiple of 16", N); return CB_ERR_ARG; }
    L.n_tiles = N / L.BN;
    L.k_chunks = (K + BK - 1) / BK;
    L.Kpad = L.k_chunks * BK;
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
                        const float w = k < K ? W[(size_t)k * N + nt * L.BN + n] * scale : 0.f;
                        const __half hi = __float2half_rn(w);
                        const __half lo = __float2half_rn(w - __half2float(hi));
                        base[(size_t)gq * L.BN * 8 + n * 8 + e] = hi;
                        base[(size_t)L.BN * BK + (size_t)gq * L.BN * 8 + n * 8 + e] = lo;
                    }
        }
    CB_CUDA(cudaMalloc(&L.img, img.size() * sizeof(__half)));
    CB_CUDA(cudaMemcpy(L.img, img.data(), img.size() * sizeof(__half), cudaMemcpyHostToDevice));
    if ((int)st->layers.size() <= layer_id) st->layers.resize(layer_id + 1, TcLayer{});
    st->layers[layer_id] = L;
    return CB_OK;
}

int cb_tc_prepare(cb_handle* h, const float* hw) {
    TcState* st = new TcState();
    st->d_range_flag = nullptr;
    h->tc = st;
    const CbConfig& c = h->cfg;
    const int C = c.channels, H = c.hidden;
    for (int b = 1; b < c.n_blocks; ++b)
        if (c.stride[b] != 1 || c.k[b] != 3) {
            cb_set_error("tensor-core path supports stride-1, width-3 residual blocks after the first; use precision fp32");
            return CB_ERR_ARG;
        }
    if (C % 32) { cb_set_error("tensor-core path needs channels %% 32 == 0"); return CB_ERR_ARG; }
    auto host = [&](const float* dev) { return hw + (dev - h->d_weights); };
    int rc;
    for (int b = 0; b < c.n_blocks; ++b) {
        if (b > 0 && (rc = cb_tc_build_layer(h, b * 4 + 0, host(h->conv2a[b].W), C, C)) != CB_OK) return rc;
        if ((rc = cb_tc_build_layer(h, b * 4 + 1, host(h->conv2b[b].W), c.k[b] * C, C)) != CB_OK) return rc;
        if ((rc = cb_tc_build_layer(h, b * 4 + 2, host(h->convc[b].W), b == 0 ? C : 2 * C, C)) != CB_OK) return rc;
    }
    // LSTM input projections.  Layer 0 reads the CNN feature image (K = C).  Later layers read the h image written by
    // the recurrence, whose planes are [fw: 13 k-groups (104 ch, 100 real)][bw: 13 k-groups]: K' = 208 with zero rows.
    const int HP = (H + 7) / 8 * 8;                         // 104
    for (int l = 0; l < c.n_layers; ++l) {
        if (l == 0) {
            if ((rc = cb_tc_build_layer(h, 32, host(h->wxcat[0]), C, 8 * H)) != CB_OK) return rc;
        } else if (c.rnn_layout == 0) {
            std::vector<float> W((size_t)2 * HP * 8 * H, 0.f);
            const float* src = host(h->wxcat[l]);               // [2H][8H]
            for (int d = 0; d < 2; ++d)
                for (int u = 0; u < H; ++u)
                    memcpy(&W[(size_t)(d * HP + u) * 8 * H], src + (size_t)(d * H + u) * 8 * H, sizeof(float) * 8 * H);
            if ((rc = cb_tc_build_layer(h, 32 + l * 2, W.data(), 2 * HP, 8 * H)) != CB_OK) return rc;
        } else {
            for (int d = 0; d < 2; ++d) {
                std::vector<float> W((size_t)HP * 4 * H, 0.f);
                memcpy(W.data(), host(h->wx[l][d]), sizeof(float) * (size_t)H * 4 * H);
                if ((rc = cb_tc_build_layer(h, 32 + l * 2 + d, W.data(), HP, 4 * H)) != CB_OK) return rc;
            }
        }
    }
    CB_CUDA(cudaMalloc(&st->d_range_flag, sizeof(int)));
    CB_CUDA(cudaMemset(st->d_range_flag, 0, sizeof(int)));
    CB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_for(256)));
    return cb_lstm_tc_prepare(h, hw);
}

void cb_tc_release(cb_handle* h) {
    cb_lstm_tc_release(h);
    TcState* st = (TcState*)h->tc;
    if (!st) return;
    for (auto& L : st->layers) if (L.img) cudaFree(L.img);
    if (st->d_range_flag) cudaFree(st->d_range_flag);
    delete st;
    h->tc = nullptr;
}

int* cb_tc_range_flag(cb_handle* h) { return h->tc ? ((TcState*)h->tc)->d_range_flag : nullptr; }

int cb_tc_check_range(cb_handle* h, cudaStream_t s) {
    TcState* st = (TcState*)h->tc;
    if (!st) return CB_OK;
    int flag = 0;
    CB_CUDA(cudaMemcpyAsync(&flag, st->d_range_flag, sizeof(int), cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    if (flag) {
        cudaMemsetAsync(st->d_range_flag, 0, sizeof(int), s);
        cb_set_error("an activation exceeded the fp16 range of the tensor-core path; rerun with precision fp32");
        return CB_ERR_RANGE;
    }
    return CB_OK;
}

int cb_launch_gemm_tc(cb_handle* h, const TcGemm& g, cudaStream_t s) {
    TcState* st = (TcState*)h->tc;
    if (!st || g.layer_id < 0 || g.layer_id >= (int)st->layers.size() || !st->layers[g.layer_id].img) {
        cb_set_error("tensor-core path: no weight image for layer %d", g.layer_id);
        return CB_ERR_ARG;
    }
    if (g.M <= 0) return CB_OK;
    const TcLayer& L = st->layers[g.layer_id];
    if (L.N != g.p.N) { cb_set_error("tensor-core path: layer %d N mismatch", g.layer_id); return CB_ERR_ARG; }
    if (g.a_mode == 1) {
        if (g.taps * g.a0_chunks_per_tap + g.a1_chunks != L.k_chunks) {
            cb_set_error("tensor-core path: layer %d K mismatch (%d chunks vs %d)", g.layer_id,
                         g.taps * g.a0_chunks_per_tap + g.a1_chunks, L.k_chunks);
            return CB_ERR_ARG;
        }
    } else if (L.K != g.p.K || (g.p.c0 & 3) || (g.p.c1 & 3) || (g.p.lda0 & 3) || (g.p.lda1 & 3)) {
        cb_set_error("tensor-core path: layer %d gather shape mismatch", g.layer_id);
        return CB_ERR_ARG;
    }
    if ((g.row_mode == 2 || g.o_tmajor) && g.Bp % BM) { cb_set_error("tensor-core path: padded batch must be a multiple of 128"); return CB_ERR_ARG; }
    TcParams q;
    q.g = g; q.img = L.img; q.BN = L.BN; q.n_tiles = L.n_tiles; q.k_chunks = L.k_chunks;
    q.m_tiles = (int)(((long long)g.M + BM - 1) / BM); q.out_scale = L.out_scale;
    q.passes = h->precision == CB_PREC_TC_FAST ? 1 : 3;
    q.range_flag = st->d_range_flag;
    const long long tiles = (long long)q.m_tiles * q.n_tiles;
    const int grid = (int)(tiles < h->sm_count ? tiles : h->sm_count);
    gemm_tc_kernel<<<grid, NTHREADS, smem_bytes_for(L.BN), s>>>(q);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}
