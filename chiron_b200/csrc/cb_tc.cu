// tcgen05 tensor-core contractions for the conv stack and the hoisted LSTM input projection
// (CB_PREC_TC_SPLIT / CB_PREC_TC_FAST).  Same GemmProblem contract as cb_gemm_simt.cu (chiron/cnn.py:60-82,251-261;
// chiron/rnn.py:49-50,64), different machine:
//
//   * operands are fp16 hi/lo splits (a = hi + lo, |lo| <= 2^-11 |a|): D = Ah*Wh + Ah*Wl + Al*Wh, three
//     tcgen05.mma.kind::f16 per K-step with fp32 accumulation in TMEM -> ~2^-21 relative error per product, i.e.
//     fp32-class, which is what bit-exact greedy bases need.  CB_PREC_TC_FAST issues only Ah*Wh.
//   * warp-specialised persistent CTAs (one per SM): 8 producer warps gather the fp32 activation rows (im2col taps,
//     'SAME' padding, strides, the appended 1x1 branch input, or the rank-1 block-1 generator), split them and store the
//     K-major no-swizzle core-matrix image; one thread streams the pre-packed weight images with cp.async.bulk;
//     one thread issues the MMAs; 4 epilogue warps drain TMEM (tcgen05.ld 32x32b) -> scale/shift/residual/ReLU -> HBM.
//   * mbarrier pipelines: smem full/empty ring (STAGES deep) and a double-buffered TMEM accumulator full/empty pair, so
//     the epilogue of tile i overlaps the MMAs of tile i+1.
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include <vector>

#include "cb_internal.cuh"

namespace {

constexpr int BM = 128;          // rows (frames) per tile = UMMA M
constexpr int BK = 32;           // K elements per pipeline stage (2 UMMA K-steps of 16)
constexpr int STAGES = 4;
constexpr int N_PROD_WARPS = 8;  // A-operand producers
constexpr int N_EPI_WARPS = 4;   // one per TMEM lane quadrant
constexpr int NTHREADS = (N_EPI_WARPS + 2 + N_PROD_WARPS) * 32;   // 448

struct TcLayer {                 // one prepared weight image
    __half* img;                 // [n_tiles][k_chunks][2 (hi,lo)][4 k-groups][BN rows][8]
    int K, Kpad, N, BN, n_tiles, k_chunks;
    float out_scale;             // 2^-s, undoes the power-of-two prescale of the weights
};

struct TcState {
    std::vector<TcLayer> layers; // indexed by GemmProblem::layer_id
    int* d_range_flag;
};

struct TcParams {
    GemmProblem p;
    const __half* img;
    int BN, n_tiles, k_chunks, m_tiles;
    float out_scale;
    int passes;                  // 3 = hi/lo split, 1 = fast
    int* range_flag;
};

// ---- PTX helpers -----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor: 16-byte k-group g of row r lives at g*lbo + (r/8)*sbo + (r%8)*16.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFF);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
    d |= 1ULL << 46;             // descriptor version (Blackwell)
    return d;                    // base_offset 0, lbo_mode 0, layout SWIZZLE_NONE
}

__device__ __forceinline__ void umma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t acc) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(acc)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&v)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]),
          "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]),
          "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr));
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

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

// ---- the kernel ----------------------------------------------------------------------------------------------------------
// shared memory: STAGES x { A_hi[4][128][8], A_lo, B_hi[4][BN][8], B_lo } halfs, then the barriers.
__global__ void __launch_bounds__(NTHREADS, 1) gemm_tc_kernel(const TcParams q) {
    extern __shared__ __align__(1024) uint8_t smem[];
    const GemmProblem& p = q.p;
    const int BN = q.BN;
    const uint32_t a_bytes = BM * BK * 2;                 // one of hi / lo
    const uint32_t b_bytes = (uint32_t)BN * BK * 2;
    const uint32_t stage_bytes = 2 * a_bytes + 2 * b_bytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + (size_t)STAGES * stage_bytes);
    uint64_t* full_bar = bars;                            // [STAGES]  producers + weight bytes landed
    uint64_t* empty_bar = bars + STAGES;                  // [STAGES]  MMAs that read the stage retired
    uint64_t* acc_full = bars + 2 * STAGES;               // [2]       accumulator ready for the epilogue
    uint64_t* acc_empty = bars + 2 * STAGES + 2;          // [2]       accumulator drained
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = q.m_tiles * q.n_tiles;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], N_PROD_WARPS * 32 + 1); mbar_init(&empty_bar[s], 1); }
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
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++it) {
            const int mt = tile / q.n_tiles, nt = tile - mt * q.n_tiles;
            const uint32_t buf = it & 1, par = (it >> 1) & 1;
            mbar_wait(&acc_full[buf], par);
            tc_fence_after();
            const long long m = (long long)mt * BM + warp * 32 + lane;
            bool row_ok = m < p.M;
            long long tm_t = 0; int tm_b = 0;
            if (p.tmajor) { tm_t = m / p.Bp; tm_b = (int)(m - tm_t * p.Bp); row_ok = row_ok && tm_b < p.Bvalid; }
            float xr = 0.f;
            if (p.res && row_ok) {
                const int b = (int)(m / p.t_out), to = (int)(m % p.t_out);
                xr = __ldg(p.x + (long long)b * p.t_inr + (long long)to * p.strider);
            }
            const uint32_t taddr = tmem_base + ((uint32_t)(warp * 32) << 16) + buf * (uint32_t)BN;
            for (int c0 = 0; c0 < BN; c0 += 16) {
                uint32_t v[16];
                tmem_ld16(taddr + c0, v);
                const int n0 = nt * BN + c0;
                if (row_ok && n0 < p.N && p.out_tlayout) {
                    float* dst = p.out + ((size_t)tm_t * p.ldo + n0) * (size_t)p.Bp + tm_b;
#pragma unroll
                    for (int e = 0; e < 16; ++e)
                        dst[(size_t)e * p.Bp] = fmaf(__uint_as_float(v[e]), q.out_scale, __ldg(p.shift + n0 + e));
                } else if (row_ok && n0 < p.N) {
                    float* dst = p.out + m * p.ldo + n0;
#pragma unroll
                    for (int g = 0; g < 4; ++g) {
                        const int n = n0 + g * 4;
                        const float4 sh = ldg4(p.shift + n);
                        float o[4] = {fmaf(__uint_as_float(v[g * 4 + 0]), q.out_scale, sh.x),
                                      fmaf(__uint_as_float(v[g * 4 + 1]), q.out_scale, sh.y),
                                      fmaf(__uint_as_float(v[g * 4 + 2]), q.out_scale, sh.z),
                                      fmaf(__uint_as_float(v[g * 4 + 3]), q.out_scale, sh.w)};
                        if (p.res) {
                            const float4 w = ldg4(p.rw + n), iv = ldg4(p.rinv + n), rs = ldg4(p.rsh + n);
                            o[0] += fmaf(xr * w.x, iv.x, rs.x);
                            o[1] += fmaf(xr * w.y, iv.y, rs.y);
                            o[2] += fmaf(xr * w.z, iv.z, rs.z);
                            o[3] += fmaf(xr * w.w, iv.w, rs.w);
                        }
                        if (p.relu) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) o[e] = fmaxf(o[e], 0.f);
                        }
                        *reinterpret_cast<float4*>(dst + g * 4) = make_float4(o[0], o[1], o[2], o[3]);
                    }
                }
            }
            tc_fence_before();
            mbar_arrive(&acc_empty[buf]);
        }
    } else if (warp == N_EPI_WARPS) {
        // ============================ MMA issuer (one elected thread) =======================================================
        if (lane == 0) {
            const uint32_t idesc = (1u << 4) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);   // f16 x f16 -> f32
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
                    const uint32_t a_hi = sa, a_lo = sa + a_bytes, b_hi = sa + 2 * a_bytes, b_lo = b_hi + b_bytes;
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        const uint32_t ao = ks * 2 * (BM * 16), bo = ks * 2 * ((uint32_t)BN * 16);
                        const uint64_t dah = make_desc(a_hi + ao, BM * 16, 128), dal = make_desc(a_lo + ao, BM * 16, 128);
                        const uint64_t dbh = make_desc(b_hi + bo, BN * 16, 128), dbl = make_desc(b_lo + bo, BN * 16, 128);
                        umma_f16(d_tmem, dah, dbh, idesc, (kc | ks) != 0);
                        if (q.passes == 3) {
                            umma_f16(d_tmem, dah, dbl, idesc, 1);
                            umma_f16(d_tmem, dal, dbh, idesc, 1);
                        }
                    }
                    umma_commit(&empty_bar[s]);               // frees the stage once these MMAs retire
                }
                umma_commit(&acc_full[buf]);
            }
        }
    } else if (warp == N_EPI_WARPS + 1) {
        // ============================ weight loader: pre-packed images, one bulk copy per stage ==============================
        if (lane == 0) {
            uint32_t kit = 0;
            for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
                const int nt = tile % q.n_tiles;
                const __half* src = q.img + (size_t)nt * q.k_chunks * (2 * (size_t)BN * BK);
                for (int kc = 0; kc < q.k_chunks; ++kc, ++kit) {
                    const uint32_t s = kit % STAGES, ph = (kit / STAGES) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_arrive_expect_tx(&full_bar[s], 2 * b_bytes);
                    bulk_g2s(smem + (size_t)s * stage_bytes + 2 * a_bytes, src + (size_t)kc * (2 * (size_t)BN * BK),
                             2 * b_bytes, &full_bar[s]);
                }
            }
        }
    } else {
        // ============================ A producers: gather fp32 rows, split to fp16 hi/lo, store the core-matrix image ========
        const int pt = threadIdx.x - (N_EPI_WARPS + 2) * 32;      // 0..255
        const int r = pt & 127, hsel = pt >> 7;                   // row of the tile, which 16-wide half of the 32-wide chunk
        const int K0 = p.taps * p.c0;
        bool overflow = false;
        uint32_t kit = 0;
        for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
            const int mt = tile / q.n_tiles;
            const long long m = (long long)mt * BM + r;
            bool row_ok = m < p.M;
            int b = 0, to = 0;
            if (row_ok) {
                if (p.tmajor) { to = (int)(m / p.Bp); b = (int)(m - (long long)to * p.Bp); row_ok = b < p.Bvalid; }
                else { b = (int)(m / p.t_out); to = (int)(m % p.t_out); }
            }
            const long long f0 = (long long)b * p.t_in0;
            const int tbase = to * p.stride0 - p.left;
            for (int kc = 0; kc < q.k_chunks; ++kc, ++kit) {
                const uint32_t s = kit % STAGES, ph = (kit / STAGES) & 1;
                float4 a[4];
                const int kk = kc * BK + hsel * 16;
#pragma unroll
                for (int g = 0; g < 4; ++g) {
                    const int kq = kk + g * 4;
                    a[g] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!row_ok || kq >= p.K) continue;
                    if (kq < K0) {
                        const int j = kq / p.c0, c = kq - j * p.c0;
                        const int ti = tbase + j;
                        if (ti < 0 || ti >= p.t_in0) continue;
                        if (p.gen) {
                            const float xv = __ldg(p.x + f0 + ti);
                            const float4 w = ldg4(p.gw + c), iv = ldg4(p.ginv + c), sh = ldg4(p.gsh + c);
                            a[g].x = fmaxf(fmaf(xv * w.x, iv.x, sh.x), 0.f);
                            a[g].y = fmaxf(fmaf(xv * w.y, iv.y, sh.y), 0.f);
                            a[g].z = fmaxf(fmaf(xv * w.z, iv.z, sh.z), 0.f);
                            a[g].w = fmaxf(fmaf(xv * w.w, iv.w, sh.w), 0.f);
                        } else if (p.a_tlayout) {
                            const float* src = p.src0 + ((size_t)ti * p.lda0 + c) * (size_t)p.Bp + b;
                            a[g].x = __ldg(src); a[g].y = __ldg(src + p.Bp);
                            a[g].z = __ldg(src + 2 * (size_t)p.Bp); a[g].w = __ldg(src + 3 * (size_t)p.Bp);
                        } else {
                            a[g] = ldg4(p.src0 + (f0 + ti) * p.lda0 + c);
                        }
                    } else {
                        a[g] = ldg4(p.src1 + ((long long)b * p.t_in1 + (long long)to * p.stride1) * p.lda1 + (kq - K0));
                    }
                }
                uint2 hi[4], lo[4];
#pragma unroll
                for (int g = 0; g < 4; ++g) split4(a[g], hi[g], lo[g], overflow);
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
static int build_layer(TcState* st, int layer_id, const float* W, int K, int N) {
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
            for (int g = 0; g < 4; ++g)
                for (int n = 0; n < L.BN; ++n)
                    for (int e = 0; e < 8; ++e) {
                        const int k = kc * BK + g * 8 + e;
                        const float w = k < K ? W[(size_t)k * N + nt * L.BN + n] * scale : 0.f;
                        const __half hi = __float2half_rn(w);
                        const __half lo = __float2half_rn(w - __half2float(hi));
                        base[(size_t)g * L.BN * 8 + n * 8 + e] = hi;
                        base[(size_t)L.BN * BK + (size_t)g * L.BN * 8 + n * 8 + e] = lo;
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
    auto host = [&](const float* dev) { return hw + (dev - h->d_weights); };
    int rc;
    for (int b = 0; b < c.n_blocks; ++b) {
        if (b > 0 && (rc = build_layer(st, b * 4 + 0, host(h->conv2a[b].W), C, C)) != CB_OK) return rc;
        if ((rc = build_layer(st, b * 4 + 1, host(h->conv2b[b].W), c.k[b] * C, C)) != CB_OK) return rc;
        if ((rc = build_layer(st, b * 4 + 2, host(h->convc[b].W), b == 0 ? C : 2 * C, C)) != CB_OK) return rc;
    }
    for (int l = 0; l < c.n_layers; ++l) {
        if (l == 0 || c.rnn_layout == 0) {
            if ((rc = build_layer(st, 32 + l * 2, host(h->wxcat[l]), l == 0 ? C : 2 * H, 8 * H)) != CB_OK) return rc;
        } else {
            for (int d = 0; d < 2; ++d)
                if ((rc = build_layer(st, 32 + l * 2 + d, host(h->wx[l][d]), H, 4 * H)) != CB_OK) return rc;
        }
    }
    CB_CUDA(cudaMalloc(&st->d_range_flag, sizeof(int)));
    CB_CUDA(cudaMemset(st->d_range_flag, 0, sizeof(int)));
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

int cb_launch_gemm_tc(cb_handle* h, const GemmProblem& p, cudaStream_t s) {
    TcState* st = (TcState*)h->tc;
    if (!st || p.layer_id < 0 || p.layer_id >= (int)st->layers.size() || !st->layers[p.layer_id].img) {
        cb_set_error("tensor-core path: no weight image for layer %d", p.layer_id);
        return CB_ERR_ARG;
    }
    if (p.M <= 0) return CB_OK;
    const TcLayer& L = st->layers[p.layer_id];
    if (L.K != p.K || L.N != p.N) { cb_set_error("tensor-core path: layer %d shape mismatch", p.layer_id); return CB_ERR_ARG; }
    if ((p.c0 & 3) || (p.c1 & 3) || (p.lda0 & 3) || (p.lda1 & 3) || (p.ldo & 3) || (p.taps > 1 && (p.c0 % 16))) {
        cb_set_error("tensor-core path: unsupported channel alignment");
        return CB_ERR_ARG;
    }
    TcParams q;
    q.p = p; q.img = L.img; q.BN = L.BN; q.n_tiles = L.n_tiles; q.k_chunks = L.k_chunks;
    q.m_tiles = (p.M + BM - 1) / BM; q.out_scale = L.out_scale;
    if (p.tmajor && (p.Bp % BM || p.taps != 1 || p.stride0 != 1 || p.c1 || p.res || p.relu)) {
        cb_set_error("tensor-core path: unsupported time-major contraction");
        return CB_ERR_ARG;
    }
    q.passes = h->precision == CB_PREC_TC_FAST ? 1 : 3;
    q.range_flag = st->d_range_flag;
    const size_t smem = smem_bytes_for(L.BN);
    static bool attr_set = false;
    if (!attr_set) {
        CB_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_bytes_for(256)));
        attr_set = true;
    }
    const long long tiles = (long long)q.m_tiles * q.n_tiles;
    const int grid = (int)(tiles < h->sm_count ? tiles : h->sm_count);
    gemm_tc_kernel<<<grid, NTHREADS, smem, s>>>(q);
    CB_CHECK_LAUNCH();
    h->launches++;
    return CB_OK;
}

