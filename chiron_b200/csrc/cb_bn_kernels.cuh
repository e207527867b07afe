// Batch-statistics BatchNorm (CB_BN_BATCH) for the fp32 path: HEAD's simple_global_bn (chiron/cnn.py:166-188) --
// tf.nn.moments(inp, [0, 1, 2]) over every frame of every window of THIS batch (zero padding included), then
// tf.nn.batch_normalization with eps 1e-5 -- which conv_layer applies after every convolution even at inference
// (chiron/cnn.py:65-68).  The shipped checkpoints use population statistics instead (folded into the weights by
// cb_create); this mode exists for models trained at HEAD.
//
// Three kernels per normalised tensor X[M,C] (row-major fp32, raw convolution output):
//   bn_col_stats_kernel   per-CTA partial sum / sum of squares of every channel, accumulated in fp64.  Lanes map to
//                         channel quads (128-bit loads, a warp reads 512 contiguous bytes), a thread walks rows, the row
//                         groups of a CTA are combined through shared memory; no atomics, so the result is deterministic.
//   bn_finalize_kernel    fixed-order sum of the partials, mean / variance in fp64, then inv = scale / sqrt(var + eps),
//                         shift = offset - mean * inv in fp32 (the two vectors tf.nn.batch_normalization multiplies by).
//   bn_apply_kernel       y = act(x * inv + shift [+ second normalised or raw tensor] [+ rank-1 branch of the raw signal]).
// Block 1's two 1x1 convolutions of the one-channel signal are rank-1 (w[c] * x), so their statistics follow from the
// statistics of the samples they read: bn_x_stats_kernel (warp-shuffle + shared-memory reduction of sum / sum of squares)
// and the rank-1 form of bn_finalize_kernel (mean_c = w_c * mean_x, var_c = w_c^2 * var_x).
// Roofline: HBM (4 B/element read by stats, 4+4 B/element by apply); this is the reference-grade path, not the fast one.
// This header holds the kernels only (no launch syntax) so that tests/cuda_emu can compile the same source for the host
// and run it thread by thread on a machine without a GPU.
#pragma once

namespace cb_bn {

constexpr int BN_THREADS = 256;
constexpr int BN_STATS_CTAS = 4;      // resident CTAs per SM of the statistics kernels (<= 64 registers)
constexpr int BN_APPLY_CTAS = 3;      // ... of the apply kernel (<= 85 registers): 3 x 256 threads x 4 x 16 B = 48 KB of loads in flight per SM
constexpr float BN_EPS = 1e-5f;   // chiron/cnn.py:187

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
    return v;
}

// part[blockIdx.x*2 + {0,1}] = sum / sum of squares of x[b*t_in + to*stride] over the CTA's share of (b < B, to < t_out).
__global__ void __launch_bounds__(BN_THREADS) bn_x_stats_kernel(const float* __restrict__ x, int B, int t_in, int stride,
                                                                int t_out, double* __restrict__ part) {
    __shared__ double red[2][BN_THREADS / 32];
    const long long n = (long long)B * t_out;
    double s = 0.0, q = 0.0;
    for (long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * BN_THREADS) {
        const long long b = i / t_out;
        const long long to = i - b * t_out;
        const double v = (double)__ldg(x + b * t_in + to * stride);
        s += v;
        q += v * v;
    }
    s = warp_sum(s);
    q = warp_sum(q);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { red[0][w] = s; red[1][w] = q; }
    __syncthreads();
    if (w == 0) {
        s = l < BN_THREADS / 32 ? red[0][l] : 0.0;
        q = l < BN_THREADS / 32 ? red[1][l] : 0.0;
        s = warp_sum(s);
        q = warp_sum(q);
        if (l == 0) { part[(size_t)blockIdx.x * 2] = s; part[(size_t)blockIdx.x * 2 + 1] = q; }
    }
}

// part[blockIdx.x][0][c] / [1][c] = sum / sum of squares of channel c over the CTA's rows.  C % 4 == 0, C <= 1024.
__global__ void __launch_bounds__(BN_THREADS, BN_STATS_CTAS) bn_col_stats_kernel(const float* __restrict__ X, long long M, int C,
                                                                  double* __restrict__ part) {
    __shared__ double red[2][BN_THREADS * 4];         // [row group][channel]: rpp * C <= 1024 entries
    const int qw = C >> 2;                            // channel quads per row
    const int rpp = BN_THREADS / qw;                  // rows one pass of the CTA covers
    const int q = threadIdx.x % qw, r = threadIdx.x / qw;
    double s[4] = {0.0, 0.0, 0.0, 0.0}, ss[4] = {0.0, 0.0, 0.0, 0.0};
    if (r < rpp) {
        auto add = [&](const float4& v) {
            const double d[4] = {(double)v.x, (double)v.y, (double)v.z, (double)v.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) { s[j] += d[j]; ss[j] += d[j] * d[j]; }
        };
        const long long step = (long long)gridDim.x * rpp;
        const float* src = X + q * 4;
        long long m = (long long)blockIdx.x * rpp + r;
        for (; m + 3 * step < M; m += 4 * step) {       // four independent 128-bit loads in flight per thread
            const float4 v0 = ldg4(src + m * C), v1 = ldg4(src + (m + step) * C), v2 = ldg4(src + (m + 2 * step) * C),
                         v3 = ldg4(src + (m + 3 * step) * C);
            add(v0); add(v1); add(v2); add(v3);
        }
        for (; m < M; m += step) add(ldg4(src + m * C));
#pragma unroll
        for (int j = 0; j < 4; ++j) { red[0][r * C + q * 4 + j] = s[j]; red[1][r * C + q * 4 + j] = ss[j]; }
    }
    __syncthreads();
    for (int c = threadIdx.x; c < C; c += BN_THREADS) {
        double a = 0.0, b = 0.0;
        for (int rr = 0; rr < rpp; ++rr) { a += red[0][rr * C + c]; b += red[1][rr * C + c]; }
        part[(size_t)blockIdx.x * 2 * C + c] = a;
        part[(size_t)blockIdx.x * 2 * C + C + c] = b;
    }
}

// w == nullptr: part is [n_part][2][C] (bn_col_stats_kernel).  w != nullptr: part is [n_part][2] statistics of the raw
// samples and the normalised tensor is the rank-1 product w[c] * x.  scale == nullptr: no BN on this branch (inv 1, shift 0).
// One CTA of 32 warps per 32 channels: lanes = channels (256-byte coalesced loads), warp w sums the partials
// w, w + 32, ... and the 32 warp sums are added in warp order -- a fixed tree, so the moments are deterministic.
constexpr int BN_FIN_THREADS = 1024;
__global__ void __launch_bounds__(BN_FIN_THREADS) bn_finalize_kernel(const double* __restrict__ part, int n_part, int C,
                                                                     double count, const float* __restrict__ w,
                                                                     const float* __restrict__ scale,
                                                                     const float* __restrict__ offset,
                                                                     float* __restrict__ inv, float* __restrict__ shift) {
    __shared__ double red[2][32][33];
    const int lane = threadIdx.x & 31, wp = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane;
    const bool ok = c < C;
    double s = 0.0, q = 0.0;
    if (scale && ok) {
        const size_t step = w ? 2 : (size_t)2 * C, off_s = w ? 0 : (size_t)c, off_q = w ? 1 : (size_t)C + c;
        for (int i = wp; i < n_part; i += 32) { s += part[i * step + off_s]; q += part[i * step + off_q]; }
    }
    red[0][wp][lane] = s;
    red[1][wp][lane] = q;
    __syncthreads();
    if (wp != 0 || !ok) return;
    if (!scale) { inv[c] = 1.0f; shift[c] = 0.0f; return; }
    s = 0.0; q = 0.0;
    for (int i = 0; i < 32; ++i) { s += red[0][i][lane]; q += red[1][i][lane]; }
    double mean = s / count;
    double var = q / count - mean * mean;
    if (var < 0.0) var = 0.0;
    if (w) { const double wc = (double)w[c]; mean *= wc; var *= wc * wc; }
    const float iv = scale[c] * (1.0f / sqrtf((float)var + BN_EPS));
    inv[c] = iv;
    shift[c] = offset[c] - (float)mean * iv;
}

struct BnApply {
    const float *a, *a_inv, *a_sh;          // y = a * a_inv + a_sh
    const float *b, *b_inv, *b_sh;          // optional second tensor; b_inv == nullptr adds it raw
    const float *x, *rw, *rinv, *rsh;       // optional rank-1 branch: (x[win*t_inr + to*strider] * rw) * rinv + rsh
    int t_out, t_inr, strider;
    int relu;
    float* out;
    long long M;
    int C;
};

// Grid-stride over float4 elements in batches of BN_APPLY_U: all loads of a batch are issued before the first store (the
// tensors may alias `out`, which would otherwise serialise every load behind the previous store), and the (row, channel)
// position advances incrementally instead of by a 64-bit division per element.
constexpr int BN_APPLY_U = 4;
__global__ void __launch_bounds__(BN_THREADS, BN_APPLY_CTAS) bn_apply_kernel(const BnApply p) {
    const int qw = p.C >> 2;
    const long long n = p.M * qw;
    const long long S = (long long)gridDim.x * BN_THREADS;
    const long long dm = S / qw;
    const int dq = (int)(S - dm * qw);
    long long i = (long long)blockIdx.x * BN_THREADS + threadIdx.x;
    long long m = i / qw;
    int cq = (int)(i - m * qw);
    while (i < n) {
        float4 va[BN_APPLY_U], vb[BN_APPLY_U];
        long long mm[BN_APPLY_U];
        int cc[BN_APPLY_U];
        int cnt = 0;
#pragma unroll
        for (int u = 0; u < BN_APPLY_U; ++u) {
            if (i < n) {
                mm[u] = m; cc[u] = cq * 4; cnt = u + 1;
                // a and b may alias out: plain loads (ld.global.nc must not see memory this kernel writes)
                va[u] = *reinterpret_cast<const float4*>(p.a + m * p.C + cq * 4);
                if (p.b) vb[u] = *reinterpret_cast<const float4*>(p.b + m * p.C + cq * 4);
                i += S; m += dm; cq += dq;
                if (cq >= qw) { cq -= qw; ++m; }
            }
        }
#pragma unroll
        for (int u = 0; u < BN_APPLY_U; ++u) {
            if (u < cnt) {
                const int c = cc[u];
                const float4 v = va[u], iv = ldg4(p.a_inv + c), sh = ldg4(p.a_sh + c);
                float o[4] = {fmaf(v.x, iv.x, sh.x), fmaf(v.y, iv.y, sh.y), fmaf(v.z, iv.z, sh.z), fmaf(v.w, iv.w, sh.w)};
                if (p.b) {
                    const float4 t = vb[u];
                    if (p.b_inv) {
                        const float4 bi = ldg4(p.b_inv + c), bs = ldg4(p.b_sh + c);
                        o[0] += fmaf(t.x, bi.x, bs.x); o[1] += fmaf(t.y, bi.y, bs.y);
                        o[2] += fmaf(t.z, bi.z, bs.z); o[3] += fmaf(t.w, bi.w, bs.w);
                    } else {
                        o[0] += t.x; o[1] += t.y; o[2] += t.z; o[3] += t.w;
                    }
                }
                if (p.x) {
                    const long long win = mm[u] / p.t_out;
                    const long long to = mm[u] - win * p.t_out;
                    const float xr = __ldg(p.x + win * p.t_inr + to * p.strider);
                    const float4 w = ldg4(p.rw + c), ri = ldg4(p.rinv + c), rs = ldg4(p.rsh + c);
                    o[0] += fmaf(xr * w.x, ri.x, rs.x); o[1] += fmaf(xr * w.y, ri.y, rs.y);
                    o[2] += fmaf(xr * w.z, ri.z, rs.z); o[3] += fmaf(xr * w.w, ri.w, rs.w);
                }
                if (p.relu) {
#pragma unroll
                    for (int j = 0; j < 4; ++j) o[j] = fmaxf(o[j], 0.f);
                }
                *reinterpret_cast<float4*>(p.out + mm[u] * p.C + c) = make_float4(o[0], o[1], o[2], o[3]);
            }
        }
    }
}

// ---- launch plans shared by cb_bn.cu and the host emulation ---------------------------------------------------------------
// Grid of a grid-stride kernel: enough CTAs for the work, at most the CTAs that are resident at once (ctas_per_sm x SM
// count: one wave, no tail) and at most CB_BN_MAX_PART (the partial-sum buffer).
inline int bn_grid(int sm_count, long long work_items, int per_block, int ctas_per_sm) {
    long long g = (work_items + per_block - 1) / per_block;
    const long long cap = (long long)(sm_count > 0 ? sm_count : 148) * ctas_per_sm;
    if (g > cap) g = cap;
    if (g > CB_BN_MAX_PART) g = CB_BN_MAX_PART;
    return g < 1 ? 1 : (int)g;
}

inline BnApply bn_apply_params(const BnApplyArgs& a, int C) {
    BnApply p;
    p.a = a.a; p.a_inv = a.a_inv; p.a_sh = a.a_sh;
    p.b = a.b; p.b_inv = a.b_inv; p.b_sh = a.b_sh;
    p.x = a.x; p.rw = a.rw; p.rinv = a.rinv; p.rsh = a.rsh;
    p.t_out = a.t_out; p.t_inr = a.t_inr; p.strider = a.strider;
    p.relu = a.relu; p.out = a.out; p.M = a.M; p.C = C;
    return p;
}

}  // namespace cb_bn
