// fp32 FFMA implicit-GEMM for the residual conv stack and the hoisted LSTM input projection (CB_PREC_FP32).
//
// Replaces tf.nn.conv2d + tf.nn.batch_normalization + relu of conv_layer (chiron/cnn.py:60-82) for conv2a/conv2b/conv2c
// and branch1 (cnn.py:251-261), and the input half of LSTMCell's matmul (chiron/rnn.py:49-50,64).  The im2col gather
// ('SAME' zero padding, stride, the 1x1 branch input appended along K) happens in the A-tile loader; the epilogue adds
// the folded BN shift, the rank-1 block-1 branch1 term and the ReLU.  Roofline: FFMA-bound (this is the slow,
// reference-grade path; the tcgen05 path in cb_gemm_tc.cu is the fast one).
// This header holds the kernel only (no launch syntax): tests/cuda_emu compiles the same source for the host.
#pragma once

namespace cb_simt {


constexpr int BM = 128, BN = 128, BK = 8, NT = 256;

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }

__global__ void __launch_bounds__(NT, 2) gemm_simt_kernel(const GemmProblem p) {
    __shared__ __align__(16) float As[2][BK][BM + 4];
    __shared__ __align__(16) float Bs[2][BK][BN];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int n_tiles = (p.N + BN - 1) / BN;       // 1-D grid, N-tiles of one row block adjacent (A tile reuse in L2)
    const int n0 = (int)(blockIdx.x % n_tiles) * BN;
    const long long m0 = (long long)(blockIdx.x / n_tiles) * BM;
    const int K0 = p.taps * p.c0;

    // A-loader coordinates: one float4 (4 consecutive k) of one row per thread per k-chunk
    const int ar = tid >> 1, akq = (tid & 1) * 4;
    const long long am = m0 + ar;
    const bool arow_ok = am < p.M;
    int ab = 0, ato = 0;
    if (arow_ok) { ab = (int)(am / p.t_out); ato = (int)(am % p.t_out); }
    const long long aframe0 = (long long)ab * p.t_in0;         // first part-0 input frame of this window
    const int abase_t = ato * p.stride0 - p.left;
    // B-loader coordinates
    const int br = tid >> 5, bc = (tid & 31) * 4;

    auto load_a = [&](int k0) -> float4 {
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const int kk = k0 + akq;
        if (!arow_ok || kk >= p.K) return v;
        if (kk < K0) {
            const int j = kk / p.c0, c = kk - j * p.c0;
            const int ti = abase_t + j;
            if (ti < 0 || ti >= p.t_in0) return v;
            if (p.gen) {
                const float xv = __ldg(p.x + aframe0 + ti);
                const float4 w = ldg4(p.gw + c), iv = ldg4(p.ginv + c), sh = ldg4(p.gsh + c);
                v.x = fmaxf(fmaf(xv * w.x, iv.x, sh.x), 0.f);
                v.y = fmaxf(fmaf(xv * w.y, iv.y, sh.y), 0.f);
                v.z = fmaxf(fmaf(xv * w.z, iv.z, sh.z), 0.f);
                v.w = fmaxf(fmaf(xv * w.w, iv.w, sh.w), 0.f);
                return v;
            }
            return ldg4(p.src0 + (aframe0 + ti) * p.lda0 + c);
        }
        const int c = kk - K0;
        return ldg4(p.src1 + ((long long)ab * p.t_in1 + (long long)ato * p.stride1) * p.lda1 + c);
    };
    auto load_b = [&](int k0) -> float4 {
        const int kk = k0 + br, n = n0 + bc;
        if (kk >= p.K || n >= p.N) return make_float4(0.f, 0.f, 0.f, 0.f);
        return ldg4(p.W + (long long)kk * p.N + n);
    };
    auto store_tiles = [&](int buf, const float4& a, const float4& b) {
        As[buf][akq + 0][ar] = a.x;
        As[buf][akq + 1][ar] = a.y;
        As[buf][akq + 2][ar] = a.z;
        As[buf][akq + 3][ar] = a.w;
        *reinterpret_cast<float4*>(&Bs[buf][br][bc]) = b;
    };

    float acc[8][8];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

    const int nk = (p.K + BK - 1) / BK;
    float4 ra = load_a(0), rb = load_b(0);
    store_tiles(0, ra, rb);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) { ra = load_a((kt + 1) * BK); rb = load_b((kt + 1) * BK); }
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
            const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            store_tiles(buf ^ 1, ra, rb);
            __syncthreads();
        }
    }

    // epilogue: + shift (+ rank-1 residual), ReLU, 128-bit stores
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
        if (m >= p.M) continue;
        float xr = 0.f;
        if (p.res) {
            const int b = (int)(m / p.t_out), to = (int)(m % p.t_out);
            xr = __ldg(p.x + (long long)b * p.t_inr + (long long)to * p.strider);
        }
#pragma unroll
        for (int hj = 0; hj < 2; ++hj) {
            const int n = n0 + hj * 64 + tx * 4;
            if (n >= p.N) continue;
            const float4 sh = ldg4(p.shift + n);
            float o[4] = {acc[i][hj * 4 + 0] + sh.x, acc[i][hj * 4 + 1] + sh.y, acc[i][hj * 4 + 2] + sh.z,
                          acc[i][hj * 4 + 3] + sh.w};
            if (p.res) {
                const float4 w = ldg4(p.rw + n), iv = ldg4(p.rinv + n), rs = ldg4(p.rsh + n);
                o[0] += fmaf(xr * w.x, iv.x, rs.x);
                o[1] += fmaf(xr * w.y, iv.y, rs.y);
                o[2] += fmaf(xr * w.z, iv.z, rs.z);
                o[3] += fmaf(xr * w.w, iv.w, rs.w);
            }
            if (p.relu) {
#pragma unroll
                for (int q = 0; q < 4; ++q) o[q] = fmaxf(o[q], 0.f);
            }
            *reinterpret_cast<float4*>(p.out + m * p.ldo + n) = make_float4(o[0], o[1], o[2], o[3]);
        }
    }
}

}  // namespace cb_simt
