// Stem convolution of the raw one-channel signal (fp32 path): conv_layer(net, [1, k, 1, C], strides=s) + BN + ReLU in front of
// the residual blocks of RNA_model2 (k=9, s=5) and RNA_model3 (k=14, s=7), chiron/cnn.py:454-476, with TF 'SAME' padding.
// One thread produces four channels of one output frame from the k samples under the window (k <= 16 samples, L1/L2 hits
// shared by the C/4 threads of the frame); the kernel is bound by its HBM write (4*C bytes per output frame).
// This header holds the kernel only (no launch syntax): tests/cuda_emu compiles the same source for the host.
#pragma once

namespace cb_stem {

constexpr int STEM_THREADS = 256;

__global__ void __launch_bounds__(STEM_THREADS) stem_conv_kernel(const StemProblem p) {
    const int qw = p.C >> 2;
    const long long n = (long long)p.B * p.t_out * qw;
    for (long long i = (long long)blockIdx.x * STEM_THREADS + threadIdx.x; i < n; i += (long long)gridDim.x * STEM_THREADS) {
        const long long m = i / qw;
        const int c = (int)(i - m * qw) * 4;
        const long long b = m / p.t_out;
        const int to = (int)(m - b * p.t_out);
        const float* xs = p.x + b * p.t_in;
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
        for (int j = 0; j < p.k; ++j) {
            const int ti = to * p.stride + j - p.left;
            if (ti < 0 || ti >= p.t_in) continue;
            const float xv = __ldg(xs + ti);
            const float4 w = __ldg(reinterpret_cast<const float4*>(p.w + (long long)j * p.C + c));
            acc[0] = fmaf(xv, w.x, acc[0]); acc[1] = fmaf(xv, w.y, acc[1]);
            acc[2] = fmaf(xv, w.z, acc[2]); acc[3] = fmaf(xv, w.w, acc[3]);
        }
        if (p.inv) {
            const float4 iv = __ldg(reinterpret_cast<const float4*>(p.inv + c)), sh = __ldg(reinterpret_cast<const float4*>(p.shift + c));
            acc[0] = fmaf(acc[0], iv.x, sh.x); acc[1] = fmaf(acc[1], iv.y, sh.y);
            acc[2] = fmaf(acc[2], iv.z, sh.z); acc[3] = fmaf(acc[3], iv.w, sh.w);
        }
        if (p.relu) {
#pragma unroll
            for (int q = 0; q < 4; ++q) acc[q] = fmaxf(acc[q], 0.f);
        }
        *reinterpret_cast<float4*>(p.out + m * p.C + c) = make_float4(acc[0], acc[1], acc[2], acc[3]);
    }
}

// One wave of resident CTAs (8 x 256 threads per SM at <= 32 registers) or fewer for small problems.
inline int stem_grid(int sm_count, long long items) {
    long long g = (items + STEM_THREADS - 1) / STEM_THREADS;
    const long long cap = (long long)(sm_count > 0 ? sm_count : 148) * 8;
    if (g > cap) g = cap;
    return g < 1 ? 1 : (int)g;
}

}  // namespace cb_stem
