// Logit head, path_prob, seq_len scaling and CTC greedy decoding -- HBM-bound streaming kernels.
//
//  head_kernel       chiron/rnn.py:89-96    h2 = fw*W[0] + bw*W[1] + bias ; logits = h2 @ Wc + bc      800 B in, 20 B out / frame
//  path_prob_kernel  chiron_eval.py:116-136 mean over ALL T frames of (top1 - top2)                        20 B in / frame
//  seq_len_kernel    chiron_eval.py:337     np.round(seq_len / ratio).astype(int32)  (half to even)
//  greedy_kernel     chiron_eval.py:486-487 tf.nn.ctc_greedy_decoder(merge_repeated=True)                  20 B in, <=1 B out / frame
// This header holds the kernels only (no launch syntax): tests/cuda_emu compiles the same source for the host.
#pragma once

namespace cb_hd {


constexpr int MAX_CLASS = 8;

// One warp per frame; lanes stride over hidden units; shuffle tree for the n_class dot products.
__global__ void __launch_bounds__(256) head_kernel(const float* __restrict__ lasth, long long M, int H, int C,
                                                   const float* __restrict__ w, const float* __restrict__ bias,
                                                   const float* __restrict__ wc, const float* __restrict__ bc,
                                                   float* __restrict__ logits) {
    const int lane = threadIdx.x & 31;
    const long long warp0 = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long m = warp0; m < M; m += nwarps) {
        const float* row = lasth + m * 2 * H;
        float part[MAX_CLASS];
#pragma unroll
        for (int c = 0; c < MAX_CLASS; ++c) part[c] = 0.f;
        for (int u = lane; u < H; u += 32) {
            const float h2 = (row[u] * __ldg(w + u) + row[H + u] * __ldg(w + H + u)) + __ldg(bias + u);
#pragma unroll
            for (int c = 0; c < MAX_CLASS; ++c)
                if (c < C) part[c] = fmaf(h2, __ldg(wc + u * C + c), part[c]);
        }
#pragma unroll
        for (int c = 0; c < MAX_CLASS; ++c) {
            if (c < C) {
#pragma unroll
                for (int o = 16; o > 0; o >>= 1) part[c] += __shfl_xor_sync(0xffffffffu, part[c], o);
            }
        }
        if (lane == 0) {
#pragma unroll
            for (int c = 0; c < MAX_CLASS; ++c)
                if (c < C) logits[m * C + c] = part[c] + __ldg(bc + c);
        }
    }
}

// Time-major variant for the tensor-core LSTM stack: lasth is [T][2 x H/4][Bp][4]; one thread per frame, consecutive
// threads = consecutive batch rows, so every load is a 128-bit piece of a 512-byte coalesced segment.
__global__ void __launch_bounds__(128) head_tmajor_kernel(const float* __restrict__ lasth, int B, int Bp, int T, int H,
                                                          int C, const float* __restrict__ w,
                                                          const float* __restrict__ bias, const float* __restrict__ wc,
                                                          const float* __restrict__ bc, float* __restrict__ logits) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    const int t = blockIdx.y;
    if (b >= B) return;
    const int H4 = H / 4;
    const float4* src = reinterpret_cast<const float4*>(lasth) + (size_t)t * 2 * H4 * Bp + b;
    float acc[MAX_CLASS];
#pragma unroll
    for (int c = 0; c < MAX_CLASS; ++c) acc[c] = 0.f;
    for (int g = 0; g < H4; ++g) {
        const float4 f = src[(size_t)g * Bp], r = src[(size_t)(H4 + g) * Bp];
        const float fw[4] = {f.x, f.y, f.z, f.w}, bw[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const int u = g * 4 + e;
            const float h2 = (fw[e] * __ldg(w + u) + bw[e] * __ldg(w + H + u)) + __ldg(bias + u);
#pragma unroll
            for (int c = 0; c < MAX_CLASS; ++c)
                if (c < C) acc[c] = fmaf(h2, __ldg(wc + u * C + c), acc[c]);
        }
    }
    float* dst = logits + ((size_t)b * T + t) * C;
#pragma unroll
    for (int c = 0; c < MAX_CLASS; ++c)
        if (c < C) dst[c] = acc[c] + __ldg(bc + c);
}

__device__ __forceinline__ void top2_argmax(const float* __restrict__ row, int C, float& d, int& am) {
    float v0 = row[0], v1 = -INFINITY;
    am = 0;
    for (int c = 1; c < C; ++c) {
        const float v = row[c];
        if (v > v0) { v1 = v0; v0 = v; am = c; }       // strict >: first maximum wins (Eigen argmax / top_k order)
        else if (v > v1) v1 = v;
    }
    d = v0 - v1;
}

// One warp per window.
__global__ void __launch_bounds__(128) path_prob_kernel(const float* __restrict__ logits, int B, int T, int C,
                                                        float* __restrict__ prob) {
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    const float* base = logits + (long long)b * T * C;
    float sum = 0.f;
    for (int t = lane; t < T; t += 32) {
        float d; int am;
        top2_argmax(base + (long long)t * C, C, d, am);
        sum += d;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    if (lane == 0) prob[b] = sum / (float)T;
}

__global__ void seq_len_kernel(const int32_t* __restrict__ in, int B, int L, int T, int32_t* __restrict__ out) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    const double ratio = (double)L / (double)T;
    out[b] = (int32_t)rint((double)in[b] / ratio);         // rint = round half to even, like np.round
}

// One warp per window: argmax per frame, drop repeats and blanks, ballot/popc stream compaction.
__global__ void __launch_bounds__(128) greedy_kernel(const float* __restrict__ logits, const int32_t* __restrict__ lens,
                                                     int B, int T, int C, int8_t* __restrict__ bases,
                                                     int32_t* __restrict__ n_bases) {
    const int lane = threadIdx.x & 31;
    const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (b >= B) return;
    int len = lens[b];
    len = len < 0 ? 0 : (len > T ? T : len);
    const float* base = logits + (long long)b * T * C;
    int8_t* dst = bases + (long long)b * T;
    const int blank = C - 1;
    int count = 0, carry = -1;
    for (int t0 = 0; t0 < len; t0 += 32) {
        const int t = t0 + lane;
        int am = -1;
        if (t < len) { float d; top2_argmax(base + (long long)t * C, C, d, am); }
        int prev = __shfl_up_sync(0xffffffffu, am, 1);
        if (lane == 0) prev = carry;
        const bool keep = t < len && am != blank && am != prev;
        const unsigned mask = __ballot_sync(0xffffffffu, keep);
        if (keep) dst[count + __popc(mask & ((1u << lane) - 1u))] = (int8_t)am;
        count += __popc(mask);
        carry = __shfl_sync(0xffffffffu, am, 31);
    }
    for (int i = count + lane; i < T; i += 32) dst[i] = 0;
    if (lane == 0) n_bases[b] = count;
}

}  // namespace cb_hd
