// fp32 FFMA persistent LSTM recurrence (CB_PREC_FP32).
//
// Replaces the tf.while_loop that dynamic_rnn builds around LSTMCell (chiron/rnn.py:49-50,64,140-143): per step
// z = pre[b,t,:] + h @ W_hh ; i,j,f,o = split(z) ; c' = sigmoid(f+1)*c + sigmoid(i)*tanh(j) ; h' = sigmoid(o)*tanh(c').
// `pre` is the hoisted input projection (x_t @ W_ih + bias) produced by the GEMM kernel.  Batch rows are independent,
// so one CTA owns R = 16*RG rows of one direction for all T steps: W_hh (H x 4H fp32, 160 KB for H=100) stays resident
// in shared memory, h ping-pongs through shared memory, c lives in registers; no grid-wide synchronisation.
// sequence_length semantics of dynamic_rnn: for t >= len the output is zero and the state frozen; the backward
// direction is reverse_sequence over the first len frames, i.e. the same recurrence walking t = len-1 .. 0.
// This header holds the kernel only (no launch syntax): tests/cuda_emu compiles the same source for the host.
#pragma once

namespace cb_lstm {


constexpr int RPT = 16;   // rows per thread

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int RG>
__global__ void __launch_bounds__(128 * RG) lstm_simt_kernel(const LstmProblem p) {
    constexpr int R = RPT * RG, RS = R + 4;
#ifdef CB_HOST_EMU
    float* smem = reinterpret_cast<float*>(emu::dyn_smem());
#else
    extern __shared__ __align__(16) float smem[];
#endif
    const int H = p.H, H4 = 4 * p.H;
    float* Ws = smem;                       // [H][4H]
    float* hs = smem + H * H4;              // [2][H][RS]
    __shared__ int lens_s[R];
    __shared__ int max_len_s;

    const int tid = threadIdx.x;
    const int u = tid & 127, rg = tid >> 7;
    const int rbase = rg * RPT;
    const int b0 = blockIdx.x * R;
    const bool u_ok = u < H;
    const int dir = blockIdx.y;             // 0 = forward, 1 = backward (reverse_sequence over the first len frames)
    const int col0 = dir * H4, ocol0 = dir * H;
    const bool reverse = dir != 0;
    const float* __restrict__ whh = p.whh[dir];

    for (int i = tid * 4; i < H * H4; i += blockDim.x * 4)
        *reinterpret_cast<float4*>(Ws + i) = __ldg(reinterpret_cast<const float4*>(whh + i));
    for (int i = tid; i < 2 * H * RS; i += blockDim.x) hs[i] = 0.f;
    if (tid == 0) max_len_s = 0;
    __syncthreads();
    if (tid < R) {
        const int b = b0 + tid;
        int l = 0;
        if (b < p.B) { l = p.lens[b]; l = l < 0 ? 0 : (l > p.T ? p.T : l); }
        lens_s[tid] = l;
        atomicMax(&max_len_s, l);
    }
    __syncthreads();
    const int max_len = max_len_s;

    float c[RPT], acc[RPT][4];
    int len_r[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) { c[r] = 0.f; len_r[r] = lens_s[rbase + r]; }

    auto load_pre = [&](int s) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            const bool act = u_ok && s < len_r[r];
            if (act) {
                const int t = reverse ? len_r[r] - 1 - s : s;
                const float* src = p.pre + ((long long)(b0 + rbase + r) * p.T + t) * p.ld_pre + col0 + u;
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[r][g] = __ldg(src + g * H);
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) acc[r][g] = 0.f;
            }
        }
    };

    load_pre(0);
    int cur = 0;
    for (int s = 0; s < p.T; ++s) {
        if (s < max_len) {
            if (u_ok) {
                const float* hb = hs + cur * H * RS + rbase;
                for (int k = 0; k < H; ++k) {
                    const float* wr = Ws + k * H4 + u;
                    const float w0 = wr[0], w1 = wr[H], w2 = wr[2 * H], w3 = wr[3 * H];
#pragma unroll
                    for (int q = 0; q < RPT / 4; ++q) {
                        const float4 hv = *reinterpret_cast<const float4*>(hb + k * RS + q * 4);
                        const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            acc[q * 4 + e][0] = fmaf(hh[e], w0, acc[q * 4 + e][0]);
                            acc[q * 4 + e][1] = fmaf(hh[e], w1, acc[q * 4 + e][1]);
                            acc[q * 4 + e][2] = fmaf(hh[e], w2, acc[q * 4 + e][2]);
                            acc[q * 4 + e][3] = fmaf(hh[e], w3, acc[q * 4 + e][3]);
                        }
                    }
                }
            }
        }
        // gates, state update, output
        if (u_ok) {
            float* hn = hs + (cur ^ 1) * H * RS + u * RS + rbase;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int b = b0 + rbase + r;
                float hnew = 0.f;
                if (s < len_r[r]) {
                    const float gi = acc[r][0], gj = acc[r][1], gf = acc[r][2], go = acc[r][3];
                    const float cn = sigmoid_acc(gf + 1.0f) * c[r] + sigmoid_acc(gi) * tanhf(gj);
                    hnew = sigmoid_acc(go) * tanhf(cn);
                    c[r] = cn;
                    const int t = reverse ? len_r[r] - 1 - s : s;
                    p.out[((long long)b * p.T + t) * p.ldo + ocol0 + u] = hnew;
                } else if (b < p.B) {
                    p.out[((long long)b * p.T + s) * p.ldo + ocol0 + u] = 0.f;   // frames t >= len are zero
                }
                hn[r] = hnew;
            }
        }
        if (s + 1 < p.T) load_pre(s + 1);
        __syncthreads();
        cur ^= 1;
    }
}

// Dynamic shared memory of lstm_simt_kernel<RG> in bytes.
inline size_t lstm_smem_bytes(int H, int RG) { return (size_t)(H * 4 * H + 2 * H * (RPT * RG + 4)) * sizeof(float); }

}  // namespace cb_lstm
