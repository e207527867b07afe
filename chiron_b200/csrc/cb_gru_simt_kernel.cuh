// fp32 FFMA persistent GRU recurrence (CB_PREC_FP32; cell_type = GRU in the CBW1 header).
//
// Replaces the tf.while_loop dynamic_rnn builds around GRUCell (chiron/rnn.py:51-53,129-131; TF 1.15
// rnn_cell_impl.GRUCell.call): per step
//   r, u = sigmoid(pre_g[b,t,:] + h @ Wg_h) ;  c = tanh(pre_c[b,t,:] + (r*h) @ Wc_h) ;  h' = u*h + (1-u)*c
// `pre` is the hoisted input projection (x_t @ [Kg_x | Kc_x] + [bg | bc], columns r,u,c per direction) produced by the GEMM
// kernel.  Same ownership as the LSTM kernel (cb_lstm_simt.cu): one CTA owns R = 16*RG batch rows of one direction for all T
// steps, Wg_h (H x 2H) and Wc_h (H x H) stay resident in shared memory, a thread owns one hidden unit of 16 rows and keeps
// that unit's state in registers.  The candidate needs r*h of EVERY unit, so a step has two phases separated by a CTA
// barrier: gates -> r*h to shared memory | candidate -> new state to shared memory.  sequence_length semantics of
// dynamic_rnn: for t >= len the output is zero and the state frozen; the backward direction walks t = len-1 .. 0.
// Roofline: FFMA / shared-memory bound latency chain (reference-grade path; no GRU checkpoint ships with the reference).
// This header holds the kernel only (no launch syntax): tests/cuda_emu compiles the same source for the host.
#pragma once

namespace cb_gru {

constexpr int RPT = 16;   // rows per thread

__device__ __forceinline__ float sigmoid_acc(float x) { return 1.0f / (1.0f + expf(-x)); }

template <int RG>
__global__ void __launch_bounds__(128 * RG) gru_simt_kernel(const GruProblem p) {
    constexpr int R = RPT * RG, RS = R + 4;
#ifdef CB_HOST_EMU
    float* smem = reinterpret_cast<float*>(emu::dyn_smem());
#else
    extern __shared__ __align__(16) float smem[];
#endif
    const int H = p.H, H2 = 2 * p.H;
    float* Wg = smem;                       // [H][2H]  recurrent gate kernel (columns r | u)
    float* Wc = Wg + H * H2;                // [H][H]   recurrent candidate kernel
    float* hs = Wc + H * H;                 // [H][RS]  state h, unit-major
    float* rh = hs + H * RS;                // [H][RS]  r * h
    __shared__ int lens_s[R];
    __shared__ int max_len_s;

    const int tid = threadIdx.x;
    const int u = tid & 127, rg = tid >> 7;
    const int rbase = rg * RPT;
    const int b0 = blockIdx.x * R;
    const bool u_ok = u < H;
    const int dir = blockIdx.y;             // 0 = forward, 1 = backward (reverse_sequence over the first len frames)
    const int col0 = dir * 3 * H, ocol0 = dir * H;
    const bool reverse = dir != 0;

    for (int i = tid * 4; i < H * H2; i += blockDim.x * 4)
        *reinterpret_cast<float4*>(Wg + i) = __ldg(reinterpret_cast<const float4*>(p.wg[dir] + i));
    for (int i = tid * 4; i < H * H; i += blockDim.x * 4)
        *reinterpret_cast<float4*>(Wc + i) = __ldg(reinterpret_cast<const float4*>(p.wc[dir] + i));
    for (int i = tid; i < 2 * H * RS; i += blockDim.x) hs[i] = 0.f;      // hs and rh
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

    float h_own[RPT], ar[RPT], au[RPT], ac[RPT];
    int len_r[RPT];
#pragma unroll
    for (int r = 0; r < RPT; ++r) { h_own[r] = 0.f; len_r[r] = lens_s[rbase + r]; }

    auto load_pre = [&](int s) {
#pragma unroll
        for (int r = 0; r < RPT; ++r) {
            if (u_ok && s < len_r[r]) {
                const int t = reverse ? len_r[r] - 1 - s : s;
                const float* src = p.pre + ((long long)(b0 + rbase + r) * p.T + t) * p.ld_pre + col0 + u;
                ar[r] = __ldg(src); au[r] = __ldg(src + H); ac[r] = __ldg(src + 2 * H);
            } else {
                ar[r] = 0.f; au[r] = 0.f; ac[r] = 0.f;
            }
        }
    };

    load_pre(0);
    for (int s = 0; s < p.T; ++s) {
        // ---- phase 1: gates r, u of this unit; publish r*h -------------------------------------------------------------------
        if (s < max_len && u_ok) {
            const float* hb = hs + rbase;
            for (int k = 0; k < H; ++k) {
                const float w0 = Wg[k * H2 + u], w1 = Wg[k * H2 + H + u];
#pragma unroll
                for (int q = 0; q < RPT / 4; ++q) {
                    const float4 hv = *reinterpret_cast<const float4*>(hb + k * RS + q * 4);
                    const float hh[4] = {hv.x, hv.y, hv.z, hv.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        ar[q * 4 + e] = fmaf(hh[e], w0, ar[q * 4 + e]);
                        au[q * 4 + e] = fmaf(hh[e], w1, au[q * 4 + e]);
                    }
                }
            }
            float* dst = rh + u * RS + rbase;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                au[r] = sigmoid_acc(au[r]);                       // update gate, kept for phase 2
                dst[r] = s < len_r[r] ? sigmoid_acc(ar[r]) * h_own[r] : 0.f;
            }
        }
        __syncthreads();
        // ---- phase 2: candidate from r*h of every unit; new state ----------------------------------------------------------
        if (u_ok) {
            if (s < max_len) {
                const float* rb = rh + rbase;
                for (int k = 0; k < H; ++k) {
                    const float w = Wc[k * H + u];
#pragma unroll
                    for (int q = 0; q < RPT / 4; ++q) {
                        const float4 v = *reinterpret_cast<const float4*>(rb + k * RS + q * 4);
                        ac[q * 4 + 0] = fmaf(v.x, w, ac[q * 4 + 0]);
                        ac[q * 4 + 1] = fmaf(v.y, w, ac[q * 4 + 1]);
                        ac[q * 4 + 2] = fmaf(v.z, w, ac[q * 4 + 2]);
                        ac[q * 4 + 3] = fmaf(v.w, w, ac[q * 4 + 3]);
                    }
                }
            }
            float* hn = hs + u * RS + rbase;
#pragma unroll
            for (int r = 0; r < RPT; ++r) {
                const int b = b0 + rbase + r;
                if (s < len_r[r]) {
                    const float c = tanhf(ac[r]);
                    const float hnew = au[r] * h_own[r] + (1.0f - au[r]) * c;
                    h_own[r] = hnew;
                    hn[r] = hnew;
                    const int t = reverse ? len_r[r] - 1 - s : s;
                    p.out[((long long)b * p.T + t) * p.ldo + ocol0 + u] = hnew;
                } else if (b < p.B) {
                    p.out[((long long)b * p.T + s) * p.ldo + ocol0 + u] = 0.f;   // frames t >= len are zero; state frozen
                }
            }
        }
        if (s + 1 < p.T) load_pre(s + 1);
        __syncthreads();
    }
}

// Dynamic shared memory of gru_simt_kernel<RG> in bytes.
inline size_t gru_smem_bytes(int H, int RG) { return (size_t)(H * 2 * H + H * H + 2 * H * (RPT * RG + 4)) * sizeof(float); }

}  // namespace cb_gru
