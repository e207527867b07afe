// Microbenchmark: what does reading a TMEM accumulator into registers (tcgen05.ld) cost while the tensor core is busy?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I chiron_b200/csrc -o build/ldtm_bench tools/experiments/ldtm_bench.cu
// One CTA per SM: warp 8 issues a stream of tcgen05.mma (M=128, N=256, K=16, fp16 -> fp32) into TMEM columns [0,256); warps 0-7
// (two per lane quadrant, 128 columns each, like the GEMM epilogue) drain columns [256,512) `reps` times with `depth` x16 loads
// in flight per wait (depth 8 = one wait per 128 columns) or with .x64 loads.  Modes: 1 = MMA only, 2 = loads only, 3 = both.
// Prints clocks per MMA and clocks per 128x256 drain, so that (3) - (1),(2) shows how much the two overlap.
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

#include "cb_tc_common.cuh"

__device__ __forceinline__ void tmem_ld64(uint32_t taddr, uint32_t (&v)[64]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, %17, %18, %19, %20, %21, %22, %23, %24, %25, "
        "%26, %27, %28, %29, %30, %31, %32, %33, %34, %35, %36, %37, %38, %39, %40, %41, %42, %43, %44, %45, %46, %47, %48, %49, "
        "%50, %51, %52, %53, %54, %55, %56, %57, %58, %59, %60, %61, %62, %63}, [%64];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]), "=r"(v[9]),
          "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]), "=r"(v[16]), "=r"(v[17]), "=r"(v[18]),
          "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]), "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]),
          "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31]), "=r"(v[32]), "=r"(v[33]), "=r"(v[34]), "=r"(v[35]), "=r"(v[36]),
          "=r"(v[37]), "=r"(v[38]), "=r"(v[39]), "=r"(v[40]), "=r"(v[41]), "=r"(v[42]), "=r"(v[43]), "=r"(v[44]), "=r"(v[45]),
          "=r"(v[46]), "=r"(v[47]), "=r"(v[48]), "=r"(v[49]), "=r"(v[50]), "=r"(v[51]), "=r"(v[52]), "=r"(v[53]), "=r"(v[54]),
          "=r"(v[55]), "=r"(v[56]), "=r"(v[57]), "=r"(v[58]), "=r"(v[59]), "=r"(v[60]), "=r"(v[61]), "=r"(v[62]), "=r"(v[63])
        : "r"(taddr));
}

template <int DEPTH>      // DEPTH x16 loads in flight per wait; DEPTH == 0: .x64 loads (one wait per 64 columns)
__global__ void __launch_bounds__(320, 1) ldtm_kernel(int mode, int n_mma, int reps, long long* out, float* sink) {
    extern __shared__ __align__(1024) uint8_t smem[];
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + 24 * 1024);
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar + 2);
    volatile int* done = reinterpret_cast<volatile int*>(bar + 3);      // mode 3: the load loops run until the MMA stream ends
    if (threadIdx.x == 0) *done = 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (uint32_t i = threadIdx.x * 16; i < 24 * 1024; i += 320 * 16) *reinterpret_cast<uint4*>(smem + i) = make_uint4(0, 0, 0, 0);
    if (threadIdx.x == 0) { mbar_init(bar, 1); fence_barrier_init(); }
    if (warp == 8) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    long long t0 = 0, t1 = 0;
    if (warp == 8) {
        if (mode & 1) {
            const bool leader = elect_one();
            const uint32_t idesc = make_idesc_f16(128, 256);
            const uint64_t da = make_desc(smem_u32(smem), 128 * 16, 128), db = make_desc(smem_u32(smem) + 8192, 256 * 16, 128);
            __syncwarp();
            t0 = clock64();
            if (leader) {
                for (int i = 0; i < n_mma; ++i) umma_f16(tmem_base, da, db, idesc, 1);
                umma_commit(bar);
                mbar_wait(bar, 0);
            }
            __syncwarp();
            t1 = clock64();
            if (lane == 0) { out[blockIdx.x * 4 + 0] = t1 - t0; *done = 1; }
        }
    } else if (warp < 8) {
        if (mode & 2) {
            const int quad = warp & 3, chalf = warp >> 2;
            const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + 256 + chalf * 128;
            float acc[16];                    // (no per-column accumulators: this measures the loads, not register pressure)
#pragma unroll
            for (int e = 0; e < 16; ++e) acc[e] = 0.f;
            __syncwarp();
            t0 = clock64();
            int r = 0;
            for (; mode == 3 ? !*done : r < reps; ++r) {
                if constexpr (DEPTH == 0) {
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        uint32_t v[64];
                        tmem_ld64(taddr + j * 64, v);
                        tmem_ld_wait();
#pragma unroll
                        for (int e = 0; e < 64; ++e) acc[e & 15] += __uint_as_float(v[e]);
                    }
                } else {
#pragma unroll
                    for (int j = 0; j < 8; j += DEPTH) {
                        uint32_t v[DEPTH][16];
#pragma unroll
                        for (int d = 0; d < DEPTH; ++d) tmem_ld16(taddr + (j + d) * 16, v[d]);
                        tmem_ld_wait();
#pragma unroll
                        for (int d = 0; d < DEPTH; ++d)
#pragma unroll
                            for (int e = 0; e < 16; ++e) acc[e] += __uint_as_float(v[d][e]);
                    }
                }
            }
            t1 = clock64();
            float s = 0.f;
#pragma unroll
            for (int e = 0; e < 16; ++e) s += acc[e];
            if (s == 12345.678f) sink[0] = s;
            if (lane == 0 && warp == 0) { out[blockIdx.x * 4 + 1] = t1 - t0; out[blockIdx.x * 4 + 2] = r; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 8) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base));
    }
}

template <int DEPTH>
void run(int mode, int n_mma, int reps, long long* d_out, float* d_sink, const char* name) {
    cudaFuncSetAttribute(ldtm_kernel<DEPTH>, cudaFuncAttributeMaxDynamicSharedMemorySize, 32 * 1024);
    long long h[148 * 4];
    for (int it = 0; it < 2; ++it) {
        cudaMemset(d_out, 0, sizeof(h));
        ldtm_kernel<DEPTH><<<148, 320, 25 * 1024>>>(mode, n_mma, reps, d_out, d_sink);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("%s: %s\n", name, cudaGetErrorString(e)); exit(1); }
    }
    cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
    double mma = 0, ld = 0, cnt = 0;
    for (int b = 0; b < 148; ++b) { mma += h[b * 4]; ld += h[b * 4 + 1]; cnt += h[b * 4 + 2]; }
    printf("%-28s mode %d: %8.1f clk per MMA (M128 N256 K16; floor 128)   %8.1f clk per 128x256 fp32 drain (%.0f drains)\n", name,
           mode, n_mma ? mma / 148 / n_mma : 0.0, cnt ? ld / cnt : 0.0, cnt / 148);
}

int main() {
    long long* d_out; float* d_sink;
    cudaMalloc(&d_out, 148 * 4 * sizeof(long long));
    cudaMalloc(&d_sink, 64);
    const int n_mma = 4096;
    run<1>(1, n_mma, 0, d_out, d_sink, "MMA only");
    run<1>(2, 0, 256, d_out, d_sink, "loads only, x16 depth 1");
    run<2>(2, 0, 256, d_out, d_sink, "loads only, x16 depth 2");
    run<4>(2, 0, 256, d_out, d_sink, "loads only, x16 depth 4");
    run<8>(2, 0, 256, d_out, d_sink, "loads only, x16 depth 8");
    run<0>(2, 0, 256, d_out, d_sink, "loads only, x64");
    // both: the load loop is sized to last about as long as the MMA stream
    run<1>(3, n_mma, 128, d_out, d_sink, "both, x16 depth 1");
    run<2>(3, n_mma, 192, d_out, d_sink, "both, x16 depth 2");
    run<4>(3, n_mma, 256, d_out, d_sink, "both, x16 depth 4");
    run<8>(3, n_mma, 256, d_out, d_sink, "both, x16 depth 8");
    run<0>(3, n_mma, 256, d_out, d_sink, "both, x64");
    return 0;
}
