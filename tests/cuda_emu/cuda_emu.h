// Minimal host emulation of the CUDA execution model -- TEST INFRASTRUCTURE ONLY (tests/test_cuda_emu.py).
//
// Lets g++ compile the kernel headers of the fp32 path (chiron_b200/csrc/cb_bn_kernels.cuh, cb_gemm_simt_kernel.cuh) and
// run them on a machine without a GPU: every CUDA thread of a block is a std::thread, __syncthreads() is a std::barrier
// over the block, warp shuffles exchange through a per-warp buffer between two warp barriers, __shared__ variables are
// function-local statics (blocks run one after the other).  Only what those kernels use is provided.  Nothing in the
// product path includes this file.
#pragma once
#include <algorithm>
#include <array>
#include <atomic>
#include <barrier>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <memory>
#include <thread>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline
#define __restrict__
#define __launch_bounds__(...)
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct alignas(16) float4 { float x, y, z, w; };
static inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }

namespace emu {
struct Dim { unsigned x, y, z; };
struct BlockCtx {
    unsigned char* dyn_smem;
    std::barrier<>* block_bar;
    std::vector<std::unique_ptr<std::barrier<>>>* warp_bar;
    std::vector<std::array<double, 32>>* shfl;
    unsigned n_threads;
};
inline thread_local Dim t_idx{0, 0, 0}, b_idx{0, 0, 0};
inline thread_local BlockCtx* ctx = nullptr;
inline Dim b_dim{1, 1, 1}, g_dim{1, 1, 1};
inline long long launches = 0, threads_run = 0;

inline unsigned char* dyn_smem() { return ctx->dyn_smem; }

// Run body() once per CUDA thread of a (grid_x, grid_y) x block launch (1-D blocks) with dyn_bytes of dynamic shared memory.
template <class F>
void launch2d(unsigned grid_x, unsigned grid_y, unsigned block, size_t dyn_bytes, F&& body) {
    g_dim = Dim{grid_x, grid_y, 1};
    b_dim = Dim{block, 1, 1};
    const unsigned warps = (block + 31) / 32;
    std::vector<unsigned char> dyn(dyn_bytes + 64);
    unsigned char* dyn_base = dyn.data() + (64 - reinterpret_cast<uintptr_t>(dyn.data()) % 64) % 64;
    for (unsigned by = 0; by < grid_y; ++by)
    for (unsigned blk = 0; blk < grid_x; ++blk) {
        std::barrier<> bb((std::ptrdiff_t)block);
        std::vector<std::unique_ptr<std::barrier<>>> wb;
        for (unsigned w = 0; w < warps; ++w)
            wb.emplace_back(new std::barrier<>((std::ptrdiff_t)std::min(32u, block - 32 * w)));
        std::vector<std::array<double, 32>> sh(warps);
        BlockCtx c{dyn_base, &bb, &wb, &sh, block};
        std::vector<std::thread> th;
        th.reserve(block);
        for (unsigned t = 0; t < block; ++t)
            th.emplace_back([&, t, blk, by] {
                t_idx = Dim{t, 0, 0};
                b_idx = Dim{blk, by, 0};
                ctx = &c;
                body();
                // a thread that leaves early must not strand the others at a later barrier
                c.block_bar->arrive_and_drop();
                (*c.warp_bar)[t / 32]->arrive_and_drop();
            });
        for (auto& x : th) x.join();
        threads_run += block;
    }
    ++launches;
}

template <class F>
void launch(unsigned grid, unsigned block, F&& body) { launch2d(grid, 1, block, 0, static_cast<F&&>(body)); }
}  // namespace emu

#define threadIdx (emu::t_idx)
#define blockIdx (emu::b_idx)
#define blockDim (emu::b_dim)
#define gridDim (emu::g_dim)

template <class T>
static inline T __ldg(const T* p) { return *p; }

static inline int atomicMax(int* addr, int v) {
    std::atomic_ref<int> a(*addr);
    int old = a.load();
    while (old < v && !a.compare_exchange_weak(old, v)) {}
    return old;
}

static inline void __syncthreads() { emu::ctx->block_bar->arrive_and_wait(); }

// Warp collectives: all live lanes of the calling warp take part (the kernels only call them warp-uniformly).  A value is
// published in the warp's exchange buffer between two warp barriers.
namespace emu {
struct WarpPos { BlockCtx* c; unsigned w, lane, lanes; };
inline WarpPos warp_pos() {
    BlockCtx* c = ctx;
    const unsigned t = t_idx.x, w = t / 32;
    return WarpPos{c, w, t % 32, std::min(32u, c->n_threads - 32 * w)};
}
template <class T, class Pick>
inline T exchange(T v, Pick pick) {           // pick(lane, lanes) -> source lane, or -1 to keep the own value
    static_assert(sizeof(T) <= sizeof(double), "exchange slot is 8 bytes");
    const WarpPos p = warp_pos();
    std::memcpy(&(*p.c->shfl)[p.w][p.lane], &v, sizeof(T));
    (*p.c->warp_bar)[p.w]->arrive_and_wait();
    const int src = pick((int)p.lane, (int)p.lanes);
    T r = v;
    if (src >= 0 && src < (int)p.lanes) std::memcpy(&r, &(*p.c->shfl)[p.w][src], sizeof(T));
    (*p.c->warp_bar)[p.w]->arrive_and_wait();
    return r;
}
}  // namespace emu

template <class T>
static inline T __shfl_down_sync(unsigned, T v, int delta) {
    return emu::exchange(v, [delta](int lane, int lanes) { return lane + delta < lanes ? lane + delta : -1; });
}
template <class T>
static inline T __shfl_up_sync(unsigned, T v, int delta) {
    return emu::exchange(v, [delta](int lane, int) { return lane - delta >= 0 ? lane - delta : -1; });
}
template <class T>
static inline T __shfl_xor_sync(unsigned, T v, int mask) {
    return emu::exchange(v, [mask](int lane, int lanes) { return (lane ^ mask) < lanes ? (lane ^ mask) : -1; });
}
template <class T>
static inline T __shfl_sync(unsigned, T v, int src) {
    return emu::exchange(v, [src](int, int) { return src & 31; });
}
static inline unsigned __ballot_sync(unsigned, bool pred) {
    unsigned m = 0;
    for (int l = 0; l < 32; ++l) {
        const int bit = emu::exchange((int)pred, [l](int, int) { return l; });
        const emu::WarpPos p = emu::warp_pos();
        if (l < (int)p.lanes && bit) m |= 1u << l;
    }
    return m;
}
static inline bool __any_sync(unsigned mask, bool pred) { return __ballot_sync(mask, pred) != 0; }
static inline int __popc(unsigned v) { return __builtin_popcount(v); }
static inline int __ffs(unsigned v) { return __builtin_ffs((int)v); }
static inline void __syncwarp(unsigned = 0xffffffffu) {
    const emu::WarpPos p = emu::warp_pos();
    (*p.c->warp_bar)[p.w]->arrive_and_wait();
}
static inline int atomicExch(int* addr, int v) { return std::atomic_ref<int>(*addr).exchange(v); }
static inline int atomicAdd(int* addr, int v) { return std::atomic_ref<int>(*addr).fetch_add(v); }
static inline double atomicAdd(double* addr, double v) {
    std::atomic_ref<double> a(*addr);
    double old = a.load();
    while (!a.compare_exchange_weak(old, old + v)) {}
    return old;
}
