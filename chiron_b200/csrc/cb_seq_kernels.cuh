// CTC beam search and overlap assembly on the GPU (latency/branch-bound integer work; see cb_seq_algos.cuh for the
// algorithms and the reference code they restate).
//
//  beam_warp_kernel one warp per window walks TF's trie beam search cooperatively over shared memory (bit-identical to the
//                   sequential routine cb_beam_decode_one); the product path.  A window whose node pool overflows is
//                   MARKED (n_bases = -1) for the next pass.
//  beam_retry_kernel second pass: the marked windows alone, one per CTA with the CTA's whole shared memory as node pool.
//  beam_kernel      one thread per window runs cb_beam_decode_one over a global workspace with a pool that cannot
//                   overflow: third pass for windows still marked (workspace slots claimed atomically), and the
//                   reference the cooperative kernels are tested against.
//  The three passes are enqueued back to back with no host synchronisation: a pass finds nothing to do unless the one
//  before it marked a window.
//  assembly         asm_compact (drop empty windows like sparse2dense, chiron_eval.py:56-66) -> asm_disp (one thread per
//                   adjacent window pair: stick / glue / difflib-exact simple displacement) -> asm_scan (prefix sum ->
//                   window coordinates, read length) -> asm_vote (count matrix [4,len] + quality sums, atomics) ->
//                   asm_finish (argmax + phred+33 string, chiron_eval.py:152-174,457).
// This header holds the kernels only (no launch syntax): tests/cuda_emu compiles the same source for the host.
#pragma once
#include "cb_seq_algos.cuh"

namespace cb_seq {


inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- beam search -------------------------------------------------------------------------------------------------
// slot_counter == nullptr: every window, workspace b.  Otherwise only the windows marked n_bases == -1, each claiming one of
// n_slots workspaces; a marked window that finds no slot raises *overflow (the launcher sizes the pool so that nothing else can).
__global__ void __launch_bounds__(64) beam_kernel(const float* __restrict__ logits, const int32_t* __restrict__ lens,
                                                  int B, int T, int C, int W, int pool, char* __restrict__ work,
                                                  size_t work_stride, int8_t* __restrict__ bases,
                                                  int32_t* __restrict__ n_bases, int* __restrict__ overflow,
                                                  int* __restrict__ slot_counter, int n_slots, float* __restrict__ scores) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b >= B) return;
    int8_t* dst = bases + (size_t)b * T;
    size_t ws = (size_t)b;
    if (slot_counter) {
        if (n_bases[b] != -1) return;
        const int slot = atomicAdd(slot_counter, 1);
        if (slot >= n_slots) {
            atomicExch(overflow, 1);
            for (int i = 0; i < T; ++i) dst[i] = 0;
            n_bases[b] = 0;
            return;
        }
        ws = (size_t)slot;
    }
    int len = lens[b];
    len = len < 0 ? 0 : (len > T ? T : len);
    CbBeamWork k = cb_beam_work_carve(work + ws * work_stride, W, pool);
    int n = cb_beam_decode_one(logits + (size_t)b * T * C, len, C, W, k, dst, scores ? scores + b : nullptr);
    if (n < 0) { atomicExch(overflow, 1); n = 0; }
    for (int i = n; i < T; ++i) dst[i] = 0;
    n_bases[b] = n;
}

// One WARP per window, workspace and the window's logits in SHARED memory.  The search is a serial, branchy walk: with 32
// windows per warp (beam_kernel) the lanes serialise each other's divergent paths and every trie / slot access is an
// uncoalesced global load.  Here a warp runs cb_beam_decode_one's frame loop COOPERATIVELY with bit-identical results:
//   * the stable descending sort of the leaves is a rank computation (rank = #greater + #equal-before), the per-branch
//     copies and the probability updates (two log-sum-exp per branch -- the transcendental bulk of a frame) are
//     independent per branch and go one branch per lane, the bottom-of-beam search is a warp reduction;
//   * the extension loop (branch x child, order-dependent: evictions move the threshold) stays on lane 0, but only for the
//     candidates a parallel pre-filter cannot rule out.  Within a frame a branch's oldp only ever drops to -inf, the
//     threshold only rises once the beam is full and the beam only fills up, so "passes with the state at the start of
//     the chunk" is a superset of "passes when reached"; candidates whose child is one of this frame's branches are always
//     kept (TF's deactivate-child reset).  Lane 0 re-evaluates every kept candidate with the current state, exactly like
//     the sequential code; chunks are aligned to branches so a branch's entry test is made once.
// The node pool is small (compacted often); if it ever overflows the launcher falls back to beam_kernel.
constexpr int BEAM_WARPS = 4;
typedef int16_t BeamIdx;             // trie index type of the shared-memory kernels (pool, T and W all < 32768)
typedef CbBeamWorkT<BeamIdx> BeamWorkS;
typedef CbBeamNodeT<BeamIdx> BeamNodeS;

// `lg` is the window's logits [len][C] in shared memory (staged by the caller) or in global memory; either way the row of
// frame t + 1 is fetched into registers while frame t is processed, so its latency hides behind the frame's work.
__device__ int beam_decode_warp(const float* __restrict__ lg, int len, int C, int W, BeamWorkS k, int8_t* out, int lane,
                                float* score) {
    const unsigned FULL = 0xffffffffu;
    const int blank = C - 1, n_child = C - 1;
    int n_nodes = 1, n_leaves = 1, n_free = 0, err = 0;
    if (lane == 0) {
        k.nodes[0].parent = -1; k.nodes[0].label = -1; k.nodes[0].slot = 0; k.nodes[0].bidx = 0; k.nodes[0].bframe = -1;
        for (int c = 0; c < CB_BEAM_MAX_CHILD; ++c) k.nodes[0].child[c] = -1;
        k.slot_node[0] = 0;
        k.ot[0] = k.ob[0] = -INFINITY;
        k.nt[0] = 0.f; k.nb[0] = 0.f; k.nl[0] = -INFINITY;
        k.leaves[0] = 0;
        for (int s = W - 1; s >= 1; --s) k.freel[n_free++] = s;            // pop order 1,2,3,...
    }
    n_free = __shfl_sync(FULL, n_free, 0);
    __syncwarp();
    const int bpc = 32 / n_child;                        // branches per chunk of the extension loop
    // candidate number -> (branch, child) without an integer division when n_child is a power of two (4 for ACGT + blank:
    // the software division was ~25 of the ~60 instructions every surviving candidate cost)
    const int cshift = (n_child & (n_child - 1)) == 0 ? __ffs(n_child) - 1 : -1;
    int n_evict = 0;                                     // evictions so far in this window (warp-uniform)
    float inp[8], row[8];
#pragma unroll
    for (int c = 0; c < 8; ++c) row[c] = (c < C && len > 0) ? lg[c] : 0.f;
    for (int t = 0; t < len; ++t) {
        float mx = row[0];
#pragma unroll
        for (int c = 1; c < 8; ++c) if (c < C && row[c] > mx) mx = row[c];
#pragma unroll
        for (int c = 0; c < 8; ++c) inp[c] = row[c] - mx;
        if (t + 1 < len) {                               // prefetch the next frame's row
            const float* nx = lg + (size_t)(t + 1) * C;
#pragma unroll
            for (int c = 0; c < 8; ++c) if (c < C) row[c] = nx[c];
        }
        const int nb = n_leaves;
        // leaves_.Extract(): descending newp.total, stable  ==  rank of every leaf
        for (int i = lane; i < nb; i += 32) {
            const int v = k.leaves[i];
            const float key = k.nt[v];
            int rank = 0;
            for (int j = 0; j < nb; ++j) {
                const float o = k.nt[k.leaves[j]];
                rank += (o > key) || (o == key && j < i);
            }
            k.branches[rank] = v;
        }
        __syncwarp();
        for (int i = lane; i < nb; i += 32) {
            const int s = k.branches[i];
            k.ot[s] = k.nt[s]; k.ob[s] = k.nb[s];
            k.bnode[i] = k.slot_node[s]; k.bo_total[i] = k.nt[s]; k.bo_blank[i] = k.nb[s];
            k.nodes[k.slot_node[s]].bidx = i; k.nodes[k.slot_node[s]].bframe = t;
        }
        __syncwarp();
        for (int i = lane; i < nb; i += 32) {
            const int s = k.branches[i];
            const BeamNodeS& nd = k.nodes[k.slot_node[s]];
            float nl = k.nl[s];
            if (nd.parent >= 0) {
                const BeamNodeS& pa = k.nodes[nd.parent];
                if (pa.slot >= 0) {
                    const float prev = (nd.label == pa.label) ? k.ob[pa.slot] : k.ot[pa.slot];
                    nl = cb_lse(nl, prev);
                }
                nl += inp[nd.label];
            }
            const float nbv = k.ot[s] + inp[blank];
            k.nl[s] = nl; k.nb[s] = nbv;
            k.nt[s] = cb_lse(nbv, nl);
            k.leaves[i] = s;
        }
        __syncwarp();
        n_leaves = nb;
        // bottom = first minimum in push order
        float bot_val = INFINITY; int bot = 0x7fffffff;
        for (int i = lane; i < n_leaves; i += 32) {
            const float v = k.nt[k.leaves[i]];
            if (v < bot_val || bot == 0x7fffffff) { bot_val = v; bot = i; }     // strided scan keeps the lowest index per value
        }
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) {
            const float ov = __shfl_xor_sync(FULL, bot_val, off);
            const int oi = __shfl_xor_sync(FULL, bot, off);
            if (oi != 0x7fffffff && (bot == 0x7fffffff || ov < bot_val || (ov == bot_val && oi < bot))) { bot_val = ov; bot = oi; }
        }
        // QUIET FRAMES.  One branch per lane, its children unrolled (the loads of a branch are made once, the child loads are
        // independent): is there any candidate in the whole frame that survives the filter and is not idle (see below)?  The
        // test uses the state the first chunk would see, and as long as nothing is processed nothing changes, so every chunk
        // would come to the same conclusion: no candidate to look at.  On real logits ~85 % of the frames are like that, and
        // the four chunk filters -- eight dependent shared-memory loads each -- were a fifth of the kernel.
        {
            bool live = false;
            for (int i = lane; i < nb; i += 32) {
                const float tot = k.bo_total[i];
                if (tot > -INFINITY && (n_leaves < W || tot > bot_val)) {
                    const int bn = k.bnode[i];
                    const int label = k.nodes[bn].label;
                    const float blank_prev = k.bo_blank[i];
#pragma unroll
                    for (int c = 0; c < CB_BEAM_MAX_CHILD; ++c) {
                        if (c < n_child) {
                            const int ch = k.nodes[bn].child[c];
                            const float lab = inp[c] + (c == label ? blank_prev : tot);
                            const bool flag = (lab > -INFINITY && (n_leaves < W || lab > bot_val)) || (ch >= 0 && k.nodes[ch].bframe == t);
                            live |= flag && !(ch >= 0 && k.nodes[ch].slot >= 0);
                        }
                    }
                }
            }
            if (!__any_sync(FULL, live)) continue;
        }
        // extension loop, chunks of bpc whole branches
        for (int ib = 0; ib < nb; ib += bpc) {
            bool flag = false, idle = false;
            {
                const int i = ib + (cshift >= 0 ? lane >> cshift : lane / n_child);
                const int c = cshift >= 0 ? lane & (n_child - 1) : lane % n_child;
                if (lane < bpc * n_child && i < nb) {
                    const float tot = k.bo_total[i];
                    if (tot > -INFINITY && (n_leaves < W || tot > bot_val)) {
                        const int bn = k.bnode[i];
                        const int ch = k.nodes[bn].child[c];
                        const float prev = (c == k.nodes[bn].label) ? k.bo_blank[i] : tot;
                        const float lab = inp[c] + prev;
                        flag = (lab > -INFINITY && (n_leaves < W || lab > bot_val)) || (ch >= 0 && k.nodes[ch].bframe == t);
                        // A candidate whose child is an active beam is skipped by the sequential code without any effect
                        // ("already an active beam"), and a beam only stops being active by an eviction: as long as no
                        // eviction has happened since this test, such a candidate needs no second look at all.  On real
                        // logits that is most of what survives the filter (the likely next base of every beam already
                        // exists as its child): ~13 survivors per frame, a new base only every ~23 frames.
                        idle = flag && ch >= 0 && k.nodes[ch].slot >= 0;
                    }
                }
            }
            unsigned mask = __ballot_sync(FULL, flag);
            const unsigned idle_mask = __ballot_sync(FULL, idle);
            const int evict0 = n_evict;
            if (mask && !err) {
                // The surviving candidates are taken in order -- an insertion moves the beam's bottom, which the next
                // candidate's test reads -- but by the WHOLE warp: every lane follows the same (warp-uniform) decisions from
                // broadcast shared-memory reads, lane 0 does the scalar writes, and the two O(W) steps of an insertion (closing
                // the gap of the evicted leaf, finding the new bottom) are spread over the lanes.  On one lane they were
                // ~3,000 clocks per insertion, and a new base is ~W insertions.
                int cur_i = -1; bool cur_pass = false; float tot = 0.f;
                if ((mask & ~idle_mask) == 0) mask = 0;         // only idle survivors: nothing can be inserted, so nothing evicted
                while (mask) {
                    const int bit = __ffs(mask) - 1;
                    mask &= mask - 1;
                    if (((idle_mask >> bit) & 1u) && n_evict == evict0) continue;
                    const int i = ib + (cshift >= 0 ? bit >> cshift : bit / n_child);
                    const int c = cshift >= 0 ? bit & (n_child - 1) : bit % n_child;
                    if (i != cur_i) {                      // the branch's entry test, with the state of this moment
                        cur_i = i;
                        tot = k.bo_total[i];
                        cur_pass = tot > -INFINITY && (n_leaves < W || tot > bot_val);
                    }
                    if (!cur_pass) continue;
                    const int bn = k.bnode[i];
                    const int ch = k.nodes[bn].child[c];
                    if (ch >= 0 && k.nodes[ch].slot >= 0) continue;           // already an active beam
                    const float prev = (c == k.nodes[bn].label) ? k.bo_blank[i] : tot;
                    const float lab = inp[c] + prev;
                    if (!(lab > -INFINITY && (n_leaves < W || lab > bot_val))) {
                        if (ch >= 0 && k.nodes[ch].bframe == t) {
                            const int bi = k.nodes[ch].bidx;
                            __syncwarp();
                            if (lane == 0) { k.bo_total[bi] = -INFINITY; k.bo_blank[bi] = -INFINITY; }
                            __syncwarp();
                        }
                        continue;
                    }
                    __syncwarp();          // every lane has made this candidate's decision: nothing it read may change before
                    if (n_leaves == W) {                                       // evict the bottom beam
                        const int bs = k.leaves[bot];
                        const int bnode = k.slot_node[bs];
                        __syncwarp();
                        for (int base = bot; base + 1 < n_leaves; base += 32) {      // close the gap, 32 leaves at a time
                            const int q = base + lane;
                            const int v = (q + 1 < n_leaves) ? k.leaves[q + 1] : 0;
                            __syncwarp();
                            if (q + 1 < n_leaves) k.leaves[q] = v;
                            __syncwarp();
                        }
                        if (lane == 0) { k.nodes[bnode].slot = -1; k.freel[n_free] = bs; }
                        --n_leaves;
                        ++n_free;
                        ++n_evict;
                        __syncwarp();
                    }
                    int node = ch;
                    if (node < 0) {
                        if (n_nodes == k.pool) {
                            __syncwarp();
                            int m = 0;
                            if (lane == 0) m = cb_beam_compact(k, n_nodes, n_leaves, nb, n_child);
                            n_nodes = __shfl_sync(FULL, m, 0);
                            __syncwarp();
                            if (n_nodes == k.pool) { err = 1; break; }
                        }
                        node = n_nodes++;
                        const int parent = k.bnode[i];                         // (after a compaction: the renumbered node)
                        if (lane == 0) {
                            BeamNodeS& nn = k.nodes[node];
                            nn.parent = (BeamIdx)parent; nn.label = (BeamIdx)c; nn.slot = -1; nn.bidx = 0; nn.bframe = -1;
                            for (int q = 0; q < CB_BEAM_MAX_CHILD; ++q) nn.child[q] = -1;
                            k.nodes[parent].child[c] = (BeamIdx)node;
                        }
                    }
                    const int s = k.freel[n_free - 1];
                    --n_free;
                    if (lane == 0) {
                        k.slot_node[s] = node;
                        k.nodes[node].slot = (BeamIdx)s;
                        k.nb[s] = -INFINITY; k.nl[s] = lab; k.nt[s] = lab;
                        k.ot[s] = k.ob[s] = -INFINITY;
                        k.leaves[n_leaves] = s;
                    }
                    ++n_leaves;
                    __syncwarp();
                    // new bottom = first minimum in push order
                    float bv = INFINITY; int bq = 0x7fffffff;
                    for (int q = lane; q < n_leaves; q += 32) {
                        const float v = k.nt[k.leaves[q]];
                        if (bq == 0x7fffffff || v < bv) { bv = v; bq = q; }  // ascending q per lane: keeps its lowest index per value
                    }
#pragma unroll
                    for (int off = 16; off > 0; off >>= 1) {
                        const float ov = __shfl_xor_sync(FULL, bv, off);
                        const int oq = __shfl_xor_sync(FULL, bq, off);
                        if (oq != 0x7fffffff && (bq == 0x7fffffff || ov < bv || (ov == bv && oq < bq))) { bv = ov; bq = oq; }
                    }
                    bot = bq; bot_val = bv;
                }
            }
            __syncwarp();          // (n_leaves, n_free, n_nodes, bot, bot_val, err are warp-uniform: every lane kept them)
        }
        if (err) return -2;
    }
    int n = 0;
    if (lane == 0) {
        int best = 0;
        for (int i = 1; i < n_leaves; ++i) if (k.nt[k.leaves[i]] > k.nt[k.leaves[best]]) best = i;
        if (score) *score = k.nt[k.leaves[best]];
        for (int cur = k.slot_node[k.leaves[best]]; k.nodes[cur].parent >= 0; cur = k.nodes[cur].parent) ++n;
        int i = n - 1;
        for (int cur = k.slot_node[k.leaves[best]]; k.nodes[cur].parent >= 0; cur = k.nodes[cur].parent)
            out[i--] = (int8_t)k.nodes[cur].label;
    }
    return __shfl_sync(FULL, n, 0);
}

// STAGED: the window's logits are copied to shared memory first (T*C*4 bytes per window).  Not staged: the search reads
// its one row per frame straight from global memory (prefetched a frame ahead), shared memory holds the workspace only, so
// more CTAs are resident per SM -- and the search is a latency chain per warp, so resident warps are its throughput.
// A window whose pool overflows is reported as n_bases = -1 (and *marked is raised) for beam_retry_kernel.
template <bool STAGED>
__global__ void __launch_bounds__(BEAM_WARPS * 32) beam_warp_kernel(const float* __restrict__ logits, const int32_t* __restrict__ lens,
                                                                    int B, int T, int C, int W, int pool, int stride,
                                                                    int8_t* __restrict__ bases, int32_t* __restrict__ n_bases,
                                                                    int* __restrict__ marked, float* __restrict__ scores) {
#ifdef CB_HOST_EMU
    char* beam_sm = reinterpret_cast<char*>(emu::dyn_smem());
#else
    extern __shared__ __align__(16) char beam_sm[];
#endif
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int b = blockIdx.x * BEAM_WARPS + w;
    if (b >= B) return;
    char* base = beam_sm + (size_t)w * stride;
    int len = lens[b];
    len = len < 0 ? 0 : (len > T ? T : len);
    const float* src = logits + (size_t)b * T * C;
    const float* lg = src;
    size_t work_off = 0;
    if (STAGED) {
        float* stage = reinterpret_cast<float*>(base);
        for (int i = lane; i < len * C; i += 32) stage[i] = src[i];
        __syncwarp();
        lg = stage;
        work_off = ((size_t)T * C * 4 + 15) & ~(size_t)15;
    }
    int8_t* dst = bases + (size_t)b * T;
    BeamWorkS k = cb_beam_work_carve<BeamIdx>(base + work_off, W, pool);
    int n = beam_decode_warp(lg, len, C, W, k, dst, lane, scores ? scores + b : nullptr);
    if (n < 0) { if (lane == 0) atomicExch(marked, 1); n = -1; }
    __syncwarp();
    for (int i = (n < 0 ? 0 : n) + lane; i < T; i += 32) dst[i] = 0;
    if (lane == 0) n_bases[b] = n;
}

// Second pass for the windows the first pass marked (n_bases == -1): the same cooperative search, one window per CTA with
// the CTA's whole shared-memory allotment as its node pool (~260 W nodes at W=30), logits from global memory.  A window that
// overflows even this pool stays marked (for beam_kernel's slot mode) and raises *marked.
__global__ void __launch_bounds__(32) beam_retry_kernel(const float* __restrict__ logits, const int32_t* __restrict__ lens, int B, int T,
                                                        int C, int W, int pool, int8_t* __restrict__ bases,
                                                        int32_t* __restrict__ n_bases, int* __restrict__ marked,
                                                        float* __restrict__ scores) {
#ifdef CB_HOST_EMU
    char* beam_sm = reinterpret_cast<char*>(emu::dyn_smem());
#else
    extern __shared__ __align__(16) char beam_sm[];
#endif
    const int b = blockIdx.x, lane = threadIdx.x;
    if (b >= B || n_bases[b] != -1) return;
    int len = lens[b];
    len = len < 0 ? 0 : (len > T ? T : len);
    int8_t* dst = bases + (size_t)b * T;
    BeamWorkS k = cb_beam_work_carve<BeamIdx>(beam_sm, W, pool);
    int n = beam_decode_warp(logits + (size_t)b * T * C, len, C, W, k, dst, lane, scores ? scores + b : nullptr);
    if (n < 0) { if (lane == 0) atomicExch(marked, 1); n = -1; }
    __syncwarp();
    for (int i = (n < 0 ? 0 : n) + lane; i < T; i += 32) dst[i] = 0;
    if (lane == 0) n_bases[b] = n;
}

// ---- assembly ----------------------------------------------------------------------------------------------------
struct AsmWork {
    int* list;       // [n_windows] indices of non-empty windows, in order
    int* n_ne;       // [1]
    int* disp;       // [n_windows] displacement of list[j] against list[j-1]; later the running position
    int* length;     // [1] consensus length before clamping to max_len
    int* counts;     // [4][max_len]
    double* qsum;    // [4][max_len]
    double* logfact; // [T+2]
    int* scratch;    // [n_windows][scratch_stride]
    size_t scratch_stride;
};

__global__ void __launch_bounds__(1024) asm_compact_kernel(const int32_t* __restrict__ n_bases, int n_windows, int T,
                                                           AsmWork w, int32_t* __restrict__ pos) {
    // single block: ordered compaction of the non-empty windows + the log-factorial table
    __shared__ int warp_tot[32];
    __shared__ int base;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) base = 0;
    __syncthreads();
    for (int start = 0; start < n_windows; start += blockDim.x) {
        const int i = start + tid;
        const int keep = (i < n_windows && n_bases[i] > 0) ? 1 : 0;
        const unsigned m = __ballot_sync(0xffffffffu, keep);
        if (lane == 0) warp_tot[wid] = __popc(m);
        __syncthreads();
        int off = base;
        for (int q = 0; q < wid; ++q) off += warp_tot[q];
        if (keep) w.list[off + __popc(m & ((1u << lane) - 1u))] = i;
        if (i < n_windows && !keep) pos[i] = -1;
        __syncthreads();
        if (tid == 0) { int tot = 0; for (int q = 0; q < (int)(blockDim.x >> 5); ++q) tot += warp_tot[q]; base += tot; }
        __syncthreads();
    }
    if (tid == 0) {
        *w.n_ne = base;
        double acc = 0.0;                  // sum([np.log(x+1) for x in range(k)]) accumulated left to right
        w.logfact[0] = 0.0;
        for (int k = 1; k <= T + 1; ++k) { acc += log((double)k); w.logfact[k] = acc; }
    }
}

__global__ void __launch_bounds__(128) asm_disp_kernel(const int8_t* __restrict__ bases, const int32_t* __restrict__ n_bases,
                                                       int T, int kernel, double jsr, AsmWork w) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int n = *w.n_ne;
    if (j >= n) return;
    if (j == 0) { w.disp[0] = 0; return; }
    const int wc = w.list[j], wp = w.list[j - 1];
    const int8_t* cur = bases + (size_t)wc * T;
    const int8_t* prev = bases + (size_t)wp * T;
    const int la = n_bases[wc], lb = n_bases[wp];
    int d;
    if (kernel == CB_ASM_STICK) d = cb_disp_stick(la, lb);
    else if (kernel == CB_ASM_GLUE) d = cb_disp_glue(cur, la, prev, lb);
    else d = cb_disp_simple(cur, la, prev, lb, jsr, w.logfact, w.scratch + (size_t)j * w.scratch_stride);
    w.disp[j] = d;
}

__global__ void __launch_bounds__(1024) asm_scan_kernel(const int32_t* __restrict__ n_bases, AsmWork w,
                                                        int32_t* __restrict__ pos, int32_t* __restrict__ out_len,
                                                        int max_len) {
    // single block: inclusive scan of the displacements -> window coordinates; length = max_{j>=1}(pos_j + len_j)
    __shared__ int warp_tot[32];
    __shared__ int base, max_end;
    const int tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    const int n = *w.n_ne;
    if (tid == 0) { base = 0; max_end = 0; }
    __syncthreads();
    for (int start = 0; start < n; start += blockDim.x) {
        const int j = start + tid;
        int v = j < n ? w.disp[j] : 0;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(0xffffffffu, v, o); if (lane >= o) v += t; }
        if (lane == 31) warp_tot[wid] = v;
        __syncthreads();
        int off = base;
        for (int q = 0; q < wid; ++q) off += warp_tot[q];
        v += off;
        if (j < n) {
            w.disp[j] = v;                               // running position of window list[j]
            pos[w.list[j]] = v;
            if (j >= 1) atomicMax(&max_end, v + n_bases[w.list[j]]);   // window 0 never updates `length` (:316-318)
        }
        __syncthreads();
        if (tid == blockDim.x - 1) base = v;
        __syncthreads();
    }
    if (tid == 0) { *w.length = max_end; *out_len = max_end < max_len ? max_end : max_len; }
}

__global__ void __launch_bounds__(128) asm_vote_kernel(const int8_t* __restrict__ bases, const int32_t* __restrict__ n_bases,
                                                       const float* __restrict__ path_prob, int T, AsmWork w, int max_len) {
    // one block per non-empty window: add_count(_qs) (easy_assembler.py:381-388,435-442)
    const int j = blockIdx.x;
    if (j >= *w.n_ne) return;
    const int wi = w.list[j];
    const int len = n_bases[wi], p0 = w.disp[j];
    int length = *w.length; if (length > max_len) length = max_len;
    const double q = path_prob ? (double)path_prob[wi] : 0.0;
    for (int i = threadIdx.x; i < len; i += blockDim.x) {
        const int col = p0 + i;
        if (col < 0 || col >= length) continue;          // negative start trims the head; beyond `length` is cut
        const int base = bases[(size_t)wi * T + i] & 3;
        atomicAdd(&w.counts[(size_t)base * max_len + col], 1);
        if (path_prob) atomicAdd(&w.qsum[(size_t)base * max_len + col], q);
    }
}

__global__ void __launch_bounds__(256) asm_finish_kernel(AsmWork w, int max_len, int8_t* __restrict__ consensus,
                                                         char* __restrict__ qual) {
    int length = *w.length; if (length > max_len) length = max_len;
    const int col = blockIdx.x * blockDim.x + threadIdx.x;
    if (col >= length) return;
    int c[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) c[r] = w.counts[(size_t)r * max_len + col];
    int best = 0;
#pragma unroll
    for (int r = 1; r < 4; ++r) if (c[r] > c[best]) best = r;        // np.argmax: first maximum
    consensus[col] = (int8_t)best;
    if (!qual) return;
    // qs(): np.argsort(axis=0) on 4 elements is a stable insertion sort; rows [2] and [3] of the sorted matrix
    int idx[4] = {0, 1, 2, 3};
#pragma unroll
    for (int x = 1; x < 4; ++x) {
        const int v = idx[x];
        int y = x;
        while (y > 0 && c[idx[y - 1]] > c[v]) { idx[y] = idx[y - 1]; --y; }
        idx[y] = v;
    }
    const double c3 = (double)c[idx[3]], c2 = (double)c[idx[2]];
    const double qs3 = w.qsum[(size_t)idx[3] * max_len + col];
    char ch = '!';
    if (c3 > 0.0) {
        const double q = 10.0 * log10((c3 + 1.0) / (c2 + 1.0)) + qs3 / c3 / log(10.0);
        ch = (char)((int)q + 33);                                     // astype(int): truncation toward zero
    }
    qual[col] = ch;
}


// Workspace of one cb_assemble call: byte offsets of the AsmWork arrays inside one allocation (shared by the launcher in
// cb_seq.cu and the host emulation).
struct AsmPlan { size_t list, nne, disp, len, counts, qsum, logfact, scratch, total, scratch_stride, ml; };
inline AsmPlan asm_plan(int n_windows, int T, int kernel, int max_len) {
    AsmPlan p;
    p.scratch_stride = kernel == CB_ASM_SIMPLE ? align_up(cb_simple_scratch_ints(T, T), 4) : 0;
    p.ml = (size_t)(max_len > 0 ? max_len : 1);
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += align_up(bytes, 256); return o; };
    p.list = carve(sizeof(int) * n_windows); p.nne = carve(sizeof(int)); p.disp = carve(sizeof(int) * n_windows);
    p.len = carve(sizeof(int)); p.counts = carve(sizeof(int) * 4 * p.ml); p.qsum = carve(sizeof(double) * 4 * p.ml);
    p.logfact = carve(sizeof(double) * (T + 2)); p.scratch = carve(sizeof(int) * p.scratch_stride * n_windows);
    p.total = off;
    return p;
}
inline AsmWork asm_work(char* base, const AsmPlan& p) {
    AsmWork w;
    w.list = (int*)(base + p.list); w.n_ne = (int*)(base + p.nne); w.disp = (int*)(base + p.disp);
    w.length = (int*)(base + p.len); w.counts = (int*)(base + p.counts); w.qsum = (double*)(base + p.qsum);
    w.logfact = (double*)(base + p.logfact); w.scratch = (int*)(base + p.scratch); w.scratch_stride = p.scratch_stride;
    return w;
}

// Shared-memory footprint of one window of beam_warp_kernel ((staged logits +) workspace) for a pool of `pool` nodes.
inline size_t beam_warp_stride(int T, int C, int W, int pool, bool staged) {
    return align_up((staged ? align_up((size_t)T * C * 4, 16) : 0) + cb_beam_work_bytes<BeamIdx>(W, pool), 16);
}

// Node pool of the first pass.  Live nodes are the ancestors of the W leaves and of the current branches.  On real logits
// (the reference's chiron/utils/logits_sample.npy: 1100 windows, T=300, W=30) 55 % of the windows fit 6W nodes, 99.3 % fit
// 12W and all fit 16W; of 64 oracle-basecalled T=512 windows of the bundled read3, 27 % fit 6W and all fit 16W
// (profiles/r01_beam_pool_survey_*.json).  The pool is therefore 16W nodes -- the rare window that still overflows is redone
// alone by beam_retry_kernel, not the batch -- grown into whatever shared memory the resulting CTAs-per-SM leaves unused
// (nodes are free up to the next occupancy step) and bounded by what four windows hold in 192 KB and by the
// never-overflows bound 2W(T+1)+2.  At W=30 that is 5 CTAs = 20 warps per SM (26 bytes per node with 16-bit indices).
constexpr long long BEAM_SMEM_BUDGET = 192 * 1024;          // of the 200 KB the launcher opts in to
inline long long beam_small_pool(int T, int W, int mult = 16) {
    const long long cap = 2LL * W * (T + 1) + 2;
    const long long node = (long long)(sizeof(BeamNodeS) + sizeof(BeamIdx)), fixed = 12LL * 4 * W + 32;
    long long pool = (long long)mult * W;
    const long long fit = (BEAM_SMEM_BUDGET / BEAM_WARPS - fixed) / node;
    if (pool > fit) pool = fit;
    if (pool < 64 && fit >= 64) pool = 64;
    // nodes are free up to the next occupancy step: grow the pool into the slack of the CTAs-per-SM it already costs
    const long long ctas = (227LL * 1024) / ((node * pool + fixed) * BEAM_WARPS + 1024);
    if (ctas >= 1) {
        const long long grown = (((227LL * 1024) / ctas - 1024) / BEAM_WARPS - fixed) / node;
        if (grown > pool) pool = grown < fit ? grown : fit;
    }
    if (pool > cap) pool = cap;
    if (pool > 32767) pool = 32767;
    return pool;                                   // usable iff >= 2W + 2
}

// Pool of beam_retry_kernel: what one window can hold in the 200 KB the launcher opts in to, at most the never-overflows bound.
inline long long beam_retry_pool(int T, int W) {
    const long long cap = 2LL * W * (T + 1) + 2;
    long long fit = (200LL * 1024 - 12LL * 4 * W - 32) / (long long)(sizeof(BeamNodeS) + sizeof(BeamIdx));
    if (fit > 32767) fit = 32767;
    return fit < cap ? fit : cap;
}

}  // namespace cb_seq
