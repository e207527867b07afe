// Sequential decode/assembly algorithms shared by the device kernels in cb_seq.cu and by the host-compiled
// self-test hooks (include/chiron_b200_selftest.h).  Everything here is __host__ __device__ so the exact code the GPU
// runs can be unit-tested on a machine without a GPU; the product path only ever calls it from kernels.
//
//  * cb_beam_decode_one   tf.nn.ctc_beam_search_decoder(merge_repeated=False, top_paths=1), i.e. TF 1.15's
//                         CTCBeamSearchDecoder::Step/TopPaths (chiron/chiron_eval.py:489-492)
//  * cb_disp_glue/stick   glue_kernal / stick_kernal               chiron/utils/easy_assembler.py:276-300
//  * cb_disp_simple       simple_assembly_kernal incl. difflib.SequenceMatcher.get_matching_blocks with autojunk
//                         chiron/utils/easy_assembler.py:212-250
#pragma once
#include <math.h>
#include <stdint.h>

#ifdef __CUDACC__
#define CB_HD __host__ __device__
#else
#define CB_HD
#endif

#define CB_BEAM_MAX_CHILD 7          // n_class - 1 <= 7

// ---------------------------------------------------------------------------------------------------------------------
// CTC beam search.  The prefix trie lives in a bounded node pool; probabilities live in W "leaf slots" because only
// beams that are currently leaves are active (TF resets newp of everything it drops).  Children are materialised only
// when they enter the beam, which is equivalent to TF's eager PopulateChildren because an inactive child carries no
// state.  When the pool fills up it is compacted in place (ancestors of leaves and of the current branches survive).
// ---------------------------------------------------------------------------------------------------------------------
// Index type I: int for the thread-per-window kernel over a global workspace (pools up to 2*W*(T+1)+2 nodes), int16_t for
// the shared-memory kernels (node pools of a few thousand nodes, T < 32768): a 24-byte node instead of 48 doubles the
// windows an SM can keep resident, and the search is a latency chain per warp.
template <typename I>
struct CbBeamNodeT {
    I parent;
    I label;
    I slot;                            // leaf slot, -1 = inactive
    I bidx, bframe;                    // index in `branches` of frame `bframe` (to reach the per-branch oldp copy)
    I child[CB_BEAM_MAX_CHILD];
};
typedef CbBeamNodeT<int> CbBeamNode;

template <typename I>
struct CbBeamWorkT {                   // per-window scratch (global memory, or shared memory in the cooperative kernels)
    CbBeamNodeT<I>* nodes; int pool;   // pool >= 2*W + 2
    I* remap;                          // [pool]   compaction scratch
    int *slot_node;                    // [W]
    float *ot, *ob, *nt, *nb, *nl;     // [W] oldp.total/blank, newp.total/blank/label per slot
    int *leaves, *branches, *freel;    // [W]
    int* bnode; float *bo_total, *bo_blank;   // [W] per-branch copies of oldp (slots may be recycled mid-step)
};
typedef CbBeamWorkT<int> CbBeamWork;

template <typename I = int>
CB_HD inline size_t cb_beam_work_bytes(int W, int pool) {
    const size_t nodes = (sizeof(CbBeamNodeT<I>) + sizeof(I)) * (size_t)pool;
    return ((nodes + 3) & ~(size_t)3) + 12 * sizeof(int) * (size_t)W;
}

template <typename I = int>
CB_HD inline CbBeamWorkT<I> cb_beam_work_carve(void* base, int W, int pool) {
    CbBeamWorkT<I> k;
    char* p = (char*)base;
    k.nodes = (CbBeamNodeT<I>*)p; p += sizeof(CbBeamNodeT<I>) * (size_t)pool;
    k.pool = pool;
    k.remap = (I*)p; p += sizeof(I) * (size_t)pool;
    p = (char*)base + (((size_t)(p - (char*)base) + 3) & ~(size_t)3);
    int* q = (int*)p;
    k.slot_node = q; q += W;
    k.ot = (float*)q; q += W; k.ob = (float*)q; q += W;
    k.nt = (float*)q; q += W; k.nb = (float*)q; q += W; k.nl = (float*)q; q += W;
    k.leaves = q; q += W; k.branches = q; q += W; k.freel = q; q += W;
    k.bnode = q; q += W; k.bo_total = (float*)q; q += W; k.bo_blank = (float*)q; q += W;
    return k;
}

CB_HD inline float cb_lse(float a, float b) {
    if (a == -INFINITY && b == -INFINITY) return -INFINITY;
    return a > b ? a + log1pf(expf(b - a)) : b + log1pf(expf(a - b));
}

// Compact the pool: keep ancestors of every leaf and of every current branch.  Returns the new node count.
template <typename I>
CB_HD inline int cb_beam_compact(CbBeamWorkT<I>& k, int n_nodes, int n_leaves, int n_branches, int n_child) {
    I* mark = k.remap;
    for (int i = 0; i < n_nodes; ++i) mark[i] = 0;
    for (int i = 0; i < n_leaves; ++i)
        for (int n = k.slot_node[k.leaves[i]]; n >= 0 && !mark[n]; n = k.nodes[n].parent) mark[n] = 1;
    for (int i = 0; i < n_branches; ++i)
        for (int n = k.bnode[i]; n >= 0 && !mark[n]; n = k.nodes[n].parent) mark[n] = 1;
    int m = 0;
    for (int i = 0; i < n_nodes; ++i) mark[i] = mark[i] ? (I)m++ : (I)-1;      // mark becomes the remap table
    for (int i = 0; i < n_nodes; ++i) {
        if (mark[i] < 0) continue;
        CbBeamNodeT<I> nd = k.nodes[i];
        nd.parent = nd.parent >= 0 ? mark[nd.parent] : (I)-1;         // parents of kept nodes are kept
        for (int c = 0; c < n_child; ++c) nd.child[c] = nd.child[c] >= 0 ? mark[nd.child[c]] : (I)-1;
        k.nodes[mark[i]] = nd;                                         // mark[i] <= i: in-place forward move is safe
    }
    for (int i = 0; i < n_leaves; ++i) k.slot_node[k.leaves[i]] = mark[k.slot_node[k.leaves[i]]];
    for (int i = 0; i < n_branches; ++i) k.bnode[i] = mark[k.bnode[i]];
    return m;
}

// logits: [T][C] row-major rows of one window; returns the number of decoded labels written to out, or -2 when the
// node pool is too small even after compaction.
// `score` (may be null): newp.total of the best beam, i.e. the path log probability TopPaths() reports -- the `log_prob`
// output of tf.nn.ctc_beam_search_decoder (chiron/export_test.py:36-40).
template <typename I>
CB_HD inline int cb_beam_decode_one(const float* logits, int len, int C, int W, CbBeamWorkT<I> k, int8_t* out,
                                    float* score = nullptr) {
    const int blank = C - 1, n_child = C - 1;
    int n_nodes = 1, n_leaves = 1, n_free = 0;
    k.nodes[0].parent = -1; k.nodes[0].label = -1; k.nodes[0].slot = 0; k.nodes[0].bidx = 0; k.nodes[0].bframe = -1;
    for (int c = 0; c < CB_BEAM_MAX_CHILD; ++c) k.nodes[0].child[c] = -1;
    k.slot_node[0] = 0;
    k.ot[0] = k.ob[0] = -INFINITY;
    k.nt[0] = 0.f; k.nb[0] = 0.f; k.nl[0] = -INFINITY;
    k.leaves[0] = 0;
    for (int s = W - 1; s >= 1; --s) k.freel[n_free++] = s;            // pop order 1,2,3,...
    float inp[8];
    for (int t = 0; t < len; ++t) {
        const float* row = logits + (size_t)t * C;
        float mx = row[0];
        for (int c = 1; c < C; ++c) if (row[c] > mx) mx = row[c];
        for (int c = 0; c < C; ++c) inp[c] = row[c] - mx;
        // leaves_.Extract(): descending newp.total, stable
        const int nb = n_leaves;
        for (int i = 0; i < nb; ++i) {
            const int v = k.leaves[i];
            int j = i;
            while (j > 0 && k.nt[k.branches[j - 1]] < k.nt[v]) { k.branches[j] = k.branches[j - 1]; --j; }
            k.branches[j] = v;
        }
        for (int i = 0; i < nb; ++i) {
            const int s = k.branches[i];
            k.ot[s] = k.nt[s]; k.ob[s] = k.nb[s];
            k.bnode[i] = k.slot_node[s]; k.bo_total[i] = k.nt[s]; k.bo_blank[i] = k.nb[s];
            k.nodes[k.slot_node[s]].bidx = i; k.nodes[k.slot_node[s]].bframe = t;
        }
        for (int i = 0; i < nb; ++i) {
            const int s = k.branches[i];
            const CbBeamNodeT<I>& nd = k.nodes[k.slot_node[s]];
            if (nd.parent >= 0) {
                const CbBeamNodeT<I>& pa = k.nodes[nd.parent];
                if (pa.slot >= 0) {
                    const float prev = (nd.label == pa.label) ? k.ob[pa.slot] : k.ot[pa.slot];
                    k.nl[s] = cb_lse(k.nl[s], prev);
                }
                k.nl[s] += inp[nd.label];
            }
            k.nb[s] = k.ot[s] + inp[blank];
            k.nt[s] = cb_lse(k.nb[s], k.nl[s]);
            k.leaves[i] = s;
        }
        n_leaves = nb;
        // bottom = first minimum in push order
        int bot = 0;
        for (int i = 1; i < n_leaves; ++i) if (k.nt[k.leaves[i]] < k.nt[k.leaves[bot]]) bot = i;
        float bot_val = k.nt[k.leaves[bot]];
        for (int i = 0; i < nb; ++i) {
            const float tot = k.bo_total[i];
            if (!(tot > -INFINITY && (n_leaves < W || tot > bot_val))) continue;
            for (int c = 0; c < n_child; ++c) {
                const int bn = k.bnode[i];
                const int ch = k.nodes[bn].child[c];
                if (ch >= 0 && k.nodes[ch].slot >= 0) continue;           // already an active beam
                const float prev = (c == k.nodes[bn].label) ? k.bo_blank[i] : tot;
                const float lab = inp[c] + prev;
                if (!(lab > -INFINITY && (n_leaves < W || lab > bot_val))) {
                    // TF "deactivate child": c.oldp.Reset(); c.newp.Reset().  If the child is a branch of this very
                    // frame that was evicted a moment ago, its oldp is what the rest of this loop will read.
                    if (ch >= 0 && k.nodes[ch].bframe == t) {
                        k.bo_total[k.nodes[ch].bidx] = -INFINITY; k.bo_blank[k.nodes[ch].bidx] = -INFINITY;
                    }
                    continue;
                }
                if (n_leaves == W) {                                       // evict the bottom beam
                    const int bs = k.leaves[bot];
                    k.nodes[k.slot_node[bs]].slot = -1;
                    for (int q = bot; q + 1 < n_leaves; ++q) k.leaves[q] = k.leaves[q + 1];
                    --n_leaves;
                    k.freel[n_free++] = bs;
                }
                int node = ch;
                if (node < 0) {
                    if (n_nodes == k.pool) {
                        n_nodes = cb_beam_compact(k, n_nodes, n_leaves, nb, n_child);
                        if (n_nodes == k.pool) return -2;
                    }
                    node = n_nodes++;
                    CbBeamNodeT<I>& nn = k.nodes[node];
                    nn.parent = (I)k.bnode[i]; nn.label = (I)c; nn.slot = -1; nn.bidx = 0; nn.bframe = -1;
                    for (int q = 0; q < CB_BEAM_MAX_CHILD; ++q) nn.child[q] = -1;
                    k.nodes[k.bnode[i]].child[c] = node;
                }
                const int s = k.freel[--n_free];
                k.slot_node[s] = node;
                k.nodes[node].slot = s;
                k.nb[s] = -INFINITY; k.nl[s] = lab; k.nt[s] = lab;
                k.ot[s] = k.ob[s] = -INFINITY;
                k.leaves[n_leaves++] = s;
                bot = 0;
                for (int q = 1; q < n_leaves; ++q) if (k.nt[k.leaves[q]] < k.nt[k.leaves[bot]]) bot = q;
                bot_val = k.nt[k.leaves[bot]];
            }
        }
    }
    int best = 0;
    for (int i = 1; i < n_leaves; ++i) if (k.nt[k.leaves[i]] > k.nt[k.leaves[best]]) best = i;
    if (score) *score = k.nt[k.leaves[best]];
    int n = 0;
    for (int cur = k.slot_node[k.leaves[best]]; k.nodes[cur].parent >= 0; cur = k.nodes[cur].parent) ++n;
    int i = n - 1;
    for (int cur = k.slot_node[k.leaves[best]]; k.nodes[cur].parent >= 0; cur = k.nodes[cur].parent)
        out[i--] = (int8_t)k.nodes[cur].label;
    return n;
}

// ---------------------------------------------------------------------------------------------------------------------
// Assembly displacement kernels.  Strings are int8 base indices (0..3).
// ---------------------------------------------------------------------------------------------------------------------
CB_HD inline int cb_disp_stick(int /*la*/, int lb) { return lb; }

CB_HD inline int cb_disp_glue(const int8_t* cur, int la, const int8_t* prev, int lb) {
    int max_overlap = (int)floor(0.1 * (double)lb);
    if (la < max_overlap) max_overlap = la;
    int best_i = 0, best_score = 0;
    for (int i = 1; i < max_overlap; ++i) {
        int hits = 0;
        for (int q = 0; q < i; ++q) hits += cur[q] == prev[lb - i + q];
        const int score = 2 * hits - i;
        if (score > best_score) { best_score = score; best_i = i; }
    }
    return lb - best_i;
}

// scratch ints needed by cb_disp_simple for strings up to (la, lb)
CB_HD inline size_t cb_simple_scratch_ints(int la, int lb) {
    const int mn = la < lb ? la : lb;
    return (size_t)2 * (lb + 2) + (size_t)lb + 8 + (size_t)4 * (la + lb + 4) + (size_t)3 * (mn + 2) +
           (size_t)2 * (la + lb + 2);
}

// difflib.SequenceMatcher(a=cur, b=prev, autojunk=True).get_matching_blocks() -> offsets -> log-probabilities -> argmax.
// `logfact[k]` = sum_{x<k} log(x+1) accumulated left to right in double (k <= max(la, lb)).
CB_HD inline int cb_disp_simple(const int8_t* a, int la, const int8_t* b, int lb, double jump_step_ratio,
                                const double* logfact, int* scratch) {
    // ---- carve scratch ----
    int* j2a = scratch;                 // [lb+2]  (index j+1)
    int* j2b = j2a + (lb + 2);          // [lb+2]
    int* bpos = j2b + (lb + 2);         // [lb] positions of b grouped by base
    int* bstart = bpos + lb;            // [5]
    int* queue = bstart + 8;            // [(la+lb+4)][4]
    int* blocks = queue + 4 * (la + lb + 4);   // [(mn+2)][3]
    const int mn = la < lb ? la : lb;
    int* ns_key = blocks + 3 * (mn + 2);       // insertion-ordered keys
    int* ns_val = ns_key + (la + lb + 2);
    // ---- __chain_b: b2j + autojunk "popular" purge ----
    int cnt[4] = {0, 0, 0, 0};
    for (int j = 0; j < lb; ++j) cnt[b[j] & 3]++;
    bstart[0] = 0;
    for (int c = 0; c < 4; ++c) bstart[c + 1] = bstart[c] + cnt[c];
    int fill[4] = {bstart[0], bstart[1], bstart[2], bstart[3]};
    for (int j = 0; j < lb; ++j) bpos[fill[b[j] & 3]++] = j;
    bool popular[4] = {false, false, false, false};
    if (lb >= 200) {
        const int ntest = lb / 100 + 1;
        for (int c = 0; c < 4; ++c) popular[c] = cnt[c] > ntest;
    }
    for (int j = 0; j < lb + 2; ++j) { j2a[j] = 0; j2b[j] = 0; }
    // ---- get_matching_blocks ----
    int nq = 0, nblk = 0;
    queue[0] = 0; queue[1] = la; queue[2] = 0; queue[3] = lb; nq = 1;
    while (nq > 0) {
        --nq;
        const int alo = queue[4 * nq], ahi = queue[4 * nq + 1], blo = queue[4 * nq + 2], bhi = queue[4 * nq + 3];
        // find_longest_match(alo, ahi, blo, bhi)
        int besti = alo, bestj = blo, bestsize = 0;
        int* prevl = j2a; int* curl = j2b;       // prevl[j+1] = length of match ending at (i-1, j)
        int last_c = -1;
        for (int i = alo; i < ahi; ++i) {
            const int c = a[i] & 3;
            if (!popular[c]) {
                for (int q = bstart[c]; q < bstart[c + 1]; ++q) {
                    const int j = bpos[q];
                    if (j < blo) continue;
                    if (j >= bhi) break;
                    const int kk = prevl[j] + 1;          // prevl[(j-1)+1]
                    curl[j + 1] = kk;
                    if (kk > bestsize) { besti = i - kk + 1; bestj = j - kk + 1; bestsize = kk; }
                }
            }
            // j2len = newj2len: clear the row before the one just written, then swap
            if (last_c >= 0 && !popular[last_c])
                for (int q = bstart[last_c]; q < bstart[last_c + 1]; ++q) prevl[bpos[q] + 1] = 0;
            int* tmp = prevl; prevl = curl; curl = tmp;
            last_c = c;
        }
        if (last_c >= 0 && !popular[last_c])
            for (int q = bstart[last_c]; q < bstart[last_c + 1]; ++q) prevl[bpos[q] + 1] = 0;
        // extend by non-junk (popular) elements on each end; there is no junk (isjunk=None)
        while (besti > alo && bestj > blo && a[besti - 1] == b[bestj - 1]) { --besti; --bestj; ++bestsize; }
        while (besti + bestsize < ahi && bestj + bestsize < bhi && a[besti + bestsize] == b[bestj + bestsize]) ++bestsize;
        if (bestsize) {
            blocks[3 * nblk] = besti; blocks[3 * nblk + 1] = bestj; blocks[3 * nblk + 2] = bestsize; ++nblk;
            if (alo < besti && blo < bestj) {
                queue[4 * nq] = alo; queue[4 * nq + 1] = besti; queue[4 * nq + 2] = blo; queue[4 * nq + 3] = bestj; ++nq;
            }
            if (besti + bestsize < ahi && bestj + bestsize < bhi) {
                queue[4 * nq] = besti + bestsize; queue[4 * nq + 1] = ahi;
                queue[4 * nq + 2] = bestj + bestsize; queue[4 * nq + 3] = bhi; ++nq;
            }
        }
    }
    // matching_blocks.sort()  (blocks are disjoint, so ordering by i orders the tuples)
    for (int x = 1; x < nblk; ++x) {
        const int bi = blocks[3 * x], bj = blocks[3 * x + 1], bk = blocks[3 * x + 2];
        int y = x;
        while (y > 0 && blocks[3 * (y - 1)] > bi) {
            blocks[3 * y] = blocks[3 * (y - 1)]; blocks[3 * y + 1] = blocks[3 * (y - 1) + 1];
            blocks[3 * y + 2] = blocks[3 * (y - 1) + 2]; --y;
        }
        blocks[3 * y] = bi; blocks[3 * y + 1] = bj; blocks[3 * y + 2] = bk;
    }
    // collapse adjacent blocks, accumulate ns[offset] in first-seen order, sentinel (la, lb, 0) last
    int nkeys = 0;
    auto add_ns = [&](int off, int size) {
        for (int q = 0; q < nkeys; ++q)
            if (ns_key[q] == off) { ns_val[q] += size; return; }
        ns_key[nkeys] = off; ns_val[nkeys] = size; ++nkeys;
    };
    int i1 = 0, j1 = 0, k1 = 0;
    for (int x = 0; x < nblk; ++x) {
        const int i2 = blocks[3 * x], j2 = blocks[3 * x + 1], k2 = blocks[3 * x + 2];
        if (i1 + k1 == i2 && j1 + k1 == j2) k1 += k2;
        else { if (k1) add_ns(j1 - i1, k1); i1 = i2; j1 = j2; k1 = k2; }
    }
    if (k1) add_ns(j1 - i1, k1);
    add_ns(lb - la, 0);
    // log_px and argmax (first maximum wins, like max() over dict keys)
    const double error_rate = 0.2;
    const double back_ratio = 6.5 * 10e-4;
    const double p_same = 1.0 - 2.0 * error_rate + 26.0 / 25.0 * (error_rate * error_rate);
    const double lsame = log(p_same / 0.25);
    const double N = (double)la;
    int best_key = ns_key[0];
    double best_lp = 0.0;
    for (int q = 0; q < nkeys; ++q) {
        const int key = ns_key[q];
        double lp;
        if (key < 0) lp = (double)(-key) * log(back_ratio * N * jump_step_ratio) - logfact[-key] + (double)ns_val[q] * lsame;
        else lp = (double)key * log(N * jump_step_ratio) - logfact[key] + (double)ns_val[q] * lsame;
        if (q == 0 || lp > best_lp) { best_lp = lp; best_key = key; }
    }
    return best_key;
}
