/* CPU ORACLE -- TEST INFRASTRUCTURE ONLY (see oracle/chiron_oracle.py header).
 *
 * Plain-C restatement of the two TensorFlow CTC decoders the reference calls:
 *   tf.nn.ctc_greedy_decoder(merge_repeated=True)                     chiron/chiron_eval.py:486-487
 *   tf.nn.ctc_beam_search_decoder(merge_repeated=False, top_paths=1)  chiron/chiron_eval.py:489-492
 * TensorFlow 1.15 is a pinned third-party dependency (setup.py:28) absent from /root/reference; the algorithm below
 * is the published CTCBeamSearchDecoder::Step/TopPaths of tensorflow/core/util/ctc/ctc_beam_search.h.
 * Same tie rules as the Python restatement: branches in descending newp.total (stable), bottom = first minimum
 * in push order.  Build: gcc -O2 -shared -fPIC -o oracle/_build/libctc_oracle.so oracle/ctc_oracle.c -lm
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int parent, label;
    int child[8];           /* node index of child with label c, -1 = children not populated */
    float o_total, o_blank, o_label;
    float n_total, n_blank, n_label;
} node_t;

static float lse(float a, float b) {
    if (a == -INFINITY && b == -INFINITY) return -INFINITY;
    return a > b ? a + log1pf(expf(b - a)) : b + log1pf(expf(a - b));
}

static int bottom_of(const node_t* nd, const int* leaves, int n) {
    int best = 0;
    for (int i = 1; i < n; ++i)
        if (nd[leaves[i]].n_total < nd[leaves[best]].n_total) best = i;
    return best;
}

/* logits [T][C] row major, blank = C-1 (C <= 9).  Returns the decoded length (labels in out[0..T)), -1 on OOM. */
static int oracle_ctc_beam_scored(const float* logits, int T, int C, int len, int W, int* out, float* score);
int oracle_ctc_beam(const float* logits, int T, int C, int len, int W, int* out) {
    return oracle_ctc_beam_scored(logits, T, C, len, W, out, 0);
}

/* score (may be NULL): newp.total of the best beam = what TopPaths() reports as the path's log probability (the
 * `log_prob` output of tf.nn.ctc_beam_search_decoder in TF 1.15: inputs are max-subtracted per frame, not soft-maxed). */
static int oracle_ctc_beam_scored(const float* logits, int T, int C, int len, int W, int* out, float* score) {
    const int blank = C - 1;
    size_t cap = 1 + (size_t)(len > 0 ? len : 1) * (size_t)W * (size_t)(C - 1);
    node_t* nd = (node_t*)malloc(cap * sizeof(node_t));
    int* leaves = (int*)malloc(sizeof(int) * (size_t)(W + 1));
    int* branches = (int*)malloc(sizeof(int) * (size_t)(W + 1));
    float inp[16];
    if (!nd || !leaves || !branches || C > 9) { free(nd); free(leaves); free(branches); return -1; }
    (void)T;
    int n_nodes = 1, n_leaves = 1;
    nd[0].parent = -1; nd[0].label = -1;
    for (int c = 0; c < 8; ++c) nd[0].child[c] = -1;
    nd[0].o_total = nd[0].o_blank = nd[0].o_label = -INFINITY;
    nd[0].n_total = 0.f; nd[0].n_blank = 0.f; nd[0].n_label = -INFINITY;
    leaves[0] = 0;
    for (int t = 0; t < len; ++t) {
        const float* row = logits + (size_t)t * C;
        float mx = row[0];
        for (int c = 1; c < C; ++c) if (row[c] > mx) mx = row[c];
        for (int c = 0; c < C; ++c) inp[c] = row[c] - mx;
        /* Extract(): stable insertion sort, descending newp.total */
        int nb = n_leaves;
        for (int i = 0; i < nb; ++i) {
            int v = leaves[i], j = i;
            while (j > 0 && nd[branches[j - 1]].n_total < nd[v].n_total) { branches[j] = branches[j - 1]; --j; }
            branches[j] = v;
        }
        n_leaves = 0;
        for (int i = 0; i < nb; ++i) {
            node_t* b = &nd[branches[i]];
            b->o_total = b->n_total; b->o_blank = b->n_blank; b->o_label = b->n_label;
        }
        for (int i = 0; i < nb; ++i) {
            node_t* b = &nd[branches[i]];
            if (b->parent >= 0) {
                const node_t* p = &nd[b->parent];
                if (p->n_total != -INFINITY) {
                    float prev = (b->label == p->label) ? p->o_blank : p->o_total;
                    b->n_label = lse(b->n_label, prev);
                }
                b->n_label += inp[b->label];
            }
            b->n_blank = b->o_total + inp[blank];
            b->n_total = lse(b->n_blank, b->n_label);
            leaves[n_leaves++] = branches[i];
        }
        for (int i = 0; i < nb; ++i) {
            const int bi = branches[i];
            float tot = nd[bi].o_total;
            if (!(tot > -INFINITY && (n_leaves < W || tot > nd[leaves[bottom_of(nd, leaves, n_leaves)]].n_total)))
                continue;
            if (nd[bi].child[0] < 0) {
                for (int c = 0; c < C - 1; ++c) {
                    node_t* ch = &nd[n_nodes];
                    ch->parent = bi; ch->label = c;
                    for (int k = 0; k < 8; ++k) ch->child[k] = -1;
                    ch->o_total = ch->o_blank = ch->o_label = -INFINITY;
                    ch->n_total = ch->n_blank = ch->n_label = -INFINITY;
                    nd[bi].child[c] = n_nodes++;
                }
            }
            for (int c = 0; c < C - 1; ++c) {
                node_t* ch = &nd[nd[bi].child[c]];
                if (ch->n_total != -INFINITY) continue;
                ch->n_blank = -INFINITY;
                float prev = (c == nd[bi].label) ? nd[bi].o_blank : nd[bi].o_total;
                ch->n_label = inp[c] + prev;
                ch->n_total = ch->n_label;
                int cand = ch->n_total > -INFINITY &&
                           (n_leaves < W || ch->n_total > nd[leaves[bottom_of(nd, leaves, n_leaves)]].n_total);
                if (cand) {
                    if (n_leaves == W) {
                        int bo = bottom_of(nd, leaves, n_leaves);
                        node_t* bt = &nd[leaves[bo]];
                        bt->n_total = bt->n_blank = bt->n_label = -INFINITY;
                        memmove(leaves + bo, leaves + bo + 1, sizeof(int) * (size_t)(n_leaves - bo - 1));
                        --n_leaves;
                    }
                    leaves[n_leaves++] = nd[bi].child[c];
                } else {
                    ch->o_total = ch->o_blank = ch->o_label = -INFINITY;
                    ch->n_total = ch->n_blank = ch->n_label = -INFINITY;
                }
            }
        }
    }
    int best = 0;
    for (int i = 1; i < n_leaves; ++i)
        if (nd[leaves[i]].n_total > nd[leaves[best]].n_total) best = i;
    if (score) *score = nd[leaves[best]].n_total;
    int n = 0, cur = leaves[best];
    while (nd[cur].parent >= 0) { ++n; cur = nd[cur].parent; }
    cur = leaves[best];
    for (int i = n - 1; i >= 0; --i) { out[i] = nd[cur].label; cur = nd[cur].parent; }
    free(nd); free(leaves); free(branches);
    return n;
}

/* Best-path decoding: argmax (first max), collapse repeats, drop blank. */
int oracle_ctc_greedy(const float* logits, int T, int C, int len, int* out) {
    int n = 0, prev = -1;
    (void)T;
    for (int t = 0; t < len; ++t) {
        const float* row = logits + (size_t)t * C;
        int m = 0;
        for (int c = 1; c < C; ++c) if (row[c] > row[m]) m = c;
        if (m != C - 1 && m != prev) out[n++] = m;
        prev = m;
    }
    return n;
}

/* Batched helpers: logits [B][T][C]; out [B][T]; out_len [B]. */
void oracle_ctc_beam_batch(const float* logits, int B, int T, int C, const int* lens, int W, int* out, int* out_len) {
    for (int b = 0; b < B; ++b)
        out_len[b] = oracle_ctc_beam(logits + (size_t)b * T * C, T, C, lens[b], W, out + (size_t)b * T);
}

void oracle_ctc_beam_batch_scored(const float* logits, int B, int T, int C, const int* lens, int W, int* out, int* out_len,
                                  float* scores) {
    for (int b = 0; b < B; ++b)
        out_len[b] = oracle_ctc_beam_scored(logits + (size_t)b * T * C, T, C, lens[b], W, out + (size_t)b * T, scores + b);
}

void oracle_ctc_greedy_batch(const float* logits, int B, int T, int C, const int* lens, int* out, int* out_len) {
    for (int b = 0; b < B; ++b)
        out_len[b] = oracle_ctc_greedy(logits + (size_t)b * T * C, T, C, lens[b], out + (size_t)b * T);
}
