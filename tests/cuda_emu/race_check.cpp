// Race check of the SIMT kernels on the host -- TEST INFRASTRUCTURE ONLY (tests/test_cuda_emu.py::test_kernels_are_race_free).
//
// Built with -fsanitize=thread: under the emulation of cuda_emu.h every CUDA thread is a real thread and __syncthreads /
// __syncwarp / the warp collectives are the only synchronisation, so a kernel that reads shared (or global) memory another
// thread of its block wrote without a barrier in between -- the bugs compute-sanitizer's racecheck looks for, including
// warp-synchronous code that forgot its __syncwarp -- is reported by ThreadSanitizer as a data race.  Results are not
// checked here (tests/test_cuda_emu.py does that against the oracle); inputs are small and synthetic.
#include "emu_fp32_path.cpp"

#include <cstdio>
#include <random>
#include <string>

namespace {
std::mt19937 rng(7);
std::vector<float> rnd(size_t n, float scale = 1.f, float bias = 0.f) {
    std::normal_distribution<float> d(0.f, 1.f);
    std::vector<float> v(n);
    for (auto& x : v) x = d(rng) * scale + bias;
    return v;
}
}  // namespace

// Positive control: a kernel that forgets the barrier between writing and reading shared memory.
static int racy_sum = 0;
static void racy_kernel(int* out) {
    __shared__ int buf[64];
    buf[threadIdx.x] = (int)threadIdx.x;
    // missing __syncthreads()
    out[threadIdx.x] = buf[(threadIdx.x + 1) % 64];
}

int main(int argc, char** argv) {
    if (argc > 1 && std::string(argv[1]) == "--self-test") {
        std::vector<int> out(64);
        emu::launch(1, 64, [&] { racy_kernel(out.data()); });
        for (int v : out) racy_sum += v;
        std::printf("self-test ran (%d)\n", racy_sum);
        return 0;
    }
    // ---- conv stack, both BatchNorm modes, with and without a stem -------------------------------------------------------------
    for (int stem = 0; stem < 2; ++stem) {
        const int C = 8, B = 3, L = 53, nb = 2;
        int geom[21] = {nb, C, 1, 3, 3, 0, 0, 0, 0, 0, 0, 1, 2, 0, 0, 0, 0, 0, 0, stem ? 9 : 0, stem ? 5 : 0};
        std::vector<std::vector<float>> keep;
        auto mk = [&](size_t n, float s = 0.5f, float b = 0.f) { keep.push_back(rnd(n, s, b)); return (const float*)keep.back().data(); };
        const float* stem_t[5] = {mk(9 * C), mk(C, 0.1f, 1.f), mk(C, 0.1f), mk(C, 0.1f, 1.f), mk(C, 0.1f)};
        std::vector<const float*> raw, folded;
        for (int b = 0; b < nb; ++b) {
            const int cin = (b == 0 && !stem) ? 1 : C;
            const size_t wn[4] = {(size_t)cin * C, (size_t)cin * C, (size_t)3 * C * C, (size_t)C * C};
            for (int i = 0; i < 4; ++i) { raw.push_back(mk(wn[i])); raw.push_back(mk(C, 0.1f, 1.f)); raw.push_back(mk(C, 0.1f)); }
            folded.push_back(mk((size_t)C * C)); folded.push_back(mk(C, 0.1f));
            folded.push_back(mk((size_t)3 * C * C)); folded.push_back(mk(C, 0.1f));
            folded.push_back(mk((size_t)2 * C * C)); folded.push_back(mk(C, 0.1f));
        }
        const float* rank1[6] = {mk(C), mk(C, 0.1f, 1.f), mk(C, 0.1f), mk(C), mk(C, 0.1f, 1.f), mk(C, 0.1f)};
        std::vector<float> x = rnd((size_t)B * L, 0.43f, -0.16f), out((size_t)B * L * C);
        long long n_launch = 0;
        for (int sms : {1, 148}) {
            if (emu_conv_stack(CB_BN_BATCH, geom, raw.data(), rank1, stem_t, x.data(), B, L, sms, out.data(), &n_launch) <= 0) return 2;
            if (emu_conv_stack(CB_BN_POPULATION, geom, folded.data(), rank1, stem_t, x.data(), B, L, sms, out.data(), &n_launch) <= 0) return 2;
        }
    }
    // ---- recurrences --------------------------------------------------------------------------------------------------------------
    {
        const int B = 5, T = 6, H = 8;
        std::vector<int32_t> lens = {6, 0, 3, 6, 1};
        std::vector<float> pre = rnd((size_t)B * T * 8 * H), whh = rnd((size_t)H * 4 * H, 0.3f), out((size_t)B * T * 2 * H);
        std::vector<float> wg = rnd((size_t)H * 2 * H, 0.3f), wc = rnd((size_t)H * H, 0.3f);
        for (int rg : {1, 2, 4}) {
            if (emu_lstm(rg, B, T, H, pre.data(), 8 * H, whh.data(), whh.data(), lens.data(), out.data(), 2 * H)) return 3;
            if (emu_gru(rg, B, T, H, pre.data(), 6 * H, wg.data(), wg.data(), wc.data(), wc.data(), lens.data(), out.data(), 2 * H)) return 3;
        }
    }
    // ---- head, path_prob, seq_len, greedy ----------------------------------------------------------------------------------------
    {
        const int B = 9, T = 40, H = 12, C = 5, Bp = 128;
        std::vector<float> lasth = rnd((size_t)B * T * 2 * H), tm = rnd((size_t)T * 2 * (H / 4) * Bp * 4), w = rnd(2 * H), bias = rnd(H);
        std::vector<float> wcl = rnd((size_t)H * C), bc = rnd(C), logits((size_t)B * T * C), prob(B);
        emu_head(lasth.data(), (long long)B * T, H, C, w.data(), bias.data(), wcl.data(), bc.data(), logits.data(), 2);
        emu_head_tmajor(tm.data(), B, Bp, T, H, C, w.data(), bias.data(), wcl.data(), bc.data(), logits.data());
        emu_path_prob(logits.data(), B, T, C, prob.data());
        std::vector<int32_t> lens = {40, 0, 1, 33, 40, 7, 32, 31, 12}, lo(B), nbs(B);
        std::vector<int8_t> bases((size_t)B * T);
        emu_seq_len(lens.data(), B, 400, T, lo.data());
        emu_greedy(logits.data(), lens.data(), B, T, C, bases.data(), nbs.data());
        // ---- beam search: the two cooperative variants and the fallback ------------------------------------------------------------
        for (int warp : {1, 2, 0})
            for (int W : {1, 5, 12})
                if (emu_beam(warp, logits.data(), lens.data(), B, T, C, W, 2 * W * (T + 1) + 2, bases.data(), nbs.data()) != 0) return 4;
        {   // the launcher's three passes with tiny pools, so that windows are marked, retried and finished by the third pass
            int marked[3] = {0, 0, 0};
            if (emu_beam_passes(logits.data(), lens.data(), B, T, C, 12, 2 * 12 + 2, 3 * 12, B, bases.data(), nbs.data(), marked) != 0) return 6;
            if (marked[2] != 0) return 7;
        }
        // ---- assembly: every kernel --------------------------------------------------------------------------------------------------
        std::uniform_int_distribution<int> base4(0, 3);
        const int n = 14, Tb = 30;
        std::vector<int8_t> segs((size_t)n * Tb);
        for (auto& v : segs) v = (int8_t)base4(rng);
        std::vector<int32_t> nb = {20, 25, 0, 30, 18, 22, 27, 0, 30, 21, 19, 24, 26, 23}, pos(n), out_len(1);
        std::vector<float> pp = rnd(n, 1.f, 6.f);
        const int max_len = 400;
        std::vector<int8_t> cons(max_len);
        std::vector<char> qual(max_len);
        for (int kernel : {CB_ASM_SIMPLE, CB_ASM_GLUE, CB_ASM_STICK})
            if (emu_assemble(segs.data(), nb.data(), pp.data(), n, Tb, kernel == CB_ASM_SIMPLE ? 200 : 390, 400, kernel, cons.data(),
                             qual.data(), pos.data(), out_len.data(), max_len)) return 5;
    }
    std::printf("race_check: %lld emulated launches, %lld CUDA threads\n", emu::launches, emu::threads_run);
    return 0;
}
