// Host-side signal preparation (no GPU work): the per-read Python hot spots in front of the GPU path.
//
//   cb_host_parse_signal   chiron_input.read_signal           chiron/chiron_input.py:527-532   text -> float32 samples
//   cb_host_normalize      read_signal / read_signal_fast5     chiron/chiron_input.py:535-538, 548-554   median / MAD
//   cb_host_windows        read_data_for_eval + padding        chiron/chiron_input.py:253-292, 681-692   sliding windows
//   cb_host_format_segments  write_output's segment records    chiron/chiron_eval.py:211-214             dense bases -> text
//
// In the reference these are a Python token loop, np.unique (a sort) and per-window list slicing; here they are single
// passes in C that release the GIL (ctypes), so the reader threads of chiron_eval.evaluation() run in parallel and keep up
// with the GPU.  Results are bit-identical to the numpy formulation (float64 statistics, (s - med) / mad rounded once to
// float32): tests/test_host_signal.py.
#include <math.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "../../include/chiron_b200.h"

void cb_set_error(const char* fmt, ...);

namespace {

inline bool is_space(unsigned char c) { return c == ' ' || c == '\n' || c == '\t' || c == '\r' || c == '\v' || c == '\f'; }

// median of v (modified): mean of the two middle order statistics for even counts, like np.median
double median_inplace(std::vector<double>& v) {
    const size_t n = v.size();
    if (n == 0) return NAN;
    const size_t hi = n / 2;
    std::nth_element(v.begin(), v.begin() + hi, v.end());
    const double b = v[hi];
    if (n & 1) return b;
    const double a = *std::max_element(v.begin(), v.begin() + hi);
    return (a + b) / 2.0;
}

constexpr double MAD_C = 0.6744897501960817;     // statsmodels.robust.mad: Phi^-1(3/4)
constexpr int INT_RANGE = 1 << 17;               // integral samples in [-2^17, 2^17) take the presence-bitmap path

}  // namespace

extern "C" long long cb_host_parse_signal(const char* text, size_t nbytes, float* out, size_t cap) {
    if (!text && nbytes) { cb_set_error("cb_host_parse_signal: bad arguments"); return CB_ERR_ARG; }
    size_t i = 0, n = 0;
    while (i < nbytes) {
        while (i < nbytes && is_space((unsigned char)text[i])) ++i;
        if (i >= nbytes) break;
        const size_t start = i;
        // fast path: [+-]digits (what .signal files hold: raw DAC integers), at most 15 digits -> exact in double
        bool neg = false;
        if (text[i] == '-' || text[i] == '+') { neg = text[i] == '-'; ++i; }
        long long v = 0;
        size_t digits = 0;
        while (i < nbytes && text[i] >= '0' && text[i] <= '9' && digits < 16) { v = v * 10 + (text[i] - '0'); ++i; ++digits; }
        double val;
        if (digits > 0 && digits <= 15 && (i >= nbytes || is_space((unsigned char)text[i]))) {
            val = neg ? -(double)v : (double)v;
        } else {                                    // general token: the C locale's strtod on a bounded copy
            i = start;
            while (i < nbytes && !is_space((unsigned char)text[i])) ++i;
            const size_t len = i - start;
            char buf[128];
            if (len >= sizeof(buf)) { cb_set_error("cb_host_parse_signal: token of %zu characters at byte %zu", len, start); return CB_ERR_ARG; }
            memcpy(buf, text + start, len);
            buf[len] = 0;
            char* end = nullptr;
            val = strtod(buf, &end);
            if (end != buf + len) { cb_set_error("cb_host_parse_signal: could not convert '%s' to a number (byte %zu)", buf, start); return CB_ERR_ARG; }
        }
        if (out) {
            if (n >= cap) { cb_set_error("cb_host_parse_signal: more than %zu samples", cap); return CB_ERR_ARG; }
            out[n] = (float)val;
        }
        ++n;
    }
    return (long long)n;
}

extern "C" int cb_host_normalize(const float* in, size_t n, int mode, float* out) {
    if ((!in || !out) && n) { cb_set_error("cb_host_normalize: bad arguments"); return CB_ERR_ARG; }
    if (mode < 0 || mode > 2) { cb_set_error("cb_host_normalize: unknown signal normalisation %d", mode); return CB_ERR_ARG; }
    if (mode == 0 || n == 0) { if (out != in && n) memmove(out, in, n * sizeof(float)); return CB_OK; }
    for (size_t i = 0; i < n; ++i)
        if (in[i] != in[i]) {                        // a NaN sample: np.median is NaN, so every output is (sorting NaNs is undefined)
            for (size_t j = 0; j < n; ++j) out[j] = NAN;
            return CB_OK;
        }
    std::vector<double> ref;
    if (mode == 1) {                                 // statistics over the UNIQUE sample values (chiron_input.py:548)
        bool integral = true;
        for (size_t i = 0; i < n && integral; ++i) {
            const float v = in[i];
            integral = v >= -(float)INT_RANGE && v < (float)INT_RANGE && v == (float)(int)v;
        }
        if (integral) {                              // presence bitmap: O(n), no sort
            std::vector<unsigned char> seen((size_t)2 * INT_RANGE, 0);
            for (size_t i = 0; i < n; ++i) seen[(size_t)((int)in[i] + INT_RANGE)] = 1;
            for (size_t k = 0; k < seen.size(); ++k) if (seen[k]) ref.push_back((double)((long long)k - INT_RANGE));
        } else {
            ref.assign(in, in + n);
            std::sort(ref.begin(), ref.end());
            ref.erase(std::unique(ref.begin(), ref.end()), ref.end());
        }
    } else {
        ref.assign(in, in + n);
    }
    std::vector<double> work(ref);
    const double med = median_inplace(work);
    for (size_t i = 0; i < ref.size(); ++i) work[i] = fabs(ref[i] - med);
    const double mad = median_inplace(work) / MAD_C;
    for (size_t i = 0; i < n; ++i) out[i] = (float)(((double)in[i] - med) / mad);
    return CB_OK;
}

extern "C" long long cb_host_windows(const float* sig, size_t n, int jump, int L, float* x, int32_t* lens, size_t cap_windows) {
    if (jump < 1 || L < 1 || (!sig && n)) { cb_set_error("cb_host_windows: bad arguments"); return CB_ERR_ARG; }
    const size_t n_win = (n + (size_t)jump - 1) / (size_t)jump;
    if (!x || !lens) return (long long)n_win;        // count only
    if (n_win > cap_windows) { cb_set_error("cb_host_windows: %zu windows, room for %zu", n_win, cap_windows); return CB_ERR_ARG; }
    for (size_t w = 0; w < n_win; ++w) {
        const size_t s = w * (size_t)jump;
        const size_t len = n - s < (size_t)L ? n - s : (size_t)L;
        memcpy(x + w * (size_t)L, sig + s, len * sizeof(float));
        if (len < (size_t)L) memset(x + w * (size_t)L + len, 0, ((size_t)L - len) * sizeof(float));
        lens[w] = (int32_t)len;
    }
    return (long long)n_win;
}

extern "C" long long cb_host_format_segments(const char* name, const int8_t* bases, const int32_t* n_bases, int n_windows, int T,
                                             char* out, size_t cap) {
    if (!name || n_windows < 0 || T < 1 || ((!bases || !n_bases) && n_windows)) { cb_set_error("cb_host_format_segments: bad arguments"); return CB_ERR_ARG; }
    static const char LUT[4] = {'A', 'C', 'G', 'T'};
    const size_t name_len = strlen(name);
    size_t pos = 0;
    int kept = 0;
    for (int w = 0; w < n_windows; ++w) {
        int n = n_bases[w];
        if (n <= 0) continue;                          // sparse2dense drops windows without bases (chiron_eval.py:56-66)
        if (n > T) n = T;
        char idx[16];
        const int idx_len = snprintf(idx, sizeof(idx), "%d", kept++);
        const size_t need = 1 + name_len + (size_t)idx_len + 1 + (size_t)n + 1;
        if (out) {
            if (pos + need > cap) { cb_set_error("cb_host_format_segments: output buffer of %zu bytes is too small", cap); return CB_ERR_ARG; }
            char* p = out + pos;
            *p++ = '>';
            memcpy(p, name, name_len); p += name_len;
            memcpy(p, idx, (size_t)idx_len); p += idx_len;
            *p++ = '\n';
            const int8_t* row = bases + (size_t)w * T;
            for (int i = 0; i < n; ++i) p[i] = LUT[row[i] & 3];
            p[n] = '\n';
        }
        pos += need;
    }
    return (long long)pos;
}
