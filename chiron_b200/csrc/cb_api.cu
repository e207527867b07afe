// C-ABI of libchiron_b200.so: handle lifecycle, weight preparation, workspace, forward orchestration, host wrappers.
// See include/chiron_b200.h for the contract and the reference call sites each entry point replaces.
#include <math.h>
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include "cb_internal.cuh"
#include "cb_conv_stack.cuh"

static thread_local char g_err[1024] = "";

void cb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

extern "C" const char* cb_last_error(void) { return g_err; }
extern "C" const char* cb_version(void) { return "chiron_b200 0.1 (sm_100a)"; }

namespace {

constexpr float BN_EPS = 1e-5f;   // chiron/cnn.py:125,187

struct BlobHeader {               // chiron_b200/model.py: "<4s8i8i8i6iq" (packed, 132 bytes)
    char magic[4];
    int32_t version, n_blocks, channels, hidden, n_layers, n_class, rnn_layout, branch1_bn_mask;
    int32_t k[8], stride[8];
    int32_t sig_norm, reverse_signal, bn_mode, cell_type, stem_k, stem_stride;
};
constexpr size_t HEADER_BYTES = 4 + 30 * 4 + 8;

struct HostBuilder {              // accumulates the derived fp32 weights; every tensor starts 16-byte aligned
    std::vector<float> data;
    size_t add(const std::vector<float>& v) {
        while (data.size() & 3) data.push_back(0.f);
        size_t off = data.size();
        data.insert(data.end(), v.begin(), v.end());
        return off;
    }
};

inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

int out_len_of(const CbConfig& c, int L) {
    int T = c.stem_k > 0 ? (L + c.stem_stride - 1) / c.stem_stride : L;
    for (int b = 0; b < c.n_blocks; ++b) T = (T + c.stride[b] - 1) / c.stride[b];
    return T;
}

}  // namespace

// -----------------------------------------------------------------------------------------------------------------
extern "C" int cb_create(const void* blob, size_t nbytes, int device, int precision, cb_handle** out) {
    if (!blob || !out || nbytes < HEADER_BYTES) { cb_set_error("cb_create: bad arguments"); return CB_ERR_ARG; }
    if (precision < CB_PREC_FP32 || precision > CB_PREC_TC_SPLIT) { cb_set_error("cb_create: unknown precision %d", precision); return CB_ERR_ARG; }
    BlobHeader hd;
    memcpy(&hd, blob, sizeof(hd));
    int64_t n_floats;
    memcpy(&n_floats, (const char*)blob + HEADER_BYTES - 8, 8);
    if (memcmp(hd.magic, "CBW1", 4) != 0 || hd.version != 1) { cb_set_error("cb_create: not a CBW1 blob"); return CB_ERR_BLOB; }
    if (hd.n_blocks < 1 || hd.n_blocks > CB_MAX_BLOCKS || hd.n_layers < 1 || hd.n_layers > CB_MAX_LAYERS ||
        hd.channels < 4 || (hd.channels & 3) || hd.hidden < 4 || (hd.hidden & 3) || hd.n_class < 2 || hd.n_class > 8) {
        cb_set_error("cb_create: unsupported topology in blob header");
        return CB_ERR_BLOB;
    }
    if (n_floats < 0 || (uint64_t)n_floats > (nbytes - HEADER_BYTES) / 4) { cb_set_error("cb_create: truncated blob"); return CB_ERR_BLOB; }
    {   // the float count the header's topology needs, BEFORE any tensor is read (a corrupt or mismatched header must
        // not walk the tensor walk below off the end of the blob)
        const uint64_t C = hd.channels, H = hd.hidden, bn = 4 * C;
        bool geom_ok = hd.stem_k >= 0 && hd.stem_k <= 64 && hd.stem_stride >= 0 && hd.stem_stride <= 64;
        uint64_t need = hd.stem_k > 0 ? (uint64_t)hd.stem_k * C + bn : 0;
        for (int b = 0; b < hd.n_blocks; ++b) {
            if (hd.k[b] < 1 || hd.k[b] > 64 || hd.stride[b] < 1 || hd.stride[b] > 64) { geom_ok = false; break; }
            const uint64_t cin = (b == 0 && hd.stem_k == 0) ? 1 : C;
            need += cin * C + (((hd.branch1_bn_mask >> b) & 1) ? bn : 0) + cin * C + bn + (uint64_t)hd.k[b] * C * C + bn + C * C + bn;
        }
        for (int l = 0; l < hd.n_layers; ++l) {
            const uint64_t in = l == 0 ? C : (hd.rnn_layout == 0 ? 2 * H : H);
            need += 2 * (hd.cell_type == CB_CELL_GRU ? (in + H) * 2 * H + 2 * H + (in + H) * H + H : (in + H) * 4 * H + 4 * H);
        }
        need += 2 * H + H + H * (uint64_t)hd.n_class + (uint64_t)hd.n_class;
        if (!geom_ok) { cb_set_error("cb_create: bad conv geometry in blob header"); return CB_ERR_BLOB; }
        if (need != (uint64_t)n_floats) {
            cb_set_error("cb_create: blob holds %lld floats, topology needs %llu", (long long)n_floats, (unsigned long long)need);
            return CB_ERR_BLOB;
        }
    }
    const float* w = (const float*)((const char*)blob + HEADER_BYTES);

    cb_handle* h = new cb_handle();
    memset(h, 0, sizeof(*h));
    if (const char* rs = getenv("CB_RESERVE_SMS")) { const int n = atoi(rs); h->reserve_sms = n > 0 && n <= 16 ? (n + 1) & ~1 : 0; }
    h->device = device;
    h->precision = precision;
    CbConfig& c = h->cfg;
    c.n_blocks = hd.n_blocks; c.channels = hd.channels; c.hidden = hd.hidden; c.n_layers = hd.n_layers;
    c.n_class = hd.n_class; c.rnn_layout = hd.rnn_layout; c.branch1_bn_mask = hd.branch1_bn_mask;
    for (int i = 0; i < CB_MAX_BLOCKS; ++i) { c.k[i] = hd.k[i]; c.stride[i] = hd.stride[i]; }
    c.sig_norm = hd.sig_norm; c.reverse_signal = hd.reverse_signal; c.cell_type = hd.cell_type;
    if (c.cell_type != CB_CELL_LSTM && c.cell_type != CB_CELL_GRU) { delete h; cb_set_error("cb_create: unknown RNN cell type %d", c.cell_type); return CB_ERR_BLOB; }
    c.stem_k = hd.stem_k; c.stem_stride = hd.stem_stride;
    if (c.stem_k < 0 || c.stem_k > 64 || (c.stem_k > 0) != (c.stem_stride > 0) || c.stem_stride < 0) {
        delete h; cb_set_error("cb_create: bad stem convolution geometry in blob header"); return CB_ERR_BLOB;
    }
    if (c.cell_type == CB_CELL_GRU && precision != CB_PREC_FP32) {
        delete h;
        cb_set_error("cb_create: GRU cells run on the CB_PREC_FP32 path only (the tensor-core recurrence is an LSTM kernel)");
        return CB_ERR_ARG;
    }
    const int C = c.channels, H = c.hidden;
    for (int b = 0; b < c.n_blocks; ++b)
        if (c.k[b] < 1 || c.stride[b] < 1) { delete h; cb_set_error("cb_create: bad conv geometry"); return CB_ERR_BLOB; }

    // ---- walk the canonical tensor order (model.py: tensor_specs) and derive the kernels' operands ----------------
    size_t pos = 0;
    auto take = [&](size_t n) -> const float* { const float* p = w + pos; pos += n; return p; };
    struct Bn { std::vector<float> inv, shift; const float *scale, *offset; };
    auto take_bn = [&]() {           // tf.nn.batch_normalization, population statistics (cnn.py:160-161)
        Bn bn; bn.inv.resize(C); bn.shift.resize(C);
        const float *scale = take(C), *offset = take(C), *mean = take(C), *var = take(C);
        bn.scale = scale; bn.offset = offset;
        for (int n = 0; n < C; ++n) {
            bn.inv[n] = scale[n] * (1.0f / sqrtf(var[n] + BN_EPS));
            bn.shift[n] = offset[n] - mean[n] * bn.inv[n];
        }
        return bn;
    };
    HostBuilder hb;
    size_t o_conv2a[CB_MAX_BLOCKS][2], o_conv2b[CB_MAX_BLOCKS][2], o_convc[CB_MAX_BLOCKS][2];
    size_t o_g[3] = {0, 0, 0}, o_r[3] = {0, 0, 0};
    struct RawOff { size_t W, scale, offset; bool bn; } o_raw[CB_MAX_BLOCKS][4];   // branch1, conv2a, conv2b, conv2c
    size_t o_stem[5] = {0, 0, 0, 0, 0};
    if (c.stem_k > 0) {                  // conv_layer/conv1 [k,C] + BN (cnn.py:454-476): raw weights, folded and raw BN vectors
        const float* ws = take((size_t)c.stem_k * C);
        Bn bns = take_bn();
        o_stem[0] = hb.add(std::vector<float>(ws, ws + (size_t)c.stem_k * C));
        o_stem[1] = hb.add(bns.inv);
        o_stem[2] = hb.add(bns.shift);
        o_stem[3] = hb.add(std::vector<float>(bns.scale, bns.scale + C));
        o_stem[4] = hb.add(std::vector<float>(bns.offset, bns.offset + C));
    }
    for (int b = 0; b < c.n_blocks; ++b) {
        const bool rank1 = b == 0 && c.stem_k == 0;       // block 1 reads the one-channel signal itself
        const int cin = rank1 ? 1 : C;
        const float* w1 = take((size_t)cin * C);
        Bn bn1; bool has_bn1 = (c.branch1_bn_mask >> b) & 1;
        if (has_bn1) bn1 = take_bn();
        const float* w2a = take((size_t)cin * C);
        Bn bna = take_bn();
        const float* w2b = take((size_t)c.k[b] * C * C);
        Bn bnb = take_bn();
        const float* w2c = take((size_t)C * C);
        Bn bnc = take_bn();
        if (rank1) {
            o_g[0] = hb.add(std::vector<float>(w2a, w2a + C));
            o_g[1] = hb.add(bna.inv);
            o_g[2] = hb.add(bna.shift);
            o_r[0] = hb.add(std::vector<float>(w1, w1 + C));
            o_r[1] = hb.add(has_bn1 ? bn1.inv : std::vector<float>(C, 1.0f));
            o_r[2] = hb.add(has_bn1 ? bn1.shift : std::vector<float>(C, 0.0f));
            o_conv2a[b][0] = o_conv2a[b][1] = 0;
        } else {
            std::vector<float> f((size_t)C * C);
            for (int k = 0; k < C; ++k)
                for (int n = 0; n < C; ++n) f[(size_t)k * C + n] = w2a[(size_t)k * C + n] * bna.inv[n];
            o_conv2a[b][0] = hb.add(f);
            o_conv2a[b][1] = hb.add(bna.shift);
        }
        {
            std::vector<float> f((size_t)c.k[b] * C * C);
            for (size_t k = 0; k < (size_t)c.k[b] * C; ++k)
                for (int n = 0; n < C; ++n) f[k * C + n] = w2b[k * C + n] * bnb.inv[n];
            o_conv2b[b][0] = hb.add(f);
            o_conv2b[b][1] = hb.add(bnb.shift);
        }
        {   // conv2c (+BN) and, for blocks >= 2, the 1x1 branch1 conv stacked along K (cnn.py:258-261)
            const int kc = rank1 ? C : 2 * C;
            std::vector<float> f((size_t)kc * C), sh(bnc.shift);
            for (int k = 0; k < C; ++k)
                for (int n = 0; n < C; ++n) f[(size_t)k * C + n] = w2c[(size_t)k * C + n] * bnc.inv[n];
            if (!rank1) {
                for (int k = 0; k < C; ++k)
                    for (int n = 0; n < C; ++n)
                        f[(size_t)(C + k) * C + n] = w1[(size_t)k * C + n] * (has_bn1 ? bn1.inv[n] : 1.0f);
                if (has_bn1) for (int n = 0; n < C; ++n) sh[n] += bn1.shift[n];
            }
            o_convc[b][0] = hb.add(f);
            o_convc[b][1] = hb.add(sh);
        }
        {   // the same four convolutions with nothing folded, for the batch-statistics BN mode (cnn.py:166-188)
            const float* ws[4] = {w1, w2a, w2b, w2c};
            const size_t wn[4] = {(size_t)cin * C, (size_t)cin * C, (size_t)c.k[b] * C * C, (size_t)C * C};
            const Bn* bns[4] = {has_bn1 ? &bn1 : nullptr, &bna, &bnb, &bnc};
            for (int i = 0; i < 4; ++i) {
                o_raw[b][i].W = hb.add(std::vector<float>(ws[i], ws[i] + wn[i]));
                o_raw[b][i].bn = bns[i] != nullptr;
                if (bns[i]) {
                    o_raw[b][i].scale = hb.add(std::vector<float>(bns[i]->scale, bns[i]->scale + C));
                    o_raw[b][i].offset = hb.add(std::vector<float>(bns[i]->offset, bns[i]->offset + C));
                }
            }
        }
    }
    const size_t o_zeros = hb.add(std::vector<float>((size_t)(C > 8 * H ? C : 8 * H), 0.0f));
    size_t o_wx[CB_MAX_LAYERS][2], o_b[CB_MAX_LAYERS][2], o_whh[CB_MAX_LAYERS][2], o_wxcat[CB_MAX_LAYERS], o_bcat[CB_MAX_LAYERS];
    size_t o_gwc[CB_MAX_LAYERS][2] = {};
    const int G = c.cell_type == CB_CELL_GRU ? 3 : 4;         // hoisted input-projection columns per direction: G*H
    for (int l = 0; l < c.n_layers; ++l) {
        const int in = l == 0 ? C : (c.rnn_layout == 0 ? 2 * H : H);
        std::vector<float> wx[2], bx[2];                      // per direction: input kernel [in, G*H], bias [G*H]
        for (int d = 0; d < 2; ++d) {
            wx[d].resize((size_t)in * G * H); bx[d].resize((size_t)G * H);
            if (c.cell_type == CB_CELL_GRU) {                 // TF GRUCell: gates/kernel [in+H,2H] (r|u), candidate/kernel [in+H,H]
                const float* gk = take((size_t)(in + H) * 2 * H); const float* gb = take((size_t)2 * H);
                const float* ck = take((size_t)(in + H) * H);     const float* cb = take((size_t)H);
                for (int k = 0; k < in; ++k) {
                    memcpy(&wx[d][(size_t)k * 3 * H], gk + (size_t)k * 2 * H, sizeof(float) * 2 * H);
                    memcpy(&wx[d][(size_t)k * 3 * H + 2 * H], ck + (size_t)k * H, sizeof(float) * H);
                }
                memcpy(&bx[d][0], gb, sizeof(float) * 2 * H);
                memcpy(&bx[d][2 * H], cb, sizeof(float) * H);
                o_whh[l][d] = hb.add(std::vector<float>(gk + (size_t)in * 2 * H, gk + (size_t)(in + H) * 2 * H));
                o_gwc[l][d] = hb.add(std::vector<float>(ck + (size_t)in * H, ck + (size_t)(in + H) * H));
            } else {                                          // TF LSTMCell: kernel [in+H,4H] (i,j,f,o), bias [4H]
                const float* kern = take((size_t)(in + H) * 4 * H); const float* bias = take((size_t)4 * H);
                wx[d].assign(kern, kern + (size_t)in * 4 * H);
                bx[d].assign(bias, bias + 4 * H);
                o_whh[l][d] = hb.add(std::vector<float>(kern + (size_t)in * 4 * H, kern + (size_t)(in + H) * 4 * H));
            }
            o_wx[l][d] = hb.add(wx[d]);
            o_b[l][d] = hb.add(bx[d]);
        }
        std::vector<float> cat((size_t)in * 2 * G * H), bcat((size_t)2 * G * H);      // fw || bw (stacked-bidirectional layout)
        for (int d = 0; d < 2; ++d) {
            for (int k = 0; k < in; ++k)
                memcpy(&cat[(size_t)k * 2 * G * H + (size_t)d * G * H], &wx[d][(size_t)k * G * H], sizeof(float) * G * H);
            memcpy(&bcat[(size_t)d * G * H], bx[d].data(), sizeof(float) * G * H);
        }
        o_wxcat[l] = hb.add(cat);
        o_bcat[l] = hb.add(bcat);
    }
    const float* hw = take((size_t)2 * H);
    const float* hbias = take(H);
    const float* hwc = take((size_t)H * c.n_class);
    const float* hbc = take(c.n_class);
    size_t o_head[4] = {hb.add(std::vector<float>(hw, hw + 2 * H)), hb.add(std::vector<float>(hbias, hbias + H)),
                        hb.add(std::vector<float>(hwc, hwc + (size_t)H * c.n_class)),
                        hb.add(std::vector<float>(hbc, hbc + c.n_class))};
    if ((int64_t)pos != n_floats) {
        delete h;
        cb_set_error("cb_create: blob holds %lld floats, topology needs %zu", (long long)n_floats, pos);
        return CB_ERR_BLOB;
    }

    // ---- device ------------------------------------------------------------------------------------------------------
    cudaError_t e = cudaSetDevice(device);
    if (e != cudaSuccess) { delete h; cb_set_error("cudaSetDevice(%d): %s", device, cudaGetErrorString(e)); return CB_ERR_CUDA; }
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) { delete h; cb_set_error("cudaGetDeviceProperties: %s", cudaGetErrorString(e)); return CB_ERR_CUDA; }
    h->sm_count = prop.multiProcessorCount;
    if (prop.major != 10) {
        delete h;
        cb_set_error("chiron_b200 is built for sm_100a only; device %d is sm_%d%d", device, prop.major, prop.minor);
        return CB_ERR_CUDA;
    }
    h->weights_floats = hb.data.size();
    e = cudaMalloc(&h->d_weights, hb.data.size() * sizeof(float));
    if (e != cudaSuccess) { delete h; cb_set_error("cudaMalloc(weights): %s", cudaGetErrorString(e)); return CB_ERR_NOMEM; }
    e = cudaMemcpy(h->d_weights, hb.data.data(), hb.data.size() * sizeof(float), cudaMemcpyHostToDevice);
    if (e != cudaSuccess) { cudaFree(h->d_weights); delete h; cb_set_error("cudaMemcpy(weights): %s", cudaGetErrorString(e)); return CB_ERR_CUDA; }
    const float* base = h->d_weights;
    h->g_w = base + o_g[0]; h->g_inv = base + o_g[1]; h->g_sh = base + o_g[2];
    h->r_w = base + o_r[0]; h->r_inv = base + o_r[1]; h->r_sh = base + o_r[2];
    for (int b = 0; b < c.n_blocks; ++b) {
        h->conv2a[b].W = base + o_conv2a[b][0]; h->conv2a[b].shift = base + o_conv2a[b][1];
        h->conv2b[b].W = base + o_conv2b[b][0]; h->conv2b[b].shift = base + o_conv2b[b][1];
        h->convc[b].W = base + o_convc[b][0];   h->convc[b].shift = base + o_convc[b][1];
        CbRawConv* raw[4] = {&h->raw1[b], &h->raw2a[b], &h->raw2b[b], &h->raw2c[b]};
        for (int i = 0; i < 4; ++i) {
            raw[i]->W = base + o_raw[b][i].W;
            raw[i]->scale = o_raw[b][i].bn ? base + o_raw[b][i].scale : nullptr;
            raw[i]->offset = o_raw[b][i].bn ? base + o_raw[b][i].offset : nullptr;
        }
    }
    h->zeros = base + o_zeros;
    if (c.stem_k > 0) h->stem = CbStem{base + o_stem[0], base + o_stem[1], base + o_stem[2], base + o_stem[3], base + o_stem[4]};
    for (int l = 0; l < c.n_layers; ++l) {
        for (int d = 0; d < 2; ++d) {
            h->wx[l][d] = base + o_wx[l][d]; h->whh[l][d] = base + o_whh[l][d]; h->bias[l][d] = base + o_b[l][d];
            h->gru_wc[l][d] = c.cell_type == CB_CELL_GRU ? base + o_gwc[l][d] : nullptr;
        }
        h->wxcat[l] = base + o_wxcat[l]; h->bcat[l] = base + o_bcat[l];
    }
    h->head_w = base + o_head[0]; h->head_b = base + o_head[1]; h->head_wc = base + o_head[2]; h->head_bc = base + o_head[3];
    e = cudaMalloc(&h->d_flag, 64);
    if (e == cudaSuccess) e = cudaMemset(h->d_flag, 0, 64);
    for (int i = 0; i < 8 && e == cudaSuccess; ++i) e = cudaEventCreate(&h->ev[i]);
    if (e != cudaSuccess) { cb_set_error("cb_create: %s", cudaGetErrorString(e)); cb_destroy(h); return CB_ERR_CUDA; }
    if (precision != CB_PREC_FP32) {
        int rc = cb_tc_prepare(h, hb.data.data());
        if (rc != CB_OK) { cb_destroy(h); return rc; }
    }
    if (hd.bn_mode != CB_BN_POPULATION) {
        int rc = cb_set_bn_mode(h, hd.bn_mode);
        if (rc != CB_OK) { cb_destroy(h); return rc; }
    }
    *out = h;
    return CB_OK;
}

extern "C" int cb_set_bn_mode(cb_handle* h, int bn_mode) {
    if (!h || (bn_mode != CB_BN_POPULATION && bn_mode != CB_BN_BATCH)) { cb_set_error("cb_set_bn_mode: bad arguments"); return CB_ERR_ARG; }
    if (bn_mode == CB_BN_BATCH) {
        if (h->precision != CB_PREC_FP32) {
            cb_set_error("batch-statistics BatchNorm (CB_BN_BATCH) runs on the CB_PREC_FP32 path only");
            return CB_ERR_ARG;
        }
        if (h->cfg.channels > 1024) { cb_set_error("batch-statistics BatchNorm needs channels <= 1024"); return CB_ERR_ARG; }
        CB_CUDA(cudaSetDevice(h->device));
        if (!h->bn_part) CB_CUDA(cudaMalloc(&h->bn_part, (size_t)CB_BN_MAX_PART * 2 * h->cfg.channels * sizeof(double)));
        if (!h->bn_vec) CB_CUDA(cudaMalloc(&h->bn_vec, (size_t)CB_BN_VECS * h->cfg.channels * sizeof(float)));
    }
    h->bn_mode = bn_mode;
    return CB_OK;
}

extern "C" int cb_bn_mode(const cb_handle* h) { return h ? h->bn_mode : CB_ERR_ARG; }

extern "C" int cb_destroy(cb_handle* h) {
    if (!h) return CB_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    cb_forward_tc_release(h);
    cb_tc_release(h);
    if (h->d_weights) cudaFree(h->d_weights);
    if (h->ws) cudaFree(h->ws);
    if (h->stage) cudaFree(h->stage);
    for (int i = 0; i < 2; ++i) {
        if (h->pipe[i].pin) cudaFreeHost(h->pipe[i].pin);
        if (h->pipe[i].dev) cudaFree(h->pipe[i].dev);
        if (h->pipe[i].h2d_done) cudaEventDestroy(h->pipe[i].h2d_done);
        if (h->pipe[i].compute_done) cudaEventDestroy(h->pipe[i].compute_done);
        if (h->pipe[i].d2h_done) cudaEventDestroy(h->pipe[i].d2h_done);
    }
    if (h->pipe_in) cudaStreamDestroy(h->pipe_in);
    if (h->pipe_compute) cudaStreamDestroy(h->pipe_compute);
    if (h->pipe_out) cudaStreamDestroy(h->pipe_out);
    if (h->asm_stream) cudaStreamDestroy(h->asm_stream);
    if (h->asm_stage) cudaFree(h->asm_stage);
    if (h->beam_ws) cudaFree(h->beam_ws);
    if (h->asm_ws) cudaFree(h->asm_ws);
    if (h->d_flag) cudaFree(h->d_flag);
    if (h->bn_part) cudaFree(h->bn_part);
    if (h->bn_vec) cudaFree(h->bn_vec);
    for (int i = 0; i < 8; ++i) if (h->ev[i]) cudaEventDestroy(h->ev[i]);
    for (int i = 0; i < CB_PROF_MAX; ++i)
        for (int j = 0; j < 2; ++j) if (h->prof_ev[i][j]) cudaEventDestroy(h->prof_ev[i][j]);
    delete h;
    return CB_OK;
}

extern "C" int cb_out_len(const cb_handle* h, int L) { return h ? out_len_of(h->cfg, L) : CB_ERR_ARG; }
extern "C" int cb_n_class(const cb_handle* h) { return h ? h->cfg.n_class : CB_ERR_ARG; }
extern "C" int cb_precision(const cb_handle* h) { return h ? h->precision : CB_ERR_ARG; }
extern "C" size_t cb_workspace_bytes(const cb_handle* h) { return h ? h->ws_bytes + h->ws_bytes_tc + h->stage_bytes + h->beam_ws_bytes + h->asm_ws_bytes : 0; }
extern "C" long long cb_launch_count(const cb_handle* h) { return h ? h->launches : 0; }
extern "C" void cb_enable_timing(cb_handle* h, int on) { if (h) h->timing = on; }

extern "C" int cb_reserve_sms(cb_handle* h, int n) {
    if (!h || n < 0 || n > 16) { cb_set_error("cb_reserve_sms: n must be 0..16"); return CB_ERR_ARG; }
    h->reserve_sms = (n + 1) & ~1;
    return CB_OK;
}

extern "C" int cb_last_forward_ms(const cb_handle* h, float* ms, int n) {
    if (!h || !ms || !h->have_ms) return 0;
    // events were recorded on the stream of the last cb_forward; the caller has synchronised it
    float v[4] = {0, 0, 0, 0};
    if (cudaEventElapsedTime(&v[0], h->ev[0], h->ev[1]) != cudaSuccess) return 0;   // residual conv stack
    cudaEventElapsedTime(&v[1], h->ev[1], h->ev[2]);                                  // BiLSTM stack
    cudaEventElapsedTime(&v[2], h->ev[2], h->ev[3]);                                  // head + path_prob
    cudaEventElapsedTime(&v[3], h->ev[0], h->ev[3]);                                  // total
    const int m = n < 4 ? n : 4;
    for (int i = 0; i < m; ++i) ms[i] = v[i];
    return m;
}

extern "C" int cb_last_forward_profile(const cb_handle* h, float* ms, int* count, int n) {
    if (!h || !ms || !count || !h->have_ms) return 0;
    const int m = n < CB_CAT_COUNT ? n : CB_CAT_COUNT;
    for (int i = 0; i < m; ++i) { ms[i] = 0.f; count[i] = 0; }
    for (int i = 0; i < h->prof_n; ++i) {
        float t = 0.f;
        if (cudaEventElapsedTime(&t, h->prof_ev[i][0], h->prof_ev[i][1]) != cudaSuccess) return 0;
        if (h->prof_cat[i] < m) { ms[h->prof_cat[i]] += t; count[h->prof_cat[i]]++; }
    }
    static const bool dump = getenv("CB_PROF_DUMP") != nullptr;      // development: every timed launch of the forward, in order
    if (dump) {
        fprintf(stderr, "cb_forward launches (ms):");
        for (int i = 0; i < h->prof_n; ++i) {
            float t = 0.f;
            cudaEventElapsedTime(&t, h->prof_ev[i][0], h->prof_ev[i][1]);
            fprintf(stderr, " %d:%.3f", h->prof_cat[i], t);
        }
        fprintf(stderr, "\n");
    }
    return m;
}

// -----------------------------------------------------------------------------------------------------------------
static int ensure_workspace(cb_handle* h, int B, int L) {
    const CbConfig& c = h->cfg;
    const int T = out_len_of(c, L);
    const size_t act_floats = (size_t)B * L * c.channels;
    const size_t pre_floats = (size_t)B * T * 8 * c.hidden;
    const size_t out_floats = (size_t)B * T * 2 * c.hidden;
    const size_t need = (3 * align_up(act_floats, 64) + align_up(pre_floats, 64) + 2 * align_up(out_floats, 64)) * sizeof(float);
    if (need > h->ws_bytes) {
        if (h->ws) { cudaFree(h->ws); h->ws = nullptr; h->ws_bytes = 0; }
        cudaError_t e = cudaMalloc(&h->ws, need);
        if (e != cudaSuccess) { cb_set_error("workspace cudaMalloc(%zu bytes): %s", need, cudaGetErrorString(e)); return CB_ERR_NOMEM; }
        h->ws_bytes = need;
    }
    float* p = (float*)h->ws;
    for (int i = 0; i < 3; ++i) { h->act[i] = p; p += align_up(act_floats, 64); }
    h->pre = p; p += align_up(pre_floats, 64);
    for (int i = 0; i < 2; ++i) { h->lstm_out[i] = p; p += align_up(out_floats, 64); }
    return CB_OK;
}

int cb_prof_begin(cb_handle* h, int cat, cudaStream_t s) {
    if (!h->timing || h->prof_n >= CB_PROF_MAX) return -1;
    const int i = h->prof_n++;
    if (!h->prof_ev[i][0]) {
        if (cudaEventCreate(&h->prof_ev[i][0]) != cudaSuccess || cudaEventCreate(&h->prof_ev[i][1]) != cudaSuccess) {
            h->prof_n--; return -1;
        }
    }
    h->prof_cat[i] = cat;
    cudaEventRecord(h->prof_ev[i][0], s);
    return i;
}
void cb_prof_end(cb_handle* h, int i, cudaStream_t s) { if (i >= 0) cudaEventRecord(h->prof_ev[i][1], s); }

static int run_gemm(cb_handle* h, const GemmProblem& p, cudaStream_t s, int cat) {
    const int pi = cb_prof_begin(h, cat, s);
    const int rc = cb_launch_gemm_simt(h, p, s);
    cb_prof_end(h, pi, s);
    return rc;
}

// The two conv-stack orchestrations (cb_conv_stack.cuh) issue their work through this adapter: CUDA launches on stream s.
namespace {
struct CudaConvOps {
    cb_handle* h; cudaStream_t s;
    int gemm(const GemmProblem& g) { return run_gemm(h, g, s, CB_CAT_CONV); }
    int bn_rank1(const float* x, int B, int t_in, int stride, int t_out, const float* w, const float* scale,
                 const float* offset, float* inv, float* shift) {
        return cb_launch_bn_rank1(h, x, B, t_in, stride, t_out, w, scale, offset, inv, shift, s);
    }
    int bn_stats(const float* X, long long M, const float* scale, const float* offset, float* inv, float* shift) {
        return cb_launch_bn_stats(h, X, M, scale, offset, inv, shift, s);
    }
    int stem(const StemProblem& p) {
        const int pi = cb_prof_begin(h, CB_CAT_CONV, s);
        const int rc = cb_launch_stem(h, p, s);
        cb_prof_end(h, pi, s);
        return rc;
    }
    int bn_apply(const BnApplyArgs& a) { return cb_launch_bn_apply(h, a, s); }
};
}  // namespace

extern "C" int cb_seq_len_out(cb_handle* h, const int32_t* seq_len_in, int B, int L, int32_t* seq_len_out, void* stream) {
    if (!h || !seq_len_in || !seq_len_out || B < 0 || L < 1) { cb_set_error("cb_seq_len_out: bad arguments"); return CB_ERR_ARG; }
    CB_CUDA(cudaSetDevice(h->device));
    return cb_launch_seq_len(h, seq_len_in, B, L, out_len_of(h->cfg, L), seq_len_out, (cudaStream_t)stream);
}

extern "C" int cb_forward(cb_handle* h, const float* x, const int32_t* seq_len_out, int B, int L, float* logits,
                          float* path_prob, void* stream) {
    if (!h || !x || !seq_len_out || !logits || B < 0 || L < 1) { cb_set_error("cb_forward: bad arguments"); return CB_ERR_ARG; }
    if (B == 0) return CB_OK;
    if ((long long)B * L > 0x7fffffffLL / 2) { cb_set_error("cb_forward: B*L too large"); return CB_ERR_ARG; }
    CB_CUDA(cudaSetDevice(h->device));
    cudaStream_t s = (cudaStream_t)stream;
    if (h->precision != CB_PREC_FP32) {
        if (h->bn_mode != CB_BN_POPULATION) { cb_set_error("cb_forward: batch-statistics BatchNorm needs CB_PREC_FP32"); return CB_ERR_ARG; }
        return cb_forward_tc(h, x, seq_len_out, B, L, logits, path_prob, s);
    }
    int rc = ensure_workspace(h, B, L);
    if (rc != CB_OK) return rc;
    const CbConfig& c = h->cfg;
    const int C = c.channels, H = c.hidden;
    h->prof_n = 0;
    if (h->timing) CB_CUDA(cudaEventRecord(h->ev[0], s));

    // ---- residual conv stack (cnn.py:234-262, 380-389) ----------------------------------------------------------
    int t_in = L;
    const float* X = nullptr;
    {
        CudaConvOps ops{h, s};
        CbConvStackBufs bufs;
        for (int i = 0; i < 3; ++i) bufs.act[i] = h->act[i];
        for (int i = 0; i < CB_BN_VECS; ++i) bufs.vec[i] = h->bn_vec ? h->bn_vec + (size_t)i * C : nullptr;
        bufs.zeros = h->zeros;
        if (h->bn_mode == CB_BN_BATCH)
            rc = cb_conv_stack_batch_bn(ops, c, h->raw1, h->raw2a, h->raw2b, h->raw2c, &h->stem, bufs, x, B, L, &X, &t_in);
        else
            rc = cb_conv_stack_folded(ops, c, h->conv2a, h->conv2b, h->convc, h->g_w, h->g_inv, h->g_sh, h->r_w, h->r_inv,
                                      h->r_sh, &h->stem, bufs, x, B, L, &X, &t_in);
        if (rc != CB_OK) return rc;
    }
    const int T = t_in;
    const int M = B * T;
    h->fea = X;
    if (h->timing) CB_CUDA(cudaEventRecord(h->ev[1], s));

    // ---- BiLSTM stack (rnn.py:20-64 stacked-bidirectional; rnn.py:99-145 per-direction MultiRNNCell) ---------------
    const float* Z = X; int ldz = C;
    const int G = c.cell_type == CB_CELL_GRU ? 3 : 4;          // pre-activation columns per direction: G*H
    for (int l = 0; l < c.n_layers; ++l) {
        GemmProblem g;
        const int n_gemm = (l == 0 || c.rnn_layout == 0) ? 1 : 2;
        for (int d = 0; d < n_gemm; ++d) {
            memset(&g, 0, sizeof(g));
            const int in = l == 0 ? C : (c.rnn_layout == 0 ? 2 * H : H);
            g.M = M; g.N = n_gemm == 1 ? 2 * G * H : G * H; g.K = in; g.t_out = T; g.t_in0 = T; g.stride0 = 1; g.taps = 1; g.c0 = in;
            g.W = n_gemm == 1 ? h->wxcat[l] : h->wx[l][d];
            g.shift = n_gemm == 1 ? h->bcat[l] : h->bias[l][d];
            g.src0 = Z + (l == 0 ? 0 : d * H); g.lda0 = ldz; g.out = h->pre + d * G * H; g.ldo = 2 * G * H;
            if ((rc = run_gemm(h, g, s, CB_CAT_LSTM_IN)) != CB_OK) return rc;
        }
        const int pi = cb_prof_begin(h, CB_CAT_LSTM_REC, s);
        if (c.cell_type == CB_CELL_GRU) {
            GruProblem gp;
            memset(&gp, 0, sizeof(gp));
            gp.B = B; gp.T = T; gp.H = H; gp.pre = h->pre; gp.ld_pre = 2 * G * H;
            for (int d = 0; d < 2; ++d) { gp.wg[d] = h->whh[l][d]; gp.wc[d] = h->gru_wc[l][d]; }
            gp.lens = seq_len_out; gp.out = h->lstm_out[l & 1]; gp.ldo = 2 * H; gp.layer = l;
            rc = cb_launch_gru_simt(h, gp, s);
        } else {
            LstmProblem lp;
            memset(&lp, 0, sizeof(lp));
            lp.B = B; lp.T = T; lp.H = H; lp.pre = h->pre; lp.ld_pre = 2 * G * H;
            lp.whh[0] = h->whh[l][0]; lp.whh[1] = h->whh[l][1];
            lp.lens = seq_len_out; lp.out = h->lstm_out[l & 1]; lp.ldo = 2 * H; lp.layer = l;
            rc = cb_launch_lstm_simt(h, lp, s);
        }
        cb_prof_end(h, pi, s);
        if (rc != CB_OK) return rc;
        Z = h->lstm_out[l & 1]; ldz = 2 * H;
    }
    if (h->timing) CB_CUDA(cudaEventRecord(h->ev[2], s));

    // ---- head + path_prob ----------------------------------------------------------------------------------------
    {
        const int pi = cb_prof_begin(h, CB_CAT_HEAD, s);
        rc = cb_launch_head(h, Z, M, logits, s);
        cb_prof_end(h, pi, s);
        if (rc != CB_OK) return rc;
    }
    h->last_Bp = B; h->last_tmajor = 0;
    if (path_prob && (rc = cb_launch_path_prob(h, logits, B, T, path_prob, s)) != CB_OK) return rc;
    if (h->timing) { CB_CUDA(cudaEventRecord(h->ev[3], s)); h->have_ms = 1; }
    h->last_B = B; h->last_T = T;
    return CB_OK;
}

extern "C" int cb_decode_greedy(cb_handle* h, const float* logits, const int32_t* seq_len_out, int B, int T,
                                int8_t* bases, int32_t* n_bases, void* stream) {
    if (!h || !logits || !seq_len_out || !bases || !n_bases || B < 0 || T < 1) { cb_set_error("cb_decode_greedy: bad arguments"); return CB_ERR_ARG; }
    CB_CUDA(cudaSetDevice(h->device));
    return cb_launch_greedy(h, logits, seq_len_out, B, T, bases, n_bases, (cudaStream_t)stream);
}

extern "C" int cb_decode_beam(cb_handle* h, const float* logits, const int32_t* seq_len_out, int B, int T, int beam_width,
                              int8_t* bases, int32_t* n_bases, void* stream) {
    if (!h || !logits || !seq_len_out || !bases || !n_bases || B < 0 || T < 1 || beam_width < 1) { cb_set_error("cb_decode_beam: bad arguments"); return CB_ERR_ARG; }
    CB_CUDA(cudaSetDevice(h->device));
    return cb_launch_beam(h, logits, seq_len_out, B, T, beam_width, bases, n_bases, nullptr, (cudaStream_t)stream);
}

extern "C" int cb_decode_beam_scored(cb_handle* h, const float* logits, const int32_t* seq_len_out, int B, int T, int beam_width,
                                     int8_t* bases, int32_t* n_bases, float* log_prob, void* stream) {
    if (!h || !logits || !seq_len_out || !bases || !n_bases || !log_prob || B < 0 || T < 1 || beam_width < 1) { cb_set_error("cb_decode_beam_scored: bad arguments"); return CB_ERR_ARG; }
    CB_CUDA(cudaSetDevice(h->device));
    return cb_launch_beam(h, logits, seq_len_out, B, T, beam_width, bases, n_bases, log_prob, (cudaStream_t)stream);
}

extern "C" int cb_assemble(cb_handle* h, const int8_t* bases, const int32_t* n_bases, const float* path_prob,
                           int n_windows, int T, int jump, int L, int kernel, int8_t* consensus, char* qual,
                           int32_t* pos, int32_t* out_len, int max_len, void* stream) {
    if (!h || !bases || !n_bases || !consensus || !pos || !out_len || n_windows < 0 || T < 1 || L < 1 || max_len < 0 ||
        kernel < CB_ASM_SIMPLE || kernel > CB_ASM_STICK || (qual && !path_prob)) {
        cb_set_error("cb_assemble: bad arguments");
        return CB_ERR_ARG;
    }
    CB_CUDA(cudaSetDevice(h->device));
    return cb_launch_assemble(h, bases, n_bases, path_prob, n_windows, T, jump, L, kernel, consensus, qual, pos, out_len,
                              max_len, (cudaStream_t)stream);
}

// ---- host-buffer wrappers ----------------------------------------------------------------------------------------
static int ensure_stage(cb_handle* h, size_t bytes) {
    if (bytes <= h->stage_bytes) return CB_OK;
    if (h->stage) { cudaFree(h->stage); h->stage = nullptr; h->stage_bytes = 0; }
    cudaError_t e = cudaMalloc(&h->stage, bytes);
    if (e != cudaSuccess) { cb_set_error("staging cudaMalloc(%zu bytes): %s", bytes, cudaGetErrorString(e)); return CB_ERR_NOMEM; }
    h->stage_bytes = bytes;
    return CB_OK;
}

extern "C" int cb_basecall_host(cb_handle* h, const float* x, const int32_t* seq_len_in, int B, int L, int beam_width,
                                int8_t* bases, int32_t* n_bases, float* path_prob, float* logits) {
    if (!h || !x || !seq_len_in || !bases || !n_bases || B < 0 || L < 1 || beam_width < 0) { cb_set_error("cb_basecall_host: bad arguments"); return CB_ERR_ARG; }
    if (B == 0) return CB_OK;
    CB_CUDA(cudaSetDevice(h->device));
    const int T = out_len_of(h->cfg, L), C = h->cfg.n_class;
    const size_t sz_x = align_up((size_t)B * L * 4, 256), sz_i = align_up((size_t)B * 4, 256);
    const size_t sz_lg = align_up((size_t)B * T * C * 4, 256), sz_b = align_up((size_t)B * T, 256);
    int rc = ensure_stage(h, sz_x + 4 * sz_i + sz_lg + sz_b);
    if (rc != CB_OK) return rc;
    char* p = (char*)h->stage;
    float* d_x = (float*)p; p += sz_x;
    int32_t* d_in = (int32_t*)p; p += sz_i;
    int32_t* d_len = (int32_t*)p; p += sz_i;
    int32_t* d_nb = (int32_t*)p; p += sz_i;
    float* d_prob = (float*)p; p += sz_i;
    float* d_lg = (float*)p; p += sz_lg;
    int8_t* d_bases = (int8_t*)p;
    cudaStream_t s = 0;
    CB_CUDA(cudaMemcpyAsync(d_x, x, (size_t)B * L * 4, cudaMemcpyHostToDevice, s));
    CB_CUDA(cudaMemcpyAsync(d_in, seq_len_in, (size_t)B * 4, cudaMemcpyHostToDevice, s));
    if ((rc = cb_launch_seq_len(h, d_in, B, L, T, d_len, s)) != CB_OK) return rc;
    if ((rc = cb_forward(h, d_x, d_len, B, L, d_lg, d_prob, s)) != CB_OK) return rc;
    if (beam_width == 0) rc = cb_launch_greedy(h, d_lg, d_len, B, T, d_bases, d_nb, s);
    else rc = cb_launch_beam(h, d_lg, d_len, B, T, beam_width, d_bases, d_nb, nullptr, s);
    if (rc != CB_OK) return rc;
    CB_CUDA(cudaMemcpyAsync(bases, d_bases, (size_t)B * T, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaMemcpyAsync(n_bases, d_nb, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    if (path_prob) CB_CUDA(cudaMemcpyAsync(path_prob, d_prob, (size_t)B * 4, cudaMemcpyDeviceToHost, s));
    if (logits) CB_CUDA(cudaMemcpyAsync(logits, d_lg, (size_t)B * T * C * 4, cudaMemcpyDeviceToHost, s));
    CB_CUDA(cudaStreamSynchronize(s));
    return cb_check_deferred(h, s);
}

// ---- two-slot asynchronous host pipeline -----------------------------------------------------------------------------
static int pipe_init(cb_handle* h) {
    if (h->pipe_compute) return CB_OK;
    CB_CUDA(cudaStreamCreateWithFlags(&h->pipe_in, cudaStreamNonBlocking));
    CB_CUDA(cudaStreamCreateWithFlags(&h->pipe_compute, cudaStreamNonBlocking));
    CB_CUDA(cudaStreamCreateWithFlags(&h->pipe_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2; ++i) {
        CB_CUDA(cudaEventCreateWithFlags(&h->pipe[i].h2d_done, cudaEventDisableTiming));
        CB_CUDA(cudaEventCreateWithFlags(&h->pipe[i].compute_done, cudaEventDisableTiming));
        CB_CUDA(cudaEventCreateWithFlags(&h->pipe[i].d2h_done, cudaEventDisableTiming));
    }
    return CB_OK;
}

struct PipeLayout {              // offsets into the pinned and the device buffer of a slot
    size_t p_x, p_len, p_bases, p_nb, p_prob, p_total;
    size_t d_x, d_in, d_len, d_nb, d_prob, d_lg, d_bases, d_total;
};
static PipeLayout pipe_layout(int B, int L, int T, int C) {
    PipeLayout o;
    const size_t sz_x = align_up((size_t)B * L * 4, 256), sz_i = align_up((size_t)B * 4, 256);
    const size_t sz_lg = align_up((size_t)B * T * C * 4, 256), sz_b = align_up((size_t)B * T, 256);
    o.p_x = 0; o.p_len = sz_x; o.p_bases = o.p_len + sz_i; o.p_nb = o.p_bases + sz_b; o.p_prob = o.p_nb + sz_i;
    o.p_total = o.p_prob + sz_i;
    o.d_x = 0; o.d_in = sz_x; o.d_len = o.d_in + sz_i; o.d_nb = o.d_len + sz_i; o.d_prob = o.d_nb + sz_i;
    o.d_lg = o.d_prob + sz_i; o.d_bases = o.d_lg + sz_lg; o.d_total = o.d_bases + sz_b;
    return o;
}

extern "C" int cb_basecall_submit(cb_handle* h, int slot, const float* x, const int32_t* seq_len_in, int B, int L,
                                  int beam_width) {
    if (!h || slot < 0 || slot > 1 || !x || !seq_len_in || B < 1 || L < 1 || beam_width < 0) { cb_set_error("cb_basecall_submit: bad arguments"); return CB_ERR_ARG; }
    CB_CUDA(cudaSetDevice(h->device));
    int rc = pipe_init(h);
    if (rc != CB_OK) return rc;
    cb_handle::PipeSlot& sl = h->pipe[slot];
    if (sl.busy) { cb_set_error("cb_basecall_submit: slot %d has not been collected", slot); return CB_ERR_ARG; }
    const int T = out_len_of(h->cfg, L), C = h->cfg.n_class;
    const PipeLayout o = pipe_layout(B, L, T, C);
    if (o.p_total > sl.pin_bytes) {
        if (sl.pin) { cudaFreeHost(sl.pin); sl.pin = nullptr; sl.pin_bytes = 0; }
        cudaError_t e = cudaMallocHost(&sl.pin, o.p_total);
        if (e != cudaSuccess) { cb_set_error("pinned staging cudaMallocHost(%zu bytes): %s", o.p_total, cudaGetErrorString(e)); return CB_ERR_NOMEM; }
        sl.pin_bytes = o.p_total;
    }
    if (o.d_total > sl.dev_bytes) {
        if (sl.dev) { cudaFree(sl.dev); sl.dev = nullptr; sl.dev_bytes = 0; }
        cudaError_t e = cudaMalloc(&sl.dev, o.d_total);
        if (e != cudaSuccess) { cb_set_error("pipeline staging cudaMalloc(%zu bytes): %s", o.d_total, cudaGetErrorString(e)); return CB_ERR_NOMEM; }
        sl.dev_bytes = o.d_total;
    }
    char* pin = (char*)sl.pin; char* dev = (char*)sl.dev;
    memcpy(pin + o.p_x, x, (size_t)B * L * 4);
    memcpy(pin + o.p_len, seq_len_in, (size_t)B * 4);
    sl.B = B; sl.L = L; sl.T = T;
    CB_CUDA(cudaMemcpyAsync(dev + o.d_x, pin + o.p_x, (size_t)B * L * 4, cudaMemcpyHostToDevice, h->pipe_in));
    CB_CUDA(cudaMemcpyAsync(dev + o.d_in, pin + o.p_len, (size_t)B * 4, cudaMemcpyHostToDevice, h->pipe_in));
    CB_CUDA(cudaEventRecord(sl.h2d_done, h->pipe_in));
    cudaStream_t s = h->pipe_compute;
    CB_CUDA(cudaStreamWaitEvent(s, sl.h2d_done, 0));
    int32_t* d_len = (int32_t*)(dev + o.d_len);
    float* d_lg = (float*)(dev + o.d_lg);
    if ((rc = cb_launch_seq_len(h, (const int32_t*)(dev + o.d_in), B, L, T, d_len, s)) != CB_OK) return rc;
    if ((rc = cb_forward(h, (const float*)(dev + o.d_x), d_len, B, L, d_lg, (float*)(dev + o.d_prob), s)) != CB_OK) return rc;
    if (beam_width == 0) rc = cb_launch_greedy(h, d_lg, d_len, B, T, (int8_t*)(dev + o.d_bases), (int32_t*)(dev + o.d_nb), s);
    else rc = cb_launch_beam(h, d_lg, d_len, B, T, beam_width, (int8_t*)(dev + o.d_bases), (int32_t*)(dev + o.d_nb), nullptr, s);
    if (rc != CB_OK) return rc;
    CB_CUDA(cudaEventRecord(sl.compute_done, s));
    CB_CUDA(cudaStreamWaitEvent(h->pipe_out, sl.compute_done, 0));
    CB_CUDA(cudaMemcpyAsync(pin + o.p_bases, dev + o.d_bases, (size_t)B * T, cudaMemcpyDeviceToHost, h->pipe_out));
    CB_CUDA(cudaMemcpyAsync(pin + o.p_nb, dev + o.d_nb, (size_t)B * 4, cudaMemcpyDeviceToHost, h->pipe_out));
    CB_CUDA(cudaMemcpyAsync(pin + o.p_prob, dev + o.d_prob, (size_t)B * 4, cudaMemcpyDeviceToHost, h->pipe_out));
    CB_CUDA(cudaEventRecord(sl.d2h_done, h->pipe_out));
    sl.busy = 1;
    return CB_OK;
}

extern "C" int cb_basecall_collect(cb_handle* h, int slot, int8_t* bases, int32_t* n_bases, float* path_prob) {
    if (!h || slot < 0 || slot > 1 || !bases || !n_bases) { cb_set_error("cb_basecall_collect: bad arguments"); return CB_ERR_ARG; }
    cb_handle::PipeSlot& sl = h->pipe[slot];
    if (!sl.busy) { cb_set_error("cb_basecall_collect: slot %d holds no batch", slot); return CB_ERR_ARG; }
    CB_CUDA(cudaSetDevice(h->device));
    sl.busy = 0;
    CB_CUDA(cudaEventSynchronize(sl.d2h_done));
    const PipeLayout o = pipe_layout(sl.B, sl.L, sl.T, h->cfg.n_class);
    const char* pin = (const char*)sl.pin;
    memcpy(bases, pin + o.p_bases, (size_t)sl.B * sl.T);
    memcpy(n_bases, pin + o.p_nb, (size_t)sl.B * 4);
    if (path_prob) memcpy(path_prob, pin + o.p_prob, (size_t)sl.B * 4);
    return cb_check_deferred(h, h->pipe_out);
}

extern "C" int cb_check_status(cb_handle* h, void* stream) {
    if (!h) { cb_set_error("cb_check_status: bad arguments"); return CB_ERR_ARG; }
    CB_CUDA(cudaSetDevice(h->device));
    CB_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    return cb_check_deferred(h, (cudaStream_t)stream);
}

extern "C" int cb_assemble_host(cb_handle* h, const int8_t* bases, const int32_t* n_bases, const float* path_prob,
                                int n_windows, int T, int jump, int L, int kernel, int8_t* consensus, char* qual,
                                int32_t* pos, int32_t* out_len, int max_len) {
    if (!h || !bases || !n_bases || !consensus || !pos || !out_len || n_windows < 0 || T < 1 || max_len < 0) { cb_set_error("cb_assemble_host: bad arguments"); return CB_ERR_ARG; }
    CB_CUDA(cudaSetDevice(h->device));
    const size_t sz_b = align_up((size_t)n_windows * T + 1, 256), sz_i = align_up((size_t)n_windows * 4 + 4, 256);
    const size_t sz_c = align_up((size_t)max_len + 1, 256);
    const size_t need = sz_b + 3 * sz_i + 2 * sz_c + 256;
    if (need > h->asm_stage_bytes) {            // grow-only: a cudaFree per read would synchronise the whole device
        if (h->asm_stage) { cudaFree(h->asm_stage); h->asm_stage = nullptr; h->asm_stage_bytes = 0; }
        cudaError_t em = cudaMalloc(&h->asm_stage, need + need / 2);
        if (em != cudaSuccess) { cb_set_error("cb_assemble_host: cudaMalloc(%zu bytes): %s", need + need / 2, cudaGetErrorString(em)); return CB_ERR_NOMEM; }
        h->asm_stage_bytes = need + need / 2;
    }
    char* buf = (char*)h->asm_stage;
    char* p = buf;
    int8_t* d_bases = (int8_t*)p; p += sz_b;
    int32_t* d_nb = (int32_t*)p; p += sz_i;
    float* d_pp = (float*)p; p += sz_i;
    int32_t* d_pos = (int32_t*)p; p += sz_i;
    int8_t* d_cons = (int8_t*)p; p += sz_c;
    char* d_qual = (char*)p; p += sz_c;
    int32_t* d_len = (int32_t*)p;
    // a stream of its own, non-blocking: the legacy default stream would serialise with every blocking stream of the caller
    if (!h->asm_stream) CB_CUDA(cudaStreamCreateWithFlags(&h->asm_stream, cudaStreamNonBlocking));
    cudaStream_t s = h->asm_stream;
    int rc = CB_OK;
    cudaError_t e = cudaSuccess;
    if (n_windows > 0) {
        e = cudaMemcpyAsync(d_bases, bases, (size_t)n_windows * T, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess) e = cudaMemcpyAsync(d_nb, n_bases, (size_t)n_windows * 4, cudaMemcpyHostToDevice, s);
        if (e == cudaSuccess && path_prob) e = cudaMemcpyAsync(d_pp, path_prob, (size_t)n_windows * 4, cudaMemcpyHostToDevice, s);
    }
    if (e == cudaSuccess)
        rc = cb_assemble(h, d_bases, d_nb, path_prob ? d_pp : nullptr, n_windows, T, jump, L, kernel, d_cons,
                         qual ? d_qual : nullptr, d_pos, d_len, max_len, s);
    if (rc == CB_OK && e == cudaSuccess) {
        e = cudaMemcpyAsync(out_len, d_len, 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && n_windows > 0) e = cudaMemcpyAsync(pos, d_pos, (size_t)n_windows * 4, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && max_len > 0) e = cudaMemcpyAsync(consensus, d_cons, max_len, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess && qual && max_len > 0) e = cudaMemcpyAsync(qual, d_qual, max_len, cudaMemcpyDeviceToHost, s);
        if (e == cudaSuccess) e = cudaStreamSynchronize(s);
    }
    if (e != cudaSuccess) { cb_set_error("cb_assemble_host: %s", cudaGetErrorString(e)); return CB_ERR_CUDA; }
    return rc;
}

extern "C" long long cb_debug_fetch(cb_handle* h, int what, float* dst, size_t max_floats) {
    if (!h || !dst || !h->last_B) { cb_set_error("cb_debug_fetch: nothing to fetch"); return CB_ERR_ARG; }
    if (h->precision != CB_PREC_FP32) { CB_CUDA(cudaSetDevice(h->device)); return cb_debug_fetch_tc(h, what, dst, max_floats); }
    const size_t M = (size_t)h->last_B * h->last_T;
    const float* src; size_t n;
    if (what == 0) { src = h->fea; n = M * h->cfg.channels; }
    else if (what >= 1 && what <= h->cfg.n_layers) { src = h->lstm_out[(what - 1) & 1]; n = M * 2 * h->cfg.hidden; }
    else { cb_set_error("cb_debug_fetch: unknown tensor %d", what); return CB_ERR_ARG; }
    if (n > max_floats) { cb_set_error("cb_debug_fetch: destination too small"); return CB_ERR_ARG; }
    CB_CUDA(cudaSetDevice(h->device));
    CB_CUDA(cudaDeviceSynchronize());
    CB_CUDA(cudaMemcpy(dst, src, n * sizeof(float), cudaMemcpyDeviceToHost));
    return (long long)n;
}

extern "C" void* cb_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cb_set_error("cudaHostAlloc(%zu) failed", bytes); return nullptr; }
    return p;
}

extern "C" void cb_host_free(void* p) { if (p) cudaFreeHost(p); }
