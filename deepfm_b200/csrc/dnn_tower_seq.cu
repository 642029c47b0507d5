// The whole DNN tower as two C-ABI calls (reference: deepfm/models/layers/dnn.py:45-59, the nn.Sequential of
// Linear -> BatchNorm1d -> activation -> Dropout blocks, and its autograd).  Nothing new is computed here: the functions
// below sequence the per-block entry points (dfm_gemm3, dfm_bn_stats, dfm_bn_act_fwd / _bwd) on the caller's stream.  They
// exist because the host side of a multi-rank step was the bottleneck: three autograd nodes with ~45 ctypes calls and
// ~50 tensor allocations per step became one node with two calls and a handful of allocations.
#include <stdint.h>
#include <string.h>

#include "common.cuh"
#include "../../include/deepfm_b200.h"

using namespace dfm;

namespace {

constexpr int MAX_LAYERS = 16;

inline size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }

// workspace shared by every block: the largest GEMM workspace of any of its three products + the column-sum partials
size_t block_ws(int64_t M, int64_t din, int64_t dout, bool backward) {
    size_t g = dfm_gemm3_workspace_bytes(0, M, dout, din);
    if (backward) {
        size_t g1 = dfm_gemm3_workspace_bytes(1, M, din, dout), g2 = dfm_gemm3_workspace_bytes(2, dout, din, M);
        if (g1 > g) g = g1;
        if (g2 > g) g = g2;
    }
    return al256(g) + al256(dfm_tower_workspace_bytes(M, (int)dout));
}

}  // namespace

extern "C" {

// Floats of the activation store of dfm_tower_fwd: per block [y (M x h)][a (M x h), not for the last block][mean h][rstd h],
// every piece 256-byte aligned.  offsets (4 per block, in floats: y, a, mean, rstd; a = -1 for the last block) are optional.
size_t dfm_tower_store_floats(int n_layers, const int64_t* dims, int64_t M, int64_t* offsets) {
    size_t off = 0;
    for (int l = 0; l < n_layers; ++l) {
        const size_t h = (size_t)dims[l + 1];
        auto take = [&](size_t n) { const size_t o = off; off += (n + 63) & ~(size_t)63; return o; };
        const size_t oy = take((size_t)M * h);
        const long long oa = l + 1 < n_layers ? (long long)take((size_t)M * h) : -1;
        const size_t om = take(h), orr = take(h);
        if (offsets) { offsets[4 * l] = (int64_t)oy; offsets[4 * l + 1] = oa; offsets[4 * l + 2] = (int64_t)om; offsets[4 * l + 3] = (int64_t)orr; }
    }
    return off;
}

size_t dfm_tower_seq_workspace_bytes(int n_layers, const int64_t* dims, int64_t M, int backward) {
    size_t ws = 0;
    int64_t hmax = 0;
    for (int l = 0; l < n_layers; ++l) {
        const size_t b = block_ws(M, dims[l], dims[l + 1], backward != 0);
        if (b > ws) ws = b;
        if (dims[l + 1] > hmax) hmax = dims[l + 1];
    }
    if (backward) ws += 3 * al256((size_t)M * hmax * 4);      // dy + two ping-pong input-gradient buffers of the inner blocks
    return ws + 256;
}

// params: 4 pointers per block (W (h, din), bias (h) or NULL, gamma (h) or NULL, beta (h) or NULL)
// bn: per block 0 none / 1 batch statistics / 2 fixed (running) statistics;  running: 2 pointers per block (running_mean,
// running_var; updated with `momentum[l]` in mode 1 when not NULL; the statistics used in mode 2)
// seeds: one dropout seed per block.  out: (M, dims[n]) the tower output.  store: dfm_tower_store_floats() floats.
int dfm_tower_fwd(int n_layers, const int64_t* dims, int64_t M, const float* x, const float* const* params, const int* bn,
                  float* const* running, const float* eps, const float* momentum, int act, float drop_p, const uint64_t* seeds,
                  float* store, float* out, void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(n_layers >= 1 && n_layers <= MAX_LAYERS && dims && M > 0 && x && params && bn && store && out && workspace,
                DFM_ERR_INVALID, "dfm_tower_fwd: bad argument (1..%d blocks)", MAX_LAYERS);
    DFM_REQUIRE(workspace_bytes >= dfm_tower_seq_workspace_bytes(n_layers, dims, M, 0), DFM_ERR_WORKSPACE, "dfm_tower_fwd: workspace too small");
    int64_t offs[4 * MAX_LAYERS];
    dfm_tower_store_floats(n_layers, dims, M, offs);
    char* ws = static_cast<char*>(workspace);
    const float* in = x;
    for (int l = 0; l < n_layers; ++l) {
        const int64_t din = dims[l], h = dims[l + 1];
        const float* W = params[4 * l];
        const float* bias = params[4 * l + 1];
        const float* gamma = params[4 * l + 2];
        const float* beta = params[4 * l + 3];
        DFM_REQUIRE(W, DFM_ERR_INVALID, "dfm_tower_fwd: block %d has no weight", l);
        float* y = store + offs[4 * l];
        float* a = offs[4 * l + 1] >= 0 ? store + offs[4 * l + 1] : out;
        float* mean = store + offs[4 * l + 2];
        float* rstd = store + offs[4 * l + 3];
        const size_t gws = al256(dfm_gemm3_workspace_bytes(0, M, h, din));
        int rc = dfm_gemm3(0, in, W, y, bias, M, h, din, ws, gws, stream);
        if (rc) return rc;
        const float* use_mean = nullptr;
        const float* use_rstd = nullptr;
        if (bn[l] == 1) {
            rc = dfm_bn_stats(y, M, (int)h, eps ? eps[l] : 1e-5f, mean, rstd, running ? running[2 * l] : nullptr,
                              running ? running[2 * l + 1] : nullptr, momentum ? momentum[l] : 0.1f, ws + gws,
                              workspace_bytes - gws, stream);
            if (rc) return rc;
            use_mean = mean; use_rstd = rstd;
        } else if (bn[l] == 2) {
            DFM_REQUIRE(running && running[2 * l] && running[2 * l + 1], DFM_ERR_INVALID, "dfm_tower_fwd: block %d needs fixed statistics", l);
            use_mean = running[2 * l]; use_rstd = running[2 * l + 1];      // caller passes (mean, rstd) of the running statistics
        }
        rc = dfm_bn_act_fwd(y, M, (int)h, bn[l], act, use_mean, use_rstd, gamma, beta, drop_p, seeds ? seeds[l] : 0ull, a, stream);
        if (rc) return rc;
        in = a;
    }
    return DFM_OK;
}

// grads: 4 pointers per block (dW (h, din), dbias (h) or NULL, dgamma, dbeta (NULL without BatchNorm)); dx (M, dims[0]) or NULL.
// fixed: for bn mode 2 blocks the (mean, rstd) pair the forward used (2 pointers per block), else ignored.
int dfm_tower_bwd(int n_layers, const int64_t* dims, int64_t M, const float* x, const float* const* params, const int* bn,
                  const float* const* fixed, int act, float drop_p, const uint64_t* seeds, const float* store, const float* g_out,
                  float* dx, float* const* grads, void* workspace, size_t workspace_bytes, void* stream) {
    DFM_REQUIRE(n_layers >= 1 && n_layers <= MAX_LAYERS && dims && M > 0 && x && params && bn && store && g_out && grads && workspace,
                DFM_ERR_INVALID, "dfm_tower_bwd: bad argument (1..%d blocks)", MAX_LAYERS);
    DFM_REQUIRE(workspace_bytes >= dfm_tower_seq_workspace_bytes(n_layers, dims, M, 1), DFM_ERR_WORKSPACE, "dfm_tower_bwd: workspace too small");
    int64_t offs[4 * MAX_LAYERS];
    dfm_tower_store_floats(n_layers, dims, M, offs);
    int64_t hmax = 0;
    size_t bws = 0;
    for (int l = 0; l < n_layers; ++l) {
        if (dims[l + 1] > hmax) hmax = dims[l + 1];
        const size_t b = block_ws(M, dims[l], dims[l + 1], true);
        if (b > bws) bws = b;
    }
    char* ws = static_cast<char*>(workspace);
    const size_t act_bytes = al256((size_t)M * hmax * 4);
    float* dy = reinterpret_cast<float*>(ws + bws);
    float* ping[2] = {reinterpret_cast<float*>(ws + bws + act_bytes), reinterpret_cast<float*>(ws + bws + 2 * act_bytes)};
    const float* g = g_out;
    for (int l = n_layers - 1; l >= 0; --l) {
        const int64_t din = dims[l], h = dims[l + 1];
        const float* W = params[4 * l];
        const float* gamma = params[4 * l + 2];
        const float* beta = params[4 * l + 3];
        const float* y = store + offs[4 * l];
        const float* in = l > 0 ? store + offs[4 * (l - 1) + 1] : x;          // the block's input = previous block's output
        const float* mean = bn[l] == 1 ? store + offs[4 * l + 2] : (bn[l] == 2 && fixed ? fixed[2 * l] : nullptr);
        const float* rstd = bn[l] == 1 ? store + offs[4 * l + 3] : (bn[l] == 2 && fixed ? fixed[2 * l + 1] : nullptr);
        float* dW = grads[4 * l];
        DFM_REQUIRE(dW, DFM_ERR_INVALID, "dfm_tower_bwd: block %d has no weight-gradient buffer", l);
        const size_t tws_off = bws - al256(dfm_tower_workspace_bytes(M, (int)h));
        int rc = dfm_bn_act_bwd(g, y, M, (int)h, bn[l], act, mean, rstd, gamma, beta, drop_p, seeds ? seeds[l] : 0ull, dy,
                                grads[4 * l + 2], grads[4 * l + 3], grads[4 * l + 1], ws + tws_off, bws - tws_off, stream);
        if (rc) return rc;
        rc = dfm_gemm3(2, dy, in, dW, nullptr, h, din, M, ws, tws_off, stream);
        if (rc) return rc;
        float* gin = l > 0 ? ping[l & 1] : dx;
        if (gin) {
            rc = dfm_gemm3(1, dy, W, gin, nullptr, M, din, h, ws, tws_off, stream);
            if (rc) return rc;
        }
        g = gin;
    }
    return DFM_OK;
}

}  // extern "C"
