// Host-side embedding plan and its by-value device image (kernel parameter).
#pragma once

#include <vector>

#include "common.cuh"

namespace dfm {

constexpr int MAX_FIELDS = 128;
constexpr int MAX_SLOTS = 1024;

struct FieldDev {
    const void* in;     // int64 ids (B,) / (B, max_len), or float x (B,)
    const float* w2;    // (V, d) table or Linear(1, d) weight
    const float* b2;    // DENSE only
    const float* w1;    // (V, 1) table or Linear(1, 1) weight
    const float* b1;    // DENSE only
    const float* proj;  // (D, d) or nullptr
    long long row_base; // first global row of this field's table
    int kind, dim, flat_off, max_len, combiner, slot_base, aux_off, vocab;
    int row_stride, w1_stride;   // floats between consecutive rows of the id table / first-order table
    int foreign;                 // 1: the table's gradient is produced elsewhere (no sort key, no table grad here)
};

struct GradDev {
    float *gw2, *gb2, *gw1, *gb1, *gproj;
};

// A run of consecutive fields of one class (forward kernel): 0 = plain SPARSE of dim D (no
// projection), 1 = plain DENSE of dim D, 2 = anything else (sequence bags, projected fields).
struct FieldRun {
    short cls, f0, n, x0;   // x0: index of the run's first DENSE field among the DENSE fields
};

struct DevPlan {
    int n_fields, D, T, S, A, aliased, max_tdim;
    unsigned pad_key;
    int n_runs, n_dense;
    FieldDev f[MAX_FIELDS];
    FieldRun runs[MAX_FIELDS];
    unsigned short slot_field[MAX_SLOTS];
    unsigned short slot_pos[MAX_SLOTS];
    unsigned short dense_field[MAX_FIELDS];   // field index of the i-th DENSE field
};

struct DevGrads {
    GradDev g[MAX_FIELDS];
};

}  // namespace dfm

struct dfm_plan {
    int n_fields = 0, fm_dim = 0;
    std::vector<int> kind, dim, max_len, combiner, flat_off, slot_base, aux_off;
    std::vector<long long> vocab, row_base;
    std::vector<int> slot_field, slot_pos;
    int T = 0, S = 0, A = 0, aliasable = 0, max_tdim = 0, key_bits = 0, vec = 1;
    long long total_rows = 0;
    int n_proj_expected = 0;
    std::vector<int> row_stride, w1_stride, foreign;   // per field, see dfm_plan_set_field_source

    // row_base / total_rows / key_bits over the fields whose gradient is produced here (foreign fields own no keys)
    void recompute_rows() {
        long long rows = 0;
        for (int f = 0; f < n_fields; ++f) {
            row_base[f] = rows;
            if (kind[f] != DFM_DENSE && !foreign[f]) rows += vocab[f];
        }
        row_base[n_fields] = rows;
        total_rows = rows;
        int bits = 1;
        while ((1ULL << bits) <= (unsigned long long)rows) ++bits;   // PAD key == rows must sort last
        key_bits = bits;
    }

    // Fill the per-call device image.  Returns the vector width usable for this call
    // (4 only if every dimension is a multiple of 4 and every pointer is 16-byte aligned).
    int fill(dfm::DevPlan& P, const void* const* inputs, const float* const* params,
             bool need_inputs) const;
};
