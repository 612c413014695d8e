/* TEST INFRASTRUCTURE (oracle) — exact CPU restatement of the scoring/masking/top-k contract.
 *
 * Follows: scores = u_emb @ items.T          reference code/model.py:122
 *          rating[exclude] = -(1<<10)        reference code/Procedure.py:177-181
 *          torch.topk(rating, k)             reference code/Procedure.py:183
 * with the arithmetic fixed the way the B200 kernel documents it (csrc/score_topk.cu): each score is
 * one fp32 FMA chain over k = 0..d-1 starting from 0; ties are broken towards the lowest item id.
 * fmaf() is exact (single rounding) whether or not the CPU has an FMA unit.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>

static int row_has(const int32_t* idx, int lo, int hi, int key) {
    int l = lo, h = hi;
    while (l < h) { int mid = (l + h) >> 1; if (idx[mid] < key) l = mid + 1; else h = mid; }
    return l < hi && idx[l] == key;
}

/* users may be NULL (identity). mask_indptr may be NULL (no mask). Returns 0. */
int oracle_score_topk(const float* U, const float* V, const int64_t* users, int Bt, int m_items, int d,
                      const int32_t* mask_indptr, const int32_t* mask_indices, int mask_col_offset,
                      int k, int64_t* idx_out, float* val_out) {
    float* bv = (float*)malloc(sizeof(float) * (size_t)k);
    int* bi = (int*)malloc(sizeof(int) * (size_t)k);
    for (int b = 0; b < Bt; ++b) {
        const int64_t u = users ? users[b] : b;
        const float* ur = U + (size_t)u * d;
        int cnt = 0;
        const int lo = mask_indptr ? mask_indptr[u] : 0, hi = mask_indptr ? mask_indptr[u + 1] : 0;
        for (int i = 0; i < m_items; ++i) {
            const float* vr = V + (size_t)i * d;
            float s = 0.f;
            for (int q = 0; q < d; ++q) s = fmaf(ur[q], vr[q], s);
            if (hi > lo && row_has(mask_indices, lo, hi, mask_col_offset + i)) s = -1024.f;
            if (cnt < k || s > bv[k - 1]) {
                int pos = cnt < k ? cnt : k - 1;
                while (pos > 0 && bv[pos - 1] < s) { bv[pos] = bv[pos - 1]; bi[pos] = bi[pos - 1]; --pos; }
                bv[pos] = s; bi[pos] = i;
                if (cnt < k) ++cnt;
            }
        }
        for (int q = 0; q < k; ++q) { idx_out[(size_t)b * k + q] = bi[q]; val_out[(size_t)b * k + q] = bv[q]; }
    }
    free(bv); free(bi);
    return 0;
}

/* dense scores with the same FMA chain */
int oracle_score_dense(const float* U, const float* V, const int64_t* users, int Bt, int m_items, int d, float* out) {
    for (int b = 0; b < Bt; ++b) {
        const float* ur = U + (size_t)(users ? users[b] : b) * d;
        for (int i = 0; i < m_items; ++i) {
            const float* vr = V + (size_t)i * d;
            float s = 0.f;
            for (int q = 0; q < d; ++q) s = fmaf(ur[q], vr[q], s);
            out[(size_t)b * m_items + i] = s;
        }
    }
    return 0;
}

/* fp32 CSR SpMM in CSR order (row by row, non-zeros in stored order), plain mul+add like a scalar
 * CPU loop — the summation order torch's CPU sparse addmm uses (reference code/model.py:217). */
int oracle_spmm_f32(const int32_t* indptr, const int32_t* indices, const float* vals, int n_rows, int d,
                    const float* X, float* Y) {
    for (int r = 0; r < n_rows; ++r) {
        float* y = Y + (size_t)r * d;
        for (int q = 0; q < d; ++q) y[q] = 0.f;
        for (int j = indptr[r]; j < indptr[r + 1]; ++j) {
            const float v = vals[j]; const float* x = X + (size_t)indices[j] * d;
            for (int q = 0; q < d; ++q) y[q] += v * x[q];
        }
    }
    return 0;
}
