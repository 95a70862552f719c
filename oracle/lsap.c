/*
 * ORACLE — test infrastructure, not product code.
 *
 * CPU restatement of the rectangular linear-sum-assignment solver that the
 * reference calls through scipy.optimize.linear_sum_assignment
 * (reference call sites: AGQA/src/lxrt/matcher.py:8, :79, :103).
 *
 * The arithmetic lives in an un-vendored third-party dependency: scipy
 * (pinned scipy==1.7.3 in AGQA/requirements.txt:82; this image has 1.18.1).
 * The algorithm restated here is scipy's published `rectangular_lsap`
 * (shortest augmenting path, Crouse 2016), including its tie-break: the
 * `remaining` column list starts REVERSED, is swap-removed, and among equal
 * reduced costs a still-unassigned column wins.
 *
 * Parity pinning: the reference ships no tests; this file is pinned against the
 * installed scipy (tests/test_oracle_lsap.py: random / tied / constant /
 * duplicate-column / empty matrices, bit-exact) and against the committed
 * known-answer vectors in tests/golden/lsap_kat.npz produced by
 * tools/gen_golden.py from scipy itself.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference
 * arm may call into this file.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define SHG_ORACLE_MAXDIM 64

/* One augmentation: shortest path from row `cur` to an unassigned column. */
static int augment(int nr, int nc, const double *C, int cur, double *u, double *v,
                   int *path, const int *row4col, double *spc, unsigned char *SR,
                   unsigned char *SC, int *remaining, double *p_min) {
    double min_val = 0.0;
    int n_rem = nc;
    for (int it = 0; it < nc; ++it) remaining[it] = nc - it - 1; /* reversed */
    memset(SR, 0, (size_t)nr);
    memset(SC, 0, (size_t)nc);
    for (int j = 0; j < nc; ++j) spc[j] = INFINITY;

    int sink = -1;
    int i = cur;
    while (sink == -1) {
        int index = -1;
        double lowest = INFINITY;
        SR[i] = 1;
        for (int it = 0; it < n_rem; ++it) {
            int j = remaining[it];
            double r = min_val + C[(size_t)i * nc + j] - u[i] - v[j];
            if (r < spc[j]) {
                path[j] = i;
                spc[j] = r;
            }
            if (spc[j] < lowest || (spc[j] == lowest && row4col[j] == -1)) {
                lowest = spc[j];
                index = it;
            }
        }
        min_val = lowest;
        if (min_val == INFINITY) return -1; /* infeasible */
        int j = remaining[index];
        if (row4col[j] == -1) sink = j;
        else i = row4col[j];
        SC[j] = 1;
        remaining[index] = remaining[--n_rem];
    }
    *p_min = min_val;
    return sink;
}

/*
 * cost: nr x nc row-major float64. Writes k = min(nr,nc) pairs (row ascending).
 * Returns k, or -1 infeasible, -2 invalid entry (NaN / -inf: scipy raises
 * ValueError), -3 dimension above SHG_ORACLE_MAXDIM.
 */
int shg_oracle_lsap_f64(int nr, int nc, const double *cost, int64_t *row_ind, int64_t *col_ind) {
    if (nr == 0 || nc == 0) return 0;
    if (nr > SHG_ORACLE_MAXDIM || nc > SHG_ORACLE_MAXDIM) return -3;
    for (int t = 0; t < nr * nc; ++t)
        if (isnan(cost[t]) || cost[t] == -INFINITY) return -2;

    double Ct[SHG_ORACLE_MAXDIM * SHG_ORACLE_MAXDIM];
    const double *C = cost;
    int transpose = nc < nr;
    if (transpose) {
        for (int i = 0; i < nr; ++i)
            for (int j = 0; j < nc; ++j) Ct[(size_t)j * nr + i] = cost[(size_t)i * nc + j];
        int t = nr; nr = nc; nc = t;
        C = Ct;
    }
    double u[SHG_ORACLE_MAXDIM], v[SHG_ORACLE_MAXDIM], spc[SHG_ORACLE_MAXDIM];
    int path[SHG_ORACLE_MAXDIM], col4row[SHG_ORACLE_MAXDIM], row4col[SHG_ORACLE_MAXDIM];
    int remaining[SHG_ORACLE_MAXDIM];
    unsigned char SR[SHG_ORACLE_MAXDIM], SC[SHG_ORACLE_MAXDIM];
    for (int i = 0; i < nr; ++i) { u[i] = 0.0; col4row[i] = -1; }
    for (int j = 0; j < nc; ++j) { v[j] = 0.0; row4col[j] = -1; path[j] = -1; }

    for (int cur = 0; cur < nr; ++cur) {
        double min_val;
        int sink = augment(nr, nc, C, cur, u, v, path, row4col, spc, SR, SC, remaining, &min_val);
        if (sink < 0) return -1;
        u[cur] += min_val;
        for (int i = 0; i < nr; ++i)
            if (SR[i] && i != cur) u[i] += min_val - spc[col4row[i]];
        for (int j = 0; j < nc; ++j)
            if (SC[j]) v[j] -= min_val - spc[j];
        int j = sink;
        for (;;) {
            int i = path[j];
            row4col[j] = i;
            int t = col4row[i]; col4row[i] = j; j = t;
            if (i == cur) break;
        }
    }
    if (transpose) {
        /* argsort(col4row): col4row is injective so a counting pass suffices */
        int k = 0;
        for (int q = 0; q < nc; ++q)
            if (row4col[q] != -1) { row_ind[k] = q; col_ind[k] = row4col[q]; ++k; }
        return k;
    }
    for (int i = 0; i < nr; ++i) { row_ind[i] = i; col_ind[i] = col4row[i]; }
    return nr;
}

/* float32 costs widened exactly to float64, as numpy does for scipy (matcher.py:79). */
int shg_oracle_lsap_f32(int nr, int nc, const float *cost, int ld, int64_t *row_ind, int64_t *col_ind) {
    if (nr > SHG_ORACLE_MAXDIM || nc > SHG_ORACLE_MAXDIM) return -3;
    double C[SHG_ORACLE_MAXDIM * SHG_ORACLE_MAXDIM];
    for (int i = 0; i < nr; ++i)
        for (int j = 0; j < nc; ++j) C[(size_t)i * nc + j] = (double)cost[(size_t)i * ld + j];
    return shg_oracle_lsap_f64(nr, nc, C, row_ind, col_ind);
}

/*
 * Batched per-frame matching on precomputed fp32 cost blocks, the shape the
 * reference feeds scipy (matcher.py:76-79): frame f owns rows [f*q_f,(f+1)*q_f)
 * of `cost` (leading dim ld) and the column range [col_off[f], col_off[f+1]).
 * Output: for frame f, pairs at out_rows/out_cols[f*q_f ...], count in out_n[f].
 */
int shg_oracle_match_frames(int n_frames, int q_f, const float *cost, int ld, const int32_t *col_off,
                            int64_t *out_rows, int64_t *out_cols, int32_t *out_n) {
    for (int f = 0; f < n_frames; ++f) {
        int n = col_off[f + 1] - col_off[f];
        int k = shg_oracle_lsap_f32(q_f, n, cost + (size_t)f * q_f * ld + col_off[f], ld,
                                    out_rows + (size_t)f * q_f, out_cols + (size_t)f * q_f);
        if (k < 0) return k;
        out_n[f] = k;
    }
    return 0;
}
