/*
 * smm_oracle.c -- CPU ORACLE for the smmregrid weight-application hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product package (smmregrid_b200/)
 * may import, link or execute this file.  It is used by tests/, by
 * __graft_entry__.smoke() as the checker, and by bench.py's cpu_baseline /
 * --impl reference legs as the timed CPU port of the reference.
 *
 * PARITY STATUS: "parity unpinned" at the sparse/dask boundary.  The reference
 * (jhardenberg/smmregrid v0.1.6, pure Python) delegates its arithmetic to
 * third-party packages that are not vendored under /root/reference and are not
 * installable in this image (pydata `sparse` -- unpinned, pyproject.toml:29;
 * `dask` -- unpinned, pyproject.toml:26; xarray).  The reference ships no
 * golden vectors, stored weights or stored outputs (every numeric test calls
 * the `cdo` binary at run time).  This file therefore restates the published
 * algorithm of those packages as called from the reference's own call sites:
 *
 *   smmregrid/weights.py:25-44   compute_weights_matrix  -> orc_coo_build
 *        src_address-1, dst_address-1, remap_matrix[:,0], shape (n_src,n_dst);
 *        sparse.COO(coords, data): coordinates sorted lexicographically
 *        (src major, dst minor; stable), duplicate coordinates summed.
 *   smmregrid/weights.py:47-52   mask_tensordot          -> orc_mask_sum
 *        t = src_mask . W  ;  where(t < 0.5, 0, 1)
 *   smmregrid/regrid.py:544-547  NaN/inf -> 1e20 fill (np.ma.fix_invalid +
 *        filled; default float fill_value 1e20 *in the input dtype*)
 *                                                        -> orc_fill_*
 *   smmregrid/regrid.py:550      tensordot(X[B,n_src], W) -> orc_dot_*
 *        pydata/sparse ndarray.COO kernel: for every batch row, for every
 *        stored link k in COO order: out[b,dst[k]] += x[b,src[k]] * w[k],
 *        product and accumulation in result_type(x,w) (= float64 for CDO's
 *        float64 weights), separate multiply and add (numba does not contract
 *        to FMA).
 *   smmregrid/regrid.py:553-570  where(imask, Y, nan); where(frac < min, nan,
 *        Y); where(Y > 1e19, nan, Y)                     -> orc_epilogue
 *
 * Build:  make -C oracle      (gcc -O2 -fno-fast-math -ffp-contract=off)
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_F32 0
#define ORC_F64 1

/* ---------------------------------------------------------------- COO build */

typedef struct {
    int64_t key; /* src * n_dst + dst : sparse.COO linear location */
    int64_t pos; /* original position: keeps the sort stable        */
} orc_key_t;

static int orc_key_cmp(const void *a, const void *b)
{
    const orc_key_t *x = (const orc_key_t *)a, *y = (const orc_key_t *)b;
    if (x->key != y->key) return x->key < y->key ? -1 : 1;
    return x->pos < y->pos ? -1 : (x->pos > y->pos);
}

/*
 * weights.py:31-39.  Inputs are CDO arrays: 1-based (index_base=1) int32
 * addresses, remap_matrix[nnz, num_wgts] of which only column 0 is used.
 * Outputs (caller-allocated, length nnz): COO sorted src-major with duplicate
 * coordinates summed in original order.  Returns the number of stored links,
 * or -1 on an out-of-range address.
 */
int64_t orc_coo_build(int64_t n_src, int64_t n_dst, int64_t nnz,
                      const int32_t *src_address, const int32_t *dst_address,
                      const double *remap_matrix, int num_wgts, int index_base,
                      int32_t *coo_src, int32_t *coo_dst, double *coo_w)
{
    if (nnz == 0) return 0;
    orc_key_t *keys = (orc_key_t *)malloc((size_t)nnz * sizeof(orc_key_t));
    if (!keys) return -2;
    for (int64_t k = 0; k < nnz; ++k) {
        int64_t s = (int64_t)src_address[k] - index_base;
        int64_t d = (int64_t)dst_address[k] - index_base;
        if (s < 0 || s >= n_src || d < 0 || d >= n_dst) { free(keys); return -1; }
        keys[k].key = s * n_dst + d;
        keys[k].pos = k;
    }
    qsort(keys, (size_t)nnz, sizeof(orc_key_t), orc_key_cmp);
    int64_t m = 0;
    for (int64_t k = 0; k < nnz; ++k) {
        double w = remap_matrix[keys[k].pos * (int64_t)num_wgts];
        if (m > 0 && keys[k].key == keys[k - 1].key) {
            coo_w[m - 1] += w; /* sum_duplicates */
        } else {
            coo_src[m] = (int32_t)(keys[k].key / n_dst);
            coo_dst[m] = (int32_t)(keys[k].key % n_dst);
            coo_w[m] = w;
            ++m;
        }
    }
    free(keys);
    return m;
}

/* ----------------------------------------------------------- mask_tensordot */

/* weights.py:47-52: t[dst] = sum_k imask[src[k]] * w[k] in COO order (float64),
 * dst_imask = t < 0.5 ? 0 : 1.  t_out may be NULL. */
void orc_mask_sum(int64_t n_dst, int64_t m, const int32_t *coo_src,
                  const int32_t *coo_dst, const double *coo_w,
                  const int32_t *src_imask, int32_t *dst_imask, double *t_out)
{
    double *t = t_out ? t_out : (double *)malloc((size_t)n_dst * sizeof(double));
    for (int64_t d = 0; d < n_dst; ++d) t[d] = 0.0;
    for (int64_t k = 0; k < m; ++k)
        t[coo_dst[k]] += (double)src_imask[coo_src[k]] * coo_w[k];
    for (int64_t d = 0; d < n_dst; ++d) dst_imask[d] = t[d] < 0.5 ? 0 : 1;
    if (!t_out) free(t);
}

/* ------------------------------------------------------------------- apply */

typedef struct {
    /* matrix */
    int64_t n_src, n_dst, m;
    const int32_t *coo_src, *coo_dst;
    const double *coo_w;
    /* data */
    const void *x; int x_dtype; int64_t ldx;
    double *y; int64_t ldy;           /* result_type(x, float64) = float64 */
    /* epilogue */
    const int32_t *dst_imask;         /* NULL <=> masked == False          */
    const double *dst_frac;           /* used when remap_area_min > 0      */
    double remap_area_min;
    /* batch range for this worker */
    int64_t b0, b1;
} orc_job_t;

static void orc_rows(const orc_job_t *j)
{
    const int64_t n_src = j->n_src, n_dst = j->n_dst, m = j->m;
    /* temporaries materialised like the reference's elementwise passes */
    float *xf = NULL; double *xd = NULL;
    if (j->x_dtype == ORC_F32) xf = (float *)malloc((size_t)n_src * sizeof(float));
    else                       xd = (double *)malloc((size_t)n_src * sizeof(double));
    for (int64_t b = j->b0; b < j->b1; ++b) {
        double *out = j->y + b * j->ldy;
        /* regrid.py:545-547 fix_invalid + filled: non-finite -> 1e20 in the input dtype */
        if (xf) {
            const float *x = (const float *)j->x + b * j->ldx;
            for (int64_t i = 0; i < n_src; ++i) xf[i] = isfinite(x[i]) ? x[i] : 1e20f;
        } else {
            const double *x = (const double *)j->x + b * j->ldx;
            for (int64_t i = 0; i < n_src; ++i) xd[i] = isfinite(x[i]) ? x[i] : 1e20;
        }
        /* regrid.py:550 tensordot -> sparse ndarray.COO loop, links in COO order */
        for (int64_t d = 0; d < n_dst; ++d) out[d] = 0.0;
        if (xf) {
            for (int64_t k = 0; k < m; ++k) {
                double p = (double)xf[j->coo_src[k]] * j->coo_w[k];
                out[j->coo_dst[k]] = out[j->coo_dst[k]] + p;
            }
        } else {
            for (int64_t k = 0; k < m; ++k) {
                double p = xd[j->coo_src[k]] * j->coo_w[k];
                out[j->coo_dst[k]] = out[j->coo_dst[k]] + p;
            }
        }
        /* regrid.py:553-559 destination mask */
        if (j->dst_imask)
            for (int64_t d = 0; d < n_dst; ++d)
                if (!j->dst_imask[d]) out[d] = NAN;
        /* regrid.py:562-565 dst_grid_frac < remap_area_min */
        if (j->remap_area_min > 0.0 && j->dst_frac)
            for (int64_t d = 0; d < n_dst; ++d)
                if (j->dst_frac[d] < j->remap_area_min) out[d] = NAN;
        /* regrid.py:570 bring the NaN back in */
        for (int64_t d = 0; d < n_dst; ++d)
            if (out[d] > 1e19) out[d] = NAN;
    }
    free(xf); free(xd);
}

static void *orc_thread(void *arg) { orc_rows((const orc_job_t *)arg); return NULL; }

/*
 * regrid.py:536-570 on B batch rows.  x is [B, ldx] (n_src fastest), y is
 * [B, ldy] float64.  nthreads > 1 splits the batch axis into contiguous
 * chunks, one pthread each, mirroring dask's threaded scheduler running one
 * sparse.tensordot task per chunk of the kept dims.
 */
int orc_apply(int64_t n_src, int64_t n_dst, int64_t m,
              const int32_t *coo_src, const int32_t *coo_dst, const double *coo_w,
              const void *x, int x_dtype, int64_t B, int64_t ldx,
              double *y, int64_t ldy,
              const int32_t *dst_imask, const double *dst_frac,
              double remap_area_min, int nthreads)
{
    if (x_dtype != ORC_F32 && x_dtype != ORC_F64) return -1;
    if (nthreads < 1) nthreads = 1;
    if (nthreads > B) nthreads = (int)(B > 0 ? B : 1);
    orc_job_t *jobs = (orc_job_t *)calloc((size_t)nthreads, sizeof(orc_job_t));
    pthread_t *tids = (pthread_t *)calloc((size_t)nthreads, sizeof(pthread_t));
    for (int t = 0; t < nthreads; ++t) {
        orc_job_t *j = &jobs[t];
        j->n_src = n_src; j->n_dst = n_dst; j->m = m;
        j->coo_src = coo_src; j->coo_dst = coo_dst; j->coo_w = coo_w;
        j->x = x; j->x_dtype = x_dtype; j->ldx = ldx; j->y = y; j->ldy = ldy;
        j->dst_imask = dst_imask; j->dst_frac = dst_frac;
        j->remap_area_min = remap_area_min;
        j->b0 = B * t / nthreads; j->b1 = B * (t + 1) / nthreads;
    }
    if (nthreads == 1) {
        orc_rows(&jobs[0]);
    } else {
        for (int t = 0; t < nthreads; ++t) pthread_create(&tids[t], NULL, orc_thread, &jobs[t]);
        for (int t = 0; t < nthreads; ++t) pthread_join(tids[t], NULL);
    }
    free(jobs); free(tids);
    return 0;
}
