/* CPU oracle, C part (TEST INFRASTRUCTURE ONLY -- never linked into or called by the product).
 *
 * Exact mutual-nearest-neighbour data for sizes whose similarity matrix does not fit in memory
 * (BASELINE config 5: 64k x 64k): the maths of the reference matchers
 *   sim = A @ B.T; nn12 = argmax_j sim[i, j]; nn21 = argmax_i sim[i, j]      (first maximum wins)
 *   evaluations/hpatches/evaluation.py:27-38, evaluations/aachen/matchers.py:5-13,
 *   losses/preprocess_utils.py:795-803
 * on the EXACT similarities: products of float32 numbers are exact in float64, sums are float64.
 * Besides the argmax it returns the largest and second largest value of every row and column, so that a
 * test can prove that a disagreement is a float64 tie.  The N x M matrix is never held: row panels of
 * PI rows are multiplied against column panels of PJ columns of B^T (converted to double once and packed in
 * groups of 8 columns so that the k loop streams contiguous memory).
 * Built by oracle/Makefile (gcc -O3 -pthread); pinned against the numpy oracle and the reference-generated
 * fixtures in tests/test_oracle_golden.py.
 */
#include <math.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define PI 8
#define PJ 512

static inline void upd(double v, int64_t idx, double* best, double* second, int64_t* arg) {
  /* first maximum wins: replace only when strictly larger, or equal with a smaller index */
  if (v > *best || (v == *best && idx < *arg)) {
    *second = *best;
    *best = v;
    *arg = idx;
  } else if (v > *second) {
    *second = v;
  }
}

typedef struct {
  const float* a;
  const double* bt;
  int64_t N, M, D, Mp;
  int64_t* nn12;
  double *row_best, *row_second;
  double* cb;      /* this thread's column state: [M][2] best, second */
  int64_t* ca;     /* [M] argmax row */
  int tid, nthreads;
} Job;

typedef double v4d __attribute__((vector_size(32), aligned(8)));

static void* worker(void* arg) {
  Job* q = (Job*)arg;
  const int64_t N = q->N, M = q->M, D = q->D, Mp = q->Mp;
  static __thread double acc[PI][PJ];
  double* ad = (double*)malloc(sizeof(double) * (size_t)PI * (size_t)D);   /* the row panel of A in double, [k][PI] */
  if (!ad) return (void*)1;
  /* row panels are dealt round-robin: every row is owned by exactly one thread */
  for (int64_t i0 = (int64_t)q->tid * PI; i0 < N; i0 += (int64_t)q->nthreads * PI) {
    const int ni = (int)((N - i0) < PI ? (N - i0) : PI);
    for (int64_t k = 0; k < D; ++k)
      for (int r = 0; r < PI; ++r) ad[k * PI + r] = r < ni ? (double)q->a[(i0 + r) * D + k] : 0.0;
    for (int64_t j0 = 0; j0 < M; j0 += PJ) {
      /* register-blocked: 4 rows x 8 columns of accumulators live in registers over the whole k loop; every
       * element is the plain sequential sum over k = 0..D-1 of exact products (no reassociation) */
      for (int rb4 = 0; rb4 < PI; rb4 += 4) {
        for (int j = 0; j < PJ; j += 8) {
          v4d c00 = {0, 0, 0, 0}, c01 = c00, c10 = c00, c11 = c00, c20 = c00, c21 = c00, c30 = c00, c31 = c00;
          const double* bk = q->bt + (j0 + j) * D;          /* 8-column group: [k][8] contiguous */
          const double* ak = ad + rb4;
          for (int64_t k = 0; k < D; ++k, bk += 8, ak += PI) {
            const v4d b0 = *(const v4d*)bk, b1 = *(const v4d*)(bk + 4);
            const v4d x0 = {ak[0], ak[0], ak[0], ak[0]}, x1 = {ak[1], ak[1], ak[1], ak[1]};
            const v4d x2 = {ak[2], ak[2], ak[2], ak[2]}, x3 = {ak[3], ak[3], ak[3], ak[3]};
            c00 += x0 * b0; c01 += x0 * b1;
            c10 += x1 * b0; c11 += x1 * b1;
            c20 += x2 * b0; c21 += x2 * b1;
            c30 += x3 * b0; c31 += x3 * b1;
          }
          *(v4d*)&acc[rb4 + 0][j] = c00; *(v4d*)&acc[rb4 + 0][j + 4] = c01;
          *(v4d*)&acc[rb4 + 1][j] = c10; *(v4d*)&acc[rb4 + 1][j + 4] = c11;
          *(v4d*)&acc[rb4 + 2][j] = c20; *(v4d*)&acc[rb4 + 2][j + 4] = c21;
          *(v4d*)&acc[rb4 + 3][j] = c30; *(v4d*)&acc[rb4 + 3][j + 4] = c31;
        }
      }
      const int nj = (int)((M - j0) < PJ ? (M - j0) : PJ);
      for (int r = 0; r < ni; ++r) {
        const int64_t i = i0 + r;
        double rb = q->row_best[i], rs = q->row_second[i];
        int64_t ra = q->nn12[i];
        for (int j = 0; j < nj; ++j) {
          const double v = acc[r][j];
          if (v > rs) {                                       /* rare once the running values have settled */
            if (v > rb) { rs = rb; rb = v; ra = j0 + j; }      /* columns ascend: strict > keeps the first maximum */
            else rs = v;
          }
          /* a thread's rows ascend, so within a thread an equal value never has a smaller index */
          if (v > q->cb[2 * (j0 + j) + 1]) upd(v, i, &q->cb[2 * (j0 + j)], &q->cb[2 * (j0 + j) + 1], &q->ca[j0 + j]);
        }
        q->row_best[i] = rb; q->row_second[i] = rs; q->nn12[i] = ra;
      }
    }
  }
  free(ad);
  return NULL;
}

int posfeat_oracle_mnn_f64(const float* a, int64_t N, const float* b, int64_t M, int64_t D, int nthreads,
                           int64_t* nn12, double* row_best, double* row_second, int64_t* nn21, double* col_best,
                           double* col_second) {
  if (N <= 0 || M <= 0 || D <= 0) return 1;
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 256) nthreads = 256;
  const int64_t Mp = (M + PJ - 1) / PJ * PJ;
  double* bt = (double*)calloc((size_t)D * (size_t)Mp, sizeof(double));   /* B^T in double, zero padded to Mp columns */
  double* tb = (double*)malloc(sizeof(double) * (size_t)nthreads * (size_t)M * 2);
  int64_t* ta = (int64_t*)malloc(sizeof(int64_t) * (size_t)nthreads * (size_t)M);
  Job* jobs = (Job*)malloc(sizeof(Job) * (size_t)nthreads);
  pthread_t* th = (pthread_t*)malloc(sizeof(pthread_t) * (size_t)nthreads);
  if (!bt || !tb || !ta || !jobs || !th) { free(bt); free(tb); free(ta); free(jobs); free(th); return 2; }
  for (int64_t j = 0; j < M; ++j)   /* packed by groups of 8 columns: bt[(j / 8) * D * 8 + k * 8 + j % 8] */
    for (int64_t k = 0; k < D; ++k) bt[(j >> 3) * D * 8 + k * 8 + (j & 7)] = (double)b[j * D + k];
  for (int64_t i = 0; i < N; ++i) { row_best[i] = -INFINITY; row_second[i] = -INFINITY; nn12[i] = 0; }
  for (int64_t j = 0; j < M; ++j) { col_best[j] = -INFINITY; col_second[j] = -INFINITY; nn21[j] = INT64_MAX; }
  for (int64_t x = 0; x < (int64_t)nthreads * M; ++x) { tb[2 * x] = -INFINITY; tb[2 * x + 1] = -INFINITY; ta[x] = INT64_MAX; }
  for (int t = 0; t < nthreads; ++t) {
    Job j = {a, bt, N, M, D, Mp, nn12, row_best, row_second, tb + (size_t)t * (size_t)M * 2, ta + (size_t)t * (size_t)M, t, nthreads};
    jobs[t] = j;
    if (pthread_create(&th[t], NULL, worker, &jobs[t]) != 0) {
      for (int u = 0; u < t; ++u) pthread_join(th[u], NULL);
      free(bt); free(tb); free(ta); free(jobs); free(th);
      return 3;
    }
  }
  for (int t = 0; t < nthreads; ++t) pthread_join(th[t], NULL);
  /* merge the per-thread column states (upd keeps the first-index rule across threads) */
  for (int t = 0; t < nthreads; ++t) {
    const double* cb = tb + (size_t)t * (size_t)M * 2;
    const int64_t* ca = ta + (size_t)t * (size_t)M;
    for (int64_t j = 0; j < M; ++j) {
      if (ca[j] == INT64_MAX) continue;
      upd(cb[2 * j], ca[j], &col_best[j], &col_second[j], &nn21[j]);
      if (cb[2 * j + 1] > col_second[j]) col_second[j] = cb[2 * j + 1];
    }
  }
  free(bt); free(tb); free(ta); free(jobs); free(th);
  return 0;
}
