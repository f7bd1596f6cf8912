/*
 * sgd_oracle.c -- plain-C CPU restatement of the reference's SGD trainers.  TEST INFRASTRUCTURE ONLY:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may link or call it.
 *
 * Follows (paths relative to the reference checkout):
 *   MatrixFactorization/Cython/MatrixFactorization_Cython_Epoch.pyx
 *       epochIteration_Cython_BPR_SGD       :583-678      sampleBPR_Cython   :943-987
 *       epochIteration_Cython_FUNK_SVD_SGD  :289-390      sampleMSE_Cython   :881-938
 *       epochIteration_Cython_ASY_SVD_SGD   :396-578      (batch size 1; USER_factors is the n_items x f matrix Y)
 *       _apply_minibatch_updates_...        :773-832      adaptive_gradient  :838-876
 *       first-touch lists                   :709-769      adam powers        :220-221, :361-364, :649-652
 *   SLIM_BPR/Cython/SLIM_BPR_Cython_Epoch.pyx
 *       epochIteration_Cython :211-335, sampleBPR_Cython :436-480, adaptive_gradient :395-433,
 *       Triangular_Matrix add_value/get_value :1272-1330 (symmetric storage)
 *       Sparse_Matrix_Tree_CSR (train_with_sparse_weights): add_value/get_value :617-733, rebalance_tree :782-802,
 *       topK_selection_from_list :954-1031 -- restated on a dense array plus a cell-exists map (slim_prune below)
 *   libc rand()/srand(): glibc TYPE_3 additive-feedback generator (not under the reference tree; its published
 *   algorithm is restated in glibc_rand_* below and pinned against libc's rand() in tests/test_oracle_sgd.py).
 *
 * All arithmetic is double precision, in the reference's order.  State lives in caller-owned arrays so that the
 * Python wrapper (oracle/sgd_oracle.py) can seed factors with numpy's legacy RNG exactly as the reference does
 * (pyx:145-147,177-178).  Pinned against the compiled reference (oracle/_ref) to <=1e-12 in
 * tests/test_oracle_sgd.py, and by golden vectors in tests/golden/sgd_golden.npz.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* ---------------------------------------------------------------- glibc rand() replay */
typedef struct { int32_t r[34]; int f, b; } glibc_rand_t; /* r[3..33] is the live 31-word table */

void glibc_rand_seed(glibc_rand_t* g, unsigned int seed) {
  int32_t r[34];
  int i;
  int32_t word = seed == 0 ? 1 : (int32_t)seed;
  r[0] = word;
  for (i = 1; i < 31; ++i) {
    /* 16807 * r[i-1] % 2147483647 without overflow (Schrage) */
    long hi = word / 127773, lo = word % 127773;
    long w = 16807 * lo - 2836 * hi;
    if (w < 0) w += 2147483647;
    word = (int32_t)w;
    r[i] = word;
  }
  memcpy(g->r, r, sizeof(int32_t) * 31);
  g->f = 3; /* front pointer = &r[sep], sep = 3 */
  g->b = 0; /* rear pointer */
  for (i = 0; i < 310; ++i) {
    g->r[g->f] = (int32_t)((uint32_t)g->r[g->f] + (uint32_t)g->r[g->b]);
    g->f = (g->f + 1) % 31;
    g->b = (g->b + 1) % 31;
  }
}

int glibc_rand_next(glibc_rand_t* g) {
  uint32_t v = (uint32_t)g->r[g->f] + (uint32_t)g->r[g->b];
  g->r[g->f] = (int32_t)v;
  g->f = (g->f + 1) % 31;
  g->b = (g->b + 1) % 31;
  return (int)(v >> 1);
}

#define ORACLE_RAND_MAX 2147483647

/* ---------------------------------------------------------------- samplers */
/* pyx:943-987 (MF) == SLIM pyx:436-480 */
static void sample_bpr(glibc_rand_t* g, const int32_t* indptr, const int32_t* indices, int n_users, int n_items,
                       long* u_out, long* i_out, long* j_out) {
  long u = 0, start = 0, end = 0, n = 0, idx, j;
  while (n == 0 || n == n_items) {
    u = glibc_rand_next(g) % n_users;
    start = indptr[u];
    end = indptr[u + 1];
    n = end - start;
  }
  idx = glibc_rand_next(g) % n;
  *i_out = indices[start + idx];
  for (;;) {
    j = glibc_rand_next(g) % n_items;
    idx = 0;
    while (idx < n && indices[start + idx] < j) ++idx; /* the reference's linear scan, pyx:976-978 */
    if (idx == n || indices[start + idx] > j) break;
  }
  *u_out = u;
  *j_out = j;
}

/* pyx:881-938; quota is the probability of sampling a POSITIVE (pyx:901) */
static void sample_mse(glibc_rand_t* g, const int32_t* indptr, const int32_t* indices, const double* data, int n_users,
                       int n_items, double quota, long* u_out, long* i_out, double* r_out) {
  long u = 0, start = 0, end = 0, n = 0, idx, item;
  int positive;
  while (n == 0 || n == n_items) {
    u = glibc_rand_next(g) % n_users;
    start = indptr[u];
    end = indptr[u + 1];
    n = end - start;
  }
  if (quota != 0.0) positive = glibc_rand_next(g) <= quota * ORACLE_RAND_MAX; else positive = 1;
  if (positive) {
    idx = glibc_rand_next(g) % n;
    *i_out = indices[start + idx];
    *r_out = data[start + idx];
  } else {
    for (;;) {
      item = glibc_rand_next(g) % n_items;
      idx = 0;
      while (idx < n && indices[start + idx] < item) ++idx;
      if (idx == n || indices[start + idx] > item) break;
    }
    *i_out = item;
    *r_out = 0.0;
  }
  *u_out = u;
}

/* ---------------------------------------------------------------- adaptive gradient, pyx:838-876 */
typedef struct {
  int mode; /* 0 sgd, 1 adagrad, 2 rmsprop, 3 adam */
  double gamma, beta1, beta2, b1_pow, b2_pow;
} adapt_t;

static double adapt(const adapt_t* a, double g, double* cache, double* m1, double* m2) {
  if (a->mode == 1) {
    *cache += g * g;
    return g / (sqrt(*cache) + 1e-8);
  } else if (a->mode == 2) {
    *cache = *cache * a->gamma + (1 - a->gamma) * g * g;
    return g / (sqrt(*cache) + 1e-8);
  } else if (a->mode == 3) {
    double mm1, mm2;
    *m1 = *m1 * a->beta1 + (1 - a->beta1) * g;
    *m2 = *m2 * a->beta2 + (1 - a->beta2) * g * g;
    mm1 = *m1 / (1 - a->b1_pow);
    mm2 = *m2 / (1 - a->b2_pow);
    return mm1 / (sqrt(mm2) + 1e-8);
  }
  return g;
}

/* ---------------------------------------------------------------- MF trainer state (caller-owned arrays) */
typedef struct {
  int n_users, n_items, f, batch_size, algorithm; /* algorithm: 0 MF_BPR, 1 FUNK_SVD, 2 ASY_SVD (U holds n_items rows) */
  int use_bias;
  double lr, user_reg, item_reg, bias_reg, positive_reg, negative_reg, quota;
  adapt_t ad;
  const int32_t* indptr; const int32_t* indices; const double* data; long nnz;
  double *U, *V, *bu, *bi, *mu;                 /* parameters */
  double *accU, *accV, *accbu, *accbi, *accmu;  /* mini-batch accumulators */
  double *cU, *cV, *cbu, *cbi, *cmu;            /* adagrad/rmsprop cache */
  double *m1U, *m2U, *m1V, *m2V, *m1bu, *m2bu, *m1bi, *m2bi, *m1mu, *m2mu; /* adam */
  long *items_list, *users_list; char *items_flag, *users_flag; long n_items_touched, n_users_touched;
  glibc_rand_t rng;
  /* optional externally supplied sample stream (u,i,j) / (u,i,rating): when non-NULL the samplers are bypassed */
  const int32_t* ext_u; const int32_t* ext_i; const int32_t* ext_j; const double* ext_r; long ext_pos;
  /* optional recording of the stream actually used */
  int32_t *rec_u, *rec_i, *rec_j; long rec_pos;
} mf_t;

static void touch_item(mf_t* s, long it) {
  if (!s->items_flag[it]) { s->items_flag[it] = 1; s->items_list[s->n_items_touched++] = it; }
}
static void touch_user(mf_t* s, long u) {
  if (!s->users_flag[u]) { s->users_flag[u] = 1; s->users_list[s->n_users_touched++] = u; }
}

/* pyx:773-832: global bias, then items in first-touch order (bias then factors), then users */
static void apply_minibatch(mf_t* s) {
  long n, k;
  int f = s->f, q;
  double g;
  if (s->use_bias) {
    g = s->accmu[0] / s->batch_size;
    g = adapt(&s->ad, g, s->cmu, s->m1mu, s->m2mu);
    s->mu[0] += s->lr * g;
    s->accmu[0] = 0.0;
  }
  for (n = 0; n < s->n_items_touched; ++n) {
    k = s->items_list[n];
    if (s->use_bias) {
      g = s->accbi[k] / s->batch_size;
      g = adapt(&s->ad, g, s->cbi ? s->cbi + k : 0, s->m1bi ? s->m1bi + k : 0, s->m2bi ? s->m2bi + k : 0);
      s->bi[k] += s->lr * g;
      s->accbi[k] = 0.0;
    }
    for (q = 0; q < f; ++q) {
      g = s->accV[k * f + q] / s->batch_size;
      g = adapt(&s->ad, g, s->cV ? s->cV + k * f + q : 0, s->m1V ? s->m1V + k * f + q : 0, s->m2V ? s->m2V + k * f + q : 0);
      s->V[k * f + q] += s->lr * g;
      s->accV[k * f + q] = 0.0;
    }
  }
  for (n = 0; n < s->n_users_touched; ++n) {
    k = s->users_list[n];
    if (s->use_bias) {
      g = s->accbu[k] / s->batch_size;
      g = adapt(&s->ad, g, s->cbu ? s->cbu + k : 0, s->m1bu ? s->m1bu + k : 0, s->m2bu ? s->m2bu + k : 0);
      s->bu[k] += s->lr * g;
      s->accbu[k] = 0.0;
    }
    for (q = 0; q < f; ++q) {
      g = s->accU[k * f + q] / s->batch_size;
      g = adapt(&s->ad, g, s->cU ? s->cU + k * f + q : 0, s->m1U ? s->m1U + k * f + q : 0, s->m2U ? s->m2U + k * f + q : 0);
      s->U[k * f + q] += s->lr * g;
      s->accU[k * f + q] = 0.0;
    }
  }
}

static void clear_minibatch(mf_t* s) { /* pyx:723-735 */
  long n;
  for (n = 0; n < s->n_items_touched; ++n) s->items_flag[s->items_list[n]] = 0;
  for (n = 0; n < s->n_users_touched; ++n) s->users_flag[s->users_list[n]] = 0;
  s->n_items_touched = 0;
  s->n_users_touched = 0;
}

/* pyx:396-578: one sample per step; the user is represented by the sum of the Y rows (s->U, n_items x f) of the items
 * in the profile divided by sqrt(profile length); Y rows of the whole profile and the X row (s->V) of the sampled item are
 * updated at once, every parameter with its own adaptive state. */
static long asy_svd_epoch(mf_t* s) {
  const int f = s->f;
  const long n_total = s->nnz / 1 + 1; /* pyx:402 with batch_size == 1 */
  long b, k, count = 0;
  int q;
  double* acc = (double*)malloc(sizeof(double) * (size_t)f);
  for (b = 0; b < n_total; ++b, ++count) {
    long u, i, start, end;
    double r = 0.0, pred, err, den, g;
    if (s->ext_u) { u = s->ext_u[s->ext_pos]; i = s->ext_i[s->ext_pos]; r = s->ext_r[s->ext_pos]; s->ext_pos++; }
    else sample_mse(&s->rng, s->indptr, s->indices, s->data, s->n_users, s->n_items, s->quota, &u, &i, &r);
    if (s->rec_u) { s->rec_u[s->rec_pos] = (int32_t)u; s->rec_i[s->rec_pos] = (int32_t)i; s->rec_j[s->rec_pos] = -1; s->rec_pos++; }
    start = s->indptr[u]; end = s->indptr[u + 1];
    for (q = 0; q < f; ++q) acc[q] = 0.0;
    for (k = start; k < end; ++k) { const long it = s->indices[k]; for (q = 0; q < f; ++q) acc[q] += s->U[it * f + q]; }
    den = sqrt((double)(end - start));
    for (q = 0; q < f; ++q) acc[q] /= den;
    pred = s->use_bias ? s->mu[0] + s->bu[u] + s->bi[i] : 0.0;
    for (q = 0; q < f; ++q) pred += acc[q] * s->V[i * f + q];
    err = r - pred;
    if (s->use_bias) {
      g = err - s->bias_reg * s->mu[0];
      g = adapt(&s->ad, g, s->cmu, s->m1mu, s->m2mu);
      s->mu[0] += s->lr * g;
      g = adapt(&s->ad, err - s->bias_reg * s->bi[i], s->cbi ? s->cbi + i : 0, s->m1bi ? s->m1bi + i : 0, s->m2bi ? s->m2bi + i : 0);
      {
        const double gu = adapt(&s->ad, err - s->bias_reg * s->bu[u], s->cbu ? s->cbu + u : 0, s->m1bu ? s->m1bu + u : 0, s->m2bu ? s->m2bu + u : 0);
        s->bi[i] += s->lr * g;
        s->bu[u] += s->lr * gu;
      }
    }
    for (k = start; k < end; ++k) { /* pyx:505-521: every Y row of the profile, H_i read from the not yet updated X row */
      const long it = s->indices[k];
      for (q = 0; q < f; ++q) {
        const double Hi = s->V[i * f + q], Wu = s->U[it * f + q];
        g = err * Hi - s->user_reg * Wu;
        g = adapt(&s->ad, g, s->cU ? s->cU + it * f + q : 0, s->m1U ? s->m1U + it * f + q : 0, s->m2U ? s->m2U + it * f + q : 0);
        s->U[it * f + q] += s->lr * g;
      }
    }
    for (q = 0; q < f; ++q) { /* pyx:524-539: the accumulated profile from BEFORE the Y update */
      const double Hi = s->V[i * f + q];
      g = err * acc[q] - s->item_reg * Hi;
      g = adapt(&s->ad, g, s->cV ? s->cV + i * f + q : 0, s->m1V ? s->m1V + i * f + q : 0, s->m2V ? s->m2V + i * f + q : 0);
      s->V[i * f + q] += s->lr * g;
    }
    if (s->ad.mode == 3) { s->ad.b1_pow *= s->ad.beta1; s->ad.b2_pow *= s->ad.beta2; } /* per sample, pyx:544-547 */
  }
  free(acc);
  return count;
}

/* one epochIteration_Cython; returns the number of samples processed */
long mf_epoch(mf_t* s) {
  const int f = s->f;
  long n_batches, b, smp, count = 0;
  int q;
  if (s->algorithm == 2) return asy_svd_epoch(s);
  if (s->algorithm == 0) n_batches = (long)(s->n_users / s->batch_size) + 1;  /* pyx:586 */
  else n_batches = (long)(s->nnz / s->batch_size) + 1;                        /* pyx:292 */
  for (b = 0; b < n_batches; ++b) {
    clear_minibatch(s);
    for (smp = 0; smp < s->batch_size; ++smp, ++count) {
      long u, i, j = -1;
      double r = 0.0;
      if (s->ext_u) {
        u = s->ext_u[s->ext_pos]; i = s->ext_i[s->ext_pos];
        if (s->algorithm == 0) j = s->ext_j[s->ext_pos]; else r = s->ext_r[s->ext_pos];
        s->ext_pos++;
      } else if (s->algorithm == 0) {
        sample_bpr(&s->rng, s->indptr, s->indices, s->n_users, s->n_items, &u, &i, &j);
      } else {
        sample_mse(&s->rng, s->indptr, s->indices, s->data, s->n_users, s->n_items, s->quota, &u, &i, &r);
      }
      if (s->rec_u) { s->rec_u[s->rec_pos] = (int32_t)u; s->rec_i[s->rec_pos] = (int32_t)i; s->rec_j[s->rec_pos] = (int32_t)j; s->rec_pos++; }
      if (s->algorithm == 0) {
        /* pyx:608-642; first-touch order: pos item, neg item, user (pyx:754-769) */
        double x = 0.0, sig;
        touch_item(s, i); touch_item(s, j); touch_user(s, u);
        for (q = 0; q < f; ++q) x += s->U[u * f + q] * (s->V[i * f + q] - s->V[j * f + q]);
        sig = 1 / (1 + exp(x));
        for (q = 0; q < f; ++q) {
          const double Hi = s->V[i * f + q], Hj = s->V[j * f + q], Wu = s->U[u * f + q];
          const double gi = sig * Wu - s->positive_reg * Hi;
          const double gj = sig * (-Wu) - s->negative_reg * Hj;
          const double gu = sig * (Hi - Hj) - s->user_reg * Wu;
          s->accU[u * f + q] += gu;
          s->accV[i * f + q] += gi;
          s->accV[j * f + q] += gj;
        }
      } else {
        /* pyx:305-354; first-touch: item then user (pyx:739-749) */
        double pred = 0.0, err;
        touch_item(s, i); touch_user(s, u);
        if (s->use_bias) pred = s->mu[0] + s->bu[u] + s->bi[i];
        for (q = 0; q < f; ++q) pred += s->U[u * f + q] * s->V[i * f + q];
        err = r - pred;
        if (s->use_bias) {
          s->accmu[0] += err - s->bias_reg * s->mu[0];
          s->accbi[i] += err - s->bias_reg * s->bi[i];
          s->accbu[u] += err - s->bias_reg * s->bu[u];
        }
        for (q = 0; q < f; ++q) {
          const double Hi = s->V[i * f + q], Wu = s->U[u * f + q];
          s->accV[i * f + q] += err * Wu - s->positive_reg * Hi; /* positive_reg, not item_reg: pyx:349 */
          s->accU[u * f + q] += err * Hi - s->user_reg * Wu;
        }
      }
    }
    apply_minibatch(s);
    if (s->ad.mode == 3) { s->ad.b1_pow *= s->ad.beta1; s->ad.b2_pow *= s->ad.beta2; } /* once per batch */
  }
  return count;
}

/* ---------------------------------------------------------------- SLIM-BPR, dense or symmetric S */
typedef struct {
  int n_users, n_items, symmetric;
  double lr, li_reg, lj_reg;
  adapt_t ad;
  const int32_t* indptr; const int32_t* indices;
  double* S;                   /* n_items x n_items, row-major; symmetric mode mirrors (i,s) and (s,i) */
  double *c, *m1, *m2;         /* per-ITEM adaptive state (pyx:395-433) */
  glibc_rand_t rng;
  const int32_t* ext_u; const int32_t* ext_i; const int32_t* ext_j; long ext_pos;
  int32_t *rec_u, *rec_i, *rec_j; long rec_pos;
  unsigned char* exists;       /* tree mode (train_with_sparse_weights): 1 where the reference's row tree holds a cell */
  int tree_topk;               /* the TopK of rebalance_tree (0 = False: nothing is removed) */
} slim_t;

static double slim_get(const slim_t* s, long a, long b) {
  if (s->symmetric && b > a) { long t = a; a = b; b = t; } /* lower-triangular storage, pyx:1309-1330 */
  return s->S[a * (long)s->n_items + b];
}
static void slim_add(slim_t* s, long a, long b, double v) {
  if (s->symmetric && b > a) { long t = a; a = b; b = t; }
  s->S[a * (long)s->n_items + b] += v;
  if (s->exists) s->exists[a * (long)s->n_items + b] = 1; /* add_value creates the cell, pyx:627-680 */
}

/* rebalance_tree(TopK) pyx:782-802 / the selection inside get_scipy_csr(TopK) pyx:762-763, through
 * topK_selection_from_list pyx:954-1031: a row that holds at least TopK cells keeps the TopK largest by value.  The list is
 * in ascending column order and glibc's qsort is a stable merge sort at these sizes, so among equal values the higher
 * columns sit last and are the ones kept.  Cells that are dropped cease to exist (they read as 0 again). */
typedef struct { double v; long col; } slim_cell_t;
static int slim_cell_cmp(const void* a, const void* b) {
  const slim_cell_t *x = (const slim_cell_t*)a, *y = (const slim_cell_t*)b;
  if (x->v < y->v) return -1;
  if (x->v > y->v) return 1;
  return x->col < y->col ? -1 : (x->col > y->col ? 1 : 0);
}
void slim_prune(slim_t* s, long topk) {
  const long n = s->n_items;
  long r, c, m, q;
  slim_cell_t* cells;
  if (!s->exists || topk <= 0) return;
  cells = (slim_cell_t*)malloc(sizeof(slim_cell_t) * (size_t)n);
  for (r = 0; r < n; ++r) {
    m = 0;
    for (c = 0; c < n; ++c)
      if (s->exists[r * n + c]) { cells[m].v = s->S[r * n + c]; cells[m].col = c; ++m; }
    if (m < topk) continue;
    qsort(cells, (size_t)m, sizeof(slim_cell_t), slim_cell_cmp);
    for (q = 0; q < m - topk; ++q) { s->S[r * n + cells[q].col] = 0.0; s->exists[r * n + cells[q].col] = 0; }
  }
  free(cells);
}
/* get_S in tree mode touches every diagonal cell (add_value(index, index, -get_value(index, index)), pyx:349-350): the
 * cell exists afterwards, with value 0 */
void slim_touch_diagonal(slim_t* s) {
  long r;
  if (!s->exists) return;
  for (r = 0; r < s->n_items; ++r) { s->S[r * (long)s->n_items + r] = 0.0; s->exists[r * (long)s->n_items + r] = 1; }
}

long slim_epoch(slim_t* s) { /* pyx:211-335 */
  long n, count = 0;
  for (n = 0; n < s->n_users; ++n, ++count) {
    long u, i, j, k, start, end;
    double x = 0.0, g, gi, gj;
    if (s->ext_u) { u = s->ext_u[s->ext_pos]; i = s->ext_i[s->ext_pos]; j = s->ext_j[s->ext_pos]; s->ext_pos++; }
    else sample_bpr(&s->rng, s->indptr, s->indices, s->n_users, s->n_items, &u, &i, &j);
    if (s->rec_u) { s->rec_u[s->rec_pos] = (int32_t)u; s->rec_i[s->rec_pos] = (int32_t)i; s->rec_j[s->rec_pos] = (int32_t)j; s->rec_pos++; }
    start = s->indptr[u]; end = s->indptr[u + 1];
    for (k = start; k < end; ++k) { const long sn = s->indices[k]; x += slim_get(s, i, sn) - slim_get(s, j, sn); }
    g = 1 / (1 + exp(x));
    gi = adapt(&s->ad, g, s->c ? s->c + i : 0, s->m1 ? s->m1 + i : 0, s->m2 ? s->m2 + i : 0); /* i first, then j */
    gj = adapt(&s->ad, g, s->c ? s->c + j : 0, s->m1 ? s->m1 + j : 0, s->m2 ? s->m2 + j : 0);
    for (k = start; k < end; ++k) {
      const long sn = s->indices[k];
      if (sn != i) slim_add(s, i, sn, s->lr * (gi - s->li_reg * slim_get(s, i, sn)));
      if (sn != j) slim_add(s, j, sn, -s->lr * (gj - s->lj_reg * slim_get(s, j, sn)));
    }
    if (s->ad.mode == 3) { s->ad.b1_pow *= s->ad.beta1; s->ad.b2_pow *= s->ad.beta2; } /* per sample, pyx:309-312 */
    /* pyx:318-319: `n_current_sample % (self.n_users/5) == 0` is a float modulo under language_level=3 */
    if (s->exists && n != 0 && fmod((double)n, (double)s->n_users / 5.0) == 0.0) slim_prune(s, s->tree_topk);
  }
  return count;
}

unsigned long sizeof_mf(void) { return sizeof(mf_t); }
unsigned long sizeof_slim(void) { return sizeof(slim_t); }
