/*
 * dgl_cpu_oracle.c -- TEST INFRASTRUCTURE ONLY.  CPU restatement of DGL v0.6.1's CPU kernels
 * for the sparse message-passing path (gspmm / gsddmm / COO->CSR), used as the parity checker
 * in tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs.
 * Nothing under dgl-0.5-benchmark_b200/ may import, link or call this file.
 *
 * PARITY UNPINNED: the reference repository (/root/reference) holds no golden vectors and no
 * tests for this path, and the arithmetic lives in the un-vendored pip
 * dependency `dgl-cu111` (docker/build.dockerfile:14; README.md:6 states v0.6.1), which is
 * absent from this image and cannot be installed offline.  This file therefore restates the
 * published algorithm of dmlc/dgl@0.6.1 from knowledge of that tree:
 *     src/array/cpu/spmat_op_impl_coo.cc :: COOToCSR        -> oracle_coo_to_csr
 *     src/array/cpu/spmm.h :: SpMMSumCsr / SpMMCmpCsr        -> oracle_spmm_csr
 *     src/array/cpu/sddmm.h :: SDDMMCoo (+ Dot functor)      -> oracle_sddmm_coo
 *     include/dgl/bcast.h :: CalcBcastOff                    -> offsets computed by the Python side
 * and is anchored on the reference's own call sites (kernel/dgl-new.py:20,39;
 * main_dgl_citation_sage.py:75-77) and its written-out twins (kernel/pyg.py:47-49,
 * kernel/utils.py:8-16, main_pyg_arxiv_gat.py:98-111).  It is pinned against the hand-derived
 * known-answer vector of SURVEY.md Appendix A.6 and against independent fp64 restatements
 * (scipy CSR matmul, torch index_add_/scatter_reduce) in tests/test_oracle.py.
 *
 * Loop order, accumulation order (sequential fp32 in CSR row order), the strict `<` compare of
 * the max reducer (first entry in row order wins ties) and the OpenMP row/edge parallelisation
 * follow the upstream kernels; compile with -ffp-contract=off so mul+add is not fused (upstream
 * wheels target baseline x86-64, no FMA).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

enum { OP_ADD = 0, OP_SUB = 1, OP_MUL = 2, OP_DIV = 3, OP_COPY_LHS = 4, OP_COPY_RHS = 5, OP_DOT = 6 };
enum { RED_SUM = 0, RED_MAX = 1, RED_MIN = 2 };
enum { TGT_U = 0, TGT_E = 1, TGT_V = 2 };

int oracle_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

void oracle_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ---- COOToCSR (dmlc/dgl@0.6.1 src/array/cpu/spmat_op_impl_coo.cc): scipy-style counting sort.
 * Stable: entries of one row keep increasing edge id.  Row-sorted input keeps data = identity
 * (upstream leaves `data` null in that case; we materialise arange). */
void oracle_coo_to_csr(int64_t n_rows, int64_t nnz, const int32_t* row, const int32_t* col,
                       int32_t* indptr, int32_t* indices, int32_t* data) {
  memset(indptr, 0, sizeof(int32_t) * (size_t)(n_rows + 1));
  for (int64_t i = 0; i < nnz; ++i) indptr[row[i]]++;
  int32_t cumsum = 0;
  for (int64_t r = 0; r < n_rows; ++r) {
    int32_t t = indptr[r];
    indptr[r] = cumsum;
    cumsum += t;
  }
  indptr[n_rows] = (int32_t)nnz;
  for (int64_t i = 0; i < nnz; ++i) {
    int32_t r = row[i];
    int32_t p = indptr[r];
    indices[p] = col[i];
    data[p] = (int32_t)i;
    indptr[r]++;
  }
  int32_t last = 0;
  for (int64_t r = 0; r <= n_rows; ++r) {
    int32_t t = indptr[r];
    indptr[r] = last;
    last = t;
  }
}

static inline float binop(int op, const float* l, const float* r) {
  switch (op) {
    case OP_ADD: return *l + *r;
    case OP_SUB: return *l - *r;
    case OP_MUL: return *l * *r;
    case OP_DIV: return *l / *r;
    case OP_COPY_LHS: return *l;
    default: return *r; /* OP_COPY_RHS */
  }
}

/* ---- SpMMSumCsr / SpMMCmpCsr (dmlc/dgl@0.6.1 src/array/cpu/spmm.h).
 * lhs_off/rhs_off: broadcast offset tables of length out_len (NULL => identity), as produced
 * by CalcBcastOff.  eids NULL => edge id = CSR position.  arg_u/arg_e only for max/min. */
void oracle_spmm_csr(int op, int reduce, int64_t n_rows, const int32_t* indptr,
                     const int32_t* indices, const int32_t* eids, const float* X, const float* W,
                     int64_t lhs_len, int64_t rhs_len, int64_t out_len, const int64_t* lhs_off,
                     const int64_t* rhs_off, float* out, int32_t* arg_u, int32_t* arg_e) {
  const int use_lhs = (op != OP_COPY_RHS);
  const int use_rhs = (op != OP_COPY_LHS);
#pragma omp parallel for schedule(static)
  for (int64_t rid = 0; rid < n_rows; ++rid) {
    const int32_t row_start = indptr[rid], row_end = indptr[rid + 1];
    float* out_row = out + rid * out_len;
    if (reduce == RED_SUM) {
      for (int64_t k = 0; k < out_len; ++k) out_row[k] = 0.f;
      for (int32_t j = row_start; j < row_end; ++j) {
        const int64_t cid = indices[j];
        const int64_t eid = eids ? eids[j] : j;
        for (int64_t k = 0; k < out_len; ++k) {
          const int64_t la = lhs_off ? lhs_off[k] : k;
          const int64_t ra = rhs_off ? rhs_off[k] : k;
          const float* l = use_lhs ? X + cid * lhs_len + la : NULL;
          const float* r = use_rhs ? W + eid * rhs_len + ra : NULL;
          out_row[k] += binop(op, l, r);
        }
      }
    } else {
      const float zero = (reduce == RED_MAX) ? -INFINITY : INFINITY;
      int32_t* au = arg_u ? arg_u + rid * out_len : NULL;
      int32_t* ae = arg_e ? arg_e + rid * out_len : NULL;
      for (int64_t k = 0; k < out_len; ++k) {
        out_row[k] = zero;
        if (au) au[k] = 0;
        if (ae) ae[k] = 0;
      }
      for (int32_t j = row_start; j < row_end; ++j) {
        const int64_t cid = indices[j];
        const int64_t eid = eids ? eids[j] : j;
        for (int64_t k = 0; k < out_len; ++k) {
          const int64_t la = lhs_off ? lhs_off[k] : k;
          const int64_t ra = rhs_off ? rhs_off[k] : k;
          const float* l = use_lhs ? X + cid * lhs_len + la : NULL;
          const float* r = use_rhs ? W + eid * rhs_len + ra : NULL;
          const float val = binop(op, l, r);
          const int better = (reduce == RED_MAX) ? (out_row[k] < val) : (out_row[k] > val);
          if (better) {
            out_row[k] = val;
            if (au) au[k] = (int32_t)cid;
            if (ae) ae[k] = (int32_t)eid;
          }
        }
      }
    }
  }
}

/* ---- SDDMMCoo (dmlc/dgl@0.6.1 src/array/cpu/sddmm.h), parallel over edges.
 * out[e*out_len + k] = op(lhs[sel_l(e)*lhs_len + lhs_off[k]*reduce_size ...], rhs[...]);
 * Dot: sequential fp32 sum over reduce_size.  Output is in edge-id order. */
void oracle_sddmm_coo(int op, int lhs_target, int rhs_target, int64_t nnz, const int32_t* src,
                      const int32_t* dst, const float* L, const float* R, int64_t lhs_len,
                      int64_t rhs_len, int64_t out_len, int64_t reduce_size,
                      const int64_t* lhs_off, const int64_t* rhs_off, float* out) {
  const int use_lhs = (op != OP_COPY_RHS);
  const int use_rhs = (op != OP_COPY_LHS);
#pragma omp parallel for schedule(static)
  for (int64_t e = 0; e < nnz; ++e) {
    const int64_t lid = lhs_target == TGT_U ? src[e] : (lhs_target == TGT_E ? e : dst[e]);
    const int64_t rid = rhs_target == TGT_U ? src[e] : (rhs_target == TGT_E ? e : dst[e]);
    float* out_row = out + e * out_len;
    for (int64_t k = 0; k < out_len; ++k) {
      const int64_t la = lhs_off ? lhs_off[k] : k;
      const int64_t ra = rhs_off ? rhs_off[k] : k;
      const float* l = use_lhs ? L + lid * lhs_len + la * reduce_size : NULL;
      const float* r = use_rhs ? R + rid * rhs_len + ra * reduce_size : NULL;
      if (op == OP_DOT) {
        float acc = 0.f;
        for (int64_t i = 0; i < reduce_size; ++i) acc += l[i] * r[i];
        out_row[k] = acc;
      } else {
        out_row[k] = binop(op, l, r);
      }
    }
  }
}
