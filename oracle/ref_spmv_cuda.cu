/*============================================================================
 * oracle/ref_spmv_cuda.cu -- TEST INFRASTRUCTURE, not product code.
 *
 * The bar SURVEY.md 2.5 sets for the SpMV kernel: the reference's OWN device kernel
 * for MSR matrices, `_mat_vect_p_l_msr` (src/alge/cs_matrix_spmv_cuda.cu:415-439; a
 * thread per row, 256 threads per block, launched as at :1338-1345), compiled for
 * sm_100a from the file where it lies (the #include below is resolved against
 * /root/reference at build time: oracle/Makefile.ref; no reference source is copied
 * here), and the cuSPARSE path the reference offers beside it
 * (cs_matrix_spmv_cuda_msr_cusparse, :1540-1615: `_mat_vect_p_l_msr_diag` for
 * y = D.x, then cusparseSpMV(ALG_DEFAULT) with beta = 1 on the extra-diagonal CSR part,
 * descriptors as in :917-1030).  bench.py times both beside k_spmv_stream /
 * k_spmv_pipeline on the same matrix in the same run (`spmv.reference_kernel_gbs`,
 * `spmv.cusparse_gbs`).  The rest of that translation unit (host wrappers around
 * cs_matrix_t) is compiled but never called: the library is loaded with RTLD_LAZY.
 *============================================================================*/

#include <cusparse_v2.h>      /* before the reference file: it redefines __ldg on the host pass */

#define HAVE_CUDA 1
#define HAVE_ACCEL 1
#include "alge/cs_matrix_spmv_cuda.cu"

namespace {
struct CusparseMsr {
  cusparseHandle_t     handle = nullptr;
  cusparseSpMatDescr_t matA = nullptr;
  cusparseDnVecDescr_t vecX = nullptr, vecY = nullptr;
  void                *buffer = nullptr;
  int                  n_rows = 0;
  const double        *d_val = nullptr;
};
}

extern "C" {

/* y = A.x with the reference's kernel; raw device pointers. */
int
csref_cuda_spmv_msr(int n_rows, const int *row_index, const int *col_id, const double *d_val,
                    const double *x_val, const double *x, double *y, void *stream)
{
  const unsigned int blocksize = 256;
  const unsigned int gridsize = (unsigned int)ceil((double)n_rows / blocksize);
  _mat_vect_p_l_msr<<<gridsize, blocksize, 0, (cudaStream_t)stream>>>
    (n_rows, row_index, col_id, d_val, x_val, x, y);
  return (int)cudaGetLastError();
}

/* cuSPARSE descriptors of the extra-diagonal part (device arrays), as the reference sets
   them up; returns an opaque handle or NULL. */
void *
csref_cusparse_msr_create(int n_rows, int n_cols, int nnz, int *row_index, int *col_id,
                          double *d_val, double *x_val, double *x, double *y, void *stream)
{
  CusparseMsr *c = new CusparseMsr;
  c->n_rows = n_rows; c->d_val = d_val;
  double alpha = 1., beta = 1.;
  size_t bsize = 0;
  if (cusparseCreate(&c->handle) != CUSPARSE_STATUS_SUCCESS
      || cusparseSetStream(c->handle, (cudaStream_t)stream) != CUSPARSE_STATUS_SUCCESS
      || cusparseCreateCsr(&c->matA, n_rows, n_cols, nnz, row_index, col_id, x_val,
                           CUSPARSE_INDEX_32I, CUSPARSE_INDEX_32I, CUSPARSE_INDEX_BASE_ZERO,
                           CUDA_R_64F) != CUSPARSE_STATUS_SUCCESS
      || cusparseCreateDnVec(&c->vecX, n_cols, x, CUDA_R_64F) != CUSPARSE_STATUS_SUCCESS
      || cusparseCreateDnVec(&c->vecY, n_rows, y, CUDA_R_64F) != CUSPARSE_STATUS_SUCCESS
      || cusparseSpMV_bufferSize(c->handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, c->matA,
                                 c->vecX, &beta, c->vecY, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT,
                                 &bsize) != CUSPARSE_STATUS_SUCCESS
      || cudaMalloc(&c->buffer, bsize > 0 ? bsize : 16) != cudaSuccess) {
    delete c;
    return nullptr;
  }
  return c;
}

/* y = D.x (reference kernel), then y += E.x (cusparseSpMV, beta = 1). */
int
csref_cusparse_msr_spmv(void *h, const double *x, double *y, void *stream)
{
  CusparseMsr *c = static_cast<CusparseMsr *>(h);
  const unsigned int blocksize = 256;
  const unsigned int gridsize = (unsigned int)ceil((double)c->n_rows / blocksize);
  _mat_vect_p_l_msr_diag<<<gridsize, blocksize, 0, (cudaStream_t)stream>>>
    (c->n_rows, c->d_val, x, y);
  double alpha = 1., beta = 1.;
  cusparseDnVecSetValues(c->vecX, (void *)x);
  cusparseDnVecSetValues(c->vecY, (void *)y);
  return (int)cusparseSpMV(c->handle, CUSPARSE_OPERATION_NON_TRANSPOSE, &alpha, c->matA, c->vecX,
                           &beta, c->vecY, CUDA_R_64F, CUSPARSE_SPMV_ALG_DEFAULT, c->buffer);
}

void
csref_cusparse_msr_destroy(void *h)
{
  CusparseMsr *c = static_cast<CusparseMsr *>(h);
  if (c == nullptr) return;
  cusparseDestroySpMat(c->matA); cusparseDestroyDnVec(c->vecX); cusparseDestroyDnVec(c->vecY);
  cusparseDestroy(c->handle);
  cudaFree(c->buffer);
  delete c;
}

}
