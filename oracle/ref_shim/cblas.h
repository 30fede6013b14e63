/* oracle/ref_shim/cblas.h -- TEST INFRASTRUCTURE.  Declarations only, so that
 * the reference's cpuLS.hpp:40 include resolves.  CBLAS/LAPACK are used solely
 * by the downlink/TX helpers (cpuLS.hpp:391-529), which are outside the
 * receive hot path and are never called; oracle/ref_shim/blas_stubs.c
 * provides aborting definitions so the reference header links. */
#ifndef REF_SHIM_CBLAS_H
#define REF_SHIM_CBLAS_H
#ifdef __cplusplus
extern "C" {
#endif
enum CBLAS_ORDER { CblasRowMajor = 101, CblasColMajor = 102 };
enum CBLAS_TRANSPOSE { CblasNoTrans = 111, CblasTrans = 112, CblasConjTrans = 113 };
void cblas_cgemm(const enum CBLAS_ORDER Order, const enum CBLAS_TRANSPOSE TransA,
                 const enum CBLAS_TRANSPOSE TransB, const int M, const int N, const int K,
                 const void *alpha, const void *A, const int lda, const void *B, const int ldb,
                 const void *beta, void *C, const int ldc);
void cblas_cgemv(const enum CBLAS_ORDER order, const enum CBLAS_TRANSPOSE TransA, const int M,
                 const int N, const void *alpha, const void *A, const int lda, const void *X,
                 const int incX, const void *beta, void *Y, const int incY);
void cblas_csscal(const int N, const float alpha, void *X, const int incX);
int cblas_icamax(const int N, const void *X, const int incX);
#ifdef __cplusplus
}
#endif
#endif
