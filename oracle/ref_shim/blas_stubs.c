/* oracle/ref_shim/blas_stubs.c -- TEST INFRASTRUCTURE.  Aborting definitions of
 * the BLAS/LAPACK symbols that the reference's TX-only helpers
 * (cpuLS.hpp:391-529) reference; the receive path never reaches them. */
#include <stdio.h>
#include <stdlib.h>
#include "cblas.h"
static void die(const char *n) { fprintf(stderr, "blas stub %s called (TX path is out of scope)\n", n); abort(); }
void cblas_cgemm(const enum CBLAS_ORDER o, const enum CBLAS_TRANSPOSE a, const enum CBLAS_TRANSPOSE b,
                 const int M, const int N, const int K, const void *al, const void *A, const int lda,
                 const void *B, const int ldb, const void *be, void *C, const int ldc)
{ (void)o;(void)a;(void)b;(void)M;(void)N;(void)K;(void)al;(void)A;(void)lda;(void)B;(void)ldb;(void)be;(void)C;(void)ldc; die("cblas_cgemm"); }
void cblas_cgemv(const enum CBLAS_ORDER o, const enum CBLAS_TRANSPOSE a, const int M, const int N,
                 const void *al, const void *A, const int lda, const void *X, const int incX,
                 const void *be, void *Y, const int incY)
{ (void)o;(void)a;(void)M;(void)N;(void)al;(void)A;(void)lda;(void)X;(void)incX;(void)be;(void)Y;(void)incY; die("cblas_cgemv"); }
void cblas_csscal(const int N, const float alpha, void *X, const int incX)
{ (void)N;(void)alpha;(void)X;(void)incX; die("cblas_csscal"); }
int cblas_icamax(const int N, const void *X, const int incX)
{ (void)N;(void)X;(void)incX; die("cblas_icamax"); return 0; }
void cgetrf_(int *m, int *n, void *A, int *lda, int *ipiv, int *info)
{ (void)m;(void)n;(void)A;(void)lda;(void)ipiv;(void)info; die("cgetrf_"); }
void cgetri_(int *n, void *A, int *lda, int *ipiv, void *work, int *lwork, int *info)
{ (void)n;(void)A;(void)lda;(void)ipiv;(void)work;(void)lwork;(void)info; die("cgetri_"); }
void csytrf_(char *u, int *n, void *A, int *lda, int *ipiv, void *work, int *lwork, int *info)
{ (void)u;(void)n;(void)A;(void)lda;(void)ipiv;(void)work;(void)lwork;(void)info; die("csytrf_"); }
void csytri_(char *u, int *n, void *A, int *lda, int *ipiv, void *work, int *info)
{ (void)u;(void)n;(void)A;(void)lda;(void)ipiv;(void)work;(void)info; die("csytri_"); }
float clange_(char *norm, int *m, int *n, void *A, int *lda, float *work)
{ (void)norm;(void)m;(void)n;(void)A;(void)lda;(void)work; die("clange_"); return 0.f; }
