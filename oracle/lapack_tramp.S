/* LAPACK/BLAS entry points for the oracle build of the unmodified reference.
 * TEST INFRASTRUCTURE ONLY.  The container has no system LAPACK; the scipy wheel
 * bundles OpenBLAS exporting every routine as scipy_<name> (LP64 ints).  Each
 * Fortran symbol declared in reference src/base/LinearAlgebra.h:68-107 is a
 * tail-jump to it. */
	.text
	.globl dgbsv_
	.type dgbsv_,@function
dgbsv_:
	jmp scipy_dgbsv_@PLT
	.globl dgbtrf_
	.type dgbtrf_,@function
dgbtrf_:
	jmp scipy_dgbtrf_@PLT
	.globl dgbtrs_
	.type dgbtrs_,@function
dgbtrs_:
	jmp scipy_dgbtrs_@PLT
	.globl dgesv_
	.type dgesv_,@function
dgesv_:
	jmp scipy_dgesv_@PLT
	.globl dgetrf_
	.type dgetrf_,@function
dgetrf_:
	jmp scipy_dgetrf_@PLT
	.globl dgetrs_
	.type dgetrs_,@function
dgetrs_:
	jmp scipy_dgetrs_@PLT
	.globl dgetri_
	.type dgetri_,@function
dgetri_:
	jmp scipy_dgetri_@PLT
	.globl dgemm_
	.type dgemm_,@function
dgemm_:
	jmp scipy_dgemm_@PLT
	.globl dgtsv_
	.type dgtsv_,@function
dgtsv_:
	jmp scipy_dgtsv_@PLT
	.globl dtpsv_
	.type dtpsv_,@function
dtpsv_:
	jmp scipy_dtpsv_@PLT
	.globl dtrtri_
	.type dtrtri_,@function
dtrtri_:
	jmp scipy_dtrtri_@PLT
	.globl dgeqrf_
	.type dgeqrf_,@function
dgeqrf_:
	jmp scipy_dgeqrf_@PLT
	.globl dorgqr_
	.type dorgqr_,@function
dorgqr_:
	jmp scipy_dorgqr_@PLT
	.globl dgesvd_
	.type dgesvd_,@function
dgesvd_:
	jmp scipy_dgesvd_@PLT
	.globl daxpy_
	.type daxpy_,@function
daxpy_:
	jmp scipy_daxpy_@PLT
	.globl dcopy_
	.type dcopy_,@function
dcopy_:
	jmp scipy_dcopy_@PLT
	.globl ddot_
	.type ddot_,@function
ddot_:
	jmp scipy_ddot_@PLT
	.globl dnrm2_
	.type dnrm2_,@function
dnrm2_:
	jmp scipy_dnrm2_@PLT
	.globl dscal_
	.type dscal_,@function
dscal_:
	jmp scipy_dscal_@PLT
	.section .note.GNU-stack,"",@progbits
