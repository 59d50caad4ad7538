! lapack_b200_iso_c.f90 -- ISO_C_BINDING view of the C ABI for Fortran hosts.
!
! The Fortran-77 symbols (dgetrf_, dpotrf_, dgeqrf_, ...) exported by liblapack_b200.so already have the
! gfortran calling convention, so an existing Fortran caller needs NO source change: link liblapack_b200.so
! before liblapack (the reference's own mechanism for SRC/VARIANTS, SRC/VARIANTS/README:66-78).
! This module is for new Fortran code that wants the device-pointer API (lb200_*): matrices stay resident in
! GPU memory (type(c_ptr) device addresses) across calls and the stream is explicit.
! Shipped as source; this image has no Fortran compiler, so it is compile-gated like the reference's own
! "CMake couldn't find a Fortran compiler" branch (CMakeLists.txt:313-317).
module lapack_b200
  use, intrinsic :: iso_c_binding
  implicit none
  interface
     ! SRC/dgetrf.f:105  DGETRF(M,N,A,LDA,IPIV,INFO) on device pointers
     integer(c_int) function lb200_dgetrf(stream, m, n, dA, lda, dipiv, dinfo) bind(C, name="lb200_dgetrf")
       import :: c_int, c_long_long, c_ptr
       type(c_ptr), value :: stream, dA, dipiv, dinfo
       integer(c_int), value :: m, n
       integer(c_long_long), value :: lda
     end function
     ! SRC/dgetrs.f:118
     integer(c_int) function lb200_dgetrs(stream, trans, n, nrhs, dA, lda, dipiv, dB, ldb) bind(C, name="lb200_dgetrs")
       import :: c_int, c_long_long, c_ptr, c_char
       type(c_ptr), value :: stream, dA, dipiv, dB
       character(kind=c_char), value :: trans
       integer(c_int), value :: n, nrhs
       integer(c_long_long), value :: lda, ldb
     end function
     ! SRC/dpotrf.f:104
     integer(c_int) function lb200_dpotrf(stream, uplo, n, dA, lda, dinfo) bind(C, name="lb200_dpotrf")
       import :: c_int, c_long_long, c_ptr, c_char
       type(c_ptr), value :: stream, dA, dinfo
       character(kind=c_char), value :: uplo
       integer(c_int), value :: n
       integer(c_long_long), value :: lda
     end function
     ! SRC/dpotrs.f:107
     integer(c_int) function lb200_dpotrs(stream, uplo, n, nrhs, dA, lda, dB, ldb) bind(C, name="lb200_dpotrs")
       import :: c_int, c_long_long, c_ptr, c_char
       type(c_ptr), value :: stream, dA, dB
       character(kind=c_char), value :: uplo
       integer(c_int), value :: n, nrhs
       integer(c_long_long), value :: lda, ldb
     end function
     ! SRC/dgeqrf.f:145 (device scratch replaces WORK/LWORK)
     integer(c_int) function lb200_dgeqrf(stream, m, n, dA, lda, dtau) bind(C, name="lb200_dgeqrf")
       import :: c_int, c_long_long, c_ptr
       type(c_ptr), value :: stream, dA, dtau
       integer(c_int), value :: m, n
       integer(c_long_long), value :: lda
     end function
     ! BLAS/SRC/dgemm.f:187
     integer(c_int) function lb200_dgemm(stream, transa, transb, m, n, k, alpha, dA, lda, dB, ldb, beta, dC, ldc) &
          bind(C, name="lb200_dgemm")
       import :: c_int, c_long_long, c_ptr, c_char, c_double
       type(c_ptr), value :: stream, dA, dB, dC
       character(kind=c_char), value :: transa, transb
       integer(c_int), value :: m, n, k
       real(c_double), value :: alpha, beta
       integer(c_long_long), value :: lda, ldb, ldc
     end function
     ! SRC/dormqr.f:165  C := Q C, Q**T C, C Q or C Q**T with the reflectors of DGEQRF (device pointers)
     integer(c_int) function lb200_dormqr(stream, side, trans, m, n, k, dA, lda, dtau, dC, ldc) bind(C, name="lb200_dormqr")
       import :: c_int, c_long_long, c_ptr, c_char
       type(c_ptr), value :: stream, dA, dtau, dC
       character(kind=c_char), value :: side, trans
       integer(c_int), value :: m, n, k
       integer(c_long_long), value :: lda, ldc
     end function
     ! SRC/dorgqr.f:126  first N columns of Q in place over the reflectors
     integer(c_int) function lb200_dorgqr(stream, m, n, k, dA, lda, dtau) bind(C, name="lb200_dorgqr")
       import :: c_int, c_long_long, c_ptr
       type(c_ptr), value :: stream, dA, dtau
       integer(c_int), value :: m, n, k
       integer(c_long_long), value :: lda
     end function
     ! SRC/dgetri.f:114  inverse from the DGETRF factors; dinfo = i if U(i,i) is exactly zero
     integer(c_int) function lb200_dgetri(stream, n, dA, lda, dipiv, dinfo) bind(C, name="lb200_dgetri")
       import :: c_int, c_long_long, c_ptr
       type(c_ptr), value :: stream, dA, dipiv, dinfo
       integer(c_int), value :: n
       integer(c_long_long), value :: lda
     end function
     ! SRC/dgeqrt.f:139  blocked QR keeping the compact-WY factors T (nb x min(m,n))
     integer(c_int) function lb200_dgeqrt(stream, m, n, nb, dA, lda, dT, ldt) bind(C, name="lb200_dgeqrt")
       import :: c_int, c_long_long, c_ptr
       type(c_ptr), value :: stream, dA, dT
       integer(c_int), value :: m, n, nb
       integer(c_long_long), value :: lda, ldt
     end function
     ! SRC/dgemqrt.f:166  apply Q or Q**T from DGEQRT
     integer(c_int) function lb200_dgemqrt(stream, side, trans, m, n, k, nb, dV, ldv, dT, ldt, dC, ldc) &
          bind(C, name="lb200_dgemqrt")
       import :: c_int, c_long_long, c_ptr, c_char
       type(c_ptr), value :: stream, dV, dT, dC
       character(kind=c_char), value :: side, trans
       integer(c_int), value :: m, n, k, nb
       integer(c_long_long), value :: ldv, ldt, ldc
     end function
  end interface
end module lapack_b200
