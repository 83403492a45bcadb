!! tfqmrgpu_Fortran_module.F90 - Fortran 2003 module for libtfQMRgpu (B200-native implementation).
!!
!! Role of the reference's include/tfqmrgpu_Fortran_module.F90: the generic names
!!     create, destroy, free, set, get, solve, print_error
!! and the quick starters tfqmrgpu_bsrsv_rectangular / tfqmrgpu_bsrsv_complete (reference lines 12-59, 294-462), so that
!! existing callers (`use tfqmrgpu`) compile unchanged.  Every status is returned in the trailing `ierr` argument.
!! This module is written against the 18 by-reference shims of libtfQMRgpu.so / libtfQMRgpu_Fortran.a
!! (tfqmrgpu_b200/csrc/fortran_wrappers.c; the shims shift the 1-based BSR index arrays themselves) through explicit
!! bind(C) interfaces, i.e. independent of the compiler's name mangling.
!! NOT compiled in this repository's build (the image has no Fortran compiler); tests/test_gpu_parity.py
!! (test_fortran_shims_full_solve) drives exactly the call sequence of tfqmrgpu_bsrsv_rectangular through the same shims.
!!     gfortran -cpp -I include -c include/tfqmrgpu_Fortran_module.F90 ; link with -ltfQMRgpu
module tfqmrgpu
  use, intrinsic :: iso_c_binding, only: c_int32_t, c_int64_t, c_double, c_char, c_size_t
  implicit none
  private

  include "tfqmrgpu_Fortran.h"

  public :: print_error, create, destroy, free, set, get, solve
  public :: tfqmrgpu_bsrsv_rectangular, tfqmrgpu_bsrsv_complete
  public :: TFQMRGPU_HANDLE_KIND, TFQMRGPU_PLAN_KIND, TFQMRGPU_PTR_KIND, cuda_stream_kind
  public :: TFQMRGPU_LAYOUT_RIRIRIRI, TFQMRGPU_LAYOUT_RRIIRRII, TFQMRGPU_LAYOUT_RRRRIIII, TFQMRGPU_LAYOUT_DEFAULT
  public :: TFQMRGPU_STATUS_SUCCESS, TFQMRGPU_STATUS_MAX_ITERATIONS, TFQMRGPU_STATUS_BREAKDOWN

  !! ---- the shims (all arguments by reference, status last) -----------------------------------------------------
  interface
    subroutine shim_print_error(status, stat) bind(C, name="tfqmrgpuprinterror_")
      import :: c_int32_t
      integer(c_int32_t), intent(in) :: status
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_create_handle(handle, stat) bind(C, name="tfqmrgpucreatehandle_")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(out) :: handle
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_destroy_handle(handle, stat) bind(C, name="tfqmrgpudestroyhandle_")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(inout) :: handle
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_set_stream(handle, stream, stat) bind(C, name="tfqmrgpusetstream_")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: handle, stream
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_get_stream(handle, stream, stat) bind(C, name="tfqmrgpugetstream_")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: handle
      integer(c_int64_t), intent(out) :: stream
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_create_plan(handle, plan, mb, rpA, nnzbA, ciA, rpX, nnzbX, ciX, rpB, nnzbB, ciB, echo, stat) &
        bind(C, name="tfqmrgpu_bsrsv_createplan_")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: handle
      integer(c_int64_t), intent(out) :: plan
      integer(c_int32_t), intent(in) :: mb, nnzbA, nnzbX, nnzbB, echo
      integer(c_int32_t), intent(in) :: rpA(*), ciA(*), rpX(*), ciX(*), rpB(*), ciB(*)
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_destroy_plan(handle, plan, stat) bind(C, name="tfqmrgpu_bsrsv_destroyplan_")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: handle
      integer(c_int64_t), intent(inout) :: plan
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_buffer_size(handle, plan, ldA, blockDim, ldB, rhsBlockDim, precision, bytes, stat) &
        bind(C, name="tfqmrgpu_bsrsv_buffersize_")
      import :: c_int32_t, c_int64_t, c_char, c_size_t
      integer(c_int64_t), intent(in) :: handle, plan
      integer(c_int32_t), intent(in) :: ldA, blockDim, ldB, rhsBlockDim
      character(kind=c_char), intent(in) :: precision
      integer(c_size_t), intent(out) :: bytes
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_create_workspace(buffer, bytes, stat) bind(C, name="tfqmrgpucreateworkspace_")
      import :: c_int32_t, c_int64_t, c_size_t
      integer(c_int64_t), intent(out) :: buffer
      integer(c_size_t), intent(in) :: bytes
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_destroy_workspace(buffer, stat) bind(C, name="tfqmrgpudestroyworkspace_")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: buffer
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_set_buffer(handle, plan, buffer, stat) bind(C, name="tfqmrgpu_bsrsv_setbuffer_")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: handle, plan, buffer
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_get_buffer(handle, plan, buffer, stat) bind(C, name="tfqmrgpu_bsrsv_getbuffer_")
      import :: c_int32_t, c_int64_t
      integer(c_int64_t), intent(in) :: handle, plan
      integer(c_int64_t), intent(out) :: buffer
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_set_matrix_z(handle, plan, var, val, ld, d2, trans, layout, stat) bind(C, name="tfqmrgpu_bsrsv_setmatrix_z_")
      import :: c_int32_t, c_int64_t, c_char
      integer(c_int64_t), intent(in) :: handle, plan
      character(kind=c_char), intent(in) :: var, trans
      type(*), intent(in) :: val(*)            !! complex(8) data as it lies in memory
      integer(c_int32_t), intent(in) :: ld, d2, layout
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_set_matrix_c(handle, plan, var, val, ld, d2, trans, layout, stat) bind(C, name="tfqmrgpu_bsrsv_setmatrix_c_")
      import :: c_int32_t, c_int64_t, c_char
      integer(c_int64_t), intent(in) :: handle, plan
      character(kind=c_char), intent(in) :: var, trans
      type(*), intent(in) :: val(*)            !! complex(4) data as it lies in memory
      integer(c_int32_t), intent(in) :: ld, d2, layout
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_get_matrix_z(handle, plan, var, val, ld, d2, trans, layout, stat) bind(C, name="tfqmrgpu_bsrsv_getmatrix_z_")
      import :: c_int32_t, c_int64_t, c_char
      integer(c_int64_t), intent(in) :: handle, plan
      character(kind=c_char), intent(in) :: var, trans
      type(*) :: val(*)
      integer(c_int32_t), intent(in) :: ld, d2, layout
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_get_matrix_c(handle, plan, var, val, ld, d2, trans, layout, stat) bind(C, name="tfqmrgpu_bsrsv_getmatrix_c_")
      import :: c_int32_t, c_int64_t, c_char
      integer(c_int64_t), intent(in) :: handle, plan
      character(kind=c_char), intent(in) :: var, trans
      type(*) :: val(*)
      integer(c_int32_t), intent(in) :: ld, d2, layout
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_solve(handle, plan, threshold, maxIterations, stat) bind(C, name="tfqmrgpu_bsrsv_solve_")
      import :: c_int32_t, c_int64_t, c_double
      integer(c_int64_t), intent(in) :: handle, plan
      real(c_double), intent(in) :: threshold
      integer(c_int32_t), intent(in) :: maxIterations
      integer(c_int32_t), intent(out) :: stat
    end subroutine
    subroutine shim_get_info(handle, plan, residuum, iterations, flops, flops_all, stat) bind(C, name="tfqmrgpu_bsrsv_getinfo_")
      import :: c_int32_t, c_int64_t, c_double
      integer(c_int64_t), intent(in) :: handle, plan
      real(c_double), intent(out) :: residuum, flops, flops_all
      integer(c_int32_t), intent(out) :: iterations
      integer(c_int32_t), intent(out) :: stat
    end subroutine
  end interface

  !! ---- the generic names of the reference module -------------------------------------------------------------------
  interface create
    module procedure new_handle, new_plan, new_workspace
  end interface
  interface destroy
    module procedure delete_handle, delete_plan
  end interface
  interface free
    module procedure delete_workspace
  end interface
  interface set
    module procedure put_stream, put_buffer, put_matrix_z, put_matrix_c
  end interface
  interface get
    module procedure fetch_stream, fetch_buffer_size, fetch_buffer, fetch_matrix_z, fetch_matrix_c, fetch_info
  end interface
  interface solve
    module procedure run_solve, tfqmrgpu_bsrsv_complete, tfqmrgpu_bsrsv_rectangular
  end interface

contains

  subroutine print_error(status, ierr)
    integer(kind=4), intent(in) :: status
    integer(kind=4), intent(out) :: ierr
    call shim_print_error(status, ierr)
  end subroutine

  subroutine new_handle(handle, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(out) :: handle
    integer(kind=4), intent(out) :: ierr
    call shim_create_handle(handle, ierr)
  end subroutine

  subroutine delete_handle(handle, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(inout) :: handle
    integer(kind=4), intent(out) :: ierr
    call shim_destroy_handle(handle, ierr)
  end subroutine

  subroutine put_stream(handle, streamId, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=cuda_stream_kind), intent(in) :: streamId
    integer(kind=4), intent(out) :: ierr
    call shim_set_stream(handle, streamId, ierr)
  end subroutine

  subroutine fetch_stream(handle, streamId, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=cuda_stream_kind), intent(out) :: streamId
    integer(kind=4), intent(out) :: ierr
    call shim_get_stream(handle, streamId, ierr)
  end subroutine

  !! analyse the BSR patterns of A, X and B (1-based row pointers and column indices, as Fortran holds them)
  subroutine new_plan(handle, plan, mb, rowPtrA, nnzbA, colIndA, rowPtrX, nnzbX, colIndX, rowPtrB, nnzbB, colIndB, echo, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=TFQMRGPU_PLAN_KIND), intent(out) :: plan
    integer(kind=4), intent(in) :: mb, nnzbA, nnzbX, nnzbB, echo
    integer(kind=4), intent(in) :: rowPtrA(*), colIndA(*), rowPtrX(*), colIndX(*), rowPtrB(*), colIndB(*)
    integer(kind=4), intent(out) :: ierr
    call shim_create_plan(handle, plan, mb, rowPtrA, nnzbA, colIndA, rowPtrX, nnzbX, colIndX, rowPtrB, nnzbB, colIndB, echo, ierr)
  end subroutine

  subroutine delete_plan(handle, plan, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=TFQMRGPU_PLAN_KIND), intent(inout) :: plan
    integer(kind=4), intent(out) :: ierr
    call shim_destroy_plan(handle, plan, ierr)
  end subroutine

  subroutine fetch_buffer_size(handle, plan, ldA, blockDim, ldB, RhsBlockDim, doublePrecision, pBufferSizeInBytes, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=TFQMRGPU_PLAN_KIND), intent(in) :: plan
    integer(kind=4), intent(in) :: ldA, blockDim, ldB, RhsBlockDim
    character, intent(in) :: doublePrecision      !! 'z' | 'c'
    integer(kind=8), intent(out) :: pBufferSizeInBytes
    integer(kind=4), intent(out) :: ierr
    integer(c_size_t) :: bytes
    call shim_buffer_size(handle, plan, ldA, blockDim, ldB, RhsBlockDim, doublePrecision, bytes, ierr)
    pBufferSizeInBytes = int(bytes, kind=8)
  end subroutine

  subroutine new_workspace(pBuffer, pBufferSizeInBytes, ierr)
    integer(kind=TFQMRGPU_PTR_KIND), intent(out) :: pBuffer
    integer(kind=8), intent(in) :: pBufferSizeInBytes
    integer(kind=4), intent(out) :: ierr
    call shim_create_workspace(pBuffer, int(pBufferSizeInBytes, kind=c_size_t), ierr)
  end subroutine

  subroutine delete_workspace(pBuffer, ierr)
    integer(kind=TFQMRGPU_PTR_KIND), intent(in) :: pBuffer
    integer(kind=4), intent(out) :: ierr
    call shim_destroy_workspace(pBuffer, ierr)
  end subroutine

  subroutine put_buffer(handle, plan, pBuffer, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=TFQMRGPU_PLAN_KIND), intent(in) :: plan
    integer(kind=TFQMRGPU_PTR_KIND), intent(in) :: pBuffer
    integer(kind=4), intent(out) :: ierr
    call shim_set_buffer(handle, plan, pBuffer, ierr)
  end subroutine

  subroutine fetch_buffer(handle, plan, pBuffer, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=TFQMRGPU_PLAN_KIND), intent(in) :: plan
    integer(kind=TFQMRGPU_PTR_KIND), intent(out) :: pBuffer
    integer(kind=4), intent(out) :: ierr
    call shim_get_buffer(handle, plan, pBuffer, ierr)
  end subroutine

  !! operands: complex arrays are passed as they lie in memory (RIRIRIRI); the shims' `val` is assumed-type (Fortran 2018)
  subroutine put_matrix_z(handle, plan, var, val, ld, d2, trans, layout, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=TFQMRGPU_PLAN_KIND), intent(in) :: plan
    character, intent(in) :: var, trans
    complex(kind=8), intent(in) :: val(*)
    integer(kind=4), intent(in) :: ld, d2, layout
    integer(kind=4), intent(out) :: ierr
    call shim_set_matrix_z(handle, plan, var, val, ld, d2, trans, layout, ierr)
  end subroutine

  subroutine put_matrix_c(handle, plan, var, val, ld, d2, trans, layout, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=TFQMRGPU_PLAN_KIND), intent(in) :: plan
    character, intent(in) :: var, trans
    complex(kind=4), intent(in) :: val(*)
    integer(kind=4), intent(in) :: ld, d2, layout
    integer(kind=4), intent(out) :: ierr
    call shim_set_matrix_c(handle, plan, var, val, ld, d2, trans, layout, ierr)
  end subroutine

  subroutine fetch_matrix_z(handle, plan, var, val, ld, d2, trans, layout, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=TFQMRGPU_PLAN_KIND), intent(in) :: plan
    character, intent(in) :: var, trans
    complex(kind=8), intent(inout) :: val(*)
    integer(kind=4), intent(in) :: ld, d2, layout
    integer(kind=4), intent(out) :: ierr
    call shim_get_matrix_z(handle, plan, var, val, ld, d2, trans, layout, ierr)
  end subroutine

  subroutine fetch_matrix_c(handle, plan, var, val, ld, d2, trans, layout, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=TFQMRGPU_PLAN_KIND), intent(in) :: plan
    character, intent(in) :: var, trans
    complex(kind=4), intent(inout) :: val(*)
    integer(kind=4), intent(in) :: ld, d2, layout
    integer(kind=4), intent(out) :: ierr
    call shim_get_matrix_c(handle, plan, var, val, ld, d2, trans, layout, ierr)
  end subroutine

  subroutine run_solve(handle, plan, threshold, maxIterations, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=TFQMRGPU_PLAN_KIND), intent(in) :: plan
    real(kind=8), intent(in) :: threshold
    integer(kind=4), intent(in) :: maxIterations
    integer(kind=4), intent(out) :: ierr
    call shim_solve(handle, plan, threshold, maxIterations, ierr)
  end subroutine

  subroutine fetch_info(handle, plan, residual_reached, iterations_needed, flops_performed, flops_performed_all, ierr)
    integer(kind=TFQMRGPU_HANDLE_KIND), intent(in) :: handle
    integer(kind=TFQMRGPU_PLAN_KIND), intent(in) :: plan
    real(kind=8), intent(out) :: residual_reached, flops_performed, flops_performed_all
    integer(kind=4), intent(out) :: iterations_needed
    integer(kind=4), intent(out) :: ierr
    call shim_get_info(handle, plan, residual_reached, iterations_needed, flops_performed, flops_performed_all, ierr)
  end subroutine

  !! ---- quick starters: the whole workflow in one call, complex double precision -----------------------------------------
  !! Solves A * X = B for BSR operands with ldA x ldA blocks in A and ldB x ldA blocks (ldB right-hand sides per block column,
  !! stored Xmat(ldB, ldA, nnzbX)) in X and B.  On entry `iterations` / `residual` are the limits, on exit what was reached.
  !! `o` > 0: unit for a one-line report; ierr /= 0 on entry switches the plan analysis to verbose.  A status of
  !! MAX_ITERATIONS (9) or BREAKDOWN (6) from the solver is returned in ierr with X downloaded; any other failure returns
  !! at once with everything released (the reference stops the program instead).
  subroutine tfqmrgpu_bsrsv_rectangular(mb, ldA, ldB, rowPtrA, colIndA, Amat, transA, rowPtrX, colIndX, Xmat, transX, &
                                        rowPtrB, colIndB, Bmat, transB, iterations, residual, o, ierr)
    integer(kind=4), intent(in) :: mb, ldA, ldB
    integer(kind=4), intent(in) :: rowPtrA(:), rowPtrX(:), rowPtrB(:), colIndA(:), colIndX(:), colIndB(:)
    character, intent(in) :: transA, transX, transB
    complex(kind=8), intent(in) :: Amat(ldA, ldA, *), Bmat(ldB, ldA, *)
    complex(kind=8), intent(out) :: Xmat(ldB, ldA, *)
    integer(kind=4), intent(inout) :: iterations
    real(kind=8), intent(inout) :: residual
    integer(kind=4), intent(in) :: o
    integer(kind=4), intent(inout) :: ierr

    integer(kind=TFQMRGPU_HANDLE_KIND) :: handle
    integer(kind=TFQMRGPU_PLAN_KIND) :: plan
    integer(kind=TFQMRGPU_PTR_KIND) :: workspace
    integer(kind=cuda_stream_kind), parameter :: default_stream = 0
    integer(kind=8) :: bytes
    integer(kind=4) :: echo, limit, status, ignore
    real(kind=8) :: tolerance, flops, flops_all
    logical :: have_handle, have_plan, have_workspace

    echo = 0 ; if (0 /= ierr) echo = 9
    limit = iterations ; tolerance = residual
    have_handle = .false. ; have_plan = .false. ; have_workspace = .false.
    status = TFQMRGPU_STATUS_SUCCESS

    call create(handle, ierr) ; have_handle = (0 == ierr)
    if (0 == ierr) call set(handle, default_stream, ierr)
    if (0 == ierr) then
      call create(handle, plan, mb, rowPtrA, size(colIndA), colIndA, rowPtrX, size(colIndX), colIndX, &
                  rowPtrB, size(colIndB), colIndB, echo, ierr)
      have_plan = (0 == ierr)
    endif
    if (0 == ierr) call get(handle, plan, ldA, ldA, ldB, ldB, 'z', bytes, ierr)
    if (0 == ierr) then
      call create(workspace, bytes, ierr) ; have_workspace = (0 == ierr)
    endif
    if (0 == ierr) call set(handle, plan, workspace, ierr)
    if (0 == ierr) call set(handle, plan, 'A', Amat(:, 1, 1), ldA, ldA, transA, TFQMRGPU_LAYOUT_RIRIRIRI, ierr)
    if (0 == ierr) call set(handle, plan, 'B', Bmat(:, 1, 1), ldB, ldA, transB, TFQMRGPU_LAYOUT_RIRIRIRI, ierr)
    if (0 == ierr) then
      call solve(handle, plan, tolerance, limit, ierr)
      if (TFQMRGPU_STATUS_MAX_ITERATIONS == ierr .or. TFQMRGPU_STATUS_BREAKDOWN == ierr) then
        status = ierr ; ierr = 0          !! not converged: still report and download what was reached
      endif
    endif
    if (0 == ierr) call get(handle, plan, residual, iterations, flops, flops_all, ierr)
    if (0 == ierr) call get(handle, plan, 'X', Xmat(:, 1, 1), ldB, ldA, transX, TFQMRGPU_LAYOUT_RIRIRIRI, ierr)
    if (0 /= ierr) call print_error(ierr, ignore)

    if (have_workspace) call free(workspace, ignore)
    if (have_plan) call destroy(handle, plan, ignore)
    if (have_handle) call destroy(handle, ignore)
    if (0 == ierr) ierr = status
    if (o > 0 .and. (0 == ierr .or. ierr == status)) write(o, "(2(a,es8.1),2(a,i0),a)") &
        " tfqmrgpu_bsrsv reached", residual, " (limit", tolerance, ") in ", iterations, " (limit ", limit, ") iterations."
  end subroutine

  !! square blocks: ldB = ldA
  subroutine tfqmrgpu_bsrsv_complete(mb, ldA, rowPtrA, colIndA, Amat, transA, rowPtrX, colIndX, Xmat, transX, &
                                     rowPtrB, colIndB, Bmat, transB, iterations, residual, o, ierr)
    integer(kind=4), intent(in) :: mb, ldA
    integer(kind=4), intent(in) :: rowPtrA(:), rowPtrX(:), rowPtrB(:), colIndA(:), colIndX(:), colIndB(:)
    character, intent(in) :: transA, transX, transB
    complex(kind=8), intent(in) :: Amat(ldA, ldA, *), Bmat(ldA, ldA, *)
    complex(kind=8), intent(out) :: Xmat(ldA, ldA, *)
    integer(kind=4), intent(inout) :: iterations
    real(kind=8), intent(inout) :: residual
    integer(kind=4), intent(in) :: o
    integer(kind=4), intent(inout) :: ierr
    call tfqmrgpu_bsrsv_rectangular(mb, ldA, ldA, rowPtrA, colIndA, Amat, transA, rowPtrX, colIndX, Xmat, transX, &
                                    rowPtrB, colIndB, Bmat, transB, iterations, residual, o, ierr)
  end subroutine

end module tfqmrgpu
