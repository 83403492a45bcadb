!! tfqmrgpu_Fortran.h - constants of the Fortran interface of libtfQMRgpu (B200-native implementation).
!! Same names and values as the reference's include/tfqmrgpu_Fortran.h:4-30 (they are part of the ABI: status codes and
!! layout keys cross the by-reference shims of tfqmrgpu_Fortran_wrappers.c unchanged).  Include inside a declaration part.

      !! status codes (tfqmrgpu.h:160-180)
      integer(kind=4), parameter :: TFQMRGPU_STATUS_SUCCESS           = 0
      integer(kind=4), parameter :: TFQMRGPU_STATUS_LAUNCH_FAILED     = 2
      integer(kind=4), parameter :: TFQMRGPU_STATUS_NO_INFO_PASSED    = 3
      integer(kind=4), parameter :: TFQMRGPU_STATUS_ALLOCATION_FAILED = 4
      integer(kind=4), parameter :: TFQMRGPU_STATUS_BREAKDOWN         = 6
      integer(kind=4), parameter :: TFQMRGPU_POINTER_INVALID          = 7
      integer(kind=4), parameter :: TFQMRGPU_STATUS_MAX_ITERATIONS    = 9
      integer(kind=4), parameter :: TFQMRGPU_B_HAS_A_ZERO_COLUMN      = 11
      integer(kind=4), parameter :: TFQMRGPU_BLOCKSIZE_MISSING        = 12
      integer(kind=4), parameter :: TFQMRGPU_B_IS_NOT_SUBSET_OF_X     = 13
      integer(kind=4), parameter :: TFQMRGPU_UNDOCUMENTED_ERROR       = 14
      integer(kind=4), parameter :: TFQMRGPU_DATALAYOUT_UNKNOWN       = 15
      integer(kind=4), parameter :: TFQMRGPU_PRECISION_MISSMATCH      = 16
      integer(kind=4), parameter :: TFQMRGPU_TANSPOSITION_UNKNOWN     = 17
      integer(kind=4), parameter :: TFQMRGPU_VARIABLENAME_UNKNOWN     = 18
      integer(kind=4), parameter :: TFQMRGPU_NO_IMPLEMENTATION        = 19
      integer(kind=4), parameter :: TFQMRGPU_CODE_LINE                = 1000
      integer(kind=4), parameter :: TFQMRGPU_CODE_CHAR                = 10000000

      !! block data layouts (tfqmrgpu.h:184-186)
      integer(kind=4), parameter :: TFQMRGPU_LAYOUT_RRRRIIII = 15 !! real and imaginary planes of a block separated (device layout)
      integer(kind=4), parameter :: TFQMRGPU_LAYOUT_RRIIRRII = 51 !! real and imaginary parts separated row by row
      integer(kind=4), parameter :: TFQMRGPU_LAYOUT_RIRIRIRI = 85 !! interleaved: the layout of Fortran complex arrays
      integer(kind=4), parameter :: TFQMRGPU_LAYOUT_DEFAULT  = 85

      !! opaque pointers are 64-bit integers on the Fortran side
      integer, parameter :: TFQMRGPU_HANDLE_KIND = 8
      integer, parameter :: TFQMRGPU_PLAN_KIND   = 8
      integer, parameter :: TFQMRGPU_PTR_KIND    = 8
      integer, parameter :: cuda_stream_kind     = 8   !! a cudaStream_t without the CUDA headers; 0 = default stream
