! beom_gpu_mod.f95 -- ISO_C_BINDING view of include/beom_gpu.h for the unchanged BEOM driver.
!
! This is the binding a maintainer of zhazorken/beom adds next to shared_mod.f95 / private_mod.f95 so
! that integrate_time (private_mod.f95:1840-1919) runs its step routines on the GPU.  It could not be
! compiled in the development environment (no Fortran compiler exists there, SURVEY.md section 0); it
! is written against the Fortran 2003 standard only (bind(C), c_ptr, c_loc, value arguments).
!
! How it is used is shown in INTEGRATION.md: three call sites of private_mod.f95 change, main.f95 and
! shared_mod.f95 stay byte-for-byte the same (plus the svis/tdrg/topt declarations that every build of
! the fork needs, SURVEY.md section 0).

module beom_gpu_mod
  use, intrinsic :: iso_c_binding
  implicit none
  private

  integer, parameter, public :: beom_maxlay = 16

! struct beom_params (include/beom_gpu.h): the parameter set of shared_mod.f95:41-111 by value.
  type, bind(C), public :: beom_params
    integer(c_int32_t) :: lm, mm, nlay, ndeg
    real(c_double)     :: dl, cext, f0
    real(c_double)     :: rhon(beom_maxlay), topl(beom_maxlay)
    real(c_double)     :: dt_s, dt_o, dt_r, dt3d, bvis, dvis, bdrg, hmin, hsbl, hbbl
    real(c_double)     :: g_fb, uadv, qdrg, ocrp, rsta, xper, yper, diag, rgld, mcbc
    real(c_double)     :: tauw(2)
    real(c_double)     :: svis, tdrg, topt, plum
    real(c_double)     :: dt, hsal, hdry, tole, pi, grav, rho0, beta, epsi, gamm, del1, del2, sor
    integer(c_int32_t) :: itmx, nsal, variant, reserved_
  end type beom_params

! struct beom_fields: addresses of the module arrays of private_mod.f95:27-93 (c_loc of element 0).
  integer, parameter, public :: BEOM_MAXLAY = 16
  ! one set of output records in page-locked host memory (beom_gpu_records_wait)
  type, bind(C), public :: beom_records
    type(c_ptr) :: eta, u, v            ! real(c_float) (count, nlay)
    type(c_ptr) :: pvor, mont, v_cc     ! c_null_ptr unless begun with_diag
    integer(c_int) :: first_point, count
    real(c_double) :: hmin(BEOM_MAXLAY), hmax(BEOM_MAXLAY)
    integer(c_int) :: thin_layer        ! first layer thinner than hmin / 2 somewhere wet, 0 = none (private_mod.f95:2798-2808)
  end type beom_records

  type, bind(C), public :: beom_fields
    type(c_ptr) :: neig, subc
    type(c_ptr) :: mk_u, mk_v, mk_n, mkpe, mkpi
    type(c_ptr) :: fcor, h_th
    type(c_ptr) :: nudg, fnud, hdot, taus, tide, bodf
    type(c_ptr) :: segm
    type(c_ptr) :: Ow, Os, Osum_, pi_s
    integer(c_int32_t) :: nseg, flag_nudging
    real(c_double)     :: invf, w_ti
  end type beom_fields

  type, bind(C), public :: beom_gpu_options
    integer(c_int32_t) :: device, fused, rank, nranks, strict
    integer(c_int32_t) :: reserved_(3)
  end type beom_gpu_options

  public :: beom_gpu_default_options, beom_gpu_init, beom_gpu_upload_state, beom_gpu_stress, &
            beom_gpu_step, beom_gpu_download_state, beom_gpu_download_pi_s, beom_gpu_download_diag, beom_gpu_diagnostics, &
            beom_gpu_sync, beom_gpu_finalize, beom_gpu_last_error, beom_gpu_check

  interface
    subroutine beom_gpu_default_options(opt) bind(C, name = 'beom_gpu_default_options')
      import :: beom_gpu_options
      type(beom_gpu_options), intent(out) :: opt
    end subroutine

    function beom_gpu_init(par, fld, opt) bind(C, name = 'beom_gpu_init') result(rc)
      import :: beom_params, beom_fields, beom_gpu_options, c_int
      type(beom_params),      intent(in) :: par
      type(beom_fields),      intent(in) :: fld
      type(beom_gpu_options), intent(in) :: opt
      integer(c_int) :: rc
    end function

!   hlay, u, v are passed as the address of element (0,1): the whole (0:ndeg,nlay) arrays.
    function beom_gpu_upload_state(hlay, u, v) bind(C, name = 'beom_gpu_upload_state') result(rc)
      import :: c_double, c_int
      real(c_double), intent(in) :: hlay(*), u(*), v(*)
      integer(c_int) :: rc
    end function

    function beom_gpu_stress() bind(C, name = 'beom_gpu_stress') result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function

    function beom_gpu_step(tstp, ctim, ramp, gene, upst, first_three) bind(C, name = 'beom_gpu_step') result(rc)
      import :: c_int, c_double
      integer(c_int), value :: tstp, upst, first_three
      real(c_double), value :: ctim, ramp, gene
      integer(c_int) :: rc
    end function

    function beom_gpu_download_state(hlay, u, v) bind(C, name = 'beom_gpu_download_state') result(rc)
      import :: c_double, c_int
      real(c_double), intent(inout) :: hlay(*), u(*), v(*)
      integer(c_int) :: rc
    end function

    function beom_gpu_download_pi_s(pi_s) bind(C, name = 'beom_gpu_download_pi_s') result(rc)
      import :: c_double, c_int
      real(c_double), intent(inout) :: pi_s(*)
      integer(c_int) :: rc
    end function

    function beom_gpu_download_diag(pvor, mont, v_cc) bind(C, name = 'beom_gpu_download_diag') result(rc)
      import :: c_float, c_int
      real(c_float), intent(inout) :: pvor(*), mont(*), v_cc(*)
      integer(c_int) :: rc
    end function

    ! conservation integrals of testcases/conservation.m:116-211 (vol(nlay), ke(nlay), pe(1)); h_0(0:ndeg, nlay)
    function beom_gpu_diagnostics(h_0, vol, ke, pe) bind(C, name = 'beom_gpu_diagnostics') result(rc)
      import :: c_double, c_int
      real(c_double), intent(in) :: h_0(*)
      real(c_double), intent(inout) :: vol(*), ke(*), pe(*)
      integer(c_int) :: rc
    end function

    ! the same plus potential enstrophy and relative vorticity (conservation.m:169-211): enst, zeta, zeta2 (nlay each), npts(1)
    function beom_gpu_diagnostics_all(h_0, vol, ke, pe, enst, zeta, zeta2, npts) bind(C, name = 'beom_gpu_diagnostics_all') result(rc)
      import :: c_double, c_int
      real(c_double), intent(in) :: h_0(*)
      real(c_double), intent(inout) :: vol(*), ke(*), pe(*), enst(*), zeta(*), zeta2(*), npts(*)
      integer(c_int) :: rc
    end function

    ! write_array's records made on the device and copied out asynchronously (private_mod.f95:2817-2974):
    !   call beom_gpu_set_rest_thickness(h_0 as written to h_0.bin)          once
    !   rc = beom_gpu_records_begin(with_diag)                                at an output step; returns at once
    !   rc = beom_gpu_records_wait(rec)                                        before writing: c_f_pointer(rec%eta, ior4, (/ndeg, nlay/)) ...
    function beom_gpu_set_rest_thickness(h_0_r4) bind(C, name = 'beom_gpu_set_rest_thickness') result(rc)
      import :: c_float, c_int
      real(c_float), intent(in) :: h_0_r4(*)
      integer(c_int) :: rc
    end function
    function beom_gpu_records_begin(with_diag) bind(C, name = 'beom_gpu_records_begin') result(rc)
      import :: c_int
      integer(c_int), value :: with_diag
      integer(c_int) :: rc
    end function
    function beom_gpu_records_wait(rec) bind(C, name = 'beom_gpu_records_wait') result(rc)
      import :: beom_records, c_int
      type(beom_records), intent(out) :: rec
      integer(c_int) :: rc
    end function

    ! sweeps of the last surf_pressure solve (the reference's `iters', private_mod.f95:1756-1803)
    function beom_gpu_pi_iterations(iters) bind(C, name = 'beom_gpu_pi_iterations') result(rc)
      import :: c_int
      integer(c_int), intent(out) :: iters
      integer(c_int) :: rc
    end function

    function beom_gpu_sync() bind(C, name = 'beom_gpu_sync') result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function

    function beom_gpu_finalize() bind(C, name = 'beom_gpu_finalize') result(rc)
      import :: c_int
      integer(c_int) :: rc
    end function

    function beom_gpu_last_error(buf, length) bind(C, name = 'beom_gpu_last_error') result(n)
      import :: c_char, c_int
      character(kind = c_char), intent(inout) :: buf(*)
      integer(c_int), value :: length
      integer(c_int) :: n
    end function
  end interface

contains

! The reference's error convention (shared_mod.f95:113-157): a non-zero code goes to errc, the text is
! appended to errm and quit() ends the run.
  subroutine beom_gpu_check(rc, where)
    use shared_mod, only: errc, errm, quit
    integer(c_int),   intent(in) :: rc
    character(len=*), intent(in) :: where
    character(kind = c_char)     :: buf(512)
    character(len = 512)         :: text
    integer                      :: i, n
    if (rc == 0) return
    n = beom_gpu_last_error(buf, 512_c_int)
    text = ' '
    do i = 1, min(n, 511)
      text(i:i) = buf(i)
    end do
    errc = int(rc)
    errm = trim(errm) // ' ' // where // ': ' // trim(text)
    call quit()
  end subroutine beom_gpu_check

end module beom_gpu_mod
