! pamg_iface.F90 -- ISO_C_BINDING interfaces to libpamg_cuda.so (include/pamg.h) for the reference's own
! Fortran driver (main.F90 / transport_tri_semi.F90).  NOT COMPILED OR TESTED IN THIS REPOSITORY: the build
! image has no Fortran compiler.  The identical call sequence is exercised from C++ (host/pamg_host.cpp)
! and Python (pamg.py).  See INTEGRATION.md for where each call replaces a contained procedure of
! Semi_implicit_iterative.
!
! Layout notes: tracer(ilevel)%tnew(nloc, totele_str, totele_unst) is passed as is (column-major, contiguous);
! meshList(:)%X / Neig / fNeig / Dir are gathered once into X(2,3,U), neig(3,U), fneig(3,U), dir(3,U)
! (Dir: .true. -> 1).  All routines return 0 on success, < 0 on error (like ierr / errorflag).
module pamg_iface
  use, intrinsic :: iso_c_binding
  implicit none

  integer(c_int), parameter :: PAMG_TNEW = 0, PAMG_TOLD = 1, PAMG_RHS = 2, PAMG_RES = 3, PAMG_TNONLIN = 5

  type, bind(c) :: pamg_params
    integer(c_int32_t) :: n_split, multi_levels, n_smooth, n_multigrid, n_coarse_smooth, solver
    integer(c_int32_t) :: face_terms, literal_source, transfer, residual_sign, halo_rule, coarse_bc_zero
    integer(c_int32_t) :: keep_tnew_gs, reserved
    real(c_double) :: theta, dt, k, omega, u_x, u_y, source_coef
  end type pamg_params

  interface
    subroutine pamg_default_params(p, literal_head) bind(c, name="pamg_default_params")
      import :: pamg_params, c_int
      type(pamg_params), intent(out) :: p
      integer(c_int), value :: literal_head
    end subroutine

    integer(c_int) function pamg_create(p, device, handle) bind(c, name="pamg_create")
      import :: pamg_params, c_int, c_ptr
      type(pamg_params), intent(in) :: p
      integer(c_int), value :: device
      type(c_ptr), intent(out) :: handle
    end function

    subroutine pamg_destroy(handle) bind(c, name="pamg_destroy")
      import :: c_ptr
      type(c_ptr), value :: handle
    end subroutine

    integer(c_int) function pamg_set_parents(handle, U, X, neig, fneig, dir) bind(c, name="pamg_set_parents")
      import :: c_int, c_ptr, c_double, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int), value :: U
      real(c_double), intent(in) :: X(2, 3, *)
      integer(c_int32_t), intent(in) :: neig(3, *), fneig(3, *), dir(3, *)
    end function

    integer(c_int) function pamg_upload_field(handle, field, level, host) bind(c, name="pamg_upload_field")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: field, level
      real(c_double), intent(in) :: host(*)
    end function

    integer(c_int) function pamg_download_field(handle, field, level, host) bind(c, name="pamg_download_field")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: field, level
      real(c_double), intent(out) :: host(*)
    end function

    integer(c_int) function pamg_copy_field(handle, level, dst_field, src_field) bind(c, name="pamg_copy_field")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int), value :: level, dst_field, src_field
    end function

    ! update_overlaps (splitting.F90:1210)
    integer(c_int) function pamg_update_overlaps(handle, level) bind(c, name="pamg_update_overlaps")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int), value :: level
    end function

    ! smoother (transport_tri_semi.F90:543)
    integer(c_int) function pamg_smooth(handle, level, solver, nsweeps) bind(c, name="pamg_smooth")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int), value :: level, solver, nsweeps
    end function

    ! get_residual (:725) + norms
    integer(c_int) function pamg_residual(handle, level, l2, linf) bind(c, name="pamg_residual")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: level
      real(c_double), intent(out) :: l2, linf
    end function

    ! get_convergence (:876)
    integer(c_int) function pamg_convergence(handle, level, conv) bind(c, name="pamg_convergence")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: level
      real(c_double), intent(out) :: conv
    end function

    ! restrictor / prolongator (splitting.F90:10,38)
    integer(c_int) function pamg_restrict(handle, fine_level) bind(c, name="pamg_restrict")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int), value :: fine_level
    end function

    integer(c_int) function pamg_prolong(handle, fine_level) bind(c, name="pamg_prolong")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int), value :: fine_level
    end function

    integer(c_int) function pamg_vcycle_solve(handle, solver, nu1, nu2, ncoarse, max_cycles, tol, cycles, hist) &
        bind(c, name="pamg_vcycle_solve")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: solver, nu1, nu2, ncoarse, max_cycles
      real(c_double), value :: tol
      integer(c_int), intent(out) :: cycles
      real(c_double), intent(out) :: hist(*)
    end function

    ! one itime of the loop at transport_tri_semi.F90:316-379
    integer(c_int) function pamg_literal_timestep(handle, solver, n_multigrid, n_smooth) &
        bind(c, name="pamg_literal_timestep")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int), value :: solver, n_multigrid, n_smooth
    end function

    ! unstr_explicit (transport_tri_unstr.F90:413)
    integer(c_int) function pamg_set_unstructured(handle, E, X, neig, fneig) bind(c, name="pamg_set_unstructured")
      import :: c_int, c_ptr, c_double, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int), value :: E
      real(c_double), intent(in) :: X(2, 3, *)
      integer(c_int32_t), intent(in) :: neig(3, *), fneig(3, *)
    end function

    integer(c_int) function pamg_unstr_upload(handle, tnew) bind(c, name="pamg_unstr_upload")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: tnew(3, *)
    end function

    integer(c_int) function pamg_unstr_download(handle, tnew) bind(c, name="pamg_unstr_download")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: tnew(3, *)
    end function

    integer(c_int) function pamg_explicit_step(handle, dt, u_x, u_y, t_bc, ntime, nits, njac_its, use_exact_minv, &
                                               use_dir) bind(c, name="pamg_explicit_step")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), value :: dt, u_x, u_y, t_bc
      integer(c_int), value :: ntime, nits, njac_its, use_exact_minv, use_dir
    end function

    ! unstr_implicit (transport_tri_unstr.F90:214-387): block-CSR assembly and solve on the device
    integer(c_int) function pamg_implicit_assemble(handle, dt, u_x, u_y, use_dir) bind(c, name="pamg_implicit_assemble")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), value :: dt, u_x, u_y
      integer(c_int), value :: use_dir
    end function

    integer(c_int) function pamg_implicit_get_bsr(handle, val, col) bind(c, name="pamg_implicit_get_bsr")
      import :: c_int, c_ptr, c_double, c_int32_t
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: val(9, 4, *)        ! block entries row-major inside a block
      integer(c_int32_t), intent(out) :: col(4, *)       ! 0-based element of the block column, -1 = none
    end function

    integer(c_int) function pamg_implicit_apply(handle, x, y) bind(c, name="pamg_implicit_apply")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: x(3, *)
      real(c_double), intent(out) :: y(3, *)
    end function

    integer(c_int) function pamg_implicit_step(handle, ntime, nits, tol, max_iters, iters_total, relres) &
        bind(c, name="pamg_implicit_step")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: ntime, nits, max_iters
      real(c_double), value :: tol
      integer(c_int), intent(out) :: iters_total
      real(c_double), intent(out) :: relres
    end function

    ! Petrov-Galerkin stabilisation (transport_tri_unstr.F90:239-267,278)
    integer(c_int) function pamg_unstr_stab(handle, told, dt, u_x, u_y, diff_coe, stab) bind(c, name="pamg_unstr_stab")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(in) :: told(3, *)
      real(c_double), value :: dt, u_x, u_y
      real(c_double), intent(out) :: diff_coe(3, *)      ! (ngi, ele)
      real(c_double), intent(out) :: stab(3, 3, *)       ! (jloc, iloc, ele): row-major blocks
    end function

    integer(c_int) function pamg_implicit_set_stab(handle, with_stab) bind(c, name="pamg_implicit_set_stab")
      import :: c_int, c_ptr
      type(c_ptr), value :: handle
      integer(c_int), value :: with_stab
    end function

    ! get_vtu / get_error (get_vtk_files.F90:10, transport_tri_semi.F90:531)
    integer(c_int) function pamg_output_fields(handle, x_all, analytical, error) bind(c, name="pamg_output_fields")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), intent(out) :: x_all(2, 3, *), analytical(3, *), error(3, *)
    end function

    integer(c_int) function pamg_write_vtu(handle, path, solve_for, binary) bind(c, name="pamg_write_vtu")
      import :: c_int, c_ptr, c_char
      type(c_ptr), value :: handle
      character(kind=c_char), intent(in) :: path(*), solve_for(*)    ! null-terminated
      integer(c_int), value :: binary
    end function

    ! smoother on host arrays, pipelined across calls (call pamg_sync before reading tnew_out)
    integer(c_int) function pamg_smooth_host(handle, solver, nsweeps, tnew_in, tnew_out) bind(c, name="pamg_smooth_host")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      integer(c_int), value :: solver, nsweeps
      real(c_double), intent(in) :: tnew_in(*)
      real(c_double), intent(inout) :: tnew_out(*)
    end function

    ! trans_rec (transport_rect.F90:7)
    integer(c_int) function pamg_trans_rec(handle, CFL, no_ele_row, no_ele_col, x_length, y_length, u_x, u_y, time, nits, &
                                           njac_its, direct_solver, volume_term, x_all, tnew, ntime) bind(c, name="pamg_trans_rec")
      import :: c_int, c_ptr, c_double
      type(c_ptr), value :: handle
      real(c_double), value :: CFL, x_length, y_length, u_x, u_y, time
      integer(c_int), value :: no_ele_row, no_ele_col, nits, njac_its, direct_solver, volume_term
      real(c_double), intent(out) :: x_all(2, 4, *), tnew(4, *)
      integer(c_int), intent(out) :: ntime
    end function

    ! FINDInv (matrices.F90:1618), batched
    integer(c_int) function pamg_apply_local_minv(handle, n, batch, M, rhs, x, Minv, status) &
        bind(c, name="pamg_apply_local_minv")
      import :: c_int, c_ptr, c_double, c_int32_t
      type(c_ptr), value :: handle
      integer(c_int), value :: n, batch
      real(c_double), intent(in) :: M(*)      ! note: row-major [batch][n][n]; pass transpose(matrix) per block
      type(c_ptr), value :: rhs, x, Minv, status
    end function
  end interface

end module pamg_iface
