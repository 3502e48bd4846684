!> Drop-in replacements for the reference's hot-path modules.  Same module names, same public
!! procedures, same argument lists as
!!   common/particle.f90:18,48        common/field.f90:22,66        common/sort.f90:16,36
!!   common/boundary_periodic.f90:26,61,99,251,357,511,571           common/mom_calc.f90:18,48,167
!! so proj/weibel/app.f90 compiles against them unchanged.  Bodies call the C ABI of
!! include/wumingpic2d.h through wm_cabi.  Link order: these objects INSTEAD of the five
!! reference objects in libwuming2d_common.a (INTEGRATION.md).
!!
!! NOT COMPILED HERE (no Fortran compiler / MPI in this image, SURVEY.md F2).
!!
!! Deviations from the reference, all forced by device residency (DESIGN.md "boundary"):
!!  * the procedure arguments of field__fdtd_i (set_boundary_dfield/curre/phi) are accepted and
!!    ignored: the boundary kind is fixed at context creation (the shim's bc module sets it);
!!  * boundary_*__dfield/__curre/__phi stay callable but are no-ops on host arrays (the device
!!    applies them inside wm_field__fdtd_i);
!!  * sort__bucket declares np2 intent(inout): it is where the host copies of up/np2/cumcnt/uf are
!!    refreshed (every WM_SYNC_INTERVAL-th call);
!!  * every procedure that receives the active range (nxs, nxe) forwards it with wm_set_xrange before its own call when
!!    the boundary module is boundary_shock (proj/shock/app.f90 `relocate` moves nxe between two steps, and
!!    particle__solv is the first call of the next step: its cell-centre fields must cover the new columns);
!!  * `gp` is never written on the host (it is device scratch); nothing in the apps reads it
!!    between particle__solv and sort__bucket except the routines replaced here.

module particle
  use wm_cabi
  implicit none
  private
  public :: particle__init, particle__solv
contains

  subroutine particle__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nys_in,nye_in, &
                            delx_in,delt_in,c_in,q_in,r_in)
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nys_in, nye_in
    real(8), intent(in) :: delx_in, delt_in, c_in, q_in(nsp_in), r_in(nsp_in)
    cfg%ndim = ndim_in; cfg%np = np_in; cfg%nsp = nsp_in
    cfg%nxgs = nxgs_in; cfg%nxge = nxge_in; cfg%nygs = nygs_in; cfg%nyge = nyge_in
    cfg%nys = nys_in; cfg%nye = nye_in
    cfg%delx = delx_in; cfg%delt = delt_in; cfg%c = c_in
    cfg%q(1:nsp_in) = q_in; cfg%r(1:nsp_in) = r_in
    have_grid = .true.
    call wm_shim__try_create()
  end subroutine particle__init

  subroutine particle__solv(gp,up,uf,cumcnt,nxs,nxe)
    integer, intent(in)  :: nxs, nxe
    integer, intent(in)  :: cumcnt(cfg%nxgs:cfg%nxge+1,cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(in)  :: up(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(in)  :: uf(6,cfg%nxgs-2:cfg%nxge+2,cfg%nys-2:cfg%nye+2)
    real(8), intent(out) :: gp(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    if (.not. c_associated(ctx)) then
       write(6,*) 'Initialize first by calling particle__init()'
       stop
    end if
    call wm_shim__set_xrange(nxs, nxe)   ! first call of a step: relocate() may have moved nxe since the last one
    call wm_shim__upload_if_dirty(up, uf, cumcnt)
    call wm_check(wm_particle__solv(ctx), 'particle__solv')
  end subroutine particle__solv

end module particle


module field
  use wm_cabi
  implicit none
  private
  public :: field__init, field__fdtd_i
contains

  subroutine field__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nys_in,nye_in, &
                         mnpr_in,ncomw_in,opsum_in,nerr_in,                                  &
                         delx_in,delt_in,c_in,q_in,r_in,gfac_in)
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nys_in, nye_in
    integer, intent(in) :: mnpr_in, ncomw_in, opsum_in, nerr_in
    real(8), intent(in) :: delx_in, delt_in, c_in, q_in(nsp_in), r_in(nsp_in), gfac_in
    cfg%gfac = gfac_in
    comm_world = ncomw_in
    have_phys = .true.
    call wm_shim__try_create()
  end subroutine field__init

  subroutine field__fdtd_i(uf,up,gp,cumcnt,nxs,nxe, &
       & set_boundary_dfield, set_boundary_curre, set_boundary_phi)
    external :: set_boundary_dfield, set_boundary_curre, set_boundary_phi   ! ignored, see header
    integer, intent(in)    :: nxs, nxe
    integer, intent(in)    :: cumcnt(cfg%nxgs:cfg%nxge+1,cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(in)    :: gp(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(in)    :: up(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(inout) :: uf(6,cfg%nxgs-2:cfg%nxge+2,cfg%nys-2:cfg%nye+2)
    if (.not. c_associated(ctx)) then
       write(6,*) 'Initialize first by calling field__init()'
       stop
    end if
    call wm_shim__set_xrange(nxs, nxe)
    ! ele_cur + bc__curre + cgm x3 + bc__dfield x2 + uf update, all on the device
    call wm_check(wm_field__fdtd_i(ctx), 'field__fdtd_i')
    ! uf is final for this step: refresh the host copy if this step's sort__bucket will sync
    if (sync_interval > 0 .and. nstep_since_sync + 1 >= sync_interval) &
         call wm_check(wm_download_field(ctx, uf), 'field__fdtd_i(download uf)')
  end subroutine field__fdtd_i

end module field


module sort
  use wm_cabi
  implicit none
  private
  public :: sort__init, sort__bucket
contains

  subroutine sort__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nys_in,nye_in)
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nys_in, nye_in
  end subroutine sort__init

  !> first argument is the OUTPUT (the app calls sort__bucket(up, gp, ...), common/sort.f90:36)
  subroutine sort__bucket(gp,up,cumcnt,np2,nxs,nxe)
    integer, intent(in)    :: nxs, nxe
    integer, intent(inout) :: np2(cfg%nys:cfg%nye,cfg%nsp)
    integer, intent(out)   :: cumcnt(cfg%nxgs:cfg%nxge+1,cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(in)    :: up(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(out)   :: gp(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    if (host_dirty) then
       ! stand-alone use (restart path, proj/weibel/app.f90:351-352): sort host rows on the device
       call wm_check(wm_upload_particles(ctx, up, np2), 'sort__bucket(upload)')
       call wm_check(wm_download_particles(ctx, gp, np2, cumcnt), 'sort__bucket(download)')
       return                      ! uf is still only on the host: stay dirty until particle__solv
    end if
    call wm_shim__set_xrange(nxs, nxe)
    call wm_check(wm_sort__bucket(ctx), 'sort__bucket')
    nstep_since_sync = nstep_since_sync + 1
    if (sync_interval > 0 .and. nstep_since_sync >= sync_interval) then
       call wm_check(wm_download_particles(ctx, gp, np2, cumcnt), 'sort__bucket(download)')
       nstep_since_sync = 0          ! uf was refreshed by field__fdtd_i of this same step
    end if
  end subroutine sort__bucket

end module sort


module boundary_periodic
  use wm_cabi
  implicit none
  private
  public :: boundary_periodic__init
  public :: boundary_periodic__dfield, boundary_periodic__particle_x, boundary_periodic__particle_y
  public :: boundary_periodic__curre, boundary_periodic__phi, boundary_periodic__mom
contains

  subroutine boundary_periodic__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nys_in,nye_in, &
                            nup_in,ndown_in,mnpi_in,mnpr_in,ncomw_in,nerr_in,nstat_in,          &
                            delx_in,delt_in,c_in)
    use mpi
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nys_in, nye_in
    integer, intent(in) :: nup_in, ndown_in, mnpi_in, mnpr_in, ncomw_in, nerr_in, nstat_in(:)
    real(8), intent(in) :: delx_in, delt_in, c_in
    integer :: nerr, nrank, nsize
    call MPI_COMM_RANK(ncomw_in, nrank, nerr)
    call MPI_COMM_SIZE(ncomw_in, nsize, nerr)
    cfg%nrank = nrank; cfg%nsize = nsize      ! ring: nup = nrank+1, ndown = nrank-1 (mpi_set.f90:44-47)
    comm_world = ncomw_in
    bc_kind = WM_BC_PERIODIC
    have_ring = .true.
    call wm_shim__try_create()
  end subroutine boundary_periodic__init

  subroutine boundary_periodic__particle_x(up,np2)
    integer, intent(in)    :: np2(cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(inout) :: up(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    call wm_check(wm_boundary__particle_x(ctx), 'boundary_periodic__particle_x')
  end subroutine boundary_periodic__particle_x

  subroutine boundary_periodic__particle_y(up,np2)
    integer, intent(inout) :: np2(cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(inout) :: up(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    call wm_check(wm_boundary__particle_y(ctx), 'boundary_periodic__particle_y')
  end subroutine boundary_periodic__particle_y

  ! the three field boundary procedures are applied on the device inside wm_field__fdtd_i
  subroutine boundary_periodic__dfield(df,nxs,nxe,nys,nye,nxgs,nxge)
    integer, intent(in)    :: nxs, nxe, nys, nye, nxgs, nxge
    real(8), intent(inout) :: df(6,nxgs-2:nxge+2,nys-2:nye+2)
  end subroutine boundary_periodic__dfield

  subroutine boundary_periodic__curre(uj,nxs,nxe,nys,nye,nxgs,nxge)
    integer, intent(in)    :: nxs, nxe, nys, nye, nxgs, nxge
    real(8), intent(inout) :: uj(3,nxgs-2:nxge+2,nys-2:nye+2)
  end subroutine boundary_periodic__curre

  subroutine boundary_periodic__phi(phi,nxs,nxe,nys,nye,l)
    integer, intent(in)    :: nxs, nxe, nys, nye, l
    real(8), intent(inout) :: phi(nxs-1:nxe+1,nys-1:nye+1)
  end subroutine boundary_periodic__phi

  subroutine boundary_periodic__mom(mom)
    real(8), intent(inout) :: mom(7,cfg%nxgs-1:cfg%nxge+1,cfg%nys-1:cfg%nye+1,cfg%nsp)
    call wm_check(wm_boundary__mom(ctx, mom), 'boundary_periodic__mom')
  end subroutine boundary_periodic__mom

end module boundary_periodic


module mom_calc
  use wm_cabi
  implicit none
  private
  public :: mom_calc__init, mom_calc__accl, mom_calc__nvt
contains

  subroutine mom_calc__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nys_in,nye_in, &
                            delx_in,delt_in,c_in,q_in,r_in)
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nys_in, nye_in
    real(8), intent(in) :: delx_in, delt_in, c_in, q_in(nsp_in), r_in(nsp_in)
  end subroutine mom_calc__init

  !> half-step acceleration into the device's idle particle store (common/mom_calc.f90:48)
  subroutine mom_calc__accl(gp,up,uf,cumcnt,nxs,nxe)
    integer, intent(in)  :: nxs, nxe
    integer, intent(in)  :: cumcnt(cfg%nxgs:cfg%nxge+1,cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(in)  :: up(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(in)  :: uf(6,cfg%nxgs-2:cfg%nxge+2,cfg%nys-2:cfg%nye+2)
    real(8), intent(out) :: gp(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    call wm_shim__set_xrange(nxs, nxe)
    call wm_shim__upload_if_dirty(up, uf, cumcnt)
    call wm_check(wm_mom_calc__accl(ctx), 'mom_calc__accl')
  end subroutine mom_calc__accl

  !> seven moments with linear weights -> host mom, ghosts unfolded (common/mom_calc.f90:167)
  subroutine mom_calc__nvt(mom,up,np2)
    integer, intent(in)  :: np2(cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(in)  :: up(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(out) :: mom(7,cfg%nxgs-1:cfg%nxge+1,cfg%nys-1:cfg%nye+1,cfg%nsp)
    call wm_check(wm_mom_calc__nvt(ctx, mom), 'mom_calc__nvt')
  end subroutine mom_calc__nvt

end module mom_calc


!> proj/reconnection/boundary_reconnection.f90 (reflecting / conducting x walls, periodic y).
!! The app selects it by `use boundary_reconnection, bc__init => boundary_reconnection__init, ...`
!! (proj/reconnection/app.f90:6-13); nothing else changes.
module boundary_reconnection
  use wm_cabi
  implicit none
  private
  public :: boundary_reconnection__init
  public :: boundary_reconnection__dfield, boundary_reconnection__particle_x, boundary_reconnection__particle_y
  public :: boundary_reconnection__curre, boundary_reconnection__phi, boundary_reconnection__mom
contains

  subroutine boundary_reconnection__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nys_in,nye_in, &
                            nup_in,ndown_in,mnpi_in,mnpr_in,ncomw_in,nerr_in,nstat_in,          &
                            delx_in,delt_in,c_in)
    use mpi
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nys_in, nye_in
    integer, intent(in) :: nup_in, ndown_in, mnpi_in, mnpr_in, ncomw_in, nerr_in, nstat_in(:)
    real(8), intent(in) :: delx_in, delt_in, c_in
    integer :: nerr, nrank, nsize
    call MPI_COMM_RANK(ncomw_in, nrank, nerr)
    call MPI_COMM_SIZE(ncomw_in, nsize, nerr)
    cfg%nrank = nrank; cfg%nsize = nsize
    comm_world = ncomw_in
    bc_kind = WM_BC_RECONNECTION
    have_ring = .true.
    call wm_shim__try_create()
  end subroutine boundary_reconnection__init

  !> the device library takes the wall positions from nxgs/nxge; the apps pass nxs = nxgs, nxe = nxge
  subroutine boundary_reconnection__particle_x(up,np2,nxs,nxe)
    integer, intent(in)    :: nxs, nxe
    integer, intent(in)    :: np2(cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(inout) :: up(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    if (nxs /= cfg%nxgs .or. nxe /= cfg%nxge) then
       write(6,*) 'boundary_reconnection__particle_x: nxs/nxe must equal nxgs/nxge'
       stop
    end if
    call wm_check(wm_boundary__particle_x(ctx), 'boundary_reconnection__particle_x')
  end subroutine boundary_reconnection__particle_x

  subroutine boundary_reconnection__particle_y(up,np2)
    integer, intent(inout) :: np2(cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(inout) :: up(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    call wm_check(wm_boundary__particle_y(ctx), 'boundary_reconnection__particle_y')
  end subroutine boundary_reconnection__particle_y

  subroutine boundary_reconnection__dfield(df,nxs,nxe,nys,nye,nxgs,nxge)
    integer, intent(in)    :: nxs, nxe, nys, nye, nxgs, nxge
    real(8), intent(inout) :: df(6,nxgs-2:nxge+2,nys-2:nye+2)
  end subroutine boundary_reconnection__dfield

  subroutine boundary_reconnection__curre(uj,nxs,nxe,nys,nye,nxgs,nxge)
    integer, intent(in)    :: nxs, nxe, nys, nye, nxgs, nxge
    real(8), intent(inout) :: uj(3,nxgs-2:nxge+2,nys-2:nye+2)
  end subroutine boundary_reconnection__curre

  subroutine boundary_reconnection__phi(phi,nxs,nxe,nys,nye,l)
    integer, intent(in)    :: nxs, nxe, nys, nye, l
    real(8), intent(inout) :: phi(nxs-1:nxe+1,nys-1:nye+1)
  end subroutine boundary_reconnection__phi

  subroutine boundary_reconnection__mom(mom)
    real(8), intent(inout) :: mom(7,cfg%nxgs-1:cfg%nxge+1,cfg%nys-1:cfg%nye+1,cfg%nsp)
    call wm_check(wm_boundary__mom(ctx, mom), 'boundary_reconnection__mom')
  end subroutine boundary_reconnection__mom

end module boundary_reconnection


!> proj/shock/boundary_shock.f90 (reflecting wall on the left, injection wall on the right, periodic y).
!! The app selects it by `use boundary_shock, bc__init => boundary_shock__init, bc__injection => ...`
!! (proj/shock/app.f90:6-14).  The active range nxs..nxe of the step is forwarded with wm_set_xrange.  The driver's
!! `inject` and `relocate` (proj/shock/app.f90:611-850) write new particles and upstream fields straight into the
!! host arrays: with device-resident state they have to hand the new records to wm_append_particles and the
!! changed field columns to wm_upload_field instead (INTEGRATION.md).
module boundary_shock
  use wm_cabi
  implicit none
  private
  public :: boundary_shock__init
  public :: boundary_shock__dfield, boundary_shock__particle_x, boundary_shock__particle_y
  public :: boundary_shock__injection
  public :: boundary_shock__curre, boundary_shock__phi, boundary_shock__mom
contains

  subroutine boundary_shock__init(ndim_in,np_in,nsp_in,nxgs_in,nxge_in,nygs_in,nyge_in,nys_in,nye_in, &
                            nup_in,ndown_in,mnpi_in,mnpr_in,ncomw_in,nerr_in,nstat_in,          &
                            delx_in,delt_in,c_in)
    use mpi
    integer, intent(in) :: ndim_in, np_in, nsp_in
    integer, intent(in) :: nxgs_in, nxge_in, nygs_in, nyge_in, nys_in, nye_in
    integer, intent(in) :: nup_in, ndown_in, mnpi_in, mnpr_in, ncomw_in, nerr_in, nstat_in(:)
    real(8), intent(in) :: delx_in, delt_in, c_in
    integer :: nerr, nrank, nsize
    call MPI_COMM_RANK(ncomw_in, nrank, nerr)
    call MPI_COMM_SIZE(ncomw_in, nsize, nerr)
    cfg%nrank = nrank; cfg%nsize = nsize
    comm_world = ncomw_in
    bc_kind = WM_BC_SHOCK
    have_ring = .true.
    call wm_shim__try_create()
  end subroutine boundary_shock__init

  subroutine boundary_shock__particle_x(up,np2,nxs,nxe)
    integer, intent(in)    :: nxs, nxe
    integer, intent(in)    :: np2(cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(inout) :: up(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    if (nxs /= cfg%nxgs .or. nxe /= cfg%nxge) then
       write(6,*) 'boundary_shock__particle_x: nxs/nxe must equal nxgs/nxge'
       stop
    end if
    call wm_check(wm_boundary__particle_x(ctx), 'boundary_shock__particle_x')
  end subroutine boundary_shock__particle_x

  subroutine boundary_shock__particle_y(up,np2)
    integer, intent(inout) :: np2(cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(inout) :: up(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    call wm_check(wm_boundary__particle_y(ctx), 'boundary_shock__particle_y')
  end subroutine boundary_shock__particle_y

  !> bc__injection(gp,np2,nxs,nxe,u0), called between particle__solv and field__fdtd_i (proj/shock/app.f90:112-113)
  subroutine boundary_shock__injection(up,np2,nxs,nxe,u0)
    integer, intent(in)    :: nxs, nxe
    integer, intent(in)    :: np2(cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(inout) :: up(cfg%ndim,cfg%np,cfg%nys:cfg%nye,cfg%nsp)
    real(8), intent(in)    :: u0
    ! the shock app moves nxe (relocate): tell the library the active range of this step's calls
    call wm_shim__set_xrange(nxs, nxe)
    call wm_check(wm_boundary__injection(ctx, u0), 'boundary_shock__injection')
  end subroutine boundary_shock__injection

  subroutine boundary_shock__dfield(df,nxs,nxe,nys,nye,nxgs,nxge)
    integer, intent(in)    :: nxs, nxe, nys, nye, nxgs, nxge
    real(8), intent(inout) :: df(6,nxgs-2:nxge+2,nys-2:nye+2)
  end subroutine boundary_shock__dfield

  subroutine boundary_shock__curre(uj,nxs,nxe,nys,nye,nxgs,nxge)
    integer, intent(in)    :: nxs, nxe, nys, nye, nxgs, nxge
    real(8), intent(inout) :: uj(3,nxgs-2:nxge+2,nys-2:nye+2)
  end subroutine boundary_shock__curre

  subroutine boundary_shock__phi(phi,nxs,nxe,nys,nye,l)
    integer, intent(in)    :: nxs, nxe, nys, nye, l
    real(8), intent(inout) :: phi(nxs-1:nxe+1,nys-1:nye+1)
  end subroutine boundary_shock__phi

  subroutine boundary_shock__mom(mom)
    real(8), intent(inout) :: mom(7,cfg%nxgs-1:cfg%nxge+1,cfg%nys-1:cfg%nye+1,cfg%nsp)
    call wm_check(wm_boundary__mom(ctx, mom), 'boundary_shock__mom')
  end subroutine boundary_shock__mom

end module boundary_shock
