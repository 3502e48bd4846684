!> wm_cabi -- ISO_C_BINDING view of include/wumingpic2d.h plus the state the shim modules share.
!!
!! The reference's hot-path modules (common/particle.f90, field.f90, sort.f90,
!! boundary_periodic.f90, mom_calc.f90) are replaced by the modules of the same name in this
!! directory; their bodies call the C ABI declared here.  The application (proj/*/app.f90) and
!! everything else stay unchanged.
!!
!! NOT COMPILED IN THIS REPOSITORY'S ENVIRONMENT: the image has no Fortran compiler and no MPI
!! (SURVEY.md F2).  The C ABI these interfaces bind is exercised by tests/ through ctypes with the
!! same Fortran array layouts.  See INTEGRATION.md for the build recipe.
!!
!! Residency (SURVEY.md section 7 "Residency vs drop-in"): the device state is authoritative between
!! calls.  Host arrays are uploaded when `host_dirty` is set (first step, after a restart load or
!! any host-side edit: call wm_shim__mark_host_dirty) and downloaded at the end of sort__bucket
!! every `sync_interval` steps (environment WM_SYNC_INTERVAL; default 1 = the host arrays are
!! valid after every step, exactly like the reference; set it to the gcd of intvl_ptcl, intvl_mom
!! and intvl_orb to keep the state resident between outputs; 0 = never, call wm_shim__download).
module wm_cabi
  use, intrinsic :: iso_c_binding
  implicit none
  public

  integer(c_int), parameter :: WM_NSP_MAX = 2
  integer(c_int), parameter :: WM_BC_PERIODIC = 0, WM_BC_RECONNECTION = 1, WM_BC_SHOCK = 2
  integer(c_int), parameter :: WM_FLAG_EXACT_PUSH = 1

  !> mirrors `struct wm_config` of include/wumingpic2d.h member by member
  type, bind(C) :: wm_config
     integer(c_int32_t) :: ndim, np, nsp
     integer(c_int32_t) :: nxgs, nxge, nygs, nyge
     integer(c_int32_t) :: nys, nye
     integer(c_int32_t) :: nrank, nsize
     integer(c_int32_t) :: bc
     integer(c_int32_t) :: device
     integer(c_int32_t) :: flags
     real(c_double)     :: delx, delt, c, gfac
     real(c_double)     :: q(WM_NSP_MAX), r(WM_NSP_MAX)
     integer(c_int64_t) :: capacity
  end type wm_config

  interface
     function wm_last_error() bind(C, name='wm_last_error') result(p)
       import :: c_ptr
       type(c_ptr) :: p
     end function wm_last_error
     function wm_create(cfg, ctx) bind(C, name='wm_create') result(ierr)
       import :: c_int, c_ptr, wm_config
       type(wm_config), intent(in) :: cfg
       type(c_ptr), intent(out)    :: ctx
       integer(c_int) :: ierr
     end function wm_create
     function wm_destroy(ctx) bind(C, name='wm_destroy') result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wm_destroy
     function wm_comm_unique_id(id128) bind(C, name='wm_comm_unique_id') result(ierr)
       import :: c_int, c_char
       character(kind=c_char), intent(out) :: id128(128)
       integer(c_int) :: ierr
     end function wm_comm_unique_id
     function wm_comm_init(ctx, id128) bind(C, name='wm_comm_init') result(ierr)
       import :: c_int, c_ptr, c_char
       type(c_ptr), value :: ctx
       character(kind=c_char), intent(in) :: id128(128)
       integer(c_int) :: ierr
     end function wm_comm_init
     function wm_host_register(ptr, bytes) bind(C, name='wm_host_register') result(ierr)
       import :: c_int, c_ptr, c_size_t
       type(c_ptr), value :: ptr
       integer(c_size_t), value :: bytes
       integer(c_int) :: ierr
     end function wm_host_register
     function wm_upload_particles_sorted(ctx, up, np2, cumcnt) bind(C, name='wm_upload_particles_sorted') result(ierr)
       import :: c_int, c_ptr, c_double, c_int32_t
       type(c_ptr), value :: ctx
       real(c_double), intent(in)     :: up(*)
       integer(c_int32_t), intent(in) :: np2(*), cumcnt(*)
       integer(c_int) :: ierr
     end function wm_upload_particles_sorted
     function wm_upload_particles(ctx, up, np2) bind(C, name='wm_upload_particles') result(ierr)
       import :: c_int, c_ptr, c_double, c_int32_t
       type(c_ptr), value :: ctx
       real(c_double), intent(in)     :: up(*)
       integer(c_int32_t), intent(in) :: np2(*)
       integer(c_int) :: ierr
     end function wm_upload_particles
     function wm_upload_field(ctx, uf) bind(C, name='wm_upload_field') result(ierr)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(in) :: uf(*)
       integer(c_int) :: ierr
     end function wm_upload_field
     function wm_download_particles(ctx, up, np2, cumcnt) bind(C, name='wm_download_particles') result(ierr)
       import :: c_int, c_ptr, c_double, c_int32_t
       type(c_ptr), value :: ctx
       real(c_double), intent(inout)     :: up(*)
       integer(c_int32_t), intent(inout) :: np2(*), cumcnt(*)
       integer(c_int) :: ierr
     end function wm_download_particles
     function wm_download_gp(ctx, gp) bind(C, name='wm_download_gp') result(ierr)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(inout) :: gp(*)
       integer(c_int) :: ierr
     end function wm_download_gp
     function wm_download_field(ctx, uf) bind(C, name='wm_download_field') result(ierr)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(inout) :: uf(*)
       integer(c_int) :: ierr
     end function wm_download_field
     function wm_particle__solv(ctx) bind(C, name='wm_particle__solv') result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wm_particle__solv
     function wm_field__fdtd_i(ctx) bind(C, name='wm_field__fdtd_i') result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wm_field__fdtd_i
     function wm_boundary__particle_x(ctx) bind(C, name='wm_boundary__particle_x') result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wm_boundary__particle_x
     function wm_boundary__injection(ctx, u0) bind(C, name='wm_boundary__injection') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value    :: ctx
       real(c_double), value :: u0
       integer(c_int)        :: ierr
     end function wm_boundary__injection
     function wm_set_xrange(ctx, nxs, nxe) bind(C, name='wm_set_xrange') result(ierr)
       import :: c_ptr, c_int, c_int32_t
       type(c_ptr), value       :: ctx
       integer(c_int32_t), value :: nxs, nxe
       integer(c_int)           :: ierr
     end function wm_set_xrange
     function wm_append_particles(ctx, isp, n, rec) bind(C, name='wm_append_particles') result(ierr)
       import :: c_ptr, c_int, c_int32_t, c_int64_t, c_double
       type(c_ptr), value        :: ctx
       integer(c_int32_t), value :: isp
       integer(c_int64_t), value :: n
       real(c_double), intent(in) :: rec(*)
       integer(c_int)            :: ierr
     end function wm_append_particles
     function wm_set_u_inject(ctx, u0) bind(C, name='wm_set_u_inject') result(ierr)
       import :: c_ptr, c_int, c_double
       type(c_ptr), value    :: ctx
       real(c_double), value :: u0
       integer(c_int)        :: ierr
     end function wm_set_u_inject
     function wm_boundary__particle_y(ctx) bind(C, name='wm_boundary__particle_y') result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wm_boundary__particle_y
     function wm_sort__bucket(ctx) bind(C, name='wm_sort__bucket') result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wm_sort__bucket
     function wm_step(ctx, nsteps) bind(C, name='wm_step') result(ierr)
       import :: c_int, c_ptr, c_int32_t
       type(c_ptr), value :: ctx
       integer(c_int32_t), value :: nsteps
       integer(c_int) :: ierr
     end function wm_step
     function wm_mom_calc__accl(ctx) bind(C, name='wm_mom_calc__accl') result(ierr)
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: ierr
     end function wm_mom_calc__accl
     function wm_mom_calc__nvt(ctx, mom) bind(C, name='wm_mom_calc__nvt') result(ierr)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(inout) :: mom(*)
       integer(c_int) :: ierr
     end function wm_mom_calc__nvt
     function wm_boundary__mom(ctx, mom) bind(C, name='wm_boundary__mom') result(ierr)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(inout) :: mom(*)
       integer(c_int) :: ierr
     end function wm_boundary__mom
     function wm_energy(ctx, out) bind(C, name='wm_energy') result(ierr)
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double), intent(out) :: out(*)
       integer(c_int) :: ierr
     end function wm_energy
     function c_strlen(s) bind(C, name='strlen') result(n)
       import :: c_ptr, c_size_t
       type(c_ptr), value :: s
       integer(c_size_t) :: n
     end function c_strlen
  end interface

  ! ---- state shared by the shim modules (one rank = one GPU = one context) -------------------
  type(c_ptr), save :: ctx = c_null_ptr
  type(wm_config), save :: cfg
  logical, save :: have_grid = .false., have_phys = .false., have_ring = .false.
  logical, save :: host_dirty = .true.      !< host arrays are newer than the device state
  integer, save :: sync_interval = 1        !< download every n-th sort__bucket (0 = never)
  integer, save :: nstep_since_sync = 0
  integer, save :: bc_kind = WM_BC_PERIODIC
  integer, save :: comm_world = -1          !< ncomw, used only to broadcast the NCCL id

contains

  !> the reference reports errors with `write + stop` (e.g. common/particle.f90:63-66)
  subroutine wm_check(ierr, who)
    integer(c_int), intent(in)   :: ierr
    character(len=*), intent(in) :: who
    character(kind=c_char), pointer :: msg(:)
    type(c_ptr) :: p
    integer :: n, k
    character(len=512) :: buf
    if (ierr == 0) return
    p = wm_last_error()
    n = int(c_strlen(p))
    call c_f_pointer(p, msg, [n])
    buf = ' '
    do k = 1, min(n, len(buf))
       buf(k:k) = msg(k)
    end do
    write(6,*) trim(who), ': ', trim(buf)
    stop
  end subroutine wm_check

  !> called by every *__init of the shim; creates the context once all scalars are known
  !! (grid from particle__init / sort__init, gfac from field__init, ring from boundary_*__init)
  subroutine wm_shim__try_create()
    use mpi
    character(kind=c_char) :: id(128)
    character(len=32) :: env
    integer :: nerr, stat, ndev_rank
    if (c_associated(ctx)) return
    if (.not.(have_grid .and. have_phys .and. have_ring)) return
    cfg%bc = bc_kind
    cfg%flags = 0
    cfg%capacity = 0
    call get_environment_variable('WM_EXACT_PUSH', env, status=stat)
    if (stat == 0 .and. trim(env) == '1') cfg%flags = WM_FLAG_EXACT_PUSH
    call get_environment_variable('WM_SYNC_INTERVAL', env, status=stat)
    if (stat == 0) read(env, *) sync_interval
    ! one rank per GPU of the node (8 per B200 box); -1 would keep the current device
    call get_environment_variable('WM_RANKS_PER_NODE', env, status=stat)
    ndev_rank = 8
    if (stat == 0) read(env, *) ndev_rank
    cfg%device = mod(cfg%nrank, ndev_rank)
    call wm_check(wm_create(cfg, ctx), 'wm_create')
    if (cfg%nsize > 1) then
       ! MPI stays the control plane of the unchanged driver: it only carries the 128-byte NCCL id
       if (cfg%nrank == 0) call wm_check(wm_comm_unique_id(id), 'wm_comm_unique_id')
       call MPI_BCAST(id, 128, MPI_CHARACTER, 0, comm_world, nerr)
       call wm_check(wm_comm_init(ctx, id), 'wm_comm_init')
    end if
  end subroutine wm_shim__try_create

  !> forward the active x range of a call (its nxs, nxe arguments) to the library.  Only boundary_shock works on a
  !! sub-range and moves nxe in time (proj/shock/app.f90:611-621); the other modules always pass nxgs, nxge.
  subroutine wm_shim__set_xrange(nxs, nxe)
    integer, intent(in) :: nxs, nxe
    if (bc_kind /= WM_BC_SHOCK) return
    call wm_check(wm_set_xrange(ctx, int(nxs, c_int32_t), int(nxe, c_int32_t)), 'wm_set_xrange')
  end subroutine wm_shim__set_xrange

  !> tell the shim that the application changed up/uf/np2/cumcnt on the host (restart load,
  !! shock `inject`/`relocate` when not using the device-side injection)
  subroutine wm_shim__mark_host_dirty()
    host_dirty = .true.
  end subroutine wm_shim__mark_host_dirty

  !> optional, once after the allocation of the particle arrays (proj/weibel/app.f90:74-82): page-lock them, so that the copies
  !! of the sync steps run asynchronously at the full PCIe rate (wm_host_register; pageable memory is staged at about half of it)
  subroutine wm_shim__pin_host(up, gp)
    real(c_double), intent(in), target :: up(*), gp(*)
    integer(c_size_t) :: bytes
    bytes = 8_c_size_t * int(cfg%ndim, c_size_t) * int(cfg%np, c_size_t) &
          * int(cfg%nye - cfg%nys + 1, c_size_t) * int(cfg%nsp, c_size_t)
    call wm_check(wm_host_register(c_loc(up), bytes), 'wm_host_register(up)')
    call wm_check(wm_host_register(c_loc(gp), bytes), 'wm_host_register(gp)')
  end subroutine wm_shim__pin_host

  !> host -> device, if needed.  np2(j,isp) = cumcnt(nxge+1,j,isp) by construction (common/sort.f90:64-69)
  subroutine wm_shim__upload_if_dirty(up, uf, cumcnt)
    real(c_double), intent(in)     :: up(*), uf(*)
    integer(c_int32_t), intent(in) :: cumcnt(cfg%nxgs:cfg%nxge+1, cfg%nys:cfg%nye, cfg%nsp)
    integer(c_int32_t) :: np2(cfg%nys:cfg%nye, cfg%nsp)
    if (.not. host_dirty) return
    np2(:,:) = cumcnt(cfg%nxge+1, :, :)
    call wm_check(wm_upload_particles_sorted(ctx, up, np2, cumcnt), 'wm_upload_particles_sorted')
    call wm_check(wm_upload_field(ctx, uf), 'wm_upload_field')
    host_dirty = .false.
  end subroutine wm_shim__upload_if_dirty

  !> device -> host on demand (before io__ptcl / io__mom / save_restart when WM_SYNC_INTERVAL /= 1)
  subroutine wm_shim__download(up, uf, np2, cumcnt)
    real(c_double), intent(inout)     :: up(*), uf(*)
    integer(c_int32_t), intent(inout) :: np2(*), cumcnt(*)
    call wm_check(wm_download_particles(ctx, up, np2, cumcnt), 'wm_download_particles')
    call wm_check(wm_download_field(ctx, uf), 'wm_download_field')
    nstep_since_sync = 0
  end subroutine wm_shim__download

end module wm_cabi
