! critic2_gpu.f90 -- ISO_C_BINDING shim between critic2's Fortran host code and the CUDA library
! libcritic2_gpu.so (C ABI: include/critic2_gpu.h).  It is meant to sit next to
! src/c_interface_module.f90 / src/libcritic2.f90 in the critic2 tree and follows their conventions
! (bind(c) interfaces, type(c_ptr) opaque handles, integer status codes turned into ferror calls).
!
! NOT COMPILED IN THIS REPOSITORY'S IMAGE (no Fortran compiler is available); the same entry points are
! exercised from Python/ctypes in tests/ and from C++ in critic2_b200/csrc/host.  See INTEGRATION.md for
! the four patch points in bader@proc.f90, yt@proc.f90, integration@proc.f90 and nci@proc.f90.
module critic2_gpu
  use iso_c_binding
  implicit none
  private

  public :: gpu_enabled, gpu_init, gpu_end
  public :: gpu_bader_integrate, gpu_yt_integrate, gpu_yt_isosurface, gpu_integrate_fields, gpu_integrate_multipoles
  public :: gpu_nci_rdg, gpu_nci_rdg_fourier, gpu_grid_fft, gpu_read_text_block, gpu_write_text_block, gpu_wcube_block, gpu_basins_remap, gpu_hirshfeld_fields, gpu_yt_export, gpu_voronoi_grid

  logical :: gpu_enabled = .false.        !< set by gpu_init (environment variable CRITIC2_GPU=1)
  type(c_ptr) :: ctx = c_null_ptr         !< c2g_context
  type(c_ptr) :: basins = c_null_ptr      !< c2g_basins of the last BADER/YT call (consumed by gpu_integrate_fields)
  integer(c_int) :: hgrid = -1            !< resident copy of bas%f

  interface
     function c2g_init(device,ctx) bind(c,name="c2g_init")
       import :: c_int, c_ptr
       integer(c_int), value :: device
       type(c_ptr) :: ctx
       integer(c_int) :: c2g_init
     end function c2g_init
     !> one process, ngpus devices (critic2 is a single process): z-slabs and NCCL live inside the library
     function c2g_init_devices(ngpus,ctx) bind(c,name="c2g_init_devices")
       import :: c_int, c_ptr
       integer(c_int), value :: ngpus
       type(c_ptr) :: ctx
       integer(c_int) :: c2g_init_devices
     end function c2g_init_devices
     !> the ytdata record (yt.f90:36-45) for hosts that keep their own yt_weights; inear/fnear may be c_null_ptr
     function c2g_yt_export(res,nlo,ibasin,iio,inear,fnear) bind(c,name="c2g_yt_export")
       import :: c_int, c_ptr
       type(c_ptr), value :: res
       integer(c_int) :: nlo(*), ibasin(*), iio(*)
       type(c_ptr), value :: inear, fnear
       integer(c_int) :: c2g_yt_export
     end function c2g_yt_export
     function c2g_voronoi_grid(ctx,n,x2c,nat,xat,res) bind(c,name="c2g_voronoi_grid")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int) :: n(3)
       real(c_double) :: x2c(3,3), xat(3,*)
       integer(c_int), value :: nat
       type(c_ptr) :: res
       integer(c_int) :: c2g_voronoi_grid
     end function c2g_voronoi_grid
     subroutine c2g_finalize(ctx) bind(c,name="c2g_finalize")
       import :: c_ptr
       type(c_ptr), value :: ctx
     end subroutine c2g_finalize
     function c2g_last_error(ctx) bind(c,name="c2g_last_error")
       import :: c_ptr
       type(c_ptr), value :: ctx
       type(c_ptr) :: c2g_last_error
     end function c2g_last_error
     function c2g_grid_upload(ctx,f,n,handle) bind(c,name="c2g_grid_upload")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double) :: f(*)
       integer(c_int) :: n(3), handle
       integer(c_int) :: c2g_grid_upload
     end function c2g_grid_upload
     function c2g_grid_free(ctx,handle) bind(c,name="c2g_grid_free")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int), value :: handle
       integer(c_int) :: c2g_grid_free
     end function c2g_grid_free
     function c2g_bader_assign(ctx,handle,car2lat,lat_i_dist,algo,order,nmax,res) bind(c,name="c2g_bader_assign")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: handle, algo, order
       real(c_double) :: car2lat(3,3), lat_i_dist(27)
       integer(c_int) :: nmax
       type(c_ptr) :: res
       integer(c_int) :: c2g_bader_assign
     end function c2g_bader_assign
     function c2g_yt_build(ctx,handle,nvec,vec,area,nmax,res) bind(c,name="c2g_yt_build")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: handle, nvec
       integer(c_int) :: vec(3,*)
       real(c_double) :: area(*)
       integer(c_int) :: nmax
       type(c_ptr) :: res
       integer(c_int) :: c2g_yt_build
     end function c2g_yt_build
     function c2g_yt_isosurface(yt,isov,nraw,nattr,res) bind(c,name="c2g_yt_isosurface")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: yt
       real(c_double), value :: isov
       integer(c_int) :: nraw, nattr
       type(c_ptr) :: res
       integer(c_int) :: c2g_yt_isosurface
     end function c2g_yt_isosurface
     function c2g_basins_nattr(res,nattr) bind(c,name="c2g_basins_nattr")
       import :: c_int, c_ptr
       type(c_ptr), value :: res
       integer(c_int) :: nattr
       integer(c_int) :: c2g_basins_nattr
     end function c2g_basins_nattr
     function c2g_basins_weight_grid(res,idb,handle) bind(c,name="c2g_basins_weight_grid")
       import :: c_int, c_ptr
       type(c_ptr), value :: res
       integer(c_int), value :: idb
       integer(c_int) :: handle
       integer(c_int) :: c2g_basins_weight_grid
     end function c2g_basins_weight_grid
     function c2g_basins_remap(ctx,res,xattr,c2x,isortho,isortho_del,x2c,x2xr,xr2c,nws,ws_ineighc,maxattn,nattn,&
        iatt,ilvec,idg1) bind(c,name="c2g_basins_remap")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx, res
       real(c_double) :: xattr(3,*), c2x(3,3), x2c(3,3), x2xr(3,3), xr2c(3,3), ws_ineighc(3,*)
       integer(c_int), value :: isortho, isortho_del, nws, maxattn
       integer(c_int) :: nattn, iatt(*), ilvec(3,*)
       type(c_ptr), value :: idg1          ! c_loc of idg1(n1,n2,n3), or c_null_ptr (yt_remap)
       integer(c_int) :: c2g_basins_remap
     end function c2g_basins_remap
     function c2g_promolecular_grid(ctx,n,x2c,nat,xat,ispc,nspc,spc_ngrid,spc_off,spc_a,spc_b,spc_rmax,spc_rcut,rtab,ftab,&
        infrag,handle) bind(c,name="c2g_promolecular_grid")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int) :: n(3), ispc(*), spc_ngrid(*), spc_off(*), handle
       integer(c_int), value :: nat, nspc
       real(c_double) :: x2c(3,3), xat(3,*), spc_a(*), spc_b(*), spc_rmax(*), spc_rcut(*), rtab(*), ftab(*)
       type(c_ptr), value :: infrag        ! c_loc of an integer(c_signed_char) mask(nat), or c_null_ptr
       integer(c_int) :: c2g_promolecular_grid
     end function c2g_promolecular_grid
     function c2g_hirshfeld_integrate(ctx,hpromol,x2c,nat,xat,ispc,nspc,spc_ngrid,spc_off,spc_a,spc_b,spc_rmax,spc_rcut,&
        rtab,ftab,domask,nprop,fieldhandles,omega,psum,vol) bind(c,name="c2g_hirshfeld_integrate")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: hpromol, nat, nspc, nprop
       integer(c_int) :: ispc(*), spc_ngrid(*), spc_off(*), fieldhandles(*)
       real(c_double) :: x2c(3,3), xat(3,*), spc_a(*), spc_b(*), spc_rmax(*), spc_rcut(*), rtab(*), ftab(*)
       type(c_ptr), value :: domask
       real(c_double), value :: omega
       real(c_double) :: psum(*), vol(*)
       integer(c_int) :: c2g_hirshfeld_integrate
     end function c2g_hirshfeld_integrate
     function c2g_basins_maxima(res,pmax) bind(c,name="c2g_basins_maxima")
       import :: c_int, c_ptr
       type(c_ptr), value :: res
       integer(c_int) :: pmax(3,*)
       integer(c_int) :: c2g_basins_maxima
     end function c2g_basins_maxima
     function c2g_basins_set_map(res,nattr,map) bind(c,name="c2g_basins_set_map")
       import :: c_int, c_ptr
       type(c_ptr), value :: res
       integer(c_int), value :: nattr
       integer(c_int) :: map(*)
       integer(c_int) :: c2g_basins_set_map
     end function c2g_basins_set_map
     function c2g_basins_relabel(res,nattr0,assigned,nattr_new) bind(c,name="c2g_basins_relabel")
       import :: c_int, c_ptr
       type(c_ptr), value :: res
       integer(c_int), value :: nattr0, nattr_new
       integer(c_int) :: assigned(*)
       integer(c_int) :: c2g_basins_relabel
     end function c2g_basins_relabel
     function c2g_basins_labels(res,idg) bind(c,name="c2g_basins_labels")
       import :: c_int, c_ptr
       type(c_ptr), value :: res
       integer(c_int) :: idg(*)
       integer(c_int) :: c2g_basins_labels
     end function c2g_basins_labels
     subroutine c2g_basins_free(res) bind(c,name="c2g_basins_free")
       import :: c_ptr
       type(c_ptr), value :: res
     end subroutine c2g_basins_free
     function c2g_integrate(ctx,res,nprop,fieldhandles,omega,psum,vol) bind(c,name="c2g_integrate")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx, res
       integer(c_int), value :: nprop
       integer(c_int) :: fieldhandles(*)
       real(c_double), value :: omega
       real(c_double) :: psum(*), vol(*)
       integer(c_int) :: c2g_integrate
     end function c2g_integrate
     function c2g_integrate_multipoles(ctx,res,fieldhandle,lmax,xattr,domask,isortho,isortho_del,x2c,x2xr,xr2c,&
        nws,ws_ineighc,omega,mpole) bind(c,name="c2g_integrate_multipoles")
       import :: c_int, c_ptr, c_double, c_signed_char
       type(c_ptr), value :: ctx, res
       integer(c_int), value :: fieldhandle, lmax, isortho, isortho_del, nws
       real(c_double) :: xattr(3,*), x2c(3,3), x2xr(3,3), xr2c(3,3), ws_ineighc(3,*)
       integer(c_signed_char) :: domask(*)
       real(c_double), value :: omega
       real(c_double) :: mpole(*)
       integer(c_int) :: c2g_integrate_multipoles
     end function c2g_integrate_multipoles
     function c2g_nci_rdg(ctx,handle,x0,xmat,nstep,c2x,x2c,c2xl,nnuc,nuc,crho,cgrad) bind(c,name="c2g_nci_rdg")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: handle, nnuc
       real(c_double) :: x0(3), xmat(3,3), c2x(3,3), x2c(3,3), c2xl(3,3), nuc(3,*)
       integer(c_int) :: nstep(3)
       real(c_double) :: crho(*), cgrad(*)
       integer(c_int) :: c2g_nci_rdg
     end function c2g_nci_rdg
     ! asynchronous upload (copy stream): f must stay allocated and unchanged until c2g_synchronize
     function c2g_grid_upload_async(ctx,f,n,handle) bind(c,name="c2g_grid_upload_async")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double) :: f(*)
       integer(c_int) :: n(3), handle
       integer(c_int) :: c2g_grid_upload_async
     end function c2g_grid_upload_async
     function c2g_basins_labels_async(res,idg) bind(c,name="c2g_basins_labels_async")
       import :: c_int, c_ptr
       type(c_ptr), value :: res
       integer(c_int) :: idg(*)
       integer(c_int) :: c2g_basins_labels_async
     end function c2g_basins_labels_async
     function c2g_synchronize(ctx) bind(c,name="c2g_synchronize")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
       integer(c_int) :: c2g_synchronize
     end function c2g_synchronize
     ! formatted-text numeric block (read_cube / read_vasp) -> resident grid
     function c2g_grid_parse_text(ctx,text,nbytes,n,order,divisor,handle,consumed,nhost) bind(c,name="c2g_grid_parse_text")
       import :: c_int, c_ptr, c_double, c_size_t, c_long_long, c_char
       type(c_ptr), value :: ctx
       character(kind=c_char) :: text(*)
       integer(c_size_t), value :: nbytes
       integer(c_int) :: n(3), handle
       integer(c_int), value :: order
       real(c_double), value :: divisor
       integer(c_size_t) :: consumed
       integer(c_long_long) :: nhost
       integer(c_int) :: c2g_grid_parse_text
     end function c2g_grid_parse_text
     ! formatted output of a resident grid (writegrid_cube / write_cube_body value loops)
     function c2g_grid_format_text(ctx,handle,layout,ishift,width,digits,scale,out,cap,nbytes) bind(c,name="c2g_grid_format_text")
       import :: c_int, c_ptr, c_size_t, c_char
       type(c_ptr), value :: ctx
       integer(c_int), value :: handle, layout, width, digits, scale
       integer(c_int) :: ishift(3)
       character(kind=c_char) :: out(*)
       integer(c_size_t), value :: cap
       integer(c_size_t) :: nbytes
       integer(c_int) :: c2g_grid_format_text
     end function c2g_grid_format_text
     function c2g_fft_derivative(ctx,handle,iff,x2c,hout) bind(c,name="c2g_fft_derivative")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: handle, iff
       real(c_double) :: x2c(3,3)
       integer(c_int) :: hout
       integer(c_int) :: c2g_fft_derivative
     end function c2g_fft_derivative
     function c2g_grid_download(ctx,handle,f) bind(c,name="c2g_grid_download")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int), value :: handle
       real(c_double) :: f(*)
       integer(c_int) :: c2g_grid_download
     end function c2g_grid_download
     function c2g_nci_rdg_fourier(ctx,h,x0,xmat,nstep,c2x,c2xl,crho,cgrad) bind(c,name="c2g_nci_rdg_fourier")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int) :: h(5), nstep(3)
       real(c_double) :: x0(3), xmat(3,3), c2x(3,3), c2xl(3,3)
       real(c_double) :: crho(*), cgrad(*)
       integer(c_int) :: c2g_nci_rdg_fourier
     end function c2g_nci_rdg_fourier
  end interface

contains

  !> Turn a non-zero status into a fatal critic2 error (tools_io ferror).
  subroutine check(ier,routine)
    use tools_io, only: ferror, faterr
    use c_interface_module, only: c_f_string_alloc
    integer(c_int), intent(in) :: ier
    character*(*), intent(in) :: routine
    character(len=:), allocatable :: msg
    if (ier /= 0) then
       call c_f_string_alloc(c2g_last_error(ctx),msg)
       call ferror(routine,"GPU: " // msg,faterr)
    end if
  end subroutine check

  subroutine gpu_init()
    character(len=8) :: val
    integer :: stat, ngpus
    call get_environment_variable("CRITIC2_GPU",val,status=stat)
    if (stat /= 0) return
    if (trim(val) /= "1") return
    ! CRITIC2_GPUS=N (2, 4, 8): the same calls drive N devices of the node; grids are sharded as z-slabs inside the library
    ngpus = 1
    call get_environment_variable("CRITIC2_GPUS",val,status=stat)
    if (stat == 0) read (val,*,iostat=stat) ngpus
    if (stat /= 0 .or. ngpus < 1) ngpus = 1
    if (ngpus > 1) then
       call check(c2g_init_devices(int(ngpus,c_int),ctx),"gpu_init")
    else
       call check(c2g_init(0_c_int,ctx),"gpu_init")
    end if
    gpu_enabled = .true.
  end subroutine gpu_init

  subroutine gpu_end()
    if (c_associated(basins)) call c2g_basins_free(basins)
    if (c_associated(ctx)) call c2g_finalize(ctx)
    basins = c_null_ptr
    ctx = c_null_ptr
    gpu_enabled = .false.
  end subroutine gpu_end

  !> Map the device's maxima to attractors with the host's own rules (atoms, known nnm, new nnm); this
  !> is the per-maximum part of bader_integrate / yt_integrate that needs the crystal object.
  subroutine identify_attractors(s,bas,nmax,pmax,vsmall_thr,map)
    use systemmod, only: system
    use types, only: basindat, realloc
    use param, only: icrd_crys
    type(system), intent(inout) :: s
    type(basindat), intent(inout) :: bas
    integer, intent(in) :: nmax
    integer(c_int), intent(in) :: pmax(3,nmax)
    real*8, intent(in) :: vsmall_thr
    integer(c_int), intent(out) :: map(nmax)
    integer :: i, l, nid
    real*8 :: dv(3)

    do i = 1, nmax
       dv = real(pmax(:,i)-1,8) / real(bas%n,8)
       map(i) = 0
       if (bas%atexist) then
          nid = s%c%identify_atom(dv,icrd_crys,distmax=bas%ratom)
          if (nid > 0) map(i) = nid
       end if
       if (map(i) == 0 .and. bas%ratom > vsmall_thr) then
          do l = 1, bas%nattr
             if (s%c%are_lclose(dv,bas%xattr(:,l),bas%ratom)) then
                map(i) = l
                exit
             end if
          end do
       end if
       if (map(i) == 0) then
          ! (a DISCARD expression would be evaluated here, exactly as in the CPU path)
          bas%nattr = bas%nattr + 1
          if (bas%nattr > size(bas%xattr,2)) call realloc(bas%xattr,3,2*bas%nattr)
          bas%xattr(:,bas%nattr) = dv
          map(i) = bas%nattr
       end if
    end do
  end subroutine identify_attractors

  !> GPU body of bader_integrate: replaces the scan + refine of src/bader@proc.f90:147-224.
  !> lat2car, car2lat, lat_i_dist are the module variables computed at :124-145.
  subroutine gpu_bader_integrate(s,bas,car2lat,lat_i_dist)
    use systemmod, only: system
    use types, only: basindat, realloc
    use param, only: vsmall
    type(system), intent(inout) :: s
    type(basindat), intent(inout) :: bas
    real*8, intent(in) :: car2lat(3,3), lat_i_dist(-1:1,-1:1,-1:1)
    integer(c_int) :: nmax, n(3)
    integer(c_int), allocatable :: pmax(:,:), map(:)
    real(c_double) :: lid(27)
    integer :: i, j, k

    n = int(bas%n,c_int)
    ! flatten lat_i_dist as (d1+1)*9+(d2+1)*3+(d3+1)
    do i = -1, 1
       do j = -1, 1
          do k = -1, 1
             lid((i+1)*9+(j+1)*3+(k+1)+1) = lat_i_dist(i,j,k)
          end do
       end do
    end do
    if (hgrid >= 0) call check(c2g_grid_free(ctx,hgrid),"gpu_bader_integrate")
    call check(c2g_grid_upload(ctx,bas%f,n,hgrid),"gpu_bader_integrate")
    if (c_associated(basins)) call c2g_basins_free(basins)
    call check(c2g_bader_assign(ctx,hgrid,car2lat,lid,0_c_int,1_c_int,nmax,basins),"gpu_bader_integrate")
    allocate(pmax(3,nmax),map(nmax))
    call check(c2g_basins_maxima(basins,pmax),"gpu_bader_integrate")
    call identify_attractors(s,bas,int(nmax),pmax,vsmall,map)
    call check(c2g_basins_set_map(basins,int(bas%nattr,c_int),map),"gpu_bader_integrate")
    if (allocated(bas%idg)) deallocate(bas%idg)
    allocate(bas%idg(bas%n(1),bas%n(2),bas%n(3)))
    call check(c2g_basins_labels(basins,bas%idg),"gpu_bader_integrate")
    call realloc(bas%xattr,3,bas%nattr)
  end subroutine gpu_bader_integrate

  !> GPU body of yt_integrate (src/yt@proc.f90:77-211).  The ytdata scratch file is not written: the
  !> weights stay on the device and gpu_integrate_fields uses them; bas%luw is left at 0.
  subroutine gpu_yt_integrate(s,bas)
    use systemmod, only: system
    use types, only: basindat, realloc
    use param, only: vsmall
    type(system), intent(inout) :: s
    type(basindat), intent(inout) :: bas
    integer(c_int) :: nmax, n(3), nvec
    integer(c_int), allocatable :: pmax(:,:), map(:)

    n = int(bas%n,c_int)
    nvec = int(s%f(s%iref)%grid%nvec,c_int)
    if (hgrid >= 0) call check(c2g_grid_free(ctx,hgrid),"gpu_yt_integrate")
    call check(c2g_grid_upload(ctx,bas%f,n,hgrid),"gpu_yt_integrate")
    if (c_associated(basins)) call c2g_basins_free(basins)
    call check(c2g_yt_build(ctx,hgrid,nvec,s%f(s%iref)%grid%vec,s%f(s%iref)%grid%area,nmax,basins),"gpu_yt_integrate")
    allocate(pmax(3,nmax),map(nmax))
    call check(c2g_basins_maxima(basins,pmax),"gpu_yt_integrate")
    call identify_attractors(s,bas,int(nmax),pmax,vsmall,map)
    call check(c2g_basins_set_map(basins,int(bas%nattr,c_int),map),"gpu_yt_integrate")
    if (allocated(bas%idg)) deallocate(bas%idg)
    allocate(bas%idg(bas%n(1),bas%n(2),bas%n(3)))
    call check(c2g_basins_labels(basins,bas%idg),"gpu_yt_integrate")   ! spatial ids, 0 = IAS point
    call realloc(bas%xattr,3,bas%nattr)
  end subroutine gpu_yt_integrate

  !> GPU body of voronoi_grid (src/hirshfeld@proc.f90:93-122): the atoms are the attractors, bas%idg = nearest atom of
  !> every grid node (crystal%nearest_atom_grid).  The basins stay on the device for gpu_integrate_fields.
  subroutine gpu_voronoi_grid(s,bas)
    use systemmod, only: system
    use types, only: basindat
    type(system), intent(inout) :: s
    type(basindat), intent(inout) :: bas
    integer :: i
    real(c_double), allocatable :: xat(:,:)

    if (allocated(bas%xattr)) deallocate(bas%xattr)
    allocate(bas%xattr(3,s%f(s%iref)%ncpcel))
    bas%xattr = 0d0
    bas%nattr = 0
    if (bas%atexist) then
       bas%nattr = s%f(s%iref)%ncpcel
       do i = 1, s%f(s%iref)%ncpcel
          bas%xattr(:,i) = s%f(s%iref)%cpcel(i)%x
       end do
    end if
    allocate(xat(3,s%c%ncel))
    do i = 1, s%c%ncel
       xat(:,i) = s%c%atcel(i)%x
    end do
    if (c_associated(basins)) call c2g_basins_free(basins)
    call check(c2g_voronoi_grid(ctx,int(bas%n,c_int),s%c%m_x2c,int(s%c%ncel,c_int),xat,basins),"gpu_voronoi_grid")
    if (allocated(bas%idg)) deallocate(bas%idg)
    allocate(bas%idg(bas%n(1),bas%n(2),bas%n(3)))
    call check(c2g_basins_labels(basins,bas%idg),"gpu_voronoi_grid")
  end subroutine gpu_voronoi_grid

  !> The ytdata record of the last gpu_yt_integrate (yt.f90:36-45), for the consumers that call the host's own
  !> yt_weights(din=...) (BASINS, DI: integration@proc.f90:1125-1158, yt@proc.f90:399-530).  Same arrays as the ones
  !> yt_integrate writes to bas%luw (:191-199), in the order of a stable (density, index) sort.
  subroutine gpu_yt_export(bas,dat)
    use yt, only: ytdata
    use systemmod, only: sy
    use types, only: basindat
    type(basindat), intent(in) :: bas
    type(ytdata), intent(inout), target :: dat
    integer :: nn, nvec

    nn = bas%n(1)*bas%n(2)*bas%n(3)
    nvec = sy%f(sy%iref)%grid%nvec
    dat%nbasin = bas%nattr
    dat%nn = nn
    dat%nvec = nvec
    if (allocated(dat%nlo)) deallocate(dat%nlo,dat%ibasin,dat%iio,dat%inear,dat%fnear)
    allocate(dat%nlo(nn),dat%ibasin(nn),dat%iio(nn),dat%inear(nvec,nn),dat%fnear(nvec,nn))
    call check(c2g_yt_export(basins,dat%nlo,dat%ibasin,dat%iio,c_loc(dat%inear),c_loc(dat%fnear)),"gpu_yt_export")
  end subroutine gpu_yt_export

  !> GPU body of yt_isosurface (src/yt@proc.f90:233-390) for an empty DISCARD expression (with an expression the
  !> caller keeps the CPU routine: the expression is evaluated by the host parser, :304-311).  Fills bas%nattr,
  !> bas%xattr and bas%idg like the reference (surviving regions keep their discovery numbers, :337-359); the
  !> regions stay on the device for gpu_integrate_fields / gpu_integrate_multipoles.
  subroutine gpu_yt_isosurface(s,bas)
    use systemmod, only: system
    use types, only: basindat, realloc
    type(system), intent(inout) :: s
    type(basindat), intent(inout) :: bas
    integer(c_int) :: nmax, n(3), nvec, nraw, nattr
    integer(c_int), allocatable :: pmax(:,:)
    type(c_ptr) :: yt, regions
    integer :: i

    n = int(bas%n,c_int)
    nvec = int(s%f(s%iref)%grid%nvec,c_int)
    if (hgrid >= 0) call check(c2g_grid_free(ctx,hgrid),"gpu_yt_isosurface")
    call check(c2g_grid_upload(ctx,bas%f,n,hgrid),"gpu_yt_isosurface")
    if (c_associated(basins)) call c2g_basins_free(basins)
    call check(c2g_yt_build(ctx,hgrid,nvec,s%f(s%iref)%grid%vec,s%f(s%iref)%grid%area,nmax,yt),"gpu_yt_isosurface")
    call check(c2g_yt_isosurface(yt,bas%isov,nraw,nattr,regions),"gpu_yt_isosurface")
    call c2g_basins_free(yt)
    basins = regions
    allocate(pmax(3,max(nraw,1)))
    call check(c2g_basins_maxima(basins,pmax),"gpu_yt_isosurface")
    if (allocated(bas%xattr)) deallocate(bas%xattr)
    allocate(bas%xattr(3,max(nraw,1)))
    do i = 1, nraw
       bas%xattr(:,i) = real(pmax(:,i)-1,8) / real(bas%n,8)      ! dv of :301
    end do
    bas%nattr = nattr                                           ! :352
    if (allocated(bas%idg)) deallocate(bas%idg)
    allocate(bas%idg(bas%n(1),bas%n(2),bas%n(3)))
    call check(c2g_basins_labels(basins,bas%idg),"gpu_yt_isosurface")
    call realloc(bas%xattr,3,bas%nattr)                         ! :359
  end subroutine gpu_yt_isosurface

  !> GPU body of the two per-attractor loops of intgrid_fields (src/integration@proc.f90:1205-1219 and
  !> :1288-1301).  fint holds the nprop integrand grids already built by the host code (:1235-1280).
  subroutine gpu_integrate_fields(bas,nprop,fint,omega,psum,vol,assigned,nattr_new)
    use types, only: basindat
    type(basindat), intent(in) :: bas
    integer, intent(in) :: nprop
    real*8, intent(in) :: fint(:,:,:,:)
    real*8, intent(in) :: omega
    real*8, intent(out) :: psum(:,:), vol(:)
    integer, intent(in), optional :: assigned(:), nattr_new
    integer(c_int) :: h(nprop), n(3), nrow
    real*8, allocatable :: psum0(:,:), vol0(:)
    integer :: k

    n = int(bas%n,c_int)
    if (present(assigned)) then   ! relabelling decided by int_reorder_gridout (:1069-1110)
       call check(c2g_basins_relabel(basins,int(size(assigned),c_int),int(assigned,c_int),int(nattr_new,c_int)),&
          "gpu_integrate_fields")
    end if
    do k = 1, nprop
       call check(c2g_grid_upload(ctx,fint(:,:,:,k),n,h(k)),"gpu_integrate_fields")
    end do
    ! rows written by the library: bas%nattr, except after gpu_yt_isosurface (region ids run beyond the number of
    ! surviving regions); the reference's loops stop at bas%nattr (:1208, :1290) and so does this copy
    call check(c2g_basins_nattr(basins,nrow),"gpu_integrate_fields")
    allocate(psum0(nrow,max(nprop,1)),vol0(nrow))
    call check(c2g_integrate(ctx,basins,int(nprop,c_int),h,omega,psum0,vol0),"gpu_integrate_fields")
    psum(1:bas%nattr,1:nprop) = psum0(1:bas%nattr,1:nprop)
    vol(1:bas%nattr) = vol0(1:bas%nattr)
    ! ONLY / ONLY_RANGE: the reference skips the attractors with docelatom = .false. and leaves their sums at zero
    ! (:1210, :1292); the device integrates every basin in the same pass, so the rows are cleared here
    if (allocated(bas%docelatom) .and. allocated(bas%icp)) then
       do k = 1, bas%nattr
          if (.not.bas%docelatom(bas%icp(k))) then
             psum(k,1:nprop) = 0d0
             vol(k) = 0d0
          end if
       end do
    end if
    do k = 1, nprop
       call check(c2g_grid_free(ctx,h(k)),"gpu_integrate_fields")
    end do
  end subroutine gpu_integrate_fields

  !> GPU body of the multipole branch of intgrid_fields (src/integration@proc.f90:1302-1361): replaces both the
  !> YT loop over basins (:1316-1336) and the Bader/isosurface loop under omp critical (:1338-1358), and the final
  !> scaling (:1360).  fint is the integrand grid of property k; mpole((lmax+1)**2,bas%nattr) = res(k)%mpole.
  subroutine gpu_integrate_multipoles(c,bas,lmax,fint,mpole)
    use crystalmod, only: crystal
    use types, only: basindat
    use tools_io, only: ferror, faterr
    type(crystal), intent(in) :: c
    type(basindat), intent(in) :: bas
    integer, intent(in) :: lmax
    real*8, intent(in) :: fint(:,:,:)
    real*8, intent(out) :: mpole(:,:)
    integer(c_int) :: h, n(3), nws, nrow
    integer(c_signed_char) :: domask(max(bas%nattr,1))
    real*8 :: wsdum(3,1)
    integer :: m

    n = int(bas%n,c_int)
    ! after gpu_yt_isosurface with merged regions the ids in idg exceed bas%nattr and the reference itself reads
    ! xattr(:,ix) out of bounds (:1346): refuse instead of guessing
    call check(c2g_basins_nattr(basins,nrow),"gpu_integrate_multipoles")
    if (nrow /= bas%nattr) &
       call ferror("gpu_integrate_multipoles","region ids exceed the number of attractors (merged isosurface regions)",faterr)
    do m = 1, bas%nattr   ! bas%docelatom(bas%icp(m)) of :1318; the Bader branch does not look at it
       domask(m) = merge(1_c_signed_char,0_c_signed_char,bas%docelatom(bas%icp(m)))
    end do
    call check(c2g_grid_upload(ctx,fint,n,h),"gpu_integrate_multipoles")
    nws = 0
    if (allocated(c%ws_ineighc)) nws = int(c%ws_nf,c_int)
    if (nws > 0) then
       call check(c2g_integrate_multipoles(ctx,basins,h,int(lmax,c_int),bas%xattr,domask,&
          merge(1_c_int,0_c_int,c%isortho),merge(1_c_int,0_c_int,c%isortho_del),c%m_x2c,c%m_x2xr,c%m_xr2c,&
          nws,c%ws_ineighc,c%omega,mpole),"gpu_integrate_multipoles")
    else
       call check(c2g_integrate_multipoles(ctx,basins,h,int(lmax,c_int),bas%xattr,domask,&
          merge(1_c_int,0_c_int,c%isortho),merge(1_c_int,0_c_int,c%isortho_del),c%m_x2c,c%m_x2xr,c%m_xr2c,&
          0_c_int,wsdum,c%omega,mpole),"gpu_integrate_multipoles")
    end if
    call check(c2g_grid_free(ctx,h),"gpu_integrate_multipoles")
  end subroutine gpu_integrate_multipoles

  !> GPU body of the nciplot loop (src/nci@proc.f90:540-606), grid interpolation mode, no fragments.
  subroutine gpu_nci_rdg(f,x0,xmat,nstep,m_c2x,m_x2c,c2xl,nuc,crho,cgrad)
    real*8, intent(in) :: f(:,:,:), x0(3), xmat(3,3), m_c2x(3,3), m_x2c(3,3), c2xl(3,3), nuc(:,:)
    integer, intent(in) :: nstep(3)
    real*8, intent(out) :: crho(0:,0:,0:), cgrad(0:,0:,0:)
    integer(c_int) :: h, n(3)

    n = int(shape(f),c_int)
    call check(c2g_grid_upload(ctx,f,n,h),"gpu_nci_rdg")
    call check(c2g_nci_rdg(ctx,h,x0,xmat,int(nstep,c_int),m_c2x,m_x2c,c2xl,int(size(nuc,2),c_int),nuc,crho,cgrad),&
       "gpu_nci_rdg")
    call check(c2g_grid_free(ctx,h),"gpu_nci_rdg")
  end subroutine gpu_nci_rdg

  !> grid3%fft on the device (grid3mod@proc.f90:1757-1872): fnew%f = FFT-derived field of fold%f.
  !> iff is one of the ifformat_as_ft_* codes of param.F90 (the C header uses the same numbers).
  !> Call site: the body of grid3%fft after copy_geometry (:1781-1784), i.e. LOAD AS LAP/GRAD/..., the
  !> lap/gmod integrables and the derived grids of NCIPLOT FOURIER.
  subroutine gpu_grid_fft(fnew,fold,x2c,iff)
    real*8, intent(inout) :: fnew(:,:,:)
    real*8, intent(in) :: fold(:,:,:)
    real*8, intent(in) :: x2c(3,3)
    integer, intent(in) :: iff
    integer(c_int) :: n(3), h, hout

    n = int(shape(fold),c_int)
    call check(c2g_grid_upload(ctx,fold,n,h),"gpu_grid_fft")
    call check(c2g_fft_derivative(ctx,h,int(iff,c_int),x2c,hout),"gpu_grid_fft")
    call check(c2g_grid_download(ctx,hout,fnew),"gpu_grid_fft")
    call check(c2g_grid_free(ctx,h),"gpu_grid_fft")
    call check(c2g_grid_free(ctx,hout),"gpu_grid_fft")
  end subroutine gpu_grid_fft

  !> NCIPLOT loop with FOURIER interpolation (nci@proc.f90:527-565): the four derived grids of :528-531 are
  !> built on the device from the reference field and never leave it.
  subroutine gpu_nci_rdg_fourier(f,x2c,x0,xmat,nstep,m_c2x,c2xl,crho,cgrad)
    use param, only: ifformat_as_ft_grad, ifformat_as_ft_xx, ifformat_as_ft_yy, ifformat_as_ft_zz
    real*8, intent(in) :: f(:,:,:)
    real*8, intent(in) :: x2c(3,3), x0(3), xmat(3,3), m_c2x(3,3), c2xl(3,3)
    integer, intent(in) :: nstep(3)
    real*8, intent(inout) :: crho(0:,0:,0:), cgrad(0:,0:,0:)
    integer(c_int) :: n(3), h(5), i
    integer, parameter :: iffs(4) = (/ifformat_as_ft_grad,ifformat_as_ft_xx,ifformat_as_ft_yy,ifformat_as_ft_zz/)

    n = int(shape(f),c_int)
    call check(c2g_grid_upload(ctx,f,n,h(1)),"gpu_nci_rdg_fourier")
    do i = 1, 4
       call check(c2g_fft_derivative(ctx,h(1),int(iffs(i),c_int),x2c,h(i+1)),"gpu_nci_rdg_fourier")
    end do
    call check(c2g_nci_rdg_fourier(ctx,h,x0,xmat,int(nstep,c_int),m_c2x,c2xl,crho,cgrad),"gpu_nci_rdg_fourier")
    do i = 1, 5
       call check(c2g_grid_free(ctx,h(i)),"gpu_nci_rdg_fourier")
    end do
  end subroutine gpu_nci_rdg_fourier

  !> The numeric block of a cube (order=1) or CHGCAR (order=0) file: text = the bytes after the header lines
  !> (read with stream access), f = the grid, divided by `divisor` (det3(x2c) for read_vasp with vscal).
  !> Replaces the list-directed READ of grid3mod@proc.f90:559 (read_cube) and :884 (read_vasp).
  subroutine gpu_read_text_block(text,order,divisor,f,consumed)
    character(kind=c_char), intent(in) :: text(:)
    integer, intent(in) :: order
    real*8, intent(in) :: divisor
    real*8, intent(inout) :: f(:,:,:)
    integer(c_size_t), intent(out) :: consumed
    integer(c_int) :: n(3), h
    integer(c_long_long) :: nhost

    n = int(shape(f),c_int)
    call check(c2g_grid_parse_text(ctx,text,int(size(text),c_size_t),n,int(order,c_int),divisor,h,consumed,nhost),&
       "gpu_read_text_block")
    call check(c2g_grid_download(ctx,h,f),"gpu_read_text_block")
    call check(c2g_grid_free(ctx,h),"gpu_read_text_block")
  end subroutine gpu_read_text_block

  !> The value block of a cube file from a host array: g(i,j,k) in cube order with the shift of writegrid_cube
  !> (layout 1: width 12, digits 5, scale 1; precisecube: 22, 14, 0), or c(k,j,i) of write_cube_body (layout 0:
  !> 13, 5, 1).  The text is written to unit lu (opened with access="stream") in one piece.
  subroutine gpu_write_text_block(lu,g,layout,ishift,width,digits,scale)
    integer, intent(in) :: lu, layout, ishift(3), width, digits, scale
    real*8, intent(in) :: g(:,:,:)
    integer(c_int) :: n(3), h
    integer(c_size_t) :: nbytes
    character(kind=c_char), allocatable :: text(:)

    n = int(shape(g),c_int)
    call check(c2g_grid_upload(ctx,g,n,h),"gpu_write_text_block")
    allocate(text(1))
    call check(c2g_grid_format_text(ctx,h,int(layout,c_int),int(ishift,c_int),int(width,c_int),int(digits,c_int),&
       int(scale,c_int),text,0_c_size_t,nbytes),"gpu_write_text_block")   ! step 1: size (cap = 0 with out ignored)
    deallocate(text)
    allocate(text(nbytes))
    call check(c2g_grid_format_text(ctx,h,int(layout,c_int),int(ishift,c_int),int(width,c_int),int(digits,c_int),&
       int(scale,c_int),text,nbytes,nbytes),"gpu_write_text_block")
    write (lu) text
    call check(c2g_grid_free(ctx,h),"gpu_write_text_block")
  end subroutine gpu_write_text_block

  !> GPU body of bader_remap (src/bader@proc.f90:237-296) and yt_remap (src/yt@proc.f90:533-594): the attractor
  !> images used by the delocalization indices.  Same outputs as the reference: nattn, iatt, ilvec and, for Bader
  !> basins, idg1.  The basins of the last gpu_bader_integrate / gpu_yt_integrate call are used.
  subroutine gpu_basins_remap(c,bas,nattn,ilvec,iatt,idg1)
    use crystalmod, only: crystal
    use types, only: basindat
    type(crystal), intent(in) :: c
    type(basindat), intent(in) :: bas
    integer, intent(out) :: nattn
    integer, allocatable, intent(inout) :: iatt(:), ilvec(:,:)
    integer, allocatable, intent(inout), target, optional :: idg1(:,:,:)
    integer(c_int) :: cap, ier, nws, nn
    integer(c_int), allocatable :: iatt_(:), ilvec_(:,:)
    type(c_ptr) :: pidg1
    real*8 :: wsdum(3,1)

    pidg1 = c_null_ptr
    if (present(idg1)) then
       if (allocated(idg1)) deallocate(idg1)
       allocate(idg1(bas%n(1),bas%n(2),bas%n(3)))
       pidg1 = c_loc(idg1)
    end if
    nws = 0
    if (allocated(c%ws_ineighc)) nws = int(c%ws_nf,c_int)
    cap = int(27*max(bas%nattr,1),c_int)
    do
       if (allocated(iatt_)) deallocate(iatt_,ilvec_)
       allocate(iatt_(cap),ilvec_(3,cap))
       if (nws > 0) then
          ier = c2g_basins_remap(ctx,basins,bas%xattr,c%m_c2x,merge(1_c_int,0_c_int,c%isortho),&
             merge(1_c_int,0_c_int,c%isortho_del),c%m_x2c,c%m_x2xr,c%m_xr2c,nws,c%ws_ineighc,cap,nn,iatt_,ilvec_,pidg1)
       else
          ier = c2g_basins_remap(ctx,basins,bas%xattr,c%m_c2x,merge(1_c_int,0_c_int,c%isortho),&
             merge(1_c_int,0_c_int,c%isortho_del),c%m_x2c,c%m_x2xr,c%m_xr2c,0_c_int,wsdum,cap,nn,iatt_,ilvec_,pidg1)
       end if
       if (ier == 6 .and. nn > cap) then     ! C2G_ERR_OVERFLOW: nn = capacity needed
          cap = nn
          cycle
       end if
       call check(ier,"gpu_basins_remap")
       exit
    end do
    nattn = nn
    if (allocated(iatt)) deallocate(iatt)
    if (allocated(ilvec)) deallocate(ilvec)
    allocate(iatt(nattn),ilvec(3,nattn))
    iatt = iatt_(1:nattn)
    ilvec = ilvec_(:,1:nattn)
  end subroutine gpu_basins_remap

  !> Pack the atomic radial grids of the species of c (grid1mod agrid) and its atoms for the HIRSHFELD entry points:
  !> cutoffs as in promolecular_atom (src/crystalmod@env.f90:671-684).
  subroutine pack_atomic_grids(c,xat,ispc,ngrid,off,a,b,rmax,rcut,rtab,ftab)
    use crystalmod, only: crystal
    use grid1mod, only: agrid
    use global, only: cutrad
    use param, only: maxzat
    type(crystal), intent(in) :: c
    real*8, allocatable, intent(out) :: xat(:,:), a(:), b(:), rmax(:), rcut(:), rtab(:), ftab(:)
    integer(c_int), allocatable, intent(out) :: ispc(:), ngrid(:), off(:)
    integer :: i, iz, ntab

    allocate(xat(3,c%ncel),ispc(c%ncel),ngrid(c%nspc),off(c%nspc),a(c%nspc),b(c%nspc),rmax(c%nspc),rcut(c%nspc))
    do i = 1, c%ncel
       xat(:,i) = c%atcel(i)%x
       ispc(i) = int(c%atcel(i)%is,c_int)
    end do
    ngrid = 0; off = 0; a = 1d0; b = 1d0; rmax = 0d0; rcut = 0d0
    ntab = 0
    do i = 1, c%nspc
       iz = c%spc(i)%z
       if (iz == 0 .or. iz > maxzat) cycle
       if (.not.agrid(iz)%isinit) cycle
       if (agrid(iz)%z - agrid(iz)%qat <= 0) cycle            ! interp returns zero (grid1mod@proc.f90:101)
       ngrid(i) = int(agrid(iz)%ngrid,c_int)
       off(i) = int(ntab,c_int)
       a(i) = agrid(iz)%a; b(i) = agrid(iz)%b; rmax(i) = agrid(iz)%rmax
       rcut(i) = min(cutrad(iz),agrid(iz)%rmax)
       ntab = ntab + agrid(iz)%ngrid
    end do
    allocate(rtab(max(ntab,1)),ftab(max(ntab,1)))
    do i = 1, c%nspc
       if (ngrid(i) == 0) cycle
       iz = c%spc(i)%z
       rtab(off(i)+1:off(i)+ngrid(i)) = agrid(iz)%r(1:ngrid(i))
       ftab(off(i)+1:off(i)+ngrid(i)) = agrid(iz)%f(1:ngrid(i))
    end do
  end subroutine pack_atomic_grids

  !> HIRSHFELD on a grid: bas%f = promolecular density (src/integration@proc.f90:264-267, promolecular_array3 without
  !> zpsp or fragment) kept resident, then the grid loop of intgrid_hirshfeld_fields (:1552-1596) for the nprop
  !> integrand grids of fint (built by the host code, :1452-1527) -- psum(nattr,nprop), vol(nattr) are the psuml
  !> columns already scaled by omega/ntot (:1590).  With download = .true. bas%f is also copied back to the host.
  subroutine gpu_hirshfeld_fields(c,bas,nprop,fint,psum,vol,download)
    use crystalmod, only: crystal
    use types, only: basindat
    type(crystal), intent(in) :: c
    type(basindat), intent(inout) :: bas
    integer, intent(in) :: nprop
    real*8, intent(in) :: fint(:,:,:,:)
    real*8, intent(out) :: psum(:,:), vol(:)
    logical, intent(in) :: download
    real*8, allocatable :: xat(:,:), a(:), b(:), rmax(:), rcut(:), rtab(:), ftab(:)
    integer(c_int), allocatable :: ispc(:), ngrid(:), off(:)
    integer(c_signed_char), allocatable, target :: domask(:)
    integer(c_int) :: hp, h(max(nprop,1)), n(3)
    integer :: k

    n = int(bas%n,c_int)
    call pack_atomic_grids(c,xat,ispc,ngrid,off,a,b,rmax,rcut,rtab,ftab)
    call check(c2g_promolecular_grid(ctx,n,c%m_x2c,int(c%ncel,c_int),xat,ispc,int(c%nspc,c_int),ngrid,off,a,b,rmax,rcut,&
       rtab,ftab,c_null_ptr,hp),"gpu_hirshfeld_fields")
    if (download) call check(c2g_grid_download(ctx,hp,bas%f),"gpu_hirshfeld_fields")
    allocate(domask(c%ncel))
    do k = 1, c%ncel          ! attractors = atoms of the complete list (hirsh_grid, src/hirshfeld@proc.f90:44-49)
       domask(k) = merge(1_c_signed_char,0_c_signed_char,bas%docelatom(bas%icp(k)))
    end do
    do k = 1, nprop
       call check(c2g_grid_upload(ctx,fint(:,:,:,k),n,h(k)),"gpu_hirshfeld_fields")
    end do
    call check(c2g_hirshfeld_integrate(ctx,hp,c%m_x2c,int(c%ncel,c_int),xat,ispc,int(c%nspc,c_int),ngrid,off,a,b,rmax,rcut,&
       rtab,ftab,c_loc(domask),int(nprop,c_int),h,c%omega,psum,vol),"gpu_hirshfeld_fields")
    do k = 1, nprop
       call check(c2g_grid_free(ctx,h(k)),"gpu_hirshfeld_fields")
    end do
    call check(c2g_grid_free(ctx,hp),"gpu_hirshfeld_fields")
  end subroutine gpu_hirshfeld_fields

  !> WCUBE (int_cubew, src/integration@proc.f90:4449-4462): the value block of the weight cube of attractor i, from
  !> the basins resident on the device -- the YT weights (:4451) or the indicator of idg == i (:4455-4458) -- written to
  !> unit lu (access="stream") after the header lines of writegrid_cube; the weights never visit the host.
  subroutine gpu_wcube_block(lu,i,ishift,precisecube)
    integer, intent(in) :: lu, i, ishift(3)
    logical, intent(in) :: precisecube
    integer(c_int) :: h, width, digits, scale
    integer(c_size_t) :: nbytes
    character(kind=c_char), allocatable :: text(:)

    width = 12; digits = 5; scale = 1                      ! (1p,6(" ",E12.5E3)), crystalmod@write.f90:3563
    if (precisecube) then
       width = 22; digits = 14; scale = 0                  ! (6(" ",E22.14E3)), :3559
    end if
    call check(c2g_basins_weight_grid(basins,int(i,c_int),h),"gpu_wcube_block")
    allocate(text(1))
    call check(c2g_grid_format_text(ctx,h,1_c_int,int(ishift,c_int),width,digits,scale,text,0_c_size_t,nbytes),"gpu_wcube_block")
    deallocate(text)
    allocate(text(nbytes))
    call check(c2g_grid_format_text(ctx,h,1_c_int,int(ishift,c_int),width,digits,scale,text,nbytes,nbytes),"gpu_wcube_block")
    write (lu) text
    call check(c2g_grid_free(ctx,h),"gpu_wcube_block")
  end subroutine gpu_wcube_block

end module critic2_gpu
