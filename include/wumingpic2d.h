/*
 * wumingpic2d.h -- C ABI of the B200-native per-timestep PIC hot path.
 *
 * This is the drop-in boundary for WumingPIC2D's hot-path module procedures
 * (Fortran 90, no existing C/FFI layer).  Each entry point names the reference
 * procedure it replaces (file:line relative to the reference tree); the Fortran
 * ISO_C_BINDING shim that binds them under the reference's own module/procedure
 * names is in fortran/ and INTEGRATION.md.
 *
 * Conventions
 *  - plain pointers and sizes only; every function returns 0 on success, non-zero
 *    on error, with the message available from wm_last_error() (the reference
 *    does `write + stop`, e.g. common/particle.f90:63-66; the shim does the same).
 *  - host arrays use the reference's Fortran (column-major) layout, per rank:
 *      up,gp  real(8) (6, np, nys:nye, nsp)           proj/weibel/app.f90:281-282
 *      uf     real(8) (6, nxgs-2:nxge+2, nys-2:nye+2) proj/weibel/app.f90:280
 *      uj     real(8) (3, nxgs-2:nxge+2, nys-2:nye+2) common/field.f90:108
 *      np2    integer (nys:nye, nsp)                  proj/weibel/app.f90:278
 *      cumcnt integer (nxgs:nxge+1, nys:nye, nsp)     proj/weibel/app.f90:279
 *      mom    real(8) (7, nxgs-1:nxge+1, nys-1:nye+1, nsp) proj/weibel/app.f90:283
 *  - device state is authoritative between calls ("resident" mode).  The library
 *    owns all device memory including the CG warm-start `df` that the reference
 *    keeps as a SAVE variable (common/field.f90:98).
 *  - one context = one GPU = one y-slab (common/mpi_set.f90:36-47).  Not re-entrant
 *    per context; contexts are independent.
 *  - there is no CPU fallback: every entry point fails if no CUDA device is usable.
 */
#ifndef WUMINGPIC2D_H
#define WUMINGPIC2D_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WM_NSP_MAX 2
#define WM_UNIQUE_ID_BYTES 128

/* boundary plugin (the module chosen by `use ... => bc__*` in proj/<problem>/app.f90:6-13) */
enum {
  WM_BC_PERIODIC = 0,      /* common/boundary_periodic.f90 (proj/weibel)                                         */
  WM_BC_RECONNECTION = 1,  /* proj/reconnection/boundary_reconnection.f90: reflecting particles / conducting     */
                           /* fields at the x walls nxs+1, nxe-1; periodic y                                     */
  WM_BC_SHOCK = 2          /* proj/shock/boundary_shock.f90: reflecting left wall, injection wall at             */
                           /* xend = nxe*delx + v0*delt on the right (bc__injection, applied BEFORE the field    */
                           /* solve: proj/shock/app.f90:112-113), df = 0 in the right ghost column; periodic y.   */
                           /* Active range nxs..nxe: wm_set_xrange.  wm_step needs wm_set_u_inject first.       */
};

/* flags */
enum {
  WM_FLAG_EXACT_PUSH = 1  /* push arithmetic without FMA contraction, IEEE sqrt/div in the
                             reference's operation order: bit-identical to the CPU path   */
};

/* All scalars the reference passes to particle__init (common/particle.f90:18),
 * field__init (common/field.f90:22), sort__init (common/sort.f90:16),
 * boundary_periodic__init (common/boundary_periodic.f90:26), mom_calc__init
 * (common/mom_calc.f90:18) and mpi_set__init (common/mpi_set.f90:19). */
typedef struct wm_config {
  int32_t ndim;             /* 6 */
  int32_t np;               /* host row capacity of up/gp (second dimension) */
  int32_t nsp;              /* 2 */
  int32_t nxgs, nxge, nygs, nyge; /* global cell index range */
  int32_t nys, nye;         /* this rank's rows */
  int32_t nrank, nsize;     /* rank / number of ranks (ring: nup=nrank+1, ndown=nrank-1, periodic) */
  int32_t bc;               /* WM_BC_* */
  int32_t device;           /* CUDA device ordinal, -1 = current */
  int32_t flags;            /* WM_FLAG_* */
  double delx, delt, c, gfac;
  double q[WM_NSP_MAX], r[WM_NSP_MAX];
  int64_t capacity;         /* device particle capacity per species; 0 = 1.25 x (np2 total at upload) */
} wm_config;

typedef struct wm_ctx wm_ctx;

const char *wm_last_error(void);
int wm_version(void);
/* hash of the sources the library was built from (the host layer refuses a binary that does not match the tree) */
const char *wm_source_hash(void);

/* ---- life cycle ------------------------------------------------------------------ */
/* the six *__init calls of proj/weibel/app.f90:331-347 */
int wm_create(const wm_config *cfg, wm_ctx **out);
int wm_destroy(wm_ctx *ctx);

/* Communicator for nsize > 1 (replaces MPI_COMM_WORLD of common/mpi_set.f90:25-28):
 * rank 0 calls wm_comm_unique_id, the host broadcasts the 128 bytes by any means,
 * every rank calls wm_comm_init.  NCCL send/recv + all-reduce over NVLink. */
int wm_comm_unique_id(void *id128);
int wm_comm_init(wm_ctx *ctx, const void *id128);

/* In-process transport for nsize > 1 without a second GPU: N contexts of ONE process on one device, each driven by its own
 * host thread, exchange through device-to-device copies and a host rendezvous instead of NCCL (same call sites, same
 * semantics: all ranks make the same calls in the same order).  The CG runs as the host loop here (N persistent kernels on one
 * device cannot wait for each other).  For tests of the ring path on a one-GPU box and for several slabs per GPU. */
typedef struct wm_loopback wm_loopback;
int wm_loopback_create(int32_t nsize, wm_loopback **out);
int wm_loopback_destroy(wm_loopback *group);
int wm_comm_init_loopback(wm_ctx *ctx, wm_loopback *group);

/* ---- residency: the only points where host arrays are read / written -------------- */
/* `up` need not be sorted: rows are taken as lists of np2(j,isp) records and bucket-
 * sorted on the device exactly like sort__bucket (common/sort.f90:36). */
int wm_upload_particles(wm_ctx *ctx, const double *up, const int32_t *np2);
/* `up` is already cell-sorted with a consistent cumcnt (the reference's loop invariant
 * at the top of a step): plain transposition, no sort. */
int wm_upload_particles_sorted(wm_ctx *ctx, const double *up, const int32_t *np2, const int32_t *cumcnt);
int wm_upload_field(wm_ctx *ctx, const double *uf);
/* sorted state -> up (rows, cell-sorted), np2, cumcnt; any may be NULL */
int wm_download_particles(wm_ctx *ctx, double *up, int32_t *np2, int32_t *cumcnt);
/* the post-push / post-boundary buffer (`gp` in the reference) in the slot order of the
 * sorted state; valid after wm_particle__solv and until wm_sort__bucket */
int wm_download_gp(wm_ctx *ctx, double *gp);
int wm_download_field(wm_ctx *ctx, double *uf);
int wm_download_current(wm_ctx *ctx, double *uj);  /* uj after the last deposit / bc */
int wm_download_dfield(wm_ctx *ctx, double *df);   /* df (CG warm start + last dE) */
/* per-species active particle counts on this rank */
int wm_particle_counts(wm_ctx *ctx, int64_t *n);

/* ---- the hot path, device resident, one call per reference procedure --------------- */
int wm_particle__solv(wm_ctx *ctx);        /* particle__solv  common/particle.f90:48        */
int wm_field__ele_cur(wm_ctx *ctx);        /* ele_cur only    common/field.f90:189 (testing) */
int wm_boundary__curre(wm_ctx *ctx);       /* bc__curre       common/boundary_periodic.f90:357 (testing) */
int wm_field__fdtd_i(wm_ctx *ctx);         /* field__fdtd_i   common/field.f90:66 (incl. ele_cur, bc__curre, cgm, bc__dfield) */
int wm_boundary__particle_x(wm_ctx *ctx);  /* bc__particle_x  common/boundary_periodic.f90:61 */
int wm_boundary__particle_y(wm_ctx *ctx);  /* bc__particle_y  common/boundary_periodic.f90:99 */
/* bc__injection(gp,np2,nxs,nxe,u0)  proj/shock/boundary_shock.f90:255 (WM_BC_SHOCK; on the pushed store, before
 * wm_field__fdtd_i).  Also records u0 for wm_step, like wm_set_u_inject. */
int wm_boundary__injection(wm_ctx *ctx, double u0);
int wm_set_u_inject(wm_ctx *ctx, double u0);
/* The active x range [nxs, nxe] that the following calls work on: the nxs, nxe arguments of particle__solv,
 * field__fdtd_i, sort__bucket, bc__injection ... (proj/shock/app.f90:111-116; `relocate` moves nxe, :611-621).
 * nxs must be nxgs; cells beyond nxe must hold no particles.  Default: nxgs..nxge. */
int wm_set_xrange(wm_ctx *ctx, int32_t nxs, int32_t nxe);
/* Particles the driver adds between two steps (`inject` / `relocate` of proj/shock/app.f90:611-850 append them to
 * the rows of `up`): n records (x,y,ux,uy,uz,id) of species isp (0-based), any order, inside this rank's slab. */
int wm_append_particles(wm_ctx *ctx, int32_t isp, int64_t n, const double *rec);
int wm_sort__bucket(wm_ctx *ctx);          /* sort__bucket    common/sort.f90:36             */
/* the five calls above fused: push + deposit + particle boundaries in one kernel,
 * then the field solve, then the scatter pass (proj/weibel/app.f90:100-107) */
int wm_step(wm_ctx *ctx, int32_t nsteps);

/* ---- the same, with HOST arrays (drop-in signatures; copies inside the call) -------- */
/* One full time step on host arrays: on entry `up` is cell-sorted with `cumcnt`
 * consistent (as after sort__bucket); on exit the same holds for the new state. */
int wm_host_step(wm_ctx *ctx, double *up, double *uf, int32_t *np2, int32_t *cumcnt);
/* The same over nsteps time steps with one upload before and one download after: what the Fortran shim does with
 * WM_SYNC_INTERVAL = n (host arrays refreshed every n-th sort__bucket; the sample config's intvl_mom is 50, i.e. the driver
 * looks at up/uf every 50 steps, proj/weibel/config_sample.json, proj/weibel/app.f90:109-126). */
int wm_host_steps(wm_ctx *ctx, double *up, double *uf, int32_t *np2, int32_t *cumcnt, int32_t nsteps);
/* wm_host_step moves the rows in chunks (WM_HOSTPIPE_ROWS rows, default 16): the upload of the next chunks, the particle pass
 * over the chunk that has arrived and the download of the rows that are final run side by side (PCIe is full duplex), the
 * field solve beside the last downloads.  Same kernels and results as upload + wm_step + download; WM_HOSTPIPE=0 selects that
 * sequence.  Returns the number of chunks of the last wm_host_step (0 = it ran unpipelined). */
int wm_host_pipe_chunks(const wm_ctx *ctx);
/* The order of work of the pipelined call over a slab of nyl rows in chunks of `rows` rows, as triples (kind, a, b): 0 = push
 * tile rows [a,b) (tile row = 8 rows), 1 = place the rim arrivals of tile rows [a,b), 2 = rows [a,b) are final and go back,
 * 3 = the ring exchange.  Returns the number of triples (or -needed if max_ops is too small).  No device needed. */
int wm_host_pipe_plan(int32_t nyl, int32_t rows, int32_t *ops, int32_t max_ops);
/* Page-lock a host array the calls above copy from / to (cudaHostRegister), so that the copies run at the full PCIe rate and
 * asynchronously: the reference's allocatables up, gp, uf, cumcnt (proj/weibel/app.f90:74-82) are pageable.  Once per array
 * after its allocation; registering an array twice is not an error.  Without it everything still works, the copies are
 * staged by the driver (about half the rate, and the pipelined wm_host_step cannot overlap them). */
int wm_host_register(void *ptr, size_t bytes);
int wm_host_unregister(void *ptr);
/* particle__solv(gp,up,uf,cumcnt,nxs,nxe)   common/particle.f90:48 */
int wm_host_particle__solv(wm_ctx *ctx, double *gp, const double *up, const double *uf,
                           const int32_t *cumcnt, const int32_t *np2);
/* sort__bucket(gp_out,up_in,cumcnt,np2,nxs,nxe)  common/sort.f90:36 (first argument is the output) */
int wm_host_sort__bucket(wm_ctx *ctx, double *gp_out, const double *up_in, int32_t *cumcnt,
                         const int32_t *np2);

/* ---- diagnostics ------------------------------------------------------------------- */
int wm_cg_iters(wm_ctx *ctx, int32_t out[3]);  /* CG iterations of the last solve, l=1..3 */
/* FP64 vector peak of the device in TFLOP/s, measured with a DFMA loop (2 flop each) and CUDA events: the denominator of the
 * FP64-pipe utilisation reported for the particle kernel. */
int wm_fp64_peak(wm_ctx *ctx, double *tflops);
/* Which implementation of cgm (common/field.f90:319-461) the next field solve uses: 0 = host loop of small kernels with NCCL
 * all-reduces / halo exchanges per iteration, 1 = one persistent cooperative kernel (one rank), 2 = the persistent kernel
 * with the ring exchange and the all-reduce done in the kernel over CUDA-IPC mapped peer memory.  WM_CG=0 forces 0. */
int wm_cg_path(wm_ctx *ctx, int32_t *path);
/* Block decomposition the persistent CG kernel uses for an nx x nyl slab on a device with nsm SMs and smem_max bytes of
 * shared memory per CTA: out = {blocks in x, blocks in y, shared-memory bytes, cells per thread}; error if the slab does not
 * fit (pure host logic, no device needed). */
int wm_cg_plan(int32_t nx, int32_t nyl, int32_t nsm, int64_t smem_max, int32_t out[4]);
/* energy_history (proj/weibel/app.f90:479-545), this rank's share:
 * out[0..nsp-1] kinetic, out[nsp] = sum E^2/8pi, out[nsp+1] = sum B^2/8pi */
int wm_energy(wm_ctx *ctx, double *out);
/* discrete Gauss law of the current state (north_star: "div E - rho/eps0 must hold to roundoff"; Gaussian units:
 * div E = 4 pi rho, common/field.f90:159): out[0] = max |div E - 4 pi rho| over the cells, rho with the deposit's
 * second-order shape, out[1] = max 4 pi sum |q| S S, the scale; both over this rank's cells.  Periodic boundaries; on a ring
 * every rank calls it (one exchange of the edge rows of rho per direction). */
int wm_gauss_residual(wm_ctx *ctx, double out[2]);
/* mom_calc__accl (common/mom_calc.f90:48): half-step momenta into the device's idle particle
 * store (the reference's `gp`); valid until the next call that moves or transfers particles */
int wm_mom_calc__accl(wm_ctx *ctx);
/* mom_calc__nvt (common/mom_calc.f90:167): seven moments with linear weights of the store
 * written by wm_mom_calc__accl -> host mom, ghost sums not folded yet */
int wm_mom_calc__nvt(wm_ctx *ctx, double *mom);
/* bc__mom (common/boundary_periodic.f90:571): x fold, y fold over the rank ring; host mom in/out */
int wm_boundary__mom(wm_ctx *ctx, double *mom);
/* the three calls above without the intermediate host copies (proj/weibel/app.f90:121-123) */
int wm_moments(wm_ctx *ctx, double *mom);

/* Synthetic uniform Maxwellian of proj/weibel/app.f90:380-474 generated on the device
 * with the counter-based RNG documented in DESIGN.md (same stream definition as the
 * test oracle; Box-Muller in device libm, so velocities agree to a few ulp only). */
int wm_ic_weibel(wm_ctx *ctx, uint64_t seed, int32_t n0, double vti, double vte, double t_ani, double b0);

/* ---- the applications' particle sources and initial conditions, generated on the device ------
 * The distributions are the reference's, formula by formula; the random numbers come from the counter-based generator of
 * wm_ic_weibel (the reference seeds from OS entropy and is not reproducible, utils/wuming_utils.f90:38-55). */
/* Harris current sheet of proj/reconnection/app.f90:368-456: By = b0 tanh((x-x0)/lcs) + localized perturbation e1, nbg
 * background pairs per cell between the walls + ncs*2*lcs sheet pairs per row with the sech^2 profile, drifts from jz/density.
 * lcs in cells (the app's lcs*c/wpi), vti / vte as the app defines them (sigma*sqrt(2)). */
int wm_ic_harris(wm_ctx *ctx, uint64_t seed, int32_t nbg, int32_t ncs, double lcs, double vti, double vte, double b0,
                 double rtemp, double e1);
/* Initial state of proj/shock/app.f90:406-470: n0 pairs per cell evenly spaced over the cells nxs+1..nxe-1 (the reference's
 * cumcnt, app.f90:341-344; its positions start one cell further left, see gen_kernels.cu), Maxwellian boosted by the
 * velocity profile, uniform B (b0, theta_bn, phi_bn in radians) with the motional E; sets the active range to nxs..nxe.
 * wm_config.capacity must be set (the box fills up). */
int wm_ic_shock(wm_ctx *ctx, uint64_t seed, int32_t n0, int32_t nxe, double v0, double vti, double vte, double b0,
                double theta_bn, double phi_bn, double l_damp_ini);
/* inject() (proj/shock/app.f90:685-850) and relocate() (:611-680) of the shock driver on the device, with the parameters of
 * wm_ic_shock: new particles are appended to their cells' segments, the upstream field columns refreshed; relocate also moves
 * nxe (wm_xrange tells where it is).  Call them where the driver does: after sort__bucket (proj/shock/app.f90:119-125). */
int wm_shock_inject(wm_ctx *ctx, uint64_t seed, int32_t it);
int wm_shock_relocate(wm_ctx *ctx, uint64_t seed, int32_t it);
int wm_xrange(wm_ctx *ctx, int32_t out[2]);

/* ---- measurement hooks ----------------------------------------------------------- */
/* device milliseconds (CUDA events on the context's stream) accumulated by wm_step since the
 * last reset: [0] the fused push+deposit+boundary kernel alone, [1] field solve,
 * [2] cell-centre fields + clears + migration + prefix scan, [3] scatter pass,
 * [4] whole wm_step calls (first launch to last completion); launches = kernels launched */
int wm_timing(wm_ctx *ctx, double ms[5], int64_t *launches, int32_t reset);
int wm_synchronize(wm_ctx *ctx);
/* how often wm_step had to rebuild the particle layout because a cell segment overflowed */
int wm_layout_rebuilds(wm_ctx *ctx, int64_t *n);

#ifdef __cplusplus
}
#endif
#endif
