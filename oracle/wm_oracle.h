/*
 * wm_oracle.h -- C interface of the CPU oracle (TEST INFRASTRUCTURE, not product code).
 *
 * The oracle is a line-by-line C++ restatement of the per-timestep hot path of
 * WumingPIC2D (Fortran 90 + MPI + OpenMP).  It exists only to check the CUDA
 * path: nothing under wumingpic2d_b200/ may include, link or call it.
 *
 * PARITY UNPINNED: the reference ships no golden vectors / known-answer tests
 * for this path (its only tests cover utils/iocore) and cannot be compiled in
 * this image (no Fortran compiler, no MPI).  The oracle is therefore pinned by
 * (a) an independent numpy restatement of the same Fortran (tests/np_restate.py),
 * (b) analytic invariants (discrete Gauss law, per-cell count identities,
 *     N-rank == 1-rank equivalence, energy behaviour),
 * (c) known answers of the physics the scheme must reproduce
 *     (tests/test_oracle_pins.py): Boris rotation angle, exact gather of linear
 *     fields, Esirkepov moments and continuity, the CG solution against a dense
 *     solve, the amplification factor of a vacuum wave under the implicit
 *     theta-scheme, G = (1 + i s (1 - gfac)) / (1 - i s gfac), step by step, and
 *     the frequency and amplitude of a cold Langmuir oscillation
 *     (omega = 0.994 omega_pe, e E_max = m v0 omega_pe), and the frequency of a
 *     light wave in a cold plasma (omega^2 = omega_pe^2 + c^2 k^2).
 *
 * All arrays use the reference's Fortran (column-major) layout, per rank:
 *   up,gp  (ndim=6, np, nys:nye, nsp)            proj/weibel/app.f90:75-76,281-282
 *   uf,df  (6, nxgs-2:nxge+2, nys-2:nye+2)       proj/weibel/app.f90:74,280
 *   uj     (3, nxgs-2:nxge+2, nys-2:nye+2)       common/field.f90:108
 *   gkl    (3, nxgs:nxge, nys:nye)               common/field.f90:107
 *   np2    (nys:nye, nsp)                        proj/weibel/app.f90:73,278
 *   cumcnt (nxgs:nxge+1, nys:nye, nsp)           proj/weibel/app.f90:73,279
 *   mom    (7, nxgs-1:nxge+1, nys-1:nye+1, nsp)  proj/weibel/app.f90:77,283
 * Several "ranks" (y-slabs, common/mpi_set.f90:36-47) live in one address space;
 * MPI_SENDRECV / MPI_ALLREDUCE are restated as copies / ordered sums between them.
 */
#ifndef WM_ORACLE_H
#define WM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ORC_NSP_MAX 4

/* boundary kinds (the "plugin" chosen by use-renaming in proj/<problem>/app.f90:6-13) */
enum {
  ORC_BC_PERIODIC = 0,      /* common/boundary_periodic.f90 (weibel)                                        */
  ORC_BC_RECONNECTION = 1,  /* proj/reconnection/boundary_reconnection.f90: conducting/reflecting x walls,   */
                            /* periodic y                                                                     */
  ORC_BC_SHOCK = 2          /* proj/shock/boundary_shock.f90: reflecting left wall, injection wall on the    */
                            /* right (bc__injection before the field solve), active x range [nxs, nxe]       */
};

typedef struct orc_config {
  int32_t nx, ny;      /* global grid: nxge-nxgs+1, nyge-nygs+1            */
  int32_t nxgs, nygs;  /* first global cell index (2 in every app)          */
  int32_t nranks;      /* number of y-slabs                                  */
  int32_t np;          /* row capacity (proj/weibel/app.f90:248)             */
  int32_t nsp;         /* species (2)                                        */
  int32_t bc;          /* ORC_BC_*                                           */
  double delx, delt, c, gfac;
  double q[ORC_NSP_MAX], r[ORC_NSP_MAX];
} orc_config;

enum {
  ORC_UP = 0, ORC_GP, ORC_UF, ORC_DF, ORC_UJ, ORC_GKL, ORC_MOM, /* double */
  ORC_NP2 = 16, ORC_CUMCNT                                       /* int32  */
};

typedef struct orc_world orc_world;

orc_world *orc_create(const orc_config *cfg);
void orc_destroy(orc_world *w);
/* slab bounds of `rank` (common/mpi_set.f90:36-41) */
void orc_bounds(const orc_world *w, int rank, int32_t *nys, int32_t *nye);
/* raw pointer to one of the arrays of `rank`; *len = element count */
void *orc_array(orc_world *w, int rank, int which, int64_t *len);

/* ---- stages, each over all ranks (order of one step: proj/weibel/app.f90:100-107) ---- */
void orc_particle_solv(orc_world *w);   /* gp <- push(up, uf, cumcnt)   common/particle.f90:48-177 */
void orc_ele_cur(orc_world *w);         /* uj <- deposit(up, gp)        common/field.f90:189-316   */
void orc_bc_curre(orc_world *w);        /* uj fold + halo               boundary_periodic.f90:357  */
int  orc_field_fdtd_i(orc_world *w);    /* full field solve; 0 ok, 1 = CG hit ite_max (field.f90:427) */
void orc_bc_particle_x(orc_world *w);   /* on gp                        boundary_periodic.f90:61   */
void orc_bc_injection(orc_world *w, double u0); /* on gp            proj/shock/boundary_shock.f90:255  */
void orc_set_u_inject(orc_world *w, double u0); /* u0 used by orc_step (proj/shock/app.f90:113)           */
int  orc_set_xrange(orc_world *w, int nxs, int nxe); /* active x range of the shock app; 0 = ok         */
int  orc_bc_particle_y(orc_world *w);   /* on gp; 1 = row overflow      boundary_periodic.f90:99   */
void orc_sort_bucket(orc_world *w);     /* up <- sort(gp), cumcnt       common/sort.f90:36         */
int  orc_step(orc_world *w, int nsteps);/* the 5 calls above in order                               */
/* wall-clock seconds accumulated by orc_step per stage: push, field(+deposit), bc_x, bc_y, sort */
void orc_stage_times(orc_world *w, double out[5], int reset);
void orc_mom_accl(orc_world *w);        /* gp <- half-step accel of up  common/mom_calc.f90:48     */
void orc_mom_nvt(orc_world *w);         /* mom <- moments of gp         common/mom_calc.f90:167    */
void orc_bc_mom(orc_world *w);          /* boundary_periodic.f90:571                                */

/* CG iteration counts of the last field solve, per component l=1..3 */
void orc_cg_iters(const orc_world *w, int32_t out[3]);

/* energy_history (proj/weibel/app.f90:479-545): out[0..nsp-1] kinetic per species,
 * out[nsp] = E^2/8pi, out[nsp+1] = B^2/8pi (global sums over ranks) */
void orc_energy(orc_world *w, double *out);

/* Gauss residual: max over interior cells of |div E - 4 pi rho| using `up`
 * (sorted state) and uf; rho with the 2nd-order shape about the cell.
 * Also returns max |4 pi rho| for scale in *scale. */
double orc_gauss_residual(orc_world *w, double *scale);

/* uniform Maxwellian IC of proj/weibel/app.f90:380-474 with a counter-based RNG
 * keyed by (seed, species, global particle id): identical for any nranks.
 * Sets up, gp(=up), np2, cumcnt, uf (Bz=b0), ids = -(global id). */
void orc_ic_weibel(orc_world *w, uint64_t seed, int n0, double vti, double vte,
                   double t_ani, double b0);

/* the RNG itself (for cross-checks against the device generator) */
uint64_t orc_rng_hash(uint64_t seed, int isp, uint64_t gid, int stream);
double orc_rng_uniform(uint64_t seed, int isp, uint64_t gid, int stream);

/* number of OpenMP threads the oracle will use */
int orc_num_threads(void);
void orc_set_num_threads(int n);

#ifdef __cplusplus
}
#endif
#endif
