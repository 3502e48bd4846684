// kernels.h -- launch wrappers of the device kernels (particle_kernels.cu, field_kernels.cu)
#pragma once
#include "wm_internal.h"

namespace wm {

enum { M_PUSH = 1, M_DEPOSIT = 2, M_BOUND = 4, M_EXACT = 8, M_NOMOVE = 16 };

// ---- particles
void launch_pass1(int mode, const DevParams &P, const Pass1Args &a, cudaStream_t st);
// fused push + deposit + particle boundaries + histogram, FMA arithmetic (fused_kernel.cu)
void launch_fused(const DevParams &P, const Pass1Args &a, cudaStream_t st);
void launch_pass2(const DevParams &P, const PartSoA &src, const PartSoA &dst, const int *cstart_old,
                  const int *cstart_new, const int *tilebase, const uint32_t *tag, PView<double> keyx, unsigned *err,
                  cudaStream_t st);
int scan_scratch_ints(int n);
// exclusive scan of cell_capacity(in[c], sl): sl = 0 gives the tight offsets (cumcnt), sl > 0 the segment offsets
// floor_n: with slack, a cell is laid out as if it held at least floor_n particles (cells a moving source will fill)
int launch_scan(const int *in, int *out, int *scratch, int n, float sl, cudaStream_t st, int floor_n = 0);
void launch_incoming_tag(const DevParams &P, const double *rec, int n, int isp, const int *spv, int *gcnt, int *rank,
                         unsigned *err, cudaStream_t st);
void launch_incoming_scatter(const DevParams &P, const double *rec, int n, int isp, const int *spv, const int *cstart_new,
                             const int *rank, const PartSoA &dst, unsigned *err, cudaStream_t st);
void launch_aos2soa(const double *rec, long long n, size_t so, const PartSoA &dst, cudaStream_t st);
void launch_soa2aos(const PartSoA &src, size_t so, long long n, double *rec, cudaStream_t st);
void launch_bcx(const DevParams &P, const PartSoA &g, const int *cstart, bool injection, cudaStream_t st);
void launch_ic_weibel(const DevParams &P, const PartSoA &dst, int *cstart, int *cnt, uint64_t seed, int n0, double vti,
                      double vte, double t_ani, float sl, cudaStream_t st);
void launch_relayout_from_aos(const DevParams &P, const double *rec, long long n, const int *tight, const int *cstart_dst,
                              const PartSoA &dst, size_t so, unsigned *err, cudaStream_t st);
void launch_relayout_to_aos(const DevParams &P, const PartSoA &key, const PartSoA &val, size_t so, const int *cstart_src,
                            const int *tight, double *rec, cudaStream_t st);
void launch_relayout_soa(const DevParams &P, const PartSoA &src, const int *cstart_src, const int *cnt_src, const PartSoA &dst,
                         const int *cstart_dst, unsigned *err, cudaStream_t st);
void launch_incoming_append(const DevParams &P, const double *rec, int n, int isp, const int *cstart, int *cnt_tail,
                            const PartSoA &dst, double *ovf, int *ovfsp, int *ovfcnt, int ovfcap, unsigned *err,
                            cudaStream_t st, const int *n_dev = nullptr, const int *cntb = nullptr);
void launch_check_counts(const int *cnt, int n, const int lim[4], unsigned *err, cudaStream_t st);
// in-place sort: append the records staged by k_fused<INPLACE> to their new segments, retire vacated slots;
// rim_only: k_fused_sm<TAIL> served the tile's own cells, only the window rim is left (k_place_rim)
void launch_place(const DevParams &P, const double *stage, const uint32_t *tag, const PartSoA &dst, const int *cstart,
                  int *cnt_new, const int *tilebase, double *ovf, int *ovfsp, int *ovfcnt, int ovfcap, unsigned *err,
                  bool rim_only, cudaStream_t st, int tile0 = 0, int ntiles = 0);
// does launch_fused_sm(variant) place the in-tile cell changers itself?
bool fused_sm_has_tail(int variant);
void launch_clamp_counts(const DevParams &P, const int *cstart, int *cnt, cudaStream_t st, const int *cntb = nullptr);
// pipelined host step (hostpipe_kernels.cu): the reference's cumcnt on the device, layout changes by row range
void launch_cum_to_counts(const DevParams &P, const int *cum, const int *np2, int *cnt, unsigned *err, cudaStream_t st);
void launch_rows_from_aos(const DevParams &P, const double *rec, int n, int isp, int ra, int rb, const int *rowoff, int base, const int *cum,
                          const int *cstart, const PartSoA &dst, unsigned *err, cudaStream_t st);
void launch_mark_gaps_rows(const DevParams &P, PView<double> x, const int *cstart, const int *cnt, int ra, int rb, cudaStream_t st);
void launch_rows_cumcnt(const DevParams &P, const int *cnt, int ra, int rb, int *cum, int *rowoff, cudaStream_t st);
void launch_rows_to_aos(const DevParams &P, const PartSoA &src, int ra, int rb, const int *cstart, const int *cnt, const int *cum,
                        const int *rowoff, double *rec, int reccap, unsigned *err, cudaStream_t st);
void launch_mark_gaps(const DevParams &P, PView<double> x, const int *cstart, const int *cnt, cudaStream_t st);
void launch_nbr_max(const DevParams &P, const int *in, int *out, int r, cudaStream_t st);
void launch_mark_dead(const DevParams &P, PView<double> x, const int *cstart, const int *cnt_old, int *cnt_new, cudaStream_t st);
// fused push + deposit + boundaries that moves cell changers itself (no tags, no scatter pass)
void launch_fused_inplace(const DevParams &P, const Pass1Args &a, cudaStream_t st);
// k_fused<INPLACE> with the deposit split into stayers (21 sums, in the loop) and movers (queued, drained per cell) (fused5_kernel.cu)
void launch_fused_sm(const DevParams &P, const Pass1Args &a, int variant, cudaStream_t st, int ntiles = 0);
// push + deposit + boundaries + sort with direct placement of the cell changers: reads a.src, writes a.dst (fused6_kernel.cu)
void launch_fused_dp(const DevParams &P, const Pass1Args &a, cudaStream_t st);
void launch_place_rim2(const DevParams &P, const uint32_t *tag, const PartSoA &dst, const int *cstart, int *cnt_new, const int *cntb_new,
                       double *ovf, int *ovfsp, int *ovfcnt, int ovfcap, unsigned *err, cudaStream_t st);
void launch_normalize(const DevParams &P, const PartSoA &st_, const int *cstart, int *cnt, int *cntb, cudaStream_t st);
void launch_kinetic(const DevParams &P, const PartSoA &src, const int *cstart, int isp, double *partial, int nblocks,
                    cudaStream_t st);
// diagnostic: discrete Gauss law residual of the sorted store against uf (periodic x).  rho: 2 planes of (nyl + 2) rows; the
// rows 0 and nyl + 1 belong to the ring neighbours and are folded by the host between the two launches
void launch_charge_density(const DevParams &P, const PartSoA &src, const int *cstart, double *rho, cudaStream_t st);
void launch_gauss(const DevParams &P, const double *uf, const double *rho, unsigned long long *out, cudaStream_t st);
void launch_moments(const DevParams &P, const PartSoA &src, PView<double> keyx, const int *cstart, double *mom,
                    cudaStream_t st);

// ---- generators of the applications' particle sources and initial fields (gen_kernels.cu)
struct GenParams {
  int nxgs, nygs, nxg, nyg;        // global grid: first cell, cells
  double delx, delt, c;
  double q[WM_NSP_MAX];
  double vti, vte, b0;
  // Harris sheet (proj/reconnection)
  int nbg, ncs;
  long long npr, ibg;              // pairs per row, of which background
  double lcs, rtemp, e1;
  // shock (proj/shock)
  int n0, it;
  double v0, theta, phi, l_damp;
};
void launch_gen_harris(const DevParams &P, const GenParams &g, uint64_t seed, double *stage0, double *stage1, double *uf,
                       cudaStream_t st);
void launch_gen_shock(const DevParams &P, const GenParams &g, uint64_t seed, int kind, int nxe, const int *rowoff, double *stage0,
                      double *stage1, cudaStream_t st);
void launch_field_shock(const DevParams &P, const GenParams &g, int kind, int nxe, double *uf, cudaStream_t st);

// ---- fields
struct FieldBufs {
  double *uf, *df, *tmpf;     // AoS6 padded
  double *uj, *gkl;           // AoS3 padded
  double *phi, *p, *r, *ap;   // CG vectors, AoS3 padded (b lives in gkl, scaled by f5)
  double *p2;                 // second p buffer of the two-kernel CG iteration
  double *red;                // reduction scratch: see field_kernels.cu
  int *cgstate;               // device CG control block
};
constexpr int RED_BLOCKS = 592;       // 148 SMs x 4
constexpr int RED_BLOCKS_MAX = 2368;  // size of the partial-sum scratch

void launch_tmpf(const DevParams &P, const double *uf, double *tmpf, cudaStream_t st);
void launch_fill_x(const DevParams &P, double *a, int ncomp, int ng, cudaStream_t st);           // periodic x ghosts (copy)
void launch_fill_y_local(const DevParams &P, double *a, int ncomp, int ng, cudaStream_t st);     // periodic y ghosts, nsize==1
void launch_fold_x(const DevParams &P, double *uj, cudaStream_t st);                              // uj x fold + copy back
void launch_fold_y_local(const DevParams &P, double *uj, cudaStream_t st);                        // uj y fold + refresh, nsize==1
void launch_mom_fold_x(const DevParams &P, double *mom, cudaStream_t st);                        // mom x fold
void launch_mom_fold_y_local(const DevParams &P, double *mom, cudaStream_t st);                  // mom y fold, nsize==1
void launch_add_rows(double *dst, const double *src, long long n, cudaStream_t st);
void launch_rhs(const DevParams &P, const FieldBufs &f, cudaStream_t st);
void launch_cg_init(const DevParams &P, const FieldBufs &f, cudaStream_t st);      // phi<-df, b, sum b^2
void launch_cg_resid0(const DevParams &P, const FieldBufs &f, cudaStream_t st);    // r, p, sum r^2
void launch_cg_begin(const DevParams &P, const FieldBufs &f, int nranks_reduced, cudaStream_t st);
void launch_cg_ap(const DevParams &P, const FieldBufs &f, cudaStream_t st);        // ap, sums
void launch_cg_update(const DevParams &P, const FieldBufs &f, cudaStream_t st);    // phi, r, sum r^2
void launch_cg_pupdate(const DevParams &P, const FieldBufs &f, cudaStream_t st);   // p
// two-kernel iteration: p update folded into A p (p double-buffered), loop control in the update kernel
void launch_cg_pap(const DevParams &P, const FieldBufs &f, const double *p_in, double *p_out, cudaStream_t st);
void launch_cg_update2(const DevParams &P, const FieldBufs &f, const double *p, cudaStream_t st);
void launch_cg_finish(const DevParams &P, const FieldBufs &f, cudaStream_t st);    // df(1:3) <- phi
void launch_efield(const DevParams &P, const FieldBufs &f, cudaStream_t st);       // df(4:6)
void launch_update_uf(const DevParams &P, const FieldBufs &f, cudaStream_t st);    // uf += df
void launch_field_energy(const DevParams &P, const double *uf, double *partial, int nblocks, cudaStream_t st);
// ---- persistent cooperative CG (cg_persist_kernel.cu): the three solves of cgm in one kernel, state on chip
constexpr int CGP_T = 1024;   // threads per CTA, one CTA per SM
constexpr int CGP_K = 16;     // cells per thread at most (r in registers): 4096 x 512 on 148 SMs takes 14, 16384 x 128 takes 15
constexpr int CGP_KM = 14, CGP_KS = 8;  // further instantiations for slabs that need at most 14 / 8 cells per thread
constexpr int CGP_MAXR = 8;   // ranks on the ring the in-kernel all-reduce supports
struct CgpShared {            // one per rank, mapped by every other rank (CUDA IPC)
  unsigned long long ll[4][CGP_MAXR][4];   // [sequence & 3][src][2 x value]: (half of a double | sequence << 32) words
  unsigned long long halo_flag[2][160];    // [0]: ndown's CTA bx has stored its part of my row nys-1 for barrier (value);
                                           // [1]: nup's CTA bx ... of my row nye+1
};
struct CgpArgs {
  int cbx, cby;               // block decomposition of the slab: cbx x cby CTAs
  int rl;                     // cells per thread (a vertical run)
  int cp;                     // column pitch of the shared-memory tile: (rows of the tallest block + 2) | 1
  double *df;                 // in: df(1:3) warm start (+ ghost rows of the ring neighbours); out: df(1:3) interior
  const double *gkl;          // right-hand side before the f5 scaling
  double *rg;                 // AoS3 padded: r of the blocks' perimeter cells, ghost rows written by the ring neighbours
  double *phipl, *bpl;        // 3 dense planes each: phi and b, cell e of CTA c at plane[l] + base(c) + e
  double *partial;            // [2][160] x (2 doubles): partial sums of the CTAs
  unsigned *bar;              // barrier counter, zero at launch
  int *abort;                 // set by a CTA whose wait timed out: everybody leaves
  int *out;                   // ite[3], stop, barriers
  unsigned *err;
  int nrank, nsize;
  unsigned long long seq0;    // barrier sequence base of this solve (same on every rank)
  CgpShared *sh[CGP_MAXR];    // every rank's block as mapped here (sh[nrank] = mine)
  double *r_up, *r_down;      // rg of nup / ndown as mapped here
  int nyl_down;               // rows of ndown: its upper ghost row is local row nyl_down
  int nup, ndown;             // ring neighbours
  int cbx_up, cbx_down;       // blocks in x of nup's / ndown's decomposition (one halo flag each)
  unsigned long long *trace;  // WM_CGTRACE: [G][4 barriers][4 events] global-timer stamps, else null
  unsigned trace_seq;
};
bool cgp_plan(int nx, int nyl, int nsm, size_t smem_max, int *cbx, int *cby, int *rl, size_t *smem);
cudaError_t cgp_prepare(size_t smem);
cudaError_t launch_cg_persist(const DevParams &P, const CgpArgs &a, size_t smem, cudaStream_t st);
struct RankPtrs { void *p[CGP_MAXR]; };
// out[k] = sum (doubles) or min (ints) over the ranks' p[q][k], in rank order (loopback all-reduce)
void launch_sum_ranks(const RankPtrs &rp, int nranks, int n, bool is_double, double *out, cudaStream_t st);
void launch_fp64_peak(double *out, int nblocks, int n, cudaStream_t st);  // 8 x n DFMA per thread, 512 threads per block
size_t cgctl_bytes();
size_t cgctl_active_offset();       // int active[3]; int ite[3]; int stop  (contiguous)
size_t cgctl_sums_offset(int which); // 0 sumb, 1 sumr(+sum2 adjacent), 2 sum2, 3 sum1

}  // namespace wm
