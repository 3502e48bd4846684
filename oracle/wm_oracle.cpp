// wm_oracle.cpp -- CPU oracle for the WumingPIC2D per-timestep hot path.
//
// TEST INFRASTRUCTURE ONLY.  Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs may load this; the product path
// (wumingpic2d_b200/) never does.
//
// PARITY UNPINNED: the reference has no golden vectors for this path and cannot be
// built here (no Fortran compiler / MPI).  See wm_oracle.h for what pins it instead.
//
// Every routine restates one reference routine literally: same loop structure,
// same operation order inside each expression, int() truncation, dsqrt, separate
// 1./gam, round-toward -inf wrap adds, and the CG loop exactly as written
// (including its convergence-test quirks).  Citations are file:line relative to
// the reference tree.  Parity build: -O2 -fno-fast-math -ffp-contract=off
// -frounding-math (see Makefile); the timing build uses -O3 -march=native.
//
// Deviations from the Fortran, all confined to *ordering that the reference itself
// leaves unspecified*:
//  * OpenMP array REDUCTION(+:uj) (field.f90:209-211) -> 5-colour row phases so the
//    sum order is fixed and independent of the thread count.
//  * OpenMP scalar reductions -> per-row partial sums added in row order, then in
//    rank order (MPI_SUM order is unspecified too).
//  * omp_lock-ordered appends in particle_y (boundary_periodic.f90:156-161) ->
//    fixed order (from row j-1 first, then from row j+1).
//  * MPI_SENDRECV between ranks -> copies between slabs in one address space.
#include "wm_oracle.h"

#include <cfenv>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

namespace {

constexpr int NDIM = 6;
const double PI = 4.0 * std::atan(1.0);  // field.f90:13, app.f90:68

struct Rank {
  int nys, nye, nyl, nup, ndown;
  std::vector<double> up, gp, uf, df, uj, gkl, mom;
  std::vector<int> np2, cumcnt;
  // cgm automatic arrays (field.f90:342-344)
  std::vector<double> phi, p, r, b, ap;
  // particle_y scratch (boundary_periodic.f90:107-111)
  std::vector<int> flag, cnt2;
  std::vector<std::vector<double>> bff;  // rows nys-1..nye+1
};

}  // namespace

struct orc_world {
  orc_config c;
  int nx, ny, nxgs, nxge, nygs, nyge, nxs, nxe, np, nsp;
  double delx, delt, cc, gfac, d_delx, d_delt;
  double f1, f2, f3, f4, f5;  // field.f90:53-57
  std::vector<Rank> R;
  int cg_ite[3];
  double t_stage[5];
  double u_inject = 0.0;  // u0 of bc__injection (proj/shock/app.f90:113)
};

namespace {

// ---------------------------------------------------------------- indexing
struct V {  // Fortran-layout views of one rank
  const orc_world *w;
  const Rank *k;
  inline size_t UP(int idim, int ii, int j, int isp) const {  // 1-based idim, ii, isp
    return (size_t)(idim - 1) + NDIM * ((size_t)(ii - 1) + (size_t)w->np * ((size_t)(j - k->nys) + (size_t)k->nyl * (isp - 1)));
  }
  inline size_t F6(int c, int i, int j) const {  // uf/df (6, nxgs-2:nxge+2, nys-2:nye+2)
    return (size_t)(c - 1) + 6 * ((size_t)(i - (w->nxgs - 2)) + (size_t)(w->nx + 4) * (j - (k->nys - 2)));
  }
  inline size_t J3(int c, int i, int j) const {  // uj (3, nxgs-2:nxge+2, nys-2:nye+2)
    return (size_t)(c - 1) + 3 * ((size_t)(i - (w->nxgs - 2)) + (size_t)(w->nx + 4) * (j - (k->nys - 2)));
  }
  inline size_t G3(int c, int i, int j) const {  // gkl (3, nxgs:nxge, nys:nye)
    return (size_t)(c - 1) + 3 * ((size_t)(i - w->nxgs) + (size_t)w->nx * (j - k->nys));
  }
  inline size_t CUM(int i, int j, int isp) const {  // cumcnt (nxgs:nxge+1, nys:nye, nsp)
    return (size_t)(i - w->nxgs) + (size_t)(w->nx + 1) * ((size_t)(j - k->nys) + (size_t)k->nyl * (isp - 1));
  }
  inline size_t NP2(int j, int isp) const { return (size_t)(j - k->nys) + (size_t)k->nyl * (isp - 1); }
  inline size_t PH(int i, int j) const {  // phi,p (nxs-1:nxe+1, nys-1:nye+1)
    return (size_t)(i - (w->nxs - 1)) + (size_t)(w->nx + 2) * (j - (k->nys - 1));
  }
  inline size_t RI(int i, int j) const {  // r,b,ap (nxs:nxe, nys:nye)
    return (size_t)(i - w->nxs) + (size_t)w->nx * (j - k->nys);
  }
  inline size_t MOM(int c, int i, int j, int isp) const {  // (7, nxgs-1:nxge+1, nys-1:nye+1, nsp)
    return (size_t)(c - 1) + 7 * ((size_t)(i - (w->nxgs - 1)) + (size_t)(w->nx + 2) * ((size_t)(j - (k->nys - 1)) + (size_t)(k->nyl + 2) * (isp - 1)));
  }
};

inline double now() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

// ---------------------------------------------------------------- push
// common/particle.f90:48-177 (dt_scale=1, move=true) and common/mom_calc.f90:48-164
// (delt*0.5 at mom_calc.f90:34, no move, positions copied :151-152).
void push_rank(const orc_world *w, Rank &k, double delt, bool move) {
  V v{w, &k};
  const int nxs = w->nxs, nxe = w->nxe, nys = k.nys, nye = k.nye;
  const double c = w->cc, d_delx = w->d_delx;
  const double *up = k.up.data();
  const double *uf = k.uf.data();
  double *gp = k.gp.data();
  // tmp(1:6, nxs-1:nxe+1, nys-1:nye+1)                          particle.f90:61
  const int tnx = w->nx + 2;
  std::vector<double> tmpv((size_t)6 * tnx * (k.nyl + 2));
  double *tmp = tmpv.data();
  auto T = [&](int cidx, int i, int j) -> size_t {
    return (size_t)(cidx - 1) + 6 * ((size_t)(i - (nxs - 1)) + (size_t)tnx * (j - (nys - 1)));
  };
  // fields at (i+1/2, j+1/2)                                     particle.f90:69-81
#pragma omp parallel for
  for (int j = nys - 1; j <= nye + 1; j++)
    for (int i = nxs - 1; i <= nxe + 1; i++) {
      tmp[T(1, i, j)] = 0.5 * (+uf[v.F6(1, i, j)] + uf[v.F6(1, i, j + 1)]);
      tmp[T(2, i, j)] = 0.5 * (+uf[v.F6(2, i, j)] + uf[v.F6(2, i + 1, j)]);
      tmp[T(3, i, j)] = 0.25 * (+uf[v.F6(3, i, j)] + uf[v.F6(3, i + 1, j)] + uf[v.F6(3, i, j + 1)] + uf[v.F6(3, i + 1, j + 1)]);
      tmp[T(4, i, j)] = 0.5 * (+uf[v.F6(4, i, j)] + uf[v.F6(4, i + 1, j)]);
      tmp[T(5, i, j)] = 0.5 * (+uf[v.F6(5, i, j)] + uf[v.F6(5, i, j + 1)]);
      tmp[T(6, i, j)] = uf[v.F6(6, i, j)];
    }

#pragma omp parallel for schedule(static)
  for (int j = nys; j <= nye; j++)
    for (int i = nxs; i <= nxe; i++)
      for (int isp = 1; isp <= w->nsp; isp++) {
        const double fac1 = w->c.q[isp - 1] / w->c.r[isp - 1] * 0.5 * delt;  // particle.f90:90
        const double txxx = fac1 * fac1;                                       // :91
        const double fac2 = w->c.q[isp - 1] * delt / w->c.r[isp - 1];          // :92
        const int lo = k.cumcnt[v.CUM(i, j, isp)] + 1, hi = k.cumcnt[v.CUM(i + 1, j, isp)];
        for (int ii = lo; ii <= hi; ii++) {
          const double *pu = up + v.UP(1, ii, j, isp);
          double *pg = gp + v.UP(1, ii, j, isp);
          double sh[3][2];  // sh(-1:1, 1:2)
          // second order shape function                          particle.f90:97-105
          double dh = pu[0] * d_delx - 0.5 - i;
          sh[0][0] = 0.5 * (0.5 - dh) * (0.5 - dh);
          sh[1][0] = 0.75 - dh * dh;
          sh[2][0] = 0.5 * (0.5 + dh) * (0.5 + dh);
          dh = pu[1] * d_delx - 0.5 - j;
          sh[0][1] = 0.5 * (0.5 - dh) * (0.5 - dh);
          sh[1][1] = 0.75 - dh * dh;
          sh[2][1] = 0.5 * (0.5 + dh) * (0.5 + dh);
          double f[6];
          for (int cidx = 1; cidx <= 6; cidx++) {  // particle.f90:107-129
            f[cidx - 1] =
                +(+tmp[T(cidx, i - 1, j - 1)] * sh[0][0] + tmp[T(cidx, i, j - 1)] * sh[1][0] + tmp[T(cidx, i + 1, j - 1)] * sh[2][0]) * sh[0][1]
                + (+tmp[T(cidx, i - 1, j)] * sh[0][0] + tmp[T(cidx, i, j)] * sh[1][0] + tmp[T(cidx, i + 1, j)] * sh[2][0]) * sh[1][1]
                + (+tmp[T(cidx, i - 1, j + 1)] * sh[0][0] + tmp[T(cidx, i, j + 1)] * sh[1][0] + tmp[T(cidx, i + 1, j + 1)] * sh[2][0]) * sh[2][1];
          }
          const double bpx = f[0], bpy = f[1], bpz = f[2], epx = f[3], epy = f[4], epz = f[5];
          // accel.                                               particle.f90:132-134
          double uvm1 = pu[2] + fac1 * epx;
          double uvm2 = pu[3] + fac1 * epy;
          double uvm3 = pu[4] + fac1 * epz;
          // rotate                                               particle.f90:137-148
          double gam = std::sqrt(c * c + uvm1 * uvm1 + uvm2 * uvm2 + uvm3 * uvm3);
          const double igam = 1. / gam;
          const double fac1r = fac1 * igam;
          const double fac2r = fac2 / (gam + txxx * (bpx * bpx + bpy * bpy + bpz * bpz) * igam);
          const double uvm4 = uvm1 + fac1r * (+uvm2 * bpz - uvm3 * bpy);
          const double uvm5 = uvm2 + fac1r * (+uvm3 * bpx - uvm1 * bpz);
          const double uvm6 = uvm3 + fac1r * (+uvm1 * bpy - uvm2 * bpx);
          uvm1 = uvm1 + fac2r * (+uvm5 * bpz - uvm6 * bpy);
          uvm2 = uvm2 + fac2r * (+uvm6 * bpx - uvm4 * bpz);
          uvm3 = uvm3 + fac2r * (+uvm4 * bpy - uvm5 * bpx);
          // accel.                                               particle.f90:151-153
          pg[2] = uvm1 + fac1 * epx;
          pg[3] = uvm2 + fac1 * epy;
          pg[4] = uvm3 + fac1 * epz;
          if (move) {
            // move                                               particle.f90:156-161
            gam = 1. / std::sqrt(1.0 + (+pg[2] * pg[2] + pg[3] * pg[3] + pg[4] * pg[4]) / (c * c));
            pg[0] = pu[0] + pg[2] * delt * gam;
            pg[1] = pu[1] + pg[3] * delt * gam;
          } else {
            pg[0] = pu[0];  // mom_calc.f90:151-152
            pg[1] = pu[1];
          }
        }
      }
  if (move) {
    // gp(6,:,:,:) = up(6,:,:,:) over the full padded capacity     particle.f90:171-175
    const size_t n = (size_t)w->np * k.nyl * w->nsp;
#pragma omp parallel for
    for (size_t s = 0; s < n; s++) gp[s * NDIM + 5] = up[s * NDIM + 5];
  }
}

// ---------------------------------------------------------------- deposit
// common/field.f90:189-316 (ele_cur).  One cell's 5x5x3 block, both species.
inline void ele_cur_cell(const orc_world *w, const Rank &k, const V &v, int i, int j, double *uj) {
  const double fac = 1.0 / 3.0;  // field.f90:198
  const double d_delx = w->d_delx, d_delt = w->d_delt, delx = w->delx, c = w->cc;
  const double *up = k.up.data();
  const double *gp = k.gp.data();
  double pjx[5][5], pjy[5][5], pjz[5][5], pjtmp[5][5];  // [jp+2][ip+2]
  for (int a = 0; a < 5; a++)
    for (int b = 0; b < 5; b++) pjx[a][b] = pjy[a][b] = pjz[a][b] = 0.0;
  for (int isp = 1; isp <= w->nsp; isp++) {
    const double q = w->c.q[isp - 1];
    const int lo = k.cumcnt[v.CUM(i, j, isp)] + 1, hi = k.cumcnt[v.CUM(i + 1, j, isp)];
    for (int ii = lo; ii <= hi; ii++) {
      const double *pu = up + v.UP(1, ii, j, isp);
      const double *pg = gp + v.UP(1, ii, j, isp);
      double s0[5][2], ds[5][2];  // (-2:2, 1:2)
      // field.f90:224-236
      double dh = pu[0] * d_delx - 0.5 - i;
      s0[0][0] = 0.0;
      s0[1][0] = 0.5 * (0.5 - dh) * (0.5 - dh);
      s0[2][0] = 0.75 - dh * dh;
      s0[3][0] = 0.5 * (0.5 + dh) * (0.5 + dh);
      s0[4][0] = 0.0;
      dh = pu[1] * d_delx - 0.5 - j;
      s0[0][1] = 0.0;
      s0[1][1] = 0.5 * (0.5 - dh) * (0.5 - dh);
      s0[2][1] = 0.75 - dh * dh;
      s0[3][1] = 0.5 * (0.5 + dh) * (0.5 + dh);
      s0[4][1] = 0.0;
      for (int ax = 0; ax < 2; ax++) {  // field.f90:238-266
        const int i2 = (int)(pg[ax] * d_delx);
        dh = pg[ax] * d_delx - 0.5 - i2;
        const int inc = i2 - (ax == 0 ? i : j);
        const double s1_1 = 0.5 * (0.5 - dh) * (0.5 - dh);
        const double s1_2 = 0.75 - dh * dh;
        const double s1_3 = 0.5 * (0.5 + dh) * (0.5 + dh);
        const double smo_1 = -(inc - std::abs(inc)) * 0.5 + 0;
        const double smo_2 = -std::abs(inc) + 1;
        const double smo_3 = (inc + std::abs(inc)) * 0.5 + 0;
        ds[0][ax] = s1_1 * smo_1;
        ds[1][ax] = s1_1 * smo_2 + s1_2 * smo_1;
        ds[2][ax] = s1_2 * smo_2 + s1_3 * smo_1 + s1_1 * smo_3;
        ds[3][ax] = s1_3 * smo_2 + s1_2 * smo_3;
        ds[4][ax] = s1_3 * smo_3;
      }
      for (int ax = 0; ax < 2; ax++)
        for (int a = 0; a < 5; a++) ds[a][ax] = ds[a][ax] - s0[a][ax];  // field.f90:268
      // field.f90:270-272
      const double gvz = pg[4] / std::sqrt(1. + (+pg[2] * pg[2] + pg[3] * pg[3] + pg[4] * pg[4]) / (c * c));
      // field.f90:274-281
      for (int a = 0; a < 5; a++)
        for (int b = 0; b < 5; b++) pjtmp[a][b] = 0.0;
      for (int jp = 0; jp < 5; jp++)
        for (int ip = 0; ip < 4; ip++)
          pjtmp[jp][ip + 1] = pjtmp[jp][ip] - q * delx * d_delt * ds[ip][0] * (s0[jp][1] + 0.5 * ds[jp][1]);
      for (int a = 0; a < 5; a++)
        for (int b = 0; b < 5; b++) pjx[a][b] = pjx[a][b] + pjtmp[a][b];
      // field.f90:283-290
      for (int a = 0; a < 5; a++)
        for (int b = 0; b < 5; b++) pjtmp[a][b] = 0.0;
      for (int jp = 0; jp < 4; jp++)
        for (int ip = 0; ip < 5; ip++)
          pjtmp[jp + 1][ip] = pjtmp[jp][ip] - q * delx * d_delt * ds[jp][1] * (s0[ip][0] + 0.5 * ds[ip][0]);
      for (int a = 0; a < 5; a++)
        for (int b = 0; b < 5; b++) pjy[a][b] = pjy[a][b] + pjtmp[a][b];
      // field.f90:292-298
      for (int jp = 0; jp < 5; jp++)
        for (int ip = 0; ip < 5; ip++)
          pjz[jp][ip] = pjz[jp][ip] + q * gvz * (+s0[ip][0] * s0[jp][1] + 0.5 * ds[ip][0] * s0[jp][1] + 0.5 * s0[ip][0] * ds[jp][1] + fac * ds[ip][0] * ds[jp][1]);
    }
  }
  // field.f90:304-310
  for (int jp = -2; jp <= 2; jp++)
    for (int ip = -2; ip <= 2; ip++) {
      uj[v.J3(1, i + ip, j + jp)] = uj[v.J3(1, i + ip, j + jp)] + pjx[jp + 2][ip + 2];
      uj[v.J3(2, i + ip, j + jp)] = uj[v.J3(2, i + ip, j + jp)] + pjy[jp + 2][ip + 2];
      uj[v.J3(3, i + ip, j + jp)] = uj[v.J3(3, i + ip, j + jp)] + pjz[jp + 2][ip + 2];
    }
}

void ele_cur_rank(const orc_world *w, Rank &k) {
  V v{w, &k};
  double *uj = k.uj.data();
  // uj(1:3,nxs-2:nxe+2,nys-2:nye+2) = 0                          field.f90:203-205
  std::fill(k.uj.begin(), k.uj.end(), 0.0);
  // rows j, j+5, j+10, ... touch disjoint uj rows (stencil +-2): five ordered phases
  for (int phase = 0; phase < 5; phase++) {
#pragma omp parallel for schedule(static)
    for (int j = k.nys + phase; j <= k.nye; j += 5)
      for (int i = w->nxs; i <= w->nxe; i++) ele_cur_cell(w, k, v, i, j, uj);
  }
}

// ---------------------------------------------------------------- boundaries (periodic)
// boundary_periodic.f90:357-508
void bc_curre(orc_world *w) {
  const int nxs = w->nxs, nxe = w->nxe;
  const int n = 6 * (nxe - nxs + 4 + 1);
  const int nr = (int)w->R.size();
  std::vector<std::vector<double>> snd(nr, std::vector<double>(n));
  auto pack = [&](int rows_of /*0: nys-2,nys-1  1: nye+1,nye+2  2: nys,nys+1  3: nye-1,nye*/) {
    for (int rk = 0; rk < nr; rk++) {
      Rank &k = w->R[rk];
      V v{w, &k};
      const int ja = rows_of == 0 ? k.nys - 2 : rows_of == 1 ? k.nye + 1 : rows_of == 2 ? k.nys : k.nye - 1;
      for (int i = nxs - 2; i <= nxe + 2; i++) {
        const int ii = 6 * (i - (nxs - 2));
        for (int cidx = 1; cidx <= 3; cidx++) {
          snd[rk][ii + cidx - 1] = k.uj[v.J3(cidx, i, ja)];
          snd[rk][ii + 3 + cidx - 1] = k.uj[v.J3(cidx, i, ja + 1)];
        }
      }
    }
  };
  auto unpack = [&](bool from_up, int ja_kind /*0 nye-1  1 nys  2 nye+1  3 nys-2*/, bool add) {
    for (int rk = 0; rk < nr; rk++) {
      Rank &k = w->R[rk];
      V v{w, &k};
      const std::vector<double> &rcv = snd[from_up ? k.nup : k.ndown];
      const int ja = ja_kind == 0 ? k.nye - 1 : ja_kind == 1 ? k.nys : ja_kind == 2 ? k.nye + 1 : k.nys - 2;
      for (int i = nxs - 2; i <= nxe + 2; i++) {
        const int ii = 6 * (i - (nxs - 2));
        for (int cidx = 1; cidx <= 3; cidx++) {
          if (add) {
            k.uj[v.J3(cidx, i, ja)] = k.uj[v.J3(cidx, i, ja)] + rcv[ii + cidx - 1];
            k.uj[v.J3(cidx, i, ja + 1)] = k.uj[v.J3(cidx, i, ja + 1)] + rcv[ii + 3 + cidx - 1];
          } else {
            k.uj[v.J3(cidx, i, ja)] = rcv[ii + cidx - 1];
            k.uj[v.J3(cidx, i, ja + 1)] = rcv[ii + 3 + cidx - 1];
          }
        }
      }
    }
  };
  // send to rank-1: my nys-2,nys-1 -> ndown; I receive nup's and add into nye-1,nye   :369-398
  pack(0);
  unpack(true, 0, true);
  // send to rank+1: my nye+1,nye+2 -> nup; add ndown's into nys,nys+1                 :400-429
  pack(1);
  unpack(false, 1, true);
  // refresh ghosts: my nys,nys+1 -> ndown; nup's go to my nye+1,nye+2                 :433-462
  pack(2);
  unpack(true, 2, false);
  // my nye-1,nye -> nup; ndown's go to my nys-2,nys-1                                 :464-493
  pack(3);
  unpack(false, 3, false);
  // the wall modules end after the y exchange (proj/reconnection/boundary_reconnection.f90:364-502)
  if (w->c.bc != ORC_BC_PERIODIC) return;
  // x fold + copy back                                                                :495-506
  for (Rank &k : w->R) {
    V v{w, &k};
    for (int j = k.nys - 2; j <= k.nye + 2; j++)
      for (int cidx = 1; cidx <= 3; cidx++) {
        k.uj[v.J3(cidx, nxe - 1, j)] = k.uj[v.J3(cidx, nxe - 1, j)] + k.uj[v.J3(cidx, nxs - 2, j)];
        k.uj[v.J3(cidx, nxe, j)] = k.uj[v.J3(cidx, nxe, j)] + k.uj[v.J3(cidx, nxs - 1, j)];
        k.uj[v.J3(cidx, nxs, j)] = k.uj[v.J3(cidx, nxs, j)] + k.uj[v.J3(cidx, nxe + 1, j)];
        k.uj[v.J3(cidx, nxs + 1, j)] = k.uj[v.J3(cidx, nxs + 1, j)] + k.uj[v.J3(cidx, nxe + 2, j)];
      }
    for (int j = k.nys - 2; j <= k.nye + 2; j++)
      for (int cidx = 1; cidx <= 3; cidx++) {
        k.uj[v.J3(cidx, nxs - 2, j)] = k.uj[v.J3(cidx, nxe - 1, j)];
        k.uj[v.J3(cidx, nxs - 1, j)] = k.uj[v.J3(cidx, nxe, j)];
        k.uj[v.J3(cidx, nxe + 1, j)] = k.uj[v.J3(cidx, nxs, j)];
        k.uj[v.J3(cidx, nxe + 2, j)] = k.uj[v.J3(cidx, nxs + 1, j)];
      }
  }
}

// boundary_periodic.f90:251-354
void bc_dfield(orc_world *w) {
  const int nxs = w->nxs, nxe = w->nxe;
  const int n = 12 * (nxe - nxs + 1);
  const int nr = (int)w->R.size();
  std::vector<std::vector<double>> snd(nr, std::vector<double>(n));
  // my nys,nys+1 -> ndown ; nup's land in nye+1,nye+2                                 :263-303
  for (int rk = 0; rk < nr; rk++) {
    Rank &k = w->R[rk];
    V v{w, &k};
    for (int i = nxs; i <= nxe; i++) {
      const int ii = 12 * (i - nxs);
      for (int cidx = 1; cidx <= 6; cidx++) {
        snd[rk][ii + cidx - 1] = k.df[v.F6(cidx, i, k.nys)];
        snd[rk][ii + 6 + cidx - 1] = k.df[v.F6(cidx, i, k.nys + 1)];
      }
    }
  }
  for (int rk = 0; rk < nr; rk++) {
    Rank &k = w->R[rk];
    V v{w, &k};
    const std::vector<double> &rcv = snd[k.nup];
    for (int i = nxs; i <= nxe; i++) {
      const int ii = 12 * (i - nxs);
      for (int cidx = 1; cidx <= 6; cidx++) {
        k.df[v.F6(cidx, i, k.nye + 1)] = rcv[ii + cidx - 1];
        k.df[v.F6(cidx, i, k.nye + 2)] = rcv[ii + 6 + cidx - 1];
      }
    }
  }
  // my nye-1,nye -> nup ; ndown's land in nys-2,nys-1                                 :305-345
  // NB: packed after the first unpack, exactly as in the reference (matters when nyl<=2).
  for (int rk = 0; rk < nr; rk++) {
    Rank &k = w->R[rk];
    V v{w, &k};
    for (int i = nxs; i <= nxe; i++) {
      const int ii = 12 * (i - nxs);
      for (int cidx = 1; cidx <= 6; cidx++) {
        snd[rk][ii + cidx - 1] = k.df[v.F6(cidx, i, k.nye - 1)];
        snd[rk][ii + 6 + cidx - 1] = k.df[v.F6(cidx, i, k.nye)];
      }
    }
  }
  for (int rk = 0; rk < nr; rk++) {
    Rank &k = w->R[rk];
    V v{w, &k};
    const std::vector<double> &rcv = snd[k.ndown];
    for (int i = nxs; i <= nxe; i++) {
      const int ii = 12 * (i - nxs);
      for (int cidx = 1; cidx <= 6; cidx++) {
        k.df[v.F6(cidx, i, k.nys - 2)] = rcv[ii + cidx - 1];
        k.df[v.F6(cidx, i, k.nys - 1)] = rcv[ii + 6 + cidx - 1];
      }
    }
  }
  if (w->c.bc == ORC_BC_SHOCK) {
    // conducting wall on the left, zero on the right  (proj/shock/boundary_shock.f90:396-405)
    for (Rank &k : w->R) {
      V v{w, &k};
      for (int j = k.nys - 2; j <= k.nye + 2; j++) {
        k.df[v.F6(1, nxs - 1, j)] = -k.df[v.F6(1, nxs, j)];
        for (int cidx = 2; cidx <= 4; cidx++) k.df[v.F6(cidx, nxs - 1, j)] = k.df[v.F6(cidx, nxs + 1, j)];
        for (int cidx = 5; cidx <= 6; cidx++) k.df[v.F6(cidx, nxs - 1, j)] = -k.df[v.F6(cidx, nxs, j)];
        for (int cidx = 1; cidx <= 6; cidx++) k.df[v.F6(cidx, nxe + 1, j)] = 0.0;
      }
    }
    return;
  }
  if (w->c.bc == ORC_BC_RECONNECTION) {
    // conducting walls: odd / even mirror  (proj/reconnection/boundary_reconnection.f90:349-359)
    for (Rank &k : w->R) {
      V v{w, &k};
      for (int j = k.nys - 2; j <= k.nye + 2; j++) {
        k.df[v.F6(1, nxs - 1, j)] = -k.df[v.F6(1, nxs, j)];
        for (int cidx = 2; cidx <= 4; cidx++) k.df[v.F6(cidx, nxs - 1, j)] = k.df[v.F6(cidx, nxs + 1, j)];
        for (int cidx = 5; cidx <= 6; cidx++) k.df[v.F6(cidx, nxs - 1, j)] = -k.df[v.F6(cidx, nxs, j)];
        k.df[v.F6(1, nxe, j)] = -k.df[v.F6(1, nxe - 1, j)];
        for (int cidx = 2; cidx <= 4; cidx++) k.df[v.F6(cidx, nxe + 1, j)] = k.df[v.F6(cidx, nxe - 1, j)];
        for (int cidx = 5; cidx <= 6; cidx++) k.df[v.F6(cidx, nxe, j)] = -k.df[v.F6(cidx, nxe - 1, j)];
      }
    }
    return;
  }
  // x ghosts by periodic copy                                                          :347-352
  for (Rank &k : w->R) {
    V v{w, &k};
    for (int j = k.nys - 2; j <= k.nye + 2; j++)
      for (int cidx = 1; cidx <= 6; cidx++) {
        k.df[v.F6(cidx, nxs - 2, j)] = k.df[v.F6(cidx, nxe - 1, j)];
        k.df[v.F6(cidx, nxs - 1, j)] = k.df[v.F6(cidx, nxe, j)];
        k.df[v.F6(cidx, nxe + 1, j)] = k.df[v.F6(cidx, nxs, j)];
        k.df[v.F6(cidx, nxe + 2, j)] = k.df[v.F6(cidx, nxs + 1, j)];
      }
  }
}

// boundary_periodic.f90:511-568 ; which: 0 = phi, 1 = p
void bc_phi(orc_world *w, int which, int l) {
  const int nxs = w->nxs, nxe = w->nxe;
  const int nr = (int)w->R.size();
  std::vector<std::vector<double>> snd(nr, std::vector<double>(w->nx));
  auto arr = [&](Rank &k) -> std::vector<double> & { return which == 0 ? k.phi : k.p; };
  for (int rk = 0; rk < nr; rk++) {  // my row nys -> ndown; nup's -> my nye+1       :523-541
    Rank &k = w->R[rk];
    V v{w, &k};
    for (int i = nxs; i <= nxe; i++) snd[rk][i - nxs] = arr(k)[v.PH(i, k.nys)];
  }
  for (int rk = 0; rk < nr; rk++) {
    Rank &k = w->R[rk];
    V v{w, &k};
    for (int i = nxs; i <= nxe; i++) arr(k)[v.PH(i, k.nye + 1)] = snd[k.nup][i - nxs];
  }
  for (int rk = 0; rk < nr; rk++) {  // my row nye -> nup; ndown's -> my nys-1       :543-561
    Rank &k = w->R[rk];
    V v{w, &k};
    for (int i = nxs; i <= nxe; i++) snd[rk][i - nxs] = arr(k)[v.PH(i, k.nye)];
  }
  for (int rk = 0; rk < nr; rk++) {
    Rank &k = w->R[rk];
    V v{w, &k};
    for (int i = nxs; i <= nxe; i++) arr(k)[v.PH(i, k.nys - 1)] = snd[k.ndown][i - nxs];
  }
  if (w->c.bc != ORC_BC_PERIODIC) {
    // l = 1 (Bx) odd, l = 2,3 even at the left wall; zero at the right wall
    // (proj/reconnection/boundary_reconnection.f90:557-577, proj/shock/boundary_shock.f90:603-623)
    for (Rank &k : w->R) {
      V v{w, &k};
      for (int j = k.nys - 1; j <= k.nye + 1; j++) {
        arr(k)[v.PH(nxs - 1, j)] = (l == 1) ? -arr(k)[v.PH(nxs, j)] : arr(k)[v.PH(nxs + 1, j)];
        arr(k)[v.PH(nxe + 1, j)] = 0.0;
      }
    }
    return;
  }
  for (Rank &k : w->R) {  // :563-566
    V v{w, &k};
    for (int j = k.nys - 1; j <= k.nye + 1; j++) {
      arr(k)[v.PH(nxs - 1, j)] = arr(k)[v.PH(nxe, j)];
      arr(k)[v.PH(nxe + 1, j)] = arr(k)[v.PH(nxs, j)];
    }
  }
}

// boundary_periodic.f90:61-96, on gp (proj/weibel/app.f90:105)
void bc_particle_x(orc_world *w) {
  const int nxgs = w->nxgs, nxge = w->nxge;
  const double delx = w->delx;
  if (w->c.bc != ORC_BC_PERIODIC) {
    // reflecting walls at nxs+1 and nxe-1  (proj/reconnection/boundary_reconnection.f90:61-99,
    // proj/shock/boundary_shock.f90:62-100)
    const int nxs = w->nxs, nxe = w->nxe;
    for (Rank &k : w->R) {
      V v{w, &k};
      for (int isp = 1; isp <= w->nsp; isp++) {
#pragma omp parallel for
        for (int j = k.nys; j <= k.nye; j++)
          for (int ii = 1; ii <= k.np2[v.NP2(j, isp)]; ii++) {
            double *r = &k.gp[v.UP(1, ii, j, isp)];
            const int ipos = (int)(r[0] / delx);
            if (ipos < nxs + 1) {
              r[0] = 2. * (nxs + 1) * delx - r[0];
              r[2] = -r[2];
              r[3] = -r[3];
              r[4] = -r[4];
            } else if (ipos >= nxe - 1) {
              r[0] = 2. * (nxe - 1) * delx - r[0];
              r[2] = -r[2];
              r[3] = -r[3];
              r[4] = -r[4];
            }
          }
      }
    }
    return;
  }
  for (Rank &k : w->R) {
    V v{w, &k};
    for (int isp = 1; isp <= w->nsp; isp++) {
#pragma omp parallel
      {
        const int old = std::fegetround();
        std::fesetround(FE_DOWNWARD);  // ieee_set_rounding_mode(ieee_down)  :74
#pragma omp for
        for (int j = k.nys; j <= k.nye; j++)
          for (int ii = 1; ii <= k.np2[v.NP2(j, isp)]; ii++) {
            volatile double *x = &k.gp[v.UP(1, ii, j, isp)];
            const int ipos = (int)(*x / delx);
            if (ipos < nxgs)
              *x = *x + (nxge - nxgs + 1) * delx;
            else if (ipos >= nxge + 1)
              *x = *x - (nxge - nxgs + 1) * delx;
          }
        std::fesetround(old);
      }
    }
  }
}

// boundary_periodic.f90:99-248, on gp (proj/weibel/app.f90:106)
int bc_particle_y(orc_world *w) {
  const int nygs = w->nygs, nyge = w->nyge, np = w->np;
  const double delx = w->delx;
  int err = 0;
  for (int isp = 1; isp <= w->nsp; isp++) {
    // scan rows, wrap y, remember holes and leavers                                   :143-169
    for (Rank &k : w->R) {
      V v{w, &k};
      k.flag.assign((size_t)np * k.nyl, 0);
      k.cnt2.assign(k.nyl, 0);
      k.bff.assign(k.nyl + 2, {});
      std::vector<std::vector<int>> goes_dn(k.nyl), goes_up(k.nyl);
#pragma omp parallel
      {
        const int old = std::fegetround();
        std::fesetround(FE_DOWNWARD);  // :124
#pragma omp for
        for (int j = k.nys; j <= k.nye; j++) {
          int c2 = 0;
          for (int ii = 1; ii <= k.np2[v.NP2(j, isp)]; ii++) {
            volatile double *y = &k.gp[v.UP(2, ii, j, isp)];
            const int jpos = (int)(*y / delx);
            if (jpos != j) {
              if (jpos <= nygs - 1)
                *y = *y + (nyge - nygs + 1) * delx;
              else if (jpos >= nyge + 1)
                *y = *y - (nyge - nygs + 1) * delx;
              if (jpos == j - 1)
                goes_dn[j - k.nys].push_back(ii);
              else if (jpos == j + 1)
                goes_up[j - k.nys].push_back(ii);
              else {
#pragma omp atomic write
                err = 2;  // moved more than one row: CFL violated
              }
              c2++;
              k.flag[(size_t)(c2 - 1) + (size_t)np * (j - k.nys)] = ii;
            }
          }
          k.cnt2[j - k.nys] = c2;
        }
        std::fesetround(old);
      }
      // bff_ptcl(:,jpos): fixed append order (reference order is lock-arbitrary)
#pragma omp parallel for
      for (int jpos = k.nys - 1; jpos <= k.nye + 1; jpos++) {
        std::vector<double> &b = k.bff[jpos - (k.nys - 1)];
        const int jm = jpos - 1, jp = jpos + 1;
        if (jm >= k.nys && jm <= k.nye)
          for (int ii : goes_up[jm - k.nys]) {
            const double *src = &k.gp[v.UP(1, ii, jm, isp)];
            b.insert(b.end(), src, src + NDIM);
          }
        if (jp >= k.nys && jp <= k.nye)
          for (int ii : goes_dn[jp - k.nys]) {
            const double *src = &k.gp[v.UP(1, ii, jp, isp)];
            b.insert(b.end(), src, src + NDIM);
          }
      }
    }
    // transfer to rank-1: bff(:,nys-1) -> ndown, appended to its bff(:,nye)           :173-180
    for (Rank &k : w->R) {
      const std::vector<double> &in = w->R[k.nup].bff[0];
      std::vector<double> &dst = k.bff[k.nyl];  // row nye
      dst.insert(dst.end(), in.begin(), in.end());
    }
    // transfer to rank+1: bff(:,nye+1) -> nup, appended to its bff(:,nys)             :182-189
    for (Rank &k : w->R) {
      const std::vector<double> &in = w->R[k.ndown].bff[w->R[k.ndown].nyl + 1];
      std::vector<double> &dst = k.bff[1];  // row nys
      dst.insert(dst.end(), in.begin(), in.end());
    }
    // fill holes / compact / append                                                    :193-236
    for (Rank &k : w->R) {
      V v{w, &k};
#pragma omp parallel for
      for (int j = k.nys; j <= k.nye; j++) {
        const std::vector<double> &b = k.bff[j - (k.nys - 1)];
        int cnt = (int)(b.size() / NDIM);
        int &np2 = k.np2[v.NP2(j, isp)];
        const int *flag = &k.flag[(size_t)np * (j - k.nys)];  // flag(1..,j) -> flag[0..]
        const int cnt2 = k.cnt2[j - k.nys];
        int iii = 0;
        int cnt_tmp = cnt2;
        for (int ii = 1; ii <= cnt2; ii++) {  // loop1
          if (cnt == 0) {
            if (np2 < flag[ii - 1]) goto done;
            while (np2 == flag[cnt_tmp - 1]) {
              np2 = np2 - 1;
              if (np2 < flag[ii - 1]) goto done;
              cnt_tmp = cnt_tmp - 1;
            }
            for (int idim = 1; idim <= NDIM; idim++)
              k.gp[v.UP(idim, flag[ii - 1], j, isp)] = k.gp[v.UP(idim, np2, j, isp)];
            np2 = np2 - 1;
          } else {
            for (int idim = 1; idim <= NDIM; idim++)
              k.gp[v.UP(idim, flag[ii - 1], j, isp)] = b[(size_t)idim - 1 + NDIM * iii];
            iii = iii + 1;
            cnt = cnt - 1;
          }
        }
      done:
        if (cnt > 0) {
          if (np2 + cnt > np) {  // "memory over (np2 > np)"  :231-234
#pragma omp atomic write
            err = 1;
            continue;
          }
          for (int ii = 1; ii <= cnt; ii++)
            for (int idim = 1; idim <= NDIM; idim++)
              k.gp[v.UP(idim, np2 + ii, j, isp)] = b[(size_t)NDIM * iii + idim - 1 + (size_t)NDIM * (ii - 1)];
        }
        np2 = np2 + cnt;
      }
    }
  }
  return err;
}

// common/sort.f90:36-82 : (out) up <- (in) gp, cumcnt
void sort_bucket(orc_world *w) {
  const int nxs = w->nxs, nxe = w->nxe;
  for (Rank &k : w->R) {
    V v{w, &k};
    for (int isp = 1; isp <= w->nsp; isp++) {
#pragma omp parallel
      {
        std::vector<int> cnt(w->nx), sum_cnt(w->nx + 1);
#pragma omp for
        for (int j = k.nys; j <= k.nye; j++) {
          std::fill(cnt.begin(), cnt.end(), 0);
          const int n = k.np2[v.NP2(j, isp)];
          for (int ii = 1; ii <= n; ii++) {
            const int i = (int)k.gp[v.UP(1, ii, j, isp)];
            cnt[i - nxs] = cnt[i - nxs] + 1;
          }
          sum_cnt[0] = 0;
          k.cumcnt[v.CUM(nxs, j, isp)] = 0;
          for (int i = nxs + 1; i <= nxe + 1; i++) {
            sum_cnt[i - nxs] = sum_cnt[i - 1 - nxs] + cnt[i - 1 - nxs];
            k.cumcnt[v.CUM(i, j, isp)] = sum_cnt[i - nxs];
          }
          for (int ii = 1; ii <= n; ii++) {
            const int i = (int)k.gp[v.UP(1, ii, j, isp)];
            std::memcpy(&k.up[v.UP(1, sum_cnt[i - nxs] + 1, j, isp)], &k.gp[v.UP(1, ii, j, isp)], NDIM * sizeof(double));
            sum_cnt[i - nxs] = sum_cnt[i - nxs] + 1;
          }
        }
      }
    }
  }
}

// ---------------------------------------------------------------- field solve
double ranksum(const orc_world *w, const std::vector<std::vector<double>> &rows) {
  // per-row partials added in row order, then in rank order
  double g = 0.0;
  for (size_t rk = 0; rk < w->R.size(); rk++) {
    double s = 0.0;
    for (double x : rows[rk]) s += x;
    g += s;
  }
  return g;
}

// common/field.f90:319-461
int cgm(orc_world *w) {
  const int ite_max = 100;
  const double err = 1e-6;
  const int nxs = w->nxs, nxe = w->nxe;
  const double f4 = w->f4, f5 = w->f5;
  const int nr = (int)w->R.size();
  std::vector<std::vector<double>> rs(nr), rs2(nr);
  for (int rk = 0; rk < nr; rk++) {
    rs[rk].assign(w->R[rk].nyl, 0.0);
    rs2[rk].assign(w->R[rk].nyl, 0.0);
  }
  for (int l = 1; l <= 3; l++) {
    int ite = 0;
    // initial guess                                                                    :349-360
    for (int rk = 0; rk < nr; rk++) {
      Rank &k = w->R[rk];
      V v{w, &k};
#pragma omp parallel for
      for (int j = k.nys; j <= k.nye; j++) {
        double s = 0.0;
        for (int i = nxs; i <= nxe; i++) {
          k.phi[v.PH(i, j)] = k.df[v.F6(l, i, j)];
          k.b[v.RI(i, j)] = f5 * k.gkl[v.G3(l, i, j)];
          s = s + k.b[v.RI(i, j)] * k.b[v.RI(i, j)];
        }
        rs[rk][j - k.nys] = s;
      }
    }
    double sum_g = ranksum(w, rs);        // MPI_ALLREDUCE :362
    const double eps = std::sqrt(sum_g) * err;  // :364
    bc_phi(w, 0, l);                      // :367
    for (int rk = 0; rk < nr; rk++) {     // :370-381
      Rank &k = w->R[rk];
      V v{w, &k};
#pragma omp parallel for
      for (int j = k.nys; j <= k.nye; j++) {
        double s = 0.0;
        for (int i = nxs; i <= nxe; i++) {
          k.r[v.RI(i, j)] = k.b[v.RI(i, j)] + k.phi[v.PH(i, j - 1)] + k.phi[v.PH(i - 1, j)] - f4 * k.phi[v.PH(i, j)] + k.phi[v.PH(i + 1, j)] + k.phi[v.PH(i, j + 1)];
          k.p[v.PH(i, j)] = k.r[v.RI(i, j)];
          s = s + k.r[v.RI(i, j)] * k.r[v.RI(i, j)];
        }
        rs[rk][j - k.nys] = s;
      }
    }
    double sumr_g = ranksum(w, rs);  // :383
    if (std::sqrt(sumr_g) > eps) {   // :385
      while (sum_g > eps) {          // :387  (first pass compares sum(b^2), not its sqrt)
        ite = ite + 1;
        bc_phi(w, 1, l);             // :392
        for (int rk = 0; rk < nr; rk++) {  // :395-407
          Rank &k = w->R[rk];
          V v{w, &k};
#pragma omp parallel for
          for (int j = k.nys; j <= k.nye; j++) {
            double s = 0.0, s2 = 0.0;
            for (int i = nxs; i <= nxe; i++) {
              k.ap[v.RI(i, j)] = -k.p[v.PH(i, j - 1)] - k.p[v.PH(i - 1, j)] + f4 * k.p[v.PH(i, j)] - k.p[v.PH(i + 1, j)] - k.p[v.PH(i, j + 1)];
              s = s + k.r[v.RI(i, j)] * k.r[v.RI(i, j)];
              s2 = s2 + k.p[v.PH(i, j)] * k.ap[v.RI(i, j)];
            }
            rs[rk][j - k.nys] = s;
            rs2[rk][j - k.nys] = s2;
          }
        }
        sumr_g = ranksum(w, rs);  // :409-413
        const double sum2_g = ranksum(w, rs2);
        const double av = sumr_g / sum2_g;  // :415
        for (int rk = 0; rk < nr; rk++) {   // :417-424
          Rank &k = w->R[rk];
          V v{w, &k};
#pragma omp parallel for
          for (int j = k.nys; j <= k.nye; j++)
            for (int i = nxs; i <= nxe; i++) {
              k.phi[v.PH(i, j)] = k.phi[v.PH(i, j)] + av * k.p[v.PH(i, j)];
              k.r[v.RI(i, j)] = k.r[v.RI(i, j)] - av * k.ap[v.RI(i, j)];
            }
        }
        sum_g = std::sqrt(sumr_g);  // :426 (residual *before* this update)
        if (ite >= ite_max) {       // :427-430
          w->cg_ite[l - 1] = ite;
          return 1;
        }
        for (int rk = 0; rk < nr; rk++) {  // :432-439
          Rank &k = w->R[rk];
          V v{w, &k};
#pragma omp parallel for
          for (int j = k.nys; j <= k.nye; j++) {
            double s = 0.0;
            for (int i = nxs; i <= nxe; i++) s = s + k.r[v.RI(i, j)] * k.r[v.RI(i, j)];
            rs[rk][j - k.nys] = s;
          }
        }
        const double sum1_g = ranksum(w, rs);  // :441
        const double bv = sum1_g / sumr_g;     // :442
        for (int rk = 0; rk < nr; rk++) {      // :444-450
          Rank &k = w->R[rk];
          V v{w, &k};
#pragma omp parallel for
          for (int j = k.nys; j <= k.nye; j++)
            for (int i = nxs; i <= nxe; i++) k.p[v.PH(i, j)] = k.r[v.RI(i, j)] + bv * k.p[v.PH(i, j)];
        }
      }
    }
    for (Rank &k : w->R) {  // :455-457
      V v{w, &k};
#pragma omp parallel for
      for (int j = k.nys; j <= k.nye; j++)
        for (int i = nxs; i <= nxe; i++) k.df[v.F6(l, i, j)] = k.phi[v.PH(i, j)];
    }
    w->cg_ite[l - 1] = ite;
  }
  return 0;
}

// common/field.f90:66-186
int field_fdtd_i(orc_world *w) {
  const int nxs = w->nxs, nxe = w->nxe;
  const double f1 = w->f1, f2 = w->f2, f3 = w->f3, gfac = w->gfac, delt = w->delt;
  for (Rank &k : w->R) ele_cur_rank(w, k);  // :121
  bc_curre(w);                              // :122
  for (Rank &k : w->R) {                    // :125-146
    V v{w, &k};
    const double *uf = k.uf.data();
    const double *uj = k.uj.data();
#pragma omp parallel for
    for (int j = k.nys; j <= k.nye; j++)
      for (int i = nxs; i <= nxe; i++) {
        k.gkl[v.G3(1, i, j)] = +f2 * (+uf[v.F6(1, i, j - 1)] + uf[v.F6(1, i - 1, j)] - 4. * uf[v.F6(1, i, j)] + uf[v.F6(1, i + 1, j)] + uf[v.F6(1, i, j + 1)] + f3 * (-uj[v.J3(3, i, j - 1)] + uj[v.J3(3, i, j)])) - f1 * (-uf[v.F6(6, i, j - 1)] + uf[v.F6(6, i, j)]);
        k.gkl[v.G3(2, i, j)] = +f2 * (+uf[v.F6(2, i, j - 1)] + uf[v.F6(2, i - 1, j)] - 4. * uf[v.F6(2, i, j)] + uf[v.F6(2, i + 1, j)] + uf[v.F6(2, i, j + 1)] - f3 * (-uj[v.J3(3, i - 1, j)] + uj[v.J3(3, i, j)])) + f1 * (-uf[v.F6(6, i - 1, j)] + uf[v.F6(6, i, j)]);
        k.gkl[v.G3(3, i, j)] = +f2 * (+uf[v.F6(3, i, j - 1)] + uf[v.F6(3, i - 1, j)] - 4. * uf[v.F6(3, i, j)] + uf[v.F6(3, i + 1, j)] + uf[v.F6(3, i, j + 1)] + f3 * (-uj[v.J3(2, i - 1, j)] + uj[v.J3(2, i, j)] + uj[v.J3(1, i, j - 1)] - uj[v.J3(1, i, j)])) - f1 * (-uf[v.F6(5, i - 1, j)] + uf[v.F6(5, i, j)] + uf[v.F6(4, i, j - 1)] - uf[v.F6(4, i, j)]);
      }
  }
  if (cgm(w)) return 1;  // :149
  bc_dfield(w);          // :151
  for (Rank &k : w->R) { // :154-171
    V v{w, &k};
    const double *uf = k.uf.data();
    const double *uj = k.uj.data();
    double *df = k.df.data();
#pragma omp parallel for
    for (int j = k.nys; j <= k.nye; j++)
      for (int i = nxs; i <= nxe; i++) {
        df[v.F6(4, i, j)] = +f1 * (+gfac * (-df[v.F6(3, i, j)] + df[v.F6(3, i, j + 1)]) + (-uf[v.F6(3, i, j)] + uf[v.F6(3, i, j + 1)])) - 4. * PI * delt * uj[v.J3(1, i, j)];
        df[v.F6(5, i, j)] = -f1 * (+gfac * (-df[v.F6(3, i, j)] + df[v.F6(3, i + 1, j)]) + (-uf[v.F6(3, i, j)] + uf[v.F6(3, i + 1, j)])) - 4. * PI * delt * uj[v.J3(2, i, j)];
        df[v.F6(6, i, j)] = +f1 * (+gfac * (-df[v.F6(2, i, j)] + df[v.F6(2, i + 1, j)] + df[v.F6(1, i, j)] - df[v.F6(1, i, j + 1)]) + (-uf[v.F6(2, i, j)] + uf[v.F6(2, i + 1, j)] + uf[v.F6(1, i, j)] - uf[v.F6(1, i, j + 1)])) - 4. * PI * delt * uj[v.J3(3, i, j)];
      }
  }
  bc_dfield(w);          // :173
  for (Rank &k : w->R) { // :176-184: uf += df over nxs-2..nxe+2 (the active range of the call), nys-2..nye+2
    V v{w, &k};
#pragma omp parallel for
    for (int j = k.nys - 2; j <= k.nye + 2; j++)
      for (int i = nxs - 2; i <= nxe + 2; i++)
        for (int ieq = 1; ieq <= 6; ieq++) k.uf[v.F6(ieq, i, j)] = k.uf[v.F6(ieq, i, j)] + k.df[v.F6(ieq, i, j)];
  }
  return 0;
}

// ---------------------------------------------------------------- moments
// common/mom_calc.f90:167-249 on gp (app.f90:122)
void mom_nvt(orc_world *w) {
  const double d_delx = w->d_delx, c = w->cc;
  for (Rank &k : w->R) {
    V v{w, &k};
    std::fill(k.mom.begin(), k.mom.end(), 0.0);
    // REDUCTION(+:mom): rows 3 apart touch disjoint mom rows (jh in {j-1,j}, +1) -> 3 phases
    for (int isp = 1; isp <= w->nsp; isp++)
      for (int phase = 0; phase < 3; phase++) {
#pragma omp parallel for
        for (int j = k.nys + phase; j <= k.nye; j += 3)
          for (int ii = 1; ii <= k.np2[v.NP2(j, isp)]; ii++) {
            const double *pu = &k.gp[v.UP(1, ii, j, isp)];
            const int ih = (int)(pu[0] * d_delx - 0.5);
            const int jh = (int)(pu[1] * d_delx - 0.5);
            const double dx = pu[0] * d_delx - 0.5 - ih;
            const double dxm = 1. - dx;
            const double dy = pu[1] * d_delx - 0.5 - jh;
            const double dym = 1. - dy;
            const double gam = 1. / std::sqrt(1.0 + (+pu[2] * pu[2] + pu[3] * pu[3] + pu[4] * pu[4]) / (c * c));
            const double val[7] = {1.0, pu[2] * gam, pu[3] * gam, pu[4] * gam, pu[2] * pu[2] * gam, pu[3] * pu[3] * gam, pu[4] * pu[4] * gam};
            for (int m = 1; m <= 7; m++) {
              // N is added as dxm*dym etc.; the others as val*dxm*dym (left to right)
              if (m == 1) {
                k.mom[v.MOM(m, ih, jh, isp)] += dxm * dym;
                k.mom[v.MOM(m, ih + 1, jh, isp)] += dx * dym;
                k.mom[v.MOM(m, ih, jh + 1, isp)] += dxm * dy;
                k.mom[v.MOM(m, ih + 1, jh + 1, isp)] += dx * dy;
              } else {
                k.mom[v.MOM(m, ih, jh, isp)] += val[m - 1] * dxm * dym;
                k.mom[v.MOM(m, ih + 1, jh, isp)] += val[m - 1] * dx * dym;
                k.mom[v.MOM(m, ih, jh + 1, isp)] += val[m - 1] * dxm * dy;
                k.mom[v.MOM(m, ih + 1, jh + 1, isp)] += val[m - 1] * dx * dy;
              }
            }
          }
      }
  }
}

// boundary_periodic.f90:571-636
void bc_mom(orc_world *w) {
  const int nxgs = w->nxgs, nxge = w->nxge, nk = 7;
  const int nr = (int)w->R.size();
  for (Rank &k : w->R) {  // :579-584
    V v{w, &k};
    for (int isp = 1; isp <= w->nsp; isp++)
      for (int j = k.nys - 1; j <= k.nye + 1; j++)
        for (int m = 1; m <= nk; m++) {
          if (w->c.bc == ORC_BC_PERIODIC) {
            k.mom[v.MOM(m, nxgs, j, isp)] = k.mom[v.MOM(m, nxgs, j, isp)] + k.mom[v.MOM(m, nxge + 1, j, isp)];
            k.mom[v.MOM(m, nxge, j, isp)] = k.mom[v.MOM(m, nxge, j, isp)] + k.mom[v.MOM(m, nxgs - 1, j, isp)];
          } else {  // walls: the ghost column folds back onto its own side (boundary_reconnection.f90:590-595)
            k.mom[v.MOM(m, nxgs, j, isp)] = k.mom[v.MOM(m, nxgs, j, isp)] + k.mom[v.MOM(m, nxgs - 1, j, isp)];
            k.mom[v.MOM(m, nxge, j, isp)] = k.mom[v.MOM(m, nxge, j, isp)] + k.mom[v.MOM(m, nxge + 1, j, isp)];
          }
        }
  }
  const int n = nk * (nxge - nxgs + 3);
  std::vector<std::vector<double>> snd(nr, std::vector<double>(n));
  for (int isp = 1; isp <= w->nsp; isp++) {
    for (int rk = 0; rk < nr; rk++) {  // my nys-1 -> ndown; add nup's into nye      :587-609
      Rank &k = w->R[rk];
      V v{w, &k};
      for (int i = nxgs - 1; i <= nxge + 1; i++)
        for (int m = 1; m <= nk; m++) snd[rk][nk * (i - (nxgs - 1)) + m - 1] = k.mom[v.MOM(m, i, k.nys - 1, isp)];
    }
    for (int rk = 0; rk < nr; rk++) {
      Rank &k = w->R[rk];
      V v{w, &k};
      for (int i = nxgs - 1; i <= nxge + 1; i++)
        for (int m = 1; m <= nk; m++) k.mom[v.MOM(m, i, k.nye, isp)] = k.mom[v.MOM(m, i, k.nye, isp)] + snd[k.nup][nk * (i - (nxgs - 1)) + m - 1];
    }
    for (int rk = 0; rk < nr; rk++) {  // my nye+1 -> nup; add ndown's into nys      :611-633
      Rank &k = w->R[rk];
      V v{w, &k};
      for (int i = nxgs - 1; i <= nxge + 1; i++)
        for (int m = 1; m <= nk; m++) snd[rk][nk * (i - (nxgs - 1)) + m - 1] = k.mom[v.MOM(m, i, k.nye + 1, isp)];
    }
    for (int rk = 0; rk < nr; rk++) {
      Rank &k = w->R[rk];
      V v{w, &k};
      for (int i = nxgs - 1; i <= nxge + 1; i++)
        for (int m = 1; m <= nk; m++) k.mom[v.MOM(m, i, k.nys, isp)] = k.mom[v.MOM(m, i, k.nys, isp)] + snd[k.ndown][nk * (i - (nxgs - 1)) + m - 1];
    }
  }
}

// ---------------------------------------------------------------- RNG (counter based)
inline uint64_t splitmix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
inline uint64_t rng_hash(uint64_t seed, int isp, uint64_t gid, int stream) {
  uint64_t k = splitmix(seed ^ (0xD1B54A32D192ED03ULL * (uint64_t)(isp + 1)));
  k = splitmix(k + gid);
  return splitmix(k + 0x8CB92BA72F3D8DD7ULL * (uint64_t)(stream + 1));
}
inline double rng_uniform(uint64_t seed, int isp, uint64_t gid, int stream) {
  return ((double)(rng_hash(seed, isp, gid, stream) >> 11) + 0.5) * (1.0 / 9007199254740992.0);  // (0,1)
}

}  // namespace

// ================================================================ C interface
extern "C" {

orc_world *orc_create(const orc_config *cfg) {
  if (cfg->nsp < 1 || cfg->nsp > ORC_NSP_MAX || cfg->nranks < 1 || cfg->ny < cfg->nranks) return nullptr;
  if (cfg->bc != ORC_BC_PERIODIC && cfg->bc != ORC_BC_RECONNECTION && cfg->bc != ORC_BC_SHOCK) return nullptr;
  orc_world *w = new orc_world();
  w->c = *cfg;
  w->nx = cfg->nx;
  w->ny = cfg->ny;
  w->nxgs = cfg->nxgs;
  w->nxge = cfg->nxgs + cfg->nx - 1;  // app.f90:250-255
  w->nygs = cfg->nygs;
  w->nyge = cfg->nygs + cfg->ny - 1;
  w->nxs = w->nxgs;
  w->nxe = w->nxge;
  w->np = cfg->np;
  w->nsp = cfg->nsp;
  w->delx = cfg->delx;
  w->delt = cfg->delt;
  w->cc = cfg->c;
  w->gfac = cfg->gfac;
  w->d_delx = 1. / w->delx;  // particle.f90:41
  w->d_delt = 1. / w->delt;  // field.f90:59
  // field.f90:53-57
  w->f1 = w->cc * w->delt / w->delx;
  w->f2 = w->gfac * w->f1 * w->f1;
  w->f3 = 4.0 * PI * w->delx / w->cc;
  {
    const double t = w->delx / (w->cc * w->delt * w->gfac);
    w->f4 = 4.0 + t * t;
    w->f5 = t * t;
  }
  w->cg_ite[0] = w->cg_ite[1] = w->cg_ite[2] = 0;
  for (double &t : w->t_stage) t = 0.0;
  const int nsize = cfg->nranks;
  w->R.resize(nsize);
  for (int nrank = 0; nrank < nsize; nrank++) {
    Rank &k = w->R[nrank];
    // common/mpi_set.f90:36-47
    const int iwork1 = (w->nyge - w->nygs + 1) / nsize;
    const int iwork2 = (w->nyge - w->nygs + 1) % nsize;
    k.nys = nrank * iwork1 + w->nygs + std::min(nrank, iwork2);
    k.nye = k.nys + iwork1 - 1;
    if (iwork2 > nrank) k.nye = k.nye + 1;
    k.nyl = k.nye - k.nys + 1;
    k.nup = nrank + 1;
    k.ndown = nrank - 1;
    if (nrank == nsize - 1) k.nup = 0;
    if (nrank == 0) k.ndown = nsize - 1;
    const size_t npart = (size_t)NDIM * w->np * k.nyl * w->nsp;
    const size_t ng = (size_t)(w->nx + 4) * (k.nyl + 4);
    k.up.assign(npart, 0.0);
    k.gp.assign(npart, 0.0);
    k.uf.assign(6 * ng, 0.0);
    k.df.assign(6 * ng, 0.0);  // SAVEd warm start, zero on first call (field.f90:105-119)
    k.uj.assign(3 * ng, 0.0);
    k.gkl.assign((size_t)3 * w->nx * k.nyl, 0.0);
    k.mom.assign((size_t)7 * (w->nx + 2) * (k.nyl + 2) * w->nsp, 0.0);
    k.np2.assign((size_t)k.nyl * w->nsp, 0);
    k.cumcnt.assign((size_t)(w->nx + 1) * k.nyl * w->nsp, 0);
    const size_t nph = (size_t)(w->nx + 2) * (k.nyl + 2), nri = (size_t)w->nx * k.nyl;
    k.phi.assign(nph, 0.0);
    k.p.assign(nph, 0.0);
    k.r.assign(nri, 0.0);
    k.b.assign(nri, 0.0);
    k.ap.assign(nri, 0.0);
  }
  return w;
}

void orc_destroy(orc_world *w) { delete w; }

void orc_bounds(const orc_world *w, int rank, int32_t *nys, int32_t *nye) {
  *nys = w->R[rank].nys;
  *nye = w->R[rank].nye;
}

void *orc_array(orc_world *w, int rank, int which, int64_t *len) {
  Rank &k = w->R[rank];
  std::vector<double> *d = nullptr;
  switch (which) {
    case ORC_UP: d = &k.up; break;
    case ORC_GP: d = &k.gp; break;
    case ORC_UF: d = &k.uf; break;
    case ORC_DF: d = &k.df; break;
    case ORC_UJ: d = &k.uj; break;
    case ORC_GKL: d = &k.gkl; break;
    case ORC_MOM: d = &k.mom; break;
    case ORC_NP2: if (len) *len = (int64_t)k.np2.size(); return k.np2.data();
    case ORC_CUMCNT: if (len) *len = (int64_t)k.cumcnt.size(); return k.cumcnt.data();
    default: if (len) *len = 0; return nullptr;
  }
  if (len) *len = (int64_t)d->size();
  return d->data();
}

void orc_particle_solv(orc_world *w) {
  for (Rank &k : w->R) push_rank(w, k, w->delt, true);
}
void orc_ele_cur(orc_world *w) {
  for (Rank &k : w->R) ele_cur_rank(w, k);
}
void orc_bc_curre(orc_world *w) { bc_curre(w); }
int orc_field_fdtd_i(orc_world *w) { return field_fdtd_i(w); }
void orc_bc_particle_x(orc_world *w) { bc_particle_x(w); }

// boundary_shock__injection, on gp                          proj/shock/boundary_shock.f90:255-297
void orc_bc_injection(orc_world *w, double u0) {
  const int nxs = w->nxs, nxe = w->nxe;
  const double delx = w->delx, c = std::sqrt(w->cc);
  w->u_inject = u0;
  const double xend = nxe * delx + u0 / std::sqrt(1 + (u0 * u0) / (c * c)) * w->delt;
  for (Rank &k : w->R) {
    V v{w, &k};
    for (int isp = 1; isp <= w->nsp; isp++) {
#pragma omp parallel for
      for (int j = k.nys; j <= k.nye; j++)
        for (int ii = 1; ii <= k.np2[v.NP2(j, isp)]; ii++) {
          double *r = &k.gp[v.UP(1, ii, j, isp)];
          const int ipos = (int)(r[0] / delx);
          if (ipos < nxs + 1) {
            r[0] = +2. * (nxs + 1) * delx - r[0];
            r[2] = -r[2];
            r[3] = -r[3];
            r[4] = -r[4];
          } else if (r[0] > xend) {
            r[0] = +2. * xend - r[0];
            r[2] = +2. * u0 - r[2];
            r[3] = -r[3];
            r[4] = -r[4];
          }
        }
    }
  }
}
void orc_set_u_inject(orc_world *w, double u0) { w->u_inject = u0; }
// the active x range [nxs, nxe] of the shock app (proj/shock/app.f90:611-621 `relocate` moves nxe)
int orc_set_xrange(orc_world *w, int nxs, int nxe) {
  if (nxs < w->nxgs || nxe > w->nxge || nxe - nxs < 4) return 1;
  w->nxs = nxs;
  w->nxe = nxe;
  return 0;
}
int orc_bc_particle_y(orc_world *w) { return bc_particle_y(w); }
void orc_sort_bucket(orc_world *w) { sort_bucket(w); }

int orc_step(orc_world *w, int nsteps) {
  for (int it = 0; it < nsteps; it++) {  // proj/weibel/app.f90:100-107
    double t0 = now();
    orc_particle_solv(w);
    double t1 = now();
    if (w->c.bc == ORC_BC_SHOCK) orc_bc_injection(w, w->u_inject);  // proj/shock/app.f90:112-113: before the field solve
    if (field_fdtd_i(w)) return 1;
    double t2 = now();
    if (w->c.bc != ORC_BC_SHOCK) bc_particle_x(w);
    double t3 = now();
    const int e = bc_particle_y(w);
    if (e) return 10 + e;
    double t4 = now();
    sort_bucket(w);
    double t5 = now();
    w->t_stage[0] += t1 - t0;
    w->t_stage[1] += t2 - t1;
    w->t_stage[2] += t3 - t2;
    w->t_stage[3] += t4 - t3;
    w->t_stage[4] += t5 - t4;
  }
  return 0;
}

void orc_stage_times(orc_world *w, double out[5], int reset) {
  for (int s = 0; s < 5; s++) {
    out[s] = w->t_stage[s];
    if (reset) w->t_stage[s] = 0.0;
  }
}

void orc_mom_accl(orc_world *w) {
  for (Rank &k : w->R) push_rank(w, k, w->delt * 0.5, false);  // mom_calc.f90:34
}
void orc_mom_nvt(orc_world *w) { mom_nvt(w); }
void orc_bc_mom(orc_world *w) { bc_mom(w); }

void orc_cg_iters(const orc_world *w, int32_t out[3]) {
  for (int l = 0; l < 3; l++) out[l] = w->cg_ite[l];
}

// proj/weibel/app.f90:479-545
void orc_energy(orc_world *w, double *out) {
  const int nsp = w->nsp;
  const double c = w->cc;
  for (int s = 0; s < nsp + 2; s++) out[s] = 0.0;
  for (Rank &k : w->R) {
    V v{w, &k};
    for (int isp = 1; isp <= nsp; isp++) {
      std::vector<double> rows(k.nyl, 0.0);
#pragma omp parallel for
      for (int j = k.nys; j <= k.nye; j++) {
        double s = 0.0;
        for (int ii = 1; ii <= k.np2[v.NP2(j, isp)]; ii++) {
          const double *pu = &k.up[v.UP(1, ii, j, isp)];
          const double u2 = pu[2] * pu[2] + pu[3] * pu[3] + pu[4] * pu[4];
          const double gam = std::sqrt(1 + u2 / (c * c));
          s = s + w->c.r[isp - 1] * (gam - 1);
        }
        rows[j - k.nys] = s;
      }
      double s = 0.0;
      for (double x : rows) s += x;
      out[isp - 1] += s;
    }
    std::vector<double> rb(k.nyl, 0.0), re(k.nyl, 0.0);
#pragma omp parallel for
    for (int j = k.nys; j <= k.nye; j++) {
      double bf = 0.0, ef = 0.0;
      for (int i = w->nxgs; i <= w->nxge; i++) {
        bf = bf + k.uf[v.F6(1, i, j)] * k.uf[v.F6(1, i, j)] + k.uf[v.F6(2, i, j)] * k.uf[v.F6(2, i, j)] + k.uf[v.F6(3, i, j)] * k.uf[v.F6(3, i, j)];
        ef = ef + k.uf[v.F6(4, i, j)] * k.uf[v.F6(4, i, j)] + k.uf[v.F6(5, i, j)] * k.uf[v.F6(5, i, j)] + k.uf[v.F6(6, i, j)] * k.uf[v.F6(6, i, j)];
      }
      rb[j - k.nys] = bf;
      re[j - k.nys] = ef;
    }
    double bf = 0.0, ef = 0.0;
    for (double x : rb) bf += x;
    for (double x : re) ef += x;
    out[nsp] += ef / (8 * PI);
    out[nsp + 1] += bf / (8 * PI);
  }
}

double orc_gauss_residual(orc_world *w, double *scale) {
  // rho(i,j) = sum_p q S2(x_p/d - i - 1/2) S2(y_p/d - j - 1/2), periodic in x and y,
  // from the *sorted* state `up`; div E at the cell centre from uf (Ex on x-faces,
  // Ey on y-faces: particle.f90:76-77).  Gaussian units: div E = 4 pi rho (field.f90:159).
  const int nx = w->nx, ny = w->ny;
  std::vector<double> rho((size_t)nx * ny, 0.0), rabs((size_t)nx * ny, 0.0);
  auto wrapx = [&](int i) { return ((i - w->nxgs) % nx + nx) % nx; };
  auto wrapy = [&](int j) { return ((j - w->nygs) % ny + ny) % ny; };
  for (Rank &k : w->R) {
    V v{w, &k};
    for (int isp = 1; isp <= w->nsp; isp++)
      for (int j = k.nys; j <= k.nye; j++)
        for (int ii = 1; ii <= k.np2[v.NP2(j, isp)]; ii++) {
          const double *pu = &k.up[v.UP(1, ii, j, isp)];
          const int ic = (int)(pu[0] * w->d_delx), jc = (int)(pu[1] * w->d_delx);
          double sx[3], sy[3];
          double dh = pu[0] * w->d_delx - 0.5 - ic;
          sx[0] = 0.5 * (0.5 - dh) * (0.5 - dh); sx[1] = 0.75 - dh * dh; sx[2] = 0.5 * (0.5 + dh) * (0.5 + dh);
          dh = pu[1] * w->d_delx - 0.5 - jc;
          sy[0] = 0.5 * (0.5 - dh) * (0.5 - dh); sy[1] = 0.75 - dh * dh; sy[2] = 0.5 * (0.5 + dh) * (0.5 + dh);
          for (int b = -1; b <= 1; b++)
            for (int a = -1; a <= 1; a++) {
              const size_t g = (size_t)wrapx(ic + a) + (size_t)nx * wrapy(jc + b);
              rho[g] += w->c.q[isp - 1] * sx[a + 1] * sy[b + 1];
              rabs[g] += std::fabs(w->c.q[isp - 1]) * sx[a + 1] * sy[b + 1];
            }
        }
  }
  double res = 0.0, sc = 0.0;
  for (Rank &k : w->R) {
    V v{w, &k};
    for (int j = k.nys; j <= k.nye; j++)
      for (int i = w->nxgs; i <= w->nxge; i++) {
        const double dive = (k.uf[v.F6(4, i + 1, j)] - k.uf[v.F6(4, i, j)] + k.uf[v.F6(5, i, j + 1)] - k.uf[v.F6(5, i, j)]) * w->d_delx;
        const size_t g = (size_t)(i - w->nxgs) + (size_t)nx * (j - w->nygs);
        res = std::max(res, std::fabs(dive - 4.0 * PI * rho[g]));
        sc = std::max(sc, 4.0 * PI * rabs[g]);
      }
  }
  if (scale) *scale = sc;
  return res;
}

void orc_ic_weibel(orc_world *w, uint64_t seed, int n0, double vti, double vte, double t_ani, double b0) {
  const int nx = w->nx;
  for (Rank &k : w->R) {
    V v{w, &k};
    // app.f90:388-399
    for (int j = k.nys - 2; j <= k.nye + 2; j++)
      for (int i = w->nxgs - 2; i <= w->nxge + 2; i++) {
        for (int cidx = 1; cidx <= 6; cidx++) k.uf[v.F6(cidx, i, j)] = 0.0;
        k.uf[v.F6(3, i, j)] = b0;
      }
    std::fill(k.df.begin(), k.df.end(), 0.0);
    std::fill(k.up.begin(), k.up.end(), 0.0);
    const int npr = n0 * nx;  // app.f90:306
    const double sd[2] = {vti, vte};
    for (int isp = 1; isp <= w->nsp; isp++) {
#pragma omp parallel for
      for (int j = k.nys; j <= k.nye; j++) {
        k.np2[v.NP2(j, isp)] = npr;
        // app.f90:315-328
        k.cumcnt[v.CUM(w->nxgs, j, isp)] = 0;
        for (int i = w->nxgs + 1; i <= w->nxge + 1; i++) k.cumcnt[v.CUM(i, j, isp)] = k.cumcnt[v.CUM(i - 1, j, isp)] + n0;
        for (int ii = 1; ii <= npr; ii++) {
          // global id: rows are equally populated, app.f90:442-474
          const uint64_t pid = (uint64_t)(j - w->nygs) * (uint64_t)npr + (uint64_t)ii;
          double *pu = &k.up[v.UP(1, ii, j, isp)];
          // app.f90:408-411 (electrons co-located with ions: position keyed on species 0)
          pu[0] = (w->nxgs + (w->nxge - w->nxgs + 1) * (ii - 0.5) / npr) * w->delx;
          pu[1] = (j + rng_uniform(seed, 0, pid, 0)) * w->delx;
          // app.f90:426-428 with Box-Muller as in utils/wuming_utils.f90:71-90
          const double u1 = rng_uniform(seed, isp, pid, 1), u2 = rng_uniform(seed, isp, pid, 2);
          const double u3 = rng_uniform(seed, isp, pid, 3), u4 = rng_uniform(seed, isp, pid, 4);
          const double rr1 = std::sqrt(-2 * std::log(1 - u1) + 1.0e-30);
          const double rr2 = std::sqrt(-2 * std::log(1 - u3) + 1.0e-30);
          const double s = sd[(isp - 1) & 1];
          pu[2] = s * (rr1 * std::sin(2 * PI * u2));
          pu[3] = s * (rr1 * std::cos(2 * PI * u2));
          pu[4] = t_ani * s * (rr2 * std::sin(2 * PI * u4));
          const int64_t neg = -(int64_t)pid;  // app.f90:468
          std::memcpy(&pu[5], &neg, 8);
        }
      }
    }
    k.gp = k.up;  // app.f90:362
  }
}

uint64_t orc_rng_hash(uint64_t seed, int isp, uint64_t gid, int stream) { return rng_hash(seed, isp, gid, stream); }
double orc_rng_uniform(uint64_t seed, int isp, uint64_t gid, int stream) { return rng_uniform(seed, isp, gid, stream); }

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// OpenMP team size of the following calls (bench.py sets it explicitly: launchers such as torch.distributed.run export
// OMP_NUM_THREADS=1 to their workers)
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  if (n > 0) omp_set_num_threads(n);
#else
  (void)n;
#endif
}

}  // extern "C"
