// field_kernels.cu -- sm_100a kernels for the implicit FDTD field solve (Hoshino 2013) of
// common/field.f90:66-186 and its conjugate-gradient solver (field.f90:319-461), plus the
// periodic grid boundary routines of common/boundary_periodic.f90:251-568.
//
// All grid arrays keep the reference's AoS layout with two ghost cells per side, so host
// uploads/downloads are plain copies.  Every kernel is a streaming stencil: HBM bound,
// one thread per cell handling all components of that cell (coalesced 24/48-byte records).
// The three CG systems (Bx, By, Bz) are independent and are iterated together, one thread
// handling the three components of its cell; each keeps its own convergence state, so the
// iteration counts equal those of the reference's sequential l=1,2,3 loop.
//
// This file is compiled with -fmad=false: expressions keep the reference's operation order
// without FMA contraction, so the only difference to the CPU path is the order of the
// global dot-product sums.
#include <cstddef>
#include <cstdlib>

#include "kernels.h"

namespace wm {

struct CgCtl {
  double sumb[3];   // sum b^2
  double sumr[3];   // sum r^2 (before the update of the current iteration)
  double sum2[3];   // sum p.Ap
  double sum1[3];   // sum r^2 after the update
  double eps[3];
  double sum_g[3];
  int active[3];
  int ite[3];
  int stop;         // ite_max reached (field.f90:427-430)
  unsigned ticket[4];
};

__device__ __forceinline__ size_t pidx(const DevParams &P, int li, int lj) {
  return (size_t)(lj + 2) * P.pitch + (li + 2);
}

// cell-centred fields                                                     particle.f90:69-81
__global__ void k_tmpf(const DevParams P, const double *__restrict__ uf, double *__restrict__ tmpf) {
  const int n = (P.nx + 2) * (P.nyl + 2);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int lj = t / (P.nx + 2) - 1, li = t % (P.nx + 2) - 1;
    const double *a = uf + pidx(P, li, lj) * 6;
    const double *ax = a + 6, *ay = a + (size_t)P.pitch * 6, *axy = ay + 6;
    double *o = tmpf + pidx(P, li, lj) * 6;
    o[0] = 0.5 * (+a[0] + ay[0]);
    o[1] = 0.5 * (+a[1] + ax[1]);
    o[2] = 0.25 * (+a[2] + ax[2] + ay[2] + axy[2]);
    o[3] = 0.5 * (+a[3] + ax[3]);
    o[4] = 0.5 * (+a[4] + ay[4]);
    o[5] = a[5];
  }
}

// periodic x ghosts by copy, ng ghost columns per side, all rows  boundary_periodic.f90:347-352,563-566
__global__ void k_fill_x(const DevParams P, double *a, int ncomp, int ng) {
  const int n = (P.nyl + 4) * 2 * ng * ncomp;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int c = t % ncomp;
    const int g = (t / ncomp) % (2 * ng);
    const int lj = t / (ncomp * 2 * ng) - 2;
    // g in [0,ng): left ghosts li = -ng+g <- li + nx ; g in [ng,2ng): right ghosts li = nx+g-ng <- li - nx
    const int li = (g < ng) ? (-ng + g) : (P.nx + g - ng);
    const int ls = (g < ng) ? li + P.nx : li - P.nx;
    a[pidx(P, li, lj) * ncomp + c] = a[pidx(P, ls, lj) * ncomp + c];
  }
}

// periodic y ghosts by copy when the ring has one rank           boundary_periodic.f90:263-345,523-561
__global__ void k_fill_y_local(const DevParams P, double *a, int ncomp, int ng) {
  const int w = P.pitch * ncomp;
  const int n = 2 * ng * w;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int e = t % w;
    const int g = t / w;
    const int lj = (g < ng) ? (-ng + g) : (P.nyl + g - ng);
    const int ls = (g < ng) ? lj + P.nyl : lj - P.nyl;
    a[(size_t)(lj + 2) * w + e] = a[(size_t)(ls + 2) * w + e];
  }
}

// df x boundary at conducting walls, all rows incl. ghosts   proj/reconnection/boundary_reconnection.f90:349-359
// (one ghost column per side; Bx, Ey, Ez odd about the wall and pinned on it, By, Bz, Ex even)
__global__ void k_wall_x_dfield(const DevParams P, double *df) {
  const int n = P.nyl + 4;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    double *row = df + pidx(P, 0, t - 2) * 6;  // row[6*li + c], li = 0 is nxs
    const int e = P.nx - 1;                    // nxe
    row[6 * (-1) + 0] = -row[6 * 0 + 0];
    for (int c = 1; c <= 3; c++) row[6 * (-1) + c] = row[6 * 1 + c];
    for (int c = 4; c <= 5; c++) row[6 * (-1) + c] = -row[6 * 0 + c];
    if (P.bc == WM_BC_SHOCK) {  // proj/shock/boundary_shock.f90:401-403
      for (int c = 0; c <= 5; c++) row[6 * (e + 1) + c] = 0.0;
    } else {                    // proj/reconnection/boundary_reconnection.f90:355-357
      row[6 * e + 0] = -row[6 * (e - 1) + 0];
      for (int c = 1; c <= 3; c++) row[6 * (e + 1) + c] = row[6 * (e - 1) + c];
      for (int c = 4; c <= 5; c++) row[6 * e + c] = -row[6 * (e - 1) + c];
    }
  }
}

// uj: x fold then copy back, all rows                               boundary_periodic.f90:495-506
__global__ void k_fold_x(const DevParams P, double *uj) {
  const int n = (P.nyl + 4) * 3;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int c = t % 3, lj = t / 3 - 2;
    double *row = uj + pidx(P, 0, lj) * 3 + c;  // row[3*li]
    const int nx = P.nx;
    row[3 * (nx - 2)] = row[3 * (nx - 2)] + row[3 * (-2)];
    row[3 * (nx - 1)] = row[3 * (nx - 1)] + row[3 * (-1)];
    row[3 * 0] = row[3 * 0] + row[3 * nx];
    row[3 * 1] = row[3 * 1] + row[3 * (nx + 1)];
    row[3 * (-2)] = row[3 * (nx - 2)];
    row[3 * (-1)] = row[3 * (nx - 1)];
    row[3 * nx] = row[3 * 0];
    row[3 * (nx + 1)] = row[3 * 1];
  }
}

// uj: y fold then ghost refresh when the ring has one rank          boundary_periodic.f90:369-493
__global__ void k_fold_y_local(const DevParams P, double *uj) {
  const int w = P.pitch * 3;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < w; e += gridDim.x * blockDim.x) {
    auto R = [&](int lj) -> double & { return uj[(size_t)(lj + 2) * w + e]; };
    const int nyl = P.nyl;
    R(nyl - 2) = R(nyl - 2) + R(-2);
    R(nyl - 1) = R(nyl - 1) + R(-1);
    R(0) = R(0) + R(nyl);
    R(1) = R(1) + R(nyl + 1);
    R(nyl) = R(0);
    R(nyl + 1) = R(1);
    R(-2) = R(nyl - 2);
    R(-1) = R(nyl - 1);
  }
}

// mom: x fold of the single ghost column, all rows incl. ghosts   boundary_periodic.f90:579-584
__global__ void k_mom_fold_x(const DevParams P, double *mom) {
  const int nxp = P.nx + 2, nyp = P.nyl + 2;  // nxgs-1:nxge+1
  const int n = P.nsp * nyp * 7;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int m = t % 7, row = t / 7;  // row = isp*nyp + j
    double *r = mom + (size_t)row * nxp * 7 + m;
    if (P.bc == WM_BC_PERIODIC) {
      r[7 * 1] = r[7 * 1] + r[7 * (P.nx + 1)];  // nxgs += nxge+1
      r[7 * P.nx] = r[7 * P.nx] + r[7 * 0];     // nxge += nxgs-1
    } else {  // walls: boundary_reconnection.f90:590-595
      r[7 * 1] = r[7 * 1] + r[7 * 0];                     // nxgs += nxgs-1
      r[7 * P.nx] = r[7 * P.nx] + r[7 * (P.nx + 1)];      // nxge += nxge+1
    }
  }
}

// mom: y fold when the ring has one rank (the neighbour is this rank)    boundary_periodic.f90:586-633
__global__ void k_mom_fold_y_local(const DevParams P, double *mom) {
  const int nxp = P.nx + 2, nyp = P.nyl + 2;  // nxgs-1:nxge+1
  const int w = nxp * 7;
  const int n = P.nsp * w;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int isp = t / w, e = t - isp * w;
    double *b = mom + (size_t)isp * nyp * w + e;
    const double lo = b[0], hi = b[(size_t)(P.nyl + 1) * w];
    b[(size_t)P.nyl * w] = b[(size_t)P.nyl * w] + lo;  // nye += (nup's) nys-1
    b[(size_t)1 * w] = b[(size_t)1 * w] + hi;          // nys += (ndown's) nye+1
  }
}

__global__ void k_add(double *__restrict__ dst, const double *__restrict__ src, long long n) {
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
    dst[t] = dst[t] + src[t];
}

// right-hand side of the delta-B systems                                  field.f90:125-146
__global__ void k_rhs(const DevParams P, const double *__restrict__ uf, const double *__restrict__ uj,
                      double *__restrict__ gkl) {
  const int n = P.nx * P.nyl;
  const double f1 = P.f1, f2 = P.f2, f3 = P.f3;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int lj = t / P.nx, li = t % P.nx;
    const size_t o = pidx(P, li, lj);
    const double *c = uf + o * 6, *xm = c - 6, *xp = c + 6, *ym = c - (size_t)P.pitch * 6, *yp = c + (size_t)P.pitch * 6;
    const double *j = uj + o * 3, *jxm = j - 3, *jym = j - (size_t)P.pitch * 3;
    double *g = gkl + o * 3;
    g[0] = +f2 * (+ym[0] + xm[0] - 4. * c[0] + xp[0] + yp[0] + f3 * (-jym[2] + j[2])) - f1 * (-ym[5] + c[5]);
    g[1] = +f2 * (+ym[1] + xm[1] - 4. * c[1] + xp[1] + yp[1] - f3 * (-jxm[2] + j[2])) + f1 * (-xm[5] + c[5]);
    g[2] = +f2 * (+ym[2] + xm[2] - 4. * c[2] + xp[2] + yp[2] + f3 * (-jxm[1] + j[1] + jym[0] - j[0])) -
           f1 * (-xm[4] + c[4] + ym[3] - c[3]);
  }
}

// ---------------------------------------------------------------- CG (field.f90:319-461)
// block partial sums -> red[block*8 + k]; the last block to arrive adds them in a fixed order (thread t takes the
// blocks t, t + 256, ...; then lanes, then warps), so the result is deterministic run to run.  The final sum is
// done by the whole block: a serial loop over the partials was a 20 us tail on every reducing kernel.
template <int NV>
__device__ __forceinline__ bool block_reduce_store(double (&v)[NV], double *red, unsigned *ticket, double *out) {
  __shared__ double s_w[8][NV];
  __shared__ bool s_last;
#pragma unroll
  for (int k = 0; k < NV; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; k++) s_w[wid][k] = v[k];
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < NV; k++) {
      double t = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) t += s_w[w][k];
      red[blockIdx.x * 8 + k] = t;
    }
    __threadfence();
    const unsigned tk = atomicAdd(ticket, 1u);
    s_last = (tk == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    __threadfence();
    double t[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) t[k] = 0.0;
    for (unsigned b = threadIdx.x; b < gridDim.x; b += blockDim.x) {
#pragma unroll
      for (int k = 0; k < NV; k++) t[k] += __ldcg(&red[b * 8 + k]);
    }
#pragma unroll
    for (int k = 0; k < NV; k++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) t[k] += __shfl_xor_sync(0xffffffffu, t[k], o);
    }
    __syncthreads();  // s_w is reused
    if (lane == 0)
#pragma unroll
      for (int k = 0; k < NV; k++) s_w[wid][k] = t[k];
    __syncthreads();
    if (threadIdx.x < NV) {
      double tt = 0.0;
      for (int w = 0; w < (int)(blockDim.x >> 5); w++) tt += s_w[w][threadIdx.x];
      out[threadIdx.x] = tt;
    }
    if (threadIdx.x == 0) *ticket = 0;
  }
  return s_last;
}

// neighbour rows with periodic wrap when this rank owns the whole ring, columns always wrap
__device__ __forceinline__ void nbr(const DevParams &P, int li, int lj, size_t &xm, size_t &xp, size_t &ym, size_t &yp) {
  const int lim = (li == 0) ? P.nx - 1 : li - 1;
  const int lip = (li == P.nx - 1) ? 0 : li + 1;
  int ljm = lj - 1, ljp = lj + 1;
  if (P.nsize == 1) {
    if (ljm < 0) ljm = P.nyl - 1;
    if (ljp >= P.nyl) ljp = 0;
  }
  xm = pidx(P, lim, lj);
  xp = pidx(P, lip, lj);
  ym = pidx(P, li, ljm);
  yp = pidx(P, li, ljp);
}

// x neighbours of component l of a CG vector.  Periodic: the wrapped column.  Conducting walls
// (proj/reconnection/boundary_reconnection.f90:557-577): left ghost = -phi(nxs) for l = 1 (Bx, odd) and
// phi(nxs+1) for l = 2,3 (even); right ghost = 0.
__device__ __forceinline__ double nb_xm(const DevParams &P, const double *a, int l, int li, size_t o, size_t xm) {
  if (P.bc != WM_BC_PERIODIC && li == 0) return (l == 0) ? -a[o * 3 + l] : a[(o + 1) * 3 + l];
  return a[xm * 3 + l];
}
__device__ __forceinline__ double nb_xp(const DevParams &P, const double *a, int l, int li, size_t xp) {
  if (P.bc != WM_BC_PERIODIC && li == P.nx - 1) return 0.0;
  return a[xp * 3 + l];
}

// phi <- df(l), b <- f5*gkl(l) (in place), sum b^2                          field.f90:349-362
__global__ void __launch_bounds__(256) k_cg_init(const DevParams P, const double *__restrict__ df, double *gkl,
                                                 double *__restrict__ phi, double *red, CgCtl *ctl) {
  const int n = P.nx * P.nyl;
  double s[3] = {0.0, 0.0, 0.0};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int lj = t / P.nx, li = t % P.nx;
    const size_t o = pidx(P, li, lj);
#pragma unroll
    for (int l = 0; l < 3; l++) {
      phi[o * 3 + l] = df[o * 6 + l];
      const double b = P.f5 * gkl[o * 3 + l];
      gkl[o * 3 + l] = b;
      s[l] = s[l] + b * b;
    }
  }
  block_reduce_store<3>(s, red, &ctl->ticket[0], ctl->sumb);
}

// r <- b + N4 phi - f4 phi ; p <- r ; sum r^2                               field.f90:370-383
__global__ void __launch_bounds__(256) k_cg_resid0(const DevParams P, const double *__restrict__ b,
                                                   const double *__restrict__ phi, double *__restrict__ r,
                                                   double *__restrict__ p, double *red, CgCtl *ctl) {
  const int n = P.nx * P.nyl;
  double s[3] = {0.0, 0.0, 0.0};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int lj = t / P.nx, li = t % P.nx;
    const size_t o = pidx(P, li, lj);
    size_t xm, xp, ym, yp;
    nbr(P, li, lj, xm, xp, ym, yp);
#pragma unroll
    for (int l = 0; l < 3; l++) {
      const double rr = b[o * 3 + l] + phi[ym * 3 + l] + nb_xm(P, phi, l, li, o, xm) - P.f4 * phi[o * 3 + l] +
                        nb_xp(P, phi, l, li, xp) + phi[yp * 3 + l];
      r[o * 3 + l] = rr;
      p[o * 3 + l] = rr;
      s[l] = s[l] + rr * rr;
    }
  }
  block_reduce_store<3>(s, red, &ctl->ticket[0], ctl->sumr);
}

// eps, first loop test                                                       field.f90:364,385-387
__global__ void k_cg_begin(CgCtl *ctl) {
  const int l = threadIdx.x;
  if (l < 3) {
    const double err = 1e-6;
    ctl->eps[l] = sqrt(ctl->sumb[l]) * err;
    ctl->ite[l] = 0;
    int act = 0;
    if (sqrt(ctl->sumr[l]) > ctl->eps[l]) {
      ctl->sum_g[l] = ctl->sumb[l];  // first test compares sum(b^2), not its sqrt
      act = ctl->sum_g[l] > ctl->eps[l];
    }
    ctl->active[l] = act;
    ctl->sum1[l] = 0.0;  // bv of the first fused p update (k_cg_pap): p <- r + 0 p
  }
  if (l == 0) ctl->stop = 0;
}

// ap <- f4 p - N4 p ; sum r^2, sum p.ap                                      field.f90:395-413
__global__ void __launch_bounds__(256) k_cg_ap(const DevParams P, const double *__restrict__ p,
                                               const double *__restrict__ r, double *__restrict__ ap, double *red,
                                               CgCtl *ctl) {
  const int a0 = ctl->active[0], a1 = ctl->active[1], a2 = ctl->active[2];
  if (!(a0 | a1 | a2)) return;
  const int act[3] = {a0, a1, a2};
  const int n = P.nx * P.nyl;
  double s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int lj = t / P.nx, li = t % P.nx;
    const size_t o = pidx(P, li, lj);
    size_t xm, xp, ym, yp;
    nbr(P, li, lj, xm, xp, ym, yp);
#pragma unroll
    for (int l = 0; l < 3; l++) {
      if (!act[l]) continue;
      const double pc = p[o * 3 + l];
      const double av = -p[ym * 3 + l] - nb_xm(P, p, l, li, o, xm) + P.f4 * pc - nb_xp(P, p, l, li, xp) - p[yp * 3 + l];
      ap[o * 3 + l] = av;
      const double rr = r[o * 3 + l];
      s[l] = s[l] + rr * rr;
      s[3 + l] = s[3 + l] + pc * av;
    }
  }
  // out: sumr[0..2] and sum2[0..2] are adjacent in CgCtl
  block_reduce_store<6>(s, red, &ctl->ticket[1], ctl->sumr);
}

// phi += av p ; r -= av ap ; sum r^2                                          field.f90:415-441
__global__ void __launch_bounds__(256) k_cg_update(const DevParams P, const double *__restrict__ p,
                                                   const double *__restrict__ ap, double *__restrict__ phi,
                                                   double *__restrict__ r, double *red, CgCtl *ctl) {
  const int a0 = ctl->active[0], a1 = ctl->active[1], a2 = ctl->active[2];
  if (!(a0 | a1 | a2)) return;
  const int act[3] = {a0, a1, a2};
  double av[3];
#pragma unroll
  for (int l = 0; l < 3; l++) av[l] = act[l] ? ctl->sumr[l] / ctl->sum2[l] : 0.0;
  const int n = P.nx * P.nyl;
  double s[3] = {0.0, 0.0, 0.0};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int lj = t / P.nx, li = t % P.nx;
    const size_t o = pidx(P, li, lj);
#pragma unroll
    for (int l = 0; l < 3; l++) {
      if (!act[l]) continue;
      phi[o * 3 + l] = phi[o * 3 + l] + av[l] * p[o * 3 + l];
      const double rr = r[o * 3 + l] - av[l] * ap[o * 3 + l];
      r[o * 3 + l] = rr;
      s[l] = s[l] + rr * rr;
    }
  }
  block_reduce_store<3>(s, red, &ctl->ticket[2], ctl->sum1);
}

// p <- r + bv p ; then the loop control for the next iteration               field.f90:426-450,387
__global__ void __launch_bounds__(256) k_cg_pupdate(const DevParams P, const double *__restrict__ r,
                                                    double *__restrict__ p, CgCtl *ctl) {
  __shared__ bool s_last;
  const int a0 = ctl->active[0], a1 = ctl->active[1], a2 = ctl->active[2];
  if (!(a0 | a1 | a2)) return;
  const int act[3] = {a0, a1, a2};
  double bv[3];
#pragma unroll
  for (int l = 0; l < 3; l++) bv[l] = act[l] ? ctl->sum1[l] / ctl->sumr[l] : 0.0;
  const int n = P.nx * P.nyl;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int lj = t / P.nx, li = t % P.nx;
    const size_t o = pidx(P, li, lj);
#pragma unroll
    for (int l = 0; l < 3; l++) {
      if (!act[l]) continue;
      p[o * 3 + l] = r[o * 3 + l] + bv[l] * p[o * 3 + l];
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    s_last = (atomicAdd(&ctl->ticket[3], 1u) == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last && threadIdx.x < 3) {
    // every other block has finished reading ctl: safe to advance the loop state in place
    const int l = threadIdx.x;
    if (act[l]) {
      ctl->ite[l] = ctl->ite[l] + 1;
      ctl->sum_g[l] = sqrt(ctl->sumr[l]);  // residual *before* this iteration's update (:426)
      if (ctl->ite[l] >= 100) {            // ite_max (:427)
        ctl->stop = 1;
        ctl->active[l] = 0;
      } else {
        ctl->active[l] = ctl->sum_g[l] > ctl->eps[l];
      }
    }
  }
  if (s_last && threadIdx.x == 0) ctl->ticket[3] = 0;
}


// ---- two-kernel CG iteration (default): the p update of field.f90:446-450 is folded into the next A p.
// k_cg_pap: pn <- r + bv p (bv = sum1/sumr of the previous iteration, 0 in the first), written to a second
// buffer because the stencil reads p of the neighbours; ap <- f4 pn - N4 pn with pn of the four neighbours
// recomputed from r and p (identical expression, so the values are those of the three-kernel form);
// sums r^2 and pn.ap.  With more than one rank the rows nys-1 and nye+1 of pn are computed here as well, from
// the exchanged halo rows of r (instead of exchanging p every iteration, set_boundary_phi at field.f90:392).
__global__ void __launch_bounds__(256) k_cg_pap(const DevParams P, const double *__restrict__ r, const double *__restrict__ p,
                                                double *__restrict__ pn, double *__restrict__ ap, double *red, CgCtl *ctl) {
  const int a0 = ctl->active[0], a1 = ctl->active[1], a2 = ctl->active[2];
  if (!(a0 | a1 | a2)) return;
  const int act[3] = {a0, a1, a2};
  double bv[3];
#pragma unroll
  for (int l = 0; l < 3; l++) bv[l] = act[l] ? ctl->sum1[l] / ctl->sumr[l] : 0.0;
  const int n = P.nx * P.nyl;
  const int nh = (P.nsize > 1) ? 2 * P.nx : 0;  // halo rows of pn
  double s[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n + nh; t += gridDim.x * blockDim.x) {
    if (t >= n) {
      const int h = t - n, li = h % P.nx, lj = (h < P.nx) ? -1 : P.nyl;
      const size_t o = pidx(P, li, lj);
#pragma unroll
      for (int l = 0; l < 3; l++)
        if (act[l]) pn[o * 3 + l] = r[o * 3 + l] + bv[l] * p[o * 3 + l];
      continue;
    }
    const int lj = t / P.nx, li = t % P.nx;
    const size_t o = pidx(P, li, lj);
    size_t xm, xp, ym, yp;
    nbr(P, li, lj, xm, xp, ym, yp);
#pragma unroll
    for (int l = 0; l < 3; l++) {
      if (!act[l]) continue;
      const double b = bv[l];
      const double rr = r[o * 3 + l];
      const double pc = rr + b * p[o * 3 + l];
      pn[o * 3 + l] = pc;
      const double pym = r[ym * 3 + l] + b * p[ym * 3 + l], pyp = r[yp * 3 + l] + b * p[yp * 3 + l];
      double pxm, pxp;
      if (P.bc != WM_BC_PERIODIC && li == 0)  // conducting wall: see nb_xm
        pxm = (l == 0) ? -pc : r[(o + 1) * 3 + l] + b * p[(o + 1) * 3 + l];
      else
        pxm = r[xm * 3 + l] + b * p[xm * 3 + l];
      if (P.bc != WM_BC_PERIODIC && li == P.nx - 1)
        pxp = 0.0;
      else
        pxp = r[xp * 3 + l] + b * p[xp * 3 + l];
      const double av = -pym - pxm + P.f4 * pc - pxp - pyp;
      ap[o * 3 + l] = av;
      s[l] = s[l] + rr * rr;
      s[3 + l] = s[3 + l] + pc * av;
    }
  }
  block_reduce_store<6>(s, red, &ctl->ticket[1], ctl->sumr);
}

// phi += av p ; r -= av ap ; sum r^2 ; then the loop control of field.f90:426-430,387 (it only needs sumr, the
// residual before this update)
__global__ void __launch_bounds__(256) k_cg_update2(const DevParams P, const double *__restrict__ p,
                                                    const double *__restrict__ ap, double *__restrict__ phi,
                                                    double *__restrict__ r, double *red, CgCtl *ctl) {
  const int a0 = ctl->active[0], a1 = ctl->active[1], a2 = ctl->active[2];
  if (!(a0 | a1 | a2)) return;
  const int act[3] = {a0, a1, a2};
  double av[3];
#pragma unroll
  for (int l = 0; l < 3; l++) av[l] = act[l] ? ctl->sumr[l] / ctl->sum2[l] : 0.0;
  const int n = P.nx * P.nyl;
  double s[3] = {0.0, 0.0, 0.0};
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int lj = t / P.nx, li = t % P.nx;
    const size_t o = pidx(P, li, lj);
#pragma unroll
    for (int l = 0; l < 3; l++) {
      if (!act[l]) continue;
      phi[o * 3 + l] = phi[o * 3 + l] + av[l] * p[o * 3 + l];
      const double rr = r[o * 3 + l] - av[l] * ap[o * 3 + l];
      r[o * 3 + l] = rr;
      s[l] = s[l] + rr * rr;
    }
  }
  const bool last = block_reduce_store<3>(s, red, &ctl->ticket[2], ctl->sum1);
  if (last && threadIdx.x < 3) {
    // every other block has read ctl (it had written its partial sums before the ticket): advance the loop state
    const int l = threadIdx.x;
    if (act[l]) {
      ctl->ite[l] = ctl->ite[l] + 1;
      ctl->sum_g[l] = sqrt(ctl->sumr[l]);  // residual *before* this iteration's update (:426)
      if (ctl->ite[l] >= 100) {            // ite_max (:427)
        ctl->stop = 1;
        ctl->active[l] = 0;
      } else {
        ctl->active[l] = ctl->sum_g[l] > ctl->eps[l];
      }
    }
  }
}

// df(l) <- phi on the interior                                               field.f90:455-457
__global__ void k_cg_finish(const DevParams P, const double *__restrict__ phi, double *__restrict__ df) {
  const int n = P.nx * P.nyl;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int lj = t / P.nx, li = t % P.nx;
    const size_t o = pidx(P, li, lj);
#pragma unroll
    for (int l = 0; l < 3; l++) df[o * 6 + l] = phi[o * 3 + l];
  }
}

// delta-E                                                                    field.f90:154-171
__global__ void k_efield(const DevParams P, const double *__restrict__ uf, const double *__restrict__ uj, double *df) {
  const int n = P.nx * P.nyl;
  const double f1 = P.f1, gfac = P.gfac, pi4dt = P.pi4dt;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int lj = t / P.nx, li = t % P.nx;
    const size_t o = pidx(P, li, lj);
    const double *u = uf + o * 6, *ux = u + 6, *uy = u + (size_t)P.pitch * 6;
    double *d = df + o * 6;
    const double *dx = d + 6, *dy = d + (size_t)P.pitch * 6;
    const double *j = uj + o * 3;
    const double e4 = +f1 * (+gfac * (-d[2] + dy[2]) + (-u[2] + uy[2])) - pi4dt * j[0];
    const double e5 = -f1 * (+gfac * (-d[2] + dx[2]) + (-u[2] + ux[2])) - pi4dt * j[1];
    const double e6 = +f1 * (+gfac * (-d[1] + dx[1] + d[0] - dy[0]) + (-u[1] + ux[1] + u[0] - uy[0])) - pi4dt * j[2];
    d[3] = e4;
    d[4] = e5;
    d[5] = e6;
  }
}

// uf += df over the whole padded array                                        field.f90:176-184
// over the columns nxs-2 .. nxe+2 of the ACTIVE range (P.nx = nxe - nxs + 1 here) and all rows incl. ghosts, as the reference
// loops: beyond a shock box that has not grown to nxge yet, uf is left alone
__global__ void k_update_uf(const DevParams P, double *__restrict__ uf, const double *__restrict__ df) {
  const long long w = (long long)(P.nx + 4) * 6, n = w * (P.nyl + 4);
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x) {
    const long long row = t / w, e = t - row * w;
    const size_t o = (size_t)row * P.pitch * 6 + e;
    uf[o] = uf[o] + df[o];
  }
}

// sum B^2, sum E^2 over the interior (app.f90:521-528): partial[block*2 + {0:B,1:E}]
__global__ void __launch_bounds__(256) k_field_energy(const DevParams P, const double *__restrict__ uf, double *partial) {
  __shared__ double s_w[8][2];
  const int n = P.nx * P.nyl;
  double sb = 0.0, se = 0.0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int lj = t / P.nx, li = t % P.nx;
    const double *u = uf + pidx(P, li, lj) * 6;
    sb = sb + u[0] * u[0] + u[1] * u[1] + u[2] * u[2];
    se = se + u[3] * u[3] + u[4] * u[4] + u[5] * u[5];
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    sb += __shfl_xor_sync(0xffffffffu, sb, o);
    se += __shfl_xor_sync(0xffffffffu, se, o);
  }
  if ((threadIdx.x & 31) == 0) {
    s_w[threadIdx.x >> 5][0] = sb;
    s_w[threadIdx.x >> 5][1] = se;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double tb = 0.0, te = 0.0;
    for (int k = 0; k < 8; k++) {
      tb += s_w[k][0];
      te += s_w[k][1];
    }
    partial[blockIdx.x * 2] = tb;
    partial[blockIdx.x * 2 + 1] = te;
  }
}

__global__ void k_sum_ranks(const RankPtrs rp, int nranks, int n, int is_double, double *out) {
  const int k = threadIdx.x;
  if (k >= n) return;
  if (is_double) {
    double s = 0.0;
    for (int q = 0; q < nranks; q++) s += static_cast<const volatile double *>(rp.p[q])[k];
    out[k] = s;
  } else {
    int m = static_cast<const volatile int *>(rp.p[0])[k];
    for (int q = 1; q < nranks; q++) m = min(m, static_cast<const volatile int *>(rp.p[q])[k]);
    reinterpret_cast<int *>(out)[k] = m;
  }
}
void launch_sum_ranks(const RankPtrs &rp, int nranks, int n, bool is_double, double *out, cudaStream_t st) {
  k_sum_ranks<<<1, 32, 0, st>>>(rp, nranks, n, is_double ? 1 : 0, out);
}

// ---- measured FP64 peak: 8 independent DFMA chains per thread, 16 warps per SM x 4 CTAs: the pipe's own rate
//      ("to be measured by the builder with a DFMA loop")
__global__ void __launch_bounds__(512) k_fp64_peak(double *out, int n, double a, double b) {
  double v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = threadIdx.x * 1e-3 + i;
  for (int it = 0; it < n; it++) {
#pragma unroll
    for (int i = 0; i < 8; i++) asm volatile("fma.rn.f64 %0, %0, %1, %2;" : "+d"(v[i]) : "d"(a), "d"(b));
  }
  double s = 0.0;
#pragma unroll
  for (int i = 0; i < 8; i++) s += v[i];
  out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}
void launch_fp64_peak(double *out, int nblocks, int n, cudaStream_t st) { k_fp64_peak<<<nblocks, 512, 0, st>>>(out, n, 0.999, 1e-3); }

// ---------------------------------------------------------------- launch wrappers
// grid of the reducing CG kernels: RED_BLOCKS by default (4 per SM); WM_CGBLOCKS overrides it up to the allocated maximum
static int cg_blocks() {
  static int nb = 0;
  if (!nb) {
    nb = RED_BLOCKS;
    if (const char *v = getenv("WM_CGBLOCKS")) nb = atoi(v);
    if (nb < 1) nb = 1;
    if (nb > RED_BLOCKS_MAX) nb = RED_BLOCKS_MAX;
  }
  return nb;
}
static inline int gblocks(long long n) { return (int)((n + 255) / 256 < RED_BLOCKS * 2 ? (n + 255) / 256 : RED_BLOCKS * 2); }

size_t cgctl_bytes() { return sizeof(CgCtl); }
size_t cgctl_active_offset() { return offsetof(CgCtl, active); }
size_t cgctl_sums_offset(int which) {
  return which == 0 ? offsetof(CgCtl, sumb) : which == 1 ? offsetof(CgCtl, sumr) : which == 2 ? offsetof(CgCtl, sum2) : offsetof(CgCtl, sum1);
}

void launch_tmpf(const DevParams &P, const double *uf, double *tmpf, cudaStream_t st) {
  k_tmpf<<<gblocks((long long)(P.nx + 2) * (P.nyl + 2)), 256, 0, st>>>(P, uf, tmpf);
}
void launch_fill_x(const DevParams &P, double *a, int ncomp, int ng, cudaStream_t st) {
  if (P.bc != WM_BC_PERIODIC) {  // only df takes this path (the CG vectors apply their wall rule in place)
    k_wall_x_dfield<<<gblocks((long long)P.nyl + 4), 256, 0, st>>>(P, a);
    return;
  }
  k_fill_x<<<gblocks((long long)(P.nyl + 4) * 2 * ng * ncomp), 256, 0, st>>>(P, a, ncomp, ng);
}
void launch_fill_y_local(const DevParams &P, double *a, int ncomp, int ng, cudaStream_t st) {
  k_fill_y_local<<<gblocks((long long)2 * ng * P.pitch * ncomp), 256, 0, st>>>(P, a, ncomp, ng);
}
void launch_fold_x(const DevParams &P, double *uj, cudaStream_t st) {
  k_fold_x<<<gblocks((long long)(P.nyl + 4) * 3), 256, 0, st>>>(P, uj);
}
void launch_fold_y_local(const DevParams &P, double *uj, cudaStream_t st) {
  k_fold_y_local<<<gblocks((long long)P.pitch * 3), 256, 0, st>>>(P, uj);
}
void launch_mom_fold_x(const DevParams &P, double *mom, cudaStream_t st) {
  k_mom_fold_x<<<gblocks((long long)P.nsp * (P.nyl + 2) * 7), 256, 0, st>>>(P, mom);
}
void launch_mom_fold_y_local(const DevParams &P, double *mom, cudaStream_t st) {
  k_mom_fold_y_local<<<gblocks((long long)P.nsp * (P.nx + 2) * 7), 256, 0, st>>>(P, mom);
}
void launch_add_rows(double *dst, const double *src, long long n, cudaStream_t st) {
  k_add<<<gblocks(n), 256, 0, st>>>(dst, src, n);
}
void launch_rhs(const DevParams &P, const FieldBufs &f, cudaStream_t st) {
  k_rhs<<<gblocks((long long)P.ncell), 256, 0, st>>>(P, f.uf, f.uj, f.gkl);
}
void launch_cg_init(const DevParams &P, const FieldBufs &f, cudaStream_t st) {
  k_cg_init<<<cg_blocks(), 256, 0, st>>>(P, f.df, f.gkl, f.phi, f.red, (CgCtl *)f.cgstate);
}
void launch_cg_resid0(const DevParams &P, const FieldBufs &f, cudaStream_t st) {
  k_cg_resid0<<<cg_blocks(), 256, 0, st>>>(P, f.gkl, f.phi, f.r, f.p, f.red, (CgCtl *)f.cgstate);
}
void launch_cg_begin(const DevParams &, const FieldBufs &f, int, cudaStream_t st) {
  k_cg_begin<<<1, 32, 0, st>>>((CgCtl *)f.cgstate);
}
void launch_cg_ap(const DevParams &P, const FieldBufs &f, cudaStream_t st) {
  k_cg_ap<<<cg_blocks(), 256, 0, st>>>(P, f.p, f.r, f.ap, f.red, (CgCtl *)f.cgstate);
}
void launch_cg_update(const DevParams &P, const FieldBufs &f, cudaStream_t st) {
  k_cg_update<<<cg_blocks(), 256, 0, st>>>(P, f.p, f.ap, f.phi, f.r, f.red, (CgCtl *)f.cgstate);
}
void launch_cg_pap(const DevParams &P, const FieldBufs &f, const double *p_in, double *p_out, cudaStream_t st) {
  k_cg_pap<<<cg_blocks(), 256, 0, st>>>(P, f.r, p_in, p_out, f.ap, f.red, (CgCtl *)f.cgstate);
}
void launch_cg_update2(const DevParams &P, const FieldBufs &f, const double *p, cudaStream_t st) {
  k_cg_update2<<<cg_blocks(), 256, 0, st>>>(P, p, f.ap, f.phi, f.r, f.red, (CgCtl *)f.cgstate);
}
void launch_cg_pupdate(const DevParams &P, const FieldBufs &f, cudaStream_t st) {
  k_cg_pupdate<<<cg_blocks(), 256, 0, st>>>(P, f.r, f.p, (CgCtl *)f.cgstate);
}
void launch_cg_finish(const DevParams &P, const FieldBufs &f, cudaStream_t st) {
  k_cg_finish<<<gblocks((long long)P.ncell), 256, 0, st>>>(P, f.phi, f.df);
}
void launch_efield(const DevParams &P, const FieldBufs &f, cudaStream_t st) {
  k_efield<<<gblocks((long long)P.ncell), 256, 0, st>>>(P, f.uf, f.uj, f.df);
}
void launch_update_uf(const DevParams &P, const FieldBufs &f, cudaStream_t st) {
  const long long n = (long long)(P.nx + 4) * (P.nyl + 4) * 6;
  k_update_uf<<<gblocks(n), 256, 0, st>>>(P, f.uf, f.df);
}
void launch_field_energy(const DevParams &P, const double *uf, double *partial, int nblocks, cudaStream_t st) {
  k_field_energy<<<nblocks, 256, 0, st>>>(P, uf, partial);
}

}  // namespace wm
