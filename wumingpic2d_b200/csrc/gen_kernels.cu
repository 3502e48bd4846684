// gen_kernels.cu -- particle / field generators of the applications, on the device: the
// initial conditions of proj/reconnection (Harris current sheet, app.f90:368-456) and proj/shock (drifting Maxwellian,
// app.f90:406-470), and the per-step particle sources of the shock driver, `inject` (app.f90:685-850) and `relocate`
// (app.f90:611-680), which in the reference write new records into the host arrays every step and so defeat residency.
//
// The DISTRIBUTIONS are the reference's, formula by formula; the random numbers are not (the reference seeds from OS
// entropy): every variate is a counter-based hash keyed by (seed, species key, particle id, stream), the generator
// of k_ic_weibel / the oracle (splitmix64), so a run is reproducible on any decomposition.  Normal variates follow
// utils/wuming_utils.f90:71-90 (Box-Muller: sqrt(-2 log(1 - u1) + 1e-30) sin(2 pi u2)), one pair of uniforms per variate.
//
// Records are written as tight AoS rows (x, y, ux, uy, uz, id) into a staging buffer; the host layer then buckets them like
// sort__bucket (initial conditions) or appends them to their cells' segments (inject / relocate).
#include <cstdint>

#include "kernels.h"

namespace wm {

namespace {

__device__ __forceinline__ uint64_t splitmix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double uni(uint64_t seed, int key, uint64_t gid, int stream) {
  uint64_t k = splitmix(seed ^ (0xD1B54A32D192ED03ULL * (uint64_t)(key + 1)));
  k = splitmix(k + gid);
  k = splitmix(k + 0x8CB92BA72F3D8DD7ULL * (uint64_t)(stream + 1));
  return ((double)(k >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
__device__ __forceinline__ double normal(uint64_t seed, int key, uint64_t gid, int stream) {  // wuming_utils.f90:71-90
  const double PI = 3.14159265358979323846;
  const double u1 = uni(seed, key, gid, stream), u2 = uni(seed, key, gid, stream + 1);
  return sqrt(-2 * log(1 - u1) + 1.0e-30) * sin(2 * PI * u2);
}
__device__ __forceinline__ void put(double *rec, double x, double y, double ux, double uy, double uz, long long id) {
  rec[0] = x;
  rec[1] = y;
  rec[2] = ux;
  rec[3] = uy;
  rec[4] = uz;
  rec[5] = __longlong_as_double(id);
}

// initial velocity profile of the shock set-up                                  proj/shock/app.f90:881-893
__device__ __forceinline__ double vprofile(const GenParams &g, double x) {
  const double x0 = g.l_damp + g.nxgs * g.delx, xs = g.l_damp * 0.1;
  return 0.5 * g.v0 * (1 + tanh((x - x0) / xs));
}
// Maxwellian in the fluid rest frame, then the Lorentz transform to the lab frame     proj/shock/app.f90:447-466
__device__ __forceinline__ void boosted_maxwellian(const GenParams &g, uint64_t seed, int key, uint64_t gid, double sd, double x,
                                                   double &ux, double &uy, double &uz) {
  ux = sd * normal(seed, key, gid, 2);
  uy = sd * normal(seed, key, gid, 4);
  uz = sd * normal(seed, key, gid, 6);
  const double v1 = vprofile(g, x);
  const double gam1 = 1 / sqrt(1 - (v1 / g.c) * (v1 / g.c));
  const double gamp = sqrt(1 + (ux * ux + uy * uy + uz * uz) / (g.c * g.c));
  ux = gam1 * (ux + v1 * gamp);
}

}  // namespace

// ---- Harris current sheet                                                    proj/reconnection/app.f90:368-456
// row j holds npr = nbg (nx - 1) + int(2 ncs lcs) pairs: the first ibg = nbg (nx - 3) uniform between the walls, the rest with
// the second-order logistic (sech^2) distribution about x0; ions and electrons at the same positions; drift velocities from
// jz / density split by the temperature ratio.
__global__ void k_gen_harris(const DevParams P, const GenParams g, uint64_t seed, double *stage0, double *stage1) {
  const double PI = 3.14159265358979323846;
  const long long npr = g.npr, ntot = npr * P.nyl;
  const int nxge = g.nxgs + g.nxg - 1, nyge = g.nygs + g.nyg - 1;
  const double x0 = 0.5 * (nxge + g.nxgs) * g.delx, y0 = 0.5 * (nyge - g.nygs) * g.delx;  // app.f90:306-307 (as written)
  const double lcs = g.lcs, b0 = g.b0, e1 = g.e1;
  const double f1 = 1.0 / ((1.0 + g.rtemp) * g.q[0]), f2 = g.rtemp / ((1.0 + g.rtemp) * g.q[1]);
  const double sq2 = (double)sqrtf(2.0f);  // `sqrt(2.)` in the source is single precision
  const double sdi = g.vti / sq2, sde = g.vte / sq2;
  const double span = g.delx * (nxge - g.nxgs - 2);
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < ntot; s += (long long)gridDim.x * blockDim.x) {
    const int lj = (int)(s / npr);
    const long long ii = s - (long long)lj * npr + 1;
    const int gj = P.nys + lj;
    const uint64_t pid = (uint64_t)(gj - g.nygs) * (uint64_t)npr + (uint64_t)ii;
    double x;
    if (ii <= g.ibg) {
      x = (g.nxgs + 1) * g.delx + uni(seed, 0, pid, 0) * span;
    } else {
      double r1 = uni(seed, 0, pid, 0);
      r1 = (2.0 * r1 - 1.0) * tanh(0.5 * span / lcs);
      x = lcs * 0.5 * (log(1.0 + r1) - log(1.0 - r1)) + x0;
    }
    const double y = (double)gj * g.delx + uni(seed, 0, pid, 1) * g.delx;
    const double ch = cosh((x - x0) / lcs), sech2 = 1.0 / (ch * ch);
    const double rho2 = (x - x0) * (x - x0) + (y - y0) * (y - y0), w2 = (2 * lcs) * (2 * lcs);
    const double jz = b0 / (4 * PI * lcs) * sech2 - 2 * e1 * b0 / (4 * PI * lcs) * (1.0 - rho2 / w2) * exp(-rho2 / w2);
    const double dens = g.ncs * sech2 + g.nbg;
    put(stage0 + s * 6, x, y, sdi * normal(seed, 1, pid, 2), sdi * normal(seed, 1, pid, 4),
        sdi * normal(seed, 1, pid, 6) + f1 * jz / dens, -(long long)pid);
    put(stage1 + s * 6, x, y, sde * normal(seed, 2, pid, 2), sde * normal(seed, 2, pid, 4),
        sde * normal(seed, 2, pid, 6) + f2 * jz / dens, -(long long)pid);
  }
}

// uf: Harris field + localized perturbation, whole padded array             proj/reconnection/app.f90:380-411
__global__ void k_field_harris(const DevParams P, const GenParams g, double *uf) {
  const int nxge = g.nxgs + g.nxg - 1, nyge = g.nygs + g.nyg - 1;
  const double x0 = 0.5 * (nxge + g.nxgs) * g.delx, y0 = 0.5 * (nyge - g.nygs) * g.delx;
  const double lcs = g.lcs, b0 = g.b0, e1 = g.e1;
  const int n = P.pitch * (P.nyl + 4);
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
    const int i = g.nxgs - 2 + t % P.pitch, j = P.nys - 2 + t / P.pitch;
    const double x = i * g.delx, y = j * g.delx;
    const double ex = exp(-((x - x0) * (x - x0) + (y - y0) * (y - y0)) / ((2 * lcs) * (2 * lcs)));
    double *u = uf + (size_t)t * 6;
    u[0] = 0.0 + e1 * b0 * ((y - y0) / lcs) * ex;
    u[1] = b0 * tanh((x - x0) / lcs) + (-e1 * b0 * ((x - x0) / lcs) * ex);
    u[2] = u[3] = u[4] = u[5] = 0.0;
  }
}

// ---- shock: initial load (kind 0), relocate (kind 1: n0 pairs per row in the cell nxe - 1), inject (kind 2: rowcnt[lj]
//      pairs per row beyond nxe, moved by (v0 + ux) delt)            proj/shock/app.f90:406-470, 611-680, 685-850
// rowoff[lj] = first record of row lj in the staging buffers (exclusive prefix of the per-row counts), rowoff[nyl] = total
__global__ void k_gen_shock(const DevParams P, const GenParams g, uint64_t seed, int kind, int nxe, const int *__restrict__ rowoff,
                            double *stage0, double *stage1) {
  const int ntot = rowoff[P.nyl];
  for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < ntot; s += gridDim.x * blockDim.x) {
    int lo = 0, hi = P.nyl;  // row of record s: the last row with rowoff <= s
    while (hi - lo > 1) {
      const int mid = (lo + hi) >> 1;
      if (rowoff[mid] <= s) lo = mid; else hi = mid;
    }
    const int lj = lo, nrow = rowoff[lj + 1] - rowoff[lj];
    const int ii = s - rowoff[lj] + 1;
    const int gj = P.nys + lj;
    // id: unique over steps, rows and kinds; negative like every non-tracer id (proj/shock/app.f90:668-669)
    const uint64_t pid = ((uint64_t)g.it << 36) + ((uint64_t)kind << 34) + (uint64_t)(gj - g.nygs) * 4194304ULL + (uint64_t)ii;
    double x;
    if (kind == 0)
      // The reference spreads its np2 = n0 (nxe - nxs - 1) particles over [nxs, nxe] (app.f90:432) although its own cumcnt puts
      // n0 in each of the cells nxs+1 .. nxe-1 (app.f90:341-344), i.e. left of the wall at nxs+1 nothing: for its first step the
      // particles of the cell nxs are pushed with the weights of the cell nxs+1.  The generator follows cumcnt (a sorted
      // state must be consistent with the positions): evenly spaced over [nxs+1, nxe).
      x = (g.nxgs + 1 + (double)(nxe - g.nxgs - 1) * (ii - 0.5) / nrow) * g.delx;
    else if (kind == 1)
      x = (nxe - 1) * g.delx + (ii - 0.5) / g.n0 * g.delx;                              // app.f90:634
    else
      x = nxe * g.delx + (ii - 0.5) / nrow * (fabs(g.v0) * g.delt);                     // app.f90:776
    const double y = (gj + uni(seed, 0, pid, 1)) * g.delx;
#pragma unroll
    for (int isp = 0; isp < 2; isp++) {
      const double sd = isp ? g.vte : g.vti;
      double ux, uy, uz, xs = x;
      if (kind == 2) {
        // injection (non-relativistic approximation): the thermal velocity moves the particle first   app.f90:799-801
        const double ut = sd * normal(seed, isp + 1, pid, 2);
        xs = x + (g.v0 + ut) * g.delt;
      }
      boosted_maxwellian(g, seed, isp + 1, pid, sd, xs, ux, uy, uz);
      put((isp ? stage1 : stage0) + (size_t)s * 6, xs, y, ux, uy, uz, -(long long)pid);
    }
  }
}

// shock fields: kind 0 = the whole array (uniform B, motional E with the velocity profile, app.f90:414-425); kind 1 = the two
// columns nxe - 1, nxe that inject / relocate refresh every step (app.f90:671-679, 841-849)
__global__ void k_field_shock(const DevParams P, const GenParams g, int kind, int nxe, double *uf) {
  const double bx = g.b0 * cos(g.theta), by = g.b0 * sin(g.theta) * cos(g.phi), bz = g.b0 * sin(g.theta) * sin(g.phi);
  if (kind == 0) {
    const int n = P.pitch * (P.nyl + 4);
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < n; t += gridDim.x * blockDim.x) {
      const int i = g.nxgs - 2 + t % P.pitch;
      double *u = uf + (size_t)t * 6;
      const double v = vprofile(g, i * g.delx);
      u[0] = bx;
      u[1] = by;
      u[2] = bz;
      u[3] = 0.0;
      u[4] = +v * bz / g.c;
      u[5] = -v * by / g.c;
    }
  } else {
    for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < P.nyl + 4; t += gridDim.x * blockDim.x) {
      double *a = uf + ((size_t)t * P.pitch + (nxe - 1 - (g.nxgs - 2))) * 6, *b = a + 6;
      a[1] = by;
      a[2] = bz;
      a[4] = +g.v0 * bz / g.c;
      a[5] = -g.v0 * by / g.c;
      b[1] = by;
      b[2] = bz;
    }
  }
}

void launch_gen_harris(const DevParams &P, const GenParams &g, uint64_t seed, double *stage0, double *stage1, double *uf,
                       cudaStream_t st) {
  k_field_harris<<<148 * 4, 256, 0, st>>>(P, g, uf);
  k_gen_harris<<<148 * 16, 256, 0, st>>>(P, g, seed, stage0, stage1);
}
void launch_gen_shock(const DevParams &P, const GenParams &g, uint64_t seed, int kind, int nxe, const int *rowoff, double *stage0,
                      double *stage1, cudaStream_t st) {
  k_gen_shock<<<148 * 8, 256, 0, st>>>(P, g, seed, kind, nxe, rowoff, stage0, stage1);
}
void launch_field_shock(const DevParams &P, const GenParams &g, int kind, int nxe, double *uf, cudaStream_t st) {
  k_field_shock<<<kind == 0 ? 148 * 4 : 8, 256, 0, st>>>(P, g, kind, nxe, uf);
}

}  // namespace wm
