// fused2_kernel.cu -- the production particle pass of wm_step on sm_100a, second generation.
//
//  k_fused2 : particle__solv (common/particle.f90:83-169) + ele_cur (common/field.f90:189-316)
//             + bc__particle_x/_y (common/boundary_periodic.f90:61-248) + the destination-cell
//             histogram of sort__bucket (common/sort.f90:57-62), one pass over the particles, in place.
//
// Same data flow as k_fused (fused_kernel.cu: TMA-staged field tile, register prefetch, ballot-ranked
// stayers, one RED.ADD.F64 flush of the current tile per CTA); what changes is who owns the
// Esirkepov accumulators.  k_fused kept all 65 sums of a cell in every lane (130 registers ->
// 255 registers/thread, 8 warps/SM, FP64 pipe 44 % busy, latency-bound).  Here each iteration has two
// phases inside the warp:
//   A  every lane pushes ONE particle and writes the per-axis factors of its current
//      (Ty, Tx, Hx: 5 each; Cx, Cy: 4 each; Vy: 5; Uy: 3 -- the factorisation
//      Jx = Cx (x) Ty, Jy = Tx (x) Cy, Jz = Tx (x) Uy + Hx (x) Vy of field.f90:274-298) to a
//      272-byte record in shared memory (conflict-free STS.128);
//   B  the 8 lanes of a cell sweep the 8 records of their cell; lane l owns 10 of the cell's sums
//      (two 5-vectors: a column of Jx, a row of Jy or a row-part of Jz) and does
//      acc[0..4] += V1 * s1, acc[5..9] += V2 * s2 with V and s picked by per-lane offsets
//      (broadcast LDS.128).  Same 80 DFMA per particle as before, but 20 accumulator registers
//      instead of 130 -> 16 warps/SM, and no shuffle reduction at the end of a cell.
#include <cstdint>

#include "kernels.h"

namespace wm {

namespace {

constexpr int FT = 128;          // threads per CTA
constexpr int FW = FT / 32;      // warps per CTA
constexpr int QX = TX / 4;       // quads per tile row
constexpr int NQ = QX * TY;      // quads per tile
// staged record of one particle (doubles): Ty[5] pad | Tx[5] pad | Hx[5] pad | Cx[4] Cy[4] Vy[5] Uy[3]
constexpr int REC = 34;
constexpr int O_TY = 0, O_TX = 6, O_HX = 12, O_SC = 18;
constexpr int S_CX = O_SC, S_CY = O_SC + 4, S_VY = O_SC + 8, S_UY1 = O_SC + 13;
constexpr int GSTR = 8 * REC + 2;   // doubles per 8-lane group; +2 skews groups by 16 B (banks)
constexpr int WSTAGE = 4 * GSTR;    // doubles per warp
constexpr int joff(int comp, int b, int a) { return (comp * JY + 2 + b) * JX + 2 + a; }
// per-lane ownership: {V1, s1, V2, s2, flush base 1, flush stride 1, flush base 2, flush stride 2}
struct LaneTab {
  int v[8][8];
};
constexpr LaneTab make_tab() {
  LaneTab t{};
  // lanes 0,1: columns a' = -1..2 of Jx (5 rows each): Ty * Cx[a']
  for (int l = 0; l < 2; l++) {
    const int u0 = 2 * l, u1 = 2 * l + 1;
    t.v[l][0] = O_TY; t.v[l][1] = S_CX + u0; t.v[l][2] = O_TY; t.v[l][3] = S_CX + u1;
    t.v[l][4] = joff(0, -2, u0 - 1); t.v[l][5] = JX; t.v[l][6] = joff(0, -2, u1 - 1); t.v[l][7] = JX;
  }
  // lanes 2,3: rows b' = -1..2 of Jy (5 columns each): Tx * Cy[b']
  for (int l = 2; l < 4; l++) {
    const int u0 = 2 * (l - 2), u1 = u0 + 1;
    t.v[l][0] = O_TX; t.v[l][1] = S_CY + u0; t.v[l][2] = O_TX; t.v[l][3] = S_CY + u1;
    t.v[l][4] = joff(1, u0 - 1, -2); t.v[l][5] = 1; t.v[l][6] = joff(1, u1 - 1, -2); t.v[l][7] = 1;
  }
  // lane 4: rows -2 and +2 of Jz: Hx * Vy[0], Hx * Vy[4]   (Uy is zero there)
  t.v[4][0] = O_HX; t.v[4][1] = S_VY + 0; t.v[4][2] = O_HX; t.v[4][3] = S_VY + 4;
  t.v[4][4] = joff(2, -2, -2); t.v[4][5] = 1; t.v[4][6] = joff(2, 2, -2); t.v[4][7] = 1;
  // lanes 5..7: row b = -1, 0, 1 of Jz as two partial sums: Tx * Uy[b] and Hx * Vy[b]
  for (int l = 5; l < 8; l++) {
    const int b = l - 6;
    t.v[l][0] = O_TX; t.v[l][1] = S_UY1 + (b + 1); t.v[l][2] = O_HX; t.v[l][3] = S_VY + (b + 2);
    t.v[l][4] = joff(2, b, -2); t.v[l][5] = 1; t.v[l][6] = joff(2, b, -2); t.v[l][7] = 1;
  }
  return t;
}
__constant__ LaneTab c_tab = make_tab();

constexpr size_t SM_F = 0;                                          // cell-centred fields of the tile (+1 halo)
constexpr size_t SM_J = SM_F + sizeof(double) * WINY * WINX * 6;    // current tile (+2 halo)
constexpr size_t SM_STAGE = SM_J + sizeof(double) * 3 * JY * JX;    // per-warp staging records
constexpr size_t SM_STAY = SM_STAGE + sizeof(double) * FW * WSTAGE;
constexpr size_t SM_ARR = SM_STAY + sizeof(int) * WM_NSP_MAX * TX * TY;
constexpr size_t SM_BAR = (SM_ARR + sizeof(int) * WM_NSP_MAX * WIN + 7) / 8 * 8;
constexpr size_t SM_TOTAL = SM_BAR + 8;

__device__ __forceinline__ double rsqrt_fast(double a) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));  // MUFU.RSQ64H, ~2^-22
  const double e = fma(a, -(y * y), 1.0);                  // 1 - a y^2
  const double p = fma(e, 0.375, 0.5);
  return fma(p, y * e, y);                                 // third order: full double
}

__device__ __forceinline__ double rcp_fast(double a) {
  double y;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(a));    // MUFU.RCP64H
  double e = fma(-a, y, 1.0);
  e = fma(e, e, e);
  y = fma(y, e, y);
  e = fma(-a, y, 1.0);
  return fma(y, e, y);
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "WAIT_%=:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
      "@p bra DONE_%=;\n"
      "bra WAIT_%=;\n"
      "DONE_%=:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
// TMA 1-D bulk copy global -> shared, completion on an mbarrier
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

}  // namespace


__global__ void __launch_bounds__(FT, 4) k_fused2(const DevParams P, const Pass1Args a) {
  extern __shared__ __align__(128) unsigned char smem[];
  double *const s_f = reinterpret_cast<double *>(smem + SM_F);
  double *const s_j = reinterpret_cast<double *>(smem + SM_J);
  double *const s_stage = reinterpret_cast<double *>(smem + SM_STAGE);
  int *const s_stay = reinterpret_cast<int *>(smem + SM_STAY);
  int *const s_arr = reinterpret_cast<int *>(smem + SM_ARR);
  uint64_t *const s_bar = reinterpret_cast<uint64_t *>(smem + SM_BAR);

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int li0 = (tile % P.ntx) * TX, lj0 = (tile / P.ntx) * TY;
  const int tw = min(TX, P.nx - li0), th = min(TY, P.nyl - lj0);

  // ---- stage the cell-centred fields of the tile (+1 halo) with TMA, zero everything else
  if (tid == 0) {
    mbar_init(s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t rowbytes = (uint32_t)(tw + 2) * 48u;
    mbar_expect_tx(s_bar, rowbytes * (uint32_t)(th + 2));
    for (int ly = 0; ly < th + 2; ly++)
      tma_load_1d(&s_f[ly * (WINX * 6)], a.tmpf + ((size_t)(lj0 + 1 + ly) * P.pitch + (li0 + 1)) * 6, rowbytes, s_bar);
  }
  for (int e = tid; e < 3 * JY * JX; e += FT) s_j[e] = 0.0;
  for (int e = tid; e < FW * WSTAGE; e += FT) s_stage[e] = 0.0;  // stale records must stay finite
  for (int e = tid; e < WM_NSP_MAX * TX * TY; e += FT) s_stay[e] = 0;
  for (int e = tid; e < WM_NSP_MAX * WIN; e += FT) s_arr[e] = 0;
  __syncthreads();
  mbar_wait(s_bar, 0);

  const int wid = tid >> 5, lane = tid & 31;
  const int grp = lane >> 3, l8 = lane & 7;
  const unsigned below = (1u << l8) - 1u;
  const size_t cstride = (size_t)P.cap * P.nsp;  // elements between component arrays (carved SoA)
  double *const px = a.src.x;
  const double qf_base = P.delx / P.delt;
  const double delt = P.delt, inv_cc = P.inv_cc, cc = P.cc;
  const double xlo = (double)P.nxgs, xhi = (double)(P.nxgs + P.nx);
  const double ylo = (double)P.nygs, yhi = (double)(P.nygs + P.ny);

  // this lane's record in phase A, this lane's group in phase B
  double *const grp_stage = s_stage + wid * WSTAGE + grp * GSTR;
  double *const my_rec = grp_stage + l8 * REC;
  const double *const b_v1 = grp_stage + c_tab.v[l8][0];
  const double *const b_s1 = grp_stage + c_tab.v[l8][1];
  const double *const b_v2 = grp_stage + c_tab.v[l8][2];
  const double *const b_s2 = grp_stage + c_tab.v[l8][3];
  const int fb1 = c_tab.v[l8][4], fs1 = c_tab.v[l8][5], fb2 = c_tab.v[l8][6], fs2 = c_tab.v[l8][7];

  for (int q = wid; q < NQ; q += FW) {
    const int cy = q / QX, cx = (q - cy * QX) * 4 + grp;
    const bool valid = (cx < tw) && (cy < th);
    const int cell = (lj0 + cy) * P.nx + (li0 + cx);
    const int gi = P.nxgs + li0 + cx, gj = P.nys + lj0 + cy;
    const double di = (double)gi, dj = (double)gj;
    const double cxh = di + 0.5, cyh = dj + 0.5, di1 = di + 1.0, dj1 = dj + 1.0;
    const double *sf0 = &s_f[(cy * WINX + cx) * 6];

    double acc[10];
#pragma unroll
    for (int e = 0; e < 10; e++) acc[e] = 0.0;

    for (int isp = 0; isp < P.nsp; isp++) {
      int beg = 0, end = 0;
      if (valid) {
        beg = a.cstart[(size_t)isp * (P.ncell + 1) + cell];
        end = beg + a.cnt[(size_t)isp * P.ncell + cell];
      }
      int nmax = end - beg;
      nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 8));
      nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 16));
      const size_t so = (size_t)isp * P.cap;
      const double qs = P.q[isp];
      // particle.f90:90-92
      const double fac1 = qs / P.r[isp] * 0.5 * delt;
      const double txxx = fac1 * fac1;
      const double fac2 = qs * delt / P.r[isp];
      const double qf = qs * qf_base;  // q*delx*d_delt, field.f90:278

      int nst = 0;  // stayers of this (cell, species) so far
      int p = beg + l8;
      double nx_ = 0.0, ny_ = 0.0, nu1 = 0.0, nu2 = 0.0, nu3 = 0.0;
      if (p < end) {
        const double *b = px + so + p;
        nx_ = b[0];
        ny_ = b[cstride];
        nu1 = b[2 * cstride];
        nu2 = b[3 * cstride];
        nu3 = b[4 * cstride];
      }
      for (int k = 0; k < nmax; k += 8) {
        const int pc = p;
        const bool active = pc < end;
        const double x = nx_, y = ny_, u1 = nu1, u2 = nu2, u3 = nu3;
        p += 8;
        if (p < end) {  // prefetch the next particle of this lane
          const double *b = px + so + p;
          nx_ = b[0];
          ny_ = b[cstride];
          nu1 = b[2 * cstride];
          nu2 = b[3 * cstride];
          nu3 = b[4 * cstride];
        }
        // ================= phase A: push one particle per lane, stage its current factors
        bool stay = false;
        double xn = 0.0, yn = 0.0;
        double2 *const rec2 = reinterpret_cast<double2 *>(my_rec);
        if (active) {
          // ---- second order shape function about the sorted cell       particle.f90:97-105
          const double hx = x - cxh, hy = y - cyh;
          const double hx2 = hx * hx, hy2 = hy * hy;
          const double ex = fma(0.5, hx2, 0.125), ey = fma(0.5, hy2, 0.125);
          const double sxm = fma(-0.5, hx, ex), sx0 = 0.75 - hx2, sxp = fma(0.5, hx, ex);
          const double sym = fma(-0.5, hy, ey), sy0 = 0.75 - hy2, syp = fma(0.5, hy, ey);
          // ---- 3x3 gather of the six cell-centred components            particle.f90:107-129
          double f0 = 0.0, f1 = 0.0, f2 = 0.0, f3 = 0.0, f4 = 0.0, f5 = 0.0;
#pragma unroll
          for (int b = 0; b < 3; b++) {
            const double syb = (b == 0) ? sym : (b == 1) ? sy0 : syp;
            const double2 *row = reinterpret_cast<const double2 *>(sf0 + b * (WINX * 6));
#pragma unroll
            for (int c = 0; c < 3; c++) {
              const double w = syb * ((c == 0) ? sxm : (c == 1) ? sx0 : sxp);
              const double2 v0 = row[c * 3], v1 = row[c * 3 + 1], v2 = row[c * 3 + 2];
              f0 = fma(w, v0.x, f0);
              f1 = fma(w, v0.y, f1);
              f2 = fma(w, v1.x, f2);
              f3 = fma(w, v1.y, f3);
              f4 = fma(w, v2.x, f4);
              f5 = fma(w, v2.y, f5);
            }
          }
          // ---- Buneman-Boris                                             particle.f90:131-153
          double uvm1 = fma(fac1, f3, u1), uvm2 = fma(fac1, f4, u2), uvm3 = fma(fac1, f5, u3);
          const double s2 = fma(uvm3, uvm3, fma(uvm2, uvm2, fma(uvm1, uvm1, cc)));
          const double igam = rsqrt_fast(s2);
          const double gam = s2 * igam;
          const double fac1r = fac1 * igam;
          const double b2 = fma(f2, f2, fma(f1, f1, f0 * f0));
          const double fac2r = fac2 * rcp_fast(fma(txxx * b2, igam, gam));
          const double uvm4 = fma(fac1r, fma(uvm2, f2, -(uvm3 * f1)), uvm1);
          const double uvm5 = fma(fac1r, fma(uvm3, f0, -(uvm1 * f2)), uvm2);
          const double uvm6 = fma(fac1r, fma(uvm1, f1, -(uvm2 * f0)), uvm3);
          uvm1 = fma(fac2r, fma(uvm5, f2, -(uvm6 * f1)), uvm1);
          uvm2 = fma(fac2r, fma(uvm6, f0, -(uvm4 * f2)), uvm2);
          uvm3 = fma(fac2r, fma(uvm4, f1, -(uvm5 * f0)), uvm3);
          const double un1 = fma(fac1, f3, uvm1), un2 = fma(fac1, f4, uvm2), un3 = fma(fac1, f5, uvm3);
          // ---- move                                                      particle.f90:156-161
          const double uu = fma(un3, un3, fma(un2, un2, un1 * un1));
          const double wmove = rsqrt_fast(fma(uu, inv_cc, 1.0));
          const double dtw = delt * wmove;
          xn = fma(un1, dtw, x);
          yn = fma(un2, dtw, y);
          {
            double *b = px + so + pc;
            b[2 * cstride] = un1;
            b[3 * cstride] = un2;
            b[4 * cstride] = un3;
          }
          // ---- new cell relative to the old one (int() truncation == floor: positions > 0)
          const bool xl = xn < di, xr = xn >= di1, yl = yn < dj, yr = yn >= dj1;
          stay = !(xl | xr | yl | yr);
          // ---- Esirkepov density decomposition, factorised               field.f90:224-298
          //  T = S0 + DS/2, H = S0/2 + DS/3, C = running sum of -q*dx/dt*DS, U = q vz S0, V = q vz DS
          double dsx0, dsx1, dsx2, dsx3, dsx4, dsy0, dsy1, dsy2, dsy3, dsy4;
          {
            const double d2 = xn - (xl ? cxh - 1.0 : (xr ? cxh + 1.0 : cxh));
            const double d22 = d2 * d2, e2 = fma(0.5, d22, 0.125);
            const double s1 = fma(-0.5, d2, e2), s2_ = 0.75 - d22, s3 = fma(0.5, d2, e2);
            dsx0 = xl ? s1 : 0.0;
            dsx1 = (xl ? s2_ : (xr ? 0.0 : s1)) - sxm;
            dsx2 = (xl ? s3 : (xr ? s1 : s2_)) - sx0;
            dsx3 = (xl ? 0.0 : (xr ? s2_ : s3)) - sxp;
            dsx4 = xr ? s3 : 0.0;
          }
          {
            const double d2 = yn - (yl ? cyh - 1.0 : (yr ? cyh + 1.0 : cyh));
            const double d22 = d2 * d2, e2 = fma(0.5, d22, 0.125);
            const double s1 = fma(-0.5, d2, e2), s2_ = 0.75 - d22, s3 = fma(0.5, d2, e2);
            dsy0 = yl ? s1 : 0.0;
            dsy1 = (yl ? s2_ : (yr ? 0.0 : s1)) - sym;
            dsy2 = (yl ? s3 : (yr ? s1 : s2_)) - sy0;
            dsy3 = (yl ? 0.0 : (yr ? s2_ : s3)) - syp;
            dsy4 = yr ? s3 : 0.0;
          }
          const double third = 1.0 / 3.0;
          const double qvz = qs * (un3 * wmove);  // q*gvz, field.f90:270-272,295
          // Ty | Tx | Hx
          rec2[0] = make_double2(0.5 * dsy0, fma(0.5, dsy1, sym));
          rec2[1] = make_double2(fma(0.5, dsy2, sy0), fma(0.5, dsy3, syp));
          rec2[2] = make_double2(0.5 * dsy4, 0.0);
          rec2[3] = make_double2(0.5 * dsx0, fma(0.5, dsx1, sxm));
          rec2[4] = make_double2(fma(0.5, dsx2, sx0), fma(0.5, dsx3, sxp));
          rec2[5] = make_double2(0.5 * dsx4, 0.0);
          rec2[6] = make_double2(third * dsx0, fma(third, dsx1, 0.5 * sxm));
          rec2[7] = make_double2(fma(third, dsx2, 0.5 * sx0), fma(third, dsx3, 0.5 * sxp));
          rec2[8] = make_double2(third * dsx4, 0.0);
          // Cx | Cy | Vy | Uy
          {
            const double c0 = -qf * dsx0, c1 = fma(-qf, dsx1, c0), c2 = fma(-qf, dsx2, c1), c3 = qf * dsx4;
            rec2[9] = make_double2(c0, c1);
            rec2[10] = make_double2(c2, c3);
          }
          {
            const double c0 = -qf * dsy0, c1 = fma(-qf, dsy1, c0), c2 = fma(-qf, dsy2, c1), c3 = qf * dsy4;
            rec2[11] = make_double2(c0, c1);
            rec2[12] = make_double2(c2, c3);
          }
          rec2[13] = make_double2(qvz * dsy0, qvz * dsy1);
          rec2[14] = make_double2(qvz * dsy2, qvz * dsy3);
          rec2[15] = make_double2(qvz * dsy4, qvz * sym);
          rec2[16] = make_double2(qvz * sy0, qvz * syp);
        } else {
          const double2 z = make_double2(0.0, 0.0);
#pragma unroll
          for (int e = 9; e < 17; e++) rec2[e] = z;  // all scalar factors zero: contributes nothing
        }
        // ---- sort bookkeeping                                             sort.f90:57-62
        const unsigned bal = __ballot_sync(0xffffffffu, stay);
        if (active) {
          const unsigned m8 = (bal >> (grp * 8)) & 0xffu;
          uint32_t tg;
          if (stay) {
            tg = ((uint32_t)((cy + 1) * WINX + (cx + 1)) << TAG_WSHIFT) | (uint32_t)(nst + __popc(m8 & below));
          } else {
            // cell changers: periodic wraps with round-toward -inf adds  boundary_periodic.f90:74,82-88,124,147-154
            int incx = (xn >= di1) - (xn < di), incy = (yn >= dj1) - (yn < dj);
            if (!(xn >= di - 1.0 && xn < di1 + 1.0 && yn >= dj - 1.0 && yn < dj1 + 1.0)) {
              atomicOr(a.err, ERR_MOVED_TOO_FAR);  // also catches NaN
              incx = (xn >= di1) ? 1 : ((xn < di) ? -1 : 0);
              incy = (yn >= dj1) ? 1 : ((yn < dj) ? -1 : 0);
            }
            const int j2 = gj + incy;  // unwrapped destination row
            if (xn < xlo)
              xn = __dadd_rd(xn, P.xlen);
            else if (xn >= xhi)
              xn = __dadd_rd(xn, -P.xlen);
            if (yn < ylo)
              yn = __dadd_rd(yn, P.ylen);
            else if (yn >= yhi)
              yn = __dadd_rd(yn, -P.ylen);
            const bool leaves = (P.nsize > 1) && (j2 < P.nys || j2 >= P.nys + P.nyl);
            if (leaves) {
              // record goes to the neighbour's edge row        boundary_periodic.f90:156-161,174-189
              const int dir = (j2 < P.nys) ? 0 : 1;
              const int pos = atomicAdd(&a.sendcnt[dir * P.nsp + isp], 1);
              if (pos < a.sendcap) {
                double *rec = a.send[dir] + ((size_t)isp * a.sendcap + pos) * 6;
                const double *b = px + so + pc;
                rec[0] = xn;
                rec[1] = yn;
                rec[2] = b[2 * cstride];
                rec[3] = b[3 * cstride];
                rec[4] = b[4 * cstride];
                rec[5] = b[5 * cstride];  // id, bit pattern
              } else {
                atomicOr(a.err, ERR_SENDBUF);
              }
              tg = TAG_DEAD;
            } else {
              const int w = (cy + 1 + incy) * WINX + (cx + 1 + incx);
              const int rk = atomicAdd(&s_arr[isp * WIN + w], 1);
              tg = TAG_ARRIVAL | ((uint32_t)w << TAG_WSHIFT) | (uint32_t)rk;
            }
          }
          double *b = px + so + pc;
          b[0] = xn;
          b[cstride] = yn;
          a.tag[so + pc] = tg;
          nst += __popc(m8);
        }
        __syncwarp();
        // ================= phase B: the 8 lanes of a cell sweep its 8 staged records   field.f90:274-298
        {
          const int jmax = min(8, nmax - k);
#pragma unroll
          for (int j = 0; j < 8; j++) {
            if (j < jmax) {
              const double2 a0 = *reinterpret_cast<const double2 *>(b_v1 + j * REC);
              const double2 a1 = *reinterpret_cast<const double2 *>(b_v1 + j * REC + 2);
              const double a2 = b_v1[j * REC + 4];
              const double s1 = b_s1[j * REC];
              const double2 c0 = *reinterpret_cast<const double2 *>(b_v2 + j * REC);
              const double2 c1 = *reinterpret_cast<const double2 *>(b_v2 + j * REC + 2);
              const double c2 = b_v2[j * REC + 4];
              const double s2 = b_s2[j * REC];
              acc[0] = fma(a0.x, s1, acc[0]);
              acc[1] = fma(a0.y, s1, acc[1]);
              acc[2] = fma(a1.x, s1, acc[2]);
              acc[3] = fma(a1.y, s1, acc[3]);
              acc[4] = fma(a2, s1, acc[4]);
              acc[5] = fma(c0.x, s2, acc[5]);
              acc[6] = fma(c0.y, s2, acc[6]);
              acc[7] = fma(c1.x, s2, acc[7]);
              acc[8] = fma(c1.y, s2, acc[8]);
              acc[9] = fma(c2, s2, acc[9]);
            }
          }
        }
        __syncwarp();
      }
      if (valid && l8 == 0) s_stay[isp * (TX * TY) + cy * TX + cx] = nst;
    }

    // ---- each lane adds the 10 sums it owns to the tile                      field.f90:304-310
    if (valid) {
      double *sj0 = &s_j[cy * JX + cx];
#pragma unroll
      for (int e = 0; e < 5; e++) {
        if (acc[e] != 0.0) atomicAdd(sj0 + fb1 + e * fs1, acc[e]);
        if (acc[5 + e] != 0.0) atomicAdd(sj0 + fb2 + e * fs2, acc[5 + e]);
      }
    }
  }
  __syncthreads();

  // ---- one flush of the tile (+2 halo) into uj: window (jx,jy) = padded (li0+jx, lj0+jy)
  {
    const int jw = tw + 4;
    for (int e = tid; e < 3 * (th + 4) * jw; e += FT) {
      const int comp = e / ((th + 4) * jw);
      const int r = e - comp * (th + 4) * jw;
      const int jy = r / jw, jx = r - jy * jw;
      const double v = s_j[(comp * JY + jy) * JX + jx];
      if (v != 0.0) atomicAdd(&a.uj[((size_t)(lj0 + jy) * P.pitch + (li0 + jx)) * 3 + comp], v);
    }
  }
  // ---- reserve this tile's share of every destination cell                  sort.f90:57-62
  for (int e = tid; e < P.nsp * WIN; e += FT) {
    const int isp = e / WIN, w = e - isp * WIN;
    const int wy = w / WINX, wx = w - wy * WINX;
    int ns = 0;
    if (wx >= 1 && wx <= TX && wy >= 1 && wy <= TY) ns = s_stay[isp * (TX * TY) + (wy - 1) * TX + (wx - 1)];
    const int n = ns + s_arr[isp * WIN + w];
    int base = 0;
    if (n > 0) {
      const int cell = window_cell(P, li0, lj0, w);
      if (cell < 0) {
        atomicOr(a.err, ERR_MOVED_TOO_FAR);
      } else {
        base = atomicAdd(&a.gcnt[(size_t)isp * P.ncell + cell], n);
      }
      if (n > (int)TAG_RANK_MASK) atomicOr(a.err, ERR_TAG_RANK);
    }
    int *tb = a.tilebase + ((size_t)tile * P.nsp + isp) * (2 * WIN);
    tb[w] = base;             // stayers (kind 0)
    tb[WIN + w] = base + ns;  // arrivals (kind 1)
  }
}

void launch_fused2(const DevParams &P, const Pass1Args &a, cudaStream_t st) {
  // per device and cheap: contexts on several GPUs may share one process
  cudaFuncSetAttribute(k_fused2, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
  k_fused2<<<P.ntx * P.nty, FT, SM_TOTAL, st>>>(P, a);
}

}  // namespace wm
