// fused5_kernel.cu -- k_fused_sm: the in-place fused particle pass with the Esirkepov deposit split by
// particle kind ("stayer / mover split").
//
//  Same work, data layout and sort protocol as k_fused<INPLACE> (fused_kernel.cu): particle__solv
//  (common/particle.f90:83-169) + ele_cur (common/field.f90:189-316) + bc__particle_x/_y
//  (common/boundary_periodic.f90:61-248) + the stayer half of sort__bucket (common/sort.f90:57-62).
//
//  What is different: the reference's 5x5 block per particle (field.f90:215-217) is the union over the three
//  possible cell shifts inc = -1, 0, +1 (field.f90:238-266).  A particle that stays in its cell (85 % at the
//  Weibel parameters) touches only the inner 3x3: Jx 2x3, Jy 3x2, Jz 3x3 = 21 sums.  So
//   * the particle loop deposits the stayers only, into 21 register accumulators per lane (instead of 65), with
//     the closed form ds(-1,0,+1) = (A - h, -2A, A + h), A = (d'-d)(d'+d)/2, h = (d'-d)/2, and no selects;
//   * a particle that changes cell ("mover") pushes a 48-byte record (old offsets, new offsets, q*vz, q*dx/dt)
//     on a shared-memory queue of its cell's 8 lanes;
//   * when the cell is finished the 8 lanes drain the queue with the full shifted-stencil deposit into the 65-sum
//     block (the 21 stayer sums are the start values of their entries), reduce-scatter it with shuffles and add
//     it once to the shared-memory current tile, as before.
//  The loop body is ~130 instructions shorter per particle and needs 88 fewer registers: 3 CTAs (12 warps) per
//  SM instead of 2.  A queue that fills up mid-cell (more than QCAP movers per cell) falls back to a slow path
//  (shared-memory atomics from the mover's lane), so any density is handled correctly.
#include <cstdint>
#include <cstdlib>

#include "kernels.h"

namespace wm {

namespace {

constexpr int FT = 128;          // threads per CTA
constexpr int FW = FT / 32;      // warps per CTA
constexpr int QX = TX / 4;       // quads per tile row
constexpr int NQ = QX * TY;      // quads per tile
constexpr int QCAP = 40;         // mover records per cell queue (mean 19 at 2 x 64 ppc, 15 % movers)

struct JoffTable {
  int v[72];
};
constexpr JoffTable make_joff() {
  JoffTable t{};
  for (int e = 0; e < 65; e++) {
    int comp = 0, a2 = 0, b2 = 0;
    if (e < 20) {  // Jx[b][a'] : b = -2..2, a = -1..2
      comp = 0; b2 = e / 4 - 2; a2 = e % 4 - 1;
    } else if (e < 40) {  // Jy[b'][a] : b = -1..2, a = -2..2
      comp = 1; b2 = (e - 20) / 5 - 1; a2 = (e - 20) % 5 - 2;
    } else {  // Jz[b][a]
      comp = 2; b2 = (e - 40) / 5 - 2; a2 = (e - 40) % 5 - 2;
    }
    t.v[e] = (comp * JY + (2 + b2)) * JX + (2 + a2);
  }
  return t;
}
__constant__ JoffTable c_joff5 = make_joff();

// entries of the 65-sum block a stayer touches: index into sa[21] -> index into acc[65]
//   Jx: acc[b*4 + q], b = 1..3, q = 1..2      Jy: acc[20 + b*5 + q], b = 1..2, q = 1..3
//   Jz: acc[40 + b*5 + q], b = 1..3, q = 1..3
__host__ __device__ constexpr int stay_slot(int e) {
  if (e < 20) {
    const int b = e / 4, q = e % 4;
    return (b >= 1 && b <= 3 && q >= 1 && q <= 2) ? (b - 1) * 2 + (q - 1) : -1;
  } else if (e < 40) {
    const int b = (e - 20) / 5, q = (e - 20) % 5;
    return (b >= 1 && b <= 2 && q >= 1 && q <= 3) ? 6 + (b - 1) * 3 + (q - 1) : -1;
  } else {
    const int b = (e - 40) / 5, q = (e - 40) % 5;
    return (b >= 1 && b <= 3 && q >= 1 && q <= 3) ? 12 + (b - 1) * 3 + (q - 1) : -1;
  }
}

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
__device__ __forceinline__ void tma_load_1d(void *dst, const void *src, uint32_t bytes, uint64_t *bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
               "l"(src), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}

// second-order shape function about a cell centre, offset d in [-1/2, 1/2)     particle.f90:97-105
__device__ __forceinline__ void shape3(double d, double &sm, double &s0, double &sp) {
  const double d2 = d * d, e = fma(0.5, d2, 0.125);
  sm = fma(-0.5, d, e);
  s0 = 0.75 - d2;
  sp = fma(0.5, d, e);
}

// DS(0..4) of field.f90:238-268 for a new offset dn in (-3/2, 3/2) relative to the OLD cell centre
__device__ __forceinline__ void ds5(double dn, double sm, double s0, double sp, double &d0, double &d1, double &d2, double &d3,
                                    double &d4) {
  const bool l = dn < -0.5, r = dn >= 0.5;
  double t1, t2, t3;
  shape3(dn - (l ? -1.0 : (r ? 1.0 : 0.0)), t1, t2, t3);
  d0 = l ? t1 : 0.0;
  d1 = (l ? t2 : (r ? 0.0 : t1)) - sm;
  d2 = (l ? t3 : (r ? t1 : t2)) - s0;
  d3 = (l ? 0.0 : (r ? t2 : t3)) - sp;
  d4 = r ? t3 : 0.0;
}

// Slow path: a mover whose cell queue is full adds its block straight to the current tile (shared-memory
// atomics, 65 of them).  Never taken at the benchmark densities; keeps any density correct.
__device__ __noinline__ void deposit_mover_slow(double *sj0, double hx, double hy, double dxn, double dyn, double qvz, double qf) {
  double sx[3], sy[3], dx[5], dy[5];
  shape3(hx, sx[0], sx[1], sx[2]);
  shape3(hy, sy[0], sy[1], sy[2]);
  ds5(dxn, sx[0], sx[1], sx[2], dx[0], dx[1], dx[2], dx[3], dx[4]);
  ds5(dyn, sy[0], sy[1], sy[2], dy[0], dy[1], dy[2], dy[3], dy[4]);
  double s0x[5] = {0.0, sx[0], sx[1], sx[2], 0.0}, s0y[5] = {0.0, sy[0], sy[1], sy[2], 0.0};
  const double third = 1.0 / 3.0;
  for (int b = 0; b < 5; b++) {
    const double ty = fma(0.5, dy[b], s0y[b]);
    double c = 0.0;
    for (int q = 0; q < 4; q++) {
      c = fma(-qf, dx[q], c);
      const double v = c * ty;
      if (v != 0.0) atomicAdd(sj0 + c_joff5.v[b * 4 + q], v);
    }
  }
  {
    double c = 0.0;
    for (int b = 0; b < 4; b++) {
      c = fma(-qf, dy[b], c);
      for (int q = 0; q < 5; q++) {
        const double v = fma(0.5, dx[q], s0x[q]) * c;
        if (v != 0.0) atomicAdd(sj0 + c_joff5.v[20 + b * 5 + q], v);
      }
    }
  }
  for (int b = 0; b < 5; b++)
    for (int q = 0; q < 5; q++) {
      const double tx = fma(0.5, dx[q], s0x[q]), hxq = fma(third, dx[q], 0.5 * s0x[q]);
      const double v = fma(hxq, qvz * dy[b], tx * (qvz * s0y[b]));
      if (v != 0.0) atomicAdd(sj0 + c_joff5.v[40 + b * 5 + q], v);
    }
}

}  // namespace

// MINB resident CTAs per SM: 3 -> at most 168 registers.  DRAIN = false drops the movers' current (wrong physics:
// only for timing the particle loop in isolation).
template <int MINB, bool DRAIN>
__global__ void __launch_bounds__(FT, MINB) k_fused_sm(const DevParams P, const Pass1Args a) {
  __shared__ __align__(128) double s_f[WINY * WINX * 6];
  __shared__ __align__(16) double s_j[3 * JY * JX];
  __shared__ __align__(16) double2 s_q[FW * 4 * QCAP * 3];  // [warp][cell of the quad][slot] x (hx hy | dxn dyn | qvz qf)
  __shared__ int s_arr[WM_NSP_MAX * WIN];
  __shared__ int s_nmv[WM_NSP_MAX * NQ];
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int li0 = (tile % P.ntx) * TX, lj0 = (tile / P.ntx) * TY;
  const int tw = min(TX, P.nx - li0), th = min(TY, P.nyl - lj0);

  // ---- stage the cell-centred fields of the tile (+1 halo) with TMA, zero the accumulators
  if (tid == 0) {
    mbar_init(&s_bar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (tid == 0) {
    const uint32_t rowbytes = (uint32_t)(tw + 2) * 48u;
    mbar_expect_tx(&s_bar, rowbytes * (uint32_t)(th + 2));
    for (int ly = 0; ly < th + 2; ly++)
      tma_load_1d(&s_f[ly * (WINX * 6)], a.tmpf + ((size_t)(lj0 + 1 + ly) * P.pitch + (li0 + 1)) * 6, rowbytes, &s_bar);
  }
  for (int e = tid; e < 3 * JY * JX; e += FT) s_j[e] = 0.0;
  for (int e = tid; e < WM_NSP_MAX * WIN; e += FT) s_arr[e] = 0;
  __syncthreads();
  mbar_wait(&s_bar, 0);

  const int wid = tid >> 5, lane = tid & 31;
  const int grp = lane >> 3, l8 = lane & 7;
  const unsigned below = (1u << l8) - 1u;
  const size_t cstride = (size_t)P.cap * P.nsp;  // elements between component arrays (carved SoA)
  double *const px = a.src.x;
  const double qf_base = P.delx / P.delt;
  const double delt = P.delt, inv_cc = P.inv_cc, cc = P.cc;
  double2 *const myq = &s_q[(wid * 4 + grp) * (QCAP * 3)];

  int nb0 = 0, nc0 = 0, nb1 = 0, nc1 = 0;
  {
    const int cy = wid / QX, cx = (wid - cy * QX) * 4 + grp;
    if (cx < tw && cy < th) {
      const int cell = (lj0 + cy) * P.nx + (li0 + cx);
      nb0 = a.cstart[cell];
      nc0 = a.cnt[cell];
      if (P.nsp > 1) {
        nb1 = a.cstart[(size_t)(P.ncell + 1) + cell];
        nc1 = a.cnt[(size_t)P.ncell + cell];
      }
    }
  }
#pragma unroll 1
  for (int q = wid; q < NQ; q += FW) {
    const int cy = q / QX, cx = (q - cy * QX) * 4 + grp;
    const bool valid = (cx < tw) && (cy < th);
    const int cell = (lj0 + cy) * P.nx + (li0 + cx);
    const int gi = P.nxgs + li0 + cx, gj = P.nys + lj0 + cy;
    const double cxh = (double)gi + 0.5, cyh = (double)gj + 0.5;
    const double *sf0 = &s_f[(cy * WINX + cx) * 6];
    double *const sj0 = &s_j[cy * JX + cx];

    double sa[21];
#pragma unroll
    for (int e = 0; e < 21; e++) sa[e] = 0.0;
    int qn = 0;  // movers queued for this cell

    // segment bounds of both species: loaded one quad ahead (nb*), so that the first particles of the next
    // quad can be prefetched while this one is being worked on
    const int beg0 = nb0, cnt0 = nc0, beg1 = nb1, cnt1 = nc1;
    {
      const int qn_ = q + FW;
      const int cyn = qn_ / QX, cxn = (qn_ - cyn * QX) * 4 + grp;
      nb0 = nc0 = nb1 = nc1 = 0;
      if (qn_ < NQ && cxn < tw && cyn < th) {
        const int celln = (lj0 + cyn) * P.nx + (li0 + cxn);
        nb0 = a.cstart[celln];
        nc0 = a.cnt[celln];
        if (P.nsp > 1) {
          nb1 = a.cstart[(size_t)(P.ncell + 1) + celln];
          nc1 = a.cnt[(size_t)P.ncell + celln];
        }
      }
    }
    for (int isp = 0; isp < P.nsp; isp++) {
      const int beg = isp ? beg1 : beg0, end = beg + (isp ? cnt1 : cnt0);
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
      int nmv = 0;  // cell changers of this (quad, species) so far
      long long qrec = 0;
      int qcap = 0;
      if (cy < th && (q - cy * QX) * 4 < tw) {
        const int c0 = (lj0 + cy) * P.nx + li0 + (q - cy * QX) * 4;
        const int *cs = a.cstart + (size_t)isp * (P.ncell + 1);
        stage_region(so_slots(P, isp) + cs[c0], so_slots(P, isp) + cs[min(c0 + 4, (lj0 + cy + 1) * P.nx)], &qrec, &qcap);
      }
      int p = beg + l8;
      double nx_ = 0.0, ny_ = 0.0, nu1 = 0.0, nu2 = 0.0, nu3 = 0.0, nid = 0.0;
      if (p < end) {
        const double *b = px + so + p;
        nx_ = b[0];
        ny_ = b[cstride];
        nu1 = b[2 * cstride];
        nu2 = b[3 * cstride];
        nu3 = b[4 * cstride];
        nid = b[5 * cstride];
      }
      {
        // first particle of what this lane works on next: the other species of this cell, then species 0 of the
        // next quad's cell (bounds loaded one quad ahead)
        const bool last = isp + 1 == P.nsp;
        const int pb = last ? nb0 : beg1, pn = last ? nc0 : cnt1;
        if (l8 < pn) {
          const double *b = px + (last ? (size_t)0 : P.cap) + pb + l8;
#pragma unroll
          for (int cpt = 0; cpt < 6; cpt++) asm volatile("prefetch.global.L1 [%0];" ::"l"(b + cpt * cstride));
        }
      }
      for (int k = 0; k < nmax; k += 8) {
        const int pc = p;
        const bool active = pc < end;
        const double x = nx_, y = ny_, u1 = nu1, u2 = nu2, u3 = nu3, idc = nid;
        p += 8;
        if (p < end) {  // prefetch the next particle of this lane
          const double *b = px + so + p;
          nx_ = b[0];
          ny_ = b[cstride];
          nu1 = b[2 * cstride];
          nu2 = b[3 * cstride];
          nu3 = b[4 * cstride];
          nid = b[5 * cstride];  // the id moves with the record (bit pattern)
        }
        bool stay = false;   // stays in its cell for the sort (after the particle boundary)
        bool dmove = false;  // changes cell for the deposit (before the particle boundary)
        int incx = 0, incy = 0;
        double xn = 0.0, yn = 0.0, un1 = 0.0, un2 = 0.0, un3 = 0.0;
        double hx = 0.0, hy = 0.0, dxn = 0.0, dyn = 0.0, qvz = 0.0;
        if (active) {
          // ---- second order shape function about the sorted cell       particle.f90:97-105
          hx = x - cxh;
          hy = y - cyh;
          double sxm, sx0, sxp, sym, sy0, syp;
          shape3(hx, sxm, sx0, sxp);
          shape3(hy, sym, sy0, syp);
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
          un1 = fma(fac1, f3, uvm1);
          un2 = fma(fac1, f4, uvm2);
          un3 = fma(fac1, f5, uvm3);
          // ---- move                                                      particle.f90:156-161
          const double uu = fma(un3, un3, fma(un2, un2, un1 * un1));
          const double wmove = rsqrt_fast(fma(uu, inv_cc, 1.0));
          const double dtw = delt * wmove;
          xn = fma(un1, dtw, x);
          yn = fma(un2, dtw, y);
          // ---- new cell relative to the old one: xn - cxh is exact, so these are the comparisons
          //      int(gp*d_delx) of field.f90:238 makes (positions > 0: truncation == floor)
          dxn = xn - cxh;
          dyn = yn - cyh;
          const bool xl = dxn < -0.5, xr = dxn >= 0.5, yl = dyn < -0.5, yr = dyn >= 0.5;
          stay = !(xl | xr | yl | yr);
          dmove = !stay;
          incx = (int)xr - (int)xl;
          incy = (int)yr - (int)yl;
          qvz = qs * (un3 * wmove);  // q*gvz, field.f90:270-272,295
          if (stay) {
            // ---- Esirkepov density decomposition of a stayer (inc = 0)     field.f90:224-298
            //  DS(-1,0,+1) = S1 - S0 = (A - h, -2A, A + h),  A = (d'-d)(d'+d)/2,  h = (d'-d)/2
            const double hdx = 0.5 * (dxn - hx), ax = hdx * (dxn + hx);
            const double hdy = 0.5 * (dyn - hy), ay = hdy * (dyn + hy);
            const double dsx1 = ax - hdx, dsx2 = -2.0 * ax, dsx3 = ax + hdx;
            const double dsy1 = ay - hdy, dsy2 = -2.0 * ay, dsy3 = ay + hdy;
            const double tx1 = fma(0.5, dsx1, sxm), tx2 = fma(0.5, dsx2, sx0), tx3 = fma(0.5, dsx3, sxp);
            const double ty1 = fma(0.5, dsy1, sym), ty2 = fma(0.5, dsy2, sy0), ty3 = fma(0.5, dsy3, syp);
            {  // Jx: running sum of -qf*DSx = (-qf dsx1, +qf dsx3)
              const double c1 = -qf * dsx1, c2 = qf * dsx3;
              sa[0] = fma(c1, ty1, sa[0]); sa[1] = fma(c2, ty1, sa[1]);
              sa[2] = fma(c1, ty2, sa[2]); sa[3] = fma(c2, ty2, sa[3]);
              sa[4] = fma(c1, ty3, sa[4]); sa[5] = fma(c2, ty3, sa[5]);
            }
            {  // Jy
              const double c1 = -qf * dsy1, c2 = qf * dsy3;
              sa[6] = fma(tx1, c1, sa[6]); sa[7] = fma(tx2, c1, sa[7]);   sa[8] = fma(tx3, c1, sa[8]);
              sa[9] = fma(tx1, c2, sa[9]); sa[10] = fma(tx2, c2, sa[10]); sa[11] = fma(tx3, c2, sa[11]);
            }
            {  // Jz = q vz (S0x S0y + DSx S0y/2 + S0x DSy/2 + DSx DSy/3) = tx*uy + hx*vy
              const double third = 1.0 / 3.0;
              const double hx1 = fma(third, dsx1, 0.5 * sxm), hx2 = fma(third, dsx2, 0.5 * sx0), hx3 = fma(third, dsx3, 0.5 * sxp);
              const double uy1 = qvz * sym, uy2 = qvz * sy0, uy3 = qvz * syp;
              const double vy1 = qvz * dsy1, vy2 = qvz * dsy2, vy3 = qvz * dsy3;
              sa[12] = fma(hx1, vy1, fma(tx1, uy1, sa[12])); sa[13] = fma(hx2, vy1, fma(tx2, uy1, sa[13])); sa[14] = fma(hx3, vy1, fma(tx3, uy1, sa[14]));
              sa[15] = fma(hx1, vy2, fma(tx1, uy2, sa[15])); sa[16] = fma(hx2, vy2, fma(tx2, uy2, sa[16])); sa[17] = fma(hx3, vy2, fma(tx3, uy2, sa[17]));
              sa[18] = fma(hx1, vy3, fma(tx1, uy3, sa[18])); sa[19] = fma(hx2, vy3, fma(tx2, uy3, sa[19])); sa[20] = fma(hx3, vy3, fma(tx3, uy3, sa[20]));
            }
          }
        }
        // ---- movers: queue the deposit for the drain at the end of the cell
        {
          const unsigned bald = __ballot_sync(0xffffffffu, dmove);
          const unsigned d8 = (bald >> (grp * 8)) & 0xffu;
          if (dmove) {
            const int slot = qn + __popc(d8 & below);
            if (slot < QCAP) {
              double2 *r = myq + slot * 3;
              r[0] = make_double2(hx, hy);
              r[1] = make_double2(dxn, dyn);
              r[2] = make_double2(qvz, qf);
            } else if (DRAIN) {
              deposit_mover_slow(sj0, hx, hy, dxn, dyn, qvz, qf);
            }
          }
          qn = min(qn + __popc(d8), QCAP);
        }
        // ---- reflecting x walls (after the deposit, which uses the position before the boundary)
        //      proj/reconnection/boundary_reconnection.f90:61-99
        if (P.bc != WM_BC_PERIODIC && active) {
          bool flip = false;
          if (xn < P.xwlo) {
            xn = P.xw2lo - xn;
            flip = true;
          } else if (xn >= P.xwhi) {
            xn = P.xw2hi - xn;
            flip = true;
          }
          if (flip) {
            un1 = -un1;
            un2 = -un2;
            un3 = -un3;
            const double d = xn - cxh;
            incx = (int)(d >= 0.5) - (int)(d < -0.5);
            stay = (incx | incy) == 0;
          }
        }
        // ---- sort bookkeeping                                             sort.f90:57-62
        const unsigned bal = __ballot_sync(0xffffffffu, stay);
        const unsigned balm = __ballot_sync(0xffffffffu, active && !stay);  // changers + leavers
        const unsigned m8 = (bal >> (grp * 8)) & 0xffu;
        if (stay) {
          // stable compaction inside the segment: slot beg + rank among the stayers <= pc
          const int ns = beg + nst + __popc(m8 & below);
          double *d = px + so + ns;
          d[0] = xn;
          d[cstride] = yn;
          d[2 * cstride] = un1;
          d[3 * cstride] = un2;
          d[4 * cstride] = un3;
          if (ns != pc) d[5 * cstride] = idc;
        } else if (active) {
          // cell changer.  |move| < 1 cell (CFL), so the new cell is (gi + incx, gj + incy); anything else
          // is an error (also catches NaN)
          if (!(fabs(xn - cxh) < 1.5 && fabs(yn - cyh) < 1.5)) atomicOr(a.err, ERR_MOVED_TOO_FAR);
          // periodic wraps with round-toward -inf adds   boundary_periodic.f90:74,82-88,124,147-154
          const int gi2 = gi + incx, j2 = gj + incy;  // unwrapped destination cell
          if (gi2 < P.nxgs)
            xn = __dadd_rd(xn, P.xlen);
          else if (gi2 >= P.nxgs + P.nx)
            xn = __dadd_rd(xn, -P.xlen);
          if (j2 < P.nygs)
            yn = __dadd_rd(yn, P.ylen);
          else if (j2 >= P.nygs + P.ny)
            yn = __dadd_rd(yn, -P.ylen);
          uint32_t tg;
          if (P.nsize > 1 && (j2 < P.nys || j2 >= P.nys + P.nyl)) {
            // record goes to the neighbour's edge row        boundary_periodic.f90:156-161,174-189
            const int dir = (j2 < P.nys) ? 0 : 1;
            const int pos = atomicAdd(&a.sendcnt[dir * P.nsp + isp], 1);
            if (pos < a.sendcap) {
              double *rec = a.send[dir] + ((size_t)isp * a.sendcap + pos) * 6;
              rec[0] = xn;
              rec[1] = yn;
              rec[2] = un1;
              rec[3] = un2;
              rec[4] = un3;
              rec[5] = idc;
            } else {
              atomicOr(a.err, ERR_SENDBUF);
            }
            tg = TAG_DEAD;
          } else {
            const int w = (cy + 1 + incy) * WINX + (cx + 1 + incx);
            tg = TAG_ARRIVAL | ((uint32_t)w << TAG_WSHIFT) | (uint32_t)atomicAdd(&s_arr[isp * WIN + w], 1);
          }
          // stage the record (64 B: x y | ux uy | uz id | tag -) in the idle store, in the shadow of
          // this quad; slot order = ballot rank, so the stores of a warp are contiguous
          const int sk = nmv + __popc(balm & ((1u << lane) - 1u));
          if (sk < qcap) {
            double2 *d = reinterpret_cast<double2 *>(a.dst.x) + (size_t)(qrec + sk) * 4;
            d[0] = make_double2(xn, yn);
            d[1] = make_double2(un1, un2);
            d[2] = make_double2(un3, idc);
            d[3] = make_double2(__longlong_as_double((long long)tg), 0.0);
          } else {
            atomicOr(a.err, ERR_OVERFLOW);
          }
        }
        nst += __popc(m8);
        nmv += __popc(balm);
      }
      if (valid && l8 == 0) a.cnt_tail[(size_t)isp * P.ncell + cell] = nst;  // arrivals are added by k_place
      if (lane == 0) s_nmv[isp * NQ + q] = nmv;
    }

    // ---- drain the mover queue of the cell into the full block; the stayer sums are the start values
    double acc[65];
#pragma unroll
    for (int e = 0; e < 65; e++) acc[e] = (stay_slot(e) >= 0) ? sa[stay_slot(e) >= 0 ? stay_slot(e) : 0] : 0.0;
    if (DRAIN) {
      __syncwarp();
      int nqm = qn;
      nqm = max(nqm, __shfl_xor_sync(0xffffffffu, nqm, 8));
      nqm = max(nqm, __shfl_xor_sync(0xffffffffu, nqm, 16));
      for (int k = l8; k - l8 < nqm; k += 8) {
        if (k < qn) {
          const double2 *r = myq + k * 3;
          const double2 r0 = r[0], r1 = r[1], r2 = r[2];
          const double qvz = r2.x, qf = r2.y;
          double sxm, sx0, sxp, sym, sy0, syp;
          shape3(r0.x, sxm, sx0, sxp);
          shape3(r0.y, sym, sy0, syp);
          double dsx0, dsx1, dsx2, dsx3, dsx4, dsy0, dsy1, dsy2, dsy3, dsy4;
          ds5(r1.x, sxm, sx0, sxp, dsx0, dsx1, dsx2, dsx3, dsx4);
          ds5(r1.y, sym, sy0, syp, dsy0, dsy1, dsy2, dsy3, dsy4);
          //  Jx block = Cx (x) Ty, Jy block = Tx (x) Cy, Jz block = Tx (x) Uy + Hx (x) Vy
          //  T = S0 + DS/2, H = S0/2 + DS/3, C = running sum of -q*dx/dt*DS
          {
            const double ty0 = 0.5 * dsy0, ty1 = fma(0.5, dsy1, sym), ty2 = fma(0.5, dsy2, sy0), ty3 = fma(0.5, dsy3, syp),
                         ty4 = 0.5 * dsy4;
            const double c0 = -qf * dsx0, c1 = fma(-qf, dsx1, c0), c2 = fma(-qf, dsx2, c1), c3 = qf * dsx4;
            acc[0] = fma(c0, ty0, acc[0]);   acc[1] = fma(c1, ty0, acc[1]);   acc[2] = fma(c2, ty0, acc[2]);   acc[3] = fma(c3, ty0, acc[3]);
            acc[4] = fma(c0, ty1, acc[4]);   acc[5] = fma(c1, ty1, acc[5]);   acc[6] = fma(c2, ty1, acc[6]);   acc[7] = fma(c3, ty1, acc[7]);
            acc[8] = fma(c0, ty2, acc[8]);   acc[9] = fma(c1, ty2, acc[9]);   acc[10] = fma(c2, ty2, acc[10]); acc[11] = fma(c3, ty2, acc[11]);
            acc[12] = fma(c0, ty3, acc[12]); acc[13] = fma(c1, ty3, acc[13]); acc[14] = fma(c2, ty3, acc[14]); acc[15] = fma(c3, ty3, acc[15]);
            acc[16] = fma(c0, ty4, acc[16]); acc[17] = fma(c1, ty4, acc[17]); acc[18] = fma(c2, ty4, acc[18]); acc[19] = fma(c3, ty4, acc[19]);
          }
          {
            const double tx0 = 0.5 * dsx0, tx1 = fma(0.5, dsx1, sxm), tx2 = fma(0.5, dsx2, sx0), tx3 = fma(0.5, dsx3, sxp),
                         tx4 = 0.5 * dsx4;
            {
              const double c0 = -qf * dsy0, c1 = fma(-qf, dsy1, c0), c2 = fma(-qf, dsy2, c1), c3 = qf * dsy4;
              acc[20] = fma(tx0, c0, acc[20]); acc[21] = fma(tx1, c0, acc[21]); acc[22] = fma(tx2, c0, acc[22]); acc[23] = fma(tx3, c0, acc[23]); acc[24] = fma(tx4, c0, acc[24]);
              acc[25] = fma(tx0, c1, acc[25]); acc[26] = fma(tx1, c1, acc[26]); acc[27] = fma(tx2, c1, acc[27]); acc[28] = fma(tx3, c1, acc[28]); acc[29] = fma(tx4, c1, acc[29]);
              acc[30] = fma(tx0, c2, acc[30]); acc[31] = fma(tx1, c2, acc[31]); acc[32] = fma(tx2, c2, acc[32]); acc[33] = fma(tx3, c2, acc[33]); acc[34] = fma(tx4, c2, acc[34]);
              acc[35] = fma(tx0, c3, acc[35]); acc[36] = fma(tx1, c3, acc[36]); acc[37] = fma(tx2, c3, acc[37]); acc[38] = fma(tx3, c3, acc[38]); acc[39] = fma(tx4, c3, acc[39]);
            }
            const double third = 1.0 / 3.0;
            const double hx0 = third * dsx0, hx1 = fma(third, dsx1, 0.5 * sxm), hx2_ = fma(third, dsx2, 0.5 * sx0),
                         hx3 = fma(third, dsx3, 0.5 * sxp), hx4 = third * dsx4;
            const double uy1 = qvz * sym, uy2 = qvz * sy0, uy3 = qvz * syp;
            const double vy0 = qvz * dsy0, vy1 = qvz * dsy1, vy2 = qvz * dsy2, vy3 = qvz * dsy3, vy4 = qvz * dsy4;
            acc[40] = fma(hx0, vy0, acc[40]); acc[41] = fma(hx1, vy0, acc[41]); acc[42] = fma(hx2_, vy0, acc[42]); acc[43] = fma(hx3, vy0, acc[43]); acc[44] = fma(hx4, vy0, acc[44]);
            acc[45] = fma(hx0, vy1, fma(tx0, uy1, acc[45])); acc[46] = fma(hx1, vy1, fma(tx1, uy1, acc[46])); acc[47] = fma(hx2_, vy1, fma(tx2, uy1, acc[47]));
            acc[48] = fma(hx3, vy1, fma(tx3, uy1, acc[48])); acc[49] = fma(hx4, vy1, fma(tx4, uy1, acc[49]));
            acc[50] = fma(hx0, vy2, fma(tx0, uy2, acc[50])); acc[51] = fma(hx1, vy2, fma(tx1, uy2, acc[51])); acc[52] = fma(hx2_, vy2, fma(tx2, uy2, acc[52]));
            acc[53] = fma(hx3, vy2, fma(tx3, uy2, acc[53])); acc[54] = fma(hx4, vy2, fma(tx4, uy2, acc[54]));
            acc[55] = fma(hx0, vy3, fma(tx0, uy3, acc[55])); acc[56] = fma(hx1, vy3, fma(tx1, uy3, acc[56])); acc[57] = fma(hx2_, vy3, fma(tx2, uy3, acc[57]));
            acc[58] = fma(hx3, vy3, fma(tx3, uy3, acc[58])); acc[59] = fma(hx4, vy3, fma(tx4, uy3, acc[59]));
            acc[60] = fma(hx0, vy4, acc[60]); acc[61] = fma(hx1, vy4, acc[61]); acc[62] = fma(hx2_, vy4, acc[62]); acc[63] = fma(hx3, vy4, acc[63]); acc[64] = fma(hx4, vy4, acc[64]);
          }
        }
      }
      __syncwarp();
    }

    // ---- reduce-scatter the 65 partial sums over the 8 lanes of the cell   field.f90:304-310
    {
      const bool h4 = (l8 & 4) != 0, h2 = (l8 & 2) != 0, h1 = (l8 & 1) != 0;
      double t64 = acc[64];
      t64 += __shfl_xor_sync(0xffffffffu, t64, 4);
      t64 += __shfl_xor_sync(0xffffffffu, t64, 2);
      t64 += __shfl_xor_sync(0xffffffffu, t64, 1);
#pragma unroll
      for (int e = 0; e < 32; e++) {
        const double snd = h4 ? acc[e] : acc[e + 32];
        const double kp = h4 ? acc[e + 32] : acc[e];
        acc[e] = kp + __shfl_xor_sync(0xffffffffu, snd, 4);
      }
#pragma unroll
      for (int e = 0; e < 16; e++) {
        const double snd = h2 ? acc[e] : acc[e + 16];
        const double kp = h2 ? acc[e + 16] : acc[e];
        acc[e] = kp + __shfl_xor_sync(0xffffffffu, snd, 2);
      }
#pragma unroll
      for (int e = 0; e < 8; e++) {
        const double snd = h1 ? acc[e] : acc[e + 8];
        const double kp = h1 ? acc[e + 8] : acc[e];
        acc[e] = kp + __shfl_xor_sync(0xffffffffu, snd, 1);
      }
      if (valid) {
        const int ebase = (h4 ? 32 : 0) + (h2 ? 16 : 0) + (h1 ? 8 : 0);
#pragma unroll
        for (int e = 0; e < 8; e++)
          if (acc[e] != 0.0) atomicAdd(sj0 + c_joff5.v[ebase + e], acc[e]);
        if (l8 == 0 && t64 != 0.0) atomicAdd(sj0 + c_joff5.v[64], t64);
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
  // ---- hand the tile's arrival counts per window cell and its staged-record counts to k_place
  int *tb = a.tilebase + (size_t)tile * P.nsp * (2 * WIN);
  for (int e = tid; e < P.nsp * WIN; e += FT) {
    const int isp = e / WIN, w = e - isp * WIN;
    tb[isp * (2 * WIN) + w] = s_arr[e];
    if (w < NQ) tb[isp * (2 * WIN) + WIN + w] = s_nmv[isp * NQ + w];
  }
}

void launch_fused_sm(const DevParams &P, const Pass1Args &a, int variant, cudaStream_t st) {
  if (variant == 9)
    k_fused_sm<3, false><<<P.ntx * P.nty, FT, 0, st>>>(P, a);  // timing experiment: movers' current dropped
  else if (variant == 2)
    k_fused_sm<2, true><<<P.ntx * P.nty, FT, 0, st>>>(P, a);
  else
    k_fused_sm<3, true><<<P.ntx * P.nty, FT, 0, st>>>(P, a);
}

}  // namespace wm
