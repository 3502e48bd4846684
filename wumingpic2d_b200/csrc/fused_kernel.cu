// fused_kernel.cu -- the production particle pass of wm_step on sm_100a.
//
//  k_fused : particle__solv (common/particle.f90:83-169) + ele_cur (common/field.f90:189-316)
//            + bc__particle_x/_y (common/boundary_periodic.f90:61-248) + the destination-cell
//            histogram of sort__bucket (common/sort.f90:57-62), one pass over the particles, in place.
//
// Design ("k_fused"):
//  * one CTA = one TX x TY tile of cells; 4 warps; a warp works on a "quad" of 4 cells adjacent in
//    x, 8 lanes per cell, all 32 lanes in lockstep.  Particles of a cell are contiguous (cell-
//    sorted SoA), so every load/store is a 64-byte run per cell and 256 bytes per warp.
//  * cell-centred E/B of the tile (+1 halo) are staged in shared memory by TMA bulk copies
//    (cp.async.bulk + mbarrier: UBLKCP in SASS), one row per copy, while the tile's current
//    accumulator is being zeroed; a particle gathers its 3x3x6 stencil with 27 LDS.128.
//  * the next particle's five doubles are loaded into registers before the current one is pushed
//    (software prefetch), so HBM latency is covered by ~1000 cycles of FP64 work per iteration.
//  * every lane keeps the cell's Esirkepov block (4x5 + 5x4 + 5x5, the reference's pjx/pjy/pjz,
//    field.f90:215-217) in registers across all its particles of both species; at the end of the
//    cell the 8 lanes reduce-scatter the 65 sums with shuffles (56 adds per lane instead of 195)
//    and each lane adds 8 of them to the shared-memory tile.  The tile (+2 halo) is flushed once
//    per CTA with RED.ADD.F64.  No per-particle atomics on FP64 data.
//  * sort bookkeeping without per-particle atomics for the 85 % of particles that stay in their
//    cell: their rank is (running count) + popc(ballot below me); only cell-changers take a
//    shared-memory integer atomic.  tag = kind<<31 | window cell<<23 | rank makes the scatter
//    pass (k_pass2) a pure copy.
//  * FP64 arithmetic is FMA-contracted; 1/sqrt and 1/x are MUFU seeds + Newton steps without the
//    libm range checks (arguments are >= c^2 > 0).  Agreement with the unfused CPU order is
//    ~1e-15 relative; WM_FLAG_EXACT_PUSH selects the bit-exact k_pass1 instead.
#include <cstdint>
#include <cstdlib>

#include "kernels.h"

namespace wm {

namespace {

constexpr int FT_DEFAULT = 128;  // threads per CTA of the production configuration
constexpr int QX = TX / 4;       // quads per tile row
constexpr int NQ = QX * TY;      // quads per tile

// offset of accumulator entry e inside the current tile relative to (comp 0, row cy, col cx):
// s_j[(comp*JY + cy+2+b)*JX + cx+2+a]
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
__constant__ JoffTable c_joff = make_joff();

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

// INPLACE = false: every particle gets a sort tag and k_pass2 scatters the whole store afterwards.
// INPLACE = true : the sort only moves the particles that change cell.  Stayers are compacted to the
//                  front of their own segment as they are written back (slot = running stayer count +
//                  popc(ballot below me); never ahead of a slot that is still to be read).  A cell
//                  changer is staged, ranked by a warp ballot, in the shadow of its quad in the idle
//                  store (a.dst) with a tag (window cell, rank); k_place appends the staged records to
//                  their new segments, k_mark_dead retires the vacated slots.  No scatter pass, no
//                  per-particle tag, no global atomics in the particle loop.
// FT threads per CTA, MINB resident CTAs per SM: (128, 2) = 8 warps/SM at 255 registers.  CTAs of 96 or
// 64 threads do not buy occupancy: registers are allocated per 4 warps, ptxas then caps at 168 and spills.
template <bool INPLACE, int FT, int MINB>
__global__ void __launch_bounds__(FT, MINB) k_fused(const DevParams P, const Pass1Args a) {
  constexpr int FW = FT / 32;  // warps per CTA
  __shared__ __align__(128) double s_f[WINY * WINX * 6];
  __shared__ __align__(16) double s_j[3 * JY * JX];
  __shared__ int s_stay[WM_NSP_MAX * TX * TY];
  __shared__ int s_arr[WM_NSP_MAX * WIN];
  __shared__ int s_nmv[INPLACE ? WM_NSP_MAX * NQ : 1];
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
  for (int e = tid; e < WM_NSP_MAX * TX * TY; e += FT) s_stay[e] = 0;
  for (int e = tid; e < WM_NSP_MAX * WIN; e += FT) s_arr[e] = 0;
  __syncthreads();
  mbar_wait(&s_bar, 0);

  const int wid = tid >> 5, lane = tid & 31;
  const int grp = lane >> 3, l8 = lane & 7;
  const unsigned below = (1u << l8) - 1u;
  // component offsets of a slot in doubles: x 0, y 1, ux 16, uy 17, uz 32, id 33 (wm_internal.h, PView)
  double *const px = a.src.x.p;
  const double qf_base = P.delx / P.delt;
  const double delt = P.delt, inv_cc = P.inv_cc, cc = P.cc;
  const double xlo = (double)P.nxgs, xhi = (double)(P.nxgs + P.nx);
  const double ylo = (double)P.nygs, yhi = (double)(P.nygs + P.ny);

  for (int q = wid; q < NQ; q += FW) {
    const int cy = q / QX, cx = (q - cy * QX) * 4 + grp;
    const bool valid = (cx < tw) && (cy < th);
    const int cell = (lj0 + cy) * P.nx + (li0 + cx);
    const int gi = P.nxgs + li0 + cx, gj = P.nys + lj0 + cy;
    const double di = (double)gi, dj = (double)gj;
    const double cxh = di + 0.5, cyh = dj + 0.5, di1 = di + 1.0, dj1 = dj + 1.0;
    const double *sf0 = &s_f[(cy * WINX + cx) * 6];

    double acc[65];
#pragma unroll
    for (int e = 0; e < 65; e++) acc[e] = 0.0;

    // segment bounds of both species at once: one memory round trip per quad instead of one per species
    int beg0 = 0, cnt0 = 0, beg1 = 0, cnt1 = 0;
    if (valid) {
      beg0 = a.cstart[cell];
      cnt0 = a.cnt[cell];
      if (P.nsp > 1) {
        beg1 = a.cstart[(size_t)(P.ncell + 1) + cell];
        cnt1 = a.cnt[(size_t)P.ncell + cell];
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
      int nmv = 0;  // cell changers of this (quad, species) so far (INPLACE)
      // staging region of this quad in the idle store: starts at the segment of the quad's first cell
      long long qrec = 0;
      int qcap = 0;
      if (INPLACE && cy < th && (q - cy * QX) * 4 < tw) {
        const int c0 = (lj0 + cy) * P.nx + li0 + (q - cy * QX) * 4;
        const int *cs = a.cstart + (size_t)isp * (P.ncell + 1);
        stage_region(so_slots(P, isp) + cs[c0], so_slots(P, isp) + cs[min(c0 + 4, (lj0 + cy + 1) * P.nx)], &qrec, &qcap);
      }
      int p = beg + l8;
      double nx_ = 0.0, ny_ = 0.0, nu1 = 0.0, nu2 = 0.0, nu3 = 0.0, nid = 0.0;
      if (p < end) {
        const double *b = px + 2 * pslot_w(so + p);
        nx_ = b[0];
        ny_ = b[1];
        nu1 = b[16];
        nu2 = b[17];
        nu3 = b[32];
        if (INPLACE) nid = b[33];
      }
      for (int k = 0; k < nmax; k += 8) {
        const int pc = p;
        const bool active = pc < end;
        const double x = nx_, y = ny_, u1 = nu1, u2 = nu2, u3 = nu3, idc = nid;
        p += 8;
        if (p < end) {  // prefetch the next particle of this lane
          const double *b = px + 2 * pslot_w(so + p);
          nx_ = b[0];
          ny_ = b[1];
          nu1 = b[16];
          nu2 = b[17];
          nu3 = b[32];
          if (INPLACE) nid = b[33];  // the id moves with the record (bit pattern)
        }
        bool stay = false;
        int incx = 0, incy = 0;  // new cell - old cell, from the same comparisons the deposit uses
        double xn = 0.0, yn = 0.0, un1 = 0.0, un2 = 0.0, un3 = 0.0;
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
          un1 = fma(fac1, f3, uvm1);
          un2 = fma(fac1, f4, uvm2);
          un3 = fma(fac1, f5, uvm3);
          // ---- move                                                      particle.f90:156-161
          const double uu = fma(un3, un3, fma(un2, un2, un1 * un1));
          const double wmove = rsqrt_fast(fma(uu, inv_cc, 1.0));
          const double dtw = delt * wmove;
          xn = fma(un1, dtw, x);
          yn = fma(un2, dtw, y);
          if (!INPLACE) {
            double *b = px + 2 * pslot_w(so + pc);
            b[16] = un1;
            b[17] = un2;
            b[32] = un3;
          }
          // ---- new cell relative to the old one (int() truncation == floor: positions > 0)
          const bool xl = xn < di, xr = xn >= di1, yl = yn < dj, yr = yn >= dj1;
          stay = !(xl | xr | yl | yr);
          incx = (int)xr - (int)xl;
          incy = (int)yr - (int)yl;
          // ---- Esirkepov density decomposition, factorised               field.f90:224-298
          //  Jx block = Cx (x) Ty, Jy block = Tx (x) Cy, Jz block = Tx (x) Uy + Hx (x) Vy
          //  T = S0 + DS/2, H = S0/2 + DS/3, C = running sum of -q*dx/dt*DS
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
          // Jx: acc[b*4 + q] += cxv[q]*ty[b]
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
            // Jy: acc[20 + b*5 + q] += tx[q]*cyv[b]
            {
              const double c0 = -qf * dsy0, c1 = fma(-qf, dsy1, c0), c2 = fma(-qf, dsy2, c1), c3 = qf * dsy4;
              acc[20] = fma(tx0, c0, acc[20]); acc[21] = fma(tx1, c0, acc[21]); acc[22] = fma(tx2, c0, acc[22]); acc[23] = fma(tx3, c0, acc[23]); acc[24] = fma(tx4, c0, acc[24]);
              acc[25] = fma(tx0, c1, acc[25]); acc[26] = fma(tx1, c1, acc[26]); acc[27] = fma(tx2, c1, acc[27]); acc[28] = fma(tx3, c1, acc[28]); acc[29] = fma(tx4, c1, acc[29]);
              acc[30] = fma(tx0, c2, acc[30]); acc[31] = fma(tx1, c2, acc[31]); acc[32] = fma(tx2, c2, acc[32]); acc[33] = fma(tx3, c2, acc[33]); acc[34] = fma(tx4, c2, acc[34]);
              acc[35] = fma(tx0, c3, acc[35]); acc[36] = fma(tx1, c3, acc[36]); acc[37] = fma(tx2, c3, acc[37]); acc[38] = fma(tx3, c3, acc[38]); acc[39] = fma(tx4, c3, acc[39]);
            }
            // Jz: acc[40 + b*5 + q] += tx[q]*uy[b] + hx[q]*vy[b]     (uy = 0 for b = 0, 4)
            const double third = 1.0 / 3.0;
            const double qvz = qs * (un3 * wmove);  // q*gvz, field.f90:270-272,295
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
            if (!INPLACE) {
              double *b = px + 2 * pslot_w(so + pc);
              b[16] = un1;
              b[17] = un2;
              b[32] = un3;
            }
            incx = (int)(xn >= di1) - (int)(xn < di);
            stay = (incx | incy) == 0;
          }
        }
        // ---- sort bookkeeping                                             sort.f90:57-62
        if constexpr (INPLACE) {
          const unsigned bal = __ballot_sync(0xffffffffu, stay);
          const unsigned balm = __ballot_sync(0xffffffffu, active && !stay);  // changers + leavers
          const unsigned m8 = (bal >> (grp * 8)) & 0xffu;
          if (stay) {
            // stable compaction inside the segment: slot beg + rank among the stayers <= pc
            const int ns = beg + nst + __popc(m8 & below);
            double *d = px + 2 * pslot_w(so + ns);
            d[0] = xn;
            d[1] = yn;
            d[16] = un1;
            d[17] = un2;
            d[32] = un3;
            if (ns != pc) d[33] = idc;
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
            // stage the record (48 B: x y | ux uy | uz id) and its tag in the idle store, in the shadow of
            // this quad; slot order = ballot rank, so the stores of a warp are contiguous
            const int sk = nmv + __popc(balm & ((1u << lane) - 1u));
            if (sk < qcap) {
              double2 *d = reinterpret_cast<double2 *>(a.dst.x.p) + (size_t)(qrec + sk) * 3;
              d[0] = make_double2(xn, yn);
              d[1] = make_double2(un1, un2);
              d[2] = make_double2(un3, idc);
              a.tag[qrec + sk] = tg;
            } else {
              atomicOr(a.err, ERR_OVERFLOW);
            }
          }
          nst += __popc(m8);
          nmv += __popc(balm);
        } else {
          const unsigned bal = __ballot_sync(0xffffffffu, stay);
          const unsigned balm = INPLACE ? __ballot_sync(0xffffffffu, active && !stay) : 0u;  // changers + leavers
          if (active) {
            const unsigned m8 = (bal >> (grp * 8)) & 0xffu;
            const int srank = nst + __popc(m8 & below);  // my rank among the stayers of this cell
            uint32_t tg;
            if (stay) {
              tg = ((uint32_t)((cy + 1) * WINX + (cx + 1)) << TAG_WSHIFT) | (uint32_t)srank;
              if (INPLACE) {
                // stable compaction inside the segment: slot beg + srank <= pc
                double *d = px + 2 * pslot_w(so + beg + srank);
                d[0] = xn;
                d[1] = yn;
                d[16] = un1;
                d[17] = un2;
                d[32] = un3;
                if (beg + srank != pc) d[33] = idc;
              }
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
              const double idv = INPLACE ? idc : (leaves ? px[2 * pslot_w(so + pc) + 33] : 0.0);  // id, bit pattern
              if (leaves) {
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
                  rec[5] = idv;
                } else {
                  atomicOr(a.err, ERR_SENDBUF);
                }
                tg = TAG_DEAD;
              } else {
                const int w = (cy + 1 + incy) * WINX + (cx + 1 + incx);
                const int rk = atomicAdd(&s_arr[isp * WIN + w], 1);
                tg = TAG_ARRIVAL | ((uint32_t)w << TAG_WSHIFT) | (uint32_t)rk;
              }
              if (INPLACE) {
                // stage the record (48 B: x y | ux uy | uz id) and its tag in the idle store, in the shadow of
                // this quad; slot order = ballot rank, so the stores of a warp are contiguous
                const int sk = nmv + __popc(balm & ((1u << lane) - 1u));
                if (sk < qcap) {
                  double2 *d = reinterpret_cast<double2 *>(a.dst.x.p) + (size_t)(qrec + sk) * 3;
                  d[0] = make_double2(xn, yn);
                  d[1] = make_double2(un1, un2);
                  d[2] = make_double2(un3, idv);
                  a.tag[qrec + sk] = tg;
                } else {
                  atomicOr(a.err, ERR_OVERFLOW);
                }
              }
            }
            if (!INPLACE) {
              double *b = px + 2 * pslot_w(so + pc);
              b[0] = xn;
              b[1] = yn;
              a.tag[so + pc] = tg;
            }
            nst += __popc(m8);
          }
          nmv += __popc(balm);
        }
        }
      if (INPLACE) {
        if (valid && l8 == 0) a.cnt_tail[(size_t)isp * P.ncell + cell] = nst;  // arrivals are added by k_place
        if (lane == 0) s_nmv[isp * NQ + q] = nmv;
      }
      if (!INPLACE && valid && l8 == 0) s_stay[isp * (TX * TY) + cy * TX + cx] = nst;
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
        double *sj0 = &s_j[cy * JX + cx];
#pragma unroll
        for (int e = 0; e < 8; e++)
          if (acc[e] != 0.0) atomicAdd(sj0 + c_joff.v[ebase + e], acc[e]);
        if (l8 == 0 && t64 != 0.0) atomicAdd(sj0 + c_joff.v[64], t64);
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
  if (INPLACE) {
    // ---- hand the tile's arrival counts per window cell and its staged-record counts to k_place
    int *tb = a.tilebase + (size_t)tile * P.nsp * (2 * WIN);
    for (int e = tid; e < P.nsp * WIN; e += FT) {
      const int isp = e / WIN, w = e - isp * WIN;
      tb[isp * (2 * WIN) + w] = s_arr[e];
      if (w < NQ) tb[isp * (2 * WIN) + WIN + w] = s_nmv[isp * NQ + w];
    }
    return;
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

void launch_fused(const DevParams &P, const Pass1Args &a, cudaStream_t st) {
  k_fused<false, FT_DEFAULT, 2><<<P.ntx * P.nty, FT_DEFAULT, 0, st>>>(P, a);
}
void launch_fused_inplace(const DevParams &P, const Pass1Args &a, cudaStream_t st) {
  k_fused<true, FT_DEFAULT, 2><<<P.ntx * P.nty, FT_DEFAULT, 0, st>>>(P, a);
}

}  // namespace wm
