// fused3_kernel.cu -- the production particle pass of wm_step on sm_100a, warp-specialised.
//
//  k_fused_ws : particle__solv (common/particle.f90:83-169) + ele_cur (common/field.f90:189-316)
//               + bc__particle_x/_y (common/boundary_periodic.f90:61-248, reflecting walls of
//               proj/reconnection/boundary_reconnection.f90:61-99) + the in-place cell sort of the
//               stayers (sort.f90:57-75), one pass over the particles.
//
// Same arithmetic and data flow as k_fused<INPLACE> (fused_kernel.cu), different execution shape.
// k_fused keeps the cell's 65 Esirkepov sums in every lane (130 registers), which caps it at 255
// registers = 8 warps/SM with the serial Boris chain and the ILP-rich deposit in one instruction
// stream: FP64 pipe 44 % busy, latency-bound (profiles/r01b).  Here one CTA of 384 threads per SM is
// split into three warpgroups with different register budgets (setmaxnreg):
//   warpgroups 0,1 (8 warps, 104 registers): PUSH.  Warp (team t, species s) walks the quads
//       t, t+4, ... of the tile for its species: load (register prefetch), gather from the
//       TMA-staged field tile, Boris, move, wall reflection / periodic wrap, in-place sort
//       bookkeeping -- and hands the deposit a 48-byte record per particle (old and new offsets
//       in the cell, q*vz, the four "left/right cell" flags) through a shared-memory ring;
//   warpgroup 2 (4 warps, 232 registers): DEPOSIT.  Warp t consumes the records of both species
//       of its team's quad (mbarrier full/empty pairs, 4 slots per ring), rebuilds the shape
//       factors and does the 80 DFMA per particle into register-resident accumulators, then the
//       shuffle reduce-scatter and the shared-memory tile add at the end of the quad.
// One team per SM sub-partition (warps t, t+4, t+8 share a scheduler): two latency-bound push
// streams and one throughput-bound deposit stream interleave on the FP64 pipe.
#include <cstdint>

#include "kernels.h"

namespace wm {

namespace {

constexpr int FT = 512;          // threads per CTA: 12 push warps + 4 deposit warps
constexpr int WS_TEAMS = 4;      // a team = one deposit warp + WS_PPT push warps, one team per SM sub-partition
constexpr int WS_PPT = 3;        // push warps per team
constexpr int WS_NS = 4;         // ring slots per push warp
constexpr int WS_REC = 6;        // doubles per record: hx hy | d2x d2y | q*vz flags
constexpr int R_PUSH = 104, R_DEP = 200;  // 12*104 + 4*200 = 2048 = 64 K registers / 32 lanes
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


__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

constexpr size_t SM_F = 0;
constexpr size_t SM_J = SM_F + sizeof(double) * WINY * WINX * 6;
constexpr size_t SM_RING = SM_J + sizeof(double) * 3 * JY * JX;
constexpr size_t SM_ARR = SM_RING + sizeof(double) * WS_PPT * WS_TEAMS * WS_NS * 32 * WS_REC;
constexpr size_t SM_NMV = SM_ARR + sizeof(int) * WM_NSP_MAX * WIN;
constexpr size_t SM_BAR = (SM_NMV + sizeof(int) * WM_NSP_MAX * NQ + 7) / 8 * 8;
constexpr size_t SM_TOTAL = SM_BAR + 8 * (1 + 2 * WS_PPT * WS_TEAMS * WS_NS);

}  // namespace

__global__ void __launch_bounds__(FT, 1) k_fused_ws(const DevParams P, const Pass1Args a) {
  constexpr bool INPLACE = true;
  extern __shared__ __align__(128) unsigned char smem[];
  double *const s_f = reinterpret_cast<double *>(smem + SM_F);
  double *const s_j = reinterpret_cast<double *>(smem + SM_J);
  double *const s_ring = reinterpret_cast<double *>(smem + SM_RING);
  int *const s_arr = reinterpret_cast<int *>(smem + SM_ARR);
  int *const s_nmv = reinterpret_cast<int *>(smem + SM_NMV);
  uint64_t *const s_bar = reinterpret_cast<uint64_t *>(smem + SM_BAR);        // TMA
  uint64_t *const s_full = s_bar + 1;                                          // [ring][slot]
  uint64_t *const s_empty = s_full + WS_PPT * WS_TEAMS * WS_NS;

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int li0 = (tile % P.ntx) * TX, lj0 = (tile / P.ntx) * TY;
  const int tw = min(TX, P.nx - li0), th = min(TY, P.nyl - lj0);

  if (tid == 0) {
    mbar_init(s_bar, 1);
    for (int i = 0; i < WS_PPT * WS_TEAMS * WS_NS; i++) {
      mbar_init(&s_full[i], 32);
      mbar_init(&s_empty[i], 32);
    }
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
  for (int e = tid; e < WM_NSP_MAX * WIN; e += FT) s_arr[e] = 0;
  for (int e = tid; e < WM_NSP_MAX * NQ; e += FT) s_nmv[e] = 0;
  __syncthreads();

  const int wid = tid >> 5, lane = tid & 31;
  const int grp = lane >> 3, l8 = lane & 7;

  if (wid < WS_PPT * WS_TEAMS) {
    // =====================================================================  PUSH warps
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(R_PUSH));
    mbar_wait(s_bar, 0);
    // work items of a team in order: (quad 0, species 0), (quad 0, species 1), (quad 1, species 0), ...;
    // push warp j of the team takes items j, j + WS_PPT, ...: every item is one (quad, species), i.e. cell
    // segments nobody else touches, so the in-place compaction needs no coordination between warps
    const int team = wid & (WS_TEAMS - 1), pj = wid / WS_TEAMS;
    double *const ring = s_ring + (size_t)wid * WS_NS * 32 * WS_REC;
    uint64_t *const full = s_full + wid * WS_NS, *const empty = s_empty + wid * WS_NS;
    int bcount = 0;
    const unsigned below = (1u << l8) - 1u;
    const size_t cstride = (size_t)P.cap * P.nsp;  // elements between component arrays (carved SoA)
    double *const px = a.src.x;
    const double delt = P.delt, inv_cc = P.inv_cc, cc = P.cc;
    const double xlo = (double)P.nxgs, xhi = (double)(P.nxgs + P.nx);
    const double ylo = (double)P.nygs, yhi = (double)(P.nygs + P.ny);
    {
      for (int item = pj; item < (NQ / WS_TEAMS) * P.nsp; item += WS_PPT) {
        const int isp = item % P.nsp, q = team + (item / P.nsp) * WS_TEAMS;
        const size_t so = (size_t)isp * P.cap;
        const double qs = P.q[isp];
        // particle.f90:90-92
        const double fac1 = qs / P.r[isp] * 0.5 * delt;
        const double txxx = fac1 * fac1;
        const double fac2 = qs * delt / P.r[isp];
        const int cy = q / QX, cx = (q - cy * QX) * 4 + grp;
        const bool valid = (cx < tw) && (cy < th);
        const int cell = (lj0 + cy) * P.nx + (li0 + cx);
        const int gj = P.nys + lj0 + cy;
        const double di = (double)(P.nxgs + li0 + cx), dj = (double)gj;
        const double cxh = di + 0.5, cyh = dj + 0.5, di1 = di + 1.0, dj1 = dj + 1.0;
        const double *sf0 = &s_f[(cy * WINX + cx) * 6];
        int beg = 0, end = 0;
        if (valid) {
          beg = a.cstart[(size_t)isp * (P.ncell + 1) + cell];
          end = beg + a.cnt[(size_t)isp * P.ncell + cell];
        }
        int nmax = end - beg;
        nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 8));
        nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 16));
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
        bool stay = false;
        double xn = 0.0, yn = 0.0, un1 = 0.0, un2 = 0.0, un3 = 0.0;
        // ring slot of this batch: wait until the deposit warp has drained it
        const int slot = bcount % WS_NS;
        mbar_wait(&empty[slot], ((bcount / WS_NS) & 1) ^ 1);
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
          // ---- new cell relative to the old one (int() truncation == floor: positions > 0)
          const bool xl = xn < di, xr = xn >= di1, yl = yn < dj, yr = yn >= dj1;
          stay = !(xl | xr | yl | yr);
          // ---- hand the deposit what it needs                                field.f90:224-272
          {
            const double d2x = xn - (xl ? cxh - 1.0 : (xr ? cxh + 1.0 : cxh));
            const double d2y = yn - (yl ? cyh - 1.0 : (yr ? cyh + 1.0 : cyh));
            const double qvz = qs * (un3 * wmove);  // q*gvz, field.f90:270-272,295
            const long long fl = (xl ? 1 : 0) | (xr ? 2 : 0) | (yl ? 4 : 0) | (yr ? 8 : 0);
            double2 *r = reinterpret_cast<double2 *>(ring + ((size_t)slot * 32 + lane) * WS_REC);
            r[0] = make_double2(hx, hy);
            r[1] = make_double2(d2x, d2y);
            r[2] = make_double2(qvz, __longlong_as_double(fl));
          }
        }
        mbar_arrive(&full[slot]);
        bcount++;
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
              double *b = px + so + pc;
              b[2 * cstride] = un1;
              b[3 * cstride] = un2;
              b[4 * cstride] = un3;
            }
            stay = !(xn < di || xn >= di1 || yn < dj || yn >= dj1);
          }
        }
        // ---- sort bookkeeping                                             sort.f90:57-62
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
              double *d = px + so + beg + srank;
              d[0] = xn;
              d[cstride] = yn;
              d[2 * cstride] = un1;
              d[3 * cstride] = un2;
              d[4 * cstride] = un3;
              if (beg + srank != pc) d[5 * cstride] = idc;
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
            const double idv = INPLACE ? idc : (leaves ? px[so + pc + 5 * cstride] : 0.0);  // id, bit pattern
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
              // stage the record (64 B: x y | ux uy | uz id | tag -) in the idle store, in the shadow of
              // this quad; slot order = ballot rank, so the stores of a warp are contiguous
              const int sk = nmv + __popc(balm & ((1u << lane) - 1u));
              if (sk < qcap) {
                double2 *d = reinterpret_cast<double2 *>(a.dst.x) + (size_t)(qrec + sk) * 4;
                d[0] = make_double2(xn, yn);
                d[1] = make_double2(un1, un2);
                d[2] = make_double2(un3, idv);
                d[3] = make_double2(__longlong_as_double((long long)tg), 0.0);
              } else {
                atomicOr(a.err, ERR_OVERFLOW);
              }
            }
          }
          if (!INPLACE) {
            double *b = px + so + pc;
            b[0] = xn;
            b[cstride] = yn;
            a.tag[so + pc] = tg;
          }
          nst += __popc(m8);
        }
        nmv += __popc(balm);
      }
        if (valid && l8 == 0) a.cnt_tail[(size_t)isp * P.ncell + cell] = nst;  // arrivals are added by k_place
        if (lane == 0) s_nmv[isp * NQ + q] = nmv;
      }
    }
  } else {
    // =====================================================================  DEPOSIT warps
    asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(R_DEP));
    const int team = wid - WS_PPT * WS_TEAMS;
    const double qf_base = P.delx / P.delt;
    int cbr0 = 0, cbr1 = 0, cbr2 = 0;  // batches consumed from the rings of the team's push warps
    for (int qi = 0; qi < NQ / WS_TEAMS; qi++) {
      const int q = team + qi * WS_TEAMS;
      const int cy = q / QX, cx = (q - cy * QX) * 4 + grp;
      const bool valid = (cx < tw) && (cy < th);
      const int cell = (lj0 + cy) * P.nx + (li0 + cx);
      int n0 = 0, n1 = 0;
      if (valid) {
        n0 = a.cnt[cell];
        if (P.nsp > 1) n1 = a.cnt[(size_t)P.ncell + cell];
      }
      int m0 = n0, m1 = n1;
      m0 = max(m0, __shfl_xor_sync(0xffffffffu, m0, 8));
      m0 = max(m0, __shfl_xor_sync(0xffffffffu, m0, 16));
      m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, 8));
      m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, 16));
      const int nb0 = (m0 + 7) >> 3, nb1 = (m1 + 7) >> 3;

      double acc[65];
#pragma unroll
      for (int e = 0; e < 65; e++) acc[e] = 0.0;

      for (int k = 0; k < max(nb0, nb1); k++) {
#pragma unroll 1
        for (int isp = 0; isp < 2; isp++) {
          if (k >= (isp ? nb1 : nb0)) continue;
          const int pj = (qi * P.nsp + isp) % WS_PPT;  // the push warp that owns this (quad, species)
          const int r = pj * WS_TEAMS + team;
          const int cb = pj == 0 ? cbr0 : (pj == 1 ? cbr1 : cbr2);
          const int slot = cb % WS_NS;
          mbar_wait(&s_full[r * WS_NS + slot], (cb / WS_NS) & 1);
          const bool active = l8 + 8 * k < (isp ? n1 : n0);
          const double2 *rr = reinterpret_cast<const double2 *>(s_ring + (((size_t)r * WS_NS + slot) * 32 + lane) * WS_REC);
          const double2 r0 = rr[0], r1 = rr[1], r2 = rr[2];
          mbar_arrive(&s_empty[r * WS_NS + slot]);
          if (pj == 0) cbr0++; else if (pj == 1) cbr1++; else cbr2++;
          if (active) {
            const double hx = r0.x, hy = r0.y, d2x = r1.x, d2y = r1.y, qvz = r2.x;
            const long long fl = __double_as_longlong(r2.y);
            const bool xl = fl & 1, xr = fl & 2, yl = fl & 4, yr = fl & 8;
            const double qf = P.q[isp] * qf_base;  // q*delx*d_delt, field.f90:278
            // second order shape function about the sorted cell            field.f90:224-236
          const double hx2 = hx * hx, hy2 = hy * hy;
          const double ex = fma(0.5, hx2, 0.125), ey = fma(0.5, hy2, 0.125);
          const double sxm = fma(-0.5, hx, ex), sx0 = 0.75 - hx2, sxp = fma(0.5, hx, ex);
          const double sym = fma(-0.5, hy, ey), sy0 = 0.75 - hy2, syp = fma(0.5, hy, ey);
          // ---- Esirkepov density decomposition, factorised               field.f90:224-298
          //  Jx block = Cx (x) Ty, Jy block = Tx (x) Cy, Jz block = Tx (x) Uy + Hx (x) Vy
          //  T = S0 + DS/2, H = S0/2 + DS/3, C = running sum of -q*dx/dt*DS
          double dsx0, dsx1, dsx2, dsx3, dsx4, dsy0, dsy1, dsy2, dsy3, dsy4;
          {
            const double d2 = d2x;
            const double d22 = d2 * d2, e2 = fma(0.5, d22, 0.125);
            const double s1 = fma(-0.5, d2, e2), s2_ = 0.75 - d22, s3 = fma(0.5, d2, e2);
            dsx0 = xl ? s1 : 0.0;
            dsx1 = (xl ? s2_ : (xr ? 0.0 : s1)) - sxm;
            dsx2 = (xl ? s3 : (xr ? s1 : s2_)) - sx0;
            dsx3 = (xl ? 0.0 : (xr ? s2_ : s3)) - sxp;
            dsx4 = xr ? s3 : 0.0;
          }
          {
            const double d2 = d2y;
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
  }
}

void launch_fused_ws(const DevParams &P, const Pass1Args &a, cudaStream_t st) {
  cudaFuncSetAttribute(k_fused_ws, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SM_TOTAL);
  k_fused_ws<<<P.ntx * P.nty, FT, SM_TOTAL, st>>>(P, a);
}

}  // namespace wm
