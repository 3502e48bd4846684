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
//   * when the cell is finished the 8 lanes drain the queue with the full shifted-stencil deposit in two passes
//     (Jz: 25 sums; Jx and Jy together: 40 sums), the 21 stayer sums being the start values
//     of their entries; each pass ends with a reduce-scatter over the 8 lanes (shuffles) and one add per entry to
//     the shared-memory current tile.
//  The loop body is ~130 instructions shorter per particle and needs 88 fewer registers: 3 CTAs (12 warps) per
//  SM instead of 2.  A queue that could fill up mid-cell (more than QCAP - 8 movers in a cell) is drained early and
//  the particle loop re-entered, so any density is handled correctly.
//  Particle store: blocks of 8 slots x three 16-byte words (wm_internal.h): a record is loaded with three LDG.128 at
//  the top of its iteration (L1 prefetch hints two iterations ahead), stayers are compacted in place with three
//  STG.128, cell changers are staged as 48-byte records + 4-byte tags for k_place.
//  template <WALL>: 0 periodic x, 1 reflecting walls after the deposit (proj/reconnection), 2 bc__injection before
//  the deposit (proj/shock).
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
  const double e = fma(-a, y, 1.0);  // 1 - a y ~ 2^-22
  const double p = fma(e, e, e);     // third order: y (1 + e + e^2), error e^3
  return fma(y, p, y);
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

// Opaque copies: after pin(v) the compiler can no longer rematerialise v from its defining expression, so the
// value stays in a register across the particle loop instead of being recomputed every iteration.
__device__ __forceinline__ void pin(double &v) { asm volatile("" : "+d"(v)); }
__device__ __forceinline__ void pin(int &v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void pin(unsigned &v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ void pin(size_t &v) { asm volatile("" : "+l"(v)); }
template <typename T>
__device__ __forceinline__ void pin(T *&v) { asm volatile("" : "+l"(v)); }

// weights of a queued mover: old shape function, DS of both directions, q*vz, q*dx/dt
struct MoverW {
  double sxm, sx0, sxp, sym, sy0, syp;
  double dsx0, dsx1, dsx2, dsx3, dsx4, dsy0, dsy1, dsy2, dsy3, dsy4;
  double qvz, qf;
};
__device__ __forceinline__ void mover_weights(const double2 *r, MoverW &w) {
  const double2 r0 = r[0], r1 = r[1], r2 = r[2];
  w.qvz = r2.x;
  w.qf = r2.y;
  shape3(r0.x, w.sxm, w.sx0, w.sxp);
  shape3(r0.y, w.sym, w.sy0, w.syp);
  ds5(r1.x, w.sxm, w.sx0, w.sxp, w.dsx0, w.dsx1, w.dsx2, w.dsx3, w.dsx4);
  ds5(r1.y, w.sym, w.sy0, w.syp, w.dsy0, w.dsy1, w.dsy2, w.dsy3, w.dsy4);
}

// reduce-scatter NP (multiple of 8) partial sums over the 8 lanes of a cell with shuffles, then every lane adds
// its NP/8 totals to the current tile; entry e < nvalid of v is entry ebase + e of the 65-sum block
template <int NP>
__device__ __forceinline__ void rs_add(double (&v)[NP], int l8, bool valid, double *sj0, int ebase, int nvalid) {
  const bool h4 = (l8 & 4) != 0, h2 = (l8 & 2) != 0, h1 = (l8 & 1) != 0;
#pragma unroll
  for (int e = 0; e < NP / 2; e++) {
    const double snd = h4 ? v[e] : v[e + NP / 2];
    const double kp = h4 ? v[e + NP / 2] : v[e];
    v[e] = kp + __shfl_xor_sync(0xffffffffu, snd, 4);
  }
#pragma unroll
  for (int e = 0; e < NP / 4; e++) {
    const double snd = h2 ? v[e] : v[e + NP / 4];
    const double kp = h2 ? v[e + NP / 4] : v[e];
    v[e] = kp + __shfl_xor_sync(0xffffffffu, snd, 2);
  }
#pragma unroll
  for (int e = 0; e < NP / 8; e++) {
    const double snd = h1 ? v[e] : v[e + NP / 8];
    const double kp = h1 ? v[e + NP / 8] : v[e];
    v[e] = kp + __shfl_xor_sync(0xffffffffu, snd, 1);
  }
  if (valid) {
    const int i0 = (h4 ? NP / 2 : 0) + (h2 ? NP / 4 : 0) + (h1 ? NP / 8 : 0);
#pragma unroll
    for (int e = 0; e < NP / 8; e++)
      if (i0 + e < nvalid && v[e] != 0.0) atomicAdd(sj0 + c_joff5.v[ebase + i0 + e], v[e]);
  }
}

}  // namespace

// MINB resident CTAs per SM: 3 -> at most 168 registers.  DRAIN = false drops the movers' current (wrong physics:
// only for timing the particle loop in isolation).
template <int MINB, bool DRAIN, int WALL, int PFD, bool TAIL>
__global__ void __launch_bounds__(FT, MINB) k_fused_sm(const DevParams P, const Pass1Args a) {
  __shared__ __align__(128) double s_f[WINY * WINX * 6];
  __shared__ __align__(16) double s_j[3 * JY * JX];
  __shared__ __align__(16) double2 s_q[FW * 4 * QCAP * 3];  // [warp][cell of the quad][slot] x (hx hy | dxn dyn | qvz qf)
  __shared__ int s_arr[WM_NSP_MAX * WIN];
  __shared__ int s_nmv[WM_NSP_MAX * NQ];
  __shared__ int s_nst[TAIL ? WM_NSP_MAX * TX * TY : 1];  // stayers per (species, cell of the tile)
  __shared__ int s_cs[TAIL ? WM_NSP_MAX * TY * (TX + 1) : 1];  // segment offsets of the tile's cells (+ one column)
  __shared__ __align__(8) uint64_t s_bar;

  const int tid = threadIdx.x;
  const int tile = blockIdx.x + a.tile0;
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
  const int grp = lane >> 3;
  int l8 = lane & 7;
  unsigned lanelt = (1u << lane) - 1u;  // lanes below me
  pin(l8);
  pin(lanelt);
  unsigned below = (1u << l8) - 1u;
  int gsh = grp * 8;
  pin(below);
  pin(gsh);
  double *const px = a.src.x.p;  // blocks of 8 slots x three 16-byte words (x y | ux uy | uz id), wm_internal.h
  const double qf_base = P.delx / P.delt;
  const double delt = P.delt, inv_cc = P.inv_cc, cc = P.cc;
  int myqi = (wid * 4 + grp) * (QCAP * 3);
  pin(myqi);
  double2 *const myq = &s_q[myqi];
  unsigned myq_a = smem_u32(myq), sarr_a = smem_u32(s_arr);  // 32-bit shared addresses for the stores / atomics of the loop
  pin(myq_a);
  pin(sarr_a);

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
    double cxh = (double)gi + 0.5, cyh = (double)gj + 0.5;
    int sfi = (cy * WINX + cx) * 6;
    double *const sj0 = &s_j[cy * JX + cx];
    pin(cxh);
    pin(cyh);
    pin(sfi);
    const double *const sf0 = &s_f[sfi];


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
    // The cell is worked on in rounds: particle loop until both species are done or the mover queue of one of
    // the quad's cells may overflow in the next iteration, then the drain.  One round per cell unless a cell has
    // more than QCAP - 8 movers.  Across a drain only (isp, k0, nst, nmv) survive: the loop re-enters at
    // iteration k0 of species isp and reloads its particle (a slot is never overwritten before it is read).
    int isp = 0, k0 = 0;
    int nst = 0;  // stayers of this (cell, species) so far
    int nmv = 0;  // cell changers of this (quad, species) so far
    do {
    double sa[21];
#pragma unroll
    for (int e = 0; e < 21; e++) sa[e] = 0.0;
    int qn = 0;  // movers queued for this cell
    bool full = false;
    while (isp < P.nsp && !full) {
      const int beg = isp ? beg1 : beg0, end = beg + (isp ? cnt1 : cnt0);
      int nmax = end - beg;
      nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 8));
      nmax = max(nmax, __shfl_xor_sync(0xffffffffu, nmax, 16));
      const size_t so = (size_t)isp * P.cap;
      double qs = P.q[isp];
      // particle.f90:90-92
      double fac1 = qs / P.r[isp] * 0.5 * delt;
      const double txxx = fac1 * fac1;
      const double fac2 = qs * delt / P.r[isp];
      double qf = qs * qf_base;  // q*delx*d_delt, field.f90:278
      pin(qs);
      pin(fac1);
      pin(qf);

      long long qrec = 0;
      int qcap = 0;
      if (cy < th && (q - cy * QX) * 4 < tw) {
        const int c0 = (lj0 + cy) * P.nx + li0 + (q - cy * QX) * 4;
        const int *cs = a.cstart + (size_t)isp * (P.ncell + 1);
        stage_region(so_slots(P, isp) + cs[c0], so_slots(P, isp) + cs[min(c0 + 4, (lj0 + cy + 1) * P.nx)], &qrec, &qcap);
      }
      if (TAIL) {  // segment offsets of the tile's cells for the tail of the CTA (end of the quad = start of the next one)
        int *const scs = &s_cs[(isp * TY + cy) * (TX + 1)];
        if (valid && l8 == 0) scs[cx] = beg;
        if (lane == 0 && qcap > 0) scs[min((q - cy * QX) * 4 + 4, tw)] = (int)(qrec - so_slots(P, isp)) + qcap;
      }
      int p = beg + l8 + k0;
      double2 *pbase = reinterpret_cast<double2 *>(px + 6 * ((size_t)isp * P.cap));  // slot 0 of this species (cap % 8 == 0)
      pin(pbase);
      __builtin_assume(__isGlobal(pbase));
      const double2 *pl = pbase + pslot_w((size_t)p);        // word 0 of the lane's current slot; + 8, + 16: words 1, 2
      const int w0 = isp * WIN + (cy + 1) * WINX + (cx + 1);  // this cell in the window of arrival counters
      // Latency hiding without registers: the lines of the iterations k + 1 and k + 2 are pulled into L1 by prefetch
      // hints (PFD iterations ahead in steady state); the loads at the top of iteration k are L1 hits.  Hints for
      // iteration k0 were issued while the previous species / quad was being worked on.
      if (p + 8 < end) {
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pl + 24));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pl + 24 + 8));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pl + 24 + 16));
      }
      {
        // first particle of what this lane works on next: the other species of this cell, then species 0 of the
        // next quad's cell (bounds loaded one quad ahead)
        const bool last = isp + 1 == P.nsp;
        const int pb = last ? nb0 : beg1, pn = last ? nc0 : cnt1;
        if (l8 < pn) {
          const double2 *b = reinterpret_cast<const double2 *>(px) + pslot_w((last ? (size_t)0 : (size_t)P.cap) + pb + l8);
          asm volatile("prefetch.global.L1 [%0];" ::"l"(b));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(b + 8));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(b + 16));
        }
        if (PFD >= 10) {
          // two-level prefetch: the WHOLE list this lane's group works on next goes to L2 now, a pass (8-9 iterations) ahead,
          // so that the L1 hints and the loads of that pass find their lines in L2 instead of DRAM.  A list of n slots is
          // ceil(n / 8) blocks of three 128-byte lines; the 8 lanes of the group take every 8th line.
          const double2 *b0 = reinterpret_cast<const double2 *>(px) + pslot_w((last ? (size_t)0 : (size_t)P.cap) + pb);
          const int nln = ((pn + 7) >> 3) * 3;
          for (int ln = l8; ln < nln; ln += 8) asm volatile("prefetch.global.L2 [%0];" ::"l"(b0 + ln * 8));
        }
      }
      int k = k0;
      for (; k < nmax; k += 8) {
        const int pc = p;
        const bool active = pc < end;
        double xn, yn, un1, un2, un3;
        double hx, hy, dxn, dyn, qvz;
        double sxm, sx0, sxp, sym, sy0, syp;
        double idc;  // the id moves with the record (bit pattern)
        if (p + 8 * (PFD % 10) < end) {
          asm volatile("prefetch.global.L1 [%0];" ::"l"(pl + 24 * (PFD % 10)));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(pl + 24 * (PFD % 10) + 8));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(pl + 24 * (PFD % 10) + 16));
        }
        if (active) {
          const double2 r0 = pl[0], r1 = pl[8], r2 = pl[16];
          const double x = r0.x, y = r0.y, u1 = r1.x, u2 = r1.y, u3 = r2.x;
          idc = r2.y;
          // ---- second order shape function about the sorted cell       particle.f90:97-105
          hx = x - cxh;
          hy = y - cyh;
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
          double wmove = rsqrt_fast(fma(uu, inv_cc, 1.0));
          const double dtw = delt * wmove;
          xn = fma(un1, dtw, x);
          yn = fma(un2, dtw, y);
          // ---- new cell relative to the old one: xn - cxh is exact, so these are the comparisons
          //      int(gp*d_delx) of field.f90:238 makes (positions > 0: truncation == floor)
          if (WALL == 2) {
            // bc__injection, before the deposit (proj/shock/app.f90:112-113, boundary_shock.f90:280-291): reflecting
            // wall at nxs+1, injection wall at xend (mirror in the frame that moves with u0)
            if (xn < P.xwlo) {
              xn = P.xw2lo - xn;
              un1 = -un1;
              un2 = -un2;
              un3 = -un3;
            } else if (xn > P.xwhi) {
              xn = P.xw2hi - xn;
              un1 = P.u0x2 - un1;
              un2 = -un2;
              un3 = -un3;
              // ele_cur takes vz = uz/gamma from the momentum it finds in gp (field.f90:270-272): the new one
              wmove = rsqrt_fast(fma(fma(un3, un3, fma(un2, un2, un1 * un1)), inv_cc, 1.0));
            }
          }
          dxn = xn - cxh;
          dyn = yn - cyh;
          qvz = qs * (un3 * wmove);  // q*gvz, field.f90:270-272,295
        }
        p += 8;
        pl += 24;
        // stays in its cell as far as the deposit is concerned (before the particle boundary): the comparisons
        // int(gp*d_delx) of field.f90:238 makes (positions > 0: truncation == floor; xn - cxh is exact)
        const bool stay = active && dxn >= -0.5 && dxn < 0.5 && dyn >= -0.5 && dyn < 0.5;
        const bool dmove = active && !stay;
        {
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
            {  // Jz = q vz (S0x S0y + DSx S0y/2 + S0x DSy/2 + DSx DSy/3) = q vz (Tx Ty + DSx DSy/12),  T = S0 + DS/2
              const double q12 = qvz * (1.0 / 12.0);
              const double uy1 = qvz * ty1, uy2 = qvz * ty2, uy3 = qvz * ty3;
              const double vy1 = q12 * dsy1, vy2 = q12 * dsy2, vy3 = q12 * dsy3;
              sa[12] = fma(dsx1, vy1, fma(tx1, uy1, sa[12])); sa[13] = fma(dsx2, vy1, fma(tx2, uy1, sa[13])); sa[14] = fma(dsx3, vy1, fma(tx3, uy1, sa[14]));
              sa[15] = fma(dsx1, vy2, fma(tx1, uy2, sa[15])); sa[16] = fma(dsx2, vy2, fma(tx2, uy2, sa[16])); sa[17] = fma(dsx3, vy2, fma(tx3, uy2, sa[17]));
              sa[18] = fma(dsx1, vy3, fma(tx1, uy3, sa[18])); sa[19] = fma(dsx2, vy3, fma(tx2, uy3, sa[19])); sa[20] = fma(dsx3, vy3, fma(tx3, uy3, sa[20]));
            }
          }
        }
        // ---- movers: queue the deposit for the drain at the end of the cell
        {
          const unsigned bald = __ballot_sync(0xffffffffu, dmove);
          const unsigned d8 = (bald >> gsh) & 0xffu;
          if (dmove) {
            const unsigned r = myq_a + (unsigned)(qn + __popc(d8 & below)) * 48u;  // qn <= QCAP - 8 here
            asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(r), "d"(hx), "d"(hy) : "memory");
            asm volatile("st.shared.v2.f64 [%0+16], {%1, %2};" ::"r"(r), "d"(dxn), "d"(dyn) : "memory");
            asm volatile("st.shared.v2.f64 [%0+32], {%1, %2};" ::"r"(r), "d"(qvz), "d"(qf) : "memory");
          }
          qn += __popc(d8);
        }
        // ---- reflecting x walls (after the deposit, which uses the position before the boundary)
        //      proj/reconnection/boundary_reconnection.f90:61-99
        bool sstay = stay;  // stays in its cell as far as the sort is concerned (after the particle boundary)
        if (WALL == 1) {
          if (active) {
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
              dxn = xn - cxh;
              sstay = dxn >= -0.5 && dxn < 0.5 && dyn >= -0.5 && dyn < 0.5;
            }
          }
        }
        // ---- sort bookkeeping                                             sort.f90:57-62
        const unsigned bal = __ballot_sync(0xffffffffu, sstay);
        const unsigned balm = __ballot_sync(0xffffffffu, active && !sstay);  // changers + leavers
        const unsigned m8 = (bal >> gsh) & 0xffu;
        if (sstay) {
          // stable compaction inside the segment: slot beg + rank among the stayers <= pc
          const int ns = beg + nst + __popc(m8 & below);
          double2 *d = pbase + pslot_w((size_t)ns);
          d[0] = make_double2(xn, yn);
          d[8] = make_double2(un1, un2);
          d[16] = make_double2(un3, idc);  // the id moves with the record (bit pattern)
        } else if (active) {
          // cell changer.  |move| < 1 cell (CFL), so the new cell is (gi + incx, gj + incy); anything else
          // is an error (also catches NaN)
          if (!(fabs(dxn) < 1.5 && fabs(dyn) < 1.5)) atomicOr(a.err, ERR_MOVED_TOO_FAR);
          const int incx = (int)(dxn >= 0.5) - (int)(dxn < -0.5), incy = (int)(dyn >= 0.5) - (int)(dyn < -0.5);
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
            const int w = w0 + incy * WINX + incx;
            unsigned rk;  // rank among the arrivals of window cell w: one shared-memory integer atomic
            asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(rk) : "r"(sarr_a + (unsigned)w * 4u) : "memory");
            tg = TAG_ARRIVAL | ((uint32_t)(w - isp * WIN) << TAG_WSHIFT) | rk;
          }
          // stage the record (48 B: x y | ux uy | uz id) and its tag in the idle store, in the shadow of
          // this quad; slot order = ballot rank, so the stores of a warp are contiguous
          const int sk = nmv + __popc(balm & lanelt);
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
        if (DRAIN && __any_sync(0xffffffffu, qn > QCAP - 8)) {  // the next iteration could overflow a queue: drain first
          k += 8;
          full = true;
          break;
        }
      }
      if (k >= nmax) {
        if (TAIL) {
          if (l8 == 0) s_nst[isp * (TX * TY) + cy * TX + cx] = nst;  // the new count is written by the tail of the CTA
        } else {
          if (valid && l8 == 0) a.cnt_tail[(size_t)isp * P.ncell + cell] = nst;  // arrivals are added by k_place
        }
        if (lane == 0) s_nmv[isp * NQ + q] = nmv;
        isp++;
        k0 = 0;
        nst = 0;
        nmv = 0;
      } else {
        k0 = k;
      }
    }

    // ---- drain the mover queue of the cell in two passes (Jz; Jx and Jy together), so that at most 40 sums are
    //      live at a time; the stayer sums are the start values of their entries.  Each pass ends with
    //      the reduce-scatter over the 8 lanes of the cell and one add per entry to the shared-memory tile
    //      (field.f90:304-310).
    {
      __syncwarp();
      int nqm = DRAIN ? qn : 0;
      nqm = max(nqm, __shfl_xor_sync(0xffffffffu, nqm, 8));
      nqm = max(nqm, __shfl_xor_sync(0xffffffffu, nqm, 16));
      {  // Jz block = Tx (x) Uy + Hx (x) Vy     acc[40 + b*5 + q]
        double v[32];
#pragma unroll
        for (int e = 0; e < 32; e++) v[e] = (e < 25 && stay_slot(40 + e) >= 0) ? sa[stay_slot(40 + (e < 25 ? e : 0)) >= 0 ? stay_slot(40 + (e < 25 ? e : 0)) : 0] : 0.0;
        for (int k = l8; k - l8 < nqm; k += 8) {
          if (k < qn) {
            MoverW w;
            mover_weights(myq + k * 3, w);
            const double third = 1.0 / 3.0;
            const double tx0 = 0.5 * w.dsx0, tx1 = fma(0.5, w.dsx1, w.sxm), tx2 = fma(0.5, w.dsx2, w.sx0), tx3 = fma(0.5, w.dsx3, w.sxp),
                         tx4 = 0.5 * w.dsx4;
            const double hx0 = third * w.dsx0, hx1 = fma(third, w.dsx1, 0.5 * w.sxm), hx2 = fma(third, w.dsx2, 0.5 * w.sx0),
                         hx3 = fma(third, w.dsx3, 0.5 * w.sxp), hx4 = third * w.dsx4;
            const double uy1 = w.qvz * w.sym, uy2 = w.qvz * w.sy0, uy3 = w.qvz * w.syp;
            const double vy0 = w.qvz * w.dsy0, vy1 = w.qvz * w.dsy1, vy2 = w.qvz * w.dsy2, vy3 = w.qvz * w.dsy3, vy4 = w.qvz * w.dsy4;
            v[0] = fma(hx0, vy0, v[0]); v[1] = fma(hx1, vy0, v[1]); v[2] = fma(hx2, vy0, v[2]); v[3] = fma(hx3, vy0, v[3]); v[4] = fma(hx4, vy0, v[4]);
            v[5] = fma(hx0, vy1, fma(tx0, uy1, v[5])); v[6] = fma(hx1, vy1, fma(tx1, uy1, v[6])); v[7] = fma(hx2, vy1, fma(tx2, uy1, v[7]));
            v[8] = fma(hx3, vy1, fma(tx3, uy1, v[8])); v[9] = fma(hx4, vy1, fma(tx4, uy1, v[9]));
            v[10] = fma(hx0, vy2, fma(tx0, uy2, v[10])); v[11] = fma(hx1, vy2, fma(tx1, uy2, v[11])); v[12] = fma(hx2, vy2, fma(tx2, uy2, v[12]));
            v[13] = fma(hx3, vy2, fma(tx3, uy2, v[13])); v[14] = fma(hx4, vy2, fma(tx4, uy2, v[14]));
            v[15] = fma(hx0, vy3, fma(tx0, uy3, v[15])); v[16] = fma(hx1, vy3, fma(tx1, uy3, v[16])); v[17] = fma(hx2, vy3, fma(tx2, uy3, v[17]));
            v[18] = fma(hx3, vy3, fma(tx3, uy3, v[18])); v[19] = fma(hx4, vy3, fma(tx4, uy3, v[19]));
            v[20] = fma(hx0, vy4, v[20]); v[21] = fma(hx1, vy4, v[21]); v[22] = fma(hx2, vy4, v[22]); v[23] = fma(hx3, vy4, v[23]); v[24] = fma(hx4, vy4, v[24]);
          }
        }
        rs_add<32>(v, l8, valid, sj0, 40, 25);
      }
      {  // Jx block = Cx (x) Ty  acc[b*4 + q] and Jy block = Tx (x) Cy  acc[20 + b*5 + q] in one pass over the queue
         // (C = running sum of -q*dx/dt*DS): 40 sums live, the weights of a mover are computed once for both
        double v[24], u[24];
#pragma unroll
        for (int e = 0; e < 24; e++) {
          v[e] = (e < 20 && stay_slot(e < 20 ? e : 0) >= 0) ? sa[stay_slot(e < 20 ? e : 0) >= 0 ? stay_slot(e < 20 ? e : 0) : 0] : 0.0;
          u[e] = (e < 20 && stay_slot(20 + (e < 20 ? e : 0)) >= 0) ? sa[stay_slot(20 + (e < 20 ? e : 0)) >= 0 ? stay_slot(20 + (e < 20 ? e : 0)) : 0] : 0.0;
        }
        for (int k = l8; k - l8 < nqm; k += 8) {
          if (k < qn) {
            MoverW w;
            mover_weights(myq + k * 3, w);
            {
              const double ty0 = 0.5 * w.dsy0, ty1 = fma(0.5, w.dsy1, w.sym), ty2 = fma(0.5, w.dsy2, w.sy0), ty3 = fma(0.5, w.dsy3, w.syp),
                           ty4 = 0.5 * w.dsy4;
              const double c0 = -w.qf * w.dsx0, c1 = fma(-w.qf, w.dsx1, c0), c2 = fma(-w.qf, w.dsx2, c1), c3 = w.qf * w.dsx4;
              v[0] = fma(c0, ty0, v[0]);   v[1] = fma(c1, ty0, v[1]);   v[2] = fma(c2, ty0, v[2]);   v[3] = fma(c3, ty0, v[3]);
              v[4] = fma(c0, ty1, v[4]);   v[5] = fma(c1, ty1, v[5]);   v[6] = fma(c2, ty1, v[6]);   v[7] = fma(c3, ty1, v[7]);
              v[8] = fma(c0, ty2, v[8]);   v[9] = fma(c1, ty2, v[9]);   v[10] = fma(c2, ty2, v[10]); v[11] = fma(c3, ty2, v[11]);
              v[12] = fma(c0, ty3, v[12]); v[13] = fma(c1, ty3, v[13]); v[14] = fma(c2, ty3, v[14]); v[15] = fma(c3, ty3, v[15]);
              v[16] = fma(c0, ty4, v[16]); v[17] = fma(c1, ty4, v[17]); v[18] = fma(c2, ty4, v[18]); v[19] = fma(c3, ty4, v[19]);
            }
            {
              const double tx0 = 0.5 * w.dsx0, tx1 = fma(0.5, w.dsx1, w.sxm), tx2 = fma(0.5, w.dsx2, w.sx0), tx3 = fma(0.5, w.dsx3, w.sxp),
                           tx4 = 0.5 * w.dsx4;
              const double c0 = -w.qf * w.dsy0, c1 = fma(-w.qf, w.dsy1, c0), c2 = fma(-w.qf, w.dsy2, c1), c3 = w.qf * w.dsy4;
              u[0] = fma(tx0, c0, u[0]);   u[1] = fma(tx1, c0, u[1]);   u[2] = fma(tx2, c0, u[2]);   u[3] = fma(tx3, c0, u[3]);   u[4] = fma(tx4, c0, u[4]);
              u[5] = fma(tx0, c1, u[5]);   u[6] = fma(tx1, c1, u[6]);   u[7] = fma(tx2, c1, u[7]);   u[8] = fma(tx3, c1, u[8]);   u[9] = fma(tx4, c1, u[9]);
              u[10] = fma(tx0, c2, u[10]); u[11] = fma(tx1, c2, u[11]); u[12] = fma(tx2, c2, u[12]); u[13] = fma(tx3, c2, u[13]); u[14] = fma(tx4, c2, u[14]);
              u[15] = fma(tx0, c3, u[15]); u[16] = fma(tx1, c3, u[16]); u[17] = fma(tx2, c3, u[17]); u[18] = fma(tx3, c3, u[18]); u[19] = fma(tx4, c3, u[19]);
            }
          }
        }
        rs_add<24>(v, l8, valid, sj0, 0, 20);
        rs_add<24>(u, l8, valid, sj0, 20, 20);
      }
      __syncwarp();
    }
    } while (isp < P.nsp);
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
  // ---- sort, in-tile part (TAIL): a cell changer whose new cell belongs to this tile (88 % of them) is appended to the
  //      tail of its new segment right here (sort.f90:71-75 reduced to the cell changers).  Every cell of the tile is
  //      finished (the barrier after the quad loop), so the slots behind the stayers of a segment are free, and this CTA
  //      alone knows the stayer counts: slot = segment start + stayers + rank, the rank being the one the particle loop
  //      drew from the shared-memory arrival counter.  One pass in staging order, a warp per staging region (quad,
  //      species), four regions x 64 records in flight per warp: tag + three 16-byte words each (ld.global.cg: this CTA
  //      wrote them a moment ago); the 16-byte pieces of a destination line written by different threads merge in L2.
  //      k_place is left with the arrivals of the window's rim (other tiles' cells).  Shared memory: the field tile is
  //      dead by now and holds the tables.
  if (TAIL) {
    constexpr int RB = 4;  // regions in flight per warp
    static_assert(2 * WM_NSP_MAX * WIN * 4 <= (int)sizeof(double) * WINY * WINX * 6, "tail tables must fit the field tile");
    static_assert(TX * TY == FT, "one thread per cell of the tile");
    static_assert((WM_NSP_MAX * NQ) % (FW * RB) == 0, "regions per warp");
    int *const t_base = reinterpret_cast<int *>(s_f);     // [nwin] first slot of the arrivals, -1 = not my cell
    int *const t_end = t_base + WM_NSP_MAX * WIN;         // [nwin] end of the segment
    const int nwin = P.nsp * WIN, nreg = P.nsp * NQ;
    for (int e = tid; e < nwin; e += FT) {  // rim of the window: not mine
      const int w = e % WIN, wy = w / WINX, wx = w - wy * WINX;
      if (!(wx >= 1 && wx <= tw && wy >= 1 && wy <= th)) t_base[e] = -1;
    }
    {  // thread = cell of the tile: its arrivals go to [start + stayers, start + capacity)
      const int cy = tid / TX, cx = tid - cy * TX;
      if (cx < tw && cy < th) {
        const int cell = (lj0 + cy) * P.nx + (li0 + cx);
#pragma unroll
        for (int isp = 0; isp < WM_NSP_MAX; isp++)
          if (isp < P.nsp) {
            const int *cs = &s_cs[(isp * TY + cy) * (TX + 1) + cx];
            const int e = isp * WIN + (cy + 1) * WINX + (cx + 1);
            const int n = s_arr[e], ns = s_nst[isp * (TX * TY) + tid];
            t_base[e] = cs[0] + ns;
            t_end[e] = cs[1];
            a.cnt_tail[(size_t)isp * P.ncell + cell] = ns + n;  // k_place_rim adds the arrivals from other tiles
            // retire the slots the cell does not fill any more (k_mark_dead); rim arrivals may overwrite some later
            const int nold = a.cnt[(size_t)isp * P.ncell + cell];
            for (int pp = ns + n; pp < nold; pp++) a.src.x[(size_t)isp * P.cap + (size_t)(cs[0] + pp)] = dead_x();
          }
      }
    }
    __syncthreads();
    const double2 *const stage = reinterpret_cast<const double2 *>(a.dst.x.p);
#pragma unroll 1
    for (int rb = wid * RB; rb < nreg; rb += FW * RB) {
      // regions rb .. rb + RB - 1 of this warp: start and number of staged records (same bounds as the staging stores)
      int roff[RB], rcnt[RB], nmaxr = 0;
#pragma unroll
      for (int v = 0; v < RB; v++) {
        const int r = rb + v, isp = r / NQ, q = r - isp * NQ;
        const int cy = q / QX, cx0 = (q - cy * QX) * 4;
        roff[v] = 0;
        rcnt[v] = 0;
        if (cy < th && cx0 < tw) {
          const int *cs = &s_cs[(isp * TY + cy) * (TX + 1)];
          roff[v] = cs[cx0];
          rcnt[v] = min(s_nmv[r], cs[min(cx0 + 4, tw)] - cs[cx0]);
        }
        nmaxr = max(nmaxr, rcnt[v]);
      }
      const int isp = rb / NQ;  // RB regions never straddle the species (NQ % RB == 0)
      const size_t so = (size_t)isp * P.cap;
#pragma unroll 1
      for (int kb = 0; kb < nmaxr; kb += 64) {
        uint32_t tg[2 * RB];
        double2 r0[2 * RB], r1[2 * RB], r2[2 * RB];
#pragma unroll
        for (int u = 0; u < 2 * RB; u++) {
          const int k = kb + lane + 32 * (u & 1);
          tg[u] = TAG_DEAD;
          if (k < rcnt[u >> 1]) {
            const size_t ri = so + (size_t)(roff[u >> 1] + k);
            tg[u] = __ldcg(a.tag + ri);
            const double2 *r = stage + ri * 3;
            r0[u] = __ldcg(r);
            r1[u] = __ldcg(r + 1);
            r2[u] = __ldcg(r + 2);
          }
        }
#pragma unroll
        for (int u = 0; u < 2 * RB; u++) {
          const uint32_t t = tg[u];
          if (t == TAG_DEAD) continue;  // left the slab: already in the send buffer
          const int e = isp * WIN + (int)((t >> TAG_WSHIFT) & 0xff);
          const int bs = t_base[e];
          if (bs < 0) continue;  // rim of the window: k_place
          const int d = bs + (int)(t & TAG_RANK_MASK);
          if (d < t_end[e]) {
            double2 *o = a.src.word(so + (size_t)d);
            o[0] = r0[u];
            o[8] = r1[u];
            o[16] = r2[u];
          } else {  // segment full: park the record; the host rebuilds the layout after this step
            const int kk = atomicAdd(a.ovfcnt, 1);
            if (kk < a.ovfcap) {
              double *o = a.ovf + (size_t)kk * 6;
              o[0] = r0[u].x; o[1] = r0[u].y; o[2] = r1[u].x; o[3] = r1[u].y; o[4] = r2[u].x; o[5] = r2[u].y;
              a.ovfsp[kk] = isp;
            } else {
              atomicOr(a.err, ERR_OVERFLOW);
            }
          }
        }
      }
    }
  }
  // ---- hand the tile's arrival counts per window cell and its staged-record counts to k_place
  int *tb = a.tilebase + (size_t)tile * P.nsp * (2 * WIN);
  for (int e = tid; e < P.nsp * WIN; e += FT) {
    const int isp = e / WIN, w = e - isp * WIN;
    const int wy = w / WINX, wx = w - wy * WINX;
    const bool mine = TAIL && wx >= 1 && wx <= tw && wy >= 1 && wy <= th;  // placed by the tail above
    tb[isp * (2 * WIN) + w] = mine ? 0 : s_arr[e];
    if (w < NQ) tb[isp * (2 * WIN) + WIN + w] = s_nmv[isp * NQ + w];
  }
}

bool fused_sm_has_tail(int variant) { return variant != 3; }

void launch_fused_sm(const DevParams &P, const Pass1Args &a, int variant, cudaStream_t st, int ntiles) {
  const int grid = ntiles > 0 ? ntiles : P.ntx * P.nty;  // (a.tile0 = the first one)
  if (P.bc == WM_BC_SHOCK)
    k_fused_sm<3, true, 2, 2, true><<<grid, FT, 0, st>>>(P, a);
  else if (P.bc == WM_BC_RECONNECTION)
    k_fused_sm<3, true, 1, 2, true><<<grid, FT, 0, st>>>(P, a);
  else if (variant == 9)
    k_fused_sm<3, false, 0, 2, true><<<grid, FT, 0, st>>>(P, a);  // timing experiment: movers' current dropped
  else if (variant == 2)
    k_fused_sm<2, true, 0, 2, true><<<grid, FT, 0, st>>>(P, a);   // 2 CTAs per SM, up to 255 registers
  else if (variant == 3)
    k_fused_sm<3, true, 0, 2, false><<<grid, FT, 0, st>>>(P, a);  // every cell changer left to k_place
  else if (variant == 10)
    k_fused_sm<3, true, 0, 3, true><<<grid, FT, 0, st>>>(P, a);   // hints three iterations ahead
  else if (variant == 11)
    k_fused_sm<3, true, 0, 12, true><<<grid, FT, 0, st>>>(P, a);  // + the next list as a whole into L2 a pass ahead (measured: slower)
  else
    k_fused_sm<3, true, 0, 2, true><<<grid, FT, 0, st>>>(P, a);
}

}  // namespace wm
