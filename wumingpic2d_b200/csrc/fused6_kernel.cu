// fused6_kernel.cu -- k_fused_dp: the fused particle pass with DIRECT PLACEMENT of the cell changers ("ping-pong" stores).
//
//  Same physics, arithmetic and deposit as k_fused_sm (fused5_kernel.cu): particle__solv (common/particle.f90:83-169) +
//  ele_cur split into stayers / movers (common/field.f90:189-316) + bc__particle_x/_y (common/boundary_periodic.f90:61-248) +
//  sort__bucket (common/sort.f90:36-82).  What is different is the sort: the pass reads store A and writes store B.
//   * A cell's segment of B is filled from both ends: the stayers from the front (slot = start + rank among the stayers),
//     the arrivals from other cells of the same tile from the back (slot = start + capacity - 1 - rank, the rank drawn from a
//     shared-memory counter).  Every record is written ONCE, at its final place: no staging, no read-back, no tail of the
//     CTA, no in-place hazards.  The next step reads the two ranges [0, nf) and [cap - nb, cap) of a segment.
//   * Only the arrivals of OTHER tiles' cells (the window's rim: 12 % of the cell changers, 1.8 % of the particles) are
//     staged -- in the memory of the tag array, in the shadow of their quad -- and appended behind the stayers by
//     k_place_rim2 once every tile has written its stayer counts.
//   * Kernels that walk a store by its liveness marks (download, moments, energy ...) see a normalised state: k_normalize
//     moves the back range behind the front range and marks the rest dead; only diagnostics steps pay for it.
//  Against k_fused_sm<TAIL> this removes the tail's 4.4 GB of DRAM traffic per launch (staged records that had left the L2
//  before they were read back) and all but 1.8 % of the work of k_place.
#include <algorithm>
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

// MINB resident CTAs per SM: 3 -> at most 168 registers.
template <int MINB, int WALL, int PFD>
__global__ void __launch_bounds__(FT, MINB) k_fused_dp(const DevParams P, const Pass1Args a) {
  constexpr bool DRAIN = true;
  __shared__ __align__(128) double s_f[WINY * WINX * 6];
  __shared__ __align__(16) double s_j[3 * JY * JX];
  __shared__ __align__(16) double2 s_q[FW * 4 * QCAP * 3];  // [warp][cell of the quad][slot] x (hx hy | dxn dyn | qvz qf)
  __shared__ int s_arr[WM_NSP_MAX * WIN];
  __shared__ int s_cs[WM_NSP_MAX * TY * (TX + 1)];  // segment offsets of the tile's cells (+ one column: the end of the last one)
  __shared__ int s_nold[WM_NSP_MAX * TX * TY];      // particles a cell holds now: its stayers cannot reach beyond this slot
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
  for (int e = tid; e < P.nsp * TY * (TX + 1); e += FT) {
    const int isp = e / (TY * (TX + 1)), r = e - isp * (TY * (TX + 1)), cy = r / (TX + 1), cx = r - cy * (TX + 1);
    int v = 0;
    if (cy < th && cx <= tw) v = a.cstart[(size_t)isp * (P.ncell + 1) + (size_t)(lj0 + cy) * P.nx + (li0 + cx)];
    s_cs[e] = v;
  }
  for (int e = tid; e < P.nsp * TX * TY; e += FT) {
    const int isp = e / (TX * TY), r = e - isp * (TX * TY), cy = r / TX, cx = r - cy * TX;
    int v = 0;
    if (cy < th && cx < tw) {
      const size_t ci = (size_t)isp * P.ncell + (size_t)(lj0 + cy) * P.nx + (li0 + cx);
      v = a.cnt[ci] + a.cntb[ci];
    }
    s_nold[e] = v;
  }
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
  double *const pdx = a.dst.x.p;  // the store this pass writes
  const double qf_base = P.delx / P.delt;
  const double delt = P.delt, inv_cc = P.inv_cc, cc = P.cc;
  int myqi = (wid * 4 + grp) * (QCAP * 3);
  pin(myqi);
  double2 *const myq = &s_q[myqi];
  unsigned myq_a = smem_u32(myq), sarr_a = smem_u32(s_arr);  // 32-bit shared addresses for the stores / atomics of the loop
  pin(myq_a);
  pin(sarr_a);

  // per (cell, species): nb* = segment start, nf* = particles at the front of the segment, nc* = front + back
  int nb0 = 0, nc0 = 0, nf0 = 0, nb1 = 0, nc1 = 0, nf1 = 0;
  {
    const int cy = wid / QX, cx = (wid - cy * QX) * 4 + grp;
    if (cx < tw && cy < th) {
      const int cell = (lj0 + cy) * P.nx + (li0 + cx);
      nb0 = a.cstart[cell];
      nf0 = a.cnt[cell];
      nc0 = nf0 + a.cntb[cell];
      if (P.nsp > 1) {
        nb1 = a.cstart[(size_t)(P.ncell + 1) + cell];
        nf1 = a.cnt[(size_t)P.ncell + cell];
        nc1 = nf1 + a.cntb[(size_t)P.ncell + cell];
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
    const int beg0 = nb0, cnt0 = nc0, fr0 = nf0, beg1 = nb1, cnt1 = nc1, fr1 = nf1;
    {
      const int qn_ = q + FW;
      const int cyn = qn_ / QX, cxn = (qn_ - cyn * QX) * 4 + grp;
      nb0 = nc0 = nf0 = nb1 = nc1 = nf1 = 0;
      if (qn_ < NQ && cxn < tw && cyn < th) {
        const int celln = (lj0 + cyn) * P.nx + (li0 + cxn);
        nb0 = a.cstart[celln];
        nf0 = a.cnt[celln];
        nc0 = nf0 + a.cntb[celln];
        if (P.nsp > 1) {
          nb1 = a.cstart[(size_t)(P.ncell + 1) + celln];
          nf1 = a.cnt[(size_t)P.ncell + celln];
          nc1 = nf1 + a.cntb[(size_t)P.ncell + celln];
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
      // the cell's particles are its front range [0, nfr) and its back range [capc - (ntot - nfr), capc): particle k of the
      // cell sits in slot k (k < nfr) or k + jump
      const int beg = isp ? beg1 : beg0, ntot = isp ? cnt1 : cnt0, nfr = isp ? fr1 : fr0;
      const int jump = valid ? (s_cs[(isp * TY + cy) * (TX + 1) + cx + 1] - beg) - ntot : 0;
      int nmax = ntot;
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

      // staging area of the quad's rim movers (cell changers whose new cell belongs to another tile): the bytes of the tag
      // array under the quad's slot range -- a 16-byte header (the count) and 48-byte records
      double2 *rimrec = nullptr;
      int rimcap = 0;
      if (cy < th && (q - cy * QX) * 4 < tw) {
        const int x4 = (q - cy * QX) * 4;
        const int *scs = &s_cs[(isp * TY + cy) * (TX + 1)];
        const long long s0 = so_slots(P, isp) + scs[x4];
        const int nsl = scs[min(x4 + 4, tw)] - scs[x4];
        rimrec = reinterpret_cast<double2 *>(a.tag + s0) + 1;
        rimcap = nsl >= 16 ? (4 * nsl - 16) / 48 : 0;
      }
      int kk = l8 + k0;  // index of the lane's current particle in the cell's list
      double2 *pbase = reinterpret_cast<double2 *>(px + 6 * ((size_t)isp * P.cap));  // slot 0 of this species (cap % 8 == 0)
      double2 *dbase = reinterpret_cast<double2 *>(pdx + 6 * ((size_t)isp * P.cap));
      pin(pbase);
      __builtin_assume(__isGlobal(pbase));
      __builtin_assume(__isGlobal(dbase));
      const double2 *segw = pbase + pslot_w((size_t)beg);    // word 0 of the segment's slot 0 (segments are whole blocks)
      auto slot_w = [&](int k) -> const double2 * {
        const int sl = k + (k >= nfr ? jump : 0);
        return segw + ((sl >> 3) * 24 + (sl & 7));
      };
      const double2 *pl = slot_w(kk);                        // word 0 of the lane's current slot; + 8, + 16: words 1, 2
      const int w0 = isp * WIN + (cy + 1) * WINX + (cx + 1);  // this cell in the window of arrival counters
      // Latency hiding without registers: the lines of the iterations k + 1 and k + 2 are pulled into L1 by prefetch
      // hints (PFD iterations ahead in steady state); the loads at the top of iteration k are L1 hits.  Hints for
      // iteration k0 were issued while the previous species / quad was being worked on.
      if (kk + 8 < ntot) {
        const double2 *pp = slot_w(kk + 8);
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pp));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pp + 8));
        asm volatile("prefetch.global.L1 [%0];" ::"l"(pp + 16));
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
      }
      int k = k0;
      for (; k < nmax; k += 8) {
        const int kc = kk;
        const bool active = kc < ntot;
        double xn, yn, un1, un2, un3;
        double hx, hy, dxn, dyn, qvz;
        double sxm, sx0, sxp, sym, sy0, syp;
        double idc;  // the id moves with the record (bit pattern)
        if (kc + 8 * PFD < ntot) {
          const double2 *pp = slot_w(kc + 8 * PFD);
          asm volatile("prefetch.global.L1 [%0];" ::"l"(pp));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(pp + 8));
          asm volatile("prefetch.global.L1 [%0];" ::"l"(pp + 16));
        }
        if (kc >= nfr && kc - 8 < nfr) pl = slot_w(kc);  // this lane crosses from the front range into the back range
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
        kk += 8;
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
        // ---- sort: every record is written once, into the other store                sort.f90:57-75
        const bool chg = active && !sstay;
        int incx = 0, incy = 0;
        bool leave = false;
        if (chg) {
          // cell changer.  |move| < 1 cell (CFL), so the new cell is (gi + incx, gj + incy); anything else
          // is an error (also catches NaN)
          if (!(fabs(dxn) < 1.5 && fabs(dyn) < 1.5)) atomicOr(a.err, ERR_MOVED_TOO_FAR);
          incx = (int)(dxn >= 0.5) - (int)(dxn < -0.5);
          incy = (int)(dyn >= 0.5) - (int)(dyn < -0.5);
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
          leave = P.nsize > 1 && (j2 < P.nys || j2 >= P.nys + P.nyl);
        }
        const int wx = cx + 1 + incx, wy = cy + 1 + incy;  // destination in the tile's window
        const bool intile = chg && !leave && wx >= 1 && wx <= tw && wy >= 1 && wy <= th;
        const bool rim = chg && !leave && !intile;
        const unsigned bal = __ballot_sync(0xffffffffu, sstay);
        const unsigned balr = __ballot_sync(0xffffffffu, rim);
        const unsigned m8 = (bal >> gsh) & 0xffu;
        if (sstay) {
          // the stayers fill the front of the cell's segment in the order they are met
          double2 *d = dbase + pslot_w((size_t)(beg + nst + __popc(m8 & below)));
          d[0] = make_double2(xn, yn);
          d[8] = make_double2(un1, un2);
          d[16] = make_double2(un3, idc);  // the id moves with the record (bit pattern)
        } else if (leave) {
          // record goes to the neighbour's edge row        boundary_periodic.f90:156-161,174-189
          const int dir = (gj + incy < P.nys) ? 0 : 1;
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
        } else if (intile) {
          // new cell in this tile: from the back of its segment, rank from one shared-memory integer atomic.  The stayers of
          // that cell (at most the s_nold particles it holds now) fill it from the front: no overlap below slot s_nold
          unsigned rk;
          asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(rk) : "r"(sarr_a + (unsigned)(w0 + incy * WINX + incx) * 4u) : "memory");
          const int *scs = &s_cs[(isp * TY + (wy - 1)) * (TX + 1) + (wx - 1)];
          const int bd = scs[0], sl = scs[1] - bd - 1 - (int)rk;
          if (sl >= s_nold[isp * (TX * TY) + (wy - 1) * TX + (wx - 1)]) {
            double2 *d = dbase + pslot_w((size_t)(bd + sl));
            d[0] = make_double2(xn, yn);
            d[8] = make_double2(un1, un2);
            d[16] = make_double2(un3, idc);
          } else {  // segment full: park the record; the host rebuilds the layout after this step
            const int kq = atomicAdd(a.ovfcnt, 1);
            if (kq < a.ovfcap) {
              double *o = a.ovf + (size_t)kq * 6;
              o[0] = xn; o[1] = yn; o[2] = un1; o[3] = un2; o[4] = un3; o[5] = idc;
              a.ovfsp[kq] = isp;
            } else {
              atomicOr(a.err, ERR_OVERFLOW);
            }
          }
        } else if (rim) {
          // new cell in another tile: staged for k_place_rim2, ballot-ranked so that a warp's stores are contiguous
          const int sk = nmv + __popc(balr & lanelt);
          if (sk < rimcap) {
            double2 *d = rimrec + (size_t)sk * 3;
            d[0] = make_double2(xn, yn);
            d[1] = make_double2(un1, un2);
            d[2] = make_double2(un3, idc);
          } else {
            const int kq = atomicAdd(a.ovfcnt, 1);
            if (kq < a.ovfcap) {
              double *o = a.ovf + (size_t)kq * 6;
              o[0] = xn; o[1] = yn; o[2] = un1; o[3] = un2; o[4] = un3; o[5] = idc;
              a.ovfsp[kq] = isp;
            } else {
              atomicOr(a.err, ERR_OVERFLOW);
            }
          }
        }
        nst += __popc(m8);
        nmv += __popc(balr);
        if (DRAIN && __any_sync(0xffffffffu, qn > QCAP - 8)) {  // the next iteration could overflow a queue: drain first
          k += 8;
          full = true;
          break;
        }
      }
      if (k >= nmax) {
        if (valid && l8 == 0) a.cnt_tail[(size_t)isp * P.ncell + cell] = nst;  // front count of the new store (rim arrivals follow)
        if (lane == 0 && rimrec) *reinterpret_cast<int *>(rimrec - 1) = min(nmv, rimcap);  // header of the quad's rim staging area
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
  // ---- the tile's cells: how many arrivals their back ranges received (the front counts were written by their owners)
  for (int e = tid; e < P.nsp * TX * TY; e += FT) {
    const int isp = e / (TX * TY), r = e - isp * (TX * TY), cy = r / TX, cx = r - cy * TX;
    if (cy < th && cx < tw) {
      // (arrivals that found the segment full -- the highest ranks -- went to the overflow list)
      const int *scs = &s_cs[(isp * TY + cy) * (TX + 1) + cx];
      a.cntb_new[(size_t)isp * P.ncell + (size_t)(lj0 + cy) * P.nx + (li0 + cx)] =
          min(s_arr[isp * WIN + (cy + 1) * WINX + (cx + 1)], max(scs[1] - scs[0] - s_nold[e], 0));
    }
  }
}

// ---- the rest of the sort: arrivals from other tiles.  One thread per (species, row, quad of 4 cells): the records its quad
//      staged in the tag array's memory are appended behind the stayers of their new cells (one global atomic per record:
//      1.8 % of the particles), unless the segment is full (-> overflow list, layout rebuild).
__global__ void __launch_bounds__(128) k_place_rim2(const DevParams P, const uint32_t *__restrict__ tag, const PartSoA dst,
                                                    const int *__restrict__ cstart, int *cnt_new, const int *__restrict__ cntb_new,
                                                    double *ovf, int *ovfsp, int *ovfcnt, int ovfcap, unsigned *err) {
  const int nq = (P.nx + 3) / 4;
  const long long nreg = (long long)P.nsp * P.nyl * nq;
  for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < nreg; t += (long long)gridDim.x * blockDim.x) {
    const int isp = (int)(t / ((long long)P.nyl * nq));
    const int r = (int)(t - (long long)isp * P.nyl * nq), lj = r / nq, m = r - lj * nq;
    const int *cs = cstart + (size_t)isp * (P.ncell + 1);
    const int c0 = lj * P.nx + 4 * m;
    // the quad's region ends with the tile (tiles are TX cells wide) or the row
    const int c1 = min(min(c0 + 4, (4 * m / TX + 1) * TX + lj * P.nx), (lj + 1) * P.nx);
    const int nsl = cs[c1] - cs[c0];
    if (nsl < 16) continue;
    const double2 *rec = reinterpret_cast<const double2 *>(tag + so_slots(P, isp) + cs[c0]);
    const int n = min(*reinterpret_cast<const int *>(rec), (4 * nsl - 16) / 48);
    rec += 1;
    for (int k = 0; k < n; k++) {
      const double2 r0 = rec[3 * k], r1 = rec[3 * k + 1], r2 = rec[3 * k + 2];
      const int li = __double2int_rz(r0.x) - P.nxgs, lj2 = __double2int_rz(r0.y) - P.nys;
      if (li < 0 || li >= P.nx || lj2 < 0 || lj2 >= P.nyl) {
        atomicOr(err, ERR_BAD_CELL);
        continue;
      }
      const int cell = lj2 * P.nx + li;
      const size_t ci = (size_t)isp * P.ncell + cell;
      const int pos = atomicAdd(&cnt_new[ci], 1);
      if (pos < cs[cell + 1] - cs[cell] - cntb_new[ci]) {
        double2 *o = dst.word((size_t)isp * P.cap + (size_t)(cs[cell] + pos));
        o[0] = r0;
        o[8] = r1;
        o[16] = r2;
      } else {
        const int kk = atomicAdd(ovfcnt, 1);
        if (kk < ovfcap) {
          double *o = ovf + (size_t)kk * 6;
          o[0] = r0.x; o[1] = r0.y; o[2] = r1.x; o[3] = r1.y; o[4] = r2.x; o[5] = r2.y;
          ovfsp[kk] = isp;
        } else {
          atomicOr(err, ERR_OVERFLOW);
        }
      }
    }
  }
}

// ---- two ranges -> one: the back range of every segment moves behind its front range, the rest of the segment is marked
//      dead, so that the kernels that walk a store by its liveness marks (download, moments, energy, layout rebuild ...) and
//      the other particle kernels see the layout they know.  A warp per (cell, species).
__global__ void __launch_bounds__(256) k_normalize(const DevParams P, const PartSoA st, const int *__restrict__ cstart, int *cnt, int *cntb) {
  const long long n = (long long)P.nsp * P.ncell;
  const int lane = threadIdx.x & 31;
  for (long long wk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wk < n; wk += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int isp = (int)(wk / P.ncell), cell = (int)(wk - (long long)isp * P.ncell);
    const int *cs = cstart + (size_t)isp * (P.ncell + 1);
    const int beg = cs[cell], capc = cs[cell + 1] - beg;
    const int nb = min(cntb[wk], capc), nf = min(cnt[wk], capc - nb);  // (appends that found the segment full went to the overflow list)
    const size_t so = (size_t)isp * P.cap + beg;
    // ascending copy, a warp-wide batch at a time (all loads of a batch before its stores): source index >= target index
    for (int b = 0; b < nb; b += 32) {
      const int i = b + lane;
      double2 r0, r1, r2;
      if (i < nb) {
        const double2 *s = st.word(so + (size_t)(capc - nb + i));
        r0 = s[0]; r1 = s[8]; r2 = s[16];
      }
      __syncwarp();
      if (i < nb) {
        double2 *d = st.word(so + (size_t)(nf + i));
        d[0] = r0; d[8] = r1; d[16] = r2;
      }
      __syncwarp();
    }
    for (int p = nf + nb + lane; p < capc; p += 32) st.x[so + p] = dead_x();
    __syncwarp();
    if (lane == 0) {
      cnt[wk] = nf + nb;
      cntb[wk] = 0;
    }
  }
}

void launch_fused_dp(const DevParams &P, const Pass1Args &a, cudaStream_t st) {
  const int grid = P.ntx * P.nty;
  if (P.bc == WM_BC_SHOCK)
    k_fused_dp<3, 2, 2><<<grid, FT, 0, st>>>(P, a);
  else if (P.bc == WM_BC_RECONNECTION)
    k_fused_dp<3, 1, 2><<<grid, FT, 0, st>>>(P, a);
  else
    k_fused_dp<3, 0, 2><<<grid, FT, 0, st>>>(P, a);
}
void launch_place_rim2(const DevParams &P, const uint32_t *tag, const PartSoA &dst, const int *cstart, int *cnt_new, const int *cntb_new,
                       double *ovf, int *ovfsp, int *ovfcnt, int ovfcap, unsigned *err, cudaStream_t st) {
  const long long nreg = (long long)P.nsp * P.nyl * ((P.nx + 3) / 4);
  const int nb = (int)std::min<long long>((nreg + 127) / 128, 148LL * 64);
  k_place_rim2<<<nb, 128, 0, st>>>(P, tag, dst, cstart, cnt_new, cntb_new, ovf, ovfsp, ovfcnt, ovfcap, err);
}
void launch_normalize(const DevParams &P, const PartSoA &st_, const int *cstart, int *cnt, int *cntb, cudaStream_t st) {
  k_normalize<<<148 * 16, 256, 0, st>>>(P, st_, cstart, cnt, cntb);
}

}  // namespace wm
