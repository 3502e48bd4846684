// cg_persist_kernel.cu -- k_cg_persist: the three conjugate-gradient solves of field__fdtd_i (cgm, common/field.f90:319-461)
// as ONE persistent cooperative kernel with the whole iteration state on chip.
//
//  * One CTA per SM (1024 threads), each owning a rectangular block of bw x bh cells of the slab (at most CGP_K cells per
//    thread).  The search direction p of the block lives in shared memory with a one-cell halo ring, the residual r in
//    registers (CGP_K doubles per thread), phi in a dense per-CTA plane in global memory that only its owner thread
//    touches (an L2-resident spill; nothing on the critical path depends on it).  The components l = 1, 2, 3 are solved one
//    after the other, as the reference does, so the on-chip state is one component's.
//  * The loop is the reference's, literally (same exit test with its quirks, SURVEY F6 / field.f90:385-450), in the
//    two-phase form of k_cg_pap / k_cg_update2 (field_kernels.cu): phase A computes A p and the sums (r.r, p.Ap), phase B
//    updates phi and r and sums the new r.r; the p update p <- r + beta p is folded into the start of the next phase A.
//    The halo ring of the new p follows from the neighbours' new r (exchanged through the global array `rg`: every CTA
//    stores the r of its perimeter cells before the barrier of phase B) and the old halo p kept in shared memory --
//    the same expression the owner evaluates, so both sides hold identical values.
//  * Two grid barriers per iteration, each fused with the reduction of the dot products: every CTA writes its partial sums
//    and arrives at a global counter; after the barrier every CTA adds the G partials in the same fixed order, so all CTAs
//    (and all ranks) hold bit-identical alpha / beta and take the same branch.  No host round trip, no kernel launch, no
//    NCCL call inside the solve.
//  * More than one rank (y slabs on a ring, common/mpi_set.f90:36-47): the rows nys-1 / nye+1 of r are stored straight into
//    the neighbours' `rg` ghost rows over NVLink (CUDA IPC mappings), and the all-reduce of the sums is done in the kernel:
//    CTA 0 of every rank stores its slab's sums into every rank's CgpShared block and raises a sequence flag there
//    (st.release.sys); every CTA spins on the flags in its own memory (ld.acquire.sys) and adds the contributions in rank
//    order.  This replaces the two MPI_ALLREDUCE + two MPI_SENDRECV per iteration of field.f90:392,411,441 /
//    boundary_periodic.f90:511-568.
//  * Boundary rules of set_boundary_phi: periodic x = the wrapped column; conducting walls (boundary_reconnection.f90:557-577,
//    boundary_shock.f90:603-623) = left ghost -phi(nxs) for l = 1 and phi(nxs+1) for l = 2, 3, right ghost 0.
//
// Compiled with -fmad=false like field_kernels.cu: every expression keeps the reference's operation order; only the order of
// the global sums differs from the CPU path.
#include <cstddef>
#include <cstdio>
#include <cstdlib>

#include "kernels.h"

namespace wm {

namespace {

__device__ __forceinline__ size_t pidx(const DevParams &P, int li, int lj) { return (size_t)(lj + 2) * P.pitch + (li + 2); }

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned *p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long ld_acquire_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.acquire.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.release.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ void st_relaxed_sys(unsigned long long *p, unsigned long long v) {
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_relaxed_sys(const unsigned long long *p) {
  unsigned long long v;
  asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ int ld_volatile(const int *p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long v;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(v));
  return v;
}
// WM_CGTRACE=1: thread 0 of every CTA records the global timer at the points of barrier number a.trace_seq .. +3
__device__ __forceinline__ void trace(const CgpArgs &a, int cta, unsigned seq, int ev) {
  if (a.trace && seq >= a.trace_seq && seq < a.trace_seq + 4) a.trace[((size_t)cta * 4 + (seq - a.trace_seq)) * 4 + ev] = gtime();
}

constexpr long long SPIN_LIMIT = 6000000000LL;  // ~3 s of SM clocks: a peer that never shows up ends the solve with an error

struct Ctx {
  const CgpArgs *a;
  int G, cta;
  unsigned nsync;  // barriers passed so far (uniform over the grid)
  bool peer_writer;  // this CTA stores into the ring neighbours' ghost rows (its block touches the slab's first / last row)
  bool w_up, w_down; // ... into nup's lower ghost row / ndown's upper ghost row
  bool r_up, r_down; // this CTA reads the slab's upper / lower ghost row
  int bx;
};

// spin until pred() or abort / timeout
template <typename F>
__device__ __forceinline__ void spin_until(const CgpArgs &a, F pred) {
  const long long t0 = clock64();
  while (!pred()) {
    if (ld_volatile(a.abort)) break;
    if (clock64() - t0 > SPIN_LIMIT) {
      atomicExch(a.abort, 1);
      atomicOr(a.err, ERR_CG_TIMEOUT);
      break;
    }
  }
}

// Sum of NV values over all cells of all ranks.  In: every thread's private sums.  Out: the global sums, bit-identical in
// every thread of every CTA of every rank.  Also a grid-wide (and ring-wide) barrier with release / acquire semantics for
// the stores made before it (perimeter r values, peers' ghost rows).  `between` runs after this CTA has arrived and before
// it waits: work that nothing on the critical path depends on (the phi update).
constexpr int GMAX = 160;  // CTAs at most (one per SM)
template <int NV, typename F>
__device__ __forceinline__ void allsum(Ctx &c, double (&v)[NV], double *s_red, F between) {
  const CgpArgs &a = *c.a;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5, nw = blockDim.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; k++) s_red[wid * 2 + k] = v[k];
  __syncthreads();
  const unsigned seq = ++c.nsync;
  // Ring: this CTA's stores into a neighbour's ghost row are released by ONE flag per CTA in the neighbour's memory, off the
  // critical path of the sums (thread 32: warp 1, while warp 0 reduces).  st.release.sys is cumulative over the stores the
  // CTA barrier above has ordered before it.
  if (c.peer_writer && t == 32) {
    const unsigned long long gs = a.seq0 + seq;
    if (c.w_up) st_release_sys(&a.sh[a.nup]->halo_flag[0][c.bx], gs);      // nup's row nys-1 comes "from its ndown"
    if (c.w_down) st_release_sys(&a.sh[a.ndown]->halo_flag[1][c.bx], gs);  // ndown's row nye+1 comes "from its nup"
  }
  double2 *part = reinterpret_cast<double2 *>(a.partial) + (size_t)(seq & 1u) * GMAX;
  if (wid == 0) {
    double w[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0; k < NV; k++) {
      w[k] = (lane < nw) ? s_red[lane * 2 + k] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w[k] += __shfl_xor_sync(0xffffffffu, w[k], o);
    }
    if (lane == 0) {
      trace(a, c.cta, seq, 0);
      part[c.cta] = make_double2(w[0], w[1]);
      // release at gpu scope: the partial sums, and (through the CTA barrier above) every thread's perimeter stores, are
      // visible to whoever acquires the counter.  (A release reduction, not __threadfence() + atomicAdd: the latter is a
      // sequentially consistent fence, MEMBAR.SC, and costs about twice as much on the arrival path.)
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(a.bar) : "memory");
      trace(a, c.cta, seq, 1);
    }
  }
  between();
  if (t == 0) trace(a, c.cta, seq, 2);
  const unsigned target = seq * (unsigned)c.G;
  // the CTAs' partial sums, added in a fixed order by every warp on its own (no broadcast through shared memory)
  auto sum_partials = [&](double (&w)[2]) {
    double2 q[GMAX / 32];
#pragma unroll
    for (int j = 0; j < GMAX / 32; j++) {
      const int b = lane + 32 * j;
      q[j] = (b < c.G) ? __ldcg(&part[b]) : make_double2(0.0, 0.0);
    }
    w[0] = q[0].x;
    w[1] = q[0].y;
#pragma unroll
    for (int j = 1; j < GMAX / 32; j++) {
      w[0] += q[j].x;
      w[1] += q[j].y;
    }
#pragma unroll
    for (int k = 0; k < NV; k++) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w[k] += __shfl_xor_sync(0xffffffffu, w[k], o);
    }
  };
  if (a.nsize == 1) {
    if (t == 0) {
      spin_until(a, [&] { return ld_acquire_gpu(a.bar) >= target; });
      trace(a, c.cta, seq, 3);
    }
    __syncthreads();
    // one warp reads the partials (every warp of every CTA doing it made a hot spot of the few L2 lines that hold them)
    if (wid == 0) {
      double w[2];
      sum_partials(w);
      if (lane == 0) {
        s_red[64] = w[0];
        s_red[65] = w[1];
      }
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = s_red[64 + k];
  } else {
    // ring: CTA 0 adds the slab's partials and publishes them to every rank (itself included) as "LL" words: each half of a
    // double travels in one 8-byte store together with the low 32 bits of the barrier's sequence number, so the data IS the
    // flag (8-byte stores are single-copy atomic): no release fence, one NVLink trip.  Everybody polls the words in its own
    // memory.
    const unsigned long long gseq = a.seq0 + seq;
    const unsigned seq32 = (unsigned)gseq;
    const int slot = (int)(gseq & 3ull);
    if (c.cta == 0) {
      if (t == 0) spin_until(a, [&] { return ld_acquire_gpu(a.bar) >= target; });
      __syncthreads();
      if (wid == 0) {
        double w[2];
        sum_partials(w);
        if (lane < a.nsize) {
          unsigned long long *dst = &a.sh[lane]->ll[slot][a.nrank][0];
#pragma unroll
          for (int k = 0; k < NV; k++) {
            const unsigned long long b = (unsigned long long)__double_as_longlong(w[k]);
            st_relaxed_sys(dst + 2 * k, (b & 0xffffffffull) | ((unsigned long long)seq32 << 32));
            st_relaxed_sys(dst + 2 * k + 1, (b >> 32) | ((unsigned long long)seq32 << 32));
          }
        }
      }
    }
    CgpShared *me = a.sh[a.nrank];
    unsigned *halves = reinterpret_cast<unsigned *>(s_red + 66);  // [rank][2 * NV]
    if (t < a.nsize * 2 * NV) {
      const int q = t / (2 * NV), j = t - q * 2 * NV;
      const unsigned long long *src = &me->ll[slot][q][j];
      unsigned long long wv = 0;
      spin_until(a, [&] { wv = ld_relaxed_sys(src); return (unsigned)(wv >> 32) == seq32; });
      halves[t] = (unsigned)wv;
    }
    // the ghost rows this CTA reads: one flag per boundary CTA of the neighbour that owns them
    if (c.r_down && t >= 64 && t < 64 + a.cbx_down) spin_until(a, [&] { return ld_acquire_sys(&me->halo_flag[0][t - 64]) >= gseq; });
    if (c.r_up && t >= 256 && t < 256 + a.cbx_up) spin_until(a, [&] { return ld_acquire_sys(&me->halo_flag[1][t - 256]) >= gseq; });
    __syncthreads();
    double w[2] = {0.0, 0.0};
    for (int q = 0; q < a.nsize; q++) {  // rank order: the same sums on every rank
#pragma unroll
      for (int k = 0; k < NV; k++) {
        const unsigned lo = halves[q * 2 * NV + 2 * k], hi = halves[q * 2 * NV + 2 * k + 1];
        w[k] += __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
      }
    }
    __syncthreads();  // `halves` is reused by the next barrier
#pragma unroll
    for (int k = 0; k < NV; k++) v[k] = w[k];
  }
}

// element k (run-time) of a register array, by a chain of selects (no local-memory indexing)
template <int K>
__device__ __forceinline__ double r_at(const double (&r)[K], int k) {
  double v = r[0];
#pragma unroll
  for (int j = 1; j < K; j++) v = (k == j) ? r[j] : v;
  return v;
}

}  // namespace

// Thread layout: a block of bw x bh cells, bw * ng <= 1024 threads; thread (cx, g) owns the vertical run of rl = ceil(bh / ng)
// cells (cx, g * rl ...) of one column.  The tile is stored COLUMN-major with an odd column pitch cp >= bh + 2: a thread's run,
// the run of its left neighbour and the run of its right neighbour are three contiguous strips, so every shared-memory access
// of the stencil is `base register + immediate` (no address arithmetic in the loops), the stencil slides down the column with
// three loads per cell, and consecutive threads (consecutive columns) hit distinct banks.  The halo ring is served by the
// threads next to it: the first / last thread of a column takes the cell below / above, the threads of the first / last column
// the cells left / right of their rows.
template <int K, bool GUARD>
__device__ __forceinline__ void cg_phase_a(const double *q, const double *ql, const double *qr, int nrow, double f4, const double (&r)[K],
                                           double (&sa)[2]) {
  double up = q[-1], cc = q[0];
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (!GUARD || k < nrow) {
      const double dn = q[k + 1];
      const double av = -up - ql[k] + f4 * cc - qr[k] - dn;      // ap <- f4 p - N4 p                 field.f90:395-405
      const double rr = r[k];
      sa[0] = sa[0] + rr * rr;
      sa[1] = sa[1] + cc * av;
      up = cc;
      cc = dn;
    }
  }
}
template <int K, bool GUARD>
__device__ __forceinline__ void cg_phase_b(const double *q, const double *ql, const double *qr, int nrow, double f4, double alpha,
                                           double (&r)[K], double (&sb)[1]) {
  double up = q[-1], cc = q[0];
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (!GUARD || k < nrow) {
      const double dn = q[k + 1];
      const double av = -up - ql[k] + f4 * cc - qr[k] - dn;
      const double rr = r[k] - alpha * av;                        // r <- r - alpha ap                 field.f90:421-424
      r[k] = rr;
      sb[0] = sb[0] + rr * rr;
      up = cc;
      cc = dn;
    }
  }
}

template <int K>
__global__ void __launch_bounds__(CGP_T, 1) k_cg_persist(const __grid_constant__ DevParams P, const __grid_constant__ CgpArgs a) {
  extern __shared__ __align__(16) double T[];  // p (phi during the set-up) of the block with its halo ring: (bw+2) columns x cp
  __shared__ double s_red[32 * 2 + 2 + CGP_MAXR * 2];  // warp sums, the totals, the halves of the ranks' sums

  const int t = threadIdx.x;
  Ctx c;
  c.a = &a;
  c.G = gridDim.x;
  c.cta = blockIdx.x;
  c.nsync = 0;
  c.peer_writer = c.w_up = c.w_down = c.r_up = c.r_down = false;
  c.bx = 0;
  const int bx = c.cta % a.cbx, by = c.cta / a.cbx;
  const int x0 = (int)((long long)bx * P.nx / a.cbx), x1 = (int)((long long)(bx + 1) * P.nx / a.cbx);
  const int y0 = (int)((long long)by * P.nyl / a.cby), y1 = (int)((long long)(by + 1) * P.nyl / a.cby);
  const int bw = x1 - x0, bh = y1 - y0;
  const int cp = a.cp;                        // column pitch of the tile (odd, >= rows of the tallest block + 2)
  const int rl = a.rl;                        // rows per thread
  const int ngr = (bh + rl - 1) / rl;         // row groups in this block
  const int nthr = bw * ngr;                  // threads that own cells
  const bool own = t < nthr;
  const int cx = own ? t % bw : 0, cyb = own ? (t / bw) * rl : 0;
  const int nrow = own ? min(rl, bh - cyb) : 0;              // cells of this thread: (cx, cyb .. cyb + nrow - 1)
  double *const q = T + (cx + 1) * cp + (cyb + 1);            // its run; q[-1] / q[nrow]: the cells below / above
  double *const ql = q - cp, *const qr = q + cp;              // the same rows of the columns left / right
  const bool full = nrow == K;
  const size_t plane = (size_t)P.nx * P.nyl;
  const size_t pbase = (size_t)y0 * P.nx + (size_t)x0 * bh + t;  // dense packing: cell k of thread t at pbase + k * nthr
  const bool wall = P.bc != WM_BC_PERIODIC;
  const double f4 = P.f4;
  const int li_own = x0 + cx;
  // ---- the halo cells this thread serves, as offsets into the AoS3 array rg (component 0)
  const bool has_bot = own && cyb == 0, has_top = own && cyb + nrow == bh;
  const bool first_col = own && cx == 0, last_col = own && cx == bw - 1;
  int ljb = y0 - 1, ljt = y0 + bh;  // rows below / above the block; on one rank the ring neighbour is this slab itself
  if (P.nsize == 1) {
    if (ljb < 0) ljb += P.nyl;
    if (ljt >= P.nyl) ljt -= P.nyl;
  }
  const size_t g_bot = pidx(P, li_own, ljb) * 3, g_top = pidx(P, li_own, ljt) * 3;
  // left / right neighbours of the block's first / last column: the wrapped column, or a wall ghost (kind 1 / 2)
  int lil = x0 - 1, lir = x0 + bw, kindl = 0, kindr = 0;
  if (lil < 0) { if (wall) kindl = 1; else lil += P.nx; }
  if (lir >= P.nx) { if (wall) kindr = 2; else lir -= P.nx; }
  const bool slab_bot = P.nsize > 1 && y0 == 0, slab_top = P.nsize > 1 && y0 + bh == P.nyl;  // rows the ring neighbours need
  c.peer_writer = slab_bot || slab_top;
  c.w_up = slab_top;
  c.w_down = slab_bot;
  c.r_down = slab_bot;  // the block that touches the slab's first row reads the ghost row below it, written by ndown
  c.r_up = slab_top;
  c.bx = bx;

  // r of the perimeter cells goes to the global array (and to the ring neighbours' ghost rows)
  auto publish_one = [&](int l, int cy, double rr) {
    const int lj = y0 + cy;
    a.rg[pidx(P, li_own, lj) * 3 + l] = rr;
    if (slab_bot && cy == 0) a.r_down[pidx(P, li_own, a.nyl_down) * 3 + l] = rr;
    if (slab_top && cy == bh - 1) a.r_up[pidx(P, li_own, -1) * 3 + l] = rr;
  };
  // the halo ring from get(global offset of component l of the cell) and mix(value, old tile value): rows first, then columns
  auto fill_halo = [&](int l, auto get, auto mix) {
    double hb = 0.0, ht = 0.0;
    if (has_bot) hb = get(g_bot + l);
    if (has_top) ht = get(g_top + l);
    if (has_bot) q[-1] = mix(hb, q[-1]);
    if (has_top) q[nrow] = mix(ht, q[nrow]);
    if (first_col && kindl == 0)
      for (int k = 0; k < nrow; k++) ql[k] = mix(get(pidx(P, lil, y0 + cyb + k) * 3 + l), ql[k]);
    if (last_col && kindr == 0)
      for (int k = 0; k < nrow; k++) qr[k] = mix(get(pidx(P, lir, y0 + cyb + k) * 3 + l), qr[k]);
  };
  // wall ghosts from the block's own cells (after a barrier: the neighbour thread's column is read)   set_boundary_phi
  auto wall_halo = [&](int l) {
    if (!wall) return;
    if (first_col && kindl == 1)
      for (int k = 0; k < nrow; k++) ql[k] = (l == 0) ? -q[k] : qr[k];
    if (last_col && kindr == 2)
      for (int k = 0; k < nrow; k++) qr[k] = 0.0;
    __syncthreads();
  };

  // ---- set-up: phi <- df(l), b <- f5 gkl(l) for the three components into the per-thread planes   field.f90:349-360
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (k < nrow) {
      const size_t o = pidx(P, li_own, y0 + cyb + k);
#pragma unroll
      for (int l = 0; l < 3; l++) {
        a.phipl[l * plane + pbase + (size_t)k * nthr] = a.df[o * 6 + l];
        a.bpl[l * plane + pbase + (size_t)k * nthr] = P.f5 * a.gkl[o * 3 + l];
      }
    }
  }

  int stop = 0;
#pragma unroll 1
  for (int l = 0; l < 3; l++) {
    double *__restrict__ const phi = a.phipl + l * plane + pbase;
    const double *__restrict__ const bb = a.bpl + l * plane + pbase;
    double r[K];
    // ---- T <- phi with the halo ring (set_boundary_phi(phi), field.f90:367): neighbours' values are df(l) itself, rows of
    //      the ring neighbours are df's ghost rows (kept current by bc__dfield, field.f90:151,173)
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++)
      if (k < nrow) q[k] = phi[(size_t)k * nthr];
    fill_halo(l, [&](size_t g) { return a.df[(g - l) * 2 + l]; }, [](double v, double) { return v; });
    __syncthreads();
    wall_halo(l);
    // ---- r <- b + N4 phi - f4 phi, sum b^2, sum r^2                                               field.f90:362-383
    double s[2] = {0.0, 0.0};
    {
      double up = q[-1], cc = q[0];
#pragma unroll
      for (int k = 0; k < K; k++) {
        r[k] = 0.0;
        if (k < nrow) {
          const double dn = q[k + 1];
          const double b = bb[(size_t)k * nthr];
          const double rr = b + up + ql[k] - f4 * cc + qr[k] + dn;
          r[k] = rr;
          s[0] = s[0] + b * b;
          s[1] = s[1] + rr * rr;
          up = cc;
          cc = dn;
        }
      }
    }
    auto publish = [&]() {
      if (first_col || last_col) {
        for (int k = 0; k < nrow; k++) publish_one(l, cyb + k, r_at(r, k));
      } else {
        if (has_bot) publish_one(l, cyb, r[0]);
        if (has_top) publish_one(l, cyb + nrow - 1, r_at(r, nrow - 1));
      }
    };
    publish();
    allsum<2>(c, s, s_red, [] {});
    const double sumb = s[0];
    double sumr = s[1];
    const double eps = sqrt(sumb) * 1e-6;                                                        // field.f90:364
    int act = 0, ite = 0;
    if (sqrt(sumr) > eps) act = sumb > eps;  // the first test compares sum(b^2), not its sqrt      field.f90:385-387
    // ---- p <- r                                                                                  field.f90:379
    __syncthreads();
#pragma unroll
    for (int k = 0; k < K; k++)
      if (k < nrow) q[k] = r[k];
    fill_halo(l, [&](size_t g) { return __ldcg(&a.rg[g]); }, [](double v, double) { return v; });
    __syncthreads();
    wall_halo(l);

#pragma unroll 1
    while (act) {
      // ---- phase A: ap <- f4 p - N4 p, sums r.r and p.ap                                        field.f90:392-413
      double sa[2] = {0.0, 0.0};
      if (full)
        cg_phase_a<K, false>(q, ql, qr, nrow, f4, r, sa);
      else
        cg_phase_a<K, true>(q, ql, qr, nrow, f4, r, sa);
      allsum<2>(c, sa, s_red, [] {});
      sumr = sa[0];
      const double alpha = sumr / sa[1];                                                          // field.f90:415
      // ---- phase B: r -= alpha ap, sum of the new r.r; phi += alpha p while the barrier gathers    field.f90:417-441
      double sb1[1] = {0.0};
      if (full)
        cg_phase_b<K, false>(q, ql, qr, nrow, f4, alpha, r, sb1);
      else
        cg_phase_b<K, true>(q, ql, qr, nrow, f4, alpha, r, sb1);
      publish();
      allsum<1>(c, sb1, s_red, [&] {
        // loads first, in two batches: one L2 round trip per batch instead of one per cell
#pragma unroll
        for (int h = 0; h < 2; h++) {
          double ph[(K + 1) / 2];
#pragma unroll
          for (int k = 0; k < (K + 1) / 2; k++) {
            const int kk = h * ((K + 1) / 2) + k;
            ph[k] = (kk < K && kk < nrow) ? phi[(size_t)kk * nthr] : 0.0;
          }
#pragma unroll
          for (int k = 0; k < (K + 1) / 2; k++) {
            const int kk = h * ((K + 1) / 2) + k;
            if (kk < K && kk < nrow) phi[(size_t)kk * nthr] = ph[k] + alpha * q[kk];
          }
        }
      });
      // ---- loop control: the test uses the residual BEFORE this update                           field.f90:426-430,387
      ite++;
      if (ite >= 100) {
        stop = 1;
        act = 0;
      } else {
        act = sqrt(sumr) > eps;
      }
      if (ld_volatile(a.abort)) act = 0;
      if (!act) break;
      // ---- p <- r + beta p (field.f90:443-450), halo ring from the neighbours' new r and the old halo p
      const double beta = sb1[0] / sumr;
      {
        // the halo loads first: their L2 latency runs under the update of the own cells
        double hb = 0.0, ht = 0.0;
        if (has_bot) hb = __ldcg(&a.rg[g_bot + l]);
        if (has_top) ht = __ldcg(&a.rg[g_top + l]);
#pragma unroll
        for (int k = 0; k < K; k++)
          if (k < nrow) q[k] = r[k] + beta * q[k];
        if (has_bot) q[-1] = hb + beta * q[-1];
        if (has_top) q[nrow] = ht + beta * q[nrow];
        if (first_col && kindl == 0)
          for (int k = 0; k < nrow; k++) ql[k] = __ldcg(&a.rg[pidx(P, lil, y0 + cyb + k) * 3 + l]) + beta * ql[k];
        if (last_col && kindr == 0)
          for (int k = 0; k < nrow; k++) qr[k] = __ldcg(&a.rg[pidx(P, lir, y0 + cyb + k) * 3 + l]) + beta * qr[k];
      }
      __syncthreads();
      wall_halo(l);
    }
    if (c.cta == 0 && t == 0) a.out[l] = ite;
  }

  // ---- df(l) <- phi on the interior                                                              field.f90:455-457
#pragma unroll
  for (int k = 0; k < K; k++) {
    if (k < nrow) {
      const size_t o = pidx(P, li_own, y0 + cyb + k);
#pragma unroll
      for (int l = 0; l < 3; l++) a.df[o * 6 + l] = a.phipl[l * plane + pbase + (size_t)k * nthr];
    }
  }
  if (c.cta == 0 && t == 0) {
    a.out[3] = stop;
    a.out[4] = (int)c.nsync;
  }
}

// Block decomposition of an nx x nyl slab over at most nsm CTAs: cbx x cby blocks of bw x bh cells; a thread owns a vertical
// run of rl <= CGP_K cells, bw * ceil(bh / rl) <= CGP_T threads, and the tile (+ halo ring) fits `smem_max` bytes of shared
// memory.  Returns false if the slab is too large for the on-chip solver.
bool cgp_plan(int nx, int nyl, int nsm, size_t smem_max, int *cbx_out, int *cby_out, int *rl_out, size_t *smem_out) {
  long long best = -1;
  int bcx = 0, bcy = 0, brl = 0;
  size_t bsm = 0;
  if (nsm > 160) nsm = 160;  // GMAX of the kernel
  int fx = 0, fy = 0;        // WM_CGPLAN=cbx,cby forces a decomposition (experiments)
  if (const char *v = getenv("WM_CGPLAN")) sscanf(v, "%d,%d", &fx, &fy);
  for (int cbx = 1; cbx <= nsm && (cbx * 4 <= nx || cbx == 1); cbx++)
    for (int cby = 1; cbx * cby <= nsm && cby <= nyl; cby++) {
      const int bw = (nx + cbx - 1) / cbx, bh = (nyl + cby - 1) / cby;
      if (fx > 0 && (cbx != fx || cby != fy)) continue;
      if (bw > CGP_T) continue;
      int ng = CGP_T / bw;
      if (ng > bh) ng = bh;
      const int rl = (bh + ng - 1) / ng;
      if (rl > CGP_K) continue;
      const size_t sm = (size_t)(bw + 2) * (size_t)((bh + 2) | 1) * sizeof(double);  // column-major tile, odd column pitch
      if (sm > smem_max) continue;
      // fewest cells per thread first; among equals the fewest CTAs (a cheaper barrier), then the shorter perimeter
      const long long score = (long long)rl * 100000000LL + (long long)cbx * cby * 10000 + (bw + bh);
      if (best < 0 || score < best) {
        best = score;
        bcx = cbx;
        bcy = cby;
        brl = rl;
        bsm = sm;
      }
    }
  if (best < 0) return false;
  *cbx_out = bcx;
  *cby_out = bcy;
  *rl_out = brl;
  *smem_out = bsm;
  return true;
}

cudaError_t cgp_prepare(size_t smem) {
  cudaError_t e = cudaFuncSetAttribute(k_cg_persist<CGP_K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  e = cudaFuncSetAttribute(k_cg_persist<CGP_KM>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) return e;
  return cudaFuncSetAttribute(k_cg_persist<CGP_KS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t launch_cg_persist(const DevParams &P, const CgpArgs &a, size_t smem, cudaStream_t st) {
  DevParams Pc = P;
  CgpArgs ac = a;
  void *args[] = {(void *)&Pc, (void *)&ac};
  const void *fn = (a.rl <= CGP_KS)   ? (const void *)k_cg_persist<CGP_KS>
                   : (a.rl <= CGP_KM) ? (const void *)k_cg_persist<CGP_KM>
                                      : (const void *)k_cg_persist<CGP_K>;
  return cudaLaunchCooperativeKernel(fn, dim3(a.cbx * a.cby), dim3(CGP_T), args, smem, st);
}

}  // namespace wm
