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
__device__ __forceinline__ int ld_volatile(const int *p) {
  int v;
  asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}

constexpr long long SPIN_LIMIT = 6000000000LL;  // ~3 s of SM clocks: a peer that never shows up ends the solve with an error

struct Ctx {
  const CgpArgs *a;
  int G, cta;
  unsigned nsync;  // barriers passed so far (uniform over the grid)
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
// the stores made before it (perimeter r values, peers' ghost rows).
template <int NV>
__device__ void allsum(Ctx &c, double (&v)[NV], double *s_red, double *s_tot) {
  const CgpArgs &a = *c.a;
  const int t = threadIdx.x, lane = t & 31, wid = t >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v[k] += __shfl_xor_sync(0xffffffffu, v[k], o);
  }
  if (lane == 0)
#pragma unroll
    for (int k = 0; k < NV; k++) s_red[wid * 4 + k] = v[k];
  if (a.nsize > 1) __threadfence_system();  // my stores into the neighbours' ghost rows
  __syncthreads();
  const unsigned seq = ++c.nsync;
  double *part = a.partial + (size_t)(seq & 1u) * c.G * 4;
  if (wid == 0) {
    double w[NV];
#pragma unroll
    for (int k = 0; k < NV; k++) {
      w[k] = (lane < CGP_T / 32) ? s_red[lane * 4 + k] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) w[k] += __shfl_xor_sync(0xffffffffu, w[k], o);
    }
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < NV; k++) part[c.cta * 4 + k] = w[k];
      __threadfence();
      atomicAdd(a.bar, 1u);
    }
  }
  const unsigned target = seq * (unsigned)c.G;
  if (a.nsize == 1) {
    if (t == 0) spin_until(a, [&] { return ld_acquire_gpu(a.bar) >= target; });
    __syncthreads();
    if (wid == 0) {
      double w[NV];
#pragma unroll
      for (int k = 0; k < NV; k++) w[k] = 0.0;
      for (int b = lane; b < c.G; b += 32) {
#pragma unroll
        for (int k = 0; k < NV; k++) w[k] += __ldcg(&part[b * 4 + k]);
      }
#pragma unroll
      for (int k = 0; k < NV; k++) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) w[k] += __shfl_xor_sync(0xffffffffu, w[k], o);
      }
      if (lane == 0)
#pragma unroll
        for (int k = 0; k < NV; k++) s_tot[k] = w[k];
    }
    __syncthreads();
  } else {
    // ring: CTA 0 adds the slab's partials and publishes them to every rank (itself included); everybody waits for the
    // flags of all ranks in its own memory
    const unsigned long long gseq = a.seq0 + seq;
    const int slot = (int)(gseq & 3ull);
    if (c.cta == 0) {
      if (t == 0) spin_until(a, [&] { return ld_acquire_gpu(a.bar) >= target; });
      __syncthreads();
      if (wid == 0) {
        double w[NV];
#pragma unroll
        for (int k = 0; k < NV; k++) w[k] = 0.0;
        for (int b = lane; b < c.G; b += 32) {
#pragma unroll
          for (int k = 0; k < NV; k++) w[k] += __ldcg(&part[b * 4 + k]);
        }
#pragma unroll
        for (int k = 0; k < NV; k++) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) w[k] += __shfl_xor_sync(0xffffffffu, w[k], o);
        }
        if (lane < a.nsize) {
          CgpShared *dst = a.sh[lane];
#pragma unroll
          for (int k = 0; k < NV; k++) dst->xsum[slot][a.nrank][k] = w[k];
          __threadfence_system();
          st_release_sys(&dst->flag[a.nrank], gseq);
        }
      }
    }
    CgpShared *me = a.sh[a.nrank];
    if (t < a.nsize) spin_until(a, [&] { return ld_acquire_sys(&me->flag[t]) >= gseq; });
    __syncthreads();
    if (t == 0) {
      double w[NV];
#pragma unroll
      for (int k = 0; k < NV; k++) w[k] = 0.0;
      for (int q = 0; q < a.nsize; q++) {
#pragma unroll
        for (int k = 0; k < NV; k++) w[k] += __ldcg(&me->xsum[slot][q][k]);
      }
#pragma unroll
      for (int k = 0; k < NV; k++) s_tot[k] = w[k];
    }
    __syncthreads();
  }
#pragma unroll
  for (int k = 0; k < NV; k++) v[k] = s_tot[k];
  __syncthreads();  // s_red / s_tot are reused by the next call
}

}  // namespace

__global__ void __launch_bounds__(CGP_T, 1) k_cg_persist(const __grid_constant__ DevParams P, const __grid_constant__ CgpArgs a) {
  extern __shared__ __align__(16) double T[];  // p (phi during the set-up) of the block with its halo ring: (bw+2) x (bh+2)
  __shared__ double s_red[32 * 4];
  __shared__ double s_tot[4];

  const int t = threadIdx.x;
  Ctx c;
  c.a = &a;
  c.G = gridDim.x;
  c.cta = blockIdx.x;
  c.nsync = 0;
  const int bx = c.cta % a.cbx, by = c.cta / a.cbx;
  const int x0 = (int)((long long)bx * P.nx / a.cbx), x1 = (int)((long long)(bx + 1) * P.nx / a.cbx);
  const int y0 = (int)((long long)by * P.nyl / a.cby), y1 = (int)((long long)(by + 1) * P.nyl / a.cby);
  const int bw = x1 - x0, bh = y1 - y0, tp = bw + 2, ncl = bw * bh;
  const size_t pbase = (size_t)y0 * P.nx + (size_t)x0 * bh;  // dense packing of the blocks: row of blocks, then block
  const size_t plane = (size_t)P.nx * P.nyl;
  const int dq = CGP_T / bw, dr = CGP_T - dq * bw;           // (cx, cy) of cell e + CGP_T from those of cell e
  const int cx0 = t % bw, cy0 = t / bw;
  const bool wall = P.bc != WM_BC_PERIODIC;
  const double f4 = P.f4;
  const int nhalo = 2 * (bw + bh);

  // halo position h -> tile index, local cell (li, lj) it mirrors, kind: 0 = a cell of this slab or of a ring neighbour
  // (read from the global array), 1 = left wall ghost, 2 = right wall ghost
  auto halo_pos = [&](int h, int &ti, int &li, int &lj, int &kind) {
    int hx, hy;
    if (h < bw) { hx = h; hy = -1; }
    else if (h < 2 * bw) { hx = h - bw; hy = bh; }
    else if (h < 2 * bw + bh) { hx = -1; hy = h - 2 * bw; }
    else { hx = bw; hy = h - 2 * bw - bh; }
    ti = (hy + 1) * tp + hx + 1;
    li = x0 + hx;
    lj = y0 + hy;
    kind = 0;
    if (li < 0) {
      if (wall) kind = 1; else li += P.nx;
    } else if (li >= P.nx) {
      if (wall) kind = 2; else li -= P.nx;
    }
    if (P.nsize == 1) {  // the ring neighbour is this slab itself
      if (lj < 0) lj += P.nyl; else if (lj >= P.nyl) lj -= P.nyl;
    }
  };

  // ---- set-up: phi <- df(l), b <- f5 gkl(l) for the three components into the per-thread planes   field.f90:349-360
#pragma unroll
  for (int k = 0, cx = cx0, cy = cy0; k < CGP_K; k++) {
    const int e = t + k * CGP_T;
    if (e < ncl) {
      const size_t o = pidx(P, x0 + cx, y0 + cy);
#pragma unroll
      for (int l = 0; l < 3; l++) {
        a.phipl[l * plane + pbase + e] = a.df[o * 6 + l];
        a.bpl[l * plane + pbase + e] = P.f5 * a.gkl[o * 3 + l];
      }
    }
    cx += dr; cy += dq;
    if (cx >= bw) { cx -= bw; cy++; }
  }

  int stop = 0;
#pragma unroll 1
  for (int l = 0; l < 3; l++) {
    double *const phi = a.phipl + l * plane + pbase;
    const double *const bb = a.bpl + l * plane + pbase;
    double r[CGP_K];
    // ---- T <- phi with the halo ring (set_boundary_phi(phi), field.f90:367): neighbours' values are df(l) itself, rows of
    //      the ring neighbours are df's ghost rows (kept current by bc__dfield, field.f90:151,173)
    __syncthreads();
#pragma unroll
    for (int k = 0, cx = cx0, cy = cy0; k < CGP_K; k++) {
      const int e = t + k * CGP_T;
      if (e < ncl) T[e + 2 * cy + tp + 1] = phi[e];
      cx += dr; cy += dq;
      if (cx >= bw) { cx -= bw; cy++; }
    }
    __syncthreads();
    for (int h = t; h < nhalo; h += CGP_T) {
      int ti, li, lj, kind;
      halo_pos(h, ti, li, lj, kind);
      double v;
      if (kind == 0) v = a.df[pidx(P, li, lj) * 6 + l];
      else if (kind == 1) v = (l == 0) ? -T[ti + 1] : T[ti + 2];
      else v = 0.0;
      T[ti] = v;
    }
    __syncthreads();
    // ---- r <- b + N4 phi - f4 phi, sum b^2, sum r^2                                               field.f90:362-383
    double s[2] = {0.0, 0.0};
#pragma unroll
    for (int k = 0, cx = cx0, cy = cy0; k < CGP_K; k++) {
      const int e = t + k * CGP_T;
      r[k] = 0.0;
      if (e < ncl) {
        const int i = e + 2 * cy + tp + 1;
        const double b = bb[e];
        const double rr = b + T[i - tp] + T[i - 1] - f4 * T[i] + T[i + 1] + T[i + tp];
        r[k] = rr;
        s[0] = s[0] + b * b;
        s[1] = s[1] + rr * rr;
        if (cx == 0 || cx == bw - 1 || cy == 0 || cy == bh - 1) {
          const int li = x0 + cx, lj = y0 + cy;
          a.rg[pidx(P, li, lj) * 3 + l] = rr;
          if (P.nsize > 1) {
            if (lj == 0) a.r_down[pidx(P, li, a.nyl_down) * 3 + l] = rr;
            if (lj == P.nyl - 1) a.r_up[pidx(P, li, -1) * 3 + l] = rr;
          }
        }
      }
      cx += dr; cy += dq;
      if (cx >= bw) { cx -= bw; cy++; }
    }
    allsum<2>(c, s, s_red, s_tot);
    const double sumb = s[0];
    double sumr = s[1];
    const double eps = sqrt(sumb) * 1e-6;                                                        // field.f90:364
    int act = 0, ite = 0;
    if (sqrt(sumr) > eps) act = sumb > eps;  // the first test compares sum(b^2), not its sqrt      field.f90:385-387
    // ---- p <- r                                                                                  field.f90:379
#pragma unroll
    for (int k = 0, cx = cx0, cy = cy0; k < CGP_K; k++) {
      const int e = t + k * CGP_T;
      if (e < ncl) T[e + 2 * cy + tp + 1] = r[k];
      cx += dr; cy += dq;
      if (cx >= bw) { cx -= bw; cy++; }
    }
    __syncthreads();
    for (int h = t; h < nhalo; h += CGP_T) {
      int ti, li, lj, kind;
      halo_pos(h, ti, li, lj, kind);
      double v;
      if (kind == 0) v = __ldcg(&a.rg[pidx(P, li, lj) * 3 + l]);
      else if (kind == 1) v = (l == 0) ? -T[ti + 1] : T[ti + 2];
      else v = 0.0;
      T[ti] = v;
    }
    __syncthreads();

#pragma unroll 1
    while (act) {
      // ---- phase A: ap <- f4 p - N4 p, sums r.r and p.ap                                        field.f90:392-413
      double sa[2] = {0.0, 0.0};
#pragma unroll
      for (int k = 0, cx = cx0, cy = cy0; k < CGP_K; k++) {
        const int e = t + k * CGP_T;
        if (e < ncl) {
          const int i = e + 2 * cy + tp + 1;
          const double pc = T[i];
          const double av = -T[i - tp] - T[i - 1] + f4 * pc - T[i + 1] - T[i + tp];
          const double rr = r[k];
          sa[0] = sa[0] + rr * rr;
          sa[1] = sa[1] + pc * av;
        }
        cx += dr; cy += dq;
        if (cx >= bw) { cx -= bw; cy++; }
      }
      allsum<2>(c, sa, s_red, s_tot);
      sumr = sa[0];
      const double alpha = sumr / sa[1];                                                          // field.f90:415
      // ---- phase B: phi += alpha p, r -= alpha ap, sum of the new r.r                            field.f90:417-441
      double sb1[1] = {0.0};
#pragma unroll
      for (int k = 0, cx = cx0, cy = cy0; k < CGP_K; k++) {
        const int e = t + k * CGP_T;
        if (e < ncl) {
          const int i = e + 2 * cy + tp + 1;
          const double pc = T[i];
          const double av = -T[i - tp] - T[i - 1] + f4 * pc - T[i + 1] - T[i + tp];
          phi[e] = phi[e] + alpha * pc;
          const double rr = r[k] - alpha * av;
          r[k] = rr;
          sb1[0] = sb1[0] + rr * rr;
          if (cx == 0 || cx == bw - 1 || cy == 0 || cy == bh - 1) {
            const int li = x0 + cx, lj = y0 + cy;
            a.rg[pidx(P, li, lj) * 3 + l] = rr;
            if (P.nsize > 1) {
              if (lj == 0) a.r_down[pidx(P, li, a.nyl_down) * 3 + l] = rr;
              if (lj == P.nyl - 1) a.r_up[pidx(P, li, -1) * 3 + l] = rr;
            }
          }
        }
        cx += dr; cy += dq;
        if (cx >= bw) { cx -= bw; cy++; }
      }
      allsum<1>(c, sb1, s_red, s_tot);
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
#pragma unroll
      for (int k = 0, cx = cx0, cy = cy0; k < CGP_K; k++) {
        const int e = t + k * CGP_T;
        if (e < ncl) {
          const int i = e + 2 * cy + tp + 1;
          T[i] = r[k] + beta * T[i];
        }
        cx += dr; cy += dq;
        if (cx >= bw) { cx -= bw; cy++; }
      }
      __syncthreads();
      for (int h = t; h < nhalo; h += CGP_T) {
        int ti, li, lj, kind;
        halo_pos(h, ti, li, lj, kind);
        double v;
        if (kind == 0) v = __ldcg(&a.rg[pidx(P, li, lj) * 3 + l]) + beta * T[ti];
        else if (kind == 1) v = (l == 0) ? -T[ti + 1] : T[ti + 2];
        else v = 0.0;
        T[ti] = v;
      }
      __syncthreads();
    }
    if (c.cta == 0 && t == 0) a.out[l] = ite;
  }

  // ---- df(l) <- phi on the interior                                                              field.f90:455-457
#pragma unroll
  for (int k = 0, cx = cx0, cy = cy0; k < CGP_K; k++) {
    const int e = t + k * CGP_T;
    if (e < ncl) {
      const size_t o = pidx(P, x0 + cx, y0 + cy);
#pragma unroll
      for (int l = 0; l < 3; l++) a.df[o * 6 + l] = a.phipl[l * plane + pbase + e];
    }
    cx += dr; cy += dq;
    if (cx >= bw) { cx -= bw; cy++; }
  }
  if (c.cta == 0 && t == 0) {
    a.out[3] = stop;
    a.out[4] = (int)c.nsync;
  }
}

// Block decomposition of an nx x nyl slab over at most nsm CTAs: cbx x cby blocks with at most CGP_K * CGP_T cells each whose
// tile (+ halo ring) fits `smem_max` bytes of shared memory.  Returns false if the slab is too large for the on-chip solver.
bool cgp_plan(int nx, int nyl, int nsm, size_t smem_max, int *cbx_out, int *cby_out, size_t *smem_out) {
  long long best = -1;
  int bcx = 0, bcy = 0;
  size_t bsm = 0;
  for (int cbx = 1; cbx <= nsm && cbx * 4 <= nx; cbx++)
    for (int cby = 1; cbx * cby <= nsm && cby <= nyl; cby++) {
      const int bw = (nx + cbx - 1) / cbx, bh = (nyl + cby - 1) / cby;
      const long long cells = (long long)bw * bh;
      if (cells > (long long)CGP_K * CGP_T) continue;
      const size_t sm = (size_t)(bw + 2) * (bh + 2) * sizeof(double);
      if (sm > smem_max) continue;
      // fewest cells per CTA first; below one cell per thread more CTAs only make the barrier slower: then the fewest CTAs;
      // ties: the shorter perimeter
      const long long work = cells < CGP_T ? CGP_T : cells;
      const long long score = work * 1000000LL + (cells < CGP_T ? (long long)cbx * cby * 1000 : 0) + (bw + bh);
      if (best < 0 || score < best) {
        best = score;
        bcx = cbx;
        bcy = cby;
        bsm = sm;
      }
    }
  if (best < 0) return false;
  *cbx_out = bcx;
  *cby_out = bcy;
  *smem_out = bsm;
  return true;
}

cudaError_t cgp_prepare(size_t smem) {
  return cudaFuncSetAttribute(k_cg_persist, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
}

cudaError_t launch_cg_persist(const DevParams &P, const CgpArgs &a, size_t smem, cudaStream_t st) {
  DevParams Pc = P;
  CgpArgs ac = a;
  void *args[] = {(void *)&Pc, (void *)&ac};
  return cudaLaunchCooperativeKernel((const void *)k_cg_persist, dim3(a.cbx * a.cby), dim3(CGP_T), args, smem, st);
}

}  // namespace wm
