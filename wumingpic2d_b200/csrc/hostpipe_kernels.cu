// hostpipe_kernels.cu -- layout kernels of the pipelined host step (wm_host_step with the upload of the next rows, the
// particle pass and the download of the finished rows running side by side, wm_api.cu "pipelined host step").
//
// They work on the reference's own index arrays, on the device: cumcnt(nxgs:nxge+1, nys:nye, nsp) as the caller holds it
// (common/sort.f90:36-82 produces it, common/particle.f90:83-177 consumes it) instead of the host-made tight offsets of
// wm_upload_particles_sorted, and on ROW RANGES of the slab, so that a chunk of rows can be laid out, pushed and sent back
// while the other rows are still on the PCIe bus.
#include "kernels.h"

namespace wm {

// cumcnt -> per-cell counts (the input of the segment layout); checks cumcnt(nxgs) = 0, counts >= 0, cumcnt(nxge+1) = np2
__global__ void k_cum_to_counts(const DevParams P, const int *__restrict__ cum, const int *__restrict__ np2, int *__restrict__ cnt,
                                unsigned *err) {
  const long long n = (long long)P.nsp * P.ncell;
  for (long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (long long)gridDim.x * blockDim.x) {
    const int isp = (int)(e / P.ncell), cell = (int)(e - (long long)isp * P.ncell);
    const int lj = cell / P.nx, li = cell - lj * P.nx;
    const int *row = cum + ((size_t)isp * P.nyl + lj) * (P.nx + 1);
    const int v = row[li + 1] - row[li];
    if (v < 0 || (li == 0 && row[0] != 0) || (li == P.nx - 1 && row[P.nx] != np2[lj + P.nyl * isp])) atomicOr(err, ERR_BAD_CELL);
    cnt[e] = v < 0 ? 0 : v;
  }
}
void launch_cum_to_counts(const DevParams &P, const int *cum, const int *np2, int *cnt, unsigned *err, cudaStream_t st) {
  k_cum_to_counts<<<148 * 8, 256, 0, st>>>(P, cum, np2, cnt, err);
}

// the slots behind the live particles of the segments of rows [ra, rb) become dead (k_mark_gaps for a row range: the chunk's
// share of the layout set-up, done while the chunk is on the bus)
__global__ void k_mark_gaps_rows(const DevParams P, PView<double> x, const int *__restrict__ cstart, const int *__restrict__ cnt, int ra, int rb) {
  const int ncl = (rb - ra) * P.nx;
  const long long n = (long long)P.nsp * ncl;
  const int lane = threadIdx.x & 31;
  for (long long wk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wk < n; wk += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int isp = (int)(wk / ncl), cell = ra * P.nx + (int)(wk - (long long)isp * ncl);
    const int *cs = cstart + (size_t)isp * (P.ncell + 1);
    const size_t so = (size_t)isp * P.cap;
    for (int p = cs[cell] + cnt[(size_t)isp * P.ncell + cell] + lane; p < cs[cell + 1]; p += 32) x[so + p] = dead_x();
  }
}
void launch_mark_gaps_rows(const DevParams &P, PView<double> x, const int *cstart, const int *cnt, int ra, int rb, cudaStream_t st) {
  if (rb > ra) k_mark_gaps_rows<<<148 * 4, 256, 0, st>>>(P, x, cstart, cnt, ra, rb);
}

// rows [ra, rb) of one species, as they lie in the host array (row after row, each sorted by cell: tight AoS records), ->
// their cell segments.  rowoff[lj - ra] - base = first record of row lj in `rec` (rowoff counts from the species' first row).
__global__ void k_rows_from_aos(const DevParams P, const double *__restrict__ rec, int n, int isp, int ra, int rb,
                                const int *__restrict__ rowoff, int base, const int *__restrict__ cum, const int *__restrict__ cstart,
                                const PartSoA dst, unsigned *err) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const double2 *r = reinterpret_cast<const double2 *>(rec + (size_t)s * 6);
  const double2 w0 = r[0], w1 = r[1], w2 = r[2];
  const int li = __double2int_rz(w0.x) - P.nxgs, lj = __double2int_rz(w0.y) - P.nys;
  if (li < 0 || li >= P.nx || lj < ra || lj >= rb) {
    atomicOr(err, ERR_BAD_CELL);
    return;
  }
  const int *row = cum + ((size_t)isp * P.nyl + lj) * (P.nx + 1);
  const int k = s + base - rowoff[lj - ra] - row[li];
  if (k < 0 || k >= row[li + 1] - row[li]) {  // cumcnt inconsistent with the positions
    atomicOr(err, ERR_BAD_CELL);
    return;
  }
  const int *cs = cstart + (size_t)isp * (P.ncell + 1);
  const int cell = lj * P.nx + li;
  if ((long long)cs[cell] + k >= P.cap || cs[cell] + k >= cs[cell + 1]) {
    atomicOr(err, ERR_CAPACITY);
    return;
  }
  double2 *o = dst.word((size_t)isp * P.cap + (size_t)cs[cell] + k);
  o[0] = w0;
  o[8] = w1;
  o[16] = w2;
}
void launch_rows_from_aos(const DevParams &P, const double *rec, int n, int isp, int ra, int rb, const int *rowoff, int base, const int *cum,
                          const int *cstart, const PartSoA &dst, unsigned *err, cudaStream_t st) {
  if (n > 0) k_rows_from_aos<<<(n + 255) / 256, 256, 0, st>>>(P, rec, n, isp, ra, rb, rowoff, base, cum, cstart, dst, err);
}

// final counts of rows [ra, rb) -> the reference's cumcnt rows (exclusive scan inside the row; entry nx = np2 of the row)
// and the first-record offsets of the rows in the download staging (species after species, row after row):
// rowoff[isp * (rb - ra) + (lj - ra)], rowoff[nsp * (rb - ra)] = records of the chunk.   One block; rows are scanned by warps.
__global__ void __launch_bounds__(1024) k_rows_cumcnt(const DevParams P, const int *__restrict__ cnt, int ra, int rb, int *__restrict__ cum,
                                                      int *__restrict__ rowoff) {
  __shared__ int s_tot[1024];
  const int nr = rb - ra, nrows = P.nsp * nr;
  const int wid = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int rr = wid; rr < nrows; rr += 32) {
    const int isp = rr / nr, lj = ra + (rr - isp * nr);
    const int *in = cnt + (size_t)isp * P.ncell + (size_t)lj * P.nx;
    int *out = cum + ((size_t)isp * P.nyl + lj) * (P.nx + 1);
    int run = 0;
    for (int b = 0; b < P.nx; b += 32) {
      const int v = (b + lane < P.nx) ? in[b + lane] : 0;
      int inc = v;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, inc, d);
        if (lane >= d) inc += t;
      }
      if (b + lane < P.nx) out[b + lane] = run + inc - v;
      run += __shfl_sync(0xffffffffu, inc, 31);
    }
    if (lane == 0) {
      out[P.nx] = run;
      s_tot[rr] = run;
    }
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    int run = 0;
    for (int rr = 0; rr < nrows; rr++) {
      rowoff[rr] = run;
      run += s_tot[rr];
    }
    rowoff[nrows] = run;
  }
}
void launch_rows_cumcnt(const DevParams &P, const int *cnt, int ra, int rb, int *cum, int *rowoff, cudaStream_t st) {
  k_rows_cumcnt<<<1, 1024, 0, st>>>(P, cnt, ra, rb, cum, rowoff);
}

// cell segments of rows [ra, rb) -> tight AoS records in the order of the host array (species, row, cell, slot order)
__global__ void k_rows_to_aos(const DevParams P, const PartSoA src, int ra, int rb, const int *__restrict__ cstart,
                              const int *__restrict__ cnt, const int *__restrict__ cum, const int *__restrict__ rowoff,
                              double *__restrict__ rec, int reccap, unsigned *err) {
  const int nr = rb - ra;
  const int ncl = nr * P.nx;
  const int lane = threadIdx.x & 31;
  const long long nwork = (long long)P.nsp * ncl;
  for (long long wk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wk < nwork; wk += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int isp = (int)(wk / ncl), cl = (int)(wk - (long long)isp * ncl);
    const int lj = ra + cl / P.nx, li = cl % P.nx;
    const int cell = lj * P.nx + li;
    const int s0 = cstart[(size_t)isp * (P.ncell + 1) + cell];
    const int n = min(cnt[(size_t)isp * P.ncell + cell], cstart[(size_t)isp * (P.ncell + 1) + cell + 1] - s0);  // (overflowed segment: rebuilt and fetched again)
    const int d0 = rowoff[isp * nr + (lj - ra)] + cum[((size_t)isp * P.nyl + lj) * (P.nx + 1) + li];
    if (d0 + n > reccap) continue;  // (the host sees the chunk's total and fetches everything again at the end)
    for (int k = lane; k < n; k += 32) {
      const double2 *w = src.word((size_t)isp * P.cap + (size_t)s0 + k);
      double2 *o = reinterpret_cast<double2 *>(rec + (size_t)(d0 + k) * 6);
      o[0] = w[0];
      o[1] = w[8];
      o[2] = w[16];
    }
  }
}
void launch_rows_to_aos(const DevParams &P, const PartSoA &src, int ra, int rb, const int *cstart, const int *cnt, const int *cum,
                        const int *rowoff, double *rec, int reccap, unsigned *err, cudaStream_t st) {
  k_rows_to_aos<<<148 * 8, 256, 0, st>>>(P, src, ra, rb, cstart, cnt, cum, rowoff, rec, reccap, err);
}

}  // namespace wm
