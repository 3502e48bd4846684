// wm_internal.h -- device data model shared by the kernels and the C-ABI host layer.
//
// Data layout in HBM (one context = one y-slab, rows nys..nye, full x extent):
//  * particles: 48-byte records (x, y, ux, uy, uz, id) blocked by 8 slots (see PView), species s at slot offset s*cap,
//    globally cell-sorted: cell = (j-nys)*nx + (i-nxgs); cell c of species s owns slots
//    [cstart[s][c], cstart[s][c] + cnt[s][c]); cstart[s][c+1] - cstart[s][c] is the segment's
//    CAPACITY (count + slack, so that a step only has to move the ~15 % of particles that change
//    cell; the layout is rebuilt when a segment overflows).  A slot is live iff x >= 0: gaps and
//    holes hold an all-ones NaN.  The reference's cumcnt(i,j,s) is the exclusive prefix sum of
//    cnt over the row and np2(j,s) the row total; both are produced exactly at download.
//  * grid arrays: the reference's own AoS layout incl. 2 ghost cells per side,
//    uf/df/tmpf (6, nx+4, nyl+4), uj/gkl/CG vectors (3, nx+4, nyl+4).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/wumingpic2d.h"

namespace wm {

constexpr int TX = 16;           // tile: cells in x
constexpr int TY = 8;            // tile: rows
constexpr int WINX = TX + 2;     // destination window (particles move < 1 cell per step)
constexpr int WINY = TY + 2;
constexpr int WIN = WINX * WINY; // 180 < 255: fits the 8-bit window field of a tag
constexpr int JX = TX + 4;       // current tile incl. the +-2 stencil halo
constexpr int JY = TY + 4;
constexpr int P1_THREADS = 256;
constexpr int GRP = 8;           // threads that share one cell's accumulators
// sort tag of a particle: kind<<31 | window cell<<23 | rank.  kind 0 = stays in its cell (rank
// counts the stayers of that cell in slot order), kind 1 = arrives in window cell w (rank from a
// shared-memory counter).  The scatter pass adds tilebase[kind][w] + cstart_new[cell(w)].
constexpr uint32_t TAG_DEAD = 0xFFFFFFFFu;   // left the slab (migrates to a neighbour rank)
constexpr int TAG_WSHIFT = 23;
constexpr uint32_t TAG_ARRIVAL = 0x80000000u;
constexpr uint32_t TAG_RANK_MASK = (1u << TAG_WSHIFT) - 1u;

// Particle store: the reference's 48-byte record (x, y, ux, uy, uz, id) = up(1:6,ii,j,isp) (proj/weibel/app.f90:75-76),
// kept as three 16-byte words w0 = (x, y), w1 = (ux, uy), w2 = (uz, id) and blocked by 8 slots:
//     block b = slot >> 3 (384 bytes) = [w0 of its 8 slots | w1 of its 8 slots | w2 of its 8 slots]   (3 x 128 bytes)
// The 8 lanes that work on a cell read 8 consecutive slots with three 128-bit loads, each a full 128-byte line
// at an immediate offset from one running pointer; consecutive destination slots of the sort are consecutive
// 16-byte pieces of a line.  PView gives the other kernels component-wise access (element i of one component).
constexpr int PBLK = 8;  // slots per block
__host__ __device__ __forceinline__ size_t pslot_w(size_t i) { return (i >> 3) * 24 + (i & 7); }  // in 16-byte words, word 0
template <typename T>
struct PView {
  T *p;  // base + offset of the component inside block 0, slot 0
  __host__ __device__ __forceinline__ T &operator[](size_t i) const { return p[2 * pslot_w(i)]; }
};
struct PartSoA {  // (the name predates the blocked record layout)
  PView<double> x, y, ux, uy, uz;
  PView<long long> id;
  // word w (0..2) of slot i as a 16-byte pointer: word(i)[8 * w]
  __host__ __device__ __forceinline__ double2 *word(size_t i) const { return reinterpret_cast<double2 *>(x.p) + pslot_w(i); }
};

struct DevParams {
  int nx, nyl;      // local cells
  int nxgs, nys;    // global index of local cell (0,0)
  int nygs, ny;     // global y range
  int nsize;        // ranks on the ring
  int bc;           // WM_BC_*: x boundary kind (y is periodic over the rank ring in every kind)
  int pitch;        // nx + 4
  int ntx, nty;     // tiles
  int nsp;
  int ncell;        // nx * nyl
  long long cap;    // per-species slot capacity
  double delx, delt, c, cc, inv_cc;
  double xlen, ylen;            // nx*delx, ny*delx
  double xwlo, xwhi;            // reflecting walls at (nxs+1)*delx, (nxe-1)*delx (wall kinds)
  double xw2lo, xw2hi;          // 2.*(nxs+1)*delx, 2.*(nxe-1)*delx as the reference computes them
  double u0x2;                  // WM_BC_SHOCK: 2.*u0 of bc__injection; then xwhi = xend, xw2hi = 2.*xend
                                // (proj/shock/boundary_shock.f90:272,287-290)
  double q[WM_NSP_MAX], r[WM_NSP_MAX];
  double f1, f2, f3, f4, f5, gfac, pi4dt;  // field.f90:53-57, 4*pi*delt
};

// error bits written by kernels into the context's device flag word
enum : unsigned {
  ERR_MOVED_TOO_FAR = 1u,   // a particle left the +-1 cell window (CFL violated / NaN)
  ERR_CAPACITY = 2u,        // particle slots exhausted ("memory over", boundary_periodic.f90:231-234)
  ERR_SENDBUF = 4u,         // migration buffer exhausted
  ERR_BAD_CELL = 8u,        // uploaded particle outside the slab
  ERR_TAG_RANK = 16u,       // more than 2^23 particles from one tile into one cell
  ERR_OVERFLOW = 32u,       // in-place sort: overflow list exhausted
  ERR_CG_TIMEOUT = 64u      // persistent CG kernel: a CTA or a ring neighbour never arrived at a barrier
};

// capacity of a cell segment that holds n particles now: slack ~ sl standard deviations of the
// change of a Poisson count between two layout rebuilds; sl = 0 -> tight
__host__ __device__ inline int cell_capacity(int n, float sl) {
  if (sl <= 0.f) return n;
  return (n + 4 + (int)ceilf(sl * sqrtf(2.0f * (float)n)) + 7) & ~7;  // whole blocks: a cell's 8 lanes read one 128-byte line per word
}

struct Pass1Args {
  PartSoA src, dst;         // dst == src for the in-place fused pass
  const int *cstart;        // [nsp][ncell+1] segment offsets (capacity)
  const int *cnt;           // [nsp][ncell]   live particles per segment (k_fused_dp: at the front of the segment)
  const int *cntb;          // [nsp][ncell]   k_fused_dp: particles at the back of the segment (arrivals of the last step)
  int *cnt_tail;            // [nsp][ncell]   in-place sort: append cursors (start = cnt); k_fused_dp: front counts of the new store
  int *cntb_new;            // [nsp][ncell]   k_fused_dp: back counts of the new store
  double *ovf;              // in-place sort: records that did not fit their segment [ovfcap][6], isp in ovfsp
  int *ovfsp;
  int *ovfcnt;
  int ovfcap;
  const double *tmpf;       // cell-centred fields, AoS6 padded
  double *uj;               // AoS3 padded, accumulated with RED.ADD.F64
  int *gcnt;                // [nsp][ncell] destination-cell counters / cursors
  int *tilebase;            // [ntiles][nsp][2][WIN]  (kind 0 = stayers, 1 = arrivals)
  uint32_t *tag;            // [nsp*cap]
  double *send[2];          // leavers: [dir][isp] AoS records, sendcap each   (dir 0 = down)
  int *sendcnt;             // [2][nsp]
  int sendcap;
  unsigned *err;
  double delt_push;         // delt (push) or delt/2 (mom_calc__accl)
  int tile0;                // k_fused_sm: first tile of this launch (row-range launches of the pipelined host step; 0 = whole slab)
};

#ifdef __CUDACC__
// a slot holds a particle iff x >= 0 (positions are >= nxgs >= 1); dead = all-ones NaN = memset 0xFF
__device__ __forceinline__ bool slot_live(double x) { return x >= 0.0; }
__device__ __forceinline__ double dead_x() { return __longlong_as_double(-1LL); }
// Staging of the cell changers (in-place sort): the idle particle store is reused as an array of 48-byte records
// (x y | ux uy | uz id, three 16-byte words) and the tag array (4 bytes per slot) holds their sort tags; the region of
// a quad is its own slot range [s0, s1): record r of the region is record s0 + r of the idle store, tag[s0 + r] its tag.
// Regions of different quads are disjoint.
__host__ __device__ inline long long so_slots(const DevParams &P, int isp) { return (long long)isp * P.cap; }
__host__ __device__ inline void stage_region(long long s0, long long s1, long long *rec0, int *cap) {
  *rec0 = s0;
  *cap = (int)(s1 > s0 ? s1 - s0 : 0);
}
// window index -> local cell index with the periodic wraps; -1 = outside the slab
__device__ __forceinline__ int window_cell(const DevParams &P, int li0, int lj0, int w) {
  int lx = w % WINX, ly = w / WINX;
  int li = li0 - 1 + lx;
  if (P.bc == WM_BC_PERIODIC) {
    if (li < 0) li += P.nx;
    if (li >= P.nx) li -= P.nx;
  }
  int lj = lj0 - 1 + ly;
  if (P.nsize == 1) {
    if (lj < 0) lj += P.nyl;
    if (lj >= P.nyl) lj -= P.nyl;
  }
  if (li < 0 || li >= P.nx || lj < 0 || lj >= P.nyl) return -1;
  return lj * P.nx + li;
}
#endif

}  // namespace wm
