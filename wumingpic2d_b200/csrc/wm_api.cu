// wm_api.cu -- host side of the C ABI (include/wumingpic2d.h): context, residency,
// per-procedure entry points, the fused step, NCCL exchanges on the y ring.
//
// The call order of one step follows proj/weibel/app.f90:100-107; what each entry point
// replaces is cited in the header.  No CPU fallback: everything fails without a device.
#include <cuda_runtime.h>
#include <nccl.h>

#include <algorithm>
#include <cmath>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "kernels.h"

using namespace wm;

namespace {

thread_local std::string g_err;

int fail(const char *fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return 1;
}

#define CU(x)                                                                                  \
  do {                                                                                         \
    cudaError_t e_ = (x);                                                                      \
    if (e_ != cudaSuccess) return fail("CUDA error %s at %s:%d (%s)", cudaGetErrorString(e_), __FILE__, __LINE__, #x); \
  } while (0)
// (a failure inside an open ncclGroupStart() must close the group before returning: ncclGroupEnd() is harmless otherwise)
#define NC(x)                                                                                  \
  do {                                                                                         \
    ncclResult_t e_ = (x);                                                                     \
    if (e_ != ncclSuccess) {                                                                   \
      (void)ncclGroupEnd();                                                                    \
      return fail("NCCL error %s at %s:%d (%s)", ncclGetErrorString(e_), __FILE__, __LINE__, #x); \
    }                                                                                          \
  } while (0)
#define WM(x)               \
  do {                      \
    int e_ = (x);           \
    if (e_) return e_;      \
  } while (0)

enum State { ST_EMPTY = 0, ST_SORTED, ST_PUSHED, ST_BOUNDED };

}  // namespace

struct wm_ctx {
  wm_config cfg;
  DevParams P;
  int dev = 0;
  cudaStream_t st = nullptr, st2 = nullptr;
  // particles: two stores of blocked 48-byte records (wm_internal.h, PView); `cur` holds the sorted state (the reference's `up`),
  // the other one is `gp` in stage mode and the scatter target in the fused step
  double *pbuf[2] = {nullptr, nullptr};
  PartSoA soa[2];
  int *cstart[2] = {nullptr, nullptr};   // [nsp][ncell+1] segment offsets of store b
  int *cnt[2] = {nullptr, nullptr};      // [nsp][ncell]   live particles per segment of store b
  int *cnt_tail = nullptr;               // [nsp][ncell]   append cursors of the in-place sort
  // k_fused_dp ("ping-pong" stores, direct placement): a segment holds cnt particles at its front and cntb at its back; the
  // pass reads store `cur` and writes the other one; has_back = the current state has back ranges (k_normalize removes them)
  int *cntb[2] = {nullptr, nullptr}, *cntb_tail = nullptr;
  bool has_back = false;
  int *tight = nullptr;                  // [nsp][ncell+1] scratch: exclusive scan of cnt (= cumcnt + row bases)
  double *ovf = nullptr;                 // in-place sort overflow list
  int *ovfsp = nullptr, *ovfcnt = nullptr, *ovfrank = nullptr, *h_ovf = nullptr;
  int ovfcap = 0;
  float slack = 6.0f;                    // segment slack in std deviations of the count change (WM_SLACK)
  int nbr_r = 0;                         // segments sized for the densest cell within nbr_r cells in x (adaptive: frequent rebuilds)
  long long nstep = 0, last_rebuild_step = -1000000;
  int cell_floor = 0;                    // every segment is laid out for at least this many particles (shock: upstream cells fill up)
  GenParams gen{};                       // parameters of the device-side particle sources (wm_ic_shock -> wm_shock_inject / _relocate)
  int *d_rowoff = nullptr;               // [nyl + 1] per-row record offsets of a generator call
  double *gen_stage = nullptr;           // records of one inject / relocate call, both species
  long long gen_stage_cap = 0;
  bool inplace = true;                   // WM_INPLACE=0 selects the tag + scatter sort for wm_step
  bool cg3 = false;                      // WM_CG3=1: three-kernel CG iteration (k_cg_ap, k_cg_update, k_cg_pupdate)
  bool rimplace = true;                  // k_place_rim after k_fused_sm<TAIL> (WM_RIMPLACE=0: the general k_place)
  int sm = 1;                            // 1: k_fused_sm<TAIL> + k_place_rim (default); 5: k_fused_dp (ping-pong stores, direct placement:
                                         // measured slower, kept as a tested variant); 0: k_fused<INPLACE> (WM_SM)
  long long rebuilds = 0;
  int cur = 0;
  int *gcnt = nullptr, *tilebase = nullptr, *scan_scratch = nullptr;
  uint32_t *tag = nullptr;
  State state = ST_EMPTY;
  // migration on the y ring
  int sendcap = 0;
  double *send[2] = {nullptr, nullptr}, *recv[2] = {nullptr, nullptr};
  int *sendcnt = nullptr, *recvcnt = nullptr;  // device [2][nsp]
  int *in_rank = nullptr;                      // [2*nsp*sendcap]
  int *h_cnt = nullptr;                        // pinned [4*nsp]
  int n_in[2][WM_NSP_MAX] = {{0}};
  bool mig_prev = false;                       // h_cnt holds the counts of the previous step's exchange
  int mig_fast = 1;                            // WM_MIGSYNC=1 keeps the count-then-payload protocol with its host round trip
  // fields
  FieldBufs f{};
  double *rowtmp = nullptr;  // 2 ghost rows x 6 comps (multi-rank fold)
  double *mom = nullptr;
  double *partial = nullptr, *h_partial = nullptr;
  unsigned *d_err = nullptr, *h_err = nullptr;
  int *h_cg = nullptr;       // pinned copies of CgCtl {active[3], ite[3], stop} x 2
  cudaEvent_t ev_cg[2] = {nullptr, nullptr};
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr, ev_b[2] = {nullptr, nullptr};  // sort on st2 beside the field solve
  bool overlap = true;                   // WM_OVERLAP=0: everything on one stream
  int nxa = 0;                           // active cells in x: nxe - nxgs + 1 (< nx only while the shock box grows)
  double u_inject = 0.0;
  bool u_inject_set = false;             // WM_BC_SHOCK: wm_set_u_inject / wm_boundary__injection has been called
  int cg_ite[3] = {0, 0, 0};
  // persistent cooperative CG (k_cg_persist): one launch per solve, state on chip, barriers and sums in the kernel
  int cg_mode = 1;                       // WM_CG=0 selects the host loop of small kernels (k_cg_pap / k_cg_update2)
  bool cgp_ok = false;                   // the slab fits the on-chip solver on this device
  bool cgp_ring = false;                 // nsize > 1: the neighbours' arrays are mapped (CUDA IPC) on every rank
  int cgp_cbx = 0, cgp_cby = 0, cgp_rl = 0, nsm = 0;
  int cgp_planned_nxa = -1;
  bool cgp_plan_ok = false;
  size_t cgp_smem = 0, cgp_smem_max = 0;
  double *cgp_partial = nullptr;
  unsigned *cgp_bar = nullptr;           // {barrier counter, abort word}
  int *cgp_out = nullptr, *h_cgp_out = nullptr;
  bool cg_pending = false;               // h_cgp_out of the last solve has not been looked at yet
  unsigned long long cgp_seq = 0;
  CgpShared *cgp_sh[CGP_MAXR] = {nullptr}, *cgp_sh_mine = nullptr;
  void *cgp_ipc_open[CGP_MAXR + 2] = {nullptr};  // mappings to close at destroy
  int cgp_nopen = 0;
  double *cgp_r_up = nullptr, *cgp_r_down = nullptr;
  int cgp_nyl_down = 0, cgp_nyl_up = 0;
  // comm
  ncclComm_t comm = nullptr;
  int nup = 0, ndown = 0;
  // timing
  cudaEvent_t ev[6] = {nullptr};
  cudaEvent_t ev_call[2] = {nullptr, nullptr};
  double ms[5] = {0, 0, 0, 0, 0};
  long long launches = 0;
  bool timing = true;
  bool accl_valid = false;   // the idle store holds mom_calc__accl's half-step momenta
};

namespace {

void carve(PartSoA &s, double *base, long long, int) {
  // component offsets inside a block: w0 = (x, y) at 0, w1 = (ux, uy) at 8 words, w2 = (uz, id) at 16 words
  s.x.p = base;
  s.y.p = base + 1;
  s.ux.p = base + 16;
  s.uy.p = base + 17;
  s.uz.p = base + 32;
  s.id.p = reinterpret_cast<long long *>(base + 33);
}

int alloc_particles(wm_ctx *c, long long need) {
  if (c->pbuf[0]) {
    if (need <= c->P.cap) return 0;
    return fail("particle capacity %lld exceeded (need %lld); recreate the context with a larger wm_config.capacity", c->P.cap, need);
  }
  // slots per species: every cell segment carries slack (wm_internal.h cell_capacity); by
  // Cauchy-Schwarz sum_c cap(n_c) <= need + 12 ncell + slack sqrt(2 ncell need)
  auto bound = [&](float sl) { return (double)need + 12.0 * c->P.ncell + (double)sl * std::sqrt(2.0 * c->P.ncell * (double)need); };  // cell_capacity adds up to 4 + ceil + 7
  long long cap = c->cfg.capacity > 0 ? c->cfg.capacity : (long long)std::ceil(1.1 * bound(c->slack)) + 4096;
  if (cap < need) return fail("wm_config.capacity %lld < particles per species %lld", cap, need);
  while (c->slack > 0.f && bound(c->slack) > (double)cap) c->slack = (c->slack > 0.5f) ? c->slack * 0.5f : 0.f;
  if (c->slack <= 0.f) c->inplace = false;
  cap = (cap + 7) & ~7LL;  // species offsets are whole blocks of the particle store
  if (cap >= (1LL << 31) - 1) return fail("more than 2^31 particles per species per GPU are not supported");
  c->P.cap = cap;
  const int nsp = c->P.nsp;
  for (int b = 0; b < 2; b++) {
    CU(cudaMalloc(&c->pbuf[b], (size_t)cap * nsp * 6 * sizeof(double)));
    carve(c->soa[b], c->pbuf[b], cap, nsp);
    CU(cudaMalloc(&c->cstart[b], (size_t)nsp * (c->P.ncell + 1) * sizeof(int)));
    CU(cudaMalloc(&c->cnt[b], (size_t)nsp * c->P.ncell * sizeof(int)));
    CU(cudaMalloc(&c->cntb[b], (size_t)nsp * c->P.ncell * sizeof(int)));
    CU(cudaMemset(c->cntb[b], 0, (size_t)nsp * c->P.ncell * sizeof(int)));
    CU(cudaMemset(c->pbuf[b], 0xFF, (size_t)cap * nsp * 6 * sizeof(double)));  // every slot dead (x = all-ones NaN)
  }
  CU(cudaMalloc(&c->cnt_tail, (size_t)nsp * c->P.ncell * sizeof(int)));
  CU(cudaMalloc(&c->cntb_tail, (size_t)nsp * c->P.ncell * sizeof(int)));
  CU(cudaMalloc(&c->tight, (size_t)nsp * (c->P.ncell + 1) * sizeof(int)));
  c->ovfcap = (int)std::min<long long>(std::max<long long>(need / 64, 1 << 16), 1 << 24);
  CU(cudaMalloc(&c->ovf, (size_t)c->ovfcap * 6 * sizeof(double)));
  CU(cudaMalloc(&c->ovfsp, (size_t)c->ovfcap * sizeof(int)));
  CU(cudaMalloc(&c->ovfrank, (size_t)c->ovfcap * sizeof(int)));
  CU(cudaMalloc(&c->ovfcnt, sizeof(int)));
  CU(cudaMemset(c->ovfcnt, 0, sizeof(int)));
  CU(cudaMallocHost(&c->h_ovf, sizeof(int)));
  CU(cudaMalloc(&c->tag, (size_t)cap * nsp * sizeof(uint32_t)));
  return 0;
}

inline size_t host_up_index(const wm_config &g, int isp, int jl) {
  const int nyl = g.nye - g.nys + 1;
  return (size_t)6 * ((size_t)g.np * ((size_t)jl + (size_t)nyl * isp));
}

// Parameters of the grid kernels: the same layout (pitch), x extent = the active range nxs..nxe of the call
// (field.f90 loops i = nxs..nxe; only the shock app moves nxe, proj/shock/app.f90:611-621).
inline DevParams fieldp(const wm_ctx *c) {
  DevParams P = c->P;
  P.nx = c->nxa;
  return P;
}

// result of the last persistent CG solve (copied to pinned memory behind the kernel): call after a synchronisation
int cg_collect(wm_ctx *c) {
  if (!c->cg_pending) return 0;
  c->cg_pending = false;
  for (int l = 0; l < 3; l++) c->cg_ite[l] = c->h_cgp_out[l];
  if (c->h_cgp_out[3]) return fail("********** stop at cgm after ite_max ********** (field.f90:427-430)");
  return 0;
}

int check_errors(wm_ctx *c, const char *where) {
  CU(cudaMemcpyAsync(c->h_err, c->d_err, sizeof(unsigned), cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  const unsigned e = *c->h_err;
  if (!e) return cg_collect(c);
  c->cg_pending = false;
  CU(cudaMemsetAsync(c->d_err, 0, sizeof(unsigned), c->st));
  std::string m = std::string(where) + ":";
  if (e & ERR_MOVED_TOO_FAR) m += " particle moved more than one cell (CFL violated or NaN);";
  if (e & ERR_CAPACITY) m += " memory over (particle slots exhausted, cf. boundary_periodic.f90:231-234);";
  if (e & ERR_SENDBUF) m += " migration buffer exhausted;";
  if (e & ERR_BAD_CELL) m += " particle outside the slab;";
  if (e & ERR_TAG_RANK) m += " more than 2^23 particles from one tile into one cell;";
  if (e & ERR_OVERFLOW) m += " in-place sort overflow list exhausted (raise WM_SLACK or wm_config.capacity);";
  if (e & ERR_CG_TIMEOUT) m += " persistent CG kernel: a barrier timed out (a CTA or a ring neighbour never arrived);";
  return fail("%s", m.c_str());
}

// ---- y-ring exchange of `rows` contiguous grid rows (ncomp doubles per cell) -------------
// send rows starting at local row lj_send to `to`, receive into rows starting at lj_recv from `from`
int ring_rows(wm_ctx *c, double *a, int ncomp, int lj_send, int to, int lj_recv, int from, int rows, double *recv_override = nullptr) {
  const size_t w = (size_t)c->P.pitch * ncomp;
  double *src = a + (size_t)(lj_send + 2) * w;
  double *dst = recv_override ? recv_override : a + (size_t)(lj_recv + 2) * w;
  NC(ncclGroupStart());
  NC(ncclSend(src, rows * w, ncclDouble, to, c->comm, c->st));
  NC(ncclRecv(dst, rows * w, ncclDouble, from, c->comm, c->st));
  NC(ncclGroupEnd());
  return 0;
}

// bc__dfield / bc__phi: ghost rows then periodic x ghosts   boundary_periodic.f90:251-354,511-568
int halo_copy(wm_ctx *c, double *a, int ncomp, int ng, bool do_x) {
  const int nyl = c->P.nyl;
  if (c->P.nsize == 1) {
    launch_fill_y_local(fieldp(c), a, ncomp, ng, c->st);
    c->launches++;
  } else {
    // my first ng rows -> ndown, nup's first rows land in my upper ghosts; my last ng rows -> nup, ndown's last
    // rows land in my lower ghosts: both directions in one NCCL group (one launch instead of two)
    const size_t w = (size_t)c->P.pitch * ncomp;
    NC(ncclGroupStart());
    NC(ncclSend(a + (size_t)(0 + 2) * w, ng * w, ncclDouble, c->ndown, c->comm, c->st));
    NC(ncclRecv(a + (size_t)(nyl + 2) * w, ng * w, ncclDouble, c->nup, c->comm, c->st));
    NC(ncclSend(a + (size_t)(nyl - ng + 2) * w, ng * w, ncclDouble, c->nup, c->comm, c->st));
    NC(ncclRecv(a + (size_t)(-ng + 2) * w, ng * w, ncclDouble, c->ndown, c->comm, c->st));
    NC(ncclGroupEnd());
  }
  if (do_x) {
    launch_fill_x(fieldp(c), a, ncomp, ng, c->st);
    c->launches++;
  }
  return 0;
}

// bc__curre: fold ghost rows into the neighbour, refresh ghosts, x fold   boundary_periodic.f90:357-508
int bc_curre(wm_ctx *c) {
  const int nyl = c->P.nyl;
  double *uj = c->f.uj;
  if (c->P.nsize == 1) {
    launch_fold_y_local(c->P, uj, c->st);
    c->launches++;
  } else {
    const size_t w = (size_t)c->P.pitch * 3;
    // my lower ghosts (nys-2,nys-1) -> ndown ; nup's are added into my nye-1,nye
    WM(ring_rows(c, uj, 3, -2, c->ndown, 0, c->nup, 2, c->rowtmp));
    launch_add_rows(uj + (size_t)(nyl - 2 + 2) * w, c->rowtmp, 2 * w, c->st);
    // my upper ghosts (nye+1,nye+2) -> nup ; ndown's are added into my nys,nys+1
    WM(ring_rows(c, uj, 3, nyl, c->nup, 0, c->ndown, 2, c->rowtmp));
    launch_add_rows(uj + (size_t)(0 + 2) * w, c->rowtmp, 2 * w, c->st);
    c->launches += 2;
    // refresh: my nys,nys+1 -> ndown (their upper ghosts) ; my nye-1,nye -> nup (their lower ghosts)
    WM(ring_rows(c, uj, 3, 0, c->ndown, nyl, c->nup, 2));
    WM(ring_rows(c, uj, 3, nyl - 2, c->nup, -2, c->ndown, 2));
  }
  if (c->P.bc == WM_BC_PERIODIC) {  // the wall modules end after the y exchange (boundary_reconnection.f90:364-502)
    launch_fold_x(c->P, uj, c->st);
    c->launches++;
  }
  return 0;
}

int allreduce_ctl(wm_ctx *c, int which, int n) {
  if (c->P.nsize == 1) return 0;
  double *p = reinterpret_cast<double *>(reinterpret_cast<char *>(c->f.cgstate) + cgctl_sums_offset(which));
  NC(ncclAllReduce(p, p, n, ncclDouble, ncclSum, c->comm, c->st));
  return 0;
}

// all-reduce of the sums `which` and the one-row ghost exchange of the CG vector a (3 components) in ONE NCCL group:
// one launch instead of two per CG iteration
int allreduce_and_halo(wm_ctx *c, int which, int n, double *a) {
  if (c->P.nsize == 1) return 0;
  double *p = reinterpret_cast<double *>(reinterpret_cast<char *>(c->f.cgstate) + cgctl_sums_offset(which));
  const int nyl = c->P.nyl;
  const size_t w = (size_t)c->P.pitch * 3;
  NC(ncclGroupStart());
  NC(ncclAllReduce(p, p, n, ncclDouble, ncclSum, c->comm, c->st));
  NC(ncclSend(a + (size_t)(0 + 2) * w, w, ncclDouble, c->ndown, c->comm, c->st));
  NC(ncclRecv(a + (size_t)(nyl + 2) * w, w, ncclDouble, c->nup, c->comm, c->st));
  NC(ncclSend(a + (size_t)(nyl - 1 + 2) * w, w, ncclDouble, c->nup, c->comm, c->st));
  NC(ncclRecv(a + (size_t)(-1 + 2) * w, w, ncclDouble, c->ndown, c->comm, c->st));
  NC(ncclGroupEnd());
  return 0;
}

// cgm for l = 1..3 in one persistent cooperative kernel (cg_persist_kernel.cu)     field.f90:319-461
int cg_solve_persist(wm_ctx *c) {
  const DevParams P = fieldp(c);
  CgpArgs a{};
  a.cbx = c->cgp_cbx;
  a.cby = c->cgp_cby;
  a.rl = c->cgp_rl;
  a.cp = ((P.nyl + a.cby - 1) / a.cby + 2) | 1;
  a.df = c->f.df;
  a.gkl = c->f.gkl;
  a.rg = c->f.r;
  a.phipl = c->f.phi;
  a.bpl = c->f.ap;
  a.partial = c->cgp_partial;
  a.bar = c->cgp_bar;
  a.abort = reinterpret_cast<int *>(c->cgp_bar + 1);
  a.out = c->cgp_out;
  a.err = c->d_err;
  a.nrank = c->cfg.nrank;
  a.nsize = P.nsize;
  a.seq0 = c->cgp_seq;
  c->cgp_seq += 1024;  // more than the 3 x (1 + 2 x 100) barriers a solve can take
  for (int q = 0; q < CGP_MAXR; q++) a.sh[q] = c->cgp_sh[q];
  a.r_up = c->cgp_r_up;
  a.r_down = c->cgp_r_down;
  a.nyl_down = c->cgp_nyl_down;
  a.nup = c->nup;
  a.ndown = c->ndown;
  a.cbx_up = a.cbx_down = a.cbx;
  if (P.nsize > 1) {  // the neighbours' decompositions follow from their row counts (the same rule on every rank)
    int cby, rl;
    size_t sm;
    if (!cgp_plan(c->nxa, c->cgp_nyl_up, c->nsm, c->cgp_smem_max, &a.cbx_up, &cby, &rl, &sm) ||
        !cgp_plan(c->nxa, c->cgp_nyl_down, c->nsm, c->cgp_smem_max, &a.cbx_down, &cby, &rl, &sm))
      return fail("cg_solve_persist: a ring neighbour's slab does not fit the on-chip solver");
  }
  static const char *tr = getenv("WM_CGTRACE");
  unsigned long long *d_tr = nullptr;
  const int G = a.cbx * a.cby;
  if (tr && atoi(tr) > 0) {
    CU(cudaMalloc(&d_tr, (size_t)G * 16 * sizeof(unsigned long long)));
    CU(cudaMemset(d_tr, 0, (size_t)G * 16 * sizeof(unsigned long long)));
    a.trace = d_tr;
    a.trace_seq = (unsigned)atoi(tr);
  }
  CU(cudaMemsetAsync(c->cgp_bar, 0, 2 * sizeof(unsigned), c->st));
  CU(launch_cg_persist(P, a, c->cgp_smem, c->st));
  if (d_tr) {  // debugging aid: spread of the CTAs' arrival / release times at four consecutive barriers
    std::vector<unsigned long long> h((size_t)G * 16);
    CU(cudaStreamSynchronize(c->st));
    CU(cudaMemcpy(h.data(), d_tr, h.size() * sizeof(unsigned long long), cudaMemcpyDeviceToHost));
    CU(cudaFree(d_tr));
    unsigned long long t0 = ~0ull;
    for (auto v : h) if (v && v < t0) t0 = v;
    static int shown = 0;
    if (shown++ < 3)
      for (int b = 0; b < 4; b++) {
        static const char *nm[4] = {"reduced", "arrived", "between done", "released"};
        for (int e = 0; e < 4; e++) {
          unsigned long long lo = ~0ull, hi = 0;
          double mean = 0;
          for (int g = 0; g < G; g++) {
            const unsigned long long v = h[((size_t)g * 4 + b) * 4 + e];
            lo = std::min(lo, v);
            hi = std::max(hi, v);
            mean += (double)(v - t0);
          }
          fprintf(stderr, "cgtrace barrier %u %-13s min %8.2f us mean %8.2f us max %8.2f us\n", a.trace_seq + b, nm[e],
                  (lo - t0) * 1e-3, mean / G * 1e-3, (hi - t0) * 1e-3);
        }
      }
  }
  CU(cudaMemcpyAsync(c->h_cgp_out, c->cgp_out, 8 * sizeof(int), cudaMemcpyDeviceToHost, c->st));
  c->cg_pending = true;
  c->launches++;
  return 0;
}

// Block decomposition for the active x range of this call (only the shock module moves nxe).  With more than one rank the
// plan must also hold for a slab with one more row (mpi_set.f90:37-41 spreads the remainder), so that all ranks decide alike.
bool cg_persist_usable(wm_ctx *c) {
  if (!c->cg_mode || !c->cgp_ok) return false;
  if (c->P.nsize > 1 && !c->cgp_ring) return false;
  if (c->cgp_planned_nxa != c->nxa) {
    int cbx, cby, rl;
    size_t sm;
    const int nyl = c->P.nyl;
    bool ok = cgp_plan(c->nxa, nyl, c->nsm, c->cgp_smem_max, &cbx, &cby, &rl, &sm);
    if (ok && c->P.nsize > 1) {  // slabs hold ny / nsize rows, the first mod(ny, nsize) ranks one more
      const int ny = c->cfg.nyge - c->cfg.nygs + 1, base = ny / c->P.nsize;
      int bx2, by2, rl2;
      size_t sm2;
      ok = cgp_plan(c->nxa, base, c->nsm, c->cgp_smem_max, &bx2, &by2, &rl2, &sm2) &&
           cgp_plan(c->nxa, base + (ny % c->P.nsize ? 1 : 0), c->nsm, c->cgp_smem_max, &bx2, &by2, &rl2, &sm2);
    }
    c->cgp_planned_nxa = c->nxa;
    c->cgp_plan_ok = ok;
    if (ok) {
      c->cgp_cbx = cbx;
      c->cgp_cby = cby;
      c->cgp_rl = rl;
      c->cgp_smem = sm;
    }
  }
  return c->cgp_plan_ok;
}

// cgm for l = 1..3 together                                                field.f90:319-461
int cg_solve(wm_ctx *c) {
  if (cg_persist_usable(c)) return cg_solve_persist(c);
  const DevParams P = fieldp(c);
  launch_cg_init(P, c->f, c->st);
  WM(allreduce_ctl(c, 0, 3));
  if (P.nsize > 1) WM(halo_copy(c, c->f.phi, 3, 1, false));  // set_boundary_phi(phi)  field.f90:367
  launch_cg_resid0(P, c->f, c->st);
  WM(allreduce_ctl(c, 1, 3));
  launch_cg_begin(P, c->f, P.nsize, c->st);
  c->launches += 3;
  char *ctl = reinterpret_cast<char *>(c->f.cgstate) + cgctl_active_offset();
  const size_t ctl_n = 7 * sizeof(int);
  int it = 0;
  bool done = false;
  double *p_in = c->f.p, *p_out = c->f.p2;
  for (; it < 101 && !done; it++) {
    if (c->cg3) {
      // three-kernel iteration (WM_CG3=1): A p, update, p update, as field.f90:392-450 is written
      if (P.nsize > 1) WM(halo_copy(c, c->f.p, 3, 1, false));  // set_boundary_phi(p)  field.f90:392
      launch_cg_ap(P, c->f, c->st);
      WM(allreduce_ctl(c, 1, 6));
      launch_cg_update(P, c->f, c->st);
      WM(allreduce_ctl(c, 3, 3));
      launch_cg_pupdate(P, c->f, c->st);
      c->launches += 3;
    } else {
      // two-kernel iteration: the p update is folded into the next A p; the neighbours' rows of p follow from the
      // exchanged rows of r (one exchange per iteration, as set_boundary_phi(p) at field.f90:392)
      if (P.nsize > 1 && it == 0) WM(halo_copy(c, c->f.r, 3, 1, false));
      launch_cg_pap(P, c->f, p_in, p_out, c->st);
      WM(allreduce_ctl(c, 1, 6));
      launch_cg_update2(P, c->f, p_out, c->st);
      WM(allreduce_and_halo(c, 3, 3, c->f.r));  // sum r^2 and the rows of the new r the next A p needs
      std::swap(p_in, p_out);
      c->launches += 2;
    }
    int *h = c->h_cg + (it & 1) * 8;
    CU(cudaMemcpyAsync(h, ctl, ctl_n, cudaMemcpyDeviceToHost, c->st));
    CU(cudaEventRecord(c->ev_cg[it & 1], c->st));
    if (it >= 1) {
      // look at the previous iteration's flags while this one is in flight
      CU(cudaEventSynchronize(c->ev_cg[(it - 1) & 1]));
      const int *hp = c->h_cg + ((it - 1) & 1) * 8;
      if (!(hp[0] | hp[1] | hp[2])) done = true;
    }
  }
  CU(cudaEventSynchronize(c->ev_cg[(it - 1) & 1]));
  const int *hl = c->h_cg + ((it - 1) & 1) * 8;
  for (int l = 0; l < 3; l++) c->cg_ite[l] = hl[3 + l];
  if (hl[6]) return fail("********** stop at cgm after ite_max ********** (field.f90:427-430)");
  if (hl[0] | hl[1] | hl[2]) return fail("cgm: internal error, loop ended while a component is still active");
  launch_cg_finish(P, c->f, c->st);
  c->launches++;
  return 0;
}

// everything of field__fdtd_i after ele_cur, in two halves: up to the end of the CG solves (field.f90:122-149) ...
int field_solve_pre(wm_ctx *c) {
  const DevParams P = fieldp(c);
  WM(bc_curre(c));
  launch_rhs(P, c->f, c->st);
  c->launches++;
  WM(cg_solve(c));
  return 0;
}
// ... and from bc__dfield to the update of uf (field.f90:151-184)
int field_solve_post(wm_ctx *c) {
  const DevParams P = fieldp(c);
  WM(halo_copy(c, c->f.df, 6, 2, true));
  launch_efield(P, c->f, c->st);
  WM(halo_copy(c, c->f.df, 6, 2, true));
  launch_update_uf(P, c->f, c->st);
  c->launches += 2;
  return 0;
}
int field_solve(wm_ctx *c) {
  WM(field_solve_pre(c));
  return field_solve_post(c);
}

Pass1Args p1args(wm_ctx *c, const PartSoA &src, const PartSoA &dst, double delt_push) {
  Pass1Args a{};
  a.src = src;
  a.dst = dst;
  a.cstart = c->cstart[c->cur];
  a.cnt = c->cnt[c->cur];
  a.cntb = c->cntb[c->cur];
  a.cnt_tail = c->cnt_tail;
  a.cntb_new = c->cntb_tail;
  a.ovf = c->ovf;
  a.ovfsp = c->ovfsp;
  a.ovfcnt = c->ovfcnt;
  a.ovfcap = c->ovfcap;
  a.tmpf = c->f.tmpf;
  a.uj = c->f.uj;
  a.gcnt = c->gcnt;
  a.tilebase = c->tilebase;
  a.tag = c->tag;
  a.send[0] = c->send[0];
  a.send[1] = c->send[1];
  a.sendcnt = c->sendcnt;
  a.sendcap = c->sendcap;
  a.err = c->d_err;
  a.delt_push = delt_push;
  return a;
}

int zero_sort_state(wm_ctx *c) {
  CU(cudaMemsetAsync(c->gcnt, 0, (size_t)c->P.nsp * c->P.ncell * sizeof(int), c->st));
  if (c->P.nsize > 1) CU(cudaMemsetAsync(c->sendcnt, 0, 2 * WM_NSP_MAX * sizeof(int), c->st));
  return 0;
}

// ring exchange of the leavers packed by the BOUND pass, then rank the arrivals
// boundary_periodic.f90:173-189
int migrate(wm_ctx *c, bool inplace = false, bool dp = false) {
  const DevParams &P = c->P;
  // append cursors of the arrivals: the in-place sort's cnt_tail, or (k_fused_dp) the front counts of the new store, limited by
  // its back ranges
  int *const cursor = dp ? c->cnt[c->cur] : c->cnt_tail;
  const int *const backs = dp ? c->cntb[c->cur] : nullptr;
  if (P.nsize == 1) return 0;
  if (!c->comm) return fail("nsize > 1 but wm_comm_init has not been called");
  const int nsp = P.nsp;
  if (inplace && c->mig_fast && c->mig_prev && nsp <= 2) {
    // No host round trip: the counts travel WITH the payload in one NCCL group.  A message is sized from the count the same
    // channel carried in the previous step (both ends know it: the sender sent it, the receiver got it in the header), twice
    // that + 4096 records; the counts of this step are checked on the device (a truncated message is an error, not a loss)
    // and copied to the host behind the kernels for the next step's sizes (wm_step synchronises at the end of every step).
    int ms[4], mr[4];  // [dir * nsp + isp]: send down / up, receive from nup (their down-going) / ndown (their up-going)
    for (int k = 0; k < 2 * nsp; k++) {
      ms[k] = (int)std::min<long long>(c->sendcap, 2LL * c->h_cnt[k] + 4096);
      mr[k] = (int)std::min<long long>(c->sendcap, 2LL * c->h_cnt[2 * nsp + k] + 4096);
    }
    for (int k = 2 * nsp; k < 4; k++) ms[k] = mr[k] = 0;
    NC(ncclGroupStart());
    NC(ncclSend(c->sendcnt, nsp, ncclInt, c->ndown, c->comm, c->st));
    NC(ncclSend(c->sendcnt + nsp, nsp, ncclInt, c->nup, c->comm, c->st));
    NC(ncclRecv(c->recvcnt, nsp, ncclInt, c->nup, c->comm, c->st));
    NC(ncclRecv(c->recvcnt + nsp, nsp, ncclInt, c->ndown, c->comm, c->st));
    for (int isp = 0; isp < nsp; isp++) {
      const size_t off = (size_t)isp * c->sendcap * 6;
      NC(ncclSend(c->send[0] + off, (size_t)ms[isp] * 6, ncclDouble, c->ndown, c->comm, c->st));
      NC(ncclSend(c->send[1] + off, (size_t)ms[nsp + isp] * 6, ncclDouble, c->nup, c->comm, c->st));
      NC(ncclRecv(c->recv[0] + off, (size_t)mr[isp] * 6, ncclDouble, c->nup, c->comm, c->st));
      NC(ncclRecv(c->recv[1] + off, (size_t)mr[nsp + isp] * 6, ncclDouble, c->ndown, c->comm, c->st));
    }
    NC(ncclGroupEnd());
    launch_check_counts(c->sendcnt, 2 * nsp, ms, c->d_err, c->st);
    for (int d = 0; d < 2; d++)
      for (int isp = 0; isp < nsp; isp++) {
        const size_t off = (size_t)isp * c->sendcap;
        launch_incoming_append(P, c->recv[d] + off * 6, mr[d * nsp + isp], isp, c->cstart[c->cur], cursor, c->soa[c->cur], c->ovf,
                               c->ovfsp, c->ovfcnt, c->ovfcap, c->d_err, c->st, c->recvcnt + d * nsp + isp, backs);
        c->launches++;
      }
    c->launches++;
    CU(cudaMemcpyAsync(c->h_cnt, c->sendcnt, 2 * nsp * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CU(cudaMemcpyAsync(c->h_cnt + 2 * nsp, c->recvcnt, 2 * nsp * sizeof(int), cudaMemcpyDeviceToHost, c->st));
    return 0;
  }
  // counts first (MPI_SENDRECV of cnt, :174,:183)
  NC(ncclGroupStart());
  NC(ncclSend(c->sendcnt, nsp, ncclInt, c->ndown, c->comm, c->st));
  NC(ncclSend(c->sendcnt + nsp, nsp, ncclInt, c->nup, c->comm, c->st));
  NC(ncclRecv(c->recvcnt, nsp, ncclInt, c->nup, c->comm, c->st));          // nup's down-going
  NC(ncclRecv(c->recvcnt + nsp, nsp, ncclInt, c->ndown, c->comm, c->st));  // ndown's up-going
  NC(ncclGroupEnd());
  CU(cudaMemcpyAsync(c->h_cnt, c->sendcnt, 2 * nsp * sizeof(int), cudaMemcpyDeviceToHost, c->st));
  CU(cudaMemcpyAsync(c->h_cnt + 2 * nsp, c->recvcnt, 2 * nsp * sizeof(int), cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  for (int k = 0; k < 4 * nsp; k++)
    if (c->h_cnt[k] > c->sendcap) return fail("migration buffer exhausted (%d > %d records)", c->h_cnt[k], c->sendcap);
  // payload (MPI_SENDRECV of bff_ptcl, :177,:186)
  NC(ncclGroupStart());
  for (int isp = 0; isp < nsp; isp++) {
    const size_t off = (size_t)isp * c->sendcap * 6;
    const int sd = c->h_cnt[isp], su = c->h_cnt[nsp + isp];
    const int ru = c->h_cnt[2 * nsp + isp], rd = c->h_cnt[3 * nsp + isp];
    if (sd) NC(ncclSend(c->send[0] + off, (size_t)sd * 6, ncclDouble, c->ndown, c->comm, c->st));
    if (su) NC(ncclSend(c->send[1] + off, (size_t)su * 6, ncclDouble, c->nup, c->comm, c->st));
    if (ru) NC(ncclRecv(c->recv[0] + off, (size_t)ru * 6, ncclDouble, c->nup, c->comm, c->st));
    if (rd) NC(ncclRecv(c->recv[1] + off, (size_t)rd * 6, ncclDouble, c->ndown, c->comm, c->st));
    c->n_in[0][isp] = ru;
    c->n_in[1][isp] = rd;
  }
  NC(ncclGroupEnd());
  c->mig_prev = inplace;  // h_cnt = {send down, send up, recv from nup, recv from ndown} of this step
  for (int d = 0; d < 2; d++)
    for (int isp = 0; isp < nsp; isp++) {
      const size_t off = (size_t)isp * c->sendcap;
      if (inplace)  // append at the tail of the destination segments of the current store
        launch_incoming_append(P, c->recv[d] + off * 6, c->n_in[d][isp], isp, c->cstart[c->cur], cursor, c->soa[c->cur],
                               c->ovf, c->ovfsp, c->ovfcnt, c->ovfcap, c->d_err, c->st, nullptr, backs);
      else
        launch_incoming_tag(P, c->recv[d] + off * 6, c->n_in[d][isp], isp, nullptr, c->gcnt,
                            c->in_rank + (size_t)d * nsp * c->sendcap + off, c->d_err, c->st);
      c->launches++;
    }
  return 0;
}

int scatter_arrivals(wm_ctx *c, int dstbuf) {
  const DevParams &P = c->P;
  if (P.nsize == 1) return 0;
  for (int d = 0; d < 2; d++)
    for (int isp = 0; isp < P.nsp; isp++) {
      const size_t off = (size_t)isp * c->sendcap;
      launch_incoming_scatter(P, c->recv[d] + off * 6, c->n_in[d][isp], isp, nullptr, c->cstart[dstbuf],
                              c->in_rank + (size_t)d * P.nsp * c->sendcap + off, c->soa[dstbuf], c->d_err, c->st);
      c->launches++;
    }
  return 0;
}

// new layout of store `dstbuf` from the destination-cell counts in gcnt: cnt = gcnt, cstart = scan of capacities
int scan_counts(wm_ctx *c, int dstbuf) {
  for (int isp = 0; isp < c->P.nsp; isp++) {
    const int *in = c->gcnt + (size_t)isp * c->P.ncell;
    if (c->nbr_r > 0 && c->slack > 0.f) {  // capacity from the densest cell within nbr_r cells in x (c->tight is scratch here)
      int *tmp = c->tight + (size_t)isp * (c->P.ncell + 1);
      launch_nbr_max(c->P, in, tmp, c->nbr_r, c->st);
      c->launches++;
      in = tmp;
    }
    if (launch_scan(in, c->cstart[dstbuf] + (size_t)isp * (c->P.ncell + 1), c->scan_scratch, c->P.ncell, c->slack, c->st, c->cell_floor))
      return fail("grid too large for the prefix scan");
    c->launches += 3;
  }
  CU(cudaMemcpyAsync(c->cnt[dstbuf], c->gcnt, (size_t)c->P.nsp * c->P.ncell * sizeof(int), cudaMemcpyDeviceToDevice, c->st));
  return 0;
}

// exclusive scan of the live counts of the current store -> c->tight (the reference's cumcnt + row bases)
int scan_tight(wm_ctx *c) {
  for (int isp = 0; isp < c->P.nsp; isp++) {
    if (launch_scan(c->cnt[c->cur] + (size_t)isp * c->P.ncell, c->tight + (size_t)isp * (c->P.ncell + 1), c->scan_scratch,
                    c->P.ncell, 0.f, c->st))
      return fail("grid too large for the prefix scan");
    c->launches += 3;
  }
  return 0;
}

int fill_dead(wm_ctx *c, int buf) {
  CU(cudaMemsetAsync(c->pbuf[buf], 0xFF, (size_t)c->P.cap * c->P.nsp * 6 * sizeof(double), c->st));
  return 0;
}

// k_fused_dp leaves the arrivals of a step at the back of their segments; everything but the next k_fused_dp wants one range
int normalize(wm_ctx *c) {
  if (!c->has_back) return 0;
  CU(cudaSetDevice(c->dev));
  launch_normalize(c->P, c->soa[c->cur], c->cstart[c->cur], c->cnt[c->cur], c->cntb[c->cur], c->st);
  c->launches++;
  c->has_back = false;
  return 0;
}
// the state after an upload / initial condition / layout rebuild has no back ranges
int reset_back(wm_ctx *c, int buf) {
  CU(cudaMemsetAsync(c->cntb[buf], 0, (size_t)c->P.nsp * c->P.ncell * sizeof(int), c->st));
  c->has_back = false;
  return 0;
}

int need_state(wm_ctx *c, State s, const char *who, bool keep_back = false) {
  if (!c) return fail("%s: null context", who);
  if (c->state != s) {
    static const char *nm[] = {"EMPTY", "SORTED", "PUSHED", "BOUNDED"};
    return fail("%s: particle state is %s, expected %s (call order of proj/weibel/app.f90:100-107)", who, nm[c->state], nm[s]);
  }
  if (!keep_back) return normalize(c);
  return 0;
}

int set_device(wm_ctx *c) {
  CU(cudaSetDevice(c->dev));
  return 0;
}

}  // namespace

// ================================================================ C ABI
extern "C" {

const char *wm_last_error(void) { return g_err.c_str(); }
int wm_version(void) { return 100; }

static int wm_create_impl(const wm_config *g, wm_ctx **out, wm_ctx **partial);
int wm_create(const wm_config *g, wm_ctx **out) {
  wm_ctx *partial = nullptr;
  const int e = wm_create_impl(g, out, &partial);
  if (e && partial) {  // a failed allocation half way: give everything back (wm_destroy copes with null members)
    const std::string keep = g_err;
    wm_destroy(partial);
    g_err = keep;
    if (out) *out = nullptr;
  }
  return e;
}
static int wm_create_impl(const wm_config *g, wm_ctx **out, wm_ctx **partial) {
  if (!g || !out) return fail("wm_create: null argument");
  *out = nullptr;
  if (g->ndim != 6) return fail("wm_create: ndim must be 6 (x,y,ux,uy,uz,id)");
  if (g->nsp < 1 || g->nsp > WM_NSP_MAX) return fail("wm_create: nsp must be 1..%d", WM_NSP_MAX);
  if (g->bc != WM_BC_PERIODIC && g->bc != WM_BC_RECONNECTION && g->bc != WM_BC_SHOCK)
    return fail("wm_create: boundary kind %d not implemented (periodic, reconnection walls and shock injection are)", g->bc);
  if (g->delx != 1.0) return fail("wm_create: delx must be 1 (common/sort.f90:60 keys on int(x) without /delx; all apps use delx=1)");
  const int nx = g->nxge - g->nxgs + 1, ny = g->nyge - g->nygs + 1, nyl = g->nye - g->nys + 1;
  if (nx < 4 || nyl < 2 || ny < nyl) return fail("wm_create: grid too small (nx>=4, rows per rank>=2)");
  if (g->nsize < 1 || g->nrank < 0 || g->nrank >= g->nsize) return fail("wm_create: bad rank/size");
  if (g->nxgs < 1 || g->nygs < 1) return fail("wm_create: nxgs, nygs must be >= 1 (int() truncation is used as floor)");
  if ((long long)nx * nyl >= (1LL << 31) / 4) return fail("wm_create: slab too large");
  int ndev = 0;
  if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
    return fail("wm_create: no CUDA device available (this library has no CPU fallback)");
  wm_ctx *c = new wm_ctx();
  *partial = c;
  c->cfg = *g;
  if (const char *v = getenv("WM_SLACK")) c->slack = (float)atof(v);
  if (const char *v = getenv("WM_INPLACE")) c->inplace = atoi(v) != 0;
  if (const char *v = getenv("WM_SM")) c->sm = atoi(v);
  if (const char *v = getenv("WM_RIMPLACE")) c->rimplace = atoi(v) != 0;
  if (const char *v = getenv("WM_CG3")) c->cg3 = atoi(v) != 0;
  if (const char *v = getenv("WM_OVERLAP")) c->overlap = atoi(v) != 0;
  if (const char *v = getenv("WM_CG")) c->cg_mode = atoi(v);
  if (const char *v = getenv("WM_MIGSYNC")) c->mig_fast = atoi(v) == 0;
  if (g->flags & WM_FLAG_EXACT_PUSH) c->inplace = false;  // the exact path keeps the reference's two-pass structure
  if (g->device >= 0) {
    c->dev = g->device;
  } else {
    CU(cudaGetDevice(&c->dev));
  }
  CU(cudaSetDevice(c->dev));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, c->dev));
  if (prop.major < 10)
    return fail("wm_create: device %d is sm_%d%d; this library is built for sm_100a only", c->dev, prop.major, prop.minor);
  DevParams &P = c->P;
  P.nx = nx;
  P.nyl = nyl;
  P.nxgs = g->nxgs;
  P.nys = g->nys;
  P.nygs = g->nygs;
  P.ny = ny;
  P.nsize = g->nsize;
  P.bc = g->bc;
  P.pitch = nx + 4;
  P.ntx = (nx + TX - 1) / TX;
  P.nty = (nyl + TY - 1) / TY;
  P.nsp = g->nsp;
  P.ncell = nx * nyl;
  P.cap = 0;
  P.delx = g->delx;
  P.delt = g->delt;
  P.c = g->c;
  P.cc = g->c * g->c;
  P.inv_cc = 1.0 / P.cc;
  P.xlen = nx * g->delx;
  P.ylen = ny * g->delx;
  // walls of boundary_reconnection.f90:82-92 (nxs = nxgs, nxe = nxge; delx = 1 so int(x/delx) < n <=> x < n)
  P.xwlo = (g->nxgs + 1) * g->delx;
  P.xwhi = (g->nxge - 1) * g->delx;
  P.xw2lo = 2. * (g->nxgs + 1) * g->delx;
  P.xw2hi = 2. * (g->nxge - 1) * g->delx;
  P.u0x2 = 0.0;  // WM_BC_SHOCK: wm_set_u_inject replaces xwhi, xw2hi by xend, 2.*xend
  c->nxa = nx;
  for (int s = 0; s < g->nsp; s++) {
    P.q[s] = g->q[s];
    P.r[s] = g->r[s];
  }
  const double pi = 4.0 * std::atan(1.0);
  // field.f90:53-57
  P.f1 = g->c * g->delt / g->delx;
  P.f2 = g->gfac * P.f1 * P.f1;
  P.f3 = 4.0 * pi * g->delx / g->c;
  const double t = g->delx / (g->c * g->delt * g->gfac);
  P.f4 = 4.0 + t * t;
  P.f5 = t * t;
  P.gfac = g->gfac;
  P.pi4dt = 4. * pi * g->delt;
  c->nup = (g->nrank == g->nsize - 1) ? 0 : g->nrank + 1;   // mpi_set.f90:44-47
  c->ndown = (g->nrank == 0) ? g->nsize - 1 : g->nrank - 1;

  {
    // the main stream outranks the second one: when the sort tail (k_place, thousands of CTAs) runs beside the
    // field solve (many small kernels), the field kernels get their SM slots first and k_place fills the rest
    int lo = 0, hi = 0;
    CU(cudaDeviceGetStreamPriorityRange(&lo, &hi));
    CU(cudaStreamCreateWithPriority(&c->st, cudaStreamNonBlocking, hi));
    CU(cudaStreamCreateWithPriority(&c->st2, cudaStreamNonBlocking, lo));
  }
  const size_t ng = (size_t)P.pitch * (nyl + 4);
  CU(cudaMalloc(&c->f.uf, ng * 6 * sizeof(double)));
  CU(cudaMalloc(&c->f.df, ng * 6 * sizeof(double)));
  CU(cudaMalloc(&c->f.tmpf, ng * 6 * sizeof(double)));
  CU(cudaMalloc(&c->f.uj, ng * 3 * sizeof(double)));
  CU(cudaMalloc(&c->f.gkl, ng * 3 * sizeof(double)));
  CU(cudaMalloc(&c->f.phi, ng * 3 * sizeof(double)));
  CU(cudaMalloc(&c->f.p, ng * 3 * sizeof(double)));
  CU(cudaMalloc(&c->f.p2, ng * 3 * sizeof(double)));
  CU(cudaMalloc(&c->f.r, ng * 3 * sizeof(double)));
  CU(cudaMalloc(&c->f.ap, ng * 3 * sizeof(double)));
  CU(cudaMalloc(&c->f.red, (size_t)RED_BLOCKS_MAX * 8 * sizeof(double)));
  CU(cudaMalloc(&c->f.cgstate, cgctl_bytes()));
  CU(cudaMemset(c->f.cgstate, 0, cgctl_bytes()));
  for (double *a : {c->f.uf, c->f.df, c->f.tmpf}) CU(cudaMemset(a, 0, ng * 6 * sizeof(double)));  // df=0: field.f90:109-111
  for (double *a : {c->f.uj, c->f.gkl, c->f.phi, c->f.p, c->f.p2, c->f.r, c->f.ap}) CU(cudaMemset(a, 0, ng * 3 * sizeof(double)));
  {
    // persistent CG: one CTA per SM, the block's p tile in (opt-in) shared memory
    c->nsm = prop.multiProcessorCount;
    c->cgp_smem_max = prop.sharedMemPerBlockOptin > 4096 ? prop.sharedMemPerBlockOptin - 2048 : 0;  // s_red, s_tot are static
    int cbx, cby, rl;
    size_t sm;
    c->cgp_ok = prop.cooperativeLaunch && cgp_plan(nx, nyl, c->nsm, c->cgp_smem_max, &cbx, &cby, &rl, &sm) &&
                cgp_prepare(c->cgp_smem_max) == cudaSuccess;
    (void)cudaGetLastError();
    CU(cudaMalloc(&c->cgp_partial, (size_t)2 * 160 * 2 * sizeof(double)));
    CU(cudaMalloc(&c->cgp_bar, 2 * sizeof(unsigned)));
    CU(cudaMalloc(&c->cgp_out, 8 * sizeof(int)));
    CU(cudaMemset(c->cgp_out, 0, 8 * sizeof(int)));
    CU(cudaMallocHost(&c->h_cgp_out, 8 * sizeof(int)));
  }
  CU(cudaMalloc(&c->rowtmp, (size_t)P.pitch * 2 * 6 * sizeof(double)));
  CU(cudaMalloc(&c->mom, (size_t)7 * (nx + 2) * (nyl + 2) * P.nsp * sizeof(double)));
  CU(cudaMalloc(&c->gcnt, (size_t)P.nsp * P.ncell * sizeof(int)));
  CU(cudaMalloc(&c->tilebase, (size_t)P.ntx * P.nty * P.nsp * 2 * WIN * sizeof(int)));
  CU(cudaMalloc(&c->scan_scratch, (size_t)scan_scratch_ints(P.ncell) * sizeof(int)));
  CU(cudaMalloc(&c->partial, 4096 * 2 * sizeof(double)));
  CU(cudaMallocHost(&c->h_partial, 4096 * 2 * sizeof(double)));
  CU(cudaMalloc(&c->d_err, sizeof(unsigned)));
  CU(cudaMemset(c->d_err, 0, sizeof(unsigned)));
  CU(cudaMallocHost(&c->h_err, sizeof(unsigned)));
  CU(cudaMallocHost(&c->h_cg, 16 * sizeof(int)));
  CU(cudaMallocHost(&c->h_cnt, 4 * WM_NSP_MAX * sizeof(int)));
  CU(cudaMalloc(&c->sendcnt, 2 * WM_NSP_MAX * sizeof(int)));
  CU(cudaMalloc(&c->recvcnt, 2 * WM_NSP_MAX * sizeof(int)));
  CU(cudaMemset(c->sendcnt, 0, 2 * WM_NSP_MAX * sizeof(int)));
  for (auto &e : c->ev_cg) CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
  for (auto &e : c->ev) CU(cudaEventCreate(&e));
  for (auto &e : c->ev_b) CU(cudaEventCreate(&e));
  CU(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
  CU(cudaEventCreateWithFlags(&c->ev_join, cudaEventDisableTiming));
  for (auto &e : c->ev_call) CU(cudaEventCreate(&e));
  *out = c;
  *partial = nullptr;
  return 0;
}

int wm_destroy(wm_ctx *c) {
  if (!c) return 0;
  cudaSetDevice(c->dev);
  cudaDeviceSynchronize();
  if (c->comm) ncclCommDestroy(c->comm);
  for (int b = 0; b < 2; b++) {
    cudaFree(c->pbuf[b]);
    cudaFree(c->cstart[b]);
    cudaFree(c->cnt[b]);
    cudaFree(c->cntb[b]);
    cudaFree(c->send[b]);
    cudaFree(c->recv[b]);
  }
  for (void *p : {(void *)c->tag, (void *)c->gcnt, (void *)c->tilebase, (void *)c->scan_scratch, (void *)c->sendcnt,
                  (void *)c->recvcnt, (void *)c->in_rank, (void *)c->f.uf, (void *)c->f.df, (void *)c->f.tmpf,
                  (void *)c->f.uj, (void *)c->f.gkl, (void *)c->f.phi, (void *)c->f.p, (void *)c->f.p2, (void *)c->f.r, (void *)c->f.ap,
                  (void *)c->f.red, (void *)c->f.cgstate, (void *)c->rowtmp, (void *)c->mom, (void *)c->partial,
                  (void *)c->d_err})
    cudaFree(p);
  for (int k = 0; k < c->cgp_nopen; k++) cudaIpcCloseMemHandle(c->cgp_ipc_open[k]);
  cudaFree(c->cgp_sh_mine);
  cudaFree(c->d_rowoff);
  cudaFree(c->gen_stage);
  cudaFree(c->cgp_partial);
  cudaFree(c->cgp_bar);
  cudaFree(c->cgp_out);
  cudaFreeHost(c->h_cgp_out);
  for (auto &e : c->ev_b) if (e) cudaEventDestroy(e);
  if (c->ev_fork) cudaEventDestroy(c->ev_fork);
  if (c->ev_join) cudaEventDestroy(c->ev_join);
  cudaFree(c->cnt_tail);
  cudaFree(c->cntb_tail);
  cudaFree(c->tight);
  cudaFree(c->ovf);
  cudaFree(c->ovfsp);
  cudaFree(c->ovfrank);
  cudaFree(c->ovfcnt);
  cudaFreeHost(c->h_ovf);
  cudaFreeHost(c->h_partial);
  cudaFreeHost(c->h_err);
  cudaFreeHost(c->h_cg);
  cudaFreeHost(c->h_cnt);
  for (auto &e : c->ev_cg) if (e) cudaEventDestroy(e);
  for (auto &e : c->ev) if (e) cudaEventDestroy(e);
  for (auto &e : c->ev_call) if (e) cudaEventDestroy(e);
  if (c->st) cudaStreamDestroy(c->st);
  if (c->st2) cudaStreamDestroy(c->st2);
  (void)cudaGetLastError();
  delete c;
  return 0;
}

// Peer memory for the kernels that exchange data themselves (k_cg_persist): every rank exports its CgpShared block and its
// CG residual array with CUDA IPC, the handles travel once through NCCL, and every rank maps what it needs.  All ranks then
// agree (all-reduce) on whether the mapped path is usable; if not, the ring falls back to NCCL calls per CG iteration.
static int ring_map_peers(wm_ctx *c) {
  const int N = c->P.nsize, me = c->cfg.nrank;
  struct Info {
    cudaIpcMemHandle_t sh, r;
    int nyl, ok, pad[2];
  };
  static_assert(sizeof(Info) % 8 == 0, "Info is exchanged as 8-byte words");
  int ok = (c->cg_mode && c->cgp_ok && N <= CGP_MAXR) ? 1 : 0;
  Info mine{};
  CgpShared *sh = nullptr;
  CU(cudaMalloc(&sh, 2u << 20));  // its own 2 MiB block: the IPC mapping exposes nothing else
  CU(cudaMemset(sh, 0, 2u << 20));
  c->cgp_sh_mine = sh;
  if (me < CGP_MAXR) c->cgp_sh[me] = sh;
  if (ok && (cudaIpcGetMemHandle(&mine.sh, sh) != cudaSuccess || cudaIpcGetMemHandle(&mine.r, c->f.r) != cudaSuccess)) ok = 0;
  (void)cudaGetLastError();
  mine.nyl = c->P.nyl;
  mine.ok = ok;
  Info *d_all = nullptr;
  std::vector<Info> all(N);
  CU(cudaMalloc(&d_all, sizeof(Info) * N));
  CU(cudaMemcpy(d_all + me, &mine, sizeof(Info), cudaMemcpyHostToDevice));
  NC(ncclAllGather(d_all + me, d_all, sizeof(Info) / 8, ncclUint64, c->comm, c->st));
  CU(cudaStreamSynchronize(c->st));
  CU(cudaMemcpy(all.data(), d_all, sizeof(Info) * N, cudaMemcpyDeviceToHost));
  for (int q = 0; q < N; q++) ok = ok && all[q].ok;
  if (ok) {
    auto open = [&](const cudaIpcMemHandle_t &h, void **out) {
      if (cudaIpcOpenMemHandle(out, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
        (void)cudaGetLastError();
        return false;
      }
      c->cgp_ipc_open[c->cgp_nopen++] = *out;
      return true;
    };
    for (int q = 0; q < N && ok; q++)
      if (q != me) ok = open(all[q].sh, reinterpret_cast<void **>(&c->cgp_sh[q]));
    if (ok) ok = open(all[c->nup].r, reinterpret_cast<void **>(&c->cgp_r_up));
    if (ok) {
      if (c->ndown == c->nup)
        c->cgp_r_down = c->cgp_r_up;
      else
        ok = open(all[c->ndown].r, reinterpret_cast<void **>(&c->cgp_r_down));
    }
    c->cgp_nyl_down = all[c->ndown].nyl;
    c->cgp_nyl_up = all[c->nup].nyl;
  }
  // everybody or nobody
  int *d_ok = reinterpret_cast<int *>(d_all);
  CU(cudaMemcpy(d_ok, &ok, sizeof(int), cudaMemcpyHostToDevice));
  NC(ncclAllReduce(d_ok, d_ok, 1, ncclInt, ncclMin, c->comm, c->st));
  CU(cudaStreamSynchronize(c->st));
  CU(cudaMemcpy(&ok, d_ok, sizeof(int), cudaMemcpyDeviceToHost));
  CU(cudaFree(d_all));
  c->cgp_ring = ok != 0;
  return 0;
}

int wm_comm_unique_id(void *id128) {
  static_assert(sizeof(ncclUniqueId) == WM_UNIQUE_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  NC(ncclGetUniqueId(&id));
  memcpy(id128, &id, sizeof id);
  return 0;
}

int wm_comm_init(wm_ctx *c, const void *id128) {
  if (!c) return fail("wm_comm_init: null context");
  WM(set_device(c));
  if (c->P.nsize == 1) return 0;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof id);
  NC(ncclCommInitRank(&c->comm, c->P.nsize, id, c->cfg.nrank));
  return ring_map_peers(c);
}

static int ensure_migration_buffers(wm_ctx *c, long long n_per_species) {
  if (c->P.nsize == 1 || c->send[0]) return 0;
  long long per_row = n_per_species / c->P.nyl + 1;
  c->sendcap = (int)std::min<long long>(per_row + 4096, (1LL << 30));
  const size_t bytes = (size_t)c->P.nsp * c->sendcap * 6 * sizeof(double);
  for (int d = 0; d < 2; d++) {
    CU(cudaMalloc(&c->send[d], bytes));
    CU(cudaMalloc(&c->recv[d], bytes));
  }
  CU(cudaMalloc(&c->in_rank, (size_t)2 * c->P.nsp * c->sendcap * sizeof(int)));
  return 0;
}

// ---------------------------------------------------------------- residency
static int count_host(const wm_ctx *c, const int32_t *np2, long long n[WM_NSP_MAX]) {
  const int nyl = c->P.nyl;
  for (int isp = 0; isp < c->P.nsp; isp++) {
    n[isp] = 0;
    for (int jl = 0; jl < nyl; jl++) {
      const int v = np2[jl + nyl * isp];
      if (v < 0 || v > c->cfg.np) return fail("np2(%d,%d)=%d outside 0..np=%d", jl, isp, v, c->cfg.np);
      n[isp] += v;
    }
  }
  return 0;
}

// rows of the host array -> tight AoS records on the device (in the idle particle store)
static int stage_rows_h2d(wm_ctx *c, const double *up, const int32_t *np2, double *stage, const long long n[WM_NSP_MAX]) {
  const int nyl = c->P.nyl;
  long long off = 0;
  for (int isp = 0; isp < c->P.nsp; isp++)
    for (int jl = 0; jl < nyl; jl++) {
      const int v = np2[jl + nyl * isp];
      if (v) CU(cudaMemcpyAsync(stage + off * 6, up + host_up_index(c->cfg, isp, jl), (size_t)v * 6 * sizeof(double), cudaMemcpyHostToDevice, c->st));
      off += v;
    }
  (void)n;
  return 0;
}

// tight AoS records in `stage` (species 0 first, then species 1), any order -> the cell-sorted store: sort__bucket on the device
// (histogram by global atomics, prefix scan into segments with slack, scatter)                 common/sort.f90:36-82
static int bucket_stage(wm_ctx *c, double *stage, const long long n[WM_NSP_MAX], const char *who) {
  const DevParams &P = c->P;
  WM(zero_sort_state(c));
  int *rank = reinterpret_cast<int *>(c->tag);
  long long off = 0;
  for (int isp = 0; isp < P.nsp; isp++) {
    launch_incoming_tag(P, stage + off * 6, (int)n[isp], isp, nullptr, c->gcnt, rank + off, c->d_err, c->st);
    off += n[isp];
  }
  WM(scan_counts(c, c->cur));
  WM(fill_dead(c, c->cur));
  WM(reset_back(c, c->cur));
  off = 0;
  for (int isp = 0; isp < P.nsp; isp++) {
    launch_incoming_scatter(P, stage + off * 6, (int)n[isp], isp, nullptr, c->cstart[c->cur], rank + off, c->soa[c->cur], c->d_err, c->st);
    off += n[isp];
  }
  c->launches += 2 * P.nsp;
  WM(check_errors(c, who));
  c->state = ST_SORTED;
  return 0;
}

int wm_upload_particles(wm_ctx *c, const double *up, const int32_t *np2) {
  if (!c || !up || !np2) return fail("wm_upload_particles: null argument");
  c->accl_valid = false;
  WM(set_device(c));
  long long n[WM_NSP_MAX];
  WM(count_host(c, np2, n));
  long long nmax = 0;
  for (int isp = 0; isp < c->P.nsp; isp++) nmax = std::max(nmax, n[isp]);
  WM(alloc_particles(c, nmax));
  WM(ensure_migration_buffers(c, nmax));
  double *stage = c->pbuf[c->cur ^ 1];
  WM(stage_rows_h2d(c, up, np2, stage, n));
  return bucket_stage(c, stage, n, "wm_upload_particles");
}

int wm_upload_particles_sorted(wm_ctx *c, const double *up, const int32_t *np2, const int32_t *cumcnt) {
  if (!c || !up || !np2 || !cumcnt) return fail("wm_upload_particles_sorted: null argument");
  c->accl_valid = false;
  WM(set_device(c));
  long long n[WM_NSP_MAX];
  WM(count_host(c, np2, n));
  long long nmax = 0;
  for (int isp = 0; isp < c->P.nsp; isp++) nmax = std::max(nmax, n[isp]);
  WM(alloc_particles(c, nmax));
  WM(ensure_migration_buffers(c, nmax));
  const DevParams &P = c->P;
  double *stage = c->pbuf[c->cur ^ 1];
  WM(stage_rows_h2d(c, up, np2, stage, n));
  // cstart from cumcnt + row bases
  std::vector<int> cs((size_t)P.nsp * (P.ncell + 1));
  for (int isp = 0; isp < P.nsp; isp++) {
    long long base = 0;
    int *o = cs.data() + (size_t)isp * (P.ncell + 1);
    for (int jl = 0; jl < P.nyl; jl++) {
      const int32_t *cc = cumcnt + (size_t)(P.nx + 1) * ((size_t)jl + (size_t)P.nyl * isp);
      if (cc[0] != 0 || cc[P.nx] != np2[jl + P.nyl * isp]) return fail("wm_upload_particles_sorted: cumcnt inconsistent with np2 in row %d", jl);
      for (int li = 0; li < P.nx; li++) o[(size_t)jl * P.nx + li] = (int)(base + cc[li]);
      base += cc[P.nx];
    }
    o[P.ncell] = (int)base;
  }
  // tight offsets -> device, counts -> gcnt, segment layout with slack, then an order-preserving copy
  std::vector<int> hc((size_t)P.nsp * P.ncell);
  for (int isp = 0; isp < P.nsp; isp++)
    for (int cl = 0; cl < P.ncell; cl++)
      hc[(size_t)isp * P.ncell + cl] = cs[(size_t)isp * (P.ncell + 1) + cl + 1] - cs[(size_t)isp * (P.ncell + 1) + cl];
  CU(cudaMemcpyAsync(c->tight, cs.data(), cs.size() * sizeof(int), cudaMemcpyHostToDevice, c->st));
  CU(cudaMemcpyAsync(c->gcnt, hc.data(), hc.size() * sizeof(int), cudaMemcpyHostToDevice, c->st));
  WM(scan_counts(c, c->cur));
  WM(fill_dead(c, c->cur));
  WM(reset_back(c, c->cur));
  long long off = 0;
  for (int isp = 0; isp < P.nsp; isp++) {
    launch_relayout_from_aos(P, stage + off * 6, n[isp], c->tight + (size_t)isp * (P.ncell + 1),
                             c->cstart[c->cur] + (size_t)isp * (P.ncell + 1), c->soa[c->cur], (size_t)isp * P.cap, c->d_err, c->st);
    off += n[isp];
  }
  WM(check_errors(c, "wm_upload_particles_sorted"));  // also synchronises: cs, hc go out of scope
  c->state = ST_SORTED;
  return 0;
}

int wm_upload_field(wm_ctx *c, const double *uf) {
  if (!c || !uf) return fail("wm_upload_field: null argument");
  WM(set_device(c));
  const size_t ng = (size_t)c->P.pitch * (c->P.nyl + 4);
  CU(cudaMemcpyAsync(c->f.uf, uf, ng * 6 * sizeof(double), cudaMemcpyHostToDevice, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

static int download_store(wm_ctx *c, int buf, double *up, int32_t *np2, int32_t *cumcnt, double *stage_override) {
  c->accl_valid = false;
  const DevParams &P = c->P;
  std::vector<int> cs((size_t)P.nsp * (P.ncell + 1));
  WM(scan_tight(c));
  CU(cudaMemcpyAsync(cs.data(), c->tight, cs.size() * sizeof(int), cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  double *stage = stage_override ? stage_override : c->pbuf[buf ^ 1];
  long long off = 0;
  for (int isp = 0; isp < P.nsp; isp++) {
    const int *o = cs.data() + (size_t)isp * (P.ncell + 1);
    const long long n = o[P.ncell];
    if (up)
      launch_relayout_to_aos(P, c->soa[c->cur], c->soa[buf], (size_t)isp * P.cap, c->cstart[c->cur] + (size_t)isp * (P.ncell + 1),
                             c->tight + (size_t)isp * (P.ncell + 1), stage + off * 6, c->st);
    for (int jl = 0; jl < P.nyl; jl++) {
      const int rb = o[(size_t)jl * P.nx], re = o[(size_t)(jl + 1) * P.nx];
      if (re - rb > c->cfg.np) return fail("memory over (np2 > np): row %d species %d holds %d > np=%d (boundary_periodic.f90:231-234)", jl, isp, re - rb, c->cfg.np);
      if (np2) np2[jl + P.nyl * isp] = re - rb;
      if (cumcnt) {
        int32_t *cc = cumcnt + (size_t)(P.nx + 1) * ((size_t)jl + (size_t)P.nyl * isp);
        for (int li = 0; li <= P.nx; li++) cc[li] = o[(size_t)jl * P.nx + li] - rb;
      }
      if (up && re > rb)
        CU(cudaMemcpyAsync(up + host_up_index(c->cfg, isp, jl), stage + (off + rb) * 6, (size_t)(re - rb) * 6 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    }
    off += n;
  }
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

int wm_download_particles(wm_ctx *c, double *up, int32_t *np2, int32_t *cumcnt) {
  WM(need_state(c, ST_SORTED, "wm_download_particles"));
  WM(set_device(c));
  return download_store(c, c->cur, up, np2, cumcnt, nullptr);
}

int wm_download_gp(wm_ctx *c, double *gp) {
  if (!c || !gp) return fail("wm_download_gp: null argument");
  if (c->state != ST_PUSHED && c->state != ST_BOUNDED) return fail("wm_download_gp: no pushed state (call wm_particle__solv first)");
  WM(set_device(c));
  int64_t n[WM_NSP_MAX];
  WM(wm_particle_counts(c, n));
  long long tot = 0;
  for (int isp = 0; isp < c->P.nsp; isp++) tot += n[isp];
  double *tmp = nullptr;
  CU(cudaMalloc(&tmp, (size_t)std::max<long long>(tot, 1) * 6 * sizeof(double)));
  const int e = download_store(c, c->cur ^ 1, gp, nullptr, nullptr, tmp);
  cudaFree(tmp);
  return e;
}

int wm_download_field(wm_ctx *c, double *uf) {
  if (!c || !uf) return fail("wm_download_field: null argument");
  WM(set_device(c));
  const size_t ng = (size_t)c->P.pitch * (c->P.nyl + 4);
  CU(cudaMemcpyAsync(uf, c->f.uf, ng * 6 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

int wm_download_current(wm_ctx *c, double *uj) {
  if (!c || !uj) return fail("wm_download_current: null argument");
  WM(set_device(c));
  const size_t ng = (size_t)c->P.pitch * (c->P.nyl + 4);
  CU(cudaMemcpyAsync(uj, c->f.uj, ng * 3 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

int wm_download_dfield(wm_ctx *c, double *df) {
  if (!c || !df) return fail("wm_download_dfield: null argument");
  WM(set_device(c));
  const size_t ng = (size_t)c->P.pitch * (c->P.nyl + 4);
  CU(cudaMemcpyAsync(df, c->f.df, ng * 6 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

int wm_particle_counts(wm_ctx *c, int64_t *n) {
  if (!c || !n) return fail("wm_particle_counts: null argument");
  WM(set_device(c));
  if (c->state == ST_EMPTY) {
    for (int isp = 0; isp < c->P.nsp; isp++) n[isp] = 0;
    return 0;
  }
  WM(normalize(c));
  WM(scan_tight(c));
  for (int isp = 0; isp < c->P.nsp; isp++) {
    int v = 0;
    CU(cudaMemcpyAsync(&v, c->tight + (size_t)isp * (c->P.ncell + 1) + c->P.ncell, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    n[isp] = v;
  }
  return 0;
}

// Particles the driver adds between two steps (proj/shock/app.f90 `inject` :685-850 and `relocate` :611-680 append
// them to the rows of `up`): records (x, y, ux, uy, uz, id) of species isp, any order, inside this rank's slab.
// They are appended at the tail of their cells' segments; a segment that is full sends the record to the overflow
// list and the layout is rebuilt.
static int rebuild_layout(wm_ctx *c, int novf);

int wm_append_particles(wm_ctx *c, int32_t isp, int64_t n, const double *rec) {
  WM(need_state(c, ST_SORTED, "wm_append_particles"));
  if (isp < 0 || isp >= c->P.nsp) return fail("wm_append_particles: species %d out of range", isp);
  if (n < 0 || (n > 0 && !rec)) return fail("wm_append_particles: bad arguments");
  if (n == 0) return 0;
  if (!c->inplace) return fail("wm_append_particles: needs the segment layout with slack (unset WM_INPLACE=0 / WM_SLACK=0)");
  if (n >= (1LL << 31)) return fail("wm_append_particles: too many records in one call");
  WM(set_device(c));
  c->accl_valid = false;
  const DevParams &P = c->P;
  double *d_rec = nullptr;
  CU(cudaMalloc(&d_rec, (size_t)n * 6 * sizeof(double)));
  CU(cudaMemcpyAsync(d_rec, rec, (size_t)n * 6 * sizeof(double), cudaMemcpyHostToDevice, c->st));
  CU(cudaMemsetAsync(c->ovfcnt, 0, sizeof(int), c->st));
  launch_incoming_append(P, d_rec, (int)n, isp, c->cstart[c->cur], c->cnt[c->cur], c->soa[c->cur], c->ovf, c->ovfsp,
                         c->ovfcnt, c->ovfcap, c->d_err, c->st);
  launch_clamp_counts(P, c->cstart[c->cur], c->cnt[c->cur], c->st);
  c->launches += 2;
  CU(cudaMemcpyAsync(c->h_ovf, c->ovfcnt, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  CU(cudaFree(d_rec));
  WM(check_errors(c, "wm_append_particles"));
  if (*c->h_ovf > 0) WM(rebuild_layout(c, *c->h_ovf));
  return 0;
}

// ---------------------------------------------------------------- stage calls
int wm_particle__solv(wm_ctx *c) {
  WM(need_state(c, ST_SORTED, "wm_particle__solv"));
  c->accl_valid = false;
  WM(set_device(c));
  launch_tmpf(fieldp(c), c->f.uf, c->f.tmpf, c->st);
  WM(fill_dead(c, c->cur ^ 1));  // gp: only live slots are written by the push
  const int mode = M_PUSH | ((c->cfg.flags & WM_FLAG_EXACT_PUSH) ? M_EXACT : 0);
  launch_pass1(mode, c->P, p1args(c, c->soa[c->cur], c->soa[c->cur ^ 1], c->P.delt), c->st);
  c->launches += 2;
  CU(cudaGetLastError());
  c->state = ST_PUSHED;
  return 0;
}

int wm_field__ele_cur(wm_ctx *c) {
  WM(need_state(c, ST_PUSHED, "wm_field__ele_cur"));
  WM(set_device(c));
  const size_t ng = (size_t)c->P.pitch * (c->P.nyl + 4);
  CU(cudaMemsetAsync(c->f.uj, 0, ng * 3 * sizeof(double), c->st));  // field.f90:203-205
  launch_pass1(M_DEPOSIT, c->P, p1args(c, c->soa[c->cur], c->soa[c->cur ^ 1], c->P.delt), c->st);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

int wm_boundary__curre(wm_ctx *c) {
  if (!c) return fail("wm_boundary__curre: null context");
  WM(set_device(c));
  WM(bc_curre(c));
  CU(cudaGetLastError());
  return 0;
}

int wm_field__fdtd_i(wm_ctx *c) {
  WM(wm_field__ele_cur(c));
  WM(field_solve(c));
  WM(check_errors(c, "wm_field__fdtd_i"));
  return 0;
}

// u0 of bc__injection: xend = nxe*delx + u0/sqrt(1 + (u0*u0)/(c*c))*delt   proj/shock/boundary_shock.f90:272
int wm_set_u_inject(wm_ctx *c, double u0) {
  if (!c) return fail("wm_set_u_inject: null context");
  if (c->P.bc != WM_BC_SHOCK) return fail("wm_set_u_inject: the context was not created with WM_BC_SHOCK");
  const double cl = c->cfg.c, delx = c->cfg.delx;
  const int nxe = c->cfg.nxgs + c->nxa - 1;
  const double xend = nxe * delx + u0 / std::sqrt(1 + (u0 * u0) / (cl * cl)) * c->cfg.delt;
  c->P.xwhi = xend;
  c->P.xw2hi = +2. * xend;
  c->P.u0x2 = +2. * u0;
  c->u_inject = u0;
  c->u_inject_set = true;
  return 0;
}

// The active x range of the calls that follow (the nxs, nxe arguments of particle__solv, field__fdtd_i, sort__bucket,
// bc__injection ...).  nxs must be nxgs (no app moves it); nxe <= nxge.  Cells beyond nxe must hold no particles.
int wm_set_xrange(wm_ctx *c, int32_t nxs, int32_t nxe) {
  if (!c) return fail("wm_set_xrange: null context");
  if (nxs != c->cfg.nxgs) return fail("wm_set_xrange: nxs must equal nxgs (%d)", c->cfg.nxgs);
  if (nxe > c->cfg.nxge || nxe - nxs < 4) return fail("wm_set_xrange: nxe must be in nxs+4..nxge");
  if (nxe != c->cfg.nxge && c->P.bc != WM_BC_SHOCK)
    return fail("wm_set_xrange: only the shock boundary module works on a sub-range (proj/shock/app.f90)");
  c->nxa = nxe - nxs + 1;
  if (c->P.bc == WM_BC_SHOCK && c->u_inject_set) WM(wm_set_u_inject(c, c->u_inject));
  return 0;
}

int wm_boundary__injection(wm_ctx *c, double u0) {
  WM(wm_set_u_inject(c, u0));
  WM(need_state(c, ST_PUSHED, "wm_boundary__injection"));
  WM(set_device(c));
  launch_bcx(fieldp(c), c->soa[c->cur ^ 1], c->cstart[c->cur], true, c->st);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

int wm_boundary__particle_x(wm_ctx *c) {
  WM(need_state(c, ST_PUSHED, "wm_boundary__particle_x"));
  WM(set_device(c));
  launch_bcx(fieldp(c), c->soa[c->cur ^ 1], c->cstart[c->cur], false, c->st);
  c->launches++;
  CU(cudaGetLastError());
  return 0;
}

int wm_boundary__particle_y(wm_ctx *c) {
  WM(need_state(c, ST_PUSHED, "wm_boundary__particle_y"));
  WM(set_device(c));
  WM(zero_sort_state(c));
  const PartSoA &gp = c->soa[c->cur ^ 1];
  launch_pass1(M_BOUND, c->P, p1args(c, gp, gp, c->P.delt), c->st);
  c->launches++;
  WM(migrate(c));
  WM(check_errors(c, "wm_boundary__particle_y"));
  c->state = ST_BOUNDED;
  return 0;
}

int wm_sort__bucket(wm_ctx *c) {
  WM(need_state(c, ST_BOUNDED, "wm_sort__bucket"));
  WM(set_device(c));
  // (out) up <- (in) gp: the scatter goes back into the store that held the old sorted state
  const int src = c->cur ^ 1, dst = c->cur;
  // the new offsets must not overwrite the old ones while pass 2 still reads them
  WM(scan_counts(c, src));
  // wm_particle__solv dead-filled the gp store before the push, so its x marks the gaps by itself
  WM(fill_dead(c, dst));
  launch_pass2(c->P, c->soa[src], c->soa[dst], c->cstart[dst], c->cstart[src], c->tilebase, c->tag, c->soa[src].x, c->d_err, c->st);
  c->launches++;
  // arrivals use cstart[src] (new offsets) and go to soa[dst]
  {
    const DevParams &P = c->P;
    if (P.nsize > 1)
      for (int d = 0; d < 2; d++)
        for (int isp = 0; isp < P.nsp; isp++) {
          const size_t off = (size_t)isp * c->sendcap;
          launch_incoming_scatter(P, c->recv[d] + off * 6, c->n_in[d][isp], isp, nullptr, c->cstart[src],
                                  c->in_rank + (size_t)d * P.nsp * c->sendcap + off, c->soa[dst], c->d_err, c->st);
          c->launches++;
        }
  }
  WM(check_errors(c, "wm_sort__bucket"));
  std::swap(c->cstart[0], c->cstart[1]);  // new offsets and counts now belong to store `cur`
  std::swap(c->cnt[0], c->cnt[1]);
  c->state = ST_SORTED;
  return 0;
}

// Layout rebuild of the in-place sort: a segment overflowed during the last step, its surplus records
// are in the overflow list.  New segment offsets with fresh slack from the current counts (+ the
// parked records), order-preserving copy into the other store, then the parked records.
static int rebuild_layout(wm_ctx *c, int novf) {
  const DevParams &P = c->P;
  WM(normalize(c));
  const int dst = c->cur ^ 1;
  WM(reset_back(c, dst));
  if (novf > c->ovfcap) return fail("in-place sort: %d records overflowed their segments, list holds %d", novf, c->ovfcap);
  launch_clamp_counts(P, c->cstart[c->cur], c->cnt[c->cur], c->st);  // the surplus of a full segment is in the overflow list
  CU(cudaMemcpyAsync(c->gcnt, c->cnt[c->cur], (size_t)P.nsp * P.ncell * sizeof(int), cudaMemcpyDeviceToDevice, c->st));
  launch_incoming_tag(P, c->ovf, novf, 0, c->ovfsp, c->gcnt, c->ovfrank, c->d_err, c->st);
  // A moving density front (the shock) overflows the Poisson slack of the cells just ahead of it step after step: when
  // rebuilds come in quick succession, size the segments for the densest cell of their neighbourhood in x
  // ... and when that is not enough (filaments that compress by tens of per cent within a few steps), widen the slack itself
  if (c->nstep - c->last_rebuild_step < 4) {
    if (c->nbr_r < 8)
      c->nbr_r = c->nbr_r ? 2 * c->nbr_r : 4;
    else if (c->slack < 40.f)
      c->slack *= 1.5f;
  }
  c->last_rebuild_step = c->nstep;
  for (;;) {  // the new layout must fit the store: drop the neighbourhood sizing, then the slack, before giving up
    WM(scan_counts(c, dst));
    int tot[WM_NSP_MAX] = {0, 0};
    for (int isp = 0; isp < P.nsp; isp++)
      CU(cudaMemcpyAsync(&tot[isp], c->cstart[dst] + (size_t)isp * (P.ncell + 1) + P.ncell, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    if ((long long)std::max(tot[0], tot[1]) <= P.cap) break;
    if (c->slack > 6.f) {
      c->slack = std::max(6.f, c->slack / 1.5f);
    } else if (c->nbr_r > 0) {
      c->nbr_r /= 2;
      if (c->nbr_r < 4) c->nbr_r = 0;
    } else if (c->cell_floor > 0) {
      c->cell_floor /= 2;
    } else if (c->slack > 0.5f) {
      c->slack *= 0.5f;
    } else {
      return fail("memory over: the particle store (%lld slots per species) cannot hold the segments any more; raise wm_config.capacity "
                  "(cf. boundary_periodic.f90:231-234)", P.cap);
    }
  }
  for (int isp = 0; isp < P.nsp; isp++)
    launch_relayout_soa(P, c->soa[c->cur], c->cstart[c->cur] + (size_t)isp * (P.ncell + 1), c->soa[dst],
                        c->cstart[dst] + (size_t)isp * (P.ncell + 1), (size_t)isp * P.cap, c->d_err, c->st);
  launch_incoming_scatter(P, c->ovf, novf, 0, c->ovfsp, c->cstart[dst], c->ovfrank, c->soa[dst], c->d_err, c->st);
  // dead marks behind the live particles of every segment (a memset of the whole store cost more than the copy itself)
  launch_mark_gaps(P, c->soa[dst].x, c->cstart[dst], c->cnt[dst], c->st);
  c->launches += 3 + P.nsp;
  c->cur = dst;
  c->rebuilds++;
  return 0;
}

static int wm_step_impl(wm_ctx *c, int32_t nsteps);
int wm_step(wm_ctx *c, int32_t nsteps) {
  const int e = wm_step_impl(c, nsteps);
  // a step that bailed out half way leaves the stores inconsistent: nothing but a fresh upload may follow
  if (e && c && c->state == ST_SORTED && e != 2) c->state = ST_EMPTY;
  return e;
}
static int wm_step_impl(wm_ctx *c, int32_t nsteps) {
  if (need_state(c, ST_SORTED, "wm_step", true)) return 2;  // wrong state on entry: nothing was touched
  c->accl_valid = false;
  WM(set_device(c));
  const DevParams &P = c->P;
  const size_t ng = (size_t)P.pitch * (P.nyl + 4);
  const int mode = M_PUSH | M_DEPOSIT | M_BOUND | ((c->cfg.flags & WM_FLAG_EXACT_PUSH) ? M_EXACT : 0);
  const bool inplace = c->inplace && !(c->cfg.flags & WM_FLAG_EXACT_PUSH);
  // k_fused_sm<TAIL> places the in-tile cell changers and retires the vacated slots itself: k_place_rim for the rest
  const bool dp = inplace && c->sm == 5;  // k_fused_dp: store `cur` -> the other store, cell changers placed directly
  const bool rim = inplace && !dp && c->sm && fused_sm_has_tail(c->sm) && c->rimplace;
  if (!dp) WM(normalize(c));
  if (P.bc == WM_BC_SHOCK) {
    if (!c->u_inject_set) return fail("wm_step: WM_BC_SHOCK needs wm_set_u_inject(u0) first (bc__injection, proj/shock/app.f90:113)");
    if (!(c->cfg.flags & WM_FLAG_EXACT_PUSH) && !(inplace && c->sm))
      return fail("wm_step: WM_BC_SHOCK is implemented by k_fused_sm and the exact path only (unset WM_SM=0 / WM_INPLACE=0)");
  }
  CU(cudaEventRecord(c->ev_call[0], c->st));
  for (int it = 0; it < nsteps; it++) {
    if (c->timing) CU(cudaEventRecord(c->ev[0], c->st));
    // particle__solv + ele_cur + bc__particle_x/y + (histogram | move of the cell changers) in one pass, in place
    launch_tmpf(fieldp(c), c->f.uf, c->f.tmpf, c->st);
    CU(cudaMemsetAsync(c->f.uj, 0, ng * 3 * sizeof(double), c->st));
    if (dp) {  // (no destination-cell counters: only the leavers' counters)
      if (P.nsize > 1) CU(cudaMemsetAsync(c->sendcnt, 0, 2 * WM_NSP_MAX * sizeof(int), c->st));
    } else {
      WM(zero_sort_state(c));
    }
    if (inplace) CU(cudaMemsetAsync(c->ovfcnt, 0, sizeof(int), c->st));
    const PartSoA &a = c->soa[c->cur];
    if (c->timing) CU(cudaEventRecord(c->ev[1], c->st));
    if (c->cfg.flags & WM_FLAG_EXACT_PUSH)
      launch_pass1(mode, P, p1args(c, a, a, P.delt), c->st);
    else if (dp)
      launch_fused_dp(P, p1args(c, a, c->soa[c->cur ^ 1], P.delt), c->st);
    else if (inplace && c->sm)
      launch_fused_sm(P, p1args(c, a, c->soa[c->cur ^ 1], P.delt), c->sm, c->st);
    else if (inplace)
      launch_fused_inplace(P, p1args(c, a, c->soa[c->cur ^ 1], P.delt), c->st);
    else
      launch_fused(P, p1args(c, a, a, P.delt), c->st);
    c->launches += 2;
    if (c->timing) CU(cudaEventRecord(c->ev[2], c->st));
    if (dp) {
      // The pass has written the other store: stayers at the front of their segments (counts in cnt_tail), arrivals from
      // cells of the same tile at the back (counts in cntb_tail).  That store is the current one from here on; the segment
      // offsets are the same, so the offset arrays swap with the stores.
      const int d = c->cur ^ 1;
      std::swap(c->cstart[0], c->cstart[1]);
      std::swap(c->cnt[d], c->cnt_tail);
      std::swap(c->cntb[d], c->cntb_tail);
      c->cur = d;
      c->has_back = true;
      WM(migrate(c, true, true));  // ring exchange of the leavers; arrivals are appended behind the stayers
      if (c->timing) CU(cudaEventRecord(c->ev[3], c->st));  // (ev[2], ev[3]) = migration
      const bool late_fork = c->overlap && cg_persist_usable(c);
      if (late_fork) WM(field_solve_pre(c));
      cudaStream_t ps = c->overlap ? c->st2 : c->st;
      if (c->overlap) {
        CU(cudaEventRecord(c->ev_fork, c->st));
        CU(cudaStreamWaitEvent(c->st2, c->ev_fork, 0));
      }
      if (c->timing) CU(cudaEventRecord(c->ev_b[0], ps));
      // arrivals from other tiles' cells (the window rim: 1.8 % of the particles) behind the stayers of their new cells
      launch_place_rim2(P, c->tag, c->soa[d], c->cstart[d], c->cnt[d], c->cntb[d], c->ovf, c->ovfsp, c->ovfcnt, c->ovfcap, c->d_err, ps);
      c->launches++;
      CU(cudaMemcpyAsync(c->h_ovf, c->ovfcnt, sizeof(int), cudaMemcpyDeviceToHost, ps));
      if (c->timing) CU(cudaEventRecord(c->ev_b[1], ps));
      if (c->overlap) CU(cudaEventRecord(c->ev_join, c->st2));
      if (late_fork)
        WM(field_solve_post(c));
      else
        WM(field_solve(c));
      if (c->timing) CU(cudaEventRecord(c->ev[4], c->st));
      if (c->overlap) CU(cudaStreamWaitEvent(c->st, c->ev_join, 0));
    } else if (inplace && c->overlap) {
      // The rest of the sort (migration, k_place_rim -- or k_place + k_mark_dead on the variant paths) only needs what the fused pass left behind, the rest
      // of field__fdtd_i only needs uj: the two run side by side, the sort on the second stream.  (NCCL calls stay
      // on the main stream, in program order.)
      WM(migrate(c, true));  // ring exchange of the leavers; arrivals are appended to their segments
      // The persistent CG kernel takes every SM (one 1024-thread CTA each): a sort tail started before it would have to
      // drain first.  So with k_place_rim (0.3 ms) the fork comes after the CG solves and the tail runs beside the
      // streaming kernels that follow (bc__dfield, delta-E, uf update); the longer variants fork before the field solve.
      const bool late_fork = rim && cg_persist_usable(c);
      if (c->timing) CU(cudaEventRecord(c->ev[3], c->st));  // (ev[2], ev[3]) = migration
      if (late_fork) WM(field_solve_pre(c));
      CU(cudaEventRecord(c->ev_fork, c->st));
      CU(cudaStreamWaitEvent(c->st2, c->ev_fork, 0));
      if (c->timing) CU(cudaEventRecord(c->ev_b[0], c->st2));
      launch_place(P, c->pbuf[c->cur ^ 1], c->tag, a, c->cstart[c->cur], c->cnt_tail, c->tilebase, c->ovf, c->ovfsp, c->ovfcnt,
                   c->ovfcap, c->d_err, rim, c->st2);
      if (!rim) launch_mark_dead(P, a.x, c->cstart[c->cur], c->cnt[c->cur], c->cnt_tail, c->st2);  // (rim: done by the fused pass)
      c->launches += rim ? 1 : 2;
      std::swap(c->cnt[c->cur], c->cnt_tail);  // cnt_tail held the new counts
      CU(cudaMemcpyAsync(c->h_ovf, c->ovfcnt, sizeof(int), cudaMemcpyDeviceToHost, c->st2));
      if (c->timing) CU(cudaEventRecord(c->ev_b[1], c->st2));
      CU(cudaEventRecord(c->ev_join, c->st2));
      if (late_fork)
        WM(field_solve_post(c));
      else
        WM(field_solve(c));
      if (c->timing) CU(cudaEventRecord(c->ev[4], c->st));  // (ev[3], ev[4]) = field solve, the sort tail running beside (part of) it
      CU(cudaStreamWaitEvent(c->st, c->ev_join, 0));
    } else if (inplace) {
      // rest of field__fdtd_i
      WM(field_solve(c));
      if (c->timing) CU(cudaEventRecord(c->ev[3], c->st));
      // ring exchange of the leavers; arrivals are appended to their segments
      WM(migrate(c, true));
      if (c->timing) CU(cudaEventRecord(c->ev[4], c->st));
      // sort__bucket, reduced to the cell changers: append them to their new segments
      launch_place(P, c->pbuf[c->cur ^ 1], c->tag, a, c->cstart[c->cur], c->cnt_tail, c->tilebase, c->ovf, c->ovfsp, c->ovfcnt,
                   c->ovfcap, c->d_err, rim, c->st);
      if (!rim) launch_mark_dead(P, a.x, c->cstart[c->cur], c->cnt[c->cur], c->cnt_tail, c->st);
      c->launches += rim ? 1 : 2;
      std::swap(c->cnt[c->cur], c->cnt_tail);  // cnt_tail held the new counts
      CU(cudaMemcpyAsync(c->h_ovf, c->ovfcnt, sizeof(int), cudaMemcpyDeviceToHost, c->st));
    } else {
      // rest of field__fdtd_i
      WM(field_solve(c));
      if (c->timing) CU(cudaEventRecord(c->ev[3], c->st));
      // migration, prefix scan
      WM(migrate(c));
      const int dst = c->cur ^ 1;
      WM(scan_counts(c, dst));
      WM(fill_dead(c, dst));
      if (c->timing) CU(cudaEventRecord(c->ev[4], c->st));
      // sort__bucket scatter
      launch_pass2(P, a, c->soa[dst], c->cstart[c->cur], c->cstart[dst], c->tilebase, c->tag, a.x, c->d_err, c->st);
      c->launches++;
      WM(scatter_arrivals(c, dst));
      c->cur = dst;
    }
    CU(cudaEventRecord(c->ev[5], c->st));
    CU(cudaEventSynchronize(c->ev[5]));
    if (c->timing) {
      float t[5];
      for (int k = 0; k < 5; k++) CU(cudaEventElapsedTime(&t[k], c->ev[k], c->ev[k + 1]));
      c->ms[0] += t[1];
      if (dp || (inplace && c->overlap)) {
        float tb;
        CU(cudaEventElapsedTime(&tb, c->ev_b[0], c->ev_b[1]));
        c->ms[1] += t[3];         // field solve
        c->ms[2] += t[0] + t[2];  // prep + migration
        c->ms[3] += tb;           // k_place_rim (k_place + k_mark_dead), overlapped with the field solve
      } else {
        c->ms[1] += t[2];
        c->ms[2] += t[0] + t[3];
        c->ms[3] += t[4];
      }
    }
    c->nstep++;
    if (inplace && *c->h_ovf > 0) WM(rebuild_layout(c, *c->h_ovf));
  }
  CU(cudaEventRecord(c->ev_call[1], c->st));
  CU(cudaEventSynchronize(c->ev_call[1]));
  {
    float t;
    CU(cudaEventElapsedTime(&t, c->ev_call[0], c->ev_call[1]));
    c->ms[4] += t;
  }
  WM(check_errors(c, "wm_step"));
  return 0;
}

// ---------------------------------------------------------------- host-array (drop-in) calls
int wm_host_steps(wm_ctx *c, double *up, double *uf, int32_t *np2, int32_t *cumcnt, int32_t nsteps) {
  WM(wm_upload_particles_sorted(c, up, np2, cumcnt));
  WM(wm_upload_field(c, uf));
  WM(wm_step(c, nsteps));
  WM(wm_download_particles(c, up, np2, cumcnt));
  WM(wm_download_field(c, uf));
  return 0;
}

int wm_host_step(wm_ctx *c, double *up, double *uf, int32_t *np2, int32_t *cumcnt) { return wm_host_steps(c, up, uf, np2, cumcnt, 1); }

int wm_host_particle__solv(wm_ctx *c, double *gp, const double *up, const double *uf, const int32_t *cumcnt, const int32_t *np2) {
  WM(wm_upload_particles_sorted(c, up, np2, cumcnt));
  WM(wm_upload_field(c, uf));
  WM(wm_particle__solv(c));
  WM(wm_download_gp(c, gp));
  return 0;
}

int wm_host_sort__bucket(wm_ctx *c, double *gp_out, const double *up_in, int32_t *cumcnt, const int32_t *np2) {
  WM(wm_upload_particles(c, up_in, np2));
  WM(wm_download_particles(c, gp_out, nullptr, cumcnt));
  return 0;
}

// ---------------------------------------------------------------- diagnostics
int wm_cg_iters(wm_ctx *c, int32_t out[3]) {
  if (!c) return fail("wm_cg_iters: null context");
  if (c->cg_pending) {
    WM(set_device(c));
    CU(cudaStreamSynchronize(c->st));
    WM(cg_collect(c));
  }
  for (int l = 0; l < 3; l++) out[l] = c->cg_ite[l];
  return 0;
}

// FP64 vector peak of this device, measured: a DFMA loop (8 independent chains per thread, 64 warps per SM) timed with CUDA
// events; flop = 2 per DFMA.  The denominator for the FP64-pipe figures of the fused particle kernel.
int wm_fp64_peak(wm_ctx *c, double *tflops) {
  if (!c || !tflops) return fail("wm_fp64_peak: null argument");
  WM(set_device(c));
  const int nb = c->nsm * 4, n = 1 << 15;
  double *out = nullptr;
  CU(cudaMalloc(&out, (size_t)nb * 512 * sizeof(double)));
  launch_fp64_peak(out, nb, 1 << 10, c->st);  // warm-up
  float best = 1e30f;
  for (int rep = 0; rep < 3; rep++) {
    CU(cudaEventRecord(c->ev_call[0], c->st));
    launch_fp64_peak(out, nb, n, c->st);
    CU(cudaEventRecord(c->ev_call[1], c->st));
    CU(cudaEventSynchronize(c->ev_call[1]));
    float ms;
    CU(cudaEventElapsedTime(&ms, c->ev_call[0], c->ev_call[1]));
    best = std::min(best, ms);
  }
  CU(cudaFree(out));
  *tflops = 2.0 * 8.0 * (double)n * 512.0 * nb / (best * 1e-3) / 1e12;
  return 0;
}

int wm_cg_path(wm_ctx *c, int32_t *path) {
  if (!c || !path) return fail("wm_cg_path: null argument");
  *path = cg_persist_usable(c) ? (c->P.nsize > 1 ? 2 : 1) : 0;
  return 0;
}

int wm_cg_plan(int32_t nx, int32_t nyl, int32_t nsm, int64_t smem_max, int32_t out[4]) {
  if (!out) return fail("wm_cg_plan: null argument");
  int cbx = 0, cby = 0, rl = 0;
  size_t sm = 0;
  if (!cgp_plan(nx, nyl, nsm, (size_t)smem_max, &cbx, &cby, &rl, &sm)) return fail("wm_cg_plan: the slab does not fit the on-chip solver");
  out[0] = cbx;
  out[1] = cby;
  out[2] = (int32_t)sm;
  out[3] = rl;
  return 0;
}

int wm_energy(wm_ctx *c, double *out) {
  WM(need_state(c, ST_SORTED, "wm_energy"));
  WM(set_device(c));
  const DevParams &P = c->P;
  const int nb = 1024;
  const double pi = 4.0 * std::atan(1.0);
  for (int isp = 0; isp < P.nsp; isp++) {
    launch_kinetic(P, c->soa[c->cur], c->cstart[c->cur], isp, c->partial, nb, c->st);
    CU(cudaMemcpyAsync(c->h_partial, c->partial, nb * sizeof(double), cudaMemcpyDeviceToHost, c->st));
    CU(cudaStreamSynchronize(c->st));
    double s = 0.0;
    for (int k = 0; k < nb; k++) s += c->h_partial[k];
    out[isp] = s;
  }
  launch_field_energy(P, c->f.uf, c->partial, nb, c->st);
  CU(cudaMemcpyAsync(c->h_partial, c->partial, nb * 2 * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  double sb = 0.0, se = 0.0;
  for (int k = 0; k < nb; k++) {
    sb += c->h_partial[2 * k];
    se += c->h_partial[2 * k + 1];
  }
  out[P.nsp] = se / (8 * pi);
  out[P.nsp + 1] = sb / (8 * pi);
  return 0;
}

// bc__mom on the device copy: x fold, then the one-row y fold on the ring   boundary_periodic.f90:571-636
static int bc_mom_device(wm_ctx *c) {
  const DevParams &P = c->P;
  launch_mom_fold_x(P, c->mom, c->st);
  c->launches++;
  if (P.nsize == 1) {
    launch_mom_fold_y_local(P, c->mom, c->st);
    c->launches++;
    return 0;
  }
  if (!c->comm) return fail("nsize > 1 but wm_comm_init has not been called");
  const size_t w = (size_t)(P.nx + 2) * 7;
  const size_t nyp = (size_t)P.nyl + 2;
  for (int isp = 0; isp < P.nsp; isp++) {
    double *b = c->mom + (size_t)isp * nyp * w;
    // my ghost row nys-1 -> ndown ; nup's is added into my nye
    NC(ncclGroupStart());
    NC(ncclSend(b, w, ncclDouble, c->ndown, c->comm, c->st));
    NC(ncclRecv(c->rowtmp, w, ncclDouble, c->nup, c->comm, c->st));
    NC(ncclGroupEnd());
    launch_add_rows(b + (size_t)P.nyl * w, c->rowtmp, (long long)w, c->st);
    // my ghost row nye+1 -> nup ; ndown's is added into my nys
    NC(ncclGroupStart());
    NC(ncclSend(b + (size_t)(P.nyl + 1) * w, w, ncclDouble, c->nup, c->comm, c->st));
    NC(ncclRecv(c->rowtmp, w, ncclDouble, c->ndown, c->comm, c->st));
    NC(ncclGroupEnd());
    launch_add_rows(b + w, c->rowtmp, (long long)w, c->st);
    c->launches += 2;
  }
  return 0;
}

static size_t mom_elems(const wm_ctx *c) { return (size_t)7 * (c->P.nx + 2) * (c->P.nyl + 2) * c->P.nsp; }

int wm_mom_calc__accl(wm_ctx *c) {
  WM(need_state(c, ST_SORTED, "wm_mom_calc__accl"));
  WM(set_device(c));
  const DevParams &P = c->P;
  // half-step acceleration into the idle store, positions copied (mom_calc.f90:34,48-164)
  launch_tmpf(fieldp(c), c->f.uf, c->f.tmpf, c->st);
  WM(fill_dead(c, c->cur ^ 1));
  const int mode = M_PUSH | M_NOMOVE | ((c->cfg.flags & WM_FLAG_EXACT_PUSH) ? M_EXACT : 0);
  launch_pass1(mode, P, p1args(c, c->soa[c->cur], c->soa[c->cur ^ 1], P.delt * 0.5), c->st);
  c->launches += 2;
  CU(cudaGetLastError());
  c->accl_valid = true;
  return 0;
}

static int mom_nvt_device(wm_ctx *c) {
  if (!c->accl_valid) return fail("wm_mom_calc__nvt: call wm_mom_calc__accl first (proj/weibel/app.f90:121-122)");
  CU(cudaMemsetAsync(c->mom, 0, mom_elems(c) * sizeof(double), c->st));
  launch_moments(c->P, c->soa[c->cur ^ 1], c->soa[c->cur].x, c->cstart[c->cur], c->mom, c->st);
  c->launches++;
  return 0;
}

int wm_mom_calc__nvt(wm_ctx *c, double *mom) {
  WM(need_state(c, ST_SORTED, "wm_mom_calc__nvt"));
  if (!mom) return fail("wm_mom_calc__nvt: null argument");
  WM(set_device(c));
  WM(mom_nvt_device(c));
  CU(cudaMemcpyAsync(mom, c->mom, mom_elems(c) * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

int wm_boundary__mom(wm_ctx *c, double *mom) {
  if (!c || !mom) return fail("wm_boundary__mom: null argument");
  WM(set_device(c));
  CU(cudaMemcpyAsync(c->mom, mom, mom_elems(c) * sizeof(double), cudaMemcpyHostToDevice, c->st));
  WM(bc_mom_device(c));
  CU(cudaMemcpyAsync(mom, c->mom, mom_elems(c) * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

int wm_moments(wm_ctx *c, double *mom) {
  if (!mom) return fail("wm_moments: null argument");
  WM(wm_mom_calc__accl(c));
  WM(mom_nvt_device(c));
  WM(bc_mom_device(c));
  CU(cudaMemcpyAsync(mom, c->mom, mom_elems(c) * sizeof(double), cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

// Discrete Gauss law of the current state: out[0] = max over this rank's cells of |div E - 4 pi rho| with
// rho = sum_p q S2 S2 (second-order shape, the deposit's weights), out[1] = max 4 pi sum_p |q| S2 S2 (its scale).
// Charge conservation of the Esirkepov deposit keeps out[0] at roundoff of out[1] for all times.  Periodic boundaries;
// on a ring every rank calls it (the charge that particles of the edge rows deposit into the neighbour's rows is folded
// with one exchange per direction, like bc__curre folds the current).
int wm_gauss_residual(wm_ctx *c, double out[2]) {
  if (!out) return fail("wm_gauss_residual: null argument");
  WM(need_state(c, ST_SORTED, "wm_gauss_residual"));
  if (c->P.bc != WM_BC_PERIODIC) return fail("wm_gauss_residual: periodic boundaries only");
  if (c->P.nsize > 1 && !c->comm) return fail("nsize > 1 but wm_comm_init has not been called");
  WM(set_device(c));
  const DevParams &P = c->P;
  WM(scan_tight(c));  // makes sure the per-species slot totals in cstart are current
  double *rho = nullptr;
  unsigned long long *d_out = nullptr;
  const size_t plane = (size_t)P.nx * (P.nyl + 2);
  CU(cudaMalloc(&rho, 2 * plane * sizeof(double)));
  CU(cudaMalloc(&d_out, 2 * sizeof(unsigned long long)));
  launch_charge_density(P, c->soa[c->cur], c->cstart[c->cur], rho, c->st);
  if (P.nsize > 1) {
    for (int k = 0; k < 2; k++) {
      double *pl = rho + k * plane;
      // my row below the slab -> ndown ; nup's lands in rowtmp and is added to my last row
      NC(ncclGroupStart());
      NC(ncclSend(pl, P.nx, ncclDouble, c->ndown, c->comm, c->st));
      NC(ncclRecv(c->rowtmp, P.nx, ncclDouble, c->nup, c->comm, c->st));
      NC(ncclGroupEnd());
      launch_add_rows(pl + (size_t)P.nyl * P.nx, c->rowtmp, P.nx, c->st);
      // my row above the slab -> nup ; ndown's is added to my first row
      NC(ncclGroupStart());
      NC(ncclSend(pl + (size_t)(P.nyl + 1) * P.nx, P.nx, ncclDouble, c->nup, c->comm, c->st));
      NC(ncclRecv(c->rowtmp, P.nx, ncclDouble, c->ndown, c->comm, c->st));
      NC(ncclGroupEnd());
      launch_add_rows(pl + (size_t)P.nx, c->rowtmp, P.nx, c->st);
    }
    c->launches += 4;
  }
  launch_gauss(P, c->f.uf, rho, d_out, c->st);
  c->launches += 2;
  unsigned long long h[2];
  CU(cudaMemcpyAsync(h, d_out, sizeof h, cudaMemcpyDeviceToHost, c->st));
  CU(cudaStreamSynchronize(c->st));
  CU(cudaFree(rho));
  CU(cudaFree(d_out));
  for (int k = 0; k < 2; k++) {
    long long v = (long long)h[k];
    std::memcpy(&out[k], &v, sizeof(double));
  }
  return 0;
}

int wm_ic_weibel(wm_ctx *c, uint64_t seed, int32_t n0, double vti, double vte, double t_ani, double b0) {
  if (!c) return fail("wm_ic_weibel: null context");
  c->accl_valid = false;
  WM(set_device(c));
  DevParams &P = c->P;
  const long long n = (long long)n0 * P.nx * P.nyl;
  if (n >= (1LL << 31) - 1) return fail("wm_ic_weibel: too many particles per species for one GPU");
  WM(alloc_particles(c, n));
  WM(ensure_migration_buffers(c, n));
  if ((long long)cell_capacity(n0, c->slack) * P.ncell > P.cap)  // before anything is written (the kernel was queued first)
    return fail("wm_ic_weibel: particle capacity %lld too small for %d cells x %d slots (segment slack)", P.cap, P.ncell, cell_capacity(n0, c->slack));
  WM(fill_dead(c, c->cur));
  WM(reset_back(c, c->cur));
  launch_ic_weibel(P, c->soa[c->cur], c->cstart[c->cur], c->cnt[c->cur], seed, n0, vti, vte, t_ani, c->slack, c->st);
  // uniform field Bz = b0 (app.f90:388-399), df = 0
  const size_t ng = (size_t)P.pitch * (P.nyl + 4);
  std::vector<double> h(ng * 6, 0.0);
  for (size_t g = 0; g < ng; g++) h[g * 6 + 2] = b0;
  CU(cudaMemcpyAsync(c->f.uf, h.data(), ng * 6 * sizeof(double), cudaMemcpyHostToDevice, c->st));
  CU(cudaMemsetAsync(c->f.df, 0, ng * 6 * sizeof(double), c->st));
  CU(cudaStreamSynchronize(c->st));
  CU(cudaGetLastError());
  c->state = ST_SORTED;
  return 0;
}

// ---------------------------------------------------------------- device-side particle sources
static void gen_common(wm_ctx *c, GenParams &g) {
  g.nxgs = c->cfg.nxgs;
  g.nygs = c->cfg.nygs;
  g.nxg = c->cfg.nxge - c->cfg.nxgs + 1;
  g.nyg = c->cfg.nyge - c->cfg.nygs + 1;
  g.delx = c->cfg.delx;
  g.delt = c->cfg.delt;
  g.c = c->cfg.c;
  for (int s = 0; s < WM_NSP_MAX; s++) g.q[s] = c->cfg.q[s];
}

int wm_ic_harris(wm_ctx *c, uint64_t seed, int32_t nbg, int32_t ncs, double lcs, double vti, double vte, double b0, double rtemp,
                 double e1) {
  if (!c) return fail("wm_ic_harris: null context");
  if (c->P.nsp != 2) return fail("wm_ic_harris: two species (ions, electrons) expected");
  if (nbg < 0 || ncs < 0 || !(lcs > 0)) return fail("wm_ic_harris: bad arguments");
  c->accl_valid = false;
  WM(set_device(c));
  const DevParams &P = c->P;
  GenParams g{};
  gen_common(c, g);
  g.nbg = nbg;
  g.ncs = ncs;
  g.lcs = lcs;
  g.vti = vti;
  g.vte = vte;
  g.b0 = b0;
  g.rtemp = rtemp;
  g.e1 = e1;
  g.npr = (long long)nbg * (g.nxg - 1) + (long long)(ncs * 2 * lcs);  // np2 = nbg*(nxge-nxgs) + ncs*2*lcs   app.f90:313
  g.ibg = (long long)nbg * (g.nxg - 3);                               // ibg = nbg*(nxge-nxgs-2)            app.f90:422
  const long long n = g.npr * P.nyl;
  if (g.npr > c->cfg.np) return fail("wm_ic_harris: %lld particles per row > np = %d", g.npr, c->cfg.np);
  if (n >= (1LL << 31) - 1) return fail("wm_ic_harris: too many particles per species for one GPU");
  WM(alloc_particles(c, n));
  WM(ensure_migration_buffers(c, n));
  if (2 * n > (long long)P.cap * P.nsp) return fail("wm_ic_harris: capacity too small for the staging area");
  double *stage = c->pbuf[c->cur ^ 1];
  launch_gen_harris(P, g, seed, stage, stage + n * 6, c->f.uf, c->st);
  const size_t ng = (size_t)P.pitch * (P.nyl + 4);
  CU(cudaMemsetAsync(c->f.df, 0, ng * 6 * sizeof(double), c->st));
  c->launches += 2;
  const long long nn[WM_NSP_MAX] = {n, n};
  return bucket_stage(c, stage, nn, "wm_ic_harris");
}

static int gen_rowoff(wm_ctx *c, const std::vector<int> &cnt, int *total) {
  std::vector<int> off(cnt.size() + 1, 0);
  for (size_t k = 0; k < cnt.size(); k++) off[k + 1] = off[k] + cnt[k];
  *total = off.back();
  if (!c->d_rowoff) CU(cudaMalloc(&c->d_rowoff, (size_t)(c->P.nyl + 1) * sizeof(int)));
  CU(cudaMemcpyAsync(c->d_rowoff, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice, c->st));
  CU(cudaStreamSynchronize(c->st));  // `off` goes out of scope
  return 0;
}

int wm_ic_shock(wm_ctx *c, uint64_t seed, int32_t n0, int32_t nxe, double v0, double vti, double vte, double b0, double theta_bn,
                double phi_bn, double l_damp_ini) {
  if (!c) return fail("wm_ic_shock: null context");
  if (c->P.bc != WM_BC_SHOCK || c->P.nsp != 2) return fail("wm_ic_shock: needs a WM_BC_SHOCK context with two species");
  if (n0 < 1 || nxe - c->cfg.nxgs < 4 || nxe > c->cfg.nxge || !(std::fabs(v0) < c->cfg.c)) return fail("wm_ic_shock: bad arguments");
  c->accl_valid = false;
  WM(set_device(c));
  const DevParams &P = c->P;
  GenParams &g = c->gen;
  g = GenParams{};
  gen_common(c, g);
  g.n0 = n0;
  g.it = 0;
  g.v0 = v0;
  g.vti = vti;
  g.vte = vte;
  g.b0 = b0;
  g.theta = theta_bn;
  g.phi = phi_bn;
  g.l_damp = l_damp_ini;
  const int npr = n0 * (nxe - c->cfg.nxgs - 1);  // np2 = n0*(nxe-nxs-1)   proj/shock/app.f90:330
  const long long n = (long long)npr * P.nyl;
  if (!c->cfg.capacity) return fail("wm_ic_shock: set wm_config.capacity (the box fills up: the reference sizes rows for n_ppc*nx*5 particles)");
  c->cell_floor = n0;  // upstream cells are laid out for the density they are going to hold
  WM(alloc_particles(c, n));
  WM(ensure_migration_buffers(c, (long long)n0 * P.nx * P.nyl));
  std::vector<int> cnt(P.nyl, npr);
  int total = 0;
  WM(gen_rowoff(c, cnt, &total));
  double *stage = c->pbuf[c->cur ^ 1];
  launch_field_shock(P, g, 0, nxe, c->f.uf, c->st);
  launch_gen_shock(P, g, seed, 0, nxe, c->d_rowoff, stage, stage + n * 6, c->st);
  const size_t ng = (size_t)P.pitch * (P.nyl + 4);
  CU(cudaMemsetAsync(c->f.df, 0, ng * 6 * sizeof(double), c->st));
  c->launches += 2;
  const long long nn[WM_NSP_MAX] = {n, n};
  WM(bucket_stage(c, stage, nn, "wm_ic_shock"));
  WM(wm_set_xrange(c, c->cfg.nxgs, nxe));
  return 0;
}

// records of one inject / relocate call -> their cells' segments (as wm_append_particles, without the host copy)
static int gen_append(wm_ctx *c, int total, const char *who) {
  const DevParams &P = c->P;
  CU(cudaMemsetAsync(c->ovfcnt, 0, sizeof(int), c->st));
  for (int isp = 0; isp < 2; isp++)
    launch_incoming_append(P, c->gen_stage + (size_t)isp * total * 6, total, isp, c->cstart[c->cur], c->cnt[c->cur], c->soa[c->cur],
                           c->ovf, c->ovfsp, c->ovfcnt, c->ovfcap, c->d_err, c->st, nullptr, c->cntb[c->cur]);
  launch_clamp_counts(P, c->cstart[c->cur], c->cnt[c->cur], c->st, c->cntb[c->cur]);
  c->launches += 3;
  CU(cudaMemcpyAsync(c->h_ovf, c->ovfcnt, sizeof(int), cudaMemcpyDeviceToHost, c->st));
  WM(check_errors(c, who));
  if (*c->h_ovf > 0) WM(rebuild_layout(c, *c->h_ovf));
  return 0;
}

static int gen_stage_reserve(wm_ctx *c, long long records) {
  if (records <= c->gen_stage_cap) return 0;
  if (c->gen_stage) CU(cudaFree(c->gen_stage));
  c->gen_stage = nullptr;
  c->gen_stage_cap = records + records / 4 + 1024;
  CU(cudaMalloc(&c->gen_stage, (size_t)c->gen_stage_cap * 2 * 6 * sizeof(double)));
  return 0;
}

// inject() of proj/shock/app.f90:685-850 on the device: n0 |v0| delt delx ny particles per species enter through the right-hand
// boundary (the fractional part by a random number), spread evenly over all rows of the ring with the remainder on random
// rows; new upstream field values in the columns nxe - 1, nxe.  `it` = time step (keys the random numbers and the ids).
int wm_shock_inject(wm_ctx *c, uint64_t seed, int32_t it) {
  WM(need_state(c, ST_SORTED, "wm_shock_inject", true));  // (the appends know about the back ranges)
  if (c->P.bc != WM_BC_SHOCK || c->gen.n0 < 1) return fail("wm_shock_inject: call wm_ic_shock first");
  WM(set_device(c));
  const DevParams &P = c->P;
  GenParams g = c->gen;
  g.it = it;
  const int nxe = c->cfg.nxgs + c->nxa - 1;
  const int ny = g.nyg;
  // (1) total over the system, (2)+(3) equal shares, remainders on random rows: one level here (rows of the whole ring),
  // which gives every row floor or floor + 1 particles like the reference's two-level split
  auto h64 = [](uint64_t z) {
    z += 0x9E3779B97F4A7C15ULL;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
    return z ^ (z >> 31);
  };
  const double pflux = g.n0 * std::fabs(g.v0) * g.delt * g.delx * ny;
  int nginj = (int)pflux;
  const double u = ((double)(h64(seed ^ h64((uint64_t)it * 2 + 1)) >> 11) + 0.5) * (1.0 / 9007199254740992.0);
  if (u < pflux - (int)pflux) nginj++;
  const int base = nginj / ny, rem = nginj % ny;
  // the `rem` rows with the smallest hash get one more
  std::vector<std::pair<uint64_t, int>> key(ny);
  for (int j = 0; j < ny; j++) key[j] = {h64(seed ^ h64(((uint64_t)it << 20) + (uint64_t)j + 77)), j};
  std::vector<int> extra(ny, 0);
  if (rem) {
    std::nth_element(key.begin(), key.begin() + rem, key.end());
    for (int k = 0; k < rem; k++) extra[key[k].second] = 1;
  }
  std::vector<int> cnt(P.nyl);
  for (int lj = 0; lj < P.nyl; lj++) cnt[lj] = base + extra[P.nys - g.nygs + lj];
  int total = 0;
  WM(gen_rowoff(c, cnt, &total));
  launch_field_shock(P, g, 1, nxe, c->f.uf, c->st);
  c->launches++;
  if (total == 0) return 0;
  WM(gen_stage_reserve(c, total));
  launch_gen_shock(P, g, seed, 2, nxe, c->d_rowoff, c->gen_stage, c->gen_stage + (size_t)total * 6, c->st);
  c->launches++;
  return gen_append(c, total, "wm_shock_inject");
}

// relocate() of proj/shock/app.f90:611-680 on the device: the box grows by one column (nxe + 1) unless it is at nxge, n0 new
// pairs per row in the cell nxe - 1, upstream fields in the columns nxe - 1, nxe; the active range follows (wm_set_xrange).
int wm_shock_relocate(wm_ctx *c, uint64_t seed, int32_t it) {
  WM(need_state(c, ST_SORTED, "wm_shock_relocate", true));
  if (c->P.bc != WM_BC_SHOCK || c->gen.n0 < 1) return fail("wm_shock_relocate: call wm_ic_shock first");
  WM(set_device(c));
  const DevParams &P = c->P;
  int nxe = c->cfg.nxgs + c->nxa - 1;
  if (nxe == c->cfg.nxge) return 0;  // app.f90:619
  nxe++;
  GenParams g = c->gen;
  g.it = it;
  std::vector<int> cnt(P.nyl, g.n0);
  int total = 0;
  WM(gen_rowoff(c, cnt, &total));
  WM(gen_stage_reserve(c, total));
  launch_gen_shock(P, g, seed, 1, nxe, c->d_rowoff, c->gen_stage, c->gen_stage + (size_t)total * 6, c->st);
  launch_field_shock(P, g, 1, nxe, c->f.uf, c->st);
  c->launches += 2;
  WM(wm_set_xrange(c, c->cfg.nxgs, nxe));
  return gen_append(c, total, "wm_shock_relocate");
}

int wm_xrange(wm_ctx *c, int32_t out[2]) {
  if (!c || !out) return fail("wm_xrange: null argument");
  out[0] = c->cfg.nxgs;
  out[1] = c->cfg.nxgs + c->nxa - 1;
  return 0;
}

int wm_timing(wm_ctx *c, double ms[5], int64_t *launches, int32_t reset) {
  if (!c) return fail("wm_timing: null context");
  for (int k = 0; k < 5; k++) {
    if (ms) ms[k] = c->ms[k];
    if (reset) c->ms[k] = 0.0;
  }
  if (launches) *launches = c->launches;
  if (reset) c->launches = 0;
  return 0;
}

int wm_layout_rebuilds(wm_ctx *c, int64_t *n) {
  if (!c || !n) return fail("wm_layout_rebuilds: null argument");
  *n = c->rebuilds;
  return 0;
}

int wm_synchronize(wm_ctx *c) {
  if (!c) return fail("wm_synchronize: null context");
  WM(set_device(c));
  CU(cudaStreamSynchronize(c->st));
  return 0;
}

}  // extern "C"
