// particle_kernels.cu -- sm_100a kernels for the particle side of the hot path.
//
//  k_pass1<MODE>   tile-per-CTA kernel.  MODE selects any of
//                    PUSH    Buneman-Boris push            common/particle.f90:83-169
//                    DEPOSIT Esirkepov 2nd-order deposit    common/field.f90:189-316
//                    BOUND   periodic wraps, row migration  common/boundary_periodic.f90:61-248
//                            + destination-cell histogram   common/sort.f90:57-62
//                  The fused step runs all three in one pass over the particles, in place.
//  k_pass2         scatter into the new cell order          common/sort.f90:71-75
//  k_scan_*        exclusive prefix scan of per-cell counts common/sort.f90:64-69
//
// Design notes:
//  * FP64 and HBM bound, no tensor cores: nothing here is a dense contraction.
//  * a CTA owns a TX x TY tile of cells; cell-centred E/B of the tile (+1 halo) are staged in
//    shared memory as AoS6 so one particle gathers with LDS.128; the current of the tile
//    (+2 halo) is accumulated in shared memory and flushed once with RED.ADD.F64.
//  * GRP=8 threads share one cell: each keeps the cell's 4x5 + 5x4 + 5x5 Esirkepov block in
//    registers (the reference's pjx/pjy/pjz, field.f90:215-217) over all its particles of both
//    species, then one shuffle reduction and one shared-memory add per cell: no per-particle
//    atomics on FP64 data.
//  * destination ranks for the counting sort come from shared-memory integer atomics in the
//    same pass; the tag (window cell << 24 | rank) makes the scatter pass a pure copy.
#include <cstdio>

#include "kernels.h"

namespace wm {

// ---------------------------------------------------------------- arithmetic policy
// EXACT: never contracted, IEEE sqrt/div, the reference's operation order -> bit-identical
// to the CPU path.  FAST: same order, FMA contraction allowed, rsqrt where the reference
// divides by a square root.
template <bool EX>
struct Ar {
  static __device__ __forceinline__ double mul(double a, double b) { return EX ? __dmul_rn(a, b) : a * b; }
  static __device__ __forceinline__ double add(double a, double b) { return EX ? __dadd_rn(a, b) : a + b; }
  static __device__ __forceinline__ double sub(double a, double b) { return EX ? __dsub_rn(a, b) : a - b; }
  static __device__ __forceinline__ double div(double a, double b) { return EX ? __ddiv_rn(a, b) : a / b; }
  static __device__ __forceinline__ double sqrt_(double a) { return EX ? __dsqrt_rn(a) : sqrt(a); }
};

template <int MODE>
__global__ void __launch_bounds__(P1_THREADS, 1) k_pass1(const DevParams P, const Pass1Args a) {
  constexpr bool PUSH = (MODE & M_PUSH) != 0;
  constexpr bool DEPOSIT = (MODE & M_DEPOSIT) != 0;
  constexpr bool BOUND = (MODE & M_BOUND) != 0;
  constexpr bool EX = (MODE & M_EXACT) != 0;
  constexpr bool NOMOVE = (MODE & M_NOMOVE) != 0;
  using A = Ar<EX>;

  __shared__ __align__(16) double s_f[PUSH ? WINY * WINX * 6 : 2];
  __shared__ double s_j[DEPOSIT ? 3 * JY * JX : 1];
  __shared__ int s_cnt[BOUND ? WM_NSP_MAX * WIN : 1];

  const int tid = threadIdx.x;
  const int tile = blockIdx.x;
  const int li0 = (tile % P.ntx) * TX, lj0 = (tile / P.ntx) * TY;
  const int tw = min(TX, P.nx - li0), th = min(TY, P.nyl - lj0);

  if (PUSH) {
    // window cell (lx,ly) = local cell (li0-1+lx, lj0-1+ly) = padded (li0+1+lx, lj0+1+ly)
    const int nw = (tw + 2) * 6;
    for (int e = tid; e < (th + 2) * nw; e += P1_THREADS) {
      const int ly = e / nw, r = e - ly * nw;
      s_f[ly * (WINX * 6) + r] = a.tmpf[((size_t)(lj0 + 1 + ly) * P.pitch + (li0 + 1)) * 6 + r];
    }
  }
  if (DEPOSIT)
    for (int e = tid; e < 3 * JY * JX; e += P1_THREADS) s_j[e] = 0.0;
  if (BOUND)
    for (int e = tid; e < WM_NSP_MAX * WIN; e += P1_THREADS) s_cnt[e] = 0;
  __syncthreads();

  const int grp = tid / GRP, lane = tid % GRP;
  constexpr int NGRP = P1_THREADS / GRP;
  const int ncells = tw * th;
  const double qf_base = P.delx / P.delt;

  for (int c0 = 0; c0 < ncells; c0 += NGRP) {
    const int c = c0 + grp;
    const bool valid = c < ncells;
    double acc[DEPOSIT ? 65 : 1];
    if (DEPOSIT) {
#pragma unroll
      for (int e = 0; e < 65; e++) acc[e] = 0.0;
    }
    int cx = 0, cy = 0;
    if (valid) {
      cy = c / tw;
      cx = c - cy * tw;
      const int cell = (lj0 + cy) * P.nx + (li0 + cx);
      const int gi = P.nxgs + li0 + cx, gj = P.nys + lj0 + cy;  // global cell indices (i, j)
      const double di = (double)gi, dj = (double)gj;
      for (int isp = 0; isp < P.nsp; isp++) {
        const int beg = a.cstart[(size_t)isp * (P.ncell + 1) + cell];
        const int end = beg + a.cnt[(size_t)isp * P.ncell + cell];
        const size_t so = (size_t)isp * P.cap;
        const double qs = P.q[isp];
        // particle.f90:90-92
        const double fac1 = A::mul(A::mul(A::div(qs, P.r[isp]), 0.5), a.delt_push);
        const double txxx = A::mul(fac1, fac1);
        const double fac2 = A::div(A::mul(qs, a.delt_push), P.r[isp]);
        const double qf = qs * qf_base;  // q*delx*d_delt, field.f90:278
        for (int p = beg + lane; p < end; p += GRP) {
          const double x = a.src.x[so + p], y = a.src.y[so + p];
          double xn, yn, un1, un2, un3;
          // second order shape function about the sorted cell     particle.f90:97-105
          // (delx == 1 is asserted at wm_create, so x*d_delx == x; sort.f90:60 assumes it too)
          double dh = A::sub(A::sub(x, 0.5), di);
          const double sxm = A::mul(A::mul(0.5, A::sub(0.5, dh)), A::sub(0.5, dh));
          const double sx0 = A::sub(0.75, A::mul(dh, dh));
          const double sxp = A::mul(A::mul(0.5, A::add(0.5, dh)), A::add(0.5, dh));
          dh = A::sub(A::sub(y, 0.5), dj);
          const double sym = A::mul(A::mul(0.5, A::sub(0.5, dh)), A::sub(0.5, dh));
          const double sy0 = A::sub(0.75, A::mul(dh, dh));
          const double syp = A::mul(A::mul(0.5, A::add(0.5, dh)), A::add(0.5, dh));
          double wmove = 0.0;  // 1/sqrt(1+u^2/c^2) of the new momentum
          if (PUSH) {
            const double u1 = a.src.ux[so + p], u2 = a.src.uy[so + p], u3 = a.src.uz[so + p];
            // 3x3 gather of the six cell-centred components           particle.f90:107-129
            double f[6];
            {
              double row[3][6];
#pragma unroll
              for (int b = 0; b < 3; b++) {
                const double2 *q0 = reinterpret_cast<const double2 *>(&s_f[((cy + b) * WINX + cx) * 6]);
                double t[18];
#pragma unroll
                for (int k = 0; k < 9; k++) {
                  const double2 v = q0[k];
                  t[2 * k] = v.x;
                  t[2 * k + 1] = v.y;
                }
#pragma unroll
                for (int k = 0; k < 6; k++)
                  row[b][k] = A::add(A::add(A::mul(t[k], sxm), A::mul(t[6 + k], sx0)), A::mul(t[12 + k], sxp));
              }
#pragma unroll
              for (int k = 0; k < 6; k++)
                f[k] = A::add(A::add(A::mul(row[0][k], sym), A::mul(row[1][k], sy0)), A::mul(row[2][k], syp));
            }
            const double bpx = f[0], bpy = f[1], bpz = f[2], epx = f[3], epy = f[4], epz = f[5];
            // accel.                                                   particle.f90:132-134
            double uvm1 = A::add(u1, A::mul(fac1, epx));
            double uvm2 = A::add(u2, A::mul(fac1, epy));
            double uvm3 = A::add(u3, A::mul(fac1, epz));
            // rotate                                                   particle.f90:137-148
            const double s2 = A::add(A::add(A::add(P.cc, A::mul(uvm1, uvm1)), A::mul(uvm2, uvm2)), A::mul(uvm3, uvm3));
            double gam, igam;
            if (EX) {
              gam = A::sqrt_(s2);
              igam = A::div(1.0, gam);
            } else {
              igam = rsqrt(s2);
              gam = s2 * igam;
            }
            const double fac1r = A::mul(fac1, igam);
            const double b2 = A::add(A::add(A::mul(bpx, bpx), A::mul(bpy, bpy)), A::mul(bpz, bpz));
            const double fac2r = A::div(fac2, A::add(gam, A::mul(A::mul(txxx, b2), igam)));
            const double uvm4 = A::add(uvm1, A::mul(fac1r, A::sub(A::mul(uvm2, bpz), A::mul(uvm3, bpy))));
            const double uvm5 = A::add(uvm2, A::mul(fac1r, A::sub(A::mul(uvm3, bpx), A::mul(uvm1, bpz))));
            const double uvm6 = A::add(uvm3, A::mul(fac1r, A::sub(A::mul(uvm1, bpy), A::mul(uvm2, bpx))));
            uvm1 = A::add(uvm1, A::mul(fac2r, A::sub(A::mul(uvm5, bpz), A::mul(uvm6, bpy))));
            uvm2 = A::add(uvm2, A::mul(fac2r, A::sub(A::mul(uvm6, bpx), A::mul(uvm4, bpz))));
            uvm3 = A::add(uvm3, A::mul(fac2r, A::sub(A::mul(uvm4, bpy), A::mul(uvm5, bpx))));
            // accel.                                                   particle.f90:151-153
            un1 = A::add(uvm1, A::mul(fac1, epx));
            un2 = A::add(uvm2, A::mul(fac1, epy));
            un3 = A::add(uvm3, A::mul(fac1, epz));
            if (NOMOVE) {  // mom_calc.f90:151-152
              xn = x;
              yn = y;
            } else {
              // move                                                   particle.f90:156-161
              const double uu = A::add(A::add(A::mul(un1, un1), A::mul(un2, un2)), A::mul(un3, un3));
              if (EX)
                wmove = A::div(1.0, A::sqrt_(A::add(1.0, A::div(uu, P.cc))));
              else
                wmove = rsqrt(1.0 + uu * P.inv_cc);
              xn = A::add(x, A::mul(A::mul(un1, P.delt), wmove));
              yn = A::add(y, A::mul(A::mul(un2, P.delt), wmove));
              if (P.bc == WM_BC_SHOCK && BOUND) {
                // fused step: bc__injection comes before the deposit (proj/shock/app.f90:112-113)
                if (xn < P.xwlo) {
                  xn = A::sub(P.xw2lo, xn);
                  un1 = -un1; un2 = -un2; un3 = -un3;
                } else if (xn > P.xwhi) {
                  xn = A::sub(P.xw2hi, xn);
                  un1 = A::sub(P.u0x2, un1); un2 = -un2; un3 = -un3;
                  // ele_cur takes vz = uz/gamma from the momentum it finds in gp (field.f90:270-272): the new one
                  const double uu2 = A::add(A::add(A::mul(un1, un1), A::mul(un2, un2)), A::mul(un3, un3));
                  if (EX)
                    wmove = A::div(1.0, A::sqrt_(A::add(1.0, A::div(uu2, P.cc))));
                  else
                    wmove = rsqrt(1.0 + uu2 * P.inv_cc);
                }
              }
            }
          } else {
            xn = a.dst.x[so + p];
            yn = a.dst.y[so + p];
            if (DEPOSIT) {
              un1 = a.dst.ux[so + p];
              un2 = a.dst.uy[so + p];
              un3 = a.dst.uz[so + p];
              wmove = rsqrt(1.0 + (un1 * un1 + un2 * un2 + un3 * un3) * P.inv_cc);
            }
          }

          int incx = 0, incy = 0;
          if (DEPOSIT || BOUND) {
            // new cell                                                  field.f90:238-243,253-255
            const int i2 = __double2int_rz(xn), j2 = __double2int_rz(yn);
            incx = i2 - gi;
            incy = j2 - gj;
            if (!PUSH) {  // stage mode: the positions may already be wrapped
              if (incx > 1) incx -= P.nx;
              if (incx < -1) incx += P.nx;
              if (incy > 1) incy -= P.ny;
              if (incy < -1) incy += P.ny;
            }
            if (incx < -1 || incx > 1 || incy < -1 || incy > 1) {
              atomicOr(a.err, ERR_MOVED_TOO_FAR);
              incx = max(-1, min(1, incx));
              incy = max(-1, min(1, incy));
            }
          }

          if (DEPOSIT) {
            // Esirkepov density decomposition, factorised:
            //  Jx block = Cx (x) Ty, Jy block = Tx (x) Cy, Jz block = Tx (x) Uy + Hx (x) Vy
            // with T = S0 + DS/2, H = S0/2 + DS/3, C = running sum of -q*dx/dt*DS  (field.f90:274-298)
            double dsx[5], dsy[5];
            {
              const double d2 = xn - 0.5 - (double)__double2int_rz(xn);
              const double s1 = 0.5 * (0.5 - d2) * (0.5 - d2), s2 = 0.75 - d2 * d2, s3 = 0.5 * (0.5 + d2) * (0.5 + d2);
              const bool m = incx < 0, z = incx == 0, pl = incx > 0;
              dsx[0] = m ? s1 : 0.0;
              dsx[1] = (m ? s2 : (z ? s1 : 0.0)) - sxm;
              dsx[2] = (m ? s3 : (z ? s2 : s1)) - sx0;
              dsx[3] = (m ? 0.0 : (z ? s3 : s2)) - sxp;
              dsx[4] = pl ? s3 : 0.0;
            }
            {
              const double d2 = yn - 0.5 - (double)__double2int_rz(yn);
              const double s1 = 0.5 * (0.5 - d2) * (0.5 - d2), s2 = 0.75 - d2 * d2, s3 = 0.5 * (0.5 + d2) * (0.5 + d2);
              const bool m = incy < 0, z = incy == 0, pl = incy > 0;
              dsy[0] = m ? s1 : 0.0;
              dsy[1] = (m ? s2 : (z ? s1 : 0.0)) - sym;
              dsy[2] = (m ? s3 : (z ? s2 : s1)) - sy0;
              dsy[3] = (m ? 0.0 : (z ? s3 : s2)) - syp;
              dsy[4] = pl ? s3 : 0.0;
            }
            const double qvz = qs * (un3 * wmove);  // q*gvz, field.f90:270-272,295
            double tx[5], ty[5], hx[5], cxv[4], cyv[4], uy[3], vy[5];
            tx[0] = 0.5 * dsx[0];
            tx[1] = sxm + 0.5 * dsx[1];
            tx[2] = sx0 + 0.5 * dsx[2];
            tx[3] = sxp + 0.5 * dsx[3];
            tx[4] = 0.5 * dsx[4];
            ty[0] = 0.5 * dsy[0];
            ty[1] = sym + 0.5 * dsy[1];
            ty[2] = sy0 + 0.5 * dsy[2];
            ty[3] = syp + 0.5 * dsy[3];
            ty[4] = 0.5 * dsy[4];
            const double third = 1.0 / 3.0;
            hx[0] = third * dsx[0];
            hx[1] = 0.5 * sxm + third * dsx[1];
            hx[2] = 0.5 * sx0 + third * dsx[2];
            hx[3] = 0.5 * sxp + third * dsx[3];
            hx[4] = third * dsx[4];
            cxv[0] = -qf * dsx[0];
            cxv[1] = cxv[0] - qf * dsx[1];
            cxv[2] = cxv[1] - qf * dsx[2];
            cxv[3] = cxv[2] - qf * dsx[3];
            cyv[0] = -qf * dsy[0];
            cyv[1] = cyv[0] - qf * dsy[1];
            cyv[2] = cyv[1] - qf * dsy[2];
            cyv[3] = cyv[2] - qf * dsy[3];
            uy[0] = qvz * sym;
            uy[1] = qvz * sy0;
            uy[2] = qvz * syp;
#pragma unroll
            for (int b = 0; b < 5; b++) vy[b] = qvz * dsy[b];
            // acc layout: [0,20) Jx[b][a'] a'=0..3 <-> a=-1..2 ; [20,40) Jy[b'][a] b'=0..3 <-> b=-1..2 ;
            //             [40,65) Jz[b][a]
#pragma unroll
            for (int b = 0; b < 5; b++)
#pragma unroll
              for (int q = 0; q < 4; q++) acc[b * 4 + q] += cxv[q] * ty[b];
#pragma unroll
            for (int b = 0; b < 4; b++)
#pragma unroll
              for (int q = 0; q < 5; q++) acc[20 + b * 5 + q] += tx[q] * cyv[b];
#pragma unroll
            for (int b = 0; b < 5; b++)
#pragma unroll
              for (int q = 0; q < 5; q++) {
                double v = hx[q] * vy[b];
                if (b >= 1 && b <= 3) v += tx[q] * uy[b - 1];
                acc[40 + b * 5 + q] += v;
              }
          }

          uint32_t tg = 0;
          if (BOUND) {
            // periodic wraps with round-toward -inf adds      boundary_periodic.f90:74,82-88,124,147-154
            const int j2 = gj + incy;  // unwrapped destination row
            const int ic = __double2int_rz(xn), jc = __double2int_rz(yn);
            if (P.bc == WM_BC_SHOCK) {
              // the shock app applies bc__injection before the field solve and has no bc__particle_x in its step
            } else if (P.bc != WM_BC_PERIODIC) {
              // reflecting walls                     proj/reconnection/boundary_reconnection.f90:82-92
              bool flip = false;
              if (xn < P.xwlo) {
                xn = A::sub(P.xw2lo, xn);
                flip = true;
              } else if (xn >= P.xwhi) {
                xn = A::sub(P.xw2hi, xn);
                flip = true;
              }
              if (flip) {
                if (PUSH) {
                  un1 = -un1;
                  un2 = -un2;
                  un3 = -un3;
                }
                incx = __double2int_rz(xn) - gi;  // the sort goes by the reflected position
              }
            } else if (ic < P.nxgs)
              xn = __dadd_rd(xn, P.xlen);
            else if (ic >= P.nxgs + P.nx)
              xn = __dadd_rd(xn, -P.xlen);
            if (jc <= P.nygs - 1)
              yn = __dadd_rd(yn, P.ylen);
            else if (jc >= P.nygs + P.ny)
              yn = __dadd_rd(yn, -P.ylen);
            const bool leaves = (P.nsize > 1) && (j2 < P.nys || j2 >= P.nys + P.nyl);
            if (leaves) {
              // record goes to the neighbour's edge row        boundary_periodic.f90:156-161,174-189
              const int dir = (j2 < P.nys) ? 0 : 1;
              const int pos = atomicAdd(&a.sendcnt[dir * P.nsp + isp], 1);
              if (pos < a.sendcap) {
                double *rec = a.send[dir] + ((size_t)isp * a.sendcap + pos) * 6;
                rec[0] = xn;
                rec[1] = yn;
                rec[2] = PUSH ? un1 : a.dst.ux[so + p];
                rec[3] = PUSH ? un2 : a.dst.uy[so + p];
                rec[4] = PUSH ? un3 : a.dst.uz[so + p];
                rec[5] = __longlong_as_double(a.src.id[so + p]);
              } else {
                atomicOr(a.err, ERR_SENDBUF);
              }
              tg = TAG_DEAD;
            } else {
              const int w = (cy + 1 + incy) * WINX + (cx + 1 + incx);
              const int rk = atomicAdd(&s_cnt[isp * WIN + w], 1);
              tg = TAG_ARRIVAL | ((uint32_t)w << TAG_WSHIFT) | (uint32_t)rk;
            }
            a.tag[so + p] = tg;
          }

          if (PUSH) {
            a.dst.x[so + p] = xn;
            a.dst.y[so + p] = yn;
            a.dst.ux[so + p] = un1;
            a.dst.uy[so + p] = un2;
            a.dst.uz[so + p] = un3;
            if (a.dst.id.p != a.src.id.p) a.dst.id[so + p] = a.src.id[so + p];  // particle.f90:171-175
          } else if (BOUND) {
            a.dst.x[so + p] = xn;
            a.dst.y[so + p] = yn;
          }
        }
      }
    }
    if (DEPOSIT) {
      // sum the GRP partial blocks, then each lane adds its share to the tile   field.f90:304-310
#pragma unroll
      for (int e = 0; e < 65; e++) {
        double v = acc[e];
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        acc[e] = v;
      }
      if (valid) {
#pragma unroll
        for (int e = 0; e < 65; e++) {
          if ((e & (GRP - 1)) == lane) {
            int comp, a2, b2;
            if (e < 20) {
              comp = 0; b2 = e / 4 - 2; a2 = e % 4 - 1;
            } else if (e < 40) {
              comp = 1; b2 = (e - 20) / 5 - 1; a2 = (e - 20) % 5 - 2;
            } else {
              comp = 2; b2 = (e - 40) / 5 - 2; a2 = (e - 40) % 5 - 2;
            }
            atomicAdd(&s_j[(comp * JY + (cy + 2 + b2)) * JX + (cx + 2 + a2)], acc[e]);
          }
        }
      }
    }
  }
  __syncthreads();

  if (DEPOSIT) {
    // one flush of the tile (+halo) into uj: window (jx,jy) = padded (li0+jx, lj0+jy)
    const int jw = tw + 4;
    for (int e = tid; e < 3 * (th + 4) * jw; e += P1_THREADS) {
      const int comp = e / ((th + 4) * jw);
      const int r = e - comp * (th + 4) * jw;
      const int jy = r / jw, jx = r - jy * jw;
      const double v = s_j[(comp * JY + jy) * JX + jx];
      if (v != 0.0) atomicAdd(&a.uj[((size_t)(lj0 + jy) * P.pitch + (li0 + jx)) * 3 + comp], v);
    }
  }
  if (BOUND) {
    // reserve this tile's share of every destination cell                 sort.f90:57-62
    for (int e = tid; e < P.nsp * WIN; e += P1_THREADS) {
      const int isp = e / WIN, w = e - isp * WIN;
      const int n = s_cnt[isp * WIN + w];
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
      a.tilebase[((size_t)tile * P.nsp + isp) * (2 * WIN) + w] = base;
      a.tilebase[((size_t)tile * P.nsp + isp) * (2 * WIN) + WIN + w] = base;
    }
  }
}

// scatter: (out) dst <- (in) src, in the new cell order                      sort.f90:71-75
__global__ void __launch_bounds__(256) k_pass2(const DevParams P, const PartSoA src, const PartSoA dst,
                                               const int *__restrict__ cstart_old, const int *__restrict__ cstart_new,
                                               const int *__restrict__ tilebase, const uint32_t *__restrict__ tag,
                                               const PView<double> keyx, unsigned *err) {
  __shared__ int s_base[WM_NSP_MAX * 512];  // [isp][kind][256]: indexed by the top 9 bits of a tag
  const int tid = threadIdx.x, tile = blockIdx.x;
  const int li0 = (tile % P.ntx) * TX, lj0 = (tile / P.ntx) * TY;
  const int tw = min(TX, P.nx - li0), th = min(TY, P.nyl - lj0);
  for (int e = tid; e < P.nsp * 2 * WIN; e += blockDim.x) {
    const int isp = e / (2 * WIN), r = e - isp * (2 * WIN), kind = r / WIN, w = r - kind * WIN;
    const int cell = window_cell(P, li0, lj0, w);
    s_base[isp * 512 + kind * 256 + w] =
        (cell < 0) ? -1 : tilebase[(size_t)tile * P.nsp * (2 * WIN) + e] + cstart_new[(size_t)isp * (P.ncell + 1) + cell];
  }
  __syncthreads();
  for (int isp = 0; isp < P.nsp; isp++) {
    const size_t so = (size_t)isp * P.cap;
    for (int cy = 0; cy < th; cy++) {
      const int c0 = (lj0 + cy) * P.nx + li0;
      const int beg = cstart_old[(size_t)isp * (P.ncell + 1) + c0];
      const int end = cstart_old[(size_t)isp * (P.ncell + 1) + c0 + tw];
      for (int p = beg + tid; p < end; p += blockDim.x) {  // capacity span of the tile row: skip gaps
        if (!slot_live(keyx[so + p])) continue;  // gaps: only the sorted store marks them
        const uint32_t t = tag[so + p];
        if (t == TAG_DEAD) continue;
        const int d = s_base[isp * 512 + (t >> TAG_WSHIFT)] + (int)(t & TAG_RANK_MASK);
        if (d < 0 || d >= P.cap) {
          atomicOr(err, ERR_CAPACITY);
          continue;
        }
        dst.x[so + d] = src.x[so + p];
        dst.y[so + d] = src.y[so + p];
        dst.ux[so + d] = src.ux[so + p];
        dst.uy[so + d] = src.uy[so + p];
        dst.uz[so + d] = src.uz[so + p];
        dst.id[so + d] = src.id[so + p];
      }
    }
  }
}

// ---------------------------------------------------------------- prefix scan (sort.f90:64-69)
constexpr int SCAN_THREADS = 256;
constexpr int SCAN_ITEMS = 16;
constexpr int SCAN_TILE = SCAN_THREADS * SCAN_ITEMS;

__device__ __forceinline__ int block_exclusive_scan(int v, int *total) {
  // exclusive scan of one int per thread over a 256-thread block
  __shared__ int s_w[SCAN_THREADS / 32];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  int inc = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, inc, o);
    if (lane >= o) inc += t;
  }
  if (lane == 31) s_w[wid] = inc;
  __syncthreads();
  if (wid == 0) {
    int ws = (lane < SCAN_THREADS / 32) ? s_w[lane] : 0;
#pragma unroll
    for (int o = 1; o < SCAN_THREADS / 32; o <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, ws, o);
      if (lane >= o) ws += t;
    }
    if (lane < SCAN_THREADS / 32) s_w[lane] = ws;
  }
  __syncthreads();
  const int wbase = wid ? s_w[wid - 1] : 0;
  *total = s_w[SCAN_THREADS / 32 - 1];
  const int r = wbase + inc - v;
  __syncthreads();
  return r;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_partial(const int *__restrict__ in, int *__restrict__ bsum, int n, float sl, int floor_n) {
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++)
    if (base + k < n) s += cell_capacity(sl > 0.f ? max(in[base + k], floor_n) : in[base + k], sl);
  int tot;
  block_exclusive_scan(s, &tot);
  if (threadIdx.x == 0) bsum[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_bsums(int *bsum, int nb) {
  // single block; nb <= SCAN_TILE
  const int base = threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = (base + k < nb) ? bsum[base + k] : 0;
    s += v[k];
  }
  int tot;
  int ex = block_exclusive_scan(s, &tot);
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < nb) bsum[base + k] = ex;
    ex += v[k];
  }
}

__global__ void __launch_bounds__(SCAN_THREADS) k_scan_final(const int *__restrict__ in, const int *__restrict__ bsum,
                                                             int *__restrict__ out, int n, float sl, int floor_n) {
  const int base = blockIdx.x * SCAN_TILE + threadIdx.x * SCAN_ITEMS;
  int v[SCAN_ITEMS], s = 0;
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    v[k] = (base + k < n) ? cell_capacity(sl > 0.f ? max(in[base + k], floor_n) : in[base + k], sl) : 0;
    s += v[k];
  }
  int tot;
  int ex = block_exclusive_scan(s, &tot) + bsum[blockIdx.x];
#pragma unroll
  for (int k = 0; k < SCAN_ITEMS; k++) {
    if (base + k < n) out[base + k] = ex;
    ex += v[k];
    if (base + k == n - 1) out[n] = ex;  // grand total -> cstart[ncell]
  }
}

// ---------------------------------------------------------------- records <-> SoA
// incoming records (uploads, migrated particles): rank inside the destination cell by global atomics
// spv != nullptr: the species of record s is spv[s] (the overflow list mixes species)
__global__ void k_incoming_tag(const DevParams P, const double *__restrict__ rec, int n, int isp, const int *__restrict__ spv,
                               int *gcnt, int *__restrict__ rank, unsigned *err) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  if (spv) isp = spv[s];
  const int li = __double2int_rz(rec[(size_t)s * 6 + 0]) - P.nxgs;
  const int lj = __double2int_rz(rec[(size_t)s * 6 + 1]) - P.nys;
  if (li < 0 || li >= P.nx || lj < 0 || lj >= P.nyl) {
    atomicOr(err, ERR_BAD_CELL);
    rank[s] = -1;
    return;
  }
  rank[s] = atomicAdd(&gcnt[(size_t)isp * P.ncell + lj * P.nx + li], 1);
}

__global__ void k_incoming_scatter(const DevParams P, const double *__restrict__ rec, int n, int isp,
                                   const int *__restrict__ spv, const int *__restrict__ cstart_new,
                                   const int *__restrict__ rank, const PartSoA dst, unsigned *err) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  if (rank[s] < 0) return;
  if (spv) isp = spv[s];
  const double *r = rec + (size_t)s * 6;
  const int cell = (__double2int_rz(r[1]) - P.nys) * P.nx + (__double2int_rz(r[0]) - P.nxgs);
  const long long d = (long long)cstart_new[(size_t)isp * (P.ncell + 1) + cell] + rank[s];
  if (d >= P.cap) {
    atomicOr(err, ERR_CAPACITY);
    return;
  }
  const size_t o = (size_t)isp * P.cap + d;
  dst.x[o] = r[0];
  dst.y[o] = r[1];
  dst.ux[o] = r[2];
  dst.uy[o] = r[3];
  dst.uz[o] = r[4];
  dst.id[o] = __double_as_longlong(r[5]);
}

// slot-identical transposes between the AoS record order of the host arrays and the SoA store
__global__ void k_aos2soa(const double *__restrict__ rec, long long n, size_t so, const PartSoA dst) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const double *r = rec + s * 6;
  dst.x[so + s] = r[0];
  dst.y[so + s] = r[1];
  dst.ux[so + s] = r[2];
  dst.uy[so + s] = r[3];
  dst.uz[so + s] = r[4];
  dst.id[so + s] = __double_as_longlong(r[5]);
}

__global__ void k_soa2aos(const PartSoA src, size_t so, long long n, double *__restrict__ rec) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  double *r = rec + s * 6;
  r[0] = src.x[so + s];
  r[1] = src.y[so + s];
  r[2] = src.ux[so + s];
  r[3] = src.uy[so + s];
  r[4] = src.uz[so + s];
  r[5] = __longlong_as_double(src.id[so + s]);
}

// stand-alone x wrap (stage mode)                                  boundary_periodic.f90:61-96
__global__ void k_bcx(const DevParams P, const PartSoA g, const int *__restrict__ cstart, const bool injection) {
  const PView<double> x = g.x;
  for (int isp = 0; isp < P.nsp; isp++) {
    const int n = cstart[(size_t)isp * (P.ncell + 1) + P.ncell];
    for (int s = blockIdx.x * blockDim.x + threadIdx.x; s < n; s += gridDim.x * blockDim.x) {
      double v = x[(size_t)isp * P.cap + s];
      if (!slot_live(v)) continue;
      const int ipos = __double2int_rz(v);
      if (P.bc != WM_BC_PERIODIC) {  // reflecting walls   proj/reconnection/boundary_reconnection.f90:82-92
        const size_t o = (size_t)isp * P.cap + s;
        if (injection) {  // bc__injection   proj/shock/boundary_shock.f90:280-291
          if (v < P.xwlo) {
            x[o] = P.xw2lo - v;
            g.ux[o] = -g.ux[o];
            g.uy[o] = -g.uy[o];
            g.uz[o] = -g.uz[o];
          } else if (v > P.xwhi) {
            x[o] = P.xw2hi - v;
            g.ux[o] = P.u0x2 - g.ux[o];
            g.uy[o] = -g.uy[o];
            g.uz[o] = -g.uz[o];
          }
          continue;
        }
        // bc__particle_x of the wall modules: walls at nxs+1 and nxe-1 (in WM_BC_SHOCK xwhi/xw2hi hold xend)
        const double whi = (P.bc == WM_BC_SHOCK) ? (double)(P.nxgs + P.nx - 2) * P.delx : P.xwhi;
        const double w2hi = (P.bc == WM_BC_SHOCK) ? 2. * (P.nxgs + P.nx - 2) * P.delx : P.xw2hi;
        if (v < P.xwlo || v >= whi) {
          x[o] = (v < P.xwlo ? P.xw2lo : w2hi) - v;
          g.ux[o] = -g.ux[o];
          g.uy[o] = -g.uy[o];
          g.uz[o] = -g.uz[o];
        }
        continue;
      }
      if (ipos < P.nxgs) {
        x[(size_t)isp * P.cap + s] = __dadd_rd(v, P.xlen);
      } else if (ipos >= P.nxgs + P.nx) {
        x[(size_t)isp * P.cap + s] = __dadd_rd(v, -P.xlen);
      }
    }
  }
}

// ---------------------------------------------------------------- synthetic IC (app.f90:380-474)
__device__ __forceinline__ uint64_t splitmix(uint64_t z) {
  z += 0x9E3779B97F4A7C15ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}
__device__ __forceinline__ double rng_uniform(uint64_t seed, int isp, uint64_t gid, int stream) {
  uint64_t k = splitmix(seed ^ (0xD1B54A32D192ED03ULL * (uint64_t)(isp + 1)));
  k = splitmix(k + gid);
  k = splitmix(k + 0x8CB92BA72F3D8DD7ULL * (uint64_t)(stream + 1));
  return ((double)(k >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}

// The x array of `dst` must have been filled with the dead pattern: only live slots are written.
__global__ void k_ic_weibel(const DevParams P, const PartSoA dst, int *cstart, int *cnt, uint64_t seed, int n0, double vti,
                            double vte, double t_ani, float sl) {
  const int capc = cell_capacity(n0, sl);
  const long long npr = (long long)n0 * P.nx;
  const long long ntot = npr * P.nyl;
  const double PI = 3.14159265358979323846;
  for (long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x; s < ntot; s += (long long)gridDim.x * blockDim.x) {
    const int lj = (int)(s / npr);
    const long long ii = s - (long long)lj * npr + 1;  // 1-based slot in the row
    const int gj = P.nys + lj;
    const uint64_t pid = (uint64_t)(gj - P.nygs) * (uint64_t)npr + (uint64_t)ii;
    const double x = (P.nxgs + P.nx * ((double)ii - 0.5) / (double)npr) * P.delx;  // app.f90:408
    const double y = (gj + rng_uniform(seed, 0, pid, 0)) * P.delx;                   // app.f90:409
    for (int isp = 0; isp < P.nsp; isp++) {
      const double u1 = rng_uniform(seed, isp + 1, pid, 1), u2 = rng_uniform(seed, isp + 1, pid, 2);
      const double u3 = rng_uniform(seed, isp + 1, pid, 3), u4 = rng_uniform(seed, isp + 1, pid, 4);
      const double rr1 = sqrt(-2 * log(1 - u1) + 1.0e-30);
      const double rr2 = sqrt(-2 * log(1 - u3) + 1.0e-30);
      const double sd = (isp & 1) ? vte : vti;
      // evenly spaced x: slot ii of the row lies in cell (ii-1)/n0   (app.f90:315-328,408)
      const long long cell = (long long)lj * P.nx + (ii - 1) / n0;
      const size_t o = (size_t)isp * P.cap + (size_t)(cell * capc + (ii - 1) % n0);
      dst.x[o] = x;
      dst.y[o] = y;
      dst.ux[o] = sd * (rr1 * sin(2 * PI * u2));
      dst.uy[o] = sd * (rr1 * cos(2 * PI * u2));
      dst.uz[o] = t_ani * sd * (rr2 * sin(2 * PI * u4));
      dst.id[o] = -(long long)pid;  // app.f90:468
    }
  }
  // analytic cell offsets for the evenly spaced load (app.f90:315-328)
  for (long long c = (long long)blockIdx.x * blockDim.x + threadIdx.x; c <= P.ncell; c += (long long)gridDim.x * blockDim.x)
    for (int isp = 0; isp < P.nsp; isp++) {
      cstart[(size_t)isp * (P.ncell + 1) + c] = (int)(c * capc);
      if (c < P.ncell) cnt[(size_t)isp * P.ncell + c] = n0;
    }
}

// kinetic energy per species (app.f90:507-519); one partial per block, summed on the host in order
__global__ void __launch_bounds__(256) k_kinetic(const DevParams P, const PartSoA src, const int *__restrict__ cstart,
                                                 int isp, double *__restrict__ partial) {
  __shared__ double s_w[8];
  const int n = cstart[(size_t)isp * (P.ncell + 1) + P.ncell];
  const size_t so = (size_t)isp * P.cap;
  double s = 0.0;
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
    if (!slot_live(src.x[so + p])) continue;
    const double u1 = src.ux[so + p], u2 = src.uy[so + p], u3 = src.uz[so + p];
    const double uu = u1 * u1 + u2 * u2 + u3 * u3;
    s += P.r[isp] * (sqrt(1 + uu / P.cc) - 1);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = s;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int k = 0; k < 8; k++) t += s_w[k];
    partial[blockIdx.x] = t;
  }
}

// moments with bilinear weights (mom_calc.f90:167-249): mom (7, nx+2, nyl+2, nsp) = (7, nxgs-1:nxge+1, nys-1:nye+1, nsp), RED.ADD.F64
__global__ void k_moments(const DevParams P, const PartSoA src, const PView<double> keyx,
                          const int *__restrict__ cstart, double *mom) {
  for (int isp = 0; isp < P.nsp; isp++) {
    const int n = cstart[(size_t)isp * (P.ncell + 1) + P.ncell];
    const size_t so = (size_t)isp * P.cap;
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
      if (!slot_live(keyx[so + p])) continue;  // gaps of the idle store hold stale data
      const double x = src.x[so + p], y = src.y[so + p];
      const double u1 = src.ux[so + p], u2 = src.uy[so + p], u3 = src.uz[so + p];
      const int ih = __double2int_rz(x - 0.5), jh = __double2int_rz(y - 0.5);
      const double dx = x - 0.5 - ih, dxm = 1. - dx, dy = y - 0.5 - jh, dym = 1. - dy;
      const double gam = 1. / sqrt(1.0 + (u1 * u1 + u2 * u2 + u3 * u3) / P.cc);
      const double val[7] = {1.0, u1 * gam, u2 * gam, u3 * gam, u1 * u1 * gam, u2 * u2 * gam, u3 * u3 * gam};
      const int mi = ih - (P.nxgs - 1), mj = jh - (P.nys - 1);
      if (mi < 0 || mi + 1 >= P.nx + 2 || mj < 0 || mj + 1 >= P.nyl + 2) continue;
      double *m00 = mom + (((size_t)isp * (P.nyl + 2) + mj) * (P.nx + 2) + mi) * 7;
      double *m01 = m00 + 7, *m10 = m00 + (size_t)(P.nx + 2) * 7, *m11 = m10 + 7;
#pragma unroll
      for (int m = 0; m < 7; m++) {
        atomicAdd(&m00[m], val[m] * dxm * dym);
        atomicAdd(&m01[m], val[m] * dx * dym);
        atomicAdd(&m10[m], val[m] * dxm * dy);
        atomicAdd(&m11[m], val[m] * dx * dy);
      }
    }
  }
}


// Charge density with the second-order shape function, periodic in x, for the discrete Gauss law
// div E = 4 pi rho (field.f90:159; the weights are those of particle.f90:97-105):
//   rho[0][j][i] += q S2(x - i - 1/2) S2(y - j - 1/2),  rho[1] the same with |q| (scale of the residual)
// Planes of (nyl + 2) rows: row lj + 1 holds local row lj; rows 0 and nyl + 1 collect what belongs to the ring neighbours
// (folded by the host with one exchange per direction); with one rank the rows wrap onto the slab itself.
__global__ void k_charge_density(const DevParams P, const PartSoA src, const int *__restrict__ cstart, double *rho) {
  const size_t plane = (size_t)P.nx * (P.nyl + 2);
  for (int isp = 0; isp < P.nsp; isp++) {
    const int n = cstart[(size_t)isp * (P.ncell + 1) + P.ncell];
    const size_t so = (size_t)isp * P.cap;
    const double q = P.q[isp], qa = fabs(P.q[isp]);
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < n; p += gridDim.x * blockDim.x) {
      const double x = src.x[so + p];
      if (!slot_live(x)) continue;
      const double y = src.y[so + p];
      const int ic = __double2int_rz(x), jc = __double2int_rz(y);
      double sx[3], sy[3];
      double dh = x - 0.5 - ic;
      sx[0] = 0.5 * (0.5 - dh) * (0.5 - dh); sx[1] = 0.75 - dh * dh; sx[2] = 0.5 * (0.5 + dh) * (0.5 + dh);
      dh = y - 0.5 - jc;
      sy[0] = 0.5 * (0.5 - dh) * (0.5 - dh); sy[1] = 0.75 - dh * dh; sy[2] = 0.5 * (0.5 + dh) * (0.5 + dh);
#pragma unroll
      for (int b = -1; b <= 1; b++) {
        int lj = jc + b - P.nys;
        if (P.nsize == 1) lj = (lj < 0) ? lj + P.nyl : (lj >= P.nyl ? lj - P.nyl : lj);
#pragma unroll
        for (int a = -1; a <= 1; a++) {
          int li = ic + a - P.nxgs;
          li = (li < 0) ? li + P.nx : (li >= P.nx ? li - P.nx : li);
          const double w = sx[a + 1] * sy[b + 1];
          atomicAdd(&rho[(size_t)(lj + 1) * P.nx + li], q * w);
          atomicAdd(&rho[plane + (size_t)(lj + 1) * P.nx + li], qa * w);
        }
      }
    }
  }
}

// max |div E - 4 pi rho| and max 4 pi rho_abs over the cells; out[0], out[1] hold non-negative doubles whose bit
// patterns order like the values, so atomicMax on the 64-bit integers is exact
__global__ void k_gauss_residual(const DevParams P, const double *__restrict__ uf, const double *__restrict__ rho,
                                 unsigned long long *out) {
  const double PI4 = 4.0 * 3.14159265358979323846;
  const size_t plane = (size_t)P.nx * (P.nyl + 2);
  rho += P.nx;  // row lj + 1
  double res = 0.0, sc = 0.0;
  for (int t = blockIdx.x * blockDim.x + threadIdx.x; t < P.nx * P.nyl; t += gridDim.x * blockDim.x) {
    const int lj = t / P.nx, li = t - lj * P.nx;
    const double *c = uf + ((size_t)(lj + 2) * P.pitch + (li + 2)) * 6;
    const double *xp = c + 6, *yp = c + (size_t)P.pitch * 6;
    const double dive = (xp[3] - c[3] + yp[4] - c[4]) / P.delx;
    res = fmax(res, fabs(dive - PI4 * rho[t]));
    sc = fmax(sc, PI4 * rho[plane + t]);
  }
  for (int o = 16; o > 0; o >>= 1) {
    res = fmax(res, __shfl_xor_sync(0xffffffffu, res, o));
    sc = fmax(sc, __shfl_xor_sync(0xffffffffu, sc, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(&out[0], (unsigned long long)__double_as_longlong(res));
    atomicMax(&out[1], (unsigned long long)__double_as_longlong(sc));
  }
}
void launch_charge_density(const DevParams &P, const PartSoA &src, const int *cstart, double *rho, cudaStream_t st) {
  cudaMemsetAsync(rho, 0, (size_t)2 * P.nx * (P.nyl + 2) * sizeof(double), st);
  k_charge_density<<<148 * 8, 256, 0, st>>>(P, src, cstart, rho);
}
void launch_gauss(const DevParams &P, const double *uf, const double *rho, unsigned long long *out, cudaStream_t st) {
  cudaMemsetAsync(out, 0, 2 * sizeof(unsigned long long), st);
  k_gauss_residual<<<148 * 4, 256, 0, st>>>(P, uf, rho, out);
}

// ---------------------------------------------------------------- layout changes (order inside a cell is kept)
__device__ __forceinline__ int cell_of(const DevParams &P, double x, double y) {
  return (__double2int_rz(y) - P.nys) * P.nx + (__double2int_rz(x) - P.nxgs);
}

// tight, cell-sorted AoS records (a host upload) -> segments with slack
__global__ void k_relayout_from_aos(const DevParams P, const double *__restrict__ rec, long long n,
                                    const int *__restrict__ tight, const int *__restrict__ cstart_dst, const PartSoA dst,
                                    size_t so, unsigned *err) {
  const long long s = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const double *r = rec + s * 6;
  const int li = __double2int_rz(r[0]) - P.nxgs, lj = __double2int_rz(r[1]) - P.nys;
  if (li < 0 || li >= P.nx || lj < 0 || lj >= P.nyl) {
    atomicOr(err, ERR_BAD_CELL);
    return;
  }
  const int cell = lj * P.nx + li;
  const int k = (int)s - tight[cell];
  if (k < 0 || k >= tight[cell + 1] - tight[cell]) {  // cumcnt inconsistent with the positions
    atomicOr(err, ERR_BAD_CELL);
    return;
  }
  if ((long long)cstart_dst[cell] + k >= P.cap) {
    atomicOr(err, ERR_CAPACITY);
    return;
  }
  const size_t o = so + (size_t)cstart_dst[cell] + k;
  dst.x[o] = r[0];
  dst.y[o] = r[1];
  dst.ux[o] = r[2];
  dst.uy[o] = r[3];
  dst.uz[o] = r[4];
  dst.id[o] = __double_as_longlong(r[5]);
}

// segments with slack -> tight AoS records.  `key` (the sorted store) gives liveness and the cell,
// `val` the values (the same store, or the post-push store for wm_download_gp).
__global__ void k_relayout_to_aos(const DevParams P, const PartSoA key, const PartSoA val, size_t so,
                                  const int *__restrict__ cstart_src, const int *__restrict__ tight,
                                  double *__restrict__ rec) {
  const int span = cstart_src[P.ncell];
  for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < span; p += gridDim.x * blockDim.x) {
    const double x = key.x[so + p];
    if (!slot_live(x)) continue;
    const int cell = cell_of(P, x, key.y[so + p]);
    double *r = rec + ((size_t)tight[cell] + (p - cstart_src[cell])) * 6;
    r[0] = val.x[so + p];
    r[1] = val.y[so + p];
    r[2] = val.ux[so + p];
    r[3] = val.uy[so + p];
    r[4] = val.uz[so + p];
    r[5] = __longlong_as_double(val.id[so + p]);
  }
}

// segments -> segments with different offsets (layout rebuild), all species in one launch.  Count-driven: a warp copies the
// cnt live records at the front of a segment (the in-place sort keeps them compact), so the cost follows the particle
// number, not the span of the store (the shock run holds 35 M particles per species in 537 M slots).
__global__ void k_relayout_soa(const DevParams P, const PartSoA src, const int *__restrict__ cstart_src, const int *__restrict__ cnt_src,
                               const PartSoA dst, const int *__restrict__ cstart_dst, unsigned *err) {
  const long long n = (long long)P.nsp * P.ncell;
  const int lane = threadIdx.x & 31;
  for (long long wk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wk < n; wk += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int isp = (int)(wk / P.ncell), cell = (int)(wk - (long long)isp * P.ncell);
    const int m = cnt_src[wk];
    if (m <= 0) continue;
    const size_t so = (size_t)isp * P.cap;
    const int s0 = cstart_src[(size_t)isp * (P.ncell + 1) + cell];
    const int d0 = cstart_dst[(size_t)isp * (P.ncell + 1) + cell], d1 = cstart_dst[(size_t)isp * (P.ncell + 1) + cell + 1];
    if ((long long)d0 + m > P.cap || d0 + m > d1) {
      if (lane == 0) atomicOr(err, ERR_CAPACITY);
      continue;
    }
    for (int k = lane; k < m; k += 32) {
      const double2 *w = src.word(so + (size_t)s0 + k);
      double2 *o = dst.word(so + (size_t)d0 + k);
      o[0] = w[0];
      o[8] = w[8];
      o[16] = w[16];
    }
  }
}

// in-place sort: records arriving from a neighbour rank (or from the overflow list after a rebuild
// is NOT needed: see wm_api.cu) are appended at the tail of their cell's segment
// n_dev != nullptr: the number of records is *n_dev (a count that arrived with the records); more than n is an error (the
// message was sized from the previous step's count and got truncated)
__global__ void k_incoming_append(const DevParams P, const double *__restrict__ rec, int n, const int *__restrict__ n_dev, int isp,
                                  const int *__restrict__ cstart, int *cnt_tail, const PartSoA dst, double *ovf, int *ovfsp,
                                  int *ovfcnt, int ovfcap, unsigned *err, const int *__restrict__ cntb) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (n_dev) {
    const int m = *n_dev;
    if (m > n) {
      if (s == 0) atomicOr(err, ERR_SENDBUF);
    } else {
      n = m;
    }
  }
  if (s >= n) return;
  const double *r = rec + (size_t)s * 6;
  const int li = __double2int_rz(r[0]) - P.nxgs, lj = __double2int_rz(r[1]) - P.nys;
  if (li < 0 || li >= P.nx || lj < 0 || lj >= P.nyl) {
    atomicOr(err, ERR_BAD_CELL);
    return;
  }
  const int cell = lj * P.nx + li;
  const int *cs = cstart + (size_t)isp * (P.ncell + 1);
  const int pos = atomicAdd(&cnt_tail[(size_t)isp * P.ncell + cell], 1);
  // (cntb: the back of the segment holds the arrivals of the last fused pass, k_fused_dp)
  if (pos < cs[cell + 1] - cs[cell] - (cntb ? cntb[(size_t)isp * P.ncell + cell] : 0)) {
    const size_t o = (size_t)isp * P.cap + cs[cell] + pos;
    dst.x[o] = r[0];
    dst.y[o] = r[1];
    dst.ux[o] = r[2];
    dst.uy[o] = r[3];
    dst.uz[o] = r[4];
    dst.id[o] = __double_as_longlong(r[5]);
  } else {
    const int k = atomicAdd(ovfcnt, 1);
    if (k < ovfcap) {
      for (int e = 0; e < 6; e++) ovf[(size_t)k * 6 + e] = r[e];
      ovfsp[k] = isp;
    } else {
      atomicOr(err, ERR_OVERFLOW);
    }
  }
}

// in-place sort, second half (the scatter of sort.f90:71-75 reduced to the particles that changed
// cell).  k_fused<INPLACE> left, per tile, the number of arrivals it sends to every cell of its window
// (tilebase[tile][isp][0][w]), the number of records it staged per quad (tilebase[tile][isp][1][q]),
// the staged records themselves in the idle store and their tags, and cnt_new[cell] = stayers.
// One CTA per tile reserves room at the tail of each destination segment and appends its records.
constexpr int PL_THREADS = 256;
constexpr int PL_NQ = (TX / 4) * TY;
constexpr int PL_MAX = 4096;  // records placed per round
__global__ void __launch_bounds__(PL_THREADS) k_place(const DevParams P, const double2 *__restrict__ stage,
                                                      const uint32_t *__restrict__ tag, const PartSoA dst,
                                                      const int *__restrict__ cstart, int *cnt_new,
                                                      const int *__restrict__ tilebase, double *ovf, int *ovfsp, int *ovfcnt,
                                                      int ovfcap, unsigned *err) {
  __shared__ int s_base[WM_NSP_MAX * WIN], s_end[WM_NSP_MAX * WIN], s_pref[WM_NSP_MAX * WIN + 1];
  __shared__ int s_off[WM_NSP_MAX * PL_NQ + 1];
  __shared__ long long s_rec0[WM_NSP_MAX * PL_NQ];
  __shared__ int s_inv[PL_MAX];
  const int tid = threadIdx.x, tile = blockIdx.x;
  const int li0 = (tile % P.ntx) * TX, lj0 = (tile / P.ntx) * TY;
  const int tw = min(TX, P.nx - li0), th = min(TY, P.nyl - lj0);
  const int *tb = tilebase + (size_t)tile * P.nsp * (2 * WIN);
  const int nreg = P.nsp * PL_NQ, nwin = P.nsp * WIN;
  // room at the tail of every destination segment; arrival counts and staged-record counts to smem
  for (int e = tid; e < nwin; e += PL_THREADS) {
    const int isp = e / WIN, w = e - isp * WIN;
    const int n = tb[isp * (2 * WIN) + w];
    int base = -1, end = 0;
    if (n > 0) {
      const int cell = window_cell(P, li0, lj0, w);
      if (cell < 0) {
        atomicOr(err, ERR_MOVED_TOO_FAR);
      } else {
        const int *cs = cstart + (size_t)isp * (P.ncell + 1);
        base = cs[cell] + atomicAdd(&cnt_new[(size_t)isp * P.ncell + cell], n);
        end = cs[cell + 1];
      }
    }
    s_base[e] = base;
    s_end[e] = end;
    s_pref[e + 1] = n;
  }
  for (int r = tid; r < nreg; r += PL_THREADS) {
    const int isp = r / PL_NQ, q = r - isp * PL_NQ;
    const int cy = q / (TX / 4), cx0 = (q - cy * (TX / 4)) * 4;
    long long rec0 = 0;
    int cap = 0;
    if (cy < th && cx0 < tw) {
      const int c0 = (lj0 + cy) * P.nx + li0 + cx0;
      const int *cs = cstart + (size_t)isp * (P.ncell + 1);
      stage_region(so_slots(P, isp) + cs[c0], so_slots(P, isp) + cs[min(c0 + 4, (lj0 + cy + 1) * P.nx)], &rec0, &cap);
    }
    s_rec0[r] = rec0;
    s_off[r + 1] = min(tb[isp * (2 * WIN) + WIN + q], cap);
  }
  __syncthreads();
  if (tid == 0) {
    s_pref[0] = 0;
    for (int e = 0; e < nwin; e++) s_pref[e + 1] += s_pref[e];
  }
  if (tid == 32) {
    s_off[0] = 0;
    for (int r = 0; r < nreg; r++) s_off[r + 1] += s_off[r];
  }
  __syncthreads();
  const int total = s_pref[nwin], nstaged = s_off[nreg];
  for (int j0 = 0; j0 < total; j0 += PL_MAX) {
    // where does every staged record go in destination order?  (window cell, rank) -> position
    for (int k = tid; k < nstaged; k += PL_THREADS) {
      int lo = 0, hi = nreg;  // largest r with s_off[r] <= k
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_off[mid] <= k) lo = mid; else hi = mid;
      }
      const long long ri = s_rec0[lo] + (k - s_off[lo]);
      const uint32_t t = tag[ri];
      if (t == TAG_DEAD) continue;  // left the slab: already in the send buffer
      const int e = (lo / PL_NQ) * WIN + (int)((t >> TAG_WSHIFT) & 0xff);
      if (s_base[e] < 0) continue;  // placed by k_fused_sm itself (new cell inside the tile)
      const int j = s_pref[e] + (int)(t & TAG_RANK_MASK) - j0;
      if (j >= 0 && j < PL_MAX) s_inv[j] = (int)(ri - s_rec0[0]);
    }
    __syncthreads();
    // consecutive threads write consecutive slots of the same destination segment
    const int nj = min(PL_MAX, total - j0);
    for (int j = tid; j < nj; j += PL_THREADS) {
      const long long ri = s_rec0[0] + s_inv[j];
      const double2 r0 = stage[ri * 3], r1 = stage[ri * 3 + 1], r2 = stage[ri * 3 + 2];
      const uint32_t t = tag[ri];
      int lo = 0, hi = nwin;  // largest e with s_pref[e] <= j0 + j
      while (hi - lo > 1) {
        const int mid = (lo + hi) >> 1;
        if (s_pref[mid] <= j0 + j) lo = mid; else hi = mid;
      }
      const int e = lo, isp = e / WIN;
      if (s_base[e] < 0) continue;
      const int d = s_base[e] + (int)(t & TAG_RANK_MASK);
      if (d < s_end[e]) {
        double2 *o = dst.word((size_t)isp * P.cap + d);  // three 16-byte words of the record, one per 128-byte row
        o[0] = r0;
        o[8] = r1;
        o[16] = r2;
      } else {  // segment full: park the record; the host rebuilds the layout after this step
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
    __syncthreads();
  }
}

// k_place for what k_fused_sm<TAIL> leaves over: the tile's own cells were served by the CTA itself, only the rim of its
// window (cells of other tiles, 12 % of the cell changers) still has arrivals.  One pass in staging order, a warp per
// staging region (quad, species): the tags of four regions x 64 records in flight per warp, then the records of the rim
// arrivals among them (compacted over the warp); slot = reserved base + rank, no destination-order map (the 16-byte pieces of a line merge in L2).
constexpr int PR_THREADS = 256;
constexpr int PR_RB = 4;
__global__ void __launch_bounds__(PR_THREADS) k_place_rim(const DevParams P, const double2 *__restrict__ stage,
                                                          const uint32_t *__restrict__ tag, const PartSoA dst,
                                                          const int *__restrict__ cstart, int *cnt_new,
                                                          const int *__restrict__ tilebase, double *ovf, int *ovfsp, int *ovfcnt,
                                                          int ovfcap, unsigned *err, int tile0) {
  __shared__ int s_base[WM_NSP_MAX * WIN], s_end[WM_NSP_MAX * WIN];
  __shared__ int s_roff[WM_NSP_MAX * PL_NQ], s_rcnt[WM_NSP_MAX * PL_NQ];
  __shared__ int2 s_list[(PR_THREADS / 32) * 2 * PR_RB * 32];
  static_assert(PL_NQ % PR_RB == 0, "regions in flight never straddle the species");
  const int tid = threadIdx.x, tile = blockIdx.x + tile0, wid = tid >> 5, lane = tid & 31;
  const int li0 = (tile % P.ntx) * TX, lj0 = (tile / P.ntx) * TY;
  const int tw = min(TX, P.nx - li0), th = min(TY, P.nyl - lj0);
  const int *tb = tilebase + (size_t)tile * P.nsp * (2 * WIN);
  const int nreg = P.nsp * PL_NQ, nwin = P.nsp * WIN;
  for (int e = tid; e < nwin; e += PR_THREADS) {
    const int isp = e / WIN, w = e - isp * WIN;
    const int n = tb[isp * (2 * WIN) + w];
    int base = -1, end = 0;
    if (n > 0) {
      const int cell = window_cell(P, li0, lj0, w);
      if (cell < 0) {
        atomicOr(err, ERR_MOVED_TOO_FAR);
      } else {
        const int *cs = cstart + (size_t)isp * (P.ncell + 1);
        base = cs[cell] + atomicAdd(&cnt_new[(size_t)isp * P.ncell + cell], n);
        end = cs[cell + 1];
      }
    }
    s_base[e] = base;
    s_end[e] = end;
  }
  for (int r = tid; r < nreg; r += PR_THREADS) {
    const int isp = r / PL_NQ, q = r - isp * PL_NQ;
    const int cy = q / (TX / 4), cx0 = (q - cy * (TX / 4)) * 4;
    long long rec0 = so_slots(P, isp);
    int cap = 0;
    if (cy < th && cx0 < tw) {
      const int c0 = (lj0 + cy) * P.nx + li0 + cx0;
      const int *cs = cstart + (size_t)isp * (P.ncell + 1);
      stage_region(so_slots(P, isp) + cs[c0], so_slots(P, isp) + cs[min(c0 + 4, (lj0 + cy + 1) * P.nx)], &rec0, &cap);
    }
    s_roff[r] = (int)(rec0 - so_slots(P, isp));
    s_rcnt[r] = min(tb[isp * (2 * WIN) + WIN + q], cap);
  }
  __syncthreads();
  int2 *const mylist = s_list + wid * (2 * PR_RB * 32);
  const unsigned lanelt = (1u << lane) - 1u;
#pragma unroll 1
  for (int rb = wid * PR_RB; rb < nreg; rb += (PR_THREADS / 32) * PR_RB) {
    int roff[PR_RB], rcnt[PR_RB], nmaxr = 0;
#pragma unroll
    for (int v = 0; v < PR_RB; v++) {
      roff[v] = s_roff[rb + v];
      rcnt[v] = s_rcnt[rb + v];
      nmaxr = max(nmaxr, rcnt[v]);
    }
    const int isp = rb / PL_NQ;
    const size_t so = (size_t)isp * P.cap;
#pragma unroll 1
    for (int kb = 0; kb < nmaxr; kb += 64) {
      uint32_t tg[2 * PR_RB];
#pragma unroll
      for (int u = 0; u < 2 * PR_RB; u++) {
        const int k = kb + lane + 32 * (u & 1);
        tg[u] = TAG_DEAD;
        if (k < rcnt[u >> 1]) tg[u] = tag[so + (size_t)(roff[u >> 1] + k)];
      }
      // the rim arrivals among them (about one in eight): compacted into the warp's list (source record, destination)
      int nact = 0;
#pragma unroll
      for (int u = 0; u < 2 * PR_RB; u++) {
        const uint32_t t = tg[u];
        int dd = -1;  // destination slot, -1 = nothing to do, -2 = segment full
        if (t != TAG_DEAD) {  // (dead: left the slab, already in the send buffer)
          const int e = isp * WIN + (int)((t >> TAG_WSHIFT) & 0xff);
          const int bs = s_base[e];
          if (bs >= 0) {  // (< 0: placed by k_fused_sm itself, new cell inside the tile)
            dd = bs + (int)(t & TAG_RANK_MASK);
            if (dd >= s_end[e]) dd = -2;
          }
        }
        const unsigned bal = __ballot_sync(0xffffffffu, dd != -1);
        if (dd != -1) mylist[nact + __popc(bal & lanelt)] = make_int2(roff[u >> 1] + kb + lane + 32 * (u & 1), dd);
        nact += __popc(bal);
      }
      __syncwarp();
      for (int i = lane; i < nact; i += 32) {
        const int2 m = mylist[i];
        const double2 *r = stage + (so + (size_t)m.x) * 3;
        const double2 r0 = r[0], r1 = r[1], r2 = r[2];
        if (m.y >= 0) {
          double2 *o = dst.word(so + (size_t)m.y);  // three 16-byte words of the record, one per 128-byte row
          o[0] = r0;
          o[8] = r1;
          o[16] = r2;
        } else {  // segment full: park the record; the host rebuilds the layout after this step
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
      __syncwarp();
    }
  }
}

// in-place sort, last step: clamp the new counts to the segment capacity (the surplus is in the
// overflow list) and retire the slots between the new and the old count
__global__ void k_mark_dead(const DevParams P, const PView<double> x, const int *__restrict__ cstart, const int *__restrict__ cnt_old,
                            int *cnt_new) {
  const long long n = (long long)P.nsp * P.ncell;
  for (long long wk = (long long)blockIdx.x * blockDim.x + threadIdx.x; wk < n; wk += (long long)gridDim.x * blockDim.x) {
    const int isp = (int)(wk / P.ncell), cell = (int)(wk - (long long)isp * P.ncell);
    const int *cs = cstart + (size_t)isp * (P.ncell + 1);
    const int capc = cs[cell + 1] - cs[cell];
    int nn = cnt_new[wk];
    if (nn > capc) {
      nn = capc;
      cnt_new[wk] = nn;
    }
    const size_t xs = (size_t)isp * P.cap + cs[cell];
    for (int p = nn; p < cnt_old[wk]; p++) x[xs + p] = dead_x();
  }
}

// counts that ran past their segment's capacity (the surplus went to the overflow list): clamp
__global__ void k_clamp_counts(const DevParams P, const int *__restrict__ cstart, int *cnt, const int *__restrict__ cntb) {
  const long long n = (long long)P.nsp * P.ncell;
  for (long long wk = (long long)blockIdx.x * blockDim.x + threadIdx.x; wk < n; wk += (long long)gridDim.x * blockDim.x) {
    const int isp = (int)(wk / P.ncell), cell = (int)(wk - (long long)isp * P.ncell);
    const int *cs = cstart + (size_t)isp * (P.ncell + 1);
    const int full = cs[cell + 1] - cs[cell];
    const int capc = full - (cntb ? min(cntb[wk], full) : 0);  // (k_fused_dp: the back of the segment is taken)
    if (cnt[wk] > capc) cnt[wk] = capc;
  }
}
// layout rebuild: the slots behind the live particles of every segment become dead (instead of a memset of the whole store)
__global__ void k_mark_gaps(const DevParams P, PView<double> x, const int *__restrict__ cstart, const int *__restrict__ cnt) {
  const long long n = (long long)P.nsp * P.ncell;
  const int lane = threadIdx.x & 31;
  for (long long wk = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5; wk < n; wk += ((long long)gridDim.x * blockDim.x) >> 5) {
    const int isp = (int)(wk / P.ncell), cell = (int)(wk - (long long)isp * P.ncell);
    const int *cs = cstart + (size_t)isp * (P.ncell + 1);
    const size_t so = (size_t)isp * P.cap;
    for (int p = cs[cell] + cnt[wk] + lane; p < cs[cell + 1]; p += 32) x[so + p] = dead_x();
  }
}
void launch_mark_gaps(const DevParams &P, PView<double> x, const int *cstart, const int *cnt, cudaStream_t st) {
  k_mark_gaps<<<148 * 16, 256, 0, st>>>(P, x, cstart, cnt);
}
// out[c] = max of in over the cells c-r .. c+r of the same row: segments sized for the densest neighbour survive a moving
// density front (shock) for ~r / front speed steps instead of one
__global__ void k_nbr_max(const DevParams P, const int *__restrict__ in, int *__restrict__ out, int r) {
  for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < P.ncell; c += gridDim.x * blockDim.x) {
    const int lj = c / P.nx, li = c - lj * P.nx;
    int m = 0;
    for (int d = max(li - r, 0); d <= min(li + r, P.nx - 1); d++) m = max(m, in[lj * P.nx + d]);
    out[c] = m;
  }
}
void launch_nbr_max(const DevParams &P, const int *in, int *out, int r, cudaStream_t st) { k_nbr_max<<<148 * 8, 256, 0, st>>>(P, in, out, r); }
void launch_clamp_counts(const DevParams &P, const int *cstart, int *cnt, cudaStream_t st, const int *cntb) {
  k_clamp_counts<<<148 * 8, 256, 0, st>>>(P, cstart, cnt, cntb);
}

// ---------------------------------------------------------------- launch wrappers
template <int MODE>
static void launch_p1(const DevParams &P, const Pass1Args &a, cudaStream_t st) {
  k_pass1<MODE><<<P.ntx * P.nty, P1_THREADS, 0, st>>>(P, a);
}

void launch_pass1(int mode, const DevParams &P, const Pass1Args &a, cudaStream_t st) {
  switch (mode) {
    case M_PUSH: launch_p1<M_PUSH>(P, a, st); break;
    case M_PUSH | M_EXACT: launch_p1<M_PUSH | M_EXACT>(P, a, st); break;
    case M_PUSH | M_NOMOVE: launch_p1<M_PUSH | M_NOMOVE>(P, a, st); break;
    case M_PUSH | M_NOMOVE | M_EXACT: launch_p1<M_PUSH | M_NOMOVE | M_EXACT>(P, a, st); break;
    case M_DEPOSIT: launch_p1<M_DEPOSIT>(P, a, st); break;
    case M_BOUND: launch_p1<M_BOUND>(P, a, st); break;
    case M_PUSH | M_DEPOSIT | M_BOUND: launch_p1<M_PUSH | M_DEPOSIT | M_BOUND>(P, a, st); break;
    case M_PUSH | M_DEPOSIT | M_BOUND | M_EXACT: launch_p1<M_PUSH | M_DEPOSIT | M_BOUND | M_EXACT>(P, a, st); break;
    default: fprintf(stderr, "wumingpic2d: unsupported pass1 mode %d\n", mode); abort();
  }
}

void launch_pass2(const DevParams &P, const PartSoA &src, const PartSoA &dst, const int *cstart_old,
                  const int *cstart_new, const int *tilebase, const uint32_t *tag, PView<double> keyx, unsigned *err,
                  cudaStream_t st) {
  k_pass2<<<P.ntx * P.nty, 256, 0, st>>>(P, src, dst, cstart_old, cstart_new, tilebase, tag, keyx, err);
}

int scan_scratch_ints(int n) { return (n + SCAN_TILE - 1) / SCAN_TILE + 1; }

int launch_scan(const int *in, int *out, int *scratch, int n, float sl, cudaStream_t st, int floor_n) {
  const int nb = (n + SCAN_TILE - 1) / SCAN_TILE;
  if (nb > SCAN_TILE) return 1;
  k_scan_partial<<<nb, SCAN_THREADS, 0, st>>>(in, scratch, n, sl, floor_n);
  k_scan_bsums<<<1, SCAN_THREADS, 0, st>>>(scratch, nb);
  k_scan_final<<<nb, SCAN_THREADS, 0, st>>>(in, scratch, out, n, sl, floor_n);
  return 0;
}

void launch_incoming_tag(const DevParams &P, const double *rec, int n, int isp, const int *spv, int *gcnt, int *rank,
                         unsigned *err, cudaStream_t st) {
  if (n > 0) k_incoming_tag<<<(n + 255) / 256, 256, 0, st>>>(P, rec, n, isp, spv, gcnt, rank, err);
}
void launch_incoming_scatter(const DevParams &P, const double *rec, int n, int isp, const int *spv, const int *cstart_new,
                             const int *rank, const PartSoA &dst, unsigned *err, cudaStream_t st) {
  if (n > 0) k_incoming_scatter<<<(n + 255) / 256, 256, 0, st>>>(P, rec, n, isp, spv, cstart_new, rank, dst, err);
}
void launch_aos2soa(const double *rec, long long n, size_t so, const PartSoA &dst, cudaStream_t st) {
  if (n > 0) k_aos2soa<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(rec, n, so, dst);
}
void launch_soa2aos(const PartSoA &src, size_t so, long long n, double *rec, cudaStream_t st) {
  if (n > 0) k_soa2aos<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(src, so, n, rec);
}
void launch_relayout_from_aos(const DevParams &P, const double *rec, long long n, const int *tight, const int *cstart_dst,
                              const PartSoA &dst, size_t so, unsigned *err, cudaStream_t st) {
  if (n > 0) k_relayout_from_aos<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(P, rec, n, tight, cstart_dst, dst, so, err);
}
void launch_relayout_to_aos(const DevParams &P, const PartSoA &key, const PartSoA &val, size_t so, const int *cstart_src,
                            const int *tight, double *rec, cudaStream_t st) {
  k_relayout_to_aos<<<148 * 16, 256, 0, st>>>(P, key, val, so, cstart_src, tight, rec);
}
void launch_relayout_soa(const DevParams &P, const PartSoA &src, const int *cstart_src, const int *cnt_src, const PartSoA &dst,
                         const int *cstart_dst, unsigned *err, cudaStream_t st) {
  k_relayout_soa<<<148 * 16, 256, 0, st>>>(P, src, cstart_src, cnt_src, dst, cstart_dst, err);
}
void launch_incoming_append(const DevParams &P, const double *rec, int n, int isp, const int *cstart, int *cnt_tail,
                            const PartSoA &dst, double *ovf, int *ovfsp, int *ovfcnt, int ovfcap, unsigned *err,
                            cudaStream_t st, const int *n_dev, const int *cntb) {
  if (n > 0)
    k_incoming_append<<<(n + 255) / 256, 256, 0, st>>>(P, rec, n, n_dev, isp, cstart, cnt_tail, dst, ovf, ovfsp, ovfcnt, ovfcap, err, cntb);
}
// the sender's side of the same check: more leavers than the message holds
__global__ void k_check_counts(const int *__restrict__ cnt, int n, int lim0, int lim1, int lim2, int lim3, unsigned *err) {
  const int lim[4] = {lim0, lim1, lim2, lim3};
  if (threadIdx.x < n && cnt[threadIdx.x] > lim[threadIdx.x]) atomicOr(err, ERR_SENDBUF);
}
void launch_check_counts(const int *cnt, int n, const int lim[4], unsigned *err, cudaStream_t st) {
  k_check_counts<<<1, 32, 0, st>>>(cnt, n, lim[0], lim[1], lim[2], lim[3], err);
}
void launch_place(const DevParams &P, const double *stage, const uint32_t *tag, const PartSoA &dst, const int *cstart,
                  int *cnt_new, const int *tilebase, double *ovf, int *ovfsp, int *ovfcnt, int ovfcap, unsigned *err,
                  bool rim_only, cudaStream_t st, int tile0, int ntiles) {
  if (rim_only) {
    k_place_rim<<<ntiles > 0 ? ntiles : P.ntx * P.nty, PR_THREADS, 0, st>>>(P, reinterpret_cast<const double2 *>(stage), tag, dst, cstart, cnt_new,
                                                      tilebase, ovf, ovfsp, ovfcnt, ovfcap, err, tile0);
    return;
  }
  k_place<<<P.ntx * P.nty, PL_THREADS, 0, st>>>(P, reinterpret_cast<const double2 *>(stage), tag, dst, cstart, cnt_new, tilebase,
                                               ovf, ovfsp, ovfcnt, ovfcap, err);
}
void launch_mark_dead(const DevParams &P, PView<double> x, const int *cstart, const int *cnt_old, int *cnt_new, cudaStream_t st) {
  k_mark_dead<<<148 * 16, 256, 0, st>>>(P, x, cstart, cnt_old, cnt_new);
}
void launch_bcx(const DevParams &P, const PartSoA &g, const int *cstart, bool injection, cudaStream_t st) {
  k_bcx<<<148 * 8, 256, 0, st>>>(P, g, cstart, injection);
}
void launch_ic_weibel(const DevParams &P, const PartSoA &dst, int *cstart, int *cnt, uint64_t seed, int n0, double vti,
                      double vte, double t_ani, float sl, cudaStream_t st) {
  k_ic_weibel<<<148 * 16, 256, 0, st>>>(P, dst, cstart, cnt, seed, n0, vti, vte, t_ani, sl);
}
void launch_kinetic(const DevParams &P, const PartSoA &src, const int *cstart, int isp, double *partial, int nblocks,
                    cudaStream_t st) {
  k_kinetic<<<nblocks, 256, 0, st>>>(P, src, cstart, isp, partial);
}
void launch_moments(const DevParams &P, const PartSoA &src, PView<double> keyx, const int *cstart, double *mom,
                    cudaStream_t st) {
  k_moments<<<148 * 8, 256, 0, st>>>(P, src, keyx, cstart, mom);
}

}  // namespace wm
