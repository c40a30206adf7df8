// dev.cuh -- device-side view of the model (dense layout) shared by all kernels.
//
// HBM layout.  The reference stores fields as compressed wet-point vectors addressed through an
// 8-neighbour table (private_mod.f95:27-33).  Every vector point is a point (i,j) of the padded
// (lm+2) x (mm+2) grid and neig(k,p) is always the vector entry of grid point (i+di, j+dj) (or 0, or a
// periodic alias; private_mod.f95:614-726).  On the device each field is therefore a dense plane
//     f[layer][Y][X],   X = i + GX0,  Y = j - jbase,   pitch NX (multiple of 16 doubles)
// with a halo of G cells on every side.  Cells that are not vector points hold 0 for ever, which is
// exactly what the reference's index 0 ("discarded cell") reads as; periodic aliases and slab halos
// are "mirror cells" refreshed after each update.  Neighbours are +-1 / +-NX: no index table is read
// in the hot path.  Conversion to/from the reference's vector layout happens only in
// upload/download (cell_of_point).
#ifndef BEOM_DEV_CUH
#define BEOM_DEV_CUH
#include <cstdint>

#include "../../../include/beom_gpu.h"
#include "layout.h"  // G, GX0 and the per-cell flag bits F_N .. F_GHOST

namespace beom {

struct Dev {
  // geometry
  int NX, NY;         // pitch and number of rows of a plane
  int x_lo, x_hi;     // active X range (inclusive): i = 1 .. lm+1
  int y_lo, y_hi;     // rows this rank computes (inclusive)
  int i_off, j_off;   // i = X - i_off, j = Y - j_off (global grid indices)
  int lm, mm, nlay;
  size_t plane;       // NX * NY
  // static fields
  const uint8_t *flags;
  const double *fcor, *h_th;
  const double *nudg;  // [3] planes or null
  const double *fnud;  // [3][nlay] planes or null
  const double *hdot;  // [nlay] or null
  const double *taus;  // [2] or null
  const double *tide;  // [3][2] planes (amp, phase) or null
  // state (current read / write buffers; equal for in-place kernels)
  double *hlay, *u, *v, *h_u, *h_v;  // [nlay]
  double *rs1, *rs2, *rs_new;        // rs_h(1), rs_h(2) and the slot receiving rs_3
  double *dx1, *dx2, *dx3, *dx_new;  // dmdx(1..3) and the slot receiving dmd4
  double *dy1, *dy2, *dy3, *dy_new;
  // per-layer temporaries of the split path
  double *mont, *rvor, *pvor, *dive, *d2hx, *d2hy, *v_cc, *v_ll;  // [nlay]
  double *tt3d, *tb3d, *tu3d;  // [nlay][2] or null
  double *layt, *layb, *layu;  // [nlay]
  double *taub, *taum;         // [2]
  double *delu, *delv, *UU4, *VV4;  // [nlay] (svis > 0)
  double *pi_s, *pi_rhs, *pi_prev;
  const double *Ow, *Os, *Osum_;
  // scalars of the step (private_mod.f95:69-73)
  double ctim, ramp, gene, invf, w_ti;
  // parameters (shared_mod.f95) and derived constants, computed on the host in the same order
  // the reference computes them
  double dl, dt, i_dl, i_gr, i_r0, i_r1, grav, uadv, ocrp, qdrg, rgld;
  double hsal, hmin, hsbl, hbbl, i_ns, two_hs;  // two_hs = 2._r8 * hs_8
  double bvis, dvis, svis, bdrg, tdrg, beta, epsi, gamm, del1, del2, plum, pi;
  double c_ab1, c_ab2;  // (1.5 + beta), (0.5 + 2 beta)
  double rhon[BEOM_MAXLAY], i_rn[BEOM_MAXLAY], bodf[2][BEOM_MAXLAY];
  double bstress_thr;   // 2*hsal (0*hsal in the 1d variant, private_mod1d.f95:2046)
  int variant, nsal;
  // feature switches (file present / parameter non-zero)
  int has_nudg, has_tide, has_hdot, has_wind, has_bdrg, has_tdrg, has_visc_term, mask_check;
};

__device__ __forceinline__ double m_n(uint8_t f) { return (f & F_N) ? 1.0 : 0.0; }
__device__ __forceinline__ double m_u(uint8_t f) { return (f & F_U) ? 1.0 : 0.0; }
__device__ __forceinline__ double m_v(uint8_t f) { return (f & F_V) ? 1.0 : 0.0; }
__device__ __forceinline__ double m_pe(uint8_t f) { return (f & F_PE) ? 1.0 : 0.0; }
__device__ __forceinline__ double m_pi(uint8_t f) { return (f & F_PI) ? 1.0 : 0.0; }

// x**3 as the reference's integer power (oracle convention: (x*x)*x)
__device__ __forceinline__ double cube(double x) { return (x * x) * x; }

}  // namespace beom
#endif
