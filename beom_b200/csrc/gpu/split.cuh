// split.cuh -- one kernel per loop of the reference (the general path: any coastline, any option).
// Each kernel cites the private_mod.f95 lines it implements and evaluates the same expressions in
// the same order (the library is built with -fmad=false, so results are bit-identical to a strict
// IEEE evaluation of the reference's formulas).  Neighbour names follow private_mod.f95:28-31:
// E=1, NE=2, N=3, NW=4, W=5, SW=6, S=7, SE=8; on the dense layout E/W = +-1 and N/S = +-NX.
#ifndef BEOM_SPLIT_CUH
#define BEOM_SPLIT_CUH
#include "dev.cuh"

namespace beom {

#define BEOM_CELL(D)                                              \
  const int x = (D).x_lo + blockIdx.x * blockDim.x + threadIdx.x; \
  const int y = (D).y_lo + blockIdx.y * blockDim.y + threadIdx.y; \
  if (x > (D).x_hi || y > (D).y_hi) return;                       \
  const int NX = (D).NX;                                          \
  const size_t c = (size_t)y * NX + x;                            \
  const uint8_t f = (D).flags[c];                                 \
  if (!(f & F_ACT)) return;

// tide term ramp*A*cos(phi - w*ctim) of private_mod.f95:1453-1454, 1538-1539, 1633-1634
// (comp 0,1,2 = eta,u,v)
__device__ __forceinline__ double tide_term(const Dev &D, size_t c, int comp) {
  const double amp = D.tide[((size_t)comp * 2 + 0) * D.plane + c];
  const double pha = D.tide[((size_t)comp * 2 + 1) * D.plane + c];
  return D.ramp * amp * cos(pha - D.w_ti * D.ctim);
}
__device__ __forceinline__ double single(double x) { return (double)(float)x; }

// The relaxation targets of one step with the tidal term added, for the fused step (whose stream table has no room for
// the six tidal planes): out[comp][layer] = fnud + ramp*A*cos(phi - w*ctim), the first two terms of the reference's
// hfor / ufor / vfor (private_mod.f95:1453-1454, 1538-1539, 1633-1634; eta: layer 1 only), same expressions as
// k_update_h / k_update_u / k_update_v above.  Every cell of the plane: the fused step also reads its halo and ghost cells.
// (Only used when the Ekman term that would come between the two is absent or exactly zero: fused_configure.)
__global__ void k_tide_targets(const __grid_constant__ Dev D, double *__restrict__ out) {
  const size_t c = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= D.plane) return;
  const int l = blockIdx.y;
  for (int comp = 0; comp < 3; comp++) {
    const size_t k = ((size_t)comp * D.nlay + l) * D.plane + c;
    double t = D.fnud[k];
    if (comp == 0) t = t + tide_term(D, c, 0) * (l == 0 ? 1.0 : 0.0);
    else t = t + tide_term(D, c, comp);
    out[k] = t;
  }
}

// ---- first_three_timesteps flux rebuild, private_mod.f95:2166-2177 (and 2208-2219) ----
__global__ void k_centred_flux(const __grid_constant__ Dev D) {
  BEOM_CELL(D)
  const size_t L = (size_t)blockIdx.z * D.plane;
  const double h = D.hlay[L + c];
  D.h_u[L + c] = D.u[L + c] * (h + D.hlay[L + c - 1]) / (1.0 + m_u(f));
  D.h_v[L + c] = D.v[L + c] * (h + D.hlay[L + c - NX]) / (1.0 + m_v(f));
}

// ---- rigid-lid flux rebuild, private_mod.f95:2238-2256, 2293-2310 (d2hx/d2hy of the LAST layer) ----
__global__ void k_upstream_flux(const __grid_constant__ Dev D) {
  BEOM_CELL(D)
  const size_t L = (size_t)blockIdx.z * D.plane, LL = (size_t)(D.nlay - 1) * D.plane;
  const double uu = D.u[L + c], vv = D.v[L + c], h = D.hlay[L + c];
  double hcen = (D.hlay[L + c - 1] + h) / (1.0 + m_u(f));
  D.h_u[L + c] = 0.5 * (uu + fabs(uu)) * (hcen - 0.16667 * D.d2hx[LL + c - 1]) + 0.5 * (uu - fabs(uu)) * (hcen - 0.16667 * D.d2hx[LL + c]);
  hcen = (D.hlay[L + c - NX] + h) / (1.0 + m_v(f));
  D.h_v[L + c] = 0.5 * (vv + fabs(vv)) * (hcen - 0.16667 * D.d2hy[LL + c - NX]) + 0.5 * (vv - fabs(vv)) * (hcen - 0.16667 * D.d2hy[LL + c]);
}

// ---- update_h, private_mod.f95:1593-1646 and the variants' epilogues ----
// One thread per water column, layers nlay..1 in sequence (the variants read other layers of the
// same column after they were updated).
__global__ void k_update_h(const __grid_constant__ Dev D) {
  BEOM_CELL(D)
  const double mk = m_n(f);
  const double nud = D.has_nudg ? D.nudg[c] : 0.0;
  const int si = x - D.i_off;
  const int half = D.lm / 2;
  for (int l = D.nlay - 1; l >= 0; l--) {
    const size_t L = (size_t)l * D.plane;
    double hold = D.hlay[L + c];
    double rs_3 = (D.h_u[L + c] - D.h_u[L + c + 1]) * D.i_dl + (D.h_v[L + c] - D.h_v[L + c + NX]) * D.i_dl;
    if (D.has_hdot) rs_3 = rs_3 + D.hdot[L + c];
    rs_3 = rs_3 * mk;
    const double rhsi = (D.c_ab1 * rs_3 - D.c_ab2 * D.rs2[L + c] + D.beta * D.rs1[L + c]) * D.dt * D.gene + rs_3 * D.dt * (1.0 - D.gene);
    hold = hold + rhsi;
    double hfor = D.fnud ? D.fnud[L + c] : 0.0;
    if (D.has_tide) hfor = hfor + tide_term(D, c, 0) * (l == 0 ? 1.0 : 0.0);
    double hnew;
    if (D.variant == BEOM_VARIANT_STANDARD) {  // pm:1637-1638
      hnew = D.has_nudg ? (hfor * nud + (1.0 - nud) * hold) : hold;
    } else {
      double h = hold;
      if (D.variant == BEOM_VARIANT_1D) {  // private_mod1d.f95:1637-1653
        const double h2 = (l == 1) ? hold : D.hlay[(size_t)1 * D.plane + c];
        if (h2 > 20 * D.hsal && si > half) {
          if (l == 0) h = h + 0 * nud + fmax(800.0 * nud + (-nud) * h, 0.0);
          else if (l == 1) h = h - 0 * nud + fmin(0.0 * nud + (-nud) * h, 0.0);
        }
      } else if (D.variant == BEOM_VARIANT_3D) {  // private_mod3d.f95:1637-1672
        const double h3 = (l == 2) ? hold : D.hlay[(size_t)2 * D.plane + c];
        if (h3 > 20 * D.hsal && si > half) {
          if (l == 0) h = h + 0 * nud + fmax(0.0 * nud + (-nud) * h, 0.0);
          else if (l == 1) h = h + 0 * nud + fmax(800.0 * nud + (-nud) * h, 0.0);
          else if (l == 2) h = h - 0 * nud + fmin(0.0 * nud + (-nud) * h, 0.0);
        } else if (h3 < 20 * D.hsal && si > half) {
          if (l == 0) h = h + 0 * nud + 1 * fmax(800.0 * nud + (-nud) * h, 0.0);
          else if (l == 1) h = h - 0 * nud + 1 * fmin(0.0 * nud + (-nud) * h, 0.0);
        }
      } else {  // private_modplumenew.f95:1637-1713, Lnud = 5000 from the start
        const double Lnud = 5000.0, Wnud = 8000.0, alph = (double)0.13f, Q0 = 250.0, gp0 = (double)0.265f;
        const double B0 = gp0 * Q0;
        const int lim = (int)floor((double)D.lm - Lnud / D.dl);
        if (si > lim) {
          if (D.plum < 0.5) {
            if (l == 0) h = h + 0 * nud + fmax(0.0 * nud + (-nud) * h, 0.0);
            else if (l == 1) h = h + 0 * nud + fmax(800.0 * nud + (-nud) * h, 0.0);
            else if (l == 2) h = h - 0 * nud + fmin(0.0 * nud + (-nud) * h, 0.0);
          } else {
            const double h2 = (l == 1) ? hold : D.hlay[(size_t)1 * D.plane + c];
            const double h3 = (l == 2) ? hold : D.hlay[(size_t)2 * D.plane + c];
            const double gp12 = 5 * B0 / (6 * alph) * 1.0 * 1.0 * (1.0 / (h2 + h3));
            const double gp23 = 5 * B0 / (6 * alph) * 1.0 * 1.0 * (1.0 / h3);
            const double w12 = 5 / (6 * alph) * 1.0 * 1.0 * 1.0, w23 = w12;
            const double R12 = 1 * alph * (h3 + h2), R23 = 1 * alph * h3;
            const double Q12 = D.pi / 2 * w12 * (R12 * R12), Q23 = D.pi / 2 * w23 * (R23 * R23);
            const double rhop1 = -gp12 * D.rhon[1] / D.grav + D.rhon[1];
            const double rhop2 = -gp23 * D.rhon[1] / D.grav + D.rhon[2];
            const double m1 = fmax(single((rhop1 - D.rhon[0]) / (D.rhon[1] - D.rhon[0])), 0.0);
            const double m2 = fmax(single((rhop2 - D.rhon[1]) / (D.rhon[2] - D.rhon[1])), 0.0);
            if (l == 0) h = h + D.dt * Q12 / (Wnud * Lnud) * (1 - m1);
            else if (l == 1) h = h + D.dt * (Q23 - Q12) / (Wnud * Lnud) - D.dt * Q23 / (Wnud * Lnud) * m2 + D.dt * Q12 / (Wnud * Lnud) * m1;
            else if (l == 2) h = h - D.dt * Q23 / (Wnud * Lnud) * (1 - m2);
          }
        }
      }
      hnew = h;
      if (si < half) hnew = hfor * nud + (1.0 - nud) * hold;
    }
    D.hlay[L + c] = hnew;
    D.rs_new[L + c] = rs_3;
  }
}

// ---- rgld column correction, private_mod.f95:1648-1654 (two layers, single precision) ----
__global__ void k_rgld_correct(const __grid_constant__ Dev D) {
  BEOM_CELL(D)
  double s = 0.0;
  for (int l = 0; l < D.nlay; l++) s += D.hlay[(size_t)l * D.plane + c];
  D.hlay[c] = D.hlay[c] - (double)(0.5f * (float)(s - D.h_th[c]));
  s = 0.0;
  for (int l = 0; l < D.nlay; l++) s += D.hlay[(size_t)l * D.plane + c];
  D.hlay[D.plane + c] = D.hlay[D.plane + c] - (double)(0.5f * (float)(s - D.h_th[c]));
}

// ---- update_mont_rvor_pvor_dive_kine, private_mod.f95:2318-2439 (all layers: grid.z) ----
__global__ void k_diag(const __grid_constant__ Dev D) {
  BEOM_CELL(D)
  const int l = blockIdx.z;
  const size_t L = (size_t)l * D.plane;
  const uint8_t fE = D.flags[c + 1], fW = D.flags[c - 1], fN = D.flags[c + NX], fS = D.flags[c - NX], fSW = D.flags[c - NX - 1];
  const double u_le = D.u[L + c], u_ri = D.u[L + c + 1], v_bo = D.v[L + c], v_to = D.v[L + c + NX];
  const double h = D.hlay[L + c], hE = D.hlay[L + c + 1], hW = D.hlay[L + c - 1], hN = D.hlay[L + c + NX], hS = D.hlay[L + c - NX],
               hSW = D.hlay[L + c - NX - 1];
  const double mk = m_n(f);
  double mpot;
  if (D.ocrp > 0.5) {  // pm:2351-2353
    mpot = h + D.hmin * (1.0 - mk);
    mpot = cube(D.hsal / mpot);
    mpot = mpot * (-D.ocrp * D.i_ns * D.hsal * mk);
  } else {
    mpot = -0.0;  // x * (-0.) for finite positive x
  }
  mpot = mpot - 0.0;  // h_to is never assigned (pm:2356)
  for (int i = 0; i < l; i++) mpot = mpot - (D.rhon[l] - D.rhon[i]) * D.i_rn[l] * D.hlay[(size_t)i * D.plane + c];  // pm:2357-2361
  if (D.rgld < 0.5) {  // pm:2365-2375
    double hcol = 0.0;
    for (int i = 0; i < D.nlay; i++) hcol = hcol + D.hlay[(size_t)i * D.plane + c];
    mpot = hcol - D.h_th[c] + mpot;
  }
  D.mont[L + c] = mpot + 0.25 * D.uadv * D.i_gr * (u_ri * u_ri + u_le * u_le + v_to * v_to + v_bo * v_bo);  // pm:2380-2383
  const double rv = (v_bo - D.v[L + c - 1] - u_le + D.u[L + c - NX]) * D.i_dl * m_pe(f);  // pm:2388
  D.rvor[L + c] = rv;
  double dx = (hE + hW - h * 2.0) * m_n(fE) * m_n(fW) * mk;  // pm:2394-2402
  double dy = (hN + hS - h * 2.0) * m_n(fN) * m_n(fS) * mk;
  if (D.ocrp > 0.5) {  // pm:2404-2416
    if (hE < D.two_hs || hW < D.two_hs || h < D.two_hs) dx = 0.0;
    if (hN < D.two_hs || hS < D.two_hs || h < D.two_hs) dy = 0.0;
  }
  D.d2hx[L + c] = dx;
  D.d2hy[L + c] = dy;
  const double have = h + hW + hSW + hS;  // pm:2421-2433
  D.pvor[L + c] = (D.fcor[c] + rv * D.uadv) * m_pi(f) * (mk + m_n(fW) + m_n(fSW) + m_n(fS)) / have;
  D.dive[L + c] = (u_ri - u_le + v_to - v_bo) * D.i_dl;  // pm:2435
}

// ---- update_viscosity (Leith part), private_mod.f95:2450-2502, and the Laplacians 2508-2550 ----
__device__ __forceinline__ double leith_ll(double r_bl, double r_br, double r_tl, double rbll, double rbbl, double d_cc,
                                           double d_le, double d_bl, double d_bo) {
  return (r_br - r_bl) * (r_br - r_bl) + (r_bl - rbll) * (r_bl - rbll) + (r_tl - r_bl) * (r_tl - r_bl) + (r_bl - rbbl) * (r_bl - rbbl) +
         (d_cc - d_le) * (d_cc - d_le) + (d_bo - d_bl) * (d_bo - d_bl) + (d_cc - d_bo) * (d_cc - d_bo) + (d_le - d_bl) * (d_le - d_bl);
}
__device__ __forceinline__ double leith_cc(double r_bl, double r_br, double r_tr, double r_tl, double d_cc, double d_ri,
                                           double d_to, double d_le, double d_bo) {
  return (r_br - r_bl) * (r_br - r_bl) + (r_tr - r_tl) * (r_tr - r_tl) + (r_tl - r_bl) * (r_tl - r_bl) + (r_tr - r_br) * (r_tr - r_br) +
         (d_ri - d_cc) * (d_ri - d_cc) + (d_cc - d_le) * (d_cc - d_le) + (d_to - d_cc) * (d_to - d_cc) + (d_cc - d_bo) * (d_cc - d_bo);
}

__global__ void k_visc(const __grid_constant__ Dev D) {
  BEOM_CELL(D)
  const size_t L = (size_t)blockIdx.z * D.plane;
  const double *rv = D.rvor + L, *dv = D.dive + L;
  const double r_bl = rv[c], r_br = rv[c + 1], r_tr = rv[c + NX + 1], r_tl = rv[c + NX], rbll = rv[c - 1], rbbl = rv[c - NX];
  const double d_cc = dv[c], d_ri = dv[c + 1], d_to = dv[c + NX], d_le = dv[c - 1], d_bl = dv[c - NX - 1], d_bo = dv[c - NX];
  D.v_ll[L + c] = sqrt(leith_ll(r_bl, r_br, r_tl, rbll, rbbl, d_cc, d_le, d_bl, d_bo)) * D.dvis * D.dl * D.dl + D.bvis;
  D.v_cc[L + c] = sqrt(leith_cc(r_bl, r_br, r_tr, r_tl, d_cc, d_ri, d_to, d_le, d_bo)) * D.dvis * D.dl * D.dl + D.bvis;
  if (D.svis > 0.0) {  // pm:2508-2550
    const uint8_t fE = D.flags[c + 1], fW = D.flags[c - 1], fN = D.flags[c + NX], fS = D.flags[c - NX];
    const double *uu = D.u + L, *vv = D.v + L;
    double du = 0.0, dw = 0.0;
    if (f & F_U) {
      du = du + 1.0 / (D.dl * D.dl) * (m_u(fE) * uu[c + 1] + m_u(fN) * uu[c + NX] + m_u(fW) * uu[c - 1] + m_u(fS) * uu[c - NX]);
      du = du - 1.0 / (D.dl * D.dl) * (m_u(fE) + m_u(fN) + m_u(fW) + m_u(fS)) * uu[c];
    }
    if (f & F_V) {
      dw = dw + 1.0 / (D.dl * D.dl) * (m_v(fE) * vv[c + 1] + m_v(fN) * vv[c + NX] + m_v(fW) * vv[c - 1] + m_v(fS) * vv[c - NX]);
      dw = dw - 1.0 / (D.dl * D.dl) * (m_v(fE) + m_v(fN) + m_v(fW) + m_v(fS)) * vv[c];
    }
    D.delu[L + c] = du;
    D.delv[L + c] = dw;
  }
}

// ---- biharmonic fluxes, private_mod.f95:2556-2599 ----
__global__ void k_biharm(const __grid_constant__ Dev D) {
  BEOM_CELL(D)
  const size_t L = (size_t)blockIdx.z * D.plane;
  const int si = x - D.i_off, sj = y - D.j_off;
  const uint8_t fW = D.flags[c - 1], fS = D.flags[c - NX], fSW = D.flags[c - NX - 1];
  const double h = D.hlay[L + c];
  const double hh_q = (double)(float)(h + m_n(fW) * D.hlay[L + c - 1] + m_n(fSW) * D.hlay[L + c - NX - 1] + m_n(fS) * D.hlay[L + c - NX]) /
                      (1.0 + m_n(fW) + m_n(fSW) + m_n(fS));
  const double *du = D.delu + L, *dv = D.delv + L;
  double U4 = 0.0, V4 = 0.0;
  U4 = U4 - D.i_dl * h * du[c] + D.i_dl * h * dv[c];
  V4 = V4 + D.i_dl * hh_q * du[c] + D.i_dl * hh_q * dv[c];
  if (si <= D.lm - 1) U4 = U4 + D.i_dl * h * du[c + 1];
  if (sj <= D.mm - 1) U4 = U4 - D.i_dl * h * dv[c + NX];
  if (si > 1) V4 = V4 - D.i_dl * hh_q * dv[c - 1];
  if (sj > 1) V4 = V4 - D.i_dl * hh_q * du[c - NX];
  if (m_u(f) * m_v(f) < 0.5) V4 = 0.0;
  D.UU4[L + c] = U4;
  D.VV4[L + c] = V4;
}

// ---- update_u, private_mod.f95:1422-1503 ----
__global__ void k_update_u(const __grid_constant__ Dev D) {
  BEOM_CELL(D)
  const int l = blockIdx.z;
  const size_t L = (size_t)l * D.plane;
  const size_t cN = c + NX, cNW = c + NX - 1, cW = c - 1;
  const double mask = m_u(f);
  const double hcen = (D.hlay[L + cW] + D.hlay[L + c]) / (1.0 + mask);
  const double i__h = 1.0 / (hcen + 1.0 - mask);
  double uold = D.u[L + c];
  const double dmd4 = (D.mont[L + cW] - D.mont[L + c]) * D.i_dl * D.grav * mask;
  const double *hv = D.h_v + L;
  double rhsi = dmd4 * (1.0 - D.gene) + 0.25 * D.pvor[L + c] * (hv[c] + hv[cW]) + 0.25 * D.pvor[L + cN] * (hv[cN] + hv[cNW]);
  if (D.has_wind) {
    const double tauw = 0.5 * (D.tt3d[((size_t)l * 2 + 0) * D.plane + cW] + D.tt3d[((size_t)l * 2 + 0) * D.plane + c]) * D.ramp;
    rhsi = rhsi + tauw * D.i_r0 * i__h;
  }
  if (D.has_bdrg) rhsi = rhsi - D.tb3d[((size_t)l * 2 + 0) * D.plane + c] * D.i_r0 * i__h;
  if (D.has_tdrg) rhsi = rhsi - D.tu3d[((size_t)l * 2 + 0) * D.plane + c] * D.i_r0 * i__h;
  rhsi = rhsi + D.bodf[0][l] + (D.del1 * dmd4 + D.del2 * D.dx3[L + c] + D.gamm * D.dx2[L + c] + D.epsi * D.dx1[L + c]) * D.gene;
  if (D.svis > 0.0) {  // pm:1471-1473: real() without kind rounds to single
    const float t4 = (float)(D.UU4[L + c] - D.UU4[L + cW] + D.VV4[L + cN] - D.VV4[L + c]);
    rhsi = rhsi - D.svis * D.i_dl * (double)t4 * i__h;
  } else {  // pm:1476-1479
    rhsi = rhsi + (D.v_cc[L + c] * D.dive[L + c] - D.v_cc[L + cW] * D.dive[L + cW]) * D.i_dl -
           (D.v_ll[L + cN] * D.rvor[L + cN] - D.v_ll[L + c] * D.rvor[L + c]) * D.i_dl;
  }
  uold = uold + rhsi * mask * D.dt;
  if (D.has_nudg) {
    double ufor = D.fnud[((size_t)1 * D.nlay + l) * D.plane + c];
    if (D.has_wind)
      ufor = ufor + 0.5 * (D.tt3d[((size_t)l * 2 + 1) * D.plane + c] + D.tt3d[((size_t)l * 2 + 1) * D.plane + cW]) * D.i_r1 * D.invf * i__h * D.ramp;
    if (D.has_tide) ufor = ufor + tide_term(D, c, 1);
    const double nu = D.nudg[D.plane + c];
    uold = ufor * nu + uold * (1.0 - nu);
  }
  D.u[L + c] = uold;
  if (D.rgld < 0.5)  // pm:1491-1496
    D.h_u[L + c] = 0.5 * (uold + fabs(uold)) * (hcen - 0.16667 * D.d2hx[L + cW]) + 0.5 * (uold - fabs(uold)) * (hcen - 0.16667 * D.d2hx[L + c]);
  D.dx_new[L + c] = dmd4;
}

// ---- update_v, private_mod.f95:1505-1591 ----
__global__ void k_update_v(const __grid_constant__ Dev D) {
  BEOM_CELL(D)
  const int l = blockIdx.z;
  const size_t L = (size_t)l * D.plane;
  const size_t cE = c + 1, cS = c - NX, cSE = c - NX + 1;
  const double mask = m_v(f);
  const double hcen = (D.hlay[L + c] + D.hlay[L + cS]) / (1.0 + mask);
  const double i__h = 1.0 / (hcen + 1.0 - mask);
  double vold = D.v[L + c];
  const double dmd4 = (D.mont[L + cS] - D.mont[L + c]) * D.i_dl * D.grav * mask;
  const double *hu = D.h_u + L;
  double rhsi = dmd4 * (1.0 - D.gene) - 0.25 * D.pvor[L + c] * (hu[c] + hu[cS]) - 0.25 * D.pvor[L + cE] * (hu[cE] + hu[cSE]);
  if (D.has_wind) {
    const double tauw = 0.5 * (D.tt3d[((size_t)l * 2 + 1) * D.plane + cS] + D.tt3d[((size_t)l * 2 + 1) * D.plane + c]) * D.ramp;
    rhsi = rhsi + tauw * D.i_r0 * i__h;
  }
  if (D.has_bdrg) rhsi = rhsi - D.tb3d[((size_t)l * 2 + 1) * D.plane + c] * D.i_r0 * i__h;
  if (D.has_tdrg) rhsi = rhsi - D.tu3d[((size_t)l * 2 + 1) * D.plane + c] * D.i_r0 * i__h;
  rhsi = rhsi + D.bodf[1][l] + (D.del1 * dmd4 + D.del2 * D.dy3[L + c] + D.gamm * D.dy2[L + c] + D.epsi * D.dy1[L + c]) * D.gene;
  if (D.svis > 0.0) {  // pm:1555-1557
    rhsi = rhsi - D.svis * D.i_dl * (D.VV4[L + cE] - D.VV4[L + c] - D.UU4[L + c] + D.UU4[L + cS]) * i__h;
  } else {  // pm:1561-1564
    rhsi = rhsi + (D.v_cc[L + c] * D.dive[L + c] - D.v_cc[L + cS] * D.dive[L + cS]) * D.i_dl +
           (D.v_ll[L + cE] * D.rvor[L + cE] - D.v_ll[L + c] * D.rvor[L + c]) * D.i_dl;
  }
  vold = vold + rhsi * mask * D.dt;
  if (D.has_nudg) {
    double vfor = D.fnud[((size_t)2 * D.nlay + l) * D.plane + c];
    if (D.has_wind)
      vfor = vfor - 0.5 * (D.tt3d[((size_t)l * 2 + 0) * D.plane + c] + D.tt3d[((size_t)l * 2 + 0) * D.plane + cS]) * D.i_r1 * D.invf * i__h * D.ramp;
    if (D.has_tide) vfor = vfor + tide_term(D, c, 2);
    const double nv = D.nudg[2 * D.plane + c];
    vold = vfor * nv + vold * (1.0 - nv);
  }
  D.v[L + c] = vold;
  if (D.rgld < 0.5)  // pm:1577-1582
    D.h_v[L + c] = 0.5 * (vold + fabs(vold)) * (hcen - 0.16667 * D.d2hy[L + cS]) + 0.5 * (vold - fabs(vold)) * (hcen - 0.16667 * D.d2hy[L + c]);
  D.dy_new[L + c] = dmd4;
}

// ---- distribute_stress, private_mod.f95:1921-2149 ----
// layt / layb / layu: fraction of the surface / bottom boundary layer occupied by each layer.
// Computed on every cell of the computed region INCLUDING inactive ones (the reference loops from
// ipnt = 0 and tb3d reads layb at the W/S neighbour, which may be the discarded cell).
__global__ void k_stress_fractions(const __grid_constant__ Dev D) {
  const int x = D.x_lo - 1 + blockIdx.x * blockDim.x + threadIdx.x;
  const int y = D.y_lo - 1 + blockIdx.y * blockDim.y + threadIdx.y;
  if (x > D.x_hi || y > D.y_hi) return;
  const size_t c = (size_t)y * D.NX + x;
  const int nlay = D.nlay;
  if (D.has_wind || D.has_tdrg) {  // pm:1945-1967 (layt) and pm:1991-2013 (layu): same formula
    double *dst = D.has_wind ? D.layt : D.layu;
    if (D.ocrp > 0.5) {
      double sofar = 0.0, hcum = 0.0;  // running sums give the same left-to-right partial sums
      for (int l = 0; l < nlay; l++) {
        hcum = hcum + fmax(0.0, D.hlay[(size_t)l * D.plane + c] - 1.5 * D.hsal);
        double t = fmin(hcum, D.hsbl) / D.hsbl - (sofar + 0.0);
        t = fmax(t, 0.0);
        dst[(size_t)l * D.plane + c] = t;
        sofar = sofar + t;
      }
    } else {
      for (int l = 0; l < nlay; l++) dst[(size_t)l * D.plane + c] = (l == 0) ? 1.0 : 0.0;
    }
    if (D.has_wind && D.has_tdrg)
      for (int l = 0; l < nlay; l++) D.layu[(size_t)l * D.plane + c] = D.layt[(size_t)l * D.plane + c];
  }
  if (D.has_bdrg) {  // pm:1969-1988
    if (D.ocrp > 0.5) {
      for (int l = nlay - 1; l >= 0; l--) {
        double sofar = 0.0, hcum = 0.0;  // sum(layb(ilay:nlay)) with layb(ilay) = 0, ascending
        for (int k = l + 1; k < nlay; k++) sofar += D.layb[(size_t)k * D.plane + c];
        for (int k = l; k < nlay; k++) hcum += D.hlay[(size_t)k * D.plane + c];
        double t = fmin(hcum, D.hbbl) / D.hbbl - sofar;
        D.layb[(size_t)l * D.plane + c] = fmax(t, 0.0);
      }
    } else {
      for (int l = 0; l < nlay; l++) D.layb[(size_t)l * D.plane + c] = (l == nlay - 1) ? 1.0 : 0.0;
    }
  }
}

// taub / taum, private_mod.f95:2015-2049 and 2075-2109 (which = 0 bottom, 1 top)
__global__ void k_stress_drag(const __grid_constant__ Dev D, int which) {
  BEOM_CELL(D)
  const int nlay = D.nlay;
  int l = which == 0 ? nlay - 1 : 0;
  if (D.ocrp > 0.5) {
    if (which == 0) {
      for (int k = nlay - 1; k >= 0; k--)
        if (D.hlay[(size_t)k * D.plane + c] > D.bstress_thr) { l = k; break; }
    } else {
      for (int k = 0; k < nlay; k++)
        if (D.hlay[(size_t)k * D.plane + c] > D.two_hs) { l = k; break; }
    }
  }
  const double *uu = D.u + (size_t)l * D.plane, *vv = D.v + (size_t)l * D.plane;
  const double vatu = 0.25 * vv[c] + 0.25 * vv[c + NX] + 0.25 * vv[c + NX - 1] + 0.25 * vv[c - 1];
  const double uatv = 0.25 * uu[c] + 0.25 * uu[c + 1] + 0.25 * uu[c - NX] + 0.25 * uu[c - NX + 1];
  const double coef = which == 0 ? D.bdrg : D.tdrg;
  double *tau = which == 0 ? D.taub : D.taum;
  tau[c] = uu[c] * coef * D.rhon[l] * (D.qdrg * sqrt(uu[c] * uu[c] + vatu * vatu) + 1.0 - D.qdrg);
  tau[D.plane + c] = vv[c] * coef * D.rhon[l] * (D.qdrg * sqrt(vv[c] * vv[c] + uatv * uatv) + 1.0 - D.qdrg);
}

// tb3d / tu3d / tt3d, private_mod.f95:2056-2071, 2116-2133, 2136-2146
__global__ void k_stress_apply(const __grid_constant__ Dev D) {
  BEOM_CELL(D)
  const int l = blockIdx.z;
  const size_t L = (size_t)l * D.plane, T0 = ((size_t)l * 2) * D.plane, T1 = T0 + D.plane;
  if (D.has_bdrg) {
    D.tb3d[T0 + c] = D.taub[c] * 0.5 * (D.layb[L + c] + D.layb[L + c - 1]);
    D.tb3d[T1 + c] = D.taub[D.plane + c] * 0.5 * (D.layb[L + c] + D.layb[L + c - NX]);
  }
  if (D.has_tdrg) {
    D.tu3d[T0 + c] = D.taum[c] * 0.5 * (D.layu[L + c] + D.layu[L + c - 1]);
    D.tu3d[T1 + c] = D.taum[D.plane + c] * 0.5 * (D.layu[L + c] + D.layu[L + c - NX]);
  }
  if (D.has_wind) {
    D.tt3d[T0 + c] = D.taus[c] * D.layt[L + c];
    D.tt3d[T1 + c] = D.taus[D.plane + c] * D.layt[L + c];
  }
}

// ---- no_gradient_obc, private_mod.f95:2613-2679: seg[k] = {flag4, flag5, c1, c10, c13, c16} as dense cells.
// The two passes of the reference are two launches (pass = 0 tangential, 1 normal).
struct SegDev {
  int zonal, merid;
  int c_face, c_wet, c_norm, c_int;  // segm columns 1, 10, 13, 16 as dense cell offsets
};
__global__ void k_obc(const __grid_constant__ Dev D, const SegDev *seg, int nseg, int pass) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const int NX = D.NX;
  const SegDev g = seg[s];
  for (int l = 0; l < D.nlay; l++) {
    const size_t L = (size_t)l * D.plane;
    const double *fu = D.fnud + ((size_t)1 * D.nlay + l) * D.plane, *fv = D.fnud + ((size_t)2 * D.nlay + l) * D.plane;
    if (pass == 0) {
      const size_t p = g.c_wet, q = g.c_int;
      const uint8_t f = D.flags[p];
      if (g.merid == 1) {
        if (f & F_U) {
          D.u[L + p] = D.u[L + q] - fu[q] + fu[p];
          D.h_u[L + p] = D.u[L + p] * (D.hlay[L + p] + D.hlay[L + p - 1]) / (1.0 + m_u(f));
        }
      } else if (g.zonal == 1) {
        if (f & F_V) {
          D.v[L + p] = D.v[L + q] - fv[q] + fv[p];
          D.h_v[L + p] = D.v[L + p] * (D.hlay[L + p] + D.hlay[L + p - NX]) / (1.0 + m_v(f));
        }
      }
    } else {
      const size_t p = g.c_face, q = g.c_norm;
      const uint8_t f = D.flags[p];
      if (g.merid == 1) {
        D.v[L + p] = D.v[L + q] - fv[q] + fv[p];
        D.h_v[L + p] = D.v[L + p] * (D.hlay[L + p] + D.hlay[L + p - NX]) / (1.0 + m_v(f));
      } else if (g.zonal == 1) {
        D.u[L + p] = D.u[L + q] - fu[q] + fu[p];
        D.h_u[L + p] = D.u[L + p] * (D.hlay[L + p] + D.hlay[L + p - 1]) / (1.0 + m_u(f));
      }
    }
  }
}

// ---- surf_pressure, private_mod.f95:1705-1838 (rigid lid) ----
// Right-hand side in gather form.  The reference scatters -h_u/(dl dt) to ipnt and +h_u/(dl dt) to its west
// neighbour while ipnt ascends (pm:1725-1750); for a given point the order of those updates is: own
// x-flux, east neighbour's x-flux, own y-flux, north neighbour's y-flux, layer by layer from nlay to 1.
__global__ void k_pi_rhs(const __grid_constant__ Dev D) {
  BEOM_CELL(D)
  const int si = x - D.i_off, sj = y - D.j_off;
  const double dd = D.dl * D.dt;
  const bool eE = D.flags[c + 1] & F_ACT, eN = D.flags[c + NX] & F_ACT;
  double r = 0.0;
  for (int l = D.nlay - 1; l >= 0; l--) {
    const size_t L = (size_t)l * D.plane;
    if (si > 1) r = r - D.h_u[L + c] / dd;
    if (eE) r = r + D.h_u[L + c + 1] / dd;
    if (sj > 1) r = r - D.h_v[L + c] / dd;
    if (eN) r = r + D.h_v[L + c + NX] / dd;
  }
  D.pi_rhs[c] = r;
}

// The reference's sweep is lexicographic Gauss-Seidel (ipnt ascending: row by row, west to east), each
// point using the already-updated west and south values and the old east and north values
// (pm:1766-1792).  Updating the anti-diagonals x + y = const one after the other, all points of a diagonal
// in parallel, has exactly the same data dependencies, hence the same iterates bit for bit.
// One thread block; max-norm stopping test as in the reference (pm:1756-1803).
__global__ void __launch_bounds__(1024, 1) k_surf_pressure(const __grid_constant__ Dev D, int maxiters, double pi_tol, int *iters_out) {
  __shared__ double red[32];
  __shared__ double s_max;
  const int NX = D.NX;
  const int w = D.x_hi - D.x_lo + 1, h = D.y_hi - D.y_lo + 1;
  const double rp = 1.0;
  int iters = 0;
  double maxdiff = pi_tol + 1;
  while (maxdiff > pi_tol && iters < maxiters) {
    double local = 0.0;
    for (int d = 0; d <= w + h - 2; d++) {
      for (int ry = threadIdx.x; ry < h; ry += blockDim.x) {
        const int rx = d - ry;
        if (rx < 0 || rx >= w) continue;
        const int x = D.x_lo + rx, y = D.y_lo + ry;
        const size_t c = (size_t)y * NX + x;
        if (!(D.flags[c] & F_ACT)) continue;
        const int si = x - D.i_off, sj = y - D.j_off;
        const double prev = D.pi_s[c], os = D.Osum_[c];
        double v = (1 - rp) * prev - rp * os * D.pi_rhs[c];
        if (si < D.lm) v = v + rp * os * D.Ow[c + 1] * D.pi_s[c + 1];
        if (sj < D.mm) v = v + rp * os * D.Os[c + NX] * D.pi_s[c + NX];
        if (si > 1) v = v + rp * os * D.Ow[c] * D.pi_s[c - 1];
        if (sj > 1) v = v + rp * os * D.Os[c] * D.pi_s[c - NX];
        D.pi_s[c] = v;
        local = fmax(local, fabs(v - prev));
      }
      __syncthreads();
    }
    for (int o = 16; o > 0; o >>= 1) local = fmax(local, __shfl_xor_sync(0xffffffffu, local, o));
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = local;
    __syncthreads();
    if (threadIdx.x < 32) {
      double m = threadIdx.x < (blockDim.x >> 5) ? red[threadIdx.x] : 0.0;
      for (int o = 16; o > 0; o >>= 1) m = fmax(m, __shfl_xor_sync(0xffffffffu, m, o));
      if (threadIdx.x == 0) s_max = m;
    }
    __syncthreads();
    maxdiff = s_max;
    iters++;
    __syncthreads();
  }
  if (threadIdx.x == 0 && iters_out) *iters_out = iters;
}

// velocity projection, private_mod.f95:1806-1833
__global__ void k_pi_correct(const __grid_constant__ Dev D) {
  BEOM_CELL(D)
  const size_t L = (size_t)blockIdx.z * D.plane;
  const int si = x - D.i_off, sj = y - D.j_off;
  const double fac = D.dt / D.dl;
  if (si > 1 && si < D.lm + 1) {
    double t = D.u[L + c] - fac * D.pi_s[c];
    D.u[L + c] = t + fac * D.pi_s[c - 1];
  }
  if (sj > 1 && sj < D.mm + 1) {
    double t = D.v[L + c] - fac * D.pi_s[c];
    D.v[L + c] = t + fac * D.pi_s[c - NX];
  }
}

// ---- mirror cells (periodic aliases, slab halos inside one device): dst <- src for np planes ----
__global__ void k_mirror(double *field, size_t plane, int nplanes, const int *dst, const int *src, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int d = dst[k], s = src[k];
  for (int p = 0; p < nplanes; p++) field[(size_t)p * plane + d] = field[(size_t)p * plane + s];
}

// ---- y-slab halos: pack the G outermost owned rows of each plane / unpack the neighbours' rows ----
constexpr int kHaloMaxItems = 16;
struct HaloTab {
  double *p[kHaloMaxItems];
  int planes[kHaloMaxItems], off[kHaloMaxItems];
  int n, total;
};
__device__ __forceinline__ double *halo_plane(const HaloTab &t, int gp, size_t plane) {
  int k = 0;
  while (k + 1 < t.n && gp >= t.off[k + 1]) k++;
  return t.p[k] + (size_t)(gp - t.off[k]) * plane;
}
// blockIdx.z = 0: G rows from row_lo -> send_lo; 1: G rows from row_hi -> send_hi (layout.h: halo_rows; plain slabs send
// y_lo .. y_lo+G-1 and y_hi-G+1 .. y_hi, the last rank of a y-periodic ring one row lower)
__global__ void k_halo_pack(const __grid_constant__ HaloTab t, size_t plane, int NX, int row_lo, int row_hi, double *send_lo, double *send_hi) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)G * NX;
  if (k >= n) return;
  const int gp = blockIdx.y;
  const double *src = halo_plane(t, gp, plane) + (size_t)(blockIdx.z == 0 ? row_lo : row_hi) * NX;
  (blockIdx.z == 0 ? send_lo : send_hi)[(size_t)gp * n + k] = src[k];
}
// blockIdx.z = 0: recv_lo -> G rows from row_lo (plain: y_lo-G); 1: recv_hi -> G rows from row_hi (plain: y_hi+1)
__global__ void k_halo_unpack(const __grid_constant__ HaloTab t, size_t plane, int NX, int row_lo, int row_hi, const double *recv_lo, const double *recv_hi,
                              int has_lo, int has_hi) {
  const size_t k = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n = (size_t)G * NX;
  if (k >= n) return;
  if (blockIdx.z == 0 ? !has_lo : !has_hi) return;
  const int gp = blockIdx.y;
  double *dst = halo_plane(t, gp, plane) + (size_t)(blockIdx.z == 0 ? row_lo : row_hi) * NX;
  dst[k] = (blockIdx.z == 0 ? recv_lo : recv_hi)[(size_t)gp * n + k];
}

// ---- vector <-> dense ----
template <class T>
__global__ void k_scatter(T *dense, const T *vec, const int *cell, int p0, int n) {  // dense[cell[p]] = vec[p]
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int cc = cell[p0 + k];
  if (cc >= 0) dense[cc] = vec[k];
}
template <class T>
__global__ void k_gather(T *vec, const T *dense, const int *cell, int p0, int n) {  // vec[p] = dense[cell[p]]
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int cc = cell[p0 + k];
  vec[k] = cc >= 0 ? dense[cc] : T(0);
}
// history arrays go out in the reference's (k, ipnt) order, k fastest
__global__ void k_gather_hist(double *vec, const double *d1, const double *d2, const double *d3, int nh, const int *cell, int p0, int n) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  const int cc = cell[p0 + k];
  vec[(size_t)k * nh + 0] = cc >= 0 ? d1[cc] : 0.0;
  vec[(size_t)k * nh + 1] = cc >= 0 ? d2[cc] : 0.0;
  if (nh == 3) vec[(size_t)k * nh + 2] = cc >= 0 ? d3[cc] : 0.0;
}

}  // namespace beom
#endif
