/* beom_oracle.c -- TEST INFRASTRUCTURE ONLY (see beom_oracle.h).
 *
 * A CPU restatement of the reference's algorithm: read_input_data and the per-timestep update of
 * private_mod.f95 (and the update_h epilogues of private_mod1d/3d/plumenew.f95), in plain C with the
 * reference's own vector layout, neighbour table and expression order.  Every function cites the
 * reference lines it follows ("pm:" = private_mod.f95, "sm:" = shared_mod.f95).
 *
 * Build: gcc -O2 -ffp-contract=off (strict IEEE double; no FMA contraction, no reassociation).
 * With -fopenmp the same `ipnt' loops the reference parallelises are parallel (bench cpu_baseline).
 *
 * PARITY PINNED (round 2): the reference holds no golden vectors and no Fortran compiler exists here, but its own
 * sources, translated to C++ by oracle/f95c and run in the development container (oracle/_ref), leave every module
 * array bit-identical to this restatement on all 16 reference scripts, 24 option / variant cases and the three variant
 * files (tests/test_reference_pin.py); tests/golden holds outputs of that program.  See oracle/README.md.
 *
 * Conventions fixed by this restatement where Fortran leaves latitude:
 *   - SUM() intrinsics are evaluated sequentially in array order;
 *   - x**3 = (x*x)*x and x**4 = (x*x)*(x*x) (what gfortran's powi expansion and libgcc's __powidf2
 *     both produce);
 *   - cext**2._rw is the correctly rounded square (gfortran folds the parameter expression).
 */
#include "beom_oracle.h"

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

struct beom_oracle {
  beom_params P;
  int lm, mm, nlay, ndeg, nd1;
  int32_t *neig, *subc, *posc, *segm;
  int nseg, flag_nudging;
  double *h_u, *h_v, *u, *v, *UU4, *VV4, *delu, *delv;
  double *tt3d, *tb3d, *tu3d, *taus;
  double *rvor, *pvor, *dive, *v_cc, *v_ll, *fcor;
  double *mk_u, *mk_v, *mk_n, *mkpe, *mkpi;
  double *mont, *d2hx, *d2hy, *h_bo, *h_to, *h_th, *Ow, *Os, *Osum_;
  double ctim, invf, ramp, gene, w_ti, tres;
  double *bodf, *fnud, *nudg, *hdot, *rs_h, *dmdx, *dmdy, *tide;
  double *hlay, *pi_s, *h_0, *h_2d;
  /* scratch of distribute_stress (pm:1925-1929) */
  double *layt, *layb, *layu, *taub, *taum;
  int nstp, notp, n_3d;
};

static char g_err[1024];
const char *beom_oracle_error(void) { return g_err; }

/* ---- index helpers (1-based layer / component, 0-based point like the reference) ---- */
#define ND1 (o->nd1)
#define A2(x, ip, il) ((x)[(size_t)((il) - 1) * ND1 + (ip)])
#define NEIG(k, ip) (o->neig[(size_t)(ip) * 8 + ((k) - 1)])
#define SUBC(ip, d) (o->subc[(size_t)((d) - 1) * ND1 + (ip)])
#define FNUD(ip, il, ix) (o->fnud[((size_t)((ix) - 1) * o->nlay + ((il) - 1)) * ND1 + (ip)])
#define NUDG(ip, ix) (o->nudg[(size_t)((ix) - 1) * ND1 + (ip)])
#define TIDE(a, ip, ix) (o->tide[((size_t)((ix) - 1) * ND1 + (ip)) * 2 + ((a) - 1)])
#define T3(x, ip, c, il) ((x)[((size_t)((il) - 1) * 2 + ((c) - 1)) * ND1 + (ip)])
#define TAUS(ip, c) (o->taus[(size_t)((c) - 1) * ND1 + (ip)])
#define TAUB(x, ip, c) ((x)[(size_t)((c) - 1) * ND1 + (ip)])
#define RSH(k, ip, il) (o->rs_h[((size_t)((il) - 1) * ND1 + (ip)) * 2 + ((k) - 1)])
#define DM(x, k, ip, il) ((x)[((size_t)((il) - 1) * ND1 + (ip)) * 3 + ((k) - 1)])
#define BODF(il, c) (o->bodf[(size_t)((c) - 1) * o->nlay + ((il) - 1)])
#define SEGM(is, col) (o->segm[(size_t)((col) - 1) * o->nseg + ((is) - 1)])
#define H2(i, j) (o->h_2d[(size_t)((j) + 1) * (o->lm + 4) + ((i) + 1)])

enum { ix_n = 1, ix_u = 2, ix_v = 3 };

static double *dalloc(size_t n) {
  double *p = (double *)calloc(n ? n : 1, sizeof(double));
  if (!p) { fprintf(stderr, "beom_oracle: out of memory\n"); abort(); }
  return p;
}
static int32_t *ialloc(size_t n) {
  int32_t *p = (int32_t *)calloc(n ? n : 1, sizeof(int32_t));
  if (!p) { fprintf(stderr, "beom_oracle: out of memory\n"); abort(); }
  return p;
}
static double dmax2(double a, double b) { return a > b ? a : b; }
static double dmin2(double a, double b) { return a < b ? a : b; }
/* Fortran NINT: round half away from zero (pm:1854-1856) */
static long f_nint(double x) { return (long)(x >= 0.0 ? floor(x + 0.5) : -floor(-x + 0.5)); }
static double pow3(double x) { return (x * x) * x; }
static double pow4(double x) { double y = x * x; return y * y; }

/* read a raw little-endian float32 file; returns NULL if absent (pm:775-776) */
static float *read_f32(const char *idir, const char *keyw, size_t n, int *err) {
  char path[2048];
  FILE *f;
  float *buf;
  *err = 0;
  if (!idir || !idir[0]) return NULL;
  snprintf(path, sizeof path, "%s%s%s.bin", idir, idir[strlen(idir) - 1] == '/' ? "" : "/", keyw);
  f = fopen(path, "rb");
  if (!f) return NULL;
  buf = (float *)malloc(n * sizeof(float));
  if (!buf || fread(buf, sizeof(float), n, f) != n) {
    snprintf(g_err, sizeof g_err, "could not open/read file %s.bin from directory %s", keyw, idir);
    *err = 1;
    free(buf);
    fclose(f);
    return NULL;
  }
  fclose(f);
  return buf;
}

/* ---------------------------------------------------------------------------------------------
 * index_grid_points, pm:567-764
 * ------------------------------------------------------------------------------------------- */
static int any_wet(const beom_oracle *o, int i, int j) {
  const double hdry = o->P.hdry;
  return H2(i, j) > hdry || H2(i - 1, j) > hdry || H2(i, j - 1) > hdry || H2(i - 1, j - 1) > hdry;
}

static int index_grid_points(beom_oracle *o) {
  const int lm = o->lm, mm = o->mm;
  const double hdry = o->P.hdry;
  const int w = lm + 4;
  int32_t *indc = ialloc((size_t)w * (mm + 4));
  int i, j, i_c = 0;
#define INDC(i, j) indc[(size_t)((j) + 1) * w + ((i) + 1)]
  for (j = 0; j <= mm + 1; j++) /* pm:592-602 */
    for (i = 0; i <= lm + 1; i++)
      if (any_wet(o, i, j)) INDC(i, j) = ++i_c;
  if (i_c != o->ndeg) { /* pm:604-610 */
    snprintf(g_err, sizeof g_err, "wrong input parameter! Please set ndeg = %d", i_c);
    free(indc);
    return -1;
  }
  if (o->P.xper > 0.5) { /* pm:614-640 */
    i = 1;
    for (j = 1; j <= mm; j++) {
      if (H2(i, j) > hdry && H2(lm, j) > hdry) {
        INDC(0, j) = INDC(lm, j);
        INDC(lm + 1, j) = INDC(1, j);
        o->mk_u[INDC(i, j)] = 1.0;
      }
      if (j > 1 && j <= mm)
        if (H2(i, j - 1) > hdry && H2(i, j) > hdry && H2(lm, j - 1) > hdry && H2(lm, j) > hdry)
          o->mkpe[INDC(i, j)] = 1.0;
      if (j == mm)
        if (H2(i, j) > hdry && H2(lm, j) > hdry) {
          INDC(0, mm + 1) = INDC(lm, mm + 1);
          INDC(lm + 1, mm + 1) = INDC(1, mm + 1);
        }
    }
  }
  if (o->P.yper > 0.5) { /* pm:642-668 */
    j = 1;
    for (i = 1; i <= lm; i++) {
      if (H2(i, j) > hdry && H2(i, mm) > hdry) {
        INDC(i, 0) = INDC(i, mm);
        INDC(i, mm + 1) = INDC(i, 1);
        o->mk_v[INDC(i, j)] = 1.0;
      }
      if (i > 1 && i <= lm)
        if (H2(i - 1, j) > hdry && H2(i, j) > hdry && H2(i - 1, mm) > hdry && H2(i, mm) > hdry)
          o->mkpe[INDC(i, j)] = 1.0;
      if (i == lm)
        if (H2(i, mm) > hdry && H2(i, j) > hdry) {
          INDC(lm + 1, 0) = INDC(lm + 1, mm);
          INDC(lm + 1, mm + 1) = INDC(lm + 1, 1);
        }
    }
  }
  if (o->P.xper > 0.5 && o->P.yper > 0.5) { /* pm:672-685 */
    if (H2(1, 1) > hdry && H2(lm, 1) > hdry && H2(1, mm) > hdry) {
      INDC(0, 0) = INDC(lm, mm);
      o->mkpe[INDC(1, 1)] = 1.0;
      INDC(0, mm + 1) = INDC(lm, 1);
    }
    if (H2(lm, mm) > hdry && H2(1, mm) > hdry && H2(lm, 1) > hdry) {
      INDC(lm + 1, 0) = INDC(1, mm);
      INDC(lm + 1, mm + 1) = INDC(1, 1);
    }
  }
  i_c = 0; /* pm:690-730 */
  for (j = 0; j <= mm + 1; j++)
    for (i = 0; i <= lm + 1; i++) {
      if (!any_wet(o, i, j)) continue;
      ++i_c;
      if (H2(i, j) > hdry) o->mk_n[i_c] = 1.0;
      if (H2(i - 1, j) > hdry && H2(i, j) > hdry) o->mk_u[i_c] = 1.0;
      if (H2(i, j - 1) > hdry && H2(i, j) > hdry) o->mk_v[i_c] = 1.0;
      if (H2(i - 1, j - 1) > hdry && H2(i, j - 1) > hdry && H2(i - 1, j) > hdry && H2(i, j) > hdry)
        o->mkpe[i_c] = 1.0;
      if (H2(i - 1, j - 1) > hdry || H2(i, j - 1) > hdry || H2(i - 1, j) > hdry || H2(i, j) > hdry)
        o->mkpi[i_c] = 1.0;
      o->posc[i_c] = i + 1 + j * (lm + 2);
      SUBC(i_c, 1) = i;
      SUBC(i_c, 2) = j;
      NEIG(1, i_c) = INDC(i + 1, j);
      NEIG(2, i_c) = INDC(i + 1, j + 1);
      NEIG(3, i_c) = INDC(i, j + 1);
      NEIG(4, i_c) = INDC(i - 1, j + 1);
      NEIG(5, i_c) = INDC(i - 1, j);
      NEIG(6, i_c) = INDC(i - 1, j - 1);
      NEIG(7, i_c) = INDC(i, j - 1);
      NEIG(8, i_c) = INDC(i + 1, j - 1);
    }
  for (i_c = 0; i_c <= o->ndeg; i_c++) /* pm:753-757 */
    o->h_th[i_c] = H2(SUBC(i_c, 1), SUBC(i_c, 2));
#undef INDC
  free(indc);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * get_equilibrium_thickness_h_0, pm:309-565
 * ------------------------------------------------------------------------------------------- */
static int get_equilibrium_thickness_h_0(beom_oracle *o) {
  const int nlay = o->nlay, ndeg = o->ndeg, lm = o->lm, mm = o->mm;
  const beom_params *P = &o->P;
  const double thre = P->tole, hsal = P->hsal, sor = P->sor;
  double dmax, rho8[BEOM_MAXLAY], gues0[BEOM_MAXLAY], cons[BEOM_MAXLAY];
  int ilay, k, ipnt, failed = 0;
  size_t n;

  dmax = o->h_2d[0]; /* pm:334 */
  for (n = 1; n < (size_t)(lm + 4) * (mm + 4); n++)
    if (o->h_2d[n] > dmax) dmax = o->h_2d[n];
  for (ilay = 1; ilay <= nlay; ilay++) rho8[ilay - 1] = P->rhon[ilay - 1];
  for (ilay = 1; ilay <= nlay; ilay++) { /* pm:339-345 */
    gues0[ilay - 1] = dmax * (1.0 - P->topl[ilay - 1]);
    if (ilay < nlay) gues0[ilay - 1] = gues0[ilay - 1] - dmax * (1.0 - P->topl[ilay]);
  }
  for (ilay = 1; ilay <= nlay; ilay++) { /* pm:349-355 */
    double s = 0.0;
    for (k = 0; k < nlay; k++) s += gues0[k];
    cons[ilay - 1] = dmax * (-1.0) + s;
    for (k = 1; k <= ilay - 1; k++)
      cons[ilay - 1] = cons[ilay - 1] - (rho8[ilay - 1] - rho8[k - 1]) * gues0[k - 1] / rho8[ilay - 1];
  }

#ifdef _OPENMP
#pragma omp parallel for schedule(dynamic, 256)
#endif
  for (ipnt = 1; ipnt <= ndeg; ipnt++) { /* pm:361-475 */
    double gues[BEOM_MAXLAY], func[BEOM_MAXLAY], maug[BEOM_MAXLAY][BEOM_MAXLAY + 1], line[BEOM_MAXLAY + 1];
    double hbot, habv, hbel, maxv, resu;
    int il, i, j, kk, l, iter, imax, conv;
    if (o->mk_n[ipnt] < 0.5 || failed) continue;
    hbot = H2(SUBC(ipnt, 1), SUBC(ipnt, 2));
    for (il = nlay; il >= 1; il--) { /* pm:370-380 */
      habv = dmax * P->topl[il - 1];
      hbel = 0.0;
      if (il < nlay)
        for (j = il + 1; j <= nlay; j++) hbel += gues[j - 1];
      gues[il - 1] = hbot - habv - hbel;
      gues[il - 1] = dmax2(gues[il - 1], hsal);
    }
    conv = 0;
    for (iter = 1; iter <= P->itmx; iter++) {
      int all_small = 1, any_le = 0;
      for (i = 1; i <= nlay; i++) { /* pm:383-393 */
        double s = 0.0;
        for (j = 0; j < nlay; j++) s += gues[j];
        func[i - 1] = (hbot - s) + 1.0 / (double)(P->nsal - 1) * hsal * pow3(hsal / gues[i - 1]) + cons[i - 1];
        func[i - 1] = func[i - 1] * (-1.0);
        for (j = 1; j <= i - 1; j++)
          func[i - 1] = func[i - 1] - (rho8[i - 1] - rho8[j - 1]) * gues[j - 1] / rho8[i - 1];
      }
      if (iter == P->itmx) break; /* pm:395-401 */
      for (i = 0; i < nlay; i++)
        if (!(fabs(func[i]) < thre)) all_small = 0;
      if (all_small) { /* pm:403-408 */
        for (i = 1; i <= nlay; i++) A2(o->h_0, ipnt, i) = gues[i - 1];
        conv = 1;
        break;
      }
      for (i = 1; i <= nlay; i++) /* pm:410-419 */
        for (j = 1; j <= nlay; j++) {
          maug[i - 1][j - 1] = dmin2(rho8[i - 1], rho8[j - 1]) / rho8[i - 1];
          if (i == j) maug[i - 1][j - 1] = maug[i - 1][j - 1] + pow4(hsal / gues[j - 1]);
        }
      for (i = 0; i < nlay; i++) maug[i][nlay] = func[i] * (-1.0); /* pm:421 */
      for (kk = 1; kk <= nlay; kk++) { /* pm:426-455 */
        imax = 0;
        maxv = 0.0;
        for (il = kk; il <= nlay; il++)
          if (fabs(maug[il - 1][kk - 1]) > maxv) {
            maxv = fabs(maug[il - 1][kk - 1]);
            imax = il;
          }
        if (imax != kk && imax > 0) {
          memcpy(line, maug[kk - 1], sizeof(double) * (nlay + 1));
          memcpy(maug[kk - 1], maug[imax - 1], sizeof(double) * (nlay + 1));
          memcpy(maug[imax - 1], line, sizeof(double) * (nlay + 1));
        }
        for (il = kk + 1; il <= nlay; il++) {
          for (l = kk; l <= nlay + 1; l++)
            maug[il - 1][l - 1] = maug[il - 1][l - 1] - maug[kk - 1][l - 1] * (maug[il - 1][kk - 1] / maug[kk - 1][kk - 1]);
          maug[il - 1][kk - 1] = 0.0;
        }
      }
      for (il = nlay; il >= 1; il--) { /* pm:459-466 */
        resu = 0.0;
        for (j = il + 1; j <= nlay; j++) resu = resu + maug[il - 1][j - 1] * maug[j - 1][nlay];
        maug[il - 1][nlay] = (maug[il - 1][nlay] - resu) / maug[il - 1][il - 1];
      }
      for (i = 0; i < nlay; i++) { /* pm:468-472 */
        gues[i] = (1.0 - sor) * gues[i] + sor * (maug[i][nlay] + gues[i]);
        if (gues[i] <= thre) any_le = 1;
      }
      if (any_le)
        for (i = 0; i < nlay; i++) gues[i] = dmax2(gues[i], thre);
    }
    if (!conv) failed = ipnt;
  }
  if (failed) {
    snprintf(g_err, sizeof g_err, "calculation of h_layers did not converge at point %d", failed);
    return -1;
  }

  if (P->rgld > 0.5) { /* pm:505-563 */
    double *Osum = dalloc(o->nd1);
    const double dl = P->dl;
    for (ipnt = 0; ipnt <= ndeg; ipnt++) {
      double s = 0.0;
      for (ilay = 1; ilay <= nlay; ilay++) s += A2(o->h_0, ipnt, ilay);
      o->pi_s[ipnt] = (s - o->h_th[ipnt]) * P->grav;
      o->Ow[ipnt] = o->Os[ipnt] = o->Osum_[ipnt] = 0.0;
    }
    for (ipnt = 1; ipnt <= ndeg; ipnt++) {
      const int si = SUBC(ipnt, 1), sj = SUBC(ipnt, 2);
      if (1 < si && si < lm + 1 && 1 < sj && sj < mm + 1) {
        o->Ow[ipnt] = 0.5 * (o->h_th[ipnt] + o->h_th[NEIG(5, ipnt)]) / (dl * dl);
        o->Os[ipnt] = 0.5 * (o->h_th[ipnt] + o->h_th[NEIG(7, ipnt)]) / (dl * dl);
      } else if (si == 1 && 1 < sj && sj < mm + 1) {
        o->Ow[ipnt] = 0.0;
        o->Os[ipnt] = 0.5 * (o->h_th[ipnt] + o->h_th[NEIG(7, ipnt)]) / (dl * dl);
      } else if (1 < si && si < lm + 1 && sj == 1) {
        o->Ow[ipnt] = 0.5 * (o->h_th[ipnt] + o->h_th[NEIG(5, ipnt)]) / (dl * dl);
        o->Os[ipnt] = 0.0;
      }
    }
    for (ipnt = 1; ipnt <= ndeg; ipnt++) {
      const int si = SUBC(ipnt, 1), sj = SUBC(ipnt, 2);
      if (si < lm && sj < mm)
        Osum[ipnt] = o->Ow[ipnt] + o->Ow[NEIG(1, ipnt)] + o->Os[ipnt] + o->Os[NEIG(3, ipnt)];
      else if (si == lm && sj < mm)
        Osum[ipnt] = o->Ow[ipnt] + o->Os[ipnt] + o->Os[NEIG(3, ipnt)];
      else if (sj == mm && si < lm)
        Osum[ipnt] = o->Ow[ipnt] + o->Os[ipnt] + o->Ow[NEIG(1, ipnt)];
      else
        Osum[ipnt] = o->Ow[ipnt] + o->Os[ipnt];
      if (si > 0 && si < lm + 1 && sj > 0 && sj < mm + 1) o->Osum_[ipnt] = 1 / Osum[ipnt];
    }
    free(Osum);
  }
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * index_boundary_points, pm:1060-1240 (nudg is the raw float32 file, (0:lm+1,0:mm+1,3))
 * ------------------------------------------------------------------------------------------- */
static int index_boundary_points(beom_oracle *o, const float *nf) {
  const int lm = o->lm, mm = o->mm, w = lm + 4;
  const double hdry = o->P.hdry;
  const float tiny4 = 1.17549435e-38f; /* tiny(0._r4) */
  int32_t *indc = ialloc((size_t)w * (mm + 4));
  int irep, i, j, i_c = 0, nseg = 0;
#define INDC(i, j) indc[(size_t)((j) + 1) * w + ((i) + 1)]
#define NF(i, j, c) nf[((size_t)((c) - 1) * (mm + 2) + (j)) * (lm + 2) + (i)]
#define NFS(i, j, c) (((i) >= 0 && (i) <= lm + 1 && (j) >= 0 && (j) <= mm + 1) ? NF(i, j, c) : 0.0f)
  for (irep = 1; irep <= 2; irep++) {
    if (irep == 2) {
      o->nseg = nseg;
      o->segm = ialloc((size_t)nseg * 18);
    }
    nseg = 0;
    for (j = 0; j <= mm + 1; j++)
      for (i = 0; i <= lm + 1; i++) {
        if (irep == 1 && any_wet(o, i, j)) INDC(i, j) = ++i_c;
        if (H2(i, j) > hdry && !(H2(i - 1, j) > hdry)) /* western, pm:1106-1133 */
          if (NFS(i, j, ix_u) > tiny4 && NFS(i - 1, j, ix_u) > tiny4 && o->P.xper < 0.5) {
            nseg++;
            if (irep == 2) {
              SEGM(nseg, 1) = INDC(i, j); SEGM(nseg, 2) = i; SEGM(nseg, 3) = j;
              SEGM(nseg, 4) = 1; SEGM(nseg, 6) = 1;
              SEGM(nseg, 7) = INDC(i - 1, j); SEGM(nseg, 8) = i - 1; SEGM(nseg, 9) = j;
              SEGM(nseg, 10) = INDC(i, j); SEGM(nseg, 11) = i; SEGM(nseg, 12) = j;
              SEGM(nseg, 13) = INDC(i + 1, j); SEGM(nseg, 14) = i + 1; SEGM(nseg, 15) = j;
              SEGM(nseg, 16) = INDC(i + 1, j); SEGM(nseg, 17) = i + 1; SEGM(nseg, 18) = j;
            }
          }
        if (!(H2(i, j) > hdry) && H2(i - 1, j) > hdry) /* eastern, pm:1134-1162 */
          if (NFS(i - 1, j, ix_u) > tiny4 && NFS(i, j, ix_u) > tiny4 && o->P.xper < 0.5) {
            nseg++;
            if (irep == 2) {
              SEGM(nseg, 1) = INDC(i, j); SEGM(nseg, 2) = i; SEGM(nseg, 3) = j;
              SEGM(nseg, 4) = 1; SEGM(nseg, 6) = -1;
              SEGM(nseg, 7) = INDC(i, j); SEGM(nseg, 8) = i; SEGM(nseg, 9) = j;
              SEGM(nseg, 10) = INDC(i - 1, j); SEGM(nseg, 11) = i - 1; SEGM(nseg, 12) = j;
              SEGM(nseg, 13) = INDC(i - 1, j); SEGM(nseg, 14) = i - 1; SEGM(nseg, 15) = j;
              SEGM(nseg, 16) = INDC(i - 2, j); SEGM(nseg, 17) = i - 2; SEGM(nseg, 18) = j;
            }
          }
        if (H2(i, j) > hdry && !(H2(i, j - 1) > hdry)) /* southern, pm:1163-1192 */
          if (NFS(i, j, ix_v) > tiny4 && NFS(i, j - 1, ix_v) > tiny4 && o->P.yper < 0.5) {
            nseg++;
            if (irep == 2) {
              SEGM(nseg, 1) = INDC(i, j); SEGM(nseg, 2) = i; SEGM(nseg, 3) = j;
              SEGM(nseg, 5) = 1; SEGM(nseg, 6) = 1;
              SEGM(nseg, 7) = INDC(i, j - 1); SEGM(nseg, 8) = i; SEGM(nseg, 9) = j - 1;
              SEGM(nseg, 10) = INDC(i, j); SEGM(nseg, 11) = i; SEGM(nseg, 12) = j;
              SEGM(nseg, 13) = INDC(i, j + 1); SEGM(nseg, 14) = i; SEGM(nseg, 15) = j + 1;
              SEGM(nseg, 16) = INDC(i, j + 1); SEGM(nseg, 17) = i; SEGM(nseg, 18) = j + 1;
            }
          }
        if (!(H2(i, j) > hdry) && H2(i, j - 1) > hdry) /* northern, pm:1193-1222 */
          if (NFS(i, j - 1, ix_v) > tiny4 && NFS(i, j, ix_v) > tiny4 && o->P.yper < 0.5) {
            nseg++;
            if (irep == 2) {
              SEGM(nseg, 1) = INDC(i, j); SEGM(nseg, 2) = i; SEGM(nseg, 3) = j;
              SEGM(nseg, 5) = 1; SEGM(nseg, 6) = -1;
              SEGM(nseg, 7) = INDC(i, j); SEGM(nseg, 8) = i; SEGM(nseg, 9) = j;
              SEGM(nseg, 10) = INDC(i, j - 1); SEGM(nseg, 11) = i; SEGM(nseg, 12) = j - 1;
              SEGM(nseg, 13) = INDC(i, j - 1); SEGM(nseg, 14) = i; SEGM(nseg, 15) = j - 1;
              SEGM(nseg, 16) = INDC(i, j - 2); SEGM(nseg, 17) = i; SEGM(nseg, 18) = j - 2;
            }
          }
      }
    if (nseg == 0) { /* pm:1226-1231 */
      snprintf(g_err, sizeof g_err, "the nudged open boundary segments could not be identified.");
      free(indc);
      return -1;
    }
  }
#undef INDC
#undef NF
#undef NFS
  free(indc);
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * read_input_file, pm:766-967
 * ------------------------------------------------------------------------------------------- */
static int read_inputs_after_h0(beom_oracle *o, const char *idir) {
  const int lm = o->lm, mm = o->mm, nlay = o->nlay, ndeg = o->ndeg;
  const size_t np = (size_t)(lm + 2) * (mm + 2);
  int err, ipnt, ilay, i, j;
  float *f;
#define F2(i, j) f[(size_t)(j) * (lm + 2) + (i)]
#define F3(i, j, k) f[((size_t)((k) - 1) * (mm + 2) + (j)) * (lm + 2) + (i)]
#define F4(i, j, k, c) f[(((size_t)((c) - 1) * nlay + ((k) - 1)) * (mm + 2) + (j)) * (lm + 2) + (i)]

  o->flag_nudging = 0; /* pm:203 */
  f = read_f32(idir, "nudg", np * 3, &err); /* pm:843-881 */
  if (err) return -1;
  if (f) {
    int any = 0;
    for (ipnt = 1; ipnt <= ndeg; ipnt++) {
      i = SUBC(ipnt, 1);
      j = SUBC(ipnt, 2);
      NUDG(ipnt, ix_n) = (double)F3(i, j, ix_n);
      if (F3(i - 1, j, ix_u) > 1.e-9f && F3(i, j, ix_u) > 1.e-9f)
        NUDG(ipnt, ix_u) = (double)F3(i, j, ix_u) * 0.5 + (double)F3(i - 1, j, ix_u) * 0.5;
      else
        NUDG(ipnt, ix_u) = 0.0;
      if (F3(i, j - 1, ix_v) > 1.e-9f && F3(i, j, ix_v) > 1.e-9f)
        NUDG(ipnt, ix_v) = (double)F3(i, j, ix_v) * 0.5 + (double)F3(i, j - 1, ix_v) * 0.5;
      else
        NUDG(ipnt, ix_v) = 0.0;
    }
    for (ipnt = 0; ipnt < 3 * o->nd1; ipnt++)
      if (o->nudg[ipnt] > 1.e-9) any = 1;
    if (any) {
      o->flag_nudging = 1;
      if (index_boundary_points(o, f)) { free(f); return -1; }
    }
    free(f);
    for (ilay = 1; ilay <= nlay; ilay++)
      for (ipnt = 1; ipnt <= ndeg; ipnt++) FNUD(ipnt, ilay, ix_n) = A2(o->hlay, ipnt, ilay);
  }

  f = read_f32(idir, "init", np * nlay * 3, &err); /* pm:882-910 */
  if (err) return -1;
  if (f) {
    for (ilay = 1; ilay <= nlay; ilay++)
      for (ipnt = 1; ipnt <= ndeg; ipnt++) {
        i = SUBC(ipnt, 1);
        j = SUBC(ipnt, 2);
        if (ilay < nlay)
          FNUD(ipnt, ilay, ix_n) = A2(o->hlay, ipnt, ilay) + (double)F4(i, j, ilay, ix_n) - (double)F4(i, j, ilay + 1, ix_n);
        else
          FNUD(ipnt, ilay, ix_n) = A2(o->hlay, ipnt, ilay) + (double)F4(i, j, ilay, ix_n);
        FNUD(ipnt, ilay, ix_n) = FNUD(ipnt, ilay, ix_n) * o->mk_n[ipnt];
        FNUD(ipnt, ilay, ix_u) = (double)F4(i, j, ilay, ix_u);
        FNUD(ipnt, ilay, ix_v) = (double)F4(i, j, ilay, ix_v);
        if (o->P.rsta < 0.5) {
          A2(o->hlay, ipnt, ilay) = FNUD(ipnt, ilay, ix_n) * o->mk_n[ipnt];
          A2(o->u, ipnt, ilay) = FNUD(ipnt, ilay, ix_u);
          A2(o->v, ipnt, ilay) = FNUD(ipnt, ilay, ix_v);
        }
      }
    free(f);
  }

  f = read_f32(idir, "bodf", (size_t)nlay * 2, &err); /* pm:840-842 */
  if (err) return -1;
  if (f) {
    for (i = 0; i < nlay * 2; i++) o->bodf[i] = (double)f[i];
    free(f);
  }

  f = read_f32(idir, "hdot", np * nlay, &err); /* pm:911-919 */
  if (err) return -1;
  if (f) {
    for (ilay = 1; ilay <= nlay; ilay++)
      for (ipnt = 1; ipnt <= ndeg; ipnt++) A2(o->hdot, ipnt, ilay) = (double)F3(SUBC(ipnt, 1), SUBC(ipnt, 2), ilay);
    free(f);
  }

  f = read_f32(idir, "taus", np * 2, &err); /* pm:920-931 */
  if (err) return -1;
  if (f) {
    memset(o->taus, 0, sizeof(double) * 2 * o->nd1);
    for (ipnt = 1; ipnt <= ndeg; ipnt++) {
      TAUS(ipnt, 1) = (double)F3(SUBC(ipnt, 1), SUBC(ipnt, 2), 1);
      TAUS(ipnt, 2) = (double)F3(SUBC(ipnt, 1), SUBC(ipnt, 2), 2);
    }
    free(f);
  }

  f = read_f32(idir, "tide", np * 2 * 3, &err); /* pm:951-964; shape (2,1,0:lm+1,0:mm+1,3) */
  if (err) return -1;
  if (f) {
    int a, c;
    o->w_ti = (double)f[0];
    for (ipnt = 1; ipnt <= ndeg; ipnt++) {
      i = SUBC(ipnt, 1);
      j = SUBC(ipnt, 2);
      for (c = 1; c <= 3; c++)
        for (a = 1; a <= 2; a++) TIDE(a, ipnt, c) = (double)f[(((size_t)(c - 1) * (mm + 2) + j) * (lm + 2) + i) * 2 + (a - 1)];
    }
    free(f);
  }

  f = read_f32(idir, "fcor", np, &err); /* pm:932-950 */
  if (err) return -1;
  if (f) {
    float s = 0.0f;
    size_t n;
    for (n = 0; n < np; n++) s = s + f[n];
    o->fcor[0] = (double)(s / (float)np);
    for (ipnt = 1; ipnt <= ndeg; ipnt++) {
      i = SUBC(ipnt, 1);
      j = SUBC(ipnt, 2);
      if (i > 0 && j > 0) {
        float t = F2(i, j) * 0.25f;
        t = t + F2(i - 1, j) * 0.25f;
        t = t + F2(i, j - 1) * 0.25f;
        t = t + F2(i - 1, j - 1) * 0.25f;
        o->fcor[ipnt] = (double)t;
      } else
        o->fcor[ipnt] = (double)F2(i, j);
    }
    free(f);
  }
#undef F2
#undef F3
#undef F4
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * read_input_data, pm:105-250 (+ initialize_variables pm:252-307)
 * ------------------------------------------------------------------------------------------- */
beom_oracle *beom_oracle_create(const beom_params *par, const char *idir) {
  beom_oracle *o = (beom_oracle *)calloc(1, sizeof *o);
  int lm, mm, nlay, ndeg, nd1, i, j, ipnt, ilay, err;
  size_t n, n2;
  double dmin, dmaxv, dtd8;
  float *f;
  g_err[0] = 0;
  o->P = *par;
  lm = o->lm = par->lm;
  mm = o->mm = par->mm;
  nlay = o->nlay = par->nlay;
  ndeg = o->ndeg = par->ndeg;
  nd1 = o->nd1 = ndeg + 1;
  if (nlay < 1 || nlay > BEOM_MAXLAY || lm < 1 || mm < 1) {
    snprintf(g_err, sizeof g_err, "bad dimensions");
    free(o);
    return NULL;
  }
  n = (size_t)nd1;
  n2 = n * nlay;
  o->neig = ialloc(n * 8); o->subc = ialloc(n * 2); o->posc = ialloc(n);
  o->h_u = dalloc(n2); o->h_v = dalloc(n2); o->u = dalloc(n2); o->v = dalloc(n2);
  o->UU4 = dalloc(n2); o->VV4 = dalloc(n2); o->delu = dalloc(n2); o->delv = dalloc(n2);
  o->tt3d = dalloc(n2 * 2); o->tb3d = dalloc(n2 * 2); o->tu3d = dalloc(n2 * 2); o->taus = dalloc(n * 2);
  o->rvor = dalloc(n); o->pvor = dalloc(n); o->dive = dalloc(n);
  o->v_cc = dalloc(n2); o->v_ll = dalloc(n2); o->fcor = dalloc(n);
  o->mk_u = dalloc(n); o->mk_v = dalloc(n); o->mk_n = dalloc(n); o->mkpe = dalloc(n); o->mkpi = dalloc(n);
  o->mont = dalloc(n); o->d2hx = dalloc(n); o->d2hy = dalloc(n);
  o->h_bo = dalloc(n); o->h_to = dalloc(n); o->h_th = dalloc(n);
  o->Ow = dalloc(n); o->Os = dalloc(n); o->Osum_ = dalloc(n);
  o->bodf = dalloc((size_t)nlay * 2); o->fnud = dalloc(n2 * 3); o->nudg = dalloc(n * 3);
  o->hdot = dalloc(n2); o->rs_h = dalloc(n2 * 2); o->dmdx = dalloc(n2 * 3); o->dmdy = dalloc(n2 * 3);
  o->tide = dalloc(n * 6);
  o->hlay = dalloc(n2); o->pi_s = dalloc(n); o->h_0 = dalloc(n2);
  o->layt = dalloc(n2); o->layb = dalloc(n2); o->layu = dalloc(n2); o->taub = dalloc(n * 2); o->taum = dalloc(n * 2);
  o->h_2d = dalloc((size_t)(lm + 4) * (mm + 4));

  /* initialize_variables, pm:252-307 */
  for (n = 0; n < n2; n++) o->v_cc[n] = o->v_ll[n] = par->bvis;
  for (n = 0; n < (size_t)nd1; n++) {
    o->fcor[n] = par->f0;
    o->taus[n] = par->tauw[0];
    o->taus[nd1 + n] = par->tauw[1];
  }

  for (j = 1; j <= mm; j++) /* pm:119-121 */
    for (i = 1; i <= lm; i++) H2(i, j) = (par->cext * par->cext) / par->grav;

  f = read_f32(idir, "h_bo", (size_t)(lm + 2) * (mm + 2), &err); /* pm:127, 827-839 */
  if (err) goto fail;
  if (f) {
    float *f2 = NULL;
    if (par->topt > 0.5) {
      f2 = read_f32(idir, "h_to", (size_t)(lm + 2) * (mm + 2), &err);
      if (!f2) { if (!err) snprintf(g_err, sizeof g_err, "topt=1 but h_to.bin is missing"); free(f); goto fail; }
    }
    memset(o->h_2d, 0, sizeof(double) * (size_t)(lm + 4) * (mm + 4));
    for (j = 0; j <= mm + 1; j++)
      for (i = 0; i <= lm + 1; i++) {
        size_t k = (size_t)j * (lm + 2) + i;
        H2(i, j) = f2 ? (double)(f[k] - f2[k]) : (double)f[k];
      }
    free(f);
    free(f2);
    for (n = 0; n < (size_t)(lm + 4) * (mm + 4); n++)
      if (o->h_2d[n] < par->hdry) o->h_2d[n] = 0.0;
    for (j = -1; j <= mm + 2; j++) H2(0, j) = H2(lm + 1, j) = 0.0;
    for (i = -1; i <= lm + 2; i++) H2(i, 0) = H2(i, mm + 1) = 0.0;
  }

  if (index_grid_points(o)) goto fail; /* pm:129 */

  dmin = HUGE_VAL; /* pm:134-135 */
  dmaxv = -HUGE_VAL;
  for (n = 0; n < (size_t)(lm + 4) * (mm + 4); n++) {
    if (o->h_2d[n] > par->hdry && o->h_2d[n] < dmin) dmin = o->h_2d[n];
    if (o->h_2d[n] > dmaxv) dmaxv = o->h_2d[n];
  }
  if (par->ocrp < 0.5 && nlay > 1) { /* pm:137-152 */
    if ((par->topl[nlay - 1] * dmaxv + 10.0 * par->hmin) >= dmin) {
      snprintf(g_err, sizeof g_err, "Please modify topl so that bathymetry is contained within lower layer.");
      goto fail;
    }
  } else if (par->ocrp < 0.5 && nlay == 1) {
    if (dmin <= 10.0 * par->hmin) {
      snprintf(g_err, sizeof g_err, "Please adjust h_bo or hmin so that min(h_bo) > 10. * hmin.");
      goto fail;
    }
  }
  if (par->ocrp < 0.5) { /* pm:154-175 */
    for (ipnt = 1; ipnt <= ndeg; ipnt++)
      if (o->mk_n[ipnt] > 0.5) {
        i = SUBC(ipnt, 1);
        j = SUBC(ipnt, 2);
        for (ilay = nlay; ilay >= 1; ilay--) {
          double habv = 0.0, hbel = 0.0;
          int k;
          if (ilay > 1) habv = dmaxv * par->topl[ilay - 1];
          if (ilay < nlay)
            for (k = ilay + 1; k <= nlay; k++) hbel += A2(o->h_0, ipnt, k);
          A2(o->h_0, ipnt, ilay) = H2(i, j) - habv - hbel;
        }
      }
  }
  if (par->ocrp > 0.5) /* pm:177-183 */
    if (get_equilibrium_thickness_h_0(o)) goto fail;

  for (ilay = 1; ilay <= nlay; ilay++) /* pm:198-200 */
    for (ipnt = 0; ipnt <= ndeg; ipnt++) A2(o->hlay, ipnt, ilay) = A2(o->h_0, ipnt, ilay) * o->mk_n[ipnt];

  if (read_inputs_after_h0(o, idir)) goto fail; /* pm:204-216 */

  { /* pm:223-229 */
    double s = 0.0;
    for (ipnt = 0; ipnt <= ndeg; ipnt++) s += o->fcor[ipnt];
    o->invf = s / (double)nd1;
    if (fabs(o->invf) > 1.25e-5) o->invf = 1.0 / o->invf;
    else o->invf = 0.0;
  }

  dtd8 = par->dt / 24.0 / 3600.0; /* pm:1853-1856 */
  o->nstp = (int)f_nint(par->dt_s / dtd8);
  o->notp = (int)f_nint(par->dt_o / dtd8);
  if (o->notp < 1) o->notp = 1;
  o->n_3d = (int)f_nint(par->dt3d / dtd8);
  if (o->n_3d < 1) o->n_3d = 1;
  o->ramp = 1.0;
  o->gene = 0.0;
  o->tres = 0.0;
  return o;
fail:
  beom_oracle_destroy(o);
  return NULL;
}

void beom_oracle_destroy(beom_oracle *o) {
  if (!o) return;
  free(o->neig); free(o->subc); free(o->posc); free(o->segm);
  free(o->h_u); free(o->h_v); free(o->u); free(o->v); free(o->UU4); free(o->VV4); free(o->delu); free(o->delv);
  free(o->tt3d); free(o->tb3d); free(o->tu3d); free(o->taus);
  free(o->rvor); free(o->pvor); free(o->dive); free(o->v_cc); free(o->v_ll); free(o->fcor);
  free(o->mk_u); free(o->mk_v); free(o->mk_n); free(o->mkpe); free(o->mkpi);
  free(o->mont); free(o->d2hx); free(o->d2hy); free(o->h_bo); free(o->h_to); free(o->h_th);
  free(o->Ow); free(o->Os); free(o->Osum_);
  free(o->bodf); free(o->fnud); free(o->nudg); free(o->hdot); free(o->rs_h); free(o->dmdx); free(o->dmdy); free(o->tide);
  free(o->hlay); free(o->pi_s); free(o->h_0); free(o->h_2d);
  free(o->layt); free(o->layb); free(o->layu); free(o->taub); free(o->taum);
  free(o);
}


/* ---------------------------------------------------------------------------------------------
 * update_u, pm:1422-1503
 * ------------------------------------------------------------------------------------------- */
void beom_oracle_update_u(beom_oracle *o, int ilay) {
  const beom_params *P = &o->P;
  const double i_dl = 1.0 / P->dl, i_r0 = 1.0 / P->rho0, i_r1 = 1.0 / P->rhon[0];
  const double grav = P->grav, dt = P->dt, gene = o->gene, ramp = o->ramp;
  const double del1 = P->del1, del2 = P->del2, gamm = P->gamm, epsi = P->epsi;
  int ipnt;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (ipnt = 1; ipnt <= o->ndeg; ipnt++) {
    const int c3 = NEIG(3, ipnt), c4 = NEIG(4, ipnt), c5 = NEIG(5, ipnt);
    const double mask = o->mk_u[ipnt];
    const double hcen = (A2(o->hlay, c5, ilay) + A2(o->hlay, ipnt, ilay)) / (1.0 + mask);
    const double i__h = 1.0 / (hcen + 1.0 - mask);
    double uold = A2(o->u, ipnt, ilay);
    const double dmd4 = (o->mont[c5] - o->mont[ipnt]) * i_dl * grav * mask;
    const double tauw = 0.5 * (T3(o->tt3d, c5, 1, ilay) + T3(o->tt3d, ipnt, 1, ilay)) * ramp;
    /* vcor (pm:1446-1449) is computed and never used: omitted */
    const double ufor = FNUD(ipnt, ilay, ix_u)
                      + 0.5 * (T3(o->tt3d, ipnt, 2, ilay) + T3(o->tt3d, c5, 2, ilay)) * i_r1 * o->invf * i__h * ramp
                      + ramp * TIDE(1, ipnt, ix_u) * cos(TIDE(2, ipnt, ix_u) - o->w_ti * o->ctim);
    double rhsi = dmd4 * (1.0 - gene)
                + 0.25 * o->pvor[ipnt] * (A2(o->h_v, ipnt, ilay) + A2(o->h_v, c5, ilay))
                + 0.25 * o->pvor[c3] * (A2(o->h_v, c3, ilay) + A2(o->h_v, c4, ilay))
                + tauw * i_r0 * i__h
                - T3(o->tb3d, ipnt, 1, ilay) * i_r0 * i__h
                - T3(o->tu3d, ipnt, 1, ilay) * i_r0 * i__h
                + BODF(ilay, 1)
                + (del1 * dmd4 + del2 * DM(o->dmdx, 3, ipnt, ilay) + gamm * DM(o->dmdx, 2, ipnt, ilay)
                   + epsi * DM(o->dmdx, 1, ipnt, ilay)) * gene;
    if (P->svis > 0.0) { /* pm:1471-1473; real() without kind = single precision */
      const float t4 = (float)(A2(o->UU4, ipnt, ilay) - A2(o->UU4, c5, ilay) + A2(o->VV4, c3, ilay) - A2(o->VV4, ipnt, ilay));
      rhsi = rhsi - P->svis * i_dl * (double)t4 * i__h;
    } else { /* pm:1476-1479 */
      rhsi = rhsi + (A2(o->v_cc, ipnt, ilay) * o->dive[ipnt] - A2(o->v_cc, c5, ilay) * o->dive[c5]) * i_dl
                  - (A2(o->v_ll, c3, ilay) * o->rvor[c3] - A2(o->v_ll, ipnt, ilay) * o->rvor[ipnt]) * i_dl;
    }
    uold = uold + rhsi * mask * dt;
    uold = ufor * NUDG(ipnt, ix_u) + uold * (1.0 - NUDG(ipnt, ix_u));
    A2(o->u, ipnt, ilay) = uold;
    if (P->rgld < 0.5) /* pm:1491-1496 */
      A2(o->h_u, ipnt, ilay) = 0.5 * (uold + fabs(uold)) * (hcen - 0.16667 * o->d2hx[c5])
                             + 0.5 * (uold - fabs(uold)) * (hcen - 0.16667 * o->d2hx[ipnt]);
    DM(o->dmdx, 1, ipnt, ilay) = DM(o->dmdx, 2, ipnt, ilay);
    DM(o->dmdx, 2, ipnt, ilay) = DM(o->dmdx, 3, ipnt, ilay);
    DM(o->dmdx, 3, ipnt, ilay) = dmd4;
  }
}

/* ---------------------------------------------------------------------------------------------
 * update_v, pm:1505-1591
 * ------------------------------------------------------------------------------------------- */
void beom_oracle_update_v(beom_oracle *o, int ilay) {
  const beom_params *P = &o->P;
  const double i_dl = 1.0 / P->dl, i_r0 = 1.0 / P->rho0, i_r1 = 1.0 / P->rhon[0];
  const double grav = P->grav, dt = P->dt, gene = o->gene, ramp = o->ramp;
  const double del1 = P->del1, del2 = P->del2, gamm = P->gamm, epsi = P->epsi;
  int ipnt;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (ipnt = 1; ipnt <= o->ndeg; ipnt++) {
    const int c1 = NEIG(1, ipnt), c7 = NEIG(7, ipnt), c8 = NEIG(8, ipnt);
    const double mask = o->mk_v[ipnt];
    const double hcen = (A2(o->hlay, ipnt, ilay) + A2(o->hlay, c7, ilay)) / (1.0 + mask);
    const double i__h = 1.0 / (hcen + 1.0 - mask);
    double vold = A2(o->v, ipnt, ilay);
    const double dmd4 = (o->mont[c7] - o->mont[ipnt]) * i_dl * grav * mask;
    const double tauw = 0.5 * (T3(o->tt3d, c7, 2, ilay) + T3(o->tt3d, ipnt, 2, ilay)) * ramp;
    const double vfor = FNUD(ipnt, ilay, ix_v)
                      - 0.5 * (T3(o->tt3d, ipnt, 1, ilay) + T3(o->tt3d, c7, 1, ilay)) * i_r1 * o->invf * i__h * ramp
                      + ramp * TIDE(1, ipnt, ix_v) * cos(TIDE(2, ipnt, ix_v) - o->w_ti * o->ctim);
    double rhsi = dmd4 * (1.0 - gene)
                - 0.25 * o->pvor[ipnt] * (A2(o->h_u, ipnt, ilay) + A2(o->h_u, c7, ilay))
                - 0.25 * o->pvor[c1] * (A2(o->h_u, c1, ilay) + A2(o->h_u, c8, ilay))
                + tauw * i_r0 * i__h
                - T3(o->tb3d, ipnt, 2, ilay) * i_r0 * i__h
                - T3(o->tu3d, ipnt, 2, ilay) * i_r0 * i__h
                + BODF(ilay, 2)
                + (del1 * dmd4 + del2 * DM(o->dmdy, 3, ipnt, ilay) + gamm * DM(o->dmdy, 2, ipnt, ilay)
                   + epsi * DM(o->dmdy, 1, ipnt, ilay)) * gene;
    if (P->svis > 0.0) { /* pm:1555-1557 (no real() here: double) */
      rhsi = rhsi - P->svis * i_dl * (A2(o->VV4, c1, ilay) - A2(o->VV4, ipnt, ilay) - A2(o->UU4, ipnt, ilay) + A2(o->UU4, c7, ilay)) * i__h;
    } else { /* pm:1561-1564 */
      rhsi = rhsi + (A2(o->v_cc, ipnt, ilay) * o->dive[ipnt] - A2(o->v_cc, c7, ilay) * o->dive[c7]) * i_dl
                  + (A2(o->v_ll, c1, ilay) * o->rvor[c1] - A2(o->v_ll, ipnt, ilay) * o->rvor[ipnt]) * i_dl;
    }
    vold = vold + rhsi * mask * dt;
    vold = vfor * NUDG(ipnt, ix_v) + vold * (1.0 - NUDG(ipnt, ix_v));
    A2(o->v, ipnt, ilay) = vold;
    if (P->rgld < 0.5) /* pm:1577-1582 */
      A2(o->h_v, ipnt, ilay) = 0.5 * (vold + fabs(vold)) * (hcen - 0.16667 * o->d2hy[c7])
                             + 0.5 * (vold - fabs(vold)) * (hcen - 0.16667 * o->d2hy[ipnt]);
    DM(o->dmdy, 1, ipnt, ilay) = DM(o->dmdy, 2, ipnt, ilay);
    DM(o->dmdy, 2, ipnt, ilay) = DM(o->dmdy, 3, ipnt, ilay);
    DM(o->dmdy, 3, ipnt, ilay) = dmd4;
  }
}

/* ---------------------------------------------------------------------------------------------
 * update_h, pm:1593-1702; variants private_mod1d.f95:1635-1663, private_mod3d.f95:1635-1683,
 * private_modplumenew.f95:1637-1724
 * ------------------------------------------------------------------------------------------- */
static double plume_real(double x) { return (double)(float)x; } /* real(x): single precision */

void beom_oracle_update_h(beom_oracle *o) {
  const beom_params *P = &o->P;
  const int nlay = o->nlay, ndeg = o->ndeg, lm = o->lm, variant = P->variant;
  const double i_dl = 1.0 / P->dl, dt = P->dt, gene = o->gene, ramp = o->ramp, beta = P->beta;
  const double hsal = P->hsal;
  int ilay, ipnt;
  for (ilay = nlay; ilay >= 1; ilay--) {
    const double vecl = (ilay == 1) ? 1.0 : 0.0;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
    for (ipnt = 1; ipnt <= ndeg; ipnt++) {
      const int c1 = NEIG(1, ipnt), c3 = NEIG(3, ipnt);
      double hold = A2(o->hlay, ipnt, ilay);
      double rs_3 = (A2(o->h_u, ipnt, ilay) - A2(o->h_u, c1, ilay)) * i_dl
                  + (A2(o->h_v, ipnt, ilay) - A2(o->h_v, c3, ilay)) * i_dl
                  + A2(o->hdot, ipnt, ilay);
      double rhsi, hfor;
      const double nud = NUDG(ipnt, ix_n);
      rs_3 = rs_3 * o->mk_n[ipnt];
      rhsi = ((1.5 + beta) * rs_3 - (0.5 + 2.0 * beta) * RSH(2, ipnt, ilay) + beta * RSH(1, ipnt, ilay)) * dt * gene
           + rs_3 * dt * (1.0 - gene);
      hold = hold + rhsi;
      hfor = FNUD(ipnt, ilay, ix_n) + ramp * TIDE(1, ipnt, ix_n) * vecl * cos(TIDE(2, ipnt, ix_n) - o->w_ti * o->ctim);
      if (variant == BEOM_VARIANT_STANDARD) { /* pm:1637-1638 */
        A2(o->hlay, ipnt, ilay) = hfor * nud + (1.0 - nud) * hold;
      } else {
        const int si = SUBC(ipnt, 1);
        double h = hold;
        A2(o->hlay, ipnt, ilay) = hold;
        if (variant == BEOM_VARIANT_1D) { /* private_mod1d.f95:1637-1653 */
          const double hfor1 = 800.0, hfor2 = 0.0;
          if (A2(o->hlay, ipnt, 2) > 20 * hsal && si > lm / 2) {
            if (ilay == 1) h = h + 0 * nud + dmax2(hfor1 * nud + (-nud) * h, 0.0);
            else if (ilay == 2) h = h - 0 * nud + dmin2(hfor2 * nud + (-nud) * h, 0.0);
          }
        } else if (variant == BEOM_VARIANT_3D) { /* private_mod3d.f95:1637-1672 */
          const double hfor1 = 0.0, hfor2 = 800.0, hfor3 = 0.0;
          if (A2(o->hlay, ipnt, 3) > 20 * hsal && si > lm / 2) {
            if (ilay == 1) h = h + 0 * nud + dmax2(hfor1 * nud + (-nud) * h, 0.0);
            else if (ilay == 2) h = h + 0 * nud + dmax2(hfor2 * nud + (-nud) * h, 0.0);
            else if (ilay == 3) h = h - 0 * nud + dmin2(hfor3 * nud + (-nud) * h, 0.0);
          } else if (A2(o->hlay, ipnt, 3) < 20 * hsal && si > lm / 2) {
            if (ilay == 1) h = h + 0 * nud + 1 * dmax2(hfor2 * nud + (-nud) * h, 0.0);
            else if (ilay == 2) h = h - 0 * nud + 1 * dmin2(hfor1 * nud + (-nud) * h, 0.0);
          }
        } else { /* BEOM_VARIANT_PLUME, private_modplumenew.f95:1637-1713.
                    Lnud is read at :1643 before it is assigned at :1645 (undefined in the reference);
                    this restatement uses the value it is assigned, 5000, throughout.  Integer
                    divisions 9/10, -1/3, -2/3, 1/3, 6/5 evaluate to 0,0,0,0,1 and -5/3 to -1. */
          const double hfor1 = 0.0, hfor2 = 800.0, hfor3 = 0.0;
          const double Lnud = 5000.0, Wnud = 8000.0, alph = (double)0.13f, Q0 = 250.0, gp0 = (double)0.265f;
          const double B0 = gp0 * Q0, pi = P->pi;
          const int lim = (int)floor((double)lm - Lnud / P->dl);
          if (si > lim) {
            if (P->plum < 0.5) {
              if (ilay == 1) h = h + 0 * nud + dmax2(hfor1 * nud + (-nud) * h, 0.0);
              else if (ilay == 2) h = h + 0 * nud + dmax2(hfor2 * nud + (-nud) * h, 0.0);
              else if (ilay == 3) h = h - 0 * nud + dmin2(hfor3 * nud + (-nud) * h, 0.0);
            } else {
              const double h2 = A2(o->hlay, ipnt, 2), h3 = A2(o->hlay, ipnt, 3);
              const double gp12 = 5 * B0 / (6 * alph) * 1.0 * 1.0 * (1.0 / (h2 + h3));
              const double gp23 = 5 * B0 / (6 * alph) * 1.0 * 1.0 * (1.0 / h3);
              const double w12 = 5 / (6 * alph) * 1.0 * 1.0 * 1.0, w23 = w12;
              const double R12 = 1 * alph * (h3 + h2), R23 = 1 * alph * h3;
              const double Q12 = pi / 2 * w12 * (R12 * R12), Q23 = pi / 2 * w23 * (R23 * R23);
              const double rhop1 = -gp12 * P->rhon[1] / P->grav + P->rhon[1];
              const double rhop2 = -gp23 * P->rhon[1] / P->grav + P->rhon[2];
              const double m1 = dmax2(plume_real((rhop1 - P->rhon[0]) / (P->rhon[1] - P->rhon[0])), 0.0);
              const double m2 = dmax2(plume_real((rhop2 - P->rhon[1]) / (P->rhon[2] - P->rhon[1])), 0.0);
              if (ilay == 1) h = h + dt * Q12 / (Wnud * Lnud) * (1 - m1);
              else if (ilay == 2) h = h + dt * (Q23 - Q12) / (Wnud * Lnud) - dt * Q23 / (Wnud * Lnud) * m2 + dt * Q12 / (Wnud * Lnud) * m1;
              else if (ilay == 3) h = h - dt * Q23 / (Wnud * Lnud) * (1 - m2);
            }
          }
        }
        A2(o->hlay, ipnt, ilay) = h;
        if (si < lm / 2) A2(o->hlay, ipnt, ilay) = hfor * nud + (1.0 - nud) * hold;
      }
      RSH(1, ipnt, ilay) = RSH(2, ipnt, ilay);
      RSH(2, ipnt, ilay) = rs_3;
    }
  }
  if (P->rgld > 0.5) { /* pm:1648-1700: hard-wired to two layers, single-precision correction */
    for (ipnt = 1; ipnt <= ndeg; ipnt++) {
      int k;
      double s = 0.0;
      for (k = 1; k <= nlay; k++) s += A2(o->hlay, ipnt, k);
      A2(o->hlay, ipnt, 1) = A2(o->hlay, ipnt, 1) - (double)(0.5f * (float)(s - o->h_th[ipnt]));
      s = 0.0;
      for (k = 1; k <= nlay; k++) s += A2(o->hlay, ipnt, k);
      A2(o->hlay, ipnt, 2) = A2(o->hlay, ipnt, 2) - (double)(0.5f * (float)(s - o->h_th[ipnt]));
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * surf_pressure, pm:1705-1838
 * ------------------------------------------------------------------------------------------- */
void beom_oracle_surf_pressure(beom_oracle *o) {
  const beom_params *P = &o->P;
  const int ndeg = o->ndeg, nlay = o->nlay, lm = o->lm, mm = o->mm;
  const double dl = P->dl, dt = P->dt, rp = 1.0, pi_tol = 1.e-5;
  const int maxiters = 1000;
  double *pi_rhs = dalloc(o->nd1), *pi_prev = dalloc(o->nd1), *pi_s = o->pi_s;
  double maxdiff;
  int iters, ilay, ipnt;
  for (ilay = nlay; ilay >= 1; ilay--) { /* pm:1722-1752 */
    for (ipnt = 1; ipnt <= ndeg; ipnt++)
      if (SUBC(ipnt, 1) > 1) {
        const int c5 = NEIG(5, ipnt);
        pi_rhs[ipnt] = pi_rhs[ipnt] - A2(o->h_u, ipnt, ilay) / (dl * dt);
        pi_rhs[c5] = pi_rhs[c5] + A2(o->h_u, ipnt, ilay) / (dl * dt);
      }
    for (ipnt = 1; ipnt <= ndeg; ipnt++)
      if (SUBC(ipnt, 2) > 1) {
        const int c7 = NEIG(7, ipnt);
        pi_rhs[ipnt] = pi_rhs[ipnt] - A2(o->h_v, ipnt, ilay) / (dl * dt);
        pi_rhs[c7] = pi_rhs[c7] + A2(o->h_v, ipnt, ilay) / (dl * dt);
      }
  }
  maxdiff = pi_tol + 1; /* pm:1756-1803 */
  iters = 0;
  while (maxdiff > pi_tol && iters < maxiters) {
    maxdiff = 0;
    for (ipnt = 1; ipnt <= ndeg; ipnt++) {
      pi_prev[ipnt] = pi_s[ipnt];
      pi_s[ipnt] = (1 - rp) * pi_s[ipnt] - rp * o->Osum_[ipnt] * pi_rhs[ipnt];
      if (SUBC(ipnt, 1) < lm) { const int c1 = NEIG(1, ipnt); pi_s[ipnt] = pi_s[ipnt] + rp * o->Osum_[ipnt] * o->Ow[c1] * pi_s[c1]; }
      if (SUBC(ipnt, 2) < mm) { const int c3 = NEIG(3, ipnt); pi_s[ipnt] = pi_s[ipnt] + rp * o->Osum_[ipnt] * o->Os[c3] * pi_s[c3]; }
      if (SUBC(ipnt, 1) > 1) { const int c5 = NEIG(5, ipnt); pi_s[ipnt] = pi_s[ipnt] + rp * o->Osum_[ipnt] * o->Ow[ipnt] * pi_s[c5]; }
      if (SUBC(ipnt, 2) > 1) { const int c7 = NEIG(7, ipnt); pi_s[ipnt] = pi_s[ipnt] + rp * o->Osum_[ipnt] * o->Os[ipnt] * pi_s[c7]; }
    }
    for (ipnt = 1; ipnt <= ndeg; ipnt++) {
      const double diff = fabs(pi_s[ipnt] - pi_prev[ipnt]);
      if (diff > maxdiff) maxdiff = diff;
    }
    iters++;
  }
  for (ilay = 1; ilay <= nlay; ilay++) /* pm:1807-1819 */
    for (ipnt = 1; ipnt <= ndeg; ipnt++)
      if (SUBC(ipnt, 1) > 1 && SUBC(ipnt, 1) < lm + 1) {
        const int c5 = NEIG(5, ipnt);
        A2(o->u, ipnt, ilay) = A2(o->u, ipnt, ilay) - dt / dl * pi_s[ipnt];
        A2(o->u, ipnt, ilay) = A2(o->u, ipnt, ilay) + dt / dl * pi_s[c5];
      }
  for (ilay = 1; ilay <= nlay; ilay++) /* pm:1823-1833 */
    for (ipnt = 1; ipnt <= ndeg; ipnt++)
      if (SUBC(ipnt, 2) > 1 && SUBC(ipnt, 2) < mm + 1) {
        const int c7 = NEIG(7, ipnt);
        A2(o->v, ipnt, ilay) = A2(o->v, ipnt, ilay) - dt / dl * pi_s[ipnt];
        A2(o->v, ipnt, ilay) = A2(o->v, ipnt, ilay) + dt / dl * pi_s[c7];
      }
  free(pi_rhs);
  free(pi_prev);
}

/* ---------------------------------------------------------------------------------------------
 * distribute_stress, pm:1921-2149 (private_mod1d.f95:2046 changes one threshold)
 * ------------------------------------------------------------------------------------------- */
static void drag_stress(beom_oracle *o, double coef, int from_bottom, double thr_mult, double *tau) {
  /* pm:2015-2049 (bottom) and pm:2075-2109 (top) */
  const beom_params *P = &o->P;
  const int nlay = o->nlay;
  const double hs_8 = P->hsal, qdrg = P->qdrg;
  int ipnt;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (ipnt = 0; ipnt <= o->ndeg; ipnt++) {
    int ilay = from_bottom ? nlay : 1, klay;
    int c1, c3, c4, c5, c7, c8;
    double vatu, uatv, uu, vv;
    if (P->ocrp > 0.5) {
      if (from_bottom) {
        for (klay = nlay; klay >= 1; klay--)
          if (A2(o->hlay, ipnt, klay) > (thr_mult * hs_8)) { ilay = klay; break; }
      } else {
        for (klay = 1; klay <= nlay; klay++)
          if (A2(o->hlay, ipnt, klay) > (2.0 * hs_8)) { ilay = klay; break; }
      }
    }
    c1 = NEIG(1, ipnt); c3 = NEIG(3, ipnt); c4 = NEIG(4, ipnt);
    c5 = NEIG(5, ipnt); c7 = NEIG(7, ipnt); c8 = NEIG(8, ipnt);
    vatu = 0.25 * A2(o->v, ipnt, ilay) + 0.25 * A2(o->v, c3, ilay) + 0.25 * A2(o->v, c4, ilay) + 0.25 * A2(o->v, c5, ilay);
    uatv = 0.25 * A2(o->u, ipnt, ilay) + 0.25 * A2(o->u, c1, ilay) + 0.25 * A2(o->u, c7, ilay) + 0.25 * A2(o->u, c8, ilay);
    uu = A2(o->u, ipnt, ilay);
    vv = A2(o->v, ipnt, ilay);
    TAUB(tau, ipnt, 1) = uu * coef * P->rhon[ilay - 1] * (qdrg * sqrt(uu * uu + vatu * vatu) + 1.0 - qdrg);
    TAUB(tau, ipnt, 2) = vv * coef * P->rhon[ilay - 1] * (qdrg * sqrt(vv * vv + uatv * uatv) + 1.0 - qdrg);
  }
}

void beom_oracle_distribute_stress(beom_oracle *o) {
  const beom_params *P = &o->P;
  const int nlay = o->nlay, ndeg = o->ndeg;
  const double hsal = P->hsal, hsbl = P->hsbl, hbbl = P->hbbl;
  int ilay, ipnt, klay, any_taus = 0;
  size_t n;
  for (n = 0; n < (size_t)2 * o->nd1; n++)
    if (fabs(o->taus[n]) > 1.e-7) { any_taus = 1; break; }

  if (any_taus && P->ocrp > 0.5) { /* pm:1945-1959 */
    for (ilay = 1; ilay <= nlay; ilay++)
#ifdef _OPENMP
#pragma omp parallel for private(klay) schedule(static)
#endif
      for (ipnt = 0; ipnt <= ndeg; ipnt++) {
        double hcum = 0.0, sofar = 0.0;
        A2(o->layt, ipnt, ilay) = 0.0;
        for (klay = 1; klay <= ilay; klay++) sofar += A2(o->layt, ipnt, klay);
        for (klay = 1; klay <= ilay; klay++) hcum = hcum + dmax2(0.0, A2(o->hlay, ipnt, klay) - 1.5 * hsal);
        A2(o->layt, ipnt, ilay) = dmin2(hcum, hsbl) / hsbl - sofar;
        A2(o->layt, ipnt, ilay) = dmax2(A2(o->layt, ipnt, ilay), 0.0);
      }
  } else if (any_taus) { /* pm:1960-1967 */
    for (ipnt = 0; ipnt <= ndeg; ipnt++) {
      for (ilay = 1; ilay <= nlay; ilay++) A2(o->layt, ipnt, ilay) = 0.0;
      A2(o->layt, ipnt, 1) = 1.0;
    }
  }

  if (P->bdrg > 1.e-7 && P->ocrp > 0.5) { /* pm:1969-1980 */
    for (ilay = nlay; ilay >= 1; ilay--)
#ifdef _OPENMP
#pragma omp parallel for private(klay) schedule(static)
#endif
      for (ipnt = 0; ipnt <= ndeg; ipnt++) {
        double hcum = 0.0, sofar = 0.0;
        A2(o->layb, ipnt, ilay) = 0.0;
        for (klay = ilay; klay <= nlay; klay++) sofar += A2(o->layb, ipnt, klay);
        for (klay = ilay; klay <= nlay; klay++) hcum += A2(o->hlay, ipnt, klay);
        A2(o->layb, ipnt, ilay) = dmin2(hcum, hbbl) / hbbl - sofar;
        A2(o->layb, ipnt, ilay) = dmax2(A2(o->layb, ipnt, ilay), 0.0);
      }
  } else if (P->bdrg > 1.e-7) { /* pm:1981-1988 */
    for (ipnt = 0; ipnt <= ndeg; ipnt++) {
      for (ilay = 1; ilay <= nlay; ilay++) A2(o->layb, ipnt, ilay) = 0.0;
      A2(o->layb, ipnt, nlay) = 1.0;
    }
  }

  if (P->tdrg > 1.e-7 && P->ocrp > 0.5) { /* pm:1991-2005 */
    for (ilay = 1; ilay <= nlay; ilay++)
#ifdef _OPENMP
#pragma omp parallel for private(klay) schedule(static)
#endif
      for (ipnt = 0; ipnt <= ndeg; ipnt++) {
        double hcum = 0.0, sofar = 0.0;
        A2(o->layu, ipnt, ilay) = 0.0;
        for (klay = 1; klay <= ilay; klay++) sofar += A2(o->layu, ipnt, klay);
        for (klay = 1; klay <= ilay; klay++) hcum = hcum + dmax2(0.0, A2(o->hlay, ipnt, klay) - 1.5 * hsal);
        A2(o->layu, ipnt, ilay) = dmin2(hcum, hsbl) / hsbl - sofar;
        A2(o->layu, ipnt, ilay) = dmax2(A2(o->layu, ipnt, ilay), 0.0);
      }
  } else if (P->tdrg > 1.e-7) { /* pm:2006-2013 */
    for (ipnt = 0; ipnt <= ndeg; ipnt++) {
      for (ilay = 1; ilay <= nlay; ilay++) A2(o->layu, ipnt, ilay) = 0.0;
      A2(o->layu, ipnt, 1) = 1.0;
    }
  }

  if (P->bdrg > 1.e-7) { /* pm:2015-2072 */
    drag_stress(o, P->bdrg, 1, P->variant == BEOM_VARIANT_1D ? 0.0 : 2.0, o->taub);
    for (ilay = 1; ilay <= nlay; ilay++)
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
      for (ipnt = 1; ipnt <= ndeg; ipnt++) {
        const int c5 = NEIG(5, ipnt), c7 = NEIG(7, ipnt);
        T3(o->tb3d, ipnt, 1, ilay) = TAUB(o->taub, ipnt, 1) * 0.5 * (A2(o->layb, ipnt, ilay) + A2(o->layb, c5, ilay));
        T3(o->tb3d, ipnt, 2, ilay) = TAUB(o->taub, ipnt, 2) * 0.5 * (A2(o->layb, ipnt, ilay) + A2(o->layb, c7, ilay));
      }
  }
  if (P->tdrg > 1.e-7) { /* pm:2075-2134 */
    drag_stress(o, P->tdrg, 0, 2.0, o->taum);
    for (ilay = 1; ilay <= nlay; ilay++)
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
      for (ipnt = 1; ipnt <= ndeg; ipnt++) {
        const int c5 = NEIG(5, ipnt), c7 = NEIG(7, ipnt);
        T3(o->tu3d, ipnt, 1, ilay) = TAUB(o->taum, ipnt, 1) * 0.5 * (A2(o->layu, ipnt, ilay) + A2(o->layu, c5, ilay));
        T3(o->tu3d, ipnt, 2, ilay) = TAUB(o->taum, ipnt, 2) * 0.5 * (A2(o->layu, ipnt, ilay) + A2(o->layu, c7, ilay));
      }
  }
  if (any_taus) { /* pm:2136-2146 */
    for (ilay = 1; ilay <= nlay; ilay++)
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
      for (ipnt = 1; ipnt <= ndeg; ipnt++) {
        T3(o->tt3d, ipnt, 1, ilay) = TAUS(ipnt, 1) * A2(o->layt, ipnt, ilay);
        T3(o->tt3d, ipnt, 2, ilay) = TAUS(ipnt, 2) * A2(o->layt, ipnt, ilay);
      }
  }
}

/* ---------------------------------------------------------------------------------------------
 * update_mont_rvor_pvor_dive_kine, pm:2318-2439
 * ------------------------------------------------------------------------------------------- */
void beom_oracle_update_mont(beom_oracle *o, int ilay) {
  const beom_params *P = &o->P;
  const int nlay = o->nlay;
  const double i_dl = 1.0 / P->dl, i_gr = 1.0 / P->grav, i_ns = 1.0 / (double)(P->nsal - 1);
  const double hs_8 = P->hsal, hmin = P->hmin, hsal = P->hsal, ocrp = P->ocrp, uadv = P->uadv;
  double i_rn[BEOM_MAXLAY];
  int ipnt, k;
  for (k = 0; k < nlay; k++) i_rn[k] = 1.0 / P->rhon[k];
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (ipnt = 1; ipnt <= o->ndeg; ipnt++) {
    const int c1 = NEIG(1, ipnt), c3 = NEIG(3, ipnt), c5 = NEIG(5, ipnt), c6 = NEIG(6, ipnt), c7 = NEIG(7, ipnt);
    const double u_le = A2(o->u, ipnt, ilay), u_ri = A2(o->u, c1, ilay);
    const double v_bo = A2(o->v, ipnt, ilay), v_to = A2(o->v, c3, ilay);
    double mpot, hcol, have;
    int i;
    mpot = A2(o->hlay, ipnt, ilay) + hmin * (1.0 - o->mk_n[ipnt]); /* pm:2351-2353 */
    mpot = pow3(hs_8 / mpot);
    mpot = mpot * (-ocrp * i_ns * hsal * o->mk_n[ipnt]);
    mpot = mpot - o->h_to[ipnt]; /* pm:2356 */
    for (i = 1; i <= ilay - 1; i++) /* pm:2357-2361 */
      mpot = mpot - (P->rhon[ilay - 1] - P->rhon[i - 1]) * i_rn[ilay - 1] * A2(o->hlay, ipnt, i);
    if (P->rgld < 0.5) { /* pm:2365-2375 */
      hcol = 0.0;
      for (i = 1; i <= nlay; i++) hcol = hcol + A2(o->hlay, ipnt, i);
      mpot = hcol - o->h_th[ipnt] + mpot;
    }
    o->mont[ipnt] = mpot + 0.25 * uadv * i_gr * (u_ri * u_ri + u_le * u_le + v_to * v_to + v_bo * v_bo); /* pm:2380-2383 */
    o->rvor[ipnt] = (v_bo - A2(o->v, c5, ilay) - u_le + A2(o->u, c7, ilay)) * i_dl * o->mkpe[ipnt]; /* pm:2388 */
    o->d2hx[ipnt] = (A2(o->hlay, c1, ilay) + A2(o->hlay, c5, ilay) - A2(o->hlay, ipnt, ilay) * 2.0)
                  * o->mk_n[c1] * o->mk_n[c5] * o->mk_n[ipnt]; /* pm:2394-2397 */
    o->d2hy[ipnt] = (A2(o->hlay, c3, ilay) + A2(o->hlay, c7, ilay) - A2(o->hlay, ipnt, ilay) * 2.0)
                  * o->mk_n[c3] * o->mk_n[c7] * o->mk_n[ipnt]; /* pm:2399-2402 */
    if (ocrp > 0.5) { /* pm:2404-2416 */
      if (A2(o->hlay, c1, ilay) < 2.0 * hs_8 || A2(o->hlay, c5, ilay) < 2.0 * hs_8 || A2(o->hlay, ipnt, ilay) < 2.0 * hs_8)
        o->d2hx[ipnt] = 0.0;
      if (A2(o->hlay, c3, ilay) < 2.0 * hs_8 || A2(o->hlay, c7, ilay) < 2.0 * hs_8 || A2(o->hlay, ipnt, ilay) < 2.0 * hs_8)
        o->d2hy[ipnt] = 0.0;
    }
    have = A2(o->hlay, ipnt, ilay) + A2(o->hlay, c5, ilay) + A2(o->hlay, c6, ilay) + A2(o->hlay, c7, ilay); /* pm:2421 */
    o->pvor[ipnt] = (o->fcor[ipnt] + o->rvor[ipnt] * uadv) * o->mkpi[ipnt]
                  * (o->mk_n[ipnt] + o->mk_n[c5] + o->mk_n[c6] + o->mk_n[c7]) / have; /* pm:2426-2433 */
    o->dive[ipnt] = (u_ri - u_le + v_to - v_bo) * i_dl; /* pm:2435 */
  }
}

/* ---------------------------------------------------------------------------------------------
 * update_viscosity, pm:2441-2611
 * ------------------------------------------------------------------------------------------- */
void beom_oracle_update_viscosity(beom_oracle *o, int ilay) {
  const beom_params *P = &o->P;
  const double dvis = P->dvis, dl = P->dl, bvis = P->bvis, svis = P->svis;
  const int lm = o->lm, mm = o->mm;
  int ipnt;
#ifdef _OPENMP
#pragma omp parallel for schedule(static)
#endif
  for (ipnt = 1; ipnt <= o->ndeg; ipnt++) {
    const int c1 = NEIG(1, ipnt), c2 = NEIG(2, ipnt), c3 = NEIG(3, ipnt), c5 = NEIG(5, ipnt), c6 = NEIG(6, ipnt), c7 = NEIG(7, ipnt);
    const double r_bl = o->rvor[ipnt], r_br = o->rvor[c1], r_tr = o->rvor[c2], r_tl = o->rvor[c3], rbll = o->rvor[c5], rbbl = o->rvor[c7];
    const double d_cc = o->dive[ipnt], d_ri = o->dive[c1], d_to = o->dive[c3], d_le = o->dive[c5], d_bl = o->dive[c6], d_bo = o->dive[c7];
    double t;
    t = (r_br - r_bl) * (r_br - r_bl) + (r_bl - rbll) * (r_bl - rbll) + (r_tl - r_bl) * (r_tl - r_bl) + (r_bl - rbbl) * (r_bl - rbbl)
      + (d_cc - d_le) * (d_cc - d_le) + (d_bo - d_bl) * (d_bo - d_bl) + (d_cc - d_bo) * (d_cc - d_bo) + (d_le - d_bl) * (d_le - d_bl);
    A2(o->v_ll, ipnt, ilay) = sqrt(t) * dvis * dl * dl + bvis; /* pm:2477-2489 */
    t = (r_br - r_bl) * (r_br - r_bl) + (r_tr - r_tl) * (r_tr - r_tl) + (r_tl - r_bl) * (r_tl - r_bl) + (r_tr - r_br) * (r_tr - r_br)
      + (d_ri - d_cc) * (d_ri - d_cc) + (d_cc - d_le) * (d_cc - d_le) + (d_to - d_cc) * (d_to - d_cc) + (d_cc - d_bo) * (d_cc - d_bo);
    A2(o->v_cc, ipnt, ilay) = sqrt(t) * dvis * dl * dl + bvis; /* pm:2492-2502 */
    if (svis > 0.0) { /* pm:2508-2550 */
      double du = 0.0, dv = 0.0;
      if (o->mk_u[ipnt] > 0.5) {
        du = du + 1.0 / (dl * dl) * (o->mk_u[c1] * A2(o->u, c1, ilay) + o->mk_u[c3] * A2(o->u, c3, ilay) + o->mk_u[c5] * A2(o->u, c5, ilay) + o->mk_u[c7] * A2(o->u, c7, ilay));
        du = du - 1.0 / (dl * dl) * (o->mk_u[c1] + o->mk_u[c3] + o->mk_u[c5] + o->mk_u[c7]) * A2(o->u, ipnt, ilay);
      }
      if (o->mk_v[ipnt] > 0.5) {
        dv = dv + 1.0 / (dl * dl) * (o->mk_v[c1] * A2(o->v, c1, ilay) + o->mk_v[c3] * A2(o->v, c3, ilay) + o->mk_v[c5] * A2(o->v, c5, ilay) + o->mk_v[c7] * A2(o->v, c7, ilay));
        dv = dv - 1.0 / (dl * dl) * (o->mk_v[c1] + o->mk_v[c3] + o->mk_v[c5] + o->mk_v[c7]) * A2(o->v, ipnt, ilay);
      }
      A2(o->delu, ipnt, ilay) = du;
      A2(o->delv, ipnt, ilay) = dv;
    }
  }
  if (svis > 0.0) { /* pm:2556-2599 (serial in the reference) */
    for (ipnt = 1; ipnt <= o->ndeg; ipnt++) {
      const int c1 = NEIG(1, ipnt), c3 = NEIG(3, ipnt), c5 = NEIG(5, ipnt), c6 = NEIG(6, ipnt), c7 = NEIG(7, ipnt);
      const double h = A2(o->hlay, ipnt, ilay), i_dl = 1.0 / dl;
      /* real(...) without kind: the numerator is rounded to single precision (pm:2565) */
      const double hh_q = (double)(float)(h + o->mk_n[c5] * A2(o->hlay, c5, ilay) + o->mk_n[c6] * A2(o->hlay, c6, ilay) + o->mk_n[c7] * A2(o->hlay, c7, ilay))
                        / (1.0 + o->mk_n[c5] + o->mk_n[c6] + o->mk_n[c7]);
      double U4 = 0.0, V4 = 0.0;
      U4 = U4 - i_dl * h * A2(o->delu, ipnt, ilay) + i_dl * h * A2(o->delv, ipnt, ilay);
      V4 = V4 + i_dl * hh_q * A2(o->delu, ipnt, ilay) + i_dl * hh_q * A2(o->delv, ipnt, ilay);
      if (SUBC(ipnt, 1) <= lm - 1) U4 = U4 + i_dl * h * A2(o->delu, c1, ilay);
      if (SUBC(ipnt, 2) <= mm - 1) U4 = U4 - i_dl * h * A2(o->delv, c3, ilay);
      if (SUBC(ipnt, 1) > 1) V4 = V4 - i_dl * hh_q * A2(o->delv, c5, ilay);
      if (SUBC(ipnt, 2) > 1) V4 = V4 - i_dl * hh_q * A2(o->delu, c7, ilay);
      if (o->mk_u[ipnt] * o->mk_v[ipnt] < 0.5) V4 = 0.0;
      A2(o->UU4, ipnt, ilay) = U4;
      A2(o->VV4, ipnt, ilay) = V4;
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * no_gradient_obc, pm:2613-2679
 * ------------------------------------------------------------------------------------------- */
void beom_oracle_no_gradient_obc(beom_oracle *o, int ilay) {
  int iseg, ipnt;
  for (iseg = 1; iseg <= o->nseg; iseg++) {
    ipnt = SEGM(iseg, 10);
    if (SEGM(iseg, 5) == 1) {
      if (o->mk_u[ipnt] > 0.5) {
        const int q = SEGM(iseg, 16);
        A2(o->u, ipnt, ilay) = A2(o->u, q, ilay) - FNUD(q, ilay, ix_u) + FNUD(ipnt, ilay, ix_u);
        A2(o->h_u, ipnt, ilay) = A2(o->u, ipnt, ilay) * (A2(o->hlay, ipnt, ilay) + A2(o->hlay, NEIG(5, ipnt), ilay)) / (1.0 + o->mk_u[ipnt]);
      }
    } else if (SEGM(iseg, 4) == 1) {
      if (o->mk_v[ipnt] > 0.5) {
        const int q = SEGM(iseg, 16);
        A2(o->v, ipnt, ilay) = A2(o->v, q, ilay) - FNUD(q, ilay, ix_v) + FNUD(ipnt, ilay, ix_v);
        A2(o->h_v, ipnt, ilay) = A2(o->v, ipnt, ilay) * (A2(o->hlay, ipnt, ilay) + A2(o->hlay, NEIG(7, ipnt), ilay)) / (1.0 + o->mk_v[ipnt]);
      }
    }
  }
  for (iseg = 1; iseg <= o->nseg; iseg++) {
    ipnt = SEGM(iseg, 1);
    if (SEGM(iseg, 5) == 1) {
      const int q = SEGM(iseg, 13);
      A2(o->v, ipnt, ilay) = A2(o->v, q, ilay) - FNUD(q, ilay, ix_v) + FNUD(ipnt, ilay, ix_v);
      A2(o->h_v, ipnt, ilay) = A2(o->v, ipnt, ilay) * (A2(o->hlay, ipnt, ilay) + A2(o->hlay, NEIG(7, ipnt), ilay)) / (1.0 + o->mk_v[ipnt]);
    } else if (SEGM(iseg, 4) == 1) {
      const int q = SEGM(iseg, 13);
      A2(o->u, ipnt, ilay) = A2(o->u, q, ilay) - FNUD(q, ilay, ix_u) + FNUD(ipnt, ilay, ix_u);
      A2(o->h_u, ipnt, ilay) = A2(o->u, ipnt, ilay) * (A2(o->hlay, ipnt, ilay) + A2(o->hlay, NEIG(5, ipnt), ilay)) / (1.0 + o->mk_u[ipnt]);
    }
  }
}

/* ---------------------------------------------------------------------------------------------
 * first_three_timesteps pm:2151-2223, gener_forward_backward pm:2225-2316
 * ------------------------------------------------------------------------------------------- */
static void centred_fluxes(beom_oracle *o) { /* pm:2166-2177, 2208-2219 */
  int ilay, ipnt;
  for (ilay = 1; ilay <= o->nlay; ilay++)
    for (ipnt = 1; ipnt <= o->ndeg; ipnt++) {
      const int c5 = NEIG(5, ipnt), c7 = NEIG(7, ipnt);
      A2(o->h_u, ipnt, ilay) = A2(o->u, ipnt, ilay) * (A2(o->hlay, ipnt, ilay) + A2(o->hlay, c5, ilay)) / (1.0 + o->mk_u[ipnt]);
      A2(o->h_v, ipnt, ilay) = A2(o->v, ipnt, ilay) * (A2(o->hlay, ipnt, ilay) + A2(o->hlay, c7, ilay)) / (1.0 + o->mk_v[ipnt]);
    }
}
static void upstream_fluxes(beom_oracle *o) { /* pm:2238-2256, 2293-2310: d2hx/d2hy of the last layer processed */
  int ilay, ipnt;
  for (ilay = 1; ilay <= o->nlay; ilay++)
    for (ipnt = 1; ipnt <= o->ndeg; ipnt++) {
      const int c5 = NEIG(5, ipnt), c7 = NEIG(7, ipnt);
      double mask = o->mk_u[ipnt], uu = A2(o->u, ipnt, ilay), vv = A2(o->v, ipnt, ilay);
      double hcen = (A2(o->hlay, c5, ilay) + A2(o->hlay, ipnt, ilay)) / (1.0 + mask);
      A2(o->h_u, ipnt, ilay) = 0.5 * (uu + fabs(uu)) * (hcen - 0.16667 * o->d2hx[c5]) + 0.5 * (uu - fabs(uu)) * (hcen - 0.16667 * o->d2hx[ipnt]);
      mask = o->mk_v[ipnt];
      hcen = (A2(o->hlay, c7, ilay) + A2(o->hlay, ipnt, ilay)) / (1.0 + mask);
      A2(o->h_v, ipnt, ilay) = 0.5 * (vv + fabs(vv)) * (hcen - 0.16667 * o->d2hy[c7]) + 0.5 * (vv - fabs(vv)) * (hcen - 0.16667 * o->d2hy[ipnt]);
    }
}

static void first_three_timesteps(beom_oracle *o, int tstp) {
  int ilay;
  centred_fluxes(o);
  beom_oracle_update_h(o);
  for (ilay = 1; ilay <= o->nlay; ilay++) {
    beom_oracle_update_mont(o, ilay);
    beom_oracle_update_viscosity(o, ilay);
    if (tstp % 2 == 0) { beom_oracle_update_u(o, ilay); beom_oracle_update_v(o, ilay); }
    else               { beom_oracle_update_v(o, ilay); beom_oracle_update_u(o, ilay); }
    if (o->flag_nudging && o->P.mcbc < 0.5) beom_oracle_no_gradient_obc(o, ilay);
  }
  if (o->P.rgld > 0.5) { centred_fluxes(o); beom_oracle_surf_pressure(o); }
}

static void gener_forward_backward(beom_oracle *o, int tstp, int upst) {
  int ilay;
  if (o->P.rgld > 0.5) upstream_fluxes(o);
  beom_oracle_update_h(o);
  for (ilay = 1; ilay <= o->nlay; ilay++) {
    beom_oracle_update_mont(o, ilay);
    if ((o->P.dvis > 1.e-3 && upst) || o->P.svis > 0) beom_oracle_update_viscosity(o, ilay);
    if (tstp % 2 == 0) { beom_oracle_update_u(o, ilay); beom_oracle_update_v(o, ilay); }
    else               { beom_oracle_update_v(o, ilay); beom_oracle_update_u(o, ilay); }
    if (o->flag_nudging && o->P.mcbc < 0.5) beom_oracle_no_gradient_obc(o, ilay);
  }
  if (o->P.rgld > 0.5) { upstream_fluxes(o); beom_oracle_surf_pressure(o); }
}

/* ---------------------------------------------------------------------------------------------
 * integrate_time, pm:1840-1919
 * ------------------------------------------------------------------------------------------- */
int beom_oracle_advance(beom_oracle *o, int tstp0, int tstp1) {
  const beom_params *P = &o->P;
  const double dtd8 = P->dt / 24.0 / 3600.0;
  int tstp;
  for (tstp = tstp0; tstp <= tstp1; tstp++) {
    o->ctim = o->tres + dtd8 * (double)tstp;
    if (tstp <= 3) {
      if (tstp == 1) { /* pm:1858-1866 */
        o->ramp = 1.0;
        o->gene = 0.0;
        beom_oracle_distribute_stress(o);
        if (P->rsta < 0.5 && o->ctim < P->dt_r) o->ramp = o->ctim / P->dt_r;
      }
      first_three_timesteps(o, tstp);
      if (tstp == 3) { /* pm:1877-1884 */
        o->gene = P->g_fb;
        if (o->gene > 0.5 && P->rgld > 0.5) o->gene = 0.0;
      }
    } else { /* pm:1886-1906 */
      const int upst = (tstp % o->n_3d == 0);
      if (tstp == 4 && tstp0 == 4) { /* entering mid-run: same switch as after step 3 */
        o->gene = P->g_fb;
        if (o->gene > 0.5 && P->rgld > 0.5) o->gene = 0.0;
      }
      if (upst) beom_oracle_distribute_stress(o);
      o->ramp = 1.0;
      if (P->rsta < 0.5 && o->ctim < P->dt_r) o->ramp = o->ctim / P->dt_r;
      gener_forward_backward(o, tstp, upst);
    }
  }
  return 0;
}

void beom_oracle_counts(const beom_oracle *o, int *nstp, int *notp, int *n_3d) {
  if (nstp) *nstp = o->nstp;
  if (notp) *notp = o->notp;
  if (n_3d) *n_3d = o->n_3d;
}
void beom_oracle_set_scalars(beom_oracle *o, double ctim, double ramp, double gene) {
  o->ctim = ctim; o->ramp = ramp; o->gene = gene;
}

/* ---------------------------------------------------------------------------------------------
 * write_array records, pm:2817-2975 (float32, ndeg x nlay, without the sentinel)
 * ------------------------------------------------------------------------------------------- */
int beom_oracle_record(beom_oracle *o, const char *var, float *out) {
  const beom_params *P = &o->P;
  const int ndeg = o->ndeg, nlay = o->nlay;
  int ilay, ipnt, i;
#define R4(ip, il) out[(size_t)((il) - 1) * ndeg + ((ip) - 1)]
  if (!strcmp(var, "eta_")) { /* pm:2848-2875; h_0 read back from its float32 file */
    for (ilay = nlay; ilay >= 2; ilay--)
      for (ipnt = 1; ipnt <= ndeg; ipnt++) {
        const double h0 = (double)(float)A2(o->h_0, ipnt, ilay);
        if (ilay == nlay) R4(ipnt, ilay) = (float)(A2(o->hlay, ipnt, ilay) - h0);
        else R4(ipnt, ilay) = (float)(A2(o->hlay, ipnt, ilay) - h0 + (double)R4(ipnt, ilay + 1));
      }
    if (P->rgld < 0.5) {
      for (ipnt = 1; ipnt <= ndeg; ipnt++) {
        const double h0 = (double)(float)A2(o->h_0, ipnt, 1);
        if (nlay > 1) R4(ipnt, 1) = (float)(A2(o->hlay, ipnt, 1) - h0 + (double)R4(ipnt, 2));
        else R4(ipnt, 1) = (float)(A2(o->hlay, ipnt, 1) - h0 + 0.0); /* ior4(ipnt,2) does not exist for nlay=1: see README */
      }
    } else
      for (ipnt = 1; ipnt <= ndeg; ipnt++) R4(ipnt, 1) = (float)o->pi_s[ipnt];
  } else if (!strcmp(var, "u___") || !strcmp(var, "v___")) { /* pm:2880-2883 */
    const double *x = var[0] == 'u' ? o->u : o->v;
    for (ilay = 1; ilay <= nlay; ilay++)
      for (ipnt = 1; ipnt <= ndeg; ipnt++) R4(ipnt, ilay) = (float)A2(x, ipnt, ilay);
  } else if (!strcmp(var, "v_cc")) { /* pm:2884-2929 */
    double *w1 = dalloc(o->nd1), *w2 = dalloc(o->nd1);
    const double dl = P->dl;
    for (ilay = 1; ilay <= nlay; ilay++) {
      memset(w1, 0, sizeof(double) * o->nd1);
      memset(w2, 0, sizeof(double) * o->nd1);
      for (ipnt = 1; ipnt <= ndeg; ipnt++) {
        const int c1 = NEIG(1, ipnt), c3 = NEIG(3, ipnt), c5 = NEIG(5, ipnt), c7 = NEIG(7, ipnt);
        w1[ipnt] = ((A2(o->v, ipnt, ilay) - A2(o->v, c5, ilay)) / dl - (A2(o->u, ipnt, ilay) - A2(o->u, c7, ilay)) / dl) * o->mkpe[ipnt];
        w2[ipnt] = (A2(o->u, c1, ilay) - A2(o->u, ipnt, ilay)) / dl + (A2(o->v, c3, ilay) - A2(o->v, ipnt, ilay)) / dl;
      }
      for (ipnt = 1; ipnt <= ndeg; ipnt++) {
        const int c1 = NEIG(1, ipnt), c2 = NEIG(2, ipnt), c3 = NEIG(3, ipnt), c5 = NEIG(5, ipnt), c7 = NEIG(7, ipnt);
        const double a = w1[c1] - w1[ipnt], b = w1[c2] - w1[c3], c = w1[c3] - w1[ipnt], d = w1[c2] - w1[c1];
        const double e = w2[c1] - w2[ipnt], f = w2[ipnt] - w2[c5], g = w2[c3] - w2[ipnt], h = w2[ipnt] - w2[c7];
        R4(ipnt, ilay) = (float)(P->bvis + P->dvis * (dl * dl) * sqrt(a * a + b * b + c * c + d * d + e * e + f * f + g * g + h * h));
      }
    }
    free(w1);
    free(w2);
  } else if (!strcmp(var, "mont")) { /* pm:2930-2950 */
    for (ilay = 1; ilay <= nlay; ilay++)
      for (ipnt = 1; ipnt <= ndeg; ipnt++) {
        double s = 0.0;
        float r = (float)(-P->ocrp / (double)(P->nsal - 1) * P->hsal * o->mk_n[ipnt]
                          * pow3(P->hsal / (P->hmin * (1.0 - o->mk_n[ipnt]) + A2(o->hlay, ipnt, ilay))));
        for (i = 1; i <= ilay - 1; i++)
          r = r - (float)((P->rhon[ilay - 1] - P->rhon[i - 1]) * A2(o->hlay, ipnt, i) / P->rhon[ilay - 1]);
        for (i = 1; i <= nlay; i++) s += A2(o->hlay, ipnt, i);
        r = r + (float)(s - o->h_th[ipnt]);
        R4(ipnt, ilay) = r;
      }
  } else if (!strcmp(var, "pvor")) { /* pm:2951-2974 */
    const double dl = P->dl;
    for (ilay = 1; ilay <= nlay; ilay++)
      for (ipnt = 1; ipnt <= ndeg; ipnt++) {
        const int c5 = NEIG(5, ipnt), c6 = NEIG(6, ipnt), c7 = NEIG(7, ipnt);
        const double zeta = ((A2(o->v, ipnt, ilay) - A2(o->v, c5, ilay)) / dl - (A2(o->u, ipnt, ilay) - A2(o->u, c7, ilay)) / dl) * o->mkpe[ipnt];
        R4(ipnt, ilay) = (float)((o->fcor[ipnt] + zeta * P->uadv) * o->mkpi[ipnt]
                                 * (o->mk_n[ipnt] + o->mk_n[c5] + o->mk_n[c7] + o->mk_n[c6])
                                 / (A2(o->hlay, ipnt, ilay) + A2(o->hlay, c5, ilay) + A2(o->hlay, c6, ilay) + A2(o->hlay, c7, ilay)));
      }
  } else
    return -1;
#undef R4
  return 0;
}

/* ---------------------------------------------------------------------------------------------
 * accessors
 * ------------------------------------------------------------------------------------------- */
double *beom_oracle_array(beom_oracle *o, const char *name) {
#define F(n) if (!strcmp(name, #n)) return o->n;
  F(h_u) F(h_v) F(u) F(v) F(UU4) F(VV4) F(delu) F(delv) F(tt3d) F(tb3d) F(tu3d) F(taus)
  F(rvor) F(pvor) F(dive) F(v_cc) F(v_ll) F(fcor) F(mk_u) F(mk_v) F(mk_n) F(mkpe) F(mkpi)
  F(mont) F(d2hx) F(d2hy) F(h_bo) F(h_to) F(h_th) F(Ow) F(Os) F(Osum_)
  F(bodf) F(fnud) F(nudg) F(hdot) F(rs_h) F(dmdx) F(dmdy) F(tide) F(hlay) F(pi_s) F(h_0) F(h_2d)
#undef F
  return NULL;
}
int32_t *beom_oracle_iarray(beom_oracle *o, const char *name) {
  if (!strcmp(name, "neig")) return o->neig;
  if (!strcmp(name, "subc")) return o->subc;
  if (!strcmp(name, "posc")) return o->posc;
  if (!strcmp(name, "segm")) return o->segm;
  return NULL;
}
double beom_oracle_scalar(const beom_oracle *o, const char *name) {
  if (!strcmp(name, "ctim")) return o->ctim;
  if (!strcmp(name, "invf")) return o->invf;
  if (!strcmp(name, "ramp")) return o->ramp;
  if (!strcmp(name, "gene")) return o->gene;
  if (!strcmp(name, "w_ti")) return o->w_ti;
  if (!strcmp(name, "tres")) return o->tres;
  if (!strcmp(name, "flag_nudging")) return (double)o->flag_nudging;
  return NAN;
}
int beom_oracle_nseg(const beom_oracle *o) { return o->nseg; }

void beom_oracle_fields(beom_oracle *o, beom_fields *f) {
  memset(f, 0, sizeof *f);
  f->neig = o->neig; f->subc = o->subc;
  f->mk_u = o->mk_u; f->mk_v = o->mk_v; f->mk_n = o->mk_n; f->mkpe = o->mkpe; f->mkpi = o->mkpi;
  f->fcor = o->fcor; f->h_th = o->h_th;
  f->nudg = o->nudg; f->fnud = o->fnud; f->hdot = o->hdot; f->taus = o->taus; f->tide = o->tide; f->bodf = o->bodf;
  f->segm = o->segm; f->nseg = o->nseg;
  f->Ow = o->Ow; f->Os = o->Os; f->Osum_ = o->Osum_; f->pi_s = o->pi_s;
  f->flag_nudging = o->flag_nudging;
  f->invf = o->invf;
  f->w_ti = o->w_ti;
}
