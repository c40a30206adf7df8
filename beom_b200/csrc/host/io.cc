// io.cc -- the reference's output side, kept on the host:
//   grid.bin            private_mod.f95:732-749   5 int32 records of length ndeg
//   h_0.bin             private_mod.f95:185-194   ndeg x nlay float32
//   param_basin.txt     private_mod.f95:1242-1297 Octave-evaluable "name = value ;" lines
//   eta_/u___/v___.bin  private_mod.f95:2817-2883 one float32 record (ndeg x nlay) per output
//   time.txt            private_mod.f95:2732-2738 one line per complete record
// and read_restart_record (private_mod.f95:1299-1420).  File formats are byte-compatible for the
// binary files; the two text files are syntactically compatible (their only consumers `eval'/`load'
// them: testcases/get_field.m:46-60, get_metadata.m:19-36).
#include <cmath>
#include <cstdio>
#include <cstring>
#include <fstream>

#include "host_model.h"

namespace {

// gfortran-style list-directed real(8): 17 significant digits, width 25
std::string fort_real(double x) {
  char b[64];
  const double a = std::fabs(x);
  if (x == 0.0) return "   0.0000000000000000     ";
  if (a >= 0.1 && a < 1.e17) {
    int before = (int)std::floor(std::log10(a)) + 1;
    if (before < 1) before = 1;
    std::snprintf(b, sizeof b, "%.*f", 17 - before, x);
    std::string s = b;
    while (s.size() < 20) s.insert(s.begin(), ' ');
    return s + "     ";
  }
  std::snprintf(b, sizeof b, "%.16E", x);  // d.ddddE+XX -> d.ddddE+0XX
  std::string s = b;
  size_t e = s.find('E');
  std::string mant = s.substr(0, e), ex = s.substr(e + 2);
  while (ex.size() < 3) ex.insert(ex.begin(), '0');
  std::string r = mant + "E" + s[e + 1] + ex;
  while (r.size() < 25) r.insert(r.begin(), ' ');
  return r;
}
std::string fort_int(long v) {
  char b[32];
  std::snprintf(b, sizeof b, "%12ld", v);
  return b;
}

bool write_record(const std::string &path, bool append_mode, int irec, const float *data, size_t count) {
  // direct access, recl = whole record, rec = irec (pm:2977-2993); `replace' on the first call
  std::FILE *f = std::fopen(path.c_str(), append_mode ? "r+b" : "wb");
  if (!f && append_mode) f = std::fopen(path.c_str(), "wb");
  if (!f) return false;
  bool ok = std::fseek(f, (long)((size_t)(irec - 1) * count * sizeof(float)), SEEK_SET) == 0 &&
            std::fwrite(data, sizeof(float), count, f) == count;
  std::fclose(f);
  return ok;
}

}  // namespace

bool beom_host_write_grid_files(beom_host *h) {
  const size_t nd = (size_t)h->ndeg;
  {  // grid.bin, pm:732-749
    std::FILE *f = std::fopen((h->odir + "grid.bin").c_str(), "wb");
    if (!f) return false;
    std::vector<int32_t> rec(nd);
    auto put = [&](const double *m) {
      for (size_t p = 0; p < nd; p++) rec[p] = (int32_t)std::lround(m[p + 1]);
      return std::fwrite(rec.data(), 4, nd, f) == nd;
    };
    bool ok;
    if (h->device_init) {  // the five records as the device made them (beom_gpu_download_grid_files)
      ok = std::fwrite(h->grid5.data(), 4, 5 * nd, f) == 5 * nd;
    } else {
      ok = std::fwrite(h->posc.data() + 1, 4, nd, f) == nd;
      ok = ok && put(h->mk_n.data()) && put(h->mk_u.data()) && put(h->mk_v.data()) && put(h->mkpi.data());
    }
    std::fclose(f);
    if (!ok) return false;
  }
  {  // h_0.bin, pm:185-194
    std::FILE *f = std::fopen((h->odir + "h_0.bin").c_str(), "wb");
    if (!f) return false;
    bool ok = std::fwrite(h->h_0_r4.data(), 4, h->h_0_r4.size(), f) == h->h_0_r4.size();
    std::fclose(f);
    if (!ok) return false;
  }
  return true;
}

bool beom_host_save_metadata(beom_host *h) {  // pm:1242-1297
  std::ofstream o(h->odir + "param_basin.txt");
  if (!o) return false;
  const beom_params &P = h->p;
  auto I = [&](const char *n, long v) { o << " " << n << " = " << fort_int(v) << " ;\n"; };
  auto R = [&](const char *n, double v) { o << " " << n << " = " << fort_real(v) << " ;\n"; };
  auto V = [&](const char *n, const double *v, int k) {
    o << " " << n << " = [";
    for (int a = 0; a < k; a++) o << fort_real(v[a]);
    o << " ];\n";
  };
  auto S = [&](const char *n, const std::string &v) { o << " " << n << " = '" << v << "';\n"; };
  I("lm            ", P.lm); I("mm            ", P.mm); I("nlay          ", P.nlay); I("ndeg          ", P.ndeg);
  R("dl            ", P.dl); R("cext          ", P.cext); R("f0            ", P.f0);
  V("rhon          ", P.rhon, P.nlay); V("topl          ", P.topl, P.nlay);
  R("dt_s          ", P.dt_s); R("dt_o          ", P.dt_o); R("dt_r          ", P.dt_r); R("dt3d          ", P.dt3d);
  R("bvis          ", P.bvis); R("dvis          ", P.dvis); R("svis          ", P.svis); R("bdrg          ", P.bdrg);
  R("tdrg          ", P.tdrg); R("tole          ", P.tole); I("nsal          ", P.nsal); R("hsal          ", P.hsal);
  R("hmin          ", P.hmin); R("hdry          ", P.hdry); R("hsbl          ", P.hsbl);
  R("hbbl          ", P.hsbl);  // the reference echoes hsbl here (pm:1276)
  R("g_fb          ", P.g_fb); R("uadv          ", P.uadv); R("qdrg          ", P.qdrg); R("ocrp          ", P.ocrp);
  R("tauwx         ", P.tauw[0]); R("tauwy         ", P.tauw[1]); R("rsta          ", P.rsta);
  R("xper          ", P.xper); R("yper          ", P.yper); R("diag          ", P.diag); R("rgld          ", P.rgld);
  R("mcbc          ", P.mcbc); R("topt          ", P.topt);
  S("idir          ", h->idir); S("desc          ", h->desc);
  R("dt            ", P.dt);
  return (bool)o;
}

extern "C" {

// write_outputs + write_array for eta_, u___, v___ (pm:2681-2883).  The diag=1 records
// (pvor/mont/v_cc) are produced by the GPU library and appended by the caller (run.cc).
int beom_host_write_outputs(beom_host *h, double ctim) {
  const int nlay = h->nlay, ndeg = h->ndeg;
  const size_t n = h->nd1, cnt = (size_t)ndeg * nlay;
  if (h->odir.empty()) { beom_host_set_error("write_outputs: odir is empty"); return -1000; }
  if (!h->out_init) h->irec = 1;
  std::vector<float> rec(cnt);
  auto R = [&](int p, int l) -> float & { return rec[(size_t)l * ndeg + (p - 1)]; };

  // eta: cumulate hlay - h_0 upward from the bottom layer, in float32 (pm:2848-2875)
  for (int l = nlay - 1; l >= 0; l--)
    for (int p = 1; p <= ndeg; p++) {
      const double dh = h->hlay[(size_t)l * n + p] - (double)h->h_0_r4[(size_t)l * ndeg + (p - 1)];
      R(p, l) = (l == nlay - 1) ? (float)dh : (float)(dh + (double)R(p, l + 1));
    }
  if (h->p.rgld > 0.5)
    for (int p = 1; p <= ndeg; p++) R(p, 0) = (float)h->pi_s[p];
  bool ok = write_record(h->odir + "eta_.bin", h->out_init, h->irec, rec.data(), cnt);
  for (int l = 0; l < nlay; l++)
    for (int p = 1; p <= ndeg; p++) R(p, l) = (float)h->u[(size_t)l * n + p];
  ok = ok && write_record(h->odir + "u___.bin", h->out_init, h->irec, rec.data(), cnt);
  for (int l = 0; l < nlay; l++)
    for (int p = 1; p <= ndeg; p++) R(p, l) = (float)h->v[(size_t)l * n + p];
  ok = ok && write_record(h->odir + "v___.bin", h->out_init, h->irec, rec.data(), cnt);
  if (!ok) { beom_host_set_error("write_array: could not write a record into " + h->odir); return -1001; }

  // time.txt only after all arrays are on disk (pm:2724-2738)
  {
    std::FILE *f = std::fopen((h->odir + "time.txt").c_str(), h->out_init ? "a" : "w");
    if (!f) { beom_host_set_error("write_outputs: cannot open time.txt"); return -1002; }
    std::fprintf(f, "%s\n", fort_real(ctim).c_str());
    std::fclose(f);
  }
  h->irec += 1;
  h->out_init = true;
  h->ctim = ctim;

  for (int l = 0; l < nlay; l++) {  // pm:2772-2794
    double lo = INFINITY, hi = -INFINITY;
    for (int p = 0; p <= ndeg; p++)
      if (h->mk_n[p] > 0.5) { lo = std::fmin(lo, h->hlay[(size_t)l * n + p]); hi = std::fmax(hi, h->hlay[(size_t)l * n + p]); }
    std::printf(" min/max h %d = %.15g %.15g\n", l + 1, lo, hi);
  }
  for (int l = 0; l < nlay; l++)  // pm:2798-2808
    for (int p = 1; p <= ndeg; p++)
      if (h->mk_n[p] > 0.5 && h->hlay[(size_t)l * n + p] < 0.5 * h->p.hmin) {
        char b[128];
        std::snprintf(b, sizeof b, " layer number %d has its thickness < hmin; Calculation halted.", l + 1);
        beom_host_set_error(std::string("In main, in subroutine write_outputs,") + b);
        return -(l + 1);
      }
  return 0;
}

// write_outputs with the records already made on the device (beom_gpu_records_begin / _wait, include/beom_gpu.h): the same
// files, the same report and the same halt as beom_host_write_outputs, without the double-precision state crossing the bus.
int beom_host_write_records(beom_host *h, double ctim, const beom_records *r) {
  const int nlay = h->nlay, ndeg = h->ndeg;
  const size_t cnt = (size_t)ndeg * nlay;
  if (h->odir.empty()) { beom_host_set_error("write_outputs: odir is empty"); return -1000; }
  if (r->first_point != 1 || r->count != ndeg) { beom_host_set_error("write_outputs: the record set does not cover 1..ndeg"); return -1003; }
  if (!h->out_init) h->irec = 1;
  bool ok = write_record(h->odir + "eta_.bin", h->out_init, h->irec, r->eta, cnt);
  ok = ok && write_record(h->odir + "u___.bin", h->out_init, h->irec, r->u, cnt);
  ok = ok && write_record(h->odir + "v___.bin", h->out_init, h->irec, r->v, cnt);
  if (r->pvor) {  // pm:2718-2722
    ok = ok && write_record(h->odir + "pvor.bin", h->out_init, h->irec, r->pvor, cnt);
    ok = ok && write_record(h->odir + "mont.bin", h->out_init, h->irec, r->mont, cnt);
    ok = ok && write_record(h->odir + "v_cc.bin", h->out_init, h->irec, r->v_cc, cnt);
  }
  if (!ok) { beom_host_set_error("write_array: could not write a record into " + h->odir); return -1001; }
  {
    std::FILE *f = std::fopen((h->odir + "time.txt").c_str(), h->out_init ? "a" : "w");
    if (!f) { beom_host_set_error("write_outputs: cannot open time.txt"); return -1002; }
    std::fprintf(f, "%s\n", fort_real(ctim).c_str());
    std::fclose(f);
  }
  h->irec += 1;
  h->out_init = true;
  h->ctim = ctim;
  for (int l = 0; l < nlay; l++) std::printf(" min/max h %d = %.15g %.15g\n", l + 1, r->hmin[l], r->hmax[l]);  // pm:2772-2794
  if (r->thin_layer) {  // pm:2798-2808
    char b[128];
    std::snprintf(b, sizeof b, " layer number %d has its thickness < hmin; Calculation halted.", r->thin_layer);
    beom_host_set_error(std::string("In main, in subroutine write_outputs,") + b);
    return -r->thin_layer;
  }
  return 0;
}

// Append one diag record (float32, ndeg x nlay) to <var>.bin at the record just written.
int beom_host_write_diag_record(beom_host *h, const char *var, const float *rec) {
  const size_t cnt = (size_t)h->ndeg * h->nlay;
  const int irec = h->irec - 1;
  if (irec < 1) return -1;
  return write_record(h->odir + var + ".bin", irec > 1, irec, rec, cnt) ? 0 : -1;
}

int beom_host_read_restart(beom_host *h) {  // pm:1299-1420
  const int nlay = h->nlay, ndeg = h->ndeg;
  const size_t n = h->nd1, cnt = (size_t)ndeg * nlay;
  int irec = 0;
  {
    std::ifstream t(h->odir + "time.txt");
    double d;
    while (t >> d) { irec++; h->tres = d; }
  }
  if (irec < 1) { beom_host_set_error("read_restart_record: time.txt is empty or missing"); return -1; }
  std::vector<float> rec(cnt);
  auto get = [&](const char *var) {
    std::FILE *f = std::fopen((h->odir + var + ".bin").c_str(), "rb");
    if (!f) return false;
    bool ok = std::fseek(f, (long)((size_t)(irec - 1) * cnt * 4), SEEK_SET) == 0 && std::fread(rec.data(), 4, cnt, f) == cnt;
    std::fclose(f);
    return ok;
  };
  auto fail = [&](const char *var) {
    beom_host_set_error(std::string(" could not open/read file ") + var + ".bin from directory " + h->odir);
    return -2;
  };
  if (!get("u___")) return fail("u___");
  for (int l = 0; l < nlay; l++)
    for (int p = 1; p <= ndeg; p++) h->u[(size_t)l * n + p] = (double)rec[(size_t)l * ndeg + p - 1];
  if (!get("v___")) return fail("v___");
  for (int l = 0; l < nlay; l++)
    for (int p = 1; p <= ndeg; p++) h->v[(size_t)l * n + p] = (double)rec[(size_t)l * ndeg + p - 1];
  if (!get("eta_")) return fail("eta_");
  for (int l = 0; l < nlay; l++)  // pm:1391-1407
    for (int p = 1; p <= ndeg; p++) {
      double t = (double)h->h_0_r4[(size_t)l * ndeg + p - 1] + (double)rec[(size_t)l * ndeg + p - 1];
      if (l < nlay - 1) t = t - (double)rec[(size_t)(l + 1) * ndeg + p - 1];
      h->hlay[(size_t)l * n + p] = t * h->mk_n[p];
    }
  h->out_init = true;  // pm:2704-2710
  h->irec = irec + 1;
  return 0;
}

}  // extern "C"
