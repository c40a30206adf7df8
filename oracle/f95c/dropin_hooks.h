// dropin_hooks.h -- the binding of INTEGRATION.md section 2, executed (TEST INFRASTRUCTURE ONLY).
//
// Included by the translator (oracle/f95c) inside namespace ref, after the module variables of the translated
// shared_mod / private_mod, when oracle/refbuild.py builds the DROP-IN flavour of the reference: the reference's own
// program -- its read_input_data, its integrate_time loop, its write_outputs / write_array, translated from its own
// sources -- with the five edits INTEGRATION.md lists applied to the text of private_mod.f95 and main.f95 before
// translation:
//     end of read_input_data            + call gpu_setup()
//     call distribute_stress()          -> call gpu_stress()
//     call first_three_timesteps(tstp)  -> call gpu_step(tstp, ctim, ramp, gene, .true., 1)
//     call gener_forward_backward(..)   -> call gpu_step(tstp, ctim, ramp, gene, upst, 0)
//     call write_outputs() in the loop  -> call gpu_download() first
//     main.f95, before call quit()      + call gpu_finalize()
// Each function below is the C++ twin of the Fortran lines INTEGRATION.md gives for that edit (no Fortran compiler
// exists to build fortran/beom_gpu_mod.f95): it hands the reference's OWN module arrays, in the reference's own layout,
// to the C ABI of include/beom_gpu.h.  The binary is linked against libbeom_gpu.so (a B200) or, for the CPU suite,
// against the emulated library of tools/emu.  Errors go the reference's way: errc / errm / quit().
#include "beom_gpu.h"

inline void gpu_check(int rc, const char* what) {
    if (rc == 0) return;
    char msg[512];
    beom_gpu_last_error(msg, sizeof msg);
    errc = rc;
    errm = f_cat(f_trim(errm), std::string(" ") + what + ": " + msg);
    quit();
}

inline void gpu_setup() {
    static beom_params par;
    static beom_fields fld;
    beom_gpu_options opt;
    std::memset(&par, 0, sizeof par);
    std::memset(&fld, 0, sizeof fld);
    par.lm = lm; par.mm = mm; par.nlay = nlay; par.ndeg = ndeg;
    par.dl = dl; par.cext = cext; par.f0 = f0;
    for (int k = 1; k <= nlay; ++k) {
        par.rhon[k - 1] = rhon(k);
        par.topl[k - 1] = topl(k);
    }
    par.dt_s = dt_s; par.dt_o = dt_o; par.dt_r = dt_r; par.dt3d = dt3d;
    par.bvis = bvis; par.dvis = dvis; par.bdrg = bdrg; par.hmin = hmin; par.hsbl = hsbl; par.hbbl = hbbl;
    par.g_fb = g_fb; par.uadv = uadv; par.qdrg = qdrg; par.ocrp = ocrp; par.rsta = rsta;
    par.xper = xper; par.yper = yper; par.diag = diag; par.rgld = rgld; par.mcbc = mcbc;
    par.tauw[0] = tauw.real(); par.tauw[1] = tauw.imag();
    par.svis = svis; par.tdrg = tdrg; par.topt = topt; par.plum = 0.0;
    par.dt = dt; par.hsal = hsal; par.hdry = hdry; par.tole = tole; par.pi = pi; par.grav = grav;
    par.rho0 = rho0; par.beta = beta; par.epsi = epsi; par.gamm = gamm; par.del1 = del1; par.del2 = del2;
    par.sor = sor; par.itmx = itmx; par.nsal = nsal;
    par.variant = BEOM_VARIANT_STANDARD;
    fld.neig = neig.d; fld.subc = subc.d;
    fld.mk_u = mk_u.d; fld.mk_v = mk_v.d; fld.mk_n = mk_n.d; fld.mkpe = mkpe.d; fld.mkpi = mkpi.d;
    fld.fcor = fcor.d; fld.h_th = h_th.d;
    fld.nudg = nudg.d; fld.fnud = fnud.d; fld.hdot = hdot.d;
    fld.taus = taus.d; fld.tide = tide.d; fld.bodf = bodf.d;
    fld.segm = nullptr; fld.nseg = 0;
    if (segm.allocated()) {
        fld.segm = segm.d;
        fld.nseg = static_cast<int32_t>(segm.ext(1));
    }
    fld.Ow = ow.d; fld.Os = os.d; fld.Osum_ = osum_.d; fld.pi_s = pi_s.d;
    fld.flag_nudging = flag_nudging ? 1 : 0;
    fld.invf = invf; fld.w_ti = w_ti(1);
    beom_gpu_default_options(&opt);
    if (const char* f = std::getenv("BEOM_DROPIN_FUSED")) opt.fused = std::atoi(f);  // tests name the kernel path
    gpu_check(beom_gpu_init(&par, &fld, &opt), "beom_gpu_init");
    gpu_check(beom_gpu_upload_state(hlay.d, u.d, v.d), "beom_gpu_upload_state");
}

inline void gpu_stress() { gpu_check(beom_gpu_stress(), "distribute_stress"); }

inline void gpu_step(int tstp, double ctim_, double ramp_, double gene_, bool upst, int first_three) {
    gpu_check(beom_gpu_step(tstp, ctim_, ramp_, gene_, upst ? 1 : 0, first_three),
              first_three ? "first_three_timesteps" : "gener_forward_backward");
}

inline void gpu_download() {
    gpu_check(beom_gpu_download_state(hlay.d, u.d, v.d), "write_outputs");
    if (rgld > 0.5) gpu_check(beom_gpu_download_pi_s(pi_s.d), "write_outputs (pi_s)");
}

inline void gpu_finalize() { beom_gpu_finalize(); }
