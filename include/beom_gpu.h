/* beom_gpu.h -- C ABI of the B200 (sm_100a) implementation of BEOM's per-timestep
 * layered shallow-water update.
 *
 * This is the drop-in boundary.  The reference (zhazorken/beom) has no FFI of its own: the hot path
 * is a set of private procedures of `module private_mod' operating on module arrays
 * (private_mod.f95:27-93) and called from `integrate_time' (private_mod.f95:1840-1919).  Each entry
 * point below replaces one of those call sites; the citation says which.  The Fortran side binds them
 * with ISO_C_BINDING (see fortran/beom_gpu_mod.f95 and INTEGRATION.md).
 *
 * Conventions
 *   - plain C, no torch / C++ types; every function returns 0 on success or a negative code,
 *     the text is available from beom_gpu_last_error().
 *   - arrays are the reference's own arrays, Fortran column-major, passed by address of their first
 *     element (index 0 = the "discarded cell" sentinel, private_mod.f95:31-33):
 *       x(0:ndeg)            -> double[ndeg+1]
 *       x(0:ndeg,nlay)       -> double[nlay][ndeg+1]
 *       fnud(0:ndeg,nlay,3)  -> double[3][nlay][ndeg+1]
 *       nudg(0:ndeg,3)       -> double[3][ndeg+1]
 *       taus(0:ndeg,2)       -> double[2][ndeg+1]
 *       tide(2,1,0:ndeg,3)   -> double[3][ndeg+1][1][2]
 *       bodf(nlay,2)         -> double[2][nlay]
 *       neig(8,0:ndeg)       -> int32[ndeg+1][8]
 *       subc(0:ndeg,2)       -> int32[2][ndeg+1]
 *       segm(nseg,18)        -> int32[18][nseg]
 *   - the library copies what it needs; host arrays stay owned by the caller.
 *   - one context per process (the reference is one model per process); one GPU per process;
 *     several processes (one per GPU, y-slabs) cooperate through beom_gpu_comm_*.
 */
#ifndef BEOM_GPU_H
#define BEOM_GPU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BEOM_MAXLAY 16
#define BEOM_ABI_VERSION 1

/* Which private_mod*.f95 the update_h epilogue follows (SURVEY section 8, row a1'). */
enum {
  BEOM_VARIANT_STANDARD = 0, /* private_mod.f95:1593-1702            */
  BEOM_VARIANT_1D       = 1, /* private_mod1d.f95:1635-1663, :2046   */
  BEOM_VARIANT_3D       = 2, /* private_mod3d.f95:1635-1683          */
  BEOM_VARIANT_PLUME    = 3  /* private_modplumenew.f95:1637-1724    */
};

/* POD mirror of the public parameters of shared_mod.f95:41-111.  All reals are the *double values the
 * Fortran parameter holds*, i.e. default-real literals are already rounded through float32
 * (e.g. grav = (double)9.8f).  beom_params_parse() produces this from a shared_mod.f95-style text. */
typedef struct beom_params {
  /* user section, shared_mod.f95:41-77 */
  int32_t lm, mm, nlay, ndeg;
  double dl, cext, f0;
  double rhon[BEOM_MAXLAY], topl[BEOM_MAXLAY];
  double dt_s, dt_o, dt_r, dt3d, bvis, dvis, bdrg, hmin, hsbl, hbbl;
  double g_fb, uadv, qdrg, ocrp, rsta, xper, yper, diag, rgld, mcbc;
  double tauw[2];
  /* used by private_mod*.f95 but absent from the shipped shared_mod.f95 */
  double svis, tdrg, topt, plum;
  /* "other constants", shared_mod.f95:83-111 */
  double dt, hsal, hdry, tole, pi, grav, rho0, beta, epsi, gamm, del1, del2, sor;
  int32_t itmx, nsal;
  /* not in shared_mod.f95: which private_mod file's update_h to follow */
  int32_t variant;
  int32_t reserved_;
} beom_params;

/* Static fields produced by the reference's read_input_data (private_mod.f95:105-250).  Pointers that
 * are optional may be NULL, meaning "all zero / feature absent". */
typedef struct beom_fields {
  const int32_t *neig;   /* required */
  const int32_t *subc;   /* required */
  const double  *mk_u, *mk_v, *mk_n, *mkpe, *mkpi; /* required */
  const double  *fcor;   /* required */
  const double  *h_th;   /* required */
  const double  *nudg;   /* optional: NULL = no sponge                   */
  const double  *fnud;   /* optional: required when nudg or tide present  */
  const double  *hdot;   /* optional */
  const double  *taus;   /* optional */
  const double  *tide;   /* optional */
  const double  *bodf;   /* optional */
  const int32_t *segm;   /* optional, with nseg */
  const double  *Ow, *Os, *Osum_; /* optional (rgld = 1 only)            */
  const double  *pi_s;   /* optional initial surface pressure (rgld = 1) */
  int32_t nseg;
  int32_t flag_nudging;  /* private_mod.f95:93, 868-871 */
  double  invf;          /* private_mod.f95:223-229 */
  double  w_ti;          /* private_mod.f95:953 */
} beom_fields;

/* Behavioural options of this library that have no counterpart in the reference. */
typedef struct beom_gpu_options {
  int32_t device;        /* CUDA device ordinal; -1 = current/LOCAL_RANK              */
  int32_t fused;         /* 1 = fused single-pass step where the case allows, 0 = one  */
                         /*     kernel per reference loop (always available), 2 = by   */
                         /*     size (the default): the fused step from 250 000        */
                         /*     cell-layers per rank on, below that the split path,    */
                         /*     which is quicker on latency-bound grids  */
  int32_t rank, nranks;  /* y-slab decomposition; nranks = 1 for a single GPU          */
  int32_t strict;        /* reserved (kernels are always built -fmad=false)            */
  int32_t reserved_[3];
} beom_gpu_options;

const char *beom_gpu_version(void);
int  beom_gpu_abi_version(void);
/* Copies the last error text (NUL-terminated, truncated to len) and returns its full length. */
int  beom_gpu_last_error(char *buf, int len);

/* Replaces nothing in the reference: fills opt with defaults. */
void beom_gpu_default_options(beom_gpu_options *opt);

/* Called once at the end of read_input_data (after private_mod.f95:243).  Uploads the static fields.
 * Fails (no CPU fallback) if no sm_100 device is usable. */
int beom_gpu_init(const beom_params *par, const beom_fields *fld, const beom_gpu_options *opt);

/* After init and after read_restart_record (private_mod.f95:238): hlay,u,v (0:ndeg,nlay).  Zeroes
 * h_u,h_v,rs_h,dmdx,dmdy like initialize_variables (private_mod.f95:252-307). */
int beom_gpu_upload_state(const double *hlay, const double *u, const double *v);

/* distribute_stress (private_mod.f95:1921-2149; call sites :1863, :1895). */
int beom_gpu_stress(void);

/* Body of first_three_timesteps (private_mod.f95:2151-2223; first_three = 1) or
 * gener_forward_backward (private_mod.f95:2225-2316; first_three = 0).  The scalars are the ones
 * integrate_time computes (private_mod.f95:1862-1901).  Asynchronous. */
int beom_gpu_step(int tstp, double ctim, double ramp, double gene, int upst, int first_three);

/* The reference's whole loop body for steps tstp0..tstp1 (private_mod.f95:1861-1906) with ctim,
 * ramp, upst, gene derived exactly as integrate_time does, tres being the restart time in days.
 * Convenience for hosts that do not need per-step control; asynchronous. */
int beom_gpu_advance(int tstp0, int tstp1, double tres);

/* Before write_outputs (private_mod.f95:1909): synchronises and copies hlay,u,v back. */
int beom_gpu_download_state(double *hlay, double *u, double *v);

/* Optional extra state for bit-level checks: any pointer may be NULL.
 * h_u,h_v: (0:ndeg,nlay); rs_h: (2,0:ndeg,nlay); dmdx,dmdy: (3,0:ndeg,nlay) in the reference's order
 * (history index fastest, slot 3 newest). */
int beom_gpu_download_aux(double *h_u, double *h_v, double *rs_h, double *dmdx, double *dmdy);

/* Diagnostic records of write_array (private_mod.f95:2884-2974) exactly as it would write them:
 * float32, ndeg x nlay (no sentinel), computed on the device with the reference's mixed
 * float32/double arithmetic; NULL = skip. */
int beom_gpu_download_diag(float *pvor, float *mont, float *v_cc);

/* The output records of write_outputs / write_array (private_mod.f95:2681-2883) produced ON THE DEVICE and copied
 * out asynchronously: float32 `eta_' (layer thickness minus the float32 rest thickness of h_0.bin, cumulated upward from the
 * bottom layer in float32 exactly as pm:2848-2870 does; pi_s on top under the rigid lid), `u___', `v___', with_diag != 0
 * also `pvor', `mont', `v_cc' (pm:2884-2974), all in the record layout of the files ([nlay][count] floats for the vector
 * points first_point .. first_point + count - 1; one rank: 1 .. ndeg, i.e. the record itself), plus the min / max
 * thickness of every layer over the wet points and the `hlay < hmin / 2' verdict of pm:2772-2808.
 *   beom_gpu_set_rest_thickness   h_0.bin's content ([nlay][ndeg] float32), once, before the first record
 *   beom_gpu_records_begin        enqueues the record kernels behind the steps issued so far and their device->host copy
 *                                 into page-locked buffers owned by the library on a copy stream; returns at once, the caller
 *                                 goes on stepping.  At most two record sets may be in flight.
 *   beom_gpu_records_wait         blocks until the OLDEST begun set is in host memory; the pointers stay valid until the
 *                                 next but one beom_gpu_records_begin. */
typedef struct beom_records {
  const float *eta, *u, *v;          /* [nlay][count] */
  const float *pvor, *mont, *v_cc;   /* NULL unless begun with_diag */
  int first_point, count;
  double hmin[BEOM_MAXLAY], hmax[BEOM_MAXLAY];  /* over this rank's wet points (+inf / -inf if it has none) */
  int thin_layer;                    /* 1-based first layer whose thinnest wet point is below hmin / 2, 0 = none */
} beom_records;
int beom_gpu_set_rest_thickness(const float *h_0_r4);
int beom_gpu_records_begin(int with_diag);
int beom_gpu_records_wait(beom_records *out);

/* Initialisation from the raw input files with read_input_data's grid-shaped work done ON THE DEVICE (private_mod.f95:105-250:
 * the masks and the vector numbering of index_grid_points as a prefix sum, the rest thickness -- layers stacked from the bottom or
 * the per-column Newton solve of get_equilibrium_thickness_h_0 --, the relaxation targets, forcing planes and the initial state
 * of read_input_file), written straight into the dense planes with the reference's arithmetic; replaces read_input_data +
 * beom_gpu_init + beom_gpu_upload_state for the cases it covers.  The arrays hold the files' contents as they are on disk
 * (little-endian float32, Fortran order, no record markers; NULL = file absent).  par carries the derived values of
 * shared_mod.f95 (dt, hsal, ...) and ndeg.  Periodic domains and sponges with the open-boundary copy are covered as well: their
 * irregular part -- the aliases of index_grid_points, the layout analysis beom_gpu_init runs on the neighbour table, the segment
 * table of index_boundary_points -- is done on the host from the depth grid (the host driver's own source, grid_index.h), the planes
 * are filled on the device.  Returns 0, a negative error, or BEOM_GRIDS_UNSUPPORTED for what is initialised on the host: the rigid
 * lid, the 1d/3d/plume variants, h_to.bin, restarts. */
typedef struct beom_grids {
  const float *h_bo;   /* (0:lm+1, 0:mm+1) */
  const float *init;   /* (0:lm+1, 0:mm+1, nlay, 3) */
  const float *nudg;   /* (0:lm+1, 0:mm+1, 3) */
  const float *taus;   /* (0:lm+1, 0:mm+1, 2) */
  const float *fcor;   /* (0:lm+1, 0:mm+1) */
  const float *hdot;   /* (0:lm+1, 0:mm+1, nlay) */
  const float *bodf;   /* (nlay, 2) */
  const float *tide;   /* (2, 1, 0:lm+1, 0:mm+1, 3) */
  int32_t has_h_to;    /* h_to.bin exists (not handled here) */
} beom_grids;
#define BEOM_GRIDS_UNSUPPORTED 1
int beom_gpu_init_grids(const beom_params *par, const beom_grids *grids, const beom_gpu_options *opt);
/* Grid coordinates (i, j) of the vector points this rank holds (beom_gpu_point_range: first .. first + count - 1). */
int beom_gpu_download_subc(int32_t *si, int32_t *sj);
/* After beom_gpu_init_grids on one rank: grid.bin's five records (posc, mk_n, mk_u, mk_v, mkpi; 5 x ndeg int32, private_mod.f95:732-749)
 * and h_0.bin's content (float32 [nlay][ndeg], :185-194), so that a host driver can write its metadata files; either may be NULL. */
int beom_gpu_download_grid_files(int32_t *grid5, float *h_0_r4);
/* A static plane in the reference's vector layout (0:ndeg), for checks: name = "fcor", "h_th", "nudg", "fnud", "taus", "hdot",
 * "h_0", "flags" (the flag byte: masks mk_n = 1, mk_u = 2, mk_v = 4, mkpe = 8, mkpi = 16, vector point = 32); index = plane. */
int beom_gpu_debug_static(const char *name, int index, double *out);

/* Rigid-lid surface pressure pi_s(0:ndeg) (private_mod.f95:91). */
int beom_gpu_download_pi_s(double *pi_s);
/* Sweeps the last surf_pressure solve took (the reference's `iters', private_mod.f95:1756-1803). */
int beom_gpu_pi_iterations(int *iters);

/* Conservation integrals (testcases/conservation.m:116-211) over the vector points (frozen periodic duplicates
 * excluded, as the script does at :196-201), per layer l:
 *   vol[l] = sum over wet points of hlay                      (the script's `volu` times the number of wet points)
 *   ke[l]  = 0.5 sum 0.5 (U + U(E)) + 0.5 sum 0.5 (V + V(N)),  U = u^2 0.5 (h(W) + h),  V = v^2 0.5 (h(S) + h), dry h = 0
 *            (the script's `kine` before the factor rhon(l) dl^2)
 *   pe[0]  = sum over wet points of eta_1^2, eta_1 = sum_l (hlay - h_0)   (`pote` before 0.5 rhon(1) grav dl^2)
 * h_0(0:ndeg, nlay) is the rest thickness in the reference layout (NULL = the one passed before).  Device reduction:
 * warp-shuffle trees, block partials added in a fixed order; ncclAllReduce(SUM) over the ranks. */
int beom_gpu_diagnostics(const double *h_0, double *vol, double *ke, double *pe);

/* The same plus the vorticity integrals of conservation.m:169-211, per layer l over the same points (npts of them):
 *   enst[l]  = sum of 0.5 pvor^2 hatp          (the script's `enst` = enst[l] / npts: domain-mean potential enstrophy)
 *   zeta[l]  = sum of (pvor hatp - fcor)        (`rvor` = zeta[l] / npts: domain-mean relative vorticity)
 *   zeta2[l] = sum of (pvor hatp - fcor)^2      (`rstd`^2 = (zeta2 - zeta^2 / npts) / (npts - 1))
 * pvor as write_array evaluates it (private_mod.f95:2951-2974, kept in double here, fcor the psi-point field), hatp the mean
 * thickness of the up to four cells around the psi point that hold a value.  Any pointer may be NULL. */
int beom_gpu_diagnostics_all(const double *h_0, double *vol, double *ke, double *pe, double *enst, double *zeta, double *zeta2, double *npts);

/* y-slab runs: the vector points this rank holds (owned rows + halo rows: first,count) and owns
 * (own_first, own_count).  beom_gpu_set_window(first, count) declares that the state arrays passed to
 * upload_state / download_state / download_aux from now on hold only points first..first+count-1 of
 * each layer ([nlay][count], must cover the rank's points); count <= 0 restores whole arrays. */
int beom_gpu_point_range(int *first, int *count, int *own_first, int *own_count);
int beom_gpu_set_window(int first, int count);

/* Page-locked host memory for state arrays that cross the boundary every output interval (the
 * Fortran side maps it with c_f_pointer); plain malloc'ed arrays work too, only slower.  The pages are placed on the NUMA
 * node of the current device (sysfs; BEOM_HOST_NUMA=0 switches that off): call it after beom_gpu_init / cudaSetDevice. */
void *beom_gpu_host_alloc(size_t bytes);
void  beom_gpu_host_free(void *p);

/* Blocks until all queued work is done; returns the sticky error if a kernel failed. */
int beom_gpu_sync(void);

/* Milliseconds of device time between two marks on the library's compute stream (CUDA events).
 * beom_gpu_mark(0) ... beom_gpu_mark(1); beom_gpu_elapsed_ms() synchronises on mark 1. */
int beom_gpu_mark(int which);
int beom_gpu_elapsed_ms(double *ms);
/* Number of kernel launches issued by this library since init. */
long long beom_gpu_launch_count(void);
/* Steps that ran as one CUDA-graph launch instead of kernel by kernel (latency-bound grids on one rank, steady steps: the launch
 * sequence of each buffer phase is captured once and replayed; BEOM_GRAPH=0 / 1 overrides the size rule).  Diagnostic. */
long long beom_gpu_graph_launch_count(void);
/* Name of the kernel path in use ("fused" / "split"), for logs. */
const char *beom_gpu_path(void);
/* Which instantiation of the fused step runs: "specialised (options <mask>, <n> layers, <g> column groups)", "general (...)", "none". */
const char *beom_gpu_fused_variant(void);

/* Multi-GPU (one process per GPU).  Rank 0 obtains an id (128 bytes), the host broadcasts it by any
 * means (torch.distributed, MPI, a file), every rank calls comm_init before beom_gpu_init. */
int beom_gpu_comm_unique_id(char id[128]);
int beom_gpu_comm_init(const char id[128], int rank, int nranks, int device);
int beom_gpu_comm_finalize(void);

/* Before quit() (main.f95:35). */
int beom_gpu_finalize(void);

#ifdef __cplusplus
}
#endif
#endif /* BEOM_GPU_H */
