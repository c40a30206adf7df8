/* beom_host.h -- host side of the drop-in: what stays Fortran in the reference (parameter set,
 * read_input_data, integrate_time, write_outputs) restated in C++ because this environment has no
 * Fortran compiler.  Same names, argument meaning and error behaviour as the reference routines
 * (private_mod.f95; cited per function in the .cc files).  The per-timestep work is done by
 * include/beom_gpu.h; nothing here computes the hot path on the CPU.
 */
#ifndef BEOM_HOST_H
#define BEOM_HOST_H
#include "../../../include/beom_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct beom_host beom_host;

/* Parse a shared_mod.f95 (or just the parameter block printed by testcases/print_params.m) with
 * Fortran literal semantics: a real literal without kind suffix is default real (float32) and is
 * widened to double afterwards (shared_mod.f95:41-111).  Strings receive idir/odir/desc (may be
 * NULL).  Returns 0 or a negative code (beom_host_last_error). */
int beom_params_parse(const char *text, beom_params *out, char *idir, char *odir, char *desc, int slen);
int beom_params_parse_file(const char *path, beom_params *out, char *idir, char *odir, char *desc, int slen);
/* Fill the "other constants" (shared_mod.f95:83-111) from the user section. */
void beom_params_derive(beom_params *p);
/* Defaults = the values in the reference's shipped shared_mod.f95 for everything print_params.m does
 * not print (rgld, mcbc, constants); user section zeroed. */
void beom_params_defaults(beom_params *p);

int beom_host_last_error(char *buf, int len);

/* read_input_data (private_mod.f95:105-250) up to and including save_metadata; writes grid.bin,
 * h_0.bin and param_basin.txt into odir when odir is non-empty.  Does NOT touch the GPU. */
beom_host *beom_host_create(const beom_params *par, const char *idir, const char *odir, const char *desc);
/* The same with read_input_data's grid-shaped work done on the device (beom_gpu_init_grids, include/beom_gpu.h): the input
 * files are memory-mapped and handed to the GPU library, which is left initialised with the initial state in place; the host
 * object keeps no big array (beom_host_array returns NULL for them).  Cases the device path does not cover (periodic domains,
 * the rigid lid, ...) fall back to beom_host_create.  beom_host_run works on either kind. */
beom_host *beom_host_create_on_device(const beom_params *par, const char *idir, const char *odir, const char *desc, const beom_gpu_options *opt);
void beom_host_destroy(beom_host *h);

/* Reference-layout arrays (see include/beom_gpu.h).  Names: neig subc posc segm (int32);
 * mk_u mk_v mk_n mkpe mkpi fcor h_th nudg fnud hdot taus tide bodf Ow Os Osum_ pi_s h_0 hlay u v h_2d. */
double  *beom_host_array(beom_host *h, const char *name);
int32_t *beom_host_iarray(beom_host *h, const char *name);
double   beom_host_scalar(const beom_host *h, const char *name);
void     beom_host_fields(beom_host *h, beom_fields *f);
const beom_params *beom_host_params(const beom_host *h);

/* nstp, notp, n_3d (private_mod.f95:1853-1856). */
void beom_host_counts(const beom_host *h, int *nstp, int *notp, int *n_3d);

/* write_outputs (private_mod.f95:2681-2815) on the host copies hlay,u,v of the model; appends a
 * record to eta_/u___/v___(.bin) (+ pvor/mont/v_cc when diag = 1) and a line to time.txt.  Returns a
 * negative layer number if a wet layer is thinner than 0.5*hmin (private_mod.f95:2798-2808). */
int beom_host_write_outputs(beom_host *h, double ctim);
/* the same files, report and halt from records made on the device (beom_gpu_records_begin / _wait) */
int beom_host_write_records(beom_host *h, double ctim, const beom_records *r);
/* read_restart_record (private_mod.f95:1299-1344). */
int beom_host_read_restart(beom_host *h);

/* run() = read_input_data + integrate_time (private_mod.f95:99-103, 1840-1919) with the step
 * routines executed by the GPU library.  max_steps > 0 stops early (testing). */
int beom_host_run(beom_host *h, const beom_gpu_options *opt, int max_steps);

#ifdef __cplusplus
}
#endif
#endif
