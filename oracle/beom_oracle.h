/* beom_oracle.h -- TEST INFRASTRUCTURE ONLY.
 *
 * CPU restatement (plain C, strict IEEE double, -ffp-contract=off) of the reference's
 * read_input_data + per-timestep update (private_mod.f95 and its 1d/3d/plume variants).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load
 * this library; nothing under beom_b200/ does.
 *
 * PARITY PINNED (round 2) against the reference itself: its Fortran sources are translated to C++ by oracle/f95c
 * (no Fortran compiler exists here), built into oracle/_ref and compared with this restatement bit for bit on every
 * module array (tests/test_reference_pin.py, oracle/README.md); the analytical solutions of the reference's test scripts
 * and its documented conservation property (tests/test_oracle_pins.py) are kept as a second, independent pin.
 */
#ifndef BEOM_ORACLE_H
#define BEOM_ORACLE_H
#include "../include/beom_gpu.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct beom_oracle beom_oracle;

/* read_input_data (private_mod.f95:105-250) with inputs read from directory idir (may be NULL/""
 * for "no input files").  Returns NULL on error (message via beom_oracle_error). */
beom_oracle *beom_oracle_create(const beom_params *par, const char *idir);
void         beom_oracle_destroy(beom_oracle *o);
const char  *beom_oracle_error(void);

/* integrate_time (private_mod.f95:1840-1919) for steps tstp0..tstp1 inclusive (1-based; the
 * reference runs 1..nstp).  No output is written. */
int  beom_oracle_advance(beom_oracle *o, int tstp0, int tstp1);
/* nstp, notp, n_3d of private_mod.f95:1853-1856 */
void beom_oracle_counts(const beom_oracle *o, int *nstp, int *notp, int *n_3d);

/* Direct access to the module arrays (reference layout, see include/beom_gpu.h). */
double  *beom_oracle_array(beom_oracle *o, const char *name);
int32_t *beom_oracle_iarray(beom_oracle *o, const char *name);
double   beom_oracle_scalar(const beom_oracle *o, const char *name);
int      beom_oracle_nseg(const beom_oracle *o);
/* Fills a beom_fields whose pointers alias the oracle's arrays (tests feeding the GPU library with
 * oracle-made inputs). */
void     beom_oracle_fields(beom_oracle *o, beom_fields *f);

/* write_array's float32 records (private_mod.f95:2817-2975): var = "eta_","u___","v___","pvor",
 * "mont","v_cc"; out has ndeg*nlay floats. */
int  beom_oracle_record(beom_oracle *o, const char *var, float *out);

/* individual routines, for unit-level parity */
void beom_oracle_distribute_stress(beom_oracle *o);
void beom_oracle_update_h(beom_oracle *o);
void beom_oracle_update_mont(beom_oracle *o, int ilay);
void beom_oracle_update_viscosity(beom_oracle *o, int ilay);
void beom_oracle_update_u(beom_oracle *o, int ilay);
void beom_oracle_update_v(beom_oracle *o, int ilay);
void beom_oracle_no_gradient_obc(beom_oracle *o, int ilay);
void beom_oracle_surf_pressure(beom_oracle *o);
void beom_oracle_set_scalars(beom_oracle *o, double ctim, double ramp, double gene);

#ifdef __cplusplus
}
#endif
#endif
