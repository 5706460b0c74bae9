/* aiolos_b200 -- C ABI of the B200-native per-timestep solver core of Aiolos, batched over an ensemble of independent
 * 1-D columns.  Plain pointers and sizes only; every entry point returns 0 on success or a negative code and never
 * throws (the message is available from aio_last_error()).  One aio_ensemble lives on one GPU; calls on one handle are
 * not thread-safe (as the reference's c_Sim), distinct handles are independent.
 *
 * What each entry point replaces in the reference (/root/reference):
 *   aio_create            c_Sim::c_Sim(par, spc, dir, ...)              init_and_bounds.cpp:26   (same three strings)
 *   aio_create_from_dump  the same constructor, with its result read from an oracle state dump (oracle/ref_dump.cpp)
 *   aio_advance / aio_run c_Sim::execute()                              advection.cpp:18-366     (the time loop)
 *   aio_status            c_Sim::globalTime / dt / steps + the soft-crash flags   advection.cpp:282-346
 *   aio_get               public members c_Species::u / prim / dS, c_Sim::Jrad_FLD / phi   aiolos.h:533-670,881-957
 *   aio_upload_state / aio_download_state   std::vector<AOS> c_Species::u for all columns (host buffers; end-to-end path)
 *   aio_write_outputs     c_Species::print_AOS_component_tofile + c_Sim::print_diagnostic_file   io.cpp:704-932
 *   aio_write_monitor     c_Sim::print_monitor                                                  io.cpp:939-1072
 * The problems/<name>.cpp hook API (aiolos.h:1050-1055) is described in INTEGRATION.md.
 */
#ifndef AIOLOS_B200_H
#define AIOLOS_B200_H

#ifdef __cplusplus
extern "C" {
#endif

typedef struct aio_ensemble aio_ensemble;

/* Per-column override of one .par key (ensemble sweeps).  column < 0 applies to every column.
 * A column-specific override may change what the engine keeps per column: the initial state the constructor builds (u,
 * eint_last, phi, Jrad_FLD, dS_band), the irradiation (PARI_UVSTAR, PARI_XSTAR, PARI_LUM..., i.e. solar_heating_final), the
 * clock (dt, CFL factor, PARI_TIME_TMAX) and T_planet; PARI_PLANET_MASS as long as neither self-gravity nor a giantplanet
 * boundary reads the mass inside the kernels.  Every other quantity (switches, opacities, species constants, band layout,
 * grid) is shared by the ensemble: an override that makes it differ between columns is refused with AIO_ERR_UNSUPPORTED,
 * naming the field. */
typedef struct aio_override {
    int column;
    const char* key;      /* a PARI_* name exactly as in the .par file, e.g. "PARI_UVSTAR" */
    const char* value;    /* parsed as the .par reader would parse the second token of that line */
} aio_override;

/* error codes */
#define AIO_OK 0
#define AIO_ERR_ARG (-1)          /* bad argument */
#define AIO_ERR_IO (-2)           /* a file could not be read */
#define AIO_ERR_UNSUPPORTED (-3)  /* the deck enables a module this build does not cover (message names it) */
#define AIO_ERR_CUDA (-4)         /* no CUDA device / CUDA runtime error: there is no CPU fallback */
#define AIO_ERR_INTERNAL (-5)

/* column state, reported by aio_status (advection.cpp:282-346) */
#define AIO_COL_OK 0
#define AIO_COL_NEG_T 1
#define AIO_COL_NEG_J 2
#define AIO_COL_NAN 3

/* field ids for aio_get; names follow the reference's members */
#define AIO_F_U1 0          /* c_Species::u[j].u1  (density)            */
#define AIO_F_U2 1          /* c_Species::u[j].u2  (momentum)           */
#define AIO_F_U3 2          /* c_Species::u[j].u3  (total energy)       */
#define AIO_F_EINT_LAST 3   /* primlast[j].internal_energy              */
#define AIO_F_PHI 4         /* c_Sim::phi[j]                            */
#define AIO_F_PHI_S 5       /* c_Species::phi_s[j]                      */
#define AIO_F_DS 6          /* c_Species::dS(j)                         */
#define AIO_F_TEMP 7        /* prim[j].temperature                      */
#define AIO_F_PRES 8        /* prim[j].pres                             */
#define AIO_F_SPEED 9       /* prim[j].speed                            */
#define AIO_F_J0 16         /* c_Sim::Jrad_FLD(j, b): field = AIO_F_J0 + b */

int aio_device_count(void);
/* Measured roofline denominators of one device (no reference counterpart; bench.py reports against them): the FP64 FMA
 * rate of the SMs in TFLOP/s (2 flop per DFMA) and the device-to-device copy bandwidth in GB/s (read + write bytes).
 * Either pointer may be NULL. */
int aio_device_peaks(int device, double* fp64_tflops, double* copy_gbs);

/* Kernel-level known-answer hook for the coupled J_b / T_s block-tridiagonal solve (block_tridiag_solve.h:57-130): runs the
 * product's block-Thomas sweep on caller-supplied block rows.  coef: [ncols][num_cells2][NC] doubles; per cell bands_out
 * band records of W doubles  [D, L, U, B, JT(s = 0..S-1), TJ(s = 0..S-1), pad]  followed by num_species species records of V
 * doubles  [D, B, TT(t = 0..S-1), pad], W = (4 + 2 S) and V = (2 + S) rounded up to multiples of 4, NC = bands_out W + S V
 * (D: diagonal of d_i; L, U: the diagonal l_i / u_i entry of the J row; B: right-hand side; JT = d_i(J_b, T_s), TJ = d_i(T_s,
 * J_b), TT = off-diagonal d_i(T_s, T_t)); n = bands_out + num_species <= 24.  x: [ncols][num_cells2][n]. */
int aio_selftest_block_thomas(int ncols, int num_cells2, int bands_out, int num_species, const double* coef, double* x);

int aio_create(const char* dir, const char* par, const char* spc, int ncols, const aio_override* ov, int n_ov,
               int device, aio_ensemble** out);
int aio_create_from_dump(const char* init_dump_path, int ncols, const aio_override* ov, int n_ov, int device,
                         aio_ensemble** out);
void aio_destroy(aio_ensemble* e);

/* sizes: num_cells is the reference's num_cells (arrays hold num_cells + 2 entries) */
int aio_dims(aio_ensemble* e, int* ncols, int* num_cells, int* num_species, int* bands_in, int* bands_out);

/* Advance every column by up to max_steps further steps (a column stops at its own t_max or when it soft-crashes).
 * device_ms (may be NULL) receives the CUDA-event time of the launch sequence on the ensemble's stream. */
int aio_advance(aio_ensemble* e, long long max_steps, double* device_ms);
/* c_Sim::execute(): run until every column reached min(t_max, t_stop), crashed, or took max_steps steps (max_steps <= 0: no
 * limit; never exceeded).  *steps_done receives the steps taken (the largest per-column count). */
int aio_run(aio_ensemble* e, double t_stop, long long max_steps, long long* steps_done);
int aio_set_tmax(aio_ensemble* e, int column, double t_max);   /* column < 0: all */
/* Advance nsteps steps with a CUDA event between the kernel launches; ms3 receives the average device milliseconds per
 * step of the {hydro, friction, radiation} kernels (0 for a kernel the deck does not launch). */
int aio_kernel_times(aio_ensemble* e, long long nsteps, double* ms3);

/* Engine options.  "tridiag" = "pcr" (default: parallel cyclic reduction across the CTA) | "thomas" (validation mode:
 * the reference's elimination order, block_tridiag_solve.h:57-130, on one lane). */
int aio_set_option(aio_ensemble* e, const char* key, const char* value);

int aio_status(aio_ensemble* e, int column, double* t, double* dt, long long* steps, int* state);
int aio_status_all(aio_ensemble* e, double* t, double* dt, long long* steps, int* state);   /* arrays of ncols */
int aio_get(aio_ensemble* e, int column, int species, int field, double* dst, int n);

/* What the front end built for one column, by the reference's member name ("x_i12", "phi", "planck_matrix", "s0.u1",
 * "s1.opacity_avg_solar", "dt", ...).  Returns the number of values (copies at most `capacity` into dst; dst may be NULL),
 * or a negative error code. */
int aio_problem_field(aio_ensemble* e, int column, const char* name, double* dst, int capacity);

/* Conserved state of all columns, host layout [3][ncols][num_species][num_cells+2] (component, column, species, cell). */
int aio_upload_state(aio_ensemble* e, const double* u_host);
int aio_download_state(aio_ensemble* e, double* u_host);
/* Pipelined variant: the copy from (pinned) host memory runs on a second stream into a staging buffer while earlier steps
 * are still executing; aio_commit_upload makes the staged state current (ordered on the device by events: it does not wait on
 * the host).  u_host must stay valid and unchanged until the next call that synchronises with the ensemble's stream
 * (aio_advance, aio_status*, aio_column_mass, aio_download_state).  The first step after an upload still takes its dt from
 * the CFL / energy-change numbers of the state that was replaced. */
int aio_upload_state_async(aio_ensemble* e, const double* u_host);
int aio_commit_upload(aio_ensemble* e);
/* Sum over all columns/species/cells of u1 (mass per column, weighted by cell volume) -- a small result to read back. */
int aio_column_mass(aio_ensemble* e, double* mass_per_column);

int aio_write_outputs(aio_ensemble* e, int column, int index, const char* outdir);
/* c_Sim::print_monitor (io.cpp:939-1072): append one row for this column's current state to <outdir>monitor_<stem>.dat
 * (outdir NULL/"" = the deck directory); truncate != 0 starts a new file, as the reference does at step 0. */
int aio_write_monitor(aio_ensemble* e, int column, const char* outdir, int truncate);

long long aio_kernel_launches(aio_ensemble* e);
/* Name(s) of the kernel(s) behind slot 0/1/2 (hydro or fused step, friction, radiation) of aio_kernel_times for this
 * ensemble, e.g. "k_step1<4,256,3>": what bench.py names in its roofline records. */
int aio_kernel_name(aio_ensemble* e, int slot, char* dst, int capacity);
long long aio_total_steps(aio_ensemble* e);      /* sum over columns of completed steps */
const char* aio_last_error(void);
const char* aio_version(void);

#ifdef __cplusplus
}
#endif
#endif
