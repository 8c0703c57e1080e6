/*
 * pyxrd_b200 -- C-ABI of the B200 (sm_100a) implementation of PyXRD's matrix-method
 * hot path (pyxrd.calculations: phase -> diffracted intensity -> specimen pattern ->
 * mixture residual).
 *
 * The reference has no FFI of its own: its boundary is a set of Python module-level
 * functions taking DataObject graphs.  The Python host layer in
 * pyxrd_b200/ flattens those graphs ("packing") into the plain arrays below and calls
 * these entry points through ctypes; INTEGRATION.md shows the binding.  Each entry
 * point names the reference function(s) (path:line under the reference tree) whose
 * arithmetic it replaces.
 *
 * Conventions: every function returns 0 on success, non-zero on failure with a
 * message available from pyxrd_last_error(); all pointers are HOST pointers unless a
 * parameter name ends in _dev; all floating point is IEEE double; all work is queued
 * on the context's CUDA stream and the pyxrd_read_* calls synchronise that stream.
 * There is no CPU fallback: without a CUDA device pyxrd_create() fails.
 */
#ifndef PYXRD_B200_H
#define PYXRD_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct pyxrd_ctx pyxrd_ctx;

/* layout strides of the packed parameter tables */
#define PYXRD_GONIO_STRIDE 12 /* wavelength, soller1, soller2, mcr_2theta, divergence_mode (0 FIXED, 1 AUTOMATIC, 2 other), divergence, has_absorption_correction, absorption, sample_surf_density, sample_length, radius, has_observations (0: the specimen's residual is 0.0, mixture.py:247-251) */
#define PYXRD_ATYPE_STRIDE 12 /* par_a[5], par_b[5], par_c, debye */
#define PYXRD_PHASE_I_STRIDE 8 /* G, rank, csds_minimum, csds_maximum, flags, comp_list_begin, matrix_offset, reserved */
#define PYXRD_PHASE_D_STRIDE 8 /* sigma_star, csds_average, alpha_scale, alpha_offset, beta_scale, beta_offset, reserved x2 */
#define PYXRD_COMP_D_STRIDE 8  /* d001, default_c, delta_c, lattice_d, volume, weight, reserved x2 */
#define PYXRD_COMP_I_STRIDE 4  /* atom_begin, n_layer_atoms, n_interlayer_atoms, reserved */
#define PYXRD_ATOM_D_STRIDE 2  /* pn, default_z */

#define PYXRD_FLAG_VALID_PROBS 1      /* PhaseData.valid_probs (phases.py:109-111) */
#define PYXRD_FLAG_APPLY_LPF 2        /* PhaseData.apply_lpf (phases.py:88) */
#define PYXRD_FLAG_APPLY_CORRECTION 4 /* PhaseData.apply_correction (specimen.py:57) */
#define PYXRD_FLAG_RAW_PATTERN 8       /* RawPatternPhase (phases.py:97-104): G = rank = 0, csds_min = n, matrix_offset -> raw_x[n] | raw_y[n] */

#define PYXRD_MAX_RANK 64
#define PYXRD_MAX_COMPONENTS 8

/* Candidate-invariant tables: what SpecimenData / GonioData / AtomTypeData carry
 * (data_objects.py:35-47,118-172,214-239).  Ragged per-specimen arrays are
 * concatenated in specimen order. */
typedef struct {
    int32_t n_specimens;      /* S */
    int32_t n_atom_types;     /* T */
    int32_t n_patterns;       /* Z = len(specimen.z_list) (mixture.py:133-138) */
    int32_t reserved;
    const int32_t *spec_npts; /* [S] X_s = len(range_theta) */
    const double *theta;      /* [sum X_s] SpecimenData.range_theta (radians) */
    const double *spec_gonio; /* [S][PYXRD_GONIO_STRIDE] */
    const int32_t *spec_nwl;  /* [S] len(goniometer.wavelength_distribution) */
    const double *wl_fraction;/* [sum nwl] the f_j of the list, in list order */
    /* per (specimen, wavelength j, point x), concatenated: the abscissa part of
     * interp1d(asin(stl*lambda_j/2), ., fill 0)(theta) (specimen.py:66-69): index of the
     * upper neighbour (or -1 when theta_x lies outside) and the two lerp weights */
    const int32_t *wl_index;
    const double *wl_w_hi;
    const double *wl_w_lo;
    const double *observed;   /* [S][Z][X_s] = observed_intensity.T */
    const uint8_t *selected;  /* [sum X_s] SpecimenData.selected_range */
    const double *atom_types; /* [T][PYXRD_ATYPE_STRIDE] */
    /* optional [sum X_s]: use these 2 sin(theta)/lambda values instead of deriving them from
     * theta and the wavelength (get_diffracted_intensity / get_intensity take range_stl as an
     * argument of their own, phases.py:68,81); NULL otherwise */
    const double *stl_override;
} pyxrd_static_desc;

/* Per-candidate parameters: what PhaseData / ComponentData / AtomData / CSDSData /
 * MixtureData carry (data_objects.py:49-116,174-271). */
typedef struct {
    int32_t n_candidates;     /* K trials */
    int32_t n_phase_slots;    /* M = mixture.m (columns of the phase matrix) */
    int32_t n_phases;         /* NP parameter blocks */
    int32_t n_components;     /* NC */
    int32_t n_atoms;          /* NA */
    int32_t n_phase_comp;     /* length of phase_comp */
    int64_t n_matrix;         /* length of matrices */
    const int32_t *phase_i;   /* [NP][PYXRD_PHASE_I_STRIDE] */
    const double *phase_d;    /* [NP][PYXRD_PHASE_D_STRIDE] */
    const int32_t *phase_comp;/* component index lists */
    const double *matrices;   /* per phase at matrix_offset: W[r*r] then P[r*r], row major */
    const double *comp_d;     /* [NC][PYXRD_COMP_D_STRIDE] */
    const int32_t *comp_i;    /* [NC][PYXRD_COMP_I_STRIDE] */
    const double *atom_d;     /* [NA][PYXRD_ATOM_D_STRIDE] */
    const int32_t *atom_type; /* [NA] index into atom_types, -1 = None */
    const int32_t *job_phase; /* [K][S][Z][M] parameter-block index, -1 = None phase */
    const double *fractions;  /* [K][M]  MixtureData.fractions */
    const double *scales;     /* [K][S]  MixtureData.scales */
    const double *bgshifts;   /* [K][S]  MixtureData.bgshifts (already 0 when settings.BGSHIFT is off) */
    /* optional (both NULL: every phase carries W / P in `matrices`): probability model of each phase
     * (PYXRD_PROB_*) and its parameters; for a model-driven phase the device computes W, P into the
     * phase's pool entry and valid_probs itself (probabilities/models/*.py, phases/models/phase.py:45) */
    const int32_t *prob_model;  /* [NP] */
    const double *prob_params;  /* [NP][PYXRD_PROB_STRIDE] */
} pyxrd_batch_desc;

/* Probability models evaluated on the device.  Parameters in the order of the
 * reference's properties, clamped to the properties' bounds like the setters do
 * (mvc/models/properties/cast_property.py:44-47):
 *   R0   (any G <= 6): F1 .. F(G-1)                                   R0models.py:68-83
 *   R1G2: W1, P11_or_P22                                              R1models.py:125-139
 *   R2G2: W1 (>= 1/2), P112_or_P211, P21, P122_or_P221                R2models.py:195-233
 *   R3G2: W1 (>= 2/3), P1111_or_P2112                                 R3models.py:158-198
 *   R1G3: W1, P11_or_P22, G1, G2, G3, G4                              R1models.py:367-410
 *   R1G4: W1, P11_or_P22, R1, R2, G1, G2, G11, G12, G21, G22, G31, G32  R1models.py:747-797
 *   R2G3: W1 (>= 1/2), P111_or_P212, G1, G2, G3, G4                   R2models.py:485-535
 * followed by solve() and validate() (base_models.py:142-241). */
#define PYXRD_PROB_FIXED 0
#define PYXRD_PROB_R0 1
#define PYXRD_PROB_R1G2 2
#define PYXRD_PROB_R2G2 3
#define PYXRD_PROB_R3G2 4
#define PYXRD_PROB_R1G3 5
#define PYXRD_PROB_R1G4 6
#define PYXRD_PROB_R2G3 7
#define PYXRD_PROB_STRIDE 12
#define PYXRD_PROB_MAX_RANK 9   /* R2G3; W_out / P_out of pyxrd_leaf_probabilities hold up to 81 doubles */

/* ---- population from a template ---------------------------------------
 * A refinement proposes K solution vectors x_k of d refinable values.  Instead of rebuilding K
 * DataObject graphs and packing each of them (refinement/refiner.py:91-109,
 * mixture/models/mixture.py:53-85: 6.4 ms per trial in the reference, 70 us per trial in packing.py),
 * the template candidate is packed once and a binding table says which packed slot every refinable
 * drives: slot = scale * x[param] + offset (or, for a slot several refinables drive together, slot *= ... / slot += ...:
 * a cell volume a * b * d001 with b and d001 both refined, a component weight summed over two refined atom contents;
 * phases/models/unit_cell_prop.py:199, atom_relations.py:244,363,498).  Rows apply in table order.  The expansion runs on the device (pyxrd_set_population and the
 * calls built on it; pyxrd_expand_population is the same expansion as native host code, for inspection and
 * tests); everything numeric that depends on the bound values -- probability matrices, CSDS distribution,
 * absolute scale, z-stretch -- is computed on the device by the batch prologue. */
#define PYXRD_BIND_PHASE_D 0    /* phase_d[index]: phase * 8 + {0 sigma_star, 1 CSDS average, 2..5 alpha / beta} */
#define PYXRD_BIND_COMP_D 1     /* comp_d[index]: component * 8 + {0 d001, 1 default_c, 2 delta_c, 3 lattice_d, 4 volume, 5 weight} */
#define PYXRD_BIND_ATOM_D 2     /* atom_d[index]: atom * 2 + {0 pn, 1 default_z} */
#define PYXRD_BIND_MATRIX 3     /* matrices[index] */
#define PYXRD_BIND_PROB_PARAM 4 /* prob_params[index]: phase * PYXRD_PROB_STRIDE + parameter number */
#define PYXRD_BIND_FRACTION 5   /* fractions[index], index < M */
#define PYXRD_BIND_SCALE 6      /* scales[index], index < S */
#define PYXRD_BIND_BGSHIFT 7    /* bgshifts[index], index < S */
#define PYXRD_BIND_MODE_SET 0       /* slot = v */
#define PYXRD_BIND_MODE_MULTIPLY 1  /* slot = slot * v (after the row that set it) */
#define PYXRD_BIND_MODE_ADD 2       /* slot = slot + v */
typedef struct {
    int32_t param;              /* column of x */
    int32_t target;             /* PYXRD_BIND_* */
    int32_t index;              /* flat index inside the TEMPLATE's array */
    int32_t mode;               /* PYXRD_BIND_MODE_*: with v = scale * x[param] + offset */
    double scale, offset;
} pyxrd_binding;

typedef struct {
    int32_t n_candidates;       /* K */
    int32_t n_params;           /* d */
    int32_t n_bindings;
    int32_t n_specimens;        /* S of the static tables (scales / bgshifts per candidate) */
    int32_t n_patterns;         /* Z of the static tables (the template's job table is [S][Z][M]) */
    int32_t reserved;
    const double *x;            /* [K][d] */
    const pyxrd_binding *bindings;
    /* optional [NP of the template]: 1 = after binding set CSDS maximum = int(csds_factor * average)
     * (phases/models/CSDS.py:127,140 with settings.LOG_NORMAL_MAX_CSDS_FACTOR, settings.py:26) */
    const uint8_t *csds_auto_maximum;
    double csds_factor;
} pyxrd_population_desc;

typedef struct pyxrd_expanded pyxrd_expanded;
/* host only (no device needed): the K-candidate batch the template and the table describe */
int pyxrd_expand_population(const pyxrd_batch_desc *template_desc, const pyxrd_population_desc *pop,
                            pyxrd_expanded **out);
void pyxrd_expanded_desc(const pyxrd_expanded *e, pyxrd_batch_desc *out);
void pyxrd_expanded_free(pyxrd_expanded *e);
const char *pyxrd_expand_error(void);   /* message of the last failed pyxrd_expand_population of this thread */
/* the same expansion on the DEVICE (k_expand_population): the packed template, x[K][d] and the table are uploaded
 * (K * d * 8 bytes per generation instead of K parameter blocks), every candidate's slice of the batch arrays, its
 * job table and job lists are written by one kernel, then the batch prologue runs as after pyxrd_set_batch */
int pyxrd_set_population(pyxrd_ctx *ctx, const pyxrd_batch_desc *template_desc, const pyxrd_population_desc *pop);
/* The next generation of the SAME refinement (same template, table, K and d as the resident population; refinement/
 * refine_async_helper.py:16-23 evaluates generation after generation): only the solution vectors x[K][d] cross PCIe
 * (NULL: keep the resident vectors), the expansion kernel and the batch prologue run again.  Fails -- and the caller
 * falls back to pyxrd_set_population -- when no population is resident (any pyxrd_set_static / pyxrd_set_batch in
 * between drops it). */
int pyxrd_update_population(pyxrd_ctx *ctx, const double *x);
/* set_population + compute_intensities + compute_residuals -> residuals [K][1+S] */
int pyxrd_evaluate_population(pyxrd_ctx *ctx, const pyxrd_batch_desc *template_desc, const pyxrd_population_desc *pop,
                              double *residuals_out);
/* set_population + compute_intensities + optimize -> residuals [K][1+S], solutions [K][M+2S] (may be NULL) */
int pyxrd_optimize_population(pyxrd_ctx *ctx, const pyxrd_batch_desc *template_desc, const pyxrd_population_desc *pop,
                              int refine_bg, int outer_iterations, int newton_iterations, double *residuals_out,
                              double *solutions_out);
/* one probability model (GUI previews, tests): W, P f64[rank*rank] row major, valid_probs, rank */
int pyxrd_leaf_probabilities(pyxrd_ctx *ctx, int model, int G, const double *params, double *W_out, double *P_out,
                             int *valid_out, int *rank_out);

/* ---- lifetime ------------------------------------------------------------------ */
int pyxrd_create(int device, pyxrd_ctx **out);
void pyxrd_destroy(pyxrd_ctx *ctx);
const char *pyxrd_last_error(const pyxrd_ctx *ctx);
/* use an existing cudaStream_t (e.g. torch's current stream); NULL = the context's own */
int pyxrd_set_stream(pyxrd_ctx *ctx, void *cuda_stream);
int pyxrd_synchronize(pyxrd_ctx *ctx);
/* number of kernels this context has launched so far */
int64_t pyxrd_launch_count(const pyxrd_ctx *ctx);
/* Explicit measurement / experiment switches of one context (the library reads nothing from the environment; the
 * defaults are what the product runs).  Names: "use_structure" (1; 0 = treat structured transition matrices as dense),
 * "series_mma" (1; 0 = CUDA-core series for the ranks that have an FP64 tensor-core kernel), "overlap_classes" (1; 0 =
 * job classes one after the other), "layer_tables" (1; 0 = candidate-invariant layer stacks are evaluated inside the
 * phase kernels instead of being tabulated once per batch), "optimizer_threads" (0 = by population size; 128 | 256), "host_chunks" (0 = by
 * population size; candidate chunks of pyxrd_evaluate_host), "optimizer_delta0" (1e-2), "optimizer_shrink" (0.5),
 * "optimizer_delta_min" (1e-7): weight-floor schedule of pyxrd_optimize relative to the mean observed intensity;
 * "optimizer_newton_tol" (1e-5): relative decrease of the majoriser below which a pass takes no further Newton iteration;
 * "optimizer_single_from" (0 = off): pass from which one Newton iteration is taken whatever the decrease. */
int pyxrd_set_option(pyxrd_ctx *ctx, const char *name, double value);

/* ---- upload -------------------------------------------------------------------- */
/* Uploads the tables and runs the table kernels: sin(theta), stl = 2 sin(theta)/lambda
 * (specimen.py:50), cos^2(2 theta), machine correction (specimen.py:21-43,
 * goniometer.py:36-37) and the atomic scattering factors (atoms.py:13-38) per
 * (atom type, point). */
int pyxrd_set_static(pyxrd_ctx *ctx, const pyxrd_static_desc *desc);
/* Uploads K candidates and runs the per-phase prologue: CSDS distribution + mean
 * (CSDS.py:13-46, math_tools.py:36-37), progression factors (phases.py:147-151),
 * absolute scale (phases.py:48-66), interlayer z-stretch (components.py:15-16,42-50). */
int pyxrd_set_batch(pyxrd_ctx *ctx, const pyxrd_batch_desc *desc);
/* Replace only the mixture solution of the resident batch ([K][M], [K][S], [K][S]). */
int pyxrd_set_solution(pyxrd_ctx *ctx, const double *fractions, const double *scales,
                       const double *bgshifts);

/* Load phase intensities computed earlier (SpecimenData.phase_intensities, [K][S][M][Z][X_s])
 * into the resident batch instead of computing them (calculate_scaled_intensities on a
 * specimen whose patterns already exist, specimen.py:84-88). */
int pyxrd_set_intensities(pyxrd_ctx *ctx, const double *intensities, int64_t count);

/* ---- compute (asynchronous on the stream) ---------------------------------------- */
/* All phase patterns of the resident batch: structure factors (atoms.py:40-67,
 * components.py:18-54, phases.py:20-39), the CSDS-weighted series Re Tr(F W S)*scale
 * (phases.py:106-158, math_tools.py:16-17), Lorentz-polarisation (goniometer.py:15-34,
 * phases.py:81-95), machine correction and the wavelength distribution
 * (specimen.py:45-82).  Result: f64 [K][S][M][Z][X_s] on the device. */
int pyxrd_compute_intensities(pyxrd_ctx *ctx);
/* Scale / sum / background and Rp per specimen (specimen.py:84-88,16-19,
 * statistics.py:22-27, mixture.py:222-257).  Result: residuals f64 [K][1+S]
 * ([k][0] = mean over specimens), totals f64 [K][S][Z][X_s]; flags: PYXRD_WANT_SCALED
 * also stores scaled_phase_intensities f64 [K][S][M][Z][X_s], PYXRD_RESIDUALS_ONLY skips the totals. */
#define PYXRD_WANT_SCALED 1     /* also scaled_phase_intensities (calculate_mixture hands them back) */
#define PYXRD_RESIDUALS_ONLY 2  /* the total patterns are summed on the fly but not stored (a refinement trial only
                                 * needs residuals [K][1+S]); pyxrd_read_totals fails afterwards */
int pyxrd_compute_residuals(pyxrd_ctx *ctx, int flags);
/* pyxrd_compute_intensities + pyxrd_compute_residuals as one call (calculate_mixture for the resident batch); the last
 * pass of the wavelength distribution runs inside the residual kernel (the patterns are not read a second time); where
 * the gathers of the distribution stay within 128 points (K-alpha doublets) ALL passes run inside it.  With
 * PYXRD_RESIDUALS_ONLY alone (Rp) that kernel does not store the final phase patterns either: afterwards
 * pyxrd_read_intensities / pyxrd_compute_residuals / pyxrd_optimize fail until the patterns are computed again. */
int pyxrd_compute_all(pyxrd_ctx *ctx, int flags);

/* Which pattern residual pyxrd_compute_residuals / pyxrd_evaluate_host report (settings.py:27-32,
 * mixture.py:22-27,89): PYXRD_RESIDUAL_RP = statistics.py:22-27 (default), PYXRD_RESIDUAL_RPDER =
 * statistics.py:29-48 (Rp of the first derivatives of the 31-tap-Blackman-smoothed clipped patterns,
 * math_tools.py:46-99).  "Rpw" and "Rphase" cannot be evaluated by the reference itself (they raise
 * for every 2-D input, statistics.py:50-68,81 / mixture.py:89); the host layer re-raises the same
 * exceptions.  The device optimiser (pyxrd_optimize) always minimises Rp. */
#define PYXRD_RESIDUAL_RP 0
#define PYXRD_RESIDUAL_RPDER 1
int pyxrd_set_residual_method(pyxrd_ctx *ctx, int method);

/* Device-resident optimiser of the mixture solution of every candidate of the resident batch
 * (optimize_mixture, mixture.py:150-219: minimise the mean Rp over fractions >= 0, scales > 0,
 * bgshifts >= 0).  Not a port of L-BFGS-B: majorise-minimise on the L1 objective;
 * the contract is residual <= the reference's optimum * (1 + 5e-3).  Needs computed intensities;
 * all fractions and all scales must be free (the reference's default masks); refine_bg = 0 keeps
 * the batch's bgshifts.  Results: pyxrd_read_solutions f64 [K][M + 2S] = fractions (sum 1) |
 * scales | bgshifts, pyxrd_read_opt_residuals f64 [K][1+S]. */
int pyxrd_optimize(pyxrd_ctx *ctx, int refine_bg, int outer_iterations, int newton_iterations);
int pyxrd_read_solutions(pyxrd_ctx *ctx, double *out, int64_t count);
int pyxrd_read_opt_residuals(pyxrd_ctx *ctx, double *out, int64_t count);

/* ---- results (synchronise, device -> host) ---------------------------------------- */
int64_t pyxrd_intensity_count(const pyxrd_ctx *ctx); /* doubles in the intensity buffer */
int64_t pyxrd_total_count(const pyxrd_ctx *ctx);     /* doubles in the totals buffer */
int pyxrd_read_intensities(pyxrd_ctx *ctx, double *out, int64_t count);
int pyxrd_read_scaled(pyxrd_ctx *ctx, double *out, int64_t count);
int pyxrd_read_totals(pyxrd_ctx *ctx, double *out, int64_t count);
int pyxrd_read_residuals(pyxrd_ctx *ctx, double *out, int64_t count);
int pyxrd_read_correction(pyxrd_ctx *ctx, double *out, int64_t count); /* [sum X_s] */
/* device addresses of the result buffers, for zero-copy consumers (NCCL gather) */
int pyxrd_device_buffers(pyxrd_ctx *ctx, double **intensities_dev, double **totals_dev,
                         double **residuals_dev);

/* device addresses of the optimiser's results (pyxrd_optimize): residuals [K][1+S], solutions [K][M+2S] */
int pyxrd_device_results(pyxrd_ctx *ctx, double **opt_residuals_dev, double **solutions_dev);

/* The exchange step of the sharded path (refinement/async_evaluatable.py:38-50 collects the results of
 * a generation's tasks): rows [0, n_rows) of the device table src_dev[n_rows][cols] (this rank's residual rows,
 * pyxrd_device_buffers / pyxrd_device_results) are stored into the gathered table of EVERY rank at row dst_row0.
 * peer_tables: n_peers (<= 16) DEVICE pointers, one per rank, this rank's own included -- the peers' tables mapped into
 * this process over NVLink / NVSwitch by the caller (torch symmetric memory, CUDA IPC).  Plain stores + a system-scope
 * fence on the context's stream; the caller's cross-GPU barrier orders them against the peers' reads. */
int pyxrd_put_rows_to_peers(pyxrd_ctx *ctx, const double *src_dev, int64_t n_rows, int cols, void *const *peer_tables,
                            int n_peers, int64_t dst_row0);

/* ---- one call, host buffers in and out (what a per-generation plugin call does) ---- */
/* set_batch + compute_intensities + compute_residuals + read residuals [K][1+S];
 * intensities_out may be NULL (a refinement trial only needs the residuals). */
int pyxrd_evaluate_host(pyxrd_ctx *ctx, const pyxrd_batch_desc *desc, double *residuals_out,
                        double *intensities_out);

/* The per-generation call of a refinement: set_batch + compute_intensities + optimize + read;
 * residuals_out [K][1+S] (get_optimized_residual of every trial = residuals_out[k][0]),
 * solutions_out [K][M+2S] (may be NULL). */
int pyxrd_optimize_host(pyxrd_ctx *ctx, const pyxrd_batch_desc *desc, int refine_bg, int outer_iterations,
                        int newton_iterations, double *residuals_out, double *solutions_out);

/* ---- leaf helpers (GUI previews; also the unit-level parity hooks) ---------------- */
/* get_atomic_scattering_factor (atoms.py:13-38): s -> asf, n points, one atom type */
int pyxrd_leaf_asf(pyxrd_ctx *ctx, const double *s, int64_t n, const double *atom_type, double *out);
/* pyxrd/calculations/statistics.py on plain arrays (GUI pattern statistics, specimen/models/statistics.py:16,115-118):
 * kind 0 Rp (:22-27), 1 Rpw (:50-68), 2 R_squared (:12-20), 3 Rpe with extra = number of parameters (:70-79),
 * 4 Rphase with the phase pattern in `phase` (:81-100); calc / phase may be NULL where the statistic does not use them */
int pyxrd_leaf_statistic(pyxrd_ctx *ctx, int kind, const double *exp, const double *calc, const double *phase, int64_t n,
                         double extra, double *out);
/* derive() = np.gradient(smooth(pattern, 15)) (statistics.py:29-41, math_tools.py:46-99: 31-tap Blackman window with
 * reflected ends); smooth_only != 0: smooth_pattern().  n >= 31, else the reference's ValueError text is the error */
int pyxrd_leaf_derive(pyxrd_ctx *ctx, const double *pattern, int64_t n, int smooth_only, double *out);
/* get_lorentz_polarisation_factor (goniometer.py:15-34) */
int pyxrd_leaf_lpf(pyxrd_ctx *ctx, const double *theta, int64_t n, double sigma_star,
                   double soller1, double soller2, double mcr_2theta, double *out);
/* get_T (goniometer.py:20-25): the Lorentz part of that factor */
int pyxrd_leaf_lorentz(pyxrd_ctx *ctx, const double *theta, int64_t n, double sigma_star, double soller1, double soller2, double *out);
/* calculate_distribution (CSDS.py:13-46): out_arr[maximum+1], out_mean[1];
 * progression factors (phases.py:147-151) into out_coef[maximum+2] (index = power of Q) */
int pyxrd_leaf_csds(pyxrd_ctx *ctx, const double *csds6, int minimum, int maximum,
                    double *out_arr, double *out_mean, double *out_coef);
/* get_factors (components.py:18-54) / get_structure_factor (atoms.py:40-67) of ONE component on
 * an arbitrary stl axis: SF and phi as interleaved complex [n]; z_out[NA] = the stretched atom
 * positions the reference writes back into atom.z (components.py:42-44).  atom_type_rows holds
 * one PYXRD_ATYPE_STRIDE row per atom; atom_has_type[a] == 0 marks a None atom / atom type. */
int pyxrd_leaf_component(pyxrd_ctx *ctx, const double *stl, int64_t n, const double *comp_d,
                         int n_layer, int n_interlayer, const double *atom_d,
                         const double *atom_type_rows, const uint8_t *atom_has_type,
                         double *sf_out, double *phi_out, double *z_out);

/* mmult (math_tools.py:16-17): out[x] = a[x] b[x], n_points complex rank x rank matrices (interleaved re / im, row
 * major), with numpy's rounding of the reference's broadcast-and-sum (no FMA, contraction index ascending) */
int pyxrd_leaf_mmult(pyxrd_ctx *ctx, const double *a, const double *b, int64_t n_points, int rank, double *out);
/* get_Q_matrices (phases.py:41-46): out[n][x] = q[x]^(n+1), n = 0 .. n_max, by repeated mmult.  The phase kernels never
 * form this stack ((N + 1) X r^2 complex numbers: 2.96 GB for BASELINE configs[2]); this is the reference's helper as such. */
int pyxrd_leaf_matrix_powers(pyxrd_ctx *ctx, const double *q, int64_t n_points, int rank, int n_max, double *out);

/* ---- measurement helpers ----------------------------------------------------------- */
/* CUDA-event durations (ms) of the most recent compute calls, on the context's stream:
 * ms_out[0] = phase-intensity kernels, [1] = wavelength kernel, [2] = mixture residual kernels,
 * [3] = per-candidate prologue (set_batch).  Synchronises the stream. */
int pyxrd_stage_times(pyxrd_ctx *ctx, double *ms_out);
/* dependent-chain-free DFMA microbenchmark; returns achieved FP64 TFLOP/s (FMA = 2 flop) */
int pyxrd_dfma_peak(pyxrd_ctx *ctx, int iterations, double *tflops_out, double *ms_out);

#ifdef __cplusplus
}
#endif
#endif /* PYXRD_B200_H */
