/* swk.h - C ABI of the B200-native shallow-water kernel library (libswk.so).
 *
 * Drop-in boundary for ONE hot path of ANUGA: the per-timestep discontinuous-
 * elevation (DE0/DE1/DE2) update of anuga.shallow_water.Domain.  Everything is
 * `extern "C"`, plain pointers and sizes, no Python.h, no torch types.  Host
 * arrays use the reference's own layouts (C-contiguous numpy: int64 indices,
 * FP64 values); the library converts them to its device layout (32-byte SoA
 * records, int32 connectivity, locality-reordered) on upload.
 *
 * Two layers:
 *
 *  (1) RESIDENT layer  - swk_create / swk_set_* / swk_evolve / swk_get_* ...
 *      The arrays live in HBM for the life of the handle; the time loop
 *      (generic_domain.py:1835-1912, evolve_one_{euler,rk2,rk3}_step :1914-2179,
 *      update_timestep :2349-2415) runs on the device, the host only sees
 *      scalars until a yield.  This is what the new multiprocessor_mode uses.
 *
 *  (2) PER-CALL layer  - swk_call_*  (host arrays in, host arrays out)
 *      One entry point per function the reference's Cython FFI binds for this
 *      path (anuga/shallow_water/sw_domain_openmp_ext.pyx:371-459,
 *      anuga/abstract_2d_finite_volumes/quantity_ext.pyx:38-117), with the same
 *      argument meaning and in-place update of the same host arrays, so that
 *      the dispatch sites in shallow_water_domain.py:1855-1874, 1893-1915,
 *      2022-2039, 2114-2154 and friction.py:40-73 keep working for per-call use
 *      (differential tests in the style of shallow_water/tests/test_DE_openmp.py).
 *
 * Every function returns an int status (SWK_OK = 0, negative = error) and never
 * throws or aborts.  swk_last_error() returns a thread-local message.
 * A handle is not thread-safe; different handles are independent.
 * There is NO CPU fallback: every entry point fails with SWK_ERR_CUDA when no
 * sm_100 device is usable.
 */
#ifndef SWK_H
#define SWK_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SWK_ABI_VERSION 2   /* 2: value tables per substep, swk_step_*, swk_comm_*, table boundaries, explicit forcing */

/* ---- status codes ------------------------------------------------------- */
#define SWK_OK                 0
#define SWK_ERR_CUDA          -1   /* no device / CUDA runtime error */
#define SWK_ERR_ARG           -2   /* bad argument */
#define SWK_ERR_DENOMINATOR   -3   /* semi-implicit denominator <= 0: the reference's
                                      _update returns -1 (quantity.c:806-808) */
#define SWK_ERR_SMALLSTEP     -4   /* "Too small timestep ... even after N steps of 1 order
                                      scheme" (generic_domain.py:2377-2388) */
#define SWK_ERR_OVERSHOOT     -5   /* time overshot finaltime (generic_domain.py:1873-1878) */
#define SWK_ERR_UNSUPPORTED   -6   /* feature outside the hot-path scope */
#define SWK_ERR_NCCL          -7

typedef struct swk_domain swk_domain;     /* opaque, owns all device memory */

/* ---- scalar parameters -----------------------------------------------------
 * Same names and meaning as the scalars of the reference's `struct domain`
 * (shallow_water/sw_domain.h:16-36) plus the time-loop constants the reference
 * keeps on the Python object (generic_domain.py:2349-2415, config.py).       */
typedef struct {
  double epsilon;                 /* config.py:12   1e-12 */
  double H0;                      /* config.py:186  1e-5  */
  double g;                       /* config.py:45   9.8   */
  double minimum_allowed_height;
  double maximum_allowed_speed;
  double evolve_max_timestep;     /* config.py:153  1000  */
  double evolve_min_timestep;     /* config.py:154  1e-6  */
  double beta_w, beta_w_dry, beta_uh, beta_uh_dry, beta_vh, beta_vh_dry;
  double CFL;
  double fixed_flux_timestep;     /* <= 0 : variable timestepping (generic_domain.py:2332-2360) */
  int64_t extrapolate_velocity_second_order;
  int64_t low_froude;             /* 0, 1, 2 (sw_domain_openmp.c:178-194) */
  int64_t timestepping_method;    /* 1 euler, 2 rk2, 3 rk3 = timestep_fluxcalls */
  int64_t use_sloped_mannings;    /* friction.py:65-73 */
  int64_t max_smallsteps;         /* config.py: 50 */
  int64_t default_order;          /* 1 or 2 (only the bookkeeping of update_timestep uses it) */
  int64_t ghost_layer_width;      /* rk2 skips the mid-step exchange when >= 4 (generic_domain.py:2014) */
  int64_t centroid_transmissive_bc;
  int64_t track_max_speed;        /* 1: the resident time loop also writes max_speed[k] (:709-710);
                                     the per-call layer always does */
} swk_params;

/* ---- static mesh description (host pointers, borrowed during swk_create) ---
 * Field names follow `struct domain` / the Mesh attributes they come from
 * (neighbour_mesh.py:94-97, general_mesh.py:134-148, generic_domain.py:285).   */
typedef struct {
  int64_t number_of_elements;            /* N */
  int64_t boundary_length;               /* M */
  const int64_t *neighbours;             /* (N,3)  <0: boundary index -(m+1) */
  const int64_t *neighbour_edges;        /* (N,3) */
  const int64_t *surrogate_neighbours;   /* (N,3) */
  const int64_t *number_of_boundaries;   /* (N,)  */
  const int64_t *tri_full_flag;          /* (N,)  1 full, 0 ghost */
  const double *normals;                 /* (N,6) */
  const double *edgelengths;             /* (N,3) */
  const double *radii;                   /* (N,)  */
  const double *areas;                   /* (N,)  */
  const double *centroid_coordinates;    /* (N,2) */
  const double *edge_coordinates;        /* (3N,2) edge midpoints */
  const double *vertex_coordinates;      /* (3N,2) (sloped Manning only; may be NULL) */
  const int64_t *boundary_cells;         /* (M,) */
  const int64_t *boundary_edges;         /* (M,) */
  /* riverwall tables (shallow_water_domain.py:371-399); all NULL/0 when absent */
  int64_t number_of_riverwall_edges;
  int64_t ncol_riverwall_hydraulic_properties;
  const int64_t *edge_flux_type;         /* (3N,) */
  const int64_t *edge_river_wall_counter;/* (3N,) */
  const double *riverwall_elevation;
  const int64_t *riverwall_rowIndex;
  const double *riverwall_hydraulic_properties;
  /* optional locality reordering: permutation[new] = old.  NULL = keep order.
   * All host-facing arrays and ids stay in the caller's ("old") order.        */
  const int64_t *permutation;
} swk_mesh;

/* ---- quantity ids for swk_set_quantity / swk_get_quantity ------------------ */
enum {
  SWK_Q_STAGE_C = 0, SWK_Q_XMOM_C = 1, SWK_Q_YMOM_C = 2,      /* (N,)   conserved, centroid  */
  SWK_Q_ELEVATION_C = 3, SWK_Q_FRICTION_C = 4,                /* (N,)   static inputs        */
  SWK_Q_HEIGHT_C = 5,                                         /* (N,)   get only: max(w-z,0) */
  SWK_Q_STAGE_E = 10, SWK_Q_XMOM_E = 11, SWK_Q_YMOM_E = 12,   /* (N,3)  edge values          */
  SWK_Q_HEIGHT_E = 13, SWK_Q_ELEVATION_E = 14,                /*        (14: get only)       */
  SWK_Q_STAGE_V = 20, SWK_Q_XMOM_V = 21, SWK_Q_YMOM_V = 22,   /* (N,3)  get only: from edges */
  SWK_Q_HEIGHT_V = 23, SWK_Q_ELEVATION_V = 24,                /*   (24: settable, sloped Manning) */
  SWK_Q_STAGE_B = 30, SWK_Q_XMOM_B = 31, SWK_Q_YMOM_B = 32,   /* (M,)   boundary values      */
  SWK_Q_STAGE_EU = 40, SWK_Q_XMOM_EU = 41, SWK_Q_YMOM_EU = 42,/* (N,)   explicit_update      */
  SWK_Q_MAX_SPEED = 50,                                       /* (N,)   get only             */
  SWK_Q_STAGE_BACKUP = 60, SWK_Q_XMOM_BACKUP = 61, SWK_Q_YMOM_BACKUP = 62
};

/* ---- boundary kinds (a4 of SURVEY.md section 8) --------------------------------
 * Only stage/xmom/ymom boundary values feed the flux (sw_domain_openmp.c:556-560). */
enum {
  SWK_BC_NONE = 0,                 /* boundary_map[tag] is None: values left untouched */
  SWK_BC_REFLECTIVE = 1,           /* boundaries.py:235-288 */
  SWK_BC_DIRICHLET = 2,            /* generic_boundary_conditions.py:221-264; also Time_boundary
                                      :370-411 and Time_stage_zero_momentum boundaries.py:616-635
                                      with host-evaluated values */
  SWK_BC_TRANSMISSIVE = 3,         /* generic_boundary_conditions.py:173-193 */
  SWK_BC_TRANSMISSIVE_N_ZERO_T_SET_STAGE = 4,   /* boundaries.py:477-517 */
  SWK_BC_TRANSMISSIVE_MOMENTUM_SET_STAGE = 5,   /* boundaries.py:344-372 */
  SWK_BC_TRANSMISSIVE_STAGE_ZERO_MOMENTUM = 6,  /* boundaries.py:543-551 */
  SWK_BC_FLATHER_EXTERNAL_STAGE_ZERO_VELOCITY = 7, /* boundaries.py:1096-1266 (evaluate_segment); v0 = external stage */
  SWK_BC_CHARACTERISTIC_STAGE = 8,                 /* boundaries.py:639-843 (evaluate_segment); v0 = outside stage */
  /* File_boundary / Time_space_boundary (generic_boundary_conditions.py:419-700) and Field_boundary
   * (boundaries.py:993-1090): values interpolated in time between two frames of a device-resident table
   * (swk_set_boundary_table); v0 = ratio, v1 = frame index, v2 = mean_stage (added to the stage by kind 10) */
  SWK_BC_TIME_SPACE_TABLE = 9,
  SWK_BC_TIME_SPACE_TABLE_MEAN_STAGE = 10,
  SWK_BC_DIRICHLET_DISCHARGE = 11                  /* boundaries.py:845-890; v0 = stage0, v1 = wh0 (momentum along the inward normal) */
};

/* ---- evolve result ------------------------------------------------------------ */
typedef struct {
  double time;                     /* relative model time reached */
  double timestep;                 /* last timestep taken */
  double flux_timestep;            /* last CFL-limiting timestep from the flux kernel */
  double recorded_min_timestep, recorded_max_timestep;
  double boundary_flux_integral;   /* boundary_flux_integral_operator.py:44-62 */
  double fractional_step_volume_integral;   /* rate_operators.py:259 */
  double mass_error;               /* accumulated return of protect (sw_domain_openmp.c:1149) */
  double boundary_flux_sum[3];
  int64_t number_of_steps;         /* since the last yield */
  int64_t number_of_first_order_steps;
  int64_t total_steps;             /* since swk_create */
  int64_t negative_cells;          /* accumulated fix_negative_cells count */
  int64_t stop_reason;             /* 0 step budget exhausted, 1 yieldtime reached, 2 finaltime reached */
  int64_t kernel_launches;         /* library kernels launched by this call */
} swk_evolve_result;

/* ---- host-side set-up helper (no GPU needed) --------------------------------------------------
 * Neighbour structure of a triangle table, the job of the reference's native
 * neighbour_table.cpp (build_neighbour_structure, semantics of neighbour_mesh.py:234-294):
 * neighbours / neighbour_edges (N,3) filled with -1 where there is no neighbour,
 * number_of_boundaries (N,).  Returns SWK_ERR_ARG when two triangles own the same directed edge. */
int swk_build_neighbour_structure(int64_t number_of_triangles, int64_t number_of_nodes,
                                  const int64_t *triangles, int64_t *neighbours,
                                  int64_t *neighbour_edges, int64_t *number_of_boundaries);

/* Triangle geometry of a mesh, the job of General_mesh.__init__ (general_mesh.py:156-277): per
 * triangle the vertex coordinates (3N,2), area, outward unit normals of the three edges (N,6),
 * edge lengths (N,3), centroid (N,2), radius (min distance centroid -> edge midpoint, or the
 * inscribed-circle radius) and edge midpoints (3N,2), with the reference's operation order.
 * *first_degenerate receives the first triangle with area <= 0, or -1. */
int swk_mesh_geometry(int64_t number_of_triangles, int64_t number_of_nodes, const double *nodes,
                      const int64_t *triangles, int64_t use_inscribed_circle, double *vertex_coordinates,
                      double *areas, double *normals, double *edgelengths, double *centroid_coordinates,
                      double *radii, double *edge_midpoint_coordinates, int64_t *first_degenerate);

/* =========================== (1) RESIDENT LAYER ================================ */

/* Number of usable sm_100 devices (0 and SWK_ERR_CUDA semantics: *count = 0). */
int swk_device_count(int *count);
const char *swk_last_error(void);
int swk_abi_version(void);

int swk_create(const swk_mesh *mesh, const swk_params *params, int device, swk_domain **out);
int swk_destroy(swk_domain *d);
/* Re-read scalars (set_flow_algorithm after creation; SURVEY.md 8(b) "State to mirror"). */
int swk_set_params(swk_domain *d, const swk_params *params);

/* Host<->device transfer of one quantity in the caller's triangle order.
 * n must equal the quantity's element count (N, 3N or M).                       */
int swk_set_quantity(swk_domain *d, int quantity_id, const double *host, int64_t n);
int swk_get_quantity(swk_domain *d, int quantity_id, double *host, int64_t n);

/* Bind a boundary kind to a list of boundary indices (Generic_Domain.set_boundary
 * :937-1033 / tag_boundary_cells).  values[0..2]: Dirichlet constants or
 * values[0] = stage for the set-stage kinds.  segment ids are boundary indices m. */
int swk_set_boundary_segment(swk_domain *d, int segment, int kind, const int64_t *ids,
                             int64_t n_ids, const double values[3]);
int swk_set_boundary_values(swk_domain *d, int segment, const double values[3]);
/* Time-space table of a segment of kind SWK_BC_TIME_SPACE_TABLE[_MEAN_STAGE]: frames[n_frames][n_points][3],
 * point j = the j-th boundary index passed to swk_set_boundary_segment; resident in HBM.  The time
 * interpolation q = Q0 + ratio*(Q1 - Q0) (fit_interpolate/interpolate.py:1056-1092) runs in the boundary
 * kernel.  swk_set_boundary_table_frame rewrites one frame in place.                              */
int swk_set_boundary_table(swk_domain *d, int segment, int64_t n_frames, int64_t n_points, const double *frames);
int swk_set_boundary_table_frame(swk_domain *d, int segment, int64_t frame, const double *values);
/* The values one RK substep sees (0: start of the step, 1: second flux evaluation, 2: third): the
 * reference evaluates time-dependent boundaries at the substep's own time (generic_domain.py:2011,
 * 2093, 2132).  swk_set_boundary_values sets all three.  Values live in a device table; changing them
 * is one small asynchronous copy and leaves the captured step valid.                              */
int swk_set_boundary_values_substep(swk_domain *d, int segment, int substep, const double values[3]);

/* Rate_operator (operators/rate_operators.py:24-269): stage += factor*dt*rate on
 * `indices` (NULL = all).  rate_array (N,) optional per-centroid rates (NULL: scalar). */
int swk_add_rate_operator(swk_domain *d, double rate, double factor, const double *rate_array,
                          const int64_t *indices, int64_t n_indices, int *op_id);
int swk_set_rate(swk_domain *d, int op_id, double rate, double factor);
/* State-independent explicit forcing: Wind_stress.__call__ + assign_windfield_values (shallow_water/forcing.py:
 * 133-215, momentum) and General_forcing.__call__ with its descendants Rainfall and Inflow (:400-640, stage over a
 * region): explicit_update[k] += force[k] at every flux evaluation.  Arrays of n = N doubles in the caller's order;
 * a NULL array is zero, all NULL switches the term off.                                                  */
int swk_set_explicit_forcing(swk_domain *d, const double *stage_force, const double *xmom_force,
                             const double *ymom_force, int64_t n);
/* Mark a Rate_operator whose rate / factor are functions of time (rate_operators.py:276-289): its
 * scalars are read from a device table that swk_set_rate refreshes, without touching the captured step. */
int swk_set_rate_dynamic(swk_domain *d, int op_id, int dynamic);
/* Forget every Rate_operator registered so far (ids become invalid). */
int swk_clear_rate_operators(swk_domain *d);

/* Small index sets (inlets, structures, gauges): read / write the centroid records of `n` triangles
 * without moving whole arrays.  out: (n,4) row-major {stage, xmomentum, ymomentum, elevation};
 * in: (n,3) {stage, xmomentum, ymomentum}.  Used by host-side operators that do scalar hydraulics on a
 * handful of cells per step (structures/inlet.py:69-190, inlet_operator.py:78-157).              */
int swk_gather_centroids(swk_domain *d, const int64_t *ids, int64_t n, double *out);
int swk_scatter_centroids(swk_domain *d, const int64_t *ids, int64_t n, const double *in);
/* Bed elevation of `n` triangles (in: (n,)): what Set_elevation writes into elevation.centroid_values
 * (operators/set_elevation.py:116-150, discontinuous-elevation branch). */
int swk_scatter_bed(swk_domain *d, const int64_t *ids, int64_t n, const double *in);
/* fractional_step_volume_integral += volume (host-side operators account their own water) */
int swk_add_fractional_step_volume(swk_domain *d, double volume);
/* Registered cell sets: the triangles an inlet reads and writes at every timestep (structures/inlet.py:135-330).
 * Device ids and a page-locked staging buffer are kept with the handle, so that Inlet.fetch is one small kernel +
 * one copy and Inlet.commit is queued without a host round trip.  Rows keep the order of `ids`.            */
int swk_register_cells(swk_domain *d, const int64_t *ids, int64_t n, int *set_id);
int swk_gather_set(swk_domain *d, int set_id, double *out);        /* (n,4) stage, xmom, ymom, elevation */
int swk_scatter_set(swk_domain *d, int set_id, const double *in);  /* (n,3) stage, xmom, ymom; asynchronous */
/* update_ghosts queued in stream order, without waiting for it */
int swk_update_ghosts_async(swk_domain *d);

/* Single-process ghost copy (Generic_Domain.update_ghosts :2448-2469):
 * centroid values of full_ids are copied onto ghost_ids after each update.       */
int swk_set_local_ghost_copy(swk_domain *d, const int64_t *full_ids, const int64_t *ghost_ids,
                             int64_t n);

/* Time state (relative times, generic_domain.py:1764-1807). */
int swk_set_time(swk_domain *d, double relative_time);

/* Individual steps of the path, operating on resident data ---------------------- */
int swk_protect(swk_domain *d, double *mass_error);                 /* sw_domain_openmp.c:1096 */
int swk_extrapolate_second_order_edge_sw(swk_domain *d);            /* :1336 (incl. loop-1/3 effects) */
int swk_distribute_to_vertices_and_edges(swk_domain *d);            /* protect + extrapolate */
int swk_update_boundary(swk_domain *d);                             /* generic_domain.py:2288 */
int swk_compute_fluxes(swk_domain *d, int substep, double *flux_timestep); /* :456 */
/* compute_forcing_terms (Manning friction, friction.py:22-73 -> sw_domain_openmp.c:1954, 1988) has no
 * entry point of its own on resident data: the semi-implicit friction term is evaluated inside
 * the update kernel, from the same protected centroid values, instead of round-tripping
 * semi_implicit_update through HBM.  swk_update_conserved_quantities = forcing terms +
 * Quantity.update x3 (quantity.c:772) + fix_negative_cells (:2037).                             */
int swk_update_conserved_quantities(swk_domain *d, double timestep, int64_t *num_negative);
int swk_backup_conserved_quantities(swk_domain *d);                 /* quantity.c:735 */
int swk_saxpy_conserved_quantities(swk_domain *d, double a, double b, double divide_by); /* quantity.c:752 (+ /3, generic_domain.py:2167-2170) */
int swk_update_ghosts(swk_domain *d);
/* apply_fractional_steps (generic_domain.py:2312-2314) for a host-driven step of length
 * `timestep`: boundary_flux_integral_operator + every registered Rate_operator.      */
int swk_apply_fractional_steps(swk_domain *d, double timestep);
/* current clock scalars (time, integrals, counters) without stepping */
int swk_get_statistics(swk_domain *d, swk_evolve_result *result);

/* Device-resident time loop.  Runs whole timesteps until relative_yieldtime or
 * relative_finaltime is reached (finaltime < 0: none) or max_steps steps were
 * taken (max_steps <= 0: unlimited).  Implements _evolve_base's loop body:
 * evolve_one_*_step, apply_fractional_steps, update_ghosts, step counters.
 * At a yield the centroid arrays are protected/extrapolated exactly as
 * distribute_to_vertices_and_edges + update_boundary do (generic_domain.py:1884-1902). */
int swk_evolve(swk_domain *d, double relative_yieldtime, double relative_finaltime,
               int64_t max_steps, swk_evolve_result *result);
/* Host-paced time loop for boundary values / rates that are host functions of time
 * (shallow_water/boundaries.py:384-517, 553-635; generic_boundary_conditions.py:297-411): same kernels,
 * same graphs, but the step is launched in two halves with one host visit in between -
 *     swk_step_begin(yieldtime, finaltime)
 *     loop:  swk_step_first(&r)    first half: extrapolate, boundary, flux, global dt, update_timestep;
 *                                  returns synchronised with r.time = start of the step, r.timestep = dt,
 *                                  r.stop_reason != 0 when the previous step reached the yield / final time
 *                                  (then nothing was computed: leave the loop)
 *            ... host evaluates f(t + dt) [, f(t + dt/2)], f(next step's t) and calls
 *                swk_set_boundary_values_substep / swk_set_rate ...
 *            swk_step_rest()       second half: update, later RK substeps, fractional steps, finish (async)
 *     swk_step_end(&r)             the yield's extrapolation + boundary update, final statistics        */
int swk_step_begin(swk_domain *d, double relative_yieldtime, double relative_finaltime);
int swk_step_first(swk_domain *d, swk_evolve_result *result);
int swk_step_rest(swk_domain *d);
int swk_step_end(swk_domain *d, swk_evolve_result *result);
/* reset per-yield statistics (generic_domain.py:1906-1912) */
int swk_reset_yield_statistics(swk_domain *d);

/* Launch exactly n_steps timesteps back to back (no host interaction inside) and time them
 * with CUDA events recorded on the library's stream; *elapsed_ms covers the whole region.
 * The clock keeps running (yield/final times are pushed out of the way).  With
 * per_kernel != 0 every launch of the four hot kernels is additionally bracketed by its
 * own event pair; read the totals with swk_kernel_timing.                              */
int swk_run_steps(swk_domain *d, int64_t n_steps, int per_kernel, float *elapsed_ms);
/* kernel ids: 0 extrapolate (pass A), 1 flux (B1), 2 update (B2), 3 flux_update (fused B).
 * total_ms[4], launches[4] accumulated by the last swk_run_steps(per_kernel=1).          */
int swk_kernel_timing(swk_domain *d, double total_ms[4], int64_t launches[4]);
int swk_stream(swk_domain *d, void **cuda_stream_out);
int swk_synchronize(swk_domain *d);
/* Page-lock / unlock a host buffer that is used repeatedly with swk_set/get_quantity (the numpy
 * arrays of Domain.quantities live as long as the domain), so that the copies run at PCIe speed. */
int swk_pin_host_buffer(swk_domain *d, void *host, size_t bytes);
int swk_unpin_host_buffer(swk_domain *d, void *host);
int swk_kernel_launch_count(swk_domain *d, int64_t *count);
/* Algorithmic HBM bytes per triangle and timestep of the current configuration
 *  and the bytes this library's layout actually moves.       */
int swk_bytes_per_triangle_step(swk_domain *d, double *algorithmic, double *layout);

/* ---- multi-GPU (one process per GPU) -----------------------
 * The caller creates the NCCL unique id on rank 0 (swk_nccl_unique_id), ships the
 * 128 bytes to every rank by its own means (torch.distributed, MPI, file), then
 * each rank calls swk_comm_init.  Halo lists are the reference's
 * full_send_dict[p][0] / ghost_recv_dict[p][0] (distribute_mesh.py:1128-1170).    */
int swk_nccl_unique_id(void *id128);
int swk_comm_init(swk_domain *d, const void *id128, int rank, int nranks);
/* Process-level communicator (one per GPU process), independent of any domain: the device time
 * loop of every domain attached to it (swk_comm_attach) and the small host-level collectives of
 * the host layer (swk_comm_allreduce: barriers, maxima of timings, the bit-exact integer merges
 * of the structure operators) share one ncclComm_t, so multi-GPU runs need nothing but NCCL.
 * Replaces the pypar/mpi4py calls of parallel/parallel_api.py and
 * parallel_generic_communications.py:35-67.                                              */
typedef struct swk_comm swk_comm;
#define SWK_F64 0
#define SWK_I64 1
#define SWK_SUM 0
#define SWK_MIN 1
#define SWK_MAX 2
int swk_comm_create(const void *id128, int rank, int nranks, int device, swk_comm **out);
int swk_comm_destroy(swk_comm *comm);
int swk_comm_allreduce(swk_comm *comm, void *host_buf, int64_t n, int dtype, int op);
int swk_comm_attach(swk_domain *d, swk_comm *comm);
int swk_set_halo(swk_domain *d, int n_peers, const int *peer_ranks,
                 const int64_t *send_counts, const int64_t *const *send_ids,
                 const int64_t *recv_counts, const int64_t *const *recv_ids);

/* ============================ (2) PER-CALL LAYER =================================
 * Host view of one reference Domain: the pointers that
 * get_python_domain_pointers (sw_domain_openmp_ext.pyx:138-365) extracts, by the
 * same names.  Arrays are read and updated IN PLACE like the reference's C code
 * does, so user code holding numpy aliases keeps seeing the results.           */
typedef struct {
  swk_mesh mesh;
  swk_params params;
  double *stage_centroid_values, *xmom_centroid_values, *ymom_centroid_values;
  double *bed_centroid_values, *height_centroid_values, *friction_centroid_values;
  double *stage_edge_values, *xmom_edge_values, *ymom_edge_values;
  double *bed_edge_values, *height_edge_values;
  double *stage_vertex_values, *xmom_vertex_values, *ymom_vertex_values;
  double *bed_vertex_values, *height_vertex_values;
  double *stage_boundary_values, *xmom_boundary_values, *ymom_boundary_values;
  double *stage_explicit_update, *xmom_explicit_update, *ymom_explicit_update;
  double *stage_semi_implicit_update, *xmom_semi_implicit_update, *ymom_semi_implicit_update;
  double *max_speed;
  double *boundary_flux_sum;          /* (timestep_fluxcalls,) */
} swk_host_view;

/* A per-call context caches the device copy of the static mesh between calls. */
int swk_call_open(const swk_host_view *v, int device, swk_domain **out);
int swk_call_close(swk_domain *d);

/* replaces compute_fluxes_ext_central(domain, timestep) -> float
 * (sw_domain_openmp_ext.pyx:371 -> sw_domain_openmp.c:456).  `substep` replaces the
 * function-static call counter (:492-505).  Reads edge + boundary values, stage/bed
 * centroids; writes explicit_update x3, max_speed, boundary_flux_sum[substep].    */
int swk_call_compute_fluxes_ext_central(swk_domain *d, const swk_host_view *v, double timestep,
                                        int substep, double *flux_timestep);
/* replaces extrapolate_second_order_edge_sw(domain) (pyx:384 -> .c:1336) */
int swk_call_extrapolate_second_order_edge_sw(swk_domain *d, const swk_host_view *v);
/* replaces protect_new(domain) -> mass_error (pyx:397 -> .c:1096) */
int swk_call_protect_new(swk_domain *d, const swk_host_view *v, double *mass_error);
/* replaces fix_negative_cells(domain) -> count (pyx:449 -> .c:2037) */
int swk_call_fix_negative_cells(swk_domain *d, const swk_host_view *v, int64_t *count);
/* replaces manning_friction_flat / _sloped (pyx:417-446 -> .c:1954, 1988) */
int swk_call_manning_friction_flat(int device, double g, double eps, int64_t N, const double *w,
                                   const double *zv, const double *uh, const double *vh,
                                   const double *eta, double *xmom_update, double *ymom_update);
int swk_call_manning_friction_sloped(int device, double g, double eps, int64_t N, const double *x,
                                     const double *w, const double *zv, const double *uh,
                                     const double *vh, const double *eta,
                                     double *xmom_update, double *ymom_update);
/* replaces quantity_ext.update / backup_centroid_values / saxpy_centroid_values
 * (quantity_ext.pyx:38-117 -> quantity.c:735-820)                                */
int swk_call_update(int device, int64_t N, double timestep, double *centroid_values,
                    const double *explicit_update, double *semi_implicit_update);
int swk_call_backup_centroid_values(int device, int64_t N, const double *centroid_values,
                                    double *centroid_backup_values);
int swk_call_saxpy_centroid_values(int device, int64_t N, double a, double b,
                                   double *centroid_values, const double *centroid_backup_values);

#ifdef __cplusplus
}
#endif
#endif /* SWK_H */
