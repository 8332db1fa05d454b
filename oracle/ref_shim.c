/* TEST INFRASTRUCTURE ONLY - never linked, loaded or called by the product.
 *
 * Adapter between the plain-pointer `orc_domain` view and the reference's own
 * C sources.  oracle/Makefile compiles
 *     /root/reference/anuga/shallow_water/sw_domain_openmp.c      (mode 2 kernels)
 *     /root/reference/anuga/abstract_2d_finite_volumes/quantity.c (update/backup/saxpy)
 * where they lie (no copies) together with this file into
 * oracle/_ref/libanuga_ref.so.  The only declarations taken from the reference
 * are its public header `sw_domain.h` (struct domain) and the prototypes below.
 */
#include <stdint.h>
#include <string.h>
#include "sw_domain.h"      /* -I/root/reference/anuga/shallow_water */
#include "orc_domain.h"

/* prototypes of the reference entry points (sw_domain_openmp.c, quantity.c) */
double _openmp_compute_fluxes_central(struct domain *D, double timestep);
double _openmp_protect(struct domain *D);
int64_t _openmp_extrapolate_second_order_edge_sw(struct domain *D);
int64_t _openmp_fix_negative_cells(struct domain *D);
void _openmp_manning_friction_flat(double g, double eps, int64_t N, double *w, double *zv,
                                   double *uh, double *vh, double *eta,
                                   double *xmom_update, double *ymom_update);
void _openmp_manning_friction_sloped(double g, double eps, int64_t N, double *x, double *w,
                                     double *zv, double *uh, double *vh, double *eta,
                                     double *xmom_update, double *ymom_update);
int64_t _update(int64_t N, double timestep, double *centroid_values,
                double *explicit_update, double *semi_implicit_update);
int64_t _backup_centroid_values(int64_t N, double *centroid_values, double *backup);
int64_t _saxpy_centroid_values(int64_t N, double a, double b, double *centroid_values,
                               double *backup);

static void fill(struct domain *R, const orc_domain *D)
{
  memset(R, 0, sizeof(*R));
  R->number_of_elements = D->number_of_elements;
  R->boundary_length = D->boundary_length;
  R->number_of_riverwall_edges = D->number_of_riverwall_edges;
  R->epsilon = D->epsilon;
  R->H0 = D->H0;
  R->g = D->g;
  R->optimise_dry_cells = D->optimise_dry_cells;
  R->evolve_max_timestep = D->evolve_max_timestep;
  R->extrapolate_velocity_second_order = D->extrapolate_velocity_second_order;
  R->minimum_allowed_height = D->minimum_allowed_height;
  R->maximum_allowed_speed = D->maximum_allowed_speed;
  R->low_froude = D->low_froude;
  R->timestep_fluxcalls = D->timestep_fluxcalls;
  R->beta_w = D->beta_w;
  R->beta_w_dry = D->beta_w_dry;
  R->beta_uh = D->beta_uh;
  R->beta_uh_dry = D->beta_uh_dry;
  R->beta_vh = D->beta_vh;
  R->beta_vh_dry = D->beta_vh_dry;
  R->max_flux_update_frequency = 1;
  R->ncol_riverwall_hydraulic_properties = D->ncol_riverwall_hydraulic_properties;
  R->neighbours = D->neighbours;
  R->neighbour_edges = D->neighbour_edges;
  R->surrogate_neighbours = D->surrogate_neighbours;
  R->normals = D->normals;
  R->edgelengths = D->edgelengths;
  R->radii = D->radii;
  R->areas = D->areas;
  R->edge_flux_type = D->edge_flux_type;
  R->tri_full_flag = D->tri_full_flag;
  R->max_speed = D->max_speed;
  R->vertex_coordinates = D->vertex_coordinates;
  R->edge_coordinates = D->edge_coordinates;
  R->centroid_coordinates = D->centroid_coordinates;
  R->number_of_boundaries = D->number_of_boundaries;
  R->stage_edge_values = D->stage_edge_values;
  R->xmom_edge_values = D->xmom_edge_values;
  R->ymom_edge_values = D->ymom_edge_values;
  R->bed_edge_values = D->bed_edge_values;
  R->height_edge_values = D->height_edge_values;
  R->stage_centroid_values = D->stage_centroid_values;
  R->xmom_centroid_values = D->xmom_centroid_values;
  R->ymom_centroid_values = D->ymom_centroid_values;
  R->bed_centroid_values = D->bed_centroid_values;
  R->height_centroid_values = D->height_centroid_values;
  R->stage_vertex_values = D->stage_vertex_values;
  R->xmom_vertex_values = D->xmom_vertex_values;
  R->ymom_vertex_values = D->ymom_vertex_values;
  R->bed_vertex_values = D->bed_vertex_values;
  R->height_vertex_values = D->height_vertex_values;
  R->stage_boundary_values = D->stage_boundary_values;
  R->xmom_boundary_values = D->xmom_boundary_values;
  R->ymom_boundary_values = D->ymom_boundary_values;
  R->stage_explicit_update = D->stage_explicit_update;
  R->xmom_explicit_update = D->xmom_explicit_update;
  R->ymom_explicit_update = D->ymom_explicit_update;
  R->x_centroid_work = D->x_centroid_work;
  R->y_centroid_work = D->y_centroid_work;
  R->boundary_flux_sum = D->boundary_flux_sum;
  R->edge_river_wall_counter = D->edge_river_wall_counter;
  R->riverwall_elevation = D->riverwall_elevation;
  R->riverwall_rowIndex = D->riverwall_rowIndex;
  R->riverwall_hydraulic_properties = D->riverwall_hydraulic_properties;
  R->stage_semi_implicit_update = D->stage_semi_implicit_update;
  R->xmom_semi_implicit_update = D->xmom_semi_implicit_update;
  R->ymom_semi_implicit_update = D->ymom_semi_implicit_update;
}

/* The reference derives the RK substep from function-static call counters
 * (sw_domain_openmp.c:492-505): base_call is reset to the current call whenever
 * D->timestep_fluxcalls differs from the remembered value.  To evaluate a given
 * `substep` we re-base with an empty (N=0) domain carrying a different
 * timestep_fluxcalls, then burn `substep` empty calls.
 */
static int64_t remembered_fluxcalls = 1;   /* mirrors the reference's static */

static void empty_call(int64_t fluxcalls)
{
  struct domain E;
  double bfs[64];
  memset(&E, 0, sizeof(E));
  E.number_of_elements = 0;
  E.timestep_fluxcalls = fluxcalls;
  E.boundary_flux_sum = bfs;
  _openmp_compute_fluxes_central(&E, 0.0);
  remembered_fluxcalls = fluxcalls;
}

double ref_compute_fluxes(orc_domain *D, double timestep, int64_t substep)
{
  struct domain R;
  fill(&R, D);
  const int64_t T = D->timestep_fluxcalls;
  /* force a re-base on the next call that carries T */
  empty_call(T + 7);
  if (substep > 0) {
    empty_call(T);                       /* re-based: this call is substep 0 */
    for (int64_t s = 1; s < substep; s++) empty_call(T);
  }
  remembered_fluxcalls = T;
  return _openmp_compute_fluxes_central(&R, timestep);
}

double ref_protect(orc_domain *D)
{
  struct domain R;
  fill(&R, D);
  return _openmp_protect(&R);
}

int64_t ref_extrapolate(orc_domain *D)
{
  struct domain R;
  fill(&R, D);
  return _openmp_extrapolate_second_order_edge_sw(&R);
}

int64_t ref_fix_negative_cells(orc_domain *D)
{
  struct domain R;
  fill(&R, D);
  return _openmp_fix_negative_cells(&R);
}

void ref_manning_friction_flat(double g, double eps, int64_t N, double *w, double *zv,
                               double *uh, double *vh, double *eta,
                               double *xmom_update, double *ymom_update)
{
  _openmp_manning_friction_flat(g, eps, N, w, zv, uh, vh, eta, xmom_update, ymom_update);
}

void ref_manning_friction_sloped(double g, double eps, int64_t N, double *x, double *w,
                                 double *zv, double *uh, double *vh, double *eta,
                                 double *xmom_update, double *ymom_update)
{
  _openmp_manning_friction_sloped(g, eps, N, x, w, zv, uh, vh, eta, xmom_update, ymom_update);
}

int64_t ref_update(int64_t N, double timestep, double *c, double *eu, double *siu)
{
  return _update(N, timestep, c, eu, siu);
}

void ref_backup_centroid_values(int64_t N, double *c, double *backup)
{
  _backup_centroid_values(N, c, backup);
}

void ref_saxpy_centroid_values(int64_t N, double a, double b, double *c, double *backup)
{
  _saxpy_centroid_values(N, a, b, c, backup);
}
