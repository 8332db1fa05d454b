/* TEST INFRASTRUCTURE ONLY - never linked or loaded by the product.
 *
 * Plain-pointer view of one shallow-water domain, shared by
 *   - oracle/sw_oracle.c : CPU restatement ("port") of the reference's DE kernels
 *   - oracle/ref_shim.c  : adapter that feeds the same view to the reference's own
 *                          C sources compiled from /root/reference (oracle/_ref/)
 * The field set mirrors what the reference's `struct domain` carries for this
 * path (anuga/shallow_water/sw_domain.h:14-113), with the same array layouts
 * (C-contiguous, int64 indices, FP64 values).
 */
#ifndef ORC_DOMAIN_H
#define ORC_DOMAIN_H
#include <stdint.h>

typedef struct {
  int64_t number_of_elements;
  int64_t boundary_length;
  int64_t number_of_riverwall_edges;
  int64_t ncol_riverwall_hydraulic_properties;
  int64_t extrapolate_velocity_second_order;
  int64_t low_froude;
  int64_t timestep_fluxcalls;
  int64_t optimise_dry_cells;
  double epsilon, H0, g, minimum_allowed_height, maximum_allowed_speed;
  double evolve_max_timestep;
  double beta_w, beta_w_dry, beta_uh, beta_uh_dry, beta_vh, beta_vh_dry;

  /* mesh (static) */
  int64_t *neighbours;           /* (N,3) <0 : boundary index -(m+1) */
  int64_t *neighbour_edges;      /* (N,3) */
  int64_t *surrogate_neighbours; /* (N,3) */
  int64_t *number_of_boundaries; /* (N,)  */
  int64_t *tri_full_flag;        /* (N,)  */
  int64_t *edge_flux_type;       /* (3N,) */
  int64_t *edge_river_wall_counter; /* (3N,) */
  double *normals;               /* (N,6) */
  double *edgelengths;           /* (N,3) */
  double *radii;                 /* (N,)  */
  double *areas;                 /* (N,)  */
  double *centroid_coordinates;  /* (N,2) */
  double *edge_coordinates;      /* (3N,2) */
  double *vertex_coordinates;    /* (3N,2) */
  double *riverwall_elevation;
  int64_t *riverwall_rowIndex;
  double *riverwall_hydraulic_properties;

  /* quantities */
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
  double *x_centroid_work, *y_centroid_work;
  double *boundary_flux_sum;     /* (timestep_fluxcalls,) */
} orc_domain;

#endif
