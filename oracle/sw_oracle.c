/* TEST INFRASTRUCTURE ONLY - never linked, loaded or called by the product.
 *
 * CPU restatement ("port") of the reference's discontinuous-elevation kernels,
 * serial, plain C, FP64.  Each function cites the reference lines it restates
 * (paths relative to /root/reference/anuga).  Build with
 *     gcc -O2 -ffp-contract=off -fPIC -shared   (see oracle/Makefile)
 * -ffp-contract=off keeps every multiply and add separately rounded, which is
 * the "parity build" of SURVEY.md section 7; the CUDA kernels are compiled with
 * -fmad=false for the same reason.
 *
 * Pinned against: the reference's own C sources compiled into oracle/_ref
 * (tests/test_oracle_vs_ref.py) and the golden fixtures generated from the
 * Python reference (tests/golden/, tests/test_oracle_golden.py).
 */
#include <math.h>
#include <stdint.h>
#include <string.h>
#include "orc_domain.h"

/* ---------------------------------------------------------------------------
 * Edge flux: Kurganov-Noelle-Petrova central-upwind with Audusse heights.
 * shallow_water/sw_domain_openmp.c:65-268
 * ------------------------------------------------------------------------- */
static void edge_flux_central(double wl, double uhl_xy, double vhl_xy,
                              double wr, double uhr_xy, double vhr_xy,
                              double h_left, double h_right,
                              double hle, double hre,
                              double n1, double n2,
                              double epsilon, double ze, double g,
                              int64_t low_froude,
                              double flux[3], double *max_speed,
                              double *pressure_flux)
{
  if (h_left == 0. && h_right == 0.) {             /* :101-108 */
    flux[0] = flux[1] = flux[2] = 0.0;
    *max_speed = 0.0;
    *pressure_flux = 0.0;
    return;
  }
  /* rotate momenta into the edge frame (:41-62, :118-119) */
  double uh_left = n1 * uhl_xy + n2 * vhl_xy;
  double vh_left = -n2 * uhl_xy + n1 * vhl_xy;
  double uh_right = n1 * uhr_xy + n2 * vhr_xy;
  double vh_right = -n2 * uhr_xy + n1 * vhr_xy;

  double u_left, v_left, u_right, v_right, inv;
  if (hle > 0.0) {                                  /* :125-139 */
    inv = 1.0 / hle;
    u_left = uh_left * inv;
    uh_left = h_left * u_left;
    v_left = vh_left * inv;
    vh_left = h_left * inv * vh_left;
  } else {
    u_left = 0.; uh_left = 0.; vh_left = 0.; v_left = 0.;
  }
  if (hre > 0.0) {                                  /* :147-161 */
    inv = 1.0 / hre;
    u_right = uh_right * inv;
    uh_right = h_right * u_right;
    v_right = vh_right * inv;
    vh_right = h_right * inv * vh_right;
  } else {
    u_right = 0.; uh_right = 0.; vh_right = 0.; v_right = 0.;
  }

  double c_left = sqrt(g * h_left);                 /* :166-167 */
  double c_right = sqrt(g * h_right);

  double local_fr;                                  /* :178-194 */
  if (low_froude == 1) {
    local_fr = sqrt(fmax(0.001, fmin(1.0,
        (u_right * u_right + u_left * u_left + v_right * v_right + v_left * v_left) /
        (c_left * c_left + c_right * c_right + 1.0e-10))));
  } else if (low_froude == 2) {
    local_fr = sqrt((u_right * u_right + u_left * u_left + v_right * v_right + v_left * v_left) /
                    (c_left * c_left + c_right * c_right + 1.0e-10));
    local_fr = sqrt(fmin(1.0, 0.01 + fmax(local_fr - 0.01, 0.0)));
  } else {
    local_fr = 1.0;
  }

  double s_max = fmax(u_left + c_left, u_right + c_right);   /* :197-211 */
  if (s_max < 0.0) s_max = 0.0;
  double s_min = fmin(u_left - c_left, u_right - c_right);
  if (s_min > 0.0) s_min = 0.0;

  double fl0 = u_left * h_left, fl1 = u_left * uh_left, fl2 = u_left * vh_left;     /* :218-224 */
  double fr0 = u_right * h_right, fr1 = u_right * uh_right, fr2 = u_right * vh_right;

  double denom = s_max - s_min;                     /* :227 */
  if (denom < epsilon) {                            /* :228-236 */
    flux[0] = flux[1] = flux[2] = 0.0;
    *max_speed = 0.0;
    *pressure_flux = 0.5 * g * 0.5 * (h_left * h_left + h_right * h_right);
    return;
  }
  *max_speed = fmax(s_max, -s_min);                 /* :240 */
  double inverse_denominator = 1.0 / fmax(denom, 1.0e-100);
  double e0 = s_max * fl0 - s_min * fr0;            /* :245-258 */
  e0 += (s_max * s_min) * (fmax(wr, ze) - fmax(wl, ze));
  e0 *= inverse_denominator;
  double e1 = s_max * fl1 - s_min * fr1;
  e1 += local_fr * (s_max * s_min) * (uh_right - uh_left);
  e1 *= inverse_denominator;
  double e2 = s_max * fl2 - s_min * fr2;
  e2 += local_fr * (s_max * s_min) * (vh_right - vh_left);
  e2 *= inverse_denominator;
  *pressure_flux = 0.5 * g * (s_max * h_left * h_left - s_min * h_right * h_right) * inverse_denominator; /* :261 */
  /* rotate back, i.e. rotate with (n1, -n2)  (:264) */
  double mn2 = -n2;
  flux[0] = e0;
  flux[1] = n1 * e1 + mn2 * e2;
  flux[2] = -mn2 * e1 + n1 * e2;
}

/* Villemonte weir blend for riverwall edges; sw_domain_openmp.c:324-426 */
static void weir_adjust(double flux[3], double h_left, double h_right, double g,
                        double weir_height, double Qfactor, double s1, double s2,
                        double h1, double h2, double *max_speed_local)
{
  const double twothirds = (2.0 / 3.0);
  if ((h_left <= 0.0) && (h_right <= 0.0)) return;
  double minhd = fmin(h_left, h_right);
  double maxhd = fmax(h_left, h_right);
  double rw = Qfactor * twothirds * maxhd * sqrt(twothirds * g * maxhd);
  double rw2 = Qfactor * twothirds * minhd * sqrt(twothirds * g * minhd);
  double rwRat = rw2 / fmax(rw, 1.0e-100);
  double hdRat = minhd / fmax(maxhd, 1.0e-100);
  double hdWrRat = minhd / fmax(weir_height, 1.0e-100);
  rw = rw * pow(1.0 - rwRat, 0.385);
  if (h_right > h_left) rw *= -1.0;
  if ((hdRat < s2) & (hdWrRat < h2)) {
    double w1 = fmin(fmax(hdRat - s1, 0.) / (s2 - s1), 1.0);
    double w2 = fmin(fmax(hdWrRat - h1, 0.) / (h2 - h1), 1.0);
    double newFlux = (rw * (1.0 - w1) + w1 * flux[0]) * (1.0 - w2) + w2 * flux[0];
    double scaleFlux;
    if (fabs(flux[0]) > 1.0e-100) scaleFlux = newFlux / flux[0];
    else scaleFlux = 0.;
    scaleFlux = fmax(scaleFlux, 0.);
    flux[0] = newFlux;
    flux[1] *= fmin(scaleFlux, 10.);
    flux[2] *= fmin(scaleFlux, 10.);
  }
  if (fabs(flux[0]) > 0.)
    *max_speed_local = sqrt(g * (maxhd + weir_height)) + fabs(flux[0] / (maxhd + 1.0e-12));
}

/* ---------------------------------------------------------------------------
 * compute_fluxes: sw_domain_openmp.c:456-772.
 * The reference derives `substep` from function-static call counters
 * (:492-505); here it is an explicit argument.
 * Returns the CFL-limiting local timestep on substep 0, `timestep` otherwise.
 * ------------------------------------------------------------------------- */
double orc_compute_fluxes(orc_domain *D, double timestep, int64_t substep)
{
  const int64_t K = D->number_of_elements;
  const double g = D->g, epsilon = D->epsilon;
  const int64_t ncol = D->ncol_riverwall_hydraulic_properties;
  double local_timestep = 1.0e+100;
  double bflux = 0.0;

  for (int64_t k = 0; k < K; k++) {
    double speed_max_last = 0.0;
    double su = 0.0, xu = 0.0, yu = 0.0;
    const double hc = D->height_centroid_values[k];
    const double zc = D->bed_centroid_values[k];
    for (int i = 0; i < 3; i++) {
      const int64_t ki = 3 * k + i;
      double wl = D->stage_edge_values[ki];
      double uhl = D->xmom_edge_values[ki];
      double vhl = D->ymom_edge_values[ki];
      double zl = D->bed_edge_values[ki];
      double hle = D->height_edge_values[ki];
      double wr, uhr, vhr, zr, hre, hc_n = hc, zc_n = zc;
      const int64_t n = D->neighbours[ki];
      if (n < 0) {                                   /* :551-561 */
        const int64_t m = -n - 1;
        wr = D->stage_boundary_values[m];
        uhr = D->xmom_boundary_values[m];
        vhr = D->ymom_boundary_values[m];
        zr = zl;
        hre = fmax(wr - zr, 0.0);
      } else {                                       /* :562-576 */
        hc_n = D->height_centroid_values[n];
        zc_n = D->bed_centroid_values[n];
        const int64_t nm = n * 3 + D->neighbour_edges[ki];
        wr = D->stage_edge_values[nm];
        uhr = D->xmom_edge_values[nm];
        vhr = D->ymom_edge_values[nm];
        zr = D->bed_edge_values[nm];
        hre = D->height_edge_values[nm];
      }
      double z_half = fmax(zl, zr);                  /* :579 */
      const int riverwall = D->edge_flux_type && D->edge_flux_type[ki] == 1;
      int64_t rwc = 0;
      if (riverwall) {                               /* :582-588 */
        rwc = D->edge_river_wall_counter[ki];
        z_half = fmax(D->riverwall_elevation[rwc - 1], z_half);
      }
      double h_left = fmax(hle + zl - z_half, 0.);   /* :591-592 */
      double h_right = fmax(hre + zr - z_half, 0.);
      const double n1 = D->normals[2 * ki], n2 = D->normals[2 * ki + 1];
      double flux[3], max_speed_local, pressure_flux;
      edge_flux_central(wl, uhl, vhl, wr, uhr, vhr, h_left, h_right, hle, hre,
                        n1, n2, epsilon, z_half, g, D->low_froude,
                        flux, &max_speed_local, &pressure_flux);
      if (riverwall) {                               /* :607-653 */
        const int64_t ii = D->riverwall_rowIndex[rwc - 1] * ncol;
        const double Qfactor = D->riverwall_hydraulic_properties[ii];
        const double s1 = D->riverwall_hydraulic_properties[ii + 1];
        const double s2 = D->riverwall_hydraulic_properties[ii + 2];
        const double h1 = D->riverwall_hydraulic_properties[ii + 3];
        const double h2 = D->riverwall_hydraulic_properties[ii + 4];
        const double weir_height = fmax(D->riverwall_elevation[rwc - 1] - fmin(zl, zr), 0.);
        const double h_left_tmp = fmax(D->stage_centroid_values[k] - z_half, 0.);
        double h_right_tmp;
        if (n >= 0) h_right_tmp = fmax(D->stage_centroid_values[n] - z_half, 0.);
        else h_right_tmp = fmax(hc_n + zr - z_half, 0.);
        if (D->riverwall_elevation[rwc - 1] > fmax(zc, zc_n))
          weir_adjust(flux, h_left_tmp, h_right_tmp, g, weir_height, Qfactor,
                      s1, s2, h1, h2, &max_speed_local);
      }
      const double length = D->edgelengths[ki];      /* :656-662 */
      flux[0] = -flux[0] * length;
      flux[1] = -flux[1] * length;
      flux[2] = -flux[2] * length;
      const double pressuregrad =
          length * (-g * 0.5 * (h_left * h_left - hle * hle - (hle + hc) * (zl - zc)) + pressure_flux);

      if (substep == 0) {                            /* :667-686 */
        const double edge_timestep = D->radii[k] * 1.0 / fmax(max_speed_local, epsilon);
        if (D->tri_full_flag[k] == 1) {
          if (max_speed_local > epsilon) {
            local_timestep = fmin(local_timestep, edge_timestep);
            speed_max_last = fmax(speed_max_last, max_speed_local);
          }
        }
      }
      su += flux[0];                                 /* :689-704 */
      xu += flux[1];
      yu += flux[2];
      if (((n < 0) & (D->tri_full_flag[k] == 1)) |
          ((n >= 0) && ((D->tri_full_flag[k] == 1) & (D->tri_full_flag[n] == 0))))
        bflux += flux[0];
      xu -= n1 * pressuregrad;
      yu -= n2 * pressuregrad;
    }
    if (substep == 0) D->max_speed[k] = speed_max_last;   /* :709-710 */
    const double inv_area = 1.0 / D->areas[k];       /* :714-717 */
    D->stage_explicit_update[k] = su * inv_area;
    D->xmom_explicit_update[k] = xu * inv_area;
    D->ymom_explicit_update[k] = yu * inv_area;
  }
  D->boundary_flux_sum[substep] = bflux;             /* :765 */
  if (substep == 0) timestep = local_timestep;       /* :768-771 */
  return timestep;
}

/* ---------------------------------------------------------------------------
 * protect: sw_domain_openmp.c:1096-1171 (xmom is zeroed twice, ymom never: kept)
 * ------------------------------------------------------------------------- */
double orc_protect(orc_domain *D)
{
  const int64_t K = D->number_of_elements;
  const double mah = D->minimum_allowed_height;
  double mass_error = 0.;
  for (int64_t k = 0; k < K; k++) {
    const double hc = D->stage_centroid_values[k] - D->bed_centroid_values[k];
    if (hc < mah * 1.0) {
      D->xmom_centroid_values[k] = 0.;
      if (hc <= 0.0) {
        const double bmin = D->bed_centroid_values[k];
        if (D->stage_centroid_values[k] < bmin) {
          mass_error += (bmin - D->stage_centroid_values[k]) * D->areas[k];
          D->stage_centroid_values[k] = bmin;
          D->stage_vertex_values[3 * k] = bmin;
          D->stage_vertex_values[3 * k + 1] = bmin;
          D->stage_vertex_values[3 * k + 2] = bmin;
        }
      }
    }
  }
  return mass_error;
}

/* limiter: sw_domain_openmp.c:1174-1231.  r0 is carried from edge to edge. */
static void limit_gradient(double dqv[3], double qmin, double qmax, double beta)
{
  double r = 1000.0, r0 = 1.0;
  const double TINY = 1.0e-100;
  for (int i = 0; i < 3; i++) {
    if (dqv[i] < -TINY) r0 = qmin / dqv[i];
    if (dqv[i] > TINY) r0 = qmax / dqv[i];
    r = fmin(r0, r);
  }
  const double phi = fmin(r * beta, 1.0);
  dqv[0] = dqv[0] * phi;
  dqv[1] = dqv[1] * phi;
  dqv[2] = dqv[2] * phi;
}

/* three-neighbour plane gradient + limiter: sw_domain_openmp.c:1233-1284 */
static void edge_values_3(double beta, double qc, double q0, double q1, double q2,
                          const double dxv[3], const double dyv[3],
                          double dx1, double dx2, double dy1, double dy2,
                          double inv_area2, double out[3])
{
  if (beta > 0.) {
    const double dq0 = q0 - qc;
    const double dq1 = q1 - q0;
    const double dq2 = q2 - q0;
    double a = dy2 * dq1 - dy1 * dq2;
    a *= inv_area2;
    double b = dx1 * dq2 - dx2 * dq1;
    b *= inv_area2;
    double dqv[3];
    dqv[0] = a * dxv[0] + b * dyv[0];
    dqv[1] = a * dxv[1] + b * dyv[1];
    dqv[2] = a * dxv[2] + b * dyv[2];
    const double qmax = fmax(fmax(dq0, fmax(dq0 + dq1, dq0 + dq2)), 0.0);
    const double qmin = fmin(fmin(dq0, fmin(dq0 + dq1, dq0 + dq2)), 0.0);
    limit_gradient(dqv, qmin, qmax, beta);
    out[0] = qc + dqv[0];
    out[1] = qc + dqv[1];
    out[2] = qc + dqv[2];
  } else {
    out[0] = out[1] = out[2] = qc;
  }
}

/* single-neighbour gradient (two boundary edges): sw_domain_openmp.c:1702-1840 */
static void edge_values_1(double beta, double qc, double q1,
                          const double dxv[3], const double dyv[3],
                          double dx2, double dy2, double out[3])
{
  const double dq1 = q1 - qc;
  const double a = dq1 * dx2;
  const double b = dq1 * dy2;
  double dqv[3], qmin, qmax;
  dqv[0] = a * dxv[0] + b * dyv[0];
  dqv[1] = a * dxv[1] + b * dyv[1];
  dqv[2] = a * dxv[2] + b * dyv[2];
  if (dq1 >= 0.0) { qmin = 0.0; qmax = dq1; }
  else { qmin = dq1; qmax = 0.0; }
  limit_gradient(dqv, qmin, qmax, beta);
  out[0] = qc + dqv[0];
  out[1] = qc + dqv[1];
  out[2] = qc + dqv[2];
}

/* ---------------------------------------------------------------------------
 * extrapolate_second_order_edge_sw: sw_domain_openmp.c:1336-1952
 * Serial in-place semantics (identical to the reference's modes 1 and 2 run
 * with one thread).
 * ------------------------------------------------------------------------- */
int64_t orc_extrapolate(orc_domain *D)
{
  const int64_t K = D->number_of_elements;
  const double mah = D->minimum_allowed_height;
  const int64_t vel2 = D->extrapolate_velocity_second_order;
  const double a_tmp = 0.3, b_tmp = 0.1;
  const double c_tmp = 1.0 / (a_tmp - b_tmp);
  const double d_tmp = 1.0 - (c_tmp * a_tmp);
  double *wc = D->stage_centroid_values, *zc = D->bed_centroid_values;
  double *uc = D->xmom_centroid_values, *vc = D->ymom_centroid_values;
  double *hcv = D->height_centroid_values;
  double *xw = D->x_centroid_work, *yw = D->y_centroid_work;

  for (int64_t k = 0; k < K; k++) {                  /* loop 1 :1373-1402 */
    const double dk = fmax(wc[k] - zc[k], 0.0);
    hcv[k] = dk;
    xw[k] = 0.0;
    yw[k] = 0.0;
    if (dk <= mah) {
      uc[k] = 0.0;
      vc[k] = 0.0;
    }
    if (vel2 == 1 && dk > mah) {
      const double inv = 1.0 / dk;
      xw[k] = uc[k];
      uc[k] = uc[k] * inv;
      yw[k] = vc[k];
      vc[k] = vc[k] * inv;
    }
  }

  for (int64_t k = 0; k < K; k++) {                  /* loop 2 :1408-1896 */
    const int64_t k3 = 3 * k;
    const double x = D->centroid_coordinates[2 * k], y = D->centroid_coordinates[2 * k + 1];
    double dxv[3], dyv[3];
    for (int i = 0; i < 3; i++) {
      dxv[i] = D->edge_coordinates[6 * k + 2 * i] - x;
      dyv[i] = D->edge_coordinates[6 * k + 2 * i + 1] - y;
    }
    const int64_t k0 = D->surrogate_neighbours[k3];
    const int64_t k1 = D->surrogate_neighbours[k3 + 1];
    const int64_t k2 = D->surrogate_neighbours[k3 + 2];
    const double x0 = D->centroid_coordinates[2 * k0], y0 = D->centroid_coordinates[2 * k0 + 1];
    const double x1 = D->centroid_coordinates[2 * k1], y1 = D->centroid_coordinates[2 * k1 + 1];
    const double x2 = D->centroid_coordinates[2 * k2], y2 = D->centroid_coordinates[2 * k2 + 1];
    double dx1 = x1 - x0, dx2 = x2 - x0, dy1 = y1 - y0, dy2 = y2 - y0;
    const double area2 = dy2 * dx1 - dy1 * dx2;

    if (((hcv[k0] < mah) | (k0 == k)) & ((hcv[k1] < mah) | (k1 == k)) &
        ((hcv[k2] < mah) | (k2 == k))) {             /* :1486-1495 */
      xw[k] = 0.; uc[k] = 0.; yw[k] = 0.; vc[k] = 0.;
    }

    double se[3], he[3], ue[3], ve[3];
    const int64_t nb = D->number_of_boundaries[k];
    if (nb == 3) {                                   /* :1498-1522 */
      for (int i = 0; i < 3; i++) { se[i] = wc[k]; ue[i] = uc[k]; ve[i] = vc[k]; he[i] = hcv[k]; }
    } else if (nb <= 1) {                            /* :1523-1645 */
      const double hc = hcv[k], h0 = hcv[k0], h1 = hcv[k1], h2 = hcv[k2];
      const double hmin = fmin(fmin(h0, fmin(h1, h2)), hc);
      const double hmax = fmax(fmax(h0, fmax(h1, h2)), hc);
      double hfactor = fmax(0., fmin(c_tmp * fmax(hmin, 0.0) / fmax(hc, 1.0e-06) + d_tmp,
                                     fmin(c_tmp * fmax(hc, 0.) / fmax(hmax, 1.0e-06) + d_tmp, 1.0)));
      hfactor = fmin(1.2 * fmax(hmin - mah, 0.) / (fmax(hmin, 0.) + 1. * mah), hfactor);
      const double inv_area2 = 1.0 / area2;
      double beta = D->beta_w_dry + (D->beta_w - D->beta_w_dry) * hfactor;
      edge_values_3(beta, wc[k], wc[k0], wc[k1], wc[k2], dxv, dyv, dx1, dx2, dy1, dy2, inv_area2, se);
      edge_values_3(beta, hcv[k], hcv[k0], hcv[k1], hcv[k2], dxv, dyv, dx1, dx2, dy1, dy2, inv_area2, he);
      beta = D->beta_uh_dry + (D->beta_uh - D->beta_uh_dry) * hfactor;
      edge_values_3(beta, uc[k], uc[k0], uc[k1], uc[k2], dxv, dyv, dx1, dx2, dy1, dy2, inv_area2, ue);
      beta = D->beta_vh_dry + (D->beta_vh - D->beta_vh_dry) * hfactor;
      edge_values_3(beta, vc[k], vc[k0], vc[k1], vc[k2], dxv, dyv, dx1, dx2, dy1, dy2, inv_area2, ve);
    } else {                                         /* two boundaries :1646-1842 */
      int64_t kk = k3;
      for (; kk < k3 + 3; kk++)
        if (D->surrogate_neighbours[kk] != k) break;
      if (kk == k3 + 3) kk = k3 + 2;   /* reference would read out of row; never happens for nb==2 */
      const int64_t kn = D->surrogate_neighbours[kk];
      dx1 = D->centroid_coordinates[2 * kn] - x;
      dy1 = D->centroid_coordinates[2 * kn + 1] - y;
      const double d2 = dx1 * dx1 + dy1 * dy1;
      dx2 = 1.0 / d2;
      dy2 = dx2 * dy1;
      dx2 *= dx1;
      edge_values_1(D->beta_w, wc[k], wc[kn], dxv, dyv, dx2, dy2, se);
      edge_values_1(D->beta_w, hcv[k], hcv[kn], dxv, dyv, dx2, dy2, he);
      edge_values_1(D->beta_w, uc[k], uc[kn], dxv, dyv, dx2, dy2, ue);
      edge_values_1(D->beta_w, vc[k], vc[kn], dxv, dyv, dx2, dy2, ve);
    }

    for (int i = 0; i < 3; i++) {
      if (vel2 == 1) {                               /* :1851-1860 */
        ue[i] = ue[i] * he[i];
        ve[i] = ve[i] * he[i];
      }
      D->stage_edge_values[k3 + i] = se[i];
      D->height_edge_values[k3 + i] = he[i];
      D->xmom_edge_values[k3 + i] = ue[i];
      D->ymom_edge_values[k3 + i] = ve[i];
      D->bed_edge_values[k3 + i] = se[i] - he[i];    /* :1863-1865 */
    }
    /* vertex values from edge values :1871-1892 */
    double *E[5] = {D->stage_edge_values, D->height_edge_values, D->xmom_edge_values,
                    D->ymom_edge_values, D->bed_edge_values};
    double *V[5] = {D->stage_vertex_values, D->height_vertex_values, D->xmom_vertex_values,
                    D->ymom_vertex_values, D->bed_vertex_values};
    for (int q = 0; q < 5; q++) {
      if (!V[q]) continue;
      const double e0 = E[q][k3], e1 = E[q][k3 + 1], e2 = E[q][k3 + 2];
      V[q][k3 + 0] = e1 + e2 - e0;
      V[q][k3 + 1] = e0 + e2 - e1;
      V[q][k3 + 2] = e0 + e1 - e2;
    }
  }

  if (vel2 == 1)                                     /* loop 3 :1899-1907 */
    for (int64_t k = 0; k < K; k++) {
      uc[k] = xw[k];
      vc[k] = yw[k];
    }
  return 0;
}

/* Manning friction, flat bed form: sw_domain_openmp.c:1954-1986 */
void orc_manning_friction_flat(double g, double eps, int64_t N, const double *w,
                               const double *zv, const double *uh, const double *vh,
                               const double *eta, double *xmom_update, double *ymom_update)
{
  const double seven_thirds = 7.0 / 3.0;
  for (int64_t k = 0; k < N; k++) {
    const double abs_mom = sqrt((uh[k] * uh[k] + vh[k] * vh[k]));
    double S = 0.0;
    if (eta[k] > eps) {
      const double h = w[k] - zv[k];
      if (h >= eps) {
        S = -g * eta[k] * eta[k] * abs_mom;
        S /= pow(h, seven_thirds);
      }
    }
    xmom_update[k] += S * uh[k];
    ymom_update[k] += S * vh[k];
  }
}

/* Manning friction, sloped form: sw_domain_openmp.c:1988-2034, gradient util_ext.h:52-92 */
void orc_manning_friction_sloped(double g, double eps, int64_t N, const double *x,
                                 const double *w, const double *zv, const double *uh,
                                 const double *vh, const double *eta,
                                 double *xmom_update, double *ymom_update)
{
  const double one_third = 1.0 / 3.0, seven_thirds = 7.0 / 3.0;
  for (int64_t k = 0; k < N; k++) {
    double S = 0.0;
    const double z0 = zv[3 * k], z1 = zv[3 * k + 1], z2 = zv[3 * k + 2];
    const double x0 = x[6 * k], y0 = x[6 * k + 1], x1 = x[6 * k + 2], y1 = x[6 * k + 3];
    const double x2 = x[6 * k + 4], y2 = x[6 * k + 5];
    if (eta[k] > eps) {
      const double det = (y2 - y0) * (x1 - x0) - (y1 - y0) * (x2 - x0);
      double zx = (y2 - y0) * (z1 - z0) - (y1 - y0) * (z2 - z0);
      zx /= det;
      double zy = (x1 - x0) * (z2 - z0) - (x2 - x0) * (z1 - z0);
      zy /= det;
      const double zs = sqrt(1.0 + zx * zx + zy * zy);
      const double z = (z0 + z1 + z2) * one_third;
      const double h = w[k] - z;
      if (h >= eps) {
        S = -g * eta[k] * eta[k] * zs * sqrt((uh[k] * uh[k] + vh[k] * vh[k]));
        S /= pow(h, seven_thirds);
      }
    }
    xmom_update[k] += S * uh[k];
    ymom_update[k] += S * vh[k];
  }
}

/* fix_negative_cells: sw_domain_openmp.c:2037-2056 */
int64_t orc_fix_negative_cells(orc_domain *D)
{
  int64_t count = 0;
  for (int64_t k = 0; k < D->number_of_elements; k++) {
    if ((D->stage_centroid_values[k] - D->bed_centroid_values[k] < 0.0) & (D->tri_full_flag[k] > 0)) {
      count++;
      D->stage_centroid_values[k] = D->bed_centroid_values[k];
      D->xmom_centroid_values[k] = 0.0;
      D->ymom_centroid_values[k] = 0.0;
    }
  }
  return count;
}

/* Quantity.update: abstract_2d_finite_volumes/quantity.c:772-820 (three sweeps) */
int64_t orc_update(int64_t N, double timestep, double *centroid_values,
                   const double *explicit_update, double *semi_implicit_update)
{
  for (int64_t k = 0; k < N; k++) {
    const double x = centroid_values[k];
    if (x == 0.0) semi_implicit_update[k] = 0.0;
    else semi_implicit_update[k] /= x;
  }
  for (int64_t k = 0; k < N; k++) centroid_values[k] += timestep * explicit_update[k];
  for (int64_t k = 0; k < N; k++) {
    const double denominator = 1.0 - timestep * semi_implicit_update[k];
    if (denominator <= 0.0) return -1;
    centroid_values[k] /= denominator;
  }
  memset(semi_implicit_update, 0, N * sizeof(double));
  return 0;
}

/* quantity.c:735-750 */
void orc_backup_centroid_values(int64_t N, const double *c, double *backup)
{
  for (int64_t k = 0; k < N; k++) backup[k] = c[k];
}

/* quantity.c:752-769 */
void orc_saxpy_centroid_values(int64_t N, double a, double b, double *c, const double *backup)
{
  for (int64_t k = 0; k < N; k++) c[k] = a * c[k] + b * backup[k];
}
