// swk_kernels.cuh - sm_100a kernels of the DE shallow-water timestep.
//
// Data layout in HBM.  N triangles, NP = N rounded up to 64,
// triangles renumbered by a locality (Morton) permutation at upload.
//   cq   [NP]      d4 {stage, xmom, ymom, bed}        centroid record   (32 B, gathered by neighbours)
//   eq   [3][NP]   d4 {stage, height, xmom, ymom}     edge record       (32 B, gathered by neighbours)
//   xg   [3][NP]   d4 extrapolation geometry          (static, streamed)
//   fg   [3][NP]   d4 flux geometry                   (static, streamed)
//   connA[NP]      i4 {s0, s1, s2, flagsA}            surrogate neighbours
//   connB[NP]      i4 {p0, p1, p2, flagsB}            p = (n<<2)|edge  or  -(m+1) boundary
//   eu   [3][NP]   double explicit updates            (substep 0 round trip)
//   bk   [3][NP]   double RK backup
//   eta  [NP]      double Manning n
//   zflag[NP]      u8  loop-2 "surrounded by dry cells" momentum zeroing
//   bq   [M]       d4 {stage, xmom, ymom, -}          boundary values
// Every record is one 256-bit load/store (LDG.E.256/STG.E.256 on sm_100a); own-cell
// streams are fully coalesced (32 consecutive records per warp), neighbour gathers
// touch exactly one 32-byte sector each.
#pragma once
#include "swk_math.cuh"

namespace swk {

#ifndef SWK_BLOCK
#define SWK_BLOCK 128
#endif
#ifndef SWK_MINB_A          // min resident blocks per SM for pass A (extrapolate)
#define SWK_MINB_A 5
#endif
#ifndef SWK_MINB_F          // ... for the flux kernel of substep 0
#define SWK_MINB_F 5
#endif
#ifndef SWK_MINB_FU         // ... for the fused flux + update kernel
#define SWK_MINB_FU 7
#endif
#ifndef SWK_FU_ROLLED       // fused kernel: one edge per trip of a rolled loop (see triangle_flux)
#define SWK_FU_ROLLED true
#endif
#ifndef SWK_MINB_U
#define SWK_MINB_U 8
#endif
#ifndef SWK_F_ROLLED        // the same for the flux kernel of substep 0
#define SWK_F_ROLLED false
#endif
constexpr int BLOCK = SWK_BLOCK;

// One tile of BLOCK consecutive triangles per CTA.  (A persistent form - as many CTAs as fit on the GPU,
// each walking a contiguous run of tiles and prefetching its own next tiles - was measured 20-40 %
// slower for every kernel: profiles/r2/sweep3_persistent.txt.)
// First triangle of the tile whose slabs a CTA asks the L2 for (SWK_PF_AHEAD tiles further on), or -1.
__device__ __forceinline__ long long prefetch_target(int k0)
{
#if SWK_PF_AHEAD > 0
  return (long long)k0 + (long long)(blockIdx.x + SWK_PF_AHEAD) * BLOCK;
#else
  (void)k0;
  return -1;
#endif
}

// ---- device-side clock: the scalars of Generic_Domain's time loop -------------
struct Clock {
  double time;                  // relative_time
  double step_start_time;       // relative time at the start of the current step
  double dt;                    // self.timestep
  double flux_dt;               // self.flux_timestep (as returned by compute_fluxes)
  unsigned long long dt_min_bits;   // running min over the flux kernel (uint64 image of a positive double)
  double yieldtime, finaltime;  // relative; finaltime < 0: none
  double recorded_min_timestep, recorded_max_timestep;
  double boundary_flux_sum[3];
  double boundary_flux_integral;
  double fractional_step_volume_integral;
  double mass_error;
  long long number_of_steps, number_of_first_order_steps, total_steps, step_budget;
  long long negative_cells;
  int smallsteps, order;
  int stop;                     // 0 run, 1 yield reached, 2 final reached, 3 budget exhausted, <0 error
  int pad;
};

struct TimeParams {
  double CFL, evolve_max_timestep, evolve_min_timestep, fixed_flux_timestep, epsilon;
  int max_smallsteps, default_order, method;   // method: 1 euler 2 rk2 3 rk3
};

// ---- pointers --------------------------------------------------------------
struct Dev {
  int N, NP, M;
  d4 *cq;
  d4 *eq;
  const d4 *xg;
  const d4 *fg;
  const i4 *connA;
  const i4 *connB;
  double *eu;
  double *bk;
  const double *eta;
  unsigned char *zflag;
  double *max_speed;
  d4 *bq;
  const double *vcoord;          // (6*NP) vertex coordinates x0,y0,x1,y1,x2,y2 planes (sloped Manning), may be null
  // per-call layer only: edge beds / centroid heights taken from the caller's arrays instead of being
  // recomputed as stage_e - height_e and max(stage_c - bed_c, 0) (any consistent or inconsistent input)
  const double *bed_e_x;         // [3][NP] or null
  const double *hc_x;            // [NP] or null
  const double *wind;            // [3][NP] state-independent explicit forcing of stage, xmom, ymom (Wind_stress,
                                 // General_forcing / Rainfall / Inflow: forcing.py:80-640), may be null
  Clock *clock;
  // boundary-flux accounting (sw_domain_openmp.c:696-701): slot per accounting edge, in (k, i) order
  double *acct_val;              // [n_acct]
  const int *pos_b;              // [M] slot of boundary edge m, -1 if its cell is a ghost
  const int *acct_keys;          // sorted (k<<2|i) of full-cell edges facing a ghost cell
  const int *acct_keys_pos;      // their slots
  int n_acct, n_acct_keys;
  // riverwalls
  const int *rw_counter;         // [3][NP] edge_river_wall_counter (1-based) or null
  const double *rw_elevation;
  const int *rw_rowIndex;
  const double *rw_hydraulic;
  int rw_ncol;
};

// =============================================================================
// Pass A: protect + second-order edge extrapolation with limiting.
// Reference: _openmp_protect (sw_domain_openmp.c:1096-1171) and
// _openmp_extrapolate_second_order_edge_sw (:1336-1952), fused; centroid values
// are NOT modified (their protected form is recomputed by the consumers).
// One thread per triangle.  Reads cq (own + 3 gathered), connA, xg; writes eq, zflag.
// =============================================================================
// Edge records of ONE triangle k from the centroid records `cq` (own + 3 surrogate neighbours).
// Returns the three {stage, height, xmom, ymom} edge records, the own effective state and the
// loop-2 "surrounded by dry cells" flag.  count_mass: add protect's mass error to the clock
// (only for triangles a block owns, not for halo re-evaluations).
__device__ __forceinline__ void extrapolate_tri(const Dev &D, const Consts &K, const d4 *__restrict__ cq,
                                                int k, bool count_mass, d4 &r0, d4 &r1, d4 &r2,
                                                Eff &e, bool &zero_mom, int &connA_flags)
{
  const int NP = D.NP;
  const i4 s = lds(&D.connA[k]);
  const d4 c = cq[k];
  const d4 c0 = cq[s.x];
  const d4 c1 = cq[s.y];
  const d4 c2 = cq[s.z];
  const d4 g0 = lds(&D.xg[k]);
  const d4 g1 = lds(&D.xg[NP + k]);
  XGeom G;
  G.dxv0 = g0.x; G.dxv1 = g0.y; G.dxv2 = g0.z; G.dyv0 = g0.w;
  G.dyv1 = g1.x; G.dyv2 = g1.y;
  const d4 g2 = lds(&D.xg[2 * NP + k]);
  G.dx1 = g1.z; G.dx2 = g1.w;
  G.dy1 = g2.x; G.dy2 = g2.y; G.inv_area2 = g2.z;
  connA_flags = s.w;

  e = effective(c, K);
  const Eff e0 = effective(c0, K);
  const Eff e1 = effective(c1, K);
  const Eff e2 = effective(c2, K);
  if (count_mass && e.mass_added != 0.0) atomicAdd(&D.clock->mass_error, e.mass_added * D.fg[2 * NP + k].w);

  const int nb = s.w & 3;
  // loop 2 head (:1486-1495): all neighbours dry (or self) -> no momentum
  const bool dry0 = (e0.h < K.mah) | (s.x == k);
  const bool dry1 = (e1.h < K.mah) | (s.y == k);
  const bool dry2 = (e2.h < K.mah) | (s.z == k);
  zero_mom = dry0 & dry1 & dry2;
  if (zero_mom) { e.u = 0.0; e.v = 0.0; }

  double w0, w1, w2, h0, h1, h2, u0, u1, u2, v0, v1, v2;
  if (nb == 3) {                                   // :1498-1522
    w0 = w1 = w2 = e.w;
    h0 = h1 = h2 = e.h;
    u0 = u1 = u2 = e.u;
    v0 = v1 = v2 = e.v;
  } else if (nb <= 1) {                            // :1523-1645
    const double a_tmp = 0.3, b_tmp = 0.1;
    const double c_tmp = 1.0 / (a_tmp - b_tmp);
    const double d_tmp = 1.0 - (c_tmp * a_tmp);
    const double hc = e.h;
    const double hmin = dmin(dmin(e0.h, dmin(e1.h, e2.h)), hc);
    const double hmax = dmax(dmax(e0.h, dmax(e1.h, e2.h)), hc);
    double hfactor;
    // Well inside the wet region all three quotients below are >= 1 and hfactor is exactly 1.0; that is
    // decided without dividing (margin 1e-12 over the roundings: x/y + d >= 1 whenever x >= y*(1 - d)*(1 + 1e-12))
    // and taken only when every active lane of the warp agrees.  Otherwise the reference's expression runs.
    const double n1_ = c_tmp * dmax0(hmin), y1_ = dmax(hc, 1.0e-06);
    const double n2_ = c_tmp * dmax0(hc), y2_ = dmax(hmax, 1.0e-06);
    const double n3_ = 1.2 * dmax0(hmin - K.mah), y3_ = dmax0(hmin) + 1. * K.mah;
    const double lift = (1.0 - d_tmp) * (1.0 + 1.0e-12);
    const bool all_one = (n1_ >= y1_ * lift) & (n2_ >= y2_ * lift) & (n3_ >= y3_ * (1.0 + 1.0e-12));
    if (__all_sync(__activemask(), all_one)) {
      hfactor = 1.0;
    } else {
      hfactor = dmax0(dmin(n1_ / y1_ + d_tmp, dmin(n2_ / y2_ + d_tmp, 1.0)));
      hfactor = dmin(n3_ / y3_, hfactor);
    }
    double beta = K.beta_w_dry + (K.beta_w - K.beta_w_dry) * hfactor;
    edge_values_3(beta, e.w, e0.w, e1.w, e2.w, G, w0, w1, w2);
    edge_values_3(beta, e.h, e0.h, e1.h, e2.h, G, h0, h1, h2);
    beta = K.beta_uh_dry + (K.beta_uh - K.beta_uh_dry) * hfactor;
    edge_values_3(beta, e.u, e0.u, e1.u, e2.u, G, u0, u1, u2);
    beta = K.beta_vh_dry + (K.beta_vh - K.beta_vh_dry) * hfactor;
    edge_values_3(beta, e.v, e0.v, e1.v, e2.v, G, v0, v1, v2);
  } else {                                         // two boundary edges :1646-1842
    const int which = (s.w >> 2) & 3;             // value selects: a reference would put e0..e2 in local memory
    const bool is0 = which == 0, is1 = which == 1;
    const double nw = is0 ? e0.w : (is1 ? e1.w : e2.w);
    const double nh = is0 ? e0.h : (is1 ? e1.h : e2.h);
    const double nu = is0 ? e0.u : (is1 ? e1.u : e2.u);
    const double nv = is0 ? e0.v : (is1 ? e1.v : e2.v);
    edge_values_1(K.beta_w, e.w, nw, G, w0, w1, w2);
    edge_values_1(K.beta_w, e.h, nh, G, h0, h1, h2);
    edge_values_1(K.beta_w, e.u, nu, G, u0, u1, u2);
    edge_values_1(K.beta_w, e.v, nv, G, v0, v1, v2);
  }
  if (K.vel2) {                                    // :1851-1860
    u0 = u0 * h0; v0 = v0 * h0;
    u1 = u1 * h1; v1 = v1 * h1;
    u2 = u2 * h2; v2 = v2 * h2;
  }
  r0.x = w0; r0.y = h0; r0.z = u0; r0.w = v0;
  r1.x = w1; r1.y = h1; r1.z = u1; r1.w = v1;
  r2.x = w2; r2.y = h2; r2.z = u2; r2.w = v2;
}

__global__ void __launch_bounds__(BLOCK, SWK_MINB_A) k_extrapolate(Dev D, Consts K)
{
  if (D.clock->stop) return;
  const int k = blockIdx.x * BLOCK + threadIdx.x;
#if SWK_PF_AHEAD > 0
  if (threadIdx.x < 5) {      // slabs of a tile further on: cq, xg x3, connA
    const long long t0 = prefetch_target(0);
    if (t0 + BLOCK <= D.NP) {
      const int j = threadIdx.x;
      if (j == 0) prefetch_l2_bulk(D.cq + t0, BLOCK * 32);
      else if (j < 4) prefetch_l2_bulk(D.xg + (long long)(j - 1) * D.NP + t0, BLOCK * 32);
      else if (j == 4) prefetch_l2_bulk(D.connA + t0, BLOCK * 16);
    }
  }
#endif
  if (k >= D.N) return;
  d4 r0, r1, r2;
  Eff e;
  bool zero_mom;
  int fl;
  extrapolate_tri(D, K, D.cq, k, true, r0, r1, r2, e, zero_mom, fl);
  D.zflag[k] = (zero_mom ? 1 : 0) | ((fl >> 3) & 2);   // bit0: zeroed momenta, bit1: tri_full_flag
  D.eq[k] = r0;
  D.eq[D.NP + k] = r1;
  D.eq[2 * D.NP + k] = r2;
}

// Write the protected/zeroed centroid values in place: what the reference's centroid
// arrays hold after distribute_to_vertices_and_edges (used at yields and by the
// per-call layer).  Must run after k_extrapolate (needs zflag).
__global__ void __launch_bounds__(BLOCK) k_materialize_centroids(Dev D, Consts K, int protect_only)
{
  if (D.clock->stop) return;
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= D.N) return;
  d4 c = D.cq[k];
  if (protect_only) {                 // exactly _openmp_protect, incl. the xmom-only zeroing (:1139-1140)
    const double hc = c.x - c.w;
    if (hc < K.mah * 1.0) {
      c.y = 0.0;
      if (hc <= 0.0 && c.x < c.w) c.x = c.w;
    }
  } else {
    const Eff e = effective(c, K);
    c.x = e.w; c.y = e.uh; c.z = e.vh;
    if (D.zflag[k] & 1) { c.y = 0.0; c.z = 0.0; }
  }
  D.cq[k] = c;
}

// protect alone with its mass error (per-call layer / swk_protect)
__global__ void __launch_bounds__(BLOCK) k_protect_mass(Dev D, Consts K)
{
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= D.N) return;
  const d4 c = D.cq[k];
  if (c.x < c.w) atomicAdd(&D.clock->mass_error, (c.w - c.x) * D.fg[2 * D.NP + k].w);
}

// =============================================================================
// Boundary values: Generic_Domain.update_boundary (generic_domain.py:2288-2306)
// with the evaluate_segment arithmetic of each boundary class.  One thread per
// boundary edge m.
// =============================================================================
// Flux geometry of triangle k: fg[i][k] = {nx_i, ny_i, length_i, aux_i}, aux = {1/area, radius, area}.
__device__ __forceinline__ void edge_normal(const Dev &D, int k, int i, double &n1, double &n2)
{
  const d4 f = D.fg[i * D.NP + k];
  n1 = f.x; n2 = f.y;
}

struct Segments {
  const int *b_cell;       // [M] triangle (device numbering)
  const int *b_edge;       // [M]
  const int *b_seg;        // [M] segment id or -1
  const int *seg_kind;     // [nseg]
  const double *seg_val;   // [nseg][3 substeps][3]: values of time-dependent kinds differ per RK substep
  int substep;
  // time-space tables (File_boundary / Field_boundary / Time_space_boundary): frames[frame][point][3] of
  // segment s start at seg_tab[s], hold seg_np[s] points per frame; b_point[m] = point of boundary edge m
  const double *const *seg_tab;
  const int *seg_np;
  const int *b_point;
};

// boundary value of one edge from the triangle's own edge record e = {stage, height, xmom, ymom},
// its outward normal and (centroid-transmissive only) its protected centroid state
__device__ __forceinline__ bool boundary_value_core(int kind, double v0, double v1, double v2, const d4 e,
                                                    double n1, double n2, int centroid_transmissive,
                                                    double cw, double cuh, double cvh, d4 &out,
                                                    double bed_c = 0.0, double g = 0.0)
{
  switch (kind) {
    case 1: {                                  // Reflective (boundaries.py:262-278)
      const double q1 = e.z, q2 = e.w;
      const double r1 = -q1 * n1 - q2 * n2;
      const double r2 = -q1 * n2 + q2 * n1;
      out.x = e.x;
      out.y = n1 * r1 - n2 * r2;
      out.z = n2 * r1 + n1 * r2;
    } break;
    case 2:                                    // Dirichlet / host-evaluated time boundary
      out.x = v0; out.y = v1; out.z = v2;
      break;
    case 3:                                    // Transmissive
      if (centroid_transmissive) { out.x = cw; out.y = cuh; out.z = cvh; }
      else { out.x = e.x; out.y = e.z; out.z = e.w; }
      break;
    case 4: {                                  // Transmissive_n_momentum_zero_t_momentum_set_stage
      const double ndotq = n1 * e.z + n2 * e.w;
      out.x = v0;
      out.y = ndotq * n1;
      out.z = ndotq * n2;
    } break;
    case 5:                                    // Transmissive_momentum_set_stage
      out.x = v0; out.y = e.z; out.z = e.w;
      break;
    case 6:                                    // Transmissive_stage_zero_momentum
      out.x = e.x; out.y = 0.0; out.z = 0.0;
      break;
    case 7: {                                  // Flather_external_stage_zero_velocity (boundaries.py:1207-1266)
      const double sb = e.x, xb = e.z, yb = e.w, eb = e.x - e.y;
      const double depth = dmax0(sb - bed_c);
      const double so = 0.0 * sb + v0;
      // "dry" also whenever the external stage is above the cell's bed (:1255) - kept as is
      if (depth == 0.0 || so > bed_c) {
        out.x = (bed_c <= so) ? so : eb;
        out.y = 0.0 * xb;
        out.z = 0.0 * yb;
      } else {
        const double s = sqrt(g / depth);
        const double ndotq = n1 * xb + n2 * yb;
        const double w1 = 0.0 - s * so;
        const double w2 = (ndotq > 0.0) ? (n2 * xb - n1 * yb) / depth : 0.0 * ndotq;
        const double w3 = ndotq / depth + s * sb;
        const double qperp = (w3 + w1) / 2.0 * depth;
        const double qpar = w2 * depth;
        out.x = (w3 - w1) / (2.0 * s);
        out.y = qperp * n1 + qpar * n2;
        out.z = qperp * n2 - qpar * n1;
      }
    } break;
    case 8: {                                  // Characteristic_stage (boundaries.py:760-843, vectorised form)
      const double sb = e.x, xb = e.z, yb = e.w, eb = e.x - e.y;
      const double h_inside = dmax0(sb - eb);
      const double w_outside = 0.0 * sb + v0;
      const double uh_inside = n1 * xb + n2 * yb;
      const double vh_inside = n2 * xb - n1 * yb;
      const double u_inside = (h_inside > 0.0) ? uh_inside / h_inside : 0.0;
      const double h_outside = dmax0(w_outside - eb);
      if (h_inside == 0.0 || h_outside == 0.0) {
        out.x = w_outside; out.y = 0.0; out.z = 0.0;
      } else {
        const double sqrt_g = g;                 // the caller passes gravity**0.5 for this kind
        const double sqrt_h_inside = sqrt(h_inside), sqrt_h_outside = sqrt(h_outside);
        const double r = 0.5 * (sqrt_h_inside + sqrt_h_outside) + u_inside / 4.0 / sqrt_g;
        const double h_m = r * r;
        const double u_m = 0.5 * u_inside + sqrt_g * (sqrt_h_inside - sqrt_h_outside);
        const double uh_m = h_m * u_m;
        const double vh_m = (uh_inside > 0.0) ? vh_inside : 0.0;    // outflow keeps its tangential momentum
        out.x = h_m + eb;
        out.y = uh_m * n1 + vh_m * n2;
        out.z = uh_m * n2 - vh_m * n1;
      }
    } break;
    case 11:                                   // Dirichlet_discharge (boundaries.py:845-890): stage0, wh0 inwards
      out.x = v0;
      out.y = -v1 * n1;
      out.z = -v1 * n2;
      break;
    default:
      return false;
  }
  out.w = 0.0;
  return true;
}

__device__ __forceinline__ void boundary_value(const Dev &D, const Segments &S, const Consts &K,
                                               int m, int centroid_transmissive, d4 &out, bool &touched)
{
  touched = false;
  const int seg = S.b_seg[m];
  if (seg < 0) return;
  const int kind = S.seg_kind[seg];
  if (kind == 0) return;
  const double *sv = S.seg_val + (3 * seg + S.substep) * 3;
  if (kind == 9 || kind == 10) {
    // frames resident in HBM, linear in time between frame idx and idx + 1 (Interpolation_function.__call__,
    // fit_interpolate/interpolate.py:1056-1092): q = Q0 + ratio*(Q1 - Q0); kind 10 adds mean_stage to the stage
    const double ratio = sv[0];
    const int idx = (int)sv[1];
    const int np_ = S.seg_np[seg];
    const double *q0 = S.seg_tab[seg] + ((long long)idx * np_ + S.b_point[m]) * 3;
    double v[3] = {q0[0], q0[1], q0[2]};
    if (ratio > 0.0) {
      const double *q1 = q0 + (long long)np_ * 3;
#pragma unroll
      for (int j = 0; j < 3; j++) v[j] = q0[j] + ratio * (q1[j] - q0[j]);
    }
    out.x = (kind == 10) ? v[0] + sv[2] : v[0];
    out.y = v[1];
    out.z = v[2];
    out.w = 0.0;
    touched = true;
    return;
  }
  const int k = S.b_cell[m];
  const int i = S.b_edge[m];
  const d4 e = D.eq[i * D.NP + k];           // {stage, height, xmom, ymom}
  double n1, n2;
  edge_normal(D, k, i, n1, n2);
  double cw = 0.0, cuh = 0.0, cvh = 0.0;
  if (centroid_transmissive && kind == 3) {   // centroid arrays as the reference sees them
    const Eff ef = effective(D.cq[k], K);
    cw = ef.w; cuh = ef.uh; cvh = ef.vh;
    if (D.zflag[k] & 1) { cuh = 0.0; cvh = 0.0; }
  }
  const double bed_c = (kind == 7) ? D.cq[k].w : 0.0;
  touched = boundary_value_core(kind, sv[0], sv[1], sv[2], e,
                                n1, n2, centroid_transmissive, cw, cuh, cvh, out, bed_c, (kind == 8) ? K.sqrt_g : K.g);
}

__global__ void __launch_bounds__(BLOCK) k_boundary_values(Dev D, Segments S, Consts K, int centroid_transmissive)
{
  if (D.clock->stop) return;
  const int m = blockIdx.x * BLOCK + threadIdx.x;
  if (m >= D.M) return;
  d4 out;
  bool touched;
  boundary_value(D, S, K, m, centroid_transmissive, out, touched);
  if (touched) D.bq[m] = out;
}

// =============================================================================
// Flux of one triangle: _openmp_compute_fluxes_central loop body (:521-719).
// =============================================================================
struct TriFlux {
  double su, xu, yu;      // explicit updates (already scaled by 1/area)
  double dtmin;           // min edge timestep of this triangle (1e100 if none)
  double speed;           // max_speed[k]
};

// slot of an accounting edge that faces a ghost cell (multi-GPU sub-domains only)
__device__ __noinline__ int acct_slot_lookup(const int *keys, const int *keys_pos, int n, int key)
{
  int lo = 0, hi = n - 1;
  while (lo <= hi) {
    const int mid = (lo + hi) >> 1;
    const int v = keys[mid];
    if (v == key) return keys_pos[mid];
    if (v < key) lo = mid + 1; else hi = mid - 1;
  }
  return -1;
}

// el[i]: own edge records; er[i]: the neighbour's record of the shared edge, or for a boundary edge
// (pn[i] < 0) the boundary value {stage, xmom, ymom, -}.
// contribution of edge i of triangle k: el = own edge record, er = the neighbour's record of the shared
// edge or, for a boundary edge (q < 0), the boundary value {stage, xmom, ymom, -}
template <bool RW, bool XB = false>
__device__ __forceinline__ void edge_contribution(const Dev &D, const Consts &K, int k, int i, int q, int flags,
                                                  const d4 el, const d4 er, double nx, double ny, double length,
                                                  const Eff &own, bool first, TriFlux &T)
{
  const int NP = D.NP;
  const bool full = flags & 1;
  const double hc = XB ? D.hc_x[k] : own.h, zc = own.z;
  const double wl = el.x, hle = el.y, uhl = el.z, vhl = el.w;
  const double zl = XB ? D.bed_e_x[i * NP + k] : wl - hle;   // bed_edge = stage_edge - height_edge (:1863)
  double wr, uhr, vhr, zr, hre;
  if (q < 0) {                                        // :551-561
    wr = er.x; uhr = er.y; vhr = er.z;
    zr = zl;
    hre = dmax0(wr - zr);
  } else {                                            // :562-576
    wr = er.x; hre = er.y; uhr = er.z; vhr = er.w;
    zr = XB ? D.bed_e_x[(q & 3) * NP + (q >> 2)] : wr - hre;
  }
  double z_half = dmax(zl, zr);
  bool rw_edge = false;
  int rwc = 0;
  if (RW) {
    rw_edge = (flags >> (1 + i)) & 1;
    if (rw_edge) {                                    // :582-588
      rwc = D.rw_counter[i * NP + k];
      z_half = dmax(D.rw_elevation[rwc - 1], z_half);
    }
  }
  const double h_left = dmax0(hle + zl - z_half);
  const double h_right = dmax0(hre + zr - z_half);
  EdgeFlux F = edge_flux_central(wl, uhl, vhl, wr, uhr, vhr, h_left, h_right, hle, hre, nx, ny, z_half, K);
  if (RW) {
    if (rw_edge) {                                    // :607-653
      const int ii = D.rw_rowIndex[rwc - 1] * D.rw_ncol;
      const double Qfactor = D.rw_hydraulic[ii];
      const double s1 = D.rw_hydraulic[ii + 1];
      const double s2 = D.rw_hydraulic[ii + 2];
      const double h1 = D.rw_hydraulic[ii + 3];
      const double h2 = D.rw_hydraulic[ii + 4];
      const double rw_elev = D.rw_elevation[rwc - 1];
      const double weir_height = dmax0(rw_elev - dmin(zl, zr));
      const double h_left_tmp = dmax0(own.w - z_half);
      double h_right_tmp, zc_n = zc;
      if (q >= 0) {
        const Eff en = effective(D.cq[q >> 2], K);
        zc_n = en.z;
        h_right_tmp = dmax0(en.w - z_half);
      } else {
        h_right_tmp = dmax0(hc + zr - z_half);
      }
      if (rw_elev > dmax(zc, zc_n))
        weir_adjust(F, h_left_tmp, h_right_tmp, K.g, weir_height, Qfactor, s1, s2, h1, h2);
    }
  }
  const double ef0 = -F.f0 * length;
  const double ef1 = -F.f1 * length;
  const double ef2 = -F.f2 * length;
  const double pressuregrad =
      length * (-K.g * 0.5 * (h_left * h_left - hle * hle - (hle + hc) * (zl - zc)) + F.pressure_flux);
  if (first) {                                        // :667-686, division hoisted out of the loop
    if (full && F.max_speed > K.epsilon) T.speed = dmax(T.speed, F.max_speed);
  }
  T.su += ef0;
  T.xu += ef1;
  T.yu += ef2;
  if (q < 0) {                                        // boundary_flux_sum terms (:696-701)
    const int slot = D.pos_b[-q - 1];
    if (slot >= 0) D.acct_val[slot] = ef0;
  } else if ((flags >> (4 + i)) & 1) {
    const int slot = acct_slot_lookup(D.acct_keys, D.acct_keys_pos, D.n_acct_keys, (k << 2) | i);
    if (slot >= 0) D.acct_val[slot] = ef0;
  }
  T.xu -= nx * pressuregrad;
  T.yu -= ny * pressuregrad;
}

__device__ __forceinline__ void finish_flux(TriFlux &T, double radius, double inv_area, bool first)
{
  // min_i fl(radius / speed_i) == fl(radius / max_i speed_i): correctly rounded division is monotone,
  // so one division per triangle gives the reference's local timestep bit for bit
  if (first && T.speed > 0.0) T.dtmin = radius * 1.0 / T.speed;
  T.su *= inv_area;
  T.xu *= inv_area;
  T.yu *= inv_area;
}

template <bool RW, bool ROLLED = false, bool XB = false>
__device__ __forceinline__ TriFlux triangle_flux(const Dev &D, const Consts &K, int k, const i4 p,
                                                 const Eff &own, bool first)
{
  const int NP = D.NP;
  TriFlux T;
  T.su = 0.0; T.xu = 0.0; T.yu = 0.0;
  T.dtmin = 1.0e+100;
  T.speed = 0.0;
  double inv_area = 0.0, radius = 0.0;
  if (ROLLED) {
  // one edge per trip of a rolled loop: the trip loads the three records of its edge (own edge
  // values, the neighbour's, the edge geometry), so only one edge's operands are live at a time
#pragma unroll 1
  for (int i = 0; i < 3; i++) {
    const int q = (i == 0) ? p.x : ((i == 1) ? p.y : p.z);
    const d4 el = D.eq[i * NP + k];
    const d4 ge = lds(&D.fg[i * NP + k]);
    d4 er;
    if (q >= 0) er = D.eq[(q & 3) * NP + (q >> 2)];
    else er = D.bq[-q - 1];
    if (i == 0) inv_area = ge.w;
    if (i == 1) radius = ge.w;
    edge_contribution<RW, XB>(D, K, k, i, q, p.w, el, er, ge.x, ge.y, ge.z, own, first, T);
  }
  } else {
  d4 el[3], ge[3], er[3];
  const int pn[3] = {p.x, p.y, p.z};
#pragma unroll
  for (int i = 0; i < 3; i++) {
    el[i] = D.eq[i * NP + k];
    ge[i] = lds(&D.fg[i * NP + k]);
  }
#pragma unroll
  for (int i = 0; i < 3; i++) {
    const int q = pn[i];
    if (q >= 0) er[i] = D.eq[(q & 3) * NP + (q >> 2)];
    else er[i] = D.bq[-q - 1];
  }
  inv_area = ge[0].w;
  radius = ge[1].w;
#pragma unroll
  for (int i = 0; i < 3; i++)
    edge_contribution<RW, XB>(D, K, k, i, pn[i], p.w, el[i], er[i], ge[i].x, ge[i].y, ge[i].z, own, first, T);
  }
  finish_flux(T, radius, inv_area, first);
  return T;
}

// L2 prefetch of the slabs of the tile that starts at triangle t0 (flux kernels)
__device__ __forceinline__ void prefetch_flux_slabs(const Dev &D, long long t0, bool with_update)
{
  const int j = threadIdx.x;
  if (j < 12 && t0 >= 0 && t0 + BLOCK <= D.NP) {
    const long long NP = D.NP;
    if (j < 3) prefetch_l2_bulk(D.eq + j * NP + t0, BLOCK * 32);
    else if (j < 6) prefetch_l2_bulk(D.fg + (j - 3) * NP + t0, BLOCK * 32);
    else if (j == 6) prefetch_l2_bulk(D.connB + t0, BLOCK * 16);
    else if (j == 7) prefetch_l2_bulk(D.cq + t0, BLOCK * 32);
    else if (with_update) {
      if (j == 8) prefetch_l2_bulk(D.eta + t0, BLOCK * 8);
      else prefetch_l2_bulk(D.bk + (j - 9) * NP + t0, BLOCK * 8);
    }
  }
}

// block-wide min of positive doubles -> one atomicMin per block
template <int NT>
__device__ __forceinline__ void block_min_to_clock_t(double v, Clock *clock)
{
  __shared__ double smin[NT / 32];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = dmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  if (lane == 0) smin[wid] = v;
  __syncthreads();
  if (wid == 0) {
    v = (lane < NT / 32) ? smin[lane] : 1.0e+100;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (lane == 0 && v < 1.0e+100) atomicMin(&clock->dt_min_bits, d2u(v));
  }
}

__device__ __forceinline__ void block_min_to_clock(double v, Clock *clock) { block_min_to_clock_t<BLOCK>(v, clock); }

// update of one triangle's conserved quantities from its explicit updates:
// friction (sw_domain_openmp.c:1954-2034) -> Quantity.update x3 (quantity.c:772-820)
// -> fix_negative_cells (:2037-2056) -> optional RK combine (quantity.c:752-769,
// generic_domain.py:2045, 2126, 2167-2170).
struct UpdateArgs {
  double a, b, divide_by;     // saxpy coefficients; combine only if do_saxpy
  double g;
  double rain_rate, rain_factor;   // fused scalar Rate_operator (rate >= 0, all cells), see k_finish_step
  int do_backup, do_saxpy, sloped, do_rain;
};

__device__ __forceinline__ void triangle_update(const Dev &D, const Consts &K, const UpdateArgs &U,
                                                int k, const d4 raw, Eff e, const unsigned zf,
                                                double su, double xu, double yu, double dt,
                                                d4 *__restrict__ cq_out, const d4 *own_edges)
{
  const int NP = D.NP;
  const bool full = (zf & 2) != 0;
  if (U.do_backup) {                                   // backup holds the RAW start-of-step values
    sts(&D.bk[k], raw.x);
    sts(&D.bk[NP + k], raw.y);
    sts(&D.bk[2 * NP + k], raw.z);
  }
  if (zf & 1) { e.uh = 0.0; e.vh = 0.0; }
  if (D.wind) {                                        // compute_forcing_terms: explicit_update += forcing
    su += lds(&D.wind[k]);
    xu += lds(&D.wind[NP + k]);
    yu += lds(&D.wind[2 * NP + k]);
  }
  double zs = 1.0;
  double h = e.w - e.z;
  if (U.sloped) {                                      // :2003-2023 with the dynamic bed vertex values
    const d4 a0 = own_edges ? own_edges[0] : D.eq[k];
    const d4 a1 = own_edges ? own_edges[1] : D.eq[NP + k];
    const d4 a2 = own_edges ? own_edges[2] : D.eq[2 * NP + k];
    const double b0 = a0.x - a0.y, b1 = a1.x - a1.y, b2 = a2.x - a2.y;
    const double z0 = b1 + b2 - b0, z1 = b0 + b2 - b1, z2 = b0 + b1 - b2;
    const double x0 = D.vcoord[k], y0 = D.vcoord[NP + k], x1 = D.vcoord[2 * NP + k];
    const double y1 = D.vcoord[3 * NP + k], x2 = D.vcoord[4 * NP + k], y2 = D.vcoord[5 * NP + k];
    const double det = (y2 - y0) * (x1 - x0) - (y1 - y0) * (x2 - x0);
    double zx = (y2 - y0) * (z1 - z0) - (y1 - y0) * (z2 - z0);
    zx /= det;
    double zy = (x1 - x0) * (z2 - z0) - (x2 - x0) * (z1 - z0);
    zy /= det;
    zs = sqrt(1.0 + zx * zx + zy * zy);
    h = e.w - (z0 + z1 + z2) * (1.0 / 3.0);
  }
  const double S = manning_S(U.g, K.mah, lds(&D.eta[k]), h, e.uh, e.vh, zs, U.sloped);
  double w = e.w, uh = e.uh, vh = e.vh;
  w += dt * su;                                        // stage has no semi-implicit term: /1.0 is exact
  bool ok = true;
  ok &= update_value(uh, dt, xu, 0.0 + S * e.uh);
  ok &= update_value(vh, dt, yu, 0.0 + S * e.vh);
  if (!ok) D.clock->stop = -3;                         // SWK_ERR_DENOMINATOR
  if ((w - e.z < 0.0) & full) {                        // fix_negative_cells
    w = e.z; uh = 0.0; vh = 0.0;
    atomicAdd((unsigned long long *)&D.clock->negative_cells, 1ULL);
  }
  if (U.do_saxpy) {
    w = U.a * w + U.b * lds(&D.bk[k]);
    uh = U.a * uh + U.b * lds(&D.bk[NP + k]);
    vh = U.a * vh + U.b * lds(&D.bk[2 * NP + k]);
    if (U.divide_by != 1.0) {
      w = w / U.divide_by;
      uh = uh / U.divide_by;
      vh = vh / U.divide_by;
    }
  }
  if (U.do_rain) w = w + U.rain_factor * dt * U.rain_rate;   // rate_operators.py:205-208 (all rates >= 0)
  d4 out;
  out.x = w; out.y = uh; out.z = vh; out.w = e.z;
  cq_out[k] = out;
}

// Pass B1 (substep 0): flux + dt partials.  writes eu, max_speed, dt_min_bits.
template <bool RW, bool XB = false>
__global__ void __launch_bounds__(BLOCK, SWK_MINB_F) k_flux(Dev D, Consts K, int first, int write_speed,
                                                            int k0, int k1)
{
  if (D.clock->stop) return;
  const int k = k0 + blockIdx.x * BLOCK + threadIdx.x;
#if SWK_PF_AHEAD > 0
  prefetch_flux_slabs(D, prefetch_target(k0), false);
#endif
  double dtmin = 1.0e+100;
  if (k < k1) {
    const i4 p = lds(&D.connB[k]);
    const Eff own = effective(D.cq[k], K);
    const TriFlux T = triangle_flux<RW, SWK_F_ROLLED, XB>(D, K, k, p, own, first != 0);
    sts(&D.eu[k], T.su);
    sts(&D.eu[D.NP + k], T.xu);
    sts(&D.eu[2 * D.NP + k], T.yu);
    if (first && write_speed) D.max_speed[k] = T.speed;
    dtmin = T.dtmin;
  }
  if (first) block_min_to_clock(dtmin, D.clock);
}

// Pass B2 (substep 0): friction + update + fix-negative (+ RK backup), dt from the clock.
__global__ void __launch_bounds__(BLOCK, SWK_MINB_U) k_update(Dev D, Consts K, UpdateArgs U, double dt_override,
                                                  int k0, int k1)
{
  if (D.clock->stop) return;
  const double dt = (dt_override >= 0.0) ? dt_override : D.clock->dt;
  const int k = k0 + blockIdx.x * BLOCK + threadIdx.x;
#if SWK_PF_AHEAD > 0
  if (threadIdx.x < 5) {
    const long long t0 = prefetch_target(k0);
    if (t0 + BLOCK <= D.NP) {
      const int j = threadIdx.x;
      if (j < 3) prefetch_l2_bulk(D.eu + (long long)j * D.NP + t0, BLOCK * 8);
      else if (j == 3) prefetch_l2_bulk(D.cq + t0, BLOCK * 32);
      else prefetch_l2_bulk(D.eta + t0, BLOCK * 8);
    }
  }
#endif
  if (k >= k1) return;
  const d4 raw = D.cq[k];
  const Eff e = effective(raw, K);
  triangle_update(D, K, U, k, raw, e, D.zflag[k], lds(&D.eu[k]), lds(&D.eu[D.NP + k]), lds(&D.eu[2 * D.NP + k]), dt,
                  D.cq, nullptr);
}

// Fused pass B (substeps >= 1): flux + friction + update + fix-negative + RK combine.
// dt is already known, so explicit updates never touch HBM.
// Riverwalls (RW): the weir branch reads the NEIGHBOUR's centroid record (sw_domain_openmp.c:635), which
// this very kernel overwrites.  Both triangles of a wall edge carry the wall flag, so it is enough that
// triangles with a wall edge do not update in here: they store their explicit updates like pass B1 and
// are updated afterwards by k_update_list over the (short) list of wall triangles; every other triangle
// reads no neighbour centroid and stays fused.
template <bool RW>
__global__ void __launch_bounds__(BLOCK, SWK_MINB_FU) k_flux_update(Dev D, Consts K, UpdateArgs U, int k0, int k1)
{
  if (D.clock->stop) return;
  const double dt = D.clock->dt;
  // (written as a loop over the CTA's single tile: with this shape ptxas keeps the edge loop at 32 bytes of
  // spills instead of 52 / 88, see profiles/r2/ptxas_table.txt)
  for (unsigned tile = blockIdx.x; tile < blockIdx.x + 1; tile++) {
  const int k = k0 + tile * BLOCK + threadIdx.x;
#if SWK_PF_AHEAD > 0
  prefetch_flux_slabs(D, prefetch_target(k0), true);
#endif
  if (k >= k1) continue;
  const i4 p = lds(&D.connB[k]);
  if (RW && (p.w & 0xE)) {                      // a triangle with a wall edge: flux only
    const Eff own = effective(D.cq[k], K);
    const TriFlux T = triangle_flux<true, SWK_FU_ROLLED>(D, K, k, p, own, false);
    D.eu[k] = T.su;
    D.eu[D.NP + k] = T.xu;
    D.eu[2 * D.NP + k] = T.yu;
    continue;
  }
  // only {h, z} of the own state live across the edge loop; the record is read again (L1) for the update
  Eff own;
  {
    const Eff e0 = effective(D.cq[k], K);
    own.h = e0.h; own.z = e0.z; own.w = e0.w;
  }
  const TriFlux T = triangle_flux<false, SWK_FU_ROLLED>(D, K, k, p, own, false);
  const d4 raw = D.cq[k];
  const Eff e = effective(raw, K);
  triangle_update(D, K, U, k, raw, e, D.zflag[k], T.su, T.xu, T.yu, dt, D.cq, nullptr);
  }
}

// update of the listed triangles from their stored explicit updates (the wall triangles of the fused pass)
__global__ void __launch_bounds__(BLOCK) k_update_list(Dev D, Consts K, UpdateArgs U, const int *list, int n)
{
  if (D.clock->stop) return;
  const int j = blockIdx.x * BLOCK + threadIdx.x;
  if (j >= n) return;
  const int k = list[j];
  const double dt = D.clock->dt;
  const d4 raw = D.cq[k];
  const Eff e = effective(raw, K);
  triangle_update(D, K, U, k, raw, e, D.zflag[k], D.eu[k], D.eu[D.NP + k], D.eu[2 * D.NP + k], dt, D.cq, nullptr);
}

// =============================================================================
// boundary_flux_sum[substep]: sum of the mass flux through edges of full cells that face a
// boundary or a ghost cell (:696-701, 765).  The flux kernels drop each such term into its
// slot (slots are in the reference's (k, i) order); one 1024-thread block adds them with a
// fixed order and tree, so the sum is reproducible run to run.  The reduction rides in the
// single-block clock kernels below.
// =============================================================================
__device__ __forceinline__ double block_sum_acct(const Dev &D)
{
  __shared__ double part[1024];
  double s = 0.0;
  for (int j = threadIdx.x; j < D.n_acct; j += 1024) s += D.acct_val[j];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  return part[0];
}

__global__ void __launch_bounds__(1024) k_boundary_flux_sum(Dev D, int substep)
{
  if (D.clock->stop) return;
  const double s = block_sum_acct(D);
  if (threadIdx.x == 0) D.clock->boundary_flux_sum[substep] = s;
}

// =============================================================================
// Clock kernels (one thread): Generic_Domain.update_timestep (generic_domain.py:
// 2349-2415) and the bookkeeping of _evolve_base (:1853-1911).
// =============================================================================
__global__ void k_begin_step(Clock *c)
{
  if (c->stop) return;
  c->step_start_time = c->time;
  c->dt_min_bits = d2u(1.0e+100);
}

__global__ void __launch_bounds__(1024) k_update_timestep(Dev D, TimeParams P, int reduce_bflux)
{
  Clock *c = D.clock;
  if (c->stop) return;
  if (reduce_bflux) {
    const double bsum = block_sum_acct(D);
    if (threadIdx.x == 0) c->boundary_flux_sum[0] = bsum;
  }
  if (threadIdx.x != 0) return;
  // compute_fluxes returned the min edge timestep (substep 0) ...
  double flux_dt = u2d(c->dt_min_bits);
  if (P.fixed_flux_timestep > 0.0) flux_dt = P.fixed_flux_timestep;
  c->flux_dt = flux_dt;
  double timestep = dmin(P.CFL * flux_dt, P.evolve_max_timestep);
  c->recorded_max_timestep = dmax(timestep, c->recorded_max_timestep);
  c->recorded_min_timestep = dmin(timestep, c->recorded_min_timestep);
  if (timestep < P.evolve_min_timestep) {
    c->smallsteps += 1;
    if (c->smallsteps > P.max_smallsteps) {
      c->smallsteps = 0;
      if (c->order == 1) { c->stop = -4; return; }      // SWK_ERR_SMALLSTEP
      c->order = 1;
    }
  } else {
    c->smallsteps = 0;
    if (c->order == 1 && P.default_order == 2) c->order = 2;
  }
  if (c->finaltime >= 0.0 && c->time + timestep > c->finaltime) timestep = c->finaltime - c->time;
  if (c->time + timestep > c->yieldtime) timestep = c->yieldtime - c->time;
  c->dt = timestep;
}

// set_relative_time between RK substeps (generic_domain.py:2011, 2093, 2132)
__global__ void k_set_substep_time(Clock *c, double fraction)
{
  if (c->stop) return;
  c->time = c->step_start_time + c->dt * fraction;
}

struct FusedRain {
  double rate, factor, full_area;   // influx = (factor*dt*rate) * sum of full-cell areas
  int on;
};

__global__ void __launch_bounds__(1024) k_finish_step(Dev D, TimeParams P, int last_substep, FusedRain R)
{
  Clock *c = D.clock;
  if (c->stop) return;
  if (last_substep >= 0) {
    const double bsum = block_sum_acct(D);
    if (threadIdx.x == 0) c->boundary_flux_sum[last_substep] = bsum;
  }
  if (threadIdx.x != 0) return;
  // boundary_flux_integral_operator.__call__ (boundary_flux_integral_operator.py:44-62)
  const double dt = c->dt;
  if (P.method == 1) c->boundary_flux_integral += dt * c->boundary_flux_sum[0];
  else if (P.method == 2) c->boundary_flux_integral += 0.5 * dt * (c->boundary_flux_sum[0] + c->boundary_flux_sum[1]);
  else c->boundary_flux_integral += 1.0 / 6.0 * dt * (c->boundary_flux_sum[0] + c->boundary_flux_sum[1] + 4.0 * c->boundary_flux_sum[2]);
  c->boundary_flux_sum[0] = c->boundary_flux_sum[1] = c->boundary_flux_sum[2] = 0.0;
  if (R.on) c->fractional_step_volume_integral += R.factor * dt * R.rate * R.full_area;   // rate_operators.py:206, 259
  c->time = c->step_start_time + c->dt;                 // :1855
  c->number_of_steps += 1;
  c->total_steps += 1;
  if (c->order == 1) c->number_of_first_order_steps += 1;
  if (P.method != 1) c->flux_dt = P.evolve_max_timestep;  // quirk (8): later substeps return evolve_max_timestep
  if (c->finaltime >= 0.0 && c->time >= c->finaltime - P.epsilon) {   // :1870-1888
    if (c->time > c->finaltime) { c->stop = -5; return; }
    c->time = c->finaltime;
    c->stop = 2;
    return;
  }
  if (c->time >= c->yieldtime) { c->stop = 1; return; }  // :1891
  if (c->step_budget > 0 && c->total_steps >= c->step_budget) { c->stop = 3; return; }
  // begin the next step
  c->step_start_time = c->time;
  c->dt_min_bits = d2u(1.0e+100);
}

// =============================================================================
// Rate_operator.__call__ (operators/rate_operators.py:149-269), device form.
// all_nonneg mirrors `num.all(rate >= 0.0)` and is decided on the host from the
// operator's rate (scalar or array); dt comes from the clock.
// =============================================================================
// rf != null: {rate, factor} live in a device table (time-dependent operator: the host refreshes the
// table every step, nothing is baked into the launch) and a scalar rate decides all_nonneg itself.
__global__ void __launch_bounds__(BLOCK) k_rate_operator(Dev D, double rate, double factor,
                                                         const double *rate_array, const int *indices,
                                                         int n, int all_nonneg, double *influx_partial,
                                                         const double *rf)
{
  if (D.clock->stop) return;
  if (rf) {
    rate = rf[0];
    factor = rf[1];
    if (!rate_array) all_nonneg = (rate >= 0.0) ? 1 : 0;
  }
  const int j = blockIdx.x * BLOCK + threadIdx.x;
  double contrib = 0.0;
  if (j < n) {
    const int k = indices ? indices[j] : j;
    const double dt = D.clock->dt;
    const double r = rate_array ? rate_array[k] : rate;
    double local_rate = factor * dt * r;
    d4 c = D.cq[k];
    if (all_nonneg) {
      c.x = c.x + local_rate;
    } else {
      const double height = c.x - c.w;
      local_rate = dmax(local_rate, -height);
      const double f = (local_rate < 0.0) ? (local_rate + height) / (height + 1.0e-10) : 1.0;
      c.x = c.x + local_rate;
      c.y = c.y * f;
      c.z = c.z * f;
    }
    D.cq[k] = c;
    if (D.connB[k].w & 1) contrib = local_rate * D.fg[2 * D.NP + k].w;
  }
  // deterministic two-stage sum: per-block partials, finished by k_rate_finish
  __shared__ double part[BLOCK];
  part[threadIdx.x] = contrib;
  __syncthreads();
  for (int o = BLOCK / 2; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) influx_partial[blockIdx.x] = part[0];
}

__global__ void __launch_bounds__(1024) k_rate_finish(Clock *c, const double *influx_partial, int nblocks)
{
  if (c->stop) return;
  __shared__ double part[1024];
  double s = 0.0;
  for (int j = threadIdx.x; j < nblocks; j += 1024) s += influx_partial[j];
  part[threadIdx.x] = s;
  __syncthreads();
  for (int o = 512; o > 0; o >>= 1) {
    if (threadIdx.x < o) part[threadIdx.x] += part[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) c->fractional_step_volume_integral += part[0];
}

// single-process ghost copy, Generic_Domain.update_ghosts (generic_domain.py:2448-2469)
__global__ void __launch_bounds__(BLOCK) k_ghost_copy(Dev D, const int *full_ids, const int *ghost_ids, int n)
{
  // no stop check: the exchange after the last step of a yield must still run; surplus
  // (speculative) exchanges after the stop are idempotent copies
  const int j = blockIdx.x * BLOCK + threadIdx.x;
  if (j >= n) return;
  const d4 s = D.cq[full_ids[j]];
  d4 t = D.cq[ghost_ids[j]];
  t.x = s.x; t.y = s.y; t.z = s.z;
  D.cq[ghost_ids[j]] = t;
}

// halo pack / unpack for the multi-GPU exchange (parallel_generic_communications.py:188-245)
__global__ void __launch_bounds__(BLOCK) k_halo_pack(Dev D, const int *ids, int n, double *buf)
{
  // no stop check: the exchange after the last step of a yield must still run; surplus
  // (speculative) exchanges after the stop are idempotent copies
  const int j = blockIdx.x * BLOCK + threadIdx.x;
  if (j >= n) return;
  const d4 s = D.cq[ids[j]];
  buf[3 * j] = s.x; buf[3 * j + 1] = s.y; buf[3 * j + 2] = s.z;
}

__global__ void __launch_bounds__(BLOCK) k_halo_unpack(Dev D, const int *ids, int n, const double *buf)
{
  // no stop check: the exchange after the last step of a yield must still run; surplus
  // (speculative) exchanges after the stop are idempotent copies
  const int j = blockIdx.x * BLOCK + threadIdx.x;
  if (j >= n) return;
  d4 t = D.cq[ids[j]];
  t.x = buf[3 * j]; t.y = buf[3 * j + 1]; t.z = buf[3 * j + 2];
  D.cq[ids[j]] = t;
}

// =============================================================================
// Host<->device marshalling kernels (caller order <-> device order and layout)
// =============================================================================
// comp: 0..3 component of a d4 record; planes: number of [NP] planes in dst
__global__ void __launch_bounds__(BLOCK) k_scatter_component(d4 *dst, int NP, int N, int planes, int comp,
                                                             const double *src, const int *new2old)
{
  const long long t = (long long)blockIdx.x * BLOCK + threadIdx.x;
  if (t >= (long long)N * planes) return;
  const int i = (int)(t / N), k = (int)(t % N);
  const double v = src[(long long)new2old[k] * planes + i];
  double *rec = reinterpret_cast<double *>(&dst[(long long)i * NP + k]);
  rec[comp] = v;
}

__global__ void __launch_bounds__(BLOCK) k_gather_component(const d4 *srcrec, int NP, int N, int planes, int comp,
                                                            double *dst, const int *new2old)
{
  const long long t = (long long)blockIdx.x * BLOCK + threadIdx.x;
  if (t >= (long long)N * planes) return;
  const int i = (int)(t / N), k = (int)(t % N);
  const double *rec = reinterpret_cast<const double *>(&srcrec[(long long)i * NP + k]);
  dst[(long long)new2old[k] * planes + i] = rec[comp];
}

// plain [planes][NP] double arrays
__global__ void __launch_bounds__(BLOCK) k_scatter_plain(double *dst, int NP, int N, int planes,
                                                         const double *src, const int *new2old)
{
  const long long t = (long long)blockIdx.x * BLOCK + threadIdx.x;
  if (t >= (long long)N * planes) return;
  const int i = (int)(t / N), k = (int)(t % N);
  dst[(long long)i * NP + k] = src[(long long)new2old[k] * planes + i];
}

__global__ void __launch_bounds__(BLOCK) k_gather_plain(const double *srcp, int NP, int N, int planes,
                                                        double *dst, const int *new2old)
{
  const long long t = (long long)blockIdx.x * BLOCK + threadIdx.x;
  if (t >= (long long)N * planes) return;
  const int i = (int)(t / N), k = (int)(t % N);
  dst[(long long)new2old[k] * planes + i] = srcp[(long long)i * NP + k];
}

// derived host views: mode 0 height_c; 1 bed_e (stage_e-height_e); 2..6 vertex values of
// stage,height,xmom,ymom,bed from the edge records (sw_domain_openmp.c:1871-1892)
__global__ void __launch_bounds__(BLOCK) k_gather_derived(Dev D, Consts K, int mode, double *dst, const int *new2old)
{
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= D.N) return;
  const long long o = new2old[k];
  if (mode == 0) {
    const Eff e = effective(D.cq[k], K);
    dst[o] = e.h;
    return;
  }
  const d4 a0 = D.eq[k], a1 = D.eq[D.NP + k], a2 = D.eq[2 * D.NP + k];
  double e0, e1, e2;
  switch (mode) {
    case 1: case 6: e0 = a0.x - a0.y; e1 = a1.x - a1.y; e2 = a2.x - a2.y; break;
    case 2: e0 = a0.x; e1 = a1.x; e2 = a2.x; break;
    case 3: e0 = a0.y; e1 = a1.y; e2 = a2.y; break;
    case 4: e0 = a0.z; e1 = a1.z; e2 = a2.z; break;
    default: e0 = a0.w; e1 = a1.w; e2 = a2.w; break;
  }
  if (mode == 1) {
    dst[3 * o] = e0; dst[3 * o + 1] = e1; dst[3 * o + 2] = e2;
  } else {
    dst[3 * o] = e1 + e2 - e0;
    dst[3 * o + 1] = e0 + e2 - e1;
    dst[3 * o + 2] = e0 + e1 - e2;
  }
}

// ---- per-call elementwise entry points (host arrays staged to plain device arrays) ----
__global__ void __launch_bounds__(BLOCK) k_manning_plain(double g, double eps, int N, int sloped, const double *x,
                                                         const double *w, const double *zv, const double *uh,
                                                         const double *vh, const double *eta,
                                                         double *xmom_update, double *ymom_update)
{
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= N) return;
  double zs = 1.0, h;
  if (sloped) {
    const double z0 = zv[3 * k], z1 = zv[3 * k + 1], z2 = zv[3 * k + 2];
    const double x0 = x[6 * k], y0 = x[6 * k + 1], x1 = x[6 * k + 2], y1 = x[6 * k + 3];
    const double x2 = x[6 * k + 4], y2 = x[6 * k + 5];
    const double det = (y2 - y0) * (x1 - x0) - (y1 - y0) * (x2 - x0);
    double zx = (y2 - y0) * (z1 - z0) - (y1 - y0) * (z2 - z0);
    zx /= det;
    double zy = (x1 - x0) * (z2 - z0) - (x2 - x0) * (z1 - z0);
    zy /= det;
    zs = sqrt(1.0 + zx * zx + zy * zy);
    h = w[k] - (z0 + z1 + z2) * (1.0 / 3.0);
  } else {
    h = w[k] - zv[k];
  }
  const double S = manning_S(g, eps, eta[k], h, uh[k], vh[k], zs, sloped != 0);
  xmom_update[k] += S * uh[k];
  ymom_update[k] += S * vh[k];
}

__global__ void __launch_bounds__(BLOCK) k_update_plain(int N, double dt, double *c, const double *eu, double *siu, int *err)
{
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= N) return;
  double x = c[k];
  if (!update_value(x, dt, eu[k], siu[k])) *err = 1;
  c[k] = x;
  siu[k] = 0.0;
}

__global__ void __launch_bounds__(BLOCK) k_saxpy_plain(int N, double a, double b, double *c, const double *bk)
{
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= N) return;
  c[k] = a * c[k] + b * bk[k];
}

}  // namespace swk

namespace swk {
// RK backup / combine on resident data (quantity.c:735-769, generic_domain.py:2167-2170)
__global__ void __launch_bounds__(BLOCK) k_backup(Dev D)
{
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= D.N) return;
  const d4 c = D.cq[k];
  D.bk[k] = c.x;
  D.bk[D.NP + k] = c.y;
  D.bk[2 * D.NP + k] = c.z;
}

__global__ void __launch_bounds__(BLOCK) k_saxpy(Dev D, double a, double b, double divide_by)
{
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= D.N) return;
  d4 c = D.cq[k];
  c.x = a * c.x + b * D.bk[k];
  c.y = a * c.y + b * D.bk[D.NP + k];
  c.z = a * c.z + b * D.bk[2 * D.NP + k];
  if (divide_by != 1.0) {
    c.x = c.x / divide_by;
    c.y = c.y / divide_by;
    c.z = c.z / divide_by;
  }
  D.cq[k] = c;
}

// _openmp_fix_negative_cells alone (sw_domain_openmp.c:2037-2056), per-call layer
__global__ void __launch_bounds__(BLOCK) k_fix_negative(Dev D)
{
  const int k = blockIdx.x * BLOCK + threadIdx.x;
  if (k >= D.N) return;
  d4 c = D.cq[k];
  if ((c.x - c.w < 0.0) & ((D.connB[k].w & 1) != 0)) {
    c.x = c.w; c.y = 0.0; c.z = 0.0;
    D.cq[k] = c;
    atomicAdd((unsigned long long *)&D.clock->negative_cells, 1ULL);
  }
}
}  // namespace swk

namespace swk {
__global__ void k_set_dt(Clock *c, double dt) { c->dt = dt; }

// boundary_flux_integral_operator.__call__ for a host-driven step
__global__ void k_bfi_update(Clock *c, TimeParams P)
{
  const double dt = c->dt;
  if (P.method == 1) c->boundary_flux_integral += dt * c->boundary_flux_sum[0];
  else if (P.method == 2) c->boundary_flux_integral += 0.5 * dt * (c->boundary_flux_sum[0] + c->boundary_flux_sum[1]);
  else c->boundary_flux_integral += 1.0 / 6.0 * dt * (c->boundary_flux_sum[0] + c->boundary_flux_sum[1] + 4.0 * c->boundary_flux_sum[2]);
  c->boundary_flux_sum[0] = c->boundary_flux_sum[1] = c->boundary_flux_sum[2] = 0.0;
}
}  // namespace swk



namespace swk {
__global__ void __launch_bounds__(BLOCK) k_gather_cells(Dev D, const int *ids, int n, double *out)
{
  const int j = blockIdx.x * BLOCK + threadIdx.x;
  if (j >= n) return;
  const d4 c = D.cq[ids[j]];
  out[4 * j] = c.x; out[4 * j + 1] = c.y; out[4 * j + 2] = c.z; out[4 * j + 3] = c.w;
}

__global__ void __launch_bounds__(BLOCK) k_scatter_cells(Dev D, const int *ids, int n, const double *in)
{
  const int j = blockIdx.x * BLOCK + threadIdx.x;
  if (j >= n) return;
  d4 c = D.cq[ids[j]];
  c.x = in[3 * j]; c.y = in[3 * j + 1]; c.z = in[3 * j + 2];
  D.cq[ids[j]] = c;
}

__global__ void __launch_bounds__(BLOCK) k_scatter_bed(Dev D, const int *ids, int n, const double *in)
{
  const int j = blockIdx.x * BLOCK + threadIdx.x;
  if (j >= n) return;
  D.cq[ids[j]].w = in[j];
}
}  // namespace swk

