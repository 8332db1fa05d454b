// swk_math.cuh - FP64 device arithmetic of the DE shallow-water path.
//
// Compiled with -fmad=false: every multiply and add is separately rounded, in
// the operation order of the reference's C code, so that results are
// bit-comparable with the reference's non-contracting build
// (anuga/shallow_water/sw_domain_openmp.c; SURVEY.md section 7 "reproducible
// arithmetic").  Double-precision division and square root are IEEE
// round-to-nearest on the device.  Explicit __fma_rn calls below are
// error-free transformations inside pow_7_3 only.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace swk {

// 32-byte record: one LDG.E.256 / STG.E.256 on sm_100a.
struct __align__(32) d4 {
  double x, y, z, w;
};

struct __align__(16) i4 {
  int x, y, z, w;
};

__device__ __forceinline__ d4 ldg4(const d4 *p) { return *p; }

// ---------------------------------------------------------------------------
// Streamed operands (static geometry, connectivity, explicit updates, backups): each byte is used
// once per pass, so they bypass L1 allocation and are marked evict-first in L2, which leaves the
// caches to the records that neighbours gather (cq in pass A, eq in pass B).
// ---------------------------------------------------------------------------
__device__ __forceinline__ d4 lds(const d4 *p)
{
  d4 r;
  asm volatile("ld.global.L1::no_allocate.L2::evict_first.v4.f64 {%0,%1,%2,%3}, [%4];"
               : "=d"(r.x), "=d"(r.y), "=d"(r.z), "=d"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ i4 lds(const i4 *p)
{
  i4 r;
  asm volatile("ld.global.cs.v4.s32 {%0,%1,%2,%3}, [%4];"
               : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
  return r;
}
__device__ __forceinline__ double lds(const double *p)
{
  double r;
  asm volatile("ld.global.cs.f64 %0, [%1];" : "=d"(r) : "l"(p));
  return r;
}
__device__ __forceinline__ void sts(double *p, double v)
{
  asm volatile("st.global.cs.f64 [%0], %1;" :: "l"(p), "d"(v) : "memory");
}

// ---------------------------------------------------------------------------
// L2 prefetch-ahead of the streamed slabs: one thread per array and block asks the L2 for the
// slab that the block SWK_PF_AHEAD positions further on will read (UBLKPF.L2, no registers, no
// smem), so the DRAM stream runs ahead of the resident warps instead of being paced by them.
// ---------------------------------------------------------------------------
#ifndef SWK_PF_AHEAD          // 555 tiles = about 3/4 of a wave of resident CTAs; 92 ... 740 measure alike
#define SWK_PF_AHEAD 555
#endif
__device__ __forceinline__ void prefetch_l1(const void *p) { asm volatile("prefetch.global.L1 [%0];" :: "l"(p)); }
__device__ __forceinline__ void prefetch_l2_bulk(const void *p, unsigned bytes)
{
  asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" :: "l"(p), "r"(bytes));
}

// ---------------------------------------------------------------------------
// h^(7/3) with the exponent the reference really uses: the double nearest to
// 7/3 (sw_domain_openmp.c:1962, 1974 `pow(h, seven_thirds)`).
// glibc's pow is accurate to ~0.52 ulp; CUDA's pow only to 2 ulp, which is the
// one place where the device arithmetic could drift from the reference.  This
// routine evaluates h^2 * cbrt(h) * (1 + delta ln h) in double-double
// arithmetic and rounds once.  The two small terms need little accuracy - ln h
// only scales delta = y - 7/3 ~ 2^-53 (1e-6 absolute is plenty) and 1/(3c^2)
// scales a Newton correction that is itself 2^-52 relative (20 bits are plenty) -
// so they come from MUFU.LG2 / MUFU.RCP64H; the result has a relative error of
// ~2^-72 before the final rounding, i.e. it is the correctly rounded value except
// for ~7e-7 of arguments; it agrees with glibc on 99.92 % of arguments and is
// 1 ulp away on the rest (2e7 random h in [1e-6, 1e4], CPU emulation).
// ---------------------------------------------------------------------------
// ln(h) to ~1e-6 absolute: exponent + MUFU.LG2 of the mantissa (it only scales the 2^-53-sized delta term)
__device__ __forceinline__ double ln_rough(double h)
{
  const int hi = __double2hiint(h);
  const int ex = ((hi >> 20) & 0x7ff) - 1023;
  const double m = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(h));
  const float lg = __log2f((float)m) + (float)ex;
  return (double)lg * 0.6931471805599453;
}
// 1/x to ~2^-20 (MUFU.RCP64H): enough for a correction term that is itself 2^-52 relative
__device__ __forceinline__ double rcp_rough(double x)
{
  double r;
  asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(r) : "d"(x));
  return r;
}

__device__ __forceinline__ double pow_7_3(double h)
{
  const double y = 7.0 / 3.0;
  const double delta = __fma_rn(y, 3.0, -7.0) / 3.0;      // y - 7/3 (3*delta is exact)
  const double c = cbrt(h);
  // Newton correction of c in double-double: r = h - c^3
  const double p = c * c;
  const double pe = __fma_rn(c, c, -p);
  const double q = p * c;
  double qe = __fma_rn(p, c, -q);
  qe = __fma_rn(pe, c, qe);
  const double r = (h - q) - qe;
  const double corr = r * rcp_rough(3.0 * p);             // |corr| <= ~2^-52 c: 20 good bits suffice
  const double chi = c + corr;
  const double clo = corr - (chi - c);
  // h^2 exactly
  const double s = h * h;
  const double se = __fma_rn(h, h, -s);
  // (s + se) * (chi + clo)
  const double P = s * chi;
  double Pe = __fma_rn(s, chi, -P);
  Pe = __fma_rn(s, clo, Pe);
  Pe = __fma_rn(se, chi, Pe);
  Pe = __fma_rn(P, delta * ln_rough(h), Pe);
  return P + Pe;
}

// ---------------------------------------------------------------------------
// min / max without the NaN plumbing of fmin / fmax (DSETP + 2 selects instead of
// DSETP + 2 selects + NaN quieting + moves).  For non-NaN operands the value is the
// same as the reference's fmin / fmax; on a +0/-0 tie either zero may be returned by
// libm as well, and no consumer distinguishes them (no division by these values).
// ---------------------------------------------------------------------------
// Written as PTX setp + selp: with a constant operand the compiler would otherwise recognise
// "(a > C) ? a : C" as fmax(a, C) and emit the NaN-aware form again (DSETP.MAX + SEL + FSEL +
// LOP3 + moves instead of DSETP + 2 FSEL).
__device__ __forceinline__ double dmax(double a, double b)
{
  double r;
  asm("{\n\t.reg .pred p;\n\tsetp.gt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b));
  return r;
}
__device__ __forceinline__ double dmin(double a, double b)
{
  double r;
  asm("{\n\t.reg .pred p;\n\tsetp.lt.f64 p, %1, %2;\n\tselp.f64 %0, %1, %2, p;\n\t}" : "=d"(r) : "d"(a), "d"(b));
  return r;
}
// max(a, 0.0) = (a > 0) ? a : +0.0 on the integer pipe: clear all bits when the sign bit is set
// (negative numbers and -0.0 give +0.0, exactly as the comparison form does).
__device__ __forceinline__ double dmax0(double a)
{
  const int hi = __double2hiint(a), lo = __double2loint(a);
  const int keep = ~(hi >> 31);
  return __hiloint2double(hi & keep, lo & keep);
}

// ---------------------------------------------------------------------------
// Scalars every kernel needs (copied into kernel parameter space).
// ---------------------------------------------------------------------------
struct Consts {
  double epsilon, g, mah;               // mah = minimum_allowed_height
  double beta_w, beta_w_dry, beta_uh, beta_uh_dry, beta_vh, beta_vh_dry;
  double evolve_max_timestep;
  int vel2;                             // extrapolate_velocity_second_order
  int low_froude;
  int protect;                          // 1: protect precedes the extrapolation (always, except the
                                        //    per-call extrapolate entry point called on its own)
  int pad;
  double sqrt_g;                        // gravity**0.5 as the host's libm gives it (Characteristic_stage_boundary)
};

// ---------------------------------------------------------------------------
// Effective centroid state: what the reference's in-place sequence
//   protect (sw_domain_openmp.c:1132-1164) -> extrapolate loop 1 (:1374-1402)
// leaves in the centroid arrays, computed on the fly from the raw record
// c = {stage, xmom, ymom, bed}.  Idempotent, so it may be re-applied to values
// that were already protected.
// ---------------------------------------------------------------------------
struct Eff {
  double w, uh, vh, z, h;     // stage, momenta (zeroed if dry), bed, height_c = max(w-z,0)
  double u, v;                // velocities (vel2) or momenta (!vel2) used by the extrapolation
  double mass_added;          // (z - w) when protect lifted the stage, else 0
};

__device__ __forceinline__ Eff effective(const d4 c, const Consts &K)
{
  Eff e;
  e.w = c.x; e.uh = c.y; e.vh = c.z; e.z = c.w;
  e.mass_added = 0.0;
  if (K.protect && e.w < e.z) {         // hc <= 0 && w < bmin  (:1141-1152)
    e.mass_added = e.z - e.w;
    e.w = e.z;
  }
  e.h = dmax0(e.w - e.z);           // :1376
  if (e.h <= K.mah) {                   // :1382-1388 (subsumes protect's xmom-only zeroing :1136-1140)
    e.uh = 0.0;
    e.vh = 0.0;
  }
  e.u = e.uh;
  e.v = e.vh;
  if (K.vel2) {                         // :1390-1401; dry cells have uh = vh = 0, so the product with
    const double inv = 1.0 / ((e.h > K.mah) ? e.h : 1.0);   // 1/1 is the reference's "untouched" value
    e.u = e.uh * inv;
    e.v = e.vh * inv;
  }
  return e;
}

// limiter, sw_domain_openmp.c:1195-1231:
//     r = 1000; r0 = 1;  for each edge i: { if (dq_i < -TINY) r0 = qmin/dq_i; else if (dq_i > TINY) r0 = qmax/dq_i;
//     r = min(r0, r); }   phi = min(r*beta, 1)
// i.e. r = min(1000, ratios of the valid edges, and 1.0 when edge 0 is not valid (the carried r0)).
// qmax >= 0 is divided by positive dq only and qmin <= 0 by negative dq only, and a correctly rounded
// quotient is monotone in its denominator, so
//     min_i fl(qmax/dq_i) = fl(qmax / max_i dq_i),   min_i fl(qmin/dq_i) = fl(qmin / min_i dq_i):
// two branch-free divisions per quantity instead of the reference's three (the sign of dq differs from
// thread to thread, so a branching form would execute every division in almost every warp anyway).
__device__ __forceinline__ void limit_gradient(double &d0, double &d1, double &d2,
                                               double qmin, double qmax, double beta)
{
  const double TINY = 1.0e-100;
  const double dhi = dmax(d0, dmax(d1, d2));
  const double dlo = dmin(d0, dmin(d1, d2));
  const bool pos = dhi > TINY, neg = dlo < -TINY;
  // Where the field is smooth nothing is limited: phi = min(r*beta, 1) is exactly 1 and d*1 = d.  That can
  // be decided without the two divisions: fl(fl(q/d)*beta) >= 1 whenever q*beta >= d*(1 + 1e-12) (the margin
  // swamps the three roundings involved), the inactive sides give r = 1000, and an invalid edge 0 caps r at
  // 1.  Taken only when every active lane of the warp agrees, so the branch costs no divergence; in all other
  // cases the full computation below runs and gives, by construction, the same bits.
  {
    const double m = 1.0 + 1.0e-12;
    const bool valid0f = (d0 < -TINY) | (d0 > TINY);
    const bool fast = (beta >= 0.001) & (valid0f | (beta >= 1.0)) &
                      (!pos | (qmax * beta >= dhi * m)) & (!neg | (qmin * beta <= dlo * m));
    if (__all_sync(__activemask(), fast)) return;
  }
  const double rp = (pos ? qmax : 1000.0) / (pos ? dhi : 1.0);
  const double rn = (neg ? qmin : 1000.0) / (neg ? dlo : 1.0);
  double r = dmin(dmin(rp, rn), 1000.0);
  const bool valid0 = (d0 < -TINY) | (d0 > TINY);
  r = valid0 ? r : dmin(r, 1.0);
  const double phi = dmin(r * beta, 1.0);
  d0 = d0 * phi;
  d1 = d1 * phi;
  d2 = d2 * phi;
}

// Static extrapolation geometry of one triangle (precomputed on the host with the
// reference's operation order, sw_domain_openmp.c:1445-1484 / 1684-1696).
struct XGeom {
  double dxv0, dxv1, dxv2, dyv0, dyv1, dyv2;   // edge midpoint - centroid
  double dx1, dx2, dy1, dy2;                   // auxiliary triangle (nb<=1) | (.,dx2,.,dy2) 1-D gradient (nb==2)
  double inv_area2;                            // 1/(dy2*dx1 - dy1*dx2)      (nb<=1)
};

// three-neighbour plane gradient + limiter, :1233-1284
__device__ __forceinline__ void edge_values_3(double beta, double qc, double q0, double q1, double q2,
                                              const XGeom &G, double &e0, double &e1, double &e2)
{
  if (beta > 0.) {
    const double dq0 = q0 - qc;
    const double dq1 = q1 - q0;
    const double dq2 = q2 - q0;
    double a = G.dy2 * dq1 - G.dy1 * dq2;
    a *= G.inv_area2;
    double b = G.dx1 * dq2 - G.dx2 * dq1;
    b *= G.inv_area2;
    double d0 = a * G.dxv0 + b * G.dyv0;
    double d1 = a * G.dxv1 + b * G.dyv1;
    double d2 = a * G.dxv2 + b * G.dyv2;
    const double qmax = dmax0(dmax(dq0, dmax(dq0 + dq1, dq0 + dq2)));
    const double qmin = dmin(dmin(dq0, dmin(dq0 + dq1, dq0 + dq2)), 0.0);
    limit_gradient(d0, d1, d2, qmin, qmax, beta);
    e0 = qc + d0;
    e1 = qc + d1;
    e2 = qc + d2;
  } else {
    e0 = e1 = e2 = qc;
  }
}

// single-neighbour gradient (triangle with two boundary edges), :1702-1840
__device__ __forceinline__ void edge_values_1(double beta, double qc, double q1, const XGeom &G,
                                              double &e0, double &e1, double &e2)
{
  const double dq1 = q1 - qc;
  const double a = dq1 * G.dx2;
  const double b = dq1 * G.dy2;
  double d0 = a * G.dxv0 + b * G.dyv0;
  double d1 = a * G.dxv1 + b * G.dyv1;
  double d2 = a * G.dxv2 + b * G.dyv2;
  double qmin, qmax;
  if (dq1 >= 0.0) { qmin = 0.0; qmax = dq1; }
  else { qmin = dq1; qmax = 0.0; }
  limit_gradient(d0, d1, d2, qmin, qmax, beta);
  e0 = qc + d0;
  e1 = qc + d1;
  e2 = qc + d2;
}

// ---------------------------------------------------------------------------
// Central-upwind (Kurganov-Noelle-Petrova) edge flux with Audusse heights,
// sw_domain_openmp.c:65-268.  Momenta are given in x/y; (n1,n2) is the outward
// unit normal.  Returns the flux in x/y, the maximal wave speed and the
// separately kept pressure flux.
// ---------------------------------------------------------------------------
struct EdgeFlux {
  double f0, f1, f2, max_speed, pressure_flux;
};

__device__ __forceinline__ EdgeFlux edge_flux_central(double wl, double uhl_xy, double vhl_xy,
                                                      double wr, double uhr_xy, double vhr_xy,
                                                      double h_left, double h_right,
                                                      double hle, double hre,
                                                      double n1, double n2, double ze,
                                                      const Consts &K)
{
  EdgeFlux F;
  if (h_left == 0. && h_right == 0.) {
    F.f0 = F.f1 = F.f2 = 0.0;
    F.max_speed = 0.0;
    F.pressure_flux = 0.0;
    return F;
  }
  double uh_left = n1 * uhl_xy + n2 * vhl_xy;
  double vh_left = -n2 * uhl_xy + n1 * vhl_xy;
  double uh_right = n1 * uhr_xy + n2 * vhr_xy;
  double vh_right = -n2 * uhr_xy + n1 * vhr_xy;
  double u_left = 0., v_left = 0., u_right = 0., v_right = 0.;
  if (hle > 0.0) {
    const double inv = 1.0 / hle;
    u_left = uh_left * inv;
    uh_left = h_left * u_left;
    v_left = vh_left * inv;
    vh_left = h_left * inv * vh_left;
  } else {
    uh_left = 0.; vh_left = 0.;
  }
  if (hre > 0.0) {
    const double inv = 1.0 / hre;
    u_right = uh_right * inv;
    uh_right = h_right * u_right;
    v_right = vh_right * inv;
    vh_right = h_right * inv * vh_right;
  } else {
    uh_right = 0.; vh_right = 0.;
  }
  const double c_left = sqrt(K.g * h_left);
  const double c_right = sqrt(K.g * h_right);

  double local_fr = 1.0;
  if (K.low_froude == 1) {
    local_fr = sqrt(dmax(0.001, dmin(1.0,
        (u_right * u_right + u_left * u_left + v_right * v_right + v_left * v_left) /
        (c_left * c_left + c_right * c_right + 1.0e-10))));
  } else if (K.low_froude == 2) {
    local_fr = sqrt((u_right * u_right + u_left * u_left + v_right * v_right + v_left * v_left) /
                    (c_left * c_left + c_right * c_right + 1.0e-10));
    local_fr = sqrt(dmin(1.0, 0.01 + dmax0(local_fr - 0.01)));
  }

  double s_max = dmax(u_left + c_left, u_right + c_right);
  if (s_max < 0.0) s_max = 0.0;
  double s_min = dmin(u_left - c_left, u_right - c_right);
  if (s_min > 0.0) s_min = 0.0;

  const double fl0 = u_left * h_left, fl1 = u_left * uh_left, fl2 = u_left * vh_left;
  const double fr0 = u_right * h_right, fr1 = u_right * uh_right, fr2 = u_right * vh_right;

  const double denom = s_max - s_min;
  if (denom < K.epsilon) {
    F.f0 = F.f1 = F.f2 = 0.0;
    F.max_speed = 0.0;
    F.pressure_flux = 0.5 * K.g * 0.5 * (h_left * h_left + h_right * h_right);
    return F;
  }
  F.max_speed = dmax(s_max, -s_min);
  const double inv_denom = 1.0 / dmax(denom, 1.0e-100);
  const double smm = s_max * s_min;
  double e0 = s_max * fl0 - s_min * fr0;
  e0 += smm * (dmax(wr, ze) - dmax(wl, ze));
  e0 *= inv_denom;
  double e1 = s_max * fl1 - s_min * fr1;
  e1 += local_fr * smm * (uh_right - uh_left);
  e1 *= inv_denom;
  double e2 = s_max * fl2 - s_min * fr2;
  e2 += local_fr * smm * (vh_right - vh_left);
  e2 *= inv_denom;
  F.pressure_flux = 0.5 * K.g * (s_max * h_left * h_left - s_min * h_right * h_right) * inv_denom;
  F.f0 = e0;
  F.f1 = n1 * e1 - n2 * e2;     // rotate back with (n1, -n2); negation is exact
  F.f2 = n2 * e1 + n1 * e2;
  return F;
}

// Villemonte weir blend on riverwall edges, sw_domain_openmp.c:324-426
__device__ __noinline__ void weir_adjust(EdgeFlux &F, double h_left, double h_right, double g,
                                         double weir_height, double Qfactor, double s1, double s2,
                                         double h1, double h2)
{
  const double twothirds = (2.0 / 3.0);
  if ((h_left <= 0.0) && (h_right <= 0.0)) return;
  const double minhd = dmin(h_left, h_right);
  const double maxhd = dmax(h_left, h_right);
  double rw = Qfactor * twothirds * maxhd * sqrt(twothirds * g * maxhd);
  const double rw2 = Qfactor * twothirds * minhd * sqrt(twothirds * g * minhd);
  const double rwRat = rw2 / dmax(rw, 1.0e-100);
  const double hdRat = minhd / dmax(maxhd, 1.0e-100);
  const double hdWrRat = minhd / dmax(weir_height, 1.0e-100);
  rw = rw * pow(1.0 - rwRat, 0.385);
  if (h_right > h_left) rw *= -1.0;
  if ((hdRat < s2) & (hdWrRat < h2)) {
    const double w1 = dmin(dmax0(hdRat - s1) / (s2 - s1), 1.0);
    const double w2 = dmin(dmax0(hdWrRat - h1) / (h2 - h1), 1.0);
    const double newFlux = (rw * (1.0 - w1) + w1 * F.f0) * (1.0 - w2) + w2 * F.f0;
    double scaleFlux;
    if (fabs(F.f0) > 1.0e-100) scaleFlux = newFlux / F.f0;
    else scaleFlux = 0.;
    scaleFlux = dmax0(scaleFlux);
    F.f0 = newFlux;
    F.f1 *= dmin(scaleFlux, 10.);
    F.f2 *= dmin(scaleFlux, 10.);
  }
  if (fabs(F.f0) > 0.)
    F.max_speed = sqrt(g * (maxhd + weir_height)) + fabs(F.f0 / (maxhd + 1.0e-12));
}

// Manning friction coefficient S (flat form), sw_domain_openmp.c:1966-1982;
// sloped form multiplies by zs = sqrt(1+zx^2+zy^2) (:2021-2026).
__device__ __forceinline__ double manning_S(double g, double eps, double eta, double h,
                                            double uh, double vh, double zs, bool sloped)
{
  double S = 0.0;
  if (eta > eps) {
    if (h >= eps) {
      const double abs_mom = sqrt((uh * uh + vh * vh));
      if (sloped) S = -g * eta * eta * zs * abs_mom;
      else S = -g * eta * eta * abs_mom;
      S /= pow_7_3(h);
    }
  }
  return S;
}

// Quantity.update for one value, quantity.c:786-812 : returns false when the
// semi-implicit denominator is <= 0.
__device__ __forceinline__ bool update_value(double &x, double dt, double explicit_update,
                                             double semi_implicit_update)
{
  const double s = (x == 0.0) ? 0.0 : semi_implicit_update / x;
  x += dt * explicit_update;
  const double denominator = 1.0 - dt * s;
  if (denominator <= 0.0) return false;
  x /= denominator;
  return true;
}

// monotone map between positive doubles and uint64 for atomicMin
__device__ __forceinline__ unsigned long long d2u(double x) { return (unsigned long long)__double_as_longlong(x); }
__device__ __forceinline__ double u2d(unsigned long long u) { return __longlong_as_double((long long)u); }

}  // namespace swk
