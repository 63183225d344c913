// Covariance-model evaluation on the device (a1 of SURVEY §8: GeoStatsFunctions formulas,
// practical-range convention) and element-centroid lookup for CartesianGrid / point domains.
#pragma once
#include "common.h"

namespace gsp {

// passed by value as a kernel parameter
struct CovDev {
  int nstruct;
  int dim;
  int kind[GSP_MAX_STRUCTS];
  double sill[GSP_MAX_STRUCTS];
  double A[GSP_MAX_STRUCTS][9];
};

struct DomDev {
  int kind;  // 0 points, 1 grid
  int dim;
  long long nelems;
  const double* coords;  // device, dim x nelems (kind 0)
  long long dims[3];
  double origin[3];
  double spacing[3];
};

// centroid of element `e` (0-based linear, column-major) - Meshes: origin + (ijk - 1/2) * spacing
GSP_DEV void centroid(const DomDev& d, long long e, double& x, double& y, double& z) {
  x = y = z = 0.0;
  if (d.kind == 1) {
    long long i = e % d.dims[0];
    long long r = e / d.dims[0];
    long long j = r % d.dims[1];
    long long k = r / d.dims[1];
    x = d.origin[0] + ((double)i + 0.5) * d.spacing[0];
    if (d.dim > 1) y = d.origin[1] + ((double)j + 0.5) * d.spacing[1];
    if (d.dim > 2) z = d.origin[2] + ((double)k + 0.5) * d.spacing[2];
  } else {
    const double* p = d.coords + e * d.dim;
    x = p[0];
    if (d.dim > 1) y = p[1];
    if (d.dim > 2) z = p[2];
  }
}

GSP_DEV double corr_eval(int kind, double u) {
  switch (kind) {
    case GSP_NUGGET:
      return u == 0.0 ? 1.0 : 0.0;
    case GSP_SPHERICAL:
      return u < 1.0 ? 1.0 - 1.5 * u + 0.5 * (u * u * u) : 0.0;
    case GSP_EXPONENTIAL:
      return exp(-3.0 * u);
    case GSP_GAUSSIAN:
      return exp(-3.0 * (u * u));
    case GSP_CUBIC: {
      if (u >= 1.0) return 0.0;
      double u2 = u * u, u3 = u2 * u, u5 = u3 * u2, u7 = u5 * u2;
      return 1.0 - (7.0 * u2 - 8.75 * u3 + 3.5 * u5 - 0.75 * u7);
    }
    case GSP_PENTASPHERICAL: {
      if (u >= 1.0) return 0.0;
      double u2 = u * u, u3 = u2 * u, u5 = u3 * u2;
      return 1.0 - (1.875 * u - 1.25 * u3 + 0.375 * u5);
    }
    case GSP_SINEHOLE: {
      if (u == 0.0) return 1.0;
      return sinpi(u) / (3.141592653589793238462643383279502884 * u);
    }
    case GSP_CIRCULAR: {
      if (u >= 1.0) return 0.0;
      return 0.636619772367581343075535053490057448 * (acos(u) - u * sqrt(1.0 - u * u));
    }
    default:
      return 0.0;
  }
}

// C(delta) = sum_k sill_k * rho_k(|A_k delta|)
GSP_DEV double cov_eval(const CovDev& m, double dx, double dy, double dz) {
  double c = 0.0;
  for (int s = 0; s < m.nstruct; ++s) {
    const double* A = m.A[s];
    double tx = A[0] * dx + A[1] * dy + A[2] * dz;
    double ty = A[3] * dx + A[4] * dy + A[5] * dz;
    double tz = A[6] * dx + A[7] * dy + A[8] * dz;
    double u = sqrt(tx * tx + ty * ty + tz * tz);
    c += m.sill[s] * corr_eval(m.kind[s], u);
  }
  return c;
}

// host helpers (api.cu)
int make_cov_dev(gsp_ctx* ctx, const gsp_cov_model* cov, int dim, int argpos, CovDev* out);
int make_dom_dev(gsp_ctx* ctx, const gsp_domain* dom, int argpos, DomDev* out);
double cov_sill(const CovDev& m);

}  // namespace gsp
