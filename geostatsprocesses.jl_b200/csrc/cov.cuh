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
  double param[GSP_MAX_STRUCTS];  // Matern: order nu
  double aux[GSP_MAX_STRUCTS];    // Matern: 2^(1-nu) / Gamma(nu)
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

// Modified Bessel function of the second kind K_nu(x), real order nu > 0, x > 0, double precision (CUDA has no such function;
// the reference calls SpecialFunctions.besselk).  Temme's series for x < 2, Steed's continued fraction CF2 for x >= 2, both for
// the fractional order mu = nu - round(nu) in [-1/2, 1/2], then upward recurrence K_{m+1} = (2m/x) K_m + K_{m-1}.
// 1/Gamma(1 -+ mu) enter through gam1 = (1/G(1-mu) - 1/G(1+mu)) / (2mu) and gam2 = (1/G(1-mu) + 1/G(1+mu)) / 2, evaluated from the
// Taylor series of 1/Gamma(1+x) (coefficients generated with mpmath at 50 digits; truncation error 4e-18 for |mu| <= 1/2).
GSP_DEV_NOINLINE double bessel_k(double nu, double x) {
  const double EPS = 1.0e-16;
  const int nl = (int)(nu + 0.5);
  const double mu = nu - nl, mu2 = mu * mu;
  const double xi = 1.0 / x, xi2 = 2.0 * xi;
  double rkmu, rk1;
  if (x < 2.0) {
    // odd / even Taylor coefficients of 1/Gamma(1+x)
    const double o[10] = {0.57721566490153286061,   -0.042002635034095235529, -0.042197734555544336748, 0.0072189432466630995424,
                          -0.00021524167411495097282, -0.000020134854780788238656, 1.1330272319816958824e-6, 6.1160951044814158179e-9,
                          -1.1812745704870201446e-9, 7.782263439905071254e-12};
    const double e[10] = {1.0, -0.65587807152025388108, 0.1665386113822914895, -0.0096219715278769735621, -0.0011651675918590651121,
                          0.00012805028238811618615, -1.2504934821426706573e-6, -2.0563384169776071035e-7, 5.0020076444692229301e-9,
                          1.0434267116911005105e-10};
    double so = o[9], se = e[9];
#pragma unroll
    for (int j = 8; j >= 0; --j) {
      so = so * mu2 + o[j];
      se = se * mu2 + e[j];
    }
    const double gam1 = -so, gam2 = se;
    const double gampl = gam2 - mu * gam1, gammi = gam2 + mu * gam1;  // 1/Gamma(1+mu), 1/Gamma(1-mu)
    const double x2 = 0.5 * x;
    const double pimu = 3.141592653589793238462643383279502884 * mu;
    const double fact = (fabs(pimu) < EPS) ? 1.0 : pimu / sin(pimu);
    double d = -log(x2);
    double ee = mu * d;
    const double fact2 = (fabs(ee) < EPS) ? 1.0 : sinh(ee) / ee;
    double ff = fact * (gam1 * cosh(ee) + gam2 * fact2 * d);
    double sum = ff;
    ee = exp(ee);
    double p = 0.5 * ee / gampl;
    double q = 0.5 / (ee * gammi);
    double c = 1.0;
    d = x2 * x2;
    double sum1 = p;
    for (int i = 1; i <= 500; ++i) {
      ff = (i * ff + p + q) / ((double)i * i - mu2);
      c *= d / i;
      p /= (i - mu);
      q /= (i + mu);
      const double del = c * ff;
      sum += del;
      sum1 += c * (p - i * ff);
      if (fabs(del) < fabs(sum) * EPS) break;
    }
    rkmu = sum;
    rk1 = sum1 * xi2;
  } else {
    double b = 2.0 * (1.0 + x);
    double d = 1.0 / b;
    double h = d, delh = d;
    double q1 = 0.0, q2 = 1.0;
    const double a1 = 0.25 - mu2;
    double q = a1, c = a1;
    double a = -a1;
    double s = 1.0 + q * delh;
    for (int i = 2; i <= 500; ++i) {
      a -= 2 * (i - 1);
      c = -a * c / i;
      const double qnew = (q1 - b * q2) / a;
      q1 = q2;
      q2 = qnew;
      q += c * qnew;
      b += 2.0;
      d = 1.0 / (b + a * d);
      delh = (b * d - 1.0) * delh;
      h += delh;
      const double dels = q * delh;
      s += dels;
      if (fabs(dels) < fabs(s) * EPS) break;
    }
    h = a1 * h;
    rkmu = sqrt(3.141592653589793238462643383279502884 / (2.0 * x)) * exp(-x) / s;
    rk1 = rkmu * (mu + x + 0.5 - h) * xi;
  }
  for (int i = 1; i <= nl; ++i) {
    const double t = (mu + i) * xi2 * rk1 + rkmu;
    rkmu = rk1;
    rk1 = t;
  }
  return rkmu;
}

// Matern correlation at normalised lag u (GeoStatsFunctions: delta = sqrt(2 nu) * 3 h / r; Omega = 2^(1-nu)/Gamma(nu) * delta^nu;
// rho = Omega * K_nu(delta)).  The reference shifts h by eps() to dodge the singularity at the origin; here rho(0) = 1 exactly and
// anything non-finite or above 1 from under/overflow at tiny lags clamps to 1 (difference <= 1e-15).
GSP_DEV double matern_corr(double nu, double aux, double u) {
  if (u == 0.0) return 1.0;
  const double dl = sqrt(2.0 * nu) * 3.0 * u;
  if (dl > 745.0) return 0.0;  // exp(-dl) underflows
  const double r = aux * pow(dl, nu) * bessel_k(nu, dl);
  return (r <= 1.0) ? r : 1.0;
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
    c += m.sill[s] * (m.kind[s] == GSP_MATERN ? matern_corr(m.param[s], m.aux[s], u) : corr_eval(m.kind[s], u));
  }
  return c;
}

// host helpers (api.cu)
int make_cov_dev(gsp_ctx* ctx, const gsp_cov_model* cov, int dim, int argpos, CovDev* out);
int make_dom_dev(gsp_ctx* ctx, const gsp_domain* dom, int argpos, DomDev* out);
double cov_sill(const CovDev& m);

}  // namespace gsp
