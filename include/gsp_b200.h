/* gsp_b200.h - C ABI of the B200-native Gaussian-field simulation engine (libgspb200.so).
 *
 * Drop-in boundary for the LUSIM / FFTSIM hot path of JuliaEarth/GeoStatsProcesses.jl v0.13.0.
 * The reference has no FFI for this path (it is pure Julia calling OpenBLAS/FFTW through
 * LinearAlgebra/FFTW.jl); the entry points below are what a Julia method plugin
 *     struct LUSIM_B200 <: FieldSimulationMethod   (src/simulation/field.jl:135)
 *     preprocess(rng, process, ::LUSIM_B200, init, domain, data)   (called at field.jl:64,87)
 *     randsingle(rng, process, ::LUSIM_B200, domain, data, preproc) (called at field.jl:67,90)
 * `ccall`s.  INTEGRATION.md shows that glue.  Each function cites the reference lines it replaces.
 *
 * Conventions: plain pointers and sizes only; Float64 everywhere; matrices column-major (Julia
 * layout); node indices are 1-BASED like Julia's (`dinds`, `inds`); the realization index is the
 * slowest dimension of W / Z.  Host pointers unless the function name ends in `_dev`.
 * Calls are blocking; calls on one plan are serialised internally; distinct plans are re-entrant.
 * The library owns all device memory inside the opaque handles and never keeps caller pointers.
 *
 * Return value: 0 = ok; >0 = LAPACK-style `info` (1-based index, in the factored ordering
 * [data nodes; simulation nodes], of the first non-positive pivot -> glue throws PosDefException,
 * as `cholesky` does at lusim.jl:92,98,103); <0 = error, see GSP_E_*; -k for 1<=k<=32 means
 * "argument k is invalid" (the reference's ArgumentError / AssertionError cases).
 */
#ifndef GSP_B200_H
#define GSP_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSP_OK 0
#define GSP_E_CUDA (-1001)        /* CUDA runtime failure; text in gsp_last_error */
#define GSP_E_UNSUPPORTED (-1002) /* e.g. dim > 3, a grid line that does not fit shared memory (extent > ~6000 off the fast path) */
#define GSP_E_NOMEM (-1003)
#define GSP_E_STATE (-1004)       /* plan not usable (failed factorization) */

/* basic structures of GeoStatsFunctions covariances / variograms (practical-range convention);
 * a variogram gamma is passed as its covariance form sill - gamma (src/utils.jl:50-62). */
enum gsp_kind {
  GSP_NUGGET = 0,         /* c * [h == 0]                                   */
  GSP_SPHERICAL = 1,      /* c * (1 - 1.5u + 0.5u^3) [u < 1]                */
  GSP_EXPONENTIAL = 2,    /* c * exp(-3u)                                   */
  GSP_GAUSSIAN = 3,       /* c * exp(-3u^2)                                 */
  GSP_CUBIC = 4,          /* c * (1 - 7u^2 + 8.75u^3 - 3.5u^5 + 0.75u^7) [u<1] */
  GSP_PENTASPHERICAL = 5, /* c * (1 - 1.875u + 1.25u^3 - 0.375u^5) [u < 1]  */
  GSP_SINEHOLE = 6,       /* c * sin(pi u) / (pi u)  (1 at u = 0)           */
  GSP_CIRCULAR = 7,       /* c * (2/pi) (acos(u) - u sqrt(1 - u^2)) [u < 1]  */
  GSP_MATERN = 8          /* c * 2^(1-nu)/Gamma(nu) d^nu K_nu(d), d = sqrt(2 nu) 3u; nu = param > 0 (1 at u = 0) */
};

/* one nested structure: sill * rho(|A * delta|).  A is 3x3 ROW-major and maps a coordinate
 * difference into the unit-range frame: isotropic range r -> A = I/r; MetricBall(radii, R)
 * (Mahalanobis) -> A = diag(1/radii) * R'.  Unused rows/cols (dim < 3) must be zero. */
typedef struct gsp_structure {
  int32_t kind;
  int32_t reserved;
  double sill;
  double A[9];
  double param; /* GSP_MATERN: the order nu (GeoStatsFunctions' `order`, default 1.0); ignored by the other kinds */
} gsp_structure;

#define GSP_MAX_STRUCTS 8
typedef struct gsp_cov_model {
  int32_t nstruct; /* <= GSP_MAX_STRUCTS */
  int32_t reserved;
  const gsp_structure* structs;
} gsp_cov_model;

/* simulation domain: element centroids.  kind 1 = CartesianGrid (centroid = origin + (ijk-1/2)*spacing,
 * column-major linear index, x fastest: Meshes convention pinned by test/initialization.jl:16-21);
 * kind 0 = explicit centroid list (PointSet / any mesh), coords is dim x nelems column-major. */
typedef struct gsp_domain {
  int32_t kind;
  int32_t dim; /* 1..3 */
  int64_t nelems;
  const double* coords; /* kind 0 only */
  int64_t dims[3];      /* kind 1 only; unused extents = 1 */
  double origin[3];
  double spacing[3];
} gsp_domain;

typedef struct gsp_ctx gsp_ctx;
typedef struct gsp_lu_plan gsp_lu_plan;
typedef struct gsp_fft_plan gsp_fft_plan;

const char* gsp_version(void);

/* One context drives `ndev` GPUs from the calling process (devs = CUDA ordinals; NULL = 0..ndev-1).
 * Replaces the reference's worker pool (src/simulation/field.jl:93-121): realizations are
 * sharded over the devices instead of over Julia worker processes. */
int gsp_ctx_create(int32_t ndev, const int32_t* devs, gsp_ctx** out);
int gsp_ctx_destroy(gsp_ctx* ctx);
const char* gsp_last_error(gsp_ctx* ctx);
int gsp_ctx_ndev(gsp_ctx* ctx);

/* pinned host buffers for W / Z (full PCIe speed); plain Julia Arrays work too, only slower */
int gsp_host_alloc(void** out, int64_t bytes);
int gsp_host_free(void* p);

/* _pairwise(fun, dom1[, dom2]) - src/utils.jl:50-62 (-> GeoStatsFunctions.pairwise).
 * X1: dim x n1, X2: dim x n2 (NULL => X1, symmetric).  out: n1 x n2 column-major. */
int gsp_pairwise(gsp_ctx* ctx, const gsp_cov_model* cov, int32_t dim, int64_t n1, const double* X1, int64_t n2,
                 const double* X2, double* out);

/* cholesky(Symmetric(A)).L - src/simulation/field/lusim.jl:92,98,103.  A: n x n column-major, only the
 * lower triangle is read; on return the lower triangle holds L and the strict upper triangle is zero. */
int gsp_potrf(gsp_ctx* ctx, int64_t n, double* A);

/* initialize + NearestInit for ONE variable on a CartesianGrid - src/processes/field.jl:43-58,
 * src/initialization/nearest.jl:12-34: every datum goes to the element whose centroid is nearest (grid arithmetic on the device
 * instead of a KD-tree over all nelems centroids), later data overwrite earlier ones, a NaN value (= missing) is skipped.
 * grid: kind 1.  dcoords: dim x nd column-major data locations, dvals: nd values.
 * Out (caller-allocated, capacity nd): dinds = findall(mask) - 1-based, ascending - and z1 = the values in that order;
 * *count = number of distinct data nodes.  These are the (nd, dinds, z1) of gsp_lu_plan_create and the knodes of
 * gsp_fft_plan_condition.  Views and non-grid domains keep the host-side search of the glue. */
int gsp_nearest_init(gsp_ctx* ctx, const gsp_domain* grid, int64_t nd, const double* dcoords, const double* dvals, int64_t* dinds,
                     double* z1, int64_t* count);

/* LUSIM preprocess for ONE variable - the body of the map at src/simulation/field/lusim.jl:66-107:
 * assembles C11/C12/C22 over the centroids (lusim.jl:81-96), factorises (lusim.jl:92 or 98-103) and
 * computes d2 (lusim.jl:102).  `cov` is the marginal covariance of the variable (lusim.jl:85).
 * nd, dinds (1-based, strictly ascending = findall(mask), lusim.jl:71), z1 (lusim.jl:75), mu (lusim.jl:78).
 * nd == 0 => unconditional.  Returns >0 (info) if the matrix is not positive definite. */
int gsp_lu_plan_create(gsp_ctx* ctx, const gsp_cov_model* cov, const gsp_domain* dom, int64_t nd, const int64_t* dinds,
                       const double* z1, double mu, gsp_lu_plan** out);
/* Plan of ANOTHER variable whose marginal covariance and data nodes are those of `base` - cosimulation with a proportional model
 * such as [1 rho; rho 1] * cov, where the `map` at lusim.jl:66-107 assembles and factors the same matrix once per variable: the new
 * plan SHARES base's factor (reference counted; either plan may be destroyed first) and only computes its own d2 from z1 (lusim.jl:102).
 * nd / dinds must equal base's.  The caller decides that the marginal covariances are equal (the glue compares the flattened structures). */
int gsp_lu_plan_create_like(gsp_lu_plan* base, int64_t nd, const int64_t* dinds, const double* z1, double mu, gsp_lu_plan** out);
int gsp_lu_plan_destroy(gsp_lu_plan* plan);
/* sizes[0] = N, sizes[1] = Nd, sizes[2] = Ns */
int gsp_lu_plan_sizes(gsp_lu_plan* plan, int64_t sizes[3]);
/* device time (ms, CUDA events) of the plan's stages: ms[0] covariance assembly, ms[1] Cholesky, ms[2] d2 solve */
int gsp_lu_plan_times(gsp_lu_plan* plan, double ms[3]);
/* inspection (tests / debugging): d2 (Ns) and L22 (Ns x Ns column-major, lower) of lusim.jl:106; either may be NULL */
int gsp_lu_plan_get(gsp_lu_plan* plan, double* d2, double* L22);

/* _lusim for R realizations at once - src/simulation/field/lusim.jl:145-175 (and randsingle :112-126):
 *   Z[sinds, r] = d2 + L22 * w_r                      (rho is NaN: first / only variable, lusim.jl:162)
 *   Z[sinds, r] = d2 + L22 * (rho*w1_r + sqrt(1-rho^2)*w_r)   (second variable, lusim.jl:164)
 *   Z[dinds, r] = z1 (lusim.jl:168);  Z += mu only when nd == 0 (lusim.jl:172).
 * W, W1: Ns x R column-major standard normals (the reference's randn at lusim.jl:160).  W == NULL =>
 * on-device counter RNG: column r of variable stream s uses Philox(seed, stream, first_real + r); pass
 * `stream` = 0 for variable 1 and 1 for variable 2 (then W1 == NULL regenerates stream 0 for w1).
 * Z: N x R column-major. */
int gsp_lu_sample(gsp_lu_plan* plan, int64_t R, const double* W, uint64_t seed, int32_t stream, int64_t first_real,
                  double rho, const double* W1, double* Z);
/* same with DEVICE pointers on device 0 of the context (ld = leading dimension of W/W1 and Z) */
int gsp_lu_sample_dev(gsp_lu_plan* plan, int64_t R, const double* W, int64_t ldw, uint64_t seed, int32_t stream,
                      int64_t first_real, double rho, const double* W1, double* Z, int64_t ldz);

/* FFTSIM preprocess, unconditional part - src/simulation/field/fftsim.jl:77-91: covariance from the
 * centre centroid (dims .÷ 2) to every node, F = sqrt.(abs.(fft(fftshift(C)))), F[1] = 0.
 * grid: dom->kind must be 1 (the parent grid). */
int gsp_fft_plan_create(gsp_ctx* ctx, const gsp_cov_model* cov, const gsp_domain* grid, gsp_fft_plan** out);
int gsp_fft_plan_destroy(gsp_fft_plan* plan);
/* inspection: full F (prod(dims) doubles, column-major) as the reference stores it */
int gsp_fft_plan_get(gsp_fft_plan* plan, double* F);

/* FFTSIM randsingle, unconditional part, for R realizations - src/simulation/field/fftsim.jl:124-135:
 *   P = F .* exp.(im .* angle.(fft(w)));  Z = real(ifft(P));  s2 = var(Z, mean=0);
 *   Z = sqrt(sill/s2) .* Z .+ mu;  out = Z[inds]
 * w: prod(dims) x R uniforms in [0,1) (rand at fftsim.jl:124); NULL => on-device Philox(seed, 0, first_real + r).
 * inds: n_inds 1-based parent indices (parentindices(sdom), fftsim.jl:120) or NULL (n_inds = 0) for the whole grid.
 * out: (n_inds or prod(dims)) x R column-major. */
int gsp_fft_sample(gsp_fft_plan* plan, int64_t R, const double* w, uint64_t seed, int64_t first_real, double sill,
                   double mu, int64_t n_inds, const int64_t* inds, double* out);
/* same with DEVICE pointers on device 0 of the context; inds_dev 1-based device array or NULL */
int gsp_fft_sample_dev(gsp_fft_plan* plan, int64_t R, const double* w, uint64_t seed, int64_t first_real, double sill,
                       double mu, int64_t n_inds, const int64_t* inds_dev, double* out);

/* FFTSIM conditioning by simple Kriging of residuals - src/simulation/field/fftsim.jl:94-101 (preprocess:
 * zbar = fitpredict(Kriging(f, mu), data, sdom; minneighbors, maxneighbors, distance)) and :140-153 (randsingle:
 * z = zbar + (zu - zbaru), zbaru = the same Kriging of the unconditional values at the data nodes view(sdom, dinds)).
 * GeoStatsModels' neighbourhood search is restated: the `maxneighbors` (default 26) nearest samples by Euclidean distance,
 * ties -> lower sample index; simple Kriging lambda = C^-1 c0 of those samples.  Both weight sets depend on geometry only
 * and are computed here once; afterwards every gsp_fft_sample* call on the plan returns CONDITIONAL realizations.
 *   mu                 the process mean (must be passed again, unchanged, to gsp_fft_sample*)
 *   nd, dcoords, dvals the conditioning table: locations as given (dim x nd column-major) and values (fftsim.jl:99)
 *   nk, knodes         dinds = findall(mask) (fftsim.jl:104): 1-based, ascending positions WITHIN the simulation domain
 *   n_inds, inds       the simulation domain as a view of the grid (parentindices, NULL: whole grid); the same n_inds / inds
 *                      must be passed to gsp_fft_sample*
 * maxneighbors outside [1, #samples] means "all samples" (as fitpredict fixes it); at most 32 after that clamp.
 * Returns GSP_E_STATE if a Kriging matrix is not positive definite (the reference's cholesky would throw). */
int gsp_fft_plan_condition(gsp_fft_plan* plan, double mu, int32_t minneighbors, int32_t maxneighbors, int64_t nd, const double* dcoords,
                           const double* dvals, int64_t nk, const int64_t* knodes, int64_t n_inds, const int64_t* inds);
/* inspection: zbar (n = n_inds or prod(dims) doubles), the posterior mean of src/expectation/field/gaussian.jl:21-25 */
int gsp_fft_plan_condmean(gsp_fft_plan* plan, double* zbar);

/* ---- Device-resident ensembles: Ensemble(domain, reals; fetch) - src/ensembles.jl:10-16.
 * R realizations of n values each stay in HBM, sharded contiguously over the devices of the context (the same rule
 * gsp_*_sample uses); `fetch` is the reference's lazy per-realization hook (ensembles.jl:27-31) and the statistics
 * replace the scalar `ereduce` loops of ensembles.jl:42-52,76-85: only n-vectors cross PCIe. */
typedef struct gsp_ensemble gsp_ensemble;
int gsp_ensemble_create(gsp_ctx* ctx, int64_t n, int64_t R, gsp_ensemble** out);
int gsp_ensemble_destroy(gsp_ensemble* e);
/* sizes[0] = n (values per realization), sizes[1] = R */
int gsp_ensemble_sizes(gsp_ensemble* e, int64_t sizes[2]);
/* realizations [r0, r0 + nr) (0-based) <-> host buffer n x nr column-major */
int gsp_ensemble_put(gsp_ensemble* e, int64_t r0, int64_t nr, const double* Z);
int gsp_ensemble_fetch(gsp_ensemble* e, int64_t r0, int64_t nr, double* Z);
/* fill by simulation: the arguments of gsp_fft_sample / gsp_lu_sample (w / W host noise or NULL for the device RNG) with R and
 * the output taken from the ensemble (n must equal n_inds or prod(dims), resp. N) */
int gsp_fft_sample_ensemble(gsp_fft_plan* plan, gsp_ensemble* e, const double* w, uint64_t seed, int64_t first_real, double sill,
                            double mu, int64_t n_inds, const int64_t* inds);
int gsp_lu_sample_ensemble(gsp_lu_plan* plan, gsp_ensemble* e, const double* W, uint64_t seed, int32_t stream, int64_t first_real,
                           double rho, const double* W1);
/* mean(e), var(e) (corrected, R - 1), cdf(e, x) = count(<= x)/R, ccdf(e, x) = count(> x)/R - ensembles.jl:42-48; out: n */
int gsp_ensemble_mean(gsp_ensemble* e, double* out);
int gsp_ensemble_var(gsp_ensemble* e, double* out);
int gsp_ensemble_cdf(gsp_ensemble* e, double x, double* out);
int gsp_ensemble_ccdf(gsp_ensemble* e, double x, double* out);
/* quantile(e, ps) - ensembles.jl:50-52, Statistics.quantile's default definition (alpha = beta = 1); out: n x np column-major */
int gsp_ensemble_quantile(gsp_ensemble* e, int64_t np, const double* ps, double* out);
/* per-node mean and centred sum of squares m2 = sum (z - mean)^2 over this context's R realizations (either may be NULL):
 * the partials a multi-process job merges (Chan's update) after an all-gather over its ranks */
int gsp_ensemble_moments(gsp_ensemble* e, double* mean, double* m2);

/* Optional per-kernel-class device timing (CUDA events on the launching stream around every launch of
 * this library).  Off by default.  gsp_profile_read writes a JSON object {"kernel": {"ms": total, "launches": n}, ...}
 * accumulated since the last enable into buf and returns its length (or -needed if buflen is too small). */
int gsp_profile_enable(gsp_ctx* ctx, int32_t on);
int64_t gsp_profile_read(gsp_ctx* ctx, char* buf, int64_t buflen);

/* counters for harnesses: kernels launched by this library in this process since load */
int64_t gsp_kernel_launches(void);
/* device-time (ms, CUDA events on the library's stream) of the last gsp_*_sample* call on device 0 */
double gsp_last_sample_ms(gsp_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif
