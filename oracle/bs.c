/* oracle/bs.c -- TEST INFRASTRUCTURE (CPU oracle). Not part of the shipped product path.
 *
 * Black-Scholes closed form with continuous dividend yield, restating /root/reference/src/bs.rs, which is the
 * analytic oracle every reference test compares against (src/mc_simd.rs:781-907, src/mc.rs:75-121).
 *   bsf_*  : f32, same operation order as bs.rs, including its Abramowitz-Stegun 7.1.26 erf (bs.rs:3-15) and the
 *            truncated pi = 3.14159 inside pdf (bs.rs:21-24), so values agree with the reference's to f32 rounding.
 *   bsd_*  : fp64 with erfc() -- the "true" closed form used for the 3-standard-error checks.
 * Units follow bs.rs: vega and rho per 1 point (/100, bs.rs:126,139,152), theta per year (bs.rs:155-184).
 * Pinned by tests/golden/bs_closed_form.json (values from an independent numpy-float32 emulation of bs.rs,
 * tools/gen_golden.py, and the table in SURVEY.md section 4).
 */
#include <math.h>

/* ----------------------------- f32, bs.rs order of operations ----------------------------- */
static float bsf_erf(float x) {                                    /* bs.rs:3-15 */
    const float sign = copysignf(1.0f, x);                        /* f32::signum */
    const float e = fabsf(x);
    const float u = 1.0f / (1.0f + 0.3275911f * e);
    const float poly = ((((1.061405429f * u + -1.453152027f) * u + 1.421413741f) * u + -0.284496736f) * u
                        + 0.254829592f) * u;
    return sign * (1.0f - poly * expf(-e * e));
}
static float bsf_cdf(float x) { return 0.5f * (1.0f + bsf_erf(x / sqrtf(2.0f))); }   /* bs.rs:17-19 */
static float bsf_pdf01(float x) {                                                    /* bs.rs:21-24, mu=0 sigma=1 */
    return expf((-1.0f * (x - 0.0f) * (x - 0.0f)) / (2.0f * 1.0f * 1.0f)) / (1.0f * sqrtf(2.0f * 3.14159f));
}
static void bsf_d(float s, float k, float v, float r, float t, float q, float *d1, float *d2) {  /* bs.rs:26-40 */
    *d1 = (logf(s / k) + (r - q + (v * v) / 2.0f) * t) / (v * sqrtf(t));
    *d2 = *d1 - v * sqrtf(t);
}

float bsf_call_price(float s, float k, float v, float r, float t, float q) {
    float d1, d2; bsf_d(s, k, v, r, t, q, &d1, &d2);
    return s * expf(-q * t) * bsf_cdf(d1) - k * expf(-r * t) * bsf_cdf(d2);
}
float bsf_put_price(float s, float k, float v, float r, float t, float q) {
    float d1, d2; bsf_d(s, k, v, r, t, q, &d1, &d2);
    return k * expf(-r * t) * (1.0f - bsf_cdf(d2)) - s * expf(-q * t) * (1.0f - bsf_cdf(d1));
}
float bsf_gamma(float s, float k, float v, float r, float t, float q) {
    float d1, d2; bsf_d(s, k, v, r, t, q, &d1, &d2);
    return (expf(-q * t) * bsf_pdf01(d1)) / (s * v * sqrtf(t));
}
float bsf_call_delta(float s, float k, float v, float r, float t, float q) {
    float d1, d2; bsf_d(s, k, v, r, t, q, &d1, &d2);
    return expf(-q * t) * bsf_cdf(d1);
}
float bsf_put_delta(float s, float k, float v, float r, float t, float q) {
    float d1, d2; bsf_d(s, k, v, r, t, q, &d1, &d2);
    return expf(-q * t) * (bsf_cdf(d1) - 1.0f);
}
float bsf_vega(float s, float k, float v, float r, float t, float q) {
    float d1, d2; bsf_d(s, k, v, r, t, q, &d1, &d2);
    return (expf(-q * t) * bsf_pdf01(d1) * (s * sqrtf(t))) / 100.0f;
}
float bsf_call_rho(float s, float k, float v, float r, float t, float q) {
    float d1, d2; bsf_d(s, k, v, r, t, q, &d1, &d2);
    return k * t * expf(-r * t) * bsf_cdf(d2) / 100.0f;
}
float bsf_put_rho(float s, float k, float v, float r, float t, float q) {
    float d1, d2; bsf_d(s, k, v, r, t, q, &d1, &d2);
    return k * t * expf(-r * t) * (bsf_cdf(d2) - 1.0f) / 100.0f;
}
static float bsf_theta_common(float s, float v, float t, float q, float d1) {
    return (-expf(-q * t) * s * bsf_pdf01(d1) * v) / (2.0f * sqrtf(t));
}
float bsf_call_theta(float s, float k, float v, float r, float t, float q) {
    float d1, d2; bsf_d(s, k, v, r, t, q, &d1, &d2);
    const float a = r * k * expf(-r * t), b = q * s * expf(-q * t);
    return bsf_theta_common(s, v, t, q, d1) - a * bsf_cdf(d2) + b * bsf_cdf(d1);
}
float bsf_put_theta(float s, float k, float v, float r, float t, float q) {
    float d1, d2; bsf_d(s, k, v, r, t, q, &d1, &d2);
    const float a = r * k * expf(-r * t), b = q * s * expf(-q * t);
    return bsf_theta_common(s, v, t, q, d1) + a * bsf_cdf(-d2) - b * bsf_cdf(-d1);
}

/* ----------------------------- fp64, exact special functions ----------------------------- */
static double bsd_cdf(double x) { return 0.5 * erfc(-x * 0.70710678118654752440); }
static double bsd_pdf(double x) { return exp(-0.5 * x * x) * 0.39894228040143267794; }
static void bsd_d(double s, double k, double v, double r, double t, double q, double *d1, double *d2) {
    *d1 = (log(s / k) + (r - q + 0.5 * v * v) * t) / (v * sqrt(t));
    *d2 = *d1 - v * sqrt(t);
}
double bsd_call_price(double s, double k, double v, double r, double t, double q) {
    double d1, d2; bsd_d(s, k, v, r, t, q, &d1, &d2);
    return s * exp(-q * t) * bsd_cdf(d1) - k * exp(-r * t) * bsd_cdf(d2);
}
double bsd_put_price(double s, double k, double v, double r, double t, double q) {
    double d1, d2; bsd_d(s, k, v, r, t, q, &d1, &d2);
    return k * exp(-r * t) * bsd_cdf(-d2) - s * exp(-q * t) * bsd_cdf(-d1);
}
double bsd_call_delta(double s, double k, double v, double r, double t, double q) {
    double d1, d2; bsd_d(s, k, v, r, t, q, &d1, &d2);
    return exp(-q * t) * bsd_cdf(d1);
}
double bsd_put_delta(double s, double k, double v, double r, double t, double q) {
    double d1, d2; bsd_d(s, k, v, r, t, q, &d1, &d2);
    return -exp(-q * t) * bsd_cdf(-d1);
}
double bsd_gamma(double s, double k, double v, double r, double t, double q) {
    double d1, d2; bsd_d(s, k, v, r, t, q, &d1, &d2);
    return exp(-q * t) * bsd_pdf(d1) / (s * v * sqrt(t));
}
double bsd_vega(double s, double k, double v, double r, double t, double q) {
    double d1, d2; bsd_d(s, k, v, r, t, q, &d1, &d2);
    return exp(-q * t) * bsd_pdf(d1) * s * sqrt(t) / 100.0;
}
double bsd_call_rho(double s, double k, double v, double r, double t, double q) {
    double d1, d2; bsd_d(s, k, v, r, t, q, &d1, &d2);
    return k * t * exp(-r * t) * bsd_cdf(d2) / 100.0;
}
double bsd_put_rho(double s, double k, double v, double r, double t, double q) {
    double d1, d2; bsd_d(s, k, v, r, t, q, &d1, &d2);
    return -k * t * exp(-r * t) * bsd_cdf(-d2) / 100.0;
}
double bsd_call_theta(double s, double k, double v, double r, double t, double q) {
    double d1, d2; bsd_d(s, k, v, r, t, q, &d1, &d2);
    return -exp(-q * t) * s * bsd_pdf(d1) * v / (2.0 * sqrt(t)) - r * k * exp(-r * t) * bsd_cdf(d2)
           + q * s * exp(-q * t) * bsd_cdf(d1);
}
double bsd_put_theta(double s, double k, double v, double r, double t, double q) {
    double d1, d2; bsd_d(s, k, v, r, t, q, &d1, &d2);
    return -exp(-q * t) * s * bsd_pdf(d1) * v / (2.0 * sqrt(t)) + r * k * exp(-r * t) * bsd_cdf(-d2)
           - q * s * exp(-q * t) * bsd_cdf(-d1);
}
