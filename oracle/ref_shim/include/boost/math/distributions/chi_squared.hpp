// TEST INFRASTRUCTURE: stand-in for the two Boost.Math names the reference uses (Update.cpp:27-34, StateManager.cpp:613-615):
// boost::math::chi_squared and quantile(dist, p).  Own implementation: regularised incomplete gamma (series / Lentz
// continued fraction) inverted by safeguarded Newton from the Wilson-Hilferty start.
#pragma once
#include <algorithm>
#include <cmath>

namespace boost { namespace math {

class chi_squared {
 public:
  explicit chi_squared(double dof) : dof_(dof) {}
  double degrees_of_freedom() const { return dof_; }
 private:
  double dof_;
};

namespace shim_detail {
inline double reg_gamma_p(double a, double x) {
  if (!(x > 0.0)) return 0.0;
  const double pre = std::exp(a * std::log(x) - x - std::lgamma(a));
  if (x < a + 1.0) {
    double term = 1.0 / a, sum = term;
    for (int k = 1; k < 2000; ++k) { term *= x / (a + k); sum += term; if (term < sum * 1e-17) break; }
    return pre * sum;
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, f = d;
  for (int k = 1; k < 2000; ++k) {
    const double an = -k * (k - a);
    b += 2.0;
    d = an * d + b; if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c; if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double delta = c * d;
    f *= delta;
    if (std::fabs(delta - 1.0) < 1e-16) break;
  }
  return 1.0 - pre * f;
}
}  // namespace shim_detail

inline double quantile(const chi_squared& dist, double p) {
  const double dof = dist.degrees_of_freedom(), a = 0.5 * dof;
  double zlo = -8.0, zhi = 8.0;
  for (int it = 0; it < 60; ++it) { const double zm = 0.5 * (zlo + zhi); if (0.5 * std::erfc(-zm / std::sqrt(2.0)) < p) zlo = zm; else zhi = zm; }
  const double z = 0.5 * (zlo + zhi), t = 2.0 / (9.0 * dof);
  double x = dof * std::pow(1.0 - t + z * std::sqrt(t), 3.0);
  if (!(x > 0.0)) x = a;
  double lo = 0.0, hi = std::max(2.0 * x, 4.0 * dof + 60.0);
  for (int it = 0; it < 300; ++it) {
    const double f = shim_detail::reg_gamma_p(a, 0.5 * x) - p;
    if (f > 0.0) hi = x; else lo = x;
    const double dens = 0.5 * std::exp((a - 1.0) * std::log(0.5 * x) - 0.5 * x - std::lgamma(a));
    double xn = x - f / dens;
    if (!(xn > lo && xn < hi)) xn = 0.5 * (lo + hi);
    const bool done = std::fabs(xn - x) <= 2e-15 * std::max(1.0, x);
    x = xn;
    if (done) break;
  }
  return x;
}

}}  // namespace boost::math
