// Stand-in for lib/dfe-snippets NumericUtils.hpp (empty submodule): almost_equal(got, exp, rel, abs) as
// called at test/test_utils.hpp:36 with (1E-8, 1E-11).  Source unavailable; argument meaning inferred.
#pragma once
#include <cmath>
namespace dfesnippets { namespace numeric_utils {
inline bool almost_equal(double a, double b, double rel, double abs_tol) {
  const double diff = std::fabs(a - b);
  return diff <= abs_tol || diff <= rel * std::fmax(std::fabs(a), std::fabs(b));
}
}}
