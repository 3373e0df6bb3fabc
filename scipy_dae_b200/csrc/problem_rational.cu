// problem_rational.cu — kernel instantiations for the "rational" residual (Radau s = 3 only).
#include "launchers.cuh"

namespace b200dae {
ProblemEntry entry_rational() { return make_entry<Rational, false, false, false>(); }
}  // namespace b200dae
