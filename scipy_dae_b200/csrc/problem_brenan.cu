// problem_brenan.cu — kernel instantiations for the "brenan" residual (Radau s = 3 only).
#include "launchers.cuh"

namespace b200dae {
ProblemEntry entry_brenan() { return make_entry<Brenan, false, false, false>(); }
}  // namespace b200dae
