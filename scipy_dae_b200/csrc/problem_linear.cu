// problem_linear.cu — kernel instantiations for the "linear" residual (Radau s = 3 only).
#include "launchers.cuh"

namespace b200dae {
ProblemEntry entry_linear() { return make_entry<Linear, false, false, false>(); }
}  // namespace b200dae
