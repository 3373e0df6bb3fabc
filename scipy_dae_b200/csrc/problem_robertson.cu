// problem_robertson.cu — kernel instantiations for the "robertson" residual (Radau stage counts compiled
// besides s = 3: the template flags are s = 1, 5, 7).
#include "launchers.cuh"

namespace b200dae {
ProblemEntry entry_robertson() { return make_entry<Robertson, true, true, true>(); }
}  // namespace b200dae
