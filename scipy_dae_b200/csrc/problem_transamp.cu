// problem_transamp.cu — kernel instantiations for the "transamp" residual (Radau stage counts compiled
// besides s = 3: the template flags are s = 1, 5, 7).
#include "launchers.cuh"

namespace b200dae {
ProblemEntry entry_transamp() { return make_entry<Transamp, false, true, false>(); }
}  // namespace b200dae
