// problem_andrews.cu — kernel instantiations for the "andrews" residual (Radau stage counts compiled
// besides s = 3: the template flags are s = 1, 5, 7).
#include "launchers.cuh"

namespace b200dae {
ProblemEntry entry_andrews() { return make_entry<Andrews, false, true, false>(); }
}  // namespace b200dae
