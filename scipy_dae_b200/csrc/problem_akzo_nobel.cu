// problem_akzo_nobel.cu — kernel instantiations for the "akzo_nobel" residual (Radau stage counts compiled
// besides s = 3: the template flags are s = 1, 5, 7).
#include "launchers.cuh"

namespace b200dae {
ProblemEntry entry_akzo_nobel() { return make_entry<AkzoNobel, false, true, false>(); }
}  // namespace b200dae
