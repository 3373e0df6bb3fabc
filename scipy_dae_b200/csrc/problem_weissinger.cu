// problem_weissinger.cu — kernel instantiations for the "weissinger" residual (Radau s = 3 only).
#include "launchers.cuh"

namespace b200dae {
ProblemEntry entry_weissinger() { return make_entry<Weissinger, false, false, false>(); }
}  // namespace b200dae
