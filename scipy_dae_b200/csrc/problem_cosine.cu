// problem_cosine.cu — kernel instantiations for the "cosine" residual (Radau s = 3 only).
#include "launchers.cuh"

namespace b200dae {
ProblemEntry entry_cosine() { return make_entry<Cosine, false, false, false>(); }
}  // namespace b200dae
