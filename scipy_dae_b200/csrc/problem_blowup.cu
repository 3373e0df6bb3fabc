// problem_blowup.cu — kernel instantiations for the "blowup" residual (Radau s = 3 only).
#include "launchers.cuh"

namespace b200dae {
ProblemEntry entry_blowup() { return make_entry<Blowup, false, false, false>(); }
}  // namespace b200dae
