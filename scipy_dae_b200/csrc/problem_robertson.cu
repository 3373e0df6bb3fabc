// problem_robertson.cu — kernel instantiations for the "robertson" residual: Radau s = 3, BDF,
// residual check, consistent initial conditions.  The other stage counts (s = 1, 5, 7) live in
// problem_robertson_stages.cu so the two halves compile in parallel.
#include "launchers.cuh"

namespace b200dae {
void entry_robertson_stages(ProblemEntry& e);
ProblemEntry entry_robertson() {
    ProblemEntry e = make_entry<Robertson, false, false, false>();
    entry_robertson_stages(e);
    return e;
}
}  // namespace b200dae
