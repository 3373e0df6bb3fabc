// problem_transamp1.cu — kernel instantiations for the "transamp1" residual, the reference's own n = 5
// circuit (examples/daes/stiff_transistor_amplifier.py): Radau s = 3 and 5, BDF.
#include "launchers.cuh"

namespace b200dae {
ProblemEntry entry_transamp1() { return make_entry<Transamp1, false, true, false>(); }
}  // namespace b200dae
