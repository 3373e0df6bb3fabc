// problem_robertson_stages.cu — Radau IIA with s = 1, 5, 7 stages for the "robertson" residual
// (option matrix of the reference's Robertson tests, tests/test_integration_robertson.py:54-62).
#include "launchers.cuh"

namespace b200dae {
void entry_robertson_stages(ProblemEntry& e) {
    e.radau[1] = launch_radau<Robertson, 1>;
    e.radau[5] = launch_radau<Robertson, 5>;
    e.radau[7] = launch_radau<Robertson, 7>;
}
}  // namespace b200dae
