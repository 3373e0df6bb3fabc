// variant_stubs.cu — empty registry entries for every problem except Robertson (s = 3), so that
// an A/B variant library (scripts/build_variant.py) is 3 MB instead of 33 MB.  Experiments only.
#include "../scipy_dae_b200/csrc/launchers.cuh"

namespace b200dae {
#define B200DAE_STUB(id)                                              \
    ProblemEntry entry_##id() {                                       \
        ProblemEntry e;                                               \
        memset(&e, 0, sizeof(e));                                     \
        e.name = #id "_stub";                                         \
        return e;                                                     \
    }
#ifdef B200DAE_STUB_STAGES
void entry_robertson_stages(ProblemEntry&) {}
#else
B200DAE_STUB(akzo_nobel)
B200DAE_STUB(transamp)
B200DAE_STUB(andrews)
B200DAE_STUB(rational)
B200DAE_STUB(linear)
B200DAE_STUB(cosine)
B200DAE_STUB(weissinger)
B200DAE_STUB(brenan)
B200DAE_STUB(transamp1)
B200DAE_STUB(blowup)
#endif
}  // namespace b200dae
