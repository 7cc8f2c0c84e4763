// smm_inst_f32_f32.cu -- kernel launchers for x = float, y = float (one translation unit per type pair so
// that the library builds in parallel; see smm_internal.h).
#include "smm_launch.cuh"

namespace smm {
SMM_DECLARE_LAUNCHERS(, float, float)
}  // namespace smm
