// smm_inst_f64_f32.cu -- kernel launchers for x = double, y = float (one translation unit per type pair so
// that the library builds in parallel; see smm_internal.h).
#include "smm_launch.cuh"

namespace smm {
SMM_DECLARE_LAUNCHERS(, double, float)
}  // namespace smm
