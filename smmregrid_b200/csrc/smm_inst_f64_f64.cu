// smm_inst_f64_f64.cu -- kernel launchers for x = double, y = double (one translation unit per type pair so
// that the library builds in parallel; see smm_internal.h).
#include "smm_launch.cuh"

namespace smm {
SMM_DECLARE_LAUNCHERS(, double, double)
}  // namespace smm
