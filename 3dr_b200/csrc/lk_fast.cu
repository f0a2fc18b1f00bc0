// Specialised LK kernels (compile-time window sizes).  Placeholder until the tuned kernels land:
// every window size currently goes through lk_generic.cu.
#include "dr3lk_internal.cuh"

namespace dr3lk {

bool launch_lk_fast(Launch& L, const LKParams& p)
{
    (void)L; (void)p;
    return false;
}

}  // namespace dr3lk
