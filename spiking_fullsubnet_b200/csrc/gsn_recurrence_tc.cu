// tcgen05 / TMEM recurrence (GSN_BACKEND_TCGEN05) -- placeholder until the kernel lands.
#include "gsn_common.cuh"

namespace gsn {

bool recurrence_tc_supported(int, int, int) { return false; }
size_t recurrence_tc_workspace(int, int, int) { return 0; }
int launch_recurrence_tc(const float*, const float*, const float*, const float*, const float*,
                         const float*, const float*, float*, float*, float*, float*, int, int, int, int,
                         void*, cudaStream_t) {
  return fail(GSN_ENOSUP, "tcgen05 recurrence not built");
}

}  // namespace gsn
