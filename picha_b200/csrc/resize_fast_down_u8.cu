// Instantiation unit of the fast resize kernels: see resize_fast.cuh / resize_fast.cu.
#include "resize_fast.cuh"

namespace picha_b200 {

cudaError_t launch_fast_down_u8(const FastLaunch &a) { return fast::launch_depth<0, false>(a); }

}  // namespace picha_b200
