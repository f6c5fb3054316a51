// Instantiation unit of the upscaling resize kernels: see resize_up.cuh.
#include "resize_up.cuh"

namespace picha_b200 {

cudaError_t launch_up_u8(const UpLaunch &a) { return up::launch_depth<false>(a); }

}  // namespace picha_b200
