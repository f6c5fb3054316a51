// Instantiation unit of the upscaling resize kernels: see resize_up.cuh.
#include "resize_up.cuh"

namespace picha_b200 {

cudaError_t launch_up_u16(const UpLaunch &a) { return up::launch_depth<true>(a); }

}  // namespace picha_b200
