// Instantiation unit of the upscaling resize kernels: see resize_up.cuh.
#include "resize_up.cuh"

namespace picha_b200 {

template <> cudaError_t launch_up<true, 1>(const UpLaunch &a) { return up::launch_depth<true, 1>(a); }

}  // namespace picha_b200
