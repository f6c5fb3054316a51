// Instantiation unit of the downscaling resize kernels: see resize_down.cuh.
#include "resize_down.cuh"

namespace picha_b200 {

template <> cudaError_t launch_down<true, 3>(const DownLaunch &a) { return down::launch_depth<true, 3>(a); }

}  // namespace picha_b200
