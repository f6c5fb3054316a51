// Instantiation unit of the downscaling resize kernels: see resize_down.cuh.
#include "resize_down.cuh"

namespace picha_b200 {

template <> cudaError_t launch_down<false, 1>(const DownLaunch &a) { return down::launch_depth<false, 1>(a); }

}  // namespace picha_b200
