// Instantiation unit of the upscaling resize kernels: see resize_up.cuh.
#define PICHA_UP_PACKED 0   // see resize_up.cuh
#include "resize_up.cuh"

namespace picha_b200 {

template <> cudaError_t launch_up<true, 3>(const UpLaunch &a) { return up::launch_depth<true, 3>(a); }

}  // namespace picha_b200
