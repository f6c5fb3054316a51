// Synthetic benchmark pixels: i.i.d. uniform bytes from a counter-based hash, so that host
// (picha_b200/synthetic.py) and device regenerate identical images from (seed, image, y, word).
#include "kernels.h"

namespace picha_b200 {

namespace {

__host__ __device__ inline uint64_t mix64(uint64_t z) {   // splitmix64 finaliser
	z += 0x9E3779B97F4A7C15ull;
	z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
	z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
	return z ^ (z >> 31);
}

__global__ void __launch_bounds__(256)
synthetic_fill_kernel(DevBatch img, int row_bytes, uint64_t seed, uint64_t first_image, int aligned) {
	const int words = (row_bytes + 3) >> 2;
	const int k = blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= words) return;
	const uint64_t image = first_image + blockIdx.z;
	for (int y = blockIdx.y; y < img.height; y += gridDim.y) {
		const uint64_t ctr = (image << 40) ^ ((uint64_t)y << 20) ^ (uint64_t)k;
		const unsigned v = (unsigned)mix64(mix64(seed) ^ ctr);
		uint8_t *p = img.base + (long long)blockIdx.z * img.step + (long long)y * img.stride + 4ll * k;
		if (aligned && 4 * k + 4 <= row_bytes) {
			*reinterpret_cast<unsigned *>(p) = v;
		} else {
			for (int b = 0; b < 4 && 4 * k + b < row_bytes; ++b) p[b] = (uint8_t)(v >> (8 * b));
		}
	}
}

}  // namespace

cudaError_t launch_synthetic_fill(const DevBatch &img, int n, uint64_t seed, uint64_t first_image,
                                  cudaStream_t stream, int *launches) {
	static const int bytes[8] = {3, 4, 1, 2, 2, 4, 6, 8};
	const int row_bytes = img.width * bytes[img.pixel];
	const int words = (row_bytes + 3) / 4;
	const int aligned = (reinterpret_cast<uintptr_t>(img.base) & 3) == 0 && (img.stride & 3) == 0 && (img.step & 3) == 0;
	for (int z0 = 0; z0 < n; z0 += 65535) {
		DevBatch b = img;
		b.base += (long long)z0 * img.step;
		const int nz = n - z0 < 65535 ? n - z0 : 65535;
		dim3 grid((words + 255) / 256, img.height < 65535 ? img.height : 65535, nz);
		synthetic_fill_kernel<<<grid, 256, 0, stream>>>(b, row_bytes, seed, first_image + z0, aligned);
		*launches += 1;
	}
	return cudaGetLastError();
}

}  // namespace picha_b200
