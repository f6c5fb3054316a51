// Launchers shared between the kernel translation units and the C-ABI layer (api.cu).
#ifndef PICHA_B200_KERNELS_H
#define PICHA_B200_KERNELS_H

#include <cuda_runtime.h>
#include <stdint.h>

namespace picha_b200 {

// A uniform batch of images in device memory: image i starts at base + i*step.
struct DevBatch {
	uint8_t *base;
	int64_t step;
	int stride, width, height, pixel;
};

// Device-resident tables of one resize plan (all pointers into one allocation).
struct ResizeTables {
	// horizontal axis, reference order
	const int *xfirst, *xcount, *xstart;
	const float *xw;
	// vertical axis: per output row a gather list (effective source row, weight)
	const int *ycount, *ystart, *yeff;
	const float *yw;
	// per band of `band_h` output rows: first source row touched and how many
	const int *band_lo, *band_rows;
	int band_h;         // output rows per CTA band
	int max_band_rows;  // max over bands of band_rows
};

// Bit-exact separable resize (reference summation order, separate mul/add). Any format,
// ratio, stride or alignment. Returns cudaErrorInvalidValue if the band does not fit in smem.
cudaError_t launch_resize_exact(const DevBatch &src, const DevBatch &dst, int n,
                                const ResizeTables &t, cudaStream_t stream, int *launches);

// Pixel-format conversion, bit-exact (integer identities for everything but luma, which is
// float without contraction). Any stride or alignment.
cudaError_t launch_color_convert(const DevBatch &src, const DevBatch &dst, int n,
                                 float rf, float gf, float bf, cudaStream_t stream, int *launches);

cudaError_t launch_synthetic_fill(const DevBatch &img, int n, uint64_t seed, uint64_t first_image,
                                  cudaStream_t stream, int *launches);

int max_dynamic_smem();

}  // namespace picha_b200
#endif
