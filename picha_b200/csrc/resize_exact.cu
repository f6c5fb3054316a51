// Bit-exact fused two-pass resize: the parity anchor and the any-shape path.
//
// One CTA produces a tile of tile_w x band_h output pixels of one image (32 x 8 unless the band's source
// rows would not fit shared memory -- extreme downscales get narrower, shorter tiles):
//   1. horizontal pass (src/resize.cc:105-119): for every source row the band's vertical taps
//      touch, TW horizontally-filtered pixels go to shared memory as floats;
//   2. vertical pass + pack (src/resize.cc:121-132) from shared memory.
// Sums run in the reference's order with separate multiply and add, so the result equals the
// reference's byte for byte -- including its ring-buffer aliasing, which the host resolved into
// the per-tap effective rows (tables.cc).
#include "kernels.h"
#include "pixel.cuh"
#include "pixel_convert.cuh"

namespace picha_b200 {

namespace {

constexpr int kThreads = 256;

template <int CH, bool DEEP>
__global__ void __launch_bounds__(kThreads)
resize_exact_kernel(DevBatch src, DevBatch dst, ResizeTables t, FuseArgs fuse) {
	extern __shared__ float tmp[];   // [band rows][tw][CH]
	constexpr int BPP = CH * Depth<DEEP>::bytes;

	const int x0 = blockIdx.x * t.tile_w;
	const int tw = min(t.tile_w, dst.width - x0);
	const int band = blockIdx.y;
	const int y0 = band * t.band_h;
	const int th = min(t.band_h, dst.height - y0);
	const uint8_t *simg = src.base + (int64_t)blockIdx.z * src.step;
	uint8_t *dimg = dst.base + (int64_t)blockIdx.z * dst.step;
	const int row_lo = t.band_lo[band];
	const int rows = t.band_rows[band];

	for (int i = threadIdx.x; i < rows * tw; i += kThreads) {
		const int r = i / tw, xx = i - r * tw, x = x0 + xx;
		const uint8_t *p = simg + (int64_t)(row_lo + r) * src.stride + (int64_t)t.xfirst[x] * BPP;
		const float *w = t.xw + t.xstart[x];
		const int taps = t.xcount[x];
		float acc[CH];
#pragma unroll
		for (int c = 0; c < CH; ++c) acc[c] = 0.0f;
		for (int k = 0; k < taps; ++k, p += BPP) {
			const float wk = w[k];
#pragma unroll
			for (int c = 0; c < CH; ++c) {
				float u = unpack_value<DEEP>(load_channel<DEEP>(p + c * Depth<DEEP>::bytes));
				acc[c] = __fadd_rn(acc[c], __fmul_rn(wk, u));
			}
		}
#pragma unroll
		for (int c = 0; c < CH; ++c) tmp[(r * tw + xx) * CH + c] = acc[c];
	}
	__syncthreads();

	for (int i = threadIdx.x; i < th * tw; i += kThreads) {
		const int yy = i / tw, xx = i - yy * tw, y = y0 + yy;
		const int taps = t.ycount[y], s = t.ystart[y];
		float acc[CH];
#pragma unroll
		for (int c = 0; c < CH; ++c) acc[c] = 0.0f;
		for (int k = 0; k < taps; ++k) {
			const float wk = t.yw[s + k];
			const float *v = tmp + ((t.yeff[s + k] - row_lo) * tw + xx) * CH;
#pragma unroll
			for (int c = 0; c < CH; ++c) acc[c] = __fadd_rn(acc[c], __fmul_rn(wk, v[c]));
		}
		if (fuse.dst_pixel < 0) {
			uint8_t *d = dimg + (int64_t)y * dst.stride + (int64_t)(x0 + xx) * BPP;
#pragma unroll
			for (int c = 0; c < CH; ++c) store_channel<DEEP>(d + c * Depth<DEEP>::bytes, pack_value<DEEP>(acc[c]));
		} else {
			// resize, then convert (src/colorconvert.cc:136-152 on the pixel src/resize.cc:128-131 would have stored)
			unsigned pv[CH];
#pragma unroll
			for (int c = 0; c < CH; ++c) pv[c] = pack_value<DEEP>(acc[c]);
			convert_store<CH, DEEP>(dimg + (int64_t)y * dst.stride + (int64_t)(x0 + xx) * pixel_bytes(fuse.dst_pixel), pv, fuse);
		}
	}
}

template <int CH, bool DEEP>
cudaError_t launch(const DevBatch &src, const DevBatch &dst, int n, const ResizeTables &t, const FuseArgs &fuse, cudaStream_t stream) {
	const size_t smem = (size_t)t.max_band_rows * t.tile_w * CH * sizeof(float);
	if (smem > (size_t)max_dynamic_smem()) return cudaErrorInvalidValue;
	auto kern = resize_exact_kernel<CH, DEEP>;
	static SmemGrant granted;   // (per instantiation: see grow_dynamic_smem)
	cudaError_t e = grow_dynamic_smem(reinterpret_cast<const void *>(kern), (int)smem, &granted);
	if (e != cudaSuccess) return e;
	const int bands = (dst.height + t.band_h - 1) / t.band_h;
	for (int z0 = 0; z0 < n; z0 += 65535) {   // gridDim.z limit
		const int nz = min(65535, n - z0);
		DevBatch s = src, d = dst;
		s.base += (int64_t)z0 * src.step;
		d.base += (int64_t)z0 * dst.step;
		dim3 grid((dst.width + t.tile_w - 1) / t.tile_w, bands, nz);
		kern<<<grid, kThreads, smem, stream>>>(s, d, t, fuse);
	}
	return cudaGetLastError();
}

}  // namespace

cudaError_t launch_resize_exact(const DevBatch &src, const DevBatch &dst, int n, const ResizeTables &t,
                                const FuseArgs &fuse, cudaStream_t stream, int *launches) {
	*launches += (n + 65534) / 65535;
	switch (src.pixel) {
		case 0: return launch<3, false>(src, dst, n, t, fuse, stream);
		case 1: return launch<4, false>(src, dst, n, t, fuse, stream);
		case 2: return launch<1, false>(src, dst, n, t, fuse, stream);
		case 3: return launch<2, false>(src, dst, n, t, fuse, stream);
		case 4: return launch<1, true>(src, dst, n, t, fuse, stream);
		case 5: return launch<2, true>(src, dst, n, t, fuse, stream);
		case 6: return launch<3, true>(src, dst, n, t, fuse, stream);
		case 7: return launch<4, true>(src, dst, n, t, fuse, stream);
	}
	return cudaErrorInvalidValue;
}

}  // namespace picha_b200
