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
	int tile_w;         // output columns per CTA
	int max_band_rows;  // max over bands of band_rows
};

// "Resize, then convert" in one kernel (picha_b200_resize_convert): when dst_pixel >= 0 the resize kernels put every
// resized pixel through the reference's format conversion in their pack stage (pixel_convert.cuh) and store it in
// the destination's format; DevBatch::pixel of the destination is then that format.
struct FuseArgs {
	int dst_pixel;      // PixelMode of the destination, or -1: no conversion (destination format = source format)
	float r, g, b;      // normalised luma weights (ColorSettings)
};

// Bit-exact separable resize (reference summation order, separate mul/add). Any format,
// ratio, stride or alignment. Returns cudaErrorInvalidValue if the band does not fit in smem.
cudaError_t launch_resize_exact(const DevBatch &src, const DevBatch &dst, int n,
                                const ResizeTables &t, const FuseArgs &fuse, cudaStream_t stream, int *launches);

// Pixel-format conversion, bit-exact (integer identities for everything but luma, which is
// float without contraction). Any stride or alignment.
cudaError_t launch_color_convert(const DevBatch &src, const DevBatch &dst, int n,
                                 float rf, float gf, float bf, cudaStream_t stream, int *launches);

// The JPEG decoder's CMYK -> RGB row loop (src/jpegcodec.cc:36-42): rgb[c] = cmyk[c] * cmyk[3] / 255.
// src: 4-byte pixels (pixel == rgba as the container), dst: rgb.
cudaError_t launch_cmyk_to_rgb(const DevBatch &src, const DevBatch &dst, int n, cudaStream_t stream, int *launches);

// Device-resident horizontal tables of the fast resize path (tables.h: FastAxisX); the vertical
// tables (FastAxisY) stay on the host and travel as kernel parameters, a slice per launch.
struct FastTables {
	const int *xfirst, *xcount;
	const float *xw;
	int xstride;
	int xtaps;    // most taps of any column
	const int *xrow;     // [dst] the column's row in xuw
	const float *xuw;    // [xunique][xstride] distinct weight rows
	int xunique;
	// flat form for 1 ([0]) and 3 ([1]) channels (tables.h: FlatRows)
	const int *xe_col[2], *xe_src[2], *xe_off[2];
	int xe_count[2];
	// host copies of xfirst / xcount / xw (launch planning; never dereferenced on the device)
	const int *h_xfirst, *h_xcount, *h_xrow;
	const float *h_xw;
	// wide-window variant of the upscaling kernel (tables.h: WideBlocks; weights already carry the kernel's 2^kHExp)
	const float *xwide;
	int xwide_window;   // 0: none
	int xshort;   // 4 or 8 when no column has more taps than that (unrolled horizontal pass), else 0
	int depth;    // vertical accumulators / window rows the kernel is instantiated with
	int tile_w;   // output columns per CTA
	int tile_w96, tile_w128;   // the same for the downscaling kernel's 96- and 128-thread variants (0: none)
	int align_px; // tile source origins are multiples of this many pixels (16-byte TMA start)
	int band_h;   // output rows per CTA
};

constexpr int kFastThreads = 128;        // threads per CTA = source column groups per tile
constexpr int kFastValuesPerThread = 8;  // channel values of one source row owned by a thread
constexpr int kFastMaxDepth = 12;

// Tile width (output columns, a multiple of `unit`) whose source span, counted from the tile origin
// (the first tap's pixel rounded down to a multiple of align_px), fits one CTA row for every tile; 0 if
// none does.
// row_values: channel values of a source row one CTA holds (1024; 1536 / 2048 for the wide downscaling variants).
int fast_tile_width(const int *xfirst, const int *xcount, int dst_w, int channels, int unit, int align_px, int cap,
                    int row_values = 1024);

struct FastAxisY;

// Fused two-pass resize for throughput: TMA-staged source rows, vertical pass in registers with
// thread-private columns and uniform (constant-bank) weights, horizontal pass from shared memory,
// coalesced stores. FMA arithmetic, vertical-then-horizontal order: within +-1 LSB of the
// reference, not bit-exact. Needs 16-byte aligned source base / stride / step (TMA). Returns
// cudaErrorNotSupported when the shape or alignment is outside what it handles (the caller then
// uses the exact kernel). One launch per group of row bands whose vertical tables fit the
// parameter block.
cudaError_t launch_resize_fast(const DevBatch &src, const DevBatch &dst, int n, const FastTables &t,
                               const FastAxisY &fy, const FuseArgs &fuse, cudaStream_t stream, int *launches);
// Frees the TMA descriptor ring launch_resize_fast keeps on `device` (current and idle); the next call makes a new one.
void release_resize_descriptors(int device);

cudaError_t launch_synthetic_fill(const DevBatch &img, int n, uint64_t seed, uint64_t first_image,
                                  cudaStream_t stream, int *launches);

int max_dynamic_smem();

// cudaFuncAttributeMaxDynamicSharedMemorySize is per function and device, and callers on several threads launch
// the same kernel with different tile sizes: the attribute only ever grows (a launch never finds it below its own
// need), under a lock, and stays at what was actually asked for.
// `granted`: the caller's per-kernel state (a function-local static of the launch template), one slot per device.
struct SmemGrant {
	int bytes[16] = {0};
};
cudaError_t grow_dynamic_smem(const void *kernel, int bytes, SmemGrant *granted);
int sm_count();   // SMs of the current device (B200: 148)

// picha_b200_last_resize_kernel(): set by the launchers, read by the C-ABI layer
extern thread_local int g_last_resize_kernel;

}  // namespace picha_b200
#endif
