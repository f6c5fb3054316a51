/*
 * picha_b200.h -- C-ABI of the B200-native pixel hot path for jhs67/picha.
 *
 * This is the drop-in boundary: the two plain-C++ seams of the reference,
 *
 *     void resizeImage(const ResizeOptions&, NativeImage& src, NativeImage& dst);   src/resize.cc:270
 *     void doColorConvert(const ColorSettings&, NativeImage& src, NativeImage& dst); src/colorconvert.cc:171
 *
 * re-expressed with plain pointers and sizes so picha's Node addon (src/picha.cc) can call
 * CUDA through it without knowing about CUDA.  The codecs, the JS facade and the option
 * parsing stay where they are; INTEGRATION.md shows the few lines of addon glue.
 *
 * Every function is thread-safe and re-entrant (picha.resize / picha.colorConvert run on
 * libuv pool threads, several at once: src/resize.cc:291-294,362-364).  Host entry points
 * block until the result is in `dst`.  Nothing here throws; a negative status is returned
 * and the addon turns it into a thrown Error (sync) or cb(err) (async).
 *
 * There is NO CPU fallback: without a CUDA device the compute entry points return
 * PICHA_B200_ERR_NO_DEVICE.
 */
#ifndef PICHA_B200_H
#define PICHA_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PICHA_B200_VERSION 100

/* enum PixelMode, src/picha.h:79-92 -- numeric values are part of the contract. */
enum picha_b200_pixel {
	PICHA_B200_INVALID_PIXEL = -1,
	PICHA_B200_RGB = 0,            /* 3 bytes, 3 channels u8  (src/picha.h:118-123) */
	PICHA_B200_RGBA = 1,           /* 4 bytes, 4 channels u8  (:125-130) */
	PICHA_B200_GREY = 2,           /* 1 byte,  1 channel  u8  (:132-137) */
	PICHA_B200_GREYA = 3,          /* 2 bytes, 2 channels u8  (:139-144) */
	PICHA_B200_R16 = 4,            /* 2 bytes, 1 channel  u16 (:146-151) */
	PICHA_B200_R16G16 = 5,         /* 4 bytes, 2 channels u16 (:153-158) */
	PICHA_B200_R16G16B16 = 6,      /* 6 bytes, 3 channels u16 (:160-165) */
	PICHA_B200_R16G16B16A16 = 7,   /* 8 bytes, 4 channels u16 (:167-172) */
	PICHA_B200_NUM_PIXELS = 8
};

/* enum ResizeFilterTag, src/resize.cc:151-160. */
enum picha_b200_filter {
	PICHA_B200_CUBIC = 0,
	PICHA_B200_LANCZOS = 1,
	PICHA_B200_CATMULROM = 2,
	PICHA_B200_MITCHEL = 3,
	PICHA_B200_BOX = 4,
	PICHA_B200_TRIANGLE = 5,
	PICHA_B200_NUM_FILTERS = 6
};

/* struct NativeImage, src/picha.h:202-218: interleaved, row-major, top-down, native-endian
 * u16; row(y) = data + y*stride; stride >= width*bytes; `data` may be unaligned (subView,
 * lib/image.js:76-87). */
typedef struct picha_b200_image {
	void *data;
	int32_t stride;
	int32_t width;
	int32_t height;
	int32_t pixel;   /* enum picha_b200_pixel */
} picha_b200_image;

enum picha_b200_status {
	PICHA_B200_OK = 0,
	PICHA_B200_ERR_INVALID_IMAGE = -1,        /* "invalid image"          src/resize.cc:337 */
	PICHA_B200_ERR_INVALID_DIMENSIONS = -2,   /* "invalid dimensions"     src/resize.cc:344 */
	PICHA_B200_ERR_INVALID_FILTER = -3,       /* "invalid filter mode"    src/resize.cc:185 */
	PICHA_B200_ERR_INVALID_FILTER_WIDTH = -4, /* "invalid filter width"   src/resize.cc:193 */
	PICHA_B200_ERR_INVALID_PIXEL = -5,        /* "expected pixel mode"    src/colorconvert.cc:237 */
	PICHA_B200_ERR_FORMAT_MISMATCH = -6,      /* resize needs src.pixel == dst.pixel   src/resize.cc:137 */
	PICHA_B200_ERR_SIZE_MISMATCH = -7,        /* convert needs equal width/height      src/colorconvert.cc:138-139 */
	PICHA_B200_ERR_NO_DEVICE = -8,            /* no CUDA device/driver; there is no CPU fallback */
	PICHA_B200_ERR_CUDA = -9,                 /* see picha_b200_last_error() */
	PICHA_B200_ERR_NOMEM = -10,
	PICHA_B200_ERR_UNSUPPORTED = -11,         /* shape outside the documented limits */
	PICHA_B200_ERR_INVALID_ARGUMENT = -12
};

/* Flags for the *_ex / *_device entry points. */
#define PICHA_B200_EXACT 1u       /* resize: force the bit-exact kernel (reference summation order, no FMA) */
#define PICHA_B200_FORCE_FAST 2u  /* resize: use the throughput kernel whenever it supports the shape, even for
                                     images so small that the default picks the bit-exact kernel (tests) */

/* ---- library ------------------------------------------------------------------------- */

int picha_b200_version(void);
/* Number of usable CUDA devices (0 when there is no driver or no GPU). */
int picha_b200_device_count(void);
/* Select the device host entry points use on this process (default 0) and create its
 * context, streams and staging.  device = -1 warms every device.  Optional. */
int picha_b200_init(int device);
/* Release every stream, staging buffer, cached table and device allocation. */
void picha_b200_shutdown(void);
const char *picha_b200_strerror(int status);
/* Detail of the last PICHA_B200_ERR_CUDA on this thread ("" if none). */
const char *picha_b200_last_error(void);
/* Kernels this library has launched so far (process-wide, all devices). */
uint64_t picha_b200_launch_count(void);
/* Which kernel served the most recent resize launched from the calling thread (diagnostics and
 * tests; there is nothing like it in the reference): 0 none yet, 1 bit-exact kernel
 * (resize_exact.cu), 2 generic throughput kernel (resize_fast.cuh), 3 downscaling kernel with
 * 4-row groups, 4 with 8-row groups, 6 with the integer-ratio horizontal pass (resize_down.cuh), 5 upscaling kernel,
 * 7 its wide-window variant (vertical upscale with a horizontal downscale; resize_up.cuh). */
int picha_b200_last_resize_kernel(void);

/* ---- format helpers: pixelBytes / pixelChannels / NativeImage::row_stride,
 *      src/picha.h:174-200,212-215 ----------------------------------------------------- */
int picha_b200_pixel_bytes(int pixel);
int picha_b200_pixel_channels(int pixel);
int picha_b200_row_stride(int width, int pixel);

/* ---- option resolution, so the addon keeps the reference's semantics ----------------- */

/* getResizeOptions, src/resize.cc:173-198: default {cubic, 0.70}; a `filter` key sets the
 * width to 1.0; a `filterScale` key overrides it; NaN or <= 0 -> ERR_INVALID_FILTER_WIDTH;
 * unknown filter -> ERR_INVALID_FILTER. */
int picha_b200_resolve_resize_options(int has_filter, int filter_tag,
                                      int has_filter_scale, double filter_scale,
                                      int *tag_out, float *width_out);
/* getSettings, src/colorconvert.cc:6-22 with the defaults of src/colorconvert.h:12: NaN
 * means "key absent"; the three weights are always rescaled by 1/(r+g+b). */
void picha_b200_resolve_color_settings(double red, double green, double blue, float out_rgb[3]);

/* ---- host entry points (caller-owned host buffers; blocking) -------------------------
 * Only width*bytes of each dst row is written; stride padding is left untouched
 * (the reference's dst buffer is uninitialised there: src/picha.cc:119-133). */

/* resizeImage, src/resize.cc:270-280 (called from :293 and :399). */
int picha_b200_resize(const picha_b200_image *src, picha_b200_image *dst,
                      int filter_tag, float filter_width);
int picha_b200_resize_ex(const picha_b200_image *src, picha_b200_image *dst,
                         int filter_tag, float filter_width, unsigned flags);
/* doColorConvert, src/colorconvert.cc:171-188 (called from :201 and :288).  The three
 * weights are ColorSettings as doColorConvert sees them, i.e. already resolved. */
int picha_b200_color_convert(const picha_b200_image *src, picha_b200_image *dst,
                             float r_factor, float g_factor, float b_factor);

/* cmyk_to_rgb, src/jpegcodec.cc:36-42 (called per decoded row from :96): rgb[c] = cmyk[c] * cmyk[3] / 255,
 * truncating integer arithmetic -- the pixel loop of the JPEG decoder's CMYK path (the
 * step in front of the hot path; the decoder itself stays on the host).  `cmyk` carries 4 bytes per
 * pixel (C, M, Y, K as libjpeg delivers them) in an image whose pixel is PICHA_B200_RGBA; `rgb` is
 * PICHA_B200_RGB of the same size.  Bit-exact. */
int picha_b200_cmyk_to_rgb(const picha_b200_image *cmyk, picha_b200_image *rgb);

/* Data-parallel batches of independent images (no cross-image step).  device >= 0 runs the
 * whole batch on that GPU; device = -1 shards contiguous blocks of the batch across every
 * GPU of the box, one host thread + streams + pinned staging per GPU, no collective. */
int picha_b200_resize_batch(int n, const picha_b200_image *srcs, picha_b200_image *dsts,
                            int filter_tag, float filter_width, unsigned flags, int device);
int picha_b200_color_convert_batch(int n, const picha_b200_image *srcs, picha_b200_image *dsts,
                                   float r_factor, float g_factor, float b_factor, int device);

/* Page-locked host memory for image buffers (the addon's newJsImage can hand these to
 * Nan::NewBuffer): host entry points copy straight from/to it with no staging pass. */
/* How the batch entry points split their work (pure host logic, no device needed -- exposed so that the
 * multi-GPU plumbing can be tested on CPU): shard `index` of `shards` owns images [lo, hi) (device = -1: one shard
 * per GPU; the reference's counterpart is many uv_queue_work items, src/resize.cc:362-364); and the
 * chunks of same-shape images a shard is cut into, each one kernel launch.  plan_batch returns the number of
 * chunks and fills up to `cap` entries. */
int picha_b200_shard_range(int n, int shards, int index, int *lo, int *hi);
int picha_b200_plan_batch(int n, const picha_b200_image *srcs, const picha_b200_image *dsts, int lanes,
                          int *chunk_first, int *chunk_count, int cap);

void *picha_b200_host_alloc(size_t bytes);
void picha_b200_host_free(void *p);

/* ---- device entry points (device-resident images; asynchronous on `stream`) ----------
 * `src0`/`dst0` describe image 0; image i lives at data + i*step bytes (n = 1: steps
 * ignored).  All images of a call share shape, format and stride.  `stream` is a
 * cudaStream_t (NULL = legacy default stream) of the current device. */
int picha_b200_resize_device(int n, const picha_b200_image *src0, int64_t src_step,
                             const picha_b200_image *dst0, int64_t dst_step,
                             int filter_tag, float filter_width, unsigned flags, void *stream);
int picha_b200_color_convert_device(int n, const picha_b200_image *src0, int64_t src_step,
                                    const picha_b200_image *dst0, int64_t dst_step,
                                    float r_factor, float g_factor, float b_factor, void *stream);
int picha_b200_cmyk_to_rgb_device(int n, const picha_b200_image *cmyk0, int64_t cmyk_step,
                                  const picha_b200_image *rgb0, int64_t rgb_step, void *stream);

/* ---- resize, then convert, in one kernel ---------------------------------
 * dst = doColorConvert(resizeImage(src)) -- src/resize.cc:270-280 followed by src/colorconvert.cc:171-188, the
 * pair a thumbnailer runs back to back (index.js:62-72, README.md:33-37) -- without the intermediate image: the
 * resize kernels put every resized pixel through the reference's conversion in their pack stage.  `dst` has the
 * destination's size AND pixel format; r/g/b are resolved luma weights (picha_b200_resolve_color_settings).
 * Results are those of the two calls in sequence (bit-identical with PICHA_B200_EXACT). */
int picha_b200_resize_convert(const picha_b200_image *src, picha_b200_image *dst, int filter_tag, float filter_width,
                              float r_factor, float g_factor, float b_factor, unsigned flags);
int picha_b200_resize_convert_batch(int n, const picha_b200_image *srcs, picha_b200_image *dsts, int filter_tag,
                                    float filter_width, float r_factor, float g_factor, float b_factor,
                                    unsigned flags, int device);
int picha_b200_resize_convert_device(int n, const picha_b200_image *src0, int64_t src_step,
                                     const picha_b200_image *dst0, int64_t dst_step, int filter_tag,
                                     float filter_width, float r_factor, float g_factor, float b_factor,
                                     unsigned flags, void *stream);
/* Synthetic pixels, i.i.d. uniform over the full channel range, from a counter-based hash of
 * (seed, image, byte offset in the payload): the same bytes picha_b200.synthetic.fill_host
 * produces, so host and device can regenerate any image of a benchmark batch. */
int picha_b200_synthetic_fill_device(int n, const picha_b200_image *img0, int64_t step,
                                     uint64_t seed, uint64_t first_image, void *stream);

/* ---- introspection (tests) ------------------------------------------------------------ */

/* The per-axis contribution table the kernels consume, in the reference's order
 * (makeContribs, src/resize.cc:19-50): for output i, taps left[i] .. left[i]+count[i]-1
 * with normalised weights at weights[offset[i] ...].  For each tap, eff_row (may be NULL)
 * receives the source row the reference's ring buffer actually reads in the vertical pass
 * (src/resize.cc:83,108,126).  Returns the number of weights, or a negative status. */
int picha_b200_contribs(int filter_tag, float filter_width, int srcsize, int dstsize,
                        int *left, int *count, int *offset,
                        float *weights, int *eff_row, int cap);

/* The dense weight blocks of the upscaling kernel's wide-window variant for one horizontal axis (csrc/tables.h:
 * WideBlocks; a host-side table, no device needed): block g = output columns 4g .. 4g+3 holds, for each of the
 * `window` source pixels from the first tap of column 4g on, the weight that pixel carries into each of the 4 columns
 * (unscaled here; 0 where it is not one of the column's taps).  Returns the window (source pixels per block), 0 if
 * this axis has no such table (no block spans more than 8 pixels, or some block more than 64), or a negative status;
 * blocks receives (dstsize + 3) / 4 * window * 4 floats if it holds at least cap of them. */
int picha_b200_wide_blocks(int filter_tag, float filter_width, int srcsize, int dstsize, float *blocks, int cap);

#ifdef __cplusplus
}
#endif
#endif /* PICHA_B200_H */
