// C-ABI layer (include/picha_b200.h): validation with the reference's error behaviour, device
// selection, per-thread-safe lanes (stream + pinned staging + device scratch), the resize plan
// cache, host batches pipelined over streams and sharded over GPUs, and the device entry points.
//
// Reference seams replaced here:
//   resizeImage(const ResizeOptions&, NativeImage&, NativeImage&)      src/resize.cc:270-280
//   doColorConvert(const ColorSettings&, NativeImage&, NativeImage&)   src/colorconvert.cc:171-188
// and the checks their NAN callers make before reaching them (src/resize.cc:321-403,
// src/colorconvert.cc:215-291, src/picha.cc:61-85).
#include "../../include/picha_b200.h"

#include <cuda_runtime.h>
#include <nvtx3/nvToolsExt.h>   // header-only: ranges cost nothing unless a tool injects itself (nsys, ncu --nvtx)

#include <atomic>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <list>
#include <map>
#include <memory>
#include <mutex>
#include <string>
#include <thread>
#include <tuple>
#include <vector>

#include "kernels.h"
#include "tables.h"

namespace picha_b200 {

namespace {

// NVTX range around a C-ABI call (tracing row)
struct Range {
	explicit Range(const char *name) { nvtxRangePushA(name); }
	~Range() { nvtxRangePop(); }
};

thread_local std::string g_last_error;
std::atomic<uint64_t> g_launches{0};
std::atomic<int> g_default_device{-1};   // -1: not chosen yet (env PICHA_B200_DEVICE or 0)

int fail_cuda(cudaError_t e, const char *what) {
	g_last_error = std::string(what) + ": " + cudaGetErrorName(e) + " (" + cudaGetErrorString(e) + ")";
	cudaGetLastError();   // clear the sticky-free error state
	if (e == cudaErrorNoDevice || e == cudaErrorInsufficientDriver || e == cudaErrorInitializationError)
		return PICHA_B200_ERR_NO_DEVICE;
	if (e == cudaErrorMemoryAllocation) return PICHA_B200_ERR_NOMEM;
	return PICHA_B200_ERR_CUDA;
}
#define CU(call)                                              \
	do {                                                      \
		cudaError_t e_ = (call);                              \
		if (e_ != cudaSuccess) return fail_cuda(e_, #call);   \
	} while (0)

size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

// ---- grow-only buffers -----------------------------------------------------------------

struct PinnedBuf {
	uint8_t *p = nullptr;
	size_t cap = 0;
	int ensure(size_t n) {
		if (n <= cap) return 0;
		if (p) cudaFreeHost(p);
		p = nullptr; cap = 0;
		size_t want = align_up(n + n / 4, 1 << 20);
		CU(cudaHostAlloc((void **)&p, want, cudaHostAllocPortable));
		cap = want;
		return 0;
	}
	void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// PICHA_B200_DEBUG_GUARD=1 (debugging aid; compute-sanitizer is not available everywhere): every lane buffer sits
// between two 8 MB guard zones filled with a pattern, checked after each host call -- a kernel that writes outside
// its destination is reported (last_error) instead of silently corrupting a neighbouring allocation.
const bool g_debug_guard = getenv("PICHA_B200_DEBUG_GUARD") != nullptr;
constexpr size_t kGuard = size_t(8) << 20;

struct DeviceBuf {
	uint8_t *p = nullptr, *raw = nullptr;
	size_t cap = 0;
	int ensure(size_t n) {
		if (n <= cap) return 0;
		if (raw) cudaFree(raw);
		p = raw = nullptr; cap = 0;
		size_t want = align_up(n + n / 4, 1 << 20);
		const size_t guard = g_debug_guard ? kGuard : 0;
		CU(cudaMalloc((void **)&raw, want + 2 * guard));
		if (guard) {
			CU(cudaMemset(raw, 0xA5, guard));
			CU(cudaMemset(raw + guard + want, 0xA5, guard));
		}
		p = raw + guard;
		cap = want;
		return 0;
	}
	// (debug) number of guard bytes that no longer hold the pattern; the stream the buffer is used on must be idle
	long long guard_damage() const {
		if (!g_debug_guard || !raw) return 0;
		std::vector<uint8_t> h(kGuard);
		long long bad = 0;
		for (int side = 0; side < 2; ++side) {
			if (cudaMemcpy(h.data(), side ? raw + kGuard + cap : raw, kGuard, cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
			for (uint8_t b : h) bad += b != 0xA5;
		}
		return bad;
	}
	void release() { if (raw) cudaFree(raw); p = raw = nullptr; cap = 0; }
};

// ---- resize plans ------------------------------------------------------------------------

struct Plan {
	AxisTable x, y;
	uint32_t *blob = nullptr;   // device
	ResizeTables t{};
	FastTables ft{};            // tile_w / band_h are filled per launch
	FastAxisY fy;               // vertical axis of the fast path (host; sliced into kernel parameters)
	FastAxisX fx;               // horizontal axis of the fast path (host copy, for launch planning)
	int fast_tile_w[kNumPixels] = {0, 0, 0, 0, 0, 0, 0, 0};
	int fast_align_px[kNumPixels] = {1, 1, 1, 1, 1, 1, 1, 1};   // tile origins are multiples of this many source pixels
	int fast_tile_w96[kNumPixels] = {0, 0, 0, 0, 0, 0, 0, 0}, fast_tile_w128[kNumPixels] = {0, 0, 0, 0, 0, 0, 0, 0};
	~Plan() { if (blob) cudaFree(blob); }
};

// Smallest pixel count whose byte size is a multiple of `quantum` bytes: tile origins (source boxes and bulk
// copies start on 16-byte boundaries) and tile widths (whole 16-byte destination stores) are multiples of it.
int align_pixels(int bytes_per_pixel, int quantum = 16) {
	int unit = quantum;
	while (unit > 1 && (unit / 2 * bytes_per_pixel) % quantum == 0) unit /= 2;
	return unit;
}

typedef std::tuple<int, uint32_t, int, int, int, int> PlanKey;   // filter, width bits, sw, sh, dw, dh

// A lane is what one in-flight host call owns: a stream, staging and device scratch.
struct Lane {
	cudaStream_t stream = nullptr;
	PinnedBuf hin, hout;
	DeviceBuf din, dout;
	// pending copy-out of staged results (batch pipelining): images whose rows wait in `hout`
	struct Pending { const picha_b200_image *dst; size_t offset, pitch; };
	std::vector<Pending> pending;
};

struct Device {
	int id = 0;
	int smem_optin = 0;
	std::mutex mu;
	std::vector<Lane *> idle;
	std::vector<Lane *> all;
	std::map<PlanKey, std::shared_ptr<Plan>> plans;
	std::list<PlanKey> lru;

	Lane *acquire() {
		{
			std::lock_guard<std::mutex> g(mu);
			if (!idle.empty()) { Lane *l = idle.back(); idle.pop_back(); return l; }
		}
		Lane *l = new Lane();
		if (cudaStreamCreateWithFlags(&l->stream, cudaStreamNonBlocking) != cudaSuccess) { delete l; return nullptr; }
		std::lock_guard<std::mutex> g(mu);
		all.push_back(l);
		return l;
	}
	void release(Lane *l) {
		std::lock_guard<std::mutex> g(mu);
		idle.push_back(l);
	}
};

std::mutex g_devices_mu;
std::vector<Device *> g_devices;   // index = CUDA ordinal
int g_device_count = -1;

int device_count() {
	std::lock_guard<std::mutex> g(g_devices_mu);
	if (g_device_count < 0) {
		int n = 0;
		cudaError_t e = cudaGetDeviceCount(&n);
		if (e != cudaSuccess) { cudaGetLastError(); n = 0; }
		g_device_count = n;
		g_devices.assign(n, nullptr);
	}
	return g_device_count;
}

Device *get_device(int ordinal) {
	int n = device_count();
	if (ordinal < 0 || ordinal >= n) return nullptr;
	std::lock_guard<std::mutex> g(g_devices_mu);
	if (!g_devices[ordinal]) {
		Device *d = new Device();
		d->id = ordinal;
		cudaDeviceGetAttribute(&d->smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, ordinal);
		g_devices[ordinal] = d;
	}
	return g_devices[ordinal];
}

int default_ordinal() {
	int d = g_default_device.load();
	if (d >= 0) return d;
	const char *env = getenv("PICHA_B200_DEVICE");
	return env ? atoi(env) : 0;
}

// Build (or fetch) the plan for one resize shape on `dev`. The current CUDA device must be dev->id.
int get_plan(Device *dev, int tag, float width, int sw, int sh, int dw, int dh, std::shared_ptr<Plan> *out) {
	uint32_t wbits;
	memcpy(&wbits, &width, 4);
	PlanKey key(tag, wbits, sw, sh, dw, dh);
	{
		std::lock_guard<std::mutex> g(dev->mu);
		auto it = dev->plans.find(key);
		if (it != dev->plans.end()) {
			dev->lru.remove(key);
			dev->lru.push_front(key);
			*out = it->second;
			return 0;
		}
	}
	// Not cached: the tables are built and uploaded WITHOUT the device's lock (first calls of different shapes on
	// different threads proceed side by side; two threads building the same shape both finish, one result is kept).
	std::shared_ptr<Plan> p(new Plan());
	build_axis(tag, width, sw, dw, p->x);
	build_axis(tag, width, sh, dh, p->y);

	// Rows of output per CTA band: as many as fit the exact kernel's shared-memory tile at the
	// widest format (4 channels), starting from 8.
	const size_t smem_budget = dev->smem_optin > 0 ? (size_t)dev->smem_optin : 48 * 1024;
	int band_h = 8, tile_w = 32, max_rows = 0;
	std::vector<int> blo, brows;
	auto fits = [&]() { return (size_t)max_rows * tile_w * 4 * sizeof(float) <= smem_budget; };
	for (;;) {
		const int nb = (dh + band_h - 1) / band_h;
		blo.assign(nb, 0); brows.assign(nb, 0);
		max_rows = 0;
		for (int b = 0; b < nb; ++b) {
			int lo = sh, hi = -1;
			for (int y = b * band_h; y < dh && y < (b + 1) * band_h; ++y)
				for (int k = 0; k < p->y.count[y]; ++k) {
					int r = p->y.eff[p->y.start[y] + k];
					if (r < lo) lo = r;
					if (r > hi) hi = r;
				}
			blo[b] = lo; brows[b] = hi - lo + 1;
			if (brows[b] > max_rows) max_rows = brows[b];
		}
		if (fits() || band_h == 1) break;
		band_h /= 2;
	}
	while (!fits() && tile_w > 1) tile_w /= 2;   // extreme downscales: one output row, few columns per CTA
	if (!fits()) return PICHA_B200_ERR_UNSUPPORTED;
	const int nb = (dh + band_h - 1) / band_h;
	if (nb > 65535) return PICHA_B200_ERR_UNSUPPORTED;

	std::vector<uint32_t> blob;
	auto put_i = [&](const std::vector<int> &v) { size_t o = blob.size(); blob.insert(blob.end(), v.begin(), v.end()); return o; };
	auto put_f = [&](const std::vector<float> &v) {
		size_t o = blob.size();
		blob.resize(o + v.size());
		if (!v.empty()) memcpy(&blob[o], v.data(), v.size() * 4);
		return o;
	};
	size_t o_xfirst = put_i(p->x.first), o_xcount = put_i(p->x.count), o_xstart = put_i(p->x.start);
	size_t o_ycount = put_i(p->y.count), o_ystart = put_i(p->y.start), o_yeff = put_i(p->y.eff);
	size_t o_blo = put_i(blo), o_brows = put_i(brows);
	size_t o_xw = put_f(p->x.w), o_yw = put_f(p->y.w);

	// fast path tables (vertical axis as accumulator ring / row window, horizontal axis padded)
	FastAxisY &fy = p->fy;
	FastAxisX &fx = p->fx;
	build_fast_y(p->y, kFastMaxDepth, fy);
	build_fast_x(p->x, fx);
	size_t o_fxw = put_f(fx.w), o_fxfirst = put_i(fx.first), o_fxcount = put_i(fx.count);
	size_t o_fxrow = put_i(fx.urow), o_fxuw = put_f(fx.uw);
	// vertical upscales whose 4-column groups touch more than 8 source pixels (horizontal downscales, wide filters):
	// the dense weight blocks of the upscaling kernel's wide-window variant (2^120 = the kernel's up::kHExp)
	WideBlocks wide;
	if (fy.variant == FastAxisY::kUp) build_wide_blocks(fx, dw, 64, std::ldexp(1.0f, 120), wide);
	while (blob.size() % 4) blob.push_back(0u);      // (the kernel reads the blocks as float4)
	size_t o_wide = put_f(wide.w);
	FlatRows flat[2];
	size_t o_ecol[2], o_esrc[2], o_eoff[2];
	for (int ci = 0; ci < 2; ++ci) {
		build_flat_rows(fx, ci ? 3 : 1, flat[ci]);
		o_ecol[ci] = put_i(flat[ci].col); o_esrc[ci] = put_i(flat[ci].src); o_eoff[ci] = put_i(flat[ci].off);
	}
	// Conditioning: sum |w| of an output's taps is how much the filter amplifies rounding differences.
	// It is <= 2 per axis for every filter at sane widths, but lanczos and catmulrom narrowed to
	// filterScale ~0.5 have upscaling phases whose weights nearly cancel before makeContribs normalises
	// them (src/resize.cc:41-47): sums of hundreds, thousands, or not finite at all.  Only the reference's
	// own summation order reproduces its result there, so those plans stay with the bit-exact kernel.
	auto amplification = [](const AxisTable &t) {
		float worst = 0.0f;
		for (int i = 0; i < t.dst_size; ++i) {
			float sum = 0.0f;
			for (int k = 0; k < t.count[i]; ++k) sum += std::fabs(t.w[t.start[i] + k]);
			if (!(sum <= 1e30f)) return INFINITY;   // infinite or NaN weights
			if (sum > worst) worst = sum;
		}
		return worst;
	};
	const bool well_conditioned = amplification(p->x) * amplification(p->y) <= 8.0f;
	for (int px = 0; px < kNumPixels; ++px) {
		const PixelInfo pi = pixel_info(px);
		// Tile widths: multiples of the 16-byte pixel group when the destination is the big side (vector
		// stores on every tile); when the image shrinks by 2x or more the destination is small, so any
		// multiple of 4 pixels will do and the source row of the tile can be filled to the brim.
		// (also what the downscaling kernel's integer-ratio horizontal pass wants: blocks of 4 columns)
		const int unit = p->x.scale >= 2.0f ? 4 : align_pixels(pi.bytes);
		// tile origins: 16-byte aligned in the source row.  Bulk copies need that by definition; a tensor-map box of
		// 32-bit words may nominally start at any word, but a start that is not 16-byte aligned faults on this
		// part (cudaErrorIllegalInstruction; tried).
		p->fast_align_px[px] = align_pixels(pi.bytes, 16);
		p->fast_tile_w[px] = fy.variant == FastAxisY::kNone || !well_conditioned ? 0
			: fast_tile_width(fx.first.data(), fx.count.data(), dw, pi.channels, unit, p->fast_align_px[px], 512);
		// the downscaling kernel's wide variants (8-bit formats, images that shrink by 2x or more)
		if (p->fast_tile_w[px] > 0 && !pi.deep && fy.variant == FastAxisY::kDown && p->x.scale >= 2.0f) {
			p->fast_tile_w96[px] = fast_tile_width(fx.first.data(), fx.count.data(), dw, pi.channels, unit, p->fast_align_px[px], 512, 1536);
			p->fast_tile_w128[px] = fast_tile_width(fx.first.data(), fx.count.data(), dw, pi.channels, unit, p->fast_align_px[px], 512, 2048);
		}
	}

	CU(cudaMalloc((void **)&p->blob, blob.size() * 4));
	// The upload must have LANDED before any stream may launch a kernel that reads the tables.  A plain cudaMemcpy
	// does not promise that for pageable sources -- it may return once the data sits in the driver's staging buffer,
	// with the DMA to the device still queued (and the queue may hold a 30 MB image upload issued just before on the
	// calling lane's non-blocking stream, which a kernel launched on that stream then overtakes the tables on).
	// That was the intermittent cudaErrorIllegalAddress / occasional wrong output of the round-1 fuzz runs: a kernel
	// reading table memory the copy had not reached yet.  An asynchronous copy on a stream
	// that is then synchronised is complete when the call returns.
	if (getenv("PICHA_B200_DEBUG_RACY_UPLOAD")) {   // (A/B evidence only: the round-1 upload, profiles/r02_fault_ab.txt)
		CU(cudaMemcpy(p->blob, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice));
	} else {
		cudaStream_t up = nullptr;
		CU(cudaStreamCreateWithFlags(&up, cudaStreamNonBlocking));
		cudaError_t ce = cudaMemcpyAsync(p->blob, blob.data(), blob.size() * 4, cudaMemcpyHostToDevice, up);
		if (ce == cudaSuccess) ce = cudaStreamSynchronize(up);
		cudaStreamDestroy(up);
		if (ce != cudaSuccess) return fail_cuda(ce, "upload of the resize tables");
	}
	const int *ib = reinterpret_cast<const int *>(p->blob);
	const float *fb = reinterpret_cast<const float *>(p->blob);
	p->t.xfirst = ib + o_xfirst; p->t.xcount = ib + o_xcount; p->t.xstart = ib + o_xstart;
	p->t.ycount = ib + o_ycount; p->t.ystart = ib + o_ystart; p->t.yeff = ib + o_yeff;
	p->t.band_lo = ib + o_blo; p->t.band_rows = ib + o_brows;
	p->t.xw = fb + o_xw; p->t.yw = fb + o_yw;
	p->t.band_h = band_h;
	p->t.tile_w = tile_w;
	p->t.max_band_rows = max_rows;
	p->ft.xfirst = ib + o_fxfirst; p->ft.xcount = ib + o_fxcount;
	p->ft.xw = fb + o_fxw; p->ft.xstride = fx.stride;
	p->ft.xtaps = fx.taps;
	p->ft.h_xfirst = fx.first.data(); p->ft.h_xcount = fx.count.data(); p->ft.h_xw = fx.w.data();
	p->ft.h_xrow = fx.urow.data();
	p->ft.xrow = ib + o_fxrow; p->ft.xuw = fb + o_fxuw; p->ft.xunique = fx.unique;
	for (int ci = 0; ci < 2; ++ci) {
		p->ft.xe_col[ci] = ib + o_ecol[ci]; p->ft.xe_src[ci] = ib + o_esrc[ci]; p->ft.xe_off[ci] = ib + o_eoff[ci];
		p->ft.xe_count[ci] = (int)flat[ci].src.size();
	}
	p->ft.xwide = fb + o_wide; p->ft.xwide_window = wide.window;
	p->ft.xshort = fx.taps <= 4 ? 4 : (fx.taps <= 8 ? 8 : 0);

	std::vector<std::shared_ptr<Plan>> evicted;   // released (cudaFree: a device-wide synchronisation) after the lock
	{
		std::lock_guard<std::mutex> g(dev->mu);
		auto it = dev->plans.find(key);
		if (it != dev->plans.end()) {             // another thread was faster: keep its plan
			evicted.push_back(p);
			p = it->second;
		} else {
			dev->plans[key] = p;
			dev->lru.push_front(key);
			while (dev->lru.size() > 64) {
				auto old = dev->plans.find(dev->lru.back());
				if (old != dev->plans.end()) { evicted.push_back(old->second); dev->plans.erase(old); }
				dev->lru.pop_back();
			}
		}
	}
	*out = p;
	return 0;
}

// ---- validation (what the reference's NAN callers check before reaching the seam) ---------

int check_image(const picha_b200_image *im) {
	if (!im || !im->data) return PICHA_B200_ERR_INVALID_IMAGE;
	PixelInfo pi = pixel_info(im->pixel);
	if (pi.bytes == 0) return PICHA_B200_ERR_INVALID_IMAGE;                    // src/picha.cc:68-69
	if (im->width < 0 || im->height <= 0) return PICHA_B200_ERR_INVALID_IMAGE;  // src/picha.cc:78 (height != 0)
	if ((int64_t)im->stride < (int64_t)im->width * pi.bytes) return PICHA_B200_ERR_INVALID_IMAGE;   // lib/image.js:13-14
	return 0;
}

int check_resize(const picha_b200_image *s, const picha_b200_image *d, int tag, float width) {
	int rc = check_image(s);
	if (rc) return rc;
	if (s->width <= 0) return PICHA_B200_ERR_INVALID_IMAGE;
	if (!d || d->width <= 0 || d->height <= 0) return PICHA_B200_ERR_INVALID_DIMENSIONS;   // src/resize.cc:343-346
	if (tag < 0 || tag >= PICHA_B200_NUM_FILTERS) return PICHA_B200_ERR_INVALID_FILTER;     // src/resize.cc:184-187
	if (width != width || width <= 0) return PICHA_B200_ERR_INVALID_FILTER_WIDTH;           // src/resize.cc:192-195
	rc = check_image(d);
	if (rc) return rc;
	if (s->pixel != d->pixel) return PICHA_B200_ERR_FORMAT_MISMATCH;                        // src/resize.cc:137
	return 0;
}

// resize, then convert: the checks of both reference entry points, minus the equal-format requirement
int check_resize_convert(const picha_b200_image *s, const picha_b200_image *d, int tag, float width) {
	int rc = check_image(s);
	if (rc) return rc;
	if (s->width <= 0) return PICHA_B200_ERR_INVALID_IMAGE;
	if (!d || d->width <= 0 || d->height <= 0) return PICHA_B200_ERR_INVALID_DIMENSIONS;
	if (tag < 0 || tag >= PICHA_B200_NUM_FILTERS) return PICHA_B200_ERR_INVALID_FILTER;
	if (width != width || width <= 0) return PICHA_B200_ERR_INVALID_FILTER_WIDTH;
	if (pixel_info(d->pixel).bytes == 0) return PICHA_B200_ERR_INVALID_PIXEL;
	return check_image(d);
}

int check_convert(const picha_b200_image *s, const picha_b200_image *d) {
	int rc = check_image(s);
	if (rc) return rc;
	if (!d) return PICHA_B200_ERR_INVALID_IMAGE;
	if (pixel_info(d->pixel).bytes == 0) return PICHA_B200_ERR_INVALID_PIXEL;              // src/colorconvert.cc:235-239
	rc = check_image(d);
	if (rc) return rc;
	if (s->width != d->width || s->height != d->height) return PICHA_B200_ERR_SIZE_MISMATCH;   // src/colorconvert.cc:138-139
	return 0;
}


// ---- host <-> device staging ---------------------------------------------------------------

bool is_pinned(const void *p) {
	cudaPointerAttributes a;
	if (cudaPointerGetAttributes(&a, p) != cudaSuccess) { cudaGetLastError(); return false; }
	return a.type == cudaMemoryTypeHost;
}

size_t device_pitch(const picha_b200_image &im) {
	return align_up((size_t)im.width * pixel_info(im.pixel).bytes, 128);
}

// Host image -> lane->din + offset (pitch `pitch`), asynchronously on the lane's stream.  Pinned sources are
// copied in place (one strided copy); pageable ones are re-pitched into the lane's pinned staging buffer, which the
// caller then sends with one copy per chunk (flush_uploads).  The buffers have been sized by the caller.
int upload(Lane *lane, const picha_b200_image &im, size_t pitch, size_t offset, bool *staged) {
	const size_t row = (size_t)im.width * pixel_info(im.pixel).bytes;
	if (row == 0) return 0;
	if (is_pinned(im.data)) {
		CU(cudaMemcpy2DAsync(lane->din.p + offset, pitch, im.data, im.stride, row, im.height, cudaMemcpyHostToDevice, lane->stream));
		return 0;
	}
	const uint8_t *s = static_cast<const uint8_t *>(im.data);
	for (int y = 0; y < im.height; ++y) memcpy(lane->hin.p + offset + y * pitch, s + (size_t)y * im.stride, row);
	*staged = true;
	return 0;
}

// lane->dout + offset -> host image (payload bytes only), asynchronously; results for pageable destinations go to
// the lane's pinned buffer (one copy per chunk: flush_downloads) and are copied out by finish().
int download(Lane *lane, const picha_b200_image &im, size_t pitch, size_t offset, bool *staged) {
	const size_t row = (size_t)im.width * pixel_info(im.pixel).bytes;
	if (row == 0) return 0;
	if (is_pinned(im.data)) {
		CU(cudaMemcpy2DAsync(im.data, im.stride, lane->dout.p + offset, pitch, row, im.height, cudaMemcpyDeviceToHost, lane->stream));
		return 0;
	}
	lane->pending.push_back(Lane::Pending{&im, offset, pitch});
	*staged = true;
	return 0;
}

int finish(Lane *lane) {
	CU(cudaStreamSynchronize(lane->stream));
	if (g_debug_guard) {
		const long long a = lane->din.guard_damage(), b = lane->dout.guard_damage();
		if (a || b) {
			g_last_error = "PICHA_B200_DEBUG_GUARD: " + std::to_string(a) + " guard bytes around the source buffer and " +
			               std::to_string(b) + " around the destination buffer were overwritten";
			return PICHA_B200_ERR_CUDA;
		}
	}
	for (const Lane::Pending &p : lane->pending) {
		const picha_b200_image &im = *p.dst;
		const size_t row = (size_t)im.width * pixel_info(im.pixel).bytes;
		uint8_t *d = static_cast<uint8_t *>(im.data);
		for (int y = 0; y < im.height; ++y) memcpy(d + (size_t)y * im.stride, lane->hout.p + p.offset + y * p.pitch, row);
	}
	lane->pending.clear();
	return 0;
}

DevBatch dev_batch(uint8_t *base, int64_t step, size_t pitch, const picha_b200_image &im) {
	DevBatch b;
	b.base = base; b.step = step; b.stride = (int)pitch;
	b.width = im.width; b.height = im.height; b.pixel = im.pixel;
	return b;
}

// luma: resize, then convert to d.pixel in the kernels' pack stage (nullptr, or same format: plain resize)
int run_resize(Device *dev, const DevBatch &s, const DevBatch &d, int n, int tag, float width, unsigned flags,
               cudaStream_t stream, const float *luma = nullptr) {
	FuseArgs fuse{-1, 0.0f, 0.0f, 0.0f};
	if (luma && d.pixel != s.pixel) fuse = FuseArgs{d.pixel, luma[0], luma[1], luma[2]};
	std::shared_ptr<Plan> plan;
	int rc = get_plan(dev, tag, width, s.width, s.height, d.width, d.height, &plan);
	if (rc) return rc;
	int launches = 0;
	cudaError_t e = cudaErrorNotSupported;
	// Small images are launch-latency bound either way, so they get the bit-exact kernel; the
	// throughput kernel (within +-1 LSB) takes everything large enough for bandwidth to matter.
	const bool large = (long long)s.width * s.height >= 128 * 128 &&
	                   (long long)d.width * d.height * pixel_info(s.pixel).channels >= 4096;
	if (!(flags & PICHA_B200_EXACT) && (large || (flags & PICHA_B200_FORCE_FAST)) && plan->fast_tile_w[s.pixel] > 0) {
		FastTables ft = plan->ft;
		ft.tile_w = plan->fast_tile_w[s.pixel];
		ft.tile_w96 = plan->fast_tile_w96[s.pixel];
		ft.tile_w128 = plan->fast_tile_w128[s.pixel];
		ft.align_px = plan->fast_align_px[s.pixel];
		e = launch_resize_fast(s, d, n, ft, plan->fy, fuse, stream, &launches);
		if (e == cudaErrorNotSupported) cudaGetLastError();
	}
	if (e == cudaErrorNotSupported) {
		launches = 0;
		e = launch_resize_exact(s, d, n, plan->t, fuse, stream, &launches);
		g_last_resize_kernel = 1;
	}
	g_launches += launches;
	if (e == cudaErrorInvalidValue) { cudaGetLastError(); return PICHA_B200_ERR_UNSUPPORTED; }
	if (e != cudaSuccess) return fail_cuda(e, "resize kernel launch");
	return 0;
}

int run_convert(const DevBatch &s, const DevBatch &d, int n, float r, float g, float b, cudaStream_t stream, bool cmyk = false) {
	int launches = 0;
	cudaError_t e = cmyk ? launch_cmyk_to_rgb(s, d, n, stream, &launches) : launch_color_convert(s, d, n, r, g, b, stream, &launches);
	g_launches += launches;
	if (e != cudaSuccess) return fail_cuda(e, "color convert kernel launch");
	return 0;
}

// Host entry points run on whichever device they were asked for without changing the calling
// thread's current device (the caller may be another CUDA user, e.g. PyTorch).
struct DeviceGuard {
	int prev = -1;
	DeviceGuard() { if (cudaGetDevice(&prev) != cudaSuccess) { cudaGetLastError(); prev = -1; } }
	~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

struct Op {
	bool resize;
	int tag; float width; unsigned flags;   // resize
	float r, g, b;                          // convert
	bool cmyk;                              // convert: the JPEG decoder's cmyk_to_rgb instead of doColorConvert
	bool fused;                             // resize, then convert to the destination's format (r, g, b as for convert)
};

bool same_shape(const picha_b200_image &a, const picha_b200_image &b) {
	return a.width == b.width && a.height == b.height && a.pixel == b.pixel;
}

// k same-shape images through one lane: uploads, ONE kernel launch over the chunk, downloads (all asynchronous on
// the lane's stream).
int submit_chunk(Device *dev, Lane *lane, const Op &op, int k, const picha_b200_image *srcs, const picha_b200_image *dsts) {
	const picha_b200_image &s = srcs[0], &d = dsts[0];
	const size_t sp = device_pitch(s), dp = device_pitch(d);
	const size_t sstep = align_up(sp * s.height, 256), dstep = align_up(dp * d.height, 256);
	int rc = lane->din.ensure(sstep * k);
	if (!rc) rc = lane->dout.ensure(dstep * k);
	bool any_pageable_src = false, any_pageable_dst = false;
	for (int i = 0; i < k; ++i) {
		any_pageable_src |= !is_pinned(srcs[i].data);
		any_pageable_dst |= !is_pinned(dsts[i].data);
	}
	if (!rc && any_pageable_src) rc = lane->hin.ensure(sstep * k);
	if (!rc && any_pageable_dst) rc = lane->hout.ensure(dstep * k);
	if (rc) return rc;
	// uploads: staged images form runs of consecutive slots, each sent with one copy
	int run_begin = -1;
	auto flush = [&](int end) -> int {
		if (run_begin < 0) return 0;
		CU(cudaMemcpyAsync(lane->din.p + sstep * run_begin, lane->hin.p + sstep * run_begin, sstep * (end - run_begin - 1) + sp * s.height,
		                   cudaMemcpyHostToDevice, lane->stream));
		run_begin = -1;
		return 0;
	};
	for (int i = 0; i < k; ++i) {
		bool staged = false;
		rc = upload(lane, srcs[i], sp, sstep * i, &staged);
		if (rc) return rc;
		if (staged && run_begin < 0) run_begin = i;
		if (!staged && (rc = flush(i))) return rc;
	}
	if ((rc = flush(k))) return rc;
	DevBatch sb = dev_batch(lane->din.p, (int64_t)sstep, sp, s), db = dev_batch(lane->dout.p, (int64_t)dstep, dp, d);
	const float luma[3] = {op.r, op.g, op.b};
	rc = op.resize ? run_resize(dev, sb, db, k, op.tag, op.width, op.flags, lane->stream, op.fused ? luma : nullptr)
	               : run_convert(sb, db, k, op.r, op.g, op.b, lane->stream, op.cmyk);
	if (rc) return rc;
	run_begin = -1;
	auto flush_down = [&](int end) -> int {
		if (run_begin < 0) return 0;
		CU(cudaMemcpyAsync(lane->hout.p + dstep * run_begin, lane->dout.p + dstep * run_begin, dstep * (end - run_begin - 1) + dp * d.height,
		                   cudaMemcpyDeviceToHost, lane->stream));
		run_begin = -1;
		return 0;
	};
	for (int i = 0; i < k; ++i) {
		bool staged = false;
		rc = download(lane, dsts[i], dp, dstep * i, &staged);
		if (rc) return rc;
		if (staged && run_begin < 0) run_begin = i;
		if (!staged && (rc = flush_down(i))) return rc;
	}
	return flush_down(k);
}

struct Chunk { int first, count; };

// Runs of same-shape images cut into chunks of one kernel launch each: at least two chunks per lane when the batch
// allows (overlap of copies and kernels), at most 32 images or ~128 MB of source per chunk.
std::vector<Chunk> plan_chunks(int n, const picha_b200_image *srcs, const picha_b200_image *dsts, int lanes) {
	std::vector<Chunk> chunks;
	if (lanes < 1) lanes = 1;
	const int by_count = n / (2 * lanes) > 0 ? n / (2 * lanes) : 1;
	for (int i = 0; i < n;) {
		const size_t bytes = align_up(device_pitch(srcs[i]) * (size_t)(srcs[i].height > 0 ? srcs[i].height : 0), 256);
		const int by_bytes = bytes > 0 && (size_t(128) << 20) / bytes > 0 ? (int)((size_t(128) << 20) / bytes) : 1;
		const int kmax = by_count < by_bytes ? (by_count < 32 ? by_count : 32) : (by_bytes < 32 ? by_bytes : 32);
		int k = 1;
		while (i + k < n && k < kmax && same_shape(srcs[i], srcs[i + k]) && same_shape(dsts[i], dsts[i + k])) ++k;
		chunks.push_back(Chunk{i, k});
		i += k;
	}
	return chunks;
}

// Image block of shard `index` of `shards`: contiguous, no exchange.
void shard_range(int n, int shards, int index, int *lo, int *hi) {
	*lo = (int)((int64_t)n * index / shards);
	*hi = (int)((int64_t)n * (index + 1) / shards);
}

// A batch on one device: runs of same-shape images are cut into chunks, chunks go round-robin over a few lanes
// so the copy engines and the SMs overlap (H2D of chunk c+1 with the kernel of chunk c and the D2H of chunk c-1).
// A chunk is one kernel launch, whatever its size (the launch of a single image is a few tiles: it cannot fill
// the GPU, and a descriptor upload per image costs more than a thumbnail's resize).
int batch_on_device(int ordinal, const Op &op, int n, const picha_b200_image *srcs, picha_b200_image *dsts) {
	Device *dev = get_device(ordinal);
	if (!dev) { g_last_error = "no such CUDA device"; return PICHA_B200_ERR_NO_DEVICE; }
	DeviceGuard guard;                 // the caller's current device is restored on return
	CU(cudaSetDevice(ordinal));
	const int kLanes = n < 4 ? (n < 1 ? 1 : n) : 4;
	Lane *lanes[4] = {nullptr, nullptr, nullptr, nullptr};
	bool busy[4] = {false, false, false, false};
	int rc = 0;
	for (int l = 0; l < kLanes; ++l) {
		lanes[l] = dev->acquire();
		if (!lanes[l]) { rc = PICHA_B200_ERR_CUDA; g_last_error = "cudaStreamCreate failed"; }
	}
	const std::vector<Chunk> chunks = plan_chunks(n, srcs, dsts, kLanes);
	// One host thread per lane takes chunks off a shared counter: re-pitching pageable images into pinned staging
	// memory is a host memcpy (~10 GB/s per thread), and four of them keep the copy engine busy where one cannot.
	std::atomic<int> next{0};
	std::atomic<int> first_rc{rc};
	std::string first_err;
	std::mutex err_mu;
	auto lane_loop = [&](int l, bool set_device) {
		if (set_device && cudaSetDevice(ordinal) != cudaSuccess) { first_rc = PICHA_B200_ERR_CUDA; return; }
		bool busy = false;
		int my = 0;
		while (!first_rc.load()) {
			const int c = next.fetch_add(1);
			if (c >= (int)chunks.size()) break;
			if (busy) { my = finish(lanes[l]); busy = false; if (my) break; }
			my = submit_chunk(dev, lanes[l], op, chunks[c].count, srcs + chunks[c].first, dsts + chunks[c].first);
			if (my) break;
			busy = true;
		}
		if (busy) { int r2 = finish(lanes[l]); if (!my) my = r2; }
		else if (my || first_rc.load()) cudaStreamSynchronize(lanes[l]->stream);
		lanes[l]->pending.clear();
		if (my) {
			std::lock_guard<std::mutex> g(err_mu);
			if (!first_rc.load()) { first_rc = my; first_err = g_last_error; }
		}
	};
	if (!rc) {
		const int workers = (int)chunks.size() < kLanes ? (int)chunks.size() : kLanes;
		if (workers <= 1) {
			lane_loop(0, false);
		} else {
			std::vector<std::thread> threads;
			for (int l = 1; l < workers; ++l) threads.emplace_back(lane_loop, l, true);
			lane_loop(0, false);
			for (auto &t : threads) t.join();
		}
		rc = first_rc.load();
		if (rc && !first_err.empty()) g_last_error = first_err;
	}
	for (int l = 0; l < kLanes; ++l)
		if (lanes[l]) dev->release(lanes[l]);
	return rc;
}

int run_batch(const Op &op, int n, const picha_b200_image *srcs, picha_b200_image *dsts, int device) {
	if (n < 0 || (n > 0 && (!srcs || !dsts))) return PICHA_B200_ERR_INVALID_ARGUMENT;
	for (int i = 0; i < n; ++i) {
		int rc = op.fused ? check_resize_convert(&srcs[i], &dsts[i], op.tag, op.width)
		                  : op.resize ? check_resize(&srcs[i], &dsts[i], op.tag, op.width) : check_convert(&srcs[i], &dsts[i]);
		if (rc) return rc;
	}
	const int ndev = device_count();
	if (ndev <= 0) { g_last_error = "no CUDA device"; return PICHA_B200_ERR_NO_DEVICE; }
	if (n == 0) return 0;
	if (device >= 0) return batch_on_device(device, op, n, srcs, dsts);

	// Shard contiguous blocks of the batch across every GPU; independent images, no collective.
	const int shards = n < ndev ? n : ndev;
	std::vector<int> rcs(shards, 0);
	std::vector<std::string> errs(shards);
	std::vector<std::thread> threads;
	for (int s = 0; s < shards; ++s) {
		int lo, hi;
		shard_range(n, shards, s, &lo, &hi);
		threads.emplace_back([&, s, lo, hi]() {
			rcs[s] = batch_on_device(s, op, hi - lo, srcs + lo, dsts + lo);
			if (rcs[s]) errs[s] = g_last_error;
		});
	}
	for (auto &t : threads) t.join();
	for (int s = 0; s < shards; ++s)
		if (rcs[s]) { g_last_error = errs[s]; return rcs[s]; }
	return 0;
}

int current_device_checked(Device **out) {
	if (device_count() <= 0) { g_last_error = "no CUDA device"; return PICHA_B200_ERR_NO_DEVICE; }
	int ord = 0;
	CU(cudaGetDevice(&ord));
	*out = get_device(ord);
	return *out ? 0 : PICHA_B200_ERR_NO_DEVICE;
}

}  // namespace

int sm_count() {
	int dev = 0, v = 0;
	if (cudaGetDevice(&dev) != cudaSuccess) return 148;
	if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
	return v;
}

cudaError_t grow_dynamic_smem(const void *kernel, int bytes, SmemGrant *granted) {
	static std::mutex mu;
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	if (dev < 0 || dev >= 16) return cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	std::lock_guard<std::mutex> g(mu);
	if (bytes <= granted->bytes[dev]) return cudaSuccess;
	e = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
	if (e == cudaSuccess) granted->bytes[dev] = bytes;
	return e;
}

int max_dynamic_smem() {
	int dev = 0, v = 0;
	if (cudaGetDevice(&dev) != cudaSuccess) return 48 * 1024;
	if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || v <= 0) return 48 * 1024;
	return v;
}

}  // namespace picha_b200

using namespace picha_b200;

extern "C" {

int picha_b200_version(void) { return PICHA_B200_VERSION; }

int picha_b200_device_count(void) { return device_count(); }

int picha_b200_init(int device) {
	const int n = device_count();
	if (n <= 0) { g_last_error = "no CUDA device"; return PICHA_B200_ERR_NO_DEVICE; }
	if (device >= n) return PICHA_B200_ERR_INVALID_ARGUMENT;
	for (int d = (device < 0 ? 0 : device); d < (device < 0 ? n : device + 1); ++d) {
		Device *dev = get_device(d);
		CU(cudaSetDevice(d));
		CU(cudaFree(0));
		Lane *l = dev->acquire();
		if (!l) return PICHA_B200_ERR_CUDA;
		dev->release(l);
	}
	if (device >= 0) g_default_device.store(device);
	return 0;
}

void picha_b200_shutdown(void) {
	std::lock_guard<std::mutex> g(g_devices_mu);
	for (Device *d : g_devices) {
		if (!d) continue;
		cudaSetDevice(d->id);
		cudaDeviceSynchronize();
		for (Lane *l : d->all) {
			l->hin.release(); l->hout.release(); l->din.release(); l->dout.release();
			if (l->stream) cudaStreamDestroy(l->stream);
			delete l;
		}
		d->plans.clear();
		release_resize_descriptors(d->id);
		delete d;
	}
	g_devices.assign(g_devices.size(), nullptr);
}

const char *picha_b200_strerror(int status) {
	switch (status) {
		case PICHA_B200_OK: return "ok";
		case PICHA_B200_ERR_INVALID_IMAGE: return "invalid image";
		case PICHA_B200_ERR_INVALID_DIMENSIONS: return "invalid dimensions";
		case PICHA_B200_ERR_INVALID_FILTER: return "invalid filter mode";
		case PICHA_B200_ERR_INVALID_FILTER_WIDTH: return "invalid filter width";
		case PICHA_B200_ERR_INVALID_PIXEL: return "expected pixel mode";
		case PICHA_B200_ERR_FORMAT_MISMATCH: return "source and destination pixel formats differ";
		case PICHA_B200_ERR_SIZE_MISMATCH: return "source and destination dimensions differ";
		case PICHA_B200_ERR_NO_DEVICE: return "no CUDA device (picha_b200 has no CPU fallback)";
		case PICHA_B200_ERR_CUDA: return "CUDA error";
		case PICHA_B200_ERR_NOMEM: return "out of memory";
		case PICHA_B200_ERR_UNSUPPORTED: return "unsupported shape";
		case PICHA_B200_ERR_INVALID_ARGUMENT: return "invalid argument";
	}
	return "unknown status";
}

const char *picha_b200_last_error(void) { return g_last_error.c_str(); }

uint64_t picha_b200_launch_count(void) { return g_launches.load(); }

int picha_b200_last_resize_kernel(void) { return g_last_resize_kernel; }

int picha_b200_pixel_bytes(int pixel) { return pixel_info(pixel).bytes; }
int picha_b200_pixel_channels(int pixel) { return pixel_info(pixel).channels; }
int picha_b200_row_stride(int width, int pixel) { return (pixel_info(pixel).bytes * width + 3) & ~3; }

int picha_b200_resolve_resize_options(int has_filter, int filter_tag, int has_filter_scale, double filter_scale,
                                      int *tag_out, float *width_out) {
	int tag = PICHA_B200_CUBIC;
	float width = 0.70f;                                     // ResizeOptions(), src/resize.cc:174
	if (has_filter) {
		width = 1.0f;                                        // :182
		if (filter_tag < 0 || filter_tag >= PICHA_B200_NUM_FILTERS) return PICHA_B200_ERR_INVALID_FILTER;
		tag = filter_tag;
	}
	if (has_filter_scale) {
		width = (float)filter_scale;                         // :191
		if (width != width || width <= 0) return PICHA_B200_ERR_INVALID_FILTER_WIDTH;
	}
	if (tag_out) *tag_out = tag;
	if (width_out) *width_out = width;
	return 0;
}

void picha_b200_resolve_color_settings(double red, double green, double blue, float out[3]) {
	float r = 0.299f, g = 0.587f, b = float(0.114);          // src/colorconvert.h:12
	if (red == red) r = (float)red;                          // src/colorconvert.cc:11
	if (green == green) g = (float)green;
	if (blue == blue) b = (float)blue;
	const float n = 1.0f / (r + g + b);                      // :18
	out[0] = r * n; out[1] = g * n; out[2] = b * n;
}

int picha_b200_resize_ex(const picha_b200_image *src, picha_b200_image *dst, int filter_tag, float filter_width,
                         unsigned flags) {
	Range nvtx_range("picha_b200_resize_ex");
	Op op{};
	op.resize = true; op.tag = filter_tag; op.width = filter_width; op.flags = flags;
	if (!src || !dst) return PICHA_B200_ERR_INVALID_IMAGE;
	return run_batch(op, 1, src, dst, default_ordinal());
}

int picha_b200_resize(const picha_b200_image *src, picha_b200_image *dst, int filter_tag, float filter_width) {
	return picha_b200_resize_ex(src, dst, filter_tag, filter_width, 0);
}

int picha_b200_color_convert(const picha_b200_image *src, picha_b200_image *dst, float r, float g, float b) {
	Range nvtx_range("picha_b200_color_convert");
	Op op{};
	op.resize = false; op.r = r; op.g = g; op.b = b;
	if (!src || !dst) return PICHA_B200_ERR_INVALID_IMAGE;
	return run_batch(op, 1, src, dst, default_ordinal());
}

int picha_b200_cmyk_to_rgb(const picha_b200_image *cmyk, picha_b200_image *rgb) {
	Range nvtx_range("picha_b200_cmyk_to_rgb");
	Op op{};
	op.resize = false; op.cmyk = true;
	if (!cmyk || !rgb) return PICHA_B200_ERR_INVALID_IMAGE;
	if (cmyk->pixel != PICHA_B200_RGBA || rgb->pixel != PICHA_B200_RGB) return PICHA_B200_ERR_FORMAT_MISMATCH;
	return run_batch(op, 1, cmyk, rgb, default_ordinal());
}

int picha_b200_resize_batch(int n, const picha_b200_image *srcs, picha_b200_image *dsts, int filter_tag,
                            float filter_width, unsigned flags, int device) {
	Range nvtx_range("picha_b200_resize_batch");
	Op op{};
	op.resize = true; op.tag = filter_tag; op.width = filter_width; op.flags = flags;
	return run_batch(op, n, srcs, dsts, device);
}

int picha_b200_color_convert_batch(int n, const picha_b200_image *srcs, picha_b200_image *dsts, float r, float g,
                                   float b, int device) {
	Range nvtx_range("picha_b200_color_convert_batch");
	Op op{};
	op.resize = false; op.r = r; op.g = g; op.b = b;
	return run_batch(op, n, srcs, dsts, device);
}

int picha_b200_resize_convert(const picha_b200_image *src, picha_b200_image *dst, int filter_tag, float filter_width,
                              float r, float g, float b, unsigned flags) {
	Range nvtx_range("picha_b200_resize_convert");
	Op op{};
	op.resize = true; op.fused = true; op.tag = filter_tag; op.width = filter_width; op.flags = flags; op.r = r; op.g = g; op.b = b;
	if (!src || !dst) return PICHA_B200_ERR_INVALID_IMAGE;
	return run_batch(op, 1, src, dst, default_ordinal());
}

int picha_b200_resize_convert_batch(int n, const picha_b200_image *srcs, picha_b200_image *dsts, int filter_tag,
                                    float filter_width, float r, float g, float b, unsigned flags, int device) {
	Range nvtx_range("picha_b200_resize_convert_batch");
	Op op{};
	op.resize = true; op.fused = true; op.tag = filter_tag; op.width = filter_width; op.flags = flags; op.r = r; op.g = g; op.b = b;
	return run_batch(op, n, srcs, dsts, device);
}

int picha_b200_resize_convert_device(int n, const picha_b200_image *src0, int64_t src_step, const picha_b200_image *dst0,
                                     int64_t dst_step, int filter_tag, float filter_width, float r, float g, float b,
                                     unsigned flags, void *stream) {
	Range nvtx_range("picha_b200_resize_convert_device");
	if (n < 0) return PICHA_B200_ERR_INVALID_ARGUMENT;
	int rc = check_resize_convert(src0, dst0, filter_tag, filter_width);
	if (rc) return rc;
	Device *dev = nullptr;
	rc = current_device_checked(&dev);
	if (rc) return rc;
	if (n == 0) return 0;
	DevBatch s = dev_batch(static_cast<uint8_t *>(src0->data), src_step, src0->stride, *src0);
	DevBatch d = dev_batch(static_cast<uint8_t *>(dst0->data), dst_step, dst0->stride, *dst0);
	const float luma[3] = {r, g, b};
	return run_resize(dev, s, d, n, filter_tag, filter_width, flags, static_cast<cudaStream_t>(stream), luma);
}

int picha_b200_shard_range(int n, int shards, int index, int *lo, int *hi) {
	if (n < 0 || shards <= 0 || index < 0 || index >= shards || !lo || !hi) return PICHA_B200_ERR_INVALID_ARGUMENT;
	shard_range(n, shards, index, lo, hi);
	return 0;
}

int picha_b200_plan_batch(int n, const picha_b200_image *srcs, const picha_b200_image *dsts, int lanes,
                          int *chunk_first, int *chunk_count, int cap) {
	if (n < 0 || (n > 0 && (!srcs || !dsts))) return PICHA_B200_ERR_INVALID_ARGUMENT;
	const std::vector<Chunk> chunks = plan_chunks(n, srcs, dsts, lanes);
	for (size_t i = 0; i < chunks.size() && (int)i < cap; ++i) {
		if (chunk_first) chunk_first[i] = chunks[i].first;
		if (chunk_count) chunk_count[i] = chunks[i].count;
	}
	return (int)chunks.size();
}

void *picha_b200_host_alloc(size_t bytes) {
	void *p = nullptr;
	if (device_count() <= 0) return nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) { cudaGetLastError(); return nullptr; }
	return p;
}

void picha_b200_host_free(void *p) {
	if (p) cudaFreeHost(p);
}

int picha_b200_resize_device(int n, const picha_b200_image *src0, int64_t src_step, const picha_b200_image *dst0,
                             int64_t dst_step, int filter_tag, float filter_width, unsigned flags, void *stream) {
	Range nvtx_range("picha_b200_resize_device");
	if (n < 0) return PICHA_B200_ERR_INVALID_ARGUMENT;
	int rc = check_resize(src0, dst0, filter_tag, filter_width);
	if (rc) return rc;
	Device *dev = nullptr;
	rc = current_device_checked(&dev);
	if (rc) return rc;
	if (n == 0) return 0;
	DevBatch s = dev_batch(static_cast<uint8_t *>(src0->data), src_step, src0->stride, *src0);
	DevBatch d = dev_batch(static_cast<uint8_t *>(dst0->data), dst_step, dst0->stride, *dst0);
	return run_resize(dev, s, d, n, filter_tag, filter_width, flags, static_cast<cudaStream_t>(stream));
}

int picha_b200_color_convert_device(int n, const picha_b200_image *src0, int64_t src_step, const picha_b200_image *dst0,
                                    int64_t dst_step, float r, float g, float b, void *stream) {
	Range nvtx_range("picha_b200_color_convert_device");
	if (n < 0) return PICHA_B200_ERR_INVALID_ARGUMENT;
	int rc = check_convert(src0, dst0);
	if (rc) return rc;
	Device *dev = nullptr;
	rc = current_device_checked(&dev);
	if (rc) return rc;
	if (n == 0) return 0;
	DevBatch s = dev_batch(static_cast<uint8_t *>(src0->data), src_step, src0->stride, *src0);
	DevBatch d = dev_batch(static_cast<uint8_t *>(dst0->data), dst_step, dst0->stride, *dst0);
	return run_convert(s, d, n, r, g, b, static_cast<cudaStream_t>(stream));
}

int picha_b200_cmyk_to_rgb_device(int n, const picha_b200_image *cmyk0, int64_t cmyk_step, const picha_b200_image *rgb0,
                                  int64_t rgb_step, void *stream) {
	Range nvtx_range("picha_b200_cmyk_to_rgb_device");
	if (n < 0) return PICHA_B200_ERR_INVALID_ARGUMENT;
	int rc = check_convert(cmyk0, rgb0);
	if (rc) return rc;
	if (cmyk0->pixel != PICHA_B200_RGBA || rgb0->pixel != PICHA_B200_RGB) return PICHA_B200_ERR_FORMAT_MISMATCH;
	Device *dev = nullptr;
	rc = current_device_checked(&dev);
	if (rc) return rc;
	if (n == 0) return 0;
	DevBatch s = dev_batch(static_cast<uint8_t *>(cmyk0->data), cmyk_step, cmyk0->stride, *cmyk0);
	DevBatch d = dev_batch(static_cast<uint8_t *>(rgb0->data), rgb_step, rgb0->stride, *rgb0);
	return run_convert(s, d, n, 0.0f, 0.0f, 0.0f, static_cast<cudaStream_t>(stream), true);
}

int picha_b200_synthetic_fill_device(int n, const picha_b200_image *img0, int64_t step, uint64_t seed,
                                     uint64_t first_image, void *stream) {
	Range nvtx_range("picha_b200_synthetic_fill_device");
	if (n < 0) return PICHA_B200_ERR_INVALID_ARGUMENT;
	int rc = check_image(img0);
	if (rc) return rc;
	Device *dev = nullptr;
	rc = current_device_checked(&dev);
	if (rc) return rc;
	if (n == 0 || img0->width == 0) return 0;
	DevBatch b = dev_batch(static_cast<uint8_t *>(img0->data), step, img0->stride, *img0);
	int launches = 0;
	cudaError_t e = launch_synthetic_fill(b, n, seed, first_image, static_cast<cudaStream_t>(stream), &launches);
	g_launches += launches;
	if (e != cudaSuccess) return fail_cuda(e, "synthetic fill launch");
	return 0;
}

int picha_b200_contribs(int filter_tag, float filter_width, int srcsize, int dstsize, int *left, int *count,
                        int *offset, float *weights, int *eff_row, int cap) {
	if (filter_tag < 0 || filter_tag >= PICHA_B200_NUM_FILTERS) return PICHA_B200_ERR_INVALID_FILTER;
	if (filter_width != filter_width || filter_width <= 0) return PICHA_B200_ERR_INVALID_FILTER_WIDTH;
	if (srcsize <= 0 || dstsize <= 0) return PICHA_B200_ERR_INVALID_DIMENSIONS;
	AxisTable t;
	build_axis(filter_tag, filter_width, srcsize, dstsize, t);
	for (int i = 0; i < dstsize; ++i) {
		if (left) left[i] = t.first[i];
		if (count) count[i] = t.count[i];
		if (offset) offset[i] = t.start[i];
	}
	const int n = (int)t.w.size();
	for (int i = 0; i < n && i < cap; ++i) {
		if (weights) weights[i] = t.w[i];
		if (eff_row) eff_row[i] = t.eff[i];
	}
	return n;
}

int picha_b200_wide_blocks(int filter_tag, float filter_width, int srcsize, int dstsize, float *blocks, int cap) {
	if (filter_tag < 0 || filter_tag >= PICHA_B200_NUM_FILTERS) return PICHA_B200_ERR_INVALID_FILTER;
	if (filter_width != filter_width || filter_width <= 0) return PICHA_B200_ERR_INVALID_FILTER_WIDTH;
	if (srcsize <= 0 || dstsize <= 0) return PICHA_B200_ERR_INVALID_DIMENSIONS;
	AxisTable t;
	build_axis(filter_tag, filter_width, srcsize, dstsize, t);
	FastAxisX fx;
	build_fast_x(t, fx);
	WideBlocks wide;
	build_wide_blocks(fx, dstsize, 64, 1.0f, wide);
	const int n = (int)wide.w.size();
	if (blocks && cap >= n)
		for (int i = 0; i < n; ++i) blocks[i] = wide.w[i];
	return wide.window;
}

}  // extern "C"
