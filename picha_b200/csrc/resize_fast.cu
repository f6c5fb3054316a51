// Launch planning of the throughput resize path (reference: src/resize.cc:66-134): picks one of three
// kernels per shape, cuts the image into row bands, slices the host-built vertical tables into kernel
// parameters and launches.
//
//   resize_down.cuh   downscales (vertical depth <= 8): vertical pass in registers on thread-private
//                     columns, horizontal pass from shared memory; TMA-staged source rows.
//   resize_up.cuh     upscales (row window <= 6, <= 8 source pixels per 4 outputs): horizontal pass from the
//                     staged row, vertical pass in registers on thread-private output pixels; no
//                     intermediate in shared memory.
//   resize_fast.cuh   the generic first-generation kernel, for everything else the fast path takes.
//
// All three read each source byte from HBM once (plus band and tile halos), compute with fused FMAs in
// an order of their own (within +-1 LSB of the reference, not bit-exact: resize_exact.cu is the
// bit-exact path) and take the vertical weights as warp-uniform operands from the constant bank
// (`VTable`, a __grid_constant__ kernel parameter; one launch per group of row bands whose tables fit it).
// Which pass runs first follows from which pass is expensive: the one at the larger resolution must be
// the thread-private one -- vertical-first for downscales (a horizontal-first pass would push every
// source value through shared memory as a float once per tap), horizontal-first for upscales (the
// vertical pass runs at output resolution).
//
// No tensor cores: FP32 FMA on a byte stream, bounded by HBM on one side and FP32 issue on the other
// (DESIGN.md sections 5.3 and 6 have the arithmetic and the measurements).
#include <algorithm>

#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <mutex>
#include <utility>

#include "resize_up.cuh"
#include "tables.h"

namespace picha_b200 {

using namespace fast;

thread_local int g_last_resize_kernel = 0;

namespace {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
	static const EncodeTiledFn fn = []() -> EncodeTiledFn {   // resolved once, thread-safe
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
		    q == cudaDriverEntryPointSuccess)
			return reinterpret_cast<EncodeTiledFn>(p);
		cudaGetLastError();
		return nullptr;
	}();
	return fn;
}

int padded_depth(int d) {   // DEPTH the kernel is instantiated with
	const int steps[] = {3, 4, 5, 6, 8, 12};
	for (int s : steps)
		if (d <= s) return s;
	return 0;
}

}  // namespace

int fast_tile_width(const int *xfirst, const int *xcount, int dst_w, int channels, int unit, int align_px, int cap, int row_values) {
	const int limit = row_values / channels;   // source pixels one CTA row holds
	// Among the widths whose source span fits, take the one that keeps both passes busiest: pass 1
	// works on all NT*NV values of the staged row whether the tile needs them or not, pass 2 on
	// rounds of NT (pixel, row) items.  The passes are weighted by which one dominates.
	const bool down = dst_w > 0 && xfirst[dst_w - 1] + xcount[dst_w - 1] > dst_w;   // source wider than destination
	const double w1 = down ? 0.7 : 0.3;
	int best = 0;
	double best_score = 0;
	for (int tw = cap / unit * unit; tw >= unit; tw -= unit) {
		bool ok = true;
		long long span_sum = 0;
		int tiles = 0;
		for (int x0 = 0; x0 < dst_w && ok; x0 += tw, ++tiles) {
			const int x1 = (x0 + tw < dst_w ? x0 + tw : dst_w) - 1;
			int hi = 0;   // the right edge is not monotone in x in general (trimmed zero taps): scan the tile
			for (int x = x0; x <= x1; ++x)
				if (xfirst[x] + xcount[x] > hi) hi = xfirst[x] + xcount[x];
			int lo = xfirst[x0];
			for (int x = x0; x <= x1; ++x)
				if (xfirst[x] < lo) lo = xfirst[x];
			// the kernels start a tile's staged row at the 16-byte boundary below its first tap (align_px pixels)
			if (lo != xfirst[x0] || hi - xfirst[x0] / align_px * align_px > limit) ok = false;
			span_sum += hi - lo;
		}
		if (!ok || tiles == 0) continue;
		const double util1 = (double)span_sum / ((double)tiles * limit);
		const int items = (tw < dst_w ? tw : dst_w) * 4;
		const int round = row_values / 16;   // granularity of the horizontal pass: one (pixel, row) item per thread
		const double util2 = (double)items / ((items + round - 1) / round * round);
		const double score = 1.0 / (w1 / util1 + (1.0 - w1) / util2);
		if (score > best_score * 1.02) { best_score = score; best = tw; }   // prefer wider tiles on near-ties
	}
	return best;
}

// Device-resident TMA descriptors: one 128-byte slot per resize call from a ring per device.  The slot is written
// on the device, in stream order ahead of the resize launches (a one-warp kernel that takes the encoded descriptor
// as a parameter), so the consumers' fence.proxy.tensormap::generic.acquire.gpu has the scope the writer needs.
// A slot's lifetime is explicit: an event recorded behind the last launch that reads it, and waited for (it has
// long completed, normally) before the slot is handed out again kMapSlots calls later -- callers may run resizes on
// any number of streams, so "long ago" is not an ordering.
namespace {

constexpr unsigned kMapSlots = 1024;

struct MapRing {
	CUtensorMap *slots = nullptr;
	cudaEvent_t events[kMapSlots] = {};
	unsigned next = 0;
};
std::mutex g_map_mu;
std::map<int, MapRing *> g_map_rings;   // device -> ring

__global__ void write_descriptor_kernel(CUtensorMap *slot, const __grid_constant__ CUtensorMap map) {
	const uint64_t *src = reinterpret_cast<const uint64_t *>(&map);
	uint64_t *dst = reinterpret_cast<uint64_t *>(slot);
	if (threadIdx.x < sizeof(CUtensorMap) / 8) dst[threadIdx.x] = src[threadIdx.x];
}

// Reserves a slot (waiting for the kernels of its previous use) and returns it with its index.
CUtensorMap *map_slot(unsigned *index) {
	int dev = 0;
	if (cudaGetDevice(&dev) != cudaSuccess) return nullptr;
	MapRing *ring;
	unsigned i;
	cudaEvent_t ev;
	{
		std::lock_guard<std::mutex> lock(g_map_mu);
		auto it = g_map_rings.find(dev);
		if (it == g_map_rings.end()) {
			ring = new MapRing();
			if (cudaMalloc((void **)&ring->slots, kMapSlots * sizeof(CUtensorMap)) != cudaSuccess) { delete ring; return nullptr; }
			it = g_map_rings.emplace(dev, ring).first;
		}
		ring = it->second;
		i = ring->next++ % kMapSlots;
		ev = ring->events[i];
	}
	if (ev && cudaEventSynchronize(ev) != cudaSuccess) return nullptr;
	*index = i;
	return ring->slots + i;
}

// Behind the last launch that reads the slot.
cudaError_t map_slot_release(unsigned index, cudaStream_t stream) {
	int dev = 0;
	cudaError_t e = cudaGetDevice(&dev);
	if (e != cudaSuccess) return e;
	cudaEvent_t ev;
	{
		std::lock_guard<std::mutex> lock(g_map_mu);
		MapRing *ring = g_map_rings[dev];
		if (!ring->events[index]) {
			e = cudaEventCreateWithFlags(&ring->events[index], cudaEventDisableTiming);
			if (e != cudaSuccess) return e;
		}
		ev = ring->events[index];
	}
	return cudaEventRecord(ev, stream);
}

}  // namespace

// picha_b200_shutdown: the calling thread has selected `device` and synchronised it.
void release_resize_descriptors(int device) {
	std::lock_guard<std::mutex> lock(g_map_mu);
	auto it = g_map_rings.find(device);
	if (it == g_map_rings.end()) return;
	for (cudaEvent_t ev : it->second->events)
		if (ev) cudaEventDestroy(ev);
	cudaFree(it->second->slots);
	delete it->second;
	g_map_rings.erase(it);
}

cudaError_t launch_resize_fast(const DevBatch &src, const DevBatch &dst, int n, const FastTables &tables,
                               const FastAxisY &fy, const FuseArgs &fuse, cudaStream_t stream, int *launches) {
	static const int kBytes[8] = {3, 4, 1, 2, 2, 4, 6, 8}, kChannels[8] = {3, 4, 1, 2, 1, 2, 3, 4};
	const int bpp = kBytes[src.pixel], channels = kChannels[src.pixel];
	const bool deep = src.pixel >= 4;
	FastTables t = tables;
	const int depth = padded_depth(fy.depth);
	if (fy.variant < 0 || t.tile_w <= 0 || depth == 0) return cudaErrorNotSupported;
	if ((reinterpret_cast<uintptr_t>(src.base) & 15) || (src.stride & 15) || (n > 1 && (src.step & 15)))
		return cudaErrorNotSupported;
	if (n > 65535) return cudaErrorNotSupported;
	EncodeTiledFn encode = encode_fn();
	if (!encode) return cudaErrorNotSupported;
	// Downscales whose accumulator ring is at most 8 deep take the kernel of resize_down.cuh.
	bool use_down = fy.variant == FastAxisY::kDown && depth <= down::kMaxDepth;
	DownLaunch dl{};
	dl.threads = down::NT;
	if (use_down && !deep && depth <= 4 && !getenv("PICHA_B200_NO_WIDE") &&
	    !(channels == 4 && src.width % dst.width == 0 && src.width / dst.width <= 4)) {
		// 8-bit formats: 96- or 128-thread CTAs where their tiles divide the row so much better that fewer source
		// columns are computed in total (tiles x threads)
		auto cost = [&](int tile_w, int threads) { return tile_w > 0 ? (long long)((dst.width + tile_w - 1) / tile_w) * threads : (1LL << 60); };
		long long best = cost(t.tile_w, 64);
		if (cost(t.tile_w96, 96) * 10 < best * 9) { best = cost(t.tile_w96, 96); dl.threads = 96; }
		if (cost(t.tile_w128, 128) * 10 < best * 9 && cost(t.tile_w128, 128) < cost(t.tile_w96, 96)) dl.threads = 128;
		if (const char *w = getenv("PICHA_B200_DOWN_THREADS")) {
			const int want = atoi(w);
			if (want == 64 || (want == 96 && t.tile_w96 > 0) || (want == 128 && t.tile_w128 > 0)) dl.threads = want;
		}
		if (dl.threads == 96) t.tile_w = t.tile_w96;
		if (dl.threads == 128) t.tile_w = t.tile_w128;
	}
	if (use_down) {
		float wmax = 0;
		for (float w : fy.wv) wmax = std::fmax(wmax, std::fabs(w));
		if (!(wmax < 128.0f)) use_down = false;      // weights carry 2^120 in that kernel
		for (int d : fy.done)
			if (d > (int)down::kEvCount) use_down = false;   // the row loop's event flag counts to 7
		dl.da.nb = (channels & 1) ? (channels * t.xtaps + 3 + 4 * channels - 1) / (4 * channels) : (t.xtaps + 3) / 4;
		// few distinct weight rows (integer and small p/q ratios): the tile keeps just those
		const int rows = (channels & 1) ? t.xe_count[channels == 3] : t.xunique;
		dl.da.uniq = rows > 0 && rows * 2 <= t.tile_w;
		dl.da.wrows = dl.da.uniq ? rows : t.tile_w;
		dl.da.xscale = std::ldexp(1.0f, 149 - down::kVExp) / (deep ? 65535.0f : 255.0f);
		const int store_unit = (bpp == 4 || bpp == 8) ? bpp : (bpp == 2 && !deep) ? 2 : deep ? 2 : 1;
		dl.da.direct = ((reinterpret_cast<uintptr_t>(dst.base) | (uintptr_t)dst.stride | (uintptr_t)dst.step) & (store_unit - 1)) == 0;
		dl.da.fuse = fuse;
		if (fuse.dst_pixel >= 0) dl.da.direct = 1;   // converted pixels are stored channel by channel: any alignment
	}
	// 4-channel upscales with a shallow row window take the kernel of resize_up.cuh.
	// Upscales with a shallow row window take the kernel of resize_up.cuh (its stores are whole words
	// up to 16 bytes: the destination must be 16-byte aligned, which the library's own layout is).
	bool use_up = fy.variant == FastAxisY::kUp && depth <= up::kMaxDepth && !getenv("PICHA_B200_OLD_UP") &&
	              (fuse.dst_pixel >= 0 ||
	               ((reinterpret_cast<uintptr_t>(dst.base) | (uintptr_t)dst.stride | (uintptr_t)(n > 1 ? dst.step : 0)) & 15) == 0);
	UpLaunch ul{};
	const int *host_xfirst = t.h_xfirst, *host_xcount = t.h_xcount;
	const float *host_xw = t.h_xw;
	if (use_up) {
		// per tile: bytes of a source row to stage; per thread (4 output pixels): source pixels touched
		int win_px = 0, wpx = 0;
		for (int x0 = 0; x0 < dst.width; x0 += up::TILE) {
			const int x1 = std::min(dst.width, x0 + up::TILE);
			const int sx0 = host_xfirst[x0] / t.align_px * t.align_px;
			for (int x = x0; x < x1; ++x) win_px = std::max(win_px, host_xfirst[x] + host_xcount[x] - sx0);
			for (int g0 = x0; g0 < x1; g0 += up::NPX) {
				int lo = host_xfirst[g0], hi = 0;
				for (int x = g0; x < std::min(x1, g0 + up::NPX); ++x) {
					lo = std::min(lo, host_xfirst[x]);
					hi = std::max(hi, host_xfirst[x] + host_xcount[x]);
				}
				if (lo != host_xfirst[g0]) use_up = false;       // the window is anchored at the group's first column
				wpx = std::max(wpx, hi - lo);
			}
		}
		float wmax = 0;
		for (int i = 0; i < dst.width * t.xstride; ++i) wmax = std::fmax(wmax, std::fabs(host_xw[i]));
		if (!(wmax < 128.0f)) use_up = false;
		// more than 8 source pixels per 4 output columns (a horizontal downscale, a very wide filter): the wide-window
		// variant walks the window from the plan's dense weight blocks -- plain resizes only
		const bool wide = wpx > 8;
		if (wide && (t.xwide_window < wpx || fuse.dst_pixel >= 0 || getenv("PICHA_B200_NO_WIDE_UP"))) use_up = false;
		ul.wpx = wide ? 0 : wpx;
		ul.ua.window = t.xwide_window;
		ul.ua.win_bytes = (win_px * bpp + 15) & ~15;
		ul.ua.hscale = std::ldexp(1.0f, up::kHExp);
		ul.ua.fuse = fuse;
		if (up::smem_bytes(ul.ua.win_bytes, wide) > max_dynamic_smem()) use_up = false;
	}
	// The generic kernel is not a default route any more: shapes neither specialised kernel takes (vertical depth
	// above 8; a vertical upscale with a horizontal downscale; misaligned destinations of upscales) get the bit-exact
	// kernel.  It once failed with an illegal address under a fuzz sequence and the cause was never pinned down
	// (DESIGN.md section 9); PICHA_B200_GENERIC=1 (and the A/B switch PICHA_B200_OLD_UP) still reach it.
	if (!use_down && !use_up && (fuse.dst_pixel >= 0 || (!getenv("PICHA_B200_GENERIC") && !getenv("PICHA_B200_OLD_UP")))) return cudaErrorNotSupported;
	if (use_down) {
		// Rows per pass-2 group.  With 4, the eight lanes that share a shared-memory phase of a float4 read
		// are four rows of two neighbouring columns; they hit distinct banks only if the columns' windows
		// start the right distance apart (in float4 units: 4 mod 8 with rows one unit apart -- even channel
		// counts -- or an odd distance with rows two units apart -- odd channel counts).  When that holds
		// for few column pairs, groups of 8 rows (one column per phase: never a conflict) are worth their
		// shared memory.
		// (measured with 2 pixels in flight per thread for odd channel counts: only 4-channel pixels gain from
		// 8-row groups; the others lose more to the smaller number of CTAs per SM than to the conflicts)
		int pairs = 0, clean = 0;
		if (channels == 4) {
			const int x1 = std::min(dst.width, t.tile_w);
			for (int x = 0; x + 1 < x1; x += 2) {
				const int a0 = (host_xfirst[x] * channels) >> 2, a1 = (host_xfirst[x + 1] * channels) >> 2;
				const int d = (a1 - a0) & 7;
				++pairs;
				clean += (channels & 1) ? (d & 1) : d == 4;
			}
		}
		dl.group = pairs > 0 && clean * 4 < pairs * 3 && dl.threads == down::NT ? 8 : 4;   // (the wide variants: 4-row groups only)
		// (by columns -- pass2_cols -- the lanes of a phase are dealt out by bank group: 4-row groups never conflict either)
		if (down::by_columns(channels, 4)) dl.group = 4;
		if (const char *g = getenv("PICHA_B200_DOWN_G")) dl.group = atoi(g) == 8 && dl.threads == down::NT ? 8 : 4;
		if (fuse.dst_pixel >= 0) dl.group = 4;   // the converting kernels exist for 4-row groups only
	}
	// 4-channel pixels at an integer ratio of 2, 3 or 4: the horizontal pass with a sliding window (pass2_int4).
	// Every column's taps must lie in its nominal window [rq * x + off0, + rq * dx); columns that differ from the
	// regular one (clipped windows at the image edges) must fall into the first 2 and last 3 blocks of 4 columns.
	if (use_down && channels == 4 && dl.da.direct && !getenv("PICHA_B200_NO_P2INT") && dst.width >= 32 && t.tile_w % down::kIntU == 0 &&
	    src.width % dst.width == 0 && src.width / dst.width >= 2 && src.width / dst.width <= 4) {
		const int rq = src.width / dst.width, dw = dst.width, xm = dw / 2;
		const int *xf = t.h_xfirst, *xc = t.h_xcount, *xr = t.h_xrow;
		const int off0 = xf[xm] - rq * xm;
		int need = 0;
		bool ok = off0 >= -8;
		for (int x = 0; x < dw && ok; ++x) {
			const int d = xf[x] - (rq * x + off0);
			if (d < 0) ok = false;
			need = std::max(need, d + xc[x]);
		}
		const int dx = need <= 2 * rq ? 2 : need <= 4 * rq ? 4 : 6;
		if (need > 6 * rq) ok = false;
		auto regular = [&](int x) { return xr[x] == xr[xm] && xf[x] == rq * x + off0; };
		int lo = 0, hi = dw;
		while (ok && lo < dw && !regular(lo)) ++lo;
		while (ok && hi > lo && !regular(hi - 1)) --hi;
		for (int x = lo; x < hi && ok; ++x) ok = regular(x);
		const int nl = (lo + down::kIntU - 1) / down::kIntU, br0 = hi / down::kIntU, nblk = (dw + down::kIntU - 1) / down::kIntU;
		if (ok && nl <= 2 && nblk - br0 <= 3 && nl < br0) {
			dl.da.rq = rq; dl.da.dx = dx; dl.da.off0 = off0; dl.da.nl = nl; dl.da.br0 = br0;
			dl.group = 4;
		}
	}
	const int smem_total = use_up ? up::smem_bytes(ul.ua.win_bytes, ul.wpx == 0) : use_down && dl.da.rq > 0 ? down::smem_layout_int(dl.da.rq * dl.da.dx).total
	                       : use_down ? down::smem_layout(dl.group, t.tile_w, bpp, channels, dl.da.nb, dl.da.wrows, dl.da.direct != 0, dl.threads, down::by_columns(channels, dl.group)).total
	                                : smem_layout(deep, t.tile_w, bpp, t.xstride).total;
	if (smem_total > max_dynamic_smem()) return cudaErrorNotSupported;

	// The batch as a 3-D tensor of 32-bit words: (words per row, rows, images).
	CUtensorMap map;
	const cuuint64_t row_words = ((cuuint64_t)src.width * bpp + 3) / 4;
	if (row_words * 4 > (cuuint64_t)src.stride) return cudaErrorNotSupported;
	cuuint64_t dims[3] = {row_words, (cuuint64_t)src.height, (cuuint64_t)n};
	cuuint64_t strides[2] = {(cuuint64_t)src.stride, (cuuint64_t)(n > 1 ? src.step : (int64_t)src.stride * src.height)};
	if (strides[1] & 15) strides[1] = (strides[1] + 15) & ~15ull;   // n == 1: never dereferenced
	cuuint32_t box[3] = {(cuuint32_t)(use_down ? down::box_bytes(deep, dl.threads) / 4 : 256), (cuuint32_t)(use_down ? down::stage_rows(deep) : stage_rows(deep)), 1};
	cuuint32_t estr[3] = {1, 1, 1};
	CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, src.base, dims, strides, box, estr,
	                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
	                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) return cudaErrorNotSupported;
	// Bands: enough CTAs for ~16 waves of 4 CTAs/SM when the batch is small, tall strips (little
	// vertical halo) when it is large; multiples of 8 rows; small enough that one band's vertical
	// tables fit a launch's parameter block.
	const int WS = use_down ? down::weight_stride(depth) : (depth + 3) & ~3;   // floats per row of the vertical table
	const int dh = dst.height;
	// (CTAs that fit an SM: the downscaling kernel's wide variants and its 8-row groups take more of it each)
	int ctas_per_sm = use_down ? (dl.threads == 128 ? 3 : dl.threads == 96 || dl.group == 8 ? 4 : 6) : use_up ? 6 : 4;
	// (shared memory may allow fewer: a tile that keeps one weight row per column)
	ctas_per_sm = std::max(1, std::min(ctas_per_sm, (228 * 1024) / (smem_total + 1024)));
	const long long tiles = (long long)((dst.width + t.tile_w - 1) / t.tile_w) * n;
	// (the new kernels have a noticeable per-CTA start-up -- tables and a zeroed intermediate in shared
	// memory, the first stage's latency -- so they get fewer, taller bands: ~6 waves instead of ~16)
	const int waves = use_down || use_up ? 6 : 16;
	long long want = ((long long)sm_count() * ctas_per_sm * waves + tiles - 1) / tiles;
	const int max_bands = dh / 16 > 0 ? dh / 16 : 1;
	if (const char *b = getenv("PICHA_B200_BANDS")) want = atoi(b);   // (tuning experiments)
	if (want > max_bands) want = max_bands;
	if (want < 1) want = 1;
	int band_h = (int)(((dh + want - 1) / want + 7) / 8 * 8);
	// The downscaling kernel starts every band on a ring-stage boundary (a multiple of `rsk` source rows): its row
	// loop learns where stages end from a flag in the weight table, which is shared by all bands of a launch.
	const int rsk = use_down ? down::stage_rows(deep) : 1;
	auto first_row = [&](int y0) { return fy.smin[y0] & ~(rsk - 1); };
	auto band_fits = [&](int y0, int y1) {
		const int rows = fy.cum[y1 - 1] - first_row(y0) + 1;
		const int outs = fy.variant == 0 ? y1 - fy.ybase[first_row(y0)] : y1 - y0;
		// (the downscaling kernel reads the weights one row ahead and may start up to depth - 1 outputs early)
		if (use_down) return (rows + 2) * WS <= kYtabMax + kWtMax;   // (no per-output table: VTable::wdown)
		return outs + kFastMaxDepth <= kYtabMax && ((fy.variant == 0 ? rows : outs) + 1) * WS <= kWtMax;
	};
	for (;;) {
		bool ok = true;
		for (int y0 = 0; y0 < dh && ok; y0 += band_h) ok = band_fits(y0, std::min(dh, y0 + band_h));
		if (ok) break;
		if (band_h <= 8) return cudaErrorNotSupported;
		band_h = (band_h / 2 + 7) / 8 * 8;
	}
	t.band_h = band_h;
	// the upscaling kernel is instantiated for windows of 3, 4 and 6 rows: slots are r % that
	const int updepth = depth <= 3 ? 3 : depth <= 4 ? 4 : 6;
	t.depth = use_up ? updepth : depth;

	// The kernels fetch the descriptor from global memory, not from their parameter block: a slot
	// of a per-device ring, written in stream order ahead of the launches (the upscaling kernel copies rows
	// without a descriptor).
	unsigned slot_index = 0;
	CUtensorMap *dmap = nullptr;
	if (!use_up) {
		dmap = map_slot(&slot_index);
		if (!dmap) return cudaErrorMemoryAllocation;
		write_descriptor_kernel<<<1, 32, 0, stream>>>(dmap, map);
		cudaError_t ce = cudaGetLastError();
		if (ce != cudaSuccess) return ce;
		*launches += 1;
	}

	FastLaunch a{};
	a.map = dmap; a.dst = &dst; a.t = &t; a.n = n; a.channels = channels; a.smem_bytes = smem_total; a.stream = stream;
	VTable vt{};   // (zeroed: the kernels never read past what is filled below, but the block travels to the device whole)
	a.vt = &vt;
	dl.map = dmap; dl.dst = &dst; dl.t = &t; dl.vt = &vt; dl.n = n; dl.channels = channels; dl.smem_bytes = smem_total;
	dl.stream = stream;
	const float vscale = std::ldexp(1.0f, down::kVExp);
	ul.src = &src; ul.dst = &dst; ul.t = &t; ul.vt = &vt; ul.n = n; ul.stream = stream;
	const float vscale_up = std::ldexp(1.0f, 149 - up::kHExp) / (deep ? 65535.0f : 255.0f);
	// Groups of consecutive bands whose tables fit one parameter block; one launch per group.
	for (int yb = 0; yb < dh;) {
		int ye = yb, bands = 0;
		while (ye < dh && bands < kMaxBands && band_fits(yb, std::min(dh, ye + band_h))) {
			ye = std::min(dh, ye + band_h);
			++bands;
		}
		if (bands == 0) return cudaErrorNotSupported;
		vt.y_begin = yb;
		vt.y_end = ye;
		const int row_lo = first_row(yb), row_hi = fy.cum[ye - 1];
		vt.row_base = row_lo;
		vt.out_base = fy.variant == 0 ? fy.ybase[row_lo] : yb;
		for (int b = 0; b < bands; ++b) {
			const int y0 = yb + b * band_h, y1 = std::min(ye, y0 + band_h);
			vt.band_rlo[b] = first_row(y0);
			vt.band_rhi[b] = fy.cum[y1 - 1];
			vt.band_ys[b] = fy.variant == 0 ? fy.ybase[first_row(y0)] : y0;
			vt.band_n0[b] = fy.variant == 0 ? fy.cum[vt.band_ys[b]] - vt.band_rlo[b] + 1 : 0;
		}
		if (!use_down) {
			const int *ysrc = fy.variant == 0 || use_up ? fy.cum.data() : fy.lo.data();
			for (int y = vt.out_base; y < ye; ++y) vt.ytab[y - vt.out_base] = ysrc[y];
			// the upscaling kernel looks one output ahead; its output loop ends on this sentinel
			vt.ytab[ye - vt.out_base] = use_up || ye >= dh ? -1 : ysrc[ye];
		}
		// weight rows are re-strided from the host table's stride to the kernel's WS
		const int first = fy.variant == 0 ? row_lo : yb, last = fy.variant == 0 ? row_hi : ye - 1;
		if (use_down) {
			// slot order: the weight of row i for output y sits at y % depth; scaled by 2^kVExp
			for (int i = first; i <= last; ++i) {
				float *w = vt.wdown + (i - first) * WS;
				for (int j = 0; j < WS; ++j) w[j] = 0.0f;
				for (int j = 0; j < fy.depth; ++j) w[(fy.ybase[i] + j) % depth] = fy.wv[(size_t)i * fy.stride + j] * vscale;
				// the row loop's event flags: how many outputs this row completes, and whether the row after it is the
				// last of its ring stage
				const uint32_t bits = (uint32_t)fy.done[i] | ((i & (rsk - 1)) == rsk - 2 ? down::kEvStage : 0u);
				memcpy(&w[depth], &bits, 4);
			}
			// the two look-ahead rows behind the last one (the block is reused from launch to launch)
			for (int j = 0; j < 2 * WS; ++j) vt.wdown[(last + 1 - first) * WS + j] = 0.0f;
		} else if (use_up) {
			// slot order: source row r sits in window slot r % depth; scaled so the result lands on [0, 1]
			for (int i = first; i <= last; ++i) {
				float *w = vt.wt + (i - first) * WS;
				for (int j = 0; j < WS; ++j) w[j] = 0.0f;
				for (int j = 0; j < fy.depth; ++j) w[(fy.lo[i] + j) % updepth] = fy.wv[(size_t)i * fy.stride + j] * vscale_up;
			}
		} else {
			for (int i = first; i <= last; ++i)
				for (int j = 0; j < WS; ++j)
					vt.wt[(i - first) * WS + j] = j < fy.stride ? fy.wv[(size_t)i * fy.stride + j] : 0.0f;
		}
		a.bands = bands;
		dl.bands = bands;
		dl.overlap = yb > 0;
		ul.bands = bands;
		ul.overlap = yb > 0;
		cudaError_t e;
		if (use_down) {
			switch (channels * 2 + (deep ? 1 : 0)) {
				case 2: e = launch_down<false, 1>(dl); break;
				case 3: e = launch_down<true, 1>(dl); break;
				case 4: e = launch_down<false, 2>(dl); break;
				case 5: e = launch_down<true, 2>(dl); break;
				case 6: e = launch_down<false, 3>(dl); break;
				case 7: e = launch_down<true, 3>(dl); break;
				case 8: e = launch_down<false, 4>(dl); break;
				default: e = launch_down<true, 4>(dl); break;
			}
		} else if (use_up) {
			switch (channels * 2 + (deep ? 1 : 0)) {
				case 2: e = launch_up<false, 1>(ul); break;
				case 3: e = launch_up<true, 1>(ul); break;
				case 4: e = launch_up<false, 2>(ul); break;
				case 5: e = launch_up<true, 2>(ul); break;
				case 6: e = launch_up<false, 3>(ul); break;
				case 7: e = launch_up<true, 3>(ul); break;
				case 8: e = launch_up<false, 4>(ul); break;
				default: e = launch_up<true, 4>(ul); break;
			}
		}
		else if (fy.variant == 0) e = deep ? launch_fast_down_u16(a) : launch_fast_down_u8(a);
		else e = deep ? launch_fast_up_u16(a) : launch_fast_up_u8(a);
		if (e != cudaSuccess) return e;
		*launches += 1;
		g_last_resize_kernel = use_up ? (ul.wpx == 0 ? 7 : 5) : use_down ? (dl.da.rq > 0 ? 6 : dl.group == 8 ? 4 : 3) : 2;
		yb = ye;
	}
	return use_up ? cudaSuccess : map_slot_release(slot_index, stream);
}

}  // namespace picha_b200
