// Upscaling variant of the fast resize path; reference: src/resize.cc:66-134.  Included by the
// resize_up_*.cu instantiation units (one per channel count and depth).
//
// An upscale is the mirror image of a downscale: the expensive pass is the one that runs at the
// OUTPUT resolution, so that is the pass that has to be thread-private.  Here that is the vertical
// pass, and the order is the reference's own -- horizontal first:
//   * A thread owns 4 consecutive output pixels (4 x C channel values) of the tile and keeps a window
//     of DEPTH horizontally-filtered rows of them in registers (DEPTH x 4C floats).
//   * Horizontal pass, once per SOURCE row: the thread reads the <= WPX source pixels its four
//     outputs touch straight from the staged row (bytes as subnormal floats, see resize_down.cuh;
//     4-channel pixels as whole words, the others value by value with zero-extending loads),
//     and multiplies them with a dense 4 x WPX weight block it holds in registers (the block is the
//     same for every row; entries outside a column's taps are zero).  No intermediate ever goes
//     through shared memory and there is no CTA-wide barrier in the kernel.
//   * Vertical pass, once per OUTPUT row: DEPTH FMAs per value with warp-uniform weights from the
//     constant bank (laid out by window slot: row r lives in slot r % DEPTH, so the code is the same
//     for every row), a saturating last FMA, two more instructions per value to pack, and 16/32-byte
//     stores of the thread's own pixels -- coalesced across the warp.
//   * Wide-window variant (WPX = 0): when a thread's 4 outputs touch more than 8 source pixels -- a vertical upscale
//     combined with a horizontal downscale, or very wide filters -- the horizontal pass walks the thread's window
//     pixel by pixel (up to 64), reading the pixel from the staged row and its four weights (one per output column,
//     zero where it is not a tap: FastTables::xwide, built with the plan) through the read-only cache; everything
//     else is the same.  (These shapes took the bit-exact kernel before: 6 - 9 % of the roofline.)
//   * Source rows arrive through a ring of plain bulk copies (cp.async.bulk, one per row, mbarrier
//     completion); stages are handed back like in the downscaling kernel.
#ifndef PICHA_B200_RESIZE_UP_CUH
#define PICHA_B200_RESIZE_UP_CUH

#include "resize_down.cuh"

// Packed window and fma.rn.f32x2 in both passes (see resize_down.cuh): the instantiation units of the
// 3-channel formats switch it off -- their horizontal pass has no channel pairs with a common weight, and
// pairing scalars up afterwards costs more than the vertical pass gains (measured).
#ifndef PICHA_UP_PACKED
#define PICHA_UP_PACKED 1
#endif

namespace picha_b200 {
namespace up {

using fast::lds;
using fast::smem;
using fast::smem_u32;
using fast::VTable;

constexpr int NT = 64;        // threads per CTA
constexpr int NPX = 4;        // output pixels per thread
constexpr int TILE = NT * NPX;   // output pixels per tile
constexpr int NS = 3;         // ring stages
constexpr int RS = 8;         // source rows per stage
constexpr int RS_WIDE = 2;    // the same for the wide-window variant: its rows are ten times the work and several times
                              // the bytes, so a shallow ring keeps 6 CTAs per SM (8-row stages: 3)
constexpr int kHExp = 120;    // horizontal weights are scaled by 2^kHExp
constexpr int kMaxDepth = 6;

struct UpArgs {
	int win_bytes;     // bytes of a source row staged per tile (multiple of 16)
	float hscale;      // 2^kHExp
	int window;        // wide-window variant: source pixels a thread walks per row (FastTables::xwide_window)
	FuseArgs fuse;     // resize, then convert: the pack stage stores the destination's pixel format
};

__host__ __device__ inline int smem_bytes(int win_bytes, bool wide = false) { return NS * (wide ? RS_WIDE : RS) * win_bytes + 2 * NS * 8; }

__device__ __forceinline__ void bulk_load(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar) {
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

struct RingState {
	int stage, slot, nstages;
	uint32_t parity;
};

// One stage: the rows [row0, row0 + RS) that exist and belong to the band, copy_bytes of each.
template <int RS>
__device__ __forceinline__ void issue_stage(uint32_t dst, uint32_t bar, const uint8_t *src, long long stride, int row0,
                                            int row_end, int win_bytes, int copy_bytes) {
	int rows = row_end - row0;
	if (rows > RS) rows = RS;
	fast::mbar_expect_tx_a(bar, rows * copy_bytes);
	// no loop here: this runs under a per-thread condition, and a loop there makes the compiler treat the
	// caller's loop state as divergent (the vertical weights would leave the uniform registers)
#pragma unroll
	for (int i = 0; i < RS; ++i)
		if (i < rows) bulk_load(dst + i * win_bytes, src + (row0 + i) * stride, copy_bytes, bar);
}

// FUSED: resize, then convert (picha_b200_resize_convert); kernels of their own, see resize_down.cuh
template <int DEPTH, bool DEEP, int WPX_, int C, bool FUSED>
__global__ void __launch_bounds__(NT, 6)
resize_up_kernel(DevBatch src, DevBatch dst, FastTables t, const __grid_constant__ VTable vt, UpArgs ua) {
	constexpr int BPP = C * Depth<DEEP>::bytes;
	constexpr int NV = NPX * C;                  // channel values a thread owns
	constexpr int WS = DEPTH <= 4 ? 4 : 8;       // vertical weights per table row
	// registers per staged pixel: 4-channel pixels travel as 32-bit words, the others as one zero-extended value each
	constexpr int WPP = C == 4 ? (DEEP ? 2 : 1) : C;
	constexpr bool WIDE = WPX_ == 0;             // the wide-window variant: no register block of weights, no raw-pixel prefetch
	constexpr int RS = WIDE ? RS_WIDE : up::RS;  // source rows per ring stage
	constexpr int WPX = WIDE ? 1 : WPX_;
	const int tid = threadIdx.x;
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");

	const int x0 = blockIdx.x * TILE;
	const int sx0 = t.xfirst[x0] / t.align_px * t.align_px;   // tile origin in the source row: 16-byte aligned
	const int band = blockIdx.y;
	const int y0 = vt.y_begin + band * t.band_h, y1 = min(vt.y_end, y0 + t.band_h);
	const int rlo = vt.band_rlo[band], rhi = vt.band_rhi[band];
	const uint8_t *const simg = src.base + (long long)blockIdx.z * src.step + (long long)sx0 * BPP;
	// never read past the row's own bytes (the last row of the last image ends the allocation)
	const int copy_bytes = min(ua.win_bytes, (src.stride - sx0 * BPP) & ~15);

	uint32_t sbase = smem_u32(smem);
	asm volatile("" : "+r"(sbase));
	const uint32_t ring = sbase, bars = sbase + NS * RS * ua.win_bytes;
	const int stage_bytes = RS * ua.win_bytes;

	RingState rs;
	rs.stage = -1; rs.slot = NS - 1; rs.parity = 1;
	rs.nstages = (rhi - rlo) / RS + 1;
	if (tid == 0) {
		for (int i = 0; i < NS; ++i) {
			down::mbar_init_a(bars + 8 * i, 1);
			down::mbar_init_a(bars + 8 * (NS + i), NT / 32);
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
#pragma unroll
		for (int k = 0; k < NS; ++k)
			if (k < rs.nstages) issue_stage<RS>(ring + k * stage_bytes, bars + 8 * k, simg, src.stride, rlo + k * RS, rhi + 1, ua.win_bytes, copy_bytes);
	}

	// ---- this thread's four output pixels: window start and dense weight block ---------------------
	const int px0 = x0 + NPX * tid;
	const bool any = px0 < dst.width;
	const int ws = any ? t.xfirst[px0] - sx0 : 0;           // first source pixel of the thread's window (tile-relative)
	float wh[NPX][WPX];
#pragma unroll
	for (int p = 0; p < (WIDE ? 0 : NPX); ++p) {
		const int px = px0 + p;
		const bool live = px < dst.width;
		const int f = live ? t.xfirst[px] - sx0 : 0, c = live ? t.xcount[px] : 0;
#pragma unroll
		for (int j = 0; j < WPX; ++j) {
			const int k = ws + j - f;
			wh[p][j] = (k >= 0 && k < c) ? t.xw[(long long)px * t.xstride + k] * ua.hscale : 0.0f;
		}
	}
	// byte offset of window pixel j in a staged row; positions behind the staged bytes carry zero weights
	uint32_t off[WPX];
#pragma unroll
	for (int j = 0; j < WPX; ++j) off[j] = min((ws + j) * BPP, copy_bytes - BPP);
	__syncthreads();   // barrier initialisation is visible

#if PICHA_UP_PACKED
	// the window as packed pairs: both passes use fma.rn.f32x2 with a broadcast weight (see resize_down.cuh)
	down::u64 win[DEPTH][NV / 2];
#pragma unroll
	for (int k = 0; k < DEPTH; ++k)
#pragma unroll
		for (int i = 0; i < NV / 2; ++i) win[k][i] = 0;
#else
	float win[DEPTH][NV];
#pragma unroll
	for (int k = 0; k < DEPTH; ++k)
#pragma unroll
		for (int i = 0; i < NV; ++i) win[k][i] = 0.0f;
#endif

	uint32_t rowaddr = 0;   // shared address of the next row to read
	int fleft = 0;
	auto advance = [&]() {
		const int prev = rs.slot;
		if (rs.stage >= 0 && rs.stage + NS < rs.nstages) {
			__syncwarp();
			if ((tid & 31) == 0) {
				uint32_t pending;
				asm volatile(
					"{\n\t.reg .b64 st;\n\t"
					"mbarrier.arrive.shared::cta.b64 st, [%1];\n\t"
					"mbarrier.pending_count.b64 %0, st;\n\t}"
					: "=r"(pending) : "r"(bars + 8 * (NS + prev)) : "memory");
				if (pending == 1)
					issue_stage<RS>(ring + prev * stage_bytes, bars + 8 * prev, simg, src.stride, rlo + (rs.stage + NS) * RS, rhi + 1,
					            ua.win_bytes, copy_bytes);
			}
		}
		++rs.stage;
		if (++rs.slot == NS) { rs.slot = 0; rs.parity ^= 1; }
		fast::mbar_wait_a(bars + 8 * rs.slot, rs.parity);
		fleft = RS;
		rowaddr = ring + rs.slot * stage_bytes;
	};
	uint32_t curaddr = 0;   // (wide-window variant) shared address of the row being processed
	auto load_row = [&](uint32_t (&raw)[WPX * WPP]) {
		curaddr = rowaddr;
#pragma unroll
		for (int j = 0; j < (WIDE ? 0 : WPX); ++j) {
			if (C == 4 && DEEP) {
				const uint2 v = lds<uint2>(rowaddr + off[j]);
				raw[(WPP * j) % (WPX * WPP)] = v.x; raw[(WPP * j + 1) % (WPX * WPP)] = v.y;
			} else if (C == 4) {
				raw[j] = (uint32_t)lds<int>(rowaddr + off[j]);
			} else {
#pragma unroll
				for (int c = 0; c < C; ++c) {
					uint32_t v;
					if (DEEP) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(rowaddr + off[j] + 2 * c) : "memory");
					else asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(rowaddr + off[j] + c) : "memory");
					raw[(C * j + c) % (WPX * WPP)] = v;
				}
			}
		}
		rowaddr += ua.win_bytes;
		--fleft;
	};

	uint8_t *const dcol = dst.base + (long long)blockIdx.z * dst.step + (long long)px0 * (FUSED ? pixel_bytes(ua.fuse.dst_pixel) : BPP);
	const int npx = any ? min(NPX, dst.width - px0) : 0;

	uint32_t raw[WPX * WPP];
	advance();
	load_row(raw);
	// Loop state that indexes the constant bank (widx, yidx) is kept apart from anything per-thread: used
	// in a per-thread address computation it would be held in a vector register, and the weight loads
	// would follow it there.
	int yidx = y0 - vt.out_base + 1;                // next entry of ytab
	int widx = (y0 - vt.out_base) * WS;
	int need = vt.ytab[yidx - 1];                   // last source row output y needs
	uint8_t *drow = dcol + (long long)y0 * dst.stride;
#if PICHA_UP_PACKED
	auto process_row = [&](int r, down::u64 (&wrow)[NV / 2]) {
#else
	auto process_row = [&](int r, float (&wrow)[NV]) {
#endif
		// ---- horizontal pass of source row r -------------------------------------------------------
		float h[NV];   // accumulated apart from the window: the window's old row may still be read by nobody, but
		               // the prefetch below wants the raw words free early
#if PICHA_UP_PACKED
		down::u64 h2[NV / 2];
#endif
		if constexpr (WIDE) {
			// walk the window: pixel j of it carries weight w4[p] into output column p (zero where it is no tap)
#pragma unroll
			for (int i = 0; i < NV; ++i) h[i] = 0.0f;
			const float4 *wt = reinterpret_cast<const float4 *>(t.xwide) + (long long)(any ? px0 / NPX : 0) * ua.window;
			const uint32_t base = curaddr + ws * BPP, last = curaddr + copy_bytes - BPP;
#pragma unroll 4
			for (int j = 0; j < ua.window; ++j) {
				const float4 w4 = __ldg(wt + j);
				const uint32_t a = min(base + j * BPP, last);   // (pixels behind the staged bytes carry zero weights)
				float u[C];
				if (C == 4 && DEEP) {
					const uint2 v = lds<uint2>(a);
					u[0] = __uint_as_float(v.x & 0xFFFFu); u[1 % C] = __uint_as_float(v.x >> 16);
					u[2 % C] = __uint_as_float(v.y & 0xFFFFu); u[3 % C] = __uint_as_float(v.y >> 16);
				} else if (C == 4) {
					const uint32_t v = (uint32_t)lds<int>(a);
#pragma unroll
					for (int c = 0; c < C; ++c) u[c] = __uint_as_float(__byte_perm(v, 0, 0x4440 + c));
				} else {
#pragma unroll
					for (int c = 0; c < C; ++c) {
						uint32_t v;
						if (DEEP) asm volatile("ld.shared.u16 %0, [%1];" : "=r"(v) : "r"(a + 2 * c) : "memory");
						else asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(a + c) : "memory");
						u[c] = __uint_as_float(v);
					}
				}
				const float wp[NPX] = {w4.x, w4.y, w4.z, w4.w};
#pragma unroll
				for (int p = 0; p < NPX; ++p)
#pragma unroll
					for (int c = 0; c < C; ++c) h[C * p + c] = fmaf(wp[p], u[c], h[C * p + c]);
			}
		}
#pragma unroll
		for (int j = 0; j < (WIDE ? 0 : WPX); ++j) {
			float u[C];
			if (C == 4 && DEEP) {
				u[0] = __uint_as_float(raw[(2 * j) % (WPX * WPP)] & 0xFFFFu);
				u[1 % C] = __uint_as_float(raw[(2 * j) % (WPX * WPP)] >> 16);
				u[2 % C] = __uint_as_float(raw[(2 * j + 1) % (WPX * WPP)] & 0xFFFFu);
				u[3 % C] = __uint_as_float(raw[(2 * j + 1) % (WPX * WPP)] >> 16);
			} else if (C == 4) {
#pragma unroll
				for (int c = 0; c < C; ++c) u[c] = __uint_as_float(__byte_perm(raw[j % (WPX * WPP)], 0, 0x4440 + c));
			} else {
#pragma unroll
				for (int c = 0; c < C; ++c) u[c] = __uint_as_float(raw[(C * j + c) % (WPX * WPP)]);
			}
#if PICHA_UP_PACKED
			if (C % 2 == 0) {
				// channel pairs of a pixel share the weight: packed MACs straight into the pairs of h
#pragma unroll
				for (int p = 0; p < NPX; ++p)
#pragma unroll
					for (int c = 0; c < C; c += 2) {
						const down::u64 uu = down::pair(u[c], u[(c + 1) % C]), ww = down::pair(wh[p][j], wh[p][j]);
						if (j == 0) asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(h2[(C * p + c) / 2]) : "l"(uu), "l"(ww));
						else down::ffma2(h2[(C * p + c) / 2], uu, ww);
					}
			} else
#endif
			{
#pragma unroll
				for (int p = 0; p < NPX; ++p)
#pragma unroll
					for (int c = 0; c < C; ++c) h[C * p + c] = j == 0 ? wh[p][0] * u[c] : fmaf(wh[p][j], u[c], h[C * p + c]);
			}
		}
		// the next row's pixels, fetched while this row's outputs are computed
		if (r < rhi) {
			if (fleft == 0) advance();
			load_row(raw);
		}
#if PICHA_UP_PACKED
#pragma unroll
		for (int i = 0; i < NV / 2; ++i) wrow[i] = (C % 2 == 0 && !WIDE) ? h2[i] : down::pair(h[2 * i], h[2 * i + 1]);
#else
#pragma unroll
		for (int i = 0; i < NV; ++i) wrow[i] = h[i];
#endif

		// ---- vertical pass: every output row whose last source row this was ----------------------------
		// (The loop deliberately has no "y < y1" test: with a second exit condition the compiler holds the
		// loop state in vector registers and the weights below stop being uniform operands.  The table ends
		// with a sentinel; at a band boundary inside a launch an output row that shares its last source row
		// with this band's last output is produced here as well as by the next band -- same values.)
		while (need == r) {
			const float *w = vt.wt + widx;
#if PICHA_UP_PACKED
			down::u64 part = 0;
#endif
			uint32_t pv[NV];
#pragma unroll
			for (int i = 0; i < NV; ++i) {
#if PICHA_UP_PACKED
				// all but the last tap as packed MACs on the pair (done once per pair, at its even member); the
				// last tap stays scalar: it carries the saturation, which the packed FMA does not have
				float a, last;
				if ((i & 1) == 0) {
					asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(part) : "l"(win[0][i >> 1]), "l"(down::pair(w[0], w[0])));
#pragma unroll
					for (int k = 1; k < DEPTH - 1; ++k) down::ffma2(part, win[k][i >> 1], down::pair(w[k], w[k]));
				}
				{
					float p0, p1, l0, l1;
					down::unpair(part, p0, p1);
					down::unpair(win[DEPTH - 1][i >> 1], l0, l1);
					a = (i & 1) ? p1 : p0;
					last = (i & 1) ? l1 : l0;
				}
				a = __saturatef(fmaf(w[DEPTH - 1], last, a));
#else
				float a = w[0] * win[0][i];
#pragma unroll
				for (int k = 1; k < DEPTH - 1; ++k) a = fmaf(w[k], win[k][i], a);
				a = __saturatef(fmaf(w[DEPTH - 1], win[DEPTH - 1][i], a));
#endif
				// floor(a * max + 0.5) in the low mantissa bits (round half up like the reference: ties are common
				// with box and triangle weights, so round-to-nearest-even in a single FMA will not do)
				pv[i] = __float_as_uint(__fadd_rd(fmaf(a, Depth<DEEP>::maxv, 0.5f), 8388608.0f));
			}
			// the thread's 4 pixels are NV / (DEEP ? 2 : 4) whole 32-bit words of the row
			constexpr int NW = DEEP ? NV / 2 : NV / 4;
			uint32_t q[NW];
#pragma unroll
			for (int i = 0; i < NW; ++i) {
				if (DEEP) {
					q[i] = __byte_perm(pv[2 * i], pv[2 * i + 1], 0x5410);
				} else {
					const uint32_t lo = __byte_perm(pv[4 * i], pv[4 * i + 1], 0x0040), hi = __byte_perm(pv[4 * i + 2], pv[4 * i + 3], 0x0040);
					q[i] = __byte_perm(lo, hi, 0x5410);
				}
			}
			if (FUSED) {
				// resize, then convert: each of the thread's pixels goes through the reference's conversion
#pragma unroll
				for (int p = 0; p < NPX; ++p)
					if (p < npx) convert_store<C, DEEP>(drow + p * pixel_bytes(ua.fuse.dst_pixel), pv + C * p, ua.fuse);
			} else if (npx == NPX) {
				if (NW % 4 == 0) {
#pragma unroll
					for (int i = 0; i < NW / 4; ++i) reinterpret_cast<uint4 *>(drow)[i] = make_uint4(q[4 * i], q[4 * i + 1], q[4 * i + 2], q[4 * i + 3]);
				} else if (NW % 2 == 0) {
#pragma unroll
					for (int i = 0; i < NW / 2; ++i) reinterpret_cast<uint2 *>(drow)[i] = make_uint2(q[2 * i], q[2 * i + 1]);
				} else {
#pragma unroll
					for (int i = 0; i < NW; ++i) reinterpret_cast<uint32_t *>(drow)[i] = q[i];
				}
			} else {
				// the last thread of a row whose width is not a multiple of 4: value by value
#pragma unroll
				for (int i = 0; i < NV; ++i)
					if (i < npx * C) {
						if (DEEP) reinterpret_cast<uint16_t *>(drow)[i] = (uint16_t)pv[i];
						else drow[i] = (uint8_t)pv[i];
					}
			}
			need = vt.ytab[yidx];
			++yidx;
			widx += WS;
			drow += dst.stride;
		}
	};
	// Window slots are static: the row loop is unrolled DEPTH times and starts at a multiple of DEPTH
	// (source row r lives in slot r % DEPTH, which is also how the host lays out the weights).
	for (int rb = rlo - rlo % DEPTH; rb <= rhi; rb += DEPTH) {
#pragma unroll
		for (int k = 0; k < DEPTH; ++k) {
			const int r = rb + k;
			if (r >= rlo && r <= rhi) process_row(r, win[k]);
		}
	}

	// Never leave with a copy still in flight: wait for every stage that was issued.
	for (int k = rs.stage + 1; k < rs.nstages && k < rs.stage + NS; ++k) {
		if (++rs.slot == NS) { rs.slot = 0; rs.parity ^= 1; }
		fast::mbar_wait_a(bars + 8 * rs.slot, rs.parity);
	}
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

struct UpLaunch {
	const DevBatch *src, *dst;
	const FastTables *t;
	const VTable *vt;
	UpArgs ua;
	int n, bands, wpx, overlap;   // wpx = 0: the wide-window variant
	cudaStream_t stream;
};

template <int DEPTH, bool DEEP, int WPX, int C, bool FUSED> cudaError_t launch_one(const UpLaunch &a) {
	auto kern = resize_up_kernel<DEPTH, DEEP, WPX, C, FUSED>;
	const int smem_total = smem_bytes(a.ua.win_bytes, WPX == 0);
	static SmemGrant granted;   // (per instantiation: see grow_dynamic_smem)
	cudaError_t e = grow_dynamic_smem(reinterpret_cast<const void *>(kern), smem_total, &granted);
	if (e != cudaSuccess) return e;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((a.dst->width + TILE - 1) / TILE, a.bands, a.n);
	cfg.blockDim = dim3(NT);
	cfg.dynamicSmemBytes = smem_total;
	cfg.stream = a.stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = a.overlap ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kern, *a.src, *a.dst, *a.t, *a.vt, a.ua);
}

template <int DEPTH, bool DEEP, int C> cudaError_t launch_wpx(const UpLaunch &a) {
	if (a.wpx == 0) return launch_one<DEPTH, DEEP, 0, C, false>(a);                 // (wide window: no converting kernels)
	if (a.ua.fuse.dst_pixel >= 0) return launch_one<DEPTH, DEEP, 8, C, true>(a);   // (converting kernels: the wider window only)
	if (a.wpx <= 6) return launch_one<DEPTH, DEEP, 6, C, false>(a);
	if (a.wpx <= 8) return launch_one<DEPTH, DEEP, 8, C, false>(a);
	return cudaErrorNotSupported;
}

template <bool DEEP, int C> cudaError_t launch_depth(const UpLaunch &a) {
	const int d = a.t->depth;
	if (d <= 3) return launch_wpx<3, DEEP, C>(a);
	if (d <= 4) return launch_wpx<4, DEEP, C>(a);
	if (d <= 6) return launch_wpx<6, DEEP, C>(a);
	return cudaErrorNotSupported;
}

}  // namespace up

using up::UpLaunch;
// One definition per instantiation unit (channels 1..4, 8- and 16-bit).
template <bool DEEP, int C> cudaError_t launch_up(const UpLaunch &a);

}  // namespace picha_b200
#endif
