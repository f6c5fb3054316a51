// The generic (first-generation) kernel of the fast resize path, and the shared-memory / mbarrier /
// TMA helpers the newer kernels (resize_down.cuh, resize_up.cuh) reuse.  Work split: one CTA (128
// threads) produces a tile of `tile_w` x `band_h` output pixels of one image from TMA-staged source rows.
// Pass 1 (vertical, in registers): a thread owns 8 consecutive channel values of the source row and keeps
// either a ring of DEPTH accumulators (kDown) or a window of DEPTH unpacked source rows (kUp), rotated
// by loop unrolling so every register index is static; weights come from the constant bank.  Pass 2
// (horizontal, from shared memory): lanes of a quarter-warp walk different rows of the group at the same
// x; results are packed into a shared-memory tile and leave as 16-byte coalesced row segments.
// Included by the four instantiation units (resize_fast_{down,up}_{u8,u16}.cu), which compile in parallel.
#ifndef PICHA_B200_RESIZE_FAST_CUH
#define PICHA_B200_RESIZE_FAST_CUH
#include <cuda.h>

#include "kernels.h"
#include "pixel.cuh"

namespace picha_b200 {

namespace fast {

constexpr int NT = kFastThreads;
constexpr int NV = kFastValuesPerThread;
#ifndef PICHA_FAST_RS
#define PICHA_FAST_RS 8
#endif
#ifndef PICHA_FAST_NS
#define PICHA_FAST_NS 2
#endif
#ifndef PICHA_FAST_G
#define PICHA_FAST_G 4
#endif

constexpr int RS = PICHA_FAST_RS;  // source rows per TMA stage (u8; 16-bit rows are twice as long: half as many)
constexpr int NS = PICHA_FAST_NS;  // stages in the ring
constexpr int G = PICHA_FAST_G;    // output rows per pass-2 group (4 or 8)
constexpr int RPT = G / 4;         // output rows a pass-2 thread produces (they share the x weights)
static_assert(G == 4 || G == 8, "pass 2 maps 4 rows to the low lane bits");
constexpr int TMPS = NT * NV + 4;  // floats per intermediate row (+4: rows land 4 banks apart)

// Vertical tables of one launch, passed by value as a __grid_constant__ kernel parameter (the
// constant bank holds 32,764 bytes of parameters since CUDA 12.1).  A launch covers the output
// rows [y_begin, y_end) in bands of band_h rows, one band per blockIdx.y.
constexpr int kMaxBands = 64;
constexpr int kYtabMax = 1024;
constexpr int kWtMax = 5632;
struct alignas(16) VTable {
	int y_begin, y_end;
	int row_base;               // kDown: source row of wt[0]
	int out_base;               // output row of ytab[0] (and, kUp, of wt[0])
	int band_rlo[kMaxBands];    // first source row the band touches
	int band_rhi[kMaxBands];    // last one
	int band_ys[kMaxBands];     // kDown: output row that is open when row band_rlo arrives (<= the band's first row)
	int band_n0[kMaxBands];     // resize_down.cuh: source rows from band_rlo that complete output band_ys
	union {
		struct {
			int ytab[kYtabMax];         // kDown: cum[out_base + i]; kUp: lo[out_base + i]
			float wt[kWtMax];           // kDown: weights of source row row_base + i / WS; kUp: of output out_base + i / WS
		};
		// the downscaling kernel (resize_down.cuh) has no per-output table: its rows' weights and event flags use both
		float wdown[kYtabMax + kWtMax];
	};
};
static_assert(sizeof(VTable) <= 28 * 1024, "kernel parameters are limited to 32,764 bytes");

// Source rows per TMA stage: 16-bit rows are twice as long, so half as many.
__host__ __device__ constexpr int stage_rows(bool deep) { return deep ? RS / 2 : RS; }

struct SmemLayout {
	int row_bytes;    // bytes per staged source row
	int ring, tmp, out, out_stride, xw, xf, xc, bars, total;
};

__host__ __device__ inline SmemLayout smem_layout(bool deep, int tile_w, int bpp, int xstride) {
	SmemLayout L;
	L.row_bytes = NT * NV * (deep ? 2 : 1);
	L.ring = 0;
	L.tmp = L.ring + NS * stage_rows(deep) * L.row_bytes;
	L.out = L.tmp + G * TMPS * 4 + 256;   // 64 zeroed floats: padded taps of the last row may read past it
	L.out_stride = ((tile_w * bpp + 127) / 128) * 128 + 16;
	L.xw = L.out + G * L.out_stride;
	L.xf = L.xw + tile_w * xstride * 4;
	L.xc = L.xf + tile_w * 4;
	L.bars = ((L.xc + tile_w * 4 + 7) / 8) * 8;
	L.total = L.bars + NS * 8;
	return L;
}

extern __shared__ __align__(128) uint8_t smem[];

// Shared-memory locations travel as 32-bit shared-window addresses (base of the CTA's dynamic
// shared memory + byte offset) and are accessed with explicit ld.shared / st.shared: through C++
// pointers the compiler re-derives the window base from SR_CgaCtaId (an S2UR with scoreboard
// latency) in front of every access of the row loop.
template <typename T> __device__ __forceinline__ T lds(uint32_t addr);
template <> __device__ __forceinline__ uint4 lds<uint4>(uint32_t addr) {
	uint4 v;
	asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(addr) : "memory");
	return v;
}
template <> __device__ __forceinline__ uint2 lds<uint2>(uint32_t addr) {
	uint2 v;
	asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(v.x), "=r"(v.y) : "r"(addr) : "memory");
	return v;
}
template <> __device__ __forceinline__ float4 lds<float4>(uint32_t addr) {
	float4 v;
	asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
	return v;
}
template <> __device__ __forceinline__ float2 lds<float2>(uint32_t addr) {
	float2 v;
	asm volatile("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
	return v;
}
template <> __device__ __forceinline__ float lds<float>(uint32_t addr) {
	float v;
	asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
	return v;
}
template <> __device__ __forceinline__ int lds<int>(uint32_t addr) {
	int v;
	asm volatile("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ void sts(uint32_t addr, const float4 &v) {
	asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts(uint32_t addr, float v) {
	asm volatile("st.shared.f32 [%0], %1;" ::"r"(addr), "f"(v) : "memory");
}
__device__ __forceinline__ void sts(uint32_t addr, int v) {
	asm volatile("st.shared.s32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
	uint32_t ok;
	uint32_t spins = 0;
	do {
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
		// A copy that never lands is a bug in this file, not a condition to wait out: fail the launch
		// (cudaErrorLaunchFailure reaches the caller as PICHA_B200_ERR_CUDA) instead of hanging the GPU.
		if (!ok && ++spins > (1u << 24)) __trap();
	} while (!ok);
}
// One box of a 3-D tensor (words, rows, images) into shared memory.
__device__ __forceinline__ void tma_load_3d(void *dst, const CUtensorMap *map, uint64_t *bar, int x, int y, int z) {
	asm volatile(
		"cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
		::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(x), "r"(y), "r"(z) : "memory");
}

// The same on 32-bit shared-window addresses, for the out-of-line stage advance below.
__device__ __forceinline__ void mbar_expect_tx_a(uint32_t bar, uint32_t bytes) {
	asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait_a(uint32_t bar, uint32_t parity) {
	uint32_t ok, spins = 0;
	do {
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(ok) : "r"(bar), "r"(parity) : "memory");
		if (!ok && ++spins > (1u << 24)) __trap();   // a copy that never lands is a bug here: fail, do not hang
	} while (!ok);
}
__device__ __forceinline__ void tma_load_3d_a(uint32_t dst, const CUtensorMap *map, uint32_t bar, int x, int y, int z) {
	asm volatile(
		"cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
		::"r"(dst), "l"(map), "r"(bar), "r"(x), "r"(y), "r"(z) : "memory");
}

// The same box, only as far as L2: issued a few stages ahead of the copy into shared memory, so that copy finds its
// rows in L2 (the ring in shared memory is too shallow to cover HBM latency under load; L2 is not).
__device__ __forceinline__ void tma_prefetch_3d_a(const CUtensorMap *map, int x, int y, int z) {
	asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global.tile [%0, {%1, %2, %3}];" ::"l"(map), "r"(x), "r"(y), "r"(z) : "memory");
}

// Stage transition of the ring, out of line: it runs once per 8 rows, and inlined into every copy of
// the unrolled row body it would triple the size of the hot loop (instruction-cache misses showed
// up as the top stall).  Issues stage `issue` (if >= 0) into its slot and waits for `wait_bar`.
template <bool DEEP>
__device__ __noinline__ void ring_advance(const CUtensorMap *map, uint32_t ring, uint32_t bars, int issue, int word0,
                                          int row0, int img, uint32_t wait_bar, uint32_t parity, int tid) {
	constexpr int RSK = stage_rows(DEEP);
	constexpr int BOXES = DEEP ? 2 : 1;          // TMA boxes are at most 256 elements wide
	constexpr int ROW_BYTES = NT * NV * (DEEP ? 2 : 1);
	__syncthreads();                             // every thread has finished the previous stage
	if (tid == 0 && issue >= 0) {
		const uint32_t bar = bars + 8 * (issue % NS);
		mbar_expect_tx_a(bar, RSK * ROW_BYTES);
		const uint32_t d = ring + (issue % NS) * RSK * ROW_BYTES;
#pragma unroll
		for (int b = 0; b < BOXES; ++b) tma_load_3d_a(d + b * RSK * 1024, map, bar, word0 + b * 256, row0 + issue * RSK, img);
	}
	mbar_wait_a(wait_bar, parity);
}

// ---- unpack: exact float(v) * (1/max) in one FMA ----------------------------------------------
// 0x4B000000 | v is the float 2^23 + v; fma(2^23 + v, inv, -2^23*inv) rounds the exact product
// v*inv once, which is the reference's float(v) * inv (src/picha.h:98-105).
// `magic` (0x4B000000) and `inv` are passed in as registers the caller made opaque to the compiler:
// as literals they are rematerialised with two extra instructions in every row body.
template <bool DEEP> __device__ __forceinline__ void unpack8(const uint32_t *w, float *u, uint32_t magic, float inv) {
	constexpr float bias = -8388608.0f * Depth<DEEP>::inv;
	if (DEEP) {
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			u[2 * i] = fmaf(__uint_as_float(__byte_perm(w[i], magic, 0x7410)), inv, bias);
			u[2 * i + 1] = fmaf(__uint_as_float(__byte_perm(w[i], magic, 0x7432)), inv, bias);
		}
	} else {
#pragma unroll
		for (int i = 0; i < 2; ++i) {
			u[4 * i + 0] = fmaf(__uint_as_float(__byte_perm(w[i], magic, 0x7440)), inv, bias);
			u[4 * i + 1] = fmaf(__uint_as_float(__byte_perm(w[i], magic, 0x7441)), inv, bias);
			u[4 * i + 2] = fmaf(__uint_as_float(__byte_perm(w[i], magic, 0x7442)), inv, bias);
			u[4 * i + 3] = fmaf(__uint_as_float(__byte_perm(w[i], magic, 0x7443)), inv, bias);
		}
	}
}

// pack: floor(clamp(f * max + 0.5)).  The clamp is a saturate to [0, 1] on f itself (it folds into the
// FMA that produced f as FFMA.SAT; clamping before or after the scaling gives the same integer), and the
// float -> integer step avoids F2I (several issue cycles here): adding 2^23 with round-toward-minus-
// infinity leaves floor(t) in the low mantissa bits; byte/halfword merging with PRMT drops the exponent.
template <bool DEEP> __device__ __forceinline__ uint32_t pack_biased(float f) {
	return __float_as_uint(__fadd_rd(fmaf(__saturatef(f), Depth<DEEP>::maxv, 0.5f), 8388608.0f));
}

// One packed pixel into the shared-memory output tile (shared-window address).
template <int C, bool DEEP> __device__ __forceinline__ void store_pixel(uint32_t d, const float *acc) {
	constexpr int BPP = C * Depth<DEEP>::bytes;
	uint32_t v[C];
#pragma unroll
	for (int ch = 0; ch < C; ++ch) v[ch] = pack_biased<DEEP>(acc[ch]);
	if (BPP == 4 && !DEEP) {
		const uint32_t lo = __byte_perm(v[0], v[1 % C], 0x0040), hi = __byte_perm(v[2 % C], v[3 % C], 0x0040);
		asm volatile("st.shared.u32 [%0], %1;" ::"r"(d), "r"(__byte_perm(lo, hi, 0x5410)) : "memory");
	} else if (BPP == 8) {
		asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(d), "r"(__byte_perm(v[0], v[1 % C], 0x5410)),
		             "r"(__byte_perm(v[2 % C], v[3 % C], 0x5410)) : "memory");
	} else if (DEEP) {
#pragma unroll
		for (int ch = 0; ch < C; ++ch) asm volatile("st.shared.u16 [%0], %1;" ::"r"(d + 2 * ch), "h"((uint16_t)v[ch]) : "memory");
	} else {
#pragma unroll
		for (int ch = 0; ch < C; ++ch) asm volatile("st.shared.u8 [%0], %1;" ::"r"(d + ch), "r"(v[ch]) : "memory");
	}
}

// ---- pass 2: horizontal filter of one group of intermediate rows, pack, store --------------------
// A thread produces the output pixels (xx, g) and (xx, g + 4): the two rows share the x weights.
// Lanes: g fastest (4 rows), then 8 different xx per warp -- float4 reads of a quarter-warp fall in
// distinct banks (rows are 4 banks apart, neighbouring columns of a 4:1 downscale 16 banks apart).
struct Pass2Args {
	uint32_t sbase;        // shared-window address of the CTA's dynamic shared memory
	int tmp;               // float [G][TMPS]
	int xw;                // float [tile_w][xstride], zero padded
	int xf, xc;            // int: first source pixel (relative to the tile origin), taps
	int outt;              // bytes [G][out_stride]
	uint8_t *gbase;        // destination of the group's first row, at the tile's first column
	int xstride, out_stride, dstride, tw, ng, tid;
};

// shared-memory tile -> global, 16 bytes per thread where the destination allows it (row by row:
// no index division -- on upscales this loop moves most of the kernel's bytes)
template <int BPP> __device__ __forceinline__ void copy_out(const Pass2Args &a) {
	__syncthreads();
	const int row_bytes = a.tw * BPP;
	const bool vec = ((reinterpret_cast<uintptr_t>(a.gbase) | (uintptr_t)a.dstride) & 15) == 0;
	const int nvec = vec ? row_bytes >> 4 : 0;
	int done = nvec << 4;
	const bool word = (((uintptr_t)a.gbase | (uintptr_t)a.dstride) & 3) == 0;
	const int nw = word ? (row_bytes - done) >> 2 : 0;
	const int tail0 = done + (nw << 2);
	if (nvec >= NT / 2) {                        // wide tiles (upscales): a row at a time
		for (int g = 0; g < a.ng; ++g) {
			uint8_t *grow = a.gbase + (long long)g * a.dstride;
			const uint32_t srow = a.sbase + a.outt + g * a.out_stride;
			for (int j = a.tid; j < nvec; j += NT) reinterpret_cast<uint4 *>(grow)[j] = lds<uint4>(srow + 16 * j);
		}
	} else {                                     // narrow tiles (downscales): all rows in one sweep
		for (int i = a.tid; i < a.ng * nvec; i += NT) {
			const int g = i / nvec, j = i - g * nvec;
			reinterpret_cast<uint4 *>(a.gbase + (long long)g * a.dstride)[j] = lds<uint4>(a.sbase + a.outt + g * a.out_stride + 16 * j);
		}
	}
	for (int i = a.tid; i < a.ng * nw; i += NT) {
		const int g = i / nw, j = done + 4 * (i - g * nw);
		*reinterpret_cast<uint32_t *>(a.gbase + (long long)g * a.dstride + j) = lds<int>(a.sbase + a.outt + g * a.out_stride + j);
	}
	const int tail = row_bytes - tail0;
	for (int i = a.tid; i < a.ng * tail; i += NT) {
		const int g = i / tail, j = tail0 + (i - g * tail);
		a.gbase[(long long)g * a.dstride + j] = smem[a.outt + g * a.out_stride + j];
	}
}

template <int C, bool DEEP>
__device__ __noinline__ void pass2(Pass2Args a) {
	constexpr int BPP = C * Depth<DEEP>::bytes;
	for (int o = a.tid; o < a.tw * 4; o += NT) {
		const int g = o & 3, xx = o >> 2;
		if (g >= a.ng) continue;
		const bool two = RPT == 2 && g + 4 < a.ng;
		const int cnt = lds<int>(a.sbase + a.xc + 4 * xx);
		const uint32_t w = a.sbase + a.xw + 4 * xx * a.xstride;
		const uint32_t v0 = a.sbase + a.tmp + 4 * g * TMPS + lds<int>(a.sbase + a.xf + 4 * xx);
		const uint32_t v1 = v0 + (two ? 16 * TMPS : 0);
		float acc0[C], acc1[C];
#pragma unroll
		for (int ch = 0; ch < C; ++ch) acc0[ch] = acc1[ch] = 0.0f;
		int k = 0;
		if (C == 4) {
			for (; k + 4 <= cnt; k += 4) {
				const float4 wq = lds<float4>(w + 4 * k);
				const float wk[4] = {wq.x, wq.y, wq.z, wq.w};
#pragma unroll
				for (int j = 0; j < 4; ++j) {
					const float4 p = lds<float4>(v0 + 16 * (k + j));
					const float4 q = lds<float4>(v1 + 16 * (k + j));
					acc0[0] = fmaf(wk[j], p.x, acc0[0]); acc0[1 % C] = fmaf(wk[j], p.y, acc0[1 % C]);
					acc0[2 % C] = fmaf(wk[j], p.z, acc0[2 % C]); acc0[3 % C] = fmaf(wk[j], p.w, acc0[3 % C]);
					acc1[0] = fmaf(wk[j], q.x, acc1[0]); acc1[1 % C] = fmaf(wk[j], q.y, acc1[1 % C]);
					acc1[2 % C] = fmaf(wk[j], q.z, acc1[2 % C]); acc1[3 % C] = fmaf(wk[j], q.w, acc1[3 % C]);
				}
			}
			for (; k < cnt; ++k) {
				const float wk = lds<float>(w + 4 * k);
				const float4 p = lds<float4>(v0 + 16 * k);
				const float4 q = lds<float4>(v1 + 16 * k);
				acc0[0] = fmaf(wk, p.x, acc0[0]); acc0[1 % C] = fmaf(wk, p.y, acc0[1 % C]);
				acc0[2 % C] = fmaf(wk, p.z, acc0[2 % C]); acc0[3 % C] = fmaf(wk, p.w, acc0[3 % C]);
				acc1[0] = fmaf(wk, q.x, acc1[0]); acc1[1 % C] = fmaf(wk, q.y, acc1[1 % C]);
				acc1[2 % C] = fmaf(wk, q.z, acc1[2 % C]); acc1[3 % C] = fmaf(wk, q.w, acc1[3 % C]);
			}
		} else {
#pragma unroll 4
			for (; k < cnt; ++k) {
				const float wk = lds<float>(w + 4 * k);
#pragma unroll
				for (int ch = 0; ch < C; ++ch) {
					acc0[ch] = fmaf(wk, lds<float>(v0 + 4 * (C * k + ch)), acc0[ch]);
					acc1[ch] = fmaf(wk, lds<float>(v1 + 4 * (C * k + ch)), acc1[ch]);
				}
			}
		}
		const uint32_t d = a.sbase + a.outt + g * a.out_stride + xx * BPP;
		store_pixel<C, DEEP>(d, acc0);
		if (two) store_pixel<C, DEEP>(d + 4 * a.out_stride, acc1);
	}
	copy_out<C * Depth<DEEP>::bytes>(a);
}

// Few taps per output (upscaling): everything unrolled, weights zero-padded to XT taps, so the
// per-pixel loop and address overhead of the general path does not dominate the 4 x C FMAs per tap.
template <int C, bool DEEP, int XT>
__device__ __noinline__ void pass2_short(Pass2Args a) {
	constexpr int BPP = C * Depth<DEEP>::bytes;
#ifndef PICHA_FAST_P2_U
#define PICHA_FAST_P2_U 2
#endif
	constexpr int U = PICHA_FAST_P2_U;   // output pixels in flight per thread: their loads are issued together
	const int total = a.tw * 4;
	for (int o0 = a.tid; o0 < total; o0 += U * NT) {
		int g[U], xx[U];
		uint32_t w[U], v0[U];
		bool live[U];
#pragma unroll
		for (int u = 0; u < U; ++u) {
			const int o = o0 + u * NT;
			g[u] = o & 3;
			xx[u] = o >> 2;
			live[u] = o < total && g[u] < a.ng;
			if (!live[u]) { xx[u] = 0; g[u] = 0; }          // a valid location: computed, not stored
			w[u] = a.sbase + a.xw + 4 * xx[u] * a.xstride;
			v0[u] = a.sbase + a.tmp + 4 * g[u] * TMPS + lds<int>(a.sbase + a.xf + 4 * xx[u]);
		}
		float wk[U][XT];
#pragma unroll
		for (int u = 0; u < U; ++u)
#pragma unroll
			for (int q = 0; q < XT / 4; ++q) {
				const float4 wq = lds<float4>(w[u] + 16 * q);
				wk[u][4 * q] = wq.x; wk[u][4 * q + 1] = wq.y; wk[u][4 * q + 2] = wq.z; wk[u][4 * q + 3] = wq.w;
			}
		float acc[U][C];
#pragma unroll
		for (int u = 0; u < U; ++u)
#pragma unroll
			for (int ch = 0; ch < C; ++ch) acc[u][ch] = 0.0f;
#pragma unroll
		for (int k = 0; k < XT; ++k) {
#pragma unroll
			for (int u = 0; u < U; ++u) {
				if (C == 4) {
					const float4 p = lds<float4>(v0[u] + 16 * k);
					acc[u][0] = fmaf(wk[u][k], p.x, acc[u][0]); acc[u][1 % C] = fmaf(wk[u][k], p.y, acc[u][1 % C]);
					acc[u][2 % C] = fmaf(wk[u][k], p.z, acc[u][2 % C]); acc[u][3 % C] = fmaf(wk[u][k], p.w, acc[u][3 % C]);
				} else if (C == 2) {
					const float2 p = lds<float2>(v0[u] + 8 * k);
					acc[u][0] = fmaf(wk[u][k], p.x, acc[u][0]); acc[u][1 % C] = fmaf(wk[u][k], p.y, acc[u][1 % C]);
				} else {
#pragma unroll
					for (int ch = 0; ch < C; ++ch) acc[u][ch] = fmaf(wk[u][k], lds<float>(v0[u] + 4 * (C * k + ch)), acc[u][ch]);
				}
			}
		}
#pragma unroll
		for (int u = 0; u < U; ++u)
			if (live[u]) store_pixel<C, DEEP>(a.sbase + a.outt + g[u] * a.out_stride + xx[u] * BPP, acc[u]);
	}
	copy_out<BPP>(a);
}

// XS: 0 = general horizontal pass, 4 / 8 = unrolled pass for columns of at most that many taps.
// It is a template parameter of the kernel so that only one family of callees is linked into each
// kernel: with all of them reachable the register allocation of the row loop suffers.
template <int C, bool DEEP, int XS> __device__ __forceinline__ void pass2_any(const Pass2Args &a) {
	if (XS == 4) pass2_short<C, DEEP, 4>(a);
	else if (XS == 8) pass2_short<C, DEEP, 8>(a);
	else pass2<C, DEEP>(a);
}

template <bool DEEP, int XS> __device__ __forceinline__ void run_pass2(const Pass2Args &a, int channels) {
	__syncthreads();           // the group's intermediate rows are complete
	switch (channels) {
		case 1: pass2_any<1, DEEP, XS>(a); break;
		case 2: pass2_any<2, DEEP, XS>(a); break;
		case 3: pass2_any<3, DEEP, XS>(a); break;
		default: pass2_any<4, DEEP, XS>(a); break;
	}
	__syncthreads();           // pass 1 may overwrite the intermediate rows again
}

#ifndef PICHA_FAST_PAIR_UNROLL
#define PICHA_FAST_PAIR_UNROLL 0
#endif
#ifndef PICHA_FAST_MIN_CTAS
#define PICHA_FAST_MIN_CTAS 4
#endif
template <int VARIANT, int DEPTH, bool DEEP, int XS>
__global__ void __launch_bounds__(NT, (VARIANT == 0 && DEPTH <= 4 ? PICHA_FAST_MIN_CTAS + 1 : DEPTH <= 6 ? PICHA_FAST_MIN_CTAS : 1))
resize_fast_kernel(const CUtensorMap *__restrict__ smap, DevBatch dst, FastTables t,
                   const __grid_constant__ VTable vt, int channels) {
	constexpr int WPT = DEEP ? 4 : 2;            // 32-bit words of a source row per thread
	constexpr int RSK = stage_rows(DEEP);        // rows per ring stage
	constexpr int WS = (DEPTH + 3) & ~3;         // vertical weights per table row
	const int bpp = channels * Depth<DEEP>::bytes;
	const int tid = threadIdx.x;

	const int x0 = blockIdx.x * t.tile_w;
	const int tw = min(t.tile_w, dst.width - x0);
	const int sx0 = t.xfirst[x0] / t.align_px * t.align_px;   // tile origin: 16-byte aligned in the row (TMA box start)
	const int word0 = sx0 * bpp / 4;
	// everything below is uniform across the CTA and comes from the constant bank
	const int band = blockIdx.y;
	const int y0 = vt.y_begin + band * t.band_h, y1 = min(vt.y_end, y0 + t.band_h);
	const int rlo = vt.band_rlo[band], rhi = vt.band_rhi[band];
	const int nstages = (rhi - rlo + RSK) / RSK;

	const SmemLayout L = smem_layout(DEEP, t.tile_w, bpp, t.xstride);
	uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
	uint32_t sbase = smem_u32(smem);
	asm volatile("" : "+r"(sbase));   // keep it in a register: never re-derived

	auto issue_stage = [&](int k) {
		constexpr int BOXES = DEEP ? 2 : 1;      // TMA boxes are at most 256 elements wide
		uint64_t *bar = bars + (k % NS);
		mbar_expect_tx(bar, RSK * L.row_bytes);
		uint8_t *d = smem + L.ring + (k % NS) * RSK * L.row_bytes;
#pragma unroll
		for (int b = 0; b < BOXES; ++b) tma_load_3d(d + b * RSK * 1024, smap, bar, word0 + b * 256, rlo + k * RSK, blockIdx.z);
	};

	if (tid == 0) {
		for (int i = 0; i < NS; ++i) mbar_init(bars + i, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(smap) : "memory");   // the descriptor slot is reused (see map_slot)
		for (int k = 0; k < NS - 1 && k < nstages; ++k) issue_stage(k);
	}
	// this tile's horizontal tables -> shared memory
	for (int i = tid; i < tw * t.xstride; i += NT) sts(sbase + L.xw + 4 * i, t.xw[(long long)x0 * t.xstride + i]);
	for (int i = tid; i < tw; i += NT) {
		sts(sbase + L.xf + 4 * i, (t.xfirst[x0 + i] - sx0) * channels * 4);   // byte offset of the column's first tap in a row
		sts(sbase + L.xc + 4 * i, t.xcount[x0 + i]);
	}
	// The unrolled horizontal pass multiplies zero-padded taps with whatever lies behind a column's
	// window (row padding, rows of a partial group, the tail): make sure that is never a NaN.
	for (int i = tid; i < G * TMPS + 64; i += NT) sts(sbase + L.tmp + 4 * i, 0.0f);
	__syncthreads();   // tables and barrier initialisation are visible to every thread

	// ---- ring consumer -------------------------------------------------------------------------
	int stage = -1, slot = NS - 1, rows_left = 0;   // uniform
	uint32_t parity = 1;
	uint32_t doff = 0;                              // shared address of this thread's words in the next row
	const int thread_byte = 4 * ((DEEP ? ((tid * WPT) >> 8) * RSK * 256 : 0) + ((tid * WPT) & 255));
	auto next_stage = [&]() {
		++stage;
		if (++slot == NS) { slot = 0; parity ^= 1; }
		ring_advance<DEEP>(smap, sbase + L.ring, sbase + L.bars, stage + NS - 1 < nstages ? stage + NS - 1 : -1, word0, rlo,
		                   blockIdx.z, sbase + L.bars + 8 * slot, parity, tid);
		rows_left = RSK;
		doff = sbase + L.ring + slot * RSK * L.row_bytes + thread_byte;
	};
	// Next row of the tile for this thread.  Called one row ahead of the row being accumulated, so
	// the LDS latency is covered by this warp's own FMAs.  Past the band's last stage it does
	// nothing; inside the last stage it may read rows beyond rhi (staged, never used).
	auto fetch = [&](uint32_t (&w)[WPT]) {
		if (rows_left == 0) {
			if (stage + 1 >= nstages) return;
			next_stage();
		}
		--rows_left;
		if (DEEP) {
			const uint4 v = lds<uint4>(doff);
			w[0] = v.x; w[1] = v.y; w[2 % WPT] = v.z; w[3 % WPT] = v.w;
		} else {
			const uint2 v = lds<uint2>(doff);
			w[0] = v.x; w[1] = v.y;
		}
		doff += 1024;
	};
	uint32_t magic;
	float inv;
	asm volatile("mov.b32 %0, 0x4B000000;" : "=r"(magic));
	asm volatile("mov.f32 %0, %1;" : "=f"(inv) : "f"(Depth<DEEP>::inv));

	Pass2Args pa;
	pa.sbase = sbase; pa.tmp = L.tmp; pa.xw = L.xw; pa.xf = L.xf; pa.xc = L.xc; pa.outt = L.out;
	pa.xstride = t.xstride; pa.out_stride = L.out_stride; pa.dstride = dst.stride; pa.tw = tw; pa.tid = tid;
	uint8_t *const dtile = dst.base + (long long)blockIdx.z * dst.step + (long long)x0 * bpp;
	const uint32_t my_tmp = sbase + L.tmp + tid * NV * 4;

	auto emit_row = [&](int g, const float *v) {
		const uint32_t d = my_tmp + g * TMPS * 4;
		sts(d, make_float4(v[0], v[1], v[2], v[3]));
		sts(d + 16, make_float4(v[4], v[5], v[6], v[7]));
	};
	auto flush_group = [&](int y_end, int gcount) {      // rows [y_end - gcount, y_end) are in the group buffer
		pa.ng = gcount;
		pa.gbase = dtile + (long long)(y_end - gcount) * dst.stride;
		run_pass2<DEEP, XS>(pa, channels);
	};

	int gcount = 0;
	if (VARIANT == 0) {
		float acc[DEPTH][NV];
#pragma unroll
		for (int j = 0; j < DEPTH; ++j)
#pragma unroll
			for (int i = 0; i < NV; ++i) acc[j][i] = 0.0f;
		int r = rlo, y = vt.band_ys[band];
		int widx = (rlo - vt.row_base) * WS;                     // uniform index of row r's weights
		// One source row into the accumulator ring; slot s holds the output row that is open first.
		// `cur` holds row r's words, row r+1 is fetched into `nxt` meanwhile.
		auto row = [&](int s, const uint32_t (&cur)[WPT], uint32_t (&nxt)[WPT]) {
			fetch(nxt);
			float u[NV];
			unpack8<DEEP>(cur, u, magic, inv);
			const float *w = vt.wt + widx;                       // constant bank -> uniform registers
#pragma unroll
			for (int j = 0; j < DEPTH - 1; ++j)
#pragma unroll
				for (int i = 0; i < NV; ++i) acc[(s + j) % DEPTH][i] = fmaf(w[j], u[i], acc[(s + j) % DEPTH][i]);
			if (w[DEPTH - 1] != 0.0f) {                          // the newest output row is touched by few rows
#pragma unroll
				for (int i = 0; i < NV; ++i)
					acc[(s + DEPTH - 1) % DEPTH][i] = fmaf(w[DEPTH - 1], u[i], acc[(s + DEPTH - 1) % DEPTH][i]);
			}
			widx += WS;
		};
		uint32_t wa[WPT], wb[WPT];
		fetch(wa);
		while (y < y1) {
#pragma unroll
			for (int s = 0; s < DEPTH; ++s) {
				if (y < y1) {
					const int need = vt.ytab[y - vt.out_base];       // output y is complete after this row
					int n = need - r + 1;
					r = need + 1;
#if PICHA_FAST_PAIR_UNROLL
					for (; n >= 2; n -= 2) { row(s, wa, wb); row(s, wb, wa); }
					if (n > 0) {
						row(s, wa, wb);
#pragma unroll
						for (int i = 0; i < WPT; ++i) wa[i] = wb[i];
					}
#else
#pragma unroll 1
					for (; n > 0; --n) {                         // one copy of the row body per slot keeps the loop in the I-cache
						row(s, wa, wb);
#pragma unroll
						for (int i = 0; i < WPT; ++i) wa[i] = wb[i];
					}
#endif
					if (y >= y0) {
						emit_row(gcount, acc[s]);
						++gcount;
					}
#pragma unroll
					for (int i = 0; i < NV; ++i) acc[s][i] = 0.0f;
					++y;
					if (gcount == G || (y == y1 && gcount > 0)) {
						flush_group(y, gcount);
						gcount = 0;
					}
				}
			}
		}
	} else {
		float win[DEPTH][NV];
		int rb = vt.ytab[y0 - vt.out_base], rnext = rb;
		uint32_t pw[WPT];
		fetch(pw);
		auto load_window_row = [&](float (&dstv)[NV]) {
			if (rnext <= rhi) {
				uint32_t w[WPT];
#pragma unroll
				for (int i = 0; i < WPT; ++i) w[i] = pw[i];
				fetch(pw);
				unpack8<DEEP>(w, dstv, magic, inv);
				++rnext;
			} else {
#pragma unroll
				for (int i = 0; i < NV; ++i) dstv[i] = 0.0f;
			}
		};
#pragma unroll
		for (int k = 0; k < DEPTH; ++k) load_window_row(win[k]);
		int y = y0;
		while (y < y1) {
#pragma unroll
			for (int s = 0; s < DEPTH; ++s) {
				if (y < y1) {
					while (y < y1 && vt.ytab[y - vt.out_base] == rb) {
						const float *w = vt.wt + (y - vt.out_base) * WS;
						float o[NV];
#pragma unroll
						for (int i = 0; i < NV; ++i) o[i] = w[0] * win[s][i];
#pragma unroll
						for (int k = 1; k < DEPTH; ++k)
#pragma unroll
							for (int i = 0; i < NV; ++i) o[i] = fmaf(w[k], win[(s + k) % DEPTH][i], o[i]);
						emit_row(gcount, o);
						++gcount;
						++y;
						if (gcount == G || y == y1) {
							flush_group(y, gcount);
							gcount = 0;
						}
					}
					if (y < y1) {                       // slide the window down one source row
						load_window_row(win[s]);
						++rb;
					}
				}
			}
		}
	}

	// Never leave with a TMA load still in flight: wait for every stage that was issued.
	for (int k = stage + 1; k < nstages && k < stage + NS; ++k) {
		if (++slot == NS) { slot = 0; parity ^= 1; }
		mbar_wait(bars + slot, parity);
	}
}

struct FastLaunch {
	const CUtensorMap *map;
	const DevBatch *dst;
	const FastTables *t;
	const VTable *vt;
	int n, channels, smem_bytes, bands;
	cudaStream_t stream;
};

template <int VARIANT, int DEPTH, bool DEEP, int XS> cudaError_t launch_one(const FastLaunch &a) {
	auto kern = resize_fast_kernel<VARIANT, DEPTH, DEEP, XS>;
	static SmemGrant granted;   // (per instantiation: see grow_dynamic_smem)
	cudaError_t e = grow_dynamic_smem(reinterpret_cast<const void *>(kern), a.smem_bytes, &granted);
	if (e != cudaSuccess) return e;
	dim3 grid((a.dst->width + a.t->tile_w - 1) / a.t->tile_w, a.bands, a.n);
	kern<<<grid, NT, a.smem_bytes, a.stream>>>(a.map, *a.dst, *a.t, *a.vt, a.channels);
	return cudaGetLastError();
}

template <int VARIANT, int DEPTH, bool DEEP> cudaError_t launch_xs(const FastLaunch &a) {
	if (a.t->xshort == 4) return launch_one<VARIANT, DEPTH, DEEP, 4>(a);
	if (a.t->xshort == 8) return launch_one<VARIANT, DEPTH, DEEP, 8>(a);
	return launch_one<VARIANT, DEPTH, DEEP, 0>(a);
}

template <int VARIANT, bool DEEP> cudaError_t launch_depth(const FastLaunch &a) {
	const int d = a.t->depth;
	if (d <= 3) return launch_xs<VARIANT, 3, DEEP>(a);
	if (d <= 4) return launch_xs<VARIANT, 4, DEEP>(a);
	if (d <= 5) return launch_xs<VARIANT, 5, DEEP>(a);
	if (d <= 6) return launch_xs<VARIANT, 6, DEEP>(a);
	if (d <= 8) return launch_xs<VARIANT, 8, DEEP>(a);
	if (d <= 12) return launch_xs<VARIANT, 12, DEEP>(a);
	return cudaErrorNotSupported;
}

}  // namespace fast

using fast::FastLaunch;

// One definition per instantiation unit.
cudaError_t launch_fast_down_u8(const FastLaunch &a);
cudaError_t launch_fast_down_u16(const FastLaunch &a);
cudaError_t launch_fast_up_u8(const FastLaunch &a);
cudaError_t launch_fast_up_u16(const FastLaunch &a);

}  // namespace picha_b200
#endif
