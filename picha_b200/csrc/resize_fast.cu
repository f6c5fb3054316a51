// Fast fused separable resize for sm_100a (the throughput path; reference: src/resize.cc:66-134).
//
// Work split: one CTA (128 threads) produces a tile of `tile_w` x `band_h` output pixels of one
// image.  The source rows the tile needs are streamed through a shared-memory ring by TMA
// (cp.async.bulk.tensor, one elected thread, mbarrier completion), so global latency is hidden by
// the copy engine, not by occupancy.
//
// Pass 1 (vertical, in registers).  A thread owns 8 consecutive channel values of the source row
// (2 words of u8 data, 4 of u16) -- columns are independent in the vertical direction, so every
// source byte is read from shared memory once, unpacked once, and used for all the output rows it
// contributes to.  The per-thread state is either
//   kDown  a ring of DEPTH accumulators (one per output row currently open), or
//   kUp    a window of DEPTH unpacked source rows,
// rotated by loop unrolling so every register index is static.  The vertical weights and the
// loop bounds are the same for every thread of the grid, so they travel as a kernel parameter
// (`VTable`, in the constant bank): the compiler keeps them in uniform registers, the FMAs take
// the weight as a uniform operand (a third vector-register operand halves the FMA issue rate on
// this part: tools/microbench/row_body.cu) and all loop control stays on the uniform datapath.
// Finished rows go to shared memory as floats, G output rows per group.
//
// Pass 2 (horizontal, from shared memory).  Each thread takes output pixels of the group; lanes of
// a quarter-warp walk different rows of the group at the same x, so the float4 reads are
// bank-conflict free and the x weights are broadcast.  Results are packed to u8/u16 into a
// shared-memory tile and leave as 16-byte coalesced row segments.
//
// Why vertical first: the reference filters horizontally first, which on a GPU forces every
// source value through shared memory as a float once per tap (about 1 byte of shared-memory
// traffic per MAC, the SM's limit).  Vertical-first keeps 80 % of the MACs of a 4x downscale on
// thread-private data.  The price is the summation order: results are within +-1 LSB of the
// reference rather than bit-exact (resize_exact.cu is the bit-exact path).
//
// No tensor cores: FP32 FMA on a byte stream, bounded by HBM on one side and FP32 issue on the
// other (DESIGN.md has the arithmetic).
#include <cuda.h>

#include <algorithm>

#include "kernels.h"
#include "pixel.cuh"
#include "tables.h"

namespace picha_b200 {

namespace {

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

constexpr int RS = PICHA_FAST_RS;  // source rows per TMA stage
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
	int ytab[kYtabMax];         // kDown: cum[out_base + i]; kUp: lo[out_base + i]
	float wt[kWtMax];           // kDown: weights of source row row_base + i / WS; kUp: of output out_base + i / WS
};
static_assert(sizeof(VTable) <= 28 * 1024, "kernel parameters are limited to 32,764 bytes");

struct SmemLayout {
	int row_bytes;    // bytes per staged source row
	int ring, tmp, out, out_stride, xw, xf, xc, bars, total;
};

__host__ __device__ inline SmemLayout smem_layout(bool deep, int tile_w, int bpp, int xstride) {
	SmemLayout L;
	L.row_bytes = NT * NV * (deep ? 2 : 1);
	L.ring = 0;
	L.tmp = L.ring + NS * RS * L.row_bytes;
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

// pack: floor(clamp(f * max + 0.5)) without F2I (which costs several issue cycles here): adding 2^23
// with round-toward-minus-infinity leaves floor(t) in the low mantissa bits; the clamp is done on
// the biased float; byte/halfword merging with PRMT drops the exponent bits.
template <bool DEEP> __device__ __forceinline__ uint32_t pack_biased(float f) {
	float t = __fadd_rd(fmaf(f, Depth<DEEP>::maxv, 0.5f), 8388608.0f);
	t = fminf(fmaxf(t, 8388608.0f), 8388608.0f + Depth<DEEP>::maxv);
	return __float_as_uint(t);
}

template <int C, bool DEEP> __device__ __forceinline__ void store_pixel(uint8_t *d, const float *acc) {
	constexpr int BPP = C * Depth<DEEP>::bytes;
	uint32_t v[C];
#pragma unroll
	for (int ch = 0; ch < C; ++ch) v[ch] = pack_biased<DEEP>(acc[ch]);
	if (BPP == 4 && !DEEP) {
		const uint32_t lo = __byte_perm(v[0], v[1 % C], 0x0040), hi = __byte_perm(v[2 % C], v[3 % C], 0x0040);
		*reinterpret_cast<uint32_t *>(d) = __byte_perm(lo, hi, 0x5410);
	} else if (BPP == 8) {
		*reinterpret_cast<uint2 *>(d) = make_uint2(__byte_perm(v[0], v[1 % C], 0x5410), __byte_perm(v[2 % C], v[3 % C], 0x5410));
	} else if (DEEP) {
#pragma unroll
		for (int ch = 0; ch < C; ++ch) reinterpret_cast<uint16_t *>(d)[ch] = (uint16_t)v[ch];
	} else {
#pragma unroll
		for (int ch = 0; ch < C; ++ch) d[ch] = (uint8_t)v[ch];
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
	int xshort;            // 4 or 8: every column has at most that many taps (unrolled path); else 0
};

// shared-memory tile -> global, 16 bytes per thread where the destination allows it
template <int BPP> __device__ __forceinline__ void copy_out(const Pass2Args &a) {
	__syncthreads();
	const int row_bytes = a.tw * BPP;
	const bool vec = ((reinterpret_cast<uintptr_t>(a.gbase) | (uintptr_t)a.dstride) & 15) == 0;
	const int nvec = vec ? row_bytes >> 4 : 0;
	for (int i = a.tid; i < a.ng * nvec; i += NT) {
		const int g = i / nvec, j = i - g * nvec;
		reinterpret_cast<uint4 *>(a.gbase + (long long)g * a.dstride)[j] = lds<uint4>(a.sbase + a.outt + g * a.out_stride + 16 * j);
	}
	const int tail0 = nvec << 4, tail = row_bytes - tail0;
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
		const uint32_t v0 = a.sbase + a.tmp + 4 * (g * TMPS + lds<int>(a.sbase + a.xf + 4 * xx) * C);
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
		uint8_t *d = smem + a.outt + g * a.out_stride + xx * BPP;
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
	for (int o = a.tid; o < a.tw * 4; o += NT) {
		const int g = o & 3, xx = o >> 2;
		if (g >= a.ng) continue;
		const uint32_t w = a.sbase + a.xw + 4 * xx * a.xstride;
		const uint32_t v0 = a.sbase + a.tmp + 4 * (g * TMPS + lds<int>(a.sbase + a.xf + 4 * xx) * C);
		float wk[XT];
#pragma unroll
		for (int q = 0; q < XT / 4; ++q) {
			const float4 wq = lds<float4>(w + 16 * q);
			wk[4 * q] = wq.x; wk[4 * q + 1] = wq.y; wk[4 * q + 2] = wq.z; wk[4 * q + 3] = wq.w;
		}
		float acc[C];
#pragma unroll
		for (int ch = 0; ch < C; ++ch) acc[ch] = 0.0f;
#pragma unroll
		for (int k = 0; k < XT; ++k) {
			if (C == 4) {
				const float4 p = lds<float4>(v0 + 16 * k);
				acc[0] = fmaf(wk[k], p.x, acc[0]); acc[1 % C] = fmaf(wk[k], p.y, acc[1 % C]);
				acc[2 % C] = fmaf(wk[k], p.z, acc[2 % C]); acc[3 % C] = fmaf(wk[k], p.w, acc[3 % C]);
			} else if (C == 2) {
				const float2 p = lds<float2>(v0 + 8 * k);
				acc[0] = fmaf(wk[k], p.x, acc[0]); acc[1 % C] = fmaf(wk[k], p.y, acc[1 % C]);
			} else {
#pragma unroll
				for (int ch = 0; ch < C; ++ch) acc[ch] = fmaf(wk[k], lds<float>(v0 + 4 * (C * k + ch)), acc[ch]);
			}
		}
		store_pixel<C, DEEP>(smem + a.outt + g * a.out_stride + xx * BPP, acc);
	}
	copy_out<BPP>(a);
}

template <int C, bool DEEP> __device__ __forceinline__ void pass2_any(const Pass2Args &a) {
	if (a.xshort == 4) pass2_short<C, DEEP, 4>(a);
	else if (a.xshort == 8) pass2_short<C, DEEP, 8>(a);
	else pass2<C, DEEP>(a);
}

template <bool DEEP> __device__ __forceinline__ void run_pass2(const Pass2Args &a, int channels) {
	__syncthreads();           // the group's intermediate rows are complete
	switch (channels) {
		case 1: pass2_any<1, DEEP>(a); break;
		case 2: pass2_any<2, DEEP>(a); break;
		case 3: pass2_any<3, DEEP>(a); break;
		default: pass2_any<4, DEEP>(a); break;
	}
	__syncthreads();           // pass 1 may overwrite the intermediate rows again
}

#ifndef PICHA_FAST_MIN_CTAS
#define PICHA_FAST_MIN_CTAS 4
#endif
template <int VARIANT, int DEPTH, bool DEEP>
__global__ void __launch_bounds__(NT, (DEPTH <= 6 ? PICHA_FAST_MIN_CTAS : 1))
resize_fast_kernel(const __grid_constant__ CUtensorMap smap, DevBatch dst, FastTables t,
                   const __grid_constant__ VTable vt, int channels) {
	constexpr int WPT = DEEP ? 4 : 2;            // 32-bit words of a source row per thread
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
	const int nstages = (rhi - rlo + RS) / RS;

	const SmemLayout L = smem_layout(DEEP, t.tile_w, bpp, t.xstride);
	uint64_t *bars = reinterpret_cast<uint64_t *>(smem + L.bars);
	uint32_t sbase = smem_u32(smem);
	asm volatile("" : "+r"(sbase));   // keep it in a register: never re-derived

	auto issue_stage = [&](int k) {
		constexpr int BOXES = DEEP ? 2 : 1;      // TMA boxes are at most 256 elements wide
		uint64_t *bar = bars + (k % NS);
		mbar_expect_tx(bar, RS * L.row_bytes);
		uint8_t *d = smem + L.ring + (k % NS) * RS * L.row_bytes;
#pragma unroll
		for (int b = 0; b < BOXES; ++b) tma_load_3d(d + b * RS * 1024, &smap, bar, word0 + b * 256, rlo + k * RS, blockIdx.z);
	};

	if (tid == 0) {
		for (int i = 0; i < NS; ++i) mbar_init(bars + i, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		for (int k = 0; k < NS - 1 && k < nstages; ++k) issue_stage(k);
	}
	// this tile's horizontal tables -> shared memory
	for (int i = tid; i < tw * t.xstride; i += NT) sts(sbase + L.xw + 4 * i, t.xw[(long long)x0 * t.xstride + i]);
	for (int i = tid; i < tw; i += NT) {
		sts(sbase + L.xf + 4 * i, t.xfirst[x0 + i] - sx0);
		sts(sbase + L.xc + 4 * i, t.xcount[x0 + i]);
	}
	if (tid < 64) sts(sbase + L.tmp + G * TMPS * 4 + 4 * tid, 0.0f);
	__syncthreads();   // tables and barrier initialisation are visible to every thread

	// ---- ring consumer -------------------------------------------------------------------------
	int stage = -1, slot = NS - 1, rows_left = 0;   // uniform
	uint32_t parity = 1;
	uint32_t doff = 0;                              // shared address of this thread's words in the next row
	const int thread_byte = 4 * ((DEEP ? ((tid * WPT) >> 8) * RS * 256 : 0) + ((tid * WPT) & 255));
	auto next_stage = [&]() {
		__syncthreads();                  // every thread has finished the previous stage
		++stage;
		if (++slot == NS) { slot = 0; parity ^= 1; }
		if (tid == 0 && stage + NS - 1 < nstages) issue_stage(stage + NS - 1);
		mbar_wait(bars + slot, parity);
		rows_left = RS;
		doff = sbase + L.ring + slot * RS * L.row_bytes + thread_byte;
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
	pa.xshort = t.xshort;
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
		run_pass2<DEEP>(pa, channels);
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
					for (; n >= 2; n -= 2) { row(s, wa, wb); row(s, wb, wa); }
					if (n > 0) {
						row(s, wa, wb);
#pragma unroll
						for (int i = 0; i < WPT; ++i) wa[i] = wb[i];
					}
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

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
	static EncodeTiledFn fn = nullptr;
	static bool tried = false;
	if (!tried) {
		tried = true;
		void *p = nullptr;
		cudaDriverEntryPointQueryResult q;
		if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
		    q == cudaDriverEntryPointSuccess)
			fn = reinterpret_cast<EncodeTiledFn>(p);
		else
			cudaGetLastError();
	}
	return fn;
}

struct LaunchArgs {
	const CUtensorMap *map;
	const DevBatch *dst;
	const FastTables *t;
	const VTable *vt;
	int n, channels, smem_bytes, bands;
	cudaStream_t stream;
};

template <int VARIANT, int DEPTH, bool DEEP> cudaError_t launch_one(const LaunchArgs &a) {
	auto kern = resize_fast_kernel<VARIANT, DEPTH, DEEP>;
	cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, a.smem_bytes);
	if (e != cudaSuccess) return e;
	dim3 grid((a.dst->width + a.t->tile_w - 1) / a.t->tile_w, a.bands, a.n);
	kern<<<grid, NT, a.smem_bytes, a.stream>>>(*a.map, *a.dst, *a.t, *a.vt, a.channels);
	return cudaGetLastError();
}

template <int VARIANT, bool DEEP> cudaError_t launch_depth(const LaunchArgs &a) {
	const int d = a.t->depth;
	if (d <= 3) return launch_one<VARIANT, 3, DEEP>(a);
	if (d <= 4) return launch_one<VARIANT, 4, DEEP>(a);
	if (d <= 5) return launch_one<VARIANT, 5, DEEP>(a);
	if (d <= 6) return launch_one<VARIANT, 6, DEEP>(a);
	if (d <= 8) return launch_one<VARIANT, 8, DEEP>(a);
	if (d <= 12) return launch_one<VARIANT, 12, DEEP>(a);
	return cudaErrorNotSupported;
}

int padded_depth(int d) {   // DEPTH the kernel is instantiated with
	const int steps[] = {3, 4, 5, 6, 8, 12};
	for (int s : steps)
		if (d <= s) return s;
	return 0;
}

}  // namespace

int fast_tile_width(const int *xfirst, const int *xcount, int dst_w, int channels, int unit, int cap) {
	const int limit = NT * NV / channels;   // source pixels one CTA row holds
	for (int tw = cap / unit * unit; tw >= unit; tw -= unit) {
		bool ok = true;
		for (int x0 = 0; x0 < dst_w && ok; x0 += tw) {
			const int x1 = (x0 + tw < dst_w ? x0 + tw : dst_w) - 1;
			int hi = 0;   // the right edge is not monotone in x in general (trimmed zero taps): scan the tile
			for (int x = x0; x <= x1; ++x)
				if (xfirst[x] + xcount[x] > hi) hi = xfirst[x] + xcount[x];
			int lo = xfirst[x0];
			for (int x = x0; x <= x1; ++x)
				if (xfirst[x] < lo) lo = xfirst[x];
			if (lo != xfirst[x0] || hi - xfirst[x0] / unit * unit > limit) ok = false;
		}
		if (ok) return tw;
	}
	return 0;
}

cudaError_t launch_resize_fast(const DevBatch &src, const DevBatch &dst, int n, const FastTables &tables,
                               const FastAxisY &fy, cudaStream_t stream, int *launches) {
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
	const SmemLayout L = smem_layout(deep, t.tile_w, bpp, t.xstride);
	if (L.total > max_dynamic_smem()) return cudaErrorNotSupported;

	// The batch as a 3-D tensor of 32-bit words: (words per row, rows, images).
	CUtensorMap map;
	const cuuint64_t row_words = ((cuuint64_t)src.width * bpp + 3) / 4;
	if (row_words * 4 > (cuuint64_t)src.stride) return cudaErrorNotSupported;
	cuuint64_t dims[3] = {row_words, (cuuint64_t)src.height, (cuuint64_t)n};
	cuuint64_t strides[2] = {(cuuint64_t)src.stride, (cuuint64_t)(n > 1 ? src.step : (int64_t)src.stride * src.height)};
	if (strides[1] & 15) strides[1] = (strides[1] + 15) & ~15ull;   // n == 1: never dereferenced
	cuuint32_t box[3] = {256, RS, 1};
	cuuint32_t estr[3] = {1, 1, 1};
	CUresult r = encode(&map, CU_TENSOR_MAP_DATA_TYPE_UINT32, 3, src.base, dims, strides, box, estr,
	                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
	                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
	if (r != CUDA_SUCCESS) return cudaErrorNotSupported;

	// Bands: enough CTAs for ~16 waves of 4 CTAs/SM when the batch is small, tall strips (little
	// vertical halo) when it is large; multiples of 8 rows; small enough that one band's vertical
	// tables fit a launch's parameter block.
	const int WS = (depth + 3) & ~3;
	const int dh = dst.height;
	const long long tiles = (long long)((dst.width + t.tile_w - 1) / t.tile_w) * n;
	long long want = (148LL * 4 * 16 + tiles - 1) / tiles;
	const int max_bands = dh / 16 > 0 ? dh / 16 : 1;
	if (want > max_bands) want = max_bands;
	if (want < 1) want = 1;
	int band_h = (int)(((dh + want - 1) / want + 7) / 8 * 8);
	auto band_fits = [&](int y0, int y1) {
		const int rows = fy.cum[y1 - 1] - fy.smin[y0] + 1;
		const int outs = fy.variant == 0 ? y1 - fy.ybase[fy.smin[y0]] : y1 - y0;
		return outs <= kYtabMax && (fy.variant == 0 ? rows : outs) * WS <= kWtMax;
	};
	for (;;) {
		bool ok = true;
		for (int y0 = 0; y0 < dh && ok; y0 += band_h) ok = band_fits(y0, std::min(dh, y0 + band_h));
		if (ok) break;
		if (band_h <= 8) return cudaErrorNotSupported;
		band_h = (band_h / 2 + 7) / 8 * 8;
	}
	t.band_h = band_h;
	t.depth = depth;

	LaunchArgs a;
	a.map = &map; a.dst = &dst; a.t = &t; a.n = n; a.channels = channels; a.smem_bytes = L.total; a.stream = stream;
	VTable vt;
	a.vt = &vt;
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
		const int row_lo = fy.smin[yb], row_hi = fy.cum[ye - 1];
		vt.row_base = row_lo;
		vt.out_base = fy.variant == 0 ? fy.ybase[row_lo] : yb;
		for (int b = 0; b < bands; ++b) {
			const int y0 = yb + b * band_h, y1 = std::min(ye, y0 + band_h);
			vt.band_rlo[b] = fy.smin[y0];
			vt.band_rhi[b] = fy.cum[y1 - 1];
			vt.band_ys[b] = fy.variant == 0 ? fy.ybase[fy.smin[y0]] : y0;
		}
		const int *ysrc = fy.variant == 0 ? fy.cum.data() : fy.lo.data();
		for (int y = vt.out_base; y < ye; ++y) vt.ytab[y - vt.out_base] = ysrc[y];
		// weight rows are re-strided from the host table's stride to the kernel's WS
		const int first = fy.variant == 0 ? row_lo : yb, last = fy.variant == 0 ? row_hi : ye - 1;
		for (int i = first; i <= last; ++i)
			for (int j = 0; j < WS; ++j)
				vt.wt[(i - first) * WS + j] = j < fy.stride ? fy.wv[(size_t)i * fy.stride + j] : 0.0f;
		a.bands = bands;
		cudaError_t e;
		if (fy.variant == 0) e = deep ? launch_depth<0, true>(a) : launch_depth<0, false>(a);
		else e = deep ? launch_depth<1, true>(a) : launch_depth<1, false>(a);
		if (e != cudaSuccess) return e;
		*launches += 1;
		yb = ye;
	}
	return cudaSuccess;
}

}  // namespace picha_b200
