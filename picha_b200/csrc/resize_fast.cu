// Fast fused separable resize for sm_100a (the bandwidth path; reference: src/resize.cc:66-134).
//
// Work split: one CTA (128 threads) produces a tile of `tile_w` x `band_h` output pixels of one
// image.  The source rows the tile needs are streamed through a 3-stage shared-memory ring by TMA
// (cp.async.bulk.tensor, one elected thread, mbarrier completion), 8 rows per stage, so global
// latency is hidden by the copy engine, not by occupancy.
//
// Pass 1 (vertical, in registers).  A thread owns 8 consecutive channel values of the source row
// (2 words of u8 data, 4 of u16) -- columns are independent in the vertical direction, so every
// source byte is read from shared memory once, unpacked once, and used for all the output rows it
// contributes to with warp-uniform weights.  The per-thread state is either
//   kDown  a ring of DEPTH accumulators (one per output row currently open), or
//   kUp    a window of DEPTH unpacked source rows,
// rotated by loop unrolling so every register index is static.  Finished rows go to shared
// memory as floats, 8 output rows per group.
//
// Pass 2 (horizontal, from shared memory).  Each thread takes output pixels of the group; lanes of
// a quarter-warp walk 8 different rows of the group at the same x, so the float4 reads are
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

#include "kernels.h"
#include "pixel.cuh"

namespace picha_b200 {

namespace {

constexpr int NT = kFastThreads;
constexpr int NV = kFastValuesPerThread;
constexpr int RS = 8;              // source rows per TMA stage
constexpr int NS = 3;              // stages in the ring
constexpr int G = 8;               // output rows per pass-2 group
constexpr int TMPS = NT * NV + 4;  // floats per intermediate row (+4: rows land 4 banks apart)

struct SmemLayout {
	int row_bytes;   // bytes per staged source row
	int ring, tmp, out, out_stride, xw, xf, xc, bars, total;
};

__host__ __device__ inline SmemLayout smem_layout(bool deep, int tile_w, int bpp, int xstride) {
	SmemLayout L;
	L.row_bytes = NT * NV * (deep ? 2 : 1);
	L.ring = 0;
	L.tmp = L.ring + NS * RS * L.row_bytes;
	L.out = L.tmp + G * TMPS * 4;
	L.out_stride = ((tile_w * bpp + 127) / 128) * 128 + 16;
	L.xw = L.out + G * L.out_stride;
	L.xf = L.xw + tile_w * xstride * 4;
	L.xc = L.xf + tile_w * 4;
	L.bars = ((L.xc + tile_w * 4 + 7) / 8) * 8;
	L.total = L.bars + NS * 8;
	return L;
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
	do {
		asm volatile(
			"{\n\t.reg .pred p;\n\t"
			"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
			"selp.u32 %0, 1, 0, p;\n\t}"
			: "=r"(ok) : "r"(smem_u32(bar)), "r"(parity) : "memory");
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
template <bool DEEP> __device__ __forceinline__ void unpack8(const uint32_t *w, float *u) {
	constexpr float inv = Depth<DEEP>::inv;
	constexpr float bias = -8388608.0f * inv;
	if (DEEP) {
#pragma unroll
		for (int i = 0; i < 4; ++i) {
			u[2 * i] = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7410)), inv, bias);
			u[2 * i + 1] = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7432)), inv, bias);
		}
	} else {
#pragma unroll
		for (int i = 0; i < 2; ++i) {
			u[4 * i + 0] = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7440)), inv, bias);
			u[4 * i + 1] = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7441)), inv, bias);
			u[4 * i + 2] = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7442)), inv, bias);
			u[4 * i + 3] = fmaf(__uint_as_float(__byte_perm(w[i], 0x4B000000u, 0x7443)), inv, bias);
		}
	}
}

template <bool DEEP> __device__ __forceinline__ unsigned pack_fast(float f) {
	float t = fmaf(f, Depth<DEEP>::maxv, 0.5f);
	t = fminf(fmaxf(t, 0.0f), Depth<DEEP>::maxv);
	return (unsigned)t;
}

// Everything a CTA keeps about its tile.
struct Tile {
	const CUtensorMap *map;
	uint8_t *smem;
	SmemLayout L;
	uint64_t *bars;
	int tid;
	int word0, row0, img;     // TMA coordinates of the tile's first staged row
	int nstages;              // stages the tile consumes
	int stage, stage_row;     // consumer position
	// pass 2
	const FastTables *t;
	DevBatch dst;
	uint8_t *dimg;
	int x0, tw, sx0;
};

template <bool DEEP> __device__ __forceinline__ void issue_stage(const Tile &c, int k) {
	constexpr int BOXES = DEEP ? 2 : 1;   // TMA boxes are at most 256 elements wide
	uint64_t *bar = c.bars + (k % NS);
	mbar_expect_tx(bar, RS * c.L.row_bytes);
	uint8_t *dst = c.smem + c.L.ring + (k % NS) * RS * c.L.row_bytes;
#pragma unroll
	for (int b = 0; b < BOXES; ++b)
		tma_load_3d(dst + b * RS * 1024, c.map, bar, c.word0 + b * 256, c.row0 + k * RS, c.img);
}

// Next source row of the tile for this thread: its words in the ring (waits for the stage).
template <bool DEEP> __device__ __forceinline__ const uint32_t *next_row(Tile &c) {
	constexpr int WPT = DEEP ? 4 : 2;
	if (c.stage_row == RS) {          // uniform: every thread has finished the previous stage
		__syncthreads();
		++c.stage;
		if (c.tid == 0 && c.stage + NS - 1 < c.nstages) issue_stage<DEEP>(c, c.stage + NS - 1);
		mbar_wait(c.bars + (c.stage % NS), (c.stage / NS) & 1);
		c.stage_row = 0;
	}
	const int word = c.tid * WPT;     // a DEEP row is two 256-word boxes, each [RS][256]
	const uint8_t *base = c.smem + c.L.ring + (c.stage % NS) * RS * c.L.row_bytes;
	const uint32_t *p = reinterpret_cast<const uint32_t *>(base) + (DEEP ? (word >> 8) * RS * 256 : 0) +
	                    c.stage_row * 256 + (word & 255);
	++c.stage_row;
	return p;
}

template <bool DEEP> __device__ __forceinline__ void load_row(Tile &c, float *u) {
	const uint32_t *p = next_row<DEEP>(c);
	uint32_t w[4];
	if (DEEP) {
		uint4 v = *reinterpret_cast<const uint4 *>(p);
		w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
	} else {
		uint2 v = *reinterpret_cast<const uint2 *>(p);
		w[0] = v.x; w[1] = v.y;
	}
	unpack8<DEEP>(w, u);
}

__device__ __forceinline__ void emit_row(const Tile &c, int g, const float *v) {
	float4 *d = reinterpret_cast<float4 *>(c.smem + c.L.tmp) + (g * TMPS + c.tid * NV) / 4;
	d[0] = make_float4(v[0], v[1], v[2], v[3]);
	d[1] = make_float4(v[4], v[5], v[6], v[7]);
}

// ---- pass 2: horizontal filter of one group of intermediate rows, pack, store --------------------
template <int C, bool DEEP>
__device__ __noinline__ void pass2(const Tile &c, int gy0, int ng) {
	constexpr int BPP = C * Depth<DEEP>::bytes;
	const float *tmp = reinterpret_cast<const float *>(c.smem + c.L.tmp);
	const float *sxw = reinterpret_cast<const float *>(c.smem + c.L.xw);
	const int *sxf = reinterpret_cast<const int *>(c.smem + c.L.xf);
	const int *sxc = reinterpret_cast<const int *>(c.smem + c.L.xc);
	uint8_t *outt = c.smem + c.L.out;
	const int xstride = c.t->xstride;

	for (int o = c.tid; o < c.tw * G; o += NT) {
		const int g = o & (G - 1), xx = o >> 3;
		if (g >= ng) continue;
		const int cnt = sxc[xx];
		const float *w = sxw + xx * xstride;
		const float *v = tmp + g * TMPS + sxf[xx] * C;
		float acc[C];
#pragma unroll
		for (int ch = 0; ch < C; ++ch) acc[ch] = 0.0f;
#pragma unroll 4
		for (int k = 0; k < cnt; ++k) {
			const float wk = w[k];
			if (C == 4) {
				float4 p = *reinterpret_cast<const float4 *>(v + 4 * k);
				acc[0] = fmaf(wk, p.x, acc[0]); acc[1] = fmaf(wk, p.y, acc[1]);
				acc[2 % C] = fmaf(wk, p.z, acc[2 % C]); acc[3 % C] = fmaf(wk, p.w, acc[3 % C]);
			} else if (C == 2) {
				float2 p = *reinterpret_cast<const float2 *>(v + 2 * k);
				acc[0] = fmaf(wk, p.x, acc[0]); acc[1 % C] = fmaf(wk, p.y, acc[1 % C]);
			} else {
#pragma unroll
				for (int ch = 0; ch < C; ++ch) acc[ch] = fmaf(wk, v[C * k + ch], acc[ch]);
			}
		}
		uint8_t *d = outt + g * c.L.out_stride + xx * BPP;
		if (BPP == 4 && !DEEP) {
			*reinterpret_cast<uint32_t *>(d) = pack_fast<false>(acc[0]) | (pack_fast<false>(acc[1 % C]) << 8) |
			                                   (pack_fast<false>(acc[2 % C]) << 16) | (pack_fast<false>(acc[3 % C]) << 24);
		} else if (BPP == 8) {
			*reinterpret_cast<uint2 *>(d) = make_uint2(pack_fast<true>(acc[0]) | (pack_fast<true>(acc[1 % C]) << 16),
			                                           pack_fast<true>(acc[2 % C]) | (pack_fast<true>(acc[3 % C]) << 16));
		} else if (DEEP) {
#pragma unroll
			for (int ch = 0; ch < C; ++ch) reinterpret_cast<uint16_t *>(d)[ch] = (uint16_t)pack_fast<true>(acc[ch]);
		} else {
#pragma unroll
			for (int ch = 0; ch < C; ++ch) d[ch] = (uint8_t)pack_fast<false>(acc[ch]);
		}
	}
	__syncthreads();

	// shared-memory tile -> global, 16 bytes per thread where the destination allows it
	const int row_bytes = c.tw * BPP;
	uint8_t *gbase = c.dimg + (long long)gy0 * c.dst.stride + (long long)c.x0 * BPP;
	const bool vec = ((reinterpret_cast<uintptr_t>(gbase) | (uintptr_t)c.dst.stride) & 15) == 0;
	const int nvec = vec ? row_bytes >> 4 : 0;
	for (int i = c.tid; i < ng * nvec; i += NT) {
		const int g = i / nvec, j = i - g * nvec;
		reinterpret_cast<uint4 *>(gbase + (long long)g * c.dst.stride)[j] =
			reinterpret_cast<const uint4 *>(outt + g * c.L.out_stride)[j];
	}
	const int tail0 = nvec << 4, tail = row_bytes - tail0;
	for (int i = c.tid; i < ng * tail; i += NT) {
		const int g = i / tail, j = tail0 + (i - g * tail);
		gbase[(long long)g * c.dst.stride + j] = outt[g * c.L.out_stride + j];
	}
}

template <bool DEEP> __device__ __forceinline__ void run_pass2(const Tile &c, int channels, int gy0, int ng) {
	__syncthreads();           // the group's intermediate rows are complete
	switch (channels) {
		case 1: pass2<1, DEEP>(c, gy0, ng); break;
		case 2: pass2<2, DEEP>(c, gy0, ng); break;
		case 3: pass2<3, DEEP>(c, gy0, ng); break;
		default: pass2<4, DEEP>(c, gy0, ng); break;
	}
	__syncthreads();           // pass 1 may overwrite the intermediate rows again
}

template <int VARIANT, int DEPTH, bool DEEP>
__global__ void __launch_bounds__(NT)
resize_fast_kernel(const __grid_constant__ CUtensorMap smap, DevBatch dst, FastTables t, int channels) {
	extern __shared__ __align__(128) uint8_t smem[];
	const int bpp = channels * Depth<DEEP>::bytes;

	Tile c;
	c.map = &smap;
	c.smem = smem;
	c.L = smem_layout(DEEP, t.tile_w, bpp, t.xstride);
	c.bars = reinterpret_cast<uint64_t *>(smem + c.L.bars);
	c.tid = threadIdx.x;
	c.t = &t;
	c.dst = dst;
	c.dimg = dst.base + (long long)blockIdx.z * dst.step;
	c.img = blockIdx.z;
	c.x0 = blockIdx.x * t.tile_w;
	c.tw = min(t.tile_w, dst.width - c.x0);
	c.sx0 = t.xfirst[c.x0] / t.align_px * t.align_px;   // tile origin: 16-byte aligned in the row (TMA box start)
	c.word0 = c.sx0 * bpp / 4;

	const int y0 = blockIdx.y * t.band_h, y1 = min(dst.height, y0 + t.band_h);
	const int rlo = t.smin[y0], rhi = t.cum[y1 - 1];
	c.row0 = rlo;
	c.nstages = (rhi - rlo + RS) / RS;
	c.stage = -1;
	c.stage_row = RS;

	if (c.tid == 0) {
		for (int i = 0; i < NS; ++i) mbar_init(c.bars + i, 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		for (int k = 0; k < NS - 1 && k < c.nstages; ++k) issue_stage<DEEP>(c, k);
	}
	// this tile's horizontal tables -> shared memory
	{
		float *sxw = reinterpret_cast<float *>(smem + c.L.xw);
		int *sxf = reinterpret_cast<int *>(smem + c.L.xf), *sxc = reinterpret_cast<int *>(smem + c.L.xc);
		for (int i = c.tid; i < c.tw * t.xstride; i += NT) sxw[i] = t.xw[(long long)c.x0 * t.xstride + i];
		for (int i = c.tid; i < c.tw; i += NT) {
			sxf[i] = t.xfirst[c.x0 + i] - c.sx0;
			sxc[i] = t.xcount[c.x0 + i];
		}
	}
	// (the first next_row() starts with a __syncthreads, which also publishes the barriers and tables)

	int gcount = 0;
	if (VARIANT == 0) {
		float acc[DEPTH][NV];
#pragma unroll
		for (int j = 0; j < DEPTH; ++j)
#pragma unroll
			for (int i = 0; i < NV; ++i) acc[j][i] = 0.0f;
		int r = rlo;
		int y = t.ybase[rlo];                           // <= y0: outputs before y0 are not emitted
		while (y < y1) {
#pragma unroll
			for (int s = 0; s < DEPTH; ++s) {
				if (y < y1) {
					const int need = t.cum[y];
					for (; r <= need; ++r) {
						float u[NV];
						load_row<DEEP>(c, u);
						const float4 *wp = reinterpret_cast<const float4 *>(t.wv + (long long)r * t.ystride);
						float w[(DEPTH + 3) & ~3];
#pragma unroll
						for (int q = 0; q < (DEPTH + 3) / 4; ++q) {
							float4 v = __ldg(wp + q);
							w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
						}
#pragma unroll
						for (int j = 0; j < DEPTH; ++j)
#pragma unroll
							for (int i = 0; i < NV; ++i)
								acc[(s + j) % DEPTH][i] = fmaf(w[j], u[i], acc[(s + j) % DEPTH][i]);
					}
					if (y >= y0) {
						emit_row(c, gcount, acc[s]);
						++gcount;
					}
#pragma unroll
					for (int i = 0; i < NV; ++i) acc[s][i] = 0.0f;
					++y;
					if (gcount == G || (y == y1 && gcount > 0)) {
						run_pass2<DEEP>(c, channels, y - gcount, gcount);
						gcount = 0;
					}
				}
			}
		}
	} else {
		float win[DEPTH][NV];
		int rb = t.lo[y0], rnext = rb;
#pragma unroll
		for (int k = 0; k < DEPTH; ++k) {
			if (rnext <= rhi) { load_row<DEEP>(c, win[k]); ++rnext; }
			else {
#pragma unroll
				for (int i = 0; i < NV; ++i) win[k][i] = 0.0f;
			}
		}
		int y = y0;
		while (y < y1) {
#pragma unroll
			for (int s = 0; s < DEPTH; ++s) {
				if (y < y1) {
					while (y < y1 && t.lo[y] == rb) {
						const float4 *wp = reinterpret_cast<const float4 *>(t.wv + (long long)y * t.ystride);
						float w[(DEPTH + 3) & ~3];
#pragma unroll
						for (int q = 0; q < (DEPTH + 3) / 4; ++q) {
							float4 v = __ldg(wp + q);
							w[4 * q] = v.x; w[4 * q + 1] = v.y; w[4 * q + 2] = v.z; w[4 * q + 3] = v.w;
						}
						float o[NV];
#pragma unroll
						for (int i = 0; i < NV; ++i) o[i] = 0.0f;
#pragma unroll
						for (int k = 0; k < DEPTH; ++k)
#pragma unroll
							for (int i = 0; i < NV; ++i) o[i] = fmaf(w[k], win[(s + k) % DEPTH][i], o[i]);
						emit_row(c, gcount, o);
						++gcount;
						++y;
						if (gcount == G || y == y1) {
							run_pass2<DEEP>(c, channels, y - gcount, gcount);
							gcount = 0;
						}
					}
					if (y < y1) {                       // slide the window down one source row
						if (rnext <= rhi) { load_row<DEEP>(c, win[s]); ++rnext; }
						else {
#pragma unroll
							for (int i = 0; i < NV; ++i) win[s][i] = 0.0f;
						}
						++rb;
					}
				}
			}
		}
	}

	// Never leave with a TMA load still in flight: wait for every stage that was issued.
	for (int k = c.stage + 1; k < c.nstages && k < c.stage + NS; ++k)
		mbar_wait(c.bars + (k % NS), (k / NS) & 1);
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

template <int VARIANT, int DEPTH, bool DEEP>
cudaError_t launch_one(const CUtensorMap &map, const DevBatch &dst, int n, const FastTables &t, int channels,
                       int smem_bytes, cudaStream_t stream) {
	auto kern = resize_fast_kernel<VARIANT, DEPTH, DEEP>;
	cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes);
	if (e != cudaSuccess) return e;
	dim3 grid((dst.width + t.tile_w - 1) / t.tile_w, (dst.height + t.band_h - 1) / t.band_h, n);
	kern<<<grid, NT, smem_bytes, stream>>>(map, dst, t, channels);
	return cudaGetLastError();
}

template <int VARIANT, bool DEEP>
cudaError_t launch_depth(const CUtensorMap &map, const DevBatch &dst, int n, const FastTables &t, int channels,
                         int smem_bytes, cudaStream_t stream) {
	const int d = t.depth;
	if (d <= 3) return launch_one<VARIANT, 3, DEEP>(map, dst, n, t, channels, smem_bytes, stream);
	if (d <= 4) return launch_one<VARIANT, 4, DEEP>(map, dst, n, t, channels, smem_bytes, stream);
	if (d <= 5) return launch_one<VARIANT, 5, DEEP>(map, dst, n, t, channels, smem_bytes, stream);
	if (d <= 6) return launch_one<VARIANT, 6, DEEP>(map, dst, n, t, channels, smem_bytes, stream);
	if (d <= 8) return launch_one<VARIANT, 8, DEEP>(map, dst, n, t, channels, smem_bytes, stream);
	if (d <= 12) return launch_one<VARIANT, 12, DEEP>(map, dst, n, t, channels, smem_bytes, stream);
	return cudaErrorNotSupported;
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

cudaError_t launch_resize_fast(const DevBatch &src, const DevBatch &dst, int n, const FastTables &t,
                               cudaStream_t stream, int *launches) {
	static const int kBytes[8] = {3, 4, 1, 2, 2, 4, 6, 8}, kChannels[8] = {3, 4, 1, 2, 1, 2, 3, 4};
	const int bpp = kBytes[src.pixel], channels = kChannels[src.pixel];
	const bool deep = src.pixel >= 4;
	if (t.variant < 0 || t.tile_w <= 0 || t.depth > kFastMaxDepth) return cudaErrorNotSupported;
	if ((reinterpret_cast<uintptr_t>(src.base) & 15) || (src.stride & 15) || (n > 1 && (src.step & 15)))
		return cudaErrorNotSupported;
	if (n > 65535 || (dst.height + t.band_h - 1) / t.band_h > 65535) return cudaErrorNotSupported;
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

	*launches += 1;
	if (t.variant == 0)
		return deep ? launch_depth<0, true>(map, dst, n, t, channels, L.total, stream)
		            : launch_depth<0, false>(map, dst, n, t, channels, L.total, stream);
	return deep ? launch_depth<1, true>(map, dst, n, t, channels, L.total, stream)
	            : launch_depth<1, false>(map, dst, n, t, channels, L.total, stream);
}

}  // namespace picha_b200
