// Pixel-format conversion kernels (reference: ColorConverter<Src,Dst>::op, src/colorconvert.cc:136-152,
// with the thirteen ChannelConvertOp specialisations of :24-134).
//
// The reference round-trips every channel through float (unpack -> op -> pack).  For everything
// except luma that round trip is an integer identity, verified for every u8/u16 value against the
// reference's own output (tests/test_oracle.py::test_depth_identities):
//     same depth: v            u8 -> u16: v * 257            u16 -> u8: (v * 255 + 32767) / 65535
//     constants:  1.0f -> max  0.0f -> 0
// Luma ((s0*r + s1*g) + s2*b, :90) is float and must not be contracted into FMAs, or 12 pixels
// of a 1080p frame come out one LSB off; pixel.cuh's *_rn intrinsics guarantee that.
//
// Two kernels:
//   convert_rows_kernel  -- the bandwidth path: a warp converts 128 pixels per step; global
//       traffic is 32-bit-per-lane fully coalesced (every 128-byte line is touched by exactly
//       one request) and the word -> pixel regrouping goes through a warp-private shared-memory
//       tile, so 3- and 6-byte pixels cost the same as the power-of-two ones.
//       Needs 4-byte aligned row starts on both sides.
//   convert_pixels_kernel -- any alignment / stride / width (subView inputs, row tails).
#include "kernels.h"
#include "pixel.cuh"
#include "pixel_convert.cuh"

namespace picha_b200 {

namespace {

constexpr int kWarps = 8;
constexpr int kGroup = 128;   // pixels per warp step: 4 per lane

// Channel i of a little-endian word array, zero-extended: one byte permute.
template <bool DEEP> __device__ __forceinline__ unsigned word_get(const unsigned *w, int i) {
	if (DEEP) return __byte_perm(w[i >> 1], 0u, (i & 1) ? 0x4432 : 0x4410);
	return __byte_perm(w[i >> 2], 0u, 0x4440 + (i & 3));
}
// The same with 0x4B in the top byte: the bit pattern of the float 2^23 + value.
template <bool DEEP> __device__ __forceinline__ unsigned word_get_magic(const unsigned *w, int i) {
	if (DEEP) return __byte_perm(w[i >> 1], 0x4B000000u, (i & 1) ? 0x7432 : 0x7410);
	return __byte_perm(w[i >> 2], 0x4B000000u, 0x7440 + (i & 3));
}
// Insert the low byte (halfword) of v as channel i; upper bits of v are ignored.
template <bool DEEP> __device__ __forceinline__ void word_put(unsigned *w, int i, unsigned v) {
	if (DEEP) w[i >> 1] = __byte_perm(w[i >> 1], v, (i & 1) ? 0x5410 : 0x3254);
	else w[i >> 2] = __byte_perm(w[i >> 2], v, (i & 3) == 0 ? 0x3214 : (i & 3) == 1 ? 0x3240 : (i & 3) == 2 ? 0x3410 : 0x4210);
}

// A lane's N consecutive words of a warp tile, with the widest shared-memory access N allows
// (scalar accesses at an even word stride would be 2- to 8-way bank conflicted).
template <int N> __device__ __forceinline__ void lane_words_load(unsigned *r, const unsigned *tile, int lane) {
	if constexpr (N % 4 == 0) {
#pragma unroll
		for (int j = 0; j < N / 4; ++j) {
			uint4 v = reinterpret_cast<const uint4 *>(tile)[lane * (N / 4) + j];
			r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
		}
	} else if constexpr (N % 2 == 0) {
#pragma unroll
		for (int j = 0; j < N / 2; ++j) {
			uint2 v = reinterpret_cast<const uint2 *>(tile)[lane * (N / 2) + j];
			r[2 * j] = v.x; r[2 * j + 1] = v.y;
		}
	} else {
#pragma unroll
		for (int j = 0; j < N; ++j) r[j] = tile[lane * N + j];
	}
}
template <int N> __device__ __forceinline__ void lane_words_store(unsigned *tile, const unsigned *r, int lane) {
	if constexpr (N % 4 == 0) {
#pragma unroll
		for (int j = 0; j < N / 4; ++j)
			reinterpret_cast<uint4 *>(tile)[lane * (N / 4) + j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
	} else if constexpr (N % 2 == 0) {
#pragma unroll
		for (int j = 0; j < N / 2; ++j)
			reinterpret_cast<uint2 *>(tile)[lane * (N / 2) + j] = make_uint2(r[2 * j], r[2 * j + 1]);
	} else {
#pragma unroll
		for (int j = 0; j < N; ++j) tile[lane * N + j] = r[j];
	}
}

// The same straight from / to global memory, for formats whose pixels are 1, 2, 4 or 8 bytes: a lane's four pixels
// are one aligned 4-, 8- or 16-byte unit (two for 8-byte pixels), consecutive lanes read consecutive units -- already
// the fully coalesced pattern, so the shared-memory regrouping (and its 2 barriers and ~10 instructions per step of
// an issue-bound kernel) is only needed on the sides with 3- and 6-byte pixels.
__host__ __device__ constexpr bool direct_words(int n) { return n == 1 || n == 2 || n == 4 || n == 8; }
template <int N> __device__ __forceinline__ void lane_words_ldg(unsigned *r, const unsigned *row, int lane) {
	if constexpr (N % 4 == 0) {
#pragma unroll
		for (int j = 0; j < N / 4; ++j) {
			const uint4 v = __ldg(reinterpret_cast<const uint4 *>(row) + lane * (N / 4) + j);
			r[4 * j] = v.x; r[4 * j + 1] = v.y; r[4 * j + 2] = v.z; r[4 * j + 3] = v.w;
		}
	} else if constexpr (N == 2) {
		const uint2 v = __ldg(reinterpret_cast<const uint2 *>(row) + lane);
		r[0] = v.x; r[1] = v.y;
	} else {
		r[0] = __ldg(row + lane);
	}
}
template <int N> __device__ __forceinline__ void lane_words_stg(unsigned *row, const unsigned *r, int lane) {
	if constexpr (N % 4 == 0) {
#pragma unroll
		for (int j = 0; j < N / 4; ++j)
			reinterpret_cast<uint4 *>(row)[lane * (N / 4) + j] = make_uint4(r[4 * j], r[4 * j + 1], r[4 * j + 2], r[4 * j + 3]);
	} else if constexpr (N == 2) {
		reinterpret_cast<uint2 *>(row)[lane] = make_uint2(r[0], r[1]);
	} else {
		row[lane] = r[0];
	}
}

template <int SC, bool SDEEP, int DC, bool DDEEP, bool CMYK = false>
__device__ __forceinline__ void convert_one_unaligned(const uint8_t *s, uint8_t *d, float rf, float gf, float bf) {
	unsigned in[4], out[4];
#pragma unroll
	for (int c = 0; c < SC; ++c) in[c] = load_channel<SDEEP>(s + c * Depth<SDEEP>::bytes);
	convert_pixel<SC, SDEEP, DC, DDEEP, false, CMYK>(in, out, rf, gf, bf);
#pragma unroll
	for (int c = 0; c < DC; ++c) store_channel<DDEEP>(d + c * Depth<DDEEP>::bytes, out[c]);
}

// DIRECT: both images are 16-byte aligned; the sides whose pixels are 1, 2, 4 or 8 bytes skip the shared-memory tile
template <int SC, bool SDEEP, int DC, bool DDEEP, bool CMYK = false, bool DIRECT = false>
__global__ void __launch_bounds__(kWarps * 32)
convert_rows_kernel(DevBatch src, DevBatch dst, int groups_per_row, float rf, float gf, float bf) {
	constexpr int SW = SC * Depth<SDEEP>::bytes;   // source words per lane per step (= bytes per pixel)
	constexpr int DW = DC * Depth<DDEEP>::bytes;
	constexpr bool SDIR = DIRECT && direct_words(SW), DDIR = DIRECT && direct_words(DW);
	__shared__ __align__(16) unsigned tile_in[kWarps][SDIR ? 4 : SW * 32];
	__shared__ __align__(16) unsigned tile_out[kWarps][DDIR ? 4 : DW * 32];

	const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
	unsigned *tin = tile_in[warp], *tout = tile_out[warp];
	// A warp owns one 128-pixel column group and walks down the rows (no per-step index division;
	// only warp-level synchronisation is used, so a warp without a group simply leaves).
	const int gx = blockIdx.x * kWarps + warp;
	if (gx >= groups_per_row) return;
	const int px0 = gx * kGroup;
	const int npx = min(kGroup, src.width - px0);
	const uint8_t *simg = src.base + (long long)blockIdx.z * src.step + (long long)px0 * SW;
	uint8_t *dimg = dst.base + (long long)blockIdx.z * dst.step + (long long)px0 * DW;

	if (npx == kGroup) {
		// full group: 32-bit coalesced loads, one row ahead of the row being converted
		unsigned win[SW];
		int y = blockIdx.y;
		auto fetch = [&](unsigned (&w)[SW], int row) {
			const unsigned *s32 = reinterpret_cast<const unsigned *>(simg + (long long)row * src.stride);
			if constexpr (SDIR) {
				lane_words_ldg<SW>(w, s32, lane);
			} else {
#pragma unroll
				for (int j = 0; j < SW; ++j) w[j] = __ldg(s32 + j * 32 + lane);
			}
		};
		if (y < src.height) fetch(win, y);
		for (; y < src.height; y += gridDim.y) {
			unsigned nxt[SW];
			if (y + (int)gridDim.y < src.height) fetch(nxt, y + gridDim.y);
			unsigned mine[SW];
			if constexpr (SDIR) {
#pragma unroll
				for (int j = 0; j < SW; ++j) mine[j] = win[j];
			} else {
				__syncwarp();   // previous step's readers are done with the tiles
#pragma unroll
				for (int j = 0; j < SW; ++j) tin[j * 32 + lane] = win[j];
				__syncwarp();
				lane_words_load<SW>(mine, tin, lane);
			}

			unsigned packed[DW];
#pragma unroll
			for (int j = 0; j < DW; ++j) packed[j] = 0u;
#pragma unroll
			for (int p = 0; p < 4; ++p) {
				unsigned in[4], out[4];
				constexpr bool LUMA = SC >= 3 && DC <= 2 && !CMYK;
#pragma unroll
				for (int c = 0; c < SC; ++c)
					in[c] = (LUMA && c < 3) ? word_get_magic<SDEEP>(mine, p * SC + c) : word_get<SDEEP>(mine, p * SC + c);
				convert_pixel<SC, SDEEP, DC, DDEEP, LUMA, CMYK>(in, out, rf, gf, bf);
#pragma unroll
				for (int c = 0; c < DC; ++c) word_put<DDEEP>(packed, p * DC + c, out[c]);
			}
			unsigned *d32 = reinterpret_cast<unsigned *>(dimg + (long long)y * dst.stride);
			if constexpr (DDIR) {
				lane_words_stg<DW>(d32, packed, lane);
			} else {
				if constexpr (SDIR) __syncwarp();   // (the staged source path has synchronised already)
				lane_words_store<DW>(tout, packed, lane);
				__syncwarp();
#pragma unroll
				for (int j = 0; j < DW; ++j) d32[j * 32 + lane] = tout[j * 32 + lane];
			}
#pragma unroll
			for (int j = 0; j < SW; ++j) win[j] = nxt[j];
		}
	} else {
		// row tail: fewer than 128 pixels left; only payload bytes may be written
		for (int y = blockIdx.y; y < src.height; y += gridDim.y) {
			const uint8_t *srow = simg + (long long)y * src.stride;
			uint8_t *drow = dimg + (long long)y * dst.stride;
			for (int p = lane; p < npx; p += 32)
				convert_one_unaligned<SC, SDEEP, DC, DDEEP, CMYK>(srow + p * SW, drow + p * DW, rf, gf, bf);
		}
	}
}

template <int SC, bool SDEEP, int DC, bool DDEEP, bool CMYK = false>
__global__ void __launch_bounds__(256)
convert_pixels_kernel(DevBatch src, DevBatch dst, float rf, float gf, float bf) {
	constexpr int SB = SC * Depth<SDEEP>::bytes, DB = DC * Depth<DDEEP>::bytes;
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	if (x >= src.width) return;
	for (int y = blockIdx.y; y < src.height; y += gridDim.y) {
		const uint8_t *s = src.base + (long long)blockIdx.z * src.step + (long long)y * src.stride + (long long)x * SB;
		uint8_t *d = dst.base + (long long)blockIdx.z * dst.step + (long long)y * dst.stride + (long long)x * DB;
		convert_one_unaligned<SC, SDEEP, DC, DDEEP, CMYK>(s, d, rf, gf, bf);
	}
}

// Same format: NativeImage::copy, src/picha.cc:27-34 -- payload bytes of each row.
__global__ void __launch_bounds__(256)
copy_rows_kernel(DevBatch src, DevBatch dst, int row_bytes, int vec_ok) {
	const uint8_t *s = src.base + (long long)blockIdx.z * src.step + (long long)blockIdx.y * src.stride;
	uint8_t *d = dst.base + (long long)blockIdx.z * dst.step + (long long)blockIdx.y * dst.stride;
	const int tid = blockIdx.x * blockDim.x + threadIdx.x, nth = gridDim.x * blockDim.x;
	int done = 0;
	if (vec_ok) {
		const int nvec = row_bytes >> 4;
		const uint4 *s4 = reinterpret_cast<const uint4 *>(s);
		uint4 *d4 = reinterpret_cast<uint4 *>(d);
		for (int i = tid; i < nvec; i += nth) d4[i] = __ldg(s4 + i);
		done = nvec << 4;
	}
	for (int i = done + tid; i < row_bytes; i += nth) d[i] = s[i];
}

bool aligned4(const DevBatch &b) {
	return (reinterpret_cast<uintptr_t>(b.base) & 3) == 0 && (b.stride & 3) == 0 && (b.step & 3) == 0;
}
bool aligned16(const DevBatch &b) {
	return (reinterpret_cast<uintptr_t>(b.base) & 15) == 0 && (b.stride & 15) == 0 && (b.step & 15) == 0;
}

int sm_count() {
	static int n = 0;
	if (!n) {
		int dev = 0;
		cudaGetDevice(&dev);
		if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
	}
	return n;
}

template <int SC, bool SDEEP, int DC, bool DDEEP, bool CMYK = false>
cudaError_t launch_pair(const DevBatch &src, const DevBatch &dst, int n, float rf, float gf, float bf,
                        cudaStream_t stream, int *launches) {
	if (aligned4(src) && aligned4(dst)) {
		const int gpr = (src.width + kGroup - 1) / kGroup;
		const int gx = (gpr + kWarps - 1) / kWarps;
		for (int z0 = 0; z0 < n; z0 += 65535) {
			const int nz = n - z0 < 65535 ? n - z0 : 65535;
			// rows are dealt out over gridDim.y so that the grid is about 32 CTAs per SM in total
			long long gy = ((long long)sm_count() * 32 + (long long)gx * nz - 1) / ((long long)gx * nz);
			if (gy > src.height) gy = src.height;
			if (gy > 65535) gy = 65535;
			if (gy < 1) gy = 1;
			DevBatch s = src, d = dst;
			s.base += (long long)z0 * src.step;
			d.base += (long long)z0 * dst.step;
			constexpr bool can_direct = direct_words(SC * Depth<SDEEP>::bytes) || direct_words(DC * Depth<DDEEP>::bytes);
			if (can_direct && aligned16(src) && aligned16(dst))
				convert_rows_kernel<SC, SDEEP, DC, DDEEP, CMYK, can_direct><<<dim3(gx, (unsigned)gy, nz), kWarps * 32, 0, stream>>>(s, d, gpr, rf, gf, bf);
			else
				convert_rows_kernel<SC, SDEEP, DC, DDEEP, CMYK><<<dim3(gx, (unsigned)gy, nz), kWarps * 32, 0, stream>>>(s, d, gpr, rf, gf, bf);
			*launches += 1;
		}
		return cudaGetLastError();
	}
	for (int z0 = 0; z0 < n; z0 += 65535) {
		const int nz = n - z0 < 65535 ? n - z0 : 65535;
		DevBatch s = src, d = dst;
		s.base += (long long)z0 * src.step;
		d.base += (long long)z0 * dst.step;
		dim3 grid((src.width + 255) / 256, src.height < 65535 ? src.height : 65535, nz);
		convert_pixels_kernel<SC, SDEEP, DC, DDEEP, CMYK><<<grid, 256, 0, stream>>>(s, d, rf, gf, bf);
		*launches += 1;
	}
	return cudaGetLastError();
}

template <int SC, bool SDEEP>
cudaError_t launch_src(const DevBatch &src, const DevBatch &dst, int n, float rf, float gf, float bf,
                       cudaStream_t stream, int *launches) {
	switch (dst.pixel) {
		case 0: return launch_pair<SC, SDEEP, 3, false>(src, dst, n, rf, gf, bf, stream, launches);
		case 1: return launch_pair<SC, SDEEP, 4, false>(src, dst, n, rf, gf, bf, stream, launches);
		case 2: return launch_pair<SC, SDEEP, 1, false>(src, dst, n, rf, gf, bf, stream, launches);
		case 3: return launch_pair<SC, SDEEP, 2, false>(src, dst, n, rf, gf, bf, stream, launches);
		case 4: return launch_pair<SC, SDEEP, 1, true>(src, dst, n, rf, gf, bf, stream, launches);
		case 5: return launch_pair<SC, SDEEP, 2, true>(src, dst, n, rf, gf, bf, stream, launches);
		case 6: return launch_pair<SC, SDEEP, 3, true>(src, dst, n, rf, gf, bf, stream, launches);
		case 7: return launch_pair<SC, SDEEP, 4, true>(src, dst, n, rf, gf, bf, stream, launches);
	}
	return cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_color_convert(const DevBatch &src, const DevBatch &dst, int n, float rf, float gf, float bf,
                                 cudaStream_t stream, int *launches) {
	if (n <= 0 || src.width <= 0 || src.height <= 0) return cudaSuccess;
	if (src.pixel == dst.pixel) {
		static const int bytes[8] = {3, 4, 1, 2, 2, 4, 6, 8};
		const int row_bytes = src.width * bytes[src.pixel];
		const int vec_ok = aligned16(src) && aligned16(dst);
		for (int z0 = 0; z0 < n; z0 += 65535)
			for (int y0 = 0; y0 < src.height; y0 += 65535) {
				DevBatch s = src, d = dst;
				s.base += (long long)z0 * src.step + (long long)y0 * src.stride;
				d.base += (long long)z0 * dst.step + (long long)y0 * dst.stride;
				const int nz = n - z0 < 65535 ? n - z0 : 65535;
				const int ny = src.height - y0 < 65535 ? src.height - y0 : 65535;
				int gx = (row_bytes / 16 + 255) / 256;
				if (gx < 1) gx = 1;
				if (gx > 8) gx = 8;
				copy_rows_kernel<<<dim3(gx, ny, nz), 256, 0, stream>>>(s, d, row_bytes, vec_ok);
				*launches += 1;
			}
		return cudaGetLastError();
	}
	switch (src.pixel) {
		case 0: return launch_src<3, false>(src, dst, n, rf, gf, bf, stream, launches);
		case 1: return launch_src<4, false>(src, dst, n, rf, gf, bf, stream, launches);
		case 2: return launch_src<1, false>(src, dst, n, rf, gf, bf, stream, launches);
		case 3: return launch_src<2, false>(src, dst, n, rf, gf, bf, stream, launches);
		case 4: return launch_src<1, true>(src, dst, n, rf, gf, bf, stream, launches);
		case 5: return launch_src<2, true>(src, dst, n, rf, gf, bf, stream, launches);
		case 6: return launch_src<3, true>(src, dst, n, rf, gf, bf, stream, launches);
		case 7: return launch_src<4, true>(src, dst, n, rf, gf, bf, stream, launches);
	}
	return cudaErrorInvalidValue;
}

// cmyk_to_rgb of every row (src/jpegcodec.cc:36-42,96): the source is 4 bytes per pixel (C, M, Y, K as the
// decoder delivers them, carried in an rgba image), the destination rgb.
cudaError_t launch_cmyk_to_rgb(const DevBatch &src, const DevBatch &dst, int n, cudaStream_t stream, int *launches) {
	if (n <= 0 || src.width <= 0 || src.height <= 0) return cudaSuccess;
	if (src.pixel != 1 || dst.pixel != 0) return cudaErrorInvalidValue;
	return launch_pair<4, false, 3, false, true>(src, dst, n, 0.0f, 0.0f, 0.0f, stream, launches);
}

}  // namespace picha_b200
