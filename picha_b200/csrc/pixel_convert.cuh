// One pixel's format conversion on integer channel values (reference: ChannelConvertOp<S,D> and
// ColorConverter<Src,Dst>::op, src/colorconvert.cc:24-152), shared by the colour-conversion kernels and by the
// resize kernels' fused "resize, then convert" epilogue (picha_b200_resize_convert).
#ifndef PICHA_B200_PIXEL_CONVERT_CUH
#define PICHA_B200_PIXEL_CONVERT_CUH

#include "kernels.h"
#include "pixel.cuh"

namespace picha_b200 {

template <bool SDEEP, bool DDEEP> __device__ __forceinline__ unsigned depth_convert(unsigned v) {
	if (SDEEP == DDEEP) return v;
	if (DDEEP) return v * 257u;
	return (v * 255u + 32767u) / 65535u;
}

// One pixel, on integer channel values. SC/DC: channel counts. src/colorconvert.cc:24-134.
// MAGIC: the luma inputs in[0..2] arrive as 0x4B000000 | v (the float 2^23 + v) straight out of a byte
// permute, instead of as integers -- one instruction less per channel on the bandwidth path.
// CMYK (4 x u8 -> 3 x u8 only): the JPEG decoder's cmyk_to_rgb, src/jpegcodec.cc:36-42 -- integer
// c * k / 255 per channel, truncating.
template <int SC, bool SDEEP, int DC, bool DDEEP, bool MAGIC = false, bool CMYK = false>
__device__ __forceinline__ void convert_pixel(const unsigned *in, unsigned *out, float rf, float gf, float bf) {
	constexpr unsigned ONE = DDEEP ? 65535u : 255u;
	if constexpr (CMYK) {
#pragma unroll
		for (int c = 0; c < 3; ++c) out[c] = (in[c] * in[3]) / 255u;
		return;
	}
	if constexpr (SC >= 3 && DC <= 2) {          // 3->1, 3->2, 4->1, 4->2: luma (alpha ignored or passed through)
		float r = MAGIC ? unpack_magic<SDEEP>(in[0]) : unpack_value<SDEEP>(in[0]);
		float g = MAGIC ? unpack_magic<SDEEP>(in[1]) : unpack_value<SDEEP>(in[1]);
		float b = MAGIC ? unpack_magic<SDEEP>(in[2]) : unpack_value<SDEEP>(in[2]);
		float l = __fadd_rn(__fadd_rn(__fmul_rn(r, rf), __fmul_rn(g, gf)), __fmul_rn(b, bf));
		out[0] = pack_value<DDEEP>(l);
		if constexpr (DC == 2) out[1] = (SC == 4) ? depth_convert<SDEEP, DDEEP>(in[3]) : ONE;
	} else {
	unsigned v[4];
#pragma unroll
	for (int c = 0; c < SC; ++c) v[c] = depth_convert<SDEEP, DDEEP>(in[c]);
	if constexpr (SC == DC) {
#pragma unroll
		for (int c = 0; c < DC; ++c) out[c] = v[c];
	} else if constexpr (SC == 1) {              // 1->2 (g,1)  1->3 (g,g,g)  1->4 (g,g,g,1)
		out[0] = v[0];
		if constexpr (DC == 2) out[1] = ONE;
		if constexpr (DC >= 3) { out[1] = v[0]; out[2] = v[0]; }
		if constexpr (DC == 4) out[3] = ONE;
	} else if constexpr (SC == 2) {              // 2->1 g   2->3 (g,a,0)   2->4 (g,g,g,a)
		out[0] = v[0];
		if constexpr (DC == 3) { out[1] = v[1]; out[2] = 0u; }
		if constexpr (DC == 4) { out[1] = v[0]; out[2] = v[0]; out[3] = v[1]; }
	} else if constexpr (SC == 3) {              // 3->4 (r,g,b,1)
		out[0] = v[0]; out[1] = v[1]; out[2] = v[2]; out[3] = ONE;
	} else {                           // 4->3 (r,g,b)
		out[0] = v[0]; out[1] = v[1]; out[2] = v[2];
	}
	}
}


// What a resize kernel's pack stage does when the destination has another pixel format: the packed channel
// values of one resized pixel (exactly what the reference's resize would have stored) go through the
// reference's conversion and are stored channel by channel (any alignment).  Same result as resizeImage
// followed by doColorConvert, without the intermediate image.
template <int SC, bool SDEEP, int DC, bool DDEEP>
__device__ __forceinline__ void convert_store_as(uint8_t *d, const unsigned *v, const FuseArgs &f) {
	unsigned in[4], out[4];
#pragma unroll
	for (int c = 0; c < SC; ++c) in[c] = v[c] & (SDEEP ? 0xFFFFu : 0xFFu);   // (pack results carry exponent bits above the value)
	convert_pixel<SC, SDEEP, DC, DDEEP>(in, out, f.r, f.g, f.b);
#pragma unroll
	for (int c = 0; c < DC; ++c) store_channel<DDEEP>(d + c * Depth<DDEEP>::bytes, out[c]);
}

__host__ __device__ constexpr int pixel_bytes(int p) { return p == 0 ? 3 : p == 1 ? 4 : p == 2 ? 1 : p == 3 ? 2 : p == 4 ? 2 : p == 5 ? 4 : p == 6 ? 6 : 8; }

// d: address of the destination pixel (in the destination's format).  Out of line on purpose: inlined, the eight
// conversions cost the resize kernels' pack stages enough registers to spill in their unfused form.
template <int SC, bool SDEEP>
__device__ __noinline__ void convert_store_call(uint8_t *d, unsigned v0, unsigned v1, unsigned v2, unsigned v3, FuseArgs f) {
	const unsigned v[4] = {v0, v1, v2, v3};
	switch (f.dst_pixel) {   // src/picha.h:79-92
		case 0: convert_store_as<SC, SDEEP, 3, false>(d, v, f); break;
		case 1: convert_store_as<SC, SDEEP, 4, false>(d, v, f); break;
		case 2: convert_store_as<SC, SDEEP, 1, false>(d, v, f); break;
		case 3: convert_store_as<SC, SDEEP, 2, false>(d, v, f); break;
		case 4: convert_store_as<SC, SDEEP, 1, true>(d, v, f); break;
		case 5: convert_store_as<SC, SDEEP, 2, true>(d, v, f); break;
		case 6: convert_store_as<SC, SDEEP, 3, true>(d, v, f); break;
		default: convert_store_as<SC, SDEEP, 4, true>(d, v, f); break;
	}
}
// The same for up to N pixels `step` bytes apart (a thread's column of a row group, or its neighbouring columns): one
// call and one dispatch on the destination format for all of them -- per pixel, the call and the switch cost more
// than the conversion (cfg5 straight to grey: 2.40 ms per 1024 thumbnails with a call per pixel against 1.97 ms for
// the plain resize).
template <int SC, bool SDEEP, int DC, bool DDEEP, int N>
__device__ __forceinline__ void convert_store_n_as(uint8_t *d, long long step, int n, const unsigned (&v)[N][4], const FuseArgs &f) {
#pragma unroll
	for (int i = 0; i < N; ++i)
		if (i < n) convert_store_as<SC, SDEEP, DC, DDEEP>(d + i * step, v[i], f);
}
template <int SC, bool SDEEP, int N> struct PixelValues { unsigned v[N][4]; };
template <int SC, bool SDEEP, int N>
__device__ __noinline__ void convert_store_n_call(uint8_t *d, long long step, int n, PixelValues<SC, SDEEP, N> pv, FuseArgs f) {
	switch (f.dst_pixel) {   // src/picha.h:79-92
		case 0: convert_store_n_as<SC, SDEEP, 3, false, N>(d, step, n, pv.v, f); break;
		case 1: convert_store_n_as<SC, SDEEP, 4, false, N>(d, step, n, pv.v, f); break;
		case 2: convert_store_n_as<SC, SDEEP, 1, false, N>(d, step, n, pv.v, f); break;
		case 3: convert_store_n_as<SC, SDEEP, 2, false, N>(d, step, n, pv.v, f); break;
		case 4: convert_store_n_as<SC, SDEEP, 1, true, N>(d, step, n, pv.v, f); break;
		case 5: convert_store_n_as<SC, SDEEP, 2, true, N>(d, step, n, pv.v, f); break;
		case 6: convert_store_n_as<SC, SDEEP, 3, true, N>(d, step, n, pv.v, f); break;
		default: convert_store_n_as<SC, SDEEP, 4, true, N>(d, step, n, pv.v, f); break;
	}
}

template <int SC, bool SDEEP>
__device__ __forceinline__ void convert_store(uint8_t *d, const unsigned *v, const FuseArgs &f) {
	convert_store_call<SC, SDEEP>(d, v[0], v[SC > 1 ? 1 : 0], v[SC > 2 ? 2 : 0], v[SC > 3 ? 3 : 0], f);
}

}  // namespace picha_b200
#endif
