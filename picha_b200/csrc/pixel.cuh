// Device-side pixel pack / unpack, bit-identical to the reference's float round trip.
//   unpack: float(v) * (1 / max)                   src/picha.h:98-105
//   pack:   T(max(0, min(max, 0 + f*max + 0.5f)))  src/picha.h:107-114 (truncating cast)
// The *_rn intrinsics are never contracted into FMAs, whatever -fmad says.
#ifndef PICHA_B200_PIXEL_CUH
#define PICHA_B200_PIXEL_CUH

#include <stdint.h>

namespace picha_b200 {

template <bool DEEP> struct Depth;
template <> struct Depth<false> {
	typedef uint8_t type;
	static constexpr int bytes = 1;
	static constexpr float maxv = 255.0f;
	static constexpr float inv = 1 / 255.0f;
};
template <> struct Depth<true> {
	typedef uint16_t type;
	static constexpr int bytes = 2;
	static constexpr float maxv = 65535.0f;
	static constexpr float inv = 1 / 65535.0f;
};

// Both directions avoid the I2F / F2I conversion instructions, which issue at a quarter of the FP32
// rate on this part and would bound the luma conversions (tools/microbench/unpack_issue.cu):
//   unpack  0x4B000000 | v is the float 2^23 + v; fma(2^23 + v, inv, -2^23 * inv) forms the exact
//           product v * inv and rounds it once -- bit-identical to float(v) * inv;
//   pack    after the clamp t is in [0, max] so the truncating cast is floor(t): adding 2^23 with
//           round-toward-minus-infinity leaves floor(t) in the low mantissa bits.
template <bool DEEP> __device__ __forceinline__ float unpack_value(unsigned v) {
	return __fmaf_rn(__uint_as_float(0x4B000000u | v), Depth<DEEP>::inv, -8388608.0f * Depth<DEEP>::inv);
}

// The same for a value that already is the bit pattern 0x4B000000 | v.
template <bool DEEP> __device__ __forceinline__ float unpack_magic(unsigned magic_bits) {
	return __fmaf_rn(__uint_as_float(magic_bits), Depth<DEEP>::inv, -8388608.0f * Depth<DEEP>::inv);
}

// Returns floor(t) in the low 23 bits; the bits above are the exponent of 2^23 (0x4B0...), which
// every consumer drops: they store or merge only the low 8 / 16 bits.
template <bool DEEP> __device__ __forceinline__ unsigned pack_value(float f) {
	// (the reference's leading "0 +" only turns -0 into +0, which the clamp and floor do as well)
	float t = __fadd_rn(__fmul_rn(f, Depth<DEEP>::maxv), 0.5f);
	t = fmaxf(0.0f, fminf(Depth<DEEP>::maxv, t));   // NaN -> max, like std::min/std::max in the reference
	return __float_as_uint(__fadd_rd(t, 8388608.0f));
}

// Unaligned-safe channel load/store (subView bases are arbitrary byte offsets).
template <bool DEEP> __device__ __forceinline__ unsigned load_channel(const uint8_t *p) {
	if (DEEP) return (unsigned)p[0] | ((unsigned)p[1] << 8);
	return p[0];
}
template <bool DEEP> __device__ __forceinline__ void store_channel(uint8_t *p, unsigned v) {
	p[0] = (uint8_t)v;
	if (DEEP) p[1] = (uint8_t)(v >> 8);
}

}  // namespace picha_b200
#endif
