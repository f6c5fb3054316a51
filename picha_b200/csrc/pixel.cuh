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

template <bool DEEP> __device__ __forceinline__ float unpack_value(unsigned v) {
	return __fmul_rn(__uint2float_rn(v), Depth<DEEP>::inv);
}

template <bool DEEP> __device__ __forceinline__ unsigned pack_value(float f) {
	float t = __fadd_rn(__fadd_rn(0.0f, __fmul_rn(f, Depth<DEEP>::maxv)), 0.5f);
	t = fmaxf(0.0f, fminf(Depth<DEEP>::maxv, t));
	return (unsigned)t;   // cvt.rzi: truncation, like the C++ cast
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
