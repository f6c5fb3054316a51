// Glue between picha's Node addon and the B200 library (include/picha_b200.h).
//
// picha's NAN methods stay as they are; only the two calls into the CPU core change:
//     resizeImage(opts, src, dst)       src/resize.cc:293 (UV_resize), :399 (resizeSync)
//     doColorConvert(cs, src, dst)      src/colorconvert.cc:201 (UV_colorConvert), :288 (colorConvertSync)
// become picha_b200::resize(...) / picha_b200::colorConvert(...), which return a status instead of
// void (the CPU core cannot fail; a GPU call can).  See INTEGRATION.md for the patch.
//
// This header only needs picha's own picha.h (NativeImage, PixelMode); it has been syntax-checked
// against inert v8/node/nan stand-ins (no Node toolchain exists in the build image) -- UNVERIFIED
// against a real Node build.
#ifndef PICHA_B200_GLUE_H
#define PICHA_B200_GLUE_H

#include "picha_b200.h"

namespace picha_b200 {

// NativeImage (src/picha.h:202-218) and picha_b200_image have the same fields; PixelMode and
// enum picha_b200_pixel the same numeric values (src/picha.h:79-92).
template <class NativeImage> inline picha_b200_image wrap(const NativeImage &im) {
	picha_b200_image r;
	r.data = im.data;
	r.stride = im.stride;
	r.width = im.width;
	r.height = im.height;
	r.pixel = static_cast<int32_t>(im.pixel);
	return r;
}

// Drop-in for picha::resizeImage(const ResizeOptions&, NativeImage&, NativeImage&).
// ResizeOptions::filter is ResizeFilterTag (src/resize.cc:151-160) == enum picha_b200_filter.
template <class ResizeOptions, class NativeImage>
inline int resize(const ResizeOptions &opts, NativeImage &src, NativeImage &dst) {
	picha_b200_image s = wrap(src), d = wrap(dst);
	return picha_b200_resize(&s, &d, static_cast<int>(opts.filter), opts.width);
}

// Drop-in for picha::doColorConvert(const ColorSettings&, NativeImage&, NativeImage&).
template <class ColorSettings, class NativeImage>
inline int colorConvert(const ColorSettings &cs, NativeImage &src, NativeImage &dst) {
	picha_b200_image s = wrap(src), d = wrap(dst);
	return picha_b200_color_convert(&s, &d, cs.rFactor, cs.gFactor, cs.bFactor);
}

// Message for a thrown Error (sync) or cb(err) (async).
inline const char *message(int status) {
	const char *detail = picha_b200_last_error();
	return (status == PICHA_B200_ERR_CUDA && detail && *detail) ? detail : picha_b200_strerror(status);
}

}  // namespace picha_b200
#endif
