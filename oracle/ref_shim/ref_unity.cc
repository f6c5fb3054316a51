// TEST INFRASTRUCTURE ONLY.
//
// Compiles the reference's own hot-path sources *where they lie* (they are
// #included by absolute path from the read-only reference checkout; nothing is
// copied into this repository) into oracle/_ref/libpicha_ref.so, so the C
// restatement in oracle/picha_oracle.c and the CUDA product can be compared
// against the real reference arithmetic.  PICHA_REF_SRC is supplied by the
// Makefile (-DPICHA_REF_SRC=/root/reference/src).
#define PICHA_STR2(x) #x
#define PICHA_STR(x) PICHA_STR2(x)
#define PICHA_INC(f) PICHA_STR(PICHA_REF_SRC/f)

#include PICHA_INC(resize.cc)
#include PICHA_INC(colorconvert.cc)

namespace picha {

// Storage for the interned-symbol externs the glue names (never dereferenced).
#define SSYMBOL(a) Nan::Persistent<String> a##_symbol;
STATIC_SYMBOLS
#undef SSYMBOL

// Row copy used by doColorConvert for same-format inputs: payload bytes only.
void NativeImage::copy(NativeImage& o) {
	const size_t payload = size_t(width) * pixelBytes(pixel);
	for (int y = 0; y < height; ++y)
		memcpy(row(y), o.row(y), payload);
}

// Glue referenced by the NAN methods; unreachable from the exported entry points.
NativeImage jsImageToNativeImage(Local<Object>&) { return NativeImage(); }
Local<Object> newJsImage(int, int, PixelMode) { return Local<Object>(); }
PixelMode pixelSymbolToEnum(Local<Value>) { return INVALID_PIXEL; }
void makeCallback(Local<Function>, const char*, Local<Value>) {}

}  // namespace picha

namespace {
picha::NativeImage wrap(char* data, int stride, int w, int h, int pixel) {
	picha::NativeImage im;
	im.data = data; im.stride = stride; im.width = w; im.height = h;
	im.pixel = static_cast<picha::PixelMode>(pixel);
	return im;
}
}

extern "C" {

// -> picha::resizeImage(const ResizeOptions&, NativeImage&, NativeImage&)
void ref_resize(int filter, float fwidth,
                char* s, int sstride, int sw, int sh,
                char* d, int dstride, int dw, int dh, int pixel) {
	picha::ResizeOptions o;
	o.filter = static_cast<picha::ResizeFilterTag>(filter);
	o.width = fwidth;
	picha::NativeImage a = wrap(s, sstride, sw, sh, pixel);
	picha::NativeImage b = wrap(d, dstride, dw, dh, pixel);
	picha::resizeImage(o, a, b);
}

// -> picha::doColorConvert(const ColorSettings&, NativeImage&, NativeImage&)
// The caller passes already-normalised weights (what getSettings produces).
void ref_colorconvert(float r, float g, float b,
                      char* s, int sstride, int w, int h, int spixel,
                      char* d, int dstride, int dpixel) {
	picha::ColorSettings cs;
	cs.rFactor = r; cs.gFactor = g; cs.bFactor = b;
	picha::NativeImage a = wrap(s, sstride, w, h, spixel);
	picha::NativeImage c = wrap(d, dstride, w, h, dpixel);
	picha::doColorConvert(cs, a, c);
}

// Table-level probe: the reference's makeContribs for one axis.
// Returns the number of weights written (or the number needed if cap is short).
int ref_contribs(int filter, float fwidth, int srcsize, int dstsize,
                 int* left, int* right, int* woff, float* weights, int cap) {
	using namespace picha;
	RangeVector ranges; ranges.resize(dstsize);
	PixelContribs storage;
	float scale = srcsize / float(dstsize);
	switch (filter) {
		case CubicFilterTag: makeContribs(ranges, ScaledFilter<CubicFilter>(fwidth), scale, storage, srcsize); break;
		case LanczosFilterTag: makeContribs(ranges, ScaledFilter<LanczosFilter>(fwidth), scale, storage, srcsize); break;
		case CatmulRomFilterTag: makeContribs(ranges, ScaledFilter<CatmulRomFilter>(fwidth), scale, storage, srcsize); break;
		case MitchelFilterTag: makeContribs(ranges, ScaledFilter<MitchelFilter>(fwidth), scale, storage, srcsize); break;
		case BoxFilterTag: makeContribs(ranges, ScaledFilter<BoxFilter>(fwidth), scale, storage, srcsize); break;
		case TriangleFilterTag: makeContribs(ranges, ScaledFilter<TriangleFilter>(fwidth), scale, storage, srcsize); break;
		default: return -1;
	}
	for (int i = 0; i < dstsize; ++i) {
		left[i] = ranges[i].left; right[i] = ranges[i].right; woff[i] = ranges[i].weights;
	}
	int n = int(storage.size());
	for (int i = 0; i < n && i < cap; ++i) weights[i] = storage[i];
	return n;
}

}  // extern "C"
