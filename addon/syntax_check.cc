// Syntax check of the glue against the reference's own NativeImage / ResizeOptions / ColorSettings
// (compiled with the inert v8/node/nan stand-ins of oracle/ref_shim; never linked or run):
//   g++ -std=c++14 -fsyntax-only -w -Ioracle/ref_shim -Iinclude -I/root/reference/src addon/syntax_check.cc
#include "picha.h"
#include "colorconvert.h"
namespace picha {
// the two types the reference defines inside resize.cc (src/resize.cc:151-177)
enum ResizeFilterTag { CubicFilterTag, LanczosFilterTag, CatmulRomFilterTag, MitchelFilterTag, BoxFilterTag, TriangleFilterTag, InvalidFilterTag };
struct ResizeOptions { ResizeOptions() : filter(CubicFilterTag), width(0.70f) {} ResizeFilterTag filter; float width; };
}
#include "../addon/picha_b200_glue.h"

int check(picha::NativeImage &a, picha::NativeImage &b) {
	picha::ResizeOptions o;
	picha::ColorSettings cs;
	int rc = picha_b200::resize(o, a, b);
	if (rc) return rc;
	rc = picha_b200::colorConvert(cs, a, b);
	return rc ? (picha_b200::message(rc) != 0) : 0;
}
