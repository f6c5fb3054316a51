/*
 * picha_oracle.c -- CPU restatement of picha's pixel hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product (picha_b200/, the C-ABI
 * library) may include, link or call this file; only tests/, the smoke check
 * and bench.py's cpu_baseline leg use it, and only as the checker.
 *
 * Parity is PINNED: this restatement is checked (tests/test_oracle.py) against
 *   - the reference's own golden fixtures (test/test2.jpg -> test/test2.png,
 *     test/test.png -> test/greytest.png; tests/golden/picha_fixtures.npz),
 *   - outputs of the reference's own C++ compiled here (oracle/_ref, built from
 *     /root/reference/src by oracle/Makefile) on a sweep of filters, formats,
 *     ratios and strides (tests/golden/ref_vectors.npz + live when present).
 *
 * Every function cites the reference lines it restates (paths relative to the
 * reference checkout).  Must be compiled with -ffp-contract=off: the reference
 * is built without FMA contraction and two of its results depend on that.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "picha_oracle.h"

/* ---- pixel formats: src/picha.h:79-92 (enum), :118-172 (traits) ---------- */

static const int kBytes[PO_NUM_PIXELS]    = {3, 4, 1, 2, 2, 4, 6, 8};
static const int kChannels[PO_NUM_PIXELS] = {3, 4, 1, 2, 1, 2, 3, 4};

int po_pixel_bytes(int p)    { return (p >= 0 && p < PO_NUM_PIXELS) ? kBytes[p] : 0; }
int po_pixel_channels(int p) { return (p >= 0 && p < PO_NUM_PIXELS) ? kChannels[p] : 0; }
int po_pixel_deep(int p)     { return p >= PO_R16; }

/* src/picha.h:212-215 */
int po_row_stride(int w, int p) { return (po_pixel_bytes(p) * w + 3) & ~3; }

/* src/picha.h:98-105: value * (1 / (max - min)), min = 0, one float multiply */
static void unpack_px(int p, const unsigned char *s, float *f) {
	int n = kChannels[p];
	if (po_pixel_deep(p)) {
		const float inv = 1 / 65535.0f;
		for (int c = 0; c < n; ++c) {
			uint16_t v;
			memcpy(&v, s + 2 * c, 2);
			f[c] = ((float)v - 0.0f) * inv;
		}
	} else {
		const float inv = 1 / 255.0f;
		for (int c = 0; c < n; ++c)
			f[c] = ((float)s[c] - 0.0f) * inv;
	}
}

/* src/picha.h:107-114: T(max(min, min(max, min + f*(max-min) + 0.5f))), truncating */
static float clampf(float v, float hi) {
	float m = (v < hi) ? v : hi;      /* std::min(hi, v) */
	return (0.0f < m) ? m : 0.0f;     /* std::max(0, m)  */
}
static void pack_px(int p, const float *f, unsigned char *d) {
	int n = kChannels[p];
	if (po_pixel_deep(p)) {
		for (int c = 0; c < n; ++c) {
			float t = 0.0f + f[c] * 65535.0f;
			t = t + 0.5f;
			uint16_t v = (uint16_t)clampf(t, 65535.0f);
			memcpy(d + 2 * c, &v, 2);
		}
	} else {
		for (int c = 0; c < n; ++c) {
			float t = 0.0f + f[c] * 255.0f;
			t = t + 0.5f;
			d[c] = (unsigned char)clampf(t, 255.0f);
		}
	}
}

/* ---- filters: src/resize.cc:200-268 ------------------------------------- */

static float mitchel_family(float B, float C, float o) {      /* :210-231 */
	float x = fabsf(o);
	if (x < 1) {
		const float a3 = (12 - 9 * B - 6 * C) / 6;
		const float a2 = (-18 + 12 * B + 6 * C) / 6;
		const float a0 = (6 - 2 * B) / 6;
		return a0 + (x * x * (a2 + x * a3));
	} else {
		const float b3 = (-B - 6 * C) / 6;
		const float b2 = (6 * B + 30 * C) / 6;
		const float b1 = (-12 * B - 48 * C) / 6;
		const float b0 = (8 * B + 24 * C) / 6;
		return b0 + (x * (b1 + x * (b2 + x * b3)));
	}
}

static float base_support(int tag) {
	switch (tag) {
		case PO_TRIANGLE: return 1.0f;                          /* :201 */
		case PO_BOX:      return 0.5f;                          /* :206 */
		default:          return 2.0f;  /* cubic :258, lanczos<2> :247, mitchel family :211 */
	}
}

static float base_eval(int tag, float o) {
	switch (tag) {
		case PO_TRIANGLE: return 1.0f - fabsf(o);                                   /* :202 */
		case PO_BOX:      return 1.0f;                                              /* :207 */
		case PO_CATMULROM: return mitchel_family(0.0f, 0.5f, o);                    /* :233-236 */
		case PO_MITCHEL:  return mitchel_family(0.333f, 0.333f, o);                 /* :238-241 */
		case PO_LANCZOS: {                                                          /* :246-253 */
			float x = o * (float)M_PI, x2 = x * x;
			return x2 == 0 ? 1.0f : 2u * sinf(x) * sinf(x / 2u) / x2;
		}
		default: {                                                                  /* cubic :257-260 */
			float a = fabsf(o);
			return 1.0f - a * a * (0.75f - 0.25f * a);
		}
	}
}

/* ScaledFilter<F>: src/resize.cc:262-268 */
typedef struct { int tag; float scale; } po_filter;
static float f_support(const po_filter *f) { return f->scale * base_support(f->tag); }
static float f_eval(const po_filter *f, float o) { return base_eval(f->tag, o / f->scale) / f->scale; }

/* ---- contribution tables: src/resize.cc:19-50 --------------------------- */

int po_make_contribs(int tag, float fwidth, int srcsize, int dstsize,
                     int *left_out, int *right_out, int *woff_out,
                     float *weights, int cap) {
	po_filter flt = { tag, fwidth };
	float scale = srcsize / (float)dstsize;                      /* :72-73 */
	float fscale = fmaxf(fmaxf(scale, 1.0f), 1.0f / f_support(&flt));
	float fsupport = f_support(&flt) * fscale;
	float iscale = 1.0f / fscale;
	int n = 0;

	float center = 0.5f * scale;
	for (int i = 0; i < dstsize; ++i, center += scale) {         /* sequential float accumulation :27 */
		float total = 0;
		int left = (int)fmaxf(0.0f, ceilf(center - fsupport));
		int right = (int)fminf((float)(srcsize - 1), floorf(center + fsupport));
		while (left < right && f_eval(&flt, (center - left) * iscale) == 0) left += 1;
		while (right > left && f_eval(&flt, (center - right) * iscale) == 0) right -= 1;
		left_out[i] = left; right_out[i] = right; woff_out[i] = n;
		int first = n;
		for (int j = left; j <= right; ++j) {
			float o = center - (float)j;
			float w = f_eval(&flt, o * iscale);
			if (n < cap) weights[n] = w;
			n++;
			total += w;
		}
		float norm = 1.0f / total;
		for (int j = first; j < n && j < cap; ++j) weights[j] *= norm;
	}
	return n;
}

/* ---- resize: src/resize.cc:66-134 (ring buffer included, so the aliasing
 *      of slot c % M when a row has more taps than M falls out by itself) --- */

int po_resize(int tag, float fwidth,
              const unsigned char *src, int sstride, int sw, int sh,
              unsigned char *dst, int dstride, int dw, int dh, int pixel) {
	if (pixel < 0 || pixel >= PO_NUM_PIXELS || sw <= 0 || sh <= 0 || dw <= 0 || dh <= 0) return -1;
	if (!(fwidth > 0)) return -1;
	const int ch = kChannels[pixel], bpp = kBytes[pixel];
	po_filter flt = { tag, fwidth };

	float xscale = sw / (float)dw, yscale = sh / (float)dh;
	float xfscale = fmaxf(fmaxf(xscale, 1.0f), 1.0f / f_support(&flt));
	float yfscale = fmaxf(fmaxf(yscale, 1.0f), 1.0f / f_support(&flt));
	float xfsupport = f_support(&flt) * xfscale;
	float yfsupport = f_support(&flt) * yfscale;
	int maxx = (int)ceilf(2 * xfsupport), maxy = (int)ceilf(2 * yfsupport);   /* :78-79 */

	/* generous capacity: every row may carry up to max+2 taps */
	int capx = (maxx + 3) * dw, capy = (maxy + 3) * dh;
	int *xl = malloc(sizeof(int) * 3 * dw), *yl = malloc(sizeof(int) * 3 * dh);
	float *xw = malloc(sizeof(float) * capx), *yw = malloc(sizeof(float) * capy);
	float *ring = calloc((size_t)maxy * dw * ch, sizeof(float));               /* :83, zero-initialised vector */
	if (!xl || !yl || !xw || !yw || !ring) { free(xl); free(yl); free(xw); free(yw); free(ring); return -2; }
	int *xr = xl + dw, *xo = xl + 2 * dw, *yr = yl + dh, *yo = yl + 2 * dh;
	po_make_contribs(tag, fwidth, sw, dw, xl, xr, xo, xw, capx);
	po_make_contribs(tag, fwidth, sh, dh, yl, yr, yo, yw, capy);

	float centery = 0.5f * yscale;                                             /* :99 */
	int srcrow = (int)fmaxf(0.0f, ceilf(centery - yfsupport));                 /* :100 */
	for (int y = 0; y < dh; ++y, centery += yscale) {
		int need = (int)(centery + yfsupport);                                  /* :104 */
		if (need > sh - 1) need = sh - 1;
		for (; srcrow <= need; ++srcrow) {                                      /* horizontal pass :105-119 */
			const unsigned char *srow = src + (size_t)srcrow * sstride;
			float *t = ring + (size_t)(srcrow % maxy) * dw * ch;
			memset(t, 0, sizeof(float) * ch * dw);
			for (int x = 0; x < dw; ++x, t += ch) {
				const float *w = xw + xo[x];
				for (int c = xl[x]; c <= xr[x]; ++c, ++w) {
					float u[4];
					unpack_px(pixel, srow + (size_t)c * bpp, u);
					for (int p = 0; p < ch; ++p) {
						float prod = *w * u[p];
						t[p] = t[p] + prod;
					}
				}
			}
		}
		unsigned char *drow = dst + (size_t)y * dstride;                        /* vertical pass :121-132 */
		for (int x = 0; x < dw; ++x, drow += bpp) {
			float acc[4] = {0, 0, 0, 0};
			const float *w = yw + yo[y];
			for (int c = yl[y]; c <= yr[y]; ++c, ++w) {
				const float *sp = ring + ((size_t)(c % maxy) * dw + x) * ch;
				for (int p = 0; p < ch; ++p) {
					float prod = *w * sp[p];
					acc[p] = acc[p] + prod;
				}
			}
			pack_px(pixel, acc, drow);
		}
	}
	free(xl); free(yl); free(xw); free(yw); free(ring);
	return 0;
}

/* ---- colour settings: src/colorconvert.h:11-14, src/colorconvert.cc:6-22 -- */

void po_resolve_color_settings(double r, double g, double b, float out[3]) {
	float rf = 0.299f, gf = 0.587f, bf = (float)0.114;
	if (r == r) rf = (float)r;      /* NaN = "not given" (:11,:14,:17) */
	if (g == g) gf = (float)g;
	if (b == b) bf = (float)b;
	float n = 1.0f / (rf + gf + bf);                                           /* :18 */
	out[0] = rf * n; out[1] = gf * n; out[2] = bf * n;
}

/* ---- colour conversion: src/colorconvert.cc:24-188 ----------------------- */

static void channel_op(int sc, int dc, const float cs[3], const float *s, float *d) {
	if (sc == dc) { for (int i = 0; i < sc; ++i) d[i] = s[i]; return; }        /* :26-32 */
	float luma = 0;
	if (sc >= 3 && dc <= 2) {                                                  /* :90,:97,:115,:122 */
		float a = s[0] * cs[0], b = s[1] * cs[1], c = s[2] * cs[2];
		luma = a + b;
		luma = luma + c;
	}
	switch (sc * 10 + dc) {
		case 12: d[0] = s[0]; d[1] = 1; break;                                  /* :34-40 */
		case 13: d[0] = d[1] = d[2] = s[0]; break;                              /* :42-49 */
		case 14: d[0] = d[1] = d[2] = s[0]; d[3] = 1; break;                    /* :51-59 */
		case 21: d[0] = s[0]; break;                                            /* :61-66 */
		case 23: d[0] = s[0]; d[1] = s[1]; d[2] = 0; break;                     /* :68-75 */
		case 24: d[0] = d[1] = d[2] = s[0]; d[3] = s[1]; break;                 /* :77-85 */
		case 31: d[0] = luma; break;                                            /* :87-92 */
		case 32: d[0] = luma; d[1] = 1; break;                                  /* :94-100 */
		case 34: d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; d[3] = 1; break;        /* :102-110 */
		case 41: d[0] = luma; break;                                            /* :112-117 */
		case 42: d[0] = luma; d[1] = s[3]; break;                               /* :119-125 */
		case 43: d[0] = s[0]; d[1] = s[1]; d[2] = s[2]; break;                  /* :127-134 */
	}
}

int po_color_convert(float r, float g, float b,
                     const unsigned char *src, int sstride, int w, int h, int spixel,
                     unsigned char *dst, int dstride, int dpixel) {
	if (spixel < 0 || spixel >= PO_NUM_PIXELS || dpixel < 0 || dpixel >= PO_NUM_PIXELS) return -1;
	if (w < 0 || h < 0) return -1;
	const int sb = kBytes[spixel], db = kBytes[dpixel];
	if (spixel == dpixel) {                                     /* :172-175 + src/picha.cc:27-34 */
		for (int y = 0; y < h; ++y)
			memcpy(dst + (size_t)y * dstride, src + (size_t)y * sstride, (size_t)w * sb);
		return 0;
	}
	const float cs[3] = { r, g, b };
	const int sc = kChannels[spixel], dc = kChannels[dpixel];
	for (int y = 0; y < h; ++y) {                               /* :143-150 */
		const unsigned char *s = src + (size_t)y * sstride;
		unsigned char *d = dst + (size_t)y * dstride;
		for (int x = 0; x < w; ++x, s += sb, d += db) {
			float in[4], out[4];
			unpack_px(spixel, s, in);
			channel_op(sc, dc, cs, in, out);
			pack_px(dpixel, out, d);
		}
	}
	return 0;
}

/* src/jpegcodec.cc:36-42: rgb[c] = int(cmyk[c]) * cmyk[3] / 255 (C integer division: truncation),
 * applied per decoded row (:96). */
int po_cmyk_to_rgb(const unsigned char *cmyk, int sstride, int w, int h, unsigned char *rgb, int dstride) {
	if (w < 0 || h < 0) return -1;
	for (int y = 0; y < h; ++y) {
		const unsigned char *s = cmyk + (size_t)y * sstride;
		unsigned char *d = rgb + (size_t)y * dstride;
		for (int i = 0; i < w; ++i, s += 4, d += 3) {
			d[0] = (unsigned char)((int)s[0] * s[3] / 255);
			d[1] = (unsigned char)((int)s[1] * s[3] / 255);
			d[2] = (unsigned char)((int)s[2] * s[3] / 255);
		}
	}
	return 0;
}
