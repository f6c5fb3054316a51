// Host model of the fast resize kernel's control flow (csrc/resize_fast.cu), used by
// tests/test_fast_model.py: same tables (tables.cc), same loop structure, plain float math.
// It checks every index the kernel would form and returns the result so it can be compared with
// the oracle on the CPU.  Not part of the product library.
//   g++ -O2 -shared -fPIC -o libfast_model.so fast_model.cc ../picha_b200/csrc/tables.cc
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <vector>

#include "../picha_b200/csrc/tables.h"

using namespace picha_b200;

extern "C" int fast_tile_width_model(const int *xfirst, const int *xcount, int dst_w, int channels, int unit, int align_px, int cap) {
	const int limit = 128 * 8 / channels;
	for (int tw = cap / unit * unit; tw >= unit; tw -= unit) {
		bool ok = true;
		for (int x0 = 0; x0 < dst_w && ok; x0 += tw) {
			const int x1 = (x0 + tw < dst_w ? x0 + tw : dst_w) - 1;
			int hi = 0, lo = xfirst[x0];
			for (int x = x0; x <= x1; ++x) {
				if (xfirst[x] + xcount[x] > hi) hi = xfirst[x] + xcount[x];
				if (xfirst[x] < lo) lo = xfirst[x];
			}
			if (lo != xfirst[x0] || hi - xfirst[x0] / align_px * align_px > limit) ok = false;
		}
		if (ok) return tw;
	}
	return 0;
}

// info[0..5] = variant, depth, tile_w, rows consumed mismatch count, max taps x, band_h
extern "C" int fast_model(int tag, float width, const uint8_t *src, int sstride, int sw, int sh, uint8_t *dst, int dstride,
                          int dw, int dh, int channels, int deep, int band_h, int force_variant, int *info) {
	AxisTable ax, ay;
	build_axis(tag, width, sw, dw, ax);
	build_axis(tag, width, sh, dh, ay);
	FastAxisY fy;
	FastAxisX fx;
	build_fast_y(ay, 12, fy);
	build_fast_x(ax, fx);
	const int bpp = channels * (deep ? 2 : 1);
	int unit = 16;
	while (unit > 1 && (unit / 2 * bpp) % 16 == 0) unit /= 2;
	const int tile_w = fy.variant < 0 ? 0 : fast_tile_width_model(fx.first.data(), fx.count.data(), dw, channels, unit, unit, 256);
	info[0] = fy.variant; info[1] = fy.depth; info[2] = tile_w; info[3] = 0; info[4] = fx.taps; info[5] = band_h;
	if (fy.variant < 0 || tile_w == 0) return 1;
	(void)force_variant;
	const float inv = deep ? 1 / 65535.0f : 1 / 255.0f, maxv = deep ? 65535.0f : 255.0f;
	const int row_values = 128 * 8;
	auto value = [&](int row, int v) -> float {   // what TMA would deliver: zero outside the image
		if (row < 0 || row >= sh) return 0.0f;
		const long long byte = (long long)v * (deep ? 2 : 1);
		if (byte + (deep ? 2 : 1) > (((long long)sw * bpp + 3) / 4) * 4) return 0.0f;
		const uint8_t *p = src + (long long)row * sstride + byte;
		unsigned x = deep ? (p[0] | (p[1] << 8)) : p[0];
		return std::fmaf(8388608.0f + (float)x, inv, -8388608.0f * inv);
	};
	int bad = 0;
	for (int y0 = 0; y0 < dh; y0 += band_h) {
		const int y1 = y0 + band_h < dh ? y0 + band_h : dh;
		for (int x0 = 0; x0 < dw; x0 += tile_w) {
			const int tw = x0 + tile_w < dw ? tile_w : dw - x0;
			const int sx0 = fx.first[x0] / unit * unit;
			const int v0 = sx0 * channels;   // first value of the tile row
			const int rlo = fy.smin[y0], rhi = fy.cum[y1 - 1];
			if (rlo < 0 || rhi >= sh || rlo > rhi) { ++bad; continue; }
			std::vector<std::vector<float>> tmp(y1 - y0, std::vector<float>(row_values, 0.0f));
			int consumed = 0;
			const int D = fy.depth;
			if (fy.variant == FastAxisY::kDown) {
				std::vector<std::vector<float>> acc(D, std::vector<float>(row_values, 0.0f));
				int r = rlo, y = fy.ybase[rlo];
				if (y > y0) ++bad;
				// fixed slots, as in csrc/resize_down.cuh: output y accumulates in slot y % D and the
				// weights of a source row are laid out in slot order
				while (y < y1) {
					const int need = fy.cum[y];
					for (; r <= need; ++r) {
						if (r > rhi) ++bad;
						++consumed;
						std::vector<float> ws(D, 0.0f);
						for (int j = 0; j < D; ++j) ws[(fy.ybase[r] + j) % D] = fy.wv[(size_t)r * fy.stride + j];
						if (fy.ybase[r] > y) ++bad;            // row r only feeds outputs y .. y + D - 1
						for (int j = 0; j < D; ++j)
							for (int i = 0; i < row_values; ++i) acc[j][i] = std::fmaf(ws[j], value(r, v0 + i), acc[j][i]);
					}
					const int s = y % D;
					if (y >= y0) tmp[y - y0] = acc[s];
					std::fill(acc[s].begin(), acc[s].end(), 0.0f);
					++y;
				}
				if (consumed != rhi - rlo + 1) ++bad;
			} else {
				std::vector<std::vector<float>> win(D, std::vector<float>(row_values, 0.0f));
				int rb = fy.lo[y0], rnext = rb, s = 0;
				auto load = [&](std::vector<float> &w) {
					for (int i = 0; i < row_values; ++i) w[i] = value(rnext, v0 + i);
					++rnext; ++consumed;
				};
				for (int k = 0; k < D; ++k) {
					if (rnext <= rhi) load(win[k]);
					else std::fill(win[k].begin(), win[k].end(), 0.0f);
				}
				int y = y0;
				while (y < y1) {
					while (y < y1 && fy.lo[y] == rb) {
						std::vector<float> o(row_values, 0.0f);
						for (int k = 0; k < D; ++k) {
							const float w = fy.wv[(size_t)y * fy.stride + k];
							for (int i = 0; i < row_values; ++i) o[i] = std::fmaf(w, win[(s + k) % D][i], o[i]);
						}
						tmp[y - y0] = o;
						++y;
					}
					if (y < y1) {
						if (fy.lo[y] < rb) { ++bad; break; }
						if (rnext <= rhi) load(win[s]);
						else std::fill(win[s].begin(), win[s].end(), 0.0f);
						++rb;
						s = (s + 1) % D;
					}
				}
				if (consumed > rhi - rlo + 1) ++bad;
			}
			for (int y = y0; y < y1; ++y)
				for (int xx = 0; xx < tw; ++xx) {
					const int x = x0 + xx, first = fx.first[x] - sx0, cnt = fx.count[x];
					if (first < 0 || (first + cnt) * channels > row_values) { ++bad; continue; }
					for (int ch = 0; ch < channels; ++ch) {
						float a = 0.0f;
						for (int k = 0; k < cnt; ++k)
							a = std::fmaf(fx.w[(size_t)x * fx.stride + k], tmp[y - y0][(first + k) * channels + ch], a);
						float t = std::fmaf(a, maxv, 0.5f);
						t = std::fmin(std::fmax(t, 0.0f), maxv);
						unsigned q = (unsigned)t;
						uint8_t *d = dst + (long long)y * dstride + (long long)x * bpp + ch * (deep ? 2 : 1);
						d[0] = (uint8_t)q;
						if (deep) d[1] = (uint8_t)(q >> 8);
					}
				}
		}
	}
	info[3] = bad;
	return 0;
}
