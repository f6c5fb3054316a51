// See tables.h.  Compiled with -ffp-contract=off: the reference build has no FMA.
#include "tables.h"

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>

namespace picha_b200 {
namespace {

// Piecewise cubic with the (B, C) parametrisation; src/resize.cc:210-231.
struct BCSpline {
	float p0, p2, p3, q0, q1, q2, q3;
	BCSpline(float B, float C) {
		p3 = (12 - 9 * B - 6 * C) / 6;
		p2 = (-18 + 12 * B + 6 * C) / 6;
		p0 = (6 - 2 * B) / 6;
		q3 = (-B - 6 * C) / 6;
		q2 = (6 * B + 30 * C) / 6;
		q1 = (-12 * B - 48 * C) / 6;
		q0 = (8 * B + 24 * C) / 6;
	}
	float at(float o) const {
		float x = std::fabs(o);
		if (x < 1) return p0 + (x * x * (p2 + x * p3));
		return q0 + (x * (q1 + x * (q2 + x * q3)));
	}
};

class Kernel1D {
public:
	Kernel1D(int tag, float width) : tag_(tag), width_(width), catmul_(0.0f, 0.5f), mitchel_(0.333f, 0.333f) {}

	// ScaledFilter::support, src/resize.cc:266
	float support() const {
		float base = 2.0f;                       // cubic, lanczos<2>, catmulrom, mitchel
		if (tag_ == 5) base = 1.0f;              // triangle, :201
		else if (tag_ == 4) base = 0.5f;         // box, :206
		return width_ * base;
	}
	// ScaledFilter::operator(), src/resize.cc:267
	float operator()(float f) const { return base(f / width_) / width_; }

private:
	float base(float o) const {
		switch (tag_) {
			case 1: {                            // lanczos A=2, :249-252
				float x = o * float(M_PI), x2 = x * x;
				return x2 == 0 ? 1.0f : 2u * std::sin(x) * std::sin(x / 2u) / x2;
			}
			case 2: return catmul_.at(o);
			case 3: return mitchel_.at(o);
			case 4: return 1.0f;                 // box has no cut-off, :207
			case 5: return 1.0f - std::fabs(o);  // :202
			default: {                           // cubic, :259
				float a = std::fabs(o);
				return 1.0f - a * a * (0.75f - 0.25f * a);
			}
		}
	}
	int tag_;
	float width_;
	BCSpline catmul_, mitchel_;
};

}  // namespace

void build_axis(int filter_tag, float width, int src_size, int dst_size, AxisTable &t) {
	const Kernel1D k(filter_tag, width);
	t = AxisTable();
	t.src_size = src_size;
	t.dst_size = dst_size;
	t.scale = src_size / float(dst_size);                                   // resize.cc:72-73
	const float fscale = std::fmax(std::fmax(t.scale, 1.0f), 1.0f / k.support());
	t.fsupport = k.support() * fscale;
	t.ring = int(std::ceil(2 * t.fsupport));
	const float inv = 1.0f / fscale;

	t.first.resize(dst_size);
	t.count.resize(dst_size);
	t.start.resize(dst_size);
	t.need.resize(dst_size);
	t.w.reserve(size_t(t.ring + 2) * dst_size);

	float centre = 0.5f * t.scale;
	for (int i = 0; i < dst_size; ++i, centre += t.scale) {                 // float accumulation on purpose
		int lo = int(std::fmax(0.0f, std::ceil(centre - t.fsupport)));
		int hi = int(std::fmin(float(src_size - 1), std::floor(centre + t.fsupport)));
		while (lo < hi && k((centre - lo) * inv) == 0) ++lo;                 // exact-zero end taps are dropped
		while (hi > lo && k((centre - hi) * inv) == 0) --hi;
		t.first[i] = lo;
		t.count[i] = hi - lo + 1;
		t.start[i] = int(t.w.size());
		if (t.count[i] > t.max_taps) t.max_taps = t.count[i];
		float sum = 0;
		for (int j = lo; j <= hi; ++j) {
			float wgt = k((centre - float(j)) * inv);
			t.w.push_back(wgt);
			sum += wgt;
		}
		const float norm = 1.0f / sum;   // the reference asserts sum > 0 in debug builds and divides regardless in release
		for (size_t j = t.start[i]; j < t.w.size(); ++j) t.w[j] *= norm;

		int need = int(centre + t.fsupport);                                // resize.cc:104
		t.need[i] = need < src_size - 1 ? need : src_size - 1;
	}

	// Ring model: when output row i is produced every source row <= need[i] has been written
	// to slot row % M, so slot c % M holds the newest such row congruent to c.
	t.eff.resize(t.w.size());
	const int M = t.ring > 0 ? t.ring : 1;
	for (int i = 0; i < dst_size; ++i)
		for (int q = 0; q < t.count[i]; ++q) {
			int c = t.first[i] + q;
			int lag = t.need[i] - c;
			t.eff[t.start[i] + q] = lag >= 0 ? c + M * (lag / M) : c;
		}
}

// End taps the fast path leaves out: the reference keeps every tap whose weight is not exactly zero
// (resize.cc:31-34), which at integer ratios includes end taps of ~1e-16 (sin(2*pi) in float).  A
// tap below 2^-30 moves a result by < 1e-6 LSB -- far inside the fast path's +-1 LSB contract --
// but costs a whole accumulator slot per thread in the vertical pass.
static void pruned_range(const AxisTable &t, int i, int &k0, int &k1) {
	const float eps = 9.3132257e-10f;   // 2^-30; the weights of an output sum to 1
	k0 = 0;
	k1 = t.count[i];
	while (k1 - k0 > 1 && std::fabs(t.w[t.start[i] + k0]) < eps) ++k0;
	while (k1 - k0 > 1 && std::fabs(t.w[t.start[i] + k1 - 1]) < eps) --k1;
}

void build_fast_y(const AxisTable &t, int max_depth, FastAxisY &f) {
	f = FastAxisY();
	const int n = t.dst_size;
	std::vector<int> first(n), last(n), pk0(n), pk1(n);
	for (int y = 0; y < n; ++y) {
		pruned_range(t, y, pk0[y], pk1[y]);
		int lo = t.src_size, hi = -1;
		for (int k = pk0[y]; k < pk1[y]; ++k) {
			int r = t.eff[t.start[y] + k];
			if (r < lo) lo = r;
			if (r > hi) hi = r;
		}
		first[y] = lo;
		last[y] = hi;
	}
	f.cum.resize(n);
	f.smin.resize(n);
	for (int y = 0, m = -1; y < n; ++y) { if (last[y] > m) m = last[y]; f.cum[y] = m; }
	for (int y = n - 1, m = t.src_size; y >= 0; --y) { if (first[y] < m) m = first[y]; f.smin[y] = m; }
	f.lo = f.smin;

	// kDown: which output is open when row r arrives
	f.ybase.assign(t.src_size, n);
	for (int r = 0, y = 0; r < t.src_size; ++r) {
		while (y < n && f.cum[y] < r) ++y;
		f.ybase[r] = y;
	}
	f.done.assign(t.src_size + 8, 0);
	for (int y = 0; y < n; ++y) ++f.done[f.cum[y]];
	int need_down = 0, need_up = 0;
	for (int y = 0; y < n; ++y) {
		for (int k = pk0[y]; k < pk1[y]; ++k) {
			int d = y - f.ybase[t.eff[t.start[y] + k]] + 1;
			if (d > need_down) need_down = d;
		}
		int span = last[y] - f.lo[y] + 1;
		if (span > need_up) need_up = span;
	}
	const bool down_ok = need_down <= max_depth, up_ok = need_up <= max_depth;
	if (!down_ok && !up_ok) return;
	// the form with the smaller register footprint; ties go with the direction of the scale
	bool use_down = down_ok && (!up_ok || need_down < need_up || (need_down == need_up && t.scale >= 1.0f));
	f.variant = use_down ? FastAxisY::kDown : FastAxisY::kUp;
	f.depth = use_down ? need_down : need_up;
	f.stride = (f.depth + 3) & ~3;
	if (use_down) {
		// 8 extra zero rows: the kernel copies the weights of whole 8-row stages
		f.wv.assign(size_t(t.src_size + 8) * f.stride, 0.0f);
		for (int y = 0; y < n; ++y)
			for (int k = pk0[y]; k < pk1[y]; ++k) {
				int r = t.eff[t.start[y] + k];
				f.wv[size_t(r) * f.stride + (y - f.ybase[r])] += t.w[t.start[y] + k];
			}
	} else {
		f.wv.assign(size_t(n) * f.stride, 0.0f);
		for (int y = 0; y < n; ++y)
			for (int k = pk0[y]; k < pk1[y]; ++k) {
				int r = t.eff[t.start[y] + k];
				f.wv[size_t(y) * f.stride + (r - f.lo[y])] += t.w[t.start[y] + k];
			}
	}
}

void build_fast_x(const AxisTable &t, FastAxisX &f) {
	f = FastAxisX();
	f.first.resize(t.dst_size);
	f.count.resize(t.dst_size);
	std::vector<int> k0(t.dst_size);
	for (int x = 0; x < t.dst_size; ++x) {
		int k1;
		pruned_range(t, x, k0[x], k1);
		f.first[x] = t.first[x] + k0[x];
		f.count[x] = k1 - k0[x];
		if (f.count[x] > f.taps) f.taps = f.count[x];
	}
	f.stride = (f.taps + 3) & ~3;          // float4 reads; 4 mod 8 keeps 8 neighbouring rows in distinct banks
	if (f.stride % 8 == 0) f.stride += 4;
	f.w.assign(size_t(t.dst_size) * f.stride, 0.0f);
	for (int x = 0; x < t.dst_size; ++x)
		for (int k = 0; k < f.count[x]; ++k) f.w[size_t(x) * f.stride + k] = t.w[t.start[x] + k0[x] + k];
	// unique rows (the padded rows compare equal exactly when counts and weights do: padding is zero and
	// a pruned row never ends in an exact zero unless it is a single tap)
	f.urow.resize(t.dst_size);
	std::map<std::vector<uint32_t>, int> seen;
	std::vector<uint32_t> key(f.stride + 1);
	for (int x = 0; x < t.dst_size; ++x) {
		key[0] = (uint32_t)f.count[x];
		memcpy(&key[1], &f.w[size_t(x) * f.stride], f.stride * sizeof(float));
		auto it = seen.find(key);
		if (it == seen.end()) {
			it = seen.emplace(key, f.unique++).first;
			f.uw.insert(f.uw.end(), f.w.begin() + size_t(x) * f.stride, f.w.begin() + size_t(x + 1) * f.stride);
		}
		f.urow[x] = it->second;
	}
}

void build_flat_rows(const FastAxisX &x, int channels, FlatRows &f) {
	f = FlatRows();
	const int n = (int)x.first.size();
	f.col.resize(n);
	std::map<std::pair<int, int>, int> seen;
	for (int i = 0; i < n; ++i) {
		const std::pair<int, int> key(x.urow[i], (x.first[i] * channels) & 3);
		auto it = seen.find(key);
		if (it == seen.end()) {
			it = seen.emplace(key, (int)f.src.size()).first;
			f.src.push_back(key.first);
			f.off.push_back(key.second);
		}
		f.col[i] = it->second;
	}
}

void build_wide_blocks(const FastAxisX &x, int dst_size, int cap, float scale, WideBlocks &out) {
	out.window = 0;
	out.w.clear();
	const int groups = (dst_size + 3) / 4;
	int window = 0;
	for (int g = 0; g < groups; ++g) {
		const int lo = x.first[4 * g];
		for (int px = 4 * g; px < dst_size && px < 4 * g + 4; ++px) {
			if (x.first[px] < lo) return;              // the walk is anchored at the group's first column
			window = std::max(window, x.first[px] + x.count[px] - lo);
		}
	}
	if (window <= 8 || window > cap) return;         // (up to 8 pixels the kernel keeps the block in registers: no table)
	out.window = window;
	out.w.assign((size_t)groups * window * 4, 0.0f);
	for (int g = 0; g < groups; ++g) {
		const int lo = x.first[4 * g];
		for (int p = 0; p < 4 && 4 * g + p < dst_size; ++p) {
			const int px = 4 * g + p;
			for (int k = 0; k < x.count[px]; ++k)
				out.w[((size_t)g * window + (x.first[px] - lo + k)) * 4 + p] = x.w[(size_t)px * x.stride + k] * scale;
		}
	}
}

}  // namespace picha_b200
