// Host-side contribution tables for the separable resize.
//
// The weights must equal the reference's bit for bit (they come out of float sinf/ceilf/floorf
// and a float-accumulated centre), so they are built on the CPU and uploaded; the GPU never
// evaluates a filter.  What the reference computes: makeContribs, src/resize.cc:19-50, with the
// filters of src/resize.cc:200-268.
#ifndef PICHA_B200_TABLES_H
#define PICHA_B200_TABLES_H

#include <vector>

namespace picha_b200 {

enum { kNumFilters = 6, kNumPixels = 8 };

struct PixelInfo { int bytes, channels, deep; };
// src/picha.h:118-172
inline PixelInfo pixel_info(int p) {
	static const PixelInfo t[kNumPixels] = {
		{3, 3, 0}, {4, 4, 0}, {1, 1, 0}, {2, 2, 0}, {2, 1, 1}, {4, 2, 1}, {6, 3, 1}, {8, 4, 1}};
	if (p < 0 || p >= kNumPixels) return PixelInfo{0, 0, 0};
	return t[p];
}

// One axis (x or y) of a resize: for output coordinate i the taps are source coordinates
// first[i] .. first[i] + count[i] - 1 with weights w[start[i] + k], already normalised.
struct AxisTable {
	int src_size = 0, dst_size = 0;
	float scale = 0, fsupport = 0;
	int ring = 0;                 // M = ceil(2*fsupport): rows in the reference's FloatBuffer (resize.cc:79,83)
	int max_taps = 0;
	std::vector<int> first, count, start;
	std::vector<float> w;
	// Vertical use only: eff[start[i] + k] is the source row whose horizontally-filtered values
	// sit in ring slot (first[i]+k) % M when output row i is produced (resize.cc:108,126); it
	// differs from first[i]+k exactly when a row has more than M taps.
	std::vector<int> eff;
	std::vector<int> need;        // needrow(i), resize.cc:104
};

// Filter tag order: src/resize.cc:151-160.  `width` is ResizeOptions::width (ScaledFilter scale).
void build_axis(int filter_tag, float width, int src_size, int dst_size, AxisTable &out);

}  // namespace picha_b200
#endif
