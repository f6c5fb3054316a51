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

// Vertical axis re-expressed for the fast kernel's first pass, where every thread owns a few
// source columns and walks down the rows (csrc/resize_fast.cu).  Two equivalent forms of the same
// banded matrix (taps that alias to one effective row are merged by adding their weights):
//   kDown  "accumulator ring": source row r adds wv[r*stride + j] * value into output
//          ybase[r] + j, j < depth; output y is complete once rows <= cum[y] are consumed.
//   kUp    "row window": output y = sum_k wv[y*stride + k] * value(row lo[y] + k), k < depth.
// All arrays are band-independent, so the launch may cut the image into bands of any height.
struct FastAxisY {
	enum { kNone = -1, kDown = 0, kUp = 1 };
	int variant = kNone;
	int depth = 0;    // accumulators (kDown) or window rows (kUp) needed
	int stride = 0;   // depth rounded up to a multiple of 4 (float4 loads)
	std::vector<int> cum;    // [dst] running max of the last effective row of outputs <= y
	std::vector<int> smin;   // [dst] first effective row touched by any output >= y
	std::vector<int> ybase;  // [src] kDown: first output still open when row r is consumed
	std::vector<int> lo;     // [dst] kUp: window base (= smin)
	std::vector<int> done;   // [src + 8] kDown: how many outputs are complete once row r is consumed (cum[y] == r)
	std::vector<float> wv;
};
void build_fast_y(const AxisTable &y, int max_depth, FastAxisY &out);

// Horizontal axis for the second pass: weights padded to a fixed row length.
struct FastAxisX {
	int taps = 0;            // max taps of any output column
	int stride = 0;          // row length >= taps: a multiple of 4 that is 4 mod 8
	std::vector<int> first, count;   // [dst] tap range without end taps below 2^-30 (see tables.cc)
	std::vector<float> w;    // [dst][stride]
	// The same rows with duplicates removed (bitwise equal tap counts and weights: every column of an
	// integer-ratio resize away from the edges, every q-th column of a p/q ratio): the downscaling
	// kernel keeps only these in shared memory when there are few of them.
	std::vector<int> urow;   // [dst] index of the column's row in uw
	std::vector<float> uw;   // [unique][stride]
	int unique = 0;
};
void build_fast_x(const AxisTable &x, FastAxisX &out);

// Odd channel counts (grey, rgb) run the horizontal pass on the row as a flat float array read in
// aligned float4 chunks: a column then needs its weight row expanded to one weight per float
// (each tap repeated `channels` times) and shifted by the column's misalignment off = (first *
// channels) mod 4.  The expansion happens on the device; this is the list of distinct (row, off)
// pairs and each column's entry in it.
struct FlatRows {
	std::vector<int> col;       // [dst] index into src/off
	std::vector<int> src, off;  // [count] weight row (FastAxisX::urow numbering), shift in floats
};
void build_flat_rows(const FastAxisX &x, int channels, FlatRows &out);

// Upscaling kernel, wide-window variant (csrc/resize_up.cuh, WPX = 0): a thread owns 4 consecutive output columns and
// walks the source pixels [first[4g], first[4g] + window) of every row; block g holds, for each of those pixels, the
// weight it carries into each of the 4 columns (zero where it is not one of the column's taps), scaled by `scale`.
// window = 0: no such table (some group's pixels do not start at its first column's, the widest group exceeds cap, or
// no group spans more than the 8 pixels the regular variant handles from registers).
struct WideBlocks {
	int window = 0;
	std::vector<float> w;   // [(dst + 3) / 4][window][4]
};
void build_wide_blocks(const FastAxisX &x, int dst_size, int cap, float scale, WideBlocks &out);

// Filter tag order: src/resize.cc:151-160.  `width` is ResizeOptions::width (ScaledFilter scale).
void build_axis(int filter_tag, float width, int src_size, int dst_size, AxisTable &out);

}  // namespace picha_b200
#endif
