// Downscaling variant of the fast resize path (reference: src/resize.cc:66-134); design notes in
// resize_fast.cu and DESIGN.md section 5.3.  Included by the resize_down_*.cu instantiation units.
//
// Compared with the generic kernel of resize_fast.cuh this one is built around the instruction
// budget of the vertical pass, which is what bounds a wide-filter downscale on this part:
//   * 64 threads x 16 channel values per row: per-row control and weight loads are amortised
//     over twice as many FMAs.
//   * No unpack arithmetic.  A byte (or 16-bit value) v moved into the low bits of an otherwise
//     zero word IS the float v * 2^-149 (a subnormal, exact), and the FMA pipe takes subnormal
//     inputs at full speed: one PRMT per value replaces PRMT + FMA.  The vertical weights carry
//     2^120 so the sums are ordinary normal floats (w*v * 2^-29); the horizontal weights carry
//     2^29 / max, so the second pass lands on the [0, 1] scale the pack expects.  Every product is
//     formed exactly inside the FMA, so nothing is lost relative to unpacking first.
//   * The accumulators are packed pairs and the vertical MACs fma.rn.f32x2 with the weight as a broadcast
//     uniform operand (FFMA2 R, R.F32x2, UR.F32, R.F32x2): the same FMA rate (an FFMA2 takes two issue
//     cycles) with half the instructions in flight -- measured 10 % faster end to end than scalar FFMAs.
//   * Accumulator slots are fixed (output row y lives in slot y % DEPTH) and the host lays the
//     weights of a source row out in slot order: the row body is the same code for every row --
//     no rotation by unrolling, no per-row branches; only the (rare) emit picks a slot.
//   * The row loop decides nothing by arithmetic: the host puts a flag word behind every source row's weights (how
//     many outputs the row completes, whether the ring stage ends with the next row).  Per emit slot there are two
//     row bodies -- the first row after an emit, which clears the slot on the way, and one loop body -- because the
//     row loop's code is what fills the instruction cache.
//   * Ring stages are handed back through `empty` mbarriers (no CTA-wide barrier in pass 1).
//   * Pass 2 comes in four forms: by columns (pass2_cols: 1-, 3-, 4-channel pixels), general
//     (pass2: 2-channel pixels, 8-row groups), integer ratios of 4-channel pixels (pass2_int4), and each of them
//     converting (FUSED); all store pixels straight to global memory where the destination's alignment allows.
#ifndef PICHA_B200_RESIZE_DOWN_CUH
#define PICHA_B200_RESIZE_DOWN_CUH

#include "pixel_convert.cuh"
#include "resize_fast.cuh"

namespace picha_b200 {
namespace down {

using fast::lds;
using fast::sts;
using fast::smem;
using fast::smem_u32;
using fast::VTable;
using fast::kMaxBands;
using fast::kYtabMax;
using fast::kWtMax;

#ifndef PICHA_DOWN_RS
#define PICHA_DOWN_RS 8
#endif
constexpr int NT = 64;             // threads per CTA
constexpr int NV = 16;             // channel values of one source row owned by a thread
constexpr int ROWV = NT * NV;      // values per staged row (the tile geometry of the generic kernel)
static_assert(ROWV == fast::NT * fast::NV, "tile widths are planned once for both kernels");
// Output rows per pass-2 group: 4 or 8 (template parameter GR).  With 8, the eight lanes that share a
// shared-memory phase of a float4 read are the eight rows of ONE column -- conflict-free whatever
// the ratio; with 4 they are four rows of two neighbouring columns, which is conflict-free only when
// the columns' windows start the right distance apart (e.g. 4:1 rgba) but costs half the memory.
#ifndef PICHA_DOWN_ODD_PAD
#define PICHA_DOWN_ODD_PAD 8
#endif
// floats per intermediate row: the rows of a group land 4 banks apart (8 for odd channel counts, whose
// neighbouring columns start 20 or 24 floats apart at a 7.5:1 ratio)
__host__ __device__ constexpr int tmps(int channels, int group, int nt = NT) {
	return nt * NV + (group == 4 && (channels & 1) ? PICHA_DOWN_ODD_PAD : 4);
}
// Wider CTAs (96 or 128 threads: 8-bit formats, general horizontal pass): a tile's source span is 1536 or 2048 bytes
// instead of 1024, for shapes where 1024-byte tiles divide the row badly (1080p rgb to 256 columns needs 8 tiles
// of 341 pixels -- 42 % more columns than the image has -- but only 4 of 512).  A staged row then arrives as two
// TMA boxes side by side, like the 16-bit rows of the 64-thread kernel.
__host__ __device__ constexpr int stage_bytes(int nt) { return PICHA_DOWN_RS * 16 * nt; }
__host__ __device__ constexpr int row_boxes(bool deep, int nt) { return deep || nt > 64 ? 2 : 1; }
__host__ __device__ constexpr int box_bytes(bool deep, int nt) { return (deep ? 32 : 16) * nt / row_boxes(deep, nt); }
#ifndef PICHA_DOWN_FRESH
#define PICHA_DOWN_FRESH 1       // 0: every emit zeroes its slot itself (the round-2 loop before the `fresh` row body; for A/B builds)
#endif
#ifndef PICHA_DOWN_ONE_BODY
#define PICHA_DOWN_ONE_BODY 1    // 0: the row loop unrolled over two row buffers (for A/B builds)
#endif
#ifndef PICHA_DOWN_SPLIT_LOAD
#define PICHA_DOWN_SPLIT_LOAD 1    // 0: the whole next row is loaded in front of the body (for A/B builds)
#endif
#ifndef PICHA_DOWN_COLS
#define PICHA_DOWN_COLS 2        // horizontal pass by columns: 2 every channel count, 1 odd ones only, 0 never (for A/B builds)
#endif
#ifndef PICHA_DOWN_NS
#define PICHA_DOWN_NS 2
#endif
constexpr int NS = PICHA_DOWN_NS;  // ring stages
#ifndef PICHA_DOWN_PF
#define PICHA_DOWN_PF 0
#endif
constexpr int PF = PICHA_DOWN_PF;  // stages prefetched into L2 beyond the ring (0: none)
constexpr int STAGE_BYTES = PICHA_DOWN_RS * 1024;  // 8 rows of 1024 bytes (u8) or 4 rows of 2048 bytes (u16)
constexpr int kVExp = 120;         // vertical weights are scaled by 2^kVExp
// Table row of a source row: its DEPTH vertical weights in slot order, then one word of event flags (an even
// number of words: the row is fetched with 8-byte uniform loads).
__host__ __device__ constexpr int weight_stride(int depth) { return (depth + 2) & ~1; }
constexpr uint32_t kEvCount = 7;   // outputs completed by this row
constexpr uint32_t kEvStage = 8;   // the row after this one is the last of its ring stage: hand the stage back, wait for the next
constexpr uint32_t kEvMask = kEvCount | kEvStage;
constexpr int kMaxDepth = 8;

__host__ __device__ constexpr int stage_rows(bool deep) { return deep ? PICHA_DOWN_RS / 2 : PICHA_DOWN_RS; }

// Odd channel counts: float4 chunks of an expanded weight row, nb blocks of `channels` chunks (a block is
// 4 * channels floats there; the host sizes nb so that the longest window plus 3 floats of misalignment fits).
__host__ __device__ constexpr int flat_chunks(int channels, int nb) { return channels * nb; }

// slots of the column-wise horizontal pass (pass2_cols): two quarter warps more than the tile has columns
__host__ __device__ constexpr int slot_count(int tile_w) { return 8 * ((tile_w + 7) / 8 + 2); }

// whether a kernel with `group` rows per horizontal pass runs it by columns (launch_one)
__host__ __device__ constexpr bool by_columns(int channels, int group) {
	// (2-channel pixels stay with the general pass: their windows start on 8-byte boundaries, the column-wise pass
	// would read them with 8-byte loads, and measured that is slower -- greya 3840x2160 -> 1000x562: 1.02 ms against 0.62)
	return group == 4 && channels != 2 && (PICHA_DOWN_COLS >= 2 || ((channels & 1) && PICHA_DOWN_COLS));
}

struct SmemLayout {
	int ring, tmp, tmp_floats, out, out_stride, xw, xs2, xf, xslots, bars, total;
};

// nb: blocks of the horizontal pass (DownArgs::nb); wrows: weight rows held in shared memory (the plan's distinct
// rows, or one per column of the tile); direct: pixels go straight to global memory (no output tile).
// cols: the horizontal pass runs by columns (pass2_cols): even channel counts then keep one float per tap instead of a pair
__host__ __device__ inline SmemLayout smem_layout(int G, int tile_w, int bpp, int channels, int nb, int wrows, bool direct, int nt = NT, bool cols = false) {
	SmemLayout L;
	L.ring = 0;
	L.tmp = L.ring + NS * stage_bytes(nt);
	// Every column runs the same nb blocks of taps; a column with a shorter window (image edges) reads on
	// behind it with zero weights -- into the next row of the group or, from the last row, into this
	// zeroed tail (at most 4 * nb pixels).
	L.tmp_floats = G * tmps(channels, G, nt) + (4 * nb * channels + 16 + 63) / 64 * 64;
	L.out = L.tmp + L.tmp_floats * 4;
	L.out_stride = ((tile_w * bpp + 127) / 128) * 128 + 16;
	L.xw = L.out + (direct ? 0 : G * L.out_stride);
	// floats per weight row, an odd number of float4s (rows spread over the banks): duplicated weights
	// for even channel counts, the expanded flat form (see PixelAcc<3>) for odd ones
	L.xs2 = 4 * (((channels & 1) ? flat_chunks(channels, nb) : cols ? nb : 2 * nb) | 1);
	L.xf = L.xw + wrows * L.xs2 * 4;         // per column: {byte offset of the first tap in a row, of the weight row}
	// the slots of the column-wise horizontal pass (pass2_cols)
	L.xslots = L.xf + tile_w * 8;
	L.bars = L.xslots + (cols ? slot_count(tile_w) * 8 : 0);
	L.total = L.bars + 2 * NS * 8;
	return L;
}

// ---- integer-ratio horizontal pass (4-channel pixels) ------------------------------------------
// A thread produces 4 neighbouring output pixels of one row from one window of source pixels (a 16-byte unit
// each): with rq source pixels per output the windows of neighbours overlap, so each unit is loaded once and
// feeds up to 4 outputs.  The intermediate row is stored with one unit of padding after every 4 * rq units,
// counted from the first block's nominal window start: the 8 lanes of a shared-memory phase then work on 8
// neighbouring blocks of the same row, (4 rq + 1) units apart -- an odd distance, so never a bank conflict --
// and every offset inside a block's window is an immediate.
constexpr int kIntU = 4;                          // outputs per thread
constexpr int kIntGuard = 128;                    // zeroed bytes in front of the first row (edge windows start before it)
constexpr int kIntRowBytes = 16 * (256 + 32 + 4); // units of a row + padding units (rq >= 2) + slack
constexpr int kIntTabs = 6;                       // weight tables: regular blocks, 2 left-edge blocks, 3 right-edge blocks
__host__ __device__ constexpr int int_tab_bytes(int nw) { return kIntU * ((nw + 3) & ~3) * 4 + 16; }

struct SmemLayoutInt {
	int ring, tmp, wtab, bars, total;
};
__host__ __device__ inline SmemLayoutInt smem_layout_int(int nw) {
	SmemLayoutInt L;
	L.ring = 0;
	L.tmp = L.ring + NS * STAGE_BYTES;
	L.wtab = L.tmp + kIntGuard + 4 * kIntRowBytes + 1024;   // 1 KB zeroed tail: the last block's window may end behind the row
	L.bars = (L.wtab + kIntTabs * int_tab_bytes(nw) + 7) & ~7;
	L.total = L.bars + 2 * NS * 8;
	return L;
}

struct DownArgs {
	float xscale;   // factor on the horizontal weights: 2^(149 - kVExp) / max
	int nb;         // blocks every column's horizontal pass runs: 4 taps each (even channel counts), 4 * C floats (odd)
	int direct;     // destination aligned to the pixel's store unit: pixels are stored from registers (no output tile)
	int wrows;      // weight rows in shared memory: FastTables::xunique (shared by all columns) or tile_w (one each)
	int uniq;       // which of the two
	// integer-ratio horizontal pass (pass2_int4; kernels instantiated with P2 = 1)
	int rq, dx;     // source pixels per output pixel; taps per output <= rq * dx
	int off0;       // nominal first tap of column x is rq * x + off0
	int nl, br0;    // blocks of 4 columns below nl and from br0 on are irregular (image edges): own weight tables
	FuseArgs fuse;  // resize, then convert: the pack stage stores the destination's pixel format
};

__device__ __forceinline__ void mbar_init_a(uint32_t bar, int count) {
	asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive_a(uint32_t bar) {
	asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

typedef unsigned long long u64;
__device__ __forceinline__ void ffma2(u64 &acc, u64 a, u64 b) {
	asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b));
}
__device__ __forceinline__ void lds_2x64(uint32_t addr, u64 &lo, u64 &hi) {
	asm volatile("ld.shared.v2.u64 {%0, %1}, [%2];" : "=l"(lo), "=l"(hi) : "r"(addr) : "memory");
}
__device__ __forceinline__ u64 lds_64(uint32_t addr) {
	u64 v;
	asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(addr) : "memory");
	return v;
}
__device__ __forceinline__ u64 pair(float lo, float hi) {
	u64 v;
	asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi));
	return v;
}
__device__ __forceinline__ void unpair(u64 v, float &lo, float &hi) {
	asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}

// Stage transition of the ring (once per 8 KB of source).  A warp that has read the last row of a
// stage counts itself out on the stage's counter; the last warp to do so refills the slot with the
// stage NS ahead -- nobody ever waits for another warp -- and then everyone waits for the next
// stage to land (normally it has).  No loop runs under a per-thread condition here: that would make
// the compiler treat the caller's loop state as divergent and move the vertical weights out of the
// uniform registers.
struct RingState {
	int stage, slot, nstages;
	uint32_t parity;
};

template <bool DEEP, int NTT>
__device__ __noinline__ RingState ring_advance(const CUtensorMap *map, uint32_t ring, uint32_t bars, RingState rs, int word0,
                                                  int row0, int img, int tid) {
	constexpr int RSK = stage_rows(DEEP);
	constexpr int BOXES = row_boxes(DEEP, NTT);  // TMA boxes are at most 256 elements wide
	constexpr int BOXB = box_bytes(DEEP, NTT);
	constexpr int STAGE_BYTES = stage_bytes(NTT);
	const int prev = rs.slot;
#ifdef PICHA_DOWN_NO_REFILL      // (timing experiments only: the ring is filled once and re-read; no copies, no waits)
	if (rs.stage >= NS - 1) {
		++rs.stage;
		if (++rs.slot == NS) rs.slot = 0;
		return rs;
	}
#endif
	// Everything below is straight-line code for the whole warp (lane 0 acts through predicates, no divergent
	// region), and the two barrier operations whose results take long -- the arrival on the hand-back barrier and
	// the first test of the next stage's full barrier -- are both issued before either result is looked at.
	const bool hand_back = rs.stage >= 0 && rs.stage + NS < rs.nstages;
	const uint32_t lane0 = (tid & 31) == 0;
	uint32_t pending = 0;
	__syncwarp();
	if (hand_back) {
		// arrive on the stage's hand-back barrier; the state it returns holds the pending count before
		// this arrival: 1 means every other warp has been here already
		asm volatile(
			"{\n\t.reg .pred p;\n\t.reg .b64 st;\n\t"
			"setp.ne.u32 p, %2, 0;\n\t"
			"@p mbarrier.arrive.shared::cta.b64 st, [%1];\n\t"
			"@p mbarrier.pending_count.b64 %0, st;\n\t}"
			: "+r"(pending) : "r"(bars + 8 * (NS + prev)), "r"(lane0) : "memory");
	}
	const int issue = rs.stage + NS;
	++rs.stage;
	if (++rs.slot == NS) { rs.slot = 0; rs.parity ^= 1; }
	uint32_t ok;
	asm volatile(
		"{\n\t.reg .pred p;\n\t"
		"mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
		"selp.u32 %0, 1, 0, p;\n\t}"
		: "=r"(ok) : "r"(bars + 8 * rs.slot), "r"(rs.parity) : "memory");
	if (hand_back) {
		// the last warp to arrive refills the slot with the stage NS ahead
		asm volatile(
			"{\n\t.reg .pred q;\n\t"
			"setp.eq.u32 q, %0, 1;\n\t"
			"@q mbarrier.arrive.expect_tx.shared::cta.b64 _, [%2], %7;\n\t"
			"@q cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%1], [%3, {%4, %5, %6}], [%2];\n\t}"
			::"r"(pending), "r"(ring + prev * STAGE_BYTES), "r"(bars + 8 * prev), "l"(map),
			  "r"(word0), "r"(row0 + issue * RSK), "r"(img), "r"(STAGE_BYTES) : "memory");
		if (BOXES == 2)
			asm volatile(
				"{\n\t.reg .pred q;\n\t"
				"setp.eq.u32 q, %0, 1;\n\t"
				"@q cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%1], [%3, {%4, %5, %6}], [%2];\n\t}"
				::"r"(pending), "r"(ring + prev * STAGE_BYTES + RSK * BOXB), "r"(bars + 8 * prev), "l"(map),
				  "r"(word0 + BOXB / 4), "r"(row0 + issue * RSK), "r"(img) : "memory");
	}
	if (!ok) fast::mbar_wait_a(bars + 8 * rs.slot, rs.parity);
	return rs;
}

// ---- pass 2 -----------------------------------------------------------------------------------
struct Pass2Args {
	uint32_t sbase;
	int tmp, xw, xf, outt;
	uint8_t *gbase;        // destination of the group's first row, at the tile's first column
	int xs2, out_stride, dstride, tw, ng, tid, direct, nb;
	int nslots;            // pass2_cols
	FuseArgs fuse;
};

// shared-memory output tile -> global memory, 16 bytes per thread where the destination allows it
template <int BPP, int NTT> __device__ __forceinline__ void copy_out(const Pass2Args &a) {
	__syncthreads();
	const int row_bytes = a.tw * BPP;
	const bool vec = ((reinterpret_cast<uintptr_t>(a.gbase) | (uintptr_t)a.dstride) & 15) == 0;
	const int nvec = vec ? row_bytes >> 4 : 0;
	const int done = nvec << 4;
	for (int i = a.tid; i < a.ng * nvec; i += NTT) {
		const int g = i / nvec, j = i - g * nvec;
		reinterpret_cast<uint4 *>(a.gbase + (long long)g * a.dstride)[j] = lds<uint4>(a.sbase + a.outt + g * a.out_stride + 16 * j);
	}
	const int tail = row_bytes - done;
	for (int i = a.tid; i < a.ng * tail; i += NTT) {
		const int g = i / tail, j = done + (i - g * tail);
		a.gbase[(long long)g * a.dstride + j] = smem[a.outt + g * a.out_stride + j];
	}
}

// Accumulators of one output pixel and one block of 4 taps into them.  Weights come duplicated
// ({w, w} pairs) so that two channels share a packed FMA; odd channel counts keep two partial sums
// per channel (even / odd taps) to shorten the dependency chains.
template <int C> struct PixelAcc;
template <> struct PixelAcc<4> {
	u64 a01 = 0, a23 = 0;
	__device__ __forceinline__ void block(uint32_t w, uint32_t v) {
		u64 w0, w1, w2, w3, p0, p1, p2, p3, p4, p5, p6, p7;
		lds_2x64(w, w0, w1);
		lds_2x64(w + 16, w2, w3);
		lds_2x64(v, p0, p1);
		lds_2x64(v + 16, p2, p3);
		lds_2x64(v + 32, p4, p5);
		lds_2x64(v + 48, p6, p7);
		ffma2(a01, p0, w0); ffma2(a23, p1, w0);
		ffma2(a01, p2, w1); ffma2(a23, p3, w1);
		ffma2(a01, p4, w2); ffma2(a23, p5, w2);
		ffma2(a01, p6, w3); ffma2(a23, p7, w3);
	}
	// by columns (pass2_cols): the block's 4 weights come as plain floats, loaded once for all rows of the group, and
	// enter the packed FMAs as broadcast scalars (FFMA2 R, R.F32x2, R.F32, R)
	static constexpr int kW = 2, kWBytes = 16;
	__device__ __forceinline__ static void weights(uint32_t w, u64 (&q)[kW]) { lds_2x64(w, q[0], q[1]); }
	__device__ __forceinline__ void mac(const u64 (&q)[kW], uint32_t v) {
		u64 p0, p1, p2, p3, p4, p5, p6, p7;
		float w0, w1, w2, w3;
		unpair(q[0], w0, w1);
		unpair(q[1], w2, w3);
		lds_2x64(v, p0, p1);
		lds_2x64(v + 16, p2, p3);
		lds_2x64(v + 32, p4, p5);
		lds_2x64(v + 48, p6, p7);
		ffma2(a01, p0, pair(w0, w0)); ffma2(a23, p1, pair(w0, w0));
		ffma2(a01, p2, pair(w1, w1)); ffma2(a23, p3, pair(w1, w1));
		ffma2(a01, p4, pair(w2, w2)); ffma2(a23, p5, pair(w2, w2));
		ffma2(a01, p6, pair(w3, w3)); ffma2(a23, p7, pair(w3, w3));
	}
	__device__ __forceinline__ void result(float *f, int) const { unpair(a01, f[0], f[1]); unpair(a23, f[2], f[3]); }
};
template <> struct PixelAcc<2> {
	u64 e = 0, od = 0;
	__device__ __forceinline__ void block(uint32_t w, uint32_t v) {
		u64 w0, w1, w2, w3;
		lds_2x64(w, w0, w1);
		lds_2x64(w + 16, w2, w3);
		const u64 p0 = lds_64(v), p1 = lds_64(v + 8), p2 = lds_64(v + 16), p3 = lds_64(v + 24);
		ffma2(e, p0, w0); ffma2(od, p1, w1);
		ffma2(e, p2, w2); ffma2(od, p3, w3);
	}
	static constexpr int kW = 2, kWBytes = 16;
	__device__ __forceinline__ static void weights(uint32_t w, u64 (&q)[kW]) { lds_2x64(w, q[0], q[1]); }
	__device__ __forceinline__ void mac(const u64 (&q)[kW], uint32_t v) {
		float w0, w1, w2, w3;
		unpair(q[0], w0, w1);
		unpair(q[1], w2, w3);
		const u64 p0 = lds_64(v), p1 = lds_64(v + 8), p2 = lds_64(v + 16), p3 = lds_64(v + 24);
		ffma2(e, p0, pair(w0, w0)); ffma2(od, p1, pair(w1, w1));
		ffma2(e, p2, pair(w2, w2)); ffma2(od, p3, pair(w3, w3));
	}
	__device__ __forceinline__ void result(float *f, int) const {
		float e0, e1, o0, o1;
		unpair(e, e0, e1);
		unpair(od, o0, o1);
		f[0] = e0 + o0;
		f[1] = e1 + o1;
	}
};
// Odd channel counts read the row as a flat float array in aligned float4 chunks, against a weight
// row expanded to one weight per float (each tap `C` times) and shifted by the window's misalignment
// `off`.  Float j of the window belongs to channel (j - off) mod C.  For C = 3 a block is three
// chunks (12 floats, a multiple of 3): the pairs (j, j+1) of a block always fall on the residue pairs
// (0,1) (2,0) (1,2) (0,1) (2,0) (1,2), so three packed accumulators suffice.
template <> struct PixelAcc<3> {
	u64 p01 = 0, p20 = 0, p12 = 0;
	__device__ __forceinline__ void block(uint32_t w, uint32_t v) {
		u64 w0, w1, w2, w3, w4, w5, f0, f1, f2, f3, f4, f5;
		lds_2x64(w, w0, w1);
		lds_2x64(w + 16, w2, w3);
		lds_2x64(w + 32, w4, w5);
		lds_2x64(v, f0, f1);
		lds_2x64(v + 16, f2, f3);
		lds_2x64(v + 32, f4, f5);
		ffma2(p01, f0, w0); ffma2(p20, f1, w1); ffma2(p12, f2, w2);
		ffma2(p01, f3, w3); ffma2(p20, f4, w4); ffma2(p12, f5, w5);
	}
	// the same with the block's weights already in registers (pass2_cols: one column, several rows)
	static constexpr int kW = 6, kWBytes = 48;
	__device__ __forceinline__ static void weights(uint32_t w, u64 (&q)[kW]) {
		lds_2x64(w, q[0], q[1]);
		lds_2x64(w + 16, q[2], q[3]);
		lds_2x64(w + 32, q[4], q[5]);
	}
	__device__ __forceinline__ void mac(const u64 (&q)[kW], uint32_t v) {
		u64 f0, f1, f2, f3, f4, f5;
		lds_2x64(v, f0, f1);
		lds_2x64(v + 16, f2, f3);
		lds_2x64(v + 32, f4, f5);
		ffma2(p01, f0, q[0]); ffma2(p20, f1, q[1]); ffma2(p12, f2, q[2]);
		ffma2(p01, f3, q[3]); ffma2(p20, f4, q[4]); ffma2(p12, f5, q[5]);
	}
	__device__ __forceinline__ void result(float *f, int off) const {
		float a, b, c, d, e, g;
		unpair(p01, a, b);
		unpair(p20, c, d);
		unpair(p12, e, g);
		const float r0 = a + d, r1 = b + e, r2 = c + g;   // sums over the floats with j mod 3 = 0, 1, 2
		const int o = off == 3 ? 0 : off;                 // channel c sits at residue (c + off) mod 3
		f[0] = o == 0 ? r0 : o == 1 ? r1 : r2;
		f[1] = o == 0 ? r1 : o == 1 ? r2 : r0;
		f[2] = o == 0 ? r2 : o == 1 ? r0 : r1;
	}
};
template <> struct PixelAcc<1> {
	u64 p = 0, q = 0;
	__device__ __forceinline__ void block(uint32_t w, uint32_t v) {
		u64 w0, w1, f0, f1;
		lds_2x64(w, w0, w1);
		lds_2x64(v, f0, f1);
		ffma2(p, f0, w0);
		ffma2(q, f1, w1);
	}
	static constexpr int kW = 2, kWBytes = 16;
	__device__ __forceinline__ static void weights(uint32_t w, u64 (&wq)[kW]) { lds_2x64(w, wq[0], wq[1]); }
	__device__ __forceinline__ void mac(const u64 (&wq)[kW], uint32_t v) {
		u64 f0, f1;
		lds_2x64(v, f0, f1);
		ffma2(p, f0, wq[0]);
		ffma2(q, f1, wq[1]);
	}
	__device__ __forceinline__ void result(float *f, int) const {
		float a, b, c, d;
		unpair(p, a, b);
		unpair(q, c, d);
		f[0] = (a + b) + (c + d);
	}
};

// One finished pixel (row g of the group, column xx of the tile): packed and stored -- straight to global memory, or into
// the shared output tile where the destination is not aligned to the pixel's store unit.
template <int C, bool DEEP, bool FUSED>
__device__ __forceinline__ void put_pixel(const Pass2Args &a, const float (&f)[C], int g, int xx) {
	constexpr int BPP = C * Depth<DEEP>::bytes;
	if (a.direct) {
		uint32_t pv[C];
#pragma unroll
		for (int ch = 0; ch < C; ++ch) pv[ch] = fast::pack_biased<DEEP>(f[ch]);
		uint8_t *gp = a.gbase + (long long)g * a.dstride + xx * (FUSED ? pixel_bytes(a.fuse.dst_pixel) : BPP);
		if (FUSED) {
			convert_store<C, DEEP>(gp, pv, a.fuse);
		} else if (BPP == 4 && !DEEP) {
			const uint32_t lo = __byte_perm(pv[0], pv[1 % C], 0x0040), hi = __byte_perm(pv[2 % C], pv[3 % C], 0x0040);
			*reinterpret_cast<uint32_t *>(gp) = __byte_perm(lo, hi, 0x5410);
		} else if (BPP == 4) {
			*reinterpret_cast<uint32_t *>(gp) = __byte_perm(pv[0], pv[1 % C], 0x5410);
		} else if (BPP == 8) {
			*reinterpret_cast<uint2 *>(gp) = make_uint2(__byte_perm(pv[0], pv[1 % C], 0x5410), __byte_perm(pv[2 % C], pv[3 % C], 0x5410));
		} else if (BPP == 2 && !DEEP) {
			*reinterpret_cast<uint16_t *>(gp) = (uint16_t)__byte_perm(pv[0], pv[1 % C], 0x0040);
		} else {
			// 1, 3 and 6 byte pixels: channel by channel (the output of a downscale is a small part of the traffic)
#pragma unroll
			for (int ch = 0; ch < C; ++ch) {
				if (DEEP) reinterpret_cast<uint16_t *>(gp)[ch] = (uint16_t)pv[ch];
				else gp[ch] = (uint8_t)pv[ch];
			}
		}
	} else {
		fast::store_pixel<C, DEEP>(a.sbase + a.outt + g * a.out_stride + xx * BPP, f);
	}
}

// A thread produces U output pixels at a time (same row of the group, columns NT / GR apart): their
// loads and FMA chains interleave, which is what hides the shared-memory and FMA latencies here --
// there are only two or three warps per scheduler.  Every column runs the same `nb` blocks of taps
// (weights are zero-padded; what lies behind a short window is finite: see the kernel).
// (One instantiation per kernel: with several widths of U linked into the same kernel the row loop's
// register allocation suffers -- measured, cfg3 +9 %.)
#ifndef PICHA_DOWN_P2_INLINE
#define PICHA_DOWN_P2_INLINE __noinline__
#endif
// FUSED: resize, then convert -- a function of its own, so that the conversion's registers (and the spills they
// cause at the kernel's register budget) stay out of the plain resize.
// U (pixel, row) items per thread, NTT items apart, their loads and FMA chains interleaved
template <int C, bool DEEP, int GR, bool FUSED, int NTT, int U>
__device__ __forceinline__ void pass2_items(const Pass2Args &a, int o0, int total, int g, uint32_t vrow) {
	constexpr int BPP = C * Depth<DEEP>::bytes;
	constexpr int GSH = GR == 8 ? 3 : 2;
	{
		int xx[U], off[U];
		bool live[U];
		uint32_t w[U], v[U];
		PixelAcc<C> acc[U];
#pragma unroll
		for (int u = 0; u < U; ++u) {
			const int o = o0 + u * NTT;
			live[u] = o < total && g < a.ng;
			xx[u] = live[u] ? o >> GSH : 0;          // idle slots recompute column 0 and store nothing
			const uint2 e = lds<uint2>(a.sbase + a.xf + 8 * xx[u]);
			v[u] = vrow + e.x;
			w[u] = a.sbase + a.xw + (e.y & ~3u);
			off[u] = e.y & 3;
		}
		constexpr int WSTEP = (C & 1) ? 16 * C : 32, VSTEP = 16 * C;   // bytes per block: 4 taps (even C), 4 * C floats (odd C)
		const int blocks = a.nb;
		for (int kb = 0; kb < blocks; ++kb) {
#pragma unroll
			for (int u = 0; u < U; ++u) {
				acc[u].block(w[u], v[u]);
				w[u] += WSTEP;
				v[u] += VSTEP;
			}
		}
#pragma unroll
		for (int u = 0; u < U; ++u) {
			if (!live[u]) continue;
			float f[C];
			acc[u].result(f, off[u]);
			put_pixel<C, DEEP, FUSED>(a, f, g, xx[u]);
		}
	}
}

template <int C, bool DEEP, int GR, bool FUSED, int NTT>
__device__ PICHA_DOWN_P2_INLINE void pass2(Pass2Args a) {
	constexpr int BPP = C * Depth<DEEP>::bytes;
	// pixels in flight per thread: 4 for even channel counts; 2 for odd ones, whose blocks hold twice as many loaded
	// values (with 4 the kernel needs 236 registers, or spills at the 168 that 6 CTAs per SM allow)
#ifndef PICHA_DOWN_P2_U
#define PICHA_DOWN_P2_U ((C & 1) ? 2 : 4)
#endif
	constexpr int U = PICHA_DOWN_P2_U;
	const int total = a.tw * GR;
	const int g = a.tid & (GR - 1);                // NTT is a multiple of GR: the same row for every item
	const uint32_t vrow = a.sbase + a.tmp + 4 * g * tmps(C, GR, NTT);
	// whole rounds of U items per thread, then the rest with as few slots as it needs (a round of idle slots costs
	// what a full one does: with 160 items on 64 threads that was two rounds of 128 for one and a quarter)
	const int full = total / (U * NTT) * (U * NTT);
	int o0 = a.tid;
	for (; o0 < full; o0 += U * NTT) pass2_items<C, DEEP, GR, FUSED, NTT, U>(a, o0, total, g, vrow);
	const int rest = (total - full + NTT - 1) / NTT;   // slots per thread the remaining items need: 0 .. U - 1
	if (rest > 2) pass2_items<C, DEEP, GR, FUSED, NTT, U>(a, o0, total, g, vrow);
	else if (rest == 2) pass2_items<C, DEEP, GR, FUSED, NTT, 2>(a, o0, total, g, vrow);
	else if (rest == 1) pass2_items<C, DEEP, GR, FUSED, NTT, 1>(a, o0, total, g, vrow);
	if (!a.direct) copy_out<BPP, NTT>(a);
}

// ---- pass 2 by columns (odd channel counts) ------------------------------------------------------
// A thread owns one column of the tile and produces it for every row of the group: a block of the column's weights is
// loaded once and serves GR rows (pass2 above loads it once per pixel; for rgb at 7.5:1 the horizontal pass was bound
// by shared-memory wavefronts, half of them weights).  The lanes of a warp are columns here, so what decides bank
// conflicts is where the columns' windows start: the kernel's prologue deals the columns to slots such that the eight
// lanes of a quarter warp start in eight different 16-byte bank groups (see assign_slots); slot entries are
// {first byte of the window in a row | column << 16, byte offset of the weight row | misalignment}.
constexpr uint32_t kNoColumn = 0xFFFFu;

template <int C, bool DEEP, int GR, bool FUSED, int NTT>
__device__ __noinline__ void pass2_cols(Pass2Args a) {
	constexpr int BSTEP = 16 * C;                  // bytes per block: 4 * C floats of the row (odd C: as many weights; even C: 4)
	constexpr uint32_t ROWB = tmps(C, GR, NTT) * 4;
	for (int sl = a.tid; sl < a.nslots; sl += NTT) {
		const uint2 e = lds<uint2>(a.sbase + a.xf + 8 * sl);
		const uint32_t col = e.x >> 16;            // (an empty slot computes column 0's window and stores nothing)
		uint32_t v = a.sbase + a.tmp + (e.x & 0xFFFFu);
		uint32_t w = a.sbase + a.xw + (e.y & ~3u);
		const int off = e.y & 3;
		PixelAcc<C> acc[GR];
		for (int kb = 0; kb < a.nb; ++kb) {
			u64 wq[PixelAcc<C>::kW];
			PixelAcc<C>::weights(w, wq);
#pragma unroll
			for (int g = 0; g < GR; ++g) acc[g].mac(wq, v + g * ROWB);
			w += PixelAcc<C>::kWBytes;
			v += BSTEP;
		}
		if (col == kNoColumn) continue;
		if constexpr (FUSED) {
			// resize, then convert: the column's pixels of the group in one call
			PixelValues<C, DEEP, GR> pv;
#pragma unroll
			for (int g = 0; g < GR; ++g) {
				float f[C];
				acc[g].result(f, off);
#pragma unroll
				for (int ch = 0; ch < 4; ++ch) pv.v[g][ch] = ch < C ? fast::pack_biased<DEEP>(f[ch % C]) : 0u;
			}
			convert_store_n_call<C, DEEP, GR>(a.gbase + (long long)col * pixel_bytes(a.fuse.dst_pixel), a.dstride, a.ng, pv, a.fuse);
		} else {
#pragma unroll
			for (int g = 0; g < GR; ++g) {
				if (g >= a.ng) break;
				float f[C];
				acc[g].result(f, off);
				put_pixel<C, DEEP, FUSED>(a, f, g, (int)col);
			}
		}
	}
	if (!a.direct) copy_out<C * Depth<DEEP>::bytes, NTT>(a);
}

// Deals the tile's columns to the slots of pass2_cols (run by one warp): the q-th column whose window starts in bank
// group r (16-byte units modulo 8) goes to slot 8 q + r -- a quarter warp then reads eight different bank groups at
// every step of the pass.  Residues that hold more than their share of
// columns (a ratio that puts every window in the same group) overflow into whatever slots stay empty, conflicts and all.
// `first` and `wrow` are what the general pass keeps per column; the slots must be initialised empty before.
__device__ __forceinline__ void assign_slots(uint32_t slots, int nslots, uint32_t colinfo, int tw, int lane) {
	// one warp, 32 columns at a time: a column's rank among the columns of its residue comes from 8 ballots
	int base[8];
#pragma unroll
	for (int r = 0; r < 8; ++r) base[r] = 0;
	uint32_t spilled = 0;
	for (int x0 = 0; x0 < tw; x0 += 32) {
		const int x = x0 + lane;
		const bool valid = x < tw;
		const uint2 e = lds<uint2>(colinfo + 8 * (valid ? x : 0));
		const uint32_t res = valid ? (e.x >> 4) & 7 : 8;
		int q = 0;
#pragma unroll
		for (int r = 0; r < 8; ++r) {
			const uint32_t m = __ballot_sync(0xFFFFFFFFu, res == (uint32_t)r);
			if (res == (uint32_t)r) q = base[r] + __popc(m & ((1u << lane) - 1u));
			base[r] += __popc(m);
		}
		const int slot = 8 * q + (int)res;
		const bool fits = valid && slot < nslots;
		if (fits) asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(slots + 8 * slot), "r"(e.x | ((uint32_t)x << 16)), "r"(e.y) : "memory");
		spilled |= __ballot_sync(0xFFFFFFFFu, valid && !fits);
	}
	__syncwarp();
	if (spilled && lane == 0) {
		// (rare: a residue holds more columns than the slots have rows for it: walk the columns again, counting per
		// residue, and drop the ones past their residue's share into whatever slots stayed empty)
		int hole = 0;
		for (int r = 0; r < 8; ++r) {
			int q = 0;
			for (int x = 0; x < tw; ++x) {
				const uint2 e = lds<uint2>(colinfo + 8 * x);
				if (((e.x >> 4) & 7) != (uint32_t)r) continue;
				if (8 * q + r < nslots) { ++q; continue; }
				while ((lds<uint2>(slots + 8 * hole).x >> 16) != kNoColumn) ++hole;
				asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(slots + 8 * hole), "r"(e.x | ((uint32_t)x << 16)), "r"(e.y) : "memory");
			}
		}
	}
}

// ---- pass 2 for integer ratios, 4-channel pixels (see smem_layout_int) ---------------------------
struct Pass2IntArgs {
	uint32_t row0;       // shared address of unit 0 of the group's first intermediate row
	uint32_t wtab;       // shared address of the weight tables
	uint8_t *gbase;      // destination of the group's first row, at the tile's first column
	int dstride, tw, ng, tid;
	int c0;              // nominal first unit of the tile's first block, relative to the tile origin (may be negative)
	int blk0;            // the tile's first block in the image (x0 / 4)
	int nl, br0;         // DownArgs
	int vec;             // destination rows are 16-byte aligned: a block leaves as 16-byte stores
	FuseArgs fuse;
};

// RQ source pixels per output pixel, at most RQ * DX taps per output.  Lanes: 8 neighbouring blocks x 4 rows.
template <bool DEEP, int RQ, int DX, bool FUSED>
__device__ __forceinline__ void pass2_int4(const Pass2IntArgs &a) {
	constexpr int U = kIntU, NW = RQ * DX, NWP = (NW + 3) & ~3, PER = U * RQ, NK = (U - 1) * RQ + NW;
	constexpr int BPP = 4 * Depth<DEEP>::bytes;
	const int nblk = (a.tw + U - 1) / U;
	const int lane = a.tid & 31, g = lane >> 3;
	for (int i = (a.tid >> 5) * 8 + (lane & 7); i < nblk; i += (NT / 32) * 8) {
		if (g >= a.ng) continue;
		const int b = a.blk0 + i;
		const int slot = b < a.nl ? 1 + b : b >= a.br0 ? 3 + min(b - a.br0, 2) : 0;
		const uint32_t wt = a.wtab + slot * int_tab_bytes(NW);
		const uint32_t v = a.row0 + g * kIntRowBytes + 16 * ((PER + 1) * i + a.c0);
		u64 acc[U][2];
#pragma unroll
		for (int u = 0; u < U; ++u) acc[u][0] = acc[u][1] = 0;
		float4 wq[U];
#pragma unroll
		for (int k = 0; k < NK; ++k) {
			u64 p01, p23;
			lds_2x64(v + 16 * (k + k / PER), p01, p23);
#pragma unroll
			for (int u = 0; u < U; ++u) {
				const int wi = k - u * RQ;       // tap of output u this unit is (compile-time)
				if (wi < 0 || wi >= NW) continue;
				if ((wi & 3) == 0) wq[u] = lds<float4>(wt + (u * NWP + wi) * 4);
				const float w = (wi & 3) == 0 ? wq[u].x : (wi & 3) == 1 ? wq[u].y : (wi & 3) == 2 ? wq[u].z : wq[u].w;
				ffma2(acc[u][0], p01, pair(w, w));
				ffma2(acc[u][1], p23, pair(w, w));
			}
		}
		uint32_t px[U][DEEP ? 2 : 1];
		PixelValues<4, DEEP, U> fpv;
#pragma unroll
		for (int u = 0; u < U; ++u) {
			float f[4];
			unpair(acc[u][0], f[0], f[1]);
			unpair(acc[u][1], f[2], f[3]);
			uint32_t pv[4];
#pragma unroll
			for (int ch = 0; ch < 4; ++ch) pv[ch] = fast::pack_biased<DEEP>(f[ch]);
			if (FUSED) {                     // resize, then convert: stored below, in the destination's format
#pragma unroll
				for (int ch = 0; ch < 4; ++ch) fpv.v[u][ch] = pv[ch];
				continue;
			}
			if (DEEP) {
				px[u][0] = __byte_perm(pv[0], pv[1], 0x5410);
				px[u][DEEP ? 1 : 0] = __byte_perm(pv[2], pv[3], 0x5410);
			} else {
				px[u][0] = __byte_perm(__byte_perm(pv[0], pv[1], 0x0040), __byte_perm(pv[2], pv[3], 0x0040), 0x5410);
			}
		}
		if (FUSED) {
			const int pb = pixel_bytes(a.fuse.dst_pixel);
			convert_store_n_call<4, DEEP, U>(a.gbase + (long long)g * a.dstride + (long long)(U * i) * pb, pb, a.tw - U * i, fpv, a.fuse);
			continue;
		}
		uint8_t *gp = a.gbase + (long long)g * a.dstride + (long long)(U * i) * BPP;
		if (a.vec && U * i + U <= a.tw) {
			if (DEEP) {
				reinterpret_cast<uint4 *>(gp)[0] = make_uint4(px[0][0], px[0][DEEP ? 1 : 0], px[1][0], px[1][DEEP ? 1 : 0]);
				reinterpret_cast<uint4 *>(gp)[1] = make_uint4(px[2][0], px[2][DEEP ? 1 : 0], px[3][0], px[3][DEEP ? 1 : 0]);
			} else {
				*reinterpret_cast<uint4 *>(gp) = make_uint4(px[0][0], px[1][0], px[2][0], px[3][0]);
			}
		} else {
#pragma unroll
			for (int u = 0; u < U; ++u) {
				if (U * i + u >= a.tw) break;
				if (DEEP) reinterpret_cast<uint2 *>(gp)[u] = make_uint2(px[u][0], px[u][DEEP ? 1 : 0]);
				else reinterpret_cast<uint32_t *>(gp)[u] = px[u][0];
			}
		}
	}
}

// One out-of-line function holds all the variants: the row loop calls it from every emit slot, and with the switch at
// the call sites (one call per variant) each slot carried 160 instructions of dispatch -- the row loop's code is what
// fills the instruction cache.
template <bool DEEP, bool FUSED> __device__ __noinline__ void pass2_int4_any(Pass2IntArgs a, int rq, int dx) {
	switch (rq * 8 + dx) {
		case 2 * 8 + 2: pass2_int4<DEEP, 2, 2, FUSED>(a); break;
		case 2 * 8 + 4: pass2_int4<DEEP, 2, 4, FUSED>(a); break;
		case 2 * 8 + 6: pass2_int4<DEEP, 2, 6, FUSED>(a); break;
		case 3 * 8 + 2: pass2_int4<DEEP, 3, 2, FUSED>(a); break;
		case 3 * 8 + 4: pass2_int4<DEEP, 3, 4, FUSED>(a); break;
		case 3 * 8 + 6: pass2_int4<DEEP, 3, 6, FUSED>(a); break;
		case 4 * 8 + 2: pass2_int4<DEEP, 4, 2, FUSED>(a); break;
		case 4 * 8 + 4: pass2_int4<DEEP, 4, 4, FUSED>(a); break;
		default: pass2_int4<DEEP, 4, 6, FUSED>(a); break;
	}
}

// ---- the kernel -------------------------------------------------------------------------------
#ifndef PICHA_DOWN_MINB
#define PICHA_DOWN_MINB(D) ((D) <= 4 ? 6 : (D) <= 6 ? 5 : 4)
#endif

// P2: 0 = general horizontal pass (pass2), 1 = integer-ratio pass for 4-channel pixels (pass2_int4), 2 = by columns
// (pass2_cols, odd channel counts).  A template
// parameter so that each kernel links one family of callees (the registers of the row loop are what is left over).
// FUSED: resize, then convert (picha_b200_resize_convert) -- kernels of their own: a plain resize kernel that merely
// CONTAINS the call of a converting horizontal pass runs 7 % slower (measured; the callee's nested call gives the
// whole kernel a stack frame).
// NTT: threads per CTA (64; 96 or 128 for the wide 8-bit variants, see stage_bytes)
template <int DEPTH, bool DEEP, int C, int GR, int P2, bool FUSED, int NTT>
#ifndef PICHA_DOWN_MINB8
#define PICHA_DOWN_MINB8 4
#endif
__global__ void __launch_bounds__(NTT, NTT == 128 ? 3 : NTT == 96 ? 4 : GR == 8 ? PICHA_DOWN_MINB8 : PICHA_DOWN_MINB(DEPTH))   // 8-row groups: shared memory allows 4 CTAs per SM anyway
resize_down_kernel(const CUtensorMap *__restrict__ smap, DevBatch dst, FastTables t,
                   const __grid_constant__ VTable vt, DownArgs da) {
	constexpr int BPP = C * Depth<DEEP>::bytes;
	constexpr int RSK = stage_rows(DEEP);
	constexpr int WS = weight_stride(DEPTH);     // floats per table row: DEPTH weights and the row's event flags
	constexpr int WPT = DEEP ? 8 : 4;            // 32-bit words of a source row per thread
	constexpr int BOXES = row_boxes(DEEP, NTT);
	constexpr int BOXB = box_bytes(DEEP, NTT);
	constexpr int STAGE_BYTES = stage_bytes(NTT);
	static_assert(P2 != 1 || NTT == NT, "the integer-ratio pass is laid out for 64 threads");
	static_assert(P2 != 2 || GR == 4, "the column-wise pass is written for 4-row groups");
	const int tid = threadIdx.x;
	asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // see launch_one

	const int x0 = blockIdx.x * t.tile_w;
	const int tw = min(t.tile_w, dst.width - x0);
	const int sx0 = t.xfirst[x0] / t.align_px * t.align_px;   // tile origin: 16-byte aligned in the row (TMA box start)
	const int word0 = sx0 * BPP / 4;
	const int band = blockIdx.y;
	const int y0 = vt.y_begin + band * t.band_h, y1 = min(vt.y_end, y0 + t.band_h);
	const int rlo = vt.band_rlo[band], rhi = vt.band_rhi[band];

	const bool direct = da.direct != 0;
	const SmemLayout L = smem_layout(GR, t.tile_w, BPP, C, da.nb, da.wrows, direct, NTT, P2 == 2);
	const SmemLayoutInt LI = smem_layout_int(da.rq * da.dx);
	uint32_t sbase = smem_u32(smem);
	asm volatile("" : "+r"(sbase));   // keep it in a register: never re-derived
	const uint32_t bars = sbase + (P2 == 1 ? LI.bars : L.bars);   // full[NS] mbarriers, then NS hand-back mbarriers
	const uint32_t ring = sbase + L.ring;

	RingState rs;
	rs.stage = -1; rs.slot = NS - 1; rs.parity = 1;
	// rows rlo .. rhi + 2: every row is prefetched one ahead, and the row before a stage's last one already waits
	// for the next stage
	rs.nstages = (rhi + 2 - rlo) / RSK + 1;

	// the descriptor slot is reused (see map_slot in resize_fast.cu); lane 0 of any warp may issue a refill
	if ((tid & 31) == 0) asm volatile("fence.proxy.tensormap::generic.acquire.gpu [%0], 128;" ::"l"(smap) : "memory");
	if (tid == 0) {
		for (int i = 0; i < NS; ++i) {
			mbar_init_a(bars + 8 * i, 1);
			mbar_init_a(bars + 8 * (NS + i), NTT / 32);   // hand-back: one arrival per warp
		}
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
		asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
		for (int k = 0; k < NS && k < rs.nstages; ++k) {
			fast::mbar_expect_tx_a(bars + 8 * k, STAGE_BYTES);
#pragma unroll
			for (int b = 0; b < BOXES; ++b)
				fast::tma_load_3d_a(ring + k * STAGE_BYTES + b * RSK * BOXB, smap, bars + 8 * k, word0 + b * (BOXB / 4), rlo + k * RSK, blockIdx.z);
		}
	}
	// this tile's horizontal tables -> shared memory: weights scaled and duplicated (packed-FMA operands)
	const bool uniq = da.uniq != 0;                // the plan's distinct rows, else one row per column
	const int wrows = uniq ? da.wrows : tw;
	// integer-ratio pass: nominal first unit of the tile's first block relative to the tile origin
	const int c0 = da.rq * x0 + da.off0 - sx0;
	if (P2 == 1) {
		// weight tables [table][output of the block][tap], zero where a column has no tap: table 0 from a regular
		// block, 1..2 the image's first blocks, 3..5 its last ones
		const int nw = da.rq * da.dx, nwp = (nw + 3) & ~3;
		for (int i = tid; i < kIntTabs * kIntU * nwp; i += NTT) {
			const int tab = i / (kIntU * nwp), u = (i / nwp) % kIntU, j = i % nwp;
			const int blk = tab == 0 ? da.nl : tab <= 2 ? tab - 1 : da.br0 + (tab - 3);
			const int x = kIntU * blk + u;
			float w = 0.0f;
			if (x < dst.width && j < nw) {
				const int k = j - (t.xfirst[x] - (da.rq * x + da.off0));
				if (k >= 0 && k < t.xcount[x]) w = t.xw[(long long)x * t.xstride + k] * da.xscale;
			}
			sts(sbase + LI.wtab + tab * int_tab_bytes(nw) + (u * nwp + j) * 4, w);
		}
		for (int i = tid; i < (kIntGuard + 4 * kIntRowBytes + 1024) / 4; i += NTT) sts(sbase + LI.tmp + 4 * i, 0.0f);
	} else
	if (C & 1) {
		// a thread expands whole rows: zero the row, then every tap C times from its position on (one global load per
		// tap and no index arithmetic per float: with one row per column -- ratios whose weights are not periodic -- a
		// float-by-float fill with three dependent loads each took a third of a band's time)
		constexpr int ci = C == 3;
		const int nq = flat_chunks(C, da.nb);
		for (int row = tid; row < wrows; row += NTT) {
			const int src = uniq ? t.xe_src[ci][row] : t.xrow[x0 + row];
			const int off = uniq ? t.xe_off[ci][row] : ((t.xfirst[x0 + row] - sx0) * C) & 3;
			const uint32_t wr = sbase + L.xw + 4 * row * L.xs2;
			for (int q = 0; q < nq; ++q) sts(wr + 16 * q, make_float4(0.0f, 0.0f, 0.0f, 0.0f));
			const float *ws = t.xuw + (long long)src * t.xstride;
			for (int kk = 0; kk < t.xstride; ++kk) {
				const float w = ws[kk] * da.xscale;
#pragma unroll
				for (int c = 0; c < C; ++c)
					if (off + C * kk + c < 4 * nq) sts(wr + 4 * (off + C * kk + c), w);
			}
		}
	} else {
		const float *wsrc = uniq ? t.xuw : t.xw + (long long)x0 * t.xstride;
		for (int i = tid; i < wrows * da.nb * 4; i += NTT) {
			const int row = i / (da.nb * 4), k = i - row * (da.nb * 4);
			const float w = k < t.xstride ? wsrc[(long long)row * t.xstride + k] * da.xscale : 0.0f;
			if (P2 == 2) sts(sbase + L.xw + 4 * (row * L.xs2 + k), w);
			else asm volatile("st.shared.v2.f32 [%0], {%1, %1};" ::"r"(sbase + L.xw + 4 * (row * L.xs2 + 2 * k)), "f"(w) : "memory");
		}
	}
	for (int i = tid; i < (P2 == 1 ? 0 : tw); i += NTT) {
		// {byte offset of the column's first (aligned) float in a row, byte offset of its weight row | misalignment}
		const int first = (t.xfirst[x0 + i] - sx0) * C, off = (C & 1) ? first & 3 : 0;
		const int wrow = !uniq ? i : (C & 1) ? t.xe_col[C == 3][x0 + i] : t.xrow[x0 + i];
		asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sbase + L.xf + 8 * i), "r"((first - off) * 4), "r"(wrow * L.xs2 * 4 + off) : "memory");
	}
	// padded taps multiply whatever lies behind a column's window by zero: make sure that is never a NaN
	for (int i = tid; i < (P2 == 1 ? 0 : L.tmp_floats); i += NTT) sts(sbase + L.tmp + 4 * i, 0.0f);
	const int nslots = slot_count(tw);
	if (P2 == 2)
		for (int i = tid; i < nslots; i += NTT)
			asm volatile("st.shared.v2.u32 [%0], {%1, %2};" ::"r"(sbase + L.xslots + 8 * i), "r"(kNoColumn << 16), "r"(0) : "memory");
	__syncthreads();
	if (P2 == 2) {
		if (tid < 32) assign_slots(sbase + L.xslots, nslots, sbase + L.xf, tw, tid);
		__syncthreads();
	}

	// This thread's share of a staged row: four chunks of 4 values, 256 values apart (consecutive
	// lanes read consecutive words and, in the emit, write consecutive float4s: no bank conflicts).
	const uint32_t thread_off = DEEP ? 8 * tid : 4 * tid;
	uint32_t faddr = 0;     // shared address of this thread's first chunk in the next row to fetch
	// chunk q = values 4 * (tid + NTT * q) ...; rows of two boxes hold chunks 0, 1 in the first and 2, 3 in the second
	auto load_chunk = [&](uint32_t (&w)[WPT], const int q) {
		if (DEEP) {
			const uint2 v = lds<uint2>(faddr + (q >> 1) * (RSK * BOXB) + (q & 1) * (BOXB / 2));
			w[(2 * q) % WPT] = v.x; w[(2 * q + 1) % WPT] = v.y;
		} else {
			w[q % WPT] = (uint32_t)lds<int>(faddr + (BOXES == 2 ? (q >> 1) * (RSK * BOXB) + (q & 1) * (BOXB / 2) : q * (BOXB / 4)));
		}
	};
	auto load_row = [&](uint32_t (&w)[WPT]) {
#pragma unroll
		for (int q = 0; q < 4; ++q) load_chunk(w, q);
		faddr += BOXB;
	};
	auto advance = [&]() {
		rs = ring_advance<DEEP, NTT>(smap, ring, bars, rs, word0, rlo, blockIdx.z, tid);
		faddr = ring + rs.slot * STAGE_BYTES + thread_off;
	};

	// accumulators as packed pairs: fma.rn.f32x2 takes the weight as a broadcast uniform operand
	// (FFMA2 R, R.F32x2, UR.F32, R.F32x2) and does two MACs in two issue cycles -- the same FMA rate with
	// half the instructions in flight
	u64 acc[DEPTH][NV / 2];
#pragma unroll
	for (int j = 0; j < DEPTH; ++j)
#pragma unroll
		for (int i = 0; i < NV / 2; ++i) acc[j][i] = 0;

	// One source row into every open output row.  The value is used as the subnormal float its bits
	// already are (see the header); weights come from the constant bank as uniform operands.
	// fresh: the slot (or -1) whose output was emitted right before this row and whose accumulators still hold that
	// output -- its FMAs take zero as the addend instead of the accumulator, which is what clears the slot (an emit
	// that zeroes the slot itself costs 8 packed multiplies, and ptxas renames them around the stores: 10 more MOVs)
	// (a plain int: the lambda is inlined and the loop over slots unrolled, so the comparison is static)
	auto body = [&](const int FRESH, const uint32_t (&cur)[WPT], const float (&w)[DEPTH + 1]) {
#ifdef PICHA_DOWN_SKIP_BODY      // (timing experiments only: the row loop without its arithmetic)
		uint32_t x = __float_as_uint(w[0]);
#pragma unroll
		for (int i = 0; i < WPT; ++i) x ^= cur[i];
		acc[0][0] ^= x;
#else
#pragma unroll
		for (int i = 0; i < NV; i += 2) {
			uint32_t b0, b1;
			if (DEEP) { b0 = cur[(i >> 1) % WPT] & 0xFFFFu; b1 = cur[(i >> 1) % WPT] >> 16; }
			else { b0 = __byte_perm(cur[(i >> 2) % WPT], 0, 0x4440 + (i & 3)); b1 = __byte_perm(cur[(i >> 2) % WPT], 0, 0x4441 + (i & 3)); }
			const u64 uu = pair(__uint_as_float(b0), __uint_as_float(b1));
#pragma unroll
			for (int j = 0; j < DEPTH; ++j) {
				if (j == FRESH) asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc[j][i >> 1]) : "l"(uu), "l"(pair(w[j], w[j])), "l"(0ull));
				else ffma2(acc[j][i >> 1], uu, pair(w[j], w[j]));
			}
		}
#endif
	};
	// The weights of a row are fetched from the constant bank into uniform registers one row ahead,
	// like the data: issued right in front of their first use, the load's latency stalls every row.
	auto load_w = [&](float (&w)[DEPTH + 1], int widx) {
#pragma unroll
		for (int j = 0; j <= DEPTH; ++j) w[j] = vt.wdown[widx + j];  // [DEPTH]: the row's event flags
	};

	// (the horizontal pass's arguments are put together at the call, from kernel parameters: held in registers
	// across the row loop they cost it a dozen registers)
	uint8_t *const dtile = dst.base + (long long)blockIdx.z * dst.step + (long long)x0 * (FUSED ? pixel_bytes(da.fuse.dst_pixel) : BPP);
	const uint32_t my_tmp = sbase + L.tmp + tid * 16;
	// integer-ratio pass: where this thread's four units of an intermediate row go (one unit of padding after every
	// 4 * rq units, counted from c0)
	uint32_t epos[4];
	if (P2 == 1) {
		const int per = kIntU * da.rq;
#pragma unroll
		for (int q = 0; q < 4; ++q) {
			const int u = tid + NTT * q, d = u - c0;
			epos[q] = sbase + LI.tmp + kIntGuard + 16 * (u + (d >= 0 ? d / per : -1));
		}
	}

	// The row loop.  Which rows complete an output is not computed here: the host has put the number of outputs
	// a row completes (and whether the ring stage ends with the next row: bands start on stage boundaries) into
	// a flag word behind the row's weights.  The weights of a row are fetched a row ahead, so the flag has been in a uniform
	// register for a whole row body when the branch at the end of the body needs it -- all the loop decides per
	// row is "did this row complete something" and "is the ring stage used up", both on values known long before.
	// (The first version walked per-output row counts from a table: the chain table load -> min -> counters ->
	// slot comparison -> branches sat on the critical path of every output, ~300 cycles of fixed latency that
	// three warps per scheduler cannot hide; measured, the row loop ran at 160 cycles per row against 89 for the
	// row body alone.)  The output loop is unrolled DEPTH times: the slot of every emit is static.
	uint32_t ra[WPT], rb[WPT];
	float wa[DEPTH + 1], wb[DEPTH + 1];
	advance();
	load_row(ra);
	int widx = (rlo - vt.row_base) * WS;           // weights of the row held in ra, then of the one after it
	load_w(wa, widx);
	widx += WS;
	// outputs are taken in rounds of DEPTH starting at slot 0: the band's oldest open output is preceded by the
	// (up to DEPTH - 1) outputs that share its round; they count as complete before the first row and are discarded
	const int ys = vt.band_ys[band];
	int y = ys - ys % DEPTH;
	int pending = ys - y;                          // outputs complete and not yet emitted
	int gcount = 0;
	auto flags = [](const float (&w)[DEPTH + 1]) { return (int)__float_as_uint(w[DEPTH]); };
	bool done = y >= y1;
	// The pieces of the row loop (lambdas, inlined where they are used):
	// the first row after the emit of slot `fresh`: clears that slot on the way (see body; at the band's start the slot
	// is zero anyway); the row after it moves to ra whatever the flags say
	auto first_row = [&](const int fresh) {
#if PICHA_DOWN_SPLIT_LOAD
		load_chunk(rb, 0);
		load_chunk(rb, 1);
		load_w(wb, widx);
		widx += WS;
		body(fresh, ra, wa);
		load_chunk(rb, 2);
		load_chunk(rb, 3);
		faddr += BOXB;
#else
		load_row(rb);
		load_w(wb, widx);
		widx += WS;
		body(fresh, ra, wa);
#endif
		const int f = flags(wa);
#pragma unroll
		for (int i = 0; i < WPT; ++i) ra[i] = rb[i];
#pragma unroll
		for (int j = 0; j <= DEPTH; ++j) wa[j] = wb[j];
		if (f) {
			if (f & kEvStage) advance();
			pending = f & kEvCount;
		}
	};
	// the rows after it, until one completes an output
	auto rest_rows = [&]() {
#if PICHA_DOWN_ONE_BODY
		// (one body per loop: ptxas keeps the row and its successor in the same registers anyway -- each word of
		// the next row is loaded right behind the last use of the current one -- so a loop unrolled over two
		// row buffers only doubles the code, and the row loop's code is what fills the instruction cache)
		for (;;) {
#if PICHA_DOWN_SPLIT_LOAD
			// (the next row's first two chunks before the body, its last two behind it: ptxas hoists loads issued
			// before the body above the last use of the registers they will end up in, and pays two MOVs per row)
			load_chunk(rb, 0);
			load_chunk(rb, 1);
			load_w(wb, widx);
			widx += WS;
			body(-1, ra, wa);
			load_chunk(rb, 2);
			load_chunk(rb, 3);
			faddr += BOXB;
#else
			load_row(rb);
			load_w(wb, widx);
			widx += WS;
			body(-1, ra, wa);
#endif
			const int f = flags(wa);
#pragma unroll
			for (int i = 0; i < WPT; ++i) ra[i] = rb[i];
#pragma unroll
			for (int j = 0; j <= DEPTH; ++j) wa[j] = wb[j];
			if (f) {
				if (f & kEvStage) advance();
				pending = f & kEvCount;
				if (pending) break;
			}
		}
#else
		for (;;) {
			load_row(rb);
			load_w(wb, widx);
			widx += WS;
			body(-1, ra, wa);
			int f = flags(wa);
			if (f) {
				// (rare path: once per output and once per ring stage) continue with the current row in ra
#pragma unroll
				for (int i = 0; i < WPT; ++i) ra[i] = rb[i];
#pragma unroll
				for (int j = 0; j <= DEPTH; ++j) wa[j] = wb[j];
				if (f & kEvStage) advance();
				pending = f & kEvCount;
				if (pending) break;
				continue;
			}
			load_row(ra);
			load_w(wa, widx);
			widx += WS;
			body(-1, rb, wb);
			f = flags(wb);
			if (f) {
				if (f & kEvStage) advance();
				pending = f & kEvCount;
				if (pending) break;
			}
		}
#endif
	};
	// emit: slot S is final; it becomes the slot of output y + DEPTH
	auto emit = [&](const int S) {
#ifdef PICHA_DOWN_SKIP_EMIT      // (timing experiments only)
		if (y >= y0) ++gcount;
		if (y == -12345)
#endif
		{
			// (predicated stores, no branch: around a branch ptxas renames the slot's registers and moves them back)
			const uint32_t keep = y >= y0;
#pragma unroll
			for (int q = 0; q < 4; ++q) {
				const uint32_t addr = P2 == 1 ? epos[q] + gcount * kIntRowBytes : my_tmp + gcount * (tmps(C, GR, NTT) * 4) + q * (16 * NTT);
				asm volatile(
					"{\n\t.reg .pred p;\n\t"
					"setp.ne.u32 p, %3, 0;\n\t"
					"@p st.shared.v2.b64 [%0], {%1, %2};\n\t}"
					::"r"(addr), "l"(acc[S][2 * q]), "l"(acc[S][2 * q + 1]), "r"(keep) : "memory");
			}
			gcount += keep;
		}
		// The slot is cleared by the next row's body (see `fresh`); only when another output follows without a row
		// in between (the discarded outputs at a band's start, ratios below 1 row per output) is it zeroed here --
		// in place (tied operand, x * 0): a plain "= 0" makes new values that ptxas pairs up for CS2R and then
		// shuffles every accumulator of the kernel between two register assignments per output row.
		if (!PICHA_DOWN_FRESH || pending != 1) {
#pragma unroll
			for (int i = 0; i < NV / 2; ++i) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(acc[S][i]) : "l"(0ull));
		}
	};
	// a group of GR output rows is complete (or the band ends): the horizontal pass
	auto group = [&]() {
		if (gcount == GR || (done && gcount > 0)) {
			uint8_t *const gbase = dtile + (long long)(y - gcount) * dst.stride;
#ifndef PICHA_DOWN_SKIP_SYNC     // (timing experiments only)
			__syncthreads();           // the group's intermediate rows are complete
#endif
#ifndef PICHA_DOWN_SKIP_P2       // (timing experiments only: pass 1 and its events without the horizontal pass)
			if constexpr (P2 == 1) {
				Pass2IntArgs pi;
				pi.row0 = sbase + LI.tmp + kIntGuard; pi.wtab = sbase + LI.wtab; pi.dstride = dst.stride; pi.tw = tw; pi.tid = tid;
				pi.c0 = da.rq * x0 + da.off0 - sx0; pi.blk0 = x0 / kIntU; pi.nl = da.nl; pi.br0 = da.br0;
				pi.vec = ((reinterpret_cast<uintptr_t>(dtile) | (uintptr_t)dst.stride) & 15) == 0;
				pi.fuse = da.fuse;
				pi.ng = gcount;
				pi.gbase = gbase;
				pass2_int4_any<DEEP, FUSED>(pi, da.rq, da.dx);
			} else {
				Pass2Args pa;
				pa.sbase = sbase; pa.tmp = L.tmp; pa.xw = L.xw; pa.xf = L.xf; pa.outt = L.out;
				pa.xs2 = L.xs2; pa.out_stride = L.out_stride; pa.dstride = dst.stride; pa.tw = tw; pa.tid = tid; pa.direct = direct; pa.nb = da.nb;
				pa.fuse = da.fuse;
				pa.ng = gcount;
				pa.gbase = gbase;
				if constexpr (P2 == 2) {
					pa.xf = L.xslots;
					pa.nslots = nslots;
					pass2_cols<C, DEEP, GR, FUSED, NTT>(pa);
				} else {
					pass2<C, DEEP, GR, FUSED, NTT>(pa);
				}
			}
#endif
#ifndef PICHA_DOWN_SKIP_SYNC
			__syncthreads();           // pass 1 may overwrite the intermediate rows again
#endif
			gcount = 0;
		}
	};
	// (One copy of the loop body and of the group code for ALL slots, the slot-specific pieces hanging off a uniform
	// counter, halves the loop's code once more and is 9 % slower: the compare chain in front of the first row and of
	// the emit sits on every output's critical path -- profiles/r02_ablation.txt, "shared1".)
	while (!done) {
#pragma unroll
		for (int s = 0; s < DEPTH; ++s) {
			if (done) break;
			if (pending == 0) {
#if PICHA_DOWN_FRESH
				first_row((s + DEPTH - 1) % DEPTH);
				if (pending == 0)
#endif
				rest_rows();
			}
			emit(s);
			--pending;
			++y;
			done = y >= y1;
			group();
		}
	}

	// Never leave with a TMA load still in flight: wait for every stage that was issued.
#ifdef PICHA_DOWN_NO_REFILL
	rs.stage = rs.nstages;
#endif
	for (int k = rs.stage + 1; k < rs.nstages && k < rs.stage + NS; ++k) {
		if (++rs.slot == NS) { rs.slot = 0; rs.parity ^= 1; }
		fast::mbar_wait_a(bars + 8 * rs.slot, rs.parity);
	}
	asm volatile("griddepcontrol.wait;" ::: "memory");
}

struct DownLaunch {
	const CUtensorMap *map;
	const DevBatch *dst;
	const FastTables *t;
	const VTable *vt;
	DownArgs da;
	int n, channels, smem_bytes, bands;
	int overlap;    // not the first launch of this resize: may start while the previous one drains
	int group;      // output rows per pass-2 group (4 or 8)
	int threads;    // threads per CTA: 64, or 96 / 128 (wide 8-bit variants: 4-row groups, general horizontal pass, depth <= 4)
	cudaStream_t stream;
};

// The launches of one resize (one per group of row bands) write disjoint rows and read the same
// source, so nothing orders them: from the second on they are launched with programmatic stream
// serialization and start filling SMs as soon as every CTA of the previous launch has started
// (griddepcontrol.launch_dependents at the top of the kernel) -- the partial last wave of each
// launch would otherwise idle a good part of the GPU.  Every CTA ends with griddepcontrol.wait, so a
// launch never completes before its predecessor and whatever follows in the stream sees all of them.
template <int DEPTH, bool DEEP, int C, int GR, int P2, bool FUSED, int NTT = NT> cudaError_t launch_group(const DownLaunch &a) {
	auto kern = resize_down_kernel<DEPTH, DEEP, C, GR, P2, FUSED, NTT>;
	static SmemGrant granted;   // (per instantiation: see grow_dynamic_smem)
	cudaError_t e = grow_dynamic_smem(reinterpret_cast<const void *>(kern), a.smem_bytes, &granted);
	if (e != cudaSuccess) return e;
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3((a.dst->width + a.t->tile_w - 1) / a.t->tile_w, a.bands, a.n);
	cfg.blockDim = dim3(NTT);
	cfg.dynamicSmemBytes = a.smem_bytes;
	cfg.stream = a.stream;
	cudaLaunchAttribute attr[1];
	attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
	attr[0].val.programmaticStreamSerializationAllowed = 1;
	cfg.attrs = attr;
	cfg.numAttrs = a.overlap ? 1 : 0;
	return cudaLaunchKernelEx(&cfg, kern, a.map, *a.dst, *a.t, *a.vt, a.da);
}

template <int DEPTH, bool DEEP, int C> cudaError_t launch_one(const DownLaunch &a) {
	// horizontal pass of the 4-row groups: by columns (pass2_cols) or the general one, see by_columns
	constexpr int PC = by_columns(C, 4) ? 2 : 0;
	// (the wide variants exist for 8-bit formats and depths up to 4 only: elsewhere the names below are the 64-thread kernel)
	constexpr int W96 = (DEEP || DEPTH > 4) ? NT : 96, W128 = (DEEP || DEPTH > 4) ? NT : 128;
	const bool wide = !DEEP && DEPTH <= 4;
	if (a.da.fuse.dst_pixel >= 0) {       // (converting kernels: 4-row groups only)
		if (C == 4 && a.da.rq > 0) return launch_group<DEPTH, DEEP, C, 4, C == 4, true>(a);
		if (wide && a.threads == 96) return launch_group<DEPTH, DEEP, C, 4, PC, true, W96>(a);
		if (wide && a.threads == 128) return launch_group<DEPTH, DEEP, C, 4, PC, true, W128>(a);
		return launch_group<DEPTH, DEEP, C, 4, PC, true>(a);
	}
	if (C == 4 && a.da.rq > 0) return launch_group<DEPTH, DEEP, C, 4, C == 4, false>(a);   // (C == 4: no instantiation for other formats)
	if (wide && a.threads == 96) return launch_group<DEPTH, DEEP, C, 4, PC, false, W96>(a);
	if (wide && a.threads == 128) return launch_group<DEPTH, DEEP, C, 4, PC, false, W128>(a);
	return a.group == 8 ? launch_group<DEPTH, DEEP, C, 8, 0, false>(a) : launch_group<DEPTH, DEEP, C, 4, PC, false>(a);
}

template <bool DEEP, int C> cudaError_t launch_depth(const DownLaunch &a) {
	const int d = a.t->depth;
	if (d <= 3) return launch_one<3, DEEP, C>(a);
	if (d <= 4) return launch_one<4, DEEP, C>(a);
	if (d <= 5) return launch_one<5, DEEP, C>(a);
	if (d <= 6) return launch_one<6, DEEP, C>(a);
	if (d <= 8) return launch_one<8, DEEP, C>(a);
	return cudaErrorNotSupported;
}

}  // namespace down

using down::DownLaunch;

// One definition per instantiation unit (channels 1..4, 8- and 16-bit).
template <bool DEEP, int C> cudaError_t launch_down(const DownLaunch &a);

}  // namespace picha_b200
#endif
