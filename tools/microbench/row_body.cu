// Micro-benchmark of the vertical-pass row body of csrc/resize_fast.cu in isolation: per row a
// thread reads 8 packed u8 values and 8 vertical weights from shared memory, unpacks the bytes and
// does 5 x 8 FMAs.  Variants differ only in how bytes become floats.  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o row_body row_body.cu && ./row_body
#include <cstdio>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#define ITERS 4096
#define ROWS 32
// 0: PRMT + fma(m, inv, bias) with immediates     1: PRMT + (m - 2^23) (inv folded into the weights)
// 2: PRMT only (magic float used as is; timing floor)   3: PRMT + fma with register constants
// 4: I2F.U8                                        5: PRMT -> half2, HADD2, cvt  (two bytes per PRMT)
// 6: no unpack at all, 4 slots instead of 5 (32 FMAs)
__constant__ float cwts[ROWS * 8];
// WSRC 0: vertical weights read from shared memory (vector registers)
// WSRC 1: vertical weights read from constant memory with a uniform index (uniform registers)
template <int MODE, int WSRC>
__global__ void __launch_bounds__(128) k(float *out, float cscale) {
	__shared__ __align__(16) unsigned data[ROWS][256];
	__shared__ __align__(16) float wts[ROWS][8];
	for (int i = threadIdx.x; i < ROWS * 256; i += 128) (&data[0][0])[i] = i * 2654435761u;
	for (int i = threadIdx.x; i < ROWS * 8; i += 128) (&wts[0][0])[i] = (i % 8 < 5) ? 1e-3f * (i % 7 + 1) : 0.f;
	__syncthreads();
	constexpr int SLOTS = MODE == 6 ? 4 : 5;
	float acc[SLOTS][8];
#pragma unroll
	for (int j = 0; j < SLOTS; ++j)
#pragma unroll
		for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
	const float inv = cscale, bias = -8388608.0f * cscale;   // runtime values: mode 3 keeps them in registers
	for (int it = 0; it < ITERS / ROWS; ++it) {
#pragma unroll 2
		for (int r = 0; r < ROWS; ++r) {
			const uint2 w = *reinterpret_cast<const uint2 *>(&data[r][threadIdx.x * 2]);
			float wt[5];
			if (WSRC == 0) {
				const float4 wa = *reinterpret_cast<const float4 *>(&wts[r][0]);
				const float4 wb = *reinterpret_cast<const float4 *>(&wts[r][4]);
				wt[0] = wa.x; wt[1] = wa.y; wt[2] = wa.z; wt[3] = wa.w; wt[4] = wb.x;
			} else {
#pragma unroll
				for (int j = 0; j < 5; ++j) wt[j] = cwts[r * 8 + j];
			}
			float u[8];
			const unsigned ww[2] = {w.x, w.y};
#pragma unroll
			for (int h = 0; h < 2; ++h) {
				if (MODE == 5) {
					const unsigned lo = __byte_perm(ww[h], 0x64646464u, 0x4140);   // half2(1024+b1, 1024+b0)
					const unsigned hi = __byte_perm(ww[h], 0x64646464u, 0x4342);
					const __half2 k1024 = __floats2half2_rn(1024.f, 1024.f);
					const __half2 a = __hsub2(*reinterpret_cast<const __half2 *>(&lo), k1024);
					const __half2 b = __hsub2(*reinterpret_cast<const __half2 *>(&hi), k1024);
					const float2 fa = __half22float2(a), fb = __half22float2(b);
					u[4 * h] = fa.x; u[4 * h + 1] = fa.y; u[4 * h + 2] = fb.x; u[4 * h + 3] = fb.y;
				} else {
#pragma unroll
					for (int i = 0; i < 4; ++i) {
						const float m = __uint_as_float(__byte_perm(ww[h], 0x4B000000u, 0x7440 + i));
						if (MODE == 0) u[4 * h + i] = fmaf(m, 1 / 255.0f, -8388608.0f / 255.0f);
						else if (MODE == 1) u[4 * h + i] = m - 8388608.0f;
						else if (MODE == 2 || MODE == 6) u[4 * h + i] = m;
						else if (MODE == 3) u[4 * h + i] = fmaf(m, inv, bias);
						else u[4 * h + i] = (float)((ww[h] >> (8 * i)) & 0xff);
					}
				}
			}
#pragma unroll
			for (int j = 0; j < SLOTS; ++j)
#pragma unroll
				for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(wt[j], u[i], acc[j][i]);
		}
	}
	float s = 0;
#pragma unroll
	for (int j = 0; j < SLOTS; ++j)
#pragma unroll
		for (int i = 0; i < 8; ++i) s += acc[j][i];
	out[blockIdx.x * 128 + threadIdx.x] = s;
}
template <int MODE, int WSRC> void run(const char *name, int ctas_per_sm) {
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	int blocks = sms * ctas_per_sm;
	float *out;
	cudaMalloc(&out, blocks * 128 * 4);
	k<MODE, WSRC><<<blocks, 128>>>(out, 1 / 255.0f);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k<MODE, WSRC><<<blocks, 128>>>(out, 1 / 255.0f);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	double rows = (double)blocks * 4 * ITERS;
	printf("%-46s CTAs/SM %d: %.3f ms -> %.1f cycles per warp-row per SMSP\n", name, ctas_per_sm, ms,
	       ms * 1e-3 * 1.965e9 * sms * 4 / rows);
	cudaFree(out);
}
int main() {
	float h[ROWS * 8];
	for (int i = 0; i < ROWS * 8; ++i) h[i] = (i % 8 < 5) ? 1e-3f * (i % 7 + 1) : 0.f;
	cudaMemcpyToSymbol(cwts, h, sizeof(h));
	for (int c = 3; c <= 6; c += 3) {
		run<2, 0>("smem weights:  PRMT only, 40 FFMA", c);
		run<6, 0>("smem weights:  PRMT only, 32 FFMA", c);
		run<0, 0>("smem weights:  PRMT + FFMA imm, 40 FFMA", c);
		run<2, 1>("const weights: PRMT only, 40 FFMA", c);
		run<6, 1>("const weights: PRMT only, 32 FFMA", c);
		run<0, 1>("const weights: PRMT + FFMA imm, 40 FFMA", c);
		run<1, 1>("const weights: PRMT + FADD, 40 FFMA", c);
		run<4, 1>("const weights: I2F.U8, 40 FFMA", c);
	}
	printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
