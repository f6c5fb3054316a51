// Micro-benchmark: the downscaling kernel's row body (4 LDS.32 + 16 PRMT + 4 x 16 MACs with warp-uniform
// weights from the constant bank) with scalar FFMAs against packed fma.rn.f32x2 with a broadcast weight.
// Reports cycles per warp-row per scheduler.  Not part of the product.
#include <cstdio>
#include <cuda_runtime.h>
#ifndef REPS
#define REPS 1
#endif
#define ROWS 32
#ifndef ITERS
#define ITERS 256
#endif
__constant__ float cw[ROWS * 4];
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 v; asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b)); return v; }
__device__ __forceinline__ void ffma2(u64 &acc, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
template <int MODE>
__global__ void __launch_bounds__(64) k(float *out, unsigned long long *clk) {
	unsigned long long c0 = clock64(), g0;
	asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g0));
	__shared__ unsigned data[ROWS][256];
	for (int i = threadIdx.x; i < ROWS * 256; i += 64) (&data[0][0])[i] = i * 2654435761u;
	__syncthreads();
	float acc[4][16];
	u64 acc2[4][8];
#pragma unroll
	for (int j = 0; j < 4; ++j) {
#pragma unroll
		for (int i = 0; i < 16; ++i) acc[j][i] = 0.f;
#pragma unroll
		for (int i = 0; i < 8; ++i) acc2[j][i] = 0;
	}
	for (int it = 0; it < ITERS; ++it) {
#pragma unroll 2
		for (int r = 0; r < ROWS; ++r) {
			unsigned w[4];
			if (MODE == 4) {
				const uint4 v = *reinterpret_cast<const uint4 *>(&data[r][4 * threadIdx.x]);
				w[0] = v.x; w[1] = v.y; w[2] = v.z; w[3] = v.w;
			} else {
#pragma unroll
				for (int q = 0; q < 4; ++q) w[q] = data[r][threadIdx.x + 64 * q];
			}
			float wt[4];
#pragma unroll
			for (int j = 0; j < 4; ++j) wt[j] = cw[r * 4 + j];
			float u[16];
#pragma unroll
			for (int i = 0; i < 16; ++i) u[i] = __uint_as_float(__byte_perm(w[i >> 2], 0, 0x4440 + (i & 3)));
			if (MODE == 2 || MODE == 3) {
				// timing experiments: what the PRMTs cost next to the FFMA2s -- values arrive as whole words
				// (MODE 3: all 16 from four LDS.128, no PRMT; MODE 2: the upper 8, 8 PRMTs left)
#pragma unroll
				for (int q = (MODE == 2 ? 2 : 0); q < 4; ++q) {
					const uint4 v = *reinterpret_cast<const uint4 *>(&data[(r + q) % ROWS][4 * threadIdx.x]);
					u[4 * q] = __uint_as_float(v.x); u[4 * q + 1] = __uint_as_float(v.y);
					u[4 * q + 2] = __uint_as_float(v.z); u[4 * q + 3] = __uint_as_float(v.w);
				}
			}
			if (MODE == 0) {
#pragma unroll
				for (int i = 0; i < 16; ++i)
#pragma unroll
					for (int j = 0; j < 4; ++j) acc[j][i] = fmaf(wt[j], u[i], acc[j][i]);
			} else {
#pragma unroll
				for (int i = 0; i < 8; ++i) {
					const u64 uu = pk(u[2 * i], u[2 * i + 1]);
#pragma unroll
					for (int j = 0; j < 4; ++j) ffma2(acc2[j][i], uu, pk(wt[j], wt[j]));
				}
			}
		}
	}
	float s = 0;
#pragma unroll
	for (int j = 0; j < 4; ++j) {
#pragma unroll
		for (int i = 0; i < 16; ++i) s += acc[j][i];
#pragma unroll
		for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)acc2[j][i]) + __uint_as_float((unsigned)(acc2[j][i] >> 32));
	}
	out[blockIdx.x * 64 + threadIdx.x] = s;
	if (threadIdx.x == 0 && clk) {
		unsigned long long g1;
		asm volatile("mov.u64 %0, %globaltimer;" : "=l"(g1));
		clk[2 * blockIdx.x] = clock64() - c0;
		clk[2 * blockIdx.x + 1] = g1 - g0;
	}
}
template <int MODE> void run(const char *name, int ctas_per_sm) {
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	const int blocks = sms * ctas_per_sm;
	float *out;
	cudaMalloc(&out, blocks * 64 * 4);
	unsigned long long *clk;
	cudaMalloc(&clk, blocks * 16);
	k<MODE><<<blocks, 64>>>(out, nullptr);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	for (int rep = 0; rep < REPS; ++rep) k<MODE><<<blocks, 64>>>(out, clk);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	unsigned long long h[2];
	cudaMemcpy(h, clk, 16, cudaMemcpyDeviceToHost);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double warp_rows = (double)blocks * 2 * ITERS * ROWS * REPS;
	const double mhz = h[1] ? (double)h[0] / (double)h[1] * 1e3 : 0;   // SM clock while the last launch ran (CTA 0)
	printf("%-26s CTAs/SM %d: %.3f ms -> %.1f cycles per warp-row per scheduler at 1965 MHz, %.1f at the measured %.0f MHz (%s)\n",
	       name, ctas_per_sm, ms, ms * 1e-3 * 1.965e9 * sms * 4 / warp_rows, ms * 1e-3 * mhz * 1e6 * sms * 4 / warp_rows, mhz,
	       cudaGetErrorString(cudaGetLastError()));
	cudaFree(clk);
	cudaFree(out);
}
int main() {
	float h[ROWS * 4];
	for (int i = 0; i < ROWS * 4; ++i) h[i] = 1e-3f * (i % 7 + 1);
	cudaMemcpyToSymbol(cw, h, sizeof(h));
	for (int c : {6}) {
		run<0>("64 FFMA (uniform weight)", c);
		run<1>("32 FFMA2 (broadcast weight)", c);
		run<2>("32 FFMA2, 8 PRMT, 2 LDS.128", c);
		run<3>("32 FFMA2, 0 PRMT, 4 LDS.128", c);
		run<4>("32 FFMA2, one LDS.128", c);
	}
	return cudaDeviceSynchronize() != cudaSuccess;
}
