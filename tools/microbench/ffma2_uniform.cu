// Micro-benchmark: does fma.rn.f32x2 with a warp-uniform weight pair (from the constant bank) issue
// faster than with three vector-register pairs?  16 MACs per thread per inner step in every mode.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__constant__ float cw[128];
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk(float a, float b) { u64 v; asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b)); return v; }
__device__ __forceinline__ void ffma2(u64 &acc, u64 a, u64 b) { asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc) : "l"(a), "l"(b)); }
template <int MODE>
__global__ void __launch_bounds__(256) k(float *out, int seed) {
	float f[16];
	u64 acc[8], d[4];
#pragma unroll
	for (int i = 0; i < 16; ++i) f[i] = (float)(seed + i + threadIdx.x);
#pragma unroll
	for (int i = 0; i < 8; ++i) acc[i] = pk(f[2 * i], f[2 * i + 1]);
#pragma unroll
	for (int i = 0; i < 4; ++i) d[i] = pk(1.0f + threadIdx.x + i, 2.0f + threadIdx.x * i);
	for (int it = 0; it < ITERS; ++it) {
		const float w = cw[it & 127];
		if (MODE == 0) {          // scalar FFMA, uniform weight
			float u[4]; u[0] = __uint_as_float((unsigned)d[0]); u[1] = __uint_as_float((unsigned)d[1]); u[2] = __uint_as_float((unsigned)d[2]); u[3] = __uint_as_float((unsigned)d[3]);
#pragma unroll
			for (int i = 0; i < 16; ++i) f[i] = fmaf(w, u[i & 3], f[i]);
		} else if (MODE == 1) {   // FFMA2, weight pair built from the uniform value
			const u64 ww = pk(w, w);
#pragma unroll
			for (int i = 0; i < 8; ++i) ffma2(acc[i], d[i & 3], ww);
		} else {                  // FFMA2, weight pair in per-thread registers
			const u64 ww = pk(w + threadIdx.x, w);
#pragma unroll
			for (int i = 0; i < 8; ++i) ffma2(acc[i], d[i & 3], ww);
		}
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) s += f[i];
#pragma unroll
	for (int i = 0; i < 8; ++i) s += __uint_as_float((unsigned)acc[i]) + __uint_as_float((unsigned)(acc[i] >> 32));
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name) {
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	const int blocks = sms * 4;
	float *out;
	cudaMalloc(&out, blocks * 256 * 4);
	k<MODE><<<blocks, 256>>>(out, 1);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k<MODE><<<blocks, 256>>>(out, 1);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double macs = (double)blocks * 256 * ITERS * 16;
	printf("%-44s %.3f ms  %.1f MAC/clk/SM (%s)\n", name, ms, macs / (ms * 1e-3 * 1.965e9 * sms), cudaGetErrorString(cudaGetLastError()));
	cudaFree(out);
}
int main() {
	float h[128];
	for (int i = 0; i < 128; ++i) h[i] = 1e-3f * (i + 1);
	cudaMemcpyToSymbol(cw, h, sizeof(h));
	run<0>("FFMA, uniform weight");
	run<1>("FFMA2, {w,w} from a uniform value");
	run<2>("FFMA2, per-thread weight pair");
	return cudaDeviceSynchronize() != cudaSuccess;
}
