// Micro-benchmark: issue rate of the integer dot-product instructions (dp2a / dp4a) against FFMA
// with a uniform weight operand, 16 independent chains per thread.  Not part of the product.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
__constant__ int cw[64];
template <int MODE>
__global__ void __launch_bounds__(256) k(int *out, int seed) {
	int acc[16];
	float facc[16];
	unsigned d[4];
#pragma unroll
	for (int i = 0; i < 16; ++i) { acc[i] = seed + i + threadIdx.x; facc[i] = (float)(seed + i); }
#pragma unroll
	for (int i = 0; i < 4; ++i) d[i] = threadIdx.x * 2654435761u + i + seed;
	for (int it = 0; it < ITERS; ++it) {
		const int w0 = cw[it & 63], w1 = cw[(it + 1) & 63];
#pragma unroll
		for (int i = 0; i < 16; ++i) {
			if (MODE == 0) facc[i] = fmaf(__int_as_float(w0), __uint_as_float(d[i & 3]), facc[i]);
			else if (MODE == 1) acc[i] = __dp2a_lo(w0, (int)d[i & 3], acc[i]);
			else if (MODE == 2) acc[i] = __dp4a((int)d[i & 3], w0, acc[i]);
			else if (MODE == 3) { if (i & 1) acc[i] = __dp2a_hi(w1, (int)d[i & 3], acc[i]); else acc[i] = __dp2a_lo(w0, (int)d[i & 3], acc[i]); }
			else acc[i] = acc[i] + w0 * (int)d[i & 3];   // IMAD
		}
	}
	int s = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) s += acc[i] + (int)facc[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int MODE> void run(const char *name) {
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	const int blocks = sms * 4;
	int *out;
	cudaMalloc(&out, blocks * 256 * 4);
	k<MODE><<<blocks, 256>>>(out, 1);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k<MODE><<<blocks, 256>>>(out, 1);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double inst = (double)blocks * 256 * ITERS * 16;
	printf("%-28s %.3f ms  %.1f thread-instr/clk/SM (%s)\n", name, ms, inst / (ms * 1e-3 * 1.965e9 * sms), cudaGetErrorString(cudaGetLastError()));
	cudaFree(out);
}
int main() {
	int h[64];
	for (int i = 0; i < 64; ++i) h[i] = 0x00010001 * (i + 1);
	cudaMemcpyToSymbol(cw, h, sizeof(h));
	run<0>("FFMA uniform weight");
	run<1>("DP2A.LO uniform weight");
	run<3>("DP2A.LO/HI uniform weight");
	run<2>("DP4A uniform weight");
	run<4>("IMAD uniform weight");
	return cudaDeviceSynchronize() != cudaSuccess;
}
