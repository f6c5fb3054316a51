// Micro-benchmark: would 16 values per thread (instead of 8) pay in the vertical-pass row body?
// Per row: NVT/4 words from shared memory, weights from the constant bank, PRMT + FMA unpack, SLOTS x NVT FMAs.
// Reported per 8 values so the two are comparable.  Not part of the product.
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
#define ROWS 16
__constant__ float cwts[ROWS * 8];
template <int NVT, int SLOTS>
__global__ void __launch_bounds__(128) k(float *out) {
	__shared__ __align__(16) unsigned data[ROWS][128 * NVT / 4];
	for (int i = threadIdx.x; i < ROWS * 128 * NVT / 4; i += 128) (&data[0][0])[i] = i * 2654435761u;
	__syncthreads();
	float acc[SLOTS][NVT];
#pragma unroll
	for (int j = 0; j < SLOTS; ++j)
#pragma unroll
		for (int i = 0; i < NVT; ++i) acc[j][i] = 0.f;
	for (int it = 0; it < ITERS / ROWS; ++it) {
#pragma unroll 1
		for (int r = 0; r < ROWS; ++r) {
			unsigned w[NVT / 4];
#pragma unroll
			for (int q = 0; q < NVT / 4; ++q) w[q] = data[r][threadIdx.x * (NVT / 4) + q];
			float u[NVT];
#pragma unroll
			for (int q = 0; q < NVT / 4; ++q)
#pragma unroll
				for (int i = 0; i < 4; ++i)
					u[4 * q + i] = fmaf(__uint_as_float(__byte_perm(w[q], 0x4B000000u, 0x7440 + i)), 1 / 255.0f, -8388608.0f / 255.0f);
#pragma unroll
			for (int j = 0; j < SLOTS; ++j)
#pragma unroll
				for (int i = 0; i < NVT; ++i) acc[j][i] = fmaf(cwts[r * 8 + j], u[i], acc[j][i]);
		}
	}
	float s = 0;
#pragma unroll
	for (int j = 0; j < SLOTS; ++j)
#pragma unroll
		for (int i = 0; i < NVT; ++i) s += acc[j][i];
	out[blockIdx.x * 128 + threadIdx.x] = s;
}
template <int NVT, int SLOTS> void run(int ctas_per_sm) {
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	int blocks = sms * ctas_per_sm;
	float *out;
	cudaMalloc(&out, blocks * 128 * 4);
	k<NVT, SLOTS><<<blocks, 128>>>(out);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k<NVT, SLOTS><<<blocks, 128>>>(out);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	double rows = (double)blocks * 4 * ITERS * (NVT / 8);   // in units of 8 values
	printf("NVT %2d  slots %d  CTAs/SM %d: %.3f ms -> %.1f cycles per warp-row of 8 values per SMSP (%s)\n", NVT, SLOTS, ctas_per_sm, ms,
	       ms * 1e-3 * 1.965e9 * sms * 4 / rows, cudaGetErrorString(cudaGetLastError()));
	cudaFree(out);
}
int main() {
	float h[ROWS * 8];
	for (int i = 0; i < ROWS * 8; ++i) h[i] = (i % 8 < 5) ? 1e-3f * (i % 7 + 1) : 0.f;
	cudaMemcpyToSymbol(cwts, h, sizeof(h));
	for (int c = 2; c <= 4; ++c) { run<8, 4>(c); run<16, 4>(c); run<8, 5>(c); run<16, 5>(c); }
	printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
