// Micro-benchmark: cost of turning 8 packed u8 channel values into floats next to 40 FFMAs
// (the shape of the resize kernel's vertical-pass row body).  Not part of the product.
//   mode 0: 40 FFMA only        mode 1: 8 x (PRMT + FFMA) + 40 FFMA       mode 2: 8 x I2F.U8 + 40 FFMA
#include <cstdio>
#include <cuda_runtime.h>
#define ITERS 4096
template <int MODE>
__global__ void __launch_bounds__(128) k(float *out, const unsigned *in, float wscale) {
	float acc[5][8];
#pragma unroll
	for (int j = 0; j < 5; ++j)
#pragma unroll
		for (int i = 0; i < 8; ++i) acc[j][i] = 0.f;
	unsigned w0 = in[threadIdx.x], w1 = in[threadIdx.x + 128];
	float wt[5];
#pragma unroll
	for (int j = 0; j < 5; ++j) wt[j] = wscale * (j + 1);
	for (int it = 0; it < ITERS; ++it) {
		float u[8];
		if (MODE == 1) {
			const float inv = 1 / 255.0f, bias = -8388608.0f * inv;
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				u[i] = fmaf(__uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7440 + i)), inv, bias);
				u[4 + i] = fmaf(__uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7440 + i)), inv, bias);
			}
		} else if (MODE == 2) {
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				u[i] = (float)((w0 >> (8 * i)) & 0xff);
				u[4 + i] = (float)((w1 >> (8 * i)) & 0xff);
			}
		} else if (MODE == 3) {
			// mask in place: bytes 0 and 1 of w and of w >> 16; the position is folded into the FMA constants
			const float inv = 1 / 255.0f;
			const unsigned t0 = w0 >> 16, t1 = w1 >> 16;
			u[0] = fmaf(__uint_as_float((w0 & 0xffu) | 0x4B000000u), inv, -8388608.0f * inv);
			u[1] = fmaf(__uint_as_float((w0 & 0xff00u) | 0x4B000000u), inv / 256, -8388608.0f * inv / 256);
			u[2] = fmaf(__uint_as_float((t0 & 0xffu) | 0x4B000000u), inv, -8388608.0f * inv);
			u[3] = fmaf(__uint_as_float((t0 & 0xff00u) | 0x4B000000u), inv / 256, -8388608.0f * inv / 256);
			u[4] = fmaf(__uint_as_float((w1 & 0xffu) | 0x4B000000u), inv, -8388608.0f * inv);
			u[5] = fmaf(__uint_as_float((w1 & 0xff00u) | 0x4B000000u), inv / 256, -8388608.0f * inv / 256);
			u[6] = fmaf(__uint_as_float((t1 & 0xffu) | 0x4B000000u), inv, -8388608.0f * inv);
			u[7] = fmaf(__uint_as_float((t1 & 0xff00u) | 0x4B000000u), inv / 256, -8388608.0f * inv / 256);
		} else if (MODE == 4) {
			// as mode 3 but without the unpack FMA: the magic floats go straight into the MACs (timing only)
			const unsigned t0 = w0 >> 16, t1 = w1 >> 16;
			u[0] = __uint_as_float((w0 & 0xffu) | 0x4B000000u); u[1] = __uint_as_float((w0 & 0xff00u) | 0x4B000000u);
			u[2] = __uint_as_float((t0 & 0xffu) | 0x4B000000u); u[3] = __uint_as_float((t0 & 0xff00u) | 0x4B000000u);
			u[4] = __uint_as_float((w1 & 0xffu) | 0x4B000000u); u[5] = __uint_as_float((w1 & 0xff00u) | 0x4B000000u);
			u[6] = __uint_as_float((t1 & 0xffu) | 0x4B000000u); u[7] = __uint_as_float((t1 & 0xff00u) | 0x4B000000u);
		} else if (MODE == 5) {
			// PRMT only (no unpack FMA): timing of the permutes themselves
#pragma unroll
			for (int i = 0; i < 4; ++i) {
				u[i] = __uint_as_float(__byte_perm(w0, 0x4B000000u, 0x7440 + i));
				u[4 + i] = __uint_as_float(__byte_perm(w1, 0x4B000000u, 0x7440 + i));
			}
		} else {
#pragma unroll
			for (int i = 0; i < 4; ++i) { u[i] = __uint_as_float(w0 + i); u[4 + i] = __uint_as_float(w1 + i); }
		}
#pragma unroll
		for (int j = 0; j < 5; ++j)
#pragma unroll
			for (int i = 0; i < 8; ++i) acc[j][i] = fmaf(wt[j], u[i], acc[j][i]);
		w0 = w0 * 1664525u + 1013904223u;
		w1 = w1 * 22695477u + 1u;
	}
	float s = 0;
#pragma unroll
	for (int j = 0; j < 5; ++j)
#pragma unroll
		for (int i = 0; i < 8; ++i) s += acc[j][i];
	out[blockIdx.x * 128 + threadIdx.x] = s;
}
template <int MODE> void run(const char *name, int ctas_per_sm) {
	int sms = 148; cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	int blocks = sms * ctas_per_sm;
	float *out; unsigned *in;
	cudaMalloc(&out, blocks * 128 * 4); cudaMalloc(&in, 256 * 4); cudaMemset(in, 0x5a, 256 * 4);
	k<MODE><<<blocks, 128>>>(out, in, 1e-3f);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0); k<MODE><<<blocks, 128>>>(out, in, 1e-3f); cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	double rows = (double)blocks * 4 * ITERS;   // warp-rows
	printf("%-28s CTAs/SM %d: %.3f ms  -> %.1f issue-cycles per warp-row at 1965 MHz (40 FFMA + unpack of 8)\n", name, ctas_per_sm, ms,
	       ms * 1e-3 * 1.965e9 * sms * 4 / rows);
	cudaFree(out); cudaFree(in);
}
int main() {
	for (int c = 3; c <= 6; c += 3) {
		run<0>("40 FFMA", c);
		run<1>("8x(PRMT+FFMA) + 40 FFMA", c);
		run<2>("8xI2F.U8 + 40 FFMA", c);
		run<3>("LOP3/SHF + 8 FFMA + 40 FFMA", c);
		run<4>("LOP3/SHF only + 40 FFMA", c);
		run<5>("8 PRMT only + 40 FFMA", c);
	}
	printf("status %s\n", cudaGetErrorString(cudaDeviceSynchronize()));
}
