// Micro-benchmark behind DESIGN.md's "FP32 issue model": how many FP32 MACs per clock per SM the
// resize inner loops can count on, and what co-issues with them.  Not part of the product.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fma_issue fma_issue.cu && ./fma_issue
#include <cstdio>
#include <cuda_runtime.h>

#define ITERS 2048

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
	unsigned long long d;
	asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
	return d;
}
__device__ __forceinline__ unsigned long long fmul2(unsigned long long a, unsigned long long b) {
	unsigned long long d;
	asm volatile("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}
__device__ __forceinline__ unsigned long long fadd2(unsigned long long a, unsigned long long b) {
	unsigned long long d;
	asm volatile("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b));
	return d;
}

// mode 0: scalar FFMA x16 chains      mode 1: FFMA2 x8 chains (same MACs)
// mode 2: FFMA2 + one PRMT per FFMA2  mode 3: FFMA2 + one LDS.32 per 2 FFMA2
// mode 4: FMUL2 + FADD2 (exact mode)  mode 5: FFMA2 + 2 PRMT per FFMA2
template <int MODE>
__global__ void __launch_bounds__(512) k(float *out, long long *cycles, float seed) {
	__shared__ float sm[1024];
	sm[threadIdx.x] = seed + threadIdx.x;
	sm[threadIdx.x + 512] = seed;
	__syncthreads();
	float a[16];
	unsigned p[8];
#pragma unroll
	for (int i = 0; i < 16; ++i) a[i] = seed + i + threadIdx.x;
#pragma unroll
	for (int i = 0; i < 8; ++i) p[i] = threadIdx.x * 2654435761u + i;
	float w0 = seed * 0.5f, w1 = seed * 0.25f;
	unsigned long long w2;
	asm("mov.b64 %0, {%1, %2};" : "=l"(w2) : "f"(w0), "f"(w1));
	unsigned long long acc[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) asm("mov.b64 %0, {%1, %2};" : "=l"(acc[i]) : "f"(a[2 * i]), "f"(a[2 * i + 1]));
	long long t0 = clock64();
	int idx = threadIdx.x;
	for (int it = 0; it < ITERS; ++it) {
		if (MODE == 0) {
#pragma unroll
			for (int i = 0; i < 16; ++i) a[i] = fmaf(a[i], w0, w1);
		} else if (MODE == 1) {
#pragma unroll
			for (int i = 0; i < 8; ++i) acc[i] = ffma2(acc[i], w2, w2);
		} else if (MODE == 2) {
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				acc[i] = ffma2(acc[i], w2, w2);
				p[i] = __byte_perm(p[i], p[(i + 1) & 7], 0x2143);
			}
		} else if (MODE == 3) {
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				acc[i] = ffma2(acc[i], w2, w2);
				if (i & 1) { a[i] += sm[idx]; idx = (idx + 33) & 1023; }
			}
		} else if (MODE == 4) {
#pragma unroll
			for (int i = 0; i < 8; ++i) acc[i] = fadd2(acc[i], fmul2(acc[i], w2));
		} else if (MODE == 5) {
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				acc[i] = ffma2(acc[i], w2, w2);
				p[i] = __byte_perm(p[i], p[(i + 1) & 7], 0x2143);
				p[(i + 3) & 7] = __byte_perm(p[(i + 3) & 7], p[i], 0x3021);
			}
		}
	}
	long long t1 = clock64();
	float s = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) s += a[i];
#pragma unroll
	for (int i = 0; i < 8; ++i) { float x, y; asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc[i])); s += x + y; s += (float)p[i]; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	if (threadIdx.x == 0) cycles[blockIdx.x] = t1 - t0;
}

template <int MODE> void run(const char *name, int blocks_per_sm) {
	int sms = 148;
	cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
	int blocks = sms * blocks_per_sm;
	float *out; long long *cyc;
	cudaMalloc(&out, blocks * 512 * sizeof(float));
	cudaMalloc(&cyc, blocks * sizeof(long long));
	k<MODE><<<blocks, 512>>>(out, cyc, 1.0f);
	cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k<MODE><<<blocks, 512>>>(out, cyc, 1.0f);
	cudaEventRecord(e1); cudaEventSynchronize(e1);
	float ms; cudaEventElapsedTime(&ms, e0, e1);
	long long *h = new long long[blocks];
	cudaMemcpy(h, cyc, blocks * sizeof(long long), cudaMemcpyDeviceToHost);
	double avg = 0; for (int i = 0; i < blocks; ++i) avg += h[i]; avg /= blocks;
	// MACs per thread per iteration: 16 in every mode (mode 4: 16 mul + 16 add = 16 "exact MACs")
	double macs_per_block = 512.0 * ITERS * 16;
	printf("%-34s blocks/SM %d  cycles/block %.0f  MAC/clk/SM %.1f  time %.3f ms  (%.2f TMAC/s)\n", name, blocks_per_sm, avg,
	       macs_per_block * blocks_per_sm / avg, ms, macs_per_block * blocks / ms / 1e9);
	cudaFree(out); cudaFree(cyc); delete[] h;
}

int main() {
	for (int b = 1; b <= 2; ++b) {
		run<0>("FFMA scalar", b);
		run<1>("FFMA2 packed", b);
		run<2>("FFMA2 + 1 PRMT each", b);
		run<5>("FFMA2 + 2 PRMT each", b);
		run<3>("FFMA2 + LDS per 2", b);
		run<4>("FMUL2+FADD2 (exact)", b);
	}
	cudaError_t e = cudaDeviceSynchronize();
	printf("status: %s\n", cudaGetErrorString(e));
	return e != cudaSuccess;
}
