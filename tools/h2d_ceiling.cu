// Platform ceiling for the end-to-end numbers: raw concurrent cudaMemcpyAsync between pinned host memory and
// 1 / 2 / 4 / 8 GPUs, one host thread and one stream pair per GPU, nothing else running.
//   nvcc -O2 -std=c++17 -o build/h2d_ceiling tools/h2d_ceiling.cu && build/h2d_ceiling [MiB per copy] [copies]
// Prints, per GPU count: aggregate H2D alone, D2H alone, and the resize mix (16 bytes up per byte down, both
// directions in flight), in GB/s.  bench.py's e2e lines are quoted against the H2D figure.
#include <cuda_runtime.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

struct PerGpu {
	void *h_up = nullptr, *h_down = nullptr, *d_up = nullptr, *d_down = nullptr;
	cudaStream_t s_up = nullptr, s_down = nullptr;
};

static double run(std::vector<PerGpu> &g, int ngpu, size_t up_bytes, size_t down_bytes, int copies) {
	std::vector<std::thread> th;
	// warm-up + barrier by joining, then the timed pass
	for (int pass = 0; pass < 2; ++pass) {
		auto t0 = std::chrono::steady_clock::now();
		th.clear();
		for (int i = 0; i < ngpu; ++i)
			th.emplace_back([&, i]() {
				CK(cudaSetDevice(i));
				for (int c = 0; c < (pass ? copies : 2); ++c) {
					if (up_bytes) CK(cudaMemcpyAsync(g[i].d_up, g[i].h_up, up_bytes, cudaMemcpyHostToDevice, g[i].s_up));
					if (down_bytes) CK(cudaMemcpyAsync(g[i].h_down, g[i].d_down, down_bytes, cudaMemcpyDeviceToHost, g[i].s_down));
				}
				CK(cudaStreamSynchronize(g[i].s_up));
				CK(cudaStreamSynchronize(g[i].s_down));
			});
		for (auto &t : th) t.join();
		if (pass) {
			double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
			return (double)(up_bytes + down_bytes) * copies * ngpu / s / 1e9;
		}
	}
	return 0;
}

int main(int argc, char **argv) {
	const size_t mib = argc > 1 ? atoi(argv[1]) : 256;
	const int copies = argc > 2 ? atoi(argv[2]) : 16;
	int ndev = 0;
	CK(cudaGetDeviceCount(&ndev));
	const size_t bytes = mib << 20;
	std::vector<PerGpu> g(ndev);
	for (int i = 0; i < ndev; ++i) {
		CK(cudaSetDevice(i));
		CK(cudaHostAlloc(&g[i].h_up, bytes, cudaHostAllocPortable));
		CK(cudaHostAlloc(&g[i].h_down, bytes, cudaHostAllocPortable));
		CK(cudaMalloc(&g[i].d_up, bytes));
		CK(cudaMalloc(&g[i].d_down, bytes));
		CK(cudaStreamCreateWithFlags(&g[i].s_up, cudaStreamNonBlocking));
		CK(cudaStreamCreateWithFlags(&g[i].s_down, cudaStreamNonBlocking));
		memset(g[i].h_up, 1, bytes);
	}
	printf("# pinned cudaMemcpyAsync, %zu MiB per copy, %d copies per GPU, one host thread per GPU; GB/s aggregate (per GPU)\n", mib, copies);
	printf("# gpus   h2d_only          d2h_only          mix_16to1 (h2d+d2h bytes)\n");
	for (int n = 1; n <= ndev; n *= 2) {
		const double up = run(g, n, bytes, 0, copies), down = run(g, n, 0, bytes, copies), mix = run(g, n, bytes, bytes / 16, copies);
		printf("%5d   %7.1f (%5.1f)   %7.1f (%5.1f)   %7.1f (%5.1f)\n", n, up, up / n, down, down / n, mix, mix / n);
	}
	// pageable source for comparison (what a Node Buffer is): one GPU, synchronous staging inside the driver
	{
		CK(cudaSetDevice(0));
		void *pg = malloc(bytes);
		memset(pg, 1, bytes);
		auto t0 = std::chrono::steady_clock::now();
		for (int c = 0; c < 4; ++c) CK(cudaMemcpy(g[0].d_up, pg, bytes, cudaMemcpyHostToDevice));
		double s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
		printf("pageable h2d, 1 GPU, cudaMemcpy: %.1f GB/s\n", bytes * 4.0 / s / 1e9);
		free(pg);
	}
	return 0;
}
