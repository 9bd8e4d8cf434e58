// Dependent-load latency seen by ONE warp of ONE block (the situation of the single-block global relabel of k_maxflow):
// a pointer chase through a 4 MB table (L2-resident, far larger than L1) with the load flavours the max-flow kernel
// uses. Prints cycles and ns per dependent load.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o l2_latency l2_latency.cu && ./l2_latency
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_runtime.h>

template <int MODE> __device__ __forceinline__ int load(const int *p) {
	int v;
	if (MODE == 0) asm volatile("ld.global.ca.s32 %0, [%1];" : "=r"(v) : "l"(p));
	if (MODE == 1) asm volatile("ld.global.cg.s32 %0, [%1];" : "=r"(v) : "l"(p));
	if (MODE == 2) asm volatile("ld.volatile.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
	if (MODE == 3) asm volatile("ld.global.nc.s32 %0, [%1];" : "=r"(v) : "l"(p));
	return v;
}

template <int MODE> __global__ void chase(const int *table, int steps, int lanes, long long *cycles, int *sink) {
	int i = threadIdx.x * 97 % 1000;
	const bool on = (int)threadIdx.x < lanes;
	const long long t0 = clock64();
	for (int s = 0; s < steps; ++s)
		if (on) i = load<MODE>(table + i);
	const long long t1 = clock64();
	if (threadIdx.x == 0) *cycles = t1 - t0;
	if (i == -1) *sink = i;
}

int main() {
	const int n = 1 << 20; // 4 MB of ints
	std::vector<int> h(n);
	// one random cycle over the table: every load lands in another 128-byte line
	std::vector<int> perm(n);
	for (int i = 0; i < n; ++i) perm[i] = i;
	srand(1);
	for (int i = n - 1; i > 0; --i) std::swap(perm[i], perm[rand() % (i + 1)]);
	for (int i = 0; i < n; ++i) h[perm[i]] = perm[(i + 1) % n];
	int *d, *sink;
	long long *cyc;
	cudaMalloc(&d, sizeof(int) * n);
	cudaMalloc(&sink, sizeof(int));
	cudaMalloc(&cyc, sizeof(long long));
	cudaMemcpy(d, h.data(), sizeof(int) * n, cudaMemcpyHostToDevice);
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	const char *names[4] = {"ld.global.ca", "ld.global.cg", "ld.volatile.global", "ld.global.nc"};
	const int steps = 20000;
	for (int lanes : {1, 32}) {
		for (int mode = 0; mode < 4; ++mode) {
			for (int rep = 0; rep < 2; ++rep) { // second pass: table warm in L2
				if (mode == 0) chase<0><<<1, 32>>>(d, steps, lanes, cyc, sink);
				if (mode == 1) chase<1><<<1, 32>>>(d, steps, lanes, cyc, sink);
				if (mode == 2) chase<2><<<1, 32>>>(d, steps, lanes, cyc, sink);
				if (mode == 3) chase<3><<<1, 32>>>(d, steps, lanes, cyc, sink);
				cudaDeviceSynchronize();
			}
			long long c = 0;
			cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
			printf("%-20s lanes=%2d  %7.1f cycles per dependent load  (%.0f ns at %d MHz)\n", names[mode], lanes, (double)c / steps,
			       (double)c / steps / (khz / 1e6), khz / 1000);
		}
	}
	return 0;
}
