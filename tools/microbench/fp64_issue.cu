// fp64_issue.cu -- what bounds an FP64-heavy loop on B200: the FP64 pipe alone, or FP64 + everything else?
// Measures cycles per warp-instruction per SM sub-partition for (a) pure DFMA, (b) DFMA with k integer instructions
// interleaved per 4 DFMA, (c) DMUL/DADD mixes, (d) DFMA with three distinct source registers.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_issue fp64_issue.cu && ./fp64_issue
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 4096;

template <int MODE, int NINT>
__global__ void __launch_bounds__(256) k(double *out, int *iout, double a, double b, int n) {
	double x[8];
	int y[8];
#pragma unroll
	for (int i = 0; i < 8; ++i) { x[i] = threadIdx.x * 1e-3 + i; y[i] = threadIdx.x + i; }
	double c = a + 1.0, d = b + 2.0;
#pragma unroll 1
	for (int it = 0; it < n; ++it) {
#pragma unroll
		for (int i = 0; i < 8; ++i) {
			if (MODE == 0) x[i] = __fma_rn(x[i], a, b);              // 2 shared sources
			if (MODE == 1) x[i] = __dadd_rn(__dmul_rn(x[i], a), b);  // DMUL + DADD
			if (MODE == 2) x[i] = __fma_rn(x[i], x[(i + 3) & 7], x[(i + 5) & 7]); // 3 distinct register sources
			if (MODE == 3) x[i] = __dmul_rn(x[i], a);
			if (MODE == 4) x[i] = __dadd_rn(x[i], a);
		}
#pragma unroll
		for (int i = 0; i < NINT; ++i) y[i & 7] = (y[i & 7] ^ (y[(i + 1) & 7] + it)) + 0x9e3779b9;
	}
	double s = c * 0 + d * 0;
	int t = 0;
#pragma unroll
	for (int i = 0; i < 8; ++i) { s += x[i]; t ^= y[i]; }
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
	iout[blockIdx.x * blockDim.x + threadIdx.x] = t;
}

template <int MODE, int NINT> void run(const char *name, int fp64_per_iter, double *out, int *iout, int sms, double ghz) {
	const int blocks = sms * 4; // 4 blocks x 8 warps = 32 warps / SM = 8 per sub-partition
	k<MODE, NINT><<<blocks, 256>>>(out, iout, 1.0000001, 1e-9, 16);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k<MODE, NINT><<<blocks, 256>>>(out, iout, 1.0000001, 1e-9, ITERS);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double warp_instr_per_smsp = 8.0 * ITERS * (fp64_per_iter + NINT); // 8 warps per sub-partition
	const double cycles = ms * 1e-3 * ghz * 1e9;
	printf("%-44s %8.3f ms  cycles/iter/warp-set %7.2f  cycles per FP64 warp-instr %.3f  (fp64 %d + int %d per iter)\n", name, ms,
	       cycles / (8.0 * ITERS), cycles / (8.0 * ITERS * fp64_per_iter), fp64_per_iter, NINT);
	(void)warp_instr_per_smsp;
}

int main() {
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	const double ghz = khz * 1e-6;
	printf("%s, %d SMs, %.3f GHz (nominal max; cycles below assume it)\n", p.name, p.multiProcessorCount, ghz);
	double *out;
	int *iout;
	cudaMalloc(&out, sizeof(double) * 256 * p.multiProcessorCount * 4);
	cudaMalloc(&iout, sizeof(int) * 256 * p.multiProcessorCount * 4);
	const int s = p.multiProcessorCount;
	run<0, 0>("DFMA x8 (2 uniform sources)", 8, out, iout, s, ghz);
	run<0, 2>("DFMA x8 + 2 int", 8, out, iout, s, ghz);
	run<0, 4>("DFMA x8 + 4 int", 8, out, iout, s, ghz);
	run<0, 8>("DFMA x8 + 8 int", 8, out, iout, s, ghz);
	run<0, 16>("DFMA x8 + 16 int", 8, out, iout, s, ghz);
	run<1, 0>("DMUL+DADD x8", 16, out, iout, s, ghz);
	run<1, 8>("DMUL+DADD x8 + 8 int", 16, out, iout, s, ghz);
	run<2, 0>("DFMA x8 (3 distinct register sources)", 8, out, iout, s, ghz);
	run<2, 4>("DFMA x8 (3 distinct) + 4 int", 8, out, iout, s, ghz);
	run<3, 0>("DMUL x8", 8, out, iout, s, ghz);
	run<4, 0>("DADD x8", 8, out, iout, s, ghz);
	cudaDeviceSynchronize();
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
