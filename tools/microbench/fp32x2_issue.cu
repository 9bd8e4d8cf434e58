// fp32x2_issue.cu -- is the packed FFMA2 (fma.rn.f32x2, sm_100) issued at the same rate as a scalar FFMA?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp32x2_issue fp32x2_issue.cu && ./fp32x2_issue
#include <cstdio>
#include <cuda_runtime.h>

constexpr int ITERS = 8192;

__device__ __forceinline__ unsigned long long ffma2(unsigned long long a, unsigned long long b, unsigned long long c) {
	unsigned long long r;
	asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
	return r;
}

template <int MODE> __global__ void __launch_bounds__(256) k(float *out, float a, float b, int n) {
	float x[16];
#pragma unroll
	for (int i = 0; i < 16; ++i) x[i] = threadIdx.x * 1e-3f + i;
	unsigned long long *x2 = reinterpret_cast<unsigned long long *>(x);
	float2 aa = make_float2(a, a), bb = make_float2(b, b);
	const unsigned long long a2 = *reinterpret_cast<unsigned long long *>(&aa), b2 = *reinterpret_cast<unsigned long long *>(&bb);
#pragma unroll 1
	for (int it = 0; it < n; ++it) {
		if (MODE == 0) {
#pragma unroll
			for (int i = 0; i < 16; ++i) x[i] = fmaf(x[i], a, b);
		} else if (MODE == 1) {
#pragma unroll
			for (int i = 0; i < 8; ++i) x2[i] = ffma2(x2[i], a2, b2);
		} else { // three distinct register sources
#pragma unroll
			for (int i = 0; i < 8; ++i) x2[i] = ffma2(x2[i], x2[(i + 3) & 7], x2[(i + 5) & 7]);
		}
	}
	float s = 0;
#pragma unroll
	for (int i = 0; i < 16; ++i) s += x[i];
	out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> void run(const char *name, int instr_per_iter, int flops_per_iter, float *out, int sms, double ghz) {
	const int blocks = sms * 4;
	k<MODE><<<blocks, 256>>>(out, 1.0000001f, 1e-9f, 16);
	cudaEvent_t e0, e1;
	cudaEventCreate(&e0);
	cudaEventCreate(&e1);
	cudaEventRecord(e0);
	k<MODE><<<blocks, 256>>>(out, 1.0000001f, 1e-9f, ITERS);
	cudaEventRecord(e1);
	cudaEventSynchronize(e1);
	float ms;
	cudaEventElapsedTime(&ms, e0, e1);
	const double cycles = ms * 1e-3 * ghz * 1e9;
	printf("%-40s %8.3f ms  cycles per warp-instr per sub-partition %.3f   %.1f TFLOP/s\n", name, ms,
	       cycles / (8.0 * ITERS * instr_per_iter), 2.0 * flops_per_iter * ITERS * 256.0 * blocks / (ms * 1e-3) / 1e12);
}

int main() {
	cudaDeviceProp p;
	cudaGetDeviceProperties(&p, 0);
	int khz = 0;
	cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
	const double ghz = khz * 1e-6;
	printf("%s, %d SMs, %.3f GHz\n", p.name, p.multiProcessorCount, ghz);
	float *out;
	cudaMalloc(&out, sizeof(float) * 256 * p.multiProcessorCount * 4);
	run<0>("FFMA x16 (scalar)", 16, 16, out, p.multiProcessorCount, ghz);
	run<1>("FFMA2 x8 (packed, 2 uniform sources)", 8, 16, out, p.multiProcessorCount, ghz);
	run<2>("FFMA2 x8 (packed, 3 register sources)", 8, 16, out, p.multiProcessorCount, ghz);
	cudaDeviceSynchronize();
	printf("%s\n", cudaGetErrorString(cudaGetLastError()));
	return 0;
}
