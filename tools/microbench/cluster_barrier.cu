// What one BFS level of k_maxflow_cluster costs in synchronisation: K threads per CTA store 2 bytes into all 16 replicas,
// then the cluster meets at a barrier. Variants of the memory ordering around the barrier:
//   0  barrier.cluster.arrive.release + wait.acquire by every thread (MEMBAR.ALL.GPU in every thread)
//   1  arrive.relaxed + wait.acquire; only the storing threads run fence.acq_rel.cluster first
//   2  arrive.relaxed + wait.acquire; the storing threads read their last store back from every replica instead of a fence
//   3  arrive.relaxed + wait (no ordering at all: lower bound)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_barrier cluster_barrier.cu && ./cluster_barrier
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned mapa(unsigned a, unsigned r) {
	unsigned o;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(o) : "r"(a), "r"(r));
	return o;
}
template <int MODE> __global__ void k(int iters, int K, long long *out, int *sink) {
	extern __shared__ unsigned short rep[];
	unsigned rank, nranks;
	asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
	asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(nranks));
	const unsigned my = (unsigned)__cvta_generic_to_shared(rep + rank * 1024 + threadIdx.x);
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
	int bad = 0;
	const long long t0 = clock64();
	for (int i = 1; i <= iters; ++i) {
		if ((int)threadIdx.x < K) {
			for (unsigned r = 0; r < nranks; ++r) asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(mapa(my, r)), "h"((unsigned short)i) : "memory");
			if (MODE == 1) asm volatile("fence.acq_rel.cluster;" ::: "memory");
			if (MODE == 2) {
				unsigned acc = 0;
				for (unsigned r = 0; r < nranks; ++r) {
					unsigned short v;
					asm volatile("ld.volatile.shared::cluster.u16 %0, [%1];" : "=h"(v) : "r"(mapa(my, r)) : "memory");
					acc += v;
				}
				if (acc != nranks * (unsigned)(unsigned short)i) bad++;
			}
		}
		if (MODE == 0) asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
		if (MODE == 1 || MODE == 2) asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
		if (MODE == 3) asm volatile("barrier.cluster.arrive.relaxed.aligned;\n\tbarrier.cluster.wait.aligned;" ::: "memory");
		// every thread checks a value another CTA stored in this iteration
		const unsigned other = (rank + 1 + threadIdx.x) % nranks;
		if ((int)(threadIdx.x % 1024) < K && ((volatile unsigned short *)rep)[other * 1024 + threadIdx.x % 1024] != (unsigned short)i) bad++;
	}
	const long long t1 = clock64();
	if (threadIdx.x == 0 && rank == 0) out[0] = t1 - t0;
	if (bad) atomicAdd(sink, bad);
	asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

template <int MODE> int run(int K, long long *d_out, int *d_sink) {
	const int iters = 2000, csize = 16;
	size_t smem = 200 << 10;
	cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
	cudaFuncSetAttribute(k<MODE>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
	cudaLaunchConfig_t cfg = {};
	cfg.gridDim = dim3(csize);
	cfg.blockDim = dim3(1024);
	cfg.dynamicSmemBytes = smem;
	cudaLaunchAttribute at[1];
	at[0].id = cudaLaunchAttributeClusterDimension;
	at[0].val.clusterDim.x = csize;
	at[0].val.clusterDim.y = 1;
	at[0].val.clusterDim.z = 1;
	cfg.attrs = at;
	cfg.numAttrs = 1;
	cudaMemset(d_sink, 0, sizeof(int));
	for (int rep = 0; rep < 2; ++rep) cudaLaunchKernelEx(&cfg, k<MODE>, iters, K, d_out, d_sink);
	cudaError_t e = cudaDeviceSynchronize();
	long long out;
	int sink;
	cudaMemcpy(&out, d_out, sizeof(out), cudaMemcpyDeviceToHost);
	cudaMemcpy(&sink, d_sink, sizeof(sink), cudaMemcpyDeviceToHost);
	printf("mode %d, %4d storing threads per CTA: %7.0f cycles per level, stale reads %d (%s)\n", MODE, K, (double)out / iters, sink, cudaGetErrorString(e));
	return 0;
}

int main() {
	long long *d_out;
	int *d_sink;
	cudaMalloc(&d_out, sizeof(long long));
	cudaMalloc(&d_sink, sizeof(int));
	for (int K : {0, 16, 64, 1024}) {
		run<0>(K, d_out, d_sink);
		run<1>(K, d_out, d_sink);
		run<2>(K, d_out, d_sink);
		run<3>(K, d_out, d_sink);
	}
	return 0;
}
