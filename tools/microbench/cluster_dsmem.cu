// What a 16-CTA thread-block cluster costs on B200 for the shared-memory-resident max-flow (k_maxflow_cluster):
//   * does a 16 x 1024-thread cluster with ~200 KB of dynamic shared memory per CTA launch at all (non-portable size)
//   * float64 atomic adds into ANOTHER CTA's shared memory (generic atomicAdd on a mapped address, and the PTX
//     red.shared::cluster form): correctness of the sums and cycles per operation
//   * cluster.sync() round trip, 16-way replicated 2-byte stores, remote vs local loads
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o cluster_dsmem cluster_dsmem.cu && ./cluster_dsmem
#include <cooperative_groups.h>
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

namespace cg = cooperative_groups;

#define CK(x)                                                                                     \
	do {                                                                                          \
		cudaError_t e = (x);                                                                      \
		if (e != cudaSuccess) {                                                                   \
			printf("%s failed: %s\n", #x, cudaGetErrorString(e));                                \
			return 1;                                                                             \
		}                                                                                         \
	} while (0)

__device__ __forceinline__ unsigned map_rank(const void *p, unsigned rank) {
	unsigned a = (unsigned)__cvta_generic_to_shared(p), r;
	asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(a), "r"(rank));
	return r;
}
__device__ __forceinline__ void red_add_f64_cluster(unsigned addr, double v) {
	asm volatile("red.relaxed.cluster.shared::cluster.add.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
__device__ __forceinline__ void st_u16_cluster(unsigned addr, unsigned short v) {
	asm volatile("st.shared::cluster.u16 [%0], %1;" ::"r"(addr), "h"(v) : "memory");
}

// out[0..3]: cycles of (generic atomicAdd, red.shared::cluster, cluster.sync, replicated stores); sums[rank]: what each
// CTA's accumulator holds at the end (expected: 2 * iters * threads per CTA when targets are spread evenly)
__global__ void k_probe(int iters, long long *out, double *sums) {
	extern __shared__ double smem[];
	__shared__ double acc[32];
	__shared__ unsigned short rep[2048];
	cg::cluster_group cluster = cg::this_cluster();
	const unsigned rank = cluster.block_rank(), nranks = cluster.num_blocks();
	if (threadIdx.x < 32) acc[threadIdx.x] = 0.0;
	smem[threadIdx.x] = 1.0;
	cluster.sync();
	// generic atomicAdd on a mapped pointer: thread t adds 1.0 to slot (t % 32) of rank (rank + 1 + i) % nranks
	long long t0 = clock64();
	for (int i = 0; i < iters; ++i) {
		double *remote = cluster.map_shared_rank(&acc[threadIdx.x & 31], (rank + 1 + i) % nranks);
		atomicAdd(remote, 1.0);
	}
	cluster.sync();
	long long t1 = clock64();
	for (int i = 0; i < iters; ++i) red_add_f64_cluster(map_rank(&acc[threadIdx.x & 31], (rank + 1 + i) % nranks), 1.0);
	cluster.sync();
	long long t2 = clock64();
	for (int i = 0; i < iters; ++i) cluster.sync();
	long long t3 = clock64();
	for (int i = 0; i < iters; ++i)
		if (threadIdx.x < 64)
#pragma unroll
			for (unsigned r = 0; r < 16; ++r)
				if (r < nranks) st_u16_cluster(map_rank(&rep[threadIdx.x + 64 * (i & 15)], r), (unsigned short)i);
	cluster.sync();
	long long t4 = clock64();
	if (threadIdx.x == 0 && rank == 0) {
		out[0] = t1 - t0;
		out[1] = t2 - t1;
		out[2] = t3 - t2;
		out[3] = t4 - t3;
	}
	if (threadIdx.x == 0) {
		double s = 0;
		for (int k = 0; k < 32; ++k) s += acc[k];
		sums[rank] = s + (smem[5] - 1.0);
	}
}

int main() {
	const int threads = 1024, iters = 256;
	long long *d_out;
	double *d_sums;
	CK(cudaMalloc(&d_out, 4 * sizeof(long long)));
	CK(cudaMalloc(&d_sums, 16 * sizeof(double)));
	for (int csize : {8, 16}) {
		for (size_t smem : {(size_t)64 << 10, (size_t)200 << 10}) {
			CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
			CK(cudaFuncSetAttribute(k_probe, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
			cudaLaunchConfig_t cfg = {};
			cfg.gridDim = dim3(csize);
			cfg.blockDim = dim3(threads);
			cfg.dynamicSmemBytes = smem;
			cudaLaunchAttribute at[1];
			at[0].id = cudaLaunchAttributeClusterDimension;
			at[0].val.clusterDim.x = csize;
			at[0].val.clusterDim.y = 1;
			at[0].val.clusterDim.z = 1;
			cfg.attrs = at;
			cfg.numAttrs = 1;
			int nclusters = -1;
			cudaError_t e = cudaOccupancyMaxActiveClusters(&nclusters, k_probe, &cfg);
			printf("cluster %2d x %d threads, %3zu KB smem: max active clusters %d (%s)\n", csize, threads, smem >> 10, nclusters,
			       cudaGetErrorString(e));
			if (e != cudaSuccess || nclusters < 1) continue;
			CK(cudaMemset(d_sums, 0, 16 * sizeof(double)));
			for (int rep = 0; rep < 2; ++rep) CK(cudaLaunchKernelEx(&cfg, k_probe, iters, d_out, d_sums));
			CK(cudaDeviceSynchronize());
			long long out[4];
			double sums[16];
			CK(cudaMemcpy(out, d_out, sizeof(out), cudaMemcpyDeviceToHost));
			CK(cudaMemcpy(sums, d_sums, sizeof(sums), cudaMemcpyDeviceToHost));
			bool ok = true;
			for (int r = 0; r < csize; ++r) ok &= sums[r] == 2.0 * iters * threads;
			printf("   sums %s (rank 0 holds %.0f, expected %.0f)\n", ok ? "exact" : "WRONG", sums[0], 2.0 * iters * threads);
			printf("   generic atomicAdd f64 to a remote CTA: %.1f cycles per warp-wide op (1024 threads x %d ops)\n",
			       (double)out[0] / iters, iters);
			printf("   red.shared::cluster.add.f64          : %.1f cycles per warp-wide op\n", (double)out[1] / iters);
			printf("   cluster.sync                         : %.1f cycles\n", (double)out[2] / iters);
			printf("   64 x 16 replicated u16 stores        : %.1f cycles per batch\n", (double)out[3] / iters);
		}
	}
	// many clusters at once: how many 16-clusters run concurrently (batch mode: several cuts in flight)
	return 0;
}
