// pxb_nccl.cu -- the multi-GPU exchange steps of the hot path over NCCL / NVLink.
//
// One process per GPU. The path has exactly two exchange steps and both are all-gathers of small records:
//   * hypothesis blocks of ONE large problem (C3 / C5): the coordinator (rank 0) broadcasts a block of minimal samples,
//     every rank solves and scores its contiguous slice over all N points (points are replicated), and one in-place
//     ncclAllGather returns every slice's (models, flags, count, value, shared) record to every rank
//     (Driver::solve_and_score_sharded in pxb_driver.cu uses shard_broadcast / shard_allgather below);
//   * independent problems (C4): pxb_allgather_instances merges the surviving instances of every rank's pairs.
//
// NCCL is resolved at RUN TIME (dlopen of the libnccl.so.2 the process already holds -- torch's bundled copy when the
// caller is a torch.distributed job -- else the system one), so libpxb200.so has no link-time dependency on it and still
// loads on a box without NCCL; every entry point here fails with PXB_ERR_UNSUPPORTED in that case. Only the handful of
// prototypes used are declared (they are stable across NCCL 2.x: nccl.h:146-186,379-430).
#include <dlfcn.h>

#include <algorithm>
#include <cstring>
#include <mutex>

#include "pxb_internal.h"

namespace pxb {
namespace {

struct NcclUniqueId { // nccl.h:37-38
	char internal[128];
};
typedef void *NcclComm;
enum { kNcclSuccess = 0, kNcclInt8 = 0 }; // ncclResult_t / ncclDataType_t (ncclChar == ncclInt8 == 0)

struct NcclApi {
	void *handle = nullptr;
	int (*GetVersion)(int *) = nullptr;
	int (*GetUniqueId)(NcclUniqueId *) = nullptr;
	int (*CommInitRank)(NcclComm *, int, NcclUniqueId, int) = nullptr;
	int (*CommDestroy)(NcclComm) = nullptr;
	int (*CommCount)(NcclComm, int *) = nullptr;
	int (*CommUserRank)(NcclComm, int *) = nullptr;
	int (*Broadcast)(const void *, void *, size_t, int, int, NcclComm, cudaStream_t) = nullptr;
	int (*AllGather)(const void *, void *, size_t, int, NcclComm, cudaStream_t) = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
	bool ok = false;
	std::string why;
};

NcclApi &api() {
	static NcclApi a;
	static std::once_flag once;
	std::call_once(once, [] {
		const char *names[] = {"libnccl.so.2", "libnccl.so"};
		for (const char *n : names) { // the copy this process already mapped, if any (torch ships its own)
			a.handle = dlopen(n, RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);
			if (a.handle) break;
		}
		if (!a.handle)
			if (const char *e = getenv("PXB_NCCL_LIB")) a.handle = dlopen(e, RTLD_NOW | RTLD_GLOBAL);
		for (const char *n : names) {
			if (a.handle) break;
			a.handle = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
		}
		if (!a.handle) {
			a.why = std::string("libnccl.so.2 not found (") + (dlerror() ? dlerror() : "dlopen failed") + ")";
			return;
		}
#define PXB_SYM(field, name)                                                  \
	do {                                                                      \
		*reinterpret_cast<void **>(&a.field) = dlsym(a.handle, name);         \
		if (!a.field) {                                                       \
			a.why = std::string("symbol ") + name + " missing from libnccl"; \
			return;                                                           \
		}                                                                     \
	} while (0)
		PXB_SYM(GetVersion, "ncclGetVersion");
		PXB_SYM(GetUniqueId, "ncclGetUniqueId");
		PXB_SYM(CommInitRank, "ncclCommInitRank");
		PXB_SYM(CommDestroy, "ncclCommDestroy");
		PXB_SYM(CommCount, "ncclCommCount");
		PXB_SYM(CommUserRank, "ncclCommUserRank");
		PXB_SYM(Broadcast, "ncclBroadcast");
		PXB_SYM(AllGather, "ncclAllGather");
		PXB_SYM(GetErrorString, "ncclGetErrorString");
#undef PXB_SYM
		a.ok = true;
	});
	return a;
}

int require_nccl() {
	NcclApi &a = api();
	if (!a.ok) {
		set_error("NCCL unavailable: %s", a.why.c_str());
		return PXB_ERR_UNSUPPORTED;
	}
	return PXB_OK;
}

#define PXB_NCCL(call)                                                                                          \
	do {                                                                                                        \
		const int r__ = (call);                                                                                 \
		if (r__ != kNcclSuccess) {                                                                              \
			set_error("%s failed: %s (%s:%d)", #call, api().GetErrorString(r__), __FILE__, __LINE__);          \
			return PXB_ERR_CUDA;                                                                                \
		}                                                                                                       \
	} while (0)

} // namespace

// ---- used by the sharded driver (device pointers, asynchronous on ctx->stream) ------------------------------------
int shard_broadcast(pxb_ctx *ctx, void *buf_dev, size_t bytes, int root) {
	PXB_TRY(require_nccl());
	if (bytes == 0) return PXB_OK;
	PXB_NCCL(api().Broadcast(buf_dev, buf_dev, bytes, kNcclInt8, root, ctx->shard_comm, ctx->stream));
	return PXB_OK;
}
// in place: rank r's record sits at recv_dev + r * bytes_per_rank before the call
int shard_allgather(pxb_ctx *ctx, void *recv_dev, size_t bytes_per_rank) {
	PXB_TRY(require_nccl());
	if (bytes_per_rank == 0) return PXB_OK;
	const char *send = reinterpret_cast<const char *>(recv_dev) + (size_t)ctx->shard_rank * bytes_per_rank;
	PXB_NCCL(api().AllGather(send, recv_dev, bytes_per_rank, kNcclInt8, ctx->shard_comm, ctx->stream));
	return PXB_OK;
}

} // namespace pxb

using namespace pxb;

extern "C" {

int pxb_nccl_version(int *version) {
	PXB_CHECK_ARG(version != nullptr, "null argument");
	PXB_TRY(require_nccl());
	PXB_NCCL(api().GetVersion(version));
	return PXB_OK;
}

int pxb_nccl_unique_id(void *id128) {
	PXB_CHECK_ARG(id128 != nullptr, "null argument");
	PXB_TRY(require_nccl());
	NcclUniqueId id;
	PXB_NCCL(api().GetUniqueId(&id));
	std::memcpy(id128, id.internal, sizeof(id.internal));
	return PXB_OK;
}

int pxb_nccl_comm_init(pxb_ctx *ctx, const void *id128, int world, int rank, void **nccl_comm_out) {
	PXB_CHECK_ARG(ctx && id128 && nccl_comm_out && world >= 1 && rank >= 0 && rank < world, "bad argument");
	PXB_TRY(require_nccl());
	PXB_CUDA(cudaSetDevice(ctx->device));
	NcclUniqueId id;
	std::memcpy(id.internal, id128, sizeof(id.internal));
	NcclComm comm = nullptr;
	PXB_NCCL(api().CommInitRank(&comm, world, id, rank));
	*nccl_comm_out = comm;
	return PXB_OK;
}

int pxb_nccl_comm_destroy(void *nccl_comm) {
	if (!nccl_comm) return PXB_OK;
	PXB_TRY(require_nccl());
	PXB_NCCL(api().CommDestroy(nccl_comm));
	return PXB_OK;
}

int pxb_ctx_set_shard(pxb_ctx *ctx, void *nccl_comm) {
	PXB_CHECK_ARG(ctx != nullptr, "null context");
	if (!nccl_comm) {
		ctx->shard_comm = nullptr;
		ctx->shard_world = 1;
		ctx->shard_rank = 0;
		return PXB_OK;
	}
	PXB_TRY(require_nccl());
	int world = 0, rank = 0;
	PXB_NCCL(api().CommCount(nccl_comm, &world));
	PXB_NCCL(api().CommUserRank(nccl_comm, &rank));
	ctx->shard_comm = nccl_comm;
	ctx->shard_world = world;
	ctx->shard_rank = rank;
	return PXB_OK;
}

int pxb_shard_info(pxb_ctx *ctx, int *world, int *rank) {
	PXB_CHECK_ARG(ctx && world && rank, "null argument");
	*world = ctx->shard_comm ? ctx->shard_world : 1;
	*rank = ctx->shard_comm ? ctx->shard_rank : 0;
	return PXB_OK;
}

// Independent problems (C4): pair p lives on rank p % world in local slot p / world. Every rank contributes
// `pairs_per_rank` fixed-size records {count int32 | models max_models x model_size f64 | labels n_points int32}; one
// ncclAllGather per field returns all of them to every rank, rank-major.
int pxb_allgather_instances(pxb_ctx *ctx, void *nccl_comm, int64_t pairs_per_rank, int64_t n_points, int32_t model_size,
                            int32_t max_models, const int32_t *counts_host, const double *models_host,
                            const int32_t *labels_host, int32_t *counts_out_host, double *models_out_host,
                            int32_t *labels_out_host) {
	PXB_CHECK_ARG(ctx && nccl_comm && counts_host && models_host && labels_host && counts_out_host && models_out_host &&
	                  labels_out_host && pairs_per_rank >= 0 && n_points > 0 && model_size > 0 && max_models > 0,
	              "bad argument");
	PXB_TRY(require_nccl());
	PXB_CUDA(cudaSetDevice(ctx->device));
	int world = 0, rank = 0;
	PXB_NCCL(api().CommCount(nccl_comm, &world));
	PXB_NCCL(api().CommUserRank(nccl_comm, &rank));
	if (pairs_per_rank == 0) return PXB_OK;
	const size_t b_cnt = sizeof(int32_t) * (size_t)pairs_per_rank;
	const size_t b_mod = sizeof(double) * (size_t)pairs_per_rank * max_models * model_size;
	const size_t b_lab = sizeof(int32_t) * (size_t)pairs_per_rank * n_points;
	auto up16 = [](size_t v) { return (v + 15) & ~size_t(15); };
	const size_t rec = up16(b_mod) + up16(b_lab) + up16(b_cnt); // one rank's record, fields 16-byte aligned
	PXB_TRY(ctx->staging.reserve(rec * (size_t)world));
	char *all = ctx->staging.as<char>();
	char *mine = all + rec * (size_t)rank;
	PXB_CUDA(cudaMemcpyAsync(mine, models_host, b_mod, cudaMemcpyHostToDevice, ctx->stream));
	PXB_CUDA(cudaMemcpyAsync(mine + up16(b_mod), labels_host, b_lab, cudaMemcpyHostToDevice, ctx->stream));
	PXB_CUDA(cudaMemcpyAsync(mine + up16(b_mod) + up16(b_lab), counts_host, b_cnt, cudaMemcpyHostToDevice, ctx->stream));
	PXB_NCCL(api().AllGather(mine, all, rec, kNcclInt8, nccl_comm, ctx->stream));
	for (int r = 0; r < world; ++r) {
		const char *src = all + rec * (size_t)r;
		PXB_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(models_out_host) + b_mod * (size_t)r, src, b_mod, cudaMemcpyDeviceToHost,
		                         ctx->stream));
		PXB_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(labels_out_host) + b_lab * (size_t)r, src + up16(b_mod), b_lab,
		                         cudaMemcpyDeviceToHost, ctx->stream));
		PXB_CUDA(cudaMemcpyAsync(reinterpret_cast<char *>(counts_out_host) + b_cnt * (size_t)r, src + up16(b_mod) + up16(b_lab),
		                         b_cnt, cudaMemcpyDeviceToHost, ctx->stream));
	}
	PXB_TRY(ctx_wait(ctx));
	return PXB_OK;
}

} // extern "C"
