// (e) multi-GPU shards behind the C ABI: one process + one hb_ctx per GPU, instances split in contiguous blocks, NO collective on
// the data path -- the only exchange is the gather of per-instance output rows (80-byte torque rows in the control step), issued on the
// shard's own stream behind an event on the context's stream so that it overlaps the next step's kernels.
// NCCL is resolved at run time (dlopen of the libnccl.so.2 the process already uses, or the system one): the library itself does not
// link against it, a single-GPU caller never loads it.
// Included at the end of hb_api.cu (same translation unit: hb_ctx, error codes).
#pragma once
#include <dlfcn.h>
#include <nccl.h>

namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
};

NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, []() {
    // RTLD_NOLOAD first: a process that already carries an NCCL (torch bundles one) must keep using that copy
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
    if (!h) return;
    api.handle = h;
    api.GetUniqueId = reinterpret_cast<decltype(api.GetUniqueId)>(dlsym(h, "ncclGetUniqueId"));
    api.CommInitRank = reinterpret_cast<decltype(api.CommInitRank)>(dlsym(h, "ncclCommInitRank"));
    api.CommDestroy = reinterpret_cast<decltype(api.CommDestroy)>(dlsym(h, "ncclCommDestroy"));
    api.AllGather = reinterpret_cast<decltype(api.AllGather)>(dlsym(h, "ncclAllGather"));
    api.GetErrorString = reinterpret_cast<decltype(api.GetErrorString)>(dlsym(h, "ncclGetErrorString"));
    api.ok = api.GetUniqueId && api.CommInitRank && api.CommDestroy && api.AllGather;
  });
  return api;
}

// block of `rank`: sizes differ by at most one (the first total % world ranks own one more)
__host__ __device__ inline void shard_block(int total, int world, int rank, int* lo, int* n) {
  const int base = total / world, rem = total % world;
  *lo = rank * base + (rank < rem ? rank : rem);
  *n = base + (rank < rem ? 1 : 0);
}

// rows of the local block into the (padded) send buffer, undoing a schedule sort on the way: out[i] = rows[inverse[i]]
__global__ void shard_pack_kernel(int n, int row, const double* __restrict__ rows, const int32_t* __restrict__ inverse, double* __restrict__ send) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n * row) return;
  const int r = i / row, c = i - r * row;
  const int src = inverse ? inverse[r] : r;
  send[i] = rows[(size_t)src * row + c];
}

// world x nmax padded blocks -> total rows in instance order
__global__ void shard_compact_kernel(int total, int world, int nmax, int row, const double* __restrict__ recv, double* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total * row) return;
  const int g = i / row, c = i - g * row;
  const int base = total / world, rem = total % world;
  // owner of global row g
  int rank = (g < rem * (base + 1)) ? g / (base + 1) : rem + (base ? (g - rem * (base + 1)) / base : 0);
  int lo, n;
  shard_block(total, world, rank, &lo, &n);
  out[i] = recv[((size_t)rank * nmax + (g - lo)) * row + c];
}

}  // namespace

struct hb_shard {
  hb_ctx* ctx = nullptr;
  int device = 0;                          // kept separately: hb_shard_destroy must not touch a context that was destroyed first
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0, total = 0, lo = 0, n = 0, nmax = 0, max_row = 0;
  cudaStream_t stream = nullptr;           // the gather runs here
  cudaEvent_t ready = nullptr, done = nullptr;
  double* send[2] = {nullptr, nullptr};    // double buffered: the gather of step k overlaps the kernels of step k+1
  double* recv = nullptr;
  double* out[2] = {nullptr, nullptr};
  int turn = 0;
  int last_nccl = 0;
};

extern "C" {

int hb_shard_partition(int total, int world, int rank, int* begin, int* count) {
  if (total < 0 || world < 1 || rank < 0 || rank >= world || !begin || !count) return HB_EINVAL;
  shard_block(total, world, rank, begin, count);
  return HB_OK;
}

int hb_shard_sort_by_schedule(int B, int nodes, const int32_t* mode, int32_t* perm, int32_t* inverse) {
  if (B < 0 || nodes < 1 || !mode || !perm) return HB_EINVAL;
  for (int i = 0; i < B; ++i) perm[i] = i;
  std::stable_sort(perm, perm + B, [&](int32_t a, int32_t b) {
    return std::lexicographical_compare(mode + (size_t)a * nodes, mode + (size_t)(a + 1) * nodes, mode + (size_t)b * nodes, mode + (size_t)(b + 1) * nodes);
  });
  if (inverse) for (int i = 0; i < B; ++i) inverse[perm[i]] = i;
  return HB_OK;
}

int hb_shard_unique_id(void* id) {
  if (!id) return HB_EINVAL;
  NcclApi& api = nccl_api();
  if (!api.ok) return HB_ECOMM;
  static_assert(sizeof(ncclUniqueId) == HB_SHARD_ID_BYTES, "ncclUniqueId is 128 bytes");
  ncclUniqueId uid;
  if (api.GetUniqueId(&uid) != ncclSuccess) return HB_ECOMM;
  memcpy(id, &uid, sizeof(uid));
  return HB_OK;
}

int hb_shard_destroy(hb_shard* s) {
  if (!s) return HB_OK;
  cudaSetDevice(s->device);
  if (s->stream) cudaStreamSynchronize(s->stream);
  if (s->comm) nccl_api().CommDestroy(s->comm);
  for (int i = 0; i < 2; ++i) { if (s->send[i]) cudaFree(s->send[i]); if (s->out[i]) cudaFree(s->out[i]); }
  if (s->recv) cudaFree(s->recv);
  if (s->ready) cudaEventDestroy(s->ready);
  if (s->done) cudaEventDestroy(s->done);
  if (s->stream) cudaStreamDestroy(s->stream);
  delete s;
  return HB_OK;
}

int hb_shard_create(hb_ctx* ctx, const void* id, int world, int rank, int total_instances, int max_row_doubles, hb_shard** out) {
  if (!ctx || !out || world < 1 || rank < 0 || rank >= world || total_instances < 1 || max_row_doubles < 1 || (world > 1 && !id)) return HB_EINVAL;
  if (set_device(ctx)) return HB_ECUDA;
  hb_shard* s = new (std::nothrow) hb_shard();
  if (!s) return HB_ENOMEM;
  s->ctx = ctx; s->device = ctx->device; s->world = world; s->rank = rank; s->total = total_instances; s->max_row = max_row_doubles;
  shard_block(total_instances, world, rank, &s->lo, &s->n);
  s->nmax = (total_instances + world - 1) / world;
  bool ok = cudaStreamCreateWithFlags(&s->stream, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&s->ready, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&s->done, cudaEventDisableTiming) == cudaSuccess;
  const size_t row = (size_t)max_row_doubles * sizeof(double);
  for (int i = 0; ok && i < 2; ++i)
    ok = cudaMalloc(&s->send[i], row * s->nmax) == cudaSuccess && cudaMemset(s->send[i], 0, row * s->nmax) == cudaSuccess &&
         cudaMalloc(&s->out[i], row * total_instances) == cudaSuccess;
  ok = ok && cudaMalloc(&s->recv, row * s->nmax * world) == cudaSuccess;
  if (!ok) { hb_shard_destroy(s); return HB_ECUDA; }
  if (world > 1) {
    NcclApi& api = nccl_api();
    if (!api.ok) { hb_shard_destroy(s); return HB_ECOMM; }
    ncclUniqueId uid;
    memcpy(&uid, id, sizeof(uid));
    const ncclResult_t r = api.CommInitRank(&s->comm, world, uid, rank);
    if (r != ncclSuccess) { s->comm = nullptr; hb_shard_destroy(s); return HB_ECOMM; }
  }
  *out = s;
  return HB_OK;
}

int hb_shard_block(const hb_shard* s, int* begin, int* count) {
  if (!s || !begin || !count) return HB_EINVAL;
  *begin = s->lo; *count = s->n;
  return HB_OK;
}

int hb_shard_gather_dev(hb_shard* s, int row_doubles, const double* rows_dev, const int32_t* inverse_dev, const double** gathered_dev) {
  if (!s || row_doubles < 1 || row_doubles > s->max_row || !rows_dev || !gathered_dev) return HB_EINVAL;
  hb_ctx* ctx = s->ctx;
  if (set_device(ctx)) return HB_ECUDA;
  const int b = s->turn & 1;
  s->turn++;
  // the buffers of this turn were last touched two calls ago: their gather has completed before the previous call's `done` was recorded,
  // and both calls ran on s->stream in order, so reusing them behind s->stream is safe; the pack runs on the CONTEXT's stream (it reads
  // the solver's output), so it must not overtake the gather that read send[b] two calls ago
  CK(cudaStreamWaitEvent(ctx->stream, s->done, 0));
  if (s->n > 0) {
    const int work = s->n * row_doubles;
    shard_pack_kernel<<<(work + 255) / 256, 256, 0, ctx->stream>>>(s->n, row_doubles, rows_dev, inverse_dev, s->send[b]);
    ctx->launches++;
  }
  CK(cudaEventRecord(s->ready, ctx->stream));
  CK(cudaStreamWaitEvent(s->stream, s->ready, 0));
  if (s->world > 1) {
    const ncclResult_t r = nccl_api().AllGather(s->send[b], s->recv, (size_t)s->nmax * row_doubles, ncclDouble, s->comm, s->stream);
    if (r != ncclSuccess) { s->last_nccl = (int)r; return HB_ECOMM; }
  } else {
    CK(cudaMemcpyAsync(s->recv, s->send[b], sizeof(double) * s->nmax * row_doubles, cudaMemcpyDeviceToDevice, s->stream));
  }
  const int work = s->total * row_doubles;
  shard_compact_kernel<<<(work + 255) / 256, 256, 0, s->stream>>>(s->total, s->world, s->nmax, row_doubles, s->recv, s->out[b]);
  ctx->launches++;
  CK(cudaGetLastError());
  CK(cudaEventRecord(s->done, s->stream));
  *gathered_dev = s->out[b];
  return HB_OK;
}

int hb_shard_wait(hb_shard* s, int block_host) {
  if (!s) return HB_EINVAL;
  hb_ctx* ctx = s->ctx;
  if (set_device(ctx)) return HB_ECUDA;
  CK(cudaStreamWaitEvent(ctx->stream, s->done, 0));
  if (block_host) CK(cudaStreamSynchronize(s->stream));
  return HB_OK;
}

const char* hb_shard_last_error(const hb_shard* s) {
  if (!s || !s->last_nccl) return "";
  NcclApi& api = nccl_api();
  return api.GetErrorString ? api.GetErrorString((ncclResult_t)s->last_nccl) : "NCCL error";
}

}  // extern "C"
