// comm_host.inl -- the global mass-balance / convergence reductions behind the C ABI.
//
// The reference keeps its mass bookkeeping per MPI rank (MPPVSFMALM_Driver.F90:124-133, 556-601, 845-898); a batch sharded by
// column over the GPUs of one box wants it globally.  Every StepDT leaves 9 doubles (4 sums, 4 maxima, the worst SNES reason)
// in the handle's device buffer; mppgpu_global_reduce_async gathers the buffers of all ranks with ONE ncclAllGather on the
// handle's stream and folds them on the device in rank order (deterministic).  NCCL is resolved with dlopen at
// mppgpu_comm_init time: the library has no link-time dependency on it and binds to the copy the host process already
// loaded (torch's, an MPI launcher's) when there is one.  A host model in Fortran + MPI gets the unique id from
// mppgpu_comm_unique_id on rank 0, MPI_Bcast's the 128 bytes, and calls mppgpu_comm_init on every rank.
#include <dlfcn.h>
#include <nccl.h>          // types and enums only; no NCCL symbol is referenced at link time

struct NcclApi {
  void *lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId *) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void *, void *, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char *(*GetErrorString)(ncclResult_t) = nullptr;
};
static NcclApi g_nccl;

static int nccl_load()
{
  if (g_nccl.lib) return 0;
  const char *names[] = {"libnccl.so.2", "libnccl.so"};
  void *lib = nullptr;
  for (const char *n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
  if (!lib) return fail("mppgpu_comm: cannot load libnccl.so.2 (%s)", dlerror());
  NcclApi a; a.lib = lib;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
  a.AllGather = (decltype(a.AllGather))dlsym(lib, "ncclAllGather");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
  if (!a.GetUniqueId || !a.CommInitRank || !a.AllGather || !a.CommDestroy || !a.GetErrorString) return fail("mppgpu_comm: libnccl lacks a required symbol");
  g_nccl = a;
  return 0;
}
#define NK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) return fail("%s:%d NCCL error %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); } while (0)

struct CommState {
  ncclComm_t comm = nullptr; int nranks = 1, rank = 0;
  DevBuf<double> gathered, result;       // nranks x 9, 9
  double *h_result = nullptr;            // pinned
};

// rank-ordered fold of the gathered per-rank buffers: sums add, maxima take the max, the SNES reason the minimum
__global__ void fold_reductions_kernel(const double *__restrict__ g, int nranks, double *__restrict__ out)
{
  const int k = threadIdx.x;
  if (k >= 9) return;
  double v = g[k];
  for (int r = 1; r < nranks; ++r) {
    const double x = g[(size_t)r * 9 + k];
    v = (k < 4) ? v + x : (k < 8 ? fmax(v, x) : fmin(v, x));
  }
  out[k] = v;
}

static void comm_destroy(CommState *c)
{
  if (!c) return;
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  if (c->h_result) cudaFreeHost(c->h_result);
  delete c;
}

extern "C" int mppgpu_comm_unique_id(void *id128)
{
  if (!id128) return fail("mppgpu_comm_unique_id: null buffer");
  if (nccl_load()) return 1;
  static_assert(sizeof(ncclUniqueId) == MPPGPU_COMM_ID_BYTES, "ncclUniqueId size");
  ncclUniqueId id;
  NK(g_nccl.GetUniqueId(&id));
  memcpy(id128, &id, sizeof(id));
  return 0;
}

extern "C" int mppgpu_comm_init(mppgpu_handle h, int nranks, int rank, const void *id128)
{
  CHECK_H(h);
  if (nranks < 1 || rank < 0 || rank >= nranks) return fail("mppgpu_comm_init: rank %d outside a world of %d", rank, nranks);
  if (h->comm) { comm_destroy(h->comm); h->comm = nullptr; }
  CommState *c = new CommState();
  c->nranks = nranks; c->rank = rank;
  h->comm = c;                                       // owned by the handle from here on (mppgpu_destroy frees it)
  CK(c->gathered.alloc((size_t)nranks * 9)); CK(c->result.alloc(9));
  CK(cudaMemsetAsync(c->result.p, 0, 9 * sizeof(double), h->stream));
  CK(cudaMallocHost((void **)&c->h_result, 9 * sizeof(double)));
  memset(c->h_result, 0, 9 * sizeof(double));
  if (nranks > 1) {
    if (!id128) return fail("mppgpu_comm_init: a unique id is required for more than one rank");
    if (nccl_load()) return 1;
    ncclUniqueId id; memcpy(&id, id128, sizeof(id));
    NK(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
  }
  return 0;
}

// queue: all-gather of this step's 9 reduction doubles + fold + copy to the pinned result (asynchronous on the handle's stream)
extern "C" int mppgpu_global_reduce_async(mppgpu_handle h)
{
  CHECK_H(h);
  if (!h->comm) return fail("mppgpu_global_reduce_async: call mppgpu_comm_init first (nranks = 1 needs no unique id)");
  CommState *c = h->comm;
  if (c->nranks > 1) {
    NK(g_nccl.AllGather(h->red_out.p, c->gathered.p, 9, ncclDouble, c->comm, h->stream));
    fold_reductions_kernel<<<1, 32, 0, h->stream>>>(c->gathered.p, c->nranks, c->result.p);
    CK(cudaGetLastError());
    h->launches += 1;
  } else {
    CK(cudaMemcpyAsync(c->result.p, h->red_out.p, 9 * sizeof(double), cudaMemcpyDeviceToDevice, h->stream));
  }
  CK(cudaMemcpyAsync(c->h_result, c->result.p, 9 * sizeof(double), cudaMemcpyDeviceToHost, h->stream));
  c_pending(h) = true;
  return 0;
}

// collective: every rank calls it after its StepDT.  sums: mass before, mass after, sources * dt, boundary mass exchanged [kg];
// maxs: |mass error| [kg], Newton iterations, any column diverged (0/1), dt cuts; worst_reason: the minimum SNESConvergedReason.
extern "C" int mppgpu_global_mass_balance(mppgpu_handle h, double sums[4], double maxs[4], int *worst_reason)
{
  CHECK_H(h);
  if (!h->comm) return fail("mppgpu_global_mass_balance: call mppgpu_comm_init first (nranks = 1 needs no unique id)");
  if (!c_pending(h) && mppgpu_global_reduce_async(h)) return 1;
  CK(cudaStreamSynchronize(h->stream));
  c_pending(h) = false;
  const double *r = h->comm->h_result;
  for (int k = 0; k < 4; ++k) { if (sums) sums[k] = r[k]; if (maxs) maxs[k] = r[4 + k]; }
  if (worst_reason) *worst_reason = (int)r[8];
  return 0;
}
