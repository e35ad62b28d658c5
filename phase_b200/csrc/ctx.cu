// ctx.cu -- context, error text, NCCL bootstrap.
#include <cstdarg>
#include <dlfcn.h>

#include "comm.cuh"
#include "common.cuh"
#include "hostcopy.cuh"

namespace phb {
static thread_local char g_err[1024] = "";
void set_error(const char *fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
}  // namespace phb

extern "C" {

const char *phb_last_error(void) { return phb::g_err; }
int phb_version(void) { return 100; }

int phb_ctx_create(int device, phb_ctx **out) {
  PHB_REQUIRE(out, "phb_ctx_create: out is NULL");
  if (device == PHB_DEVICE_HOST_ONLY) {
    // mesh / partition / halo-map construction only (host logic); no kernels can run
    phb_ctx *c = new phb_ctx();
    c->device = -1;
    *out = c;
    return PHB_OK;
  }
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0) {
    phb::set_error("phb_ctx_create: no CUDA device (%s); this library has no CPU fallback",
                   cudaGetErrorString(e));
    return PHB_ERR_CUDA;
  }
  PHB_REQUIRE(device >= 0 && device < count, "phb_ctx_create: device %d out of range", device);
  PHB_CUDA(cudaSetDevice(device));
  phb_ctx *c = new phb_ctx();
  c->device = device;
  PHB_CUDA(cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking));
  PHB_CUDA(cudaStreamCreateWithFlags(&c->commStream, cudaStreamNonBlocking));
  PHB_CUDA(cudaDeviceGetAttribute(&c->numSMs, cudaDevAttrMultiProcessorCount, device));
  PHB_CUDA(cudaMallocHost((void **)&c->pinned, 256 * sizeof(double)));
  *out = c;
  return PHB_OK;
}

int phb_ctx_destroy(phb_ctx *c) {
  if (!c) return PHB_OK;
  if (c->device < 0) {
    delete c;
    return PHB_OK;
  }
  cudaSetDevice(c->device);
  // a captured graph with NCCL nodes keeps the communicator alive: ncclCommDestroy would wait for it forever
  for (phb_solver *s : c->solvers) phb::solver_drop_graph(s);
  cudaStreamSynchronize(c->stream);
  phb::peer_destroy(c);
  phb::comm_destroy(c);
  if (c->stream) cudaStreamDestroy(c->stream);
  if (c->commStream) cudaStreamDestroy(c->commStream);
  if (c->pinned) cudaFreeHost(c->pinned);
  if (c->stage) { c->stage->release(); delete c->stage; }
  delete c;
  return PHB_OK;
}

int phb_ctx_rank(const phb_ctx *c) { return c ? c->rank : 0; }
int phb_ctx_nprocs(const phb_ctx *c) { return c ? c->nProcs : 1; }
long long phb_ctx_kernel_launches(const phb_ctx *c) { return c ? c->launches : 0; }
void *phb_ctx_stream(phb_ctx *c) { return c ? (void *)c->stream : nullptr; }

int phb_ctx_sync(phb_ctx *c) {
  PHB_REQUIRE(c, "phb_ctx_sync: ctx is NULL");
  if (c->device < 0) return PHB_OK;
  PHB_CUDA(cudaStreamSynchronize(c->stream));
  PHB_CUDA(cudaStreamSynchronize(c->commStream));
  return phb::launch_status(c);
}

int phb_comm_unique_id(void *out128) { return phb::comm_unique_id(out128); }
int phb_ctx_init_comm(phb_ctx *c, int rank, int nProcs, const void *id128) {
  PHB_REQUIRE(c, "phb_ctx_init_comm: ctx is NULL");
  if (c->device < 0) {  // host-only: just record the rank layout
    PHB_REQUIRE(nProcs >= 1 && rank >= 0 && rank < nProcs, "phb_ctx_init_comm: bad rank %d/%d", rank, nProcs);
    c->rank = rank;
    c->nProcs = nProcs;
    return PHB_OK;
  }
  return phb::comm_init(c, rank, nProcs, id128);
}

int phb_ctx_peer_arena_create(phb_ctx *c, long long maxCols, int maxSolvers, void *handle64) {
  return phb::peer_arena_create(c, maxCols, maxSolvers, handle64);
}
int phb_ctx_peer_arena_open(phb_ctx *c, const void *handles) { return phb::peer_arena_open(c, handles); }

}  // extern "C"
