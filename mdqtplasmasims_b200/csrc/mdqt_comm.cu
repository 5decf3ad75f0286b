// mdqt_comm.cu -- row-decomposed large-N runs INSIDE the library (BASELINE configs[4]): rank g of G owns
// the ion rows [g N/G, (g+1) N/G) -- their forces, velocities, wavefunctions never leave the rank -- and the only data every
// rank needs from the others are positions, once per MD step. One process (or thread) per GPU, one NCCL communicator.
//
// Per MD step on the handle's stream S and the communication stream C:
//     S:  substep kernel (own rows)  ->  pack own fixed-point rows into the exchange buffer  -> event
//     C:  ncclAllGather of ONE [3][rows] int64 block per rank (in place, rank-major)  ->  unpack remote rows into Rfix  -> event
//     S:  force kernel over the j chunks that lie inside the rank's OWN rows (they need no remote data: overlaps C)
//     S:  wait for C, force kernel over the remaining j chunks, substep kernel, ...
// What travels is the 64-bit periodic FIXED-POINT copy of the positions -- exactly what the pair kernels read -- so no
// conversion pass follows the collective, and the chunk partial sums are added in ascending chunk order whatever launch
// produced them: forces are bitwise identical for any number of ranks.
// The observables of output() (SU:934-979) are completed by two small ncclAllReduce calls (mdqt_diagnostics / mdqt_vel_dist
// on a communicating handle return the whole-system values on every rank).
// NCCL is bound at run time (dlopen "libnccl.so.2"): single-GPU users need no NCCL, and a process that already carries NCCL
// (PyTorch) shares its copy.
#include "mdqt_handle.h"
#include <dlfcn.h>
#include <nccl.h>
#include <string.h>
#include <algorithm>

using namespace mdqt;

namespace {
struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  std::string error;
};
NcclApi* nccl_api() {
  static NcclApi api = [] {
    NcclApi a;
    for (const char* name : {"libnccl.so.2", "libnccl.so"}) {
      a.lib = dlopen(name, RTLD_NOW | RTLD_GLOBAL);
      if (a.lib) break;
    }
    if (!a.lib) { a.error = std::string("cannot load libnccl.so.2: ") + dlerror(); return a; }
#define BIND(field, sym) *(void**)(&a.field) = dlsym(a.lib, sym); if (!a.field) a.error = std::string("libnccl lacks ") + sym;
    BIND(GetUniqueId, "ncclGetUniqueId") BIND(CommInitRank, "ncclCommInitRank") BIND(CommDestroy, "ncclCommDestroy")
    BIND(AllGather, "ncclAllGather") BIND(AllReduce, "ncclAllReduce") BIND(GetErrorString, "ncclGetErrorString")
#undef BIND
    return a;
  }();
  return &api;
}
}  // namespace

struct mdqt_comm {
  ncclComm_t comm;
  int rank, world, rows;
  long long* xbuf;        // [world][3][rows] fixed-point positions, rank-major: the all-gather runs in place on it
  double* red;            // device scratch of the observable all-reduces
  cudaStream_t cstream;   // communication stream
  cudaEvent_t ev_packed, ev_unpacked;
  int c_lo, c_hi;         // j chunks [c_lo, c_hi) of the CTA-tile force plan lie inside the rank's own rows
  bool pending;           // an exchange is in flight on cstream: the next force call must wait for ev_unpacked
};

#define NC(call)                                                                                                   \
  do {                                                                                                             \
    ncclResult_t r_ = (call);                                                                                      \
    if (r_ != ncclSuccess) return mdqt_fail(MDQT_ECUDA, std::string(#call) + ": " + nccl_api()->GetErrorString(r_)); \
  } while (0)

__global__ void k_pack_rows(const long long* __restrict__ Rfix, long long* __restrict__ block, int row0, int rows, int own, int ld) {
  const long long g = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= 3LL * rows) return;
  const int c = (int)(g / rows), i = (int)(g % rows);
  block[g] = i < own ? Rfix[(size_t)c * ld + row0 + i] : 0;  // the last rank's block is padded
}
// Rfix[c][g*rows + i] = xbuf[g][c][i] for every rank g but `skip`
__global__ void k_unpack_rows(const long long* __restrict__ xbuf, long long* __restrict__ Rfix, int world, int rows, int N, int ld, int skip) {
  const long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= 3LL * rows * world) return;
  const int g = (int)(k / (3LL * rows));
  if (g == skip) return;
  const long long r = k % (3LL * rows);
  const int c = (int)(r / rows), i = (int)(r % rows);
  const long long row = (long long)g * rows + i;
  if (row < N) Rfix[(size_t)c * ld + row] = xbuf[k];
}

extern "C" {

int mdqt_comm_unique_id(void* id_out) {
  if (!id_out) return mdqt_fail(MDQT_EINVAL, "null argument");
  NcclApi* n = nccl_api();
  if (!n->error.empty()) return mdqt_fail(MDQT_ESTATE, n->error);
  ncclUniqueId id;
  NC(n->GetUniqueId(&id));
  memcpy(id_out, &id, sizeof(id));
  return MDQT_OK;
}

int mdqt_comm_init(mdqt_handle* h, const void* unique_id, int rank, int world) {
  if (!h || !unique_id) return mdqt_fail(MDQT_EINVAL, "null argument");
  if (h->comm) return mdqt_fail(MDQT_ESTATE, "handle already has a communicator");
  if (world < 1 || rank < 0 || rank >= world) return mdqt_fail(MDQT_EINVAL, "rank outside [0, world)");
  if (h->B != 1) return mdqt_fail(MDQT_ESTATE, "row decomposition needs n_traj == 1");
  const int rows = (h->N + world - 1) / world;  // rows per rank; the last rank holds the remainder
  if (h->row0 != rank * rows || h->nrows != std::min(rows, h->N - h->row0) || h->nrows < 1)
    return mdqt_fail(MDQT_EINVAL, "handle must own rows [rank * R, min(N, (rank+1) * R)) with R = ceil(N / world)");
  NcclApi* n = nccl_api();
  if (!n->error.empty()) return mdqt_fail(MDQT_ESTATE, n->error);
  CU(cudaSetDevice(h->p.device));
  mdqt_comm* c = new mdqt_comm();
  memset(c, 0, sizeof(*c));
  c->rank = rank; c->world = world; c->rows = rows;
  ncclUniqueId id;
  memcpy(&id, unique_id, sizeof(id));
  ncclResult_t r = n->CommInitRank(&c->comm, world, id, rank);
  if (r != ncclSuccess) { delete c; return mdqt_fail(MDQT_ECUDA, std::string("ncclCommInitRank: ") + n->GetErrorString(r)); }
  cudaError_t e = cudaMalloc((void**)&c->xbuf, sizeof(long long) * 3 * (size_t)c->rows * world);
  if (e == cudaSuccess) e = cudaMalloc((void**)&c->red, sizeof(double) * (16 + 3 * kVelBins));
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&c->cstream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_packed, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&c->ev_unpacked, cudaEventDisableTiming);
  if (e != cudaSuccess) {
    h->comm = c;
    mdqt_comm_release(h);
    return mdqt_fail(MDQT_ECUDA, std::string("communicator buffers: ") + cudaGetErrorString(e));
  }
  // chunks of the force plan that need own positions only (the CTA-tile kernel; the item kernel runs unsplit)
  c->c_lo = c->c_hi = 0;
  if (!h->items && h->nsplit > 1) {
    const int lo = (h->row0 + h->jlen - 1) / h->jlen;                              // first chunk starting inside the rows
    const int end = h->row0 + h->nrows;
    int hi = end / h->jlen;                                                        // chunks ending at or before the rows' end
    if (end == h->N) hi = h->nsplit;                                               // the last chunk ends at N
    if (hi > lo) { c->c_lo = lo; c->c_hi = hi; }
  }
  h->comm = c;
  return MDQT_OK;
}

int mdqt_comm_destroy(mdqt_handle* h) {
  if (!h) return mdqt_fail(MDQT_EINVAL, "null handle");
  if (h->stream) cudaStreamSynchronize(h->stream);
  mdqt_comm_release(h);
  return MDQT_OK;
}

// all-gather of the positions after the caller's own mdqt_substeps (the fused path is mdqt_md_steps): blocking on the stream
int mdqt_comm_exchange_positions(mdqt_handle* h);

}  // extern "C"

void mdqt_comm_release(mdqt_handle* h) {
  mdqt_comm* c = h->comm;
  if (!c) return;
  cudaSetDevice(h->p.device);
  if (c->cstream) cudaStreamSynchronize(c->cstream);
  if (c->comm) nccl_api()->CommDestroy(c->comm);
  if (c->xbuf) cudaFree(c->xbuf);
  if (c->red) cudaFree(c->red);
  if (c->ev_packed) cudaEventDestroy(c->ev_packed);
  if (c->ev_unpacked) cudaEventDestroy(c->ev_unpacked);
  if (c->cstream) cudaStreamDestroy(c->cstream);
  delete c;
  h->comm = nullptr;
}

// enqueue: pack own rows (stream S) -> all-gather + unpack (stream C). The consumer waits for ev_unpacked.
static int start_exchange(mdqt_handle* h) {
  mdqt_comm* c = h->comm;
  const long long n = 3LL * c->rows;
  long long* own = c->xbuf + (size_t)c->rank * 3 * c->rows;
  k_pack_rows<<<(unsigned)((n + 255) / 256), 256, 0, h->stream>>>(h->Rfix, own, h->row0, c->rows, h->nrows, h->ld);
  CU(cudaEventRecord(c->ev_packed, h->stream));
  CU(cudaStreamWaitEvent(c->cstream, c->ev_packed, 0));
  NC(nccl_api()->AllGather(own, c->xbuf, (size_t)n, ncclInt64, c->comm, c->cstream));
  const long long m = n * c->world;
  k_unpack_rows<<<(unsigned)((m + 255) / 256), 256, 0, c->cstream>>>(c->xbuf, h->Rfix, c->world, c->rows, h->N, h->ld, c->rank);
  CU(cudaEventRecord(c->ev_unpacked, c->cstream));
  c->pending = true;
  return MDQT_OK;
}
static int finish_exchange(mdqt_handle* h) {
  mdqt_comm* c = h->comm;
  if (!c->pending) return MDQT_OK;
  CU(cudaStreamWaitEvent(h->stream, c->ev_unpacked, 0));
  c->pending = false;
  return MDQT_OK;
}

extern "C" int mdqt_comm_exchange_positions(mdqt_handle* h) {
  if (!h || !h->comm) return mdqt_fail(MDQT_ESTATE, "handle has no communicator (mdqt_comm_init)");
  CU(cudaSetDevice(h->p.device));
  mdqt_refresh_fixed(h);  // own rows may have been uploaded
  int rc = start_exchange(h);
  if (rc) return rc;
  rc = finish_exchange(h);
  if (rc) return rc;
  CU(cudaGetLastError());
  return MDQT_OK;
}

// nsteps x { forces(); ratio x { step(); qstep(); } ; exchange } on a row-decomposed handle. Precondition (as for any force
// call): the handle's positions are complete -- after an upload of all N positions, or after the previous call's exchange.
int mdqt_comm_md_steps(mdqt_handle* h, int nsteps) {
  mdqt_comm* c = h->comm;
  const int ratio = h->p.substeps_per_md;
  for (int k = 0; k < nsteps; k++) {
    if (h->rfix_dirty) mdqt_refresh_fixed(h);  // after an upload (waits for an exchange in flight); otherwise keep the overlap
    ForceArgs fa = mdqt_force_args(h);
    const bool split = c->pending && c->c_hi > c->c_lo;
    if (split) {
      // Two concurrent launches of the same kernel: the chunks inside the own rows need no remote positions and start at once on
      // S; the other chunks are queued on C behind the all-gather and the unpack, and join the first launch on the SMs as soon as
      // the positions have landed -- the machine stays full (a serial "local chunks, wait, remote chunks" leaves it 5/6 empty
      // during the first part at 8 ranks: 10.4 instead of 9.9 ms per MD step at N = 2e5). The arrival counters and the
      // ascending-chunk reduction work across the two launches; S continues when both are done.
      ForceArgs la = fa;
      la.js0 = c->c_lo; la.js_count = c->c_hi - c->c_lo;
      launch_forces(la, h->stream, false);
      if (fa.nsplit > la.js_count) {
        ForceArgs ra = fa;  // chunks [0, c_lo) and [c_hi, nsplit)
        ra.js0 = 0; ra.js_skip0 = c->c_lo; ra.js_skipn = la.js_count; ra.js_count = fa.nsplit - la.js_count;
        launch_forces(ra, c->cstream, false);
        CU(cudaEventRecord(c->ev_unpacked, c->cstream));
      }
      int rc = finish_exchange(h);  // S waits for C: exchange, unpack and the remote-chunk launch
      if (rc) return rc;
    } else {
      int rc = finish_exchange(h);
      if (rc) return rc;
      launch_forces(fa, h->stream, false);
    }
    int rc = mdqt_enqueue_substeps(h, ratio, 1, 1, /*forces_partial=*/true);
    if (rc) return rc;
    if (c->world > 1) {
      rc = start_exchange(h);
      if (rc) return rc;
    }
  }
  CU(cudaGetLastError());
  return MDQT_OK;
}

// ---- observables of a row-decomposed run: partial sums over the own rows + ncclAllReduce --------------------------------------
extern "C" int mdqt_comm_allreduce(mdqt_handle* h, double* values, int n) {
  if (!h || !h->comm || !values) return mdqt_fail(MDQT_ESTATE, "handle has no communicator (mdqt_comm_init)");
  if (n < 1 || n > 16 + 3 * kVelBins) return mdqt_fail(MDQT_EINVAL, "mdqt_comm_allreduce: n outside [1, 6019]");
  mdqt_comm* c = h->comm;
  CU(cudaSetDevice(h->p.device));
  CU(cudaMemcpyAsync(c->red, values, sizeof(double) * n, cudaMemcpyHostToDevice, h->stream));
  NC(nccl_api()->AllReduce(c->red, c->red, (size_t)n, ncclDouble, ncclSum, c->comm, h->stream));
  CU(cudaMemcpyAsync(values, c->red, sizeof(double) * n, cudaMemcpyDeviceToHost, h->stream));
  CU(cudaStreamSynchronize(h->stream));
  return MDQT_OK;
}

int mdqt_comm_rank_world(const mdqt_handle* h, int* rank, int* world) {
  if (!h->comm) return 0;
  *rank = h->comm->rank; *world = h->comm->world;
  return 1;
}
int mdqt_comm_sync_pending(mdqt_handle* h) { return h->comm ? finish_exchange(h) : MDQT_OK; }
