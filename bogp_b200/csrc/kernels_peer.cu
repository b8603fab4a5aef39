// Final reduction across the GPUs of one box without a collective library: every rank owns a small device buffer that
// its peers map through CUDA IPC; one tiny kernel per step stores this rank's 16-byte (index, value) record straight
// into every peer's buffer over NVLink (P2P stores, every word tagged with the step's epoch) and waits for the peers'
// records of the same step -- or, pipelined, picks up those of the step before.  This is
// the exchange the reference does on the host with np.argmax over the collated surfaces
// (ref/projects/swellex96_localization/data/collate.py:50); an NCCL all-gather of 16 bytes per rank costs ~25 us per
// step on 4-8 GPUs, a fixed cost that is 12 % of the 0.19 ms step.
//
// Buffer layout per rank: records[2][world] slots of 32 bytes, double-buffered by epoch parity; every 8-byte word of a
// slot carries the epoch it was written in (peer_device.cuh), so there are no flags and no fences.  Rank A writes
// records[e & 1][A] in every peer; a peer reads records[e & 1][*] once all words carry e.  A rank can only start epoch
// e + 2 (same parity) after its epoch e + 1 kernel has finished, i.e. after every peer has published e + 1 -- which a
// peer does only after its epoch-e kernel (the reader of the slot that is about to be overwritten) has completed.
#include <cuda_runtime.h>

#include <cstring>
#include <vector>

#include "bogp_internal.h"
#include "peer_device.cuh"


namespace bogp {

constexpr int kPeerMaxWorld = 64;

// One thread per peer.  Collective form (pipelined == 0): publish this rank's record of `epoch`, wait for every peer's
// record of the same epoch.  Pipelined form: first pick up the peers' records of epoch - 1 (published one whole step
// ago, so the wait is normally over before it starts), then publish this rank's record of `epoch` without waiting for
// anybody -- the step no longer ends on the slowest rank of the box; the records of the last step are picked up by
// peer_collect_kernel.  Collect-before-publish keeps the two record slots safe: a peer can only publish epoch e + 1
// (the slot of e - 1) after it has seen this rank's epoch e, which is stored after this rank has read e - 1.
__global__ void peer_exchange_kernel(unsigned char* const* bufs, int rank, int world, unsigned long long epoch, int pipelined,
                                     const long long* rec_in, long long* recs_out, int* err) {
  const int t = threadIdx.x;
  if (t >= world) return;
  peer_step(bufs, t, rank, world, epoch, pipelined, rec_in[0], rec_in[1], recs_out, err);
}

// the records of the latest published epoch (closes a pipelined sequence)
__global__ void peer_collect_kernel(unsigned char* const* bufs, int rank, int world, unsigned long long epoch,
                                    long long* recs_out, int* err) {
  const int t = threadIdx.x;
  if (t >= world) return;
  if (epoch == 0) { recs_out[2 * t] = -1; recs_out[2 * t + 1] = 0; return; }
  peer_wait_read(bufs[rank], t, world, epoch, recs_out, err);
}

int peer_next(bogp_peer_s* p, long long index_offset, long long* recs_out, PeerArgs& a) {
  BOGP_REQUIRE(p && p->mapped_dev, "peer exchange: call bogp_peer_open first");
  ++p->epoch;
  a.bufs = p->mapped_dev; a.rank = p->rank; a.world = p->world; a.epoch = p->epoch; a.index_offset = index_offset;
  a.recs_out = recs_out; a.pipelined = p->pipelined ? 1 : (p->last_pipelined ? 2 : 0); a.err = reinterpret_cast<int*>(p->local + p->bytes - 64);
  p->last_pipelined = p->pipelined;
  return 0;
}

int peer_exchange_launch(const PeerArgs& a, const long long* rec_dev, cudaStream_t s) {
  peer_exchange_kernel<<<1, 64, 0, s>>>(a.bufs, a.rank, a.world, a.epoch, a.pipelined, rec_dev, a.recs_out, a.err);
  count_launch();
  BOGP_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace bogp

extern "C" int bogp_peer_create(bogp_peer_t* out, int rank, int world, unsigned char* ipc_handle_out) {
  BOGP_REQUIRE(out && ipc_handle_out, "bogp_peer_create: NULL argument");
  *out = nullptr;
  BOGP_REQUIRE(world >= 1 && world <= bogp::kPeerMaxWorld && rank >= 0 && rank < world, "bogp_peer_create: bad rank %d / world %d", rank, world);
  static_assert(sizeof(cudaIpcMemHandle_t) == BOGP_IPC_HANDLE_BYTES, "IPC handle size");
  bogp_peer_s* p = new bogp_peer_s();
  p->rank = rank; p->world = world;
  p->bytes = (size_t)2 * world * bogp::kPeerSlotBytes + 64;
  if (cudaMalloc(&p->local, p->bytes) != cudaSuccess || cudaMemset(p->local, 0, p->bytes) != cudaSuccess) {
    bogp::set_error("bogp_peer_create: %s", cudaGetErrorString(cudaGetLastError()));
    delete p;
    return BOGP_ERR_CUDA;
  }
  cudaIpcMemHandle_t h;
  if (cudaIpcGetMemHandle(&h, p->local) != cudaSuccess) {
    bogp::set_error("bogp_peer_create: cudaIpcGetMemHandle: %s", cudaGetErrorString(cudaGetLastError()));
    cudaFree(p->local);
    delete p;
    return BOGP_ERR_CUDA;
  }
  std::memcpy(ipc_handle_out, &h, sizeof(h));
  BOGP_CUDA(cudaDeviceSynchronize());
  *out = p;
  return 0;
}

extern "C" int bogp_peer_open(bogp_peer_t p, const unsigned char* all_handles) {
  BOGP_REQUIRE(p && all_handles, "bogp_peer_open: NULL argument");
  BOGP_REQUIRE(p->mapped.empty(), "bogp_peer_open: already open");
  p->mapped.assign(p->world, nullptr);
  for (int r = 0; r < p->world; ++r) {
    if (r == p->rank) { p->mapped[r] = p->local; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, all_handles + (size_t)r * sizeof(h), sizeof(h));
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      bogp::set_error("bogp_peer_open: cudaIpcOpenMemHandle(rank %d): %s", r, cudaGetErrorString(e));
      cudaGetLastError();
      return BOGP_ERR_CUDA;
    }
    p->mapped[r] = static_cast<unsigned char*>(ptr);
  }
  BOGP_CUDA(cudaMalloc(&p->mapped_dev, p->world * sizeof(unsigned char*)));
  BOGP_CUDA(cudaMemcpy(p->mapped_dev, p->mapped.data(), p->world * sizeof(unsigned char*), cudaMemcpyHostToDevice));
  return 0;
}

extern "C" int bogp_peer_exchange(bogp_peer_t p, const long long* rec_dev, long long* recs_out_dev, void* stream) {
  BOGP_REQUIRE(p && rec_dev && recs_out_dev, "bogp_peer_exchange: NULL argument");
  BOGP_REQUIRE(p->mapped_dev, "bogp_peer_exchange: call bogp_peer_open first");
  bogp::PeerArgs a;
  if (int rc = bogp::peer_next(p, 0, recs_out_dev, a)) return rc;
  return bogp::peer_exchange_launch(a, rec_dev, static_cast<cudaStream_t>(stream));
}

extern "C" int bogp_peer_set_pipelined(bogp_peer_t p, int on) {
  BOGP_REQUIRE(p, "bogp_peer_set_pipelined: NULL handle");
  p->pipelined = on ? 1 : 0;
  return 0;
}

extern "C" int bogp_peer_collect(bogp_peer_t p, long long* recs_out_dev, void* stream) {
  BOGP_REQUIRE(p && recs_out_dev, "bogp_peer_collect: NULL argument");
  BOGP_REQUIRE(p->mapped_dev, "bogp_peer_collect: call bogp_peer_open first");
  bogp::peer_collect_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(p->mapped_dev, p->rank, p->world, p->epoch, recs_out_dev,
                                                                            reinterpret_cast<int*>(p->local + p->bytes - 64));
  bogp::count_launch();
  BOGP_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int bogp_peer_destroy(bogp_peer_t p) {
  if (!p) return 0;
  cudaDeviceSynchronize();
  for (int r = 0; r < (int)p->mapped.size(); ++r)
    if (r != p->rank && p->mapped[r]) cudaIpcCloseMemHandle(p->mapped[r]);
  if (p->mapped_dev) cudaFree(p->mapped_dev);
  if (p->local) cudaFree(p->local);
  delete p;
  return 0;
}
