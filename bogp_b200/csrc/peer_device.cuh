// Device side of the peer-memory record exchange (protocol: kernels_peer.cu), shared by the stand-alone exchange kernel
// and the surface kernel that runs the exchange from its last CTA.
#pragma once
#include <cuda_runtime.h>

namespace bogp {

// A record slot is four 8-byte words, each (32 payload bits | 32-bit epoch tag << 32): index low / high, value low / high.
// An aligned 8-byte store is single-copy atomic, so every word validates itself (the scheme of NCCL's LL protocol): the
// writer issues four plain stores and is done -- no fence, no flag, no round trip over NVLink before the kernel can end --
// and the reader polls until all four words carry the epoch it waits for.  Slot (e & 1) cannot be overwritten by epoch
// e + 2 before it has been read (kernels_peer.cu), so a tag is either e - 2 (stale) or e.
constexpr int kPeerSlotBytes = 32;

__device__ __forceinline__ void peer_wait_read(const unsigned char* mine, int t, int world, unsigned long long epoch,
                                               long long* recs_out, int* err) {
  const unsigned long long* src = reinterpret_cast<const unsigned long long*>(mine + ((size_t)(epoch & 1ull) * world + t) * kPeerSlotBytes);
  const unsigned int tag = (unsigned int)epoch;
  unsigned long long w0, w1, w2, w3;
  const long long t0 = clock64();
  for (unsigned spin = 0;; ++spin) {
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w0) : "l"(src) : "memory");
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w1) : "l"(src + 1) : "memory");
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w2) : "l"(src + 2) : "memory");
    asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(w3) : "l"(src + 3) : "memory");
    if ((unsigned int)(w0 >> 32) == tag && (unsigned int)(w1 >> 32) == tag && (unsigned int)(w2 >> 32) == tag && (unsigned int)(w3 >> 32) == tag) break;
    if ((spin & 1023u) == 1023u && clock64() - t0 > 20000000000LL) {      // ~10 s: a peer never reached this step
      atomicExch(err, 1 + t);
      __threadfence_system();
      __trap();                                                               // surfaces as a CUDA error at the next sync
    }
  }
  recs_out[2 * t] = (long long)((w0 & 0xffffffffull) | (w1 << 32));
  recs_out[2 * t + 1] = (long long)((w2 & 0xffffffffull) | (w3 << 32));
}

__device__ __forceinline__ void peer_publish(unsigned char* const* bufs, int t, int rank, int world, unsigned long long epoch,
                                             long long r0, long long r1) {
  unsigned long long* dst = reinterpret_cast<unsigned long long*>(bufs[t] + ((size_t)(epoch & 1ull) * world + rank) * kPeerSlotBytes);
  const unsigned long long tag = (unsigned long long)(unsigned int)epoch << 32;
  const unsigned long long a = (unsigned long long)r0, b = (unsigned long long)r1;
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst), "l"((a & 0xffffffffull) | tag) : "memory");
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 1), "l"((a >> 32) | tag) : "memory");
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 2), "l"((b & 0xffffffffull) | tag) : "memory");
  asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(dst + 3), "l"((b >> 32) | tag) : "memory");
}

// One step of the exchange for peer `t` (one thread per peer).  mode 0: collective (publish, wait for the peers' records
// of the same epoch).  mode 1: pipelined (records of epoch - 1, then publish; nobody is waited for).  mode 2: a collective
// step right after a pipelined one (the peers may still be reading the slot of epoch - 2 until they have published
// epoch - 1, so that is waited for before the slot is overwritten).
__device__ __forceinline__ void peer_step(unsigned char* const* bufs, int t, int rank, int world, unsigned long long epoch, int mode,
                                          long long r0, long long r1, long long* recs_out, int* err) {
  if (mode) {
    if (epoch > 1) peer_wait_read(bufs[rank], t, world, epoch - 1, recs_out, err);
    else { recs_out[2 * t] = -1; recs_out[2 * t + 1] = 0; }
  }
  peer_publish(bufs, t, rank, world, epoch, r0, r1);
  if (mode != 1) peer_wait_read(bufs[rank], t, world, epoch, recs_out, err);
}

}  // namespace bogp
