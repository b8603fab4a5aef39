// K1-UMMA-H: the separable (range x depth) Bartlett surface on the tcgen05 tensor cores with FP16-split operands.
//
// Same mathematics as kernels_umma.cu (mode-space operator, D[i, row] = sum_m E_f^T[i, m] * Bop[row, m], fused Bartlett
// epilogue), different arithmetic and pipeline:
//   * Operands are split x * 2^s = hi + lo into two FP16 numbers (s: a per-tonal power of two that puts the largest
//     magnitude at 2^14, so hi and lo carry 11 significant bits each down to an absolute floor of 2^-38 of the largest
//     value) and D = A_lo B_hi + A_hi B_lo + A_hi B_hi accumulates in fp32 in TMEM: the same three products and the same
//     22-bit operand precision as the 3xTF32 split (both drop the lo * lo term, 2^-22), at the kind::f16 rate (K = 16
//     per instruction: half the MMA instructions and half the operand bytes of kind::tf32).  The Bartlett value is a
//     ratio of two quadratic forms of the same scaled D, so the scales cancel.
//   * A = E_f^T comes from SHARED memory (one bulk copy per (range tile, tonal) of a table kept across calls for the
//     same range vector), stages are N = 128 wide (8 depths of a 16-row tonal, 4 of a 32-row tonal): at N = 128 an MMA
//     with A in shared memory runs at the full N/2 cycles (tools/micro/tc_peak.cu), the whole of TMEM is accumulators
//     (2 buffers x (re | im) x 128 columns), and a stage carries twice the work of the TF32 kernel's N = 64 stage per
//     barrier hand-shake -- the small tonals of that kernel were bound by the hand-shakes, not by the tensor pipe.
//   * All 16 epilogue warps drain EVERY stage (4 per TMEM lane quadrant, depth j -> warp j % 4): a buffer is held for one
//     TMEM round trip, and since a depth always belongs to the same warp the multi-frequency accumulation needs no
//     ordering between warps and the write-out no block barrier.
// Replaces the loop `for z in src_z: amb[zz,:] = MFP({"src_z": z, "rec_r": rvec})`
// (ref/projects/swellex96_localization/data/mfp.py:213-214) for a vertical array.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>

#include <cuda_fp16.h>

#include "bogp_internal.h"
#include "device_math.cuh"
#include "peer_device.cuh"
#include "umma_ptx.cuh"

// Per-tonal time stamps of CTA 0 (issuer 0 and epilogue warp 0) are compiled in only with -DBOGP_UMMA_DBG=1
// (BOGP_UMMA_DBG_BUILD=1 python -m bogp_b200.build) and printed when BOGP_UMMA_DBG=1 is set at run time.
#ifndef BOGP_UMMA_DBG
#define BOGP_UMMA_DBG 0
#endif

namespace bogp {
namespace {

constexpr int kTile = 128;              // ranges per tile = MMA M = TMEM lanes
constexpr int kNA = 128;                // accumulator columns per part; buffer b = columns [256 b, 256 b + 256): re | im
constexpr int kMaxStages = 6;           // Bop ring depth (as many as fit)
constexpr int kEpiW = 16;               // epilogue warps: 4 TMEM lane quadrants x 4 depth residues
constexpr int kProdW = 16;              // Bop producer
constexpr int kIssW0 = 17, kIssW1 = 18; // MMA issuers (stage parity)
constexpr int kAProdW = 19;             // E-image producer
constexpr int kThr = 20 * 32;
constexpr unsigned kSmemMax = 227u * 1024u;
constexpr float kTargetMax = 16384.f;   // 2^14: largest scaled operand magnitude (fp16 max 65504)

struct HTonal {
  int Kp, rowsP;               // modes padded to 16, operator rows padded to 16
  int nq, n1, nj, fold, src;
  unsigned a_bytes, a_off;     // E image of one range tile: 1024 Kp bytes = [re_hi | re_lo | im_hi | im_lo] x [128 x Kp] halves
  unsigned b_zbytes;           // rowsP * Kp * 2: one depth of one split of Bop
  unsigned long long b_off;
  float trK, c1r, c1i, c2r, c2i;
  float sE, sB;                // operand scales (powers of two)
};

struct HParams {
  int F, Nr, Nz, nrt, ZC, units, atype, mf, S, two_slot;
  unsigned e_rt_stride, stage_bytes, AR, smem_B, smem_acc, smem_bar;
  const unsigned char* E;
  const unsigned char* Bhi;
  const unsigned char* Blo;
  float* out;
  int* err;
  int ext_mode;
  unsigned long long* ext_key;
  unsigned int* ext_done;
  float* ext_val;
  long long* ext_idx;
  // fused final reduction across the GPUs of a box: the last CTA publishes (index_offset + index, value) to every peer's
  // buffer and collects theirs (what csrc/kernels_peer.cu does in a launch of its own); peer_bufs == null: off
  unsigned char* const* peer_bufs;
  int peer_rank, peer_world, peer_mode;
  unsigned long long peer_epoch;
  long long peer_offset;
  long long* peer_out;
  int* peer_err;
  // streamed Bop tables: tab_cnt[f] counts the finished table blocks of tonal position f (monotonic over calls), the
  // tonal's tables of this call are complete when it has reached tab_target[f]; null = wait for the whole table grid
  const unsigned int* tab_cnt;
  unsigned int tab_target[BOGP_MAX_TONALS];
  const float* phi;            // cached source mode shapes [tonal position][depth][Kp] (streamed tables)
  unsigned int phi_off[BOGP_MAX_TONALS];
  long long* dbg;              // debug build: [0, 64) issuer 0 tonal starts, [64, 128) epilogue warp 0 tonal starts (CTA 0)
  int skip;                    // debug build: 1 = no MMAs issued, 2 = epilogue releases without reading
  HTonal t[BOGP_MAX_TONALS];
};

// barrier slots (8 bytes each) at smem_bar
enum { BAR_BFULL = 0, BAR_BEMPTY = kMaxStages, BAR_AFULL = 2 * kMaxStages, BAR_ADONE = 2 * kMaxStages + 2,
       BAR_TFULL = 2 * kMaxStages + 4, BAR_TEMPTY = 2 * kMaxStages + 6, BAR_TURN = 2 * kMaxStages + 8, BAR_COUNT = 2 * kMaxStages + 10 };

// Hot-path wait of the lean epilogue / issuers: try_wait, branch, counter -- no clock reads inside the loop (the spinning
// roles share their scheduler with four epilogue warps that are bound by instruction issue).  Still bounded: a
// protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait_lean(uint32_t bar, uint32_t parity, int* err, int code) {
  uint32_t ok = 0;
#pragma unroll 1
  for (uint32_t spin = 0; spin < (1u << 26); ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
  }
  atomicExch(err, code + 1000 * (int)blockIdx.x);
  __threadfence_system();
  __trap();
}

// kind::f16 MMA, both operands from shared memory; the 64-bit descriptors are passed as halves (A and B share the high
// half: same LBO / SBO for K-major no-swizzle blocks of the same Kp).
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t desc_hi, uint32_t idesc,
                                           bool accumulate) {
  if (accumulate)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 da, db;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "mov.b64 da, {%1, %3};\n\t"
        "mov.b64 db, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(desc_hi), "r"(idesc)
        : "memory");
}

// All MMAs of one stage: per K step of 16 modes, D += A_lo B_hi + A_hi B_lo + A_hi B_hi for the re and im accumulators.
// a0 = descriptor low word of the image's first block (re_hi); blocks follow at `blk` (16-byte units): re_lo, im_hi, im_lo.
// The turn is passed to the other issuer kTurnLeft MMAs before the end of the stage: the hop (mbarrier arrive -> the other
// warp's try_wait returns) takes ~300 cycles, about the time the pipe needs for the MMAs still to be issued here, so the
// next stage's first MMA reaches the queue as this stage's last one leaves it.
#ifndef BOGP_U16_TURN_LEFT
#define BOGP_U16_TURN_LEFT 4      // measured: 6 (turn passed after the first MMA of a K = 16 stage) is 1.5 % slower
#endif
template <int KS>
__device__ __forceinline__ void issue_stage16(uint32_t d_re, uint32_t d_im, uint32_t a0, uint32_t blk, uint32_t bh, uint32_t bl,
                                              uint32_t desc_hi, uint32_t idesc, uint32_t turn_bar) {
  constexpr int kTurnAt = (6 * KS - BOGP_U16_TURN_LEFT) > 1 ? (6 * KS - BOGP_U16_TURN_LEFT) : 1;      // after this many MMAs
  asm volatile("mov.u32 %0, %0;" : "+r"(a0));
  asm volatile("mov.u32 %0, %0;" : "+r"(bh));
  asm volatile("mov.u32 %0, %0;" : "+r"(bl));
#pragma unroll
  for (int part = 0; part < 2; ++part) {
    const uint32_t d = part ? d_im : d_re;
    const uint32_t a_hi = a0 + (uint32_t)(2 * part) * blk, a_lo = a_hi + blk;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      const int n0 = (part * KS + ks) * 3;
      mma_f16_ss(d, a_lo + 16u * ks, bh + 16u * ks, desc_hi, idesc, ks > 0);
      if (n0 + 1 == kTurnAt) mbar_arrive(turn_bar);
      mma_f16_ss(d, a_hi + 16u * ks, bl + 16u * ks, desc_hi, idesc, true);
      if (n0 + 2 == kTurnAt) mbar_arrive(turn_bar);
      mma_f16_ss(d, a_hi + 16u * ks, bh + 16u * ks, desc_hi, idesc, true);
      if (n0 + 3 == kTurnAt) mbar_arrive(turn_bar);
    }
  }
}

// debug build: time stamp `ev` of stage sc (CTA 0, first 128 stages): 0 issuer ready (all waits passed), 1 issued + committed,
// 2 epilogue warp 0 woken, 3 epilogue warp 0 released, 4 epilogue warp 0 done with the stage, 5 epilogue warp 15 released
__device__ __forceinline__ void stamp16(const HParams& p, int ev, uint32_t sc) {
#if BOGP_UMMA_DBG
  if (p.dbg && blockIdx.x == 0 && sc < 128u && (threadIdx.x & 31) == 0) p.dbg[128 + ev * 128 + (int)sc] = clock64();
#endif
}

// shared-memory offset of tonal g's E image: even tonals at the low end of the A region, odd ones at the high end
__device__ __forceinline__ uint32_t a_offset(const HParams& p, int g, unsigned a_bytes) {
  return (g & 1) ? p.AR - a_bytes : 0u;
}

struct IssState {
  const HParams* p;
  uint32_t tmem_base, s_base, bar0;
  uint32_t sc;
  int g, nzu, id;
};

template <int KS>
__device__ __forceinline__ void issue_tonal(IssState& st, const HTonal& u) {
  const HParams& p = *st.p;
  const uint32_t Kp = 16u * KS;
  // K-major, no swizzle: LBO (next 16-byte K chunk) = 128 B, SBO (next 8-row group) = 16 Kp B
  const uint32_t desc_hi = (uint32_t)(make_desc(0, 128u, 16u * Kp) >> 32);
  const uint32_t desc_lo0 = (uint32_t)make_desc(0, 128u, 16u * Kp);
  const uint32_t idesc0 = (1u << 4) | ((kTile >> 4) << 24);       // D fp32, A / B fp16 (format 0), K-major both
  auto bar = [&](int i) { return st.bar0 + 8u * (uint32_t)i; };
  mbar_wait(bar(BAR_AFULL + (st.g & 1)), (uint32_t)((st.g >> 1) & 1), p.err, 4);
  if (BOGP_UMMA_DBG && p.dbg && st.id == 0 && blockIdx.x == 0 && st.g < 63 && (threadIdx.x & 31) == 0) p.dbg[st.g] = clock64();
  const uint32_t a_addr = st.s_base + a_offset(p, st.g, u.a_bytes);
  const uint32_t a0 = desc_lo0 | ((a_addr >> 4) & 0x3fffu);
  const uint32_t blk = 16u * Kp;                                    // 256 Kp bytes per block, in 16-byte units
  const int nj = u.nj, nzu = st.nzu;
  uint32_t sc = st.sc;
#pragma unroll 1
  for (int j0 = 0; j0 < nzu; j0 += nj, ++sc) {
    if ((int)(sc & 1u) != st.id) continue;
    const int njs = min(nj, nzu - j0);
    const uint32_t slot = sc % (uint32_t)p.S, ph = (sc / (uint32_t)p.S) & 1u;
    const uint32_t buf = sc & 1u, tph = (sc >> 1) & 1u;
    mbar_wait_lean(bar(BAR_BFULL + slot), ph, p.err, 5);
    mbar_wait_lean(bar(BAR_TEMPTY + buf), tph ^ 1u, p.err, 6);
    tc_fence_after();
    const uint32_t N = (uint32_t)(njs * u.rowsP);
    const uint32_t idesc = idesc0 | ((N >> 3) << 17);
    const uint32_t b_hi = st.s_base + p.smem_B + slot * p.stage_bytes;
    const uint32_t b_lo = b_hi + (uint32_t)njs * u.b_zbytes;
    const uint32_t d_re = st.tmem_base + buf * (uint32_t)(2 * kNA), d_im = d_re + (uint32_t)kNA;
    // stages enter the tensor pipe in order (see kernels_umma.cu: without the turn both buffers reach the epilogue together)
    if (sc > 0) mbar_wait_lean(bar(BAR_TURN + (st.id ^ 1)), ((sc - 1u) >> 1) & 1u, p.err, 8);
    stamp16(p, 0, sc);
    if (elect_one()) {
      if (!(BOGP_UMMA_DBG && (p.skip & 1)))
      if (BOGP_UMMA_DBG && (p.skip & 1)) mbar_arrive(bar(BAR_TURN + st.id));
      issue_stage16<KS>(d_re, d_im, a0, blk, desc_lo0 | ((b_hi >> 4) & 0x3fffu), desc_lo0 | ((b_lo >> 4) & 0x3fffu), desc_hi, idesc, bar(BAR_TURN + st.id));
      tc_commit(bar(BAR_BEMPTY + slot));
      tc_commit(bar(BAR_TFULL + buf));
    }
    __syncwarp();
    stamp16(p, 1, sc);
  }
  st.sc = sc;
  if (elect_one()) tc_commit(bar(BAR_ADONE + (st.g & 1)));
  __syncwarp();
}

// Bartlett value of one depth slot from its accumulator columns (real receiver modes: every |T|^2 term is a plain
// square; the numerator sits in the first two columns -- folded operator -- or in the nq leading (re, im) row pairs).
struct EpiTonal {
  float c1r, c1i, c2r, c2i, trK;
  int fold, nq, atype;
};
__device__ __forceinline__ void numer_first(const EpiTonal& e, const float* re, const float* im, float& num, float& numsq) {
  if (e.fold) {
    const float tr = fmaf(e.c1r, re[0], fmaf(-e.c1i, im[0], fmaf(e.c2r, re[1], -e.c2i * im[1])));
    const float ti = fmaf(e.c1r, im[0], fmaf(e.c1i, re[0], fmaf(e.c2r, im[1], e.c2i * re[1])));
    num = fmaf(tr, tr, ti * ti);
    numsq = 0.f;
  } else if (e.nq == 1) {
    const float tr = re[0] - im[1], ti = im[0] + re[1];
    num = fmaf(tr, tr, ti * ti);
    numsq = fmaf(re[0], re[0], fmaf(im[0], im[0], fmaf(re[1], re[1], im[1] * im[1])));
  } else {
    num = 0.f; numsq = 0.f;
    numerator16(re, im, e.nq, num, numsq);
  }
}
__device__ __forceinline__ void sq16(const float* re, const float* im, unsigned long long& t0, unsigned long long& t1,
                                     unsigned long long& t2, unsigned long long& t3) {
#pragma unroll
  for (int q = 0; q < 16; q += 4) {
    sq_acc2(t0, re[q], re[q + 1]);
    sq_acc2(t1, im[q], im[q + 1]);
    sq_acc2(t2, re[q + 2], re[q + 3]);
    sq_acc2(t3, im[q + 2], im[q + 3]);
  }
}


// Lean epilogue (all tonals folded, real modes, cbf + mean): Bartlett value of one depth slot.
struct FoldC { float c1r, c1i, c2r, c2i; };
__device__ __forceinline__ float fold_num(const FoldC& c, const float* re, const float* im) {
  const float tr = fmaf(c.c1r, re[0], fmaf(-c.c1i, im[0], fmaf(c.c2r, re[1], -c.c2i * im[1])));
  const float ti = fmaf(c.c1r, im[0], fmaf(c.c1i, re[0], fmaf(c.c2r, im[1], c.c2i * re[1])));
  return fmaf(tr, tr, ti * ti);
}
__device__ __forceinline__ float sum4(f32x2 t0, f32x2 t1, f32x2 t2, f32x2 t3) {
  float x, y;
  upk(add2(add2(t0, t1), add2(t2, t3)), x, y);
  return x + y;
}
__device__ __forceinline__ float lean_slot16(const FoldC& c, const float* re, const float* im) {
  const float num = fold_num(c, re, im);
  f32x2 t0 = 0ull, t1 = 0ull, t2 = 0ull, t3 = 0ull;
  sq16(re, im, t0, t1, t2, t3);
  return __fdividef(num, sum4(t0, t1, t2, t3));
}
__device__ __forceinline__ float lean_slot32(const FoldC& c, const float* reA, const float* imA, const float* reB, const float* imB) {
  const float num = fold_num(c, reA, imA);
  f32x2 t0 = 0ull, t1 = 0ull, t2 = 0ull, t3 = 0ull;
  sq16(reA, imA, t0, t1, t2, t3);
  sq16(reB, imB, t0, t1, t2, t3);
  return __fdividef(num, sum4(t0, t1, t2, t3));
}

// The lean instance has to stay at 88 registers: 640 x 88 = 56 320 leave 9 216 of the SM's 65 536 for the 128-thread
// block of the streamed table kernel that runs next to it (at 89 the allocation granularity makes it 96 and the two no
// longer fit one SM -- the launcher checks the compiled counts and falls back to the sequential table kernel).
template <bool LEAN>
__device__ __forceinline__ void umma16_grid_body(const HParams& p) {
  extern __shared__ __align__(1024) unsigned char smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t s_base = smem_u32(smem);
  const uint32_t bar0 = s_base + p.smem_bar;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem + p.smem_bar + BAR_COUNT * 8);
  auto bar = [&](int i) { return bar0 + 8u * (uint32_t)i; };

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < kMaxStages; ++s) { mbar_init(bar(BAR_BFULL + s), 1); mbar_init(bar(BAR_BEMPTY + s), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(bar(BAR_AFULL + b), 1);
      mbar_init(bar(BAR_ADONE + b), 2);          // both issuers commit
      mbar_init(bar(BAR_TFULL + b), 1);
      mbar_init(bar(BAR_TURN + b), 1);
      mbar_init(bar(BAR_TEMPTY + b), kEpiW);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == kIssW0) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512));
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);

  if (warp == kProdW) {
    // ============================================================== Bop producer
    // The Bop tables of this call.  Streamed tables (tab_cnt): the table grid is still running -- it was launched as this
    // grid's programmatic primary and produces the tonals in the order they are consumed here -- and a tonal's copies
    // start as soon as its counter has reached the call's target.  Otherwise: wait for the whole table grid (no-op
    // without a programmatic launch).
    if (!p.tab_cnt) asm volatile("griddepcontrol.wait;" ::: "memory");
    uint32_t sc = 0;
    int f_ready = 0;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      const int zc = unit / p.nrt, z0 = zc * p.ZC, nzu = min(p.ZC, p.Nz - z0);
      for (int f = 0; f < p.F; ++f) {
        const HTonal& u = p.t[f];
        if (p.tab_cnt && f >= f_ready) {
          const long long t0 = clock64();
          for (unsigned spin = 0;; ++spin) {
            unsigned int seen;
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.tab_cnt + f) : "memory");
            if ((int)(seen - p.tab_target[f]) >= 0) break;
            if ((spin & 255u) == 255u && clock64() - t0 > 4000000000LL) {      // ~2 s: the table grid never got there
              atomicExch(p.err, 9 + 1000 * (int)blockIdx.x);
              __threadfence_system();
              __trap();
            }
          }
          asm volatile("fence.proxy.async.global;" ::: "memory");      // generic-proxy table stores -> this warp's bulk copies
          f_ready = f + 1;
        }
        for (int j0 = 0; j0 < nzu; j0 += u.nj, ++sc) {
          const int njs = min(u.nj, nzu - j0);
          const uint32_t slot = sc % (uint32_t)p.S, ph = (sc / (uint32_t)p.S) & 1u;
          mbar_wait_relaxed<128>(bar(BAR_BEMPTY + slot), ph ^ 1u, p.err, 3);
          const uint32_t bytes = (uint32_t)njs * u.b_zbytes;
          const uint32_t dst = s_base + p.smem_B + slot * p.stage_bytes;
          const size_t src = (size_t)u.b_off + (size_t)(z0 + j0) * u.b_zbytes;
          if (elect_one()) {
            if (BOGP_UMMA_DBG && (p.skip & 4)) mbar_arrive(bar(BAR_BFULL + slot));
            else {
            mbar_expect_tx(bar(BAR_BFULL + slot), 2 * bytes);
            bulk_g2s(dst, p.Bhi + src, bytes, bar(BAR_BFULL + slot));
            bulk_g2s(dst + bytes, p.Blo + src, bytes, bar(BAR_BFULL + slot));
            }
          }
          __syncwarp();
        }
      }
    }
  } else if (warp == kAProdW) {
    // ============================================================== E-image producer (one bulk copy per tonal, one tonal ahead)
    int g = 0;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      const int rt = unit % p.nrt;
      for (int f = 0; f < p.F; ++f, ++g) {
        const HTonal& u = p.t[f];
        if (g >= 2) mbar_wait_relaxed<256>(bar(BAR_ADONE + (g & 1)), (uint32_t)(((g - 2) >> 1) & 1), p.err, 1);
        if (g >= 1) {
          const HTonal& up = p.t[f == 0 ? p.F - 1 : f - 1];
          if (up.a_bytes + u.a_bytes > p.AR)
            mbar_wait_relaxed<256>(bar(BAR_ADONE + ((g - 1) & 1)), (uint32_t)(((g - 1) >> 1) & 1), p.err, 2);
        }
        if (elect_one()) {
          mbar_expect_tx(bar(BAR_AFULL + (g & 1)), u.a_bytes);
          bulk_g2s(s_base + a_offset(p, g, u.a_bytes), p.E + (size_t)rt * p.e_rt_stride + u.a_off, u.a_bytes, bar(BAR_AFULL + (g & 1)));
        }
        __syncwarp();
      }
    }
  } else if (warp == kIssW0 || warp == kIssW1) {
    // ============================================================== MMA issuers
    IssState st{};
    st.id = warp == kIssW0 ? 0 : 1;
    st.p = &p; st.tmem_base = tmem_base; st.s_base = s_base; st.bar0 = bar0;
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      const int zc = unit / p.nrt, z0 = zc * p.ZC;
      st.nzu = min(p.ZC, p.Nz - z0);
      for (int f = 0; f < p.F; ++f, ++st.g) {
        switch (p.t[f].Kp >> 4) {
          case 1: issue_tonal<1>(st, p.t[f]); break;
          case 2: issue_tonal<2>(st, p.t[f]); break;
          case 3: issue_tonal<3>(st, p.t[f]); break;
          default: issue_tonal<4>(st, p.t[f]); break;
        }
      }
    }
  } else {
    // ============================================================== epilogue
    const int quad = warp & 3;               // TMEM lane quadrant this warp may read
    const int sub = warp >> 2;               // depth residue: this warp owns the depths j with j % 4 == sub
    const int tl = quad * 32 + lane;
    const uint32_t t_lane0 = tmem_base + ((uint32_t)(quad * 32) << 16);
    float* acc = reinterpret_cast<float*>(smem + p.smem_acc);     // [ZC][128]
    uint32_t sc = 0;
    // err / arg-ext words are reset by the table kernel: with streamed tables the wait for that grid's completion moves
    // to the first write-out (long over by then)
    if (!p.tab_cnt) asm volatile("griddepcontrol.wait;" ::: "memory");
    if (LEAN) {
      // ---- lean path: every tonal folded with 16 or 32 padded rows, cbf + mean.  Straight-line stages: wait, four
      // loads, one wait, release, arithmetic; acc starts at zero so that every tonal is a plain +=.
      const uint32_t tb0 = t_lane0 + (uint32_t)(sub * 16), tb1 = t_lane0 + (uint32_t)(sub * 32);
      for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
        const int rt = unit % p.nrt, zc = unit / p.nrt;
        const int z0 = zc * p.ZC, nzu = min(p.ZC, p.Nz - z0);
        for (int j = sub; j < nzu; j += 4) acc[(size_t)j * kTile + tl] = 0.f;
        for (int f = 0; f < p.F; ++f) {
          const HTonal& u = p.t[f];
          FoldC fc;
          fc.c1r = u.c1r; fc.c1i = u.c1i; fc.c2r = u.c2r; fc.c2i = u.c2i;
          if (BOGP_UMMA_DBG && p.dbg && warp == 0 && lane == 0 && blockIdx.x == 0 && unit == 0 && f < 63) p.dbg[64 + f] = clock64();
          float* a = acc + (size_t)sub * kTile + tl;
          if (u.rowsP == 16) {
            const int full = nzu >> 3;
#pragma unroll 1
            for (int st = 0; st < full; ++st, ++sc, a += 8 * kTile) {
              const uint32_t buf = sc & 1u;
              mbar_wait_lean(bar(BAR_TFULL + buf), (sc >> 1) & 1u, p.err, 7);
              tc_fence_after();
              if (warp == 0) stamp16(p, 2, sc);
              const uint32_t tA = tb0 + buf * (uint32_t)(2 * kNA);
              float reA[16], imA[16], reB[16], imB[16];
              tc_ld16(tA, reA); tc_ld16(tA + kNA, imA);
              tc_ld16(tA + 64u, reB); tc_ld16(tA + kNA + 64u, imB);
              tc_wait_ld();
              tc_fence_before();
              if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + buf));
              if (warp == 0) stamp16(p, 3, sc); else if (warp == 15) stamp16(p, 5, sc);
              a[0] += lean_slot16(fc, reA, imA);
              a[4 * kTile] += lean_slot16(fc, reB, imB);
              if (warp == 0) stamp16(p, 4, sc);
            }
            if (nzu & 7) {          // ragged last stage: slots sub and sub + 4 if they exist
              const int njs = nzu & 7;
              const uint32_t buf = sc & 1u;
              mbar_wait_lean(bar(BAR_TFULL + buf), (sc >> 1) & 1u, p.err, 7);
              tc_fence_after();
              const uint32_t tA = tb0 + buf * (uint32_t)(2 * kNA);
              float reA[16], imA[16], reB[16], imB[16];
              tc_ld16(tA, reA); tc_ld16(tA + kNA, imA);
              tc_ld16(tA + 64u, reB); tc_ld16(tA + kNA + 64u, imB);
              tc_wait_ld();
              tc_fence_before();
              if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + buf));
              if (sub < njs) a[0] += lean_slot16(fc, reA, imA);
              if (sub + 4 < njs) a[4 * kTile] += lean_slot16(fc, reB, imB);
              ++sc;
            }
          } else {
            const int full = nzu >> 2;
#pragma unroll 1
            for (int st = 0; st < full; ++st, ++sc, a += 4 * kTile) {
              const uint32_t buf = sc & 1u;
              mbar_wait_lean(bar(BAR_TFULL + buf), (sc >> 1) & 1u, p.err, 7);
              tc_fence_after();
              if (warp == 0) stamp16(p, 2, sc);
              const uint32_t tA = tb1 + buf * (uint32_t)(2 * kNA);
              float reA[16], imA[16], reB[16], imB[16];
              tc_ld16(tA, reA); tc_ld16(tA + kNA, imA);
              tc_ld16(tA + 16u, reB); tc_ld16(tA + kNA + 16u, imB);
              tc_wait_ld();
              tc_fence_before();
              if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + buf));
              if (warp == 0) stamp16(p, 3, sc); else if (warp == 15) stamp16(p, 5, sc);
              a[0] += lean_slot32(fc, reA, imA, reB, imB);
              if (warp == 0) stamp16(p, 4, sc);
            }
            if (nzu & 3) {
              const int njs = nzu & 3;
              const uint32_t buf = sc & 1u;
              mbar_wait_lean(bar(BAR_TFULL + buf), (sc >> 1) & 1u, p.err, 7);
              tc_fence_after();
              const uint32_t tA = tb1 + buf * (uint32_t)(2 * kNA);
              float reA[16], imA[16], reB[16], imB[16];
              tc_ld16(tA, reA); tc_ld16(tA + kNA, imA);
              tc_ld16(tA + 16u, reB); tc_ld16(tA + kNA + 16u, imB);
              tc_wait_ld();
              tc_fence_before();
              if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + buf));
              if (sub < njs) a[0] += lean_slot32(fc, reA, imA, reB, imB);
              ++sc;
            }
          }
        }
        // write-out: this warp's own depths (j % 4 == sub), nobody else touched them -- no barrier
        __syncwarp();
        if (p.tab_cnt) asm volatile("griddepcontrol.wait;" ::: "memory");
        const int i = rt * kTile + tl;
        const float scale = 1.f / (float)p.F;
        unsigned long long best = 0ull;
        if (i < p.Nr)
          for (int j = sub; j < nzu; j += 4) {
            const float v = acc[(size_t)j * kTile + tl] * scale;
            p.out[(size_t)(z0 + j) * p.Nr + i] = v;
            if (p.ext_mode && v == v) {
              uint32_t uu = __float_as_uint(v);
              uu = (uu & 0x80000000u) ? ~uu : (uu | 0x80000000u);
              if (p.ext_mode == 2) uu = ~uu;
              const unsigned long long key = ((unsigned long long)uu << 32) | (0xffffffffu - ((uint32_t)(z0 + j) * (uint32_t)p.Nr + (uint32_t)i));
              best = key > best ? key : best;
            }
          }
        if (p.ext_mode) {
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
            best = other > best ? other : best;
          }
          if (lane == 0 && best) atomicMax(p.ext_key, best);
        }
      }
    } else
    for (int unit = blockIdx.x; unit < p.units; unit += gridDim.x) {
      const int rt = unit % p.nrt, zc = unit / p.nrt;
      const int z0 = zc * p.ZC, nzu = min(p.ZC, p.Nz - z0);
      for (int f = 0; f < p.F; ++f) {
        const HTonal& u = p.t[f];
        const int rowsP = u.rowsP;
        EpiTonal et;
        et.c1r = u.c1r; et.c1i = u.c1i; et.c2r = u.c2r; et.c2i = u.c2i; et.trK = u.trK; et.fold = u.fold; et.nq = u.nq; et.atype = p.atype;
        const bool fast = (u.n1 == 2 * u.nq) && u.nq <= 8;
        const bool simple = fast && (u.nq == 1 || u.fold) && !(BOGP_UMMA_DBG && (p.skip & 2));
        if (BOGP_UMMA_DBG && p.dbg && warp == 0 && lane == 0 && blockIdx.x == 0 && unit == 0 && f < 63) p.dbg[64 + f] = clock64();
        auto store = [&](int j, float num, float norm) {
          const float b = __fdividef(num, norm);
          const float val = et.atype == BOGP_ATYPE_CBF ? b : et.trK - b;
          float* a = acc + (size_t)j * kTile + tl;
          *a = (f == 0) ? val : (p.mf == BOGP_MF_MEAN ? *a + val : *a * val);
        };
        for (int j0 = 0; j0 < nzu; j0 += u.nj, ++sc) {
          const int njs = min(u.nj, nzu - j0);
          const uint32_t buf = sc & 1u;
          const uint32_t t_lane = t_lane0 + buf * (uint32_t)(2 * kNA);
          mbar_wait_lean(bar(BAR_TFULL + buf), (sc >> 1) & 1u, p.err, 7);
          tc_fence_after();
          if (warp == 0) stamp16(p, 2, sc);
          const int s_first = (sub - (j0 & 3) + 4) & 3;      // first slot of this stage with (j0 + s) % 4 == sub
          bool released = false;
          if ((p.two_slot & 1) && simple && rowsP == 16 && s_first + 4 < njs && s_first + 8 >= njs) {
            // two 16-column slots: all four loads in flight, one wait, release, then the arithmetic
            const int sA = s_first, sB = s_first + 4;
            float reA[16], imA[16], reB[16], imB[16];
            const uint32_t tA = t_lane + (uint32_t)(sA * 16), tB = t_lane + (uint32_t)(sB * 16);
            tc_ld16(tA, reA); tc_ld16(tA + kNA, imA);
            tc_ld16(tB, reB); tc_ld16(tB + kNA, imB);
            tc_wait_ld();
            tc_fence_before();
            if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + buf));
            released = true;
            if (warp == 0) stamp16(p, 3, sc); else if (warp == 15) stamp16(p, 5, sc);
            {
              float num, numsq;
              numer_first(et, reA, imA, num, numsq);
              unsigned long long t0 = 0ull, t1 = 0ull, t2 = 0ull, t3 = 0ull;
              sq16(reA, imA, t0, t1, t2, t3);
              store(j0 + sA, num, ((sum2(t0) + sum2(t1)) + (sum2(t2) + sum2(t3))) - numsq);
            }
            {
              float num, numsq;
              numer_first(et, reB, imB, num, numsq);
              unsigned long long t0 = 0ull, t1 = 0ull, t2 = 0ull, t3 = 0ull;
              sq16(reB, imB, t0, t1, t2, t3);
              store(j0 + sB, num, ((sum2(t0) + sum2(t1)) + (sum2(t2) + sum2(t3))) - numsq);
            }
          } else if ((p.two_slot & 2) && simple && rowsP == 32 && s_first < njs && s_first + 4 >= njs) {
            // one slot of two 16-column chunks: all four loads in flight, one wait
            float reA[16], imA[16], reB[16], imB[16];
            const uint32_t tA = t_lane + (uint32_t)(s_first * 32);
            tc_ld16(tA, reA); tc_ld16(tA + kNA, imA);
            tc_ld16(tA + 16u, reB); tc_ld16(tA + kNA + 16u, imB);
            tc_wait_ld();
            tc_fence_before();
            if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + buf));
            released = true;
            if (warp == 0) stamp16(p, 3, sc); else if (warp == 15) stamp16(p, 5, sc);
            float num, numsq;
            numer_first(et, reA, imA, num, numsq);
            unsigned long long t0 = 0ull, t1 = 0ull, t2 = 0ull, t3 = 0ull;
            sq16(reA, imA, t0, t1, t2, t3);
            sq16(reB, imB, t0, t1, t2, t3);
            store(j0 + s_first, num, ((sum2(t0) + sum2(t1)) + (sum2(t2) + sum2(t3))) - numsq);
          } else {
            for (int s = s_first; s < njs && !(BOGP_UMMA_DBG && (p.skip & 2)); s += 4) {
              const bool last_slot = s + 4 >= njs;
              const uint32_t t_re = t_lane + (uint32_t)(s * rowsP);
              float norm = 0.f, num = 0.f;
              if (fast) {
                unsigned long long t0 = 0ull, t1 = 0ull, t2 = 0ull, t3 = 0ull;
                float numsq = 0.f;
                const int c_last = rowsP - 16;
                for (int c0 = 0; c0 <= c_last; c0 += 16) {
                  float re[16], im[16];
                  tc_ld16(t_re + (uint32_t)c0, re);
                  tc_ld16(t_re + kNA + (uint32_t)c0, im);
                  tc_wait_ld();
                  if (last_slot && c0 == c_last) {
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + buf));
                    released = true;
                  }
                  if (c0 == 0) numer_first(et, re, im, num, numsq);
                  sq16(re, im, t0, t1, t2, t3);
                }
                norm = ((sum2(t0) + sum2(t1)) + (sum2(t2) + sum2(t3))) - numsq;
              } else {
                // complex receiver modes / general layout: (re, im) row pairs below n1, plain rows from n1 on
                const int nq2 = 2 * u.nq, n1 = u.n1;
                for (int c0 = 0; c0 < rowsP; c0 += 16) {
                  float re[16], im[16];
                  tc_ld16(t_re + (uint32_t)c0, re);
                  tc_ld16(t_re + kNA + (uint32_t)c0, im);
                  tc_wait_ld();
                  if (last_slot && c0 + 16 >= rowsP) {
                    tc_fence_before();
                    if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + buf));
                    released = true;
                  }
#pragma unroll
                  for (int q = 0; q < 8; ++q) {
                    const int c = c0 + 2 * q;
                    const float r0 = re[2 * q], r1 = re[2 * q + 1], i0 = im[2 * q], i1 = im[2 * q + 1];
                    if (c >= n1) {
                      norm += r0 * r0 + i0 * i0 + r1 * r1 + i1 * i1;
                    } else {
                      const float tr = r0 - i1, ti = i0 + r1;
                      const float v = tr * tr + ti * ti;
                      if (c < nq2) num += v; else norm += v;
                    }
                  }
                }
              }
              store(j0 + s, num, norm);
            }
          }
          if (!released) {      // no slot of this stage belongs to this warp
            tc_fence_before();
            if (lane == 0) mbar_arrive(bar(BAR_TEMPTY + buf));
            if (warp == 0) stamp16(p, 3, sc); else if (warp == 15) stamp16(p, 5, sc);
          }
          if (warp == 0) stamp16(p, 4, sc);
        }
      }
      // write-out: this warp's own depths (j % 4 == sub), nobody else touched them -- no barrier
      __syncwarp();
      if (p.tab_cnt) asm volatile("griddepcontrol.wait;" ::: "memory");
      const int i = rt * kTile + tl;
      const float scale = p.mf == BOGP_MF_MEAN ? 1.f / (float)p.F : 1.f;
      unsigned long long best = 0ull;
      if (i < p.Nr)
        for (int j = sub; j < nzu; j += 4) {
          const float v = acc[(size_t)j * kTile + tl] * scale;
          p.out[(size_t)(z0 + j) * p.Nr + i] = v;
          if (p.ext_mode && v == v) {
            uint32_t uu = __float_as_uint(v);
            uu = (uu & 0x80000000u) ? ~uu : (uu | 0x80000000u);         // order-preserving map float -> uint32
            if (p.ext_mode == 2) uu = ~uu;                               // arg-min: largest key = smallest value
            const unsigned long long key = ((unsigned long long)uu << 32) | (0xffffffffu - ((uint32_t)(z0 + j) * (uint32_t)p.Nr + (uint32_t)i));
            best = key > best ? key : best;                              // ties: lowest flat index (np.argmax rule)
          }
        }
      if (p.ext_mode) {
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
          const unsigned long long other = __shfl_xor_sync(0xffffffffu, best, o);
          best = other > best ? other : best;
        }
        if (lane == 0 && best) atomicMax(p.ext_key, best);
      }
    }
  }
  if (BOGP_UMMA_DBG && p.dbg && blockIdx.x == 0 && (threadIdx.x & 31) == 0) {
    if (warp == kIssW0) p.dbg[63] = clock64();
    if (warp == 0) p.dbg[127] = clock64();
  }
  tc_fence_before();
  __syncthreads();
  long long* xrec = reinterpret_cast<long long*>(smem + p.smem_acc);      // (index, value bits, last-CTA flag): acc is free now
  if (p.ext_mode && threadIdx.x == 0) {
    xrec[2] = 0;
    __threadfence();
    if (atomicAdd(p.ext_done, 1u) == gridDim.x - 1) {
      __threadfence();
      const unsigned long long key = atomicMax(p.ext_key, 0ull);
      float val = 0.f;
      long long idx = -1;
      if (key != 0ull) {
        uint32_t uu = (uint32_t)(key >> 32);
        if (p.ext_mode == 2) uu = ~uu;
        uu = (uu & 0x80000000u) ? (uu & 0x7fffffffu) : ~uu;
        val = __uint_as_float(uu);
        idx = (long long)(0xffffffffu - (uint32_t)(key & 0xffffffffu));
      }
      *p.ext_val = val;
      *p.ext_idx = idx;
      xrec[0] = idx < 0 ? idx : idx + p.peer_offset;
      xrec[1] = (long long)__float_as_uint(val);
      xrec[2] = 1;
    }
  }
  if (p.ext_mode && p.peer_bufs) {
    __syncthreads();
    if (xrec[2] == 1 && (int)threadIdx.x < p.peer_world) {
      // thread t: this rank's record into peer t's buffer, peer t's record out of this rank's buffer (peer_device.cuh:
      // collective or pipelined by one step)
      peer_step(p.peer_bufs, threadIdx.x, p.peer_rank, p.peer_world, p.peer_epoch, p.peer_mode, xrec[0], xrec[1], p.peer_out, p.peer_err);
    }
  }
  if (warp == kIssW0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512));
  }
}

// (nvcc does not take __launch_bounds__ and __maxnreg__ on the same kernel, and with __maxnreg__(88) alone the lean
// instance compiles to 85 registers and runs 2 % slower: the launch bounds stay, the launcher checks the count.)
__global__ void __launch_bounds__(kThr, 1) umma16_grid_kernel_lean(const __grid_constant__ HParams p) { umma16_grid_body<true>(p); }
__global__ void __launch_bounds__(kThr, 1) umma16_grid_kernel_general(const __grid_constant__ HParams p) { umma16_grid_body<false>(p); }

// ------------------------------------------------------------------------------------------------ operand tables
// byte offset of the 16-byte chunk (row, K chunk kc) inside a K-major no-swizzle block with nch chunks per row
__device__ __forceinline__ size_t chunk_off(int row, int kc, int nch) {
  return ((size_t)((row >> 3) * nch + kc) * 8 + (row & 7)) * 16;
}
// x (already scaled) -> fp16 hi + fp16 lo, packed two modes per 32-bit word (packed conversions: one instruction per pair)
__device__ __forceinline__ void split8(const float* x, uint4& hi, uint4& lo) {
  uint32_t h[4], l[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const __half2 hh = __floats2half2_rn(x[2 * e], x[2 * e + 1]);          // low half = even mode
    const float2 hf = __half22float2(hh);
    const __half2 ll = __floats2half2_rn(x[2 * e] - hf.x, x[2 * e + 1] - hf.y);
    h[e] = *reinterpret_cast<const uint32_t*>(&hh);
    l[e] = *reinterpret_cast<const uint32_t*>(&ll);
  }
  hi = make_uint4(h[0], h[1], h[2], h[3]);
  lo = make_uint4(l[0], l[1], l[2], l[3]);
}

// E images: per (range tile, tonal) four blocks [re_hi | re_lo | im_hi | im_lo], each the [128 ranges x Kp] matrix
// E_f^T * sE in K-major no-swizzle order.  E[m, i] = exp(-i k_m r_i) / sqrt(k_m r_i) in float64 (k r reaches 1.6e4 rad).
// One thread = one (range, 8-mode chunk).  Kept across calls while the range vector is the same (the reference sweeps
// one grid over all time steps of an event, ref/projects/swellex96_localization/data/mfp.py:166-216).
constexpr int kETabRows = 32;
__global__ void __launch_bounds__(128) umma16_E_kernel(ModesetDev ms, HParams p, const double* __restrict__ r) {
  __shared__ double2 kinv[64];
  const int per_f = p.nrt * (kTile / kETabRows);
  const int f = blockIdx.x / per_f, rem = blockIdx.x % per_f;
  const int rt = rem / (kTile / kETabRows), row0 = (rem % (kTile / kETabRows)) * kETabRows;
  const HTonal& u = p.t[f];
  const TonalMeta& t = ms.t[u.src];
  for (int m = threadIdx.x; m < u.Kp; m += blockDim.x) {
    double cr = 0.0, ci = 0.0;
    if (m < t.M) {
      const double kre = ms.k_re[t.offM + m], kim = ms.k_im[t.offM + m];
      const double mag = rsqrt(sqrt(kre * kre + kim * kim)), ang = -0.5 * atan2(kim, kre);
      sincos(ang, &ci, &cr);
      cr *= mag; ci *= mag;
    }
    kinv[m] = make_double2(cr, ci);
  }
  __syncthreads();
  const int nch = u.Kp >> 3;
  unsigned char* base = const_cast<unsigned char*>(p.E) + (size_t)rt * p.e_rt_stride + u.a_off;
  const size_t blk = (size_t)256 * u.Kp;
  const double sE = (double)u.sE;
  for (int idx = threadIdx.x; idx < kETabRows * nch; idx += blockDim.x) {
    const int row = row0 + (idx % kETabRows), kc = idx / kETabRows, i = rt * kTile + row;
    float xr[8], xi[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int m = kc * 8 + e;
      double vr = 0.0, vi = 0.0;
      if (i < p.Nr && m < t.M) {
        const double kre = ms.k_re[t.offM + m], kim = ms.k_im[t.offM + m], rr = r[i];
        double s, c;
        sincos(kre * rr, &s, &c);
        const double amp = exp(kim * rr) * rsqrt(rr) * sE;
        const double2 q = kinv[m];
        vr = amp * (q.x * c + q.y * s);          // (q.x + i q.y) (c - i s)
        vi = amp * (q.y * c - q.x * s);
      }
      xr[e] = (float)vr; xi[e] = (float)vi;
    }
    uint4 rh, rl, ih, il;
    split8(xr, rh, rl);
    split8(xi, ih, il);
    const size_t o = chunk_off(row, kc, nch);
    *reinterpret_cast<uint4*>(base + o) = rh;
    *reinterpret_cast<uint4*>(base + blk + o) = rl;
    *reinterpret_cast<uint4*>(base + 2 * blk + o) = ih;
    *reinterpret_cast<uint4*>(base + 3 * blk + o) = il;
  }
}

// Bop tables: per (tonal, depth) the block W_f diag(phi(z_j)) * sB, [rowsP x Kp] halves in K-major no-swizzle order, hi
// and lo arrays.  One block = kBDepths depths of one tonal; a thread keeps its 8-mode chunks of W_f in registers.
constexpr int kBDepths = 8;
constexpr int kBThreads = 128;
__device__ __forceinline__ void b_block(const ModesetDev& ms, const HParams& p, const double* __restrict__ z, int f, int j0,
                                        float (*phi_s)[64]) {
  const HTonal& u = p.t[f];
  const TonalMeta& t = ms.t[u.src];
  const int nd = min(kBDepths, p.Nz - j0), tid = threadIdx.x;
  for (int e = tid; e < nd * u.Kp; e += kBThreads) {
    const int dj = e / u.Kp, m = e - dj * u.Kp;
    float v = 0.f;
    if (m < t.M) {
      int i0; float w;
      depth_index(z[j0 + dj], ms.z0, ms.dz, ms.nz_tab, i0, w);
      const float* tr = ms.tab_re + t.offTab + (size_t)i0 * t.Mp + m;
      v = fmaf(w, __ldg(tr + t.Mp) - __ldg(tr), __ldg(tr)) * u.sB;
    }
    phi_s[dj][m] = v;
  }
  __syncthreads();
  const int nch = u.Kp >> 3, n_q = u.rowsP * nch;
  const size_t zstride = (size_t)u.b_zbytes;
  unsigned char* Bhi = const_cast<unsigned char*>(p.Bhi) + u.b_off + (size_t)j0 * zstride;
  unsigned char* Blo = const_cast<unsigned char*>(p.Blo) + u.b_off + (size_t)j0 * zstride;
  for (int q = tid; q < n_q; q += kBThreads) {
    const int pc = q >> 3, kc = pc % nch, rg = pc / nch, row = rg * 8 + (q & 7);
    float w[8];
    if (kc * 8 < t.Mp) {
      const float4 a = __ldg(reinterpret_cast<const float4*>(ms.Wu + t.offWu + (size_t)row * t.Mp + kc * 8));
      const float4 b = __ldg(reinterpret_cast<const float4*>(ms.Wu + t.offWu + (size_t)row * t.Mp + kc * 8) + 1);
      w[0] = a.x; w[1] = a.y; w[2] = a.z; w[3] = a.w; w[4] = b.x; w[5] = b.y; w[6] = b.z; w[7] = b.w;
    } else {
#pragma unroll
      for (int e = 0; e < 8; ++e) w[e] = 0.f;
    }
    for (int dj = 0; dj < nd; ++dj) {
      const float4 pa = *reinterpret_cast<const float4*>(&phi_s[dj][kc * 8]), pb = *reinterpret_cast<const float4*>(&phi_s[dj][kc * 8 + 4]);
      const float x[8] = {w[0] * pa.x, w[1] * pa.y, w[2] * pa.z, w[3] * pa.w, w[4] * pb.x, w[5] * pb.y, w[6] * pb.z, w[7] * pb.w};
      uint4 hi, lo;
      split8(x, hi, lo);
      *reinterpret_cast<uint4*>(Bhi + (size_t)dj * zstride + (size_t)q * 16) = hi;
      *reinterpret_cast<uint4*>(Blo + (size_t)dj * zstride + (size_t)q * 16) = lo;
    }
  }
}

__global__ void __launch_bounds__(kBThreads) umma16_B_kernel(ModesetDev ms, HParams p, const double* __restrict__ z) {
  __shared__ __align__(16) float phi_s[kBDepths][64];
  const int nzb = (p.Nz + kBDepths - 1) / kBDepths;
  // programmatic dependent launch: the surface kernel may start its prologue (barriers, TMEM, E images) while these
  // blocks run; its Bop producer and epilogue warps wait for this grid's completion with griddepcontrol.wait
  asm volatile("griddepcontrol.launch_dependents;");
  if (blockIdx.x == 0 && threadIdx.x == 0) { *p.err = 0; if (p.ext_mode) { *p.ext_key = 0ull; *p.ext_done = 0u; } }
  b_block(ms, p, z, blockIdx.x / nzb, (blockIdx.x % nzb) * kBDepths, phi_s);
}

// Source mode shapes at the call's depths, interpolated from the handle's depth table: phi[f][j][m] (unscaled, zero for
// m >= M).  Covariance-independent: kept while the depth vector is the same (one grid for all time steps of a sweep).
__global__ void __launch_bounds__(256) umma16_phi_kernel(ModesetDev ms, HParams p, const double* __restrict__ z, float* __restrict__ phi) {
  const int f = blockIdx.y;
  const HTonal& u = p.t[f];
  const TonalMeta& t = ms.t[u.src];
  const int n = p.Nz * u.Kp;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
    const int j = e / u.Kp, m = e - j * u.Kp;
    float v = 0.f;
    if (m < t.M) {
      int i0; float w;
      depth_index(z[j], ms.z0, ms.dz, ms.nz_tab, i0, w);
      const float* tr = ms.tab_re + t.offTab + (size_t)i0 * t.Mp + m;
      v = fmaf(w, __ldg(tr + t.Mp) - __ldg(tr), __ldg(tr));
    }
    phi[p.phi_off[f] + e] = v;
  }
}

// Streamed form: one small block per SM (128 threads, <= 64 registers: it has to fit NEXT to a surface-kernel block,
// which leaves 9 216 registers and ~9 KB of shared memory on its SM) walks the (tonal, 8-depth block) items in the order
// the surface kernel consumes the tonals and counts every finished item in tab_cnt[tonal position].  The surface kernel
// is launched as this grid's programmatic dependent and runs AT THE SAME TIME: its Bop producer starts a tonal's bulk
// copies when the counter has reached the call's target, so only the first (smallest) tonal's tables are waited for
// and the rest of the table build hides under the MMAs of the tonals before.  With four warps per SM every latency is
// exposed, so: no shared memory and no block barrier (mode shapes come from the cached phi table, the operator chunk
// from L1), and the (depth, row, K chunk) triples of an item are independent iterations the compiler can overlap.
// Same values as umma16_B_kernel: (W * sB) * phi == W * (phi * sB) for a power-of-two sB.
__global__ void __launch_bounds__(kBThreads, 8) umma16_B_stream_kernel(ModesetDev ms, HParams p, unsigned int* tab_cnt) {
  asm volatile("griddepcontrol.launch_dependents;");
  if (blockIdx.x == 0 && threadIdx.x == 0) { *p.err = 0; if (p.ext_mode) { *p.ext_key = 0ull; *p.ext_done = 0u; } }
  const int nzb = (p.Nz + kBDepths - 1) / kBDepths, items = p.F * nzb;
  const int tid = threadIdx.x;
  for (int it = blockIdx.x; it < items; it += gridDim.x) {
    const int f = it / nzb, j0 = (it - f * nzb) * kBDepths;
    const HTonal& u = p.t[f];
    const TonalMeta& t = ms.t[u.src];
    const int nd = min(kBDepths, p.Nz - j0);
    const int Kp = u.Kp, nch = Kp >> 3, n_q = u.rowsP * nch, Mp = t.Mp;
    const float sB = u.sB;
    const size_t zstride = (size_t)u.b_zbytes;
    unsigned char* Bhi = const_cast<unsigned char*>(p.Bhi) + u.b_off + (size_t)j0 * zstride;
    unsigned char* Blo = const_cast<unsigned char*>(p.Blo) + u.b_off + (size_t)j0 * zstride;
    const float* W = ms.Wu + t.offWu;
    const float* ph = p.phi + p.phi_off[f] + (size_t)j0 * Kp;
    // a thread keeps its (row, K chunk) of the scaled operator in registers and walks the item's depths (independent
    // iterations: the loads of all eight are issued ahead of the arithmetic)
    for (int q = tid; q < n_q; q += kBThreads) {
      const int pc = q >> 3, kc = pc % nch, row = (pc / nch) * 8 + (q & 7);
      float w[8];
      if (kc * 8 < Mp) {
        const float4 a = __ldg(reinterpret_cast<const float4*>(W + (size_t)row * Mp + kc * 8));
        const float4 b = __ldg(reinterpret_cast<const float4*>(W + (size_t)row * Mp + kc * 8) + 1);
        w[0] = a.x * sB; w[1] = a.y * sB; w[2] = a.z * sB; w[3] = a.w * sB; w[4] = b.x * sB; w[5] = b.y * sB; w[6] = b.z * sB; w[7] = b.w * sB;
      } else {
#pragma unroll
        for (int e = 0; e < 8; ++e) w[e] = 0.f;
      }
      const float* pq = ph + kc * 8;
      unsigned char* oh = Bhi + (size_t)q * 16;
      unsigned char* ol = Blo + (size_t)q * 16;
      if (nd == kBDepths) {
        // full item: no predicates, so the eight pairs of loads are issued ahead of the arithmetic
        float4 pa[kBDepths], pb[kBDepths];
#pragma unroll
        for (int dj = 0; dj < kBDepths; ++dj) {
          pa[dj] = __ldg(reinterpret_cast<const float4*>(pq + dj * Kp));
          pb[dj] = __ldg(reinterpret_cast<const float4*>(pq + dj * Kp) + 1);
        }
#pragma unroll
        for (int dj = 0; dj < kBDepths; ++dj) {
          const float x[8] = {w[0] * pa[dj].x, w[1] * pa[dj].y, w[2] * pa[dj].z, w[3] * pa[dj].w, w[4] * pb[dj].x, w[5] * pb[dj].y, w[6] * pb[dj].z, w[7] * pb[dj].w};
          uint4 hi, lo;
          split8(x, hi, lo);
          *reinterpret_cast<uint4*>(oh + (size_t)dj * zstride) = hi;
          *reinterpret_cast<uint4*>(ol + (size_t)dj * zstride) = lo;
        }
      } else {
        for (int dj = 0; dj < nd; ++dj) {
          const float4 pa = __ldg(reinterpret_cast<const float4*>(pq + dj * Kp)), pb = __ldg(reinterpret_cast<const float4*>(pq + dj * Kp) + 1);
          const float x[8] = {w[0] * pa.x, w[1] * pa.y, w[2] * pa.z, w[3] * pa.w, w[4] * pb.x, w[5] * pb.y, w[6] * pb.z, w[7] * pb.w};
          uint4 hi, lo;
          split8(x, hi, lo);
          *reinterpret_cast<uint4*>(oh + (size_t)dj * zstride) = hi;
          *reinterpret_cast<uint4*>(ol + (size_t)dj * zstride) = lo;
        }
      }
    }
    // one count per warp (target: 4 per item): no block barrier, the warps of a block run decoupled
    // (release at gpu scope: the lanes' stores are ordered before lane 0 by the warp barrier; no L1 invalidation, the
    // operator chunks and mode shapes stay cached for the next item)
    __syncwarp();
    if ((tid & 31) == 0) {
      asm volatile("fence.proxy.async.global;" ::: "memory");
      asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(tab_cnt + f) : "memory");
    }
  }
}

// ------------------------------------------------------------------------------------------------ host side
int device_sms16() {
  int dev = 0, n = 148;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  return n;
}

float pow2_scale(double bound) {
  if (!(bound > 0.0) || !std::isfinite(bound)) return 1.f;
  return (float)std::ldexp(1.0, (int)std::floor(std::log2((double)kTargetMax / bound)));
}

// smallest, largest, second smallest, ... by padded mode count: consecutive E images fit side by side in the A region
void tonal_order16(const ModesetDev& d, int* order) {
  int idx[BOGP_MAX_TONALS];
  for (int f = 0; f < d.F; ++f) idx[f] = f;
  std::stable_sort(idx, idx + d.F, [&](int a, int b) { return d.t[a].M < d.t[b].M; });
  const char* e = std::getenv("BOGP_UMMA_TONAL_ORDER");
  const bool interleave = !(e && e[0] == '0');
  for (int i = 0, lo = 0, hi = d.F - 1; i < d.F; ++i) order[i] = !interleave ? i : ((i & 1) ? idx[hi--] : idx[lo++]);
}

bool plan16(const bogp_modeset_s* h, HParams& p, int ZC) {
  const ModesetDev& d = h->dev;
  if (d.src_complex || !h->has_cov) return false;
  if ((int)h->u16.wu_max.size() != d.F || (int)h->u16.phi_max.size() != d.F) return false;
  int order[BOGP_MAX_TONALS];
  tonal_order16(d, order);
  unsigned a_off = 0, max_stage = 0, max_a = 0;
  for (int f = 0; f < d.F; ++f) {
    const TonalMeta& t = d.t[order[f]];
    HTonal& u = p.t[f];
    u.src = order[f];
    u.Kp = pad_to(t.M, 16);
    u.rowsP = t.u_rowsP;
    if (u.Kp > 64 || u.rowsP > kNA) return false;
    u.n1 = t.u_p0; u.nq = t.u_fold ? 0 : t.n_qpair; u.trK = t.trK;
    u.fold = t.u_fold; u.c1r = t.u_c[0]; u.c1i = t.u_c[1]; u.c2r = t.u_c[2]; u.c2i = t.u_c[3];
    u.a_bytes = 1024u * (unsigned)u.Kp;
    u.a_off = a_off;
    a_off += u.a_bytes;
    u.b_zbytes = (unsigned)(u.rowsP * u.Kp * 2);
    u.nj = kNA / u.rowsP;
    u.sB = pow2_scale((double)h->u16.wu_max[order[f]] * (double)h->u16.phi_max[order[f]]);
    max_stage = std::max(max_stage, 2u * (unsigned)u.nj * u.b_zbytes);
    max_a = std::max(max_a, u.a_bytes);
  }
  p.F = d.F;
  p.e_rt_stride = a_off;
  // A region: the largest pair of consecutive images (cyclic) when the shared memory allows it, else one image
  unsigned pair = 0;
  for (int f = 0; f < d.F; ++f) pair = std::max(pair, p.t[f].a_bytes + p.t[(f + 1) % d.F].a_bytes);
  const unsigned acc_bytes = (unsigned)ZC * kTile * 4u, bar_bytes = 256u;
  const unsigned avail = kSmemMax - acc_bytes - bar_bytes;
  if (max_a + 2u * max_stage > avail) return false;
  p.AR = (pair + 3u * max_stage <= avail) ? pair : max_a;
  p.S = (int)std::min<unsigned>(kMaxStages, (avail - p.AR) / max_stage);
  if (p.S < 2) return false;
  p.stage_bytes = max_stage;
  p.smem_B = p.AR;
  p.smem_acc = p.smem_B + (unsigned)p.S * max_stage;
  p.smem_bar = p.smem_acc + acc_bytes;
  return true;
}

// Depths per unit.  waves = 1: one unit per SM.  waves = 2 (output in mapped host memory): half-size units, two per SM, so
// that the first half of the surface crosses PCIe while the second half is being computed (all units of a single wave
// finish -- and write -- at the same time).
int choose_zc16(int Nr, int Nz, int sms, int waves) {
  if (const char* e = std::getenv("BOGP_UMMA_ZC")) { const int v = std::atoi(e); if (v >= 8) return std::min(64, v & ~7); }
  const int nrt = (Nr + kTile - 1) / kTile;
  const int chunks = std::max(1, waves * sms / nrt);
  int zc = (Nz + chunks - 1) / chunks;
  zc = std::max(8, std::min(64, (zc + 7) & ~7));
  return zc;
}

int grow(void** ptr, size_t* have, size_t need) {
  if (*have >= need) return 0;
  if (*ptr) cudaFree(*ptr);
  *ptr = nullptr; *have = 0;
  BOGP_CUDA(cudaMalloc(ptr, need + need / 8 + 256));
  *have = need + need / 8 + 256;
  return 0;
}

}  // namespace

bool umma16_supported(const bogp_modeset_s* h) {
  if (const char* e = std::getenv("BOGP_UMMA_KIND")) { if (e[0] == 't') return false; }      // "tf32": the 3xTF32 kernel
  int dev = 0, major = 0;
  if (cudaGetDevice(&dev) != cudaSuccess) return false;
  cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (major != 10 || !h->has_cov) return false;
  HParams p{};
  return plan16(h, p, 64);
}

int launch_grid_umma16(bogp_modeset_s* h, const double* r_host, int Nr, const double* z_host, int Nz, int atype, int mf,
                       float* out_dev, cudaStream_t s, int ext_mode, float* ext_val_dev, long long* ext_idx_dev,
                       const PeerArgs* peer) {
  const ModesetDev& d = h->dev;
  const int sms = device_sms16();
  HParams p{};
  cudaPointerAttributes pat{};
  const bool host_out = cudaPointerGetAttributes(&pat, out_dev) == cudaSuccess && pat.type == cudaMemoryTypeHost;
  if (!host_out) cudaGetLastError();
  int waves = host_out ? 2 : 1;
  if (const char* e = std::getenv("BOGP_UMMA_WAVES")) waves = std::max(1, std::min(4, std::atoi(e)));
  const int ZC = choose_zc16(Nr, Nz, sms, waves);
  if (!plan16(h, p, ZC)) { set_error("UMMA-H engine: mode set does not fit (operator rows > 128, more than 64 modes or complex source table)"); return BOGP_ERR_UNSUPPORTED; }
  p.Nr = Nr; p.Nz = Nz; p.ZC = ZC; p.atype = atype; p.mf = mf;
  { const char* e = std::getenv("BOGP_UMMA_TWO_SLOT"); p.two_slot = e ? std::atoi(e) : 3; }
  p.nrt = (Nr + kTile - 1) / kTile;
  p.units = p.nrt * ((Nz + ZC - 1) / ZC);
  unsigned long long b_off = 0;
  for (int f = 0; f < d.F; ++f) { p.t[f].b_off = b_off; b_off += (unsigned long long)Nz * p.t[f].b_zbytes; }
  // E scale per tonal: |E| <= |k|^-1/2 r^-1/2 e^{Im k r} over the range vector
  double rmin = r_host[0], rmax = r_host[0];
  for (int i = 1; i < Nr; ++i) { rmin = std::min(rmin, r_host[i]); rmax = std::max(rmax, r_host[i]); }
  for (int f = 0; f < d.F; ++f) {
    double bound = 0.0;
    for (const cplx& k : h->u16.k_host[p.t[f].src]) {
      const double a = std::exp(std::max(k.imag() * rmin, k.imag() * rmax)) / std::sqrt(std::abs(k) * rmin);
      bound = std::max(bound, a);
    }
    p.t[f].sE = pow2_scale(bound);
  }
  bogp_modeset_s::U16& c = h->u16;
  const size_t bytes_E = (size_t)p.nrt * p.e_rt_stride;
  const size_t off_z = ((size_t)Nr * sizeof(double) + 255) & ~(size_t)255;
  const size_t off_err = (off_z + (size_t)Nz * sizeof(double) + 255) & ~(size_t)255;
  if (int rc = grow(&c.misc, &c.misc_bytes, off_err + 256)) return rc;
  if (int rc = grow(&c.B, &c.B_bytes, 2 * (size_t)b_off + 512)) return rc;
  const bool e_cached = c.E && c.r_cached.size() == (size_t)Nr && std::memcmp(c.r_cached.data(), r_host, (size_t)Nr * sizeof(double)) == 0 &&
                        !(std::getenv("BOGP_NO_TABLE_CACHE") && std::getenv("BOGP_NO_TABLE_CACHE")[0] == '1');
  if (!e_cached) {
    c.r_cached.clear();
    c.z_cached.clear();
    if (int rc = grow(&c.E, &c.E_bytes, bytes_E + 256)) return rc;
  }
  unsigned char* misc = static_cast<unsigned char*>(c.misc);
  double* r_dev = reinterpret_cast<double*>(misc);
  double* z_dev = reinterpret_cast<double*>(misc + off_z);
  p.err = reinterpret_cast<int*>(misc + off_err);
  if (ext_mode && (long long)Nr * Nz < (1ll << 32)) {
    p.ext_mode = ext_mode;
    p.ext_key = reinterpret_cast<unsigned long long*>(misc + off_err + 16);
    p.ext_done = reinterpret_cast<unsigned int*>(misc + off_err + 32);
    p.ext_val = ext_val_dev; p.ext_idx = ext_idx_dev;
    if (peer && peer->bufs) {
      p.peer_bufs = peer->bufs; p.peer_rank = peer->rank; p.peer_world = peer->world; p.peer_epoch = peer->epoch;
      p.peer_offset = peer->index_offset; p.peer_out = peer->recs_out; p.peer_err = peer->err; p.peer_mode = peer->pipelined;
    }
  }
  p.E = static_cast<unsigned char*>(c.E);
  p.Bhi = static_cast<unsigned char*>(c.B);
  p.Blo = p.Bhi + (((size_t)b_off + 255) & ~(size_t)255);
  p.out = out_dev;
  if (!e_cached) {
    BOGP_CUDA(h2d_async(r_dev, r_host, Nr * sizeof(double), s));
    umma16_E_kernel<<<d.F * p.nrt * (kTile / kETabRows), 128, 0, s>>>(d, p, r_dev);
    count_launch();
    BOGP_CUDA(cudaGetLastError());
    c.r_cached.assign(r_host, r_host + Nr);
  }
  // the depth vector stays on the device while it is the same (one grid for all time steps of a sweep)
  if (!(e_cached && c.z_cached.size() == (size_t)Nz && std::memcmp(c.z_cached.data(), z_host, (size_t)Nz * sizeof(double)) == 0)) {
    BOGP_CUDA(h2d_async(z_dev, z_host, Nz * sizeof(double), s));
    c.z_cached.assign(z_host, z_host + Nz);
    c.phi_valid = false;
  }
  if (BOGP_UMMA_DBG) { const char* sk = std::getenv("BOGP_UMMA_SKIP"); p.skip = sk ? std::atoi(sk) : 0; }
  // lean instance: every tonal folded (rank-1 covariance, real modes) with 16 or 32 padded rows, cbf + mean
  bool lean = atype == BOGP_ATYPE_CBF && mf == BOGP_MF_MEAN && !(BOGP_UMMA_DBG && (p.skip & 2));
  for (int f = 0; f < d.F; ++f) lean = lean && p.t[f].fold && (p.t[f].rowsP == 16 || p.t[f].rowsP == 32);
  if (const char* e = std::getenv("BOGP_UMMA_LEAN")) lean = lean && e[0] != '0';
  // Streamed tables: see umma16_B_stream_kernel.  Only when the two launches are chained programmatically (no timing
  // event between them) and with the lean instance of the surface kernel: 88 registers x 640 threads leave room for
  // the table block on the same SM, the general instance (96 allocated) does not.
  const char* pdl_env = std::getenv("BOGP_UMMA_PDL");
  const char* ts_env = std::getenv("BOGP_UMMA_TABSTREAM");
  const bool chained = !timing_enabled() && !(pdl_env && pdl_env[0] == '0') && !(BOGP_UMMA_DBG && std::getenv("BOGP_UMMA_DBG"));
  static int fits_next_to_tables = -1;      // the compiled register counts decide whether the two kernels share an SM
  if (fits_next_to_tables < 0) {
    cudaFuncAttributes fa{}, fb{};
    fits_next_to_tables = cudaFuncGetAttributes(&fa, umma16_grid_kernel_lean) == cudaSuccess &&
                          cudaFuncGetAttributes(&fb, umma16_B_stream_kernel) == cudaSuccess &&
                          ((fa.numRegs + 7) & ~7) * kThr + ((fb.numRegs + 7) & ~7) * kBThreads <= 65536;
  }
  const bool stream_tabs = chained && lean && fits_next_to_tables == 1 && !(ts_env && ts_env[0] == '0');
  const int nzb = (Nz + kBDepths - 1) / kBDepths;
  if (stream_tabs) {
    if (!c.tab_cnt) {
      BOGP_CUDA(cudaMalloc(&c.tab_cnt, BOGP_MAX_TONALS * sizeof(unsigned int)));
      BOGP_CUDA(cudaMemsetAsync(c.tab_cnt, 0, BOGP_MAX_TONALS * sizeof(unsigned int), s));
      std::memset(c.tab_target, 0, sizeof(c.tab_target));
    }
    for (int f = 0; f < d.F; ++f) { c.tab_target[f] += (unsigned int)nzb * (kBThreads / 32); p.tab_target[f] = c.tab_target[f]; }
    p.tab_cnt = c.tab_cnt;
    size_t phi_n = 0;
    for (int f = 0; f < d.F; ++f) { p.phi_off[f] = (unsigned int)phi_n; phi_n += (size_t)Nz * p.t[f].Kp; }
    if (c.phi_bytes < phi_n * sizeof(float)) c.phi_valid = false;
    if (int rc = grow(&c.phi, &c.phi_bytes, phi_n * sizeof(float))) return rc;
    p.phi = static_cast<const float*>(c.phi);
    if (!c.phi_valid) {
      umma16_phi_kernel<<<dim3((unsigned)std::min(64, (Nz * 64 + 255) / 256), (unsigned)d.F), 256, 0, s>>>(d, p, z_dev, static_cast<float*>(c.phi));
      count_launch();
      BOGP_CUDA(cudaGetLastError());
      c.phi_valid = true;
    }
    // the SM's shared-memory carve-out cannot change while a block is resident: ask for the largest one, or the
    // surface kernel's 217 KB block would have to wait for the table block to leave
    static bool carve_set = false;
    if (!carve_set) {
      BOGP_CUDA(cudaFuncSetAttribute(umma16_B_stream_kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
      carve_set = true;
    }
    umma16_B_stream_kernel<<<std::min(sms, d.F * nzb), kBThreads, 0, s>>>(d, p, c.tab_cnt);
  } else {
    umma16_B_kernel<<<d.F * nzb, kBThreads, 0, s>>>(d, p, z_dev);
  }
  count_launch();
  if (cudaError_t e = cudaGetLastError(); e != cudaSuccess) {
    if (stream_tabs) for (int f = 0; f < d.F; ++f) c.tab_target[f] -= (unsigned int)nzb * (kBThreads / 32);      // the counters did not move
    set_error("UMMA-H engine: table kernel launch: %s", cudaGetErrorString(e));
    return BOGP_ERR_CUDA;
  }
  static long long* dbg_dev = nullptr;
  const char* dbg_env = BOGP_UMMA_DBG ? std::getenv("BOGP_UMMA_DBG") : nullptr;
  if (dbg_env && dbg_env[0] == '1') {
    if (!dbg_dev) BOGP_CUDA(cudaMalloc(&dbg_dev, 1024 * sizeof(long long)));
    BOGP_CUDA(cudaMemsetAsync(dbg_dev, 0, 1024 * sizeof(long long), s));
    p.dbg = dbg_dev;
  }
  const size_t smem = (size_t)p.smem_bar + 256;
  auto kern = lean ? umma16_grid_kernel_lean : umma16_grid_kernel_general;
  BOGP_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  cudaEvent_t ev = timing_begin(s);
  if (!ev && !p.dbg && !(pdl_env && pdl_env[0] == '0')) {
    // no event between the two launches: let the surface kernel's prologue overlap the table kernel
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3((unsigned)std::min(p.units, sms)); cfg.blockDim = dim3(kThr); cfg.dynamicSmemBytes = smem; cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at; cfg.numAttrs = 1;
    BOGP_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  } else {
    kern<<<std::min(p.units, sms), kThr, smem, s>>>(p);
  }
  timing_end(ev, s);
  count_launch();
  BOGP_CUDA(cudaGetLastError());
  if (p.dbg) {
    static long long hd[1024];
    BOGP_CUDA(cudaMemcpyAsync(hd, dbg_dev, sizeof(hd), cudaMemcpyDeviceToHost, s));
    BOGP_CUDA(cudaStreamSynchronize(s));
    const int nzu = std::min(ZC, Nz);
    fprintf(stderr, "[umma16 dbg] S=%d AR=%u stage=%u ZC=%d units=%d; per tonal of unit 0 (CTA 0): Kp rowsP stages | MMA cycles | issuer | epilogue\n", p.S, p.AR, p.stage_bytes, ZC, p.units);
    long long tot_mma = 0;
    for (int f = 0; f < d.F; ++f) {
      const HTonal& u = p.t[f];
      const int stages = (nzu + u.nj - 1) / u.nj;
      const long long mma = (long long)6 * (u.Kp / 16) * (long long)(nzu * u.rowsP) / 2;
      tot_mma += mma;
      const long long iss = (f + 1 < d.F ? hd[f + 1] : hd[d.F]) - hd[f], epi = (f + 1 < d.F ? hd[64 + f + 1] : hd[127]) - hd[64 + f];
      fprintf(stderr, "[umma16 dbg]   %2d %3d %3d | %6lld | %7lld | %7lld\n", u.Kp, u.rowsP, stages, mma, iss, epi);
    }
    fprintf(stderr, "[umma16 dbg] stage | ready->issued | issued->wake | wake->release(w0) | release->done(w0) | w15 release - w0 release | release(w0)->ready(sc+2) | ready delta\n");
    {
      const long long* e = hd + 128;
      int sc = 0;
      for (int f = 0; f < d.F && sc < 126; ++f) {
        const int stages = (nzu + p.t[f].nj - 1) / p.t[f].nj;
        for (int k = 0; k < stages && sc < 126; ++k, ++sc)
          if (k < 6) fprintf(stderr, "   %3d (Kp %2d rows %2d)  %6lld %6lld %6lld %6lld %6lld %6lld   %6lld\n", sc, p.t[f].Kp, p.t[f].rowsP, e[128 + sc] - e[sc], e[256 + sc] - e[128 + sc],
                  e[384 + sc] - e[256 + sc], e[512 + sc] - e[384 + sc], e[640 + sc] - e[384 + sc], e[sc + 2] - e[384 + sc], e[sc + 1] - e[sc]);
      }
    }
    fprintf(stderr, "[umma16 dbg] unit 0: MMA %lld cycles, issuer first tonal -> end of kernel %lld\n", tot_mma, hd[63] - hd[0]);
  }
  return 0;
}

}  // namespace bogp
