// PTX wrappers shared by the tcgen05 (UMMA) kernels: mbarriers, bulk copies, tcgen05.mma / ld / st / commit,
// packed fp32x2 arithmetic, shared-memory matrix descriptors.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

namespace bogp {

// ------------------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
// Bounded wait: a protocol bug must trap (an error the host sees), never hang the GPU.
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity, int* err, int code) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    if (spin == 0) t0 = clock64();
    if ((spin & 1023u) == 1023u && clock64() - t0 > 3000000000LL) {
      atomicExch(err, code + 1000 * (int)blockIdx.x);
      __threadfence_system();
      __trap();
    }
  }
}
// One lane of a converged warp (elect.sync): the tcgen05 / bulk-copy instructions take warp-uniform operands, so the
// issuing roles run their loops with the whole warp and predicate only the instruction itself -- issuing from inside
// an `if (lane == 0)` region makes ptxas wrap every UTCHMMA in a per-lane waterfall loop (~90 cycles per MMA).
__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, px;\n\t}"
      : "=r"(pred));
  return pred != 0;
}
// Waiting flavour for the roles that run ahead of the pipeline (producer, loaders): back off between polls so the
// spinning warp does not take issue slots from the epilogue warps on its scheduler (higher warp ids win arbitration).
template <unsigned kSleepNs = 256>
__device__ __forceinline__ void mbar_wait_relaxed(uint32_t bar, uint32_t parity, int* err, int code) {
  uint32_t ok = 0;
  long long t0 = 0;
  for (uint32_t spin = 0;; ++spin) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
    if (ok) return;
    __nanosleep(kSleepNs);
    if (spin == 0) t0 = clock64();
    if ((spin & 255u) == 255u && clock64() - t0 > 3000000000LL) {
      atomicExch(err, code + 1000 * (int)blockIdx.x);
      __threadfence_system();
      __trap();
    }
  }
}
// mbar_wait that also accumulates the waiting time into *acc when profiling is on
__device__ __forceinline__ void mbar_wait_t(uint32_t bar, uint32_t parity, int* err, int code, long long* acc) {
  if (acc) {
    const long long t0 = clock64();
    mbar_wait(bar, parity, err, code);
    *acc += clock64() - t0;
  } else {
    mbar_wait(bar, parity, err, code);
  }
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
               "l"(src), "r"(bytes), "r"(bar)
               : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc,
                                               uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// Same with the shared-memory descriptor passed as two 32-bit halves (only the low half changes between K steps).
__device__ __forceinline__ void tc_mma_tf32_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_desc_lo, uint32_t b_desc_hi,
                                                uint32_t idesc, bool accumulate) {
  if (accumulate)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 d;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "mov.b64 d, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], d, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_desc_lo), "r"(b_desc_hi), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 d;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "mov.b64 d, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], d, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_desc_lo), "r"(b_desc_hi), "r"(idesc)
        : "memory");
}
// kind::f16 with A in TMEM (two FP16 values per 32-bit column, K = 16 per instruction = 8 columns): same operand
// addressing as the TF32 form above.
__device__ __forceinline__ void tc_mma_f16_ts2(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_desc_lo, uint32_t b_desc_hi,
                                               uint32_t idesc, bool accumulate) {
  if (accumulate)
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 d;\n\t"
        "setp.eq.b32 p, 0, 0;\n\t"
        "mov.b64 d, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], d, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_desc_lo), "r"(b_desc_hi), "r"(idesc)
        : "memory");
  else
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b64 d;\n\t"
        "setp.ne.b32 p, 0, 0;\n\t"
        "mov.b64 d, {%2, %3};\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], d, %4, p;\n\t}" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_desc_lo), "r"(b_desc_hi), "r"(idesc)
        : "memory");
}
// All MMAs of one stage, fully unrolled over the KS K-steps: per K-step D += A_lo B_hi + A_hi B_lo + A_hi B_hi for
// the re and the im accumulator.  Called by ONE elected lane; a single warp must keep the tensor pipe fed with
// 128 x N x 8 MMAs of only N/2 cycles each, so the instruction count per MMA matters (v1 spent ~90 cycles per MMA
// in address arithmetic and R2UR moves).
template <int KS, bool F16 = false>
__device__ __forceinline__ void issue_stage(uint32_t d_re, uint32_t d_im, uint32_t a_col, uint32_t Mp, uint32_t bh_lo,
                                            uint32_t bl_lo, uint32_t desc_hi, uint32_t idesc) {
  // Opaque copies: keeps ptxas from hoisting the 4 KS derived addresses out of the stage loop into vector registers
  // (it then re-moves every one of them to a uniform register per stage); derived here they stay in the uniform path.
  asm volatile("mov.u32 %0, %0;" : "+r"(a_col));
  asm volatile("mov.u32 %0, %0;" : "+r"(bh_lo));
  asm volatile("mov.u32 %0, %0;" : "+r"(bl_lo));
#pragma unroll
  for (int part = 0; part < 2; ++part) {
    const uint32_t d = part ? d_im : d_re;
    const uint32_t a_hi = a_col + (uint32_t)(2 * part) * Mp, a_lo = a_hi + Mp;
#pragma unroll
    for (int ks = 0; ks < KS; ++ks) {
      if (F16) {
        tc_mma_f16_ts2(d, a_lo + 8u * ks, bh_lo + 16u * ks, desc_hi, idesc, ks > 0);
        tc_mma_f16_ts2(d, a_hi + 8u * ks, bl_lo + 16u * ks, desc_hi, idesc, true);
        tc_mma_f16_ts2(d, a_hi + 8u * ks, bh_lo + 16u * ks, desc_hi, idesc, true);
      } else {
        tc_mma_tf32_ts2(d, a_lo + 8u * ks, bh_lo + 16u * ks, desc_hi, idesc, ks > 0);
        tc_mma_tf32_ts2(d, a_hi + 8u * ks, bl_lo + 16u * ks, desc_hi, idesc, true);
        tc_mma_tf32_ts2(d, a_hi + 8u * ks, bh_lo + 16u * ks, desc_hi, idesc, true);
      }
    }
  }
}
// Packed fp32x2 FMA (FFMA2 on sm_100): acc.{lo,hi} += {a,b}^2 -- halves the epilogue's dominant instruction count.
__device__ __forceinline__ void sq_acc2(unsigned long long& acc, float a, float b) {
  unsigned long long v;
  asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(a), "f"(b));
  asm("fma.rn.f32x2 %0, %1, %1, %0;" : "+l"(acc) : "l"(v));
}
__device__ __forceinline__ float sum2(unsigned long long acc) {
  float x, y;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(x), "=f"(y) : "l"(acc));
  return x + y;
}
// Numerator pairs live in the first 2 nq <= 16 columns: T = (r0 + i i0) + i (r1 + i i1) per (re, im) row pair.
__device__ __forceinline__ void numerator16(const float* re, const float* im, int nq, float& num, float& numsq) {
#pragma unroll
  for (int q = 0; q < 8; ++q) {
    if (q < nq) {
      const float r0 = re[2 * q], r1 = re[2 * q + 1], i0 = im[2 * q], i1 = im[2 * q + 1];
      const float tr = r0 - i1, ti = i0 + r1;
      num = fmaf(tr, tr, fmaf(ti, ti, num));
      numsq = fmaf(r0, r0, fmaf(i0, i0, fmaf(r1, r1, fmaf(i1, i1, numsq))));
    }
  }
}
__device__ __forceinline__ void tc_st16(uint32_t taddr, const float4& a, const float4& b, const float4& c, const float4& d) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16};" ::"r"(taddr),
      "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)), "r"(__float_as_uint(a.w)),
      "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)), "r"(__float_as_uint(b.z)), "r"(__float_as_uint(b.w)),
      "r"(__float_as_uint(c.x)), "r"(__float_as_uint(c.y)), "r"(__float_as_uint(c.z)), "r"(__float_as_uint(c.w)),
      "r"(__float_as_uint(d.x)), "r"(__float_as_uint(d.y)), "r"(__float_as_uint(d.z)), "r"(__float_as_uint(d.w))
      : "memory");
}
__device__ __forceinline__ void tc_st8(uint32_t taddr, const float* v) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
               "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
               "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7]))
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld16(uint32_t taddr, float* v) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,"
      "%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tc_ld2(uint32_t taddr, float& a, float& b) {
  uint32_t r0, r1;
  asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(r0), "=r"(r1) : "r"(taddr));
  a = __uint_as_float(r0);
  b = __uint_as_float(r1);
}
// Packed fp32x2 arithmetic (sm_100 FADD2 / FMUL2 / FFMA2): two receivers per instruction in the tilt epilogue.
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pk(float lo, float hi) { f32x2 v; asm("mov.b64 %0, {%1, %2};" : "=l"(v) : "f"(lo), "f"(hi)); return v; }
__device__ __forceinline__ void upk(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 sub2(f32x2 a, f32x2 b) { f32x2 d; asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }

// MUFU.RSQ alone: rsqrtf() adds a denormal-range rescue (FSETP + two predicated FMULs) that 1 - l sin(tau) / r ~ 1 never needs
__device__ __forceinline__ float rsqrt_fast(float x) {
  float y;
  asm("rsqrt.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor, K-major, no swizzle (cute::UMMA::SmemDescriptor bit layout): start address
// [0,14), leading byte offset [16,30), stride byte offset [32,46) -- all in 16-byte units -- version 1 at [46,48).
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3fffu) | ((uint64_t)((lbo_bytes >> 4) & 0x3fffu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3fffu) << 32) | (1ull << 46);
}

}  // namespace bogp
