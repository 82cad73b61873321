// Inline-PTX wrappers for sm_100a: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (alloc/mma/commit/ld).
// No CUTLASS dependency; field layouts follow the PTX ISA "tcgen05 shared memory descriptor" and
// "instruction descriptor" tables (cross-checked against cute/arch/mma_sm100_desc.hpp in the image).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

namespace tgis {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ uint32_t lane_id() { return threadIdx.x & 31; }

__device__ __forceinline__ bool elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .pred P;\n\t.reg .b32 R;\n\t"
      "elect.sync R|P, 0xffffffff;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(pred));
  return pred != 0;
}

// ---------------------------------------------------------------- mbarrier
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;\n" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;\n" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];\n" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(smem_u32(bar)),
               "r"(bytes)
               : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred P;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 P, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, P;\n\t}\n"
      : "=r"(ok)
      : "r"(smem_u32(bar)), "r"(parity)
      : "memory");
  return ok != 0;
}
// Bounded spin: a protocol bug traps (-> sticky CUDA error surfaced to the host) instead of hanging the GPU box.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
#ifdef TGIS_MBAR_TIMEOUT
  const long long t0 = clock64();
  for (;;) {
    if (mbar_try_wait(bar, parity)) return;
    if (clock64() - t0 > 4000000000ll) break;  // ~2 s at 2 GHz: no legitimate wait in this code base is that long
  }
  printf("mbar_wait timeout block %d thread %d\n", blockIdx.x, threadIdx.x);
  __trap();
#else
  while (!mbar_try_wait(bar, parity)) {
  }
#endif
}

// ---------------------------------------------------------------- TMA
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* m) {
  asm volatile("prefetch.tensormap [%0];\n" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
// 2D tiled load global->smem, completes on mbarrier (SASS: UTMALDG)
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d_hint(const CUtensorMap* m, uint64_t* bar, void* dst, int c0, int c1,
                                                 uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;\n" ::"r"(smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "l"(policy)
      : "memory");
}
// 2D tile prefetch into L2 only (SASS: UTMAPF.L2): no smem destination, no completion tracking
__device__ __forceinline__ void tma_prefetch_2d(const CUtensorMap* m, int c0, int c1) {
  asm volatile("cp.async.bulk.prefetch.tensor.2d.L2.global.tile [%0, {%1, %2}];\n" ::"l"(
                   reinterpret_cast<uint64_t>(m)),
               "r"(c0), "r"(c1)
               : "memory");
}
// 1D bulk copy global->smem (SASS: UBLKCP); size multiple of 16, both addresses 16B aligned
__device__ __forceinline__ void bulk_load_1d(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar))
      : "memory");
}
// 1D bulk copy with an L2 cache policy (streamed-once data: evict_first)
__device__ __forceinline__ void bulk_load_1d_hint(void* dst, const void* src, uint32_t bytes, uint64_t* bar,
                                                  uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;\n" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(src)), "r"(bytes), "r"(smem_u32(bar)), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;\n" : "=l"(p));
  return p;
}

// ---------------------------------------------------------------- tcgen05 / TMEM
template <uint32_t kCols>
__device__ __forceinline__ void tmem_alloc(uint32_t* smem_dst) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;\n" ::"r"(smem_u32(smem_dst)),
               "n"(kCols));
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;\n");
}
template <uint32_t kCols>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;\n" ::"r"(taddr), "n"(kCols));
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;\n" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;\n" ::: "memory"); }

// tcgen05.commit: arrive on an mbarrier once all previously issued MMAs of this thread retire
// (implies tcgen05.fence::before_thread_sync). SASS: UTCBAR
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];\n" ::"r"(
                   smem_u32(bar))
               : "memory");
}

// D[tmem] (+)= A[smem] * B[smem], kind::f16 (bf16/fp16 in, fp32 accumulate). SASS: UTCHMMA
__device__ __forceinline__ void tc_mma_f16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                           uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}\n" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// K-major operand tile in smem, rows of 64 bf16 (=128 B) under the 128-byte swizzle TMA wrote it with:
// 8-row groups are 1024 B apart (SBO), LBO unused for swizzled K-major, descriptor version 1 (sm_100),
// layout type 2 = SWIZZLE_128B. The tile base must be 1024-B aligned (base_offset = 0).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);  // start address  [0,14)
  d |= (uint64_t)1 << 16;                       // LBO (ignored)  [16,30)
  d |= (uint64_t)(1024 >> 4) << 32;             // SBO = 1024 B   [32,46)
  d |= (uint64_t)1 << 46;                       // version = 1    [46,48)
  d |= (uint64_t)2 << 61;                       // SWIZZLE_128B   [61,64)
  return d;
}
// Instruction descriptor, kind::f16: C=F32, A=B=BF16, both K-major, dense, no negate.
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t umma_m, uint32_t umma_n) {
  return (1u << 4)                // c_format = F32
         | (1u << 7)              // a_format = BF16
         | (1u << 10)             // b_format = BF16
         | ((umma_n >> 3) << 17)  // n_dim
         | ((umma_m >> 4) << 24); // m_dim
}

// TMEM -> registers: 32 lanes (this warp's sub-partition) x N consecutive 32-bit columns. SASS: LDTM
__device__ __forceinline__ void tmem_ld_32x32b_x16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];\n"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;\n" ::: "memory"); }

// ---------------------------------------------------------------- thread-block cluster / distributed shared memory
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;\n" : "=r"(r));
  return r;
}
// all threads of all CTAs of the cluster; release/acquire orders shared-memory writes before peers' DSMEM reads
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}
__device__ __forceinline__ uint32_t dsmem_addr(uint32_t local_smem_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;\n" : "=r"(r) : "r"(local_smem_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ float4 ld_dsmem_f4(uint32_t addr) {
  float4 v;
  asm volatile("ld.shared::cluster.v4.f32 {%0,%1,%2,%3}, [%4];\n"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w)
               : "r"(addr)
               : "memory");
  return v;
}

// ---------------------------------------------------------------- misc
__device__ __forceinline__ float bf16_bits_to_f32(uint16_t b) { return __uint_as_float(((uint32_t)b) << 16); }
__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ void griddep_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void griddep_launch() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

// ---------------------------------------------------------------- in-situ step timeline (debug builds only)
// -DTGIS_STEP_TIMELINE: CTA 0 / thread 0 of every kernel launch appends {kernel id, %globaltimer at entry, after the
// grid-dependency wait, at exit} to a device buffer (scripts/step_timeline.py) -- shows, inside CUDA-graph replays with
// PDL, where a decode step's time goes.  Release builds compile the marks away.
#ifdef TGIS_STEP_TIMELINE
__device__ __forceinline__ unsigned long long stl_timer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  return t;
}
#define TGIS_STL_DEFINE(prefix)                                                                       \
  static __device__ unsigned long long* g_stl = nullptr;                                              \
  int prefix##_set_step_timeline(unsigned long long* p) {                                             \
    return cudaMemcpyToSymbol(g_stl, &p, sizeof(p)) == cudaSuccess ? 0 : -1;                          \
  }
#define STL_ENTER(kid)                                                                                \
  int _stl = -1;                                                                                      \
  if (g_stl != nullptr && blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) {                   \
    _stl = (int)(atomicAdd(g_stl, 1ull) & 4095ull);                                                   \
    unsigned long long* _r = g_stl + 8 + _stl * 4;                                                    \
    _r[0] = (unsigned long long)(kid);                                                                \
    _r[1] = stl_timer();                                                                              \
    _r[2] = 0;                                                                                        \
    _r[3] = 0;                                                                                        \
    unsigned long long* _x = g_stl + 8 + 4096 * 4 + _stl * 4;                                         \
    _x[0] = _x[1] = _x[2] = _x[3] = 0;                                                                \
  }
// header word [1] = latest exit stamp of ANY warp of any instrumented kernel; the next kernel's CTA 0 stores
// (its wait-return time - that stamp) in bits 32..51 of its record's word 0, so the "gap" after a kernel splits into
// the kernel's own tail (last warp exit - CTA 0 exit) and the dependency-resolution latency.
#define STL_WAITED() do { if (_stl >= 0) {                                                            \
    const unsigned long long _t = stl_timer();                                                        \
    g_stl[8 + _stl * 4 + 2] = _t;                                                                     \
    const unsigned long long _last = *((volatile unsigned long long*)(g_stl + 1));                    \
    g_stl[8 + _stl * 4 + 0] |= ((_t > _last ? _t - _last : 0ull) & 0xfffffull) << 32; } } while (0)
#define STL_EXIT() do {                                                                               \
    if (g_stl != nullptr && (threadIdx.x & 31) == 0) atomicMax(g_stl + 1, stl_timer());               \
    if (_stl >= 0) g_stl[8 + _stl * 4 + 3] = stl_timer(); } while (0)
// extra stamps of the instrumented CTA, second half of the buffer: [8 + 4096 * 4 + record * 4 + k].  Other warps learn the
// record index through a shared-memory word (STL_SHARE by thread 0 before a CTA barrier).
#define STL_SHARE(smem_int_ptr) do { if (threadIdx.x == 0) *(smem_int_ptr) = _stl; } while (0)
#define STL_EXTRA(rec, k) do { const int _rc = (rec);                                                 \
    if (g_stl != nullptr && _rc >= 0) g_stl[8 + 4096 * 4 + _rc * 4 + (k)] = stl_timer(); } while (0)
#define STL_MINE() (_stl)
#else
#define TGIS_STL_DEFINE(prefix) \
  int prefix##_set_step_timeline(unsigned long long*) { return -2; }
#define STL_ENTER(kid) do {} while (0)
#define STL_WAITED() do {} while (0)
#define STL_EXIT() do {} while (0)
#define STL_SHARE(smem_int_ptr) do {} while (0)
#define STL_EXTRA(rec, k) do {} while (0)
#define STL_MINE() (-1)
#endif

}  // namespace tgis
