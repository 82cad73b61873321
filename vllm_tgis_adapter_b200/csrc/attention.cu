// Paged-KV attention for the TGIS hot path (reference call site: grpc_server.py:222 ->
// vllm v1/attention/backends/flashinfer.py:1665,1803).  Causal, GQA, head_dim 128, bf16 in / fp32 softmax.
//
// HBM layout of the cache (ours to choose; see kernels.h): per (block, kv_head) one contiguous 8 KiB K tile stored
// "8-dim chunk major" [16][32 tok][8] and one 8 KiB V tile [32 tok][16 chunks ^ (tok&7)][8].  Both tiles are fetched
// with ONE 1-D TMA bulk copy each (cp.async.bulk -> UBLKCP) completing on an mbarrier, and both are bank-conflict
// free for the two consumers below without any in-kernel transposition:
//   decode : QK^T with lane = token (16-B LDS per 8-dim chunk), PV with lane = 4 dims (8-B LDS per token)
//   prefill: QK^T B-fragments are plain 4-B LDS, PV B-fragments are ldmatrix.trans rows spread over 8 bank groups
//
// decode  (q_len = 1): grid (seq x split, kv_head); split-KV in fixed 128-token chunks (batch-invariant numerics);
//                      one warp per 32-token block; the G = n_q/n_kv query heads of the group share every K/V byte;
//                      last-arriving split combines partials in split order (deterministic).
// prefill (q_len >= 1): grid (16-query tile, kv_head); one warp per query head of the group; mma.sync m16n8k16 with
//                      online softmax; KV blocks double-buffered through smem by TMA bulk copies.
#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace tgis {

constexpr int DEC_TOK = DECODE_SPLIT;        // tokens per split (8 KV blocks streamed through a 4-stage ring)
constexpr int DEC_BLOCKS = DEC_TOK / KV_BLOCK;
constexpr int TILE_BYTES = KV_BLOCK * HEAD_DIM * 2;  // 8192
static_assert(DEC_TOK == DECODE_SPLIT, "split size mismatch");

__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
__device__ __forceinline__ float warp_add(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ================================================================================================ decode
// Streaming flash-decoding.  One CTA = (sequence, 256-token split, kv head): a producer warp feeds a 4-stage ring of
// KV blocks (one 8 KiB K tile + one 8 KiB V tile per stage, two 1-D TMA bulk copies completing on the stage's
// mbarrier); four compute warps consume it: warps {0,1} take even blocks, {2,3} odd blocks, and inside a pair the
// G query heads of the group are split between the two warps.  Every warp keeps an online-softmax state (m, l, o)
// across its blocks, the two block lanes are combined through shared memory, and splits (kv_len > 256) are merged by
// the last-arriving CTA in split order.  The split size is a constant, so the arithmetic of a sequence depends only
// on its own length, never on what else is in the batch.
constexpr int DEC_STAGES = 4;
constexpr int DEC_THREADS = 160;  // 4 compute warps + 1 producer warp

template <int G>
__global__ void __launch_bounds__(DEC_THREADS)
attn_decode_kernel(const __nv_bfloat16* __restrict__ qkv, int qkv_ld, const __nv_bfloat16* __restrict__ k_cache,
                   const __nv_bfloat16* __restrict__ v_cache, const AttnSeq* __restrict__ seqs,
                   const int32_t* __restrict__ seq_ids, const int32_t* __restrict__ block_table, int bt_stride,
                   int max_splits, float* __restrict__ part_o, float* __restrict__ part_ml,
                   int* __restrict__ counters, __nv_bfloat16* __restrict__ out, int out_ld, int n_kv, float scale) {
  extern __shared__ __align__(128) uint8_t smem[];
  griddep_launch();
  griddep_wait();
  constexpr int GH = (G + 1) / 2;  // heads per warp of a pair
  uint8_t* kv_s = smem;                                                        // DEC_STAGES x (K 8K | V 8K)
  float* q_s = reinterpret_cast<float*>(smem + DEC_STAGES * 2 * TILE_BYTES);   // [G][128]
  float* p_s = q_s + G * HEAD_DIM;                                             // [4 warps][GH][32]
  float* ml_s = p_s + 4 * GH * 32;                                             // [2 lanes][G][2]
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(ml_s + 2 * G * 2);          // [DEC_STAGES]
  uint64_t* empty_bar = full_bar + DEC_STAGES;                                 // [DEC_STAGES]
  int* flag_s = reinterpret_cast<int*>(empty_bar + DEC_STAGES);

  const int sidx = blockIdx.x / max_splits, split = blockIdx.x % max_splits;
  const int kvh = blockIdx.y;
  const AttnSeq sq = seqs[seq_ids ? seq_ids[sidx] : sidx];  // engine: decode sequences are entries 0..n-1
  const int kv_len = sq.kv_len;
  const int n_splits = (kv_len + DEC_TOK - 1) / DEC_TOK;
  if (split >= n_splits) return;
  const int tok0 = split * DEC_TOK;
  const int n_tok = min(DEC_TOK, kv_len - tok0);
  const int n_blk = (n_tok + KV_BLOCK - 1) / KV_BLOCK;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  if (threadIdx.x == 0) {
    for (int i = 0; i < DEC_STAGES; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 2);  // the two warps of the pair that consumed the block
    }
    fence_barrier_init();
  }
  __syncthreads();

  if (warp == 4) {
    // ===================== producer =====================
    if (lane == 0) {
      const int32_t* bt = block_table + (size_t)sq.block_row * bt_stride + tok0 / KV_BLOCK;
      for (int j = 0; j < n_blk; ++j) {
        const int st = j & (DEC_STAGES - 1);
        if (j >= DEC_STAGES) mbar_wait(&empty_bar[st], ((j / DEC_STAGES) - 1) & 1);
        const size_t tile = ((size_t)bt[j] * n_kv + kvh) * (KV_BLOCK * HEAD_DIM);
        mbar_arrive_expect_tx(&full_bar[st], 2 * TILE_BYTES);
        bulk_load_1d(kv_s + st * 2 * TILE_BYTES, k_cache + tile, TILE_BYTES, &full_bar[st]);
        bulk_load_1d(kv_s + st * 2 * TILE_BYTES + TILE_BYTES, v_cache + tile, TILE_BYTES, &full_bar[st]);
      }
    }
    return;
  }

  // ===================== compute warps (128 threads) =====================
  {
    const __nv_bfloat16* q = qkv + (size_t)sq.q_start * qkv_ld + (size_t)kvh * G * HEAD_DIM;
    for (int i = threadIdx.x; i < G * HEAD_DIM; i += 128) q_s[i] = __bfloat162float(q[i]);
  }
  asm volatile("bar.sync 1, 128;\n" ::: "memory");

  const int blane = warp >> 1, half = warp & 1;   // block lane (even / odd blocks), head half
  const int g0 = half * GH;
  const int gh = half == 0 ? GH : G - GH;         // heads of this warp (0 when G == 1 and half == 1)
  const float* q_w = q_s + g0 * HEAD_DIM;
  float* p_w = p_s + warp * GH * 32;
  float m_w[GH], l_w[GH], o_w[GH][4];
#pragma unroll
  for (int g = 0; g < GH; ++g) {
    m_w[g] = -INFINITY;
    l_w[g] = 0.f;
#pragma unroll
    for (int e = 0; e < 4; ++e) o_w[g][e] = 0.f;
  }
  const int lchunk = lane >> 1, lhalf = lane & 1;

  for (int j = blane; j < n_blk; j += 2) {
    const int st = j & (DEC_STAGES - 1);
    mbar_wait(&full_bar[st], (j / DEC_STAGES) & 1);
    if (gh > 0) {
      const int valid = min(KV_BLOCK, n_tok - j * KV_BLOCK);
      const uint8_t* k_t = kv_s + st * 2 * TILE_BYTES;
      const uint8_t* v_t = k_t + TILE_BYTES;
      // ---- scores: lane = token
      float s[GH];
#pragma unroll
      for (int g = 0; g < GH; ++g) s[g] = 0.f;
#pragma unroll 4
      for (int c = 0; c < HEAD_DIM / 8; ++c) {
        const uint4 kk = *reinterpret_cast<const uint4*>(k_t + (c * KV_BLOCK + lane) * 16);
        float kf[8];
        kf[0] = __uint_as_float(kk.x << 16); kf[1] = __uint_as_float(kk.x & 0xffff0000u);
        kf[2] = __uint_as_float(kk.y << 16); kf[3] = __uint_as_float(kk.y & 0xffff0000u);
        kf[4] = __uint_as_float(kk.z << 16); kf[5] = __uint_as_float(kk.z & 0xffff0000u);
        kf[6] = __uint_as_float(kk.w << 16); kf[7] = __uint_as_float(kk.w & 0xffff0000u);
#pragma unroll
        for (int g = 0; g < GH; ++g) {
          if (g < gh) {
            const float4 qa = *reinterpret_cast<const float4*>(q_w + g * HEAD_DIM + c * 8);
            const float4 qb = *reinterpret_cast<const float4*>(q_w + g * HEAD_DIM + c * 8 + 4);
            s[g] += qa.x * kf[0] + qa.y * kf[1] + qa.z * kf[2] + qa.w * kf[3] + qb.x * kf[4] + qb.y * kf[5] +
                    qb.z * kf[6] + qb.w * kf[7];
          }
        }
      }
      // ---- online softmax update (per head: running max m, running sum l, rescale of o)
      float corr[GH];
#pragma unroll
      for (int g = 0; g < GH; ++g) {
        const float sv = (lane < valid && g < gh) ? s[g] * scale : -INFINITY;
        const float m_new = fmaxf(m_w[g], warp_max(sv));
        corr[g] = (m_w[g] == -INFINITY) ? 0.f : __expf(m_w[g] - m_new);
        const float p = (lane < valid && g < gh) ? __expf(sv - m_new) : 0.f;
        l_w[g] = l_w[g] * corr[g] + warp_add(p);
        m_w[g] = m_new;
        p_w[g * 32 + lane] = p;
#pragma unroll
        for (int e = 0; e < 4; ++e) o_w[g][e] *= corr[g];
      }
      __syncwarp();
      // ---- PV: lane = dims [lane*4, lane*4+4); all 32 slots unconditionally (p == 0 beyond `valid`, cache slots
      // always hold finite values) -> constant trip count, loads pipeline
#pragma unroll 2
      for (int tk0 = 0; tk0 < KV_BLOCK; tk0 += 4) {
        uint2 vv[4];
#pragma unroll
        for (int jj = 0; jj < 4; ++jj)
          vv[jj] = *reinterpret_cast<const uint2*>(v_t + (tk0 + jj) * (HEAD_DIM * 2) +
                                                   ((lchunk ^ ((tk0 + jj) & 7)) * 16) + lhalf * 8);
        float4 pp[GH];
#pragma unroll
        for (int g = 0; g < GH; ++g) pp[g] = *reinterpret_cast<const float4*>(p_w + g * 32 + tk0);
#pragma unroll
        for (int jj = 0; jj < 4; ++jj) {
          const float v0 = __uint_as_float(vv[jj].x << 16), v1 = __uint_as_float(vv[jj].x & 0xffff0000u);
          const float v2 = __uint_as_float(vv[jj].y << 16), v3 = __uint_as_float(vv[jj].y & 0xffff0000u);
#pragma unroll
          for (int g = 0; g < GH; ++g) {
            const float p = jj == 0 ? pp[g].x : jj == 1 ? pp[g].y : jj == 2 ? pp[g].z : pp[g].w;
            o_w[g][0] += p * v0; o_w[g][1] += p * v1; o_w[g][2] += p * v2; o_w[g][3] += p * v3;
          }
        }
      }
      __syncwarp();  // p_w is rewritten by the next block of this warp
    }
    if (lane == 0) mbar_arrive(&empty_bar[st]);
  }
  // ---- combine the two block lanes.  All blocks are consumed (no TMA write can still be in flight), so the ring
  // memory is reused as staging: ow_s [2 lanes][G][128] floats.
  asm volatile("bar.sync 1, 128;\n" ::: "memory");
  float* ow_s = reinterpret_cast<float*>(kv_s);
#pragma unroll
  for (int g = 0; g < GH; ++g) {
    if (g < gh) {
      *reinterpret_cast<float4*>(ow_s + (blane * G + g0 + g) * HEAD_DIM + lane * 4) =
          make_float4(o_w[g][0], o_w[g][1], o_w[g][2], o_w[g][3]);
      if (lane == 0) {
        ml_s[(blane * G + g0 + g) * 2 + 0] = m_w[g];
        ml_s[(blane * G + g0 + g) * 2 + 1] = l_w[g];
      }
    }
  }
  asm volatile("bar.sync 1, 128;\n" ::: "memory");
  const int d = threadIdx.x;  // 128 compute threads = 128 dims
  float o_c[G], m_c[G], l_c[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    const float m0 = ml_s[(0 * G + g) * 2], m1 = ml_s[(1 * G + g) * 2];
    const float m = fmaxf(m0, m1);
    const float f0 = (m0 == -INFINITY) ? 0.f : __expf(m0 - m), f1 = (m1 == -INFINITY) ? 0.f : __expf(m1 - m);
    l_c[g] = ml_s[(0 * G + g) * 2 + 1] * f0 + ml_s[(1 * G + g) * 2 + 1] * f1;
    o_c[g] = ow_s[(0 * G + g) * HEAD_DIM + d] * f0 + ow_s[(1 * G + g) * HEAD_DIM + d] * f1;
    m_c[g] = m;
  }
  __nv_bfloat16* o_dst = out + (size_t)sq.q_start * out_ld + (size_t)kvh * G * HEAD_DIM;
  if (n_splits == 1) {
#pragma unroll
    for (int g = 0; g < G; ++g) o_dst[g * HEAD_DIM + d] = __float2bfloat16_rn(o_c[g] / l_c[g]);
    return;
  }
  // ---- multi-split: publish partial, last arriver merges in split order
  const size_t pbase = ((size_t)(sidx * n_kv + kvh) * max_splits);
#pragma unroll
  for (int g = 0; g < G; ++g) {
    part_o[((pbase + split) * G + g) * HEAD_DIM + d] = o_c[g];
    if (d == 0) {
      part_ml[((pbase + split) * G + g) * 2 + 0] = m_c[g];
      part_ml[((pbase + split) * G + g) * 2 + 1] = l_c[g];
    }
  }
  // publish: barrier + one acq_rel atomic (cumulative over the barrier) instead of membar.gl on every thread
  asm volatile("bar.sync 1, 128;\n" ::: "memory");
  if (threadIdx.x == 0) {
    int old;
    asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;\n"
                 : "=r"(old) : "l"(counters + sidx * n_kv + kvh) : "memory");
    *flag_s = (old == n_splits - 1);
  }
  asm volatile("bar.sync 1, 128;\n" ::: "memory");
  if (!*flag_s) return;
  // ---- last arriver: (m, l) of every split staged in smem by all threads in parallel (ring memory, past ow_s), then the
  // o rows fetched SPB splits at a time so that SPB*G L2 loads are in flight per thread
  float* ml_all = reinterpret_cast<float*>(kv_s + 2 * TILE_BYTES);  // [n_splits][G][2]
  for (int i = threadIdx.x; i < n_splits * G * 2; i += 128) ml_all[i] = __ldcg(&part_ml[pbase * G * 2 + i]);
  asm volatile("bar.sync 1, 128;\n" ::: "memory");
  float m_f[G], l_f[G], o_f[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    float m = -INFINITY;
    for (int sp = 0; sp < n_splits; ++sp) m = fmaxf(m, ml_all[(sp * G + g) * 2]);
    m_f[g] = m;
    l_f[g] = 0.f;
    o_f[g] = 0.f;
  }
  constexpr int SPB = 4;
  for (int sp0 = 0; sp0 < n_splits; sp0 += SPB) {
    float ov[SPB][G];
#pragma unroll
    for (int j = 0; j < SPB; ++j)
#pragma unroll
      for (int g = 0; g < G; ++g)
        ov[j][g] = (sp0 + j < n_splits) ? __ldcg(&part_o[((pbase + sp0 + j) * G + g) * HEAD_DIM + d]) : 0.f;
#pragma unroll
    for (int j = 0; j < SPB; ++j) {
      if (sp0 + j < n_splits) {
#pragma unroll
        for (int g = 0; g < G; ++g) {
          const float f = __expf(ml_all[((sp0 + j) * G + g) * 2] - m_f[g]);
          l_f[g] += ml_all[((sp0 + j) * G + g) * 2 + 1] * f;
          o_f[g] += ov[j][g] * f;
        }
      }
    }
  }
#pragma unroll
  for (int g = 0; g < G; ++g) o_dst[g * HEAD_DIM + d] = __float2bfloat16_rn(o_f[g] / l_f[g]);
  if (threadIdx.x == 0) counters[sidx * n_kv + kvh] = 0;
}

template <int G>
static cudaError_t decode_launch_g(const __nv_bfloat16* qkv, int qkv_ld, const __nv_bfloat16* k_cache,
                                   const __nv_bfloat16* v_cache, const AttnSeq* seqs, const int32_t* seq_ids,
                                   int n_seqs, const int32_t* block_table, int bt_stride, int max_splits,
                                   float* part_o, float* part_ml, int* counters, __nv_bfloat16* out, int out_ld,
                                   int n_kv, float scale, cudaStream_t stream) {
  constexpr int GH = (G + 1) / 2;
  const int smem = DEC_STAGES * 2 * TILE_BYTES + (G * HEAD_DIM + 4 * GH * 32 + 2 * G * 2) * 4 + 2 * DEC_STAGES * 8 + 16;
  static bool attr = false;
  if (!attr) {
    cudaError_t e =
        cudaFuncSetAttribute(attn_decode_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  dim3 grid(n_seqs * max_splits, n_kv);
  return launch_k(attn_decode_kernel<G>, grid, dim3(DEC_THREADS), smem, stream, qkv, qkv_ld, k_cache, v_cache, seqs,
                  seq_ids, block_table, bt_stride, max_splits, part_o, part_ml, counters, out, out_ld, n_kv, scale);
}

cudaError_t attn_decode_launch(const __nv_bfloat16* qkv, int qkv_ld, const __nv_bfloat16* k_cache,
                               const __nv_bfloat16* v_cache, const AttnSeq* seqs, const int32_t* seq_ids, int n_seqs,
                               const int32_t* block_table, int bt_stride, int max_splits, float* part_o,
                               float* part_ml, int* counters, __nv_bfloat16* out, int out_ld, int n_q, int n_kv,
                               float scale, cudaStream_t stream) {
  if (n_seqs <= 0) return cudaSuccess;
  if (n_q % n_kv != 0) return cudaErrorInvalidValue;
  const int G = n_q / n_kv;
#define TGIS_DEC(GG)                                                                                            \
  case GG:                                                                                                      \
    return decode_launch_g<GG>(qkv, qkv_ld, k_cache, v_cache, seqs, seq_ids, n_seqs, block_table, bt_stride,     \
                               max_splits, part_o, part_ml, counters, out, out_ld, n_kv, scale, stream)
  switch (G) {
    TGIS_DEC(1);
    TGIS_DEC(2);
    TGIS_DEC(3);
    TGIS_DEC(4);
    TGIS_DEC(8);
    default: return cudaErrorInvalidValue;
  }
#undef TGIS_DEC
}

// ================================================================================================ prefill
__device__ __forceinline__ void mma_bf16_16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

constexpr int PF_QROWS = 16;  // query tokens per tile

// grid (n_tiles, n_kv); blockDim = 32 * G (warp = query head inside the GQA group)
__global__ void __launch_bounds__(256)
attn_prefill_kernel(const __nv_bfloat16* __restrict__ qkv, int qkv_ld, const __nv_bfloat16* __restrict__ k_cache,
                    const __nv_bfloat16* __restrict__ v_cache, const AttnSeq* __restrict__ seqs,
                    const int32_t* __restrict__ tile_seq, const int32_t* __restrict__ tile_q0,
                    const int32_t* __restrict__ block_table, int bt_stride, __nv_bfloat16* __restrict__ out,
                    int out_ld, int n_kv, int G, float scale) {
  __shared__ __align__(128) uint8_t kv_s[2 * 2 * TILE_BYTES];  // 2 stages x (K | V)
  __shared__ uint64_t bars[2];
  griddep_launch();
  griddep_wait();
  const int tile = blockIdx.x, kvh = blockIdx.y;
  const AttnSeq sq = seqs[tile_seq[tile]];
  const int q0 = tile_q0[tile];                      // first query (index within this step's q_len)
  const int n_rows = min(PF_QROWS, sq.q_len - q0);   // valid query rows in the tile
  const int pos0 = sq.kv_len - sq.q_len + q0;        // absolute position of row 0
  const int last_pos = pos0 + n_rows - 1;
  const int n_blk = last_pos / KV_BLOCK + 1;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int g = lane >> 2, t = lane & 3;
  const int head = kvh * G + warp;

  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_barrier_init();
  }
  __syncthreads();
  const int32_t* bt = block_table + (size_t)sq.block_row * bt_stride;
  auto issue = [&](int j) {
    const int st = j & 1;
    const size_t tl = ((size_t)bt[j] * n_kv + kvh) * (KV_BLOCK * HEAD_DIM);
    mbar_arrive_expect_tx(&bars[st], 2 * TILE_BYTES);
    bulk_load_1d(kv_s + st * 2 * TILE_BYTES, k_cache + tl, TILE_BYTES, &bars[st]);
    bulk_load_1d(kv_s + st * 2 * TILE_BYTES + TILE_BYTES, v_cache + tl, TILE_BYTES, &bars[st]);
  };
  if (threadIdx.x == 0) {
    issue(0);
    if (n_blk > 1) issue(1);
  }

  // Q fragments (A operand), rows g and g+8 of the tile; rows past n_rows are clamped (results discarded)
  uint32_t qa[8][4];
  {
    const int r0 = min(g, n_rows - 1), r1 = min(g + 8, n_rows - 1);
    const __nv_bfloat16* q_r0 = qkv + (size_t)(sq.q_start + q0 + r0) * qkv_ld + (size_t)head * HEAD_DIM;
    const __nv_bfloat16* q_r1 = qkv + (size_t)(sq.q_start + q0 + r1) * qkv_ld + (size_t)head * HEAD_DIM;
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      qa[ks][0] = *reinterpret_cast<const uint32_t*>(q_r0 + ks * 16 + t * 2);
      qa[ks][1] = *reinterpret_cast<const uint32_t*>(q_r1 + ks * 16 + t * 2);
      qa[ks][2] = *reinterpret_cast<const uint32_t*>(q_r0 + ks * 16 + 8 + t * 2);
      qa[ks][3] = *reinterpret_cast<const uint32_t*>(q_r1 + ks * 16 + 8 + t * 2);
    }
  }
  float o[16][4];
#pragma unroll
  for (int i = 0; i < 16; ++i)
#pragma unroll
    for (int e = 0; e < 4; ++e) o[i][e] = 0.f;
  float m_r[2] = {-INFINITY, -INFINITY}, l_r[2] = {0.f, 0.f};
  const float sl2 = scale * 1.4426950408889634f;  // exp2 domain
  const int row_pos[2] = {pos0 + g, pos0 + g + 8};

  for (int j = 0; j < n_blk; ++j) {
    const int st = j & 1;
    mbar_wait(&bars[st], (j >> 1) & 1);
    const uint8_t* k_t = kv_s + st * 2 * TILE_BYTES;
    const uint8_t* v_t = k_t + TILE_BYTES;
    // ---- S = Q K^T : 4 n-tiles (8 tokens each) x 8 k-steps
    float s[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      s[nt][0] = s[nt][1] = s[nt][2] = s[nt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const int tok = nt * 8 + g;
        const uint32_t b0 = *reinterpret_cast<const uint32_t*>(k_t + ((2 * ks) * KV_BLOCK + tok) * 16 + t * 4);
        const uint32_t b1 = *reinterpret_cast<const uint32_t*>(k_t + ((2 * ks + 1) * KV_BLOCK + tok) * 16 + t * 4);
        mma_bf16_16816(s[nt], qa[ks], b0, b1);
      }
    }
    // ---- causal mask + online softmax (rows g, g+8; cols nt*8 + t*2 + {0,1})
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < 4; ++nt)
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int kpos = j * KV_BLOCK + nt * 8 + t * 2 + (e & 1);
        const int r = e >> 1;
        s[nt][e] = (kpos <= row_pos[r]) ? s[nt][e] * sl2 : -INFINITY;
        mx[r] = fmaxf(mx[r], s[nt][e]);
      }
    float corr[2];
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float m_new = fmaxf(m_r[r], mx[r]);
      corr[r] = (m_r[r] == -INFINITY) ? 0.f : exp2f(m_r[r] - m_new);
      m_r[r] = m_new;
      l_r[r] *= corr[r];
    }
#pragma unroll
    for (int i = 0; i < 16; ++i) {
      o[i][0] *= corr[0]; o[i][1] *= corr[0]; o[i][2] *= corr[1]; o[i][3] *= corr[1];
    }
    // P is fed to the tensor cores as bf16 hi + bf16 lo (two MMAs): ~16 mantissa bits instead of 8, so the result
    // tracks the fp32-softmax oracle instead of the usual flash-attention bf16-P rounding.
    uint32_t pa[2][4], pl[2][4];
#pragma unroll
    for (int nt = 0; nt < 4; ++nt) {
      float p[4], q[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const int r = e >> 1;
        p[e] = (m_r[r] == -INFINITY) ? 0.f : exp2f(s[nt][e] - m_r[r]);
        l_r[r] += p[e];
        q[e] = p[e] - bf16_round(p[e]);
      }
      const int kk = nt >> 1, hi = nt & 1;
      pa[kk][hi * 2 + 0] = pack_bf16x2(p[0], p[1]);
      pa[kk][hi * 2 + 1] = pack_bf16x2(p[2], p[3]);
      pl[kk][hi * 2 + 0] = pack_bf16x2(q[0], q[1]);
      pl[kk][hi * 2 + 1] = pack_bf16x2(q[2], q[3]);
    }
    // ---- O += P V : 2 k-steps (16 tokens) x 16 n-tiles (8 dims), V B-fragments by ldmatrix.trans
    const uint32_t v_base = smem_u32(v_t);
#pragma unroll
    for (int kk = 0; kk < 2; ++kk) {
#pragma unroll
      for (int nd = 0; nd < 16; nd += 2) {
        // lane -> (matrix = lane/8, row = lane%8): matrices (tok lo, nd), (tok hi, nd), (tok lo, nd+1), (tok hi, nd+1)
        const int mi = lane >> 3, ri = lane & 7;
        const int tok = kk * 16 + (mi & 1) * 8 + ri;
        const int chunk = nd + (mi >> 1);
        uint32_t vb[4];
        ldmatrix_x4_trans(vb, v_base + tok * (HEAD_DIM * 2) + ((chunk ^ (tok & 7)) * 16));
        mma_bf16_16816(o[nd], pa[kk], vb[0], vb[1]);
        mma_bf16_16816(o[nd + 1], pa[kk], vb[2], vb[3]);
        mma_bf16_16816(o[nd], pl[kk], vb[0], vb[1]);
        mma_bf16_16816(o[nd + 1], pl[kk], vb[2], vb[3]);
      }
    }
    __syncthreads();  // every warp is done with stage st
    if (threadIdx.x == 0 && j + 2 < n_blk) issue(j + 2);
  }
  // ---- normalise and store
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    l_r[r] += __shfl_xor_sync(0xffffffffu, l_r[r], 1);
    l_r[r] += __shfl_xor_sync(0xffffffffu, l_r[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int row = g + r * 8;
    if (row < n_rows) {
      const float inv = 1.f / l_r[r];
      __nv_bfloat16* dst = out + (size_t)(sq.q_start + q0 + row) * out_ld + (size_t)head * HEAD_DIM;
#pragma unroll
      for (int nd = 0; nd < 16; ++nd)
        *reinterpret_cast<uint32_t*>(dst + nd * 8 + t * 2) = pack_bf16x2(o[nd][r * 2] * inv, o[nd][r * 2 + 1] * inv);
    }
  }
}

cudaError_t attn_prefill_launch(const __nv_bfloat16* qkv, int qkv_ld, const __nv_bfloat16* k_cache,
                                const __nv_bfloat16* v_cache, const AttnSeq* seqs, const int32_t* tile_seq,
                                const int32_t* tile_q0, int n_tiles, const int32_t* block_table, int bt_stride,
                                __nv_bfloat16* out, int out_ld, int n_q, int n_kv, float scale, cudaStream_t stream) {
  if (n_tiles <= 0) return cudaSuccess;
  if (n_q % n_kv != 0) return cudaErrorInvalidValue;
  const int G = n_q / n_kv;
  if (G > 8) return cudaErrorInvalidValue;
  dim3 grid(n_tiles, n_kv);
  return launch_k(attn_prefill_kernel, grid, dim3(32 * G), 0, stream, qkv, qkv_ld, k_cache, v_cache, seqs, tile_seq,
                  tile_q0, block_table, bt_stride, out, out_ld, n_kv, G, scale);
}

}  // namespace tgis
