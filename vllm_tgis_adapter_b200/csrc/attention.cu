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
#include <algorithm>

#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace tgis {

TGIS_STL_DEFINE(attention)

constexpr int DEC_TOK = DECODE_SPLIT;        // tokens per split (8 KV blocks streamed through a 4-stage ring)
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

__device__ __forceinline__ uint32_t movmatrix_trans(uint32_t a) {  // 8x8 b16 transpose across the warp
  uint32_t d;
  asm volatile("movmatrix.sync.aligned.m8n8.trans.b16 %0, %1;\n" : "=r"(d) : "r"(a));
  return d;
}
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

// ================================================================================================ decode
// Persistent, warp-granular flash-decoding on the tensor cores.
//
// Work item = (sequence, 128-token split, kv head); the host lists the (sequence, split) entries of the step
// (DecItem), the kv head is the fastest index.  The grid is 2 CTAs per SM x 4 warps, and every WARP walks its own
// strided slice of the item list, alone: no CTA-wide synchronisation exists after the barrier initialisation.
// A warp owns a ring of three 8 KiB tile buffers fed by 1-D TMA bulk copies (K0 V0 K1 V1 ... of its items, the
// stream does not stop at item boundaries) and a Q buffer; the copy of tile t+3 is requested the moment tile t has
// been consumed, and the next item's metadata / block ids are fetched one item / one block ahead, so a warp keeps
// 16-24 KiB in flight while it computes and never waits on a dependent global load in steady state (8 warps per SM).
//
// Per 32-token block the tokens / dims ride the M dimension of m16n8k16 and the (up to 8) query heads of the group the
// N dimension, so nothing is padded: S^T = K Q^T is 16 MMAs with K A-fragments straight from the chunk-major tile by
// ldmatrix; P^T (bf16 hi + lo, moved into B-fragment layout by movmatrix) feeds O^T = V^T P^T, 32 MMAs with
// ldmatrix.trans on the swizzled V tile; the online-softmax state (m, l, o) stays in registers for the whole item.
// kv_len > 128: the warp stores its (o, m, l) partial and attn_merge_kernel (next launch, PDL-chained) merges the
// splits in split order: no atomics or fences in the streaming kernel (a per-item release fence cost 20 % of its
// time).  The split size is a constant, so the arithmetic of a sequence depends only on its own length, never on what
// else is in the batch or on which warp ran it.
constexpr int DEC_WARPS = 4;
constexpr int DEC_THREADS = DEC_WARPS * 32;
#ifndef TGIS_DEC_RING
#define TGIS_DEC_RING 3  // tile buffers per warp
#endif
#ifndef TGIS_DEC_MINB
#define TGIS_DEC_MINB 2  // CTAs per SM (register cap); shared memory must agree: MINB * 4 * (RING * 8 + 2) KiB <= 227
#endif
constexpr int DEC_RING = TGIS_DEC_RING;
constexpr int DEC_Q_BYTES = 8 * HEAD_DIM * 2;                           // up to 8 heads
constexpr int DEC_WARP_SMEM = DEC_RING * TILE_BYTES + DEC_Q_BYTES;      // 26 KiB
constexpr int DEC_SMEM = DEC_WARPS * DEC_WARP_SMEM + DEC_WARPS * 4 * 8; // + 4 mbarriers per warp

template <int G>
__global__ void __launch_bounds__(DEC_THREADS, TGIS_DEC_MINB)
attn_decode_kernel(const __nv_bfloat16* __restrict__ qkv, int qkv_ld, const __nv_bfloat16* __restrict__ k_cache,
                   const __nv_bfloat16* __restrict__ v_cache, const DecItem* __restrict__ items,
                   int max_splits, float* __restrict__ part_o, float* __restrict__ part_ml,
                   __nv_bfloat16* __restrict__ out, int out_ld, int n_kv, float scale, int* __restrict__ arrive,
                   int prefetch_old) {
  static_assert(G <= 8, "query heads of a group are rows 0..7 of the MMA tile");
  extern __shared__ __align__(128) uint8_t smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint8_t* ring = smem + warp * DEC_WARP_SMEM;
  uint8_t* q_s = ring + DEC_RING * TILE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + DEC_WARPS * DEC_WARP_SMEM) + warp * 4;  // [3] tiles, [3] = Q
  uint64_t* q_bar = full + 3;
  if (lane == 0) {
    for (int i = 0; i < 4; ++i) mbar_init(&full[i], 1);
    fence_barrier_init();
  }
  __syncwarp();
  STL_ENTER(3);
  griddep_launch();
  // ---- item list: host-written for this step, not produced by a preceding kernel -> readable before griddep_wait
  const int4* it4 = reinterpret_cast<const int4*>(items);  // record e = it4[2e] (q_row, kv_len, seq_split, -), it4[2e+1] (blocks)
  const int n_flat = it4[0].x * n_kv;
  const int n_slots = gridDim.x * DEC_WARPS;
  const int gw = warp * gridDim.x + blockIdx.x;  // consecutive items land on different SMs
  // The host lists the items longest first; round k hands item k*n_slots + (k even ? gw : n_slots-1-gw) to warp gw
  // (boustrophedon), so the warps that got the longest items of one round get the shortest of the next.
  auto item_of_round = [&](int k) { return k * n_slots + ((k & 1) ? n_slots - 1 - gw : gw); };
  int f = gw;
  if (f >= n_flat) return;
  auto item_at = [&](int ff) { return __ldg(&it4[2 * (1 + ff / n_kv)]); };
  auto blocks_at = [&](int ff) { return __ldg(&it4[2 * (1 + ff / n_kv) + 1]); };
  auto n_blk_of = [&](const int4& it) {
    const int n_tok = min(DEC_TOK, it.y - (it.z >> 16) * DEC_TOK);
    return (n_tok + KV_BLOCK - 1) / KV_BLOCK;
  };
  auto pick = [](const int4& b, int j) { return j == 0 ? b.x : j == 1 ? b.y : j == 2 ? b.z : b.w; };
  const uint64_t kv_policy = policy_evict_first();
  int4 c_it = item_at(f);
  int4 c_nxt = c_it;
  // ---- issue cursor: runs DEC_RING tiles ahead of consumption, across item boundaries
  int i_f = f, i_kvh = f % n_kv, i_jb = 0, i_nblk = n_blk_of(c_it), i_kv = 0, i_buf = 0;
  int4 i_nxt = c_it, i_blks = blocks_at(f), i_blks_nxt = i_blks;  // block ids of the item being issued / the next one
  int4 i_cur = c_it;                                                // record of the item being issued
  int i_k = 0, c_k = 0;  // round of the issue cursor / of the consumer
  if (item_of_round(1) < n_flat) {
    i_nxt = item_at(item_of_round(1));
    i_blks_nxt = blocks_at(item_of_round(1));
  }
  auto issue_tile = [&]() {
    if (i_f >= n_flat) return;
    if (lane == 0) {
      const __nv_bfloat16* src =
          (i_kv ? v_cache : k_cache) + ((size_t)pick(i_blks, i_jb) * n_kv + i_kvh) * (KV_BLOCK * HEAD_DIM);
      mbar_arrive_expect_tx(&full[i_buf], TILE_BYTES);
      bulk_load_1d_hint(ring + i_buf * TILE_BYTES, src, TILE_BYTES, &full[i_buf], kv_policy);  // read once per step
    }
    i_buf = i_buf == DEC_RING - 1 ? 0 : i_buf + 1;
    if (i_kv == 0) {
      i_kv = 1;
      return;
    }
    i_kv = 0;
    if (++i_jb == i_nblk) {  // next item of this warp; its record was requested one item ago
      i_f = item_of_round(++i_k);
      if (i_f >= n_flat) return;
      i_kvh = i_f % n_kv;
      i_jb = 0;
      i_nblk = n_blk_of(i_nxt);
      i_cur = i_nxt;
      i_blks = i_blks_nxt;
      if (item_of_round(i_k + 1) < n_flat) {
        i_nxt = item_at(item_of_round(i_k + 1));
        i_blks_nxt = blocks_at(item_of_round(i_k + 1));
      }
    }
  };
  auto issue_q = [&](const int4& it, int kvh) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_bar, G * HEAD_DIM * 2);
      bulk_load_1d(q_s, qkv + (size_t)it.x * qkv_ld + (size_t)kvh * G * HEAD_DIM, G * HEAD_DIM * 2, q_bar);
    }
  };

  // Cache blocks that hold only tokens of EARLIER steps do not depend on the preceding kernels: stream them while the
  // qkv projection is still draining.  The block holding position kv_len-1 (last block of the last split) was written
  // this step; the in-order issue cursor stops in front of it until the dependency resolves.
  int n_pre = 0;
  if (prefetch_old) {
    while (n_pre < DEC_RING && i_f < n_flat) {
      const bool tail_block = (i_cur.z >> 16) == (i_cur.y + DEC_TOK - 1) / DEC_TOK - 1 && i_jb == i_nblk - 1;
      if (tail_block) break;
      issue_tile();
      ++n_pre;
    }
  }
  griddep_wait();  // q and the newest cache slot come from the preceding kernels
  STL_WAITED();
  issue_q(c_it, f % n_kv);
  for (int i = n_pre; i < DEC_RING; ++i) issue_tile();

  const int hr = lane >> 2, t4 = lane & 3;
  const int mi = lane >> 3, ri = lane & 7;          // ldmatrix: lane -> (matrix, row)
  const float sl2 = scale * 1.4426950408889634f;    // exp2 domain
  const uint32_t ring_base = smem_u32(ring);
  int c_buf = 0, c_ph = 0, q_ph = 0;

  for (; f < n_flat; f = item_of_round(++c_k)) {
    const int kvh = f % n_kv;
    const int seq = c_it.z & 0xffff, split = c_it.z >> 16;
    const int kv_len = c_it.y;
    const int n_splits = (kv_len + DEC_TOK - 1) / DEC_TOK;
    const int n_tok = min(DEC_TOK, kv_len - split * DEC_TOK);
    const int n_blk = (n_tok + KV_BLOCK - 1) / KV_BLOCK;
    const int f_next = item_of_round(c_k + 1);
    const bool has_next = f_next < n_flat;
    if (has_next) c_nxt = item_at(f_next);
    // ---- Q^T as B fragments: n = head (lane/4), 8 k-steps of 16 dims; heads >= G are zero columns
    mbar_wait(q_bar, q_ph);
    q_ph ^= 1;
    uint32_t qb[8][2];
#pragma unroll
    for (int ks = 0; ks < 8; ++ks) {
      const uint8_t* qp = q_s + (hr * HEAD_DIM + ks * 16 + t4 * 2) * 2;
      qb[ks][0] = hr < G ? *reinterpret_cast<const uint32_t*>(qp) : 0u;
      qb[ks][1] = hr < G ? *reinterpret_cast<const uint32_t*>(qp + 16) : 0u;
    }
    // this thread's accumulators: heads 2*t4 + {0,1} (e & 1), dims md*16 + hr + {0,8} (e >> 1)
    float ot[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i) ot[i][0] = ot[i][1] = ot[i][2] = ot[i][3] = 0.f;
    float m_r[2] = {-INFINITY, -INFINITY}, l_r[2] = {0.f, 0.f};  // l_r: this thread's share (reduced at the end)

    for (int jb = 0; jb < n_blk; ++jb) {
      // ---- S^T = K Q^T : 2 m-tiles (16 tokens) x 8 k-steps (16 dims = chunks 2ks, 2ks+1); ldmatrix.x4 = one A
      // fragment: matrices (tok lo, chunk 2ks), (tok hi, chunk 2ks), (tok lo, chunk 2ks+1), (tok hi, chunk 2ks+1)
      mbar_wait(&full[c_buf], c_ph);
      const uint32_t k_base = ring_base + c_buf * TILE_BYTES;
      float st[2][4];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) st[mt][0] = st[mt][1] = st[mt][2] = st[mt][3] = 0.f;
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
#pragma unroll
        for (int mt = 0; mt < 2; ++mt) {
          uint32_t ka[4];
          ldmatrix_x4(ka, k_base + ((2 * ks + (mi >> 1)) * KV_BLOCK + mt * 16 + (mi & 1) * 8 + ri) * 16);
          mma_bf16_16816(st[mt], ka, qb[ks][0], qb[ks][1]);
        }
      }
      __syncwarp();
      issue_tile();  // refills the K buffer just consumed (tile t + DEC_RING of this warp's stream)
      if (++c_buf == DEC_RING) { c_buf = 0; c_ph ^= 1; }
      if (jb == 0) {  // q_s was copied to registers above; by now the next item's record has arrived
        __syncwarp();
        if (has_next) issue_q(c_nxt, f_next % n_kv);
      }
      // ---- online softmax; st[mt][e]: token mt*16 + hr + (e>>1)*8, head 2*t4 + (e&1)
      const int valid = min(KV_BLOCK, n_tok - jb * KV_BLOCK);
      float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
      for (int mt = 0; mt < 2; ++mt)
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          st[mt][e] = (mt * 16 + hr + (e >> 1) * 8 < valid) ? st[mt][e] * sl2 : -INFINITY;
          mx[e & 1] = fmaxf(mx[e & 1], st[mt][e]);
        }
      float corr[2];
#pragma unroll
      for (int hh = 0; hh < 2; ++hh) {
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 4));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 8));
        mx[hh] = fmaxf(mx[hh], __shfl_xor_sync(0xffffffffu, mx[hh], 16));
        const float m_new = fmaxf(m_r[hh], mx[hh]);  // finite: token 0 of every block of the split is valid
        corr[hh] = (m_r[hh] == -INFINITY) ? 0.f : exp2f(m_r[hh] - m_new);
        m_r[hh] = m_new;
        l_r[hh] *= corr[hh];
      }
      // P^T as B fragments, bf16 hi + bf16 lo (two MMAs): ~16 mantissa bits, tracks the fp32 oracle.  The score layout
      // (row = token, column pair = heads) is the transpose of the B layout (k = token pair, n = head): movmatrix.
      uint32_t bh[2][2], bl[2][2];
#pragma unroll
      for (int mt = 0; mt < 2; ++mt) {
        float p[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          p[e] = exp2f(st[mt][e] - m_r[e & 1]);
          l_r[e & 1] += p[e];
        }
        bh[mt][0] = movmatrix_trans(pack_bf16x2(p[0], p[1]));
        bh[mt][1] = movmatrix_trans(pack_bf16x2(p[2], p[3]));
        bl[mt][0] = movmatrix_trans(pack_bf16x2(p[0] - bf16_round(p[0]), p[1] - bf16_round(p[1])));
        bl[mt][1] = movmatrix_trans(pack_bf16x2(p[2] - bf16_round(p[2]), p[3] - bf16_round(p[3])));
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        ot[i][0] *= corr[0];
        ot[i][1] *= corr[1];
        ot[i][2] *= corr[0];
        ot[i][3] *= corr[1];
      }
      // ---- O^T += V^T P^T : 8 m-tiles (16 dims = chunks 2md, 2md+1) x 2 k-steps (16 tokens); V^T A-fragments by
      // ldmatrix.trans: matrices (tok lo, chunk 2md), (tok lo, chunk 2md+1), (tok hi, chunk 2md), (tok hi, chunk 2md+1).
      // Slots >= valid carry p = 0 (cache slots hold finite values).
      mbar_wait(&full[c_buf], c_ph);
      const uint32_t v_base = ring_base + c_buf * TILE_BYTES;
#pragma unroll
      for (int kk = 0; kk < 2; ++kk) {
        const int tok = kk * 16 + (mi >> 1) * 8 + ri;
#pragma unroll
        for (int md = 0; md < 8; md += 2) {
          uint32_t va[2][4];
#pragma unroll
          for (int u = 0; u < 2; ++u)
            ldmatrix_x4_trans(va[u], v_base + tok * (HEAD_DIM * 2) + (((2 * (md + u) + (mi & 1)) ^ (tok & 7)) * 16));
          mma_bf16_16816(ot[md], va[0], bh[kk][0], bh[kk][1]);
          mma_bf16_16816(ot[md + 1], va[1], bh[kk][0], bh[kk][1]);
          mma_bf16_16816(ot[md], va[0], bl[kk][0], bl[kk][1]);
          mma_bf16_16816(ot[md + 1], va[1], bl[kk][0], bl[kk][1]);
        }
      }
      __syncwarp();
      issue_tile();
      if (++c_buf == DEC_RING) { c_buf = 0; c_ph ^= 1; }
    }
#pragma unroll
    for (int hh = 0; hh < 2; ++hh) {
      l_r[hh] += __shfl_xor_sync(0xffffffffu, l_r[hh], 4);
      l_r[hh] += __shfl_xor_sync(0xffffffffu, l_r[hh], 8);
      l_r[hh] += __shfl_xor_sync(0xffffffffu, l_r[hh], 16);
    }

    // ---- item epilogue
    if (n_splits == 1) {
      __nv_bfloat16* o_dst = out + (size_t)c_it.x * out_ld + (size_t)(kvh * G + 2 * t4) * HEAD_DIM + hr;
      const float inv[2] = {1.f / l_r[0], 1.f / l_r[1]};
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (2 * t4 + (e & 1) < G) {
#pragma unroll
          for (int md = 0; md < 8; ++md)
            o_dst[(e & 1) * HEAD_DIM + md * 16 + (e >> 1) * 8] = __float2bfloat16_rn(ot[md][e] * inv[e & 1]);
        }
      }
    } else {
      // store the partial (m in the exp2 domain); attn_merge_kernel combines the splits in split order
      const size_t pbase = (size_t)(seq * n_kv + kvh) * max_splits;
      float* po = part_o + ((pbase + split) * G + 2 * t4) * HEAD_DIM + hr;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (2 * t4 + (e & 1) < G) {
#pragma unroll
          for (int md = 0; md < 8; ++md) po[(e & 1) * HEAD_DIM + md * 16 + (e >> 1) * 8] = ot[md][e];
        }
      }
      if (hr == 0) {
#pragma unroll
        for (int hh = 0; hh < 2; ++hh)
          if (2 * t4 + hh < G)
            *reinterpret_cast<float2*>(part_ml + ((pbase + split) * G + 2 * t4 + hh) * 2) = make_float2(m_r[hh], l_r[hh]);
      }
      if (arrive != nullptr) {
        // In-kernel split merge: the warp that delivers the LAST partial of a (sequence, kv head) merges all of them,
        // in split order (the arithmetic of attn_merge_kernel, so the result does not depend on who merges).  Publish:
        // warp barrier, then ONE acq_rel atomic by lane 0 (cumulative over the barrier) -- no membar.
        __syncwarp();
        int old = 0;
        if (lane == 0)
          asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;\n"
                       : "=r"(old)
                       : "l"(arrive + seq * n_kv + kvh)
                       : "memory");
        old = __shfl_sync(0xffffffffu, old, 0);
        if (old == n_splits - 1) {
          // lane -> dims 4*lane .. 4*lane+3 of every head; all loads of a chunk of splits are issued before the math
          float m_f[G], l_f[G];
          float4 o_f[G];
#pragma unroll
          for (int g = 0; g < G; ++g) {
            m_f[g] = -INFINITY;
            l_f[g] = 0.f;
            o_f[g] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
          for (int sp = 0; sp < n_splits; ++sp)
#pragma unroll
            for (int g = 0; g < G; ++g) m_f[g] = fmaxf(m_f[g], __ldcg(part_ml + ((pbase + sp) * G + g) * 2));
          constexpr int SPB = 4;
          for (int sp0 = 0; sp0 < n_splits; sp0 += SPB) {
            float4 ov[SPB][G];
            float2 ml[SPB][G];
#pragma unroll
            for (int j = 0; j < SPB; ++j)
#pragma unroll
              for (int g = 0; g < G; ++g) {
                const bool v = sp0 + j < n_splits;
                ov[j][g] = v ? __ldcg(reinterpret_cast<const float4*>(part_o + ((pbase + sp0 + j) * G + g) * HEAD_DIM) + lane)
                             : make_float4(0.f, 0.f, 0.f, 0.f);
                ml[j][g] = v ? __ldcg(reinterpret_cast<const float2*>(part_ml + ((pbase + sp0 + j) * G + g) * 2))
                             : make_float2(-INFINITY, 0.f);
              }
#pragma unroll
            for (int j = 0; j < SPB; ++j) {
              if (sp0 + j < n_splits) {
#pragma unroll
                for (int g = 0; g < G; ++g) {
                  const float fs = exp2f(ml[j][g].x - m_f[g]);
                  l_f[g] += ml[j][g].y * fs;
                  o_f[g].x += ov[j][g].x * fs;
                  o_f[g].y += ov[j][g].y * fs;
                  o_f[g].z += ov[j][g].z * fs;
                  o_f[g].w += ov[j][g].w * fs;
                }
              }
            }
          }
          __nv_bfloat16* o_row = out + (size_t)c_it.x * out_ld + (size_t)kvh * G * HEAD_DIM + 4 * lane;
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const __nv_bfloat162 lo2 = __floats2bfloat162_rn(o_f[g].x / l_f[g], o_f[g].y / l_f[g]);
            const __nv_bfloat162 hi2 = __floats2bfloat162_rn(o_f[g].z / l_f[g], o_f[g].w / l_f[g]);
            uint2 pk;
            pk.x = *reinterpret_cast<const uint32_t*>(&lo2);
            pk.y = *reinterpret_cast<const uint32_t*>(&hi2);
            *reinterpret_cast<uint2*>(o_row + g * HEAD_DIM) = pk;
          }
          if (lane == 0) arrive[seq * n_kv + kvh] = 0;  // re-armed for the next launch (stream order)
        }
      }
    }
    c_it = c_nxt;
  }
  STL_EXIT();
}

// Split merge: one CTA per (decode sequence, kv head) with more than one split, thread = dim.  Reads the partials in
// split order (deterministic), writes the normalised bf16 rows.
template <int G>
__global__ void __launch_bounds__(HEAD_DIM)
attn_merge_kernel(const AttnSeq* __restrict__ seqs, const int32_t* __restrict__ seq_ids, int max_splits,
                  const float* __restrict__ part_o, const float* __restrict__ part_ml,
                  __nv_bfloat16* __restrict__ out, int out_ld, int n_kv) {
  __shared__ float ml_s[64 * G * 2];
  STL_ENTER(4);
  griddep_launch();
  const int sidx = blockIdx.x, kvh = blockIdx.y, d = threadIdx.x;
  const AttnSeq sq = seqs[seq_ids ? seq_ids[sidx] : sidx];  // host-written
  const int n_splits = (sq.kv_len + DEC_TOK - 1) / DEC_TOK;
  if (n_splits == 1) return;  // the streaming kernel wrote the row itself
  griddep_wait();
  STL_WAITED();
  const size_t pbase = (size_t)(sidx * n_kv + kvh) * max_splits;
  float m_f[G], l_f[G], o_f[G];
#pragma unroll
  for (int g = 0; g < G; ++g) m_f[g] = -INFINITY, l_f[g] = 0.f, o_f[g] = 0.f;
  // pass 1: row maxima (the (m, l) pairs go through shared memory in chunks of 64 splits)
  for (int c0 = 0; c0 < n_splits; c0 += 64) {
    const int nc = min(64, n_splits - c0);
    __syncthreads();
    for (int i = d; i < nc * G * 2; i += HEAD_DIM) ml_s[i] = __ldcg(&part_ml[(pbase + c0) * G * 2 + i]);
    __syncthreads();
    for (int sp = 0; sp < nc; ++sp)
#pragma unroll
      for (int g = 0; g < G; ++g) m_f[g] = fmaxf(m_f[g], ml_s[(sp * G + g) * 2]);
  }
  // pass 2: weighted sums in split order, SPB splits of loads in flight per thread
  constexpr int SPB = 4;
  for (int c0 = 0; c0 < n_splits; c0 += 64) {
    const int nc = min(64, n_splits - c0);
    if (n_splits > 64) {
      __syncthreads();
      for (int i = d; i < nc * G * 2; i += HEAD_DIM) ml_s[i] = __ldcg(&part_ml[(pbase + c0) * G * 2 + i]);
      __syncthreads();
    }
    for (int sp0 = 0; sp0 < nc; sp0 += SPB) {
      float ov[SPB][G];
#pragma unroll
      for (int j = 0; j < SPB; ++j)
#pragma unroll
        for (int g = 0; g < G; ++g)
          ov[j][g] = (sp0 + j < nc) ? __ldcg(&part_o[((pbase + c0 + sp0 + j) * G + g) * HEAD_DIM + d]) : 0.f;
#pragma unroll
      for (int j = 0; j < SPB; ++j) {
        if (sp0 + j < nc) {
#pragma unroll
          for (int g = 0; g < G; ++g) {
            const float fs = exp2f(ml_s[((sp0 + j) * G + g) * 2] - m_f[g]);
            l_f[g] += ml_s[((sp0 + j) * G + g) * 2 + 1] * fs;
            o_f[g] += ov[j][g] * fs;
          }
        }
      }
    }
  }
  __nv_bfloat16* o_dst = out + (size_t)sq.q_start * out_ld + (size_t)kvh * G * HEAD_DIM;
#pragma unroll
  for (int g = 0; g < G; ++g) o_dst[g * HEAD_DIM + d] = __float2bfloat16_rn(o_f[g] / l_f[g]);
  STL_EXIT();
}

template <int G>
static cudaError_t decode_launch_g(const __nv_bfloat16* qkv, int qkv_ld, const __nv_bfloat16* k_cache,
                                   const __nv_bfloat16* v_cache, const DecItem* items, int max_entries,
                                   const AttnSeq* seqs, const int32_t* seq_ids, int n_seqs, int max_splits,
                                   float* part_o, float* part_ml, __nv_bfloat16* out, int out_ld, int n_kv,
                                   float scale, int num_sms, cudaStream_t stream, int* arrive) {
  static bool attr = false;
  if (!attr) {
    cudaError_t e =
        cudaFuncSetAttribute(attn_decode_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, DEC_SMEM);
    if (e != cudaSuccess) return e;
    attr = true;
  }
  const long long max_flat = (long long)max_entries * n_kv;
  const int grid = (int)std::min<long long>((long long)TGIS_DEC_MINB * num_sms, (max_flat + DEC_WARPS - 1) / DEC_WARPS);
  // split merge: the PDL-chained attn_merge_kernel (default), or inside the streaming kernel by the last-arriving warp
  // (TGIS_ATTN_INKERNEL_MERGE=1, bit-identical).  Measured (profiles/r02_step_timeline_*): folding the merge saves its
  // launch (-100 us per step at batch 32) but the o-proj GEMM that follows then loses the window in which it prefetched
  // its weights next to the small merge kernel (attention CTAs fill the SM's shared memory): +140 us on the GEMM, net
  // -1 % at batch 64.  So the separate kernel stays the default.
  const char* env_merge = getenv("TGIS_ATTN_INKERNEL_MERGE");  // read per launch: the A/B test flips it in-process
  const int inkernel = (env_merge && env_merge[0] == '1') ? 1 : 0;
  const char* env_pre = getenv("TGIS_ATTN_PREFETCH");  // 0: no cache reads before the grid-dependency wait
  const int prefetch_old = (env_pre && env_pre[0] == '0') ? 0 : 1;
  int* arr = (inkernel && max_splits > 1) ? arrive : nullptr;
  cudaError_t e = launch_k(attn_decode_kernel<G>, dim3(grid), dim3(DEC_THREADS), DEC_SMEM, stream, qkv, qkv_ld,
                           k_cache, v_cache, items, max_splits, part_o, part_ml, out, out_ld, n_kv, scale, arr, prefetch_old);
  if (e != cudaSuccess || max_splits <= 1 || arr != nullptr) return e;
  return launch_k(attn_merge_kernel<G>, dim3(n_seqs, n_kv), dim3(HEAD_DIM), 0, stream, seqs, seq_ids, max_splits,
                  (const float*)part_o, (const float*)part_ml, out, out_ld, n_kv);
}

cudaError_t attn_decode_launch(const __nv_bfloat16* qkv, int qkv_ld, const __nv_bfloat16* k_cache,
                               const __nv_bfloat16* v_cache, const DecItem* items, int max_entries,
                               const AttnSeq* seqs, const int32_t* seq_ids, int n_seqs, int max_splits, float* part_o,
                               float* part_ml, __nv_bfloat16* out, int out_ld, int n_q, int n_kv, float scale,
                               int num_sms, cudaStream_t stream, int* arrive) {
  if (max_entries <= 0 || n_seqs <= 0) return cudaSuccess;
  if (n_q % n_kv != 0) return cudaErrorInvalidValue;
  const int G = n_q / n_kv;
#define TGIS_DEC(GG)                                                                                               \
  case GG:                                                                                                         \
    return decode_launch_g<GG>(qkv, qkv_ld, k_cache, v_cache, items, max_entries, seqs, seq_ids, n_seqs, max_splits, \
                               part_o, part_ml, out, out_ld, n_kv, scale, num_sms, stream, arrive)
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
