// ONE fused TGIS sampling kernel: everything between the lm_head logits and the (token, logprob, rank, top-n) record.
//
// It replaces ~30 small launches + a per-request Python loop in the reference stack:
//   reference-owned : tgis_utils/logits_processors.py:24-47  ExpDecayLengthPenaltyWarper      (R7)
//                     tgis_utils/logits_processors.py:7-21   TypicalLogitsWarperWrapper       (R8)
//                       -> hf generation/logits_process.py:836-856
//   vllm 0.22       : v1/sample/sampler.py:67-144 (raw log-softmax, min_tokens, penalties, greedy/temperature,
//                     top-k/top-p, exponential-race sampling, gather_logprobs + rank)         (S1-S8)
// Contract order per row: raw log-softmax -> typical-p -> ExpDecay(EOS) -> min_tokens(EOS)
// -> repetition penalty -> greedy argmax | temperature -> top-k -> top-p -> sample -> logprob / rank / top-n (raw).
//
// Shape of the problem: a [rows, V=128256] scan with integer-ish selection work; a sampling row needs ~10 passes of
// 50-100 instructions per element, which is ISSUE bound on one SM (1 ms per row).  So a row is owned by a thread-block
// CLUSTER of 8 CTAs x 1024 threads on 8 SMs: each CTA scans one eighth of the vocabulary, block results are exchanged
// through distributed shared memory and combined in rank order (deterministic).  The fp32 row (512 KiB) is read from
// HBM once and stays in the 126 MB L2 for the selection passes (radix-select thresholds instead of the reference's
// full sorts).  Loads are 16-byte vectorised and coalesced.
#include <cooperative_groups.h>

#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace cg = cooperative_groups;

namespace tgis {

TGIS_STL_DEFINE(sampler)

constexpr int SAMP_CL = 8;  // CTAs (SMs) per row
constexpr int SAMP_THREADS = 1024;
constexpr int SAMP_WARPS = SAMP_THREADS / 32;

// LT = logits element type: __nv_bfloat16 on the product path (vLLM's lm_head emits model-dtype logits and the sampler
// casts them to fp32: vllm v1/sample/sampler.py:91 -- every bf16 value is exactly representable), float for the
// kernel-level golden tests (fp32 fixtures generated from the reference's own code).
template <class LT>
struct RowCtx {
  const LT* x;             // raw logits
  int V;
  int lo, hi;              // this CTA's slice of the vocabulary (multiples of 8)
  const uint32_t* seen;    // bitmap of prompt U output tokens (may be null)
  SampleRow p;
  float lenfac_m1;         // (float)(decay^n - 1), 0 => inactive
  bool mask_eos;           // min_tokens not reached
  // typical-p state
  bool typical;
  float raw_max, raw_logz, ent, typ_thr;
};

__device__ __forceinline__ float lt2f(float v) { return v; }
__device__ __forceinline__ float lt2f(__nv_bfloat16 v) { return __bfloat162float(v); }
template <class LT>
__device__ __forceinline__ float load_x(const RowCtx<LT>& c, int i) { return lt2f(c.x[i]); }
// 8 consecutive logits (i0 % 8 == 0) as fp32
__device__ __forceinline__ void load_x8(const float* x, int i0, float (&o)[8]) {
  const float4 ra = *reinterpret_cast<const float4*>(x + i0);
  const float4 rb = *reinterpret_cast<const float4*>(x + i0 + 4);
  o[0] = ra.x; o[1] = ra.y; o[2] = ra.z; o[3] = ra.w; o[4] = rb.x; o[5] = rb.y; o[6] = rb.z; o[7] = rb.w;
}
__device__ __forceinline__ void load_x8(const __nv_bfloat16* x, int i0, float (&o)[8]) {
  const uint4 r = *reinterpret_cast<const uint4*>(x + i0);
  const uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    o[2 * e] = __uint_as_float(w[e] << 16);
    o[2 * e + 1] = __uint_as_float(w[e] & 0xffff0000u);
  }
}

template <class LT>
__device__ __forceinline__ bool is_seen(const RowCtx<LT>& c, int i) {
  return c.seen != nullptr && ((c.seen[i >> 5] >> (i & 31)) & 1u);
}

// processed logit (before temperature) of vocabulary entry i with raw value x
template <class LT>
__device__ __forceinline__ float process(const RowCtx<LT>& c, int i, float x) {
  float y = x;
  if (c.typical) {
    const float lp = (x - c.raw_max) - c.raw_logz;
    const float s = fabsf((-lp) - c.ent);
    if (s > c.typ_thr) y = -INFINITY;
  }
  if (i == c.p.eos_id) {
    if (c.lenfac_m1 != 0.f && isfinite(y)) y = __fadd_rn(y, __fmul_rn(fabsf(y), c.lenfac_m1));
    if (c.mask_eos) y = -INFINITY;
  }
  if (c.p.rep_penalty != 1.0f && is_seen(c, i)) y = (y > 0.f) ? __fdiv_rn(y, c.p.rep_penalty) : __fmul_rn(y, c.p.rep_penalty);
  return y;
}

// ---------------------------------------------------------------- block reductions (1024 threads)
struct MaxSum {
  float m, s;
};
__device__ __forceinline__ MaxSum ms_combine(MaxSum a, MaxSum b) {
  if (a.m == -INFINITY) return b;
  if (b.m == -INFINITY) return a;
  const float m = fmaxf(a.m, b.m);
  return {m, a.s * __expf(a.m - m) + b.s * __expf(b.m - m)};
}
__device__ MaxSum block_maxsum(MaxSum v, float* red /*[2*SAMP_WARPS]*/) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxSum t{__shfl_xor_sync(0xffffffffu, v.m, o), __shfl_xor_sync(0xffffffffu, v.s, o)};
    v = ms_combine(v, t);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) {
    red[2 * w] = v.m;
    red[2 * w + 1] = v.s;
  }
  __syncthreads();
  MaxSum r{red[2 * l], red[2 * l + 1]};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    MaxSum t{__shfl_xor_sync(0xffffffffu, r.m, o), __shfl_xor_sync(0xffffffffu, r.s, o)};
    r = ms_combine(r, t);
  }
  return r;
}
// argmax with lowest-index tie break (torch.argmax semantics relied on by greedy parity)
struct ValIdx {
  float v;
  int i;
};
__device__ __forceinline__ ValIdx vi_better(ValIdx a, ValIdx b) {
  if (b.i < 0) return a;
  if (a.i < 0) return b;
  if (b.v > a.v || (b.v == a.v && b.i < a.i)) return b;
  return a;
}
__device__ ValIdx block_argmax(ValIdx v, float* redf, int* redi) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx t{__shfl_xor_sync(0xffffffffu, v.v, o), __shfl_xor_sync(0xffffffffu, v.i, o)};
    v = vi_better(v, t);
  }
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) {
    redf[w] = v.v;
    redi[w] = v.i;
  }
  __syncthreads();
  ValIdx r{redf[l], redi[l]};
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ValIdx t{__shfl_xor_sync(0xffffffffu, r.v, o), __shfl_xor_sync(0xffffffffu, r.i, o)};
    r = vi_better(r, t);
  }
  return r;
}
__device__ float block_sumf(float v, float* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  float r = red[l];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}
__device__ int block_sumi(int v, int* red) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  __syncthreads();
  if (l == 0) red[w] = v;
  __syncthreads();
  int r = red[l];
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
  return r;
}

// ---------------------------------------------------------------- cluster reductions (SAMP_CL CTAs, rank order)
// xch: 4 words of shared memory per CTA at the same offset in every CTA of the cluster
template <class T>
__device__ __forceinline__ T dsmem_read(T* local, int rank) {
  return *cg::this_cluster().map_shared_rank(local, rank);
}
__device__ MaxSum cluster_maxsum(MaxSum v, float* red, float* xch) {
  v = block_maxsum(v, red);
  if (threadIdx.x == 0) {
    xch[0] = v.m;
    xch[1] = v.s;
  }
  cg::this_cluster().sync();
  MaxSum acc{dsmem_read(xch, 0), dsmem_read(xch + 1, 0)};
  for (int r = 1; r < SAMP_CL; ++r) acc = ms_combine(acc, MaxSum{dsmem_read(xch, r), dsmem_read(xch + 1, r)});
  cg::this_cluster().sync();  // xch may be rewritten
  return acc;
}
__device__ ValIdx cluster_argmax(ValIdx v, float* redf, int* redi, float* xch) {
  v = block_argmax(v, redf, redi);
  if (threadIdx.x == 0) {
    xch[0] = v.v;
    xch[1] = __int_as_float(v.i);
  }
  cg::this_cluster().sync();
  ValIdx acc{dsmem_read(xch, 0), __float_as_int(dsmem_read(xch + 1, 0))};
  for (int r = 1; r < SAMP_CL; ++r) acc = vi_better(acc, ValIdx{dsmem_read(xch, r), __float_as_int(dsmem_read(xch + 1, r))});
  cg::this_cluster().sync();
  return acc;
}
__device__ float cluster_sumf(float v, float* red, float* xch) {
  v = block_sumf(v, red);
  if (threadIdx.x == 0) xch[0] = v;
  cg::this_cluster().sync();
  float acc = dsmem_read(xch, 0);
  for (int r = 1; r < SAMP_CL; ++r) acc += dsmem_read(xch, r);
  cg::this_cluster().sync();
  return acc;
}
__device__ int cluster_sumi(int v, int* red, float* xch) {
  v = block_sumi(v, red);
  if (threadIdx.x == 0) xch[0] = __int_as_float(v);
  cg::this_cluster().sync();
  int acc = 0;
  for (int r = 0; r < SAMP_CL; ++r) acc += __float_as_int(dsmem_read(xch, r));
  cg::this_cluster().sync();
  return acc;
}

// ---------------------------------------------------------------- radix select helpers
__device__ __forceinline__ uint32_t f2key(float f) {  // order preserving float -> uint
  const uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float key2f(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7fffffffu) : ~k);
}

// Smallest key K (ascending) such that  sum_{key_i <= K} w_i  >= target  (strict: > target).
// keyf(i) -> uint32 key, wf(i) -> weight.  If never reached returns the largest present key.
// Cluster-wide: every CTA histograms its slice [lo, hi), the 8 histograms are summed in rank order through DSMEM
// into histsum, and every CTA runs the same scan on the same numbers.
template <class KeyF, class WF>
__device__ uint32_t select_weighted_asc(int lo, int hi, KeyF keyf, WF wf, float target, bool strict,
                                        float* hist /*[256]*/, float* histsum /*[256]*/, uint32_t* bcast) {
  uint32_t prefix = 0;
  float below = 0.f;  // weight of keys strictly below the current prefix range
  for (int round = 0; round < 4; ++round) {
    const int shift = 24 - 8 * round;
    cg::this_cluster().sync();  // nobody still reads last round's histogram
    if (threadIdx.x < 256) hist[threadIdx.x] = 0.f;
    __syncthreads();
    // Warp-aggregated histogram: in the first rounds nearly every key of a warp falls into the same one or two
    // digits (shared exponent bits), and 1024 threads hammering one shared-memory word serialise.  Lanes are grouped
    // by digit, each group is summed with shuffles (lane order: deterministic) and its leader issues ONE atomic.
    for (int i0 = lo; i0 < hi; i0 += SAMP_THREADS) {
      const int i = i0 + threadIdx.x;
      float w = 0.f;
      uint32_t digit = 0;
      if (i < hi) {
        w = wf(i);
        if (w > 0.f) {
          const uint32_t k = keyf(i);
          if (round == 0 || (k >> (shift + 8)) == (prefix >> (shift + 8))) digit = (k >> shift) & 255;
          else w = 0.f;
        }
      }
      uint32_t todo = __ballot_sync(0xffffffffu, w > 0.f);
      while (todo) {
        const int leader = __ffs(todo) - 1;
        const uint32_t ld = __shfl_sync(0xffffffffu, digit, leader);
        const bool mine = (w > 0.f) && digit == ld;
        float part = mine ? w : 0.f;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
        if ((threadIdx.x & 31) == leader) atomicAdd(&hist[ld], part);
        todo &= ~__ballot_sync(0xffffffffu, mine);
      }
    }
    cg::this_cluster().sync();
    if (threadIdx.x < 256) {
      float t = 0.f;
      for (int r = 0; r < SAMP_CL; ++r) t += dsmem_read(hist + threadIdx.x, r);
      histsum[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      float cum = below;
      int sel = -1, last_nonempty = -1;
      for (int b = 0; b < 256; ++b) {
        const float h = histsum[b];
        if (h > 0.f) last_nonempty = b;
        const float nc = cum + h;
        if (h > 0.f && (strict ? (nc > target) : (nc >= target))) {
          sel = b;
          break;
        }
        cum = nc;
      }
      if (sel < 0) {  // rounding: total mass < target -> keep everything: choose the largest key
        sel = last_nonempty < 0 ? 255 : last_nonempty;
        cum = below;
        for (int b = 0; b < sel; ++b) cum += histsum[b];
      }
      bcast[0] = (uint32_t)sel;
      bcast[1] = __float_as_uint(cum);
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    below = __uint_as_float(bcast[1]);
  }
  return prefix;
}

// k-th largest key (k >= 1) by count
template <class KeyF>
__device__ uint32_t select_kth_largest(int lo, int hi, KeyF keyf, int k, int* hist /*[256]*/, int* histsum /*[256]*/,
                                       uint32_t* bcast) {
  uint32_t prefix = 0;
  int above = 0;
  for (int round = 0; round < 4; ++round) {
    const int shift = 24 - 8 * round;
    cg::this_cluster().sync();
    if (threadIdx.x < 256) hist[threadIdx.x] = 0;
    __syncthreads();
    for (int i0 = lo; i0 < hi; i0 += SAMP_THREADS) {  // warp-aggregated: one atomic per distinct digit per warp
      const int i = i0 + threadIdx.x;
      uint32_t digit = 0xffffffffu;  // not counted
      if (i < hi) {
        const uint32_t key = keyf(i);
        if (round == 0 || (key >> (shift + 8)) == (prefix >> (shift + 8))) digit = (key >> shift) & 255;
      }
      const uint32_t grp = __match_any_sync(0xffffffffu, digit);
      if (digit != 0xffffffffu && (int)(threadIdx.x & 31) == __ffs(grp) - 1) atomicAdd(&hist[digit], __popc(grp));
    }
    cg::this_cluster().sync();
    if (threadIdx.x < 256) {
      int t = 0;
      for (int r = 0; r < SAMP_CL; ++r) t += dsmem_read(hist + threadIdx.x, r);
      histsum[threadIdx.x] = t;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      int cum = above, sel = 0;
      for (int b = 255; b >= 0; --b) {
        if (cum + histsum[b] >= k) {
          sel = b;
          break;
        }
        cum += histsum[b];
      }
      bcast[0] = (uint32_t)sel;
      bcast[1] = (uint32_t)cum;
    }
    __syncthreads();
    prefix |= bcast[0] << shift;
    above = (int)bcast[1];
  }
  return prefix;
}

// ---------------------------------------------------------------- Philox4x32-10
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, ctr.x), lo0 = 0xD2511F53u * ctr.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, ctr.z), lo1 = 0xCD9E8D57u * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += 0x9E3779B9u;
    key.y += 0xBB67AE85u;
  }
  return ctr;
}

// ---------------------------------------------------------------- the kernel
template <class LT>
__global__ void __cluster_dims__(SAMP_CL, 1, 1) __launch_bounds__(SAMP_THREADS, 1)
tgis_sampler_kernel(const LT* __restrict__ logits, int ld, int V, const SampleRow* __restrict__ rows,
                    uint32_t* __restrict__ seen_bitmap, int bitmap_words, float* __restrict__ scratch,
                    SampleOut* __restrict__ outs) {
  __shared__ float redf[2 * SAMP_WARPS];
  __shared__ int redi[SAMP_WARPS];
  __shared__ float histf[256];
  __shared__ float histsumf[256];
  __shared__ uint32_t bcast[4];
  __shared__ float xch[4];
  int* histi = reinterpret_cast<int*>(histf);
  int* histsumi = reinterpret_cast<int*>(histsumf);
  STL_ENTER(6);
  griddep_launch();
  griddep_wait();
  STL_WAITED();

  const int r = blockIdx.x / SAMP_CL;
  const int crank = (int)cg::this_cluster().block_rank();
  RowCtx<LT> c;
  c.p = rows[r];
  c.V = V;
  {
    const int per = ((V + SAMP_CL - 1) / SAMP_CL + 7) / 8 * 8;
    c.lo = min(V, crank * per);
    c.hi = min(V, c.lo + per);
  }
  const int lo = c.lo, hi = c.hi;
  c.x = logits + (size_t)c.p.logits_row * ld;
  c.seen = (c.p.seq_slot >= 0 && c.p.rep_penalty != 1.0f) ? seen_bitmap + (size_t)c.p.seq_slot * bitmap_words : nullptr;
  c.lenfac_m1 = (c.p.flags & SAMPLE_LENPEN) ? c.p.len_decay_factor : 0.f;
  c.mask_eos = c.p.n_out < c.p.min_tokens;
  c.typical = false;
  // FORCED rows (prompt logprobs): the token is given, only its raw logprob / rank / top-n are wanted
  const bool forced = (c.p.flags & SAMPLE_FORCED) != 0;
  const bool greedy = (c.p.flags & SAMPLE_GREEDY) && !forced;
  const bool want_lp = (c.p.flags & SAMPLE_LOGPROBS) != 0;
  float* y = scratch + (size_t)r * V;  // processed logits (sampling rows only)

  // ---- pass 1: raw max / sum-exp (+ greedy argmax of the processed logits in the same sweep)
  MaxSum ms{-INFINITY, 0.f};
  ValIdx best{-INFINITY, -1};
  const bool do_typ = (c.p.flags & SAMPLE_TYPICAL) != 0 && !forced;
  const bool greedy_fast = greedy && !do_typ;  // argmax fused into the first sweep
  for (int i0 = lo + threadIdx.x * 8; i0 < hi; i0 += SAMP_THREADS * 8) {
    float xs[8];
    load_x8(c.x, i0, xs);
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float x = xs[e];
      if (x > ms.m) {
        ms.s = ms.s * __expf(ms.m - x) + 1.f;
        ms.m = x;
      } else if (x != -INFINITY) {
        ms.s += __expf(x - ms.m);
      }
      if (greedy_fast) {
        const float yy = process(c, i0 + e, x);
        if (best.i < 0 || yy > best.v) best = {yy, i0 + e};  // ascending i: strict > keeps the lowest index
      }
    }
  }
  ms = cluster_maxsum(ms, redf, xch);
  c.raw_max = ms.m;
  c.raw_logz = logf(ms.s);

  // ---- typical-p threshold (R8): entropy, then weighted select over s = |-lp - H| ascending
  if (do_typ) {
    {
      float part = 0.f;
      for (int i = lo + threadIdx.x; i < hi; i += SAMP_THREADS) {
        const float lp = (load_x(c, i) - c.raw_max) - c.raw_logz;
        const float p = expf(lp);
        const float term = lp * p;
        if (term == term) part += term;  // nansum
      }
      c.ent = -cluster_sumf(part, redf, xch);
      const float rm = c.raw_max, lz = c.raw_logz, ent = c.ent;
      const LT* xx = c.x;
      auto keyf = [=](int i) {
        const float lp = (lt2f(xx[i]) - rm) - lz;
        return __float_as_uint(fabsf((-lp) - ent));
      };
      auto wf = [=](int i) { return expf((lt2f(xx[i]) - rm) - lz); };
      const uint32_t k = select_weighted_asc(lo, hi, keyf, wf, c.p.typical_p, false, histf, histsumf, bcast);
      c.typ_thr = __uint_as_float(k);
      c.typical = true;
    }
  }
  int token;
  if (forced) {
    token = (int)c.p.seed_lo;
  } else if (greedy) {
    if (!greedy_fast) {  // typical-p + greedy (method SAMPLE, temperature 0): argmax after the mask is known
      for (int i = lo + threadIdx.x; i < hi; i += SAMP_THREADS) {
        const float yy = process(c, i, load_x(c, i));
        if (best.i < 0 || yy > best.v) best = {yy, i};
      }
    }
    best = cluster_argmax(best, redf, redi, xch);
    token = best.i;
  } else {
    // ---- processed logits / temperature -> scratch, running max
    float mx = -INFINITY;
    const float temp = c.p.temperature;
    for (int i = lo + threadIdx.x; i < hi; i += SAMP_THREADS) {
      float v = process(c, i, load_x(c, i));
      v = __fdiv_rn(v, temp);
      y[i] = v;
      mx = fmaxf(mx, v);
    }
    {
      MaxSum t = cluster_maxsum(MaxSum{mx, 0.f}, redf, xch);
      mx = t.m;
    }
    // ---- top-k (S6): keep y >= k-th largest value
    float lo_thr = -INFINITY;
    if (c.p.top_k > 0 && c.p.top_k < V) {
      auto keyf = [=](int i) { return f2key(y[i]); };
      lo_thr = key2f(select_kth_largest(lo, hi, keyf, c.p.top_k, histi, histsumi, bcast));
    }
    // ---- top-p (S6): drop the low-probability tail whose cumulative mass <= 1 - p
    if (c.p.top_p < 1.0f) {
      float part = 0.f;
      for (int i = lo + threadIdx.x; i < hi; i += SAMP_THREADS) {
        const float v = y[i];
        if (v >= lo_thr) part += expf(v - mx);
      }
      const float z = cluster_sumf(part, redf, xch);
      const float lt = lo_thr;
      auto keyf = [=](int i) { return f2key(y[i]); };
      auto wf = [=](int i) {
        const float v = y[i];
        return (v >= lt) ? expf(v - mx) / z : 0.f;
      };
      const uint32_t k = select_weighted_asc(lo, hi, keyf, wf, 1.0f - c.p.top_p, true, histf, histsumf, bcast);
      lo_thr = fmaxf(lo_thr, key2f(k));
    }
    // ---- exponential race == Gumbel max over the kept set
    const uint2 key = make_uint2(c.p.seed_lo, c.p.seed_hi);
    ValIdx bs{-INFINITY, -1};
    for (int i = lo + threadIdx.x; i < hi; i += SAMP_THREADS) {
      const float v = y[i];
      if (v >= lo_thr && v != -INFINITY) {
        const uint4 rnd = philox4x32(make_uint4((uint32_t)i, c.p.step, 0u, 0u), key);
        const float u = ((float)(rnd.x >> 8) + 0.5f) * (1.0f / 16777216.0f);
        const float score = (v - mx) - logf(-logf(u));
        if (bs.i < 0 || score > bs.v) bs = {score, i};
      }
    }
    bs = cluster_argmax(bs, redf, redi, xch);
    token = bs.i;
    if (token < 0) {  // everything masked (degenerate): fall back to raw argmax
      ValIdx b2{-INFINITY, -1};
      for (int i = lo + threadIdx.x; i < hi; i += SAMP_THREADS) {
        const float v = load_x(c, i);
        if (b2.i < 0 || v > b2.v) b2 = {v, i};
      }
      b2 = cluster_argmax(b2, redf, redi, xch);
      token = b2.i;
    }
  }

  // ---- logprob / rank / top-n on the RAW log-softmax (S1, S8)
  float tok_lp = 0.f;
  int rank = 0;
  if (want_lp) {
    tok_lp = (load_x(c, token) - c.raw_max) - c.raw_logz;
    int cnt = 0;
    for (int i = lo + threadIdx.x; i < hi; i += SAMP_THREADS) {
      const float lp = (load_x(c, i) - c.raw_max) - c.raw_logz;
      cnt += (lp >= tok_lp) ? 1 : 0;
    }
    rank = cluster_sumi(cnt, redi, xch);
  }
  SampleOut* o = outs + r;
  const int n_topn = min(c.p.n_topn, MAX_TOPN);
  float prev_v = INFINITY;
  int prev_i = -1;
  for (int n = 0; n < n_topn; ++n) {
    // next element in (value desc, index asc) order strictly after (prev_v, prev_i)
    ValIdx b{-INFINITY, -1};
    for (int i = lo + threadIdx.x; i < hi; i += SAMP_THREADS) {
      const float v = load_x(c, i);
      const bool after = (v < prev_v) || (v == prev_v && i > prev_i);
      if (after && (b.i < 0 || v > b.v)) b = {v, i};
    }
    b = cluster_argmax(b, redf, redi, xch);
    prev_v = b.v;
    prev_i = b.i;
    if (threadIdx.x == 0 && crank == 0) {
      o->topn_ids[n] = b.i;
      o->topn_lps[n] = (b.v - c.raw_max) - c.raw_logz;
    }
  }
  if (threadIdx.x == 0 && crank == 0) {
    o->token = token;
    o->logprob = tok_lp;
    o->rank = rank;
    o->n_topn = n_topn;
    if (c.p.seq_slot >= 0 && token >= 0)
      atomicOr(&seen_bitmap[(size_t)c.p.seq_slot * bitmap_words + (token >> 5)], 1u << (token & 31));
  }
  STL_EXIT();
}

cudaError_t sampler_launch(const void* logits, int logits_bf16, int ld, int vocab, const SampleRow* rows, int n_rows,
                           const uint32_t* seen_bitmap, int bitmap_words, float* scratch, SampleOut* out,
                           cudaStream_t stream) {
  if (n_rows <= 0) return cudaSuccess;
  if (vocab % 8 != 0 || ld % 8 != 0) return cudaErrorInvalidValue;
  if (logits_bf16)
    return launch_k(tgis_sampler_kernel<__nv_bfloat16>, dim3(n_rows * SAMP_CL), dim3(SAMP_THREADS), 0, stream,
                    static_cast<const __nv_bfloat16*>(logits), ld, vocab, rows, const_cast<uint32_t*>(seen_bitmap),
                    bitmap_words, scratch, out);
  return launch_k(tgis_sampler_kernel<float>, dim3(n_rows * SAMP_CL), dim3(SAMP_THREADS), 0, stream,
                  static_cast<const float*>(logits), ld, vocab, rows, const_cast<uint32_t*>(seen_bitmap), bitmap_words,
                  scratch, out);
}

size_t sampler_scratch_floats(int vocab) { return (size_t)vocab; }

// ---------------------------------------------------------------- seen-token bitmap maintenance
__global__ void bitmap_clear_kernel(uint32_t* bm, int words) {
  griddep_launch();
  griddep_wait();
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < words; i += gridDim.x * blockDim.x) bm[i] = 0u;
}
__global__ void bitmap_set_kernel(uint32_t* bm, int words, const int32_t* slots, const int32_t* tokens, int n) {
  griddep_launch();
  griddep_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int s = slots[i], t = tokens[i];
  if (s < 0 || t < 0 || (t >> 5) >= words) return;
  atomicOr(&bm[(size_t)s * words + (t >> 5)], 1u << (t & 31));
}
cudaError_t bitmap_clear_launch(uint32_t* bitmap, int bitmap_words, int slot, cudaStream_t stream) {
  return launch_k(bitmap_clear_kernel, dim3(8), dim3(256), 0, stream, bitmap + (size_t)slot * bitmap_words, bitmap_words);
}
cudaError_t bitmap_set_launch(uint32_t* bitmap, int bitmap_words, const int32_t* slots, const int32_t* tokens, int n,
                              cudaStream_t stream) {
  if (n <= 0) return cudaSuccess;
  return launch_k(bitmap_set_kernel, dim3((n + 255) / 256), dim3(256), 0, stream, bitmap, bitmap_words, slots, tokens, n);
}

}  // namespace tgis
