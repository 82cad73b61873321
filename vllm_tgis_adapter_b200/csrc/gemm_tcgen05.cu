// Dense bf16 linear layer  Y[T,N] = X[T,K] · W[N,K]^T  on the 5th-gen tensor cores (tcgen05 + TMEM + TMA).
//
// Replaces, for the decode/prefill hot path, what the reference reaches through
//   vllm: model_executor/layers/linear.py (QKVParallelLinear / RowParallelLinear / MergedColumnParallelLinear
//   -> F.linear -> cuBLAS)  behind grpc_server.py:222 `self.engine.generate(...)`   (SURVEY.md §2.2 K3/K8/K9/K10).
//
// Design (B200-first, not a cuBLAS clone):
//  * swap-AB: the WEIGHT tile [128 rows of N x 64 of K] is the MMA "A" operand (fills the 128 TMEM lanes), the
//    TOKEN tile [BT x 64] is the "B" operand (UMMA_N = BT in {16..256}).  A decode step (T = 32..256 tokens)
//    therefore streams every weight byte exactly once at HBM rate with full-width MMAs.
//  * both operands K-major, 128-byte swizzled, staged by TMA into a STAGES-deep mbarrier ring.
//  * warp roles: warp0 = TMA producer, warp1 = MMA issuer (single elected lane) + TMEM owner,
//    warps2..5 = epilogue (TMEM -> registers -> bf16 global).  TMEM accumulator is double-buffered so the
//    epilogue of unit i overlaps the mainloop of unit i+1.
//  * persistent stream-K: the (tile, k-block) iteration space is cut into gridDim.x equal contiguous ranges, so all
//    148 SMs stream the same number of weight bytes whatever N is.  A tile split across CTAs is reduced by the
//    LAST-arriving CTA summing the fp32 partials in fixed CTA order -> bit-deterministic and independent of timing.
//  * few-tile GEMMs (qkv / o / down at decode: every tile is split by the same factor s <= 8): the s CTAs of a tile
//    are launched as ONE thread-block cluster and the split-K reduction never leaves the chip: every CTA parks its
//    fp32 partial in its own shared memory (the drained smem ring), one cluster barrier, then CTA r reduces the tokens
//    t = r (mod s) by reading its peers through distributed shared memory in rank order (same sums, same order as the
//    global-memory fix-up: bit-identical) and writes the final rows.
#include <string.h>

#include "kernels.h"
#include "launch.cuh"
#include "ptx.cuh"

namespace tgis {

TGIS_STL_DEFINE(gemm)

constexpr int GEMM_BN = 128;  // weight rows per tile (UMMA_M)
constexpr int GEMM_BK = 64;   // k per stage (one 128-byte swizzle row of bf16)
constexpr int GEMM_THREADS = 192;
constexpr int GEMM_THREADS_NORM = 256;  // + 2 activation-transform warps (fused RMSNorm consumer, GemmNorm::h)

// NW = weight tiles (of GEMM_BN = 128 rows) per unit that share ONE activation tile: every unit re-reads its [BT x 64]
// activation tile per k-block, and with NW = 2 that tile feeds two MMAs (two accumulators), halving the activation bytes
// per weight byte.  An experiment (TGIS_GEMM_NW=2, see gemm_nw()): measured slower, so NW = 1 is what runs.
template <int BT, int NW>
struct GemmCfg {
  static constexpr int W_TILE_BYTES = GEMM_BN * GEMM_BK * 2;
  static constexpr int W_BYTES = NW * W_TILE_BYTES;
  static constexpr int X_BYTES = BT * GEMM_BK * 2;
  static constexpr int STAGE_BYTES = W_BYTES + X_BYTES;
#ifndef TGIS_GEMM_SMEM_KB
#define TGIS_GEMM_SMEM_KB 200
#endif
  static constexpr int STAGES_RAW = (TGIS_GEMM_SMEM_KB * 1024) / STAGE_BYTES;
#ifndef TGIS_GEMM_DECODE_STAGES
#define TGIS_GEMM_DECODE_STAGES 8
#endif
  static constexpr int STAGES_CAP = BT <= 64 ? TGIS_GEMM_DECODE_STAGES : 12;
  static constexpr int STAGES = STAGES_RAW > STAGES_CAP ? STAGES_CAP : STAGES_RAW;
  static constexpr int ACC_COLS = NW * BT;                               // TMEM columns of one unit's accumulators
  static constexpr int N_ACC = (2 * ACC_COLS <= 512) ? 2 : 1;            // double-buffered when it fits the 512 columns
  static constexpr int TMEM_COLS = (N_ACC * ACC_COLS) < 32 ? 32 : (N_ACC * ACC_COLS);
  static constexpr int PART_FLOATS = NW * BT * GEMM_BN;                  // fp32 partial of one unit
  static constexpr int BAR_BYTES = 512;   // mbarriers (full, empty, tmem, ready), TMEM base, flag
  static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + BAR_BYTES + 256 /*rstd*/;
  // fused-norm consumers append the norm weights (+ K * 2 bytes, K <= 8192)
  static constexpr int SMEM_MAX = (SMEM_BYTES + 16384) < 227 * 1024 ? (SMEM_BYTES + 16384) : 227 * 1024;
};

// Prefill-shaped launches (more than one token tile): "data-parallel waves + stream-K tail".  Linear tile index i
// (see tile_decode) -> the first dp_waves * ncta tiles are whole tiles dealt round-robin, wave w = tiles
// [w * ncta, (w+1) * ncta): at any moment the 148 CTAs work on ONE compact 8 x ~18 rectangle of (weight tile, token
// tile) pairs, so both operands of a wave stay in L2 (with contiguous per-CTA ranges the CTAs drift apart, the working
// set is every weight tile + every activation tile, and ncu showed 11 GB of DRAM reads for 0.3 GB of operands).  The
// remaining < ncta tiles are cut stream-K style as before.  Decode-shaped launches have dp_waves = 0: unchanged.
//
// Decode-shaped launches with at least one tile per CTA (gate_up: 224 tiles on 148 CTAs, lm_head: 1002) use the same two
// regions in the OPPOSITE order ("sk_first"): every CTA first works through its < 1 tile share of the stream-K region --
// the shared tiles, whose last-arriver fix-up (atomic + L2 round trip + ordered sum) then runs in the epilogue warps
// UNDER the main loop of the following whole tile -- and ends on a whole tile with a direct epilogue.  With the pure
// stream-K cut most CTAs END on a shared tile piece, and the fix-ups (4-6.5 us at 32 tokens, growing with T) are the
// launch's tail.  Same tiles, same split points inside a shared tile's k-range order, same rank-ordered sums.
struct GemmSched {
  int n_tiles, t_tiles, KB, ncta;
  int dp_waves;        // whole-tile waves
  int sk_tile0;        // first tile of the stream-K region
  long long sk_total;  // (tile, k-block) units in the stream-K region
  bool sk_first;       // stream-K region before the whole-tile waves
  __device__ GemmSched(int n_tiles_, int t_tiles_, int KB_, int ncta_, bool sk_first_)
      : n_tiles(n_tiles_), t_tiles(t_tiles_), KB(KB_), ncta(ncta_) {
    const int tiles = n_tiles * t_tiles;
    sk_first = sk_first_ && t_tiles == 1 && tiles >= ncta;
    dp_waves = (t_tiles > 1 || sk_first) ? tiles / ncta : 0;
    sk_tile0 = dp_waves * ncta;
    sk_total = (long long)(tiles - sk_tile0) * KB;
  }
};
// linear tile index -> (weight tile, token tile).  One token tile: identity.  Otherwise blocks of 8 weight tiles, token
// tile next, weight tile fastest: 148 consecutive indices = 8 weight tiles x ~18 token tiles.
__device__ __forceinline__ void tile_decode(const GemmSched& sc, int tile, int& n_tile, int& t_tile) {
  if (sc.t_tiles == 1) {
    n_tile = tile;
    t_tile = 0;
    return;
  }
  constexpr int NB = 8;
  const int full = NB * sc.t_tiles;
  const int nb = tile / full, rem = tile - nb * full;
  const int nb_n = min(NB, sc.n_tiles - nb * NB);
  t_tile = rem / nb_n;
  n_tile = nb * NB + rem - t_tile * nb_n;
}

struct UnitIter {
  // whole tiles wave * ncta + cta for wave < dp_waves, then the contiguous range [pos, end) of the stream-K region's
  // (tile, kblock) space owned by this CTA
  long long pos, end;
  int kb_per_tile, wave, dp_waves, cta, ncta, sk_tile0;
  bool first, sk_first;
  __device__ UnitIter(const GemmSched& sc, int cta_)
      : kb_per_tile(sc.KB), wave(0), dp_waves(sc.dp_waves), cta(cta_), ncta(sc.ncta), sk_tile0(sc.sk_tile0), first(true),
        sk_first(sc.sk_first) {
    pos = (sc.sk_total * cta_) / sc.ncta;
    end = (sc.sk_total * (cta_ + 1)) / sc.ncta;
  }
  __device__ bool next(int& tile, int& kb0, int& kb1, int& slot) {
    if (wave < dp_waves && (!sk_first || pos >= end)) {
      tile = wave * ncta + cta;
      kb0 = 0;
      kb1 = kb_per_tile;
      slot = 0;
      ++wave;
      return true;
    }
    if (pos >= end) return false;
    tile = sk_tile0 + (int)(pos / kb_per_tile);
    kb0 = (int)(pos % kb_per_tile);
    long long rem = end - pos;
    int room = kb_per_tile - kb0;
    int take = rem < room ? (int)rem : room;
    kb1 = kb0 + take;
    slot = first ? 0 : 1;
    first = false;
    pos += take;
    return true;
  }
};

__device__ __forceinline__ int unit_owner(long long p, long long total, int ncta) {
  return (int)(((p + 1) * ncta - 1) / total);
}

#ifdef TGIS_GEMM_TIMELINE
__device__ unsigned long long g_gemm_timeline[4][16];
__device__ __forceinline__ unsigned long long gtimer() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
#define TL(slot)                                                                            \
  do {                                                                                      \
    if ((blockIdx.x == 0 || blockIdx.x == gridDim.x / 2) && (slot) < 16)                    \
      g_gemm_timeline[blockIdx.x == 0 ? 0 : 1][slot] = gtimer();                            \
  } while (0)
#else
#define TL(slot) do {} while (0)
#endif

// SwiGLU on one (gate, up) accumulator pair with the rounding points of the unfused path
// (bf16 GEMM outputs -> bf16(silu(g)) -> bf16(. * u); vllm activation_kernels.cu / HF LlamaMLP)
__device__ __forceinline__ __nv_bfloat16 swiglu_bf16(float gate_acc, float up_acc) {
  const float g = bf16_round(gate_acc), u = bf16_round(up_acc);
  const float sl = bf16_round(g / (1.0f + expf(-g)));
  return __float2bfloat16_rn(sl * u);
}

// Fused RoPE + KV-cache scatter of the qkv projection's final epilogue (tile = one 128-dim head; arithmetic and rounding
// points of rope_kvwrite_kernel).  Called by a full warp with a warp-uniform (head, t): this thread holds the fp32 sums of
// dims r4..r4+3 of token t; the rotary partner dims (+-64) live in lane ^ 16.
__device__ __forceinline__ void qkv_rope_store(const GemmRope& rope, int head, int t, int r4, const float (&av)[4],
                                               __nv_bfloat16* y_dst) {
  const int lane = lane_id();
  float x[4], pr[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) x[e] = bf16_round(av[e]);
#pragma unroll
  for (int e = 0; e < 4; ++e) pr[e] = __shfl_xor_sync(0xffffffffu, x[e], 16);
  const int d = r4 & 63;
  if (head < rope.n_q + rope.n_kv) {
    const int pos = __ldg(rope.positions + t);
    const uint2 cs = __ldg(reinterpret_cast<const uint2*>(rope.cos_sin + (size_t)pos * HEAD_DIM + d));
    const uint2 sn = __ldg(reinterpret_cast<const uint2*>(rope.cos_sin + (size_t)pos * HEAD_DIM + 64 + d));
    const float cc[4] = {__uint_as_float(cs.x << 16), __uint_as_float(cs.x & 0xffff0000u), __uint_as_float(cs.y << 16),
                         __uint_as_float(cs.y & 0xffff0000u)};
    const float ss[4] = {__uint_as_float(sn.x << 16), __uint_as_float(sn.x & 0xffff0000u), __uint_as_float(sn.y << 16),
                         __uint_as_float(sn.y & 0xffff0000u)};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float a = bf16_round(x[e] * cc[e]), b = bf16_round(pr[e] * ss[e]);
      x[e] = lane < 16 ? a - b : a + b;  // lo' = x1 c - x2 s ; hi' = x2 c + x1 s
    }
  }
  __nv_bfloat162 p0 = __floats2bfloat162_rn(x[0], x[1]), p1 = __floats2bfloat162_rn(x[2], x[3]);
  uint2 pk;
  pk.x = *reinterpret_cast<uint32_t*>(&p0);
  pk.y = *reinterpret_cast<uint32_t*>(&p1);
  *reinterpret_cast<uint2*>(y_dst) = pk;
  if (head >= rope.n_q) {
    const int slot = __ldg(rope.slot_mapping + t);
    if (slot >= 0) {
      const int blk = slot / KV_BLOCK, off = slot % KV_BLOCK;
      const int chunk = r4 >> 3, sub = r4 & 7;  // 16-byte chunk of the head, 0 or 4 inside it
      if (head < rope.n_q + rope.n_kv) {  // K tile [chunk][token][8]
        __nv_bfloat16* kb = rope.k_cache + ((size_t)blk * rope.n_kv + (head - rope.n_q)) * (KV_BLOCK * HEAD_DIM);
        *reinterpret_cast<uint2*>(kb + (chunk * KV_BLOCK + off) * 8 + sub) = pk;
      } else {  // V tile [token][chunk ^ (token & 7)][8]
        __nv_bfloat16* vb =
            rope.v_cache + ((size_t)blk * rope.n_kv + (head - rope.n_q - rope.n_kv)) * (KV_BLOCK * HEAD_DIM);
        *reinterpret_cast<uint2*>(vb + off * HEAD_DIM + ((chunk ^ (off & 7)) * 8) + sub) = pk;
      }
    }
  }
}

template <int BT, int NW>
__global__ void __launch_bounds__(GEMM_THREADS_NORM, 1)
gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap wmap, const __grid_constant__ CUtensorMap xmap,
                         void* __restrict__ Yv, int ldy, int T, int N, int K, float* __restrict__ ws,
                         int* __restrict__ counters, int stream_weights, int out_f32,
                         const __grid_constant__ CUtensorMap next_wmap, GemmNext nxt, int cluster_split,
                         GemmRope rope, GemmNorm norm) {
  // output: bf16 (rounded once from the fp32 accumulator, = F.linear in model dtype) or raw fp32 (lm_head logits)
  __nv_bfloat16* __restrict__ Y = reinterpret_cast<__nv_bfloat16*>(Yv);
  float* __restrict__ Yf = reinterpret_cast<float*>(Yv);
  using Cfg = GemmCfg<BT, NW>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int BNW = GEMM_BN * NW;            // weight rows per unit
  constexpr int ACC_COLS = Cfg::ACC_COLS;
  constexpr int N_ACC = Cfg::N_ACC;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_w = smem;
  uint8_t* smem_x = smem + STAGES * Cfg::W_BYTES;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
  uint64_t* full_bar = bars;
  uint64_t* empty_bar = bars + STAGES;
  uint64_t* tmem_full = bars + 2 * STAGES;
  uint64_t* tmem_empty = bars + 2 * STAGES + 2;
  uint32_t* tmem_base_smem = reinterpret_cast<uint32_t*>(bars + 2 * STAGES + 4);
  int* flag_smem = reinterpret_cast<int*>(tmem_base_smem + 1);
  uint64_t* ready_bar = bars + 40;  // [STAGES <= 12] fused-norm consumer: token tile of the stage normalised in place
  uint64_t* xfull_bar = bars + 52;  // [STAGES <= 12] fused-norm consumer: raw token tile landed (weights: full_bar)
  float* rstd_smem = reinterpret_cast<float*>(smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES);  // [64]
  uint8_t* wn_smem = smem + STAGES * Cfg::STAGE_BYTES + Cfg::BAR_BYTES + 256;                      // [K] bf16 norm weights
  const bool norm_in = norm.h != nullptr;  // the token operand is RMSNorm(h): warps 6..7 rewrite each landed stage in place

  const int warp = threadIdx.x >> 5;
  STL_ENTER(2 | (N << 8));
  int* stl_slot_smem = flag_smem + 1;
  STL_SHARE(stl_slot_smem);  // read by the other warps after the set-up barrier
  if (threadIdx.x == 0) TL(0);  // kernel entry
  const int n_tiles = (N + BNW - 1) / BNW;     // units along N
  const int t_tiles = (T + BT - 1) / BT;
  const int KB = (K + GEMM_BK - 1) / GEMM_BK;
  const int ncta = gridDim.x, cta = blockIdx.x;
  const GemmSched sched(n_tiles, t_tiles, KB, ncta, (stream_weights & 2) != 0);
  // cluster mode, push variant: every CTA sends its partial rows straight into the OWNER's receive buffer (tokens
  // t = r (mod s) belong to rank r), ONE cluster barrier, then each owner sums its own shared memory in rank order -- no
  // remote read latency and no exit barrier (pull variant: park locally, barrier, read peers, barrier)
  const bool cl_push = (stream_weights & 4) != 0;
  stream_weights &= 1;
  const long long total = sched.sk_total;  // stream-K region (== everything for decode-shaped launches)

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&wmap);
    tma_prefetch_desc(&xmap);
  }
  if (warp == 1) {
    if (elect_one()) {
      for (int i = 0; i < STAGES; ++i) {
        mbar_init(&full_bar[i], 1);
        mbar_init(&empty_bar[i], 1);
        mbar_init(&ready_bar[i], 1);  // the transform warp that owns the stage's k-block (fused-norm consumer)
        mbar_init(&xfull_bar[i], 1);
      }
      for (int i = 0; i < 2; ++i) {
        mbar_init(&tmem_full[i], 1);
        mbar_init(&tmem_empty[i], 4);
      }
      fence_barrier_init();
    }
    __syncwarp();
    tmem_alloc<Cfg::TMEM_COLS>(tmem_base_smem);
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  // push variant: peers write into this CTA's shared memory as soon as THEY are done -- every CTA of the cluster must be
  // running before the first remote store (arrive now, wait right before pushing: free by then)
  if (cl_push) asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  const uint32_t tmem_base = *tmem_base_smem;
  griddep_launch();  // PDL: the next kernel may start its prologue now
  if (threadIdx.x == 0) TL(1);  // setup done (barriers, TMEM)
  int cl_tile = -1, cl_tvalid = 0;  // cluster mode: the split tile this CTA contributed to (epilogue warps)

  // ---- fused residual-stream RMSNorm (GemmNorm consumer side): xmap points at the residual stream h, so the producer
  // warp's TMA stages the RAW [BT x 64] token tile (own barrier xfull_bar: it lands long before the weight tile).  A warp
  // rewrites one landed stage IN PLACE as bf16(bf16(h * rstd) * w) -- 16-byte chunk c of row r sits at
  // r * 128 + ((c ^ (r & 7)) << 4) under the 128-byte swizzle; lane -> chunk lane & 7, rows (lane >> 3) + 4 j -- and
  // signals ready_bar, which the MMA warp waits for.  Warps 6 / 7 take the even / odd k-blocks; when the dependency
  // resolves a whole ring of tiles lands at once, so the (then idle) epilogue warps take k-blocks 2..5 of that burst.
  constexpr int NJ = BT <= GEMM_NORM_MAX_T ? BT / 4 : 1;
  const int norm_n_kb = (int)((total * (cta + 1)) / ncta - (total * cta) / ncta);  // decode-shaped: no whole-tile waves
  const int norm_kb0 = (int)(((total * cta) / ncta) % KB);
  const bool norm_helpers = norm_in && STAGES >= 6 && norm_n_kb >= 6;
  auto norm_stage = [&](int i, const float (&rs)[NJ]) {
    const int lane = lane_id();
    const int c = lane & 7, r0 = lane >> 3;
    const int stage = i % STAGES;
    const uint32_t phase = (uint32_t)(i / STAGES) & 1u;
    const int kb = (norm_kb0 + i) % KB;
    uint32_t wb[4];
    asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n"
                 : "=r"(wb[0]), "=r"(wb[1]), "=r"(wb[2]), "=r"(wb[3])
                 : "r"(smem_u32(wn_smem) + (uint32_t)((kb * GEMM_BK + c * 8) * 2)));
    mbar_wait(&xfull_bar[stage], phase);
    const uint32_t st_base = smem_u32(smem_x) + (uint32_t)(stage * Cfg::X_BYTES);
    uint32_t hb[NJ][4];
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int r = r0 + 4 * j;
      asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];\n"
                   : "=r"(hb[j][0]), "=r"(hb[j][1]), "=r"(hb[j][2]), "=r"(hb[j][3])
                   : "r"(st_base + (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4))));
    }
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int r = r0 + 4 * j;
      uint32_t ob[4];
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        const float h0 = __uint_as_float(hb[j][e] << 16), h1 = __uint_as_float(hb[j][e] & 0xffff0000u);
        const float w0 = __uint_as_float(wb[e] << 16), w1 = __uint_as_float(wb[e] & 0xffff0000u);
        const __nv_bfloat162 o2 = __floats2bfloat162_rn(bf16_round(h0 * rs[j]) * w0, bf16_round(h1 * rs[j]) * w1);
        ob[e] = *reinterpret_cast<const uint32_t*>(&o2);
      }
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};\n" ::"r"(st_base + (uint32_t)(r * 128 + ((c ^ (r & 7)) << 4))),
                   "r"(ob[0]), "r"(ob[1]), "r"(ob[2]), "r"(ob[3])
                   : "memory");
    }
    fence_proxy_async();  // generic-proxy stores -> visible to the tensor core's async-proxy reads
    __syncwarp();
    if (lane == 0) mbar_arrive(&ready_bar[stage]);
  };

  if (warp == 0) {
    // ===================== TMA producer =====================
    // The CTA's k-blocks are one contiguous range [begin, end) of the global (tile, k-block) space.  Weights are
    // static, so the first STAGES weight tiles are requested BEFORE the grid-dependency wait: under PDL this
    // overlaps the previous kernel's tail (and the small kernel in between) with useful HBM traffic.
    if (elect_one()) {
      const uint64_t pol_w = policy_evict_first();
      const uint64_t pol_x = policy_evict_last();
      const long long begin = (total * cta) / ncta, end = (total * (cta + 1)) / ncta;
      const int n_kb = sched.dp_waves * KB + (int)(end - begin);
      // (tile, kb) cursors advanced incrementally: no divisions on the per-k-block issue path.  A cursor walks the
      // CTA's whole tiles (wave order) and then its stream-K range; rows are re-derived only when the tile changes.
      struct Cur {
        int wave, tile, kb, w_row, x_row;
        int sk_left;  // sk_first: k-blocks left in the stream-K share (then the whole-tile waves)
      };
      auto cur_set_tile = [&](Cur& c, int tile) {
        c.tile = tile;
        int n_tile, t_tile;
        tile_decode(sched, tile, n_tile, t_tile);
        c.w_row = n_tile * BNW;
        c.x_row = t_tile * BT;
      };
      auto cur_init = [&](Cur& c) {
        c.wave = 0;
        c.sk_left = sched.sk_first ? (int)(end - begin) : 0;
        if (sched.dp_waves > 0 && c.sk_left == 0) {
          c.kb = 0;
          cur_set_tile(c, cta);
        } else {
          c.kb = (int)(begin % KB);
          cur_set_tile(c, sched.sk_tile0 + (int)(begin / KB));
        }
      };
      auto cur_next = [&](Cur& c) {
        if (c.sk_left > 0) {  // sk_first: inside the stream-K share
          if (--c.sk_left == 0) {  // -> first whole tile
            c.kb = 0;
            if (sched.dp_waves > 0) cur_set_tile(c, cta);
            return;
          }
          if (++c.kb == KB) {
            c.kb = 0;
            cur_set_tile(c, c.tile + 1);
          }
          return;
        }
        if (++c.kb < KB) return;
        c.kb = 0;
        if (c.wave < sched.dp_waves) {
          if (++c.wave < sched.dp_waves) {
            cur_set_tile(c, c.wave * ncta + cta);
          } else {  // into the stream-K tail
            c.kb = (int)(begin % KB);
            cur_set_tile(c, sched.sk_tile0 + (int)(begin / KB));
          }
        } else {
          cur_set_tile(c, c.tile + 1);
        }
      };
      auto issue_w = [&](const Cur& c, int stage) {
#pragma unroll
        for (int h = 0; h < NW; ++h) {  // NW boxes of 128 rows (rows beyond N are zero-filled by TMA and still counted)
          void* dst = smem_w + stage * Cfg::W_BYTES + h * Cfg::W_TILE_BYTES;
          if (stream_weights) tma_load_2d_hint(&wmap, &full_bar[stage], dst, c.kb * GEMM_BK, c.w_row + h * GEMM_BN, pol_w);
          else tma_load_2d(&wmap, &full_bar[stage], dst, c.kb * GEMM_BK, c.w_row + h * GEMM_BN);
        }
      };
      const int n_pre = n_kb < STAGES ? n_kb : STAGES;
      Cur wc, xc;
      cur_init(wc);
      xc = wc;
      // fused-norm consumer: the (small, L2-resident) token tile gets its own barrier, so it is normalised while the
      // stage's weight tile is still in flight and the rewrite adds nothing to the time a ring slot stays occupied
      const uint32_t w_tx = norm_in ? (uint32_t)Cfg::W_BYTES : (uint32_t)Cfg::STAGE_BYTES;
      for (int i = 0; i < n_pre; ++i) {
        mbar_arrive_expect_tx(&full_bar[i], w_tx);
        issue_w(wc, i);
        cur_next(wc);
      }
      TL(2);           // weight prefetch issued
      griddep_wait();  // activations (and everything the epilogue will touch) are now final
      TL(3);           // dependency wait returned
      if (threadIdx.x == 0) STL_WAITED();
      int stage = 0;
      uint32_t phase = 0;
      for (int i = 0; i < n_kb; ++i) {
        if (i >= n_pre) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          mbar_arrive_expect_tx(&full_bar[stage], w_tx);
          issue_w(wc, stage);
          cur_next(wc);
        }
        uint64_t* x_bar = &full_bar[stage];
        if (norm_in) {
          x_bar = &xfull_bar[stage];
          mbar_arrive_expect_tx(x_bar, Cfg::X_BYTES);
        }
        tma_load_2d_hint(&xmap, x_bar, smem_x + stage * Cfg::X_BYTES, xc.kb * GEMM_BK, xc.x_row, pol_x);
        cur_next(xc);
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      TL(4);  // last TMA issued
      STL_EXTRA(STL_MINE(), 1);
      // Cross-kernel weight prefetch: HBM goes idle while this kernel drains its pipeline, reduces split tiles and
      // exits, and the next GEMM of the layer stack needs ~10 us before its own loads are in flight.  So the boxes
      // that the next GEMM's CTAs will read FIRST are pulled into the 126 MB L2 now (TMA prefetch, no smem, no
      // completion tracking): next CTA c2 starts at k-block (total2 * c2 / grid2) of its (tile, k-block) space.
      if (nxt.kb_prefetch > 0) {
        const long long total2 = (long long)nxt.n_tiles * nxt.KB;
        for (int c2 = cta; c2 < nxt.grid; c2 += ncta) {
          const long long b2 = (total2 * c2) / nxt.grid, e2 = (total2 * (c2 + 1)) / nxt.grid;
          const int n2 = (int)((e2 - b2) < nxt.kb_prefetch ? (e2 - b2) : nxt.kb_prefetch);
          int tile2 = (int)(b2 / nxt.KB), kb2 = (int)(b2 % nxt.KB);
          for (int i = 0; i < n2; ++i) {
            for (int h = 0; h < NW; ++h) tma_prefetch_2d(&next_wmap, kb2 * GEMM_BK, tile2 * BNW + h * GEMM_BN);
            if (++kb2 == nxt.KB) { kb2 = 0; ++tile2; }
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = make_idesc_bf16(GEMM_BN, BT);
    UnitIter it(sched, cta);
    int tile, kb0, kb1, slot;
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_bits = 0;  // per-buffer phase parity
    bool stl_first_done = false;
    while (it.next(tile, kb0, kb1, slot)) {
      mbar_wait(&tmem_empty[acc], ((acc_bits >> acc) & 1) ^ 1);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * ACC_COLS;
      for (int kb = kb0; kb < kb1; ++kb) {
        mbar_wait(&full_bar[stage], phase);
        if (norm_in) mbar_wait(&ready_bar[stage], phase);  // the token tile has been normalised in place
        tc_fence_after();
        if (elect_one()) {
          if (!stl_first_done) STL_EXTRA(*stl_slot_smem, 0);  // first MMA of the CTA
          stl_first_done = true;
          const uint64_t b_desc = make_smem_desc_sw128(smem_u32(smem_x + stage * Cfg::X_BYTES));
#pragma unroll
          for (int h = 0; h < NW; ++h) {
            const uint64_t a_desc =
                make_smem_desc_sw128(smem_u32(smem_w + stage * Cfg::W_BYTES + h * Cfg::W_TILE_BYTES));
#pragma unroll
            for (int k = 0; k < GEMM_BK / 16; ++k) {
              // advance 16 bf16 = 32 B inside the 128-B swizzle atom: +2 in 16-byte units on the start address
              tc_mma_f16(d_tmem + h * BT, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                         (kb > kb0 || k > 0) ? 1u : 0u);
            }
          }
          tc_commit(&empty_bar[stage]);  // frees the smem slot when these MMAs retire
          if (kb == kb1 - 1) tc_commit(&tmem_full[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) {
          stage = 0;
          phase ^= 1;
        }
      }
      acc_bits ^= (1u << acc);
      acc = (acc + 1) % N_ACC;
    }
  } else if (warp < 6) {
    // ===================== epilogue (4 warps = 128 TMEM lanes) =====================
    const int sub = warp & 3;          // TMEM sub-partition this warp may read
    const int row = sub * 32 + lane_id();  // weight row within the tile
    const int ep_tid = (warp - 2) * 32 + lane_id();
    if constexpr (BT <= GEMM_NORM_MAX_T) {
      if (norm_helpers) {  // start-up burst of the fused-norm consumer: k-blocks 2..5
        asm volatile("bar.sync 2, 192;\n" ::: "memory");  // rstd is in shared memory
        float rs[NJ];
#pragma unroll
        for (int j = 0; j < NJ; ++j) rs[j] = rstd_smem[(lane_id() >> 3) + 4 * j];
        norm_stage(warp, rs);  // warps 2..5
      }
    }
    UnitIter it(sched, cta);
    int tile, kb0, kb1, slot;
    int acc = 0;
    uint32_t acc_bits = 0;  // per-buffer phase parity
    while (it.next(tile, kb0, kb1, slot)) {
      int n_tile, t_tile;
      tile_decode(sched, tile, n_tile, t_tile);
      const int t_base = t_tile * BT;
      const int t_valid = min(BT, T - t_base);
      const bool partial = (kb0 > 0) || (kb1 < KB);
      mbar_wait(&tmem_full[acc], (acc_bits >> acc) & 1);
      tc_fence_after();
      if (cl_push) asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");  // all peers are running
      if (ep_tid == 0) TL(5);  // accumulator ready (all MMAs of the unit retired)
      if (ep_tid == 0 && !norm_in) STL_EXTRA(*stl_slot_smem, 2);
      float* my_ws = ws + (size_t)(cta * 2 + slot) * Cfg::PART_FLOATS;
#pragma unroll 1
      for (int h = 0; h < NW; ++h) {  // the unit's NW accumulators = weight tiles n_tile * NW + h
        const int n = (n_tile * NW + h) * GEMM_BN + row;
        const uint32_t taddr = tmem_base + acc * ACC_COLS + h * BT + ((uint32_t)(sub * 32) << 16);
#pragma unroll 1
        for (int c0 = 0; c0 < BT; c0 += 16) {
          if (c0 >= t_valid) break;
          uint32_t r[16];
          tmem_ld_32x32b_x16(taddr + c0, r);
          tmem_ld_wait();
          if (!partial) {
            if (out_f32 == 2) {
              // fused SwiGLU: weight rows are interleaved (2j = gate_j, 2j+1 = up_j) so the pair sits in adjacent
              // lanes; even lanes produce act[t, n/2]
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float other = __shfl_xor_sync(0xffffffffu, __uint_as_float(r[j]), 1);
                if ((row & 1) == 0 && n < N && c0 + j < t_valid)
                  Y[(size_t)(t_base + c0 + j) * ldy + (n >> 1)] = swiglu_bf16(__uint_as_float(r[j]), other);
              }
            } else if (n < N) {
#pragma unroll
              for (int j = 0; j < 16; ++j)
                if (c0 + j < t_valid) {
                  const size_t o = (size_t)(t_base + c0 + j) * ldy + n;
                  if (out_f32) Yf[o] = __uint_as_float(r[j]);
                  else Y[o] = __float2bfloat16_rn(__uint_as_float(r[j]));
                }
            }
          } else {
            // split tile: fp32 partial to the global workspace, or (cluster mode) to this CTA's own shared memory --
            // the ring is idle by now: this CTA's only unit has retired all its MMAs
            if (cl_push) {
              // receive buffer of owner o (behind the ring, never aliased with in-flight stages -- the owner may still be
              // in its main loop): [source rank][local token index i = t / s][128 rows]; a warp's 32 lanes write one
              // 128-byte line of the owner's shared memory
              const int crank_e = (int)cluster_ctarank();
              const int tpo = (BT + cluster_split - 1) / cluster_split;
              const uint32_t rbase = smem_u32(wn_smem) + (uint32_t)(row * 4);
              int o = c0 % cluster_split, i_loc = c0 / cluster_split;
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                if (c0 + j < t_valid) {
                  const uint32_t ra = dsmem_addr(rbase + (uint32_t)(((crank_e * tpo + i_loc) * GEMM_BN) * 4), (uint32_t)o);
                  asm volatile("st.shared::cluster.f32 [%0], %1;\n" ::"r"(ra), "f"(__uint_as_float(r[j])) : "memory");
                }
                if (++o == cluster_split) {
                  o = 0;
                  ++i_loc;
                }
              }
            } else {
            float* dst = (cluster_split > 0 ? reinterpret_cast<float*>(smem) : my_ws) + h * (BT * GEMM_BN);
#pragma unroll
            for (int j = 0; j < 16; ++j)
              if (c0 + j < t_valid) dst[(size_t)(c0 + j) * GEMM_BN + row] = __uint_as_float(r[j]);
            }
          }
        }
      }
      // accumulators drained -> hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      __syncwarp();
      if (lane_id() == 0) mbar_arrive(&tmem_empty[acc]);
      acc_bits ^= (1u << acc);
      acc = (acc + 1) % N_ACC;

      if (ep_tid == 0) TL(6);  // TMEM drained, partial/direct stores issued
      if (partial && cluster_split > 0) {
        cl_tile = tile;  // reduced after the cluster barrier below
        cl_tvalid = t_valid;
      } else if (partial) {
        // stream-K fix-up: last arriver reduces all partials of this tile in CTA order (deterministic)
        // publish: CTA-wide barrier, then ONE acq_rel atomic (cumulative over the barrier) instead of membar.gl
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
        const long long p0 = (long long)(tile - sched.sk_tile0) * KB;  // position inside the stream-K region
        const int c_first = unit_owner(p0, total, ncta);
        const int c_last = unit_owner(p0 + KB - 1, total, ncta);
        if (ep_tid == 0) {
          int old;
          asm volatile("atom.acq_rel.gpu.global.add.s32 %0, [%1], 1;\n" : "=r"(old) : "l"(counters + tile) : "memory");
          *flag_smem = (old == (c_last - c_first)) ? 1 : 0;
        }
        asm volatile("bar.sync 1, 128;\n" ::: "memory");
        if (ep_tid == 0) TL(7);  // arrival atomic returned
        const int is_last = *flag_smem;
        if (is_last) {
          // Ordered (CTA-rank) reduction, one L2 round trip for the common case: thread -> 4 consecutive rows
          // (one 16-B load per contributor and token), tokens t == ep_tid/32 (mod 4); all FIX_T x FIX_C loads of a
          // batch are issued before the first add.  Sum order per element is contributor rank order -> deterministic.
          constexpr int FIX_C = 3, FIX_T = 8;
          const int r4 = (ep_tid & 31) * 4;   // first of this thread's 4 rows
          const int tq = ep_tid >> 5;         // token phase 0..3
#pragma unroll 1
          for (int h = 0; h < NW; ++h) {
          const int nt = n_tile * NW + h;     // weight tile (= head of the qkv projection)
          const int n4 = nt * GEMM_BN + r4;
          for (int tb = tq; tb < t_valid; tb += 4 * FIX_T) {
            float4 acc[FIX_T];
#pragma unroll
            for (int j = 0; j < FIX_T; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
            for (int c0g = c_first; c0g <= c_last; c0g += FIX_C) {
              float4 v[FIX_C][FIX_T];
#pragma unroll
              for (int ci = 0; ci < FIX_C; ++ci) {
                const int c = c0g + ci;
                const bool cv = c <= c_last;
                const long long cb = cv ? (total * c) / ncta : 0;
                const int cslot = ((int)(cb / KB) == tile - sched.sk_tile0) ? 0 : 1;
                const float* p = ws + (size_t)((cv ? c : c_first) * 2 + cslot) * Cfg::PART_FLOATS + h * (BT * GEMM_BN) + r4;
#pragma unroll
                for (int j = 0; j < FIX_T; ++j) {
                  const int t = tb + 4 * j;
                  v[ci][j] = (cv && t < t_valid) ? __ldcg(reinterpret_cast<const float4*>(p + (size_t)t * GEMM_BN))
                                                 : make_float4(0.f, 0.f, 0.f, 0.f);
                }
              }
#pragma unroll
              for (int ci = 0; ci < FIX_C; ++ci)
#pragma unroll
                for (int j = 0; j < FIX_T; ++j) {
                  acc[j].x += v[ci][j].x; acc[j].y += v[ci][j].y; acc[j].z += v[ci][j].z; acc[j].w += v[ci][j].w;
                }
            }
#pragma unroll
            for (int j = 0; j < FIX_T; ++j) {
              const int t = tb + 4 * j;
              if (t < t_valid) {
                const size_t o = (size_t)(t_base + t) * ldy + n4;
                const float av[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
                if (rope.positions != nullptr) {  // qkv projection: RoPE + KV scatter (warp-uniform t and tile)
                  qkv_rope_store(rope, nt, t_base + t, r4, av, Y + o);
                } else if (out_f32 == 2) {  // fused SwiGLU: rows (n4, n4+1) and (n4+2, n4+3) are (gate, up) pairs
                  __nv_bfloat16* yo = Y + (size_t)(t_base + t) * ldy + (n4 >> 1);
                  if (n4 + 1 < N) yo[0] = swiglu_bf16(av[0], av[1]);
                  if (n4 + 3 < N) yo[1] = swiglu_bf16(av[2], av[3]);
                } else if (n4 + 3 < N && (ldy & 3) == 0) {
                  if (out_f32 == 1) {
                    *reinterpret_cast<float4*>(Yf + o) = acc[j];
                  } else {
                    __nv_bfloat162 lo = __floats2bfloat162_rn(av[0], av[1]), hi = __floats2bfloat162_rn(av[2], av[3]);
                    uint2 pk;
                    pk.x = *reinterpret_cast<uint32_t*>(&lo);
                    pk.y = *reinterpret_cast<uint32_t*>(&hi);
                    *reinterpret_cast<uint2*>(Y + o) = pk;
                  }
                } else {
                  for (int e = 0; e < 4; ++e)
                    if (n4 + e < N) {
                      if (out_f32) Yf[o + e] = av[e];
                      else Y[o + e] = __float2bfloat16_rn(av[e]);
                    }
                }
              }
            }
          }
          }
          if (ep_tid == 0) counters[tile] = 0;
          if (ep_tid == 0) TL(8);  // ordered reduction done (last arriver only)
        }
        asm volatile("bar.sync 1, 128;\n" ::: "memory");  // flag_smem reuse safety
      }
    }
  } else if (norm_in) {
    // ===================== activation transform: fused residual-stream RMSNorm (warps 6..7) =====================
    // k-blocks alternate between the two warps; the four epilogue warps take k-blocks 2..5 of the start-up burst (see
    // norm_stage above).  rstd first: lane p fetches partial sum p of every row of this warp's half of the token tile
    // (all loads independent: one L2 round trip), a fixed shuffle tree adds them (deterministic).
    if constexpr (BT <= GEMM_NORM_MAX_T) {
      const int tw = warp - 6;                     // 0 / 1
      const int tid = threadIdx.x - GEMM_THREADS;  // 0..63
      const int lane = lane_id();
      // the norm weights are static: all of them into shared memory before the dependency wait
      for (int i = tid; i < K / 8; i += 64)
        reinterpret_cast<uint4*>(wn_smem)[i] = __ldg(reinterpret_cast<const uint4*>(norm.w_norm) + i);
      griddep_wait();  // the partial sums (and h, through the producer warp's TMA) come from the preceding GEMM
      if (tid == 0) STL_EXTRA(*stl_slot_smem, 2);  // (timeline builds: fused consumers report this instead of acc-ready)
      constexpr int RW = BT / 2;
      const int row0 = tw * RW, np = norm.n_parts;
      const bool ok0 = lane < np, ok1 = lane + 32 < np;
      float v[RW];
#pragma unroll
      for (int r = 0; r < RW; ++r) {
        const int row = row0 + r;
        const float* pp = norm.sumsq_in + (size_t)(row < T ? row : 0) * np;
        const float a0 = __ldcg(pp + (ok0 ? lane : 0)), a1 = __ldcg(pp + (ok1 ? lane + 32 : 0));
        v[r] = (ok0 ? a0 : 0.f) + (ok1 ? a1 : 0.f);
      }
      // level by level over ALL rows (RW independent shuffles in flight per level), then lane r finishes row r
#pragma unroll
      for (int sh = 16; sh > 0; sh >>= 1) {
#pragma unroll
        for (int r = 0; r < RW; ++r) v[r] += __shfl_xor_sync(0xffffffffu, v[r], sh);
      }
      float mine = 0.f;
#pragma unroll
      for (int r = 0; r < RW; ++r) mine = lane == r ? v[r] : mine;
      if (lane < RW) rstd_smem[row0 + lane] = row0 + lane < T ? rsqrtf(mine / (float)K + norm.eps) : 0.f;
      if (norm_helpers) asm volatile("bar.sync 2, 192;\n" ::: "memory");
      else asm volatile("bar.sync 2, 64;\n" ::: "memory");
      if (tid == 0) STL_EXTRA(*stl_slot_smem, 3);
      float rs[NJ];
#pragma unroll
      for (int j = 0; j < NJ; ++j) rs[j] = rstd_smem[(lane >> 3) + 4 * j];
      int i = tw;
      if (i < norm_n_kb) norm_stage(i, rs);
      i += norm_helpers ? 6 : 2;
      for (; i < norm_n_kb; i += 2) norm_stage(i, rs);
    }
  }

  if (cluster_split > 0) {
    // ---- on-chip split-K reduction across the cluster (all threads take part in the cluster barriers)
    // (push variant: the epilogue warps consumed the start-up phase before their remote stores, the others do it here)
    if (cl_push && !(warp >= 2 && warp < 6)) asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
    cluster_sync_all();  // every CTA's partial is in its (pull) / its owner's (push) shared memory
    if (warp >= 2 && warp < 6 && cl_tile >= 0) {
      const int ep_tid = (warp - 2) * 32 + lane_id();
      const int crank = (int)cluster_ctarank();
      const int r4 = (ep_tid & 31) * 4;   // this thread's 4 consecutive weight rows
      const int tq = ep_tid >> 5;         // token phase 0..3
      const uint32_t stg = smem_u32(smem);
      constexpr int FIX_T = 4;
#pragma unroll 1
      for (int h = 0; h < NW; ++h) {
      const int nt = cl_tile * NW + h;    // weight tile (= head of the qkv projection)
      const int n4 = nt * GEMM_BN + r4;
      const uint32_t hoff = (uint32_t)(h * BT * GEMM_BN * 4);
      // tokens of this CTA: t = crank + split * i; thread takes i = tq (mod 4)
      for (int i0 = tq; crank + cluster_split * i0 < cl_tvalid; i0 += 4 * FIX_T) {
        float4 acc[FIX_T];
#pragma unroll
        for (int j = 0; j < FIX_T; ++j) acc[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        uint2 resv[FIX_T];  // fused residual add: this thread's residual values, in flight during the DSMEM reads
        if (norm.sumsq_out != nullptr) {
#pragma unroll
          for (int j = 0; j < FIX_T; ++j) {
            const int t = crank + cluster_split * (i0 + 4 * j);
            resv[j] = t < cl_tvalid ? __ldcg(reinterpret_cast<const uint2*>(norm.residual + (size_t)t * ldy + n4))
                                    : make_uint2(0u, 0u);
          }
        }
        for (int c = 0; c < cluster_split; ++c) {  // contributor rank order = k order: deterministic
          const uint32_t peer = dsmem_addr(stg, (uint32_t)c) + hoff;
          const int tpo = (BT + cluster_split - 1) / cluster_split;
          float4 v[FIX_T];
#pragma unroll
          for (int j = 0; j < FIX_T; ++j) {
            const int t = crank + cluster_split * (i0 + 4 * j);
            if (cl_push)  // own receive buffer: [source rank c][local token index][128 rows]
              v[j] = t < cl_tvalid ? *reinterpret_cast<const float4*>(wn_smem + (size_t)(((c * tpo + i0 + 4 * j) * GEMM_BN + r4) * 4))
                                   : make_float4(0.f, 0.f, 0.f, 0.f);
            else
            v[j] = t < cl_tvalid ? ld_dsmem_f4(peer + (uint32_t)((t * GEMM_BN + r4) * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
          }
#pragma unroll
          for (int j = 0; j < FIX_T; ++j) {
            acc[j].x += v[j].x; acc[j].y += v[j].y; acc[j].z += v[j].z; acc[j].w += v[j].w;
          }
        }
#pragma unroll
        for (int j = 0; j < FIX_T; ++j) {
          const int t = crank + cluster_split * (i0 + 4 * j);
          if (t < cl_tvalid) {
            const size_t o = (size_t)t * ldy + n4;
            const float av[4] = {acc[j].x, acc[j].y, acc[j].z, acc[j].w};
            if (norm.sumsq_out != nullptr) {
              // h = bf16(bf16(y) + residual) (fused_add_rms_norm's add), back into the residual stream, + this tile's
              // share of sum(h^2) for the consumer's rstd; t and the tile are warp-uniform
              const float rr[4] = {__uint_as_float(resv[j].x << 16), __uint_as_float(resv[j].x & 0xffff0000u),
                                   __uint_as_float(resv[j].y << 16), __uint_as_float(resv[j].y & 0xffff0000u)};
              float hh[4], ssq = 0.f;
#pragma unroll
              for (int e = 0; e < 4; ++e) {
                hh[e] = bf16_round(bf16_round(av[e]) + rr[e]);
                ssq += hh[e] * hh[e];
              }
              __nv_bfloat162 lo = __floats2bfloat162_rn(hh[0], hh[1]), hi = __floats2bfloat162_rn(hh[2], hh[3]);
              uint2 pk;
              pk.x = *reinterpret_cast<uint32_t*>(&lo);
              pk.y = *reinterpret_cast<uint32_t*>(&hi);
              *reinterpret_cast<uint2*>(norm.residual + o) = pk;
#pragma unroll
              for (int sh = 16; sh > 0; sh >>= 1) ssq += __shfl_xor_sync(0xffffffffu, ssq, sh);
              if (lane_id() == 0) norm.sumsq_out[(size_t)t * (N / GEMM_BN) + nt] = ssq;
            } else if (rope.positions != nullptr) {
              qkv_rope_store(rope, nt, t, r4, av, Y + o);
            } else if (out_f32 == 2) {
              __nv_bfloat16* yo = Y + (size_t)t * ldy + (n4 >> 1);
              if (n4 + 1 < N) yo[0] = swiglu_bf16(av[0], av[1]);
              if (n4 + 3 < N) yo[1] = swiglu_bf16(av[2], av[3]);
            } else if (n4 + 3 < N && (ldy & 3) == 0) {
              if (out_f32 == 1) {
                *reinterpret_cast<float4*>(Yf + o) = acc[j];
              } else {
                __nv_bfloat162 lo = __floats2bfloat162_rn(av[0], av[1]), hi = __floats2bfloat162_rn(av[2], av[3]);
                uint2 pk;
                pk.x = *reinterpret_cast<uint32_t*>(&lo);
                pk.y = *reinterpret_cast<uint32_t*>(&hi);
                *reinterpret_cast<uint2*>(Y + o) = pk;
              }
            } else {
              for (int e = 0; e < 4; ++e)
                if (n4 + e < N) {
                  if (out_f32) Yf[o + e] = av[e];
                  else Y[o + e] = __float2bfloat16_rn(av[e]);
                }
            }
          }
        }
      }
      }
    }
    if (!cl_push) cluster_sync_all();  // nobody exits while a peer may still read its shared memory
  }
  tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) TL(9);  // all roles done
  STL_EXIT();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc<Cfg::TMEM_COLS>(tmem_base);
  }
}

#ifdef TGIS_GEMM_TIMELINE
int gemm_timeline_read(unsigned long long* out) {
  return cudaMemcpyFromSymbol(out, g_gemm_timeline, sizeof(unsigned long long) * 64) == cudaSuccess ? 0 : -1;
}
#else
int gemm_timeline_read(unsigned long long*) { return -2; }
#endif

// ----------------------------------------------------------------------------------------------------------------
typedef CUresult (*PFN_encodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_encodeTiled get_encode_fn() {
  static PFN_encodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !p) return nullptr;
    fn = reinterpret_cast<PFN_encodeTiled>(p);
  }
  return fn;
}

int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols) {
  PFN_encodeTiled fn = get_encode_fn();
  if (!fn) return -1;
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld_elems * 2};
  cuuint32_t box[2] = {box_cols, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : -(int)r - 1000;
}


int gemm_pick_bt(int T) {
  if (T <= 16) return 16;
  if (T <= 32) return 32;
  if (T <= 64) return 64;
  if (T <= 128) return 128;
  return 256;
}

// Weight tiles per unit (GemmCfg).  1 everywhere by default.  TGIS_GEMM_NW=2 makes decode-shaped launches share one
// activation tile between two weight tiles (half the activation re-reads per weight byte): built, parity-green, and
// MEASURED SLOWER (profiles/r02_gemm_nw_ab.log: gate_up 46 -> 55 us at T = 32, down 31 -> 50 us at T = 64, decode step
// 3.93 -> 5.05 ms at batch 32): halving the tile count doubles the partial size of every split-K reduction and leaves 5
// ring stages instead of 8, while the activation re-reads it removes were not what bounds these launches.
int gemm_nw(int T) {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TGIS_GEMM_NW");
    v = (e && e[0] == '2') ? 2 : 1;
  }
  return T <= 256 ? v : 1;
}
static int gemm_bn(int T) { return GEMM_BN * gemm_nw(T); }

static bool gemm_cluster_enabled() {  // TGIS_GEMM_CLUSTER=0: always reduce split tiles through global memory
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TGIS_GEMM_CLUSTER");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

template <int BT, int NW>
__global__ void gemm_bf16_tcgen05_kernel(const __grid_constant__ CUtensorMap, const __grid_constant__ CUtensorMap, void*, int,
                                         int, int, int, float*, int*, int, int, const __grid_constant__ CUtensorMap,
                                         GemmNext, int, GemmRope, GemmNorm);

// How many clusters of `s` GEMM CTAs (one CTA per SM) the device can hold at once: a GPC hosts floor(its SMs / s) of them
// (148 SMs: 74 of 2, 45 of 3, 33 of 4, ...).  Queried once per size; without a device (host-only plan tests) the
// optimistic num_sms / s.
static int gemm_max_clusters(int s, int num_sms) {
  static int cache[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  if (s < 2 || s > 8) return num_sms / (s > 0 ? s : 1);
  if (cache[s] == 0) {
    using Cfg = GemmCfg<32, 1>;
    cudaLaunchConfig_t qc{};
    qc.gridDim = dim3(s * (num_sms / s));
    qc.blockDim = dim3(GEMM_THREADS);
    qc.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = s;
    qa[0].val.clusterDim.y = 1;
    qa[0].val.clusterDim.z = 1;
    qc.attrs = qa;
    qc.numAttrs = 1;
    int n = 0;
    if (cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<32, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_MAX) ==
            cudaSuccess &&
        cudaOccupancyMaxActiveClusters(&n, gemm_bf16_tcgen05_kernel<32, 1>, &qc) == cudaSuccess && n > 0) {
      cache[s] = n;
    } else {
      cudaGetLastError();
      cache[s] = num_sms / s;
    }
  }
  return cache[s];
}
static bool gemm_fit_clusters() {  // TGIS_GEMM_FIT_CLUSTERS=0: round-1 policy (largest split, cluster mode only if it fits)
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("TGIS_GEMM_FIT_CLUSTERS");
    v = (e && e[0] == '0') ? 0 : 1;
  }
  return v == 1;
}

size_t gemm_workspace_bytes(int num_sms) { return (size_t)num_sms * 2 * 2 * 256 * GEMM_BN * sizeof(float); }

// Grid size of a launch (pure function of the shape: the NEXT kernel's grid must be known one launch ahead)
int gemm_grid_size(int T, int N, int K, int num_sms) {
  const int BT = gemm_pick_bt(T), BN = gemm_bn(T);
  const int n_tiles = (N + BN - 1) / BN, t_tiles = (T + BT - 1) / BT, KB = (K + GEMM_BK - 1) / GEMM_BK;
  long long total = (long long)n_tiles * t_tiles * KB;
  // do not cut finer than 4 k-blocks per CTA: tiny problems use fewer CTAs
  long long max_ctas = (total + 3) / 4;
  int grid = (int)(max_ctas < num_sms ? (max_ctas < 1 ? 1 : max_ctas) : num_sms);
  // few tiles (decode-shaped qkv / o / down at decode): split every tile by the same integer factor so each CTA owns exactly
  // one unit inside one tile (one fix-up round instead of two; costs <= 1/(split+1) of the SMs)
  const long long tiles = (long long)n_tiles * t_tiles;
  if (tiles <= num_sms / 2) {
    int split = (int)(num_sms / tiles);
    if (split > KB / 4) split = KB / 4 > 0 ? KB / 4 : 1;
    // More than 8 contributors per tile would have to meet in a global-memory fix-up in which ONE last-arriving CTA sums
    // all partials (70B TP=8 qkv at T = 256: 1.8 MB per tile through 128 threads).  Cap the split at the largest
    // cluster: the reduction then runs in distributed shared memory, spread over the cluster's CTAs by token.
    if (split > 8) split = 8;
    // large token tiles: a partial is >= 64 KB per weight tile; splitting a short k-range (o-proj shards) costs more in
    // partial traffic than the extra SMs bring -- keep at least 16 k-blocks per CTA
    while (BT > 64 && split > 1 && KB / split < 16) --split;
    // Prefer a split whose clusters are all co-resident (qkv of the 8B model: 48 tiles x 3 CTAs -- only 45 clusters of 3
    // fit, so round 1 reduced it through global memory; 48 clusters of 2 fit): the on-chip reduction spreads the fused
    // RoPE + KV-scatter epilogue over the cluster's CTAs, which is what lets it replace rope_kvwrite_kernel above 32 tokens
    if (gemm_fit_clusters() && gemm_cluster_enabled() && t_tiles == 1)
      while (split > 2 && tiles > gemm_max_clusters(split, num_sms)) --split;
    grid = (int)(tiles * split);
  }
  if (const char* e = getenv("TGIS_GEMM_MAX_CTAS")) {  // experiment knob (scripts/gemm_cta_sweep.py)
    const int cap = atoi(e);
    if (cap > 0 && grid > cap) grid = cap;
  }
  return grid;
}

// Even-split launches (every tile cut into `split` k-ranges owned by `split` consecutive CTAs): the factor, or 0.
int gemm_even_split(int T, int N, int K, int num_sms) {
  const int BT = gemm_pick_bt(T), BN = gemm_bn(T);
  const int n_tiles = (N + BN - 1) / BN, t_tiles = (T + BT - 1) / BT, KB = (K + GEMM_BK - 1) / GEMM_BK;
  const long long tiles = (long long)n_tiles * t_tiles;
  if (t_tiles != 1 || tiles > num_sms / 2) return 0;
  const int grid = gemm_grid_size(T, N, K, num_sms);
  if (grid % tiles != 0) return 0;
  const int split = (int)(grid / tiles);
  // grid == tiles * split: CTA c's k-range [total*c/grid, total*(c+1)/grid) starts a tile exactly at every multiple of
  // split, so each CTA owns ONE unit inside ONE tile even when split does not divide KB (qkv: 64 k-blocks over 3 CTAs)
  return (split >= 2 && KB >= split) ? split : 0;
}

GemmNext gemm_next_desc(int T_next, int N_next, int K_next, int num_sms, int kb_prefetch) {
  GemmNext n{};
  const int BT = gemm_pick_bt(T_next);
  if ((T_next + BT - 1) / BT != 1) return n;  // only decode-shaped successors (one token tile) are prefetched
  n.n_tiles = (N_next + gemm_bn(T_next) - 1) / gemm_bn(T_next);
  n.KB = (K_next + GEMM_BK - 1) / GEMM_BK;
  n.grid = gemm_grid_size(T_next, N_next, K_next, num_sms);
  n.kb_prefetch = kb_prefetch;
  return n;
}

// cluster mode: on-chip split-K reduction when every tile is split evenly over <= 8 consecutive CTAs, the fp32 partial
// fits the drained ring, and all clusters can be co-resident (checked once per (BT, NW, split, grid)).  Returns the
// split or 0.
template <int BT, int NW>
static int cluster_split_bt(int T, int N, int K, int num_sms) {
  using Cfg = GemmCfg<BT, NW>;
  static int cluster_ok[9][160];  // [split][grid / split]: 0 unknown, 1 yes, -1 no
  int split = gemm_cluster_enabled() ? gemm_even_split(T, N, K, num_sms) : 0;
  if (split > 8 || (size_t)Cfg::PART_FLOATS * sizeof(float) > (size_t)Cfg::STAGES * Cfg::STAGE_BYTES) split = 0;
  if (split == 0) return 0;
  const int grid = gemm_grid_size(T, N, K, num_sms);
  const int nc = grid / split;
  if (nc >= 160) return 0;
  if (cluster_ok[split][nc] == 0) {
    static bool attr_set = false;
    if (!attr_set) {
      cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BT, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_MAX);
      attr_set = true;
    }
    cudaLaunchConfig_t qc{};
    qc.gridDim = dim3(split * (num_sms / split));
    qc.blockDim = dim3(GEMM_THREADS);
    qc.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cudaLaunchAttribute qa[1];
    qa[0].id = cudaLaunchAttributeClusterDimension;
    qa[0].val.clusterDim.x = split;
    qa[0].val.clusterDim.y = 1;
    qa[0].val.clusterDim.z = 1;
    qc.attrs = qa;
    qc.numAttrs = 1;
    int n_clusters = 0;
    const cudaError_t qe = cudaOccupancyMaxActiveClusters(&n_clusters, gemm_bf16_tcgen05_kernel<BT, NW>, &qc);
    // All clusters of the launch must be co-resident: a GPC only hosts floor(SMs / split) of them, and one cluster left
    // over for a second wave doubles the launch time (measured: qkv with 48 clusters of 3 when 47 fit).
    cluster_ok[split][nc] = (qe == cudaSuccess && n_clusters >= nc) ? 1 : -1;
    if (getenv("TGIS_GEMM_DEBUG"))
      fprintf(stderr, "[gemm] BT=%d NW=%d split=%d: %d clusters fit, grid %d -> cluster mode %s\n", BT, NW, split,
              n_clusters, grid, cluster_ok[split][nc] == 1 ? "on" : "off");
    if (qe != cudaSuccess) cudaGetLastError();
  }
  return cluster_ok[split][nc] == 1 ? split : 0;
}

template <int BT>
static int cluster_split_nw(int T, int N, int K, int num_sms) {
  return gemm_nw(T) == 2 ? cluster_split_bt<BT, 2>(T, N, K, num_sms) : cluster_split_bt<BT, 1>(T, N, K, num_sms);
}

// The split factor gemm_bf16_launch will run this shape with in cluster mode (0: global-memory fix-up) -- callers that
// want the fused RoPE epilogue ask first.
int gemm_cluster_split(int T, int N, int K, int num_sms) {
  switch (gemm_pick_bt(T)) {
    case 16: return cluster_split_nw<16>(T, N, K, num_sms);
    case 32: return cluster_split_nw<32>(T, N, K, num_sms);
    case 64: return cluster_split_nw<64>(T, N, K, num_sms);
    case 128: return cluster_split_nw<128>(T, N, K, num_sms);
    default: return cluster_split_nw<256>(T, N, K, num_sms);
  }
}

template <int BT, int NW>
static cudaError_t launch_bt(const CUtensorMap& wmap, const CUtensorMap& xmap, void* Y, int ldy, int T, int N, int K,
                             float* ws, int* counters, int num_sms, int out_f32, cudaStream_t stream,
                             const CUtensorMap& next_wmap, const GemmNext& nxt, const GemmRope& rope, const GemmNorm& norm) {
  using Cfg = GemmCfg<BT, NW>;
  static bool attr_set = false;
  if (!attr_set) {
    cudaError_t e = cudaFuncSetAttribute(gemm_bf16_tcgen05_kernel<BT, NW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         Cfg::SMEM_MAX);
    if (e != cudaSuccess) return e;
    attr_set = true;
  }
  const int t_tiles = (T + BT - 1) / BT;
  const int grid = gemm_grid_size(T, N, K, num_sms);
  // bit 0: weights are read once (evict-first); bit 1: stream-K share before the whole tiles (GemmSched::sk_first;
  // TGIS_GEMM_SK_FIRST=0: pure stream-K as in round 1).  The fused-norm consumer's transform warps assume the pure cut.
  // Token tiles above 64 rows keep the pure cut: there the whole tile's direct epilogue (128+ tokens through the TMEM
  // loads) is a longer tail than the fix-up it replaces (measured: batch 128 +1.2 %, batch 32 / 64 -2.3 / -1.2 %).
  static int sk_first_env = -1;
  if (sk_first_env < 0) {
    const char* e = getenv("TGIS_GEMM_SK_FIRST");
    sk_first_env = (e && e[0] == '0') ? 0 : 1;
  }
  int stream_weights = (t_tiles == 1) ? (1 | ((sk_first_env && norm.h == nullptr && BT <= 64) ? 2 : 0)) : 0;
  const int split = cluster_split_bt<BT, NW>(T, N, K, num_sms);
  // bit 2: push variant of the cluster reduction (receive buffer of (BT + 8) x 128 floats behind the ring; token tiles of
  // at most 32 rows, where it fits next to the full ring).  TGIS_GEMM_CLUSTER_PUSH=1 enables it (default: pull variant).
  static int push_env = -1;
  if (push_env < 0) {
    const char* e = getenv("TGIS_GEMM_CLUSTER_PUSH");
    push_env = (e && e[0] == '1') ? 1 : 0;
  }
  const bool push = push_env && split > 0 && BT <= 32 && NW == 1 && norm.h == nullptr;
  if (push) stream_weights |= 4;
  // the fused RoPE epilogue lives in the two split-tile reductions (cluster and global fix-up): every tile must be split
  if (rope.positions != nullptr && (gemm_even_split(T, N, K, num_sms) < 2 || out_f32 != 0 ||
                                    N != (rope.n_q + 2 * rope.n_kv) * HEAD_DIM))
    return cudaErrorInvalidValue;
  // fused residual add / RMSNorm (GemmNorm): the producer side lives in the cluster reduction only, the consumer side in
  // decode-shaped launches with a token tile of at most 64 rows
  if (norm.sumsq_out != nullptr && (split == 0 || out_f32 != 0 || rope.positions != nullptr || N % GEMM_BN != 0 ||
                                    norm.residual == nullptr || ldy != N))
    return cudaErrorInvalidValue;
  const size_t smem_bytes = (size_t)Cfg::SMEM_BYTES + (norm.h != nullptr ? (size_t)K * 2 : 0) +
                            (push ? (size_t)(BT + 8) * GEMM_BN * 4 : 0);
  if (norm.h != nullptr && (t_tiles != 1 || BT > GEMM_NORM_MAX_T || K % GEMM_BK != 0 || NW != 1 || norm.n_parts <= 0 || norm.n_parts > GEMM_NORM_MAX_PARTS ||
                            smem_bytes > (size_t)Cfg::SMEM_MAX))
    return cudaErrorInvalidValue;
  const int threads = norm.h != nullptr ? GEMM_THREADS_NORM : GEMM_THREADS;
  if (split > 0)
    return launch_k_cluster(gemm_bf16_tcgen05_kernel<BT, NW>, dim3(grid), dim3(threads), smem_bytes, stream, split,
                            wmap, xmap, Y, ldy, T, N, K, ws, counters, stream_weights, out_f32, next_wmap, nxt, split, rope,
                            norm);
  return launch_k(gemm_bf16_tcgen05_kernel<BT, NW>, dim3(grid), dim3(threads), smem_bytes, stream, wmap, xmap, Y,
                  ldy, T, N, K, ws, counters, stream_weights, out_f32, next_wmap, nxt, 0, rope, norm);
}

// xmap must have been built with box_rows == gemm_pick_bt(T)
cudaError_t gemm_bf16_launch(const CUtensorMap& wmap, const CUtensorMap& xmap, void* Y, int ldy, int T, int N, int K,
                             float* ws, int* counters, int num_sms, cudaStream_t stream, int out_f32,
                             const CUtensorMap* next_wmap, const GemmNext* next, const GemmRope* rope, const GemmNorm* norm) {
  GemmRope rp{};
  if (rope) rp = *rope;
  GemmNorm nr{};
  if (norm) nr = *norm;
  GemmNext nx{};
  if (next && next_wmap) nx = *next;
  const CUtensorMap& nm = (next && next_wmap) ? *next_wmap : wmap;
  const int nw = gemm_nw(T);
#define TGIS_GEMM_CASE(B)                                                                                              \
  return nw == 2 ? launch_bt<B, 2>(wmap, xmap, Y, ldy, T, N, K, ws, counters, num_sms, out_f32, stream, nm, nx, rp, nr) \
                 : launch_bt<B, 1>(wmap, xmap, Y, ldy, T, N, K, ws, counters, num_sms, out_f32, stream, nm, nx, rp, nr)
  switch (gemm_pick_bt(T)) {
    case 16: TGIS_GEMM_CASE(16);
    case 32: TGIS_GEMM_CASE(32);
    case 64: TGIS_GEMM_CASE(64);
    case 128: TGIS_GEMM_CASE(128);
    default: TGIS_GEMM_CASE(256);
  }
#undef TGIS_GEMM_CASE
}

}  // namespace tgis
