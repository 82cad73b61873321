// Internal launcher declarations for the sm_100a kernels of the TGIS decode/prefill hot path.
// (Internal C++ header; the drop-in C ABI is include/tgis_engine.h.)
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

namespace tgis {

// ---- gemm_tcgen05.cu --------------------------------------------------------------------------------------------
int make_tmap_bf16_2d(CUtensorMap* out, const void* base, uint64_t rows, uint64_t cols, uint64_t ld_elems,
                      uint32_t box_rows, uint32_t box_cols);
int gemm_pick_bt(int T);
int gemm_nw(int T);  // weight tiles (128 rows each) per unit sharing one activation tile: 2 for decode-shaped launches
size_t gemm_workspace_bytes(int num_sms);
int gemm_timeline_read(unsigned long long* out64);  // debug builds (-DTGIS_GEMM_TIMELINE) only
// -DTGIS_STEP_TIMELINE builds: point every translation unit's kernels at the timeline buffer (ptx.cuh); else -2
int gemm_set_step_timeline(unsigned long long* buf);
int elementwise_set_step_timeline(unsigned long long* buf);
int attention_set_step_timeline(unsigned long long* buf);
int sampler_set_step_timeline(unsigned long long* buf);
// Decode-shaped successor of a GEMM launch: lets the running kernel pull the first `kb_prefetch` weight boxes of every
// CTA of the NEXT GEMM into L2 while its own tail drains (kb_prefetch == 0: off).
struct GemmNext {
  int n_tiles, KB, grid, kb_prefetch;
};
int gemm_grid_size(int T, int N, int K, int num_sms);
int gemm_even_split(int T, int N, int K, int num_sms);  // every tile cut over this many consecutive CTAs, or 0
GemmNext gemm_next_desc(int T_next, int N_next, int K_next, int num_sms, int kb_prefetch);
// Y: bf16 [T, ldy] (out_f32 = 0), fp32 [T, ldy] (out_f32 = 1, lm_head logits), or out_f32 = 2: fused SwiGLU — W rows are
// interleaved (gate_j, up_j) pairs and Y is bf16 [T, ldy >= N/2] = bf16(bf16(silu(gate)) * up)
// Optional fused epilogue of the qkv projection (cluster mode only, gemm_cluster_split() > 0): RoPE on the q / k heads
// and scatter of k / v into the paged cache, replacing rope_kvwrite_kernel.  N must be (n_q + 2 n_kv) * 128.
struct GemmRope {
  const int32_t* positions;     // [T]; nullptr = no fusion
  const int32_t* slot_mapping;  // [T], < 0: do not cache
  const __nv_bfloat16* cos_sin; // [max_pos][128]: cos (64) | sin (64)
  __nv_bfloat16* k_cache;       // this layer's caches
  __nv_bfloat16* v_cache;
  int32_t n_q, n_kv;
};
// Optional fusion of the layer stack's  residual add + RMSNorm  (vllm layernorm.py fused_add_rms_norm; same rounding
// points as add_rmsnorm_launch) into the two GEMMs either side of it, decode-shaped launches only (one token tile,
// T <= 64):
//  * producer (o / down projection, cluster mode required): the split-tile reduction does not write Y; it computes
//    h = bf16(bf16(acc) + residual), stores h back to `residual` [T, N] and the tile's sum of h^2 to
//    sumsq_out[t * (N / 128) + tile]  (N % 128 == 0);
//  * consumer (qkv / gate_up projection): the activation operand is not loaded by TMA; two extra warps build
//    bf16(bf16(h * rstd) * w_norm) straight into the swizzled shared-memory stage, rstd[t] from the n_parts partial sums
//    added in tile order (deterministic).  K % 64 == 0.
struct GemmNorm {
  __nv_bfloat16* residual;       // producer: in/out [T, N]; nullptr = off
  float* sumsq_out;              // producer: [T][N / 128]
  const __nv_bfloat16* h;        // consumer: [T, K]; nullptr = off
  const float* sumsq_in;         // consumer: [T][n_parts]
  const __nv_bfloat16* w_norm;   // consumer: [K]
  int32_t n_parts;
  float eps;
};
constexpr int GEMM_NORM_MAX_T = 64;
constexpr int GEMM_NORM_MAX_PARTS = 64;  // hidden <= 8192
int gemm_cluster_split(int T, int N, int K, int num_sms);
cudaError_t gemm_bf16_launch(const CUtensorMap& wmap, const CUtensorMap& xmap, void* Y, int ldy, int T, int N, int K,
                             float* ws, int* counters, int num_sms, cudaStream_t stream, int out_f32 = 0,
                             const CUtensorMap* next_wmap = nullptr, const GemmNext* next = nullptr,
                             const GemmRope* rope = nullptr, const GemmNorm* norm = nullptr);

// ---- gemm_ref.cu (debug cross-check only; never on the product path) ---------------------------------------------
cudaError_t gemm_bf16_ref_launch(const __nv_bfloat16* X, int ldx, const __nv_bfloat16* W, void* Y, int ldy, int T,
                                 int N, int K, cudaStream_t stream, int out_f32 = 0);

// ---- elementwise.cu ----------------------------------------------------------------------------------------------
cudaError_t stl_marker_launch(int kid, cudaStream_t stream);  // -DTGIS_STEP_TIMELINE builds only (else a no-op)
cudaError_t embed_gather_launch(const int32_t* token_ids, const __nv_bfloat16* table, __nv_bfloat16* out, int T,
                                int hidden, int vocab, cudaStream_t stream);
// out = rmsnorm(x) * w                                   (first layer: residual := x is done by the caller)
cudaError_t rmsnorm_launch(const __nv_bfloat16* x, const __nv_bfloat16* w, __nv_bfloat16* out, int T, int hidden,
                           float eps, cudaStream_t stream);
// residual = bf16(x + residual); out = rmsnorm(residual) * w     (vllm: layernorm.py fused_add_rms_norm)
cudaError_t add_rmsnorm_launch(const __nv_bfloat16* x, __nv_bfloat16* residual, const __nv_bfloat16* w,
                               __nv_bfloat16* out, int T, int hidden, float eps, cudaStream_t stream);
// Tensor parallelism, decode-shaped steps: one-shot all-reduce of the row-parallel GEMM partials over NVLink peer memory
// fused with the residual add + RMSNorm (elementwise.cu).  own: this rank's partial [T, hidden]; recv[q]: rank q's receive
// area for this exchange parity, mapped into this process (recv[own rank] = local): [8 source ranks][AR_MAX_ROWS rows]
// [hidden / 8 vectors][2 lines of 16 bytes = {data, epoch, data, epoch}].
constexpr int AR_MAX_ROWS = 256;
inline size_t ar_recv_bytes(int hidden) { return (size_t)8 * AR_MAX_ROWS * (hidden / 8) * 2 * 16; }
struct ArPeers {
  const __nv_bfloat16* own;
  uint4* recv[8];
};
// The exchange's epoch is *epoch_base + epoch_idx (epoch_base: device memory written by the step's metadata copy).
cudaError_t ar_add_rmsnorm_launch(const ArPeers& peers, int tp, int rank, const uint32_t* epoch_base, uint32_t epoch_idx,
                                  __nv_bfloat16* residual, const __nv_bfloat16* w, __nv_bfloat16* out, int T, int hidden,
                                  float eps, cudaStream_t stream);
// Two-shot variant (reduce-scatter by row owner t mod tp, all-gather of the bf16 reduced rows; elementwise.cu) for
// exchanges whose one-shot egress would be large.  area[q]: rank q's two-shot area for this exchange parity, mapped here.
inline size_t ar2_area_bytes(int hidden) { return (size_t)9 * AR_MAX_ROWS * hidden * 2 + (size_t)9 * AR_MAX_ROWS * 4; }
struct Ar2Peers {
  const __nv_bfloat16* own;
  uint8_t* area[8];
};
cudaError_t ar2_add_rmsnorm_launch(const Ar2Peers& peers, int tp, int rank, const uint32_t* epoch_base, uint32_t epoch_idx,
                                   __nv_bfloat16* residual, const __nv_bfloat16* w, __nv_bfloat16* out, int T, int hidden,
                                   float eps, cudaStream_t stream);
// Vocab-parallel lm_head without NCCL: ranks > 0 push their shard into rank 0's logits buffer (peer memory) and raise a
// flag there; rank 0 waits for the flags (elementwise.cu).  Sizes in bytes (multiples of 16).
cudaError_t logits_push_launch(const void* shard, void* dst_peer0, int R, int Vl_bytes, int V_bytes, int col_bytes,
                               uint32_t* remote_flag, int* local_counter, const uint32_t* epoch_base, uint32_t epoch_idx,
                               cudaStream_t stream);
cudaError_t logits_wait_launch(const uint32_t* flags, int tp, const uint32_t* epoch_base, uint32_t epoch_idx,
                               cudaStream_t stream);
// act[t, i] = bf16(silu(gate_up[t, i])) * gate_up[t, F + i]
cudaError_t silu_mul_launch(const __nv_bfloat16* gate_up, __nv_bfloat16* act, int T, int ffn, cudaStream_t stream);
// ---- multi-adapter LoRA (lora.cu): y[t, col0 + n] += bf16(B_s[n, :] . bf16(A_s x[t])), s = tok_slot[t] (0 = no adapter)
struct LoraModule {
  const __nv_bfloat16* A;  // [slots][Rm][K]  (slot s >= 1 lives at index s - 1), zero padded to Rm rows
  const __nv_bfloat16* B;  // [slots][N][Rm]  (alpha / r already folded in), zero padded to Rm columns
  int K, N, Rm;            // Rm: rank capacity of the module (multiple of 8, <= 128)
  int col0;                // first output column of the module in y (multiple of 8)
  int v_off;               // offset of the module's rank vector inside a token's row of the shrink buffer
};
constexpr int LORA_GROUP_MAX = 3;  // modules that read the same x (q, k, v)
struct LoraGroup {
  LoraModule mod[LORA_GROUP_MAX];
  int n_mods;
  int v_ld;                // floats per token in the shrink buffer
};
cudaError_t lora_shrink_launch(const __nv_bfloat16* x, int ldx, const int32_t* tok_slot, const LoraGroup& g, float* v, int T,
                               cudaStream_t stream);
cudaError_t lora_expand_launch(const float* v, const int32_t* tok_slot, const LoraGroup& g, __nv_bfloat16* y, int ldy, int T,
                               cudaStream_t stream);
// act[t, j] = bf16(silu(gate_up[t, 2j])) * gate_up[t, 2j + 1]  (the interleaved gate_up layout the fused GEMM epilogue reads)
cudaError_t silu_mul_interleaved_launch(const __nv_bfloat16* gate_up, __nv_bfloat16* act, int T, int ffn,
                                        cudaStream_t stream);
// ---- opt.cu (OPT layer stack: learned positions, biased projections, ReLU, LayerNorm) ------------------------------
// out[t] = bf16(tok_table[tok[t]] + pos_table[pos[t] + offset])
cudaError_t opt_embed_launch(const int32_t* token_ids, const int32_t* positions, const __nv_bfloat16* tok_table,
                             const __nv_bfloat16* pos_table, __nv_bfloat16* out, int T, int hidden, int vocab,
                             int n_pos_rows, int offset, cudaStream_t stream);
// acc != nullptr: residual = bf16(residual + bf16(acc + acc_bias)) first (acc: fp32 GEMM accumulators [T, hidden]);
// out = LayerNorm(residual) * w + b
cudaError_t opt_layernorm_launch(const float* acc, const __nv_bfloat16* acc_bias, __nv_bfloat16* residual,
                                 const __nv_bfloat16* w, const __nv_bfloat16* b, __nv_bfloat16* out, int T, int hidden,
                                 float eps, cudaStream_t stream);
// out[t, n] = bf16(act(acc[t, n] + bias[n])), act = ReLU or identity
cudaError_t opt_bias_act_launch(const float* acc, int ld_acc, const __nv_bfloat16* bias, __nv_bfloat16* out, int ld_out,
                                int T, int N, int relu, int num_sms, cudaStream_t stream);
// qkv epilogue: q_out[t, head, :] = bf16(acc + bias) for the q heads; k / v heads go straight into the paged cache
// (slot_mapping[t] < 0: not cached).  acc fp32 [T, (n_q + 2 n_kv) * 128], q_out bf16 with the same row stride.
cudaError_t opt_qkv_bias_kvwrite_launch(const float* acc, const __nv_bfloat16* bias, __nv_bfloat16* q_out,
                                        const int32_t* slot_mapping, __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, int T,
                                        int n_q, int n_kv, cudaStream_t stream);
// gather rows: out[r, :] = x[rows[r], :]
cudaError_t gather_rows_launch(const __nv_bfloat16* x, const int32_t* rows, __nv_bfloat16* out, int R, int hidden,
                               cudaStream_t stream);

// Paged KV cache layout (per layer), chosen for the decode kernel's TMA bulk loads:
//   K: [num_blocks][n_kv][HEAD_DIM/8][BLOCK][8] bf16  (8-dim chunk major: lanes=tokens read conflict-free 16 B)
//   V: [num_blocks][n_kv][BLOCK][16 chunks ^ (tok&7)][8] bf16  (row-major with the 16-B chunk index XOR-swizzled)
// one (block, kv_head) K or V tile is a contiguous BLOCK*HEAD_DIM*2 bytes.
constexpr int KV_BLOCK = 32;
constexpr int HEAD_DIM = 128;

// neox RoPE on q,k in-place inside qkv[T, (nq+2nkv)*128] and scatter of k,v into the paged cache.
// cos_sin: [max_pos][128] bf16 = cos(64) | sin(64)   (vllm: rotary_embedding/base.py cache.to(dtype))
cudaError_t rope_kvwrite_launch(__nv_bfloat16* qkv, const int32_t* positions, const int32_t* slot_mapping,
                                const __nv_bfloat16* cos_sin, __nv_bfloat16* k_cache, __nv_bfloat16* v_cache, int T,
                                int n_q, int n_kv, cudaStream_t stream);

// ---- attention.cu ------------------------------------------------------------------------------------------------
struct AttnSeq {        // one entry per sequence scheduled this step (device array)
  int32_t q_start;      // first row of this sequence's queries in the flat token batch
  int32_t q_len;        // number of query tokens this step (1 = decode)
  int32_t kv_len;       // context length INCLUDING this step's tokens
  int32_t block_row;    // row in the block table
};
// Decode (q_len == 1): split-KV over chunks of DECODE_SPLIT tokens; the host lists the (sequence, split) entries of
// the step, the kernel crosses them with the kv heads.  items[0].q_row = number of entries, entries follow from [1].
constexpr int DECODE_SPLIT = 128;
struct DecItem {       // 32 bytes
  int32_t q_row;       // row of the sequence's token in qkv / out            ([0]: number of entries)
  int32_t kv_len;      // of the whole sequence
  int32_t seq_split;   // seq | split << 16; seq = index among this step's decode sequences (partial buffers)
  int32_t reserved;
  int32_t blocks[DECODE_SPLIT / KV_BLOCK];  // physical KV blocks of the split (resolved on the host: no table walk)
};
static_assert(DECODE_SPLIT / KV_BLOCK == 4 && sizeof(DecItem) == 32, "DecItem is read as two int4");
// Entries are listed longest first (all full splits, then the partial tails by decreasing length): the kernel deals
// them to its warps round-robin, so the order is the load balance.  Ties keep (sequence, split) order: deterministic.
inline int decode_items_build(DecItem* items, const AttnSeq* seqs, const int32_t* seq_ids, int n_seqs,
                              const int32_t* block_table, int bt_stride) {
  constexpr int BPS = DECODE_SPLIT / KV_BLOCK;
  int n = 0;
  auto emit = [&](int i, const AttnSeq& sq, int sp) {
    const int n_blocks = (sq.kv_len + KV_BLOCK - 1) / KV_BLOCK;
    const int32_t* row = block_table + (size_t)sq.block_row * bt_stride;
    DecItem& it = items[++n];
    it = DecItem{sq.q_start, sq.kv_len, i | (sp << 16), 0, {0, 0, 0, 0}};
    for (int j = 0; j < BPS; ++j) it.blocks[j] = sp * BPS + j < n_blocks ? row[sp * BPS + j] : 0;
  };
  for (int i = 0; i < n_seqs; ++i) {  // full splits
    const AttnSeq& sq = seqs[seq_ids ? seq_ids[i] : i];
    for (int sp = 0; sp < sq.kv_len / DECODE_SPLIT; ++sp) emit(i, sq, sp);
  }
  for (int nb = BPS; nb >= 1; --nb) {  // tails: 4, 3, 2, 1 blocks (a tail of exactly DECODE_SPLIT tokens was a full split)
    for (int i = 0; i < n_seqs; ++i) {
      const AttnSeq& sq = seqs[seq_ids ? seq_ids[i] : i];
      const int rem = sq.kv_len % DECODE_SPLIT;
      if (rem > 0 && (rem + KV_BLOCK - 1) / KV_BLOCK == nb) emit(i, sq, sq.kv_len / DECODE_SPLIT);
    }
  }
  items[0] = DecItem{n, 0, 0, 0, {0, 0, 0, 0}};
  return n;
}
// Streaming kernel + (max_splits > 1) the split-merge kernel.  seqs / seq_ids / n_seqs: this step's decode sequences in
// the order decode_items_build numbered them.
cudaError_t attn_decode_launch(const __nv_bfloat16* qkv, int qkv_ld, const __nv_bfloat16* k_cache,
                               const __nv_bfloat16* v_cache, const DecItem* items, int max_entries,
                               const AttnSeq* seqs, const int32_t* seq_ids, int n_seqs, int max_splits, float* part_o,
                               float* part_ml, __nv_bfloat16* out, int out_ld, int n_q, int n_kv, float scale,
                               int num_sms, cudaStream_t stream, int* arrive = nullptr);
// arrive: [n_seqs * n_kv] ints, zero before the first launch (re-armed by the kernel): with it, the split merge runs inside
// the streaming kernel (last-arriving warp per (sequence, kv head)) and attn_merge_kernel is not launched.
// Prefill / chunked prefill (q_len >= 1), causal over the paged cache.
cudaError_t attn_prefill_launch(const __nv_bfloat16* qkv, int qkv_ld, const __nv_bfloat16* k_cache,
                                const __nv_bfloat16* v_cache, const AttnSeq* seqs, const int32_t* tile_seq,
                                const int32_t* tile_q0, int n_tiles, const int32_t* block_table, int bt_stride,
                                __nv_bfloat16* out, int out_ld, int n_q, int n_kv, float scale, cudaStream_t stream);

// ---- sampler.cu --------------------------------------------------------------------------------------------------
constexpr int MAX_TOPN = 12;  // reference forces max_logprobs >= 11 (tgis_utils/args.py:214-216)
struct SampleRow {            // per sampled row parameters (device array, 64 B)
  int32_t flags;              // bit0 greedy, bit1 want_logprobs, bit2 typical, bit3 len_penalty, bit4 seeded
  int32_t n_topn;             // number of top-n entries to return (0..MAX_TOPN)
  float temperature;
  int32_t top_k;              // <=0: off
  float top_p;                // >=1: off
  float typical_p;
  float rep_penalty;          // 1.0: off
  float len_decay_factor;     // decay ** max(0, n_out - start)  precomputed on host in double, cast to fp32
  int32_t eos_id;
  int32_t n_out;              // tokens generated so far
  int32_t min_tokens;
  int32_t seq_slot;           // row of the seen-token bitmap
  uint32_t seed_lo, seed_hi;  // philox key
  uint32_t step;              // philox counter offset
  int32_t logits_row;         // row in the logits matrix
};
struct SampleOut {            // per row result (pinned host readable)
  int32_t token;
  float logprob;
  int32_t rank;
  int32_t n_topn;
  int32_t topn_ids[MAX_TOPN];
  float topn_lps[MAX_TOPN];
};
constexpr int SAMPLE_GREEDY = 1, SAMPLE_LOGPROBS = 2, SAMPLE_TYPICAL = 4, SAMPLE_LENPEN = 8, SAMPLE_SEEDED = 16;
constexpr int SAMPLE_MASKED = 64;  // guided decoding: row seq_slot of the allow bitmap holds this step's allowed-token bits
constexpr int SAMPLE_FORCED = 32;  // token = seed_lo is given (prompt logprobs): report its logprob / rank / top-n only
// logits: [rows, ld] bf16 (logits_bf16 = 1: the lm_head GEMM's model-dtype output, what vLLM's sampler sees after its
// fp32 cast) or fp32 (kernel-level golden tests).  any_complex: some row needs selection passes (sampling rows: typical-p /
// top-k / top-p / race) -> 8 CTAs per row; otherwise the cluster size shrinks with the row count (sampler_cluster_size).
cudaError_t sampler_launch(const void* logits, int logits_bf16, int ld, int vocab, const SampleRow* rows, int n_rows,
                           const uint32_t* seen_bitmap, int bitmap_words, float* scratch, SampleOut* out,
                           cudaStream_t stream, int any_complex = 1, int num_sms = 148,
                           const uint32_t* allow_bitmap = nullptr /* [slots][bitmap_words], rows flagged SAMPLE_MASKED */);
int sampler_cluster_size(int n_rows, int any_complex, int num_sms);
// seen-token bitmap maintenance
cudaError_t bitmap_clear_launch(uint32_t* bitmap, int bitmap_words, int slot, cudaStream_t stream);
cudaError_t bitmap_set_launch(uint32_t* bitmap, int bitmap_words, const int32_t* slots, const int32_t* tokens, int n,
                              cudaStream_t stream);
size_t sampler_scratch_floats(int vocab);

}  // namespace tgis
