/*
 * tgis_kernels.h — C ABI test hooks of libtgis_engine.so: each sm_100a kernel of the hot path callable on raw device
 * pointers, so the parity tests (tests/ -m gpu) can compare every kernel with the CPU oracle through the same shared
 * library the server loads.  All pointers marked "dev" are device memory; metadata arrays are HOST memory and are
 * staged internally.  Every call synchronises before returning; 0 = ok, <0 = error (tgis_last_error()).
 *
 * Reference semantics restated by these kernels (vLLM 0.22 = the arithmetic behind
 * /root/reference/src/vllm_tgis_adapter/grpc/grpc_server.py:222):
 *   tgis_k_gemm        vllm model_executor/layers/linear.py (F.linear, bf16 in, fp32 accumulate, bf16 out)
 *   tgis_k_rmsnorm     vllm model_executor/layers/layernorm.py:104,173 (rms_norm / fused_add_rms_norm)
 *   tgis_k_silu_mul    vllm model_executor/layers/activation.py:117-143
 *   tgis_k_rope_kv     vllm model_executor/layers/rotary_embedding/base.py:200 + reshape_and_cache_flash
 *   tgis_k_attention   vllm v1/attention/backends/flashinfer.py:1665,1803 (paged prefill / decode)
 *   tgis_k_sampler     vllm v1/sample/sampler.py:67-144 + tgis_utils/logits_processors.py:7-47
 */
#ifndef TGIS_KERNELS_H_
#define TGIS_KERNELS_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* y[T,N] = x[T,K] . w[N,K]^T ; y is bf16 (out_f32 = 0) or fp32 (out_f32 = 1, the lm_head logits path);
 * impl 0 = tcgen05 kernel, 1 = SIMT cross-check kernel.  x must have >= 256 rows
 * allocated (TMA box height) or T rows when impl == 1. */
int tgis_k_gemm(const void* x_dev, const void* w_dev, void* y_dev, int32_t T, int32_t N, int32_t K, int32_t x_rows_alloc,
                int32_t impl, int32_t iters, float* ms_out, int32_t out_f32);
/* residual_dev may be NULL (plain rmsnorm); otherwise residual is updated in place */
int tgis_k_rmsnorm(const void* x_dev, void* residual_dev, const void* w_dev, void* out_dev, int32_t T, int32_t hidden,
                   float eps);
int tgis_k_silu_mul(const void* gate_up_dev, void* act_dev, int32_t T, int32_t ffn);
/* OPT (csrc/opt.cu; vllm model_executor/models/opt.py:61-70,148-197 = torch.nn.LayerNorm / F.linear(x, W, b) / ReLU):
 * tgis_k_opt_layernorm: acc_dev != NULL: residual = bf16(residual + bf16(acc + acc_bias)) first (acc fp32 [T, hidden], the
 *   row-"producer" GEMM's accumulators); out = bf16((residual - mean) * rstd * w + b).  residual is bf16 [T, hidden].
 * tgis_k_opt_bias_act: out[T, N] = bf16(act(acc[T, N] + bias[N])), act = ReLU (relu != 0) or identity.
 * tgis_k_opt_embed: out[t] = bf16(tok_table[tok[t]] + pos_table[pos[t] + offset]); tok / pos are int32 DEVICE arrays. */
int tgis_k_opt_layernorm(const void* acc_dev, const void* acc_bias_dev, void* residual_dev, const void* w_dev,
                         const void* b_dev, void* out_dev, int32_t T, int32_t hidden, float eps);
int tgis_k_opt_bias_act(const void* acc_dev, const void* bias_dev, void* out_dev, int32_t T, int32_t N, int32_t relu);
int tgis_k_opt_embed(const void* tok_dev, const void* pos_dev, const void* tok_table_dev, const void* pos_table_dev,
                     void* out_dev, int32_t T, int32_t hidden, int32_t vocab, int32_t n_pos_rows, int32_t offset);
int tgis_k_rope_kv(void* qkv_dev, const int32_t* positions_host, const int32_t* slot_mapping_host,
                   const void* cos_sin_dev, void* k_cache_dev, void* v_cache_dev, int32_t T, int32_t n_q, int32_t n_kv);
/* qkv projection y[T, (n_q + 2 n_kv) * 128] = x . w^T with RoPE and the KV-cache scatter fused into the GEMM's
 * split-tile reduction (cluster / DSMEM or global fix-up); returns 1 (and computes nothing) when the launch plan does not
 * split every weight tile */
int tgis_k_gemm_rope(const void* x_dev, const void* w_dev, void* y_dev, int32_t T, int32_t n_q, int32_t n_kv, int32_t K,
                     int32_t x_rows_alloc, const int32_t* positions_host, const int32_t* slot_mapping_host,
                     const void* cos_sin_dev, void* k_cache_dev, void* v_cache_dev);
/* a[T, K1] -> GEMM w1[H, K1] -> residual += . ; RMSNorm(residual) * w_norm -> GEMM w2[N2, H] (out_mode2: 0 bf16, 2 fused
 * SwiGLU) -> y2.  fused = 0: three launches (gemm, add_rmsnorm, gemm); fused = 1: the add + norm folded into the two GEMMs
 * (GemmNorm, kernels.h).  Returns 1 (nothing computed) when the launch plan cannot fuse the shape. */
int tgis_k_gemm_norm_chain(const void* a_dev, int32_t a_rows_alloc, const void* w1_dev, void* residual_dev,
                           const void* w_norm_dev, const void* w2_dev, void* y2_dev, int32_t T, int32_t K1, int32_t H,
                           int32_t N2, float eps, int32_t out_mode2, int32_t fused, int32_t iters, float* us_out);
/* seqs_host: n_seqs x {q_start, q_len, kv_len, block_row}; block_table_host: [rows][bt_stride] */
int tgis_k_attention(const void* qkv_dev, const void* k_cache_dev, const void* v_cache_dev, const int32_t* seqs_host,
                     int32_t n_seqs, const int32_t* block_table_host, int32_t bt_rows, int32_t bt_stride, void* out_dev,
                     int32_t n_q, int32_t n_kv, float scale);
/* host-only: GEMM launch plan for a shape: token tile (16..256), grid, even-split factor (0: stream-K + global fix-up) */
int tgis_k_gemm_plan(int32_t T, int32_t N, int32_t K, int32_t num_sms, int32_t* bt_out, int32_t* grid_out,
                     int32_t* even_split_out);
/* weight rows per GEMM unit for a T-token launch (128 x weight tiles per shared activation tile) */
int tgis_k_gemm_unit_rows(int32_t T);
/* host-only: the decode work-item list of a step (csrc/kernels.h DecItem: q_row, kv_len, seq | split << 16, 0,
 * blocks[4]); record 0 holds the entry count; entries are listed longest first */
int tgis_k_decode_items(const int32_t* seqs_host, int32_t n_seqs, const int32_t* block_table_host, int32_t bt_stride,
                        int32_t* items_out, int32_t capacity);
/* decode-only, timed over `iters` launches rotating through n_layers cache copies; us_out = avg device us per launch */
int tgis_k_attention_bench(const void* qkv_dev, const void* k_cache_dev, const void* v_cache_dev,
                           const int32_t* seqs_host, int32_t n_seqs, const int32_t* block_table_host, int32_t bt_rows,
                           int32_t bt_stride, void* out_dev, int32_t n_q, int32_t n_kv, float scale, int32_t n_layers,
                           int64_t layer_stride_bytes, int32_t iters, float* us_out);
/* rows_host: n_rows x 64-byte SampleRow records (see csrc/kernels.h); out_host: n_rows x 112-byte SampleOut records;
 * seen_bitmap_dev: [slots][ceil(vocab/32)] uint32 or NULL */
/* logits_dev: fp32 [rows, ld] */
int tgis_k_sampler(const void* logits_dev, int32_t ld, int32_t vocab, const void* rows_host, int32_t n_rows,
                   void* seen_bitmap_dev, void* out_host);
/* same, logits_bf16 = 1: logits_dev is bf16 [rows, ld] (the product path's dtype); iters > 1 repeats the launch and
 * us_out (may be NULL) receives the average device time per launch */
int tgis_k_sampler_ex(const void* logits_dev, int32_t logits_bf16, int32_t ld, int32_t vocab, const void* rows_host,
                      int32_t n_rows, void* seen_bitmap_dev, void* out_host, int32_t iters, float* us_out);
/* same, with the guided-decoding allow bitmap: allow_bitmap_dev [slots][ceil(vocab/32)] uint32 (bit set = token allowed;
 * xgrammar's token-bitmask layout) is applied to the rows flagged SAMPLE_MASKED (64) through their seq_slot; a cleared bit
 * makes the raw logit -inf before every other stage (vllm v1/structured_output/utils.py apply_grammar_bitmask) */
int tgis_k_sampler_masked(const void* logits_dev, int32_t logits_bf16, int32_t ld, int32_t vocab, const void* rows_host,
                          int32_t n_rows, void* seen_bitmap_dev, const void* allow_bitmap_dev, void* out_host,
                          int32_t iters, float* us_out);
/* LoRA (csrc/lora.cu; vllm lora/punica_wrapper/punica_gpu.py add_lora_linear): x [T, ldx] bf16, tok_slot [T] int32 (0 = no
 * adapter, s >= 1 = slot s), A [slots][Rm][K] bf16, B [slots][N][Rm] bf16 (scaling folded in); in place on y [T, ldy] bf16:
 * y[t, col0 + n] = bf16(y + bf16(B_s[n, :] . bf16(A_s x[t]))) */
int tgis_k_lora(const void* x_dev, int32_t ldx, const int32_t* tok_slot_dev, const void* a_dev, const void* b_dev, int32_t K,
                int32_t N, int32_t Rm, int32_t col0, void* y_dev, int32_t ldy, int32_t T);
/* same kernels, col0 = 0, timed: iters launches of each, average device microseconds per launch */
int tgis_k_lora_bench(const void* x_dev, int32_t ldx, const int32_t* tok_slot_dev, const void* a_dev, const void* b_dev,
                      int32_t K, int32_t N, int32_t Rm, void* y_dev, int32_t ldy, int32_t T, int32_t iters,
                      float* us_shrink, float* us_expand);
/* act[t, j] = bf16(silu(gate_up[t, 2j])) * gate_up[t, 2j + 1] */
int tgis_k_silu_mul_interleaved(const void* gate_up_dev, void* act_dev, int32_t T, int32_t ffn);
const char* tgis_k_last_error(void);
/* debug builds only (-DTGIS_GEMM_TIMELINE): %globaltimer stamps of CTA 0 and CTA grid/2, [4][16] u64; else -2 */
int tgis_k_gemm_timeline(uint64_t* out64);
/* debug builds only (-DTGIS_STEP_TIMELINE; else -2): CTA 0 of every kernel launch records {kernel id, %globaltimer at
 * entry / after the grid-dependency wait / at exit}; read: out[0] = records, then out[8 + 4 i ..] (4096-entry ring) */
int tgis_k_step_timeline_enable(void);
int tgis_k_step_timeline_read(uint64_t* out);
int tgis_k_sizeof_sample_row(void);
int tgis_k_sizeof_sample_out(void);
int tgis_k_kv_block(void);

#ifdef __cplusplus
}
#endif
#endif
