// Internal launchers of the non-GEMM encoder kernels (encoder_kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace crag {

int launch_embed_layernorm(const int32_t* token_ids, const int32_t* cu_seqlens, int n_seqs, int T, int H, int vocab,
                           int max_pos, int pos_offset, const void* word_emb, const void* pos_emb,
                           const void* type_emb, const float* gamma, const float* beta, float eps, void* out,
                           cudaStream_t stream);
int launch_layernorm(const void* in, int T, int H, const float* gamma, const float* beta, float eps, void* out,
                     cudaStream_t stream);
int launch_attention(const void* qkv, const int32_t* cu_seqlens, int n_seqs, int max_len, int H, int heads, void* ctx,
                     cudaStream_t stream);
int launch_attention_tc(const void* qkv, const int32_t* cu_seqlens, int n_seqs, int total_tokens, int max_len, int H,
                        int heads, void* ctx, cudaStream_t stream);
int launch_pool_normalize(const void* hidden, const int32_t* cu_seqlens, int n_seqs, int H, int normalize,
                          float* out_f32, void* out_bf16, int64_t out_bf16_stride, cudaStream_t stream);
int launch_cls_head(const void* hidden, const int32_t* cu_seqlens, int n_seqs, int H, const void* w_dense,
                    const float* b_dense, const void* w_out, const float* b_out, int n_labels, float* logits,
                    cudaStream_t stream);

}  // namespace crag
