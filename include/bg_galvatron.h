/*
 * bg_galvatron.h -- C ABI of the B200-native hot path behind Hetu-Galvatron's per-layer strategy API.
 *
 * The reference (PKU-DAIR/Hetu-Galvatron v2.4.1) has no FFI at this seam: every per-layer collective is a
 * torch.distributed (ProcessGroupNCCL) call made from Python.  Each entry point below names the reference
 * call site it replaces (file:line under /root/reference unless prefixed torch/).  The host side that binds
 * these (ctypes) is hetu-galvatron_b200/_bg.py; INTEGRATION.md shows the binding a maintainer would add.
 *
 * Conventions: plain pointers and sizes, no torch types; every call returns 0 on success or a negative
 * BG_E* code (message via bg_last_error()); nothing throws across the ABI.  All device work is asynchronous
 * on the cudaStream_t passed as `stream` (void*).  One bg_ctx per rank (one process per GPU; several ctx in
 * one process = "virtual ranks" on one device, used by the single-GPU parity tests).  Thread-compatible:
 * callable from the main and the autograd thread; per-(group,lane) ordering is the caller's stream order.
 *
 * Peer memory model: each rank owns ONE symmetric arena (cudaMalloc, exported with cudaIpcGetMemHandle,
 * mapped by its peers once).  A "symmetric buffer" is an array of per-member arena offsets (group order).
 * Cross-rank synchronisation is device-side: per-(group,lane,CTA) flags in the arena head, CAS put/wait
 * with release/acquire at .sys scope; no host sync, no NCCL on these paths.
 */
#ifndef BG_GALVATRON_H
#define BG_GALVATRON_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BG_ABI_VERSION 1
#define BG_MAX_PEERS 8      /* one NVSwitch domain */
#define BG_MAX_WORLD 64
#define BG_LANES 6          /* independent barrier lanes per group: unshard / grad-reduce / activations / misc / fused-op push / spare */
#define BG_MAX_CHANNELS 256 /* max CTAs of a cross-rank kernel (one barrier channel per CTA; the default is one slim CTA per SM) */

typedef struct bg_ctx* bg_ctx_t;

enum bg_dtype { BG_BF16 = 0, BG_F32 = 1 };
enum bg_redop { BG_SUM = 0, BG_MAX = 1 };
enum bg_err {
    BG_OK = 0, BG_EINVAL = -1, BG_ECUDA = -2, BG_ENOMEM = -3, BG_ENOTMAPPED = -4, BG_EGROUP = -5, BG_ETIMEOUT = -6,
    BG_EUNSUPPORTED = -7
};

/* ---- library ------------------------------------------------------------------------------------------ */
int bg_abi_version(void);
const char* bg_last_error(void);                       /* thread-local message of the last failing call */
/* "comm_ctas" (CTAs of a cross-rank kernel, default 148 = one 128-thread / <=64-register CTA per SM), "local_ctas", "timeout_ms",
 * "oneshot_bytes", "nvls_min_bytes" (multicast paths above this size), "nvls_min_ranks" (... and from this group size on), "nvls_gather" [0] / "nvls_reduce" [1]
 * (multimem.st for all-gather / multimem.ld_reduce for reduce-scatter on multicast-bound buffers; defaults from the p = 8 measurement),
 * "nvls_bcast" [1] (the fused GEMM + all-reduce writes a reduced tile to every member with one multimem.st) */
int bg_set_tunable(const char* name, long long value);
long long bg_get_tunable(const char* name);
/* how many kernels this library has launched since load (bench.py's gpu_launches) */
unsigned long long bg_launch_count(void);

/* ---- context + symmetric arena (replaces ProcessGroupNCCL communicator state) --------------------------- */
int bg_ctx_create(int rank, int world, int device, size_t arena_bytes, bg_ctx_t* out);
/* flags: BG_CTX_VMM allocates the arena with the virtual-memory API (cuMemCreate) instead of cudaMalloc; that is what NVSwitch
 * multicast objects can bind.  Such an arena is shared between processes as a POSIX file descriptor (bg_arena_export_fd /
 * bg_arena_import_fd; the host passes the descriptor over a unix socket) instead of a cudaIpc handle.  OPT-IN. */
#define BG_CTX_VMM 1u
int bg_ctx_create_ex(int rank, int world, int device, size_t arena_bytes, unsigned flags, bg_ctx_t* out);
int bg_ctx_destroy(bg_ctx_t ctx);
int bg_arena_info(bg_ctx_t ctx, void** base, size_t* bytes, size_t* used);
int bg_arena_alloc(bg_ctx_t ctx, size_t bytes, size_t* offset);        /* 256-B aligned bump allocation */
int bg_arena_export(bg_ctx_t ctx, void* handle64);                      /* cudaIpcMemHandle_t, 64 bytes */
int bg_arena_import(bg_ctx_t ctx, int peer_rank, const void* handle64); /* map a peer process's arena */
int bg_arena_attach_local(bg_ctx_t ctx, int peer_rank, bg_ctx_t peer);  /* peer ctx in this process */
int bg_arena_alloc_aligned(bg_ctx_t ctx, size_t bytes, size_t align, size_t* offset);
int bg_arena_mode(bg_ctx_t ctx, int* vmm, int* multicast_supported, size_t* multicast_granularity);
int bg_arena_export_fd(bg_ctx_t ctx, int* fd);                          /* BG_CTX_VMM arenas */
int bg_arena_import_fd(bg_ctx_t ctx, int peer_rank, int fd);
int bg_ctx_error_flag(bg_ctx_t ctx, int* flag);                         /* device-side timeout report */
/* info8[0] = status; [1] = kind (1 signalling a peer, 2 waiting for a peer, 3 fused-GEMM tile reducer);
 * [2] = CTA; [3] = thread or tile; [4] = value last seen; [5], [6] = group index/size or expected count/tiles;
 * [7] = launch site of a barrier timeout (1 all-gather, 2 reduce-scatter, 3 all-reduce, 6 all-to-all, 7 entry barrier of a fused
 * GEMM, 8 exit barrier of the all-reduce tile reducer, 10/11 p2p flags, 12 push kernel of the fused all-gather+GEMM), kind 4 = the
 * gathering GEMM's TMA producer waiting for a block.
 * The reference's analogue is the NCCL watchdog's timeout dump (ProcessGroupNCCL); here a lost peer traps the
 * kernel and leaves this record in mapped host memory. */
int bg_ctx_error_info(bg_ctx_t ctx, int* info8);

/* ---- groups (galvatron/core/runtime/comm_groups.py:7-29 CommGroup / :416 gen_comm_groups) ---------------- */
/* A group is its rank list; creation is O(1), local, and needs no collective (the reference pays one
 * torch.distributed.new_group per group on all ranks, comm_groups.py:14).  Rank lists must be arithmetic
 * progressions (every Galvatron group is). */
int bg_group_create(bg_ctx_t ctx, const int* ranks, int n, int* gid);
int bg_group_info(bg_ctx_t ctx, int gid, int* n, int* my_index, int* ranks_out /* BG_MAX_PEERS */);
/* C mirror of gen_comm_groups for one rank: for each of the n_layers whole-model rows writes the caller's
 * tp/sp/cp/dp/sdp rank lists as (count, ranks[BG_MAX_WORLD]) records; used by the bit-exact mapping test. */
int bg_build_groups(int rank, int world, int pp_size, int n_layers, const int* tp, const int* sp, const int* cp,
                    int* out_counts /* [5][n_layers] */, int* out_ranks /* [5][n_layers][BG_MAX_WORLD] */,
                    int* pp_count, int* pp_ranks /* [BG_MAX_WORLD] */);

/* ---- collectives ----------------------------------------------------------------------------------------- */
/* device-side barrier over the group (replaces torch.distributed.barrier, pipeline.py:370,698,882) */
int bg_barrier(bg_ctx_t ctx, int gid, int lane, void* stream);

/* C1  param all-gather fused with the fp32->bf16 cast.
 * Replaces torch/distributed/fsdp/_flat_param.py:1477 all_gather_into_tensor (+ the MixedPrecision shard cast,
 * galvatron/core/runtime/parallel.py:116-122); also the plain activation all-gathers mappings_group.py:100,
 * layers.py:410-412, redistribute.py:65,110 when src_dtype == dst_dtype.
 * Push: every member casts its shard once and stores it into slot `my_index` of every member's dst. */
int bg_all_gather_cast(bg_ctx_t ctx, int gid, int lane, const void* src, int src_dtype, const size_t* dst_offs,
                       int dst_dtype, size_t shard_elems, void* stream);

/* C2  gradient reduce-scatter fused with pre/post-divide, cast and accumulate.
 * Replaces torch/distributed/fsdp/_runtime_utils.py:852 (prediv) :858 reduce_scatter_tensor :879 (postdiv)
 * :917-924 (cast to param dtype, += _saved_grad_shard), reached from sp_grad_reduce.py:125-126; with
 * dst_dtype bf16 and accumulate 0 it is the Megatron-SP reduce-scatter (mappings_group.py:120, layers.py:488-494).
 * Pull: member i reads slice i of every member's src, sums in fp32, dst = [dst +] sum * prescale * postscale. */
int bg_reduce_scatter_acc(bg_ctx_t ctx, int gid, int lane, const size_t* src_offs, int src_dtype, void* dst,
                          int dst_dtype, size_t shard_elems, float prescale, float postscale, int accumulate,
                          void* stream);

/* C2 + optimizer: the same pull reduce-scatter, but the reduced gradient is consumed in registers by an AdamW
 * step on the caller's fp32 shard (param, exp_avg, exp_avg_sq) -- the fp32 gradient shard is never written.  Update rule of
 * torch.optim.AdamW / apex FusedAdam adam_w_mode (galvatron/core/runtime/utils.py:137-150); `step` >= 1 for bias correction. */
int bg_reduce_scatter_adamw(bg_ctx_t ctx, int gid, int lane, const size_t* src_offs, int src_dtype, float* param, float* exp_avg,
                            float* exp_avg_sq, size_t shard_elems, float prescale, float postscale, float lr, float beta1,
                            float beta2, float eps, float weight_decay, long long step, void* stream);

/* C3/C5/C6/C13/C14/C16  all-reduce (sum|max), out of place: src is a symmetric buffer, dst any local pointer.
 * Replaces _runtime_utils.py:940 (DDP grads), mappings_group.py:19 _reduce (row-parallel fwd, layers.py:1114;
 * column-parallel bwd, mappings_group.py:139), cross_entropy.py:22-30,61-72,78-89, grad_reduce.py:121-124.
 * One-shot below the "oneshot_bytes" tunable, two-shot (reduce own slice, then gather) above. */
int bg_all_reduce(bg_ctx_t ctx, int gid, int lane, const size_t* src_offs, void* dst, size_t elems, int dtype,
                  int redop, float scale, void* stream);

/* C10  Ulysses all-to-all fused with the head/seq transpose (one pass, q/k/v in one launch).
 * Replaces transformer.py:1928-1987 single_all_to_all (permute+contiguous, dist.all_to_all_single :1977,
 * post_all2all :1904-1925).  Pull: for each peer q and tensor t, rows of `row_elems[t]` contiguous elements:
 *   dst_t[b*dst_bs + r*dst_rs + q*dst_peer_off + c] = src_t(peer q)[b*src_bs + r*src_rs + me*src_me_off + c] */
typedef struct bg_a2a_desc {
    const size_t* src_offs; /* symmetric source (arena offsets, group order) */
    void* dst;              /* local destination */
    long long batch, rows, row_elems;
    long long src_bs, src_rs, src_me_off; /* element strides in the peer's source */
    long long dst_bs, dst_rs, dst_peer_off;
} bg_a2a_desc;
int bg_all_to_all_rows(bg_ctx_t ctx, int gid, int lane, const bg_a2a_desc* descs, int n_descs, int dtype,
                       void* stream);

/* C11  pipeline p2p: copy `bytes` into the peer's arena at dst_off on `stream`, then raise flag `flag_id` there;
 * the receiver's bg_p2p_wait orders its stream after the flag (no device-wide sync) and bg_p2p_release hands
 * the slot back (a sender's 2nd..nth use of a flag first waits for that release).
 * Replaces pipeline.py:1095-1127 batch_isend_irecv + :1244 torch.cuda.synchronize(). */
int bg_p2p_send(bg_ctx_t ctx, int peer_rank, size_t dst_off, const void* src, size_t bytes, int flag_id,
                void* stream);
int bg_p2p_wait(bg_ctx_t ctx, int peer_rank, int flag_id, void* stream);
int bg_p2p_release(bg_ctx_t ctx, int peer_rank, int flag_id, void* stream);

/* ---- local fused elementwise ops adjacent to the collectives (K5/K6/K9/a10 in SURVEY 2.3) ------------------ */
int bg_cast(const void* src, int src_dtype, void* dst, int dst_dtype, size_t elems, float scale, int accumulate,
            void* stream);
int bg_rmsnorm_fwd(const void* x, const void* w, void* y, float* rstd, long long rows, long long cols, float eps,
                   void* stream);
int bg_rmsnorm_bwd(const void* dy, const void* x, const void* w, const float* rstd, void* dx, float* dw_partial,
                   long long rows, long long cols, int n_partial, void* stream);
/* LayerNorm with bias (GPT / BERT families: gpt_hf/GPTModel_tensor_parallel.py:34,56; fp32 math, one rounding).  backward:
 * dw_partial / db_partial are [n_partial][cols] fp32 per-CTA partial sums (the caller adds them up). */
int bg_layernorm_fwd(const void* x, const void* w, const void* b, void* y, float* mean, float* rstd, long long rows, long long cols,
                     float eps, void* stream);
int bg_layernorm_bwd(const void* dy, const void* x, const void* w, const float* mean, const float* rstd, void* dx, float* dw_partial,
                     float* db_partial, long long rows, long long cols, int n_partial, void* stream);
/* bias + GeLU of the GPT / BERT MLP (transformer.py:150-160 bias_gelu_impl): out = gelu(x + bias) when dy == NULL, else
 * out = dy * gelu'(x + bias).  tanh_form 1 = Megatron's fused / HF gelu_new, 0 = exact erf.  bias may be NULL. */
int bg_bias_gelu(const void* x, const void* bias, const void* dy, void* out, long long rows, long long cols, int tanh_form, void* stream);
int bg_swiglu_fwd(const void* gate_up, void* y, long long rows, long long ffn, void* stream);
int bg_swiglu_bwd(const void* dy, const void* gate_up, void* dgate_up, long long rows, long long ffn, void* stream);
/* fused QKV split + RoPE + [s,b,ng,(r+2)*hn] -> q [b,s,ng*r,hn], k/v [b,s,ng,hn] relayout; backward=1 is the exact
 * transpose (reads dq,dk,dv, writes dmixed).  Replaces the split / repeat_interleave / apply_rotary_pos_emb /
 * rearrange(...).contiguous() chain of transformer.py:731-767,842-867 (GQA stays un-expanded: flash-attn is
 * called with ng KV heads).  cos/sin: [s, hn/2] fp32 tables for this rank's positions. */
int bg_qkv_rope(void* mixed, void* q, void* k, void* v, const float* cos_t, const float* sin_t, long long s, long long b,
                long long ng, long long r, long long hn, int backward, void* stream);

/* a10  vocab-parallel cross-entropy (cross_entropy.py:14-152) as three row kernels around the two small
 * all-reduces (MAX of rowmax; SUM of (sum_exp, predicted_logit)).  bg_ce_bwd overwrites logits with dlogits. */
int bg_ce_rowmax(const void* logits, int dtype, float* rowmax, long long rows, long long vocab_local, void* stream);
int bg_ce_sumexp(const void* logits, int dtype, const long long* target, const float* rowmax, float* out2 /* [rows][2] */,
                 long long rows, long long vocab_local, long long vocab_start, void* stream);
int bg_ce_bwd(void* logits, int dtype, const long long* target, const float* rowmax, const float* sum2,
              const float* grad_loss, long long rows, long long vocab_local, long long vocab_start, void* stream);

/* ---- GEMM (K1): bf16 x bf16 -> fp32 accumulate in TMEM -> bf16, tcgen05 + TMA ------------------------------ */
/* C[M,N] (+)= A op B.  layout: 0 = "TN" C = A[M,K] * B[N,K]^T (forward, layers.py:417);
 * 1 = "NN" C = A[M,K] * B[K,N] (dgrad, layers.py:462); 2 = "NT" C = A[K,M]^T * B[K,N] (wgrad, layers.py:534). */
int bg_gemm_bf16(const void* a, const void* b, void* c, long long m, long long n, long long k, int layout,
                 int accumulate, void* stream);
/* C = A op B + addend ([M][N] bf16): the residual add behind a projection (llama_hf/LlamaModel_tensor_parallel.py:83,100
 * `hidden_states + input_tensor`) in the GEMM epilogue -- one rounding, no separate elementwise pass.  addend may alias c. */
int bg_gemm_bf16_add(const void* a, const void* b, void* c, const void* addend, long long m, long long n, long long k, int layout,
                     void* stream);

/* C5/C8 fused with K1: C = A op B is computed in 128x256 tcgen05 tiles and REDUCE-SCATTERED along M over the group inside
 * the same operation -- every finished partial tile is TMA-stored into the owning rank's arena (peer HBM over NVLink) and
 * counted there; a tile reducer on the owner -- launched into the same stream as the GEMM's programmatic dependent, so it
 * runs beside the GEMM CTAs once all of them are resident -- sums the p partials as they land and writes out[M/p, N].
 * Replaces layers.py:1061-1109 (row-parallel GEMM then mappings_group.py:120 reduce-scatter) and :462,488-494 (dgrad +
 * reduce-scatter).  partial_offs: symmetric bf16 buffer of M*N elements; flag_offs: symmetric u32[(M/p/128)*ceil(N/256)],
 * zero-initialised.  M must be a multiple of p*128.  `out` is complete in stream order. */
int bg_gemm_reduce_scatter(bg_ctx_t ctx, int gid, int lane, const void* a, const void* b, long long m, long long n, long long k,
                           int layout, const size_t* partial_offs, const size_t* flag_offs, void* out, void* stream);

/* C5/C6 fused with K1: GEMM + ALL-REDUCE (row-parallel forward, layers.py:1110-1114: matmul then reduce_from_tensor_model_parallel
 * _region; column-parallel dgrad, mappings_group.py:139).  Both shots of the two-shot all-reduce happen inside the fused
 * operation: partial tiles are scattered to their owners as in bg_gemm_reduce_scatter; the owner's tile reducer sums a tile when
 * its p partials have landed and immediately broadcasts the rows into EVERY member's out buffer (peer stores, or one multimem.st
 * when out is multicast-bound); the reducers leave through a cross-rank barrier, so out -- a symmetric bf16 [M][N] buffer,
 * out_offs -- is complete on every member in stream order.  Deterministic (fixed summation order), fp32 accumulation. */
int bg_gemm_all_reduce(bg_ctx_t ctx, int gid, int lane, const void* a, const void* b, long long m, long long n, long long k, int layout,
                       const size_t* partial_offs, const size_t* flag_offs, const size_t* out_offs, void* stream);

/* C7 fused with K1: ALL-GATHER + GEMM (column-parallel forward under Megatron-SP, layers.py:399-417: _all_gather_base into the
 * global buffer, then matmul; also the row-parallel dgrad, whose dY is gathered, mappings_group.py:243-258).
 * C[M,N] = gather_M(a_local[M/p,K]) op B.  A slim push kernel on comm_stream (resident beside the GEMM) sends the local rows to
 * every member's staging buffer (stage_offs: symmetric bf16 [M][K]) in 128-row blocks and counts each block in on the receiver
 * (flag_offs: symmetric u32[p][M/p/128], zero-initialised); the GEMM's TMA producer reads the rank's own rows from a_local and
 * every remote block from staging as soon as its counter is complete, walking the blocks in arrival order.  layout 0 (TN) or
 * 1 (NN); M a multiple of p*128.  C is complete in `stream` order; a_local may be reused after the call in `stream` order. */
int bg_all_gather_gemm(bg_ctx_t ctx, int gid, int lane, const void* a_local, const size_t* stage_offs, const size_t* flag_offs,
                       const void* b, void* c, long long m, long long n, long long k, int layout, void* stream, void* comm_stream);

/* ---- NVLS: all-reduce reduced and replicated INSIDE the NVSwitch (multimem.ld_reduce / multimem.st) ---------------------
 * Replaces the tensor-parallel all-reduces (mappings_group.py:19, layers.py:474-480) for large messages, where NCCL itself
 * switches to NVLS: N/p bytes in and N/p out per GPU instead of 2(p-1)/p*N.  One multicast object per group over one
 * symmetric buffer: the group's first rank creates it (-> fd), every other member imports the fd, ALL join (device added),
 * barrier on the host, ALL bind their own arena range (multicast-granularity aligned), then bg_all_reduce_nvls works in
 * place on that buffer and copies the result to dst (dst may be NULL: result left in the buffer).  BG_CTX_VMM contexts only.
 * The bound range may be any part of the arena (the host binds the whole region that holds its symmetric buffers): every
 * collective whose buffer lies inside it at the same offset on all members then uses the switch on its own --
 * bg_all_reduce (ld_reduce + st), bg_all_gather_cast and the bg_gemm_all_reduce broadcast (multimem.st), bg_reduce_scatter_acc /
 * _adamw (multimem.ld_reduce) -- above "nvls_min_bytes". */
int bg_group_mc_create(bg_ctx_t ctx, int gid, size_t bytes, int* fd_out);
int bg_group_mc_join(bg_ctx_t ctx, int gid, int fd /* -1 on the creator */, size_t bytes);
int bg_group_mc_bind(bg_ctx_t ctx, int gid, size_t arena_offset);
int bg_group_mc_disable(bg_ctx_t ctx, int gid);   /* setup failed on some member: the group keeps the peer-to-peer kernels */
int bg_all_reduce_nvls(bg_ctx_t ctx, int gid, int lane, size_t byte_offset, void* dst, size_t elems, int dtype, float scale,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* BG_GALVATRON_H */
