// bg_gemm.cu -- K1: bf16 GEMM on the 5th-gen tensor cores (tcgen05.mma, fp32 accumulators in TMEM, operands staged
// by TMA with 128-B swizzle), persistent, warp-specialised.  Replaces the torch.matmul -> cuBLAS calls of
// galvatron/site_package/megatron/core/tensor_parallel/layers.py:417 (fwd), :462 (dgrad), :534 (wgrad).
//
//   tile            : BLOCK_M 128 x BLOCK_N 256 x BLOCK_K 64, one CTA per SM, UMMA 128x256x16 (cta_group::1)
//   smem pipeline   : 4 stages x (A 16 KiB + B 32 KiB), full/empty mbarriers (TMA <-> MMA)
//   TMEM            : 2 accumulator buffers x 256 columns (MMA of tile i+1 overlaps epilogue of tile i)
//   warps           : 0 = TMA producer, 1 = MMA issuer, 2 = TMEM alloc, 4-7 = epilogue (TMEM -> regs -> swizzled smem
//                     -> TMA store), 256 threads
//   layouts         : TN  C = A[M,K] * B[N,K]^T   (A, B K-major)
//                     NN  C = A[M,K] * B[K,N]     (B MN-major: TMA boxes of 64 N-elements x 64 K-rows)
//                     NT  C = A[K,M]^T * B[K,N]   (A, B MN-major)
//   edges           : TMA zero-fills out-of-bounds loads and clips stores, so M, N, K only need to be multiples of 8.
#include <stdlib.h>

#include "bg_ctx.cuh"

using namespace bg;

namespace {

constexpr int BLOCK_M = 128, BLOCK_N = 256, BLOCK_K = 64, UMMA_K = 16;
constexpr int kStages = 4, kAccStages = 2;
constexpr int kABytes = BLOCK_M * BLOCK_K * 2, kBBytes = BLOCK_N * BLOCK_K * 2, kStageBytes = kABytes + kBBytes;
constexpr int kStoreCols = 64;                                 // columns per TMA store box (128 B)
constexpr int kStoreBytes = BLOCK_M * kStoreCols * 2;          // 16 KiB per staging buffer
constexpr int kNumStoreBufs = 2;
constexpr int kSmemBytes = kStages * kStageBytes + kNumStoreBufs * kStoreBytes + 1024 /*align slack*/ + 256 /*barriers*/;
constexpr int kThreads = 256, kEpiThreads = 128;
constexpr int kGroupM = 16;  // tile raster: 16 m-blocks share each sweep over n (L2 reuse)

enum Layout { kTN = 0, kNN = 1, kNT = 2 };

// ---- PTX wrappers ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(bar), "r"(parity) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(map), "r"(bar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, uint32_t src, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(map), "r"(src), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_all() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
        ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
          "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
          "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(kEpiThreads) : "memory"); }
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout): start[0,14) lbo[16,30) sbo[32,46)
// version=1 [46,48) layout_type[61,64) (2 = SWIZZLE_128B)
__device__ __forceinline__ uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr >> 4) & 0x3fff);
    d |= (uint64_t)((lbo_bytes >> 4) & 0x3fff) << 16;
    d |= (uint64_t)((sbo_bytes >> 4) & 0x3fff) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}

// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b format BF16 [7,10)/[10,13)=1,
// a_major bit 15, b_major bit 16 (0 = K-major, 1 = MN-major), n>>3 [17,23), m>>4 [24,29)
__host__ __device__ constexpr uint32_t make_idesc(bool a_mn, bool b_mn) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((a_mn ? 1u : 0u) << 15) | ((b_mn ? 1u : 0u) << 16) |
           ((uint32_t)(BLOCK_N >> 3) << 17) | ((uint32_t)(BLOCK_M >> 4) << 24);
}

struct TileCoord { int m, n; };
// kMode: which collective is fused into the GEMM
enum { kPlain = 0, kScatterMode = 1, kGatherMode = 2 };

template <int kMode>
struct FuseParams {};

__device__ __forceinline__ TileCoord tile_of_virtual(int t, int m_blocks, int n_blocks) {
    const int per_group = kGroupM * n_blocks;
    const int g = t / per_group, first_m = g * kGroupM;
    const int rows = min(kGroupM, m_blocks - first_m);
    const int r = t - g * per_group;
    return {first_m + r % rows, r / rows};
}

// Fused GEMM + reduce-scatter (C5/C8): instead of storing C locally, the epilogue TMA-stores every finished 128x256 partial
// tile straight into the HBM of the rank that owns those rows (peer store over NVLink) and bumps that rank's per-tile
// arrival counter; a small reducer kernel on the owner sums the p partials of a tile as soon as all have landed.  Transfer and
// math overlap tile by tile; no NCCL, no separate collective pass over the full activation.
template <>
struct FuseParams<kScatterMode> {
    CUtensorMap dst[BG_MAX_PEERS];      // owner o's partial buffer viewed as [p * rows_per_rank][N] (block r = source rank r)
    uint32_t* flags[BG_MAX_PEERS];      // owner o's arrival counters, one per local tile
    int p, me, rows_per_rank;
};

// Fused all-gather + GEMM (C7): A[M,K] is the concatenation of the members' [M/p, K] shards.  The rank's own rows are read
// from its local shard; every other 128-row block is read from the staging slot the owner's push kernel fills, as soon as that
// block's arrival counter shows all pushing CTAs have delivered it.  Blocks are walked in arrival order: block c of slot me,
// me+1, ..., me-1, then block c+1 of every slot, so the first sweep over N needs only the first block of every slot.
template <>
struct FuseParams<kGatherMode> {
    CUtensorMap a_local;                // [M/p][K]
    const uint32_t* flags;              // [p][blocks_per_rank] arrival counters in THIS rank's arena
    uint32_t target;                    // CTAs of the push kernel
    int p, me, blocks_per_rank;
    unsigned long long timeout_ns;
    int* err;
};

template <int kMode>
__device__ __forceinline__ int map_m(int mv, int m_blocks, const FuseParams<kMode>& fp) {
    if constexpr (kMode == kScatterMode) {
        // rank r walks the owners in the order r+1, r+2, ..., r (ring schedule): every owner receives from ONE peer at a time
        int m = mv + ((fp.me + 1) % fp.p) * (fp.rows_per_rank / BLOCK_M);
        return m >= m_blocks ? m - m_blocks : m;
    } else if constexpr (kMode == kGatherMode) {
        const int c = mv / fp.p, k = mv - c * fp.p;
        int slot = fp.me + k; if (slot >= fp.p) slot -= fp.p;
        return slot * fp.blocks_per_rank + c;
    } else {
        return mv;
    }
}
template <int kMode>
__device__ __forceinline__ TileCoord tile_of(int t, int m_blocks, int n_blocks, const FuseParams<kMode>& fp) {
    TileCoord tc = tile_of_virtual(t, m_blocks, n_blocks);
    tc.m = map_m<kMode>(tc.m, m_blocks, fp);
    return tc;
}

template <int kLayout, int kMode = kPlain>
__global__ void __launch_bounds__(kThreads, 1) gemm_bf16_kernel(const __grid_constant__ CUtensorMap map_a,
                                                                const __grid_constant__ CUtensorMap map_b,
                                                                const __grid_constant__ CUtensorMap map_c,
                                                                const __nv_bfloat16* __restrict__ c_old, int M, int N, int K,
                                                                int accumulate,
                                                                const __grid_constant__ FuseParams<kMode> sp) {
    constexpr bool kScatter = kMode == kScatterMode, kGather = kMode == kGatherMode;
    constexpr bool kAMn = kLayout == kNT, kBMn = kLayout != kTN;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_store = smem + kStages * kStageBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_store + kNumStoreBufs * kStoreBytes);
    uint64_t* full_bar = bars;                       // [kStages]
    uint64_t* empty_bar = bars + kStages;            // [kStages]
    uint64_t* tmem_full = bars + 2 * kStages;        // [kAccStages]
    uint64_t* tmem_empty = tmem_full + kAccStages;   // [kAccStages]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + kAccStages);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m_blocks = (M + BLOCK_M - 1) / BLOCK_M, n_blocks = (N + BLOCK_N - 1) / BLOCK_N;
    const int num_tiles = m_blocks * n_blocks, k_blocks = (K + BLOCK_K - 1) / BLOCK_K;

    // Fused scatter: rank r walks the owners in the order r+1, r+2, ..., r (ring schedule, map_m), so at any moment every owner
    // receives from ONE peer instead of all p-1 at once (incast would serialise the job on one GPU's NVLink ingress), and
    // the rank's own rows -- which need no NVLink -- come last, when the links are draining.
    if constexpr (kScatter) {
        // Programmatic dependent launch: the tile reducer (next kernel in this stream) may be scheduled once EVERY CTA of
        // this grid is running.  It spins on tiles this grid (and the peers') produce, so it must never take an SM's
        // registers before the GEMM CTA of that SM is resident.
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_c) : "memory");
    }
    if (warp == 1 && lane == 0) {
        for (int i = 0; i < kStages; ++i) { mbar_init(smem_u32(full_bar + i), 1); mbar_init(smem_u32(empty_bar + i), 1); }
        for (int i = 0; i < kAccStages; ++i) { mbar_init(smem_u32(tmem_full + i), 1); mbar_init(smem_u32(tmem_empty + i), kEpiThreads); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 2) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "r"(512) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ================= TMA producer =================
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                const TileCoord tc = tile_of<kMode>(t, m_blocks, n_blocks, sp);
                const CUtensorMap* amap = &map_a;
                int a_row = tc.m * BLOCK_M;
                if constexpr (kGather) {
                    const int slot = tc.m / sp.blocks_per_rank, c = tc.m - slot * sp.blocks_per_rank;
                    if (slot == sp.me) {
                        amap = &sp.a_local; a_row = c * BLOCK_M;          // own rows: straight from the local shard
                    } else {
                        // the owner's push kernel counts this 128-row block in once per pushing CTA
                        const uint32_t* f = sp.flags + slot * sp.blocks_per_rank + c;
                        unsigned long long t0 = 0; unsigned spins = 0;
                        while (true) {
                            uint32_t v;
                            asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
                            if (v >= sp.target) break;
                            if ((++spins & 0x3ff) == 0) {
                                unsigned long long now = gtimer();
                                if (t0 == 0) t0 = now;
                                else if (now - t0 > sp.timeout_ns) {
                                    if (atomicCAS(sp.err + 1, 0, 4) == 0) { sp.err[2] = (int)blockIdx.x; sp.err[3] = tc.m; sp.err[4] = (int)v; sp.err[5] = (int)sp.target; sp.err[6] = slot; }
                                    *sp.err = BG_ETIMEOUT; __threadfence_system(); __trap();
                                }
                            }
                        }
                        asm volatile("fence.proxy.async.global;" ::: "memory");   // the peer's stores (generic proxy) before my TMA reads (async proxy)
                    }
                }
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(smem_u32(empty_bar + stage), phase ^ 1);
                    const uint32_t bar = smem_u32(full_bar + stage);
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes), sb = sa + kABytes;
                    mbar_expect_tx(bar, kStageBytes);
                    if (kAMn) {  // A stored [K][M]: two boxes of 64 M-elements x 64 K-rows
#pragma unroll
                        for (int j = 0; j < BLOCK_M / 64; ++j) tma_load_2d(sa + j * (BLOCK_K * 128), &map_a, bar, tc.m * BLOCK_M + j * 64, kb * BLOCK_K);
                    } else {     // A stored [M][K]: one box of 64 K-elements x 128 rows
                        tma_load_2d(sa, amap, bar, kb * BLOCK_K, a_row);
                    }
                    if (kBMn) {
#pragma unroll
                        for (int j = 0; j < BLOCK_N / 64; ++j) tma_load_2d(sb + j * (BLOCK_K * 128), &map_b, bar, tc.n * BLOCK_N + j * 64, kb * BLOCK_K);
                    } else {
                        tma_load_2d(sb, &map_b, bar, kb * BLOCK_K, tc.n * BLOCK_N);
                    }
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        if (elect_one()) {
            constexpr uint32_t idesc = make_idesc(kAMn, kBMn);
            // K-major: 8-row groups 1024 B apart (SBO), K step 32 B inside the 128-B swizzle row
            // MN-major: 64-element MN chunks BLOCK_K*128 B apart (LBO), 8-k-row groups 1024 B apart (SBO), K step 16 rows
            constexpr uint32_t a_lbo = kAMn ? BLOCK_K * 128 : 0, b_lbo = kBMn ? BLOCK_K * 128 : 0;
            constexpr uint32_t a_kstep = kAMn ? UMMA_K * 128 : UMMA_K * 2, b_kstep = kBMn ? UMMA_K * 128 : UMMA_K * 2;
            int stage = 0; uint32_t phase = 0;
            int acc = 0; uint32_t acc_phase = 0;
            for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
                mbar_wait(smem_u32(tmem_empty + acc), acc_phase ^ 1);
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(smem_u32(full_bar + stage), phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + stage * kStageBytes), sb = sa + kABytes;
#pragma unroll
                    for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
                        const uint64_t da = make_smem_desc(sa + k * a_kstep, a_lbo, 1024);
                        const uint64_t db = make_smem_desc(sb + k * b_kstep, b_lbo, 1024);
                        umma_bf16(d_tmem, da, db, idesc, (kb | k) != 0);
                    }
                    umma_commit(smem_u32(empty_bar + stage));  // frees the smem stage when these MMAs retire
                    if (++stage == kStages) { stage = 0; phase ^= 1; }
                }
                umma_commit(smem_u32(tmem_full + acc));          // accumulator complete -> epilogue
                if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
            }
        }
    } else if (warp >= 4) {
        // ================= epilogue: TMEM -> registers -> (+C) -> bf16 -> swizzled smem -> TMA store =================
        const int ew = warp - 4;                      // == warp % 4: TMEM lanes [32*ew, 32*ew+32)
        const int row = ew * 32 + lane;               // row inside the tile
        const bool issuer = threadIdx.x == 4 * 32;
        int acc = 0; uint32_t acc_phase = 0;
        int buf = 0;
        uint32_t* prev_flag = nullptr;                // fused scatter: arrival counter of the tile whose stores are in flight
        for (int t = blockIdx.x; t < num_tiles; t += gridDim.x) {
            const TileCoord tc = tile_of<kMode>(t, m_blocks, n_blocks, sp);
            mbar_wait(smem_u32(tmem_full + acc), acc_phase);
            tc_fence_after();
#pragma unroll 1
            for (int c = 0; c < BLOCK_N / kStoreCols; ++c) {
                const int n0 = tc.n * BLOCK_N + c * kStoreCols;
                if (n0 >= N) break;  // whole chunk out of bounds (uniform across the CTA)
                uint32_t v[64];
                const uint32_t taddr = tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(acc * BLOCK_N + c * kStoreCols);
                tmem_ld32(taddr, v);
                tmem_ld32(taddr + 32, v + 32);
                tmem_ld_wait();
                if (c == BLOCK_N / kStoreCols - 1 || n0 + kStoreCols >= N) {
                    // last TMEM read of this tile: hand the accumulator back to the MMA warp
                    tc_fence_before();
                    mbar_arrive(smem_u32(tmem_empty + acc));
                }
                if (accumulate) {
                    const long long grow = (long long)tc.m * BLOCK_M + row;
                    if (grow < M) {
#pragma unroll
                        for (int j = 0; j < 8; ++j) {
                            if (n0 + j * 8 < N) {
                                float o[8];
                                unpack8(*reinterpret_cast<const uint4*>(c_old + grow * N + n0 + j * 8), o);
#pragma unroll
                                for (int e = 0; e < 8; ++e) v[j * 8 + e] = __float_as_uint(__uint_as_float(v[j * 8 + e]) + o[e]);
                            }
                        }
                    }
                }
                // staging buffer `buf` must be free: the store issued two chunks ago has finished reading it
                if (issuer) tma_store_wait_read<kNumStoreBufs - 1>();
                epi_bar_sync();
                uint8_t* sbuf = smem_store + buf * kStoreBytes;
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    uint4 pk;
                    pk.x = f2_to_bf2(__uint_as_float(v[j * 8 + 0]), __uint_as_float(v[j * 8 + 1]));
                    pk.y = f2_to_bf2(__uint_as_float(v[j * 8 + 2]), __uint_as_float(v[j * 8 + 3]));
                    pk.z = f2_to_bf2(__uint_as_float(v[j * 8 + 4]), __uint_as_float(v[j * 8 + 5]));
                    pk.w = f2_to_bf2(__uint_as_float(v[j * 8 + 6]), __uint_as_float(v[j * 8 + 7]));
                    // 128-B swizzle: 16-B chunk j of row r lives at chunk (j ^ (r & 7))
                    *reinterpret_cast<uint4*>(sbuf + row * 128 + ((j ^ (row & 7)) << 4)) = pk;
                }
                fence_proxy_async();
                epi_bar_sync();
                if (issuer) {
                    if constexpr (kScatter) {
                        const int row0 = tc.m * BLOCK_M, owner = row0 / sp.rows_per_rank;
                        tma_store_2d(&sp.dst[owner], smem_u32(sbuf), n0, sp.me * sp.rows_per_rank + (row0 - owner * sp.rows_per_rank));
                    } else {
                        tma_store_2d(&map_c, smem_u32(sbuf), n0, tc.m * BLOCK_M);
                    }
                    tma_store_commit();
                }
                buf ^= 1;
            }
            if constexpr (kScatter) {
                if (issuer) {
                    // Publish the PREVIOUS tile: every bulk group except this tile's (<= 4 chunks) has completed, so its peer
                    // stores are done -- no stall on this tile's NVLink latency.
                    if (prev_flag != nullptr) {
                        asm volatile("cp.async.bulk.wait_group %0;" ::"n"(BLOCK_N / kStoreCols) : "memory");
                        __threadfence_system();
                        asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(prev_flag) : "memory");
                    }
                    const int row0 = tc.m * BLOCK_M, owner = row0 / sp.rows_per_rank;
                    prev_flag = sp.flags[owner] + ((row0 - owner * sp.rows_per_rank) / BLOCK_M) * n_blocks + tc.n;
                }
            }
            if (++acc == kAccStages) { acc = 0; acc_phase ^= 1; }
        }
        if (issuer) {
            tma_store_wait_all<0>();
            if constexpr (kScatter) {
                if (prev_flag != nullptr) {
                    __threadfence_system();
                    asm volatile("red.release.sys.global.add.u32 [%0], 1;" ::"l"(prev_flag) : "memory");
                }
            }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512) : "memory");
    }
}

// ---- host: tensor maps --------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
    static EncodeTiledFn fn = nullptr;
    static std::once_flag once;
    std::call_once(once, [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess && q == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    });
    return fn;
}

// 2-D row-major bf16 tensor [rows][cols] (cols contiguous); box = box_cols x box_rows, 128-B swizzle
int make_map(CUtensorMap* map, const void* ptr, long long rows, long long cols, int box_cols, int box_rows) {
    EncodeTiledFn fn = encode_fn();
    if (!fn) return fail(BG_ECUDA, "cuTensorMapEncodeTiled entry point not available");
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)cols * 2};
    cuuint32_t box[2] = {(cuuint32_t)box_cols, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r == CUDA_ERROR_INVALID_CONTEXT) {
        // a thread (e.g. the autograd engine's) whose first CUDA call is this driver-API encode has no context bound yet:
        // touching the runtime binds the device's primary context, then retry
        cudaFree(0);
        r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
               CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
               CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    }
    if (r != CUDA_SUCCESS) return fail(BG_ECUDA, "cuTensorMapEncodeTiled failed (%d) rows=%lld cols=%lld", (int)r, rows, cols);
    return BG_OK;
}

int g_num_sms = 0;

}  // namespace

int bg_preload_gemm();

static int gemm_setup() {
    if (g_num_sms == 0) {
        int dev = 0;
        BG_CUDA(cudaGetDevice(&dev));
        int sms = 0;
        BG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
#define BG_SMEM(K) BG_CUDA(cudaFuncSetAttribute(K, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes))
        BG_SMEM(gemm_bf16_kernel<kTN>); BG_SMEM(gemm_bf16_kernel<kNN>); BG_SMEM(gemm_bf16_kernel<kNT>);
        BG_SMEM((gemm_bf16_kernel<kTN, kScatterMode>)); BG_SMEM((gemm_bf16_kernel<kNN, kScatterMode>)); BG_SMEM((gemm_bf16_kernel<kNT, kScatterMode>));
        BG_SMEM((gemm_bf16_kernel<kTN, kGatherMode>)); BG_SMEM((gemm_bf16_kernel<kNN, kGatherMode>));
#undef BG_SMEM
        g_num_sms = sms;
    }
    return BG_OK;
}



static int make_ab_maps(CUtensorMap* ma, CUtensorMap* mb, const void* a, const void* b, long long m, long long n, long long k, int layout) {
    // A: TN/NN stored [M][K] (K-major: box 64 K x 128 rows); NT stored [K][M] (MN-major: box 64 M x 64 K-rows)
    int rc = layout == kNT ? make_map(ma, a, k, m, 64, BLOCK_K) : make_map(ma, a, m, k, BLOCK_K, BLOCK_M);
    if (rc) return rc;
    // B: TN stored [N][K] (box 64 K x 256 rows); NN/NT stored [K][N] (box 64 N x 64 K-rows)
    return layout == kTN ? make_map(mb, b, n, k, BLOCK_K, BLOCK_N) : make_map(mb, b, k, n, 64, BLOCK_K);
}

static int gemm_launch(const void* a, const void* b, void* c, const void* addend, long long m, long long n, long long k, int layout, void* stream);

extern "C" int bg_gemm_bf16(const void* a, const void* b, void* c, long long m, long long n, long long k, int layout,
                            int accumulate, void* stream) {
    return gemm_launch(a, b, c, accumulate ? c : nullptr, m, n, k, layout, stream);
}

// C = A op B + addend: the residual add that follows a row-parallel projection (LlamaModel_tensor_parallel.py:83,100: out + residual)
// rides in the GEMM epilogue -- fp32 accumulator + bf16 addend, ONE rounding, no separate elementwise pass.
extern "C" int bg_gemm_bf16_add(const void* a, const void* b, void* c, const void* addend, long long m, long long n, long long k,
                                int layout, void* stream) {
    if (addend == nullptr || (uintptr_t)addend % 16) return fail(BG_EINVAL, "bg_gemm_bf16_add: addend must be a 16-B aligned [M][N] bf16 tensor");
    return gemm_launch(a, b, c, addend, m, n, k, layout, stream);
}

static int gemm_launch(const void* a, const void* b, void* c, const void* addend, long long m, long long n, long long k, int layout, void* stream) {
    const int accumulate = addend != nullptr;
    if (layout < 0 || layout > 2) return fail(BG_EINVAL, "bg_gemm_bf16: layout %d", layout);
    if (m <= 0 || n <= 0 || k <= 0 || m % 8 || n % 8 || k % 8)
        return fail(BG_EINVAL, "bg_gemm_bf16: m,n,k (%lld,%lld,%lld) must be positive multiples of 8", m, n, k);
    if (((uintptr_t)a | (uintptr_t)b | (uintptr_t)c) % 16) return fail(BG_EINVAL, "bg_gemm_bf16: pointers must be 16-B aligned");
    CUtensorMap ma, mb, mc;
    int rc = make_ab_maps(&ma, &mb, a, b, m, n, k, layout);
    if (rc) return rc;
    rc = make_map(&mc, c, m, n, kStoreCols, BLOCK_M);
    if (rc) return rc;
    rc = gemm_setup();
    if (rc) return rc;
    const long long tiles = ((m + BLOCK_M - 1) / BLOCK_M) * ((n + BLOCK_N - 1) / BLOCK_N);
    const int grid = (int)(tiles < g_num_sms ? tiles : g_num_sms);
    cudaStream_t st = (cudaStream_t)stream;
    const __nv_bfloat16* c_old = (const __nv_bfloat16*)addend;
    FuseParams<kPlain> none;
    if (layout == kTN) gemm_bf16_kernel<kTN><<<grid, kThreads, kSmemBytes, st>>>(ma, mb, mc, c_old, (int)m, (int)n, (int)k, accumulate, none);
    else if (layout == kNN) gemm_bf16_kernel<kNN><<<grid, kThreads, kSmemBytes, st>>>(ma, mb, mc, c_old, (int)m, (int)n, (int)k, accumulate, none);
    else gemm_bf16_kernel<kNT><<<grid, kThreads, kSmemBytes, st>>>(ma, mb, mc, c_old, (int)m, (int)n, (int)k, accumulate, none);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

// ---- fused GEMM + reduce-scatter / all-reduce ---------------------------------------------------------------------------------
namespace {

struct Bcast {
    char* out[BG_MAX_PEERS];     // every member's full [M][N] output (all-reduce), or all null (reduce-scatter)
    char* mc;                    // multicast address of that buffer (one multimem.st reaches every member), or null
    int on;
};

__device__ __forceinline__ void mm_st_16g(void* mc, const uint4& v) {
    asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(mc), "f"(__uint_as_float(v.x)),
                 "f"(__uint_as_float(v.y)), "f"(__uint_as_float(v.z)), "f"(__uint_as_float(v.w)) : "memory");
}

// One CTA per local tile (grid-strided): wait until all p partial tiles have landed, sum them in fp32, write bf16 -- into the
// local [M/p][N] result (reduce-scatter) or into rows [me*M/p, ...) of EVERY member's [M][N] result (all-reduce: the second
// shot of the two-shot algorithm happens here, tile by tile, while the GEMMs are still producing).
__global__ void __launch_bounds__(128, 8) tile_reduce_kernel(const __nv_bfloat16* __restrict__ partial, uint32_t* __restrict__ flags,
                                                          __nv_bfloat16* __restrict__ out, int p, int me, int rows_per_rank, int N,
                                                          int n_blocks, int local_tiles, const __grid_constant__ Bcast bc,
                                                          unsigned long long timeout_ns, int* err) {
    for (int lt = blockIdx.x; lt < local_tiles; lt += gridDim.x) {
        if (threadIdx.x == 0) {
            unsigned long long t0 = 0;
            unsigned spins = 0;
            while (true) {
                uint32_t v;
                asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(flags + lt) : "memory");
                if (v >= (uint32_t)p) break;
                if ((++spins & 0x3ff) == 0) {
                    unsigned long long now = gtimer();
                    if (t0 == 0) t0 = now;
                    else if (now - t0 > timeout_ns / 2) {     // (half: a missing tile is reported before the peers' barriers time out)
                        if (atomicCAS(err + 1, 0, 3) == 0) { err[2] = (int)blockIdx.x; err[3] = lt; err[4] = (int)v; err[5] = p; err[6] = local_tiles; }
                        *err = BG_ETIMEOUT; __threadfence_system(); __trap();
                    }
                }
            }
        }
        __syncthreads();
        const int mb = lt / n_blocks, nb = lt % n_blocks;
        const int row0 = mb * BLOCK_M, col0 = nb * BLOCK_N;
        const int cols = min(BLOCK_N, N - col0), rows = min(BLOCK_M, rows_per_rank - row0);
        const int vec_per_row = cols / 8;
        for (int i = threadIdx.x; i < rows * vec_per_row; i += blockDim.x) {
            const int r = i / vec_per_row, c = i - r * vec_per_row;
            const size_t off = (size_t)(row0 + r) * N + col0 + c * 8;
            uint4 in[BG_MAX_PEERS];
#pragma unroll
            for (int src = 0; src < BG_MAX_PEERS; ++src)
                if (src < p) in[src] = ld16_stream(partial + (size_t)src * rows_per_rank * N + off);
            float acc[8];
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[e] = 0.f;
#pragma unroll
            for (int src = 0; src < BG_MAX_PEERS; ++src)
                if (src < p) {
                    float f[8];
                    unpack8(in[src], f);
#pragma unroll
                    for (int e = 0; e < 8; ++e) acc[e] += f[e];
                }
            const uint4 o = pack8(acc);
            if (bc.on) {
                const size_t goff = ((size_t)me * rows_per_rank * N + off) * 2;
                if (bc.mc != nullptr) {
                    mm_st_16g(bc.mc + goff, o);
                } else {
                    for (int k = 0; k < p; ++k) {
                        int q = me + k; if (q >= p) q -= p;
                        st16(bc.out[q] + goff, o);
                    }
                }
            } else {
                st16(out + off, o);
            }
        }
        __syncthreads();
        if (threadIdx.x == 0) flags[lt] = 0;   // ready for the next use (peers only write again after the next entry barrier)
    }
    // (all-reduce: the cross-rank exit barrier is the next kernel of the stream, bg_coll.cu)
    // programmatic dependent of the GEMM: do not complete (and release the stream) before the GEMM grid has
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

}  // namespace

// Internal entry used by bg_coll.cu (which owns contexts, groups and peer pointers).  partial/flags: per-member pointers.
// Both kernels go to ONE stream: the GEMM, then the reducer as its programmatic dependent (it starts when all GEMM CTAs are
// resident, not when they finish; tiles are handed over through the arrival counters).
int bg_gemm_scatter_launch(const void* a, const void* b, long long m, long long n, long long k, int layout, int p, int me,
                           void* const* partial_ptrs, uint32_t* const* flag_ptrs, void* out, void* const* bcast_ptrs, char* bcast_mc,
                           unsigned long long timeout_ns, int* err_dev, cudaStream_t st) {
    if (layout < 0 || layout > 2) return fail(BG_EINVAL, "bg_gemm_reduce_scatter: layout %d", layout);
    if (m <= 0 || n <= 0 || k <= 0 || n % 8 || k % 8) return fail(BG_EINVAL, "bg_gemm_reduce_scatter: bad dims");
    if (m % ((long long)p * BLOCK_M)) return fail(BG_EINVAL, "bg_gemm_reduce_scatter: M=%lld must be a multiple of p*%d", m, BLOCK_M);
    const int rows_per_rank = (int)(m / p);
    CUtensorMap ma, mb;
    int rc = make_ab_maps(&ma, &mb, a, b, m, n, k, layout);
    if (rc) return rc;
    FuseParams<kScatterMode> sp;
    sp.p = p; sp.me = me; sp.rows_per_rank = rows_per_rank;
    for (int i = 0; i < BG_MAX_PEERS; ++i) {
        sp.flags[i] = i < p ? flag_ptrs[i] : nullptr;
        if (i < p) { rc = make_map(&sp.dst[i], partial_ptrs[i], (long long)p * rows_per_rank, n, kStoreCols, BLOCK_M); if (rc) return rc; }
        else sp.dst[i] = sp.dst[0];
    }
    rc = gemm_setup();
    if (rc) return rc;
    const int m_blocks = (int)((m + BLOCK_M - 1) / BLOCK_M), n_blocks = (int)((n + BLOCK_N - 1) / BLOCK_N);
    const long long tiles = (long long)m_blocks * n_blocks;
    const int grid = (int)(tiles < g_num_sms ? tiles : g_num_sms);
    const int local_tiles = (rows_per_rank / BLOCK_M) * n_blocks;
    const CUtensorMap& mc_unused = sp.dst[0];
    if (layout == kTN) gemm_bf16_kernel<kTN, kScatterMode><<<grid, kThreads, kSmemBytes, st>>>(ma, mb, mc_unused, nullptr, (int)m, (int)n, (int)k, 0, sp);
    else if (layout == kNN) gemm_bf16_kernel<kNN, kScatterMode><<<grid, kThreads, kSmemBytes, st>>>(ma, mb, mc_unused, nullptr, (int)m, (int)n, (int)k, 0, sp);
    else gemm_bf16_kernel<kNT, kScatterMode><<<grid, kThreads, kSmemBytes, st>>>(ma, mb, mc_unused, nullptr, (int)m, (int)n, (int)k, 0, sp);
    BG_CHECK_LAUNCH();
    // one reducer CTA fits beside a GEMM CTA (registers); more CTAs than SMs only queue
    const int rgrid = local_tiles < g_num_sms ? local_tiles : g_num_sms;
    Bcast bc = {};
    if (bcast_ptrs != nullptr) {
        bc.on = 1; bc.mc = bcast_mc;
        for (int i = 0; i < p; ++i) bc.out[i] = (char*)bcast_ptrs[i];
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)rgrid); cfg.blockDim = dim3(128);   // slim: fits beside the GEMM CTA and two more collectives cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    static const bool no_pdl = getenv("HGB_NO_PDL") != nullptr;      // debugging aid: launch the reducer as an ordinary kernel
    cfg.attrs = attr; cfg.numAttrs = no_pdl ? 0 : 1;
    BG_CUDA(cudaLaunchKernelEx(&cfg, tile_reduce_kernel, (const __nv_bfloat16*)partial_ptrs[me], (uint32_t*)flag_ptrs[me],
                               (__nv_bfloat16*)out, p, me, rows_per_rank, (int)n, n_blocks, local_tiles, bc, timeout_ns, err_dev));
    bg::g_launches.fetch_add(1, std::memory_order_relaxed);
    return BG_OK;
}

// Fused all-gather + GEMM: the consumer half (the push kernel is launched by bg_coll.cu on the communication stream).
int bg_gemm_gather_launch(const void* a_local, const void* a_staged, const void* b, void* c, long long m, long long n, long long k,
                          int layout, int p, int me, const uint32_t* flags, uint32_t target, unsigned long long timeout_ns, int* err_dev,
                          cudaStream_t st) {
    if (layout != kTN && layout != kNN) return fail(BG_EINVAL, "bg_all_gather_gemm: layout %d", layout);
    if (((uintptr_t)a_local | (uintptr_t)a_staged | (uintptr_t)b | (uintptr_t)c) % 16) return fail(BG_EINVAL, "bg_all_gather_gemm: pointers must be 16-B aligned");
    CUtensorMap ma, mb, mc;
    int rc = make_ab_maps(&ma, &mb, a_staged, b, m, n, k, layout);
    if (rc) return rc;
    rc = make_map(&mc, c, m, n, kStoreCols, BLOCK_M);
    if (rc) return rc;
    FuseParams<kGatherMode> gp;
    rc = make_map(&gp.a_local, a_local, m / p, k, BLOCK_K, BLOCK_M);
    if (rc) return rc;
    gp.flags = flags; gp.target = target; gp.p = p; gp.me = me; gp.blocks_per_rank = (int)(m / p / BLOCK_M);
    gp.timeout_ns = timeout_ns; gp.err = err_dev;
    rc = gemm_setup();
    if (rc) return rc;
    const long long tiles = (m / BLOCK_M) * ((n + BLOCK_N - 1) / BLOCK_N);
    const int grid = (int)(tiles < g_num_sms ? tiles : g_num_sms);
    if (layout == kTN) gemm_bf16_kernel<kTN, kGatherMode><<<grid, kThreads, kSmemBytes, st>>>(ma, mb, mc, nullptr, (int)m, (int)n, (int)k, 0, gp);
    else gemm_bf16_kernel<kNN, kGatherMode><<<grid, kThreads, kSmemBytes, st>>>(ma, mb, mc, nullptr, (int)m, (int)n, (int)k, 0, gp);
    BG_CHECK_LAUNCH();
    return BG_OK;
}

// loads every kernel of this file and sets the dynamic shared-memory limits (see bg_preload_coll in bg_coll.cu)
int bg_preload_gemm() {
    int rc = gemm_setup();
    if (rc) return rc;
    cudaFuncAttributes attr;
    BG_CUDA(cudaFuncGetAttributes(&attr, reinterpret_cast<const void*>(&tile_reduce_kernel)));
    return BG_OK;
}
