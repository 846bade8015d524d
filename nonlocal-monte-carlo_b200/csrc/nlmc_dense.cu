// nlmc_dense.cu -- K3: dense-J production path (config C3, Sherrington-Kirkpatrick) with the local-field
// contraction H = S . J on the 5th-generation tensor cores (tcgen05 + TMEM + TMA), hand-written.
//
// What it replaces.  The reference recomputes x = J.dot(m) + h for every single-spin attempt
// (NMC/nmc.py:86): O(N^2) per attempt on a dense J.  Here R replicas share one J, so the fields of a
// block of B consecutive sites for ALL replicas are one GEMM,
//     H_blk[R x B] = S[R x N] . J[N x B]          (S = spins as bf16 +-1, exact)
// and a sweep visits the blocks in order ("left-looking"): GEMM for the block from the CURRENT spins,
// then the B sites of the block are updated one after the other per replica, correcting the field of
// site k by the flips of earlier sites of the same block (J_bb, B x B).  Every field is therefore
// computed from the up-to-date configuration -- the same sequential-scan heat bath as the reference's
// sweep with a fixed visiting order -- and 2*R*N^2 flop per sweep run on the tensor pipe.
//
// Precision.  Spins are exact in bf16; J (normalised, |J| <= 1) is split into n_split bf16 pieces
// (J = J1 + J2 + J3 to ~2^-24 relative) that accumulate into the same fp32 TMEM tile, the A tile being
// loaded once per k-step.  Reported energies are computed in fp64 on CUDA cores (nlmc_energy_states).
//
// GEMM kernel (gemm_bf16_tn_kernel): C^T[n][m] (+)= sum_k A[m][k] * B_s[n][k], both operands K-major.
// 128x128 output tile per CTA, BLOCK_K = 64 (one 128-byte swizzle atom), 3-stage TMA->smem ring with
// full/empty mbarriers; warp 0 = TMA producer, warp 1 = TMEM allocator + single-thread tcgen05.mma
// issuer (tcgen05.commit releases the smem stage / signals the epilogue), warps 2-5 = epilogue
// (tcgen05.ld 32x32b.x32 -> registers -> coalesced transposed stores, or fp32 reductions for split-K).
#include <cuda.h>

#include <algorithm>
#include <cmath>
#include <cstdlib>

#include "nlmc_common.cuh"
#include "nlmc_exchange.cuh"

namespace nlmc {

constexpr int kBM = 128, kBN = 128, kBK = 64, kMaxSplit = 3;
constexpr int kTileBytes = kBM * kBK * 2;  // one 128 x 64 bf16 tile = 16 KiB (A and each B piece)
constexpr int kGemmThreads = 192;
// shared memory of the GEMM with kSt ring stages: 3 for the stand-alone contraction (192 KB), 2 inside the sweep, where a
// GEMM CTA must fit on an SM next to a CTA of the block-update kernel it overlaps with (128 KB + 87 KB)
__host__ __device__ constexpr size_t gemm_smem_bytes(int stages, int wide = 1) {   // wide = 2: 128 x 256 output tiles (B tiles of 256 rows)
    return (size_t)stages * (1 + kMaxSplit * wide) * kTileBytes + 1024 /*align*/ + 256 /*barriers*/;
}

// ---- PTX wrappers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t *bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra WAIT_DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "WAIT_DONE:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void *smem_dst, const CUtensorMap *map, uint64_t *bar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(smem_u32(smem_dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tcgen05_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tcgen05_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void umma_commit(uint64_t *bar) {  // arrives on bar when all prior MMAs of this thread are done
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem), "l"(a_desc), "l"(b_desc), "r"(idesc),
        "r"(accumulate) : "memory");
}
// K-major, 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (=1), version 1 (sm_100)
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFF) >> 4);        // start address      bits [0,14)
    d |= (uint64_t)1 << 16;                             // leading byte off.  bits [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                   // stride byte off.   bits [32,46)
    d |= (uint64_t)1 << 46;                             // version            bits [46,48)
    d |= (uint64_t)2 << 61;                             // SWIZZLE_128B       bits [61,64)
    return d;
}
// kind::f16 instruction descriptor: D = F32, A = B = BF16, both K-major, M x N
__host__ __device__ constexpr uint32_t make_idesc_bf16(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

struct GemmParams {
    int M, N, n_base, k_blocks_total, k_splits, n_split, ldc;  // C^T is [N][ldc], ldc >= M; columns start at n_base
    // the contraction runs over k_blocks_total blocks of 64 columns of K: block i is kb_offset + i, and blocks at or above
    // skip_begin are shifted up by skip_count (look-ahead: the columns of the block being updated are left out of the main
    // part and contracted afterwards)
    int kb_offset, skip_begin, skip_count;
    int m_tile_base;  // first 128-row tile of this launch (a chain of the sweep owns a range of replica tiles)
    int pdl;  // launched with programmatic stream serialization: the prologue overlaps the tail of the preceding kernel
    float *Ct;
    int accumulate;  // 1: red.add into C^T (split-K or accumulate), 0: plain store
};

// kWide = 2: 128 x 256 output tiles.  A 128 x 128 x 16 MMA reads 4 KB + 4 KB of operands from shared memory for 64 cycles of
// tensor work -- exactly the 128 B/clk of the shared-memory port --, a 128 x 256 x 16 one reads 4 + 8 KB for 128 cycles (96 B/clk):
// the wide tile is what lets the tensor pipe run ahead of its operand supply.  Used by the stand-alone contraction when the
// column count is a multiple of 256 (2 ring stages of 112 KB).
// kPair (experiment, off by default): launched as clusters of 2 CTAs along M.  The two CTAs need the same B tiles, so each loads
// half of them (one of the two 128-row boxes of every piece) and TMA MULTICASTS it into both shared memories: the L2 -> SM
// traffic of B (470 MB per C3 contraction at ~10 TB/s) halves.  A ring stage may be refilled once BOTH CTAs have multiplied it,
// so the MMA threads commit to the `empty` barrier of both CTAs (multicast commit, count 2).  Correct, but not faster: every SM
// still takes in 112 KB per k-block, and that is the bound (see launch_fields_part).
template <int kStages, int kWide = 1, bool kPair = false>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b0,
                    const __grid_constant__ CUtensorMap map_b1, const __grid_constant__ CUtensorMap map_b2,
                    GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    constexpr int kBNt = kBN * kWide;                                   // columns of the output tile
    constexpr size_t kStageBytes = (size_t)(1 + kMaxSplit * kWide) * kTileBytes;
    const int stage_bytes = (1 + p.n_split * kWide) * kTileBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)kStages * kStageBytes);
    uint64_t *full = bars, *empty = bars + kStages, *tmem_full = bars + 2 * kStages;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t pair_rank = 0;
    if (kPair) asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(pair_rank));
    const int m0 = (blockIdx.x + p.m_tile_base) * kBM, n0 = p.n_base + blockIdx.y * kBNt;
    const int kb_per = (p.k_blocks_total + p.k_splits - 1) / p.k_splits;
    const int kb_begin = blockIdx.z * kb_per, kb_end = min(p.k_blocks_total, kb_begin + kb_per);
    const int n_kb = kb_end - kb_begin;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b0) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, kPair ? 2 : 1); }
            mbar_init(tmem_full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(kBNt) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    if (kPair) {   // both CTAs have their barriers ready before either multicasts into the other
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (p.pdl) {
        // barriers, TMEM and tensor-map prefetch are done: let the next kernel of the chain (the block update) start ITS
        // prologue now, and wait here until the kernel before us (the previous block update: spins, cleared field rows)
        // has completed and flushed
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        asm volatile("griddepcontrol.wait;" ::: "memory");
    }

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            const CUtensorMap *maps_b[kMaxSplit] = {&map_b0, &map_b1, &map_b2};
            for (int i = 0; i < n_kb; ++i) {
                const int s = i % kStages;
                const uint32_t ph = (uint32_t)(i / kStages) & 1u;
                mbar_wait(empty + s, ph ^ 1u);
                uint8_t *st = smem + (size_t)s * kStageBytes;
                mbar_expect_tx(full + s, (uint32_t)stage_bytes);
                int kb = kb_begin + i;
                if (kb >= p.skip_begin) kb += p.skip_count;
                const int k0 = (p.kb_offset + kb) * kBK;
                tma_load_2d(st, &map_a, full + s, k0, m0);
                for (int q = 0; q < p.n_split; ++q)   // a piece of B: kWide boxes of 128 rows, back to back (rows 128 B apart)
                    for (int w = 0; w < kWide; ++w) {
                        uint8_t *dst = st + (size_t)(1 + q * kWide + w) * kTileBytes;
                        if (!kPair) {
                            tma_load_2d(dst, maps_b[q], full + s, k0, n0 + w * kBN);
                        } else if ((uint32_t)w == pair_rank) {   // this CTA's half, into both CTAs (same offsets, each one's own barrier)
                            asm volatile(
                                "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%3, %4}], [%2], %5;"
                                ::"r"(smem_u32(dst)), "l"(maps_b[q]), "r"(smem_u32(full + s)), "r"(k0), "r"(n0 + w * kBN), "h"((uint16_t)3) : "memory");
                        }
                    }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer (single thread) =====
            const uint32_t idesc = make_idesc_bf16(kBM, kBNt);
            for (int i = 0; i < n_kb; ++i) {
                const int s = i % kStages;
                const uint32_t ph = (uint32_t)(i / kStages) & 1u;
                mbar_wait(full + s, ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)s * kStageBytes);
                const uint64_t a_desc = make_smem_desc_sw128(a_addr);
                for (int q = 0; q < p.n_split; ++q) {
                    const uint64_t b_desc = make_smem_desc_sw128(a_addr + (uint32_t)((1 + q * kWide) * kTileBytes));
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)  // UMMA_K = 16 bf16 = 32 bytes inside the swizzle atom
                        umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                                  (uint32_t)((i | q | k) != 0));
                }
                // smem stage reusable once these MMAs have read it (kPair: told to both CTAs, either may write into it)
                if (kPair)
                    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                                 ::"r"(smem_u32(empty + s)), "h"((uint16_t)3) : "memory");
                else
                    umma_commit(empty + s);
            }
            umma_commit(tmem_full);      // accumulator complete
        }
    } else {  // ===== epilogue: warps 2..5 own TMEM lane quarters (warp % 4) =====
        const int quarter = warp & 3;
        const int row = quarter * 32 + lane;
        mbar_wait(tmem_full, 0u);
        tcgen05_fence_after();
        const int m = m0 + row;
#pragma unroll 1
        for (int c = 0; c < kBNt; c += 32) {
            uint32_t v[32];
            const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c;
            asm volatile(
                "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                  "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15]),
                  "=r"(v[16]), "=r"(v[17]), "=r"(v[18]), "=r"(v[19]), "=r"(v[20]), "=r"(v[21]), "=r"(v[22]), "=r"(v[23]),
                  "=r"(v[24]), "=r"(v[25]), "=r"(v[26]), "=r"(v[27]), "=r"(v[28]), "=r"(v[29]), "=r"(v[30]), "=r"(v[31])
                : "r"(taddr));
            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
            if (n_kb > 0 && m < p.M) {
#pragma unroll
                for (int j = 0; j < 32; ++j) {
                    const int n = n0 + c + j;
                    if (n < p.N) {
                        float *dst = p.Ct + (size_t)n * p.ldc + m;  // transposed: lanes of a warp hit consecutive m
                        if (p.accumulate) atomicAdd(dst, __uint_as_float(v[j]));
                        else *dst = __uint_as_float(v[j]);
                    }
                }
            }
        }
    }
    tcgen05_fence_before();
    __syncthreads();
    if (kPair) {   // neither CTA leaves while the other may still write into it or arrive on its barriers
        asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
        asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    }
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kBNt) : "memory");
    }
}

// ---- host side -----------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

// row-major bf16 matrix [rows][cols] (cols contiguous, padded to a multiple of 64), box = 64 x 128, 128B swizzle
static int make_map(CUtensorMap *map, const void *base, uint64_t rows, uint64_t cols, uint64_t ld_elems) {
    EncodeTiledFn fn = encode_tiled_fn();
    if (!fn) {
        set_error("cuTensorMapEncodeTiled is not available from the driver");
        return NLMC_ERR_CUDA;
    }
    const cuuint64_t dims[2] = {cols, rows};
    const cuuint64_t strides[1] = {ld_elems * 2};
    const cuuint32_t box[2] = {(cuuint32_t)kBK, (cuuint32_t)kBM};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void *>(base), dims, strides, box, estr,
                          CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                          CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed with %d", (int)r);
        return NLMC_ERR_CUDA;
    }
    return NLMC_OK;
}

static inline uint16_t f32_to_bf16_rn(float f) {
    uint32_t x;
    memcpy(&x, &f, 4);
    const uint32_t lsb = (x >> 16) & 1u;
    x += 0x7FFFu + lsb;
    return (uint16_t)(x >> 16);
}
static inline float bf16_to_f32(uint16_t b) {
    const uint32_t x = (uint32_t)b << 16;
    float f;
    memcpy(&f, &x, 4);
    return f;
}

}  // namespace nlmc

struct nlmc_dense {
    nlmc_instance *inst = nullptr;
    int n = 0, n_pad = 0, R = 0, R_pad = 0, n_split = 1, block = 128;
    int upd_rpc = 8;                   // replicas (warps) per CTA of the block-update kernel: one CTA per SM
    uint16_t *S = nullptr;             // [R_pad][n_pad] bf16 spins (+-1; 0 in the padding)
    uint16_t *Jp[3] = {nullptr, nullptr, nullptr};  // [n_pad][n_pad] bf16 pieces of J (row-major, J symmetric)
    float *Jf = nullptr;               // [n_pad][n_pad] fp32 J TRANSPOSED (JfT[a][b] = J[b][a]; in-block corrections)
    float *hf = nullptr;               // [n_pad]
    float *Ht = nullptr;               // [n_pad][R_pad] fields, transposed
    float *beta = nullptr;             // [R_pad]
    double *E = nullptr;               // [R_pad]
    double *bestE = nullptr;           // [R_pad] lowest energy seen since nlmc_dense_best_reset
    uint16_t *bestS = nullptr;         // [R_pad][n_pad] configuration of that energy
    uint8_t *modes = nullptr;          // [R_pad][n_pad] per-site NMC phase mode: 0 normal, 1 hot (beta/temp_x), 2 frozen
    bool modes_on = false;
    float temp_x = 1.0f;
    CUtensorMap map_S, map_J[3];
    unsigned long long seed = 0;
    uint32_t *d_sweep = nullptr;       // [1] sweep counter on the device (read by the kernels of the captured graph)
    cudaGraphExec_t sweep_graph = nullptr;  // one whole sweep: 1 memset + n/128 x (split-K GEMM, block update) + counter
    nlmc::LadderExchange xch;          // replica exchange by beta labels (nlmc_dense_ladders / nlmc_dense_exchange)
    cudaStream_t stream = nullptr;
    cudaStream_t stream2 = nullptr;     // look-ahead GEMMs of the sweep (forked from / joined into `stream` inside the captured graph)
    std::vector<cudaEvent_t> fork_ev;  // one event per fork / join point of the captured sweep
    std::vector<cudaStream_t> chain_streams;  // streams of the independent replica chains of the sweep (beyond `stream`)
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    unsigned long long *fused_prof = nullptr;   // NLMC_DENSE_FUSED_PROF: phase timestamps of the last fused sweep (printed at destroy)
};

namespace nlmc {

// fields of columns [col0, col0 + n_cols) for all replicas: Ht[col][r] = sum_k S[r][k] J[col][k], the sum running over
// the k-blocks [kb_first, kb_first + kb_count) of 64 columns with the blocks [skip_first, skip_first + skip_count) left out
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, bool pdl, Args... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, args...);
}

static int launch_fields_part(nlmc_dense *D, cudaStream_t st, int col0, int n_cols, int k_splits, bool clear, int kb_first,
                              int kb_count, int skip_first, int skip_count, int stages, bool pdl = false, int m_tile_base = 0,
                              int m_tiles = 0) {
    NLMC_REQUIRE(col0 % kBN == 0, "launch_fields: col0 must be a multiple of %d", kBN);
    GemmParams p;
    p.M = D->R_pad;
    p.N = std::min(D->n_pad, col0 + n_cols);
    p.n_base = col0;
    p.kb_offset = kb_first;
    p.skip_begin = skip_count > 0 ? skip_first - kb_first : (1 << 30);
    p.skip_count = skip_count;
    p.k_blocks_total = kb_count - skip_count;
    p.k_splits = std::max(1, std::min(k_splits, p.k_blocks_total));
    p.n_split = D->n_split;
    p.ldc = D->R_pad;
    p.Ct = D->Ht;
    p.accumulate = k_splits > 1 || !clear;
    p.pdl = pdl ? 1 : 0;
    p.m_tile_base = m_tile_base;
    if (m_tiles <= 0) m_tiles = D->R_pad / kBM - m_tile_base;
    if (p.accumulate && clear)
        NLMC_CUDA(cudaMemsetAsync(D->Ht + (size_t)col0 * D->R_pad, 0, sizeof(float) * (size_t)(p.N - col0) * D->R_pad, st));
    // stand-alone contraction over a multiple of 256 columns: 128 x 256 tiles (NLMC_DENSE_WIDE=0 keeps the 128 x 128 ones)
    static const bool wide_ok = [] { const char *e = getenv("NLMC_DENSE_WIDE"); return e ? atoi(e) != 0 : true; }();
    if (wide_ok && stages == 3 && !pdl && n_cols % (2 * kBN) == 0 && col0 % (2 * kBN) == 0 && p.N == col0 + n_cols) {
        const dim3 wgrid((unsigned)m_tiles, (unsigned)(n_cols / (2 * kBN)), (unsigned)p.k_splits);
        // NLMC_DENSE_PAIR=1: clusters of 2 CTAs along M share their B tiles by TMA multicast.  Measured at C3 size: 48.4 us against
        // 46.6 us without (29.2 / 31.0 us with one piece of J) -- halving the L2 traffic does not help, the fill rate of each SM's
        // shared memory (112 KB per k-block either way) is what bounds the wide kernel; off by default.
        static const bool pair_ok = [] { const char *e = getenv("NLMC_DENSE_PAIR"); return e ? atoi(e) != 0 : false; }();
        if (pair_ok && m_tiles % 2 == 0) {
            cudaLaunchConfig_t cfg = {};
            cfg.gridDim = wgrid; cfg.blockDim = dim3(kGemmThreads); cfg.dynamicSmemBytes = gemm_smem_bytes(2, 2); cfg.stream = st;
            cudaLaunchAttribute attr[1];
            attr[0].id = cudaLaunchAttributeClusterDimension;
            attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
            cfg.attrs = attr; cfg.numAttrs = 1;
            NLMC_CUDA(cudaLaunchKernelEx(&cfg, gemm_bf16_tn_kernel<2, 2, true>, D->map_S, D->map_J[0], D->map_J[D->n_split > 1 ? 1 : 0],
                                         D->map_J[D->n_split > 2 ? 2 : 0], p));
            return NLMC_OK;
        }
        gemm_bf16_tn_kernel<2, 2><<<wgrid, kGemmThreads, gemm_smem_bytes(2, 2), st>>>(
            D->map_S, D->map_J[0], D->map_J[D->n_split > 1 ? 1 : 0], D->map_J[D->n_split > 2 ? 2 : 0], p);
        NLMC_CUDA(cudaGetLastError());
        return NLMC_OK;
    }
    const dim3 grid((unsigned)m_tiles, (unsigned)((n_cols + kBN - 1) / kBN), (unsigned)p.k_splits);
    if (stages == 2)
        NLMC_CUDA(launch_pdl(gemm_bf16_tn_kernel<2>, grid, dim3(kGemmThreads), gemm_smem_bytes(2), st, pdl, D->map_S, D->map_J[0],
                             D->map_J[D->n_split > 1 ? 1 : 0], D->map_J[D->n_split > 2 ? 2 : 0], p));
    else
        gemm_bf16_tn_kernel<3><<<grid, kGemmThreads, gemm_smem_bytes(3), st>>>(
            D->map_S, D->map_J[0], D->map_J[D->n_split > 1 ? 1 : 0], D->map_J[D->n_split > 2 ? 2 : 0], p);
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
}

static int launch_fields(nlmc_dense *D, int col0, int n_cols, int k_splits, bool clear = true) {
    return launch_fields_part(D, D->stream, col0, n_cols, k_splits, clear, 0, D->n_pad / kBK, 0, 0, 3);
}

constexpr int kBlk = 128;  // sites per sequential block == kBN

struct PhiloxD {
    uint32_t k0, k1;
    __device__ __forceinline__ uint4 operator()(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3) const {
        uint32_t a = k0, b = k1;
#pragma unroll
        for (int i = 0; i < 10; ++i) {
            const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
            const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
            const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c1 ^ a;
            const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c3 ^ b;
            c1 = (uint32_t)p1; c3 = (uint32_t)p0; c0 = n0; c2 = n2;
            a += 0x9E3779B9u; b += 0xBB67AE85u;
        }
        return make_uint4(c0, c1, c2, c3);
    }
};

// Sequential heat-bath update of the sites [c0, c0+kBlk) for kRepPerCta replicas per CTA, one WARP per replica.
// Ht holds the fields of the block from the current spins (GEMM); hf the external fields.
// s_k <- +1 with probability 1/(1+exp(-2 beta_r field_k))  == sign(tanh(beta x) - 2u + 1), NMC/nmc.py:87
//
// The chain over k is sequential per replica and there are only R chains, so what matters is the latency of one
// step.  Round 1 (8 lanes per replica, Philox + logit inside the chain): 20.5 us per block, 6 % of the warp slots, three
// quarters of the C3 sweep.  Now:
//   0. everything that does not depend on the chain is done up front, in parallel over the block: the thresholds
//      theta_k = logit(u_k) / (2 beta) of all 128 sites (one Philox call and four logits per lane), with the NMC phase
//      modes folded in (hot sites: theta * temp_x; frozen sites: theta = -+inf so that the spin keeps its value);
//   1. the block is processed in sub-blocks of 8 sites: the 8 fields, thresholds and 28 in-sub-block couplings come in
//      with 128-bit broadcast loads and the 8 decisions are taken one after the other in registers
//      (compare, select, and the right-looking FMAs into the later fields of the sub-block);
//   2. the 8 flips are propagated to the later sites of the block by all 32 lanes (lane t owns k' = 8m + t + 32 q,
//      conflict-free reads of the transposed J_bb, two FMA chains of 4).
// The random stream is the one of round 1 (same Philox keys, same logit), so trajectories are the same up to the
// rounding of the propagation sums.
// CTA size: one warp per replica, as many replicas per CTA as it takes to give every SM ONE CTA (ncu on the 8-warp
// version: 256 CTAs on 148 SMs left the SMs with two CTAs as the critical path); the 64 KB coupling block is staged
// once per SM.  The sub-block loop is fully unrolled (16 x 8 sites, every shared-memory address an immediate); sites
// past the end of J are frozen through their threshold, so the trip count is a constant.
constexpr int kMaxRepPerCta = 16;
constexpr int kFldStride = kBlk + 8;    // floats per replica row (16-byte aligned rows, rows of a CTA in different banks)
__host__ __device__ constexpr size_t update_smem_bytes(int rep_per_cta) {
    return sizeof(float) * ((size_t)kBlk * kBlk + 3 * (size_t)rep_per_cta * kFldStride);
}

// Thresholds theta_k = logit(u_k) / (2 beta) and old spins of the block's 128 sites for one replica (one warp): lane ->
// sites 4*lane .. 4*lane+3, one Philox call and four logits per lane; NMC phase modes folded in.
__device__ __forceinline__ void block_thresholds(const uint16_t *S, const uint8_t *__restrict__ modes, int n_pad, int r, int c0,
                                                 int k_end, int lane, float beta_r, uint32_t seed_lo, uint32_t seed_hi,
                                                 uint32_t sweep, float temp_x, float *trow, float *srow) {
    {
        // spins and thresholds of the lane's four sites 4*lane .. 4*lane+3
        const uint2 sv = *reinterpret_cast<const uint2 *>(S + (size_t)r * n_pad + c0 + 4 * lane);   // 256 B per warp
        const uint16_t s16[4] = {(uint16_t)(sv.x & 0xffffu), (uint16_t)(sv.x >> 16), (uint16_t)(sv.y & 0xffffu), (uint16_t)(sv.y >> 16)};
        const float inv2b = 0.5f / beta_r;
        const PhiloxD rng{seed_lo, seed_hi ^ 0x44454e53u};
        // the stream of round 1: call (r, c0 + 8*(lane/2), sweep, lane & 1) serves the four sites 4*lane .. 4*lane+3
        const uint4 rv = rng((uint32_t)r, (uint32_t)(c0 + 8 * (lane >> 1)), sweep, (uint32_t)(lane & 1));
        const uint32_t ub[4] = {rv.x, rv.y, rv.z, rv.w};
        uint32_t md4 = 0u;
        if (modes != nullptr) md4 = *reinterpret_cast<const uint32_t *>(modes + (size_t)r * n_pad + c0 + 4 * lane);
        float th[4], so[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            // logit of a uniform in (0,1) from all 32 bits, evaluated on the side of the nearer tail so that both tails
            // are symmetric and reach 2^-33 (a 24-bit uniform rounded to 1.0 forced the spin down once in 2^24 draws)
            const bool hi = (ub[i] >> 31) != 0u;
            const uint32_t m = hi ? ~ub[i] : ub[i];
            const float v = ((float)m + 0.5f) * (1.0f / 4294967296.0f);       // in (0, 1/2]
            const float t = __logf(v) - __logf(1.0f - v);
            th[i] = (hi ? -t : t) * inv2b;
            so[i] = (s16[i] == 0) ? 0.0f : ((s16[i] & 0x8000u) ? -1.0f : 1.0f);
            const uint32_t md = (md4 >> (8 * i)) & 0xffu;
            if (md == 1u) th[i] *= temp_x;                                      // hot backbone: beta / temp_x
            if (md == 2u || 4 * lane + i >= k_end)                              // frozen (or padding): the spin keeps its value
                th[i] = so[i] > 0.0f ? -__int_as_float(0x7f800000) : __int_as_float(0x7f800000);
        }
        *reinterpret_cast<float4 *>(trow + 4 * lane) = make_float4(th[0], th[1], th[2], th[3]);
        *reinterpret_cast<float4 *>(srow + 4 * lane) = make_float4(so[0], so[1], so[2], so[3]);
    }
}

// The sequential chain over the 128 sites of a block for one replica (one warp), in sub-blocks of 8 (see the kernel below).
__device__ __forceinline__ void block_update_chain(const float *Jt, float *frow, float *srow, const float *trow, int lane) {
#pragma unroll
    for (int sb = 0; sb < kBlk / 8; ++sb) {
        const int base = sb * 8;
        // ---- 1. the sub-block, sequentially, in registers (identical in all lanes of the warp)
        float F[8], T[8], so[8], d[8];
        {
            const float4 f0 = *reinterpret_cast<const float4 *>(frow + base), f1 = *reinterpret_cast<const float4 *>(frow + base + 4);
            const float4 t0 = *reinterpret_cast<const float4 *>(trow + base), t1 = *reinterpret_cast<const float4 *>(trow + base + 4);
            const float4 s0 = *reinterpret_cast<const float4 *>(srow + base), s1 = *reinterpret_cast<const float4 *>(srow + base + 4);
            F[0] = f0.x; F[1] = f0.y; F[2] = f0.z; F[3] = f0.w; F[4] = f1.x; F[5] = f1.y; F[6] = f1.z; F[7] = f1.w;
            T[0] = t0.x; T[1] = t0.y; T[2] = t0.z; T[3] = t0.w; T[4] = t1.x; T[5] = t1.y; T[6] = t1.z; T[7] = t1.w;
            so[0] = s0.x; so[1] = s0.y; so[2] = s0.z; so[3] = s0.w; so[4] = s1.x; so[5] = s1.y; so[6] = s1.z; so[7] = s1.w;
        }
        float sn[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            sn[i] = F[i] > T[i] ? 1.0f : -1.0f;
            d[i] = sn[i] - so[i];
            if (i < 7) {   // right-looking inside the sub-block: row base+i of the transposed block, columns base+i+1 .. base+7
                const float4 lo = *reinterpret_cast<const float4 *>(Jt + (base + i) * kBlk + base);
                const float4 hi4 = *reinterpret_cast<const float4 *>(Jt + (base + i) * kBlk + base + 4);
                const float Ji[8] = {lo.x, lo.y, lo.z, lo.w, hi4.x, hi4.y, hi4.z, hi4.w};
#pragma unroll
                for (int i2 = i + 1; i2 < 8; ++i2) F[i2] = fmaf(Ji[i2], d[i], F[i2]);
            }
        }
        if (lane == 0) {
            *reinterpret_cast<float4 *>(srow + base) = make_float4(sn[0], sn[1], sn[2], sn[3]);
            *reinterpret_cast<float4 *>(srow + base + 4) = make_float4(sn[4], sn[5], sn[6], sn[7]);
        }
        // ---- 2. propagate the 8 flips to the later sites of the block: lane t owns k' = base + 8 + t + 32 q
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int kp = base + 8 + lane + 32 * q;
            if (base + 8 + 32 * q < kBlk && kp < kBlk) {   // two independent FMA chains: half the dependent latency
                float a = frow[kp], b = 0.0f;
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    a = fmaf(Jt[(base + j) * kBlk + kp], d[j], a);
                    b = fmaf(Jt[(base + 4 + j) * kBlk + kp], d[4 + j], b);
                }
                frow[kp] = a + b;
            }
        }
        __syncwarp();
    }
}

__device__ __forceinline__ void store_block_spins(uint16_t *S, const float *srow, int n_pad, int r, int c0, int k_end, int lane) {
    // write the replica's block segment back as bf16 (+1 = 0x3F80, -1 = 0xBF80): lane -> 4 consecutive sites, 256 B per warp
    {
        const float4 c = *reinterpret_cast<const float4 *>(srow + 4 * lane);
        const float sv[4] = {c.x, c.y, c.z, c.w};
        uint32_t w[2] = {0u, 0u};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = 4 * lane + i;
            const uint32_t b16 = (k < k_end) ? (sv[i] > 0.0f ? 0x3F80u : (sv[i] < 0.0f ? 0xBF80u : 0u)) : 0u;
            w[i >> 1] |= b16 << (16 * (i & 1));
        }
        *reinterpret_cast<uint2 *>(S + (size_t)r * n_pad + c0 + 4 * lane) = make_uint2(w[0], w[1]);
    }
}

__global__ void __launch_bounds__(kMaxRepPerCta * 32) dense_block_update_kernel(int n, int n_pad, int R_pad, int c0, const float *__restrict__ Ht,
                                                                               const float *__restrict__ Jf, const float *__restrict__ hf,
                                                                               const float *__restrict__ beta, uint16_t *S,
                                                                               uint32_t seed_lo, uint32_t seed_hi,
                                                                               const uint32_t *__restrict__ sweep_ptr, float *Ht_zero,
                                                                               const uint8_t *__restrict__ modes, float temp_x, int pdl,
                                                                               int r_base, int r_end) {
    extern __shared__ __align__(16) uint8_t dsm[];
    const uint32_t sweep = *sweep_ptr;   // bumped by the kernel at the end of the previous sweep, several launches back
    const int rpc = blockDim.x >> 5;                                   // replicas (warps) per CTA
    float *Jt = reinterpret_cast<float *>(dsm);                        // [kBlk j][kBlk k] = J[c0+k][c0+j] (transposed)
    float *fld = Jt + (size_t)kBlk * kBlk;                             // [rpc][kFldStride]  running fields
    float *thr = fld + (size_t)rpc * kFldStride;                       // [rpc][kFldStride]  thresholds
    float *spn = thr + (size_t)rpc * kFldStride;                       // [rpc][kFldStride]  spins as floats
    const int tid = threadIdx.x;
    const int rep = tid >> 5, lane = tid & 31;   // replica within the CTA = warp
    const int r = r_base + blockIdx.x * rpc + rep;   // the launch covers the replicas [r_base, r_end)
    const int k_end = min(kBlk, n - c0);
    // stage J_bb transposed, Jt[j][k] = J[c0+k][c0+j], from the transposed copy of J kept in global memory
    // (JfT[a][b] = J[b][a]): asynchronous 16-byte copies, all 64 KB in flight while the thresholds are computed
    for (int i = tid; i < kBlk * kBlk / 4; i += blockDim.x) {
        const int j = i / (kBlk / 4), k4 = i % (kBlk / 4);
        const uint32_t dst = smem_u32(reinterpret_cast<float4 *>(Jt) + i);
        const float *src = Jf + (size_t)(c0 + j) * n_pad + c0 + k4 * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    float *frow = fld + rep * kFldStride;
    float *trow = thr + rep * kFldStride;
    float *srow = spn + rep * kFldStride;
    const bool active = r < r_end;
    if (active) {
        block_thresholds(S, modes, n_pad, r, c0, k_end, lane, beta[r], seed_lo, seed_hi, sweep, temp_x, trow, srow);
    }
    if (pdl) {
        // everything above (coupling block, spins of this block, thresholds) is independent of the field GEMM that runs
        // just before this kernel: with programmatic stream serialization it overlapped that GEMM.  The fields are not.
        asm volatile("griddepcontrol.wait;" ::: "memory");
        if (pdl > 1) asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    }
    if (active) {
        // fields of the replica: lane -> sites lane + 32 q (the warps of a CTA read the same 32-byte sectors of Ht)
        float hv[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) hv[q] = Ht[(size_t)(c0 + lane + 32 * q) * R_pad + r] + hf[c0 + lane + 32 * q];
#pragma unroll
        for (int q = 0; q < 4; ++q) frow[lane + 32 * q] = hv[q];
        // the next block's split-K GEMM accumulates with atomics: clear its field rows for this replica
        if (Ht_zero != nullptr)
#pragma unroll
            for (int q = 0; q < 4; ++q) Ht_zero[(size_t)(lane + 32 * q) * R_pad + r] = 0.f;
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    if (!active) return;
    block_update_chain(Jt, frow, srow, trow, lane);
    store_block_spins(S, srow, n_pad, r, c0, k_end, lane);
}

// ---- the whole sweep as ONE kernel: clusters of 4 CTAs, split-K over the cluster, reduction through distributed shared memory
//
// Replicas are independent, so a tile of 128 replicas never has to wait for another tile: a CLUSTER of 4 CTAs owns one
// tile for the whole sweep and synchronises only with itself (two cluster barriers per block of sites; no launch
// boundaries, no global atomics, no field matrix in global memory).  33 clusters of 4 are co-resident on a B200 (15 of 8:
// tools/micro/cluster_occupancy.cu), and 4 x 32 replicas gives every CTA one warp's worth of replicas.  Per block b of 128 sites:
//   G  the contraction, in two parts.  (i) Columns outside block b-1: CTA c takes the k-blocks c, c + 4, ... for all 128
//      replicas of the tile -- TMA -> 2-stage smem ring -> tcgen05.mma into TMEM accumulator 0 (warp 0: producer, warp 1:
//      issuer).  None of it depends on the update of block b-1, so it is loaded and multiplied WHILE block b-1 is being
//      updated.  (ii) The two k-blocks of block b-1 itself, for the CTA's OWN 32 replicas: the update threads write the new
//      spins (bf16, 128-byte swizzle) straight into the spin slot of the ring stage that already holds the coupling tiles,
//      and the product goes to accumulator 1 (rows of other replicas hold whatever was in the slot: rows are independent
//      and never read).  Nothing the update produces travels through global memory and TMA before the fields are complete.
//   R  the update warps read the accumulators (tcgen05.ld: the warps of lane quarter k hold the 32 replicas of CTA k) and
//      PUSH the partial fields into the owner's receive buffer [source CTA][site][replica] with st.shared::cluster
//      (128-byte runs; the own quarter adds accumulator 1 and stores locally); cluster barrier; every update thread adds
//      the 4 partial values and h for its sites.
//   U  the sequential chain over the 128 sites.  An update thread owns ONE sub-block of 8 sites for kE replicas (fields in
//      registers, packed in pairs of sites); a warp holds kE sub-blocks x 32 / kE replica groups.  Step s = 0..15: the
//      owners of sub-block s take its 8 decisions one after the other with the right-looking corrections inside the
//      sub-block and publish the flips d[s][8][32] (mbarrier per sub-block); every thread whose sub-block comes later
//      applies them (couplings as quarter-warp broadcast loads from the staged J_bb, FFMA2).  Same thresholds, same random
//      stream and the same order of the floating-point corrections as dense_block_update_kernel: given equal fields the
//      two paths take identical decisions (tests/test_gpu_dense_tcgen05.py::test_cluster_sweep_equals_launch_chain).
//      The thresholds of block b+1 (Philox + logits) are computed in the shadow of the chain.
// The new spins also go to global memory (generic stores, fence.proxy.async, cluster barrier U) for the TMA loads of later
// blocks.  The 64 KB receive buffer and the staged coupling block J_bb share one region: pushes for block b+1 can only
// happen after every CTA of the cluster has finished the update of block b.
// Measured at C3 size (N = 2000, 2048 replicas = 16 clusters on 64 SMs): 0.2175 ms per sweep against 0.353 ms for the chain of
// launches; per block 13.0 us = chain 8.4 + drain / push / barrier 2.5 + partial sums 0.4 + staging J_bb 0.6 + stores 0.7
// (NLMC_DENSE_FUSED_PROF=1 prints these; profiles/r2d_dense_fused_summary.md).
constexpr int kFusedCluster = 4;
constexpr int kFusedRep = kBM / kFusedCluster;             // 32 replicas per CTA = the lanes of a warp
constexpr int kFusedSub = kBlk / 8;                        // 16 sub-blocks of 8 sites
#ifndef NLMC_FUSED_H
#define NLMC_FUSED_H 1
#endif
// sub-blocks per update thread (1 or 2): its tile is 8 kH sites x kE replicas.  Measured at C3 size: kH = 1, kE = 2 (8 update
// warps) 0.234 ms per sweep; kH = 2, kE = 2 (4 warps, the flips of a thread's first sub-block enter its second without a
// hand-off) 0.244 ms -- the chain gets shorter (8.5 vs 9.1 us per block) but the 4 warps stage J_bb and store more slowly;
// kH = 1, kE = 4 (4 warps) 0.265 ms.
constexpr int kH = NLMC_FUSED_H;
constexpr int kTiles = kFusedSub / kH;                     // steps of the chain per block
static_assert(kH == 1 || kH == 2, "one or two sub-blocks of 8 sites per thread");
#ifndef NLMC_FUSED_E
#define NLMC_FUSED_E 2
#endif
constexpr int kE = NLMC_FUSED_E;                           // replicas per update thread (2 or 4): a thread's tile is 8 sites x kE replicas
constexpr int kLanesPerSub = 32 / kE;                      // lanes that share a tile of sites; a warp holds kE tiles
constexpr int kFusedUpdWarps = kTiles / kE;
static_assert(kFusedUpdWarps % 4 == 0, "the update warps cover the four TMEM lane quarters evenly");
static_assert(kE == 2 || kE == 4, "thread tile of 8 sites x 2 or 4 replicas");
constexpr int kFusedUpdThreads = 32 * kFusedUpdWarps;      // 128
constexpr int kFusedThreads = 64 + kFusedUpdThreads;       // + producer warp + MMA warp
constexpr int kFusedStages = 2;
constexpr size_t kFusedRingBytes = (size_t)kFusedStages * (1 + kMaxSplit) * kTileBytes;     // 128 KB
constexpr size_t kFusedXBytes = sizeof(float) * kBlk * kBlk;                                // 64 KB: receive buffer / J_bb
constexpr size_t kFusedDBytes = sizeof(float) * kBlk * kFusedRep;                           // 16 KB: flips d[site][replica]
constexpr size_t kFusedSmemBytes = 1024 + kFusedRingBytes + kFusedXBytes + kFusedDBytes + 512;
static_assert(kFusedRep == 32, "one replica per lane");
static_assert(kFusedCluster * kBlk * kFusedRep * sizeof(float) == kFusedXBytes, "receive buffer and J_bb share one region");

struct FusedParams {
    int n, n_pad, R_pad, n_blocks, kb_total, n_split;
    const float *Jf, *hf, *beta;
    uint16_t *S;
    uint32_t seed_lo, seed_hi;
    const uint32_t *sweep_ptr;
    const uint8_t *modes;
    float temp_x;
    unsigned long long *prof;   // NLMC_DENSE_FUSED_PROF: %globaltimer at the phase boundaries of CTA 0, [n_blocks][10]
};

__device__ __forceinline__ unsigned long long global_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define NLMC_FUSED_MARK(i) do { if (p.prof && blockIdx.x == 0 && threadIdx.x == kFusedThreads - 1) p.prof[b * 10 + (i)] = global_ns(); } while (0)

__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async;" ::: "memory"); }

// packed pairs of floats (FFMA2 / FMUL2 / FADD2 on sm_100: two fp32 operations per issue slot, each rounded like the scalar one)
typedef unsigned long long f32x2;
__device__ __forceinline__ f32x2 pack2(float lo, float hi) { f32x2 r; asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi)); return r; }
__device__ __forceinline__ float lo2(f32x2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return lo; }
__device__ __forceinline__ float hi2(f32x2 v) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); return hi; }
__device__ __forceinline__ f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) { f32x2 d; asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c)); return d; }
__device__ __forceinline__ f32x2 mul2(f32x2 a, f32x2 b) { f32x2 d; asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }
__device__ __forceinline__ f32x2 add2(f32x2 a, f32x2 b) { f32x2 d; asm("add.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(a), "l"(b)); return d; }

// kE consecutive floats of shared memory
__device__ __forceinline__ void ld_e(const float *ptr, float (&v)[kE]) {
    if constexpr (kE == 4) { const float4 t = *reinterpret_cast<const float4 *>(ptr); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
    else { const float2 t = *reinterpret_cast<const float2 *>(ptr); v[0] = t.x; v[1] = t.y; }
}
__device__ __forceinline__ void st_e(float *ptr, const float (&v)[kE]) {
    if constexpr (kE == 4) *reinterpret_cast<float4 *>(ptr) = make_float4(v[0], v[1], v[2], v[3]);
    else *reinterpret_cast<float2 *>(ptr) = make_float2(v[0], v[1]);
}

// thresholds and old spins of the sites c0 + 8w .. c0 + 8w + 7 of replica r: the two Philox calls that serve these sites
// in block_thresholds (lanes 2w and 2w+1 there)
__device__ __forceinline__ void site_thresholds8(const uint16_t *S, const uint8_t *__restrict__ modes, int n_pad, int r, int c0, int w,
                                                 int k_end, float beta_r, uint32_t seed_lo, uint32_t seed_hi, uint32_t sweep,
                                                 float temp_x, float (&T)[8], float (&so)[8]) {
    const uint4 sv = *reinterpret_cast<const uint4 *>(S + (size_t)r * n_pad + c0 + 8 * w);
    const uint32_t sw[4] = {sv.x, sv.y, sv.z, sv.w};
    uint2 md8 = make_uint2(0u, 0u);
    if (modes != nullptr) md8 = *reinterpret_cast<const uint2 *>(modes + (size_t)r * n_pad + c0 + 8 * w);
    const float inv2b = 0.5f / beta_r;
    const PhiloxD rng{seed_lo, seed_hi ^ 0x44454e53u};
#pragma unroll
    for (int half = 0; half < 2; ++half) {
        const uint4 rv = rng((uint32_t)r, (uint32_t)(c0 + 8 * w), sweep, (uint32_t)half);
        const uint32_t ub[4] = {rv.x, rv.y, rv.z, rv.w};
        const uint32_t md4 = half ? md8.y : md8.x;
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const int k = 4 * half + i;
            const uint32_t s16 = (sw[k >> 1] >> (16 * (k & 1))) & 0xffffu;
            const bool hi = (ub[i] >> 31) != 0u;
            const uint32_t m = hi ? ~ub[i] : ub[i];
            const float v = ((float)m + 0.5f) * (1.0f / 4294967296.0f);       // in (0, 1/2]
            const float t = __logf(v) - __logf(1.0f - v);
            float th = (hi ? -t : t) * inv2b;
            const float s = (s16 == 0u) ? 0.0f : ((s16 & 0x8000u) ? -1.0f : 1.0f);
            const uint32_t md = (md4 >> (8 * i)) & 0xffu;
            if (md == 1u) th *= temp_x;                                       // hot backbone: beta / temp_x
            if (md == 2u || 8 * w + k >= k_end)                               // frozen (or padding): the spin keeps its value
                th = s > 0.0f ? -__int_as_float(0x7f800000) : __int_as_float(0x7f800000);
            T[k] = th;
            so[k] = s;
        }
    }
}

#define NLMC_TMEM_LD32(x, taddr)                                                                                           \
    asm volatile(                                                                                                          \
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "                                                                          \
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "                                          \
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"                          \
        : "=r"(x[0]), "=r"(x[1]), "=r"(x[2]), "=r"(x[3]), "=r"(x[4]), "=r"(x[5]), "=r"(x[6]), "=r"(x[7]), "=r"(x[8]),      \
          "=r"(x[9]), "=r"(x[10]), "=r"(x[11]), "=r"(x[12]), "=r"(x[13]), "=r"(x[14]), "=r"(x[15]), "=r"(x[16]),           \
          "=r"(x[17]), "=r"(x[18]), "=r"(x[19]), "=r"(x[20]), "=r"(x[21]), "=r"(x[22]), "=r"(x[23]), "=r"(x[24]),          \
          "=r"(x[25]), "=r"(x[26]), "=r"(x[27]), "=r"(x[28]), "=r"(x[29]), "=r"(x[30]), "=r"(x[31])                        \
        : "r"(taddr))

__global__ void __cluster_dims__(kFusedCluster, 1, 1) __launch_bounds__(kFusedThreads, 1)
dense_fused_sweep_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b0,
                         const __grid_constant__ CUtensorMap map_b1, const __grid_constant__ CUtensorMap map_b2,
                         const FusedParams p) {
    extern __shared__ uint8_t smem_raw[];
    // 1024-byte alignment as an OFFSET from the shared array, so that the compiler still knows these are shared-memory
    // pointers (LDS / STS instead of generic loads and stores in the chain)
    uint8_t *smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
    float *X = reinterpret_cast<float *>(smem + kFusedRingBytes);        // receive buffer [4][128][32]  /  Jt [128][128]
    float *dbuf = X + (size_t)kBlk * kBlk;                               // flips d[site of the block][replica]
    uint64_t *bars = reinterpret_cast<uint64_t *>(dbuf + (size_t)kBlk * kFusedRep);
    uint64_t *full = bars, *empty = bars + kFusedStages, *tmem_full = bars + 2 * kFusedStages;
    uint64_t *adep = tmem_full + 1;                     // the new spins of a block are in the spin slots of the ring (one phase per block)
    uint64_t *slots_free = adep + 1;                    // the tiles before the two spin-slot stages of the next block have been multiplied
    uint64_t *sub_done = slots_free + 1;                // [16] one phase per block: the flips of sub-block s are published
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(sub_done + kFusedSub);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const int tile = blockIdx.x / kFusedCluster;
    const int m0 = tile * kBM;
    const int stage_bytes = (1 + p.n_split) * kTileBytes;
    // Work list of this CTA for block b.  (i) its share of the contraction over the columns OUTSIDE block b-1: the k-blocks
    // rank, rank + 4, ... without the two of block b-1; spins and couplings come by TMA and the partial sums of all 128
    // replicas go to accumulator 0.  (ii) b > 0: BOTH k-blocks of block b-1, for its OWN 32 replicas only: the update
    // threads write the new spins straight into the spin slot of the ring stage (the rows of the other replicas hold
    // whatever was there: rows are independent), the couplings come by TMA, the result goes to accumulator 1 and is added
    // to this CTA's own partial rows.  So nothing the update produces has to travel through global memory and TMA before
    // the next block's fields are complete.
    auto is_dep = [](int kb, int b) { return b > 0 && (kb >> 1) == b - 1; };   // kBlk / kBK == 2 k-blocks per site block
    auto n_indep = [&](int b) {
        int c = 0;
        for (int kb = (int)rank; kb < p.kb_total; kb += kFusedCluster) c += is_dep(kb, b) ? 0 : 1;
        return c;
    };

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b0) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kFusedStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
            mbar_init(tmem_full, 1);
            mbar_init(adep, kFusedUpdWarps);
            mbar_init(slots_free, 1);
            for (int i = 0; i < kFusedSub; ++i) mbar_init(sub_done + i, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(2 * kBN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;
    // every CTA of the cluster is running before anybody touches a neighbour's shared memory
    cluster_arrive();
    cluster_wait();

    if (warp == 0) {  // ===== TMA producer =====
        const CUtensorMap *maps_b[kMaxSplit] = {&map_b0, &map_b1, &map_b2};
        uint32_t it = 0;
        auto issue = [&](int kb, int b, bool with_spins) {
            const int s = (int)(it % kFusedStages);
            const uint32_t ph = (it / kFusedStages) & 1u;
            uint8_t *st = smem + (size_t)s * (1 + kMaxSplit) * kTileBytes;
            mbar_wait(empty + s, ph ^ 1u);
            mbar_expect_tx(full + s, (uint32_t)(with_spins ? stage_bytes : stage_bytes - kTileBytes));
            if (with_spins) tma_load_2d(st, &map_a, full + s, kb * kBK, m0);
            for (int q = 0; q < p.n_split; ++q) tma_load_2d(st + (1 + q) * kTileBytes, maps_b[q], full + s, kb * kBK, b * kBlk);
            ++it;
        };
        for (int b = 0; b < p.n_blocks; ++b) {
            if (b > 0) cluster_arrive();                       // U(b-1): nothing of ours to publish
            if (lane == 0) {
                fence_proxy_async();                           // spins stored by the update threads before the last barrier we passed
                for (int kb = (int)rank; kb < p.kb_total; kb += kFusedCluster) if (!is_dep(kb, b)) issue(kb, b, true);
                if (b > 0) { issue(2 * (b - 1), b, false); issue(2 * (b - 1) + 1, b, false); }   // couplings only: the spins come from the update threads
            }
            __syncwarp();
            if (b > 0) cluster_wait();
            cluster_arrive();                                  // R(b)
            cluster_wait();
        }
        cluster_arrive();                                      // U(last)
        cluster_wait();
    } else if (warp == 1) {  // ===== MMA issuer =====
        const uint32_t idesc = make_idesc_bf16(kBM, kBN);
        uint32_t it = 0;
        auto multiply = [&](uint32_t acc, bool first) {
            const int s = (int)(it % kFusedStages);
            const uint32_t ph = (it / kFusedStages) & 1u;
            mbar_wait(full + s, ph);
            tcgen05_fence_after();
            const uint32_t a_addr = smem_u32(smem + (size_t)s * (1 + kMaxSplit) * kTileBytes);
            const uint64_t a_desc = make_smem_desc_sw128(a_addr);
            for (int q = 0; q < p.n_split; ++q) {
                const uint64_t b_desc = make_smem_desc_sw128(a_addr + (1 + q) * kTileBytes);
#pragma unroll
                for (int k = 0; k < kBK / 16; ++k)
                    umma_bf16(acc, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc, (uint32_t)(!first || q != 0 || k != 0));
            }
            umma_commit(empty + s);
            ++it;
        };
        for (int b = 0; b < p.n_blocks; ++b) {
            if (b > 0) cluster_arrive();                       // U(b-1)
            if (lane == 0) {
                tcgen05_fence_after();
                const int ni = n_indep(b);
                for (int i = 0; i < ni; ++i) multiply(tmem_base, i == 0);
                if (b > 0) {
                    // Both ring stages are now free of spin tiles that are still to be read: the update threads of block b-1 may
                    // write the new spins into them.  (A parity wait on `empty` from their side could be a whole phase early.)
                    umma_commit(slots_free);
                    mbar_wait(adep, (uint32_t)((b - 1) & 1));  // the new spins of block b-1 are in the spin slots of the next two stages
                    tcgen05_fence_after();
                    multiply(tmem_base + (uint32_t)kBN, true);
                    multiply(tmem_base + (uint32_t)kBN, false);
                }
                if (ni > 0 || b > 0) umma_commit(tmem_full);
            }
            __syncwarp();
            if (b > 0) cluster_wait();                         // U(b-1)
            cluster_arrive();                                  // R(b)
            cluster_wait();
        }
        cluster_arrive();                                      // U(last)
        cluster_wait();
    } else {  // ===== update warps: thread = (tile kE v + lane / kLanesPerSub of 8 kH sites, replicas kE (lane % kLanesPerSub) .. + kE - 1); they also drain =====
        const int v = warp - 2;
        const int g = lane / kLanesPerSub, q = lane % kLanesPerSub;
        const int tb = kE * v + g;                             // the tile (sub-blocks kH tb .. kH tb + kH - 1 of every block) this thread owns
        const int ut = (int)threadIdx.x - 64;
        const int rq = m0 + (int)rank * kFusedRep + kE * q;    // its kE replicas
        float beta_e[kE];
#pragma unroll
        for (int e = 0; e < kE; ++e) beta_e[e] = p.beta[rq + e];
        const uint32_t sweep = *p.sweep_ptr;
        // thresholds of the thread's sites x kE replicas and their old spins as one bit each (a site past the end of J counts
        // as -1: its threshold is +inf, so it "stays" -1, its flip is 0 and its couplings are 0)
        float T[kH][8][kE];
        uint32_t so_up[kH];   // bit kE i + e: site i of replica e is +1
        auto thresholds = [&](int b, float (&Tn)[kH][8][kE], uint32_t (&up)[kH]) {
            const int c0 = b * kBlk;
            const int k_end = min(kBlk, p.n - c0);
#pragma unroll
            for (int h = 0; h < kH; ++h) {
                up[h] = 0u;
#pragma unroll
                for (int e = 0; e < kE; ++e) {
                    float t8[8], s8[8];
                    site_thresholds8(p.S, p.modes, p.n_pad, rq + e, c0, kH * tb + h, k_end, beta_e[e], p.seed_lo, p.seed_hi, sweep, p.temp_x, t8, s8);
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        Tn[h][i][e] = t8[i];
                        up[h] |= (s8[i] > 0.0f ? 1u : 0u) << (kE * i + e);
                    }
                }
            }
        };
        thresholds(0, T, so_up);
        uint32_t it_next = 0;      // ring iteration at which the NEXT block starts (the producer's count)
        uint32_t tf_phase = 0;     // completed phases of tmem_full
        for (int b = 0; b < p.n_blocks; ++b) {
            const int c0 = b * kBlk;
            const int k_end = min(kBlk, p.n - c0);
            const int ni = n_indep(b);
            it_next += (uint32_t)(ni + (b > 0 ? 2 : 0));
            NLMC_FUSED_MARK(0);
            NLMC_FUSED_MARK(1);
            if (b > 0) cluster_wait();                         // U(b-1): every CTA is done with its J_bb, the receive buffers are free
            NLMC_FUSED_MARK(2);
            {
                // TMEM lanes 32 k .. 32 k + 31 = the replicas of CTA k of the cluster: a warp (lane quarter warp % 4) pushes to ONE owner
                const int quarter = warp & 3;
                const uint32_t local = smem_u32(X) + (uint32_t)(((int)rank * kBlk) * kFusedRep + lane) * 4u;
                uint32_t dst;
                asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(dst) : "r"(local), "r"((uint32_t)quarter));
                if (ni > 0 || b > 0) {
                    mbar_wait(tmem_full, tf_phase & 1u);
                    ++tf_phase;
                    tcgen05_fence_after();
                }
                NLMC_FUSED_MARK(3);
                constexpr int kColsPerWarp = kBN / (kFusedUpdWarps / 4);   // the warps of a lane quarter share the 128 columns
#pragma unroll 1
                for (int c = (v >> 2) * kColsPerWarp; c < ((v >> 2) + 1) * kColsPerWarp; c += 32) {
                    uint32_t x[32];
                    const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)c;
                    if (ni > 0) {
                        NLMC_TMEM_LD32(x, taddr);
                        asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j) x[j] = 0u;
                    }
                    // site c + j of this replica: the 32 lanes of the warp write 128 contiguous bytes
                    if ((uint32_t)quarter == rank) {
                        if (b > 0) {   // this CTA's own replicas: add the part contracted from the spin slots (accumulator 1)
                            uint32_t y[32];
                            NLMC_TMEM_LD32(y, taddr + (uint32_t)kBN);
                            asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                            for (int j = 0; j < 32; ++j) x[j] = __float_as_uint(__uint_as_float(x[j]) + __uint_as_float(y[j]));
                        }
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(local + (uint32_t)((c + j) * kFusedRep * 4)), "r"(x[j]) : "memory");
                    } else {
#pragma unroll
                        for (int j = 0; j < 32; ++j)
                            asm volatile("st.shared::cluster.b32 [%0], %1;" ::"r"(dst + (uint32_t)((c + j) * kFusedRep * 4)), "r"(x[j]) : "memory");
                    }
                }
                tcgen05_fence_before();
            }
            NLMC_FUSED_MARK(4);
            // the spins of block b-1 went to global memory after barrier U(b-1): by now the stores have long been performed, so the
            // proxy fence that the TMA loads behind barrier R(b) need costs nothing here (right after the stores it cost 0.5 us)
            asm volatile("fence.proxy.async.global;" ::: "memory");
            cluster_arrive();                                  // R(b): the partial fields are in their owners' buffers
            cluster_wait();
            NLMC_FUSED_MARK(5);
            // fields of the thread's sites x kE replicas as packed pairs over the SITES: F2[h][ip][e] = (F[2 ip][e], F[2 ip + 1][e]) of sub-block kH tb + h
            f32x2 F2[kH][4][kE];
#pragma unroll
            for (int h = 0; h < kH; ++h)
#pragma unroll
                for (int ip = 0; ip < 4; ++ip) {
                    float f[2][kE];
#pragma unroll
                    for (int h2 = 0; h2 < 2; ++h2) {
                        const int site = 8 * (kH * tb + h) + 2 * ip + h2;
                        ld_e(X + (0 * kBlk + site) * kFusedRep + kE * q, f[h2]);
#pragma unroll
                        for (int src = 1; src < kFusedCluster; ++src) {
                            float t[kE];
                            ld_e(X + (src * kBlk + site) * kFusedRep + kE * q, t);
#pragma unroll
                            for (int e = 0; e < kE; ++e) f[h2][e] += t[e];
                        }
                        const float hk = p.hf[c0 + site];
#pragma unroll
                        for (int e = 0; e < kE; ++e) f[h2][e] += hk;
                    }
#pragma unroll
                    for (int e = 0; e < kE; ++e) F2[h][ip][e] = pack2(f[0][e], f[1][e]);
                }
            asm volatile("bar.sync 1, %0;" ::"n"(kFusedUpdThreads) : "memory");
            NLMC_FUSED_MARK(6);
            // J_bb transposed into the region the receive buffer occupied
            // (row j is read at the columns of its own sub-block and of the later ones only: about half of the 64 KB)
            for (int i = ut; i < kBlk * kBlk / 4; i += kFusedUpdThreads) {
                const int j = i / (kBlk / 4), k4 = i % (kBlk / 4);
                if (4 * k4 < (j & ~7)) continue;
                const uint32_t dj = smem_u32(reinterpret_cast<float4 *>(X) + i);
                const float *src = p.Jf + (size_t)(c0 + j) * p.n_pad + c0 + k4 * 4;
                asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dj), "l"(src) : "memory");
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
            asm volatile("cp.async.wait_group 0;" ::: "memory");
            asm volatile("bar.sync 1, %0;" ::"n"(kFusedUpdThreads) : "memory");
            NLMC_FUSED_MARK(7);
            const float *Jt = X;
            // The thresholds of the NEXT block are computed in the shadow of the chain: by the last warp before it (it is on the
            // critical path only at the very end), by the others once their own sub-blocks are decided.
            float Tn[kH][8][kE];
            uint32_t so_up_n[kH];
            const bool more = b + 1 < p.n_blocks;
            if (more && v == kFusedUpdWarps - 1) thresholds(b + 1, Tn, so_up_n);
            float dmine[kH][8][kE];
            // the flips dl (sites 0..3) / dh (sites 4..7) of sub-block s8 applied to the 8 x kE fields Fh of sub-block my8: two FMA
            // chains of four, summed (the rounding of block_update_chain); a coupling value is loaded once per kE replicas,
            // packed FFMA2 over pairs of sites
            auto apply = [&](int s8, int my8, const float (&dl)[4][kE], const float (&dh)[4][kE], f32x2 (&Fh)[4][kE]) {
                f32x2 C2[4][kE];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const ulonglong2 l0 = *reinterpret_cast<const ulonglong2 *>(Jt + (8 * s8 + j) * kBlk + 8 * my8);
                    const ulonglong2 l1 = *reinterpret_cast<const ulonglong2 *>(Jt + (8 * s8 + j) * kBlk + 8 * my8 + 4);
                    const ulonglong2 h0 = *reinterpret_cast<const ulonglong2 *>(Jt + (8 * s8 + 4 + j) * kBlk + 8 * my8);
                    const ulonglong2 h1 = *reinterpret_cast<const ulonglong2 *>(Jt + (8 * s8 + 4 + j) * kBlk + 8 * my8 + 4);
                    const f32x2 Jl[4] = {l0.x, l0.y, l1.x, l1.y}, Jh[4] = {h0.x, h0.y, h1.x, h1.y};
#pragma unroll
                    for (int e = 0; e < kE; ++e) {
                        const f32x2 ddl = pack2(dl[j][e], dl[j][e]), ddh = pack2(dh[j][e], dh[j][e]);
#pragma unroll
                        for (int ip = 0; ip < 4; ++ip) {
                            Fh[ip][e] = fma2(Jl[ip], ddl, Fh[ip][e]);
                            C2[ip][e] = j == 0 ? mul2(Jh[ip], ddh) : fma2(Jh[ip], ddh, C2[ip][e]);
                        }
                    }
                }
#pragma unroll
                for (int ip = 0; ip < 4; ++ip)
#pragma unroll
                    for (int e = 0; e < kE; ++e) Fh[ip][e] = add2(Fh[ip][e], C2[ip][e]);
            };
            // the 8 decisions of sub-block my8, one after the other, right-looking inside the sub-block
            auto decide = [&](int my8, f32x2 (&Fh)[4][kE], const float (&Th)[8][kE], uint32_t up_bits, float (&dm)[8][kE]) {
                // the couplings inside the sub-block (rows i = 0..6, pairs of columns) before the sequential part
                f32x2 Jr[7][4];
#pragma unroll
                for (int i = 0; i < 7; ++i) {
                    const ulonglong2 lo = *reinterpret_cast<const ulonglong2 *>(Jt + (8 * my8 + i) * kBlk + 8 * my8);
                    const ulonglong2 hi4 = *reinterpret_cast<const ulonglong2 *>(Jt + (8 * my8 + i) * kBlk + 8 * my8 + 4);
                    Jr[i][0] = lo.x; Jr[i][1] = lo.y; Jr[i][2] = hi4.x; Jr[i][3] = hi4.y;
                }
                // the flip of a site is one of two values known beforehand: (+1 - s_old) if its field clears the threshold,
                // (-1 - s_old) otherwise -- compare, select, FFMA2 is the whole dependent chain of a decision
                float dup[8][kE], ddn[8][kE];
#pragma unroll
                for (int i = 0; i < 8; ++i)
#pragma unroll
                    for (int e = 0; e < kE; ++e) {
                        const bool up = ((up_bits >> (kE * i + e)) & 1u) != 0u;
                        dup[i][e] = up ? 0.0f : 2.0f;
                        ddn[i][e] = up ? -2.0f : 0.0f;
                    }
#pragma unroll
                for (int i = 0; i < 8; ++i) {
#pragma unroll
                    for (int e = 0; e < kE; ++e) {
                        const float f = (i & 1) ? hi2(Fh[i >> 1][e]) : lo2(Fh[i >> 1][e]);
                        const float d = f > Th[i][e] ? dup[i][e] : ddn[i][e];
                        dm[i][e] = d;
                        if (i < 7) {   // the pair that holds site i itself is updated too (its own half is dead)
                            const f32x2 dd = pack2(d, d);
#pragma unroll
                            for (int ip = (i + 1) >> 1; ip < 4; ++ip) Fh[ip][e] = fma2(Jr[i][ip], dd, Fh[ip][e]);
                        }
                    }
                }
            };
            // The tiles in order.  Step s: the owners of tile s (kLanesPerSub lanes of warp s / kE) decide its sub-blocks -- the
            // flips of the first go into the fields of the second without leaving the thread -- and publish the flips; every
            // thread whose tile comes later applies them.
            for (int s = 0; s < kTiles; ++s) {
                const int vs = s / kE;
                if (v < vs) break;                             // every tile of this warp is done
                if (v == vs) {
                    if (g == s % kE) {
#pragma unroll
                        for (int h = 0; h < kH; ++h) {
                            decide(kH * tb + h, F2[h], T[h], so_up[h], dmine[h]);
#pragma unroll
                            for (int i = 0; i < 8; ++i) st_e(dbuf + (8 * (kH * tb + h) + i) * kFusedRep + kE * q, dmine[h][i]);
#pragma unroll
                            for (int h2 = h + 1; h2 < kH; ++h2) {
                                float dl[4][kE], dh[4][kE];
#pragma unroll
                                for (int j = 0; j < 4; ++j)
#pragma unroll
                                    for (int e = 0; e < kE; ++e) { dl[j][e] = dmine[h][j][e]; dh[j][e] = dmine[h][4 + j][e]; }
                                apply(kH * tb + h, kH * tb + h2, dl, dh, F2[h2]);
                            }
                        }
                    }
                    __syncwarp();
                    if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(sub_done + s)) : "memory");
                } else {
                    mbar_wait(sub_done + s, (uint32_t)(b & 1));    // hardware wait, no polling traffic on the shared-memory pipe
                }
                if (tb > s) {
#pragma unroll
                    for (int hs = 0; hs < kH; ++hs) {
                        const int s8 = kH * s + hs;
                        float dl[4][kE], dh[4][kE];
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            ld_e(dbuf + (8 * s8 + j) * kFusedRep + kE * q, dl[j]);
                            ld_e(dbuf + (8 * s8 + 4 + j) * kFusedRep + kE * q, dh[j]);
                        }
#pragma unroll
                        for (int h = 0; h < kH; ++h) apply(s8, kH * tb + h, dl, dh, F2[h]);
                    }
                }
            }
            // new spins = old spins + flips, as bf16 (+1 = 0x3F80, -1 = 0xBF80; 0 past the end of J)
            uint32_t bits[kH][kE][4];
#pragma unroll
            for (int h = 0; h < kH; ++h)
#pragma unroll
                for (int e = 0; e < kE; ++e)
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        const float sn = (((so_up[h] >> (kE * i + e)) & 1u) ? 1.0f : -1.0f) + dmine[h][i][e];
                        const uint32_t b16 = (8 * (kH * tb + h) + i < k_end) ? (sn > 0.0f ? 0x3F80u : 0xBF80u) : 0u;
                        if ((i & 1) == 0) bits[h][e][i >> 1] = b16; else bits[h][e][i >> 1] |= b16 << 16;
                    }
            NLMC_FUSED_MARK(8);
            if (more) {
                // a sub-block's 8 sites are one 16-byte chunk of a spin row of k-block sb / 8 of this block: into the spin slot of the
                // ring stage that will hold that k-block for block b+1 (128-byte swizzle: chunk index XOR row % 8)
                mbar_wait(slots_free, (uint32_t)(b & 1));   // the stages' previous tiles have been multiplied
                const uint32_t it_d0 = it_next + (uint32_t)n_indep(b + 1);
#pragma unroll
                for (int h = 0; h < kH; ++h) {
                    const int sb = kH * tb + h;
                    const int st = (int)((it_d0 + (uint32_t)(sb >> 3)) % kFusedStages);
                    uint8_t *slot = smem + (size_t)st * (1 + kMaxSplit) * kTileBytes;
#pragma unroll
                    for (int e = 0; e < kE; ++e) {
                        const int row = (int)rank * kFusedRep + kE * q + e;
                        *reinterpret_cast<uint4 *>(slot + (row >> 3) * 1024 + (row & 7) * 128 + (((sb & 7) ^ (row & 7)) << 4)) =
                            make_uint4(bits[h][e][0], bits[h][e][1], bits[h][e][2], bits[h][e][3]);
                    }
                }
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic stores to shared memory -> read by the tensor core (async proxy)
                __syncwarp();
                if (lane == 0) asm volatile("mbarrier.arrive.release.cta.shared::cta.b64 _, [%0];" ::"r"(smem_u32(adep)) : "memory");
            }
            NLMC_FUSED_MARK(9);
            cluster_arrive();                                  // U(b): this thread is done with the chain, J_bb and the flip buffer
            // The new spins to global memory, for the TMA loads of LATER blocks (block b+2 at the earliest, issued after barrier
            // R(b+1), which these stores and their proxy fence -- executed just before that arrive -- precede in program order)
            // and for the next sweep.
#pragma unroll
            for (int h = 0; h < kH; ++h)
#pragma unroll
                for (int e = 0; e < kE; ++e)
                    *reinterpret_cast<uint4 *>(p.S + (size_t)(rq + e) * p.n_pad + c0 + 8 * (kH * tb + h)) =
                        make_uint4(bits[h][e][0], bits[h][e][1], bits[h][e][2], bits[h][e][3]);
            if (more && v != kFusedUpdWarps - 1) thresholds(b + 1, Tn, so_up_n);
            if (more) {
#pragma unroll
                for (int h = 0; h < kH; ++h) {
#pragma unroll
                    for (int i = 0; i < 8; ++i)
#pragma unroll
                        for (int e = 0; e < kE; ++e) T[h][i][e] = Tn[h][i][e];
                    so_up[h] = so_up_n[h];
                }
            }
        }
        cluster_wait();                                        // U(last): nobody writes into this CTA's buffers any more
    }
    tcgen05_fence_before();
    __syncthreads();
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(2 * kBN) : "memory");
    }
}

// E_r = -(1/2) sum_k s_k (J s)_k - sum_k h_k s_k from the full field matrix Ht (fp32 products, fp64 accumulation)
__global__ void dense_energy_kernel(int n, int n_pad, int R_pad, const float *__restrict__ Ht, const float *__restrict__ hf,
                                    const uint16_t *__restrict__ S, double *E) {
    const int r = blockIdx.x * blockDim.x + threadIdx.x;
    if (r >= R_pad) return;
    double quad = 0.0, lin = 0.0;
    for (int k = 0; k < n; ++k) {
        const uint16_t b16 = S[(size_t)r * n_pad + k];
        const double sk = b16 == 0 ? 0.0 : ((b16 & 0x8000u) ? -1.0 : 1.0);
        quad += sk * (double)Ht[(size_t)k * R_pad + r];
        lin += sk * (double)hf[k];
    }
    E[r] = -(0.5 * quad + lin);
}

__global__ void dense_bump_kernel(uint32_t *counter) { *counter += 1u; }

// NMC bookkeeping m_init = M[:, argmin E] (first minimum wins, NMC/nmc.py:394-395): one CTA per replica copies
// the row when the current energy is strictly below the best seen since the last reset.
__global__ void dense_best_kernel(int n_pad, const double *__restrict__ E, double *bestE, const uint16_t *__restrict__ S,
                                  uint16_t *bestS) {
    const int r = blockIdx.x;
    const bool better = E[r] < bestE[r];
    if (better) {
        const uint4 *src = reinterpret_cast<const uint4 *>(S + (size_t)r * n_pad);
        uint4 *dst = reinterpret_cast<uint4 *>(bestS + (size_t)r * n_pad);
        for (int i = threadIdx.x; i < n_pad / 8; i += blockDim.x) dst[i] = src[i];
    }
    __syncthreads();
    if (better && threadIdx.x == 0) bestE[r] = E[r];
}
__global__ void dense_fill_kernel(int count, double *p, double v) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = v;
}

__global__ void dense_pack_kernel(int n, int n_pad, int R, const int8_t *in, uint16_t *S) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)R * n) return;
    const int r = (int)(i / n), k = (int)(i % n);
    const int v = in[i];
    S[(size_t)r * n_pad + k] = v > 0 ? 0x3F80 : (v < 0 ? 0xBF80 : 0);
}
__global__ void dense_unpack_kernel(int n, int n_pad, int R, const uint16_t *S, int8_t *out) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)R * n) return;
    const int r = (int)(i / n), k = (int)(i % n);
    const uint16_t b = S[(size_t)r * n_pad + k];
    out[i] = b == 0 ? 0 : ((b & 0x8000u) ? -1 : 1);
}
__global__ void dense_init_kernel(int n, int n_pad, int R_pad, uint16_t *S, uint32_t seed_lo, uint32_t seed_hi) {
    const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
    if (i >= (size_t)R_pad * (n_pad / 4)) return;
    const int r = (int)(i / (n_pad / 4)), k4 = (int)(i % (n_pad / 4)) * 4;
    const PhiloxD rng{seed_lo, seed_hi ^ 0x494e4954u};
    const uint4 x = rng((uint32_t)r, (uint32_t)k4, 0u, 1u);
    const uint32_t b[4] = {x.x, x.y, x.z, x.w};
    for (int e = 0; e < 4; ++e) S[(size_t)r * n_pad + k4 + e] = (k4 + e < n) ? ((b[e] & 1u) ? 0x3F80 : 0xBF80) : 0;
}

}  // namespace nlmc

extern "C" {

int nlmc_dense_destroy(nlmc_dense *D) {
    if (!D) return NLMC_OK;
    cudaSetDevice(D->inst->device);
    if (D->fused_prof) {   // development aid: where one block step of the cluster sweep spends its time (CTA 0, first update warp)
        const int nb = (D->n + nlmc::kBlk - 1) / nlmc::kBlk;
        std::vector<unsigned long long> t((size_t)nb * 10);
        cudaDeviceSynchronize();
        cudaMemcpy(t.data(), D->fused_prof, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost);
        static const char *names[9] = {"thresholds", "wait U(b-1)", "wait accumulator", "drain + push", "cluster barrier R", "sum",
                                       "stage J_bb", "chain", "store"};
        for (int b = 0; b < nb; ++b) {
            fprintf(stderr, "fused block %2d:", b);
            for (int i = 0; i < 9; ++i) fprintf(stderr, " %s %.2f us |", names[i], 1e-3 * (double)(t[(size_t)b * 10 + i + 1] - t[(size_t)b * 10 + i]));
            if (b + 1 < nb) fprintf(stderr, " step %.2f us", 1e-3 * (double)(t[(size_t)(b + 1) * 10] - t[(size_t)b * 10]));
            fprintf(stderr, "\n");
        }
        cudaFree(D->fused_prof);
    }
    void *ptrs[] = {D->S, D->Jp[0], D->Jp[1], D->Jp[2], D->Jf, D->hf, D->Ht, D->beta, D->E, D->d_sweep, D->bestE, D->bestS, D->modes};
    for (void *p : ptrs) if (p) cudaFree(p);
    D->xch.release();
    if (D->sweep_graph) cudaGraphExecDestroy(D->sweep_graph);
    if (D->ev0) cudaEventDestroy(D->ev0);
    if (D->ev1) cudaEventDestroy(D->ev1);
    for (cudaEvent_t e : D->fork_ev) cudaEventDestroy(e);
    for (cudaStream_t st : D->chain_streams) cudaStreamDestroy(st);
    if (D->stream2) cudaStreamDestroy(D->stream2);
    if (D->stream) cudaStreamDestroy(D->stream);
    delete D;
    return NLMC_OK;
}

int nlmc_dense_set_betas(nlmc_dense *D, const double *betas) {
    NLMC_REQUIRE(D && betas, "nlmc_dense_set_betas: NULL argument");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    std::vector<float> b((size_t)D->R_pad, 1.0f);
    for (int r = 0; r < D->R; ++r) b[(size_t)r] = (float)betas[r];
    NLMC_CUDA(cudaMemcpyAsync(D->beta, b.data(), sizeof(float) * b.size(), cudaMemcpyHostToDevice, D->stream));
    NLMC_CUDA(cudaStreamSynchronize(D->stream));
    return NLMC_OK;
}

int nlmc_dense_create(nlmc_instance *I, int n_replicas, const double *betas, int n_split, unsigned long long seed,
                      nlmc_dense **out) {
    using namespace nlmc;
    NLMC_REQUIRE(I && out && betas, "nlmc_dense_create: NULL argument");
    *out = nullptr;
    NLMC_REQUIRE(n_replicas >= 1 && n_split >= 1 && n_split <= 3, "nlmc_dense_create: n_replicas >= 1 and n_split in 1..3");
    NLMC_REQUIRE(nlmc::instance_value_symmetric(I), "nlmc_dense_create: J must be symmetric (J_ij == J_ji) without repeated entries");
    int cc_major = 0;
    NLMC_CUDA(cudaDeviceGetAttribute(&cc_major, cudaDevAttrComputeCapabilityMajor, I->device));
    if (cc_major != 10) {
        set_error("nlmc_dense_create: the tcgen05 path needs an sm_100 device (found sm_%d x)", cc_major);
        return NLMC_ERR_UNSUPPORTED;
    }
    NLMC_CUDA(cudaSetDevice(I->device));
    auto *D = new nlmc_dense();
    D->inst = I;
    D->n = I->n;
    D->n_pad = ((I->n + kBN - 1) / kBN) * kBN;
    D->R = n_replicas;
    D->R_pad = ((n_replicas + kBM - 1) / kBM) * kBM;
    {
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, I->device);
        D->upd_rpc = std::max(1, std::min(kMaxRepPerCta, (D->R_pad + sms - 1) / std::max(1, sms)));
        if (const char *e = getenv("NLMC_DENSE_RPC")) D->upd_rpc = std::max(1, std::min(kMaxRepPerCta, atoi(e)));
    }
    D->n_split = n_split;
    D->seed = seed;
    const size_t np = (size_t)D->n_pad, nn = np * np;
    // dense J from the CSR mirror; bf16 pieces J = J1 + J2 + J3 (each the bf16 rounding of the remaining residual)
    std::vector<float> Jf(nn, 0.f), hf(np, 0.f);
    std::vector<uint16_t> piece[3];
    for (int q = 0; q < n_split; ++q) piece[q].assign(nn, 0);
    for (int i = 0; i < I->n; ++i) {
        hf[(size_t)i] = (float)I->h_h[(size_t)i];
        for (int p = I->h_row_ptr[i]; p < I->h_row_ptr[i + 1]; ++p) {
            const size_t idx = (size_t)i * np + (size_t)I->h_col[(size_t)p];
            double rest = I->h_val[(size_t)p];
            Jf[(size_t)I->h_col[(size_t)p] * np + (size_t)i] += (float)rest;  // transposed copy
            for (int q = 0; q < n_split; ++q) {
                const uint16_t b = f32_to_bf16_rn((float)rest);
                piece[q][idx] = b;
                rest -= (double)bf16_to_f32(b);
            }
        }
    }
    bool ok = cudaStreamCreateWithFlags(&D->stream, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&D->stream2, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreate(&D->ev0) == cudaSuccess && cudaEventCreate(&D->ev1) == cudaSuccess &&
              cudaMalloc(&D->S, sizeof(uint16_t) * (size_t)D->R_pad * np) == cudaSuccess &&
              cudaMalloc(&D->Jf, sizeof(float) * nn) == cudaSuccess && cudaMalloc(&D->hf, sizeof(float) * np) == cudaSuccess &&
              cudaMalloc(&D->Ht, sizeof(float) * np * (size_t)D->R_pad) == cudaSuccess &&
              cudaMalloc(&D->beta, sizeof(float) * (size_t)D->R_pad) == cudaSuccess &&
              cudaMalloc(&D->E, sizeof(double) * (size_t)D->R_pad) == cudaSuccess &&
              cudaMalloc(&D->d_sweep, sizeof(uint32_t)) == cudaSuccess && cudaMemset(D->d_sweep, 0, sizeof(uint32_t)) == cudaSuccess &&
              cudaMemcpy(D->Jf, Jf.data(), sizeof(float) * nn, cudaMemcpyHostToDevice) == cudaSuccess &&
              cudaMemcpy(D->hf, hf.data(), sizeof(float) * np, cudaMemcpyHostToDevice) == cudaSuccess;
    for (int q = 0; ok && q < n_split; ++q)
        ok = cudaMalloc(&D->Jp[q], sizeof(uint16_t) * nn) == cudaSuccess &&
             cudaMemcpy(D->Jp[q], piece[q].data(), sizeof(uint16_t) * nn, cudaMemcpyHostToDevice) == cudaSuccess;
    if (!ok) {
        set_error("nlmc_dense_create: CUDA allocation/copy failed: %s", cudaGetErrorString(cudaGetLastError()));
        nlmc_dense_destroy(D);
        return NLMC_ERR_CUDA;
    }
    int rc = make_map(&D->map_S, D->S, (uint64_t)D->R_pad, np, np);
    for (int q = 0; !rc && q < n_split; ++q) rc = make_map(&D->map_J[q], D->Jp[q], np, np, np);
    if (!rc) {
        if (cudaFuncSetAttribute(gemm_bf16_tn_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes(3)) != cudaSuccess ||
            cudaFuncSetAttribute(gemm_bf16_tn_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes(2)) != cudaSuccess ||
            cudaFuncSetAttribute(gemm_bf16_tn_kernel<2, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes(2, 2)) != cudaSuccess ||
            cudaFuncSetAttribute(gemm_bf16_tn_kernel<2, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)gemm_smem_bytes(2, 2)) != cudaSuccess ||
            cudaFuncSetAttribute(dense_block_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)update_smem_bytes(kMaxRepPerCta)) != cudaSuccess ||
            cudaFuncSetAttribute(dense_fused_sweep_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kFusedSmemBytes) != cudaSuccess) {
            set_error("nlmc_dense_create: cudaFuncSetAttribute failed: %s", cudaGetErrorString(cudaGetLastError()));
            rc = NLMC_ERR_CUDA;
        }
    }
    if (!rc) rc = nlmc_dense_set_betas(D, betas);
    if (rc) {
        nlmc_dense_destroy(D);
        return rc;
    }
    const size_t items = (size_t)D->R_pad * (np / 4);
    dense_init_kernel<<<(unsigned)((items + 255) / 256), 256, 0, D->stream>>>(D->n, D->n_pad, D->R_pad, D->S, (uint32_t)seed,
                                                                             (uint32_t)(seed >> 32));
    if (cudaGetLastError() != cudaSuccess || cudaStreamSynchronize(D->stream) != cudaSuccess) {
        set_error("nlmc_dense_create: init kernel failed");
        nlmc_dense_destroy(D);
        return NLMC_ERR_CUDA;
    }
    *out = D;
    return NLMC_OK;
}

int nlmc_dense_set_spins(nlmc_dense *D, const int8_t *spins) {
    using namespace nlmc;
    NLMC_REQUIRE(D && spins, "nlmc_dense_set_spins: NULL argument");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    int8_t *tmp = nullptr;
    const size_t cnt = (size_t)D->R * D->n;
    NLMC_CUDA(cudaMalloc(&tmp, cnt));
    cudaError_t e = cudaMemcpyAsync(tmp, spins, cnt, cudaMemcpyHostToDevice, D->stream);
    if (e == cudaSuccess) {
        dense_pack_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, D->stream>>>(D->n, D->n_pad, D->R, tmp, D->S);
        e = cudaGetLastError();
    }
    if (e == cudaSuccess) e = cudaStreamSynchronize(D->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { set_error("nlmc_dense_set_spins: %s", cudaGetErrorString(e)); return NLMC_ERR_CUDA; }
    return NLMC_OK;
}

int nlmc_dense_get_spins(nlmc_dense *D, int8_t *out) {
    using namespace nlmc;
    NLMC_REQUIRE(D && out, "nlmc_dense_get_spins: NULL argument");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    int8_t *tmp = nullptr;
    const size_t cnt = (size_t)D->R * D->n;
    NLMC_CUDA(cudaMalloc(&tmp, cnt));
    dense_unpack_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, D->stream>>>(D->n, D->n_pad, D->R, D->S, tmp);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(out, tmp, cnt, cudaMemcpyDeviceToHost, D->stream);
    if (e == cudaSuccess) e = cudaStreamSynchronize(D->stream);
    cudaFree(tmp);
    if (e != cudaSuccess) { set_error("nlmc_dense_get_spins: %s", cudaGetErrorString(e)); return NLMC_ERR_CUDA; }
    return NLMC_OK;
}

/* full field recompute H = S . J for all replicas (one tensor-core GEMM); out_H [R][n] optional */
int nlmc_dense_fields(nlmc_dense *D, float *out_H) {
    using namespace nlmc;
    NLMC_REQUIRE(D, "nlmc_dense_fields: NULL handle");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    int rc = launch_fields(D, 0, D->n_pad, 1);
    if (rc) return rc;
    if (out_H) {
        std::vector<float> ht((size_t)D->n_pad * D->R_pad);
        NLMC_CUDA(cudaMemcpyAsync(ht.data(), D->Ht, sizeof(float) * ht.size(), cudaMemcpyDeviceToHost, D->stream));
        NLMC_CUDA(cudaStreamSynchronize(D->stream));
        for (int r = 0; r < D->R; ++r)
            for (int k = 0; k < D->n; ++k) out_H[(size_t)r * D->n + k] = ht[(size_t)k * D->R_pad + r];
    }
    return NLMC_OK;
}

// One sweep = the blocks of 128 sites in order.  The fields of block b+1 need the spins AFTER the update of block b, but
// only through the 128 columns of block b: the contraction over all other columns ("main part", 15/16 of the work) is
// issued on a second stream as soon as the update of block b-1 is done and overlaps the latency-bound update of block b;
// the two k-blocks of block b ("correction") follow the update.  Both parts accumulate atomically into the field rows of
// block b+1, which the update of block b-1 cleared.
static int enqueue_sweep(nlmc_dense *D, int k_splits) {
    using namespace nlmc;
    const int nb = (D->n + kBlk - 1) / kBlk, kb_all = D->n_pad / kBK, kb_blk = kBlk / kBK;
    // Measured on B200 at C3 size: 0.470 ms per sweep with the look-ahead against 0.355 ms without -- the third kernel per
    // block (the correction, 16 CTAs for two k-blocks) and the extra dependency edges cost more than the overlap hides.
    // Kept behind a switch; the default is the plain GEMM -> update chain.
    const bool look_ahead = nb > 1 && getenv("NLMC_DENSE_LOOKAHEAD") != nullptr;
    // Programmatic dependent launch of the GEMM -> update chain (the update's coupling-block staging and thresholds overlap
    // the GEMM, a 2-stage GEMM CTA and an update CTA fit on one SM together): measured 0.455 ms per sweep against 0.357 ms
    // with plain stream order on B200 at C3 size, so it is off unless NLMC_DENSE_PDL is set.
    const bool pdl = getenv("NLMC_DENSE_PDL") != nullptr;
    auto launch_update = [&](int b, float *zero_rows) {
        launch_pdl(dense_block_update_kernel, dim3((unsigned)((D->R_pad + D->upd_rpc - 1) / D->upd_rpc)), dim3((unsigned)(32 * D->upd_rpc)),
                   update_smem_bytes(D->upd_rpc), D->stream, pdl && !look_ahead, D->n, D->n_pad, D->R_pad, b * kBlk, (const float *)D->Ht,
                   (const float *)D->Jf, (const float *)D->hf, (const float *)D->beta, D->S, (uint32_t)D->seed, (uint32_t)(D->seed >> 32),
                   (const uint32_t *)D->d_sweep, zero_rows, (const uint8_t *)(D->modes_on ? D->modes : nullptr), D->temp_x,
                   (int)((pdl && !look_ahead) ? (getenv("NLMC_DENSE_PDL_EARLY") ? 2 : 1) : 0), 0, D->R_pad);
    };
    int rc;
    // The whole sweep as one launch of the cluster kernel (dense_fused_sweep_kernel); NLMC_DENSE_FUSED=0 selects the
    // GEMM -> update chain below.
    {
        const char *e = getenv("NLMC_DENSE_FUSED");
        const bool fused = e ? atoi(e) != 0 : true;
        if (fused && !look_ahead && !pdl && !getenv("NLMC_DENSE_CHAINS") && !getenv("NLMC_DENSE_SKIP")) {
            FusedParams fp;
            fp.n = D->n; fp.n_pad = D->n_pad; fp.R_pad = D->R_pad; fp.n_blocks = nb; fp.kb_total = kb_all; fp.n_split = D->n_split;
            fp.Jf = D->Jf; fp.hf = D->hf; fp.beta = D->beta; fp.S = D->S;
            fp.seed_lo = (uint32_t)D->seed; fp.seed_hi = (uint32_t)(D->seed >> 32);
            fp.sweep_ptr = D->d_sweep; fp.modes = D->modes_on ? D->modes : nullptr; fp.temp_x = D->temp_x;
            fp.prof = D->fused_prof;
            dense_fused_sweep_kernel<<<(unsigned)(kFusedCluster * (D->R_pad / kBM)), kFusedThreads, kFusedSmemBytes, D->stream>>>(
                D->map_S, D->map_J[0], D->map_J[D->n_split > 1 ? 1 : 0], D->map_J[D->n_split > 2 ? 2 : 0], fp);
            NLMC_CUDA(cudaGetLastError());
            dense_bump_kernel<<<1, 1, 0, D->stream>>>(D->d_sweep);
            NLMC_CUDA(cudaGetLastError());
            return NLMC_OK;
        }
    }
    // Replicas are independent, so the GEMM -> update chain of one range of replica tiles never has to wait for another
    // range: the sweep can run as n_chains independent chains on their own streams (forked and joined inside the captured
    // graph), the GEMM of one chain resident next to the update of another.  Measured on B200 at C3 size: 0.358 ms per
    // sweep with one chain, 0.464 / 0.426 / 0.511 ms with 2 / 4 / 8 -- like the look-ahead and the dependent launch above,
    // concurrency between these latency-bound kernels does not pay.  Default 1; NLMC_DENSE_CHAINS selects more.
    const int m_tiles_all = D->R_pad / kBM;
    int n_chains = 1;
    if (const char *e = getenv("NLMC_DENSE_CHAINS")) n_chains = std::max(1, std::min(m_tiles_all, atoi(e)));
    if (n_chains > 1 && !look_ahead && !pdl) {
        while ((int)D->chain_streams.size() < n_chains - 1) {
            cudaStream_t st = nullptr;
            NLMC_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
            D->chain_streams.push_back(st);
        }
        while ((int)D->fork_ev.size() < n_chains) {
            cudaEvent_t e = nullptr;
            NLMC_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
            D->fork_ev.push_back(e);
        }
        int sms = 148;
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, D->inst->device);
        NLMC_CUDA(cudaMemsetAsync(D->Ht, 0, sizeof(float) * (size_t)kBlk * D->R_pad, D->stream));   // field rows of block 0
        NLMC_CUDA(cudaEventRecord(D->fork_ev[0], D->stream));
        for (int c = 0; c < n_chains; ++c) {
            cudaStream_t st = c == 0 ? D->stream : D->chain_streams[(size_t)c - 1];
            if (c > 0) NLMC_CUDA(cudaStreamWaitEvent(st, D->fork_ev[0], 0));
            const int t0 = (int)((long long)m_tiles_all * c / n_chains), t1 = (int)((long long)m_tiles_all * (c + 1) / n_chains);
            const int r_base = t0 * kBM, r_end = t1 * kBM;
            const int rpc = std::max(1, std::min(kMaxRepPerCta, (r_end - r_base + sms - 1) / sms));
            int ks = std::max(1, std::min(kb_all / 2, sms / std::max(1, t1 - t0)));
            if (const char *e = getenv("NLMC_DENSE_KSPLIT")) ks = std::max(1, atoi(e));
            for (int b = 0; b < nb; ++b) {
                if ((rc = launch_fields_part(D, st, b * kBlk, kBlk, ks, /*clear=*/false, 0, kb_all, 0, 0, 2, false, t0, t1 - t0))) return rc;
                float *zero_rows = b + 1 < nb ? D->Ht + (size_t)(b + 1) * kBlk * D->R_pad : nullptr;
                dense_block_update_kernel<<<(unsigned)((r_end - r_base + rpc - 1) / rpc), (unsigned)(32 * rpc), update_smem_bytes(rpc), st>>>(
                    D->n, D->n_pad, D->R_pad, b * kBlk, D->Ht, D->Jf, D->hf, D->beta, D->S, (uint32_t)D->seed, (uint32_t)(D->seed >> 32),
                    D->d_sweep, zero_rows, D->modes_on ? D->modes : nullptr, D->temp_x, 0, r_base, r_end);
            }
            if (c > 0) {
                NLMC_CUDA(cudaEventRecord(D->fork_ev[(size_t)c], st));
                NLMC_CUDA(cudaStreamWaitEvent(D->stream, D->fork_ev[(size_t)c], 0));
            }
        }
    } else if (!look_ahead) {
        const char *skip = getenv("NLMC_DENSE_SKIP");   // timing experiments only: "gemm" or "update"
        for (int b = 0; b < nb; ++b) {
            if (!(skip && skip[0] == 'g'))
                if ((rc = launch_fields_part(D, D->stream, b * kBlk, kBlk, k_splits, /*clear=*/b == 0, 0, kb_all, 0, 0, (pdl || getenv("NLMC_DENSE_STAGES2")) ? 2 : 3,
                                             pdl && b > 0))) return rc;
            if (!(skip && skip[0] == 'u'))
                launch_update(b, (k_splits > 1 && b + 1 < nb) ? D->Ht + (size_t)(b + 1) * kBlk * D->R_pad : nullptr);
        }
    } else {
        size_t ev = 0;
        auto next_event = [&]() -> cudaEvent_t {
            if (ev == D->fork_ev.size()) {
                cudaEvent_t e = nullptr;
                if (cudaEventCreateWithFlags(&e, cudaEventDisableTiming) != cudaSuccess) return nullptr;
                D->fork_ev.push_back(e);
            }
            return D->fork_ev[ev++];
        };
        // field rows of blocks 0 and 1 start from zero (the later ones are cleared by the update two blocks before)
        NLMC_CUDA(cudaMemsetAsync(D->Ht, 0, sizeof(float) * (size_t)std::min(2 * kBlk, D->n_pad) * D->R_pad, D->stream));
        cudaEvent_t e_ready = next_event();          // "spins of all blocks before b are final, rows of block b+1 are clear"
        if (!e_ready) return NLMC_ERR_CUDA;
        NLMC_CUDA(cudaEventRecord(e_ready, D->stream));
        if ((rc = launch_fields_part(D, D->stream, 0, kBlk, k_splits, /*clear=*/false, 0, kb_all, 0, 0, 2))) return rc;
        for (int b = 0; b < nb; ++b) {
            cudaEvent_t e_main = nullptr;
            if (b + 1 < nb) {   // main part of block b+1 on the second stream: every column except those of block b
                NLMC_CUDA(cudaStreamWaitEvent(D->stream2, e_ready, 0));
                if ((rc = launch_fields_part(D, D->stream2, (b + 1) * kBlk, kBlk, k_splits, /*clear=*/false, 0, kb_all,
                                             b * kb_blk, kb_blk, 2))) return rc;
                if (!(e_main = next_event())) return NLMC_ERR_CUDA;
                NLMC_CUDA(cudaEventRecord(e_main, D->stream2));
            }
            launch_update(b, b + 2 < nb ? D->Ht + (size_t)(b + 2) * kBlk * D->R_pad : nullptr);
            if (b + 1 < nb) {
                if (!(e_ready = next_event())) return NLMC_ERR_CUDA;
                NLMC_CUDA(cudaEventRecord(e_ready, D->stream));
                // correction: the columns of block b, now final
                if ((rc = launch_fields_part(D, D->stream, (b + 1) * kBlk, kBlk, 1, /*clear=*/false, b * kb_blk, kb_blk, 0, 0, 2)))
                    return rc;
                NLMC_CUDA(cudaStreamWaitEvent(D->stream, e_main, 0));
            }
        }
    }
    dense_bump_kernel<<<1, 1, 0, D->stream>>>(D->d_sweep);
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
}

int nlmc_dense_sweep(nlmc_dense *D, int n_sweeps) {
    using namespace nlmc;
    NLMC_REQUIRE(D && n_sweeps >= 0, "nlmc_dense_sweep: bad arguments");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    if (!D->sweep_graph) {  // capture one sweep once; replays cost one launch each
        if (getenv("NLMC_DENSE_FUSED_PROF") && !D->fused_prof) {
            const size_t cnt = 10 * (size_t)((D->n + kBlk - 1) / kBlk);
            NLMC_CUDA(cudaMalloc(&D->fused_prof, sizeof(unsigned long long) * cnt));
            NLMC_CUDA(cudaMemset(D->fused_prof, 0, sizeof(unsigned long long) * cnt));
        }
        const int m_tiles = D->R_pad / kBM;
        int k_splits = std::max(1, std::min(D->n_pad / kBK, 148 / std::max(1, m_tiles)));
        if (const char *e = getenv("NLMC_DENSE_KSPLIT")) k_splits = std::max(1, atoi(e));
        cudaGraph_t graph = nullptr;
        NLMC_CUDA(cudaStreamBeginCapture(D->stream, cudaStreamCaptureModeThreadLocal));
        const int rc = enqueue_sweep(D, k_splits);
        const cudaError_t e = cudaStreamEndCapture(D->stream, &graph);
        if (rc || e != cudaSuccess) {
            if (graph) cudaGraphDestroy(graph);
            if (!rc) set_error("nlmc_dense_sweep: stream capture failed: %s", cudaGetErrorString(e));
            return rc ? rc : NLMC_ERR_CUDA;
        }
        const cudaError_t e2 = cudaGraphInstantiate(&D->sweep_graph, graph, 0);
        cudaGraphDestroy(graph);
        if (e2 != cudaSuccess) {
            set_error("nlmc_dense_sweep: cudaGraphInstantiate failed: %s", cudaGetErrorString(e2));
            D->sweep_graph = nullptr;
            return NLMC_ERR_CUDA;
        }
    }
    for (int s = 0; s < n_sweeps; ++s) NLMC_CUDA(cudaGraphLaunch(D->sweep_graph, D->stream));
    return NLMC_OK;
}

int nlmc_dense_energies(nlmc_dense *D, double *out_E) {
    using namespace nlmc;
    NLMC_REQUIRE(D && out_E, "nlmc_dense_energies: NULL argument");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    int rc = launch_fields(D, 0, D->n_pad, 1);
    if (rc) return rc;
    dense_energy_kernel<<<(D->R_pad + 127) / 128, 128, 0, D->stream>>>(D->n, D->n_pad, D->R_pad, D->Ht, D->hf, D->S, D->E);
    NLMC_CUDA(cudaGetLastError());
    NLMC_CUDA(cudaMemcpyAsync(out_E, D->E, sizeof(double) * (size_t)D->R, cudaMemcpyDeviceToHost, D->stream));
    NLMC_CUDA(cudaStreamSynchronize(D->stream));
    return NLMC_OK;
}

/* ---- replica exchange by beta labels (nlmc_exchange.cuh) ---- */
int nlmc_dense_ladders(nlmc_dense *D, int n_beta, const double *betas) {
    NLMC_REQUIRE(D, "nlmc_dense_ladders: NULL handle");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    return nlmc::exchange_setup<float>(D->xch, D->R, n_beta, betas, D->beta, D->stream);
}

int nlmc_dense_exchange(nlmc_dense *D, int num_swapping_pairs) {
    using namespace nlmc;
    NLMC_REQUIRE(D && D->xch.active(), "nlmc_dense_exchange: call nlmc_dense_ladders first");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    int rc = launch_fields(D, 0, D->n_pad, 1);
    if (rc) return rc;
    dense_energy_kernel<<<(D->R_pad + 127) / 128, 128, 0, D->stream>>>(D->n, D->n_pad, D->R_pad, D->Ht, D->hf, D->S, D->E);
    NLMC_CUDA(cudaGetLastError());
    NLMC_CUDA(cudaMemcpyAsync(D->xch.E, D->E, sizeof(double) * (size_t)D->R, cudaMemcpyDeviceToDevice, D->stream));
    return exchange_launch<float>(D->xch, num_swapping_pairs, D->beta, D->seed, 0, D->stream);
}

int nlmc_dense_labels(nlmc_dense *D, int32_t *out_labels, int n_rounds, int32_t *out_counts) {
    NLMC_REQUIRE(D, "nlmc_dense_labels: NULL handle");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    return nlmc::exchange_fetch(D->xch, D->R, out_labels, n_rounds, out_counts, D->stream);
}

int nlmc_dense_set_site_modes(nlmc_dense *D, const uint8_t *modes, double temp_x) {
    NLMC_REQUIRE(D, "nlmc_dense_set_site_modes: NULL handle");
    NLMC_REQUIRE(!modes || temp_x > 0.0, "nlmc_dense_set_site_modes: temp_x must be positive");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    const bool was_on = D->modes_on;
    if (!modes) {
        D->modes_on = false;
    } else {
        if (!D->modes) {
            NLMC_CUDA(cudaMalloc(&D->modes, (size_t)D->R_pad * D->n_pad));
            NLMC_CUDA(cudaMemset(D->modes, 0, (size_t)D->R_pad * D->n_pad));
        }
        NLMC_CUDA(cudaMemcpy2DAsync(D->modes, (size_t)D->n_pad, modes, (size_t)D->n, (size_t)D->n, (size_t)D->R,
                                    cudaMemcpyHostToDevice, D->stream));
        NLMC_CUDA(cudaStreamSynchronize(D->stream));
        D->modes_on = true;
        D->temp_x = (float)temp_x;
    }
    if (D->sweep_graph && (was_on != D->modes_on || modes)) {  // kernel arguments are baked into the captured graph
        cudaGraphExecDestroy(D->sweep_graph);
        D->sweep_graph = nullptr;
    }
    return NLMC_OK;
}

int nlmc_dense_best_reset(nlmc_dense *D) {
    using namespace nlmc;
    NLMC_REQUIRE(D, "nlmc_dense_best_reset: NULL handle");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    if (!D->bestE) {
        NLMC_CUDA(cudaMalloc(&D->bestE, sizeof(double) * (size_t)D->R_pad));
        NLMC_CUDA(cudaMalloc(&D->bestS, sizeof(uint16_t) * (size_t)D->R_pad * D->n_pad));
    }
    dense_fill_kernel<<<(D->R_pad + 127) / 128, 128, 0, D->stream>>>(D->R_pad, D->bestE, 1e300);
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
}

/* energies of the current states (returned if out_E != NULL) and best-state tracking in one call */
int nlmc_dense_best_update(nlmc_dense *D, double *out_E) {
    using namespace nlmc;
    NLMC_REQUIRE(D && D->bestE, "nlmc_dense_best_update: call nlmc_dense_best_reset first");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    int rc = launch_fields(D, 0, D->n_pad, 1);
    if (rc) return rc;
    dense_energy_kernel<<<(D->R_pad + 127) / 128, 128, 0, D->stream>>>(D->n, D->n_pad, D->R_pad, D->Ht, D->hf, D->S, D->E);
    dense_best_kernel<<<(unsigned)D->R_pad, 128, 0, D->stream>>>(D->n_pad, D->E, D->bestE, D->S, D->bestS);
    NLMC_CUDA(cudaGetLastError());
    if (out_E) {
        NLMC_CUDA(cudaMemcpyAsync(out_E, D->E, sizeof(double) * (size_t)D->R, cudaMemcpyDeviceToHost, D->stream));
        NLMC_CUDA(cudaStreamSynchronize(D->stream));
    }
    return NLMC_OK;
}

int nlmc_dense_best_get(nlmc_dense *D, int8_t *out_spins, double *out_E) {
    using namespace nlmc;
    NLMC_REQUIRE(D && D->bestE, "nlmc_dense_best_get: call nlmc_dense_best_reset first");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    if (out_spins) {
        int8_t *tmp = nullptr;
        const size_t cnt = (size_t)D->R * D->n;
        NLMC_CUDA(cudaMalloc(&tmp, cnt));
        dense_unpack_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, D->stream>>>(D->n, D->n_pad, D->R, D->bestS, tmp);
        cudaError_t e = cudaGetLastError();
        if (e == cudaSuccess) e = cudaMemcpyAsync(out_spins, tmp, cnt, cudaMemcpyDeviceToHost, D->stream);
        if (e == cudaSuccess) e = cudaStreamSynchronize(D->stream);
        cudaFree(tmp);
        if (e != cudaSuccess) { set_error("nlmc_dense_best_get: %s", cudaGetErrorString(e)); return NLMC_ERR_CUDA; }
    }
    if (out_E) {
        NLMC_CUDA(cudaMemcpyAsync(out_E, D->bestE, sizeof(double) * (size_t)D->R, cudaMemcpyDeviceToHost, D->stream));
        NLMC_CUDA(cudaStreamSynchronize(D->stream));
    }
    return NLMC_OK;
}

int nlmc_dense_sync(nlmc_dense *D) {
    NLMC_REQUIRE(D, "nlmc_dense_sync: NULL handle");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    NLMC_CUDA(cudaStreamSynchronize(D->stream));
    return NLMC_OK;
}

/* average duration (ms) of the full field GEMM over `repeats` back-to-back launches, CUDA events on the stream */
int nlmc_dense_time_fields(nlmc_dense *D, int repeats, float *out_ms) {
    using namespace nlmc;
    NLMC_REQUIRE(D && out_ms && repeats >= 1, "nlmc_dense_time_fields: bad arguments");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    int rc = launch_fields(D, 0, D->n_pad, 1);  // warm-up
    if (rc) return rc;
    NLMC_CUDA(cudaEventRecord(D->ev0, D->stream));
    for (int i = 0; i < repeats; ++i)
        if ((rc = launch_fields(D, 0, D->n_pad, 1))) return rc;
    NLMC_CUDA(cudaEventRecord(D->ev1, D->stream));
    NLMC_CUDA(cudaEventSynchronize(D->ev1));
    float ms = 0.f;
    NLMC_CUDA(cudaEventElapsedTime(&ms, D->ev0, D->ev1));
    *out_ms = ms / repeats;
    return NLMC_OK;
}

int nlmc_dense_time_sweeps(nlmc_dense *D, int n_sweeps, float *out_ms) {
    NLMC_REQUIRE(D && out_ms && n_sweeps >= 1, "nlmc_dense_time_sweeps: bad arguments");
    NLMC_CUDA(cudaSetDevice(D->inst->device));
    int rc = nlmc_dense_sweep(D, 1);
    if (rc) return rc;
    NLMC_CUDA(cudaEventRecord(D->ev0, D->stream));
    if ((rc = nlmc_dense_sweep(D, n_sweeps))) return rc;
    NLMC_CUDA(cudaEventRecord(D->ev1, D->stream));
    NLMC_CUDA(cudaEventSynchronize(D->ev1));
    float ms = 0.f;
    NLMC_CUDA(cudaEventElapsedTime(&ms, D->ev0, D->ev1));
    *out_ms = ms / n_sweeps;
    return NLMC_OK;
}

}  // extern "C"
