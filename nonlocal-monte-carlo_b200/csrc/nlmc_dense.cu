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

constexpr int kBM = 128, kBN = 128, kBK = 64, kStages = 3, kMaxSplit = 3;
constexpr int kTileBytes = kBM * kBK * 2;  // one 128 x 64 bf16 tile = 16 KiB (A and each B piece)
constexpr int kGemmThreads = 192;
constexpr size_t kGemmSmem = (size_t)kStages * (1 + kMaxSplit) * kTileBytes + 1024 /*align*/ + 256 /*barriers*/;

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
    float *Ct;
    int accumulate;  // 1: red.add into C^T (split-K or accumulate), 0: plain store
};

__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_bf16_tn_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b0,
                    const __grid_constant__ CUtensorMap map_b1, const __grid_constant__ CUtensorMap map_b2,
                    GemmParams p) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t *smem = reinterpret_cast<uint8_t *>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    const int stage_bytes = (1 + p.n_split) * kTileBytes;
    uint64_t *bars = reinterpret_cast<uint64_t *>(smem + (size_t)kStages * (1 + kMaxSplit) * kTileBytes);
    uint64_t *full = bars, *empty = bars + kStages, *tmem_full = bars + 2 * kStages;
    uint32_t *tmem_ptr = reinterpret_cast<uint32_t *>(bars + 2 * kStages + 1);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int m0 = blockIdx.x * kBM, n0 = p.n_base + blockIdx.y * kBN;
    const int kb_per = (p.k_blocks_total + p.k_splits - 1) / p.k_splits;
    const int kb_begin = blockIdx.z * kb_per, kb_end = min(p.k_blocks_total, kb_begin + kb_per);
    const int n_kb = kb_end - kb_begin;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_a) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&map_b0) : "memory");
    }
    if (warp == 1) {
        if (lane == 0) {
            for (int s = 0; s < kStages; ++s) { mbar_init(full + s, 1); mbar_init(empty + s, 1); }
            mbar_init(tmem_full, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr)), "r"(kBN) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tcgen05_fence_before();
    __syncthreads();
    tcgen05_fence_after();
    const uint32_t tmem_base = *tmem_ptr;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            const CUtensorMap *maps_b[kMaxSplit] = {&map_b0, &map_b1, &map_b2};
            for (int i = 0; i < n_kb; ++i) {
                const int s = i % kStages;
                const uint32_t ph = (uint32_t)(i / kStages) & 1u;
                mbar_wait(empty + s, ph ^ 1u);
                uint8_t *st = smem + (size_t)s * (1 + kMaxSplit) * kTileBytes;
                mbar_expect_tx(full + s, (uint32_t)stage_bytes);
                const int k0 = (kb_begin + i) * kBK;
                tma_load_2d(st, &map_a, full + s, k0, m0);
                for (int q = 0; q < p.n_split; ++q) tma_load_2d(st + (1 + q) * kTileBytes, maps_b[q], full + s, k0, n0);
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer (single thread) =====
            const uint32_t idesc = make_idesc_bf16(kBM, kBN);
            for (int i = 0; i < n_kb; ++i) {
                const int s = i % kStages;
                const uint32_t ph = (uint32_t)(i / kStages) & 1u;
                mbar_wait(full + s, ph);
                tcgen05_fence_after();
                const uint32_t a_addr = smem_u32(smem + (size_t)s * (1 + kMaxSplit) * kTileBytes);
                const uint64_t a_desc = make_smem_desc_sw128(a_addr);
                for (int q = 0; q < p.n_split; ++q) {
                    const uint64_t b_desc = make_smem_desc_sw128(a_addr + (1 + q) * kTileBytes);
#pragma unroll
                    for (int k = 0; k < kBK / 16; ++k)  // UMMA_K = 16 bf16 = 32 bytes inside the swizzle atom
                        umma_bf16(tmem_base, a_desc + (uint64_t)(k * 2), b_desc + (uint64_t)(k * 2), idesc,
                                  (uint32_t)((i | q | k) != 0));
                }
                umma_commit(empty + s);  // smem stage reusable once these MMAs have read it
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
        for (int c = 0; c < kBN; c += 32) {
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
    if (warp == 1) {
        tcgen05_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kBN) : "memory");
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
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
};

namespace nlmc {

// fields of columns [col0, col0 + n_cols) for all replicas: Ht[col][r] = sum_k S[r][k] J[col][k]
static int launch_fields(nlmc_dense *D, int col0, int n_cols, int k_splits, bool clear = true) {
    NLMC_REQUIRE(col0 % kBN == 0, "launch_fields: col0 must be a multiple of %d", kBN);
    GemmParams p;
    p.M = D->R_pad;
    p.N = std::min(D->n_pad, col0 + n_cols);
    p.n_base = col0;
    p.k_blocks_total = D->n_pad / kBK;
    p.k_splits = k_splits;
    p.n_split = D->n_split;
    p.ldc = D->R_pad;
    p.Ct = D->Ht;
    p.accumulate = k_splits > 1;
    if (p.accumulate && clear)
        NLMC_CUDA(cudaMemsetAsync(D->Ht + (size_t)col0 * D->R_pad, 0, sizeof(float) * (size_t)(p.N - col0) * D->R_pad, D->stream));
    const dim3 grid((unsigned)(D->R_pad / kBM), (unsigned)((n_cols + kBN - 1) / kBN), (unsigned)k_splits);
    gemm_bf16_tn_kernel<<<grid, kGemmThreads, kGemmSmem, D->stream>>>(
        D->map_S, D->map_J[0], D->map_J[D->n_split > 1 ? 1 : 0], D->map_J[D->n_split > 2 ? 2 : 0], p);
    NLMC_CUDA(cudaGetLastError());
    return NLMC_OK;
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

// Sequential heat-bath update of the sites [c0, c0+kBlk) for kRepPerCta replicas per CTA.
// field_k = Ht[c0+k][r] (from the GEMM, spins as of the start of the block) + h_k
//           + sum_{j<k} J[c0+k][c0+j] * (s_j^new - s_j^old)           (flips made earlier in this block)
// s_k <- +1 with probability 1/(1+exp(-2 beta_r field_k))  == sign(tanh(beta x) - 2u + 1), NMC/nmc.py:87
//
// The chain over k is sequential per replica and there are only R chains, so what matters is the latency
// of one step (ncu on the first versions: ~1000 cycles/step, one warp per scheduler, every stall exposed).
// The block is therefore processed in sub-blocks of 8 sites.  Each replica owns 8 lanes:
//   1. the 8 fields of the sub-block and its 28 in-sub-block couplings are read by all 8 lanes (broadcast);
//      the 8 decisions are then taken one after the other entirely in registers (no shuffles, no stores);
//   2. the 8 flips are propagated to the fields of the later sites of the block, lane t taking the sites
//      k' = 8m + t (conflict-free reads of the transposed J_bb, 8 FMAs each).
constexpr int kRepPerCta = 16;
constexpr int kFldStride = kBlk + 8;    // floats per replica row: the 4 replicas of a warp fall into disjoint bank groups
constexpr int kSpinStride = kBlk + 32;  // bytes per replica row, same reason
constexpr size_t kUpdateSmem = sizeof(float) * ((size_t)kBlk * kBlk + (size_t)kRepPerCta * kFldStride) + (size_t)kRepPerCta * kSpinStride;

__global__ void __launch_bounds__(128) dense_block_update_kernel(int n, int n_pad, int R_pad, int c0, const float *__restrict__ Ht,
                                                                 const float *__restrict__ Jf, const float *__restrict__ hf,
                                                                 const float *__restrict__ beta, uint16_t *S,
                                                                 uint32_t seed_lo, uint32_t seed_hi,
                                                                 const uint32_t *__restrict__ sweep_ptr, float *Ht_zero,
                                                                 const uint8_t *__restrict__ modes, float temp_x) {
    extern __shared__ __align__(16) uint8_t dsm[];
    const uint32_t sweep = *sweep_ptr;
    float *Jt = reinterpret_cast<float *>(dsm);                        // [kBlk j][kBlk k] = J[c0+k][c0+j] (transposed)
    float *fld = Jt + (size_t)kBlk * kBlk;                             // [kRepPerCta][kBlk]   running fields
    int8_t *spin = reinterpret_cast<int8_t *>(fld + (size_t)kRepPerCta * kFldStride);  // [kRepPerCta][kSpinStride]
    const int tid = threadIdx.x;
    const int rep = tid >> 3, t = tid & 7;     // replica within the CTA, lane within the replica's group
    const int r0 = blockIdx.x * kRepPerCta;
    const int r = r0 + rep;
    // stage J_bb transposed, Jt[j][k] = J[c0+k][c0+j], from the transposed copy of J kept in global memory
    // (JfT[a][b] = J[b][a]): coalesced float4 reads, conflict-free float4 writes
    // asynchronous 16-byte copies (no register staging): all 64 KB are in flight at once and overlap the field/spin
    // loads below; ncu showed 40 % of this kernel's time in the prologue's load latency
    for (int i = tid; i < kBlk * kBlk / 4; i += 128) {
        const int j = i / (kBlk / 4), k4 = i % (kBlk / 4);
        const uint32_t dst = smem_u32(reinterpret_cast<float4 *>(Jt) + i);
        const float *src = Jf + (size_t)(c0 + j) * n_pad + c0 + k4 * 4;
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
    {   // all 16 field / spin loads of a thread are issued before the first use (one exposed round trip, not 16)
        constexpr int kPer = kRepPerCta * kBlk / 128;
        float hv[kPer], fv[kPer];
        uint16_t sv[kPer];
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const int i = tid + q * 128;
            const int k = i / kRepPerCta, rr = i % kRepPerCta;  // consecutive threads -> consecutive replicas (coalesced Ht row)
            hv[q] = Ht[(size_t)(c0 + k) * R_pad + r0 + rr];
            fv[q] = hf[c0 + k];
            sv[q] = S[(size_t)(r0 + rr) * n_pad + c0 + k];
        }
#pragma unroll
        for (int q = 0; q < kPer; ++q) {
            const int i = tid + q * 128;
            const int k = i / kRepPerCta, rr = i % kRepPerCta;
            fld[rr * kFldStride + k] = hv[q] + fv[q];
            spin[rr * kSpinStride + k] = (sv[q] == 0) ? 0 : ((sv[q] & 0x8000u) ? -1 : 1);
        }
    }
    // the next block's split-K GEMM accumulates with atomics: clear its field rows for this CTA's replicas
    if (Ht_zero != nullptr)
        for (int i = tid; i < kRepPerCta * kBlk; i += 128) Ht_zero[(size_t)(i / kRepPerCta) * R_pad + r0 + (i % kRepPerCta)] = 0.f;
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();
    const float inv2b = 0.5f / beta[r];
    const uint8_t *mrow = modes ? modes + (size_t)r * n_pad + c0 : nullptr;
    const PhiloxD rng{seed_lo, seed_hi ^ 0x44454e53u};
    const int k_end = min(kBlk, n - c0);
    float *frow = fld + rep * kFldStride;
    int8_t *srow = spin + rep * kSpinStride;
    for (int base = 0; base < k_end; base += 8) {
        // ---- 1. the sub-block, sequentially, in registers (identical in the 8 lanes of the replica).
        // up  <=>  u < 1/(1+exp(-2 beta f))  <=>  f > logit(u)/(2 beta) =: theta, which depends on the random
        // number only: exp/log stay off the sequential chain, a decision is compare + select + FMA.
        float F[8], Jss[28], d[8], theta[8], dp[8], dm[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            F[i] = frow[base + i];
            const float so = (float)srow[base + i];
            dp[i] = 1.0f - so;    // delta if the site ends up +1
            dm[i] = -1.0f - so;   // delta if it ends up -1
        }
#pragma unroll
        for (int i = 1, q = 0; i < 8; ++i)
#pragma unroll
            for (int j = 0; j < i; ++j, ++q) Jss[q] = Jt[(base + j) * kBlk + base + i];  // J[base+i][base+j]
        const uint4 ra = rng((uint32_t)r, (uint32_t)(c0 + base), sweep, 0u);
        const uint4 rb = rng((uint32_t)r, (uint32_t)(c0 + base), sweep, 1u);
        const uint32_t ub[8] = {ra.x, ra.y, ra.z, ra.w, rb.x, rb.y, rb.z, rb.w};
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            // logit of a uniform in (0,1) from all 32 bits, evaluated on the side of the nearer tail so that both tails
            // are symmetric and reach 2^-33 (a 24-bit uniform rounded to 1.0 forced the spin down once in 2^24 draws)
            const bool hi = (ub[i] >> 31) != 0u;
            const uint32_t m = hi ? ~ub[i] : ub[i];
            const float v = ((float)m + 0.5f) * (1.0f / 4294967296.0f);       // in (0, 1/2]
            const float t = __logf(v) - __logf(1.0f - v);
            theta[i] = (hi ? -t : t) * inv2b;
        }
        uint32_t frozen = 0;  // NMC phases: frozen sites keep their spin (the reference pins them with h = +-1e4)
        if (mrow != nullptr) {
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const uint8_t md = mrow[base + i];
                if (md == 1) theta[i] *= temp_x;
                frozen |= (uint32_t)(md == 2) << i;
            }
        }
        uint32_t newbits = 0;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const bool live = (base + i < k_end) && !((frozen >> i) & 1u);
            const bool up = live ? (F[i] > theta[i]) : (dp[i] == 0.0f);  // not live: keep the old spin
            d[i] = live ? (up ? dp[i] : dm[i]) : 0.0f;
            newbits |= (uint32_t)up << i;
#pragma unroll
            for (int i2 = i + 1; i2 < 8; ++i2) F[i2] = fmaf(Jss[i2 * (i2 - 1) / 2 + i], d[i], F[i2]);  // right-looking
        }
        if (t == 0) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
                if (base + i < k_end) srow[base + i] = (int8_t)(((newbits >> i) & 1u) ? 1 : -1);
        }
        // ---- 2. propagate the 8 flips to the later sites of the block: lane t owns k' = 8m + t
        for (int kp = base + 8 + t; kp < k_end; kp += 8) {  // two independent FMA chains: half the dependent latency
            float a = frow[kp], b = 0.0f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                a = fmaf(Jt[(base + j) * kBlk + kp], d[j], a);
                b = fmaf(Jt[(base + 4 + j) * kBlk + kp], d[4 + j], b);
            }
            frow[kp] = a + b;
        }
        __syncwarp();
    }
    __syncthreads();
    // write the block segments back as bf16 (+1 = 0x3F80, -1 = 0xBF80): thread -> (replica, 16 consecutive sites)
    {
        const int seg = t * 16;
        uint32_t w[8];
#pragma unroll
        for (int e = 0; e < 16; e += 2) {
            uint32_t pair = 0;
#pragma unroll
            for (int h2 = 0; h2 < 2; ++h2) {
                const int k = seg + e + h2;
                const int sv = k < k_end ? spin[rep * kSpinStride + k] : 0;
                pair |= (sv > 0 ? 0x3F80u : (sv < 0 ? 0xBF80u : 0u)) << (16 * h2);
            }
            w[e >> 1] = pair;
        }
        uint4 *dst = reinterpret_cast<uint4 *>(S + (size_t)r * n_pad + c0 + seg);
        dst[0] = make_uint4(w[0], w[1], w[2], w[3]);
        dst[1] = make_uint4(w[4], w[5], w[6], w[7]);
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
    void *ptrs[] = {D->S, D->Jp[0], D->Jp[1], D->Jp[2], D->Jf, D->hf, D->Ht, D->beta, D->E, D->d_sweep, D->bestE, D->bestS, D->modes};
    for (void *p : ptrs) if (p) cudaFree(p);
    D->xch.release();
    if (D->sweep_graph) cudaGraphExecDestroy(D->sweep_graph);
    if (D->ev0) cudaEventDestroy(D->ev0);
    if (D->ev1) cudaEventDestroy(D->ev1);
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
    NLMC_REQUIRE(I->value_symmetric, "nlmc_dense_create: J must be symmetric (J_ij == J_ji) without repeated entries");
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
        if (cudaFuncSetAttribute(gemm_bf16_tn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kGemmSmem) != cudaSuccess ||
            cudaFuncSetAttribute(dense_block_update_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kUpdateSmem) != cudaSuccess) {
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

static int enqueue_sweep(nlmc_dense *D, int k_splits) {
    using namespace nlmc;
    for (int c0 = 0; c0 < D->n; c0 += kBlk) {
        int rc = launch_fields(D, c0, kBlk, k_splits, /*clear=*/c0 == 0);  // later blocks are cleared by the update kernel
        if (rc) return rc;
        float *zero_next = (k_splits > 1 && c0 + kBlk < D->n) ? D->Ht + (size_t)(c0 + kBlk) * D->R_pad : nullptr;
        dense_block_update_kernel<<<(unsigned)(D->R_pad / kRepPerCta), 128, kUpdateSmem, D->stream>>>(
            D->n, D->n_pad, D->R_pad, c0, D->Ht, D->Jf, D->hf, D->beta, D->S, (uint32_t)D->seed, (uint32_t)(D->seed >> 32),
            D->d_sweep, zero_next, D->modes_on ? D->modes : nullptr, D->temp_x);
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
