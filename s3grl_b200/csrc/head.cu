// Kernel 7 — fused scoring head of SIGNNet for the fixed-row flows: the one
// GEMM-shaped piece next to the hot path, on the 5th-generation tensor cores.
//
// Replaces, in evaluation mode, reference models.py:370-376 + :339-346:
//      xs_cat = torch.cat(xs, -1)                       (the loader's joint matrix, loader.cu)
//      h      = operator_diff(xs_cat)                   Linear((K+1)F', 256) -> ELU -> BatchNorm (eval) [-> dropout = id]
//      h      = h[center] * h[center + 1]               _centre_pool_helper, k_heuristic = 0 (rows 2i, 2i+1 of link i)
// as ONE kernel: pooled[i, :] = bn(elu(X[2i] W^T + b)) * bn(elu(X[2i+1] W^T + b)).  The remaining
// link_pred_mlp works on [L, 256] and stays with the caller.
//
// Persistent CTAs (one per SM) over 256-row tiles of X, N = 256 output channels, TF32 inputs straight from the fp32 joint matrix
// (no conversion pass), fp32 accumulation in TMEM:
//   warp 0    TMA producer: cp.async.bulk.tensor 2-D boxes, two [128 x 32] of X and one [256 x 32] of W (128-byte
//             swizzle), 3-stage mbarrier ring (64 KB per stage); the K tail is zero-filled by TMA
//   warp 1    allocates all 512 TMEM columns (two accumulators), one lane issues tcgen05.mma.kind::tf32 (M128 N256 K8,
//             2 x 4 per stage: both row halves share the weight stage), tcgen05.commit releases the stage / signals the accumulator
//   warps 2-9 epilogue (the TMA loads of the next tile run ahead meanwhile; its MMAs wait for acc_empty): tcgen05.ld 32x32b (one accumulator row
//             per thread), bias + ELU + BN affine in registers, the two rows of a link meet by one shuffle,
//             128-byte vector stores of the product
// The GEMM is HBM-bound on X (4*(K+1)F' bytes per row against 2*256*(K+1)F' flops): tensor cores are what
// keeps the math under the copy time.
#include <cuda.h>

#include "common.cuh"

namespace s3 {
namespace {

constexpr int kHeadN = 256;        // hidden channels (every paper config: hidden_channels = 256)
constexpr int kHeadM = 128;        // rows per MMA (UMMA_M)
constexpr int kHeadTileM = 256;    // rows per CTA tile: two MMAs share every [256 x 32] weight stage
constexpr int kHeadKB = 32;        // tf32 elements per stage along K = one 128-byte swizzle atom
constexpr int kHeadStages = 3;
constexpr int kHeadEpiWarps = 8;
constexpr int kHeadThreads = 64 + 32 * kHeadEpiWarps;
constexpr uint32_t kABytes = kHeadTileM * kHeadKB * 4;  // 32 KB: rows 0-127 | rows 128-255
constexpr uint32_t kBBytes = kHeadN * kHeadKB * 4;  // 32 KB
constexpr uint32_t kStageBytes = kABytes + kBBytes;
constexpr size_t kHeadSmem = (size_t)kHeadStages * kStageBytes + 1024;

struct HeadParams {
    const float* __restrict__ bias;
    const float* __restrict__ scale;
    const float* __restrict__ shift;
    float* __restrict__ pooled;  // [rows / 2, 256], or [rows, 256] when pool == 0
    int pool;
    int64_t rows;
    int num_kb;
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done;
    do {
        asm volatile(
            "{\n\t"
            ".reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t"
            "}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
    } while (!done);
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, uint32_t dst, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
        "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
        : "memory");
}
// K-major operand tile with 128-byte swizzle: rows of 128 B, 8-row groups 1024 B apart (SBO), LBO unused (1),
// descriptor version 1 (sm_100), layout type 2 = SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | (1ull << 16) | (64ull << 32) | (1ull << 46) | (2ull << 61);
}
__device__ __forceinline__ void umma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}" ::"r"(tmem_d), "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// Persistent: one CTA per SM walks the 256-row tiles blockIdx.x, blockIdx.x + gridDim.x, ... Each tile is two
// M = 128 accumulators (2 x 256 TMEM columns = all of it) fed from the SAME weight stage, so the [256 x 32] weight
// tile is read from L2 once per 256 rows (the kernel is L2-throughput bound on those re-reads; measured).
__global__ void __launch_bounds__(kHeadThreads, 1)
sign_head_kernel(const __grid_constant__ CUtensorMap map_x, const __grid_constant__ CUtensorMap map_w, HeadParams p) {
    extern __shared__ uint8_t smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[kHeadStages], empty_bar[kHeadStages], acc_full, acc_empty;
    __shared__ uint32_t s_tmem;
    __shared__ float s_bias[kHeadN], s_scale[kHeadN], s_shift[kHeadN];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const uint32_t tiles = (smem_u32(smem_raw) + 1023u) & ~1023u;  // the swizzle pattern is a function of the address
    const int num_tiles = (int)((p.rows + kHeadTileM - 1) / kHeadTileM);

    for (int i = threadIdx.x; i < kHeadN; i += kHeadThreads) {
        s_bias[i] = p.bias[i];
        s_scale[i] = p.scale[i];
        s_shift[i] = p.shift[i];
    }
    if (threadIdx.x == 0) {
        for (int s = 0; s < kHeadStages; ++s) {
            mbar_init(&full_bar[s], 1);
            mbar_init(&empty_bar[s], 1);
        }
        mbar_init(&acc_full, 1);
        mbar_init(&acc_empty, kHeadEpiWarps);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem)), "n"(2 * kHeadN));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem = s_tmem;

    if (warp == 0) {
        if (lane == 0) {  // ===== TMA producer =====
            int it = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
                for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                    const int s = it % kHeadStages;
                    const uint32_t ph = (it / kHeadStages) & 1;
                    mbar_wait(&empty_bar[s], ph ^ 1);
                    mbar_expect_tx(&full_bar[s], kStageBytes);
                    const uint32_t a = tiles + s * kStageBytes;
                    tma_load_2d(&map_x, &full_bar[s], a, kb * kHeadKB, tile * kHeadTileM);
                    tma_load_2d(&map_w, &full_bar[s], a + kABytes, kb * kHeadKB, 0);
                }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {  // ===== MMA issuer =====
            // instruction descriptor: D = F32 (bits 4-5 = 1), A = B = TF32 (bits 7-9, 10-12 = 2), both K-major, N >> 3 at 17, M >> 4 at 24
            constexpr uint32_t idesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(kHeadN >> 3) << 17) | ((uint32_t)(kHeadM >> 4) << 24);
            int it = 0, t = 0;
            for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
                mbar_wait(&acc_empty, (t & 1) ^ 1);  // the epilogue has drained both accumulators
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                for (int kb = 0; kb < p.num_kb; ++kb, ++it) {
                    const int s = it % kHeadStages;
                    const uint32_t ph = (it / kHeadStages) & 1;
                    mbar_wait(&full_bar[s], ph);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    const uint32_t a = tiles + s * kStageBytes;
#pragma unroll
                    for (int k = 0; k < kHeadKB / 8; ++k) {  // UMMA_K = 8 tf32 = 32 bytes inside the swizzle atom
                        const uint64_t db = umma_desc(a + kABytes + k * 32);
                        umma_tf32(tmem, umma_desc(a + k * 32), db, idesc, (kb | k) != 0);
                        umma_tf32(tmem + kHeadN, umma_desc(a + kABytes / 2 + k * 32), db, idesc, (kb | k) != 0);
                    }
                    umma_commit(&empty_bar[s]);  // the stage is free once these MMAs have read it
                }
                umma_commit(&acc_full);  // both accumulators of this tile complete
            }
        }
    } else {
        // ===== epilogue: 8 warps; warp w owns TMEM lanes 32*(w % 4) .. +32 of accumulator (w - 2) / 4 =====
        const int q = warp & 3, half = (warp - 2) >> 2;
        int t = 0;
        for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++t) {
            const int64_t row = (int64_t)tile * kHeadTileM + half * kHeadM + q * 32 + lane;
            mbar_wait(&acc_full, t & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            float* out = p.pooled + (p.pool ? (row >> 1) : row) * kHeadN;
            for (int c0 = 0; c0 < kHeadN; c0 += 32) {
                uint32_t r[32];
                const uint32_t taddr = tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * kHeadN + c0);
                asm volatile(
                    "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
                    "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
                    "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
                    : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
                      "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
                      "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
                      "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
                    : "r"(taddr));
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) {
                    float z = __uint_as_float(r[i]) + s_bias[c0 + i];
                    z = z > 0.0f ? z : __expf(z) - 1.0f;             // ELU (alpha = 1)
                    z = fmaf(z, s_scale[c0 + i], s_shift[c0 + i]);   // BatchNorm1d in eval mode as an affine map
                    const float other = __shfl_xor_sync(0xffffffffu, z, 1);
                    v[i] = p.pool ? z * other : z;                   // h_src * h_dst: rows 2i and 2i+1 sit in adjacent lanes
                }
                if ((!(lane & 1) || !p.pool) && row < p.rows) {
#pragma unroll
                    for (int i = 0; i < 32; i += 4)
                        *reinterpret_cast<float4*>(out + c0 + i) = make_float4(v[i], v[i + 1], v[i + 2], v[i + 3]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&acc_empty);  // this warp's part of the accumulators may be overwritten
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "n"(2 * kHeadN));
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

cudaError_t make_map(EncodeTiledFn enc, CUtensorMap* map, const float* base, int64_t rows, int64_t cols, int64_t ld, int box_rows) {
    cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    cuuint64_t strides[1] = {(cuuint64_t)ld * 4};
    cuuint32_t box[2] = {(cuuint32_t)kHeadKB, (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS ? cudaSuccess : cudaErrorInvalidValue;
}

}  // namespace

cudaError_t launch_sign_head(const float* x, int64_t rows, int64_t kdim, int64_t ldx, const float* w, int64_t ldw,
                             const float* bias, const float* scale, const float* shift, float* pooled, int pool, cudaStream_t st) {
    if (rows == 0) return cudaSuccess;
    static EncodeTiledFn enc = nullptr;
    if (!enc) {
        void* fn = nullptr;
        cudaDriverEntryPointQueryResult qres;
        cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
        if (e != cudaSuccess) return e;
        if (qres != cudaDriverEntryPointSuccess || !fn) return cudaErrorNotSupported;
        enc = reinterpret_cast<EncodeTiledFn>(fn);
    }
    static LaunchCache cache;  // per device: shared-memory opt-in and SM count
    int sms = 0;
    cudaError_t e = cache.get(reinterpret_cast<const void*>(sign_head_kernel), kHeadThreads, kHeadSmem, &sms, nullptr);
    if (e != cudaSuccess) return e;
    CUtensorMap map_x, map_w;
    e = make_map(enc, &map_x, x, rows, kdim, ldx, kHeadTileM);
    if (e != cudaSuccess) return e;
    e = make_map(enc, &map_w, w, kHeadN, kdim, ldw, kHeadN);
    if (e != cudaSuccess) return e;
    HeadParams p;
    p.bias = bias;
    p.scale = scale;
    p.shift = shift;
    p.pooled = pooled;
    p.pool = pool;
    p.rows = rows;
    p.num_kb = (int)((kdim + kHeadKB - 1) / kHeadKB);
    const int64_t num_tiles = (rows + kHeadTileM - 1) / kHeadTileM;
    sign_head_kernel<<<(unsigned)(num_tiles < sms ? num_tiles : sms), kHeadThreads, kHeadSmem, st>>>(map_x, map_w, p);
    return cudaGetLastError();
}

}  // namespace s3
