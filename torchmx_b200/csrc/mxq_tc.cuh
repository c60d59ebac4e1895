// Shared device / host helpers of the tcgen05 MX GEMM kernels (K3 family): PTX wrappers (mbarrier, TMA, tcgen05, cluster),
// shared-memory / instruction descriptors, the scale-factor loader warps and the tensor-map builders.
#pragma once
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "mxq_common.cuh"

namespace mxq {
namespace gemm {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_K = 128;  // bytes == elements (1-byte codes): one 128B swizzle row, 4 MX blocks
constexpr int UMMA_K = 32;
constexpr int kThreads = 256;
constexpr int kEpilogueThreads = 128;

// ---- PTX wrappers -------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n\t.reg .pred p;\n\t.reg .b32 r;\n\t"
        "elect.sync r|p, 0xFFFFFFFF;\n\t"
        "selp.b32 %0, 1, 0, p;\n\t}"
        : "=r"(pred));
    return pred != 0;
}

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

__device__ __forceinline__ void tma_load_3d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
// shared -> global tile store (bulk async group of the issuing thread)
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, const void* src, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%2, %3, %4}], [%1];" ::"l"(map), "r"(smem_u32(src)), "r"(c0), "r"(c1),
                 "r"(c2)
                 : "memory");
}
__device__ __forceinline__ void tma_store_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N>
__device__ __forceinline__ void tma_store_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void tma_load_4d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(const CUtensorMap* map, uint64_t* bar, void* dst, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(smem_u32(dst)),
                 "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
                 : "memory");
}
// pull a box from DRAM into L2 only (no shared-memory destination, nothing to wait for)
__device__ __forceinline__ void tma_prefetch_l2_3d(const CUtensorMap* map, int c0, int c1, int c2) {
    asm volatile("cp.async.bulk.prefetch.tensor.3d.L2.global [%0, {%1, %2, %3}];" ::"l"(map), "r"(c0), "r"(c1), "r"(c2) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}

template <int COLS>
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t addr) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}

// smem -> TMEM, 32 rows x 128 bit, replicated to the four 32-lane quadrants (scale factors)
__device__ __forceinline__ void tc_copy_sf(uint32_t tmem_addr, uint64_t smem_desc) {
    asm volatile("tcgen05.cp.cta_group::1.32x128b.warpx4 [%0], %1;" ::"r"(tmem_addr), "l"(smem_desc) : "memory");
}
__device__ __forceinline__ void tc_mma_mx(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                          uint32_t tmem_sfa, uint32_t tmem_sfb) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld_32x32b_x32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]),
          "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
          "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]),
          "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// ---- NVLink multicast (NVLS): add into the same address of every GPU mapped behind a multicast pointer ------------
__device__ __forceinline__ void multimem_red_add_bf16x8(void* mc_addr, uint4 v) {
    asm volatile("multimem.red.relaxed.sys.global.add.v4.bf16x2 [%0], {%1, %2, %3, %4};" ::"l"(mc_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ void multimem_red_add_bf16x2(void* mc_addr, uint32_t v) {
    asm volatile("multimem.red.relaxed.sys.global.add.bf16x2 [%0], %1;" ::"l"(mc_addr), "r"(v) : "memory");
}

// ---- cluster / cta_group::2 flavours ------------------------------------------------------------------
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
// address of the same shared-memory object in CTA `rank` of the cluster (shared::cluster window)
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, uint32_t rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
    // default semantics (release at CTA scope), as cutlass::arch::ClusterBarrier::arrive(cta_id): a cluster-scope release costs a
    // full membar that also waits for the loader warps' in-flight global prefetches
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_LOOP:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE;\n\t"
        "bra WAIT_LOOP;\n\t"
        "DONE:\n\t}" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// TMA load whose completion bytes are posted on an mbarrier of either CTA of the pair (cluster address)
__device__ __forceinline__ void tma_load_3d_pair(const CUtensorMap* map, uint32_t bar_cluster_addr, void* dst, int c0, int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(bar_cluster_addr), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)), "n"(COLS) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
template <int COLS>
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t addr) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(addr), "n"(COLS) : "memory");
}
__device__ __forceinline__ void tc_copy_sf_pair(uint32_t tmem_addr, uint64_t smem_desc) {
    asm volatile("tcgen05.cp.cta_group::2.32x128b.warpx4 [%0], %1;" ::"r"(tmem_addr), "l"(smem_desc) : "memory");
}
__device__ __forceinline__ void tc_mma_mx_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                               uint32_t tmem_sfa, uint32_t tmem_sfb) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::mxf8f6f4.block_scale [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}
// fp4 x fp4 at twice the rate: kind::mxf4, K = 64 per instruction, operands are the DENSE 4-bit streams (128 bytes of a 128B
// swizzle row = 256 elements), one E8M0 scale per 32 elements = two scale bytes per row and instruction (scale_vec::2X): the
// scale-factor id in the descriptor selects byte pair 0 or 2 of the row's 32-bit scale word in TMEM
__device__ __forceinline__ void tc_mma_mxf4_pair(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate,
                                                 uint32_t tmem_sfa, uint32_t tmem_sfb) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::mxf4.block_scale.scale_vec::2X [%0], %1, %2, %3, [%5], [%6], p;\n\t}" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(tmem_sfa), "r"(tmem_sfb)
        : "memory");
}
// arrive on the barrier at the same shared-memory offset in both CTAs of the pair once all prior MMAs retire
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
                 "h"((uint16_t)3)
                 : "memory");
}

// ---- descriptors ------------------------------------------------------------------------------------
// shared-memory matrix descriptor (sm_100 format: version 1 in bits [46,48))
__device__ __forceinline__ uint64_t smem_desc(uint32_t smem_addr, uint32_t sbo_bytes, uint32_t layout_type) {
    return (uint64_t)((smem_addr >> 4) & 0x3FFF) | ((uint64_t)(sbo_bytes >> 4) << 32) | ((uint64_t)1 << 46) | ((uint64_t)layout_type << 61);
}
constexpr uint32_t kLayoutNone = 0, kLayoutSw128 = 2;

// instruction descriptor for kind::mxf8f6f4.block_scale: E4M3 x E4M3 (K-major both), UE8M0 scales, dense K=32
__host__ __device__ constexpr uint32_t make_idesc(int m, int n) {
    return (0u << 7) | (0u << 10) | (0u << 15) | (0u << 16) | ((uint32_t)(n >> 3) << 17) | (1u << 23) | ((uint32_t)(m >> 4) << 24);
}
// operand storage format (MXQ_OPERAND_*) -> element format field of the descriptor (a_format bits [7,10), b_format bits
// [10,13)): E4M3 = 0, E2M3 = 3, E3M2 = 4, E2M1 = 5; and -> bits per element as TMA counts them in complete_tx
__host__ __device__ constexpr uint32_t umma_format(int operand_format) {
    return operand_format == MXQ_OPERAND_E2M1_PACKED ? 5u
           : (operand_format == MXQ_OPERAND_E3M2_PACKED ? 4u : (operand_format == MXQ_OPERAND_E2M3_PACKED ? 3u : (operand_format == MXQ_OPERAND_E5M2_BYTES ? 1u : 0u)));
}
__host__ __device__ constexpr int operand_bits(int operand_format) {
    return operand_format == MXQ_OPERAND_E2M1_PACKED ? 4 : ((operand_format == MXQ_OPERAND_E4M3_BYTES || operand_format == MXQ_OPERAND_E5M2_BYTES) ? 8 : 6);
}
__host__ __device__ constexpr uint32_t idesc_formats(int a_format, int b_format) { return (umma_format(a_format) << 7) | (umma_format(b_format) << 10); }
constexpr uint32_t kIdescMxf4Formats = (1u << 7) | (1u << 10);  // kind::mxf4: E2M1 = 1 for both operands
__device__ __forceinline__ uint32_t idesc_with_sf(uint32_t idesc, uint32_t sfa_id, uint32_t sfb_id) {
    return idesc | (sfb_id << 4) | (sfa_id << 29);
}

struct Params {
    const uint8_t* sfa; const uint8_t* sfb; const uint16_t* bias; uint16_t* d;
    uint16_t* d_mc;  // multicast alias of the output on every rank: the epilogue adds (multimem.red) instead of storing
    int64_t ld_sfa, ld_sfb, sfa_batch, sfb_batch, ldd, d_batch;
    int M, N, K, batch, m_blocks, n_blocks;
    uint32_t idesc_fmt;  // element-format bits of the instruction descriptor (idesc_formats)
    uint32_t tx_a, tx_b; // bytes one 128-row x 128-element box of A / B posts on the mbarrier (packed formats count packed bytes)
#ifdef MXQ_DEV  // developer builds only (python -m torchmx_b200.build with MXQ_DEV=1): never in the shipped library
    int dbg;           // MXQ_GEMM_DBG: bit0 = epilogue skips the global stores
    long long* trace;  // MXQ_GEMM_TRACE=<device pointer>: clock64 stamps of pair 0's leader, 8 slots per tile
#endif
};
#ifdef MXQ_DEV
#define MXQ_DEV_ONLY(...) __VA_ARGS__
#else
#define MXQ_DEV_ONLY(...)
#endif


// ---- scale-factor loader (one warp, one output tile) ---------------------------------------------------
// The reference keeps scales as [rows, K/32] bytes, i.e. one 32-bit word per (row, 128-wide K block).
// tcgen05.cp.32x128b.warpx4 wants, per 128-row group and K block, 32 chunks of 16 B: chunk i = the words of
// rows i, i+32, i+64, i+96.  Lane i therefore owns those four rows of each of the GROUPS 128-row groups; it
// fetches 16 B per row (four K blocks) per load, two load groups (8 K blocks) ahead of use, and emits one
// 16-byte shared-memory store per group and K block.  `arrive(stage)` publishes the stage (fence + mbarrier).
template <int GROUPS, int STAGES, typename Arrive>
__device__ __forceinline__ void sf_load_tile(const uint8_t* base, int64_t ld, int row0, int row_lim, int k_blocks, uint8_t* sf_smem,
                                             int sf_stage_bytes, uint64_t* empty, uint32_t& stage, uint32_t& phase, int lane, Arrive&& arrive) {
    constexpr int KB_PER_LOAD = 4;
    const uint8_t* rows[GROUPS][4];
#pragma unroll
    for (int g = 0; g < GROUPS; ++g)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = row0 + g * 128 + q * 32 + lane;
            const int rc = r < row_lim ? r : row_lim - 1;  // clamp: rows past the edge only feed masked outputs
            rows[g][q] = base + (int64_t)rc * ld;
        }
    const int n_loads = (k_blocks + KB_PER_LOAD - 1) / KB_PER_LOAD;
    const bool vec_ok = (ld % 16 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && (k_blocks % KB_PER_LOAD == 0);
    uint4 buf[3][GROUPS][4];
    auto issue = [&](int l, uint4 (&dst)[GROUPS][4]) {
        if (l >= n_loads) return;
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint8_t* src = rows[g][q] + 16 * l;
                if (vec_ok) {
                    dst[g][q] = *reinterpret_cast<const uint4*>(src);
                } else {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) w[j] = (l * KB_PER_LOAD + j < k_blocks) ? *reinterpret_cast<const uint32_t*>(src + 4 * j) : 0u;
                    dst[g][q] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    };
    auto word = [](const uint4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); };
    issue(0, buf[0]);
    issue(1, buf[1]);
    for (int l0 = 0; l0 < n_loads; l0 += 3) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int l = l0 + u;
            if (l < n_loads) {
                issue(l + 2, buf[(u + 2) % 3]);
#pragma unroll
                for (int j = 0; j < KB_PER_LOAD; ++j) {
                    if (l * KB_PER_LOAD + j < k_blocks) {
                        mbar_wait(&empty[stage], phase ^ 1);
                        uint8_t* dst = sf_smem + stage * sf_stage_bytes;
#pragma unroll
                        for (int g = 0; g < GROUPS; ++g)
                            *reinterpret_cast<uint4*>(dst + 512 * g + 16 * lane) =
                                make_uint4(word(buf[u][g][0], j), word(buf[u][g][1], j), word(buf[u][g][2], j), word(buf[u][g][3], j));
                        arrive(stage);
                        if (++stage == STAGES) { stage = 0; phase ^= 1; }
                    }
                }
            }
        }
    }
}

// scale-factor ring shared by the pair and skinny kernels: stages x K blocks per stage
constexpr int SF_STAGES = 4, SF_KB = 4;

// one warp, one tile: rows row0 + g*128 + q*32 + lane; SF ring stage = 4 K blocks (one 16-byte load per row)
template <int GROUPS, typename Arrive>
__device__ __forceinline__ void sf_load_tile4(const uint8_t* base, int64_t ld, int row0, int row_lim, int k_blocks, uint8_t* sf_smem, int sf_kb_bytes,
                                              uint64_t* sf_empty, uint32_t& sfs, uint32_t& sf_phase, int lane, Arrive&& arrive) {
    const uint8_t* rows[GROUPS][4];
#pragma unroll
    for (int g = 0; g < GROUPS; ++g)
#pragma unroll
        for (int q = 0; q < 4; ++q) {
            const int r = row0 + g * 128 + q * 32 + lane;
            const int rc = r < row_lim ? r : row_lim - 1;  // clamp: rows past the edge only feed masked outputs
            rows[g][q] = base + (int64_t)rc * ld;
        }
    const int n_loads = (k_blocks + SF_KB - 1) / SF_KB;
    const bool vec_ok = (ld % 16 == 0) && ((reinterpret_cast<uintptr_t>(base) & 15) == 0) && (k_blocks % SF_KB == 0);
    uint4 buf[3][GROUPS][4];
    auto issue = [&](int l, uint4 (&dst)[GROUPS][4]) {
        if (l >= n_loads) return;
#pragma unroll
        for (int g = 0; g < GROUPS; ++g) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const uint8_t* src = rows[g][q] + 16 * l;
                if (vec_ok) {
                    dst[g][q] = *reinterpret_cast<const uint4*>(src);
                } else {
                    uint32_t w[4];
#pragma unroll
                    for (int j = 0; j < 4; ++j) w[j] = (l * SF_KB + j < k_blocks) ? *reinterpret_cast<const uint32_t*>(src + 4 * j) : 0u;
                    dst[g][q] = make_uint4(w[0], w[1], w[2], w[3]);
                }
            }
        }
    };
    auto word = [](const uint4& v, int j) { return j == 0 ? v.x : (j == 1 ? v.y : (j == 2 ? v.z : v.w)); };
    issue(0, buf[0]);
    issue(1, buf[1]);
    for (int l0 = 0; l0 < n_loads; l0 += 3) {
#pragma unroll
        for (int u = 0; u < 3; ++u) {
            const int l = l0 + u;
            if (l < n_loads) {
                issue(l + 2, buf[(u + 2) % 3]);
                mbar_wait(&sf_empty[sfs], sf_phase ^ 1);
                uint8_t* dst = sf_smem + sfs * (SF_KB * sf_kb_bytes);
#pragma unroll
                for (int j = 0; j < SF_KB; ++j)
#pragma unroll
                    for (int g = 0; g < GROUPS; ++g)
                        *reinterpret_cast<uint4*>(dst + j * sf_kb_bytes + 512 * g + 16 * lane) =
                            make_uint4(word(buf[u][g][0], j), word(buf[u][g][1], j), word(buf[u][g][2], j), word(buf[u][g][3], j));
                arrive(sfs);
                if (++sfs == SF_STAGES) { sfs = 0; sf_phase ^= 1; }
            }
        }
    }
}


// TMA-fed variant for streaming kernels (K3c), GROUPS == 1: the scales of 128 rows x 4 K blocks arrive as one
// [128 rows][16 B] box in a RAW_STAGES-deep raw ring (prefetch depth RAW_STAGES * 4 K blocks, no registers, no
// exposed global latency), and the warp only re-tiles shared -> shared into the tcgen05.cp chunk layout.
// raw_bars: RAW_STAGES mbarriers (count 1 + tx).  byte0 = first scale byte of this CTA's K range (multiple of 16).
template <int RAW_STAGES, typename Arrive>
__device__ __forceinline__ void sf_tma_tile4(const CUtensorMap* map_sf, int byte0, int row0, int k_blocks, uint8_t* raw, uint64_t* raw_bars,
                                             uint8_t* sf_smem, uint64_t* sf_empty, uint32_t& sfs, uint32_t& sf_phase, int lane, Arrive&& arrive) {
    constexpr int RAW_BYTES = 128 * 16;
    const int n_groups = (k_blocks + SF_KB - 1) / SF_KB;
    if (lane == 0) {
        for (int g = 0; g < RAW_STAGES && g < n_groups; ++g) {
            mbar_arrive_expect_tx(&raw_bars[g], RAW_BYTES);
            tma_load_2d(map_sf, &raw_bars[g], raw + g * RAW_BYTES, byte0 + 16 * g, row0);
        }
    }
    uint32_t rs = 0, rphase = 0;
    for (int g = 0; g < n_groups; ++g) {
        mbar_wait(&raw_bars[rs], rphase);
        const uint8_t* src = raw + rs * RAW_BYTES;
        uint4 w[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = *reinterpret_cast<const uint4*>(src + (q * 32 + lane) * 16);
        fence_proxy_async_smem();  // our generic reads of the raw stage are ordered before the TMA refill below
        __syncwarp();
        if (lane == 0 && g + RAW_STAGES < n_groups) {
            mbar_arrive_expect_tx(&raw_bars[rs], RAW_BYTES);
            tma_load_2d(map_sf, &raw_bars[rs], raw + rs * RAW_BYTES, byte0 + 16 * (g + RAW_STAGES), row0);
        }
        mbar_wait(&sf_empty[sfs], sf_phase ^ 1);
        uint8_t* dst = sf_smem + sfs * (SF_KB * 512);
        *reinterpret_cast<uint4*>(dst + 0 * 512 + 16 * lane) = make_uint4(w[0].x, w[1].x, w[2].x, w[3].x);
        *reinterpret_cast<uint4*>(dst + 1 * 512 + 16 * lane) = make_uint4(w[0].y, w[1].y, w[2].y, w[3].y);
        *reinterpret_cast<uint4*>(dst + 2 * 512 + 16 * lane) = make_uint4(w[0].z, w[1].z, w[2].z, w[3].z);
        *reinterpret_cast<uint4*>(dst + 3 * 512 + 16 * lane) = make_uint4(w[0].w, w[1].w, w[2].w, w[3].w);
        arrive(sfs);
        if (++sfs == SF_STAGES) { sfs = 0; sf_phase ^= 1; }
        if (++rs == RAW_STAGES) { rs = 0; rphase ^= 1; }
    }
}

// ---- host side: tensor maps (mxq_tmap.cu) ---------------------------------------------------------------------
// cuTensorMapEncodeTiled is a pure function of (base, extents, strides, box, type): the encoded 128-byte descriptors are kept in a
// small process-wide table keyed by exactly those arguments, so a layer's weight (and the handful of activation / output
// addresses the caching allocator cycles through) is encoded once, not on every launch.  Thread-safe.

// [batch][rows][K elements], K contiguous, box = 128 elements x box_rows, 128B swizzle, OOB rows read as zero.  The shared
// memory image is always 128 bytes per row and K block: one byte per element for E4M3 / E5M2 bytes, and for the packed 4 / 6-bit
// formats the TMA unit expands every 16 elements (8 / 12 bytes) to a 16-byte slot -- the layout kind::mxf8f6f4 reads.
bool cached_operand_map(CUtensorMap* map, const void* base, int64_t K, int64_t rows, int64_t batch, int64_t ld, int64_t batch_stride, int box_rows,
                        int operand_format, int device);
// E8M0 scales [rows][K/32 bytes] row-major: box = 16 bytes (4 K blocks) x 128 rows, rows / bytes past the edge read as zero
bool cached_scale_map(CUtensorMap* map, const void* base, int64_t scale_bytes_per_row, int64_t rows, int64_t ld, int device);
// raw code bytes [batch][rows][k_bytes] (K3d): box = box_bytes x box_rows, no swizzle, bytes / rows past the edge read as zero
bool cached_raw_map(CUtensorMap* map, const void* base, int64_t k_bytes, int64_t rows, int64_t batch, int64_t ld, int64_t batch_stride, int box_bytes,
                    int box_rows, int device);
// bf16 operand [batch][rows][K], K contiguous: box = 64 elements (one 128-byte swizzle row) x box_rows, OOB reads as zero
bool cached_bf16_operand_map(CUtensorMap* map, const void* base, int64_t K, int64_t rows, int64_t batch, int64_t ld, int64_t batch_stride, int box_rows,
                             int device);
// D: [batch][M][N] bf16, box = 64 columns x 32 rows, 128B swizzle (one epilogue warp's staging buffer)
bool cached_d_map(CUtensorMap* map, void* base, int64_t N, int64_t M, int64_t batch, int64_t ldd, int64_t batch_stride, int device);
// The dynamic shared-memory opt-in is a per-(kernel, device) attribute: set the first time a kernel is launched on a device,
// not on every call.
cudaError_t ensure_smem_attr(const void* kernel, int bytes, int device);
#ifdef MXQ_DEV
// developer builds: integer value of an environment variable (0 when unset), re-read on every call
static inline int dev_env(const char* name) { const char* v = getenv(name); return v ? atoi(v) : 0; }
#endif

}  // namespace gemm
}  // namespace mxq
