// tcgen05 implicit-GEMM kernel template + launch helpers shared by conv_tc.cu (fp32 activations, split-TF32 plans) and
// conv_tc_pair.cu (bf16-pair activations, the bf16x2 plan).  See conv_tc.cu for the design notes.
#pragma once
#include "common.cuh"
#include <cuda.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

namespace xfrb {

// tuning knobs of the bf16x2 (PAIRA) epilogue, overridable at build time for A/B runs (python -m xfr_b200.build --define ...)
// Measured on a B200, 256-probe ResNet-101 sweeps (profiles/r2_notes.md): none of the variants beats the layout of the TF32 kernels
// (XFRB_PAIRA_EW 0: idle split warpgroup kept, 8 epilogue warps at 200 registers, 12 at 128 for JOIN) - 3,855-3,897 maps/s against
// 3,741 (EW 8, 224 registers, JOIN loads ahead), 3,737 (EW 12), 3,701 (EW 8 unpaired), 3,670 (+ TMEM prefetch), 3,597 (no per-tile barrier),
// 3,747 vs 3,796 in the same call (JOIN operands staged by cp.async)
#ifndef XFRB_PAIRA_EW
#define XFRB_PAIRA_EW 0               /* 0: the warp layout of the TF32 kernels; 8 / 12: the split warpgroup's warps become epilogue warps (224 / 152 registers) */
#endif
#ifndef XFRB_PAIRA_JOIN_EW16
#define XFRB_PAIRA_JOIN_EW16 0        /* JOIN only: 16 epilogue warps (the split warpgroup joins in), 4 slabs each of a 256-wide tile instead of 6 / 5 / 5 over 12: measured 551 vs 493 us per layer3 launch (104 registers per epilogue thread) - slower */
#endif
#ifndef XFRB_PAIRA_PAIRED
#define XFRB_PAIRA_PAIRED 0           /* a warp takes the two 16-column slabs of a 32-column group back to back: both halves of every 128-byte line */
#endif
#ifndef XFRB_PAIRA_JOIN_STAGE
#define XFRB_PAIRA_JOIN_STAGE 0       /* JOIN: the next slab's four operand tensors are copied global -> shared by cp.async while this slab is computed (measured: 526 vs 486 us per layer3 launch - slower) */
#endif
#ifndef XFRB_PAIRA_JOIN_L2_PREFETCH
#define XFRB_PAIRA_JOIN_L2_PREFETCH 0 /* JOIN: prefetch.global.L2 of the next slab's operand lines (measured -2 %) */
#endif
#ifndef XFRB_PAIRA_PRM_DIRECT
#define XFRB_PAIRA_PRM_DIRECT 0       /* per-channel constants straight from global memory, no per-tile barrier (measured slower: warps drift apart) */
#endif
#ifndef XFRB_PAIRA_ACC_PREFETCH
#define XFRB_PAIRA_ACC_PREFETCH 0     /* request the next slab's accumulator from TMEM before this slab's math */
#endif

constexpr int TC_BM = 128;
constexpr int TC_BK = 32;                       // 32 fp32 = one 128-byte swizzle row
constexpr int TC_FIRST_SPLIT_WARP = 4;           // warpgroup 0: TMA (activations), MMA, TMA (weights), idle; warpgroup 1: split
constexpr int TC_FIRST_EPI_WARP = 8;             // warpgroups 2.. : epilogue (8 or 12 warps, TcCfg)
constexpr uint32_t A_TILE_BYTES = TC_BM * TC_BK * 4;   // 16 KB

struct TcGeom {
    int a4d;            // 0: A is a 2-D [M, C] matrix (1x1 conv / linear); 1: 4-D NHWC box loads (3x3)
    int R, Cin, kchunks, num_k;
    int H, W, bh, bimg, tiles_per_img, Nimg;    // 4-D box: bimg images x bh rows; tiles_per_img = row strips per image group
    int n_m_tiles, n_n_tiles;
    int b_rows;         // rows of one plane of B (the lo plane of the 3xTF32 split starts at row b_rows)
    int groups, tiles_per_group;   // gradient-row groups whose m-tiles are visited interleaved (1: natural order)
    uint32_t a_bytes;   // bytes one A load deposits (box volume * element size)
    int bk;             // K elements per k-block: 32 fp32 / 64 bf16 (one 128-byte swizzle row either way)
};

// ------------------------------------------------------------------ PTX wrappers
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::"r"(dst),
                 "l"(tm), "r"(c0), "r"(c1), "r"(bar)
                 : "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* tm, int c0, int c1, int c2, int c3, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(dst),
        "l"(tm), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
        : "memory");
}
// ---- CTA-pair (cluster) variants
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same smem offset in CTA `cta` of the cluster.  Plain forms, as CUTLASS's ClusterBarrier uses
// them: with .release.cluster here and try_wait.acquire.cluster on the waiting side every k-block paid a cluster-scope fence
// and the pair kernels ran 1.05-1.5x SLOWER than single CTAs; with the plain forms they run 1.1-1.15x faster.
__device__ __forceinline__ void mbar_arrive_remote(uint32_t bar, uint32_t cta) {
    asm volatile(
        "{\n\t"
        ".reg .b32 ra;\n\t"
        "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
        "mbarrier.arrive.shared::cluster.b64 _, [ra];\n\t"
        "}\n" ::"r"(bar), "r"(cta)
        : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(uint32_t bar, uint32_t parity) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAITC_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONEC_%=;\n\t"
        "bra WAITC_%=;\n\t"
        "DONEC_%=:\n\t"
        "}\n" ::"r"(bar), "r"(parity)
        : "memory");
}
// TMA load whose box lands at the same smem offset in every CTA of `mask` and signals each one's barrier at `bar`'s offset
__device__ __forceinline__ void tma_load_2d_mc(uint32_t dst, const CUtensorMap* tm, int c0, int c1, uint32_t bar, uint16_t mask) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes.multicast::cluster [%0], [%1, {%2, %3}], [%4], %5;" ::"r"(dst),
        "l"(tm), "r"(c0), "r"(c1), "r"(bar), "h"(mask)
        : "memory");
}
__device__ __forceinline__ void tc_commit_mc(uint32_t bar) {    // cta_group::1 commit arriving on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_commit2(uint32_t bar) {      // arrives on `bar` in BOTH CTAs of the pair
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
                 "h"((uint16_t)3)
                 : "memory");
}
__device__ __forceinline__ void tc_mma_tf32_2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_bf16_2(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_mma_bf16(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t da, uint64_t db, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(da), "l"(db), "r"(idesc), "r"(accumulate)
        : "memory");
}
// K-major operand tile in 128B-swizzled smem: rows of 128 bytes, 8-row groups 1024 bytes apart.
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);          // start address
    d |= (uint64_t)1 << 16;                            // leading byte offset (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;                  // stride byte offset between 8-row groups
    d |= (uint64_t)1 << 46;                            // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;                            // SWIZZLE_128B
    return d;
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float* v) {
    uint32_t r[16];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
#pragma unroll
    for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
// wait for outstanding tcgen05.ld; the registers are tied to the asm so no use of them can be scheduled above it
__device__ __forceinline__ void tmem_ld_wait(float* v) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+f"(v[0]), "+f"(v[1]), "+f"(v[2]), "+f"(v[3]), "+f"(v[4]), "+f"(v[5]), "+f"(v[6]), "+f"(v[7]), "+f"(v[8]),
                   "+f"(v[9]), "+f"(v[10]), "+f"(v[11]), "+f"(v[12]), "+f"(v[13]), "+f"(v[14]), "+f"(v[15])
                 :
                 : "memory");
}

// ------------------------------------------------------------------ kernel
// KIND only matters for the stage count: the JOIN kernels (K = Cout of a 1x1 conv, 2-8 k-blocks per tile) are bound by
// their epilogue's global loads, which like a large L1: two stages leave ~70 KB more of the unified L1/shared memory to it.
template <int BN, int SPLIT, int PAIR = 0, int KIND = EPI_PLAIN>
struct TcCfg {
    static constexpr bool CTA2 = PAIR == 1;
    static constexpr int B_ROWS = CTA2 ? BN / 2 : BN;            // weight rows this CTA stages (a cta_group::2 pair shares the tile)
    static constexpr uint32_t B_TILE_BYTES = B_ROWS * TC_BK * 4;
    static constexpr bool PAIRA = SPLIT >= 4;                     // bf16x2 plan: activations arrive as (hi, lo) bf16 tiles, nothing is split here
    static constexpr uint32_t B_LO_BYTES = SPLIT == 1 ? B_TILE_BYTES : (SPLIT == 3 || SPLIT == 5) ? B_TILE_BYTES / 2 : 0;
    // Two independent rings: activations (raw + lo tile) and weights (hi + lo planes).  Activation tiles are unique to the
    // CTA and partly come from HBM; weight tiles are re-read by every CTA and sit in L2 - so the activation ring gets the
    // depth (FWD dual tiles: 3 + 2 where one coupled ring held 2; W+ dgrads: 4 + 2 instead of 3).  A CTA pair keeps one
    // coupled ring (its leader learns that the peer's data landed from the peer's split warps).
    static constexpr uint32_t A_BYTES = A_TILE_BYTES * (SPLIT ? 2 : 1);
    static constexpr uint32_t B_BYTES = B_TILE_BYTES + B_LO_BYTES;
    static constexpr uint32_t RING_BUDGET = 192 * 1024;
    static constexpr bool SPLITRING = !CTA2;
    static constexpr bool SHORT = (KIND == EPI_JOIN && (SPLIT == 2 || SPLIT == 4));      // epilogue-bound: leave the memory to L1
    // bf16x2 JOIN: every epilogue warp owns an 8 KB staging buffer (4 operand tensors x 32 rows x 64 bytes) that cp.async fills with
    // the NEXT slab's operands while the current slab is computed - the load latency that was a third of the kernel's stall samples
    // (ncu r2: first use of a loaded operand) hides behind math and stores without a second set of 64 load registers.  The main-loop
    // ring shrinks to ONE stage to make room: JOIN tiles have 1 - 8 k-blocks and their epilogue, not their main loop, is the bound.
    static constexpr bool STAGE_LOADS = KIND == EPI_JOIN && PAIRA && XFRB_PAIRA_JOIN_STAGE;
    static constexpr int SHORT_STAGES = STAGE_LOADS ? 1 : 2;
    static constexpr int COUPLED_RAW = RING_BUDGET / (A_BYTES + B_BYTES);
    static constexpr int COUPLED = SHORT ? SHORT_STAGES : (COUPLED_RAW > 8 ? 8 : COUPLED_RAW);
    static constexpr int NB_PICK = B_BYTES >= 32 * 1024 ? 2 : (B_BYTES >= 16 * 1024 ? 3 : 4);
    static constexpr int NB = SPLITRING ? (SHORT ? SHORT_STAGES : NB_PICK) : COUPLED;
    static constexpr int NA_RAW = (RING_BUDGET - NB * B_BYTES) / A_BYTES;
    static constexpr int NA = SPLITRING ? (SHORT ? SHORT_STAGES : (NA_RAW > 6 ? 6 : NA_RAW)) : COUPLED;
    static constexpr uint32_t RING_BYTES = NA * A_BYTES + NB * B_BYTES;
    static constexpr uint32_t TMEM_COLS = 2 * BN;          // two accumulator stages (128 or 256: powers of two)
    // per-channel constants staged per accumulator stage: rows bn[0..3] (alpha, beta, sp, tp), bias_t, bias_p, PRM_LD wide
    static constexpr int PRM_LD = KIND == EPI_FWD_DUAL ? BN / 2 : BN;
    static constexpr int PRM_ROWS = KIND == EPI_FWD_DUAL ? 6 : (KIND == EPI_PLAIN ? 1 : 4);       // PLAIN: the bias row alone
    static constexpr int BIAS_ROW = KIND == EPI_PLAIN ? 0 : 4;
    static constexpr uint32_t PRM_BYTES = 2 * PRM_ROWS * PRM_LD * 4;
    // Epilogue warps: two per TMEM lane quarter, or three for the JOIN kernels, whose epilogue (4 tensor reads, 2 writes, the
    // longest hook chain) is what bounds them: 707 -> 615 us per launch with 12 warps; the others lose 6 % to the smaller
    // register budget.  setmaxnreg moves registers between whole warpgroups inside the pool the CTA was LAUNCHED with
    // (threads x the most __launch_bounds__ allows; asking for more blocks setmaxnreg.inc forever).
    // bf16x2 plan (PAIRA): no operand-split warpgroup - its four warps become epilogue warps: 4 producer warps + 12 epilogue
    // warps = 512 threads for every kind (a CTA pair's "landed" relay moves to the idle warp 3 of the producer warpgroup)
    static constexpr bool JOIN16 = PAIRA && KIND == EPI_JOIN && XFRB_PAIRA_JOIN_EW16 != 0;
    static constexpr bool REPURPOSE = (PAIRA && XFRB_PAIRA_EW != 0) || JOIN16;
    static constexpr int FIRST_EPI_WARP = REPURPOSE ? TC_FIRST_SPLIT_WARP : TC_FIRST_EPI_WARP;
    static constexpr int EPI_WARPS = JOIN16 ? 16 : (REPURPOSE ? XFRB_PAIRA_EW : (KIND == EPI_JOIN ? 12 : 8));
    static constexpr int THREADS = (FIRST_EPI_WARP + EPI_WARPS) * 32;                     // 512 / 640
    static constexpr int REGS_LAUNCH = (65536 / THREADS) / 8 * 8;                          // 128 / 96
    static constexpr int REGS_PRODUCER = JOIN16 ? 40 : (REPURPOSE ? 56 : (EPI_WARPS == 12 ? 48 : 56));
    static constexpr int REGS_EPILOGUE = JOIN16 ? 104 : (REPURPOSE ? (EPI_WARPS == 12 ? 152 : 224) : (EPI_WARPS == 12 ? 128 : 200));
    static_assert(FIRST_EPI_WARP * 32 * REGS_PRODUCER + EPI_WARPS * 32 * REGS_EPILOGUE <= THREADS * REGS_LAUNCH,
                  "setmaxnreg budgets exceed the CTA's register pool");
    static constexpr uint32_t TR_BYTES = EPI_WARPS * 2048;      // per epilogue warp: 32 rows x 16 columns transpose slab
    static constexpr uint32_t BAR_BYTES = 512;
    static constexpr uint32_t LD_BYTES = STAGE_LOADS ? EPI_WARPS * 8192 : 0;
    static constexpr uint32_t SMEM_BYTES = RING_BYTES + 1024 /*align slack*/ + BAR_BYTES + PRM_BYTES + TR_BYTES + LD_BYTES;
    static_assert(SMEM_BYTES <= 232448, "exceeds the 227 KB of dynamic shared memory a CTA can opt into");
};

// MODE: the ebp_subtree_mode id of the MID / JOIN hook chains as a compile-time constant (the chains are ~2x cheaper once
// the mode branches fold away: tools/epi_probe.py), or -1 to read it from EpiParams at run time.
template <int BN, int SPLIT, int KIND, int PAIR, int MODE = -1>
__global__ void __launch_bounds__((TcCfg<BN, SPLIT, PAIR, KIND>::THREADS), 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
               const __grid_constant__ CUtensorMap tmBlo, const TcGeom g, const EpiParams ep) {
    using Cfg = TcCfg<BN, SPLIT, PAIR, KIND>;
    constexpr bool CTA2 = PAIR == 1;             // cta_group::2 pair: one M = 256 MMA, the leader issues
    constexpr bool MC = PAIR == 2;               // multicast pair: private M = 128 MMAs, shared weight loads
    constexpr bool CLUSTERED = PAIR != 0;
    static_assert(!CTA2 || SPLIT != 0, "the CTA-pair kernel signals the leader from the split warps");
    const uint32_t rank = CLUSTERED ? cluster_ctarank() : 0u;
    const bool leader = rank == 0 || MC;         // who issues MMAs: every CTA of a multicast pair
    const int worker = CLUSTERED ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;     // tile stream this CTA (pair) walks
    const int nworkers = CLUSTERED ? (int)(gridDim.x >> 1) : (int)gridDim.x;
    constexpr bool SPLIT3 = SPLIT != 0;          // the activation operand is a (hi, lo) pair of tiles
    constexpr bool PAIRA = SPLIT >= 4;           // ... that arrives split from HBM as bf16 (hi, lo) half-rows: kind::f16 MMAs, no split warps
    constexpr bool DUALLO = SPLIT == 3 || SPLIT == 5;   // dual forward pack: a lo weight plane for the W half of the tile
    constexpr int NA = Cfg::NA, NB = Cfg::NB;
    constexpr int EW = Cfg::EPI_WARPS;
    constexpr bool SPLITRING = Cfg::SPLITRING;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
    // barrier block lives after the stages
    const uint32_t bar_base = smem_base + Cfg::RING_BYTES;
    auto full_bar = [&](int s) { return bar_base + 8u * s; };                       // activation ring (coupled ring: both)
    auto empty_bar = [&](int s) { return bar_base + 8u * (NA + s); };
    auto split_bar = [&](int s) { return bar_base + 8u * (2 * NA + s); };
    auto fullb_bar = [&](int s) { return bar_base + 8u * (3 * NA + s); };            // weight ring
    auto emptyb_bar = [&](int s) { return bar_base + 8u * (3 * NA + NB + s); };
    auto tfull_bar = [&](int a) { return bar_base + 8u * (3 * NA + 2 * NB + a); };
    auto tempty_bar = [&](int a) { return bar_base + 8u * (3 * NA + 2 * NB + 2 + a); };
    auto land_bar = [&](int s) { return bar_base + 8u * (3 * NA + 2 * NB + 4 + s); };   // CTA pair: "both CTAs' TMA data landed"
    static_assert(8 * (4 * NA + 2 * NB + 4) + 4 <= Cfg::BAR_BYTES, "barrier block");
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(smem_gen + Cfg::RING_BYTES + 8 * (4 * NA + 2 * NB + 4));
    float* prm_s = reinterpret_cast<float*>(smem_gen + Cfg::RING_BYTES + Cfg::BAR_BYTES);   // [2][rows][PRM_LD]
    float4* tr_s = reinterpret_cast<float4*>(smem_gen + Cfg::RING_BYTES + Cfg::BAR_BYTES + Cfg::PRM_BYTES);

    auto a_hi = [&](int s) { return smem_base + s * Cfg::A_BYTES; };
    auto a_lo = [&](int s) { return smem_base + s * Cfg::A_BYTES + A_TILE_BYTES; };                           // split only
    auto b_hi = [&](int s) { return smem_base + NA * Cfg::A_BYTES + s * Cfg::B_BYTES; };
    auto b_lo = [&](int s) { return smem_base + NA * Cfg::A_BYTES + s * Cfg::B_BYTES + Cfg::B_TILE_BYTES; };   // SPLIT 1 / 3

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int total_tiles = (CLUSTERED ? (g.n_m_tiles + 1) / 2 : g.n_m_tiles) * g.n_n_tiles;

    if (warp == 0 && lane == 0) {
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
        asm volatile("prefetch.tensormap [%0];" ::"l"(&tmB) : "memory");
        if (DUALLO) asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBlo) : "memory");
        for (int s = 0; s < NA; ++s) {
            mbar_init(full_bar(s), 1);
            mbar_init(empty_bar(s), 1);
            mbar_init(split_bar(s), CTA2 ? 8 : 4);               // pair: the peer's split warps arrive here too (leader's copy)
            if (CTA2) mbar_init(land_bar(s), PAIRA ? 2 : 8);         // PAIRA: one relay lane per CTA (warp 3)
        }
        for (int s = 0; s < NB; ++s) {
            mbar_init(fullb_bar(s), 1);
            mbar_init(emptyb_bar(s), MC ? 2 : 1);                // multicast pair: both CTAs' MMAs must have released the stage
        }
        for (int a = 0; a < 2; ++a) {
            mbar_init(tfull_bar(a), 1);
            mbar_init(tempty_bar(a), CTA2 ? 2 * EW : EW);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        if (CTA2) {
            asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;");
        } else {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
    }
    tc_fence_before();
    if (CLUSTERED) cluster_sync_all();  // both CTAs' barriers are initialised before any remote arrive / multicast
    else __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = *tmem_slot;
    // Programmatic dependent launch (conv_tc_pdl()): everything above - barrier set-up, TMEM allocation, tensor-map prefetch - may
    // run while the previous kernel of the stream is still finishing on other SMs (a batch-1 sweep is ~230 kernels of a few tiles
    // each); nothing below touches global memory before the previous kernel has completed and flushed.  No-ops in a normal launch.
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
    asm volatile("griddepcontrol.wait;" ::: "memory");

    // tile -> coordinates
    auto tile_coords = [&](int tile, int& m0, int& mvalid, int& n_img0, int& h0, int& ncol0) {
        int mt = tile / g.n_n_tiles, nt = tile - mt * g.n_n_tiles;
        if (CLUSTERED) mt = 2 * mt + (int)rank;        // the pair's tile is 256 rows: two consecutive m-tiles
        const bool phantom = mt >= g.n_m_tiles;        // odd tile count: the last pair's second half loads zeros, stores nothing
        // gradient-row groups (mate / non-mate rows of the same probes) read the same saved tensors: visit group 0's
        // tile i, then group 1's tile i, ... so the second read of a saved tile hits L2 instead of HBM
        if (g.groups > 1 && !phantom) mt = (mt % g.groups) * g.tiles_per_group + mt / g.groups;
        ncol0 = nt * BN;
        if (!g.a4d) {
            m0 = mt * TC_BM;
            mvalid = min(TC_BM, ep.M - m0);
            n_img0 = 0;
            h0 = 0;
        } else {
            // 4-D box: bimg images x bh image rows x W pixels; tile mt = (image group, row strip).  Rows of the tile are NOT
            // contiguous in the [N*H*W, C] matrix when bimg > 1 and bh < H: the epilogue maps them one by one (row_of).
            const int gi = mt / g.tiles_per_img;
            n_img0 = gi * g.bimg;
            h0 = (mt - gi * g.tiles_per_img) * g.bh;
            m0 = (n_img0 * g.H + h0) * g.W;
            mvalid = TC_BM;
        }
        if (phantom) mvalid = 0;
    };

    if (warp < TC_FIRST_SPLIT_WARP) {
    asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_PRODUCER));      // whole warpgroup 0
    if (warp == 0) {
        // ===================== TMA producer =====================
        if (lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = worker; tile < total_tiles; tile += nworkers) {
                int m0, mvalid, n_img0, h0, ncol0;
                tile_coords(tile, m0, mvalid, n_img0, h0, ncol0);
                for (int kb = 0; kb < g.num_k; ++kb) {
                    mbar_wait(empty_bar(s), ph ^ 1u);
                    mbar_expect_tx(full_bar(s), (PAIRA ? 2u : 1u) * g.a_bytes + (SPLITRING ? 0u : Cfg::B_BYTES));
                    int tap = kb / g.kchunks;
                    int c0 = (kb - tap * g.kchunks) * g.bk;
                    if (g.a4d) {
                        int dr = tap / g.R - (g.R >> 1), ds = tap % g.R - (g.R >> 1);
                        tma_load_4d(a_hi(s), &tmA, c0, ds, h0 + dr, n_img0, full_bar(s));
                        if (PAIRA) tma_load_4d(a_lo(s), &tmA, g.Cin + c0, ds, h0 + dr, n_img0, full_bar(s));   // lo half-row of the pair
                    } else {
                        tma_load_2d(a_hi(s), &tmA, c0, m0, full_bar(s));
                        if (PAIRA) tma_load_2d(a_lo(s), &tmA, g.Cin + c0, m0, full_bar(s));
                    }
                    if (!SPLITRING) {                                                   // coupled ring (CTA pair)
                        const int brow = ncol0 + (CTA2 ? (int)rank * (BN / 2) : 0);    // pair: this CTA stages its half of the tile
                        tma_load_2d(b_hi(s), &tmB, kb * g.bk, brow, full_bar(s));
                        if (SPLIT == 1) tma_load_2d(b_lo(s), &tmB, kb * g.bk, g.b_rows + brow, full_bar(s));
                        if (DUALLO)
                            tma_load_2d(b_lo(s), &tmBlo, kb * g.bk, g.b_rows + ncol0 + (CTA2 ? (int)rank * (BN / 4) : 0), full_bar(s));
                    }
                    if (++s == NA) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 2) {
        // ===================== TMA producer of the weight ring =====================
        if (SPLITRING && lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = worker; tile < total_tiles; tile += nworkers) {
                const int ncol0 = (tile % g.n_n_tiles) * BN;
                for (int kb = 0; kb < g.num_k; ++kb) {
                    if (MC) mbar_wait_cluster(emptyb_bar(s), ph ^ 1u);      // released by both CTAs (the peer writes into this stage too)
                    else mbar_wait(emptyb_bar(s), ph ^ 1u);
                    mbar_expect_tx(fullb_bar(s), Cfg::B_BYTES);             // both halves land here: mine and the peer's multicast
                    if (MC) {
                        // my half of every plane, multicast into both CTAs at the half's own offset
                        constexpr uint32_t HALF = Cfg::B_TILE_BYTES / 2;
                        const int r0 = (int)rank * (BN / 2);
                        tma_load_2d_mc(b_hi(s) + rank * HALF, &tmB, kb * g.bk, ncol0 + r0, fullb_bar(s), 3);
                        if (SPLIT == 1) tma_load_2d_mc(b_lo(s) + rank * HALF, &tmB, kb * g.bk, g.b_rows + ncol0 + r0, fullb_bar(s), 3);
                        if (DUALLO)
                            tma_load_2d_mc(b_lo(s) + rank * (HALF / 2), &tmBlo, kb * g.bk, g.b_rows + ncol0 + (int)rank * (BN / 4), fullb_bar(s), 3);
                    } else {
                        tma_load_2d(b_hi(s), &tmB, kb * g.bk, ncol0, fullb_bar(s));
                        if (SPLIT == 1) tma_load_2d(b_lo(s), &tmB, kb * g.bk, g.b_rows + ncol0, fullb_bar(s));   // host-split lo plane
                        if (DUALLO) tma_load_2d(b_lo(s), &tmBlo, kb * g.bk, g.b_rows + ncol0, fullb_bar(s));  // lo of the W half
                    }
                    if (++s == NB) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 3) {
        // ===================== bf16x2 CTA pair: relay "this CTA's stage landed" to the leader's MMA thread =====================
        if (PAIRA && CTA2 && lane == 0) {
            int s = 0;
            uint32_t ph = 0;
            for (int tile = worker; tile < total_tiles; tile += nworkers) {
                for (int kb = 0; kb < g.num_k; ++kb) {
                    mbar_wait(full_bar(s), ph);
                    if (!leader) mbar_arrive_remote(land_bar(s), 0);
                    else mbar_arrive(land_bar(s));
                    if (++s == NA) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else if (warp == 1) {
        // ===================== MMA issuer (cta_group::2 pair: the leader CTA only) =====================
        constexpr uint32_t MM = CTA2 ? 2 * TC_BM : TC_BM;      // cta_group::2: M = 256, rows 128.. live in the peer's TMEM
        // instruction descriptor: fp32 accumulate (bit 4), A / B format (bits 7-9 / 10-12: 2 = tf32, 1 = bf16), N >> 3, M >> 4
        constexpr uint32_t FMT = PAIRA ? 1u : 2u;
        constexpr uint32_t idesc = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)(MM >> 4) << 24);
        constexpr uint32_t idesc_half = (1u << 4) | (FMT << 7) | (FMT << 10) | ((uint32_t)(BN >> 4) << 17) | ((uint32_t)(MM >> 4) << 24);
        auto mma = [&](uint32_t d, uint64_t da, uint64_t db, uint32_t id, uint32_t acc) {
            if (PAIRA) {
                if (CTA2) tc_mma_bf16_2(d, da, db, id, acc);
                else tc_mma_bf16(d, da, db, id, acc);
            } else {
                if (CTA2) tc_mma_tf32_2(d, da, db, id, acc);
                else tc_mma_tf32(d, da, db, id, acc);
            }
        };
        int s = 0, sb = 0;
        uint32_t ph = 0, phb = 0;
        int it = 0;
        if (leader)
        for (int tile = worker; tile < total_tiles; tile += nworkers, ++it) {
            const int a = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            if (lane == 0) {
                if (CTA2) mbar_wait_cluster(tempty_bar(a), aph ^ 1u);
                else mbar_wait(tempty_bar(a), aph ^ 1u);
            }
            __syncwarp();
            tc_fence_after();
            const uint32_t tacc = tmem_base + (uint32_t)(a * BN);
            // The hi pass of a k-block needs only the TMA data (the tensor core truncates the raw fp32 activations to TF32
            // itself), so it is issued as soon as the stage lands and runs while the split warps produce the lo tile.
            for (int kb = 0; kb < g.num_k; ++kb) {
                if (lane == 0) {
                    if (CTA2) mbar_wait_cluster(land_bar(s), ph);             // both CTAs' split warps saw their TMA data land
                    else mbar_wait(full_bar(s), ph);
                    if (SPLITRING) mbar_wait(fullb_bar(sb), phb);
                    tc_fence_after();
                    const int bs = SPLITRING ? sb : s;
                    const uint64_t dah = make_desc(a_hi(s)), dbh = make_desc(b_hi(bs));
#pragma unroll
                    for (int k = 0; k < TC_BK / 8; ++k) {
                        const uint64_t koff = (uint64_t)((k * 32) >> 4);
                        mma(tacc, dah + koff, dbh + koff, idesc, (kb | k) != 0 ? 1u : 0u);
                    }
                    if (SPLIT3) {
                        if (!PAIRA) {          // the lo tile of a bf16 pair landed with the hi tile
                            if (CTA2) mbar_wait_cluster(split_bar(s), ph);
                            else mbar_wait(split_bar(s), ph);
                            tc_fence_after();
                        }
                        const uint64_t dal = make_desc(a_lo(s)), dbl = make_desc(b_lo(bs));
#pragma unroll
                        for (int k = 0; k < TC_BK / 8; ++k) {
                            const uint64_t koff = (uint64_t)((k * 32) >> 4);
                            mma(tacc, dal + koff, dbh + koff, idesc, 1u);
                            if (SPLIT == 1) mma(tacc, dah + koff, dbl + koff, idesc, 1u);
                            if (DUALLO) mma(tacc, dah + koff, dbl + koff, idesc_half, 1u);   // columns [0, BN/2): the W half
                        }
                    }
                    if (CTA2) {
                        tc_commit2(empty_bar(s));                // both CTAs' stage s
                        if (kb == g.num_k - 1) tc_commit2(tfull_bar(a));
                    } else {
                        tc_commit(empty_bar(s));                 // smem stages reusable once these MMAs retire
                        if (SPLITRING) {
                            if (MC) tc_commit_mc(emptyb_bar(sb));    // the peer's weight producer waits for this CTA too
                            else tc_commit(emptyb_bar(sb));
                        }
                        if (kb == g.num_k - 1) tc_commit(tfull_bar(a));
                    }
                }
                __syncwarp();
                if (++s == NA) { s = 0; ph ^= 1u; }
                if (++sb == NB) { sb = 0; phb ^= 1u; }
            }
        }
    }
    } else if (!Cfg::REPURPOSE && warp < TC_FIRST_EPI_WARP) {
        // ===================== operand split (3xTF32; idle in the bf16x2 plan) =====================
        asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_PRODUCER));  // whole warpgroup 1
        if (SPLIT3 && !PAIRA) {
            const int t = threadIdx.x - TC_FIRST_SPLIT_WARP * 32;    // 0..127
            int s = 0;
            uint32_t ph = 0;
            for (int tile = worker; tile < total_tiles; tile += nworkers) {
                for (int kb = 0; kb < g.num_k; ++kb) {
                    mbar_wait(full_bar(s), ph);
                    if (CTA2 && lane == 0) {           // tell the leader's MMA thread that this CTA's stage landed: its hi pass can go
                        if (!leader) mbar_arrive_remote(land_bar(s), 0);
                        else mbar_arrive(land_bar(s));
                    }
                    // only the activation tile is split here; the weight tile arrives as (hi, lo) planes split on the host
                    float4* hi = reinterpret_cast<float4*>(smem_gen + s * Cfg::A_BYTES);
                    float4* lo = reinterpret_cast<float4*>(smem_gen + s * Cfg::A_BYTES + A_TILE_BYTES);
                    constexpr int NV = A_TILE_BYTES / 16;
#pragma unroll 4
                    for (int i = t; i < NV; i += 128) {
                        // tcgen05 kind::tf32 reads only the top 19 bits of an fp32 operand (measured: tools/trunc_probe.py), so
                        // the raw tile IS the hi operand, hi = trunc(x).  x - trunc(x) is exact in fp32 but carries up to 13
                        // significant bits: it is rounded to nearest TF32 here so that the hardware truncation of the lo
                        // operand loses nothing more and the residual (<= 2^-21 |x|) stays unbiased.
                        const float4 v = hi[i];
                        float4 l;
                        uint32_t u;
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.x - __uint_as_float(__float_as_uint(v.x) & 0xFFFFE000u))); l.x = __uint_as_float(u);
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.y - __uint_as_float(__float_as_uint(v.y) & 0xFFFFE000u))); l.y = __uint_as_float(u);
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.z - __uint_as_float(__float_as_uint(v.z) & 0xFFFFE000u))); l.z = __uint_as_float(u);
                        asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(v.w - __uint_as_float(__float_as_uint(v.w) & 0xFFFFE000u))); l.w = __uint_as_float(u);
                        lo[i] = l;
                    }
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy writes -> visible to UMMA
                    __syncwarp();
                    if (lane == 0) {
                        if (CTA2 && !leader) mbar_arrive_remote(split_bar(s), 0);     // the leader's MMA thread waits for both halves
                        else mbar_arrive(split_bar(s));
                    }
                    if (++s == NA) { s = 0; ph ^= 1u; }
                }
            }
        }
    } else {
        // ===================== epilogue =====================
        // TMEM hands each thread one pixel ROW of the accumulator, but NHWC tensors want consecutive lanes on
        // consecutive CHANNELS: a 32-row x 16-column slab is therefore transposed through a 2 KB XOR-swizzled smem
        // buffer per warp, after which lane l owns channels 4*(l%4)..+3 of rows 8i + l/4 (i = 0..3).  Every global
        // access of the epilogue is then a 64-byte row segment per 4 lanes (8 rows per request instead of 32), and all
        // loads of a slab are issued before the accumulator is waited for.  The two warps of a TMEM lane quarter
        // alternate slabs; per-channel constants are staged in smem once per tile.
        asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(Cfg::REGS_EPILOGUE));  // the epilogue warpgroups
        const int ew = warp - Cfg::FIRST_EPI_WARP;     // 0..EW-1
        const int q = warp & 3;                        // TMEM lane quarter this warp may access
        const int part = ew >> 2;                      // which slabs of the tile this warp owns: part, part + 3, ...
        const int et = threadIdx.x - Cfg::FIRST_EPI_WARP * 32;   // 0..32*EW-1
        const int cgl = lane & 3;                      // my 4-channel group inside the slab
        const int rsub = lane >> 2;                    // my row inside each group of 8 rows
        float4* tbuf = tr_s + ew * 128;
        uint8_t* ldbuf = smem_gen + Cfg::RING_BYTES + Cfg::BAR_BYTES + Cfg::PRM_BYTES + Cfg::TR_BYTES + (Cfg::STAGE_LOADS ? ew * 8192 : 0);
        constexpr int CH = (KIND == EPI_FWD_DUAL) ? BN / 2 : BN;      // channels per tile
        constexpr int NL = (KIND == EPI_JOIN) ? 4 : (KIND == EPI_MID) ? 2 : 1;     // tensors loaded per output element
        struct Loads { float4 v[NL][4]; };             // one slab's global loads: [tensor][row group]
        struct Acc { float vt[16]; float vp[KIND == EPI_FWD_DUAL ? 16 : 1]; };   // one slab of the accumulator(s), row per lane
        constexpr bool ACC_PREFETCH = PAIRA && XFRB_PAIRA_ACC_PREFETCH;     // TF32 plans: measured +-0, costs 16-32 registers of the 136
        constexpr bool LOAD_AHEAD = KIND != EPI_JOIN || (Cfg::REPURPOSE && EW == 8);   // the next slab's global loads stay in flight (JOIN: 4 tensors, needs the 224-register budget)
        constexpr int SLAB_STRIDE = 16 * (EW / 4);         // columns between two slabs of the same warp
        constexpr bool PAIRED = PAIRA && XFRB_PAIRA_PAIRED;
        int it = 0;
        for (int tile = worker; tile < total_tiles; tile += nworkers, ++it) {
            const int a = it & 1;
            const uint32_t aph = (it >> 1) & 1;
            int m0, mvalid, n_img0, h0, ncol0;
            tile_coords(tile, m0, mvalid, n_img0, h0, ncol0);
            const int cbase = (KIND == EPI_FWD_DUAL) ? (ncol0 / BN) * (BN / 2) : ncol0;
            constexpr int PLD = Cfg::PRM_LD;
            float* prm = prm_s + a * Cfg::PRM_ROWS * PLD;
            // stage the per-channel constants of this tile: rows 0-3 bn (alpha, beta, sp, tp), 4 bias_t, 5 bias_p.
            // bf16x2 kernels (PRM_DIRECT) read them straight from global memory instead (L1-resident: every row of the tile uses
            // the same few hundred floats): no staging, and above all no CTA-wide barrier per tile - with it the epilogue warps
            // moved in lockstep and the ones with fewer slabs idled (ncu: 1.8 - 4.6 warps stalled on the barrier per issue)
            constexpr bool PRM_DIRECT = PAIRA && XFRB_PAIRA_PRM_DIRECT;
            if (!PRM_DIRECT && KIND != EPI_PLAIN) {
                for (int i = et; i < 4 * CH; i += EW * 32) {
                    int r = i / CH, j = i - r * CH;
                    prm[r * PLD + j] = __ldg(ep.bn + (size_t)r * ep.C + cbase + j);
                }
            }
            if (PRM_DIRECT) {
            } else if (KIND == EPI_FWD_DUAL) {
                for (int i = et; i < BN; i += EW * 32) {
                    int r = i / CH, j = i - r * CH;            // r = 0: true bias, 1: positive twin
                    prm[(4 + r) * PLD + j] = __ldg(ep.bias + ncol0 + r * CH + j);
                }
            } else if (KIND == EPI_PLAIN) {
                for (int i = et; i < BN; i += EW * 32) prm[Cfg::BIAS_ROW * PLD + i] = ep.bias ? __ldg(ep.bias + ncol0 + i) : 0.f;
            }
            // rows this thread finishes (coalesced orientation)
            int mrow[4], msav[4];
            bool vrow[4];
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const int r = q * 32 + 8 * i + rsub;
                if (!g.a4d) {
                    vrow[i] = r < mvalid;
                    mrow[i] = m0 + r;
                } else {                                   // row r of the box -> (image, image row, pixel)
                    const int strip = g.bh * g.W;
                    const int bi = r / strip, rem = r - bi * strip;
                    const int hh = h0 + rem / g.W, img = n_img0 + bi;
                    vrow[i] = mvalid > 0 && bi < g.bimg && img < g.Nimg && hh < g.H;
                    mrow[i] = (img * g.H + hh) * g.W + (rem % g.W);
                }
                msav[i] = (KIND == EPI_MID || KIND == EPI_JOIN) ? mrow[i] % ep.Ms : mrow[i];
            }
            // ---- every global load of slab j (issued one slab ahead of its use)
            const int mode = MODE >= 0 ? MODE : ep.mode;
            const int dbg = ep.hooks >> 8;              // profiling switches (tools/epi_probe.py): 1 no loads, 2 no stores, 4 no math
            auto issue_loads = [&](int j, Loads& L) {
                const int c = cbase + j + 4 * cgl;
                const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
#pragma unroll
                    for (int t = 0; t < NL; ++t) L.v[t][i] = z4;
                    if (!vrow[i] || (dbg & 1)) continue;
                    if (KIND == EPI_FWD_DUAL) {
                        if (ep.res != nullptr && c < ep.res_c) L.v[0][i] = __ldg(reinterpret_cast<const float4*>(ep.res + (size_t)mrow[i] * ep.res_c + c));
                    } else if (KIND == EPI_PLAIN) {
                        if (ep.g_res != nullptr) L.v[0][i] = *reinterpret_cast<const float4*>(ep.g_res + (size_t)mrow[i] * ep.C + c);
                    } else {
                        const size_t offs = (size_t)msav[i] * ep.C + c;
                        // plain cached loads: L1 allocation merges the two 64-byte halves of a line that the two warps of a
                        // lane quarter read (ld.global.cs / L1::no_allocate measured 35-50 % slower here)
                        L.v[0][i] = __ldg(reinterpret_cast<const float4*>(ep.o + offs));
                        L.v[1][i] = __ldg(reinterpret_cast<const float4*>(ep.xr + offs));
                        if (KIND == EPI_JOIN) {
                            L.v[2][i] = __ldg(reinterpret_cast<const float4*>(ep.outp + offs));
                            L.v[3][i] = __ldg(reinterpret_cast<const float4*>(ep.g_res + (size_t)mrow[i] * ep.C + c));
                        }
                    }
                }
            };
            const uint32_t tacc = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(a * BN);
            // ---- accumulator slab j (requested from TMEM one slab ahead) -> epilogue math -> stores
            auto request_acc = [&](int j, Acc& V) {
                tmem_ld16(tacc + j, V.vt);
                if (KIND == EPI_FWD_DUAL) tmem_ld16(tacc + CH + j, V.vp);
            };
            auto process = [&](int j, const Loads& L, Acc& V, int jn, Acc& Vn) {
                const int c = cbase + j + 4 * cgl;
                float* vt = V.vt;
                float* vp = V.vp;
                // transpose the accumulator slab(s): row-per-lane -> channel-group-per-lane
                float4 at[4], apv[4];
                if (!ACC_PREFETCH) request_acc(j, V);
                tmem_ld_wait(vt);
                if (KIND == EPI_FWD_DUAL) tmem_ld_wait(vp);
                if (ACC_PREFETCH && jn < CH) request_acc(jn, Vn);      // the next slab's TMEM read overlaps this slab's math
#pragma unroll
                for (int g4 = 0; g4 < 4; ++g4)
                    tbuf[lane * 4 + (g4 ^ ((lane >> 1) & 3))] = make_float4(vt[4 * g4], vt[4 * g4 + 1], vt[4 * g4 + 2], vt[4 * g4 + 3]);
                __syncwarp();
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int rl = 8 * i + rsub;
                    at[i] = tbuf[rl * 4 + (cgl ^ ((rl >> 1) & 3))];
                }
                __syncwarp();
                if (KIND == EPI_FWD_DUAL) {
#pragma unroll
                    for (int g4 = 0; g4 < 4; ++g4)
                        tbuf[lane * 4 + (g4 ^ ((lane >> 1) & 3))] = make_float4(vp[4 * g4], vp[4 * g4 + 1], vp[4 * g4 + 2], vp[4 * g4 + 3]);
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const int rl = 8 * i + rsub;
                        apv[i] = tbuf[rl * 4 + (cgl ^ ((rl >> 1) & 3))];
                    }
                    __syncwarp();
                }
                // per-channel constants of my 4 channels
                const int pj = j + 4 * cgl;
                BnC b[4];
                float bt[4] = {0.f, 0.f, 0.f, 0.f}, bp[4] = {0.f, 0.f, 0.f, 0.f};
                if (KIND != EPI_PLAIN) {
                    float4 al, be, sp, tp;
                    if (PRM_DIRECT) {
                        const float4* bnp = reinterpret_cast<const float4*>(ep.bn + cbase + pj);
                        const size_t C4 = (size_t)ep.C / 4;
                        al = __ldg(bnp); be = __ldg(bnp + C4); sp = __ldg(bnp + 2 * C4); tp = __ldg(bnp + 3 * C4);
                    } else {
                        al = *reinterpret_cast<const float4*>(prm + pj); be = *reinterpret_cast<const float4*>(prm + PLD + pj);
                        sp = *reinterpret_cast<const float4*>(prm + 2 * PLD + pj); tp = *reinterpret_cast<const float4*>(prm + 3 * PLD + pj);
                    }
                    b[0] = {al.x, be.x, sp.x, tp.x}; b[1] = {al.y, be.y, sp.y, tp.y}; b[2] = {al.z, be.z, sp.z, tp.z}; b[3] = {al.w, be.w, sp.w, tp.w};
                }
                if (KIND == EPI_PLAIN || KIND == EPI_FWD_DUAL) {
                    float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (!PRM_DIRECT) t = *reinterpret_cast<const float4*>(prm + Cfg::BIAS_ROW * PLD + pj);
                    else if (ep.bias != nullptr) t = __ldg(reinterpret_cast<const float4*>(ep.bias + ncol0 + pj));
                    bt[0] = t.x; bt[1] = t.y; bt[2] = t.z; bt[3] = t.w;
                }
                if (KIND == EPI_FWD_DUAL) {
                    const float4 t = PRM_DIRECT ? __ldg(reinterpret_cast<const float4*>(ep.bias + ncol0 + CH + pj))
                                                : *reinterpret_cast<const float4*>(prm + 5 * PLD + pj);
                    bp[0] = t.x; bp[1] = t.y; bp[2] = t.z; bp[3] = t.w;
                }
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    if (!vrow[i]) continue;
                    const size_t off = (size_t)mrow[i] * ep.C + c;
                    const float av[4] = {at[i].x, at[i].y, at[i].z, at[i].w};
                    const float la[4] = {L.v[0][i].x, L.v[0][i].y, L.v[0][i].z, L.v[0][i].w};
                    float r0[4], r1[4], r2[4];
                    if (KIND == EPI_PLAIN) {
#pragma unroll
                        for (int e = 0; e < 4; ++e) r0[e] = av[e] + bt[e] + la[e];
                    } else if (KIND == EPI_FWD_DUAL) {
                        const float pv[4] = {apv[i].x, apv[i].y, apv[i].z, apv[i].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) {
                            r0[e] = __fadd_rn(av[e], bt[e]);
                            r1[e] = fmaxf(__fadd_rn(pv[e], bp[e]), 0.f);
                            r2[e] = __fadd_rn(__fadd_rn(__fmul_rn(r0[e], b[e].alpha), b[e].beta), la[e]);
                            if (!(ep.hooks & 1)) r2[e] = fmaxf(r2[e], 0.f);
                        }
                    } else if (KIND == EPI_MID) {
                        const float lb[4] = {L.v[NL > 1 ? 1 : 0][i].x, L.v[NL > 1 ? 1 : 0][i].y, L.v[NL > 1 ? 1 : 0][i].z, L.v[NL > 1 ? 1 : 0][i].w};
#pragma unroll
                        for (int e = 0; e < 4; ++e) r0[e] = (dbg & 4) ? av[e] + la[e] + lb[e] : mid_chain(av[e], la[e], lb[e], b[e], mode, ep.eps);
                    } else {
                        const float4 v1 = L.v[NL > 1 ? 1 : 0][i], v2 = L.v[NL > 2 ? 2 : 0][i], v3 = L.v[NL > 3 ? 3 : 0][i];
                        const float lb[4] = {v1.x, v1.y, v1.z, v1.w};
                        const float lc[4] = {v2.x, v2.y, v2.z, v2.w};
                        const float ld[4] = {v3.x, v3.y, v3.z, v3.w};
                        float rr[4] = {0.f, 0.f, 0.f, 0.f};
                        if (mode == XFRB_MODE_ALL && ep.res != nullptr && c < ep.res_c) {
                            const float4 t = __ldg(reinterpret_cast<const float4*>(ep.res + (size_t)msav[i] * ep.res_c + c));
                            rr[0] = t.x; rr[1] = t.y; rr[2] = t.z; rr[3] = t.w;
                        }
                        if (dbg & 4) {
#pragma unroll
                            for (int e = 0; e < 4; ++e) { r0[e] = av[e] + ld[e]; r1[e] = lc[e] + la[e] + lb[e]; }
                        } else {
#pragma unroll
                            for (int e = 0; e < 4; ++e)
                                join_chain(__fadd_rn(av[e], ld[e]), lc[e], la[e], lb[e], rr[e], b[e], ep.hooks & 255, mode, ep.eps, r0[e], r1[e]);
                        }
                    }
                    if (dbg & 2) {      // keep the math alive without storing
                        float chk = r0[0] + r0[1] + r0[2] + r0[3];
                        if (KIND == EPI_FWD_DUAL || KIND == EPI_JOIN) chk += r1[0] + r1[1] + r1[2] + r1[3];
                        if (KIND == EPI_FWD_DUAL) chk += r2[0] + r2[1] + r2[2] + r2[3];
                        if (chk != 1.2345e-30f) continue;
                    }
                    // bf16x2 plan: every tensor that is the A operand of the next GEMM (act, y_out, y3_out) is stored as a
                    // bf16 (hi, lo) pair row - same bytes as fp32 - so that no kernel ever splits an operand again
                    if (PAIRA && KIND == EPI_MID) st_pair4(ep.out0, (size_t)mrow[i], ep.C, c, r0);
                    else *reinterpret_cast<float4*>(ep.out0 + off) = make_float4(r0[0], r0[1], r0[2], r0[3]);
                    if (KIND == EPI_FWD_DUAL || KIND == EPI_JOIN) {
                        if (PAIRA && KIND == EPI_JOIN) st_pair4(ep.out1, (size_t)mrow[i], ep.C, c, r1);
                        else *reinterpret_cast<float4*>(ep.out1 + off) = make_float4(r1[0], r1[1], r1[2], r1[3]);
                    }
                    if (KIND == EPI_FWD_DUAL) {
                        if (PAIRA) {
                            if (ep.out2 != nullptr) st_pair4(ep.out2, (size_t)mrow[i], ep.C, c, r2);
                            if (ep.out3 != nullptr) *reinterpret_cast<float4*>(ep.out3 + off) = make_float4(r2[0], r2[1], r2[2], r2[3]);
                        } else {
                            *reinterpret_cast<float4*>(ep.out2 + off) = make_float4(r2[0], r2[1], r2[2], r2[3]);
                        }
                    }
                }
            };
            // the first slab's loads go out before the accumulator is waited for; afterwards slab j+1 is always in flight
            Loads La, Lb;
            Acc Va, Vb;
            // slab t of this warp starts at column col(t): every (EW/4)-th 16-column slab, or - PAIRED - every (EW/4)-th 32-column
            // group, its two slabs back to back: the warp then touches both 64-byte halves of each 128-byte line of the operand and
            // output tensors within one slab time (the second load hits L1, the two stores merge in L2)
            auto col = [&](int t) { return PAIRED ? part * 32 + (t >> 1) * (32 * (EW / 4)) + (t & 1) * 16 : part * 16 + t * SLAB_STRIDE; };
            // lane (rsub, cgl) copies, and later reads back, the same 16 bytes: [tensor][row 8i + rsub][cgl] - nothing crosses lanes
            auto stage_slab = [&](int j) {
                const int c = cbase + j + 4 * cgl;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const uint32_t dst = smem_u32(ldbuf) + (uint32_t)((8 * i + rsub) * 64 + cgl * 16);
                    const size_t offs = (size_t)msav[i] * ep.C + c;
                    const uint32_t n = vrow[i] ? 16u : 0u;             // invalid rows: zero-fill, nothing is read
                    const float* src[4] = {ep.o + offs, ep.xr + offs, ep.outp + offs, ep.g_res + (size_t)mrow[i] * ep.C + c};
#pragma unroll
                    for (int tz = 0; tz < 4; ++tz)
                        asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst + tz * 2048u), "l"(vrow[i] ? src[tz] : ep.o), "r"(n) : "memory");
                }
                asm volatile("cp.async.commit_group;" ::: "memory");
            };
            auto fetch_slab = [&](Loads& L) {
                asm volatile("cp.async.wait_group 0;" ::: "memory");
#pragma unroll
                for (int i = 0; i < 4; ++i)
#pragma unroll
                    for (int tz = 0; tz < NL; ++tz)
                        L.v[tz][i] = *reinterpret_cast<const float4*>(ldbuf + tz * 2048 + (8 * i + rsub) * 64 + cgl * 16);
            };
            const int j0 = col(0);
            if (LOAD_AHEAD && j0 < CH) issue_loads(j0, La);
            if (Cfg::STAGE_LOADS && !(dbg & 1) && j0 < CH) stage_slab(j0);      // in flight while the main loop of this tile finishes
            if (!PRM_DIRECT) asm volatile("bar.sync 1, %0;" ::"n"(EW * 32) : "memory");      // constants staged
            mbar_wait(tfull_bar(a), aph);
            tc_fence_after();
            if (LOAD_AHEAD) {
                if (ACC_PREFETCH && j0 < CH) request_acc(j0, Va);
#pragma unroll 1
                for (int t = 0; col(t) < CH; t += 2) {
                    const int ja = col(t), jb = col(t + 1), jc = col(t + 2);
                    if (jb < CH) issue_loads(jb, Lb);
                    process(ja, La, Va, jb, ACC_PREFETCH ? Vb : Va);
                    if (jc < CH) issue_loads(jc, La);
                    if (jb < CH) process(jb, Lb, ACC_PREFETCH ? Vb : Va, jc, Va);
                }
            } else {
                // JOIN: four operand tensors per element leave no registers for a second slab of loads; the next slab's
                // lines are pulled into L2 instead (fire-and-forget), so that its loads pay L2 latency, not DRAM latency
                auto prefetch_slab = [&](int j) {
                    if (!(PAIRA && XFRB_PAIRA_JOIN_L2_PREFETCH) || cgl != 0) return;        // one lane per 64-byte row segment
                    const int c = cbase + j;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        if (!vrow[i]) continue;
                        const size_t offs = (size_t)msav[i] * ep.C + c;
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.o + offs));
                        asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.xr + offs));
                        if (KIND == EPI_JOIN) {
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.outp + offs));
                            asm volatile("prefetch.global.L2 [%0];" ::"l"(ep.g_res + (size_t)mrow[i] * ep.C + c));
                        }
                    }
                };
                if (Cfg::STAGE_LOADS && !(dbg & 1)) {
#pragma unroll 1
                    for (int t = 0; col(t) < CH; ++t) {
                        fetch_slab(La);
                        if (col(t + 1) < CH) stage_slab(col(t + 1));
                        process(col(t), La, Va, CH, Va);
                    }
                } else {
#pragma unroll 1
                for (int t = 0; col(t) < CH; ++t) {
                    if (col(t + 1) < CH) prefetch_slab(col(t + 1));
                    issue_loads(col(t), La);
                    process(col(t), La, Va, CH, Va);
                }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (CTA2 && !leader) mbar_arrive_remote(tempty_bar(a), 0);
                else mbar_arrive(tempty_bar(a));
            }
        }
    }

    tc_fence_before();
    if (CLUSTERED) cluster_sync_all();  // neither CTA may leave (or free TMEM) while its peer can still touch its smem / barriers
    else __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        if (CTA2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS));
        else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS));
    }
}

// ------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
    static EncodeTiledFn fn = nullptr;
    if (fn == nullptr) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
            qres == cudaDriverEntryPointSuccess)
            fn = (EncodeTiledFn)p;
    }
    return fn;
}

static bool encode(CUtensorMap* tm, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides_bytes,
                   const cuuint32_t* box, bool bf16 = false) {
    EncodeTiledFn fn = get_encode();
    if (!fn) return false;
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = fn(tm, bf16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank, const_cast<void*>(base), dims, strides_bytes, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    return r == CUDA_SUCCESS;
}

// XFRB_PDL=1: tcgen05 kernels are launched with programmatic stream serialization (see the griddepcontrol pair in the kernel)
inline bool conv_tc_pdl() {
    static const int on = [] { const char* e = getenv("XFRB_PDL"); return e ? atoi(e) : 0; }();
    return on != 0;
}

template <int BN, int SPLIT, int KIND, int MODE = -1>
static cudaError_t launch_cfg(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBlo, const TcGeom& g,
                              const EpiParams& ep, cudaStream_t st) {
    using Cfg = TcCfg<BN, SPLIT, 0, KIND>;
    static int sms_of[XFRB_MAX_DEV] = {};          // 0: this device has not been set up for this instantiation yet
    const int slot = current_device_slot();
    if (sms_of[slot] == 0) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, SPLIT, KIND, 0, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        int dev = 0, n = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
        sms_of[slot] = n;
    }
    const int sms = sms_of[slot];
    int total = g.n_m_tiles * g.n_n_tiles;
    int grid = total < sms ? total : sms;
    if (conv_tc_pdl()) {
        cudaLaunchConfig_t cfg;
        memset(&cfg, 0, sizeof(cfg));
        cudaLaunchAttribute at[1];
        at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        at[0].val.programmaticStreamSerializationAllowed = 1;
        cfg.gridDim = dim3((unsigned)grid);
        cfg.blockDim = dim3(Cfg::THREADS);
        cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
        cfg.stream = st;
        cfg.attrs = at;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, conv_tc_kernel<BN, SPLIT, KIND, 0, MODE>, tmA, tmB, tmBlo, g, ep);
    }
    conv_tc_kernel<BN, SPLIT, KIND, 0, MODE><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tmA, tmB, tmBlo, g, ep);
    return cudaGetLastError();
}

// CTA-pair launch (PAIR 1: cta_group::2, PAIR 2: multicast weights): clusters of 2, persistent over the 256-row pair tiles.
template <int BN, int SPLIT, int KIND, int PAIR = 1, int MODE = -1>
static cudaError_t launch_cfg2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmBlo, const TcGeom& g,
                               const EpiParams& ep, cudaStream_t st) {
    using Cfg = TcCfg<BN, SPLIT, PAIR, KIND>;
    auto kern = conv_tc_kernel<BN, SPLIT, KIND, PAIR, MODE>;
    static int max_clusters_of[XFRB_MAX_DEV] = {};    // 0: not set up on this device yet
    int& max_clusters = max_clusters_of[current_device_slot()];
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.blockDim = dim3(Cfg::THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = st;
    cfg.attrs = at;
    cfg.numAttrs = conv_tc_pdl() ? 2 : 1;
    if (max_clusters <= 0) {
        cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return e;
        int dev = 0, sms = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
        cfg.gridDim = dim3((unsigned)(sms & ~1));
        int n = 0;
        e = cudaOccupancyMaxActiveClusters(&n, kern, &cfg);
        if (e != cudaSuccess) return e;
        max_clusters = n < sms / 2 ? n : sms / 2;
        if (max_clusters < 1) return cudaErrorLaunchOutOfResources;
    }
    int pairs = ((g.n_m_tiles + 1) / 2) * g.n_n_tiles;
    int clusters = pairs < max_clusters ? pairs : max_clusters;
    cfg.gridDim = dim3((unsigned)(2 * clusters));
    return cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmBlo, g, ep);
}

}  // namespace xfrb
