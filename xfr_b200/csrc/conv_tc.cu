// tcgen05 implicit-GEMM convolution for sm_100a: the forward convs and the W+ transposed convs (dgrads) of the whitebox
// sweep with the excitation-backprop hook chains fused into the epilogue.
//
//   D[m, n] = sum_k A[m, k] * B[n, k]      A: NHWC activations (R x R taps, stride 1, zero pad), B: K-major weights
//
// One persistent CTA per SM, warp-specialised in whole warpgroups so that setmaxnreg can move the registers the producers
// do not need to the epilogue (512 threads; 640 for the JOIN kernels):
//   warp 0      TMA producer, activation ring - cp.async.bulk.tensor loads of the A tile (128 pixel rows x 32 channels;
//               for 3x3 convs a 4-D box (32 ch, W, bh image rows, bimg images) shifted by the tap, out-of-bounds rows /
//               columns zero-filled by the TMA unit = the conv padding) into 128B-swizzled shared memory (mbarrier complete_tx)
//   warp 2      TMA producer, weight ring - the B tile (BN x 32, hi plane and, where the plan needs it, the lo plane).  The
//               two rings are independently deep: 3 + 2 stages for the forward dual tiles, 4 + 2 for the W+ dgrads
//   warp 1      MMA issuer - one lane issues tcgen05.mma.kind::tf32 (M = 128, N = BN, K = 8) into TMEM; two accumulator
//               stages so the epilogue of tile i overlaps the main loop of tile i+1; tcgen05.commit releases ring stages and
//               publishes accumulators.  The hi pass of a k-block is issued as soon as its TMA data lands, the lo passes when
//               the split warps are done
//   warps 4-7   operand split - lo = rna_tf32(x - trunc_tf32(x)) of the landed activation tile into a twin buffer; the raw
//               tile itself is the hi operand (tcgen05 truncates fp32 operands to TF32: tools/trunc_probe.py); weights arrive
//               pre-split from the host as two planes
//   warps 8-15  epilogue (8-19 for JOIN) - tcgen05.ld a 32-row x 16-column slab of the accumulator (lane = pixel row),
//               transpose it through swizzled smem so that lanes own consecutive channels, apply the fused EBP epilogue of
//               common.cuh against the saved tensors (their global loads are issued one slab ahead), store NHWC fp32 with
//               128-bit accesses.  The hook mode is a template constant (MODE) in the product plans
//
// SPLIT (pass plan of the GEMM):
//   0  single TF32 pass
//   1  full 3xTF32: A_hi*B_hi + A_lo*B_hi + A_hi*B_lo (signed weights: cancellation amplifies weight rounding)
//   2  W+ GEMMs (non-negative weights and activations, no cancellation): A_hi*B_hi + A_lo*B_hi only, i.e. the exact
//      product with relu(W) rounded to TF32; the lo weight plane is never loaded.  Using the same rounded W+ for the X of
//      the forward twin and for the dgrad keeps excitation backprop mass-conserving.
//      With the dual forward pack (opt-in plan XFRB_IMPL_TF32X2) the signed W half is rounded to TF32 as well: the forward is
//      shared by the mate and the non-mate sweep, so its weight rounding cancels in the contrastive map (tools/hybrid_parity_emul.py).
//   3  dual forward pack [W rows | relu(W) rows]: A_hi*B_hi + A_lo*B_hi over the whole tile, A_hi*B_lo over the W half only
//
// PAIR (clusters of two CTAs working on two consecutive m-tiles of the same n-tile; all variants are bit-identical):
//   0  single CTAs.
//   1  cta_group::2 pair (CTA2; default for the forward dual conv and the MID dgrads): one M = 256 x BN tile per pair.
//      Each CTA loads and splits its own 128 activation rows and stages HALF of the weight tile; the leader issues
//      tcgen05.mma.cta_group::2, whose tensor cores read each CTA's own A and both B halves, so the shared-memory traffic
//      per SM per k-block drops from 224 KB to 160 KB (forward dual) - the split-TF32 main loop is shared-memory-bandwidth
//      bound (profiles/r1_notes.md).  Peer -> leader: remote mbarrier arrives (data landed, split done, accumulator
//      drained); leader -> both: multicast tcgen05.commit.  The remote arrives and the leader's waits are the PLAIN forms:
//      with .release.cluster / .acquire.cluster the same kernel was 1.05-1.5x slower than single CTAs.
//   2  multicast pair (XFRB_MC=1): each CTA's weight producer loads HALF of the weight tile and TMA-multicasts it into both
//      CTAs' rings (the weight ring's empty barriers collect both CTAs' commits through a multicast commit); MMAs, TMEM and
//      epilogue stay private.  Halves the weight traffic out of L2 (2/3 of the ~7.5 TB/s L2 -> SM stream of a 3x3 dgrad) and
//      measured neutral: L2 -> SM bandwidth is not what bounds the main loop.  Off by default.
#include "conv_tc.cuh"

namespace xfrb {

bool conv_tc_available() { return true; }

// Box of a 3x3 conv's activation tile = bimg images x bh image rows x W pixels <= 128 GEMM rows.  Picks the (bh, bimg) that
// wastes the fewest of the 128 rows, counting the ragged last strip and the ragged last image group: 14x14 maps go from 7 rows
// of one image (98 of 128 rows = 77 %) to 1 row of 9 images (126 rows, 97 %), 7x7 maps to 1 row of 18 images (93 %).
// Returns the fraction of useful rows.
double conv_tc_tile_geometry(int H, int W, int Nimg, int* bh_out, int* bimg_out) {
    double best = -1.0;
    int lim = TC_BM / W;
    if (lim > H) lim = H;
    *bh_out = 1;
    *bimg_out = 1;
    for (int bh = lim; bh >= 1; --bh) {         // ties keep the larger strip (more contiguous rows per image)
        int bimg = TC_BM / (bh * W);
        if (bimg > Nimg) bimg = Nimg;
        if (bimg > 256) bimg = 256;
        if (bimg < 1) continue;
        const int th = (H + bh - 1) / bh, ng = (Nimg + bimg - 1) / bimg;
        const double eff = ((double)H / (th * bh)) * ((double)(bh * W * bimg) / TC_BM) * ((double)Nimg / (ng * bimg));
        if (eff > best + 1e-9) { best = eff; *bh_out = bh; *bimg_out = bimg; }
    }
    return best;
}

static int g_mc = -1;        // multicast pairs: -1 from the environment (XFRB_MC=1 enables), else 0 / 1
static bool mc_enabled() {
    if (g_mc < 0) {
        const char* e = getenv("XFRB_MC");
        g_mc = (e != nullptr && e[0] == '1') ? 1 : 0;
    }
    return g_mc != 0;
}
int conv_tc_set_mc(int on) {
    int prev = mc_enabled() ? 1 : 0;
    g_mc = on ? 1 : 0;
    return prev;
}
static int g_cta2 = -1;      // -1: from the environment (XFRB_CTA2=0 disables), else 0 / 1
bool conv_tc_cta2_enabled();
static bool cta2_enabled() { return conv_tc_cta2_enabled(); }
bool conv_tc_cta2_enabled() {
    if (g_cta2 < 0) {
        const char* e = getenv("XFRB_CTA2");
        g_cta2 = (e != nullptr && e[0] == '0') ? 0 : 1;
    }
    return g_cta2 != 0;
}
int conv_tc_set_cta2(int on) {
    int prev = cta2_enabled() ? 1 : 0;
    g_cta2 = on ? 1 : 0;
    return prev;
}

cudaError_t launch_conv_tc(const float* A, const float* B, const ConvGeom& cg, const EpiParams& ep_in, int split, int tn,
                           cudaStream_t st) {
    EpiParams ep = ep_in;
    {   // profiling switches for every epilogue kind (tools/epi_probe.py uses the `hooks` bits of xfrb_dgrad_join directly)
        static int dbg_env = -1;
        if (dbg_env < 0) { const char* e = getenv("XFRB_DBG"); dbg_env = e ? atoi(e) : 0; }
        if (dbg_env) ep.hooks |= dbg_env << 8;
    }
    if (cg.Cin % TC_BK != 0 || split < 0 || split > 3) return cudaErrorInvalidValue;
    if (split == 3 && ep.kind != EPI_FWD_DUAL) return cudaErrorInvalidValue;
    int BN;
    if (ep.kind == EPI_FWD_DUAL) {
        BN = tn;                                  // the dual pack fixes the tile width
        if ((BN != 128 && BN != 256) || cg.Nn % BN) return cudaErrorInvalidValue;
    } else if (cg.Nn % 256 == 0 && tn != 128) BN = 256;
    else if (cg.Nn % 128 == 0) BN = 128;
    else if (cg.Nn % 64 == 0) BN = 64;
    else return cudaErrorInvalidValue;

    TcGeom g;
    g.R = cg.R;
    g.bk = TC_BK;
    g.Cin = cg.Cin;
    g.kchunks = cg.Cin / TC_BK;
    g.num_k = cg.R * cg.R * g.kchunks;
    g.H = cg.H;
    g.W = cg.W;
    g.n_n_tiles = cg.Nn / BN;
    g.b_rows = cg.Nn;
    const int HW = cg.H * cg.W;
    g.Nimg = ep.M / HW;
    CUtensorMap tmA, tmB, tmBlo;
    if (cg.R == 1) {
        g.a4d = 0;
        g.bh = g.bimg = g.tiles_per_img = 1;
        g.n_m_tiles = (ep.M + TC_BM - 1) / TC_BM;
        g.a_bytes = A_TILE_BYTES;
        cuuint64_t dims[2] = {(cuuint64_t)cg.Cin, (cuuint64_t)ep.M};
        cuuint64_t strides[1] = {(cuuint64_t)cg.Cin * 4};
        cuuint32_t box[2] = {TC_BK, TC_BM};
        if (!encode(&tmA, A, 2, dims, strides, box)) return cudaErrorInvalidValue;
    } else {
        g.a4d = 1;
        if (cg.W > 128) return cudaErrorInvalidValue;
        conv_tc_tile_geometry(cg.H, cg.W, g.Nimg, &g.bh, &g.bimg);
        g.tiles_per_img = (cg.H + g.bh - 1) / g.bh;             // row strips per image group
        g.n_m_tiles = ((g.Nimg + g.bimg - 1) / g.bimg) * g.tiles_per_img;
        g.a_bytes = (uint32_t)(TC_BK * cg.W * g.bh * g.bimg * 4);
        cuuint64_t dims[4] = {(cuuint64_t)cg.Cin, (cuuint64_t)cg.W, (cuuint64_t)cg.H, (cuuint64_t)g.Nimg};
        cuuint64_t strides[3] = {(cuuint64_t)cg.Cin * 4, (cuuint64_t)cg.W * cg.Cin * 4, (cuuint64_t)HW * cg.Cin * 4};
        cuuint32_t box[4] = {TC_BK, (cuuint32_t)cg.W, (cuuint32_t)g.bh, (cuuint32_t)g.bimg};
        if (!encode(&tmA, A, 4, dims, strides, box)) return cudaErrorInvalidValue;
    }
    // CTA pairs (cta_group::2) for the split-TF32 plans of the product path when there are enough pair tiles for every TPC
    const bool enough_pairs = ((g.n_m_tiles + 1) / 2) * g.n_n_tiles >= 74;
    // cta_group::2 pairs for the main-loop-bound kinds of the product plans: the forward dual conv and the MID dgrads in the
    // default hook mode (measured per launch: 331 -> 297 us and 370 -> 322 us).  The JOIN dgrads are epilogue-bound and lose
    // (623 -> 776 us as pairs); other hook modes keep the single-CTA kernels specialised per mode.
    const bool cta2 = cta2_enabled() && enough_pairs &&
                      (((split == 3 || split == 2) && ep.kind == EPI_FWD_DUAL) || (split == 2 && ep.kind == EPI_MID && ep.mode == 0));
    // multicast pairs: the product plans, for the default hook mode (other modes keep the single-CTA kernels whose hook chains
    // are specialised per mode)
    const bool mc = !cta2 && (split == 2 || split == 3) && mc_enabled() && enough_pairs &&
                    !(split == 2 && ep.kind == EPI_FWD_DUAL) &&            // the two-pass forward plan has no multicast variant
                    (ep.kind == EPI_FWD_DUAL || ep.kind == EPI_PLAIN || ep.mode == 0);
    const bool half_boxes = cta2 || mc;
    {
        // B is [planes][Nn][K]: plane 0 = rna_tf32(W) (or W itself for single-pass TF32), plane 1 = W - plane 0
        cuuint64_t dims[2] = {(cuuint64_t)cg.K, (cuuint64_t)cg.Nn * (split ? 2 : 1)};
        cuuint64_t strides[1] = {(cuuint64_t)cg.K * 4};
        cuuint32_t box[2] = {TC_BK, (cuuint32_t)(half_boxes ? BN / 2 : BN)};     // a CTA pair loads half of the tile per CTA
        if (!encode(&tmB, B, 2, dims, strides, box)) return cudaErrorInvalidValue;
        tmBlo = tmB;
        if (split == 3) {
            cuuint32_t box_half[2] = {TC_BK, (cuuint32_t)(half_boxes ? BN / 4 : BN / 2)};
            if (!encode(&tmBlo, B, 2, dims, strides, box_half)) return cudaErrorInvalidValue;
        }
    }
    // interleave the m-tiles of the gradient-row groups when every group is a whole number of tiles
    g.groups = 1;
    g.tiles_per_group = g.n_m_tiles;
    if ((ep.kind == EPI_MID || ep.kind == EPI_JOIN) && ep.Ms > 0 && ep.M > ep.Ms && ep.M % ep.Ms == 0) {
        int G = ep.M / ep.Ms;
        if (g.n_m_tiles % G == 0 && (cg.R != 1 || ep.Ms % TC_BM == 0)) {
            g.groups = G;
            g.tiles_per_group = g.n_m_tiles / G;
        }
    }
#define XFRB_TC_PAIR(BN_)                                                                                \
    switch (ep.kind) {                                                                                   \
        case EPI_FWD_DUAL:                                                                               \
            if (split == 2) return launch_cfg2<BN_, 2, EPI_FWD_DUAL, 1>(tmA, tmB, tmBlo, g, ep, st);      \
            return launch_cfg2<BN_, 3, EPI_FWD_DUAL, 1>(tmA, tmB, tmBlo, g, ep, st);                      \
        case EPI_MID: return launch_cfg2<BN_, 2, EPI_MID, 1, 0>(tmA, tmB, tmBlo, g, ep, st);              \
        default: return cudaErrorInvalidValue;                                                           \
    }
    if (cta2) {
        if (BN == 256) { XFRB_TC_PAIR(256) }
        else if (BN == 128) { XFRB_TC_PAIR(128) }
        else { XFRB_TC_PAIR(64) }
    }
#undef XFRB_TC_PAIR
#define XFRB_TC_MC(BN_)                                                                                  \
    switch (ep.kind) {                                                                                   \
        case EPI_FWD_DUAL: return launch_cfg2<BN_, 3, EPI_FWD_DUAL, 2>(tmA, tmB, tmBlo, g, ep, st);       \
        case EPI_PLAIN: return launch_cfg2<BN_, 2, EPI_PLAIN, 2>(tmA, tmB, tmBlo, g, ep, st);             \
        case EPI_MID: return launch_cfg2<BN_, 2, EPI_MID, 2, 0>(tmA, tmB, tmBlo, g, ep, st);              \
        case EPI_JOIN: return launch_cfg2<BN_, 2, EPI_JOIN, 2, 0>(tmA, tmB, tmBlo, g, ep, st);            \
        default: return cudaErrorInvalidValue;                                                           \
    }
    if (mc) {
        if (BN == 256) { XFRB_TC_MC(256) }
        else if (BN == 128) { XFRB_TC_MC(128) }
        else { XFRB_TC_MC(64) }
    }
#undef XFRB_TC_MC
#define XFRB_TC_KINDS(BN_, SP_)                                                                          \
    switch (ep.kind) {                                                                                   \
        case EPI_PLAIN: return launch_cfg<BN_, SP_, EPI_PLAIN>(tmA, tmB, tmBlo, g, ep, st);               \
        case EPI_MID: return launch_cfg<BN_, SP_, EPI_MID>(tmA, tmB, tmBlo, g, ep, st);                   \
        case EPI_JOIN: return launch_cfg<BN_, SP_, EPI_JOIN>(tmA, tmB, tmBlo, g, ep, st);                 \
        default: break;                                                                                  \
    }
#define XFRB_TC_DISPATCH(BN_)                                                                            \
    if (ep.kind == EPI_FWD_DUAL) {                                                                       \
        if (split == 0) return launch_cfg<BN_, 0, EPI_FWD_DUAL>(tmA, tmB, tmBlo, g, ep, st);              \
        if (split == 1) return launch_cfg<BN_, 1, EPI_FWD_DUAL>(tmA, tmB, tmBlo, g, ep, st);              \
        if (split == 2) return launch_cfg<BN_, 2, EPI_FWD_DUAL>(tmA, tmB, tmBlo, g, ep, st);              \
        return launch_cfg<BN_, 3, EPI_FWD_DUAL>(tmA, tmB, tmBlo, g, ep, st);                              \
    }                                                                                                    \
    if (split == 0) { XFRB_TC_KINDS(BN_, 0) }                                                            \
    else if (split == 1) { XFRB_TC_KINDS(BN_, 1) }                                                       \
    else {                                                                                               \
        /* the product plan: hook chains specialised per ebp_subtree_mode */                             \
        if (ep.kind == EPI_MID && ep.mode == 0) return launch_cfg<BN_, 2, EPI_MID, 0>(tmA, tmB, tmBlo, g, ep, st);    \
        if (ep.kind == EPI_MID && ep.mode == 1) return launch_cfg<BN_, 2, EPI_MID, 1>(tmA, tmB, tmBlo, g, ep, st);    \
        if (ep.kind == EPI_MID && ep.mode == 2) return launch_cfg<BN_, 2, EPI_MID, 2>(tmA, tmB, tmBlo, g, ep, st);    \
        if (ep.kind == EPI_JOIN && ep.mode == 0) return launch_cfg<BN_, 2, EPI_JOIN, 0>(tmA, tmB, tmBlo, g, ep, st);  \
        if (ep.kind == EPI_JOIN && ep.mode == 1) return launch_cfg<BN_, 2, EPI_JOIN, 1>(tmA, tmB, tmBlo, g, ep, st);  \
        if (ep.kind == EPI_JOIN && ep.mode == 2) return launch_cfg<BN_, 2, EPI_JOIN, 2>(tmA, tmB, tmBlo, g, ep, st);  \
        XFRB_TC_KINDS(BN_, 2)                                                                            \
    }                                                                                                    \
    return cudaErrorInvalidValue;
    if (BN == 256) { XFRB_TC_DISPATCH(256) }
    else if (BN == 128) { XFRB_TC_DISPATCH(128) }
    else { XFRB_TC_DISPATCH(64) }
#undef XFRB_TC_KINDS
#undef XFRB_TC_DISPATCH
}

}  // namespace xfrb
