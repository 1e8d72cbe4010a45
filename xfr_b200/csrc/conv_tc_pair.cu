// bf16x2 plan of the tcgen05 implicit-GEMM convolution (kernel template: conv_tc.cuh, design notes: conv_tc.cu).
//
// Excitation backprop is a ratio in which the same W+ appears in the numerator and the denominator, so it tolerates coarse
// WEIGHTS but not coarse ACTIVATIONS (tools/bf16_plan_emul.py).  This plan therefore carries every GEMM operand as bf16 terms:
//   activations / gradients   2 terms (hi, lo): 16 significant bits, written ONCE by the producing epilogue as a "pair" row
//                             [C bf16 hi | C bf16 lo] (same bytes as fp32) - no kernel splits an operand in shared memory again
//   W+ = relu(W)              1 term
//   signed W (true forward)   2 terms
// and runs tcgen05.mma.kind::f16 (bf16 in, fp32 accumulate in TMEM), K = 64 per k-block, at twice the kind::tf32 rate
// (measured: tools/mma_peak.cu, 2,232 vs 1,116 TFLOP/s):
//   SPLIT 4  W+ GEMMs (MID / JOIN / PLAIN dgrads):  A_hi*B + A_lo*B                      = 1.0 TF32-pass equivalents (was 2.0)
//   SPLIT 5  forward dual pack [W | relu(W)]:        A_hi*B_hi + A_lo*B_hi + A_hi*B_lo|W  = 1.25                      (was 2.5)
// The operand-split warps of the TF32 plans have nothing to do here (in a CTA pair they only relay "landed" to the leader).
#include "conv_tc.cuh"

namespace xfrb {

cudaError_t launch_conv_tc_pair(const float* A, const void* B, const ConvGeom& cg, const EpiParams& ep_in, int tn, cudaStream_t st) {
    EpiParams ep = ep_in;
    {
        static int dbg_env = -1;
        if (dbg_env < 0) { const char* e = getenv("XFRB_DBG"); dbg_env = e ? atoi(e) : 0; }
        if (dbg_env) ep.hooks |= dbg_env << 8;
    }
    constexpr int BK = 64;                        // bf16 elements per 128-byte swizzle row
    if (cg.Cin % BK != 0) return cudaErrorInvalidValue;
    const bool dual = ep.kind == EPI_FWD_DUAL;
    int BN;
    if (dual) {
        BN = tn;
        if ((BN != 128 && BN != 256) || cg.Nn % BN) return cudaErrorInvalidValue;
    } else if (cg.Nn % 256 == 0 && tn != 128) BN = 256;
    else if (cg.Nn % 128 == 0) BN = 128;
    else if (cg.Nn % 64 == 0) BN = 64;
    else return cudaErrorInvalidValue;

    TcGeom g;
    g.R = cg.R;
    g.bk = BK;
    g.Cin = cg.Cin;
    g.kchunks = cg.Cin / BK;
    g.num_k = cg.R * cg.R * g.kchunks;
    g.H = cg.H;
    g.W = cg.W;
    g.n_n_tiles = cg.Nn / BN;
    g.b_rows = cg.Nn;
    const int HW = cg.H * cg.W;
    g.Nimg = ep.M / HW;
    CUtensorMap tmA, tmB, tmBlo;
    // A: pair rows of 2*Cin bf16 (hi half-row, lo half-row); the lo box sits Cin elements to the right of the hi box
    if (cg.R == 1) {
        g.a4d = 0;
        g.bh = g.bimg = g.tiles_per_img = 1;
        g.n_m_tiles = (ep.M + TC_BM - 1) / TC_BM;
        g.a_bytes = A_TILE_BYTES;
        cuuint64_t dims[2] = {(cuuint64_t)2 * cg.Cin, (cuuint64_t)ep.M};
        cuuint64_t strides[1] = {(cuuint64_t)cg.Cin * 4};
        cuuint32_t box[2] = {BK, TC_BM};
        if (!encode(&tmA, A, 2, dims, strides, box, true)) return cudaErrorInvalidValue;
    } else {
        g.a4d = 1;
        if (cg.W > 128) return cudaErrorInvalidValue;
        conv_tc_tile_geometry(cg.H, cg.W, g.Nimg, &g.bh, &g.bimg);
        g.tiles_per_img = (cg.H + g.bh - 1) / g.bh;
        g.n_m_tiles = ((g.Nimg + g.bimg - 1) / g.bimg) * g.tiles_per_img;
        g.a_bytes = (uint32_t)(BK * 2 * cg.W * g.bh * g.bimg);
        cuuint64_t dims[4] = {(cuuint64_t)2 * cg.Cin, (cuuint64_t)cg.W, (cuuint64_t)cg.H, (cuuint64_t)g.Nimg};
        cuuint64_t strides[3] = {(cuuint64_t)cg.Cin * 4, (cuuint64_t)cg.W * cg.Cin * 4, (cuuint64_t)HW * cg.Cin * 4};
        cuuint32_t box[4] = {BK, (cuuint32_t)cg.W, (cuuint32_t)g.bh, (cuuint32_t)g.bimg};
        if (!encode(&tmA, A, 4, dims, strides, box, true)) return cudaErrorInvalidValue;
    }
    // Small problems (batch-1 calls: a few hundred GEMM rows): a 256-wide tile leaves a handful of CTAs streaming the whole weight
    // matrix through one SM each (26 - 42 us per launch, profiles/r2r_ncu_launches_batch1.csv) - narrower tiles spread the weight
    // rows over more SMs.  The k-loop of an output element is unchanged, so the results are bit-identical.  The dual forward pack
    // is tiled for its BN and keeps it.
    // The K >= 1,024 MID dgrads (conv3 dgrad of layer3 / layer4: 256 or 512 output columns) run on 128-wide tiles: twice the tiles -
    // 10.6 rounds over the SMs instead of 5.3 run as 6 - each streaming the same A for half the MMA work: 174 -> 153 us per layer3
    // launch (gpurun_out/r2ac).  Same k-loop per output element: bit-identical.  XFRB_MID_BN=256 restores the wide tiles.
    static const int mid_bn = [] { const char* e = getenv("XFRB_MID_BN"); return e ? atoi(e) : 128; }();
    if (mid_bn == 128 && ep.kind == EPI_MID && cg.R == 1 && cg.K >= 1024 && BN == 256) { BN = 128; g.n_n_tiles = cg.Nn / BN; }
    static const int join_bn = [] { const char* e = getenv("XFRB_JOIN_BN"); return e ? atoi(e) : 0; }();      // A/B probe: JOIN dgrads on 128-wide tiles
    if (join_bn == 128 && ep.kind == EPI_JOIN && BN == 256) { BN = 128; g.n_n_tiles = cg.Nn / BN; }
    static const int small_bn = [] { const char* e = getenv("XFRB_SMALL_BN"); return e ? atoi(e) : 1; }();
    if (!dual && small_bn) {
        while (BN > 64 && g.n_m_tiles * (cg.Nn / BN) < 64 && cg.Nn % (BN / 2) == 0) BN /= 2;
        g.n_n_tiles = cg.Nn / BN;
    }
    // CTA pairs whenever there is at least one pair of m-tiles: measured faster at both ends - 256-probe sweeps and
    // batch-1 calls (4.27 -> 4.05 ms per contrastive_ebp: each SM stages half of the weight tile); XFRB_PAIR_MIN=74 is the old rule
    static const int pair_min = [] { const char* e = getenv("XFRB_PAIR_MIN"); return e ? atoi(e) : 1; }();
    const bool enough_pairs = ((g.n_m_tiles + 1) / 2) * g.n_n_tiles >= pair_min;
    static const int pair_kinds = [] { const char* e = getenv("XFRB_PAIR_KINDS"); return e ? atoi(e) : 3; }();   // bit 0: forward, bit 1: MID
    const bool cta2 = conv_tc_cta2_enabled() && enough_pairs &&
                      ((dual && (pair_kinds & 1)) || (ep.kind == EPI_MID && ep.mode == 0 && (pair_kinds & 2)));
    {
        cuuint64_t dims[2] = {(cuuint64_t)cg.K, (cuuint64_t)cg.Nn * (dual ? 2 : 1)};
        cuuint64_t strides[1] = {(cuuint64_t)cg.K * 2};
        cuuint32_t box[2] = {BK, (cuuint32_t)(cta2 ? BN / 2 : BN)};
        if (!encode(&tmB, B, 2, dims, strides, box, true)) return cudaErrorInvalidValue;
        tmBlo = tmB;
        if (dual) {
            cuuint32_t box_half[2] = {BK, (cuuint32_t)(cta2 ? BN / 4 : BN / 2)};
            if (!encode(&tmBlo, B, 2, dims, strides, box_half, true)) return cudaErrorInvalidValue;
        }
    }
    g.groups = 1;
    g.tiles_per_group = g.n_m_tiles;
    if ((ep.kind == EPI_MID || ep.kind == EPI_JOIN) && ep.Ms > 0 && ep.M > ep.Ms && ep.M % ep.Ms == 0) {
        int G = ep.M / ep.Ms;
        if (g.n_m_tiles % G == 0 && (cg.R != 1 || ep.Ms % TC_BM == 0)) {
            g.groups = G;
            g.tiles_per_group = g.n_m_tiles / G;
        }
    }
#define XFRB_PAIR_DISPATCH(BN_)                                                                                         \
    if (cta2) {                                                                                                         \
        if (dual) return launch_cfg2<BN_, 5, EPI_FWD_DUAL, 1>(tmA, tmB, tmBlo, g, ep, st);                              \
        return launch_cfg2<BN_, 4, EPI_MID, 1, 0>(tmA, tmB, tmBlo, g, ep, st);                                          \
    }                                                                                                                   \
    switch (ep.kind) {                                                                                                  \
        case EPI_FWD_DUAL: return launch_cfg<BN_, 5, EPI_FWD_DUAL>(tmA, tmB, tmBlo, g, ep, st);                         \
        case EPI_PLAIN: return launch_cfg<BN_, 4, EPI_PLAIN>(tmA, tmB, tmBlo, g, ep, st);                               \
        case EPI_MID:                                                                                                   \
            if (ep.mode == 0) return launch_cfg<BN_, 4, EPI_MID, 0>(tmA, tmB, tmBlo, g, ep, st);                        \
            if (ep.mode == 1) return launch_cfg<BN_, 4, EPI_MID, 1>(tmA, tmB, tmBlo, g, ep, st);                        \
            if (ep.mode == 2) return launch_cfg<BN_, 4, EPI_MID, 2>(tmA, tmB, tmBlo, g, ep, st);                        \
            return launch_cfg<BN_, 4, EPI_MID>(tmA, tmB, tmBlo, g, ep, st);                                             \
        case EPI_JOIN:                                                                                                  \
            if (ep.mode == 0) return launch_cfg<BN_, 4, EPI_JOIN, 0>(tmA, tmB, tmBlo, g, ep, st);                       \
            if (ep.mode == 1) return launch_cfg<BN_, 4, EPI_JOIN, 1>(tmA, tmB, tmBlo, g, ep, st);                       \
            if (ep.mode == 2) return launch_cfg<BN_, 4, EPI_JOIN, 2>(tmA, tmB, tmBlo, g, ep, st);                       \
            return launch_cfg<BN_, 4, EPI_JOIN>(tmA, tmB, tmBlo, g, ep, st);                                            \
        default: return cudaErrorInvalidValue;                                                                          \
    }
    if (BN == 256) { XFRB_PAIR_DISPATCH(256) }
    else if (BN == 128) { XFRB_PAIR_DISPATCH(128) }
    else { XFRB_PAIR_DISPATCH(64) }
#undef XFRB_PAIR_DISPATCH
}

}  // namespace xfrb
