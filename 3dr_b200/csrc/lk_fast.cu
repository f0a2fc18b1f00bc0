// Specialised pyramidal-LK tracker kernels (compile-time window sizes) for sm_100a -- warp per feature, all levels
// in one launch.  Same arithmetic as lk_generic.cu / OpenCV's LKTrackerInvoker (reference call
// site src/initialization.cpp:608-613), bit-identical results, but organised for the SM:
//
//   * Per level the warp stages three regions in shared memory with 16-byte row-coalesced loads (every image row of
//     the window is touched once per level): the template window of the previous image (u8), its Scharr derivatives
//     (packed s16x2) and a search region of the next image (window + >= 11 px margin in x, 8 px in y).  Every level
//     carries an apron (reflected pixels / zero derivatives, dr3lk_internal.cuh), so staging is plain 16-byte row
//     copies (cp.async) with one address computation per region and neither staging nor compute ever branches on
//     borders.  The search region is re-staged only when the window leaves it.
//   * The window is cut into runs of R consecutive pixels of one row; each lane owns NRUN runs.  A run's R+1 source
//     bytes of two rows are fetched as 3-4 aligned 32-bit shared loads and realigned with PRMT, and the 14-bit
//     bilinear blend is two IDP.2A (dp2a: 2 x (s16 weight * u8 pixel)) per pixel instead of four IMADs.
//   * The template (Ix, Iy per pixel) lives in registers for the whole level, and so does 2^8 - 512 * I per pixel: used
//     as the dp2a addend it makes the shifted blend come out as (J - I) directly (a multiple of 512 passes through the
//     arithmetic >> 9 exactly), so an iteration costs 2 IDP + 1 shift + 2 IMAD per pixel.  All sums are exact
//     integers: per-lane int32 partials, warp totals through REDUX on the 16-bit halves, one rounding to fp32 --
//     results do not depend on summation order.
//   * Scalar fp32 steps use explicitly rounded intrinsics (no FMA contraction) and match the CPU oracle bit for bit.
#include <float.h>
#include <stdlib.h>

#include <algorithm>
#include <mutex>

#include "dr3lk_internal.cuh"

namespace dr3lk {

namespace {


constexpr int pitch_words_for(int need)  // smallest p >= need with p % 8 == 4: 16-B aligned rows, rows 4 banks apart
{
    int p = need;
    while (p % 8 != 4) p++;
    return p;
}

__host__ __device__ constexpr int pow2_at_least(int v) { return v <= 1 ? 1 : (v <= 2 ? 2 : (v <= 4 ? 4 : (v <= 8 ? 8 : (v <= 16 ? 16 : 32)))); }
// rows a region of `rows` x `ch` 16-byte chunks occupies when staged in whole rounds of 32 lanes
constexpr int staged_rows(int rows, int ch) { return (rows + 32 / pow2_at_least(ch) - 1) / (32 / pow2_at_least(ch)) * (32 / pow2_at_least(ch)); }

template <int WW_, int WH_, int R_>
struct Geo {
    static constexpr int WW = WW_, WH = WH_, R = R_;
    static constexpr int NCB = (WW + R - 1) / R;         // runs per window row
    static constexpr int NRUNS = NCB * WH;
    static constexpr int NRUN = (NRUNS + 31) / 32;       // runs per lane
    static constexpr int NWD = (R + 1 + 3 + 3) / 4;      // 32-bit words covering R+1 bytes at any byte alignment
    static constexpr int NEO = (R + 3) / 4;              // realigned registers per parity
    static constexpr bool RAGGED = (WW % R) != 0;
    static constexpr int WARPS = 4;                      // warps per CTA (each warp is an independent worker)
    // keep the search window's realigned bytes in registers across iterations: -4 % on 21x21 (16 registers), but
    // -15 % throughput on 31x31 where the 32 extra registers do not fit next to the template
    static constexpr bool CACHE_J = NRUN * R <= 16;
#ifndef DR3LK_BLOCKS_SMALL
#define DR3LK_BLOCKS_SMALL 4
#endif
#ifndef DR3LK_BLOCKS_LARGE
#define DR3LK_BLOCKS_LARGE 3
#endif
    static constexpr int MIN_BLOCKS = (NRUN * R <= 16) ? DR3LK_BLOCKS_SMALL : DR3LK_BLOCKS_LARGE;  // measured best: 16 warps/SM at <= 128 registers (21x21), 12 at <= 168 (31x31, 30x30)
    // Search-region margins (the x margin is >= MX after alignment): how far the window may move at one level before the
    // region is staged again.  21x21: 5 px each way -- a 48-byte x 32-row region instead of 64 x 40 (one staging round
    // and 40 % of the region's L2 traffic less: 191.4 -> 189.1 ms on C3; the coarse-to-fine estimate moves the window by
    // a pixel or two per level).  The larger windows keep wider margins (their aprons are sized for them).
    static constexpr int MX = WW_ == 21 ? 5 : 13, MY = WW_ == 21 ? 5 : 8;
    static constexpr int J_CH = (WW + 1 + 2 * MX + 15 + 15) / 16;  // 16-B chunks per search-region row
    static constexpr int J_W = J_CH * 16;
    static constexpr int J_H = WH + 1 + 2 * MY;
    static constexpr int J_PW = pitch_words_for(J_CH * 4);
    static constexpr int I_CH = (WW + 1 + 15 + 15) / 16;
    static constexpr int I_W = I_CH * 16;
    static constexpr int I_PW = pitch_words_for(I_CH * 4);
    static constexpr int D_CH = (WW + 1 + 3 + 3) / 4;    // chunks of 4 derivative words
    static constexpr int D_PW = pitch_words_for(D_CH * 4);
    // Staging copies whole rounds of 32 lanes (no partial round, no idle-lane branch): regions are rounded up to whole
    // rounds of rows; the surplus rows are copied (the apron makes them addressable) and never read.
    static constexpr int J_ROWS = staged_rows(J_H, J_CH), I_ROWS = staged_rows(WH + 1, I_CH), D_ROWS = staged_rows(WH + 1, D_CH);
    // region sizes in words, rounded to 128 bytes (TMA destinations must be 128-byte aligned)
    static constexpr int J_WORDS = (J_PW * J_ROWS + 4 + 31) / 32 * 32;    // +4: realignment may read one word past the last row
    static constexpr int I_WORDS = (I_PW * I_ROWS + 4 + 31) / 32 * 32;
    static constexpr int D_ZERO = D_PW * D_ROWS;         // two rows of zeros for runs that do not exist
    static constexpr int D_WORDS = (D_PW * (D_ROWS + 2) + 4 + 31) / 32 * 32;
    // TMA staging (cp.async.bulk.tensor): one box per region, exactly the rows that are read; the box rows land densely,
    // which IS the padded layout above because every region is an odd number of 16-byte chunks wide
    static_assert(J_PW == J_CH * 4 && I_PW == I_CH * 4 && D_PW == D_CH * 4, "TMA boxes need dense region rows");
    static constexpr int TMA_I_BYTES = I_W * (WH + 1), TMA_D_BYTES = D_CH * 16 * (WH + 1), TMA_J_BYTES = J_W * J_H;
    static constexpr int PX = kApronX, PY = apron_y(WH), DPX = deriv_apron_x(WW);  // aprons of every level (bytes, rows, ints)
    static_assert(PY >= I_ROWS && PY >= D_ROWS && PY >= J_ROWS - 2 * MY, "apron rows must cover the rounded-up staging");
    // smallest level the kernel accepts: the aprons are filled by ONE reflection of the level
    static constexpr int MIN_H = PY + 1;
    static constexpr int MIN_W = PX + 1;
    static_assert(PX >= WW + 1 && PX % 16 == 0 && DPX >= WW + 1 && DPX % 4 == 0, "aprons must cover a window hanging over the border");
    static constexpr int WARP_WORDS = J_WORDS + I_WORDS + D_WORDS;
    static_assert((long long)NRUN * R * 8160LL * 4080LL < 2147483647LL, "per-lane int32 partial sums would overflow");
    static_assert(J_W >= WW + 1 + MX + MX + 15, "search region too narrow");
};

__device__ __forceinline__ int dp2a_lo(int a, unsigned b, int c)
{
    int d;
    asm("dp2a.lo.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ int dp2a_hi(int a, unsigned b, int c)
{
    int d;
    asm("dp2a.hi.s32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// exact warp total of per-lane int32 partials, as (hi, lo) with total = hi * 65536 + lo
struct HiLo {
    int hi, lo;
};
__device__ __forceinline__ HiLo warp_sum_hilo(int v)
{
    HiLo r;
    r.lo = __reduce_add_sync(0xffffffffu, v & 0xffff);
    r.hi = __reduce_add_sync(0xffffffffu, v >> 16);
    return r;
}
// rn_fp32(hi * 65536 + lo): both conversions are exact (|hi| < 2^24, |lo| < 2^24), the fma rounds once
__device__ __forceinline__ float hilo_to_float(int hi, int lo) { return __fmaf_rn((float)hi, 65536.f, (float)lo); }

struct Weights {
    int w00, w01, w10, w11;
    int wt, wb;  // (w00 | w01 << 16), (w10 | w11 << 16) for dp2a
};
__device__ __forceinline__ Weights make_weights(float a, float b)
{
    Weights q;
    // rn(x * y) * 2^14 == rn(x * (y * 2^14)): scaling by a power of two is exact (no under/overflow here), so the
    // scale is folded into one factor; the products round exactly as OpenCV's (1-a)*(1-b)*(1<<14) does.
    const float oma = __fsub_rn(1.f, a);
    const float ombs = __fmul_rn(__fsub_rn(1.f, b), 16384.f), bs = __fmul_rn(b, 16384.f);
    q.w00 = __float2int_rn(__fmul_rn(oma, ombs));
    q.w01 = __float2int_rn(__fmul_rn(a, ombs));
    q.w10 = __float2int_rn(__fmul_rn(oma, bs));
    q.w11 = (1 << W_BITS) - q.w00 - q.w01 - q.w10;
    q.wt = (int)__byte_perm((unsigned)q.w00, (unsigned)q.w01, 0x5410);  // low halves: w00 | w01 << 16
    q.wb = (int)__byte_perm((unsigned)q.w10, (unsigned)q.w11, 0x5410);
    return q;
}

// R+1 consecutive bytes of two shared-memory rows (byte offset `boff` in a region of `PW` words per row), realigned so
// that pixel k's (p[k], p[k+1]) pair sits in the low or high half of a register: even k -> E[k/4], odd k -> O[k/4],
// high half when (k & 2).
template <int R, int NWD, int NEO, int PW>
struct RunBytes {
    unsigned tE[NEO], tO[NEO], bE[NEO], bO[NEO];
    __device__ __forceinline__ void load(const unsigned* __restrict__ region, int boff, int region_words = 0x7fffffff, int staged_bytes = 0x7fffffff)
    {
        // the two rows start inside the staged rows; the last realignment word may lie behind them but inside the region
        DR3LK_CHECK(boff >= 0 && boff + PW * 4 + R + 1 <= staged_bytes && (boff >> 2) + PW + NWD <= region_words, 3, boff);
        DR3LK_CHECK_COUNT();
        const unsigned* p = region + (boff >> 2);
        const unsigned o = boff & 3;
        const unsigned selE = 0x3210u + o * 0x1111u, selO = selE + 0x1111u;
        unsigned t[NWD], b[NWD];
#pragma unroll
        for (int i = 0; i < NWD; i++) { t[i] = p[i]; b[i] = p[PW + i]; }
#pragma unroll
        for (int j = 0; j < NEO; j++) {
            tE[j] = __byte_perm(t[j], t[j + 1 < NWD ? j + 1 : j], selE);
            bE[j] = __byte_perm(b[j], b[j + 1 < NWD ? j + 1 : j], selE);
            tO[j] = __byte_perm(t[j], t[j + 1 < NWD ? j + 1 : j], selO);
            bO[j] = __byte_perm(b[j], b[j + 1 < NWD ? j + 1 : j], selO);
        }
    }
    // (S + init) >> 9 for pixel k of the run.  init = 2^8 gives the window value with 5 fractional bits; init =
    // 2^8 - 512 * I gives (J - I) directly: subtracting a multiple of 512 before the arithmetic shift is exact.
    __device__ __forceinline__ int sample(int k, const Weights& q, int init = 1 << (W_BITS - 5 - 1)) const
    {
        return sum(k, q, init) >> (W_BITS - 5);
    }
    // S + init, before the shift
    __device__ __forceinline__ int sum(int k, const Weights& q, int init = 1 << (W_BITS - 5 - 1)) const
    {
        const unsigned tt = (k & 1) ? tO[k >> 2] : tE[k >> 2];
        const unsigned bb = (k & 1) ? bO[k >> 2] : bE[k >> 2];
        int s;
        if (k & 2) { s = dp2a_hi(q.wt, tt, init); s = dp2a_hi(q.wb, bb, s); }
        else { s = dp2a_lo(q.wt, tt, init); s = dp2a_lo(q.wb, bb, s); }
        return s;
    }
};

__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// ---- TMA staging primitives (per-warp mbarriers: one elected lane arms the barrier and issues the boxes) ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, int bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, unsigned parity)
{
    // the whole wait loop in PTX (try_wait blocks for a hardware-defined time before it returns false)
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// exactly one lane of the (converged) warp: lets ptxas issue the TMA from the uniform datapath without a divergence region
__device__ __forceinline__ bool elect_one()
{
    unsigned pred;
    asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; selp.u32 %0, 1, 0, p; }" : "=r"(pred));
    return pred != 0;
}
// one box of a rank-3 tensor [pair][row][column] at (x, y, pair) -> dense rows at dst (128-byte aligned)
__device__ __forceinline__ void tma_box(void* dst, const CUtensorMap* map, int x, int y, int pair, uint64_t* bar)
{
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                 ::"r"(smem_u32(dst)), "l"(reinterpret_cast<unsigned long long>(map)), "r"(x), "r"(y), "r"(pair), "r"(smem_u32(bar)) : "memory");
}

template <typename G>
struct Tracker {
    static constexpr int WW = G::WW, WH = G::WH;

    // ROWS (a whole number of rounds) rows of CH 16-byte chunks, LPR (power of two) lanes per row: global -> shared
    // with cp.async (LDGSTS: no registers, completion tracked per commit group).  `g` points at the first chunk of the
    // first row; thanks to the aprons every row is a plain copy.  Surplus lanes of a row repeat its last chunk (same
    // bytes to the same place), so there is no divergence; offsets are 32-bit (a level image is < 2 GiB).
    template <int ROWS, int CH, int PITCH_BYTES>
    static __device__ __forceinline__ void stage_rows(void* region, int lane, const uint8_t* g, int pitch_bytes)
    {
        constexpr int LPR = pow2_at_least(CH), RPR = 32 / LPR, ROUNDS = ROWS / RPR;
        static_assert(ROWS % RPR == 0, "regions are staged in whole rounds");
        const int ch = CH < LPR ? min(lane & (LPR - 1), CH - 1) : (lane & (LPR - 1)), rr = lane / LPR;
        const unsigned sbase = (unsigned)__cvta_generic_to_shared(region) + rr * PITCH_BYTES + ch * 16;
        g += rr * pitch_bytes + ch * 16;
        const long long step = RPR * pitch_bytes;
#pragma unroll
        for (int i = 0; i < ROUNDS; i++) {
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sbase + i * RPR * PITCH_BYTES), "l"(g) : "memory");
            g += step;
        }
    }

    // ---- staging -------------------------------------------------------------------------------------------
    // All stage_* functions only ISSUE the copies; the caller commits the group, waits for it and __syncwarp()s
    // before any lane reads the region.  Coordinates are relative to pixel (0, 0) of the level; the aprons make
    // x in [-PX, w + PX) and y in [-PY, h + PY) addressable.
    // Search region of the next image around window origin (inx, iny): J_W x J_H bytes from (rx0, ry0); window origins
    // rx0 .. rx0 + J_W - (WW+1), ry0 .. ry0 + 2*MY are inside it.
    // checked build: the rectangle [first, first + rows x width) read with row step `pitch` lies inside the allocation [lo, hi)
    // and inside its rows (x0_in_row = byte offset of the rectangle's first column inside an allocation row)
    static __device__ __forceinline__ void check_rect(const uint8_t* first, int rows, int width, int pitch, int x0_in_row, const uint8_t* lo,
                                                      const uint8_t* hi, int kind)
    {
        DR3LK_CHECK(first >= lo && first + (long long)(rows - 1) * pitch + width <= hi && x0_in_row >= 0 && x0_in_row + width <= pitch, kind,
                    first - lo);
        DR3LK_CHECK_COUNT();
    }
    static __device__ __forceinline__ void origin_J(int pitch, int h, int inx, int iny, int& rx0, int& ry0)
    {
        rx0 = max(-G::PX, min((inx - G::MX) & ~15, pitch - G::PX - G::J_W));
        ry0 = max(-G::PY, min(iny - G::MY, h + G::PY - G::J_ROWS));
    }
    static __device__ __forceinline__ void stage_J(unsigned* sJ, const uint8_t* __restrict__ img, int pitch, int h, int inx, int iny, int lane,
                                                   int& rx0, int& ry0)
    {
        origin_J(pitch, h, inx, iny, rx0, ry0);
        stage_rows<G::J_ROWS, G::J_CH, G::J_PW * 4>(sJ, lane, img + (ry0 * pitch + rx0), pitch);
    }
    // first staged column of the template window / its derivatives for window origin x = ipx
    static __device__ __forceinline__ int x0_I(int ipx, int pitch) { return min(ipx & ~15, pitch - G::PX - G::I_W); }
    static __device__ __forceinline__ int x0_D(int ipx, int dpitch) { return min(ipx & ~3, dpitch - G::DPX - G::D_CH * 4); }

    // Template window of the previous image: WH+1 rows from ipy, I_W bytes from x0_I
    static __device__ __forceinline__ void stage_I(unsigned* sI, const uint8_t* __restrict__ img, int pitch, int ipx, int ipy, int lane)
    {
        stage_rows<G::I_ROWS, G::I_CH, G::I_PW * 4>(sI, lane, img + (ipy * pitch + x0_I(ipx, pitch)), pitch);
    }
    // Scharr derivatives of the template window (the apron holds the zeros outside the image): WH+1 rows, D_CH*4 words
    static __device__ __forceinline__ void stage_D(unsigned* sD, const int* __restrict__ der, int dpitch, int ipx, int ipy, int lane)
    {
        stage_rows<G::D_ROWS, G::D_CH, G::D_PW * 4>(sD, lane, reinterpret_cast<const uint8_t*>(der + (ipy * dpitch + x0_D(ipx, dpitch))), dpitch * 4);
    }
};

// Persistent warps: every warp pulls the next feature index from a global counter until the batch is exhausted, so a
// slow feature (many iterations) never holds other warps' slots.
// TMA = true: the three regions are staged with cp.async.bulk.tensor boxes (one elected lane, completion on per-warp
// mbarriers) instead of 11 - 13 cp.async rounds of all 32 lanes; same shared-memory layout, same arithmetic.
template <typename G, bool TMA>
__global__ void __launch_bounds__(G::WARPS * 32, G::MIN_BLOCKS)
lk_fast_kernel(const __grid_constant__ LKParams P)
{
    constexpr int WW = G::WW, WH = G::WH, R = G::R, NRUN = G::NRUN;
    using T = Tracker<G>;
    extern __shared__ __align__(128) unsigned smem_u[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    unsigned* sJ = smem_u + warp * G::WARP_WORDS;
    unsigned* sI = sJ + G::J_WORDS;
    unsigned* sD = sI + G::I_WORDS;
    // per-warp mbarriers behind the regions: [0] template window + derivatives, [1] search region; phase parities
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem_u + G::WARPS * G::WARP_WORDS) + 2 * warp;
    unsigned ph_t = 0, ph_j = 0;
    if (TMA) {
        if (lane == 0) {
            mbar_init(bars, 1);
            mbar_init(bars + 1, 1);
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncwarp();
    }

    // ---- run geometry of this lane (constant for the whole kernel) ----
    int jofs[NRUN], iofs[NRUN], dofs[NRUN];
    bool rvalid[NRUN], rlast[NRUN];
#pragma unroll
    for (int s = 0; s < NRUN; s++) {
        const int q = lane + 32 * s;
        rvalid[s] = q < G::NRUNS;
        const int r = rvalid[s] ? q / G::NCB : 0, c = rvalid[s] ? q - r * G::NCB : 0;
        rlast[s] = c == G::NCB - 1;
        jofs[s] = r * (G::J_PW * 4) + c * R;
        iofs[s] = r * (G::I_PW * 4) + c * R;
        // runs that do not exist read two rows of zeros (kept behind the derivative region): their Ix = Iy = 0
        dofs[s] = rvalid[s] ? r * G::D_PW + c * R : G::D_ZERO;
    }
    if (lane < 2 * (R + 1)) sD[G::D_ZERO + (lane / (R + 1)) * G::D_PW + lane % (R + 1)] = 0;
    __syncwarp();

    constexpr float hwx = (WW - 1) * 0.5f, hwy = (WH - 1) * 0.5f;
    const float FLT_SCALE = 1.f / (1 << 20);
    const bool want_err = P.err != nullptr;
    const bool get_min_eig = (P.flags & DR3LK_GET_MIN_EIGENVALS) != 0;

    // Template window of previous-frame point pp at `level`: origin (ipx, ipy), sub-pixel position (fx, fy), in-frame flag
    struct Origin {
        int ipx, ipy;
        float fx, fy;
        bool inb;
    };
    auto template_origin = [&](const float2 pp, int level) -> Origin {
        Origin o;
        const float sc = __int_as_float((127 - level) << 23);
        o.fx = __fsub_rn(__fmul_rn(pp.x, sc), hwx);
        o.fy = __fsub_rn(__fmul_rn(pp.y, sc), hwy);
        o.ipx = __float2int_rd(o.fx); o.ipy = __float2int_rd(o.fy);
        const LevelDesc& L = P.lv[level];
        o.inb = (unsigned)(o.ipx + WW) < (unsigned)(L.w + WW) && (unsigned)(o.ipy + WH) < (unsigned)(L.h + WH);
        return o;
    };
    // issue the copies of the template window (image + derivatives) at origin o of (pair, level) into sI / sD
    auto issue_template = [&](int pair, const Origin& o, int level) {
        if (!o.inb) return;
        const LevelDesc& L = P.lv[level];
#ifdef DR3LK_CHECKED
        {   // the rectangles about to be copied lie inside the apron-carrying allocations of this (pair, level)
            const uint8_t* im = L.prev + (unsigned long long)(unsigned)pair * L.prev_stride;
            const uint8_t* lo = im - (G::PY * L.pitch_p + G::PX);
            const int xi = T::x0_I(o.ipx, L.pitch_p);
            // (flags bit 30, checked build only: pretend the allocation is one window short at the bottom -- the negative control
            //  of tests/test_gpu_checked.py, which must then see violations of kind 1 for windows at the lower border)
            const int short_rows = (P.flags & 0x40000000) ? WH : 0;
            T::check_rect(im + (o.ipy * L.pitch_p + xi), TMA ? WH + 1 : G::I_ROWS, G::I_W, L.pitch_p, xi + G::PX, lo,
                          lo + (long long)L.pitch_p * (L.h + 2 * G::PY - short_rows), 1);
            const int* de = L.deriv + (unsigned long long)(unsigned)pair * L.deriv_stride;
            const uint8_t* dlo = reinterpret_cast<const uint8_t*>(de - (G::PY * L.dpitch + G::DPX));
            const int xd = T::x0_D(o.ipx, L.dpitch);
            T::check_rect(reinterpret_cast<const uint8_t*>(de + (o.ipy * L.dpitch + xd)), TMA ? WH + 1 : G::D_ROWS, G::D_CH * 16, L.dpitch * 4, (xd + G::DPX) * 4, dlo,
                          dlo + 4ll * L.dpitch * (L.h + 2 * G::PY), 2);
            // ... and the window itself inside what is staged
            DR3LK_CHECK(o.ipx - xi >= 0 && o.ipx - xi + WW + 1 <= G::I_W && o.ipx - xd >= 0 && o.ipx - xd + WW + 1 <= G::D_CH * 4, 5, o.ipx);
        }
#endif
        if (TMA) {
            if (elect_one()) {
                mbar_expect_tx(bars, G::TMA_I_BYTES + G::TMA_D_BYTES);
                tma_box(sI, &P.tma[level].prev, T::x0_I(o.ipx, L.pitch_p) + G::PX, o.ipy + G::PY, pair, bars);
                tma_box(sD, &P.tma[level].deriv, T::x0_D(o.ipx, L.dpitch) + G::DPX, o.ipy + G::PY, pair, bars);
            }
            return;
        }
        T::stage_I(sI, L.prev + (unsigned long long)(unsigned)pair * L.prev_stride, L.pitch_p, o.ipx, o.ipy, lane);
        T::stage_D(sD, L.deriv + (unsigned long long)(unsigned)pair * L.deriv_stride, L.dpitch, o.ipx, o.ipy, lane);
    };
    // wait for the template window issued last (none is in flight when its origin was out of frame)
    auto wait_template = [&](bool issued) {
        if (TMA) {
            if (issued) { mbar_wait(bars, ph_t); ph_t ^= 1; }
        } else {
            cp_async_wait<0>();
        }
        __syncwarp();
    };
    // Work distribution: every warp reserves P.fetch_n consecutive features with one atomic (1 for small batches, where
    // every resident warp should get its own feature; up to 8 for large ones: fewer atomics on the one counter)
    int f_have = 0, f_end = 0;  // indices f_have .. f_end - 1 are already reserved for this warp
    auto fetch = [&]() -> int {
        if (f_have < f_end) return f_have++;
        int f = 0;
        if (lane == 0) f = atomicAdd(P.work_counter + (P.work_epoch & 1), P.fetch_n);
        f = __shfl_sync(0xffffffffu, f, 0);
        f_have = f + 1; f_end = f + P.fetch_n;
        return f;
    };
    auto pair_of = [&](int f) -> int { return P.uniform_n > 0 ? f / P.uniform_n : __ldg(P.pair_idx + f); };

    if (blockIdx.x == 0 && threadIdx.x == 0) P.work_counter[(P.work_epoch & 1) ^ 1] = 0;  // ready for the next launch
    // everything above is launch-local set-up; the pyramids, the points and the work counter come from the operations before
    // this kernel in the stream (Launch::pdl)
    grid_dependency_wait();
    int f = fetch();
    if (f >= P.n_total) return;
    float2 pp = sanitize_point(P.prev_pts[f]);
    int pair = pair_of(f);
    Origin org = template_origin(pp, P.max_level);  // always the origin of the template window that is in flight / staged
    issue_template(pair, org, P.max_level);
    if (!TMA) cp_async_commit();

    for (;;) {
        // look ahead: the next feature's coarsest template window is prefetched while this one finishes
        const int f_next = fetch();
        float2 pp_next = make_float2(0.f, 0.f);
        int pair_next = 0;
        if (f_next < P.n_total) { pp_next = sanitize_point(P.prev_pts[f_next]); pair_next = pair_of(f_next); }

        float2 np = make_float2(0.f, 0.f);
        if (P.flags & DR3LK_USE_INITIAL_FLOW) np = sanitize_point(P.next_pts[f]);

        int status = 1;
        float err = 0.f;
        unsigned n_iters = 0, n_templates = 0, err_pass = 0;

        for (int level = P.max_level; level >= 0; --level) {
            const LevelDesc& L = P.lv[level];
            const int w = L.w, h = L.h;
            const float sc = __int_as_float((127 - level) << 23);
            const uint8_t* imgJ = L.next + (unsigned long long)(unsigned)pair * L.next_stride;

            const int ipx = org.ipx, ipy = org.ipy;
            const float px = org.fx, py = org.fy;
            const bool inb = org.inb;
            float nx, ny;
            if (level == P.max_level) {
                if (P.flags & DR3LK_USE_INITIAL_FLOW) { nx = __fmul_rn(np.x, sc); ny = __fmul_rn(np.y, sc); }
                else { nx = __fmul_rn(pp.x, sc); ny = __fmul_rn(pp.y, sc); }
            } else {
                nx = __fmul_rn(np.x, 2.f); ny = __fmul_rn(np.y, 2.f);
            }
            np.x = nx; np.y = ny;
            nx = __fsub_rn(nx, hwx); ny = __fsub_rn(ny, hwy);

            // The staged search region as the iterations see it: window origins (vx0 .. vx0 + vxs, vy0 .. vy0 + vys) are
            // BOTH inside the staged region and inside the frame (OpenCV's bounds test), so one range test per axis
            // serves both; jb0 turns an origin into its byte offset in sJ.  Nothing staged: an empty range.
            int vx0 = 0x40000000, vy0 = 0x40000000, vxs = 0, vys = 0, jb0 = 0;
            bool j_pending = false;  // TMA: a search-region box is in flight on bars[1]
            auto stage_search = [&](int ox, int oy) {
                int rx0, ry0;
                if (TMA) {
                    T::origin_J(L.pitch_n, h, ox, oy, rx0, ry0);
                    if (elect_one()) {
                        mbar_expect_tx(bars + 1, G::TMA_J_BYTES);
                        tma_box(sJ, &P.tma[level].next, rx0 + G::PX, ry0 + G::PY, pair, bars + 1);
                    }
                    j_pending = true;
                } else {
                    T::stage_J(sJ, imgJ, L.pitch_n, h, ox, oy, lane, rx0, ry0);
                }
                vx0 = max(rx0, -WW); vxs = min(rx0 + (G::J_W - (WW + 1)), w - 1) - vx0;
                vy0 = max(ry0, -WH); vys = min(ry0 + 2 * G::MY, h - 1) - vy0;
                jb0 = -(ry0 * (G::J_PW * 4) + rx0);
#ifdef DR3LK_CHECKED
                {
                    const uint8_t* lo = imgJ - (G::PY * L.pitch_n + G::PX);
                    T::check_rect(imgJ + (ry0 * L.pitch_n + rx0), TMA ? G::J_H : G::J_ROWS, G::J_W, L.pitch_n, rx0 + G::PX, lo,
                                  lo + (long long)L.pitch_n * (h + 2 * G::PY), 4);
                    // the origin that triggered the staging is inside the range it produced
                    DR3LK_CHECK((unsigned)(ox - vx0) <= (unsigned)vxs && (unsigned)(oy - vy0) <= (unsigned)vys, 6, ox);
                }
#endif
            };
            auto wait_search = [&]() {
                if (TMA) {
                    if (j_pending) { mbar_wait(bars + 1, ph_j); ph_j ^= 1; j_pending = false; }
                } else {
                    cp_async_wait<0>();
                }
            };
            // Copy discipline: at most ONE cp.async group is in flight whenever the warp waits, and every wait is a
            // wait_group 0.  (An earlier version kept the template prefetch and the search region in flight together and
            // relied on wait_group 1 retiring them in issue order; under back-to-back levels -- zero iterations -- that
            // let a template phase start on a window that had not landed, about once in 10^5 features.)
            wait_template(inb);  // this level's template window (prefetched during the previous level) has landed
            // search region around the initial estimate: issued now, consumed after the template phase
            if (inb) {
                const int jx = __float2int_rd(nx), jy = __float2int_rd(ny);
                if ((unsigned)(jx + WW) < (unsigned)(w + WW) && (unsigned)(jy + WH) < (unsigned)(h + WH)) stage_search(jx, jy);
            }
            if (!TMA) cp_async_commit();

            int dxr[NRUN][R], dyr[NRUN][R];
            int jinit[NRUN][R];  // 2^8 - 512 * I: the dp2a addend that turns the J sample into (J - I)
            int a11 = 0, a12 = 0, a22 = 0;
            Weights q;
            if (inb) {
                // ---- template: (I, Ix, Iy) of the window into registers, Gram matrix ----
                n_templates++;
                q = make_weights(__fsub_rn(px, (float)ipx), __fsub_rn(py, (float)ipy));
                const int ox = ipx - T::x0_I(ipx, L.pitch_p), oxw = ipx - T::x0_D(ipx, L.dpitch);
#pragma unroll
                for (int s = 0; s < NRUN; s++) {
                    RunBytes<R, G::NWD, G::NEO, G::I_PW> rb;
                    rb.load(sI, iofs[s] + ox, G::I_WORDS, G::I_PW * 4 * (WH + 1));
                    const unsigned* dp = sD + dofs[s] + (rvalid[s] ? oxw : 0);
                    DR3LK_CHECK(dp >= sD && dp + G::D_PW + R + 1 <= sD + G::D_WORDS && (!rvalid[s] || dofs[s] + oxw + G::D_PW + R + 1 <= G::D_PW * (WH + 1) + 4), 7, dofs[s] + oxw);
                    int tx[R + 1], ty[R + 1], bx[R + 1], by[R + 1];
#pragma unroll
                    for (int k = 0; k <= R; k++) {
                        const int dt = (int)dp[k], db = (int)dp[G::D_PW + k];
                        tx[k] = (short)dt; ty[k] = dt >> 16;
                        bx[k] = (short)db; by[k] = db >> 16;
                    }
#pragma unroll
                    for (int k = 0; k < R; k++) {
                        // 2^8 - 512 * I with I = (S + 2^8) >> 9: clear the low 9 bits of the sum instead of shifting down and
                        // multiplying up again (two ALU-pipe instructions instead of a shift and an IMAD on the FMA pipe)
                        const int i_sum = rb.sum(k, q);
                        int ix = (tx[k] * q.w00 + tx[k + 1] * q.w01 + bx[k] * q.w10 + bx[k + 1] * q.w11 + (1 << (W_BITS - 1))) >> W_BITS;
                        int iy = (ty[k] * q.w00 + ty[k + 1] * q.w01 + by[k] * q.w10 + by[k + 1] * q.w11 + (1 << (W_BITS - 1))) >> W_BITS;
                        if (G::RAGGED && k >= WW - (G::NCB - 1) * R && rlast[s]) { ix = 0; iy = 0; }  // pixels past the window edge
                        dxr[s][k] = ix; dyr[s][k] = iy;
                        jinit[s][k] = (1 << (W_BITS - 5 - 1)) - (i_sum & ~((1 << (W_BITS - 5)) - 1));
                        a11 += ix * ix; a12 += ix * iy; a22 += iy * iy;
                    }
                }
            }
            // the search region has landed behind the template phase ...
            wait_search();
            // ... and the template regions are free again: prefetch the next template window (next finer level of this
            // feature, or the coarsest level of the next feature) behind the iterations
            __syncwarp();
            if (level > 0 || f_next < P.n_total) {  // one copy of the staging code serves both cases
                const float2 pq = level > 0 ? pp : pp_next;
                const int pair_q = level > 0 ? pair : pair_next, level_q = level > 0 ? level - 1 : P.max_level;
                org = template_origin(pq, level_q);
                issue_template(pair_q, org, level_q);
            }
            if (!TMA) cp_async_commit();

            if (!inb) {
                if (level == 0) { status = 0; err = 0.f; }
                continue;
            }
            const HiLo s11 = warp_sum_hilo(a11), s12 = warp_sum_hilo(a12), s22 = warp_sum_hilo(a22);

            const float A11 = __fmul_rn(hilo_to_float(s11.hi, s11.lo), FLT_SCALE);
            const float A12 = __fmul_rn(hilo_to_float(s12.hi, s12.lo), FLT_SCALE);
            const float A22 = __fmul_rn(hilo_to_float(s22.hi, s22.lo), FLT_SCALE);
            float D = __fsub_rn(__fmul_rn(A11, A22), __fmul_rn(A12, A12));
            const float dA = __fsub_rn(A11, A22);
            const float rad = __fadd_rn(__fmul_rn(dA, dA), __fmul_rn(__fmul_rn(4.f, A12), A12));
            const float minEig = __fdiv_rn(__fsub_rn(__fadd_rn(A22, A11), __fsqrt_rn(rad)), (float)(2 * WW * WH));
            if (want_err && get_min_eig) err = minEig;
            if (minEig < P.min_eig_thr || D < FLT_EPSILON) {
                if (level == 0) status = 0;
                continue;
            }
            // 1/D with the 2^-20 scale of the mismatch vector folded in: scaling by a power of two commutes with every
            // rounding below (no under/overflow: |b| < 2^35, 2^-23 <= D), so b1, b2 stay unscaled
            const float Ds = __fmul_rn(__fdiv_rn(1.f, D), FLT_SCALE);

            // ---- iterations (the search region landed above; only the template prefetch is in flight) ----
            float pdx = 0.f, pdy = 0.f;
            bool moved = false;
            // The realigned source bytes of the search window stay in registers across iterations (G::CACHE_J): after
            // the first iteration of a level most updates are sub-pixel, the integer window origin does not move, and
            // only the weights change -- the 12 LDS + 16 PRMT of a window load are then skipped.
            RunBytes<R, G::NWD, G::NEO, G::J_PW> rbj[NRUN];
            int jloaded = 0x7fffffff;  // sJ byte offset rbj was loaded from (none yet / invalid after a re-stage)
            for (int j = 0; j < P.max_count; ++j) {
                const int inx = __float2int_rd(nx), iny = __float2int_rd(ny);
                if ((unsigned)(inx - vx0) > (unsigned)vxs || (unsigned)(iny - vy0) > (unsigned)vys) {
                    // left the staged search region -- or the frame
                    if ((unsigned)(inx + WW) >= (unsigned)(w + WW) || (unsigned)(iny + WH) >= (unsigned)(h + WH)) {
                        if (level == 0) status = 0;
                        break;
                    }
                    __syncwarp();
                    stage_search(inx, iny);
                    if (!TMA) cp_async_commit();
                    wait_search();
                    __syncwarp();
                    jloaded = 0x7fffffff;
                }
                q = make_weights(__fsub_rn(nx, (float)inx), __fsub_rn(ny, (float)iny));
                const int jbase = iny * (G::J_PW * 4) + inx + jb0;
                if (G::CACHE_J && jbase != jloaded) {
#pragma unroll
                    for (int s = 0; s < NRUN; s++) rbj[s].load(sJ, jbase + jofs[s], G::J_WORDS, G::J_PW * 4 * G::J_H);
                    jloaded = jbase;
                }
                int b1 = 0, b2 = 0;
#pragma unroll
                for (int s = 0; s < NRUN; s++) {
                    RunBytes<R, G::NWD, G::NEO, G::J_PW> rbl;
                    if (!G::CACHE_J) rbl.load(sJ, jbase + jofs[s], G::J_WORDS, G::J_PW * 4 * G::J_H);
                    const RunBytes<R, G::NWD, G::NEO, G::J_PW>& rb = G::CACHE_J ? rbj[s] : rbl;
#pragma unroll
                    for (int k = 0; k < R; k++) {
                        const int diff = rb.sample(k, q, jinit[s][k]);  // J - I
                        b1 += diff * dxr[s][k];
                        b2 += diff * dyr[s][k];
                    }
                }
                const HiLo sb1 = warp_sum_hilo(b1), sb2 = warp_sum_hilo(b2);
                n_iters++;
                // exact integer sums, one rounding to fp32 (the 2^-20 scale lives in Ds)
                const float fb1 = hilo_to_float(sb1.hi, sb1.lo);
                const float fb2 = hilo_to_float(sb2.hi, sb2.lo);
                const float ddx = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb2), __fmul_rn(A22, fb1)), Ds);
                const float ddy = __fmul_rn(__fsub_rn(__fmul_rn(A12, fb1), __fmul_rn(A11, fb2)), Ds);
                nx = __fadd_rn(nx, ddx); ny = __fadd_rn(ny, ddy);
                moved = true;
                // |delta|^2 <= eps^2 in double like OpenCV; the fp32 estimate decides unless it is within 1e-6 of the threshold
                const float t = __fmaf_rn(ddx, ddx, __fmul_rn(ddy, ddy));
                bool conv = t < P.eps2_lo;
                if (!conv && !(t > P.eps2_hi))
                    conv = __dadd_rn(__dmul_rn((double)ddx, (double)ddx), __dmul_rn((double)ddy, (double)ddy)) <= P.eps2;
                if (conv) break;
                // fabs(float) < 0.01 (double)  <=>  fabs(float) <= 0.01f, because 0.01f is the largest float below 0.01
                if (j > 0 && fabsf(__fadd_rn(ddx, pdx)) <= 0.01f && fabsf(__fadd_rn(ddy, pdy)) <= 0.01f) {
                    np.x = __fsub_rn(__fadd_rn(nx, hwx), __fmul_rn(ddx, 0.5f));
                    np.y = __fsub_rn(__fadd_rn(ny, hwy), __fmul_rn(ddy, 0.5f));
                    moved = false;
                    break;
                }
                pdx = ddx; pdy = ddy;
            }
            if (moved) { np.x = __fadd_rn(nx, hwx); np.y = __fadd_rn(ny, hwy); }

            if (status && want_err && level == 0 && !get_min_eig) {
                const float qx = __fsub_rn(np.x, hwx), qy = __fsub_rn(np.y, hwy);
                const int iqx = __float2int_rd(qx), iqy = __float2int_rd(qy);
                if ((unsigned)(iqx + WW) >= (unsigned)(w + WW) || (unsigned)(iqy + WH) >= (unsigned)(h + WH)) {
                    status = 0;
                    continue;
                }
                q = make_weights(__fsub_rn(qx, (float)iqx), __fsub_rn(qy, (float)iqy));
                if ((unsigned)(iqx - vx0) > (unsigned)vxs || (unsigned)(iqy - vy0) > (unsigned)vys) {
                    __syncwarp();
                    stage_search(iqx, iqy);
                    if (!TMA) cp_async_commit();
                    wait_search();
                    __syncwarp();
                    jloaded = 0x7fffffff;
                }
                const int jbase = iqy * (G::J_PW * 4) + iqx + jb0;
                if (G::CACHE_J && jbase != jloaded) {
#pragma unroll
                    for (int s = 0; s < NRUN; s++) rbj[s].load(sJ, jbase + jofs[s], G::J_WORDS, G::J_PW * 4 * G::J_H);
                }
                int es = 0;
#pragma unroll
                for (int s = 0; s < NRUN; s++) {
                    RunBytes<R, G::NWD, G::NEO, G::J_PW> rbl;
                    if (!G::CACHE_J) rbl.load(sJ, jbase + jofs[s], G::J_WORDS, G::J_PW * 4 * G::J_H);
                    const RunBytes<R, G::NWD, G::NEO, G::J_PW>& rb = G::CACHE_J ? rbj[s] : rbl;
                    int e = 0;
#pragma unroll
                    for (int k = 0; k < R; k++) {
                        const int d = abs(rb.sample(k, q, jinit[s][k]));
                        if (G::RAGGED && k >= WW - (G::NCB - 1) * R) e += rlast[s] ? 0 : d; else e += d;
                    }
                    es += rvalid[s] ? e : 0;
                }
                es = __reduce_add_sync(0xffffffffu, es);
                err_pass = 1;
                err = __fdiv_rn(__fmul_rn((float)es, 1.f), (float)(32 * WW * WH));
            }
        }

        DR3LK_CHECK((unsigned)f < (unsigned)P.n_total, 8, f);
        if (lane == 0) {
            P.next_out[f] = np;
            P.status[f] = (uint8_t)status;
            if (want_err) P.err[f] = err;
            if (P.stats) P.stats[f] = (n_iters & 0xffffu) | ((n_templates & 0xffu) << 16) | (err_pass << 24);
        }
        if (f_next >= P.n_total) break;
        f = f_next; pp = pp_next; pair = pair_next;
    }
    if (!TMA) cp_async_wait<0>();
}

constexpr int kMaxDevices = 64;

template <typename G, bool TMA>
bool launch_one(Launch& L, const LKParams& p)
{
    const size_t smem = (size_t)G::WARPS * G::WARP_WORDS * sizeof(unsigned) + (TMA ? G::WARPS * 16 : 0);
    // shared-memory opt-in, SM count and occupancy are per (kernel, device): looked up once, not on every call of the latency path
    struct PerDevice { int sms = 0, per_sm = 0; };
    static PerDevice cache[kMaxDevices];
    static std::mutex cache_mutex;
    int dev = 0, sms = 0, per_sm = 0;
    cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> g(cache_mutex);
        PerDevice local;
        PerDevice& c = (dev >= 0 && dev < kMaxDevices) ? cache[dev] : local;
        if (c.sms == 0) {
            L.err = cudaFuncSetAttribute(lk_fast_kernel<G, TMA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
            if (L.err != cudaSuccess) return false;
            int n = 0, o = 0;
            cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
            L.err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&o, lk_fast_kernel<G, TMA>, G::WARPS * 32, smem);
            if (L.err != cudaSuccess) return false;
            c.sms = n; c.per_sm = o;
        }
        sms = c.sms; per_sm = c.per_sm;
    }
    const int want = (p.n_total + G::WARPS - 1) / G::WARPS;
    const int blocks = std::min(want, std::max(1, per_sm) * sms);
    LKParams q = p;
    static const int fetch_max = getenv("DR3LK_FETCH_MAX") ? atoi(getenv("DR3LK_FETCH_MAX")) : 8;
    // reserve several features per atomic only when every warp has at least 64 features to work through
    q.fetch_n = std::max(1, std::min(fetch_max, (int)(p.n_total / (64LL * blocks * G::WARPS))));
    L.err = launch_kernel(L, lk_fast_kernel<G, TMA>, dim3(blocks), dim3(G::WARPS * 32), smem, q);
    if (L.err == cudaSuccess) L.err = cudaGetLastError();
    L.launches++;
    return L.err == cudaSuccess;  // true: the persistent kernel ran and consumed its work counter
}

template <typename G>
bool launch_geo(Launch& L, const LKParams& p)
{
    for (int l = 0; l <= p.max_level; l++)
        if (p.lv[l].w < G::MIN_W || p.lv[l].h < G::MIN_H) return false;  // tiny level: lk_generic handles it
    return p.use_tma ? launch_one<G, true>(L, p) : launch_one<G, false>(L, p);
}

template <typename G>
LkFastBoxes boxes_of()
{
    return LkFastBoxes{G::I_W, G::WH + 1, G::D_CH * 4, G::WH + 1, G::J_W, G::J_H};
}

}  // namespace

bool lk_fast_supported(int win_w, int win_h)
{
    return (win_w == 21 && win_h == 21) || (win_w == 31 && win_h == 31) || (win_w == 30 && win_h == 30);
}

bool launch_lk_fast(Launch& L, const LKParams& p)
{
    if (!lk_fast_supported(p.win_w, p.win_h) || !p.fast_ok) return false;
    if (L.err != cudaSuccess || p.n_total <= 0) return false;  // nothing launched; the generic launcher is a no-op here too
    if (p.win_w == 21) return launch_geo<Geo<21, 21, 7>>(L, p);
    if (p.win_w == 31) return launch_geo<Geo<31, 31, 8>>(L, p);
    return launch_geo<Geo<30, 30, 10>>(L, p);
}

// checked build: violations / kind / detail / checks executed since the last read (and resets them); false in the default build
bool lk_fast_check_read(unsigned long long out[4]) { return check_read_tu(out); }

bool lk_fast_boxes(int win_w, int win_h, LkFastBoxes* b)
{
    if (!lk_fast_supported(win_w, win_h)) return false;
    *b = win_w == 21 ? boxes_of<Geo<21, 21, 7>>() : (win_w == 31 ? boxes_of<Geo<31, 31, 8>>() : boxes_of<Geo<30, 30, 10>>());
    return true;
}

}  // namespace dr3lk
