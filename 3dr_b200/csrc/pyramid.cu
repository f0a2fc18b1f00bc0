// Pyramid kernels of the dr3lk path (sm_100a).
//
//  * pyr_level_kernel  -- one launch per pyramid level for a batch of images: stages a 160 x 36 uint8 source tile
//    (16-byte vectorised row loads, halo, REFLECT_101 by index reflection) in shared memory once and produces from it
//      - the next Gaussian level: OpenCV pyrDown, separable [1 4 6 4 1], (sum+128)>>8   (SURVEY.md A.2), and
//      - the Scharr derivative of the source level as packed int16x2 (Ix, Iy)            (SURVEY.md A.3),
//    i.e. the derivative is fused into the pyramid build: the level is read from HBM exactly once.
//  * box_half_kernel   -- utils::reduce_to_half of the reference (src/utils.cpp:382-419) including its
//    rounding modes and the odd-width pointer walk.
#include "dr3lk_internal.cuh"

namespace dr3lk {

namespace {

constexpr int TOX = 64;            // output (down-sampled) tile
constexpr int TOY_BATCH = 16;       // output rows per tile: 16 for batches, 8 on the few-frame latency path (twice the blocks,
constexpr int TOY_FEW = 8;          // about half the dependent work per block)
constexpr int HX = 16;             // x halo of the source tile: 16 keeps every tile row 16-byte aligned (2 are needed)
constexpr int SW = 2 * TOX + 2 * HX;   // 160 source bytes per tile row = 10 chunks of 16
// source rows per tile: 2 * TOY + 4 (halo 2) = 36 / 20
constexpr int SPITCH = SW + 16;        // 176
constexpr int PYR_THREADS = 256;

__device__ __forceinline__ unsigned byte_of(unsigned w, int i) { return (w >> (8 * i)) & 0xffu; }

// One pyramid level for a batch of images.  The source tile is staged once with 16-byte row-coalesced loads (interior
// tiles of 16-B aligned images) or byte-wise with REFLECT_101 index reflection (edge tiles / unaligned sources); the
// next Gaussian level (pyrDown) and the Scharr derivative of the source level are both produced from that tile.
// blockIdx.z walks two image sets back to back: n_a "A" images (the previous frames: derivative + optional down-sample)
// followed by the "B" images (the next frames: down-sample only), so one launch per level serves both pyramids.
struct PyrSet {
    const uint8_t* src;
    uint8_t* dst;
    unsigned src_stride, dst_stride;  // bytes between images
};

template <bool DOWN, int TOY>
__global__ void __launch_bounds__(PYR_THREADS)
pyr_level_kernel(PyrSet A, PyrSet B, int n_a, int w, int h, int src_pitch, int src_aligned, int dst_pitch, int* __restrict__ deriv,
                 int dpitch, unsigned deriv_stride, int apron_x, int apron_y, int src_ax, int src_ay)
{
    constexpr int SH = 2 * TOY + 4;
    __shared__ __align__(16) uint8_t tile[SH][SPITCH];
    __shared__ __align__(8) short hrow[DOWN ? SH : 1][DOWN ? TOX : 4];

    grid_dependency_wait();  // the source level is the output of the kernel before this one (Launch::pdl)
    const int tid = threadIdx.x;
    const bool set_b = (int)blockIdx.z >= n_a;
    const int img = set_b ? blockIdx.z - n_a : blockIdx.z;
    const PyrSet& S = set_b ? B : A;
    const bool DERIV = !set_b && deriv != nullptr;
    const int xs = blockIdx.x * (2 * TOX);      // first source column this tile produces derivatives for
    const int ys = blockIdx.y * (2 * TOY);
    const int x0 = xs - HX;                     // source coordinate of tile[0][0]
    const int y0 = ys - 2;
    const uint8_t* s = S.src + (unsigned long long)(unsigned)img * S.src_stride;

    // columns xs-2 .. xs+2*TOX+1 are needed; the vector path also requires the 16-byte chunks to stay inside the row
    const bool xin = src_aligned && blockIdx.x >= 1 && xs + 2 * TOX + 2 <= w && x0 + SW <= src_pitch;
    if (src_aligned && src_ax >= HX && src_ay >= 2 && w >= 3 && h >= 3) {  // (the two apron pixels a tile uses are single reflections)
        // the source level carries its REFLECT_101 apron: every row / column the tile needs is in memory, so every tile is
        // plain 16-byte row copies (chunks / rows beyond the apron are never read back from the tile)
        for (int idx = tid; idx < SH * (SW / 16); idx += PYR_THREADS) {
            const int ty = idx / (SW / 16), ch = idx - ty * (SW / 16);
            const int sx = x0 + ch * 16, sy = min(y0 + ty, h + src_ay - 1);
            if (sx + 16 <= src_pitch - src_ax)
                *reinterpret_cast<uint4*>(&tile[ty][ch * 16]) = __ldg(reinterpret_cast<const uint4*>(s + (long long)sy * src_pitch + sx));
        }
    } else if (xin) {
        for (int idx = tid; idx < SH * (SW / 16); idx += PYR_THREADS) {
            const int ty = idx / (SW / 16), ch = idx - ty * (SW / 16);
            const int sy = reflect101(y0 + ty, h);
            const uint4 v = __ldg(reinterpret_cast<const uint4*>(s + (long long)sy * src_pitch + x0 + ch * 16));
            *reinterpret_cast<uint4*>(&tile[ty][ch * 16]) = v;
        }
    } else {
        for (int idx = tid; idx < SH * (2 * TOX + 4); idx += PYR_THREADS) {
            const int ty = idx / (2 * TOX + 4), tx = HX - 2 + (idx - ty * (2 * TOX + 4));
            const int sx = reflect101(x0 + tx, w), sy = reflect101(y0 + ty, h);
            tile[ty][tx] = __ldg(s + (long long)sy * src_pitch + sx);
        }
    }
    __syncthreads();

    if (DERIV) {
        // 4 consecutive pixels per thread: source (xs + 4g .. +3, ys + ly) <-> tile[ly + 2][HX + 4g ..]
        const int g = tid & 31;
        int* d = deriv + (unsigned long long)(unsigned)img * deriv_stride;
        for (int ly = tid >> 5; ly < 2 * TOY; ly += PYR_THREADS / 32) {
            const int sx = xs + 4 * g, sy = ys + ly;
            if (sx < w && sy < h) {
                const unsigned* r0 = reinterpret_cast<const unsigned*>(&tile[ly + 1][HX - 4 + 4 * g]);
                const unsigned* r1 = reinterpret_cast<const unsigned*>(&tile[ly + 2][HX - 4 + 4 * g]);
                const unsigned* r2 = reinterpret_cast<const unsigned*>(&tile[ly + 3][HX - 4 + 4 * g]);
                // Bytes 3..8 of the 12 loaded per row are columns sx-1 .. sx+4.  Two columns per register as 16-bit lanes:
                // P[0] = columns (1,3), P[1] = (4,6), P[2] = (5,7), P[3] = (8,10) -- odd / even bytes of the three words.
                // All lane values stay in [0, 65535], so plain 32-bit adds / small multiplies act lane-wise.
                unsigned T0[4], T1[4];
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int wi = (i + 1) >> 1;                         // word 0, 1, 1, 2
                    const unsigned sel = (i == 0 || i == 2) ? 0x4341u : 0x4240u;  // odd bytes (b1, b3) / even bytes (b0, b2)
                    const unsigned a = __byte_perm(r0[wi], 0, sel), b = __byte_perm(r1[wi], 0, sel), c = __byte_perm(r2[wi], 0, sel);
                    T0[i] = (a + c) * 3u + b * 10u;          // vertical smooth, <= 4080 per lane
                    T1[i] = c + 0x08000800u - a;             // vertical difference + 2048 per lane
                }
                // horizontal pass; the lanes of ix / iy carry +32768 (removed by the final xor): no borrows between lanes
                const unsigned ixe = T0[2] + 0x80008000u - __byte_perm(T0[0], T0[2], 0x5432);   // columns 4, 6: (5,7) - (3,5)
                const unsigned ixo = __byte_perm(T0[1], T0[3], 0x5432) + 0x80008000u - T0[1];   // columns 5, 7: (6,8) - (4,6)
                const unsigned iye = (__byte_perm(T1[0], T1[2], 0x5432) + T1[2]) * 3u + T1[1] * 10u;
                const unsigned iyo = (T1[1] + __byte_perm(T1[1], T1[3], 0x5432)) * 3u + T1[2] * 10u;
                int o[4];
                o[0] = (int)(__byte_perm(ixe, iye, 0x5410) ^ 0x80008000u);
                o[1] = (int)(__byte_perm(ixo, iyo, 0x5410) ^ 0x80008000u);
                o[2] = (int)(__byte_perm(ixe, iye, 0x7632) ^ 0x80008000u);
                o[3] = (int)(__byte_perm(ixo, iyo, 0x7632) ^ 0x80008000u);
                int* out = d + (long long)sy * dpitch + sx;
                DR3LK_CHECK(sx >= 0 && sy >= 0 && sx < dpitch, 20, (long long)sy * dpitch + sx);  // inside the level: the apron is never written
                DR3LK_CHECK_COUNT();
                if (sx + 3 < w) {
                    *reinterpret_cast<int4*>(out) = make_int4(o[0], o[1], o[2], o[3]);
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (sx + j < w) out[j] = o[j];
                }
            }
        }
    }

    if (DOWN) {
        // horizontal [1 4 6 4 1]: 2 outputs per thread, output ox reads tile columns HX - 2 + 2*ox .. + 4
        const int m = tid & 31;
        for (int ty = tid >> 5; ty < SH; ty += PYR_THREADS / 32) {
            const unsigned* r = reinterpret_cast<const unsigned*>(&tile[ty][HX - 4 + 4 * m]);
            const unsigned w0 = r[0], w1 = r[1], w2 = r[2];
            // two outputs per register on 16-bit lanes (values <= 16 * 255): bytes b2..b8 of the three words, even bytes
            // E = (b0, b2), odd bytes O = (b1, b3) of a word; (hi of one, lo of the next) pairs come from one PRMT
            const unsigned e0 = __byte_perm(w0, 0, 0x4240), o0 = __byte_perm(w0, 0, 0x4341), e1 = __byte_perm(w1, 0, 0x4240),
                           o1 = __byte_perm(w1, 0, 0x4341), e2 = __byte_perm(w2, 0, 0x4240);
            const unsigned p24 = __byte_perm(e0, e1, 0x5432), p35 = __byte_perm(o0, o1, 0x5432), p68 = __byte_perm(e1, e2, 0x5432);
            // (h0, h1) = (b2, b4) + 4 (b3, b5) + 6 (b4, b6) + 4 (b5, b7) + (b6, b8)
            *reinterpret_cast<unsigned*>(&hrow[ty][2 * m]) = (p24 + p68) + (p35 + o1) * 4u + e1 * 6u;
        }
        __syncthreads();
        const int dw = (w + 1) >> 1, dh = (h + 1) >> 1;
        uint8_t* o = S.dst + (unsigned long long)(unsigned)img * S.dst_stride;
        const int q = tid & 15, oy = tid >> 4;   // 4 outputs per thread, 16 threads per row, 16 rows
        const int gx = blockIdx.x * TOX + 4 * q, gy = blockIdx.y * TOY + oy;
        if (oy < TOY && gx < dw && gy < dh) {
            // vertical [1 4 6 4 1] on the same 16-bit lanes: 128 + 16 * 4080 < 65536, so (sum + 128) >> 8 is the high byte
            // of each lane and one PRMT packs the four output pixels
            unsigned v01 = 0x00800080u, v23 = 0x00800080u;
#pragma unroll
            for (int k = 0; k < 5; k++) {
                const uint2 hv = *reinterpret_cast<const uint2*>(&hrow[2 * oy + k][4 * q]);
                const unsigned c = (k == 0 || k == 4) ? 1u : (k == 2 ? 6u : 4u);
                v01 += c * hv.x; v23 += c * hv.y;
            }
            const unsigned word = __byte_perm(v01, v23, 0x7531);
            unsigned v[4];
#pragma unroll
            for (int j = 0; j < 4; j++) v[j] = ((word >> (8 * j)) & 0xffu) << 8;  // kept in the old "v >> 8" form for the stores below
            // the row itself plus its REFLECT_101 images in the apron rows (row g mirrors to -g and to 2(dh-1)-g)
            int ys[3] = {gy, gy, gy};
            int ny = 1;
            if (apron_y > 0) {
                if (gy >= 1 && gy <= apron_y) ys[ny++] = -gy;
                const int m = 2 * (dh - 1) - gy;
                if (gy <= dh - 2 && m <= dh - 1 + apron_y) ys[ny++] = m;
            }
#pragma unroll
            for (int i = 0; i < 3; i++) {
                if (i >= ny) break;
                uint8_t* row = o + (long long)ys[i] * dst_pitch;
                uint8_t* out = row + gx;
                // rows -apron_y .. dh - 1 + apron_y, columns -apron_x .. dst_pitch - apron_x - 1 belong to this image
                DR3LK_CHECK(ys[i] >= -apron_y && ys[i] <= dh - 1 + apron_y && gx >= 0 && gx < dw && (apron_x > 0 || apron_y == 0), 21, ys[i]);
                DR3LK_CHECK_COUNT();
                if (gx + 3 < dw && (dst_pitch & 3) == 0) {
                    *reinterpret_cast<unsigned*>(out) = word;
                } else {
#pragma unroll
                    for (int j = 0; j < 4; j++)
                        if (gx + j < dw) out[j] = (uint8_t)(v[j] >> 8);
                }
                if (apron_x > 0 && (gx <= apron_x || gx + 3 >= dw - 1 - apron_x)) {
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const int x = gx + j, mx = 2 * (dw - 1) - x;
                        if (x >= dw) continue;
                        if (x >= 1 && x <= apron_x) row[-x] = (uint8_t)(v[j] >> 8);
                        if (x <= dw - 2 && mx <= dw - 1 + apron_x) {
                            DR3LK_CHECK(mx >= dw && mx < dst_pitch - apron_x, 22, mx);  // the mirrored column stays inside the row
                            row[mx] = (uint8_t)(v[j] >> 8);
                        }
                    }
                }
            }
        }
    }
}

// Level 0 with its REFLECT_101 apron: one 16-byte destination chunk per thread.  Interior chunks read the source row
// through aligned 32-bit words realigned with funnel shifts (the caller's rows may have any alignment, e.g. continuous
// 1241-wide images).
constexpr int PAD_TX = 32, PAD_TY = 8;  // block = 32 chunk columns x 8 rows; PAD_RPT rows per thread (template parameter)

// 16 destination bytes at column x of one row (x is a multiple of 16, -ax <= x): dst[x + i] = srow[reflect101(x + i)].
// A chunk that lies completely inside the row, or completely in the left / right apron, is 16 CONSECUTIVE source
// bytes, forwards or backwards: one branch-free path (aligned 32-bit loads, funnel shifts, byte reversal) serves all
// three, so warps do not diverge.  Only the chunk that straddles the right image edge gathers byte by byte.
__device__ __forceinline__ uint4 pad_chunk(const uint8_t* __restrict__ srow, int x, int w)
{
    // first source byte: x (interior), -(x + 15) (left apron, reversed), 2(w-1) - (x + 15) (right apron, reversed)
    const bool left = x + 15 < 0 && -x <= w - 1, right = x >= w && 2 * (w - 1) - (x + 15) >= 0;
    if (left || right || (x >= 0 && x + 16 <= w)) {
        const int s0 = left ? -(x + 15) : (right ? 2 * (w - 1) - (x + 15) : x);
        const uint8_t* p = srow + s0;
        const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(p) & 3);
        const unsigned* pw = reinterpret_cast<const unsigned*>(p - mis);
        const unsigned sh = mis * 8;
        // the fifth word is only touched when the chunk is misaligned: it ends inside the word holding byte s0 + 15
        const unsigned w0 = __ldg(pw), w1 = __ldg(pw + 1), w2 = __ldg(pw + 2), w3 = __ldg(pw + 3), w4 = mis ? __ldg(pw + 4) : 0u;
        const unsigned f0 = __funnelshift_r(w0, w1, sh), f1 = __funnelshift_r(w1, w2, sh), f2 = __funnelshift_r(w2, w3, sh),
                       f3 = __funnelshift_r(w3, w4, sh);
        if (!(left || right)) return make_uint4(f0, f1, f2, f3);
        return make_uint4(__byte_perm(f3, 0, 0x0123), __byte_perm(f2, 0, 0x0123), __byte_perm(f1, 0, 0x0123), __byte_perm(f0, 0, 0x0123));
    }
    unsigned o[4] = {0, 0, 0, 0};
#pragma unroll
    for (int i = 0; i < 16; i++) {
        o[i >> 2] |= (unsigned)__ldg(srow + reflect101(x + i, w)) << (8 * (i & 3));
    }
    return make_uint4(o[0], o[1], o[2], o[3]);
}

// PAD_RPT = 4 for batches (loads of four rows in flight per thread: the kernel is HBM-bound there), 1 for the few-frame
// latency path (four times as many blocks, a quarter of the per-thread chain).
template <int PAD_RPT>
__global__ void __launch_bounds__(PAD_TX * PAD_TY)
pad_level0_kernel(const uint8_t* __restrict__ src_a, const uint8_t* __restrict__ src_b, long long src_pitch, long long src_stride,
                  uint8_t* __restrict__ dst_a, uint8_t* __restrict__ dst_b, int dst_pitch, long long dst_stride, int w, int h, int ax, int ay,
                  int n_a)
{
    const int c = blockIdx.x * PAD_TX + threadIdx.x;
    if (c * 16 >= dst_pitch) return;
    const int x = c * 16 - ax;
    const bool set_b = (int)blockIdx.z >= n_a;
    const int img = set_b ? blockIdx.z - n_a : blockIdx.z;
    const uint8_t* src = (set_b ? src_b : src_a) + (long long)img * src_stride;
    uint8_t* dst = (set_b ? dst_b : dst_a) + (long long)img * dst_stride + x;
    const int r0 = blockIdx.y * (PAD_TY * PAD_RPT) + threadIdx.y;  // row index in the apron-carrying image
    uint4 v[PAD_RPT];
#pragma unroll
    for (int k = 0; k < PAD_RPT; k++) {
        const int r = r0 + k * PAD_TY;
        if (r < h + 2 * ay) v[k] = pad_chunk(src + (long long)reflect101(r - ay, h) * src_pitch, x, w);
    }
#pragma unroll
    for (int k = 0; k < PAD_RPT; k++) {
        const int r = r0 + k * PAD_TY;
        if (r < h + 2 * ay) {
            DR3LK_CHECK(x >= -ax && x + 16 <= dst_pitch - ax && r >= 0, 23, x);  // the chunk lies inside row r of the apron-carrying image
            DR3LK_CHECK_COUNT();
            *reinterpret_cast<uint4*>(dst + (long long)(r - ay) * dst_pitch) = v[k];
        }
    }
}

// 16 consecutive bytes at any alignment through aligned 32-bit loads and funnel shifts (the 5th word is only touched
// when the address is misaligned: it then holds byte 15, so every load stays inside words that hold requested bytes)
__device__ __forceinline__ void load16(const uint8_t* __restrict__ p, unsigned (&o)[4])
{
    const unsigned mis = (unsigned)(reinterpret_cast<uintptr_t>(p) & 3);
    const unsigned* pw = reinterpret_cast<const unsigned*>(p - mis);
    const unsigned sh = mis * 8;
    const unsigned w0 = __ldg(pw), w1 = __ldg(pw + 1), w2 = __ldg(pw + 2), w3 = __ldg(pw + 3), w4 = mis ? __ldg(pw + 4) : 0u;
    o[0] = __funnelshift_r(w0, w1, sh); o[1] = __funnelshift_r(w1, w2, sh); o[2] = __funnelshift_r(w2, w3, sh); o[3] = __funnelshift_r(w3, w4, sh);
}

// utils::reduce_to_half (src/utils.cpp:382-419).  The scalar walk advances `top` by 2 per pixel and by `stride` per row
// (lines 410-417), so output row i starts at flat offset i*(2*out_w + stride): for even widths the plain 2x2 box, for odd
// widths the one-byte-per-row shear of the reference.  Eight outputs per thread: 16 + 16 source bytes at any alignment,
// two outputs per register on 16-bit lanes (even / odd source bytes), both rounding modes; the ragged end of a row
// falls back to one output at a time.
__global__ void __launch_bounds__(128)
box_half_kernel(const uint8_t* __restrict__ src, int out_w, int out_h, long long row_stride, long long src_img_stride,
                uint8_t* __restrict__ dst, long long dst_img_stride, int sse2_rounding)
{
    const int ngroups = (out_w + 7) / 8;                       // groups of 8 outputs per row
    const int idx = blockIdx.x * blockDim.x + threadIdx.x;     // (row, group) flattened: no idle threads at row ends
    const int i = idx / ngroups, g = idx - i * ngroups;
    if (i >= out_h) return;
    const uint8_t* s = src + (long long)blockIdx.y * src_img_stride + (long long)i * (2LL * out_w + row_stride) + 16 * g;
    uint8_t* d = dst + (long long)blockIdx.y * dst_img_stride + (long long)i * out_w + 8 * g;
    if (8 * g + 8 <= out_w) {
        unsigned t[4], b[4], o[4];
        load16(s, t);
        load16(s + row_stride, b);
#pragma unroll
        for (int k = 0; k < 4; k++) {  // word k: source bytes 4k .. 4k+3 of both rows -> outputs 2k, 2k+1 (lanes)
            const unsigned te = t[k] & 0x00ff00ffu, to = (t[k] >> 8) & 0x00ff00ffu, be = b[k] & 0x00ff00ffu, bo = (b[k] >> 8) & 0x00ff00ffu;
            if (sse2_rounding) {  // _mm_avg_epu8(top, bottom) then the average of the even / odd bytes, both rounding up
                const unsigned v0 = ((te + be + 0x00010001u) >> 1) & 0x00ff00ffu, v1 = ((to + bo + 0x00010001u) >> 1) & 0x00ff00ffu;
                o[k] = ((v0 + v1 + 0x00010001u) >> 1) & 0x00ff00ffu;
            } else {
                o[k] = ((te + to + be + bo) >> 2) & 0x00ff00ffu;
            }
        }
        const unsigned w0 = __byte_perm(o[0], o[1], 0x6420), w1 = __byte_perm(o[2], o[3], 0x6420);
        const unsigned al = (unsigned)(reinterpret_cast<uintptr_t>(d) & 3);
        if (al == 0) {
            reinterpret_cast<unsigned*>(d)[0] = w0; reinterpret_cast<unsigned*>(d)[1] = w1;
        } else if (al == 2) {
            reinterpret_cast<unsigned short*>(d)[0] = (unsigned short)w0;
            *reinterpret_cast<unsigned*>(d + 2) = __funnelshift_r(w0, w1, 16);
            reinterpret_cast<unsigned short*>(d)[3] = (unsigned short)(w1 >> 16);
        } else {
#pragma unroll
            for (int k = 0; k < 4; k++) { d[k] = (uint8_t)(w0 >> (8 * k)); d[4 + k] = (uint8_t)(w1 >> (8 * k)); }
        }
    } else {
        for (int k = 0; 8 * g + k < out_w; k++) {
            const unsigned a = __ldg(s + 2 * k), bb = __ldg(s + 2 * k + 1), c = __ldg(s + 2 * k + row_stride), dd = __ldg(s + 2 * k + row_stride + 1);
            unsigned v;
            if (sse2_rounding) {
                const unsigned v0 = (a + c + 1u) >> 1, v1 = (bb + dd + 1u) >> 1;
                v = (v0 + v1 + 1u) >> 1;
            } else {
                v = (a + bb + c + dd) >> 2;
            }
            d[k] = (uint8_t)v;
        }
    }
}

}  // namespace

void launch_pyr_level(Launch& L, const PyrLevelArgs& a)
{
    if (L.err != cudaSuccess || a.n_prev + a.n_next <= 0) return;
    // few images: smaller tiles, so that the stage is spread over more SMs and each block's dependent chain is shorter
    const bool few = a.n_prev + a.n_next <= 8;
    const int toy = few ? TOY_FEW : TOY_BATCH;
    dim3 grid((a.w + 2 * TOX - 1) / (2 * TOX), (a.h + 2 * toy - 1) / (2 * toy), a.n_prev + a.n_next);
    auto al = [](const void* p, size_t x, size_t y) { return ((reinterpret_cast<uintptr_t>(p) | x | y) & 15) == 0; };
    const int aligned = al(a.prev_src, a.src_pitch, a.prev_src_stride) && (a.n_next == 0 || al(a.next_src, a.src_pitch, a.next_src_stride));
    PyrSet A{a.prev_src, a.prev_dst, a.prev_src_stride, a.prev_dst_stride};
    PyrSet B{a.next_src, a.next_dst, a.next_src_stride, a.next_dst_stride};
    auto go = [&](auto kernel, int ax, int ay) {
        L.err = launch_kernel(L, kernel, grid, dim3(PYR_THREADS), 0, A, B, a.n_prev, a.w, a.h, a.src_pitch, aligned, a.dst_pitch, a.deriv, a.dpitch,
                              a.deriv_stride, ax, ay, a.src_apron_x, a.src_apron_y);
    };
    if (a.down) {
        if (few) go(pyr_level_kernel<true, TOY_FEW>, a.dst_apron_x, a.dst_apron_y);
        else go(pyr_level_kernel<true, TOY_BATCH>, a.dst_apron_x, a.dst_apron_y);
    } else {
        if (few) go(pyr_level_kernel<false, TOY_FEW>, 0, 0);
        else go(pyr_level_kernel<false, TOY_BATCH>, 0, 0);
    }
    if (L.err == cudaSuccess) L.err = cudaGetLastError();
    L.launches++;
}

void launch_pad_level0(Launch& L, const uint8_t* src_a, const uint8_t* src_b, size_t src_pitch, size_t src_stride, uint8_t* dst_a,
                       uint8_t* dst_b, int dst_pitch, size_t dst_stride, int w, int h, int ax, int ay, int n_a, int n_b)
{
    if (L.err != cudaSuccess || n_a + n_b <= 0) return;
    const int rpt = (n_a + n_b <= 8) ? 1 : 4;
    dim3 grid((dst_pitch / 16 + PAD_TX - 1) / PAD_TX, (h + 2 * ay + PAD_TY * rpt - 1) / (PAD_TY * rpt), n_a + n_b);
    if (rpt == 1)
        pad_level0_kernel<1><<<grid, dim3(PAD_TX, PAD_TY), 0, L.stream>>>(src_a, src_b, (long long)src_pitch, (long long)src_stride, dst_a, dst_b,
                                                                        dst_pitch, (long long)dst_stride, w, h, ax, ay, n_a);
    else
        pad_level0_kernel<4><<<grid, dim3(PAD_TX, PAD_TY), 0, L.stream>>>(src_a, src_b, (long long)src_pitch, (long long)src_stride, dst_a, dst_b,
                                                                        dst_pitch, (long long)dst_stride, w, h, ax, ay, n_a);
    L.err = cudaGetLastError();
    L.launches++;
}

bool pyramid_check_read(unsigned long long out[4]) { return check_read_tu(out); }

void launch_box_half(Launch& L, const uint8_t* src, int w, int h, long long row_stride, long long src_img_stride,
                     uint8_t* dst, long long dst_img_stride, int n_img, int sse2_rounding)
{
    if (L.err != cudaSuccess || n_img <= 0) return;
    const int out_w = w / 2, out_h = h / 2;
    dim3 grid((unsigned)(((long long)((out_w + 7) / 8) * out_h + 127) / 128), n_img);
    box_half_kernel<<<grid, 128, 0, L.stream>>>(src, out_w, out_h, row_stride, src_img_stride, dst, dst_img_stride,
                                                sse2_rounding);
    L.err = cudaGetLastError();
    L.launches++;
}

}  // namespace dr3lk
