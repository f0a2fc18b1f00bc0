// C ABI of dr3lk (include/dr3lk.h): context, device scratch management, parameter normalisation and the
// orchestration of the pyramid + LK kernels.  Host code only; the kernels live in pyramid.cu / lk_*.cu.
#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <new>
#include <string>
#include <thread>
#include <vector>

#include "dr3lk_internal.cuh"

using namespace dr3lk;

#ifndef DR3LK_TMA_DEFAULT
#define DR3LK_TMA_DEFAULT 0
#endif

namespace {

std::string g_create_error;

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    // != 0: the buffer holds a derivative pyramid of this layout whose aprons are zero (kernels only ever write the interior
    // of a level), so a pooled buffer that is reused for the same layout needs no clearing
    unsigned long long zero_sig = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0; zero_sig = 0;
        // grow with slack so that slowly growing workloads do not reallocate every call
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want; else p = nullptr;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; zero_sig = 0; }
};

struct HostBuf {
    void* p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes)
    {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        cudaError_t e = cudaMallocHost(&p, bytes + bytes / 8 + 256);
        if (e == cudaSuccess) cap = bytes + bytes / 8 + 256;
        return e;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

inline int align_up(int v, int a) { return (v + a - 1) / a * a; }
inline size_t align_up_sz(size_t v, size_t a) { return (v + a - 1) / a * a; }

// Sizes / pitches of the Gaussian pyramid calcOpticalFlowPyrLK builds.
struct PyrLayout {
    int ml = 0;  // effective maxLevel
    int w[kMaxLevels], h[kMaxLevels];
    int pitch[kMaxLevels];   // bytes, 16-B aligned (scratch levels)
    int dpitch[kMaxLevels];  // ints, 16-B aligned
    size_t img_bytes[kMaxLevels];   // per image, aprons included
    size_t der_ints[kMaxLevels];    // per image, aprons included
    // Aprons (dr3lk_internal.cuh) around every level when the window has a specialised LK kernel, else 0; with aprons
    // level 0 is an apron-carrying scratch copy too.  img_org / der_org: offset of pixel (0, 0) inside one image.
    int ax = 0, ay = 0, dax = 0;
    size_t img_org[kMaxLevels], der_org[kMaxLevels];
};

PyrLayout make_layout(int w, int h, int win_w, int win_h, int max_level)
{
    PyrLayout P;
    int ws[kMaxLevels], hs[kMaxLevels];
    P.ml = dr3lk_lk_level_sizes(w, h, win_w, win_h, std::min(max_level, kMaxLevels - 1), ws, hs);
    if (lk_fast_supported(win_w, win_h)) { P.ax = kApronX; P.ay = apron_y(win_h); P.dax = deriv_apron_x(win_w); }
    for (int l = 0; l <= P.ml; l++) {
        P.w[l] = ws[l]; P.h[l] = hs[l];
        P.pitch[l] = align_up(ws[l] + 2 * P.ax, 16);
        P.dpitch[l] = align_up(ws[l] + 2 * P.dax, 4);
        P.img_bytes[l] = (size_t)P.pitch[l] * (hs[l] + 2 * P.ay);
        P.der_ints[l] = (size_t)P.dpitch[l] * (hs[l] + 2 * P.ay);
        P.img_org[l] = (size_t)P.ay * P.pitch[l] + P.ax;
        P.der_org[l] = (size_t)P.ay * P.dpitch[l] + P.dax;
    }
    return P;
}

// Device scratch for one batch of frame pairs: Gaussian levels >= 1 of both frames, derivatives of the previous
// frame at every level, optionally level-0 copies (host-buffer entry points), and the point arrays.
struct Workspace {
    DevBuf pyr_prev, pyr_next, deriv, lvl0_prev, lvl0_next, pts, offs, pair_idx, counter;
    int epoch = 0;  // LK launches on this workspace (selects the work counter)
    // The derivative aprons are zeros that no kernel ever writes: they are cleared once per (allocation, layout).
    struct DerivSig {
        const void* p = nullptr;
        size_t cap = 0;
        int w = 0, h = 0, win_w = 0, win_h = 0, ml = 0, batch = 0;
        bool operator==(const DerivSig& o) const
        {
            return p == o.p && cap == o.cap && w == o.w && h == o.h && win_w == o.win_w && win_h == o.win_h && ml == o.ml && batch == o.batch;
        }
    } deriv_sig;
    void release()
    {
        pyr_prev.release(); pyr_next.release(); deriv.release(); lvl0_prev.release(); lvl0_next.release(); pts.release();
        offs.release(); pair_idx.release(); counter.release();
        deriv_sig = DerivSig();
    }
};

struct LKArgs {
    int win_w, win_h, max_level, crit_type, crit_max_count, flags;
    double crit_eps, min_eig_threshold;
};

}  // namespace

struct dr3lk_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr;
    cudaStream_t stream = nullptr;
    std::string err;
    uint64_t launches = 0;
    int fast_arc = 10;  // dr3lk_debug_set_fast_arc
    Workspace ws;                 // single-call / device-batch scratch
    HostBuf pinned;               // staging for the single-pair host call
    std::vector<DevBuf> pool;     // device buffers of destroyed dr3lk_pyramid objects, reused by the next create
    std::vector<struct dr3lk_pyramid*> live;  // pyramid objects of this context that have not been destroyed yet: dr3lk_destroy
                                              // frees their device buffers and orphans them (ctx = nullptr), see dr3lk.h
    DevBuf take(size_t bytes)
    {
        for (size_t i = 0; i < pool.size(); i++)
            if (pool[i].cap >= bytes && pool[i].cap <= bytes + bytes / 4 + 4096) {
                DevBuf b = pool[i];
                pool.erase(pool.begin() + i);
                return b;
            }
        return DevBuf();
    }
    // TMA descriptors already encoded for this context's buffers (key: address + geometry); encoding costs about a
    // microsecond each, the single-call path would pay 12 of them per call
    struct TmaKey {
        const void* base;
        unsigned long long dim[3], stride[2];
        unsigned box[2];
        int dtype;
        bool operator==(const TmaKey& o) const { return memcmp(this, &o, sizeof(TmaKey)) == 0; }
    };
    std::vector<std::pair<TmaKey, CUtensorMap>> tma_cache;
    // compacted (ref, cur) points of the last dr3lk_init_second_frame, kept on the device for dr3lk_init_score_fundamental
    DevBuf tracks;
    int n_tracks = 0;
    size_t tracks_cur_offset = 0;  // byte offset of the compacted current points inside `tracks`
    bool profiling = false;
    struct Prof { cudaEvent_t e[3]; };  // pyramid start, LK start, LK end
    std::vector<Prof> prof;
    static constexpr int kSlots = 6;   // capacity; dr3lk_track_batch_host uses n_slots() of them (3 unless DR3LK_SLOTS says otherwise)
    Workspace slot_ws[kSlots];    // chunk pipeline of dr3lk_track_batch_host
    cudaStream_t slot_stream[kSlots] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    static int n_slots()
    {
        static const int n = getenv("DR3LK_SLOTS") ? std::min(kSlots, std::max(1, atoi(getenv("DR3LK_SLOTS")))) : 3;  // tuning knob
        return n;
    }
};

namespace {

int fail(dr3lk_ctx* ctx, int code, const std::string& msg)
{
    if (ctx) ctx->err = msg;
    return code;
}

int fail_cuda(dr3lk_ctx* ctx, cudaError_t e, const char* what)
{
    // clear the sticky-free error state so a later call can succeed
    cudaGetLastError();
    return fail(ctx, DR3LK_E_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

#define CU_TRY(ctx, expr)                                           \
    do {                                                            \
        cudaError_t e__ = (expr);                                   \
        if (e__ != cudaSuccess) return fail_cuda(ctx, e__, #expr);  \
    } while (0)

int check_lk_args(dr3lk_ctx* ctx, int w, int h, const LKArgs& a)
{
    // CV_Assert( maxLevel >= 0 && winSize.width > 2 && winSize.height > 2 )
    if (a.max_level < 0 || a.win_w <= 2 || a.win_h <= 2)
        return fail(ctx, DR3LK_E_ARG, "(-215:Assertion failed) maxLevel >= 0 && winSize.width > 2 && winSize.height > 2");
    if (w < 1 || h < 1) return fail(ctx, DR3LK_E_SIZE, "empty image");
    if ((long long)a.win_w * a.win_h > 96 * 96) return fail(ctx, DR3LK_E_SIZE, "window larger than 96x96 is not supported");
    return DR3LK_OK;
}

// Builds both Gaussian pyramids and the Scharr derivatives for `batch` pairs.  prev0/next0: device level-0 images (any
// alignment).  With aprons (P.ax > 0) level 0 is first copied into its apron-carrying scratch image.  Fills the `lk`
// level descriptors (pointers at pixel (0, 0)) and lk.fast_ok.  Scratch comes from `W`.
// Latency path (a few frame pairs at most): the pyramid stage is a chain of small dependent kernels, launched with
// programmatic stream serialization (Launch::pdl).  Measured on one box (tools/latency_ab.sh, profiles/latency_r02_ab.txt): PDL
// takes 9 us off a 106 us call.  Running the level-0 apron copy on a second stream beside the level kernels was also tried: it
// shortens the stage without PDL (35.6 -> 30.8 us) but costs 2 us with it (the event join breaks a dependent-launch edge and
// the first level kernel reads the raw image through its slower edge path), so it is not done.
constexpr int kLatencyMaxBatch = 4;
static bool pdl_enabled()
{
    static const bool on = getenv("DR3LK_NO_PDL") == nullptr;  // measurement knob
    return on;
}

int build_pyramids(dr3lk_ctx* ctx, Workspace& W, cudaStream_t stream, const uint8_t* prev0, const uint8_t* next0, size_t pitch0,
                   size_t stride0, int batch, const PyrLayout& P, int win_w, int win_h, LKParams& lk)
{
    const bool apr = P.ax > 0;
    // the kernels address inside one level image / derivative image with 32-bit byte offsets
    if (stride0 >= (1ull << 32) || P.img_bytes[0] >= (1ull << 31) || P.der_ints[0] >= (1ull << 29) || pitch0 >= (1ull << 31))
        return fail(ctx, DR3LK_E_SIZE, "images of 2 GiB or more (derivatives included) are not supported");
    size_t pyr_bytes = 0, der_ints = 0;
    size_t lvl_off[kMaxLevels] = {0}, der_off[kMaxLevels] = {0};
    for (int l = 0; l <= P.ml; l++) {
        if (l >= 1 || apr) { lvl_off[l] = pyr_bytes; pyr_bytes += P.img_bytes[l] * batch; }
        der_off[l] = der_ints; der_ints += P.der_ints[l] * batch;
    }
    cudaSetDevice(ctx->device);
    if (pyr_bytes) {
        CU_TRY(ctx, W.pyr_prev.reserve(pyr_bytes));
        CU_TRY(ctx, W.pyr_next.reserve(pyr_bytes));
    }
    CU_TRY(ctx, W.deriv.reserve(der_ints * sizeof(int)));
    if (apr) {
        Workspace::DerivSig sig;
        sig.p = W.deriv.p; sig.cap = W.deriv.cap;
        sig.w = P.w[0]; sig.h = P.h[0]; sig.win_w = win_w; sig.win_h = win_h; sig.ml = P.ml; sig.batch = batch;
        if (!(sig == W.deriv_sig)) {
            CU_TRY(ctx, cudaMemsetAsync(W.deriv.p, 0, der_ints * sizeof(int), stream));
            W.deriv_sig = sig;
        }
    } else {
        W.deriv_sig = Workspace::DerivSig();  // an apron-less layout is about to overwrite the zeros
    }

    for (int l = 0; l <= P.ml; l++) {
        LevelDesc& d = lk.lv[l];
        d.w = P.w[l]; d.h = P.h[l];
        if (l == 0 && !apr) {
            d.prev = prev0; d.next = next0;
            d.pitch_p = d.pitch_n = (int)pitch0;
            d.prev_stride = d.next_stride = (unsigned)stride0;
        } else {
            d.prev = (const uint8_t*)W.pyr_prev.p + lvl_off[l] + P.img_org[l];
            d.next = (const uint8_t*)W.pyr_next.p + lvl_off[l] + P.img_org[l];
            d.pitch_p = d.pitch_n = P.pitch[l];
            d.prev_stride = d.next_stride = (unsigned)P.img_bytes[l];
        }
        d.deriv = (const int*)W.deriv.p + der_off[l] + P.der_org[l];
        d.dpitch = P.dpitch[l];
        d.deriv_stride = (unsigned)P.der_ints[l];
    }
    lk.max_level = P.ml;
    lk.fast_ok = apr ? 1 : 0;

    Launch L{stream, cudaSuccess, 0};
    L.pdl = batch <= kLatencyMaxBatch && pdl_enabled();  // latency path: dependent launches overlap the tail of their predecessor
    if (apr)
        launch_pad_level0(L, prev0, next0, pitch0, stride0, const_cast<uint8_t*>(lk.lv[0].prev), const_cast<uint8_t*>(lk.lv[0].next), P.pitch[0],
                          P.img_bytes[0], P.w[0], P.h[0], P.ax, P.ay, batch, next0 ? batch : 0);
    for (int l = 0; l <= P.ml; l++) {
        const LevelDesc& s = lk.lv[l];
        const bool down = l < P.ml;
        PyrLevelArgs a{};
        a.prev_src = s.prev; a.next_src = s.next;
        a.prev_src_stride = s.prev_stride; a.next_src_stride = s.next_stride;
        a.w = s.w; a.h = s.h; a.src_pitch = s.pitch_p;
        a.deriv = const_cast<int*>(s.deriv); a.dpitch = s.dpitch; a.deriv_stride = s.deriv_stride;
        a.n_prev = batch; a.n_next = (down && next0) ? batch : 0;
        a.down = down;
        a.dst_apron_x = P.ax; a.dst_apron_y = P.ay;
        a.src_apron_x = P.ax; a.src_apron_y = P.ay;
        if (down) {
            a.prev_dst = const_cast<uint8_t*>(lk.lv[l + 1].prev); a.next_dst = const_cast<uint8_t*>(lk.lv[l + 1].next);
            a.prev_dst_stride = lk.lv[l + 1].prev_stride; a.next_dst_stride = lk.lv[l + 1].next_stride;
            a.dst_pitch = lk.lv[l + 1].pitch_p;
        }
        launch_pyr_level(L, a);
    }
    ctx->launches += L.launches;
    if (L.err != cudaSuccess) return fail_cuda(ctx, L.err, "pyramid kernel launch");
    return DR3LK_OK;
}

void fill_lk_scalars(LKParams& lk, const LKArgs& a)
{
    // parameter normalisation of calcOpticalFlowPyrLK
    lk.max_count = (a.crit_type & DR3LK_TERM_COUNT) ? std::min(std::max(a.crit_max_count, 0), 100) : 30;
    double eps = (a.crit_type & DR3LK_TERM_EPS) ? std::min(std::max(a.crit_eps, 0.), 10.) : 0.01;
    lk.eps2 = eps * eps;
    lk.eps2_lo = std::nextafterf((float)(lk.eps2 * (1.0 - 1e-6)), 0.f);
    lk.eps2_hi = std::nextafterf((float)(lk.eps2 * (1.0 + 1e-6)), INFINITY);
    lk.min_eig_thr = (float)a.min_eig_threshold;
    lk.win_w = a.win_w; lk.win_h = a.win_h;
    lk.flags = a.flags;
}

// ---- TMA descriptors for the specialised LK kernels (cp.async.bulk.tensor staging) ----
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeTiledFn tma_encoder()
{
    static EncodeTiledFn fn = []() -> EncodeTiledFn {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) {
            cudaGetLastError();
            return nullptr;
        }
        return (EncodeTiledFn)p;
    }();
    return fn;
}

// One rank-3 box descriptor over [batch][rows][pitch] elements of `esize` bytes starting at `base`; cached per context.
bool tma_map(dr3lk_ctx* ctx, CUtensorMap* out, const void* base, int esize, unsigned long long pitch_elems, unsigned long long rows,
             unsigned long long batch, unsigned long long img_stride_bytes, unsigned box_w, unsigned box_h)
{
    dr3lk_ctx::TmaKey k;
    memset(&k, 0, sizeof(k));
    k.base = base; k.dim[0] = pitch_elems; k.dim[1] = rows; k.dim[2] = batch;
    k.stride[0] = pitch_elems * esize; k.stride[1] = img_stride_bytes; k.box[0] = box_w; k.box[1] = box_h; k.dtype = esize;
    for (auto& e : ctx->tma_cache)
        if (e.first == k) { *out = e.second; return true; }
    EncodeTiledFn enc = tma_encoder();
    if (!enc) return false;
    const cuuint64_t gdim[3] = {k.dim[0], k.dim[1], k.dim[2]};
    const cuuint64_t gstr[2] = {k.stride[0], batch > 1 ? k.stride[1] : k.stride[0] * rows};
    const cuuint32_t box[3] = {box_w, box_h, 1}, estr[3] = {1, 1, 1};
    if (enc(out, esize == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_INT32, 3, const_cast<void*>(base), gdim, gstr, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (ctx->tma_cache.size() >= 256) ctx->tma_cache.clear();
    ctx->tma_cache.emplace_back(k, *out);
    return true;
}

// Fills lk.tma[] for levels that carry the aprons (lk.fast_ok) and sets lk.use_tma.  Staging with TMA is opt-in / opt-out
// through DR3LK_TMA (see DESIGN.md for the measurement behind the default).
void fill_tma(dr3lk_ctx* ctx, LKParams& lk, int batch)
{
    static const int want = getenv("DR3LK_TMA") ? atoi(getenv("DR3LK_TMA")) : DR3LK_TMA_DEFAULT;
    lk.use_tma = 0;
    LkFastBoxes b;
    if (!want || !lk.fast_ok || lk.max_level >= kTmaLevels || !lk_fast_boxes(lk.win_w, lk.win_h, &b)) return;
    const int ax = kApronX, ay = apron_y(lk.win_h), dax = deriv_apron_x(lk.win_w);
    for (int l = 0; l <= lk.max_level; l++) {
        const LevelDesc& d = lk.lv[l];
        const unsigned long long rows = (unsigned long long)d.h + 2 * ay;
        if (!tma_map(ctx, &lk.tma[l].prev, d.prev - ((size_t)ay * d.pitch_p + ax), 1, d.pitch_p, rows, batch, d.prev_stride, b.i_w, b.i_h)) return;
        if (!tma_map(ctx, &lk.tma[l].next, d.next - ((size_t)ay * d.pitch_n + ax), 1, d.pitch_n, rows, batch, d.next_stride, b.j_w, b.j_h)) return;
        if (!tma_map(ctx, &lk.tma[l].deriv, d.deriv - ((size_t)ay * d.dpitch + dax), 4, d.dpitch, rows, batch, (unsigned long long)d.deriv_stride * 4,
                     b.d_w, b.d_h))
            return;
    }
    lk.use_tma = 1;
}

int run_lk(dr3lk_ctx* ctx, cudaStream_t stream, const LKParams& lk, Workspace& W)
{
    Launch L{stream, cudaSuccess, 0};
    L.pdl = lk.batch <= kLatencyMaxBatch && pdl_enabled();  // latency path, see Launch::pdl
    if (launch_lk_fast(L, lk)) W.epoch++;  // the persistent kernel consumed one work counter and re-armed the other
    else launch_lk_generic(L, lk);
    ctx->launches += L.launches;
    if (L.err != cudaSuccess) return fail_cuda(ctx, L.err, "LK kernel launch");
    return DR3LK_OK;
}

// LK over level descriptors that are already filled in (lk.lv[0..max_level], lk.max_level, lk.fast_ok).
int run_tracking(dr3lk_ctx* ctx, Workspace& W, cudaStream_t stream, LKParams& lk, int batch, const float* prev_pts_dev,
                 float* next_pts_dev, uint8_t* status_dev, float* err_dev, const int* pts_offset, const int* pts_offset_dev, int n_total,
                 uint32_t* stats_dev, const LKArgs& a, float* next_out_dev = nullptr)
{
    fill_lk_scalars(lk, a);
    lk.prev_pts = (const float2*)prev_pts_dev;
    lk.next_pts = (float2*)next_pts_dev;
    lk.next_out = (float2*)(next_out_dev ? next_out_dev : next_pts_dev);
    lk.status = status_dev;
    lk.err = err_dev;
    lk.stats = stats_dev;
    lk.batch = batch;
    lk.n_total = n_total;
    if (!W.counter.p) {
        CU_TRY(ctx, W.counter.reserve(2 * sizeof(int)));
        CU_TRY(ctx, cudaMemsetAsync(W.counter.p, 0, 2 * sizeof(int), stream));
    }
    lk.work_counter = (int*)W.counter.p;
    lk.work_epoch = W.epoch;
    // point -> pair mapping: a division when every pair has the same number of points, else a lookup table
    bool uniform = n_total % batch == 0;
    for (int b = 0; uniform && b < batch; b++) uniform = (pts_offset[b + 1] - pts_offset[b]) == n_total / batch;
    if (uniform) {
        lk.uniform_n = n_total / batch;
    } else {
        CU_TRY(ctx, W.pair_idx.reserve(sizeof(int) * (size_t)n_total));
        Launch L{stream, cudaSuccess, 0};
        launch_pair_index(L, pts_offset_dev, batch, n_total, (int*)W.pair_idx.p);
        ctx->launches += L.launches;
        if (L.err != cudaSuccess) return fail_cuda(ctx, L.err, "pair index kernel launch");
        lk.pair_idx = (const int*)W.pair_idx.p;
    }
    fill_tma(ctx, lk, batch);
    return run_lk(ctx, stream, lk, W);
}

// Device-resident batch on (W, stream).  pts_offset: host offsets, pts_offset_dev: device copy of the same.
int track_batch_device(dr3lk_ctx* ctx, Workspace& W, cudaStream_t stream, const uint8_t* prev_dev, const uint8_t* next_dev, int w,
                       int h, size_t pitch, size_t image_stride, int batch, const float* prev_pts_dev, float* next_pts_dev,
                       uint8_t* status_dev, float* err_dev, const int* pts_offset, const int* pts_offset_dev, int n_total,
                       uint32_t* stats_dev, const LKArgs& a, float* next_out_dev = nullptr)
{
    // the pyramid kernels put the image index (both frames of every pair) in gridDim.z
    if (batch > 32767) return fail(ctx, DR3LK_E_SIZE, "at most 32767 frame pairs per device-resident call: split the batch");
    LKParams lk;
    memset(&lk, 0, sizeof(lk));
    PyrLayout P = make_layout(w, h, a.win_w, a.win_h, a.max_level);
    dr3lk_ctx::Prof pr;
    if (ctx->profiling) {
        for (int i = 0; i < 3; i++) CU_TRY(ctx, cudaEventCreate(&pr.e[i]));
        CU_TRY(ctx, cudaEventRecord(pr.e[0], stream));
    }
    int rc = build_pyramids(ctx, W, stream, prev_dev, next_dev, pitch, image_stride, batch, P, a.win_w, a.win_h, lk);
    if (rc != DR3LK_OK) return rc;
    if (ctx->profiling) CU_TRY(ctx, cudaEventRecord(pr.e[1], stream));
    rc = run_tracking(ctx, W, stream, lk, batch, prev_pts_dev, next_pts_dev, status_dev, err_dev, pts_offset, pts_offset_dev, n_total,
                      stats_dev, a, next_out_dev);
    if (ctx->profiling) {
        CU_TRY(ctx, cudaEventRecord(pr.e[2], stream));
        ctx->prof.push_back(pr);
    }
    return rc;
}

int check_offsets(dr3lk_ctx* ctx, const int* pts_offset, int batch)
{
    if (!pts_offset || pts_offset[0] != 0) return fail(ctx, DR3LK_E_ARG, "pts_offset[0] must be 0");
    for (int b = 0; b < batch; b++)
        if (pts_offset[b + 1] < pts_offset[b]) return fail(ctx, DR3LK_E_ARG, "pts_offset must be non-decreasing");
    return DR3LK_OK;
}

}  // namespace

extern "C" {

int dr3lk_create(dr3lk_ctx** out, int device)
{
    if (!out) return DR3LK_E_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n == 0) {
        cudaGetLastError();
        g_create_error = std::string("no usable CUDA device (there is no CPU fallback): ") + cudaGetErrorString(e);
        return DR3LK_E_CUDA;
    }
    if (device < 0 || device >= n) { g_create_error = "device index out of range"; return DR3LK_E_ARG; }
    e = cudaSetDevice(device);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); return DR3LK_E_CUDA; }
    dr3lk_ctx* c = new (std::nothrow) dr3lk_ctx();
    if (!c) return DR3LK_E_CUDA;
    c->device = device;
    e = cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking);
    if (e != cudaSuccess) { g_create_error = cudaGetErrorString(e); delete c; return DR3LK_E_CUDA; }
    c->stream = c->own_stream;
    *out = c;
    return DR3LK_OK;
}

static void orphan_pyramids(dr3lk_ctx* ctx);

void dr3lk_destroy(dr3lk_ctx* ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    orphan_pyramids(ctx);
    ctx->ws.release();
    ctx->pinned.release();
    ctx->tracks.release();
    for (auto& b : ctx->pool) b.release();
    for (int i = 0; i < dr3lk_ctx::kSlots; i++) {
        ctx->slot_ws[i].release();
        if (ctx->slot_stream[i]) cudaStreamDestroy(ctx->slot_stream[i]);
    }
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

const char* dr3lk_last_error(const dr3lk_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int dr3lk_set_stream(dr3lk_ctx* ctx, void* cuda_stream)
{
    if (!ctx) return DR3LK_E_ARG;
    cudaStream_t want = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    if (want != ctx->stream) {
        // Scratch state is ordered by the stream it was last used on (cleared derivative aprons, the LK work counters,
        // pooled pyramid buffers): finish that work before anything is enqueued on another stream.
        cudaSetDevice(ctx->device);
        CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
        ctx->stream = want;
    }
    return DR3LK_OK;
}

int dr3lk_synchronize(dr3lk_ctx* ctx)
{
    if (!ctx) return DR3LK_E_ARG;
    cudaSetDevice(ctx->device);
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DR3LK_OK;
}

uint64_t dr3lk_launch_count(const dr3lk_ctx* ctx) { return ctx ? ctx->launches : 0; }

int dr3lk_debug_set_fast_arc(dr3lk_ctx* ctx, int arc)
{
    if (!ctx) return DR3LK_E_ARG;
    if (arc != 9 && arc != 10) return fail(ctx, DR3LK_E_ARG, "fast arc length must be 9 (OpenCV FAST-9, parity hook) or 10 (the reference's detector)");
    ctx->fast_arc = arc;
    return DR3LK_OK;
}

int dr3lk_debug_check_read(dr3lk_ctx* ctx, unsigned long long* out4)
{
    if (!ctx || !out4) return DR3LK_E_ARG;
    cudaSetDevice(ctx->device);
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    unsigned long long pyr[4] = {0, 0, 0, 0};
    if (!lk_fast_check_read(out4) || !pyramid_check_read(pyr))
        return fail(ctx, DR3LK_E_UNSUPPORTED, "this library was built without -DDR3LK_CHECKED (make -C 3dr_b200/csrc checked)");
    // one set of counters per translation unit: LK kernels (kinds 1..8) + pyramid kernels (kinds 20..)
    if (out4[0] == 0) { out4[1] = pyr[1]; out4[2] = pyr[2]; }
    out4[0] += pyr[0];
    out4[3] += pyr[3];
    return DR3LK_OK;
}

int dr3lk_set_profiling(dr3lk_ctx* ctx, int on)
{
    if (!ctx) return DR3LK_E_ARG;
    ctx->profiling = on != 0;
    return DR3LK_OK;
}

int dr3lk_profile_read(dr3lk_ctx* ctx, float* lk_ms, int* lk_launches, float* pyramid_ms, int* pyramid_builds)
{
    if (!ctx) return DR3LK_E_ARG;
    cudaSetDevice(ctx->device);
    float lk = 0.f, py = 0.f;
    int n = 0;
    for (auto& pr : ctx->prof) {
        CU_TRY(ctx, cudaEventSynchronize(pr.e[2]));
        float a = 0.f, b = 0.f;
        CU_TRY(ctx, cudaEventElapsedTime(&a, pr.e[0], pr.e[1]));
        CU_TRY(ctx, cudaEventElapsedTime(&b, pr.e[1], pr.e[2]));
        py += a; lk += b; n++;
        for (int i = 0; i < 3; i++) cudaEventDestroy(pr.e[i]);
    }
    ctx->prof.clear();
    if (lk_ms) *lk_ms = lk;
    if (lk_launches) *lk_launches = n;
    if (pyramid_ms) *pyramid_ms = py;
    if (pyramid_builds) *pyramid_builds = n;
    return DR3LK_OK;
}

void* dr3lk_host_alloc(size_t bytes)
{
    void* p = nullptr;
    if (cudaMallocHost(&p, bytes) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return p;
}

void dr3lk_host_free(void* p) { if (p) cudaFreeHost(p); }

int dr3lk_host_register(void* p, size_t bytes)
{
    if (!p || bytes == 0) return DR3LK_E_ARG;
    if (cudaHostRegister(p, bytes, cudaHostRegisterPortable) != cudaSuccess) { cudaGetLastError(); return DR3LK_E_CUDA; }
    return DR3LK_OK;
}

int dr3lk_host_unregister(void* p)
{
    if (!p) return DR3LK_E_ARG;
    if (cudaHostUnregister(p) != cudaSuccess) { cudaGetLastError(); return DR3LK_E_CUDA; }
    return DR3LK_OK;
}

int dr3lk_lk_level_sizes(int w, int h, int win_w, int win_h, int max_level, int* ws, int* hs)
{
    // buildOpticalFlowPyramid: level l+1 = ((w+1)/2, (h+1)/2); stop when the NEXT level would not exceed the window
    int level = 0;
    for (;; level++) {
        ws[level] = w; hs[level] = h;
        if (level >= max_level) break;
        const int nw = (w + 1) / 2, nh = (h + 1) / 2;
        if (nw <= win_w || nh <= win_h) break;
        w = nw; h = nh;
    }
    return level;
}

/* ---------------------------------------------------------------------------------------------- */
/* box pyramid                                                                                     */
/* ---------------------------------------------------------------------------------------------- */

static int box_level_mode(dr3lk_ctx* ctx, int w, int h, long long stride, int mode, bool aligned16, int* sse2)
{
    if (w < 2 || h < 2 || stride < w) return fail(ctx, DR3LK_E_SIZE, "box pyramid: level smaller than 2x2 or step < width");
    *sse2 = (mode == DR3LK_BOX_SSE2) || (mode == DR3LK_BOX_AUTO_X86 && (w % 16) == 0 && aligned16);
    if (*sse2) {
        if (w % 16) return fail(ctx, DR3LK_E_ARG, "box pyramid: SSE2 rounding needs cols % 16 == 0 (src/utils.cpp:389)");
        if (stride != w) return fail(ctx, DR3LK_E_UNSUPPORTED, "box pyramid: halfSampleSSE2 assumes a continuous image");
        return DR3LK_OK;
    }
    // Scalar walk of reduce_to_half (src/utils.cpp:401-418): make sure the reference itself stays in bounds.
    const long long out_w = w / 2, out_h = h / 2, end = stride * (long long)h;
    long long bottom = stride, rows = 0;
    while (bottom < end) {
        if (bottom + 2 * out_w - 1 >= end) return fail(ctx, DR3LK_E_UNSUPPORTED, "box pyramid: the reference reads past its input for this shape");
        bottom += 2 * out_w + stride;
        rows++;
    }
    if (rows > out_h) return fail(ctx, DR3LK_E_UNSUPPORTED, "box pyramid: the reference writes past its output for this shape (odd cols with odd rows)");
    if (rows < out_h) return fail(ctx, DR3LK_E_UNSUPPORTED, "box pyramid: the reference leaves output rows unwritten for this shape");
    return DR3LK_OK;
}

int dr3lk_box_pyramid_device(dr3lk_ctx* ctx, const uint8_t* img_dev, int w, int h, size_t pitch, size_t image_stride, int batch,
                             int n_levels, uint8_t* const* out_levels_dev, int mode)
{
    if (!ctx) return DR3LK_E_ARG;
    if (!img_dev || n_levels < 1 || batch < 1 || (n_levels > 1 && !out_levels_dev)) return fail(ctx, DR3LK_E_ARG, "box pyramid: bad argument");
    if (mode < DR3LK_BOX_AUTO_X86 || mode > DR3LK_BOX_SSE2) return fail(ctx, DR3LK_E_ARG, "box pyramid: bad mode");
    if (batch > 65535) return fail(ctx, DR3LK_E_SIZE, "box pyramid: at most 65535 images per call (the image index is gridDim.y): split the batch");
    cudaSetDevice(ctx->device);
    Launch L{ctx->stream, cudaSuccess, 0};
    const uint8_t* src = img_dev;
    long long row_stride = (long long)pitch, img_stride = (long long)image_stride;
    for (int l = 1; l < n_levels; l++) {
        int sse2 = 0;
        int rc = box_level_mode(ctx, w, h, row_stride, mode, true, &sse2);
        if (rc != DR3LK_OK) return rc;
        const long long dst_stride = (long long)(w / 2) * (h / 2);
        launch_box_half(L, src, w, h, row_stride, img_stride, out_levels_dev[l - 1], dst_stride, batch, sse2);
        src = out_levels_dev[l - 1];
        w /= 2; h /= 2;
        row_stride = w; img_stride = dst_stride;
    }
    ctx->launches += L.launches;
    if (L.err != cudaSuccess) return fail_cuda(ctx, L.err, "box pyramid kernel launch");
    return DR3LK_OK;
}

int dr3lk_box_pyramid(dr3lk_ctx* ctx, const uint8_t* img, int w, int h, size_t step, int n_levels, uint8_t* const* out_levels,
                      int mode)
{
    if (!ctx) return DR3LK_E_ARG;
    if (!img || n_levels < 1 || (n_levels > 1 && !out_levels)) return fail(ctx, DR3LK_E_ARG, "box pyramid: bad argument");
    if (mode < DR3LK_BOX_AUTO_X86 || mode > DR3LK_BOX_SSE2) return fail(ctx, DR3LK_E_ARG, "box pyramid: bad mode");
    if (n_levels == 1) return DR3LK_OK;
    if (n_levels > kMaxLevels) return fail(ctx, DR3LK_E_ARG, "box pyramid: too many levels");
    cudaSetDevice(ctx->device);
    // validate every level first (the reference would have crashed / corrupted memory on the rejected shapes)
    {
        int lw = w, lh = h;
        long long st = (long long)step;
        bool aligned = (reinterpret_cast<uintptr_t>(img) & 0xF) == 0;  // is_aligned16(in.data), src/utils.cpp:387
        for (int l = 1; l < n_levels; l++) {
            int sse2;
            int rc = box_level_mode(ctx, lw, lh, st, mode, aligned, &sse2);
            if (rc != DR3LK_OK) return rc;
            lw /= 2; lh /= 2; st = lw; aligned = true;  // fresh cv::Mat levels are 16-B aligned and continuous
        }
    }
    // level 0 is copied verbatim (h*step bytes, the walk addresses the flat buffer); levels >= 1 are packed behind it
    const size_t l0_bytes = (size_t)step * (h - 1) + w;  // the last row of an ROI may end before `step`
    size_t total = align_up_sz(l0_bytes, 256);
    size_t off[kMaxLevels] = {0};
    {
        int lw = w, lh = h;
        for (int l = 1; l < n_levels; l++) {
            lw /= 2; lh /= 2;
            off[l] = total;
            total += align_up_sz((size_t)lw * lh, 256);
        }
    }
    CU_TRY(ctx, ctx->ws.lvl0_prev.reserve(total));
    uint8_t* base = (uint8_t*)ctx->ws.lvl0_prev.p;
    CU_TRY(ctx, cudaMemcpyAsync(base, img, l0_bytes, cudaMemcpyHostToDevice, ctx->stream));
    Launch L{ctx->stream, cudaSuccess, 0};
    const uint8_t* src = base;
    long long row_stride = (long long)step;
    bool aligned = (reinterpret_cast<uintptr_t>(img) & 0xF) == 0;
    int lw = w, lh = h;
    for (int l = 1; l < n_levels; l++) {
        int sse2 = 0;
        box_level_mode(ctx, lw, lh, row_stride, mode, aligned, &sse2);
        launch_box_half(L, src, lw, lh, row_stride, 0, base + off[l], 0, 1, sse2);
        lw /= 2; lh /= 2;
        src = base + off[l]; row_stride = lw; aligned = true;
    }
    ctx->launches += L.launches;
    if (L.err != cudaSuccess) return fail_cuda(ctx, L.err, "box pyramid kernel launch");
    lw = w; lh = h;
    for (int l = 1; l < n_levels; l++) {
        lw /= 2; lh /= 2;
        CU_TRY(ctx, cudaMemcpyAsync(out_levels[l - 1], base + off[l], (size_t)lw * lh, cudaMemcpyDeviceToHost, ctx->stream));
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DR3LK_OK;
}

/* ---------------------------------------------------------------------------------------------- */
/* LK                                                                                              */
/* ---------------------------------------------------------------------------------------------- */

int dr3lk_track_batch(dr3lk_ctx* ctx, const uint8_t* prev_dev, const uint8_t* next_dev, int w, int h, size_t pitch,
                      size_t image_stride, int batch, const float* prev_pts_dev, float* next_pts_dev, uint8_t* status_dev,
                      float* err_dev, const int* pts_offset, uint32_t* stats_dev, int win_w, int win_h, int max_level,
                      int crit_type, int crit_max_count, double crit_eps, int flags, double min_eig_threshold)
{
    if (!ctx) return DR3LK_E_ARG;
    LKArgs a{win_w, win_h, max_level, crit_type, crit_max_count, flags, crit_eps, min_eig_threshold};
    int rc = check_lk_args(ctx, w, h, a);
    if (rc != DR3LK_OK) return rc;
    if (batch < 1 || !prev_dev || !next_dev || pitch < (size_t)w) return fail(ctx, DR3LK_E_ARG, "track_batch: bad image arguments");
    rc = check_offsets(ctx, pts_offset, batch);
    if (rc != DR3LK_OK) return rc;
    const int n_total = pts_offset[batch];
    if (n_total == 0) return DR3LK_OK;
    if (!prev_pts_dev || !next_pts_dev || !status_dev) return fail(ctx, DR3LK_E_ARG, "track_batch: null point / status buffer");
    cudaSetDevice(ctx->device);
    CU_TRY(ctx, ctx->ws.offs.reserve(sizeof(int) * (size_t)(batch + 1)));
    CU_TRY(ctx, cudaMemcpyAsync(ctx->ws.offs.p, pts_offset, sizeof(int) * (size_t)(batch + 1), cudaMemcpyHostToDevice, ctx->stream));
    return track_batch_device(ctx, ctx->ws, ctx->stream, prev_dev, next_dev, w, h, pitch, image_stride, batch, prev_pts_dev,
                              next_pts_dev, status_dev, err_dev, pts_offset, (const int*)ctx->ws.offs.p, n_total, stats_dev, a);
}

// Latency path: the LK kernel writes next_pts / status / err of a single call straight into the context's pinned mirror
// (mapped into the device's address space under unified addressing), so that no device-to-host copy is queued behind it:
// the results are in host memory when the stream synchronises.  Returns the device alias of `host` or nullptr (more points
// than the posted PCIe writes are worth, or no mapping: the caller then copies back as before).
constexpr int kDirectOutMaxPoints = 16384;
static bool mapped_points_enabled()
{
    static const bool on = getenv("DR3LK_NO_MAPPED_PTS") == nullptr;  // measurement knob
    return on;
}
static uint8_t* mapped_alias(void* host, int n)
{
    static const bool disabled = getenv("DR3LK_NO_DIRECT_OUT") != nullptr;  // measurement knob
    if (disabled || n > kDirectOutMaxPoints) return nullptr;
    void* d = nullptr;
    if (cudaHostGetDevicePointer(&d, host, 0) != cudaSuccess) { cudaGetLastError(); return nullptr; }
    return (uint8_t*)d;
}

// The device row pitch at which an image is uploaded as it is -- ONE contiguous copy at its own row step, no packing into the
// context's pinned mirror -- or 0.  That is every image whose rows are continuous (step == w, the usual cv::Mat) or sit at
// the 16-byte aligned pitch; the level-0 apron copy reads rows of any alignment.  From page-locked memory (dr3lk_host_alloc,
// dr3lk_host_register, cudaHostAlloc) the copy is an asynchronous DMA: 85 us per KITTI call.  From pageable memory
// cudaMemcpyAsync stages through the driver's own pinned buffers and returns when the source has been consumed: 112 us,
// against 120 us for packing it ourselves (2 x 15 us of memcpy on the critical path).  Rows at any other step (ROIs) are packed:
// a 2-D DMA of 1241-byte rows into an aligned pitch was measured 36 us per call SLOWER than packing.
static size_t direct_pitch(const void* img, size_t step, int w, int pitch0)
{
    if (step != (size_t)pitch0 && step != (size_t)w) return 0;
    static const bool pack_pageable = getenv("DR3LK_PACK_PAGEABLE") != nullptr;  // measurement knob: the old behaviour
    if (!pack_pageable) return step;
    cudaPointerAttributes at;
    if (cudaPointerGetAttributes(&at, img) != cudaSuccess) { cudaGetLastError(); return 0; }
    return at.type == cudaMemoryTypeHost ? step : 0;
}
static size_t direct_bytes(size_t step, int w, int h) { return (size_t)(h - 1) * step + (size_t)w; }  // never past the last pixel

// Level-0 upload of one image (rows of w bytes at `step`); *dev_pitch receives the row pitch of the device copy (the image's own
// step when it went directly, else pitch0).  The caller synchronises the stream before it returns to its own caller, so the
// source only has to stay valid for the duration of the call.
static cudaError_t upload_image(uint8_t* dev, uint8_t* stage, int pitch0, const uint8_t* img, size_t step, int w, int h, cudaStream_t st,
                                size_t* dev_pitch)
{
    const size_t direct = direct_pitch(img, step, w, pitch0);
    *dev_pitch = direct ? direct : (size_t)pitch0;
    if (direct) return cudaMemcpyAsync(dev, img, direct_bytes(step, w, h), cudaMemcpyHostToDevice, st);
    for (int y = 0; y < h; y++) memcpy(stage + (size_t)y * pitch0, img + (size_t)y * step, (size_t)w);
    return cudaMemcpyAsync(dev, stage, (size_t)pitch0 * h, cudaMemcpyHostToDevice, st);
}

int dr3lk_calc_optical_flow_pyr_lk(dr3lk_ctx* ctx, const uint8_t* prev, size_t prev_step, const uint8_t* next, size_t next_step,
                                   int w, int h, const float* prev_pts, float* next_pts, uint8_t* status, float* err, int n,
                                   int win_w, int win_h, int max_level, int crit_type, int crit_max_count, double crit_eps,
                                   int flags, double min_eig_threshold)
{
    if (!ctx) return DR3LK_E_ARG;
    LKArgs a{win_w, win_h, max_level, crit_type, crit_max_count, flags, crit_eps, min_eig_threshold};
    int rc = check_lk_args(ctx, w, h, a);
    if (rc != DR3LK_OK) return rc;
    if (n < 0) return fail(ctx, DR3LK_E_ARG, "negative point count");
    if (n == 0) return DR3LK_OK;  // OpenCV releases the outputs and returns
    if (!prev || !next || prev_step < (size_t)w || next_step < (size_t)w) return fail(ctx, DR3LK_E_ARG, "bad image arguments");
    if (!prev_pts || !next_pts || !status) return fail(ctx, DR3LK_E_ARG, "null point / status buffer");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    Workspace& W = ctx->ws;
    const int pitch0 = align_up(w, 16);
    const size_t img_bytes = (size_t)pitch0 * h;
    // One device block and one pinned mirror of it: [prev image][next image][prev_pts 8n][offsets 16][next_pts 8n][err 4n][status n].
    // Inputs are packed on the host and cross PCIe as ONE copy; outputs come back as ONE copy (latency path of C1/C2).
    const size_t n8 = align_up_sz(8 * (size_t)n, 16);
    const size_t o_prev = 2 * img_bytes, o_offs = o_prev + n8, o_next = o_offs + 16, o_err = o_next + n8, o_status = o_err + align_up_sz(4 * (size_t)n, 16);
    const size_t total = o_status + align_up_sz((size_t)n, 16);
    CU_TRY(ctx, W.lvl0_prev.reserve(total));
    CU_TRY(ctx, ctx->pinned.reserve(total));
    uint8_t* dp = (uint8_t*)W.lvl0_prev.p;
    uint8_t* hp = (uint8_t*)ctx->pinned.p;
    // Images that are continuous or at the aligned pitch are uploaded as they are (direct_pitch; both frames share one device
    // pitch).  Any other is packed into the mirror first; then the previous image crosses PCIe while the host is still packing
    // the next one (the staging memcpy of ~0.5 MB per image is as long as its DMA).
    const size_t dp_prev = direct_pitch(prev, prev_step, w, pitch0), dp_next = direct_pitch(next, next_step, w, pitch0);
    const size_t pitch = (dp_prev && dp_prev == dp_next) ? dp_prev : (size_t)pitch0;
    const bool prev_direct = dp_prev == pitch, next_direct = dp_next == pitch;
    if (prev_direct) {
        CU_TRY(ctx, cudaMemcpyAsync(dp, prev, direct_bytes(prev_step, w, h), cudaMemcpyHostToDevice, st));
    } else {
        for (int y = 0; y < h; y++) memcpy(hp + (size_t)y * pitch0, prev + (size_t)y * prev_step, (size_t)w);
        CU_TRY(ctx, cudaMemcpyAsync(dp, hp, img_bytes, cudaMemcpyHostToDevice, st));
    }
    if (next_direct) CU_TRY(ctx, cudaMemcpyAsync(dp + img_bytes, next, direct_bytes(next_step, w, h), cudaMemcpyHostToDevice, st));
    else for (int y = 0; y < h; y++) memcpy(hp + img_bytes + (size_t)y * pitch0, next + (size_t)y * next_step, (size_t)w);
    memcpy(hp + o_prev, prev_pts, 8 * (size_t)n);
    const int offs[2] = {0, n};
    memcpy(hp + o_offs, offs, sizeof(offs));
    size_t in_bytes = o_next;
    if (flags & DR3LK_USE_INITIAL_FLOW) { memcpy(hp + o_next, next_pts, 8 * (size_t)n); in_bytes = o_next + 8 * (size_t)n; }
    uint8_t* const out = mapped_alias(hp, n);  // results straight into the pinned mirror, or ...
    uint8_t* const ob = out ? out : dp;
    // The points ride behind the next image when it was packed.  When it was not, a third copy would sit in
    // front of the kernels for a few hundred bytes: the LK kernel reads the points from the mapped mirror instead.
    const bool mapped_pts = next_direct && out && mapped_points_enabled();
    const size_t in_from = next_direct ? o_prev : img_bytes;
    if (!mapped_pts) CU_TRY(ctx, cudaMemcpyAsync(dp + in_from, hp + in_from, in_bytes - in_from, cudaMemcpyHostToDevice, st));
    const uint8_t* const pb = mapped_pts ? out : dp;
    // (the device copy of the offsets is only read for batches with differing point counts)
    rc = track_batch_device(ctx, W, st, dp, dp + img_bytes, w, h, pitch, img_bytes, 1, (const float*)(pb + o_prev), (float*)(pb + o_next),
                            ob + o_status, err ? (float*)(ob + o_err) : nullptr, offs, (const int*)(dp + o_offs), n, nullptr, a,
                            (float*)(ob + o_next));
    if (rc != DR3LK_OK) return rc;
    if (!out) CU_TRY(ctx, cudaMemcpyAsync(hp + o_next, dp + o_next, total - o_next, cudaMemcpyDeviceToHost, st));  // ... one copy back
    CU_TRY(ctx, cudaStreamSynchronize(st));
    memcpy(next_pts, hp + o_next, 8 * (size_t)n);
    memcpy(status, hp + o_status, (size_t)n);
    if (err) memcpy(err, hp + o_err, 4 * (size_t)n);
    return DR3LK_OK;
}

int dr3lk_track_batch_host(dr3lk_ctx* ctx, const uint8_t* prev, const uint8_t* next, int w, int h, size_t step,
                           size_t image_stride, int batch, const float* prev_pts, float* next_pts, uint8_t* status, float* err,
                           const int* pts_offset, uint32_t* stats, int chunk_pairs, int win_w, int win_h, int max_level,
                           int crit_type, int crit_max_count, double crit_eps, int flags, double min_eig_threshold)
{
    if (!ctx) return DR3LK_E_ARG;
    LKArgs a{win_w, win_h, max_level, crit_type, crit_max_count, flags, crit_eps, min_eig_threshold};
    int rc = check_lk_args(ctx, w, h, a);
    if (rc != DR3LK_OK) return rc;
    if (batch < 1 || !prev || !next || step < (size_t)w || image_stride < step * (size_t)h)
        return fail(ctx, DR3LK_E_ARG, "track_batch_host: bad image arguments");
    rc = check_offsets(ctx, pts_offset, batch);
    if (rc != DR3LK_OK) return rc;
    if (pts_offset[batch] == 0) return DR3LK_OK;
    if (!prev_pts || !next_pts || !status) return fail(ctx, DR3LK_E_ARG, "track_batch_host: null point / status buffer");
    cudaSetDevice(ctx->device);
    const int n_slots = dr3lk_ctx::n_slots();
    for (int i = 0; i < n_slots; i++)
        if (!ctx->slot_stream[i]) CU_TRY(ctx, cudaStreamCreateWithFlags(&ctx->slot_stream[i], cudaStreamNonBlocking));
    // Whatever way this function returns, no copy may still be in flight: earlier chunks read `offs_host` (a local) and write
    // the caller's output buffers.  The guard drains the slot streams on every exit, error returns included.
    struct Drain {
        dr3lk_ctx* c;
        ~Drain() { for (int i = 0; i < dr3lk_ctx::kSlots; i++) if (c->slot_stream[i]) cudaStreamSynchronize(c->slot_stream[i]); }
    } drain{ctx};
    // the pipeline streams start after whatever the caller queued on the context stream ...
    cudaEvent_t ev_start;
    CU_TRY(ctx, cudaEventCreateWithFlags(&ev_start, cudaEventDisableTiming));
    CU_TRY(ctx, cudaEventRecord(ev_start, ctx->stream));
    for (int i = 0; i < n_slots; i++) CU_TRY(ctx, cudaStreamWaitEvent(ctx->slot_stream[i], ev_start, 0));
    cudaEventDestroy(ev_start);

    // Chunk boundaries.  chunk_pairs > 0: equal chunks of that many pairs.  0 = choose: ~128 MB of level-0 pixels per
    // chunk (few launches, short tails), at most the batch split in 2 * kSlots pieces, and a geometric ramp-up from
    // small first chunks so that the first kernels start after microseconds of H2D instead of a full chunk's copy.
    std::vector<int> cb;  // chunk c = pairs cb[c] .. cb[c + 1]
    cb.push_back(0);
    if (chunk_pairs <= 0) {
        const size_t per_pair = 2 * (size_t)w * h;
        static const size_t chunk_mb = getenv("DR3LK_CHUNK_MB") ? (size_t)atoi(getenv("DR3LK_CHUNK_MB")) : 128;  // tuning knob
        int full = (int)std::max<size_t>(1, (chunk_mb << 20) / per_pair);
        full = std::min(full, std::max(1, (batch + 2 * n_slots - 1) / (2 * n_slots)));
        int next = std::max(1, full / 16);
        while (cb.back() < batch) {
            cb.push_back(std::min(batch, cb.back() + next));
            next = std::min(full, next * 2);
        }
    } else {
        while (cb.back() < batch) cb.push_back(std::min(batch, cb.back() + chunk_pairs));
    }
    const int n_chunks = (int)cb.size() - 1;
    std::vector<int> offs_host;  // all chunks' rebased offsets; must outlive the async copies
    offs_host.reserve((size_t)batch + n_chunks);
    std::vector<size_t> offs_pos(n_chunks);
    for (int c = 0; c < n_chunks; c++) {
        const int b0 = cb[c], b1 = cb[c + 1];
        offs_pos[c] = offs_host.size();
        for (int b = b0; b <= b1; b++) offs_host.push_back(pts_offset[b] - pts_offset[b0]);
    }
    for (int c = 0; c < n_chunks; c++) {
        const int slot = c % n_slots;
        Workspace& W = ctx->slot_ws[slot];
        cudaStream_t st = ctx->slot_stream[slot];
        const int b0 = cb[c], b1 = cb[c + 1], nb = b1 - b0;
        const int p0 = pts_offset[b0], n = pts_offset[b1] - p0;
        if (n == 0) continue;
        // Level-0 pixels cross PCIe as ONE contiguous copy per frame set in the caller's own layout (2-D copies with an
        // odd row step are DMA-unfriendly); track_batch_device re-pitches on the device when the layout is not 16-B aligned.
        const size_t span = image_stride * (size_t)(nb - 1) + step * (size_t)(h - 1) + w;
        CU_TRY(ctx, W.lvl0_prev.reserve(span + 16));
        CU_TRY(ctx, W.lvl0_next.reserve(span + 16));
        const size_t o_prev = 0, o_next = 8 * (size_t)n, o_err = 16 * (size_t)n, o_stats = 20 * (size_t)n, o_status = 24 * (size_t)n;
        CU_TRY(ctx, W.pts.reserve(25 * (size_t)n + 16));
        CU_TRY(ctx, W.offs.reserve(sizeof(int) * (size_t)(nb + 1)));
        uint8_t* dp = (uint8_t*)W.pts.p;
        CU_TRY(ctx, cudaMemcpyAsync(W.lvl0_prev.p, prev + (size_t)b0 * image_stride, span, cudaMemcpyHostToDevice, st));
        CU_TRY(ctx, cudaMemcpyAsync(W.lvl0_next.p, next + (size_t)b0 * image_stride, span, cudaMemcpyHostToDevice, st));
        CU_TRY(ctx, cudaMemcpyAsync(dp + o_prev, prev_pts + 2 * (size_t)p0, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
        if (flags & DR3LK_USE_INITIAL_FLOW)
            CU_TRY(ctx, cudaMemcpyAsync(dp + o_next, next_pts + 2 * (size_t)p0, 8 * (size_t)n, cudaMemcpyHostToDevice, st));
        CU_TRY(ctx, cudaMemcpyAsync(W.offs.p, offs_host.data() + offs_pos[c], sizeof(int) * (size_t)(nb + 1), cudaMemcpyHostToDevice, st));
        rc = track_batch_device(ctx, W, st, (const uint8_t*)W.lvl0_prev.p, (const uint8_t*)W.lvl0_next.p, w, h, step, image_stride, nb,
                                (const float*)(dp + o_prev), (float*)(dp + o_next), dp + o_status, err ? (float*)(dp + o_err) : nullptr,
                                offs_host.data() + offs_pos[c], (const int*)W.offs.p, n, stats ? (uint32_t*)(dp + o_stats) : nullptr, a);
        if (rc != DR3LK_OK) return rc;
        CU_TRY(ctx, cudaMemcpyAsync(next_pts + 2 * (size_t)p0, dp + o_next, 8 * (size_t)n, cudaMemcpyDeviceToHost, st));
        CU_TRY(ctx, cudaMemcpyAsync(status + p0, dp + o_status, (size_t)n, cudaMemcpyDeviceToHost, st));
        if (err) CU_TRY(ctx, cudaMemcpyAsync(err + p0, dp + o_err, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
        if (stats) CU_TRY(ctx, cudaMemcpyAsync(stats + p0, dp + o_stats, 4 * (size_t)n, cudaMemcpyDeviceToHost, st));
    }
    // ... and the context stream continues after them, so events on it bracket the whole call
    for (int i = 0; i < n_slots; i++) {
        cudaEvent_t ev;
        CU_TRY(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
        CU_TRY(ctx, cudaEventRecord(ev, ctx->slot_stream[i]));
        CU_TRY(ctx, cudaStreamWaitEvent(ctx->stream, ev, 0));
        cudaEventDestroy(ev);
    }
    CU_TRY(ctx, cudaStreamSynchronize(ctx->stream));
    return DR3LK_OK;
}

/* ---------------------------------------------------------------------------------------------- */
/* multi-GPU: one context + one worker thread per device, contiguous blocks of frame pairs          */
/* ---------------------------------------------------------------------------------------------- */

struct dr3lk_multi {
    std::vector<dr3lk_ctx*> ctx;   // one per entry of `devices` (a device may be listed more than once)
    std::string err;
};

int dr3lk_multi_create(dr3lk_multi** out, const int* devices, int n_devices)
{
    if (!out) return DR3LK_E_ARG;
    *out = nullptr;
    if (!devices || n_devices < 1 || n_devices > 64) { g_create_error = "multi_create: 1..64 devices"; return DR3LK_E_ARG; }
    dr3lk_multi* m = new (std::nothrow) dr3lk_multi();
    if (!m) return DR3LK_E_CUDA;
    for (int i = 0; i < n_devices; i++) {
        dr3lk_ctx* c = nullptr;
        const int rc = dr3lk_create(&c, devices[i]);
        if (rc != DR3LK_OK) {
            for (dr3lk_ctx* x : m->ctx) dr3lk_destroy(x);
            delete m;
            return rc;  // g_create_error was set by dr3lk_create
        }
        m->ctx.push_back(c);
    }
    *out = m;
    return DR3LK_OK;
}

void dr3lk_multi_destroy(dr3lk_multi* m)
{
    if (!m) return;
    for (dr3lk_ctx* c : m->ctx) dr3lk_destroy(c);
    delete m;
}

int dr3lk_multi_size(const dr3lk_multi* m) { return m ? (int)m->ctx.size() : 0; }
dr3lk_ctx* dr3lk_multi_context(dr3lk_multi* m, int i) { return (m && i >= 0 && i < (int)m->ctx.size()) ? m->ctx[i] : nullptr; }
const char* dr3lk_multi_last_error(const dr3lk_multi* m) { return m ? m->err.c_str() : g_create_error.c_str(); }

void dr3lk_shard_range(int n_pairs, int rank, int world, int* lo, int* hi)
{
    // pair p lives on rank floor(p * world / n_pairs): contiguous blocks, the same rule as 3dr_b200/sharding.py
    *lo = (int)(((long long)rank * n_pairs) / world);
    *hi = (int)(((long long)(rank + 1) * n_pairs) / world);
}

int dr3lk_multi_track_batch_host(dr3lk_multi* m, const uint8_t* prev, const uint8_t* next, int w, int h, size_t step, size_t image_stride,
                                 int batch, const float* prev_pts, float* next_pts, uint8_t* status, float* err, const int* pts_offset,
                                 uint32_t* stats, int chunk_pairs, int win_w, int win_h, int max_level, int crit_type, int crit_max_count,
                                 double crit_eps, int flags, double min_eig_threshold)
{
    if (!m) return DR3LK_E_ARG;
    const int G = (int)m->ctx.size();
    // argument errors are reported once, by the first context, before any thread starts
    if (batch < 1 || !pts_offset) { m->err = "multi_track_batch_host: bad batch / offsets"; return DR3LK_E_ARG; }
    {
        int rc = check_offsets(m->ctx[0], pts_offset, batch);
        if (rc != DR3LK_OK) { m->err = m->ctx[0]->err; return rc; }
    }
    std::vector<int> rcs(G, DR3LK_OK);
    std::vector<std::vector<int>> offs(G);
    std::vector<std::thread> workers;
    workers.reserve(G);
    for (int r = 0; r < G; r++) {
        int lo, hi;
        dr3lk_shard_range(batch, r, G, &lo, &hi);
        if (hi <= lo) continue;
        const int p0 = pts_offset[lo];
        offs[r].resize(hi - lo + 1);
        for (int b = lo; b <= hi; b++) offs[r][b - lo] = pts_offset[b] - p0;
        // every rank owns disjoint slices of the caller's arrays: no exchange, no lock (the "final gather" of SURVEY.md 8e is
        // each device's own D2H copy into its slice)
        auto work = [=, &rcs, &offs]() {
            rcs[r] = dr3lk_track_batch_host(m->ctx[r], prev + (size_t)lo * image_stride, next + (size_t)lo * image_stride, w, h, step,
                                            image_stride, hi - lo, prev_pts + 2 * (size_t)p0, next_pts + 2 * (size_t)p0, status + p0,
                                            err ? err + p0 : nullptr, offs[r].data(), stats ? stats + p0 : nullptr, chunk_pairs, win_w,
                                            win_h, max_level, crit_type, crit_max_count, crit_eps, flags, min_eig_threshold);
        };
        try {
            workers.emplace_back(work);
        } catch (...) {  // no thread to be had (std::system_error must not cross the C boundary): this rank runs on the calling thread
            work();
        }
    }
    for (auto& t : workers) t.join();
    for (int r = 0; r < G; r++)
        if (rcs[r] != DR3LK_OK) {
            m->err = "device " + std::to_string(m->ctx[r]->device) + " (rank " + std::to_string(r) + "): " + m->ctx[r]->err;
            return rcs[r];
        }
    return DR3LK_OK;
}

int dr3lk_build_lk_pyramid(dr3lk_ctx* ctx, const uint8_t* img, int w, int h, size_t step, int win_w, int win_h, int max_level,
                           uint8_t* const* out_levels, int16_t* const* out_derivs, int* eff_max_level)
{
    if (!ctx) return DR3LK_E_ARG;
    LKArgs a{win_w, win_h, max_level, 0, 0, 0, 0., 0.};
    int rc = check_lk_args(ctx, w, h, a);
    if (rc != DR3LK_OK) return rc;
    if (!img || step < (size_t)w || !out_levels) return fail(ctx, DR3LK_E_ARG, "build_lk_pyramid: bad argument");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    Workspace& W = ctx->ws;
    const int pitch0 = align_up(w, 16);
    const size_t img_bytes = (size_t)pitch0 * h;
    CU_TRY(ctx, W.lvl0_prev.reserve(img_bytes));
    CU_TRY(ctx, cudaMemcpy2DAsync(W.lvl0_prev.p, pitch0, img, step, w, h, cudaMemcpyHostToDevice, st));
    LKParams lk;
    memset(&lk, 0, sizeof(lk));
    PyrLayout P = make_layout(w, h, win_w, win_h, max_level);
    // previous-frame side only (next0 == nullptr): Gaussian levels + derivatives
    rc = build_pyramids(ctx, W, st, (const uint8_t*)W.lvl0_prev.p, nullptr, pitch0, img_bytes, 1, P, win_w, win_h, lk);
    if (rc != DR3LK_OK) return rc;
    for (int l = 0; l <= P.ml; l++) {
        const LevelDesc& d = lk.lv[l];
        CU_TRY(ctx, cudaMemcpy2DAsync(out_levels[l], d.w, d.prev, d.pitch_p, d.w, d.h, cudaMemcpyDeviceToHost, st));
        if (out_derivs)
            CU_TRY(ctx, cudaMemcpy2DAsync(out_derivs[l], (size_t)d.w * 4, d.deriv, (size_t)d.dpitch * 4, (size_t)d.w * 4, d.h,
                                          cudaMemcpyDeviceToHost, st));
    }
    CU_TRY(ctx, cudaStreamSynchronize(st));
    if (eff_max_level) *eff_max_level = P.ml;
    return DR3LK_OK;
}


/* ---------------------------------------------------------------------------------------------- */
/* f-2: cached pyramids                                                                            */
/* ---------------------------------------------------------------------------------------------- */

struct dr3lk_pyramid {
    dr3lk_ctx* ctx;
    int w, h, win_w, win_h;
    PyrLayout P;
    DevBuf img, deriv;
    bool has_deriv = true;     // false: Gaussian levels only (usable as the `next` side of a call only)
    size_t der_total = 0;      // ints in `deriv`
    LevelDesc lv[kMaxLevels];  // prev / deriv fields describe this frame
};

// Allocates the device buffers of a pyramid object (from the context's pool when possible) and fills its level
// descriptors.  with_deriv = false: Gaussian levels only (a frame that is only ever the `next` side of a call).
static int pyramid_alloc(dr3lk_ctx* ctx, int w, int h, int win_w, int win_h, int max_level, bool with_deriv, dr3lk_pyramid** out)
{
    dr3lk_pyramid* p = new (std::nothrow) dr3lk_pyramid();
    if (!p) return fail(ctx, DR3LK_E_CUDA, "out of host memory");
    p->ctx = ctx; p->w = w; p->h = h; p->win_w = win_w; p->win_h = win_h; p->has_deriv = with_deriv;
    p->P = make_layout(w, h, win_w, win_h, max_level);
    const PyrLayout& P = p->P;
    if (P.img_bytes[0] >= (1ull << 31) || P.der_ints[0] >= (1ull << 29)) { delete p; return fail(ctx, DR3LK_E_SIZE, "images of 2 GiB or more (derivatives included) are not supported"); }
    size_t img_total = 0, der_total = 0, ioff[kMaxLevels], doff[kMaxLevels];
    for (int l = 0; l <= P.ml; l++) { ioff[l] = img_total; img_total += P.img_bytes[l]; doff[l] = der_total; der_total += P.der_ints[l]; }
    p->der_total = der_total;
    p->img = ctx->take(img_total);
    p->img.zero_sig = 0;  // whatever the pooled buffer held before, it is an image buffer now (its aprons will not be zeros)
    cudaError_t e = p->img.reserve(img_total);
    if (e == cudaSuccess && with_deriv) {
        p->deriv = ctx->take(der_total * sizeof(int));
        e = p->deriv.reserve(der_total * sizeof(int));
    }
    if (e != cudaSuccess) { p->img.release(); p->deriv.release(); delete p; return fail_cuda(ctx, e, "pyramid: allocation"); }
    for (int l = 0; l <= P.ml; l++) {
        LevelDesc& d = p->lv[l];
        d.w = P.w[l]; d.h = P.h[l];
        d.prev = d.next = (const uint8_t*)p->img.p + ioff[l] + P.img_org[l];
        d.pitch_p = d.pitch_n = P.pitch[l];
        d.prev_stride = d.next_stride = (unsigned)P.img_bytes[l];
        d.deriv = with_deriv ? (const int*)p->deriv.p + doff[l] + P.der_org[l] : nullptr;
        d.dpitch = P.dpitch[l];
        d.deriv_stride = (unsigned)P.der_ints[l];
    }
    ctx->live.push_back(p);
    *out = p;
    return DR3LK_OK;
}

// dr3lk_destroy: pyramid objects that outlive their context keep a valid handle (so that dr3lk_pyramid_destroy stays legal in
// any order, e.g. from a garbage collector or a static destructor) but lose their device memory and their context.
static void orphan_pyramids(dr3lk_ctx* ctx)
{
    for (dr3lk_pyramid* p : ctx->live) {
        p->img.release(); p->deriv.release();
        p->ctx = nullptr;
    }
    ctx->live.clear();
}

// Enqueues the build of a pyramid object from a level-0 image that is already on the device (rows `pitch0` bytes apart, any
// alignment): apron copy (or plain copy for windows without aprons), Gaussian levels, derivatives.  No sync.
static void pyramid_enqueue(dr3lk_ctx* ctx, dr3lk_pyramid* p, const uint8_t* l0_dev, size_t pitch0, cudaStream_t st, Launch& L)
{
    const PyrLayout& P = p->P;
    const size_t l0_bytes = pitch0 * p->h;
    if (L.err != cudaSuccess) return;
    const bool pdl_before = L.pdl;
    L.pdl = pdl_enabled();  // one frame: a chain of small dependent kernels, see Launch::pdl
    if (P.ax > 0) {
        // level 0 is copied into its apron-carrying image; the derivative aprons are zeros: cleared once per (buffer, layout)
        if (p->has_deriv) {
            const unsigned long long sig = 1ull | ((unsigned long long)p->w << 1) ^ ((unsigned long long)p->h << 21) ^ ((unsigned long long)p->win_w << 41) ^
                                           ((unsigned long long)p->win_h << 49) ^ ((unsigned long long)P.ml << 57);
            if (p->deriv.zero_sig != sig) {
                L.err = cudaMemsetAsync(p->deriv.p, 0, p->der_total * sizeof(int), st);
                p->deriv.zero_sig = sig;
            }
        }
        launch_pad_level0(L, l0_dev, nullptr, pitch0, l0_bytes, const_cast<uint8_t*>(p->lv[0].prev), nullptr, P.pitch[0], P.img_bytes[0], p->w,
                          p->h, P.ax, P.ay, 1, 0);
    } else {
        L.err = cudaMemcpy2DAsync(p->img.p, P.pitch[0], l0_dev, pitch0, p->w, p->h, cudaMemcpyDeviceToDevice, st);
    }
    for (int l = 0; l <= P.ml; l++) {
        const LevelDesc& s = p->lv[l];
        PyrLevelArgs pa{};
        pa.prev_src = s.prev; pa.prev_src_stride = s.prev_stride;
        pa.w = s.w; pa.h = s.h; pa.src_pitch = s.pitch_p;
        pa.deriv = const_cast<int*>(s.deriv); pa.dpitch = s.dpitch; pa.deriv_stride = s.deriv_stride;
        pa.n_prev = 1; pa.n_next = 0;
        pa.down = l < P.ml;
        if (!pa.down && !pa.deriv) break;  // nothing to produce from the last level
        pa.dst_apron_x = P.ax; pa.dst_apron_y = P.ay;
        pa.src_apron_x = P.ax; pa.src_apron_y = P.ay;
        if (pa.down) { pa.prev_dst = const_cast<uint8_t*>(p->lv[l + 1].prev); pa.prev_dst_stride = p->lv[l + 1].prev_stride; pa.dst_pitch = p->lv[l + 1].pitch_p; }
        launch_pyr_level(L, pa);
    }
    L.pdl = pdl_before;
    ctx->launches += L.launches;
    L.launches = 0;
}

static void pyramid_free(dr3lk_pyramid* p, bool to_pool)
{
    dr3lk_ctx* ctx = p->ctx;
    if (!ctx) { delete p; return; }  // orphaned by dr3lk_destroy: the device buffers are gone already
    ctx->live.erase(std::remove(ctx->live.begin(), ctx->live.end(), p), ctx->live.end());
    if (to_pool && ctx->pool.size() < 16) {
        if (p->img.p) ctx->pool.push_back(p->img);
        if (p->deriv.p) ctx->pool.push_back(p->deriv);
    } else {
        p->img.release(); p->deriv.release();
    }
    delete p;
}

int dr3lk_pyramid_create(dr3lk_ctx* ctx, const uint8_t* img, int w, int h, size_t step, int win_w, int win_h, int max_level,
                         dr3lk_pyramid** out)
{
    if (!ctx || !out) return DR3LK_E_ARG;
    *out = nullptr;
    LKArgs a{win_w, win_h, max_level, 0, 0, 0, 0., 0.};
    int rc = check_lk_args(ctx, w, h, a);
    if (rc != DR3LK_OK) return rc;
    if (!img || step < (size_t)w) return fail(ctx, DR3LK_E_ARG, "pyramid_create: bad image arguments");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    dr3lk_pyramid* p = nullptr;
    rc = pyramid_alloc(ctx, w, h, win_w, win_h, max_level, true, &p);
    if (rc != DR3LK_OK) return rc;
    const int pitch0 = align_up(w, 16);
    const size_t l0_bytes = (size_t)pitch0 * h;
    cudaError_t e = ctx->pinned.reserve(l0_bytes);
    if (e == cudaSuccess) e = ctx->ws.lvl0_prev.reserve(l0_bytes);
    if (e != cudaSuccess) { pyramid_free(p, false); return fail_cuda(ctx, e, "pyramid_create: allocation"); }
    uint8_t* hp = (uint8_t*)ctx->pinned.p;
    Launch L{st, cudaSuccess, 0};
    size_t pitch = 0;
    L.err = upload_image((uint8_t*)ctx->ws.lvl0_prev.p, hp, pitch0, img, step, w, h, st, &pitch);
    pyramid_enqueue(ctx, p, (const uint8_t*)ctx->ws.lvl0_prev.p, pitch, st, L);
    // the pinned staging buffer is reused by the next call: wait for the upload
    if (L.err == cudaSuccess) L.err = cudaStreamSynchronize(st);
    if (L.err != cudaSuccess) { pyramid_free(p, false); return fail_cuda(ctx, L.err, "pyramid_create"); }
    *out = p;
    return DR3LK_OK;
}

void dr3lk_pyramid_destroy(dr3lk_pyramid* pyr)
{
    if (!pyr) return;
    if (pyr->ctx) cudaSetDevice(pyr->ctx->device);
    // stream-ordered reuse is safe: every consumer of these buffers was enqueued on the context's stream before this point
    pyramid_free(pyr, true);
}

int dr3lk_pyramid_levels(const dr3lk_pyramid* pyr) { return pyr ? pyr->P.ml + 1 : 0; }

int dr3lk_calc_optical_flow_pyr_lk_cached(dr3lk_ctx* ctx, const dr3lk_pyramid* prev, const dr3lk_pyramid* next, const float* prev_pts,
                                          float* next_pts, uint8_t* status, float* err, int n, int win_w, int win_h, int max_level,
                                          int crit_type, int crit_max_count, double crit_eps, int flags, double min_eig_threshold)
{
    if (!ctx || !prev || !next) return DR3LK_E_ARG;
    LKArgs a{win_w, win_h, max_level, crit_type, crit_max_count, flags, crit_eps, min_eig_threshold};
    int rc = check_lk_args(ctx, prev->w, prev->h, a);
    if (rc != DR3LK_OK) return rc;
    if (prev->ctx != ctx || next->ctx != ctx) return fail(ctx, DR3LK_E_ARG, "pyramids belong to another context");
    if (prev->w != next->w || prev->h != next->h)
        return fail(ctx, DR3LK_E_SIZE, "(-215:Assertion failed) prevPyr[level * lvlStep1].size() == nextPyr[level * lvlStep2].size()");
    if (prev->win_w != win_w || prev->win_h != win_h || next->win_w != win_w || next->win_h != win_h)
        return fail(ctx, DR3LK_E_ARG, "pyramids were built for another window size");
    if (!prev->has_deriv) return fail(ctx, DR3LK_E_ARG, "the previous-frame pyramid was built without derivatives");
    if (n < 0) return fail(ctx, DR3LK_E_ARG, "negative point count");
    if (n == 0) return DR3LK_OK;
    if (!prev_pts || !next_pts || !status) return fail(ctx, DR3LK_E_ARG, "null point / status buffer");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    Workspace& W = ctx->ws;
    LKParams lk;
    memset(&lk, 0, sizeof(lk));
    const int ml = std::min(std::min(prev->P.ml, next->P.ml), max_level);
    for (int l = 0; l <= ml; l++) {
        lk.lv[l] = prev->lv[l];
        lk.lv[l].next = next->lv[l].prev;
        lk.lv[l].pitch_n = next->lv[l].pitch_p;
        lk.lv[l].next_stride = next->lv[l].prev_stride;
    }
    lk.max_level = ml;
    lk.fast_ok = prev->P.ax > 0;  // pyramid objects carry the aprons whenever the window has a specialised kernel
    const size_t n8 = align_up_sz(8 * (size_t)n, 16);
    const size_t o_prev = 0, o_offs = n8, o_next = o_offs + 16, o_err = o_next + n8, o_status = o_err + align_up_sz(4 * (size_t)n, 16);
    const size_t total = o_status + align_up_sz((size_t)n, 16);
    CU_TRY(ctx, W.pts.reserve(total));
    CU_TRY(ctx, ctx->pinned.reserve(total));
    uint8_t* dp = (uint8_t*)W.pts.p;
    uint8_t* hp = (uint8_t*)ctx->pinned.p;
    memcpy(hp + o_prev, prev_pts, 8 * (size_t)n);
    const int offs[2] = {0, n};
    memcpy(hp + o_offs, offs, sizeof(offs));
    size_t in_bytes = o_next;
    if (flags & DR3LK_USE_INITIAL_FLOW) { memcpy(hp + o_next, next_pts, 8 * (size_t)n); in_bytes = o_next + 8 * (size_t)n; }
    CU_TRY(ctx, cudaMemcpyAsync(dp, hp, in_bytes, cudaMemcpyHostToDevice, st));
    uint8_t* const out = mapped_alias(hp, n);
    uint8_t* const ob = out ? out : dp;
    rc = run_tracking(ctx, W, st, lk, 1, (const float*)(dp + o_prev), (float*)(dp + o_next), ob + o_status,
                      err ? (float*)(ob + o_err) : nullptr, offs, (const int*)(dp + o_offs), n, nullptr, a, (float*)(ob + o_next));
    if (rc != DR3LK_OK) return rc;
    if (!out) CU_TRY(ctx, cudaMemcpyAsync(hp + o_next, dp + o_next, total - o_next, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    memcpy(next_pts, hp + o_next, 8 * (size_t)n);
    memcpy(status, hp + o_status, (size_t)n);
    if (err) memcpy(err, hp + o_err, 4 * (size_t)n);
    return DR3LK_OK;
}

int dr3lk_track_frame(dr3lk_ctx* ctx, const dr3lk_pyramid* prev, const uint8_t* next_img, size_t next_step, const float* prev_pts,
                      float* next_pts, uint8_t* status, float* err, int n, int win_w, int win_h, int max_level, int crit_type,
                      int crit_max_count, double crit_eps, int flags, double min_eig_threshold, int keep_next, dr3lk_pyramid** next_out)
{
    if (!ctx || !prev) return DR3LK_E_ARG;
    if (next_out) *next_out = nullptr;
    LKArgs a{win_w, win_h, max_level, crit_type, crit_max_count, flags, crit_eps, min_eig_threshold};
    const int w = prev->w, h = prev->h;
    int rc = check_lk_args(ctx, w, h, a);
    if (rc != DR3LK_OK) return rc;
    if (prev->ctx != ctx) return fail(ctx, DR3LK_E_ARG, "the pyramid belongs to another context");
    if (prev->win_w != win_w || prev->win_h != win_h) return fail(ctx, DR3LK_E_ARG, "the pyramid was built for another window size");
    if (!prev->has_deriv) return fail(ctx, DR3LK_E_ARG, "the previous-frame pyramid was built without derivatives");
    if (!next_img || next_step < (size_t)w) return fail(ctx, DR3LK_E_ARG, "track_frame: bad image arguments");
    if (keep_next < 0 || keep_next > 2 || (keep_next && !next_out)) return fail(ctx, DR3LK_E_ARG, "track_frame: keep_next must be 0..2 and needs next_out");
    if (n < 0) return fail(ctx, DR3LK_E_ARG, "negative point count");
    if (n > 0 && (!prev_pts || !next_pts || !status)) return fail(ctx, DR3LK_E_ARG, "null point / status buffer");
    if (n == 0 && !keep_next) return DR3LK_OK;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    Workspace& W = ctx->ws;
    dr3lk_pyramid* p2 = nullptr;
    rc = pyramid_alloc(ctx, w, h, win_w, win_h, max_level, keep_next == 2, &p2);
    if (rc != DR3LK_OK) return rc;
    // One device block and one pinned mirror of it: [new image][prev_pts 8n][offsets 16][next_pts 8n][err 4n][status n]:
    // the image and the points cross PCIe as ONE copy, the results come back as ONE copy, the call synchronises once.
    const int pitch0 = align_up(w, 16);
    const size_t img_block = align_up_sz((size_t)pitch0 * h, 256);
    const size_t n8 = align_up_sz(8 * (size_t)n, 16);
    const size_t o_prev = img_block, o_offs = o_prev + n8, o_next = o_offs + 16, o_err = o_next + n8, o_status = o_err + align_up_sz(4 * (size_t)n, 16);
    const size_t total = o_status + align_up_sz((size_t)n, 16);
    cudaError_t e = W.lvl0_prev.reserve(total);
    if (e == cudaSuccess) e = ctx->pinned.reserve(total);
    if (e != cudaSuccess) { pyramid_free(p2, false); return fail_cuda(ctx, e, "track_frame: allocation"); }
    uint8_t* dp = (uint8_t*)W.lvl0_prev.p;
    uint8_t* hp = (uint8_t*)ctx->pinned.p;
    // a pinned image (continuous, or at the aligned pitch) goes to the copy engine as it is; any other is packed in front of the points
    const size_t pitch_direct = direct_pitch(next_img, next_step, w, pitch0);
    const bool img_direct = pitch_direct != 0;
    Launch L{st, cudaSuccess, 0};
    if (img_direct) L.err = cudaMemcpyAsync(dp, next_img, direct_bytes(next_step, w, h), cudaMemcpyHostToDevice, st);
    else for (int y = 0; y < h; y++) memcpy(hp + (size_t)y * pitch0, next_img + (size_t)y * next_step, (size_t)w);
    const int offs[2] = {0, n};
    size_t in_bytes = img_block;
    if (n > 0) {
        memcpy(hp + o_prev, prev_pts, 8 * (size_t)n);
        memcpy(hp + o_offs, offs, sizeof(offs));
        in_bytes = o_next;
        if (flags & DR3LK_USE_INITIAL_FLOW) { memcpy(hp + o_next, next_pts, 8 * (size_t)n); in_bytes = o_next + 8 * (size_t)n; }
    }
    const size_t in_from = img_direct ? o_prev : 0;
    uint8_t* const out = n > 0 ? mapped_alias(hp, n) : nullptr;  // results (and, with a pinned image, the points) through the mapped mirror
    const bool mapped_pts = img_direct && out && mapped_points_enabled();                    // no second copy in front of the kernels for a few hundred bytes
    if (L.err == cudaSuccess && in_bytes > in_from && !mapped_pts)
        L.err = cudaMemcpyAsync(dp + in_from, hp + in_from, in_bytes - in_from, cudaMemcpyHostToDevice, st);
    pyramid_enqueue(ctx, p2, dp, img_direct ? pitch_direct : (size_t)pitch0, st, L);
    if (L.err != cudaSuccess) { pyramid_free(p2, false); return fail_cuda(ctx, L.err, "track_frame: pyramid of the new frame"); }
    if (n > 0) {
        LKParams lk;
        memset(&lk, 0, sizeof(lk));
        const int ml = std::min(std::min(prev->P.ml, p2->P.ml), max_level);
        for (int l = 0; l <= ml; l++) {
            lk.lv[l] = prev->lv[l];
            lk.lv[l].next = p2->lv[l].prev;
            lk.lv[l].pitch_n = p2->lv[l].pitch_p;
            lk.lv[l].next_stride = p2->lv[l].prev_stride;
        }
        lk.max_level = ml;
        lk.fast_ok = prev->P.ax > 0;
        uint8_t* const ob = out ? out : dp;
        const uint8_t* const pb = mapped_pts ? out : dp;
        rc = run_tracking(ctx, W, st, lk, 1, (const float*)(pb + o_prev), (float*)(pb + o_next), ob + o_status,
                          err ? (float*)(ob + o_err) : nullptr, offs, (const int*)(dp + o_offs), n, nullptr, a, (float*)(ob + o_next));
        if (rc != DR3LK_OK) { cudaStreamSynchronize(st); pyramid_free(p2, false); return rc; }
        e = out ? cudaSuccess : cudaMemcpyAsync(hp + o_next, dp + o_next, total - o_next, cudaMemcpyDeviceToHost, st);
        if (e != cudaSuccess) { cudaStreamSynchronize(st); pyramid_free(p2, false); return fail_cuda(ctx, e, "track_frame: D2H"); }
    }
    e = cudaStreamSynchronize(st);
    if (e != cudaSuccess) { pyramid_free(p2, false); return fail_cuda(ctx, e, "track_frame"); }
    if (n > 0) {
        memcpy(next_pts, hp + o_next, 8 * (size_t)n);
        memcpy(status, hp + o_status, (size_t)n);
        if (err) memcpy(err, hp + o_err, 4 * (size_t)n);
    }
    if (keep_next) *next_out = p2; else pyramid_free(p2, true);
    return DR3LK_OK;
}

/* ---------------------------------------------------------------------------------------------- */
/* f-3: status filter + disparity + bearing vectors                                                */
/* ---------------------------------------------------------------------------------------------- */

int dr3lk_filter_tracks(dr3lk_ctx* ctx, const float* ref_pts, const float* cur_pts, const uint8_t* status, int n, double fx, double fy,
                        double cx, double cy, const double* distortion, float* out_ref, float* out_cur, double* out_disparity,
                        double* out_bearing, int* n_kept)
{
    if (!ctx) return DR3LK_E_ARG;
    if (n < 0 || !n_kept) return fail(ctx, DR3LK_E_ARG, "filter_tracks: bad argument");
    *n_kept = 0;
    if (n == 0) return DR3LK_OK;
    if (!ref_pts || !cur_pts || !status || !out_ref || !out_cur || !out_disparity) return fail(ctx, DR3LK_E_ARG, "filter_tracks: null buffer");
    if (out_bearing && (fx == 0.0 || fy == 0.0)) return fail(ctx, DR3LK_E_ARG, "filter_tracks: zero focal length");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    Workspace& W = ctx->ws;
    // [ref 8n][cur 8n][status n (pad 16)] in; [out_ref 8n][out_cur 8n][disp 8n][bearing 24n][count 16] out
    const size_t n8 = align_up_sz(8 * (size_t)n, 16), n1 = align_up_sz((size_t)n, 16), n24 = align_up_sz(24 * (size_t)n, 16);
    const size_t i_ref = 0, i_cur = n8, i_st = 2 * n8, o_ref = i_st + n1, o_cur = o_ref + n8, o_disp = o_cur + n8, o_bear = o_disp + n8,
                 o_cnt = o_bear + n24, total = o_cnt + 16;
    CU_TRY(ctx, W.pts.reserve(total));
    CU_TRY(ctx, ctx->pinned.reserve(total));
    uint8_t* dp = (uint8_t*)W.pts.p;
    uint8_t* hp = (uint8_t*)ctx->pinned.p;
    memcpy(hp + i_ref, ref_pts, 8 * (size_t)n);
    memcpy(hp + i_cur, cur_pts, 8 * (size_t)n);
    memcpy(hp + i_st, status, (size_t)n);
    CU_TRY(ctx, cudaMemcpyAsync(dp, hp, o_ref, cudaMemcpyHostToDevice, st));
    Launch L{st, cudaSuccess, 0};
    launch_filter_tracks(L, (const float*)(dp + i_ref), (const float*)(dp + i_cur), dp + i_st, n, fx, fy, cx, cy, distortion, (float*)(dp + o_ref),
                         (float*)(dp + o_cur), (double*)(dp + o_disp), out_bearing ? (double*)(dp + o_bear) : nullptr, (int*)(dp + o_cnt));
    ctx->launches += L.launches;
    if (L.err != cudaSuccess) return fail_cuda(ctx, L.err, "filter kernel launch");
    CU_TRY(ctx, cudaMemcpyAsync(hp + o_ref, dp + o_ref, total - o_ref, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    const int k = *reinterpret_cast<const int*>(hp + o_cnt);
    *n_kept = k;
    memcpy(out_ref, hp + o_ref, 8 * (size_t)k);
    memcpy(out_cur, hp + o_cur, 8 * (size_t)k);
    memcpy(out_disparity, hp + o_disp, 8 * (size_t)k);
    if (out_bearing) memcpy(out_bearing, hp + o_bear, 24 * (size_t)k);
    return DR3LK_OK;
}


/* ---------------------------------------------------------------------------------------------- */
/* f-1: FAST-10 + grid Shi-Tomasi selection                                                        */
/* ---------------------------------------------------------------------------------------------- */

// FastDetector::detect on one upload of the image; lk_win_w > 0: the LK pyramid (Gaussian levels + Scharr derivatives) of
// the same device copy is built behind it and returned in *lk_pyr (dr3lk_init_first_frame).
static int fast_detect_impl(dr3lk_ctx* ctx, const uint8_t* img, int w, int h, size_t step, int n_levels, int cell_size, int fast_threshold,
                            double detection_threshold, int box_mode, const uint8_t* occupancy, int* out_xy, int* out_level, float* out_score,
                            int* n_out, int lk_win_w, int lk_win_h, int lk_max_level, dr3lk_pyramid** lk_pyr)
{
    if (!ctx) return DR3LK_E_ARG;
    if (!img || !out_xy || !out_level || !out_score || !n_out || step < (size_t)w) return fail(ctx, DR3LK_E_ARG, "fast_detect: bad argument");
    if (n_levels < 1 || n_levels > 8 || cell_size < 1 || fast_threshold < 0 || fast_threshold > 254)
        return fail(ctx, DR3LK_E_ARG, "fast_detect: n_levels must be 1..8, cell_size >= 1, threshold 0..254");
    if (box_mode < DR3LK_BOX_AUTO_X86 || box_mode > DR3LK_BOX_SSE2) return fail(ctx, DR3LK_E_ARG, "fast_detect: bad box mode");
    if (w < 7 || h < 7 || (long long)w * h >= (1ll << 28)) return fail(ctx, DR3LK_E_SIZE, "fast_detect: image must be 7x7 .. 2^28 pixels");
    *n_out = 0;
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    Workspace& W = ctx->ws;
    // box pyramid shapes (validated like dr3lk_box_pyramid) and the packed device layout
    int lw[8], lh[8], sse2[8] = {0};
    size_t off[8], total = 0;
    {
        long long row = (long long)step;
        bool aligned = (reinterpret_cast<uintptr_t>(img) & 0xF) == 0;
        lw[0] = w; lh[0] = h;
        for (int l = 0; l < n_levels; l++) {
            off[l] = total;
            total += align_up_sz((size_t)lw[l] * lh[l], 256);
            if (l + 1 < n_levels) {
                int rc = box_level_mode(ctx, lw[l], lh[l], l == 0 ? row : (long long)lw[l], box_mode, l == 0 ? aligned : true, &sse2[l + 1]);
                if (rc != DR3LK_OK) return rc;
                lw[l + 1] = lw[l] / 2; lh[l + 1] = lh[l] / 2;
            }
        }
    }
    const int gc = (w + cell_size - 1) / cell_size, gr = (h + cell_size - 1) / cell_size, ncell = gc * gr;
    // device block: [levels][score maps][cell keys 8*ncell][occupancy ncell][level widths 32][out xy 8n][level 4n][score 4n][count 16]
    const size_t o_score = total, o_keys = align_up_sz(2 * total, 256), o_occ = o_keys + align_up_sz(8 * (size_t)ncell, 16),
                 o_lw = o_occ + align_up_sz((size_t)ncell, 16), o_xy = o_lw + 32, o_lv = o_xy + align_up_sz(8 * (size_t)ncell, 16),
                 o_sc = o_lv + align_up_sz(4 * (size_t)ncell, 16), o_cnt = o_sc + align_up_sz(4 * (size_t)ncell, 16), dev_total = o_cnt + 16;
    CU_TRY(ctx, W.lvl0_next.reserve(dev_total));
    uint8_t* dp = (uint8_t*)W.lvl0_next.p;
    // pinned mirror: level 0 rows packed to w bytes, occupancy, level widths; outputs come back into the same mirror
    const size_t h_img = align_up_sz((size_t)w * h, 16), h_total = h_img + align_up_sz((size_t)ncell, 16) + 32 + (dev_total - o_xy);
    CU_TRY(ctx, ctx->pinned.reserve(h_total));
    uint8_t* hp = (uint8_t*)ctx->pinned.p;
    // the Frame's image is a continuous cv::Mat; the box-pyramid walk of a strided ROI is garbage in the reference itself
    if (step != (size_t)w) return fail(ctx, DR3LK_E_UNSUPPORTED, "fast_detect: continuous images only (step == width)");
    memcpy(hp, img, (size_t)w * h);
    uint8_t* h_occ = hp + h_img;
    if (occupancy) memcpy(h_occ, occupancy, (size_t)ncell);
    int* h_lw = (int*)(h_occ + align_up_sz((size_t)ncell, 16));
    for (int l = 0; l < 8; l++) h_lw[l] = l < n_levels ? lw[l] : 1;
    CU_TRY(ctx, cudaMemcpyAsync(dp + off[0], hp, (size_t)w * h, cudaMemcpyHostToDevice, st));
    if (occupancy) CU_TRY(ctx, cudaMemcpyAsync(dp + o_occ, h_occ, (size_t)ncell, cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaMemcpyAsync(dp + o_lw, h_lw, 32, cudaMemcpyHostToDevice, st));
    CU_TRY(ctx, cudaMemsetAsync(dp + o_keys, 0, 8 * (size_t)ncell, st));
    Launch L{st, cudaSuccess, 0};
    for (int l = 1; l < n_levels; l++)
        launch_box_half(L, dp + off[l - 1], lw[l - 1], lh[l - 1], lw[l - 1], 0, dp + off[l], 0, 1, sse2[l]);
    for (int l = 0; l < n_levels; l++)
        if (lw[l] >= 7 && lh[l] >= 7)
            launch_fast_level(L, dp + off[l], dp + o_score + off[l], lw[l], lh[l], l, fast_threshold, cell_size, gc, (float)detection_threshold,
                              detection_threshold, occupancy ? dp + o_occ : nullptr, (unsigned long long*)(dp + o_keys), ctx->fast_arc);
    launch_fast_gather(L, (const unsigned long long*)(dp + o_keys), ncell, (const int*)(dp + o_lw), (int*)(dp + o_xy), (int*)(dp + o_lv),
                       (float*)(dp + o_sc), (int*)(dp + o_cnt));
    ctx->launches += L.launches;
    L.launches = 0;
    if (L.err != cudaSuccess) return fail_cuda(ctx, L.err, "fast_detect kernel launch");
    dr3lk_pyramid* pyr = nullptr;
    if (lk_pyr) {
        int rc = pyramid_alloc(ctx, w, h, lk_win_w, lk_win_h, lk_max_level, true, &pyr);
        if (rc != DR3LK_OK) { cudaStreamSynchronize(st); return rc; }
        pyramid_enqueue(ctx, pyr, dp + off[0], (size_t)w, st, L);  // the level-0 copy FAST just read: no second upload
        if (L.err != cudaSuccess) { cudaStreamSynchronize(st); pyramid_free(pyr, false); return fail_cuda(ctx, L.err, "init_first_frame: LK pyramid"); }
    }
    uint8_t* h_out = (uint8_t*)h_lw + 32;
    cudaError_t ce = cudaMemcpyAsync(h_out, dp + o_xy, dev_total - o_xy, cudaMemcpyDeviceToHost, st);
    if (ce == cudaSuccess) ce = cudaStreamSynchronize(st);
    if (ce != cudaSuccess) { if (pyr) pyramid_free(pyr, false); return fail_cuda(ctx, ce, "fast_detect: download"); }
    if (lk_pyr) *lk_pyr = pyr;
    const int n = *reinterpret_cast<const int*>(h_out + (o_cnt - o_xy));
    *n_out = n;
    memcpy(out_xy, h_out, 8 * (size_t)n);
    memcpy(out_level, h_out + (o_lv - o_xy), 4 * (size_t)n);
    memcpy(out_score, h_out + (o_sc - o_xy), 4 * (size_t)n);
    return DR3LK_OK;
}

int dr3lk_fast_detect(dr3lk_ctx* ctx, const uint8_t* img, int w, int h, size_t step, int n_levels, int cell_size, int fast_threshold,
                      double detection_threshold, int box_mode, const uint8_t* occupancy, int* out_xy, int* out_level, float* out_score,
                      int* n_out)
{
    return fast_detect_impl(ctx, img, w, h, step, n_levels, cell_size, fast_threshold, detection_threshold, box_mode, occupancy, out_xy,
                            out_level, out_score, n_out, 0, 0, 0, nullptr);
}

/* ---------------------------------------------------------------------------------------------- */
/* the two-frame initialiser front end, device-resident between its steps                          */
/* ---------------------------------------------------------------------------------------------- */

int dr3lk_init_first_frame(dr3lk_ctx* ctx, const uint8_t* img, int w, int h, size_t step, int n_levels, int cell_size, int fast_threshold,
                           double detection_threshold, int box_mode, const uint8_t* occupancy, int win_w, int win_h, int max_level,
                           int* out_xy, int* out_level, float* out_score, int* n_out, dr3lk_pyramid** ref_pyramid)
{
    if (!ctx || !ref_pyramid) return DR3LK_E_ARG;
    *ref_pyramid = nullptr;
    LKArgs a{win_w, win_h, max_level, 0, 0, 0, 0., 0.};
    int rc = check_lk_args(ctx, w, h, a);
    if (rc != DR3LK_OK) return rc;
    return fast_detect_impl(ctx, img, w, h, step, n_levels, cell_size, fast_threshold, detection_threshold, box_mode, occupancy, out_xy,
                            out_level, out_score, n_out, win_w, win_h, max_level, ref_pyramid);
}

int dr3lk_init_second_frame(dr3lk_ctx* ctx, const dr3lk_pyramid* ref, const uint8_t* cur_img, size_t cur_step, const float* kps_ref,
                            const float* kps_cur, int n, int win_w, int win_h, int max_level, int crit_type, int crit_max_count,
                            double crit_eps, int flags, double min_eig_threshold, double fx, double fy, double cx, double cy,
                            const double* distortion, float* out_ref, float* out_cur, double* out_disparity, double* out_bearing,
                            uint8_t* out_status, float* out_err, int* n_kept)
{
    if (!ctx || !ref) return DR3LK_E_ARG;
    if (!n_kept) return fail(ctx, DR3LK_E_ARG, "init_second_frame: n_kept is null");
    *n_kept = 0;
    ctx->n_tracks = 0;
    LKArgs a{win_w, win_h, max_level, crit_type, crit_max_count, flags, crit_eps, min_eig_threshold};
    const int w = ref->w, h = ref->h;
    int rc = check_lk_args(ctx, w, h, a);
    if (rc != DR3LK_OK) return rc;
    if (ref->ctx != ctx) return fail(ctx, DR3LK_E_ARG, "the pyramid belongs to another context");
    if (ref->win_w != win_w || ref->win_h != win_h) return fail(ctx, DR3LK_E_ARG, "the pyramid was built for another window size");
    if (!ref->has_deriv) return fail(ctx, DR3LK_E_ARG, "the reference pyramid was built without derivatives");
    if (!cur_img || cur_step < (size_t)w) return fail(ctx, DR3LK_E_ARG, "init_second_frame: bad image arguments");
    if (n < 0) return fail(ctx, DR3LK_E_ARG, "negative point count");
    if (n == 0) return DR3LK_OK;
    if (!kps_ref || !out_ref || !out_cur || !out_disparity) return fail(ctx, DR3LK_E_ARG, "init_second_frame: null buffer");
    if ((flags & DR3LK_USE_INITIAL_FLOW) && !kps_cur) return fail(ctx, DR3LK_E_ARG, "OPTFLOW_USE_INITIAL_FLOW needs kps_cur");
    if (out_bearing && (fx == 0.0 || fy == 0.0)) return fail(ctx, DR3LK_E_ARG, "init_second_frame: zero focal length");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    Workspace& W = ctx->ws;
    dr3lk_pyramid* p2 = nullptr;
    rc = pyramid_alloc(ctx, w, h, win_w, win_h, max_level, false, &p2);
    if (rc != DR3LK_OK) return rc;
    // ONE upload: [new image][kps_ref 8n][offsets 16][kps_cur 8n]; device-only behind it: [err 4n][status n]
    const int pitch0 = align_up(w, 16);
    const size_t img_block = align_up_sz((size_t)pitch0 * h, 256);
    const size_t n8 = align_up_sz(8 * (size_t)n, 16), n4 = align_up_sz(4 * (size_t)n, 16), n1 = align_up_sz((size_t)n, 16), n24 = align_up_sz(24 * (size_t)n, 16);
    const size_t o_prev = img_block, o_offs = o_prev + n8, o_next = o_offs + 16, o_err = o_next + n8, o_status = o_err + n4, in_total = o_status + n1;
    // ONE download, from the resident track block: [out_ref 8n][out_cur 8n][disp 8n][bearing 24n][count 16][status n][err 4n]
    const size_t t_ref = 0, t_cur = n8, t_disp = 2 * n8, t_bear = 3 * n8, t_cnt = t_bear + n24, t_status = t_cnt + 16, t_err = t_status + n1,
                 t_total = t_err + n4;
    cudaError_t e = W.lvl0_prev.reserve(in_total);
    if (e == cudaSuccess) e = ctx->tracks.reserve(t_total);
    if (e == cudaSuccess) e = ctx->pinned.reserve(std::max(in_total, t_total));
    if (e != cudaSuccess) { pyramid_free(p2, false); return fail_cuda(ctx, e, "init_second_frame: allocation"); }
    uint8_t* dp = (uint8_t*)W.lvl0_prev.p;
    uint8_t* tp = (uint8_t*)ctx->tracks.p;
    uint8_t* hp = (uint8_t*)ctx->pinned.p;
    for (int y = 0; y < h; y++) memcpy(hp + (size_t)y * pitch0, cur_img + (size_t)y * cur_step, (size_t)w);
    const int offs[2] = {0, n};
    memcpy(hp + o_prev, kps_ref, 8 * (size_t)n);
    memcpy(hp + o_offs, offs, sizeof(offs));
    size_t in_bytes = o_next;
    if (flags & DR3LK_USE_INITIAL_FLOW) { memcpy(hp + o_next, kps_cur, 8 * (size_t)n); in_bytes = o_next + 8 * (size_t)n; }
    Launch L{st, cudaSuccess, 0};
    L.err = cudaMemcpyAsync(dp, hp, in_bytes, cudaMemcpyHostToDevice, st);
    pyramid_enqueue(ctx, p2, dp, pitch0, st, L);
    if (L.err != cudaSuccess) { cudaStreamSynchronize(st); pyramid_free(p2, false); return fail_cuda(ctx, L.err, "init_second_frame: pyramid of the new frame"); }
    LKParams lk;
    memset(&lk, 0, sizeof(lk));
    const int ml = std::min(std::min(ref->P.ml, p2->P.ml), max_level);
    for (int l = 0; l <= ml; l++) {
        lk.lv[l] = ref->lv[l];
        lk.lv[l].next = p2->lv[l].prev;
        lk.lv[l].pitch_n = p2->lv[l].pitch_p;
        lk.lv[l].next_stride = p2->lv[l].prev_stride;
    }
    lk.max_level = ml;
    lk.fast_ok = ref->P.ax > 0;
    // status / err go straight into the track block so that everything comes back in one copy
    rc = run_tracking(ctx, W, st, lk, 1, (const float*)(dp + o_prev), (float*)(dp + o_next), tp + t_status, (float*)(tp + t_err), offs,
                      (const int*)(dp + o_offs), n, nullptr, a);
    if (rc != DR3LK_OK) { cudaStreamSynchronize(st); pyramid_free(p2, false); return rc; }
    // src/initialization.cpp:615-635 on the device: erase !status (order kept), disparity, cam2world of the current point
    launch_filter_tracks(L, (const float*)(dp + o_prev), (const float*)(dp + o_next), tp + t_status, n, fx != 0.0 ? fx : 1.0, fy != 0.0 ? fy : 1.0,
                         cx, cy, distortion, (float*)(tp + t_ref), (float*)(tp + t_cur), (double*)(tp + t_disp),
                         out_bearing ? (double*)(tp + t_bear) : nullptr, (int*)(tp + t_cnt));
    ctx->launches += L.launches;
    if (L.err == cudaSuccess) L.err = cudaMemcpyAsync(hp, tp, t_total, cudaMemcpyDeviceToHost, st);
    if (L.err == cudaSuccess) L.err = cudaStreamSynchronize(st); else cudaStreamSynchronize(st);
    pyramid_free(p2, true);
    if (L.err != cudaSuccess) return fail_cuda(ctx, L.err, "init_second_frame");
    const int k = *reinterpret_cast<const int*>(hp + t_cnt);
    *n_kept = k;
    ctx->n_tracks = k;
    ctx->tracks_cur_offset = t_cur;
    memcpy(out_ref, hp + t_ref, 8 * (size_t)k);
    memcpy(out_cur, hp + t_cur, 8 * (size_t)k);
    memcpy(out_disparity, hp + t_disp, 8 * (size_t)k);
    if (out_bearing) memcpy(out_bearing, hp + t_bear, 24 * (size_t)k);
    if (out_status) memcpy(out_status, hp + t_status, (size_t)n);
    if (out_err) memcpy(out_err, hp + t_err, 4 * (size_t)n);
    return DR3LK_OK;
}


/* ---------------------------------------------------------------------------------------------- */
/* f-4: RANSAC fundamental-matrix hypothesis scoring                                               */
/* ---------------------------------------------------------------------------------------------- */

// dev1 / dev2 != nullptr: the matched points are already on the device (dr3lk_init_score_fundamental), only F goes up
static int score_fundamental_impl(dr3lk_ctx* ctx, const float* F21, int n_hyp, const float* pts1, const float* pts2, const float* dev1,
                                  const float* dev2, int n, float sigma, float* out_scores, uint8_t* out_inliers, int* best)
{
    if (!ctx) return DR3LK_E_ARG;
    if (n_hyp < 0 || n < 0 || !best) return fail(ctx, DR3LK_E_ARG, "score_fundamental: bad argument");
    *best = -1;
    if (n_hyp == 0) return DR3LK_OK;
    const bool resident = dev1 != nullptr;
    if (!F21 || !out_scores || (n > 0 && !resident && (!pts1 || !pts2)) || !(sigma > 0.f)) return fail(ctx, DR3LK_E_ARG, "score_fundamental: bad argument");
    cudaSetDevice(ctx->device);
    cudaStream_t st = ctx->stream;
    Workspace& W = ctx->ws;
    const size_t bF = align_up_sz(36 * (size_t)n_hyp, 16), bP = align_up_sz(8 * (size_t)n, 16), bS = align_up_sz(4 * (size_t)n_hyp, 16),
                 bI = out_inliers ? align_up_sz((size_t)n_hyp * n, 16) : 0;
    const size_t i_F = 0, i_p1 = bF, i_p2 = bF + bP, o_s = bF + 2 * bP, o_i = o_s + bS, total = o_i + bI;
    CU_TRY(ctx, W.pts.reserve(total));
    CU_TRY(ctx, ctx->pinned.reserve(total));
    uint8_t* dp = (uint8_t*)W.pts.p;
    uint8_t* hp = (uint8_t*)ctx->pinned.p;
    memcpy(hp + i_F, F21, 36 * (size_t)n_hyp);
    if (n > 0 && !resident) { memcpy(hp + i_p1, pts1, 8 * (size_t)n); memcpy(hp + i_p2, pts2, 8 * (size_t)n); }
    CU_TRY(ctx, cudaMemcpyAsync(dp, hp, resident ? bF : o_s, cudaMemcpyHostToDevice, st));
    Launch L{st, cudaSuccess, 0};
    // const float invSigmaSquare = 1.0/(sigma*sigma): float product, double division, rounded to float
    const float inv_sigma2 = (float)(1.0 / (double)(sigma * sigma));
    launch_score_fundamental(L, (const float*)(dp + i_F), n_hyp, resident ? dev1 : (const float*)(dp + i_p1), resident ? dev2 : (const float*)(dp + i_p2), n, inv_sigma2,
                             (float*)(dp + o_s), out_inliers ? dp + o_i : nullptr);
    ctx->launches += L.launches;
    if (L.err != cudaSuccess) return fail_cuda(ctx, L.err, "score_fundamental kernel launch");
    CU_TRY(ctx, cudaMemcpyAsync(hp + o_s, dp + o_s, total - o_s, cudaMemcpyDeviceToHost, st));
    CU_TRY(ctx, cudaStreamSynchronize(st));
    memcpy(out_scores, hp + o_s, 4 * (size_t)n_hyp);
    if (out_inliers) memcpy(out_inliers, hp + o_i, (size_t)n_hyp * n);
    // FindFundamental keeps a hypothesis only when currentScore > score (score starts at 0)
    float bs = 0.f;
    for (int i = 0; i < n_hyp; i++)
        if (out_scores[i] > bs) { bs = out_scores[i]; *best = i; }
    return DR3LK_OK;
}

int dr3lk_score_fundamental(dr3lk_ctx* ctx, const float* F21, int n_hyp, const float* pts1, const float* pts2, int n, float sigma,
                            float* out_scores, uint8_t* out_inliers, int* best)
{
    return score_fundamental_impl(ctx, F21, n_hyp, pts1, pts2, nullptr, nullptr, n, sigma, out_scores, out_inliers, best);
}

int dr3lk_init_score_fundamental(dr3lk_ctx* ctx, const float* F21, int n_hyp, float sigma, float* out_scores, uint8_t* out_inliers, int* best)
{
    if (!ctx) return DR3LK_E_ARG;
    if (ctx->n_tracks <= 0 || !ctx->tracks.p) return fail(ctx, DR3LK_E_ARG, "init_score_fundamental: no tracks resident (call dr3lk_init_second_frame first)");
    // layout of the resident block: [out_ref 8n'][out_cur 8n'] with n' = the point count of the second-frame call
    const float* d1 = (const float*)ctx->tracks.p;
    const float* d2 = (const float*)((const uint8_t*)ctx->tracks.p + ctx->tracks_cur_offset);
    return score_fundamental_impl(ctx, F21, n_hyp, nullptr, nullptr, d1, d2, ctx->n_tracks, sigma, out_scores, out_inliers, best);
}

}  // extern "C"
