// nphusl_cuda_api.cu -- the C ABI of include/nphusl_cuda.h: argument checking,
// per-device state, kernel dispatch and the host-buffer streaming pipeline.
//
// No CPU implementation of any conversion lives here (or anywhere in this
// library): a call either runs the CUDA kernels or returns a non-zero status.
#include <cuda_runtime.h>

#include <atomic>
#include <condition_variable>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/nphusl_cuda.h"
#include "husl_kernels.cuh"
#include "husl_stream.cuh"
#include "husl_lut.cuh"
#include "husl_lut_tile.cuh"
#include "husl_nv12.cuh"
#include "husl_nv12_out.cuh"
#include "husl_float.cuh"
#include "husl_fused.cuh"
#include "husl_strided.cuh"
#include "husl_rows.cuh"

namespace {

using namespace husl;

thread_local std::string t_err;
thread_local bool t_host_async = false;
// set by run_streamed around a kernel whose OUTPUT pointer is the device alias of a page-locked host buffer: kernels that
// would write it in 4-byte pieces (k_inv_direct) give way to the ones that write whole tiles (bulk stores)
thread_local bool t_out_is_host_alias = false;
// host -> host calls: the small side in place (NPHUSL_ZERO_COPY=0 turns it off, an A/B knob)
bool zero_copy_enabled() {
    static const bool v = [] {
        const char *e = std::getenv("NPHUSL_ZERO_COPY");
        return !(e && e[0] == '0');
    }();
    return v;
}
std::atomic<unsigned long long> g_launches{0};
std::atomic<size_t> g_chunk_px{0};

constexpr int MAX_DEV = 16;
constexpr int NSLOT = 3;                       // host streaming pipeline depth
constexpr size_t DEFAULT_CHUNK_PX = 1u << 21;  // at most 2 Mpx per slot (<= 48 MB per buffer)
constexpr size_t MIN_CHUNK_PX = 1u << 18;
constexpr size_t PAR_COPY_MIN = 1u << 20;      // host memcpy goes multi-threaded above this

// The entry points select `device` for the calling thread; the caller's own
// current device (torch tracks it through the same runtime state) is put back
// when the call returns.
struct DeviceGuard {
    int prev = -1;
    DeviceGuard() {
        if (cudaGetDevice(&prev) != cudaSuccess) {
            cudaGetLastError();
            prev = -1;
        }
    }
    ~DeviceGuard() {
        if (prev >= 0) cudaSetDevice(prev);
    }
};

int fail(int code, const std::string &msg) {
    t_err = msg;
    return code;
}

#define CU(call)                                                                         \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            cudaGetLastError(); /* reported here: must not surface again at the next launch check */ \
            return fail(NPHUSL_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); \
        }                                                                                \
    } while (0)

size_t dsize(int dt) { return dt == NPHUSL_U8 ? 1 : dt == NPHUSL_F32 ? 4 : dt == NPHUSL_F64 ? 8 : 0; }

struct Lut {
    bool ready = false;
    int seg = 0, nh = 0, nl = 0;
    int type = NPHUSL_LUT_U16;           // element type of light / chroma
    void *light = nullptr, *chroma = nullptr;
    double *linear = nullptr, *light_q = nullptr;   // light_q: derived table (husl_lut.cuh)
    double light_scale = 100, chroma_scale = 100;
    LutParams p{};
};
size_t lut_esize(int type) { return type == NPHUSL_LUT_U16 ? 2 : type == NPHUSL_LUT_F32 ? 4 : type == NPHUSL_LUT_F64 ? 8 : 0; }

// One staging set = NSLOT streams, each with a device-side chunk buffer per
// direction and (for pageable callers) a pinned bounce buffer.  A device owns
// several sets so that an upload-type call (host -> device result), a
// download-type call (device -> host result) and a host -> host call never
// queue behind one another: H2D of one call overlaps D2H of another.
// Host -> host calls have TWO sets, one for calls that download more than they
// upload and one for the rest: in the reference's data flow -- to_husl (3 B/px up,
// 24 down) followed by to_rgb (24 up, 3 down) per frame -- non-blocking calls on two
// user streams then run the HUSL download of one frame beside the HUSL upload of
// another (full-duplex link) instead of queueing every chunk of every call on the
// same three streams.
struct Staging {
    std::mutex mu;                       // one host-streamed call at a time per set (NOT DevCtx::mu: device-pointer calls never wait here)
    cudaStream_t st[NSLOT] = {};
    cudaEvent_t done[NSLOT] = {};
    void *din[NSLOT] = {}, *dout[NSLOT] = {};
    void *pin[NSLOT] = {}, *pout[NSLOT] = {};
    size_t din_cap = 0, dout_cap = 0, pin_cap = 0, pout_cap = 0;
    cudaEvent_t gate = nullptr;          // the caller's stream at the time of the call (recorded and waited on under mu)
};
enum { SET_UP = 0, SET_DOWN = 1, SET_BOTH = 2, SET_BOTH_B = 3, NSETS = 4 };

struct DevCtx {
    std::mutex mu;
    bool init = false;
    int dev = -1, sms = 0;
    Staging set[NSETS];
    std::atomic<bool> ready{false};      // init done: ctx_get's lock-free fast path
    std::atomic<bool> attr_set{false};   // device-pointer calls reach set_attrs() from any thread without c.mu
    Lut lut;
};

DevCtx g_ctx[MAX_DEV];

// ---- per-device initialisation -------------------------------------------------
int ctx_get(int device, DevCtx **out) {
    // fast path of every call after the first on a device: no device-count query, no mutex
    if (device >= 0 && device < MAX_DEV && g_ctx[device].ready.load(std::memory_order_acquire)) {
        CU(cudaSetDevice(device));
        *out = &g_ctx[device];
        return NPHUSL_OK;
    }
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess || n <= 0) {
        cudaGetLastError();
        return fail(NPHUSL_ERR_NO_DEVICE, "no CUDA device available");
    }
    if (device < 0 || device >= n || device >= MAX_DEV)
        return fail(NPHUSL_ERR_ARG, "device index out of range");
    DevCtx &c = g_ctx[device];
    std::lock_guard<std::mutex> lk(c.mu);
    CU(cudaSetDevice(device));
    if (!c.init) {
        cudaDeviceProp prop;
        CU(cudaGetDeviceProperties(&prop, device));
        if (prop.major < 10)
            return fail(NPHUSL_ERR_NO_DEVICE, std::string("device is sm_") + std::to_string(prop.major) +
                                                  std::to_string(prop.minor) + "; this library is built for sm_100a only");
        c.dev = device;
        c.sms = prop.multiProcessorCount;
        // linearisation table in float64 (nphusl.py:278-286 at c/255), split hi+lo
        float2 tab[256];
        for (int i = 0; i < 256; ++i) {
            const double x = i / 255.0;
            const double v = x > 0.04045 ? std::pow((x + 0.055) / 1.055, 2.4) : x / 12.92;
            tab[i].x = (float)v;
            tab[i].y = (float)(v - (double)tab[i].x);
        }
        CU(cudaMemcpyToSymbol(g_lin_tab, tab, sizeof tab));
        for (int k = 0; k < NSETS; ++k) {
            for (int s = 0; s < NSLOT; ++s) {
                CU(cudaStreamCreateWithFlags(&c.set[k].st[s], cudaStreamNonBlocking));
                CU(cudaEventCreateWithFlags(&c.set[k].done[s], cudaEventDisableTiming));
            }
            CU(cudaEventCreateWithFlags(&c.set[k].gate, cudaEventDisableTiming));
        }

        c.init = true;
        c.ready.store(true, std::memory_order_release);
    }
    *out = &c;
    return NPHUSL_OK;
}

// ---- kernel dispatch --------------------------------------------------------------
constexpr int FWD_BLOCK = 512;
constexpr int INV_BLOCK = 256;
// launch shapes of the float32 streaming kernels (threads per CTA, CTAs per SM), tuned on B200
// (profiles/r01_ab_variants.txt); overridable at build time for A/B runs
#ifndef NPHUSL_FWD32_BLOCK
#define NPHUSL_FWD32_BLOCK 1024      // one CTA of 32 warps (64 registers) beat 2 x 512 (48) by 2.8 %
#define NPHUSL_FWD32_MINB 1
#endif
#ifndef NPHUSL_HUE32_BLOCK
#define NPHUSL_HUE32_BLOCK 1024
#define NPHUSL_HUE32_MINB 2
#endif
#ifndef NPHUSL_INV32_BLOCK
#define NPHUSL_INV32_BLOCK 384
#define NPHUSL_INV32_MINB 4
#endif
#ifndef NPHUSL_INV32_IST
#define NPHUSL_INV32_IST 3
#endif
#ifndef NPHUSL_FWD32_TMA
#define NPHUSL_FWD32_TMA 0          // 1: float32 results leave through TMA bulk stores as well
#endif
#ifndef NPHUSL_FWD32_OST
#define NPHUSL_FWD32_OST 2
#endif
#ifndef NPHUSL_EDIT_BLOCK
#define NPHUSL_EDIT_BLOCK 1024      // one CTA of 32 warps: interleaved A/B (ab_interleaved.py) 0.3002 -> 0.2956 ms, planes 0.3176 -> 0.3000
#define NPHUSL_EDIT_MINB 1
#endif
constexpr int EDIT_BLOCK = NPHUSL_EDIT_BLOCK, EDIT_MINB = NPHUSL_EDIT_MINB, EDIT_W = NPHUSL_EDIT_BLOCK / 32;
#ifndef NPHUSL_FLT_BLOCK
#define NPHUSL_FLT_BLOCK 512        // tiled float-RGB kernels (husl_float.cuh); 512 x 1 best of 5 shapes for 3 of the 4 dtype pairs
#define NPHUSL_FLT_MINB 1
#define NPHUSL_FLT_IST 2
#endif
constexpr int FLT_BLOCK = NPHUSL_FLT_BLOCK, FLT_MINB = NPHUSL_FLT_MINB, FLT_IST = NPHUSL_FLT_IST;
#ifndef NPHUSL_F32_BLOCK
#define NPHUSL_F32_BLOCK 512        // k_fwd_flt<float, ...>: float32 channels, no FP64 (husl_math.cuh: diff_from_f32)
#define NPHUSL_F32_MINB 2         // float32 out (issue-bound): 512 x 2 0.0792 ms, 1024 x 1 0.0792, 512 x 1 0.0813, 768 x 1 0.0812
#define NPHUSL_F32_MINB64 1       // float64 out (HBM-bound):   512 x 1 0.1061 ms, 512 x 2 0.1092, 1024 x 1 0.1116
#define NPHUSL_F32_IST 2
#endif
constexpr int F32_BLOCK = NPHUSL_F32_BLOCK, F32_MINB = NPHUSL_F32_MINB, F32_MINB64 = NPHUSL_F32_MINB64, F32_IST = NPHUSL_F32_IST;
#ifndef NPHUSL_HUSL_EDIT_CTAS
#define NPHUSL_HUSL_EDIT_CTAS 8     // CTAs of 256 threads per SM for k_husl_edit_wide (A/B knob)
#endif
constexpr int HUSL_EDIT_CTAS = NPHUSL_HUSL_EDIT_CTAS;
#ifndef NPHUSL_FWD64_BLOCK
#define NPHUSL_FWD64_BLOCK 512
#endif
#ifndef NPHUSL_FWD64_OST
#define NPHUSL_FWD64_OST 2
#endif
#ifndef NPHUSL_INV64_BLOCK
#define NPHUSL_INV64_BLOCK 384      // round 2 (profiles/r02_ab_inv.txt): 384 x 2 with a 2-stage ring beats 256 x 2 / 3 stages inside the
#define NPHUSL_INV64_MINB 2        // to_husl -> to_rgb(edit) step (0.6141 -> 0.5975 ms) and without the edit (0.6104 -> 0.6039)
#endif
#ifndef NPHUSL_INV64_IST
#define NPHUSL_INV64_IST 2
#endif
#ifndef NPHUSL_INVD_BLOCK
#define NPHUSL_INVD_BLOCK 1024      // k_inv_direct: one CTA of 32 warps per SM (session 3 sweep: 1024 x 1 0.5798 ms per step, 768 x 1 0.5814,
#endif                              // 512 x 2 0.5849, 384 x 2 0.5881, 256 x 4 0.5925; 16 warps 0.64, 36 - 64 warps 0.60 - 0.65; TMA ring 0.6008)
constexpr int INVD_BLOCK = NPHUSL_INVD_BLOCK;
constexpr int FWD64_BLOCK = NPHUSL_FWD64_BLOCK, FWD64_OST = NPHUSL_FWD64_OST;
constexpr int INV64_BLOCK = NPHUSL_INV64_BLOCK, INV64_MINB = NPHUSL_INV64_MINB, INV64_IST = NPHUSL_INV64_IST;
[[maybe_unused]] constexpr int FWD32_OST = NPHUSL_FWD32_OST;
constexpr int FWD32_BLOCK = NPHUSL_FWD32_BLOCK, FWD32_MINB = NPHUSL_FWD32_MINB;
constexpr int HUE32_BLOCK = NPHUSL_HUE32_BLOCK, HUE32_MINB = NPHUSL_HUE32_MINB;
constexpr int INV32_BLOCK = NPHUSL_INV32_BLOCK, INV32_MINB = NPHUSL_INV32_MINB, INV32_IST = NPHUSL_INV32_IST;

template <typename K>
int set_smem(K kernel, size_t bytes) {
    CU(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    return NPHUSL_OK;
}

int set_attrs(DevCtx &c) {
    if (c.attr_set.load(std::memory_order_acquire)) return NPHUSL_OK;
    int rc;
    constexpr int W = FWD_BLOCK / 32;
    if ((rc = set_smem(k_fwd_u8<float, false, FWD32_BLOCK, FWD32_MINB>, TAB_BYTES + (FWD32_BLOCK / 32) * Pack<float>::STAGE))) return rc;
#if NPHUSL_FWD32_TMA
    if ((rc = set_smem(k_fwd_tma<float, FWD32_BLOCK, FWD32_MINB, FWD32_OST>, TAB_BYTES + (FWD32_BLOCK / 32) * FWD32_OST * Tile<float>::BYTES))) return rc;
#endif
    if ((rc = set_smem(k_fwd_u8<double, false, FWD_BLOCK, 1>, TAB_BYTES + W * Pack<double>::STAGE))) return rc;
    if ((rc = set_smem(k_fwd_u8<double, true, FWD_BLOCK, 3>, TAB_BYTES))) return rc;
    if ((rc = set_smem(k_fwd_tma<double, FWD64_BLOCK, 1, FWD64_OST>, TAB_BYTES + (FWD64_BLOCK / 32) * FWD64_OST * Tile<double>::BYTES))) return rc;
    if ((rc = set_smem(k_inv_tma<double, INV64_BLOCK, INV64_MINB, INV64_IST>, (INV64_BLOCK / 32) * InvSmem<double, INV64_IST>::PER_WARP))) return rc;
    if ((rc = set_smem(k_fwd_u8<float, true, HUE32_BLOCK, HUE32_MINB>, TAB_BYTES))) return rc;
    if ((rc = set_smem(k_inv_cp<INV32_BLOCK, INV32_MINB, INV32_IST>, (INV32_BLOCK / 32) * InvCpSmem<INV32_IST>::PER_WARP))) return rc;
    if ((rc = set_smem(k_inv_planes_cp<INV32_BLOCK, INV32_MINB, INV32_IST, false>, (INV32_BLOCK / 32) * InvCpSmem<INV32_IST>::PER_WARP))) return rc;
    if ((rc = set_smem(k_inv_planes_cp<INV32_BLOCK, INV32_MINB, INV32_IST, true>, (INV32_BLOCK / 32) * InvCpSmem<INV32_IST>::PER_WARP))) return rc;
    if ((rc = set_smem(k_fwd_planar<float, 512, 2>, TAB_BYTES))) return rc;
    if ((rc = set_smem(k_fwd_planar<double, 512, 2>, TAB_BYTES))) return rc;
#define EDIT_ATTRS(EH, ES, RSL) \
    if ((rc = set_smem(k_edit<EDIT_BLOCK, EDIT_MINB, EH, ES, RSL>, TAB_BYTES))) return rc; \
    if ((rc = set_smem(k_edit_chw<EDIT_BLOCK, EDIT_MINB, EH, ES, RSL>, TAB_BYTES))) return rc;
    EDIT_ATTRS(false, false, false) EDIT_ATTRS(false, false, true) EDIT_ATTRS(false, true, false) EDIT_ATTRS(false, true, true)
    EDIT_ATTRS(true, false, false) EDIT_ATTRS(true, false, true) EDIT_ATTRS(true, true, false) EDIT_ATTRS(true, true, true)
#undef EDIT_ATTRS
    if ((rc = set_smem(k_bands<EDIT_BLOCK, EDIT_MINB, false>, TAB_BYTES))) return rc;
    if ((rc = set_smem(k_bands<EDIT_BLOCK, EDIT_MINB, true>, TAB_BYTES))) return rc;
    if ((rc = set_smem(k_fwd_chw<float, 512, 2>, TAB_BYTES))) return rc;
    if ((rc = set_smem(k_fwd_chw<double, 512, 2>, TAB_BYTES))) return rc;

    c.attr_set.store(true, std::memory_order_release);
    return NPHUSL_OK;
}

// opt-in dynamic shared memory, once per device; `done` is a static of the CALL SITE (one per kernel
// instantiation -- kernels of equal signature share a function-pointer type, so it cannot live here)
template <typename K>
int flt_attr(std::atomic<bool> *done, DevCtx &c, K kernel, int bytes) {
    if (done[c.dev].load(std::memory_order_acquire)) return NPHUSL_OK;
    int rc = set_smem(kernel, (size_t)bytes);
    if (rc) return rc;
    done[c.dev].store(true, std::memory_order_release);
    return NPHUSL_OK;
}

int grid_for(long long work_items, int per_block, int max_blocks) {
    long long g = (work_items + per_block - 1) / per_block;
    if (g > max_blocks) g = max_blocks;
    if (g < 1) g = 1;
    return (int)g;
}

int check_launch(const char *what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return fail(NPHUSL_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
    return NPHUSL_OK;
}

template <typename TIn, typename TOut, bool HUE>
int launch_fwd_any(DevCtx &c, const void *in, void *out, long long n, int channels, cudaStream_t s) {
    if (n <= 0) return NPHUSL_OK;
    const int grid = grid_for(n, 256, c.sms * 8);
    k_fwd_any<TIn, TOut, HUE><<<grid, 256, 0, s>>>((const TIn *)in, (TOut *)out, n, channels);
    g_launches++;
    return check_launch("k_fwd_any");
}

template <typename TIn, typename TOut>
int launch_inv_any(DevCtx &c, const void *in, void *out, long long n, cudaStream_t s, const nphusl_edit *edit = nullptr,
                   const BandsDev *bands = nullptr) {
    if (n <= 0) return NPHUSL_OK;
    const int grid = grid_for(n, 256, c.sms * 8);
    if (bands) k_inv_any<TIn, TOut, 9><<<grid, 256, 0, s>>>((const TIn *)in, (TOut *)out, n, EditArg<9>{*bands});
    else if (edit) k_inv_any<TIn, TOut, 7><<<grid, 256, 0, s>>>((const TIn *)in, (TOut *)out, n, EditArg<7>{*edit});
    else k_inv_any<TIn, TOut><<<grid, 256, 0, s>>>((const TIn *)in, (TOut *)out, n);
    g_launches++;
    return check_launch("k_inv_any");
}

bool aligned(const void *p, size_t a) { return (reinterpret_cast<uintptr_t>(p) & (a - 1)) == 0; }

// NPHUSL_STREAM selects the streaming-kernel family for A/B measurements:
//   "ldg"  register/LDG+STG transposition kernels (husl_kernels.cuh)
//   "tma"  bulk-async-copy staged kernels (husl_stream.cuh)            [default]
bool use_tma() {
    static const bool v = [] {
        const char *e = std::getenv("NPHUSL_STREAM");
        return !(e && std::strcmp(e, "ldg") == 0);
    }();
    return v;
}

// forward on device pointers: streaming kernel on the 128-pixel tiles when the
// layout allows it, generic kernel on the tail / on everything else
template <bool HUE>
int run_fwd_dev(DevCtx &c, const void *in, int in_dt, int channels, void *out, int out_dt,
                long long n, cudaStream_t s) {
    int rc;
    const size_t out_px = (HUE ? 1 : 3) * dsize(out_dt);
    long long done = 0;
    if (HUE && channels == 1) {
        // grayscale -> hue is a constant per lit pixel: one streaming kernel, no per-pixel colour math
        if (n <= 0) return NPHUSL_OK;
        const int grid = grid_for(n, 256, c.sms * 8);
#define GRAY_GO(TI, TO) k_gray_hue<TI, TO><<<grid, 256, 0, s>>>((const TI *)in, (TO *)out, n)
        if (in_dt == NPHUSL_U8) { if (out_dt == NPHUSL_F32) GRAY_GO(uint8_t, float); else GRAY_GO(uint8_t, double); }
        else if (in_dt == NPHUSL_F32) { if (out_dt == NPHUSL_F32) GRAY_GO(float, float); else GRAY_GO(float, double); }
        else { if (out_dt == NPHUSL_F32) GRAY_GO(double, float); else GRAY_GO(double, double); }
#undef GRAY_GO
        g_launches++;
        return check_launch("k_gray_hue");
    }
    if (in_dt == NPHUSL_U8 && channels == 3 && aligned(in, 4) && aligned(out, HUE ? 32 : 16) && n >= TILE_PX) {
        if ((rc = set_attrs(c))) return rc;
        const long long tiles = n / TILE_PX;
        constexpr int W = FWD_BLOCK / 32;
        if (!HUE && use_tma() && out_dt == NPHUSL_F64) {
            // float64 output is HBM-bound: results leave through TMA bulk stores;
            // float32 output is issue-bound and keeps the leaner LDS/STG transposition
            const int grid = grid_for(tiles, FWD64_BLOCK / 32, c.sms);
            k_fwd_tma<double, FWD64_BLOCK, 1, FWD64_OST><<<grid, FWD64_BLOCK, TAB_BYTES + (FWD64_BLOCK / 32) * FWD64_OST * Tile<double>::BYTES, s>>>(
                (const uint32_t *)in, (double *)out, tiles);
        } else if (HUE) {
            // hue only: no staging buffer, 64 KB of table per CTA -> 3 CTAs (48 warps) per SM
            const int grid = grid_for(tiles, W, c.sms * 3);
            if (out_dt == NPHUSL_F32)   // float32 hue: 2 x 1024 threads (64 warps/SM, 32 registers) measured best
                k_fwd_u8<float, true, HUE32_BLOCK, HUE32_MINB><<<grid_for(tiles, HUE32_BLOCK / 32, c.sms * HUE32_MINB), HUE32_BLOCK, TAB_BYTES, s>>>(
                    (const uint32_t *)in, (float *)out, tiles);
            else
                k_fwd_u8<double, true, FWD_BLOCK, 3><<<grid, FWD_BLOCK, TAB_BYTES, s>>>((const uint32_t *)in, (double *)out, tiles);
        } else if (out_dt == NPHUSL_F32) {
            constexpr int W32 = FWD32_BLOCK / 32;
            const int grid = grid_for(tiles, W32, c.sms * FWD32_MINB);
#if NPHUSL_FWD32_TMA
            k_fwd_tma<float, FWD32_BLOCK, FWD32_MINB, FWD32_OST><<<grid, FWD32_BLOCK, TAB_BYTES + W32 * FWD32_OST * Tile<float>::BYTES, s>>>(
                (const uint32_t *)in, (float *)out, tiles);
#else
            const size_t sm = TAB_BYTES + W32 * Pack<float>::STAGE;
            k_fwd_u8<float, false, FWD32_BLOCK, FWD32_MINB><<<grid, FWD32_BLOCK, sm, s>>>((const uint32_t *)in, (float *)out, tiles);
#endif
        } else {
            const size_t sm = TAB_BYTES + W * Pack<double>::STAGE;
            const int grid = grid_for(tiles, W, c.sms);
            k_fwd_u8<double, false, FWD_BLOCK, 1><<<grid, FWD_BLOCK, sm, s>>>((const uint32_t *)in, (double *)out, tiles);
        }
        g_launches++;
        if ((rc = check_launch("k_fwd_u8"))) return rc;
        done = tiles * TILE_PX;
    }
    if (in_dt != NPHUSL_U8 && channels == 3 && aligned(in, 16) && aligned(out, HUE ? 32 : 16) && n >= TILE_PX) {
        // float RGB in [0,1] (the Cython backend's own input type): cp.async input ring + staged stores
        const long long tiles = n / TILE_PX;
#define FLT_GO(TI, TO, BLK, MB, ST) do { \
        constexpr int sm = (BLK / 32) * (HUE ? FltHueSmem<TI, ST>::PER_WARP : FltSmem<TI, TO, ST>::PER_WARP); \
        static std::atomic<bool> attr_done[MAX_DEV]; \
        if ((rc = flt_attr(attr_done, c, k_fwd_flt<TI, TO, HUE, BLK, MB, ST>, sm))) return rc; \
        k_fwd_flt<TI, TO, HUE, BLK, MB, ST><<<grid_for(tiles, BLK / 32, c.sms * MB), BLK, sm, s>>>((const TI *)in, (TO *)out, tiles); } while (0)
        if (in_dt == NPHUSL_F32) {
            if (out_dt == NPHUSL_F32) FLT_GO(float, float, F32_BLOCK, F32_MINB, F32_IST);
            else FLT_GO(float, double, F32_BLOCK, F32_MINB64, F32_IST);
        } else {
            if (out_dt == NPHUSL_F32) FLT_GO(double, float, FLT_BLOCK, FLT_MINB, FLT_IST);
            else FLT_GO(double, double, FLT_BLOCK, FLT_MINB, FLT_IST);
        }
#undef FLT_GO
        g_launches++;
        if ((rc = check_launch("k_fwd_flt"))) return rc;
        done = tiles * TILE_PX;
    }
    const long long rest = n - done;
    if (rest <= 0) return NPHUSL_OK;
    const char *ip = (const char *)in + done * channels * dsize(in_dt);
    char *op = (char *)out + done * out_px;
#define FWD_ANY(TI, TO) return launch_fwd_any<TI, TO, HUE>(c, ip, op, rest, channels, s)
    if (in_dt == NPHUSL_U8) { if (out_dt == NPHUSL_F32) FWD_ANY(uint8_t, float); else FWD_ANY(uint8_t, double); }
    if (in_dt == NPHUSL_F32) { if (out_dt == NPHUSL_F32) FWD_ANY(float, float); else FWD_ANY(float, double); }
    if (out_dt == NPHUSL_F32) FWD_ANY(double, float); else FWD_ANY(double, double);
#undef FWD_ANY
}

// `edit` != NULL: the HUSL-space edit is applied to every loaded pixel first (to_rgb(hsl, edit=...));
// the ring kernels have an EDIT instantiation, every other layout takes the generic kernel
template <int ED>
EditArg<ED> edit_arg(const nphusl_edit *edit, const BandsDev *bands) {
    if constexpr (ED >= 8) return EditArg<ED>{*bands};
    else return EditArg<ED>{*edit};
}

// `bands` != NULL (instead of `edit`): a banded map in the same stage (husl_bands_to_rgb)
int run_inv_dev(DevCtx &c, const void *in, int in_dt, void *out, int out_dt, long long n, cudaStream_t s,
                const nphusl_edit *edit = nullptr, const BandsDev *bands = nullptr) {
    int rc;
    long long done = 0;
    if ((edit || bands) && out_dt == NPHUSL_U8 && aligned(in, 16) && aligned(out, 16) && n >= TILE_PX) {
        const long long tiles = n / TILE_PX;
        const int code = bands ? (bands->b.s_half > 0.0f ? 9 : 8) : EditShape(*edit).code();
        long long tile_px = TILE_PX;            // pixels per unit of the kernel that ran (tiles of 128, or groups of 4)
        // instantiated per edit shape like k_edit: README-style edits (L scaled inside / outside a hue arc) carry no
        // hue-wrap, saturation or range arithmetic
#define INV_EDIT(ED) do { \
        const EditArg<ED> ea = edit_arg<ED>(edit, bands); \
        if (in_dt == NPHUSL_F32) { \
            constexpr int sm = (INV32_BLOCK / 32) * InvCpSmem<INV32_IST>::PER_WARP; \
            static std::atomic<bool> attr32[MAX_DEV]; \
            if ((rc = flt_attr(attr32, c, k_inv_cp<INV32_BLOCK, INV32_MINB, INV32_IST, ED>, sm))) return rc; \
            k_inv_cp<INV32_BLOCK, INV32_MINB, INV32_IST, ED><<<grid_for(tiles, INV32_BLOCK / 32, c.sms * INV32_MINB), INV32_BLOCK, sm, s>>>( \
                (const float *)in, (uint32_t *)out, tiles, ea); \
        } else if (aligned(in, 32) && !t_out_is_host_alias) { \
            k_inv_direct<INVD_BLOCK, 1, ED><<<grid_for(n / 4, INVD_BLOCK, c.sms), INVD_BLOCK, 0, s>>>((const double *)in, (uint32_t *)out, n / 4, ea); \
            tile_px = 4; \
        } else { \
            constexpr int sm = (INV64_BLOCK / 32) * InvSmem<double, INV64_IST>::PER_WARP; \
            static std::atomic<bool> attr64[MAX_DEV]; \
            if ((rc = flt_attr(attr64, c, k_inv_tma<double, INV64_BLOCK, INV64_MINB, INV64_IST, ED>, sm))) return rc; \
            k_inv_tma<double, INV64_BLOCK, INV64_MINB, INV64_IST, ED><<<grid_for(tiles, INV64_BLOCK / 32, c.sms * INV64_MINB), INV64_BLOCK, sm, s>>>( \
                (const double *)in, (uint32_t *)out, tiles, ea); \
        } } while (0)
        switch (code) {
            case 0: INV_EDIT(0); break;
            case 1: INV_EDIT(1); break;
            case 2: INV_EDIT(2); break;
            case 3: INV_EDIT(3); break;
            case 4: INV_EDIT(4); break;
            case 5: INV_EDIT(5); break;
            case 6: INV_EDIT(6); break;
            case 8: INV_EDIT(8); break;
            case 9: INV_EDIT(9); break;
            default: INV_EDIT(7); break;
        }
#undef INV_EDIT
        g_launches++;
        if ((rc = check_launch("k_inv (edit on load)"))) return rc;
        done = n / tile_px * tile_px;
    }
    if (!edit && !bands && out_dt == NPHUSL_U8 && aligned(in, 16) && aligned(out, 4) && n >= TILE_PX) {
        const long long tiles = n / TILE_PX;
        long long tile_px = TILE_PX;
        constexpr int W = INV_BLOCK / 32;
        if (use_tma() && aligned(out, 16)) {
            if ((rc = set_attrs(c))) return rc;
            // tuned on B200 (profiles/r01_ab_variants.txt).  float32 input is bound by the math
            // pipes: 3-stage LDGSTS ring; float64 input is HBM-bound: 3-stage bulk-copy (TMA)
            // ring with mbarriers, 2 CTAs/SM
            if (in_dt == NPHUSL_F32 && !t_out_is_host_alias) {
                // no ring either: 4 pixels (three 128-bit loads) per thread, 2 x 1024 threads per SM at 32 registers keep the same
                // ~96 KB per SM in flight as k_inv_direct: 0.1694 -> 0.1646 ms (session 3 sweep; every smaller shape loses)
                k_inv_direct32<1024, 2><<<grid_for(n / 4, 1024, c.sms * 2), 1024, 0, s>>>((const float4 *)in, (uint32_t *)out, n / 4);
                tile_px = 4;
            } else
            if (in_dt == NPHUSL_F32)   // 4 CTAs x 384 threads = 48 warps/SM at 40 registers
                k_inv_cp<INV32_BLOCK, INV32_MINB, INV32_IST><<<grid_for(tiles, INV32_BLOCK / 32, c.sms * INV32_MINB), INV32_BLOCK,
                                                               (INV32_BLOCK / 32) * InvCpSmem<INV32_IST>::PER_WARP, s>>>(
                    (const float *)in, (uint32_t *)out, tiles);
            else if (aligned(in, 32) && !t_out_is_host_alias) {
                k_inv_direct<INVD_BLOCK, 1><<<grid_for(n / 4, INVD_BLOCK, c.sms), INVD_BLOCK, 0, s>>>((const double *)in, (uint32_t *)out, n / 4);
                tile_px = 4;
            } else
                k_inv_tma<double, INV64_BLOCK, INV64_MINB, INV64_IST><<<grid_for(tiles, INV64_BLOCK / 32, c.sms * INV64_MINB), INV64_BLOCK,
                                                                        (INV64_BLOCK / 32) * InvSmem<double, INV64_IST>::PER_WARP, s>>>(
                    (const double *)in, (uint32_t *)out, tiles);
        } else if (in_dt == NPHUSL_F32) {
            const int grid = grid_for(tiles, W, c.sms * 4);
            k_inv_u8<float, INV_BLOCK, 4><<<grid, INV_BLOCK, W * Pack<float>::STAGE, s>>>((const float *)in, (uint32_t *)out, tiles);
        } else {
            const int grid = grid_for(tiles, W, c.sms * 4);
            k_inv_u8<double, INV_BLOCK, 4><<<grid, INV_BLOCK, W * Pack<double>::STAGE, s>>>((const double *)in, (uint32_t *)out, tiles);
        }
        g_launches++;
        if ((rc = check_launch("k_inv_u8"))) return rc;
        done = n / tile_px * tile_px;
    }
    if (!edit && !bands && out_dt != NPHUSL_U8 && aligned(in, 16) && aligned(out, 16) && n >= TILE_PX) {
        // float RGB in [0,1] out (what the reference's _husl_to_rgb returns, _cython_opt.pyx:177-192)
        if (in_dt == NPHUSL_F64 && out_dt == NPHUSL_F64 && aligned(in, 32) && aligned(out, 32)) {
            const long long groups = n / 4;
            k_inv_flt_wide<<<grid_for(groups, 256, c.sms * 8), 256, 0, s>>>((const double *)in, (double *)out, groups);
            g_launches++;
            if ((rc = check_launch("k_inv_flt_wide"))) return rc;
            done = groups * 4;
        } else {
        const long long tiles = n / TILE_PX;
        const int grid = grid_for(tiles, FLT_BLOCK / 32, c.sms * FLT_MINB);
#define FLT_GO(TI, TO) do { \
        constexpr int sm = (FLT_BLOCK / 32) * FltSmem<TI, TO, FLT_IST>::PER_WARP; \
        static std::atomic<bool> attr_done[MAX_DEV]; \
        if ((rc = flt_attr(attr_done, c, k_inv_flt<TI, TO, FLT_BLOCK, FLT_MINB, FLT_IST>, sm))) return rc; \
        k_inv_flt<TI, TO, FLT_BLOCK, FLT_MINB, FLT_IST><<<grid, FLT_BLOCK, sm, s>>>((const TI *)in, (TO *)out, tiles); } while (0)
        if (in_dt == NPHUSL_F32) { if (out_dt == NPHUSL_F32) FLT_GO(float, float); else FLT_GO(float, double); }
        else { if (out_dt == NPHUSL_F32) FLT_GO(double, float); else FLT_GO(double, double); }
#undef FLT_GO
        g_launches++;
        if ((rc = check_launch("k_inv_flt"))) return rc;
        done = tiles * TILE_PX;
        }
    }
    const long long rest = n - done;
    if (rest <= 0) return NPHUSL_OK;
    const char *ip = (const char *)in + done * 3 * dsize(in_dt);
    char *op = (char *)out + done * 3 * dsize(out_dt);
#define INV_ANY(TI, TO) return launch_inv_any<TI, TO>(c, ip, op, rest, s, edit, bands)
    if (in_dt == NPHUSL_F32) {
        if (out_dt == NPHUSL_U8) INV_ANY(float, uint8_t);
        if (out_dt == NPHUSL_F32) INV_ANY(float, float);
        INV_ANY(float, double);
    }
    if (out_dt == NPHUSL_U8) INV_ANY(double, uint8_t);
    if (out_dt == NPHUSL_F32) INV_ANY(double, float);
    INV_ANY(double, double);
#undef INV_ANY
}

#ifndef NPHUSL_LUT_BLOCK
#define NPHUSL_LUT_BLOCK 640       // A/B on B200 (profiles/r02_ab_lut.txt): 640 x 1 with 4 pixels per lane in flight
#define NPHUSL_LUT_MINB 1         // 0.523 ms; 512 x 1 0.547; 768 x 1 0.588; 384 x 2 0.585
#endif
#ifndef NPHUSL_LUT_NP
#define NPHUSL_LUT_NP 4             // pixels of a lane walked through the stages together (1, 2 or 4)
#endif
constexpr int LUT_BLOCK = NPHUSL_LUT_BLOCK, LUT_MINB = NPHUSL_LUT_MINB, LUT_NP = NPHUSL_LUT_NP;

int run_lut_dev(DevCtx &c, const void *in, void *out, long long n, int mode, cudaStream_t s) {
    if (!c.lut.ready) return fail(NPHUSL_ERR_NO_LUT, "LUT mode: call nphusl_cuda_lut_generate or nphusl_cuda_lut_upload first");
    if (n <= 0) return NPHUSL_OK;
    int rc;
    long long done = 0;
    if (aligned(in, 4) && aligned(out, 16) && n >= TILE_PX) {
        // whole 128-pixel tiles: the streaming kernel (husl_lut_tile.cuh)
        const long long tiles = n / TILE_PX;
        const int grid = grid_for(tiles, LUT_BLOCK / 32, c.sms * LUT_MINB);
        const bool light_smem = 3 * c.lut.seg * 8 <= LUT_LIGHT_SMEM_MAX;
#define LUT_TILE(INTERP, TT, LS) do { \
        constexpr int sm = lut_tile_smem<LS, LUT_BLOCK>(); \
        static std::atomic<bool> attr_done[MAX_DEV]; \
        if ((rc = flt_attr(attr_done, c, k_fwd_lut_tile<INTERP, TT, LUT_BLOCK, LUT_MINB, LS, LUT_NP>, sm))) return rc; \
        k_fwd_lut_tile<INTERP, TT, LUT_BLOCK, LUT_MINB, LS, LUT_NP><<<grid, LUT_BLOCK, sm, s>>>((const uint32_t *)in, (double *)out, tiles, c.lut.p); } while (0)
#define LUT_TILE_L(INTERP, TT) do { if (light_smem) LUT_TILE(INTERP, TT, true); else LUT_TILE(INTERP, TT, false); } while (0)
#define LUT_TILE_T(TT) do { if (mode == NPHUSL_LUT) LUT_TILE_L(false, TT); else LUT_TILE_L(true, TT); } while (0)
        if (c.lut.type == NPHUSL_LUT_U16) LUT_TILE_T(uint16_t);
        else if (c.lut.type == NPHUSL_LUT_F32) LUT_TILE_T(float);
        else LUT_TILE_T(double);
#undef LUT_TILE_T
#undef LUT_TILE_L
#undef LUT_TILE
        g_launches++;
        if ((rc = check_launch("k_fwd_lut_tile"))) return rc;
        done = tiles * TILE_PX;
    }
    const long long rest = n - done;
    if (rest <= 0) return NPHUSL_OK;
    const uint8_t *ip = (const uint8_t *)in + 3 * done;
    double *op = (double *)out + 3 * done;
    const int grid = grid_for(rest, 256, c.sms * 8);
#define LUT_GO(TT) do { if (mode == NPHUSL_LUT) k_fwd_lut<false, TT><<<grid, 256, 0, s>>>(ip, op, rest, c.lut.p); \
                        else k_fwd_lut<true, TT><<<grid, 256, 0, s>>>(ip, op, rest, c.lut.p); } while (0)
    if (c.lut.type == NPHUSL_LUT_U16) LUT_GO(uint16_t);
    else if (c.lut.type == NPHUSL_LUT_F32) LUT_GO(float);
    else LUT_GO(double);
#undef LUT_GO
    g_launches++;
    return check_launch("k_fwd_lut");
}

// ---- host-buffer streaming ----------------------------------------------------------
// Pageable <-> pinned bounce copies are the bottleneck of plain-numpy calls
// (a 1080p float64 result is 50 MB of first-touch page faults), so they run on
// a small persistent pool of host threads instead of one memcpy.
class CopyPool {
  public:
    static CopyPool &get() {
        static CopyPool p;
        return p;
    }
    void copy(void *dst, const void *src, size_t bytes) {
        if (bytes < PAR_COPY_MIN || workers_.empty()) {
            std::memcpy(dst, src, bytes);
            return;
        }
        std::unique_lock<std::mutex> call(call_mu_);          // one parallel copy at a time
        const size_t parts = workers_.size() + 1;
        const size_t per = ((bytes / parts) + 4095) & ~size_t(4095);
        {
            std::lock_guard<std::mutex> lk(mu_);
            dst_ = (char *)dst; src_ = (const char *)src; bytes_ = bytes; per_ = per;
            pending_ = (int)workers_.size();
            ++epoch_;
        }
        cv_.notify_all();
        std::memcpy(dst, src, per < bytes ? per : bytes);     // part 0 on the calling thread
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }

  private:
    CopyPool() {
        // 8 copy threads saturate the host memory system of the B200 boxes (16-24
        // logical cores); more only adds wake-up latency (gpurun_out/hostpath.jsonl)
        unsigned n = std::thread::hardware_concurrency() / 2;
        n = n < 2 ? 1 : (n > 8 ? 8 : n);
        if (const char *e = std::getenv("NPHUSL_COPY_THREADS")) {
            const int v = std::atoi(e);
            if (v >= 1 && v <= 64) n = (unsigned)v;
        }
        for (unsigned i = 1; i < n; ++i) workers_.emplace_back([this, i] { run(i); });
    }
    ~CopyPool() {
        {
            std::lock_guard<std::mutex> lk(mu_);
            stop_ = true;
        }
        cv_.notify_all();
        for (auto &t : workers_) t.join();
    }
    void run(unsigned idx) {
        unsigned long long seen = 0;
        for (;;) {
            std::unique_lock<std::mutex> lk(mu_);
            cv_.wait(lk, [&] { return stop_ || epoch_ != seen; });
            if (stop_) return;
            seen = epoch_;
            char *d = dst_; const char *s = src_; const size_t bytes = bytes_, per = per_;
            lk.unlock();
            const size_t off = per * idx;
            if (off < bytes) std::memcpy(d + off, s + off, off + per > bytes ? bytes - off : per);
            lk.lock();
            if (--pending_ == 0) done_.notify_one();
        }
    }
    std::vector<std::thread> workers_;
    std::mutex mu_, call_mu_;
    std::condition_variable cv_, done_;
    char *dst_ = nullptr; const char *src_ = nullptr;
    size_t bytes_ = 0, per_ = 0;
    int pending_ = 0;
    unsigned long long epoch_ = 0;
    bool stop_ = false;
};

void par_memcpy(void *dst, const void *src, size_t bytes) { CopyPool::get().copy(dst, src, bytes); }

bool is_pinned(const void *p) {
    cudaPointerAttributes a;
    if (cudaPointerGetAttributes(&a, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeHost || a.type == cudaMemoryTypeManaged;
}

int ensure(void **buf, size_t *cap, size_t need, bool pinned, int nbuf) {
    if (*cap >= need) return NPHUSL_OK;
    for (int s = 0; s < nbuf; ++s) {
        if (buf[s]) {
            if (pinned) cudaFreeHost(buf[s]); else cudaFree(buf[s]);
            buf[s] = nullptr;
        }
    }
    *cap = 0;
    for (int s = 0; s < nbuf; ++s) {
        if (pinned) CU(cudaHostAlloc(&buf[s], need, cudaHostAllocDefault));
        else CU(cudaMalloc(&buf[s], need));
    }
    *cap = need;
    return NPHUSL_OK;
}

// Generic driver for "n pixels, in_px bytes in, out_px bytes out per pixel".
// kernel(in_dev, out_dev, npx, stream) runs the conversion on device pointers.
//
// Host buffers are streamed in chunks through an NSLOT-deep pipeline (H2D,
// kernel, D2H of different chunks overlap).  Pageable host memory is bounced
// through pinned slots with a multi-threaded memcpy and the call returns when
// the result is in the caller's buffer.  When every host buffer is pinned the
// copies are DMA'd directly and the whole call is expressed in stream order;
// with nphusl_cuda_set_host_async(1) it then returns WITHOUT waiting: `user`
// (the caller's stream) is made to wait on the work instead.
template <typename F>
int run_streamed(DevCtx &c, const void *src, int src_loc, size_t in_px, void *dst, int dst_loc,
                 size_t out_px, long long n, cudaStream_t user, F kernel) {
    if (n <= 0) return NPHUSL_OK;
    if (src_loc == NPHUSL_DEVICE && dst_loc == NPHUSL_DEVICE) return kernel(src, dst, n, user);

    size_t chunk = g_chunk_px.load();
    if (!chunk) {
        // default: about 8 chunks per call so copies and kernels of neighbouring
        // chunks overlap, within [256 Kpx, 2 Mpx] per chunk
        chunk = (size_t)((n + 7) / 8);
        if (chunk < MIN_CHUNK_PX) chunk = MIN_CHUNK_PX;
        if (chunk > DEFAULT_CHUNK_PX) chunk = DEFAULT_CHUNK_PX;
    }
    chunk = (chunk + TILE_PX - 1) / TILE_PX * TILE_PX;
    if ((long long)chunk > n) chunk = (size_t)((n + TILE_PX - 1) / TILE_PX * TILE_PX);
    const bool src_host = src_loc == NPHUSL_HOST, dst_host = dst_loc == NPHUSL_HOST;
    const bool src_pin = src_host && is_pinned(src), dst_pin = dst_host && is_pinned(dst);
    const bool all_pinned = (!src_host || src_pin) && (!dst_host || dst_pin);
    // host -> host calls: one set for calls that bring more bytes down than they send up (to_husl: 3 up, 24 down),
    // the other for the rest (to_rgb: 24 up, 3 down), so that one kind runs beside the other
    const int which = src_host && dst_host ? (out_px >= in_px ? SET_BOTH : SET_BOTH_B) : (src_host ? SET_UP : SET_DOWN);
    Staging &g = c.set[which];
    std::lock_guard<std::mutex> lk(g.mu);   // the set's staging buffers and slot streams
    int rc;
    // Host -> host calls on pinned buffers: the SMALL side (uint8, 3 B/px against 12 or 24 B/px of HUSL) is read or
    // written by the kernel in place, through the device alias of the page-locked buffer, instead of through a copy.
    // A copy of the small side shares the DMA queue of its direction with the 25 MB HUSL chunks of whatever call runs
    // the other way (to_husl of one frame beside to_rgb of another: profiles/tools/host_husl_probe.py) and waits
    // behind them chunk after chunk; loads and stores issued by the SMs interleave with that traffic packet by packet.
    const char *src_alias = nullptr;
    char *dst_alias = nullptr;
    if (src_host && dst_host && all_pinned && zero_copy_enabled()) {
        void *p = nullptr;
        if (in_px * 4 <= out_px && cudaHostGetDevicePointer(&p, const_cast<void *>(src), 0) == cudaSuccess) src_alias = (const char *)p;
        else if (out_px * 4 <= in_px && cudaHostGetDevicePointer(&p, dst, 0) == cudaSuccess) dst_alias = (char *)p;
        cudaGetLastError();
    }
    if (src_host && !src_alias && (rc = ensure(g.din, &g.din_cap, chunk * in_px, false, NSLOT))) return rc;
    if (dst_host && !dst_alias && (rc = ensure(g.dout, &g.dout_cap, chunk * out_px, false, NSLOT))) return rc;
    if (src_host && !src_pin && (rc = ensure(g.pin, &g.pin_cap, chunk * in_px, true, NSLOT))) return rc;
    if (dst_host && !dst_pin && (rc = ensure(g.pout, &g.pout_cap, chunk * out_px, true, NSLOT))) return rc;

    // work submitted earlier on the caller's stream must be visible to ours
    CU(cudaEventRecord(g.gate, user));
    for (int s = 0; s < NSLOT; ++s) CU(cudaStreamWaitEvent(g.st[s], g.gate, 0));

    struct Pending { char *dst; size_t bytes; bool live; } pend[NSLOT] = {};
    long long off = 0, nchunks = 0;
    // one chunk: refill slot s, convert, start the download.  Any failure leaves the loop; the copies
    // already enqueued from / into the caller's buffers are drained below before the error is returned
    // (the caller may drop those buffers as soon as it sees the status).
    auto submit = [&](long long ci, int s, long long npx) -> int {
        // slot reuse: device-side buffers are protected by the slot stream's own
        // order; the pinned bounce buffers are touched by the HOST, so a pageable
        // call must wait for the slot's previous chunk before refilling it
        if (!all_pinned && ci >= NSLOT) {
            CU(cudaEventSynchronize(g.done[s]));
            if (pend[s].live) { par_memcpy(pend[s].dst, g.pout[s], pend[s].bytes); pend[s].live = false; }
        }
        const void *in_dev;
        if (src_alias) {
            in_dev = src_alias + off * in_px;
        } else if (src_host) {
            const char *hp = (const char *)src + off * in_px;
            if (!src_pin) { par_memcpy(g.pin[s], hp, npx * in_px); hp = (const char *)g.pin[s]; }
            CU(cudaMemcpyAsync(g.din[s], hp, npx * in_px, cudaMemcpyHostToDevice, g.st[s]));
            in_dev = g.din[s];
        } else {
            in_dev = (const char *)src + off * in_px;
        }
        void *out_dev = dst_alias ? (void *)(dst_alias + off * out_px) : (dst_host ? g.dout[s] : (void *)((char *)dst + off * out_px));
        t_out_is_host_alias = dst_alias != nullptr;
        int krc = kernel(in_dev, out_dev, npx, g.st[s]);
        t_out_is_host_alias = false;
        if (krc) return krc;
        if (dst_host && !dst_alias) {
            char *hp = (char *)dst + off * out_px;
            if (dst_pin) {
                CU(cudaMemcpyAsync(hp, g.dout[s], npx * out_px, cudaMemcpyDeviceToHost, g.st[s]));
            } else {
                CU(cudaMemcpyAsync(g.pout[s], g.dout[s], npx * out_px, cudaMemcpyDeviceToHost, g.st[s]));
                pend[s] = {hp, (size_t)npx * out_px, true};
            }
        }
        CU(cudaEventRecord(g.done[s], g.st[s]));
        return NPHUSL_OK;
    };
    for (long long ci = 0; off < n; ++ci, off += chunk, ++nchunks) {
        const long long npx = (n - off < (long long)chunk) ? n - off : (long long)chunk;
        if ((rc = submit(ci, (int)(ci % NSLOT), npx))) {
            const std::string why = t_err;
            for (int s = 0; s < NSLOT; ++s) cudaStreamSynchronize(g.st[s]);
            cudaGetLastError();
            return fail(rc, why);
        }
    }
    const int used = nchunks < NSLOT ? (int)nchunks : NSLOT;
    if (all_pinned && t_host_async) {
        for (int s = 0; s < used; ++s) CU(cudaStreamWaitEvent(user, g.done[s], 0));
        return NPHUSL_OK;
    }
    for (int s = 0; s < used; ++s) {
        CU(cudaStreamSynchronize(g.st[s]));
        if (pend[s].live) { par_memcpy(pend[s].dst, g.pout[s], pend[s].bytes); pend[s].live = false; }
    }
    return NPHUSL_OK;
}

bool ok_loc(int l) { return l == NPHUSL_HOST || l == NPHUSL_DEVICE; }

// ---- NV12 decoder surfaces ------------------------------------------------------------
int nv12_src(const nphusl_nv12 *src, const char *what, Nv12Src *out) {
    if (!src || !src->y || !src->uv) return fail(NPHUSL_ERR_ARG, std::string(what) + ": null surface");
    if (src->width <= 0 || src->height <= 0) return fail(NPHUSL_ERR_ARG, std::string(what) + ": width and height must be positive");
    if (src->y_pitch < src->width || src->uv_pitch < ((src->width + 1) / 2) * 2)
        return fail(NPHUSL_ERR_ARG, std::string(what) + ": pitch smaller than a row");
    if (src->matrix != NPHUSL_BT601 && src->matrix != NPHUSL_BT709) return fail(NPHUSL_ERR_ARG, std::string(what) + ": matrix must be BT601 or BT709");
    const double kr = src->matrix == NPHUSL_BT601 ? 0.299 : 0.2126, kb = src->matrix == NPHUSL_BT601 ? 0.114 : 0.0722;
    const double kg = 1.0 - kr - kb;
    const double ys = src->full_range ? 1.0 : 255.0 / 219.0, cs = src->full_range ? 1.0 : 255.0 / 224.0;
    Nv12Src o;
    o.y = (const uint8_t *)src->y; o.uv = (const uint8_t *)src->uv;
    o.y_pitch = src->y_pitch; o.uv_pitch = src->uv_pitch;
    o.width = src->width; o.height = src->height;
    o.k.ky = (int)std::lround(65536.0 * ys);
    o.k.rv = (int)std::lround(65536.0 * cs * 2.0 * (1.0 - kr));
    o.k.gu = (int)std::lround(65536.0 * cs * 2.0 * kb * (1.0 - kb) / kg);
    o.k.gv = (int)std::lround(65536.0 * cs * 2.0 * kr * (1.0 - kr) / kg);
    o.k.bu = (int)std::lround(65536.0 * cs * 2.0 * (1.0 - kb));
    const int y_off = src->full_range ? 0 : 16;
    o.k.cr = 32768 - o.k.ky * y_off - 128 * o.k.rv;
    o.k.cg = 32768 - o.k.ky * y_off + 128 * (o.k.gu + o.k.gv);
    o.k.cb = 32768 - o.k.ky * y_off - 128 * o.k.bu;
    *out = o;
    return NPHUSL_OK;
}

// whole groups of 4 pixels per lane: width % 4 == 0, word-aligned planes and pitches
bool nv12_vector_ok(const Nv12Src &s) {
    return s.width % 4 == 0 && aligned(s.y, 4) && aligned(s.uv, 4) && s.y_pitch % 4 == 0 && s.uv_pitch % 4 == 0;
}

template <int MODE, typename TOut, int BLOCK, int MINB, bool EH = true, bool ES = true, bool RSL = true>
int nv12_launch(DevCtx &c, const Nv12Src &s, void *out, size_t out_align, const nphusl_edit &e, cudaStream_t st) {
    int rc;
    if (nv12_vector_ok(s) && aligned(out, out_align)) {
        static std::atomic<bool> attr_done[MAX_DEV];
        const int smem = Nv12Smem<MODE, TOut>::bytes(BLOCK);
        if ((rc = flt_attr(attr_done, c, k_nv12<MODE, TOut, BLOCK, MINB, EH, ES, RSL>, smem))) return rc;
        const long long tiles = (long long)((s.width + TILE_PX - 1) / TILE_PX) * s.height;
        k_nv12<MODE, TOut, BLOCK, MINB, EH, ES, RSL><<<grid_for(tiles, BLOCK / 32, c.sms * MINB), BLOCK, smem, st>>>(s, (TOut *)out, e);
        g_launches++;
        return check_launch("k_nv12");
    }
    const long long n = (long long)s.width * s.height;
    k_nv12_any<MODE, TOut><<<grid_for(n, 256, c.sms * 8), 256, 0, st>>>(s, (TOut *)out, e);
    g_launches++;
    return check_launch("k_nv12_any");
}

int nv12_dst(const nphusl_nv12 *d, const char *what, Nv12Dst *out) {
    if (!d || !d->y || !d->uv) return fail(NPHUSL_ERR_ARG, std::string(what) + ": null destination surface");
    if (d->width <= 0 || d->height <= 0 || (d->width & 1) || (d->height & 1))
        return fail(NPHUSL_ERR_ARG, std::string(what) + ": destination width and height must be positive and even");
    if (d->y_pitch < d->width || d->uv_pitch < d->width) return fail(NPHUSL_ERR_ARG, std::string(what) + ": pitch smaller than a row");
    if (d->matrix != NPHUSL_BT601 && d->matrix != NPHUSL_BT709) return fail(NPHUSL_ERR_ARG, std::string(what) + ": matrix must be BT601 or BT709");
    const double kr = d->matrix == NPHUSL_BT601 ? 0.299 : 0.2126, kb = d->matrix == NPHUSL_BT601 ? 0.114 : 0.0722;
    const double kg = 1.0 - kr - kb;
    const double ys = d->full_range ? 1.0 : 219.0 / 255.0, cs = d->full_range ? 1.0 : 224.0 / 255.0;
    auto q = [](double v) { return (int)std::lround(65536.0 * v); };
    Nv12Dst o;
    o.y = (uint8_t *)d->y; o.uv = (uint8_t *)d->uv; o.y_pitch = d->y_pitch; o.uv_pitch = d->uv_pitch;
    o.k.yr = q(ys * kr); o.k.yg = q(ys * kg); o.k.yb = q(ys * kb);
    o.k.yc = 32768 + ((d->full_range ? 0 : 16) << 16);
    o.k.ur = q(-cs * 0.5 * kr / (1.0 - kb)); o.k.ug = q(-cs * 0.5 * kg / (1.0 - kb)); o.k.ub = q(cs * 0.5);
    o.k.vr = q(cs * 0.5); o.k.vg = q(-cs * 0.5 * kg / (1.0 - kr)); o.k.vb = q(-cs * 0.5 * kb / (1.0 - kr));
    o.k.cc = 32768 + (128 << 16);
    *out = o;
    return NPHUSL_OK;
}

bool nv12_dst_vector_ok(const Nv12Dst &d, int width) {
    return width % 4 == 0 && aligned(d.y, 4) && aligned(d.uv, 4) && d.y_pitch % 4 == 0 && d.uv_pitch % 4 == 0;
}

template <bool SRC_NV12, bool EH, bool ES, bool RSL>
int to_nv12_launch(DevCtx &c, const Nv12Src &s, const void *rgb, int W, int H, const Nv12Dst &d, const nphusl_edit &e,
                   bool vector_ok, cudaStream_t st) {
    int rc;
    if (vector_ok) {
        constexpr int BLK = 512, MB = 2;
        static std::atomic<bool> attr_done[MAX_DEV];
        const int smem = SRC_NV12 ? TAB_BYTES : 0;
        if (smem && (rc = flt_attr(attr_done, c, k_to_nv12<SRC_NV12, BLK, MB, EH, ES, RSL>, smem))) return rc;
        const long long tiles = (long long)((W + TILE_PX - 1) / TILE_PX) * (H / 2);
        k_to_nv12<SRC_NV12, BLK, MB, EH, ES, RSL><<<grid_for(tiles, BLK / 32, c.sms * MB), BLK, smem, st>>>(s, (const uint8_t *)rgb, W, H, d, e);
        g_launches++;
        return check_launch("k_to_nv12");
    }
    const long long n = (long long)(W / 2) * (H / 2);
    k_to_nv12_any<SRC_NV12><<<grid_for(n, 256, c.sms * 8), 256, 0, st>>>(s, (const uint8_t *)rgb, W, H, d, e);
    g_launches++;
    return check_launch("k_to_nv12_any");
}

// A dense view that collapsed to ONE long row (e.g. a whole contiguous RGBA batch seen through
// rgba[..., :3]) gives the row walk nothing to step down: re-tile it as rows of FLAT_ROW pixels plus a
// remainder row (row strides stay multiples of every alignment the row kernels need).
template <typename F>
int with_flat_split(const View2 &i, const View2 &o, F launch) {
    constexpr long long FLAT_ROW = 8192;
    if (i.frames * i.rows != 1 || i.cols < 2 * FLAT_ROW) return launch(i, o);
    const long long R = i.cols / FLAT_ROW, rem = i.cols - R * FLAT_ROW;
    View2 a = i, b = o;
    a.rows = b.rows = R; a.cols = b.cols = FLAT_ROW; a.rs = FLAT_ROW * i.cs; b.rs = FLAT_ROW * o.cs;
    const int rc = launch(a, b);
    if (rc || !rem) return rc;
    a = i; b = o;
    a.base += R * FLAT_ROW * i.cs; b.base += R * FLAT_ROW * o.cs; a.cols = b.cols = rem;
    return launch(a, b);
}


}  // namespace

// =====================================================================================
// C ABI
// =====================================================================================
extern "C" {

int nphusl_cuda_abi_version(void) { return NPHUSL_CUDA_ABI_VERSION; }

int nphusl_cuda_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

const char *nphusl_cuda_last_error(void) { return t_err.c_str(); }

unsigned long long nphusl_cuda_launch_count(void) { return g_launches.load(); }

int nphusl_cuda_set_chunk_pixels(size_t pixels) {
    g_chunk_px.store(pixels);
    return NPHUSL_OK;
}

int nphusl_cuda_set_host_async(int enabled) {
    t_host_async = enabled != 0;
    return NPHUSL_OK;
}

int nphusl_cuda_rgb_to_husl(const void *rgb, int rgb_dtype, int rgb_loc, int channels,
                            void *hsl, int hsl_dtype, int hsl_loc,
                            size_t n_pixels, int mode, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!rgb || !hsl) return fail(NPHUSL_ERR_ARG, "rgb_to_husl: null buffer");
    if (!dsize(rgb_dtype) || (hsl_dtype != NPHUSL_F32 && hsl_dtype != NPHUSL_F64))
        return fail(NPHUSL_ERR_ARG, "rgb_to_husl: rgb must be u8/f32/f64 and hsl f32/f64");
    if (channels != 3 && channels != 1 && channels != 4)
        return fail(NPHUSL_ERR_ARG, "rgb_to_husl: channels must be 3 (RGB), 1 (grayscale) or 4 (RGBA)");
    if (!ok_loc(rgb_loc) || !ok_loc(hsl_loc)) return fail(NPHUSL_ERR_ARG, "rgb_to_husl: bad buffer location");
    if (mode != NPHUSL_EXACT && mode != NPHUSL_LUT && mode != NPHUSL_LUT_INTERP)
        return fail(NPHUSL_ERR_ARG, "rgb_to_husl: unknown mode");
    if (mode != NPHUSL_EXACT && (rgb_dtype != NPHUSL_U8 || hsl_dtype != NPHUSL_F64 || channels != 3))
        return fail(NPHUSL_ERR_ARG, "rgb_to_husl: LUT modes take uint8 RGB and produce float64, like the reference's _simd.c");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    const size_t in_px = channels * dsize(rgb_dtype), out_px = 3 * dsize(hsl_dtype);
    auto kern = [=](const void *i, void *o, long long n, cudaStream_t s) -> int {
        if (mode != NPHUSL_EXACT) return run_lut_dev(*c, i, o, n, mode, s);
        return run_fwd_dev<false>(*c, i, rgb_dtype, channels, o, hsl_dtype, n, s);
    };
    return run_streamed(*c, rgb, rgb_loc, in_px, hsl, hsl_loc, out_px, (long long)n_pixels, (cudaStream_t)stream, kern);
}

int nphusl_cuda_rgb_to_hue(const void *rgb, int rgb_dtype, int rgb_loc, int channels,
                           void *hue, int hue_dtype, int hue_loc,
                           size_t n_pixels, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!rgb || !hue) return fail(NPHUSL_ERR_ARG, "rgb_to_hue: null buffer");
    if (!dsize(rgb_dtype) || (hue_dtype != NPHUSL_F32 && hue_dtype != NPHUSL_F64))
        return fail(NPHUSL_ERR_ARG, "rgb_to_hue: rgb must be u8/f32/f64 and hue f32/f64");
    if (channels != 3 && channels != 1 && channels != 4) return fail(NPHUSL_ERR_ARG, "rgb_to_hue: channels must be 3, 1 or 4");
    if (!ok_loc(rgb_loc) || !ok_loc(hue_loc)) return fail(NPHUSL_ERR_ARG, "rgb_to_hue: bad buffer location");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    const size_t in_px = channels * dsize(rgb_dtype), out_px = dsize(hue_dtype);
    auto kern = [=](const void *i, void *o, long long n, cudaStream_t s) -> int {
        return run_fwd_dev<true>(*c, i, rgb_dtype, channels, o, hue_dtype, n, s);
    };
    return run_streamed(*c, rgb, rgb_loc, in_px, hue, hue_loc, out_px, (long long)n_pixels, (cudaStream_t)stream, kern);
}

int nphusl_cuda_husl_to_rgb(const void *hsl, int hsl_dtype, int hsl_loc,
                            void *rgb, int rgb_dtype, int rgb_loc,
                            size_t n_pixels, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!hsl || !rgb) return fail(NPHUSL_ERR_ARG, "husl_to_rgb: null buffer");
    if ((hsl_dtype != NPHUSL_F32 && hsl_dtype != NPHUSL_F64) || !dsize(rgb_dtype))
        return fail(NPHUSL_ERR_ARG, "husl_to_rgb: hsl must be f32/f64 and rgb u8/f32/f64");
    if (!ok_loc(hsl_loc) || !ok_loc(rgb_loc)) return fail(NPHUSL_ERR_ARG, "husl_to_rgb: bad buffer location");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    const size_t in_px = 3 * dsize(hsl_dtype), out_px = 3 * dsize(rgb_dtype);
    auto kern = [=](const void *i, void *o, long long n, cudaStream_t s) -> int {
        return run_inv_dev(*c, i, hsl_dtype, o, rgb_dtype, n, s);
    };
    return run_streamed(*c, hsl, hsl_loc, in_px, rgb, rgb_loc, out_px, (long long)n_pixels, (cudaStream_t)stream, kern);
}

int nphusl_cuda_husl_edit_to_rgb(const void *hsl, int hsl_dtype, int hsl_loc, const nphusl_edit *edit,
                                 void *rgb, int rgb_dtype, int rgb_loc,
                                 size_t n_pixels, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!hsl || !rgb || !edit) return fail(NPHUSL_ERR_ARG, "husl_edit_to_rgb: null argument");
    if ((hsl_dtype != NPHUSL_F32 && hsl_dtype != NPHUSL_F64) || !dsize(rgb_dtype))
        return fail(NPHUSL_ERR_ARG, "husl_edit_to_rgb: hsl must be f32/f64 and rgb u8/f32/f64");
    if (!ok_loc(hsl_loc) || !ok_loc(rgb_loc)) return fail(NPHUSL_ERR_ARG, "husl_edit_to_rgb: bad buffer location");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    const nphusl_edit e = *edit;
    const size_t in_px = 3 * dsize(hsl_dtype), out_px = 3 * dsize(rgb_dtype);
    auto kern = [=](const void *i, void *o, long long n, cudaStream_t s) -> int {
        return run_inv_dev(*c, i, hsl_dtype, o, rgb_dtype, n, s, &e);
    };
    return run_streamed(*c, hsl, hsl_loc, in_px, rgb, rgb_loc, out_px, (long long)n_pixels, (cudaStream_t)stream, kern);
}

// nphusl_bands as the kernels take it, or an error for a descriptor that describes no bands
static int bands_check(const nphusl_bands *b, const char *what, BandsDev *out) {
    if (!b) return fail(NPHUSL_ERR_ARG, std::string(what) + ": null argument");
    if (b->src != NPHUSL_SRC_H && b->src != NPHUSL_SRC_S && b->src != NPHUSL_SRC_L)
        return fail(NPHUSL_ERR_ARG, std::string(what) + ": bands.src must be NPHUSL_SRC_H, _S or _L");
    if (!(b->width > 0.0f) || !(b->total > 0.0f) || !std::isfinite(b->width) || !std::isfinite(b->total))
        return fail(NPHUSL_ERR_ARG, std::string(what) + ": bands.width and bands.total must be positive and finite");
    if (!(b->total / b->width <= 1048576.0f))
        return fail(NPHUSL_ERR_ARG, std::string(what) + ": more than 2^20 bands");
    *out = bands_dev(*b);
    return NPHUSL_OK;
}

int nphusl_cuda_husl_bands_to_rgb(const void *hsl, int hsl_dtype, int hsl_loc, const nphusl_bands *bands,
                                  void *rgb, int rgb_dtype, int rgb_loc,
                                  size_t n_pixels, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!hsl || !rgb) return fail(NPHUSL_ERR_ARG, "husl_bands_to_rgb: null argument");
    if ((hsl_dtype != NPHUSL_F32 && hsl_dtype != NPHUSL_F64) || rgb_dtype != NPHUSL_U8)
        return fail(NPHUSL_ERR_ARG, "husl_bands_to_rgb: hsl must be f32/f64 and rgb u8");
    if (!ok_loc(hsl_loc) || !ok_loc(rgb_loc)) return fail(NPHUSL_ERR_ARG, "husl_bands_to_rgb: bad buffer location");
    BandsDev bd;
    int rc = bands_check(bands, "husl_bands_to_rgb", &bd);
    if (rc) return rc;
    DevCtx *c;
    if ((rc = ctx_get(device, &c))) return rc;
    const size_t in_px = 3 * dsize(hsl_dtype), out_px = 3;
    auto kern = [=](const void *i, void *o, long long n, cudaStream_t s) -> int {
        return run_inv_dev(*c, i, hsl_dtype, o, rgb_dtype, n, s, nullptr, &bd);
    };
    return run_streamed(*c, hsl, hsl_loc, in_px, rgb, rgb_loc, out_px, (long long)n_pixels, (cudaStream_t)stream, kern);
}

int nphusl_cuda_rgb_bands_rgb(const void *rgb_in, int in_loc, void *rgb_out, int out_loc,
                              size_t n_pixels, const nphusl_bands *bands, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!rgb_in || !rgb_out) return fail(NPHUSL_ERR_ARG, "rgb_bands_rgb: null argument");
    if (!ok_loc(in_loc) || !ok_loc(out_loc)) return fail(NPHUSL_ERR_ARG, "rgb_bands_rgb: bad buffer location");
    BandsDev bd;
    int rc = bands_check(bands, "rgb_bands_rgb", &bd);
    if (rc) return rc;
    DevCtx *c;
    if ((rc = ctx_get(device, &c))) return rc;
    const bool sat = bd.b.s_half > 0.0f;
    auto kern = [=](const void *i, void *o, long long n, cudaStream_t st) -> int {
        int r;
        long long done = 0;
        if (aligned(i, 4) && aligned(o, 4) && n >= TILE_PX) {
            if ((r = set_attrs(*c))) return r;
            const long long tiles = n / TILE_PX;
            const int grid = grid_for(tiles, EDIT_W, c->sms * EDIT_MINB);
            if (sat) k_bands<EDIT_BLOCK, EDIT_MINB, true><<<grid, EDIT_BLOCK, TAB_BYTES, st>>>((const uint32_t *)i, (uint32_t *)o, tiles, bd);
            else k_bands<EDIT_BLOCK, EDIT_MINB, false><<<grid, EDIT_BLOCK, TAB_BYTES, st>>>((const uint32_t *)i, (uint32_t *)o, tiles, bd);
            g_launches++;
            if ((r = check_launch("k_bands"))) return r;
            done = tiles * TILE_PX;
        }
        if (n > done) {
            k_bands_any<<<grid_for(n - done, 256, c->sms * 8), 256, 0, st>>>((const uint8_t *)i + 3 * done, (uint8_t *)o + 3 * done, n - done, bd);
            g_launches++;
            if ((r = check_launch("k_bands_any"))) return r;
        }
        return NPHUSL_OK;
    };
    return run_streamed(*c, rgb_in, in_loc, 3, rgb_out, out_loc, 3, (long long)n_pixels, (cudaStream_t)stream, kern);
}

int nphusl_cuda_husl_edit(const void *hsl_in, void *hsl_out, int hsl_dtype, size_t n_pixels,
                          const nphusl_edit *edit, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!hsl_in || !hsl_out || !edit) return fail(NPHUSL_ERR_ARG, "husl_edit: null argument");
    if (hsl_dtype != NPHUSL_F32 && hsl_dtype != NPHUSL_F64) return fail(NPHUSL_ERR_ARG, "husl_edit: hsl must be f32/f64");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    const long long n = (long long)n_pixels;
    const cudaStream_t st = (cudaStream_t)stream;
    // whole 96-byte groups (4 float64 / 8 float32 pixels) with 256-bit accesses, the rest one pixel per thread
    const long long per = hsl_dtype == NPHUSL_F32 ? 8 : 4;
    long long done = 0;
    if (aligned(hsl_in, 32) && aligned(hsl_out, 32) && n >= per) {
        const long long groups = n / per;
        const int grid = grid_for(groups, 256, c->sms * HUSL_EDIT_CTAS);
        if (hsl_dtype == NPHUSL_F32) k_husl_edit_wide<float><<<grid, 256, 0, st>>>((const unsigned char *)hsl_in, (unsigned char *)hsl_out, groups, *edit);
        else k_husl_edit_wide<double><<<grid, 256, 0, st>>>((const unsigned char *)hsl_in, (unsigned char *)hsl_out, groups, *edit);
        g_launches++;
        if ((rc = check_launch("k_husl_edit_wide"))) return rc;
        done = groups * per;
    }
    if (n > done) {
        const size_t off = (size_t)done * 3 * dsize(hsl_dtype);
        const int grid = grid_for(n - done, 256, c->sms * 8);
        if (hsl_dtype == NPHUSL_F32) k_husl_edit<float><<<grid, 256, 0, st>>>((const float *)((const char *)hsl_in + off), (float *)((char *)hsl_out + off), n - done, *edit);
        else k_husl_edit<double><<<grid, 256, 0, st>>>((const double *)((const char *)hsl_in + off), (double *)((char *)hsl_out + off), n - done, *edit);
        g_launches++;
        if ((rc = check_launch("k_husl_edit"))) return rc;
    }
    return NPHUSL_OK;
}

// ---- planar HUSL and the fused edit --------------------------------------------------------
int nphusl_cuda_rgb_to_husl_planar(const void *rgb, void *h, void *s, void *l, int hsl_dtype,
                                   size_t n_pixels, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!rgb || !h || !s || !l) return fail(NPHUSL_ERR_ARG, "rgb_to_husl_planar: null buffer");
    if (hsl_dtype != NPHUSL_F32 && hsl_dtype != NPHUSL_F64) return fail(NPHUSL_ERR_ARG, "rgb_to_husl_planar: planes must be f32/f64");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)n_pixels;
    const size_t es = dsize(hsl_dtype), va = hsl_dtype == NPHUSL_F32 ? 16 : 32;
    long long done = 0;
    if (aligned(rgb, 4) && aligned(h, va) && aligned(s, va) && aligned(l, va) && n >= TILE_PX) {
        if ((rc = set_attrs(*c))) return rc;
        const long long tiles = n / TILE_PX;
        const int grid = grid_for(tiles, 16, c->sms * 2);
        if (hsl_dtype == NPHUSL_F32)
            k_fwd_planar<float, 512, 2><<<grid, 512, TAB_BYTES, st>>>((const uint32_t *)rgb, (float *)h, (float *)s, (float *)l, tiles);
        else
            k_fwd_planar<double, 512, 2><<<grid, 512, TAB_BYTES, st>>>((const uint32_t *)rgb, (double *)h, (double *)s, (double *)l, tiles);
        g_launches++;
        if ((rc = check_launch("k_fwd_planar"))) return rc;
        done = tiles * TILE_PX;
    }
    if (n > done) {
        const long long rest = n - done;
        const int grid = grid_for(rest, 256, c->sms * 8);
        const uint8_t *ip = (const uint8_t *)rgb + 3 * done;
        if (hsl_dtype == NPHUSL_F32)
            k_fwd_planar_any<float><<<grid, 256, 0, st>>>(ip, (float *)h + done, (float *)s + done, (float *)l + done, rest);
        else
            k_fwd_planar_any<double><<<grid, 256, 0, st>>>(ip, (double *)h + done, (double *)s + done, (double *)l + done, rest);
        (void)es;
        g_launches++;
        if ((rc = check_launch("k_fwd_planar_any"))) return rc;
    }
    return NPHUSL_OK;
}

int nphusl_cuda_husl_planar_to_rgb(const void *h, const void *s, const void *l, int hsl_dtype,
                                   void *rgb, size_t n_pixels, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!rgb || !h || !s || !l) return fail(NPHUSL_ERR_ARG, "husl_planar_to_rgb: null buffer");
    if (hsl_dtype != NPHUSL_F32 && hsl_dtype != NPHUSL_F64) return fail(NPHUSL_ERR_ARG, "husl_planar_to_rgb: planes must be f32/f64");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)n_pixels;
    const size_t va = hsl_dtype == NPHUSL_F32 ? 16 : 32;
    long long done = 0;
    if (aligned(rgb, 4) && aligned(h, va) && aligned(s, va) && aligned(l, va) && n >= TILE_PX) {
        const long long tiles = n / TILE_PX;
        const int grid = grid_for(tiles, 8, c->sms * 4);
        if (hsl_dtype == NPHUSL_F32) {
            if ((rc = set_attrs(*c))) return rc;
            k_inv_planes_cp<INV32_BLOCK, INV32_MINB, INV32_IST, false><<<grid_for(tiles, INV32_BLOCK / 32, c->sms * INV32_MINB), INV32_BLOCK,
                                                                         (INV32_BLOCK / 32) * InvCpSmem<INV32_IST>::PER_WARP, st>>>(
                (const float *)h, (const float *)s, (const float *)l, (uint32_t *)rgb, nullptr, nullptr, tiles);
        }
        else
            k_inv_planar<double, 256, 4><<<grid, 256, 8 * 384, st>>>((const double *)h, (const double *)s, (const double *)l, (uint32_t *)rgb, tiles);
        g_launches++;
        if ((rc = check_launch("k_inv_planar"))) return rc;
        done = tiles * TILE_PX;
    }
    if (n > done) {
        const long long rest = n - done;
        const int grid = grid_for(rest, 256, c->sms * 8);
        uint8_t *op = (uint8_t *)rgb + 3 * done;
        if (hsl_dtype == NPHUSL_F32)
            k_inv_planar_any<float><<<grid, 256, 0, st>>>((const float *)h + done, (const float *)s + done, (const float *)l + done, op, rest);
        else
            k_inv_planar_any<double><<<grid, 256, 0, st>>>((const double *)h + done, (const double *)s + done, (const double *)l + done, op, rest);
        g_launches++;
        if ((rc = check_launch("k_inv_planar_any"))) return rc;
    }
    return NPHUSL_OK;
}

int nphusl_cuda_rgb_planes_to_husl_planes(const void *r, const void *g, const void *b,
                                          void *h, void *s, void *l, int hsl_dtype,
                                          size_t n_pixels, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!r || !g || !b || !h || !s || !l) return fail(NPHUSL_ERR_ARG, "rgb_planes_to_husl_planes: null buffer");
    if (hsl_dtype != NPHUSL_F32 && hsl_dtype != NPHUSL_F64) return fail(NPHUSL_ERR_ARG, "rgb_planes_to_husl_planes: planes must be f32/f64");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)n_pixels;
    const size_t va = hsl_dtype == NPHUSL_F32 ? 16 : 32;
    long long done = 0;
    if (aligned(r, 4) && aligned(g, 4) && aligned(b, 4) && aligned(h, va) && aligned(s, va) && aligned(l, va) && n >= TILE_PX) {
        if ((rc = set_attrs(*c))) return rc;
        const long long tiles = n / TILE_PX;
        const int grid = grid_for(tiles, 16, c->sms * 2);
        if (hsl_dtype == NPHUSL_F32)
            k_fwd_chw<float, 512, 2><<<grid, 512, TAB_BYTES, st>>>((const uint32_t *)r, (const uint32_t *)g, (const uint32_t *)b, (float *)h, (float *)s, (float *)l, tiles);
        else
            k_fwd_chw<double, 512, 2><<<grid, 512, TAB_BYTES, st>>>((const uint32_t *)r, (const uint32_t *)g, (const uint32_t *)b, (double *)h, (double *)s, (double *)l, tiles);
        g_launches++;
        if ((rc = check_launch("k_fwd_chw"))) return rc;
        done = tiles * TILE_PX;
    }
    if (n > done) {
        const long long rest = n - done;
        const int grid = grid_for(rest, 256, c->sms * 8);
        const uint8_t *pr = (const uint8_t *)r + done, *pg = (const uint8_t *)g + done, *pb = (const uint8_t *)b + done;
        if (hsl_dtype == NPHUSL_F32)
            k_fwd_chw_any<float><<<grid, 256, 0, st>>>(pr, pg, pb, (float *)h + done, (float *)s + done, (float *)l + done, rest);
        else
            k_fwd_chw_any<double><<<grid, 256, 0, st>>>(pr, pg, pb, (double *)h + done, (double *)s + done, (double *)l + done, rest);
        g_launches++;
        if ((rc = check_launch("k_fwd_chw_any"))) return rc;
    }
    return NPHUSL_OK;
}

int nphusl_cuda_husl_planes_to_rgb_planes(const void *h, const void *s, const void *l, int hsl_dtype,
                                          void *r, void *g, void *b,
                                          size_t n_pixels, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!r || !g || !b || !h || !s || !l) return fail(NPHUSL_ERR_ARG, "husl_planes_to_rgb_planes: null buffer");
    if (hsl_dtype != NPHUSL_F32 && hsl_dtype != NPHUSL_F64) return fail(NPHUSL_ERR_ARG, "husl_planes_to_rgb_planes: planes must be f32/f64");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)n_pixels;
    const size_t va = hsl_dtype == NPHUSL_F32 ? 16 : 32;
    long long done = 0;
    if (aligned(r, 4) && aligned(g, 4) && aligned(b, 4) && aligned(h, va) && aligned(s, va) && aligned(l, va) && n >= TILE_PX) {
        const long long tiles = n / TILE_PX;
        const int grid = grid_for(tiles, 12, c->sms * 4);
        if (hsl_dtype == NPHUSL_F32) {
            if ((rc = set_attrs(*c))) return rc;
            k_inv_planes_cp<INV32_BLOCK, INV32_MINB, INV32_IST, true><<<grid_for(tiles, INV32_BLOCK / 32, c->sms * INV32_MINB), INV32_BLOCK,
                                                                        (INV32_BLOCK / 32) * InvCpSmem<INV32_IST>::PER_WARP, st>>>(
                (const float *)h, (const float *)s, (const float *)l, (uint32_t *)r, (uint32_t *)g, (uint32_t *)b, tiles);
        }
        else
            k_inv_chw<double, 384, 4><<<grid, 384, 0, st>>>((const double *)h, (const double *)s, (const double *)l, (uint32_t *)r, (uint32_t *)g, (uint32_t *)b, tiles);
        g_launches++;
        if ((rc = check_launch("k_inv_chw"))) return rc;
        done = tiles * TILE_PX;
    }
    if (n > done) {
        const long long rest = n - done;
        const int grid = grid_for(rest, 256, c->sms * 8);
        uint8_t *pr = (uint8_t *)r + done, *pg = (uint8_t *)g + done, *pb = (uint8_t *)b + done;
        if (hsl_dtype == NPHUSL_F32)
            k_inv_chw_any<float><<<grid, 256, 0, st>>>((const float *)h + done, (const float *)s + done, (const float *)l + done, pr, pg, pb, rest);
        else
            k_inv_chw_any<double><<<grid, 256, 0, st>>>((const double *)h + done, (const double *)s + done, (const double *)l + done, pr, pg, pb, rest);
        g_launches++;
        if ((rc = check_launch("k_inv_chw_any"))) return rc;
    }
    return NPHUSL_OK;
}

int nphusl_cuda_rgb_edit_rgb(const void *rgb_in, int in_loc, void *rgb_out, int out_loc,
                             size_t n_pixels, const nphusl_edit *edit, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!rgb_in || !rgb_out || !edit) return fail(NPHUSL_ERR_ARG, "rgb_edit_rgb: null argument");
    if (!ok_loc(in_loc) || !ok_loc(out_loc)) return fail(NPHUSL_ERR_ARG, "rgb_edit_rgb: bad buffer location");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    const nphusl_edit e = *edit;
    auto kern = [=](const void *i, void *o, long long n, cudaStream_t st) -> int {
        int r;
        long long done = 0;
        if (aligned(i, 4) && aligned(o, 4) && n >= TILE_PX) {
            if ((r = set_attrs(*c))) return r;
            const long long tiles = n / TILE_PX;
            const EditShape sh(e);
            const int grid = grid_for(tiles, EDIT_W, c->sms * EDIT_MINB);
#define EDIT_GO(EH, ES, RSL) \
    if (sh.eh == EH && sh.es == ES && sh.rsl == RSL) \
        k_edit<EDIT_BLOCK, EDIT_MINB, EH, ES, RSL><<<grid, EDIT_BLOCK, TAB_BYTES, st>>>((const uint32_t *)i, (uint32_t *)o, tiles, e);
            EDIT_GO(false, false, false) EDIT_GO(false, false, true) EDIT_GO(false, true, false) EDIT_GO(false, true, true)
            EDIT_GO(true, false, false) EDIT_GO(true, false, true) EDIT_GO(true, true, false) EDIT_GO(true, true, true)
#undef EDIT_GO
            g_launches++;
            if ((r = check_launch("k_edit"))) return r;
            done = tiles * TILE_PX;
        }
        if (n > done) {
            k_edit_any<<<grid_for(n - done, 256, c->sms * 8), 256, 0, st>>>((const uint8_t *)i + 3 * done, (uint8_t *)o + 3 * done, n - done, e);
            g_launches++;
            if ((r = check_launch("k_edit_any"))) return r;
        }
        return NPHUSL_OK;
    };
    return run_streamed(*c, rgb_in, in_loc, 3, rgb_out, out_loc, 3, (long long)n_pixels, (cudaStream_t)stream, kern);
}

int nphusl_cuda_rgb_planes_edit(const void *r, const void *g, const void *b,
                                void *r_out, void *g_out, void *b_out,
                                size_t n_pixels, const nphusl_edit *edit, int device, void *stream) {
    DeviceGuard guard;
    if (n_pixels == 0) return NPHUSL_OK;
    if (!r || !g || !b || !r_out || !g_out || !b_out || !edit) return fail(NPHUSL_ERR_ARG, "rgb_planes_edit: null argument");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long n = (long long)n_pixels;
    long long done = 0;
    if (aligned(r, 4) && aligned(g, 4) && aligned(b, 4) && aligned(r_out, 4) && aligned(g_out, 4) && aligned(b_out, 4) && n >= TILE_PX) {
        if ((rc = set_attrs(*c))) return rc;
        const long long tiles = n / TILE_PX;
        const EditShape sh(*edit);
        const int grid = grid_for(tiles, EDIT_W, c->sms * EDIT_MINB);
#define EDIT_GO(EH, ES, RSL) \
    if (sh.eh == EH && sh.es == ES && sh.rsl == RSL) \
        k_edit_chw<EDIT_BLOCK, EDIT_MINB, EH, ES, RSL><<<grid, EDIT_BLOCK, TAB_BYTES, st>>>((const uint32_t *)r, (const uint32_t *)g, (const uint32_t *)b, \
                                                                      (uint32_t *)r_out, (uint32_t *)g_out, (uint32_t *)b_out, tiles, *edit);
        EDIT_GO(false, false, false) EDIT_GO(false, false, true) EDIT_GO(false, true, false) EDIT_GO(false, true, true)
        EDIT_GO(true, false, false) EDIT_GO(true, false, true) EDIT_GO(true, true, false) EDIT_GO(true, true, true)
#undef EDIT_GO
        g_launches++;
        if ((rc = check_launch("k_edit_chw"))) return rc;
        done = tiles * TILE_PX;
    }
    if (n > done) {
        const long long rest = n - done;
        k_edit_chw_any<<<grid_for(rest, 256, c->sms * 8), 256, 0, st>>>(
            (const uint8_t *)r + done, (const uint8_t *)g + done, (const uint8_t *)b + done,
            (uint8_t *)r_out + done, (uint8_t *)g_out + done, (uint8_t *)b_out + done, rest, *edit);
        g_launches++;
        if ((rc = check_launch("k_edit_chw_any"))) return rc;
    }
    return NPHUSL_OK;
}

// ---- NV12 decoder surfaces ------------------------------------------------------------
int nphusl_cuda_nv12_to_rgb(const nphusl_nv12 *src, void *rgb, int device, void *stream) {
    DeviceGuard guard;
    Nv12Src s;
    int rc = nv12_src(src, "nv12_to_rgb", &s);
    if (rc) return rc;
    if (!rgb) return fail(NPHUSL_ERR_ARG, "nv12_to_rgb: null output");
    DevCtx *c;
    if ((rc = ctx_get(device, &c))) return rc;
    return nv12_launch<NV12_RGB, uint8_t, 512, 4>(*c, s, rgb, 4, nphusl_edit{}, (cudaStream_t)stream);
}

int nphusl_cuda_nv12_to_husl(const nphusl_nv12 *src, void *hsl, int hsl_dtype, int hue_only, int device, void *stream) {
    DeviceGuard guard;
    Nv12Src s;
    int rc = nv12_src(src, "nv12_to_husl", &s);
    if (rc) return rc;
    if (!hsl) return fail(NPHUSL_ERR_ARG, "nv12_to_husl: null output");
    if (hsl_dtype != NPHUSL_F32 && hsl_dtype != NPHUSL_F64) return fail(NPHUSL_ERR_ARG, "nv12_to_husl: output must be f32/f64");
    DevCtx *c;
    if ((rc = ctx_get(device, &c))) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const nphusl_edit none{};
    if (hue_only)
        return hsl_dtype == NPHUSL_F32 ? nv12_launch<NV12_HUE, float, 1024, 2>(*c, s, hsl, 16, none, st)
                                       : nv12_launch<NV12_HUE, double, 512, 3>(*c, s, hsl, 32, none, st);
    return hsl_dtype == NPHUSL_F32 ? nv12_launch<NV12_HUSL, float, 1024, 1>(*c, s, hsl, 16, none, st)
                                   : nv12_launch<NV12_HUSL, double, 512, 1>(*c, s, hsl, 16, none, st);
}

int nphusl_cuda_nv12_edit_rgb(const nphusl_nv12 *src, const nphusl_edit *edit, void *rgb, int device, void *stream) {
    DeviceGuard guard;
    Nv12Src s;
    int rc = nv12_src(src, "nv12_edit_rgb", &s);
    if (rc) return rc;
    if (!rgb || !edit) return fail(NPHUSL_ERR_ARG, "nv12_edit_rgb: null argument");
    DevCtx *c;
    if ((rc = ctx_get(device, &c))) return rc;
    const EditShape sh(*edit);
#define EDIT_GO(EH, ES, RSL) \
    if (sh.eh == EH && sh.es == ES && sh.rsl == RSL) \
        return nv12_launch<NV12_EDIT, uint8_t, EDIT_BLOCK, EDIT_MINB, EH, ES, RSL>(*c, s, rgb, 4, *edit, (cudaStream_t)stream);
    EDIT_GO(false, false, false) EDIT_GO(false, false, true) EDIT_GO(false, true, false) EDIT_GO(false, true, true)
    EDIT_GO(true, false, false) EDIT_GO(true, false, true) EDIT_GO(true, true, false)
#undef EDIT_GO
    return nv12_launch<NV12_EDIT, uint8_t, EDIT_BLOCK, EDIT_MINB, true, true, true>(*c, s, rgb, 4, *edit, (cudaStream_t)stream);
}

int nphusl_cuda_rgb_to_nv12(const void *rgb, const nphusl_nv12 *dst, int device, void *stream) {
    DeviceGuard guard;
    Nv12Dst d;
    int rc = nv12_dst(dst, "rgb_to_nv12", &d);
    if (rc) return rc;
    if (!rgb) return fail(NPHUSL_ERR_ARG, "rgb_to_nv12: null input");
    DevCtx *c;
    if ((rc = ctx_get(device, &c))) return rc;
    const bool vec = nv12_dst_vector_ok(d, dst->width) && aligned(rgb, 4);
    return to_nv12_launch<false, true, true, true>(*c, Nv12Src{}, rgb, dst->width, dst->height, d, nphusl_edit{}, vec, (cudaStream_t)stream);
}

int nphusl_cuda_nv12_edit_nv12(const nphusl_nv12 *src, const nphusl_edit *edit, const nphusl_nv12 *dst, int device, void *stream) {
    DeviceGuard guard;
    Nv12Src s;
    Nv12Dst d;
    int rc = nv12_src(src, "nv12_edit_nv12", &s);
    if (rc) return rc;
    if ((rc = nv12_dst(dst, "nv12_edit_nv12", &d))) return rc;
    if (!edit) return fail(NPHUSL_ERR_ARG, "nv12_edit_nv12: null edit");
    if (src->width != dst->width || src->height != dst->height) return fail(NPHUSL_ERR_ARG, "nv12_edit_nv12: source and destination sizes differ");
    DevCtx *c;
    if ((rc = ctx_get(device, &c))) return rc;
    const bool vec = nv12_vector_ok(s) && nv12_dst_vector_ok(d, dst->width);
    const EditShape sh(*edit);
    cudaStream_t st = (cudaStream_t)stream;
#define EDIT_GO(EH, ES, RSL) \
    if (sh.eh == EH && sh.es == ES && sh.rsl == RSL) \
        return to_nv12_launch<true, EH, ES, RSL>(*c, s, nullptr, src->width, src->height, d, *edit, vec, st);
    EDIT_GO(false, false, false) EDIT_GO(false, false, true) EDIT_GO(false, true, false) EDIT_GO(false, true, true)
    EDIT_GO(true, false, false) EDIT_GO(true, false, true) EDIT_GO(true, true, false)
#undef EDIT_GO
    return to_nv12_launch<true, true, true, true>(*c, s, nullptr, src->width, src->height, d, *edit, vec, st);
}

// ---- strided device images ------------------------------------------------------------
static int view_check(const nphusl_view *v, const char *what, bool want_u8, bool want_float) {
    if (!v || !v->ptr) return fail(NPHUSL_ERR_ARG, std::string(what) + ": null view");
    if (v->frames < 0 || v->rows < 0 || v->cols < 0) return fail(NPHUSL_ERR_ARG, std::string(what) + ": negative extent");
    const bool is_u8 = v->dtype == NPHUSL_U8, is_f = v->dtype == NPHUSL_F32 || v->dtype == NPHUSL_F64;
    if (!((want_u8 && is_u8) || (want_float && is_f))) return fail(NPHUSL_ERR_ARG, std::string(what) + ": dtype not allowed here");
    const long long a = (long long)dsize(v->dtype);
    if ((reinterpret_cast<uintptr_t>(v->ptr) % a) || (v->frame_stride % a) || (v->row_stride % a) || (v->col_stride % a) || (v->ch_stride % a))
        return fail(NPHUSL_ERR_ARG, std::string(what) + ": pointer and strides must be multiples of the element size");
    return NPHUSL_OK;
}

static View2 view_dev(const nphusl_view *v) {
    return View2{(char *)v->ptr, v->frames, v->rows, v->cols, v->frame_stride, v->row_stride, v->col_stride, v->ch_stride};
}

static dim3 grid_3d(const DevCtx &c, const View2 &v) {
    // enough CTAs to fill the GPU; rows / frames beyond that are grid-strided
    const long long cap = (long long)c.sms * 16;
    long long gx = (v.cols + 255) / 256, gy = v.rows, gz = v.frames;
    if (gx > cap) gx = cap;
    if (gx * gy > cap) gy = cap / gx > 0 ? cap / gx : 1;
    if (gx * gy * gz > cap) gz = cap / (gx * gy) > 0 ? cap / (gx * gy) : 1;
    if (gy > 65535) gy = 65535;
    if (gz > 65535) gz = 65535;
    return dim3((unsigned)gx, (unsigned)gy, (unsigned)gz);
}
static bool same_extent(const nphusl_view *a, const nphusl_view *b) {
    return a->frames == b->frames && a->rows == b->rows && a->cols == b->cols;
}
static bool empty_view(const nphusl_view *a) { return a->frames == 0 || a->rows == 0 || a->cols == 0; }

// Rows of contiguous pixels with aligned starts: the row-tiled kernels of husl_rows.cuh apply.
// `unit` = elements per pixel in this view (3, or 1 for a hue image), `align` = bytes every row start must honour.
static bool dense_rows(const nphusl_view *v, int unit, long long align) {
    const long long a = (long long)dsize(v->dtype);
    if (v->col_stride != unit * a || (unit == 3 && v->ch_stride != a) || v->cols % 4) return false;
    return reinterpret_cast<uintptr_t>(v->ptr) % align == 0 && v->frame_stride % align == 0 && v->row_stride % align == 0;
}
static long long row_tiles(const View2 &v) { return (v.cols + TILE_PX - 1) / TILE_PX * v.frames * v.rows; }
int nphusl_cuda_rgb_to_husl_strided(const nphusl_view *rgb, const nphusl_view *hsl, int hue_only,
                                    int device, void *stream) {
    DeviceGuard guard;
    int rc;
    if ((rc = view_check(rgb, "rgb_to_husl_strided: rgb", true, true))) return rc;
    if ((rc = view_check(hsl, "rgb_to_husl_strided: hsl", false, true))) return rc;
    if (!same_extent(rgb, hsl)) return fail(NPHUSL_ERR_ARG, "rgb_to_husl_strided: extents differ");
    if (empty_view(rgb)) return NPHUSL_OK;
    DevCtx *c;
    if ((rc = ctx_get(device, &c))) return rc;
    const View2 i = view_dev(rgb), o = view_dev(hsl);
    const dim3 grid = grid_3d(*c, i);
    cudaStream_t st = (cudaStream_t)stream;
    // uint8 side: 3 = packed RGB rows; 4 = RGBA / RGBX memory seen through rgba[..., :3]; -4 = BGRA memory
    // presented as RGB (channel stride -1, the view's pointer is the R byte of the first pixel)
    int pxb = 0;
    if (rgb->dtype == NPHUSL_U8) {
        if (dense_rows(rgb, 3, 4)) pxb = 3;
        else if (rgb->col_stride == 4 && rgb->cols % 4 == 0 && rgb->frame_stride % 16 == 0 && rgb->row_stride % 16 == 0) {
            const uintptr_t p = reinterpret_cast<uintptr_t>(rgb->ptr);
            if (rgb->ch_stride == 1 && p % 16 == 0) pxb = 4;
            else if (rgb->ch_stride == -1 && p % 16 == 2) pxb = -4;
        }
    }
    if (pxb && dense_rows(hsl, hue_only ? 1 : 3, hue_only ? 4 * (long long)dsize(hsl->dtype) : 16)) {
        // crops, pitched frames, batches of ROIs: 128-pixel row tiles (husl_rows.cuh)
        return with_flat_split(i, o, [&](const View2 &vi, const View2 &vo) -> int {
            const long long tiles = row_tiles(vi);
#define ROWS_GO1(TO, HUE, BLK, MB, SM, PXB, BGR) do { \
            static std::atomic<bool> attr_done[MAX_DEV]; \
            if ((rc = flt_attr(attr_done, *c, k_rows_fwd<TO, HUE, BLK, MB, PXB, BGR>, SM))) return rc; \
            k_rows_fwd<TO, HUE, BLK, MB, PXB, BGR><<<grid_for(tiles, BLK / 32, c->sms * MB), BLK, SM, st>>>(vi, vo); } while (0)
#define ROWS_GO(TO, HUE, BLK, MB, SM) do { \
            if (pxb == 3) ROWS_GO1(TO, HUE, BLK, MB, SM, 3, false); \
            else if (pxb == 4) ROWS_GO1(TO, HUE, BLK, MB, SM, 4, false); \
            else ROWS_GO1(TO, HUE, BLK, MB, SM, 4, true); } while (0)
            if (hsl->dtype == NPHUSL_F32) {
                if (hue_only) ROWS_GO(float, true, 512, 3, TAB_BYTES);
                else ROWS_GO(float, false, 512, 2, TAB_BYTES + 16 * Pack<float>::STAGE);
            } else {
                if (hue_only) ROWS_GO(double, true, 512, 3, TAB_BYTES);
                else ROWS_GO(double, false, 512, 1, TAB_BYTES + 16 * Pack<double>::STAGE);
            }
#undef ROWS_GO
#undef ROWS_GO1
            g_launches++;
            return check_launch("k_rows_fwd");
        });
    }
#define GO(TI, TO) do { if (hue_only) k_fwd_strided<TI, TO, true><<<grid, 256, 0, st>>>(i, o); \
                        else k_fwd_strided<TI, TO, false><<<grid, 256, 0, st>>>(i, o); } while (0)
    const bool o32 = hsl->dtype == NPHUSL_F32;
    if (rgb->dtype == NPHUSL_U8) { if (o32) GO(uint8_t, float); else GO(uint8_t, double); }
    else if (rgb->dtype == NPHUSL_F32) { if (o32) GO(float, float); else GO(float, double); }
    else { if (o32) GO(double, float); else GO(double, double); }
#undef GO
    g_launches++;
    return check_launch("k_fwd_strided");
}

int nphusl_cuda_husl_to_rgb_strided(const nphusl_view *hsl, const nphusl_view *rgb, int device, void *stream) {
    DeviceGuard guard;
    int rc;
    if ((rc = view_check(hsl, "husl_to_rgb_strided: hsl", false, true))) return rc;
    if ((rc = view_check(rgb, "husl_to_rgb_strided: rgb", true, true))) return rc;
    if (!same_extent(rgb, hsl)) return fail(NPHUSL_ERR_ARG, "husl_to_rgb_strided: extents differ");
    if (empty_view(rgb)) return NPHUSL_OK;
    DevCtx *c;
    if ((rc = ctx_get(device, &c))) return rc;
    const View2 i = view_dev(hsl), o = view_dev(rgb);
    const dim3 grid = grid_3d(*c, i);
    cudaStream_t st = (cudaStream_t)stream;
    if (rgb->dtype == NPHUSL_U8 && dense_rows(rgb, 3, 4) && dense_rows(hsl, 3, hsl->dtype == NPHUSL_F32 ? 16 : 32)) {
        return with_flat_split(i, o, [&](const View2 &vi, const View2 &vo) -> int {
            const int g = grid_for(row_tiles(vi), 8, c->sms * 4);
            if (hsl->dtype == NPHUSL_F32) k_rows_inv<float, 256, 4><<<g, 256, 0, st>>>(vi, vo);
            else k_rows_inv<double, 256, 4><<<g, 256, 0, st>>>(vi, vo);
            g_launches++;
            return check_launch("k_rows_inv");
        });
    }
#define GO(TI, TO) k_inv_strided<TI, TO><<<grid, 256, 0, st>>>(i, o)
    if (hsl->dtype == NPHUSL_F32) {
        if (rgb->dtype == NPHUSL_U8) GO(float, uint8_t); else if (rgb->dtype == NPHUSL_F32) GO(float, float); else GO(float, double);
    } else {
        if (rgb->dtype == NPHUSL_U8) GO(double, uint8_t); else if (rgb->dtype == NPHUSL_F32) GO(double, float); else GO(double, double);
    }
#undef GO
    g_launches++;
    return check_launch("k_inv_strided");
}

int nphusl_cuda_rgb_edit_rgb_strided(const nphusl_view *rgb_in, const nphusl_view *rgb_out,
                                     const nphusl_edit *edit, int device, void *stream) {
    DeviceGuard guard;
    int rc;
    if (!edit) return fail(NPHUSL_ERR_ARG, "rgb_edit_rgb_strided: null edit");
    if ((rc = view_check(rgb_in, "rgb_edit_rgb_strided: in", true, false))) return rc;
    if ((rc = view_check(rgb_out, "rgb_edit_rgb_strided: out", true, false))) return rc;
    if (!same_extent(rgb_in, rgb_out)) return fail(NPHUSL_ERR_ARG, "rgb_edit_rgb_strided: extents differ");
    if (empty_view(rgb_in)) return NPHUSL_OK;
    DevCtx *c;
    if ((rc = ctx_get(device, &c))) return rc;
    const View2 i = view_dev(rgb_in), o = view_dev(rgb_out);
    // 3: packed RGB rows on both sides; 4 / -4: RGBA / BGRA memory edited IN PLACE through rgba[..., :3]
    // (the same view on both sides: the fourth byte of each pixel is carried over from the input word)
    int pxb = 0;
    if (dense_rows(rgb_in, 3, 4) && dense_rows(rgb_out, 3, 4)) pxb = 3;
    else if (rgb_in->ptr == rgb_out->ptr && rgb_in->col_stride == 4 && rgb_out->col_stride == 4 &&
             rgb_in->ch_stride == rgb_out->ch_stride && rgb_in->row_stride == rgb_out->row_stride &&
             rgb_in->frame_stride == rgb_out->frame_stride && rgb_in->cols % 4 == 0 &&
             rgb_in->frame_stride % 16 == 0 && rgb_in->row_stride % 16 == 0) {
        const uintptr_t p = reinterpret_cast<uintptr_t>(rgb_in->ptr);
        if (rgb_in->ch_stride == 1 && p % 16 == 0) pxb = 4;
        else if (rgb_in->ch_stride == -1 && p % 16 == 2) pxb = -4;
    }
    if (pxb) {
        const EditShape sh(*edit);
        return with_flat_split(i, o, [&](const View2 &vi, const View2 &vo) -> int {
            const int g = grid_for(row_tiles(vi), EDIT_BLOCK / 32, c->sms * EDIT_MINB);
#define EDIT_GO1(EH, ES, RSL, PXB, BGR) do { \
            static std::atomic<bool> attr_done[MAX_DEV]; \
            if ((rc = flt_attr(attr_done, *c, k_rows_edit<EDIT_BLOCK, EDIT_MINB, EH, ES, RSL, PXB, BGR>, TAB_BYTES))) return rc; \
            k_rows_edit<EDIT_BLOCK, EDIT_MINB, EH, ES, RSL, PXB, BGR><<<g, EDIT_BLOCK, TAB_BYTES, (cudaStream_t)stream>>>(vi, vo, *edit); } while (0)
#define EDIT_GO(EH, ES, RSL) \
        if (sh.eh == EH && sh.es == ES && sh.rsl == RSL) { \
            if (pxb == 3) EDIT_GO1(EH, ES, RSL, 3, false); else if (pxb == 4) EDIT_GO1(EH, ES, RSL, 4, false); else EDIT_GO1(EH, ES, RSL, 4, true); }
            EDIT_GO(false, false, false) EDIT_GO(false, false, true) EDIT_GO(false, true, false) EDIT_GO(false, true, true)
            EDIT_GO(true, false, false) EDIT_GO(true, false, true) EDIT_GO(true, true, false) EDIT_GO(true, true, true)
#undef EDIT_GO
#undef EDIT_GO1
            g_launches++;
            return check_launch("k_rows_edit");
        });
    }
    k_edit_strided<<<grid_3d(*c, i), 256, 0, (cudaStream_t)stream>>>(i, o, *edit);
    g_launches++;
    return check_launch("k_edit_strided");
}

// ---- lookup tables ---------------------------------------------------------------------
// derived tables and reciprocal steps (husl_lut.cuh), once the tables and steps are installed
static int lut_finish(DevCtx &c) {
    Lut &l = c.lut;
    for (int i = 0; i < 3; ++i) l.p.r_ystep[i] = 1.0 / l.p.ystep[i];
    l.p.r_hstep = 1.0 / l.p.hstep;
    l.p.r_lstep = 1.0 / l.p.lstep;
    for (int i = 0; i < 3; ++i) {
        l.p.yorg[i] = i ? l.p.ythresh[i - 1] : 0.0;
        l.p.ybase[i] = (double)(l.seg * i);
    }
    const int n_light = 3 * l.seg;
    if (l.type == NPHUSL_LUT_U16) k_lut_prepare<uint16_t><<<(n_light + 255) / 256, 256>>>((const uint16_t *)l.light, n_light, l.light_scale, l.light_q);
    else if (l.type == NPHUSL_LUT_F32) k_lut_prepare<float><<<(n_light + 255) / 256, 256>>>((const float *)l.light, n_light, l.light_scale, l.light_q);
    else k_lut_prepare<double><<<(n_light + 255) / 256, 256>>>((const double *)l.light, n_light, l.light_scale, l.light_q);
    g_launches++;
    int rc = check_launch("k_lut_prepare");
    if (rc) return rc;
    CU(cudaDeviceSynchronize());
    l.ready = true;
    return NPHUSL_OK;
}

static int lut_alloc(DevCtx &c, int seg, int nh, int nl, int type, double light_scale, double chroma_scale) {
    Lut &l = c.lut;
    if (l.light) cudaFree(l.light);
    if (l.chroma) cudaFree(l.chroma);
    if (!l.linear) CU(cudaMalloc(&l.linear, 256 * sizeof(double)));
    if (l.light_q) cudaFree(l.light_q);
    l.light = l.chroma = nullptr;
    l.light_q = nullptr;
    l.ready = false;
    CU(cudaMalloc(&l.light_q, sizeof(double) * 3 * seg));
    CU(cudaMalloc(&l.light, lut_esize(type) * 3 * seg));
    CU(cudaMalloc(&l.chroma, lut_esize(type) * (size_t)nh * nl));
    l.seg = seg; l.nh = nh; l.nl = nl; l.type = type;
    l.light_scale = light_scale; l.chroma_scale = chroma_scale;
    l.p.linear = l.linear; l.p.light = l.light; l.p.chroma = l.chroma;
    l.p.light_q = l.light_q;
    l.p.seg = seg; l.p.nh = nh; l.p.nl = nl;
    l.p.light_scale = light_scale;
    l.p.sat_scale = 100 * chroma_scale;
    return NPHUSL_OK;
}

int nphusl_cuda_lut_generate_typed(int device, int light_seg, int nh, int nl, int table_type, int int_scale) {
    DeviceGuard guard;
    if (light_seg < 2 || nh < 2 || nl < 2 || light_seg > 21845 || nh > 65535 || nl > 65535)
        return fail(NPHUSL_ERR_ARG, "lut_generate: table sizes must fit 16 bits (LUT/light.py:27, LUT/chroma.py:27-28)");
    if (!lut_esize(table_type)) return fail(NPHUSL_ERR_ARG, "lut_generate: table_type must be NPHUSL_LUT_U16 / _F32 / _F64");
    if (table_type == NPHUSL_LUT_U16 && (int_scale < 1 || int_scale > 655))
        return fail(NPHUSL_ERR_ARG, "lut_generate: int_scale must keep 100 * int_scale inside uint16");
    // LUT/light.py:31-32: the int scale applies to integer table types only
    const double scale = table_type == NPHUSL_LUT_U16 ? (double)int_scale : 1.0;
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    if ((rc = lut_alloc(*c, light_seg, nh, nl, table_type, scale, scale))) return rc;
    Lut &l = c->lut;
    // index steps exactly as LUT/light.py:62-80 and LUT/chroma.py:63-65 compute them
    const double ysteps[3] = {0.070, 0.350, 1.0};
    double start = 0.0, seg_start[3];
    for (int i = 0; i < 3; ++i) {
        double stop = ysteps[i], step;
        if (i == 2) {
            step = (stop - start) / (light_seg - 1);
            stop += step;
            step = (stop - start) / light_seg;
        } else {
            step = (stop - start) / light_seg;
        }
        l.p.ystep[i] = step;
        seg_start[i] = start;
        if (i < 2) {   // "{:0.04f}".format(step), light.py:77
            char buf[64];
            std::snprintf(buf, sizeof buf, "%0.4f", stop);
            l.p.ythresh[i] = std::strtod(buf, nullptr);
        }
        start = stop;
    }
    l.p.hstep = 360.0 / (nh - 1);
    l.p.lstep = 100.0 / (nl - 1);
    // linear table: compute_linear((float)j/255.0) printed with "%f" (_linear_lookup.c:45-75);
    // the decimal rounding is a formatting step, done on the host from device values
    k_gen_linear<<<1, 256>>>(l.linear);
    g_launches++;
    if ((rc = check_launch("k_gen_linear"))) return rc;
    double lin[256];
    CU(cudaMemcpy(lin, l.linear, sizeof lin, cudaMemcpyDeviceToHost));
    for (int j = 0; j < 256; ++j) {
        char buf[64];
        std::snprintf(buf, sizeof buf, "%f", lin[j]);
        lin[j] = std::strtod(buf, nullptr);
    }
    CU(cudaMemcpy(l.linear, lin, sizeof lin, cudaMemcpyHostToDevice));
    const long long total = (long long)nh * nl;
#define GEN_GO(TT) do { \
        k_gen_light<TT><<<(3 * light_seg + 255) / 256, 256>>>((TT *)l.light, light_seg, seg_start[0], seg_start[1], seg_start[2], \
                                                             l.p.ystep[0], l.p.ystep[1], l.p.ystep[2], scale); \
        k_gen_chroma<TT><<<(int)((total + 255) / 256), 256>>>((TT *)l.chroma, nh, nl, l.p.hstep, l.p.lstep, scale); } while (0)
    if (table_type == NPHUSL_LUT_U16) GEN_GO(uint16_t);
    else if (table_type == NPHUSL_LUT_F32) GEN_GO(float);
    else GEN_GO(double);
#undef GEN_GO
    g_launches += 2;
    if ((rc = check_launch("k_gen_light / k_gen_chroma"))) return rc;
    return lut_finish(*c);
}

int nphusl_cuda_lut_generate(int device, int light_seg, int nh, int nl) {
    return nphusl_cuda_lut_generate_typed(device, light_seg, nh, nl, NPHUSL_LUT_U16, 100);
}

int nphusl_cuda_lut_upload_typed(int device, int table_type, double light_scale, double chroma_scale,
                                 const void *light, int light_seg,
                                 const double *y_idx_step3, const double *y_thresh2,
                                 const void *chroma, int nh, int nl,
                                 double h_idx_step, double l_idx_step, const double *linear256) {
    DeviceGuard guard;
    if (!light || !chroma || !y_idx_step3 || !y_thresh2 || !linear256 || light_seg < 2 || nh < 2 || nl < 2)
        return fail(NPHUSL_ERR_ARG, "lut_upload: null table or bad size");
    if (!lut_esize(table_type)) return fail(NPHUSL_ERR_ARG, "lut_upload: table_type must be NPHUSL_LUT_U16 / _F32 / _F64");
    if (!(light_scale > 0) || !(chroma_scale > 0)) return fail(NPHUSL_ERR_ARG, "lut_upload: scales must be positive");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    if ((rc = lut_alloc(*c, light_seg, nh, nl, table_type, light_scale, chroma_scale))) return rc;
    Lut &l = c->lut;
    const size_t es = lut_esize(table_type);
    CU(cudaMemcpy(l.light, light, es * 3 * light_seg, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(l.chroma, chroma, es * (size_t)nh * nl, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(l.linear, linear256, 256 * sizeof(double), cudaMemcpyHostToDevice));
    for (int i = 0; i < 3; ++i) l.p.ystep[i] = y_idx_step3[i];
    for (int i = 0; i < 2; ++i) l.p.ythresh[i] = y_thresh2[i];
    l.p.hstep = h_idx_step;
    l.p.lstep = l_idx_step;
    return lut_finish(*c);
}

int nphusl_cuda_lut_upload(int device, const uint16_t *light, int light_seg,
                           const double *y_idx_step3, const double *y_thresh2,
                           const uint16_t *chroma, int nh, int nl,
                           double h_idx_step, double l_idx_step, const double *linear256) {
    return nphusl_cuda_lut_upload_typed(device, NPHUSL_LUT_U16, 100, 100, light, light_seg, y_idx_step3, y_thresh2,
                                        chroma, nh, nl, h_idx_step, l_idx_step, linear256);
}

int nphusl_cuda_lut_download_typed(int device, void *light, void *chroma, double *linear256,
                                   int *dims4, double *steps7, double *scales2) {
    DeviceGuard guard;
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    std::lock_guard<std::mutex> lk(c->mu);
    Lut &l = c->lut;
    if (!l.ready) return fail(NPHUSL_ERR_NO_LUT, "lut_download: no tables installed");
    const size_t es = lut_esize(l.type);
    if (light) CU(cudaMemcpy(light, l.light, es * 3 * l.seg, cudaMemcpyDeviceToHost));
    if (chroma) CU(cudaMemcpy(chroma, l.chroma, es * (size_t)l.nh * l.nl, cudaMemcpyDeviceToHost));
    if (linear256) CU(cudaMemcpy(linear256, l.linear, 256 * sizeof(double), cudaMemcpyDeviceToHost));
    if (dims4) { dims4[0] = l.seg; dims4[1] = l.nh; dims4[2] = l.nl; dims4[3] = l.type; }
    if (steps7) {
        for (int i = 0; i < 3; ++i) steps7[i] = l.p.ystep[i];
        steps7[3] = l.p.ythresh[0]; steps7[4] = l.p.ythresh[1];
        steps7[5] = l.p.hstep; steps7[6] = l.p.lstep;
    }
    if (scales2) { scales2[0] = l.light_scale; scales2[1] = l.chroma_scale; }
    return NPHUSL_OK;
}

int nphusl_cuda_lut_download(int device, uint16_t *light, uint16_t *chroma, double *linear256,
                             int *dims4, double *steps7) {
    if (light || chroma) {
        DevCtx *c;
        int rc = ctx_get(device, &c);
        if (rc) return rc;
        if (c->lut.ready && c->lut.type != NPHUSL_LUT_U16)
            return fail(NPHUSL_ERR_ARG, "lut_download: the installed tables are not uint16; use nphusl_cuda_lut_download_typed");
    }
    return nphusl_cuda_lut_download_typed(device, light, chroma, linear256, dims4, steps7, nullptr);
}

// ---- memory / stream helpers --------------------------------------------------------------
int nphusl_cuda_malloc(void **ptr, size_t bytes, int device) {
    DeviceGuard guard;
    if (!ptr) return fail(NPHUSL_ERR_ARG, "malloc: null out pointer");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    CU(cudaMalloc(ptr, bytes ? bytes : 1));
    return NPHUSL_OK;
}

int nphusl_cuda_free(void *ptr, int device) {
    DeviceGuard guard;
    if (!ptr) return NPHUSL_OK;
    CU(cudaSetDevice(device));
    CU(cudaFree(ptr));
    return NPHUSL_OK;
}

int nphusl_cuda_host_alloc(void **ptr, size_t bytes) {
    if (!ptr) return fail(NPHUSL_ERR_ARG, "host_alloc: null out pointer");
    if (nphusl_cuda_device_count() <= 0) return fail(NPHUSL_ERR_NO_DEVICE, "no CUDA device available");
    CU(cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocPortable));
    return NPHUSL_OK;
}

int nphusl_cuda_host_free(void *ptr) {
    if (!ptr) return NPHUSL_OK;
    CU(cudaFreeHost(ptr));
    return NPHUSL_OK;
}

int nphusl_cuda_memcpy(void *dst, const void *src, size_t bytes, int kind, int device, void *stream) {
    DeviceGuard guard;
    static const cudaMemcpyKind k[4] = {cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice};
    if (kind < 0 || kind > 3) return fail(NPHUSL_ERR_ARG, "memcpy: bad kind");
    if (!bytes) return NPHUSL_OK;
    CU(cudaSetDevice(device));
    if (stream) CU(cudaMemcpyAsync(dst, src, bytes, k[kind], (cudaStream_t)stream));
    else CU(cudaMemcpy(dst, src, bytes, k[kind]));
    return NPHUSL_OK;
}

int nphusl_cuda_stream_create(void **stream, int device) {
    DeviceGuard guard;
    if (!stream) return fail(NPHUSL_ERR_ARG, "stream_create: null out pointer");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    cudaStream_t s;
    CU(cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));
    *stream = (void *)s;
    return NPHUSL_OK;
}

int nphusl_cuda_stream_destroy(void *stream, int device) {
    DeviceGuard guard;
    CU(cudaSetDevice(device));
    CU(cudaStreamDestroy((cudaStream_t)stream));
    return NPHUSL_OK;
}

int nphusl_cuda_stream_synchronize(void *stream, int device) {
    DeviceGuard guard;
    CU(cudaSetDevice(device));
    CU(cudaStreamSynchronize((cudaStream_t)stream));
    return NPHUSL_OK;
}

int nphusl_cuda_event_record(void **event, void *stream, int device) {
    DeviceGuard guard;
    if (!event) return fail(NPHUSL_ERR_ARG, "event_record: null out pointer");
    CU(cudaSetDevice(device));
    if (!*event) {
        cudaEvent_t e;
        CU(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        *event = (void *)e;
    }
    CU(cudaEventRecord((cudaEvent_t)*event, (cudaStream_t)stream));
    return NPHUSL_OK;
}

int nphusl_cuda_event_query(void *event, int *done) {
    if (!event || !done) return fail(NPHUSL_ERR_ARG, "event_query: null argument");
    const cudaError_t e = cudaEventQuery((cudaEvent_t)event);
    if (e == cudaSuccess) { *done = 1; return NPHUSL_OK; }
    if (e == cudaErrorNotReady) { cudaGetLastError(); *done = 0; return NPHUSL_OK; }
    return fail(NPHUSL_ERR_CUDA, std::string("cudaEventQuery: ") + cudaGetErrorString(e));
}

int nphusl_cuda_stream_wait_event(void *stream, void *event, int device) {
    DeviceGuard guard;
    if (!event) return fail(NPHUSL_ERR_ARG, "stream_wait_event: null event");
    CU(cudaSetDevice(device));
    CU(cudaStreamWaitEvent((cudaStream_t)stream, (cudaEvent_t)event, 0));
    return NPHUSL_OK;
}

int nphusl_cuda_event_destroy(void *event, int device) {
    DeviceGuard guard;
    if (!event) return NPHUSL_OK;
    CU(cudaSetDevice(device));
    CU(cudaEventDestroy((cudaEvent_t)event));
    return NPHUSL_OK;
}

int nphusl_cuda_graph_begin(void *stream, int device) {
    DeviceGuard guard;
    if (!stream) return fail(NPHUSL_ERR_ARG, "graph_begin: capture needs a stream of its own (not the NULL stream)");
    DevCtx *c;
    int rc = ctx_get(device, &c);
    if (rc) return rc;
    if ((rc = set_attrs(*c))) return rc;       // one-time kernel attributes are not stream work: do them before capturing
    CU(cudaStreamBeginCapture((cudaStream_t)stream, cudaStreamCaptureModeRelaxed));
    return NPHUSL_OK;
}

int nphusl_cuda_graph_end(void *stream, int device, void **graph_exec) {
    DeviceGuard guard;
    if (!stream || !graph_exec) return fail(NPHUSL_ERR_ARG, "graph_end: null argument");
    CU(cudaSetDevice(device));
    cudaGraph_t g = nullptr;
    CU(cudaStreamEndCapture((cudaStream_t)stream, &g));
    cudaGraphExec_t e = nullptr;
    const cudaError_t err = cudaGraphInstantiate(&e, g, 0);
    cudaGraphDestroy(g);
    if (err != cudaSuccess) {
        cudaGetLastError();
        return fail(NPHUSL_ERR_CUDA, std::string("cudaGraphInstantiate: ") + cudaGetErrorString(err));
    }
    *graph_exec = (void *)e;
    return NPHUSL_OK;
}

int nphusl_cuda_graph_launch(void *graph_exec, void *stream, int device) {
    DeviceGuard guard;
    if (!graph_exec) return fail(NPHUSL_ERR_ARG, "graph_launch: null graph");
    CU(cudaSetDevice(device));
    CU(cudaGraphLaunch((cudaGraphExec_t)graph_exec, (cudaStream_t)stream));
    return NPHUSL_OK;
}

int nphusl_cuda_graph_destroy(void *graph_exec, int device) {
    DeviceGuard guard;
    if (!graph_exec) return NPHUSL_OK;
    CU(cudaSetDevice(device));
    CU(cudaGraphExecDestroy((cudaGraphExec_t)graph_exec));
    return NPHUSL_OK;
}

int nphusl_cuda_device_synchronize(int device) {
    DeviceGuard guard;
    CU(cudaSetDevice(device));
    CU(cudaDeviceSynchronize());
    return NPHUSL_OK;
}

int nphusl_cuda_pointer_info(const void *ptr, int *loc_out, int *device_out, int *pinned_out) {
    cudaPointerAttributes a;
    int loc = NPHUSL_HOST, dev = -1, pinned = 0;
    if (cudaPointerGetAttributes(&a, ptr) == cudaSuccess) {
        if (a.type == cudaMemoryTypeDevice) { loc = NPHUSL_DEVICE; dev = a.device; }
        else if (a.type == cudaMemoryTypeManaged) { loc = NPHUSL_DEVICE; dev = a.device; pinned = 1; }
        else if (a.type == cudaMemoryTypeHost) { pinned = 1; }
    } else {
        cudaGetLastError();
    }
    if (loc_out) *loc_out = loc;
    if (device_out) *device_out = dev;
    if (pinned_out) *pinned_out = pinned;
    return NPHUSL_OK;
}

int nphusl_cuda_device_info(int device, char *name, size_t name_len, int *sm_count, size_t *total_mem) {
    cudaDeviceProp prop;
    CU(cudaGetDeviceProperties(&prop, device));
    if (name && name_len) {
        std::strncpy(name, prop.name, name_len - 1);
        name[name_len - 1] = 0;
    }
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (total_mem) *total_mem = prop.totalGlobalMem;
    return NPHUSL_OK;
}

}  // extern "C"
