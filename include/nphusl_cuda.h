/*
 * nphusl_cuda.h -- C ABI of the B200-native HUSL conversion backend
 * (libnphusl_cuda.so, built from husl-numpy_b200/csrc with nvcc for sm_100a).
 *
 * This is the drop-in boundary for the hot path of TadLeonard/husl-numpy.  The
 * reference's only native entry point is
 *
 *     double *rgb_to_husl_nd(uint8_t *rgb, size_t size);      nphusl/_simd.h:4-5
 *
 * (callee mallocs, caller frees, errors call exit(): nphusl/_simd.c:87-124),
 * wrapped by nphusl/_simd_opt.pyx:20-39; HUSL->RGB and hue-only exist only as
 * Cython kernels (nphusl/_cython_opt.pyx:96-246).  The functions below replace
 * those three backend callables (_rgb_to_husl, _husl_to_rgb, _rgb_to_hue) with:
 *
 *   - caller-owned buffers (plain pointers + sizes, no torch / numpy types);
 *   - an int status (0 = ok) and a thread-local message instead of exit();
 *   - host OR device pointers on either side: device-resident frames
 *     (__cuda_array_interface__ / torch tensors) never touch the host, host
 *     arrays are streamed through pinned staging buffers in chunks;
 *   - an explicit device index and CUDA stream (device-pointer calls are
 *     asynchronous on that stream; host-pointer calls return when the result is
 *     in the caller's buffer).
 *
 * There is no CPU fallback anywhere behind this header: every conversion runs
 * in the CUDA kernels or fails with a non-zero status.
 *
 * Reference-side binding: see INTEGRATION.md (ctypes stub for nphusl/_cuda.py).
 */
#ifndef NPHUSL_CUDA_H
#define NPHUSL_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define NPHUSL_CUDA_ABI_VERSION 1

/* status codes */
enum {
    NPHUSL_OK = 0,
    NPHUSL_ERR_ARG = 1,        /* bad dtype / mode / null pointer */
    NPHUSL_ERR_CUDA = 2,       /* a CUDA runtime call failed (see last_error) */
    NPHUSL_ERR_NO_DEVICE = 3,  /* no usable GPU */
    NPHUSL_ERR_NO_LUT = 4      /* LUT mode requested before lut_generate/lut_upload */
};

/* element types of the image buffers */
enum { NPHUSL_U8 = 0, NPHUSL_F32 = 1, NPHUSL_F64 = 2 };
/* where a buffer lives */
enum { NPHUSL_HOST = 0, NPHUSL_DEVICE = 1 };
/* RGB -> HUSL arithmetic:
 *   EXACT      closed-form max chroma, atan2 hue, cube-root lightness
 *              (== reference NumPy/Cython float64 path within 1e-3 abs)
 *   LUT        reference default _simd.c build: light LUT + rounded chroma LUT +
 *              atan approximations, IEEE double in the reference's operation
 *              order (bit-exact vs _simd.c compiled without -ffast-math)
 *   LUT_INTERP same with -DINTERPOLATE_CHROMA (bilinear chroma LUT) */
enum { NPHUSL_EXACT = 0, NPHUSL_LUT = 1, NPHUSL_LUT_INTERP = 2 };

/* ---- library / device ---------------------------------------------------- */
int nphusl_cuda_abi_version(void);
/* number of CUDA devices (0 when there is no driver / GPU) */
int nphusl_cuda_device_count(void);
/* message of the last failing call on this thread ("" if none) */
const char *nphusl_cuda_last_error(void);
/* kernels launched by this library since load (all threads); the bench reports
 * the difference across its timed region as "gpu_launches" */
unsigned long long nphusl_cuda_launch_count(void);
/* streaming chunk (pixels) used for host-pointer calls; 0 restores the default */
int nphusl_cuda_set_chunk_pixels(size_t pixels);
/* Per calling thread.  0 (default): a call with a host buffer returns when the
 * result is in the caller's memory, like the reference's synchronous functions.
 * 1: when EVERY host buffer of a call is pinned (nphusl_cuda_host_alloc /
 * cudaHostAlloc), the call only enqueues its copies and kernels and makes
 * `stream` wait for them; the caller synchronises `stream` before touching the
 * buffers.  Lets H2D of one batch overlap D2H of the previous one. */
int nphusl_cuda_set_host_async(int enabled);

/* ---- the hot path ---------------------------------------------------------- */

/* RGB -> HUSL.  Replaces rgb_to_husl_nd (_simd.h:4-5, _simd.c:87-112) and
 * _cython_opt._rgb_to_husl (_cython_opt.pyx:22-93).
 *   rgb       n_pixels * channels elements, C-contiguous.  channels = 3:
 *             interleaved R,G,B; channels = 1: grayscale (the value is used for
 *             R, G and B: transform.from_grayscale, transform.py:210-215);
 *             channels = 4: interleaved R,G,B,A, colour premultiplied by alpha in
 *             float (transform.reshape_rgba_input, transform.py:221-230).
 *             NPHUSL_U8 holds 0..255, NPHUSL_F32/F64 hold [0,1].
 *   hsl       n_pixels * 3 elements (H,S,L interleaved), NPHUSL_F32 or F64.
 *   mode      NPHUSL_EXACT | NPHUSL_LUT | NPHUSL_LUT_INTERP (LUT modes: uint8
 *             RGB in, F64 out only, like the reference).
 *   stream    cudaStream_t (NULL = default stream) on `device`. */
int nphusl_cuda_rgb_to_husl(const void *rgb, int rgb_dtype, int rgb_loc, int channels,
                            void *hsl, int hsl_dtype, int hsl_loc,
                            size_t n_pixels, int mode, int device, void *stream);

/* RGB -> hue only (degrees, [0,360]).  Replaces _cython_opt._rgb_to_hue
 * (_cython_opt.pyx:96-156).  hue: n_pixels elements, F32 or F64. */
int nphusl_cuda_rgb_to_hue(const void *rgb, int rgb_dtype, int rgb_loc, int channels,
                           void *hue, int hue_dtype, int hue_loc,
                           size_t n_pixels, int device, void *stream);

/* HUSL -> RGB.  Replaces _cython_opt._husl_to_rgb (_cython_opt.pyx:177-246) and,
 * for rgb_dtype == NPHUSL_U8, the to_rgb_int tail fused in (transform.py:16-20:
 * round-half-even of c*255, abs, uint8 wrap; no clamp).
 *   hsl       n_pixels * 3, F32 or F64
 *   rgb       n_pixels * 3, U8 (0..255) or F32/F64 (float RGB in [0,1]) */
int nphusl_cuda_husl_to_rgb(const void *hsl, int hsl_dtype, int hsl_loc,
                            void *rgb, int rgb_dtype, int rgb_loc,
                            size_t n_pixels, int device, void *stream);

/* ---- callers' side of the path, on the GPU (no reference equivalent) --------- */

/* Planar HUSL: three separate planes of n_pixels elements (H[], S[], L[]) instead
 * of interleaved triplets -- what reference callers slice out with hsl[..., 0] and
 * hsl[..., 2] (README.md:77-98).  uint8 interleaved RGB on the other side.  DEVICE
 * pointers only; planes F32 or F64.  Asynchronous on `stream`. */
int nphusl_cuda_rgb_to_husl_planar(const void *rgb, void *h, void *s, void *l, int hsl_dtype,
                                   size_t n_pixels, int device, void *stream);
int nphusl_cuda_husl_planar_to_rgb(const void *h, const void *s, const void *l, int hsl_dtype,
                                   void *rgb, size_t n_pixels, int device, void *stream);

/* Channels-first images (torch / vision layout): separate uint8 planes R[], G[],
 * B[] on the RGB side as well.  DEVICE pointers only. */
int nphusl_cuda_rgb_planes_to_husl_planes(const void *r, const void *g, const void *b,
                                          void *h, void *s, void *l, int hsl_dtype,
                                          size_t n_pixels, int device, void *stream);
int nphusl_cuda_husl_planes_to_rgb_planes(const void *h, const void *s, const void *l, int hsl_dtype,
                                          void *r, void *g, void *b,
                                          size_t n_pixels, int device, void *stream);

/* Fused to_husl -> select -> edit -> to_rgb on uint8 pixels (the pipeline of
 * README.md:77-98 in one pass, 6 bytes of memory traffic per pixel).
 * A pixel is selected when H lies on the arc [h_lo, h_hi] (h_lo > h_hi: the arc
 * through 0), S in [s_lo, s_hi] and L in [l_lo, l_hi]; invert != 0 selects the
 * complement (use -INFINITY / +INFINITY for an unbounded range: S may exceed
 * 100 by an ulp on the gamut boundary).  Selected pixels get H' = H*h_mul + h_add wrapped to [0, 360),
 * S' = clamp(S*s_mul + s_add, 0, 100), L' = clamp(L*l_mul + l_add, 0, 100); a channel whose
 * (mul, add) is (1, 0) is left untouched (not clamped).
 * Results equal husl_to_rgb(edit(rgb_to_husl(rgb -> F32))) bit for bit. */
typedef struct nphusl_edit {
    float h_lo, h_hi, s_lo, s_hi, l_lo, l_hi;
    int invert;
    float h_mul, h_add, s_mul, s_add, l_mul, l_add;
} nphusl_edit;
/* rgb_in / rgb_out: n_pixels * 3 uint8, host or device (may alias each other). */
int nphusl_cuda_rgb_edit_rgb(const void *rgb_in, int in_loc, void *rgb_out, int out_loc,
                             size_t n_pixels, const nphusl_edit *edit, int device, void *stream);

/* The three-call form of the same pipeline, for callers that keep the HUSL array (README.md:77-98:
 * hsl = to_husl(img); <edit hsl>; out = to_rgb(hsl)):
 *  - husl_edit_to_rgb: husl_to_rgb with the edit applied to every pixel as it is loaded -- the HUSL
 *    array is read once, nothing is written back to it; equals husl_to_rgb(husl_edit(hsl)).  Same
 *    buffer rules as nphusl_cuda_husl_to_rgb (host or device on either side).
 *  - husl_edit: the edit written into a HUSL array (hsl_out may be hsl_in); DEVICE pointers.
 * The edit arithmetic is float32 on the float32 value of each channel (what the fused kernel applies
 * in registers); unselected pixels and channels whose (mul, add) is (1, 0) are passed on bit for bit. */
int nphusl_cuda_husl_edit_to_rgb(const void *hsl, int hsl_dtype, int hsl_loc, const nphusl_edit *edit,
                                 void *rgb, int rgb_dtype, int rgb_loc,
                                 size_t n_pixels, int device, void *stream);
int nphusl_cuda_husl_edit(const void *hsl_in, void *hsl_out, int hsl_dtype, size_t n_pixels,
                          const nphusl_edit *edit, int device, void *stream);

/* A banded map, the other HUSL-space edit of the reference's callers (examples.py:56-103: hue_watermelon
 * cuts the hue circle into bands of 45 degrees and paints them alternately green and pink, melonize does
 * the same over lightness bands and saturates a stripe around every band centre).  The range [0, total)
 * of channel `src` is cut the way chunk(total, width) cuts it (transform.py:300-308): band k covers the
 * OPEN interval (k*width, min((k+1)*width, total)) -- a value exactly on a band edge, below 0 or at / above
 * `total` belongs to no band and the pixel is passed on as it is, like the callers'
 * np.logical_and(x > low, x < high).  A pixel in an even band gets H = h_even, in an odd band H = h_odd
 * (examples.py:66-68: is_odd = low % (2*chunksize)).  If s_half > 0, a pixel whose `src` value lies strictly
 * within s_half of a band centre (low + high)/2 additionally gets S = s_value (examples.py:97-99).  The
 * selections use the pixel's ORIGINAL value of `src`; band edges and centres are formed in float32
 * (k*width, 0.5*(low + high), centre -+ s_half), exact for the integer parameters the callers use.
 * width > 0, total > 0, at most 2^20 bands. */
enum { NPHUSL_SRC_H = 0, NPHUSL_SRC_S = 1, NPHUSL_SRC_L = 2 };
typedef struct nphusl_bands {
    int src;                         /* NPHUSL_SRC_H | _S | _L: the channel the bands are cut from */
    float width, total;              /* chunk(total, width) */
    float h_even, h_odd;             /* the hue painted into even / odd bands */
    float s_half, s_value;           /* s_half > 0: S = s_value within s_half of a band centre */
} nphusl_bands;
/* husl_to_rgb with the banded map applied to every pixel as it is loaded (melonize renders ~200 frames
 * from ONE HUSL array, examples.py:84-103: the array stays on the device untouched); same buffer rules as
 * nphusl_cuda_husl_to_rgb, rgb_dtype U8. */
int nphusl_cuda_husl_bands_to_rgb(const void *hsl, int hsl_dtype, int hsl_loc, const nphusl_bands *bands,
                                  void *rgb, int rgb_dtype, int rgb_loc,
                                  size_t n_pixels, int device, void *stream);
/* fused to_husl -> banded map -> to_rgb on uint8 pixels (hue_watermelon in one pass); host or device,
 * rgb_out may be rgb_in.  Equals husl_bands_to_rgb(rgb_to_husl(rgb -> F32)) bit for bit. */
int nphusl_cuda_rgb_bands_rgb(const void *rgb_in, int in_loc, void *rgb_out, int out_loc,
                              size_t n_pixels, const nphusl_bands *bands, int device, void *stream);

/* The same fused edit on channels-first uint8 planes (what nvJPEG and video
 * decoders hand out); DEVICE pointers only; outputs may alias the inputs. */
int nphusl_cuda_rgb_planes_edit(const void *r, const void *g, const void *b,
                                void *r_out, void *g_out, void *b_out,
                                size_t n_pixels, const nphusl_edit *edit, int device, void *stream);

/* ---- strided device images ---------------------------------------------------- */

/* A batch of `frames` (rows x cols) images of 3-element pixels in DEVICE memory,
 * described by BYTE strides (any sign): element c of pixel (y, x) of frame f is at
 *     ptr + f*frame_stride + y*row_stride + x*col_stride + c*ch_stride.
 * Covers cropped views img[y0:y1, x0:x1], channel slices rgba[..., :3], BGR seen
 * as RGB (negative ch_stride), row-padded (pitched) decoder frames and (3, H, W)
 * planes seen as (H, W, 3).  The reference makes such arrays contiguous with a
 * host copy before a backend sees them (the C-contiguous memoryview casts of
 * _simd_opt.pyx:36 and _cython_opt.pyx:22); here the kernels walk the view.
 * For a hue image (rgb_to_husl_strided with hue_only != 0) ch_stride is unused.
 * Views whose rows are dense and aligned -- col_stride = 3 elements with ch_stride = 1 element (or uint8
 * col_stride = 4 with ch_stride = +-1: RGBA / BGRA memory), cols % 4 == 0, pointer / row / frame strides
 * multiples of 4 bytes (uint8 RGB), 16 bytes (RGBA, float HUSL; 32 for float64 HUSL input) or 4 elements
 * (hue) -- run in 128-pixel row tiles; every other view one pixel per thread.  Same values either way. */
typedef struct nphusl_view {
    void *ptr;
    int dtype;                       /* NPHUSL_U8 | NPHUSL_F32 | NPHUSL_F64 */
    long long frames, rows, cols;
    long long frame_stride, row_stride, col_stride, ch_stride;
} nphusl_view;

/* rgb: U8 (0..255) or F32/F64 ([0,1]); hsl (or hue): F32/F64; same frames x rows x cols */
int nphusl_cuda_rgb_to_husl_strided(const nphusl_view *rgb, const nphusl_view *hsl, int hue_only,
                                    int device, void *stream);
/* hsl: F32/F64; rgb: U8 (to_rgb_int tail fused, transform.py:16-20) or F32/F64 in [0,1] */
int nphusl_cuda_husl_to_rgb_strided(const nphusl_view *hsl, const nphusl_view *rgb,
                                    int device, void *stream);
/* fused edit on a U8 view, e.g. one region of a frame, in place when in == out */
int nphusl_cuda_rgb_edit_rgb_strided(const nphusl_view *rgb_in, const nphusl_view *rgb_out,
                                     const nphusl_edit *edit, int device, void *stream);

/* ---- video-decoder surfaces (NV12) as the load stage ------------------------------ */

/* The reference's callers decode video on the CPU and hand RGB arrays to to_husl
 * (examples.py:117-134, moviepy / imageio readers).  A hardware decoder (NVDEC)
 * leaves NV12 in DEVICE memory: a pitched luma plane and a pitched plane of
 * interleaved half-resolution U,V pairs (one pair per 2 x 2 luma block).  These
 * entry points make the YUV -> RGB step the load stage of the HUSL kernels, so a
 * decoded frame is read once (1.5 B/px) instead of being converted to an RGB copy
 * first.  Colour conversion: integer, 16.16 fixed-point coefficients of the
 * BT.601 / BT.709 matrix, round half up, clamp to 0..255, chroma replicated over
 * its 2 x 2 block:
 *     R = clamp8((ky*(Y-yoff)              + rv*(V-128) + 32768) >> 16)
 *     G = clamp8((ky*(Y-yoff) - gu*(U-128) - gv*(V-128) + 32768) >> 16)
 *     B = clamp8((ky*(Y-yoff) + bu*(U-128)              + 32768) >> 16)
 * ky = round(65536*ys), rv = round(65536*cs*2(1-Kr)), gu = round(65536*cs*2Kb(1-Kb)/Kg),
 * gv = round(65536*cs*2Kr(1-Kr)/Kg), bu = round(65536*cs*2(1-Kb)); limited range:
 * ys = 255/219, cs = 255/224, yoff = 16; full range: ys = cs = 1, yoff = 0.
 * HUSL results equal rgb_to_husl(nv12_to_rgb(frame)) bit for bit. */
enum { NPHUSL_BT601 = 0, NPHUSL_BT709 = 1 };
typedef struct nphusl_nv12 {
    const void *y;                   /* DEVICE: height rows of width luma bytes, y_pitch apart */
    const void *uv;                  /* DEVICE: ceil(height/2) rows of ceil(width/2) U,V byte pairs, uv_pitch apart */
    long long y_pitch, uv_pitch;     /* bytes */
    int width, height;               /* luma size in pixels */
    int matrix;                      /* NPHUSL_BT601 | NPHUSL_BT709 */
    int full_range;                  /* 0: 16..235 / 16..240 (video), 1: 0..255 (JPEG) */
} nphusl_nv12;
/* rgb: DEVICE, height*width*3 uint8, C-contiguous */
int nphusl_cuda_nv12_to_rgb(const nphusl_nv12 *src, void *rgb, int device, void *stream);
/* hsl: DEVICE, height*width*3 (or height*width when hue_only) F32 / F64, C-contiguous */
int nphusl_cuda_nv12_to_husl(const nphusl_nv12 *src, void *hsl, int hsl_dtype, int hue_only,
                             int device, void *stream);
/* fused NV12 -> HUSL -> select + edit -> uint8 RGB (DEVICE, height*width*3) */
int nphusl_cuda_nv12_edit_rgb(const nphusl_nv12 *src, const nphusl_edit *edit, void *rgb,
                              int device, void *stream);

/* The encode side (what a hardware encoder takes): uint8 RGB -> NV12, and the whole video loop
 * NV12 -> HUSL -> select + edit -> NV12 in one pass (3 bytes of memory traffic per pixel).
 * `dst` describes the destination surface (its y / uv planes are WRITTEN; width and height must
 * be even; matrix and range may differ from the source's).  RGB -> YUV: integer, 16.16 coefficients,
 *     Y = clamp8((yr*R + yg*G + yb*B + 32768 + (yoff << 16)) >> 16)
 *     U = clamp8((ur*Ra + ug*Ga + ub*Ba + 32768 + (128 << 16)) >> 16), V likewise,
 * (Ra, Ga, Ba) = (sum of the 2 x 2 block + 2) >> 2 per channel; yr = round(65536*ys*Kr) ...,
 * ur = round(-65536*cs*0.5*Kr/(1-Kb)), ug = round(-65536*cs*0.5*Kg/(1-Kb)), ub = vr = round(65536*cs*0.5),
 * vg = round(-65536*cs*0.5*Kg/(1-Kr)), vb = round(-65536*cs*0.5*Kb/(1-Kr)); video range: ys = 219/255,
 * cs = 224/255, yoff = 16; full range: ys = cs = 1, yoff = 0.
 * nv12_edit_nv12 equals rgb_to_nv12(nv12_edit_rgb(src)) bit for bit; dst may be src itself. */
int nphusl_cuda_rgb_to_nv12(const void *rgb, const nphusl_nv12 *dst, int device, void *stream);
int nphusl_cuda_nv12_edit_nv12(const nphusl_nv12 *src, const nphusl_edit *edit, const nphusl_nv12 *dst,
                               int device, void *stream);

/* ---- lookup tables (optional LUT mode) -------------------------------------- */

/* Generate, ON THE DEVICE, the tables make_lookup_tables:8-9 produces with
 * LUT/light.py (3 x light_seg uint16, L*100, Y segments split at 0.07 / 0.35)
 * and LUT/chroma.py (nh x nl uint16, max chroma * 100), plus the 6-decimal
 * linear table of nphusl/_linear_lookup.c:9-42, and keep them on `device`. */
int nphusl_cuda_lut_generate(int device, int light_seg, int nh, int nl);
/* Install caller-provided tables instead (e.g. the reference's own). */
int nphusl_cuda_lut_upload(int device, const uint16_t *light, int light_seg,
                           const double *y_idx_step3, const double *y_thresh2,
                           const uint16_t *chroma, int nh, int nl,
                           double h_idx_step, double l_idx_step,
                           const double *linear256);
/* Copy the installed tables back to the host (any pointer may be NULL).
 * dims4 = {light_seg, nh, nl, 0}; steps7 = {Y_IDX_STEP_0..2, Y_THRESH_0..1,
 * H_IDX_STEP, L_IDX_STEP}. */
int nphusl_cuda_lut_download(int device, uint16_t *light, uint16_t *chroma,
                             double *linear256, int *dims4, double *steps7);

/* The same three with the table element type the reference generator takes as its first
 * argument (make_lookup_tables:4 `table_type`, LUT/light.py:21-32, LUT/chroma.py:20-33):
 *   NPHUSL_LUT_U16  integer tables holding round(value * int_scale) (default: uint16_t, scale 100)
 *   NPHUSL_LUT_F32 / _F64  float / double tables holding the value as the generated source spells
 *                   it ("{:7.4f}" light, "{:7.3f}" chroma), LIGHT_SCALE = CHROMA_SCALE = 1
 * LUT / LUT_INTERP modes are bit-exact against _simd.c built (strictly) with those tables. */
enum { NPHUSL_LUT_U16 = 0, NPHUSL_LUT_F32 = 1, NPHUSL_LUT_F64 = 2 };
int nphusl_cuda_lut_generate_typed(int device, int light_seg, int nh, int nl, int table_type, int int_scale);
/* light / chroma: arrays of the element type; light_scale / chroma_scale: LIGHT_SCALE / CHROMA_SCALE */
int nphusl_cuda_lut_upload_typed(int device, int table_type, double light_scale, double chroma_scale,
                                 const void *light, int light_seg,
                                 const double *y_idx_step3, const double *y_thresh2,
                                 const void *chroma, int nh, int nl,
                                 double h_idx_step, double l_idx_step,
                                 const double *linear256);
/* dims4 = {light_seg, nh, nl, table_type}; scales2 = {LIGHT_SCALE, CHROMA_SCALE} */
int nphusl_cuda_lut_download_typed(int device, void *light, void *chroma, double *linear256,
                                   int *dims4, double *steps7, double *scales2);

/* ---- memory / stream helpers (so a host needs no other CUDA binding) -------- */
int nphusl_cuda_malloc(void **ptr, size_t bytes, int device);
int nphusl_cuda_free(void *ptr, int device);
int nphusl_cuda_host_alloc(void **ptr, size_t bytes);       /* pinned */
int nphusl_cuda_host_free(void *ptr);
/* kind: 0 h2h, 1 h2d, 2 d2h, 3 d2d; asynchronous when stream != NULL and the
 * host side is pinned */
int nphusl_cuda_memcpy(void *dst, const void *src, size_t bytes, int kind, int device, void *stream);
int nphusl_cuda_stream_create(void **stream, int device);
int nphusl_cuda_stream_destroy(void *stream, int device);
int nphusl_cuda_stream_synchronize(void *stream, int device);
/* Events (timing disabled): the host side orders the reuse of device blocks across streams with them.
 * event_record creates the event when *event == NULL and records it on `stream`; event_query returns
 * NPHUSL_OK with *done = 1 when everything recorded before it has finished, 0 when not yet;
 * stream_wait_event makes later work on `stream` wait for the event. */
int nphusl_cuda_event_record(void **event, void *stream, int device);
int nphusl_cuda_event_query(void *event, int *done);
int nphusl_cuda_stream_wait_event(void *stream, void *event, int device);
int nphusl_cuda_event_destroy(void *event, int device);
/* CUDA graphs: per-frame loops are bound by launch / call overhead (a 1080p frame is 10 us of kernel time), so
 * a sequence of DEVICE-pointer calls issued on `stream` between graph_begin and graph_end is recorded instead of
 * executed, and graph_launch replays the whole sequence with one call.  The buffers the recorded calls named
 * must stay allocated; their CONTENTS are read at replay time.  Host-pointer calls cannot be recorded. */
int nphusl_cuda_graph_begin(void *stream, int device);
int nphusl_cuda_graph_end(void *stream, int device, void **graph_exec);
int nphusl_cuda_graph_launch(void *graph_exec, void *stream, int device);
int nphusl_cuda_graph_destroy(void *graph_exec, int device);
int nphusl_cuda_device_synchronize(int device);
/* loc_out: NPHUSL_HOST (pageable or pinned; *pinned_out says which) or NPHUSL_DEVICE */
int nphusl_cuda_pointer_info(const void *ptr, int *loc_out, int *device_out, int *pinned_out);
/* "NVIDIA B200", SM count, bytes of global memory */
int nphusl_cuda_device_info(int device, char *name, size_t name_len, int *sm_count, size_t *total_mem);

#ifdef __cplusplus
}
#endif
#endif /* NPHUSL_CUDA_H */
