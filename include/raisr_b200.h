/*
 * raisr_b200.h -- C-ABI of the B200-native RAISR hot path (libraisr_b200.so).
 *
 * This is the drop-in boundary for the one path of saturdaycoder/oclComputerVision that this
 * repository replaces: ClRaisr.upsample() in super_resolution/raisr.py together with the OpenCL
 * kernel super_resolution/raisr.cl.  Every entry point names the reference interface it stands in
 * for (paths relative to /root/reference/).  Plain pointers and sizes only -- no torch types.
 *
 * Conventions
 *   - every function returns 0 on success or a negative RAISR_E_* code; raisr_last_error() returns
 *     a thread-local human-readable message for the last failure on the calling thread.
 *   - a handle is bound to one CUDA device and is not thread-safe (one handle per thread / GPU),
 *     like the reference's single in-order queue (raisr.py:72).
 *   - there is NO CPU fallback: without a CUDA device raisr_create() fails with RAISR_E_CUDA.
 *   - images are 8-bit luma ("gray mode", raisr.py:97-100: CL_R / UNORM_INT8), row-major with a
 *     byte pitch; a batch is n_frames images back to back (frame stride = pitch * height).
 *   - dst dimensions must equal scale * src dimensions (the reference derives the output size from
 *     dst.shape and never checks it, raisr.py:86-89).
 */
#ifndef RAISR_B200_H
#define RAISR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct raisr_ctx raisr_t;

enum {
    RAISR_OK = 0,
    RAISR_E_ARG = -1,    /* bad argument (shape, scale, NULL pointer, ...)                      */
    RAISR_E_CUDA = -2,   /* CUDA runtime error or no usable device                              */
    RAISR_E_STATE = -3,  /* call order: e.g. upsample before set_filters for that scale         */
    RAISR_E_NOMEM = -4,  /* device / pinned allocation failed                                   */
    RAISR_E_UNSUPPORTED = -5 /* scale factor without a table ("not trained", raisr.py:90-94)    */
};

/* where the image pointers of a call live */
enum { RAISR_HOST = 0, RAISR_DEVICE = 1 };

/* Replaces ClRaisr.__init__ (raisr.py:62-83): device pick, kernel build, Gaussian constants.
 * n_angle/n_strength/n_coherence/filter_len are the -D macros of raisr.cl:5-19 (24/3/3/11);
 * filter_len must be 11 (the 9x9 Gaussian window is FILTER_LEN-2, raisr.cl:38). */
int raisr_create(raisr_t** h, int device, int n_angle, int n_strength, int n_coherence,
                 int filter_len);

/* Replaces `del` of the pyopencl objects; frees every device / pinned buffer the handle owns. */
void raisr_destroy(raisr_t* h);

/* Replaces the clFilters buffer (raisr.py:77-78,111; indexed at raisr.cl:316-317).  `table` is a
 * HOST pointer to float32 (n_angle, n_strength, n_coherence, scale*scale, 121), C order; it is
 * copied (and re-laid-out per pixel type) so the caller may free it.  n_floats is checked. */
int raisr_set_filters(raisr_t* h, int scale, const float* table, size_t n_floats);

/* What the filter kernel of the gray path multiplies by (no reference counterpart; the parity tests feed it to the
 * oracle).  table_out (may be NULL) receives the effective taps in the layout of raisr_set_filters: the table as
 * given for tap format 0 (fp32), every tap rounded to fp16 for format 1, and the 24-bit values for format 2 (sign,
 * exponent and 15 mantissa bits stored, see csrc/raisr_octet.cuh).  tap_format / b24_bound (may be NULL): the format
 * in use for this scale, and the bound on |output - fp32-tap output| of the 24-bit records: max over filters of
 * max(sum of the positive, sum of the negative tap errors) -- the patch values lie in [0,1]. */
int raisr_get_effective_filters(raisr_t* h, int scale, float* table_out, size_t n_floats, int* tap_format,
                                float* b24_bound);

/* Host-only helper (needs no device): packs one 11x11 filter (row-major taps) into the 384-byte 24-bit record the
 * filter kernel keeps in shared memory for this scale and returns the 121 tap values the kernel will decode from it.
 * Record layout: 24 chunks of 16 bytes; lane p (0..7) owns chunks p, p+8, p+16 = a circular stream of 48 bytes;
 * slot k (0..15) of the lane is the little-endian 32-bit word starting at stream byte 3k (wrapping at 48). */
int raisr_pack_taps_b24(const float* filter121, int scale, unsigned char record384[384], float effective121[121]);

/* Replaces clStreQ / clCoheQ (raisr.py:112-115; used at raisr.cl:301-314).  n_sq must be
 * n_strength-1 and n_cq n_coherence-1.  Defaults are {1e-4,1e-3} and {0.25,0.5}. */
int raisr_set_quantizers(raisr_t* h, const float* strength_q, int n_sq, const float* coherence_q,
                         int n_cq);

/* Run subsequent device work of this handle on `cuda_stream` (a cudaStream_t / CUstream; NULL =
 * the handle's own stream).  Stands in for the reference's CommandQueue (raisr.py:72). */
int raisr_set_stream(raisr_t* h, void* cuda_stream);

/* (further tuning keys, results unchanged: "filter_duo" 1 = both pixel types of an output row per CTA for s = 2 with
 * 24-bit records (default), 0 = one type per CTA; "eigen_in_filter" 1 = the prep kernel stores the structure tensor and
 * the s = 2 filter kernel does the eigen-solve / hash (default 0: measured slower); "resize_fast" 0 = generic resize
 * kernel only; "prep_ctas_per_sm" n > 0 = persistent prep grid (default 0).) */
/* Tuning knobs without a reference counterpart.  Keys: "filter_impl" (1 = octet kernel, default;
 * 0 = block kernel), "chunk_budget_bytes" (size of the per-launch upscaled-image scratch, default 208 MiB),
 * "overlap" (1 = experimental two-stream pipeline that co-schedules the prep kernel of the next chunk
 * with the filter kernel of the current one; default 0, slower on B200), "prep_impl"
 * (2 = packed-fp32 prep2_kernel, default; 1 = scalar prep_kernel; bit-identical results), "filter_pipe"
 * (1 = barrier-free tile pipeline in the octet filter kernel, default; 0 = CTA-wide barriers per tile; identical
 * results), "color_filter_impl" (2 = two planes per CTA in the colour path, default; 1 = one launch per plane;
 * identical results).
 * Semantics switches, both default 0 = the intended fp32 algorithm:
 *   "quirks"    1 = "as written": the three slips of the shipped kernel text are reproduced --
 *               ma accumulates gx*gy (raisr.cl:271), the coherence bucket compares L1 (raisr.cl:310),
 *               strength is left out of the hash (raisr.cl:316)
 *   "cheap_upscaler" 1 = stage 1 uses the reference's cubic_sample (raisr.cl:63-106, present in the kernel
 *               source but never called) instead of linear_sample; gray and colour paths (packed prep kernel only)
 *   "taps"      precision of the taps in the resident shared-memory table; arithmetic is fp32 in every mode.
 *               0 = fp32.  1 = fp16, as the reference's `(half)pf[...]` (raisr.cl:328) does.  2 = b24: sign, exponent
 *               and 15 mantissa bits (three quarters of the tap stream that bounds the filter kernel).  3 = auto
 *               (default): b24 when, for every filter of the table, the positive and the negative tap errors each sum
 *               to <= 5e-5, i.e. when the output provably stays within half the 1e-4 parity tolerance of the
 *               fp32-tap result (patch values lie in [0,1]), else fp32.
 *               Gray path only; the colour path keeps fp32 (or fp16) taps.  Re-packs the tables already set.
 *   "taps_fp16" older spelling: 1 = "taps" 1, 0 = "taps" 3. */
int raisr_set_option(raisr_t* h, const char* key, long long value);

/* Replaces ClRaisr.upsample (raisr.py:85-135) for gray frames: H2D copy, the fused RAISR kernels
 * (raisr.cl:108-338 with the intended semantics of SURVEY.md 8(a)), D2H copy, blocking wait.
 *   where = RAISR_HOST   : src/dst are host pointers (pageable or pinned); copies are inside the
 *                          call and chunks of frames are pipelined H2D / compute / D2H.
 *   where = RAISR_DEVICE : src/dst are device pointers on the handle's device; no copies.
 * ms[3] (may be NULL) receives {h2d_ms, kernel_ms, d2h_ms} like the reference's event list
 * (raisr.py:12-16,135); for RAISR_DEVICE h2d and d2h are 0.  The call blocks until dst is valid
 * unless where == RAISR_DEVICE and ms == NULL, in which case it only enqueues on the stream. */
int raisr_upsample_u8(raisr_t* h, const uint8_t* src, int sw, int sh, size_t src_pitch,
                      uint8_t* dst, int dw, int dh, size_t dst_pitch, int scale, int n_frames,
                      int where, float ms[3]);

/* Same, but dst is float32 in [0,1] (the value write_imagef would saturate and quantise,
 * raisr.cl:337); dst_pitch in bytes.  Used by the 1e-4 parity check. */
int raisr_upsample_f32(raisr_t* h, const uint8_t* src, int sw, int sh, size_t src_pitch,
                       float* dst, int dw, int dh, size_t dst_pitch, int scale, int n_frames,
                       int where, float ms[3]);

/* Colour form of ClRaisr.upsample, i.e. grayMode == 0 (raisr.py:101-104), the mode the reference's own
 * __main__ runs (raisr.py:139,163-164): interleaved 8-bit BGRA in and out (pitches in bytes, multiples
 * of 4).  RGB->YUV (raisr.py:20-25, raisr.cl:211-214), hash from Y, the hashed filter on all four planes
 * (raisr.cl:322-330), YUV->RGB (raisr.py:26-31, raisr.cl:333-336), saturating store.  The _f32 variant
 * writes the four clamped float components per pixel (memory order B,G,R,A) for the 1e-4 parity check. */
int raisr_upsample_bgra_u8(raisr_t* h, const uint8_t* src, int sw, int sh, size_t src_pitch,
                           uint8_t* dst, int dw, int dh, size_t dst_pitch, int scale, int n_frames,
                           int where, float ms[3]);
int raisr_upsample_bgra_f32(raisr_t* h, const uint8_t* src, int sw, int sh, size_t src_pitch,
                            float* dst, int dw, int dh, size_t dst_pitch, int scale, int n_frames,
                            int where, float ms[3]);

/* What the SHIPPED kernel computes (it returns after the cheap upscale, raisr.cl:219-230) and
 * what basic/interpolation.cl:17-71 (bilinear_lds) computes for one channel: align-corners
 * bilinear, u8 -> u8.  No filter table needed. */
int raisr_bilinear_u8(raisr_t* h, const uint8_t* src, int sw, int sh, size_t src_pitch,
                      uint8_t* dst, int dw, int dh, size_t dst_pitch, int scale, int n_frames,
                      int where, float ms[3]);

/* Stand-alone interpolation, SURVEY.md 8(f) row N2: replaces clUtility.bilinear / bilinear_lds / bicubic /
 * bicubic_lds (basic/interpolation.py:37-107, kernels basic/interpolation.cl:3-211).  Interleaved u8
 * image with 1 or 4 channels (the reference uses BGRA), any output size >= 2x2.
 * mode: 0 = bilinear_lds (align-corners), 1 = bicubic / bicubic_lds (Catmull-Rom), 2 = bilinear (the
 * normalised-coordinate CLK_FILTER_LINEAR sampler of bilinear_simple). */
int raisr_resize_u8(raisr_t* h, const uint8_t* src, int sw, int sh, size_t src_pitch, int channels,
                    uint8_t* dst, int dw, int dh, size_t dst_pitch, int mode, int n_frames, int where,
                    float ms[3]);

/* Histogram equalisation, SURVEY.md 8(f) row N4: the device side of clHistEq (histeq/eq_opencl.py:37-89,
 * kernels histeq/hist.cl:41-147).  8-bit gray images, `where` as above; the handle only supplies the
 * device and stream.
 *   ocv_hist_grid_u8            replaces clHistEq.histGrid (eq_opencl.py:37-51): 256-bin histogram of every
 *                               256x32 tile, hist_out = uint32 (h/32, w/256, 256); remainders are ignored
 *                               like the reference's integer divisions
 *   ocv_histeq_global_u8        replaces clHistEq.histeqGlobal (eq_opencl.py:53-68): dst = mapping[src]
 *   ocv_histeq_local_block_u8   replaces clHistEq.histeqLocalBlock (eq_opencl.py:70-89): bilinear blend of the
 *                               four neighbouring block mappings, `mappings` = float32 (ny, nx, 256) */
int ocv_hist_grid_u8(raisr_t* h, const uint8_t* img, int w, int hgt, size_t pitch, uint32_t* hist_out, int where,
                     float ms[3]);
int ocv_histeq_global_u8(raisr_t* h, const uint8_t* src, int w, int hgt, size_t src_pitch, uint8_t* dst,
                         size_t dst_pitch, const uint8_t* mapping256, int where, float ms[3]);
int ocv_histeq_local_block_u8(raisr_t* h, const uint8_t* src, int w, int hgt, size_t src_pitch, uint8_t* dst,
                              size_t dst_pitch, const float* mappings, int nx, int ny, int block_w,
                              int block_h, int where, float ms[3]);

/* Parity probe: the per-pixel quantities of raisr.cl:278-317 for one frame.  All outputs are
 * dense dh x dw arrays in `where` memory, any may be NULL: hash (int32, full index into the
 * filter table incl. pixel type), angle (theta in [0,pi)), l1 (strength), coherence, and the
 * upscaled image U (stage 1, float32). */
int raisr_debug_hash(raisr_t* h, const uint8_t* src, int sw, int sh, size_t src_pitch, int scale,
                     int32_t* hash, float* angle, float* l1, float* coherence, float* upscaled,
                     int where);

/* The same probe for the colour path (raisr.py:101-104): BGRA source, the quantities are derived from the Y
 * plane of the upscaled, colour-converted image (raisr.cl:211-214,235-317).  No U output. */
int raisr_debug_hash_bgra(raisr_t* h, const uint8_t* src, int sw, int sh, size_t src_pitch, int scale,
                          int32_t* hash, float* angle, float* l1, float* coherence, int where);

/* Row-band form for one very large image split across GPUs (no counterpart in the reference,
 * which is single-device; the coordinate map of raisr.cl:209 must use GLOBAL dimensions).
 * The caller passes a window of source rows [src_row0, src_row0 + src_rows) of a global image
 * of global_sh rows (pointer `src_rows_ptr` addresses row src_row0) and receives output rows
 * [dst_row0, dst_row0 + dst_rows) (pointer `dst` addresses row dst_row0).  The window must
 * contain every source row the band needs: use raisr_band_src_rows() to compute it (owned rows
 * plus <=3 halo rows per side, fetched from the neighbour GPU over NVLink P2P by the caller or
 * by raisr_p2p_copy).  Device pointers only. */
int raisr_upsample_band_u8(raisr_t* h, const uint8_t* src_rows_ptr, int sw, int global_sh,
                           size_t src_pitch, int src_row0, int src_rows, uint8_t* dst,
                           size_t dst_pitch, int dst_row0, int dst_rows, int scale);

/* First and last (inclusive) global source row needed to produce output rows
 * [dst_row0, dst_row0+dst_rows) of a global image with global_sh source rows. */
int raisr_band_src_rows(int global_sh, int scale, int dst_row0, int dst_rows, int* first,
                        int* last);

/* NVLink peer-to-peer helpers for the halo rows (one process per GPU: CUDA IPC).
 * raisr_ipc_export writes a 64-byte cudaIpcMemHandle for a device allocation of this process;
 * raisr_ipc_open maps a peer's handle and returns a device pointer valid in this process;
 * raisr_p2p_copy2d copies `rows` rows of `width_bytes` from a (peer) pointer into local memory
 * on the handle's stream. */
int raisr_ipc_export(const void* dev_ptr, unsigned char handle_out[64]);
int raisr_ipc_open(const unsigned char handle_in[64], void** dev_ptr_out);
int raisr_ipc_close(void* dev_ptr);
int raisr_p2p_copy2d(raisr_t* h, void* dst, size_t dst_pitch, const void* src, size_t src_pitch,
                     size_t width_bytes, size_t rows);

/* Stream-ordered hand-shake between the GPUs of a row-band job, no host synchronisation and no collective: a
 * 32-bit sequence word in device memory (normally inside an allocation the peers have mapped with raisr_ipc_open).
 * raisr_flag_set enqueues "publish `value`" after everything already enqueued on the handle's stream (with a
 * system-scope fence, so rows copied before it are visible to peers that see the value); raisr_flag_wait enqueues
 * "do not start later work of this stream before *flag has reached `value`" (wrap-safe comparison).  A wait that
 * is not satisfied within timeout_ms gives up; the next raisr_sync() then returns RAISR_E_STATE. */
int raisr_flag_set(raisr_t* h, void* dev_flag, unsigned value);
int raisr_flag_wait(raisr_t* h, const void* dev_flag, unsigned value, int timeout_ms);

/* 2-D copy on the handle's stream: direction 0 = host to device, 1 = device to host, 2 = device to device (also
 * from peer memory, like raisr_p2p_copy2d).  Asynchronous; host memory should be pinned. */
int raisr_copy2d(raisr_t* h, void* dst, size_t dst_pitch, const void* src, size_t src_pitch, size_t width_bytes,
                 size_t rows, int direction);

/* CUDA-event stopwatch on the handle's stream (the reference times its queue with OpenCL profiling events,
 * raisr.py:12-16): raisr_timer_mark records slot 0..7; raisr_timer_elapsed_ms waits for slot_to and returns the
 * device time between the two marks. */
int raisr_timer_mark(raisr_t* h, int slot);
int raisr_timer_elapsed_ms(raisr_t* h, int slot_from, int slot_to, float* ms);

/* Plain device allocations on the handle's device (cudaMalloc, so they can be exported with
 * raisr_ipc_export; framework caching allocators hand out sub-blocks that cannot). */
int raisr_dev_alloc(raisr_t* h, void** p, size_t bytes);
int raisr_dev_free(raisr_t* h, void* p);

/* Pinned host memory for the HOST path (stands in for mem_flags.USE_HOST_PTR, raisr.py:99-115). */
int raisr_host_alloc(void** p, size_t bytes);
int raisr_host_free(void* p);
/* Page-lock memory the caller already owns (e.g. the numpy arrays a reference-style loop passes to
 * upsample() again and again, raisr.py:166-182) so that the HOST path copies at full PCIe speed.
 * The range must be unregistered before it is freed. */
int raisr_host_register(void* p, size_t bytes);
int raisr_host_unregister(void* p);

/* Block until all work enqueued by this handle is complete. */
int raisr_sync(raisr_t* h);

/* Introspection used by bench.py: number of kernels this handle has launched so far, the
 * device's SM count and max SM clock (kHz), and the durations (ms) of the two kernels of the last
 * raisr_upsample_* call that asked for ms[]: prep (upscale+hash) and filter (gather-dot). */
long long raisr_launch_count(const raisr_t* h);
int raisr_device_info(const raisr_t* h, int* sm_count, int* sm_clock_khz, char* name, int name_len);
int raisr_last_kernel_ms(const raisr_t* h, float* prep_ms, float* filter_ms);

/* Register-only FFMA micro-benchmark: measured FP32 TFLOP/s of the handle's device (the FFMA
 * roofline denominator of SURVEY.md 8(d) is reported against both this and 2*128*SMs*clock). */
int raisr_measure_ffma_tflops(raisr_t* h, float* tflops);

const char* raisr_last_error(void);
const char* raisr_version(void);

#ifdef __cplusplus
}
#endif
#endif /* RAISR_B200_H */
