/*
 * psg_b200.h -- C ABI of libpsgb200.so: the B200 (sm_100a) PSD / STI hot path of
 * PySpectrogram's drfProc.py.
 *
 * The reference has no FFI of its own: its boundary is the Python surface of drfProc.py
 * (drfview.py:89 `import drfProc as dp`).  Each entry point below replaces the arithmetic
 * of the reference lines it cites; the Python module pyspectrogram_b200.drfProc keeps the
 * reference's call signatures and binds these symbols with ctypes.
 *
 * Conventions
 *   - plain pointers and sizes, no C++/torch types; every function returns 0 on success or a
 *     negative psg_status; psg_last_error() gives a thread-local message for the last failure.
 *   - "dev" pointers are CUDA device pointers on the plan's device; "host" pointers are
 *     ordinary (preferably pinned) host memory.  The caller owns every buffer; the library
 *     owns only the plan's tables and scratch.
 *   - all device work is enqueued on the caller's stream (a cudaStream_t passed as void*;
 *     NULL = legacy default stream) and is asynchronous unless stated otherwise.  A plan may be
 *     used from one host thread at a time (it owns scratch); create one plan per worker thread.
 *   - there is no CPU fallback anywhere: without a usable sm_100 device the calls fail.
 *
 * Data model
 *   complex64 IQ element e(n, s) of sample n, sub-channel s lives at
 *       iq[ n * sample_stride + s * sub_stride ]                 (strides in complex elements)
 *   STI column c is the mean over k = 0..frames_per_col-1 of the periodograms of the frames that
 *   start at element offset  col_offset[c] + k * hop * sample_stride :
 *       P_c[j] = in_scale^2 / frames_per_col * sum_k | FFT_nfft( w/sum(w) * x_k ) [j] |^2
 *   stored fftshifted (output index (j + nfft/2) mod nfft), laid out  out[s][c][i]  (i fastest).
 *     Mode R (reference sti_proc_data as shipped, drfProc.py:364-403): frames_per_col = 1
 *     Mode A (per-bin averaging north_star names / read_sti reads for, drfProc.py:158):
 *             frames_per_col = nint, hop = nfft
 *     Mode S (proc_data, drfProc.py:406-453): frames_per_col = n_int, hop = nfft - nfft/8
 */
#ifndef PSG_B200_H
#define PSG_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PSG_ABI_VERSION 2

typedef struct psg_plan psg_plan; /* opaque */

typedef enum psg_status {
    PSG_OK = 0,
    PSG_ERR_ARG = -1,         /* bad argument (message says which) */
    PSG_ERR_UNSUPPORTED = -2, /* nfft outside [PSG_MIN_NFFT, PSG_MAX_NFFT] / layout not implemented */
    PSG_ERR_CUDA = -3,        /* CUDA runtime error (message carries cudaGetErrorString) */
    PSG_ERR_NODEVICE = -4,    /* no CUDA device, or device is not sm_100 */
    PSG_ERR_NOMEM = -5
} psg_status;

enum { PSG_WINDOW_KAISER = 0, PSG_WINDOW_BOXCAR = 1 };

/* Library / ABI version (PSG_ABI_VERSION). */
int psg_version(void);

/* Thread-local text of the last error on this thread ("" if none). */
const char* psg_last_error(void);

/* Number of CUDA devices visible, or a negative psg_status. */
int psg_device_count(void);

/*
 * Build a plan for one FFT length on one device: the fp32 table w[n]/sum(w) of the periodic
 * Kaiser window (drfProc.py:386 / :435 `sig.get_window(("kaiser", 1.7), nfft)`; scipy
 * windows/_windows.py:1318-1320, :2551), computed in float64 on the host, the spectrum scaling
 * 1/sum(w)^2 (scipy _spectral_py.py:2277) folded in, and the float64-accurate twiddle table.
 * nfft: any integer in [PSG_MIN_NFFT, PSG_MAX_NFFT] (the viewer allows any length, drfview.py:474-479).
 * Powers of two run the tuned kernels; 2^a 3^b 5^c up to 15000 a direct mixed-radix transform; every other
 * length Bluestein's algorithm on the GPU.
 */
#define PSG_MIN_NFFT 2
#define PSG_MAX_NFFT 1048576
int psg_plan_create(psg_plan** out, int nfft, int window_kind, double beta, int device);
int psg_plan_destroy(psg_plan* plan);
int psg_plan_nfft(const psg_plan* plan);

/* Copy the plan's window table (w/sum(w), fp32, nfft values) to a host buffer (for tests). */
int psg_plan_window(const psg_plan* plan, float* host_out);

/*
 * The fused hot path: frame -> window -> FFT -> |X|^2 -> mean over the bin's frames -> fftshift
 * -> (linear and/or 10*log10(p + eps)), one read of every sample, one write of every column.
 * Replaces sig.periodogram / welch (drfProc.py:387-396), fftshift (drfProc.py:398-399), the
 * spectrogram+mean loop of proc_data (drfProc.py:436-449) and the dB step (drfProc.py:308-310).
 *
 *   iq_dev          complex64 (interleaved re,im fp32) on the device, 8-byte aligned.  With
 *                   sample_stride == 1 and a 16-byte aligned iq_dev the TMA loader is used; it
 *                   fetches 16-byte aligned spans, so the allocation must be readable up to the
 *                   next 16-byte boundary past its last sample (true of any cudaMalloc / torch
 *                   allocation).  Other layouts take the strided LDG loader.
 *   col_offset_dev  int64[ncol] element offsets of every column's first sample (frame index
 *                   table of DrfInput.read_sti, drfProc.py:158-159, times sample_stride)
 *   out_lin_dev / out_db_dev   fp32 [nsub][ncol][nfft], 16-byte aligned, either may be NULL (not both)
 *   eps             added before the log (drfProc.py:308: 1e-15)
 */
int psg_sti_run(psg_plan* plan, const void* iq_dev,
                int64_t sample_stride, int64_t sub_stride, int nsub,
                const int64_t* col_offset_dev, int ncol,
                int frames_per_col, int64_t hop,
                float in_scale, float eps,
                float* out_lin_dev, float* out_db_dev, void* cuda_stream);

/*
 * Raw integer IQ ingest: the same fused path reading Digital RF's native
 * complex int16 / int8 samples (interleaved re, im) instead of complex64, so the host-side cast to
 * complex64 and the divide by the full-scale reference that precede the path in the reference
 * (DrfInput.read, drfProc.py:124-129; get_ref, drfProc.py:182-201) disappear: pass in_scale = 1/ref.
 * Strides and offsets stay in complex elements.  psg_sti_run == psg_sti_run_typed(PSG_IQ_C64).
 */
enum { PSG_IQ_C64 = 0, PSG_IQ_CI16 = 1, PSG_IQ_CI8 = 2 };
int psg_sti_run_typed(psg_plan* plan, const void* iq_dev, int iq_type,
                      int64_t sample_stride, int64_t sub_stride, int nsub,
                      const int64_t* col_offset_dev, int ncol,
                      int frames_per_col, int64_t hop,
                      float in_scale, float eps,
                      float* out_lin_dev, float* out_db_dev, void* cuda_stream);

/*
 * The same with a bound on what the kernels may address: iq_elems complex elements are readable from iq_dev
 * (the recording's extent travels with it).  psg_sti_run[_typed] trust the offset table;
 * here a one-CTA kernel first clamps every col_offset into [0, iq_elems - column extent] (the table stays on the
 * device: no host round trip, ~2 us) and writes 1 to *oob_flag_dev when it had to move one, 0 otherwise, then
 * the transform runs on the clamped table -- a stale or wrong table gives a flagged image, not an out-of-bounds
 * read.  A column extent larger than iq_elems is refused on the host (PSG_ERR_ARG).  oob_flag_dev: int32 on the
 * device, valid when the stream has passed this call.
 */
int psg_sti_run_checked(psg_plan* plan, const void* iq_dev, int iq_type, int64_t iq_elems,
                        int64_t sample_stride, int64_t sub_stride, int nsub,
                        const int64_t* col_offset_dev, int ncol,
                        int frames_per_col, int64_t hop,
                        float in_scale, float eps,
                        float* out_lin_dev, float* out_db_dev, int32_t* oob_flag_dev, void* cuda_stream);
int psg_sti_host_typed(psg_plan* plan, const void* iq_host, int iq_type, int64_t iq_host_elems,
                       int64_t sample_stride, int64_t sub_stride, int nsub,
                       const int64_t* col_offset_host, int ncol,
                       int frames_per_col, int64_t hop,
                       float in_scale, float eps,
                       float* out_lin_host, float* out_db_host,
                       float* med_lin_host, float* med_db_host);

/*
 * Median over the time axis of a finished linear-power image (np.median(sxx, axis=1),
 * drfProc.py:401 / :451): img_dev is [nsub][ncol][nfft]; med_lin_dev / med_db_dev are
 * [nsub][nfft] (either may be NULL).  Even ncol gives the fp32 mean of the two middle values,
 * exactly like numpy.  Exact order statistic, no approximation: a bit-by-bit search for the key of the middle
 * rank (one warp per bin, keys in shared memory, the candidate set compacted every four bits; csrc/sti_kernels.cuh,
 * median_select_kernel), bit-identical to np.median for finite inputs.
 */
int psg_median_time(psg_plan* plan, const float* img_dev, int nsub, int ncol, int nfft,
                    float eps, float* med_lin_dev, float* med_db_dev, void* cuda_stream);

/*
 * Viewer-side reductions on the finished image, so that only what is
 * drawn crosses PCIe:
 *   psg_minmax_time   minimum and maximum over the time axis per (sub-channel, bin) -- the "min" and
 *                     "max" spectra proc_data's docstring promises next to the median
 *                     (drfProc.py:430-433; np.min / np.max semantics, NaN propagates).  Any output may
 *                     be NULL (not all four).
 *   psg_gather_bins   out[r][j] = clamp(img[r][idx[j]], clamp_lo, clamp_hi) for the rows = nsub*ncol
 *                     rows of an image (or rows = nsub for a median vector): idx is the viewer's
 *                     plotindices list -- frequency-range selection and decimation to at most 2^15
 *                     points (drfview.py:1005-1023) -- and the clamp is the colour-range clip of the
 *                     PNG export (drfview.py:1515-1518); clamp_lo > clamp_hi disables it.
 */
int psg_minmax_time(psg_plan* plan, const float* img_dev, int nsub, int ncol, int nfft, float eps,
                    float* min_lin_dev, float* max_lin_dev, float* min_db_dev, float* max_db_dev,
                    void* cuda_stream);
int psg_gather_bins(psg_plan* plan, const float* img_dev, int64_t rows, int nfft,
                    const int32_t* idx_dev, int count, float clamp_lo, float clamp_hi,
                    float* out_dev, void* cuda_stream);

/*
 * Host-buffer entry point (what drfProc.sti_proc_data / proc_data call): takes the IQ array in
 * host memory, copies the span the columns touch to the device -- in one piece up to
 * 1 GiB (psg_debug_set_host_chunk), beyond that in groups of consecutive columns, the copy of
 * group j+1 on a second stream overlapping the kernels of group j through two staging buffers, so
 * recordings larger than device memory work -- runs psg_sti_run + psg_median_time and copies the
 * results back.  Synchronous: results are valid on return.
 *
 *   iq_host         complex64 host array; element (n, s) of column c at
 *                   col_offset_host[c] + n*sample_stride + s*sub_stride
 *   iq_host_elems   number of complex elements addressable from iq_host (bounds check)
 *   out_lin_host / out_db_host   fp32 [nsub][ncol][nfft] or NULL
 *   med_lin_host / med_db_host   fp32 [nsub][nfft] or NULL
 */
int psg_sti_host(psg_plan* plan, const void* iq_host, int64_t iq_host_elems,
                 int64_t sample_stride, int64_t sub_stride, int nsub,
                 const int64_t* col_offset_host, int ncol,
                 int frames_per_col, int64_t hop,
                 float in_scale, float eps,
                 float* out_lin_host, float* out_db_host,
                 float* med_lin_host, float* med_db_host);

/*
 * Debugging and tuning knobs.  NOT part of the stable surface a viewer binds: they exist for the parity tests
 * (cross-checking one kernel against another) and for measurement scripts.  Each applies to the CALLING THREAD
 * only (thread-local state; a new thread starts with the defaults), so one of the viewer's worker threads
 * (drfview.py:177-178) cannot change the kernel another one launches.
 *
 *   psg_debug_set_force_generic   the simple radix-2 kernel (and the CTA-wide median) for subsequent calls
 *   psg_debug_set_variant(name)   prefer the named variant for its FFT length when the layout allows it; NULL or ""
 *       restores the automatic choice.  Path names for the large lengths: "r32" (three-pass radix-32 kernels,
 *       sti_r32.cuh: the default at 16384 / 32768 / 65536, and at 8192 with one frame per column), "split"
 *       (two-phase path through an HBM scratch, 8192..65536), "cluster_ldg" / "cluster" / "cluster_dsmem" (one CTA
 *       per 4096-point row), "whole" (four-pass whole-frame kernels: 8192 / 16384 in one CTA, 32768 / 65536 on
 *       clusters of 2 / 4), "whole_r2", "whole_r4", "whole_s2" / "whole_s8", "whole_f" (16 x 2 x 16 x 2 x 16,
 *       sti_whole16.cuh); "bluestein" / "bluestein_r2" for non powers of two; "mixed_rt" (the run-time mixed-radix
 *       kernel where a compile-time plan, sti_mixct.cuh, is the default)
 *   psg_debug_set_split_scratch   bytes of scratch per chunk of the split path (default cap 2 GiB)
 *   psg_debug_set_host_chunk      psg_sti_host streams spans above this many bytes in column chunks (default 1 GiB)
 *   psg_debug_set_mode_r_multi    0: one-frame-per-column launches use the one-column-per-CTA kernels
 *   psg_debug_set_items_per_slot  work items per resident CTA slot the column split aims for (default 24)
 */
int psg_debug_set_force_generic(int on);
int psg_debug_set_variant(const char* name);
int psg_debug_set_split_scratch(int64_t bytes);
int psg_debug_set_host_chunk(int64_t bytes);
int psg_debug_set_mode_r_multi(int on);
int psg_debug_set_items_per_slot(int n);
int psg_variant_count(void);
const char* psg_variant_name(int index);
int psg_variant_logn(int index);

/*
 * The plan's window table computed on the host only (no device needed): w[n]/sum(w) as fp32 from
 * fp64 math, sum(w) in *sum_out (may be NULL).  Same routine psg_plan_create uses.
 */
int psg_window_table(int nfft, int window_kind, double beta, float* host_out, double* sum_out);

/* Counters since load: kernels launched by this library (all kinds) -- for bench "gpu_launches". */
int64_t psg_launch_count(void);

/* Name of the kernel variant psg_sti_run would use for this plan ("fused16x16x16", ...). */
const char* psg_plan_variant(const psg_plan* plan);

#ifdef __cplusplus
}
#endif
#endif /* PSG_B200_H */
