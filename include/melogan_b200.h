/*
 * melogan_b200.h -- C ABI of the B200-native (sm_100a) Melo-GAN hot path.
 *
 * The reference (kaushik87599/Melo-GAN) is pure Python and has no FFI; its only stable
 * boundary is the Python module/class surface.  Each entry point below
 * therefore replaces the torch-level work behind one reference symbol, cited as
 * file:line into the reference tree, and is what a maintainer binds with ctypes
 * (INTEGRATION.md shows the stubs).  Plain pointers and sizes only: no torch types.
 *
 * Conventions
 *   - every `*_dev` / unmarked pointer is a DEVICE pointer on the current CUDA device;
 *     `*_host` pointers are host memory (pinned or pageable);
 *   - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream);
 *   - tensors are contiguous row-major; a "roll" is (T<=512 rows, 4) float32 with
 *     columns (pitch, velocity, duration, step) -- config/gan_config.yaml:43-44;
 *   - every function returns MG_OK (0) or a negative mg_status; nothing falls back
 *     to the CPU: without a usable sm_100 device the calls fail with MG_ERR_CUDA.
 */
#ifndef MELOGAN_B200_H
#define MELOGAN_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum mg_status {
    MG_OK = 0,
    MG_ERR_INVALID = -1,   /* bad argument (shape, null pointer, unsupported size)          */
    MG_ERR_CUDA = -2,      /* a CUDA runtime call failed; see mg_last_error()               */
    MG_ERR_NONFINITE = -3, /* the reference would raise (int(nan) / int(inf)) on this input */
    MG_ERR_STATE = -4      /* call order violated (e.g. backward before forward)            */
} mg_status;

/* Human-readable text of the last error on this thread ("" if none). */
const char* mg_last_error(void);
/* Library/ABI version and a string naming the compiled arch ("sm_100a"). */
int mg_abi_version(void);
const char* mg_build_info(void);

/* ------------------------------------------------------------------------------------------
 * N-1  save_piano_roll_to_midi row loop          reference src/gan/utils.py:95-161
 *      (velocity gate, pitch/velocity quantisation, scale snap, onset clock, onset/offset)
 *
 * rolls      (nrolls, nrows, 4) float32
 * bpm        as passed by the caller; clamped to [60,180] like utils.py:102
 * allowed_mask  bit k set <=> pitch class k is in the scale after the root shift
 *               (utils.py:119-121; mg_scale_mask() builds it from the SCALES table)
 * counts     (nrolls)        int32   notes emitted per roll, or -1 where the reference would raise
 * pitch, velocity (nrolls, nrows) uint8, start, end (nrolls, nrows) float64:
 *            the first counts[r] entries of row r are the pretty_midi.Note arguments in
 *            emission order, bit-exact (float32 results are widened exactly); the rest of
 *            the row is unspecified.
 * ---------------------------------------------------------------------------------------- */
int mg_extract_notes_gan(const float* rolls, long long nrolls, int nrows, double bpm,
                         uint32_t allowed_mask, int32_t* counts, uint8_t* pitch, uint8_t* velocity,
                         double* start, double* end, void* stream);

/* Same, with HOST buffers: copies in, runs, copies out on an internal stream, synchronises.
 * Returns MG_ERR_NONFINITE if any counts[r] == -1 (outputs of the other rolls are valid). */
int mg_extract_notes_gan_host(const float* rolls_host, long long nrolls, int nrows, double bpm,
                              uint32_t allowed_mask, int32_t* counts_host, uint8_t* pitch_host,
                              uint8_t* velocity_host, double* start_host, double* end_host);

/* N-2  tools/roll_to_midi.py:10-21 row loop: one note per row, absolute start times.
 * pitch, velocity (nrolls*nrows) uint8; start, end float64 (end = start + duration in float64).
 * status_dev (1 int32, may be NULL) is set to 1 if a NaN pitch was met (reference raises). */
int mg_extract_notes_abs(const float* rolls, long long nrolls, int nrows, uint8_t* pitch,
                         uint8_t* velocity, double* start, double* end, int32_t* status_dev, void* stream);
int mg_extract_notes_abs_host(const float* rolls_host, long long nrolls, int nrows, uint8_t* pitch_host,
                              uint8_t* velocity_host, double* start_host, double* end_host);

/* utils.py:14-26,119-121: mask of allowed pitch classes for a named scale and root key;
 * unknown names select the chromatic scale like SCALES.get(scale, SCALES['chromatic']). */
uint32_t mg_scale_mask(const char* scale_name, int root_key);

/* ------------------------------------------------------------------------------------------
 * A-10  torch.optim.Adam.step / AdamW.step       reference src/gan/train_gan.py:136-145,204,248
 *       (src/ae/train_ae.py:79, src/emotion_discriminator/train_ed.py:97 for the AdamW form)
 * One fused, vectorised launch over a flat parameter segment of n float32 values.
 *   m = lerp(m, g, 1-beta1); v = beta2*v + (1-beta2)*g*g;
 *   p -= lr/(1-beta1^t) * m / (sqrt(v)/sqrt(1-beta2^t) + eps)        (amsgrad off)
 * weight_decay: 0 for Adam; for AdamW (decoupled != 0) p *= 1 - lr*weight_decay first.
 * step (t, 1-based) is taken from `step` unless step_dev != NULL, in which case *step_dev
 * (int64 on the device) is incremented first and used -- this form is CUDA-graph safe.
 * grad_scale multiplies g on load (1.0 normally; 1/world for summed all-reduce results).
 * bf16_copy (may be NULL): also writes the updated parameters as bfloat16 (compute copy).
 * ---------------------------------------------------------------------------------------- */
int mg_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n,
                 double lr, double beta1, double beta2, double eps, double weight_decay, int decoupled,
                 float grad_scale, long long step, long long* step_dev, uint16_t* bf16_copy, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MELOGAN_B200_H */
