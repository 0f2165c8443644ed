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
 *     to the CPU: without a usable sm_100 device the calls fail with MG_ERR_CUDA;
 *   - ONE device and ONE calling thread per process (the deployment model is one process per
 *     GPU): kernel-selection switches, the packed-weight scratch ring, the launch counter and
 *     the probe state are process-global; contexts created on different devices of the same
 *     process, or calls from concurrent host threads, are not supported.  Streams: calls that
 *     share a context must be issued on one stream (or be ordered by the caller).
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

/* 8(f)-3  MIDIDataset.__getitem__ normalisation of raw rolls     reference src/ae/dataset.py:72-89,105
 * notes/out (nrolls, nrows, 4) float32 device buffers, rows (pitch, start, duration, velocity) in raw units (MIDI
 * numbers, beats); rows with pitch == -1 are padding and are copied.  pitch, velocity -> (x/128)*2-1 (velocity clipped to
 * [0,127] first), start /= max_start_beat (cfg MAX_START_BEAT, 100), duration /= max_duration_beat (MAX_DURATION_BEAT,
 * 20), then nan_to_num(0, 0, 0).  Bit-exact with the reference's float32 numpy arithmetic; out may alias notes. */
int mg_ae_normalize(const float* notes, float* out, long long nrolls, int nrows, double max_start_beat,
                    double max_duration_beat, void* stream);

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
/* Same update with torch.nn.utils.clip_grad_norm_(params, max_norm) fused in front of it (reference
 * src/ae/train_ae.py:121): the L2 norm of grad_scale * grad over the whole flat group is reduced on the device, the
 * gradient is scaled by min(1, max_norm / (norm + 1e-6)) inside the update; norm_out_dev (1 float, may be NULL) receives
 * the pre-clip norm.  No host synchronisation: CUDA-graph capturable. */
int mg_adam_step_clipped(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, long long n, double lr,
                         double beta1, double beta2, double eps, double weight_decay, int decoupled, float grad_scale,
                         float max_norm, float* norm_out_dev, long long step, long long* step_dev, uint16_t* bf16_copy,
                         void* stream);


/* ==========================================================================================
 * GAN training hot path (SURVEY.md 8a rows A-1 .. A-9).
 *
 * One opaque context owns every workspace (activations saved for backward, scratch) for a fixed
 * per-rank batch size; parameters, gradients and optimizer state stay in caller-owned device
 * memory (the host side keeps them as views of flat torch buffers) and are bound by pointer.
 * All tensors at this boundary are float32, contiguous, channels-last exactly as the reference
 * modules exchange them: notes (B, max_notes, note_dim), embeddings (B, embed_dim), ...
 * `precision` selects the arithmetic of the big contractions: 0 = fp32 on CUDA cores (parity
 * mode, 1e-5), 1 = bf16 operands on tcgen05 tensor cores with fp32 TMEM accumulation (1e-2).
 * ========================================================================================== */
typedef struct mg_gan mg_gan;

typedef struct mg_gan_config {
    int batch;          /* B: samples per step on this rank                                  */
    int precision;      /* 0 = fp32, 1 = bf16 tensor cores                                   */
    int max_notes;      /* gan_config.yaml MAX_NOTES (multiple of 8)                         */
    int note_dim;       /* NOTE_DIM (must be 4)                                              */
    int noise_dim;      /* NOISE_DIM                                                         */
    int latent_dim;     /* LATENT_DIM                                                        */
    int gen_hidden;     /* NoiseToLatent hidden width (512: train_gan.py:95-102 default)     */
    int numeric_dim;    /* NUMERIC_INPUT_DIM                                                 */
    int enc_hidden1;    /* ENCODER_HIDDEN[0]                                                 */
    int enc_hidden2;    /* ENCODER_HIDDEN[1]                                                 */
    int embed_dim;      /* ENCODER_OUT_DIM                                                   */
    int n_classes;      /* emotion classes (4)                                               */
    double enc_dropout; /* FeatureEncoder dropout p (0.2: constructor default)               */
    double lambda_gp;   /* LAMBDA_GP                                                         */
    double lambda_emotion; /* LAMBDA_EMOTION                                                 */
    double bn_momentum; /* 0.1                                                               */
    double bn_eps;      /* 1e-5                                                              */
    int cond_dim;       /* INTEGRATION_MODE 'conditioning': width of the AE latent appended to G's input
                           (models.py:99-100,121-123); 0 = 'warm_start'                       */
} mg_gan_config;

enum { MG_MOD_ENCODER = 0, MG_MOD_GENERATOR = 1, MG_MOD_CRITIC = 2, MG_MOD_EMOTION = 3 };

int mg_gan_create(const mg_gan_config* cfg, mg_gan** out);
void mg_gan_destroy(mg_gan* ctx);
long long mg_gan_workspace_bytes(const mg_gan* ctx);

/* Bind parameter (and optionally gradient) pointers of one module, in the reference's
 * state_dict order (buffers last):
 *   ENCODER   (8):  net.0.{weight,bias} net.1.{w,b} net.4.{w,b} net.7.{w,b}       feature_encoder.py:17-41
 *   GENERATOR (22): noise_to_latent.net.{0,2}.{w,b} decoder.pre.{0,2}.{w,b} decoder.deconv.0.{w,b}
 *                   deconv.1.{w,b} deconv.3.{w,b} deconv.4.{w,b} deconv.6.{w,b}   then the buffers
 *                   deconv.1.running_{mean,var} deconv.4.running_{mean,var}       models.py:20-64
 *   CRITIC    (10): conv.{0,2,4}.{w,b} fc.1.{w,b} real_fake.{w,b}                 models.py:140-156
 *   EMOTION   (32): for i in 0..3: encoder.conv.i.net.0.{w,b} net.1.{w,b,running_mean,running_var};
 *                   encoder.project.{w,b} classifier.net.{0,3}.{w,b} classifier.head.{w,b}
 * grads has the same order and length as the trainable prefix (8 / 18 / 10; EMOTION: the 24 trainable tensors,
 * i.e. the parameter list without the running statistics); NULL = no gradients. */
int mg_gan_bind(mg_gan* ctx, int module, void* const* params, int nparams, void* const* grads, int ngrads);
/* The frozen emotion discriminator's BatchNorm is folded once per bind; call again after loading new weights. */

/* A-1  FeatureEncoder.forward / backward          src/gan/feature_encoder.py:43-45
 * mask1 (B, enc_hidden1), mask2 (B, enc_hidden2): dropout keep-masks (1/0) when train != 0 (NULL with
 * train == 0 = eval).  backward adds the parameter gradients of d(emb) into the bound grad buffers. */
int mg_feature_encoder_forward(mg_gan* ctx, const float* numeric, const float* mask1, const float* mask2,
                               int train, float* emb_out, void* stream);
int mg_feature_encoder_backward(mg_gan* ctx, const float* demb, void* stream);

/* A-2..A-4  Generator.forward (warm_start: cat[noise, emb])   src/gan/models.py:108-130
 * train != 0: BatchNorm uses batch statistics and updates the bound running stats (momentum);
 * train == 0: running statistics.  notes_out (B, max_notes, 4), latent_out (B, latent_dim) may be NULL.
 * backward: dnotes (B, max_notes, 4) [+ dlatent] -> G parameter gradients (added) and demb_out (B, embed_dim). */
int mg_generator_forward(mg_gan* ctx, const float* noise, const float* emb, int train, float* notes_out,
                         float* latent_out, void* stream);
/* 'conditioning' mode (cond_dim > 0): encoder_latent (B, cond_dim) is the third block of G's input row
 * [noise | numeric embedding | encoder latent] (models.py:112-126).  The pointer is remembered and read by every later
 * generator forward / fused step of this context; it carries no gradient (the reference feeds detached AE latents). */
int mg_generator_set_condition(mg_gan* ctx, const float* encoder_latent);
int mg_generator_backward(mg_gan* ctx, const float* dnotes, const float* dlatent, float* demb_out, void* stream);

/* A-5  Discriminator.forward (WGAN critic)         src/gan/models.py:158-169
 * nsamples <= 3*batch; emb row of sample r is emb[r % batch] (NULL: no conditioning term).
 * backward: dscore (nsamples) -> optional parameter gradients (added), dnotes_out, demb_out (may be NULL). */
int mg_discriminator_forward(mg_gan* ctx, const float* notes, const float* emb, int nsamples, float* score_out,
                             void* stream);
int mg_discriminator_backward(mg_gan* ctx, const float* dscore, int param_grads, float* dnotes_out,
                              float* demb_out, void* stream);
/* Same plus the critic's parameter gradients (added into the bound grads); `notes` is the tensor the
 * preceding mg_discriminator_forward read (conv.0's weight gradient needs it). */
int mg_discriminator_backward_ex(mg_gan* ctx, const float* notes, const float* dscore, float* dnotes_out,
                                 float* demb_out, void* stream);

/* A-6 + A-7  critic loss, forward and backward in one call:
 *   loss_d = mean(D(fake)) - mean(D(real)) + lambda_gp * GP(real, fake, alpha)
 * (src/gan/train_gan.py:191-203 and compute_gradient_penalty src/gan/utils.py:75-90, including the
 * double backward through the critic).  alpha (B) is the per-sample interpolation weight.
 * Adds d(loss_d)/d(theta_D) into the bound critic grads; metrics_out (4 floats, device):
 * [loss_d, gp, mean D(real), mean D(fake)]. */
int mg_critic_loss_backward(mg_gan* ctx, const float* real, const float* fake, const float* emb,
                            const float* alpha, float* metrics_out, void* stream);

/* A-6 alone  compute_gradient_penalty (src/gan/utils.py:75-90): metrics_out[1] = GP (unweighted); adds
 * d(GP)/d(theta_D) -- the double backward through the critic -- into the bound critic grads. */
int mg_gradient_penalty(mg_gan* ctx, const float* real, const float* fake, const float* emb, const float* alpha,
                        float* metrics_out, void* stream);

/* A-8  EmotionDiscriminator.forward in eval mode + input gradient (frozen weights)
 *      src/emotion_discriminator/ed_model.py:147-165 */
int mg_emotion_forward(mg_gan* ctx, const float* notes, float* logits_out, void* stream);
int mg_emotion_backward_input(mg_gan* ctx, const float* dlogits, float* dnotes_out, int accumulate, void* stream);

/* A-13  EmotionDiscriminator in TRAIN mode (BASELINE config #3): BatchNorm batch statistics (running statistics
 *       updated), MLP dropout with caller-supplied keep-masks (B,256) and (B,128), full backward.
 *       reference src/emotion_discriminator/train_ed.py:61-74, ed_model.py:35-42,63-69,92-95.
 *       Bind the module with 24 gradient pointers: for i in 0..3 conv.i.net.0.{w,b} net.1.{w,b}; then project,
 *       classifier.net.0, classifier.net.3, classifier.head {w,b}.  backward adds into them; dnotes_out may be NULL.
 *       mg_cross_entropy: out2 = [mean CE, accuracy], dlogits = (softmax - onehot)/batch (may be NULL). */
int mg_emotion_train_forward(mg_gan* ctx, const float* notes, const float* mask1, const float* mask2, double dropout_p,
                             float* logits_out, void* stream);
int mg_emotion_train_backward(mg_gan* ctx, const float* notes, const float* dlogits, float* dnotes_out, void* stream);
int mg_cross_entropy(const float* logits, const long long* labels, int batch, int n_classes, float* dlogits, float* out2,
                     void* stream);

/* ------------------------------------------------------------------------------------------
 * A-12  VAE (BASELINE config #2): reference src/ae/model.py:4-148 forward, its backward, and the loss of
 *       src/ae/train_ae.py:35-51.  Own context (different widths from the GAN step).
 * params (42) in state_dict order: encoder.conv.{0,3,6}.{w,b} each followed by its BatchNorm conv.{1,4,7}.{w,b};
 *   encoder._linear.1.{w,b}; fc_mu.{w,b}; fc_log_var.{w,b}; decoder.pre.{0,2}.{w,b}; decoder.deconv.0.{w,b},
 *   deconv.1.{w,b}, deconv.3.{w,b}, deconv.4.{w,b}, deconv.6.{w,b}  (32 trainable), then running_{mean,var} of
 *   encoder.conv.{1,4,7} and decoder.deconv.{1,4} (10).  grads: the 32 trainable ones.
 * eps (B, latent) is the reparameterisation noise (torch.randn_like in the reference).
 * mg_vae_backward takes d(loss)/d(recon) (+ optional d/dz, d/dmu, d/dlog_var), adds parameter gradients.
 * mg_vae_loss_step = forward + vae_loss + backward; metrics_out = [total, recon MSE, KLD]. */
typedef struct mg_vae mg_vae;
int mg_vae_create(int batch, int max_notes, int latent_dim, int precision, mg_vae** out);
void mg_vae_destroy(mg_vae* ctx);
int mg_vae_bind(mg_vae* ctx, void* const* params, int nparams, void* const* grads, int ngrads);
int mg_vae_forward(mg_vae* ctx, const float* x, const float* eps, int train, float* recon_out, float* z_out,
                   float* mu_out, float* logvar_out, void* stream);
int mg_vae_backward(mg_vae* ctx, const float* x, const float* drecon, const float* dz, const float* dmu,
                    const float* dlogvar, void* stream);
/* Debug view of a named workspace buffer of the VAE context (tests compare intermediate activations / gradients). */
int mg_vae_buffer(mg_vae* ctx, const char* name, void** ptr, long long* nbytes);
int mg_vae_loss_step(mg_vae* ctx, const float* x, const float* eps, double beta, float* metrics_out, void* stream);

/* A-7 composite: the whole critic step body up to (not including) opt_D.step()
 *      src/gan/train_gan.py:185-203: E_num and G forward without grad (dropout on, BN batch stats,
 *      running stats updated), then mg_critic_loss_backward.  Caller zeroes the critic grads first. */
int mg_critic_step(mg_gan* ctx, const float* real, const float* numeric, const float* noise, const float* alpha,
                   const float* mask1, const float* mask2, float* metrics_out, void* stream);

/* A-9 composite: generator step body up to opt_G.step()   src/gan/train_gan.py:215-247
 *      loss_g = -mean(D(G(z))) + lambda_emotion * CE(ED(G(z)), labels); adds gradients of G and E_num.
 *      The critic's own (discarded) weight gradients of this backward are not computed.
 *      labels (B) int64; metrics_out (2 floats): [loss_g_adv, loss_g_emo]. */
int mg_generator_step(mg_gan* ctx, const float* numeric, const float* noise, const long long* labels,
                      const float* mask1, const float* mask2, float* metrics_out, void* stream);

/* Instrumentation used by bench.py: number of kernels this library has launched so far, and a probe that
 * brackets every launch of one kernel family (1 = SIMT tap-GEMM, 2 = SIMT wgrad, 3 = tcgen05 GEMM,
 * 4 = tcgen05 wgrad, 5 = note extraction, 6 = Adam) with CUDA events on the launching stream.
 * mg_probe_end: out[0] launches, out[1] total ms, out[2] algorithmic FLOPs, out[3] algorithmic bytes. */
long long mg_launch_count(void);
int mg_probe_begin(int family);
int mg_probe_end(double* out);

/* bf16 mode only: 1 (default) routes tensor-core-shaped contractions to the tcgen05 kernels, 0 keeps them on
 * the CUDA-core kernels (same bf16 operands; used for A/B parity tests).  Returns the previous setting.
 * The environment variable MELOGAN_DISABLE_TC=1 sets the initial value to 0. */
int mg_tc_enable(int on);

/* Packed-weight cache of the tensor-core kernels (bf16 mode), per context.  Off (default): every contraction re-packs its
 * weights into [tap][n][k] tiles right before the launch, so weights may be changed by anybody at any time.  On: one
 * persistent packed copy per (weight tensor, layout), re-packed only after mg_adam_step / mg_adam_step_clipped updated the
 * flat group the tensor lives in, after mg_gan_bind, or after mg_weight_cache_invalidate() -- the caller promises that
 * nothing else writes the bound parameters (melogan.trainer does; load_state_dict / rebind invalidate).  The frozen
 * emotion discriminator of the generator step (reference src/gan/train_gan.py:131-133) is then packed once, the
 * generator's weights once per cycle instead of once per launch.  Returns the previous setting. */
int mg_gan_weight_cache(mg_gan* ctx, int on);
int mg_weight_cache_invalidate(void);

/* ------------------------------------------------------------------------------------------
 * Stand-alone forms of the MLP-type inner blocks (fused inside mg_generator_* / mg_emotion_* on the hot path):
 *   NoiseToLatent.forward          reference src/gan/models.py:20-29
 *   GeneratorDecoder.pre           reference src/gan/models.py:46-51
 *   MLPClassifier.forward          reference src/emotion_discriminator/ed_model.py:74-101
 *   EmotionDiscriminator.forward with input_mode 'latent'   reference ed_model.py:128-136,156-160
 * are chains of these four float32 operators (nn.Linear, then activation [+ nn.Dropout keep-mask]):
 *   z  = x W^T + bias                         x (rows, K), W (N, K) row-major as nn.Linear.weight, z (rows, N)
 *   h  = act(z) * (mask ? mask * scale : 1)   act: 0 none, 1 ReLU, 2 LeakyReLU(0.2), 3 GELU (erf form); mask 0/1 floats
 *   dz = dh * act'(z) * (mask ? mask * scale : 1)
 *   dx = dz W (if dx), dW += dz^T x (if dW; caller zeroes), db += column sums of dz (if db)
 * ---------------------------------------------------------------------------------------- */
int mg_linear_forward(const float* x, const float* W, const float* bias, float* z, int rows, int K, int N, void* stream);
int mg_linear_backward(const float* x, const float* W, const float* dz, float* dx, float* dW, float* db, int rows, int K,
                       int N, void* stream);
int mg_act_dropout_forward(const float* z, const float* mask, float scale, int act, float* h, long long n, void* stream);
int mg_act_dropout_backward(const float* dh, const float* z, const float* mask, float scale, int act, float* dz,
                            long long n, void* stream);

/* ------------------------------------------------------------------------------------------
 * SyncBatchNorm for data parallelism (what nn.SyncBatchNorm.convert_sync_batchnorm would do to the reference's
 * GeneratorDecoder, src/gan/models.py:57,60): the batch statistics of the generator's two BatchNorm layers (forward: sum x,
 * sum x^2; backward: sum dy, sum dy*xhat) are summed over all ranks of ONE node through NVLink peer memory -- a one-CTA kernel
 * per sync point that writes the rank's vector into its own exchange buffer and reads the peers' (CUDA IPC), no NCCL call,
 * capturable in a CUDA graph.  Every rank adds in rank order: statistics are bit-identical on all ranks.
 *   1. every rank: mg_gan_sync_bn_export(ctx, rank, world, handle[64])   allocates the exchange buffer, returns its IPC handle
 *   2. exchange the handles between the processes (any transport; melogan.engine uses torch.distributed.all_gather)
 *   3. every rank: mg_gan_sync_bn_connect(ctx, handles[world * 64])       opens the peers' buffers and switches the mode on
 * world <= 8, one process per GPU.  Every rank must then run the same sequence of generator forwards / backwards.
 * ---------------------------------------------------------------------------------------- */
int mg_gan_sync_bn_export(mg_gan* ctx, int rank, int world, unsigned char* handle_out);
int mg_gan_sync_bn_connect(mg_gan* ctx, const unsigned char* handles);

/* ------------------------------------------------------------------------------------------
 * Gradient exchange of data parallelism over NVLink peer memory (the all-reduce torch DDP would add behind
 * loss.backward(), reference src/gan/train_gan.py:203,247): in-place SUM over the ranks of one node (<= 8, one process per
 * GPU) of a float32 vector, as three kernel launches on the caller's stream -- stage into the rank's CUDA-IPC exchange
 * region, publish an epoch flag (system-scope release), wait for the peers' flags and add their staged copies in RANK order
 * (bit-identical result on every rank).  No NCCL and no host synchronisation: graph-capturable.
 *   mg_peer_create(rank, world, max_floats, &peer, handle[64]) on every rank; exchange the handles;
 *   mg_peer_connect(peer, handles[world * 64]); then mg_peer_allreduce_sum(peer, data, n <= max_floats, stream), the same
 *   sequence of calls on every rank.
 * ---------------------------------------------------------------------------------------- */
typedef struct mg_peer mg_peer;
int mg_peer_create(int rank, int world, long long max_floats, mg_peer** out, unsigned char* handle_out);
int mg_peer_connect(mg_peer* peer, const unsigned char* handles);
int mg_peer_allreduce_sum(mg_peer* peer, float* data, long long n, void* stream);
void mg_peer_destroy(mg_peer* peer);

/* ------------------------------------------------------------------------------------------
 * Stand-alone forms of the conv-type inner blocks (fused inside their parents on the hot path):
 *   ConvBlock1D.forward, NotesEncoder.forward      reference src/emotion_discriminator/ed_model.py:24-69
 *   GeneratorDecoder.forward (deconv stack)        reference src/gan/models.py:52-83
 *   ConvEncoder.forward, ConvDecoder.forward       reference src/ae/model.py:9-48,64-98
 * are chains of ONE float32 operator, the "conv unit": Conv1d (stride 1; or k5 s2 p2) or ConvTranspose1d (k5 s2 p2 op1)
 * [+ BatchNorm1d in train or eval mode] + activation (0 none, 1 ReLU, 2 GELU, 3 Tanh), channels-last activations
 * x [R][Lin][Cin] -> y [R][Lout][Cout], weights in the reference's layouts (Conv1d [Cout][Cin][ks], ConvTranspose1d
 * [Cin][Cout][5]).  kind: 0 Conv1d, 1 ConvTranspose1d.  The forward also returns what the backward needs (z = pre-BatchNorm
 * output, mean / invstd, the GELU derivative tile); the backward accumulates parameter gradients (+=).
 * ---------------------------------------------------------------------------------------- */
typedef struct mg_convunit mg_convunit;
int mg_convunit_create(mg_convunit** out);
void mg_convunit_destroy(mg_convunit* unit);
int mg_convunit_forward(mg_convunit* unit, int kind, const float* x, const float* W, const float* bias, int R, int Lin, int Cin,
                        int Cout, int ks, int stride, int pad, const float* gamma, const float* beta, float* running_mean,
                        float* running_var, int train, int act, float* z, float* y, float* gd, float* mean, float* invstd,
                        void* stream);
int mg_convunit_backward(mg_convunit* unit, int kind, const float* x, const float* W, int R, int Lin, int Cin, int Cout, int ks,
                         int stride, int pad, const float* gamma, int train, int act, const float* z, const float* y,
                         const float* gd, const float* mean, const float* invstd, const float* dy, float* scratch0,
                         float* scratch1, float* dx, float* dW, float* dbias, float* dgamma, float* dbeta, void* stream);

/* Stream-ordered copy between any two device/pinned-host pointers (cudaMemcpyDefault). */
int mg_device_copy(void* dst, const void* src, long long nbytes, void* stream);

/* Test/diagnostic access to a named internal workspace buffer (see DESIGN.md for the names). */
int mg_gan_buffer(mg_gan* ctx, const char* name, void** ptr, long long* nbytes);

/* Counter-based RNG fill used by the throughput path (noise ~ N(0,1), alpha ~ U[0,1), keep-masks):
 * kind 0 = normal, 1 = uniform [0,1), 2 = Bernoulli(p) as 0/1 floats. */
int mg_rng_fill(float* out, long long n, int kind, float p, unsigned long long seed, unsigned long long offset,
                void* stream);
/* CUDA-graph-safe form: the Philox offset is (*counter_dev) * counter_mul, read on the device at run
 * time; mg_counter_add advances the counter inside the same stream/graph. */
int mg_rng_fill_counter(float* out, long long n, int kind, float p, unsigned long long seed,
                        const unsigned long long* counter_dev, unsigned long long counter_mul, void* stream);
int mg_counter_add(unsigned long long* counter_dev, unsigned long long inc, void* stream);

/* ------------------------------------------------------------------------------------------
 * Per-layer harness (tests/test_tc_layers_gpu.py, scripts/bench_layers.py): runs ONE contraction of the
 * training cycle -- exactly the call the step bodies make for that layer (same dispatch, same kernel
 * selection) -- on caller-supplied device tensors, so that every (layer shape x kernel variant) of the
 * benched configuration can be pinned against a float64 contraction of the same operands and timed alone.
 * The reference computes these with torch's conv1d / conv_transpose1d / linear and their autograd:
 * src/gan/models.py:20-29,46-83,140-169, src/emotion_discriminator/ed_model.py:35-69.
 *
 * op: 0 conv1d forward          in [R, Lin, Cin] -> out [R, Lin/stride, Cout]; W [Cout][Cin][ks] (or the strides given)
 *     1 stride-1 conv dgrad     in = dOut [R, L, Cout] -> out = dIn [R, L, Cin];  W [Cout][Cin][ks]
 *     2 k5 s2 up-sampling       in [R, Lin, Cin] -> out [R, 2 Lin, Cout] (ConvTranspose1d forward / strided-conv dgrad),
 *                               W element (t, n, k) at W[t + n * w_nstride + k * w_kstride]
 *     3 linear forward          in [R, Cin] -> out [R, Cout]; W [Cout][Cin]; n_perm_* permutes the output columns
 *     4 linear dgrad            in = dZ [R, Cout] -> out = dX [R, Cin]; W [Cout][Cin]
 *     5 conv1d wgrad            in = dOut [R, Lin/stride, Cout], in2 = x [R, Lin, Cin] -> dW [Cout][Cin][ks] +=
 *     6 ConvTranspose1d wgrad   in = x [R, Lin, Cin], in2 = dOut [R, 2 Lin, Cout] -> dW [Cin][Cout][5] +=
 *     7 linear wgrad            in = dZ [R, Cout], in2 = A [R, Cin] -> dW [Cout][Cin] +=
 * dtype flags: 0 = float32, 1 = bfloat16 (in / in2, out / aux, mul_src).  act: 0 none, 1 ReLU, 2 LeakyReLU(0.2),
 * 3 GELU (+ aux = GELU'), mul_mode: 0 none, 1 LeakyReLU' sign of mul_src, 2 ReLU' sign, 3 value.
 * tf32 = 1 lets float32 Linears take the kind::tf32 tensor-core path as bf16 mode does.
 * ---------------------------------------------------------------------------------------- */
typedef struct mg_debug_layer {
    int op, in_bf16, out_bf16, mask_bf16, tf32;
    int R, Lin, Cin, Cout, ks, stride, pad;
    int act, mul_mode, accumulate;
    int w_nstride, w_kstride;          /* -1 = the layer's natural layout */
    int n_perm_q, n_perm_p;
    const void* in; const void* in2; void* out; void* aux; const void* mul_src;
    const float* W; const float* bias; const float* col_scale; float* dW;
    /* op 0 only, optional: fused AdaptiveAvgPool1d(1) -- pool_out[r, n] = pool_scale * sum_l out[r, l, n]; *pool_done = 1 if
     * the kernel that ran produced it (weight-stationary tensor-core kernels), left untouched otherwise */
    float* pool_out; float pool_scale; int* pool_done;
    /* op 2 only, optional: fused column sums (bias gradient of the layer below) -- colsum_out[n] += sum of the stored
     * out[r, l, n] over the samples r < colsum_samples; *colsum_done = 1 if the kernels that ran produced it */
    float* colsum_out; int colsum_samples; int* colsum_done;
    /* op 2 only, optional (float32 out): fused BatchNorm statistics -- stats_out[n] += sum out[., ., n], stats_out[Cout + n] +=
     * sum out[., ., n]^2 (caller zeroes); *stats_done = 1 if the kernels that ran produced them */
    float* stats_out; int* stats_done;
} mg_debug_layer;
int mg_debug_layer_run(const mg_debug_layer* layer, void* stream);
/* Kernel-selection overrides of the harness: "force_bn" (64/128), "max_stages", "staging_bufs" (1/2), "no_ws",
 * "dbg" (ablation bits, -1 = off), "reverse" (-1 = alternate), "no_tma_store", "no_tma_mask", "no_reuse", "no_pair"
 * (never the CTA-pair cta_group::2 kernel), "no_rot", "no_fuse" (bits: 1 no pooling, 2 no column sums, 4 no BatchNorm statistics in
 * the epilogues), "fp32_tc" (1 = float32 contractions of the fp32 parity mode as six bf16 tensor-core terms -- every float32
 * operand is split exactly into three bf16 parts -- instead of CUDA-core FFMA; 0 = off; -1 = MELOGAN_FP32_TC, default off);
 * "reset" restores the product heuristics.  Returns MG_ERR_INVALID for an unknown key. */
int mg_debug_set(const char* key, int value);
/* One line describing the last tensor-core launch of this thread (kernel variant, grid, stages, ...); "" if the
 * last contraction did not run on the tensor cores. */
const char* mg_debug_last_launch(void);

#ifdef __cplusplus
}
#endif
#endif /* MELOGAN_B200_H */
