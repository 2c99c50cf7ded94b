/*
 * gsn_b200.h -- C ABI of libgsn_b200.so: the B200 (sm_100a) implementation of the GSN hot path of
 * Spiking-FullSubNet.  Plain device pointers, sizes and a CUDA stream; no torch types.
 *
 * Reference interfaces replaced (paths relative to the reference root):
 *   ESN = audiozen/models/spiking_fullsubnet/efficient_spiking_neuron.py
 *   MSF = audiozen/models/spiking_fullsubnet/modeling_spiking_fullsubnet.py
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to contiguous fp32 unless stated otherwise;
 *   - nothing is allocated inside: outputs and workspaces are caller-owned, sizes come from the
 *     *_workspace_bytes() queries;
 *   - all work is enqueued on `stream` (a cudaStream_t passed as void*); calls are asynchronous;
 *   - return value 0 = success, otherwise a GSN_E* code; gsn_last_error() gives the message of the
 *     last failure on the calling thread (mirrors the Python exceptions of the reference);
 *   - "rows" R are the independent recurrences of one sequence model: R = batch (full-band model) or
 *     batch * num_subbands (sub-band model), row index r = b * N + n  (MSF:155).
 */
#ifndef GSN_B200_H_
#define GSN_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define GSN_ABI_VERSION 1

#define GSN_OK 0
#define GSN_EINVAL 1   /* bad argument (ValueError / AssertionError in the reference) */
#define GSN_ECUDA 2    /* CUDA runtime error (RuntimeError) */
#define GSN_ENOSUP 3   /* shape not supported by the requested backend (NotImplementedError) */

/* recurrence back ends */
#define GSN_BACKEND_AUTO 0
#define GSN_BACKEND_SIMT 1     /* fp32 CUDA-core kernel: any H <= 512, the on-device fp32 arbiter */
#define GSN_BACKEND_TCGEN05 2  /* tcgen05/TMEM kernel, recurrent weights split in exact bf16 planes (H <= 320) */
#define GSN_BACKEND_TCGEN05_I8 3  /* tcgen05 kind::i8: weights as 32-bit (H <= 448) / 24-bit row-scaled fixed point
                                     in byte planes, exact int32 accumulation (H <= 512) */

typedef void* gsn_stream_t;

#if defined(__GNUC__)
#define GSN_API __attribute__((visibility("default")))
#else
#define GSN_API
#endif

GSN_API int gsn_abi_version(void);
GSN_API const char* gsn_last_error(void);
/* cudaSetDevice for the library's runtime instance (call when the caller switches device). */
GSN_API int gsn_bind_device(int device);
/* sm count, compute capability and max opt-in shared memory of the bound device. */
GSN_API int gsn_device_info(int* sm_count, int* cc_major, int* cc_minor, int* smem_optin_bytes);

/* ---- front end ------------------------------------------------------------------------------
 * MSF:434-436 + the 'b f t -> t b f' of MSF:108:  cm[t, b, f] = mag[b, f, t] ** fdrc  for f < f_keep.
 * mag [B, F, T] (STFT magnitude), cm [T, B, f_keep].  fdrc == 0.5 is evaluated as sqrt (as torch
 * does), fdrc == 1 as a copy.                                                                    */
GSN_API int gsn_compress_mag(const float* mag, float* cm, int B, int F, int f_keep, int T, float fdrc,
                     gsn_stream_t stream);

/* MSF:241-258 (+ _freq_unfold MSF:265-312, concat with the tiled full-band output MSF:443) and the
 * pre-LayerNorm MSF:111-112, as ONE gather:
 *   x[t, b*N + n, j] = LN_j( j <  ctr+2*nbr : cm[t, b, reflect(lo + n*ctr - nbr + j)]
 *                            j >= ctr+2*nbr : fb[t, b, (lo + n*ctr + j - (ctr+2*nbr)) mod f_fb] )
 * reflect(q) = -q for q < 0, 2*(f_cm-1) - q for q > f_cm-1.   K = ctr + 2*nbr + (fb ? ctr : 0).
 * fb == NULL drops the second part (full-band model input: N = 1, lo = 0, ctr = K, nbr = 0).
 * ln_weight == NULL skips the LayerNorm (use_pre_layer_norm = False).  K <= 1024.               */
GSN_API int gsn_subband_features(const float* cm, int f_cm, const float* fb, int f_fb, float* x, int T, int B,
                         int N, int lo, int ctr, int nbr, const float* ln_weight, const float* ln_bias,
                         float ln_eps, gsn_stream_t stream);
/* rowsum[t, b*N + n] = sum over j of the UN-normalised gathered features above (no x written): what surface B's laplace
 * norms reduce (model_low_freq.py:146-171: mean over an utterance; model_low_freq_count_time.py:173-204: running mean of a
 * row) before gsn_xplanes_stream gathers the same features again and divides by the result.  rowsum [T, B*N].        */
GSN_API int gsn_subband_rowsums(const float* cm, int f_cm, const float* fb, int f_fb, float* rowsum, int T, int B, int N,
                                int lo, int ctr, int nbr, gsn_stream_t stream);

/* ---- dense fp32 linear (input-to-hidden product ESN:141, proj MSF:118) -------------------------
 * out[M, N] = a[M, K] @ w[N, K]^T + bias[N]   (bias may be NULL).  fp32 FMA, k ascending.
 * act: 0 none, 1 tanh, 2 sigmoid, 3 relu applied to a SECOND output out_act (may be NULL) so that
 * both the pre-activation trace entry (MSF:119) and the activated output (MSF:122) are produced.  */
GSN_API int gsn_linear_f32(const float* a, const float* w, const float* bias, float* out, float* out_act,
                   int act, int64_t M, int K, int N, gsn_stream_t stream);

/* Same product for SPIKE inputs (layers >= 1: the previous layer's h; proj: the last layer's h) on tcgen05:
 * `a` must hold values exactly representable in bf16 ({0,1} spikes); the fp32 weights are split into three
 * exact bf16 planes kept in tensor memory, so the result is an fp32 sum of exact products.  K <= 320, K % 4 == 0,
 * a 16-byte aligned.  sm_budget as for gsn_layer_recurrence (0 = whole device).                            */
GSN_API int gsn_linear_spikes(const float* a, const float* w, const float* bias, float* out, float* out_act,
                              int act, int64_t M, int K, int N, int sm_budget, gsn_stream_t stream);

/* gsn_linear_spikes with the spike trace bit-packed: a_bits [M, ceil(K/32)] uint32 as written by
 * gsn_layer_recurrence_bits / gsn_pack_spikes.  K <= 320; no alignment requirement.                       */
GSN_API int gsn_linear_spike_bits(const uint32_t* a_bits, const float* w, const float* bias, float* out,
                                  float* out_act, int act, int64_t M, int K, int N, int sm_budget,
                                  gsn_stream_t stream);

/* ---- the recurrence: GSULayer.forward ESN:75-81 over GSUCell.forward ESN:132-153 ----------------
 * For t = 0..T-1, rows r, neurons j:
 *   z      = xproj[t, r, :] + h_{t-1}[r, :] @ w_hh^T            (xproj = x @ w_ih^T, no bias)
 *   f      = sigmoid(z_f + bias[j]);  g = z_g + bias[H + j]     (shared: z_f = z_g = z[j]; else
 *                                                                z_f = z[j], z_g = z[H + j])
 *   c_t    = (f * c_{t-1} + (1 - f) * g) * bn_scale[j] + bn_shift[j]   (eval BatchNorm folded the way
 *                                                                torch's CPU kernel folds it; NULL = no BN)
 *   h_t    = c_t >= 0
 * xproj [T, R, gH], w_hh [gH, H] (g = 1 shared, 2 unshared), bias [2H].
 * h0 / c0 [R, H] may be NULL (zeros, MSF:100-106).  h_out [T, R, H] fp32 {0,1} (the trace entry of
 * all_layer_outputs, ESN:60); c_out [T, R, H] optional (NULL) membrane trace; hT / cT [R, H] optional.
 * workspace: gsn_layer_recurrence_workspace_bytes(R, H, shared, backend) bytes, 256-byte aligned.
 * sm_budget: number of SMs this launch should plan for (0 = the whole device).  Callers that run several
 * recurrences concurrently (sub-band models, wavefront schedule) pass each one's share, so that the row tile
 * NT is chosen large enough for all of them to be co-resident.                                          */
GSN_API size_t gsn_layer_recurrence_workspace_bytes(int R, int H, int shared, int backend);
GSN_API int gsn_layer_recurrence(const float* xproj, const float* w_hh, const float* bias,
                         const float* bn_scale, const float* bn_shift, const float* h0,
                         const float* c0, float* h_out, float* c_out, float* hT, float* cT, int T,
                         int R, int H, int shared, int backend, int sm_budget, void* workspace,
                         gsn_stream_t stream);
/* Same call with one more output: h_bits [T, R, W] uint32, W = ceil(H/32), the SAME spike trace bit-packed
 * (neuron n of row r at frame t = bit n%32 of h_bits[t, r, n/32]; bits of neurons >= H are 0).  The tcgen05 kernel
 * writes the ballot words it exchanges between CTAs anyway; other back ends pack h_out afterwards
 * (gsn_pack_spikes).  It is what gsn_linear_spike_bits reads: 1 bit instead of 4 bytes per spike.  May be NULL. */
GSN_API int gsn_layer_recurrence_bits(const float* xproj, const float* w_hh, const float* bias,
                         const float* bn_scale, const float* bn_shift, const float* h0,
                         const float* c0, float* h_out, float* c_out, float* hT, float* cT, uint32_t* h_bits,
                         int T, int R, int H, int shared, int backend, int sm_budget, void* workspace,
                         gsn_stream_t stream);
/* ---- spectral front / back end on the COMPLEX STFT (gsn_spectral.cu; interleaved re/im as torch.stft returns it) ----
 * gsn_compress_spec   : cm[t, b, f] = |spec[b, f, t]|^fdrc for f < f_keep (MSF:434-436 + "b f t -> t b f", MSF:108);
 *                       spec_ri [B, F, T, 2].  Same result as torch.abs + gsn_compress_mag without the |stft| tensor.
 * gsn_deepfilter_spec : gsn_deepfilter_band with complex in / complex out: out_ri [B, S, F_out, T, 2]; layout 0 = proj
 *                       features ordered (c fc df s) (MSF:160-167), 1 = (c df s fc) (cirm_gsn, CGN:230).
 * gsn_spec_passthrough: out[b, s, f, t] = spec[b, f, t] for f in [f_lo, F): the bins no band filters (MSF:461-468).
 *                       mag_out (both; may be NULL): also |out| as fp32 in the same index order (enh_mag, MSF:472).
 * time_major != 0     : the spectra are [B, T, F] / [B, S, T, F_out] -- the layout cuFFT reads and writes (torch.stft
 *                       returns a transposed VIEW of it), so the whole forward() runs without a transpose copy.
 * gsn_overlap_add     : synthesis half of torch.istft(center=True) (audio_feature.py:297-347): frames [B, T, n_fft] =
 *                       inverse real FFT of every frame (unwindowed), window [n_fft]; y[b, s] = sum over the frames
 *                       covering sample s + n_fft/2 of frame * window, divided by the overlap-added squared window;
 *                       samples past the last frame are zero.  y [B, length].                                    */
GSN_API int gsn_compress_spec(const float* spec_ri, float* cm, int B, int F, int f_keep, int T, float fdrc,
                              int time_major, gsn_stream_t stream);
GSN_API int gsn_deepfilter_spec(const float* proj, const float* spec_ri, float* out_ri, float* mag_out, int T, int B,
                                int N, int ctr, int df, int S, int lo, int F, int F_out, int layout, int time_major,
                                gsn_stream_t stream);
GSN_API int gsn_spec_passthrough(const float* spec_ri, float* out_ri, float* mag_out, int T, int B, int S, int f_lo,
                                 int F, int F_out, int time_major, gsn_stream_t stream);
/* gsn_frame_signal    : analysis half of torch.stft(center=True, pad_mode="constant") in front of the real FFT
 *                       (audio_feature.py:236-294): y [B, L] -> frames [B, T, n_fft], T = 1 + L / hop, zero padding of
 *                       n_fft/2 on both sides, framing, analysis window -- one pass instead of pad + unfold + multiply. */
GSN_API int gsn_frame_signal(const float* y, const float* window, float* frames, int B, int L, int T, int n_fft, int hop,
                             gsn_stream_t stream);
GSN_API int gsn_overlap_add(const float* frames, const float* window, float* y, int B, int T, int n_fft, int hop,
                            int length, gsn_stream_t stream);
/* ---- the recipes' 512-point real FFTs fused with their neighbours (gsn_fft.cu; n_fft = win_length = 512 only) ----
 * gsn_stft_compress   : torch.stft(center=True, pad_mode="constant", hann window; audio_feature.py:236-294) AND
 *                       |X|^fdrc (MSF:434-436, "b f t -> t b f" MSF:108) in one pass over the waveform: y [B, L] ->
 *                       spec_ri [B, T, 257, 2] (time-major complex spectrum, the layout of `time_major` above) and
 *                       cm [T, B, f_keep] (may be NULL).  Replaces gsn_frame_signal + cuFFT R2C + gsn_compress_spec.
 * gsn_irfft_frames    : spec_ri [B, T, 257, 2] -> frames [B, T, 512]: inverse real FFT of every frame with the 1/n
 *                       normalisation of torch.fft.irfft, unwindowed -- what gsn_overlap_add reads
 *                       (audio_feature.py:297-347).  Imaginary parts of the DC and Nyquist bins are ignored, as cuFFT does.
 * gsn_deepfilter_irfft: the same inverse transform of the DEEP-FILTERED spectrum (MSF:315-346, 449-472; one speaker),
 *                       computed on the fly: projs[i] [T, B*N[i], 2*ctr[i]*df[i]] are the bands' proj outputs in
 *                       frequency order from bin 0 (feature order `layout` as gsn_deepfilter_spec), bins above the
 *                       last band pass through (MSF:461-468).  mag_out [B, T, 257] = |enhanced| (enh_mag, MSF:472;
 *                       may be NULL), enh_ri [B, T, 257, 2] = the enhanced spectrum itself (may be NULL: it then never
 *                       exists in memory).  projs / N / ctr / df are HOST arrays of n_bands <= 4 entries.            */
GSN_API int gsn_stft_compress(const float* y, const float* window, float* spec_ri, float* cm, int B, int L, int T,
                              int n_fft, int hop, int f_keep, float fdrc, gsn_stream_t stream);
GSN_API int gsn_irfft_frames(const float* spec_ri, float* frames, int B, int T, int n_fft, gsn_stream_t stream);
GSN_API int gsn_deepfilter_irfft(const float* const* projs, const int* N, const int* ctr, const int* df, int n_bands,
                                 int layout, const float* spec_ri, float* frames, float* mag_out, float* enh_ri, int B,
                                 int T, int n_fft, gsn_stream_t stream);

/* ---- streaming recurrence (gsn_recurrence_stream.cu): StackedGSU.forward ESN:50-62 as a frame-granular pipeline ---
 * One persistent, warp-specialised tcgen05 launch runs GSULayer.forward (ESN:75-81) of one layer for all T frames
 * from a ZERO initial state (MSF:100-106), shared gate weights only, and is chained to concurrently running producer /
 * consumer kernels through per-frame counters in global memory:
 *   - input, one of
 *       xproj [T,R,H]               input projection without bias (layer 0 / wide layers), staged by bulk copies, or
 *       in_bits [T,R,ceil(K_in/32)] bit-packed spikes of the layer below + w_ih [H,K_in]: the input-to-hidden product
 *                                   runs inside the kernel (both weight matrices resident in tensor memory; needs
 *                                   3*ceil16(H)/2 + 3*ceil16(K_in)/2 + 2*NT <= 512 columns, e.g. H = K_in = 160), or
 *       in_planes                   REAL-valued layer-0 input as the bf16x3 operand images gsn_xplanes_stream writes
 *                                   (row tile = gsn_recurrence_stream_tile(.., fused = 1, ..)) + w_ih [H,K_in]: the
 *                                   product x_t . w_ih^T (ESN:141) runs inside the kernel too (8 of the 9 plane pairs,
 *                                   fp32-faithful), one bulk copy per frame; same tensor-memory condition, K_in <= 256;
 *                                   planes_ring = frames the image buffer holds (frame t in slot t % planes_ring;
 *                                   <= 0 or >= T: one slot per frame; a shorter ring needs out_cnt, which the
 *                                   producer reads as back-pressure), or
 *       in_image                    the spikes of the layer below as the bf16 operand image that layer wrote through
 *                                   its img_out (same rows, 16-row tiles, K_in = its H; ring = planes_ring) + w_ih:
 *                                   like in_bits, but the input arrives with one bulk copy per frame instead of
 *                                   being expanded from bits by the loader warp (0.15 us per frame at H = 160);
 *   - img_out (may be NULL; 16-row tiles only): this layer's spikes of frame t as the operand image of the layer
 *     above, [img_ring][ceil(R/16)][16 x ceil16(H) bf16] = gsn_spike_image_bytes(img_ring, R, H) bytes, slot
 *     t % img_ring; with a ring shorter than T the input of frame t waits for bp_cnt[t - img_ring] >= bp_target
 *     (bp_cnt = the out_cnt of the consuming launch, bp_target = its CTA count), so the ring stays L2-resident;
 *   - in_cnt [T] (may be NULL): frame t of the input may be read once in_cnt[t] >= in_target (acquire);
 *   - h_bits [T,R,ceil(H/32)]: the spike trace, bit-packed (always); h_out / c_out [T,R,H] fp32 optional (NULL);
 *     hT / cT [R,H] optional;
 *   - out_cnt [T] (may be NULL): every CTA adds 1 to out_cnt[t] (release) when its part of frame t (h_bits, h_out,
 *     c_out) is globally visible; the frame is complete at gsn_recurrence_stream_ctas(...);
 *   - spike_count (may be NULL): += number of spikes emitted (firing-rate numerator of SynOps, metric.py:303-340).
 * Results are bit-identical to gsn_layer_recurrence(TCGEN05) fed by gsn_linear_spike_bits / the same xproj.
 * Counters must be zeroed by the caller before the first producer starts.  workspace may be NULL.               */
GSN_API int gsn_recurrence_stream(const float* xproj, const uint32_t* in_bits, const void* in_planes,
                                  int planes_ring, const float* w_ih, int K_in, const float* w_hh, const float* bias,
                                  const float* bn_scale, const float* bn_shift, uint32_t* h_bits, float* h_out,
                                  float* c_out, float* hT, float* cT, const unsigned int* in_cnt,
                                  unsigned int in_target, unsigned int* out_cnt, unsigned long long* spike_count,
                                  const void* in_image, void* img_out, int img_ring, const unsigned int* bp_cnt,
                                  unsigned int bp_target, int T, int R, int H, int sm_budget, void* workspace,
                                  gsn_stream_t stream);
GSN_API size_t gsn_spike_image_bytes(int frames, int R, int H);
/* Loads every kernel of the streaming pipeline into the context.  The pipeline's kernels spin on counters their
 * producers advance, and with lazy module loading the first launch of a kernel synchronises with running kernels:
 * call once per process and device BEFORE the first pipeline launch (a spinning consumer would otherwise wait for a
 * producer that cannot be loaded, and trap on its wall-clock bound). */
GSN_API int gsn_stream_preload(void);
/* Row tile (16 / 32 / 64; 0 = unsupported) and number of CTAs (= out_cnt target) of that launch. */
GSN_API int gsn_recurrence_stream_tile(int R, int H, int K_in, int fused, int sm_budget);
GSN_API int gsn_recurrence_stream_ctas(int R, int H, int K_in, int fused, int sm_budget);

/* gsn_linear_spike_bits as a stage of the streaming pipeline: a_bits [T*R, ceil(K/32)] is produced frame by frame by
 * a concurrently running gsn_recurrence_stream; the persistent kernel waits for in_cnt[t] >= in_target before it reads
 * frame t and adds the rows it has written to out_cnt[t] per 128-feature slice (frame complete at R * ceil(N/128)).
 * `ctas` persistent CTAs in total (>= ceil(N/128)).  Either counter array may be NULL.  Same results.          */
GSN_API int gsn_linear_spike_bits_stream(const uint32_t* a_bits, const float* w, const float* bias, float* out,
                                         float* out_act, int act, int T, int R, int K, int N, int ctas,
                                         const unsigned int* in_cnt, unsigned int in_target, unsigned int* out_cnt,
                                         gsn_stream_t stream);

/* Streaming front end for the fused layer-0 recurrence (gsn_xplanes_stream.cu): the gather of gsn_subband_features +
 * LayerNorm + truncation split x = hi + mid + lo into three bf16 planes, written as the tcgen05 B-operand images
 * xop[t][tile][plane lo,mid,hi][nt rows x ceil16(K), K-major 8x16-byte core matrices] that gsn_recurrence_stream
 * (in_planes) fetches with one bulk copy per frame.  CUDA cores only; `ctas` persistent CTAs.  xop must hold
 * gsn_xplanes_bytes(T, R, K, nt) bytes, 128-byte aligned, and be ZEROED once by the caller (padding rows of the last
 * tile).  in_cnt / in_target as below; out_cnt[t] += rows written (frame complete at R = B*N).  K <= 256.       */
GSN_API size_t gsn_xplanes_bytes(int T, int R, int K, int nt);
/* ring: frames the buffer holds (gsn_xplanes_bytes(ring, ...) bytes; <= 0 or >= T: T).  With a ring shorter than T the
 * images stay L2-resident instead of making a DRAM round trip; frame t then waits for bp_cnt[t - ring] >= bp_target
 * (bp_cnt = the out_cnt of the consuming gsn_recurrence_stream, bp_target = its CTA count).
 * row_div (may be NULL): every feature of a row is DIVIDED by row_div[b] (div_mode 1, b = utterance: surface B's
 * offline_laplace_norm, model_low_freq.py:146-171, with mu + eps precomputed) or by row_div[t * R + r] (div_mode 2: the
 * cumulative norm of model_low_freq_count_time.py:173-204) after the optional LayerNorm. */
GSN_API int gsn_xplanes_stream(const float* cm, int f_cm, const float* fb, int f_fb, const float* ln_weight,
                               const float* ln_bias, float ln_eps, const float* row_div, int div_mode, float* x_out,
                               void* xop, int ring,
                               const unsigned int* in_cnt, unsigned int in_target, unsigned int* out_cnt,
                               const unsigned int* bp_cnt, unsigned int bp_target, int T, int B, int N, int lo,
                               int ctr, int nbr, int nt, int ctas, gsn_stream_t stream);

/* Streaming front end of one sequence model as a separate tensor-core stage (gsn_stage_stream.cu; used when the fused
 * layer-0 recurrence does not fit tensor memory): the same gather + LayerNorm (x within 2e-6 of gsn_subband_features)
 * fused with the layer-0 input-to-hidden product xproj[t, r, :] = x[t, r, :] @ w_ih^T
 * (ESN:141; no bias) on tcgen05: both operands as three exact bf16 planes, 8 of the 9 plane pairs accumulated in fp32
 * (fp32-faithful; not bit-identical to gsn_linear_f32).  Persistent; walks the [T*R] rows in order.
 *   in_cnt [T] (may be NULL): `fb` of frame t may be read once in_cnt[t] >= in_target;
 *   out_cnt [T] (may be NULL): += rows written per (row tile, 128-feature slice); frame t of xproj is complete at
 *   R * ceil(H/128).  x_out [T,R,K] (may be NULL) receives the normalised input (all_layer_outputs[0]).
 * ctas_per_slice: persistent CTAs per 128-feature slice.  K <= 256.                                             */
GSN_API int gsn_pre_stream_supported(int K, int H);
GSN_API int gsn_pre_stream(const float* cm, int f_cm, const float* fb, int f_fb, const float* ln_weight,
                           const float* ln_bias, float ln_eps, const float* w_ih, float* x_out, float* xproj,
                           const unsigned int* in_cnt, unsigned int in_target, unsigned int* out_cnt, int T, int B,
                           int N, int lo, int ctr, int nbr, int H, int ctas_per_slice, gsn_stream_t stream);

/* bits[r, w] (W = ceil(H/32) words per row) from an fp32 {0,1} trace h [rows, H]. */
GSN_API int gsn_pack_spikes(const float* h, uint32_t* bits, int64_t rows, int H, gsn_stream_t stream);
/* Per-thread launch options for the following calls.  GSN_OPT_PDL != 0: gsn_layer_recurrence(_bits) launches of the
 * tcgen05 back end are enqueued as PROGRAMMATIC DEPENDENTS of the previous kernel in their stream (CUDA programmatic
 * dependent launch): the grid may be scheduled while that kernel drains and waits (griddepcontrol.wait) before it
 * reads anything.  Meant for back-to-back frame chunks of one layer on one stream (the wavefront schedule), where the
 * previous kernel is the previous chunk; the caller must make sure that kernel does not WRITE this call's weights.  */
#define GSN_OPT_PDL 1
/* GSN_OPT_F32_MAX_CTAS = n > 0: gsn_subband_features and gsn_linear_f32 launch at most n CTAs and stride over their
 * work (same results).  The wavefront schedule uses it so that these short kernels cannot occupy the SMs the
 * latency-critical recurrence chunks are about to be placed on.  0 = one CTA per tile (default).               */
#define GSN_OPT_F32_MAX_CTAS 2
GSN_API int gsn_set_option(int option, int value);
/* Row tile NT (rows per thread-block cluster: 16, 32 or 64) the tcgen05 back ends use for this shape and SM budget
 * (0 for the SIMT back end / unsupported shapes): what gsn_layer_recurrence(_bits) will launch.  Parity tests use it
 * to make sure every tile instantiation is exercised.                                                        */
GSN_API int gsn_layer_recurrence_tile(int R, int H, int shared, int backend, int sm_budget);
/* which backend GSN_BACKEND_AUTO resolves to for this shape (GSN_BACKEND_SIMT / _TCGEN05 / _TCGEN05_I8). */
GSN_API int gsn_layer_recurrence_pick_backend(int R, int H, int shared);

/* ---- training path of the recurrence (fp32 CUDA cores, cooperative launch) -------------------------------
 * Forward: GSUCell.forward ESN:132-153 over all frames with nn.BatchNorm1d in TRAINING mode when training != 0
 * (batch statistics over the R rows per frame, biased variance; running_mean / running_var updated in place
 * every frame with `momentum` and the unbiased variance, ESN:149-150), eval-mode statistics otherwise;
 * bn_weight == NULL means no BatchNorm.  Besides h_out / c_out [T,R,H] it saves what BPTT needs:
 * f_out = sigmoid(forget gate), g_out = cell-gate pre-activation, xhat_out = normalised pre-BN membrane
 * [T,R,H] and invstd_out [T,H] (the last two only with batch statistics).  Zero initial state (MSF:100-106).
 * Backward: BPTT with the Triangle surrogate max(0, 1-|c|) (ESN:95-101) and the batch-statistics BatchNorm
 * backward.  dh_out [T,R,H] = dL/dh_t; dz [T,R,gH] = dL/d(gate pre-activations) = dL/dxproj,
 * from which dW_hh = dz^T h_{t-1}, dW_ih = dz^T x, dx = dz W_ih are plain GEMMs; dbias_part [ceil(R/8), 2H]
 * per-CTA partial sums of dL/dbias; dgamma / dbeta [H] BatchNorm affine gradients (batch statistics only).
 * workspace: gsn_layer_train_workspace_bytes(R, H, shared) bytes, 256-byte aligned.  H <= 512.            */
GSN_API size_t gsn_layer_train_workspace_bytes(int R, int H, int shared);
GSN_API int gsn_layer_train_forward(const float* xproj, const float* w_hh, const float* bias,
                                    const float* bn_weight, const float* bn_bias, float* running_mean,
                                    float* running_var, float* h_out, float* c_out, float* f_out, float* g_out,
                                    float* xhat_out, float* invstd_out, int T, int R, int H, int shared,
                                    int training, float momentum, float eps, void* workspace,
                                    gsn_stream_t stream);
/* The same training forward on tcgen05 (gsn_recurrence_tc.cu with TRAIN = true: weights stationary in TMEM, the
 * per-frame BatchNorm statistics reduced over row groups, row tiles and a grid barrier; cooperative cluster
 * launch).  Shared gate weights, H <= 320, H % 4 == 0; gsn_layer_train_tc_supported() tells.  Same outputs.  */
GSN_API int gsn_layer_train_tc_supported(int R, int H, int shared);
GSN_API size_t gsn_layer_train_tc_workspace_bytes(int R, int H);
GSN_API int gsn_layer_train_forward_tc(const float* xproj, const float* w_hh, const float* bias,
                                       const float* bn_weight, const float* bn_bias, float* running_mean,
                                       float* running_var, float* h_out, float* c_out, float* f_out, float* g_out,
                                       float* xhat_out, float* invstd_out, int T, int R, int H, int training,
                                       float momentum, float eps, int sm_budget, void* workspace,
                                       gsn_stream_t stream);
GSN_API int gsn_layer_train_backward(const float* dh_out, const float* w_hh, const float* c, const float* f,
                                     const float* g, const float* xhat, const float* invstd,
                                     const float* bn_weight, const float* running_var, float* dz,
                                     float* dbias_part, float* dgamma, float* dbeta, int T, int R, int H,
                                     int shared, int training, float eps, void* workspace, gsn_stream_t stream);

/* ---- back end: deep filter (MSF:315-346) -- "next" row f1 ----------------------
 * Applies the coefficients straight from the sub-band proj output, skipping the 6-D rearrangement:
 *   proj [T, B*N, P] with p = ((c2*ctr + fc)*df + d)*S + s   (MSF:160-167)
 *   spec_re/spec_im [B, F, T] noisy STFT;  out_re/out_im [B, S, F_out, T]
 *   out[b, s, lo + n*ctr + fc, t] = sum_d spec[b, lo + n*ctr + fc, t - (df-1) + d] * coef[d]      */
GSN_API int gsn_deepfilter_band(const float* proj, const float* spec_re, const float* spec_im, float* out_re,
                        float* out_im, int T, int B, int N, int ctr, int df, int S, int lo, int F,
                        int F_out, gsn_stream_t stream);

/* ---- optional launch trace (development / profiling aid) ------------------------------------------------
 * device_buffer (>= 64 + 32*n bytes of device memory) receives one 32-byte record per traced kernel launch:
 * {u64 start_ns, u64 end_ns (%globaltimer of CTA 0), i32 kind (1 linear, 2 recurrence, 3 features), a, b, c}.
 * Header: u32 count, u32 capacity.  NULL disables.  The pointer is baked into launches when they are enqueued
 * (so enable it before capturing a CUDA graph).                                                          */
GSN_API int gsn_trace_set(void* device_buffer, size_t bytes);

#ifdef __cplusplus
}
#endif
#endif /* GSN_B200_H_ */
