/*
 * flic_b200.h -- C ABI of libflic_b200.so, the B200 (sm_100a) implementation of the entropy
 * coding hot path of lym01803/FinalProject-LosslessImageCompression.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / Python types.  Every entry
 * point names the reference interface it replaces (paths relative to the reference repo).
 * INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Conventions
 *   - Every function returns 0 on success.  A positive value is a cudaError_t, a negative value
 *     one of FLIC_E_*.  flic_last_error() returns a thread-local description.
 *   - Symbols are float32 multiples of 1/256 ("grid floats"), exactly what the reference passes
 *     to rans.encode (trainer.py:311).  mean is float32, scale is the LINEAR scale
 *     exp(logscale) (trainer.py:313), float32.
 *   - A "stream" is one independent rANS bitstream: a contiguous slice
 *     [stream_offsets[s], stream_offsets[s+1]) of the symbol arrays, coded from state 1<<32
 *     (trainer.py:310) unless init_states says otherwise.  Its result is exactly the
 *     (state, buffer) pair the reference's encode() returns for that slice: final_states[s] and
 *     packed[word_offsets[s] .. word_offsets[s+1]) in emission order.
 *   - status[s] is a bit set of FLIC_ST_*; 0 means the stream is valid.  The reference has no
 *     such channel: it raises ZeroDivisionError for scale==0 (rans/rans.cpp:1435) and silently
 *     corrupts out-of-window symbols; both are reported here instead.
 *   - "device" entry points take device pointers and a cudaStream_t (as void*; NULL = default
 *     stream) and never synchronise.  "host" entry points take host pointers, do the copies and
 *     synchronise before returning.
 */
#ifndef FLIC_B200_H
#define FLIC_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FLIC_ABI_VERSION 2

/* negative return codes */
#define FLIC_E_ARG (-1)      /* bad argument (null pointer, negative size, offsets not monotone) */
#define FLIC_E_CAPACITY (-2) /* an output buffer or the codec workspace is too small */
#define FLIC_E_NOMEM (-3)    /* host allocation failed */
#define FLIC_E_STATUS (-4)   /* single-stream drop-ins only: the stream's status word is non-zero */

/* per-stream status bits */
#define FLIC_ST_ZERO_SCALE 1    /* scale == 0: reference raises ZeroDivisionError (rans/rans.cpp:1435-1437) */
#define FLIC_ST_OUT_OF_WINDOW 2 /* symbol not in [lower, lower+2047] or not on the 1/256 grid */
#define FLIC_ST_UNDERRUN 4      /* decoder ran out of words (reference: unchecked read, rans/rans.cpp:2109) */
#define FLIC_ST_NONFINITE 8     /* NaN/inf scale or |mean| > 16384 */
#define FLIC_ST_BAD_END_STATE 16 /* decoder did not end at 1<<32 with all words consumed (rans/test.py:26) */
#define FLIC_ST_NO_SYMBOL 32    /* decoder: no symbol in the window matches (corrupt input) */
#define FLIC_ST_TOO_LONG 64     /* decoder: one stream holds 2^32 words or more (not supported) */

typedef void* flic_cuda_stream_t; /* a cudaStream_t */

int flic_abi_version(void);
const char* flic_last_error(void);
/* Number of kernels launched by this library in the calling process since load (bench evidence). */
int64_t flic_kernel_launches(void);
/* Name of the kernel the most recent flic_rans_encode (which = 0) / flic_rans_decode (which = 1)
 * launched: the coder has a lane-per-stream and a warp-cooperative variant of each (profiling aid). */
const char* flic_last_coder_kernel(int which);
/* Which decode kernel flic_rans_decode launches: -1 = by stream count (default: a thread-block
 * cluster of 8 / 4 / 2 CTAs per stream while all streams' clusters are resident at once -- the
 * reference's own partitions of one to a few streams, trainer.py:308-318, rans/test.py:6-22 --,
 * the CTA-per-stream kernel up to two streams per SM, the lane-per-stream kernel above),
 * 0 = always lane-per-stream, 1 = always CTA-per-stream, 2 / 4 / 8 = always a cluster of that many
 * CTAs per stream.  Process-wide; for measurements and for tests that compare them (their results
 * are bit-identical by construction).  Returns the previous setting. */
int flic_set_decode_kernel(int which);
/* Streams the cluster-per-stream decode kernel can hold at once on the current device with clusters of
 * `cluster` (2, 4 or 8) CTAs; 0 if the device cannot launch them.  The default choice uses the largest
 * cluster size whose capacity covers the call's stream count. */
int64_t flic_decode_cluster_capacity(int cluster);

/* ------------------------------------------------------------------------------------------
 * Device entry points
 * ------------------------------------------------------------------------------------------ */

/* K1.  Replaces encode pass 1, rans/rans.pyx:49-56 (+ CDF(), rans.pyx:31-35): per symbol
 * start = CDF(x - 1/256), freq = CDF(x) - start.  status_word (one int32, device) is OR-ed. */
int flic_cdf_tables(const float* x, const float* mean, const float* scale, int64_t n_symbols,
                    uint32_t* start, uint32_t* freq, int32_t* status_word,
                    flic_cuda_stream_t stream);

/* Diagnostic.  y[i] = the device restatement of glibc expf (the `exp` of rans/rans.pyx:26 as the
 * reference links it, expf@GLIBC_2.27, rans/rans.cpp:1301).  Lets a test sweep it against libm. */
int flic_debug_expf(const float* x, float* y, int64_t n, flic_cuda_stream_t stream);

/* Diagnostic: y[i] = part1 of CDF() for the float argument arg[i] alone, i.e.
 * (int) roundf((float)(logistic(arg) * 16775168)) of rans/rans.pyx:25-26,34 -- every operation
 * of CDF() after the division.  A function of 32 bits, so tests sweep all of them. */
int flic_debug_part1(const float* arg, int32_t* y, int64_t n, flic_cuda_stream_t stream);

/* Diagnostics: the two divisions the coder does with reciprocals, checked on the device against the
 * hardware's exact division over n generated operand pairs; *mismatches (uint64, device, caller
 * zeroes it) receives the number of disagreements.
 *   div_check:  (xq + 1/512 - mean) / scale of CDF(), rans/rans.pyx:34 (rans.cpp:1434-1439);
 *               mode 0 = realistic magnitudes, 1 = any finite positive scale / any |mean| <= 16384
 *   push_check: the encoder step state -> (state / freq << 24) + state % freq + start with its
 *               renormalisation, rans/rans.pyx:62-65 */
int flic_debug_div_check(int64_t n, uint64_t seed, int mode, uint64_t* mismatches, flic_cuda_stream_t stream);
int flic_debug_push_check(int64_t n, uint64_t seed, uint64_t* mismatches, flic_cuda_stream_t stream);

/* Bytes of device workspace flic_rans_encode needs (worst-case word scratch + scan temporaries). */
int64_t flic_encode_workspace_bytes(int64_t n_symbols, int64_t n_streams);

/* K1+K2+K4.  Replaces rans.encode, rans/rans.pyx:37-67, for every stream of the partition.
 *   stream_offsets  int64[n_streams+1], device, non-decreasing, [0] == 0, [n_streams] == n_symbols
 *   init_states     uint64[n_streams] or NULL (= 1<<32 each, trainer.py:310)
 *   packed          uint32[packed_capacity], device; n_symbols words always suffice
 *   word_offsets    int64[n_streams+1], device (out)
 *   final_states    uint64[n_streams], device (out)
 *   status          int32[n_streams], device (out) */
int flic_rans_encode(const float* x, const float* mean, const float* scale,
                     const int64_t* stream_offsets, int64_t n_streams, int64_t n_symbols,
                     const uint64_t* init_states, void* workspace, int64_t workspace_bytes,
                     uint32_t* packed, int64_t packed_capacity, int64_t* word_offsets,
                     uint64_t* final_states, int32_t* status, flic_cuda_stream_t stream);

/* K3.  Replaces rans.decode, rans/rans.pyx:69-110, for every stream.  All arrays are in FORWARD
 * order (the reversal the reference's caller does at trainer.py:317-318 is internal here).
 *   check_end  non-zero: flag FLIC_ST_BAD_END_STATE unless the stream ends at 1<<32 with all
 *              its words consumed. */
int flic_rans_decode(const uint32_t* packed, const int64_t* word_offsets,
                     const uint64_t* final_states, const float* mean, const float* scale,
                     const int64_t* stream_offsets, int64_t n_streams, float* x_out,
                     uint64_t* end_states, int32_t* status, int check_end,
                     flic_cuda_stream_t stream);

/* K3 with a continuation (ABI 2).  coder.Encode / coder.Decode (coder.py:18-38) carry ONE state
 * through the latent levels of a batch: encode level 0, 1, 2 from the state the previous level
 * ended in, decode level 2, 1, 0 from the state the later level's decode ended in.  rANS is LIFO, so
 * that decode order is also the dependency order of a real decompress (the prior of level k needs
 * the decoded level k + 1).  A chained stream is encoded with flic_rans_encode's init_states and
 * its levels' words concatenated in emission order (flic_gather_words); each level is then decoded
 * with this call, which starts from states_in with words_left_in[s] unread words at
 * packed[word_offsets[s] ...] (null: the stream's whole word_offsets range) and reports the state
 * and the unread count it stops at.  The word a later level's FIRST symbol pushed is pulled when
 * the earlier level's decode starts -- from the same array, which is what the reference's
 * per-level buffers get wrong.  check_end belongs on the last call (level 0).
 *   words_left_in   int64[n_streams], device, or null
 *   words_left_out  int64[n_streams], device (out), or null */
int flic_rans_decode_resume(const uint32_t* packed, const int64_t* word_offsets,
                            const uint64_t* states_in, const int64_t* words_left_in, const float* mean,
                            const float* scale, const int64_t* stream_offsets, int64_t n_streams,
                            float* x_out, uint64_t* end_states, int64_t* words_left_out, int32_t* status,
                            int check_end, flic_cuda_stream_t stream);

/* K4 for chained streams: dst[dst_starts[s] + i] = src[src_offsets[s] + i] for every word i of
 * stream s (src_offsets has n_streams + 1 entries).  Replaces the list appends of
 * coder.py:26 / trainer.py:321.  A destination outside [0, dst_capacity) flags the stream
 * (FLIC_ST_UNDERRUN in status, which may be null) and copies nothing. */
int flic_gather_words(const uint32_t* src, const int64_t* src_offsets, const int64_t* dst_starts,
                      int64_t n_streams, uint32_t* dst, int64_t dst_capacity, int32_t* status,
                      flic_cuda_stream_t stream);

/* K5.  Replaces AdditiveCouple.forward / .backward's elementwise tail, couplelib.py:49-52 and
 * :58-60, with Round from roundlib.py:18-38:  x[:, a_ch:] += direction * Round_nbits(t), in place.
 *   x  float32 (batch, channels, hw) contiguous;  t  float32 (batch, channels - a_ch, hw)
 *   direction  +1 forward (zb = xb + round(t)), -1 backward (xb = zb - round(t)) */
int flic_couple_add_round(float* x, const float* t, int64_t batch, int64_t channels, int64_t a_ch,
                          int64_t hw, int direction, int nbits, flic_cuda_stream_t stream);

/* Input quantisation, trainer.py:61,72 (ToTensor then Round(nbits=8)), and its inverse. */
int flic_u8_to_grid(const uint8_t* src, float* dst, int64_t n, flic_cuda_stream_t stream);
int flic_grid_to_u8(const float* src, uint8_t* dst, int64_t n, int32_t* status_word,
                    flic_cuda_stream_t stream);

/* N1.  Permute.forward/.backward, invertible.py:38-48, as a channel gather:
 *   dst[b, i, :] = src[b, perm[i], :]      perm int32[channels] on the device.
 * Permute.forward uses ids (P[i, ids[i]] = 1, invertible.py:34), .backward the inverse permutation. */
int flic_permute_channels(const float* src, float* dst, const int32_t* perm, int64_t batch,
                          int64_t channels, int64_t hw, flic_cuda_stream_t stream);

/* N1.  ExtendDim.forward (direction +1; extenddim.py:23-29) / .backward (-1; :31-37).
 * (batch, C, H, W) always names the UNSQUEEZED shape; the squeezed one is (batch, C*s*s, H/s, W/s). */
int flic_squeeze(const float* src, float* dst, int64_t batch, int64_t C, int64_t H, int64_t W,
                 int scale, int direction, flic_cuda_stream_t stream);

/* N4: the ideal code length, fused.
 * DLogistic.log_prob (distlib.py:40-55) for x, mean, logscale of shape (batch, per_item), with
 * the per-image sum of IDFlows.log_likelihood (flows.py:154-169) computed in the same pass:
 *   logp_out  float32[batch * per_item], device, or NULL
 *   sum_out   float32[batch], device, or NULL (natural-log units; the caller divides by H*W*C)
 * Floating point (not part of the bitstream): same float operations as the torch formula. */
int flic_dlogistic_log_prob(const float* x, const float* mean, const float* logscale, int64_t batch,
                            int64_t per_item, int nbits, float eps, float* logp_out, float* sum_out,
                            flic_cuda_stream_t stream);

/* DLogistic.sample (distlib.py:57-70) from caller-supplied uniforms u in (0, 1):
 * out = Round_nbits(log(u / (1 - u)) * exp(logscale) + mean). */
int flic_dlogistic_sample(const float* u, const float* mean, const float* logscale, int64_t n, int nbits,
                          float* out, flic_cuda_stream_t stream);

/* ------------------------------------------------------------------------------------------
 * Host entry points (copies inside)
 * ------------------------------------------------------------------------------------------ */

typedef struct flic_codec flic_codec;

/* A codec owns device workspace for up to max_symbols symbols / max_streams streams per call,
 * on `device`, plus its own CUDA streams.  Not thread-safe: one codec per calling thread. */
int flic_codec_create(int device, int64_t max_symbols, int64_t max_streams, flic_codec** out);
void flic_codec_destroy(flic_codec* codec);

/* Pinned host memory for the arrays below (optional; pageable memory works but copies slower). */
int flic_host_alloc(void** ptr, int64_t bytes);
void flic_host_free(void* ptr);

/* rans.encode over a partition; all pointers are HOST memory.  words_out needs capacity for the
 * result; n_symbols words always suffice.  *n_words_out = total words written. */
int flic_codec_encode(flic_codec* codec, const float* x, const float* mean, const float* scale,
                      const int64_t* stream_offsets, int64_t n_streams, uint32_t* words_out,
                      int64_t words_capacity, int64_t* word_offsets_out, uint64_t* states_out,
                      int32_t* status_out, int64_t* n_words_out);

/* rans.decode over a partition; all pointers are HOST memory, forward order. */
int flic_codec_decode(flic_codec* codec, const uint32_t* words, const int64_t* word_offsets,
                      const uint64_t* states, const float* mean, const float* scale,
                      const int64_t* stream_offsets, int64_t n_streams, float* x_out,
                      uint64_t* end_states_out, int32_t* status_out);

/* Measurement aid: the host<->device copies of flic_codec_encode followed by flic_codec_decode for
 * these buffers -- same chunks, same three CUDA streams -- with no kernel in between.  What the
 * host side of a box can move for this call pattern is the ceiling of the end-to-end number
 * (bench.py reports both).  words / n_words stand for the compressed payload. */
int flic_codec_probe_copies(flic_codec* codec, const float* x, const float* mean, const float* scale,
                            const int64_t* stream_offsets, int64_t n_streams, uint32_t* words,
                            int64_t n_words, float* x_out);

/* Exact argument-for-argument drop-ins for the reference's two functions (single stream,
 * caller-supplied state), HOST memory:
 *   encode(state, n, x_, mean_, scale_) -> (state, buffer)             rans/rans.pyx:37
 *   decode(state, buffer_, n, mean_, scale_) -> (state, message)       rans/rans.pyx:69
 * As in the reference, decode's buffer_/mean_/scale_ are REVERSED by the caller and message
 * comes back reversed (trainer.py:317-318).  buffer_out needs n words.  Returns FLIC_E_STATUS
 * with *status_out set when the stream is invalid (the cases where the reference raises). */
int flic_rans_encode_single(flic_codec* codec, uint64_t state, int64_t n, const float* x,
                            const float* mean, const float* scale, uint32_t* buffer_out,
                            int64_t* n_words_out, uint64_t* state_out, int32_t* status_out);
int flic_rans_decode_single(flic_codec* codec, uint64_t state, const uint32_t* buffer_reversed,
                            int64_t n_buffer, int64_t n, const float* mean_reversed,
                            const float* scale_reversed, float* message_out, uint64_t* state_out,
                            int32_t* status_out);

#ifdef __cplusplus
}
#endif
#endif /* FLIC_B200_H */
