/*
 * crdr_rans -- host range coder of the CRDR bitstream (libcrdr_rans.so, plain C ABI).
 *
 * The reference reaches its coder through CompressAI 1.2.4's pybind module `compressai.ans`
 * (RansEncoder.encode_with_indexes inside EntropyModel.compress; RansDecoder.set_stream /
 * decode_stream at src/models/subnet/context_model/minnen20_charm_context_model.py:201-202,222-224)
 * and `compressai._CXX.pmf_to_quantized_cdf` (EntropyModel._pmf_to_cdf).  Those calls marshal Python
 * lists; these entry points take flat int32 arrays instead (zero copy from numpy / pinned torch
 * buffers) and add batch variants that code independent streams on a thread pool.
 * The byte format is the reference's: rANS, 64-bit state, 32-bit words, 16-bit precision, 4-bit bypass.
 */
#ifndef CRDR_RANS_H_
#define CRDR_RANS_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* CDF tables: `cdfs` is [n_cdf][cdf_stride] int32, `cdf_sizes[i]` entries of row i are valid,
 * `offsets[i]` is the symbol value of bin 0. */
typedef struct {
  const int32_t* cdfs;
  int32_t cdf_stride;
  const int32_t* cdf_sizes;
  const int32_t* offsets;
  int32_t n_cdf;
  /* Handle from crdr_rans_tables_prepare() or NULL.  A prepared handle owns private copies of the arrays and the
   * per-table acceleration structures (reciprocal frequencies, decoder bucket index); with NULL they are rebuilt
   * for every call.  Nothing is cached by address, so rebuilt or reloaded tables can never be confused. */
  const void* prepared;
} crdr_cdf_tables;

/* Copies the tables and builds the coder's acceleration structures once; the arrays of `t` may be freed or
 * overwritten afterwards.  Returns NULL on invalid input.  Free with crdr_rans_tables_free(). */
void* crdr_rans_tables_prepare(const crdr_cdf_tables* t);
void crdr_rans_tables_free(void* prepared);

/* pmf[n] (float32) -> cdf[n+1] (uint32, cdf[n] == 2^precision, every bin non-empty). 0 on success. */
int crdr_pmf_to_quantized_cdf(const float* pmf, int64_t n, int32_t precision, uint32_t* cdf);

/* Encode n symbols; returns the stream length in bytes, or -(needed bytes) if out_cap is too small,
 * or INT64_MIN on invalid input (index out of range). */
int64_t crdr_rans_encode_with_indexes(const int32_t* symbols, const int32_t* indexes, int64_t n,
                                      const crdr_cdf_tables* t, uint8_t* out, int64_t out_cap);
/* `count` independent streams coded concurrently; stream i uses symbols[i]/indexes[i] (n[i] entries) and
 * writes to out[i] (capacity out_cap[i]); lengths[i] receives the per-stream return value. */
int crdr_rans_encode_batch(int32_t count, const int32_t* const* symbols, const int32_t* const* indexes,
                           const int64_t* n, const crdr_cdf_tables* t, uint8_t* const* out, const int64_t* out_cap,
                           int64_t* lengths, int32_t threads);

/* Same with the compact arrays the CUDA kernels write for the coder (crdr_gauss_desc.symbols16 / indexes8). */
int crdr_rans_encode_batch_i16u8(int32_t count, const int16_t* const* symbols, const uint8_t* const* indexes,
                                 const int64_t* n, const crdr_cdf_tables* t, uint8_t* const* out, const int64_t* out_cap,
                                 int64_t* lengths, int32_t threads);
/* Coder thread pool of this process: persistent workers pinned to the LOCAL_RANK-th of LOCAL_WORLD_SIZE equal slices
 * of the CPUs the process may use (CRDR_CODER_THREADS caps the size, CRDR_CODER_PIN=0 disables pinning).  The
 * `threads` argument of the batch calls is an upper bound (0 = the whole pool). */
int crdr_rans_pool_info(int32_t* threads, int32_t* first_cpu);

void* crdr_rans_decoder_new(void);
void crdr_rans_decoder_free(void* dec);
/* Copies the stream; subsequent decode_stream calls continue from the same coder state. */
int crdr_rans_decoder_set_stream(void* dec, const uint8_t* stream, int64_t nbytes);
/* Same without the copy: the decoder reads the caller's bytes, which must stay valid and unchanged until the next
 * set_stream* call on this decoder or its release (the Python front-end keeps a reference to the bytes object).
 * Streams whose length is not a multiple of 4 or whose address is not 4-byte aligned are copied as above. */
int crdr_rans_decoder_set_stream_view(void* dec, const uint8_t* stream, int64_t nbytes);
int crdr_rans_decoder_decode_stream(void* dec, const int32_t* indexes, int64_t n, const crdr_cdf_tables* t,
                                    int32_t* out);
/* decoders[i] decodes n[i] symbols with indexes[i] into out[i], concurrently. */
int crdr_rans_decode_batch(int32_t count, void* const* decoders, const int32_t* const* indexes, const int64_t* n,
                           const crdr_cdf_tables* t, int32_t* const* out, int32_t threads);

/* Same with the uint8 table indexes the CUDA kernels write for the coder (crdr_gauss_desc.indexes8). */
int crdr_rans_decode_batch_u8(int32_t count, void* const* decoders, const uint8_t* const* indexes, const int64_t* n,
                              const crdr_cdf_tables* t, int32_t* const* out, int32_t threads);

#ifdef __cplusplus
}
#endif
#endif
