/*
 * starch3_b200.h -- C ABI of the B200-native starch3 compression hot path.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch / C++ types.
 * Every entry point names the reference interface it replaces
 * (hpp = /root/reference/include/starch3api.hpp, bz/ = the reference's vendored
 * third-party/bzip2-1.0.6.tar.gz).  All entry points run on the GPU; there is
 * no CPU fallback -- without a usable CUDA device s3g_init fails.
 *
 * Conventions: every function returns S3G_OK (0) or a negative S3G_E_* code;
 * s3g_last_error() returns a thread-local message for the last failure.
 * A context is bound to one device and one CUDA stream and is single-caller.
 */
#ifndef STARCH3_B200_H_
#define STARCH3_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define S3G_API __attribute__((visibility("default")))
#else
#define S3G_API
#endif

#define S3G_OK            0
#define S3G_E_CUDA       -1   /* CUDA runtime error / no device                      */
#define S3G_E_PARAM      -2   /* bad argument (hpp:842-848 maps these to EINVAL)      */
#define S3G_E_NOMEM      -3   /* allocation failed (hpp:176 ... ENOMEM)               */
#define S3G_E_MALFORMED  -4   /* BED line with fewer than three fields               */
#define S3G_E_CAPACITY   -5   /* caller-provided output buffer too small             */
#define S3G_E_LIMIT      -6   /* input exceeds a documented limit                     */

typedef struct s3g_ctx s3g_ctx;

/* Replaces Starch::initialize_out_compression_stream / initialize_bz_stream_ptr
 * (hpp:771-785, :819-855): creates the compressor state, here a device context. */
S3G_API int  s3g_init(int device, s3g_ctx **out);
/* Replaces Starch::delete_out_compression_stream / delete_bz_stream_ptr (hpp:787-801, :864-880). */
S3G_API void s3g_destroy(s3g_ctx *ctx);
S3G_API const char *s3g_last_error(void);
/* Run all kernels of this context on `cuda_stream` (a cudaStream_t); NULL = the context's own stream. */
S3G_API int  s3g_set_stream(s3g_ctx *ctx, void *cuda_stream);
/* Number of kernels launched by this context since creation (bench.py `gpu_launches`). */
S3G_API uint64_t s3g_launch_count(const s3g_ctx *ctx);
/* Times the block sort had to repeat its radix passes with peer-mask ranking because the keys left by the
 * ordered-atomic ranking were not ascending (bwt.cu, k_sweep).  Expected to stay 0; a diagnostic. */
S3G_API uint64_t s3g_sort_retries(const s3g_ctx *ctx);
/* How the last s3g_compress_bed on this context overlapped the upload: 0 = one piece (input below 48 MB, or a fallback),
 * n > 0 = n ranges by chromosome (two worker contexts), n < 0 = -n ranges chained at bzip2-block granularity (an input of
 * few chromosomes); DESIGN.md section 5b.  A diagnostic: the archive is the same bytes in every case. */
S3G_API int s3g_last_host_entry(const s3g_ctx *ctx);
/* Block-sort diagnostics: out[0] = bzip2 blocks sorted by the bucket form (bwt_bucket.cu) since the context was created,
 * out[1] = batches in which it handed blocks back to the radix form (a sub-bucket of more than 256 equal keys, or a
 * bucket beyond its shared-memory capacity), out[2] = s3g_sort_retries. */
S3G_API int s3g_sort_stats(const s3g_ctx *ctx, uint64_t out[3]);
/* Per-kernel timing with CUDA events on the launching stream.  s3g_profile(ctx, 1) starts
 * recording; s3g_profile_report synchronises and writes one line per kernel name
 * ("name\tlaunches\ttotal_ms\talgorithmic_bytes\n") into buf, then clears the records. */
S3G_API int  s3g_profile(s3g_ctx *ctx, int enable);
S3G_API int  s3g_profile_report(s3g_ctx *ctx, char *buf, uint64_t cap);
/* Restrict the events to launches of one kernel (name as printed by s3g_profile_report);
 * NULL or "" = every kernel.  Lets a caller time the dominant kernel inside a timed region
 * without paying two event records around each of the ~90 other launches. */
S3G_API int  s3g_profile_filter(s3g_ctx *ctx, const char *kernel_name);

/* One entry per chromosome stream, in input order.  Mirrors what
 * process_tf_buffer (hpp:393-407) is handed: current_chr, line_count, the
 * transformed buffer -- plus the declared-but-never-computed base counts
 * (hpp:61-62) and the position of the chromosome's bzip2 stream. */
typedef struct s3g_chrom {
    uint64_t name_off;         /* offset of the name in the INPUT buffer */
    uint32_t name_len;
    uint32_t n_blocks;         /* bzip2 blocks in this stream */
    uint64_t tf_off, tf_len;   /* transformed stream inside the tf buffer */
    int64_t  line_count;       /* hpp:503 */
    int64_t  bases_nonunique;  /* sum(stop-start) */
    int64_t  bases_unique;     /* bases covered by the union of the intervals */
    uint64_t bz_off, bz_len;   /* compressed stream inside the streams buffer */
} s3g_chrom;

typedef struct s3g_result {
    uint8_t   *archive;        /* host (pinned, owned by the context): complete archive (ARCHIVE_FORMAT.md);
                                  valid until the next compress call on ctx or s3g_destroy */
    uint64_t   archive_size;
    uint64_t   streams_off;    /* where the concatenated bzip2 streams start inside archive */
    s3g_chrom *chroms;         /* host */
    uint64_t   n_chroms;
    uint64_t   n_lines, n_blocks;
    uint64_t   tf_bytes;       /* total transformed bytes */
    uint64_t   dropped_tail_bytes; /* unterminated last line, dropped like produce_line hpp:181-190 */
    void      *d_streams;      /* device: concatenated bzip2 streams (valid until the next call on ctx) */
    uint64_t   streams_size;
    double     device_ms;      /* CUDA-event time, first kernel to last kernel */
    /* measured sizes of the intermediate stages (B_blk and M of the roofline formula) */
    uint64_t   rle_bytes;      /* bytes after RLE1, all blocks (sum of nblock) */
    uint64_t   mtf_symbols;    /* MTF/RUNA/RUNB symbols incl. EOB, all blocks (sum of nMTF) */
    /* CUDA-event time per stage of the device-resident call, ms: [0] tokenise + transform, [1] RLE1 + cut + CRC,
     * [2] block sort, [3] MTF + zero runs, [4] Huffman + pool, [5] assembly; zero after the pipelined host entry */
    double     stage_ms[8];
    /* input hardening: none of these changes the output, they tell the caller that the input is outside
     * the sorted-BED domain the reference assumes */
    uint64_t   unsorted_lines;     /* elements that start before the previous element of their chromosome */
    uint64_t   crlf_lines;         /* lines that end in "\r\n": the carriage return stays part of the line's last field */
    uint64_t   reappearing_chroms; /* streams whose chromosome name already opened an earlier stream (hpp:331 compares with
                                      the previous line only); 0 when the call did not build the archive */
} s3g_result;
#define S3G_STAGE_NAMES "tokenise+transform", "rle1+cut+crc", "blocksort", "mtf", "huffman", "assemble"

/*
 * The whole hot path: tokenise + transform + per-chromosome bzip2 + container.
 * Replaces the produce_line / consume_line / update_chr / consume_tf_buffer
 * pipeline (hpp:158-391, started at /root/reference/src/starch3.cpp:36-59) and
 * the BZ2_bzCompress loop process_tf_buffer (hpp:393) was evidently meant to run.
 *   bed, n          host buffer with the sorted BED text
 *   block_size_100k 1..9 (the reference hard-codes 9, hpp:837)
 *   note            metadata note (hpp:803-809), may be NULL
 */
S3G_API int s3g_compress_bed(s3g_ctx *ctx, const uint8_t *bed, uint64_t n, int block_size_100k,
                     const char *note, s3g_result *res);
/* Same, input already resident in device memory (16-byte aligned); the archive
 * is still assembled on the host unless `want_archive` is 0, in which case only
 * d_streams / chroms are produced (the device-resident measurement of bench.py). */
S3G_API int s3g_compress_bed_device(s3g_ctx *ctx, const void *d_bed, uint64_t n, int block_size_100k,
                            const char *note, int want_archive, s3g_result *res);
S3G_API void s3g_result_free(s3g_result *res);
/* Bounded-memory ingestion: the same archive from input handed over in pieces of any size -- replaces the
 * byte-wise reader of produce_line (hpp:158-199) over stdin or a file (hpp:728-736, :890-905).  range_bytes (0 = 256 MiB)
 * bounds what is resident: one range in pinned host memory, the chromosome still open plus one range on the device, and
 * the compressed streams so far.  A chromosome is compressed once the input has moved on to the next one. */
S3G_API int s3g_stream_begin(s3g_ctx *ctx, int block_size_100k, const char *note, uint64_t range_bytes);
S3G_API int s3g_stream_write(s3g_ctx *ctx, const uint8_t *bed, uint64_t n);
S3G_API int s3g_stream_end(s3g_ctx *ctx, s3g_result *res);
/* Copy the device-resident streams of the last compress call (s3g_result.d_streams) to the host. */
S3G_API int s3g_read_streams(s3g_ctx *ctx, uint8_t *dst, uint64_t cap, uint64_t *n);

/* ---- one archive from several GPUs ----
 * The path cut into phases that a host runs on every GPU, with small exchanges in between: one process per GPU
 * (starch3_b200/multigpu.py, torch.distributed) or one process with a context per device (s3g_multi_*, below).
 * GPU r owns a newline-aligned byte range of the input; update_transformation_state (hpp:428-504) reads only the
 * previous line, so the range is handed over together with the one line before it (the "halo").
 * All pointers named d_* are device pointers on the context's device. */
typedef struct s3g_shard_summary {
    uint64_t n_lines;            /* lines of the range (the halo line not counted) */
    int64_t  tail_max;           /* largest stop since the last chromosome change, INT64_MIN if none: uniqueBases of the
                                    next range needs the largest stop of ALL earlier lines of its first chromosome */
    uint32_t continues;          /* the first line of the range has the halo line's chromosome (hpp:331) */
    uint32_t single_piece;       /* no chromosome change inside the range */
    uint64_t dropped_tail_bytes; /* unterminated last line (hpp:181-190); only the last range can have one */
    uint64_t tf_bytes;           /* transformed bytes the range will write in phase 2 */
} s3g_shard_summary;
/* phase 1: tokenizer over d_range[0, n); its first halo_bytes bytes are the line before the range (0: none). */
S3G_API int s3g_shard_tokenize(s3g_ctx *ctx, const void *d_range, uint64_t n, uint64_t halo_bytes, s3g_shard_summary *out);
/* phase 2: transform + statistics.  carry_max = largest stop of all earlier lines of the range's first chromosome on
 * other GPUs (INT64_MIN if the range does not continue one).  pieces = the chromosome pieces of the range in order
 * (name_off relative to d_range; tf_off relative to *d_tf; bz_* unused); *d_tf stays valid until the next call. */
S3G_API int s3g_shard_transform(s3g_ctx *ctx, int64_t carry_max, s3g_chrom *pieces, uint64_t cap, uint64_t *n_pieces,
                                void **d_tf, uint64_t *tf_len);
/* phase 2, fused with the exchange that follows it: the transform kernel stores the range's transformed bytes at dst_off
 * into EVERY buffer of peer_bufs (device addresses valid on this GPU: its own copy of the transformed buffer first, then the
 * peers' copies through their NVLink-mapped pointers, e.g. torch symmetric memory or cudaIpc / cudaDeviceEnablePeerAccess
 * mappings; 16-byte aligned, at most 8).  dst_off = sum of tf_bytes of the ranges before this one (s3g_shard_summary).
 * multicast_buf: the NVLS multicast address of the same buffers (a store to it is replicated by the NVSwitch into every
 * GPU's copy; torch symmetric memory: multicast_ptr), or 0: then the kernel stores into each buffer in turn.
 * When every GPU has returned from this call, every buffer holds all transformed bytes: no all-gather follows. */
S3G_API int s3g_shard_transform_peers(s3g_ctx *ctx, int64_t carry_max, s3g_chrom *pieces, uint64_t cap, uint64_t *n_pieces,
                                      const uint64_t *peer_bufs, uint32_t n_peers, uint64_t multicast_buf, uint64_t dst_off,
                                      uint64_t *tf_len);
/* phase 3: RLE1 lengths + block cut (bz/bzlib.c:225-338, :370-412) over ALL transformed bytes, stream s =
 * d_tf_all[soff[s], soff[s+1]) (soff on the host).  Every GPU computes the same plan; nblock / stream_of describe
 * its blocks in archive order.  d_tf_all must stay valid until s3g_shard_compress returns. */
S3G_API int s3g_shard_plan(s3g_ctx *ctx, const void *d_tf_all, uint64_t tf_total, const uint64_t *soff, uint64_t n_streams,
                           int block_size_100k, uint64_t *n_blocks, uint32_t *nblock, uint32_t *stream_of, uint64_t cap);
/* phase 4: RLE1 bytes, CRC, BZ2_blockSort, MTF, Huffman of blocks [b_lo, b_hi) of the plan; per-block results out. */
S3G_API int s3g_shard_compress(s3g_ctx *ctx, uint64_t b_lo, uint64_t b_hi, uint64_t *n_bits, uint32_t *crc, uint32_t *n_mtf);
/* phase 5: given bit length and CRC of EVERY block, place the own blocks (and the "BZh9" header / trailer +
 * combined CRC, bz/compress.c:607-609, :622-628, :657-666, of the streams that begin / end among them) into a byte
 * string that covers bytes [byte_lo, byte_hi) of the concatenated streams.  Its first and last byte may be shared
 * with the neighbouring GPUs' strings (blocks are not byte aligned): the host ORs them.  stream_off / stream_len
 * (n_streams each, may be NULL) = the layout of the streams. */
S3G_API int s3g_shard_assemble(s3g_ctx *ctx, const uint64_t *n_bits_all, const uint32_t *crc_all, uint64_t b_lo, uint64_t b_hi,
                               void **d_bytes, uint64_t *byte_lo, uint64_t *byte_hi, uint64_t *stream_off, uint64_t *stream_len);
/* phase 5, gather fused: store the byte string s3g_shard_assemble left on this GPU at byte_lo of `gather_buf` -- a device
 * address valid on this GPU, typically the NVLink-mapped pointer of the buffer on the GPU that collects the archive.  The
 * buffer must be zero where nothing was placed yet: the string's end bytes are ORed in (shared with the neighbours). */
S3G_API int s3g_shard_place(s3g_ctx *ctx, uint64_t gather_buf, uint64_t byte_lo, uint64_t byte_hi);
/* The phases above driven by ONE process: ctxs[k] = a context per GPU (s3g_init(device_k, &ctxs[k]); a device may appear
 * more than once), a host thread each.  The small tables are exchanged in host memory; the two bulk exchanges are stores
 * over NVLink through peer access (s3g_shard_transform_peers, s3g_shard_place).  The archive -- owned by ctxs[0] like the
 * archive of s3g_compress_bed -- is the same bytes as the single-GPU archive.  What `starch3 --devices=0,1,...` calls. */
S3G_API int s3g_multi_compress_bed(s3g_ctx **ctxs, int n_ctx, const uint8_t *bed, uint64_t n, int block_size_100k,
                                   const char *note, s3g_result *res);
/* CUDA-event time per stage (indices as s3g_result.stage_ms) of the phases run on ctx since s3g_shard_tokenize. */
S3G_API int s3g_stage_times(s3g_ctx *ctx, double *stage_ms8);
/* Host logic only (no device needed): how a batch of n_blocks bzip2 blocks is dealt to the MTF (stage 3) or Huffman (stage 4)
 * kernels -- chunks of consecutive blocks, ctas[i] CTAs per block (0: the Huffman form with two 512-thread CTAs per SM and one
 * CTA per block).  DESIGN.md section 4, plan_chunks.  The tests check that the chunks tile the batch and fit the GPU. */
/* Host logic only: the bit layout of one step of the chained entries.  state[4] = {a stream is open, its
 * bits so far, its combined CRC so far, bytes of the closed streams}, carried from step to step (all zero before the first).
 * The step has n_streams streams and n_blocks blocks in stream order (stream_of ascending from 0), the first n_final of them
 * final; its first stream continues the open stream when first_continues.  Out: block_pos[n_final] = bit position of every final
 * block in the streams buffer; patch[2 i], patch[2 i + 1] = bit position and 32-bit word of the stream headers and trailers;
 * stream_start[s] = byte offset a stream starts at in this step (~0: it continues), stream_len[s] = its byte length if it is closed
 * in this step (0: still open). */
S3G_API int s3g_chain_layout(uint64_t *state, int block_size_100k, uint64_t n_streams, int first_continues, const uint32_t *stream_of,
                             const uint64_t *n_bits, const uint32_t *crc, uint64_t n_blocks, uint64_t n_final, uint64_t *block_pos,
                             uint64_t *patch, uint64_t patch_cap, uint64_t *n_patch, uint64_t *stream_start, uint64_t *stream_len);
S3G_API int s3g_batch_chunks(uint64_t n_blocks, int stage, uint64_t *first, uint64_t *count, uint32_t *ctas, uint64_t cap, uint64_t *n_chunks);

/* ---- stage entry points (host buffers in / out), for the parity tests ---- */

/* Kernel (1): produce_line + consume_line tokenizer (hpp:158-199, :220-309).
 * Outputs (each may be NULL): line_start[n_lines+1], start[], stop[], rem_off[]
 * (offset of the remainder relative to the line start; == line length when the
 * line has no fourth field), chrom_change[] (1 where strcmp(chr, previous chr) != 0, hpp:331). */
S3G_API int s3g_tokenize(s3g_ctx *ctx, const uint8_t *bed, uint64_t n, uint64_t cap_lines, uint64_t *n_lines,
                 uint64_t *line_start, int64_t *start, int64_t *stop, uint32_t *rem_off, uint8_t *chrom_change);

/* Kernels (1)+(2): update_transformation_state (hpp:428-504) for all lines. */
S3G_API int s3g_transform(s3g_ctx *ctx, const uint8_t *bed, uint64_t n, uint8_t *tf, uint64_t tf_cap, uint64_t *tf_len,
                  s3g_chrom *chroms, uint64_t chrom_cap, uint64_t *n_chroms, uint64_t *dropped_tail_bytes);

/* Kernel (3a): RLE1 + block CRC + block cut (bz/bzlib.c:225-338, :370-412) of ONE stream. */
typedef struct s3g_blockdesc {
    uint64_t in_start, in_end; /* input bytes committed to this block */
    uint32_t nblock;           /* bytes after RLE1 */
    uint32_t crc;              /* finalised block CRC (bz/compress.c:606) */
    uint8_t  in_use[256];      /* bz/bzlib.c:232, :247 */
} s3g_blockdesc;
S3G_API int s3g_rle1(s3g_ctx *ctx, const uint8_t *in, uint64_t n, int block_size_100k,
             s3g_blockdesc *desc, uint64_t desc_cap, uint64_t *n_blocks,
             uint8_t *rle_out, uint64_t rle_cap);

/* Kernel (3b): BZ2_blockSort (bz/blocksort.c:1031-1089) for a batch of blocks.
 * blocks = concatenated block bytes, off[n_blocks+1] their boundaries (each block <= 900000 bytes).
 * ptr_out receives the sorted rotation starts at the same offsets; orig_ptr[b] as bz/blocksort.c:1083-1086. */
S3G_API int s3g_bwt(s3g_ctx *ctx, const uint8_t *blocks, const uint64_t *off, uint64_t n_blocks,
            uint32_t *ptr_out, int32_t *orig_ptr);

/* Kernel (3c): generateMTFValues (bz/compress.c:120-231) of one block given its sorted order.
 * mtfv needs n+1 slots; freq[258]. */
S3G_API int s3g_mtf(s3g_ctx *ctx, const uint8_t *block, uint32_t n, const uint32_t *ptr, const uint8_t *in_use,
            uint16_t *mtfv, uint32_t *n_mtf, int32_t *freq);

/* Kernel (3d): sendMTFValues (bz/compress.c:239-598) of one block: table selection and emission.
 * selector[18002], len[6*258]; bits receives the block body (mapping table .. last symbol), MSB first. */
S3G_API int s3g_huff(s3g_ctx *ctx, const uint16_t *mtfv, uint32_t n_mtf, const int32_t *freq, const uint8_t *in_use,
             int32_t *n_groups, int32_t *n_selectors, uint8_t *selector, uint8_t *len,
             uint8_t *bits, uint64_t bits_cap, uint64_t *n_bits);

/* Kernels (3a-e): one complete bzip2 stream; the unit the reference's patched
 * BZ2_bzCompressInit / BZ2_bzCompress(BZ_FINISH) / BZ2_bzCompressEnd trio (bz/bzlib.h:103-117)
 * produces for one chromosome. */
S3G_API int s3g_bz_compress(s3g_ctx *ctx, const uint8_t *in, uint64_t n, int block_size_100k,
                    uint8_t *out, uint64_t out_cap, uint64_t *out_len);

/* ---- the decoder path: archive -> BED, on the GPU ----
 * The reference has no decoder (its libbz2 carries BZ2_bzDecompress, bz/bzlib.c:551-900 + bz/decompress.c:106-646, which
 * nothing calls); these entry points are what an `unstarch3` would bind.  The inverse of update_transformation_state
 * (hpp:428-504) is the one ARCHIVE_FORMAT.md states. */
typedef struct s3g_decode_info {
    uint64_t n_streams, n_blocks;  /* chromosome streams / bzip2 blocks decoded */
    uint64_t tf_bytes;             /* transformed bytes (all streams) */
    double   device_ms;            /* CUDA-event time, upload to last kernel */
    void    *d_bed;                /* device: the BED text (valid until the next call on ctx) */
} s3g_decode_info;
/* A whole archive (ARCHIVE_FORMAT.md) back to BED text.  bed may be NULL (only *bed_len and info are produced).
 * Every block CRC and every stream's combined CRC is checked (bz/bzlib.c:843-866); a mismatch is S3G_E_PARAM. */
S3G_API int s3g_decompress_archive(s3g_ctx *ctx, const uint8_t *archive, uint64_t n, uint8_t *bed, uint64_t bed_cap,
                                   uint64_t *bed_len, s3g_decode_info *info);
/* One complete bzip2 stream -> its bytes; what BZ2_bzDecompressInit / BZ2_bzDecompress / BZ2_bzDecompressEnd
 * (bz/bzlib.h:121-135) do for one stream. */
S3G_API int s3g_bz_decompress(s3g_ctx *ctx, const uint8_t *in, uint64_t n, uint8_t *out, uint64_t out_cap, uint64_t *out_len);
/* One transformed stream -> the BED lines of chromosome `name`. */
S3G_API int s3g_inverse_transform(s3g_ctx *ctx, const uint8_t *tf, uint64_t n, const uint8_t *name, uint32_t name_len,
                                  uint8_t *bed, uint64_t bed_cap, uint64_t *bed_len);

#ifdef __cplusplus
}
#endif
#endif /* STARCH3_B200_H_ */
