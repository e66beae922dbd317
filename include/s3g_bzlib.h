/*
 * s3g_bzlib.h -- a libbz2-shaped front for the GPU compressor ("Compressor C ABI (as patched)").
 *
 * The reference initialises a bz_stream of its vendored, patched libbz2 (bz/bzlib.h:48-69: the stock fields plus
 * `handler` and `block_close_functor`) with BZ2_bzCompressInit (starch3api.hpp:835-837) and would feed it with
 * BZ2_bzCompress / BZ2_bzCompressEnd (bz/bzlib.h:103-117).  These three functions keep that struct layout, those
 * signatures and return codes, and produce the bytes libbz2 produces for the same input and block size -- on the GPU
 * (s3g_bz_compress).  Differences, all allowed by the interface: output appears only once BZ_FINISH has been asked for
 * (a stream is compressed as one batch), BZ_FLUSH is refused with BZ_PARAM_ERROR, `bzalloc` / `bzfree` are not used, and
 * `block_close_functor` -- which the patch calls once per STREAM end (bz/bzlib.c:470) -- is called only if non-NULL.
 * The CUDA device is $S3G_DEVICE (default 0).  There is no CPU fallback: without a device Init returns BZ_CONFIG_ERROR.
 *
 * Define S3G_BZLIB_NAMES before including to get the libbz2 names (bz_stream, BZ2_bzCompressInit, ...) as aliases.
 */
#ifndef S3G_BZLIB_H_
#define S3G_BZLIB_H_

#ifdef __cplusplus
extern "C" {
#endif

#define S3G_BZ_RUN               0
#define S3G_BZ_FLUSH             1
#define S3G_BZ_FINISH            2
#define S3G_BZ_OK                0
#define S3G_BZ_RUN_OK            1
#define S3G_BZ_FLUSH_OK          2
#define S3G_BZ_FINISH_OK         3
#define S3G_BZ_STREAM_END        4
#define S3G_BZ_SEQUENCE_ERROR    (-1)
#define S3G_BZ_PARAM_ERROR       (-2)
#define S3G_BZ_MEM_ERROR         (-3)
#define S3G_BZ_CONFIG_ERROR      (-9)

typedef struct {
    char *next_in;
    unsigned int avail_in;
    unsigned int total_in_lo32;
    unsigned int total_in_hi32;

    char *next_out;
    unsigned int avail_out;
    unsigned int total_out_lo32;
    unsigned int total_out_hi32;

    void *state;

    void *(*bzalloc)(void *, int, int);
    void (*bzfree)(void *, void *);
    void *opaque;

    void *handler;                         /* the reference's patch: bz/bzlib.h:66-67 */
    void (*block_close_functor)(void *);
} s3g_bz_stream;

#if defined(__GNUC__)
#define S3G_BZ_API __attribute__((visibility("default")))
#else
#define S3G_BZ_API
#endif

S3G_BZ_API int s3g_BZ2_bzCompressInit(s3g_bz_stream *strm, int blockSize100k, int verbosity, int workFactor);
S3G_BZ_API int s3g_BZ2_bzCompress(s3g_bz_stream *strm, int action);
S3G_BZ_API int s3g_BZ2_bzCompressEnd(s3g_bz_stream *strm);

#ifdef S3G_BZLIB_NAMES
typedef s3g_bz_stream bz_stream;
#define BZ2_bzCompressInit s3g_BZ2_bzCompressInit
#define BZ2_bzCompress     s3g_BZ2_bzCompress
#define BZ2_bzCompressEnd  s3g_BZ2_bzCompressEnd
#define BZ_RUN S3G_BZ_RUN
#define BZ_FLUSH S3G_BZ_FLUSH
#define BZ_FINISH S3G_BZ_FINISH
#define BZ_OK S3G_BZ_OK
#define BZ_RUN_OK S3G_BZ_RUN_OK
#define BZ_FLUSH_OK S3G_BZ_FLUSH_OK
#define BZ_FINISH_OK S3G_BZ_FINISH_OK
#define BZ_STREAM_END S3G_BZ_STREAM_END
#define BZ_SEQUENCE_ERROR S3G_BZ_SEQUENCE_ERROR
#define BZ_PARAM_ERROR S3G_BZ_PARAM_ERROR
#define BZ_MEM_ERROR S3G_BZ_MEM_ERROR
#define BZ_CONFIG_ERROR S3G_BZ_CONFIG_ERROR
#endif

#ifdef __cplusplus
}
#endif
#endif /* S3G_BZLIB_H_ */
