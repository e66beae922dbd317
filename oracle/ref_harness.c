/*
 * ref_harness.c -- thin C entry points over the REFERENCE's own vendored,
 * patched libbz2 1.0.6 (third-party/bzip2-1.0.6.tar.gz).  TEST INFRASTRUCTURE.
 *
 * build_ref.sh compiles this file together with the reference's bzip2 sources
 * where they were extracted (a scratch directory outside the repository);
 * nothing from the tarball is copied into the repo.  The #include of
 * compress.c below is deliberate: generateMTFValues / sendMTFValues are
 * `static` there (bz/compress.c:120, :239) and the stage-wise parity tests
 * need their outputs.
 */
#include <string.h>
#include <stdlib.h>
#include <stdint.h>
#include "bzlib_private.h"
#include "compress.c"     /* from the extracted reference tarball (-I) */

#define API __attribute__((visibility("default")))

static void noop_functor(void *h) { (void)h; }

/* One stream, fed with a single BZ_FINISH action.  The functor must be set
 * AFTER BZ2_bzCompressInit, which NULLs it (bz/bzlib.c:211-212, :470). */
API int64_t s3ref_bz_compress(const uint8_t *in, uint64_t n, int level, uint8_t *out, uint64_t cap)
{
    bz_stream s; memset(&s, 0, sizeof s);
    if (BZ2_bzCompressInit(&s, level, 0, 30) != BZ_OK) return -1;
    s.block_close_functor = noop_functor;
    s.next_in = (char *)in; s.next_out = (char *)out;
    uint64_t in_left = n, out_left = cap; int rc;
    do {
        /* avail_* are 32-bit; BZ_FINISH demands avail_in stay consistent, so
         * inputs >= 4 GiB are outside this harness's domain */
        if (n > 0xFFFFFFF0ull) { BZ2_bzCompressEnd(&s); return -2; }
        s.avail_in = (unsigned)in_left;
        unsigned chunk = out_left > 0x40000000u ? 0x40000000u : (unsigned)out_left;
        s.avail_out = chunk;
        rc = BZ2_bzCompress(&s, BZ_FINISH);
        in_left = s.avail_in; out_left -= chunk - s.avail_out;
        if (rc != BZ_FINISH_OK && rc != BZ_STREAM_END) { BZ2_bzCompressEnd(&s); return -3; }
        if (rc == BZ_FINISH_OK && out_left == 0) { BZ2_bzCompressEnd(&s); return -4; }
    } while (rc != BZ_STREAM_END);
    BZ2_bzCompressEnd(&s);
    return (int64_t)(cap - out_left);
}

API int64_t s3ref_bz_decompress(const uint8_t *in, uint64_t n, uint8_t *out, uint64_t cap)
{
    bz_stream s; memset(&s, 0, sizeof s);
    if (BZ2_bzDecompressInit(&s, 0, 0) != BZ_OK) return -1;
    s.next_in = (char *)in; s.next_out = (char *)out;
    uint64_t in_left = n, out_left = cap; int rc;
    do {
        unsigned ci = in_left > 0x40000000u ? 0x40000000u : (unsigned)in_left;
        unsigned co = out_left > 0x40000000u ? 0x40000000u : (unsigned)out_left;
        s.avail_in = ci; s.avail_out = co;
        rc = BZ2_bzDecompress(&s);
        in_left -= ci - s.avail_in; out_left -= co - s.avail_out;
        if (rc != BZ_OK && rc != BZ_STREAM_END) { BZ2_bzDecompressEnd(&s); return -3; }
        if (rc == BZ_OK && (out_left == 0 || (in_left == 0 && s.avail_out == co))) { BZ2_bzDecompressEnd(&s); return -4; }
    } while (rc != BZ_STREAM_END);
    BZ2_bzDecompressEnd(&s);
    return (int64_t)(cap - out_left);
}

static EState *mk_state(const uint8_t *block, int32_t n, const uint8_t *in_use)
{
    EState *s = calloc(1, sizeof *s);
    size_t cap = 900000;
    s->arr1 = calloc(cap, 4);
    s->arr2 = calloc(cap + BZ_N_OVERSHOOT, 4);
    s->ftab = calloc(65537, 4);
    s->block = (UChar *)s->arr2; s->mtfv = (UInt16 *)s->arr1; s->ptr = s->arr1;
    s->nblock = n; s->workFactor = 30; s->blockSize100k = 9; s->verbosity = 0;
    memcpy(s->block, block, (size_t)n);
    if (in_use) for (int i = 0; i < 256; i++) s->inUse[i] = in_use[i];
    return s;
}
static void rm_state(EState *s) { free(s->arr1); free(s->arr2); free(s->ftab); free(s); }

/* BZ2_blockSort (bz/blocksort.c:1031) on one block; n <= 900000 */
API int32_t s3ref_block_sort(const uint8_t *block, int32_t n, uint32_t *ptr_out)
{
    EState *s = mk_state(block, n, NULL);
    BZ2_blockSort(s);
    memcpy(ptr_out, s->ptr, (size_t)n * 4);
    int32_t o = s->origPtr;
    rm_state(s);
    return o;
}

/* generateMTFValues + sendMTFValues (bz/compress.c:120, :239) given the sorted order.
 * Outputs: mtfv[nMTF], freq[258], selector[], len[6][258], and the emitted bits
 * (body only: from the mapping table to the last symbol) as bytes + bit count. */
API int32_t s3ref_mtf_huff(const uint8_t *block, int32_t n, const uint32_t *ptr, const uint8_t *in_use,
                           uint16_t *mtfv, int32_t *freq, uint8_t *selector, uint8_t *len,
                           uint8_t *bits, uint64_t bits_cap, uint64_t *nbits)
{
    EState *s = mk_state(block, n, in_use);
    memcpy(s->ptr, ptr, (size_t)n * 4);
    generateMTFValues(s);
    int32_t nmtf = s->nMTF;
    memcpy(mtfv, s->mtfv, (size_t)nmtf * 2);
    memcpy(freq, s->mtfFreq, 258 * sizeof(int32_t));
    s->zbits = malloc(2000000); s->numZ = 0;
    BZ2_bsInitWrite(s);
    sendMTFValues(s);
    int32_t nsel = (nmtf + BZ_G_SIZE - 1) / BZ_G_SIZE;
    memcpy(selector, s->selector, (size_t)nsel);
    memcpy(len, s->len, sizeof s->len);
    uint64_t nb = (uint64_t)s->numZ * 8 + (uint64_t)s->bsLive;
    bsFinishWrite(s);
    if ((uint64_t)s->numZ <= bits_cap) memcpy(bits, s->zbits, (size_t)s->numZ);
    *nbits = nb;
    free(s->zbits);
    rm_state(s);
    return nmtf;
}
