// bzshim.cu -- the libbz2-shaped front of include/s3g_bzlib.h: BZ2_bzCompressInit / BZ2_bzCompress / BZ2_bzCompressEnd
// (bz/bzlib.c:148-220, :413-473, :476-500 of the reference's vendored libbz2) over s3g_bz_compress.  Host code only.
#include <new>
#include <vector>
#include "common.cuh"
#include "../../include/s3g_bzlib.h"

namespace {
enum Mode { M_IDLE = 1, M_RUNNING = 2, M_FINISHING = 4 };     // bz/bzlib_private.h:199-202
struct ShimState {
    s3g_bz_stream *strm;
    s3g_ctx *ctx;
    int level, mode;
    unsigned avail_in_expect;
    std::vector<unsigned char> in, out;
    size_t out_pos;
    bool compressed;
};
inline void add64(unsigned &lo, unsigned &hi, unsigned n) { unsigned old = lo; lo += n; if (lo < old) hi++; }
// take whatever input the caller offers
bool consume(ShimState *s)
{
    s3g_bz_stream *z = s->strm;
    if (!z->avail_in) return false;
    s->in.insert(s->in.end(), reinterpret_cast<unsigned char *>(z->next_in), reinterpret_cast<unsigned char *>(z->next_in) + z->avail_in);
    add64(z->total_in_lo32, z->total_in_hi32, z->avail_in);
    z->next_in += z->avail_in; z->avail_in = 0;
    return true;
}
bool produce(ShimState *s)
{
    s3g_bz_stream *z = s->strm;
    size_t left = s->out.size() - s->out_pos, k = left < z->avail_out ? left : z->avail_out;
    if (!k) return false;
    memcpy(z->next_out, s->out.data() + s->out_pos, k);
    s->out_pos += k; z->next_out += k; z->avail_out -= (unsigned)k;
    add64(z->total_out_lo32, z->total_out_hi32, (unsigned)k);
    return true;
}
}  // namespace

extern "C" {

int s3g_BZ2_bzCompressInit(s3g_bz_stream *strm, int blockSize100k, int verbosity, int workFactor)
{
    (void)verbosity;
    if (!strm || blockSize100k < 1 || blockSize100k > 9 || workFactor < 0 || workFactor > 250) return S3G_BZ_PARAM_ERROR;   // bz/bzlib.c:160-163
    ShimState *s = new (std::nothrow) ShimState();
    if (!s) return S3G_BZ_MEM_ERROR;
    int dev = 0;
    if (const char *e = getenv("S3G_DEVICE")) dev = atoi(e);
    if (s3g_init(dev, &s->ctx) != S3G_OK) { delete s; return S3G_BZ_CONFIG_ERROR; }
    s->strm = strm; s->level = blockSize100k; s->mode = M_RUNNING; s->out_pos = 0; s->compressed = false; s->avail_in_expect = 0;
    strm->state = s;
    strm->total_in_lo32 = strm->total_in_hi32 = strm->total_out_lo32 = strm->total_out_hi32 = 0;
    strm->block_close_functor = NULL;                      // as the patched Init does (bz/bzlib.c:211): the caller sets it afterwards
    return S3G_BZ_OK;
}

int s3g_BZ2_bzCompress(s3g_bz_stream *strm, int action)
{
    if (!strm) return S3G_BZ_PARAM_ERROR;
    ShimState *s = static_cast<ShimState *>(strm->state);
    if (!s || s->strm != strm) return S3G_BZ_PARAM_ERROR;
    switch (s->mode) {
        case M_IDLE: return S3G_BZ_SEQUENCE_ERROR;
        case M_RUNNING:
            if (action == S3G_BZ_RUN) return consume(s) ? S3G_BZ_RUN_OK : S3G_BZ_PARAM_ERROR;
            if (action != S3G_BZ_FINISH) return S3G_BZ_PARAM_ERROR;         // BZ_FLUSH would cut a block where libbz2 would: not offered
            s->avail_in_expect = strm->avail_in;
            s->mode = M_FINISHING;
            /* fall through */
        case M_FINISHING: {
            if (action != S3G_BZ_FINISH) return S3G_BZ_SEQUENCE_ERROR;
            if (s->avail_in_expect != strm->avail_in) return S3G_BZ_SEQUENCE_ERROR;
            bool progress = consume(s);
            s->avail_in_expect = 0;
            if (!s->compressed) {
                uint64_t cap = s->in.size() + s->in.size() / 50 + 1024, len = 0;
                s->out.resize(cap);
                int rc = s3g_bz_compress(s->ctx, s->in.data(), s->in.size(), s->level, s->out.data(), cap, &len);
                if (rc == S3G_E_CAPACITY) { s->out.resize(len); cap = len; rc = s3g_bz_compress(s->ctx, s->in.data(), s->in.size(), s->level, s->out.data(), cap, &len); }
                if (rc != S3G_OK) return rc == S3G_E_NOMEM ? S3G_BZ_MEM_ERROR : S3G_BZ_CONFIG_ERROR;
                s->out.resize(len);
                s->compressed = true; progress = true;
                std::vector<unsigned char>().swap(s->in);
            }
            progress = produce(s) || progress;
            if (!progress) return S3G_BZ_SEQUENCE_ERROR;
            if (s->out_pos < s->out.size()) return S3G_BZ_FINISH_OK;
            s->mode = M_IDLE;
            if (strm->block_close_functor) strm->block_close_functor(strm->handler);     // once per stream end (bz/bzlib.c:470)
            return S3G_BZ_STREAM_END;
        }
    }
    return S3G_BZ_OK;
}

int s3g_BZ2_bzCompressEnd(s3g_bz_stream *strm)
{
    if (!strm) return S3G_BZ_PARAM_ERROR;
    ShimState *s = static_cast<ShimState *>(strm->state);
    if (!s || s->strm != strm) return S3G_BZ_PARAM_ERROR;
    s3g_destroy(s->ctx);
    delete s;
    strm->state = NULL;
    return S3G_BZ_OK;
}

}  // extern "C"
