// api.cu -- the C ABI (include/starch3_b200.h) and the host-side orchestration of the
// GPU pipeline.  Host code is plain C++; nothing here computes on the CPU except the
// few-hundred-byte archive header (ARCHIVE_FORMAT.md).
#include <cstdarg>
#include <ctime>
#include <algorithm>
#include <new>
#include "common.cuh"

namespace s3g {

static thread_local char g_err[1024] = "";

void set_error(const char *fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}

int check_launch(const char *what)
{
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("kernel launch failed (%s): %s", what, cudaGetErrorString(e)); return S3G_E_CUDA; }
    return S3G_OK;
}

int prof_begin(Ctx *ctx, const char *name)
{
    double bytes = ctx->prof_next_bytes;
    ctx->prof_next_bytes = 0;
    if (!ctx->prof) return -1;
    if (!ctx->prof_filter.empty() && ctx->prof_filter != name) return -1;
    while (ctx->prof_used + 2 > ctx->prof_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return -1; }
        ctx->prof_pool.push_back(e);
    }
    Ctx::ProfRec r;
    r.name = name; r.e0 = ctx->prof_pool[ctx->prof_used++]; r.e1 = ctx->prof_pool[ctx->prof_used++]; r.bytes = bytes;
    cudaEventRecord(r.e0, ctx->stream);
    ctx->prof_recs.push_back(r);
    return (int)ctx->prof_recs.size() - 1;
}

void prof_end(Ctx *ctx, int idx)
{
    if (idx >= 0) cudaEventRecord(ctx->prof_recs[idx].e1, ctx->stream);
}

static DevBuf *const *all_bufs(Ctx *c, size_t *n)
{
    static thread_local DevBuf *list[64];
    size_t k = 0;
#define B(x) list[k++] = &c->x
    B(bed); B(tile_cnt); B(line_start); B(start); B(stop); B(rem_off); B(flags); B(line_tf_off); B(chrom_first);
    B(scan_a); B(scan_b); B(scan_c); B(scalars); B(tf); B(chroms); B(stat_a); B(stat_b); B(soff);
    B(rle_carry); B(rle_ebase); B(blocks); B(blk_prov); B(blk_bytes); B(in_use); B(seq_map); B(stream_tab);
    B(sa); B(rk); B(kv0); B(kv1); B(hist); B(bwt_misc); B(bwt_ghist); B(lcol);
    B(mtf0); B(mtfv16); B(mtf_freq); B(bits); B(pool); B(pool_woff); B(streams); B(stream_meta);
    B(io_a); B(io_b); B(io_c); B(io_d); B(io_e);
#undef B
    *n = k;
    return list;
}

// device bytes needed per bzip2 block inside one batch of stages 3b..3d
static size_t batch_bytes_per_block()
{
    return (size_t)BLK_STRIDE * (4 + 4 + 8 + 8 + 1 + 1 + 2) + (size_t)1024 * 220 * 4 + (size_t)BITS_WORDS * 4 + 258 * 4 + 4096;
}

static uint64_t pick_batch(uint64_t n_blocks)
{
    const char *env = getenv("S3G_BATCH");
    if (env && atoll(env) > 0) return std::min<uint64_t>(n_blocks, (uint64_t)atoll(env));
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); fr = (size_t)8 << 30; }
    uint64_t fit = (uint64_t)((double)fr * 0.85 / (double)batch_bytes_per_block());
    if (fit < 1) fit = 1;
    return std::min<uint64_t>(n_blocks, std::min<uint64_t>(fit, 1024));
}

__global__ void k_soff_from_chroms(const s3g_chrom *chroms, uint64_t n_chroms, uint64_t tf_len, uint64_t *soff)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c < n_chroms) soff[c] = chroms[c].tf_off;
    if (c == n_chroms) soff[c] = tf_len;
}

// stages 3a..3e over `n_streams` streams laid out back to back in d_in
static double host_ms()
{
    timespec ts;
    clock_gettime(CLOCK_MONOTONIC, &ts);
    return ts.tv_sec * 1e3 + ts.tv_nsec * 1e-6;
}

static int compress_streams(Ctx *ctx, const uint8_t *d_in, uint64_t n, const uint64_t *d_soff, uint64_t n_streams, int level,
                            uint64_t *n_blocks_out, uint64_t *total_bytes)
{
    const bool timing = getenv("S3G_TIMING") != nullptr;
    double t0 = timing ? host_ms() : 0, t1 = 0, t2 = 0, t3 = 0, t4 = 0, t5 = 0;
    CutResult cut;
    S3G_TRY(run_rle_cut(ctx, d_in, n, d_soff, n_streams, level, &cut));
    if (timing) t1 = host_ms();
    uint64_t nb = cut.n_blocks;
    *n_blocks_out = nb;
    ctx->pool_words = 0;
    S3G_TRY(ctx->pool_woff.ensure((nb + 1) * 8));
    if (nb) {
        // the batch buffers are grow-only: size them once for the batch this input needs
        uint64_t held = 0;
        DevBuf *batch_bufs[] = {&ctx->sa, &ctx->rk, &ctx->kv0, &ctx->kv1, &ctx->hist, &ctx->lcol, &ctx->mtf0, &ctx->mtfv16, &ctx->bits};
        for (DevBuf *b : batch_bufs) held += b->cap;
        // steady state: the buffers already hold this many blocks; cudaMemGetInfo is only asked when they must grow
        // (it takes anywhere from 0.1 to 20 ms on a shared host)
        uint64_t have = held / batch_bytes_per_block();
        uint64_t batch = have >= nb && !getenv("S3G_BATCH") ? nb : pick_batch(nb);
        if (have > batch) batch = std::min<uint64_t>(nb, have);
        for (uint64_t b0 = 0; b0 < nb; b0 += batch) {
            uint64_t cnt = std::min<uint64_t>(batch, nb - b0);
            if (timing) t2 = host_ms();
            S3G_TRY(run_bwt(ctx, b0, cnt));
            if (timing) t3 = host_ms();
            S3G_TRY(run_mtf(ctx, b0, cnt));
            S3G_TRY(run_huff(ctx, b0, cnt, 1, nullptr, nullptr));
            S3G_TRY(run_pool_append(ctx, b0, cnt));
            if (timing) t4 = host_ms();
        }
    }
    S3G_TRY(run_assemble(ctx, nb, n_streams, level, total_bytes));
    if (timing) {
        t5 = host_ms();
        fprintf(stderr, "[s3g timing] rle+cut %.2f  batch setup %.2f  bwt %.2f  mtf+huff+pool (enqueue) %.2f  assemble %.2f  total %.2f ms (host clock)\n",
                t1 - t0, t2 - t1, t3 - t2, t4 - t3, t5 - t4, t5 - t0);
    }
    return S3G_OK;
}

// ---- archive header (ARCHIVE_FORMAT.md) -----------------------------------------
static void json_string(std::string &o, const uint8_t *s, size_t n)
{
    // same escapes as jansson's dump_string (jansson-2.9 src/dump.c:70-160) without JSON_ESCAPE_SLASH / JSON_ENSURE_ASCII
    o.push_back('"');
    for (size_t i = 0; i < n; i++) {
        uint8_t c = s[i];
        switch (c) {
            case '\\': o += "\\\\"; break;
            case '"': o += "\\\""; break;
            case '\b': o += "\\b"; break;
            case '\f': o += "\\f"; break;
            case '\n': o += "\\n"; break;
            case '\r': o += "\\r"; break;
            case '\t': o += "\\t"; break;
            default:
                if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04X", (unsigned)c); o += b; }
                else o.push_back((char)c);
        }
    }
    o.push_back('"');
}

static std::string build_header(const uint8_t *names, const std::vector<uint64_t> &name_off, const std::vector<s3g_chrom> &ch,
                                int level, const char *note)
{
    std::string o;
    o.reserve(256 + ch.size() * 160);
    o += "{\"archive\":{\"type\":\"starch\",\"version\":{\"major\":3,\"minor\":0,\"revision\":0},\"creator\":\"starch3_b200\","
         "\"compression\":\"bzip2\",\"blockSize100k\":";
    o += std::to_string(level);
    o += ",\"note\":";
    const char *nt = note ? note : "";
    json_string(o, reinterpret_cast<const uint8_t *>(nt), strlen(nt));
    o += "},\"streams\":[";
    for (size_t i = 0; i < ch.size(); i++) {
        const s3g_chrom &c = ch[i];
        if (i) o.push_back(',');
        o += "{\"chromosome\":";
        json_string(o, names + name_off[i], c.name_len);
        o += ",\"offset\":" + std::to_string(c.bz_off);
        o += ",\"size\":" + std::to_string(c.bz_len);
        o += ",\"lines\":" + std::to_string(c.line_count);
        o += ",\"blocks\":" + std::to_string(c.n_blocks);
        o += ",\"transformedBytes\":" + std::to_string(c.tf_len);
        o += ",\"nonUniqueBases\":" + std::to_string(c.bases_nonunique);
        o += ",\"uniqueBases\":" + std::to_string(c.bases_unique);
        o.push_back('}');
    }
    o += "]}";
    return o;
}

__global__ void k_gather_names(const uint8_t *bed, const s3g_chrom *chroms, uint64_t n_chroms, const uint64_t *dst_off, uint8_t *dst)
{
    uint64_t c = blockIdx.x;
    if (c >= n_chroms) return;
    for (uint32_t i = threadIdx.x; i < chroms[c].name_len; i += blockDim.x) dst[dst_off[c] + i] = bed[chroms[c].name_off + i];
}

static int compress_bed_impl(Ctx *ctx, const uint8_t *d_bed, uint64_t n, int level, const char *note, int want_archive,
                             s3g_result *res)
{
    memset(res, 0, sizeof *res);
    if (level < 1 || level > 9) { set_error("block_size_100k must be 1..9"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    S3G_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    TfResult tr;
    double tt0 = getenv("S3G_TIMING") ? host_ms() : 0;
    S3G_TRY(run_transform(ctx, d_bed, n, &tr, false));
    if (getenv("S3G_TIMING")) fprintf(stderr, "[s3g timing] transform %.2f ms (host clock)\n", host_ms() - tt0);
    res->n_lines = tr.n_lines; res->n_chroms = tr.n_chroms; res->tf_bytes = tr.tf_len; res->dropped_tail_bytes = tr.dropped;
    uint64_t total_bytes = 0, n_blocks = 0;
    if (tr.n_chroms) {
        S3G_TRY(ctx->soff.ensure((tr.n_chroms + 2) * 8));
        S3G_LAUNCH(ctx, k_soff_from_chroms, (unsigned)((tr.n_chroms + 1 + 127) / 128), 128, 0, ctx->chroms.as<s3g_chrom>(),
                   tr.n_chroms, tr.tf_len, ctx->soff.as<uint64_t>());
        S3G_TRY(compress_streams(ctx, ctx->tf.as<uint8_t>(), tr.tf_len, ctx->soff.as<uint64_t>(), tr.n_chroms, level, &n_blocks,
                                 &total_bytes));
    }
    S3G_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    res->n_blocks = n_blocks;
    res->d_streams = ctx->streams.p;
    res->streams_size = total_bytes;
    ctx->last_streams_size = total_bytes;
    // ---- metadata back to the host ----
    ctx->h_chroms.resize(tr.n_chroms);
    std::vector<StreamMeta> meta(tr.n_chroms);
    if (tr.n_chroms) {
        S3G_CUDA(cudaMemcpyAsync(ctx->h_chroms.data(), ctx->chroms.p, tr.n_chroms * sizeof(s3g_chrom), cudaMemcpyDeviceToHost, ctx->stream));
        S3G_CUDA(cudaMemcpyAsync(meta.data(), ctx->stream_meta.p, tr.n_chroms * sizeof(StreamMeta), cudaMemcpyDeviceToHost, ctx->stream));
    }
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    float ms = 0;
    S3G_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    res->device_ms = ms;
    for (uint64_t c = 0; c < tr.n_chroms; c++) {
        ctx->h_chroms[c].bz_off = meta[c].byte_off; ctx->h_chroms[c].bz_len = meta[c].byte_len;
        ctx->h_chroms[c].n_blocks = (uint32_t)meta[c].n_blocks;
    }
    res->chroms = (s3g_chrom *)malloc(std::max<size_t>(1, tr.n_chroms) * sizeof(s3g_chrom));
    if (!res->chroms) { set_error("out of host memory"); return S3G_E_NOMEM; }
    if (tr.n_chroms) memcpy(res->chroms, ctx->h_chroms.data(), tr.n_chroms * sizeof(s3g_chrom));
    if (!want_archive) return S3G_OK;
    // chromosome names (a few bytes each) come back through one gather
    std::vector<uint64_t> name_off(tr.n_chroms + 1, 0);
    for (uint64_t c = 0; c < tr.n_chroms; c++) name_off[c + 1] = name_off[c] + ctx->h_chroms[c].name_len;
    std::vector<uint8_t> names(name_off[tr.n_chroms] + 1);
    if (tr.n_chroms) {
        S3G_TRY(ctx->io_a.ensure((tr.n_chroms + 1) * 8));
        S3G_TRY(ctx->io_b.ensure(name_off[tr.n_chroms] + 16));
        S3G_CUDA(cudaMemcpyAsync(ctx->io_a.p, name_off.data(), (tr.n_chroms + 1) * 8, cudaMemcpyHostToDevice, ctx->stream));
        S3G_LAUNCH(ctx, k_gather_names, (unsigned)tr.n_chroms, 32, 0, d_bed, ctx->chroms.as<s3g_chrom>(), tr.n_chroms,
                   ctx->io_a.as<uint64_t>(), ctx->io_b.as<uint8_t>());
        S3G_CUDA(cudaMemcpyAsync(names.data(), ctx->io_b.p, name_off[tr.n_chroms], cudaMemcpyDeviceToHost, ctx->stream));
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    std::string hdr = build_header(names.data(), name_off, ctx->h_chroms, level, note);
    uint64_t streams_off = 4 + hdr.size() + 1;
    res->archive_size = streams_off + total_bytes;
    // the archive is assembled in context-owned pinned memory: the device-to-host copy of the
    // streams lands directly in its final place
    if (res->archive_size > ctx->h_archive_cap) {
        if (ctx->h_archive) cudaFreeHost(ctx->h_archive);
        ctx->h_archive = nullptr; ctx->h_archive_cap = 0;
        size_t want = res->archive_size + res->archive_size / 8 + 4096;
        if (cudaMallocHost(&ctx->h_archive, want) != cudaSuccess) { cudaGetLastError(); set_error("out of pinned host memory"); return S3G_E_NOMEM; }
        ctx->h_archive_cap = want;
    }
    res->archive = ctx->h_archive;
    static const uint8_t magic[4] = {0xca, 0x5c, 0xad, 0x1a};      // hpp:907-910
    memcpy(res->archive, magic, 4);
    memcpy(res->archive + 4, hdr.data(), hdr.size());
    res->archive[4 + hdr.size()] = '\n';
    res->streams_off = streams_off;
    if (total_bytes) {
        S3G_CUDA(cudaMemcpyAsync(res->archive + streams_off, ctx->streams.p, total_bytes, cudaMemcpyDeviceToHost, ctx->stream));
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return S3G_OK;
}

// copy a host buffer into a context staging buffer (16-byte padded)
static int stage_in(Ctx *ctx, DevBuf &buf, const void *src, uint64_t n)
{
    S3G_TRY(buf.ensure(n + 64));
    if (n) S3G_CUDA(cudaMemcpyAsync(buf.p, src, n, cudaMemcpyHostToDevice, ctx->stream));
    S3G_CUDA(cudaMemsetAsync((uint8_t *)buf.p + n, 0, 64, ctx->stream));
    return S3G_OK;
}

}  // namespace s3g

using namespace s3g;

extern "C" {

const char *s3g_last_error(void) { return s3g::g_err; }

int s3g_init(int device, s3g_ctx **out)
{
    if (!out) { set_error("null out pointer"); return S3G_E_PARAM; }
    *out = nullptr;
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        set_error("no CUDA device available (%s); this library has no CPU fallback", e == cudaSuccess ? "device count 0" : cudaGetErrorString(e));
        return S3G_E_CUDA;
    }
    if (device < 0 || device >= ndev) { set_error("device %d out of range (0..%d)", device, ndev - 1); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(device));
    s3g_ctx *c = new (std::nothrow) s3g_ctx();
    if (!c) { set_error("out of host memory"); return S3G_E_NOMEM; }
    c->device = device;
    S3G_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
    c->stream = c->own_stream;
    S3G_CUDA(cudaEventCreate(&c->ev0));
    S3G_CUDA(cudaEventCreate(&c->ev1));
    S3G_CUDA(cudaMallocHost(&c->h_scalars, 64 * 8));
    *out = c;
    return S3G_OK;
}

void s3g_destroy(s3g_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaStreamSynchronize(ctx->stream);
    size_t k; DevBuf *const *bl = all_bufs(ctx, &k);
    for (size_t i = 0; i < k; i++) bl[i]->release();
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    if (ctx->h_archive) cudaFreeHost(ctx->h_archive);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int s3g_set_stream(s3g_ctx *ctx, void *cuda_stream)
{
    if (!ctx) { set_error("null context"); return S3G_E_PARAM; }
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return S3G_OK;
}

uint64_t s3g_launch_count(const s3g_ctx *ctx) { return ctx ? ctx->launches : 0; }

int s3g_profile(s3g_ctx *ctx, int enable)
{
    if (!ctx) { set_error("null context"); return S3G_E_PARAM; }
    ctx->prof = enable != 0;
    if (!enable) { ctx->prof_recs.clear(); ctx->prof_used = 0; }
    return S3G_OK;
}

int s3g_profile_filter(s3g_ctx *ctx, const char *kernel_name)
{
    if (!ctx) { set_error("null context"); return S3G_E_PARAM; }
    ctx->prof_filter = kernel_name ? kernel_name : "";
    return S3G_OK;
}

int s3g_profile_report(s3g_ctx *ctx, char *buf, uint64_t cap)
{
    if (!ctx || !buf || cap == 0) { set_error("null argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    std::vector<std::string> names; std::vector<double> ms, by; std::vector<uint64_t> cnt;
    for (const Ctx::ProfRec &r : ctx->prof_recs) {
        float t = 0;
        if (cudaEventElapsedTime(&t, r.e0, r.e1) != cudaSuccess) { cudaGetLastError(); continue; }
        size_t k = 0;
        for (; k < names.size(); k++) if (names[k] == r.name) break;
        if (k == names.size()) { names.push_back(r.name); ms.push_back(0); by.push_back(0); cnt.push_back(0); }
        ms[k] += t; by[k] += r.bytes; cnt[k]++;
    }
    std::string o;
    for (size_t k = 0; k < names.size(); k++) {
        char line[256];
        snprintf(line, sizeof line, "%s\t%llu\t%.6f\t%.0f\n", names[k].c_str(), (unsigned long long)cnt[k], ms[k], by[k]);
        o += line;
    }
    ctx->prof_recs.clear(); ctx->prof_used = 0;
    if (o.size() + 1 > cap) { set_error("profile buffer too small"); return S3G_E_CAPACITY; }
    memcpy(buf, o.c_str(), o.size() + 1);
    return S3G_OK;
}

int s3g_compress_bed_device(s3g_ctx *ctx, const void *d_bed, uint64_t n, int level, const char *note, int want_archive, s3g_result *res)
{
    if (!ctx || !res || (!d_bed && n)) { set_error("null argument"); return S3G_E_PARAM; }
    if (((uintptr_t)d_bed & 15) != 0) { set_error("device input must be 16-byte aligned"); return S3G_E_PARAM; }
    return compress_bed_impl(ctx, (const uint8_t *)d_bed, n, level, note, want_archive, res);
}

int s3g_compress_bed(s3g_ctx *ctx, const uint8_t *bed, uint64_t n, int level, const char *note, s3g_result *res)
{
    if (!ctx || !res || (!bed && n)) { set_error("null argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    S3G_TRY(stage_in(ctx, ctx->bed, bed, n));
    return compress_bed_impl(ctx, ctx->bed.as<uint8_t>(), n, level, note, 1, res);
}

int s3g_read_streams(s3g_ctx *ctx, uint8_t *dst, uint64_t cap, uint64_t *n)
{
    if (!ctx || !dst || !n) { set_error("null argument"); return S3G_E_PARAM; }
    *n = ctx->last_streams_size;
    if (*n > cap) { set_error("cap too small: need %llu", (unsigned long long)*n); return S3G_E_CAPACITY; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    if (*n) S3G_CUDA(cudaMemcpyAsync(dst, ctx->streams.p, *n, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S3G_OK;
}

void s3g_result_free(s3g_result *res)
{
    if (!res) return;
    free(res->chroms);                  // the archive buffer belongs to the context
    memset(res, 0, sizeof *res);
}

// ---- stage entry points ------------------------------------------------------------

int s3g_tokenize(s3g_ctx *ctx, const uint8_t *bed, uint64_t n, uint64_t cap_lines, uint64_t *n_lines, uint64_t *line_start,
                 int64_t *start, int64_t *stop, uint32_t *rem_off, uint8_t *chrom_change)
{
    if (!ctx || !n_lines) { set_error("null argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    S3G_TRY(stage_in(ctx, ctx->bed, bed, n));
    TfResult tr;
    S3G_TRY(run_transform(ctx, ctx->bed.as<uint8_t>(), n, &tr, true));
    *n_lines = tr.n_lines;
    if (tr.n_lines > cap_lines) { set_error("cap_lines too small: need %llu", (unsigned long long)tr.n_lines); return S3G_E_CAPACITY; }
    uint64_t m = tr.n_lines;
    if (line_start) S3G_CUDA(cudaMemcpyAsync(line_start, ctx->line_start.p, (m + 1) * 8, cudaMemcpyDeviceToHost, ctx->stream));
    if (m) {
        if (start) S3G_CUDA(cudaMemcpyAsync(start, ctx->start.p, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (stop) S3G_CUDA(cudaMemcpyAsync(stop, ctx->stop.p, m * 8, cudaMemcpyDeviceToHost, ctx->stream));
        if (rem_off) S3G_CUDA(cudaMemcpyAsync(rem_off, ctx->rem_off.p, m * 4, cudaMemcpyDeviceToHost, ctx->stream));
        if (chrom_change) S3G_CUDA(cudaMemcpyAsync(chrom_change, ctx->flags.p, m, cudaMemcpyDeviceToHost, ctx->stream));
    }
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    if (chrom_change) for (uint64_t i = 0; i < m; i++) chrom_change[i] &= 1;
    return S3G_OK;
}

int s3g_transform(s3g_ctx *ctx, const uint8_t *bed, uint64_t n, uint8_t *tf, uint64_t tf_cap, uint64_t *tf_len, s3g_chrom *chroms,
                  uint64_t chrom_cap, uint64_t *n_chroms, uint64_t *dropped)
{
    if (!ctx || !tf_len || !n_chroms) { set_error("null argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    S3G_TRY(stage_in(ctx, ctx->bed, bed, n));
    TfResult tr;
    S3G_TRY(run_transform(ctx, ctx->bed.as<uint8_t>(), n, &tr, false));
    *tf_len = tr.tf_len; *n_chroms = tr.n_chroms;
    if (dropped) *dropped = tr.dropped;
    if (tr.tf_len > tf_cap || tr.n_chroms > chrom_cap) { set_error("output capacity too small"); return S3G_E_CAPACITY; }
    if (tr.tf_len && tf) S3G_CUDA(cudaMemcpyAsync(tf, ctx->tf.p, tr.tf_len, cudaMemcpyDeviceToHost, ctx->stream));
    if (tr.n_chroms && chroms) S3G_CUDA(cudaMemcpyAsync(chroms, ctx->chroms.p, tr.n_chroms * sizeof(s3g_chrom), cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S3G_OK;
}

static int single_stream_soff(Ctx *ctx, uint64_t n)
{
    S3G_TRY(ctx->soff.ensure(2 * 8));
    uint64_t h[2] = {0, n};
    S3G_CUDA(cudaMemcpyAsync(ctx->soff.p, h, 16, cudaMemcpyHostToDevice, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S3G_OK;
}

int s3g_rle1(s3g_ctx *ctx, const uint8_t *in, uint64_t n, int level, s3g_blockdesc *desc, uint64_t desc_cap, uint64_t *n_blocks,
             uint8_t *rle_out, uint64_t rle_cap)
{
    if (!ctx || !n_blocks) { set_error("null argument"); return S3G_E_PARAM; }
    if (level < 1 || level > 9) { set_error("block_size_100k must be 1..9"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    S3G_TRY(stage_in(ctx, ctx->io_a, in, n));
    S3G_TRY(single_stream_soff(ctx, n));
    CutResult cut;
    S3G_TRY(run_rle_cut(ctx, ctx->io_a.as<uint8_t>(), n, ctx->soff.as<uint64_t>(), 1, level, &cut));
    *n_blocks = cut.n_blocks;
    if (cut.n_blocks > desc_cap) { set_error("desc_cap too small"); return S3G_E_CAPACITY; }
    ctx->h_blocks.resize(cut.n_blocks);
    std::vector<uint8_t> use(cut.n_blocks * 256);
    if (cut.n_blocks) {
        S3G_CUDA(cudaMemcpyAsync(ctx->h_blocks.data(), ctx->blocks.p, cut.n_blocks * sizeof(BlockInfo), cudaMemcpyDeviceToHost, ctx->stream));
        S3G_CUDA(cudaMemcpyAsync(use.data(), ctx->in_use.p, cut.n_blocks * 256, cudaMemcpyDeviceToHost, ctx->stream));
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    uint64_t off = 0;
    for (uint64_t b = 0; b < cut.n_blocks; b++) {
        const BlockInfo &B = ctx->h_blocks[b];
        if (desc) {
            desc[b].in_start = B.in_start; desc[b].in_end = B.in_end; desc[b].nblock = B.nblock; desc[b].crc = B.crc;
            memcpy(desc[b].in_use, &use[b * 256], 256);
        }
        if (rle_out) {
            if (off + B.nblock > rle_cap) { set_error("rle_cap too small"); return S3G_E_CAPACITY; }
            S3G_CUDA(cudaMemcpyAsync(rle_out + off, ctx->blk_bytes.as<uint8_t>() + b * (uint64_t)BLK_STRIDE, B.nblock, cudaMemcpyDeviceToHost, ctx->stream));
        }
        off += B.nblock;
    }
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S3G_OK;
}

// stage helper: load caller-provided blocks into the batched device layout
__global__ void k_in_use_from_bytes(const uint8_t *blk, const BlockInfo *blocks, uint8_t *in_use)
{
    uint64_t b = blockIdx.y;
    uint32_t n = blocks[b].nblock;
    const uint8_t *p = blk + b * (uint64_t)BLK_STRIDE;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) in_use[b * 256 + p[i]] = 1;
}
__global__ void k_block_maps_api(const uint8_t *in_use, BlockInfo *blocks, uint8_t *seq_map)
{
    __shared__ uint32_t s[33];
    uint64_t b = blockIdx.x;
    uint32_t u = in_use[b * 256 + threadIdx.x] ? 1u : 0u;
    uint32_t tot;
    uint32_t ex = block_excl_sum<uint32_t>(u, s, &tot);
    seq_map[b * 256 + threadIdx.x] = (uint8_t)ex;
    if (threadIdx.x == 0) blocks[b].n_in_use = tot;
}

__global__ void k_lcol_from_ptr(const uint8_t *blk, const uint32_t *ptr, const uint8_t *seq, uint32_t n, uint8_t *L)
{
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
        uint32_t s = ptr[i];
        L[i] = seq[blk[s ? s - 1 : n - 1]];       // bz/compress.c:166-167
    }
}

static int load_blocks(Ctx *ctx, const uint8_t *blocks, const uint64_t *off, uint64_t nb, const uint8_t *in_use /*nullable, 256 per block*/)
{
    ctx->h_blocks.assign(nb, BlockInfo());
    S3G_TRY(ctx->blocks.ensure(nb * sizeof(BlockInfo)));
    S3G_TRY(ctx->blk_bytes.ensure(nb * (uint64_t)BLK_STRIDE));
    S3G_TRY(ctx->in_use.ensure(nb * 256));
    S3G_TRY(ctx->seq_map.ensure(nb * 256));
    for (uint64_t b = 0; b < nb; b++) {
        uint64_t len = off[b + 1] - off[b];
        if (len == 0 || len > 900000) { set_error("block %llu has invalid length %llu", (unsigned long long)b, (unsigned long long)len); return S3G_E_PARAM; }
        BlockInfo &B = ctx->h_blocks[b];
        memset(&B, 0, sizeof B);
        B.nblock = (uint32_t)len; B.orig_ptr = -1;
        S3G_CUDA(cudaMemcpyAsync(ctx->blk_bytes.as<uint8_t>() + b * (uint64_t)BLK_STRIDE, blocks + off[b], len, cudaMemcpyHostToDevice, ctx->stream));
    }
    S3G_CUDA(cudaMemcpyAsync(ctx->blocks.p, ctx->h_blocks.data(), nb * sizeof(BlockInfo), cudaMemcpyHostToDevice, ctx->stream));
    if (in_use) S3G_CUDA(cudaMemcpyAsync(ctx->in_use.p, in_use, nb * 256, cudaMemcpyHostToDevice, ctx->stream));
    else {
        S3G_CUDA(cudaMemsetAsync(ctx->in_use.p, 0, nb * 256, ctx->stream));
        dim3 g(64, (unsigned)nb);
        S3G_LAUNCH(ctx, k_in_use_from_bytes, g, 256, 0, ctx->blk_bytes.as<uint8_t>(), ctx->blocks.as<BlockInfo>(), ctx->in_use.as<uint8_t>());
    }
    S3G_LAUNCH(ctx, k_block_maps_api, (unsigned)nb, 256, 0, ctx->in_use.as<uint8_t>(), ctx->blocks.as<BlockInfo>(), ctx->seq_map.as<uint8_t>());
    // host mirror with the alphabet sizes (the later stages size their launches from it)
    S3G_CUDA(cudaMemcpyAsync(ctx->h_blocks.data(), ctx->blocks.p, nb * sizeof(BlockInfo), cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    return check_launch("load blocks");
}

int s3g_bwt(s3g_ctx *ctx, const uint8_t *blocks, const uint64_t *off, uint64_t nb, uint32_t *ptr_out, int32_t *orig_ptr)
{
    if (!ctx || !blocks || !off) { set_error("null argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    if (nb == 0) return S3G_OK;
    S3G_TRY(load_blocks(ctx, blocks, off, nb, nullptr));
    uint64_t batch = pick_batch(nb);
    for (uint64_t b0 = 0; b0 < nb; b0 += batch) {
        uint64_t cnt = std::min<uint64_t>(batch, nb - b0);
        S3G_TRY(run_bwt(ctx, b0, cnt));
        if (ptr_out)
            for (uint64_t b = 0; b < cnt; b++)
                S3G_CUDA(cudaMemcpyAsync(ptr_out + off[b0 + b], ctx->sa.as<uint32_t>() + b * (uint64_t)BLK_STRIDE, (off[b0 + b + 1] - off[b0 + b]) * 4,
                                         cudaMemcpyDeviceToHost, ctx->stream));
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    S3G_CUDA(cudaMemcpyAsync(ctx->h_blocks.data(), ctx->blocks.p, nb * sizeof(BlockInfo), cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    if (orig_ptr) for (uint64_t b = 0; b < nb; b++) orig_ptr[b] = ctx->h_blocks[b].orig_ptr;
    return S3G_OK;
}

int s3g_mtf(s3g_ctx *ctx, const uint8_t *block, uint32_t n, const uint32_t *ptr, const uint8_t *in_use, uint16_t *mtfv, uint32_t *n_mtf, int32_t *freq)
{
    if (!ctx || !block || !ptr || !in_use) { set_error("null argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    uint64_t off[2] = {0, n};
    S3G_TRY(load_blocks(ctx, block, off, 1, in_use));
    // L column from the caller's sorted order
    S3G_TRY(ctx->lcol.ensure(BLK_STRIDE));
    S3G_TRY(ctx->sa.ensure((size_t)BLK_STRIDE * 4));
    S3G_CUDA(cudaMemcpyAsync(ctx->sa.p, ptr, (size_t)n * 4, cudaMemcpyHostToDevice, ctx->stream));
    S3G_LAUNCH(ctx, k_lcol_from_ptr, 256, 256, 0, ctx->blk_bytes.as<uint8_t>(), ctx->sa.as<uint32_t>(), ctx->seq_map.as<uint8_t>(), n,
               ctx->lcol.as<uint8_t>());
    S3G_TRY(run_mtf(ctx, 0, 1));
    S3G_CUDA(cudaMemcpyAsync(ctx->h_blocks.data(), ctx->blocks.p, sizeof(BlockInfo), cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    uint32_t m = ctx->h_blocks[0].n_mtf;
    if (n_mtf) *n_mtf = m;
    if (mtfv) S3G_CUDA(cudaMemcpyAsync(mtfv, ctx->mtfv16.p, (size_t)m * 2, cudaMemcpyDeviceToHost, ctx->stream));
    if (freq) S3G_CUDA(cudaMemcpyAsync(freq, ctx->mtf_freq.p, 258 * 4, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S3G_OK;
}

int s3g_huff(s3g_ctx *ctx, const uint16_t *mtfv, uint32_t n_mtf, const int32_t *freq, const uint8_t *in_use, int32_t *n_groups,
             int32_t *n_selectors, uint8_t *selector, uint8_t *len, uint8_t *bits, uint64_t bits_cap, uint64_t *n_bits)
{
    if (!ctx || !mtfv || !freq || !in_use) { set_error("null argument"); return S3G_E_PARAM; }
    if (n_mtf == 0 || n_mtf > 900001) { set_error("n_mtf out of range"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    ctx->h_blocks.assign(1, BlockInfo());
    BlockInfo &B = ctx->h_blocks[0];
    memset(&B, 0, sizeof B);
    B.n_mtf = n_mtf; B.nblock = n_mtf;
    for (int i = 0; i < 256; i++) B.n_in_use += in_use[i] ? 1 : 0;
    S3G_TRY(ctx->blocks.ensure(sizeof(BlockInfo)));
    S3G_TRY(ctx->in_use.ensure(256));
    S3G_TRY(ctx->mtfv16.ensure((size_t)BLK_STRIDE * 2));
    S3G_TRY(ctx->mtf_freq.ensure(258 * 4));
    S3G_TRY(ctx->io_b.ensure(18004 + 6 * 258 + 64));
    S3G_CUDA(cudaMemcpyAsync(ctx->blocks.p, &B, sizeof B, cudaMemcpyHostToDevice, ctx->stream));
    S3G_CUDA(cudaMemcpyAsync(ctx->in_use.p, in_use, 256, cudaMemcpyHostToDevice, ctx->stream));
    S3G_CUDA(cudaMemcpyAsync(ctx->mtfv16.p, mtfv, (size_t)n_mtf * 2, cudaMemcpyHostToDevice, ctx->stream));
    S3G_CUDA(cudaMemcpyAsync(ctx->mtf_freq.p, freq, 258 * 4, cudaMemcpyHostToDevice, ctx->stream));
    uint8_t *d_sel = ctx->io_b.as<uint8_t>(), *d_len = d_sel + 18004;
    S3G_TRY(run_huff(ctx, 0, 1, 0, d_sel, d_len));
    S3G_CUDA(cudaMemcpyAsync(&B, ctx->blocks.p, sizeof B, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    int nsel = (int)((n_mtf + 49) / 50);
    if (n_groups) *n_groups = n_mtf < 200 ? 2 : n_mtf < 600 ? 3 : n_mtf < 1200 ? 4 : n_mtf < 2400 ? 5 : 6;
    if (n_selectors) *n_selectors = nsel;
    if (n_bits) *n_bits = B.n_bits;
    uint64_t nbytes = (B.n_bits + 7) / 8;
    if (bits && nbytes > bits_cap) { set_error("bits_cap too small"); return S3G_E_CAPACITY; }
    if (selector) S3G_CUDA(cudaMemcpyAsync(selector, d_sel, (size_t)nsel, cudaMemcpyDeviceToHost, ctx->stream));
    if (len) S3G_CUDA(cudaMemcpyAsync(len, d_len, 6 * 258, cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<uint32_t> w((B.n_bits + 31) / 32);
    if (bits && !w.empty()) S3G_CUDA(cudaMemcpyAsync(w.data(), ctx->bits.p, w.size() * 4, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    if (bits) for (uint64_t i = 0; i < nbytes; i++) bits[i] = (uint8_t)(w[i >> 2] >> (24 - 8 * (i & 3)));
    return S3G_OK;
}

int s3g_bz_compress(s3g_ctx *ctx, const uint8_t *in, uint64_t n, int level, uint8_t *out, uint64_t out_cap, uint64_t *out_len)
{
    if (!ctx || !out_len) { set_error("null argument"); return S3G_E_PARAM; }
    if (level < 1 || level > 9) { set_error("block_size_100k must be 1..9"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    S3G_TRY(stage_in(ctx, ctx->io_a, in, n));
    S3G_TRY(single_stream_soff(ctx, n));
    uint64_t nblocks = 0, total = 0;
    S3G_TRY(compress_streams(ctx, ctx->io_a.as<uint8_t>(), n, ctx->soff.as<uint64_t>(), 1, level, &nblocks, &total));
    *out_len = total;
    if (total > out_cap) { set_error("out_cap too small: need %llu", (unsigned long long)total); return S3G_E_CAPACITY; }
    if (out && total) S3G_CUDA(cudaMemcpyAsync(out, ctx->streams.p, total, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S3G_OK;
}

}  // extern "C"
