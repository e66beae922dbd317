// api.cu -- the C ABI (include/starch3_b200.h) and the host-side orchestration of the
// GPU pipeline.  Host code is plain C++; nothing here computes on the CPU except the
// few-hundred-byte archive header (ARCHIVE_FORMAT.md).
#include <cstdarg>
#include <ctime>
#include <condition_variable>
#include <mutex>
#include <thread>
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

void stage_mark(Ctx *ctx, int stage)
{
    size_t k = ctx->marks.size();
    if (k >= ctx->mark_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) { cudaGetLastError(); return; }
        ctx->mark_pool.push_back(e);
    }
    cudaEventRecord(ctx->mark_pool[k], ctx->stream);
    ctx->marks.emplace_back(stage, ctx->mark_pool[k]);
}

// after the call's final synchronise: per-stage sums of the intervals between consecutive marks
void stage_collect(Ctx *ctx, double *stage_ms)
{
    for (int i = 0; i < 8; i++) stage_ms[i] = 0;
    for (size_t k = 0; k + 1 < ctx->marks.size(); k++) {
        float t = 0;
        if (cudaEventElapsedTime(&t, ctx->marks[k].second, ctx->marks[k + 1].second) != cudaSuccess) { cudaGetLastError(); continue; }
        int st = ctx->marks[k].first;
        if (st >= 0 && st < 8) stage_ms[st] += t;
    }
    ctx->marks.clear();
}

__global__ void k_fetch_host(const uint64_t *__restrict__ src, uint64_t *__restrict__ dst, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int upload_small(Ctx *ctx, int slot, void *d_dst, const void *h_src, size_t bytes)
{
    if (bytes == 0) return S3G_OK;
    if (bytes & 7) { set_error("upload_small: size must be a multiple of 8"); return S3G_E_PARAM; }
    if (bytes > ctx->h_small_cap[slot]) {
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));             // nobody reads the old area any more
        if (ctx->h_small[slot]) cudaFreeHost(ctx->h_small[slot]);
        ctx->h_small[slot] = nullptr; ctx->h_small_cap[slot] = 0;
        const size_t want = std::max<size_t>(bytes * 2, 1 << 16);
        if (cudaMallocHost(&ctx->h_small[slot], want) != cudaSuccess) { cudaGetLastError(); set_error("out of pinned host memory"); return S3G_E_NOMEM; }
        ctx->h_small_cap[slot] = want;
    }
    memcpy(ctx->h_small[slot], h_src, bytes);
    const uint64_t n = bytes / 8;
    S3G_LAUNCH(ctx, k_fetch_host, (unsigned)std::min<uint64_t>(64, (n + 255) / 256), 256, 0, ctx->h_small[slot], static_cast<uint64_t *>(d_dst), n);
    return S3G_OK;
}

static DevBuf *const *all_bufs(Ctx *c, size_t *n)
{
    static thread_local DevBuf *list[64];
    size_t k = 0;
#define B(x) list[k++] = &c->x
    B(bed); B(tile_cnt); B(line_start); B(start); B(stop); B(rem_off); B(flags); B(chrom_first);
    B(scan_a); B(scan_b); B(scan_c); B(scalars); B(tf); B(chroms); B(stat_b); B(soff);
    B(rle_carry); B(rle_ebase); B(blocks); B(blk_prov); B(blk_bytes); B(in_use); B(seq_map); B(stream_tab);
    B(sa); B(rk); B(kv0); B(kv1); B(hist); B(bwt_misc); B(bwt_ghist); B(lcol);
    B(mtf0); B(mtfv16); B(mtf_freq); B(ztiles); B(bits); B(pool); B(pool_woff); B(streams); B(stream_meta);
    B(io_a); B(io_b); B(io_c); B(io_d); B(io_e); B(chain_tf[0]); B(chain_tf[1]); B(chain_out); B(front_incl); B(front_flag);
#undef B
    *n = k;
    return list;
}

// device bytes needed per bzip2 block inside one batch of stages 3b..3d
static size_t batch_bytes_per_block()
{
    return (size_t)BLK_STRIDE * (4 + 4 + 8 + 8 + 1 + 1 + 2) + (size_t)1024 * 220 * 4 + (size_t)BITS_WORDS * 4 + 258 * 4 + 4096;
}

static uint64_t pick_batch(uint64_t n_blocks, double mem_frac = 0.85)
{
    const char *env = getenv("S3G_BATCH");
    if (env && atoll(env) > 0) return std::min<uint64_t>(n_blocks, (uint64_t)atoll(env));
    size_t fr = 0, tot = 0;
    if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) { cudaGetLastError(); fr = (size_t)8 << 30; }
    uint64_t fit = (uint64_t)((double)fr * mem_frac / (double)batch_bytes_per_block());
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

// stages 3b..3d + pool for blocks [b_lo, b_hi) of ctx->blocks (their RLE1 bytes are in place), in batches
int compress_block_range(Ctx *ctx, uint64_t b_lo, uint64_t b_hi)
{
    const uint64_t nb = b_hi > b_lo ? b_hi - b_lo : 0;
    if (!nb) return S3G_OK;
    // the batch buffers are grow-only: size them once for the batch this input needs
    uint64_t held = 0;
    DevBuf *batch_bufs[] = {&ctx->sa, &ctx->rk, &ctx->kv0, &ctx->kv1, &ctx->hist, &ctx->lcol, &ctx->mtf0, &ctx->mtfv16, &ctx->bits};
    for (DevBuf *b : batch_bufs) held += b->cap;
    // steady state: the buffers already hold this many blocks; cudaMemGetInfo is only asked when they must grow
    // (it takes anywhere from 0.1 to 20 ms on a shared host)
    uint64_t have = held / batch_bytes_per_block();
    uint64_t batch = have >= nb && !getenv("S3G_BATCH") ? nb : pick_batch(nb, ctx->mem_frac);
    if (have > batch && !getenv("S3G_BATCH")) batch = std::min<uint64_t>(nb, have);
    for (uint64_t b0 = b_lo; b0 < b_hi; b0 += batch) {
        uint64_t cnt = std::min<uint64_t>(batch, b_hi - b0);
        stage_mark(ctx, 2);
        S3G_TRY(run_bwt(ctx, b0, cnt));
        stage_mark(ctx, 3);
        S3G_TRY(run_mtf(ctx, b0, cnt));
        stage_mark(ctx, 4);
        S3G_TRY(run_huff(ctx, b0, cnt, 1, nullptr, nullptr));
        S3G_TRY(run_pool_append(ctx, b0, cnt));
    }
    return S3G_OK;
}

// stages 3a..3e over `n_streams` streams laid out back to back in d_in
static int compress_streams(Ctx *ctx, const uint8_t *d_in, uint64_t n, const uint64_t *d_soff, uint64_t n_streams, int level,
                            uint64_t *n_blocks_out, uint64_t *total_bytes)
{
    CutResult cut;
    stage_mark(ctx, 1);
    S3G_TRY(run_rle_cut(ctx, d_in, n, d_soff, n_streams, level, &cut));
    ctx->last_rle_bytes = cut.rle_bytes;
    uint64_t nb = cut.n_blocks;
    *n_blocks_out = nb;
    ctx->pool_words = 0;
    S3G_TRY(ctx->pool_woff.ensure((nb + 1) * 8));
    S3G_TRY(compress_block_range(ctx, 0, nb));
    stage_mark(ctx, 5);
    S3G_TRY(run_assemble(ctx, nb, n_streams, level, total_bytes));
    stage_mark(ctx, -1);
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

// ---- one range of the input: [off, off + len) of the device buffer d_bed --------------------------
// A range starts at the first line of a chromosome.  Unless it is the last range of the input, its
// last chromosome may continue in the bytes that follow, so it is left to the next range
// (PartOut.cs_next = where it starts).
struct PartOut {
    std::vector<s3g_chrom> chroms;     // the chromosomes compressed here; name_off absolute, bz_off relative to this part's streams
    std::vector<uint8_t> names;        // their names, back to back
    uint64_t n_seen = 0;               // chromosomes found in the range
    uint64_t cs_next = 0;              // absolute offset of the chromosome left for the next range
    uint64_t streams_size = 0, n_blocks = 0, dropped = 0, tf_kept = 0, rle_bytes = 0, mtf_symbols = 0;
    uint64_t unsorted = 0, crlf = 0;   // input diagnostics over the lines of the chromosomes compressed here
    double stage_ms[8] = {0, 0, 0, 0, 0, 0, 0, 0};
};

// tokenise + transform the range; decides how many chromosomes are compressed here
static int part_front(Ctx *ctx, const uint8_t *d_bed, uint64_t off, uint64_t len, bool last_part, PartOut &po, TfResult &tr)
{
    const uint64_t base = off & ~(uint64_t)15;                 // the tokenizer loads 16-byte vectors
    const bool timing = getenv("S3G_TIMING") != nullptr;
    double tt0 = timing ? host_ms() : 0;
    ctx->marks.clear();
    ctx->last_rle_bytes = 0;
    stage_mark(ctx, 0);
    S3G_TRY(run_transform(ctx, d_bed + base, len + (off - base), &tr, false, (uint32_t)(off - base), last_part));
    if (timing) fprintf(stderr, "[s3g timing] transform %.2f ms (host clock)\n", host_ms() - tt0);
    po.n_seen = tr.n_chroms;
    po.dropped = tr.dropped;
    ctx->h_chroms.resize(tr.n_chroms);
    if (tr.n_chroms) {
        S3G_CUDA(cudaMemcpyAsync(ctx->h_chroms.data(), ctx->chroms.p, tr.n_chroms * sizeof(s3g_chrom), cudaMemcpyDeviceToHost, ctx->stream));
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
        po.unsorted = front_unsorted(ctx); po.crlf = front_crlf(ctx);
    }
    const uint64_t kept = last_part ? tr.n_chroms : (tr.n_chroms ? tr.n_chroms - 1 : 0);
    po.cs_next = (!last_part && tr.n_chroms) ? base + ctx->h_chroms[tr.n_chroms - 1].name_off : off + len;
    po.tf_kept = kept == tr.n_chroms ? tr.tf_len : ctx->h_chroms[kept].tf_off;
    po.chroms.assign(ctx->h_chroms.begin(), ctx->h_chroms.begin() + kept);
    return S3G_OK;
}

// RLE1 .. bit assembly for the chromosomes kept by part_front; stream table and names back to the host
static int part_back(Ctx *ctx, const uint8_t *d_bed, uint64_t off, int level, bool want_names, PartOut &po)
{
    const uint64_t base = off & ~(uint64_t)15;
    const uint64_t kept = po.chroms.size();
    uint64_t total_bytes = 0, n_blocks = 0;
    if (kept) {
        S3G_TRY(ctx->soff.ensure((kept + 2) * 8));
        S3G_LAUNCH(ctx, k_soff_from_chroms, (unsigned)((kept + 1 + 127) / 128), 128, 0, ctx->chroms.as<s3g_chrom>(),
                   kept, po.tf_kept, ctx->soff.as<uint64_t>());
        S3G_TRY(compress_streams(ctx, ctx->tf.as<uint8_t>(), po.tf_kept, ctx->soff.as<uint64_t>(), kept, level, &n_blocks,
                                 &total_bytes));
    }
    else stage_mark(ctx, -1);
    S3G_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    po.n_blocks = n_blocks;
    po.streams_size = total_bytes;
    po.rle_bytes = ctx->last_rle_bytes;
    ctx->last_streams_size = total_bytes;
    ctx->last_streams_host = nullptr;
    ctx->h_blocks.resize(n_blocks);
    if (n_blocks) S3G_CUDA(cudaMemcpyAsync(ctx->h_blocks.data(), ctx->blocks.p, n_blocks * sizeof(BlockInfo), cudaMemcpyDeviceToHost, ctx->stream));
    std::vector<StreamMeta> meta(kept);
    std::vector<uint64_t> name_off(kept + 1, 0);
    for (uint64_t c = 0; c < kept; c++) name_off[c + 1] = name_off[c] + po.chroms[c].name_len;
    po.names.assign(name_off[kept] + 1, 0);
    if (kept) {
        S3G_CUDA(cudaMemcpyAsync(meta.data(), ctx->stream_meta.p, kept * sizeof(StreamMeta), cudaMemcpyDeviceToHost, ctx->stream));
        if (want_names) {
            // chromosome names (a few bytes each) come back through one gather
            S3G_TRY(ctx->io_a.ensure((kept + 1) * 8));
            S3G_TRY(ctx->io_b.ensure(name_off[kept] + 16));
            S3G_TRY(upload_small(ctx, 0, ctx->io_a.p, name_off.data(), (kept + 1) * 8));     // not behind the next range's upload on the copy engine
            S3G_LAUNCH(ctx, k_gather_names, (unsigned)kept, 32, 0, d_bed + base, ctx->chroms.as<s3g_chrom>(), kept,
                       ctx->io_a.as<uint64_t>(), ctx->io_b.as<uint8_t>());
            S3G_CUDA(cudaMemcpyAsync(po.names.data(), ctx->io_b.p, name_off[kept], cudaMemcpyDeviceToHost, ctx->stream));
        }
    }
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    for (const BlockInfo &B : ctx->h_blocks) po.mtf_symbols += B.n_mtf;
    stage_collect(ctx, po.stage_ms);
    for (uint64_t c = 0; c < kept; c++) {
        po.chroms[c].bz_off = meta[c].byte_off; po.chroms[c].bz_len = meta[c].byte_len;
        po.chroms[c].n_blocks = (uint32_t)meta[c].n_blocks;
        po.chroms[c].name_off += base;
    }
    return S3G_OK;
}

// chromosomes whose name already opened an earlier stream (interleaved / unsorted input: hpp:331 compares with the
// previous line only, so each reappearance is a stream of its own)
static uint64_t count_reappearing(const std::vector<s3g_chrom> &ch, const uint8_t *names)
{
    std::vector<std::string> v;
    v.reserve(ch.size());
    uint64_t off = 0;
    for (const s3g_chrom &c : ch) { v.emplace_back(reinterpret_cast<const char *>(names + off), c.name_len); off += c.name_len; }
    std::sort(v.begin(), v.end());
    uint64_t dup = 0;
    for (size_t i = 1; i < v.size(); i++) if (v[i] == v[i - 1]) dup++;
    return dup;
}

static int ensure_archive(Ctx *ctx, uint64_t bytes)
{
    if (bytes <= ctx->h_archive_cap) return S3G_OK;
    uint8_t *p = nullptr;
    size_t want = bytes + bytes / 8 + 4096;
    if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); set_error("out of pinned host memory"); return S3G_E_NOMEM; }
    if (ctx->h_archive) cudaFreeHost(ctx->h_archive);
    ctx->h_archive = p; ctx->h_archive_cap = want;
    return S3G_OK;
}

static void fill_result(s3g_result *res, const std::vector<s3g_chrom> &chroms)
{
    res->n_chroms = chroms.size();
    res->n_lines = 0; res->tf_bytes = 0;
    for (const s3g_chrom &c : chroms) { res->n_lines += (uint64_t)c.line_count; res->tf_bytes += c.tf_len; }
    res->chroms = (s3g_chrom *)malloc(std::max<size_t>(1, chroms.size()) * sizeof(s3g_chrom));
    if (res->chroms && !chroms.empty()) memcpy(res->chroms, chroms.data(), chroms.size() * sizeof(s3g_chrom));
}

static int compress_bed_impl(Ctx *ctx, const uint8_t *d_bed, uint64_t n, int level, const char *note, int want_archive,
                             s3g_result *res)
{
    memset(res, 0, sizeof *res);
    if (level < 1 || level > 9) { set_error("block_size_100k must be 1..9"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    S3G_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    PartOut po;
    TfResult tr;
    S3G_TRY(part_front(ctx, d_bed, 0, n, true, po, tr));
    S3G_TRY(part_back(ctx, d_bed, 0, level, want_archive != 0, po));
    float ms = 0;
    S3G_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    res->device_ms = ms;
    res->n_blocks = po.n_blocks;
    res->d_streams = ctx->streams.p;
    res->streams_size = po.streams_size;
    res->dropped_tail_bytes = po.dropped;
    res->rle_bytes = po.rle_bytes; res->mtf_symbols = po.mtf_symbols;
    res->unsorted_lines = po.unsorted; res->crlf_lines = po.crlf;
    if (want_archive) res->reappearing_chroms = count_reappearing(po.chroms, po.names.data());
    memcpy(res->stage_ms, po.stage_ms, sizeof res->stage_ms);
    ctx->h_chroms = po.chroms;
    fill_result(res, po.chroms);
    if (!res->chroms) { set_error("out of host memory"); return S3G_E_NOMEM; }
    res->n_lines = tr.n_lines; res->tf_bytes = tr.tf_len;
    if (!want_archive) return S3G_OK;
    std::vector<uint64_t> name_off(po.chroms.size() + 1, 0);
    for (size_t c = 0; c < po.chroms.size(); c++) name_off[c + 1] = name_off[c] + po.chroms[c].name_len;
    std::string hdr = build_header(po.names.data(), name_off, po.chroms, level, note);
    uint64_t streams_off = 4 + hdr.size() + 1;
    res->archive_size = streams_off + po.streams_size;
    // the archive is assembled in context-owned pinned memory: the device-to-host copy of the
    // streams lands directly in its final place
    S3G_TRY(ensure_archive(ctx, res->archive_size));
    res->archive = ctx->h_archive;
    static const uint8_t magic[4] = {0xca, 0x5c, 0xad, 0x1a};      // hpp:907-910
    memcpy(res->archive, magic, 4);
    memcpy(res->archive + 4, hdr.data(), hdr.size());
    res->archive[4 + hdr.size()] = '\n';
    res->streams_off = streams_off;
    if (po.streams_size) {
        S3G_CUDA(cudaMemcpyAsync(res->archive + streams_off, ctx->streams.p, po.streams_size, cudaMemcpyDeviceToHost, ctx->stream));
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    return S3G_OK;
}

// ---- pipelined host entry --------------------------------------------------------------------------
// The upload of a large input takes a third as long as compressing it, so s3g_compress_bed cuts the input
// into ranges at line starts, queues all uploads on a copy stream, and lets two worker contexts (own
// stream, own buffers, own host thread) compress range after range as their bytes arrive.  A range
// compresses the chromosomes that END inside it; the chromosome still open at its end is handed to the next
// range, which starts over at that chromosome's first line (already on the device).  Chromosomes are
// independent bzip2 streams, so the archive is the header plus the streams of the ranges in order -- the
// same bytes as the one-shot path.  Transforms are serialised (each range needs the previous hand-over),
// everything after the transform overlaps.
constexpr uint64_t PIPE_MIN_BYTES = 48ull << 20;      // smaller inputs go through in one piece
constexpr uint64_t HDR_RESERVE = 1ull << 20;          // the header is right-aligned in front of the streams

struct PipeShared {
    std::mutex mu;
    std::condition_variable cv;
    std::vector<char> queued;      // range i's upload has been handed to the copy stream (its event is recorded)
    int next_front = 0;            // range whose transform may start
    uint64_t cs = 0;               // where it starts
    bool single_chrom = false;     // a range ended inside the chromosome it began with: stop cutting, the last range takes all
    int next_copy = 0;             // range whose streams are copied out next
    uint64_t streams_so_far = 0, archive_cap = 0;
    int rc = S3G_OK;
    std::string err;
    double t0 = 0;                 // host clock at the start of the call (S3G_TIMING)
};

static void pipe_worker(Ctx *main_ctx, Ctx *w, int wid, int nparts, const std::vector<uint64_t> &cut, int level,
                        std::vector<PartOut> &parts, PipeShared &sh)
{
    cudaSetDevice(w->device);
    const uint8_t *d_bed = main_ctx->bed.as<uint8_t>();
    for (int i = wid; i < nparts; i += 2) {
        const bool last = i == nparts - 1;
        uint64_t off;
        {
            std::unique_lock<std::mutex> lk(sh.mu);
            sh.cv.wait(lk, [&] { return sh.next_front == i || sh.rc != S3G_OK; });
            if (sh.rc != S3G_OK) return;
            off = sh.cs;
            if (sh.single_chrom && !last) {
                // nothing to do for this range: pass both turns on
                sh.next_front = i + 1;
                sh.cv.notify_all();
                sh.cv.wait(lk, [&] { return sh.next_copy == i || sh.rc != S3G_OK; });
                if (sh.rc != S3G_OK) return;
                sh.next_copy = i + 1;
                sh.cv.notify_all();
                continue;
            }
        }
        int rc = S3G_OK;
        TfResult tr;
        PartOut &po = parts[i];
        {
            // a pageable input is staged by copier threads: the range's event exists only once its last piece is queued
            std::unique_lock<std::mutex> lk(sh.mu);
            sh.cv.wait(lk, [&] { return sh.queued[i] || sh.rc != S3G_OK; });
            if (sh.rc != S3G_OK) return;
        }
        const bool timing = sh.t0 != 0;
        const double tf0 = timing ? host_ms() : 0;
        if (cudaStreamWaitEvent(w->stream, main_ctx->part_ev[i], 0) != cudaSuccess) rc = S3G_E_CUDA;
        if (rc == S3G_OK) rc = part_front(w, d_bed, off, cut[i + 1] - off, last, po, tr);
        const double tf1 = timing ? host_ms() : 0;
        {
            std::unique_lock<std::mutex> lk(sh.mu);
            if (rc != S3G_OK) { sh.rc = rc; sh.err = g_err; sh.cv.notify_all(); return; }
            if (!last && po.chroms.empty()) sh.single_chrom = true;
            sh.cs = po.cs_next;
            sh.next_front = i + 1;
            sh.cv.notify_all();
        }
        rc = part_back(w, d_bed, off, level, true, po);
        if (timing)
            fprintf(stderr, "[s3g timing] range %d (worker %d): front %.2f .. %.2f ms, %llu chromosomes, %llu blocks coded by %.2f ms\n", i, wid, tf0 - sh.t0,
                    tf1 - sh.t0, (unsigned long long)po.chroms.size(), (unsigned long long)po.n_blocks, host_ms() - sh.t0);
        // streams of the ranges go out in order, straight into their place in the pinned archive
        {
            std::unique_lock<std::mutex> lk(sh.mu);
            if (rc != S3G_OK) { sh.rc = rc; sh.err = g_err; sh.cv.notify_all(); return; }
            sh.cv.wait(lk, [&] { return sh.next_copy == i || sh.rc != S3G_OK; });
            if (sh.rc != S3G_OK) return;
            uint64_t at = HDR_RESERVE + sh.streams_so_far;
            if (at + po.streams_size > sh.archive_cap) { sh.rc = S3G_E_CAPACITY; sh.err = "pinned archive buffer too small"; sh.cv.notify_all(); return; }
            for (s3g_chrom &c : po.chroms) c.bz_off += sh.streams_so_far;
            sh.streams_so_far += po.streams_size;
            sh.next_copy = i + 1;
            sh.cv.notify_all();
            lk.unlock();
            if (po.streams_size) {
                if (cudaMemcpyAsync(main_ctx->h_archive + at, w->streams.p, po.streams_size, cudaMemcpyDeviceToHost, w->stream) != cudaSuccess ||
                    cudaStreamSynchronize(w->stream) != cudaSuccess) {
                    std::unique_lock<std::mutex> lk2(sh.mu);
                    sh.rc = S3G_E_CUDA; sh.err = "device-to-host copy of the streams failed"; sh.cv.notify_all();
                    return;
                }
            }
            if (timing) fprintf(stderr, "[s3g timing] range %d: streams on the host %.2f ms\n", i, host_ms() - sh.t0);
        }
    }
}

// Pageable input (the CLI reads files into malloc'd memory): cudaMemcpyAsync from it is synchronous and staged by the
// driver at a third of the link's speed, and it would hold up the calling thread while the workers wait.  Instead
// stage_threads() copier threads (8; S3G_STAGE_THREADS) move the input through pinned pieces of their own (two each, so a thread copies one
// while the other is on its way to the device); the thread that queues the last piece of a range records the range's event.
constexpr int STAGE_THREADS_MAX = 16;
static int stage_threads()
{
    static const int n = [] { int v = 8; if (const char *e = getenv("S3G_STAGE_THREADS")) v = atoi(e); return std::max(1, std::min(STAGE_THREADS_MAX, v)); }();
    return n;
}
constexpr uint64_t STAGE_PIECE = 8ull << 20;
static void stage_copier(Ctx *ctx, int tid, const uint8_t *bed, uint64_t n, const std::vector<uint64_t> &cut, int nparts,
                         std::vector<int> &left, PipeShared &sh)
{
    cudaSetDevice(ctx->device);
    uint8_t *slot[2] = {ctx->h_stage + (uint64_t)(2 * tid) * STAGE_PIECE, ctx->h_stage + (uint64_t)(2 * tid + 1) * STAGE_PIECE};
    cudaEvent_t ev[2] = {ctx->stage_ev[2 * tid], ctx->stage_ev[2 * tid + 1]};
    bool used[2] = {false, false};
    uint64_t piece = 0;
    int k = 0;
    for (int i = 0; i < nparts; i++) {
        for (uint64_t off = cut[i]; off < cut[i + 1] || (off == cut[i] && cut[i] == cut[i + 1]); off += STAGE_PIECE, piece++) {
            const uint64_t len = std::min<uint64_t>(STAGE_PIECE, cut[i + 1] - off);
            if ((int)(piece % stage_threads()) == tid) {
                bool ok = true;
                if (len) {
                    if (used[k]) ok = cudaEventSynchronize(ev[k]) == cudaSuccess;
                    memcpy(slot[k], bed + off, len);
                    ok = ok && cudaMemcpyAsync(ctx->bed.as<uint8_t>() + off, slot[k], len, cudaMemcpyHostToDevice, ctx->copy_stream) == cudaSuccess;
                    ok = ok && cudaEventRecord(ev[k], ctx->copy_stream) == cudaSuccess;
                    used[k] = true; k ^= 1;
                }
                std::unique_lock<std::mutex> lk(sh.mu);
                if (!ok) { sh.rc = S3G_E_CUDA; sh.err = "staged upload failed"; sh.cv.notify_all(); return; }
                if (--left[i] == 0) {
                    if (i == nparts - 1) cudaMemsetAsync(ctx->bed.as<uint8_t>() + n, 0, 64, ctx->copy_stream);
                    cudaEventRecord(ctx->part_ev[i], ctx->copy_stream);
                    sh.queued[i] = 1;
                    sh.cv.notify_all();
                }
            }
            if (cut[i] == cut[i + 1]) break;
        }
    }
    for (int q = 0; q < 2; q++) if (used[q]) cudaEventSynchronize(ev[q]);
}

// Queues the upload of the ranges cut[0..nparts] of `bed` into ctx->bed on the copy stream, one event (ctx->part_ev[i]) per
// range; sh.queued[i] says that range i's event has been recorded.  Pinned input: everything is queued before this returns.
// Pageable input: copier threads (join them) stage it through pinned pieces.
static int start_upload(Ctx *ctx, const uint8_t *bed, uint64_t n, const std::vector<uint64_t> &cut, int nparts, PipeShared &sh,
                        std::vector<int> &left, std::vector<std::thread> &copiers)
{
    // is the caller's buffer pinned?  (an unregistered pointer reports cudaMemoryTypeUnregistered, or an error on old drivers)
    bool pinned_input = false;
    {
        cudaPointerAttributes at;
        if (cudaPointerGetAttributes(&at, bed) == cudaSuccess) pinned_input = at.type == cudaMemoryTypeHost || at.type == cudaMemoryTypeManaged;
        else cudaGetLastError();
        if (getenv("S3G_NO_STAGING")) pinned_input = true;
    }
    sh.queued.assign(nparts, 0);
    left.assign(nparts, 0);
    if (pinned_input) {
        for (int i = 0; i < nparts; i++) {
            if (cut[i + 1] > cut[i])
                S3G_CUDA(cudaMemcpyAsync(ctx->bed.as<uint8_t>() + cut[i], bed + cut[i], cut[i + 1] - cut[i], cudaMemcpyHostToDevice, ctx->copy_stream));
            if (i == nparts - 1) S3G_CUDA(cudaMemsetAsync(ctx->bed.as<uint8_t>() + n, 0, 64, ctx->copy_stream));
            S3G_CUDA(cudaEventRecord(ctx->part_ev[i], ctx->copy_stream));
            sh.queued[i] = 1;
        }
    } else {
        if (!ctx->h_stage) {
            if (cudaMallocHost(&ctx->h_stage, 2ull * STAGE_THREADS_MAX * STAGE_PIECE) != cudaSuccess) { cudaGetLastError(); set_error("out of pinned host memory"); return S3G_E_NOMEM; }
            for (int q = 0; q < 2 * STAGE_THREADS_MAX; q++) {
                cudaEvent_t e;
                S3G_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
                ctx->stage_ev.push_back(e);
            }
        }
        for (int i = 0; i < nparts; i++) left[i] = cut[i + 1] > cut[i] ? (int)((cut[i + 1] - cut[i] + STAGE_PIECE - 1) / STAGE_PIECE) : 1;
        for (int t = 0; t < stage_threads(); t++)
            copiers.emplace_back(stage_copier, ctx, t, bed, n, std::cref(cut), nparts, std::ref(left), std::ref(sh));
    }
    return S3G_OK;
}

// copy stream and one event per range
static int ensure_pipe_objects(Ctx *ctx, int nparts)
{
    if (!ctx->copy_stream) S3G_CUDA(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    while ((int)ctx->part_ev.size() < nparts) {
        cudaEvent_t e;
        S3G_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        ctx->part_ev.push_back(e);
    }
    return S3G_OK;
}

static int compress_bed_pipelined(Ctx *ctx, const uint8_t *bed, uint64_t n, int level, const char *note, int nparts, s3g_result *res)
{
    memset(res, 0, sizeof *res);
    if (level < 1 || level > 9) { set_error("block_size_100k must be 1..9"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    S3G_TRY(ensure_pipe_objects(ctx, nparts));
    for (int k = 0; k < 2; k++)
        if (!ctx->sub[k]) {
            s3g_ctx *c = nullptr;
            S3G_TRY(s3g_init(ctx->device, &c));
            c->mem_frac = 0.4;                      // two workers size their batches at the same time
            ctx->sub[k] = c;
        }
    // cut points at line starts; the first range is shorter than the others (nothing can run until it has arrived)
    std::vector<uint64_t> cut(nparts + 1, n);
    cut[0] = 0;
    double first = 0.5 / nparts;
    if (const char *e = getenv("S3G_FIRST")) { double v = atof(e); if (v > 0 && v < 1) first = v; }
    for (int i = 1; i < nparts; i++) {
        double frac = first + (1.0 - first) * (i - 1) / (nparts - 1);
        uint64_t p = std::max<uint64_t>(cut[i - 1], (uint64_t)(frac * (double)n));
        const void *nl = p < n ? memchr(bed + p, '\n', n - p) : nullptr;
        cut[i] = nl ? (uint64_t)((const uint8_t *)nl - bed) + 1 : n;
    }
    S3G_TRY(ctx->bed.ensure(n + 64));
    // the archive buffer must exist before the workers copy into it: the compressed size is not known yet,
    // half of the input is generous for BED (the one-shot path takes over if it ever is not)
    S3G_TRY(ensure_archive(ctx, HDR_RESERVE + std::max<uint64_t>(n / 2, ctx->archive_hint + (ctx->archive_hint >> 3)) + 4096));
    cudaEvent_t t0 = ctx->ev0, t1 = ctx->ev1;
    std::vector<PartOut> parts(nparts);
    PipeShared sh;
    sh.archive_cap = ctx->h_archive_cap;
    if (getenv("S3G_TIMING")) sh.t0 = host_ms();
    std::vector<int> left;
    std::vector<std::thread> copiers;
    S3G_CUDA(cudaEventRecord(t0, ctx->copy_stream));
    {
        int rc = start_upload(ctx, bed, n, cut, nparts, sh, left, copiers);
        if (rc != S3G_OK) { for (std::thread &c : copiers) c.join(); return rc; }
    }
    std::thread th0(pipe_worker, ctx, ctx->sub[0], 0, nparts, std::cref(cut), level, std::ref(parts), std::ref(sh));
    std::thread th1(pipe_worker, ctx, ctx->sub[1], 1, nparts, std::cref(cut), level, std::ref(parts), std::ref(sh));
    th0.join(); th1.join();
    for (std::thread &c : copiers) c.join();
    if (sh.rc == S3G_E_CAPACITY) return S3G_E_CAPACITY;       // caller falls back to the one-shot path
    if (sh.rc != S3G_OK) { set_error("%s", sh.err.c_str()); return sh.rc; }
    S3G_CUDA(cudaEventRecord(t1, ctx->sub[(nparts - 1) & 1]->stream));
    S3G_CUDA(cudaEventSynchronize(t1));
    float ms = 0;
    S3G_CUDA(cudaEventElapsedTime(&ms, t0, t1));
    res->device_ms = ms;                                       // first upload to last kernel
    // merge
    std::vector<s3g_chrom> chroms;
    std::vector<uint8_t> names;
    for (PartOut &po : parts) {
        chroms.insert(chroms.end(), po.chroms.begin(), po.chroms.end());
        if (!po.chroms.empty()) names.insert(names.end(), po.names.begin(), po.names.end() - 1);
        res->n_blocks += po.n_blocks;
        res->rle_bytes += po.rle_bytes; res->mtf_symbols += po.mtf_symbols;
        res->unsorted_lines += po.unsorted; res->crlf_lines += po.crlf;
    }
    names.push_back(0);
    res->reappearing_chroms = count_reappearing(chroms, names.data());
    { uint64_t t = 0; for (s3g_chrom &c : chroms) { c.tf_off = t; t += c.tf_len; } }    // as in one transformed buffer
    res->dropped_tail_bytes = parts[nparts - 1].dropped;
    for (int k = 0; k < 2; k++) { ctx->launches += ctx->sub[k]->launches; ctx->sub[k]->launches = 0; }
    ctx->h_chroms = chroms;
    fill_result(res, chroms);
    if (!res->chroms) { set_error("out of host memory"); return S3G_E_NOMEM; }
    std::vector<uint64_t> name_off(chroms.size() + 1, 0);
    for (size_t c = 0; c < chroms.size(); c++) name_off[c + 1] = name_off[c] + chroms[c].name_len;
    std::string hdr = build_header(names.data(), name_off, chroms, level, note);
    const uint64_t streams_off = 4 + hdr.size() + 1;
    if (streams_off > HDR_RESERVE) return S3G_E_CAPACITY;      // enormous chromosome table: one-shot path
    uint8_t *arc = ctx->h_archive + HDR_RESERVE - streams_off;
    static const uint8_t magic[4] = {0xca, 0x5c, 0xad, 0x1a};      // hpp:907-910
    memcpy(arc, magic, 4);
    memcpy(arc + 4, hdr.data(), hdr.size());
    arc[4 + hdr.size()] = '\n';
    res->archive = arc;
    res->streams_off = streams_off;
    res->streams_size = sh.streams_so_far;
    res->archive_size = streams_off + sh.streams_so_far;
    res->d_streams = nullptr;                                  // the streams of the ranges lived in the worker contexts
    ctx->last_streams_size = sh.streams_so_far;                // s3g_read_streams serves them from the pinned archive
    ctx->last_streams_host = ctx->h_archive + HDR_RESERVE;
    ctx->archive_hint = sh.streams_so_far;
    return S3G_OK;
}

// ---- chained host entry: ranges of ONE chromosome ---------------------------------------------------
// The pipelined entry above hands out whole chromosomes, so an input that is one long chromosome (BASELINE.json
// configs 1 and 3) gets no overlap from it: the upload and the compression run one after the other.  Here the unit is the
// bzip2 BLOCK.  The input is cut into ranges at line starts; as soon as a range has arrived it is tokenised and
// transformed with the phases of the N-GPU path (shard.cu: the line before the range is its "halo", the largest stop so
// far is carried), its transformed bytes go behind the tail the step before left unfinished, the block cut runs over
// tail + new bytes, and every block whose cut can no longer move is sorted and coded while the next range is still on its
// way.  A block's cut depends on the bytes up to the first byte of the run that did not fit (bz/bzlib.c:225-293, :307),
// so a block that ends at least two bytes before the end of what is there is final; the bytes after the last final
// block (less than two blocks) are the next step's tail, and a step starts there with the clean run state the reference
// has at that point (the pending run goes to the next block whole).  The bits of a step continue where the step before
// stopped (bz/compress.c:609 keeps bsBuff across blocks): stream headers, trailers and combined CRCs come from the
// carried state, the byte strings meet in a device buffer whose seam bytes are ORed (k_place_bytes).  The archive is
// the same bytes as the one-shot path's.
static int grow_keep(Ctx *ctx, DevBuf &b, uint64_t need, uint64_t keep);
struct ChainStream {
    std::string name;
    s3g_chrom c;
};

// ---- the bit layout of one step: host arithmetic only (s3g_chain_layout exposes it to the CPU tests) ----
// Carried from step to step: is the last stream still open, its bits so far (header included) and combined CRC so far, and the
// bytes of the closed streams (= where the open or the next stream starts).
struct ChainCarry {
    bool open = false;
    uint64_t open_bits = 0; uint32_t open_comb = 0;
    uint64_t out_bytes = 0;
};
struct StepLayout {
    std::vector<uint64_t> pos;        // bit position of every final block, in order
    std::vector<uint64_t> patches;    // (bit position, 32-bit word) pairs: stream headers and trailers
    uint64_t bit_lo = ~0ull, bit_hi = 0;
    std::vector<uint64_t> started;    // per step stream: byte offset it starts at in this step, ~0 if it continues the open stream
    std::vector<uint64_t> closed_len; // per step stream: its length in bytes if it is closed in this step, 0 if still open
    std::vector<uint32_t> n_final;    // per step stream: blocks placed in this step
};
// blocks [0, nb) of the step in stream order (stream_of ascending from 0 to ns - 1), the first b_fin of them final; the
// step's first stream continues the open stream when first_continues.  bz/compress.c:607-609, :622-628, :657-666.
static int chain_layout_step(ChainCarry &st, int level, uint64_t ns, bool first_continues, const uint32_t *stream_of, const uint64_t *n_bits,
                             const uint32_t *crc, uint64_t nb, uint64_t b_fin, StepLayout &L)
{
    L.pos.clear(); L.patches.clear(); L.bit_lo = ~0ull; L.bit_hi = 0;
    L.started.assign(ns, ~0ull); L.closed_len.assign(ns, 0); L.n_final.assign(ns, 0);
    auto mark = [&](uint64_t at, uint64_t nbits) { L.bit_lo = std::min(L.bit_lo, at); L.bit_hi = std::max(L.bit_hi, at + nbits); };
    uint64_t b = 0;
    for (uint64_t s = 0; s < ns; s++) {
        uint64_t bits; uint32_t comb;
        if (s == 0 && first_continues) {
            if (!st.open) return S3G_E_PARAM;
            bits = st.open_bits; comb = st.open_comb;
        } else {
            L.started[s] = st.out_bytes;
            L.patches.push_back(st.out_bytes * 8); L.patches.push_back(0x425a6800u | (uint32_t)('0' + level));
            mark(st.out_bytes * 8, 32);
            bits = 32; comb = 0;
        }
        bool all_final = true;
        for (; b < nb && stream_of[b] == s; b++) {
            if (b >= b_fin) { all_final = false; continue; }
            L.pos.push_back(st.out_bytes * 8 + bits);
            mark(st.out_bytes * 8 + bits, n_bits[b]);
            bits += n_bits[b];
            comb = ((comb << 1) | (comb >> 31)) ^ crc[b];
            L.n_final[s]++;
        }
        if (all_final) {
            const uint64_t end = st.out_bytes * 8 + bits;
            L.patches.push_back(end); L.patches.push_back(0x17724538u);
            L.patches.push_back(end + 32); L.patches.push_back(0x50900000u | (comb >> 16));
            L.patches.push_back(end + 64); L.patches.push_back((uint64_t)(uint32_t)(comb << 16));
            mark(end, 80);
            bits += 80;
            L.closed_len[s] = (bits + 7) >> 3;
            st.out_bytes += L.closed_len[s];
            st.open = false;
        } else {
            st.open = true; st.open_bits = bits; st.open_comb = comb;
        }
    }
    return b == nb ? S3G_OK : S3G_E_PARAM;
}

// the state that goes from one step to the next, and one step
struct Chain {
    Ctx *ctx = nullptr;
    int level = 9;
    std::vector<ChainStream> streams;            // the archive's streams, in order; the last one may be open
    ChainCarry cst;                              // open: streams.back() has blocks still to come
    StepLayout lay;
    std::vector<uint32_t> l_stream, l_crc; std::vector<uint64_t> l_bits;
    DevBuf *tf[2] = {nullptr, nullptr};          // the step's transformed bytes, ping-pong (the caller's: the tail lives there between steps)
    uint64_t tail_len = 0;                       // unfinished bytes of the open stream at the start of tf[cur]
    int cur = 0;
    int64_t run_max = INT64_MIN;                 // largest stop since the last chromosome start (carry_chain of multi.cu)
    uint64_t tf_room = 4ull << 20;               // transformed bytes a step is expected to add (sizes the other buffer)
    uint64_t n_lines = 0, n_blocks = 0, rle_bytes = 0, mtf_symbols = 0, tf_bytes = 0, dropped = 0, unsorted = 0, crlf = 0;
    // where the compressed bytes go.  Either a zeroed device buffer of out_cap bytes (the seam bytes of consecutive steps are
    // ORed there by k_place_bytes) whose complete bytes are copied to h_out on the context's out_stream beside the next
    // step's kernels; or a host vector (the seam byte is ORed on the host).
    uint8_t *d_out = nullptr; uint64_t out_cap = 0; uint8_t *h_out = nullptr; uint64_t copied = 0;
    std::vector<uint8_t> *v_out = nullptr;
    std::vector<s3g_chrom> pc = std::vector<s3g_chrom>(256);
    std::vector<uint64_t> soff, items, patches;
    std::vector<uint32_t> gidx;                  // step stream -> index into `streams`
    std::vector<uint8_t> tmp;
    bool timing = false; double tt0 = 0; int step_no = 0;

    // d_range: the range on the device, its first `halo` bytes the line before it; h_own: host copy of the range's own bytes
    // (chromosome names are read there); name_base: offset of the range's own bytes in the caller's input (s3g_chrom.name_off)
    int step(const uint8_t *d_range, uint64_t len, uint64_t halo, const uint8_t *h_own, uint64_t name_base, bool last);
    void result(s3g_result *res, std::vector<s3g_chrom> &chroms, std::vector<uint8_t> &names);
};

int Chain::step(const uint8_t *d_range, uint64_t len, uint64_t halo, const uint8_t *h_own, uint64_t name_base, bool last)
{
    s3g_ctx *cx = static_cast<s3g_ctx *>(ctx);           // the phases of shard.cu take the C handle
    const double ts0 = timing ? host_ms() : 0;
    double ts1 = 0, ts2 = 0, ts3 = 0;
    // ---- tokenise + transform the range behind the tail ----
    s3g_shard_summary sm;
    S3G_TRY(s3g_shard_tokenize(cx, d_range, len, halo, &sm));
    if (timing) ts1 = host_ms();
    const int64_t carry = sm.continues ? run_max : INT64_MIN;
    if (sm.n_lines) run_max = (sm.single_piece && sm.continues) ? std::max(sm.tail_max, carry) : sm.tail_max;
    n_lines += sm.n_lines;
    if (last) dropped = sm.dropped_tail_bytes;
    DevBuf &T = *tf[cur];
    S3G_TRY(grow_keep(ctx, T, tail_len + sm.tf_bytes + 256, tail_len));
    uint64_t np = 0, tl = 0;
    for (;;) {
        const uint64_t buf = (uint64_t)(uintptr_t)T.p;
        int r2 = s3g_shard_transform_peers(cx, carry, pc.data(), pc.size(), &np, &buf, 1, 0, tail_len, &tl);
        if (r2 == S3G_E_CAPACITY && pc.size() < (1u << 24)) { pc.resize(pc.size() * 16); continue; }
        S3G_TRY(r2);
        break;
    }
    if (tl != sm.tf_bytes) { set_error("transformed size differs from the measured one"); return S3G_E_CUDA; }
    if (sm.n_lines) { unsorted += front_unsorted(ctx); crlf += front_crlf(ctx); }
    tf_bytes += tl;
    // ---- the step's streams: the tail (+ the piece that continues it), then the range's other chromosomes ----
    const uint64_t n_step = tail_len + tl;
    soff.clear(); gidx.clear();
    if (tail_len) { soff.push_back(0); gidx.push_back((uint32_t)streams.size() - 1); }
    uint64_t off = tail_len;
    for (uint64_t q = 0; q < np; q++) {
        const s3g_chrom &p = pc[q];
        if (q == 0 && sm.continues && cst.open) {
            s3g_chrom &c = streams.back().c;
            c.tf_len += p.tf_len; c.line_count += p.line_count; c.bases_nonunique += p.bases_nonunique; c.bases_unique += p.bases_unique;
            if (!tail_len) { set_error("chained entry: an open stream without a tail"); return S3G_E_CUDA; }
        } else {
            if (p.name_off < halo) { set_error("chained entry: a new chromosome that starts in the halo line"); return S3G_E_CUDA; }
            ChainStream cs;
            cs.name.assign(reinterpret_cast<const char *>(h_own + (p.name_off - halo)), p.name_len);
            memset(&cs.c, 0, sizeof cs.c);
            cs.c.name_off = name_base + (p.name_off - halo); cs.c.name_len = p.name_len;
            cs.c.tf_len = p.tf_len; cs.c.line_count = p.line_count; cs.c.bases_nonunique = p.bases_nonunique; cs.c.bases_unique = p.bases_unique;
            streams.push_back(cs);
            soff.push_back(off); gidx.push_back((uint32_t)streams.size() - 1);
        }
        off += p.tf_len;
    }
    if (off != n_step) { set_error("chained entry: piece table and transformed size disagree"); return S3G_E_CUDA; }
    const uint64_t ns = gidx.size();
    step_no++;
    if (ns == 0) return S3G_OK;                              // nothing yet (a range without a complete line)
    soff.push_back(n_step);
    // ---- block cut over tail + new bytes; the blocks whose cut is final ----
    uint64_t nb = 0;
    if (timing) ts2 = host_ms();
    S3G_TRY(s3g_shard_plan(cx, T.p, n_step, soff.data(), ns, level, &nb, nullptr, nullptr, ~0ull));
    const std::vector<BlockInfo> &hb = ctx->h_blocks;
    uint64_t b_fin = nb;
    if (!last) while (b_fin > 0 && hb[b_fin - 1].chrom == ns - 1 && hb[b_fin - 1].in_end + 2 > n_step) b_fin--;
    if (timing) ts3 = host_ms();
    S3G_TRY(s3g_shard_compress(cx, 0, b_fin, nullptr, nullptr, nullptr));
    if (timing)
        fprintf(stderr, "[s3g timing] step %d: starts %.2f ms, range measured +%.2f, transformed +%.2f, plan +%.2f (%llu blocks, %llu final), coded +%.2f\n",
                step_no - 1, ts0 - tt0, ts1 - ts0, ts2 - ts1, ts3 - ts2, (unsigned long long)nb, (unsigned long long)b_fin, host_ms() - ts3);
    // ---- where the step's bits go ----
    l_stream.resize(nb); l_bits.resize(nb); l_crc.resize(nb);
    for (uint64_t k = 0; k < nb; k++) { l_stream[k] = hb[k].chrom; l_bits[k] = hb[k].n_bits; l_crc[k] = hb[k].crc; }
    if (chain_layout_step(cst, level, ns, tail_len != 0, l_stream.data(), l_bits.data(), l_crc.data(), nb, b_fin, lay) != S3G_OK) {
        set_error("chained entry: block table and stream table disagree"); return S3G_E_CUDA;
    }
    for (uint64_t sidx = 0; sidx < ns; sidx++) {
        s3g_chrom &c = streams[gidx[sidx]].c;
        if (lay.started[sidx] != ~0ull) { c.bz_off = lay.started[sidx]; c.tf_off = 0; }
        c.n_blocks += lay.n_final[sidx];
        if (lay.closed_len[sidx]) c.bz_len = lay.closed_len[sidx];
    }
    for (uint64_t k = 0; k < b_fin; k++) { n_blocks++; rle_bytes += hb[k].nblock; mtf_symbols += hb[k].n_mtf; }
    items = lay.pos; patches = lay.patches;
    const uint64_t bit_lo = lay.bit_lo, bit_hi = lay.bit_hi;
    if (bit_hi > bit_lo) {
        const uint64_t byte_lo = bit_lo >> 3, byte_hi = (bit_hi + 7) >> 3, nbytes = byte_hi - byte_lo;
        if (d_out && byte_hi > out_cap) return S3G_E_CAPACITY;               // caller falls back to the one-shot path
        for (uint64_t &v : items) v -= byte_lo * 8;
        const uint64_t n_patch = patches.size() / 2;
        for (uint64_t k = 0; k < patches.size(); k += 2) patches[k] -= byte_lo * 8;
        items.insert(items.end(), patches.begin(), patches.end());
        S3G_TRY(run_assemble_items(ctx, 0, b_fin, items, n_patch, nbytes));
        if (d_out) {
            S3G_TRY(run_place_bytes(ctx, d_out, byte_lo, nbytes));
            // the bytes that are complete leave for the host beside the next step's kernels
            const uint64_t done = last ? cst.out_bytes : (bit_hi >> 3);
            if (h_out && done > copied) {
                S3G_CUDA(cudaEventRecord(ctx->out_ev, ctx->stream));
                S3G_CUDA(cudaStreamWaitEvent(ctx->out_stream, ctx->out_ev, 0));
                S3G_CUDA(cudaMemcpyAsync(h_out + copied, d_out + copied, done - copied, cudaMemcpyDeviceToHost, ctx->out_stream));
                copied = done;
            }
        } else {
            // the first byte may be shared with the step before (blocks are not byte aligned): OR it, copy the rest
            tmp.resize(nbytes);
            S3G_CUDA(cudaMemcpyAsync(tmp.data(), ctx->streams.p, nbytes, cudaMemcpyDeviceToHost, ctx->stream));
            S3G_CUDA(cudaStreamSynchronize(ctx->stream));
            if (v_out->size() < byte_hi) v_out->resize(byte_hi, 0);
            (*v_out)[byte_lo] |= tmp[0];
            if (nbytes > 1) memcpy(v_out->data() + byte_lo + 1, tmp.data() + 1, nbytes - 1);
        }
    }
    // ---- the tail moves to the front of the other buffer ----
    if (last) { tail_len = 0; return S3G_OK; }
    const uint64_t tail_start = (b_fin > 0 && hb[b_fin - 1].chrom == ns - 1) ? hb[b_fin - 1].in_end : soff[ns - 1];
    const uint64_t new_tail = n_step - tail_start;
    DevBuf &N = *tf[cur ^ 1];
    S3G_TRY(N.ensure(new_tail + tf_room));
    if (new_tail) S3G_CUDA(cudaMemcpyAsync(N.p, static_cast<const uint8_t *>(T.p) + tail_start, new_tail, cudaMemcpyDeviceToDevice, ctx->stream));
    tail_len = new_tail; cur ^= 1;
    return S3G_OK;
}

// statistics and the chromosome table of the finished chain
void Chain::result(s3g_result *res, std::vector<s3g_chrom> &chroms, std::vector<uint8_t> &names)
{
    res->n_lines = n_lines; res->n_blocks = n_blocks; res->rle_bytes = rle_bytes; res->mtf_symbols = mtf_symbols; res->tf_bytes = tf_bytes;
    res->dropped_tail_bytes = dropped; res->unsorted_lines = unsorted; res->crlf_lines = crlf;
    chroms.clear(); names.clear();
    uint64_t t = 0;
    for (ChainStream &cs : streams) { cs.c.tf_off = t; t += cs.c.tf_len; chroms.push_back(cs.c); names.insert(names.end(), cs.name.begin(), cs.name.end()); }
    names.push_back(0);
    res->reappearing_chroms = count_reappearing(chroms, names.data());
}

static int compress_bed_chained(Ctx *ctx, const uint8_t *bed, uint64_t n, int level, const char *note, int nparts, s3g_result *res)
{
    memset(res, 0, sizeof *res);
    if (level < 1 || level > 9) { set_error("block_size_100k must be 1..9"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    // ranges: equal sizes, cut at line starts, none empty; halo = the line before a range
    std::vector<uint64_t> cut(1, 0);
    for (int i = 1; i < nparts; i++) {
        uint64_t p = std::max<uint64_t>(cut.back(), (uint64_t)((__uint128_t)n * i / nparts));
        const void *nl = p < n ? memchr(bed + p, '\n', n - p) : nullptr;
        p = nl ? (uint64_t)((const uint8_t *)nl - bed) + 1 : n;
        if (p > cut.back() && p < n) cut.push_back(p);
    }
    cut.push_back(n);
    nparts = (int)cut.size() - 1;
    std::vector<uint64_t> halo(nparts, 0);
    for (int i = 1; i < nparts; i++) {
        uint64_t q = cut[i] - 1;                         // bed[cut[i] - 1] is the line feed that ends the halo line
        while (q > 0 && bed[q - 1] != '\n') q--;
        halo[i] = cut[i] - q;
    }
    S3G_TRY(ensure_pipe_objects(ctx, nparts));
    if (!ctx->out_stream) S3G_CUDA(cudaStreamCreateWithFlags(&ctx->out_stream, cudaStreamNonBlocking));
    if (!ctx->out_ev) S3G_CUDA(cudaEventCreateWithFlags(&ctx->out_ev, cudaEventDisableTiming));
    S3G_TRY(ctx->bed.ensure(n + 64));
    S3G_TRY(ensure_archive(ctx, HDR_RESERVE + std::max<uint64_t>(n / 2, ctx->archive_hint + (ctx->archive_hint >> 3)) + 4096));
    const uint64_t out_cap = ctx->h_archive_cap - HDR_RESERVE;
    S3G_TRY(ctx->chain_out.ensure(out_cap + 64));
    S3G_CUDA(cudaMemsetAsync(ctx->chain_out.p, 0, out_cap + 64, ctx->stream));     // seam bytes are ORed into it
    uint64_t range_max = 0;
    for (int i = 0; i < nparts; i++) range_max = std::max(range_max, cut[i + 1] - cut[i]);
    for (int k = 0; k < 2; k++) S3G_TRY(ctx->chain_tf[k].ensure(range_max + (4ull << 20)));
    cudaEvent_t t0 = ctx->ev0, t1 = ctx->ev1;
    PipeShared sh;
    std::vector<int> left;
    std::vector<std::thread> copiers;
    S3G_CUDA(cudaEventRecord(t0, ctx->copy_stream));
    int rc = start_upload(ctx, bed, n, cut, nparts, sh, left, copiers);

    Chain ch;
    ch.ctx = ctx; ch.level = level; ch.tf_room = range_max + (4ull << 20);
    ch.tf[0] = &ctx->chain_tf[0]; ch.tf[1] = &ctx->chain_tf[1];
    ch.d_out = ctx->chain_out.as<uint8_t>(); ch.out_cap = out_cap; ch.h_out = ctx->h_archive + HDR_RESERVE;
    ch.timing = getenv("S3G_TIMING") != nullptr; ch.tt0 = ch.timing ? host_ms() : 0;
    const uint8_t *d_bed = ctx->bed.as<uint8_t>();
    for (int i = 0; rc == S3G_OK && i < nparts;) {
        // A step takes the next range and every range behind it that has already arrived: when the GPU is the slower side
        // (cfg4: 21 ms of upload, 69 ms of work) the steps grow by themselves and the batches of blocks fill the GPU; when
        // the upload is (cfg3) a step is one range.
        int j = i;
        {
            std::unique_lock<std::mutex> lk(sh.mu);
            sh.cv.wait(lk, [&] { return sh.queued[i] || sh.rc != S3G_OK; });
            if (sh.rc != S3G_OK) { set_error("%s", sh.err.c_str()); rc = sh.rc; break; }
            while (j + 1 < nparts && sh.queued[j + 1] && cudaEventQuery(ctx->part_ev[j + 1]) == cudaSuccess) j++;
            cudaGetLastError();                          // cudaErrorNotReady is not an error
        }
        for (int k = i; k <= j; k++)                     // each one: copier threads finish the ranges of a pageable input in any order
            if (cudaStreamWaitEvent(ctx->stream, ctx->part_ev[k], 0) != cudaSuccess) { set_error("cudaStreamWaitEvent failed"); rc = S3G_E_CUDA; }
        if (rc != S3G_OK) break;
        const uint64_t lo = cut[i] - halo[i];
        rc = ch.step(d_bed + lo, cut[j + 1] - lo, halo[i], bed + cut[i], cut[i], j == nparts - 1);
        i = j + 1;
    }
    if (ch.timing) fprintf(stderr, "[s3g timing] last step ends %.2f ms\n", host_ms() - ch.tt0);
    for (std::thread &c : copiers) c.join();
    if (rc != S3G_OK) { cudaStreamSynchronize(ctx->copy_stream); cudaStreamSynchronize(ctx->out_stream); cudaStreamSynchronize(ctx->stream); return rc; }
    if (ch.cst.open) { set_error("chained entry: a stream was left open"); return S3G_E_CUDA; }
    S3G_CUDA(cudaEventRecord(t1, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->out_stream));
    S3G_CUDA(cudaEventSynchronize(t1));
    float ms = 0;
    S3G_CUDA(cudaEventElapsedTime(&ms, t0, t1));
    if (ch.timing) fprintf(stderr, "[s3g timing] all bytes on the host %.2f ms\n", host_ms() - ch.tt0);
    // ---- the archive ----
    std::vector<s3g_chrom> chroms;
    std::vector<uint8_t> names;
    ch.result(res, chroms, names);
    res->device_ms = ms;                                       // first upload to last kernel
    ctx->h_chroms = chroms;
    fill_result(res, chroms);
    if (!res->chroms) { set_error("out of host memory"); return S3G_E_NOMEM; }
    std::vector<uint64_t> name_off(chroms.size() + 1, 0);
    for (size_t c = 0; c < chroms.size(); c++) name_off[c + 1] = name_off[c] + chroms[c].name_len;
    std::string hdr = build_header(names.data(), name_off, chroms, level, note);
    const uint64_t streams_off = 4 + hdr.size() + 1;
    if (streams_off > HDR_RESERVE) { s3g_result_free(res); return S3G_E_CAPACITY; }      // enormous chromosome table: one-shot path
    uint8_t *arc = ctx->h_archive + HDR_RESERVE - streams_off;
    static const uint8_t magic[4] = {0xca, 0x5c, 0xad, 0x1a};      // hpp:907-910
    memcpy(arc, magic, 4);
    memcpy(arc + 4, hdr.data(), hdr.size());
    arc[4 + hdr.size()] = '\n';
    res->archive = arc;
    res->streams_off = streams_off;
    res->streams_size = ch.cst.out_bytes;
    res->archive_size = streams_off + ch.cst.out_bytes;
    res->d_streams = ctx->chain_out.p;
    ctx->last_streams_size = ch.cst.out_bytes;
    ctx->last_streams_host = ctx->h_archive + HDR_RESERVE;
    ctx->archive_hint = ch.cst.out_bytes;
    return S3G_OK;
}

// ---- bounded-memory ingestion ---------------------------------------------------------------
// s3g_stream_begin / _write / _end take the input in pieces of any size.  Bytes collect in a pinned staging buffer of
// `range_bytes`; a full buffer is cut at its last line feed, uploaded behind a copy of the line before it (the halo) and
// goes through one step of the chain above: the bzip2 blocks whose cut is final leave as bits that continue the step
// before, the unfinished tail of the transformed bytes (less than two blocks) waits on the device.  Resident: one range
// on the host, one range plus that tail on the device, and the compressed bytes so far (host) -- whatever the size of a
// chromosome.
struct StreamState {
    std::string note;
    uint64_t range_bytes = 0;
    uint8_t *h_stage = nullptr; uint64_t stage_fill = 0;
    DevBuf dev;
    DevBuf tf[2];                                // the chain's transformed bytes: the tail waits here between calls, whatever else the context does
    std::vector<uint8_t> halo;                   // the last line of the range before
    std::vector<uint8_t> streams;
    Chain chain;
    uint64_t ranges = 0;
    double device_ms = 0;
    ~StreamState() { if (h_stage) cudaFreeHost(h_stage); dev.release(); tf[0].release(); tf[1].release(); }
};
void stream_state_free(Ctx *ctx) { delete static_cast<StreamState *>(ctx->stream_state); ctx->stream_state = nullptr; }

// grow a device buffer keeping its first `keep` bytes
static int grow_keep(Ctx *ctx, DevBuf &b, uint64_t need, uint64_t keep)
{
    if (need <= b.cap) return S3G_OK;
    DevBuf nb;
    S3G_TRY(nb.ensure(need + need / 4));
    if (keep) S3G_CUDA(cudaMemcpyAsync(nb.p, b.p, keep, cudaMemcpyDeviceToDevice, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    b.release();
    b = nb;
    return S3G_OK;
}

// upload stage[0, take) behind the halo line and run one step of the chain
static int stream_flush(Ctx *ctx, StreamState &S, uint64_t take, bool last)
{
    const uint64_t halo = S.halo.size();
    if (take == 0 && !last) return S3G_OK;
    if (take == 0 && S.ranges == 0) return S3G_OK;          // an empty input: no step at all
    S3G_TRY(S.dev.ensure(halo + take + 64));
    uint8_t *d = S.dev.as<uint8_t>();
    if (halo) S3G_CUDA(cudaMemcpyAsync(d, S.halo.data(), halo, cudaMemcpyHostToDevice, ctx->stream));
    if (take) S3G_CUDA(cudaMemcpyAsync(d + halo, S.h_stage, take, cudaMemcpyHostToDevice, ctx->stream));
    S3G_CUDA(cudaMemsetAsync(d + halo + take, 0, 64, ctx->stream));
    S3G_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    S.chain.tf_room = S.range_bytes + (4ull << 20);
    S3G_TRY(S.chain.step(d, halo + take, halo, S.h_stage, 0, last));
    S3G_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));          // the staging buffer is refilled by the caller
    float ms = 0;
    if (cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1) == cudaSuccess) S.device_ms += ms; else cudaGetLastError();
    S.ranges++;
    if (!last && take) {
        // the next range's halo: this range's last line (a range ends with a line feed)
        uint64_t q = take - 1;
        while (q > 0 && S.h_stage[q - 1] != '\n') q--;
        S.halo.assign(S.h_stage + q, S.h_stage + take);
    }
    if (S.stage_fill > take) memmove(S.h_stage, S.h_stage + take, S.stage_fill - take);
    S.stage_fill -= take;
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
    auto setup = [&]() -> int {
        S3G_CUDA(cudaStreamCreateWithFlags(&c->own_stream, cudaStreamNonBlocking));
        c->stream = c->own_stream;
        S3G_CUDA(cudaEventCreate(&c->ev0));
        S3G_CUDA(cudaEventCreate(&c->ev1));
        S3G_CUDA(cudaMallocHost(&c->h_scalars, 64 * 8));
        return S3G_OK;
    };
    int rc = setup();
    if (rc != S3G_OK) { s3g_destroy(c); return rc; }       // s3g_destroy releases whatever was created
    *out = c;
    return S3G_OK;
}

void s3g_destroy(s3g_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    size_t k; DevBuf *const *bl = all_bufs(ctx, &k);
    for (size_t i = 0; i < k; i++) bl[i]->release();
    for (int k2 = 0; k2 < 2; k2++) if (ctx->sub[k2]) { s3g_destroy(static_cast<s3g_ctx *>(ctx->sub[k2])); ctx->sub[k2] = nullptr; }
    for (cudaEvent_t e : ctx->part_ev) cudaEventDestroy(e);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->out_stream) cudaStreamDestroy(ctx->out_stream);
    if (ctx->out_ev) cudaEventDestroy(ctx->out_ev);
    if (ctx->h_scalars) cudaFreeHost(ctx->h_scalars);
    if (ctx->h_archive) cudaFreeHost(ctx->h_archive);
    if (ctx->h_stage) cudaFreeHost(ctx->h_stage);
    for (int k2 = 0; k2 < 2; k2++) if (ctx->h_small[k2]) cudaFreeHost(ctx->h_small[k2]);
    for (cudaEvent_t e : ctx->stage_ev) cudaEventDestroy(e);
    if (ctx->ev0) cudaEventDestroy(ctx->ev0);
    if (ctx->ev1) cudaEventDestroy(ctx->ev1);
    for (cudaEvent_t e : ctx->prof_pool) cudaEventDestroy(e);
    for (cudaEvent_t e : ctx->mark_pool) cudaEventDestroy(e);
    shard_state_free(ctx);
    stream_state_free(ctx);
    if (ctx->own_stream) cudaStreamDestroy(ctx->own_stream);
    delete ctx;
}

int s3g_set_stream(s3g_ctx *ctx, void *cuda_stream)
{
    if (!ctx) { set_error("null context"); return S3G_E_PARAM; }
    ctx->stream = cuda_stream ? (cudaStream_t)cuda_stream : ctx->own_stream;
    return S3G_OK;
}

int s3g_chain_layout(uint64_t *state, int level, uint64_t n_streams, int first_continues, const uint32_t *stream_of, const uint64_t *n_bits,
                     const uint32_t *crc, uint64_t n_blocks, uint64_t n_final, uint64_t *block_pos, uint64_t *patch, uint64_t patch_cap,
                     uint64_t *n_patch, uint64_t *stream_start, uint64_t *stream_len)
{
    if (!state || !n_patch || (n_blocks && (!stream_of || !n_bits || !crc)) || n_final > n_blocks) { set_error("bad argument"); return S3G_E_PARAM; }
    ChainCarry st;
    st.open = state[0] != 0; st.open_bits = state[1]; st.open_comb = (uint32_t)state[2]; st.out_bytes = state[3];
    StepLayout L;
    if (chain_layout_step(st, level, n_streams, first_continues != 0, stream_of, n_bits, crc, n_blocks, n_final, L) != S3G_OK) {
        set_error("block table and stream table disagree"); return S3G_E_PARAM;
    }
    *n_patch = L.patches.size() / 2;
    if (L.patches.size() > 2 * patch_cap) { set_error("patch table too small"); return S3G_E_CAPACITY; }
    for (size_t k = 0; k < L.pos.size(); k++) if (block_pos) block_pos[k] = L.pos[k];
    for (size_t k = 0; k < L.patches.size(); k++) if (patch) patch[k] = L.patches[k];
    for (uint64_t q = 0; q < n_streams; q++) { if (stream_start) stream_start[q] = L.started[q]; if (stream_len) stream_len[q] = L.closed_len[q]; }
    state[0] = st.open; state[1] = st.open_bits; state[2] = st.open_comb; state[3] = st.out_bytes;
    return S3G_OK;
}

uint64_t s3g_launch_count(const s3g_ctx *ctx) { return ctx ? ctx->launches : 0; }
int s3g_last_host_entry(const s3g_ctx *ctx) { return ctx ? ctx->last_host_entry : 0; }

uint64_t s3g_sort_retries(const s3g_ctx *ctx)
{
    if (!ctx) return 0;
    uint64_t r = ctx->sort_retries;
    for (int k = 0; k < 2; k++) if (ctx->sub[k]) r += ctx->sub[k]->sort_retries;
    return r;
}

int s3g_sort_stats(const s3g_ctx *ctx, uint64_t out[3])
{
    if (!ctx || !out) { set_error("null argument"); return S3G_E_PARAM; }
    out[0] = ctx->bucket_blocks; out[1] = ctx->bucket_handed_back; out[2] = ctx->sort_retries;
    for (int k = 0; k < 2; k++)
        if (ctx->sub[k]) { out[0] += ctx->sub[k]->bucket_blocks; out[1] += ctx->sub[k]->bucket_handed_back; out[2] += ctx->sub[k]->sort_retries; }
    return S3G_OK;
}

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
    uint64_t pipe_min = PIPE_MIN_BYTES;
    if (const char *e = getenv("S3G_PIPE_MIN")) { long long v = atoll(e); if (v > 0) pipe_min = (uint64_t)v; }
    int nparts = n >= pipe_min ? 3 : 1;
    if (const char *e = getenv("S3G_PARTS")) { int v = atoi(e); if (v >= 1 && v <= 64) nparts = v; }
    if (nparts > 1 && !ctx->prof) {
        // Few chromosome changes (one long chromosome, BASELINE.json configs 1 and 3): ranges of whole chromosomes would not
        // overlap anything, so the ranges are chained at block granularity instead.  Sampled at eight line starts.
        bool chain = false;
        if (const char *e = getenv("S3G_CHAIN")) chain = atoi(e) != 0;
        else {
            int changes = 0;
            std::string prev;
            for (int k = 0; k < 8; k++) {
                uint64_t p = (uint64_t)((__uint128_t)n * k / 8);
                if (k) { const void *nl = memchr(bed + p, '\n', n - p); if (!nl) break; p = (uint64_t)((const uint8_t *)nl - bed) + 1; }
                uint64_t q = p;
                while (q < n && q - p < 256 && bed[q] != '\t' && bed[q] != '\n') q++;
                std::string name(reinterpret_cast<const char *>(bed + p), q - p);
                if (k && name != prev) changes++;
                prev = name;
            }
            chain = changes < 4;
        }
        int rc;
        if (chain) {
            // Range size: while a range is on its way every launch and read-back of a step is slower (the GPU fetches its commands
            // over the PCIe direction the upload saturates: a step of 96 MiB takes 2.9 ms beside an upload and 1.65 ms after
            // it), so a step costs about 2 ms + 10 us per MiB and keeps up with the 19 us per MiB of the upload from about
            // 224 MiB; larger ranges only lengthen the last step, which nothing hides (measured: 96 MiB 64.5 ms, 160 MiB
            // 55.3 ms, 224 and 320 MiB 53.1 ms on 2.6 GB).
            uint64_t range = 256ull << 20;
            if (const char *e = getenv("S3G_CHAIN_BYTES")) { long long v = atoll(e); if (v > 0) range = (uint64_t)v; }
            const int steps = (int)std::min<uint64_t>(256, std::max<uint64_t>(2, (n + range - 1) / range));
            rc = compress_bed_chained(ctx, bed, n, level, note, steps, res);
            ctx->last_host_entry = -steps;
        } else {
            rc = compress_bed_pipelined(ctx, bed, n, level, note, nparts, res);
            ctx->last_host_entry = nparts;
        }
        if (rc != S3G_E_CAPACITY) return rc;
        s3g_result_free(res);                                  // does not fit the pipelined buffers: one piece
    }
    ctx->last_host_entry = 0;
    S3G_TRY(stage_in(ctx, ctx->bed, bed, n));
    return compress_bed_impl(ctx, ctx->bed.as<uint8_t>(), n, level, note, 1, res);
}

int s3g_stream_begin(s3g_ctx *ctx, int level, const char *note, uint64_t range_bytes)
{
    if (!ctx) { set_error("null context"); return S3G_E_PARAM; }
    if (level < 1 || level > 9) { set_error("block_size_100k must be 1..9"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    stream_state_free(ctx);
    StreamState *S = new (std::nothrow) StreamState();
    if (!S) { set_error("out of host memory"); return S3G_E_NOMEM; }
    S->note = note ? note : "";
    S->range_bytes = range_bytes ? std::max<uint64_t>(range_bytes, 4096) : (256ull << 20);
    if (cudaMallocHost(&S->h_stage, S->range_bytes) != cudaSuccess) { cudaGetLastError(); delete S; set_error("out of pinned host memory"); return S3G_E_NOMEM; }
    S->chain.ctx = ctx; S->chain.level = level; S->chain.v_out = &S->streams;
    S->chain.tf[0] = &S->tf[0]; S->chain.tf[1] = &S->tf[1];
    ctx->stream_state = S;
    return S3G_OK;
}

int s3g_stream_write(s3g_ctx *ctx, const uint8_t *bed, uint64_t n)
{
    if (!ctx || (!bed && n)) { set_error("null argument"); return S3G_E_PARAM; }
    StreamState *S = static_cast<StreamState *>(ctx->stream_state);
    if (!S) { set_error("s3g_stream_write without s3g_stream_begin"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    while (n) {
        const uint64_t room = S->range_bytes - S->stage_fill, k = std::min(room, n);
        memcpy(S->h_stage + S->stage_fill, bed, k);
        S->stage_fill += k; bed += k; n -= k;
        if (S->stage_fill == S->range_bytes) {
            // cut at the last line feed
            uint64_t take = S->stage_fill;
            while (take && S->h_stage[take - 1] != '\n') take--;
            if (take) { int rc = stream_flush(ctx, *S, take, false); if (rc != S3G_OK) { stream_state_free(ctx); return rc; } }
            else {
                // no line ends in this range (a line longer than the range): the range grows to hold it
                uint8_t *p = nullptr;
                if (cudaMallocHost(&p, S->range_bytes * 2) != cudaSuccess) { cudaGetLastError(); stream_state_free(ctx); set_error("out of pinned host memory"); return S3G_E_NOMEM; }
                memcpy(p, S->h_stage, S->stage_fill);
                cudaFreeHost(S->h_stage);
                S->h_stage = p; S->range_bytes *= 2;
            }
        }
    }
    return S3G_OK;
}

int s3g_stream_end(s3g_ctx *ctx, s3g_result *res)
{
    if (!ctx || !res) { set_error("null argument"); return S3G_E_PARAM; }
    StreamState *S = static_cast<StreamState *>(ctx->stream_state);
    if (!S) { set_error("s3g_stream_end without s3g_stream_begin"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    memset(res, 0, sizeof *res);
    int rc = stream_flush(ctx, *S, S->stage_fill, true);
    if (rc == S3G_OK && S->chain.cst.open) { set_error("stream entry: a stream was left open"); rc = S3G_E_CUDA; }
    if (rc != S3G_OK) { stream_state_free(ctx); return rc; }
    std::vector<s3g_chrom> chroms;
    std::vector<uint8_t> names;
    S->chain.result(res, chroms, names);
    if (S->ranges == 0) res->dropped_tail_bytes = S->stage_fill;          // nothing but an unterminated fragment (or nothing at all)
    for (s3g_chrom &c : chroms) c.name_off = 0;                           // no input buffer to point into
    std::vector<uint64_t> name_off(chroms.size() + 1, 0);
    for (size_t c = 0; c < chroms.size(); c++) name_off[c + 1] = name_off[c] + chroms[c].name_len;
    std::string hdr = build_header(names.data(), name_off, chroms, S->chain.level, S->note.c_str());
    const uint64_t streams_off = 4 + hdr.size() + 1;
    rc = ensure_archive(ctx, streams_off + S->streams.size());
    if (rc != S3G_OK) { stream_state_free(ctx); return rc; }
    static const uint8_t magic[4] = {0xca, 0x5c, 0xad, 0x1a};      // hpp:907-910
    memcpy(ctx->h_archive, magic, 4);
    memcpy(ctx->h_archive + 4, hdr.data(), hdr.size());
    ctx->h_archive[4 + hdr.size()] = '\n';
    if (!S->streams.empty()) memcpy(ctx->h_archive + streams_off, S->streams.data(), S->streams.size());
    res->archive = ctx->h_archive; res->archive_size = streams_off + S->streams.size(); res->streams_off = streams_off;
    res->streams_size = S->streams.size(); res->d_streams = nullptr;
    res->device_ms = S->device_ms;
    ctx->h_chroms = chroms;
    fill_result(res, chroms);
    ctx->last_streams_size = S->streams.size();
    ctx->last_streams_host = ctx->h_archive + streams_off;
    stream_state_free(ctx);
    if (!res->chroms) { set_error("out of host memory"); return S3G_E_NOMEM; }
    return S3G_OK;
}

int s3g_read_streams(s3g_ctx *ctx, uint8_t *dst, uint64_t cap, uint64_t *n)
{
    if (!ctx || !dst || !n) { set_error("null argument"); return S3G_E_PARAM; }
    *n = ctx->last_streams_size;
    if (*n > cap) { set_error("cap too small: need %llu", (unsigned long long)*n); return S3G_E_CAPACITY; }
    if (ctx->last_streams_host) {          // pipelined host entry: the streams went straight into the pinned archive
        memcpy(dst, ctx->last_streams_host, *n);
        return S3G_OK;
    }
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
            S3G_CUDA(cudaMemcpyAsync(rle_out + off, ctx->blk_bytes.as<uint8_t>() + B.blk_off, B.nblock, cudaMemcpyDeviceToHost, ctx->stream));
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
    const uint8_t *p = blk + blocks[b].blk_off;
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
    if (nb > 65535) { set_error("at most 65535 blocks per stage call"); return S3G_E_LIMIT; }
    S3G_TRY(ctx->in_use.ensure(nb * 256));
    S3G_TRY(ctx->seq_map.ensure(nb * 256));
    uint64_t packed = 0;
    for (uint64_t b = 0; b < nb; b++) {
        uint64_t len = off[b + 1] - off[b];
        if (len == 0 || len > 900000) { set_error("block %llu has invalid length %llu", (unsigned long long)b, (unsigned long long)len); return S3G_E_PARAM; }
        BlockInfo &B = ctx->h_blocks[b];
        memset(&B, 0, sizeof B);
        B.nblock = (uint32_t)len; B.orig_ptr = -1; B.blk_off = packed;
        packed += blk_slot_bytes(B.nblock);
    }
    S3G_TRY(ctx->blk_bytes.ensure(packed + 256));
    for (uint64_t b = 0; b < nb; b++)
        S3G_CUDA(cudaMemcpyAsync(ctx->blk_bytes.as<uint8_t>() + ctx->h_blocks[b].blk_off, blocks + off[b], off[b + 1] - off[b], cudaMemcpyHostToDevice, ctx->stream));
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
