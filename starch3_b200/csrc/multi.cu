// multi.cu -- one archive from several GPUs, driven by ONE C++ process.
//
// s3g_multi_compress_bed runs the phases of shard.cu (s3g_shard_*) on a context per GPU, a host thread each, and does
// the small exchanges between them in host memory.  The two bulk exchanges need no copies at all: with peer access
// enabled, a pointer into another GPU's memory is valid on this GPU, so
//   * the transform kernel stores its output into EVERY GPU's copy of the transformed buffer (s3g_shard_transform_peers),
//   * every GPU stores its finished byte string into the first GPU's gather buffer (s3g_shard_place),
// over NVLink.  The archive is the same bytes as the single-GPU archive (tests/test_gpu_multi.py).
// The Python orchestration of the same phases, one process per GPU with NCCL for the exchanges, is
// starch3_b200/multigpu.py -- the form bench.py launches; the host logic here is the same, statement for statement.
#include <algorithm>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include "common.cuh"

namespace s3g {

namespace {

// a barrier for the worker threads that any of them can break (a failed phase must not leave the others waiting)
struct Barrier {
    std::mutex mu;
    std::condition_variable cv;
    int n, waiting = 0, generation = 0;
    bool broken = false;
    explicit Barrier(int count) : n(count) {}
    bool wait()                                   // false: somebody failed
    {
        std::unique_lock<std::mutex> lk(mu);
        if (broken) return false;
        const int gen = generation;
        if (++waiting == n) { waiting = 0; generation++; cv.notify_all(); return !broken; }
        cv.wait(lk, [&] { return generation != gen || broken; });
        return !broken;
    }
    void abort() { std::unique_lock<std::mutex> lk(mu); broken = true; cv.notify_all(); }
};

struct Piece { std::string name; uint64_t tf_len; int64_t lines, nonuniq, uniq; };
struct Stream { std::string name; uint64_t tf_off, tf_len; int64_t lines, nonuniq, uniq; };

// cut points at line starts and, per range, the length of the line before it (multigpu.py plan_ranges)
void plan_ranges(const uint8_t *bed, uint64_t n, int world, std::vector<uint64_t> &cut, std::vector<uint64_t> &halo)
{
    cut.assign(1, 0);
    for (int r = 1; r < world; r++) {
        uint64_t p = std::max<uint64_t>(cut.back(), (uint64_t)((__uint128_t)n * r / world));
        if (p > 0 && p < n && bed[p - 1] != '\n') {
            const void *nl = memchr(bed + p, '\n', n - p);
            p = nl ? (uint64_t)((const uint8_t *)nl - bed) + 1 : n;
        }
        cut.push_back(std::min(p, n));
    }
    cut.push_back(n);
    halo.assign(world, 0);
    for (int r = 0; r < world; r++) {
        const uint64_t c = cut[r];
        if (c == 0) continue;
        uint64_t q = c - 1;                         // bed[c - 1] is the line feed that ends the halo line
        while (q > 0 && bed[q - 1] != '\n') q--;
        halo[r] = c - q;
    }
}

struct Shared {
    int world;
    Barrier bar;
    std::vector<s3g_shard_summary> summ;
    std::vector<std::vector<s3g_chrom>> pieces;     // per rank, as returned (name_off relative to the rank's upload)
    std::vector<uint64_t> tf_off;                   // where rank r's transformed bytes start; [world] = total
    std::vector<int64_t> carry;
    std::vector<uint64_t> tf_buf;                   // device address of rank r's copy of the transformed buffer
    uint64_t gather_buf = 0;                        // on rank 0's device
    std::vector<Stream> streams;
    std::vector<uint64_t> soff;
    std::vector<uint32_t> nblock, stream_of;
    std::vector<uint64_t> bounds;
    std::vector<uint64_t> n_bits; std::vector<uint32_t> crc; std::vector<uint32_t> n_mtf;
    std::vector<uint64_t> stream_off, stream_len;
    uint64_t total = 0;
    std::mutex err_mu; int rc = S3G_OK; std::string err;
    explicit Shared(int w) : world(w), bar(w), summ(w), pieces(w), tf_off(w + 1, 0), carry(w, INT64_MIN), tf_buf(w, 0) {}
    void fail(int code)
    {
        { std::unique_lock<std::mutex> lk(err_mu); if (rc == S3G_OK) { rc = code; err = s3g_last_error(); } }
        bar.abort();
    }
};

// largest stop of all earlier lines of the chromosome a range continues (multigpu.py carry_chain)
void carry_chain(Shared &S)
{
    int64_t run = INT64_MIN;
    for (int r = 0; r < S.world; r++) {
        const s3g_shard_summary &m = S.summ[r];
        const int64_t c = m.continues ? run : INT64_MIN;
        S.carry[r] = c;
        if (m.n_lines == 0) continue;
        run = (m.single_piece && m.continues) ? std::max(m.tail_max, c) : m.tail_max;
    }
}

void block_shares(Shared &S)
{
    const uint64_t nb = S.nblock.size();
    const int world = S.world;
    std::vector<uint64_t> cum(nb + 1, 0);
    for (uint64_t b = 0; b < nb; b++) cum[b + 1] = cum[b] + S.nblock[b];
    const uint64_t total = cum[nb], cap = (nb + world - 1) / world;
    S.bounds.assign(1, 0);
    for (int r = 1; r < world; r++) {
        const uint64_t want = (uint64_t)(((__uint128_t)total * r + world / 2) / world);
        uint64_t b = (uint64_t)(std::lower_bound(cum.begin(), cum.end(), want) - cum.begin());
        b = std::min(b, S.bounds.back() + cap);
        const uint64_t need = (uint64_t)(world - r) * cap;
        if (nb > need) b = std::max(b, nb - need);
        S.bounds.push_back(std::min(std::max(b, S.bounds.back()), nb));
    }
    S.bounds.push_back(nb);
}

void worker(int r, Shared &S, s3g_ctx **ctxs, const uint8_t *bed, const std::vector<uint64_t> &cut, const std::vector<uint64_t> &halo, int level)
{
    s3g_ctx *ctx = ctxs[r];
    const int world = S.world;
#define MG_TRY(call) do { int rc_ = (call); if (rc_ != S3G_OK) { S.fail(rc_); return; } } while (0)
#define MG_CUDA(call) do { if ((call) != cudaSuccess) { set_error("%s failed: %s", #call, cudaGetErrorString(cudaGetLastError())); S.fail(S3G_E_CUDA); return; } } while (0)
#define MG_SYNC() do { if (!S.bar.wait()) return; } while (0)
    MG_CUDA(cudaSetDevice(ctx->device));
    // ---- phase 1: upload the range with its halo line, tokenizer ----
    const uint64_t lo = cut[r] - halo[r], hi = cut[r + 1], len = hi - lo;
    if (ctx->bed.ensure(len + 64) != S3G_OK) { S.fail(S3G_E_NOMEM); return; }
    if (len) MG_CUDA(cudaMemcpyAsync(ctx->bed.p, bed + lo, len, cudaMemcpyHostToDevice, ctx->stream));
    MG_CUDA(cudaMemsetAsync(ctx->bed.as<uint8_t>() + len, 0, 64, ctx->stream));
    MG_TRY(s3g_shard_tokenize(ctx, ctx->bed.p, len, halo[r], &S.summ[r]));
    MG_SYNC();
    if (r == 0) {
        carry_chain(S);
        for (int k = 0; k < world; k++) S.tf_off[k + 1] = S.tf_off[k] + S.summ[k].tf_bytes;
    }
    MG_SYNC();
    // ---- phase 2: every GPU holds a copy of the transformed buffer; the transform stores into all of them ----
    const uint64_t tf_total = S.tf_off[world];
    if (ctx->io_a.ensure(tf_total + 256) != S3G_OK) { S.fail(S3G_E_NOMEM); return; }
    S.tf_buf[r] = (uint64_t)(uintptr_t)ctx->io_a.p;
    if (r == 0) {
        if (ctx->io_b.ensure(tf_total + tf_total / 64 + (1u << 20)) != S3G_OK) { S.fail(S3G_E_NOMEM); return; }
        S.gather_buf = (uint64_t)(uintptr_t)ctx->io_b.p;
        MG_CUDA(cudaMemsetAsync(ctx->io_b.p, 0, ctx->io_b.cap, ctx->stream));
        MG_CUDA(cudaStreamSynchronize(ctx->stream));
    }
    MG_SYNC();
    {
        std::vector<uint64_t> bufs;
        bufs.push_back(S.tf_buf[r]);
        for (int k = 0; k < world; k++) if (k != r) bufs.push_back(S.tf_buf[k]);
        std::vector<s3g_chrom> pc(4096);
        uint64_t np = 0, tl = 0;
        for (;;) {
            int rc = s3g_shard_transform_peers(ctx, S.carry[r], pc.data(), pc.size(), &np, bufs.data(), (uint32_t)bufs.size(), 0, S.tf_off[r], &tl);
            if (rc == S3G_E_CAPACITY && pc.size() < (1u << 24)) { pc.resize(pc.size() * 16); continue; }
            MG_TRY(rc);
            break;
        }
        if (tl != S.summ[r].tf_bytes) { set_error("transformed size differs from the measured one"); S.fail(S3G_E_CUDA); return; }
        pc.resize(np);
        S.pieces[r] = pc;
    }
    MG_SYNC();                                     // every copy of the transformed buffer is complete
    if (r == 0) {
        // a range's first piece that continues the previous range's last chromosome is the same stream (merge_pieces)
        uint64_t off = 0;
        for (int k = 0; k < world; k++) {
            const uint64_t base = cut[k] - halo[k];
            for (size_t q = 0; q < S.pieces[k].size(); q++) {
                const s3g_chrom &p = S.pieces[k][q];
                if (q == 0 && S.summ[k].continues && !S.streams.empty()) {
                    Stream &s = S.streams.back();
                    s.tf_len += p.tf_len; s.lines += p.line_count; s.nonuniq += p.bases_nonunique; s.uniq += p.bases_unique;
                } else {
                    Stream s;
                    s.name.assign(reinterpret_cast<const char *>(bed + base + p.name_off), p.name_len);
                    s.tf_off = off; s.tf_len = p.tf_len; s.lines = p.line_count; s.nonuniq = p.bases_nonunique; s.uniq = p.bases_unique;
                    S.streams.push_back(s);
                }
                off += p.tf_len;
            }
        }
        S.soff.clear();
        for (const Stream &s : S.streams) S.soff.push_back(s.tf_off);
        S.soff.push_back(tf_total);
    }
    MG_SYNC();
    // ---- phase 3: the block plan, the same on every GPU ----
    const uint64_t n_streams = S.streams.size();
    {
        const uint64_t cap = tf_total / (100000ull * level - 19 - 260) + n_streams + 8;
        std::vector<uint32_t> nbk(cap), sof(cap);
        uint64_t nb = 0;
        MG_TRY(s3g_shard_plan(ctx, ctx->io_a.p, tf_total, S.soff.data(), n_streams, level, &nb, nbk.data(), sof.data(), cap));
        if (r == 0) {
            nbk.resize(nb); sof.resize(nb);
            S.nblock = nbk; S.stream_of = sof;
            block_shares(S);
            S.n_bits.assign(nb, 0); S.crc.assign(nb, 0); S.n_mtf.assign(nb, 0);
        }
    }
    MG_SYNC();
    // ---- phase 4: the GPU's share of the blocks ----
    const uint64_t b_lo = S.bounds[r], b_hi = S.bounds[r + 1];
    {
        std::vector<uint64_t> nbits(std::max<uint64_t>(1, b_hi - b_lo));
        std::vector<uint32_t> crc(nbits.size()), nmtf(nbits.size());
        MG_TRY(s3g_shard_compress(ctx, b_lo, b_hi, nbits.data(), crc.data(), nmtf.data()));
        for (uint64_t b = b_lo; b < b_hi; b++) { S.n_bits[b] = nbits[b - b_lo]; S.crc[b] = crc[b - b_lo]; S.n_mtf[b] = nmtf[b - b_lo]; }   // disjoint ranges
    }
    MG_SYNC();
    // ---- phase 5: place the share; every GPU stores its string into the first GPU's gather buffer ----
    {
        void *d_bytes = nullptr;
        uint64_t blo = 0, bhi = 0;
        std::vector<uint64_t> so(std::max<uint64_t>(1, n_streams)), sl(so.size());
        uint64_t no_bits = 0; uint32_t no_crc = 0;           // an input without a single block: the tables are empty
        MG_TRY(s3g_shard_assemble(ctx, S.n_bits.empty() ? &no_bits : S.n_bits.data(), S.crc.empty() ? &no_crc : S.crc.data(), b_lo, b_hi, &d_bytes, &blo, &bhi,
                                  so.data(), sl.data()));
        if (r == 0) {
            S.stream_off.assign(so.begin(), so.begin() + n_streams); S.stream_len.assign(sl.begin(), sl.begin() + n_streams);
            S.total = n_streams ? so[n_streams - 1] + sl[n_streams - 1] : 0;
        }
        MG_TRY(s3g_shard_place(ctx, S.gather_buf, blo, bhi));
    }
    MG_SYNC();
#undef MG_TRY
#undef MG_CUDA
#undef MG_SYNC
}

void json_escape(std::string &o, const std::string &s)
{
    o.push_back('"');
    for (unsigned char c : s) {
        switch (c) {
            case '\\': o += "\\\\"; break; case '"': o += "\\\""; break; case '\b': o += "\\b"; break; case '\f': o += "\\f"; break;
            case '\n': o += "\\n"; break; case '\r': o += "\\r"; break; case '\t': o += "\\t"; break;
            default:
                if (c < 0x20) { char b[8]; snprintf(b, sizeof b, "\\u%04X", (unsigned)c); o += b; }
                else o.push_back((char)c);
        }
    }
    o.push_back('"');
}

}  // namespace
}  // namespace s3g

using namespace s3g;

extern "C" int s3g_multi_compress_bed(s3g_ctx **ctxs, int n_ctx, const uint8_t *bed, uint64_t n, int level, const char *note, s3g_result *res)
{
    if (!ctxs || n_ctx < 1 || n_ctx > 8 || !res || (!bed && n)) { set_error("bad argument (1..8 contexts)"); return S3G_E_PARAM; }
    if (level < 1 || level > 9) { set_error("block_size_100k must be 1..9"); return S3G_E_PARAM; }
    for (int k = 0; k < n_ctx; k++) if (!ctxs[k]) { set_error("null context"); return S3G_E_PARAM; }
    memset(res, 0, sizeof *res);
    // every GPU reaches every other GPU's memory through its own address space (NVLink / NVSwitch)
    for (int a = 0; a < n_ctx; a++)
        for (int b = 0; b < n_ctx; b++) {
            if (ctxs[a]->device == ctxs[b]->device) continue;
            S3G_CUDA(cudaSetDevice(ctxs[a]->device));
            int can = 0;
            S3G_CUDA(cudaDeviceCanAccessPeer(&can, ctxs[a]->device, ctxs[b]->device));
            if (!can) { set_error("device %d cannot access device %d's memory", ctxs[a]->device, ctxs[b]->device); return S3G_E_CUDA; }
            cudaError_t e = cudaDeviceEnablePeerAccess(ctxs[b]->device, 0);
            if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) { set_error("cudaDeviceEnablePeerAccess: %s", cudaGetErrorString(e)); cudaGetLastError(); return S3G_E_CUDA; }
            cudaGetLastError();
        }
    std::vector<uint64_t> cut, halo;
    plan_ranges(bed, n, n_ctx, cut, halo);
    Shared S(n_ctx);
    cudaEvent_t t0 = ctxs[0]->ev0, t1 = ctxs[0]->ev1;
    S3G_CUDA(cudaSetDevice(ctxs[0]->device));
    S3G_CUDA(cudaEventRecord(t0, ctxs[0]->stream));
    std::vector<std::thread> th;
    for (int r = 0; r < n_ctx; r++) th.emplace_back(worker, r, std::ref(S), ctxs, bed, std::cref(cut), std::cref(halo), level);
    for (std::thread &t : th) t.join();
    if (S.rc != S3G_OK) { set_error("%s", S.err.c_str()); return S.rc; }
    Ctx *c0 = ctxs[0];
    S3G_CUDA(cudaSetDevice(c0->device));
    S3G_CUDA(cudaEventRecord(t1, c0->stream));
    // ---- the archive: magic, metadata, the gathered streams (ARCHIVE_FORMAT.md) ----
    const uint64_t n_streams = S.streams.size();
    std::vector<uint64_t> blocks_of(n_streams, 0);
    for (uint32_t s : S.stream_of) blocks_of[s]++;
    std::string hdr;
    hdr += "{\"archive\":{\"type\":\"starch\",\"version\":{\"major\":3,\"minor\":0,\"revision\":0},\"creator\":\"starch3_b200\","
           "\"compression\":\"bzip2\",\"blockSize100k\":" + std::to_string(level) + ",\"note\":";
    json_escape(hdr, note ? note : "");
    hdr += "},\"streams\":[";
    std::vector<s3g_chrom> chroms(n_streams);
    for (uint64_t i = 0; i < n_streams; i++) {
        const Stream &s = S.streams[i];
        if (i) hdr.push_back(',');
        hdr += "{\"chromosome\":";
        json_escape(hdr, s.name);
        hdr += ",\"offset\":" + std::to_string(S.stream_off[i]) + ",\"size\":" + std::to_string(S.stream_len[i]) + ",\"lines\":" + std::to_string(s.lines) +
               ",\"blocks\":" + std::to_string(blocks_of[i]) + ",\"transformedBytes\":" + std::to_string(s.tf_len) + ",\"nonUniqueBases\":" +
               std::to_string(s.nonuniq) + ",\"uniqueBases\":" + std::to_string(s.uniq) + "}";
        s3g_chrom &c = chroms[i];
        memset(&c, 0, sizeof c);
        c.name_len = (uint32_t)s.name.size(); c.n_blocks = (uint32_t)blocks_of[i]; c.tf_off = s.tf_off; c.tf_len = s.tf_len;
        c.line_count = s.lines; c.bases_nonunique = s.nonuniq; c.bases_unique = s.uniq; c.bz_off = S.stream_off[i]; c.bz_len = S.stream_len[i];
    }
    hdr += "]}";
    const uint64_t streams_off = 4 + hdr.size() + 1;
    if (c0->h_archive_cap < streams_off + S.total) {
        uint8_t *p = nullptr;
        const size_t want = streams_off + S.total + (streams_off + S.total) / 8 + 4096;
        if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); set_error("out of pinned host memory"); return S3G_E_NOMEM; }
        if (c0->h_archive) cudaFreeHost(c0->h_archive);
        c0->h_archive = p; c0->h_archive_cap = want;
    }
    static const uint8_t magic[4] = {0xca, 0x5c, 0xad, 0x1a};      // hpp:907-910
    memcpy(c0->h_archive, magic, 4);
    memcpy(c0->h_archive + 4, hdr.data(), hdr.size());
    c0->h_archive[4 + hdr.size()] = '\n';
    if (S.total) S3G_CUDA(cudaMemcpyAsync(c0->h_archive + streams_off, reinterpret_cast<const void *>((uintptr_t)S.gather_buf), S.total, cudaMemcpyDeviceToHost, c0->stream));
    S3G_CUDA(cudaStreamSynchronize(c0->stream));
    float ms = 0;
    if (cudaEventElapsedTime(&ms, t0, t1) == cudaSuccess) res->device_ms = ms; else cudaGetLastError();
    res->archive = c0->h_archive; res->archive_size = streams_off + S.total; res->streams_off = streams_off; res->streams_size = S.total;
    res->d_streams = reinterpret_cast<void *>((uintptr_t)S.gather_buf);
    res->n_chroms = n_streams; res->n_blocks = S.nblock.size();
    for (int r = 0; r < n_ctx; r++) res->n_lines += S.summ[r].n_lines;
    res->dropped_tail_bytes = S.summ[n_ctx - 1].dropped_tail_bytes;
    for (uint32_t v : S.nblock) res->rle_bytes += v;
    for (uint32_t v : S.n_mtf) res->mtf_symbols += v;
    res->tf_bytes = S.tf_off[n_ctx];
    res->chroms = (s3g_chrom *)malloc(std::max<size_t>(1, n_streams) * sizeof(s3g_chrom));
    if (!res->chroms) { set_error("out of host memory"); return S3G_E_NOMEM; }
    if (n_streams) memcpy(res->chroms, chroms.data(), n_streams * sizeof(s3g_chrom));
    c0->h_chroms = chroms;
    c0->last_streams_size = S.total;
    c0->last_streams_host = c0->h_archive + streams_off;
    return S3G_OK;
}
