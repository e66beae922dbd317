// shard.cu -- one archive from several GPUs: the hot path cut into phases that a host
// runs on every GPU with small exchanges in between.  The host is either one process per GPU
// (starch3_b200/multigpu.py: torch.distributed over NCCL/NVLink for the exchanges) or one process with a context per
// device (s3g_multi_compress_bed, multi.cu: stores over NVLink through peer access).
//
//   phase               per GPU                                                    exchanged afterwards
//   s3g_shard_tokenize  line framing + tokenizer of the GPU's newline-aligned      summary: largest stop since the last
//                       byte range (plus the line before it, the "halo")           chromosome change (for uniqueBases)
//   s3g_shard_transform transform + statistics, given the carried maximum          the transformed bytes (all-gather over
//                                                                                  NVLink) and the piece table
//   s3g_shard_plan      RLE1 lengths + block cut over ALL transformed bytes        -- (every GPU computes the same plan)
//   s3g_shard_compress  RLE1 bytes, CRC, block sort, MTF, Huffman of the GPU's     bit length + CRC of every block
//                       share of the blocks
//   s3g_shard_assemble  bit-level placement of the share (bz/compress.c:609 keeps   the byte strings, gathered by the host
//                       bsBuff/bsLive across blocks) + stream headers/trailers      in archive order (seam bytes ORed)
//
// Why the hand-over is this small: update_transformation_state (hpp:428-504) reads only the previous line (stop,
// length, chromosome), a transformed line never starts with '\n' while every piece ends with one, so no RLE1 run
// crosses a piece boundary (bz/bzlib.c:269-293), and a bzip2 block depends only on its own bytes; the block cut
// (bz/bzlib.c:307, :399) is a sequential chain over cheap per-tile prefix sums, which every GPU simply repeats.
#include <algorithm>
#include "common.cuh"

namespace s3g {

struct ShardState {
    const uint8_t *base = nullptr;      // range rounded down to 16 bytes
    uint64_t n = 0;                     // bytes from base
    uint32_t skip = 0, halo = 0;
    TfResult tr;
    uint64_t n_streams = 0;
    int level = 9;
};
static ShardState &st(Ctx *ctx)
{
    if (!ctx->shard) ctx->shard = new ShardState();
    return *static_cast<ShardState *>(ctx->shard);
}
void shard_state_free(Ctx *ctx) { delete static_cast<ShardState *>(ctx->shard); ctx->shard = nullptr; }

}  // namespace s3g

using namespace s3g;

extern "C" {

int s3g_shard_tokenize(s3g_ctx *ctx, const void *d_range, uint64_t n, uint64_t halo_bytes, s3g_shard_summary *out)
{
    if (!ctx || !out || (!d_range && n) || halo_bytes > n) { set_error("bad argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    memset(out, 0, sizeof *out);
    ShardState &S = st(ctx);
    const uintptr_t p = (uintptr_t)d_range;
    S.skip = (uint32_t)(p & 15);
    S.base = reinterpret_cast<const uint8_t *>(p - S.skip);
    S.n = n + S.skip;
    S.halo = halo_bytes ? 1u : 0u;
    ctx->marks.clear();
    ctx->last_rle_bytes = 0;
    stage_mark(ctx, 0);
    S3G_TRY(run_tokenize(ctx, S.base, S.n, S.skip, &S.tr, S.halo));
    uint64_t last_flag = 0; uint32_t cont = 0; int64_t tmax = INT64_MIN;
    S3G_TRY(run_range_summary(ctx, S.tr.n_lines, S.halo, &tmax, &last_flag, &cont));
    stage_mark(ctx, -1);
    if (S.halo && S.tr.n_lines == 0) { set_error("halo line is not newline-terminated"); return S3G_E_PARAM; }
    out->n_lines = S.tr.n_lines - (S.halo ? 1 : 0);
    out->tail_max = tmax;
    out->continues = cont;
    out->single_piece = last_flag == 0 ? 1u : 0u;
    out->dropped_tail_bytes = S.tr.dropped;
    out->tf_bytes = S.tr.tf_len;
    return S3G_OK;
}

int s3g_shard_transform(s3g_ctx *ctx, int64_t carry_max, s3g_chrom *pieces, uint64_t cap, uint64_t *n_pieces, void **d_tf, uint64_t *tf_len)
{
    if (!ctx || !n_pieces || !d_tf || !tf_len) { set_error("null argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    ShardState &S = st(ctx);
    *n_pieces = 0; *d_tf = nullptr; *tf_len = 0;
    if (S.tr.n_lines == 0) return S3G_OK;
    stage_mark(ctx, 0);
    S3G_TRY(run_transform_rest(ctx, S.base, S.n, &S.tr, S.halo, carry_max));
    ctx->h_chroms.resize(S.tr.n_chroms);
    S3G_CUDA(cudaMemcpyAsync(ctx->h_chroms.data(), ctx->chroms.p, S.tr.n_chroms * sizeof(s3g_chrom), cudaMemcpyDeviceToHost, ctx->stream));
    stage_mark(ctx, -1);
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t k = 0;
    for (uint64_t c = 0; c < S.tr.n_chroms; c++) {
        s3g_chrom ch = ctx->h_chroms[c];
        if (S.halo && c == 0 && ch.line_count == 0) continue;      // the halo line's chromosome ended with it
        if (k >= cap) { set_error("piece table too small"); return S3G_E_CAPACITY; }
        ch.name_off -= S.skip;
        if (pieces) pieces[k] = ch;
        k++;
    }
    *n_pieces = k; *d_tf = ctx->tf.p; *tf_len = S.tr.tf_len;
    return S3G_OK;
}

int s3g_shard_transform_peers(s3g_ctx *ctx, int64_t carry_max, s3g_chrom *pieces, uint64_t cap, uint64_t *n_pieces,
                              const uint64_t *peer_bufs, uint32_t n_peers, uint64_t multicast_buf, uint64_t dst_off, uint64_t *tf_len)
{
    if (!ctx || !n_pieces || !tf_len || !peer_bufs || n_peers == 0) { set_error("null argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    ShardState &S = st(ctx);
    *n_pieces = 0; *tf_len = 0;
    if (S.tr.n_lines == 0) return S3G_OK;
    stage_mark(ctx, 0);
    S3G_TRY(run_transform_rest(ctx, S.base, S.n, &S.tr, S.halo, carry_max, false, true, peer_bufs, n_peers, dst_off, multicast_buf));
    ctx->h_chroms.resize(S.tr.n_chroms);
    S3G_CUDA(cudaMemcpyAsync(ctx->h_chroms.data(), ctx->chroms.p, S.tr.n_chroms * sizeof(s3g_chrom), cudaMemcpyDeviceToHost, ctx->stream));
    stage_mark(ctx, -1);
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t k = 0;
    for (uint64_t c = 0; c < S.tr.n_chroms; c++) {
        s3g_chrom ch = ctx->h_chroms[c];
        if (S.halo && c == 0 && ch.line_count == 0) continue;      // the halo line's chromosome ended with it
        if (k >= cap) { set_error("piece table too small"); return S3G_E_CAPACITY; }
        ch.name_off -= S.skip;
        if (pieces) pieces[k] = ch;
        k++;
    }
    *n_pieces = k; *tf_len = S.tr.tf_len;
    return S3G_OK;
}

int s3g_shard_plan(s3g_ctx *ctx, const void *d_tf_all, uint64_t tf_total, const uint64_t *soff, uint64_t n_streams, int level,
                   uint64_t *n_blocks, uint32_t *nblock, uint32_t *stream_of, uint64_t cap)
{
    if (!ctx || !n_blocks || !soff || (!d_tf_all && tf_total)) { set_error("null argument"); return S3G_E_PARAM; }
    if (level < 1 || level > 9) { set_error("block_size_100k must be 1..9"); return S3G_E_PARAM; }
    if (((uintptr_t)d_tf_all & 15) != 0) { set_error("transformed buffer must be 16-byte aligned"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    ShardState &S = st(ctx);
    S.n_streams = n_streams; S.level = level;
    *n_blocks = 0;
    if (n_streams == 0) { ctx->h_blocks.clear(); ctx->rle_blocks = 0; return S3G_OK; }
    S3G_TRY(ctx->soff.ensure((n_streams + 2) * 8));
    S3G_TRY(upload_small(ctx, 0, ctx->soff.p, soff, (n_streams + 1) * 8));     // run_rle_plan synchronises before the slot is used again
    CutResult cut;
    stage_mark(ctx, 1);
    S3G_TRY(run_rle_plan(ctx, static_cast<const uint8_t *>(d_tf_all), tf_total, ctx->soff.as<uint64_t>(), n_streams, level, &cut));
    stage_mark(ctx, -1);
    ctx->last_rle_bytes = cut.rle_bytes;
    *n_blocks = cut.n_blocks;
    if (cut.n_blocks > cap) { set_error("block table too small: need %llu", (unsigned long long)cut.n_blocks); return S3G_E_CAPACITY; }
    for (uint64_t b = 0; b < cut.n_blocks; b++) {
        if (nblock) nblock[b] = ctx->h_blocks[b].nblock;
        if (stream_of) stream_of[b] = ctx->h_blocks[b].chrom;
    }
    return S3G_OK;
}

int s3g_shard_compress(s3g_ctx *ctx, uint64_t b_lo, uint64_t b_hi, uint64_t *n_bits, uint32_t *crc, uint32_t *n_mtf)
{
    if (!ctx) { set_error("null argument"); return S3G_E_PARAM; }
    if (b_lo > b_hi || b_hi > ctx->rle_blocks) { set_error("block range outside the plan"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    ctx->pool_words = 0;
    S3G_TRY(ctx->pool_woff.ensure((ctx->rle_blocks + 1) * 8));
    if (b_lo == b_hi) return S3G_OK;
    stage_mark(ctx, 1);
    S3G_TRY(run_rle_fill(ctx, b_lo, b_hi));
    S3G_TRY(compress_block_range(ctx, b_lo, b_hi));
    stage_mark(ctx, -1);
    S3G_CUDA(cudaMemcpyAsync(ctx->h_blocks.data() + b_lo, ctx->blocks.as<BlockInfo>() + b_lo, (b_hi - b_lo) * sizeof(BlockInfo),
                             cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    for (uint64_t b = b_lo; b < b_hi; b++) {
        if (n_bits) n_bits[b - b_lo] = ctx->h_blocks[b].n_bits;
        if (crc) crc[b - b_lo] = ctx->h_blocks[b].crc;
        if (n_mtf) n_mtf[b - b_lo] = ctx->h_blocks[b].n_mtf;
    }
    return S3G_OK;
}

int s3g_shard_assemble(s3g_ctx *ctx, const uint64_t *n_bits_all, const uint32_t *crc_all, uint64_t b_lo, uint64_t b_hi,
                       void **d_bytes, uint64_t *byte_lo, uint64_t *byte_hi, uint64_t *stream_off, uint64_t *stream_len)
{
    if (!ctx || !n_bits_all || !crc_all || !d_bytes || !byte_lo || !byte_hi) { set_error("null argument"); return S3G_E_PARAM; }
    if (b_lo > b_hi || b_hi > ctx->rle_blocks) { set_error("block range outside the plan"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    ShardState &S = st(ctx);
    for (uint64_t b = 0; b < ctx->rle_blocks; b++) { ctx->h_blocks[b].n_bits = n_bits_all[b]; ctx->h_blocks[b].crc = crc_all[b]; }
    std::vector<StreamMeta> metas;
    stage_mark(ctx, 5);
    S3G_TRY(run_assemble_range(ctx, S.n_streams, S.level, b_lo, b_hi, byte_lo, byte_hi, &metas));
    stage_mark(ctx, -1);
    for (uint64_t s = 0; s < S.n_streams; s++) {
        if (stream_off) stream_off[s] = metas[s].byte_off;
        if (stream_len) stream_len[s] = metas[s].byte_len;
    }
    *d_bytes = ctx->streams.p;
    return S3G_OK;
}

int s3g_shard_place(s3g_ctx *ctx, uint64_t gather_buf, uint64_t byte_lo, uint64_t byte_hi)
{
    if (!ctx || (!gather_buf && byte_hi > byte_lo)) { set_error("null argument"); return S3G_E_PARAM; }
    if (gather_buf & 3) { set_error("gather buffer must be 4-byte aligned"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    if (byte_hi > byte_lo) {
        stage_mark(ctx, 5);
        S3G_TRY(run_place_bytes(ctx, reinterpret_cast<uint8_t *>(gather_buf), byte_lo, byte_hi - byte_lo));
        stage_mark(ctx, -1);
    }
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    return S3G_OK;
}

int s3g_stage_times(s3g_ctx *ctx, double *stage_ms)
{
    if (!ctx || !stage_ms) { set_error("null argument"); return S3G_E_PARAM; }
    S3G_CUDA(cudaSetDevice(ctx->device));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    stage_collect(ctx, stage_ms);
    return S3G_OK;
}

}  // extern "C"
