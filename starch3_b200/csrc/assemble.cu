// assemble.cu -- kernel (3e): bit-level concatenation of the compressed blocks into
// per-stream bzip2 byte strings.
//
// Reference framing (bz/compress.c:602-667): "BZh" + level once per stream, then the
// blocks back to back with NO byte alignment between them (:609 keeps bsBuff/bsLive),
// then 0x177245385090 + combined CRC, padded to a byte.  combined = rotl(combined,1) ^
// blockCRC (:607-608).  Here every block was emitted into its own word buffer starting
// at bit 0; its place in the stream is an exclusive scan over block bit lengths and the
// merge is a shift-and-OR of 32-bit words.
#include "common.cuh"

namespace s3g {

// ---- pool: compact storage of finished blocks ---------------------------------
__global__ void __launch_bounds__(1024) k_pool_offsets(const BlockInfo *blocks, uint32_t nb, uint64_t pool_base, uint64_t *woff, uint64_t *total)
{
    __shared__ uint64_t sm[33];
    uint64_t carry = pool_base;
    for (uint32_t b0 = 0; b0 < nb; b0 += 1024) {
        uint32_t b = b0 + threadIdx.x;
        uint64_t w = b < nb ? (blocks[b].n_bits + 31) >> 5 : 0, tot;
        uint64_t ex = block_excl_sum<uint64_t>(w, sm, &tot);
        if (b < nb) woff[b] = carry + ex;
        carry += tot;
        __syncthreads();
    }
    if (threadIdx.x == 0) *total = carry;
}

__global__ void k_pool_copy(const BlockInfo *blocks, const uint32_t *bits, const uint64_t *woff, uint32_t *pool)
{
    uint32_t lb = blockIdx.y;
    uint64_t nw = (blocks[lb].n_bits + 31) >> 5;
    const uint32_t *src = bits + (uint64_t)lb * BITS_WORDS;
    uint32_t *dst = pool + woff[lb];
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nw; i += (uint64_t)gridDim.x * blockDim.x) dst[i] = src[i];
}

int run_pool_append(Ctx *ctx, uint64_t b0, uint64_t nb)
{
    if (nb == 0) return S3G_OK;
    uint64_t *d_sc = ctx->scalars.as<uint64_t>();
    uint64_t *woff = ctx->pool_woff.as<uint64_t>() + b0;
    S3G_LAUNCH(ctx, k_pool_offsets, 1, 1024, 0, ctx->blocks.as<BlockInfo>() + b0, (uint32_t)nb, ctx->pool_words, woff, d_sc + 20);
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars + 20, d_sc + 20, 8, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    S3G_TRY(check_launch("pool offsets"));
    uint64_t new_total = ctx->h_scalars[20];
    if (new_total * 4 > ctx->pool.cap) {      // grow, keeping what is there
        DevBuf nbuf;
        S3G_TRY(nbuf.ensure(new_total * 4 + new_total));
        if (ctx->pool_words) S3G_CUDA(cudaMemcpyAsync(nbuf.p, ctx->pool.p, ctx->pool_words * 4, cudaMemcpyDeviceToDevice, ctx->stream));
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
        ctx->pool.release();
        ctx->pool = nbuf;
    }
    dim3 grid(64, (unsigned)nb);
    S3G_LAUNCH(ctx, k_pool_copy, grid, 256, 0, ctx->blocks.as<BlockInfo>() + b0, ctx->bits.as<uint32_t>(), woff, ctx->pool.as<uint32_t>());
    ctx->pool_words = new_total;
    return check_launch("pool copy");
}

// ---- stream layout --------------------------------------------------------------
__global__ void k_stream_layout(BlockInfo *blocks, const uint64_t *first_block, uint64_t n_streams, StreamMeta *meta)
{
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    uint64_t bits = 32;                // "BZh9"
    uint32_t comb = 0;
    for (uint64_t b = first_block[s]; b < first_block[s + 1]; b++) {
        blocks[b].bit_off = bits;
        bits += blocks[b].n_bits;
        comb = ((comb << 1) | (comb >> 31)) ^ blocks[b].crc;
    }
    bits += 48 + 32;
    meta[s].byte_len = (bits + 7) >> 3;
    meta[s].n_blocks = first_block[s + 1] - first_block[s];
    meta[s].combined_crc = comb;
    meta[s].pad = 0;
}

__global__ void k_stream_offsets(StreamMeta *meta, uint64_t n_streams, uint64_t *total)
{
    if (blockIdx.x || threadIdx.x) return;
    uint64_t acc = 0;
    for (uint64_t s = 0; s < n_streams; s++) { meta[s].byte_off = acc; acc += meta[s].byte_len; }
    *total = acc;
}

__device__ __forceinline__ void or_bits(uint32_t *dst, uint64_t bitpos, uint32_t w)
{
    // OR the 32 bits of w (MSB first) into the big-endian bit string dst at bitpos
    uint64_t idx = bitpos >> 5; uint32_t sh = (uint32_t)(bitpos & 31);
    if (w == 0) return;
    atomicOr(&dst[idx], w >> sh);
    if (sh) { uint32_t lo = w << (32 - sh); if (lo) atomicOr(&dst[idx + 1], lo); }
}

__global__ void k_concat(const BlockInfo *blocks, const uint32_t *pool, const uint64_t *woff, const StreamMeta *meta, uint32_t *dst)
{
    // 32 CTAs per block, block index in grid.x (grid.y stops at 65535; many small chromosomes make more blocks than that)
    uint64_t b = blockIdx.x >> 5, part = blockIdx.x & 31;
    const BlockInfo &B = blocks[b];
    uint64_t nw = (B.n_bits + 31) >> 5;
    const uint32_t *src = pool + woff[b];
    uint64_t base = meta[B.chrom].byte_off * 8 + B.bit_off;
    for (uint64_t i = part * blockDim.x + threadIdx.x; i < nw; i += (uint64_t)32 * blockDim.x)
        or_bits(dst, base + i * 32, src[i]);
}

__global__ void k_stream_frame(const BlockInfo *blocks, const uint64_t *first_block, const StreamMeta *meta, uint64_t n_streams,
                               int level, uint32_t *dst)
{
    uint64_t s = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= n_streams) return;
    uint64_t base = meta[s].byte_off * 8;
    or_bits(dst, base, 0x425a6800u | (uint32_t)('0' + level));        // bz/compress.c:622-628
    uint64_t end = base + 32;
    if (first_block[s + 1] > first_block[s]) { const BlockInfo &L = blocks[first_block[s + 1] - 1]; end = base + L.bit_off + L.n_bits; }
    or_bits(dst, end, 0x17724538u);                                    // :657-666
    or_bits(dst, end + 32, 0x50900000u | (meta[s].combined_crc >> 16));
    or_bits(dst, end + 64, meta[s].combined_crc << 16);
}

__global__ void k_bswap(uint32_t *w, uint64_t n)
{
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x)
        w[i] = __byte_perm(w[i], 0, 0x0123);
}

int run_assemble(Ctx *ctx, uint64_t n_blocks, uint64_t n_streams, int level, uint64_t *total_bytes)
{
    *total_bytes = 0;
    if (n_streams == 0) return S3G_OK;
    uint64_t *d_sc = ctx->scalars.as<uint64_t>();
    S3G_TRY(ctx->stream_meta.ensure(n_streams * sizeof(StreamMeta)));
    const uint64_t *first_block = ctx->stream_tab.as<uint64_t>() + (n_streams + 2);
    StreamMeta *meta = ctx->stream_meta.as<StreamMeta>();
    BlockInfo *blocks = ctx->blocks.as<BlockInfo>();
    S3G_LAUNCH(ctx, k_stream_layout, (unsigned)((n_streams + 63) / 64), 64, 0, blocks, first_block, n_streams, meta);
    S3G_LAUNCH(ctx, k_stream_offsets, 1, 1, 0, meta, n_streams, d_sc + 21);
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars + 21, d_sc + 21, 8, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    S3G_TRY(check_launch("stream layout"));
    uint64_t total = ctx->h_scalars[21];
    uint64_t words = (total + 3) / 4 + 4;
    S3G_TRY(ctx->streams.ensure(words * 4));
    uint32_t *dst = ctx->streams.as<uint32_t>();
    S3G_CUDA(cudaMemsetAsync(dst, 0, words * 4, ctx->stream));
    if (n_blocks) {
        if (n_blocks > (1ull << 26)) { set_error("too many bzip2 blocks in one call"); return S3G_E_LIMIT; }
        S3G_LAUNCH(ctx, k_concat, (unsigned)(n_blocks * 32), 256, 0, blocks, ctx->pool.as<uint32_t>(), ctx->pool_woff.as<uint64_t>(), meta, dst);
    }
    S3G_LAUNCH(ctx, k_stream_frame, (unsigned)((n_streams + 63) / 64), 64, 0, blocks, first_block, meta, n_streams, level, dst);
    S3G_LAUNCH(ctx, k_bswap, 592, 256, 0, dst, words);
    *total_bytes = total;
    return check_launch("assemble");
}

}  // namespace s3g
