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

// ---- a GPU's share of the blocks (multi-GPU, shard.cu) ---------------------------------------------------
// The block table (ctx->h_blocks: n_bits, crc, chrom of EVERY block, own or not) fixes the layout of all streams; this
// GPU places its blocks [b_lo, b_hi), plus the header of every stream whose first block and the trailer of every
// stream whose last block it owns, into a local byte string that covers bytes [byte_lo, byte_hi) of the global
// streams buffer.  The two end bytes may be shared with the neighbouring shares (blocks are not byte aligned): the
// host ORs them.
__global__ void k_concat_range(const BlockInfo *blocks, const uint32_t *pool, const uint64_t *woff, const uint64_t *pos, uint32_t *dst)
{
    uint64_t b = blockIdx.x >> 5, part = blockIdx.x & 31;            // relative to the first own block
    uint64_t nw = (blocks[b].n_bits + 31) >> 5;
    const uint32_t *src = pool + woff[b];
    uint64_t base = pos[b];
    for (uint64_t i = part * blockDim.x + threadIdx.x; i < nw; i += (uint64_t)32 * blockDim.x)
        or_bits(dst, base + i * 32, src[i]);
}
__global__ void k_patch_words(const uint64_t *patch, uint64_t n, uint32_t *dst)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) or_bits(dst, patch[2 * i], (uint32_t)patch[2 * i + 1]);
}

// the share's byte string (ctx->streams, `len` bytes) into a gather buffer at `at`.  The buffer may live on another GPU
// (an NVLink-mapped peer pointer) and is zero where nobody has written: the string's first and last byte may be shared
// with the neighbouring shares, so they are ORed in (an atomic on the 32-bit word that holds them); the bytes between
// belong to this share alone and are stored.
__global__ void k_place_bytes(const uint8_t *__restrict__ src, uint8_t *dst, uint64_t at, uint64_t len)
{
    const uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (uint64_t)gridDim.x * blockDim.x;
    if (t == 0) {
        auto or_byte = [&](uint64_t pos, uint8_t v) {
            const uint64_t a = at + pos;
            atomicOr(reinterpret_cast<uint32_t *>(dst + (a & ~3ull)), (uint32_t)v << (8 * (a & 3)));
        };
        or_byte(0, src[0]);
        if (len > 1) or_byte(len - 1, src[len - 1]);
    }
    if (len <= 2) return;
    // interior [1, len - 1): bytes up to the first 16-byte boundary of the destination, aligned vectors, the rest
    const uint64_t d0 = at + 1, d1 = at + len - 1;                     // destination range
    const uint64_t v0 = (d0 + 15) & ~15ull, v1 = d1 & ~15ull;
    if (v0 >= v1) {
        for (uint64_t p = d0 + t; p < d1; p += nt) dst[p] = src[p - at];
        return;
    }
    for (uint64_t p = d0 + t; p < v0; p += nt) dst[p] = src[p - at];
    for (uint64_t p = v1 + t; p < d1; p += nt) dst[p] = src[p - at];
    const uint32_t sh = (uint32_t)((v0 - at) & 3);                     // source phase relative to its 4-byte words
    const uint32_t *sw = reinterpret_cast<const uint32_t *>(src + ((v0 - at) & ~3ull));
    for (uint64_t u = t; u < (v1 - v0) / 16; u += nt) {
        const uint32_t *q = sw + 4 * u;
        uint32_t w[5] = {q[0], q[1], q[2], q[3], sh ? q[4] : 0u};
        uint4 o;
        o.x = __funnelshift_r(w[0], w[1], 8 * sh); o.y = __funnelshift_r(w[1], w[2], 8 * sh);
        o.z = __funnelshift_r(w[2], w[3], 8 * sh); o.w = __funnelshift_r(w[3], w[4], 8 * sh);
        *reinterpret_cast<uint4 *>(dst + v0 + 16 * u) = o;
    }
}
int run_place_bytes(Ctx *ctx, uint8_t *dst, uint64_t at, uint64_t len)
{
    if (!len) return S3G_OK;
    S3G_BYTES(ctx, 2.0 * (double)len);
    S3G_LAUNCH(ctx, k_place_bytes, 592, 256, 0, ctx->streams.as<uint8_t>(), dst, at, len);
    return check_launch("place bytes");
}

// Blocks [b_lo, b_hi) of ctx->blocks (their bits in the pool) at the bit positions items[0 .. b_hi - b_lo), then
// n_patch (bit position, 32-bit word) pairs -- stream headers and trailers -- all relative to the first bit of a byte
// string of `nbytes` bytes, which is left in ctx->streams.
int run_assemble_items(Ctx *ctx, uint64_t b_lo, uint64_t b_hi, const std::vector<uint64_t> &items, uint64_t n_patch, uint64_t nbytes)
{
    const uint64_t words = (nbytes + 3) / 4 + 4;
    S3G_TRY(ctx->streams.ensure(words * 4));
    uint32_t *dst = ctx->streams.as<uint32_t>();
    S3G_CUDA(cudaMemsetAsync(dst, 0, words * 4, ctx->stream));
    if (items.size() != (b_hi - b_lo) + 2 * n_patch) { set_error("assemble: item table does not match the block range"); return S3G_E_PARAM; }
    if (items.empty()) return S3G_OK;
    const uint64_t patch_at = b_hi - b_lo;
    S3G_TRY(ctx->io_d.ensure(items.size() * 8 + 64));
    S3G_TRY(upload_small(ctx, 1, ctx->io_d.p, items.data(), items.size() * 8));   // staged: `items` may die with the caller's frame; the callers
                                                                                  // synchronise (place / read-back) before the next table
    if ((b_hi - b_lo) > (1ull << 26)) { set_error("too many bzip2 blocks in one call"); return S3G_E_LIMIT; }
    if (b_hi > b_lo)
        S3G_LAUNCH(ctx, k_concat_range, (unsigned)((b_hi - b_lo) * 32), 256, 0, ctx->blocks.as<BlockInfo>() + b_lo, ctx->pool.as<uint32_t>(),
                   ctx->pool_woff.as<uint64_t>() + b_lo, ctx->io_d.as<uint64_t>(), dst);
    if (n_patch)
        S3G_LAUNCH(ctx, k_patch_words, (unsigned)((n_patch + 127) / 128), 128, 0, ctx->io_d.as<uint64_t>() + patch_at, n_patch, dst);
    S3G_LAUNCH(ctx, k_bswap, 592, 256, 0, dst, words);
    return check_launch("assemble items");
}

int run_assemble_range(Ctx *ctx, uint64_t n_streams, int level, uint64_t b_lo, uint64_t b_hi, uint64_t *byte_lo, uint64_t *byte_hi,
                       std::vector<StreamMeta> *metas)
{
    const std::vector<BlockInfo> &hb = ctx->h_blocks;
    const uint64_t nb = hb.size();
    metas->assign(n_streams, StreamMeta());
    std::vector<uint64_t> gpos(nb, 0), first(n_streams + 1, nb);     // global bit position of every block; first block of a stream
    {
        uint64_t b = 0, byte_off = 0;
        for (uint64_t s = 0; s < n_streams; s++) {
            first[s] = b;
            uint64_t bits = 32; uint32_t comb = 0;
            for (; b < nb && hb[b].chrom == s; b++) {
                gpos[b] = byte_off * 8 + bits;
                bits += hb[b].n_bits;
                comb = ((comb << 1) | (comb >> 31)) ^ hb[b].crc;                 // bz/compress.c:607-608
            }
            bits += 48 + 32;
            StreamMeta &m = (*metas)[s];
            m.byte_off = byte_off; m.byte_len = (bits + 7) >> 3; m.n_blocks = b - first[s]; m.combined_crc = comb; m.pad = 0;
            byte_off += m.byte_len;
        }
        first[n_streams] = nb;
        if (b != nb) { set_error("block table and stream table disagree"); return S3G_E_PARAM; }
    }
    *byte_lo = *byte_hi = 0;
    if (b_lo >= b_hi) return S3G_OK;
    // bit range of the share
    const uint64_t s_lo = hb[b_lo].chrom, s_hi = hb[b_hi - 1].chrom;
    const bool own_head = first[s_lo] == b_lo, own_tail = first[s_hi + 1] == b_hi;
    const uint64_t bit_lo = own_head ? (*metas)[s_lo].byte_off * 8 : gpos[b_lo];
    const uint64_t bit_hi = own_tail ? ((*metas)[s_hi].byte_off + (*metas)[s_hi].byte_len) * 8 : gpos[b_hi - 1] + hb[b_hi - 1].n_bits;
    *byte_lo = bit_lo >> 3; *byte_hi = (bit_hi + 7) >> 3;
    const uint64_t base_bit = *byte_lo * 8;
    // positions of the own blocks, and the header / trailer words of the streams that begin / end in the share
    std::vector<uint64_t> up;
    up.reserve((b_hi - b_lo) + 8 * (s_hi - s_lo + 1));
    for (uint64_t b = b_lo; b < b_hi; b++) up.push_back(gpos[b] - base_bit);
    uint64_t n_patch = 0;
    for (uint64_t s = s_lo; s <= s_hi; s++) {
        const StreamMeta &m = (*metas)[s];
        if (first[s] >= b_lo && first[s] < b_hi) {
            up.push_back(m.byte_off * 8 - base_bit); up.push_back(0x425a6800u | (uint32_t)('0' + level)); n_patch++;   // bz/compress.c:622-628
        }
        const uint64_t last = first[s + 1] - 1;
        if (last >= b_lo && last < b_hi) {
            const uint64_t end = gpos[last] + hb[last].n_bits - base_bit;                                                 // :657-666
            up.push_back(end); up.push_back(0x17724538u);
            up.push_back(end + 32); up.push_back(0x50900000u | (m.combined_crc >> 16));
            up.push_back(end + 64); up.push_back((uint64_t)(uint32_t)(m.combined_crc << 16));
            n_patch += 3;
        }
    }
    return run_assemble_items(ctx, b_lo, b_hi, up, n_patch, *byte_hi - *byte_lo);
}

}  // namespace s3g
