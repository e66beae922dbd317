// common.cuh -- shared host/device plumbing for the starch3 B200 hot path.
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <utility>
#include "../../include/starch3_b200.h"

namespace s3g {

void set_error(const char *fmt, ...);

#define S3G_CUDA(call)                                                                   \
    do {                                                                                 \
        cudaError_t e_ = (call);                                                         \
        if (e_ != cudaSuccess) {                                                         \
            s3g::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); \
            return e_ == cudaErrorMemoryAllocation ? S3G_E_NOMEM : S3G_E_CUDA;           \
        }                                                                                \
    } while (0)

#define S3G_TRY(call)                 \
    do {                              \
        int rc_ = (call);             \
        if (rc_ != S3G_OK) return rc_; \
    } while (0)

// Grow-only device buffer; contexts keep them across calls so steady-state
// steps do not pay cudaMalloc / cudaFree.
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return S3G_OK;
        if (p) { cudaFree(p); p = nullptr; cap = 0; }
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            e = cudaMalloc(&p, bytes);
            want = bytes;
        }
        if (e != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            set_error("cudaMalloc(%zu bytes) failed: %s", bytes, cudaGetErrorString(e));
            return S3G_E_NOMEM;
        }
        cap = want;
        return S3G_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// ---- geometry of one bzip2 block in the batched device layout -------------
// Inside a batch of stages 3b..3d every block owns a fixed-stride slot, so batch block b's arrays start at
// b*BLK_STRIDE.  The post-RLE1 bytes of ALL blocks of a call are packed instead (BlockInfo.blk_off): an input of
// many small chromosomes has many small blocks.
// nblock never exceeds 100000*9-19 + 9 (bz/bzlib.c:194, :284-289 and the final flush).
constexpr uint32_t BLK_STRIDE = 900096;         // elements per block slot (multiple of 128)
constexpr uint32_t BITS_WORDS = 491520;         // 32-bit words per block of emitted bits (>= 17 bit * 900001 + header)
constexpr int      SM_COUNT   = 148;

// Block descriptor shared by stages 3a..3e (device and host).
struct BlockInfo {
    uint64_t in_start, in_end;   // byte range of the (concatenated) tf buffer committed to this block
    uint64_t e_base;             // RLE1 output offset (stream-relative prefix) at in_start
    uint32_t nblock;             // bytes after RLE1
    uint32_t chrom;              // owning stream
    uint32_t crc;                // finalised block CRC
    uint32_t n_in_use;
    int32_t  orig_ptr;
    uint32_t tie;                // 1 if equal rotations remained (periodic block)
    uint32_t n_mtf;
    uint32_t pad;
    uint64_t n_bits;             // bits of this block incl. its 105-bit block header
    uint64_t bit_off;            // bit offset inside its stream
    uint64_t blk_off;            // where the block's post-RLE1 bytes start inside ctx->blk_bytes (packed, 128-byte aligned)
};
// bytes a block of n post-RLE1 bytes takes in the packed blk_bytes buffer
__host__ __device__ inline uint64_t blk_slot_bytes(uint32_t n) { return ((uint64_t)n + 64 + 127) & ~(uint64_t)127; }

// ---- per-device context -----------------------------------------------------
struct Ctx {
    int device = 0;
    cudaStream_t stream = nullptr;
    cudaStream_t own_stream = nullptr;
    uint64_t launches = 0;
    // cudaFuncSetAttribute(MaxDynamicSharedMemorySize) is per device: each context sets it once on its own device
    bool attr_bwt = false, attr_mtf = false, attr_huff = false, attr_bs = false;
    uint64_t bucket_blocks = 0, bucket_handed_back = 0;   // blocks sorted by the bucket form; batches in which it handed blocks back
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    // pinned host scratch for small read-backs
    uint64_t *h_scalars = nullptr;   // 64 x u64
    // stage buffers (grow-only)
    DevBuf bed, tile_cnt, line_start, start, stop, rem_off, flags, chrom_first;
    DevBuf scan_a, scan_b, scan_c, scalars;
    DevBuf front_incl, front_flag;         // one-pass front end: inclusive prefix and status word per chunk
    uint32_t front_gen = 0;                // generation tag of the status words (nothing is cleared between calls)
    DevBuf tf, chroms, stat_b, soff;
    // what run_tokenize measured (tokenize_transform.cu): inputs of run_transform_rest / run_range_summary
    uint32_t front_halo = 0, front_skip = 0, front_line1_flag = 0;
    int64_t front_tail_max = INT64_MIN;
    DevBuf rle_carry, rle_ebase, blocks, blk_prov, blk_bytes, in_use, seq_map, stream_tab;
    DevBuf sa, rk, kv0, kv1, hist, bwt_misc, bwt_ghist, lcol;
    size_t sweep_cap = 0;                  // capacity of `hist` when its look-back status words were last cleared
    uint32_t sweep_gen = 0;                // generation tag of the last radix pass (1..255)
    int last_host_entry = 0;               // s3g_last_host_entry
    uint64_t sort_retries = 0;             // times the radix passes had to be repeated with peer-mask ranking (expected: never)
    DevBuf mtf0, mtfv16, mtf_freq, ztiles, bits, pool, pool_woff, streams, stream_meta;
    DevBuf io_a, io_b, io_c, io_d, io_e;   // staging for the stage entry points
    // host mirrors
    std::vector<BlockInfo> h_blocks;
    std::vector<s3g_chrom> h_chroms;
    uint64_t pool_words = 0;               // words appended to the pool so far
    uint64_t last_streams_size = 0;        // bytes of ctx->streams written by the last call on this context
    uint64_t archive_hint = 0;             // compressed size of the last host-entry call: sizes the next pinned archive buffer
    const uint8_t *last_streams_host = nullptr;   // after the pipelined host entry the streams live in the pinned archive only
    // stage marks of the current call: (stage id, event); consecutive marks bracket a stage
    std::vector<cudaEvent_t> mark_pool;
    std::vector<std::pair<int, cudaEvent_t>> marks;
    uint64_t last_rle_bytes = 0;
    uint64_t last_dec_blocks = 0;          // bzip2 blocks decoded by the last decoder call (decode.cu)
    void *shard = nullptr;                 // state between the s3g_shard_* phases (shard.cu)
    void *stream_state = nullptr;          // state between s3g_stream_begin and s3g_stream_end (api.cu)
    // the plan run_rle_plan left (input of run_rle_fill)
    const uint8_t *rle_in = nullptr; uint64_t rle_n = 0; const uint64_t *rle_soff = nullptr; uint64_t rle_streams = 0, rle_blocks = 0;
    uint8_t *h_archive = nullptr;          // pinned; holds the archive of the last compress call
    size_t h_archive_cap = 0;
    // pipelined host entry (s3g_compress_bed on large inputs): upload stream, one event per input range,
    // and two worker contexts that compress ranges while later ranges are still on their way
    cudaStream_t copy_stream = nullptr;
    uint8_t *h_stage = nullptr;            // pinned pieces the copier threads stage a pageable input through
    std::vector<cudaEvent_t> stage_ev;
    std::vector<cudaEvent_t> part_ev;
    Ctx *sub[2] = {nullptr, nullptr};
    // chained host entry (one chromosome spans several ranges): transformed bytes of the current step behind the tail the step
    // before left unfinished (ping-pong), and the compressed bytes of all steps
    DevBuf chain_tf[2], chain_out;
    cudaStream_t out_stream = nullptr;     // device-to-host copies of finished bytes beside the next step's kernels
    // small host tables reach the device through a kernel that reads pinned host memory (upload_small): a cudaMemcpyAsync
    // would queue on the copy engine behind a range of the input that is on its way (up to 3 ms at 160 MB)
    uint64_t *h_small[2] = {nullptr, nullptr};
    size_t h_small_cap[2] = {0, 0};
    cudaEvent_t out_ev = nullptr;
    double mem_frac = 0.85;                // share of the free device memory a batch of blocks may take (worker contexts: less)
    // per-kernel profiling (off by default)
    bool prof = false;
    std::string prof_filter;               // non-empty: only this kernel is timed
    struct ProfRec { const char *name; cudaEvent_t e0, e1; double bytes; };
    double prof_next_bytes = 0;
    std::vector<ProfRec> prof_recs;
    std::vector<cudaEvent_t> prof_pool;
    size_t prof_used = 0;
};

// Every kernel launch goes through this macro: it counts launches and, when
// per-kernel profiling is on (s3g_profile), brackets the launch with CUDA events
// on the launching stream.
// S3G_BYTES(ctx, b) before a launch records its algorithmic bytes for the
// roofline figure of bench.py.
#define S3G_BYTES(ctx, b) ((ctx)->prof_next_bytes = (double)(b))
#define S3G_LAUNCH(ctx, kernel, grid, block, smem, ...)                     \
    do {                                                                    \
        int pi_ = s3g::prof_begin((ctx), #kernel);                          \
        kernel<<<(grid), (block), (smem), (ctx)->stream>>>(__VA_ARGS__);    \
        s3g::prof_end((ctx), pi_);                                          \
        (ctx)->launches++;                                                  \
    } while (0)

// the same for a kernel whose CTAs form thread-block clusters of `cs` along x (cs == 1: an ordinary launch)
#define S3G_LAUNCH_CLUSTER(ctx, kernel, grid, block, smem, cs, ...)          \
    do {                                                                    \
        int pi_ = s3g::prof_begin((ctx), #kernel);                          \
        cudaLaunchConfig_t cfg_ = {};                                       \
        cfg_.gridDim = (grid); cfg_.blockDim = (block);                     \
        cfg_.dynamicSmemBytes = (smem); cfg_.stream = (ctx)->stream;        \
        cudaLaunchAttribute at_[1];                                         \
        at_[0].id = cudaLaunchAttributeClusterDimension;                    \
        at_[0].val.clusterDim.x = (cs); at_[0].val.clusterDim.y = 1; at_[0].val.clusterDim.z = 1; \
        cfg_.attrs = at_; cfg_.numAttrs = (cs) > 1 ? 1 : 0;                 \
        cudaLaunchKernelEx(&cfg_, kernel, __VA_ARGS__);                     \
        s3g::prof_end((ctx), pi_);                                          \
        (ctx)->launches++;                                                  \
    } while (0)

int check_launch(const char *what);
// `bytes` (a multiple of 8) of a small host table -> device memory, in stream order, without the copy engine; `slot` (0, 1)
// names the pinned staging area: a slot must not be reused before the stream has passed the earlier upload
int upload_small(Ctx *ctx, int slot, void *d_dst, const void *h_src, size_t bytes);
// stage mark: the time until the next mark is charged to `stage` (s3g_result.stage_ms)
void stage_mark(Ctx *ctx, int stage);
void stage_collect(Ctx *ctx, double *stage_ms);
void shard_state_free(Ctx *ctx);
void stream_state_free(Ctx *ctx);
int prof_begin(Ctx *ctx, const char *name);
void prof_end(Ctx *ctx, int idx);

// ---- device helpers ---------------------------------------------------------
#ifdef __CUDACC__

__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31; }

template <class T> __device__ __forceinline__ T warp_incl_sum(T v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (unsigned)d) v += o;
    }
    return v;
}
template <class T> __device__ __forceinline__ T warp_incl_max(T v)
{
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        T o = __shfl_up_sync(0xffffffffu, v, d);
        if (lane_id() >= (unsigned)d) v = o > v ? o : v;
    }
    return v;
}

// Exclusive block-wide sum over one value per thread.  `sm` needs 33 slots.
// Returns the exclusive prefix; *total gets the block total.  All threads must call.
template <class T> __device__ __forceinline__ T block_excl_sum(T v, T *sm, T *total)
{
    T inc = warp_incl_sum(v);
    unsigned w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane_id() == 31) sm[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = lane_id() < nw ? sm[lane_id()] : T(0);
        T xi = warp_incl_sum(x);
        sm[lane_id()] = xi - x;
        if (lane_id() == 31) sm[32] = xi;
    }
    __syncthreads();
    *total = sm[32];
    return sm[w] + inc - v;
}

// Exclusive block-wide max over one value per thread (identity 0): the max over
// threads with a smaller index.  `sm` needs 33 slots; *total gets the block max.
template <class T> __device__ __forceinline__ T block_excl_max(T v, T *sm, T *total)
{
    T inc = warp_incl_max(v);
    unsigned w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    __syncthreads();
    if (lane_id() == 31) sm[w] = inc;
    __syncthreads();
    if (w == 0) {
        T x = lane_id() < nw ? sm[lane_id()] : T(0);
        T xi = warp_incl_max(x);
        T ex = __shfl_up_sync(0xffffffffu, xi, 1);
        if (lane_id() == 0) ex = T(0);
        sm[lane_id()] = ex;
        if (lane_id() == 31) sm[32] = xi;
    }
    __syncthreads();
    *total = sm[32];
    T up = __shfl_up_sync(0xffffffffu, inc, 1);
    if (lane_id() == 0) up = T(0);
    T pre = sm[w];
    return pre > up ? pre : up;
}

#endif  // __CUDACC__

// ---- stage drivers (device pointers in, device pointers out) ---------------
struct TfResult {
    uint64_t n_lines = 0, n_chroms = 0, tf_len = 0, dropped = 0;
};
// kernels (1)+(2); leaves ctx->tf (bytes), ctx->chroms (s3g_chrom[n_chroms]) and the per-line arrays on device
// `skip` (< 16): leading bytes of d_bed that belong to the line before the range (see k_count_newlines)
int run_transform(Ctx *ctx, const uint8_t *d_bed, uint64_t n, TfResult *out, bool tokenize_only, uint32_t skip = 0, bool last_part = true);
// after the synchronise that follows run_transform: diagnostics of the lines of the chromosomes the range keeps
inline uint64_t front_unsorted(const Ctx *ctx);
inline uint64_t front_crlf(const Ctx *ctx);
// the same in pieces, for ranges handed between GPUs (shard.cu): `halo` = line 0 is the last line before the range
int run_tokenize(Ctx *ctx, const uint8_t *d_bed, uint64_t n, uint32_t skip, TfResult *out, uint32_t halo);
int run_range_summary(Ctx *ctx, uint64_t n_lines, uint32_t halo, int64_t *tail_max, uint64_t *last_flag, uint32_t *continues);
// peer_bufs (n_peers > 0): the transformed bytes go into these buffers at peer_off instead of ctx->tf (N-GPU path: this GPU's
// own buffer first, then the peers' copies of it through their NVLink-mapped pointers)
int run_transform_rest(Ctx *ctx, const uint8_t *d_bed, uint64_t n, TfResult *out, uint32_t halo, int64_t carry_max, bool dump = false, bool last_part = true,
                       const uint64_t *peer_bufs = nullptr, uint32_t n_peers = 0, uint64_t peer_off = 0, uint64_t multicast_buf = 0);

struct CutResult {
    uint64_t n_blocks = 0;
    uint64_t rle_bytes = 0;        // bytes after RLE1, all blocks
};
// kernel (3a) over a concatenated buffer of `n_streams` streams; stream s covers [d_soff[s], d_soff[s+1]).
// Leaves ctx->blocks (BlockInfo[n_blocks]), ctx->blk_bytes (block b at BlockInfo.blk_off), ctx->in_use (256 B per block).
int run_rle_cut(Ctx *ctx, const uint8_t *d_in, uint64_t n, const uint64_t *d_soff, uint64_t n_streams,
                int level, CutResult *out);
// the same in two steps: the plan of ALL blocks, then the bytes / CRC / maps of a range of them (a GPU's share)
int run_rle_plan(Ctx *ctx, const uint8_t *d_in, uint64_t n, const uint64_t *d_soff, uint64_t n_streams,
                 int level, CutResult *out);
int run_rle_fill(Ctx *ctx, uint64_t b_lo, uint64_t b_hi);
int run_block_crc(Ctx *ctx, const uint8_t *d_in, BlockInfo *d_blocks, uint64_t nb, double bytes);

// kernels (3b..3d) on blocks [b0, b0+nb) of ctx->blocks
int run_bwt(Ctx *ctx, uint64_t b0, uint64_t nb);                 // -> ctx->sa, ctx->lcol, BlockInfo.orig_ptr / tie
int run_mtf(Ctx *ctx, uint64_t b0, uint64_t nb);                 // -> ctx->mtfv16, mtf_freq, BlockInfo.n_mtf
// -> ctx->bits (BITS_WORDS words per block), BlockInfo.n_bits; optional selector / len dumps for the parity tests
int run_huff(Ctx *ctx, uint64_t b0, uint64_t nb, int with_block_header, uint8_t *d_sel_out, uint8_t *d_len_out);
// append the bits of blocks [b0, b0+nb) to ctx->pool (word offsets in ctx->pool_woff)
int run_pool_append(Ctx *ctx, uint64_t b0, uint64_t nb);
// kernel (3e): bit-level concatenation of the pooled blocks into per-stream byte strings (ctx->streams);
// fills ctx->stream_meta (StreamMeta[n_streams])
struct StreamMeta { uint64_t byte_off, byte_len, n_blocks; uint32_t combined_crc, pad; };
int run_assemble(Ctx *ctx, uint64_t n_blocks, uint64_t n_streams, int level, uint64_t *total_bytes);
// the same for a GPU's share [b_lo, b_hi) of the blocks; needs n_bits / crc / chrom of EVERY block in ctx->h_blocks.
// Leaves bytes [byte_lo, byte_hi) of the global streams buffer in ctx->streams.
int run_assemble_range(Ctx *ctx, uint64_t n_streams, int level, uint64_t b_lo, uint64_t b_hi, uint64_t *byte_lo, uint64_t *byte_hi,
                       std::vector<StreamMeta> *metas);
// the common back end: blocks [b_lo, b_hi) at the bit positions items[0 .. b_hi - b_lo), then n_patch (bit position, word)
// pairs, relative to the first bit of the `nbytes` bytes left in ctx->streams
int run_assemble_items(Ctx *ctx, uint64_t b_lo, uint64_t b_hi, const std::vector<uint64_t> &items, uint64_t n_patch, uint64_t nbytes);
// the byte string run_assemble_range left in ctx->streams -> dst[at, at + len) (dst may be a peer GPU's buffer)
int run_place_bytes(Ctx *ctx, uint8_t *dst, uint64_t at, uint64_t len);
int compress_block_range(Ctx *ctx, uint64_t b_lo, uint64_t b_hi);

inline uint64_t front_unsorted(const Ctx *ctx) { return ctx->h_scalars[4]; }
inline uint64_t front_crlf(const Ctx *ctx) { return ctx->h_scalars[5]; }

}  // namespace s3g

struct s3g_ctx : s3g::Ctx {};
