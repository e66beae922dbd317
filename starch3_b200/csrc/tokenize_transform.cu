// tokenize_transform.cu -- kernels (1) and (2) of the hot path.
//
// (1) newline scan + tab split + integer parse + chromosome-boundary flags:
//     replaces produce_line / consume_line (hpp:158-199, :201-309) and the
//     strcmp chromosome test (hpp:325-342).
// (2) the starch coordinate transform update_transformation_state (hpp:428-504)
//     with the per-chromosome reset (hpp:523-532) and the statistics the
//     reference declares but never computes (hpp:61-62).
// hpp = /root/reference/include/starch3api.hpp.
//
// The per-line dependencies are radius-1 (previous stop, previous length,
// previous chromosome); everything else is scans for output offsets and
// chromosome ids.
#include "common.cuh"
#include "scan.cuh"

namespace s3g {

constexpr int NL_THREADS = 256;
constexpr int NL_SUB = NL_THREADS * 16;     // bytes per sub-iteration (one uint4 per thread)
constexpr int NL_ITERS = 4;
constexpr int NL_TILE = NL_SUB * NL_ITERS;  // 16 KiB of input per CTA

__device__ __forceinline__ unsigned nl_mask16(const uint8_t *bed, uint64_t pos, uint64_t n)
{
    // bit k set <=> bed[pos+k] == '\n'
    unsigned m = 0;
    if (pos + 16 <= n) {
        uint4 v = *reinterpret_cast<const uint4 *>(bed + pos);
        unsigned w[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int q = 0; q < 4; q++) {
            unsigned eq = __vcmpeq4(w[q], 0x0a0a0a0au);   // 0xff per matching byte
            // gather the low bit of each byte into 4 bits
            unsigned b = eq & 0x01010101u;
            b = (b | (b >> 7) | (b >> 14) | (b >> 21)) & 0xfu;
            m |= b << (4 * q);
        }
    } else {
        for (int k = 0; k < 16; k++)
            if (pos + k < n && bed[pos + k] == '\n') m |= 1u << k;
    }
    return m;
}

// `skip` (< 16): the first bytes of the buffer belong to the line before the range being processed (a range that
// starts at an arbitrary line start is handed over with its base rounded down to 16 bytes); their newline is ignored.
__global__ void __launch_bounds__(NL_THREADS) k_count_newlines(const uint8_t *bed, uint64_t n, uint64_t *tile_cnt, uint32_t skip)
{
    __shared__ uint32_t sm[33];
    uint64_t tile0 = (uint64_t)blockIdx.x * NL_TILE;
    uint32_t c = 0;
#pragma unroll
    for (int it = 0; it < NL_ITERS; it++) {
        uint64_t pos = tile0 + (uint64_t)it * NL_SUB + (uint64_t)threadIdx.x * 16;
        if (pos < n) {
            unsigned m = nl_mask16(bed, pos, n);
            if (pos == 0) m &= 0xffffffffu << skip;
            c += __popc(m);
        }
    }
    uint32_t tot;
    block_excl_sum<uint32_t>(c, sm, &tot);
    if (threadIdx.x == 0) tile_cnt[blockIdx.x] = tot;
}

struct SumU64 {
    typedef uint64_t T;
    __host__ __device__ static T identity() { return 0; }
    __host__ __device__ static T op(T a, T b) { return a + b; }
};

__global__ void __launch_bounds__(NL_THREADS) k_line_starts(const uint8_t *bed, uint64_t n, const uint64_t *tile_base,
                                                            uint64_t *line_start, uint32_t skip)
{
    __shared__ uint32_t sm[33];
    uint64_t tile0 = (uint64_t)blockIdx.x * NL_TILE;
    uint64_t run = tile_base[blockIdx.x];
    if (blockIdx.x == 0 && threadIdx.x == 0) line_start[0] = skip;
    for (int it = 0; it < NL_ITERS; it++) {
        uint64_t pos = tile0 + (uint64_t)it * NL_SUB + (uint64_t)threadIdx.x * 16;
        unsigned m = pos < n ? nl_mask16(bed, pos, n) : 0;
        if (pos == 0) m &= 0xffffffffu << skip;
        uint32_t tot;
        uint32_t ex = block_excl_sum<uint32_t>(__popc(m), sm, &tot);
        uint64_t idx = run + ex + 1;
        while (m) {
            int k = __ffs(m) - 1;
            m &= m - 1;
            line_start[idx++] = pos + k + 1;
        }
        run += tot;
    }
}

// flags: bit0 = chromosome differs from the previous line (hpp:331), bit1 = malformed
__global__ void k_parse_lines(const uint8_t *__restrict__ bed, const uint64_t *__restrict__ line_start, uint64_t n_lines,
                              int64_t *__restrict__ start, int64_t *__restrict__ stop, uint32_t *__restrict__ rem_off,
                              uint8_t *__restrict__ flags, unsigned long long *malformed)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_lines) return;
    uint64_t p = line_start[i], e = line_start[i + 1] - 1;   // bed[e] == '\n'
    int field = 0;
    uint64_t acc = 0;
    int neg = 0, st = 0;   // st: 0 = expecting sign/digit, 1 = in digits, 2 = number ended
    int64_t v_start = 0, v_stop = 0;
    uint64_t q = p, t0 = e;
    for (; q < e; q++) {
        uint8_t c = bed[q];
        if (c == '\t') {
            if (field == 0) t0 = q;
            else if (field == 1) v_start = neg ? (int64_t)(0 - acc) : (int64_t)acc;
            else { break; }
            field++;
            acc = 0; neg = 0; st = 0;
            continue;
        }
        if (field >= 1) {
            // sscanf("%lld") over the documented domain: [sign] digits, stop at the first other byte
            if (st == 0 && (c == '-' || c == '+')) { neg = (c == '-'); st = 1; }
            else if (st <= 1 && c >= '0' && c <= '9') { acc = acc * 10 + (uint64_t)(c - '0'); st = 1; }
            else st = 2;
        }
    }
    uint8_t fl = 0;
    if (field == 2) v_stop = neg ? (int64_t)(0 - acc) : (int64_t)acc;   // ended by '\n' (BED3) or by the third tab
    else fl |= 2;
    // q == e (no fourth field) or q at the third tab
    uint32_t ro = (uint32_t)((q < e ? q + 1 : e) - p);
    if (i == 0) fl |= 1;
    else {
        // strcmp(chr, previous chr) != 0 (hpp:331); the previous line's field ends at its first tab
        uint64_t pp = line_start[i - 1], cl = t0 - p;
        bool diff = false;
        for (uint64_t k = 0; k < cl; k++)
            if (bed[p + k] != bed[pp + k]) { diff = true; break; }
        if (!diff && bed[pp + cl] != '\t') diff = true;
        if (diff) fl |= 1;
    }
    if (fl & 2) atomicAdd(malformed, 1ull);
    start[i] = v_start; stop[i] = v_stop; rem_off[i] = ro; flags[i] = fl;
}

__constant__ uint64_t POW10[20] = {1ull, 10ull, 100ull, 1000ull, 10000ull, 100000ull, 1000000ull, 10000000ull, 100000000ull, 1000000000ull,
                                   10000000000ull, 100000000000ull, 1000000000000ull, 10000000000000ull, 100000000000000ull,
                                   1000000000000000ull, 10000000000000000ull, 100000000000000000ull, 1000000000000000000ull,
                                   10000000000000000000ull};
// decimal digits of a (1..20): floor(bits * log10(2)) is the count or one short of it (no division)
__device__ __forceinline__ int dec_digits(uint64_t a)
{
    int b = 64 - __clzll((long long)(a | 1));
    int t = (b * 1233) >> 12;
    int d = t + (a >= POW10[t] ? 1 : 0);
    return d < 1 ? 1 : d;
}
__device__ __forceinline__ int dec_len(int64_t v)
{
    // printed length of "%lld": n_digits (hpp:559-581) plus the sign it does not count
    uint64_t a = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    return dec_digits(a) + (v < 0);
}

struct LineView {
    const uint64_t *line_start;
    const int64_t *start, *stop;
    const uint32_t *rem_off;
    const uint8_t *flags;
    uint32_t halo;      // 1: line 0 is the last line BEFORE the range (multi-GPU ranges, shard.cu): it hands its stop,
                        // length and chromosome to line 1 and is itself neither written nor counted
    __device__ __forceinline__ void prev(uint64_t i, int64_t *p_stop, int64_t *p_len) const
    {
        if (flags[i] & 1) { *p_stop = 0; *p_len = 0; }                         // hpp:523-532
        else { int64_t s0 = start[i - 1], s1 = stop[i - 1]; *p_stop = s1; *p_len = (int64_t)((uint64_t)s1 - (uint64_t)s0); }
    }
    __device__ __forceinline__ uint32_t out_len(uint64_t i) const
    {
        if (halo && i == 0) return 0;
        int64_t ps, pl; prev(i, &ps, &pl);
        int64_t s = start[i], t = stop[i];
        int64_t len = (int64_t)((uint64_t)t - (uint64_t)s), d = (int64_t)((uint64_t)s - (uint64_t)ps);
        uint32_t rem_len = (uint32_t)(line_start[i + 1] - 1 - line_start[i]) - rem_off[i];
        uint32_t o = (uint32_t)dec_len(d) + 1 + (rem_len ? rem_len + 1 : 0);
        if (len != pl) o += 2 + (uint32_t)dec_len(len);
        return o;
    }
};

struct OutLenScan : SumU64 {
    LineView lv;
    uint64_t *line_tf_off;
    __device__ T load(uint64_t i) const { return lv.out_len(i); }
    __device__ void store(uint64_t i, T excl, T) const { line_tf_off[i] = excl; }
};

// segmented running max of stop (exclusive), producing each line's unique-base contribution
struct SegMax {
    int64_t v; int32_t seg; int32_t pad;
};
struct UniqScan {
    typedef SegMax T;
    const int64_t *start, *stop;
    const uint8_t *flags;
    int64_t *uniq;
    uint32_t halo;          // see LineView
    int64_t carry_max;      // halo: the largest stop among ALL earlier lines of the halo line's chromosome (they live on other GPUs)
    __host__ __device__ static T identity() { T t; t.v = INT64_MIN; t.seg = 0; t.pad = 0; return t; }
    __host__ __device__ static T op(T a, T b)
    {
        if (b.seg) return b;
        T r; r.v = a.v > b.v ? a.v : b.v; r.seg = a.seg; r.pad = 0; return r;
    }
    __device__ T load(uint64_t i) const
    {
        T t; t.v = stop[i]; t.seg = flags[i] & 1; t.pad = 0;
        if (halo && i == 0) t.v = carry_max;
        return t;
    }
    __device__ void store(uint64_t i, T excl, T) const
    {
        if (halo && i == 0) { uniq[0] = 0; return; }
        int64_t rm = (flags[i] & 1) ? INT64_MIN : excl.v;
        int64_t s = start[i], t = stop[i];
        int64_t lo = s > rm ? s : rm;
        uniq[i] = t > lo ? t - lo : 0;
    }
};

struct Stat3 {
    int64_t len_sum, uniq_sum; uint64_t chroms;
};
struct StatScan {
    typedef Stat3 T;
    const int64_t *start, *stop, *uniq;
    const uint8_t *flags;
    uint64_t *chrom_first;     // [n_chroms]
    Stat3 *chrom_pref;         // [n_chroms]
    uint32_t halo;             // see LineView: the halo line opens piece 0 and contributes nothing
    __host__ __device__ static T identity() { T t; t.len_sum = 0; t.uniq_sum = 0; t.chroms = 0; return t; }
    __host__ __device__ static T op(T a, T b) { T r; r.len_sum = a.len_sum + b.len_sum; r.uniq_sum = a.uniq_sum + b.uniq_sum; r.chroms = a.chroms + b.chroms; return r; }
    __device__ T load(uint64_t i) const
    {
        T t; t.len_sum = (int64_t)((uint64_t)stop[i] - (uint64_t)start[i]); t.uniq_sum = uniq[i]; t.chroms = flags[i] & 1;
        if (halo && i == 0) { t.len_sum = 0; t.uniq_sum = 0; }
        return t;
    }
    __device__ void store(uint64_t i, T excl, T) const
    {
        if (flags[i] & 1) { chrom_first[excl.chroms] = i; chrom_pref[excl.chroms] = excl; }
    }
};

__global__ void k_chrom_table(const uint8_t *bed, const uint64_t *line_start, const uint64_t *line_tf_off,
                              const uint64_t *chrom_first, const Stat3 *chrom_pref, const Stat3 *stat_total,
                              uint64_t n_chroms, uint64_t n_lines, uint64_t tf_total, s3g_chrom *out, uint32_t halo)
{
    uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n_chroms) return;
    uint64_t first = chrom_first[c];
    uint64_t next = c + 1 < n_chroms ? chrom_first[c + 1] : n_lines;
    Stat3 a = chrom_pref[c];
    Stat3 b = c + 1 < n_chroms ? chrom_pref[c + 1] : *stat_total;
    s3g_chrom r;
    r.name_off = line_start[first];
    uint32_t nl = 0;
    while (bed[r.name_off + nl] != '\t') nl++;
    r.name_len = nl;
    r.n_blocks = 0;
    r.tf_off = line_tf_off[first];
    r.tf_len = (next < n_lines ? line_tf_off[next] : tf_total) - r.tf_off;
    r.line_count = (int64_t)(next - first) - (halo && c == 0 ? 1 : 0);
    r.bases_nonunique = b.len_sum - a.len_sum;
    r.bases_unique = b.uniq_sum - a.uniq_sum;
    r.bz_off = 0; r.bz_len = 0;
    out[c] = r;
}

__device__ __forceinline__ uint64_t put_dec(uint8_t *dst, int64_t v)
{
    uint64_t a = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    uint64_t k = 0;
    if (v < 0) dst[k++] = '-';
    const int nd = dec_digits(a);
    if (a <= 0xffffffffull) {            // the usual case: 32-bit divisions by a constant
        uint32_t x = (uint32_t)a;
        for (int j = nd - 1; j >= 0; j--) { uint32_t q = x / 10u; dst[k + j] = (uint8_t)('0' + (x - q * 10u)); x = q; }
    } else {
        for (int j = nd - 1; j >= 0; j--) { uint64_t q = a / 10u; dst[k + j] = (uint8_t)('0' + (uint32_t)(a - q * 10u)); a = q; }
    }
    return k + nd;
}

// One thread formats one line.  The 256 lines of a CTA produce one contiguous piece of the transformed
// buffer, so they are staged in shared memory at the destination's 16-byte phase and written out with
// aligned vector stores (per-thread byte stores to HBM kept the load/store queues full).
constexpr int WT_LINES = 256;
constexpr int WT_BUF = 16384;

__global__ void __launch_bounds__(WT_LINES) k_write_tf(const uint8_t *__restrict__ bed, LineView lv, const uint64_t *__restrict__ line_tf_off,
                                                       uint64_t n_lines, uint64_t tf_total, uint8_t *__restrict__ tf)
{
    __shared__ __align__(16) uint8_t s_buf[WT_BUF + 16];
    const uint64_t i0 = (uint64_t)blockIdx.x * WT_LINES;
    const uint64_t i1 = i0 + WT_LINES < n_lines ? i0 + WT_LINES : n_lines;
    const uint64_t o_begin = line_tf_off[i0], o_end = i1 < n_lines ? line_tf_off[i1] : tf_total;
    const bool staged = o_end - o_begin <= WT_BUF;
    const uint32_t ph = (uint32_t)o_begin & 15u;
    uint64_t i = i0 + threadIdx.x;
    if (i < n_lines && !(lv.halo && i == 0)) {
        int64_t ps, pl; lv.prev(i, &ps, &pl);
        int64_t s = lv.start[i], t = lv.stop[i];
        int64_t len = (int64_t)((uint64_t)t - (uint64_t)s), d = (int64_t)((uint64_t)s - (uint64_t)ps);
        uint8_t *o = staged ? s_buf + ph + (line_tf_off[i] - o_begin) : tf + line_tf_off[i];
        if (len != pl) { *o++ = 'p'; o += put_dec(o, len); *o++ = '\n'; }     // hpp:438-455
        o += put_dec(o, d);                                                   // hpp:456-500
        uint64_t p = lv.line_start[i], e = lv.line_start[i + 1] - 1;
        uint64_t r = p + lv.rem_off[i];
        if (r < e) {
            *o++ = '\t';
            for (uint64_t q = r; q < e; q++) *o++ = bed[q];
        }
        *o = '\n';
    }
    if (!staged) return;
    __syncthreads();
    uint8_t *dst = tf + (o_begin - ph);                                       // 16-byte aligned (tf comes from cudaMalloc)
    const uint32_t lo_b = ph, hi_b = ph + (uint32_t)(o_end - o_begin);
    for (uint32_t c = threadIdx.x * 16; c < hi_b; c += WT_LINES * 16) {
        if (c >= lo_b && c + 16 <= hi_b) *reinterpret_cast<uint4 *>(dst + c) = *reinterpret_cast<const uint4 *>(s_buf + c);
        else for (uint32_t j = c > lo_b ? c : lo_b; j < c + 16 && j < hi_b; j++) dst[j] = s_buf[j];
    }
}

// ---- tail summary of a range (multi-GPU ranges: what the next range must know) ---------------------------
__global__ void k_last_flag(const uint8_t *flags, uint64_t n_lines, unsigned long long *last)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    bool f = i < n_lines && (flags[i] & 1);
    // the largest flagged index of the warp, then one atomic per warp that has one
    unsigned m = __ballot_sync(0xffffffffu, f);
    if (m && (threadIdx.x & 31) == 31 - __clz(m)) atomicMax(last, (unsigned long long)i);
}
__global__ void k_tail_max(const int64_t *stop, uint64_t n_lines, const unsigned long long *last, uint32_t halo, long long *tail_max)
{
    uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    uint64_t lo = *last;
    if (halo && lo == 0) lo = 1;                   // the halo line itself is covered by the carry it brings
    long long v = (i < n_lines && i >= lo) ? (long long)stop[i] : (long long)INT64_MIN;
    for (int d = 16; d; d >>= 1) { long long o = __shfl_xor_sync(0xffffffffu, v, d); v = o > v ? o : v; }
    if ((threadIdx.x & 31) == 0 && v != (long long)INT64_MIN) atomicMax(tail_max, v);
}

// kernels (1): line framing + tokenizer over d_bed[0, n).  Leaves the per-line arrays in ctx.
int run_tokenize(Ctx *ctx, const uint8_t *d_bed, uint64_t n, uint32_t skip, TfResult *out)
{
    *out = TfResult();
    uint64_t ntiles = (n + NL_TILE - 1) / NL_TILE;
    if (ntiles == 0) ntiles = 1;
    if (ntiles > 0x7fffffffull) { set_error("input too large"); return S3G_E_LIMIT; }
    S3G_TRY(ctx->tile_cnt.ensure((ntiles + 1) * 8));
    S3G_TRY(ctx->scalars.ensure(64 * 8));
    uint64_t *d_sc = ctx->scalars.as<uint64_t>();
    S3G_CUDA(cudaMemsetAsync(d_sc, 0, 64 * 8, ctx->stream));
    uint64_t *tile_cnt = ctx->tile_cnt.as<uint64_t>();
    S3G_BYTES(ctx, n);
    S3G_LAUNCH(ctx, k_count_newlines, (unsigned)ntiles, NL_THREADS, 0, d_bed, n, tile_cnt, skip);
    S3G_LAUNCH(ctx, k_scan_agg<SumU64>, 1, AGG_THREADS, 0, tile_cnt, ntiles, d_sc + 0);
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars, d_sc, 8, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    uint64_t n_lines = ctx->h_scalars[0];
    out->n_lines = n_lines;
    S3G_TRY(ctx->line_start.ensure((n_lines + 1) * 8));
    uint64_t *line_start = ctx->line_start.as<uint64_t>();
    S3G_BYTES(ctx, n + 8 * n_lines);
    S3G_LAUNCH(ctx, k_line_starts, (unsigned)ntiles, NL_THREADS, 0, d_bed, n, tile_cnt, line_start, skip);
    if (n_lines == 0) {
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
        out->dropped = n - skip;               // the skip bytes belong to the line before the range
        return check_launch("line_starts");
    }
    S3G_TRY(ctx->start.ensure(n_lines * 8));
    S3G_TRY(ctx->stop.ensure(n_lines * 8));
    S3G_TRY(ctx->rem_off.ensure(n_lines * 4));
    S3G_TRY(ctx->flags.ensure(n_lines));
    unsigned lgrid = (unsigned)((n_lines + 255) / 256);
    S3G_BYTES(ctx, n + 29 * n_lines);
    S3G_LAUNCH(ctx, k_parse_lines, lgrid, 256, 0, d_bed, line_start, n_lines, ctx->start.as<int64_t>(), ctx->stop.as<int64_t>(),
               ctx->rem_off.as<uint32_t>(), ctx->flags.as<uint8_t>(), (unsigned long long *)(d_sc + 1));
    // last line start tells how many unterminated tail bytes are dropped (hpp:181-190)
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars + 40, line_start + n_lines, 8, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars + 1, d_sc + 1, 8, cudaMemcpyDeviceToHost, ctx->stream));
    return S3G_OK;
}

// summary of the tokenised range for the hand-over between GPUs (needs run_tokenize's arrays); synchronises
int run_range_summary(Ctx *ctx, uint64_t n_lines, uint32_t halo, int64_t *tail_max, uint64_t *last_flag, uint32_t *continues)
{
    uint64_t *d_sc = ctx->scalars.as<uint64_t>();
    *tail_max = INT64_MIN; *last_flag = 0; *continues = 0;
    if (n_lines == 0) { S3G_CUDA(cudaStreamSynchronize(ctx->stream)); return S3G_OK; }
    long long neg = (long long)INT64_MIN;
    S3G_CUDA(cudaMemsetAsync(d_sc + 44, 0, 8, ctx->stream));
    S3G_CUDA(cudaMemcpyAsync(d_sc + 45, &neg, 8, cudaMemcpyHostToDevice, ctx->stream));
    unsigned g = (unsigned)((n_lines + 255) / 256);
    S3G_LAUNCH(ctx, k_last_flag, g, 256, 0, ctx->flags.as<uint8_t>(), n_lines, (unsigned long long *)(d_sc + 44));
    S3G_LAUNCH(ctx, k_tail_max, g, 256, 0, ctx->stop.as<int64_t>(), n_lines, (const unsigned long long *)(d_sc + 44), halo, (long long *)(d_sc + 45));
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars + 44, d_sc + 44, 16, cudaMemcpyDeviceToHost, ctx->stream));
    uint8_t *hf = reinterpret_cast<uint8_t *>(ctx->h_scalars + 46);
    hf[0] = 1;
    if (halo && n_lines > 1) S3G_CUDA(cudaMemcpyAsync(hf, ctx->flags.as<uint8_t>() + 1, 1, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    *last_flag = ctx->h_scalars[44];
    *tail_max = (int64_t)ctx->h_scalars[45];
    *continues = (halo && n_lines > 1 && !(hf[0] & 1)) ? 1u : 0u;
    return check_launch("range summary");
}

// kernels (2) over the lines left by run_tokenize
int run_transform_rest(Ctx *ctx, const uint8_t *d_bed, uint64_t n, TfResult *out, uint32_t halo, int64_t carry_max)
{
    uint64_t *d_sc = ctx->scalars.as<uint64_t>();
    const uint64_t n_lines = out->n_lines;
    uint64_t *line_start = ctx->line_start.as<uint64_t>();
    int64_t *start = ctx->start.as<int64_t>(), *stop = ctx->stop.as<int64_t>();
    uint32_t *rem_off = ctx->rem_off.as<uint32_t>();
    uint8_t *flags = ctx->flags.as<uint8_t>();
    unsigned lgrid = (unsigned)((n_lines + 255) / 256);
    // ---- sizes: output offsets, unique-base contributions, chromosome count ----
    uint64_t stiles = (n_lines + SCAN_TILE - 1) / SCAN_TILE + 1;
    S3G_TRY(ctx->line_tf_off.ensure(n_lines * 8));
    S3G_TRY(ctx->stat_a.ensure(n_lines * 8));                 // uniq[]
    S3G_TRY(ctx->scan_a.ensure(stiles * sizeof(uint64_t)));
    S3G_TRY(ctx->scan_b.ensure(stiles * sizeof(SegMax)));
    S3G_TRY(ctx->scan_c.ensure(stiles * sizeof(Stat3)));
    LineView lv{line_start, start, stop, rem_off, flags, halo};
    OutLenScan f1; f1.lv = lv; f1.line_tf_off = ctx->line_tf_off.as<uint64_t>();
    S3G_TRY(device_scan(ctx, f1, n_lines, ctx->scan_a.as<uint64_t>(), d_sc + 3));
    UniqScan f2; f2.start = start; f2.stop = stop; f2.flags = flags; f2.uniq = ctx->stat_a.as<int64_t>(); f2.halo = halo; f2.carry_max = carry_max;
    S3G_TRY(device_scan(ctx, f2, n_lines, ctx->scan_b.as<SegMax>(), (SegMax *)nullptr));
    StatScan f3; f3.start = start; f3.stop = stop; f3.uniq = ctx->stat_a.as<int64_t>(); f3.flags = flags; f3.halo = halo;
    f3.chrom_first = nullptr; f3.chrom_pref = nullptr;
    Stat3 *d_stat_total = reinterpret_cast<Stat3 *>(d_sc + 8);
    unsigned sgrid = (unsigned)(stiles - 1 ? stiles - 1 : 1);
    S3G_LAUNCH(ctx, k_scan_reduce<StatScan>, sgrid, SCAN_THREADS, 0, f3, n_lines, ctx->scan_c.as<Stat3>());
    S3G_LAUNCH(ctx, k_scan_agg<StatScan>, 1, AGG_THREADS, 0, ctx->scan_c.as<Stat3>(), (uint64_t)sgrid, d_stat_total);
    S3G_CUDA(cudaMemcpyAsync(ctx->h_scalars, d_sc, 16 * 8, cudaMemcpyDeviceToHost, ctx->stream));
    S3G_CUDA(cudaStreamSynchronize(ctx->stream));
    S3G_TRY(check_launch("transform sizes"));
    if (ctx->h_scalars[1]) { set_error("%llu malformed BED line(s): fewer than three fields", (unsigned long long)ctx->h_scalars[1]); return S3G_E_MALFORMED; }
    out->dropped = n - ctx->h_scalars[40];
    out->tf_len = ctx->h_scalars[3];
    out->n_chroms = ctx->h_scalars[10];
    // ---- write ----
    S3G_TRY(ctx->tf.ensure(out->tf_len + 64));
    S3G_TRY(ctx->chrom_first.ensure(out->n_chroms * 8));
    S3G_TRY(ctx->stat_b.ensure(out->n_chroms * sizeof(Stat3)));
    S3G_TRY(ctx->chroms.ensure((out->n_chroms + 1) * sizeof(s3g_chrom)));
    f3.chrom_first = ctx->chrom_first.as<uint64_t>(); f3.chrom_pref = ctx->stat_b.as<Stat3>();
    S3G_LAUNCH(ctx, k_scan_apply<StatScan>, sgrid, SCAN_THREADS, 0, f3, n_lines, ctx->scan_c.as<Stat3>());
    S3G_LAUNCH(ctx, k_chrom_table, (unsigned)((out->n_chroms + 127) / 128), 128, 0, d_bed, line_start,
               ctx->line_tf_off.as<uint64_t>(), ctx->chrom_first.as<uint64_t>(), ctx->stat_b.as<Stat3>(), d_stat_total,
               out->n_chroms, n_lines, out->tf_len, ctx->chroms.as<s3g_chrom>(), halo);
    S3G_BYTES(ctx, n + out->tf_len + 37 * n_lines);
    S3G_LAUNCH(ctx, k_write_tf, lgrid, WT_LINES, 0, d_bed, lv, ctx->line_tf_off.as<uint64_t>(), n_lines, out->tf_len, ctx->tf.as<uint8_t>());
    return check_launch("transform write");
}

int run_transform(Ctx *ctx, const uint8_t *d_bed, uint64_t n, TfResult *out, bool tokenize_only, uint32_t skip)
{
    S3G_TRY(run_tokenize(ctx, d_bed, n, skip, out));
    if (out->n_lines == 0) return S3G_OK;
    if (tokenize_only) {
        S3G_CUDA(cudaStreamSynchronize(ctx->stream));
        out->dropped = n - ctx->h_scalars[40];
        if (ctx->h_scalars[1]) { set_error("%llu malformed BED line(s): fewer than three fields", (unsigned long long)ctx->h_scalars[1]); return S3G_E_MALFORMED; }
        return check_launch("tokenize");
    }
    return run_transform_rest(ctx, d_bed, n, out, 0, INT64_MIN);
}

}  // namespace s3g
