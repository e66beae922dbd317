/*
 * s3_oracle.c -- CPU ORACLE for the starch3 compression hot path.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  Only tests/, bench.py's
 * cpu_baseline / --impl reference legs and __graft_entry__.smoke() may load
 * it.  The product path (starch3_b200/csrc) never calls into this file.
 *
 * It is a from-scratch restatement, in plain C, of the algorithms on the
 * path named by BASELINE.json `north_star`:
 *
 *   (1) BED tokenizer + starch coordinate transform
 *         /root/reference/include/starch3api.hpp:158-199 (line framing),
 *         :201-309 (tab split, sscanf), :325-342 (chromosome change),
 *         :428-504 (update_transformation_state), :523-532 (reset).
 *   (2) bzip2 1.0.6 compression as vendored (patched) by the reference in
 *       /root/reference/third-party/bzip2-1.0.6.tar.gz ("bz/" below), driven
 *       the way starch3api.hpp:835-837 configures it (blockSize100k 9,
 *       workFactor 30) -- one bzip2 stream per chromosome, the whole
 *       chromosome fed with a single BZ_FINISH action:
 *         bz/bzlib.c:225-338 (RLE1, CRC, block cut), bz/blocksort.c:212-329
 *         and :1031-1089 (block sort; see s3o_bwt), bz/compress.c:106-231
 *         (MTF + RUNA/RUNB), :239-598 (Huffman table selection + emission),
 *         :602-667 (framing), bz/huffman.c:63-166.
 *   (3) The archive container of ARCHIVE_FORMAT.md (the reference pins only
 *       the 4 magic bytes, starch3api.hpp:907-910 -- the rest is "parity
 *       unpinned by the reference" and is specified by this repository).
 *
 * Parity pins (tests/test_oracle_*.py): bzip2's own golden vectors
 * sample{1,2,3}.bz2 (bz/Makefile:56-69), the reference-compiled libbz2
 * (oracle/_ref/libs3ref.so), CPython's bz2 module, and the transformed
 * stream dumped by the reference-compiled starch3 binary on
 * single-chromosome inputs (tests/golden/).
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define S3O_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------ */
/* Part 1: tokenizer + transform                                      */
/* ------------------------------------------------------------------ */

typedef struct {
    uint64_t name_off;      /* byte offset of the chromosome name in the input */
    uint32_t name_len;
    uint32_t pad;
    uint64_t tf_off;        /* offset of this chromosome's stream in the tf buffer */
    uint64_t tf_len;
    int64_t  line_count;    /* hpp:503 */
    int64_t  bases_nonunique; /* sum(stop-start); hpp:62 declares it, nothing computes it */
    int64_t  bases_unique;    /* union coverage; hpp:61 likewise */
} s3o_chrom_t;

/* decimal digits of |v| (hpp:559-581); the sign is NOT counted there */
static int ndigits_u64(uint64_t v)
{
    int d = 1;
    while (v >= 10) { v /= 10; d++; }
    return d;
}

static size_t put_i64(uint8_t *dst, int64_t v)
{
    /* what sprintf("%" PRId64) at hpp:451/:471/:474 produces */
    uint64_t a = v < 0 ? (uint64_t)0 - (uint64_t)v : (uint64_t)v;
    int nd = ndigits_u64(a);
    size_t k = 0;
    if (v < 0) dst[k++] = '-';
    for (int i = nd - 1; i >= 0; i--) { dst[k + i] = (uint8_t)('0' + a % 10); a /= 10; }
    return k + nd;
}

/* sscanf("%lld") restricted to the documented domain: optional sign, digits,
 * stops at the first non-digit; empty -> 0 (hpp:306-307). Wraps on overflow. */
static int64_t parse_i64(const uint8_t *p, size_t n)
{
    size_t i = 0; int neg = 0; uint64_t a = 0;
    if (i < n && (p[i] == '-' || p[i] == '+')) { neg = (p[i] == '-'); i++; }
    while (i < n && p[i] >= '0' && p[i] <= '9') { a = a * 10 + (uint64_t)(p[i] - '0'); i++; }
    return neg ? (int64_t)((uint64_t)0 - a) : (int64_t)a;
}

/*
 * Returns 0 on success, -1 if tf_cap / chrom_cap too small, -2 on a malformed
 * line (fewer than three fields).  An unterminated final line is dropped, as
 * the reference does (produce_line hpp:181-190); *dropped gets its length.
 */
S3O_API int s3o_transform(const uint8_t *bed, uint64_t n,
                          uint8_t *tf, uint64_t tf_cap, uint64_t *tf_len,
                          s3o_chrom_t *chroms, uint64_t chrom_cap, uint64_t *n_chroms,
                          uint64_t *dropped)
{
    uint64_t pos = 0, out = 0, nc = 0;
    const uint8_t *cur_chr = NULL; size_t cur_chr_len = 0;
    int64_t prev_stop = 0, prev_len = 0, run_max = 0;
    *dropped = 0;
    while (pos < n) {
        const uint8_t *nl = memchr(bed + pos, '\n', n - pos);
        if (!nl) { *dropped = n - pos; break; }
        uint64_t eol = (uint64_t)(nl - bed);
        /* first three tabs split the line; rem keeps its inner tabs (hpp:221) */
        uint64_t t[3]; int nt = 0;
        for (uint64_t i = pos; i < eol && nt < 3; i++) if (bed[i] == '\t') t[nt++] = i;
        if (nt < 2) return -2;
        uint64_t stop_end = nt == 3 ? t[2] : eol;
        int64_t start = parse_i64(bed + t[0] + 1, t[1] - t[0] - 1);
        int64_t stop  = parse_i64(bed + t[1] + 1, stop_end - t[1] - 1);
        uint64_t rem_off = nt == 3 ? t[2] + 1 : eol;
        uint64_t rem_len = eol - rem_off;
        size_t chr_len = t[0] - pos;
        /* chromosome change: strcmp != 0 (hpp:331); reappearing names open a new stream */
        if (!cur_chr || chr_len != cur_chr_len || memcmp(cur_chr, bed + pos, chr_len) != 0) {
            if (nc) chroms[nc - 1].tf_len = out - chroms[nc - 1].tf_off;
            if (nc == chrom_cap) return -1;
            memset(&chroms[nc], 0, sizeof chroms[nc]);
            chroms[nc].name_off = pos; chroms[nc].name_len = (uint32_t)chr_len;
            chroms[nc].tf_off = out;
            nc++;
            cur_chr = bed + pos; cur_chr_len = chr_len;
            prev_stop = 0; prev_len = 0;            /* hpp:523-532 */
            run_max = INT64_MIN;
        }
        if (out + 64 + rem_len > tf_cap) return -1;
        int64_t len = (int64_t)((uint64_t)stop - (uint64_t)start);
        if (len != prev_len) {                       /* hpp:438-455 */
            tf[out++] = 'p'; out += put_i64(tf + out, len); tf[out++] = '\n';
            prev_len = len;
        }
        int64_t d = (int64_t)((uint64_t)start - (uint64_t)prev_stop);  /* hpp:456-500 */
        out += put_i64(tf + out, d);
        if (rem_len) { tf[out++] = '\t'; memcpy(tf + out, bed + rem_off, rem_len); out += rem_len; }
        tf[out++] = '\n';
        prev_stop = stop;                            /* hpp:501-503 */
        s3o_chrom_t *c = &chroms[nc - 1];
        c->line_count++;
        c->bases_nonunique += len;
        int64_t lo = start > run_max ? start : run_max;
        if (stop > lo) c->bases_unique += stop - lo;
        if (stop > run_max) run_max = stop;
        pos = eol + 1;
    }
    if (nc) chroms[nc - 1].tf_len = out - chroms[nc - 1].tf_off;
    *tf_len = out; *n_chroms = nc;
    return 0;
}

/* ------------------------------------------------------------------ */
/* Part 2: bzip2                                                       */
/* ------------------------------------------------------------------ */

static uint32_t crc_tab[256];
static void crc_init(void)
{
    /* bz/crctable.c:32 is this table: CRC-32/BZIP2, poly 0x04C11DB7, MSB first */
    if (crc_tab[1]) return;
    for (uint32_t i = 0; i < 256; i++) {
        uint32_t c = i << 24;
        for (int k = 0; k < 8; k++) c = (c & 0x80000000u) ? (c << 1) ^ 0x04C11DB7u : (c << 1);
        crc_tab[i] = c;
    }
}

S3O_API uint32_t s3o_crc32(const uint8_t *p, uint64_t n)
{
    crc_init();
    uint32_t c = 0xFFFFFFFFu;                        /* bz/bzlib_private.h:157-172 */
    for (uint64_t i = 0; i < n; i++) c = (c << 8) ^ crc_tab[(c >> 24) ^ p[i]];
    return ~c;
}

typedef struct {
    uint64_t in_start, in_end;  /* input bytes whose runs were committed to this block */
    uint32_t nblock;
    uint32_t crc;
    uint8_t  in_use[256];
    uint8_t *data;              /* nblock bytes (post-RLE1) */
} s3o_block_t;

typedef struct {
    s3o_block_t *b; size_t n, cap;
} blocklist_t;

static s3o_block_t *new_block(blocklist_t *L, uint64_t in_start, uint32_t cap_bytes)
{
    if (L->n == L->cap) { L->cap = L->cap ? L->cap * 2 : 8; L->b = realloc(L->b, L->cap * sizeof *L->b); }
    s3o_block_t *b = &L->b[L->n++];
    memset(b, 0, sizeof *b);
    b->in_start = in_start; b->crc = 0xFFFFFFFFu;
    b->data = malloc(cap_bytes);
    return b;
}

/* commit one (byte,len) run: bz/bzlib.c:225-256 */
static void commit_run(s3o_block_t *b, uint8_t ch, int len)
{
    for (int i = 0; i < len; i++) b->crc = (b->crc << 8) ^ crc_tab[(b->crc >> 24) ^ ch];
    b->in_use[ch] = 1;
    int c = len < 4 ? len : 4;
    for (int i = 0; i < c; i++) b->data[b->nblock++] = ch;
    if (len >= 4) { b->in_use[len - 4] = 1; b->data[b->nblock++] = (uint8_t)(len - 4); }
}

/*
 * RLE1 + block cut, following the byte-serial state machine of
 * copy_input_until_stop / ADD_CHAR_TO_BLOCK / handle_compress
 * (bz/bzlib.c:269-338, :370-412) for a stream fed in ONE BZ_FINISH call:
 *   - the "block full" test (nblock >= nblockMAX) happens before each input byte,
 *   - a pending run is committed when the byte changes or it reaches 255,
 *   - closing a full block does NOT flush the pending run,
 *   - at end of input the pending run is flushed into the current block, even
 *     if that block is already full (handle_compress tests the finish
 *     condition first, bz/bzlib.c:393-396).
 */
static void rle1_cut(const uint8_t *in, uint64_t n, int level, blocklist_t *L)
{
    crc_init();
    uint32_t nmax = 100000u * (uint32_t)level - 19;  /* bz/bzlib.c:194 */
    uint32_t cap = 100000u * (uint32_t)level + 16;
    int ch = 256, len = 0;
    uint64_t run_start = 0;
    s3o_block_t *b = new_block(L, 0, cap);
    for (uint64_t i = 0; i < n; i++) {
        if (b->nblock >= nmax) {                     /* close, keep pending run */
            b->in_end = run_start;
            b = new_block(L, run_start, cap);
        }
        int c = in[i];
        if (c != ch || len == 255) {
            if (ch < 256) commit_run(b, (uint8_t)ch, len);
            ch = c; len = 1; run_start = i;
        } else len++;
    }
    if (ch < 256) commit_run(b, (uint8_t)ch, len);   /* flush_RL bz/bzlib.c:261-265 */
    b->in_end = n;
}

/* --- block sort ---------------------------------------------------- */
/*
 * Restatement of fallbackSort (bz/blocksort.c:212-329) with its helper sorts
 * (:32-61 simple sort, :93-180 ternary quicksort).  libbz2 guarantees the BWT
 * is the same whichever of mainSort / fallbackSort runs (bz/blocksort.c:
 * 1058-1061), and for blocks whose rotations are not all distinct (periodic
 * blocks) SURVEY.md section 7 shows libbz2 always ends in fallbackSort from the
 * intact block -- so emulating it step for step reproduces origPtr too.
 */
static void fb_small_sort(uint32_t *fmap, const uint32_t *ec, int lo, int hi)
{
    if (lo == hi) return;
    if (hi - lo > 3) {
        for (int i = hi - 4; i >= lo; i--) {
            uint32_t t = fmap[i], e = ec[t]; int j;
            for (j = i + 4; j <= hi && e > ec[fmap[j]]; j += 4) fmap[j - 4] = fmap[j];
            fmap[j - 4] = t;
        }
    }
    for (int i = hi - 1; i >= lo; i--) {
        uint32_t t = fmap[i], e = ec[t]; int j;
        for (j = i + 1; j <= hi && e > ec[fmap[j]]; j++) fmap[j - 1] = fmap[j];
        fmap[j - 1] = t;
    }
}

static inline void fb_swap(uint32_t *fmap, int a, int b) { uint32_t t = fmap[a]; fmap[a] = fmap[b]; fmap[b] = t; }
static inline void fb_vswap(uint32_t *fmap, int a, int b, int n) { while (n-- > 0) fb_swap(fmap, a++, b++); }

static void fb_qsort3(uint32_t *fmap, const uint32_t *ec, int lo0, int hi0)
{
    int slo[100], shi[100], sp = 0;
    uint32_t r = 0;
    slo[sp] = lo0; shi[sp] = hi0; sp++;
    while (sp > 0) {
        sp--; int lo = slo[sp], hi = shi[sp];
        if (hi - lo < 10) { fb_small_sort(fmap, ec, lo, hi); continue; }
        r = (r * 7621 + 1) % 32768;                  /* bz/blocksort.c:126 */
        uint32_t med;
        switch (r % 3) {
            case 0: med = ec[fmap[lo]]; break;
            case 1: med = ec[fmap[(lo + hi) >> 1]]; break;
            default: med = ec[fmap[hi]]; break;
        }
        int unLo = lo, ltLo = lo, unHi = hi, gtHi = hi;
        for (;;) {
            while (unLo <= unHi) {
                int32_t d = (int32_t)ec[fmap[unLo]] - (int32_t)med;
                if (d == 0) { fb_swap(fmap, unLo, ltLo); ltLo++; unLo++; continue; }
                if (d > 0) break;
                unLo++;
            }
            while (unLo <= unHi) {
                int32_t d = (int32_t)ec[fmap[unHi]] - (int32_t)med;
                if (d == 0) { fb_swap(fmap, unHi, gtHi); gtHi--; unHi--; continue; }
                if (d < 0) break;
                unHi--;
            }
            if (unLo > unHi) break;
            fb_swap(fmap, unLo, unHi); unLo++; unHi--;
        }
        if (gtHi < ltLo) continue;                   /* everything equal to the pivot */
        int a = ltLo - lo, b = unLo - ltLo; int n = a < b ? a : b;
        fb_vswap(fmap, lo, unLo - n, n);
        a = hi - gtHi; b = gtHi - unHi; int m = a < b ? a : b;
        fb_vswap(fmap, unLo, hi - m + 1, m);
        n = lo + unLo - ltLo - 1;
        m = hi - (gtHi - unHi) + 1;
        if (n - lo > hi - m) { slo[sp] = lo; shi[sp] = n; sp++; slo[sp] = m; shi[sp] = hi; sp++; }
        else                 { slo[sp] = m; shi[sp] = hi; sp++; slo[sp] = lo; shi[sp] = n; sp++; }
    }
}

#define BH_SET(z)   (bh[(z) >> 5] |= (1u << ((z) & 31)))
#define BH_CLR(z)   (bh[(z) >> 5] &= ~(1u << ((z) & 31)))
#define BH_GET(z)   (bh[(z) >> 5] & (1u << ((z) & 31)))

/* ptr[0..n) = sorted rotation starts; returns origPtr (bz/blocksort.c:1083-1086) */
S3O_API int32_t s3o_bwt(const uint8_t *block, int32_t n, uint32_t *fmap)
{
    uint32_t *ec = malloc(((size_t)n + 8) * sizeof *ec);
    uint32_t *bh = calloc((size_t)n / 32 + 8, sizeof *bh);
    int32_t ftab[257] = {0};
    for (int i = 0; i < n; i++) ftab[block[i]]++;
    { int32_t acc = 0; for (int i = 0; i < 256; i++) { acc += ftab[i]; ftab[i] = acc; } }
    for (int i = 0; i < n; i++) { int k = --ftab[block[i]]; fmap[k] = (uint32_t)i; }
    for (int i = 0; i < 256; i++) BH_SET(ftab[i]);
    for (int i = 0; i < 32; i++) { BH_SET(n + 2 * i); BH_CLR(n + 2 * i + 1); }

    for (int64_t H = 1;; ) {
        int j = 0;
        for (int i = 0; i < n; i++) {
            if (BH_GET(i)) j = i;
            int k = (int)fmap[i] - (int)H; if (k < 0) k += n;
            ec[k] = (uint32_t)j;
        }
        int notdone = 0, r = -1;
        for (;;) {
            int k = r + 1;
            while (BH_GET(k)) k++;                   /* skip singleton headers */
            int l = k - 1;
            if (l >= n) break;
            while (!BH_GET(k)) k++;
            r = k - 1;
            if (r >= n) break;
            if (r > l) {
                notdone += r - l + 1;
                fb_qsort3(fmap, ec, l, r);
                int32_t cc = -1;
                for (int i = l; i <= r; i++) {
                    int32_t c1 = (int32_t)ec[fmap[i]];
                    if (cc != c1) { BH_SET(i); cc = c1; }
                }
            }
        }
        H *= 2;
        if (H > n || notdone == 0) break;
    }
    int32_t orig = -1;
    for (int i = 0; i < n; i++) if (fmap[i] == 0) { orig = i; break; }
    free(ec); free(bh);
    return orig;
}

/* --- MTF + zero-run coding (bz/compress.c:106-231) ------------------- */
/* returns nMTF; mtfv needs n+1 slots; freq[258] */
S3O_API int32_t s3o_mtf(const uint8_t *block, int32_t n, const uint32_t *ptr,
                        const uint8_t *in_use, uint16_t *mtfv, int32_t *freq, int32_t *n_in_use)
{
    uint8_t seq[256], yy[256];
    int nu = 0;
    for (int i = 0; i < 256; i++) if (in_use[i]) seq[i] = (uint8_t)nu++;
    int eob = nu + 1;
    for (int i = 0; i <= eob; i++) freq[i] = 0;
    for (int i = 0; i < nu; i++) yy[i] = (uint8_t)i;
    int32_t wr = 0, zpend = 0;
    for (int i = 0; i <= n; i++) {
        int sym = -1;
        if (i < n) {
            int32_t j = (int32_t)ptr[i] - 1; if (j < 0) j += n;
            sym = seq[block[j]];
        }
        if (i < n && yy[0] == sym) { zpend++; continue; }
        if (zpend > 0) {                             /* bijective base-2, LSB first */
            zpend--;
            for (;;) {
                int s = zpend & 1; mtfv[wr++] = (uint16_t)s; freq[s]++;
                if (zpend < 2) break;
                zpend = (zpend - 2) / 2;
            }
            zpend = 0;
        }
        if (i == n) break;
        int p = 0; uint8_t carry = yy[0];
        while (carry != sym) { p++; uint8_t t = yy[p]; yy[p] = carry; carry = t; }
        yy[0] = carry;
        mtfv[wr++] = (uint16_t)(p + 1); freq[p + 1]++;
    }
    mtfv[wr++] = (uint16_t)eob; freq[eob]++;
    *n_in_use = nu;
    return wr;
}

/* --- Huffman code lengths (bz/huffman.c:63-148) ---------------------- */
static void hb_make_lengths(uint8_t *len, const int32_t *freq, int alpha, int max_len)
{
    int32_t heap[260], weight[520], parent[520];
    for (int i = 0; i < alpha; i++) weight[i + 1] = (freq[i] == 0 ? 1 : freq[i]) << 8;
    for (;;) {
        int nnodes = alpha, nheap = 0;
        heap[0] = 0; weight[0] = 0; parent[0] = -2;
        for (int i = 1; i <= alpha; i++) {
            parent[i] = -1;
            int z = ++nheap; heap[z] = i;
            int t = heap[z];
            while (weight[t] < weight[heap[z >> 1]]) { heap[z] = heap[z >> 1]; z >>= 1; }
            heap[z] = t;
        }
        while (nheap > 1) {
            int n12[2];
            for (int q = 0; q < 2; q++) {
                n12[q] = heap[1]; heap[1] = heap[nheap--];
                int z = 1, t = heap[1];
                for (;;) {
                    int y = z << 1;
                    if (y > nheap) break;
                    if (y < nheap && weight[heap[y + 1]] < weight[heap[y]]) y++;
                    if (weight[t] < weight[heap[y]]) break;
                    heap[z] = heap[y]; z = y;
                }
                heap[z] = t;
            }
            nnodes++;
            parent[n12[0]] = parent[n12[1]] = nnodes;
            uint32_t w1 = (uint32_t)weight[n12[0]], w2 = (uint32_t)weight[n12[1]];
            uint32_t d1 = w1 & 0xff, d2 = w2 & 0xff;
            weight[nnodes] = (int32_t)(((w1 & 0xffffff00u) + (w2 & 0xffffff00u)) | (1 + (d1 > d2 ? d1 : d2)));
            parent[nnodes] = -1;
            int z = ++nheap; heap[z] = nnodes;
            int t = heap[z];
            while (weight[t] < weight[heap[z >> 1]]) { heap[z] = heap[z >> 1]; z >>= 1; }
            heap[z] = t;
        }
        int too_long = 0;
        for (int i = 1; i <= alpha; i++) {
            int j = 0, k = i;
            while (parent[k] >= 0) { k = parent[k]; j++; }
            len[i - 1] = (uint8_t)j;
            if (j > max_len) too_long = 1;
        }
        if (!too_long) break;
        for (int i = 1; i <= alpha; i++) { int j = weight[i] >> 8; j = 1 + j / 2; weight[i] = j << 8; }
    }
}

/* --- bit writer (bz/compress.c:37-98) -------------------------------- */
typedef struct { uint8_t *p; uint64_t nbits, cap; } bitw_t;
static void bw_put(bitw_t *w, int n, uint32_t v)
{
    if ((w->nbits + (uint64_t)n + 7) / 8 + 8 > w->cap) {
        uint64_t nc = w->cap ? w->cap * 2 : 1 << 16;
        w->p = realloc(w->p, nc); memset(w->p + w->cap, 0, nc - w->cap); w->cap = nc;
    }
    for (int i = n - 1; i >= 0; i--) {
        if ((v >> i) & 1) w->p[w->nbits >> 3] |= (uint8_t)(0x80u >> (w->nbits & 7));
        w->nbits++;
    }
}

/* selection result, exposed for stage-wise parity checks */
typedef struct {
    int32_t n_groups, n_selectors, alpha;
    uint8_t len[6][258];
    uint8_t selector[18002 + 8];
} s3o_huff_t;

/* bz/compress.c:239-452: table selection */
S3O_API void s3o_huff_select(const uint16_t *mtfv, int32_t nmtf, const int32_t *mtf_freq,
                             int32_t n_in_use, s3o_huff_t *h)
{
    int alpha = n_in_use + 2;
    int ng = nmtf < 200 ? 2 : nmtf < 600 ? 3 : nmtf < 1200 ? 4 : nmtf < 2400 ? 5 : 6;
    memset(h, 0, sizeof *h);
    h->alpha = alpha; h->n_groups = ng;
    for (int t = 0; t < 6; t++) for (int v = 0; v < alpha; v++) h->len[t][v] = 15;
    {   /* initial partition, :280-317 */
        int npart = ng, remf = nmtf, gs = 0;
        while (npart > 0) {
            int tfreq = remf / npart, ge = gs - 1, afreq = 0;
            while (afreq < tfreq && ge < alpha - 1) { ge++; afreq += mtf_freq[ge]; }
            if (ge > gs && npart != ng && npart != 1 && ((ng - npart) % 2 == 1)) { afreq -= mtf_freq[ge]; ge--; }
            for (int v = 0; v < alpha; v++) h->len[npart - 1][v] = (v >= gs && v <= ge) ? 0 : 15;
            npart--; gs = ge + 1; remf -= afreq;
        }
    }
    static int32_t rfreq[6][258];
    int nsel = 0;
    for (int iter = 0; iter < 4; iter++) {
        memset(rfreq, 0, sizeof rfreq);
        nsel = 0;
        for (int gs = 0; gs < nmtf; gs += 50) {
            int ge = gs + 49; if (ge >= nmtf) ge = nmtf - 1;
            uint16_t cost[6] = {0, 0, 0, 0, 0, 0};
            for (int i = gs; i <= ge; i++) for (int t = 0; t < ng; t++) cost[t] = (uint16_t)(cost[t] + h->len[t][mtfv[i]]);
            int bt = -1, bc = 999999999;
            for (int t = 0; t < ng; t++) if (cost[t] < bc) { bc = cost[t]; bt = t; }
            h->selector[nsel++] = (uint8_t)bt;
            for (int i = gs; i <= ge; i++) rfreq[bt][mtfv[i]]++;
        }
        for (int t = 0; t < ng; t++) hb_make_lengths(h->len[t], rfreq[t], alpha, 17);
    }
    h->n_selectors = nsel;
}

/* bz/compress.c:461-598 + bz/huffman.c:152-166: emission of one block body */
static void huff_emit(bitw_t *w, const uint16_t *mtfv, int32_t nmtf, const uint8_t *in_use, const s3o_huff_t *h)
{
    int ng = h->n_groups, alpha = h->alpha, nsel = h->n_selectors;
    static int32_t code[6][258];
    for (int t = 0; t < ng; t++) {
        int mn = 32, mx = 0;
        for (int i = 0; i < alpha; i++) { if (h->len[t][i] > mx) mx = h->len[t][i]; if (h->len[t][i] < mn) mn = h->len[t][i]; }
        int vec = 0;
        for (int n = mn; n <= mx; n++) { for (int i = 0; i < alpha; i++) if (h->len[t][i] == n) code[t][i] = vec++; vec <<= 1; }
    }
    /* mapping table :495-511 */
    int used16[16];
    for (int i = 0; i < 16; i++) { used16[i] = 0; for (int j = 0; j < 16; j++) if (in_use[i * 16 + j]) used16[i] = 1; }
    for (int i = 0; i < 16; i++) bw_put(w, 1, (uint32_t)used16[i]);
    for (int i = 0; i < 16; i++) if (used16[i]) for (int j = 0; j < 16; j++) bw_put(w, 1, in_use[i * 16 + j] ? 1 : 0);
    /* selectors, MTF + unary :462-478, :519-524 */
    bw_put(w, 3, (uint32_t)ng); bw_put(w, 15, (uint32_t)nsel);
    {
        uint8_t pos[6]; for (int i = 0; i < ng; i++) pos[i] = (uint8_t)i;
        for (int i = 0; i < nsel; i++) {
            uint8_t v = h->selector[i]; int j = 0; uint8_t carry = pos[0];
            while (carry != v) { j++; uint8_t t = pos[j]; pos[j] = carry; carry = t; }
            pos[0] = carry;
            for (int k = 0; k < j; k++) bw_put(w, 1, 1);
            bw_put(w, 1, 0);
        }
    }
    /* coding tables, delta coded :531-539 */
    for (int t = 0; t < ng; t++) {
        int cur = h->len[t][0];
        bw_put(w, 5, (uint32_t)cur);
        for (int i = 0; i < alpha; i++) {
            while (cur < h->len[t][i]) { bw_put(w, 2, 2); cur++; }
            while (cur > h->len[t][i]) { bw_put(w, 2, 3); cur--; }
            bw_put(w, 1, 0);
        }
    }
    /* symbols :545-594 */
    int sc = 0;
    for (int gs = 0; gs < nmtf; gs += 50, sc++) {
        int ge = gs + 49; if (ge >= nmtf) ge = nmtf - 1;
        int t = h->selector[sc];
        for (int i = gs; i <= ge; i++) bw_put(w, h->len[t][mtfv[i]], (uint32_t)code[t][mtfv[i]]);
    }
}

/* One whole bzip2 stream (bz/compress.c:602-667).  Returns bytes written or -1. */
S3O_API int64_t s3o_bz_compress(const uint8_t *in, uint64_t n, int level, uint8_t *out, uint64_t out_cap)
{
    blocklist_t L = {0};
    rle1_cut(in, n, level, &L);
    bitw_t w = {0};
    bw_put(&w, 8, 'B'); bw_put(&w, 8, 'Z'); bw_put(&w, 8, 'h'); bw_put(&w, 8, (uint32_t)('0' + level));
    uint32_t combined = 0;
    uint32_t *ptr = malloc((100000u * (size_t)level + 16) * sizeof *ptr);
    uint16_t *mtfv = malloc((100000u * (size_t)level + 16) * sizeof *mtfv);
    for (size_t k = 0; k < L.n; k++) {
        s3o_block_t *b = &L.b[k];
        if (b->nblock == 0) { free(b->data); continue; }   /* empty input: header+trailer only */
        uint32_t crc = ~b->crc;
        combined = ((combined << 1) | (combined >> 31)) ^ crc;
        int32_t orig = s3o_bwt(b->data, (int32_t)b->nblock, ptr);
        bw_put(&w, 24, 0x314159); bw_put(&w, 24, 0x265359);
        bw_put(&w, 32, crc); bw_put(&w, 1, 0); bw_put(&w, 24, (uint32_t)orig);
        int32_t freq[258], nu;
        int32_t nmtf = s3o_mtf(b->data, (int32_t)b->nblock, ptr, b->in_use, mtfv, freq, &nu);
        s3o_huff_t h;
        s3o_huff_select(mtfv, nmtf, freq, nu, &h);
        huff_emit(&w, mtfv, nmtf, b->in_use, &h);
        free(b->data);
    }
    bw_put(&w, 24, 0x177245); bw_put(&w, 24, 0x385090); bw_put(&w, 32, combined);
    int64_t nbytes = (int64_t)((w.nbits + 7) / 8);
    int64_t rc = nbytes;
    if ((uint64_t)nbytes > out_cap) rc = -1; else memcpy(out, w.p, (size_t)nbytes);
    free(w.p); free(ptr); free(mtfv); free(L.b);
    return rc;
}

/* stage-wise view of RLE1 + cut: fills up to cap descriptors; returns #blocks.
 * rle_out (may be NULL) receives the concatenated post-RLE1 bytes. */
typedef struct { uint64_t in_start, in_end; uint32_t nblock, crc; uint8_t in_use[256]; } s3o_blockdesc_t;
S3O_API int64_t s3o_rle1_blocks(const uint8_t *in, uint64_t n, int level, s3o_blockdesc_t *d, uint64_t cap,
                                uint8_t *rle_out, uint64_t rle_cap)
{
    blocklist_t L = {0};
    rle1_cut(in, n, level, &L);
    uint64_t off = 0; int64_t rc = (int64_t)L.n;
    for (size_t k = 0; k < L.n; k++) {
        s3o_block_t *b = &L.b[k];
        if (k < cap) {
            d[k].in_start = b->in_start; d[k].in_end = b->in_end; d[k].nblock = b->nblock; d[k].crc = ~b->crc;
            memcpy(d[k].in_use, b->in_use, 256);
        } else rc = -1;
        if (rle_out) { if (off + b->nblock <= rle_cap) memcpy(rle_out + off, b->data, b->nblock); else rc = -1; }
        off += b->nblock;
        free(b->data);
    }
    free(L.b);
    return rc;
}
