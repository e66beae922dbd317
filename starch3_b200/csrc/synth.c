/*
 * synth.c -- deterministic synthetic sorted-BED generators for the shapes named in
 * BASELINE.json `configs`.  Host-only helper used by bench.py and the
 * tests to build inputs; not part of the compression path.  splitmix64, fixed
 * seeds, so every machine produces the same bytes.
 *
 *   cfg 1  BED3, one chromosome (chr1), len U[50,500], start step U[1,200] + U[0,len]
 *   cfg 2  BED6 hg38-shaped: 24 chromosomes (lexicographic order) in proportion to
 *          their hg38 lengths, len U[100,300], id "id-<serial>", score U[0,1000], strand +/-
 *   cfg 3  dense BED3, chr1, constant length 20, gap U[1,12]   (variant 1: constant gap 5)
 *   cfg 4  BED6, chr1, gap U[1000,200000], len U[5000,500000], 12 random base-62
 *          characters as name, score "%.6f" of U[0,1000), strand from "+-."
 *   cfg 5  whole-genome mix: 24 chromosomes, style of chromosome c = c mod 3 of {2,3,4}
 */
#include <stdint.h>
#include <string.h>

#define API __attribute__((visibility("default")))

static inline uint64_t sm64(uint64_t *s)
{
    uint64_t z = (*s += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}
static inline uint64_t urand(uint64_t *s, uint64_t lo, uint64_t hi) { return lo + sm64(s) % (hi - lo + 1); }

static inline uint8_t *put_u64(uint8_t *p, uint64_t v)
{
    char b[24]; int n = 0;
    do { b[n++] = (char)('0' + v % 10); v /= 10; } while (v);
    while (n) *p++ = (uint8_t)b[--n];
    return p;
}
static inline uint8_t *put_str(uint8_t *p, const char *s) { while (*s) *p++ = (uint8_t)*s++; return p; }

static const char *CHR[24] = {"chr1", "chr10", "chr11", "chr12", "chr13", "chr14", "chr15", "chr16", "chr17", "chr18", "chr19",
                              "chr2", "chr20", "chr21", "chr22", "chr3", "chr4", "chr5", "chr6", "chr7", "chr8", "chr9", "chrX", "chrY"};
static const uint64_t CHRLEN[24] = {248956422, 133797422, 135086622, 133275309, 114364328, 107043718, 101991189, 90338345, 83257441,
                                    80373285, 58617616, 242193529, 64444167, 46709983, 50818468, 198295559, 190214555, 181538259,
                                    170805979, 159345973, 145138636, 138394717, 156040895, 57227415};

typedef struct { uint64_t pos; uint64_t prev_len; uint64_t serial; } gen_t;

/* one line of the given style; returns the new write pointer */
static uint8_t *line(uint8_t *p, int style, int variant, const char *chr, uint64_t step_hi, gen_t *g, uint64_t *rs)
{
    static const char B62[] = "0123456789ABCDEFGHIJKLMNOPQRSTUVWXYZabcdefghijklmnopqrstuvwxyz";
    uint64_t start, len;
    switch (style) {
        case 1: start = g->pos + urand(rs, 1, 200) + urand(rs, 0, g->prev_len); len = urand(rs, 50, 500); break;
        case 2: start = g->pos + urand(rs, 1, step_hi); len = urand(rs, 100, 300); break;
        case 3: len = 20; start = g->pos + g->prev_len + (variant == 1 ? 5 : urand(rs, 1, 12)); break;
        default: start = g->pos + g->prev_len + urand(rs, 1000, 200000); len = urand(rs, 5000, 500000); break;
    }
    p = put_str(p, chr); *p++ = '\t';
    p = put_u64(p, start); *p++ = '\t';
    p = put_u64(p, start + len);
    if (style == 2) {
        *p++ = '\t'; p = put_str(p, "id-"); p = put_u64(p, ++g->serial);
        *p++ = '\t'; p = put_u64(p, urand(rs, 0, 1000));
        *p++ = '\t'; *p++ = (sm64(rs) & 1) ? '+' : '-';
    } else if (style == 4) {
        *p++ = '\t';
        for (int i = 0; i < 12; i++) *p++ = (uint8_t)B62[sm64(rs) % 62];
        *p++ = '\t';
        uint64_t micro = sm64(rs) % 1000000000ull;         /* U[0,1000) with 6 decimals */
        p = put_u64(p, micro / 1000000); *p++ = '.';
        uint64_t fr = micro % 1000000; char b[6];
        for (int i = 5; i >= 0; i--) { b[i] = (char)('0' + fr % 10); fr /= 10; }
        memcpy(p, b, 6); p += 6;
        *p++ = '\t'; *p++ = (uint8_t)"+-."[sm64(rs) % 3];
    }
    *p++ = '\n';
    g->pos = start; g->prev_len = len;
    return p;
}

/* Upper bound of bytes per line for sizing the output buffer. */
API uint64_t s3synth_max_line_bytes(int cfg) { (void)cfg; return 96; }

/* Writes n_lines lines; returns bytes written, or 0 if cap could be exceeded. */
API uint64_t s3synth_bed(int cfg, int variant, uint64_t n_lines, uint64_t seed, uint8_t *out, uint64_t cap)
{
    if (cap < n_lines * 96) return 0;
    uint64_t rs = seed * 0x2545F4914F6CDD1Dull + (uint64_t)cfg;
    uint8_t *p = out;
    gen_t g = {0, 0, 0};
    if (cfg == 1 || cfg == 3 || cfg == 4) {
        for (uint64_t i = 0; i < n_lines; i++) p = line(p, cfg, variant, "chr1", 0, &g, &rs);
        return (uint64_t)(p - out);
    }
    /* multi-chromosome: lines per chromosome in proportion to hg38 lengths */
    uint64_t total = 0;
    for (int c = 0; c < 24; c++) total += CHRLEN[c];
    uint64_t done = 0, acc = 0;
    for (int c = 0; c < 24; c++) {
        acc += CHRLEN[c];
        uint64_t upto = c == 23 ? n_lines : (uint64_t)((__uint128_t)n_lines * acc / total);
        uint64_t cnt = upto - done;
        done = upto;
        if (cnt == 0) continue;
        int style = cfg == 2 ? 2 : (c % 3 == 0 ? 2 : c % 3 == 1 ? 3 : 4);
        uint64_t step_hi = 2 * (CHRLEN[c] / cnt);
        if (step_hi < 2) step_hi = 2;
        g.pos = 0; g.prev_len = 0;
        for (uint64_t i = 0; i < cnt; i++) p = line(p, style, variant, CHR[c], step_hi, &g, &rs);
    }
    return (uint64_t)(p - out);
}
