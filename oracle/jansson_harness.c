/* jansson_harness.c -- TEST INFRASTRUCTURE.  Prints the archive metadata object (ARCHIVE_FORMAT.md) with the
 * reference's own vendored jansson 2.9 (third-party/jansson-2.9.tar.gz; the reference includes it at
 * starch3api.hpp:17 and links it at makefile:32 but never calls it), so that the hand-written writers
 * (starch3_b200/csrc/api.cu build_header, oracle/oracle.py archive) are pinned against the real
 * json_dumps(obj, JSON_COMPACT) instead of a reading of dump.c.  Built by oracle/build_ref.sh into
 * oracle/_ref/libs3jansson.so; loaded only by tests/. */
#include <jansson.h>
#include <stdint.h>
#include <string.h>
#include <stdlib.h>

/* fields: 7 int64 per stream: offset, size, lines, blocks, transformedBytes, nonUniqueBases, uniqueBases.
 * names: concatenated, name_len[i] bytes each.  Returns the text length, -1 if jansson refuses (e.g. a
 * name that is not valid UTF-8), -2 if cap is too small. */
__attribute__((visibility("default")))
int64_t s3ref_archive_header(int level, const char *note, uint64_t note_len, uint64_t n_streams, const uint8_t *names,
                             const uint32_t *name_len, const int64_t *fields, char *out, uint64_t cap)
{
    json_t *root = json_object(), *arc = json_object(), *ver = json_object(), *streams = json_array();
    json_object_set_new(ver, "major", json_integer(3));
    json_object_set_new(ver, "minor", json_integer(0));
    json_object_set_new(ver, "revision", json_integer(0));
    json_object_set_new(arc, "type", json_string("starch"));
    json_object_set_new(arc, "version", ver);
    json_object_set_new(arc, "creator", json_string("starch3_b200"));
    json_object_set_new(arc, "compression", json_string("bzip2"));
    json_object_set_new(arc, "blockSize100k", json_integer(level));
    json_object_set_new(arc, "note", json_stringn_nocheck(note ? note : "", note ? note_len : 0));
    json_object_set_new(root, "archive", arc);
    static const char *keys[7] = {"offset", "size", "lines", "blocks", "transformedBytes", "nonUniqueBases", "uniqueBases"};
    uint64_t off = 0;
    for (uint64_t i = 0; i < n_streams; i++) {
        json_t *s = json_object();
        json_object_set_new(s, "chromosome", json_stringn_nocheck((const char *)names + off, name_len[i]));
        off += name_len[i];
        for (int k = 0; k < 7; k++) json_object_set_new(s, keys[k], json_integer((json_int_t)fields[i * 7 + k]));
        json_array_append_new(streams, s);
    }
    json_object_set_new(root, "streams", streams);
    char *txt = json_dumps(root, JSON_COMPACT);
    json_decref(root);
    if (!txt) return -1;
    int64_t len = (int64_t)strlen(txt);
    if ((uint64_t)len > cap) { free(txt); return -2; }
    memcpy(out, txt, (size_t)len);
    free(txt);
    return len;
}
