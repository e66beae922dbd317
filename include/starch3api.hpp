// starch3api.hpp -- host-side C++ surface of the B200-native starch3 hot path.
//
// Same class, member and function names as the reference header
// (/root/reference/include/starch3api.hpp:99-149, :158-581, :921), so a client written
// against the reference -- in particular its main(), /root/reference/src/starch3.cpp:14-70 --
// compiles and links unchanged.  What sits behind the names is different:
//
//   * the compression stream (reference: a bz_stream that is initialised and never fed,
//     hpp:819-855) is a device context of the C ABI in include/starch3_b200.h;
//   * the four pthread entry points (hpp:158, :201, :347, :371) keep their signatures, but
//     the one-line-at-a-time mailbox is gone: produce_line reads the whole input in bulk,
//     consume_line hands it to process_tf_buffer, update_chr / consume_tf_buffer return at once;
//   * process_tf_buffer (hpp:393-407), the seam where the reference prints the transformed
//     chromosome, submits the batch to the GPU (tokenise + transform + per-chromosome bzip2)
//     and writes the archive (ARCHIVE_FORMAT.md) after the magic bytes;
//   * there is no CPU implementation of the path: the per-line helpers
//     update_transformation_state / append_tf_line_to_buffer exist for link compatibility and
//     refuse to run.
//
// Error behaviour follows the reference: "Error: ..." on stderr, then std::exit with an
// errno-style code (ENOMEM hpp:176, ENODATA :733, ENOSYS :779, EINVAL :842).
#ifndef STARCH3_H_
#define STARCH3_H_

#include <string>
#include <vector>
#include <new>
#include <cstdio>
#include <cstdlib>
#include <cerrno>
#include <cinttypes>
#include <cstring>
#include <getopt.h>
#include <unistd.h>
#include <sys/stat.h>
#include <pthread.h>
#include "starch3_b200.h"

namespace starch3
{
    class Starch
    {
    public:
        typedef enum compression_method { k_bzip2 = 0, k_gzip, k_compression_method_undefined } compression_method_t;
        typedef enum bed_token { k_chromosome_token = 0, k_start_token, k_stop_token, k_remainder_token, k_bed_token_undefined } bed_token_t;

        // Kept for source compatibility (hpp:37-63).  On this path the fields are filled per
        // chromosome from the device results, not per line.
        typedef struct bed {
            char* chr; size_t chr_capacity; char* start_str; size_t start_str_capacity; int64_t start;
            char* stop_str; size_t stop_str_capacity; int64_t stop; char* rem; size_t rem_capacity; int token;
        } bed_t;
        typedef struct transform_state {
            int64_t line_count; char* last_chr; int64_t last_start; int64_t last_stop; int64_t last_coord_diff;
            char* current_chr; int64_t current_start; int64_t current_stop; int64_t current_coord_diff;
            int64_t base_count_unique; int64_t base_count_nonunique;
        } transform_state_t;

        // The mailbox of the reference (hpp:66-88) with the same member names; in_line now
        // holds the WHOLE input (in_line_size bytes) and tf_buffer the finished archive.
        typedef struct shared_buffer {
            pthread_mutex_t lock;
            pthread_cond_t new_line_is_available;
            pthread_cond_t new_line_is_empty;
            pthread_cond_t new_chromosome_is_available;
            pthread_cond_t new_tf_buffer_is_available;
            char* in_line; size_t in_line_capacity; size_t in_line_size;
            int next_in; int next_out;
            bool is_new_line_available; bool is_new_chromosome_available; bool is_new_tf_buffer_available; bool is_eof;
            FILE* in_stream;
            bed_t* bed; transform_state_t* tf_state;
            char* tf_line; size_t tf_line_capacity;
            char* tf_buffer; size_t tf_buffer_capacity; size_t tf_buffer_size;
        } shared_buffer_t;

    private:
        std::string _input_fn;
        std::string _note;
        s3g_ctx* _bz_stream_ptr;           // the compression stream: a device context
        FILE* _in_stream;
        FILE* _out_stream;
        compression_method_t _compression_method;
        unsigned char _header_magic_bytes[4];
        int _device;
        int _block_size_100k;
        bool _unstarch;                    // --unstarch: the decoder path (archive in, BED out)
        std::vector<int> _devices;         // --devices=0,1,..: one archive from several GPUs (s3g_multi_compress_bed)
        std::vector<s3g_ctx*> _more_ctx;   // the contexts of the devices after the first

    public:
        Starch();
        ~Starch();

        pthread_t produce_line_thread;
        pthread_t consume_line_thread;
        pthread_t update_chr_thread;
        pthread_t consume_tf_buffer_thread;

        shared_buffer_t buffer;

        void initialize_shared_buffer(shared_buffer_t* b);
        void delete_shared_buffer(shared_buffer_t* b);
        FILE* get_in_stream(void) { return _in_stream; }
        void initialize_in_stream(void);
        void set_in_stream(FILE* s) { _in_stream = s; }
        std::string get_input_fn(void) { return _input_fn; }
        void set_input_fn(std::string s);
        void set_out_stream(FILE* s) { _out_stream = s; }
        FILE* get_out_stream(void) { return _out_stream; }
        void initialize_out_stream(void);
        void initialize_out_compression_stream(void);
        void delete_out_compression_stream(void);
        std::string get_note(void) { return _note; }
        void set_note(std::string s) { _note = s; }
        compression_method_t get_compression_method(void) { return _compression_method; }
        void set_compression_method(compression_method_t t) { _compression_method = t; }
        void initialize_bz_stream_ptr(void);
        void setup_bz_stream_callbacks(Starch* h);
        void delete_bz_stream_ptr(void);
        void bzip2_block_close_callback(void);
        void initialize_command_line_options(int argc, char** argv);
        void test_stdin_availability(void);
        void initialize_header_magic_bytes(void);
        void print_usage(FILE* wo_stream);
        void print_version(FILE* wo_stream);
        // additions of this implementation
        s3g_ctx* get_device_context(void) { return _bz_stream_ptr; }
        void set_device(int d) { _device = d; }
        void set_block_size_100k(int k) { _block_size_100k = k; }
        int get_block_size_100k(void) { return _block_size_100k; }
        void set_devices(const std::vector<int>& d) { _devices = d; if (!d.empty()) _device = d[0]; }
        const std::vector<int>& get_devices(void) { return _devices; }
        std::vector<s3g_ctx*>& get_more_contexts(void) { return _more_ctx; }
        void set_unstarch(bool u) { _unstarch = u; }
        bool get_unstarch(void) { return _unstarch; }

        static const compression_method_t client_starch_default_compression_method;
        static const std::string client_name;
        static const std::string client_version;
        static const std::string client_authors;
        std::string get_client_starch_opt_string(void);
        struct option* get_client_starch_long_options(void);
        std::string get_client_starch_name(void);
        std::string get_client_starch_version(void);
        std::string get_client_starch_authors(void);
        std::string get_client_starch_usage(void);
        std::string get_client_starch_description(void);
        std::string get_client_starch_io_options(void);
        std::string get_client_starch_general_options(void);

        static const int in_line_initial_length = 1 << 20;
        static const int in_field_initial_length = 128;
        static const int tf_line_initial_length = 1024;
        static const int tf_buffer_initial_length = 1024;
        static const char field_delimiter = '\t';
        static const char line_delimiter = '\n';

        static void fail(int code, const char* msg) { std::fprintf(stderr, "Error: %s\n", msg); std::exit(code); }

        // ---- thread entry points (pthread signature, as hpp:158 / :201 / :347 / :371) ----
        // Line reader.  A regular file of moderate size is read whole (the batch call then overlaps its upload with the
        // kernels); stdin, pipes, files beyond 4 GiB and S3G_STREAM=1 go through the bounded-memory entry instead: 64 MiB
        // at a time into s3g_stream_write (the getc loop of hpp:158-199 becomes fread + one call per piece).
        static void* produce_line(void* arg);
        // Waits for the input, then runs the whole path once.
        static void* consume_line(void* arg) {
            shared_buffer_t* sb = static_cast<shared_buffer_t*>(arg);
            pthread_mutex_lock(&sb->lock);
            while (!sb->is_new_line_available) pthread_cond_wait(&sb->new_line_is_available, &sb->lock);
            sb->is_new_tf_buffer_available = true;
            process_tf_buffer(sb);
            sb->is_new_tf_buffer_available = false;
            sb->is_new_line_available = false;
            pthread_cond_broadcast(&sb->new_tf_buffer_is_available);
            pthread_mutex_unlock(&sb->lock);
            return NULL;
        }
        // Chromosome boundaries and the per-chromosome hand-off happen on the device.
        static void* update_chr(void*) { return NULL; }
        static void* consume_tf_buffer(void*) { return NULL; }

        // The seam (hpp:393): one batch call does tokenise + transform + compress for all chromosomes.
        static void process_tf_buffer(shared_buffer_t* sb);

        static void append_tf_line_to_buffer(shared_buffer_t*) { fail(ENOSYS, "append_tf_line_to_buffer: the transform runs on the GPU; no per-line CPU path exists"); }
        static void update_transformation_state(shared_buffer_t*) { fail(ENOSYS, "update_transformation_state: the transform runs on the GPU; no per-line CPU path exists"); }

        static void initialize_transformation_state(transform_state_t** tfs) { std::memset(*tfs, 0, sizeof(**tfs)); }
        static void reset_transformation_state(transform_state_t** tfs) {
            char* last = (*tfs)->last_chr; char* cur = (*tfs)->current_chr;
            std::memset(*tfs, 0, sizeof(**tfs));
            (*tfs)->last_chr = last; (*tfs)->current_chr = cur;      // names survive a reset (hpp:523-532)
        }
        static void delete_transformation_state(transform_state_t** tfs) { free((*tfs)->last_chr); free((*tfs)->current_chr); }
        static inline void update_str(char** dest, char* src) {
            free(*dest); *dest = NULL;
            if (!src) return;
            size_t n = std::strlen(src) + 1;
            *dest = static_cast<char*>(std::malloc(n));
            if (!*dest) fail(ENOMEM, "Not enough memory for allocation of transformation buffer chromosome");
            std::memcpy(*dest, src, n);
        }
        static inline short n_digits(int64_t i) {   // digits of |i|, sign not counted (hpp:559-581)
            uint64_t a = i < 0 ? 0 - static_cast<uint64_t>(i) : static_cast<uint64_t>(i);
            short d = 1;
            while (a >= 10 && d < 19) { a /= 10; d++; }
            return d;
        }
        static void bzip2_block_close_static_callback(void* s) { reinterpret_cast<Starch*>(s)->bzip2_block_close_callback(); }
    };

    extern Starch* self;

    inline void Starch::initialize_shared_buffer(shared_buffer_t* sb) {
        std::memset(sb, 0, sizeof(*sb));
        sb->in_line = static_cast<char*>(malloc(in_line_initial_length));
        sb->bed = static_cast<bed_t*>(calloc(1, sizeof(bed_t)));
        sb->tf_state = static_cast<transform_state_t*>(calloc(1, sizeof(transform_state_t)));
        if (!sb->in_line || !sb->bed || !sb->tf_state) fail(ENOMEM, "Not enough memory for shared_buffer_t");
        sb->in_line_capacity = in_line_initial_length;
        pthread_mutex_init(&sb->lock, NULL);
        pthread_cond_init(&sb->new_line_is_available, NULL);
        pthread_cond_init(&sb->new_line_is_empty, NULL);
        pthread_cond_init(&sb->new_chromosome_is_available, NULL);
        pthread_cond_init(&sb->new_tf_buffer_is_available, NULL);
        sb->in_stream = get_in_stream();
    }

    inline void Starch::delete_shared_buffer(shared_buffer_t* sb) {
        pthread_mutex_destroy(&sb->lock);
        pthread_cond_destroy(&sb->new_line_is_available);
        pthread_cond_destroy(&sb->new_line_is_empty);
        pthread_cond_destroy(&sb->new_chromosome_is_available);
        pthread_cond_destroy(&sb->new_tf_buffer_is_available);
        if (sb->in_stream) fclose(sb->in_stream);
        free(sb->in_line); sb->in_line = NULL;
        if (sb->tf_state) { delete_transformation_state(&sb->tf_state); free(sb->tf_state); sb->tf_state = NULL; }
        free(sb->bed); sb->bed = NULL;
        free(sb->tf_line); sb->tf_line = NULL;
        free(sb->tf_buffer); sb->tf_buffer = NULL; sb->tf_buffer_size = sb->tf_buffer_capacity = 0;
    }

    inline void Starch::initialize_in_stream(void) {
        FILE* fp = get_input_fn().empty() ? stdin : fopen(get_input_fn().c_str(), "r");
        if (!fp) fail(ENODATA, "Input file handle could not be created");
        set_in_stream(fp);
    }

    inline void Starch::set_input_fn(std::string s) {
        struct stat st;
        if (stat(s.c_str(), &st) != 0) { std::fprintf(stderr, "Error: Input file does not exist (%s)\n", s.c_str()); std::exit(ENODATA); }
        _input_fn = s;
    }

    inline void Starch::initialize_out_stream(void) {
        set_out_stream(stdout);
        if (!_unstarch) std::fwrite(_header_magic_bytes, 1, 4, stdout);     // the only bytes the reference ever writes (hpp:765-769)
    }

    inline void Starch::initialize_out_compression_stream(void) {
        switch (get_compression_method()) {
        case k_bzip2: initialize_bz_stream_ptr(); setup_bz_stream_callbacks(this); break;
        case k_gzip: fail(ENOSYS, "This method is unsupported at this time");
        default: fail(ENOSYS, "This method is undefined");
        }
    }

    inline void Starch::delete_out_compression_stream(void) {
        switch (get_compression_method()) {
        case k_bzip2: delete_bz_stream_ptr(); break;
        case k_gzip: fail(ENOSYS, "This method is unsupported at this time");
        default: fail(ENOSYS, "This method is undefined");
        }
    }

    inline void Starch::initialize_bz_stream_ptr(void) {
        int rc = s3g_init(_device, &_bz_stream_ptr);
        if (rc == S3G_E_NOMEM) { std::fprintf(stderr, "Error: bzip2 initialization failed - insufficient memory\n"); std::exit(EINVAL); }
        if (rc != S3G_OK) { std::fprintf(stderr, "Error: bzip2 initialization failed - %s\n", s3g_last_error()); std::exit(EINVAL); }
    }

    inline void Starch::setup_bz_stream_callbacks(Starch*) { /* the stream-end hook (bz/bzlib.c:470) has no per-stream work left: metadata comes back with the batch */ }

    inline void Starch::delete_bz_stream_ptr(void) {
        for (size_t k = 0; k < _more_ctx.size(); k++) s3g_destroy(_more_ctx[k]);
        _more_ctx.clear();
        if (!_bz_stream_ptr) return;
        s3g_destroy(_bz_stream_ptr);
        _bz_stream_ptr = NULL;
    }

    inline void Starch::bzip2_block_close_callback(void) { }

    inline void Starch::test_stdin_availability(void) {
        struct stat st;
        if (fstat(STDIN_FILENO, &st) == -1) {
            int e = errno;
            std::fprintf(stderr, "Error: fstat() call failed (%s)", e == EBADF ? "EBADF" : (e == EIO ? "EIO" : "EOVERFLOW"));
            print_usage(stderr);
            std::exit(e);
        }
        if (S_ISCHR(st.st_mode) && !S_ISREG(st.st_mode) && get_input_fn().empty()) {
            std::fprintf(stderr, "Error: No input is specified; please redirect or pipe in formatted data, or specify filename\n");
            print_usage(stderr);
            std::exit(ENODATA);
        }
    }

    inline void Starch::initialize_header_magic_bytes(void) {
        static const unsigned char mb[4] = { 0xca, 0x5c, 0xad, 0x1a };
        std::memcpy(_header_magic_bytes, mb, 4);
    }

    inline Starch::Starch() : _bz_stream_ptr(NULL), _in_stream(NULL), _out_stream(NULL), _device(0), _block_size_100k(9), _unstarch(false) {
        set_note(std::string());
        set_compression_method(k_compression_method_undefined);
        initialize_header_magic_bytes();
        std::memset(&buffer, 0, sizeof(buffer));
    }

    inline Starch::~Starch() { }

    inline void* Starch::produce_line(void* arg) {
        shared_buffer_t* sb = static_cast<shared_buffer_t*>(arg);
        pthread_mutex_lock(&sb->lock);
        struct stat st;
        bool streaming = std::getenv("S3G_STREAM") != NULL;
        if (fstat(fileno(sb->in_stream), &st) != 0 || !S_ISREG(st.st_mode) || st.st_size > (off_t)(4ll << 30)) streaming = true;
        if (self && self->get_unstarch()) streaming = false;
        if (streaming && self && self->get_device_context()) {
            const size_t piece = 64u << 20;
            if (sb->in_line_capacity < piece) {
                char* grown = static_cast<char*>(realloc(sb->in_line, piece));
                if (!grown) fail(ENOMEM, "Not enough memory for reallocation of shared_buffer_t line character buffer");
                sb->in_line = grown; sb->in_line_capacity = piece;
            }
            size_t range = 0;
            if (const char* e = std::getenv("S3G_STREAM_RANGE")) range = static_cast<size_t>(std::atoll(e));
            int rc = s3g_stream_begin(self->get_device_context(), self->get_block_size_100k(), self->get_note().c_str(), range);
            for (size_t got; rc == S3G_OK && (got = fread(sb->in_line, 1, piece, sb->in_stream)) > 0;)
                rc = s3g_stream_write(self->get_device_context(), reinterpret_cast<const uint8_t*>(sb->in_line), got);
            if (rc != S3G_OK) {
                std::fprintf(stderr, "Error: %s\n", s3g_last_error());
                std::exit(rc == S3G_E_NOMEM ? ENOMEM : rc == S3G_E_MALFORMED ? EINVAL : rc == S3G_E_CUDA ? ENODEV : EINVAL);
            }
            sb->in_line_size = 0;
            sb->next_in = 1;                      // the input went through s3g_stream_write
        } else {
            size_t used = 0;
            for (;;) {
                if (used == sb->in_line_capacity) {
                    char* grown = static_cast<char*>(realloc(sb->in_line, sb->in_line_capacity * 2));
                    if (!grown) fail(ENOMEM, "Not enough memory for reallocation of shared_buffer_t line character buffer");
                    sb->in_line = grown; sb->in_line_capacity *= 2;
                }
                size_t got = fread(sb->in_line + used, 1, sb->in_line_capacity - used, sb->in_stream);
                used += got;
                if (got == 0) break;
            }
            sb->in_line_size = used;
        }
        sb->is_eof = true;
        sb->is_new_line_available = true;
        pthread_cond_broadcast(&sb->new_line_is_available);
        pthread_mutex_unlock(&sb->lock);
        return NULL;
    }

    inline void Starch::process_tf_buffer(shared_buffer_t* sb) {
        Starch* me = self;
        if (!me || !me->get_device_context()) fail(EINVAL, "compression stream is not initialised");
        FILE* out = me->get_out_stream() ? me->get_out_stream() : stdout;
        if (me->get_unstarch()) {
            // the decoder path: sb->in_line holds an archive, the BED text goes out
            uint64_t need = 0;
            int rc = s3g_decompress_archive(me->get_device_context(), reinterpret_cast<const uint8_t*>(sb->in_line), sb->in_line_size, NULL, 0, &need, NULL);
            uint8_t* bed = rc == S3G_OK ? static_cast<uint8_t*>(std::malloc(need ? need : 1)) : NULL;
            if (rc == S3G_OK && !bed) fail(ENOMEM, "Not enough memory for the decoded BED text");
            if (rc == S3G_OK) rc = s3g_decompress_archive(me->get_device_context(), reinterpret_cast<const uint8_t*>(sb->in_line), sb->in_line_size, bed, need, &need, NULL);
            if (rc != S3G_OK) { std::fprintf(stderr, "Error: %s\n", s3g_last_error()); std::exit(rc == S3G_E_NOMEM ? ENOMEM : rc == S3G_E_CUDA ? ENODEV : EINVAL); }
            if (need && std::fwrite(bed, 1, need, out) != need) fail(EIO, "could not write the BED text");
            std::fflush(out);
            std::free(bed);
            return;
        }
        s3g_result res;
        int rc;
        if (sb->next_in) rc = s3g_stream_end(me->get_device_context(), &res);
        else if (me->get_devices().size() > 1) {
            // one archive from several GPUs: a context per device, the phases driven by s3g_multi_compress_bed
            std::vector<s3g_ctx*> ctxs(1, me->get_device_context());
            rc = S3G_OK;
            for (size_t k = 1; k < me->get_devices().size() && rc == S3G_OK; k++) {
                s3g_ctx* c = NULL;
                rc = s3g_init(me->get_devices()[k], &c);
                if (rc == S3G_OK) { ctxs.push_back(c); me->get_more_contexts().push_back(c); }
            }
            if (rc == S3G_OK)
                rc = s3g_multi_compress_bed(ctxs.data(), static_cast<int>(ctxs.size()), reinterpret_cast<const uint8_t*>(sb->in_line), sb->in_line_size,
                                            me->get_block_size_100k(), me->get_note().c_str(), &res);
        } else
            rc = s3g_compress_bed(me->get_device_context(), reinterpret_cast<const uint8_t*>(sb->in_line), sb->in_line_size,
                                  me->get_block_size_100k(), me->get_note().c_str(), &res);
        if (rc != S3G_OK) {
            std::fprintf(stderr, "Error: %s\n", s3g_last_error());
            std::exit(rc == S3G_E_NOMEM ? ENOMEM : rc == S3G_E_MALFORMED ? EINVAL : rc == S3G_E_CUDA ? ENODEV : EINVAL);
        }
        if (res.dropped_tail_bytes)
            std::fprintf(stderr, "Warning: the last line is not newline-terminated; %llu byte(s) ignored, as the reference does\n",
                         static_cast<unsigned long long>(res.dropped_tail_bytes));
        if (res.unsorted_lines)
            std::fprintf(stderr, "Warning: %llu element(s) start before the previous element of their chromosome: the input is not sorted\n",
                         static_cast<unsigned long long>(res.unsorted_lines));
        if (res.reappearing_chroms)
            std::fprintf(stderr, "Warning: %llu chromosome stream(s) repeat an earlier chromosome name: the input is not sorted by chromosome\n",
                         static_cast<unsigned long long>(res.reappearing_chroms));
        if (res.crlf_lines)
            std::fprintf(stderr, "Warning: %llu line(s) end in CR LF; the carriage return is kept as part of the line's last field\n",
                         static_cast<unsigned long long>(res.crlf_lines));
        sb->tf_state->line_count = static_cast<int64_t>(res.n_lines);
        // the magic bytes are already out (initialize_out_stream); write the rest of the archive
        if (res.archive_size > 4 && std::fwrite(res.archive + 4, 1, res.archive_size - 4, out) != res.archive_size - 4)
            fail(EIO, "could not write the archive");
        std::fflush(out);
        s3g_result_free(&res);
    }
}

#endif // STARCH3_H_
