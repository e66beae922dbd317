// starch3 -- command-line client of the B200-native hot path.
//
// Same call sequence as the reference client (/root/reference/src/starch3.cpp:14-70):
// options, stdin test, input stream, output stream (magic bytes), compression stream,
// shared buffer, four pthreads, join, teardown.  The work happens in
// Starch::process_tf_buffer (include/starch3api.hpp), i.e. on the GPU.
#include "starch3api.hpp"

const std::string starch3::Starch::client_name = "starch3";
const std::string starch3::Starch::client_version = "0.1-b200";
const std::string starch3::Starch::client_authors = "starch3_b200 (interface after Alex Reynolds and Shane Neph's starch3)";
const starch3::Starch::compression_method_t starch3::Starch::client_starch_default_compression_method = k_bzip2;

starch3::Starch* starch3::self = NULL;

int main(int argc, char** argv)
{
    starch3::Starch starch;
    starch3::self = &starch;
    starch.initialize_command_line_options(argc, argv);
    starch.test_stdin_availability();
    starch.initialize_in_stream();
    starch.initialize_out_stream();
    starch.initialize_out_compression_stream();
    starch.initialize_shared_buffer(&starch.buffer);
    pthread_create(&starch.produce_line_thread, NULL, starch3::Starch::produce_line, &starch.buffer);
    pthread_create(&starch.consume_line_thread, NULL, starch3::Starch::consume_line, &starch.buffer);
    pthread_create(&starch.update_chr_thread, NULL, starch3::Starch::update_chr, &starch.buffer);
    pthread_create(&starch.consume_tf_buffer_thread, NULL, starch3::Starch::consume_tf_buffer, &starch.buffer);
    pthread_join(starch.produce_line_thread, NULL);
    pthread_join(starch.consume_line_thread, NULL);
    pthread_join(starch.update_chr_thread, NULL);
    pthread_join(starch.consume_tf_buffer_thread, NULL);
    starch.delete_shared_buffer(&starch.buffer);
    starch.delete_out_compression_stream();
    return EXIT_SUCCESS;
}

std::string starch3::Starch::get_client_starch_opt_string(void) { return "n:bghvud:k:D:?"; }

struct option* starch3::Starch::get_client_starch_long_options(void)
{
    static struct option opts[] = {
        {"note", required_argument, NULL, 'n'},   {"bzip2", no_argument, NULL, 'b'},       {"gzip", no_argument, NULL, 'g'},
        {"help", no_argument, NULL, 'h'},         {"version", no_argument, NULL, 'v'},     {"device", required_argument, NULL, 'd'},
        {"block-size", required_argument, NULL, 'k'}, {"unstarch", no_argument, NULL, 'u'}, {"devices", required_argument, NULL, 'D'}, {NULL, no_argument, NULL, 0}};
    return opts;
}

void starch3::Starch::initialize_command_line_options(int argc, char** argv)
{
    int methods = 0, idx = 0, c;
    opterr = 0;
    while ((c = getopt_long(argc, argv, get_client_starch_opt_string().c_str(), get_client_starch_long_options(), &idx)) != -1) {
        switch (c) {
        case 'n': set_note(optarg); break;
        case 'b': set_compression_method(k_bzip2); methods++; break;
        case 'g': set_compression_method(k_gzip); methods++; break;
        case 'd': set_device(std::atoi(optarg)); break;
        case 'u': set_unstarch(true); break;
        case 'D': {
            std::vector<int> devs;
            for (const char* p = optarg; *p;) {
                char* e = NULL;
                long v = std::strtol(p, &e, 10);
                if (e == p || v < 0) fail(EINVAL, "--devices takes a comma-separated list of CUDA device numbers");
                devs.push_back(static_cast<int>(v));
                p = *e == ',' ? e + 1 : e;
                if (*e && *e != ',') fail(EINVAL, "--devices takes a comma-separated list of CUDA device numbers");
            }
            if (devs.empty() || devs.size() > 8) fail(EINVAL, "--devices takes 1 to 8 device numbers");
            set_devices(devs);
            break;
        }
        case 'k': {
            int k = std::atoi(optarg);
            if (k < 1 || k > 9) fail(EINVAL, "bzip2 initialization failed - incorrect parameters");
            set_block_size_100k(k);
            break;
        }
        case 'v': print_version(stdout); std::exit(EXIT_SUCCESS);
        case 'h':
        case '?': print_usage(stdout); std::exit(EXIT_SUCCESS);
        default: break;
        }
    }
    for (; optind < argc; optind++) {
        if (get_input_fn().empty()) set_input_fn(argv[optind]);
        else std::fprintf(stderr, "Warning: Ignoring additional input file [%s]\n", argv[optind]);
    }
    if (methods > 1) {
        std::fprintf(stderr, "Error: Only one compression method may be set\n");
        print_usage(stderr);
        std::exit(EXIT_FAILURE);
    }
    if (methods == 0) set_compression_method(client_starch_default_compression_method);
}

std::string starch3::Starch::get_client_starch_name(void) { return client_name; }
std::string starch3::Starch::get_client_starch_version(void) { return client_version; }
std::string starch3::Starch::get_client_starch_authors(void) { return client_authors; }
std::string starch3::Starch::get_client_starch_usage(void)
{
    return "\n  Usage:\n\n  $ starch3 [options] < input > output\n\n  Or:\n\n  $ starch3 [options] input > output\n";
}
std::string starch3::Starch::get_client_starch_description(void)
{
    return "  Compress sorted BED data to a starch3 archive (per-chromosome bzip2 streams) on an NVIDIA B200.\n";
}
std::string starch3::Starch::get_client_starch_io_options(void)
{
    return "  General Options:\n\n"
           "  --note=\"foo bar...\"   Append note to output archive metadata (optional)\n"
           "  --bzip2 | --gzip      Compression backend (bzip2 is the default; gzip is unsupported)\n"
           "  --device=N            CUDA device to use (default 0)\n"
           "  --devices=A,B,...     One archive from several GPUs (up to 8; byte ranges and bzip2 blocks are dealt to them)\n"
           "  --block-size=K        bzip2 block size in 100 kB units, 1..9 (default 9)\n"
           "  --unstarch            Decode: the input is a starch3 archive, the output its BED text\n";
}
std::string starch3::Starch::get_client_starch_general_options(void)
{
    return "  Process Flags:\n\n  --help                  Show this usage message\n  --version               Show binary version\n";
}

void starch3::Starch::print_usage(FILE* w)
{
    std::fprintf(w, "%s\n  version: %s\n  author:  %s\n%s\n%s\n%s\n%s\n", get_client_starch_name().c_str(), get_client_starch_version().c_str(),
                 get_client_starch_authors().c_str(), get_client_starch_usage().c_str(), get_client_starch_description().c_str(),
                 get_client_starch_io_options().c_str(), get_client_starch_general_options().c_str());
}

void starch3::Starch::print_version(FILE* w)
{
    std::fprintf(w, "%s\n  version: %s\n  author:  %s\n", get_client_starch_name().c_str(), get_client_starch_version().c_str(),
                 get_client_starch_authors().c_str());
}
