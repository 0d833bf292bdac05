// CPU check of include/kh/stream_reader.hpp with a fake sink (tests/test_stream_reader.py).
//   stream_reader_check <file> <k> <first_line> <n_lines> <chunk_lines> <buffers> <mode>
// mode "copy": chunks are written to stdout back to back; stderr gets "first:lines" per chunk.
// mode "throw<N>": the sink throws on chunk N; mode "rank<P>:<r>": the block of rank r of P (n_lines = total).
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>

#include "kh/stream_reader.hpp"

int main(int argc, char** argv) {
    if (argc != 8) return 2;
    const std::string fname = argv[1], mode = argv[7];
    const int k = std::atoi(argv[2]);
    size_t first = std::strtoull(argv[3], nullptr, 10), n = std::strtoull(argv[4], nullptr, 10);
    kh_stream::Options opt;
    opt.chunk_lines = std::strtoull(argv[5], nullptr, 10);
    opt.buffers = std::atoi(argv[6]);
    if (mode.rfind("rank", 0) == 0) {
        int p = 1, r = 0;
        sscanf(mode.c_str(), "rank%d:%d", &p, &r);
        kh_stream::block_of_rank(n, p, r, first, n);
    }
    int throw_at = -1;
    if (mode.rfind("throw", 0) == 0) throw_at = std::atoi(mode.c_str() + 5);
    int chunk = 0, live_buffers = 0, max_live = 0;
    try {
        const size_t got = kh_stream::for_each_chunk(
            fname, k, first, n, opt,
            [&](size_t bytes) { ++live_buffers; max_live = std::max(max_live, live_buffers); return std::malloc(bytes ? bytes : 1); },
            [&](void* p) { --live_buffers; std::free(p); },
            [&](const char* text, size_t lines, size_t first_of_chunk) {
                if (chunk == throw_at) throw std::runtime_error("sink refused chunk " + std::to_string(chunk));
                fwrite(text, 1, lines * (size_t)(k + 4), stdout);
                fprintf(stderr, "%zu:%zu\n", first_of_chunk, lines);
                if (chunk % 3 == 0 && chunk < 60) std::this_thread::sleep_for(std::chrono::milliseconds(2));   // let the reader run ahead
                ++chunk;
            });
        fprintf(stderr, "delivered %zu buffers %d leaked %d\n", got, max_live, live_buffers);
    } catch (const std::exception& e) {
        fprintf(stderr, "error: %s leaked %d\n", e.what(), live_buffers);
        return 1;
    }
    return 0;
}
