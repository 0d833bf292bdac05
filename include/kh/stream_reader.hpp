// kh/stream_reader.hpp -- streamed ingestion of a k-mer file (the input side).
//
// The reference reads its whole block of the file into one buffer and parses it before the timer starts
// (read_kmers.hpp:54-79, kmer_hash.cpp:122).  At 2-35 GB of text that read IS the wall clock.  This header
// reads the same block of fixed-width lines (K bases, separator, two extensions, '\n' = K+4 bytes) in chunks of
// whole lines on a background thread, into a small ring of caller-provided buffers (pinned host memory for the
// GPU path), and hands every chunk to a sink IN FILE ORDER while the next one is being read:
//
//     kh_stream::for_each_chunk(fname, K, first_line, n_lines, opt, alloc, release,
//                               [&](const char* text, size_t n, size_t first) { kh_insert_lines(t, text, n); });
//
// kh_insert_lines may be called repeatedly and start nodes accumulate in call order (kh_capi.h), so the table and
// the contig order are exactly those of one big insert.  The reader itself knows nothing about CUDA: the sink and
// the allocator are parameters, which is how tests/test_stream_reader.py exercises it on a CPU-only box.
#pragma once

#include <algorithm>
#include <condition_variable>
#include <cstdio>
#include <cstdlib>
#include <exception>
#include <mutex>
#include <stdexcept>
#include <string>
#include <thread>
#include <vector>

namespace kh_stream {

struct Options {
    size_t chunk_lines = size_t(4) << 20;   // lines per chunk (4 Mi lines = 92 MB of 19-mer text)
    int buffers = 3;                        // ring size: one being consumed, the rest being filled
};

// Block of rank `rank` of `nprocs` over `total` lines, as read_kmers.hpp:55-58 cuts it (a rank past the end gets
// an empty block; the reference underflows there).
inline void block_of_rank(size_t total, int nprocs, int rank, size_t& first, size_t& count) {
    const size_t per_rank = (total + nprocs - 1) / nprocs;
    first = std::min(total, per_rank * static_cast<size_t>(rank));
    count = std::min(per_rank, total - first);
}

// Calls sink(text, n_lines_in_chunk, index_of_first_line) for consecutive chunks of lines [first_line,
// first_line + n_lines) of `fname`.  alloc(bytes) / release(ptr) provide the ring buffers.  Returns the number of
// lines delivered (== n_lines).  Throws std::runtime_error if the file cannot be opened or ends early; an exception
// thrown by the sink stops the reader and is rethrown to the caller.
template <class Alloc, class Release, class Sink>
size_t for_each_chunk(const std::string& fname, int k, size_t first_line, size_t n_lines, const Options& opt,
                      Alloc&& alloc, Release&& release, Sink&& sink) {
    if (k < 1) throw std::runtime_error("stream_reader: k must be positive");
    const size_t line_len = static_cast<size_t>(k) + 4;
    const size_t chunk_lines = std::max<size_t>(1, opt.chunk_lines);
    const int nbuf = std::max(2, opt.buffers);
    if (n_lines == 0) return 0;

    FILE* f = fopen(fname.c_str(), "rb");
    if (f == nullptr) throw std::runtime_error("read_kmers: could not open " + fname);
    if (fseeko(f, static_cast<off_t>(line_len * first_line), SEEK_SET) != 0) {
        fclose(f);
        throw std::runtime_error("read_kmers: cannot seek in " + fname);
    }

    struct Slot {
        char* data = nullptr;
        size_t lines = 0, first = 0;
        bool full = false;
    };
    std::vector<Slot> ring(nbuf);
    const size_t buf_bytes = line_len * std::min(chunk_lines, n_lines);
    try {
        for (Slot& s : ring) {
            s.data = static_cast<char*>(alloc(buf_bytes));
            if (s.data == nullptr) throw std::runtime_error("stream_reader: buffer allocation failed");
        }
    } catch (...) {
        for (Slot& s : ring) if (s.data) release(s.data);
        fclose(f);
        throw;
    }

    std::mutex mu;
    std::condition_variable cv;
    bool stop = false;                 // consumer gave up (sink threw)
    std::exception_ptr reader_error;   // reader hit a short read

    std::thread reader([&] {
        size_t done = 0;
        int slot = 0;
        while (done < n_lines) {
            Slot& s = ring[slot];
            {
                std::unique_lock<std::mutex> lk(mu);
                cv.wait(lk, [&] { return !s.full || stop; });
                if (stop) return;
            }
            const size_t want = std::min(chunk_lines, n_lines - done);
            const size_t got = fread(s.data, 1, want * line_len, f);
            {
                std::lock_guard<std::mutex> lk(mu);
                if (got != want * line_len) {
                    reader_error = std::make_exception_ptr(std::runtime_error(
                        "read_kmers: " + fname + " ends inside line " + std::to_string(first_line + done + got / line_len)));
                    s.lines = 0;
                } else {
                    s.lines = want;
                }
                s.first = first_line + done;
                s.full = true;
            }
            cv.notify_all();
            if (got != want * line_len) return;
            done += want;
            slot = (slot + 1) % nbuf;
        }
    });

    size_t delivered = 0;
    std::exception_ptr sink_error;
    int slot = 0;
    while (delivered < n_lines) {
        Slot& s = ring[slot];
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return s.full; });
        }
        if (s.lines == 0) break;       // short read: reader_error is set
        try {
            sink(static_cast<const char*>(s.data), s.lines, s.first);
        } catch (...) {
            sink_error = std::current_exception();
        }
        delivered += s.lines;
        {
            std::lock_guard<std::mutex> lk(mu);
            s.full = false;
            if (sink_error) stop = true;
        }
        cv.notify_all();
        if (sink_error) break;
        slot = (slot + 1) % nbuf;
    }
    reader.join();
    fclose(f);
    for (Slot& s : ring) release(s.data);
    if (sink_error) std::rethrow_exception(sink_error);
    if (reader_error) std::rethrow_exception(reader_error);
    return delivered;
}

// Convenience overload with malloc/free buffers.
template <class Sink>
size_t for_each_chunk(const std::string& fname, int k, size_t first_line, size_t n_lines, const Options& opt, Sink&& sink) {
    return for_each_chunk(fname, k, first_line, n_lines, opt, [](size_t bytes) { return std::malloc(std::max<size_t>(bytes, 1)); },
                          [](void* p) { std::free(p); }, std::forward<Sink>(sink));
}

}  // namespace kh_stream
