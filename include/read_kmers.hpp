// read_kmers.hpp -- the reference's input/stringification helpers with the same names and
// behaviour (read_kmers.hpp:14-92): kmer_size, line_count, read_kmers, extract_contig.
// File format: fixed-width lines of KMER_LEN+4 bytes -- K bases, one separator byte that is
// never inspected, backward ext, forward ext, '\n' (read_kmers.hpp:64, 72-76).
#pragma once

#include <algorithm>
#include <cstdio>
#include <fstream>
#include <list>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include "kmer_t.hpp"

// Length of the first whitespace-delimited token of the file (read_kmers.hpp:14-25).
inline int kmer_size(const std::string& fname) {
    std::ifstream in(fname);
    if (!in.is_open()) throw std::runtime_error("kmer_size: could not open " + fname);
    std::string first;
    in >> first;
    return static_cast<int>(first.size());
}

// Number of '\n' bytes in the file (read_kmers.hpp:28-49).
inline size_t line_count(const std::string& fname) {
    FILE* f = fopen(fname.c_str(), "r");
    if (f == nullptr) throw std::runtime_error("line_count: could not open " + fname);
    std::vector<char> buf(1 << 20);
    size_t lines = 0, got;
    while ((got = fread(buf.data(), 1, buf.size(), f)) != 0)
        lines += static_cast<size_t>(std::count(buf.data(), buf.data() + got, '\n'));
    fclose(f);
    return lines;
}

// This rank's block of ceil(n/nprocs) lines, parsed into kmer_pair records (read_kmers.hpp:54-79).
// A rank past the end of the file gets an empty vector (the reference underflows there).
inline std::vector<kmer_pair> read_kmers(const std::string& fname, int nprocs = 1, int rank = 0) {
    const size_t total = line_count(fname);
    const size_t per_rank = (total + nprocs - 1) / nprocs;
    const size_t first = std::min(total, per_rank * static_cast<size_t>(rank));
    const size_t count = std::min(per_rank, total - first);
    const size_t line_len = KMER_LEN + 4;

    FILE* f = fopen(fname.c_str(), "r");
    if (f == nullptr) throw std::runtime_error("read_kmers: could not open " + fname);
    std::vector<char> text(line_len * count);
    fseek(f, static_cast<long>(line_len * first), SEEK_SET);
    const size_t got = fread(text.data(), 1, text.size(), f);
    fclose(f);

    std::vector<kmer_pair> kmers;
    kmers.reserve(count);
    for (size_t off = 0; off + line_len <= got; off += line_len)
        kmers.emplace_back(std::string(&text[off], KMER_LEN), std::string(&text[off + KMER_LEN + 1], 2));
    return kmers;
}

// First k-mer's letters, then every forward extension that is not 'F' (read_kmers.hpp:81-92).
inline std::string extract_contig(const std::list<kmer_pair>& contig) {
    std::string out = contig.front().kmer_str();
    for (const kmer_pair& node : contig)
        if (node.forwardExt() != 'F') out.push_back(node.forwardExt());
    return out;
}
