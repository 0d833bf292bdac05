// kh/kmer_types.hpp -- host value types with the reference's names and byte layouts.
//
// Drop-in for /root/reference packing.hpp, pkmer_t.hpp and kmer_t.hpp: same macros
// (KMER_LEN, PACKED_KMER_LEN), same free functions (packFourMer, packKmer, unpackKmer),
// same structs and member functions (pkmer_t, kmer_pair), and -- the part that matters
// across the C ABI -- identical object bytes:
//     pkmer_t   = unsigned char[(K+3)/4], 2 bits per base, A=0 C=1 G=2 T=3, first base in the
//                 top bits of byte 0, unused tail bits 0 (= 'A' padding)   (packing.hpp:50-92)
//     kmer_pair = pkmer_t followed by char fb_ext[2] = {backward, forward}  (kmer_t.hpp:6-8,43-45)
// Differences are limited to removing undefined behaviour: everything is
// `inline` (header may be included from several translation units), packKmer writes exactly
// PACKED_KMER_LEN bytes even when K % 4 == 0, unpackKmer writes exactly KMER_LEN characters,
// and a letter outside ACGT packs as 'A' instead of reading an uninitialised variable.
#pragma once

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>

#ifndef KMER_LEN
#define KMER_LEN 19
#endif
#define PACKED_KMER_LEN ((KMER_LEN + 3) / 4)

static_assert(KMER_LEN >= 2 && KMER_LEN <= 61, "libkh_b200 supports 2 <= KMER_LEN <= 61");

namespace kh_host {
constexpr char kLetters[4] = {'A', 'C', 'G', 'T'};
constexpr unsigned code_of(char base) {
    return base == 'C' ? 1u : base == 'G' ? 2u : base == 'T' ? 3u : 0u;
}
}  // namespace kh_host

// Four letters -> one byte, first letter in bits 7..6 (packing.hpp:50-75).
inline unsigned char packFourMer(const char* fourMer) {
    unsigned v = 0;
    for (int i = 0; i < 4; ++i) v = (v << 2) | kh_host::code_of(fourMer[i]);
    return static_cast<unsigned char>(v);
}

// K letters -> PACKED_KMER_LEN bytes; a short last group is completed with 'A' (packing.hpp:77-92).
inline void packKmer(const char* kmer, unsigned char* packed_kmer) {
    for (int byte = 0; byte < PACKED_KMER_LEN; ++byte) {
        char group[4] = {'A', 'A', 'A', 'A'};
        for (int j = 0; j < 4 && 4 * byte + j < KMER_LEN; ++j) group[j] = kmer[4 * byte + j];
        packed_kmer[byte] = packFourMer(group);
    }
}

// PACKED_KMER_LEN bytes -> K letters (packing.hpp:94-107, without the lookup table).
inline void unpackKmer(const unsigned char* packed_kmer, char* kmer) {
    for (int i = 0; i < KMER_LEN; ++i)
        kmer[i] = kh_host::kLetters[(packed_kmer[i >> 2] >> (6 - 2 * (i & 3))) & 3];
}

struct pkmer_t {
    unsigned char data[PACKED_KMER_LEN];

    pkmer_t() = default;
    pkmer_t(const pkmer_t&) = default;
    pkmer_t& operator=(const pkmer_t&) = default;
    pkmer_t(const std::string& kmer) { packKmer(kmer.data(), data); }   // pkmer_t.hpp:39

    std::string get() const noexcept {                                  // pkmer_t.hpp:25-29
        std::string s(KMER_LEN, 'A');
        unpackKmer(data, &s[0]);
        return s;
    }
    uint64_t hash() const noexcept {                                    // pkmer_t.hpp:31-37 (djb2 over the bytes)
        uint64_t h = 5381;
        for (unsigned char b : data) h = h * 33 + b;
        return h;
    }
    void init(const unsigned char bytes[PACKED_KMER_LEN]) { std::memcpy(data, bytes, PACKED_KMER_LEN); }
    bool operator==(const pkmer_t& o) const noexcept { return std::memcmp(data, o.data, PACKED_KMER_LEN) == 0; }
    bool operator!=(const pkmer_t& o) const noexcept { return !(*this == o); }
};

struct kmer_pair {
    pkmer_t kmer;
    char fb_ext[2];

    kmer_pair() = default;
    kmer_pair(const kmer_pair&) = default;
    kmer_pair& operator=(const kmer_pair&) = default;
    kmer_pair(const std::string& kmer_s, const std::string& fb) { init(kmer_s, fb); }

    void init(const std::string& kmer_s, const std::string& fb) {       // kmer_t.hpp:67-76
        if (kmer_s.length() != static_cast<size_t>(KMER_LEN) || fb.length() != 2) {
            fprintf(stderr, "error: tried to initialize a kmer pair with too short a string.\n");
            return;
        }
        kmer = pkmer_t(kmer_s);
        fb_ext[0] = fb[0];
        fb_ext[1] = fb[1];
    }
    void init(const kmer_pair& o) { *this = o; }

    std::string kmer_str() const noexcept { return kmer.get(); }
    std::string fb_ext_str() const noexcept { return std::string(fb_ext, 2); }
    char forwardExt() const noexcept { return fb_ext[1]; }              // kmer_t.hpp:43
    char backwardExt() const noexcept { return fb_ext[0]; }             // kmer_t.hpp:45
    pkmer_t next_kmer() const noexcept {                                // kmer_t.hpp:51-53
        std::string s = kmer_str();
        s.erase(0, 1);
        s.push_back(forwardExt());
        return pkmer_t(s);
    }
    pkmer_t last_kmer() const noexcept {                                // kmer_t.hpp:55-57
        std::string s = kmer_str();
        s.pop_back();
        s.insert(s.begin(), backwardExt());
        return pkmer_t(s);
    }
    void print() const noexcept { printf("%s %s\n", kmer_str().c_str(), fb_ext_str().c_str()); }
    uint64_t hash() const noexcept { return kmer.hash(); }
    bool operator==(const kmer_pair& o) const noexcept {
        return kmer == o.kmer && fb_ext[0] == o.fb_ext[0] && fb_ext[1] == o.fb_ext[1];
    }
    bool operator!=(const kmer_pair& o) const noexcept { return !(*this == o); }
};

static_assert(sizeof(pkmer_t) == PACKED_KMER_LEN, "pkmer_t must be exactly the packed bytes");
static_assert(sizeof(kmer_pair) == PACKED_KMER_LEN + 2 && alignof(kmer_pair) == 1,
              "kmer_pair must match the record format of kh_capi.h");
