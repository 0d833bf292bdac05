// ref_probe: unit-level window onto the reference's own value types -- TEST INFRASTRUCTURE.
//
// Includes /root/reference/kmer_t.hpp (and through it pkmer_t.hpp, packing.hpp) from
// where they lie (-I$(REFERENCE)); nothing is copied.  For each input line in the
// reference's text format it prints, as hex, the bytes of the kmer_pair the
// reference builds (kmer_t.hpp:67-76 -> packing.hpp:77-92) and of next_kmer()
// (kmer_t.hpp:51-53).  tests/golden/make_golden.py stores the output so the
// oracle restatement and the CUDA pack kernel are pinned to the reference at
// record level, not only through the final contigs.
// (the reference headers rely on these being included first, as <upcxx/upcxx.hpp> does)
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <iostream>
#include <string>
#include "kmer_t.hpp"

static void hex(const unsigned char* p, size_t n) {
    for (size_t i = 0; i < n; ++i) printf("%02x", p[i]);
}

int main() {
    std::string line;
    while (std::getline(std::cin, line)) {
        if (line.size() < (size_t)KMER_LEN + 3) continue;
        kmer_pair kp(line.substr(0, KMER_LEN), line.substr(KMER_LEN + 1, 2));
        hex(reinterpret_cast<const unsigned char*>(&kp), sizeof(kp));
        printf(" ");
        if (kp.forwardExt() != 'F') {
            pkmer_t nx = kp.next_kmer();
            hex(nx.data, sizeof(nx.data));
        } else {
            printf("-");
        }
        printf("\n");
    }
    return 0;
}
