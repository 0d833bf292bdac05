// kmer_t.hpp -- same include name as the reference's kmer_t.hpp; the definitions live in
// kh/kmer_types.hpp (byte-identical layouts, undefined behaviour removed).
#pragma once
#include "kh/kmer_types.hpp"
