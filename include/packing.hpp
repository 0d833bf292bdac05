// packing.hpp -- same include name as the reference's packing.hpp; the definitions live in
// kh/kmer_types.hpp (byte-identical layouts, undefined behaviour removed).
#pragma once
#include "kh/kmer_types.hpp"
