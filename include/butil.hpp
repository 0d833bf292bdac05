// butil.hpp -- BUtil::print as the reference spells it (butil.hpp:6-14).  The reference
// brackets a rank-0 printf with two UPC++ barriers; this process owns all of its GPUs, so
// there is nobody to wait for: flush, print, flush.
#pragma once
#include <cstdio>
#include <string>

namespace BUtil {
template <typename... Args> void print(std::string format, Args... args) {
    fflush(stdout);
    if constexpr (sizeof...(Args) == 0) fputs(format.c_str(), stdout);
    else printf(format.c_str(), args...);
    fflush(stdout);
}
}  // namespace BUtil
