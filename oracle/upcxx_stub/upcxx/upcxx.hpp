// Single-rank stand-in for <upcxx/upcxx.hpp>  --  TEST INFRASTRUCTURE ONLY.
//
// The reference (kmer_hash.cpp:8, hash_map.hpp:2, butil.hpp:3) includes UPC++,
// which is not installable offline.  This header supplies exactly the symbols
// the reference touches so that /root/reference/kmer_hash.cpp compiles
// UNMODIFIED into a one-rank ("serial") build under oracle/_ref/.  With one
// rank every rpc() target equals rank_me(), so the remote branches only have
// to type-check; they are still executed correctly (locally) if reached.
//
// Symbols covered: init, finalize, rank_me, rank_n,
// barrier, progress + progress_level, future<T>::wait, make_future,
// dist_object<T> (ctor, *, ->), rpc(rank, fn, args...).
#pragma once
// The real upcxx.hpp pulls these in transitively; the reference relies on it
// (pkmer_t.hpp:42 uses memcmp, kmer_hash.cpp:48 runtime_error without including them).
#include <cstdint>
#include <cstring>
#include <functional>
#include <stdexcept>
#include <string>
#include <type_traits>
#include <utility>

namespace upcxx {

inline void init() {}
inline void finalize() {}
inline int rank_me() { return 0; }
inline int rank_n() { return 1; }
inline void barrier() {}

enum class progress_level { internal, user };
inline void progress(progress_level = progress_level::internal) {}

template <class T> class future {
    T held_;
  public:
    future() = default;
    explicit future(T v) : held_(std::move(v)) {}
    T wait() const { return held_; }
    bool ready() const { return true; }
};

template <> class future<void> {
  public:
    void wait() const {}
    bool ready() const { return true; }
};

template <class T> future<std::decay_t<T>> make_future(T&& v) {
    return future<std::decay_t<T>>(std::forward<T>(v));
}
inline future<void> make_future() { return {}; }

template <class T> class dist_object {
    T local_;
  public:
    dist_object(T v) : local_(std::move(v)) {}
    dist_object(const dist_object&) = delete;
    T& operator*() { return local_; }
    const T& operator*() const { return local_; }
    T* operator->() { return &local_; }
    const T* operator->() const { return &local_; }
};

namespace stub_detail {
template <class R> struct flatten { using type = future<R>; };
template <class R> struct flatten<future<R>> { using type = future<R>; };
template <class R> future<R> lift(future<R> f) { return f; }
template <class R> future<std::decay_t<R>> lift(R&& r) { return make_future(std::forward<R>(r)); }
}  // namespace stub_detail

// rpc: run the callable here and now; a callable that itself returns a
// future<R> yields future<R> (UPC++ flattens), anything else future<R>.
template <class F, class... A>
auto rpc(int /*target*/, F&& fn, A&&... a) {
    using R = std::invoke_result_t<F, A&&...>;
    if constexpr (std::is_void_v<R>) {
        std::forward<F>(fn)(std::forward<A>(a)...);
        return future<void>();
    } else {
        return stub_detail::lift(std::forward<F>(fn)(std::forward<A>(a)...));
    }
}

}  // namespace upcxx
