// TEST INFRASTRUCTURE ONLY (oracle build shim): serial stand-ins.
#pragma once
#include <cstddef>
#include <algorithm>
namespace tbb {
template <typename T> class blocked_range {
    T b_, e_;
public:
    blocked_range(T b, T e) : b_(b), e_(e) {}
    T begin() const { return b_; } T end() const { return e_; }
};
template <typename R, typename F> void parallel_for(const R& r, const F& f) { f(r); }
template <typename It> void parallel_sort(It a, It b) { std::sort(a, b); }
template <typename It, typename F> void parallel_for_each(It a, It b, const F& f) { for (; a != b; ++a) f(*a); }
}
