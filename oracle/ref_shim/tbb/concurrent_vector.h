// TEST INFRASTRUCTURE ONLY (oracle build shim)
#pragma once
#include <deque>
namespace tbb {
template <typename T> class concurrent_vector : public std::deque<T> {
public:
    typename std::deque<T>::iterator grow_by(size_t n) {
        size_t old = this->size(); this->resize(old + n); return this->begin() + old;
    }
};
}
