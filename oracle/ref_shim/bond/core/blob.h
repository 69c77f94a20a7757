// TEST INFRASTRUCTURE ONLY (oracle build shim): non-owning byte range with the
// handful of bond::blob members the reference uses.
#pragma once
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
namespace bond {
class blob {
    const char* p_; size_t n_;
public:
    blob() : p_(nullptr), n_(0) {}
    blob(const void* p, size_t n) : p_((const char*)p), n_(n) {}
    void assign(const void* p, size_t n) { p_ = (const char*)p; n_ = n; }
    const char* data() const { return p_; }
    const char* content() const { return p_; }
    size_t size() const { return n_; }
    size_t length() const { return n_; }
    bool empty() const { return n_ == 0; }
};
}
