// TEST INFRASTRUCTURE ONLY (oracle build shim): tbbmalloc entry points mapped to libc.
#pragma once
#include <cstdlib>
#include <cstring>
static inline void* scalable_aligned_malloc(size_t size, size_t align) {
    void* p = nullptr; if (align < 64) align = 64;
    if (posix_memalign(&p, align, size ? size : align)) return nullptr; return p;
}
static inline void scalable_aligned_free(void* p) { free(p); }
static inline void* scalable_malloc(size_t size) { return scalable_aligned_malloc(size, 64); }
static inline void* scalable_calloc(size_t n, size_t sz) {
    void* p = scalable_aligned_malloc(n * sz, 64); if (p) memset(p, 0, n * sz); return p;
}
static inline void scalable_free(void* p) { free(p); }
