// TEST INFRASTRUCTURE ONLY (oracle build shim)
#pragma once
#include <mutex>
namespace tbb { class mutex { std::mutex m_; public: void lock() { m_.lock(); } void unlock() { m_.unlock(); } }; }
