// TEST INFRASTRUCTURE ONLY (oracle build shim)
#pragma once
namespace tbb { class reader_writer_lock { public: void lock() {} void unlock() {} void lock_read() {} }; }
