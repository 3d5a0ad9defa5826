#pragma once
#include <mutex>
#include <shared_mutex>
namespace boost {
using std::mutex;
using std::recursive_mutex;
typedef std::recursive_timed_mutex recursive_try_mutex;
using std::shared_mutex;
using std::defer_lock_t;
using std::try_to_lock_t;
using std::defer_lock;
using std::try_to_lock;
template <typename M> using unique_lock = std::unique_lock<M>;
template <typename M> using shared_lock = std::shared_lock<M>;
template <typename M> using lock_guard = std::lock_guard<M>;
}
