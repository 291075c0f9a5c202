// boost/thread/lock_types.hpp — stand-in (TEST INFRASTRUCTURE ONLY, see oracle/refshim/README.md):
// boost::unique_lock / boost::recursive_mutex as the std equivalents, for dpose_costmap.cpp:138.
#ifndef DPOSE_REFSHIM_BOOST_LOCK_TYPES
#define DPOSE_REFSHIM_BOOST_LOCK_TYPES
#include <mutex>
namespace boost {
using recursive_mutex = std::recursive_mutex;
template <typename M>
using unique_lock = std::unique_lock<M>;
}  // namespace boost
#endif
