// boost/interprocess/managed_heap_memory.hpp — the driver only names bip::allocator<T, SegmentManager> in typedefs.
#pragma once
#include <memory>

namespace boost {
namespace interprocess {
template <typename T, typename SegmentManager>
class allocator : public std::allocator<T> {
 public:
  allocator() {}
  template <typename U>
  allocator(const allocator<U, SegmentManager>&) {}
  template <typename U>
  struct rebind { typedef allocator<U, SegmentManager> other; };
};
}  // namespace interprocess
}  // namespace boost
namespace bip = boost::interprocess;
