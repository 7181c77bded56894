// boost/dynamic_bitset.hpp — only dynamic_bitset<>::size_type is named (an unused vertex state class).
#pragma once
#include <cstddef>
#include <vector>
namespace boost {
template <typename Block = unsigned long>
class dynamic_bitset {
 public:
  typedef std::size_t size_type;
  dynamic_bitset() {}
  explicit dynamic_bitset(size_type n) : m_bits(n, false) {}
  size_type size() const { return m_bits.size(); }
 private:
  std::vector<bool> m_bits;
};
}  // namespace boost
