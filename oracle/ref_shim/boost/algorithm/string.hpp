// boost/algorithm/string.hpp — trim and iequals, the two calls the pattern readers make.
#pragma once
#include <cctype>
#include <string>

namespace boost {
inline void trim(std::string& s) {
  size_t b = 0, e = s.size();
  while (b < e && std::isspace((unsigned char)s[b])) ++b;
  while (e > b && std::isspace((unsigned char)s[e - 1])) --e;
  s = s.substr(b, e - b);
}
inline bool iequals(const std::string& a, const std::string& b) {
  if (a.size() != b.size()) return false;
  for (size_t i = 0; i < a.size(); ++i)
    if (std::tolower((unsigned char)a[i]) != std::tolower((unsigned char)b[i])) return false;
  return true;
}
}  // namespace boost
