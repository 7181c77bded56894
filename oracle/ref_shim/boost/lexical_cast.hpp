// boost/lexical_cast.hpp — string -> number through a stream, which is all the pattern readers ask of it.
#pragma once
#include <sstream>
#include <stdexcept>
#include <string>
namespace boost {
struct bad_lexical_cast : std::runtime_error {
  bad_lexical_cast() : std::runtime_error("bad lexical cast") {}
};
template <typename Target, typename Source>
Target lexical_cast(const Source& s) {
  std::stringstream ss;
  Target t;
  if (!(ss << s) || !(ss >> t) || !(ss >> std::ws).eof()) throw bad_lexical_cast();
  return t;
}
}  // namespace boost
