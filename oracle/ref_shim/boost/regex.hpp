// boost/regex.hpp — std::regex for the file-name filter of the metadata loader.
#pragma once
#include <regex>
#include <string>
namespace boost {
using std::regex;
using std::smatch;
// the loader matches a temporary file name and never reads the match results afterwards
inline bool regex_match(const std::string& s, smatch&, const regex& r) { return std::regex_match(s, r); }
}  // namespace boost
