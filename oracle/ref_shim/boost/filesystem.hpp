// boost/filesystem.hpp — std::filesystem has the same interface for what the metadata loader uses.
#pragma once
#include <filesystem>
namespace boost {
namespace filesystem = std::filesystem;
}
