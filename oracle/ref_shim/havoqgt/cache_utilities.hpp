// havoqgt/cache_utilities.hpp — page-cache helpers of the memory-mapped graph image; nothing to do for an in-memory graph.
#pragma once
