// boost/function.hpp — included by the driver, nothing of it is used on this path.
#pragma once
