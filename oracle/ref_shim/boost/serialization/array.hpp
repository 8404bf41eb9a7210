// TEST INFRASTRUCTURE: forwards <boost/serialization/array.hpp> to the stand-in (see _common.hpp).
#pragma once
#include "_common.hpp"
