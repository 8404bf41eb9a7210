// TEST INFRASTRUCTURE: forwards <boost/serialization/base_object.hpp> to the stand-in (see _common.hpp).
#pragma once
#include "_common.hpp"
