// TEST INFRASTRUCTURE: forwards <boost/serialization/split_member.hpp> to the stand-in (see _common.hpp).
#pragma once
#include "_common.hpp"
