// TEST / OFFLINE-BUILD INFRASTRUCTURE ONLY.  <sdsl/rank_support.hpp> stand-in: the rank support lives next to the
// bit vector in int_vector.hpp (see there).
#pragma once
#include "int_vector.hpp"
