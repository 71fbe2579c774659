// generated stand-in for the CMake-configured header (include/tudocomp/config.h.in): no Judy, no Boost
#pragma once
