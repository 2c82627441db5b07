// TEST STUB of the CMake-generated parpeConfig.h: no MPI, no optional optimizers
#pragma once
#define PARPE_VERSION "reference (headers compiled in place by tests/cpp)"
